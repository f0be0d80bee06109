"""B200-native (sm_100a) RandLA-Net hot path, drop-in for matthiasverstraete/3d_recognizer's
``randlanet.utils.modules`` / ``randlanet.Model`` on that path.  See DESIGN.md and INTEGRATION.md.

The package name starts with a digit, so import it with
``importlib.import_module("3d_recognizer_b200")`` (or via ``r3d = __import__("3d_recognizer_b200")``).
"""
from .build import build_library  # noqa: F401

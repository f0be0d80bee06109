"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's RandLA-Net hot path (matthiasverstraete/3d_recognizer),
used as the parity checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the
product package ``3d_recognizer_b200`` never does (tests/test_host_logic.py (test_product_never_imports_oracle)
enforces it).

Parity status: the reference ships no tests or golden vectors, so the oracle
is pinned against the reference itself, run in the build container:
  * ``oracle.knn.knn_exact`` vs the reference's nanoflann extension compiled from
    /root/reference (``oracle/_ref/libref_knn.so``, recipe in ``oracle/Makefile``) — squared
    distances bit-identical on every row, indices identical on tie-free rows;
  * ``oracle.network`` vs the reference's own ``randlanet.utils.modules`` imported from
    /root/reference (``oracle/make_golden.py`` writes ``tests/golden/*.npz``).
"""

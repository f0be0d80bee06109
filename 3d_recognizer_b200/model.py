"""Drop-in for the reference's ``randlanet.model.Model`` façade on the hot path (randlanet/model.py).

Same constructor, ``load`` / ``save`` (zip archive holding ``config`` = JSON of the settings and
``model`` = ``torch.save(state_dict)``, model.py:77-121), ``predict`` (consistent pre-sampling, eval
forward, class soft-max, up-sampling back to the full cloud, model.py:146-235) and ``upsample``
(model.py:123-144).  ``train_step`` is the per-batch body of ``Trainer.train``
(randlanet/utils/trainer.py:107-119): H2D copy, forward, loss, backward, optimiser step.  The epoch
loop, data loaders, metrics and TensorBoard logging around it (trainer.py, utils/dataset.py) are host
orchestration outside this package's scope; the reference's ``Trainer`` can drive
``Model.module`` unchanged because it only calls ``model(input)``, ``.parameters()``, ``.train()``,
``.device`` and ``state_dict()``.

There is no CPU device choice: ``use_gpu=False`` raises, the sm_100a library is the only back-end.
"""
import json
import os
import shutil
import tempfile
from collections import OrderedDict
from dataclasses import asdict
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import losses
from .modules import RandLANet, RandLANetSettings, UpSampler
from .preprocessing import sample_points


class Model:
    def __init__(self, settings: RandLANetSettings, weights: Optional[OrderedDict] = None, use_gpu: bool = True,
                 device: Optional[torch.device] = None):
        if not use_gpu:
            raise RuntimeError("3d_recognizer_b200 has no CPU path (use_gpu=False is not supported)")
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("3d_recognizer_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._model = RandLANet(settings, self.device)
        if weights is not None:
            self._model.load_state_dict(weights)
        self._model.eval()
        self._upsampler = UpSampler(settings.upsampling, self.device)

    def __str__(self) -> str:
        return str(self._model)

    @property
    def settings(self) -> RandLANetSettings:
        return self._model.settings

    @property
    def module(self) -> torch.nn.Module:
        return self._model

    # ------------------------------------------------------------------ checkpoint I/O (model.py:77-121)
    @staticmethod
    def load(path: Path, use_gpu: bool = True, **kwargs) -> "Model":
        path = Path(path)
        assert path.is_file(), f"Could not find model file at {path}!"
        with tempfile.TemporaryDirectory() as tmp_str:
            tmp = Path(tmp_str)
            shutil.unpack_archive(str(path), tmp, format="zip")
            with (tmp / "config").open("r") as f:
                config = json.load(f)
            settings = RandLANetSettings(**config)
            state_dict = torch.load(tmp / "model", map_location="cpu")
            if "model" in state_dict.keys():
                state_dict = state_dict["model"]
        for key, value in kwargs.items():
            if hasattr(settings, key):
                setattr(settings, key, value)
        return Model(settings, weights=state_dict, use_gpu=use_gpu, device=kwargs.get("device"))

    def save(self, path: Path) -> None:
        path = Path(path)
        os.makedirs(path.parent, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp_str:
            tmp = Path(tmp_str)
            with (tmp / "config").open("w") as f:
                json.dump(asdict(self.settings), f)
            torch.save(self._model.state_dict(), tmp / "model")
            with tempfile.TemporaryDirectory() as tmp2:
                shutil.make_archive(str(Path(tmp2) / "file"), "zip", tmp)
                shutil.move(str(Path(tmp2) / "file.zip"), path)

    # ------------------------------------------------------------------ inference (model.py:123-235)
    def upsample(self, logits: torch.Tensor, xyz: torch.Tensor, xyz_upsampled: torch.Tensor) -> torch.Tensor:
        """logits (B,C,N1), xyz (B,N1,3), xyz_upsampled (B,N2,3) -> class confidences (B,C,N2)."""
        confidences = torch.softmax(logits, dim=-2).unsqueeze(3)
        return self._upsampler(confidences, xyz, xyz_upsampled).squeeze(-1)

    def predict(self, xyz: np.ndarray, features: Optional[np.ndarray] = None, prepostprocess: bool = True) -> np.ndarray:
        """xyz (B,N,3) or (N,3) [, features (B,N,F) or (N,F)] -> class confidences (B,C,N) or (C,N)."""
        assert xyz.shape[-1] == 3, "xyz should have shape (B) x N x 3!"
        batched = True
        if len(xyz.shape) == 2:
            xyz = np.expand_dims(xyz, 0)
            batched = False
        if features is not None and len(features.shape) == 2:
            features = np.expand_dims(features, 0)
        input = xyz
        if features is not None:
            assert xyz.shape[0] == features.shape[0], "xyz and features should have same batch size!"
            assert xyz.shape[1] == features.shape[1], "xyz and features should have same number of points!"
            input = np.concatenate((xyz, features), axis=-1)
        if self.settings.upsampling == "none":
            prepostprocess = False
        dev = self._model.device
        with torch.no_grad():
            input_t = torch.from_numpy(np.ascontiguousarray(input, dtype=np.float32)).to(dev, non_blocking=True)
            if prepostprocess:
                # host RNG draw identical to the reference (preprocessing.py:35-62, consistent=True)
                indices = sample_points(input.shape[1], self.settings.n_points, consistent=True)
                idx_t = torch.from_numpy(indices).to(dev, non_blocking=True)
                sampled = input_t.index_select(1, idx_t)
                logits = self._model(sampled)
                predictions = self.upsample(logits, sampled[:, :, :3], input_t[:, :, :3]).cpu().numpy()
            else:
                predictions = torch.softmax(self._model(input_t), dim=-2).cpu().numpy()
        if not batched:
            predictions = predictions[0]
        return predictions

    # ------------------------------------------------------------------ training step (trainer.py:107-119)
    def make_optimizer(self, learning_rate: float = 1e-2) -> torch.optim.Optimizer:
        """Adam with the trainer's default learning rate (trainer.py:78-81)."""
        return torch.optim.Adam(self._model.parameters(), lr=learning_rate)

    def train_step(self, input, labels, optimizer: torch.optim.Optimizer, loss_function: str = "dice",
                   flat_grads=None) -> torch.Tensor:
        """One optimisation step on a batch: input (B,N,3+F) fp32, labels (B,N) int64 — host (pinned) or
        device tensors.  Returns the loss as a device scalar (the caller decides when to sync).
        ``flat_grads`` (parallel.FlatGradients) turns on the data-parallel gradient all-reduce between
        backward and the optimiser step."""
        dev = self._model.device
        self._model.train()
        input = input.to(dev, non_blocking=True)
        labels = labels.to(dev, non_blocking=True)
        logits = self._model(input)
        loss = losses.get_loss(loss_function)(logits, labels)
        if flat_grads is not None:
            flat_grads.zero()
        else:
            optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if flat_grads is not None:
            flat_grads.rebind()
            flat_grads.allreduce_mean()
        optimizer.step()
        return loss.detach()

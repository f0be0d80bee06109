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

from . import engine, losses, ops
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
        # Model.predict runs the network on a fixed-size sample of every frame (preprocessing.py:35-62), one small cloud
        # per call: ~60 launches that are bound by host launch overhead, not by the GPU.  The eval forward of that shape is
        # captured once into a CUDA graph and replayed (rebuilt when a parameter or buffer changes).
        self.use_cuda_graphs = True
        self._eval_graphs = {}

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
        input_t = torch.from_numpy(np.ascontiguousarray(input, dtype=np.float32)).to(dev, non_blocking=True)
        predictions = self.predict_device(input_t, prepostprocess).cpu().numpy()
        if not batched:
            predictions = predictions[0]
        return predictions

    def predict_device(self, input_t: torch.Tensor, prepostprocess: bool = True) -> torch.Tensor:
        """The device part of ``predict``: input_t (B,N,3+F) fp32 on the model's device -> class confidences (B,C,N) on
        the device.  Pre-sampling indices come from the host RNG exactly as in the reference (preprocessing.py:35-62,
        consistent=True); soft-max, the K=1 / K=8 neighbour search and the (weighted) gather back to the full cloud
        (model.py:123-144, modules.py:328-456) run on the device."""
        dev = self._model.device
        with torch.no_grad():
            if prepostprocess and self.settings.upsampling != "none":
                indices = sample_points(input_t.shape[1], self.settings.n_points, consistent=True)
                idx_t = torch.from_numpy(indices).to(dev, non_blocking=True)
                sampled = input_t.index_select(1, idx_t)
                logits = self._forward_eval(sampled)
                return self.upsample(logits, sampled[:, :, :3], input_t[:, :, :3])
            return torch.softmax(self._model(input_t), dim=-2)

    def infer(self, x: torch.Tensor) -> torch.Tensor:
        """Class logits (B,C,N) of a device batch x (B,N,3+F) in eval mode, without gradients: ``self.module(x)`` with the
        launches replayed from a CUDA graph of the input shape (small per-GPU batches are bound by host launch overhead:
        4 x 65 536 points take 4.6 ms eager and ~3 ms of kernels).  The permutation draw stays on the host (modules.py:571)."""
        with torch.no_grad():
            self._model.eval()
            return self._forward_eval(x)

    def _forward_eval(self, x: torch.Tensor) -> torch.Tensor:
        """``self._model(x)`` for the pre-sampled cloud of ``predict``: the same host-side permutation draw at the same
        place of the numpy stream (modules.py:571), the kernels replayed from a CUDA graph of this input shape."""
        net = self._model
        if not (self.use_cuda_graphs and x.is_cuda and not net.training and not torch.is_grad_enabled()):
            return net(x)
        key = tuple(x.shape)
        g = self._eval_graphs.get(key)
        if g is None or g.version != engine._state_version(net):
            if len(self._eval_graphs) >= 4:
                self._eval_graphs.clear()
            g = self._eval_graphs[key] = GraphedEvalForward(net, key)
        return g(x, np.random.permutation(x.shape[1]))

    # ------------------------------------------------------------------ training step (trainer.py:107-119)
    def make_optimizer(self, learning_rate: float = 1e-2, capturable: bool = False,
                       flat: bool = True) -> torch.optim.Optimizer:
        """Adam with the trainer's default learning rate (trainer.py:78-81).  ``capturable`` keeps the step
        counter on the device so that the step can live inside a CUDA graph (GraphedTrainStep).  On a CUDA device
        the update runs as torch's fused Adam, by default (``flat``) over ONE flat buffer that every parameter is
        re-pointed into (FlatAdam: one launch; the multi-tensor form takes four 18 us launches for the network's ~100
        small parameters, the foreach form sixteen -- all of them serial at the end of a 2.7 ms step)."""
        on_gpu = self._model.device.type == "cuda"
        capturable = capturable and on_gpu
        # a capturable optimiser keeps the learning rate in a device tensor: the reference's StepLR scheduler
        # (trainer.py:82) then updates it in place and a replayed graph sees the new value
        lr = torch.tensor(float(learning_rate), device=self._model.device) if capturable else learning_rate
        if on_gpu and flat:
            return FlatAdam(self._model, lr, capturable)
        return torch.optim.Adam(self._model.parameters(), lr=lr, capturable=capturable, fused=on_gpu)

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
            if hasattr(optimizer, "bind_flat_gradients"):
                optimizer.bind_flat_gradients(flat_grads.flat)
        optimizer.step()
        return loss.detach()


class FlatAdam(torch.optim.Adam):
    """torch.optim.Adam (fused kernel) over one flat fp32 buffer.

    Every trainable parameter of ``module`` is re-pointed to a view of the buffer (``load_state_dict`` /
    ``state_dict`` keep working: they copy into / read from the views), so the update of the whole network is one
    element-wise launch.  ``step()`` first gathers the gradients backward left on the parameters into a flat
    gradient buffer with one multi-tensor copy (parameters without a gradient — conv biases in front of a
    train-mode BatchNorm — keep a zero slot, which leaves them untouched exactly as skipping them does: no weight
    decay).  With data parallelism the all-reduced buffer of ``parallel.FlatGradients`` (same parameter order) is
    used directly (``bind_flat_gradients``).  The update itself is the library's ``r3d_adam_step`` (torch.optim.Adam's
    arithmetic per element; weight decay / amsgrad / maximize fall back to torch's fused kernel)."""

    def __init__(self, module: torch.nn.Module, lr, capturable: bool):
        self._module_params = [p for p in module.parameters() if p.requires_grad]
        dev = self._module_params[0].device
        n = sum(p.numel() for p in self._module_params)
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        self._flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self._grad_views = []
        off = 0
        with torch.no_grad():
            for p in self._module_params:
                view = flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                self._grad_views.append(self._flat_grad[off:off + p.numel()].view_as(p))
                off += p.numel()
        self._flat = torch.nn.Parameter(flat)
        self._flat.grad = self._flat_grad
        self._bound = None
        super().__init__([self._flat], lr=lr, capturable=capturable, fused=True)

    def bind_flat_gradients(self, flat: torch.Tensor) -> None:
        assert flat.numel() == self._flat.numel() and flat.dtype == torch.float32
        self._bound = flat
        self._flat.grad = flat

    @torch.no_grad()
    def step(self, closure=None):
        if self._bound is None:
            self._flat_grad.zero_()
            dst, src = [], []
            for p, v in zip(self._module_params, self._grad_views):
                if p.grad is not None:
                    dst.append(v)
                    src.append(p.grad)
            if dst:
                torch._foreach_copy_(dst, src)
            self._flat.grad = self._flat_grad
        group = self.param_groups[0]
        if not self._flat.is_cuda or group["amsgrad"] or group["weight_decay"] != 0 or group["maximize"]:
            return super().step(closure)
        # the library's own update kernel (csrc/elementwise.cu): the state keeps torch.optim.Adam's layout (step as a
        # device scalar, exp_avg, exp_avg_sq), so state_dict() / load_state_dict() and lr schedulers work unchanged
        loss = closure() if closure is not None else None
        st = self.state[self._flat]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=self._flat.device)
            st["exp_avg"] = torch.zeros_like(self._flat.data)
            st["exp_avg_sq"] = torch.zeros_like(self._flat.data)
        beta1, beta2 = group["betas"]
        ops.adam_step(self._flat.data, self._flat.grad, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], beta1,
                      beta2, group["eps"])
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self._module_params:
            if set_to_none or p.grad is None:
                p.grad = None
            else:
                p.grad.zero_()


class GraphedEvalForward:
    """The kernel-only eval forward (engine.forward_kernels) of one input shape as a CUDA graph.  The host keeps the one
    thing the reference does on the host — the permutation draw — and hands it over through a pinned staging buffer
    guarded by an event (the host may run ahead of the device).  ``version``: the parameter / buffer version counters the
    captured launches were built from (the folded weights are baked into the graph)."""

    def __init__(self, net: RandLANet, shape, warmup: int = 2):
        B, N, C = shape
        assert C == 3 + net.settings.n_features, "Input should have shape (B, N, 3 + F)!"
        assert N >= net._min_n_points, f"Input point cloud should have at least {net._min_n_points} points!"
        dev = net.device
        self.version = engine._state_version(net)
        self.x = torch.rand((B, N, C), dtype=torch.float32, device=dev)
        self.perm = torch.arange(N, dtype=torch.int64, device=dev)
        self._perm_host = torch.empty(N, dtype=torch.int64).pin_memory()
        self._copied = None
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        with torch.no_grad():
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    engine.forward_kernels(net, self.x, self.perm)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = engine.forward_kernels(net, self.x, self.perm)

    def __call__(self, x: torch.Tensor, permutation: np.ndarray) -> torch.Tensor:
        if self._copied is not None:
            self._copied.synchronize()               # the previous call's H2D copy has left the staging buffer
        self._perm_host.copy_(torch.from_numpy(np.ascontiguousarray(permutation, dtype=np.int64)))
        self.perm.copy_(self._perm_host, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out.clone()                      # the graph's output buffer is overwritten by the next replay


class GraphedTrainStep:
    """The training step of ``Model.train_step`` captured once into CUDA graphs and replayed.

    At the reference's cloud size (2 500 points, train.py:50) a step is a few hundred small launches and is
    bound by host launch overhead, not by the GPU; replaying a graph removes that.  Per step the host only draws
    the point permutation from the global numpy RNG (same call, same position in the stream as
    modules.py:571), copies it and the batch into static device buffers, and replays.  With ``flat_grads``
    (data parallel) forward+backward and the optimiser step are two graphs with the NCCL all-reduce between.
    Shapes are fixed at construction; the optimiser must be ``capturable`` (Model.make_optimizer(capturable=True))."""

    def __init__(self, model: "Model", optimizer: torch.optim.Optimizer, batch_shape, loss_function: str = "dice",
                 flat_grads=None, warmup: int = 3):
        self.model, self.optimizer, self.flat = model, optimizer, flat_grads
        if flat_grads is not None and hasattr(optimizer, "bind_flat_gradients"):
            optimizer.bind_flat_gradients(flat_grads.flat)
        net = model.module
        dev = net.device
        B, N, C = batch_shape
        self.x = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
        self.y = torch.zeros((B, N), dtype=torch.int64, device=dev)
        self.perm = torch.arange(N, dtype=torch.int64, device=dev)
        # The host runs ahead of the GPU (no sync per step): the pinned staging buffer of step i may still be in flight
        # when step i+1 is prepared, so the permutation rotates through a small ring of pinned buffers, each guarded by
        # an event recorded after its H2D copy; a slot is rewritten only once its copy has completed.
        self._perm_ring = [torch.empty(N, dtype=torch.int64).pin_memory() for _ in range(self.RING)]
        self._perm_done = [None] * self.RING
        self._slot = 0
        crit = losses.get_loss(loss_function)
        net.train()

        def fwd_bwd():
            logits = engine.forward_autograd(net, self.x, self.perm)
            loss = crit(logits, self.y)
            loss.backward()
            return loss.detach()

        # warm-up on a side stream (allocator, cuBLAS handles, lazy kernel attributes), as torch.cuda.graphs asks
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.perm.copy_(torch.from_numpy(np.random.permutation(N)))
                self._zero()
                fwd_bwd()
                if self.flat is not None:
                    self.flat.rebind()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        with torch.cuda.graph(self.graph):
            if self.flat is not None:
                self.flat.zero()
            self.loss = fwd_bwd()
            if self.flat is None:
                optimizer.step()
            else:
                self.flat.rebind()          # the gather of the fresh gradients into the flat buffer is part of the graph
        if self.flat is not None:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                optimizer.step()

    def _zero(self):
        if self.flat is not None:
            self.flat.zero()
        else:
            self.optimizer.zero_grad(set_to_none=True)

    RING = 3

    def __call__(self, input: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """input (B,N,3+F) fp32 and labels (B,N) int64, host (pinned) or device.  Returns the loss (a static
        device scalar, overwritten by the next call).  Host ``input`` / ``labels`` are copied asynchronously: the caller
        must leave them untouched until ``self.inputs_consumed`` (a CUDA event recorded after the copies) has
        completed — ``inputs_consumed.synchronize()`` — or simply not reuse a buffer for the next RING steps."""
        slot = self._slot
        self._slot = (slot + 1) % self.RING
        if self._perm_done[slot] is not None:
            self._perm_done[slot].synchronize()                 # this slot's previous copy has left the host buffer
        host = self._perm_ring[slot]
        host.copy_(torch.from_numpy(np.random.permutation(self.perm.shape[0])))
        self.perm.copy_(host, non_blocking=True)
        ev = self._perm_done[slot] or torch.cuda.Event()
        ev.record()
        self._perm_done[slot] = ev
        self.x.copy_(input, non_blocking=True)
        self.y.copy_(labels, non_blocking=True)
        self.inputs_consumed = torch.cuda.Event()
        self.inputs_consumed.record()
        self.graph.replay()
        if self.graph_opt is not None:
            self.flat.allreduce_mean()
            self.graph_opt.replay()
        return self.loss

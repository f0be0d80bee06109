"""Multi-GPU (NCCL) correctness check of the data-parallel path; launched by tests/test_nccl_gpu.py as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/nccl_check.py

(a) inference: whole clouds sharded over the ranks, no collective on the data path -> the gathered logits must equal (fp32 round-off) the
    logits rank 0 computes for the full batch (same weights, same permutation seed on every rank);
(b) training: after one forward/backward per rank on its shard, FlatGradients.allreduce_mean() must leave on every rank
    the mean of the per-rank gradients (gathered separately and averaged on the host side of the check);
(c) one GraphedTrainStep per rank from identical weights keeps the weights identical across ranks.
Rank 0 prints one JSON line."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    modules = importlib.import_module("3d_recognizer_b200.modules")
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    parallel = importlib.import_module("3d_recognizer_b200.parallel")
    syn = importlib.import_module("3d_recognizer_b200.synthetic")
    losses = importlib.import_module("3d_recognizer_b200.losses")
    N, B = 4096, 2 * world
    st = modules.RandLANetSettings(n_classes=2, n_points=N, n_features=0, n_neighbors=16, knn="naive")
    torch.manual_seed(100 + rank)                                  # different initial weights: the broadcast must fix that
    model = model_mod.Model(st, device=dev)
    parallel.broadcast_parameters(model.module)
    x, y = syn.fingertip_batch(7, B, N, n_raw=20000)
    x, y = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    lo, hi = parallel.shard_range(B, rank, world)
    res = {}

    # (a) sharded inference == single-GPU inference
    net = model.module
    net.eval()
    with torch.no_grad():
        np.random.seed(5)
        mine = net(x[lo:hi]).contiguous()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            np.random.seed(5)
            full = net(x)
            # per-cloud results do not depend on the batch they are computed in, up to the kernel variant the
            # dispatcher picks by row count (FP32 tiles vs tcgen05): fp32 round-off, not bit equality
            res["inference_max_rel_diff"] = float((torch.cat(parts, 0) - full).abs().max() / full.abs().max())
            res["inference_equal"] = bool(res["inference_max_rel_diff"] < 1e-5)

    # (b) all-reduced flat gradient == mean of the per-rank gradients
    net.train()
    flat = parallel.FlatGradients(net)
    flat.zero()
    np.random.seed(9)
    loss = losses.get_loss("dice")(net(x[lo:hi]), y[lo:hi])
    loss.backward()
    flat.rebind()
    local_grad = flat.flat.clone()
    flat.allreduce_mean()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean = torch.stack(gathered).double().mean(0)
    err = float((flat.flat.double() - mean).abs().max() / mean.abs().max())
    res["allreduce_mean_rel_err"] = err
    res["ranks_differ_before"] = bool(not torch.equal(gathered[0], gathered[-1]))
    del loss          # a live autograd graph keeps AccumulateGrad nodes bound to this stream: the capture below would trip on them

    # (c) a graphed data-parallel step keeps the replicas' weights identical
    opt = model.make_optimizer(1e-2, capturable=True)
    step = model_mod.GraphedTrainStep(model, opt, (hi - lo, N, 3), "dice", flat)
    for i in range(3):
        step(x[lo:hi], y[lo:hi])
    w = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    res["weights_identical_after_steps"] = bool(all(torch.equal(ws[0], t) for t in ws))
    res["weights_finite"] = bool(torch.isfinite(w).all())
    ok = torch.tensor([1 if (err < 1e-6 and res["weights_identical_after_steps"]) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        res["world"] = world
        res["ok"] = bool(ok.item() == 1 and res.get("inference_equal", False) and res["weights_finite"])
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

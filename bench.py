#!/usr/bin/env python
"""bench.py — the measurement contract of this repo.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One SHORT JSON line (< 1 200 characters) on rank 0; the per-kernel table and everything bulky goes to
profiles/bench_<workload>_n<N>.json (and gpurun_out/).  A "step" is one pass of the RandLA-Net hot path over one batch of synthetic
clouds.  Workloads (BASELINE.json `configs`):

  train40960  (default; configs[3]) training step — forward + dice loss + backward + Adam — on fingertip-style
              clouds of 40 960 points, K=16, 4 encoder levels [16,64,128,256], global batch 64 split over the
              GPUs (strong scaling; NCCL gradient all-reduce at N>1)
  train2500   (configs[1]) the same step on 8 clouds x 2 500 points (train.py:50-56 cloud size) per GPU (weak)
  infer16k | infer64k | infer256k   (configs[2]) eval forward, global batch 32 split over the GPUs (Model.infer: CUDA-graph
                                    replay per input shape; the instrumented pass launches eagerly)
  train40960_b8 | train40960_b16 | infer64k_b4   one rank's shard of the 8- / 4-GPU runs on ONE GPU (diagnostics)
  knn1m_k16 | knn1m_k32             (configs[4]) 1 M x 1 M exact KNN micro-benchmark (1 GPU)
  predict160k                       (configs[0]'s call, on the GPU) Model.predict of one 160 998-point LiDAR-shaped frame
                                    with the repo-default model: the UI's 250 ms budget (main.py:49)

`value`  : points/sec (queries/sec for knn*) with the batch already resident in HBM.
`e2e`    : the same through the public API (Model.train_step / Model.predict / ops.knn_host) with HOST
           buffers: pinned host -> device copy of the batch and a device -> host read of the result
           inside the timed region, every step.
`--impl reference` times the reference's own CPU implementation of the same step on the host cores
(the oracle port of randlanet/utils/modules.py driven by the exact KNN; nanoflann from oracle/_ref for
the KNN micro-benchmark when it was built) on a bounded sample of the workload.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n_points, K, global_batch or None (=> per-GPU batch below), per_gpu_batch)
    "train2500": dict(kind="train", n=2500, k=16, per_gpu_batch=8, scaling="weak"),
    "train40960": dict(kind="train", n=40960, k=16, global_batch=64, scaling="strong"),
    # one rank's shard of train40960 on 8 / 4 GPUs, run on ONE GPU: the diagnostic for the strong-scaling tail
    "train40960_b8": dict(kind="train", n=40960, k=16, global_batch=8, scaling="strong"),
    "train40960_b16": dict(kind="train", n=40960, k=16, global_batch=16, scaling="strong"),
    "infer64k_b4": dict(kind="infer", n=65536, k=16, global_batch=4, scaling="strong"),
    "infer16k": dict(kind="infer", n=16384, k=16, global_batch=32, scaling="strong"),
    "infer64k": dict(kind="infer", n=65536, k=16, global_batch=32, scaling="strong"),
    "infer256k": dict(kind="infer", n=262144, k=16, global_batch=32, scaling="strong"),
    # the UI's prediction call (main.py:49: one Model.predict every 250 ms on the Tk thread): one 160 998-point frame
    # (the mock L515 frames' size), repo-default model (train.py:50-51: 2 500 points, K = 32), host array in and out
    "predict160k": dict(kind="predict", n=160998, k=32, per_gpu_batch=1, scaling="weak"),
    "knn1m_k16": dict(kind="knn", n=1 << 20, k=16, per_gpu_batch=1, scaling="weak"),
    "knn1m_k32": dict(kind="knn", n=1 << 20, k=32, per_gpu_batch=1, scaling="weak"),
}
DEFAULT_WORKLOAD = "train40960"
SETTINGS = dict(n_classes=2, n_features=0, decimation=4, layer_sizes=[16, 64, 128, 256], knn="naive",
                upsampling="nni")
L2_FLUSH_BYTES = 256 << 20        # > the 126 MB L2


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ helpers
def dist_setup(n_gpus):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, world) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def timed_steps(step_fn, steps, warmup, world, flush):
    """W untimed warm-up steps, then exactly K steps, each bracketed by CUDA events on the current
    stream with an L2 flush in between (outside the events); barrier + synchronize on both sides.
    Returns (sum of the K step times in ms — max over ranks, wall seconds of the whole region)."""
    for i in range(warmup):
        step_fn(i)
    barrier(world)
    evs = []
    t0 = time.perf_counter()
    for i in range(steps):
        flush.add_(1.0)                       # writes 256 MB: evicts the previous step from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_fn(warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier(world)
    wall = time.perf_counter() - t0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    return max_over_ranks(total_ms, world), wall


def measure_fp32_peak(cabi):
    """Live FP32-pipe peak (TFLOP/s) from the library's probe kernel: best of 3, scalar FFMA and FFMA2."""
    import ctypes
    L = cabi.lib()
    out = torch.empty(L.r3d_fp32_probe_floats(), dtype=torch.float32, device="cuda")
    res = {}
    for mode, name in ((0, "ffma"), (1, "ffma2")):
        flops = ctypes.c_double(0)
        best = None
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cabi.check(L.r3d_fp32_probe(mode, 4000, cabi.ptr(out), ctypes.byref(flops), cabi.stream_ptr(out.device)),
                       "r3d_fp32_probe")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if it > 0:
                best = ms if best is None else min(best, ms)
        res[name] = flops.value / best * 1e-9
    return res


def event_pair_overhead_ms(n=64):
    """What a CUDA-event pair measures around nothing at all, with the GPU busy before and after (median of n): the
    fixed cost every per-kernel sample below carries, subtracted in kernel_table."""
    torch.cuda._sleep(int(2e6))
    pairs = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in pairs)[n // 2]


def kernel_table(timers, overhead_ms=0.0):
    """name -> dict(launches, ms_total, ms_avg, flops, bytes) from the C-ABI wrappers' event pairs."""
    tab = {}
    for name, lst in timers.items():
        ms = [max(a.elapsed_time(b) - overhead_ms, 1e-4) for a, b, _ in lst]
        tab[name] = dict(launches=len(lst), ms_total=sum(ms), ms_avg=sum(ms) / len(ms),
                         flops=sum(w.get("flops", 0.0) for _, _, w in lst) / len(lst),
                         bytes=sum(w.get("bytes", 0.0) for _, _, w in lst) / len(lst))
    return tab


def kernel_family(name):
    """Timer name -> CUDA kernel behind it: shapes dropped, the two halves of an LFA block share one kernel template
    (lfa_pool_kernel<D,K,STAGE> / lfa_pool_bwd_kernel<D,K,NT,STAGE>)."""
    base = name.split("[")[0]
    for a, b in (("lfa_pool1_bwd", "lfa_pool_bwd"), ("lfa_pool2_bwd", "lfa_pool_bwd"), ("lfa_pool1", "lfa_pool"),
                 ("lfa_pool2", "lfa_pool"), ("lfa_cl_fwd1", "lfa_cl_fwd"), ("lfa_cl_fwd2", "lfa_cl_fwd"),
                 ("lfa_cl_bwd1", "lfa_cl_bwd"), ("lfa_cl_bwd2", "lfa_cl_bwd")):
        if base == a:
            return b
    return base


def roofline_of(tab, peaks, fp32_peak):
    """Roofline entry of the dominant kernel: the kernel (all its launches of the step, whatever their shapes) with
    the largest share of the timed region.  achieved = algorithmic work of those launches / their summed duration,
    i.e. the launch-time-weighted mean over the shapes the step runs the kernel on."""
    if not tab:
        return None
    fam = {}
    for name, k in tab.items():
        f = fam.setdefault(kernel_family(name), dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, shapes=[]))
        n = k["launches"]
        f["ms"] += k["ms_avg"] * n
        f["flops"] += k["flops"] * n
        f["bytes"] += k["bytes"] * n
        f["launches"] += n
        f["shapes"].append(name)
    name = max(fam, key=lambda n: fam[n]["ms"])
    f = fam[name]
    t_s = f["ms"] * 1e-3
    n = max(f["launches"], 1)
    t_hbm = f["bytes"] / (peaks["hbm_gbs"] * 1e9)
    t_fp32 = f["flops"] / (fp32_peak["ffma"] * 1e12)
    if name.startswith("lfa_cl_"):
        # tensor-core (tcgen05, split-fp16) kernels: the contraction runs on the tensor pipe.  achieved = ALGORITHMIC
        # fp32-equivalent flops / time; each algorithmic product costs three fp16 MMAs, so 1/3 of the peak is the ceiling
        tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        ach = f["flops"] / t_s * 1e-12
        return dict(kernel=name, launches=f["launches"], ms_avg=f["ms"] / n, traffic=None, shapes=sorted(f["shapes"]),
                    algorithmic_flops_per_launch=f["flops"] / n, algorithmic_bytes_per_launch=f["bytes"] / n,
                    bound="tensor", achieved=ach, peak=tpeak, unit="TFLOP/s", frac=ach / tpeak,
                    vs_fp32_pipe=ach / fp32_peak["ffma"],       # > 1: more than the FP32 pipe could deliver at its peak
                    peak_source=peaks["source"] + ": dense bf16/fp16 tensor peak, sustained; fp32-accurate split-fp16 "
                                "issues 3 MMAs per algorithmic product (attainable <= 1/3 of this peak)")
    common = dict(kernel=name, launches=f["launches"], ms_avg=f["ms"] / n, traffic=None, shapes=sorted(f["shapes"]),
                  algorithmic_flops_per_launch=f["flops"] / n, algorithmic_bytes_per_launch=f["bytes"] / n)
    if t_fp32 >= t_hbm:
        ach = f["flops"] / t_s * 1e-12
        return dict(common, bound="fp32", achieved=ach, peak=fp32_peak["ffma"], unit="TFLOP/s",
                    frac=ach / fp32_peak["ffma"],
                    peak_source="FFMA probe kernel measured in this run (r3d_fp32_probe); FFMA2 packed: %.1f"
                                % fp32_peak["ffma2"])
    ach = f["bytes"] / t_s * 1e-9
    return dict(common, bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"],
                peak_source=peaks["source"])


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_train_or_infer(kind, n, k, batch, budget_s, max_steps, seed=0):
    """The reference's CPU implementation of one step (oracle port of modules.py + exact KNN, torch
    intra-op threads = all host cores).  Returns (points/sec, steps run, batch used, cores)."""
    from oracle import network as onet
    syn = importlib.import_module("3d_recognizer_b200.synthetic")
    losses = importlib.import_module("3d_recognizer_b200.losses")
    if kind == "predict":
        # Model.predict of the reference on the CPU (oracle port: consistent pre-sampling to 2 500 points, eval forward,
        # soft-max, 1-NN up-sampling back to the full frame), one frame per step
        st = dict(SETTINGS, n_points=2500, n_neighbors=k)
        sd = onet.synth_state_dict(st, seed)
        frame = syn.fingertip_cloud(np.random.RandomState(seed), n)[0]
        times = []
        t_start = time.perf_counter()
        for i in range(max_steps + 1):
            t0 = time.perf_counter()
            onet.predict(sd, st, frame)
            if i > 0:
                times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget_s and times:
                break
        return n / statistics.mean(times), len(times), 1, torch.get_num_threads()
    st = dict(SETTINGS, n_points=n, n_neighbors=k)
    sd = onet.synth_state_dict(st, seed)
    params = []
    if kind == "train":
        for name, t in sd.items():
            if t.is_floating_point() and "running" not in name:
                t.requires_grad_(True)
                params.append(t)
        opt = torch.optim.Adam(params, lr=1e-2)
    x, lab = syn.fingertip_batch(seed, batch, n, n_raw=max(150_000, 2 * n))
    x, lab = torch.from_numpy(x), torch.from_numpy(lab)
    crit = losses.get_loss("dice")
    times = []
    t_start = time.perf_counter()
    for i in range(max_steps + 1):                    # first step is warm-up
        t0 = time.perf_counter()
        if kind == "train":
            logits = onet.forward(sd, st, x, training=True)
            loss = crit(logits, lab)
            opt.zero_grad()
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                onet.forward(sd, st, x, training=False)
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    return batch * n / statistics.mean(times), len(times), batch, torch.get_num_threads()


def cpu_knn(n, k, budget_s):
    """KNN queries/sec on the host: the reference's nanoflann extension (oracle/_ref, single threaded by
    construction, knn.cpp:53) when it was built, else the oracle brute force on all cores; on a bounded
    sample of the queries against the full 1 M support."""
    from oracle import knn as oknn
    syn = importlib.import_module("3d_recognizer_b200.synthetic")
    s = syn.uniform_clouds(0, 1, n)
    nq = 100_000 if oknn.have_ref() else 2_000
    q = syn.uniform_clouds(1, 1, nq)
    t0 = time.perf_counter()
    if oknn.have_ref():
        oknn.ref_knn_tpk(s, q, k)
        kind, cores = "reference", 1
    else:
        oknn.knn_exact(s, q, k)
        kind, cores = "port", os.cpu_count()
    dt = time.perf_counter() - t0
    return nq / dt, kind, cores, f"{nq} queries against the full {n}-point support (tree build included)"


def run_reference(args, wl, name):
    """--impl reference: rank 0 only, host cores only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.set_num_threads(os.cpu_count())
    if wl["kind"] == "knn":
        v, kind, cores, sample = cpu_knn(wl["n"], wl["k"], 30)
        unit, metric = "queries/s", "knn_queries_per_sec"
        ms = 1e3 * wl["n"] / v
        cfg = dict(workload=name, n_points=wl["n"], k=wl["k"])
    else:
        gb = wl.get("global_batch") or wl["per_gpu_batch"] * world
        b = min(gb, 8 if wl["n"] <= 4096 else 1)
        v, nsteps, b, cores = cpu_train_or_infer(wl["kind"], wl["n"], wl["k"], b, 25 * max(1, args.steps) / 5,
                                                 max(1, args.steps))
        kind = "port"
        sample = f"{nsteps} step(s) of batch {b} x {wl['n']} points after 1 warm-up"
        unit = "points/s"
        metric = {"train": "train_step_points_per_sec", "predict": "predict_points_per_sec"}.get(
            wl["kind"], "forward_points_per_sec")
        ms = 1e3 * b * wl["n"] / v
        cfg = dict(workload=name, n_points=wl["n"], k=wl["k"], global_batch=gb, layer_sizes=SETTINGS["layer_sizes"])
    line = dict(impl="reference", metric=metric, value=v, unit=unit, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling=wl["scaling"], vs_baseline=None,
                dtype="f32", data="synthetic", config=cfg,
                cpu_baseline=dict(value=v, unit=unit, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=v, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------ output
MAX_LINE_CHARS = 1200          # the driver keeps a short tail of stdout: the bench line must fit in it


def _r(x, nd=4):
    """Round floats (significant digits for small values) so that the line stays short."""
    if isinstance(x, float):
        return float(f"{x:.{nd + 2}g}")
    return x


def compose_line(*, metric, value, unit, world, args, ms_per_step, wl, name, n, k, gbatch, batch, e2e_value, h2d,
                 d2h, launches, clocks, roof, cpu_base, tab, fp32_peak, wall, graphed, eager_ms_per_step, extras):
    """(the ONE short JSON line printed on stdout, the full record for the side file).  Everything bulky — the
    per-kernel table, the roofline's shape list and prose, the extras — goes to the side file only."""
    roof_short = None
    if roof:
        roof_short = {a: _r(roof[a]) for a in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic")}
        roof_short["launches"] = roof["launches"]
        if "vs_fp32_pipe" in roof:
            roof_short["vs_fp32_pipe"] = _r(roof["vs_fp32_pipe"])
    cpu_short = None
    if cpu_base:
        cpu_short = dict(value=_r(cpu_base["value"]), unit=cpu_base["unit"], cores=cpu_base["cores"],
                         kind=cpu_base["kind"], sample=cpu_base["sample"][:80])
    clocks_short = dict(sm_mhz=clocks.get("sm_mhz"), sm_max_mhz=clocks.get("sm_max_mhz"),
                        reasons=clocks.get("reasons", []))
    par = f"dp{world}" + ("+nccl_allreduce" if wl["kind"] == "train" and world > 1 else "")
    short = dict(metric=metric, value=_r(value), unit=unit, n_gpus=world, steps=args.steps, warmup=args.warmup,
                 ms_per_step=_r(ms_per_step), higher_is_better=True, scaling=wl["scaling"], vs_baseline=None,
                 dtype="f32", data="synthetic",
                 config=dict(workload=name, n_points=n, k=k, global_batch=gbatch, per_gpu_batch=batch,
                             l2="flushed between steps", parallelism=par),
                 e2e=dict(value=_r(e2e_value), unit=unit, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                 gpu_launches=int(launches), clocks=clocks_short, roofline=roof_short, cpu_baseline=cpu_short,
                 execution=("cuda-graph replay" if graphed else "eager"))
    line = json.dumps(short, separators=(",", ":"))
    if len(line) > MAX_LINE_CHARS:             # never let prose push the line out of the driver's tail
        if cpu_short:
            cpu_short.pop("sample", None)
        short["config"].pop("l2", None)
        line = json.dumps(short, separators=(",", ":"))
    assert len(line) <= MAX_LINE_CHARS, len(line)
    side = dict(short, config=dict(short["config"], layer_sizes=SETTINGS["layer_sizes"],
                                   l2="flushed between steps (256 MB write)"),
                roofline=roof, cpu_baseline=cpu_base, clocks=clocks,
                kernels={kn: {a: (round(b, 6) if isinstance(b, float) else b) for a, b in kv.items()}
                         for kn, kv in tab.items()},
                fp32_peak_tflops=fp32_peak, wall_s_timed_region=wall, eager_ms_per_step=eager_ms_per_step,
                extras=extras)
    return line, side


def write_side_file(side, name, world):
    """Full record (kernel table, roofline shapes, extras) next to the short line: profiles/ (tracked) and
    gpurun_out/ (what comes back from a GPU box)."""
    for d in ("profiles", "gpurun_out"):
        try:
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, f"bench_{name}_n{world}.json"), "w") as fh:
                json.dump(side, fh, indent=1)
        except OSError:
            pass


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="training step eager instead of CUDA-graph replay")
    ap.add_argument("--profile-steps", type=int, default=0,
                    help="run only P eager steps between cudaProfilerStart/Stop (for `ncu --profile-from-start off`); "
                         "prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    name = args.workload
    wl = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference(args, wl, name)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); use --impl reference for the CPU arm")
    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    cabi = importlib.import_module("3d_recognizer_b200._cabi")
    ops = importlib.import_module("3d_recognizer_b200.ops")
    modules = importlib.import_module("3d_recognizer_b200.modules")
    model_mod = importlib.import_module("3d_recognizer_b200.model")
    parallel = importlib.import_module("3d_recognizer_b200.parallel")
    syn = importlib.import_module("3d_recognizer_b200.synthetic")
    L = cabi.lib()
    peaks = load_peaks()
    flush = torch.zeros(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)

    n, k = wl["n"], wl["k"]
    if "global_batch" in wl:
        lo, hi = parallel.shard_range(wl["global_batch"], rank, world)
        batch, gbatch = hi - lo, wl["global_batch"]
    else:
        batch, gbatch = wl["per_gpu_batch"], wl["per_gpu_batch"] * world
    fp32_peak = measure_fp32_peak(cabi)
    extras = {}
    eager_step = build_graph = None

    if wl["kind"] == "knn":
        metric, unit = "knn_queries_per_sec", "queries/s"
        s_h = torch.from_numpy(syn.uniform_clouds(2 * rank, batch, n)).pin_memory()
        q_h = torch.from_numpy(syn.uniform_clouds(2 * rank + 1, batch, n)).pin_memory()
        s_d, q_d = s_h.to(dev), q_h.to(dev)
        units_per_step = batch * n

        def dev_step(i):
            ops.knn(s_d, q_d, k, idx64=True, dist=False, dist_sq=True)

        def e2e_step(i):
            ops.knn_host(s_h.numpy(), q_h.numpy(), k)

        h2d, d2h = 2 * batch * n * 12, batch * n * k * 12
    else:
        if wl["kind"] != "predict":
            st = modules.RandLANetSettings(**dict(SETTINGS, n_points=n, n_neighbors=k))
            torch.manual_seed(0)
            model = model_mod.Model(st, device=dev)
            parallel.broadcast_parameters(model.module)
        pool = 4
        if wl["kind"] != "predict":
            data = [syn.fingertip_batch(1000 * rank + j, batch, n, n_raw=max(150_000, 2 * n)) for j in range(pool)]
            x_h = [torch.from_numpy(x).pin_memory() for x, _ in data]
            y_h = [torch.from_numpy(y).pin_memory() for _, y in data]
            x_d = [x.to(dev) for x in x_h]
            y_d = [y.to(dev) for y in y_h]
        units_per_step = batch * n
        unit = "points/s"
        if wl["kind"] == "predict":
            metric = "predict_points_per_sec"
            st = modules.RandLANetSettings(**dict(SETTINGS, n_points=2500, n_neighbors=k))
            torch.manual_seed(0)
            model = model_mod.Model(st, device=dev)
            frames = [syn.fingertip_cloud(np.random.RandomState(100 * rank + j), n)[0] for j in range(pool)]
            x_d = [torch.from_numpy(f[None]).to(dev) for f in frames]

            def dev_step(i):
                model.predict_device(x_d[i % pool])

            def eager_step(i):                           # instrumented pass: the network's launches one by one
                model.use_cuda_graphs = False
                try:
                    model.predict_device(x_d[i % pool])
                finally:
                    model.use_cuda_graphs = True

            def build_graph():                           # Model.predict captures the network's eval forward itself
                model.predict_device(x_d[0])

            if args.no_graph:
                model.use_cuda_graphs, build_graph = False, None

            def e2e_step(i):
                model.predict(frames[i % pool])          # numpy in, numpy out: H2D, forward, up-sampling, D2H

            h2d, d2h = n * 12, n * 2 * 4
        elif wl["kind"] == "train":
            metric = "train_step_points_per_sec"
            opt = model.make_optimizer(1e-2, capturable=not args.no_graph)
            flat = parallel.FlatGradients(model.module) if world > 1 else None
            sink = torch.zeros((), device=dev)

            def eager_step(i):
                sink.copy_(model.train_step(x_d[i % pool], y_d[i % pool], opt, "dice", flat))

            dev_step = eager_step
            if args.no_graph:
                def e2e_step(i):
                    loss = model.train_step(x_h[i % pool], y_h[i % pool], opt, "dice", flat)
                    loss.item()                                        # D2H read of the step's result
            else:
                gstep = None

                def build_graph():
                    nonlocal gstep
                    gstep = model_mod.GraphedTrainStep(model, opt, (batch, n, 3), "dice", flat)

                def dev_step(i):                                       # noqa: F811 (graph replay)
                    sink.copy_(gstep(x_d[i % pool], y_d[i % pool]))

                def e2e_step(i):
                    gstep(x_h[i % pool], y_h[i % pool]).item()         # H2D of the batch + D2H of the loss

            h2d, d2h = batch * n * (12 + 8), 4
        else:
            metric = "forward_points_per_sec"
            model.module.eval()
            res_h = torch.empty((batch, 2, n), dtype=torch.float32).pin_memory()

            def eager_step(i):                           # instrumented pass: the launches one by one
                with torch.no_grad():
                    model.module(x_d[i % pool])

            def dev_step(i):                             # Model.infer: CUDA-graph replay per input shape
                model.infer(x_d[i % pool])

            def build_graph():
                model.infer(x_d[0])

            def e2e_step(i):
                logits = model.infer(x_h[i % pool].to(dev, non_blocking=True))
                res_h.copy_(logits, non_blocking=True)
                torch.cuda.synchronize()

            if args.no_graph:
                model.use_cuda_graphs, build_graph = False, None

            h2d, d2h = batch * n * 12, batch * n * 2 * 4

    if args.profile_steps:
        step = eager_step or dev_step
        for i in range(args.warmup):
            step(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.profile_steps):
            step(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_steps": args.profile_steps, "workload": name}))
        return

    # ---- eager instrumented pass: per-kernel CUDA events (C-ABI wrappers) and launch counting
    eager = eager_step or dev_step
    for i in range(args.warmup):
        eager(i)
    torch.cuda.synchronize()
    n_eager = min(args.steps, 5)
    launches0 = L.r3d_launch_count()
    eager_ms, _ = timed_steps(eager, n_eager, 0, world, flush)
    launches_per_step = (L.r3d_launch_count() - launches0) // n_eager
    eager_ms_per_step = eager_ms / n_eager
    # Per-kernel events.  A small step is host-bound in eager mode: an event pair around one launch would time the
    # host's launch gap, not the kernel.  So the GPU is given a head start of one eager step (a clock spin of
    # torch.cuda._sleep) while the host enqueues the launches; they then run back to back and the events bracket
    # kernel time only.
    head_start = int(min(eager_ms_per_step * 1.2, 60.0) * 1.9e6)

    def instrumented(i):
        torch.cuda._sleep(head_start)
        eager(i)

    # per-kernel events must bracket ONE kernel: the side-stream forks of the step run inline for this pass only
    engine = importlib.import_module("3d_recognizer_b200.engine")
    engine.SERIALIZE_FORKS = True
    cabi.KERNEL_TIMERS = {}
    timed_steps(instrumented, n_eager, 0, world, flush)
    engine.SERIALIZE_FORKS = False
    timer_overhead_ms = event_pair_overhead_ms()
    tab = kernel_table(cabi.KERNEL_TIMERS, timer_overhead_ms)
    cabi.KERNEL_TIMERS = None
    extras["kernel_timer_overhead_us"] = round(timer_overhead_ms * 1e3, 2)
    if wl["kind"] == "knn":
        # The default search for clouds this large is the uniform-grid back-end, whose work is O(N K), not the
        # 8 N^2 flop of the exhaustive scan, so it has no FP32 roofline to speak of.  The roofline entry is taken
        # on the tiled brute-force kernel (same results, forced through r3d_knn_set_algorithm), timed here.
        prev = L.r3d_knn_set_algorithm(1)
        cabi.KERNEL_TIMERS = {}
        for i in range(3):
            ops.knn(s_d, q_d, k, idx64=True, dist=False, dist_sq=True)
        torch.cuda.synchronize()
        brute = kernel_table(cabi.KERNEL_TIMERS)
        cabi.KERNEL_TIMERS = None
        L.r3d_knn_set_algorithm(prev)
        for kn, kv in brute.items():
            kv["ms_total"] = kv["ms_avg"] * n_eager * 1.0      # comparable with the per-step table below
            kv["launches"] = n_eager                           # one launch per step
            tab["forced_" + kn] = kv
        grid_name = next(kn for kn in tab if kn.startswith("knn_grid"))
        extras["default_search"] = "uniform grid (r3d_knn algorithm 0/2)"
        extras["brute_force_queries_per_sec"] = units_per_step / (next(iter(brute.values()))["ms_avg"] * 1e-3)
        extras["grid_ms_per_launch"] = tab[grid_name]["ms_avg"]
    graphed = build_graph is not None
    if graphed:
        build_graph()

    # ---- device-resident timed region (value)
    with ClockSampler(local) as clk:
        total_ms, wall = timed_steps(dev_step, args.steps, args.warmup, world, flush)
    launches = launches_per_step * args.steps          # a graph replay re-issues the launches of one eager step
    ms_per_step = total_ms / args.steps
    value = sum_over_ranks(units_per_step, world) / (ms_per_step * 1e-3)
    clocks = clk.summary()

    # ---- end-to-end through the public API with host buffers
    e2e_ms, _ = timed_steps(e2e_step, args.steps, 2, world, flush)
    e2e_value = sum_over_ranks(units_per_step, world) / (e2e_ms / args.steps * 1e-3)

    # per-kernel shares are relative to the EAGER instrumented step (event records are not capturable)
    for kname in tab:
        kv = tab[kname]
        # per-kernel roofline fractions (algorithmic work / measured launch time / peak of the binding resource)
        kv["tflops"] = kv["flops"] / kv["ms_avg"] * 1e-9 if kv["ms_avg"] > 0 else 0.0
        kv["frac_fp32_peak"] = kv["tflops"] / fp32_peak["ffma"]
        kv["gbs"] = kv["bytes"] / kv["ms_avg"] * 1e-6 if kv["ms_avg"] > 0 else 0.0
        kv["frac_hbm_peak"] = kv["gbs"] / peaks["hbm_gbs"]
        tab[kname]["ms_per_step"] = tab[kname]["ms_total"] / n_eager
        tab[kname]["share_of_eager_step"] = tab[kname]["ms_per_step"] / eager_ms_per_step
        tab[kname]["launches"] //= n_eager
    roof = roofline_of(tab, peaks, fp32_peak)
    # DRAM traffic of the dominant kernel comes from a committed ncu --set full capture (never measured under this run)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as fh:
            ent = json.load(fh).get(name, {}).get(roof["kernel"]) if roof else None
        if ent:
            roof["traffic"] = ent["bytes_per_launch"]
            roof["traffic_source"] = ent["source"]
    except (OSError, ValueError):
        pass

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        if wl["kind"] == "knn":
            v, kind, cores, sample = cpu_knn(n, k, 30)
            cpu_base = dict(value=v, unit=unit, cores=cores, kind=kind, sample=sample)
        else:
            b = min(batch, 8 if n <= 4096 else 1)
            v, nsteps, b, cores = cpu_train_or_infer(wl["kind"], n, k, b, 15, 8)
            cpu_base = dict(value=v, unit=unit, cores=cores, kind="port",
                            sample=f"{nsteps} step(s) of batch {b} x {n} points after 1 warm-up, oracle port of "
                                   "randlanet/utils/modules.py + exact KNN, torch threads = all host cores")

    if rank == 0:
        line, side = compose_line(
            metric=metric, value=value, unit=unit, world=world, args=args, ms_per_step=ms_per_step, wl=wl, name=name,
            n=n, k=k, gbatch=gbatch, batch=batch, e2e_value=e2e_value, h2d=h2d, d2h=d2h, launches=launches,
            clocks=clocks, roof=roof, cpu_base=cpu_base, tab=tab, fp32_peak=fp32_peak, wall=wall, graphed=graphed,
            eager_ms_per_step=eager_ms_per_step, extras=extras)
        write_side_file(side, name, world)
        print(line, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

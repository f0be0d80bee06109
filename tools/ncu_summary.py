"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "dram__cycles_active.avg.pct",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "smsp__sass_thread_inst_executed_op_fadd", "smsp__sass_thread_inst_executed_op_ffma",
        "smsp__sass_thread_inst_executed_op_fmul", "sm__sass_inst_executed_op_shared"]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"== kernel {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
    for h, u in zip(hdr, units):
        if any(w in h for w in WANT) and "realtime" not in h and ".max" not in h and ".min" not in h and ".sum." not in h:
            print(f"  {h} [{u}] = {d[h]}")

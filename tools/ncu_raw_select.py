"""Select the roofline counters from an `ncu --set full --csv --page raw` log (one row per launch) and print them per
launch; with --traffic, also print the mean DRAM bytes per launch of each kernel family as JSON.
usage: python tools/ncu_raw_select.py gpurun_out/r02_ncu_lfa_cl_raw.csv [--traffic]"""
import collections, csv, json, re, sys

WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 20]
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
fam_bytes = collections.defaultdict(list)
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
    print(name)
    for w in WANT:
        if w in d:
            print(f"    {w:<80s} {d[w]} {u[w]}")
    fam = re.sub(r"_kernel.*", "", name)
    mode = re.findall(r"\d+", name)
    if fam == "lfa_cl_bwd" or fam == "lfa_cl_wide":
        # MODE: bwd <D,K,MODE,NG>, wide <K,MODE,NG>; forward launches of the wide kernel have MODE 0
        m = int(mode[2] if fam == "lfa_cl_bwd" else mode[1])
        fam = "lfa_cl_fwd" if m == 0 else {1: "lfa_cl_bwd", 2: "lfa_cl_bwd", 3: "lfa_cl_bn2", 4: "lfa_cl_mom"}[m]
    fam_bytes[fam].append(float(d["dram__bytes_read.sum"].replace(",", "")) + float(d["dram__bytes_write.sum"].replace(",", "")))
if "--traffic" in sys.argv:
    print(json.dumps({k: {"bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in fam_bytes.items()}, indent=1))

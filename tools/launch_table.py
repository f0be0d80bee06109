"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_table.py launches.csv <steps in the capture> [top]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
steps = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
skip = ("fp32_probe",)
for r in rows[1:]:
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[ki]))[:78]
    if any(s in name for s in skip): continue
    us = float(r[vi].replace(',', '')) * (1e-3 if r[ui].startswith('n') else 1.0)
    agg[name][0] += 1; agg[name][1] += us
tot = sum(v for _, v in agg.values()); n = sum(c for c, _ in agg.values())
mine = sum(v for k, (c, v) in agg.items() if k.startswith('r3d::')); nm = sum(c for k, (c, v) in agg.items() if k.startswith('r3d::'))
print(f"# {n/steps:.0f} launches/step, {tot/steps:.0f} us/step (cold-cache, serialised);  r3d:: kernels {nm/steps:.0f} launches, {mine/steps:.0f} us/step ({mine/tot:.1%})")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:<80s} n/step {c/steps:6.1f}  us/step {v/steps:8.1f}  avg us {v/c:7.1f}  share {v/tot:.3f}")

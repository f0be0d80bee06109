import importlib, sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, os.getcwd())
import torch
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
kb = importlib.import_module("knn_bench")
for dens in (0.0, 1.0, 3.0, 4.0, 6.0, 8.0, 12.0):
    kb.L.r3d_knn_set_grid_density(dens)
    row = []
    for shp in [(8, 2500, 2500, 16), (64, 40960, 40960, 16), (1, 1 << 20, 1 << 20, 16), (8, 2500, 2500, 32), (1, 1 << 20, 1 << 20, 32), (64, 10240, 40960, 1)]:
        r = kb.time_knn(*shp, variant=2, algo=2)
        row.append("%.3f" % r["ms"])
    print("density", dens, row, flush=True)

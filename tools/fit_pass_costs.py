"""Dev tool: fit per-op-class costs of the tile kernel from an ncu launch list of bench.py
(`ncu --metrics gpu__time_duration.sum ... python bench.py --steps 1 --warmup 1`) and the plan's own pass
descriptors.  Usage: python tools/fit_pass_costs.py gpurun_out/launches.csv [n_qubits]"""
import csv
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402
from tests import tile_emulator as te  # noqa: E402

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tiles = [float(r[vi].replace(",", "")) * 1e-6 for r in data if "k_tile2" in r[ki]]
p = eng.Plan(n, wl.to_ops(wl.brickwork(n, 20)), eng.PLAN_FUSE)
P = int(p.info()["n_fused_passes"])
last = tiles[P:2 * P]          # second run of the plan (the first is cold)
print(f"{P} passes, {sum(last):.2f} ms")
cols = ["base", "shr", "shr_cy", "shi", "shi_cy", "shi_py", "gen", "perm", "phase", "resplit"]
X, Y = [], []
for i in range(P):
    head, segs, ops = te.parse(p.export_pass(i))
    c = dict.fromkeys(cols, 0.0)
    c["base"] = 1
    for o in ops:
        if o.kind == 0:
            c["shr_cy" if o.flags & 64 else "shr"] += 1
        elif o.kind == 1:
            c["shi_py" if (o.flags & 8 and (not o.flags & 64 or o.flags & 48)) else ("shi_cy" if o.flags & 64 else "shi")] += 1
        elif o.kind == 2:
            c["gen"] += bin(o.mask).count("1") / 16
        elif o.kind in (3, 4):
            c["perm"] += 1
        else:
            c["phase"] += 1
    c["resplit"] = sum(s.resplit for s in segs)
    print(i, f"{last[i]:6.2f} ms", len(ops), {k: round(v, 1) for k, v in c.items() if v and k != "base"})
    X.append([c[k] for k in cols])
    Y.append(last[i])
X, Y = np.array(X), np.array(Y)
from scipy.optimize import nnls  # noqa: E402
coef, _ = nnls(X, Y)
print("ms per unit:", {k: round(float(v), 3) for k, v in zip(cols, coef)})
print("totals ms:  ", {k: round(float(v * X[:, j].sum()), 1) for j, (k, v) in enumerate(zip(cols, coef))})
print("residuals:", np.round(X @ coef - Y, 2))

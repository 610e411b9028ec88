"""Dev tool (GPU): fourier_transform(n) fused on ONE GPU (n = 34 is a 128 GiB state: fits one 180 GB B200).
Usage: python tools/qft_single.py [n]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 34
eng.init(0)
gates = wl.qft(n)
plan = eng.Plan(n, wl.to_ops(gates), eng.PLAN_FUSE | eng.PLAN_JIT)
s = eng.State(n)
x = 0x2b3c4d5e6 & ((1 << n) - 1)
t = eng.Timer()
ms = []
for rep in range(3):
    s.set_basis(x)
    t.start(s)
    s.run(plan)
    t.stop(s)
    ms.append(t.elapsed_ms())
rev = int(format(x, f"0{n}b")[::-1], 2)
y = np.random.default_rng(1).integers(0, 1 << n, 64)
want = np.exp(2j * np.pi * ((rev * y.astype(object)) % (1 << n)).astype(np.float64) / float(1 << n)) / np.sqrt(float(1 << n))
got = np.array([s.amp(int(k)) for k in y])
err = float(np.max(np.abs(got - want)) * np.sqrt(float(1 << n)))
print(json.dumps({"workload": f"fourier_transform({n}) on one GPU, state {8 * 2 ** n / 2 ** 30:.0f} GiB", "ms": min(ms), "ms_all": ms,
                  "gate_apps": len(gates), "passes": int(plan.info()["n_fused_passes"]), "max_rel_amp_error_vs_closed_form": err,
                  "norm2": s.norm2()}))

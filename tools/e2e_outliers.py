"""Dev tool: find which part of the e2e step produces the occasional +500 ms outlier."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import aqs  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = 30
aqs.initialize(0)
gates = wl.brickwork(n, 20)
u = np.random.default_rng(0).random(1000, dtype=np.float32)
for it in range(14):
    t = [time.perf_counter()]
    qc = aqs.QCircuit(n).extend(gates); t.append(time.perf_counter())
    qs = aqs.QSimulator(n); t.append(time.perf_counter())
    qs.simulate(qc); t.append(time.perf_counter())
    out = qs.sample(u); t.append(time.perf_counter())
    del qs; t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print(f"it {it:2d}: build {d[0]:7.1f}  alloc {d[1]:7.1f}  simulate {d[2]:7.1f}  sample {d[3]:7.1f}  free {d[4]:7.1f}  total {sum(d):7.1f}")

"""Dev tool: where does the end-to-end step (bench.py `e2e`) spend its time? (run on the GPU box)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import aqs  # noqa: E402
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
aqs.initialize(0)
gates = wl.brickwork(n, 20)
u = np.random.default_rng(0).random(1000, dtype=np.float32)


def tick(label, fn, sync=None):
    t0 = time.perf_counter()
    out = fn()
    if sync is not None:
        sync.sync()
    print(f"{label:42s} {1e3 * (time.perf_counter() - t0):9.2f} ms")
    return out


for rep in range(3):
    print("--- rep", rep)
    qc = tick("build QCircuit (890 ctypes calls)", lambda: aqs.QCircuit(n).extend(gates))
    qs = tick("QSimulator(n): cudaMalloc + memset", lambda: aqs.QSimulator(n))
    qs.sync()
    ops = tick("lower to ops", lambda: qc.ops())
    plan = tick("aqs_plan_build (fusion planner)", lambda: eng.Plan(n, ops, eng.PLAN_FUSE))
    st = qs.engine_state()
    tick("plan run (first: uploads descriptors)", lambda: st.run(plan), sync=st)
    tick("plan run (second)", lambda: st.run(plan), sync=st)
    tick("simulate(qc) whole call", lambda: qs.simulate(qc), sync=qs)
    tick("sample 1000 draws", lambda: qs.sample(u))
    tick("free QSimulator (cudaFree)", lambda: qs.__del__())
    qs._h = None

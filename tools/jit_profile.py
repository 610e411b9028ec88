#!/usr/bin/env python
"""Dev tool for ncu: brickwork-n, specialised kernels, `runs` plan runs back to back (profile the last one)."""
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = eng.PLAN_FUSE | (0 if (len(sys.argv) > 3 and sys.argv[3] == "interp") else eng.PLAN_JIT)
eng.init(0)
plan = eng.Plan(n, wl.to_ops(wl.brickwork(n, 20)), flags)
st = eng.State(n)
for _ in range(runs):
    st.run(plan)
st.sync()
print("passes", plan.info()["n_fused_passes"], "jit_ready", plan.jit_ready(), "norm2", st.norm2())

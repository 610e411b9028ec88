#!/usr/bin/env python
"""Dev tool: brickwork-n through the generic tile kernel and through the specialised kernels, whole plan and pass by pass."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng.init(0)
import os
if os.environ.get("BENCH_OPS") == "aqs":
    from afquantumsim_b200 import aqs
    aqs.initialize(0)
    ops = aqs.QCircuit(n).extend(wl.brickwork(n, 20)).ops()
else:
    ops = wl.to_ops(wl.brickwork(n, 20))
st = eng.State(n)
timer = eng.Timer()
out = {"n": n}
legs = (("jit", eng.PLAN_FUSE | eng.PLAN_JIT),) if os.environ.get("BENCH_JIT_ONLY") else (("interp", eng.PLAN_FUSE), ("jit", eng.PLAN_FUSE | eng.PLAN_JIT))
for name, flags in legs:
    t0 = time.perf_counter()
    plan = eng.Plan(n, ops, flags)
    build_s = time.perf_counter() - t0
    passes = plan.info()["n_fused_passes"]
    st.set_basis(0)
    for _ in range(3):
        st.run(plan)
    st.sync()
    timer.start(st)
    for _ in range(K):
        st.run(plan)
    timer.stop(st)
    ms = timer.elapsed_ms() / K
    per = []
    for i in range(passes):
        st.run_shard(plan, i, 1, 0, 0)
        timer.start(st)
        for _ in range(3):
            st.run_shard(plan, i, 1, 0, 0)
        timer.stop(st)
        per.append(round(timer.elapsed_ms() / 3, 3))
    out[name] = {"ms": ms, "build_s": build_s, "passes": passes, "jit_ready": plan.jit_ready(), "per_pass_ms": per, "norm2": st.norm2()}
out["jit_info"] = eng.jit_info()
print(json.dumps(out))

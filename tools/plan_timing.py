"""Dev tool (no GPU needed): host time of the fusion planner with 1 / 4 / 12 seeded variants."""
import os
import sys
import time

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ops = wl.to_ops(wl.brickwork(n, 20))
print("cpus", os.cpu_count())
for v in ("1", "4", "12"):
    os.environ["AQS_PLAN_VARIANTS"] = v
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        p = eng.Plan(n, ops, eng.PLAN_FUSE)
        ts.append((time.perf_counter() - t) * 1e3)
    print(f"variants={v}: {min(ts):.1f} ms (best of 5), passes {p.info()['n_fused_passes']}")

"""Dev tool (no GPU needed): fusion-planner statistics for the BASELINE circuits."""
import os
import sys

sys.path.insert(0, ".")
os.environ.setdefault("AQS_PLAN_DUMP", "1")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.lowering import lower_array  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "brickwork"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
if which == "brickwork":
    circ = orc.Circ(n, wl.brickwork(n, 20))
elif which == "qft":
    circ = orc.Circ(n, wl.qft(n))
elif which == "grover":
    circ = orc.grover_search(n, orc.grover_oracle(n, 5), 8)
else:
    circ = orc.Circ(n, wl.ghz(n))
plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
print(plan.info())

"""Dev tool (GPU): per-op cost of the tile kernel in controlled settings.
Usage: python tools/tile_microbench.py [n_qubits]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.lowering import lower_array  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
eng.init(0)
s = eng.State(n)
rng = np.random.default_rng(0)


def timed(gates, reps=3):
    plan = eng.Plan(n, lower_array(orc.Circ(n, gates)), eng.PLAN_FUSE)
    info = plan.info()
    best = 1e9
    for _ in range(reps):
        t = eng.Timer()
        t.start(s)
        s.run(plan)
        t.stop(s)
        best = min(best, t.elapsed_ms())
    return best, info["n_fused_passes"]


def ang():
    return float(np.float32(rng.uniform(0.1, 3.0)))


# API qubit q <-> bit n-1-q
def qb(bit):
    return n - 1 - bit


cases = {}
for N in (1, 16, 64, 128):
    cases[f"alternating RotY/RotX on bit 20, N={N}"] = [("RotY" if i % 2 == 0 else "RotX", qb(20), ang()) for i in range(N)]
for N in (16, 64):
    cases[f"RotY/RotX cycling over bits 20..24, N={N}"] = [("RotY" if (i // 5) % 2 == 0 else "RotX", qb(20 + i % 5), ang()) for i in range(N)]
    cases[f"RotZ cycling over bits 20..24 then RotY (phases), N={N}"] = [("RotZ" if i % 2 else "RotY", qb(20 + (i // 2) % 5), ang()) for i in range(N)]
    cases[f"RotY + CX neighbour (mux) over bits 20..24, N={N}"] = sum(([("RotY", qb(20 + i % 5), ang()), ("CX", qb(20 + (i + 1) % 5), qb(20 + i % 5))] for i in range(N // 2)), [])
    cases[f"RotY on low bits 0..4 cycling, N={N}"] = [("RotY" if (i // 5) % 2 == 0 else "RotX", qb(i % 5), ang()) for i in range(N)]
for name, gates in cases.items():
    ms, passes = timed(gates)
    print(f"{name:70s} {ms:8.3f} ms  passes {passes}  -> {ms / len(gates) * 1e3:8.1f} us/gate")

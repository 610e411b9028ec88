"""Dev tool: per-gate kernel bandwidth by target bit position (run on the GPU box).
Usage: python tools/sweep_gate_bw.py [n_qubits]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng.init(0)
s = eng.State(n)
S = 8.0 * (1 << n)
H = np.float32(0.70710678118)
t = eng.Timer()
rows = []


def timeit(ops, reps=5):
    for _ in range(2):
        s.apply_ops(ops)
    t.start(s)
    for _ in range(reps):
        s.apply_ops(ops)
    t.stop(s)
    return t.elapsed_ms() / reps


for p in list(range(0, 12)) + [15, 20, 25, n - 1]:
    q = n - 1 - p
    ms = timeit(eng.op_record(eng.OP_U2, q, [H, H, H, -H]))
    rows.append(("u2", p, ms, 2 * S / ms / 1e6))
    ms = timeit(eng.op_record(eng.OP_DIAG, q, [np.exp(-0.3j), 0, 0, np.exp(0.3j)]))
    rows.append(("rotz", p, ms, 2 * S / ms / 1e6))
for (pc, pt) in [(1, 0), (0, 1), (2, 1), (10, 9), (9, 10), (20, 3), (3, 20), (n - 1, n - 2), (n - 2, n - 1)]:
    ms = timeit(eng.op_record(eng.OP_X, n - 1 - pt, controls=(n - 1 - pc,)))
    rows.append((f"cx c{pc}", pt, ms, S / ms / 1e6))
    ms = timeit(eng.op_record(eng.OP_DIAG, n - 1 - pt, [1, 0, 0, 1j], controls=(n - 1 - pc,)))
    rows.append((f"cphase c{pc}", pt, ms, 0.5 * S / ms / 1e6))
for kind, p, ms, gbs in rows:
    print(f"{kind:12s} p={p:2d}  {ms:8.3f} ms  {gbs:8.1f} GB/s (algorithmic)")
json.dump(rows, open("gpurun_out/sweep_gate_bw.json", "w"))

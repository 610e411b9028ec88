"""Dev tool: cost model of the fused tile kernel (run on the GPU box).
Single-pass plans with chosen tile bits / op counts at n qubits."""
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng.init(0)
s = eng.State(n)
S = 8.0 * (1 << n)
t = eng.Timer()
rng = np.random.default_rng(0)


def rot(q, ctrl=()):
    th = rng.uniform(0, 6.28)
    c, sn = np.cos(th / 2), np.sin(th / 2)
    return eng.op_record(eng.OP_U2, q, [c, -1j * sn, -1j * sn, c], controls=ctrl)


import os
ONLY = os.environ.get("SWEEP_ONLY", "")


def timeit(ops, reps=5, label=""):
    if ONLY and not any(tok in label for tok in ONLY.split("|")):
        return 0.0
    if ONLY:
        reps = 1
    plan = eng.Plan(n, np.concatenate(ops), eng.PLAN_FUSE)
    info = plan.info()
    for _ in range(2):
        s.run(plan)
    t.start(s)
    for _ in range(reps):
        s.run(plan)
    t.stop(s)
    ms = t.elapsed_ms() / reps
    print(f"{label:46s} passes={info['n_fused_passes']:2d} ops={len(ops):3d}  {ms:8.3f} ms  "
          f"{ms / max(1, info['n_fused_passes']):7.3f} ms/pass  {info['bytes_planned'] / ms / 1e6:8.1f} GB/s")
    return ms


def bits_to_q(bits):
    return [n - 1 - b for b in bits]


for name, bits in (("low bits 5..11 (contiguous 32 KiB tile)", range(5, 12)), ("bits 12..18", range(12, 19)),
                   ("bits 16..22", range(16, 23)), ("high bits 23..29", range(n - 7, n)),
                   ("spread 5,9,13,17,21,25,29", (5, 9, 13, 17, 21, 25, n - 1))):
    qs = bits_to_q(bits)
    timeit([rot(qs[0])], label=name + " 1 op")
    timeit([rot(q) for q in qs[:4]], label=name + " 4 reg ops")
    timeit([rot(q) for q in qs], label=name + " 7 ops (2 segs)")
    many = []
    for r in range(4):
        many += [rot(q, ctrl=(qs[(i + 1) % 4],)) for i, q in enumerate(qs[:4])]
    timeit(many, label=name + " 16 ctrl-ops 1 seg")
lane = bits_to_q(range(0, 5))
timeit([rot(q) for q in lane], label="lane bits 0..4: 5 ops")
many = []
for r in range(4):
    many += [rot(q, ctrl=(lane[(i + 1) % 5],)) for i, q in enumerate(lane)]
timeit(many, label="lane bits: 20 ctrl-ops")
diag = [eng.op_record(eng.OP_DIAG, n - 1 - b, [np.exp(-0.3j), 0, 0, np.exp(0.3j)]) for b in range(n)]
timeit(diag, label="30 RotZ (diag only)")
cph = [eng.op_record(eng.OP_DIAG, n - 1 - b, [1, 0, 0, np.exp(0.3j)], controls=(n - 1 - ((b + 7) % n),)) for b in range(n)]
timeit(cph, label="30 CPhase (diag only)")

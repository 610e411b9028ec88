"""Dev tool (GPU): time of aqs_apply_dense at n qubits for k = 1..6 on high target bits and on the low bits."""
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from tests.dense_cases import random_unitary  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng.init(0)
s = eng.State(n)
rng = np.random.default_rng(0)
S = 8.0 * (1 << n)
for k in range(1, 7):
    U = random_unitary(k, rng)
    for name, qubits in (("high bits", list(range(3, 3 + k))), ("low bits", list(range(n - k, n)))):
        s.apply_dense(qubits, U)
        s.sync()
        t = eng.Timer()
        t.start(s)
        for _ in range(3):
            s.apply_dense(qubits, U)
        t.stop(s)
        ms = t.elapsed_ms() / 3
        print(f"k={k} {name:9s} {ms:8.3f} ms   {2 * S / ms / 1e6:7.1f} GB/s algorithmic   {8.0 * (1 << k) * (1 << n) / ms / 1e9:8.2f} TFLOP/s")
print("norm2", s.norm2())

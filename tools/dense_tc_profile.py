"""Dev tool for ncu: one aqs_apply_dense (k = 6, high target bits) on the tensor-core kernel at n qubits."""
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from tests.dense_cases import random_unitary  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng.init(0)
s = eng.State(n)
U = random_unitary(6, np.random.default_rng(0))
for _ in range(3):
    s.apply_dense(list(range(3, 9)), U)
s.sync()
print("norm2", s.norm2())

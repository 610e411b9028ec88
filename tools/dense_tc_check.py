"""Dev tool (GPU): aqs_apply_dense on the tensor-core path against numpy, structured matrices first."""
import sys

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from tests.dense_cases import dense_reference, random_unitary, rel_l2  # noqa: E402

eng.init(0)
rng = np.random.default_rng(1)
n = 14
a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
a /= np.float32(np.linalg.norm(a))


def run(name, qubits, U, controls=(), cv=None):
    s = eng.State(n)
    s.upload(a)
    s.apply_dense(list(qubits), U, list(controls), cv)
    got = s.download()
    s.close()
    want = dense_reference(a, n, list(qubits), U, list(controls), cv)
    print(f"{name:40s} rel_l2 {rel_l2(got, want):.3e}", flush=True)


k = 6
D = 1 << k
hi = list(range(1, 1 + k))
run("identity k=6 high", hi, np.eye(D, dtype=np.complex64))
P = np.eye(D, dtype=np.complex64)[np.roll(np.arange(D), 1)]
run("cyclic permutation k=6 high", hi, P)
run("diag phases k=6 high", hi, np.diag(np.exp(1j * rng.uniform(0, 6, D))).astype(np.complex64))
R = rng.standard_normal((D, D)).astype(np.float32).astype(np.complex64)
run("real random (non-unitary) k=6 high", hi, (R / 8).astype(np.complex64))
run("random unitary k=6 high", hi, random_unitary(k, rng))
run("random unitary k=6 low", list(range(n - k, n)), random_unitary(k, rng))
run("random unitary k=6 scattered", [0, 3, 5, 8, 10, 13], random_unitary(k, rng))
run("random unitary k=5", [2, 4, 6, 9, 11], random_unitary(5, rng))
run("random unitary k=4", [1, 7, 8, 12], random_unitary(4, rng))
run("random unitary k=5 controlled", [2, 4, 6, 9, 11], random_unitary(5, rng), [0, 13], 1 << 13)

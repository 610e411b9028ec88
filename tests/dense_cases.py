"""aqs_apply_dense checks shared by the CPU (ABI stand-in) and GPU suites.

  python tests/dense_cases.py --abi cpu

1. arbitrary target qubits, with and without controls, against a numpy restatement (einsum over the target axes);
2. the reference's own placement — a compiled k-qubit circuit on contiguous qubits [b, b + k), optionally under a
   control (Gate / ControlGate, src/quantum.cpp:1760-1814, 1888-1950) — against the oracle's literal dense embedding.
Tolerance 1e-5 relative L2 (BASELINE.json north_star).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TOL = 1e-5


def random_unitary(k, rng):
    z = rng.standard_normal((1 << k, 1 << k)) + 1j * rng.standard_normal((1 << k, 1 << k))
    q, r = np.linalg.qr(z)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).astype(np.complex64)


def dense_reference(a, n, qubits, U, controls=(), ctrl_value=None):
    """numpy: apply U (row-major, qubits[0] = most significant matrix-index bit) to the axes of `qubits`"""
    k = len(qubits)
    t = a.astype(np.complex128).reshape([2] * n)            # axis q = API qubit q (qubit 0 is the MSB)
    sel = [slice(None)] * n
    cm = 0
    for c in controls:
        cm |= 1 << c
    cv = cm if ctrl_value is None else ctrl_value
    for c in controls:
        sel[c] = (cv >> c) & 1
    sub = t[tuple(sel)]
    rest = [q for q in range(n) if q not in controls]
    axes = [rest.index(q) for q in qubits]
    Ut = U.astype(np.complex128).reshape([2] * (2 * k))
    out = np.tensordot(Ut, sub, axes=(list(range(k, 2 * k)), axes))      # new axes first, in `qubits` order
    out = np.moveaxis(out, list(range(k)), axes)
    t[tuple(sel)] = out
    return t.reshape(-1).astype(np.complex64)


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / np.linalg.norm(b.astype(np.complex128)))


def run_cases(eng, orc, sizes=((9, 40), (13, 30))):
    rng = np.random.default_rng(2024)
    for n, reps in sizes:
        for rep in range(reps):
            k = int(rng.integers(1, min(6, n - 2) + 1))
            qubits = [int(x) for x in rng.choice(n, k, replace=False)]
            others = [q for q in range(n) if q not in qubits]
            nc = int(rng.integers(0, 3))
            controls = [int(x) for x in rng.choice(others, nc, replace=False)] if nc else []
            cv = 0
            for c in controls:
                cv |= int(rng.integers(0, 2)) << c
            U = random_unitary(k, rng)
            a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
            a /= np.float32(np.linalg.norm(a))
            s = eng.State(n)
            s.upload(a)
            s.apply_dense(qubits, U, controls, cv if controls else None)
            got = s.download()
            s.close()
            err = rel_l2(got, dense_reference(a, n, qubits, U, controls, cv if controls else None))
            assert err < TOL, (n, qubits, controls, cv, err)
    # the reference's placements: Gate{circ, b} and ControlGate{circ, c, b}
    for n, k, b, ctrl in ((8, 3, 2, None), (9, 4, 5, 1), (9, 4, 0, 7), (12, 5, 3, 10), (12, 6, 6, 0), (14, 6, 2, None)):
        gates = []
        for _ in range(6 * k):
            q = int(rng.integers(k))
            gates.append([("H", q), ("RotX", q, float(rng.uniform(-3, 3))), ("RotY", q, float(rng.uniform(-3, 3))),
                          ("Phase", q, float(rng.uniform(-3, 3)))][int(rng.integers(4))])
            if k > 1:
                c, t = [int(x) for x in rng.choice(k, 2, replace=False)]
                gates.append(("CX", c, t))
        inner = orc.Circ(k, gates)
        U = orc.circuit_matrix(inner, mode="dense")                 # U[r, c]
        outer = orc.Circ(n, [("Gate", inner, b)] if ctrl is None else [("ControlGate", inner, ctrl, b)])
        a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
        a /= np.float32(np.linalg.norm(a))
        want = orc.simulate(a.copy(), outer, mode="dense")
        s = eng.State(n)
        s.upload(a)
        s.apply_dense(list(range(b, b + k)), U, [] if ctrl is None else [ctrl])
        err = rel_l2(s.download(), want)
        s.close()
        assert err < TOL, (n, k, b, ctrl, err)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--abi", default="cpu")
    a = ap.parse_args()
    from afquantumsim_b200 import engine as eng
    from oracle import oracle as orc
    if a.abi == "cpu":
        eng.LIB_PATH = os.path.join(ROOT, "oracle", "_build", "cpu_abi", "libaqs_engine.so")   # test double
    eng.init(0)
    run_cases(eng, orc)
    print("ok dense_cases")

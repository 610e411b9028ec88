"""Test / benchmark infrastructure: the reference's own ALGORITHM restated on the CPU (never imported by the package).

afQuantumSim applies every gate by building an explicit 2^n x 2^n operator and multiplying it in: a CSR matrix with one
non-zero per row for permutation / diagonal gates (X, Z, Phase, CX, CZ, CPhase, CCNot, ... : src/quantum.cpp:540-1653) and a
DENSE Kronecker product I (x) U (x) I for H and the rotations (src/quantum.cpp:671-831, src/utils.cpp:169-195), single precision.
This module does exactly that with numpy / scipy.sparse, so that the cost of the reference's algorithm can be timed on the
box's host cores beside the matrix-free oracle — ArrayFire itself is not installable here.  Sizes are the
reference's own published ones (benchmark/results.md:9-16: QFT-10, Grover-10) plus GHZ-12 / 13 to show the 4^n wall.
Checked against the oracle in tests/test_oracle.py."""
import time

import numpy as np
import scipy.sparse as sp

C = np.complex64


def _u2(name, theta=0.0):
    c, s = np.float32(np.cos(np.float32(theta) / 2)), np.float32(np.sin(np.float32(theta) / 2))
    h = np.float32(0.70710678118)
    return {"H": np.array([[h, h], [h, -h]], dtype=C), "RotX": np.array([[c, -1j * s], [-1j * s, c]], dtype=C),
            "RotY": np.array([[c, -s], [s, c]], dtype=C), "RotZ": np.array([[c - 1j * s, 0], [0, c + 1j * s]], dtype=C)}[name]


def gate_operator(n, gate):
    """the explicit 2^n x 2^n operator the reference builds for one gate (qubit 0 = most significant index bit)"""
    N = 1 << n
    name, args = gate[0], gate[1:]
    m = lambda q: 1 << (n - 1 - q)
    r = np.arange(N, dtype=np.int64)
    if name in ("H", "RotX", "RotY", "RotZ"):
        t = args[0]
        u = _u2(name, args[1] if len(args) > 1 else 0.0)
        return np.kron(np.kron(np.eye(1 << t, dtype=C), u), np.eye(1 << (n - 1 - t), dtype=C))       # dense, like tensor_product
    if name in ("X", "CX", "CCNot"):
        t, ctrls = args[-1], args[:-1]
        on = np.ones(N, dtype=bool)
        for c_ in ctrls:
            on &= (r & m(c_)) != 0
        cols = np.where(on, r ^ m(t), r)
        return sp.csr_matrix((np.ones(N, dtype=C), cols, np.arange(N + 1)), shape=(N, N))
    if name in ("Z", "Phase", "CZ", "CPhase"):
        ang = args[-1] if name in ("Phase", "CPhase") else np.pi
        qs = args[:-1] if name in ("Phase", "CPhase") else args
        on = np.ones(N, dtype=bool)
        for q in qs:
            on &= (r & m(q)) != 0
        f = C(np.cos(np.float32(ang)) + 1j * np.sin(np.float32(ang))) if name in ("Phase", "CPhase") else C(-1)
        return sp.csr_matrix((np.where(on, f, C(1)).astype(C), r, np.arange(N + 1)), shape=(N, N))
    if name == "MCZ":                      # (n-1)-controlled Z over all qubits: NControl_Gate(Z) in grover_oracle / the diffuser
        d = np.ones(N, dtype=C)
        d[N - 1] = -1
        return sp.csr_matrix((d, r, np.arange(N + 1)), shape=(N, N))
    raise ValueError(name)


def simulate(n, gates, state=None):
    a = np.zeros(1 << n, dtype=C) if state is None else state.astype(C)
    if state is None:
        a[0] = 1
    for g in gates:
        op = gate_operator(n, g)
        a = (op @ a).astype(C)
    return a


def qft(n):
    g = []
    for i in range(n - 1, -1, -1):
        g.append(("H", i))
        for j in range(i):
            g.append(("CPhase", j, i, float(np.pi / (1 << (i - j)))))
    return g


def grover(n, marked, iterations):
    zero_bits = [i for i in range(n) if not (marked >> i) & 1]
    oracle = [("X", q) for q in zero_bits] + [("MCZ",)] + [("X", q) for q in zero_bits]
    diff = [("H", q) for q in range(n)] + [("X", q) for q in range(n)] + [("MCZ",)] + [("X", q) for q in range(n)] + [("H", q) for q in range(n)]
    return [("H", q) for q in range(n)] + (oracle + diff) * iterations


def ghz(n):
    return [("H", 0)] + [("CX", i, i + 1) for i in range(n - 1)]


def timed(n, gates, runs):
    if runs > 1:
        simulate(n, gates)          # warm-up
    ts = []
    for _ in range(runs):
        t0 = time.perf_counter()
        simulate(n, gates)
        ts.append((time.perf_counter() - t0) * 1e3)
    return {"qubits": n, "gates": len(gates), "runs": runs, "ms_mean": float(np.mean(ts)), "ms_sd": float(np.std(ts, ddof=1)) if runs > 1 else 0.0}


def published_sizes(budget_s=12.0):
    """mean +- sd like benchmark/benchmark.cpp:174-197; published on other hardware: QFT-10 6.05 ms, Grover-10 192 ms"""
    out = {"qft10": timed(10, qft(10), 20), "grover10": timed(10, grover(10, 5, 25), 1),
           "ghz12": timed(12, ghz(12), 3), "ghz13": timed(13, ghz(13), 1),
           "published": {"qft10_ms": 6.05, "grover10_ms": 192.0, "source": "benchmark/results.md:9-16 (MacBook Pro i9-9980H, ArrayFire 3.8.2 OpenCL)"},
           "what": "the reference's algorithm (explicit CSR / dense-Kronecker operator per gate, complex64) in numpy + scipy.sparse on the host cores"}
    return out

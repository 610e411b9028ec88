"""Synthetic circuits of BASELINE.json's configs, as plain gate tuples
``(ClassName, *ctor_args)`` in the reference's constructor order.  Both the
product path (afquantumsim_b200.aqs) and the test oracle consume this format."""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


def ghz(n: int):
    """config 1: H{0} then CX{i,i+1}."""
    return [("H", 0)] + [("CX", i, i + 1) for i in range(n - 1)]


def brickwork(n: int, depth: int = 20, seed: int | None = None):
    """config 3/5: per layer d, one random RotX/RotY/RotZ(theta) on every qubit
    (theta ~ U[0, 2pi) as f32), then CX{q,q+1} for q = d mod 2, +2, ...
    numpy PCG64(seed = n by default).  30 qubits x depth 20 = 600 rotations +
    290 CX = 890 gate applications."""
    rng = np.random.Generator(np.random.PCG64(n if seed is None else seed))
    names = ("RotX", "RotY", "RotZ")
    gates = []
    for d in range(depth):
        for q in range(n):
            k = int(rng.integers(0, 3))
            theta = float(np.float32(rng.random() * TWO_PI))
            gates.append((names[k], q, theta))
        for q in range(d % 2, n - 1, 2):
            gates.append(("CX", q, q + 1))
    return gates


def qft(n: int):
    """config 2/5: fourier_transform(n), src/quantum_algo.cpp:103-114."""
    pi = np.float32(3.14159265358979323846)
    gates = []
    for i in range(n - 1, -1, -1):
        gates.append(("H", i))
        for j in range(i):
            gates.append(("CPhase", j, i, float(pi / np.float32(1 << (i - j)))))
    return gates


def to_ops(gates):
    """Lower workload gate tuples straight to engine ops (struct aqs_op records), for states beyond the
    30-qubit limit of the drop-in QCircuit API.  Matrices as in the host layer.
    One preallocated record array, filled gate by gate (a record per call cost 40 us: 37 ms for brickwork-33)."""
    from . import engine as eng
    f32 = np.float32
    h = f32(0.70710678118)
    ops = eng.make_ops(len(gates))
    kind, target, cmask, m = ops["kind"], ops["target"], ops["ctrl_mask"], ops["m"]
    for i, g in enumerate(gates):
        name = g[0]
        if name == "H":
            kind[i], target[i] = eng.OP_U2, g[1]
            m[i] = (h, 0, h, 0, h, 0, -h, 0)
        elif name == "X":
            kind[i], target[i] = eng.OP_X, g[1]
            m[i] = (1, 0, 0, 0, 0, 0, 1, 0)
        elif name == "CX":
            kind[i], target[i], cmask[i] = eng.OP_X, g[2], 1 << g[1]
            m[i] = (1, 0, 0, 0, 0, 0, 1, 0)
        elif name in ("RotX", "RotY", "RotZ"):
            a = f32(g[2])
            c, s = np.cos(a / f32(2), dtype=f32), np.sin(a / f32(2), dtype=f32)
            target[i] = g[1]
            if name == "RotX":
                kind[i] = eng.OP_U2
                m[i] = (c, 0, 0, -s, 0, -s, c, 0)
            elif name == "RotY":
                kind[i] = eng.OP_U2
                m[i] = (c, 0, -s, 0, s, 0, c, 0)
            else:
                kind[i] = eng.OP_DIAG
                m[i] = (c, -s, 0, 0, 0, 0, c, s)
        elif name == "CPhase":
            a = f32(g[3])
            kind[i], target[i], cmask[i] = eng.OP_DIAG, g[2], 1 << g[1]
            m[i] = (1, 0, 0, 0, 0, 0, np.cos(a, dtype=f32), np.sin(a, dtype=f32))
        else:
            raise KeyError(name)
    ops["ctrl_value"] = ops["ctrl_mask"]
    return ops

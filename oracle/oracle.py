"""Python front end of the CPU parity oracle.

TEST INFRASTRUCTURE ONLY — see the header of ``aqs_oracle.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Nothing in ``afquantumsim_b200/`` does.

A circuit is plain data: ``Circ(n, gates)`` where every gate is a tuple whose
first element is the reference class name (include/quantum.h:857-1589) and the
rest are the constructor arguments in the reference order::

    ("H", t)  ("CX", c, t)  ("RotX", t, angle)  ("CPhase", c, t, angle)
    ("Swap", a, b)  ("CSwap", c, a, b)  ("CCNot", a, b, t)  ("Or", a, b, t)
    ("Gate", Circ, begin)  ("ControlGate", Circ, ctrl, begin)  ("Barrier",)

Two ways to apply ``Gate``/``ControlGate``:
  * ``mode="dense"``  — literally what the reference does: compile the inner
    circuit to a 2^k x 2^k matrix and embed it (src/quantum.cpp:1760-1814,
    1888-1950).  Only for small inner circuits.
  * ``mode="flatten"`` — recurse into the inner gate list, adding the qubit
    offset and accumulating the control mask.  The reference's own tests equate
    the two (test/tests.cpp:909-961, 1049-1108); tests/test_oracle.py checks it.

The builders at the bottom restate src/quantum_gates.cpp and
src/quantum_algo.cpp:16-129.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libaqs_oracle.so")

PI_F32 = float(np.float32(3.14159265358979323846))  # aqs::pi, include/quantum.h:121


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "aqs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_float
        L.orc_apply_gate.argtypes = [vp, i32, i32, ctypes.POINTER(i32), f32, u64]
        L.orc_apply_gate.restype = i32
        L.orc_apply_dense.argtypes = [vp, i32, i32, i32, i32, vp, u64]
        L.orc_apply_dense.restype = i32
        L.orc_set_basis.argtypes = [vp, i32, u64]
        L.orc_set_product.argtypes = [vp, i32, vp]
        L.orc_probabilities.argtypes = [vp, i32, vp]
        L.orc_prob_fixed.argtypes = [vp, i32, u64, u64]
        L.orc_prob_fixed.restype = u64
        L.orc_qubit_prob1.argtypes = [vp, i32, i32]
        L.orc_qubit_prob1.restype = ctypes.c_double
        L.orc_qubit_prob1_seq_f32.argtypes = [vp, i32, i32]
        L.orc_qubit_prob1_seq_f32.restype = f32
        L.orc_norm2.argtypes = [vp, i32]
        L.orc_norm2.restype = ctypes.c_double
        L.orc_collapse_qubit.argtypes = [vp, i32, i32, i32, f32]
        L.orc_sample.argtypes = [vp, i32, vp, u64, vp, i32]
        L.orc_sample.restype = i32
        L.orc_rel_l2.argtypes = [vp, vp, u64]
        L.orc_rel_l2.restype = ctypes.c_double
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


TYPE_ID = {
    "Barrier": 0, "X": 1, "Y": 2, "Z": 3, "H": 4, "Phase": 5, "Swap": 6,
    "RotX": 7, "RotY": 8, "RotZ": 9, "CX": 10, "CY": 11, "CZ": 12, "CH": 13,
    "CPhase": 14, "CSwap": 15, "CRotX": 16, "CRotY": 17, "CRotZ": 18,
    "CCNot": 19, "Or": 20, "Gate": 21, "ControlGate": 22,
}
ALIASES = {"Not": "X", "CNot": "CX", "Xor": "CX", "And": "CCNot"}  # quantum.h:916,1291,1298,1485
N_QUBIT_ARGS = {
    "X": 1, "Y": 1, "Z": 1, "H": 1, "Phase": 1, "RotX": 1, "RotY": 1, "RotZ": 1,
    "Swap": 2, "CX": 2, "CY": 2, "CZ": 2, "CH": 2, "CPhase": 2,
    "CRotX": 2, "CRotY": 2, "CRotZ": 2, "CSwap": 3, "CCNot": 3, "Or": 3,
}
HAS_ANGLE = {"Phase", "RotX", "RotY", "RotZ", "CPhase", "CRotX", "CRotY", "CRotZ"}


@dataclass
class Circ:
    n: int
    gates: List[tuple] = field(default_factory=list)

    def add(self, *gate) -> "Circ":
        self.gates.append(tuple(gate))
        return self

    def extend(self, gates) -> "Circ":
        self.gates.extend(tuple(g) for g in gates)
        return self


def _ptr(arr: np.ndarray):
    return arr.ctypes.data_as(ctypes.c_void_p)


def new_state(n: int, basis: int = 0) -> np.ndarray:
    a = np.zeros(1 << n, dtype=np.complex64)
    a[basis] = 1.0
    return a


def product_state(qstates: Sequence[Tuple[complex, complex]]) -> np.ndarray:
    """generate_statevector (src/quantum.cpp:261-275)."""
    n = len(qstates)
    q = np.asarray(qstates, dtype=np.complex64).reshape(n, 2).copy()
    a = np.empty(1 << n, dtype=np.complex64)
    lib().orc_set_product(_ptr(a), n, _ptr(q))
    return a


def qstate(z: complex, o: complex) -> Tuple[np.complex64, np.complex64]:
    """QState(z, o): normalise in f32 (src/quantum.cpp:145-157)."""
    z = np.complex64(z); o = np.complex64(o)
    f = np.float32
    # ((zr*zr + zi*zi) + or*or) + oi*oi, every step rounded to f32, as the C++ evaluates it
    mag2 = f(f(f(f(z.real) * f(z.real)) + f(f(z.imag) * f(z.imag))) + f(f(o.real) * f(o.real)))
    mag2 = f(mag2 + f(f(o.imag) * f(o.imag)))
    if mag2 == 0:
        raise ValueError("Cannot normalize a null state")
    mag = f(np.sqrt(mag2))
    den = f(mag * mag)   # af::cfloat / float promotes the divisor to complex: x*c/(c*c) (tests.cpp:986-1002 pins it)

    def div(x):
        return f(f(f(x) * mag) / den)
    return (np.complex64(complex(div(z.real), div(z.imag))), np.complex64(complex(div(o.real), div(o.imag))))


def apply_gate(a: np.ndarray, n: int, gate: tuple, offset: int = 0, xctrl: int = 0, mode: str = "flatten") -> None:
    name = ALIASES.get(gate[0], gate[0])
    L = lib()
    if name == "Barrier":
        return
    if name in ("Gate", "ControlGate"):
        inner: Circ = gate[1]
        if name == "Gate":
            ctrl, begin = -1, gate[2] + offset
        else:
            ctrl, begin = gate[2] + offset, gate[3] + offset
        if mode == "dense":
            U = circuit_matrix(inner, mode="dense")  # column-major like af::array
            Uf = np.asfortranarray(U)
            rc = L.orc_apply_dense(_ptr(a), n, inner.n, begin, ctrl, Uf.ctypes.data_as(ctypes.c_void_p), xctrl)
            if rc != 0:
                raise ValueError("orc_apply_dense rejected the gate placement")
        else:
            x2 = xctrl | ((1 << (n - 1 - ctrl)) if ctrl >= 0 else 0)
            for g in inner.gates:
                apply_gate(a, n, g, offset=begin, xctrl=x2, mode=mode)
        return
    nq = N_QUBIT_ARGS[name]
    qs = [int(q) + offset for q in gate[1:1 + nq]]
    angle = float(np.float32(gate[1 + nq])) if name in HAS_ANGLE else 0.0
    arr = (ctypes.c_int * 3)(*(qs + [0] * (3 - nq)))
    rc = L.orc_apply_gate(_ptr(a), n, TYPE_ID[name], arr, angle, xctrl)
    if rc != 0:
        raise ValueError(f"unknown gate {name}")


def simulate(a: np.ndarray, circ: Circ, mode: str = "flatten") -> np.ndarray:
    """QSimulator::simulate (src/quantum.cpp:277-291), in place."""
    assert a.dtype == np.complex64 and a.size == (1 << circ.n)
    for g in circ.gates:
        apply_gate(a, circ.n, g, mode=mode)
    return a


def circuit_matrix(circ: Circ, mode: str = "flatten") -> np.ndarray:
    """QCircuit::compile (src/quantum.cpp:199-210): the dense unitary, U[r, c]."""
    N = 1 << circ.n
    U = np.zeros((N, N), dtype=np.complex64)
    for c in range(N):
        col = new_state(circ.n, c)
        simulate(col, circ, mode=mode)
        U[:, c] = col
    return U


def probabilities(a: np.ndarray) -> np.ndarray:
    n = int(math.log2(a.size))
    out = np.empty(a.size, dtype=np.float32)
    lib().orc_probabilities(_ptr(a), n, _ptr(out))
    return out


def qubit_prob1(a: np.ndarray, qubit: int) -> float:
    n = int(math.log2(a.size))
    return float(lib().orc_qubit_prob1(_ptr(a), n, qubit))


def prob_fixed(a: np.ndarray, mask: int = 0, value: int = 0) -> int:
    n = int(math.log2(a.size))
    return int(lib().orc_prob_fixed(_ptr(a), n, mask, value))


def norm2(a: np.ndarray) -> float:
    n = int(math.log2(a.size))
    return float(lib().orc_norm2(_ptr(a), n))


def sample(a: np.ndarray, u: np.ndarray, mode: str = "exact") -> np.ndarray:
    """Outcome index per draw (peek_measure_all / profile_measure_all rule)."""
    n = int(math.log2(a.size))
    u = np.ascontiguousarray(u, dtype=np.float32)
    out = np.empty(u.size, dtype=np.uint64)
    rc = lib().orc_sample(_ptr(a), n, _ptr(u), u.size, _ptr(out), 0 if mode == "exact" else 1)
    if rc != 0:
        raise MemoryError("orc_sample")
    return out


def histogram(a: np.ndarray, u: np.ndarray, mode: str = "exact") -> np.ndarray:
    """profile_measure_all (src/quantum.cpp:467-501): counts per basis state."""
    idx = sample(a, u, mode)
    return np.bincount(idx.astype(np.int64), minlength=a.size).astype(np.uint32)


def measure(a: np.ndarray, qubit: int, u: float) -> bool:
    """measure(qubit) (src/quantum.cpp:308-342) with the draw passed in."""
    n = int(math.log2(a.size))
    p1 = np.float32(qubit_prob1(a, qubit))
    outcome = bool(np.float32(u) < p1)
    p = p1 if outcome else np.float32(np.float32(1.0) - p1)
    lib().orc_collapse_qubit(_ptr(a), n, qubit, int(outcome), float(p))
    return outcome


def measure_all(a: np.ndarray, u: float) -> int:
    """measure_all (src/quantum.cpp:361-370)."""
    k = int(sample(a, np.array([u], dtype=np.float32))[0])
    a[:] = 0
    a[k] = 1.0
    return k


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, dtype=np.complex64)
    b = np.ascontiguousarray(b, dtype=np.complex64)
    assert a.size == b.size
    return float(lib().orc_rel_l2(_ptr(a), _ptr(b), a.size))


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ---------------------------------------------------------------------------
# Composite-gate builders (src/quantum_gates.cpp) and algorithms
# (src/quantum_algo.cpp:16-129), restated on Circ.
# ---------------------------------------------------------------------------
def single(name: str, *args) -> Circ:
    """X::gate(), Z::gate(), RotX::gate(angle) ... : a 1-qubit circuit."""
    return Circ(1, [(name, 0) + tuple(args)])


def group_gate(qubits: int, targets: Sequence[int], gate: Circ) -> Circ:
    """Group_Gate, src/quantum_gates.cpp:16-34."""
    if gate.n != 1:
        raise ValueError("Gate not supported")
    qc = Circ(qubits)
    for t in sorted(targets):
        qc.add("Gate", gate, t)
    return qc


def control_group_gate(qubits: int, control: int, targets: Sequence[int], gate: Circ) -> Circ:
    """Control_Group_Gate, src/quantum_gates.cpp:36-78."""
    if control >= qubits:
        raise ValueError("Invalid control qubit position")
    if gate.n != 1:
        raise ValueError("Gate not supported")
    ts = sorted(targets)
    if control in ts:
        raise ValueError("Cannot add control gate at the target qubit positions")
    top = [t for t in ts if t < control]
    bottom = [t for t in ts if t > control]
    qc = Circ(qubits)
    # NB the reference places Gate(gate, rank-in-list) inside a |list|-qubit
    # circuit and anchors it at the first target (:59-73): targets must be
    # contiguous within each side for that to mean what it says.
    if top:
        tmp = Circ(len(top))
        for i, _ in enumerate(top):
            tmp.add("Gate", gate, i)
        qc.add("ControlGate", tmp, control, ts[0])
    if bottom:
        tmp = Circ(len(bottom))
        for i, _ in enumerate(bottom):
            tmp.add("Gate", gate, i)
        qc.add("ControlGate", tmp, control, bottom[0])
    return qc


def ncontrol_gate_range(qubits: int, cbegin: int, ccount: int, tbegin: int, gate: Circ) -> Circ:
    """NControl_Gate (contiguous controls), src/quantum_gates.cpp:80-114."""
    if gate.n >= qubits:
        raise ValueError("Gate not supported")
    if ccount == 0:
        raise ValueError("The number of control qubits must be at least 1")
    if tbegin + gate.n > qubits:
        raise ValueError("Invalid target qubit_begin position")
    if cbegin + ccount > tbegin:
        raise ValueError("Invalid control_qubit position")
    qc = Circ(qubits)
    if ccount == 1:
        qc.add("ControlGate", gate, cbegin, tbegin)
    else:
        temp = Circ(gate.n + 1 + tbegin - cbegin - ccount)
        temp.add("ControlGate", gate, 0, tbegin - cbegin - ccount + 1)
        for _ in range(ccount - 1):
            tmp = Circ(temp.n + 1)
            tmp.add("ControlGate", temp, 0, 1)
            temp = tmp
        qc.add("Gate", temp, cbegin)
    return qc


def ncontrol_gate_list(qubits: int, controls: Sequence[int], tbegin: int, gate: Circ) -> Circ:
    """NControl_Gate (control list), src/quantum_gates.cpp:116-182."""
    if len(controls) == 0:
        raise ValueError("Number of control qubits must be at least one")
    if gate.n + len(controls) > qubits:
        raise ValueError("Invalid number of qubits")
    if tbegin + gate.n > qubits:
        raise ValueError("Invalid target qubit begin position")
    cs = sorted(controls)
    if cs[-1] >= qubits:
        raise ValueError("Cannot add control gate at the given position")
    top = [c for c in cs if c < tbegin]
    bottom = [c for c in cs if c >= tbegin]
    if bottom and bottom[0] < tbegin + gate.n:
        raise ValueError("Cannot add control gate at the target qubit positions")
    current = gate
    if bottom:
        for c in bottom[:-1]:
            temp = Circ(c - tbegin + 1)
            temp.add("ControlGate", current, c - tbegin, 0)
            current = temp
        temp = Circ(qubits - tbegin)
        temp.add("ControlGate", current, bottom[-1] - tbegin, 0)
        current = temp
    if top:
        prev = tbegin
        for c in list(reversed(top))[:-1]:
            temp = Circ(qubits - c)
            temp.add("ControlGate", current, 0, prev - c)
            prev = c
            current = temp
        temp = Circ(qubits)
        temp.add("ControlGate", current, top[0], qubits - current.n)
        current = temp
    return current


def rewire_gate(qubits: int, new_pos: Sequence[int], gate: Circ) -> Circ:
    """Rewire_Gate, src/quantum_gates.cpp:184-231."""
    if len(new_pos) != gate.n:
        raise ValueError("New qubit positions must map all the qubits in the gate")
    if gate.n > qubits:
        raise ValueError("Cannot rewire circuit to a lower number of qubits")
    if len(set(new_pos)) != len(new_pos):
        raise ValueError("Cannot rewire multiple qubits to the same qubit")
    qc = Circ(qubits)
    swapped = set()
    swaps = []
    for i in range(len(new_pos)):
        if i in swapped:
            continue
        swapped.add(i)
        cur = i
        while i != new_pos[cur]:
            swaps.append(("Swap", cur, new_pos[cur]))
            cur = new_pos[cur]
            swapped.add(cur)
    qc.extend(swaps)
    qc.add("Gate", gate, 0)
    qc.extend(reversed(swaps))
    return qc


def adjoint_gate(gate: Circ) -> Circ:
    """Adjoint_Gate, src/quantum_gates.cpp:233-338 (without its aliasing bug:
    the reference negates angles on gate objects shared with the source)."""
    out = Circ(gate.n)
    for g in reversed(gate.gates):
        name = ALIASES.get(g[0], g[0])
        if name in HAS_ANGLE:
            out.add(*g[:-1], -float(np.float32(g[-1])))
        elif name == "Gate":
            out.add("Gate", adjoint_gate(g[1]), g[2])
        elif name == "ControlGate":
            out.add("ControlGate", adjoint_gate(g[1]), g[2], g[3])
        else:
            out.add(*g)
    return out


def fourier_transform(qubits: int) -> Circ:
    """src/quantum_algo.cpp:103-114 (no final swaps)."""
    qc = Circ(qubits)
    for i in range(qubits - 1, -1, -1):
        qc.add("H", i)
        for j in range(i):
            qc.add("CPhase", j, i, float(np.float32(PI_F32) / np.float32(1 << (i - j))))
    return qc


def inverse_fourier_transform(qubits: int) -> Circ:
    """src/quantum_algo.cpp:116-129."""
    qc = Circ(qubits)
    for i in range(qubits):
        for j in range(i - 1, -1, -1):
            qc.add("CPhase", j, i, float(-np.float32(PI_F32) / np.float32(1 << (i - j))))
        qc.add("H", i)
    return qc


def grover_oracle(search_qubits: int, marked_state: int) -> Circ:
    """src/quantum_algo.cpp:16-40: bit i of marked_state <-> qubit i."""
    if marked_state >= (1 << search_qubits):
        raise ValueError("Marked state should be in the range [0, 2^search_qubits)")
    qc = Circ(search_qubits)
    for i in range(search_qubits):
        if not (marked_state & (1 << i)):
            qc.add("X", i)
    qc.add("Gate", ncontrol_gate_range(search_qubits, 0, search_qubits - 1, search_qubits - 1, single("Z")), 0)
    for i in range(search_qubits):
        if not (marked_state & (1 << i)):
            qc.add("X", i)
    return qc


def grover_search(search_qubits: int, oracle: Circ, iterations: int) -> Circ:
    """src/quantum_algo.cpp:42-77."""
    if oracle.n < search_qubits:
        raise ValueError("Cannot use given oracle for this qubit circuit")
    qc = Circ(oracle.n)
    for i in range(search_qubits):
        qc.add("H", i)
    mcz = ncontrol_gate_range(search_qubits, 0, search_qubits - 1, search_qubits - 1, single("Z"))
    for _ in range(iterations):
        qc.add("Barrier")
        qc.add("Gate", oracle, 0)
        qc.add("Barrier")
        for j in range(search_qubits):
            qc.add("H", j)
        for j in range(search_qubits):
            qc.add("X", j)
        qc.add("Gate", mcz, 0)
        for j in range(search_qubits):
            qc.add("X", j)
            qc.add("H", j)
    return qc


def grover_iteration(search_qubits: int, oracle: Circ, iterations: int) -> Circ:
    """src/quantum_algo.cpp:79-101."""
    qc = Circ(search_qubits)
    mcz = ncontrol_gate_range(search_qubits, 0, search_qubits - 1, search_qubits - 1, single("Z"))
    for _ in range(iterations):
        qc.add("Gate", oracle, 0)
        for j in range(search_qubits):
            qc.add("H", j)
        for j in range(search_qubits):
            qc.add("X", j)
        qc.add("Gate", mcz, 0)
        for j in range(search_qubits):
            qc.add("X", j)
            qc.add("H", j)
    return qc


def count_primitives(circ: Circ) -> int:
    """Gate applications after flattening composites (SURVEY §8d unit of work)."""
    k = 0
    for g in circ.gates:
        name = ALIASES.get(g[0], g[0])
        if name == "Barrier":
            continue
        if name in ("Gate", "ControlGate"):
            k += count_primitives(g[1])
        else:
            k += 1
    return k

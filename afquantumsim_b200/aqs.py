"""Python mirror of the reference's ``aqs::`` API.

Every class here is a thin handle on the C++14 host layer
(afquantumsim_b200/host -> lib/libafquantum.so), which in turn calls the CUDA
engine through include/aqs_engine.h.  Names, argument order and error classes
follow the reference (include/quantum.h, quantum_gates.h, quantum_algo.h):

    qc = QCircuit(2); qc << H(0) << CX(0, 1)
    qs = QSimulator(2); qs.simulate(qc); qs.profile_measure_all(1000)

Reference exception classes map to: std::out_of_range -> OutOfRange(IndexError),
std::invalid_argument -> InvalidArgument(ValueError), std::domain_error ->
DomainError(ArithmeticError), anything else -> EngineFailure(RuntimeError).
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import engine as _engine

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libafquantum.so")
PI = float(np.float32(3.14159265358979323846))


class OutOfRange(IndexError):
    pass


class InvalidArgument(ValueError):
    pass


class DomainError(ArithmeticError):
    pass


class EngineFailure(RuntimeError):
    pass


_ERR = {1: OutOfRange, 2: InvalidArgument, 3: DomainError, 4: EngineFailure, 5: EngineFailure}
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise EngineFailure(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        _engine.load()   # lib/libaqs_engine.so, by absolute path
        L = ctypes.CDLL(HOST_LIB_PATH)
        L.aqsh_last_error.restype = ctypes.c_char_p
        L.aqsh_circuit_qubits.restype = ctypes.c_uint32
        L.aqsh_circuit_gate_count.restype = ctypes.c_uint64
        L.aqsh_circuit_cached_index.restype = ctypes.c_uint64
        L.aqsh_sim_qubits.restype = ctypes.c_uint32
        L.aqsh_sim_engine_handle.restype = ctypes.c_void_p
        L.aqsh_set_seed.argtypes = [ctypes.c_uint64]
        _lib = L
    return _lib


def _ck(rc: int):
    if rc:
        raise _ERR.get(rc, EngineFailure)(_load().aqsh_last_error().decode())


def _u32s(xs: Iterable[int]):
    xs = [int(x) for x in xs]
    return (ctypes.c_uint32 * len(xs))(*xs), len(xs)


_initialized = False


def initialize(device: int = 0) -> None:
    """aqs::initialize (reference src/quantum.cpp:69-86)."""
    global _initialized
    _ck(_load().aqsh_initialize(int(device)))
    _initialized = True
    _engine._inited = True


def _ensure():
    if not _initialized:
        initialize(int(os.environ.get("LOCAL_RANK", "0")))


def set_seed(seed: int) -> None:
    _load().aqsh_set_seed(seed)


def set_fusion(on: bool) -> None:
    _load().aqsh_set_fusion(1 if on else 0)


def set_jit_min_qubits(n: int) -> None:
    """Specialised pass kernels are requested for circuits on at least n qubits (default 26)."""
    _load().aqsh_set_jit_min_qubits(int(n))


def jit_wait() -> None:
    """Block until no kernel compilation is pending."""
    _load().aqsh_jit_wait()


def get_fusion() -> bool:
    return bool(_load().aqsh_get_fusion())


def clear_circuit_cache() -> None:
    _load().aqsh_clear_circuit_cache()


# ---------------------------------------------------------------------------
# gates: plain value objects with the reference's field names
# ---------------------------------------------------------------------------
class QGate:
    name = ""
    fields: Sequence[str] = ()
    has_angle = False

    def __init__(self, *args):
        want = len(self.fields) + (1 if self.has_angle else 0)
        if len(args) != want:
            raise TypeError(f"{type(self).__name__} takes {want} arguments")
        for f, v in zip(self.fields, args):
            setattr(self, f, int(v))
        self.angle = float(np.float32(args[-1])) if self.has_angle else 0.0

    def qubits(self) -> List[int]:
        return [getattr(self, f) for f in self.fields]

    def _append_to(self, qc: "QCircuit"):
        arr, n = _u32s(self.qubits())
        _ck(_load().aqsh_circuit_add(qc._h, self.name.encode(), arr, n, ctypes.c_float(self.angle)))


def _gate(cls_name, fields, angle=False):
    return type(cls_name, (QGate,), {"name": cls_name, "fields": tuple(fields), "has_angle": angle})


X = _gate("X", ["target_qubit"])
Y = _gate("Y", ["target_qubit"])
Z = _gate("Z", ["target_qubit"])
H = _gate("H", ["target_qubit"])
Phase = _gate("Phase", ["target_qubit"], True)
RotX = _gate("RotX", ["target_qubit"], True)
RotY = _gate("RotY", ["target_qubit"], True)
RotZ = _gate("RotZ", ["target_qubit"], True)
Swap = _gate("Swap", ["target_qubit_A", "target_qubit_B"])
CX = _gate("CX", ["control_qubit", "target_qubit"])
CY = _gate("CY", ["control_qubit", "target_qubit"])
CZ = _gate("CZ", ["control_qubit", "target_qubit"])
CH = _gate("CH", ["control_qubit", "target_qubit"])
CPhase = _gate("CPhase", ["control_qubit", "target_qubit"], True)
CRotX = _gate("CRotX", ["control_qubit", "target_qubit"], True)
CRotY = _gate("CRotY", ["control_qubit", "target_qubit"], True)
CRotZ = _gate("CRotZ", ["control_qubit", "target_qubit"], True)
CSwap = _gate("CSwap", ["control_qubit", "target_qubit_A", "target_qubit_B"])
CCNot = _gate("CCNot", ["control_qubit_A", "control_qubit_B", "target_qubit"])
Or = _gate("Or", ["control_qubit_A", "control_qubit_B", "target_qubit"])
Not, CNot, Xor, And = X, CX, CX, CCNot
GATE_CLASSES = {c.name: c for c in (X, Y, Z, H, Phase, RotX, RotY, RotZ, Swap, CX, CY, CZ, CH, CPhase, CRotX, CRotY,
                                    CRotZ, CSwap, CCNot, Or)}
GATE_CLASSES.update({"Not": X, "CNot": CX, "Xor": CX, "And": CCNot})


class Barrier(QGate):
    name = "Barrier"

    def __init__(self, visible: bool = True):
        self.visible = bool(visible)

    def _append_to(self, qc):
        arr, n = _u32s([1 if self.visible else 0])
        _ck(_load().aqsh_circuit_add(qc._h, b"Barrier", arr, n, ctypes.c_float(0.0)))


class Gate(QGate):
    """A circuit used as a gate at target_qubit_begin (reference include/quantum.h:1497-1541)."""

    def __init__(self, circuit: "QCircuit", target_qubit_begin: int, name: str = ""):
        self.circuit, self.target_qubit_begin, self.gate_name = circuit, int(target_qubit_begin), name

    def _append_to(self, qc):
        _ck(_load().aqsh_circuit_add_gate(qc._h, self.circuit._h, self.target_qubit_begin, self.gate_name.encode()))


class ControlGate(QGate):
    def __init__(self, circuit: "QCircuit", control_qubit: int, target_qubit_begin: int, name: str = ""):
        self.circuit, self.control_qubit = circuit, int(control_qubit)
        self.target_qubit_begin, self.gate_name = int(target_qubit_begin), name

    def _append_to(self, qc):
        _ck(_load().aqsh_circuit_add_control_gate(qc._h, self.circuit._h, self.control_qubit, self.target_qubit_begin,
                                                  self.gate_name.encode()))


# ---------------------------------------------------------------------------
class QCircuit:
    def __init__(self, qubit_count: int, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            self._h = ctypes.c_void_p()
            _ck(_load().aqsh_circuit_new(int(qubit_count), ctypes.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _load().aqsh_circuit_free(self._h)
        except Exception:
            pass

    def __lshift__(self, gate):
        if isinstance(gate, (list, tuple)) and gate and isinstance(gate[0], QGate):
            for g in gate:
                g._append_to(self)
        else:
            gate._append_to(self)
        return self

    def set_matrix(self, matrix) -> "QCircuit":
        """Make this gate-less circuit (at most 6 qubits) the opaque matrix `matrix` (2^n x 2^n, matrix[r, c]): what the
        reference's `qc.circuit() = M` does (docs/USAGE.md:121-125).  Gate / ControlGate of it apply the matrix through the
        engine's dense-matrix kernel (tensor cores from 5 qubits on)."""
        m = np.ascontiguousarray(np.asarray(matrix, dtype=np.complex64))
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise InvalidArgument("set_matrix: a square matrix is required")
        _ck(_load().aqsh_circuit_set_matrix(self._h, m.ctypes.data_as(ctypes.c_void_p), int(m.shape[0])))
        return self

    def copy(self) -> "QCircuit":
        h = ctypes.c_void_p()
        _ck(_load().aqsh_circuit_copy(self._h, ctypes.byref(h)))
        return QCircuit(0, _handle=h)

    def qubit_count(self) -> int:
        return _load().aqsh_circuit_qubits(self._h)

    def state_count(self) -> int:
        return 1 << self.qubit_count()

    def gate_count(self) -> int:
        return _load().aqsh_circuit_gate_count(self._h)

    def cached_index(self) -> int:
        return _load().aqsh_circuit_cached_index(self._h)

    def compile(self):
        _ensure()
        _ck(_load().aqsh_circuit_compile(self._h))

    def clear(self):
        _ck(_load().aqsh_circuit_clear(self._h))

    def clear_cache(self):
        _ck(_load().aqsh_circuit_clear_cache(self._h))

    def representation(self) -> str:
        need = ctypes.c_size_t()
        _ck(_load().aqsh_circuit_representation(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _ck(_load().aqsh_circuit_representation(self._h, buf, need.value, None))
        return buf.value.decode()

    def ops(self) -> np.ndarray:
        """The primitive ops the whole gate list lowers to (struct aqs_op records)."""
        cnt = ctypes.c_uint64()
        _ck(_load().aqsh_circuit_ops(self._h, None, 0, ctypes.byref(cnt)))
        out = np.zeros(cnt.value, dtype=_engine.OP_DTYPE)
        _ck(_load().aqsh_circuit_ops(self._h, out.ctypes.data_as(ctypes.c_void_p), cnt.value, ctypes.byref(cnt)))
        return out

    def circuit(self) -> np.ndarray:
        """Dense matrix of the compiled prefix, U[row, col] (materialised on the device, n <= 13)."""
        _ensure()
        n = self.state_count()
        buf = np.empty(n * n, dtype=np.complex64)
        _ck(_load().aqsh_circuit_matrix(self._h, buf.ctypes.data_as(ctypes.c_void_p)))
        return buf.reshape(n, n).T     # column-major -> [row, col]

    def add(self, *gate) -> "QCircuit":
        """Append one gate given as a plain tuple ("Name", *ctor_args) (workloads.py format)."""
        name, args = gate[0], gate[1:]
        if name == "Barrier":
            return self << Barrier(*args)
        if name == "Gate":
            return self << Gate(*args)
        if name == "ControlGate":
            return self << ControlGate(*args)
        return self << GATE_CLASSES[name](*args)

    def extend(self, gates) -> "QCircuit":
        for g in gates:
            self.add(*g)
        return self


def _new_circuit(fn, *args) -> QCircuit:
    h = ctypes.c_void_p()
    _ck(fn(*args, ctypes.byref(h)))
    return QCircuit(0, _handle=h)


def single(name: str, *angle) -> QCircuit:
    """X::gate(), RotX::gate(angle), ...: a compiled 1-qubit circuit."""
    qc = QCircuit(1)
    qc.add(name, 0, *angle)
    return qc


def Group_Gate(qubits, target_qubits, gate: QCircuit) -> QCircuit:
    arr, n = _u32s(target_qubits)
    return _new_circuit(_load().aqsh_group_gate, int(qubits), arr, n, gate._h)


def Control_Group_Gate(qubits, control_qubit, target_qubits, gate: QCircuit) -> QCircuit:
    arr, n = _u32s(target_qubits)
    return _new_circuit(_load().aqsh_control_group_gate, int(qubits), int(control_qubit), arr, n, gate._h)


def NControl_Gate(qubits, controls, *rest) -> QCircuit:
    """Both reference overloads: (qubits, cbegin, ccount, tbegin, gate) or (qubits, [controls], tbegin, gate)."""
    if isinstance(controls, (list, tuple, np.ndarray)):
        tbegin, gate = rest
        arr, n = _u32s(controls)
        return _new_circuit(_load().aqsh_ncontrol_gate_list, int(qubits), arr, n, int(tbegin), gate._h)
    ccount, tbegin, gate = rest
    return _new_circuit(_load().aqsh_ncontrol_gate_range, int(qubits), int(controls), int(ccount), int(tbegin), gate._h)


def Rewire_Gate(qubits, new_qubit_positions, gate: QCircuit) -> QCircuit:
    arr, n = _u32s(new_qubit_positions)
    return _new_circuit(_load().aqsh_rewire_gate, int(qubits), arr, n, gate._h)


def Adjoint_Gate(gate: QCircuit) -> QCircuit:
    _ensure()
    return _new_circuit(_load().aqsh_adjoint_gate, gate._h)


def fourier_transform(qubits: int) -> QCircuit:
    return _new_circuit(_load().aqsh_fourier_transform, int(qubits), 0)


def inverse_fourier_transform(qubits: int) -> QCircuit:
    return _new_circuit(_load().aqsh_fourier_transform, int(qubits), 1)


def grover_oracle(search_qubits: int, marked_state: int) -> QCircuit:
    return _new_circuit(_load().aqsh_grover_oracle, int(search_qubits), int(marked_state))


def grover_search(search_qubits: int, oracle: QCircuit, iterations: int, oracle_name: str = "") -> QCircuit:
    return _new_circuit(_load().aqsh_grover_search, int(search_qubits), oracle._h, int(iterations), oracle_name.encode())


def grover_iteration(search_qubits: int, oracle: QCircuit, iterations: int) -> QCircuit:
    return _new_circuit(_load().aqsh_grover_iteration, int(search_qubits), oracle._h, int(iterations))


def gen_circuit_text_image(circuit_or_schematic, simulator: Optional["QSimulator"] = None) -> str:
    need = ctypes.c_size_t()
    L = _load()
    if isinstance(circuit_or_schematic, str):
        s = circuit_or_schematic.encode()
        _ck(L.aqsh_schematic_text_image(s, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _ck(L.aqsh_schematic_text_image(s, buf, need.value, None))
    else:
        _ck(L.aqsh_circuit_text_image(circuit_or_schematic._h, simulator._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _ck(L.aqsh_circuit_text_image(circuit_or_schematic._h, simulator._h, buf, need.value, None))
    return buf.value.decode()


# ---------------------------------------------------------------------------
class QState:
    """One qubit (reference include/quantum.h:151-342); normalised on construction in f32."""

    def __init__(self, zero_state: complex = 1.0, one_state: complex = 0.0):
        z, o = np.complex64(zero_state), np.complex64(one_state)
        f = np.float32
        # ((zr*zr + zi*zi) + or*or) + oi*oi, every step rounded to f32 (src/quantum.cpp:145-157)
        mag2 = f(f(f(f(z.real) * f(z.real)) + f(f(z.imag) * f(z.imag))) + f(f(o.real) * f(o.real)))
        mag2 = f(mag2 + f(f(o.imag) * f(o.imag)))
        if f(mag2) == 0:
            raise InvalidArgument("Cannot normalize a null state")
        mag = f(np.sqrt(f(mag2)))
        den = f(mag * mag)   # complex division by (mag, 0): x*mag/(mag*mag), as the C++ host layer does

        def div(x):
            return f(f(f(x) * mag) / den)
        self.state = (np.complex64(complex(div(z.real), div(z.imag))), np.complex64(complex(div(o.real), div(o.imag))))

    def __getitem__(self, i):
        return self.state[int(i)]

    @staticmethod
    def zero():
        return QState(1.0, 0.0)

    @staticmethod
    def one():
        return QState(0.0, 1.0)

    @staticmethod
    def plus():
        return QState(0.70710678118, 0.70710678118)

    @staticmethod
    def minus():
        return QState(0.70710678118, -0.70710678118)

    def probability_true(self) -> float:
        o = self.state[1]
        return float(np.float32(o.real) * np.float32(o.real) + np.float32(o.imag) * np.float32(o.imag))


def _qstate_raw(q) -> List[float]:
    if isinstance(q, QState):
        z, o = q.state
    else:
        z, o = np.complex64(q[0]), np.complex64(q[1])
    return [float(z.real), float(z.imag), float(o.real), float(o.imag)]


class QSimulator:
    Z, Y, X = 0, 1, 2   # Basis

    def __init__(self, qubit_count: int, initial=None, _handle=None):
        _ensure()
        L = _load()
        self._h = ctypes.c_void_p()
        n = int(qubit_count)
        if _handle is not None:
            self._h = _handle
        elif initial is None:
            _ck(L.aqsh_sim_new(n, ctypes.byref(self._h)))
        elif isinstance(initial, np.ndarray):
            v = np.ascontiguousarray(initial, dtype=np.complex64)
            if v.size != (1 << n):
                raise InvalidArgument("Invalid initial statevector shape")
            _ck(L.aqsh_sim_new_vector(n, v.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)))
        else:
            states = [initial] * n if isinstance(initial, QState) else list(initial)
            if len(states) != n:
                raise InvalidArgument("The number of initial states must match the number of qubits in the circuit")
            raw = np.asarray([_qstate_raw(q) for q in states], dtype=np.float32)
            _ck(L.aqsh_sim_new_states(n, raw.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _load().aqsh_sim_free(self._h)
        except Exception:
            pass

    def clone(self) -> "QSimulator":
        h = ctypes.c_void_p()
        _ck(_load().aqsh_sim_clone(self._h, ctypes.byref(h)))
        return QSimulator(self.qubit_count(), _handle=h)

    def qubit_count(self) -> int:
        return _load().aqsh_sim_qubits(self._h)

    def state_count(self) -> int:
        return 1 << self.qubit_count()

    def set_qubit(self, index: int, q) -> None:
        """qs.qubit(i) = state"""
        raw = (ctypes.c_float * 4)(*_qstate_raw(q))
        _ck(_load().aqsh_sim_set_qubit(self._h, int(index), raw))

    def qubit(self, index: int) -> QState:
        raw = (ctypes.c_float * 4)()
        _ck(_load().aqsh_sim_get_qubit(self._h, int(index), raw))
        q = QState.__new__(QState)
        q.state = (np.complex64(complex(raw[0], raw[1])), np.complex64(complex(raw[2], raw[3])))
        return q

    def generate_statevector(self):
        _ck(_load().aqsh_sim_generate_statevector(self._h))

    def simulate(self, circuit: QCircuit):
        _ck(_load().aqsh_sim_simulate(self._h, circuit._h))

    def peek_measure(self, qubit: int) -> bool:
        out = ctypes.c_int()
        _ck(_load().aqsh_sim_peek_measure(self._h, int(qubit), ctypes.byref(out)))
        return bool(out.value)

    def measure(self, qubit: int) -> bool:
        out = ctypes.c_int()
        _ck(_load().aqsh_sim_measure(self._h, int(qubit), ctypes.byref(out)))
        return bool(out.value)

    def measure_all(self) -> int:
        out = ctypes.c_uint32()
        _ck(_load().aqsh_sim_measure_all(self._h, ctypes.byref(out)))
        return out.value

    def peek_measure_all(self) -> int:
        out = ctypes.c_uint32()
        _ck(_load().aqsh_sim_peek_measure_all(self._h, ctypes.byref(out)))
        return out.value

    def profile_measure(self, qubit: int, rep_count: int):
        out = (ctypes.c_uint32 * 2)()
        _ck(_load().aqsh_sim_profile_measure(self._h, int(qubit), int(rep_count), out))
        return [out[0], out[1]]

    def profile_measure_all(self, rep_count: int) -> np.ndarray:
        out = np.empty(self.state_count(), dtype=np.uint32)
        _ck(_load().aqsh_sim_profile_measure_all(self._h, int(rep_count), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def sample(self, draws: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(draws, dtype=np.float32)
        out = np.empty(u.size, dtype=np.uint64)
        _ck(_load().aqsh_sim_sample(self._h, u.ctypes.data_as(ctypes.c_void_p), u.size, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def qubit_probability_true(self, qubit: int) -> float:
        out = ctypes.c_float()
        _ck(_load().aqsh_sim_qubit_probability_true(self._h, int(qubit), ctypes.byref(out)))
        return out.value

    def qubit_probability_false(self, qubit: int) -> float:
        return float(np.float32(1.0) - np.float32(self.qubit_probability_true(qubit)))

    def state_probability(self, state: int) -> float:
        out = ctypes.c_float()
        _ck(_load().aqsh_sim_state_probability(self._h, int(state), ctypes.byref(out)))
        return out.value

    def probabilities(self) -> np.ndarray:
        out = np.empty(self.state_count(), dtype=np.float32)
        _ck(_load().aqsh_sim_probabilities(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def state(self, index: int) -> np.complex64:
        out = (ctypes.c_float * 2)()
        _ck(_load().aqsh_sim_state(self._h, int(index), out))
        return np.complex64(complex(out[0], out[1]))

    def statevector(self) -> np.ndarray:
        out = np.empty(self.state_count(), dtype=np.complex64)
        _ck(_load().aqsh_sim_statevector(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def set_basis(self, basis: int):
        _ck(_load().aqsh_sim_set_basis(self._h, int(basis)))

    def get_basis(self) -> int:
        return _load().aqsh_sim_get_basis(self._h)

    def norm2(self) -> float:
        out = ctypes.c_double()
        _ck(_load().aqsh_sim_norm2(self._h, ctypes.byref(out)))
        return out.value

    def sync(self):
        _ck(_load().aqsh_sim_sync(self._h))

    def engine_state(self) -> "_engine.State":
        """Borrowed engine.State view of this simulator's device state (not owned)."""
        st = _engine.State.__new__(_engine.State)
        st.n, st.size = self.qubit_count(), self.state_count()
        st._h = ctypes.c_void_p(_load().aqsh_sim_engine_handle(self._h))
        st.close = lambda: None
        return st

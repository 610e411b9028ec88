"""ctypes binding of the engine C ABI (include/aqs_engine.h).

Thin by design: every method is one C call.  There is no CPU fallback — if
libaqs_engine.so is missing, or no sm_100 device is visible, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AQS_ENGINE_LIB") or os.path.join(_HERE, "lib", "libaqs_engine.so")   # the override is for A/B builds of the engine

OP_U2, OP_DIAG, OP_X, OP_SWAP = 0, 1, 2, 3
PLAN_FUSE, PLAN_GRAPH, PLAN_JIT, PLAN_JIT_ASYNC = 1, 2, 4, 8

# struct aqs_op (64 bytes) as a numpy record, so op lists are one contiguous buffer
OP_DTYPE = np.dtype([
    ("kind", "<i4"), ("target", "<i4"), ("target2", "<i4"), ("reserved", "<i4"),
    ("ctrl_mask", "<u8"), ("ctrl_value", "<u8"), ("m", "<f4", (8,)),
], align=True)
assert OP_DTYPE.itemsize == 64


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("n_ops", ctypes.c_uint64), ("n_launches", ctypes.c_uint64),
        ("n_fused_passes", ctypes.c_uint64), ("n_single_ops", ctypes.c_uint64),
        ("bytes_unfused", ctypes.c_double), ("bytes_planned", ctypes.c_double),
        ("n_qubits", ctypes.c_int32), ("tile_bits", ctypes.c_int32),
    ]


class Counters(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("gate_ops", ctypes.c_uint64),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64)]


class JitInfo(ctypes.Structure):
    _fields_ = [("compiled", ctypes.c_uint64), ("cache_hits", ctypes.c_uint64), ("failed", ctypes.c_uint64),
                ("pending", ctypes.c_uint64), ("compile_seconds", ctypes.c_double)]


class EngineError(RuntimeError):
    pass


_lib = None
_inited = False

# every symbol include/aqs_engine.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "aqs_engine_init", "aqs_engine_shutdown", "aqs_engine_device", "aqs_engine_abi_version", "aqs_last_error",
    "aqs_state_create", "aqs_state_wrap", "aqs_state_destroy", "aqs_state_clone", "aqs_state_qubits", "aqs_state_set_basis",
    "aqs_state_set_product", "aqs_state_set_identity", "aqs_state_upload", "aqs_state_download",
    "aqs_state_get_amp", "aqs_state_device_ptr", "aqs_state_set_stream", "aqs_state_get_stream", "aqs_sync",
    "aqs_apply_op", "aqs_apply_ops", "aqs_plan_build", "aqs_plan_run", "aqs_plan_get_info", "aqs_plan_destroy", "aqs_plan_export_pass",
    "aqs_norm2", "aqs_scale", "aqs_prob_fixed", "aqs_qubit_prob1", "aqs_probabilities", "aqs_collapse_qubit",
    "aqs_sample", "aqs_sample_fixed", "aqs_sample_hist", "aqs_timer_create", "aqs_timer_start", "aqs_timer_stop",
    "aqs_timer_elapsed_ms", "aqs_timer_destroy", "aqs_counters_get", "aqs_counters_reset",
    "aqs_state_ipc_export", "aqs_ipc_open", "aqs_ipc_close_all", "aqs_peer_bitswap",
    "aqs_apply_dense", "aqs_flat_create", "aqs_flat_attach", "aqs_flat_ptr", "aqs_flat_destroy", "aqs_plan_run_shard", "aqs_plan_pass_span", "aqs_plan_shard_cut",
    "aqs_plan_pass_source", "aqs_plan_pass_coefs", "aqs_plan_jit_ready", "aqs_jit_wait", "aqs_jit_get_info",
    "aqs_sample_hist_sparse", "aqs_pool_trim",
    "aqs_flat_view_create", "aqs_memcpy_async", "aqs_plan_run_tiles", "aqs_plan_pass_tile",
]


def load():
    """dlopen libaqs_engine.so (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_float
    P = ctypes.POINTER
    sig = {
        "aqs_engine_init": [i32], "aqs_engine_shutdown": [],
        "aqs_engine_device": [P(i32), P(i32), P(ctypes.c_size_t)],
        "aqs_state_create": [i32, P(vp)], "aqs_state_wrap": [i32, vp, P(vp)], "aqs_state_destroy": [vp], "aqs_state_clone": [vp, P(vp)],
        "aqs_state_qubits": [vp, P(i32)], "aqs_state_set_basis": [vp, u64], "aqs_state_set_product": [vp, vp],
        "aqs_state_set_identity": [vp], "aqs_state_upload": [vp, vp, u64, u64],
        "aqs_state_download": [vp, vp, u64, u64], "aqs_state_get_amp": [vp, u64, vp],
        "aqs_state_device_ptr": [vp, P(vp)], "aqs_state_set_stream": [vp, vp], "aqs_state_get_stream": [vp, P(vp)],
        "aqs_sync": [vp], "aqs_apply_op": [vp, vp], "aqs_apply_ops": [vp, vp, u64],
        "aqs_plan_build": [i32, vp, u64, ctypes.c_uint32, P(vp)], "aqs_plan_run": [vp, vp],
        "aqs_plan_get_info": [vp, P(PlanInfo)], "aqs_plan_destroy": [vp],
        "aqs_plan_export_pass": [vp, u64, vp, u64, P(u64)],
        "aqs_norm2": [vp, P(ctypes.c_double)], "aqs_scale": [vp, f32],
        "aqs_prob_fixed": [vp, u64, u64, P(u64)], "aqs_qubit_prob1": [vp, i32, P(ctypes.c_double)],
        "aqs_probabilities": [vp, vp, u64, u64], "aqs_collapse_qubit": [vp, i32, i32, f32],
        "aqs_sample": [vp, vp, u64, vp], "aqs_sample_fixed": [vp, vp, u64, vp], "aqs_sample_hist": [vp, vp, u64, vp],
        "aqs_timer_create": [P(vp)], "aqs_timer_start": [vp, vp], "aqs_timer_stop": [vp, vp],
        "aqs_timer_elapsed_ms": [vp, P(ctypes.c_double)], "aqs_timer_destroy": [vp],
        "aqs_counters_get": [P(Counters)], "aqs_counters_reset": [],
        "aqs_state_ipc_export": [vp, vp], "aqs_ipc_open": [vp, P(vp)], "aqs_ipc_close_all": [],
        "aqs_peer_bitswap": [vp, P(vp), i32, P(i32), ctypes.c_uint32],
        "aqs_apply_dense": [vp, P(i32), i32, u64, u64, vp],
        "aqs_flat_create": [u64, i32, i32, P(vp), P(i32)], "aqs_flat_attach": [vp, i32, i32],
        "aqs_flat_ptr": [vp, P(vp), P(vp)], "aqs_flat_destroy": [vp],
        "aqs_plan_run_shard": [vp, vp, u64, u64, i32, i32], "aqs_plan_pass_span": [vp, u64, i32, P(i32)],
        "aqs_plan_shard_cut": [vp, u64, i32, i32, P(ctypes.c_uint32), P(ctypes.c_uint32), P(ctypes.c_uint8)],
        "aqs_plan_pass_source": [vp, u64, vp, u64, P(u64), P(u64)], "aqs_plan_pass_coefs": [vp, u64, vp, u64, P(u64)],
        "aqs_plan_jit_ready": [vp, P(u64)], "aqs_jit_wait": [], "aqs_jit_get_info": [P(JitInfo)],
        "aqs_sample_hist_sparse": [vp, vp, u64, vp, vp, u64, P(u64)], "aqs_pool_trim": [],
        "aqs_plan_pass_tile": [vp, u64, P(ctypes.c_uint8), P(i32)],
        "aqs_flat_view_create": [vp, vp, u64, P(vp)], "aqs_memcpy_async": [vp, vp, u64, vp],
        "aqs_plan_run_tiles": [vp, vp, u64, vp, ctypes.c_uint32, P(ctypes.c_uint8), ctypes.c_uint32, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.aqs_last_error.restype = ctypes.c_char_p
    L.aqs_engine_abi_version.restype = i32
    _lib = L
    return L


IPC_HANDLE_BYTES = 64


def ipc_open(handle: bytes) -> int:
    """Map another process's exported state into this one; returns the device pointer (cached by the engine)."""
    p = ctypes.c_void_p()
    _check(load().aqs_ipc_open(handle, ctypes.byref(p)))
    return p.value


def _check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"aqs engine error {rc}: {load().aqs_last_error().decode()}")


def init(device: int = 0) -> None:
    """aqs::initialize (src/quantum.cpp:69-86)."""
    global _inited
    _check(load().aqs_engine_init(device))
    _inited = True


def ensure_init(device: Optional[int] = None) -> None:
    if not _inited:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        init(device)


def device_info():
    d, sm, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    _check(load().aqs_engine_device(ctypes.byref(d), ctypes.byref(sm), ctypes.byref(mem)))
    return {"device": d.value, "sm_count": sm.value, "hbm_bytes": mem.value}


def counters() -> dict:
    c = Counters()
    _check(load().aqs_counters_get(ctypes.byref(c)))
    return {k: int(getattr(c, k)) for k, _ in Counters._fields_}


def counters_reset() -> None:
    _check(load().aqs_counters_reset())


# ---------------------------------------------------------------------------
# op construction helpers
# ---------------------------------------------------------------------------
def qmask(qubits: Iterable[int]) -> int:
    m = 0
    for q in qubits:
        m |= 1 << int(q)
    return m


def make_ops(n: int) -> np.ndarray:
    ops = np.zeros(n, dtype=OP_DTYPE)
    ops["target2"] = -1
    return ops


def op_record(kind: int, target: int, m: Sequence[complex] = (1, 0, 0, 1), controls: Iterable[int] = (),
              target2: int = -1, ctrl_value: Optional[int] = None) -> np.ndarray:
    r = make_ops(1)
    r["kind"], r["target"], r["target2"] = kind, target, target2
    cm = qmask(controls)
    r["ctrl_mask"] = cm
    r["ctrl_value"] = cm if ctrl_value is None else ctrl_value
    mm = np.asarray(m, dtype=np.complex64).reshape(4)
    r["m"][0] = mm.view(np.float32)
    return r


class FlatSpace:
    """aqs_flat_t: the shards of every GPU of the node mapped back to back into one virtual address range."""

    def __init__(self, shard_bytes: int, world: int, rank: int):
        ensure_init()
        self._h = ctypes.c_void_p()
        fd = ctypes.c_int(-1)
        _check(load().aqs_flat_create(shard_bytes, world, rank, ctypes.byref(self._h), ctypes.byref(fd)))
        self.fd = fd.value          # POSIX handle of this rank's shard, to be passed to the other processes
        self._views = {}            # views of the state for staged passes, by their block lists (sharded.py)

    def attach(self, peer_rank: int, fd: int):
        _check(load().aqs_flat_attach(self._h, peer_rank, fd))

    def pointers(self):
        base, own = ctypes.c_void_p(), ctypes.c_void_p()
        _check(load().aqs_flat_ptr(self._h, ctypes.byref(base), ctypes.byref(own)))
        return base.value, own.value

    def view(self, blocks) -> int:
        """Base of a view of the state: own shard as usual, the listed (state_offset, bytes) blocks of peer shards backed
        by local staging memory, everything else unmapped."""
        arr = np.ascontiguousarray(np.asarray(blocks, dtype=np.uint64).reshape(-1, 2))
        p = ctypes.c_void_p()
        _check(load().aqs_flat_view_create(self._h, arr.ctypes.data_as(ctypes.c_void_p), len(arr), ctypes.byref(p)))
        return p.value

    def close(self):
        if getattr(self, "_h", None):
            load().aqs_flat_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """aqs_plan_t: a compiled op list (QCircuit::compile, src/quantum.cpp:199-210)."""

    def __init__(self, n_qubits: int, ops: np.ndarray, flags: int = 0):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self._h = ctypes.c_void_p()
        self.n_qubits = n_qubits
        _check(load().aqs_plan_build(n_qubits, ops.ctypes.data_as(ctypes.c_void_p), len(ops), flags,
                                     ctypes.byref(self._h)))

    def info(self) -> dict:
        pi = PlanInfo()
        _check(load().aqs_plan_get_info(self._h, ctypes.byref(pi)))
        return {k: getattr(pi, k) for k, _ in PlanInfo._fields_}

    def pass_span(self, index: int, log2_world: int) -> int:
        """How many rank bits (top log2_world index bits) the tile of fused pass `index` contains."""
        v = ctypes.c_int()
        _check(load().aqs_plan_pass_span(self._h, index, log2_world, ctypes.byref(v)))
        return v.value

    def pass_tile(self, index: int):
        """Index-bit positions of the tile of fused pass `index`, ascending."""
        pos = (ctypes.c_uint8 * 16)()
        t = ctypes.c_int()
        _check(load().aqs_plan_pass_tile(self._h, index, pos, ctypes.byref(t)))
        return [int(pos[j]) for j in range(t.value)]

    def shard_cut(self, index: int, rank: int, log2_world: int):
        """(positions, value): rank `rank` runs the tiles of pass `index` whose number has these bits at this value."""
        n, v = ctypes.c_uint32(), ctypes.c_uint32()
        pos = (ctypes.c_uint8 * 8)()
        _check(load().aqs_plan_shard_cut(self._h, index, rank, log2_world, ctypes.byref(n), ctypes.byref(v), pos))
        return [int(pos[i]) for i in range(n.value)], int(v.value)

    def export_pass(self, index: int) -> bytes:
        """Raw launch descriptors of fused pass `index` (tests/tile_emulator.py parses them)."""
        need = ctypes.c_uint64()
        _check(load().aqs_plan_export_pass(self._h, index, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        _check(load().aqs_plan_export_pass(self._h, index, buf, need.value, ctypes.byref(need)))
        return buf.raw

    def pass_source(self, index: int):
        """(source, coefs, threads, smem_bytes, n_ctas) of the specialised kernel of fused pass `index` (specialize.cu)."""
        need = ctypes.c_uint64()
        geom = (ctypes.c_uint64 * 3)()
        _check(load().aqs_plan_pass_source(self._h, index, None, 0, ctypes.byref(need), geom))
        buf = ctypes.create_string_buffer(need.value)
        _check(load().aqs_plan_pass_source(self._h, index, buf, need.value, ctypes.byref(need), geom))
        _check(load().aqs_plan_pass_coefs(self._h, index, None, 0, ctypes.byref(need)))
        coefs = np.zeros(need.value, dtype=np.uint64)
        _check(load().aqs_plan_pass_coefs(self._h, index, coefs.ctypes.data_as(ctypes.c_void_p), need.value, ctypes.byref(need)))
        return buf.value.decode(), coefs, int(geom[0]), int(geom[1]), int(geom[2])

    def jit_ready(self) -> int:
        """Number of fused passes that would run on their specialised kernel if the plan ran now."""
        v = ctypes.c_uint64()
        _check(load().aqs_plan_jit_ready(self._h, ctypes.byref(v)))
        return int(v.value)

    def close(self):
        if self._h:
            load().aqs_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def jit_wait() -> None:
    """Block until every queued specialisation has been compiled."""
    _check(load().aqs_jit_wait())


def jit_info() -> dict:
    ji = JitInfo()
    _check(load().aqs_jit_get_info(ctypes.byref(ji)))
    return {k: getattr(ji, k) for k, _ in JitInfo._fields_}


class Timer:
    def __init__(self):
        self._h = ctypes.c_void_p()
        _check(load().aqs_timer_create(ctypes.byref(self._h)))

    def start(self, state: "State"):
        _check(load().aqs_timer_start(self._h, state._h))

    def stop(self, state: "State"):
        _check(load().aqs_timer_stop(self._h, state._h))

    def elapsed_ms(self) -> float:
        ms = ctypes.c_double()
        _check(load().aqs_timer_elapsed_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self._h:
                load().aqs_timer_destroy(self._h)
        except Exception:
            pass


class State:
    """aqs_state_t: a 2^n complex64 state vector resident in HBM."""

    def __init__(self, n_qubits: int, _handle=None):
        ensure_init()
        self.n = n_qubits
        self.size = 1 << n_qubits
        if _handle is not None:
            self._h = _handle
        else:
            self._h = ctypes.c_void_p()
            _check(load().aqs_state_create(n_qubits, ctypes.byref(self._h)))

    @classmethod
    def wrap(cls, n_qubits: int, device_ptr: int) -> "State":
        """Engine state on caller-owned device memory (e.g. a torch tensor's data_ptr())."""
        ensure_init()
        h = ctypes.c_void_p()
        _check(load().aqs_state_wrap(n_qubits, ctypes.c_void_p(device_ptr), ctypes.byref(h)))
        return cls(n_qubits, _handle=h)

    def close(self):
        if getattr(self, "_h", None):
            load().aqs_state_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clone(self) -> "State":
        h = ctypes.c_void_p()
        _check(load().aqs_state_clone(self._h, ctypes.byref(h)))
        return State(self.n, _handle=h)

    # -- preparation / transfer --------------------------------------------
    def set_basis(self, index: int = 0):
        _check(load().aqs_state_set_basis(self._h, index))

    def set_product(self, qstates):
        q = np.ascontiguousarray(np.asarray(qstates, dtype=np.complex64).reshape(self.n, 2))
        _check(load().aqs_state_set_product(self._h, q.ctypes.data_as(ctypes.c_void_p)))

    def set_identity(self):
        _check(load().aqs_state_set_identity(self._h))

    def upload(self, host: np.ndarray, offset: int = 0):
        host = np.ascontiguousarray(host, dtype=np.complex64)
        _check(load().aqs_state_upload(self._h, host.ctypes.data_as(ctypes.c_void_p), offset, host.size))

    def download(self, offset: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        count = self.size - offset if count is None else count
        if out is None:
            out = np.empty(count, dtype=np.complex64)
        _check(load().aqs_state_download(self._h, out.ctypes.data_as(ctypes.c_void_p), offset, count))
        return out

    def amp(self, index: int) -> np.complex64:
        out = np.empty(1, dtype=np.complex64)
        _check(load().aqs_state_get_amp(self._h, index, out.ctypes.data_as(ctypes.c_void_p)))
        return out[0]

    def device_ptr(self) -> int:
        p = ctypes.c_void_p()
        _check(load().aqs_state_device_ptr(self._h, ctypes.byref(p)))
        return p.value

    def set_stream(self, cuda_stream: int):
        _check(load().aqs_state_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        _check(load().aqs_sync(self._h))

    # -- peer memory (sharded states) -----------------------------------------
    def ipc_export(self) -> bytes:
        """CUDA IPC handle of this engine-allocated state (64 bytes) for the other ranks of the node."""
        buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        _check(load().aqs_state_ipc_export(self._h, buf))
        return buf.raw

    def peer_bitswap(self, members: Sequence[int], local_bits: Sequence[int], my_value: int):
        """Swap k (rank bit, local index bit) pairs in place over peer memory (aqs_peer_bitswap).
        members[v] = device pointer of the group member whose selected rank bits read v."""
        k = len(local_bits)
        arr = (ctypes.c_void_p * (1 << k))(*[ctypes.c_void_p(int(p)) if p else None for p in members])
        lb = (ctypes.c_int * k)(*[int(b) for b in local_bits])
        _check(load().aqs_peer_bitswap(self._h, arr, k, lb, my_value))

    # -- gates ----------------------------------------------------------------
    def apply_ops(self, ops: np.ndarray):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        _check(load().aqs_apply_ops(self._h, ops.ctypes.data_as(ctypes.c_void_p), len(ops)))

    def apply_dense(self, qubits: Sequence[int], matrix: np.ndarray, controls: Iterable[int] = (), ctrl_value: Optional[int] = None):
        """An opaque 2^k x 2^k matrix (row-major, qubits[0] = most significant index bit) on k qubits, k <= 6."""
        k = len(qubits)
        m = np.ascontiguousarray(np.asarray(matrix, dtype=np.complex64).reshape(1 << k, 1 << k))
        q = (ctypes.c_int * k)(*[int(x) for x in qubits])
        cm = qmask(controls)
        _check(load().aqs_apply_dense(self._h, q, k, cm, cm if ctrl_value is None else ctrl_value,
                                      m.ctypes.data_as(ctypes.c_void_p)))

    def run(self, plan: Plan):
        _check(load().aqs_plan_run(self._h, plan._h))

    def run_tiles(self, plan: Plan, index: int, load_base: int, fix_pos, fix_or: int, stream: int = 0):
        """One fused pass on the tiles whose number has the bits `fix_pos` pinned to those of `fix_or`, reading the state at
        `load_base` (0: in place), writing in place, on CUDA stream `stream` (0: the state's own)."""
        pos = (ctypes.c_uint8 * 8)(*fix_pos)
        _check(load().aqs_plan_run_tiles(self._h, plan._h, index, ctypes.c_void_p(load_base or None), len(fix_pos), pos, fix_or,
                                         ctypes.c_void_p(stream or None)))

    def run_shard(self, plan: Plan, first: int, count: int, rank: int, log2_world: int):
        """This rank's share of passes [first, first + count) on a flat multi-GPU state (aqs_plan_run_shard)."""
        _check(load().aqs_plan_run_shard(self._h, plan._h, first, count, rank, log2_world))

    # -- measurement ------------------------------------------------------------
    def norm2(self) -> float:
        v = ctypes.c_double()
        _check(load().aqs_norm2(self._h, ctypes.byref(v)))
        return v.value

    def scale(self, f: float):
        _check(load().aqs_scale(self._h, f))

    def prob_fixed(self, qubit_mask: int = 0, qubit_value: int = 0) -> int:
        v = ctypes.c_uint64()
        _check(load().aqs_prob_fixed(self._h, qubit_mask, qubit_value, ctypes.byref(v)))
        return v.value

    def qubit_prob1(self, qubit: int) -> float:
        v = ctypes.c_double()
        _check(load().aqs_qubit_prob1(self._h, qubit, ctypes.byref(v)))
        return v.value

    def probabilities(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.size - offset if count is None else count
        out = np.empty(count, dtype=np.float32)
        _check(load().aqs_probabilities(self._h, out.ctypes.data_as(ctypes.c_void_p), offset, count))
        return out

    def collapse_qubit(self, qubit: int, outcome: int, p: float):
        _check(load().aqs_collapse_qubit(self._h, qubit, outcome, p))

    def sample(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float32)
        out = np.empty(u.size, dtype=np.uint64)
        _check(load().aqs_sample(self._h, u.ctypes.data_as(ctypes.c_void_p), u.size,
                                 out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def sample_fixed(self, u_fixed: np.ndarray) -> np.ndarray:
        """Local indices for fixed-point thresholds (2^-62 units); UINT64_MAX = not in this shard."""
        u = np.ascontiguousarray(u_fixed, dtype=np.uint64)
        out = np.empty(u.size, dtype=np.uint64)
        _check(load().aqs_sample_fixed(self._h, u.ctypes.data_as(ctypes.c_void_p), u.size,
                                       out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def sample_hist(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float32)
        hist = np.empty(self.size, dtype=np.uint32)
        _check(load().aqs_sample_hist(self._h, u.ctypes.data_as(ctypes.c_void_p), u.size,
                                      hist.ctypes.data_as(ctypes.c_void_p)))
        return hist

    def sample_hist_sparse(self, u: np.ndarray):
        """profile_measure_all as sorted (index, count) pairs: 8 bytes per draw cross the bus, no dense 2^n vector."""
        u = np.ascontiguousarray(u, dtype=np.float32)
        idx = np.empty(max(1, u.size), dtype=np.uint64)
        cnt = np.empty(max(1, u.size), dtype=np.uint32)
        bins = ctypes.c_uint64()
        _check(load().aqs_sample_hist_sparse(self._h, u.ctypes.data_as(ctypes.c_void_p), u.size, idx.ctypes.data_as(ctypes.c_void_p),
                                             cnt.ctypes.data_as(ctypes.c_void_p), idx.size, ctypes.byref(bins)))
        return idx[:bins.value].copy(), cnt[:bins.value].copy()


def memcpy_async(dst: int, src: int, nbytes: int, stream: int) -> None:
    _check(load().aqs_memcpy_async(ctypes.c_void_p(dst), ctypes.c_void_p(src), nbytes, ctypes.c_void_p(stream or None)))


def pool_trim() -> None:
    _check(load().aqs_pool_trim())

"""Test infrastructure: run the GENERATED source of specialised passes (afquantumsim_b200/csrc/specialize.cu) on the CPU.

The generator emits CUDA C++ that is also valid host C++ under -DAQS_HOST_EMU (threads of a CTA = std::thread,
__syncthreads = a pthread barrier, fma.rn.f32x2 = two fmaf).  This module compiles it with g++ and runs it on a numpy
state, so the generated code itself — not a description of it — is checked against the oracle without a GPU.  Never
imported by the package."""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_CACHE = os.path.join(tempfile.gettempdir(), "aqs_spec_emu")


def compile_source(src: str) -> ctypes.CDLL:
    os.makedirs(_CACHE, exist_ok=True)
    key = hashlib.sha1(src.encode()).hexdigest()[:20]
    so = os.path.join(_CACHE, f"p{key}.so")
    if not os.path.exists(so):
        cpp = os.path.join(_CACHE, f"p{key}.cpp")
        with open(cpp, "w") as fh:
            fh.write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DAQS_HOST_EMU", "-shared", "-fPIC", "-pthread",
                               "-o", so + ".tmp", cpp])
        os.replace(so + ".tmp", so)
    lib = ctypes.CDLL(so)
    lib.aqs_pass_emu_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    lib.aqs_pass_emu_run.restype = None
    return lib


def run_pass(plan, index: int, state: np.ndarray, cut=None):
    """apply specialised pass `index` of `plan` to `state` (complex64, in place).  cut = (positions, value): a sharded launch."""
    src, coefs, threads, smem, n_ctas = plan.pass_source(index)
    lib = compile_source(src)
    fix_pos = (ctypes.c_uint8 * 8)()
    fix_n = fix_or = 0
    if cut is not None:
        pos, fix_or = cut
        fix_n = len(pos)
        for i, p in enumerate(pos):
            fix_pos[i] = p
        n_ctas >>= fix_n
    assert state.dtype == np.complex64 and state.flags.c_contiguous
    lib.aqs_pass_emu_run(state.ctypes.data_as(ctypes.c_void_p), coefs.ctypes.data_as(ctypes.c_void_p), n_ctas, fix_n, fix_or, fix_pos)
    return state


def run_plan(plan, state: np.ndarray):
    state = np.array(state, dtype=np.complex64)
    for i in range(plan.info()["n_fused_passes"]):
        run_pass(plan, i, state)
    return state

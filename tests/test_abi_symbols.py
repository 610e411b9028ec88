"""CPU: the C-ABI library loads and exports every symbol include/aqs_engine.h declares."""
import ctypes
import os
import re

from afquantumsim_b200 import engine as eng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aqs_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aqs_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(eng.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(eng.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.aqs_engine_abi_version() == 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    L = eng.load()
    assert L.aqs_engine_init(0) != 0
    assert b"CUDA" in L.aqs_last_error() or b"device" in L.aqs_last_error()
    h = ctypes.c_void_p()
    assert L.aqs_state_create(4, ctypes.byref(h)) != 0      # not initialised -> error, never a CPU path


def test_op_record_layout():
    assert eng.OP_DTYPE.itemsize == 64
    assert eng.OP_DTYPE.fields["ctrl_mask"][1] == 16 and eng.OP_DTYPE.fields["m"][1] == 32

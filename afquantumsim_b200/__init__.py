"""afquantumsim_b200 — B200-native state-vector engine behind the afQuantumSim API.

Layout:
  csrc/      CUDA kernels + the C ABI (include/aqs_engine.h) -> lib/libaqs_engine.so
  host/      C++14 host layer mirroring the reference's aqs:: API -> lib/libafquantum.so
  engine.py  ctypes binding of the engine ABI
  aqs.py     Python mirror of the aqs:: classes, over the host layer's C wrapper
  workloads.py  synthetic circuits of BASELINE.json's configs
"""
__version__ = "0.1.0"

"""Peer-memory remap (aqs_peer_bitswap) checks, shared by the CPU and the GPU suites.

  python tests/peer_cases.py --abi cpu     in-process members on the oracle-backed ABI stand-in
  (the GPU suite imports `check_bitswap` and runs it on the CUDA engine)

2^k states of one process play the members of a rank group.  Every member issues the call; together
the calls must realise the permutation "swap rank bit i with local index bit local_bits[i]" of the
concatenated 2^(n_local + k) vector, which numpy restates below.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def bitswap_reference(full: np.ndarray, n_local: int, local_bits) -> np.ndarray:
    """out[index with (rank bit i, local bit local_bits[i]) exchanged for every i] = full[index]"""
    idx = np.arange(full.size, dtype=np.int64)
    dst = idx.copy()
    for i, lb in enumerate(local_bits):
        gb = n_local + i
        a, b = (idx >> gb) & 1, (idx >> lb) & 1
        dst = dst & ~((1 << gb) | (1 << lb)) | (b << gb) | (a << lb)
    out = np.empty_like(full)
    out[dst] = full
    return out


def check_bitswap(eng, n_local: int, local_bits, seed: int):
    k = len(local_bits)
    rng = np.random.default_rng(seed)
    full = (rng.standard_normal(1 << (n_local + k)) + 1j * rng.standard_normal(1 << (n_local + k))).astype(np.complex64)
    states = []
    for v in range(1 << k):
        s = eng.State(n_local)
        s.upload(full[v << n_local:(v + 1) << n_local])
        states.append(s)
    ptrs = [s.device_ptr() for s in states]
    for s in states:
        s.sync()
    for v, s in enumerate(states):          # members run one after the other: their swaps are disjoint
        s.peer_bitswap(ptrs, local_bits, v)
        s.sync()
    got = np.concatenate([s.download() for s in states])
    want = bitswap_reference(full, n_local, local_bits)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (n_local, local_bits)
    for s in states:
        s.close()


CASES = [(4, [1], 1), (5, [3, 1], 2), (6, [2, 5, 3], 3), (10, [9], 4), (11, [4, 10], 5), (12, [11, 6, 1], 6), (14, [7, 12, 13], 7)]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--abi", default="cpu")
    a = ap.parse_args()
    from afquantumsim_b200 import engine as eng
    if a.abi == "cpu":
        eng.LIB_PATH = os.path.join(ROOT, "oracle", "_build", "cpu_abi", "libaqs_engine.so")   # test double
    eng.init(0)
    for n_local, lbits, seed in CASES:
        check_bitswap(eng, n_local, lbits, seed)
    print("ok peer_cases")

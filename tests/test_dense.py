"""aqs_apply_dense: an opaque k-qubit matrix on arbitrary qubits (tests/dense_cases.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import dense_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_numpy_restatement_matches_the_oracle_embedding():
    """the einsum restatement used by the checks agrees with the oracle's literal Gate embedding"""
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    n, k, b = 7, 3, 2
    inner = orc.Circ(k, [("H", 0), ("CX", 0, 1), ("RotY", 2, 0.7), ("CX", 1, 2), ("Phase", 0, 1.1)])
    U = orc.circuit_matrix(inner, mode="dense")
    a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
    want = orc.simulate(a.copy(), orc.Circ(n, [("Gate", inner, b)]), mode="dense")
    assert dense_cases.rel_l2(dense_cases.dense_reference(a, n, [b, b + 1, b + 2], U), want) < 1e-6


def test_apply_dense_cpu_abi():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dense_cases.py"), "--abi", "cpu"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "ok dense_cases" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_apply_dense_gpu():
    from afquantumsim_b200 import engine as eng
    from oracle import oracle as orc
    eng.ensure_init()
    dense_cases.run_cases(eng, orc, sizes=((9, 40), (13, 30), (20, 12)))


@pytest.mark.gpu
def test_apply_dense_rejects_bad_arguments():
    from afquantumsim_b200 import engine as eng
    eng.ensure_init()
    s = eng.State(8)
    eye = np.eye(4, dtype=np.complex64)
    with pytest.raises(eng.EngineError):
        s.apply_dense([1, 1], eye)                      # duplicate target
    with pytest.raises(eng.EngineError):
        s.apply_dense([1, 9], eye)                      # out of range
    with pytest.raises(eng.EngineError):
        s.apply_dense([1, 2], eye, controls=[2])        # control is a target
    with pytest.raises(eng.EngineError):
        s.apply_dense(list(range(7)), np.eye(128, dtype=np.complex64))   # k > 6

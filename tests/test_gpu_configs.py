"""GPU: oracle parity on the BENCHMARKED configurations (BASELINE.json configs 3 and 4 at sizes the CPU oracle still
finishes in about a minute): full-depth brickwork-26 and a Grover-26 slice, through the public aqs API, both on the
generic tile kernel and on the specialised kernels.  Amplitudes <= 1e-5 relative L2 against the oracle (the north-star
tolerance); sampling on the resulting state bit-identical given the same draws."""
import os

import numpy as np
import pytest

from afquantumsim_b200 import aqs
from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l2_chunked(state: "eng.State", ref: np.ndarray) -> float:
    num = den = 0.0
    chunk = 1 << 24
    for off in range(0, ref.size, chunk):
        got = state.download(off, min(chunk, ref.size - off))
        d = got - ref[off:off + got.size]
        num += float(np.vdot(d, d).real)
        den += float(np.vdot(ref[off:off + got.size], ref[off:off + got.size]).real)
    return (num / den) ** 0.5


@pytest.fixture(scope="module")
def brickwork26():
    n = 26
    gates = wl.brickwork(n, 20)
    return n, gates, orc.simulate(orc.new_state(n), orc.Circ(n, gates))


@pytest.mark.parametrize("jit", [False, True])
def test_brickwork26_full_depth_matches_oracle(brickwork26, jit):
    n, gates, want = brickwork26
    ops = aqs.QCircuit(n).extend(gates).ops()
    plan = eng.Plan(n, ops, eng.PLAN_FUSE | (eng.PLAN_JIT if jit else 0))
    if jit:
        assert plan.jit_ready() == plan.info()["n_fused_passes"]
    st = eng.State(n)
    st.run(plan)
    assert rel_l2_chunked(st, want) < TOL
    u = np.random.default_rng(26).random(2000, dtype=np.float32)
    assert np.array_equal(st.sample(u), orc.sample(st.download(), u, "exact"))      # same state, same draws: bit-identical
    st.close()


def test_brickwork26_through_the_host_layer(brickwork26):
    """aqs.QCircuit -> compile() (specialised kernels on >= 26 qubits) -> QSimulator.simulate, as a user would"""
    n, gates, want = brickwork26
    qc = aqs.QCircuit(n).extend(gates)
    qc.compile()
    qs = aqs.QSimulator(n)
    qs.simulate(qc)
    got = qs.statevector()
    assert orc.rel_l2(got, want) < TOL


def test_grover26_slice_matches_oracle():
    """examples/grover_search.cpp at 26 qubits, marked state 5: a slice of the 6433 iterations against the oracle
    (amplitudes, not only the marked probability) and against the closed form sin^2((2k+1) theta)"""
    n, marked = 26, 5
    iters = int(os.environ.get("AQS_GROVER_TEST_ITERS", "64"))
    circ = orc.grover_search(n, orc.grover_oracle(n, marked), iters)
    want = orc.simulate(orc.new_state(n), circ)
    qc = aqs.QCircuit(n)
    qc << aqs.Gate(aqs.grover_search(n, aqs.grover_oracle(n, marked), iters, "Oracle"), 0)
    ops = qc.ops()
    w = int(format(marked, f"0{n}b")[::-1], 2)
    closed = float(np.sin((2 * iters + 1) * np.arcsin(2.0 ** (-n / 2))) ** 2)
    for flags in (eng.PLAN_FUSE, eng.PLAN_FUSE | eng.PLAN_JIT):
        st = eng.State(n)
        st.run(eng.Plan(n, ops, flags))
        assert rel_l2_chunked(st, want) < TOL
        p_w = float(abs(st.amp(w)) ** 2)
        assert abs(p_w - float(abs(want[w]) ** 2)) <= 5e-5 * p_w          # oracle-relative (amplitude 1e-5 -> probability 2e-5, plus margin)
        assert abs(p_w - closed) / closed < 2e-3                           # closed form (the reference's float H drifts the norm)
        st.close()

"""GPU: the specialised (run-time compiled) pass kernels of afquantumsim_b200/csrc/specialize.cu against the CPU oracle
and against the generic tile kernel running the same plan.  Every test asserts that the specialised kernels are the
ones that ran (Plan.jit_ready() == number of fused passes): a silent fallback to the interpreter would pass parity."""
import numpy as np
import pytest

from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc
from tests.lowering import lower_array
from tests.test_gpu_engine import random_circuit, random_state

pytestmark = pytest.mark.gpu
TOL = 1e-5
JIT = eng.PLAN_FUSE | eng.PLAN_JIT


def run(n, init, ops, flags, expect_jit=True):
    plan = eng.Plan(n, ops, flags)
    passes = plan.info()["n_fused_passes"]
    assert passes > 0
    if flags & eng.PLAN_JIT and expect_jit:
        assert plan.jit_ready() == passes, (plan.jit_ready(), passes, eng.jit_info())
    s = eng.State(n)
    s.upload(init)
    s.run(plan)
    out = s.download()
    s.close()
    return out


@pytest.mark.parametrize("n,gates,seed", [(10, 40, 1), (12, 80, 2), (13, 120, 3), (16, 150, 4), (20, 200, 5)])
def test_random_circuits_match_oracle_and_interpreter(n, gates, seed):
    circ = random_circuit(n, gates, 7000 + seed)
    init = random_state(n, seed)
    ops = lower_array(circ)
    got = run(n, init, ops, JIT)
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
    assert orc.rel_l2(got, run(n, init, ops, eng.PLAN_FUSE)) < 2e-6


@pytest.mark.parametrize("n", [20, 23])
def test_brickwork_matches_oracle(n):
    circ = orc.Circ(n, wl.brickwork(n, 20))
    init = orc.new_state(n)
    got = run(n, init, lower_array(circ), JIT)
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL


def test_permutation_circuits_are_exact():
    n = 14
    r = np.random.default_rng(11)
    gates = []
    for _ in range(80):
        q = [int(x) for x in r.choice(n, 3, replace=False)]
        gates.append([("X", q[0]), ("CX", q[0], q[1]), ("Swap", q[0], q[1]), ("CCNot", q[0], q[1], q[2]),
                      ("CSwap", q[0], q[1], q[2])][int(r.integers(5))])
    circ = orc.Circ(n, gates)
    init = random_state(n, 11)
    assert np.array_equal(run(n, init, lower_array(circ), JIT), orc.simulate(init.copy(), circ))


def test_qft_closed_form_and_grover():
    n = 22
    x = 0x2a5a5
    init = np.zeros(1 << n, dtype=np.complex64)
    init[x] = 1
    got = run(n, init, lower_array(orc.Circ(n, wl.qft(n))), JIT)
    rev = int(format(x, f"0{n}b")[::-1], 2)
    y = np.arange(1 << n, dtype=np.float64)
    want = np.exp(2j * np.pi * ((rev * y) % (1 << n)) / (1 << n)) / np.sqrt(float(1 << n))
    assert np.max(np.abs(got - want)) * np.sqrt(float(1 << n)) < 2e-5
    n = 16
    circ = orc.grover_search(n, orc.grover_oracle(n, 5), 4)
    init = orc.new_state(n)
    assert orc.rel_l2(run(n, init, lower_array(circ), JIT), orc.simulate(init.copy(), circ)) < TOL


@pytest.mark.parametrize("tile_bits", [10, 11, 12])
def test_other_tile_sizes(tile_bits, monkeypatch):
    monkeypatch.setenv("AQS_TILE_BITS", str(tile_bits))
    n = 16
    circ = random_circuit(n, 120, 99 + tile_bits)
    init = random_state(n, tile_bits)
    assert orc.rel_l2(run(n, init, lower_array(circ), JIT), orc.simulate(init.copy(), circ)) < TOL


def test_long_pass_with_hundreds_of_coefficient_parameters():
    n = 12
    circ = orc.Circ(n, wl.brickwork(n, 12))         # one tile: every gate lands in one or two passes
    init = random_state(n, 12)
    plan = eng.Plan(n, lower_array(circ), JIT)
    assert max(len(plan.pass_source(i)[1]) for i in range(plan.info()["n_fused_passes"])) > 300
    assert orc.rel_l2(run(n, init, lower_array(circ), JIT), orc.simulate(init.copy(), circ)) < TOL


@pytest.mark.parametrize("n,g", [(16, 1), (17, 2), (18, 3)])
def test_sharded_launches_of_specialised_passes(n, g):
    """aqs_plan_run_shard on ONE GPU, the 2^g ranks played one after the other: bit-identical to the ordinary run"""
    circ = random_circuit(n, 100, 300 + n)
    ops = lower_array(circ)
    init = random_state(n, n)
    plan = eng.Plan(n, ops, JIT)
    passes = plan.info()["n_fused_passes"]
    assert plan.jit_ready() == passes
    want = run(n, init, ops, JIT)
    s = eng.State(n)
    s.upload(init)
    for i in range(passes):
        for rank in range(1 << g):
            s.run_shard(plan, i, 1, rank, g)
    assert np.array_equal(s.download(), want)


def test_background_compilation_falls_back_then_switches():
    n = 18
    circ = orc.Circ(n, wl.brickwork(n, 9))
    ops = lower_array(circ)
    init = random_state(n, 18)
    want = orc.simulate(init.copy(), circ)
    plan = eng.Plan(n, ops, eng.PLAN_FUSE | eng.PLAN_JIT_ASYNC)      # returns at once; passes run on whatever is ready
    s = eng.State(n)
    s.upload(init)
    s.run(plan)
    assert orc.rel_l2(s.download(), want) < TOL
    eng.jit_wait()
    assert eng.jit_info()["pending"] == 0
    assert plan.jit_ready() == plan.info()["n_fused_passes"]
    s.upload(init)
    s.run(plan)
    assert orc.rel_l2(s.download(), want) < TOL
    # a second plan of the same circuit finds every shape in the cache
    before = eng.jit_info()
    plan2 = eng.Plan(n, ops, JIT)
    after = eng.jit_info()
    assert after["compiled"] == before["compiled"] and after["cache_hits"] - before["cache_hits"] == plan2.info()["n_fused_passes"]

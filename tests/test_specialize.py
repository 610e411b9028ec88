"""CPU: the GENERATED source of specialised passes (afquantumsim_b200/csrc/specialize.cu), compiled with g++ under
-DAQS_HOST_EMU and executed on numpy states (tests/spec_emu.py), against the oracle and against the numpy emulator of
the generic tile kernel.  This pins the code generator — register renaming of exact permutations, skipped identity
branches, per-thread coefficient selects, immediate re-split addresses, ladders — without a GPU; the GPU tests then run
the same source through NVRTC."""
import numpy as np
import pytest

from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc
from tests import spec_emu
from tests import tile_emulator as te
from tests.lowering import lower_array
from tests.test_gpu_engine import random_circuit
from tests.test_tile_plan import random_state

TOL = 1e-5


def both(n, circ, init):
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    assert plan.info()["n_fused_passes"] > 0
    return spec_emu.run_plan(plan, init), te.run_plan(plan, init), plan


@pytest.mark.parametrize("seed", range(4))
def test_random_circuits_all_gate_classes(seed):
    n = 10 + seed
    circ = random_circuit(n, 60, 4000 + seed)
    init = random_state(n, seed)
    got, ref, _ = both(n, circ, init)
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
    assert orc.rel_l2(got, ref) < 2e-6          # same plan, same arithmetic up to packed / scalar contraction


def test_brickwork_and_controlled_gates_on_every_bit_kind():
    n = 13
    circ = orc.Circ(n, wl.brickwork(n, 6))
    init = random_state(n, 3)
    got, ref, plan = both(n, circ, init)
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
    assert orc.rel_l2(got, ref) < 2e-6
    src = plan.pass_source(0)[0]
    assert "fma2(" in src and "aqs_pass" in src


def test_qft_ladders_and_grover():
    n = 12
    init = random_state(n, 5)
    for circ in (orc.Circ(n, wl.qft(n)), orc.grover_search(n, orc.grover_oracle(n, 5), 2)):
        got, ref, _ = both(n, circ, init)
        assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
        assert orc.rel_l2(got, ref) < 2e-6


def test_permutation_circuits_are_exact():
    """X / CX / Swap / CCNot / CSwap only move amplitudes: renamed variables and conditional swaps must not round them"""
    n = 12
    r = np.random.default_rng(7)
    gates = []
    for _ in range(40):
        q = [int(x) for x in r.choice(n, 3, replace=False)]
        gates.append([("X", q[0]), ("CX", q[0], q[1]), ("Swap", q[0], q[1]), ("CCNot", q[0], q[1], q[2]),
                      ("CSwap", q[0], q[1], q[2])][int(r.integers(5))])
    circ = orc.Circ(n, gates)
    init = random_state(n, 7)
    got, _, _ = both(n, circ, init)
    assert np.array_equal(got, orc.simulate(init.copy(), circ))


def test_ghz_is_exact_and_sharded_launches_partition_the_pass():
    n = 14
    circ = orc.Circ(n, wl.ghz(n))
    got, _, _ = both(n, circ, orc.new_state(n))
    assert np.array_equal(got, orc.simulate(orc.new_state(n), circ))
    # the fix_n / fix_or / fix_pos launch parameters: 4 "ranks" one after the other == one full launch
    n = 15
    circ = orc.Circ(n, wl.brickwork(n, 4))
    init = random_state(n, 9)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    full = spec_emu.run_plan(plan, init)
    st = np.array(init, dtype=np.complex64)
    for i in range(plan.info()["n_fused_passes"]):
        for rank in range(4):
            spec_emu.run_pass(plan, i, st, cut=plan.shard_cut(i, rank, 2))
    assert np.array_equal(st, full)


def test_generation_is_deterministic_and_value_free():
    """the same circuit generates the same source and table (that is what the process-wide kernel cache keys on);
    matrix entries appear only in the coefficient table, never as literals in the source"""
    n = 12
    g1 = wl.brickwork(n, 4)
    p1 = eng.Plan(n, lower_array(orc.Circ(n, g1)), eng.PLAN_FUSE)
    p2 = eng.Plan(n, lower_array(orc.Circ(n, list(g1))), eng.PLAN_FUSE)
    s1, c1 = p1.pass_source(0)[:2]
    s2, c2 = p2.pass_source(0)[:2]
    assert s1 == s2 and np.array_equal(c1, c2)
    body = s1[s1.index("u32 tile_no"):]
    import re
    assert not re.search(r"\d\.\d+f?\b", body), "a floating-point literal in the kernel body"


def test_rotations_cost_two_packed_fmas_per_pair():
    """deferred scales: a plain rotation is 2 FFMA2 per amplitude pair (16 pairs per thread), not three shears"""
    n = 12
    circ = orc.Circ(n, [("RotY", 3, 0.7), ("RotX", 5, -1.1)])
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    src = plan.pass_source(0)[0]
    body = src[src.index("u32 tile_no"):]
    n_fp = body.count("fma2(") + body.count("mul2(")
    assert n_fp <= 2 * 2 * 16 + 32 + 32, n_fp        # two ops x 2 x 16 pairs, at most one materialisation and the store scale
    init = random_state(n, 1)
    assert orc.rel_l2(spec_emu.run_plan(plan, init), orc.simulate(init.copy(), circ)) < TOL

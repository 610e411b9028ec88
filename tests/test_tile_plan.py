"""CPU: the fusion planner's launch descriptors, executed by the numpy emulator of the tile kernel
(tests/tile_emulator.py), against the oracle.  This pins everything about a fused plan except the
CUDA code itself: op rewriting (merges, CX folded into multiplexed rotations), pass/segment grouping,
layouts, shared-memory swizzles, predicates, in-place shear decompositions.  The GPU tests then
compare the real kernel with the same oracle."""
import numpy as np
import pytest

from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc
from tests import tile_emulator as te
from tests.lowering import lower_array
from tests.test_gpu_engine import random_circuit

TOL = 1e-5


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
    return (a / np.float32(np.sqrt(orc.norm2(a)))).astype(np.complex64)


def emulate(n, circ, init, stats=None):
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    assert plan.info()["n_fused_passes"] > 0
    return te.run_plan(plan, init, stats)


@pytest.mark.parametrize("seed", range(12))
def test_random_circuits_all_gate_classes(seed):
    n = 10 + seed % 4
    circ = random_circuit(n, 70, 1000 + seed)
    init = random_state(n, seed)
    stats = {}
    got = emulate(n, circ, init, stats)
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
    assert stats.get("conflicts", 0) == 0, "a re-split swizzle leaves shared-memory bank conflicts"


@pytest.mark.parametrize("n,depth", [(10, 4), (12, 8), (14, 10)])
def test_brickwork(n, depth):
    circ = orc.Circ(n, wl.brickwork(n, depth))
    init = random_state(n, n)
    assert orc.rel_l2(emulate(n, circ, init), orc.simulate(init.copy(), circ)) < TOL


@pytest.mark.parametrize("n", [10, 13])
def test_qft_and_grover(n):
    init = random_state(n, 7)
    circ = orc.Circ(n, wl.qft(n))
    assert orc.rel_l2(emulate(n, circ, init), orc.simulate(init.copy(), circ)) < TOL
    circ = orc.grover_search(n, orc.grover_oracle(n, 5), 3)
    assert orc.rel_l2(emulate(n, circ, init), orc.simulate(init.copy(), circ)) < TOL


def test_ghz_is_exact():
    for n in (10, 12, 14):
        circ = orc.Circ(n, wl.ghz(n))
        got = emulate(n, circ, orc.new_state(n))
        assert np.array_equal(got, orc.simulate(orc.new_state(n), circ))


@pytest.mark.parametrize("seed", range(4))
def test_permutation_circuits_are_exact(seed):
    """X / CX / Swap / CCNot / CSwap only move amplitudes: the fused path must not round them"""
    n = 12
    r = np.random.default_rng(seed)
    gates = []
    for _ in range(50):
        q = [int(x) for x in r.choice(n, 3, replace=False)]
        gates.append([("X", q[0]), ("CX", q[0], q[1]), ("Swap", q[0], q[1]), ("CCNot", q[0], q[1], q[2]),
                      ("CSwap", q[0], q[1], q[2])][int(r.integers(5))])
    circ = orc.Circ(n, gates)
    init = random_state(n, seed)
    assert np.array_equal(emulate(n, circ, init), orc.simulate(init.copy(), circ))


def test_cx_folds_into_rotations():
    """brickwork: every CX next to a rotation on its target becomes part of a multiplexed shear op"""
    n = 14
    plan = eng.Plan(n, lower_array(orc.Circ(n, wl.brickwork(n, 10))), eng.PLAN_FUSE)
    kinds = []
    for i in range(plan.info()["n_fused_passes"]):
        _, _, ops = te.parse(plan.export_pass(i))
        kinds += [op.kind for op in ops]
    n_gates = len(wl.brickwork(n, 10))
    assert len(kinds) < 0.75 * n_gates, (len(kinds), n_gates)
    assert sum(k in (te.PERM_R, te.PERM_I) for k in kinds) < 0.15 * len(kinds)
    assert sum(k == te.GEN for k in kinds) < 0.15 * len(kinds)     # edge qubits: rotations with no CX between them merge


def test_diagonal_gates_ride_inside_the_next_butterfly():
    """brickwork: a RotZ slides forward to the next rotation / multiplexed CX on its qubit and becomes that op's
    complex factor on y (TF_CY) instead of a phase op of its own"""
    n = 14
    gates = wl.brickwork(n, 10)
    plan = eng.Plan(n, lower_array(orc.Circ(n, gates)), eng.PLAN_FUSE)
    ops = []
    for i in range(plan.info()["n_fused_passes"]):
        ops += te.parse(plan.export_pass(i))[2]
    n_rotz = sum(g[0] == "RotZ" for g in gates)
    n_phase = sum(op.kind in (te.PHASE, te.PHASE_N) for op in ops)
    n_cy = sum(bool(op.flags & te.TF_CY) for op in ops)
    assert n_rotz > 30 and n_cy > 0.5 * n_rotz, (n_rotz, n_cy)
    assert n_phase < 0.35 * n_rotz, (n_rotz, n_phase)
    assert all(op.kind <= te.SHI and op.flags & te.TF_PY for op in ops if op.flags & te.TF_CY)
    init = random_state(n, 11)
    assert orc.rel_l2(te.run_plan(plan, init), orc.simulate(init.copy(), orc.Circ(n, gates))) < TOL


def test_qft_ladders_become_one_op_per_target():
    """QFT: the CPhase(j, i) of one target merge into a TK_LADDER op whose controls are spread over register,
    thread and tile-number bits; amplitudes still match the oracle"""
    for n, x in ((13, 0), (15, 0b101100111010110), (17, 0x1a2b3)):
        circ = orc.Circ(n, wl.qft(n))
        plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
        ops = []
        for i in range(plan.info()["n_fused_passes"]):
            ops += te.parse(plan.export_pass(i))[2]
        n_ladders = sum(op.kind == te.LADDER for op in ops)
        n_phase = sum(op.kind in (te.PHASE, te.PHASE_N, te.SCALE_R, te.SCALE_I) for op in ops)
        assert n_ladders >= n - 4, (n, n_ladders)
        assert n_phase <= 8, (n, n_phase)                  # the short ladders of the last targets stay plain phases
        init = orc.new_state(n, x)
        got = te.run_plan(plan, init)
        assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL
        init = random_state(n, n)
        assert orc.rel_l2(te.run_plan(plan, init), orc.simulate(init.copy(), circ)) < TOL


def test_descriptor_budget_splits_passes():
    """a 12-qubit state is one tile: 40 QFTs (3120 gates, ~700 descriptors with their controlled-phase ladders
    merged) exceed the per-pass descriptor budget and must split cleanly"""
    n = 12
    gates = []
    for _ in range(40):
        gates += wl.qft(n)
    circ = orc.Circ(n, gates)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    assert plan.info()["n_fused_passes"] >= 2
    init = random_state(n, 3)
    assert orc.rel_l2(te.run_plan(plan, init), orc.simulate(init.copy(), circ)) < TOL


@pytest.mark.parametrize("n,g,seed", [(14, 1, 21), (15, 2, 22), (16, 3, 23)])
def test_sharded_launches_partition_every_pass(n, g, seed):
    """Flat multi-GPU address space (aqs_plan_run_shard), on the CPU: the R ranks' launches of a pass touch disjoint
    tiles whose union is the whole state, a pass without rank bits in its tile stays inside each rank's own shard,
    and running the ranks one after the other reproduces the unsharded plan bit for bit."""
    circ = random_circuit(n, 120, seed)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    P = plan.info()["n_fused_passes"]
    spans = [plan.pass_span(i, g) for i in range(P)]
    assert any(spans)
    init = random_state(n, seed)
    whole = te.run_plan(plan, init)
    state = np.array(init, dtype=np.complex64)
    shard = 1 << (n - g)
    for i in range(P):
        raw = plan.export_pass(i)
        stats = {}
        for r in np.random.default_rng(seed + i).permutation(1 << g):
            pos, val = plan.shard_cut(i, int(r), g)
            assert len(pos) == g
            te.run_pass(state, raw, stats, cut=(pos, val))
            touched = stats["touched"][-1]
            if spans[i] == 0:
                assert touched.min() >= int(r) * shard and touched.max() < (int(r) + 1) * shard, "a local pass left its shard"
            else:
                own = np.count_nonzero((touched >= int(r) * shard) & (touched < (int(r) + 1) * shard))
                assert own * (1 << spans[i]) == touched.size, "a spanning pass should read 1 / 2^j of its tile from its own shard"
        allt = np.sort(np.concatenate(stats["touched"]))
        assert np.array_equal(allt, np.arange(1 << n)), "the ranks' launches do not partition the state"
    assert np.array_equal(state.view(np.uint32), whole.view(np.uint32))
    assert orc.rel_l2(state, orc.simulate(init.copy(), circ)) < TOL

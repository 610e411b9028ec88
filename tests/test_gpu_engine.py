"""GPU parity tests proper, at the C-ABI level (include/aqs_engine.h): every
case runs the CUDA path through ctypes and compares with the CPU oracle on the
same seeded inputs.  Amplitude tolerance: 1e-5 relative L2 (BASELINE.json
north_star); integer results (sampling, fixed-point probabilities) bit-exact."""
import numpy as np
import pytest

from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc
from tests.helpers import check_amp, decode_gates, load_kat, qstate_of
from tests.lowering import lower_array

pytestmark = pytest.mark.gpu
KAT = load_kat()
TOL = 1e-5


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
    a /= np.float32(np.sqrt(orc.norm2(a)))
    return a


def gpu_run(n, init, circ, fuse=False):
    s = eng.State(n)
    s.upload(init)
    ops = lower_array(circ)
    if fuse is None:
        s.apply_ops(ops)
    else:
        plan = eng.Plan(n, ops, eng.PLAN_FUSE if fuse else 0)
        s.run(plan)
    out = s.download()
    s.close()
    return out


def random_circuit(n, n_gates, seed):
    rng = np.random.default_rng(seed)
    gates = []
    one = ["X", "Y", "Z", "H", "Phase", "RotX", "RotY", "RotZ"]
    two = ["CX", "CY", "CZ", "CH", "CPhase", "CRotX", "CRotY", "CRotZ", "Swap"]
    three = ["CSwap", "CCNot", "Or"]
    for _ in range(n_gates):
        r = rng.random()
        if r < 0.45 or n < 2:
            name = one[rng.integers(len(one))]
            q = [int(rng.integers(n))]
        elif r < 0.85 or n < 3:
            name = two[rng.integers(len(two))]
            q = [int(x) for x in rng.choice(n, 2, replace=False)]
        else:
            name = three[rng.integers(len(three))]
            q = [int(x) for x in rng.choice(n, 3, replace=False)]
        if name in orc.HAS_ANGLE:
            gates.append((name, *q, float(np.float32(rng.uniform(-np.pi, np.pi)))))
        else:
            gates.append((name, *q))
    return orc.Circ(n, gates)


@pytest.mark.parametrize("case", KAT["cases"], ids=lambda c: c["name"])
def test_reference_kat_on_gpu(case):
    n = case["n"]
    s = eng.State(n)
    s.set_product([qstate_of(x) for x in case["init"]])
    s.apply_ops(lower_array(orc.Circ(n, decode_gates(case["circuit"]))))
    a = s.download()
    for idx, re, im in case["expect"]:
        assert check_amp(a[idx], re, im, case["tol"]), (case["name"], idx, a[idx])
        assert check_amp(s.amp(idx), re, im, case["tol"])


@pytest.mark.parametrize("case", KAT["equiv"], ids=lambda c: c["name"])
def test_reference_matrix_equivalences_on_gpu(case):
    """tests.cpp compares whole circuit matrices: apply both sides to I(2^n)
    laid out as a 2n-qubit state (qubits n..2n-1 index the rows)."""
    n = case["n"]
    mats = []
    for side in ("lhs", "rhs"):
        inner = orc.Circ(n, decode_gates(case[side]))
        s = eng.State(2 * n)
        s.set_identity()
        s.apply_ops(lower_array(orc.Circ(2 * n, [("Gate", inner, n)])))
        mats.append(s.download().reshape(1 << n, 1 << n).T)   # column-major -> U[r, c]
        s.close()
    if case["tol"] == 0.0:
        assert np.array_equal(mats[0], mats[1])
    else:
        assert np.max(np.abs(mats[0] - mats[1])) < case["tol"]
    want = orc.circuit_matrix(orc.Circ(n, decode_gates(case["rhs"])))
    assert np.max(np.abs(mats[1] - want)) < 1e-6


def test_product_state_bit_exact():
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 11, 14):
        qs = [orc.qstate(complex(*rng.standard_normal(2)), complex(*rng.standard_normal(2))) for _ in range(n)]
        s = eng.State(n)
        s.set_product(qs)
        assert np.array_equal(s.download(), orc.product_state(qs))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 12, 13])
def test_every_bit_position_every_kernel(n):
    """one op per (kind, target, control placement): exercises the 128-bit,
    scalar and in-vector kernels on every bit."""
    init = random_state(n, 100 + n)
    for t in range(n):
        others = [q for q in range(n) if q != t]
        ctrl_sets = [()]
        if others:
            ctrl_sets += [(others[0],), (others[-1],)]
        if len(others) >= 3:
            ctrl_sets += [(others[0], others[-1]), tuple(others[:3])]
        for cs in ctrl_sets:
            gates = []
            if len(cs) == 0:
                gates = [("RotX", t, 0.37), ("RotZ", t, -1.1), ("Phase", t, 0.9), ("X", t), ("Y", t), ("H", t)]
            elif len(cs) == 1:
                gates = [("CRotY", cs[0], t, 0.77), ("CPhase", cs[0], t, 0.3), ("CX", cs[0], t), ("CRotZ", cs[0], t, 0.5)]
            elif len(cs) == 2:
                gates = [("CCNot", cs[0], cs[1], t), ("Or", cs[0], cs[1], t), ("CSwap", cs[0], cs[1], t)]
            else:
                inner = orc.ncontrol_gate_list(n, list(cs), t, orc.single("RotX", 1.234))
                gates = [("Gate", inner, 0), ("Gate", orc.ncontrol_gate_list(n, list(cs), t, orc.single("Z")), 0)]
            circ = orc.Circ(n, gates)
            want = orc.simulate(init.copy(), circ)
            got = gpu_run(n, init, circ, fuse=None)
            assert orc.rel_l2(got, want) < TOL, (n, t, cs)
    for a in range(n):
        for b in range(n):
            if a != b:
                circ = orc.Circ(n, [("Swap", a, b)])
                assert np.array_equal(gpu_run(n, init, circ, fuse=None), orc.simulate(init.copy(), circ))


@pytest.mark.parametrize("n,gates,seed", [(3, 60, 1), (6, 200, 2), (10, 300, 3), (14, 300, 4), (18, 200, 5), (21, 120, 6)])
@pytest.mark.parametrize("fuse", [False, True])
def test_random_circuits_match_oracle(n, gates, seed, fuse):
    circ = random_circuit(n, gates, seed)
    init = random_state(n, seed + 50)
    want = orc.simulate(init.copy(), circ)
    got = gpu_run(n, init, circ, fuse=fuse)
    assert orc.rel_l2(got, want) < TOL


@pytest.mark.parametrize("flags", [eng.PLAN_GRAPH, eng.PLAN_FUSE | eng.PLAN_GRAPH])
def test_cuda_graph_plans(flags):
    """AQS_PLAN_GRAPH: captured once, replayed, re-captured for another state."""
    n = 11
    circ = random_circuit(n, 120, 31)
    init = random_state(n, 32)
    ops = lower_array(circ)
    plan = eng.Plan(n, ops, flags)
    want = orc.simulate(init.copy(), circ)
    for _ in range(2):                      # second state forces a re-capture
        s = eng.State(n)
        s.upload(init)
        s.run(plan)
        assert orc.rel_l2(s.download(), want) < TOL
        s.run(plan)                         # replay composes like a second simulate()
        assert orc.rel_l2(s.download(), orc.simulate(want.copy(), circ)) < TOL


def test_state_wrap_and_sample_fixed():
    import torch
    n = 12
    a = random_state(n, 71)
    buf = torch.from_numpy(a.copy()).cuda()
    s = eng.State.wrap(n, buf.data_ptr())
    circ = random_circuit(n, 40, 72)
    s.apply_ops(lower_array(circ))
    s.sync()
    want = orc.simulate(a.copy(), circ)
    assert orc.rel_l2(buf.cpu().numpy(), want) < TOL          # the kernels ran on the caller's tensor
    s.upload(want)
    total = s.prob_fixed()
    U = np.array([0, total // 3, total - 1, total, total + 5], dtype=np.uint64)
    got = s.sample_fixed(U)
    F = (orc.probabilities(want).astype(np.float32) * np.float32(2.0 ** 62)).astype(np.uint64)
    cs = np.cumsum(F, dtype=np.uint64)
    exp = np.searchsorted(cs, U, side="right").astype(np.uint64)
    exp[exp == (1 << n)] = np.uint64(0xFFFFFFFFFFFFFFFF)
    assert np.array_equal(got, exp)
    s.close()
    assert orc.rel_l2(buf.cpu().numpy(), want) < 1e-7         # wrap does not own (or free) the memory


@pytest.mark.parametrize("fuse", [False, True])
def test_nested_composites_match_oracle(fuse):
    inner = orc.Circ(2, [("H", 0), ("CRotY", 0, 1, 0.7), ("RotZ", 1, -1.3), ("CY", 1, 0), ("Swap", 0, 1)])
    mid = orc.Circ(4, [("ControlGate", inner, 3, 0), ("RotX", 2, 0.4), ("Or", 0, 1, 3), ("ControlGate", inner, 0, 2)])
    outer = orc.Circ(9, [("H", 5), ("ControlGate", mid, 5, 1), ("Gate", mid, 2), ("ControlGate", mid, 0, 4), ("H", 8)])
    init = random_state(9, 77)
    want = orc.simulate(init.copy(), outer, "dense")       # the reference's own embedding semantics
    got = gpu_run(9, init, outer, fuse=fuse)
    assert orc.rel_l2(got, want) < TOL


@pytest.mark.parametrize("fuse", [False, True])
def test_ghz16_and_histogram(fuse):
    """BASELINE config 1: exact amplitudes and a bit-exact 1000-draw histogram."""
    n = 16
    circ = orc.Circ(n, wl.ghz(n))
    s = eng.State(n)
    s.run(eng.Plan(n, lower_array(circ), eng.PLAN_FUSE if fuse else 0))
    a = s.download()
    assert a[0] == a[-1] == np.complex64(np.float32(0.70710678118)) and np.count_nonzero(a) == 2
    u = np.random.default_rng(16).random(1000, dtype=np.float32)
    want = orc.histogram(orc.simulate(orc.new_state(n), circ), u)
    assert np.array_equal(s.sample_hist(u), want)
    assert np.array_equal(np.bincount(s.sample(u).astype(np.int64), minlength=1 << n).astype(np.uint32), want)


@pytest.mark.parametrize("n", [1, 2, 5, 8, 11, 12, 13, 17, 20])
def test_sampling_and_probabilities_bit_exact(n):
    a = random_state(n, 900 + n)
    s = eng.State(n)
    s.upload(a)
    rng = np.random.default_rng(n)
    u = np.concatenate([rng.random(3000, dtype=np.float32), np.array([0.0, 0.99999994, 0.5], np.float32)])
    assert np.array_equal(s.sample(u), orc.sample(a, u, "exact"))
    assert s.prob_fixed() == orc.prob_fixed(a)
    for q in {0, n // 2, n - 1}:
        assert s.prob_fixed(1 << q, 1 << q) == orc.prob_fixed(a, 1 << (n - 1 - q), 1 << (n - 1 - q))
        assert s.qubit_prob1(q) == orc.qubit_prob1(a, q)
    assert np.array_equal(s.probabilities(), orc.probabilities(a))
    assert abs(s.norm2() - orc.norm2(a)) < 1e-9


def test_sampling_unnormalised_tail_returns_zero():
    """u beyond the total probability: peek_measure_all's 'empty -> 0' rule (quantum.cpp:356)."""
    n = 10
    a = random_state(n, 5) * np.float32(0.5)        # total probability 0.25
    s = eng.State(n)
    s.upload(a)
    u = np.array([0.1, 0.2499, 0.26, 0.9], np.float32)
    got = s.sample(u)
    assert np.array_equal(got, orc.sample(a, u, "exact"))
    assert got[2] == 0 and got[3] == 0


def test_collapse_matches_oracle_bitwise():
    c = KAT["collapse"]
    for n, q in ((c["n"], c["qubit"]), (12, 5), (12, 11), (12, 0)):
        for u in (0.2, 0.8):
            a = random_state(n, 40 + n + q) if n != c["n"] else orc.product_state([qstate_of(x) for x in c["init"]])
            s = eng.State(n)
            s.upload(a)
            p1 = np.float32(s.qubit_prob1(q))
            outcome = bool(np.float32(u) < p1)
            s.collapse_qubit(q, int(outcome), float(p1 if outcome else np.float32(1) - p1))
            assert orc.measure(a, q, u) == outcome
            assert np.array_equal(s.download(), a)


def test_measure_all_collapse_and_clone():
    n = 9
    a = random_state(n, 9)
    s = eng.State(n)
    s.upload(a)
    t = s.clone()
    k = int(s.sample(np.array([0.42], np.float32))[0])
    s.set_basis(k)
    b = s.download()
    assert b[k] == 1 and np.count_nonzero(b) == 1
    assert np.array_equal(t.download(), a)


@pytest.mark.parametrize("fuse", [False, True])
def test_qft_closed_form(fuse):
    """fourier_transform(n)|x> = N^-1/2 exp(2 pi i rev(x) y / N)."""
    n, x = 22, 0b1011000111010010110101
    circ = orc.Circ(n, wl.qft(n))
    s = eng.State(n)
    s.set_basis(x)
    s.run(eng.Plan(n, lower_array(circ), eng.PLAN_FUSE if fuse else 0))
    a = s.download()
    rev = int(format(x, f"0{n}b")[::-1], 2)
    y = np.random.default_rng(1).integers(0, 1 << n, 4096)
    want = np.exp(2j * np.pi * ((rev * y) % (1 << n)) / (1 << n)) / np.sqrt(float(1 << n))
    assert np.max(np.abs(a[y] - want)) * np.sqrt(float(1 << n)) < 2e-4
    assert abs(s.norm2() - 1) < 1e-4


@pytest.mark.parametrize("n", [20, 23])
@pytest.mark.parametrize("fuse", [False, True])
def test_brickwork_matches_oracle(n, fuse):
    """BASELINE config 3 at oracle-sized n: same generator, same seed rule."""
    circ = orc.Circ(n, wl.brickwork(n, depth=20))
    want = orc.simulate(orc.new_state(n), circ)
    s = eng.State(n)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE if fuse else 0)
    s.run(plan)
    assert orc.rel_l2(s.download(), want) < TOL
    u = np.random.default_rng(n).random(2000, dtype=np.float32)
    got = s.sample(u)
    ref = orc.sample(want, u)
    # amplitudes differ in the last bits (FMA vs no FMA) and a bin is only ~2^-n wide, so a draw
    # may land a few bins away; bit-exactness is asserted on identical states elsewhere
    assert np.mean(np.abs(got.astype(np.int64) - ref.astype(np.int64)) > 64) < 0.01
    s2 = eng.State(n)
    s2.upload(want)
    assert np.array_equal(s2.sample(u), ref)


@pytest.mark.parametrize("seed", range(3))
def test_fused_permutation_circuits_are_exact(seed):
    """X / CX / Swap / CCNot / CSwap only move amplitudes: the tile kernel's TK_PERM path must not round them"""
    n = 14
    r = np.random.default_rng(seed)
    gates = []
    for _ in range(80):
        q = [int(x) for x in r.choice(n, 3, replace=False)]
        gates.append([("X", q[0]), ("CX", q[0], q[1]), ("Swap", q[0], q[1]), ("CCNot", q[0], q[1], q[2]),
                      ("CSwap", q[0], q[1], q[2])][int(r.integers(5))])
    circ = orc.Circ(n, gates)
    init = random_state(n, seed)
    assert np.array_equal(gpu_run(n, init, circ, fuse=True), orc.simulate(init.copy(), circ))


@pytest.mark.parametrize("tile_bits", [10, 11, 12])
def test_other_tile_sizes(tile_bits, monkeypatch):
    """the tile kernel is instantiated for 10..13 tile bits; AQS_TILE_BITS selects one at plan time"""
    monkeypatch.setenv("AQS_TILE_BITS", str(tile_bits))
    n = 16
    circ = random_circuit(n, 150, 40 + tile_bits)
    init = random_state(n, tile_bits)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    assert plan.info()["tile_bits"] == tile_bits
    s = eng.State(n)
    s.upload(init)
    s.run(plan)
    assert orc.rel_l2(s.download(), orc.simulate(init.copy(), circ)) < TOL


def test_long_circuit_on_one_tile_splits_passes():
    """n = tile size: everything fits one tile, so passes are cut by the per-pass op budget"""
    n = 12
    circ = random_circuit(n, 1500, 77)
    init = random_state(n, 5)
    plan = eng.Plan(n, lower_array(circ), eng.PLAN_FUSE)
    assert plan.info()["n_fused_passes"] >= 3
    s = eng.State(n)
    s.upload(init)
    s.run(plan)
    assert orc.rel_l2(s.download(), orc.simulate(init.copy(), circ)) < 3 * TOL      # 1500 gates of fp32 rounding


def test_error_paths():
    s = eng.State(4)
    with pytest.raises(eng.EngineError):
        s.apply_ops(eng.op_record(eng.OP_X, 4))
    with pytest.raises(eng.EngineError):
        s.apply_ops(eng.op_record(eng.OP_X, 1, controls=(1,)))
    with pytest.raises(eng.EngineError):
        s.apply_ops(eng.op_record(eng.OP_SWAP, 1, target2=1))
    with pytest.raises(eng.EngineError):
        s.set_basis(16)
    with pytest.raises(eng.EngineError):
        eng.State(0)
    with pytest.raises(eng.EngineError):
        s.run(eng.Plan(5, eng.make_ops(0)))


# ---- sharded pass launches (aqs_plan_run_shard) on ONE GPU ----------------------------------------
# The R ranks of a flat multi-GPU state each run 1/R of the tiles of every pass.  On one GPU the
# ranks can be played one after the other on the same buffer: the union must be the ordinary run.
@pytest.mark.parametrize("n,g,seed", [(14, 1, 1), (15, 2, 2), (16, 3, 3), (18, 3, 4), (20, 2, 5)])
def test_sharded_pass_launches_cover_every_tile_once(n, g, seed):
    circ = random_circuit(n, 120, seed)
    init = random_state(n, seed)
    ops = lower_array(circ)
    plan = eng.Plan(n, ops, eng.PLAN_FUSE)
    n_passes = int(plan.info()["n_fused_passes"])
    assert n_passes > 0
    spans = [plan.pass_span(i, g) for i in range(n_passes)]
    assert any(spans), "the random circuit should put a rank bit into some tile"
    whole = eng.State(n)
    whole.upload(init)
    whole.run(plan)
    want = whole.download()
    s = eng.State(n)
    s.upload(init)
    order = np.random.default_rng(seed).permutation(1 << g)
    for i in range(n_passes):
        for r in order:
            s.run_shard(plan, i, 1, int(r), g)
    got = s.download()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert orc.rel_l2(got, orc.simulate(init.copy(), circ)) < TOL


def test_sampling_many_draws_matches_oracle():
    """10^5 draws on a 2^22 state (2^10 tile sums, one CTA per draw): bit-identical to the oracle on the same state"""
    n = 22
    circ = orc.Circ(n, wl.brickwork(n, 6))
    s = eng.State(n)
    s.run(eng.Plan(n, lower_array(circ), eng.PLAN_FUSE))
    a = s.download()
    u = np.random.default_rng(5).random(100000, dtype=np.float32)
    assert np.array_equal(s.sample(u), orc.sample(a, u, "exact"))
    idx, cnt = s.sample_hist_sparse(u)
    want = np.bincount(orc.sample(a, u, "exact").astype(np.int64), minlength=1 << n)
    assert int(cnt.sum()) == u.size and np.array_equal(want[idx.astype(np.int64)], cnt) and np.count_nonzero(want) == idx.size
    s.close()

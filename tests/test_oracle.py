"""CPU: pins the oracle against the reference's own known-answer vectors
(tests/golden/reference_kat.json, transcribed from test/tests.cpp) and checks
its two composite-gate modes against each other."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import build_circ, check_amp, decode_gates, load_kat, qstate_of

KAT = load_kat()


@pytest.mark.parametrize("case", KAT["cases"], ids=lambda c: c["name"])
@pytest.mark.parametrize("mode", ["flatten", "dense"])
def test_reference_amplitudes(case, mode):
    a = orc.product_state([qstate_of(s) for s in case["init"]])
    orc.simulate(a, orc.Circ(case["n"], decode_gates(case["circuit"])), mode=mode)
    for idx, re, im in case["expect"]:
        assert check_amp(a[idx], re, im, case["tol"]), (case["name"], idx, a[idx], (re, im))


@pytest.mark.parametrize("case", KAT["state_equiv"], ids=lambda c: c["name"])
def test_reference_state_equivalence(case):
    a = orc.product_state([qstate_of(s) for s in case["init"]])
    orc.simulate(a, orc.Circ(case["n"], decode_gates(case["circuit"])))
    ref = orc.product_state([qstate_of(s) for s in case["expect_init"]])
    assert np.max(np.abs(a - ref)) <= case["tol"]


@pytest.mark.parametrize("case", KAT["equiv"], ids=lambda c: c["name"])
@pytest.mark.parametrize("mode", ["flatten", "dense"])
def test_reference_matrix_equivalences(case, mode):
    lhs = orc.circuit_matrix(orc.Circ(case["n"], decode_gates(case["lhs"])), mode=mode)
    rhs = orc.circuit_matrix(orc.Circ(case["n"], decode_gates(case["rhs"])), mode=mode)
    if case["tol"] == 0.0 and mode == "flatten":
        assert np.array_equal(lhs, rhs)          # tests.cpp uses exact == on the matrices
    else:
        assert np.max(np.abs(lhs - rhs)) < max(case["tol"], 1e-6)


@pytest.mark.parametrize("case", KAT["basis_outcomes"], ids=lambda c: c["name"])
def test_reference_basis_outcomes(case):
    a = orc.product_state([qstate_of(s) for s in case["init"]])
    orc.simulate(a, orc.Circ(case["n"], decode_gates(case["circuit"])))
    for u in (0.0, 0.3, 0.999):
        for mode in ("exact", "seq_f32"):
            assert int(orc.sample(a, np.array([u], np.float32), mode)[0]) == case["outcome"]


def test_reference_collapse():
    c = KAT["collapse"]
    for u, key in ((0.1, "if_true"), (0.9, "if_false")):
        a = orc.product_state([qstate_of(s) for s in c["init"]])
        got = orc.measure(a, c["qubit"], u)      # p1 = 1.0 here only if... see below
        exp = c["if_true"] if got else c["if_false"]
        for idx, re, im in exp:
            assert check_amp(a[idx], re, im, c["tol"])
        k = int(orc.sample(a, np.array([0.5], np.float32))[0])
        assert bool(k & (1 << (c["n"] - 1))) == got      # tests.cpp:318-320


def test_reference_collapse_both_branches():
    # qubit 0 of that state is (i, -1)/sqrt2: p1 = 0.5, so u picks the branch
    c = KAT["collapse"]
    seen = set()
    for u in (0.25, 0.75):
        a = orc.product_state([qstate_of(s) for s in c["init"]])
        seen.add(orc.measure(a, c["qubit"], u))
        assert abs(orc.norm2(a) - 1.0) < 1e-6
    assert seen == {True, False}


def test_reference_probabilities():
    for blk in KAT["probabilities"]:
        a = orc.product_state([qstate_of(s) for s in blk["init"]])
        for q, p, tol in blk["p1"]:
            got = np.float32(orc.qubit_prob1(a, q))
            if tol == "exact":
                assert got == np.float32(p)
                assert np.float32(1.0) - got == np.float32(1.0 - p)
            else:
                assert abs(got - p) <= max(abs(got), p) * tol


def test_qstate_normalisation():
    for blk in KAT["qstate"]["normalisation"]:
        z, o = orc.qstate(complex(*blk["args"][0]), complex(*blk["args"][1]))
        (zr, zi), (orr, oi) = blk["expect"]
        assert z == np.complex64(complex(np.float32(zr), np.float32(zi)))
        assert o == np.complex64(complex(np.float32(orr), np.float32(oi)))
    with pytest.raises(ValueError):
        orc.qstate(0, 0)


def test_reference_sampling_statistics():
    """tests.cpp:343-406: every bin inside the 99.9 % normal interval."""
    s = KAT["sampling"]
    rng = np.random.default_rng(343)
    for init in s["inits"]:
        a = orc.product_state([qstate_of(x) for x in init])
        u = rng.random(s["reps"], dtype=np.float32)
        for mode in ("exact", "seq_f32"):
            hist = orc.histogram(a, u, mode)
            assert hist.sum() == s["reps"]
            p = orc.probabilities(a).astype(np.float64)
            sd = np.sqrt(p * (1 - p))
            lo = np.floor(s["reps"] * p) - np.floor(np.sqrt(s["reps"]) * s["z"] * sd)
            hi = np.floor(s["reps"] * p) + np.floor(np.sqrt(s["reps"]) * s["z"] * sd)
            assert np.all((lo <= hist) & (hist <= hi))


def test_sampling_is_searchsorted_right():
    """the sort/countByKey pipeline == searchsorted(cumsum, u, 'right')."""
    rng = np.random.default_rng(5)
    a = (rng.standard_normal(1 << 10) + 1j * rng.standard_normal(1 << 10)).astype(np.complex64)
    a /= np.float32(np.sqrt(orc.norm2(a)))
    u = rng.random(4096, dtype=np.float32)
    got = orc.sample(a, u, "seq_f32")
    cs = np.cumsum(orc.probabilities(a), dtype=np.float32)   # numpy's f32 cumsum is sequential
    want = np.searchsorted(cs, u, side="right")
    want[want == a.size] = 0
    assert np.array_equal(got, want.astype(np.uint64))
    # the exact-sum contract agrees with it except (rarely) at bin boundaries
    ex = orc.sample(a, u, "exact")
    assert np.mean(ex != got) < 0.01
    assert np.all(np.abs(ex.astype(np.int64) - got.astype(np.int64))[ex != got] <= 1)


def test_flatten_equals_dense_random_nest():
    rng = np.random.default_rng(11)
    inner = orc.Circ(2, [("H", 0), ("CRotY", 0, 1, 0.7), ("RotZ", 1, -1.3), ("CY", 1, 0), ("Swap", 0, 1)])
    mid = orc.Circ(4, [("ControlGate", inner, 3, 0), ("RotX", 2, 0.4), ("ControlGate", inner, 0, 2)])
    outer = orc.Circ(6, [("H", 5), ("ControlGate", mid, 5, 1), ("Gate", mid, 2), ("ControlGate", mid, 0, 1)])
    a = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    a /= np.float32(np.linalg.norm(a))
    b = a.copy()
    orc.simulate(a, outer, "flatten")
    orc.simulate(b, outer, "dense")
    assert orc.rel_l2(a, b) < 1e-6


def test_or_under_control_matches_dense():
    inner = orc.Circ(3, [("Or", 0, 1, 2), ("Or", 2, 0, 1)])
    outer = orc.Circ(4, [("H", 0), ("H", 1), ("H", 2), ("ControlGate", inner, 0, 1)])
    a, b = orc.new_state(4), orc.new_state(4)
    orc.simulate(a, outer, "flatten"); orc.simulate(b, outer, "dense")
    assert orc.rel_l2(a, b) < 1e-6


def test_closed_forms():
    """GHZ, QFT of a basis state, Grover amplitudes."""
    n = 10
    ghz = orc.Circ(n, [("H", 0)] + [("CX", i, i + 1) for i in range(n - 1)])
    a = orc.simulate(orc.new_state(n), ghz)
    assert a[0] == a[-1] == np.complex64(np.float32(0.70710678118)) and np.count_nonzero(a) == 2

    n, x = 8, 0b10110001
    a = orc.simulate(orc.new_state(n, x), orc.fourier_transform(n))
    rev = int(format(x, f"0{n}b")[::-1], 2)
    y = np.arange(1 << n)
    want = np.exp(2j * np.pi * rev * y / (1 << n)) / np.sqrt(1 << n)
    assert np.max(np.abs(a - want)) < 2e-6
    b = orc.simulate(a.copy(), orc.inverse_fourier_transform(n))
    # inverse_fourier_transform is the adjoint of fourier_transform
    assert abs(abs(b[x]) - 1) < 1e-5

    n, M, k = 8, 5, 6
    a = orc.simulate(orc.new_state(n), orc.grover_search(n, orc.grover_oracle(n, M), k))
    th = np.arcsin(2.0 ** (-n / 2))
    w = int(format(M, f"0{n}b")[::-1], 2)
    assert abs(a[w] - (-1) ** k * np.sin((2 * k + 1) * th)) < 1e-4
    others = np.delete(a, w)
    assert np.max(np.abs(others - (-1) ** k * np.cos((2 * k + 1) * th) / np.sqrt((1 << n) - 1))) < 1e-4


def test_reference_algorithm_restatement_matches_the_oracle():
    """oracle/ref_algorithm.py (explicit CSR / dense-Kronecker operator per gate, what bench.py times as
    cpu_baseline_ref_algorithm) computes the same states as the matrix-free oracle"""
    from oracle import ref_algorithm as ra
    a = ra.simulate(8, ra.qft(8))
    assert orc.rel_l2(a, orc.simulate(orc.new_state(8), orc.fourier_transform(8))) < 1e-6
    a = ra.simulate(7, ra.grover(7, 5, 3))
    assert orc.rel_l2(a, orc.simulate(orc.new_state(7), orc.grover_search(7, orc.grover_oracle(7, 5), 3))) < 1e-6
    a = ra.simulate(9, ra.ghz(9))
    assert np.array_equal(a, orc.simulate(orc.new_state(9), orc.Circ(9, [("H", 0)] + [("CX", i, i + 1) for i in range(8)])))

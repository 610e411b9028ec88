"""Checks of the aqs:: host layer through its Python mirror (afquantumsim_b200.aqs).

The same functions run in two settings:
  * GPU  (tests/test_gpu_host.py, in-process): host layer -> real CUDA engine;
  * CPU  (tests/test_host_cpu.py, in a SUBPROCESS that preloads oracle/_build/cpu_abi/
    libaqs_engine.so): host layer -> oracle-backed ABI stand-in.  That validates the
    host logic (lowering, composites, string grammar, measurement rules, exceptions)
    where there is no GPU.  The subprocess keeps the stand-in out of any other test.

Run as a script:  python tests/host_cases.py --abi cpu [case ...]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TOL = 1e-5


def _mods():
    from afquantumsim_b200 import aqs
    from oracle import oracle as orc
    from tests import helpers
    return aqs, orc, helpers


def to_aqs(aqs, circ):
    """oracle.Circ -> aqs.QCircuit (recursing through Gate / ControlGate)."""
    qc = aqs.QCircuit(circ.n)
    for g in circ.gates:
        if g[0] == "Gate":
            qc << aqs.Gate(to_aqs(aqs, g[1]), g[2])
        elif g[0] == "ControlGate":
            qc << aqs.ControlGate(to_aqs(aqs, g[1]), g[2], g[3])
        else:
            qc.add(*g)
    return qc


def sim_with(aqs, helpers, init_specs):
    return aqs.QSimulator(len(init_specs), [helpers.qstate_of(s) for s in init_specs])


def case_reference_kat():
    aqs, orc, helpers = _mods()
    kat = helpers.load_kat()
    for case in kat["cases"]:
        for compiled in (True, False):
            qs = sim_with(aqs, helpers, case["init"])
            qc = to_aqs(aqs, orc.Circ(case["n"], helpers.decode_gates(case["circuit"])))
            if compiled:
                qc.compile()
            qs.simulate(qc)
            for idx, re, im in case["expect"]:
                assert helpers.check_amp(qs.state(idx), re, im, case["tol"]), (case["name"], idx, qs.state(idx))
    for case in kat["state_equiv"]:
        qs = sim_with(aqs, helpers, case["init"])
        qs.simulate(to_aqs(aqs, orc.Circ(case["n"], helpers.decode_gates(case["circuit"]))))
        ref = sim_with(aqs, helpers, case["expect_init"])
        assert np.max(np.abs(qs.statevector() - ref.statevector())) <= case["tol"], case["name"]


def _build(aqs, helpers, spec):
    """KAT circuit spec -> aqs.QCircuit using the HOST LAYER's own builders."""
    if "builder" in spec:
        b, a = spec["builder"], spec["args"]
        if b == "single":
            return aqs.single(*a)
        if b == "ncontrol_list":
            return aqs.NControl_Gate(a[0], list(a[1]), a[2], _build(aqs, helpers, a[3]))
        if b == "ncontrol_range":
            return aqs.NControl_Gate(a[0], a[1], a[2], a[3], _build(aqs, helpers, a[4]))
        if b == "control_group":
            return aqs.Control_Group_Gate(a[0], a[1], a[2], _build(aqs, helpers, a[3]))
        if b == "group":
            return aqs.Group_Gate(a[0], a[1], _build(aqs, helpers, a[2]))
        if b == "rewire":
            return aqs.Rewire_Gate(a[0], a[1], _build(aqs, helpers, a[2]))
        if b == "adjoint":
            return aqs.Adjoint_Gate(_build(aqs, helpers, a[0]))
        raise KeyError(b)
    qc = aqs.QCircuit(spec["n"])
    for g in spec["gates"]:
        if g[0] == "Gate":
            qc << aqs.Gate(_build(aqs, helpers, g[1]), g[2])
        elif g[0] == "ControlGate":
            qc << aqs.ControlGate(_build(aqs, helpers, g[1]), g[2], g[3])
        else:
            qc.add(*g)
    return qc


def case_reference_equivalences():
    """tests.cpp:909-961, 1049-1132: composites == primitive sequences, as whole matrices."""
    aqs, orc, helpers = _mods()
    kat = helpers.load_kat()
    for case in kat["equiv"]:
        lhs = _build(aqs, helpers, {"n": case["n"], "gates": case["lhs"]})
        rhs = _build(aqs, helpers, {"n": case["n"], "gates": case["rhs"]})
        lhs.compile(); rhs.compile()
        A, B = lhs.circuit(), rhs.circuit()
        if case["tol"] == 0.0:
            assert np.array_equal(A, B), case["name"]
        else:
            assert np.max(np.abs(A - B)) < case["tol"], case["name"]
        want = orc.circuit_matrix(orc.Circ(case["n"], helpers.decode_gates(case["rhs"])), mode="dense")
        assert np.max(np.abs(B - want)) < 1e-6, case["name"]


def case_reference_measurement():
    aqs, orc, helpers = _mods()
    kat = helpers.load_kat()
    for case in kat["basis_outcomes"]:
        qs = sim_with(aqs, helpers, case["init"])
        qs.simulate(_build(aqs, helpers, {"n": case["n"], "gates": case["circuit"]}))
        assert qs.peek_measure_all() == case["outcome"], case["name"]
        for q in range(case["n"]):
            assert qs.peek_measure(q) == bool(case["outcome"] >> (case["n"] - 1 - q) & 1)
    c = kat["collapse"]
    seen = set()
    for seed in range(8):
        aqs.set_seed(seed)
        qs = sim_with(aqs, helpers, c["init"])
        got = qs.measure(c["qubit"])
        seen.add(got)
        k = qs.peek_measure_all()
        assert bool(k & (1 << (c["n"] - 1))) == got
        for idx, re, im in (c["if_true"] if got else c["if_false"]):
            assert helpers.check_amp(qs.state(idx), re, im, c["tol"])
        m = qs.measure_all()
        assert qs.state(m) == np.complex64(1.0)
        assert abs(qs.norm2() - 1) < 1e-6
    assert seen == {True, False}
    for blk in kat["probabilities"]:
        qs = sim_with(aqs, helpers, blk["init"])
        for q, p, tol in blk["p1"]:
            got = np.float32(qs.qubit_probability_true(q))
            if tol == "exact":
                assert got == np.float32(p) and np.float32(qs.qubit_probability_false(q)) == np.float32(1.0 - p)
            else:
                assert abs(got - p) <= max(abs(got), p) * tol


def case_reference_sampling_statistics():
    """tests.cpp:343-406 (99.9 % interval per bin, 1e4 draws) on the host layer's own RNG."""
    aqs, orc, helpers = _mods()
    s = helpers.load_kat()["sampling"]
    aqs.set_seed(2022)
    for init in s["inits"]:
        qs = sim_with(aqs, helpers, init)
        hist = qs.profile_measure_all(s["reps"])
        assert hist.sum() == s["reps"]
        p = qs.probabilities().astype(np.float64)
        sd = np.sqrt(p * (1 - p))
        half = np.floor(np.sqrt(s["reps"]) * s["z"] * sd)
        assert np.all((np.floor(s["reps"] * p) - half <= hist) & (hist <= np.floor(s["reps"] * p) + half))
        for q in range(s["n"]):
            p1 = qs.qubit_probability_true(q)
            c0, c1 = qs.profile_measure(q, s["reps"])
            assert c0 + c1 == s["reps"]
            half = int(np.sqrt(s["reps"]) * s["z"] * np.sqrt(max(0.0, p1 * (1 - p1))))
            assert int(s["reps"] * p1) - half <= c1 <= int(s["reps"] * p1) + half


def case_representation_strings():
    """the circuit string grammar."""
    aqs, orc, helpers = _mods()
    qc = aqs.QCircuit(5)
    qc << aqs.X(0) << aqs.Y(1) << aqs.Z(2) << aqs.H(3) << aqs.RotX(0, 0.1) << aqs.RotY(1, 0.2) << aqs.RotZ(2, 0.3)
    qc << aqs.Phase(0, aqs.PI / 2) << aqs.Phase(0, -aqs.PI / 2) << aqs.Phase(1, aqs.PI / 4) << aqs.Phase(1, -aqs.PI / 4)
    qc << aqs.Phase(2, 0.5) << aqs.Swap(0, 4) << aqs.CX(0, 1) << aqs.CY(1, 2) << aqs.CZ(2, 3) << aqs.CH(3, 4)
    qc << aqs.CPhase(0, 1, aqs.PI / 2) << aqs.CPhase(0, 1, 0.7) << aqs.CRotX(0, 1, 1.0) << aqs.CRotY(1, 0, 1.0)
    qc << aqs.CRotZ(2, 4, 1.0) << aqs.CSwap(0, 1, 2) << aqs.CCNot(0, 1, 2) << aqs.Or(0, 1, 2)
    qc << aqs.Barrier() << aqs.Barrier(False)
    want = ("X,0,1:0;Y,0,1:1;Z,0,1:2;H,0,1:3;RotX,0,1:0;RotY,0,1:1;RotZ,0,1:2;"
            "S,0,1:0;S†,0,1:0;T,0,1:1;T†,0,1:1;Phase,0,1:2;Swap,0,2:0,4;"
            "X,1,1:0,1;Y,1,1:1,2;Z,1,1:2,3;H,1,1:3,4;S,1,1:0,1;Phase,1,1:0,1;RotX,1,1:0,1;RotY,1,1:1,0;"
            "RotZ,1,1:2,4;Swap,1,2:0,12;X,2,1:0,1,2;"
            "P;X,0,1:0;X,0,1:1;X,0,1:2;X,2,1:0,1,2;X,0,1:0;X,0,1:1;P;B;P;")
    assert qc.representation() == want, qc.representation()
    assert qc.gate_count() == 27
    inner = aqs.QCircuit(2)
    inner << aqs.H(0) << aqs.CX(0, 1) << aqs.Barrier()
    outer = aqs.QCircuit(5)
    outer << aqs.Gate(inner, 2) << aqs.ControlGate(inner, 0, 3) << aqs.Gate(inner, 1, "Bell") << aqs.ControlGate(inner, 4, 1, "CBell")
    assert outer.representation() == "H,0,1:2;X,1,1:2,3;B;H,1,1:0,3;X,2,1:0,3,4;B;Bell,0,2:1,2;CBell,1,2:4,1,2;", outer.representation()
    adj = aqs.Adjoint_Gate(inner)
    assert adj.representation() == "B;X,1,1:0,1;H,0,1:0;"
    qc.clear()
    assert qc.gate_count() == 0 and qc.cached_index() == 0


def case_exceptions():
    """SURVEY Appendix A, last paragraph: same exception classes as the reference."""
    aqs, orc, helpers = _mods()

    def raises(exc, fn):
        try:
            fn()
        except exc:
            return
        raise AssertionError(f"expected {exc.__name__}")

    raises(aqs.InvalidArgument, lambda: aqs.QCircuit(0))
    raises(aqs.InvalidArgument, lambda: aqs.QCircuit(31))
    qc = aqs.QCircuit(3)
    raises(aqs.OutOfRange, lambda: qc << aqs.X(3))
    raises(aqs.OutOfRange, lambda: qc << aqs.CX(0, 3))
    raises(aqs.InvalidArgument, lambda: qc << aqs.CX(1, 1))
    raises(aqs.InvalidArgument, lambda: qc << aqs.Swap(2, 2))
    raises(aqs.InvalidArgument, lambda: qc << aqs.CCNot(0, 1, 1))
    raises(aqs.InvalidArgument, lambda: qc << aqs.CRotX(2, 2, 0.1))
    raises(aqs.OutOfRange, lambda: qc << aqs.CRotX(2, 5, 0.1))
    one = aqs.QCircuit(1)
    raises(aqs.DomainError, lambda: one << aqs.CX(0, 1))
    raises(aqs.DomainError, lambda: one << aqs.Swap(0, 1))
    two = aqs.QCircuit(2)
    raises(aqs.DomainError, lambda: two << aqs.CCNot(0, 1, 2))
    raises(aqs.DomainError, lambda: two << aqs.CSwap(0, 1, 2))
    raises(aqs.OutOfRange, lambda: two << aqs.Gate(qc, 0))
    raises(aqs.OutOfRange, lambda: qc << aqs.Gate(two, 2))
    raises(aqs.OutOfRange, lambda: qc << aqs.ControlGate(two, 1, 0))
    raises(aqs.InvalidArgument, lambda: two << aqs.ControlGate(two, 5, 0))      # bigger gate than the circuit
    raises(aqs.InvalidArgument, lambda: qc << aqs.Gate(two, 0, "a,b"))
    assert qc.gate_count() == 0
    qs = aqs.QSimulator(2)
    raises(aqs.InvalidArgument, lambda: qs.simulate(qc))
    raises(aqs.OutOfRange, lambda: qs.measure(2))
    raises(aqs.OutOfRange, lambda: qs.peek_measure(2))
    raises(aqs.OutOfRange, lambda: qs.qubit_probability_true(2))
    raises(aqs.OutOfRange, lambda: qs.state_probability(4))
    raises(aqs.OutOfRange, lambda: qs.profile_measure(2, 10))
    raises(aqs.InvalidArgument, lambda: aqs.QSimulator(2, [aqs.QState.zero()]))
    raises(aqs.InvalidArgument, lambda: aqs.QSimulator(2, np.zeros(4, np.complex64)))
    raises(aqs.InvalidArgument, lambda: aqs.QState(0, 0))
    raises(aqs.InvalidArgument, lambda: aqs.NControl_Gate(3, [], 0, aqs.single("X")))
    raises(aqs.InvalidArgument, lambda: aqs.NControl_Gate(3, [1], 1, aqs.single("X")))
    raises(aqs.InvalidArgument, lambda: aqs.NControl_Gate(3, 1, 0, 1, aqs.single("X")))
    raises(aqs.InvalidArgument, lambda: aqs.Rewire_Gate(3, [0, 0], two))
    raises(aqs.InvalidArgument, lambda: aqs.grover_oracle(3, 8))
    raises(aqs.InvalidArgument, lambda: aqs.Control_Group_Gate(3, 1, [0, 1], aqs.single("X")))


def _random_circuit(orc, n, n_gates, seed, nested=True):
    rng = np.random.default_rng(seed)
    one = ["X", "Y", "Z", "H", "Phase", "RotX", "RotY", "RotZ"]
    two = ["CX", "CY", "CZ", "CH", "CPhase", "CRotX", "CRotY", "CRotZ", "Swap"]
    three = ["CSwap", "CCNot", "Or"]

    def gates(width, count, depth):
        out = []
        for _ in range(count):
            r = rng.random()
            if nested and depth < 2 and width >= 3 and r < 0.12:
                k = int(rng.integers(1, width))           # inner width 1..width-1
                inner = orc.Circ(k, gates(k, int(rng.integers(1, 5)), depth + 1))
                if rng.random() < 0.5:
                    out.append(("Gate", inner, int(rng.integers(0, width - k + 1))))
                else:
                    begin = int(rng.integers(0, width - k + 1))
                    free = [q for q in range(width) if not begin <= q < begin + k]
                    if free:
                        out.append(("ControlGate", inner, int(rng.choice(free)), begin))
                continue
            if r < 0.5 or width < 2:
                name, q = one[rng.integers(len(one))], [int(rng.integers(width))]
            elif r < 0.88 or width < 3:
                name, q = two[rng.integers(len(two))], [int(x) for x in rng.choice(width, 2, replace=False)]
            else:
                name, q = three[rng.integers(len(three))], [int(x) for x in rng.choice(width, 3, replace=False)]
            if name in orc.HAS_ANGLE:
                out.append((name, *q, float(np.float32(rng.uniform(-np.pi, np.pi)))))
            else:
                out.append((name, *q))
        return out

    return orc.Circ(n, gates(n, n_gates, 0))


def case_random_circuits_vs_oracle(sizes=((3, 40, 1), (5, 120, 2), (8, 200, 3), (11, 150, 4), (14, 100, 5))):
    aqs, orc, helpers = _mods()
    for n, count, seed in sizes:
        circ = _random_circuit(orc, n, count, seed)
        rng = np.random.default_rng(seed)
        init = (rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)).astype(np.complex64)
        want = orc.simulate((init / np.float32(np.sqrt(orc.norm2(init)))).astype(np.complex64), circ,
                            "dense" if n <= 8 else "flatten")
        for fusion in (True, False):
            for compiled in (True, False):
                aqs.set_fusion(fusion)
                qs = aqs.QSimulator(n, init)
                qc = to_aqs(aqs, circ)
                if compiled:
                    qc.compile()
                qs.simulate(qc)
                err = orc.rel_l2(qs.statevector(), want)
                assert err < TOL, (n, seed, fusion, compiled, err)
        aqs.set_fusion(True)


def case_partial_compile_and_chained_simulate():
    """simulate = compiled prefix then uncompiled tail (src/quantum.cpp:283-290); calls compose."""
    aqs, orc, helpers = _mods()
    n = 6
    c1 = _random_circuit(orc, n, 30, 21, nested=False)
    c2 = _random_circuit(orc, n, 30, 22, nested=False)
    qc = to_aqs(aqs, c1)
    qc.compile()
    assert qc.cached_index() == qc.gate_count()
    qc.extend(c2.gates)
    assert qc.cached_index() < qc.gate_count()
    qs = aqs.QSimulator(n)
    qs.simulate(qc)
    qs.simulate(qc)
    want = orc.new_state(n)
    for _ in range(2):
        orc.simulate(want, c1); orc.simulate(want, c2)
    assert orc.rel_l2(qs.statevector(), want) < TOL
    # a copy shares nothing mutable with the original
    cp = qc.copy()
    cp << aqs.X(0)
    cp.compile()
    assert qc.gate_count() + 1 == cp.gate_count()
    qs2 = aqs.QSimulator(n)
    qs2.simulate(qc)
    want = orc.simulate(orc.simulate(orc.new_state(n), c1), c2)
    assert orc.rel_l2(qs2.statevector(), want) < TOL
    # clone is a deep copy of the device state
    t = qs2.clone()
    qs2.measure_all()
    assert orc.rel_l2(t.statevector(), want) < TOL


def case_algorithms_vs_oracle():
    aqs, orc, helpers = _mods()
    for n in (1, 2, 5, 9):
        x = (0b101101011 & ((1 << n) - 1))
        init = orc.new_state(n, x)
        qs = aqs.QSimulator(n, init)
        qs.simulate(aqs.fourier_transform(n))
        want = orc.simulate(init.copy(), orc.fourier_transform(n))
        assert orc.rel_l2(qs.statevector(), want) < TOL
        qs.simulate(aqs.inverse_fourier_transform(n))
        assert abs(abs(qs.state(x)) - 1) < 1e-5
        assert aqs.fourier_transform(n).gate_count() == n + n * (n - 1) // 2
    for n, marked, iters in ((3, 5, 2), (6, 37, 6), (10, 5, 25)):
        oracle_c = aqs.grover_oracle(n, marked)
        qc = aqs.QCircuit(n)
        qc << aqs.Gate(aqs.grover_search(n, oracle_c, iters, "Oracle"), 0)    # examples/grover_search.cpp:31-40
        qc.compile()
        qs = aqs.QSimulator(n)
        qs.simulate(qc)
        want = orc.simulate(orc.new_state(n), orc.grover_search(n, orc.grover_oracle(n, marked), iters))
        assert orc.rel_l2(qs.statevector(), want) < TOL
        w = int(format(marked, f"0{n}b")[::-1], 2)
        th = np.arcsin(2.0 ** (-n / 2))
        assert abs(qs.state_probability(w) - np.sin((2 * iters + 1) * th) ** 2) < 1e-4
        gi = aqs.QSimulator(n, aqs.QState.plus())
        gi.simulate(aqs.grover_iteration(n, oracle_c, iters))
        assert orc.rel_l2(gi.statevector(), want) < 1e-4


def case_lowering_agrees_with_test_lowering():
    """two independent flatteners (C++ host layer, tests/lowering.py) emit the same ops."""
    aqs, orc, helpers = _mods()
    from tests.lowering import lower_array
    for seed in range(6):
        circ = _random_circuit(orc, 7, 60, 100 + seed)
        a, b = to_aqs(aqs, circ).ops(), lower_array(circ)
        assert len(a) == len(b)
        for f in ("kind", "target", "target2", "ctrl_mask", "ctrl_value"):
            assert np.array_equal(a[f], b[f]), f
        assert np.allclose(a["m"], b["m"], rtol=0, atol=2e-7)     # numpy vs libm trig: <= 1 ulp apart


def case_set_basis():
    aqs, orc, helpers = _mods()
    n = 3
    h = np.float32(0.70710678118)
    qs = aqs.QSimulator(n, [aqs.QState(1, 1), aqs.QState(1, -1), aqs.QState(1, 1)])
    qs.set_basis(aqs.QSimulator.X)                 # |+-+> in the X basis is |010>
    assert qs.get_basis() == aqs.QSimulator.X and qs.peek_measure_all() == 0b010
    qs.set_basis(aqs.QSimulator.Z)
    assert abs(qs.state(0) - h * h * h) < 1e-6
    # Y basis: the reference's matrix is h*[[1, 1], [-i, i]] (src/quantum.cpp:428-435, column-major + .T())
    rng = np.random.default_rng(8)
    init = (rng.standard_normal(8) + 1j * rng.standard_normal(8)).astype(np.complex64)
    qs = aqs.QSimulator(n, init)
    start = qs.statevector()
    qs.set_basis(aqs.QSimulator.Y)
    zy = np.array([[h, h], [-1j * h, 1j * h]], dtype=np.complex64)
    want = np.kron(np.kron(zy, zy), zy) @ start
    assert np.max(np.abs(qs.statevector() - want)) < 1e-6
    qs.set_basis(aqs.QSimulator.X)                 # Y -> X goes back through Z with the Y matrix (reference fall-through fixed)
    zx = np.array([[h, h], [h, -h]], dtype=np.complex64)
    want = np.kron(np.kron(zx, zx), zx) @ start
    assert np.max(np.abs(qs.statevector() - want)) < 1e-5
    qs.set_basis(aqs.QSimulator.Z)
    assert np.max(np.abs(qs.statevector() - start)) < 1e-5


def case_text_renderer():
    aqs, orc, helpers = _mods()
    golden = ("\n"
              "     ┌───┐           \n"
              "|1⟩──┤ H ├──────█────\n"
              "     └───┘      │    \n"
              "              ┌─┴─┐  \n"
              "|1⟩───────────┤ X ├──\n"
              "              └───┘  \n")                      # docs/USAGE.md:533-540
    qc = aqs.QCircuit(2)
    qc << aqs.H(0) << aqs.CX(0, 1)
    qs = aqs.QSimulator(2, aqs.QState.one())
    assert aqs.gen_circuit_text_image(qc, qs) == golden
    assert aqs.gen_circuit_text_image("2; 0,1; 1,1; H,0,1: 0; X,1,1: 0 , 1;") == golden   # README.md:44-66
    img = aqs.gen_circuit_text_image("3;0,0;1,0;2,0;Swap,1,2:1,0,2;B;U,0,2:0,1;X,1,1:2,0;")
    lines = img.split("\n")
    assert len(lines) == 3 * 3 + 2 and len({len(l) for l in lines[1:-1]}) == 1
    assert "─╳─" in lines[2] and "─█─" in lines[5] and "─╳─" in lines[8] and "▒" in lines[2]
    for bad, exc in (("0;", aqs.OutOfRange), ("31;", aqs.OutOfRange), ("2;0,0;0,1;", aqs.InvalidArgument),
                     ("2;0,0;1,0;X,0,1:2;", aqs.OutOfRange), ("2;0,0;1,0;X,1,1:0,0;", aqs.InvalidArgument),
                     ("2;0,0;1,0;X,1,1:0;", aqs.InvalidArgument), ("2;0,0;1,0;X,1,0:0;", aqs.InvalidArgument),
                     ("3;0,0;1,0;2,0;Swap,0,3:0,1,2;", aqs.InvalidArgument), ("2;0,0;1,0;garbage;", aqs.InvalidArgument)):
        try:
            aqs.gen_circuit_text_image(bad)
        except exc:
            continue
        raise AssertionError(bad)
    sup = aqs.QSimulator(2, aqs.QState.plus())
    try:
        aqs.gen_circuit_text_image(qc, sup)
    except aqs.InvalidArgument:
        pass
    else:
        raise AssertionError("superposed initial states must be refused")


def case_ghz16_profile():
    """BASELINE config 1 through the public API."""
    aqs, orc, helpers = _mods()
    from afquantumsim_b200 import workloads as wl
    n = 16
    qc = aqs.QCircuit(n).extend(wl.ghz(n))
    qs = aqs.QSimulator(n)
    qs.simulate(qc)
    hist = qs.profile_measure_all(1000)
    assert hist[0] + hist[-1] == 1000 and 400 < hist[0] < 600
    assert qs.state(0) == qs.state((1 << n) - 1) == np.complex64(np.float32(0.70710678118))
    u = np.random.default_rng(1).random(1000, dtype=np.float32)
    want = orc.sample(orc.simulate(orc.new_state(n), orc.Circ(n, wl.ghz(n))), u)
    assert np.array_equal(qs.sample(u), want)


def case_opaque_matrices_write_through():
    """a user-written circuit matrix (reference: `qc.circuit() = M`, docs/USAGE.md:121-125) inside Gate / ControlGate,
    compiled and uncompiled, against numpy on the oracle's states (reference src/quantum.cpp:1760-1814, 1888-1950)"""
    from tests.dense_cases import dense_reference, random_unitary
    aqs, orc, helpers = _mods()
    rng = np.random.default_rng(77)
    for n, k, b, ctrl in ((7, 3, 2, 0), (9, 5, 1, 8), (13, 6, 4, 1)):
        U1, U2 = random_unitary(k, rng), random_unitary(k, rng)
        pre = [("H", q) for q in range(n)] + [("CX", q, q + 1) for q in range(0, n - 1, 2)]
        mid = [("RotY", q, 0.3 + 0.1 * q) for q in range(n)]
        a = orc.simulate(orc.new_state(n), orc.Circ(n, pre))
        a = dense_reference(a, n, list(range(b, b + k)), U1)
        a = orc.simulate(a, orc.Circ(n, mid))
        want = dense_reference(a, n, list(range(b, b + k)), U2, [ctrl])
        for compiled in (False, True):
            in1 = aqs.QCircuit(k).set_matrix(U1)
            in2 = aqs.QCircuit(k).set_matrix(U2)
            qc = aqs.QCircuit(n).extend(pre)
            qc << aqs.Gate(in1, b)
            qc.extend(mid)
            qc << aqs.ControlGate(in2, ctrl, b)
            if compiled:
                qc.compile()
            qs = aqs.QSimulator(n)
            qs.simulate(qc)
            err = orc.rel_l2(qs.statevector(), want)
            assert err < 1e-5, (n, k, compiled, err)
    # the opaque circuit simulated directly, and the errors of misuse
    U = random_unitary(2, rng)
    qs = aqs.QSimulator(2)
    qs.simulate(aqs.QCircuit(2).set_matrix(U))
    assert np.allclose(qs.statevector(), U[:, 0], atol=1e-6)
    with pytest_raises(aqs.InvalidArgument):
        aqs.QCircuit(2).set_matrix(np.eye(8))
    with pytest_raises(aqs.EngineFailure):
        (aqs.QCircuit(2) << aqs.H(0)).set_matrix(np.eye(4))


class pytest_raises:
    def __init__(self, exc):
        self.exc = exc

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        assert et is not None and issubclass(et, self.exc), f"expected {self.exc.__name__}, got {et}"
        return True


ALL = [v for k, v in sorted(globals().items()) if k.startswith("case_")]


def preload_cpu_abi():
    path = os.path.join(ROOT, "oracle", "_build", "cpu_abi", "libaqs_engine.so")
    ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args[:2] == ["--abi", "cpu"]:
        preload_cpu_abi()
        args = args[2:]
    names = args or [f.__name__ for f in ALL]
    for name in names:
        globals()[name]()
        print("ok", name, flush=True)

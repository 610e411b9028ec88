#!/usr/bin/env python
"""Transcribes the known-answer vectors of the reference's own test suite
(/root/reference/test/tests.cpp) into tests/golden/reference_kat.json.

The reference cannot be executed here (ArrayFire is absent), so these are the
constants its asserts hold, copied as DATA with the source line of each block.
Angles the reference writes as float expressions (std::atan(0.75f), aqs::pi/2.f)
are evaluated here in float32, as the C++ would.

Run:  python tests/golden/make_reference_kat.py   (needs nothing but numpy)
"""
import json
import os

import numpy as np

f32 = np.float32
PI = f32(3.14159265358979323846)  # aqs::pi


def fl(x):
    return float(f32(x))


ZERO, ONE, PLUS, MINUS = "zero", "one", "plus", "minus"
Q68 = [[0.6, 0.0], [0.0, 0.8]]      # QState({0.6f,0.0f},{0.0f,0.8f})
Q86 = [[0.0, 0.8], [0.6, 0.0]]      # QState({0.0f,0.8f},{0.6f,0.0f})
S8 = 0.3535533906                   # invsqrt8, tests.cpp:747

ATAN075 = fl(np.arctan(f32(0.75)))
ATAN1 = fl(np.arctan(f32(1.0)))
ACOS06x2 = fl(f32(np.arccos(f32(0.6))) * f32(2.0))
PI2 = fl(PI / f32(2.0))
PI3 = fl(PI / f32(3.0))
PI4 = fl(PI / f32(4.0))

cases = []


def case(name, source, init, circuit, expect, tol, n=4):
    cases.append(dict(name=name, source="test/tests.cpp:" + source, n=n, init=init,
                      circuit=circuit, expect=expect, tol=tol))


# --- test_qsim_gates (tests.cpp:414-1005), 4 qubits ------------------------
case("X", "424-431", [ZERO, ZERO, ZERO, ONE], [["X", 2], ["X", 3]], [[0b0010, 1.0, 0.0]], "exact")
case("Y", "441-451", [ZERO, ZERO, ZERO, ONE], [["Y", 2], ["Y", 3]], [[0b0010, 1.0, 0.0]], "exact")
case("Z", "461-471", [ZERO, ZERO, ZERO, ONE], [["Z", 2], ["Z", 3]], [[0b0001, -1.0, 0.0]], "exact")
case("Not", "481-491", [ZERO, ZERO, ZERO, ONE], [["Not", 2], ["Not", 3]], [[0b0010, 1.0, 0.0]], "exact")
case("H", "493-507", [ZERO, ZERO, ZERO, ONE], [["H", 2], ["H", 3]],
     [[0, 0.5, 0.0], [1, -0.5, 0.0], [2, 0.5, 0.0], [3, -0.5, 0.0]], 1e-4)
case("Phase", "521-535", [ZERO, ZERO, Q68, ZERO], [["Phase", 2, ATAN075], ["Phase", 3, ATAN1]],
     [[0, 0.6, 0.0], [1, 0.0, 0.0], [2, -0.48, 0.64], [3, 0.0, 0.0]], 1e-4)
case("RotX", "549-565", [ZERO, ZERO, MINUS, ZERO], [["RotX", 2, PI2]],
     [[0, 0.5, 0.5], [1, 0.0, 0.0], [2, -0.5, -0.5], [3, 0.0, 0.0]], 1e-4)
case("RotY", "567-583", [ZERO, ZERO, ONE, ZERO], [["RotY", 2, ACOS06x2]],
     [[0, -0.8, 0.0], [1, 0.0, 0.0], [2, 0.6, 0.0], [3, 0.0, 0.0]], 1e-4)
case("RotZ", "585-601", [ZERO, ZERO, MINUS, ZERO], [["RotZ", 2, -PI2]],
     [[0, 0.5, 0.5], [1, 0.0, 0.0], [2, -0.5, 0.5], [3, 0.0, 0.0]], 1e-4)
XOR_EXPECT = [[0b0100, 0.0, 0.48], [0b0101, 0.36, 0.0], [0b0110, -0.64, 0.0], [0b0111, 0.0, 0.48]]
case("Xor", "603-619", [ZERO, ONE, Q68, Q68], [["Xor", 0, 2], ["Xor", 1, 3]], XOR_EXPECT, 1e-4)
case("Swap", "621-637", [ZERO, ONE, Q68, Q86], [["Swap", 0, 2], ["Swap", 1, 3]],
     [[0b0001, 0.0, 0.48], [0b0101, 0.36, 0.0], [0b1001, -0.64, 0.0], [0b1101, 0.0, 0.48]], 1e-4)
case("CNot", "639-655", [ZERO, ONE, Q68, Q68], [["CNot", 0, 2], ["CNot", 1, 3]], XOR_EXPECT, 1e-4)
case("CX", "657-672", [[[1.0, 0.0], [0.0, 0.0]], [[0.0, 0.0], [1.0, 0.0]], Q68, Q68],
     [["CX", 0, 2], ["CX", 1, 3]], XOR_EXPECT, 1e-4)
case("CY", "687-703", [ZERO, ONE, Q68, Q68], [["CY", 0, 2], ["CY", 1, 3]],
     [[0b0100, 0.48, 0.0], [0b0101, 0.0, 0.36], [0b0110, 0.0, 0.64], [0b0111, -0.48, 0.0]], 1e-4)
case("CZ", "705-720", [ZERO, ONE, Q68, Q68], [["CZ", 0, 2], ["CZ", 1, 3]],
     [[0b0100, 0.36, 0.0], [0b0101, 0.0, -0.48], [0b0110, 0.0, 0.48], [0b0111, 0.64, 0.0]], 1e-4)
case("CPhase", "722-738", [ZERO, ONE, Q68, Q68], [["CPhase", 0, 2, ATAN075], ["CPhase", 1, 3, ATAN075]],
     [[0b0100, 0.36, 0.0], [0b0101, -0.288, 0.384], [0b0110, 0.0, 0.48], [0b0111, -0.512, -0.384]], 1e-4)
case("CRotX", "740-760", [ZERO, ONE, MINUS, MINUS], [["CRotX", 0, 2, PI2], ["CRotX", 1, 3, PI2]],
     [[0b0100, S8, S8], [0b0101, -S8, -S8], [0b0110, -S8, -S8], [0b0111, S8, S8]], 1e-4)
case("CRotY", "762-779", [ZERO, ONE, ONE, ONE], [["CRotY", 0, 2, ACOS06x2], ["CRotY", 1, 3, ACOS06x2]],
     [[0b0100, 0.0, 0.0], [0b0101, 0.0, 0.0], [0b0110, -0.8, 0.0], [0b0111, 0.6, 0.0]], 1e-4)
case("CRotZ", "781-798", [ZERO, ONE, MINUS, MINUS], [["CRotZ", 0, 2, -PI2], ["CRotZ", 1, 3, -PI2]],
     [[0b0100, S8, S8], [0b0101, -S8, S8], [0b0110, -S8, -S8], [0b0111, S8, -S8]], 1e-4)
OR = [["Or", 0, 1, 2], ["Or", 1, 0, 3]]
case("Or-01", "800-812", [ZERO, ONE, ZERO, ZERO], OR, [[0b0111, 1.0, 0.0]], 1e-4)
case("Or-11", "814-818", [ONE, ONE, ZERO, ZERO], OR, [[0b1111, 1.0, 0.0]], 1e-4)
case("Or-00", "820-825", [ZERO, ZERO, ZERO, ZERO], OR, [[0b0000, 1.0, 0.0]], 1e-4)
AND = [["And", 0, 1, 2], ["And", 1, 0, 3]]
case("And-01", "827-839", [ZERO, ONE, ZERO, ZERO], AND, [[0b0100, 1.0, 0.0]], 1e-4)
case("And-11", "841-845", [ONE, ONE, ZERO, ZERO], AND, [[0b1111, 1.0, 0.0]], 1e-4)
case("And-00", "847-852", [ZERO, ZERO, ZERO, ZERO], AND, [[0b0000, 1.0, 0.0]], 1e-4)
case("CSwap-ctrl0", "854-869", [ZERO, ONE, Q68, Q86], [["CSwap", 0, 2, 3]],
     [[0b0100, 0.0, 0.48], [0b0101, 0.36, 0.0], [0b0110, -0.64, 0.0], [0b0111, 0.0, 0.48]], 1e-4)
case("CSwap-ctrl1", "871-880", [ZERO, ONE, Q68, Q86], [["CSwap", 1, 2, 3]],
     [[0b0100, 0.0, 0.48], [0b0101, -0.64, 0.0], [0b0110, 0.36, 0.0], [0b0111, 0.0, 0.48]], 1e-4)
CCN = [["CCNot", 0, 1, 2], ["CCNot", 1, 0, 3]]
case("CCNot-01", "882-894", [ZERO, ONE, ZERO, ZERO], CCN, [[0b0100, 1.0, 0.0]], 1e-4)
case("CCNot-11", "896-900", [ONE, ONE, ZERO, ZERO], CCN, [[0b1111, 1.0, 0.0]], 1e-4)
case("CCNot-00", "902-907", [ZERO, ZERO, ZERO, ZERO], CCN, [[0b0000, 1.0, 0.0]], 1e-4)

# CHadamard (tests.cpp:963-1002): the result must equal a product state exactly
RS2 = fl(np.sqrt(f32(1.0) / f32(2.0)))
state_equiv = [
    dict(name="CH-inactive", source="test/tests.cpp:963-984", n=4,
         init=[ZERO, ZERO, ONE, ZERO], circuit=[["CH", 0, 1], ["CH", 3, 2]],
         expect_init=[ZERO, ZERO, ONE, ZERO], tol=0.0),
    dict(name="CH-active", source="test/tests.cpp:986-1002", n=4,
         init=[ONE, ZERO, ONE, ONE], circuit=[["CH", 0, 1], ["CH", 3, 2]],
         expect_init=[ONE, [[RS2, 0.0], [RS2, 0.0]], [[RS2, 0.0], [-RS2, 0.0]], ONE], tol=0.0),
]

# --- composite equivalences, compared as whole circuit matrices ------------
TEMP_A = {"n": 2, "gates": [["H", 0], ["CNot", 0, 1], ["Z", 0], ["Swap", 0, 1]]}
TEMP_B = {"n": 2, "gates": [["H", 0], ["CPhase", 0, 1, 1.0], ["X", 1], ["Swap", 0, 1]]}
XG = {"builder": "single", "args": ["X"]}
equiv = [
    dict(name="ControlGate-above", source="test/tests.cpp:909-928", n=4, tol=0.0,
         lhs=[["ControlGate", TEMP_A, 0, 2]],
         rhs=[["CH", 0, 2], ["CCNot", 0, 2, 3], ["CZ", 0, 2], ["CSwap", 0, 2, 3]]),
    dict(name="ControlGate-below", source="test/tests.cpp:930-941", n=4, tol=0.0,
         lhs=[["ControlGate", TEMP_A, 2, 0]],
         rhs=[["CH", 2, 0], ["CCNot", 0, 2, 1], ["CZ", 2, 0], ["CSwap", 2, 0, 1]]),
    dict(name="Gate-offset", source="test/tests.cpp:943-961", n=4, tol=0.0,
         lhs=[["Gate", TEMP_B, 1]],
         rhs=[["H", 1], ["CPhase", 1, 2, 1.0], ["X", 2], ["Swap", 1, 2]]),
    dict(name="NControl-1", source="test/tests.cpp:1049-1059", n=4, tol=0.0,
         lhs=[["Gate", {"builder": "ncontrol_list", "args": [4, [3], 0, XG]}, 0]],
         rhs=[["CX", 3, 0]]),
    dict(name="NControl-2", source="test/tests.cpp:1061-1070", n=4, tol=0.0,
         lhs=[["Gate", {"builder": "ncontrol_list", "args": [4, [0, 3], 2, XG]}, 0]],
         rhs=[["CCNot", 0, 3, 2]]),
    dict(name="Control_Group_Gate", source="test/tests.cpp:1073-1084", n=4, tol=0.0,
         lhs=[["Gate", {"builder": "control_group", "args": [4, 1, [0, 2, 3], XG]}, 0]],
         rhs=[["CX", 1, 0], ["CX", 1, 2], ["CX", 1, 3]]),
    dict(name="Group_Gate", source="test/tests.cpp:1086-1094", n=4, tol=0.0,
         lhs=[["Gate", {"builder": "group", "args": [4, [0, 2, 3], XG]}, 0]],
         rhs=[["X", 0], ["X", 2], ["X", 3]]),
    dict(name="Rewire_Gate", source="test/tests.cpp:1096-1108", n=4, tol=0.0,
         lhs=[["Gate", {"builder": "rewire", "args": [4, [2, 1, 3, 0],
                        {"n": 4, "gates": [["H", 0], ["CX", 0, 2], ["H", 1], ["Z", 3]]}]}, 0]],
         rhs=[["H", 2], ["CX", 2, 3], ["H", 1], ["Z", 0]]),
    dict(name="Adjoint_Gate", source="test/tests.cpp:1110-1132", n=4, tol=1e-5,
         lhs=[["Gate", {"builder": "adjoint", "args": [
             {"n": 4, "gates": [["Swap", 1, 2], ["CPhase", 2, 1, -PI3], ["CRotX", 1, 3, PI4],
                                ["X", 0], ["CRotY", 3, 0, -PI2]]}]}, 0]],
         rhs=[["CRotY", 3, 0, PI2], ["X", 0], ["CRotX", 1, 3, -PI4], ["CPhase", 2, 1, PI3], ["Swap", 1, 2]]),
]

# --- NControl_Gate with non-contiguous controls (tests.cpp:1015-1047) ------
NC6 = [["Gate", {"builder": "ncontrol_list", "args": [6, [0, 2, 4, 5], 3, XG]}, 0]]
basis_outcomes = [
    dict(name="NControl6-all-one", source="test/tests.cpp:1021-1029", n=6,
         init=[ONE] * 6, circuit=NC6, outcome=0b111011),
    dict(name="NControl6-target-zero", source="test/tests.cpp:1031-1034", n=6,
         init=[ONE, ONE, ONE, ZERO, ONE, ONE], circuit=NC6, outcome=0b111111),
    dict(name="NControl6-q1-zero-t0", source="test/tests.cpp:1036-1040", n=6,
         init=[ONE, ZERO, ONE, ZERO, ONE, ONE], circuit=NC6, outcome=0b101111),
    dict(name="NControl6-q1-zero-t1", source="test/tests.cpp:1042-1046", n=6,
         init=[ONE, ZERO, ONE, ONE, ONE, ONE], circuit=NC6, outcome=0b101011),
    dict(name="measure-zero", source="test/tests.cpp:279-288", n=3, init=[ZERO] * 3, circuit=[], outcome=0),
    dict(name="measure-101", source="test/tests.cpp:290-301", n=3, init=[ONE, ZERO, ONE], circuit=[], outcome=0b101),
]

# --- measure(0) collapse (tests.cpp:313-332) --------------------------------
QI = [[0.0, 1.0], [-1.0, 0.0]]   # QState({0,1},{-1,0})
collapse = dict(
    name="measure-collapse", source="test/tests.cpp:313-332", n=3, init=[QI, ZERO, Q68], qubit=0, tol=1e-4,
    if_true=[[0b100, -0.6, 0.0], [0b101, 0.0, -0.8]],
    if_false=[[0b000, 0.0, 0.6], [0b001, -0.8, 0.0]],
)

# --- qubit probabilities (tests.cpp:1141-1194), 4 qubits --------------------
probabilities = [
    dict(source="test/tests.cpp:1148-1153", init=[ZERO] * 4, p1=[[0, 0.0, "exact"], [1, 0.0, "exact"], [2, 0.0, "exact"], [3, 0.0, "exact"]]),
    dict(source="test/tests.cpp:1155-1165", init=[ZERO, ZERO, ZERO, [[0.0, 0.0], [1.0, 0.0]]],
         p1=[[0, 0.0, "exact"], [1, 0.0, "exact"], [2, 0.0, "exact"], [3, 1.0, "exact"]]),
    dict(source="test/tests.cpp:1167-1177", init=[ZERO, ZERO, ZERO, [[1.0, 0.0], [1.0, 0.0]]],
         p1=[[0, 0.0, "exact"], [1, 0.0, "exact"], [2, 0.0, "exact"], [3, 0.5, 1e-4]]),
    dict(source="test/tests.cpp:1179-1191", init=[Q68, ZERO, ZERO, [[1.0, 0.0], [1.0, 0.0]]],
         p1=[[0, 0.64, 1e-4], [1, 0.0, "exact"], [2, 0.0, "exact"], [3, 0.5, 1e-4]]),
]

# --- QState (tests.cpp:42-245) ----------------------------------------------
qstate = dict(
    source="test/tests.cpp:42-157",
    normalisation=[
        dict(args=[[0.6, 0.0], [0.0, 0.8]], expect=[[0.6, 0.0], [0.0, 0.8]], tol="exact"),
        dict(args=[[3.0, 0.0], [4.0, 0.0]], expect=[[0.6, 0.0], [0.8, 0.0]], tol="exact"),
    ],
)

# --- statistical sampling check (tests.cpp:343-406): 1e4 draws, 99.9 % CI ---
sampling = dict(
    source="test/tests.cpp:343-406", n=5, reps=10000, z=3.291,
    inits=[
        [ZERO, ZERO, ZERO, ONE, ZERO],
        [[[1.0, 0.0], [1.0, 0.0]], [[-2.0, 0.0], [3.0, 0.0]], ZERO, ONE, [[0.0, 0.6], [0.8, 0.0]]],
    ],
)

out = dict(
    _about="Known-answer vectors transcribed from the reference's test/tests.cpp by make_reference_kat.py",
    fequal="|a-b| <= max(|a|,|b|) * tol, per component (tests.cpp:19-27); tol 'exact' means ==",
    cases=cases, state_equiv=state_equiv, equiv=equiv, basis_outcomes=basis_outcomes,
    collapse=collapse, probabilities=probabilities, qstate=qstate, sampling=sampling,
)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", path, "-", len(cases), "amplitude cases,", len(equiv), "matrix equivalences")

"""Shared test helpers: golden-vector loading and circuit-description decoding."""
import json
import os

import numpy as np

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def load_kat():
    with open(os.path.join(HERE, "golden", "reference_kat.json")) as fh:
        return json.load(fh)


_NAMED = {
    "zero": (1.0, 0.0),            # QState::zero(), include/quantum.h:300-303
    "one": (0.0, 1.0),
    "plus": (0.70710678118, 0.70710678118),
    "minus": (0.70710678118, -0.70710678118),
}


def qstate_of(spec):
    """Decode a qubit-state spec of the KAT file into a normalised (z, o) pair."""
    if isinstance(spec, str):
        z, o = _NAMED[spec]
        return orc.qstate(np.float32(z), np.float32(o))
    (zr, zi), (orr, oi) = spec
    return orc.qstate(complex(np.float32(zr), np.float32(zi)), complex(np.float32(orr), np.float32(oi)))


def build_circ(spec):
    """Decode {'n','gates'} or {'builder','args'} into an oracle Circ."""
    if "builder" in spec:
        b, args = spec["builder"], spec["args"]
        if b == "single":
            return orc.single(*args)
        if b == "ncontrol_list":
            return orc.ncontrol_gate_list(args[0], args[1], args[2], build_circ(args[3]))
        if b == "ncontrol_range":
            return orc.ncontrol_gate_range(args[0], args[1], args[2], args[3], build_circ(args[4]))
        if b == "control_group":
            return orc.control_group_gate(args[0], args[1], args[2], build_circ(args[3]))
        if b == "group":
            return orc.group_gate(args[0], args[1], build_circ(args[2]))
        if b == "rewire":
            return orc.rewire_gate(args[0], args[1], build_circ(args[2]))
        if b == "adjoint":
            return orc.adjoint_gate(build_circ(args[0]))
        raise KeyError(b)
    return orc.Circ(spec["n"], decode_gates(spec["gates"]))


def decode_gates(gates):
    out = []
    for g in gates:
        if g[0] in ("Gate", "ControlGate"):
            out.append((g[0], build_circ(g[1])) + tuple(g[2:]))
        else:
            out.append(tuple(g))
    return out


def fequal(a, b, tol):
    """tests.cpp:19-22"""
    a = np.float32(a); b = np.float32(b)
    return abs(a - b) <= max(abs(a), abs(b)) * np.float32(tol)


def check_amp(got, re, im, tol):
    if tol == "exact":
        return np.float32(got.real) == np.float32(re) and np.float32(got.imag) == np.float32(im)
    return fequal(got.real, re, tol) and fequal(got.imag, im, tol)

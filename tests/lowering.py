"""Test-side lowering of circuit descriptions (oracle.Circ tuples) to aqs_op
records, written independently of the C++ host layer so that ABI-level GPU
tests do not depend on it.  Matrices follow SURVEY.md Appendix A."""
import math

import numpy as np

from afquantumsim_b200 import engine as eng
from oracle.oracle import ALIASES, Circ

H = np.float32(0.70710678118)


def _rot(name, angle):
    a = np.float32(angle)
    if name.endswith("Phase"):
        return eng.OP_DIAG, [1, 0, 0, complex(np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32))]
    c = np.cos(a / np.float32(2), dtype=np.float32)
    s = np.sin(a / np.float32(2), dtype=np.float32)
    if name.endswith("RotX"):
        return eng.OP_U2, [c, complex(0, -s), complex(0, -s), c]
    if name.endswith("RotY"):
        return eng.OP_U2, [c, -s, s, c]
    return eng.OP_DIAG, [complex(c, -s), 0, 0, complex(c, s)]


def lower(circ: Circ, offset=0, ctrls=()):
    """-> list of op records (numpy, dtype eng.OP_DTYPE)."""
    out = []
    for g in circ.gates:
        name = ALIASES.get(g[0], g[0])
        a = g[1:]
        q = lambda i: a[i] + offset
        if name == "Barrier":
            continue
        if name == "Gate":
            out += lower(a[0], offset + a[1], ctrls)
        elif name == "ControlGate":
            out += lower(a[0], offset + a[2], tuple(ctrls) + (q(1),))
        elif name == "X":
            out.append(eng.op_record(eng.OP_X, q(0), controls=ctrls))
        elif name == "Y":
            out.append(eng.op_record(eng.OP_U2, q(0), [0, -1j, 1j, 0], ctrls))
        elif name == "Z":
            out.append(eng.op_record(eng.OP_DIAG, q(0), [1, 0, 0, -1], ctrls))
        elif name == "H":
            out.append(eng.op_record(eng.OP_U2, q(0), [H, H, H, -H], ctrls))
        elif name in ("Phase", "RotX", "RotY", "RotZ"):
            k, m = _rot(name, a[1])
            out.append(eng.op_record(k, q(0), m, ctrls))
        elif name == "Swap":
            out.append(eng.op_record(eng.OP_SWAP, q(0), controls=ctrls, target2=q(1)))
        elif name == "CX":
            out.append(eng.op_record(eng.OP_X, q(1), controls=tuple(ctrls) + (q(0),)))
        elif name == "CY":
            out.append(eng.op_record(eng.OP_U2, q(1), [0, -1j, 1j, 0], tuple(ctrls) + (q(0),)))
        elif name == "CZ":
            out.append(eng.op_record(eng.OP_DIAG, q(1), [1, 0, 0, -1], tuple(ctrls) + (q(0),)))
        elif name == "CH":
            out.append(eng.op_record(eng.OP_U2, q(1), [H, H, H, -H], tuple(ctrls) + (q(0),)))
        elif name in ("CPhase", "CRotX", "CRotY", "CRotZ"):
            k, m = _rot(name, a[2])
            out.append(eng.op_record(k, q(1), m, tuple(ctrls) + (q(0),)))
        elif name == "CSwap":
            out.append(eng.op_record(eng.OP_SWAP, q(1), controls=tuple(ctrls) + (q(0),), target2=q(2)))
        elif name == "CCNot":
            out.append(eng.op_record(eng.OP_X, q(2), controls=tuple(ctrls) + (q(0), q(1))))
        elif name == "Or":
            # t ^= (a | b)  ==  X(t) ; X(t) where a == 0 and b == 0   (both under the outer controls)
            out.append(eng.op_record(eng.OP_X, q(2), controls=ctrls))
            cm = eng.qmask(tuple(ctrls) + (q(0), q(1)))
            out.append(eng.op_record(eng.OP_X, q(2), controls=tuple(ctrls) + (q(0), q(1)),
                                     ctrl_value=eng.qmask(ctrls)))
            assert cm
        else:
            raise KeyError(name)
    return out


def lower_array(circ: Circ) -> np.ndarray:
    recs = lower(circ)
    return np.concatenate(recs) if recs else eng.make_ops(0)

"""CPU: the fusion planner builds plans without a GPU (descriptors are uploaded on first run);
these checks pin its structural guarantees on the BASELINE circuits."""
import numpy as np

from afquantumsim_b200 import engine as eng
from afquantumsim_b200 import workloads as wl
from oracle import oracle as orc
from tests.lowering import lower_array


def info(n, ops, flags=eng.PLAN_FUSE):
    return eng.Plan(n, ops, flags).info()


def test_unfused_plan_counts_every_gate():
    ops = wl.to_ops(wl.brickwork(30, 20))
    i = info(30, ops, 0)
    assert i["n_ops"] == 890 and i["n_launches"] == 890 and i["n_fused_passes"] == 0
    S = 8.0 * 2 ** 30
    assert i["bytes_unfused"] == 600 * 2 * S + 290 * S == i["bytes_planned"]      # 12.80 TB


def test_brickwork30_fuses_into_few_passes():
    i = info(30, wl.to_ops(wl.brickwork(30, 20)))
    assert i["n_single_ops"] == 0 and i["tile_bits"] == 13
    assert i["n_fused_passes"] <= 40
    assert i["bytes_planned"] == i["n_fused_passes"] * 2 * 8.0 * 2 ** 30
    assert i["bytes_planned"] < i["bytes_unfused"] / 15


def test_qft_needs_one_pass_per_seven_new_qubits():
    """diagonal gates and controls never constrain a tile, so only the n H gates do"""
    for n in (16, 22, 28):
        i = info(n, wl.to_ops(wl.qft(n)))
        assert i["n_fused_passes"] <= -(-(n - 5) // 7) + 1, (n, i["n_fused_passes"])


def test_million_op_circuits_plan_in_linear_time():
    import time
    circ = orc.grover_search(20, orc.grover_oracle(20, 5), 300)
    ops = lower_array(circ)
    t0 = time.perf_counter()
    i = info(20, ops)
    assert time.perf_counter() - t0 < 5.0
    assert i["n_ops"] == len(ops) and i["n_fused_passes"] <= 4 * 300 + 4


def test_small_states_stay_on_the_per_gate_kernels():
    i = info(6, wl.to_ops(wl.ghz(6)))
    assert i["n_fused_passes"] == 0 and i["n_launches"] == 6


def test_plan_rejects_bad_ops():
    import pytest
    bad = eng.op_record(eng.OP_X, 5)
    with pytest.raises(eng.EngineError):
        eng.Plan(4, bad)
    with pytest.raises(eng.EngineError):
        eng.Plan(0, eng.make_ops(0))

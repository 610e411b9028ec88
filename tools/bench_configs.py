"""BASELINE.json configs 1, 2 and 4 on one B200 through the public aqs API (config 3 is bench.py,
config 5 is `bench.py --gpus N`).  Writes gpurun_out/r01_configs.json.  Parity checks use the closed
forms of SURVEY.md Appendix D; timing is CUDA events on the engine stream for the resident legs and
host wall clock for the benchmark.cpp-style leg."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from afquantumsim_b200 import aqs  # noqa: E402
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402

aqs.initialize(0)
aqs.set_seed(1)
res = {}
timer = eng.Timer()


def device_ms(state, plan, reps=5, warm=2):
    for _ in range(warm):
        state.run(plan)
    timer.start(state)
    for _ in range(reps):
        state.run(plan)
    timer.stop(state)
    return timer.elapsed_ms() / reps


# ---- config 1: 16-qubit GHZ + profile_measure_all(1000), benchmark.cpp style (construct + simulate + profile) x100
def ghz_once():
    n = 16
    qc = aqs.QCircuit(n)
    qc << aqs.H(0)
    for i in range(n - 1):
        qc << aqs.CX(i, i + 1)
    qs = aqs.QSimulator(n)
    qs.simulate(qc)
    return qs.profile_measure_all(1000)


for fusion in (True, False):
    aqs.set_fusion(fusion)
    for _ in range(3):
        hist = ghz_once()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        hist = ghz_once()
        ts.append((time.perf_counter() - t0) * 1e3)
    assert hist[0] + hist[-1] == 1000 and hist.sum() == 1000
    res[f"config1_ghz16_fusion_{'on' if fusion else 'off'}"] = {
        "ms_mean": float(np.mean(ts)), "ms_sd": float(np.std(ts, ddof=1)), "runs": 100, "gates": 16,
        "what": "QCircuit build + QSimulator(16) + simulate + profile_measure_all(1000), host wall clock"}
aqs.set_fusion(True)

# ---- config 2: 28-qubit QFT
n = 28
S = 8.0 * (1 << n)
qc = aqs.fourier_transform(n)
ops = qc.ops()
st = eng.State(n)
for label, flags in (("fused", eng.PLAN_FUSE), ("unfused", 0)):
    plan = eng.Plan(n, ops, flags)
    info = plan.info()
    ms = device_ms(st, plan)
    res[f"config2_qft28_{label}"] = {"ms": ms, "gate_apps": len(ops), "gate_apps_per_s": len(ops) / (ms * 1e-3),
                                     "launches": int(info["n_launches"]),
                                     "algorithmic_GB": info["bytes_planned"] / 1e9,
                                     "GBps_algorithmic": info["bytes_planned"] / ms / 1e6}
# parity: QFT|x> = N^-1/2 exp(2 pi i rev(x) y / N) at sampled y, both paths
x = int(np.random.Generator(np.random.PCG64(2028)).integers(0, 1 << n))
rev = int(format(x, f"0{n}b")[::-1], 2)
ys = np.random.default_rng(2).integers(0, 1 << n, 2048)
for label, flags in (("fused", eng.PLAN_FUSE), ("unfused", 0)):
    st.set_basis(x)
    st.run(eng.Plan(n, ops, flags))
    got = np.array([st.amp(int(y)) for y in ys[:256]])
    want = np.exp(2j * np.pi * ((rev * ys[:256].astype(object)) % (1 << n)).astype(np.float64) / (1 << n)) / np.sqrt(float(1 << n))
    err = float(np.max(np.abs(got - want)) * np.sqrt(float(1 << n)))
    res[f"config2_qft28_{label}"]["max_rel_amp_error_vs_closed_form"] = err
    res[f"config2_qft28_{label}"]["norm2"] = st.norm2()
    assert err < 1e-3, err
del st

# ---- config 4: 26-qubit Grover, marked state 5 (examples/grover_search.cpp)
n, marked = 26, 5
full_iters = int(np.float32(np.pi) * np.sqrt(np.float32(1 << n)) / np.float32(4))     # 6433
w = int(format(marked, f"0{n}b")[::-1], 2)
theta = np.arcsin(2.0 ** (-n / 2))
oracle = aqs.grover_oracle(n, marked)
st = eng.State(n)
for label, flags, iters in (("fused_64it", eng.PLAN_FUSE, 64), ("unfused_64it", 0, 64), ("fused_full", eng.PLAN_FUSE, full_iters)):
    t0 = time.perf_counter()
    qc = aqs.QCircuit(n)
    qc << aqs.Gate(aqs.grover_search(n, oracle, iters, "Oracle"), 0)
    ops = qc.ops()
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = eng.Plan(n, ops, flags)
    t_plan = time.perf_counter() - t0
    st.set_basis(0)
    timer.start(st)
    st.run(plan)
    timer.stop(st)
    ms = timer.elapsed_ms()
    p_w = float(abs(st.amp(w)) ** 2)
    want = float(np.sin((2 * iters + 1) * theta) ** 2)
    info = plan.info()
    res[f"config4_grover26_{label}"] = {
        "iterations": iters, "gate_apps": len(ops), "ms": ms, "ms_per_iteration": ms / iters,
        "gate_apps_per_s": len(ops) / (ms * 1e-3), "launches": int(info["n_launches"]),
        "host_build_s": t_build, "host_plan_s": t_plan, "p_marked": p_w, "p_marked_closed_form": want,
        "norm2": st.norm2()}
    assert abs(p_w - want) < 5e-3 * max(1.0, iters / 64), (p_w, want)
# measurement on the final (amplified) state
u = np.random.default_rng(4).random(10000, dtype=np.float32)
t0 = time.perf_counter()
idx = st.sample(u)
res["config4_grover26_fused_full"]["sample_10000_ms"] = (time.perf_counter() - t0) * 1e3
res["config4_grover26_fused_full"]["fraction_of_draws_on_marked_state"] = float(np.mean(idx == w))
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/r01_configs.json", "w"), indent=1)

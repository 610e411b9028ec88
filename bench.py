#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on its own config.

  metric   : 30-qubit random brickwork circuit (600 random RotX/RotY/RotZ + 290 CX,
             depth 20, numpy PCG64 seed 30), gate applications per second
  step     : one simulate() of the whole 890-gate circuit on the 2^30 complex64 state
  value    : device-timed, compiled plan (QCircuit::compile's counterpart: fused passes with their
             specialised kernels) and state already resident in HBM
  e2e      : the same circuit through the public aqs API with HOST inputs each step:
             build QCircuit, QSimulator(n), simulate (lower + plan + H2D of the descriptors; the
             specialised kernels of the pass shapes come from the process-wide cache after the first
             step), profile 1000 host-generated draws (H2D) and read the outcomes back (D2H)
  roofline : algorithmic HBM bytes / CUDA-event time, for the dominant
             kernel of the headline (fused) leg and, beside it, for the per-gate kernels
             of an unfused leg of the same circuit
  parity   : the gate prefix the CPU oracle ran for cpu_baseline, re-run on a fresh GPU state through
             the headline path, relative L2 over all 2^30 amplitudes
  configs  : BASELINE.json configs 1, 2 and 4 (GHZ-16 benchmark.cpp style, QFT-28, Grover-26 slice)
  cpu_baseline / --impl reference : the CPU oracle (restated reference, OpenMP on all
             host cores) on a bounded prefix of the same circuit.  ArrayFire is not
             installable, so the reference's own binary cannot be timed.

One JSON line on stdout (rank 0).  Launch: python bench.py --gpus N --steps K --warmup W
(torchrun for N > 1: one rank per GPU; ONE state of 31 + log2(N) qubits, 16 GiB per GPU).
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0,
                    help="qubits PER GPU (default 30 on one GPU = BASELINE config 3, 31 per GPU when sharded = config 5: 32/33/34 qubits on 2/4/8 GPUs)")
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--draws", type=int, default=1000)
    ap.add_argument("--workload", default="brickwork", choices=["brickwork", "qft"],
                    help="brickwork = the headline circuit; qft = fourier_transform(n) (BASELINE configs 2 and 5)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 1, 2, 4")
    ap.add_argument("--no-jit", action="store_true", help="generic (interpreting) tile kernel only")
    a = ap.parse_args()
    if a.qubits == 0:
        a.qubits = 30 if a.gpus == 1 else 31
    return a


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the loaded interval (B200_PROFILING.md).  Samples carry
    nvidia-smi's own timestamp; only those inside [mark_start, mark_end] count — the process takes about a
    second to produce its first row, so it is started early and the loaded interval is made long enough."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device, period_ms=50):
        self.device, self.period, self.rows, self.proc = device, period_ms, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period), "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t = time.time()
            while not self.rows and time.time() - t < 4.0:
                time.sleep(0.05)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = time.time()
            self.rows.append((ts, f))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(2.5 * self.period / 1e3)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        t0, t1 = self.t0 or 0.0, self.t1 or 1e18
        inside = [f for ts, f in self.rows if t0 <= ts <= t1]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None

        sm = [v for v in (num(f[2]) for f in inside) if v is not None]
        mx = [v for v in (num(r[1][3]) for r in self.rows) if v is not None]
        pw = [v for v in (num(f[4]) for f in inside) if v is not None]
        counts = {name: sum(1 for f in inside if f[6 + i].lower().startswith("active")) for i, name in enumerate(self.NAMES)}
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(k for k, v in counts.items() if v), "reason_samples": counts,
                "samples": len(sm), "samples_total": len(self.rows), "period_ms": self.period,
                "window_s": (self.t1 - self.t0) if (self.t0 and self.t1) else None,
                "what": "nvidia-smi samples whose own timestamp falls inside the loaded interval (warm-up + timed steps + sustained loop of the headline leg)"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch at n = 30, from the committed ncu --set full captures
NCU_TRAFFIC_TILE = 17.125e9  # profiles/r02_spec_pass_ncu.txt (specialised pass kernel; algorithmic 2*S = 17.18e9: no re-reads)
NCU_TRAFFIC_PAIR = 17.12e9   # profiles/r01_pair_kernel_ncu.txt


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
def cpu_sample(n, gates, budget_s, steps=1, warmup=0, min_gates=2, keep_state=False):
    """Time the oracle (all host cores) on a prefix of `gates`, every step from |0...0>.
    Returns (gate-apps/s, info); info["state"] is the final oracle state when keep_state."""
    from oracle import oracle as orc
    try:
        orc.lib().orc_set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1
    except Exception:
        pass
    cores = orc.num_threads()
    avail_gib = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
    n_cpu = n
    while 8.0 * (1 << n_cpu) / 2**30 > 0.4 * avail_gib and n_cpu > 20:
        n_cpu -= 1
    gl = gates if n_cpu == n else __import__("afquantumsim_b200.workloads", fromlist=["x"]).brickwork(n_cpu, 20)
    a = orc.new_state(n_cpu)

    def reset():
        a.fill(0)
        a[0] = 1

    orc.simulate(a, orc.Circ(n_cpu, gl[:2]))           # first touch
    t0 = time.perf_counter()
    orc.simulate(a, orc.Circ(n_cpu, gl[2:6]))
    per_gate = max((time.perf_counter() - t0) / 4, 1e-6)
    per_step = budget_s / max(1, steps + warmup)
    count = int(max(min_gates, min(len(gl), per_step / per_gate)))
    circ = orc.Circ(n_cpu, gl[:count])
    for _ in range(warmup):
        reset()
        orc.simulate(a, circ)
    dts = []
    for _ in range(steps):
        reset()
        t0 = time.perf_counter()
        orc.simulate(a, circ)
        dts.append(time.perf_counter() - t0)
    dt = sum(dts) / len(dts)
    # scale to the n-qubit workload if the host could not hold it (stated in `sample`)
    scale = float(1 << (n - n_cpu))
    value = count / dt / scale
    sample = (f"first {count} of {len(gl)} gates of the {n_cpu}-qubit brickwork circuit per step from |0...0>, {steps} step(s), "
              f"OpenMP x{cores}" + ("" if n_cpu == n else f"; run at {n_cpu} qubits and divided by {int(scale)} (host RAM)"))
    info = {"cores": cores, "sample": sample, "ms_per_step": dt * 1e3, "gates_per_step": count, "qubits": n_cpu}
    if keep_state and n_cpu == n:
        info["state"] = a
    return value, info


def run_reference(args):
    from afquantumsim_b200 import workloads as wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = int(np.log2(max(1, args.gpus)))
    n = args.qubits + g                      # the N-GPU arm simulates ONE state of (qubits per GPU) + log2(N) qubits
    gates = wl.brickwork(n, args.depth)
    value, info = cpu_sample(n, gates, budget_s=150.0, steps=args.steps, warmup=args.warmup, min_gates=30)
    value *= float(2.0 ** (n - 30))          # 30-qubit equivalents, like the N-GPU arm
    unit = "gate-apps/s" if n == 30 else "gate-apps/s (30-qubit equivalents: gate applications x 2^(n-30))"
    line = {
        "impl": "reference", "metric": "30q random-circuit gate-apps/s", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": {"workload": f"brickwork-{n} depth {args.depth} (numpy PCG64 seed {n})",
                   "gates": len(gates), "note": "restated reference (CPU oracle): ArrayFire is not installable; a bounded gate prefix per step, "
                                                "extrapolated per gate (every gate is one streaming pass over the state)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def parity_vs_oracle(eng, n, ops_prefix, oracle_state, flags):
    """||gpu - oracle|| / ||oracle|| over ALL 2^n amplitudes, the GPU state downloaded in 512 MiB chunks."""
    st = eng.State(n)
    plan = eng.Plan(n, ops_prefix, flags)
    st.run(plan)
    chunk = 1 << 26
    buf = np.empty(min(chunk, 1 << n), dtype=np.complex64)
    num = den = 0.0
    for off in range(0, 1 << n, buf.size):
        st.download(off, buf.size, out=buf)
        ref = oracle_state[off:off + buf.size]
        d = buf - ref
        num += float(np.vdot(d, d).real)
        den += float(np.vdot(ref, ref).real)
    jit_ready = plan.jit_ready()
    passes = int(plan.info()["n_fused_passes"])
    st.close()
    return (num / den) ** 0.5, jit_ready, passes


def extra_configs(aqs, eng, wl, jit):
    """BASELINE.json configs 1, 2 and 4 through the public API / the engine, with their own parity checks."""
    res = {}
    timer = eng.Timer()
    # -- config 1: 16-qubit GHZ + profile_measure_all(1000), benchmark.cpp style (benchmark/benchmark.cpp:174-197):
    #    construct + simulate + profile, 100 runs, mean +- sample sd
    def ghz_once():
        n = 16
        qc = aqs.QCircuit(n)
        qc << aqs.H(0)
        for i in range(n - 1):
            qc << aqs.CX(i, i + 1)
        qs = aqs.QSimulator(n)
        qs.simulate(qc)
        return qs.profile_measure_all(1000)

    for _ in range(3):
        hist = ghz_once()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        hist = ghz_once()
        ts.append((time.perf_counter() - t0) * 1e3)
    assert int(hist[0]) + int(hist[-1]) == 1000 and int(hist.sum()) == 1000, "GHZ-16 histogram"
    res["ghz16"] = {"ms_mean": float(np.mean(ts)), "ms_sd": float(np.std(ts, ddof=1)), "runs": 100, "gates": 16,
                    "what": "QCircuit build + QSimulator(16) + simulate + profile_measure_all(1000), host wall clock, "
                            "benchmark/benchmark.cpp style; histogram exact (two bins)"}
    # -- config 2: 28-qubit QFT on a random basis state, closed-form check on sampled amplitudes
    n = 28
    ops = aqs.fourier_transform(n).ops()
    st = eng.State(n)
    plan = eng.Plan(n, ops, eng.PLAN_FUSE | jit)
    x = int(np.random.Generator(np.random.PCG64(2028)).integers(0, 1 << n))
    for _ in range(2):
        st.run(plan)
    timer.start(st)
    for _ in range(5):
        st.run(plan)
    timer.stop(st)
    ms = timer.elapsed_ms() / 5
    st.set_basis(x)
    st.run(plan)
    rev = int(format(x, f"0{n}b")[::-1], 2)
    ys = np.random.default_rng(2).integers(0, 1 << n, 256)
    got = np.array([st.amp(int(y)) for y in ys])
    want = np.exp(2j * np.pi * ((rev * ys.astype(object)) % (1 << n)).astype(np.float64) / (1 << n)) / np.sqrt(float(1 << n))
    err = float(np.max(np.abs(got - want)) * np.sqrt(float(1 << n)))
    assert err < 1e-4, f"QFT-28 closed-form error {err}"
    info = plan.info()
    res["qft28"] = {"ms": ms, "gate_apps": len(ops), "gate_apps_per_s": len(ops) / (ms * 1e-3), "passes": int(info["n_fused_passes"]),
                    "jit_passes": plan.jit_ready(), "algorithmic_GB": info["bytes_planned"] / 1e9,
                    "GBps_algorithmic": info["bytes_planned"] / ms / 1e6,
                    "max_rel_amp_error_vs_closed_form": err, "norm2": st.norm2()}
    st.close()
    # -- config 4: 26-qubit Grover (examples/grover_search.cpp:26-40), a 64-iteration slice of the 6433
    n, marked, iters = 26, 5, 64
    w = int(format(marked, f"0{n}b")[::-1], 2)
    theta = np.arcsin(2.0 ** (-n / 2))
    qc = aqs.QCircuit(n)
    qc << aqs.Gate(aqs.grover_search(n, aqs.grover_oracle(n, marked), iters, "Oracle"), 0)
    ops = qc.ops()
    st = eng.State(n)
    plan = eng.Plan(n, ops, eng.PLAN_FUSE | jit)
    st.run(plan)
    st.set_basis(0)
    timer.start(st)
    st.run(plan)
    timer.stop(st)
    ms = timer.elapsed_ms()
    p_w = float(abs(st.amp(w)) ** 2)
    want = float(np.sin((2 * iters + 1) * theta) ** 2)
    info = plan.info()
    res["grover26_64it"] = {"iterations": iters, "gate_apps": len(ops), "ms": ms, "ms_per_iteration": ms / iters,
                            "gate_apps_per_s": len(ops) / (ms * 1e-3), "passes": int(info["n_fused_passes"]), "jit_passes": plan.jit_ready(),
                            "p_marked": p_w, "p_marked_closed_form": want, "p_marked_rel_error": abs(p_w - want) / want,
                            "norm2": st.norm2(),
                            "note": "oracle parity of the same slice: tests/test_gpu_configs.py (amplitudes <= 1e-5 relative L2)"}
    assert abs(p_w - want) / want < 2e-3, (p_w, want)
    st.close()
    return res


def run_ours(args):
    import torch
    from afquantumsim_b200 import aqs
    from afquantumsim_b200 import engine as eng
    from afquantumsim_b200 import workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        torch.cuda.set_device(local)
        aqs.initialize(local)
        aqs.set_seed(30 + rank)
        return run_sharded(args, torch, dist, aqs, eng, wl, rank, world, local)
    torch.cuda.set_device(local)
    in_tree = os.path.join(ROOT, "afquantumsim_b200", "lib", "libaqs_engine.so")
    assert os.path.realpath(eng.LIB_PATH) == os.path.realpath(in_tree), "bench.py measures the in-tree engine only (AQS_ENGINE_LIB is for A/B builds)"
    sampler = ClockSampler(local).start()
    aqs.initialize(local)
    aqs.set_seed(30 + rank)
    jit = 0 if args.no_jit else eng.PLAN_JIT

    n, K, W = args.qubits, args.steps, max(args.warmup, 3)
    gates = wl.brickwork(n, args.depth)
    S = 8.0 * (1 << n)

    # ---- resident leg(s): state + plan in HBM, device-timed ----------------------------
    qc = aqs.QCircuit(n).extend(gates)
    ops = qc.ops()
    state = eng.State(n)
    timer = eng.Timer()

    def timed_plan(flags, sustain_s=0.0, marks=False):
        t0 = time.perf_counter()
        plan = eng.Plan(n, ops, flags)
        build_s = time.perf_counter() - t0
        info = plan.info()
        if marks:
            sampler.mark_start()
        for _ in range(W):
            state.run(plan)
        state.sync()
        c0 = eng.counters()
        timer.start(state)
        for _ in range(K):
            state.run(plan)
        timer.stop(state)
        ms = timer.elapsed_ms()
        c1 = eng.counters()
        sustained = None
        if sustain_s > 0:
            reps = max(K, int(sustain_s / max(ms / K * 1e-3, 1e-4)))
            timer.start(state)
            for _ in range(reps):
                state.run(plan)
            timer.stop(state)
            sustained = {"ms_per_step": timer.elapsed_ms() / reps, "steps": reps,
                         "what": "the same plan looped for ~%.0f s right after the timed steps (clock / power-cap behaviour)" % sustain_s}
        if marks:
            sampler.mark_end()
        return ms / K, info, c1["kernel_launches"] - c0["kernel_launches"], plan.jit_ready(), build_s, sustained

    ms_fused, info_f, launches_f, jit_ready, build_s, sustained = timed_plan(eng.PLAN_FUSE | jit, sustain_s=3.0, marks=True)
    clocks = sampler.stop()
    if jit:
        assert jit_ready == info_f["n_fused_passes"], f"specialised kernels missing: {jit_ready} of {info_f['n_fused_passes']} ({eng.jit_info()})"
    ms_generic, _, launches_g, _, _, _ = timed_plan(eng.PLAN_FUSE)
    ms_unfused, info_u, launches_u, _, _, _ = timed_plan(0)
    norm2 = state.norm2()
    assert abs(norm2 - 1.0) < 1e-3, f"state norm drifted: {norm2}"

    peak, peak_src = measured_peak_gbs()
    gate_apps = len(gates)
    value = gate_apps / (ms_fused * 1e-3)
    gbs_f = info_f["bytes_planned"] / (ms_fused * 1e-3) / 1e9
    gbs_u = info_u["bytes_planned"] / (ms_unfused * 1e-3) / 1e9

    # ---- e2e leg: public API, host inputs every step -------------------------------------
    state.close()
    del state
    rng = np.random.default_rng(rank)
    h2d = d2h = 0
    if args.no_jit:
        aqs.set_jit_min_qubits(99)

    def e2e_step():
        nonlocal h2d, d2h
        c0 = eng.counters()
        circ = aqs.QCircuit(n).extend(gates)           # host: 890 gate objects
        qs = aqs.QSimulator(n)                          # device state at |0...0>
        qs.simulate(circ)                               # lower + plan + H2D descriptors + kernels
        u = rng.random(args.draws, dtype=np.float32)    # host draws
        out = qs.sample(u)                              # H2D draws, D2H outcomes
        c1 = eng.counters()
        h2d = c1["h2d_bytes"] - c0["h2d_bytes"]
        d2h = c1["d2h_bytes"] - c0["d2h_bytes"]
        return out

    t1 = time.perf_counter()
    e2e_step()
    first_ms = (time.perf_counter() - t1) * 1e3
    eng.jit_wait()                                      # (the kernels of these shapes are in the cache already: the resident leg compiled them)
    for _ in range(max(1, W - 1)):
        e2e_step()
    torch.cuda.synchronize()
    e2e_steps = []
    t0 = time.perf_counter()
    for _ in range(K):
        t1 = time.perf_counter()
        e2e_step()                      # returns after the D2H read of the outcomes
        e2e_steps.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    e2e_value = gate_apps / (e2e_ms * 1e-3)

    line = {
        "metric": "30q random-circuit gate-apps/s", "value": value, "unit": "gate-apps/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_fused,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64", "data": "synthetic",
        "config": {
            "workload": f"brickwork-{n} depth {args.depth}: {gate_apps} gate applications "
                        f"(600 random RotX/RotY/RotZ + 290 CX at n=30), numpy PCG64 seed {n}, state {S / 2**30:.0f} GiB",
            "fusion": "on (headline); the unfused leg is reported under `unfused`, the generic tile kernel under `generic_kernel`",
            "jit": ("off" if args.no_jit else
                    "fused passes run specialised kernels: straight-line sm_100a code generated per pass shape, compiled with NVRTC when the "
                    "circuit is compiled (%.2f s for this plan, once per process), cached by shape; matrix entries are kernel parameters" % build_s),
            "parallelism": "single GPU",
            "l2": "state is 8 GiB >> 126 MB L2: every pass streams from HBM, no flush needed",
        },
        "e2e": {"value": e2e_value, "unit": "gate-apps/s", "ms_per_step": e2e_ms, "steps_ms": [round(x, 1) for x in e2e_steps],
                "first_step_ms": first_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": "QCircuit build + QSimulator(n) + simulate (lower, plan, descriptor upload; kernels of known pass shapes come from the "
                        "process-wide cache) + 1000-draw sample readback, host wall clock"},
        "gpu_launches": int(launches_f),
        "clocks": clocks,
        "sustained": sustained,
        "roofline": {"bound": "hbm", "achieved": gbs_f, "peak": peak, "unit": "GB/s", "frac": gbs_f / peak,
                     "traffic": NCU_TRAFFIC_TILE if info_f["n_fused_passes"] else NCU_TRAFFIC_PAIR,
                     "traffic_source": "ncu --set full capture committed under profiles/ (r02_spec_pass_ncu.txt), per launch; not re-measured in this run",
                     "peak_source": peak_src,
                     "note": "light passes of the specialised kernel run at the HBM floor (2.6 ms per pass), heavy ones are bound by the FP32 pipe "
                             "(ncu on the heaviest: fma pipe 67 %, DRAM 26 %; on a light one: DRAM 76 %; DESIGN.md 3.3); the HBM-bound per-gate kernels are under `unfused.roofline`",
                     "kernel": ("specialised pass kernels (aqs_pass)" if jit_ready else "generic tile kernel (k_tile2)") if info_f["n_fused_passes"] else "per-gate kernels",
                     "algorithmic_bytes_per_step": info_f["bytes_planned"], "launches_per_step": info_f["n_launches"]},
        "generic_kernel": {"value": gate_apps / (ms_generic * 1e-3), "unit": "gate-apps/s", "ms_per_step": ms_generic, "gpu_launches": int(launches_g),
                           "what": "the same fused plan on the interpreting tile kernel (what runs while specialised kernels compile in the background)"},
        "unfused": {"value": gate_apps / (ms_unfused * 1e-3), "unit": "gate-apps/s", "ms_per_step": ms_unfused,
                    "gpu_launches": int(launches_u),
                    "roofline": {"bound": "hbm", "achieved": gbs_u, "peak": peak, "unit": "GB/s", "frac": gbs_u / peak,
                                 "traffic": NCU_TRAFFIC_PAIR, "kernel": "k_pair / k_diag per-gate kernels",
                                 "algorithmic_bytes_per_step": info_u["bytes_planned"],
                                 "frac_of_8TBs_nominal": gbs_u / 8000.0}},
        "plan": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in info_f.items()},
        "jit": eng.jit_info(),
    }
    if not args.no_cpu_baseline:
        v, info = cpu_sample(n, gates, budget_s=args.cpu_seconds, keep_state=True)
        line["cpu_baseline"] = {"value": v, "unit": "gate-apps/s", "cores": info["cores"], "kind": "port",
                                "sample": info["sample"]}
        if "state" in info:
            # parity on the headline configuration: the same gate prefix through the headline path
            cnt = info["gates_per_step"]
            pre_ops = aqs.QCircuit(n).extend(gates[:cnt]).ops()
            rel, jr, ps = parity_vs_oracle(eng, n, pre_ops, info["state"], eng.PLAN_FUSE | jit)
            line["parity"] = {"n": n, "gates": cnt, "rel_l2": rel, "amplitudes_compared": 1 << n, "tolerance": 1e-5,
                              "path": f"fused plan, {jr} of {ps} passes on specialised kernels", "ok": bool(rel < 1e-5)}
            assert rel < 1e-5, f"n = {n} parity against the oracle: rel L2 {rel}"
            del info["state"]
        # the reference's own ALGORITHM (an explicit 2^n x 2^n operator per gate) at its own published sizes, same host cores
        from oracle import ref_algorithm
        line["cpu_baseline_ref_algorithm"] = ref_algorithm.published_sizes()
    if not args.no_configs:
        eng.pool_trim()
        line["configs"] = extra_configs(aqs, eng, wl, jit)
    print(json.dumps(line), flush=True)


def sharded_parity(ShardedState, orc, wl, torch, dist, rank, world, g, jit):
    """Before timing: the data path of this run (flat address space where available) against the unsharded CPU oracle."""
    n = 22 + g
    gates = wl.brickwork(n, 8)
    st = ShardedState(n, jit=jit)
    plan = st.compile(wl.to_ops(gates))          # (waits for the specialised kernels: the path that is timed below)
    st.run(plan)
    jit_passes, passes = plan.jit_ready(), plan.n_passes
    del plan
    u = np.random.default_rng(5).random(500, dtype=np.float32)
    out = st.sample(u)
    full = st.gather()
    res = [None]
    if rank == 0:
        want = orc.simulate(orc.new_state(n), orc.Circ(n, gates))
        res[0] = {"n": n, "gates": len(gates), "rel_l2": orc.rel_l2(full, want),
                  "sampling_bit_identical_to_oracle": bool(np.array_equal(out, orc.sample(full, u, "exact"))),   # same state, same draws
                  "flat_address_space": st.flat_state is not None, "world": world,
                  "path": f"{jit_passes} of {passes} passes on specialised kernels"}
    dist.broadcast_object_list(res, src=0)
    st.close()
    del st
    assert res[0]["rel_l2"] < 1e-5 and res[0]["sampling_bit_identical_to_oracle"], res[0]
    return res[0]


def run_sharded(args, torch, dist, aqs, eng, wl, rank, world, local):
    """N > 1: ONE state of (qubits per GPU) + log2(N) qubits sharded over the N GPUs (weak scaling, 16 GiB per GPU by
    default: BASELINE config 5); tiles that contain rank bits load / store peer HBM over NVLink inside the pass kernel."""
    from afquantumsim_b200.sharded import ShardedState
    from oracle import oracle as orc
    g = int(np.log2(world))
    n, K, W = args.qubits + g, args.steps, max(args.warmup, 1)
    make_gates = (lambda: wl.qft(n)) if args.workload == "qft" else (lambda: wl.brickwork(n, args.depth))
    gates = make_gates()
    ops = wl.to_ops(gates)
    S_shard = 8.0 * (1 << (n - g))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    parity = sharded_parity(ShardedState, orc, wl, torch, dist, rank, world, g, not args.no_jit)
    eng.pool_trim()
    st = ShardedState(n, jit=not args.no_jit)

    # like the single-GPU leg, the resident leg times a COMPILED circuit (QCircuit::compile's counterpart:
    # one fused plan over the whole state, its specialised kernels compiled); every step starts from |0...0>
    # (the shard memset is inside the timed region); planning time is part of `e2e` below
    t0 = time.perf_counter()
    plan = st.compile(ops)
    build_s = time.perf_counter() - t0

    def step():
        st.set_basis(0)
        st.run(plan)

    sampler = ClockSampler(local).start()
    sampler.mark_start()
    for _ in range(W):
        step()
    barrier()
    c0 = eng.counters()
    s0 = dict(st.stats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier()
    sampler.mark_end()
    c1 = eng.counters()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    exchanges = (st.stats["exchanges"] - s0["exchanges"]) / K
    xbytes = (st.stats["exchange_bytes"] - s0["exchange_bytes"]) / K
    norm2 = st.norm2()
    assert abs(norm2 - 1.0) < 1e-3, f"state norm drifted: {norm2}"
    remap_p2p, plan_passes, plan_local_ops = bool(st.p2p), plan.n_passes, plan.n_local_ops
    flat_mode = st.flat_state is not None
    schedule = st.stats.get("schedule") or ("remap" if not flat_mode else "flat")
    jit_passes = plan.jit_ready()
    spans = plan.spans() if schedule == "flat" else []
    del plan

    # e2e: host-built gate list -> ops -> sharded simulate -> 1000-draw sample read back, every step
    rng = np.random.default_rng(1234)
    u = rng.random(args.draws, dtype=np.float32)

    def e2e_step():
        s2 = ShardedState(n, jit=not args.no_jit)
        s2.apply_ops(wl.to_ops(make_gates()))
        out = s2.sample(u)
        s2.close()
        return out

    st.close()
    del st
    e2e_step()
    eng.jit_wait()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    barrier()
    if rank != 0:
        dist.destroy_process_group()
        return
    # a gate application on the 2^n state is 2^(n-30) times the amplitude work of a 30-qubit one
    scale = float(2.0 ** (n - 30))
    gate_apps = len(gates)
    peak, peak_src = measured_peak_gbs()
    unit = "gate-apps/s (30-qubit equivalents: gate applications x 2^(n-30))"
    line = {
        "metric": "30q random-circuit gate-apps/s", "value": scale * gate_apps / (ms * 1e-3), "unit": unit,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": {
            "workload": (f"brickwork-{n} depth {args.depth}" if args.workload == "brickwork" else f"fourier_transform({n})")
                        + f": {gate_apps} gate applications on ONE 2^{n} state; {S_shard / 2**30:.0f} GiB shard per GPU",
            "parallelism": f"state sharded over {world} GPUs on the top {g} qubits; "
                           + ("all shards mapped into one flat NVLink address space (CUDA VMM); " if flat_mode else "")
                           + ("schedule 'flat': one fused plan over the whole state, each GPU runs 1/N of the tiles of every pass, tiles that "
                              "contain rank bits move their remote part over NVLink (staged through the copy engines, written back from the kernel)"
                              if schedule == "flat" else
                              "schedule 'remap': lazy global-qubit swaps around fused local plans, " +
                              ("each swap one in-place exchange kernel over peer memory (aqs_peer_bitswap)" if remap_p2p else "each swap a half-shard NCCL send/recv"))
                           + "; both schedules are planned and the one that moves fewer bytes runs",
            "fusion": "on", "jit": f"{jit_passes} of {plan_passes} passes on specialised kernels (compiled in {build_s:.2f} s)",
            "l2": f"shard is {S_shard / 2**30:.0f} GiB >> 126 MB L2",
        },
        "raw_gate_apps_per_s": gate_apps / (ms * 1e-3),
        "parity_sharded": parity,
        "e2e": {"value": scale * gate_apps / (e2e_ms * 1e-3), "unit": unit,
                "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(len(ops) * 64 + args.draws * 8),
                "d2h_bytes_per_step": int(args.draws * 8),
                "what": "gate list -> ops -> ShardedState simulate (plan + cached kernels) -> 1000-draw distributed sample, host wall clock; "
                        "byte counts by formula (op records + draws in, outcomes out)"},
        "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
        "clocks": clocks,
        "exchange": {"per_step": exchanges, "bytes_per_rank_per_step": xbytes, "peer_memory": remap_p2p or flat_mode, "flat_address_space": flat_mode,
                     "schedule": schedule,
                     "local_passes_per_step": plan_passes, "local_ops_per_step": plan_local_ops, "rank_bits_per_pass": spans,
                     "note": "bytes each rank writes to its peers over NVLink per step (it reads as many)"},
        "roofline": {"bound": "hbm", "achieved": plan_passes * 2.0 * S_shard / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": plan_passes * 2.0 * S_shard / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "specialised pass kernels" if jit_passes else "generic tile kernel", "launches_per_step": plan_passes,
                     "note": "per GPU: algorithmic bytes = passes x 2 x shard bytes (every GPU runs 1/N of the tiles of every pass); "
                             "passes whose tiles span GPUs are bound by NVLink (DESIGN.md 4)"},
    }
    print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

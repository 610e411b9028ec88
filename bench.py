#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on its own config.

  metric   : 30-qubit random brickwork circuit (600 random RotX/RotY/RotZ + 290 CX,
             depth 20, numpy PCG64 seed 30), gate applications per second
  step     : one simulate() of the whole 890-gate circuit on the 2^30 complex64 state
  value    : device-timed, plan and state already resident in HBM (fusion on)
  e2e      : the same circuit through the public aqs API with HOST inputs each step:
             build QCircuit, QSimulator(n), simulate (plan build + H2D of the op
             descriptors), profile 1000 host-generated draws (H2D) and read the
             outcomes back (D2H)
  roofline : algorithmic HBM bytes / CUDA-event time, for the dominant
             kernel of the headline (fused) leg and, beside it, for the per-gate kernels
             of an unfused leg of the same circuit
  cpu_baseline / --impl reference : the CPU oracle (restated reference, OpenMP on all
             host cores) on a bounded prefix of the same circuit.  ArrayFire is not
             installable, so the reference's own binary cannot be timed.

One JSON line on stdout (rank 0).  Launch: python bench.py --gpus N --steps K --warmup W
(torchrun for N > 1: one rank per GPU).
"""
import argparse
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
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--draws", type=int, default=1000)
    ap.add_argument("--workload", default="brickwork", choices=["brickwork", "qft"],
                    help="brickwork = the headline circuit; qft = fourier_transform(n) (BASELINE configs 2 and 5)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch at n = 30, from the committed ncu --set full captures
NCU_TRAFFIC_TILE = 17.12e9   # profiles/r01_tile_kernel_ncu.txt (algorithmic 2*S = 17.18e9: no re-reads)
NCU_TRAFFIC_PAIR = 17.12e9   # profiles/r01_pair_kernel_ncu.txt


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n, gates):
    """bytes per gate application if every gate ran alone."""
    S = 8.0 * (1 << n)
    tot = 0.0
    for g in gates:
        tot += {"RotX": 2 * S, "RotY": 2 * S, "RotZ": 2 * S, "H": 2 * S, "X": 2 * S, "CX": S,
                "CPhase": S / 2, "Phase": S, "Z": S}.get(g[0], 2 * S)
    return tot


# ---------------------------------------------------------------------------
def cpu_sample(n, gates, budget_s, steps=1, warmup=0):
    """Time the oracle (all host cores) on a prefix of `gates`; returns (gate-apps/s, info)."""
    from oracle import oracle as orc
    cores = orc.num_threads()
    avail_gib = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
    n_cpu = n
    while 8.0 * (1 << n_cpu) / 2**30 > 0.4 * avail_gib and n_cpu > 20:
        n_cpu -= 1
    gl = gates if n_cpu == n else __import__("afquantumsim_b200.workloads", fromlist=["x"]).brickwork(n_cpu, 20)
    a = orc.new_state(n_cpu)
    t0 = time.perf_counter()
    orc.simulate(a, orc.Circ(n_cpu, gl[:2]))           # first touch + calibration
    per_gate = max((time.perf_counter() - t0) / 2, 1e-6)
    t0 = time.perf_counter()
    orc.simulate(a, orc.Circ(n_cpu, gl[2:4]))
    per_gate = max((time.perf_counter() - t0) / 2, 1e-6)
    per_step = budget_s / max(1, steps + warmup)
    count = int(max(2, min(len(gl), per_step / per_gate)))
    circ = orc.Circ(n_cpu, gl[:count])
    for _ in range(warmup):
        orc.simulate(a, circ)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.simulate(a, circ)
    dt = (time.perf_counter() - t0) / steps
    # scale to the n-qubit workload if the host could not hold it (stated in `sample`)
    scale = float(1 << (n - n_cpu))
    value = count / dt / scale
    sample = (f"first {count} of {len(gl)} gates of the {n_cpu}-qubit brickwork circuit per step, {steps} step(s), "
              f"OpenMP x{cores}" + ("" if n_cpu == n else f"; run at {n_cpu} qubits and divided by {int(scale)} (host RAM)"))
    return value, {"cores": cores, "sample": sample, "ms_per_step": dt * 1e3, "gates_per_step": count, "qubits": n_cpu}


def run_reference(args):
    from afquantumsim_b200 import workloads as wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = int(np.log2(max(1, args.gpus)))
    n = args.qubits + g                      # the N-GPU arm simulates ONE state of 30 + log2(N) qubits
    gates = wl.brickwork(n, args.depth)
    value, info = cpu_sample(n, gates, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    value *= float(1 << g)                   # 30-qubit equivalents, like the N-GPU arm
    line = {
        "impl": "reference", "metric": "30q random-circuit gate-apps/s", "value": value,
        "unit": "gate-apps/s" if g == 0 else "gate-apps/s (30-qubit equivalents: gate applications x 2^(n-30))",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": {"workload": f"brickwork-{n} depth {args.depth} (numpy PCG64 seed {n})",
                   "gates": len(gates), "note": "restated reference (CPU oracle): ArrayFire is not installable"},
        "cpu_baseline": {"value": value, "unit": "gate-apps/s", "cores": info["cores"], "kind": "port",
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": "gate-apps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    from afquantumsim_b200 import aqs
    from afquantumsim_b200 import engine as eng
    from afquantumsim_b200 import workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    aqs.initialize(local)
    aqs.set_seed(30 + rank)

    n, K, W = args.qubits, args.steps, max(args.warmup, 0)
    if world > 1:
        return run_sharded(args, torch, dist, aqs, eng, wl, rank, world, local)
    gates = wl.brickwork(n, args.depth)
    S = 8.0 * (1 << n)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not dist:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident leg(s): state + plan in HBM, device-timed ----------------------------
    qc = aqs.QCircuit(n).extend(gates)
    ops = qc.ops()
    state = eng.State(n)
    timer = eng.Timer()

    def timed_plan(flags):
        plan = eng.Plan(n, ops, flags)
        info = plan.info()
        for _ in range(max(W, 3)):
            state.run(plan)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        c0 = eng.counters()
        timer.start(state)
        for _ in range(K):
            state.run(plan)
        timer.stop(state)
        ms = timer.elapsed_ms()
        barrier()
        c1 = eng.counters()
        clocks = sampler.stop()
        ms = max_over_ranks(ms)
        return ms / K, info, c1["kernel_launches"] - c0["kernel_launches"], clocks

    ms_fused, info_f, launches_f, clocks = timed_plan(eng.PLAN_FUSE)
    ms_unfused, info_u, launches_u, clocks_u = timed_plan(0)
    norm2 = state.norm2()
    assert abs(norm2 - 1.0) < 1e-3, f"state norm drifted: {norm2}"

    peak, peak_src = measured_peak_gbs()
    gate_apps = len(gates)
    value = world * gate_apps / (ms_fused * 1e-3)
    gbs_f = info_f["bytes_planned"] / (ms_fused * 1e-3) / 1e9
    gbs_u = info_u["bytes_planned"] / (ms_unfused * 1e-3) / 1e9

    # ---- e2e leg: public API, host inputs every step -------------------------------------
    del state
    rng = np.random.default_rng(rank)
    h2d = d2h = 0

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

    for _ in range(max(1, W)):
        e2e_step()
    barrier()
    e2e_steps = []
    t0 = time.perf_counter()
    for _ in range(K):
        t1 = time.perf_counter()
        e2e_step()                      # returns after the D2H read of the outcomes
        e2e_steps.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    barrier()
    e2e_value = world * gate_apps / (e2e_ms * 1e-3)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    line = {
        "metric": "30q random-circuit gate-apps/s", "value": value, "unit": "gate-apps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_fused,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64", "data": "synthetic",
        "config": {
            "workload": f"brickwork-{n} depth {args.depth}: {gate_apps} gate applications "
                        f"(600 random RotX/RotY/RotZ + 290 CX at n=30), numpy PCG64 seed {n}, state {S / 2**30:.0f} GiB",
            "fusion": "on (headline); the unfused leg is reported under `unfused`",
            "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one circuit per GPU)",
            "l2": "state is 8 GiB >> 126 MB L2: every pass streams from HBM, no flush needed",
        },
        "e2e": {"value": e2e_value, "unit": "gate-apps/s", "ms_per_step": e2e_ms, "steps_ms": [round(x, 1) for x in e2e_steps],
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "what": "QCircuit build + QSimulator(n) + simulate (plan build, descriptor upload) + 1000-draw sample readback, host wall clock"},
        "gpu_launches": int(launches_f),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": gbs_f, "peak": peak, "unit": "GB/s", "frac": gbs_f / peak,
                     "traffic": NCU_TRAFFIC_TILE if info_f["n_fused_passes"] else NCU_TRAFFIC_PAIR, "peak_source": peak_src,
                     "note": "the fused kernel is bound by FP32 work and interpreter dispatch latency, not HBM "
                             "(DESIGN.md 3.2: ncu fma pipe 40%, issue 56%); the HBM-bound per-gate kernels are under `unfused.roofline`",
                     "kernel": "fused tile kernel" if info_f["n_fused_passes"] else "per-gate kernels",
                     "algorithmic_bytes_per_step": info_f["bytes_planned"], "launches_per_step": info_f["n_launches"]},
        "unfused": {"value": world * gate_apps / (ms_unfused * 1e-3), "unit": "gate-apps/s", "ms_per_step": ms_unfused,
                    "gpu_launches": int(launches_u),
                    "roofline": {"bound": "hbm", "achieved": gbs_u, "peak": peak, "unit": "GB/s", "frac": gbs_u / peak,
                                 "traffic": NCU_TRAFFIC_PAIR, "kernel": "k_pair / k_diag per-gate kernels",
                                 "algorithmic_bytes_per_step": info_u["bytes_planned"],
                                 "frac_of_8TBs_nominal": gbs_u / 8000.0}},
        "plan": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in info_f.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        v, info = cpu_sample(n, gates, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": "gate-apps/s", "cores": info["cores"], "kind": "port",
                                "sample": info["sample"]}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def run_sharded(args, torch, dist, aqs, eng, wl, rank, world, local):
    """N > 1: ONE state of 30 + log2(N) qubits sharded over the N GPUs (8 GiB per GPU, weak scaling);
    gates on the global qubits exchange half-shards over NVLink with NCCL send/recv."""
    from afquantumsim_b200.sharded import ShardedState
    g = int(np.log2(world))
    n, K, W = args.qubits + g, args.steps, max(args.warmup, 0)
    make_gates = (lambda: wl.qft(n)) if args.workload == "qft" else (lambda: wl.brickwork(n, args.depth))
    gates = make_gates()
    ops = wl.to_ops(gates)
    S_shard = 8.0 * (1 << (n - g))
    st = ShardedState(n)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # like the single-GPU leg, the resident leg times a COMPILED circuit (QCircuit::compile's counterpart:
    # fused local plans + remap schedule, built once); every step starts from |0...0> in the canonical layout
    # (the 8 GiB memset is inside the timed region); planning time is part of `e2e` below
    plan = st.compile(ops)

    def step():
        st.set_basis(0)
        st.run(plan)

    for _ in range(max(W, 1)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
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
    del plan

    # e2e: host-built gate list -> ops -> sharded simulate -> 1000-draw sample read back, every step
    rng = np.random.default_rng(1234)
    u = rng.random(args.draws, dtype=np.float32)

    def e2e_step():
        s2 = ShardedState(n)
        s2.apply_ops(wl.to_ops(make_gates()))
        return s2.sample(u)

    del st
    e2e_step()
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
    # a gate application on the 2^n state is 2^g times the amplitude work of a 30-qubit one
    scale = float(2.0 ** (n - 30))
    gate_apps = len(gates)
    peak, peak_src = measured_peak_gbs()
    line = {
        "metric": "30q random-circuit gate-apps/s", "value": scale * gate_apps / (ms * 1e-3),
        "unit": "gate-apps/s (30-qubit equivalents: gate applications x 2^(n-30))",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": {
            "workload": (f"brickwork-{n} depth {args.depth}" if args.workload == "brickwork" else f"fourier_transform({n})")
                        + f": {gate_apps} gate applications on ONE 2^{n} state; {S_shard / 2**30:.0f} GiB shard per GPU",
            "parallelism": f"state sharded over {world} GPUs on the top {g} qubits; "
                           + ("all shards in one flat NVLink address space: one fused plan over the whole state, each GPU runs "
                              "1/N of the tiles of every pass, tiles that contain rank bits load/store peer memory" if flat_mode
                              else "global-qubit remaps " + ("in place over NVLink peer memory (aqs_peer_bitswap)" if remap_p2p
                                                            else "as half-shard NCCL send/recv")),
            "fusion": "on", "l2": "shard is 8 GiB >> 126 MB L2",
        },
        "raw_gate_apps_per_s": gate_apps / (ms * 1e-3),
        "e2e": {"value": scale * gate_apps / (e2e_ms * 1e-3), "unit": "gate-apps/s (30-qubit equivalents)",
                "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(len(ops) * 64 + args.draws * 8),
                "d2h_bytes_per_step": int(args.draws * 8),
                "what": "gate list -> ops -> ShardedState simulate -> 1000-draw distributed sample, host wall clock"},
        "gpu_launches": int(c1["kernel_launches"] - c0["kernel_launches"]),
        "clocks": clocks,
        "exchange": {"per_step": exchanges, "bytes_per_rank_per_step": xbytes, "peer_memory": remap_p2p or flat_mode, "flat_address_space": flat_mode,
                     "local_passes_per_step": plan_passes, "local_ops_per_step": plan_local_ops,
                     "note": "bytes each rank writes to its peers over NVLink per step (it reads as many)"},
        "roofline": {"bound": "hbm", "achieved": plan_passes * 2.0 * S_shard / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": plan_passes * 2.0 * S_shard / (ms * 1e-3) / 1e9 / peak, "traffic": NCU_TRAFFIC_TILE if S_shard == 8.0 * 2 ** 30 else None,
                     "peak_source": peak_src, "kernel": "fused tile kernel", "launches_per_step": plan_passes,
                     "note": "per GPU: algorithmic bytes = passes x 2 x shard bytes (every GPU runs 1/N of the tiles of every pass); "
                             "the kernel is bound by FP32 work and dispatch, and passes whose tiles span GPUs by NVLink (DESIGN.md 3.2, 4)"},
    }
    print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

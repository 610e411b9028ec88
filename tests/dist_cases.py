"""Sharded-state checks, run as R cooperating processes (torch.distributed).

  python tests/dist_cases.py --backend cpu --world 2     gloo + the oracle-backed ABI stand-in
  python tests/dist_cases.py --backend cuda --world 2    nccl + the CUDA engine, one rank per GPU

Every rank reconstructs the full state (all_gather) and compares it with the CPU
oracle run on the whole, unsharded circuit.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

TOL = 1e-5


def _setup(rank, world, backend, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from afquantumsim_b200 import engine as eng
    if backend == "cpu":
        eng.LIB_PATH = os.path.join(ROOT, "oracle", "_build", "cpu_abi", "libaqs_engine.so")   # test double
        dist.init_process_group("gloo", rank=rank, world_size=world)
        eng.init(0)
        return torch.device("cpu")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng.init(rank)
    return torch.device("cuda", rank)


def _random_circuit(orc, n, count, seed):
    rng = np.random.default_rng(seed)
    one = ["X", "Y", "Z", "H", "Phase", "RotX", "RotY", "RotZ"]
    two = ["CX", "CY", "CZ", "CH", "CPhase", "CRotX", "CRotY", "CRotZ", "Swap"]
    three = ["CSwap", "CCNot", "Or"]
    gates = []
    for _ in range(count):
        r = rng.random()
        if r < 0.45:
            name, q = one[rng.integers(len(one))], [int(rng.integers(n))]
        elif r < 0.85:
            name, q = two[rng.integers(len(two))], [int(x) for x in rng.choice(n, 2, replace=False)]
        else:
            name, q = three[rng.integers(len(three))], [int(x) for x in rng.choice(n, 3, replace=False)]
        if name in orc.HAS_ANGLE:
            gates.append((name, *q, float(np.float32(rng.uniform(-np.pi, np.pi)))))
        else:
            gates.append((name, *q))
    return orc.Circ(n, gates)


def worker(rank, world, backend, port):
    device = _setup(rank, world, backend, port)
    from afquantumsim_b200 import workloads as wl
    from afquantumsim_b200.sharded import ShardedState
    from oracle import oracle as orc
    from tests.lowering import lower_array

    g = int(np.log2(world))
    # 1. random circuits touching global qubits every way (target, control, diagonal, swap);
    #    on GPUs both remap back ends: in-place swaps over peer memory and NCCL half-shard send/recv
    modes = (True, False) if backend == "cuda" else (False,)
    for n, count, seed in ((g + 3, 60, 1), (g + 6, 150, 2), (g + 9, 200, 3), (12, 250, 4)):
        circ = _random_circuit(orc, n, count, seed)
        want = orc.simulate(orc.new_state(n), circ)
        for fuse in (True, False):
            for p2p in modes:
                st = ShardedState(n, device=device, fuse=fuse, p2p=p2p)
                if backend == "cuda" and world > 1:
                    assert st.p2p == p2p, "peer memory should be available between the GPUs of one node"
                st.apply_ops(lower_array(circ))
                got = st.gather()
                err = orc.rel_l2(got, want)
                assert err < TOL, (n, seed, fuse, p2p, err)
                assert abs(st.norm2() - 1) < 1e-4
                if world > 1 and n > g + 3:
                    assert st.stats["exchanges"] > 0
    # a compiled plan is reusable: same entry layout, same result
    n = 12
    circ = _random_circuit(orc, n, 200, 5)
    want = orc.simulate(orc.new_state(n), circ)
    st = ShardedState(n, device=device)
    plan = st.compile(lower_array(circ))
    for _ in range(2):
        st.set_basis(0)
        st.run(plan)
        assert orc.rel_l2(st.gather(), want) < TOL
        st.set_basis(0)
    # flat address space (GPUs only): one fused plan over the whole state, tiles that span GPUs go over NVLink
    if backend == "cuda" and world > 1:
        n = 19 + g
        u = np.random.default_rng(11).random(300, dtype=np.float32)
        for name, circ in (("random", _random_circuit(orc, n, 160, 6)), ("brickwork", orc.Circ(n, wl.brickwork(n, 8))),
                           ("qft", orc.Circ(n, wl.qft(n))), ("ghz", orc.Circ(n, wl.ghz(n)))):
            st = ShardedState(n, device=device)
            assert st.flat_state is not None, "the flat address space should be available between the GPUs of one node"
            x = 0 if name != "qft" else 0b1011001110100110101 & ((1 << n) - 1)
            st.set_basis(x)
            plan = st.compile(lower_array(circ))
            st.run(plan)
            want = orc.simulate(orc.new_state(n, x), circ)
            err = orc.rel_l2(st.gather(), want)
            assert err < TOL, (name, err)
            assert plan.n_exchanges > 0, name           # some pass had a rank bit in its tile
            if name == "ghz":
                assert np.array_equal(st.gather(), want)
            assert np.array_equal(st.sample(u), orc.sample(st.gather(), u, "exact")), name
            st.set_basis(x)                              # a compiled plan is reusable
            st.run(plan)
            assert orc.rel_l2(st.gather(), want) < TOL, name
            del plan, st
    # 2. BASELINE circuits: brickwork and QFT (QFT's CPhase ladder needs no exchange beyond the g H gates)
    n = 12
    st = ShardedState(n, device=device)
    st.apply_ops(lower_array(orc.Circ(n, wl.brickwork(n, 8))))
    want = orc.simulate(orc.new_state(n), orc.Circ(n, wl.brickwork(n, 8)))
    assert orc.rel_l2(st.gather(), want) < TOL
    # sampling and probabilities: bit-exact on identical states
    st2 = ShardedState(n, device=device)
    st2.buf.copy_(torch.from_numpy(want[rank << st2.n_local:(rank + 1) << st2.n_local]).to(device))
    u = np.concatenate([np.random.default_rng(9).random(2000, dtype=np.float32), np.array([0.0, 0.99999994], np.float32)])
    assert np.array_equal(st2.sample(u), orc.sample(want, u, "exact"))
    assert st2.prob_fixed() == orc.prob_fixed(want)
    for q in (0, g, n - 1):
        m = 1 << (n - 1 - q)
        assert st2.prob_fixed(1 << q, 1 << q) == orc.prob_fixed(want, m, m)
        assert st2.qubit_prob1(q) == orc.qubit_prob1(want, q)

    x = 0b101100111010 & ((1 << n) - 1)
    st = ShardedState(n, device=device)
    st.set_basis(x)
    st.apply_ops(lower_array(orc.Circ(n, wl.qft(n))))
    exchanges_qft = st.stats["exchanges"]
    want = orc.simulate(orc.new_state(n, x), orc.Circ(n, wl.qft(n)))
    assert orc.rel_l2(st.gather(), want) < TOL
    assert exchanges_qft <= g, exchanges_qft          # only H on the g global qubits moves data
    # GHZ: exact, and sampling across shards
    st = ShardedState(n, device=device)
    st.apply_ops(lower_array(orc.Circ(n, wl.ghz(n))))
    u = np.random.default_rng(3).random(500, dtype=np.float32)
    want = orc.simulate(orc.new_state(n), orc.Circ(n, wl.ghz(n)))
    assert np.array_equal(st.sample(u), orc.sample(want, u, "exact"))
    assert np.array_equal(st.gather(), want)
    if rank == 0:
        print("ok dist_cases world=%d backend=%s" % (world, backend), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="cpu")
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--port", type=int, default=29611)
    a = ap.parse_args()
    mp.spawn(worker, args=(a.world, a.backend, a.port), nprocs=a.world, join=True)

"""Multi-GPU check of STAGED spanning passes (sharded.py / flat.cu), run under torch.distributed:
  python tests/staged_cases.py --world 2
Brickwork and random circuits on the flat address space with staging on and off against the unsharded CPU oracle;
the two modes must agree bit for bit (same kernels, same arithmetic, only the source of the loads differs)."""
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


def worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from afquantumsim_b200 import engine as eng
    from afquantumsim_b200 import workloads as wl
    from afquantumsim_b200.sharded import ShardedState
    from oracle import oracle as orc
    from tests.dist_cases import _random_circuit
    from tests.lowering import lower_array
    eng.init(rank)
    os.environ["AQS_SHARD_SCHEDULE"] = "flat"       # (the whole-state plan: the schedule whose spanning passes are staged)
    g = int(np.log2(world))
    n = 27 + g          # 1 GiB per shard: the blocks of the staged passes stay above the 2 MiB mapping granularity
    for name, circ, jit in (("brickwork", orc.Circ(n, wl.brickwork(n, 6)), True), ("random", _random_circuit(orc, n, 90, 8), False)):
        want = orc.simulate(orc.new_state(n), circ) if rank == 0 else None
        outs = {}
        for mode in ("1", "0"):
            os.environ["AQS_STAGED"] = mode
            st = ShardedState(n, jit=jit)
            assert st.flat_state is not None
            plan = st.compile(lower_array(circ))
            staged = plan.steps[0][4]
            if mode == "1":
                assert plan.n_exchanges > 0 and len(staged) > 0, (name, plan.spans())
            else:
                assert not staged
            for _ in range(2):                   # a plan (views, staging buffer) is reusable
                st.set_basis(0)
                st.run(plan)
            outs[mode] = st.gather()
            st.close()
            del plan, st
        assert np.array_equal(outs["1"], outs["0"]), name
        if rank == 0:
            err = orc.rel_l2(outs["1"], want)
            assert err < 1e-5, (name, err)
            print(f"ok staged {name} world={world} rel_l2={err:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--port", type=int, default=29641)
    a = ap.parse_args()
    mp.spawn(worker, args=(a.world, a.port), nprocs=a.world, join=True)

"""Dev tool (torchrun): one sharded brickwork step on every data path.  python -m torch.distributed.run ... tools/sharded_modes.py [qubits_per_gpu]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402
from afquantumsim_b200.sharded import ShardedState  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng.init(local)
world, rank = dist.get_world_size(), dist.get_rank()
g = int(np.log2(world))
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 31) + g
ops = wl.to_ops(wl.brickwork(n, 20))
for name, kw, env in (("auto", dict(), {"AQS_SHARD_SCHEDULE": "auto"}), ("flat staged", dict(), {"AQS_STAGED": "1", "AQS_SHARD_SCHEDULE": "flat"}),
                      ("remap on flat", dict(), {"AQS_SHARD_SCHEDULE": "remap"}),
                      ("remap ipc", dict(flat=False, p2p=True), {}), ("remap nccl", dict(flat=False, p2p=False), {})):
    os.environ.update(env)
    st = ShardedState(n, jit=True, **kw)
    t0 = time.perf_counter()
    plan = st.compile(ops)
    tb = time.perf_counter() - t0
    for _ in range(2):
        st.set_basis(0)
        st.run(plan)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        st.set_basis(0)
        st.run(plan)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:12s} n={n} {float(t.item()):8.1f} ms/step  exchanges {plan.n_exchanges}  exchange GiB/rank {plan.exchange_bytes / 2**30:.1f}  passes {plan.n_passes}  "
              f"jit {plan.jit_ready()}  schedule {st.stats.get('schedule')}  norm2 {st.norm2():.6f}  build {tb:.1f}s", flush=True)
    else:
        st.norm2()
    st.close()
    del plan, st
    eng.pool_trim()
dist.barrier()
dist.destroy_process_group()

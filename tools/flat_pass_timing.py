"""Dev tool (multi-GPU, run under torchrun): per-pass device time of a sharded brickwork circuit on the flat
address space, split into passes that stay inside a shard and passes whose tiles span GPUs (NVLink).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/flat_pass_timing.py [qubits_per_gpu]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng  # noqa: E402
from afquantumsim_b200 import workloads as wl  # noqa: E402
from afquantumsim_b200.sharded import ShardedState  # noqa: E402

os.environ.setdefault("AQS_SHARD_SCHEDULE", "flat")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng.init(local)
world, rank = dist.get_world_size(), dist.get_rank()
g = int(np.log2(world))
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 30) + g
st = ShardedState(n, jit=(os.environ.get("AQS_NOJIT") is None))
assert st.flat_state is not None
plan = st.compile(wl.to_ops(wl.brickwork(n, 20)))
_, ep, spans, _, staged = plan.steps[0]
if rank == 0:
    print("jit passes", ep.jit_ready(), "of", len(spans))
for rep in range(3):
    st.set_basis(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(spans) + 1)]
    dist.barrier()
    ev[0].record()
    for i, j in enumerate(spans):
        if j or (i and spans[i - 1]):
            st._stream_barrier()
        if staged.get(i) is not None:
            st._run_staged(ep, i, staged[i])
        else:
            st.flat_state.run_shard(ep, i, 1, rank, g)
        ev[i + 1].record()
    st._stream_barrier()
    torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(spans))]
if rank == 0:
    for i, (j, t) in enumerate(zip(spans, ms)):
        print(f"pass {i:2d}  rank bits in tile {j}  {t:7.2f} ms" + ("  staged: %d chunks, %d copies" % (len(staged[i][1]), sum(len(c[0]) for c in staged[i][1])) if staged.get(i) else ""))
    loc = [t for j, t in zip(spans, ms) if not j]
    rem = [t for j, t in zip(spans, ms) if j]
    print(f"total {sum(ms):.1f} ms; {len(loc)} local passes {sum(loc):.1f} ms; {len(rem)} spanning passes {sum(rem):.1f} ms")
dist.barrier()
dist.destroy_process_group()

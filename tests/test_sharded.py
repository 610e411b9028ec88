"""Sharded states: world_size-2 (and 4) gloo runs on CPU with the oracle-backed engine
stand-in; the same cases over NCCL on real GPUs when at least two are visible."""
import os
import subprocess
import sys

import pytest
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(backend, world, port):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_cases.py"), "--backend", backend,
                        "--world", str(world), "--port", str(port)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "ok dist_cases" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_state_gloo_cpu(world):
    run("cpu", world, 29610 + world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_sharded_state_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    run("cuda", world, 29620 + world)


def test_staged_pass_geometry_covers_exactly_the_remote_inputs():
    """CPU: ShardedState._stage_pass (no GPU needed for the geometry).  For every rank and every spanning pass of a
    brickwork plan: the chunks' launches partition the rank's tiles, and the blocks copied for a chunk are exactly the
    amplitudes of that chunk's tiles that live in peer shards."""
    from types import SimpleNamespace

    from afquantumsim_b200 import engine as eng
    from afquantumsim_b200 import workloads as wl
    from afquantumsim_b200.sharded import ShardedState
    from tests.tile_emulator import deposit

    g, nl = 2, 27          # (plans only: no state of this size is allocated)
    n = nl + g
    plan = eng.Plan(n, wl.to_ops(wl.brickwork(n, 8)), eng.PLAN_FUSE)
    passes = plan.info()["n_fused_passes"]
    checked = 0
    for i in range(passes):
        if not plan.pass_span(i, g):
            continue
        tile = plan.pass_tile(i)
        nontile = [b for b in range(n) if b not in tile]
        for rank in range(1 << g):
            me = SimpleNamespace(n=n, n_local=nl, g=g, rank=rank)
            geo = ShardedState._stage_pass(me, plan, i)
            assert geo is not None
            blocks, chunks = geo
            fix_pos, fix_or = plan.shard_cut(i, rank, g)
            all_tiles = deposit(np.arange(1 << (n - len(tile) - len(fix_pos)), dtype=np.uint64), fix_pos) | np.uint64(fix_or)
            seen = []
            for copies, cpos, cor in chunks:
                t = deposit(np.arange(1 << (n - len(tile) - len(cpos)), dtype=np.uint64), cpos) | np.uint64(cor)
                seen.append(t)
                # base index of every tile of the chunk (non-tile bits deposited at their positions), a sample of them
                base = np.zeros(t.size, dtype=np.uint64)
                for c, b in enumerate(nontile):
                    base |= ((t >> np.uint64(c)) & np.uint64(1)) << np.uint64(b)
                sample = base[:: max(1, base.size // 64)]
                # every (tile base, rank-bit setting) that lands in a peer shard must lie inside a copied block
                rank_bits = [b for b in tile if b >= nl]
                for v in range(1 << len(rank_bits)):
                    off = sum(((v >> k) & 1) << b for k, b in enumerate(rank_bits))
                    idx = sample | np.uint64(off)
                    shard = idx >> np.uint64(nl)
                    for x in idx[shard != rank]:
                        byte = int(x) * 8
                        assert any(po <= byte < po + nb for po, nb in copies), (i, rank, hex(int(x)))
                total = sum(nb for _, nb in copies)
                expect = t.size * (1 << len(tile)) * 8 * ((1 << len(rank_bits)) - 1) // (1 << len(rank_bits))
                assert total == expect, (total, expect)
            seen = np.sort(np.concatenate(seen))
            assert np.array_equal(seen, np.sort(all_tiles)), "the chunks must partition the rank's tiles"
            assert len({b[0] for b in blocks}) == len(blocks) and all(b[0] % (2 << 20) == 0 and b[1] % (2 << 20) == 0 for b in blocks)
            checked += 1
    assert checked > 0

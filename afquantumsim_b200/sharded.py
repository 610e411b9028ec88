"""States larger than one GPU: the 2^n amplitudes are sharded over R = 2^g ranks
(one process per GPU, torch.distributed over NCCL / NVLink).

Layout.  Index bit p of an amplitude is *local* for p < n_local = n - g and
*global* otherwise: global bits are the bits of the rank number.  Initially
logical qubit q sits at bit n-1-q (reference convention, src/quantum.cpp:546),
so the first g API qubits are global.  `pos[q]` tracks where each logical qubit
currently lives; gates never move data unless they must:

  * target local                     -> plain local op (fused plans, no communication)
  * control on a global bit          -> a per-rank predicate: the rank runs the op or skips it
  * diagonal gate on a global target -> local multiply by the rank's diagonal entry
  * non-diagonal gate on a global target -> the qubit is first swapped with a local
    one: the rank pairs (r, r ^ 2^j) exchange HALF a shard each way with NCCL
    send/recv; afterwards the gate is local.  The swap is lazy (the map just changes),
    victims are chosen by farthest next use, and everything executable between two
    exchanges is batched into one fused local plan.

Measurement first restores the canonical layout, then uses the exact-sum contract
(integer probability sums are order-free, so R-GPU sampling is bit-identical to one GPU).

The reference has no multi-device support; this module is the
north-star extension, with a 64-bit index API beside the 30-qubit drop-in one.
"""
from __future__ import annotations

import atexit
import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import engine as eng

FIX62 = np.float32(2.0 ** 62)


def fix62(u: np.ndarray) -> np.ndarray:
    """trunc(u * 2^62) exactly as the engine and the oracle compute it."""
    return (np.asarray(u, dtype=np.float32) * FIX62).astype(np.uint64)


class _Op:
    __slots__ = ("kind", "t", "t2", "ctrl", "m", "nd", "dg")

    def __init__(self, rec):
        self.kind = int(rec["kind"])
        self.t = int(rec["target"])
        self.t2 = int(rec["target2"])
        cm, cv = int(rec["ctrl_mask"]), int(rec["ctrl_value"])
        self.ctrl = [(q, (cv >> q) & 1) for q in range(64) if (cm >> q) & 1]
        self.m = np.array(rec["m"], dtype=np.float32).view(np.complex64).copy()
        if self.kind == eng.OP_U2 and self.m[1] == 0 and self.m[2] == 0:
            self.kind = eng.OP_DIAG
        cq = {q for q, _ in self.ctrl}
        if self.kind == eng.OP_DIAG:
            self.nd, self.dg = set(), cq | {self.t}
        elif self.kind == eng.OP_SWAP:
            self.nd, self.dg = {self.t, self.t2}, cq
        else:
            self.nd, self.dg = {self.t}, cq


# Flat address spaces are recycled like the engine's state buffers: creating one means an 8 GiB physical
# allocation, a file-descriptor exchange between the processes and page-table updates for every peer shard,
# and the API's normal use is a new state per run.  Keyed by geometry; reused only if EVERY rank has one.
_FLAT_CACHE: dict = {}


def _drop_flat_cache():
    while _FLAT_CACHE:
        _FLAT_CACHE.popitem()[1].close()


atexit.register(_drop_flat_cache)


class ShardedPlan:
    """A circuit compiled for one entry layout of a ShardedState: fused local plans interleaved with
    global-qubit remaps (the sharded counterpart of QCircuit::compile, src/quantum.cpp:199-210).
    steps: ("plan", eng.Plan, n_ops) | ("p2p", rank_bits, local_bits) | ("nccl", rank_bit)."""
    __slots__ = ("steps", "entry_pos", "exit_pos", "n_exchanges", "exchange_bytes", "n_local_ops", "n_passes")

    def __init__(self, steps, entry_pos, exit_pos):
        self.steps, self.entry_pos, self.exit_pos = steps, list(entry_pos), list(exit_pos)
        self.n_exchanges = self.exchange_bytes = self.n_local_ops = self.n_passes = 0

    def jit_ready(self) -> int:
        """fused passes that would run on their specialised kernel now (flat plans)"""
        return sum(step[1].jit_ready() for step in self.steps if step[0] in ("flat", "plan"))

    def spans(self) -> list:
        """rank bits inside the tile of every pass of a flat plan (0 = the pass stays inside the shards)"""
        return [j for step in self.steps if step[0] == "flat" for j in step[2]]


class _CudaMem:
    """__cuda_array_interface__ view of engine-owned device memory (lets torch / NCCL address a shard)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class _Sched:
    """Layout bookkeeping while a schedule is emitted: local op records pile up in `pending` until a remap
    (or the end) turns them into one fused plan."""

    def __init__(self, st: "ShardedState", wait: bool = True):
        self.st = st
        self.wait = wait        # do fused local plans wait for their specialised kernels?
        self.pos = list(st.pos)
        self.steps: list = []
        self.pending: List[np.ndarray] = []
        self.entry = list(st.pos)

    def qubit_at(self, p: int) -> int:
        return self.pos.index(p)

    def flush(self):
        if self.pending:
            st = self.st
            ops = np.concatenate(self.pending)
            flags = (eng.PLAN_FUSE | (st._jit_flags(self.wait) if self.wait is not None else 0)) if st.fuse else 0
            self.steps.append(("plan", eng.Plan(st.n_local, ops, flags), len(ops)))
            self.pending = []

    def local_swap(self, pa: int, pb: int):
        """exchange the logical qubits at local index bits pa and pb (a SWAP folded into the pending batch)"""
        if pa == pb:
            return
        st = self.st
        self.pending.append(eng.op_record(eng.OP_SWAP, st._local_qubit(pa), target2=st._local_qubit(pb)))
        qa, qb = self.qubit_at(pa), self.qubit_at(pb)
        self.pos[qa], self.pos[qb] = pb, pa
        st.stats["local_swaps"] += 1

    def exchange(self, pairs):
        """pairs: [(q_global, q_local)] logical qubits to trade places, all at once where peer memory allows"""
        st = self.st
        k = len(pairs)
        lbits = [self.pos[ql] for _, ql in pairs]
        if st.p2p and st.n_local >= k + 2 and all(b >= 1 for b in lbits):
            self.flush()
            self.steps.append(("p2p", [self.pos[qg] - st.n_local for qg, _ in pairs], lbits))
            for qg, ql in pairs:
                self.pos[qg], self.pos[ql] = self.pos[ql], self.pos[qg]
            return
        top = st.n_local - 1
        for qg, ql in pairs:
            self.local_swap(self.pos[ql], top)      # the half that leaves is then one contiguous block
            self.flush()
            pg = self.pos[qg]
            self.steps.append(("nccl", pg - st.n_local))
            self.pos[qg], self.pos[ql] = top, pg

    def finish(self) -> ShardedPlan:
        self.flush()
        plan = ShardedPlan(self.steps, self.entry, self.pos)
        shard_bytes = 8 << self.st.n_local
        for step in self.steps:
            if step[0] == "plan":
                plan.n_local_ops += step[2]
                plan.n_passes += int(step[1].info()["n_launches"])
            elif step[0] == "p2p":
                k = len(step[1])
                plan.n_exchanges += 1
                plan.exchange_bytes += shard_bytes * ((1 << k) - 1) // (1 << (k + 1))
            else:
                plan.n_exchanges += 1
                plan.exchange_bytes += shard_bytes // 2
        return plan


class ShardedState:
    """A 2^n complex64 state vector sharded over the ranks of a torch.distributed group."""

    def __init__(self, n_qubits: int, device: Optional[torch.device] = None, group=None, fuse: bool = True,
                 p2p: Optional[bool] = None, flat: Optional[bool] = None, jit: Optional[bool] = None):
        self.group = group
        self.jit = jit          # specialised pass kernels: None = automatic (shards of 2^26 amplitudes and more)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.g = int(math.log2(self.world))
        if (1 << self.g) != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n = int(n_qubits)
        self.n_local = self.n - self.g
        if self.n_local < 2:
            raise ValueError("each shard needs at least 2 local qubits")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.fuse = fuse
        self.stats = {"exchanges": 0, "exchange_bytes": 0, "local_plans": 0, "local_ops": 0, "local_swaps": 0}
        self.stage = None           # half-shard staging buffer of the NCCL path, made on first use
        self._copy_stream = None    # staged passes: the stream the copy engines work on
        self.p2p = False
        self.peer_ptrs: List[int] = []
        self.flat = None            # eng.FlatSpace: every shard of the node in one virtual address range
        self.flat_state = None      # engine handle on the WHOLE 2^n state (fused plans run on it shard by shard)
        if p2p is None:
            p2p = os.environ.get("AQS_SHARD_P2P", "1") != "0"
        if flat is None:
            flat = p2p and os.environ.get("AQS_SHARD_FLAT", "1") != "0"
        if self.device.type == "cuda" and flat and fuse and self.world > 1 and self._open_flat():
            pass
        elif self.device.type == "cuda":
            # engine-owned shard (the base of a cudaMalloc allocation, so that it can be exported over CUDA IPC);
            # torch sees it through __cuda_array_interface__
            self.state = eng.State(self.n_local)
            self._mem = _CudaMem(self.state.device_ptr(), 2 << self.n_local)
            self.buf = torch.view_as_complex(torch.as_tensor(self._mem, device=self.device).view(-1, 2))
            self.state.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            if p2p and self.world > 1:
                self._open_peers()
        else:
            self.buf = torch.zeros(1 << self.n_local, dtype=torch.complex64, device=self.device)
            self.state = eng.State.wrap(self.n_local, self.buf.data_ptr())
        self.pos: List[int] = [self.n - 1 - q for q in range(self.n)]   # logical qubit -> index bit
        self.set_basis(0)

    def _open_flat(self) -> bool:
        """Map the shards of all ranks back to back into one virtual address range (engine: flat.cu).
        Needs shards that are a multiple of the 2 MiB mapping granularity; all-or-nothing across ranks."""
        shard_bytes = 8 << self.n_local
        ok, space, fds = 1, None, {}
        if shard_bytes % (2 << 20) or self.n_local < 13 + self.g:
            return False            # the same answer on every rank: no consensus round needed

        def agreed(value: int) -> bool:
            flag = torch.tensor([value], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            return bool(int(flag.item()))

        self._flat_key = (shard_bytes, self.world, self.rank, id(self.group))
        space = _FLAT_CACHE.pop(self._flat_key, None)
        if not agreed(1 if space is not None else 0):
            if space is not None:
                space.close()
            space = None
            try:
                space = eng.FlatSpace(shard_bytes, self.world, self.rank)
            except eng.EngineError:
                ok = 0
            if not agreed(ok):
                if space is not None:
                    space.close()
                return False
            try:
                fds = self._exchange_fds(space.fd)
                for r, fd in fds.items():
                    space.attach(r, fd)
            except (eng.EngineError, OSError):
                ok = 0
            finally:
                for fd in fds.values():
                    os.close(fd)
            if not agreed(ok):
                space.close()
                return False
        base, own = space.pointers()
        self.flat = space
        self.flat_state = eng.State.wrap(self.n, base)
        self.state = eng.State.wrap(self.n_local, own)
        self._mem = _CudaMem(own, 2 << self.n_local)
        self.buf = torch.view_as_complex(torch.as_tensor(self._mem, device=self.device).view(-1, 2))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.state.set_stream(stream)
        self.flat_state.set_stream(stream)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        # every shard is addressable through the flat range: the in-place remap kernel (aqs_peer_bitswap) needs no CUDA IPC
        shard_bytes = 8 << self.n_local
        self.peer_ptrs = [base + r * shard_bytes for r in range(self.world)]
        self.p2p = True
        return True

    def close(self):
        """Give the flat address space back for the next state of the same geometry."""
        space, self.flat = self.flat, None
        if space is not None:
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
            self.flat_state = None
            old = _FLAT_CACHE.pop(self._flat_key, None)
            if old is not None:
                old.close()
            _FLAT_CACHE[self._flat_key] = space

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _exchange_fds(self, fd: int) -> dict:
        """Hand this rank's shard (a POSIX file descriptor) to every other rank of the node and collect
        theirs: SCM_RIGHTS over unix-domain sockets."""
        import socket
        import threading
        import uuid
        token = [uuid.uuid4().hex[:12] if self.rank == 0 else None]
        dist.broadcast_object_list(token, src=0, group=self.group)
        path = lambda r: f"/tmp/aqs_flat_{token[0]}_{r}.sock"
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path(self.rank))
        srv.listen(self.world)
        srv.settimeout(120)

        def serve():
            for _ in range(self.world - 1):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"f"], [fd])
                conn.close()

        th = threading.Thread(target=serve, daemon=True)
        th.start()
        dist.barrier(group=self.group)          # every rank is listening
        fds = {}
        try:
            for r in range(self.world):
                if r == self.rank:
                    continue
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                c.settimeout(120)
                c.connect(path(r))
                _, got, _, _ = socket.recv_fds(c, 16, 1)
                c.close()
                fds[r] = got[0]
        finally:
            th.join(timeout=130)
            srv.close()
            try:
                os.unlink(path(self.rank))
            except OSError:
                pass
        return fds

    def _open_peers(self):
        """Exchange CUDA IPC handles of the shards; peer memory is used only if EVERY rank mapped every shard."""
        ok = 1
        ptrs: List[int] = []
        try:
            handle = self.state.ipc_export()
        except eng.EngineError:
            handle, ok = bytes(eng.IPC_HANDLE_BYTES), 0
        t = torch.zeros(self.world, eng.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        t[self.rank] = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(self.device)
        dist.all_reduce(t, group=self.group)
        handles = t.cpu().numpy()
        if ok:
            try:
                for r in range(self.world):
                    ptrs.append(self.state.device_ptr() if r == self.rank else eng.ipc_open(handles[r].tobytes()))
            except eng.EngineError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.p2p = bool(int(flag.item()))
        self.peer_ptrs = ptrs if self.p2p else []
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ helpers
    def _qubit_at(self, p: int) -> int:
        return self.pos.index(p)

    def _rank_bit(self, p: int) -> int:
        return (self.rank >> (p - self.n_local)) & 1

    def _local_qubit(self, p: int) -> int:
        """engine (API-style) qubit number of local index bit p"""
        return self.n_local - 1 - p

    def _translate(self, op: _Op, pos: Sequence[int]) -> List[np.ndarray]:
        """One logical op -> local engine ops for THIS rank under the layout `pos`."""
        ctrl_local, cval = [], 0
        for q, v in op.ctrl:
            p = pos[q]
            if p >= self.n_local:
                if self._rank_bit(p) != v:
                    return []                      # control not satisfied on this rank
            else:
                lq = self._local_qubit(p)
                ctrl_local.append(lq)
                cval |= v << lq
        if op.kind == eng.OP_DIAG:
            p = pos[op.t]
            if p >= self.n_local:
                f = op.m[3] if self._rank_bit(p) else op.m[0]
                if f == 1:
                    return []
                free = next(lq for lq in range(self.n_local) if lq not in ctrl_local)
                return [eng.op_record(eng.OP_DIAG, free, [f, 0, 0, f], ctrl_local, ctrl_value=cval)]
            return [eng.op_record(eng.OP_DIAG, self._local_qubit(p), op.m, ctrl_local, ctrl_value=cval)]
        if op.kind == eng.OP_SWAP:
            return [eng.op_record(eng.OP_SWAP, self._local_qubit(pos[op.t]), controls=ctrl_local,
                                  target2=self._local_qubit(pos[op.t2]), ctrl_value=cval)]
        return [eng.op_record(op.kind, self._local_qubit(pos[op.t]), op.m, ctrl_local, ctrl_value=cval)]

    # ------------------------------------------------------------------ remaps
    def _stream_barrier(self):
        """Cross-rank barrier ORDERED ON THE STREAM (no host synchronisation): a one-element all_reduce."""
        dist.all_reduce(self._flag, group=self.group)

    def _p2p_swap(self, rank_bits: Sequence[int], local_bits: Sequence[int]):
        """Trade k rank bits with k local index bits in place over NVLink peer memory (aqs_peer_bitswap)."""
        k = len(rank_bits)
        my = sum(((self.rank >> j) & 1) << i for i, j in enumerate(rank_bits))
        members = []
        for v in range(1 << k):
            r = self.rank
            for i, j in enumerate(rank_bits):
                r = (r & ~(1 << j)) | (((v >> i) & 1) << j)
            members.append(self.peer_ptrs[r])
        self._stream_barrier()      # every member has finished the kernels that precede the remap
        self.state.peer_bitswap(members, local_bits, my)
        self._stream_barrier()      # nobody touches its shard while a partner still writes into it

    def _nccl_half_exchange(self, rank_bit: int):
        """Trade rank bit `rank_bit` with the TOP local bit: half a shard each way with send/recv."""
        b = (self.rank >> rank_bit) & 1
        peer = self.rank ^ (1 << rank_bit)
        halves = self.buf.view(2, -1)
        out_half = halves[1 - b]                 # the half whose top local bit differs from my rank bit
        if self.stage is None:
            self.stage = torch.empty(1 << (self.n_local - 1), dtype=torch.complex64, device=self.device)
        if self.device.type == "cuda":
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, out_half, peer, self.group),
                                           dist.P2POp(dist.irecv, self.stage, peer, self.group)])
            for r in reqs:
                r.wait()
        else:   # gloo: order the blocking pair by rank to avoid a deadlock
            if self.rank < peer:
                dist.send(out_half, peer, self.group); dist.recv(self.stage, peer, self.group)
            else:
                tmp = out_half.clone()
                dist.recv(self.stage, peer, self.group); dist.send(tmp, peer, self.group)
        out_half.copy_(self.stage)

    @staticmethod
    def _next_use(q: int, ops: Sequence[_Op]) -> int:
        for i, op in enumerate(ops):
            if q in op.nd:
                return i
        return 1 << 30

    # ------------------------------------------------------------------ gates
    def _jit_flags(self, wait: bool) -> int:
        """specialised pass kernels: on by request, else for shards of 2^26 amplitudes and more (like the host layer)"""
        on = self.jit if self.jit is not None else self.n_local >= 26
        if not on or self.device.type != "cuda":
            return 0
        return eng.PLAN_JIT if wait else eng.PLAN_JIT_ASYNC

    def compile(self, records: np.ndarray, wait: bool = True) -> ShardedPlan:
        """Schedule primitive ops (struct aqs_op records over all n qubits, API numbering) for the CURRENT
        layout: everything executable between two remaps becomes one fused local plan.  On a flat
        address space there is nothing to schedule: ONE fused plan over all n qubits, no remaps."""
        if self.flat_state is not None:
            # Two schedules exist on a flat address space.  "flat": ONE fused plan over all n qubits, passes whose tiles
            # hold rank bits move their remote part over NVLink — there and back, every such pass.  "remap": lazy global
            # qubit swaps (one in-place exchange moves half a shard once, the qubit then stays local) around fused LOCAL
            # plans.  Which one moves fewer bytes depends on the circuit (brickwork-32, depth 20 on 2 GPUs: 48 GiB per
            # direction against 8), so both are planned — without their kernels, that takes milliseconds — and priced.
            mode = os.environ.get("AQS_SHARD_SCHEDULE", "auto")
            cands = {}
            canonical = self.pos == [self.n - 1 - q for q in range(self.n)]      # (the whole-state plan assumes it)
            if mode in ("auto", "flat") and canonical:
                cands["flat"] = self._compile_flat(records, 0, stage=False)
            if mode in ("auto", "remap") or not cands:
                cands["remap"] = self._compile_remap(records, None)
            shard_bytes = 8 << self.n_local

            def cost(pl):      # seconds: streaming passes over the shard + NVLink bytes per direction (reads and the peers' writes)
                return pl.n_passes * 2.0 * shard_bytes / 4.5e12 + 2.0 * pl.exchange_bytes / 0.7e12

            pick = min(cands, key=lambda k: cost(cands[k]))
            self.stats["schedule"] = pick
            jit = self._jit_flags(wait)
            if not jit:
                return cands[pick]
            del cands
            return self._compile_flat(records, jit) if pick == "flat" else self._compile_remap(records, wait)
        return self._compile_remap(records, wait)

    def _compile_flat(self, records: np.ndarray, jit_flags: int, stage: bool = True) -> ShardedPlan:
        recs = np.ascontiguousarray(records, dtype=eng.OP_DTYPE)
        if True:
            plan = ShardedPlan([], self.pos, self.pos)
            if len(recs):
                ep = eng.Plan(self.n, recs, eng.PLAN_FUSE | jit_flags)
                n_passes = int(ep.info()["n_fused_passes"])
                spans = [ep.pass_span(i, self.g) for i in range(n_passes)]
                # spanning passes run STAGED where their geometry allows (bulk copies of the peers' blocks into local
                # staging memory, pipelined with the kernel; AQS_STAGED=0 reads the peers from inside the kernel instead)
                staged = {}
                if stage and os.environ.get("AQS_STAGED", "1") != "0":
                    for i, j in enumerate(spans):
                        geo = self._stage_pass(ep, i) if j else None
                        if geo is not None:
                            key = tuple(geo[0])
                            if key not in self.flat._views:
                                self.flat._views[key] = self.flat.view(geo[0])
                            staged[i] = (self.flat._views[key], geo[1])
                plan.steps.append(("flat", ep, spans, len(recs), staged))
                plan.n_local_ops, plan.n_passes = len(recs), n_passes
                # NVLink bytes this rank writes: in a pass whose tile holds j rank bits, (2^j - 1) / 2^j of the
                # amplitudes it processes (one shard's worth) live on peers
                plan.n_exchanges = sum(1 for j in spans if j)
                plan.exchange_bytes = sum((8 << self.n_local) * ((1 << j) - 1) // (1 << j) for j in spans)
            return plan

    def _compile_remap(self, records: np.ndarray, wait) -> ShardedPlan:
        """wait: True / False = fused local plans with specialised kernels (waiting for them or not), None = without"""
        sch = _Sched(self, wait)
        pos = sch.pos
        remaining = [_Op(r) for r in np.ascontiguousarray(records, dtype=eng.OP_DTYPE)]
        while remaining:
            batch, rest = [], []
            blocked_nd, blocked_d = set(), set()
            for op in remaining:
                conflict = (op.nd & (blocked_nd | blocked_d)) or (op.dg & blocked_nd)
                local = all(pos[q] < self.n_local for q in op.nd)
                if not conflict and local:
                    batch.append(op)
                else:
                    blocked_nd |= op.nd
                    blocked_d |= op.dg
                    rest.append(op)
            for op in batch:
                sch.pending += self._translate(op, pos)
            remaining = rest
            if not remaining:
                break
            # bring in the global qubits needed soonest; evict the local qubits needed latest
            wanted: List[int] = []
            for op in remaining:
                for q in sorted(op.nd):
                    if pos[q] >= self.n_local and q not in wanted:
                        wanted.append(q)
                if len(wanted) >= self.g:
                    break
            protect = set()
            for op in remaining[:1]:
                protect |= op.nd
            pairs, victims = [], set()
            for qg in wanted[: self.g]:
                locals_ = [q for q in range(self.n) if pos[q] < self.n_local and q not in protect and q not in wanted
                           and q not in victims]
                if self.p2p:
                    # peer-memory remaps move runs of 2^bit amplitudes: keep them long when there is a choice
                    high = [q for q in locals_ if pos[q] >= 6]
                    locals_ = high or [q for q in locals_ if pos[q] >= 1] or locals_
                victim = max(locals_, key=lambda q: (self._next_use(q, remaining), pos[q]))
                victims.add(victim)
                pairs.append((qg, victim))
            sch.exchange(pairs)
        return sch.finish()

    def _bind_stream(self):
        """The engine states follow torch's CURRENT stream (the torch ops, NCCL calls and stream-ordered barriers of this
        class run there): re-bound at every entry point, so that a caller's `with torch.cuda.stream(...)` keeps kernels,
        copies and collectives on one stream."""
        if self.device.type != "cuda":
            return
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != getattr(self, "_bound_stream", None):
            self.state.set_stream(s)
            if self.flat_state is not None:
                self.flat_state.set_stream(s)
            self._bound_stream = s

    def run(self, plan: ShardedPlan) -> None:
        """Execute a plan compiled for the current layout (asynchronous on the stream)."""
        self._bind_stream()
        if self.pos != plan.entry_pos:
            raise ValueError("the plan was compiled for a different qubit layout; compile() again")
        for step in plan.steps:
            if step[0] == "plan":
                self.state.run(step[1])
                self.stats["local_plans"] += 1
                self.stats["local_ops"] += step[2]
            elif step[0] == "p2p":
                self._p2p_swap(step[1], step[2])
            elif step[0] == "flat":
                self._run_flat(step[1], step[2], step[4])
                self.stats["local_plans"] += 1
                self.stats["local_ops"] += step[3]
            else:
                self._nccl_half_exchange(step[1])
        self.pos = list(plan.exit_pos)
        self.stats["exchanges"] += plan.n_exchanges
        self.stats["exchange_bytes"] += plan.exchange_bytes

    # ------------------------------------------------------------------ staged passes
    def _stage_pass(self, ep: "eng.Plan", i: int, max_chunk_bits: int = 3):
        """Geometry of a STAGED run of spanning pass i (engine: flat.cu).  The tiles this rank runs are those whose
        pinned bits (aqs_plan_shard_cut) match the rank; the peers' amplitudes they need form, in every involved peer
        shard, the set { local index x : x[b] = v_b for the pinned spare bits b } — large contiguous blocks.  The pass
        is cut into chunks by pinning up to three more high local bits.  Returns (view blocks, chunks) with
        chunks = [(copies, fix_pos, fix_or)], copies = [(state byte offset, bytes)], or None when
        the blocks would be smaller than the 2 MiB mapping granularity (the pass then reads its peers directly)."""
        n, nl, g, rank = self.n, self.n_local, self.g, self.rank
        tile = set(ep.pass_tile(i))
        nontile = [b for b in range(n) if b not in tile]              # compact tile-number position c <-> index bit nontile[c]
        fix_pos, fix_or = ep.shard_cut(i, rank, g)
        pinned = {nontile[c]: (fix_or >> c) & 1 for c in fix_pos}
        spare = {b: v for b, v in pinned.items() if b < nl}           # local bits pinned because a rank bit sits in the tile
        in_tile_rank = [b for b in range(nl, n) if b in tile]
        if not in_tile_rank or len(spare) != len(in_tile_rank):
            return None
        free_hi = [b for b in range(nl - 1, -1, -1) if b not in tile and b not in pinned]
        gran_amps = (2 << 20) // 8
        chunk_bits = []
        for b in free_hi[:max_chunk_bits]:
            if (1 << min([b] + list(spare))) >= gran_amps:
                chunk_bits.append(b)
        lowest = min(list(spare) + chunk_bits)
        if (1 << lowest) < gran_amps:
            return None
        fixed_bits = sorted(list(spare) + chunk_bits)
        free_above = [b for b in range(lowest + 1, nl) if b not in fixed_bits]
        if len(free_above) > 10:
            return None
        peers = []
        for v in range(1 << len(in_tile_rank)):
            s_ = rank
            for k, b in enumerate(in_tile_rank):
                s_ = (s_ & ~(1 << (b - nl))) | (((v >> k) & 1) << (b - nl))
            if s_ != rank:
                peers.append(s_)
        shard_bytes = 8 << nl
        blocks, chunks, staged_bytes = [], [], 0
        for cv in range(1 << len(chunk_bits)):
            vals = dict(spare)
            for k, b in enumerate(chunk_bits):
                vals[b] = (cv >> k) & 1
            base = sum(v << b for b, v in vals.items())
            copies = []
            for fa in range(1 << len(free_above)):
                start = base + sum(((fa >> k) & 1) << b for k, b in enumerate(free_above))
                for s_ in peers:
                    state_off = s_ * shard_bytes + start * 8
                    blocks.append((state_off, 8 << lowest))
                    copies.append((state_off, 8 << lowest))
                    staged_bytes += 8 << lowest
            # the launch: the rank's pinned bits plus this chunk's
            pos = dict(pinned)
            for k, b in enumerate(chunk_bits):
                pos[b] = (cv >> k) & 1
            cpos = sorted(nontile.index(b) for b in pos)
            cor = sum(pos[nontile[c]] << c for c in cpos)
            chunks.append((copies, cpos, cor))
        assert staged_bytes <= shard_bytes
        return blocks, chunks

    def _run_staged(self, ep: "eng.Plan", i: int, staged):
        """One spanning pass, staged: the copy engines fetch chunk c + 1's remote blocks while chunk c computes."""
        view, chunks = staged
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        base, _ = self.flat.pointers()
        ev0 = torch.cuda.Event()
        ev0.record(main)
        cs.wait_event(ev0)                      # (after the barrier: every peer has finished the previous pass)
        events = []
        exp = os.environ.get("AQS_STAGED_EXP", "")          # (dev experiments: timing of the parts; results are wrong)
        for copies, _, _ in chunks:
            for po, nb in copies:
                if exp not in ("nocopy", "nocopy_localstore"):
                    eng.memcpy_async(view + po, base + po, nb, cs.cuda_stream)       # peer HBM -> the view's local backing
            ev = torch.cuda.Event()
            ev.record(cs)
            events.append(ev)
        target = self.flat_state
        if exp in ("localstore", "nocopy_localstore"):
            if not hasattr(self, "_exp_state"):
                self._exp_state = {}
            if view not in self._exp_state:
                self._exp_state[view] = eng.State.wrap(self.n, view)
                self._exp_state[view].set_stream(main.cuda_stream)
            target = self._exp_state[view]
        for (copies, cpos, cor), ev in zip(chunks, events):
            main.wait_event(ev)
            if exp != "nokernel":
                target.run_tiles(ep, i, view, cpos, cor)

    def _run_flat(self, plan: "eng.Plan", spans: Sequence[int], staged=None):
        """This rank's share of every pass of a whole-state fused plan.  A pass whose tile holds no rank
        bit touches only this rank's shard; one that does reads and writes peer memory over NVLink, so a
        stream-ordered barrier separates it from its neighbours (all of it asynchronous on the stream)."""
        i, n, dirty = 0, len(spans), False
        while i < n:
            if spans[i]:
                self._stream_barrier()
                if staged is not None and staged.get(i) is not None:
                    self._run_staged(plan, i, staged[i])
                else:
                    self.flat_state.run_shard(plan, i, 1, self.rank, self.g)
                dirty = True
                i += 1
            else:
                j = i
                while j < n and not spans[j]:
                    j += 1
                if dirty:
                    self._stream_barrier()
                    dirty = False
                self.flat_state.run_shard(plan, i, j - i, self.rank, self.g)
                i = j
        if dirty:
            self._stream_barrier()

    def apply_ops(self, records: np.ndarray):
        """Apply primitive ops: compile for the current layout (kernels of new pass shapes compile in the background), then run."""
        self.run(self.compile(records, wait=False))

    def run_circuit(self, qc) -> None:
        """simulate() for an afquantumsim_b200.aqs.QCircuit on more qubits than one GPU holds."""
        self.apply_ops(qc.ops())

    # ------------------------------------------------------------------ layout
    def canonicalize(self):
        """Bring every logical qubit q back to index bit n-1-q."""
        home = lambda q: self.n - 1 - q
        if all(self.pos[q] == home(q) for q in range(self.n)):
            return
        sch = _Sched(self)
        pos = sch.pos
        # 1. global positions
        for pg in range(self.n - 1, self.n_local - 1, -1):
            q_home = self.n - 1 - pg
            if pos[q_home] == pg:
                continue
            if pos[q_home] >= self.n_local:      # sits on another rank bit: bring it local first
                locals_ = [q for q in range(self.n) if pos[q] < self.n_local]
                sch.exchange([(q_home, max(locals_, key=lambda q: pos[q]))])
            sch.exchange([(sch.qubit_at(pg), q_home)])
        # 2. local permutation (SWAPs, fused into one plan)
        for p in range(self.n_local):
            q_home = self.n - 1 - p
            if pos[q_home] != p:
                sch.local_swap(pos[q_home], p)
        assert all(pos[q] == home(q) for q in range(self.n))
        self.run(sch.finish())

    # ------------------------------------------------------------------ measurement
    def _allreduce_i64(self, value: int) -> int:
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def norm2(self) -> float:
        v = self.state.norm2()
        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return float(t.item())

    def prob_fixed(self, qubit_mask: int = 0, qubit_value: int = 0) -> int:
        """Exact probability mass (2^-62 units) of the indices whose masked qubits read `value`."""
        lmask = lval = 0
        for q in range(self.n):
            if not (qubit_mask >> q) & 1:
                continue
            v = (qubit_value >> q) & 1
            p = self.pos[q]
            if p >= self.n_local:
                if self._rank_bit(p) != v:
                    return self._allreduce_i64(0)
            else:
                lq = self._local_qubit(p)
                lmask |= 1 << lq
                lval |= v << lq
        return self._allreduce_i64(self.state.prob_fixed(lmask, lval))

    def qubit_prob1(self, qubit: int) -> float:
        return self.prob_fixed(1 << qubit, 1 << qubit) * 2.0 ** -62

    def sample(self, u: np.ndarray) -> np.ndarray:
        """Outcome index (64-bit) per uniform draw: first k with cumulative probability > u, 0 if none."""
        self._bind_stream()
        self.canonicalize()
        U = fix62(u)
        mine = self.state.prob_fixed()
        if self.world > 1:
            t = torch.zeros(self.world, dtype=torch.int64, device=self.device)
            t[self.rank] = mine
            dist.all_reduce(t, group=self.group)
            totals = t.cpu().numpy().astype(np.uint64)
        else:
            totals = np.array([mine], dtype=np.uint64)
        incl = np.cumsum(totals, dtype=np.uint64)
        owner = np.searchsorted(incl, U, side="right")          # first rank with incl > U
        out = np.zeros(U.size, dtype=np.int64)
        sel = np.nonzero(owner == self.rank)[0]
        if sel.size:
            excl = incl[self.rank] - totals[self.rank]
            loc = self.state.sample_fixed(U[sel] - excl)
            assert not np.any(loc == np.uint64(0xFFFFFFFFFFFFFFFF))
            out[sel] = (np.int64(self.rank) << self.n_local) | loc.astype(np.int64)
        if self.world > 1:
            t = torch.from_numpy(out).to(self.device)
            dist.all_reduce(t, group=self.group)
            out = t.cpu().numpy()
        return out.astype(np.uint64)

    def gather(self) -> Optional[np.ndarray]:
        """The whole state on every rank (tests / small n only)."""
        self.canonicalize()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self.world == 1:
            return self.buf.cpu().numpy().copy()
        parts = [torch.empty_like(self.buf) for _ in range(self.world)]
        dist.all_gather(parts, self.buf, group=self.group)
        return torch.cat(parts).cpu().numpy()

    def set_basis(self, index: int = 0):
        self._bind_stream()
        self.buf.zero_()
        if (index >> self.n_local) == self.rank:
            self.buf[index & ((1 << self.n_local) - 1)] = 1.0
        self.pos = [self.n - 1 - q for q in range(self.n)]

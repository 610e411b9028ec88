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

import math
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


class ShardedState:
    """A 2^n complex64 state vector sharded over the ranks of a torch.distributed group."""

    def __init__(self, n_qubits: int, device: Optional[torch.device] = None, group=None, fuse: bool = True):
        self.group = group
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
        self.buf = torch.zeros(1 << self.n_local, dtype=torch.complex64, device=self.device)
        self.stage = torch.empty(1 << (self.n_local - 1), dtype=torch.complex64, device=self.device)
        if self.rank == 0:
            self.buf[0] = 1.0
        self.state = eng.State.wrap(self.n_local, self.buf.data_ptr())
        if self.device.type == "cuda":
            self.state.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.pos: List[int] = [self.n - 1 - q for q in range(self.n)]   # logical qubit -> index bit
        self.stats = {"exchanges": 0, "exchange_bytes": 0, "local_plans": 0, "local_ops": 0, "local_swaps": 0}

    # ------------------------------------------------------------------ helpers
    def _qubit_at(self, p: int) -> int:
        return self.pos.index(p)

    def _rank_bit(self, p: int) -> int:
        return (self.rank >> (p - self.n_local)) & 1

    def _local_qubit(self, p: int) -> int:
        """engine (API-style) qubit number of local index bit p"""
        return self.n_local - 1 - p

    def _run_local(self, recs: List[np.ndarray]):
        if not recs:
            return
        ops = np.concatenate(recs)
        plan = eng.Plan(self.n_local, ops, eng.PLAN_FUSE if self.fuse else 0)
        self.state.run(plan)
        self.state.sync()          # the plan's descriptors are freed with `plan`
        self.stats["local_plans"] += 1
        self.stats["local_ops"] += len(ops)

    def _translate(self, op: _Op) -> List[np.ndarray]:
        """One logical op -> local engine ops for THIS rank under the current layout."""
        ctrl_local, cval = [], 0
        for q, v in op.ctrl:
            p = self.pos[q]
            if p >= self.n_local:
                if self._rank_bit(p) != v:
                    return []                      # control not satisfied on this rank
            else:
                lq = self._local_qubit(p)
                ctrl_local.append(lq)
                cval |= v << lq
        if op.kind == eng.OP_DIAG:
            p = self.pos[op.t]
            if p >= self.n_local:
                f = op.m[3] if self._rank_bit(p) else op.m[0]
                if f == 1:
                    return []
                free = next(lq for lq in range(self.n_local) if lq not in ctrl_local)
                return [eng.op_record(eng.OP_DIAG, free, [f, 0, 0, f], ctrl_local, ctrl_value=cval)]
            return [eng.op_record(eng.OP_DIAG, self._local_qubit(p), op.m, ctrl_local, ctrl_value=cval)]
        if op.kind == eng.OP_SWAP:
            return [eng.op_record(eng.OP_SWAP, self._local_qubit(self.pos[op.t]), controls=ctrl_local,
                                  target2=self._local_qubit(self.pos[op.t2]), ctrl_value=cval)]
        return [eng.op_record(op.kind, self._local_qubit(self.pos[op.t]), op.m, ctrl_local, ctrl_value=cval)]

    # ------------------------------------------------------------------ exchanges
    def _local_swap_to_top(self, p: int):
        top = self.n_local - 1
        if p == top:
            return
        self._run_local([eng.op_record(eng.OP_SWAP, self._local_qubit(p), target2=self._local_qubit(top))])
        qa, qb = self._qubit_at(p), self._qubit_at(top)
        self.pos[qa], self.pos[qb] = top, p
        self.stats["local_swaps"] += 1

    def _exchange(self, q_global: int, q_local: int):
        """Swap logical qubits q_global (on a rank bit) and q_local (on a local bit)."""
        pg = self.pos[q_global]
        assert pg >= self.n_local > self.pos[q_local]
        self._local_swap_to_top(self.pos[q_local])
        b = self._rank_bit(pg)
        peer = self.rank ^ (1 << (pg - self.n_local))
        halves = self.buf.view(2, -1)
        out_half = halves[1 - b]                 # the half whose top local bit differs from my rank bit
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
        self.pos[q_global], self.pos[q_local] = self.n_local - 1, pg
        self.stats["exchanges"] += 1
        self.stats["exchange_bytes"] += out_half.numel() * 8

    def _next_use(self, q: int, ops: Sequence[_Op]) -> int:
        for i, op in enumerate(ops):
            if q in op.nd:
                return i
        return 1 << 30

    # ------------------------------------------------------------------ gates
    def apply_ops(self, records: np.ndarray):
        """Apply primitive ops (struct aqs_op records over all n qubits, API numbering)."""
        remaining = [_Op(r) for r in np.ascontiguousarray(records, dtype=eng.OP_DTYPE)]
        while remaining:
            batch, rest = [], []
            blocked_nd, blocked_d = set(), set()
            for op in remaining:
                conflict = (op.nd & (blocked_nd | blocked_d)) or (op.dg & blocked_nd)
                local = all(self.pos[q] < self.n_local for q in op.nd)
                if not conflict and local:
                    batch.append(op)
                else:
                    blocked_nd |= op.nd
                    blocked_d |= op.dg
                    rest.append(op)
            recs: List[np.ndarray] = []
            for op in batch:
                recs += self._translate(op)
            self._run_local(recs)
            remaining = rest
            if not remaining:
                break
            # bring in the global qubits needed soonest; evict the local qubits needed latest
            wanted: List[int] = []
            for op in remaining:
                for q in sorted(op.nd):
                    if self.pos[q] >= self.n_local and q not in wanted:
                        wanted.append(q)
                if len(wanted) >= self.g:
                    break
            protect = set()
            for op in remaining[:1]:
                protect |= op.nd
            for qg in wanted[: self.g]:
                locals_ = [q for q in range(self.n) if self.pos[q] < self.n_local and q not in protect and q not in wanted]
                victim = max(locals_, key=lambda q: (self._next_use(q, remaining), self.pos[q]))
                self._exchange(qg, victim)

    def run_circuit(self, qc) -> None:
        """simulate() for an afquantumsim_b200.aqs.QCircuit on more qubits than one GPU holds."""
        self.apply_ops(qc.ops())

    # ------------------------------------------------------------------ layout
    def canonicalize(self):
        """Bring every logical qubit q back to index bit n-1-q."""
        home = lambda q: self.n - 1 - q
        # 1. global positions
        for pg in range(self.n - 1, self.n_local - 1, -1):
            q_home = self.n - 1 - pg
            if self.pos[q_home] == pg:
                continue
            if self.pos[q_home] >= self.n_local:      # sits on another rank bit: bring it local first
                locals_ = [q for q in range(self.n) if self.pos[q] < self.n_local]
                self._exchange(q_home, locals_[0])
            self._exchange(self._qubit_at(pg), q_home)
        # 2. local permutation
        for p in range(self.n_local):
            q_home = self.n - 1 - p
            if self.pos[q_home] != p:
                other = self._qubit_at(p)
                self._run_local([eng.op_record(eng.OP_SWAP, self._local_qubit(self.pos[q_home]),
                                               target2=self._local_qubit(p))])
                self.pos[other], self.pos[q_home] = self.pos[q_home], p
                self.stats["local_swaps"] += 1
        assert all(self.pos[q] == home(q) for q in range(self.n))

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
        self.buf.zero_()
        if (index >> self.n_local) == self.rank:
            self.buf[index & ((1 << self.n_local) - 1)] = 1.0
        self.pos = [self.n - 1 - q for q in range(self.n)]

"""Test infrastructure: a numpy interpreter of the tile kernel's launch descriptors.

`Plan.export_pass` returns the exact bytes a fused pass hands to `k_tile2`
(afquantumsim_b200/csrc/tile_kernel.cuh: PassParams / TileSeg / TileOp).  This module
re-executes them on the CPU with the kernel's semantics — layouts (register bits /
thread bits), shared-memory swizzles, predicates, coefficient sets, in-place shears —
so the planner can be checked against the oracle without a GPU, and so that a GPU
mismatch can be pinned on the kernel rather than on the plan.  It is never imported by
the package."""
import ctypes

import numpy as np

K_LANE, K_REG, K_MAX_THREAD_BITS, K_MAX_BITS = 5, 5, 8, 38
SHR, SHI, GEN, PERM_R, PERM_I, PHASE, SCALE_R, SCALE_I, PHASE_N, LADDER, LADDER_CONT = range(11)
TF_MUX, TF_REGMUX, TF_PRED, TF_PY, TF_IMAG_A, TF_IMAG_B, TF_CY = 1, 2, 4, 8, 16, 32, 64


class BitList(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("pos", ctypes.c_uint8 * K_MAX_BITS)]


class Head(ctypes.Structure):
    _fields_ = [("tile_bits", ctypes.c_uint32), ("n_segs", ctypes.c_uint32), ("n_ops", ctypes.c_uint32),
                ("has_scale", ctypes.c_uint32), ("scale", ctypes.c_float * 2), ("n_tiles", ctypes.c_uint64),
                ("tile", BitList),
                ("ld_toff", ctypes.c_uint64 * K_MAX_THREAD_BITS), ("ld_roff", ctypes.c_uint64 * K_REG),
                ("st_toff", ctypes.c_uint64 * K_MAX_THREAD_BITS), ("st_roff", ctypes.c_uint64 * K_REG)]


class TileSeg(ctypes.Structure):
    _fields_ = [("rd_tcol", ctypes.c_uint16 * K_MAX_THREAD_BITS), ("rd_rcol", ctypes.c_uint16 * K_REG),
                ("wr_tcol", ctypes.c_uint16 * K_MAX_THREAD_BITS), ("wr_rcol", ctypes.c_uint16 * K_REG),
                ("first_op", ctypes.c_uint16), ("n_ops", ctypes.c_uint16), ("resplit", ctypes.c_uint8),
                ("pad", ctypes.c_uint8 * 7)]


class TileOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint8), ("tk", ctypes.c_uint8), ("flags", ctypes.c_uint8), ("mj", ctypes.c_uint8),
                ("mask", ctypes.c_uint32), ("t_mask", ctypes.c_uint16), ("t_val", ctypes.c_uint16), ("code", ctypes.c_uint32),
                ("b_mask", ctypes.c_uint32), ("b_val", ctypes.c_uint32), ("sx", ctypes.c_float * 2),
                ("a", ctypes.c_float * 8), ("b", ctypes.c_float * 8), ("qy", ctypes.c_float * 2), ("pad", ctypes.c_uint32 * 2)]


assert ctypes.sizeof(TileSeg) == 64 and ctypes.sizeof(TileOp) == 112


def parse(raw: bytes):
    head = Head.from_buffer_copy(raw)
    off = ctypes.sizeof(Head)
    segs = []
    for _ in range(head.n_segs):
        segs.append(TileSeg.from_buffer_copy(raw, off))
        off += ctypes.sizeof(TileSeg)
    ops = []
    for _ in range(head.n_ops):
        ops.append(TileOp.from_buffer_copy(raw, off))
        off += ctypes.sizeof(TileOp)
    assert off == len(raw)
    return head, segs, ops


def _xor_cols(bits, cols, nbits):
    out = np.zeros_like(bits)
    for j in range(nbits):
        out ^= np.where((bits >> j) & 1, np.uint32(cols[j]), np.uint32(0)).astype(np.uint32)
    return out


def _sum_cols(bits, cols, nbits):
    out = np.zeros(bits.shape, dtype=np.uint64)
    for j in range(nbits):
        out += np.where((bits >> j) & 1, np.uint64(cols[j]), np.uint64(0)).astype(np.uint64)
    return out


def deposit(j, positions):
    """insert a zero bit at every listed position, lowest first (common.cuh deposit_zeros)"""
    j = j.astype(np.uint64)
    for p in positions:
        lo = j & np.uint64((1 << p) - 1)
        j = ((j >> np.uint64(p)) << np.uint64(p + 1)) | lo
    return j


def butterfly(kind, x, y, c, py=False, sx=1.0, imag=False, cy=False, qy=0.0):
    """the kernel's in-place sequences, in float32 (tile_kernel.cuh shear<> / butterfly_direct<>)"""
    f = np.float32
    c = [f(v) for v in c]
    if py:
        if imag:
            assert kind == SHI
            x = x * (np.complex64(1j) * f(sx))
        else:
            x = x * f(sx)
        if cy:
            y = y * np.complex64(complex(c[3], f(qy)))       # TF_CY: explicit complex factor on y
        elif imag:
            y = y * (np.complex64(1j) * c[3])
        else:
            y = y * c[3]
    if kind == SHR:
        x = x + c[0] * y
        y = y + c[1] * x
        x = x + c[2] * y
    elif kind == SHI:
        j = np.complex64(1j)
        x = x + (j * c[0]) * y
        y = y + (j * c[1]) * x
        x = x + (j * c[2]) * y
    elif kind == PERM_R:
        x, y = c[0] * y, c[1] * x
    elif kind == PERM_I:
        x, y = (np.complex64(1j) * c[0]) * y, (np.complex64(1j) * c[1]) * x
    else:
        m = [np.complex64(complex(c[2 * i], c[2 * i + 1])) for i in range(4)]
        x, y = m[0] * x + m[1] * y, m[2] * x + m[3] * y
    return x.astype(np.complex64), y.astype(np.complex64)


def op_code(kind, tk, mj, flags):
    """tile_kernel.cuh tile_op_code"""
    if kind == 0 and flags & TF_CY:
        return ((32 + tk) << 3) | mj
    if kind == 1 and flags & TF_CY and not flags & (TF_IMAG_A | TF_IMAG_B):
        return ((37 + tk) << 3) | mj
    if kind <= 1:
        return ((kind * 5 + tk + (10 if flags & TF_PY else 0)) << 3) | mj
    if kind == LADDER:
        return (42 << 3) | mj
    if kind == LADDER_CONT:
        return 63 << 3
    if kind <= 4:
        return ((20 + kind - 2) << 3) | tk
    pat = mj - 1 if mj >= 8 else mj
    return ((23 + (kind - 5) * 2 + (pat >> 3)) << 3) | (pat & 7)


def factor_mask(op):
    """registers a factor op touches, from the compile-time pattern the kernel uses (must agree with op.mask)"""
    k = np.arange(32)
    if op.mj < 5:
        sel = (k >> op.mj) & 1 == 1
    elif op.mj == 5:
        sel = np.ones(32, dtype=bool)
    elif 8 <= op.mj <= 12:
        sel = (k >> (op.mj - 8)) & 1 == 0
    else:
        sel = ((op.mask >> k) & 1).astype(bool)
    assert np.array_equal(sel, ((op.mask >> k) & 1).astype(bool)), "mj pattern and mask disagree"
    return sel


def pair_uses_a(op, p):
    """shears: does register pair p take coefficient set a (else set b)?"""
    assert op.mj < 4, "shear ops resolve pair subsets at compile time"
    use = bool((p >> op.mj) & 1)
    if op.flags & TF_REGMUX:
        assert use == bool((op.mask >> p) & 1), "mj pattern and pair mask disagree"
    return use


def check_swizzle(seg, T, stats):
    """every re-split must be a bijection and keep each half-warp's 64-bit accesses on 16 distinct bank pairs"""
    TB = T - K_REG
    tid = np.arange(1 << TB, dtype=np.uint32)
    k = np.arange(32, dtype=np.uint32)
    for tcol, rcol, what in ((seg.wr_tcol, seg.wr_rcol, "store"), (seg.rd_tcol, seg.rd_rcol, "load")):
        slot = _xor_cols(tid, tcol, TB)[:, None] ^ _xor_cols(k, rcol, K_REG)[None, :]
        assert slot.max() < (1 << T)
        assert len(np.unique(slot)) == slot.size, f"{what} layout is not a bijection"
        banks = (slot & 15).reshape(-1, 16, 32)      # half-warps x lanes x registers
        worst = max(16 - len(np.unique(banks[h, :, r])) for h in range(0, banks.shape[0], max(1, banks.shape[0] // 8)) for r in (0, 31))
        stats["conflicts"] = max(stats.get("conflicts", 0), worst)


def run_pass(state: np.ndarray, raw: bytes, stats=None, cut=None):
    """apply one exported pass to `state` (complex64, length 2^n) in place.  cut = (positions, value): a sharded
    launch (aqs_plan_run_shard) — only the tiles whose number has the bits `positions` equal to those of `value`."""
    stats = stats if stats is not None else {}
    head, segs, ops = parse(raw)
    T = head.tile_bits
    TB = T - K_REG
    n_tiles = head.n_tiles
    tile_pos = [head.tile.pos[i] for i in range(head.tile.n)]
    assert head.tile.n == T and tile_pos[:5] == [0, 1, 2, 3, 4]
    blk = np.arange(n_tiles, dtype=np.uint64)
    if cut is not None:
        # kernel: tile number = blockIdx.x with a zero inserted at every pinned position (ascending), OR value
        fix_pos, fix_or = cut
        blk = deposit(np.arange(n_tiles >> len(fix_pos), dtype=np.uint64), fix_pos) | np.uint64(fix_or)
        n_tiles = blk.size
        stats.setdefault("touched", []).append(None)
    gbase = deposit(blk, tile_pos)                                       # (tiles,)
    tid = np.arange(1 << TB, dtype=np.uint32)
    k = np.arange(32, dtype=np.uint32)

    def global_index(toff, roff):
        g_t = (tid & 31).astype(np.uint64) + _sum_cols(tid >> 5, [toff[j] for j in range(5, TB)], TB - 5)
        g_r = _sum_cols(k, roff, K_REG)
        return gbase[:, None, None] + g_t[None, :, None] + g_r[None, None, :]

    gi = global_index(head.ld_toff, head.ld_roff)
    if cut is None:
        assert len(np.unique(gi)) == gi.size == state.size, "entry layout does not cover the state exactly once"
    else:
        assert len(np.unique(gi)) == gi.size, "entry layout touches an amplitude twice"
        stats["touched"][-1] = np.sort(gi.reshape(-1))
    a = state[gi]                                                        # (tiles, threads, 32)

    for si, sg in enumerate(segs):
        if sg.resplit:
            assert si > 0
            check_swizzle(sg, T, stats)
            w = _xor_cols(tid, sg.wr_tcol, TB)[:, None] ^ _xor_cols(k, sg.wr_rcol, K_REG)[None, :]
            r = _xor_cols(tid, sg.rd_tcol, TB)[:, None] ^ _xor_cols(k, sg.rd_rcol, K_REG)[None, :]
            sm = np.empty((n_tiles, 1 << T), dtype=np.complex64)
            sm[:, w.reshape(-1)] = a.reshape(n_tiles, -1)
            a = sm[:, r.reshape(-1)].reshape(a.shape)
            stats["resplits"] = stats.get("resplits", 0) + 1
        else:
            assert si == 0, "only the first segment may skip the re-split"
        seg_ops = ops[sg.first_op: sg.first_op + sg.n_ops]
        skip = 0
        for oi, op in enumerate(seg_ops):
            if skip:
                assert op.kind == LADDER_CONT
                skip -= 1
                continue
            mux = bool(op.flags & TF_MUX)
            blk_ok = (blk & np.uint64(op.b_mask)) == np.uint64(op.b_val)                 # (tiles,)
            thr_ok = (tid & np.uint32(op.t_mask)) == np.uint32(op.t_val)                   # (threads,)
            ok = blk_ok[:, None] & thr_ok[None, :]                                         # (tiles, threads)
            ca, cb = list(op.a), list(op.b)
            assert op.code == (op_code(op.kind, op.tk, op.mj, op.flags) | (op.flags << 16)), "dispatch code does not match (kind, tk, mj, flags)"
            assert bool(op.flags & TF_PRED) == bool(op.t_mask or op.b_mask), "TF_PRED must mirror the predicate fields"
            if op.kind == LADDER:
                # header, register-control record, n_cont records of four thread / block controls
                n_cont = int(np.array([op.sx[0]], dtype=np.float32).view(np.uint32)[0])
                skip = 1 + n_cont
                reg = seg_ops[oi + 1]
                w = [np.complex64(complex(reg.a[2 * r], reg.a[2 * r + 1])) for r in range(4)] + [np.complex64(complex(reg.sx[0], reg.sx[1]))]
                E = np.full((n_tiles, 1 << TB), np.complex64(complex(op.a[0], op.a[1])), dtype=np.complex64)
                for c in range(n_cont):
                    cr = seg_ops[oi + 2 + c]
                    assert cr.kind == LADDER_CONT
                    for q in range(4):
                        code = (cr.mask >> (8 * q)) & 0xff
                        wq = np.complex64(complex(cr.a[2 * q], cr.a[2 * q + 1]))
                        if code & 0x20:
                            bit = ((blk >> np.uint64(code & 0x1f)) & np.uint64(1)).astype(bool)[:, None]
                        else:
                            bit = ((tid >> np.uint32(code & 0x1f)) & np.uint32(1)).astype(bool)[None, :]
                        E = np.where(bit, (E * wq).astype(np.complex64), E)
                F = np.repeat(E[:, :, None], 32, axis=2)
                for r in range(5):
                    on = ((k >> r) & 1).astype(bool)[None, None, :]
                    F = np.where(on, (F * w[r]).astype(np.complex64), F)
                sel = factor_mask(op)
                if op.mj < 5:
                    assert w[op.mj] == 1, "no control on the hub's own register bit"
                m = ok[:, :, None] & sel[None, None, :]
                a = np.where(m, (a * F).astype(np.complex64), a)
                continue
            assert op.kind != LADDER_CONT, "stray ladder record"
            if op.kind >= PHASE:
                assert not mux
                if op.kind in (PHASE, PHASE_N):
                    # three shears on (re, im): rotation by theta with c = {-tan(theta/2), sin(theta)}
                    # (PHASE_N: of the negated amplitude)
                    t, sn = np.float32(ca[0]), np.float32(ca[1])
                    sgn = np.float32(-1.0 if op.kind == PHASE_N else 1.0)
                    xr, xi = sgn * a.real.astype(np.float32), sgn * a.imag.astype(np.float32)
                    xr = xr + t * xi
                    xi = xi + sn * xr
                    xr = xr + t * xi
                    rot = (xr + 1j * xi).astype(np.complex64)
                    sel = factor_mask(op)
                    m = ok[:, :, None] & sel[None, None, :]
                    a = np.where(m, rot, a)
                    continue
                elif op.kind == SCALE_R:
                    fac = np.complex64(np.float32(ca[0]))
                else:
                    fac = np.complex64(1j) * np.float32(ca[0])
                sel = factor_mask(op)
                m = ok[:, :, None] & sel[None, None, :]
                a = np.where(m, (a * fac).astype(np.complex64), a)
                continue
            tk = op.tk
            shear = op.kind <= SHI
            regmux = bool(op.flags & TF_REGMUX)
            assert not (mux and regmux)
            if not shear:
                assert not mux and not regmux
            for p in range(16):
                k0 = ((p >> tk) << (tk + 1)) | (p & ((1 << tk) - 1))
                k1 = k0 | (1 << tk)
                x, y = a[:, :, k0], a[:, :, k1]
                if not shear:
                    if not (op.mask >> p) & 1:
                        continue
                    xa, ya = butterfly(op.kind, x, y, ca)
                    a[:, :, k0] = np.where(ok, xa, x)
                    a[:, :, k1] = np.where(ok, ya, y)
                    continue
                # kernel: ka = (mux && !ok) ? b : a;  kb = regmux ? b : ka;  pair subset picks ka / kb;
                # threads with !ok and no mux skip the op
                py, cy = bool(op.flags & TF_PY), bool(op.flags & TF_CY)
                assert py or not cy, "TF_CY needs TF_PY"
                if not (mux or regmux):
                    cb, sxb, qyb = ca, op.sx[0], op.qy[0]       # kernel: kb = ka for a plain op
                else:
                    sxb, qyb = op.sx[1], op.qy[1]
                xa, ya = butterfly(op.kind, x, y, ca, py, op.sx[0], bool(op.flags & TF_IMAG_A), cy, op.qy[0])
                xb, yb = butterfly(op.kind, x, y, cb, py, sxb, bool(op.flags & (TF_IMAG_B if (mux or regmux) else TF_IMAG_A)), cy, qyb)
                if regmux:
                    xs, ys = (xa, ya) if pair_uses_a(op, p) else (xb, yb)
                    a[:, :, k0] = np.where(ok, xs, x)
                    a[:, :, k1] = np.where(ok, ys, y)
                elif mux:
                    a[:, :, k0] = np.where(ok, xa, xb)
                    a[:, :, k1] = np.where(ok, ya, yb)
                else:
                    assert op.mask == 0xffff
                    a[:, :, k0] = np.where(ok, xa, x)
                    a[:, :, k1] = np.where(ok, ya, y)

    go = global_index(head.st_toff, head.st_roff)
    if cut is None:
        assert len(np.unique(go)) == go.size == state.size, "exit layout does not cover the state exactly once"
    else:
        assert np.array_equal(np.sort(go.reshape(-1)), stats["touched"][-1]), "a sharded launch stores where it did not load"
    if head.has_scale:
        a = (a * np.complex64(complex(head.scale[0], head.scale[1]))).astype(np.complex64)
    state[go] = a
    return state


def run_plan(plan, state: np.ndarray, stats=None):
    info = plan.info()
    assert info["n_fused_passes"] > 0, "plan has no fused passes"
    state = np.array(state, dtype=np.complex64)
    for i in range(info["n_fused_passes"]):
        run_pass(state, plan.export_pass(i), stats)
    return state

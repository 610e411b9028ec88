// tile_kernel.cuh — the gate-fusion pass kernel ("tile kernel"), second generation.
//
// One launch applies a whole group of gates to the state with ONE read and ONE
// write of HBM (2*S bytes), instead of 2*S per gate.
//
// Tile.  A pass owns T index bits, the "tile bits": always the low 5 bits (so
// that every warp-level global access is a contiguous 256-byte run) plus T-5
// arbitrary higher bits.  One CTA of 2^(T-5) threads owns one tile = the 2^T
// amplitudes that differ only in the tile bits; each thread keeps 32 of them in
// registers for the whole pass.
//
// Layouts.  A tile-local index has T bits.  A LAYOUT picks 5 of them as
// "register bits" (they select one of the thread's 32 amplitudes) and maps the
// other T-5 to the bits of threadIdx.x.  A gate is only ever applied to a
// register bit: its butterflies are between two registers of the same thread,
// pure FP32 work with no data movement.  Controls cost nothing wherever their bit
// lives: on a register bit they select register pairs at plan time, on a thread
// bit they are a per-thread predicate, outside the tile a per-CTA predicate.
// When the next gates need other target bits the CTA changes layout through
// shared memory (a "segment" boundary).  The shared-memory slot of local index L
// is L ^ swz(L), an XOR swizzle chosen by the planner per re-split so that the
// 64-bit stores of the old layout and the 64-bit loads of the new one are both
// bank-conflict free.  ANY tile bit can become a register bit, including the low
// five: there are no shuffle butterflies.
//
// Arithmetic.  Amplitudes are (re, im) pairs in 64-bit registers and all math is
// Blackwell's packed FFMA2/FMUL2 (fma.rn.f32x2): a real coefficient is a scalar
// broadcast, an imaginary one uses the instruction's operand swap and per-half
// negate modifiers, so a complex multiply-add is two instructions.  2x2 matrices
// are classified at plan time: REAL (RotY, H, X and CX folded in), XLIKE / AXLIKE
// (checkerboard real/imaginary: RotX, Y, and those times X) cost 4 FFMA2 per
// amplitude pair, a general matrix 8, a phase 2 per touched amplitude.
//
// The op list is interpreted; descriptors travel as kernel parameters (constant
// bank), so decoding runs on the uniform datapath and the matrix entries never
// occupy vector registers longer than one op.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kLaneBits = 5;
constexpr int kRegBits = 5;
constexpr int kRegs = 1 << kRegBits;       // amplitudes per thread
constexpr int kPairs = kRegs / 2;
constexpr int kMinTileBits = kLaneBits + kRegBits;   // 10
constexpr int kMaxTileBits = 13;                     // 8192 amplitudes = 64 KiB of shared memory
constexpr int kMaxThreadBits = kMaxTileBits - kRegBits;
constexpr int kMaxSegs = 24;
constexpr int kOpsLarge = 300;            // ops per pass (shared memory: 64 bytes each)

enum TileKind : uint8_t {
    // butterflies on register pairs (mask = 16-bit pair mask)
    TK_SHR = 0,      // real shears       x += a*y; y += b*x; x += g*y           c = {a, b, g}
    TK_SHI = 1,      // imaginary shears  x += i*a*y; y += i*b*x; x += i*g*y     c = {a, b, g}
    TK_GEN = 2,      // general complex 2x2, direct: c = {m00.re, m00.im, m01.re, ... m11.im}
    TK_PERM_R = 3,   // anti-diagonal, real entries:  x' = c0*y, y' = c1*x  (X, CX, Swap: exact data movement)
    TK_PERM_I = 4,   // anti-diagonal, imaginary:     x' = i*c0*y, y' = i*c1*x  (Y)
    // factors on single registers (mask = 32-bit register mask)
    TK_PHASE = 5,    // unit-modulus factor e^{i theta}, |theta| <= pi/2:  c = {-tan(theta/2), sin(theta)}
    TK_SCALE_R = 6,  // real factor     c = {s}
    TK_SCALE_I = 7,  // imaginary factor i*s
    TK_PHASE_N = 8,  // -e^{i theta}: the same three shears with negated accumulators (pi/2 < |angle| <= pi)
    // A LADDER of controlled phases that share their target ("hub") qubit — QFT's CPhase(j, i) for all j < i — as
    // ONE op: every selected amplitude is multiplied by  self * PROD_{controls c whose bit is 1} w_c,  |w_c| = 1.
    // The header is followed by one TK_LADDER_CONT record with the factors of the controls that sit on register
    // bits (a[0..7], sx[0..1] = w_0 .. w_4, (1, 0) where there is none) and by sx[0] (as an integer) further
    // records with four thread / block controls each: byte q of `mask` = source of entry q (0..7: bit of
    // threadIdx.x; 0x20 | b: bit b of the tile number; 0x3f: empty), a[2q], a[2q + 1] = w.
    TK_LADDER = 9,
    TK_LADDER_CONT = 10
};
enum TileFlags : uint8_t {
    TF_MUX = 1,      // threads/CTAs whose predicate is false use coefficient set b instead of skipping
    TF_REGMUX = 2,   // TK_SHR / TK_SHI: pairs in `mask` use set a, all other pairs use set b
    TF_PRED = 4,     // t_mask or b_mask is non-zero (set by the planner; lets plain ops skip the predicate code)
    TF_PY = 8,       // shears: each set carries factors (sx, sy) applied first to x and y: reflections (CX
                     // folded into a rotation) and sign fixes ride inside the op.  sy = c[3]; sx = TileOp.sx[set]
    TF_IMAG_A = 16,  // TK_SHI with TF_PY: set a's factors are i*sx, i*sy (the X * RotX family)
    TF_IMAG_B = 32,  // same for set b
    TF_CY = 64       // (with TF_PY) the factor on y is complex, sy + i*qy per set: a diagonal gate that precedes the op on
                     // its target qubit (RotZ, Phase, ...) rides inside it as a phase on y; the x factor is as without TF_CY
};

struct alignas(16) TileOp {
    uint8_t kind;
    uint8_t tk;          // register bit of the target (butterfly kinds)
    uint8_t flags;
    uint8_t mj;          // how `mask` is structured, so that the kernel resolves it at compile time:
                         //   shears: 0..3 = pairs whose pair-index bit mj is set (set a; the others set b), 4 = generic
                         //   factors: 0..4 = registers whose bit mj is set, 8..12 = whose bit mj-8 is clear, 5 = all registers, 6 = generic
    uint32_t mask;       // butterfly kinds: bit p = register pair p takes part (16 bits); factor kinds: bit k = register k
    uint16_t t_mask;     // predicate on threadIdx.x: (tid & t_mask) == t_val
    uint16_t t_val;
    uint32_t code;       // tile_op_code(kind, tk, mj) | flags << 16: the one word the interpreter loop reads
    uint32_t b_mask;     // predicate on blockIdx.x (control bits outside the tile, in compact tile-number space)
    uint32_t b_val;
    float sx[2];         // TF_PY: factor on x for set a / set b
    float a[8];          // coefficient set used where the predicate holds
    float b[8];          // TF_MUX: coefficient set used where it does not
    float qy[2];         // TF_CY: imaginary part of the factor on y for set a / set b
    uint32_t pad[2];
};
static_assert(sizeof(TileOp) == 112, "TileOp layout");

// dispatch code = group * 8 + sub:
//   shears:  group = kind * 5 + tk (+ 10 with TF_PY)  (0..19),  sub = mj (0..4);  TK_SHR with TF_CY: group = 32 + tk;
//            TK_SHI with TF_CY and real factors on x: group = 37 + tk
//   direct:  group = 20 + kind - TK_GEN (20..22),                sub = tk
//   factors: group = 23 + (kind - TK_PHASE) * 2 + hi (23..30),   sub = pattern & 7, pattern = mj (0..6) or mj - 1 (7..11), hi = pattern >> 3
//   ladder:  group = 42, sub = mj (0..4: registers whose bit mj is set; 5: all registers);  its records: group 63
__host__ __device__ constexpr uint32_t tile_op_code(uint32_t kind, uint32_t tk, uint32_t mj, uint32_t flags) {
    return kind == 0 && (flags & 64u) ? ((32u + tk) << 3) | mj
         : kind == 1 && (flags & 64u) && !(flags & 48u) ? ((37u + tk) << 3) | mj
         : kind <= 1 ? ((kind * 5u + tk + ((flags & 8u) ? 10u : 0u)) << 3) | mj
         : kind == 9 ? (42u << 3) | mj
         : kind == 10 ? (63u << 3)
         : kind <= 4 ? ((20u + kind - 2u) << 3) | tk
                     : ((23u + (kind - 5u) * 2u + ((mj >= 8u ? mj - 1u : mj) >> 3)) << 3) | ((mj >= 8u ? mj - 1u : mj) & 7u);
}

// One layout plus the ops executed in it.  Slot index (in float2 units) of the amplitude held by
// thread `tid` in register k:   XOR_j (tid bit j ? tcol[j] : 0)  ^  XOR_i (k bit i ? rcol[i] : 0).
struct alignas(16) TileSeg {
    uint16_t rd_tcol[kMaxThreadBits];   // this segment's layout under the swizzle of the re-split that enters it
    uint16_t rd_rcol[kRegBits];
    uint16_t wr_tcol[kMaxThreadBits];   // the PREVIOUS segment's layout under the same swizzle
    uint16_t wr_rcol[kRegBits];
    uint16_t first_op;
    uint16_t n_ops;
    uint8_t resplit;                    // 0: same layout as the previous segment, no shared-memory trip
    uint8_t pad[7];
};
static_assert(sizeof(TileSeg) == 64, "TileSeg layout");

// What the kernel reads per op: the 64-byte compaction of a TileOp.  A pass's DevOps live in global
// memory (plan arena); every CTA copies them into shared memory once, because the interpreter's
// dependent loads must be cheap: an indexed LDC from the kernel-parameter bank costs >100 cycles,
// four of them per op in a chain made the whole kernel latency-bound.
struct alignas(16) DevOp {
    uint32_t word;       // tile_op_code(kind, tk, mj) | flags << 16
    uint32_t mask;       // shears (which resolve pair subsets at compile time) keep qy of set a here instead (float bits)
    uint32_t tpred;      // t_mask | t_val << 16
    float sx_a;          // TF_PY: factor on x, set a
    uint32_t b_mask;
    uint32_t b_val;
    float sx_b;          // TF_PY: factor on x, set b
    float qy_b;          // TF_CY: imaginary part of the factor on y, set b
    float a[4];          // TK_GEN: c[0..3]
    float b[4];          // TK_GEN: c[4..7]
};
static_assert(sizeof(DevOp) == 64, "DevOp layout");
constexpr uint32_t kDevOpEnd = 0x3fu;   // group value of the sentinel that follows the last op

struct alignas(16) PassParams {
    float2* state;         // loads
    float2* state_out;     // stores (== state, except for staged passes of sharded runs: flat.cu)
    const DevOp* ops;      // n_ops + 1 entries (sentinel), global memory
    uint32_t n_segs;
    uint32_t n_ops;
    float2 scale;          // global factor of the pass (phases folded out of RotZ-like ops)
    uint32_t has_scale;
    uint32_t tile_bits;
    // global offsets (in amplitudes) of the entry layout (first segment) and the exit layout (last
    // segment); both have threadIdx bits 0..4 = index bits 0..4
    uint64_t ld_toff[kMaxThreadBits];
    uint64_t ld_roff[kRegBits];
    uint64_t st_toff[kMaxThreadBits];
    uint64_t st_roff[kRegBits];
    BitList tile;          // global positions of the tile bits, ascending (tile.pos[0..4] = 0..4)
    // Sharded runs (aqs_plan_run_shard): this launch covers only the tiles whose number has the bits
    // fix_pos[0..fix_n) (ascending positions in the compact tile-number space) equal to those of fix_or;
    // blockIdx.x enumerates the remaining bits.  fix_n = 0: the launch covers every tile.
    uint32_t fix_n;
    uint32_t fix_or;
    uint8_t fix_pos[8];
    TileSeg segs[kMaxSegs];
};
static_assert(sizeof(PassParams) <= 4096, "kernel parameter space");

// ---- packed f32x2 helpers ---------------------------------------------------
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo(f2 x) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
    return a;
}
__device__ __forceinline__ float hi(f2 x) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
    return b;
}
__device__ __forceinline__ f2 sw(f2 x) { return pk(hi(x), lo(x)); }   // (re, im) -> (im, re): an operand modifier in SASS
__device__ __forceinline__ f2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
// i*s*v = s*(-v.im, v.re) accumulated onto acc
__device__ __forceinline__ f2 fma_i(float s, f2 v, f2 acc) { return fma2(pk(-s, s), sw(v), acc); }

// In-place butterflies.  A 2x2 is applied as   M = Shear(a, b, g) * diag(sx, sy):
//     [x *= sx; y *= sy;]   x += a*y;   y += b*x;   x += g*y
// Every step overwrites one of its own operands, so an amplitude never leaves its register and
// ptxas has nothing to rename across the interpreter's switch (a direct "x' = m00 x + m01 y"
// needs both old values for both outputs; ptxas then parks results in fresh registers and copies
// 32 pairs back at every join).  A rotation by |phi| <= pi/2 is a = g = -tan(phi/2), b = sin(phi)
// (all <= 1 in magnitude); reflections and sign fixes are separate TK_SCALE_* ops.  The same
// identity holds with imaginary shear coefficients for the RotX family.  3 FFMA2 per pair.
//
// Pair subsets are resolved at COMPILE time: an op carries two coefficient sets and the index J of
// a pair-index bit; pairs with that bit set use set a, the others set b (J is where a control that
// sits on a register bit lands in the pair index).  A plain op passes a == b, a controlled op
// passes b = 0 (x += 0*y leaves x untouched), a multiplexed op (CX folded into a rotation) both.
// Predicated FFMA2 is not an option: ptxas turns "@p FFMA2" into FFMA2 + 2 SEL.
struct ShearCoef {
    float a, b, g, sy, sx;
    float qy, qx;      // TK_SHI with a prescale: the factors are (sx + i qx), (sy + i qy)
};

template <int KIND, int PY>
__device__ __forceinline__ void shear(f2& x, f2& y, const ShearCoef& k) {
    if (KIND == TK_SHR) {
        if (PY == 1) {
            x = mul2(bc(k.sx), x);
            y = mul2(bc(k.sy), y);
        } else if (PY == 2) {
            // y *= sy + i*qy (a diagonal gate folded into the op), x *= sx
            x = mul2(bc(k.sx), x);
            const float yr = lo(y), yi = hi(y);
            const float ty = yr * k.sy - yi * k.qy;
            y = pk(ty, fmaf(yr, k.qy, yi * k.sy));
        }
        asm("{\n\t.reg .b64 ka, kb, kg;\n\t"
            "mov.b64 ka, {%2, %2};\n\tmov.b64 kb, {%3, %3};\n\tmov.b64 kg, {%4, %4};\n\t"
            "fma.rn.f32x2 %0, ka, %1, %0;\n\tfma.rn.f32x2 %1, kb, %0, %1;\n\tfma.rn.f32x2 %0, kg, %1, %0;\n\t}"
            : "+l"(x.v), "+l"(y.v) : "f"(k.a), "f"(k.b), "f"(k.g));
    } else {
        // x += i*a*y etc. on the halves with scalar FFMA (same FMA-pipe time as three FFMA2; the packed
        // form needs (-c, c) operand pairs that ptxas keeps rebuilding with MOVs)
        float xr = lo(x), xi = hi(x), yr = lo(y), yi = hi(y);
        if (PY == 2) {
            // real factor on x, complex factor on y (a diagonal gate folded into the op)
            xr *= k.sx; xi *= k.sx;
            const float ty = yr * k.sy - yi * k.qy;
            yi = fmaf(yr, k.qy, yi * k.sy);
            yr = ty;
        } else if (PY) {
            // complex factors (real or purely imaginary in practice; the set decides at run time)
            const float tx = xr * k.sx - xi * k.qx;
            xi = fmaf(xr, k.qx, xi * k.sx);
            xr = tx;
            const float ty = yr * k.sy - yi * k.qy;
            yi = fmaf(yr, k.qy, yi * k.sy);
            yr = ty;
        }
        xr = fmaf(-k.a, yi, xr); xi = fmaf(k.a, yr, xi);
        yr = fmaf(-k.b, xi, yr); yi = fmaf(k.b, xr, yi);
        xr = fmaf(-k.g, yi, xr); xi = fmaf(k.g, yr, xi);
        x = pk(xr, xi);
        y = pk(yr, yi);
    }
}

// Per-op prelude shared by every body: evaluates the predicate only when the planner flagged one.
// Returns false when this thread skips the op; `use_b` = take coefficient set b (TF_MUX).
__device__ __forceinline__ bool op_predicate(const DevOp& op, uint32_t flags, uint32_t tpred, uint32_t tid, uint32_t tile_no, bool& use_b) {
    use_b = false;
    if (flags & TF_PRED) {
        const uint2 blk = *reinterpret_cast<const uint2*>(&op.b_mask);
        const bool ok = ((tile_no & blk.x) == blk.y) && ((tid & (tpred & 0xffffu)) == (tpred >> 16));
        if (!ok) {
            if (!(flags & TF_MUX)) return false;
            use_b = true;
        }
    }
    return true;
}

// the rare kinds: direct general 2x2 and the exact anti-diagonal moves
template <int KIND>
__device__ __forceinline__ void butterfly_direct(f2& x, f2& y, const float (&c)[8]) {
    const f2 x0 = x, y0 = y;
    if (KIND == TK_PERM_R) {
        x = mul2(bc(c[0]), y0);
        y = mul2(bc(c[1]), x0);
    } else if (KIND == TK_PERM_I) {
        x = mul2(pk(-c[0], c[0]), sw(y0));
        y = mul2(pk(-c[1], c[1]), sw(x0));
    } else {
        f2 u = mul2(bc(c[0]), x0);
        u = fma_i(c[1], x0, u);
        u = fma2(bc(c[2]), y0, u);
        u = fma_i(c[3], y0, u);
        f2 w = mul2(bc(c[4]), x0);
        w = fma_i(c[5], x0, w);
        w = fma2(bc(c[6]), y0, w);
        w = fma_i(c[7], y0, w);
        x = u;
        y = w;
    }
}

// Pairs whose pair-index bit J is set use ka, the others kb (ops that need any other subset of the
// pairs are emitted as TK_GEN by the planner).
template <int KIND, int TK, int J, int PY>
__device__ __forceinline__ void apply_shear(f2 (&a)[kRegs], const ShearCoef& ka, const ShearCoef& kb) {
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        const int k0 = ((p >> TK) << (TK + 1)) | (p & ((1 << TK) - 1));
        const int k1 = k0 | (1 << TK);
        shear<KIND, PY>(a[k0], a[k1], (p >> J & 1) ? ka : kb);
    }
}
// factor kinds: MJ < 5: registers with bit MJ set; MJ == 5: all registers; 8..12: registers with bit MJ-8 clear;
// otherwise a generic 32-bit mask (uniform branches)
template <int KIND, int MJ>
__device__ __forceinline__ void apply_factor(f2 (&a)[kRegs], float fr, float fi, uint32_t mask) {
    // TK_PHASE multiplies by e^{i theta} as three shears on (re, im) — in place, no temporaries
    // (fr = -tan(theta/2), fi = sin(theta), |theta| <= pi/2; a sign or modulus is a separate TK_SCALE_R)
    const f2 re = bc(fr), ims = pk(-fr, fr);
#pragma unroll
    for (int k = 0; k < kRegs; ++k) {
        const bool on = (MJ < 5) ? ((k >> MJ & 1) != 0) : (MJ == 5 ? true : (MJ >= 8 ? ((k >> (MJ - 8) & 1) == 0) : ((mask >> k & 1u) != 0)));
        if (on) {
            if (KIND == TK_PHASE) {
                float xr = lo(a[k]), xi = hi(a[k]);
                xr = fmaf(fr, xi, xr);
                xi = fmaf(fi, xr, xi);
                xr = fmaf(fr, xi, xr);
                a[k] = pk(xr, xi);
            } else if (KIND == TK_PHASE_N) {
                // R(-v) = -R(v): start from the negated amplitude
                float xr = lo(a[k]), xi = hi(a[k]);
                xr = fmaf(-fr, xi, -xr);
                xi = fmaf(fi, xr, -xi);
                xr = fmaf(fr, xi, xr);
                a[k] = pk(xr, xi);
            } else if (KIND == TK_SCALE_R) {
                a[k] = mul2(re, a[k]);
            } else {
                a[k] = mul2(ims, sw(a[k]));
            }
        }
    }
}
template <int KIND, int TK>
__device__ __forceinline__ void apply_direct(f2 (&a)[kRegs], const float (&c)[8], uint32_t mask) {
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        const int k0 = ((p >> TK) << (TK + 1)) | (p & ((1 << TK) - 1));
        const int k1 = k0 | (1 << TK);
        if (mask >> p & 1u) butterfly_direct<KIND>(a[k0], a[k1], c);
    }
}
__host__ __device__ __forceinline__ constexpr int ctz_c(int i) { return (i & 1) ? 0 : (i & 2) ? 1 : (i & 4) ? 2 : (i & 8) ? 3 : 4; }
// Ladder of controlled phases: registers are visited in Gray-code order, so the running factor F changes by one
// multiplication with w_r or its conjugate per step (|w_r| = 1) and only F and the five w_r are live.
// MJ < 5: the hub is register bit MJ (only registers with that bit set are touched); MJ == 5: all 32.
template <int MJ>
__device__ __forceinline__ void apply_ladder(f2 (&a)[kRegs], float fr, float fi, const float (&wr)[kRegBits], const float (&wi)[kRegBits]) {
    constexpr int NB = (MJ < 5) ? kRegBits - 1 : kRegBits;
#pragma unroll
    for (int i = 0; i < (1 << NB); ++i) {
        const int g = i ^ (i >> 1);
        if (i) {
            const int cb = ctz_c(i);
            const int r = (MJ < 5 && cb >= MJ) ? cb + 1 : cb;
            const float s = (g >> cb & 1) ? wi[r] : -wi[r];     // entering the bit: * w_r, leaving it: * conj(w_r)
            const float t = fr * wr[r] - fi * s;
            fi = fmaf(fr, s, fi * wr[r]);
            fr = t;
        }
        const int k = (MJ < 5) ? ((((g >> MJ) << (MJ + 1)) | (g & ((1 << MJ) - 1))) | (1 << MJ)) : g;
        const float xr = lo(a[k]), xi = hi(a[k]);
        a[k] = pk(xr * fr - xi * fi, fmaf(xr, fi, xi * fr));
    }
}

// Header of an op.  Each body reloads it for the NEXT op as soon as it has consumed the current
// values, so the shared-memory latency hides behind the body's FP work and the loop carries no
// register rotation.
struct OpHead {
    uint4 h;        // {word, mask, tpred, sx_a}
    float4 a;
};
#ifndef AQS_HEAD_PREFETCH
#define AQS_HEAD_PREFETCH 0       // 0: every op loads its own header at the top of the loop; 1: the previous op prefetches it
                                  // (measured: 161.8 ms against 166.3 ms on brickwork-30 — the prefetch cost eight register copies per op)
#endif
__device__ __forceinline__ void load_head_now(OpHead& hd, const DevOp& op) {
    hd.h = *reinterpret_cast<const uint4*>(&op);
    hd.a = *reinterpret_cast<const float4*>(&op.a[0]);
}
__device__ __forceinline__ void load_head(OpHead& hd, const DevOp& op) {
    if (!AQS_HEAD_PREFETCH) return;
    hd.h = *reinterpret_cast<const uint4*>(&op);
    hd.a = *reinterpret_cast<const float4*>(&op.a[0]);
}

// Class preludes, one copy each in the interpreter loop: predicate, coefficient sets, header of the next op.
__device__ __forceinline__ ShearCoef make_coef(bool shi_py, float a, float b, float g, float sy, float sx, bool imag, bool cy, float qy) {
    ShearCoef k;
    k.a = a; k.b = b; k.g = g;
    const bool im = shi_py && imag;
    k.sx = im ? 0.f : sx; k.qx = im ? sx : 0.f;
    k.sy = (im && !cy) ? 0.f : sy; k.qy = cy ? qy : (im ? sy : 0.f);
    return k;
}
// RARE = false: the lean kernel has no body with an imaginary prescale, so that part of the decode is compiled
// out and the coefficient sets are taken as stored (the planner writes sx = 1, qy = 0 where a set has none; qy of
// set a shares its slot with the pair mask, which only the TF_CY bodies read as a float).  Three exits — plain
// op, register-multiplexed op, predicated op — each with the least work: the prelude is ~40 % of a shear op's
// instructions.
template <bool RARE>
__device__ __forceinline__ bool shear_prelude(const DevOp& op, OpHead& hd, uint32_t tid, uint32_t tile_no, ShearCoef& ka, ShearCoef& kb) {
    const uint32_t flags = hd.h.x >> 16;
    const bool py = (flags & TF_PY) != 0, cy = RARE ? (flags & TF_CY) != 0 : true;
    const uint32_t grp = (hd.h.x >> 3) & 0x3fu;
    const bool shi_py = RARE && py && grp >= 15u && grp < 20u;
    ka = make_coef(shi_py, hd.a.x, hd.a.y, hd.a.z, hd.a.w, __uint_as_float(hd.h.w), (flags & TF_IMAG_A) != 0, cy, __uint_as_float(hd.h.y));
#ifndef AQS_SETB_ALWAYS
#define AQS_SETB_ALWAYS 1         // lean kernel: plain ops read set b too (the planner stores a copy of set a there): one exit fewer,
                                  // no register copies for kb (measured 159.6 ms against 162.6 ms on brickwork-30)
#endif
    if (!(AQS_SETB_ALWAYS && !RARE) && !(flags & (TF_PRED | TF_REGMUX))) {
        kb = ka;
        load_head(hd, (&op)[1]);
        return true;
    }
    const float4 cb = *reinterpret_cast<const float4*>(&op.b[0]);
    const ShearCoef kset_b = make_coef(shi_py, cb.x, cb.y, cb.z, cb.w, (RARE && !py) ? 1.f : op.sx_b, (flags & TF_IMAG_B) != 0, cy, op.qy_b);
    if (!(flags & TF_PRED)) {
        kb = kset_b;                     // register-multiplexed: pairs pick their set at compile time
        load_head(hd, (&op)[1]);
        return true;
    }
    bool use_b;
    const bool run = op_predicate(op, flags, hd.h.z, tid, tile_no, use_b);
    if (use_b) ka = kset_b;
    kb = (flags & TF_REGMUX) ? kset_b : ka;
    load_head(hd, (&op)[1]);
    return run;
}

// dispatch on the low bits of `sub` with plain bit tests (nvcc lowers a switch to compare chains plus
// small jump tables whose entries are again constant-bank loads)
#define AQS_DISPATCH5(sub, F0, F1, F2, F3, F4)        \
    do {                                              \
        if ((sub) & 4u) { F4; }                       \
        else if ((sub) & 2u) { if ((sub) & 1u) { F3; } else { F2; } } \
        else { if ((sub) & 1u) { F1; } else { F0; } } \
    } while (0)

#ifndef AQS_TILE_MINB
#define AQS_TILE_MINB 4
#endif
#ifndef AQS_TILE_MINB13
#define AQS_TILE_MINB13 2
#endif
constexpr int tile_min_blocks(int T) { return T >= 13 ? AQS_TILE_MINB13 : AQS_TILE_MINB; }

// RARE = false leaves out the bodies that few circuits need — shears with a general complex prescale on both
// amplitudes, the direct 2x2, the imaginary anti-diagonal, ladders (tile_op_is_rare) — about 40 % of the SASS:
// the instruction footprint is what the interpreter stalls on ("no_inst"), and brickwork / Grover passes run 3 %
// faster on the lean instantiation.  The planner marks a pass RARE when one of its ops needs such a body.
__host__ __device__ constexpr bool tile_op_is_rare(uint32_t code) {
    const uint32_t grp = (code >> 3) & 0x3fu;
    return (grp >= 15u && grp < 20u) || grp == 20u || grp == 22u || grp == 42u;
}

template <int T, bool RARE>
__global__ void __launch_bounds__(1 << (T - kRegBits), tile_min_blocks(T)) k_tile2(const __grid_constant__ PassParams P) {
    constexpr int TB = T - kRegBits;   // thread bits
    extern __shared__ __align__(16) float2 sm[];                       // the tile, then the pass's DevOps
    const DevOp* sops = reinterpret_cast<const DevOp*>(sm + (1u << T));

    const uint32_t tid = threadIdx.x;
    // tile number: blockIdx.x, with the bits this rank holds fixed inserted (sharded runs)
    uint32_t tile_no = blockIdx.x;
    for (uint32_t i = 0; i < P.fix_n; ++i) {
        const uint32_t p = P.fix_pos[i];
        tile_no = ((tile_no >> p) << (p + 1)) | (tile_no & ((1u << p) - 1u));
    }
    tile_no |= P.fix_or;
    const uint64_t gbase = deposit_zeros((uint64_t)tile_no, P.tile);
    {
        // descriptors -> shared memory (16 bytes per thread per step; the list ends with a sentinel)
        const uint4* src = reinterpret_cast<const uint4*>(P.ops);
        uint4* dst = reinterpret_cast<uint4*>(sm + (1u << T));
        const uint32_t n16 = (P.n_ops + 1u) * (uint32_t)(sizeof(DevOp) / 16);
        for (uint32_t i = tid; i < n16; i += (1u << TB)) dst[i] = __ldg(src + i);
    }

    f2 a[kRegs];
    {
        uint64_t g = gbase | (uint64_t)(tid & 31u);
#pragma unroll
        for (int j = kLaneBits; j < TB; ++j)
            if (tid >> j & 1u) g += P.ld_toff[j];
        const float2* src = P.state + g;
#pragma unroll
        for (int k = 0; k < kRegs; ++k) {
            uint64_t off = 0;
#pragma unroll
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) off += P.ld_roff[i];
            const float2 v = src[off];
            a[k] = pk(v.x, v.y);
        }
    }

    const uint32_t n_segs = P.n_segs;
    for (uint32_t s = 0; s < n_segs; ++s) {
        const TileSeg& sg = P.segs[s];
        if (sg.resplit) {
            uint32_t wb = 0, rb = 0;
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                if (tid >> j & 1u) {
                    wb ^= sg.wr_tcol[j];
                    rb ^= sg.rd_tcol[j];
                }
            }
            // Registers are visited in Gray-code order so that each shared-memory address is the
            // previous one XOR one column: a single live address register instead of 32 (the
            // unrolled address set was the kernel's register-pressure peak).
            __syncthreads();
            {
                uint32_t addr = wb;
#pragma unroll
                for (int i = 0; i < kRegs; ++i) {
                    const int k = i ^ (i >> 1);
                    if (i) addr ^= sg.wr_rcol[ctz_c(i)];
                    sm[addr] = make_float2(lo(a[k]), hi(a[k]));
                }
            }
            __syncthreads();
            {
                uint32_t addr = rb;
#pragma unroll
                for (int i = 0; i < kRegs; ++i) {
                    const int k = i ^ (i >> 1);
                    if (i) addr ^= sg.rd_rcol[ctz_c(i)];
                    const float2 v = sm[addr];
                    a[k] = pk(v.x, v.y);
                }
            }
        }
        const uint32_t first = sg.first_op, end = first + sg.n_ops;
        if (first == end) continue;
        if (s == 0) __syncthreads();          // descriptors visible (later segments pass the re-split barriers)
        OpHead hd;
        load_head(hd, sops[first]);
        for (uint32_t o = first; o < end; ++o) {
            const DevOp& op = sops[o];          // every body leaves the header of op o + 1 in hd (the sentinel keeps it in bounds)
            if (!AQS_HEAD_PREFETCH) load_head_now(hd, op);
            const uint32_t word = hd.h.x, sub = word & 7u, grp = (word >> 3) & 0x3fu;
#define AQS_SH(K, TKV, PYV)                                                          \
    do {                                                                             \
        if (sub & 2u) { if (sub & 1u) apply_shear<K, TKV, 3, PYV>(a, ka, kb); else apply_shear<K, TKV, 2, PYV>(a, ka, kb); } \
        else { if (sub & 1u) apply_shear<K, TKV, 1, PYV>(a, ka, kb); else apply_shear<K, TKV, 0, PYV>(a, ka, kb); } \
    } while (0)
#define AQS_SH5(K, PYV, g)                                                           \
    do {                                                                             \
        if ((g) & 4u) AQS_SH(K, 4, PYV);                                             \
        else if ((g) & 2u) { if ((g) & 1u) AQS_SH(K, 3, PYV); else AQS_SH(K, 2, PYV); } \
        else { if ((g) & 1u) AQS_SH(K, 1, PYV); else AQS_SH(K, 0, PYV); }            \
    } while (0)
#define AQS_DI(K) AQS_DISPATCH5(sub, (apply_direct<K, 0>(a, c, mask)), (apply_direct<K, 1>(a, c, mask)), \
                                (apply_direct<K, 2>(a, c, mask)), (apply_direct<K, 3>(a, c, mask)), (apply_direct<K, 4>(a, c, mask)))
#define AQS_FA(K, HI)                                                                                          \
    do {                                                                                                       \
        if (!(HI)) {                                                                                           \
            if (sub & 4u) {                                                                                    \
                if (sub & 2u) { if (sub & 1u) apply_factor<K, 8>(a, fr, fi, mask); else apply_factor<K, 6>(a, fr, fi, mask); } \
                else { if (sub & 1u) apply_factor<K, 5>(a, fr, fi, mask); else apply_factor<K, 4>(a, fr, fi, mask); } \
            } else if (sub & 2u) { if (sub & 1u) apply_factor<K, 3>(a, fr, fi, mask); else apply_factor<K, 2>(a, fr, fi, mask); } \
            else { if (sub & 1u) apply_factor<K, 1>(a, fr, fi, mask); else apply_factor<K, 0>(a, fr, fi, mask); }    \
        } else {                                                                                               \
            if (sub & 2u) { if (sub & 1u) apply_factor<K, 12>(a, fr, fi, mask); else apply_factor<K, 11>(a, fr, fi, mask); } \
            else { if (sub & 1u) apply_factor<K, 10>(a, fr, fi, mask); else apply_factor<K, 9>(a, fr, fi, mask); }    \
        }                                                                                                      \
    } while (0)
            if (grp < 20u || (grp >= 32u && grp < 42u)) {
                // shears: grp = kind * 5 + tk (+ 10 with a prescale); 32 + tk: real shears, complex factor on y
                ShearCoef ka, kb;
                if (!shear_prelude<RARE>(op, hd, tid, tile_no, ka, kb)) continue;
                if (grp >= 32u) {
                    if (grp < 37u) AQS_SH5(TK_SHR, 2, grp - 32u);
                    else AQS_SH5(TK_SHI, 2, grp - 37u);
                } else if (grp < 10u) {
                    if (grp < 5u) AQS_SH5(TK_SHR, 0, grp);
                    else AQS_SH5(TK_SHI, 0, grp - 5u);
                } else {
                    if (grp < 15u) AQS_SH5(TK_SHR, 1, grp - 10u);
                    else if (RARE) AQS_SH5(TK_SHI, 1, grp - 15u);
                }
            } else if (RARE && grp == 42u) {
                // ladder of controlled phases: header, register-control record, n_cont thread/block-control records
                bool use_b;
                const bool run = op_predicate(op, word >> 16, hd.h.z, tid, tile_no, use_b);
                const uint32_t n_cont = hd.h.w;
                float fr = hd.a.x, fi = hd.a.y;
                const DevOp* rec = &op + 1;
                load_head(hd, (&op)[2u + n_cont]);
                o += 1u + n_cont;
                if (!run) continue;
                const float4 w01 = *reinterpret_cast<const float4*>(&rec->a[0]);
                const float4 w23 = *reinterpret_cast<const float4*>(&rec->b[0]);
                const float wr[kRegBits] = {w01.x, w01.z, w23.x, w23.z, rec->sx_a};
                const float wi[kRegBits] = {w01.y, w01.w, w23.y, w23.w, rec->sx_b};
                for (uint32_t c = 0; c < n_cont; ++c) {
                    const DevOp& cr = rec[1u + c];
                    const uint32_t codes = cr.mask;
                    const float4 ca = *reinterpret_cast<const float4*>(&cr.a[0]);
                    const float4 cb = *reinterpret_cast<const float4*>(&cr.b[0]);
                    const float cs[4] = {ca.x, ca.z, cb.x, cb.z}, sn[4] = {ca.y, ca.w, cb.y, cb.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t code = (codes >> (8 * q)) & 0xffu;
                        const uint32_t src = (code & 0x20u) ? tile_no : tid;
                        if ((src >> (code & 0x1fu)) & 1u) {
                            const float t = fr * cs[q] - fi * sn[q];
                            fi = fmaf(fr, sn[q], fi * cs[q]);
                            fr = t;
                        }
                    }
                }
                if (sub & 4u) { if (sub & 1u) apply_ladder<5>(a, fr, fi, wr, wi); else apply_ladder<4>(a, fr, fi, wr, wi); }
                else if (sub & 2u) { if (sub & 1u) apply_ladder<3>(a, fr, fi, wr, wi); else apply_ladder<2>(a, fr, fi, wr, wi); }
                else { if (sub & 1u) apply_ladder<1>(a, fr, fi, wr, wi); else apply_ladder<0>(a, fr, fi, wr, wi); }
            } else {
                bool use_b;
                const bool run = op_predicate(op, word >> 16, hd.h.z, tid, tile_no, use_b);
                const uint32_t mask = hd.h.y;
                const float4 c0 = hd.a;
                if (grp >= 23u) {
                    // factors: grp = 23 + (kind - TK_PHASE) * 2 + hi
                    load_head(hd, (&op)[1]);
                    if (!run) continue;
                    const float fr = c0.x, fi = c0.y;
                    const uint32_t g23 = grp - 23u;
                    if (g23 < 2u) AQS_FA(TK_PHASE, g23 & 1u);
                    else if (g23 < 4u) AQS_FA(TK_SCALE_R, g23 & 1u);
                    else if (g23 < 6u) AQS_FA(TK_SCALE_I, g23 & 1u);
                    else AQS_FA(TK_PHASE_N, g23 & 1u);
                } else {
                    const float4 c1 = *reinterpret_cast<const float4*>(&op.b[0]);
                    load_head(hd, (&op)[1]);
                    if (!run) continue;
                    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    if (grp == 21u) AQS_DI(TK_PERM_R);
                    else if (RARE) {
                        if (grp == 20u) AQS_DI(TK_GEN);
                        else AQS_DI(TK_PERM_I);
                    }
                }
            }
#undef AQS_SH
#undef AQS_SH5
#undef AQS_DI
#undef AQS_FA
        }
    }

    {
        uint64_t g = gbase | (uint64_t)(tid & 31u);
#pragma unroll
        for (int j = kLaneBits; j < TB; ++j)
            if (tid >> j & 1u) g += P.st_toff[j];
        float2* dst = P.state_out + g;
        const bool sc = P.has_scale != 0;
        const f2 re = bc(P.scale.x);
        const f2 im = pk(-P.scale.y, P.scale.y);
#pragma unroll
        for (int k = 0; k < kRegs; ++k) {
            uint64_t off = 0;
#pragma unroll
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) off += P.st_roff[i];
            const f2 v = sc ? fma2(im, sw(a[k]), mul2(re, a[k])) : a[k];
            dst[off] = make_float2(lo(v), hi(v));
        }
    }
}

}  // namespace aqs

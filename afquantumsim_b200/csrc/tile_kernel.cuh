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
constexpr int kOpsSmall = 36;
constexpr int kOpsLarge = 300;

enum TileKind : uint8_t {
    TK_SHR = 0,      // real shears:                      c = {a, b, g}
    TK_SHR_P = 1,    // real prescale then real shears:   c = {a, b, g, sx, sy}
    TK_SHI = 2,      // imaginary shears i*a, i*b, i*g:   c = {a, b, g}
    TK_SHI_P = 3,    // real prescale, imaginary shears:  c = {a, b, g, sx, sy}
    TK_SHI_Q = 4,    // imaginary prescale i*sx, i*sy, imaginary shears
    TK_GEN = 5,      // general complex 2x2, direct:      c = {m00.re, m00.im, m01.re, ... m11.im}
    TK_PHASE = 6,    // one complex factor on selected registers: c = {re, im}
    TK_PERM_R = 7,   // anti-diagonal, real entries:  x' = c0*y, y' = c1*x  (X, CX, Swap: exact data movement)
    TK_PERM_I = 8    // anti-diagonal, imaginary:     x' = i*c0*y, y' = i*c1*x  (Y)
};
enum TileFlags : uint8_t {
    TF_MUX = 1       // threads/CTAs whose predicate is false use coefficient set b instead of skipping
};

struct alignas(16) TileOp {
    uint8_t kind;
    uint8_t tk;          // register bit of the target (butterfly kinds)
    uint8_t flags;
    uint8_t pad0;
    uint32_t mask;       // butterfly kinds: bit p = register pair p takes part (16 bits); PHASE: bit k = register k
    uint16_t t_mask;     // predicate on threadIdx.x: (tid & t_mask) == t_val
    uint16_t t_val;
    uint32_t b_mask;     // predicate on blockIdx.x (control bits outside the tile, in compact tile-number space)
    uint32_t b_val;
    uint32_t pad1[3];
    float a[8];          // coefficient set used where the predicate holds
    float b[8];          // TF_MUX: coefficient set used where it does not
};
static_assert(sizeof(TileOp) == 96, "TileOp layout");

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

template <int CAP>
struct alignas(16) PassParams {
    float2* state;
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
    TileSeg segs[kMaxSegs];
    TileOp ops[CAP];
};
static_assert(sizeof(PassParams<kOpsLarge>) <= 32764, "kernel parameter space");

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
// (all <= 1 in magnitude); reflections and sign fixes go into (sx, sy).  The same identity holds
// with imaginary shear coefficients for the RotX family.  3 FFMA2 per pair instead of 4.
template <int KIND>
__device__ __forceinline__ void butterfly(f2& x, f2& y, const float (&c)[8]) {
    if (KIND == TK_SHR || KIND == TK_SHR_P) {
        if (KIND == TK_SHR_P) {
            x = mul2(bc(c[3]), x);
            y = mul2(bc(c[4]), y);
        }
        x = fma2(bc(c[0]), y, x);
        y = fma2(bc(c[1]), x, y);
        x = fma2(bc(c[2]), y, x);
    } else if (KIND == TK_SHI || KIND == TK_SHI_P || KIND == TK_SHI_Q) {
        if (KIND == TK_SHI_P) {
            x = mul2(bc(c[3]), x);
            y = mul2(bc(c[4]), y);
        } else if (KIND == TK_SHI_Q) {
            x = mul2(pk(-c[3], c[3]), sw(x));      // x *= i sx
            y = mul2(pk(-c[4], c[4]), sw(y));
        }
        x = fma_i(c[0], y, x);
        y = fma_i(c[1], x, y);
        x = fma_i(c[2], y, x);
    } else if (KIND == TK_PERM_R) {
        const f2 x0 = x;
        x = mul2(bc(c[0]), y);
        y = mul2(bc(c[1]), x0);
    } else if (KIND == TK_PERM_I) {
        const f2 x0 = x;
        x = mul2(pk(-c[0], c[0]), sw(y));
        y = mul2(pk(-c[1], c[1]), sw(x0));
    } else {
        // general complex 2x2, c = {m00.re, m00.im, m01.re, m01.im, m10.re, m10.im, m11.re, m11.im}
        const f2 x0 = x, y0 = y;
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

template <int KIND, int TK, bool MASKED>
__device__ __forceinline__ void apply_pairs(f2 (&a)[kRegs], const float (&c)[8], uint32_t pair_mask) {
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        const int k0 = ((p >> TK) << (TK + 1)) | (p & ((1 << TK) - 1));
        const int k1 = k0 | (1 << TK);
        if (!MASKED || (pair_mask >> p & 1u)) butterfly<KIND>(a[k0], a[k1], c);
    }
}

template <int KIND, bool MASKED>
__device__ __forceinline__ void apply_tk(f2 (&a)[kRegs], const float (&c)[8], uint32_t tk, uint32_t pair_mask) {
    switch (tk) {
        case 0: apply_pairs<KIND, 0, MASKED>(a, c, pair_mask); break;
        case 1: apply_pairs<KIND, 1, MASKED>(a, c, pair_mask); break;
        case 2: apply_pairs<KIND, 2, MASKED>(a, c, pair_mask); break;
        case 3: apply_pairs<KIND, 3, MASKED>(a, c, pair_mask); break;
        default: apply_pairs<KIND, 4, MASKED>(a, c, pair_mask); break;
    }
}

template <int KIND>
__device__ __forceinline__ void apply_kind(f2 (&a)[kRegs], const float (&c)[8], uint32_t tk, uint32_t pair_mask) {
    if (pair_mask == 0xffffu) apply_tk<KIND, false>(a, c, tk, pair_mask);
    else apply_tk<KIND, true>(a, c, tk, pair_mask);
}

constexpr int tile_min_blocks(int T) { return T >= 13 ? 2 : 4; }

template <int T, int CAP>
__global__ void __launch_bounds__(1 << (T - kRegBits), tile_min_blocks(T)) k_tile2(const __grid_constant__ PassParams<CAP> P) {
    constexpr int TB = T - kRegBits;   // thread bits
    extern __shared__ __align__(16) float2 sm[];

    const uint32_t tid = threadIdx.x;
    const uint64_t gbase = deposit_zeros((uint64_t)blockIdx.x, P.tile);

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
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRegs; ++k) {
                uint32_t off = 0;
#pragma unroll
                for (int i = 0; i < kRegBits; ++i)
                    if (k >> i & 1) off ^= sg.wr_rcol[i];
                sm[wb ^ off] = make_float2(lo(a[k]), hi(a[k]));
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRegs; ++k) {
                uint32_t off = 0;
#pragma unroll
                for (int i = 0; i < kRegBits; ++i)
                    if (k >> i & 1) off ^= sg.rd_rcol[i];
                const float2 v = sm[rb ^ off];
                a[k] = pk(v.x, v.y);
            }
        }
        const uint32_t first = sg.first_op, end = first + sg.n_ops;
        for (uint32_t o = first; o < end; ++o) {
            const TileOp& op = P.ops[o];
            const bool mux = (op.flags & TF_MUX) != 0;
            const bool blk_ok = (blockIdx.x & op.b_mask) == op.b_val;
            if (!blk_ok && !mux) continue;                               // CTA-uniform
            const bool ok = blk_ok && ((tid & op.t_mask) == op.t_val);
            float c[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = (mux && !ok) ? op.b[i] : op.a[i];
            if (ok || mux) {
                const uint32_t tk = op.tk, mask = op.mask;
                switch (op.kind) {
                    case TK_SHR: apply_kind<TK_SHR>(a, c, tk, mask); break;
                    case TK_SHR_P: apply_kind<TK_SHR_P>(a, c, tk, mask); break;
                    case TK_SHI: apply_kind<TK_SHI>(a, c, tk, mask); break;
                    case TK_SHI_P: apply_kind<TK_SHI_P>(a, c, tk, mask); break;
                    case TK_SHI_Q: apply_kind<TK_SHI_Q>(a, c, tk, mask); break;
                    case TK_GEN: apply_kind<TK_GEN>(a, c, tk, mask); break;
                    case TK_PERM_R: apply_kind<TK_PERM_R>(a, c, tk, mask); break;
                    case TK_PERM_I: apply_kind<TK_PERM_I>(a, c, tk, mask); break;
                    default: {
                        const f2 re = bc(c[0]);
                        const f2 im = pk(-c[1], c[1]);
                        if (mask == 0xffffffffu) {
#pragma unroll
                            for (int k = 0; k < kRegs; ++k) a[k] = fma2(im, sw(a[k]), mul2(re, a[k]));
                        } else {
#pragma unroll
                            for (int k = 0; k < kRegs; ++k)
                                if (mask >> k & 1u) a[k] = fma2(im, sw(a[k]), mul2(re, a[k]));
                        }
                    } break;
                }
            }
        }
    }

    {
        uint64_t g = gbase | (uint64_t)(tid & 31u);
#pragma unroll
        for (int j = kLaneBits; j < TB; ++j)
            if (tid >> j & 1u) g += P.st_toff[j];
        float2* dst = P.state + g;
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

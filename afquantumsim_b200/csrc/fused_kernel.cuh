// fused_kernel.cuh — the gate-fusion pass kernel ("tile kernel").
//
// One launch applies a whole group of gates to the state with ONE read and ONE
// write of HBM (2*S bytes), instead of 2*S per gate.
//
// Data layout.  A pass picks T index bits, the "tile bits": always the low 5
// bits (so that every warp-level access is a contiguous 256-byte run) plus up
// to T-5 arbitrary higher bits.  One CTA owns one tile = the 2^T amplitudes
// that differ only in the tile bits.  Inside the CTA a tile-local index has T
// bits and is split three ways:
//     lane bits      local bits 0..4        <-> the 32 lanes of a warp
//     register bits  4 local bits (R[0..3]) <-> the 16 amplitudes a thread holds
//     warp bits      the remaining T-9 bits <-> the warp number
// A gate whose target is a register bit is a butterfly between two registers of
// the same thread; a target on a lane bit is a butterfly between two lanes
// (__shfl_xor); diagonal gates and controls need no data movement at all, their
// bit may live anywhere (register, lane, warp or outside the tile = CTA-uniform).
// When the next gate targets a warp bit, the CTA re-shuffles through shared
// memory into a new register/warp split (a "segment" boundary).  Shared memory
// is indexed by the plain local index, so lanes always touch 32 consecutive
// 8-byte words: conflict-free for any split.
//
// The op list is interpreted at run time.  Everything that depends only on the
// op and the segment (which register pairs a control enables, which registers
// see the target bit set, the structure of the 2x2) is precomputed by the
// planner into bit masks, so the inner loops are the FP32 work plus one uniform
// bit test; the only divergence is a control that lives on a lane.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kLaneBits = 5;
constexpr int kRegBits = 4;
constexpr int kRegs = 1 << kRegBits;     // amplitudes per thread
constexpr int kMaxTileBits = 12;         // 4096 amplitudes = 32 KiB of shared memory
constexpr int kMinTileBits = kLaneBits + kRegBits;

// op modes (TileOp::mode)
enum TileMode : uint8_t {
    TM_REG_GEN = 0,    // general complex 2x2 on a register bit
    TM_REG_REAL = 1,   // all four entries real            (RotY, H)
    TM_REG_XLIKE = 2,  // m00,m11 real; m01,m10 imaginary  (RotX)
    TM_REG_PERM = 3,   // bit flip                         (X, CX, CCNot, ...)
    TM_LANE_GEN = 4,   // general 2x2 on a lane bit
    TM_LANE_PERM = 5,  // bit flip on a lane bit
    TM_PHASE = 6       // one factor on the selected amplitudes: every diagonal gate is lowered to these
                       // (Z, Phase, CZ, CPhase directly; diag(d0,d1) as d0 on the pass scale or as two phases)
};

struct alignas(16) TileOp {
    uint8_t mode;
    uint8_t tk;          // REG_*: register bit 0..3; LANE_*: lane bit 0..4
    uint16_t amp_mask;   // REG_*: bit p = pair p enabled (8 bits); others: bit k = register k enabled
    uint16_t pad2;
    uint16_t tl_mask;    // control predicate on the thread's local index (lane + warp bits): (base & tl_mask) == tl_val
    uint16_t tl_val;
    uint16_t pad0;
    uint32_t pad1;
    uint64_t g_mask;     // control predicate on the tile's global base index (bits outside the tile)
    uint64_t g_val;
    float2 m[4];         // 2x2 row-major; PHASE uses m[0]
};
static_assert(sizeof(TileOp) == 64, "TileOp layout");

struct TileSeg {
    uint8_t R[kRegBits];   // local positions of the register bits (each >= 5)
    uint8_t W[3];          // local positions of the warp bits (first T-9 entries used)
    uint8_t pad;
    uint32_t first_op;
    uint32_t n_ops;
};

struct TileArgs {
    float2* state;
    const TileSeg* segs;
    const TileOp* ops;
    uint32_t n_segs;
    uint32_t tile_bits;    // T
    float2 scale;          // global factor of the pass (product of the phases folded out of RotZ-like ops)
    uint32_t has_scale;
    BitList tile;          // global positions of the tile bits, ascending (tile.pos[0..4] = 0..4)
};

template <int TK, int MODE>
__device__ __forceinline__ void reg_pairs(float2 (&a)[kRegs], const float2 (&m)[4], uint32_t pair_mask) {
#pragma unroll
    for (int p = 0; p < kRegs / 2; ++p) {
        const int k0 = ((p >> TK) << (TK + 1)) | (p & ((1 << TK) - 1));
        const int k1 = k0 | (1 << TK);
        if (pair_mask >> p & 1u) {
            const float2 x = a[k0], y = a[k1];
            if (MODE == TM_REG_PERM) {
                a[k0] = y; a[k1] = x;
            } else if (MODE == TM_REG_REAL) {
                a[k0] = make_float2(fmaf(m[1].x, y.x, m[0].x * x.x), fmaf(m[1].x, y.y, m[0].x * x.y));
                a[k1] = make_float2(fmaf(m[3].x, y.x, m[2].x * x.x), fmaf(m[3].x, y.y, m[2].x * x.y));
            } else if (MODE == TM_REG_XLIKE) {
                // (r + i s)(yr + i yi) with r = 0: i*s*y = (-s*yi, s*yr)
                a[k0] = make_float2(fmaf(-m[1].y, y.y, m[0].x * x.x), fmaf(m[1].y, y.x, m[0].x * x.y));
                a[k1] = make_float2(fmaf(-m[2].y, x.y, m[3].x * y.x), fmaf(m[2].y, x.x, m[3].x * y.y));
            } else {
                a[k0] = cdot2(m[0], x, m[1], y);
                a[k1] = cdot2(m[2], x, m[3], y);
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ void reg_dispatch(float2 (&a)[kRegs], const float2 (&m)[4], uint32_t tk, uint32_t pair_mask) {
    switch (tk) {
        case 0: reg_pairs<0, MODE>(a, m, pair_mask); break;
        case 1: reg_pairs<1, MODE>(a, m, pair_mask); break;
        case 2: reg_pairs<2, MODE>(a, m, pair_mask); break;
        default: reg_pairs<3, MODE>(a, m, pair_mask); break;
    }
}

template <int WARPS_LOG2>
__global__ void __launch_bounds__(32 << WARPS_LOG2, 4) k_tile(const __grid_constant__ TileArgs P) {
    constexpr int T = kMinTileBits + WARPS_LOG2;
    __shared__ float2 sm[1 << T];

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t gbase = deposit_zeros((uint64_t)blockIdx.x, P.tile);

    float2 a[kRegs];
    uint32_t base_local = 0;       // this thread's local index with the register bits clear
    uint32_t roff[kRegBits];       // local-index offset of each register bit

    auto enter_segment = [&](const TileSeg& sg) {
        base_local = lane;
#pragma unroll
        for (int j = 0; j < WARPS_LOG2; ++j) base_local |= ((warp >> j) & 1u) << sg.W[j];
#pragma unroll
        for (int i = 0; i < kRegBits; ++i) roff[i] = 1u << sg.R[i];
    };
    auto local_of = [&](int k) {
        uint32_t x = base_local;
#pragma unroll
        for (int i = 0; i < kRegBits; ++i)
            if (k >> i & 1) x |= roff[i];
        return x;
    };
    auto spread_thread = [&]() {
        uint64_t g = gbase | (uint64_t)lane;   // tile.pos[0..4] == 0..4
        for (int j = kLaneBits; j < T; ++j)
            if (base_local >> j & 1u) g |= 1ull << P.tile.pos[j];
        return g;
    };

    TileSeg sg = P.segs[0];
    enter_segment(sg);
    {
        const uint64_t gt = spread_thread();
        uint64_t go[kRegBits];
#pragma unroll
        for (int i = 0; i < kRegBits; ++i) go[i] = 1ull << P.tile.pos[sg.R[i]];
#pragma unroll
        for (int k = 0; k < kRegs; ++k) {
            uint64_t g = gt;
#pragma unroll
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) g |= go[i];
            a[k] = P.state[g];
        }
    }

    for (uint32_t s = 0; s < P.n_segs; ++s) {
        if (s) {
            // re-split through shared memory (plain local index => conflict-free)
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRegs; ++k) sm[local_of(k)] = a[k];
            __syncthreads();
            sg = P.segs[s];
            enter_segment(sg);
#pragma unroll
            for (int k = 0; k < kRegs; ++k) a[k] = sm[local_of(k)];
        }
        const uint32_t end = sg.first_op + sg.n_ops;
        for (uint32_t o = sg.first_op; o < end; ++o) {
            const TileOp* op = P.ops + o;
            const ulonglong2 gm = *reinterpret_cast<const ulonglong2*>(&op->g_mask);
            if ((gbase & gm.x) != gm.y) continue;                       // CTA-uniform control outside the tile
            const uint4 hd = *reinterpret_cast<const uint4*>(op);       // mode,tk,amp_mask | -,tl_mask | tl_val,..
            const uint32_t mode = hd.x & 0xffu, tk = (hd.x >> 8) & 0xffu, amp_mask = hd.x >> 16;
            const uint32_t tl_mask = hd.y >> 16, tl_val = hd.z & 0xffffu;
            const bool thr_ok = ((base_local & tl_mask) == tl_val);
            float2 m[4];
            {
                const float4 m01 = *reinterpret_cast<const float4*>(&op->m[0]);
                const float4 m23 = *reinterpret_cast<const float4*>(&op->m[2]);
                m[0] = make_float2(m01.x, m01.y); m[1] = make_float2(m01.z, m01.w);
                m[2] = make_float2(m23.x, m23.y); m[3] = make_float2(m23.z, m23.w);
            }
            if (mode <= TM_REG_PERM) {
                if (thr_ok) {
                    switch (mode) {
                        case TM_REG_GEN: reg_dispatch<TM_REG_GEN>(a, m, tk, amp_mask); break;
                        case TM_REG_REAL: reg_dispatch<TM_REG_REAL>(a, m, tk, amp_mask); break;
                        case TM_REG_XLIKE: reg_dispatch<TM_REG_XLIKE>(a, m, tk, amp_mask); break;
                        default: reg_dispatch<TM_REG_PERM>(a, m, tk, amp_mask); break;
                    }
                }
            } else if (mode == TM_PHASE) {
                if (thr_ok) {
#pragma unroll
                    for (int k = 0; k < kRegs; ++k)
                        if (amp_mask >> k & 1u) a[k] = cmul(m[0], a[k]);
                }
            } else {
                // lane-bit target: every lane takes part in the shuffles; thr_ok only gates the update
                const uint32_t xm = 1u << tk;
                const bool hi = (lane & xm) != 0;
                const uint32_t act = thr_ok ? amp_mask : 0u;
                if (mode == TM_LANE_GEN) {
                    const float2 m_own = hi ? m[3] : m[0];
                    const float2 m_oth = hi ? m[2] : m[1];
#pragma unroll
                    for (int k = 0; k < kRegs; ++k) {
                        float2 other;
                        other.x = __shfl_xor_sync(0xffffffffu, a[k].x, xm);
                        other.y = __shfl_xor_sync(0xffffffffu, a[k].y, xm);
                        if (act >> k & 1u) a[k] = cdot2(m_own, a[k], m_oth, other);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kRegs; ++k) {
                        float2 other;
                        other.x = __shfl_xor_sync(0xffffffffu, a[k].x, xm);
                        other.y = __shfl_xor_sync(0xffffffffu, a[k].y, xm);
                        if (act >> k & 1u) a[k] = other;
                    }
                }
            }
        }
    }

    {
        const uint64_t gt = spread_thread();
        uint64_t go[kRegBits];
#pragma unroll
        for (int i = 0; i < kRegBits; ++i) go[i] = 1ull << P.tile.pos[sg.R[i]];
        const bool sc = P.has_scale != 0;
        const float2 f = P.scale;
#pragma unroll
        for (int k = 0; k < kRegs; ++k) {
            uint64_t g = gt;
#pragma unroll
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) g |= go[i];
            P.state[g] = sc ? cmul(f, a[k]) : a[k];
        }
    }
}

}  // namespace aqs

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
// The op list is interpreted at run time (uniform control flow; the only
// divergence is a control bit that lives on a lane).
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kLaneBits = 5;
constexpr int kRegBits = 4;
constexpr int kRegs = 1 << kRegBits;     // amplitudes per thread
constexpr int kMaxTileBits = 12;         // 4096 amplitudes = 32 KiB of shared memory
constexpr int kMinTileBits = kLaneBits + kRegBits;

enum TileMode : uint8_t { TM_REG_U2 = 0, TM_REG_PERM = 1, TM_LANE_U2 = 2, TM_LANE_PERM = 3, TM_DIAG = 4 };
enum DiagTarget : uint8_t { DT_THREAD = 4, DT_CTA = 5 };   // 0..3: register bit

struct alignas(16) TileOp {
    uint8_t mode;
    uint8_t tk;         // REG_*: register bit 0..3; LANE_*: lane bit 0..4; DIAG: 0..3 | DT_THREAD | DT_CTA
    uint8_t rk_mask;    // control predicate in register-index space: (k & rk_mask) == rk_val
    uint8_t rk_val;
    uint16_t tl_mask;   // control predicate on the thread's local index (lane + warp bits)
    uint16_t tl_val;
    uint16_t tl_tbit;   // DIAG / DT_THREAD: local-index mask of the target bit
    uint8_t d0_one;     // DIAG: m[0] == 1
    uint8_t pad;
    uint32_t pad2;
    uint64_t g_mask;    // control predicate on the tile's global base index (bits outside the tile)
    uint64_t g_val;
    uint64_t g_tbit;    // DIAG / DT_CTA: global mask of the target bit
    float2 m[4];
};
static_assert(sizeof(TileOp) == 80, "TileOp layout");

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
    BitList tile;          // global positions of the tile bits, ascending (tile.pos[0..4] = 0..4)
};

template <int TK, bool PERM>
__device__ __forceinline__ void reg_pairs(float2 (&a)[kRegs], const TileOp& op, bool thr_ok) {
    const float2 m00 = op.m[0], m01 = op.m[1], m10 = op.m[2], m11 = op.m[3];
#pragma unroll
    for (int p = 0; p < kRegs / 2; ++p) {
        const int k0 = ((p >> TK) << (TK + 1)) | (p & ((1 << TK) - 1));
        const int k1 = k0 | (1 << TK);
        if (thr_ok && ((k0 & op.rk_mask) == op.rk_val)) {
            const float2 x = a[k0], y = a[k1];
            if (PERM) {
                a[k0] = y; a[k1] = x;
            } else {
                a[k0] = cdot2(m00, x, m01, y);
                a[k1] = cdot2(m10, x, m11, y);
            }
        }
    }
}

template <bool PERM>
__device__ __forceinline__ void reg_dispatch(float2 (&a)[kRegs], const TileOp& op, bool thr_ok) {
    switch (op.tk) {
        case 0: reg_pairs<0, PERM>(a, op, thr_ok); break;
        case 1: reg_pairs<1, PERM>(a, op, thr_ok); break;
        case 2: reg_pairs<2, PERM>(a, op, thr_ok); break;
        default: reg_pairs<3, PERM>(a, op, thr_ok); break;
    }
}

template <int WARPS_LOG2>
__global__ void __launch_bounds__(32 << WARPS_LOG2, (WARPS_LOG2 == 3 ? 3 : 4)) k_tile(const __grid_constant__ TileArgs P) {
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
    // global address pieces of the current split
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
            const TileOp& op = P.ops[o];
            if ((gbase & op.g_mask) != op.g_val) continue;              // CTA-uniform control outside the tile
            const bool thr_ok = ((base_local & op.tl_mask) == op.tl_val);
            switch (op.mode) {
                case TM_REG_U2: reg_dispatch<false>(a, op, thr_ok); break;
                case TM_REG_PERM: reg_dispatch<true>(a, op, thr_ok); break;
                case TM_LANE_U2: {
                    const uint32_t xm = 1u << op.tk;
                    const bool hi = (lane & xm) != 0;
                    const float2 mx = hi ? op.m[2] : op.m[0];
                    const float2 my = hi ? op.m[3] : op.m[1];
#pragma unroll
                    for (int k = 0; k < kRegs; ++k) {
                        float2 other;
                        other.x = __shfl_xor_sync(0xffffffffu, a[k].x, xm);
                        other.y = __shfl_xor_sync(0xffffffffu, a[k].y, xm);
                        if (thr_ok && ((k & op.rk_mask) == op.rk_val)) {
                            const float2 x = hi ? other : a[k];
                            const float2 y = hi ? a[k] : other;
                            a[k] = cdot2(mx, x, my, y);
                        }
                    }
                    break;
                }
                case TM_LANE_PERM: {
                    const uint32_t xm = 1u << op.tk;
#pragma unroll
                    for (int k = 0; k < kRegs; ++k) {
                        float2 other;
                        other.x = __shfl_xor_sync(0xffffffffu, a[k].x, xm);
                        other.y = __shfl_xor_sync(0xffffffffu, a[k].y, xm);
                        if (thr_ok && ((k & op.rk_mask) == op.rk_val)) a[k] = other;
                    }
                    break;
                }
                default: {   // TM_DIAG
                    const float2 d0 = op.m[0], d1 = op.m[3];
                    bool tbit = false;
                    if (op.tk == DT_THREAD) tbit = (base_local & op.tl_tbit) != 0;
                    else if (op.tk == DT_CTA) tbit = (gbase & op.g_tbit) != 0;
                    if (thr_ok) {
#pragma unroll
                        for (int k = 0; k < kRegs; ++k) {
                            if ((k & op.rk_mask) == op.rk_val) {
                                const bool bit = op.tk < kRegBits ? ((k >> op.tk) & 1) != 0 : tbit;
                                if (bit) a[k] = cmul(d1, a[k]);
                                else if (!op.d0_one) a[k] = cmul(d0, a[k]);
                            }
                        }
                    }
                    break;
                }
            }
        }
    }

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
            P.state[g] = a[k];
        }
    }
}

}  // namespace aqs

// dense_tc.cuh — opaque 2^k x 2^k complex matrices (k = 4, 5, 6) on the 5th-generation tensor cores.
//
// What Gate / ControlGate do with a compiled inner circuit's matrix (reference src/quantum.cpp:1760-1814,
// 1888-1950) is, for a state vector, a batch of small dense contractions: every GROUP of 2^k amplitudes (the
// settings of the k target bits for one setting of all other bits) is multiplied by the same matrix.  As a real
// GEMM:   [Re out; Im out] (128 x G)  =  [[Mr, -Mi], [Mi, Mr]] (128 x 128)  x  [Re in; Im in] (128 x G)
// for k = 6 (k = 4, 5 are padded with spectator bits: I (x) M) — M = 128, K = 128, N = groups: a tcgen05 shape.
//
// Precision.  tcgen05 has no fp32 kind and the parity bar is 1e-5 relative L2 in complex64, so every operand is
// split into two TF32 numbers, x = hi + lo (hi = rna(x), lo = rna(x - hi): 22 mantissa bits together), and the
// product is three MMAs with fp32 accumulation in tensor memory:  Wh Xh + Wh Xl + Wl Xh  (the dropped Wl Xl term is
// 2^-22 relative).  That is 3 x 2 x 128 x 128 flops per group of 64 amplitudes = 1.65 TFLOP for a 2^30 state, under
// the HBM time of the pass at tensor-core rates — the FP32 SIMT kernel (dense_kernels.cuh) needs 16.5 ms for it.
//
// One CTA (256 threads) per SM, persistent over tiles of 64 groups; the HBM loads of tile t + 1 are in flight (in registers)
// while the MMAs of tile t run:
//   * the matrix lives in TENSOR MEMORY for the whole kernel (operand A from TMEM: 128 lanes x 256 columns, hi and
//     lo), written once with tcgen05.st;
//   * per tile, every thread loads 16 amplitudes (lanes = consecutive groups: 256-byte runs), splits them and stores
//     the hi / lo planes into shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core
//     matrices: 16-byte chunk kc of row n at kc * 1024 + n * 16), one elected thread issues the 48 tcgen05.mma
//     (K = 8 each) and commits them to an mbarrier;
//   * the accumulator (128 lanes x 64 columns fp32) comes back with tcgen05.ld, is transposed through shared memory
//     (re and im rows sit in different warps) and stored as interleaved complex64, coalesced like the loads.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kTcGroups = 64;               // groups per tile = N of the MMA
constexpr int kTcThreads = 256;             // 8 warps: two per quarter of the 128 tensor-memory lanes
constexpr uint32_t kTcPlaneBytes = 32768;   // one operand plane (hi or lo) of a tile: 128 K x 64 N x 4 bytes
constexpr uint32_t kTcStagePitch = 65;      // epilogue staging: float2 [group][65] (padded: conflict-free both ways)
// Draining tile t - 1 while the MMAs of tile t run (double-buffered planes and accumulators) was measured SLOWER than
// draining each tile right after its own MMAs (6.75 against 5.89 ms for k = 6 at n = 30): with the next tile's loads already
// in flight the tensor core is not what the four warps wait for — their own split / transpose / store work is.
constexpr bool kTcPipeline = false;
constexpr uint32_t kTcStageBytes = 64 * kTcStagePitch * 8;                 // 33280
constexpr uint32_t kTcSmemBytes = 4 * kTcPlaneBytes + kTcStageBytes + 256;

struct DenseTcArgs {
    float2* a;
    const float2* m;          // 64 x 64 complex, row-major (device)
    uint64_t n_groups;        // a multiple of 64
    uint64_t ctrl_or;         // control bits at their required values
    uint64_t toff[6];         // index offset of matrix-index bit i (bit 0 = least significant)
    BitList fixed;            // target, spectator and control positions, ascending
};

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
// shared-memory matrix descriptor: K-major, no swizzle; core matrices of 8 rows x 16 bytes are contiguous (128 bytes),
// the next 8 rows follow at SBO = 128, the second 16-byte chunk along K sits at LBO = 1024 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)(1024u >> 4) << 16;
    d |= (uint64_t)(128u >> 4) << 32;
    d |= 1ull << 46;                        // descriptor version of sm_100
    return d;
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) k_dense_tc(const __grid_constant__ DenseTcArgs P) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    float* b_hi = reinterpret_cast<float*>(tc_smem);
    float* b_lo = reinterpret_cast<float*>(tc_smem + kTcPlaneBytes);
    float* stage = reinterpret_cast<float*>(tc_smem + 4 * kTcPlaneBytes);   // epilogue staging (the planes of parity 1 follow those of parity 0)
    uint64_t* bar = reinterpret_cast<uint64_t*>(tc_smem + 4 * kTcPlaneBytes + kTcStageBytes);          // two mbarriers, one per parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + 4 * kTcPlaneBytes + kTcStageBytes + 32);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = ((warp & 3u) * 32u) << 16;                    // a warp reaches the quarter (warp % 4) of the 128 lanes
    const uint32_t col_half = warp >> 2;                                     // the two warps of a quarter split the columns

    // ---- A = [[Mr, -Mi], [Mi, Mr]] into tensor memory: row m in lane m, K along the columns; hi at [0, 128), lo at [128, 256)
    {
        const uint32_t mrow = (warp & 3u) * 32u + (tid & 31u);
        const uint32_t part = mrow >> 6, r = mrow & 63u;
        for (uint32_t c0 = col_half * 64u; c0 < col_half * 64u + 64u; c0 += 16u) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (uint32_t j = 0; j < 16u; ++j) {
                const uint32_t kk = c0 + j, pc = kk >> 6, c = kk & 63u;
                const float2 e = P.m[r * 64u + c];
                const float w = (part == 0u) ? (pc == 0u ? e.x : -e.y) : (pc == 0u ? e.y : e.x);
                tf32_split(w, hi[j], lo[j]);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                         :: "r"(tmem + lane_base + c0), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]),
                            "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]), "r"(hi[12]), "r"(hi[13]), "r"(hi[14]), "r"(hi[15]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                         :: "r"(tmem + lane_base + 128u + c0), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]),
                            "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]), "r"(lo[12]), "r"(lo[13]), "r"(lo[14]), "r"(lo[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcGroups >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_tmem = tmem + 256u;
    const uint32_t b_hi_addr = (uint32_t)__cvta_generic_to_shared(b_hi), b_lo_addr = (uint32_t)__cvta_generic_to_shared(b_lo);
    const uint32_t gn = tid & 63u, h = tid >> 6;                              // this thread's group of the tile, and its quarter of the 64 elements
    const uint32_t n_tiles = (uint32_t)(P.n_groups / kTcGroups);
    uint32_t phase[2] = {0u, 0u};

    // this thread's 16 amplitudes of a tile: c = h * 16 + i
    uint64_t coff[16];
#pragma unroll
    for (uint32_t i = 0; i < 16u; ++i) {
        const uint32_t c = h * 16u + i;
        uint64_t off = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b)
            if (c >> b & 1u) off += P.toff[b];
        coff[i] = off;
    }
    auto tile_base = [&](uint32_t tile) { return P.a + (deposit_zeros((uint64_t)tile * kTcGroups + gn, P.fixed) | P.ctrl_or); };
    float2 v[16];                                                             // the NEXT tile's inputs: loaded while this tile's MMAs run
    if (blockIdx.x < n_tiles) {
        const float2* b0 = tile_base(blockIdx.x);
#pragma unroll
        for (uint32_t i = 0; i < 16u; ++i) v[i] = b0[coff[i]];
    }

    // Software pipeline over the tiles of this CTA, two stages deep: while the tensor core multiplies tile t (operand planes and
    // accumulator of parity t & 1), the threads fetch tile t + 1 from HBM into registers and drain tile t - 1 (tcgen05.ld,
    // transpose through the staging area, coalesced stores).
    auto epilogue = [&](uint32_t par, float2* obase) {
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar_addr + par * 8u), "r"(phase[par]) : "memory");
            phase[par] ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // D: row m = (warp % 4) * 32 + lane, this warp's 32 of the 64 columns (= groups of the tile)
        uint32_t d[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
            "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]),
              "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]),
              "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
            : "r"(d_tmem + par * 64u + col_half * 32u + lane_base) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        {
            const uint32_t mrow = (warp & 3u) * 32u + (tid & 31u);
            const uint32_t part = mrow >> 6, r = mrow & 63u;
#pragma unroll
            for (uint32_t j = 0; j < 32u; ++j) stage[((col_half * 32u + j) * kTcStagePitch + r) * 2u + part] = __uint_as_float(d[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (uint32_t i = 0; i < 16u; ++i)
            obase[coff[i]] = *reinterpret_cast<const float2*>(stage + (gn * kTcStagePitch + h * 16u + i) * 2u);
    };

    uint32_t it = 0;
    float2* prev_base = nullptr;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = kTcPipeline ? (it & 1u) : 0u;
        float2* base = tile_base(tile);
        float* bh = b_hi + par * (2u * kTcPlaneBytes / 4u);
        float* bl = b_lo + par * (2u * kTcPlaneBytes / 4u);
        // ---- B = [Re in; Im in]: this thread's 32 amplitudes, split, as 16-byte chunks of 4 consecutive K.  (The planes of this
        // parity were last read by the MMAs of tile t - 2, whose completion the epilogue of tile t - 2 has waited for.)
#pragma unroll
        for (uint32_t q = 0; q < 4u; ++q) {
            uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                tf32_split(v[q * 4u + j].x, rh[j], rl[j]);
                tf32_split(v[q * 4u + j].y, ih[j], il[j]);
            }
            const uint32_t kc = h * 4u + q;                                   // chunk of the real part; the imaginary part is 16 chunks on
            *reinterpret_cast<uint4*>(bh + (kc * 64u + gn) * 4u) = make_uint4(rh[0], rh[1], rh[2], rh[3]);
            *reinterpret_cast<uint4*>(bl + (kc * 64u + gn) * 4u) = make_uint4(rl[0], rl[1], rl[2], rl[3]);
            *reinterpret_cast<uint4*>(bh + ((16u + kc) * 64u + gn) * 4u) = make_uint4(ih[0], ih[1], ih[2], ih[3]);
            *reinterpret_cast<uint4*>(bl + ((16u + kc) * 64u + gn) * 4u) = make_uint4(il[0], il[1], il[2], il[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                      // (also: the previous epilogue's read-out of the staging area is over)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ha = b_hi_addr + par * 2u * kTcPlaneBytes, la = b_lo_addr + par * 2u * kTcPlaneBytes;
#pragma unroll 1
            for (uint32_t ks = 0; ks < 16u; ++ks) {                           // K = 8 per instruction: two 16-byte chunks
                const uint64_t dh = tc_smem_desc(ha + ks * 2048u), dl = tc_smem_desc(la + ks * 2048u);
                tc_mma_ts(d_tmem + par * 64u, tmem + ks * 8u, dh, idesc, ks ? 1u : 0u);       // Wh Xh
                tc_mma_ts(d_tmem + par * 64u, tmem + ks * 8u, dl, idesc, 1u);                 // Wh Xl
                tc_mma_ts(d_tmem + par * 64u, tmem + 128u + ks * 8u, dh, idesc, 1u);          // Wl Xh
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_addr + par * 8u) : "memory");
        }
        if (tile + gridDim.x < n_tiles) {                                     // next tile's inputs fly while the tensor core works
            const float2* nb = tile_base(tile + gridDim.x);
#pragma unroll
            for (uint32_t i = 0; i < 16u; ++i) v[i] = nb[coff[i]];
        }
        if (kTcPipeline) {
            if (it) epilogue(par ^ 1u, prev_base);                            // tile t - 1, while the MMAs of tile t run
            prev_base = base;
        } else {
            epilogue(0u, base);                                               // this tile, as soon as its MMAs are done
        }
    }
    if (kTcPipeline && it) {
        __syncthreads();
        epilogue((it - 1u) & 1u, prev_base);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

}  // namespace aqs

// measure_kernels.cuh — state preparation, probabilities, collapse and sampling.
//
// Replaces the ArrayFire pipelines of QSimulator's measurement methods
// (src/quantum.cpp:293-531): |a|^2 is computed on the fly and never
// materialised; the cumulative distribution exists only as one 64-bit sum per
// 4096-amplitude tile.  All probability sums follow the exact-sum contract of
// include/aqs_engine.h (integer sums of trunc(p*2^62)), so results do not
// depend on the reduction order.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kTileBits = 12;                 // sampling tile: 4096 amplitudes = 32 KiB
constexpr uint64_t kTileAmps = 1ull << kTileBits;

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- state preparation ------------------------------------------------------
struct ProductArgs {
    float2 q[AQS_MAX_QUBITS][2];
    int n;
};
// generate_statevector (src/quantum.cpp:261-275): a[r] = prod_k q[k][bit_k(r)],
// multiplied from qubit 0 (the MSB) down, each product rounded like the oracle.
__global__ void __launch_bounds__(256) k_set_product(float2* __restrict__ a, uint64_t N, const __grid_constant__ ProductArgs P) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (uint64_t)gridDim.x * blockDim.x) {
        float2 v = P.q[0][(r >> (P.n - 1)) & 1ull];
        for (int k = 1; k < P.n; ++k) v = cmul_rn(v, P.q[k][(r >> (P.n - 1 - k)) & 1ull]);
        a[r] = v;
    }
}
__global__ void k_set_one(float2* a, uint64_t idx) { a[idx] = make_float2(1.f, 0.f); }
// a viewed as a column-major 2^m x 2^m matrix: ones on the diagonal
__global__ void __launch_bounds__(256) k_set_diag_ones(float2* a, int m) {
    const uint64_t M = 1ull << m;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += (uint64_t)gridDim.x * blockDim.x)
        a[c * M + c] = make_float2(1.f, 0.f);
}

// ---- reductions ---------------------------------------------------------------
// sum of F_k over indices with (k & mask) == value  (mask == 0: the whole state)
__global__ void __launch_bounds__(256) k_prob_fixed(const float2* __restrict__ a, uint64_t N, uint64_t mask,
                                                    uint64_t value, unsigned long long* out) {
    unsigned long long acc = 0;
    const uint64_t nv = N >> 1;
    if (nv == 0) {  // single amplitude pair cannot happen (N >= 2); keep for safety
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (uint64_t k = 0; k < N; ++k) if ((k & mask) == value) acc += fix62(prob_rn(a[k]));
    }
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (uint64_t)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[v];
        const uint64_t k = v << 1;
        if ((k & mask) == value) acc += fix62(prob_rn(make_float2(x.x, x.y)));
        if (((k | 1ull) & mask) == value) acc += fix62(prob_rn(make_float2(x.z, x.w)));
    }
    acc = warp_sum_u64(acc);
    __shared__ unsigned long long sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long t = threadIdx.x < 8 ? sm[threadIdx.x] : 0ull;
        t = warp_sum_u64(t);
        if (threadIdx.x == 0 && t) atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256) k_norm2(const float2* __restrict__ a, uint64_t N, double* out) {
    double acc = 0.0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (uint64_t)gridDim.x * blockDim.x) {
        const float2 x = a[r];
        acc += (double)x.x * x.x + (double)x.y * x.y;
    }
    acc = warp_sum_f64(acc);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < 8 ? sm[threadIdx.x] : 0.0;
        t = warp_sum_f64(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256) k_scale(float2* a, uint64_t N, float f) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (uint64_t)gridDim.x * blockDim.x) {
        float2 x = a[r];
        a[r] = make_float2(x.x * f, x.y * f);
    }
}

__global__ void __launch_bounds__(256) k_probabilities(const float2* __restrict__ a, uint64_t count, float* __restrict__ out) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < count; r += (uint64_t)gridDim.x * blockDim.x)
        out[r] = prob_rn(a[r]);
}

// measure(): src/quantum.cpp:336-339
__global__ void __launch_bounds__(256) k_collapse(float2* a, uint64_t N, uint64_t bitmask, int outcome, float p) {
    const float s = __fsqrt_rn(p);
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (uint64_t)gridDim.x * blockDim.x) {
        const int bit = (r & bitmask) != 0;
        float2 x = a[r];
        if (bit == outcome) x = make_float2(__fdiv_rn(x.x, s), __fdiv_rn(x.y, s));
        else x = make_float2(0.f, 0.f);
        a[r] = x;
    }
}

// ---- sampling -------------------------------------------------------------------
// pass 1: one 64-bit sum per tile (the only full read of the state: 1*S bytes)
__global__ void __launch_bounds__(256) k_tile_sums(const float2* __restrict__ a, uint64_t tile_amps,
                                                   unsigned long long* __restrict__ sums) {
    const float4* base = reinterpret_cast<const float4*>(a + (uint64_t)blockIdx.x * tile_amps);
    unsigned long long acc = 0;
    for (uint32_t v = threadIdx.x; v < (uint32_t)(tile_amps >> 1); v += 256) {
        const float4 x = base[v];
        acc += fix62(prob_rn(make_float2(x.x, x.y))) + fix62(prob_rn(make_float2(x.z, x.w)));
    }
    acc = warp_sum_u64(acc);
    __shared__ unsigned long long sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long t = threadIdx.x < 8 ? sm[threadIdx.x] : 0ull;
        t = warp_sum_u64(t);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}

// pass 2: in-place inclusive scan of the tile sums (one block; <= 2^22 entries)
__global__ void __launch_bounds__(1024) k_scan_tiles(unsigned long long* s, uint64_t n) {
    const uint64_t chunk = (n + 1023) / 1024;
    const uint64_t b = (uint64_t)threadIdx.x * chunk;
    const uint64_t e = (b + chunk < n) ? b + chunk : n;
    unsigned long long acc = 0;
    for (uint64_t i = b; i < e; ++i) acc += s[i];
    // block exclusive scan of acc
    __shared__ unsigned long long wsum[32];
    unsigned long long incl = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long w = wsum[threadIdx.x];
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o) wi += t;
        }
        wsum[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    unsigned long long run = wsum[threadIdx.x >> 5] + (incl - acc);
    for (uint64_t i = b; i < e; ++i) { run += s[i]; s[i] = run; }
}

// pass 3: one block per draw.  Binary search over the inclusive tile sums, then
// resolve inside the owning tile (re-reads 32 KiB, L2-resident in practice).
// out = min{ k : S_k > U }, or 0 when U >= total (peek_measure_all :353-356).
__global__ void __launch_bounds__(256) k_sample(const float2* __restrict__ a, uint64_t tile_amps, uint64_t n_tiles,
                                                const unsigned long long* __restrict__ incl,
                                                const float* __restrict__ u, const unsigned long long* __restrict__ u_fixed,
                                                unsigned long long* __restrict__ out, uint32_t* __restrict__ hist) {
    const uint64_t d = blockIdx.x;
    const unsigned long long U = u_fixed ? u_fixed[d] : fix62(u[d]);
    // first tile t with incl[t] > U  (every thread runs the same uniform search)
    uint64_t lo = 0, hi = n_tiles;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (incl[mid] > U) hi = mid; else lo = mid + 1;
    }
    if (lo == n_tiles) {
        // no index qualifies: 0 for the public rule (peek_measure_all), "not in this shard" for the fixed-point entry
        if (threadIdx.x == 0) { if (out) out[d] = u_fixed ? ~0ull : 0ull; if (hist) atomicAdd(&hist[0], 1u); }
        return;
    }
    const uint64_t t = lo;
    const unsigned long long base = t ? incl[t - 1] : 0ull;
    const uint32_t per = (uint32_t)(tile_amps >> 8);      // amplitudes per thread (contiguous), 16 for a full tile
    const float2* tp = a + t * tile_amps;

    __shared__ unsigned long long wsum[8];
    __shared__ uint32_t best;
    if (threadIdx.x == 0) best = 0xffffffffu;

    unsigned long long f[16];
    unsigned long long acc = 0;
    if (per > 0) {
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) {
            f[i] = 0;
            if (i < per) { f[i] = fix62(prob_rn(tp[threadIdx.x * per + i])); acc += f[i]; }
        }
    } else {  // tiles smaller than 256 amplitudes: one amplitude per thread at most
        f[0] = (threadIdx.x < tile_amps) ? fix62(prob_rn(tp[threadIdx.x])) : 0ull;
        acc = f[0];
    }
    unsigned long long incl_t = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long x = __shfl_up_sync(0xffffffffu, incl_t, o);
        if ((threadIdx.x & 31) >= o) incl_t += x;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl_t;
    __syncthreads();
    unsigned long long run = base + (incl_t - acc);
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) run += wsum[w];
    const uint32_t cnt = per > 0 ? per : 1u;
    const uint32_t first = per > 0 ? threadIdx.x * per : threadIdx.x;
    uint32_t mine = 0xffffffffu;
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) {
        if (i < cnt) {
            run += f[i];
            if (run > U && mine == 0xffffffffu) mine = first + i;
        }
    }
    if (mine != 0xffffffffu) atomicMin(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long k = t * tile_amps + best;
        if (out) out[d] = k;
        if (hist) atomicAdd(&hist[k], 1u);
    }
}

}  // namespace aqs

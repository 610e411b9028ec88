// dense_kernels.cuh — an opaque 2^k x 2^k complex matrix on k arbitrary qubits (k <= 6), under controls.
//
// Replaces, for a state vector, what Gate / ControlGate do with a compiled inner circuit
// (reference src/quantum.cpp:1760-1814, 1888-1950): identity(2^n) + scatter of 4^k 2^(n-k) entries +
// matmul.  Here one thread owns one group of 2^k amplitudes (the 2^k settings of the target bits for one
// setting of all other bits): it loads the group into registers, then produces the outputs one by one,
//   out[r] = SUM_c M[r][c] * in[c],
// storing each as soon as it is complete (every input is already in registers, so the update is in
// place).  The matrix sits in shared memory as packed operand pairs {(re, re), (-im, im)}: one broadcast
// LDS.128 feeds the two FFMA2 of a complex multiply-add (the swap of the input's halves is an operand
// modifier).  FP32 SIMT on purpose — tcgen05 has no fp32 kind and the 1e-5 parity bar rules out one-pass
// TF32.  Work is 2^k complex MACs per amplitude: k = 5 is about at the HBM roofline's
// doorstep (3.7 ms of FMA pipe at n = 30 against 2.6 ms of HBM), k = 6 is FMA-bound.
// Coalescing: lanes enumerate the free index bits, lowest first — full 256-byte runs whenever the five
// low index bits are neither targets nor controls.
#pragma once
#include "common.cuh"

namespace aqs {

constexpr int kDenseMaxK = 6;
constexpr int kDenseThreads = 128;

struct DenseArgs {
    float2* a;
    const float4* m;          // device: 4^k entries, row-major, {re, re, -im, im}
    uint64_t n_groups;
    uint64_t ctrl_or;         // control bits at their required values
    uint64_t toff[kDenseMaxK]; // index offset of matrix-index bit i (bit 0 = least significant)
    BitList fixed;            // target and control positions, ascending
};

template <int K>
__global__ void __launch_bounds__(kDenseThreads) k_dense(const __grid_constant__ DenseArgs P) {
    constexpr int D = 1 << K;
    extern __shared__ __align__(16) float4 sm_m[];
    for (int i = threadIdx.x; i < D * D; i += kDenseThreads) sm_m[i] = P.m[i];
    __syncthreads();
    const uint64_t j = (uint64_t)blockIdx.x * kDenseThreads + threadIdx.x;
    if (j >= P.n_groups) return;
    float2* base = P.a + (deposit_zeros(j, P.fixed) | P.ctrl_or);
    unsigned long long in[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint64_t off = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (c >> i & 1) off += P.toff[i];
        const float2 v = base[off];
        asm("mov.b64 %0, {%1, %2};" : "=l"(in[c]) : "f"(v.x), "f"(v.y));
    }
#pragma unroll 1
    for (int r = 0; r < D; ++r) {
        unsigned long long acc0 = 0ull, acc1 = 0ull;    // two accumulators: independent FFMA2 chains
        const float4* row = sm_m + r * D;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float4 e = row[c];
            unsigned long long re2, im2, sw;
            asm("mov.b64 %0, {%1, %2};" : "=l"(re2) : "f"(e.x), "f"(e.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(im2) : "f"(e.z), "f"(e.w));
            {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(in[c]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(sw) : "f"(hi), "f"(lo));
            }
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc0) : "l"(re2), "l"(in[c]));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc1) : "l"(im2), "l"(sw));
        }
        float ar, ai, br, bi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(ar), "=f"(ai) : "l"(acc0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(br), "=f"(bi) : "l"(acc1));
        uint64_t off = 0;
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (r >> i & 1) off += P.toff[i];
        base[off] = make_float2(ar + br, ai + bi);
    }
}

}  // namespace aqs

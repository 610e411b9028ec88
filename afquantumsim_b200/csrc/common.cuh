// common.cuh — shared device/host helpers of the aqs engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "aqs_engine.h"

namespace aqs {

constexpr int kMaxBits = AQS_MAX_QUBITS + 2;

// Sorted (ascending) list of index-bit positions that a kernel holds fixed;
// the remaining ("free") bits are enumerated by the work-item index.
struct BitList {
    int n;
    uint8_t pos[kMaxBits];
};

// Spread the bits of j over the free positions: insert a zero at every listed
// position, lowest first.  This is how kernels enumerate exactly the amplitudes
// an op touches (no index arrays, cf. the CSR builders of src/quantum.cpp).
__host__ __device__ __forceinline__ uint64_t deposit_zeros(uint64_t j, const BitList& f) {
#pragma unroll 4
    for (int k = 0; k < f.n; ++k) {
        const int p = f.pos[k];
        const uint64_t lo = j & ((1ull << p) - 1ull);
        j = ((j >> p) << (p + 1)) | lo;
    }
    return j;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// m0*x + m1*y
__device__ __forceinline__ float2 cdot2(float2 m0, float2 x, float2 m1, float2 y) {
    float re = m0.x * x.x;
    re = fmaf(-m0.y, x.y, re);
    re = fmaf(m1.x, y.x, re);
    re = fmaf(-m1.y, y.y, re);
    float im = m0.x * x.y;
    im = fmaf(m0.y, x.x, im);
    im = fmaf(m1.x, y.y, im);
    im = fmaf(m1.y, y.x, im);
    return make_float2(re, im);
}
// complex multiply with every operation individually rounded (no FMA), to be
// bit-identical with the -ffp-contract=off CPU oracle where exactness is part
// of the contract (state preparation, |a|^2).
__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b) {
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// p = fl(fl(re*re) + fl(im*im))
__device__ __forceinline__ float prob_rn(float2 a) {
    return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y));
}
// F = trunc(p * 2^62): the exact-sum contract (include/aqs_engine.h)
__device__ __forceinline__ unsigned long long fix62(float p) {
    return __float2ull_rz(__fmul_rn(p, 4611686018427387904.0f));
}

}  // namespace aqs

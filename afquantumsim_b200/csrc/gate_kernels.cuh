// gate_kernels.cuh — per-gate ("fusion off") kernels: one launch applies one
// primitive op in place and touches only the amplitudes that op changes.
//
// They replace, for a state vector, the reference's per-gate
// `circuit = af::matmul(M_gate, circuit)` (src/quantum.cpp:287-289 and every
// QGate::operator()), without ever building M_gate.
//
// Memory behaviour (HBM-bound): every work item moves whole
// 128-bit vectors (two adjacent amplitudes) whenever index bit 0 is neither the
// target nor a control.  Work items are numbered so that consecutive lanes of a
// warp map to consecutive vectors inside each run between fixed bits: a target
// bit p >= 6 gives two fully coalesced 512-byte runs per warp-level load.
#pragma once
#include "common.cuh"

namespace aqs {

enum MatKind { MK_GENERAL = 0, MK_PERM = 1 };

struct PairArgs {
    float2* a;
    uint64_t n_items;   // work items (vectors of VEC amplitudes)
    uint64_t offA;      // bits OR-ed into the base index for element A (control values, swap bit A)
    uint64_t offB;      // ... for element B (control values | target bit, swap bit B)
    BitList fixed;      // positions not enumerated by the item index
    float2 m00, m01, m10, m11;
};

constexpr int kThreads = 256;

// (A, B) <- M (A, B) for every pair.  VEC = amplitudes per memory access (2 -> float4).
template <int VEC, int MK, int ITEMS>
__global__ void __launch_bounds__(kThreads) k_pair(const __grid_constant__ PairArgs P) {
    const uint64_t j0 = (uint64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    if (VEC == 2) {
        float4 va[ITEMS], vb[ITEMS];
        uint64_t ia[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            ia[u] = deposit_zeros(j, P.fixed);
            if (j < P.n_items) {
                va[u] = *reinterpret_cast<const float4*>(P.a + (ia[u] | P.offA));
                vb[u] = *reinterpret_cast<const float4*>(P.a + (ia[u] | P.offB));
            }
        }
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            if (j < P.n_items) {
                float4 ra, rb;
                if (MK == MK_PERM) {
                    ra = vb[u]; rb = va[u];
                } else {
                    const float2 a0 = make_float2(va[u].x, va[u].y), a1 = make_float2(va[u].z, va[u].w);
                    const float2 b0 = make_float2(vb[u].x, vb[u].y), b1 = make_float2(vb[u].z, vb[u].w);
                    const float2 x0 = cdot2(P.m00, a0, P.m01, b0), y0 = cdot2(P.m10, a0, P.m11, b0);
                    const float2 x1 = cdot2(P.m00, a1, P.m01, b1), y1 = cdot2(P.m10, a1, P.m11, b1);
                    ra = make_float4(x0.x, x0.y, x1.x, x1.y);
                    rb = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
                *reinterpret_cast<float4*>(P.a + (ia[u] | P.offA)) = ra;
                *reinterpret_cast<float4*>(P.a + (ia[u] | P.offB)) = rb;
            }
        }
    } else {
        float2 va[ITEMS], vb[ITEMS];
        uint64_t ia[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            ia[u] = deposit_zeros(j, P.fixed);
            if (j < P.n_items) {
                va[u] = P.a[ia[u] | P.offA];
                vb[u] = P.a[ia[u] | P.offB];
            }
        }
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            if (j < P.n_items) {
                float2 ra, rb;
                if (MK == MK_PERM) {
                    ra = vb[u]; rb = va[u];
                } else {
                    ra = cdot2(P.m00, va[u], P.m01, vb[u]);
                    rb = cdot2(P.m10, va[u], P.m11, vb[u]);
                }
                P.a[ia[u] | P.offA] = ra;
                P.a[ia[u] | P.offB] = rb;
            }
        }
    }
}

// Target on index bit 0: the pair is the two halves of one 128-bit vector.
template <int MK, int ITEMS>
__global__ void __launch_bounds__(kThreads) k_pair_bit0(const __grid_constant__ PairArgs P) {
    const uint64_t j0 = (uint64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    float4 v[ITEMS];
    uint64_t ia[ITEMS];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
        const uint64_t j = j0 + (uint64_t)u * kThreads;
        ia[u] = deposit_zeros(j, P.fixed) | P.offA;
        if (j < P.n_items) v[u] = *reinterpret_cast<const float4*>(P.a + ia[u]);
    }
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
        const uint64_t j = j0 + (uint64_t)u * kThreads;
        if (j < P.n_items) {
            float4 r;
            if (MK == MK_PERM) {
                r = make_float4(v[u].z, v[u].w, v[u].x, v[u].y);
            } else {
                const float2 x = make_float2(v[u].x, v[u].y), y = make_float2(v[u].z, v[u].w);
                const float2 rx = cdot2(P.m00, x, P.m01, y), ry = cdot2(P.m10, x, P.m11, y);
                r = make_float4(rx.x, rx.y, ry.x, ry.y);
            }
            *reinterpret_cast<float4*>(P.a + ia[u]) = r;
        }
    }
}

struct DiagArgs {
    float2* a;
    uint64_t n_items;
    uint64_t off;      // control values (| target bit when d0 == 1)
    BitList fixed;
    int p;             // target bit position
    int d0_one;        // d0 == 1+0i: amplitudes with target bit 0 are left untouched
    float2 d0, d1;
};

// a[r] <- d(bit p of r) * a[r] on the enumerated amplitudes.
template <int VEC, int ITEMS>
__global__ void __launch_bounds__(kThreads) k_diag(const __grid_constant__ DiagArgs P) {
    const uint64_t j0 = (uint64_t)blockIdx.x * (kThreads * ITEMS) + threadIdx.x;
    if (VEC == 2) {
        float4 v[ITEMS];
        uint64_t ia[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            ia[u] = deposit_zeros(j, P.fixed) | P.off;
            if (j < P.n_items) v[u] = *reinterpret_cast<const float4*>(P.a + ia[u]);
        }
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            if (j < P.n_items) {
                const int b0 = (int)((ia[u] >> P.p) & 1ull);
                const int b1 = (int)(((ia[u] | 1ull) >> P.p) & 1ull);
                float2 x = make_float2(v[u].x, v[u].y), y = make_float2(v[u].z, v[u].w);
                if (b0) x = cmul(P.d1, x); else if (!P.d0_one) x = cmul(P.d0, x);
                if (b1) y = cmul(P.d1, y); else if (!P.d0_one) y = cmul(P.d0, y);
                *reinterpret_cast<float4*>(P.a + ia[u]) = make_float4(x.x, x.y, y.x, y.y);
            }
        }
    } else {
        float2 v[ITEMS];
        uint64_t ia[ITEMS];
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            ia[u] = deposit_zeros(j, P.fixed) | P.off;
            if (j < P.n_items) v[u] = P.a[ia[u]];
        }
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) {
            const uint64_t j = j0 + (uint64_t)u * kThreads;
            if (j < P.n_items) {
                const int b = (int)((ia[u] >> P.p) & 1ull);
                float2 x = v[u];
                if (b) x = cmul(P.d1, x); else if (!P.d0_one) x = cmul(P.d0, x);
                P.a[ia[u]] = x;
            }
        }
    }
}

}  // namespace aqs

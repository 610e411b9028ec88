// Standalone micro-benchmark for the "rounds" structure of the tile kernel (dev tool).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I afquantumsim_b200/csrc \
//        tools/proto/rounds_proto.cu -o gpurun_out/rounds_proto
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
using namespace aqs;

enum { K_GENM = 0, K_REAL, K_XLIKE, K_PERM, K_PHASE, K_LGEN, K_LPERM, K_N };
struct alignas(16) POp { uint8_t mode, tk; uint16_t tl_mask, tl_val, pad0; uint32_t amp_mask, pad1; uint64_t g_mask, g_val; float2 m[4]; };
struct alignas(16) Round { uint8_t swap_j; uint8_t pad[3]; uint16_t cnt[K_N]; uint16_t pad2; uint32_t first_op; uint32_t pad3[2]; };
struct RArgs { float2* state; const Round* rounds; const POp* ops; uint32_t n_rounds; uint32_t n_ops; BitList tile; };

__device__ __forceinline__ void xor_swap(float2& u, float2& v) {
    uint32_t ux = __float_as_uint(u.x), uy = __float_as_uint(u.y), vx = __float_as_uint(v.x), vy = __float_as_uint(v.y);
    ux ^= vx; vx ^= ux; ux ^= vx; uy ^= vy; vy ^= uy; uy ^= vy;
    u = make_float2(__uint_as_float(ux), __uint_as_float(uy)); v = make_float2(__uint_as_float(vx), __uint_as_float(vy));
}
template <int RB, int J> __device__ __forceinline__ void swap_reg_bits(float2 (&a)[1 << RB]) {
#pragma unroll
    for (int k = 0; k < (1 << RB); ++k) if ((k & 1) && !(k >> J & 1)) xor_swap(a[k], a[(k ^ 1) | (1 << J)]);
}
template <int RB, int MODE, bool ALL> __device__ __forceinline__ void reg_pairs(float2 (&a)[1 << RB], const float2 (&m)[4], uint32_t pm) {
#pragma unroll
    for (int p = 0; p < (1 << RB) / 2; ++p) {
        const int k0 = 2 * p, k1 = 2 * p + 1;
        if (ALL || (pm >> p & 1u)) {
            const float2 x = a[k0], y = a[k1];
            if (MODE == K_PERM) xor_swap(a[k0], a[k1]);
            else if (MODE == K_REAL) {
                a[k0] = make_float2(fmaf(m[1].x, y.x, m[0].x * x.x), fmaf(m[1].x, y.y, m[0].x * x.y));
                a[k1] = make_float2(fmaf(m[3].x, y.x, m[2].x * x.x), fmaf(m[3].x, y.y, m[2].x * x.y));
            } else if (MODE == K_XLIKE) {
                a[k0] = make_float2(fmaf(-m[1].y, y.y, m[0].x * x.x), fmaf(m[1].y, y.x, m[0].x * x.y));
                a[k1] = make_float2(fmaf(-m[2].y, x.y, m[3].x * y.x), fmaf(m[2].y, x.x, m[3].x * y.y));
            } else { a[k0] = cdot2(m[0], x, m[1], y); a[k1] = cdot2(m[2], x, m[3], y); }
        }
    }
}

template <int RB>
__global__ void __launch_bounds__(256, (RB == 4 ? 4 : 2)) k_rounds(const __grid_constant__ RArgs P) {
    constexpr int kRegs = 1 << RB;
    float2 a[kRegs];
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t gbase = deposit_zeros((uint64_t)blockIdx.x, P.tile);
    const uint32_t base_local = threadIdx.x;
    __shared__ __align__(16) unsigned char prog[12288];
    {
        const uint4* s1 = reinterpret_cast<const uint4*>(P.rounds); uint4* d1 = reinterpret_cast<uint4*>(prog);
        for (uint32_t i = threadIdx.x; i < P.n_rounds * sizeof(Round) / 16; i += blockDim.x) d1[i] = s1[i];
        const uint4* s2 = reinterpret_cast<const uint4*>(P.ops); uint4* d2 = reinterpret_cast<uint4*>(prog + 4096);
        for (uint32_t i = threadIdx.x; i < P.n_ops * sizeof(POp) / 16; i += blockDim.x) d2[i] = s2[i];
    }
    const Round* rounds = reinterpret_cast<const Round*>(prog);
    const POp* ops = reinterpret_cast<const POp*>(prog + 4096);
#pragma unroll
    for (int k = 0; k < kRegs; ++k) a[k] = P.state[gbase + threadIdx.x + 256 * k];
    __syncthreads();
    auto header = [&](const POp* op, uint32_t& amp_mask, uint32_t& tk, float2 (&m)[4]) -> bool {
        const ulonglong2 gm = *reinterpret_cast<const ulonglong2*>(&op->g_mask);
        const uint4 hd = *reinterpret_cast<const uint4*>(op);
        tk = (hd.x >> 8) & 0xffu;
        const uint32_t tl_mask = hd.x >> 16, tl_val = hd.y & 0xffffu;
        amp_mask = hd.z;
        const float4 m01 = *reinterpret_cast<const float4*>(&op->m[0]);
        const float4 m23 = *reinterpret_cast<const float4*>(&op->m[2]);
        m[0] = make_float2(m01.x, m01.y); m[1] = make_float2(m01.z, m01.w);
        m[2] = make_float2(m23.x, m23.y); m[3] = make_float2(m23.z, m23.w);
        return ((gbase & gm.x) == gm.y) && ((base_local & tl_mask) == tl_val);
    };
    for (uint32_t r = 0; r < P.n_rounds; ++r) {
        const Round rd = rounds[r];
        if (rd.swap_j == 1) swap_reg_bits<RB, 1>(a);
        if (rd.swap_j == 2) swap_reg_bits<RB, 2>(a);
        if (rd.swap_j == 3) swap_reg_bits<RB, 3>(a);
        if (RB > 4 && rd.swap_j == 4) swap_reg_bits<RB, (RB > 4 ? 4 : 3)>(a);
        uint32_t o = rd.first_op;
        uint32_t am, tk; float2 m[4];
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_GENM]; ++i, ++o) if (header(ops + o, am, tk, m)) reg_pairs<RB, K_GENM, false>(a, m, am);
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_REAL]; ++i, ++o) if (header(ops + o, am, tk, m)) reg_pairs<RB, K_REAL, true>(a, m, am);
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_XLIKE]; ++i, ++o) if (header(ops + o, am, tk, m)) reg_pairs<RB, K_XLIKE, true>(a, m, am);
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_PERM]; ++i, ++o) if (header(ops + o, am, tk, m)) reg_pairs<RB, K_PERM, false>(a, m, am);
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_PHASE]; ++i, ++o) if (header(ops + o, am, tk, m)) {
#pragma unroll
            for (int k = 0; k < kRegs; ++k) if (am >> k & 1u) a[k] = cmul(m[0], a[k]);
        }
#pragma unroll 1
        for (uint32_t i = 0; i < rd.cnt[K_LGEN]; ++i, ++o) {
            const bool ok = header(ops + o, am, tk, m);
            const uint32_t xm = 1u << tk; const bool hi = (lane & xm) != 0;
            const uint32_t act = ok ? am : 0u;
            const float2 m_own = hi ? m[3] : m[0], m_oth = hi ? m[2] : m[1];
#pragma unroll
            for (int k = 0; k < kRegs; ++k) {
                float2 other;
                other.x = __shfl_xor_sync(0xffffffffu, a[k].x, xm);
                other.y = __shfl_xor_sync(0xffffffffu, a[k].y, xm);
                if (act >> k & 1u) a[k] = cdot2(m_own, a[k], m_oth, other);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kRegs; ++k) P.state[gbase + threadIdx.x + 256 * k] = a[k];
}

static int g_rb = 4;
struct Prog { std::vector<Round> rounds; std::vector<POp> ops; };
static POp mk(int mode, uint32_t amp_mask, int tk = 0, uint16_t tlm = 0, uint16_t tlv = 0) {
    POp o; memset(&o, 0, sizeof o); o.mode = mode; o.tk = tk; o.amp_mask = amp_mask; o.tl_mask = tlm; o.tl_val = tlv;
    o.m[0] = make_float2(0.8f, 0.f); o.m[1] = make_float2(0.f, -0.6f); o.m[2] = make_float2(0.f, -0.6f); o.m[3] = make_float2(0.8f, 0.f);
    if (mode == K_REAL) { o.m[1] = make_float2(-0.6f, 0.f); o.m[2] = make_float2(0.6f, 0.f); }
    if (mode == K_PHASE) o.m[0] = make_float2(0.8f, 0.6f);
    return o;
}
static void add_round(Prog& p, int swap_j, std::vector<POp> ops) {   // ops must be sorted by kind
    Round r; memset(&r, 0, sizeof r); r.swap_j = swap_j; r.first_op = p.ops.size();
    for (auto& o : ops) { r.cnt[o.mode]++; p.ops.push_back(o); }
    p.rounds.push_back(r);
}

int main(int argc, char** argv) {
    if (argc > 1) g_rb = atoi(argv[1]);
    const uint32_t FULLP = g_rb == 4 ? 0xffu : 0xffffu, FULLA = g_rb == 4 ? 0xffffu : 0xffffffffu, HALFP = g_rb == 4 ? 0x55u : 0x5555u, HALFA = g_rb == 4 ? 0xaaaau : 0xaaaaaaaau;
    const int n = 30;
    float2* st; cudaMalloc(&st, sizeof(float2) << n); cudaMemset(st, 0, sizeof(float2) << n);
    BitList tile; tile.n = 8 + g_rb; for (int i = 0; i < tile.n; ++i) tile.pos[i] = i;   // contiguous tile (memory path is placement-insensitive)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, const Prog& p, int nops) {
        Round* dr; POp* dop;
        cudaMalloc(&dr, p.rounds.size() * sizeof(Round) + 16); cudaMalloc(&dop, p.ops.size() * sizeof(POp) + 16);
        cudaMemcpy(dr, p.rounds.data(), p.rounds.size() * sizeof(Round), cudaMemcpyHostToDevice);
        cudaMemcpy(dop, p.ops.data(), p.ops.size() * sizeof(POp), cudaMemcpyHostToDevice);
        RArgs A; A.state = st; A.rounds = dr; A.ops = dop; A.n_rounds = p.rounds.size(); A.n_ops = p.ops.size(); A.tile = tile;
        for (int i = 0; i < 2; ++i) { if (g_rb == 4) k_rounds<4><<<1u << (n - 12), 256>>>(A); else k_rounds<5><<<1u << (n - 13), 256>>>(A); }
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) { if (g_rb == 4) k_rounds<4><<<1u << (n - 12), 256>>>(A); else k_rounds<5><<<1u << (n - 13), 256>>>(A); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-44s rounds=%3zu ops=%3d  %8.3f ms  (%6.3f ms/op over the 2.45 ms memory floor)  err=%s\n", name, p.rounds.size(), nops, ms,
               nops ? (ms - 2.45f) / nops : 0.f, cudaGetErrorString(cudaGetLastError()));
        cudaFree(dr); cudaFree(dop);
    };
    { Prog p; add_round(p, 0, {mk(K_REAL, FULLP)}); run("1 REAL op", p, 1); }
    { Prog p; std::vector<POp> v(16, mk(K_REAL, FULLP)); add_round(p, 0, v); run("16 REAL ops, one round", p, 16); }
    { Prog p; std::vector<POp> v(16, mk(K_XLIKE, FULLP)); add_round(p, 0, v); run("16 XLIKE ops, one round", p, 16); }
    { Prog p; std::vector<POp> v(16, mk(K_GENM, FULLP)); add_round(p, 0, v); run("16 GEN(masked, all on) ops, one round", p, 16); }
    { Prog p; for (int i = 0; i < 16; ++i) add_round(p, 1 + i % 3, {mk(K_XLIKE, FULLP)}); run("16 x (SWAPBITS + XLIKE)", p, 16); }
    { Prog p; for (int i = 0; i < 16; ++i) add_round(p, 1 + i % 3, {mk(K_REAL, FULLP)}); run("16 x (SWAPBITS + REAL)", p, 16); }
    { Prog p; for (int i = 0; i < 16; ++i) add_round(p, 1 + i % 3, {mk(K_PERM, HALFP)}); run("16 x (SWAPBITS + PERM half pairs)", p, 16); }
    { Prog p; for (int i = 0; i < 16; ++i) add_round(p, 1 + i % 3, {mk(K_PERM, FULLP, 0, 0x2, 0x2)}); run("16 x (SWAPBITS + PERM lane-ctrl)", p, 16); }
    { Prog p; std::vector<POp> v(30, mk(K_PHASE, HALFA)); add_round(p, 0, v); run("30 PHASE half-mask, one round", p, 30); }
    { Prog p; std::vector<POp> v(30, mk(K_PHASE, FULLA, 0, 0x4, 0x4)); add_round(p, 0, v); run("30 PHASE lane-selected, one round", p, 30); }
    { Prog p; std::vector<POp> v; for (int i = 0; i < 5; ++i) v.push_back(mk(K_LGEN, FULLA, i)); add_round(p, 0, v); run("5 LANE_GEN ops", p, 5); }
    { Prog p; for (int l = 0; l < 6; ++l) { for (int j = 0; j < 4; ++j) add_round(p, j ? j : 1, {mk(l % 2 ? K_REAL : K_XLIKE, FULLP)}); add_round(p, 2, {mk(K_PERM, HALFP)}); add_round(p, 3, {mk(K_PERM, HALFP)}); std::vector<POp> v(2, mk(K_PHASE, HALFA)); add_round(p, 0, v); }
      run("brickwork-like mix: 6 x (4 rot + 2 CX + 2 phase)", p, 48); }
    return 0;
}

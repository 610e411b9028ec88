// Microbenchmark: issue/pipe throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, alone and mixed
// with ALU work, to decide whether the fused tile kernel should be written with packed f32x2 math.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b){ uint64_t d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){ uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
constexpr int NACC = 16;
// MODE 0: scalar FFMA (2*NACC chains)   1: FFMA2 (NACC chains)   2: FFMA2 + 1 LOP3 per FFMA2
// MODE 3: FFMA2 with swapped operand + broadcast scalar   4: FFMA2 + 1 MOV-like (alu) per 2   5: scalar FFMA + 1 LOP per FFMA
template <int MODE>
__global__ void __launch_bounds__(256) kb(float2* out, float2 m, int iters, uint32_t seed) {
    float2 a[NACC];
    uint32_t z[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, 1.0f - i); z[i] = seed + i + threadIdx.x; }
    const uint64_t M = pk(m.x, m.y), MB = pk(m.x, m.x), C = pk(0.001f, -0.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0 || MODE == 5) {
                a[i].x = fmaf(m.x, a[i].x, 0.001f);
                a[i].y = fmaf(m.y, a[i].y, -0.001f);
                if (MODE == 5) { z[i] = (z[i] ^ (z[i] >> 3)) & seed; asm volatile("" : "+r"(z[i])); z[i] = (z[i] | (z[i] << 1)) ^ seed; asm volatile("" : "+r"(z[i])); }
            } else {
                uint64_t A = pk(a[i].x, a[i].y);
                uint64_t R;
                if (MODE == 3) R = fma2(MB, pk(a[i].y, a[i].x), C);
                else R = fma2(M, A, C);
                a[i] = *reinterpret_cast<float2*>(&R);
                if (MODE == 2) { z[i] = (z[i] ^ (z[i] >> 3)) & seed; asm volatile("" : "+r"(z[i])); }
                if (MODE == 4 && (i & 1)) { z[i] = (z[i] ^ (z[i] >> 3)) & seed; asm volatile("" : "+r"(z[i])); }
            }
        }
    }
    float2 s = make_float2(0, 0);
    uint32_t zz = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { s.x += a[i].x; s.y += a[i].y; zz ^= z[i]; }
    s.x += (float)zz;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float2* out, int sms, double fma_per_iter_thread, double instr_per_iter_thread) {
    const int iters = 4096, grid = sms * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kb<MODE><<<grid, block>>>(out, make_float2(0.999f, 0.998f), 64, 0xffffu);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        kb<MODE><<<grid, block>>>(out, make_float2(0.999f, 0.998f), iters, 0xffffu);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double threads = (double)grid * block;
    const double warp_instr = threads / 32 * iters * instr_per_iter_thread;
    const double fmas = threads * iters * fma_per_iter_thread;
    std::printf("%-40s %8.3f ms  %7.2f Tfma/s  %6.2f warp-instr/ns/chip  (%.3f per SM per clk @1.965GHz)\n", name, best, fmas / best / 1e9,
                warp_instr / best / 1e6, warp_instr / best / 1e6 / sms / 1.965);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    std::printf("%s, %d SMs\n", p.name, sms);
    float2* out; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256);
    run<0>("scalar FFMA", out, sms, 2.0 * NACC, 2.0 * NACC);
    run<1>("FFMA2", out, sms, 2.0 * NACC, 1.0 * NACC);
    run<2>("FFMA2 + 2 alu per FFMA2 (LOP3,LOP3)", out, sms, 2.0 * NACC, 3.0 * NACC);
    run<3>("FFMA2 swapped+broadcast operands", out, sms, 2.0 * NACC, 1.0 * NACC);
    run<4>("FFMA2 + 1 alu per FFMA2", out, sms, 2.0 * NACC, 2.0 * NACC);
    run<5>("scalar FFMA + 2 alu per FFMA", out, sms, 2.0 * NACC, 6.0 * NACC);
    return 0;
}

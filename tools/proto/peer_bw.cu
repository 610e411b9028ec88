// Dev microbenchmark (2 GPUs, one process): NVLink peer bandwidth of the access patterns the tile kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_bw peer_bw.cu && ./peer_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
typedef unsigned long long u64;

// every warp handles 256-byte runs; run r of the launch sits at (r * stride_runs) % n_runs  (stride_runs odd: a permutation)
__global__ void k_read(const u64* __restrict__ src, u64 n_runs, u64 stride_runs, int per_thread, u64* sink) {
    const u64 warp = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    u64 acc = 0;
    u64 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < per_thread) {
            const u64 run = ((warp * per_thread + i) * stride_runs) % n_runs;
            v[i] = src[run * 32 + lane];
        }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < per_thread) acc ^= v[i];
    if (acc == 0x1234567ull) sink[0] = acc;
}
__global__ void k_write(u64* __restrict__ dst, u64 n_runs, u64 stride_runs, int per_thread) {
    const u64 warp = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i < per_thread) {
            const u64 run = ((warp * per_thread + i) * stride_runs) % n_runs;
            dst[run * 32 + lane] = warp + i;
        }
    }
}
// read remote, write remote (the spanning pass: load a tile part from the peer, store it back there)
__global__ void k_rw(u64* __restrict__ buf, u64 n_runs, u64 stride_runs, int per_thread) {
    const u64 warp = (blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    u64 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < per_thread) v[i] = buf[(((warp * per_thread + i) * stride_runs) % n_runs) * 32 + lane];
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < per_thread) buf[(((warp * per_thread + i) * stride_runs) % n_runs) * 32 + lane] = v[i] + 1;
}
// bulk copies (TMA, 1-D): each CTA pulls `chunk`-byte pieces from src into shared memory, 4 in flight, and pushes them to dst
__global__ void k_bulk(const char* __restrict__ src, char* __restrict__ dst, u64 bytes, int chunk) {
    extern __shared__ __align__(128) char sm[];
    __shared__ __align__(8) u64 bar[4];
    const int S = 4;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const u64 n_chunks = bytes / chunk;
    u64 issued = 0, done = 0;
    unsigned phase[4] = {0, 0, 0, 0};
    for (u64 c = blockIdx.x; c < n_chunks || done < issued; c += gridDim.x) {
        if (c < n_chunks) {
            const int s = issued % S;
            if (issued >= (u64)S) {
                // the store that used this slot must have finished reading shared memory
                asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]), d = (unsigned)__cvta_generic_to_shared(sm + (size_t)s * chunk);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src + c * (u64)chunk), "r"(chunk), "r"(b) : "memory");
            ++issued;
        }
        if (issued - done == (u64)S || c >= n_chunks) {
            const int s = done % S;
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
            unsigned ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(phase[s]) : "memory");
            phase[s] ^= 1;
            if (dst) {
                const u64 cc = blockIdx.x + done * gridDim.x;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + cc * (u64)chunk), "r"((unsigned)__cvta_generic_to_shared(sm + (size_t)s * chunk)), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++done;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const u64 bytes = 4ull << 30, n_runs = bytes / 256;
    u64 *loc, *rem, *sink;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 1, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&loc, bytes)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(loc, 1, bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto report = [&](const char* what, double gb, float ms) { printf("%-58s %8.2f ms  %8.1f GB/s\n", what, ms, gb / (ms * 1e-3)); };
    const u64 strides[] = {1, 4097, 1048577};
    for (int which = 0; which < 2; ++which) {
        u64* p = which ? rem : loc;
        for (u64 st : strides)
            for (int per : {8, 32}) {
                const u64 warps = n_runs / per, blocks = warps * 32 / 256;
                char name[128];
                float ms;
                k_read<<<(unsigned)blocks, 256>>>(p, n_runs, st, per, sink);
                CK(cudaEventRecord(e0)); k_read<<<(unsigned)blocks, 256>>>(p, n_runs, st, per, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                snprintf(name, sizeof name, "%s read   256B runs, run stride %llu, %d loads/thread", which ? "PEER " : "local", st, per);
                report(name, bytes / 1e9, ms);
                k_write<<<(unsigned)blocks, 256>>>(p, n_runs, st, per);
                CK(cudaEventRecord(e0)); k_write<<<(unsigned)blocks, 256>>>(p, n_runs, st, per); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                snprintf(name, sizeof name, "%s write  256B runs, run stride %llu, %d stores/thread", which ? "PEER " : "local", st, per);
                report(name, bytes / 1e9, ms);
                k_rw<<<(unsigned)blocks, 256>>>(p, n_runs, st, per);
                CK(cudaEventRecord(e0)); k_rw<<<(unsigned)blocks, 256>>>(p, n_runs, st, per); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                snprintf(name, sizeof name, "%s rd+wr  256B runs, run stride %llu, %d per thread (bytes each way)", which ? "PEER " : "local", st, per);
                report(name, bytes / 1e9, ms);
            }
    }
    for (int chunk : {256, 1024, 4096, 16384}) {
        for (int mode = 0; mode < 3; ++mode) {
            const char* src = (const char*)(mode == 1 ? loc : rem);
            char* dst = mode == 0 ? nullptr : (char*)(mode == 1 ? rem : rem);
            const char* nm = mode == 0 ? "bulk PEER -> smem" : (mode == 1 ? "bulk local -> smem -> PEER" : "bulk PEER -> smem -> PEER");
            CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
            float ms;
            k_bulk<<<148 * 2, 32, 4 * chunk>>>(src, dst, bytes, chunk);
            CK(cudaEventRecord(e0)); k_bulk<<<148 * 2, 32, 4 * chunk>>>(src, dst, bytes, chunk); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char name[128];
            snprintf(name, sizeof name, "%s, %d-byte pieces, 296 CTAs x 4 in flight", nm, chunk);
            report(name, bytes / 1e9, ms);
        }
    }
    CK(cudaGetLastError());
    return 0;
}

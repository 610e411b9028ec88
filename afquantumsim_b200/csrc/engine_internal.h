// engine_internal.h — types shared by the translation units of libaqs_engine.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "aqs_engine.h"

struct aqs_state_s {
    int n = 0;
    uint64_t N = 0;
    float2* d = nullptr;          // 2^n complex64, interleaved, in HBM
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t alloc_bytes = 0;       // size of the pooled allocation behind d
    bool own_memory = true;       // false for aqs_state_wrap (caller-owned allocation)
    void* scratch = nullptr;      // reductions / sampling workspace
    size_t scratch_bytes = 0;
};

namespace aqs {

// An op translated to index-bit positions (bit p = n-1-qubit).
struct CanonOp {
    int kind;          // aqs_op_kind after demotion (diagonal U2 -> DIAG, 0/1 antidiagonal -> X)
    int p;             // target bit (for SWAP: the higher of the two)
    int p2;            // SWAP: the lower bit, else -1
    uint64_t cmask;    // control bits
    uint64_t cval;     // required values on the control bits
    float2 m[4];       // row-major 2x2
    bool d0_one;       // DIAG with m00 == 1: only amplitudes with the target bit set change
    bool identity;     // no-op
};

int fail(int code, const std::string& msg);
int fail_cuda(cudaError_t e, const char* what, int line);
void count_launch(uint64_t n);
void count_ops(uint64_t n);
void count_h2d(uint64_t bytes);
int sm_count();

int canonicalize(int n, const aqs_op& op, CanonOp& out);
double op_bytes(int n, const CanonOp& c);
int launch_canon(float2* a, int n, const CanonOp& c, cudaStream_t st);   // one per-gate kernel

cudaError_t pool_alloc(void** out, size_t bytes);   // recycled device buffers (engine.cu)
void pool_free(void* p, size_t bytes);

int fused_init();   // opt-in shared-memory size etc. for the tile kernel

}  // namespace aqs

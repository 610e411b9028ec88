// plan_internal.h — the compiled-circuit types shared by plan.cu (planner + launches) and
// specialize.cu (per-pass specialised kernels).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "engine_internal.h"
#include "tile_kernel.cuh"

namespace aqs {

struct SpecKernel;   // specialize.cu: one runtime-compiled kernel, shared between passes of the same shape

struct FusedPass {
    int T = 0;
    uint64_t n_tiles = 0;
    float2 scale = make_float2(1.f, 0.f);
    bool has_scale = false;
    BitList tile;
    uint64_t ld_toff[kMaxThreadBits], ld_roff[kRegBits], st_toff[kMaxThreadBits], st_roff[kRegBits];
    std::vector<TileSeg> segs;
    std::vector<TileOp> ops;
    const DevOp* d_ops = nullptr;   // device copy (plan arena), ops.size() + 1 entries
    bool rare = false;              // some op needs a body that only the full kernel instantiation has (tile_op_is_rare)
    // specialised form (specialize.cu): straight-line kernel for this pass's shape + this pass's coefficient table
    std::shared_ptr<SpecKernel> spec;
    std::vector<uint64_t> spec_coefs;
};

// What the generator produces for one pass.
struct SpecSource {
    std::string src;                 // CUDA C++ (also valid host C++ under -DAQS_HOST_EMU)
    std::vector<uint64_t> coefs;     // this pass's packed coefficient table (kernel parameter)
    int threads = 0;
    size_t smem_bytes = 0;
    uint64_t key = 0;                // hash of src: passes with equal keys share one kernel
};

// specialize.cu
bool spec_generate(int n, const FusedPass& fp, SpecSource& out, std::string& why_not);
int spec_attach(int n, std::vector<FusedPass>& passes, bool wait);     // look up / enqueue kernels for every pass
bool spec_ready(const FusedPass& fp);                                       // kernel compiled and loaded on the current device?
int spec_launch(const FusedPass& fp, float2* state, float2* state_out, uint64_t n_ctas, uint32_t fix_n, uint32_t fix_or,
                const uint8_t* fix_pos, cudaStream_t st);
int spec_wait_all();                                                    // block until every queued compilation has finished
void spec_stats(uint64_t* compiled, uint64_t* cache_hits, uint64_t* failed, double* compile_seconds, uint64_t* pending);

}  // namespace aqs

struct aqs_plan_s {
    int n = 0;
    uint32_t flags = 0;
    std::vector<aqs::CanonOp> ops;       // per-gate path
    std::vector<aqs::FusedPass> passes;  // fused path (empty => run ops one by one)
    void* arena = nullptr;               // device copy of every pass's DevOps
    size_t arena_bytes = 0;
    cudaGraphExec_t graph = nullptr;     // AQS_PLAN_GRAPH: the launch sequence captured for `graph_state`
    const void* graph_state = nullptr;
    bool ran = false;
    bool spec_requested = false;
    aqs_plan_info info{};
};

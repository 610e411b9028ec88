// plan.cu — compiled circuits: aqs_plan_build / aqs_plan_run.
// Replaces QCircuit::compile (src/quantum.cpp:199-210): the reference multiplies
// every gate into a dense 2^n x 2^n unitary; here "compiling" means turning the
// op list into a launch plan (and, with AQS_PLAN_FUSE, fusing it into tile passes).
#include <new>

#include "engine_internal.h"

struct aqs_plan_s {
    int n = 0;
    uint32_t flags = 0;
    std::vector<aqs::CanonOp> ops;
    aqs_plan_info info{};
};

namespace aqs {
int fused_init() { return AQS_OK; }
}  // namespace aqs

using namespace aqs;

extern "C" {

int aqs_plan_build(int n, const aqs_op* ops, uint64_t n_ops, uint32_t flags, aqs_plan_t* out) {
    if (!out) return fail(AQS_ERR_INVALID, "null output handle");
    if (n < 1 || n > AQS_MAX_QUBITS) return fail(AQS_ERR_INVALID, "qubit count out of range");
    if (!ops && n_ops) return fail(AQS_ERR_INVALID, "null op list");
    aqs_plan_s* p = new (std::nothrow) aqs_plan_s();
    if (!p) return fail(AQS_ERR_NOMEM, "host allocation failed");
    p->n = n;
    p->flags = flags;
    p->ops.reserve(n_ops);
    double bytes = 0.0;
    for (uint64_t i = 0; i < n_ops; ++i) {
        CanonOp c;
        int rc = canonicalize(n, ops[i], c);
        if (rc) { delete p; return rc; }
        bytes += op_bytes(n, c);
        p->ops.push_back(c);
    }
    p->info.n_ops = n_ops;
    p->info.n_launches = n_ops;
    p->info.n_fused_passes = 0;
    p->info.n_single_ops = n_ops;
    p->info.bytes_unfused = bytes;
    p->info.bytes_planned = bytes;
    p->info.n_qubits = n;
    p->info.tile_bits = 0;
    *out = p;
    return AQS_OK;
}

int aqs_plan_run(aqs_state_t s, aqs_plan_t p) {
    if (!s || !p) return fail(AQS_ERR_INVALID, "null handle");
    if (s->n != p->n) return fail(AQS_ERR_INVALID, "plan and state have different qubit counts");
    for (const CanonOp& c : p->ops) {
        int rc = launch_canon(s->d, s->n, c, s->stream);
        if (rc) return rc;
    }
    count_ops(p->ops.size());
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "plan run", __LINE__);
    return AQS_OK;
}

int aqs_plan_get_info(aqs_plan_t p, aqs_plan_info* info) {
    if (!p || !info) return fail(AQS_ERR_INVALID, "null argument");
    *info = p->info;
    return AQS_OK;
}

int aqs_plan_destroy(aqs_plan_t p) {
    delete p;
    return AQS_OK;
}

}  // extern "C"

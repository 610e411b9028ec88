// engine.cu — C ABI of libaqs_engine.so (see include/aqs_engine.h).
// State management, per-gate dispatch, measurement entry points, timers.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "aqs_engine.h"
#include "engine_internal.h"
#include "gate_kernels.cuh"
#include "measure_kernels.cuh"
#include "peer_kernels.cuh"
#include "dense_kernels.cuh"
#include "dense_tc.cuh"

namespace aqs {

static thread_local std::string t_err;
static bool g_inited = false;
static int g_device = 0;
static int g_sm_count = 148;
static size_t g_hbm_bytes = 0;
static std::atomic<uint64_t> c_launches{0}, c_ops{0}, c_h2d{0}, c_d2h{0};

int fail(int code, const std::string& msg) {
    t_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at engine line %d: %s", (int)e, cudaGetErrorString(e), line, what);
    t_err = buf;
    return e == cudaErrorMemoryAllocation ? AQS_ERR_NOMEM : AQS_ERR_CUDA;
}
void count_launch(uint64_t n) { c_launches.fetch_add(n, std::memory_order_relaxed); }
void count_ops(uint64_t n) { c_ops.fetch_add(n, std::memory_order_relaxed); }
void count_h2d(uint64_t b) { c_h2d.fetch_add(b, std::memory_order_relaxed); }
int sm_count() { return g_sm_count; }

#define CUDA_TRY(x)                                                           \
    do {                                                                      \
        cudaError_t e_ = (x);                                                 \
        if (e_ != cudaSuccess) return aqs::fail_cuda(e_, #x, __LINE__);       \
    } while (0)
#define REQUIRE_INIT()                                                                          \
    do {                                                                                        \
        if (!aqs::g_inited) return aqs::fail(AQS_ERR_STATE, "aqs_engine_init has not been called"); \
    } while (0)
#define REQUIRE(cond, msg)                                      \
    do {                                                        \
        if (!(cond)) return aqs::fail(AQS_ERR_INVALID, msg);    \
    } while (0)

static inline uint64_t qmask_to_pos(int n, uint64_t qmask) {
    uint64_t m = 0;
    for (int q = 0; q < n; ++q)
        if (qmask >> q & 1ull) m |= 1ull << (n - 1 - q);
    return m;
}

static void bitlist_from_mask(uint64_t mask, BitList& f) {
    f.n = 0;
    for (int p = 0; p < 64; ++p)
        if (mask >> p & 1ull) f.pos[f.n++] = (uint8_t)p;
}

static inline bool is_one(aqs_c32 z) { return z.re == 1.f && z.im == 0.f; }
static inline bool is_zero(aqs_c32 z) { return z.re == 0.f && z.im == 0.f; }

// Canonical form of an op in bit-position space.
int canonicalize(int n, const aqs_op& op, CanonOp& c) {
    REQUIRE(op.kind >= AQS_OP_U2 && op.kind <= AQS_OP_SWAP, "unknown op kind");
    REQUIRE(op.target >= 0 && op.target < n, "target qubit out of range");
    const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    REQUIRE((op.ctrl_mask & ~all) == 0, "control mask names a qubit outside the state");
    REQUIRE(!(op.ctrl_mask >> op.target & 1ull), "control qubit cannot be the target qubit");
    c.kind = op.kind;
    c.p = n - 1 - op.target;
    c.p2 = -1;
    c.cmask = qmask_to_pos(n, op.ctrl_mask);
    c.cval = qmask_to_pos(n, op.ctrl_value & op.ctrl_mask);
    for (int i = 0; i < 4; ++i) c.m[i] = make_float2(op.m[i].re, op.m[i].im);
    c.d0_one = false;
    if (op.kind == AQS_OP_SWAP) {
        REQUIRE(op.target2 >= 0 && op.target2 < n, "second swap qubit out of range");
        REQUIRE(op.target2 != op.target, "cannot swap a qubit with itself");
        REQUIRE(!(op.ctrl_mask >> op.target2 & 1ull), "control qubit cannot be a swap target");
        c.p2 = n - 1 - op.target2;
        if (c.p2 > c.p) std::swap(c.p, c.p2);
    } else if (op.kind == AQS_OP_U2) {
        // demote structured matrices so they run on the cheaper kernels
        if (is_zero(op.m[1]) && is_zero(op.m[2])) c.kind = AQS_OP_DIAG;
        else if (is_zero(op.m[0]) && is_zero(op.m[3]) && is_one(op.m[1]) && is_one(op.m[2])) c.kind = AQS_OP_X;
    }
    if (c.kind == AQS_OP_DIAG) {
        c.m[1] = c.m[2] = make_float2(0.f, 0.f);
        c.d0_one = is_one(op.m[0]);
    }
    c.identity = (c.kind == AQS_OP_DIAG && c.d0_one && is_one(op.m[3]));
    return AQS_OK;
}

// algorithmic HBM bytes of one op run alone
double op_bytes(int n, const CanonOp& c) {
    const double S = 8.0 * std::ldexp(1.0, n);
    int nc = __builtin_popcountll(c.cmask);
    double frac = std::ldexp(1.0, -nc);
    if (c.identity) return 0.0;
    if (c.kind == AQS_OP_DIAG && c.d0_one) frac *= 0.5;
    if (c.kind == AQS_OP_SWAP) frac *= 0.5;
    return 2.0 * S * frac;
}

template <typename K, typename A>
static int launch_items(K kernel, const A& args, uint64_t items, int per_block, cudaStream_t st) {
    const uint64_t blocks = (items + per_block - 1) / per_block;
    if (blocks == 0) return AQS_OK;
    if (blocks > 0x7fffffffull) return fail(AQS_ERR_INVALID, "grid too large");
    kernel<<<(unsigned)blocks, kThreads, 0, st>>>(args);
    count_launch(1);
    return AQS_OK;
}

constexpr int kPairItems = 2;   // 2 x 2 x 16 B in flight per thread
constexpr int kBit0Items = 4;
constexpr int kDiagItems = 4;

int launch_canon(float2* a, int n, const CanonOp& c, cudaStream_t st) {
    if (c.identity) return AQS_OK;
    const uint64_t N = 1ull << n;
    if (c.kind == AQS_OP_DIAG) {
        DiagArgs A;
        A.a = a; A.p = c.p; A.d0_one = c.d0_one ? 1 : 0; A.d0 = c.m[0]; A.d1 = c.m[3];
        uint64_t fixedmask = c.cmask;
        A.off = c.cval;
        if (c.d0_one) { fixedmask |= 1ull << c.p; A.off |= 1ull << c.p; }
        if (!(fixedmask & 1ull) && n >= 1 && N >= 2) {
            bitlist_from_mask(fixedmask | 1ull, A.fixed);
            A.n_items = N >> A.fixed.n;
            return launch_items(k_diag<2, kDiagItems>, A, A.n_items, kThreads * kDiagItems, st);
        }
        bitlist_from_mask(fixedmask, A.fixed);
        A.n_items = N >> A.fixed.n;
        return launch_items(k_diag<1, kDiagItems>, A, A.n_items, kThreads * kDiagItems, st);
    }
    PairArgs A;
    A.a = a;
    A.m00 = c.m[0]; A.m01 = c.m[1]; A.m10 = c.m[2]; A.m11 = c.m[3];
    const bool perm = (c.kind == AQS_OP_X || c.kind == AQS_OP_SWAP);
    uint64_t fixedmask = c.cmask | (1ull << c.p);
    if (c.kind == AQS_OP_SWAP) {
        fixedmask |= 1ull << c.p2;
        A.offA = c.cval | (1ull << c.p);
        A.offB = c.cval | (1ull << c.p2);
    } else {
        A.offA = c.cval;
        A.offB = c.cval | (1ull << c.p);
    }
    if (c.kind != AQS_OP_SWAP && c.p == 0) {
        // pair inside one 128-bit vector
        bitlist_from_mask(fixedmask, A.fixed);
        A.n_items = N >> A.fixed.n;
        if (perm) return launch_items(k_pair_bit0<MK_PERM, kBit0Items>, A, A.n_items, kThreads * kBit0Items, st);
        return launch_items(k_pair_bit0<MK_GENERAL, kBit0Items>, A, A.n_items, kThreads * kBit0Items, st);
    }
    if (!(fixedmask & 1ull)) {
        bitlist_from_mask(fixedmask | 1ull, A.fixed);
        A.n_items = N >> A.fixed.n;
        if (perm) return launch_items(k_pair<2, MK_PERM, kPairItems>, A, A.n_items, kThreads * kPairItems, st);
        return launch_items(k_pair<2, MK_GENERAL, kPairItems>, A, A.n_items, kThreads * kPairItems, st);
    }
    bitlist_from_mask(fixedmask, A.fixed);
    A.n_items = N >> A.fixed.n;
    if (perm) return launch_items(k_pair<1, MK_PERM, kPairItems>, A, A.n_items, kThreads * kPairItems, st);
    return launch_items(k_pair<1, MK_GENERAL, kPairItems>, A, A.n_items, kThreads * kPairItems, st);
}

// Device buffers and streams are recycled.  cudaMalloc / cudaFree / cudaStreamDestroy take
// milliseconds to (measured, on a shared host) hundreds of milliseconds and synchronise the
// device, and the API's normal use — a QSimulator and a plan per run — would call them every
// time.  Sizes are rounded to powers of two (>= 256 B) so that buffers are interchangeable.
struct PoolEntry { void* p; size_t bytes; };
static std::vector<PoolEntry> g_pool;
static std::vector<cudaStream_t> g_stream_pool;
static std::mutex g_pool_mu;
constexpr size_t kPoolMaxEntries = 16;
constexpr size_t kPoolMaxBigEntries = 2;            // buffers of 1 GiB and more

static size_t pool_round(size_t bytes) {
    size_t r = 256;
    while (r < bytes) r <<= 1;
    return r;
}
cudaError_t pool_alloc(void** out, size_t bytes) {
    bytes = pool_round(bytes);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].bytes == bytes) {
                *out = g_pool[i].p;
                g_pool.erase(g_pool.begin() + i);
                return cudaSuccess;
            }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {          // give cached buffers back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (auto& pe : g_pool) cudaFree(pe.p);
        g_pool.clear();
        e = cudaMalloc(out, bytes);
    }
    return e;
}
void pool_free(void* p, size_t bytes) {
    if (!p) return;
    bytes = pool_round(bytes);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t big = 0;
    for (auto& pe : g_pool) big += pe.bytes >= (1ull << 30);
    if (g_pool.size() >= kPoolMaxEntries || (bytes >= (1ull << 30) && big >= kPoolMaxBigEntries)) {
        // evict the oldest entry of the same class
        for (size_t i = 0; i < g_pool.size(); ++i)
            if ((g_pool[i].bytes >= (1ull << 30)) == (bytes >= (1ull << 30)) || g_pool.size() >= kPoolMaxEntries) {
                cudaFree(g_pool[i].p);
                g_pool.erase(g_pool.begin() + i);
                break;
            }
    }
    g_pool.push_back({p, bytes});
}
static void pool_release_all() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& pe : g_pool) cudaFree(pe.p);
    g_pool.clear();
    for (auto st : g_stream_pool) cudaStreamDestroy(st);
    g_stream_pool.clear();
}
static cudaError_t stream_acquire(cudaStream_t* out) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_stream_pool.empty()) {
            *out = g_stream_pool.back();
            g_stream_pool.pop_back();
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
static void stream_release(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_stream_pool.size() < 16) g_stream_pool.push_back(st);
    else cudaStreamDestroy(st);
}

static int ensure_scratch(aqs_state_s* s, size_t bytes) {
    if (s->scratch_bytes >= bytes) return AQS_OK;
    if (s->scratch) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        pool_free(s->scratch, s->scratch_bytes);
    }
    s->scratch = nullptr; s->scratch_bytes = 0;
    CUDA_TRY(pool_alloc(&s->scratch, bytes));
    s->scratch_bytes = bytes;
    return AQS_OK;
}

static unsigned stream_blocks(uint64_t work_threads) {
    uint64_t b = (work_threads + 255) / 256;
    uint64_t cap = (uint64_t)g_sm_count * 8;
    return (unsigned)std::max<uint64_t>(1, std::min(b, cap));
}

}  // namespace aqs

using namespace aqs;

extern "C" {

int aqs_engine_abi_version(void) { return AQS_ENGINE_ABI_VERSION; }
const char* aqs_last_error(void) { return t_err.c_str(); }

int aqs_engine_init(int device) {
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    REQUIRE(count > 0, "no CUDA device visible");
    REQUIRE(device >= 0 && device < count, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    g_device = device;
    g_sm_count = prop.multiProcessorCount;
    g_hbm_bytes = prop.totalGlobalMem;
    if (prop.major < 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
        return fail(AQS_ERR_STATE, buf);
    }
    int rc = aqs::fused_init();
    if (rc != AQS_OK) return rc;
    g_inited = true;
    return AQS_OK;
}

int aqs_ipc_close_all(void);
int aqs_engine_shutdown(void) {
    aqs_ipc_close_all();
    pool_release_all();
    g_inited = false;
    return AQS_OK;
}

int aqs_engine_device(int* device, int* sms, size_t* hbm) {
    REQUIRE_INIT();
    if (device) *device = g_device;
    if (sms) *sms = g_sm_count;
    if (hbm) *hbm = g_hbm_bytes;
    return AQS_OK;
}

static int state_alloc(int n, aqs_state_t* out, bool set_zero_state) {
    REQUIRE_INIT();
    REQUIRE(out != nullptr, "null output handle");
    REQUIRE(n >= 1 && n <= AQS_MAX_QUBITS, "qubit count must be in [1, AQS_MAX_QUBITS]");
    CUDA_TRY(cudaSetDevice(g_device));
    aqs_state_s* s = new (std::nothrow) aqs_state_s();
    if (!s) return fail(AQS_ERR_NOMEM, "host allocation failed");
    s->n = n;
    s->N = 1ull << n;
    const size_t bytes = std::max<size_t>(s->N * sizeof(float2), 16);
    cudaError_t e = pool_alloc((void**)&s->d, bytes);
    if (e != cudaSuccess) { delete s; return fail_cuda(e, "cudaMalloc(state)", __LINE__); }
    s->alloc_bytes = bytes;
    e = stream_acquire(&s->stream);
    if (e != cudaSuccess) { pool_free(s->d, bytes); delete s; return fail_cuda(e, "cudaStreamCreate", __LINE__); }
    s->own_stream = true;
    *out = s;
    return set_zero_state ? aqs_state_set_basis(s, 0) : AQS_OK;
}

int aqs_state_create(int n, aqs_state_t* out) { return state_alloc(n, out, true); }

int aqs_state_wrap(int n, void* device_ptr, aqs_state_t* out) {
    REQUIRE_INIT();
    REQUIRE(out != nullptr && device_ptr != nullptr, "null argument");
    REQUIRE(n >= 1 && n <= AQS_MAX_QUBITS, "qubit count must be in [1, AQS_MAX_QUBITS]");
    REQUIRE(((uintptr_t)device_ptr & 15u) == 0, "device pointer must be 16-byte aligned");
    CUDA_TRY(cudaSetDevice(g_device));
    aqs_state_s* s = new (std::nothrow) aqs_state_s();
    if (!s) return fail(AQS_ERR_NOMEM, "host allocation failed");
    s->n = n;
    s->N = 1ull << n;
    s->d = (float2*)device_ptr;
    s->own_memory = false;
    cudaError_t e = stream_acquire(&s->stream);
    if (e != cudaSuccess) { delete s; return fail_cuda(e, "cudaStreamCreate", __LINE__); }
    s->own_stream = true;
    *out = s;
    return AQS_OK;
}

int aqs_state_destroy(aqs_state_t s) {
    if (!s) return AQS_OK;
    cudaStreamSynchronize(s->stream);
    if (s->scratch) pool_free(s->scratch, s->scratch_bytes);
    if (s->d && s->own_memory) pool_free(s->d, s->alloc_bytes);
    if (s->own_stream) stream_release(s->stream);
    delete s;
    return AQS_OK;
}

int aqs_state_clone(aqs_state_t src, aqs_state_t* out) {
    REQUIRE_INIT();
    REQUIRE(src && out, "null handle");
    int rc = state_alloc(src->n, out, false);
    if (rc) return rc;
    // The copy runs on the SOURCE's stream: it is ordered after everything queued on src, later work on src (and
    // src's destruction, which synchronises that stream before the buffer returns to the pool) is ordered after it,
    // and the clone's own stream waits for it through an event.  No host synchronisation.
    cudaError_t e = cudaMemcpyAsync((*out)->d, src->d, src->N * sizeof(float2), cudaMemcpyDeviceToDevice, src->stream);
    cudaEvent_t ev = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev, src->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent((*out)->stream, ev, 0);
    if (ev) cudaEventDestroy(ev);
    if (e != cudaSuccess) { aqs_state_destroy(*out); *out = nullptr; return fail_cuda(e, "state clone", __LINE__); }
    return AQS_OK;
}

int aqs_state_qubits(aqs_state_t s, int* n) {
    REQUIRE(s && n, "null handle");
    *n = s->n;
    return AQS_OK;
}

int aqs_state_set_basis(aqs_state_t s, uint64_t index) {
    REQUIRE_INIT();
    REQUIRE(s, "null handle");
    REQUIRE(index < s->N, "basis index out of range");
    CUDA_TRY(cudaMemsetAsync(s->d, 0, s->N * sizeof(float2), s->stream));
    k_set_one<<<1, 1, 0, s->stream>>>(s->d, index);
    count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

int aqs_state_set_product(aqs_state_t s, const aqs_c32* q) {
    REQUIRE_INIT();
    REQUIRE(s && q, "null argument");
    ProductArgs P;
    P.n = s->n;
    for (int k = 0; k < s->n; ++k) {
        P.q[k][0] = make_float2(q[2 * k].re, q[2 * k].im);
        P.q[k][1] = make_float2(q[2 * k + 1].re, q[2 * k + 1].im);
    }
    k_set_product<<<stream_blocks(s->N), 256, 0, s->stream>>>(s->d, s->N, P);
    count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

int aqs_state_set_identity(aqs_state_t s) {
    REQUIRE_INIT();
    REQUIRE(s, "null handle");
    REQUIRE((s->n & 1) == 0, "identity needs an even qubit count (2m)");
    CUDA_TRY(cudaMemsetAsync(s->d, 0, s->N * sizeof(float2), s->stream));
    k_set_diag_ones<<<stream_blocks(1ull << (s->n / 2)), 256, 0, s->stream>>>(s->d, s->n / 2);
    count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

int aqs_state_upload(aqs_state_t s, const aqs_c32* host, uint64_t offset, uint64_t count) {
    REQUIRE_INIT();
    REQUIRE(s && host, "null argument");
    REQUIRE(offset <= s->N && count <= s->N - offset, "range outside the state");
    CUDA_TRY(cudaMemcpyAsync(s->d + offset, host, count * sizeof(float2), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    c_h2d += count * sizeof(float2);
    return AQS_OK;
}

int aqs_state_download(aqs_state_t s, aqs_c32* host, uint64_t offset, uint64_t count) {
    REQUIRE_INIT();
    REQUIRE(s && host, "null argument");
    REQUIRE(offset <= s->N && count <= s->N - offset, "range outside the state");
    CUDA_TRY(cudaMemcpyAsync(host, s->d + offset, count * sizeof(float2), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    c_d2h += count * sizeof(float2);
    return AQS_OK;
}

int aqs_state_get_amp(aqs_state_t s, uint64_t index, aqs_c32* out) { return aqs_state_download(s, out, index, 1); }

int aqs_state_device_ptr(aqs_state_t s, void** dptr) {
    REQUIRE(s && dptr, "null argument");
    *dptr = s->d;
    return AQS_OK;
}

int aqs_state_set_stream(aqs_state_t s, void* stream) {
    REQUIRE(s, "null handle");
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->own_stream) stream_release(s->stream);
    s->stream = (cudaStream_t)stream;
    s->own_stream = false;
    return AQS_OK;
}

int aqs_state_get_stream(aqs_state_t s, void** stream) {
    REQUIRE(s && stream, "null argument");
    *stream = (void*)s->stream;
    return AQS_OK;
}

int aqs_sync(aqs_state_t s) {
    REQUIRE(s, "null handle");
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return AQS_OK;
}

int aqs_apply_op(aqs_state_t s, const aqs_op* op) {
    REQUIRE_INIT();
    REQUIRE(s && op, "null argument");
    CanonOp c;
    int rc = canonicalize(s->n, *op, c);
    if (rc) return rc;
    rc = launch_canon(s->d, s->n, c, s->stream);
    if (rc) return rc;
    count_ops(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

int aqs_apply_ops(aqs_state_t s, const aqs_op* ops, uint64_t n_ops) {
    REQUIRE_INIT();
    REQUIRE(s && (ops || n_ops == 0), "null argument");
    for (uint64_t i = 0; i < n_ops; ++i) {
        int rc = aqs_apply_op(s, ops + i);
        if (rc) return rc;
    }
    return AQS_OK;
}

// ---- probabilities / measurement ------------------------------------------------
int aqs_norm2(aqs_state_t s, double* out) {
    REQUIRE_INIT();
    REQUIRE(s && out, "null argument");
    int rc = ensure_scratch(s, 64);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(s->scratch, 0, sizeof(double), s->stream));
    k_norm2<<<stream_blocks(s->N), 256, 0, s->stream>>>(s->d, s->N, (double*)s->scratch);
    count_launch(1);
    CUDA_TRY(cudaMemcpyAsync(out, s->scratch, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    c_d2h += 8;
    return AQS_OK;
}

int aqs_scale(aqs_state_t s, float f) {
    REQUIRE_INIT();
    REQUIRE(s, "null handle");
    k_scale<<<stream_blocks(s->N), 256, 0, s->stream>>>(s->d, s->N, f);
    count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

int aqs_prob_fixed(aqs_state_t s, uint64_t qmask, uint64_t qvalue, uint64_t* out) {
    REQUIRE_INIT();
    REQUIRE(s && out, "null argument");
    const uint64_t all = (1ull << s->n) - 1ull;
    REQUIRE((qmask & ~all) == 0, "mask names a qubit outside the state");
    int rc = ensure_scratch(s, 64);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(s->scratch, 0, sizeof(unsigned long long), s->stream));
    k_prob_fixed<<<stream_blocks(s->N / 2), 256, 0, s->stream>>>(s->d, s->N, qmask_to_pos(s->n, qmask),
                                                              qmask_to_pos(s->n, qvalue & qmask),
                                                              (unsigned long long*)s->scratch);
    count_launch(1);
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, s->scratch, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    c_d2h += 8;
    *out = h;
    return AQS_OK;
}

int aqs_qubit_prob1(aqs_state_t s, int qubit, double* out) {
    REQUIRE(s && out, "null argument");
    REQUIRE(qubit >= 0 && qubit < s->n, "qubit out of range");
    uint64_t f = 0;
    int rc = aqs_prob_fixed(s, 1ull << qubit, 1ull << qubit, &f);
    if (rc) return rc;
    *out = (double)f * 0x1p-62;
    return AQS_OK;
}

int aqs_probabilities(aqs_state_t s, float* host_out, uint64_t offset, uint64_t count) {
    REQUIRE_INIT();
    REQUIRE(s && host_out, "null argument");
    REQUIRE(offset <= s->N && count <= s->N - offset, "range outside the state");
    if (count == 0) return AQS_OK;
    float* tmp = nullptr;
    CUDA_TRY(pool_alloc((void**)&tmp, count * sizeof(float)));
    k_probabilities<<<stream_blocks(count), 256, 0, s->stream>>>(s->d + offset, count, tmp);
    count_launch(1);
    cudaError_t e = cudaMemcpyAsync(host_out, tmp, count * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    else cudaStreamSynchronize(s->stream);
    pool_free(tmp, count * sizeof(float));
    if (e != cudaSuccess) return fail_cuda(e, "probabilities copy", __LINE__);
    c_d2h += count * sizeof(float);
    return AQS_OK;
}

int aqs_collapse_qubit(aqs_state_t s, int qubit, int outcome, float p) {
    REQUIRE_INIT();
    REQUIRE(s, "null handle");
    REQUIRE(qubit >= 0 && qubit < s->n, "qubit out of range");
    REQUIRE(outcome == 0 || outcome == 1, "outcome must be 0 or 1");
    k_collapse<<<stream_blocks(s->N), 256, 0, s->stream>>>(s->d, s->N, 1ull << (s->n - 1 - qubit), outcome, p);
    count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return AQS_OK;
}

static int sample_impl(aqs_state_t s, const float* u_host, const uint64_t* ufix_host, uint64_t n_draws, uint64_t* out_host,
                       uint32_t* hist_host) {
    REQUIRE_INIT();
    REQUIRE(s && (u_host || ufix_host || n_draws == 0), "null argument");
    REQUIRE(n_draws <= 0x7fffffffull, "too many draws for one call");
    const uint64_t tile_amps = std::min<uint64_t>(s->N, kTileAmps);
    const uint64_t n_tiles = s->N / tile_amps;
    // scratch layout: [tile sums: n_tiles u64][u: n_draws f32 or u64 (padded)][out: n_draws u64]
    const size_t off_u = ((n_tiles * 8 + 255) / 256) * 256;
    const size_t off_o = off_u + ((n_draws * 8 + 255) / 256) * 256;
    const size_t need = off_o + n_draws * 8 + 256;
    int rc = ensure_scratch(s, need);
    if (rc) return rc;
    char* base = (char*)s->scratch;
    unsigned long long* sums = (unsigned long long*)base;
    float* u_dev = (float*)(base + off_u);
    unsigned long long* out_dev = (unsigned long long*)(base + off_o);
    k_tile_sums<<<(unsigned)n_tiles, 256, 0, s->stream>>>(s->d, tile_amps, sums);
    k_scan_tiles<<<1, 1024, 0, s->stream>>>(sums, n_tiles);
    count_launch(2);
    cudaError_t e = cudaSuccess;
    if (n_draws) {
        const size_t ub = ufix_host ? 8 : 4;
        e = cudaMemcpyAsync(u_dev, ufix_host ? (const void*)ufix_host : (const void*)u_host, n_draws * ub,
                            cudaMemcpyHostToDevice, s->stream);
        c_h2d += n_draws * ub;
        if (e == cudaSuccess) {
            k_sample<<<(unsigned)n_draws, 256, 0, s->stream>>>(s->d, tile_amps, n_tiles, sums, ufix_host ? nullptr : u_dev,
                                                              ufix_host ? (const unsigned long long*)u_dev : nullptr,
                                                              out_dev, nullptr);
            count_launch(1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && out_host) {
            e = cudaMemcpyAsync(out_host, out_dev, n_draws * 8, cudaMemcpyDeviceToHost, s->stream);
            c_d2h += n_draws * 8;
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    else cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail_cuda(e, "sampling", __LINE__);
    return AQS_OK;
}

// Histograms are built on the HOST from the n outcome indices (8 bytes per draw come back from the device, not a
// dense uint32[2^n]: 4 GiB at n = 30 for 1000 draws).  Sorted (index, count) pairs first; the dense vector of the
// reference's return type (src/quantum.cpp:467-501) is filled from them.
static int sample_sorted(aqs_state_t s, const float* u, uint64_t n, std::vector<uint64_t>& out) {
    out.resize(n);
    int rc = sample_impl(s, u, nullptr, n, out.data(), nullptr);
    if (rc) return rc;
    std::sort(out.begin(), out.end());
    return AQS_OK;
}

int aqs_sample(aqs_state_t s, const float* u, uint64_t n, uint64_t* out) {
    REQUIRE(out || n == 0, "null output");
    return sample_impl(s, u, nullptr, n, out, nullptr);
}
int aqs_sample_fixed(aqs_state_t s, const uint64_t* u, uint64_t n, uint64_t* out) {
    REQUIRE((out && u) || n == 0, "null argument");
    return sample_impl(s, nullptr, u, n, out, nullptr);
}
int aqs_sample_hist(aqs_state_t s, const float* u, uint64_t n, uint32_t* hist) {
    REQUIRE(s && hist, "null argument");
    REQUIRE(u || n == 0, "null draws");
    std::vector<uint64_t> idx;
    int rc = sample_sorted(s, u, n, idx);
    if (rc) return rc;
    std::memset(hist, 0, s->N * sizeof(uint32_t));
    for (uint64_t k : idx) ++hist[k];
    return AQS_OK;
}
int aqs_sample_hist_sparse(aqs_state_t s, const float* u, uint64_t n, uint64_t* index, uint32_t* count, uint64_t cap, uint64_t* n_bins) {
    REQUIRE(s && n_bins, "null argument");
    REQUIRE(u || n == 0, "null draws");
    std::vector<uint64_t> idx;
    int rc = sample_sorted(s, u, n, idx);
    if (rc) return rc;
    uint64_t bins = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && idx[j] == idx[i]) ++j;
        if (index && count && bins < cap) { index[bins] = idx[i]; count[bins] = (uint32_t)(j - i); }
        ++bins;
        i = j;
    }
    *n_bins = bins;
    if (index && count && bins > cap) return fail(AQS_ERR_INVALID, "histogram has more bins than the output arrays hold");
    return AQS_OK;
}
int aqs_pool_trim(void) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& pe : g_pool) cudaFree(pe.p);
    g_pool.clear();
    return AQS_OK;
}

// ---- opaque k-qubit matrices ----------------------------------------------------------
extern "C++" {
template <int K>
static cudaError_t launch_dense(const DenseArgs& A, cudaStream_t st) {
    const size_t smem = sizeof(float4) << (2 * K);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_dense<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const uint64_t blocks = (A.n_groups + kDenseThreads - 1) / kDenseThreads;
    k_dense<K><<<(unsigned)blocks, kDenseThreads, smem, st>>>(A);
    return cudaGetLastError();
}
}  // extern "C++"

int aqs_apply_dense(aqs_state_t s, const int* qubits, int k, uint64_t ctrl_mask, uint64_t ctrl_value, const aqs_c32* m) {
    REQUIRE_INIT();
    REQUIRE(s && qubits && m, "null argument");
    REQUIRE(k >= 1 && k <= kDenseMaxK, "dense blocks of 1 to 6 qubits");
    REQUIRE(k <= s->n, "more target qubits than the state has");
    const int n = s->n;
    const int D = 1 << k;
    DenseArgs A;
    std::memset(&A, 0, sizeof A);
    A.a = s->d;
    uint64_t fixedmask = 0;
    for (int i = 0; i < k; ++i) {
        REQUIRE(qubits[i] >= 0 && qubits[i] < n, "target qubit out of range");
        const int p = n - 1 - qubits[i];
        REQUIRE(!(fixedmask >> p & 1ull), "duplicate target qubit");
        fixedmask |= 1ull << p;
        A.toff[k - 1 - i] = 1ull << p;                  // qubits[0] is the most significant bit of the matrix index
    }
    const uint64_t cm = qmask_to_pos(n, ctrl_mask), cv = qmask_to_pos(n, ctrl_value & ctrl_mask);
    REQUIRE((ctrl_mask >> n) == 0 || n >= 64, "control qubit out of range");
    REQUIRE(!(cm & fixedmask), "a control qubit is also a target");
    A.ctrl_or = cv;
    bitlist_from_mask(fixedmask | cm, A.fixed);
    A.n_groups = s->N >> A.fixed.n;
    // ---- tensor-core path (dense_tc.cuh): k = 4, 5, 6 as a 64 x 64 matrix (k < 6: I (x) M on the lowest free index bits,
    // which also lengthens the contiguous runs), 3xTF32 on tcgen05 with the accumulator in tensor memory
    {
        const char* e = std::getenv("AQS_DENSE_TC");
        const int min_k = e ? (std::atoi(e) > 0 ? std::atoi(e) : 99) : 5;      // AQS_DENSE_TC=0: off; =k: from k qubits on (k = 4: 6.8 ms against 4.4 on the SIMT kernel)
        const int pad = 6 - k;
        if (k >= min_k && n - k - __builtin_popcountll(cm) >= pad + 6) {
            DenseTcArgs T;
            std::memset(&T, 0, sizeof T);
            T.a = s->d;
            T.ctrl_or = cv;
            for (int i = 0; i < k; ++i) T.toff[i] = A.toff[i];
            uint64_t allmask = fixedmask | cm;
            for (int i = k, b = 0; i < 6; ++b)
                if (!(allmask >> b & 1ull)) { T.toff[i++] = 1ull << b; allmask |= 1ull << b; }
            bitlist_from_mask(allmask, T.fixed);
            T.n_groups = s->N >> T.fixed.n;
            // W' = I (x) M, row-major 64 x 64 complex
            std::vector<float> w((size_t)64 * 64 * 2, 0.f);
            for (int sp = 0; sp < (1 << pad); ++sp)
                for (int r = 0; r < D; ++r)
                    for (int c = 0; c < D; ++c) {
                        const size_t idx = ((size_t)((sp << k) | r) * 64 + (size_t)((sp << k) | c)) * 2;
                        w[idx] = m[r * D + c].re;
                        w[idx + 1] = m[r * D + c].im;
                    }
            const size_t bytes = w.size() * sizeof(float);
            int rc = ensure_scratch(s, bytes);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(s->scratch, w.data(), bytes, cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            count_h2d(bytes);
            T.m = (const float2*)s->scratch;
            static bool opted = false;
            if (!opted) {
                CUDA_TRY(cudaFuncSetAttribute(k_dense_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
                opted = true;
            }
            const uint64_t tiles = T.n_groups / kTcGroups;
            const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)g_sm_count);
            k_dense_tc<<<grid, kTcThreads, kTcSmemBytes, s->stream>>>(T);
            cudaError_t e2 = cudaGetLastError();
            if (e2 != cudaSuccess) return fail_cuda(e2, "tensor-core dense kernel launch", __LINE__);
            count_launch(1);
            count_ops(1);
            return AQS_OK;
        }
    }
    // matrix -> device as packed operand pairs {re, re, -im, im}
    std::vector<float> packed((size_t)D * D * 4);
    for (int i = 0; i < D * D; ++i) {
        packed[4 * i] = m[i].re; packed[4 * i + 1] = m[i].re;
        packed[4 * i + 2] = -m[i].im; packed[4 * i + 3] = m[i].im;
    }
    const size_t bytes = packed.size() * sizeof(float);
    int rc = ensure_scratch(s, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->scratch, packed.data(), bytes, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));          // `packed` is pageable host memory on this frame
    count_h2d(bytes);
    A.m = (const float4*)s->scratch;
    REQUIRE((A.n_groups + kDenseThreads - 1) / kDenseThreads <= 0x7fffffffull, "grid too large");
    cudaError_t e;
    switch (k) {
        case 1: e = launch_dense<1>(A, s->stream); break;
        case 2: e = launch_dense<2>(A, s->stream); break;
        case 3: e = launch_dense<3>(A, s->stream); break;
        case 4: e = launch_dense<4>(A, s->stream); break;
        case 5: e = launch_dense<5>(A, s->stream); break;
        default: e = launch_dense<6>(A, s->stream); break;
    }
    if (e != cudaSuccess) return fail_cuda(e, "dense kernel launch", __LINE__);
    count_launch(1);
    count_ops(1);
    return AQS_OK;
}

// ---- peer memory (sharded states on one NVLink / NVSwitch node) ---------------
// Opened handles are cached for the life of the engine: state buffers are pooled, so the same
// allocations (and handles) come back run after run, and cudaIpcOpenMemHandle / Close cost milliseconds.
struct IpcEntry { cudaIpcMemHandle_t h; void* p; };
static std::vector<IpcEntry> g_ipc;

int aqs_state_ipc_export(aqs_state_t s, void* handle_out) {
    REQUIRE_INIT();
    REQUIRE(s && handle_out, "null argument");
    REQUIRE(s->own_memory, "only engine-allocated states can be exported (the buffer must be the base of its allocation)");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->d));
    static_assert(sizeof(cudaIpcMemHandle_t) == AQS_IPC_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof h);
    return AQS_OK;
}

int aqs_ipc_open(const void* handle, void** peer_ptr) {
    REQUIRE_INIT();
    REQUIRE(handle && peer_ptr, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& e : g_ipc)
        if (std::memcmp(&e.h, &h, sizeof h) == 0) { *peer_ptr = e.p; return AQS_OK; }
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcOpenMemHandle", __LINE__);
    g_ipc.push_back({h, p});
    *peer_ptr = p;
    return AQS_OK;
}

int aqs_ipc_close_all(void) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& e : g_ipc) cudaIpcCloseMemHandle(e.p);
    g_ipc.clear();
    return AQS_OK;
}

int aqs_peer_bitswap(aqs_state_t s, void* const* members, int k, const int* local_bits, uint32_t my_value) {
    REQUIRE_INIT();
    REQUIRE(s && members && local_bits, "null argument");
    REQUIRE(k >= 1 && k <= kPeerMaxK, "1 to 3 (global, local) bit pairs per remap");
    REQUIRE(my_value < (1u << k), "member value out of range");
    REQUIRE(s->n >= k + 2, "shard too small for a peer remap");
    PeerSwapArgs A;
    std::memset(&A, 0, sizeof A);
    A.mine = s->d;
    A.k = (uint32_t)k;
    A.my = my_value;
    uint64_t fixedmask = 1ull;                               // bit 0: work items are 128-bit vectors
    for (int i = 0; i < k; ++i) {
        REQUIRE(local_bits[i] >= 1 && local_bits[i] < s->n, "local bit out of range (bit 0 cannot be remapped)");
        REQUIRE(!(fixedmask >> local_bits[i] & 1ull), "duplicate local bit");
        fixedmask |= 1ull << local_bits[i];
    }
    int hbit = -1;
    for (int b = s->n - 1; b >= 1; --b)
        if (!(fixedmask >> b & 1ull)) { hbit = b; break; }
    REQUIRE(hbit >= 1, "no free local bit to split the work");
    A.hbit = (uint32_t)hbit;
    fixedmask |= 1ull << hbit;
    bitlist_from_mask(fixedmask, A.fixed);
    for (uint32_t v = 0; v < (1u << k); ++v) {
        uint64_t off = 0;
        for (int i = 0; i < k; ++i)
            if (v >> i & 1u) off |= 1ull << local_bits[i];
        A.voff[v] = off;
        A.peer[v] = (float2*)members[v];
        REQUIRE(v == my_value || members[v] != nullptr, "null member pointer");
    }
    A.n_items = s->N >> A.fixed.n;
    const uint64_t per_block = (uint64_t)kPeerThreads * kPeerItems;
    const uint64_t gx = (A.n_items + per_block - 1) / per_block;
    REQUIRE(gx * ((1ull << k) - 1ull) <= 0x7fffffffull, "grid too large");
    k_peer_bitswap<<<(unsigned)(gx * ((1ull << k) - 1ull)), kPeerThreads, 0, s->stream>>>(A);
    CUDA_TRY(cudaGetLastError());
    count_launch(1);
    return AQS_OK;
}

// ---- timers -------------------------------------------------------------------------
struct aqs_timer_s {
    cudaEvent_t a, b;
};
int aqs_timer_create(aqs_timer_t* out) {
    REQUIRE_INIT();
    REQUIRE(out, "null output");
    aqs_timer_s* t = new (std::nothrow) aqs_timer_s();
    if (!t) return fail(AQS_ERR_NOMEM, "host allocation failed");
    CUDA_TRY(cudaEventCreate(&t->a));
    CUDA_TRY(cudaEventCreate(&t->b));
    *out = t;
    return AQS_OK;
}
int aqs_timer_start(aqs_timer_t t, aqs_state_t s) {
    REQUIRE(t && s, "null argument");
    CUDA_TRY(cudaEventRecord(t->a, s->stream));
    return AQS_OK;
}
int aqs_timer_stop(aqs_timer_t t, aqs_state_t s) {
    REQUIRE(t && s, "null argument");
    CUDA_TRY(cudaEventRecord(t->b, s->stream));
    return AQS_OK;
}
int aqs_timer_elapsed_ms(aqs_timer_t t, double* ms) {
    REQUIRE(t && ms, "null argument");
    CUDA_TRY(cudaEventSynchronize(t->b));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, t->a, t->b));
    *ms = f;
    return AQS_OK;
}
int aqs_timer_destroy(aqs_timer_t t) {
    if (!t) return AQS_OK;
    cudaEventDestroy(t->a);
    cudaEventDestroy(t->b);
    delete t;
    return AQS_OK;
}

int aqs_counters_get(aqs_counters* out) {
    REQUIRE(out, "null output");
    out->kernel_launches = c_launches.load();
    out->gate_ops = c_ops.load();
    out->h2d_bytes = c_h2d.load();
    out->d2h_bytes = c_d2h.load();
    return AQS_OK;
}
int aqs_counters_reset(void) {
    c_launches = 0; c_ops = 0; c_h2d = 0; c_d2h = 0;
    return AQS_OK;
}

}  // extern "C"

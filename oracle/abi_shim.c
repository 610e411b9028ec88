/*
 * abi_shim.c — CPU stand-in for libaqs_engine.so.   TEST INFRASTRUCTURE ONLY.
 *
 * Implements every entry point of include/aqs_engine.h on top of the oracle
 * (aqs_oracle.c), so that the C++ host layer (afquantumsim_b200/host) — circuit
 * construction, flattening of composite gates, measurement logic, the string
 * grammar — can be exercised by `-m "not gpu"` tests in a container without a
 * GPU.  It is built into oracle/_build/cpu_abi/libaqs_engine.so and is loaded
 * ONLY by tests (via LD_LIBRARY_PATH / an explicit preload).  The product never
 * loads it: afquantumsim_b200 always dlopens afquantumsim_b200/lib/libaqs_engine.so
 * by absolute path, and that library has no CPU path.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "aqs_engine.h"

typedef struct { float re, im; } c32;
int orc_apply_prim(c32* a, int n, int kind, int p, int p2, uint64_t cmask, uint64_t cval, const c32* m);
void orc_set_basis(c32* a, int n, uint64_t idx);
void orc_set_product(c32* a, int n, const c32* q);
void orc_probabilities(const c32* a, int n, float* out);
uint64_t orc_prob_fixed(const c32* a, int n, uint64_t mask, uint64_t value);
double orc_norm2(const c32* a, int n);
void orc_collapse_qubit(c32* a, int n, int qubit, int outcome, float p);
int orc_sample(const c32* a, int n, const float* u, uint64_t draws, uint64_t* out, int mode);

struct aqs_state_s { int n; uint64_t N; c32* a; int own; };
int orc_sample_fixed(const c32* a, int n, const uint64_t* U, uint64_t draws, uint64_t* out);
struct aqs_plan_s { int n; uint64_t n_ops; aqs_op* ops; double bytes; };
struct aqs_timer_s { struct timespec a, b; };

static __thread char g_err[256];
static int g_init = 0;
static aqs_counters g_cnt;

static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
#define REQ(c, msg) do { if (!(c)) return fail(AQS_ERR_INVALID, msg); } while (0)

const char* aqs_last_error(void) { return g_err; }
int aqs_engine_abi_version(void) { return AQS_ENGINE_ABI_VERSION; }
int aqs_engine_init(int device) { (void)device; g_init = 1; return AQS_OK; }
int aqs_engine_shutdown(void) { g_init = 0; return AQS_OK; }
int aqs_engine_device(int* d, int* sm, size_t* mem) { if (d) *d = -1; if (sm) *sm = 0; if (mem) *mem = 0; return AQS_OK; }

int aqs_state_create(int n, aqs_state_t* out) {
    if (!g_init) return fail(AQS_ERR_STATE, "aqs_engine_init has not been called");
    REQ(out, "null output handle");
    REQ(n >= 1 && n <= 30, "qubit count must be in [1, 30] (cpu shim)");
    aqs_state_t s = (aqs_state_t)calloc(1, sizeof *s);
    s->n = n; s->N = 1ULL << n;
    s->a = (c32*)calloc(s->N, sizeof(c32));
    if (!s->a) { free(s); return fail(AQS_ERR_NOMEM, "allocation failed"); }
    s->a[0].re = 1.f;
    s->own = 1;
    *out = s;
    return AQS_OK;
}
int aqs_state_wrap(int n, void* ptr, aqs_state_t* out) {
    if (!g_init) return fail(AQS_ERR_STATE, "aqs_engine_init has not been called");
    REQ(out && ptr, "null argument"); REQ(n >= 1 && n <= 30, "qubit count must be in [1, 30] (cpu shim)");
    aqs_state_t s = (aqs_state_t)calloc(1, sizeof *s);
    s->n = n; s->N = 1ULL << n; s->a = (c32*)ptr; s->own = 0;
    *out = s;
    return AQS_OK;
}
int aqs_state_destroy(aqs_state_t s) { if (s) { if (s->own) free(s->a); free(s); } return AQS_OK; }
int aqs_state_clone(aqs_state_t src, aqs_state_t* out) {
    REQ(src && out, "null handle");
    int rc = aqs_state_create(src->n, out);
    if (rc) return rc;
    memcpy((*out)->a, src->a, src->N * sizeof(c32));
    return AQS_OK;
}
int aqs_state_qubits(aqs_state_t s, int* n) { REQ(s && n, "null"); *n = s->n; return AQS_OK; }
int aqs_state_set_basis(aqs_state_t s, uint64_t idx) { REQ(s, "null handle"); REQ(idx < s->N, "basis index out of range"); orc_set_basis(s->a, s->n, idx); return AQS_OK; }
int aqs_state_set_product(aqs_state_t s, const aqs_c32* q) { REQ(s && q, "null"); orc_set_product(s->a, s->n, (const c32*)q); return AQS_OK; }
int aqs_state_set_identity(aqs_state_t s) {
    REQ(s, "null handle"); REQ((s->n & 1) == 0, "identity needs an even qubit count (2m)");
    memset(s->a, 0, s->N * sizeof(c32));
    uint64_t M = 1ULL << (s->n / 2);
    for (uint64_t c = 0; c < M; ++c) s->a[c * M + c].re = 1.f;
    return AQS_OK;
}
int aqs_state_upload(aqs_state_t s, const aqs_c32* h, uint64_t off, uint64_t cnt) {
    REQ(s && h, "null"); REQ(off <= s->N && cnt <= s->N - off, "range outside the state");
    memcpy(s->a + off, h, cnt * sizeof(c32)); g_cnt.h2d_bytes += cnt * 8; return AQS_OK;
}
int aqs_state_download(aqs_state_t s, aqs_c32* h, uint64_t off, uint64_t cnt) {
    REQ(s && h, "null"); REQ(off <= s->N && cnt <= s->N - off, "range outside the state");
    memcpy(h, s->a + off, cnt * sizeof(c32)); g_cnt.d2h_bytes += cnt * 8; return AQS_OK;
}
int aqs_state_get_amp(aqs_state_t s, uint64_t i, aqs_c32* out) { return aqs_state_download(s, out, i, 1); }
int aqs_state_device_ptr(aqs_state_t s, void** p) { REQ(s && p, "null"); *p = s->a; return AQS_OK; }
int aqs_state_set_stream(aqs_state_t s, void* st) { (void)s; (void)st; return AQS_OK; }
int aqs_state_get_stream(aqs_state_t s, void** st) { (void)s; if (st) *st = NULL; return AQS_OK; }
int aqs_sync(aqs_state_t s) { (void)s; return AQS_OK; }

static uint64_t to_pos(int n, uint64_t qmask) {
    uint64_t m = 0;
    for (int q = 0; q < n; ++q) if (qmask >> q & 1ULL) m |= 1ULL << (n - 1 - q);
    return m;
}
static int validate(int n, const aqs_op* op) {
    REQ(op->kind >= AQS_OP_U2 && op->kind <= AQS_OP_SWAP, "unknown op kind");
    REQ(op->target >= 0 && op->target < n, "target qubit out of range");
    REQ((op->ctrl_mask >> n) == 0, "control mask names a qubit outside the state");
    REQ(!(op->ctrl_mask >> op->target & 1ULL), "control qubit cannot be the target qubit");
    if (op->kind == AQS_OP_SWAP) {
        REQ(op->target2 >= 0 && op->target2 < n, "second swap qubit out of range");
        REQ(op->target2 != op->target, "cannot swap a qubit with itself");
        REQ(!(op->ctrl_mask >> op->target2 & 1ULL), "control qubit cannot be a swap target");
    }
    return AQS_OK;
}
int aqs_apply_op(aqs_state_t s, const aqs_op* op) {
    REQ(s && op, "null argument");
    int rc = validate(s->n, op);
    if (rc) return rc;
    int p2 = op->kind == AQS_OP_SWAP ? s->n - 1 - op->target2 : -1;
    orc_apply_prim(s->a, s->n, op->kind, s->n - 1 - op->target, p2, to_pos(s->n, op->ctrl_mask),
                   to_pos(s->n, op->ctrl_value & op->ctrl_mask), (const c32*)op->m);
    g_cnt.gate_ops++; g_cnt.kernel_launches++;
    return AQS_OK;
}
int aqs_apply_ops(aqs_state_t s, const aqs_op* ops, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) { int rc = aqs_apply_op(s, ops + i); if (rc) return rc; }
    return AQS_OK;
}

int aqs_plan_build(int n, const aqs_op* ops, uint64_t n_ops, uint32_t flags, aqs_plan_t* out) {
    (void)flags;
    REQ(out, "null output handle"); REQ(n >= 1 && n <= AQS_MAX_QUBITS, "qubit count out of range");
    for (uint64_t i = 0; i < n_ops; ++i) { int rc = validate(n, ops + i); if (rc) return rc; }
    aqs_plan_t p = (aqs_plan_t)calloc(1, sizeof *p);
    p->n = n; p->n_ops = n_ops;
    p->ops = (aqs_op*)malloc(sizeof(aqs_op) * (n_ops ? n_ops : 1));
    if (n_ops) memcpy(p->ops, ops, sizeof(aqs_op) * n_ops);
    *out = p;
    return AQS_OK;
}
int aqs_plan_run(aqs_state_t s, aqs_plan_t p) {
    REQ(s && p, "null handle"); REQ(s->n == p->n, "plan and state have different qubit counts");
    return aqs_apply_ops(s, p->ops, p->n_ops);
}
int aqs_plan_get_info(aqs_plan_t p, aqs_plan_info* info) {
    REQ(p && info, "null"); memset(info, 0, sizeof *info);
    info->n_ops = info->n_launches = info->n_single_ops = p->n_ops; info->n_qubits = p->n;
    return AQS_OK;
}
int aqs_plan_destroy(aqs_plan_t p) { if (p) { free(p->ops); free(p); } return AQS_OK; }
int aqs_plan_export_pass(aqs_plan_t p, uint64_t index, void* buf, uint64_t cap, uint64_t* needed) {
    (void)p; (void)index; (void)buf; (void)cap; (void)needed;
    return fail(AQS_ERR_INVALID, "the CPU stand-in has no fused passes");
}

int aqs_norm2(aqs_state_t s, double* out) { REQ(s && out, "null"); *out = orc_norm2(s->a, s->n); return AQS_OK; }
int aqs_scale(aqs_state_t s, float f) { REQ(s, "null"); for (uint64_t r = 0; r < s->N; ++r) { s->a[r].re *= f; s->a[r].im *= f; } return AQS_OK; }
int aqs_prob_fixed(aqs_state_t s, uint64_t qm, uint64_t qv, uint64_t* out) {
    REQ(s && out, "null"); REQ((qm >> s->n) == 0, "mask names a qubit outside the state");
    *out = orc_prob_fixed(s->a, s->n, to_pos(s->n, qm), to_pos(s->n, qv & qm)); return AQS_OK;
}
int aqs_qubit_prob1(aqs_state_t s, int q, double* out) {
    REQ(s && out, "null"); REQ(q >= 0 && q < s->n, "qubit out of range");
    uint64_t f; aqs_prob_fixed(s, 1ULL << q, 1ULL << q, &f); *out = (double)f * 0x1p-62; return AQS_OK;
}
int aqs_probabilities(aqs_state_t s, float* out, uint64_t off, uint64_t cnt) {
    REQ(s && out, "null"); REQ(off <= s->N && cnt <= s->N - off, "range outside the state");
    float* tmp = (float*)malloc(sizeof(float) * s->N);
    orc_probabilities(s->a, s->n, tmp); memcpy(out, tmp + off, cnt * sizeof(float)); free(tmp); return AQS_OK;
}
int aqs_collapse_qubit(aqs_state_t s, int q, int outcome, float p) {
    REQ(s, "null"); REQ(q >= 0 && q < s->n, "qubit out of range"); REQ(outcome == 0 || outcome == 1, "outcome must be 0 or 1");
    orc_collapse_qubit(s->a, s->n, q, outcome, p); return AQS_OK;
}
int aqs_sample(aqs_state_t s, const float* u, uint64_t n, uint64_t* out) {
    REQ(s && (out || n == 0), "null"); if (n == 0) return AQS_OK;
    return orc_sample(s->a, s->n, u, n, out, 0) ? fail(AQS_ERR_NOMEM, "sample") : AQS_OK;
}
int aqs_sample_fixed(aqs_state_t s, const uint64_t* u, uint64_t n, uint64_t* out) {
    REQ(s && ((out && u) || n == 0), "null"); if (n == 0) return AQS_OK;
    return orc_sample_fixed(s->a, s->n, u, n, out) ? fail(AQS_ERR_NOMEM, "sample") : AQS_OK;
}
int aqs_sample_hist(aqs_state_t s, const float* u, uint64_t n, uint32_t* hist) {
    REQ(s && hist, "null"); memset(hist, 0, s->N * sizeof(uint32_t)); if (n == 0) return AQS_OK;
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * n);
    if (orc_sample(s->a, s->n, u, n, idx, 0)) { free(idx); return fail(AQS_ERR_NOMEM, "sample"); }
    for (uint64_t i = 0; i < n; ++i) hist[idx[i]]++;
    free(idx); return AQS_OK;
}

static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
int aqs_sample_hist_sparse(aqs_state_t s, const float* u, uint64_t n, uint64_t* index, uint32_t* count, uint64_t cap, uint64_t* n_bins) {
    REQ(s && n_bins && (u || n == 0), "null"); *n_bins = 0; if (n == 0) return AQS_OK;
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * n);
    if (orc_sample(s->a, s->n, u, n, idx, 0)) { free(idx); return fail(AQS_ERR_NOMEM, "sample"); }
    qsort(idx, n, sizeof(uint64_t), cmp_u64);
    uint64_t bins = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i; while (j < n && idx[j] == idx[i]) ++j;
        if (index && count && bins < cap) { index[bins] = idx[i]; count[bins] = (uint32_t)(j - i); }
        ++bins; i = j;
    }
    free(idx); *n_bins = bins;
    if (index && count && bins > cap) return fail(AQS_ERR_INVALID, "histogram has more bins than the output arrays hold");
    return AQS_OK;
}
int aqs_pool_trim(void) { return AQS_OK; }
int aqs_jit_wait(void) { return AQS_OK; }
/* specialised pass kernels exist only in the CUDA engine: the stand-in has no fused passes to specialise */
int aqs_plan_pass_source(aqs_plan_t p, uint64_t index, char* buf, uint64_t cap, uint64_t* needed, uint64_t* geom) {
    (void)p; (void)index; (void)buf; (void)cap; (void)needed; (void)geom; return fail(AQS_ERR_STATE, "no fused passes in the CPU stand-in");
}
int aqs_plan_pass_coefs(aqs_plan_t p, uint64_t index, uint64_t* buf, uint64_t cap, uint64_t* needed) {
    (void)p; (void)index; (void)buf; (void)cap; (void)needed; return fail(AQS_ERR_STATE, "no fused passes in the CPU stand-in");
}
int aqs_plan_jit_ready(aqs_plan_t p, uint64_t* n_ready) { REQ(p && n_ready, "null"); *n_ready = 0; return AQS_OK; }
int aqs_jit_get_info(aqs_jit_info* out) { REQ(out, "null"); memset(out, 0, sizeof *out); return AQS_OK; }

/* opaque k-qubit matrix on arbitrary qubits: plain loops (the oracle's orc_apply_dense restates the
 * reference's contiguous Gate / ControlGate embedding; tests compare the two) */
int aqs_apply_dense(aqs_state_t s, const int* qubits, int k, uint64_t ctrl_mask, uint64_t ctrl_value, const aqs_c32* m) {
    REQ(s && qubits && m, "null argument");
    REQ(k >= 1 && k <= 6 && k <= s->n, "dense blocks of 1 to 6 qubits");
    const int n = s->n, D = 1 << k;
    uint64_t tmask = 0, toff[6], cm = 0, cv = 0;
    for (int i = 0; i < k; ++i) {
        REQ(qubits[i] >= 0 && qubits[i] < n, "target qubit out of range");
        const uint64_t b = 1ULL << (n - 1 - qubits[i]);
        REQ(!(tmask & b), "duplicate target qubit");
        tmask |= b;
        toff[k - 1 - i] = b;
    }
    for (int q = 0; q < n; ++q)
        if ((ctrl_mask >> q) & 1ULL) { cm |= 1ULL << (n - 1 - q); if ((ctrl_value >> q) & 1ULL) cv |= 1ULL << (n - 1 - q); }
    REQ(!(cm & tmask), "a control qubit is also a target");
    c32 in[64];
    for (uint64_t x = 0; x < s->N; ++x) {
        if ((x & tmask) || (x & cm) != cv) continue;
        for (int c = 0; c < D; ++c) {
            uint64_t off = 0;
            for (int i = 0; i < k; ++i) if ((c >> i) & 1) off |= toff[i];
            in[c] = s->a[x | off];
        }
        for (int r = 0; r < D; ++r) {
            float re = 0.f, im = 0.f;
            for (int c = 0; c < D; ++c) {
                const aqs_c32 e = m[r * D + c];
                re += e.re * in[c].re - e.im * in[c].im;
                im += e.re * in[c].im + e.im * in[c].re;
            }
            uint64_t off = 0;
            for (int i = 0; i < k; ++i) if ((r >> i) & 1) off |= toff[i];
            s->a[x | off].re = re; s->a[x | off].im = im;
        }
    }
    g_cnt.gate_ops += 1;
    return AQS_OK;
}

/* peer memory: there is none across CPU processes.  In-process "members" (plain host pointers)
 * are supported so that tests can check the remap's index arithmetic against a numpy restatement:
 * member `my` performs every swap it shares with a member of higher value. */
int aqs_state_ipc_export(aqs_state_t s, void* h) { (void)s; (void)h; return fail(AQS_ERR_STATE, "no peer memory on the cpu shim"); }
int aqs_ipc_open(const void* h, void** p) { (void)h; (void)p; return fail(AQS_ERR_STATE, "no peer memory on the cpu shim"); }
int aqs_ipc_close_all(void) { return AQS_OK; }
int aqs_peer_bitswap(aqs_state_t s, void* const* members, int k, const int* local_bits, uint32_t my) {
    REQ(s && members && local_bits, "null argument");
    REQ(k >= 1 && k <= 3 && my < (1u << k), "bad remap");
    uint64_t sel = 0;
    for (int i = 0; i < k; ++i) { REQ(local_bits[i] >= 1 && local_bits[i] < s->n, "local bit out of range"); sel |= 1ULL << local_bits[i]; }
    for (uint64_t x = 0; x < s->N; ++x) {
        uint32_t v = 0;
        for (int i = 0; i < k; ++i) v |= (uint32_t)((x >> local_bits[i]) & 1ULL) << i;
        if (v <= my) continue;
        uint64_t y = x & ~sel;
        for (int i = 0; i < k; ++i) y |= (uint64_t)((my >> i) & 1u) << local_bits[i];
        c32* other = (c32*)members[v];
        c32 t = s->a[x]; s->a[x] = other[y]; other[y] = t;
    }
    return AQS_OK;
}

/* flat multi-GPU address space: CUDA virtual memory management has no CPU counterpart */
int aqs_flat_create(uint64_t b, int w, int r, aqs_flat_t* o, int* fd) { (void)b; (void)w; (void)r; (void)o; (void)fd; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_flat_attach(aqs_flat_t f, int r, int fd) { (void)f; (void)r; (void)fd; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_flat_ptr(aqs_flat_t f, void** b, void** o) { (void)f; (void)b; (void)o; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_flat_destroy(aqs_flat_t f) { (void)f; return AQS_OK; }
int aqs_plan_pass_tile(aqs_plan_t p, uint64_t i, uint8_t* pos, int* t) { (void)p; (void)i; (void)pos; (void)t; return fail(AQS_ERR_STATE, "no fused passes on the cpu shim"); }
int aqs_flat_view_create(aqs_flat_t f, const aqs_flat_block* b, uint64_t n, void** v) { (void)f; (void)b; (void)n; (void)v; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_memcpy_async(void* d, const void* s, uint64_t b, void* st) { (void)d; (void)s; (void)b; (void)st; return fail(AQS_ERR_STATE, "no device memory on the cpu shim"); }
int aqs_plan_run_tiles(aqs_state_t s, aqs_plan_t p, uint64_t i, const void* l, uint32_t n, const uint8_t* fp, uint32_t fo, void* st) {
    (void)s; (void)p; (void)i; (void)l; (void)n; (void)fp; (void)fo; (void)st; return fail(AQS_ERR_STATE, "no fused passes on the cpu shim");
}
int aqs_plan_run_shard(aqs_state_t s, aqs_plan_t p, uint64_t a, uint64_t c, int r, int g) { (void)s; (void)p; (void)a; (void)c; (void)r; (void)g; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_plan_shard_cut(aqs_plan_t p, uint64_t i, int r, int g, uint32_t* n, uint32_t* v, uint8_t* pos) { (void)p; (void)i; (void)r; (void)g; (void)n; (void)v; (void)pos; return fail(AQS_ERR_STATE, "no flat address space on the cpu shim"); }
int aqs_plan_pass_span(aqs_plan_t p, uint64_t i, int g, int* out) { (void)p; (void)i; (void)g; if (out) *out = 0; return AQS_OK; }

int aqs_timer_create(aqs_timer_t* out) { REQ(out, "null"); *out = (aqs_timer_t)calloc(1, sizeof **out); return AQS_OK; }
int aqs_timer_start(aqs_timer_t t, aqs_state_t s) { (void)s; clock_gettime(CLOCK_MONOTONIC, &t->a); return AQS_OK; }
int aqs_timer_stop(aqs_timer_t t, aqs_state_t s) { (void)s; clock_gettime(CLOCK_MONOTONIC, &t->b); return AQS_OK; }
int aqs_timer_elapsed_ms(aqs_timer_t t, double* ms) { *ms = (t->b.tv_sec - t->a.tv_sec) * 1e3 + (t->b.tv_nsec - t->a.tv_nsec) * 1e-6; return AQS_OK; }
int aqs_timer_destroy(aqs_timer_t t) { free(t); return AQS_OK; }
int aqs_counters_get(aqs_counters* out) { REQ(out, "null"); *out = g_cnt; return AQS_OK; }
int aqs_counters_reset(void) { memset(&g_cnt, 0, sizeof g_cnt); return AQS_OK; }

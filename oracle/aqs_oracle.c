/*
 * aqs_oracle.c — CPU restatement of afQuantumSim's state-vector hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  Nothing under afquantumsim_b200/ links, imports or calls it.
 *
 * The reference (arrayfire/afQuantumSim 1.0.0) builds an explicit 2^n x 2^n
 * operator per gate with ArrayFire and multiplies it into the state
 * (src/quantum.cpp:277-291).  ArrayFire is not installable here, so the
 * reference cannot be compiled; this file restates, gate by gate, what those
 * operators do to a 2^n complex64 state, in place and matrix-free.  Every
 * function cites the reference lines it follows.  The restatement is pinned
 * against the known-answer amplitudes of the reference's own test suite
 * (test/tests.cpp, transcribed in tests/golden/reference_kat.json).
 * Not pinned by any reference test: RNG streams and the f32 rounding order of
 * ArrayFire's accum/sum ("parity unpinned" items).
 *
 * Conventions (src/quantum.cpp:546, include/quantum.h:101-111):
 *   - amplitudes are complex64, interleaved (re, im);
 *   - qubit 0 is the MOST significant index bit: qubit q <-> bit n-1-q.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 * -ffp-contract=off matters: the sampling contract computes |a|^2 as
 * fl(fl(re*re)+fl(im*im)) with no FMA, and the CUDA side does the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } c32;

/* Gate type ids: same order as the reference enum QGate::GateTypes
 * (include/quantum.h:799-823). */
enum {
    G_BARRIER = 0, G_X, G_Y, G_Z, G_H, G_PHASE, G_SWAP, G_ROTX, G_ROTY, G_ROTZ,
    G_CX, G_CY, G_CZ, G_CH, G_CPHASE, G_CSWAP, G_CROTX, G_CROTY, G_CROTZ,
    G_CCX, G_OR, G_CIRCUIT, G_CTRL_CIRCUIT
};

static inline c32 cmul(c32 a, c32 b) {
    c32 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline c32 cadd(c32 a, c32 b) { c32 r = {a.re + b.re, a.im + b.im}; return r; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* insert a zero bit at position p of j */
static inline uint64_t ins0(uint64_t j, int p) {
    uint64_t lo = j & ((1ULL << p) - 1ULL);
    return ((j >> p) << (p + 1)) | lo;
}

/* ---------------------------------------------------------------------------
 * Generic "controlled 2x2 on one index bit":   for every pair (r0, r1=r0|m)
 * with all bits of cmask set in r0:  (a[r0], a[r1]) <- M (a[r0], a[r1]).
 * This is what I_L (x) U (x) I_R (src/quantum.cpp:676-696, 817-829 via
 * src/utils.cpp:169-195) and the dense ControlGate embedding
 * (src/quantum.cpp:1902-1947) do to a state vector.
 * ------------------------------------------------------------------------- */
static void apply_2x2(c32* a, int n, int p, uint64_t cmask, const c32 m[4]) {
    const uint64_t half = 1ULL << (n - 1);
    const uint64_t tm   = 1ULL << p;
    const c32 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)half; ++j) {
        uint64_t r0 = ins0((uint64_t)j, p);
        if ((r0 & cmask) != cmask) continue;
        uint64_t r1 = r0 | tm;
        c32 x = a[r0], y = a[r1];
        a[r0] = cadd(cmul(m00, x), cmul(m01, y));
        a[r1] = cadd(cmul(m10, x), cmul(m11, y));
    }
}

/* out[r] = in[r ^ tm] where cond(r) (a permutation that flips one bit under a
 * predicate on the OTHER bits): X :546-559, CX :992-1011, CCNot :1558-1577,
 * Or :1630-1650.  any_mode: 0 => all bits of cmask set; 1 => any bit set. */
static void apply_flip(c32* a, int n, int p, uint64_t cmask, int any_mode) {
    const uint64_t half = 1ULL << (n - 1);
    const uint64_t tm   = 1ULL << p;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)half; ++j) {
        uint64_t r0 = ins0((uint64_t)j, p);
        int on = any_mode ? ((r0 & cmask) != 0) : ((r0 & cmask) == cmask);
        if (!on) continue;
        uint64_t r1 = r0 | tm;
        c32 t = a[r0]; a[r0] = a[r1]; a[r1] = t;
    }
}

/* out[r] = d(r) * in[r]; d = d1 where all bits of (cmask|tm) are set, else
 * (d0 where the cmask bits are set and the tm bit is clear), else 1.
 * Z :633-650, Phase :855-874, CZ :1128-1149, CPhase :1192-1215, RotZ :777-787 */
static void apply_diag(c32* a, int n, int p, uint64_t cmask, c32 d0, c32 d1, int d0_is_one) {
    const uint64_t N  = 1ULL << n;
    const uint64_t tm = 1ULL << p;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) {
        if (((uint64_t)r & cmask) != cmask) continue;
        if ((uint64_t)r & tm) a[r] = cmul(d1, a[r]);
        else if (!d0_is_one)  a[r] = cmul(d0, a[r]);
    }
}

/* Swap :925-949 / CSwap :1284-1310: out[r] = in[r with bits pa,pb exchanged]
 * where the control bits are set. */
static void apply_swap(c32* a, int n, int pa, int pb, uint64_t cmask) {
    const uint64_t N = 1ULL << n;
    const uint64_t ma = 1ULL << pa, mb = 1ULL << pb;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) {
        uint64_t u = (uint64_t)r;
        if ((u & cmask) != cmask) continue;
        /* visit each (01,10) pair once: from the member with bit pa set, pb clear */
        if ((u & ma) && !(u & mb)) {
            uint64_t v = (u ^ ma) | mb;
            c32 t = a[u]; a[u] = a[v]; a[v] = t;
        }
    }
}

/* ---------------------------------------------------------------------------
 * orc_apply_gate: one reference gate object applied to the state.
 *   type   : GateTypes id
 *   q      : qubit arguments in the reference constructor order
 *            (controls first, then targets), API numbering (0 = MSB)
 *   angle  : rotation / phase angle where the gate has one
 *   xctrl  : extra control mask in BIT-POSITION space (accumulated while
 *            flattening ControlGate nests, src/quantum.cpp:1929-1936);
 *            the gate acts only where all those bits are 1.  0 for none.
 * Returns 0, or -1 for an unknown type.
 * ------------------------------------------------------------------------- */
int orc_apply_gate(c32* a, int n, int type, const int* q, float angle, uint64_t xctrl) {
#define POS(k) (n - 1 - q[k])
#define MSK(k) (1ULL << POS(k))
    const float h = 0.70710678118f; /* src/quantum.cpp:46-49 */
    switch (type) {
    case G_BARRIER: return 0;                                   /* quantum.h:857-876 */
    case G_X:  apply_flip(a, n, POS(0), xctrl, 0); return 0;    /* :546-559 */
    case G_Y: {                                                 /* :586-606  Y=[[0,-i],[i,0]] */
        c32 m[4] = {{0, 0}, {0, -1.f}, {0, 1.f}, {0, 0}};
        apply_2x2(a, n, POS(0), xctrl, m); return 0;
    }
    case G_Z: {                                                 /* :633-650 */
        c32 one = {1.f, 0}, neg = {-1.f, 0};
        apply_diag(a, n, POS(0), xctrl, one, neg, 1); return 0;
    }
    case G_H: {                                                 /* :817-829 */
        c32 m[4] = {{h, 0}, {h, 0}, {h, 0}, {-h, 0}};
        apply_2x2(a, n, POS(0), xctrl, m); return 0;
    }
    case G_PHASE: {                                             /* :855-874 */
        c32 one = {1.f, 0}, d = {cosf(angle), sinf(angle)};
        apply_diag(a, n, POS(0), xctrl, one, d, 1); return 0;
    }
    case G_SWAP: apply_swap(a, n, POS(0), POS(1), xctrl); return 0;   /* :925-949 */
    case G_ROTX: {                                              /* :683-693 */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 m[4] = {{c, 0}, {0, -s}, {0, -s}, {c, 0}};
        apply_2x2(a, n, POS(0), xctrl, m); return 0;
    }
    case G_ROTY: {                                              /* :730-740 (column-major host array, then .T()) */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 m[4] = {{c, 0}, {-s, 0}, {s, 0}, {c, 0}};
        apply_2x2(a, n, POS(0), xctrl, m); return 0;
    }
    case G_ROTZ: {                                              /* :777-787 */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 d0 = {c, -s}, d1 = {c, s};
        apply_diag(a, n, POS(0), xctrl, d0, d1, 0); return 0;
    }
    case G_CX: apply_flip(a, n, POS(1), xctrl | MSK(0), 0); return 0;  /* :992-1011 */
    case G_CY: {                                                /* :1054-1085 */
        c32 m[4] = {{0, 0}, {0, -1.f}, {0, 1.f}, {0, 0}};
        apply_2x2(a, n, POS(1), xctrl | MSK(0), m); return 0;
    }
    case G_CZ: {                                                /* :1128-1149 */
        c32 one = {1.f, 0}, neg = {-1.f, 0};
        apply_diag(a, n, POS(1), xctrl | MSK(0), one, neg, 1); return 0;
    }
    case G_CH: {                                                /* :1355-1397, sqrt2 = 0.70710678118f */
        c32 m[4] = {{h, 0}, {h, 0}, {h, 0}, {-h, 0}};
        apply_2x2(a, n, POS(1), xctrl | MSK(0), m); return 0;
    }
    case G_CPHASE: {                                            /* :1192-1215 */
        c32 one = {1.f, 0}, d = {cosf(angle), sinf(angle)};
        apply_diag(a, n, POS(1), xctrl | MSK(0), one, d, 1); return 0;
    }
    case G_CSWAP: apply_swap(a, n, POS(1), POS(2), xctrl | MSK(0)); return 0;  /* :1284-1310 */
    case G_CROTX: {                                             /* :1433 ControlGate(RotX::gate) */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 m[4] = {{c, 0}, {0, -s}, {0, -s}, {c, 0}};
        apply_2x2(a, n, POS(1), xctrl | MSK(0), m); return 0;
    }
    case G_CROTY: {                                             /* :1470 */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 m[4] = {{c, 0}, {-s, 0}, {s, 0}, {c, 0}};
        apply_2x2(a, n, POS(1), xctrl | MSK(0), m); return 0;
    }
    case G_CROTZ: {                                             /* :1507 */
        float c = cosf(angle / 2.0f), s = sinf(angle / 2.0f);
        c32 d0 = {c, -s}, d1 = {c, s};
        apply_diag(a, n, POS(1), xctrl | MSK(0), d0, d1, 0); return 0;
    }
    case G_CCX: apply_flip(a, n, POS(2), xctrl | MSK(0) | MSK(1), 0); return 0;  /* :1558-1577 */
    case G_OR: {                                                /* :1630-1650: flip where (a OR b) */
        if (xctrl == 0) { apply_flip(a, n, POS(2), MSK(0) | MSK(1), 1); return 0; }
        /* under extra controls: t ^= a ^ b ^ ab, each under xctrl */
        apply_flip(a, n, POS(2), xctrl | MSK(0), 0);
        apply_flip(a, n, POS(2), xctrl | MSK(1), 0);
        apply_flip(a, n, POS(2), xctrl | MSK(0) | MSK(1), 0);
        return 0;
    }
    default: return -1;
    }
#undef POS
#undef MSK
}

/* ---------------------------------------------------------------------------
 * Dense embedding, faithful to Gate::operator() (src/quantum.cpp:1760-1814)
 * and ControlGate::operator() (:1888-1950, bit placement src/utils.cpp:137-167):
 * a k-qubit matrix U (column-major 2^k x 2^k, as af::array stores it) acts on
 * contiguous qubits [begin, begin+k); inner qubit j <-> outer qubit begin+j;
 * optional control qubit ctrl (-1 = none) must be 1.  xctrl as above.
 * ------------------------------------------------------------------------- */
int orc_apply_dense(c32* a, int n, int k, int begin, int ctrl, const c32* U, uint64_t xctrl) {
    if (k < 1 || begin < 0 || begin + k > n) return -1;
    const int sh = n - begin - k;                 /* lowest bit of the block */
    const uint64_t K = 1ULL << k, N = 1ULL << n;
    uint64_t cmask = xctrl;
    if (ctrl >= 0) {
        if (ctrl >= begin && ctrl < begin + k) return -1;
        cmask |= 1ULL << (n - 1 - ctrl);
    }
    const uint64_t groups = N >> k;
#pragma omp parallel
    {
        c32* in  = (c32*)malloc(sizeof(c32) * K);
        c32* out = (c32*)malloc(sizeof(c32) * K);
#pragma omp for schedule(static)
        for (int64_t g = 0; g < (int64_t)groups; ++g) {
            uint64_t lo = (uint64_t)g & ((1ULL << sh) - 1ULL);
            uint64_t hi = ((uint64_t)g >> sh) << (sh + k);
            uint64_t base = hi | lo;
            if ((base & cmask) != cmask) continue;
            for (uint64_t c = 0; c < K; ++c) in[c] = a[base | (c << sh)];
            for (uint64_t r = 0; r < K; ++r) {
                c32 acc = {0.f, 0.f};
                for (uint64_t c = 0; c < K; ++c) acc = cadd(acc, cmul(U[c * K + r], in[c]));
                out[r] = acc;
            }
            for (uint64_t r = 0; r < K; ++r) a[base | (r << sh)] = out[r];
        }
        free(in); free(out);
    }
    return 0;
}

/* ---------------------------------------------------------------------------
 * State preparation (src/quantum.cpp:212-275).
 * ------------------------------------------------------------------------- */
void orc_set_basis(c32* a, int n, uint64_t idx) {
    const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) { a[r].re = 0.f; a[r].im = 0.f; }
    a[idx].re = 1.f;
}

/* generate_statevector :261-275: kron of the per-qubit 2-vectors, multiplied
 * left to right starting from qubit 0, in complex64.  q is n x 2. */
void orc_set_product(c32* a, int n, const c32* q) {
    const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) {
        c32 v = q[0 * 2 + (((uint64_t)r >> (n - 1)) & 1ULL)];
        for (int k = 1; k < n; ++k) v = cmul(v, q[k * 2 + (((uint64_t)r >> (n - 1 - k)) & 1ULL)]);
        a[r] = v;
    }
}

/* ---------------------------------------------------------------------------
 * Probabilities and the exact-sum contract.
 *
 * The reference sums f32 probabilities with ArrayFire's accum/sum
 * (src/quantum.cpp:349-353, 388, 481); their rounding order is not visible in
 * the tree ("parity unpinned").  To make histograms bit-exact between this
 * oracle, one GPU and R GPUs in any summation order, the engine contract is:
 *     p_k  = fl32( fl32(re*re) + fl32(im*im) )           (no FMA)
 *     F_k  = trunc( p_k * 2^62 )   as uint64             (exact scaling)
 *     S_k  = sum_{j<=k} F_j        in uint64 (associative => order-free)
 *     U    = trunc( u * 2^62 )     for a draw u in [0,1)
 *     outcome = min{ k : S_k > U }, or 0 if none (peek_measure_all :353-356)
 * A second mode, seq_f32, is the literal sequential f32 inclusive scan, kept
 * to show how close the contract is to a sequential ArrayFire-CPU accum.
 * ------------------------------------------------------------------------- */
static inline float prob32(c32 v) {
    float x = v.re * v.re;
    float y = v.im * v.im;
    return x + y;
}
static inline uint64_t fix62(float p) { return (uint64_t)(p * 0x1p62f); }

void orc_probabilities(const c32* a, int n, float* out) {       /* :404-414 */
    const uint64_t N = 1ULL << n;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) out[r] = prob32(a[r]);
}

/* sum of F_k over indices with (r & mask) == value; mask = 0 -> total */
uint64_t orc_prob_fixed(const c32* a, int n, uint64_t mask, uint64_t value) {
    const uint64_t N = 1ULL << n;
    uint64_t tot = 0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t r = 0; r < (int64_t)N; ++r)
        if (((uint64_t)r & mask) == value) tot += fix62(prob32(a[r]));
    return tot;
}

/* qubit_probability_true :372-391 under the exact-sum contract */
double orc_qubit_prob1(const c32* a, int n, int qubit) {
    uint64_t m = 1ULL << (n - 1 - qubit);
    return (double)orc_prob_fixed(a, n, m, m) * 0x1p-62;
}

/* literal f32 sequential version of the same quantity (what a single-threaded
 * af::sum<float> would give) */
float orc_qubit_prob1_seq_f32(const c32* a, int n, int qubit) {
    const uint64_t N = 1ULL << n, m = 1ULL << (n - 1 - qubit);
    float s = 0.f;
    for (uint64_t r = 0; r < N; ++r) if (r & m) s += prob32(a[r]);
    return s;
}

double orc_norm2(const c32* a, int n) {
    const uint64_t N = 1ULL << n;
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t r = 0; r < (int64_t)N; ++r)
        tot += (double)a[r].re * a[r].re + (double)a[r].im * a[r].im;
    return tot;
}

/* measure :336-339: keep the half matching `outcome`, zero the other, divide
 * by sqrtf(p) (p = prob1 for outcome 1, 1.f - prob1 for outcome 0; the caller
 * passes the p to divide by). */
void orc_collapse_qubit(c32* a, int n, int qubit, int outcome, float p) {
    const uint64_t N = 1ULL << n, m = 1ULL << (n - 1 - qubit);
    const float s = sqrtf(p);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)N; ++r) {
        int bit = ((uint64_t)r & m) != 0;
        if (bit == outcome) { a[r].re = a[r].re / s; a[r].im = a[r].im / s; }
        else { a[r].re = 0.f; a[r].im = 0.f; }
    }
}

/* Sampling: peek_measure_all :344-359, profile_measure_all :467-501.
 * mode 0 = exact-sum contract, mode 1 = seq_f32.  out[i] = outcome of draw i. */
int orc_sample(const c32* a, int n, const float* u, uint64_t draws, uint64_t* out, int mode) {
    const uint64_t N = 1ULL << n;
    if (mode == 0) {
        uint64_t* S = (uint64_t*)malloc(sizeof(uint64_t) * N);
        if (!S) return -1;
        /* two-level parallel inclusive scan of F_k (integer => exact) */
        const uint64_t B = 1ULL << 16;
        const uint64_t nb = (N + B - 1) / B;
        uint64_t* bs = (uint64_t*)calloc(nb + 1, sizeof(uint64_t));
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)nb; ++b) {
            uint64_t s = 0, e = ((uint64_t)b + 1) * B; if (e > N) e = N;
            for (uint64_t r = (uint64_t)b * B; r < e; ++r) { s += fix62(prob32(a[r])); S[r] = s; }
            bs[b + 1] = s;
        }
        for (uint64_t b = 0; b < nb; ++b) bs[b + 1] += bs[b];
#pragma omp parallel for schedule(static)
        for (int64_t b = 1; b < (int64_t)nb; ++b) {
            uint64_t e = ((uint64_t)b + 1) * B; if (e > N) e = N;
            for (uint64_t r = (uint64_t)b * B; r < e; ++r) S[r] += bs[b];
        }
        free(bs);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)draws; ++i) {
            uint64_t U = fix62(u[i]);
            /* first k with S[k] > U */
            uint64_t lo = 0, hi = N;
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (S[mid] > U) hi = mid; else lo = mid + 1; }
            out[i] = (lo == N) ? 0 : lo;
        }
        free(S);
        return 0;
    } else {
        float* S = (float*)malloc(sizeof(float) * N);
        if (!S) return -1;
        float s = 0.f;
        for (uint64_t r = 0; r < N; ++r) { s += prob32(a[r]); S[r] = s; }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)draws; ++i) {
            float v = u[i];
            uint64_t lo = 0, hi = N;   /* S is non-decreasing */
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (S[mid] > v) hi = mid; else lo = mid + 1; }
            out[i] = (lo == N) ? 0 : lo;
        }
        free(S);
        return 0;
    }
}

/* Local step of sharded sampling: thresholds already in the fixed-point domain.
 * out = first local k with S_k > U, or UINT64_MAX when U >= the shard's total. */
int orc_sample_fixed(const c32* a, int n, const uint64_t* U, uint64_t draws, uint64_t* out) {
    const uint64_t N = 1ULL << n;
    uint64_t* S = (uint64_t*)malloc(sizeof(uint64_t) * N);
    if (!S) return -1;
    uint64_t s = 0;
    for (uint64_t r = 0; r < N; ++r) { s += fix62(prob32(a[r])); S[r] = s; }
    for (uint64_t i = 0; i < draws; ++i) {
        uint64_t lo = 0, hi = N;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (S[mid] > U[i]) hi = mid; else lo = mid + 1; }
        out[i] = (lo == N) ? UINT64_MAX : lo;
    }
    free(S);
    return 0;
}

/* relative L2 distance ||a-b|| / ||b||, in double — used by the parity tests */
double orc_rel_l2(const c32* a, const c32* b, uint64_t count) {
    double num = 0.0, den = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : num, den)
    for (int64_t r = 0; r < (int64_t)count; ++r) {
        double dr = (double)a[r].re - b[r].re, di = (double)a[r].im - b[r].im;
        num += dr * dr + di * di;
        den += (double)b[r].re * b[r].re + (double)b[r].im * b[r].im;
    }
    return den > 0 ? sqrt(num / den) : sqrt(num);
}

/* ---------------------------------------------------------------------------
 * One engine primitive (include/aqs_engine.h, struct aqs_op) in bit-position
 * space, with arbitrary control VALUES.  Used by oracle/abi_shim.c, the CPU
 * stand-in for libaqs_engine.so that lets the C++ host layer be tested in a
 * GPU-less container.  kind: 0 U2, 1 DIAG, 2 X, 3 SWAP.
 * ------------------------------------------------------------------------- */
int orc_apply_prim(c32* a, int n, int kind, int p, int p2, uint64_t cmask, uint64_t cval, const c32* m) {
    const uint64_t N = 1ULL << n, tm = 1ULL << p;
    if (kind == 0 || kind == 2) {
        const c32 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < (int64_t)(N >> 1); ++j) {
            uint64_t r0 = ins0((uint64_t)j, p);
            if ((r0 & cmask) != cval) continue;
            uint64_t r1 = r0 | tm;
            c32 x = a[r0], y = a[r1];
            if (kind == 2) { a[r0] = y; a[r1] = x; }
            else { a[r0] = cadd(cmul(m00, x), cmul(m01, y)); a[r1] = cadd(cmul(m10, x), cmul(m11, y)); }
        }
        return 0;
    }
    if (kind == 1) {
        const c32 d0 = m[0], d1 = m[3];
        const int d0_one = (d0.re == 1.f && d0.im == 0.f);
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < (int64_t)N; ++r) {
            if (((uint64_t)r & cmask) != cval) continue;
            if ((uint64_t)r & tm) a[r] = cmul(d1, a[r]);
            else if (!d0_one) a[r] = cmul(d0, a[r]);
        }
        return 0;
    }
    if (kind == 3) {
        const uint64_t mb = 1ULL << p2;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < (int64_t)N; ++r) {
            uint64_t u = (uint64_t)r;
            if ((u & cmask) != cval) continue;
            if ((u & tm) && !(u & mb)) {
                uint64_t v = (u ^ tm) | mb;
                c32 t = a[u]; a[u] = a[v]; a[v] = t;
            }
        }
        return 0;
    }
    return -1;
}

// plan.cu — compiled circuits: aqs_plan_build / aqs_plan_run and the gate-fusion planner.
//
// Replaces QCircuit::compile (reference src/quantum.cpp:199-210).  The reference
// multiplies every gate into a dense 2^n x 2^n unitary; here "compiling" turns the
// op list into a launch plan.  With AQS_PLAN_FUSE the planner
//   1. rewrites the op list (simplify): SWAP -> three controlled flips; runs of
//      single-qubit gates on one qubit merge when the product is no more expensive
//      than its factors; a singly-controlled gate becomes a MULTIPLEXED gate ("matrix
//      A where the control bit is 0, matrix B where it is 1") and merges with its
//      neighbours on the target qubit the same way, so a CX followed by a rotation is
//      one butterfly pass, not two (CX vanishes from brickwork circuits);
//   2. greedily groups ops into PASSES: a pass owns up to T index bits (the low 5
//      always, for coalescing) and takes, in program order, every op whose target
//      lies in those bits, skipping over ops it cannot take as long as the skipped
//      op commutes with everything taken later (two ops commute when on every
//      shared bit both act diagonally — as a control or a diagonal target);
//   3. splits each pass into SEGMENTS by the same rule with 5 register bits, and
//      picks for every change of layout an XOR swizzle of shared memory that makes
//      both sides bank-conflict free;
//   4. decomposes every 2x2 into in-place shears and writes the descriptors that
//      tile_kernel.cuh interprets (they travel as kernel parameters).
// Diagonal ops and controls never constrain a tile: QFT's CPhase ladder fuses freely.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "plan_internal.h"

namespace aqs {

typedef std::complex<double> cd;

// planner op: a (possibly multiplexed, possibly controlled) 2x2 on one target bit
struct POp {
    int p = 0;               // target bit
    uint64_t cmask = 0;      // controls: the op is the identity where (index & cmask) != cval
    uint64_t cval = 0;
    int mux = -1;            // bit selecting m[1] over m[0]; -1: m[0] everywhere
    bool diag = false;       // m[0] diagonal and mux == -1: never constrains a tile
    bool dead = false;
    cd m[2][4];              // row-major 2x2
    cd pre = cd(1, 0);       // diag(1, pre) is applied BEFORE m (both branches): a diagonal gate on the target folded
                             // into this op as a phase on the target-bit-1 amplitude (absorb_diagonals)
    std::vector<std::pair<int, cd>> ladder;   // non-empty: a LADDER of controlled phases with hub bit p (diag == true): the amplitudes
                             // with bit p set are multiplied by m[0][3] and by w for every listed (bit, w) whose bit is 1;
                             // cmask holds the listed bits (for the commutation rules only)
    cd gs = cd(1, 0);        // scalar that came with it (diag(d0, d1) = d0 * diag(1, d1/d0)): goes to the pass-wide scale
};

}  // namespace aqs

namespace aqs {

static inline int popc(uint64_t x) { return __builtin_popcountll(x); }

// ---- 2x2 helpers ------------------------------------------------------------
static void mat_mul(const cd* B, const cd* A, cd* C) {   // C = B * A
    cd t[4];
    t[0] = B[0] * A[0] + B[1] * A[2];
    t[1] = B[0] * A[1] + B[1] * A[3];
    t[2] = B[2] * A[0] + B[3] * A[2];
    t[3] = B[2] * A[1] + B[3] * A[3];
    for (int i = 0; i < 4; ++i) C[i] = t[i];
}
static void mat_identity(cd* M) { M[0] = M[3] = cd(1, 0); M[1] = M[2] = cd(0, 0); }
static bool z0(cd z) { return std::abs(z) < 1e-14; }
static bool mat_is_diag(const cd* M) { return z0(M[1]) && z0(M[2]); }
static bool mat_is_identity(const cd* M) { return mat_is_diag(M) && z0(M[0] - 1.0) && z0(M[3] - 1.0); }
static bool mat_equal(const cd* A, const cd* B) {
    for (int i = 0; i < 4; ++i)
        if (!z0(A[i] - B[i])) return false;
    return true;
}

// In-place form of a 2x2 for the tile kernel:  M = Shear3(a, b, g) * diag(sx, sy), with real or
// imaginary shear coefficients (tile_kernel.cuh, butterfly<>).  Falls back to TK_GEN (the direct
// 8-instruction form) when the matrix has no such structure or the shears would be ill-conditioned.
struct Decomp {
    int kind = TK_GEN;       // TK_SHR, TK_SHI, TK_GEN, TK_PERM_R or TK_PERM_I
    float c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double sx = 1.0, sy = 1.0;   // shears only: factors applied first to the target-bit-0 / target-bit-1 amplitudes
    bool pre_imag = false;       // the prescale is (i*sx, i*sy)
    double rho = 1.0;            // shears only: common modulus, M = rho * (shears * prescale).  The reference's
                                 // float matrices are orthogonal only to ~1e-7 (h = 0.70710677f, fl(cos)^2 + fl(sin)^2 != 1)
                                 // but they ARE exact scalar multiples of a rotation / reflection: rho carries the
                                 // scalar, so the fused path drifts in norm exactly like the reference does
    int cost = 13;               // FFMA2 per amplitude pair, prescale included.  The direct form is 8 instructions, but it is
                                 // not in place (register copies at every join) and cannot carry a folded diagonal: measured
                                 // 0.47 ms per op at n = 30 against 0.18 - 0.28 ms for a shear op, so two shear ops beat one direct op
};

// N (2x2 real, det 1, N00 >= 0 expected) = [[1+ab, a+g+abg],[b, 1+bg]]
static bool shear3(double n00, double n01, double n10, double n11, double& a, double& b, double& g) {
    b = n10;
    if (std::fabs(b) > 1e-7) {
        a = (n00 - 1.0) / b;
        g = (n11 - 1.0) / b;
    } else {
        if (std::fabs(n00 - 1.0) > 1e-7 || std::fabs(n11 - 1.0) > 1e-7) return false;
        a = n01;
        g = 0.0;
    }
    if (std::fabs(a) > 4.0 || std::fabs(g) > 4.0 || std::fabs(b) > 4.0) return false;
    const double r00 = 1 + a * b, r01 = a + g + a * b * g, r10 = b, r11 = 1 + b * g;
    const double err = std::max(std::max(std::fabs(r00 - n00), std::fabs(r01 - n01)), std::max(std::fabs(r10 - n10), std::fabs(r11 - n11)));
    return err < 1e-9;
}

static Decomp decompose(const cd* M) {
    Decomp d;
    auto gen = [&]() {
        d.kind = TK_GEN;
        d.cost = 13;
        for (int i = 0; i < 4; ++i) { d.c[2 * i] = (float)M[i].real(); d.c[2 * i + 1] = (float)M[i].imag(); }
        return d;
    };
    const bool re00 = std::fabs(M[0].imag()) < 1e-14, re01 = std::fabs(M[1].imag()) < 1e-14;
    const bool re10 = std::fabs(M[2].imag()) < 1e-14, re11 = std::fabs(M[3].imag()) < 1e-14;
    const bool im00 = std::fabs(M[0].real()) < 1e-14, im01 = std::fabs(M[1].real()) < 1e-14;
    const bool im10 = std::fabs(M[2].real()) < 1e-14, im11 = std::fabs(M[3].real()) < 1e-14;
    if (z0(M[0]) && z0(M[3])) {
        // anti-diagonal: pure data movement, kept exact (X, CX, Swap, Y)
        if (re01 && re10) { d.kind = TK_PERM_R; d.c[0] = (float)M[1].real(); d.c[1] = (float)M[2].real(); d.cost = 2; return d; }
        if (im01 && im10) { d.kind = TK_PERM_I; d.c[0] = (float)M[1].imag(); d.c[1] = (float)M[2].imag(); d.cost = 2; return d; }
        return gen();
    }
    double A, B, C, D;   // the four real numbers of the structured matrix
    int cls;             // 0 real, 1 XLIKE [[A, iB],[iC, D]], 2 AXLIKE [[iA, B],[C, iD]]
    if (re00 && re01 && re10 && re11) { cls = 0; A = M[0].real(); B = M[1].real(); C = M[2].real(); D = M[3].real(); }
    else if (re00 && re11 && im01 && im10) { cls = 1; A = M[0].real(); B = M[1].imag(); C = M[2].imag(); D = M[3].real(); }
    else if (im00 && im11 && re01 && re10) { cls = 2; A = M[0].imag(); B = -M[1].real(); C = -M[2].real(); D = M[3].imag(); }
    else return gen();
    const double det = (cls == 0) ? (A * D - B * C) : (A * D + B * C);
    if (std::fabs(det) < 1e-12) return gen();
    double sx, sy, a, b, g;
    // Unitary case first (every reference gate): the float entries are orthonormal only to ~1e-7,
    // so take the exact rotation nearest to M — parametrised by its angle — rather than solving the
    // shear equations for a matrix whose determinant is not quite 1 (that error would be divided by b).
    const double orth = (cls == 0) ? (A * B + C * D) : (A * B - C * D);
    if (std::fabs(A * A + C * C - 1.0) < 2e-6 && std::fabs(B * B + D * D - 1.0) < 2e-6 && std::fabs(orth) < 2e-6) {
        sx = (A >= 0 ? 1.0 : -1.0);
        sy = (det >= 0 ? sx : -sx);
        d.rho = std::sqrt(std::fabs(det));
        const double phi = std::atan2(C / sx, A / sx);      // |phi| <= pi/2
        b = std::sin(phi);
        a = g = (cls == 0 ? -1.0 : 1.0) * std::tan(0.5 * phi);
    } else {
        const double r = std::sqrt(std::fabs(det));
        sx = (A >= 0 ? r : -r);
        sy = det / sx;
        const double n00 = A / sx, n10 = C / sx, n01 = B / sy, n11 = D / sy;
        if (cls == 0) {
            if (!shear3(n00, n01, n10, n11, a, b, g)) return gen();
        } else {
            // imaginary shears: [[n00, i n01],[i n10, n11]] = [[1-ab, i(a+g-abg)],[ib, 1-bg]]
            b = n10;
            if (std::fabs(b) > 1e-7) { a = (1.0 - n00) / b; g = (1.0 - n11) / b; }
            else if (std::fabs(n00 - 1.0) < 1e-7 && std::fabs(n11 - 1.0) < 1e-7) { a = n01; g = 0.0; }
            else return gen();
            const double q00 = 1 - a * b, q01 = a + g - a * b * g, q11 = 1 - b * g;
            const double err = std::max(std::max(std::fabs(q00 - n00), std::fabs(q01 - n01)), std::fabs(q11 - n11));
            if (err > 1e-9 || std::fabs(a) > 4.0 || std::fabs(g) > 4.0 || std::fabs(b) > 4.0) return gen();
        }
    }
    d.c[0] = (float)a; d.c[1] = (float)b; d.c[2] = (float)g;
    d.kind = (cls == 0) ? TK_SHR : TK_SHI;
    d.sx = sx; d.sy = sy;
    d.pre_imag = (cls == 2);
    // real factors ride inside the op (2 FMUL2 per pair); an imaginary prescale costs a separate op
    d.cost = 3 + (d.pre_imag ? 4 : ((sx != 1.0 || sy != 1.0) ? 2 : 0));   // the prescale rides inside the op
    return d;
}

// Cost model used to decide merges, in FFMA2 per amplitude pair.  Every emitted tile op also pays
// the interpreter's dispatch (~200 cycles per warp, i.e. about kDispatch FFMA2 per pair): what
// counts is mostly HOW MANY ops a rewrite leaves.
constexpr int kDispatch = 6;
static int mat_cost(const cd* M) {        // one matrix applied alone, dispatch included
    if (mat_is_identity(M)) return 0;
    if (mat_is_diag(M)) return kDispatch + (z0(M[0] - 1.0) ? 1 : 2);
    const Decomp d = decompose(M);
    return kDispatch + d.cost;
}
// twice the average cost over the two branches; a common factor of -1 / +-i is free (global phase)
static int pair_cost(const cd* M0, const cd* M1, cd* best_f = nullptr) {
    static const cd fs[4] = {cd(1, 0), cd(-1, 0), cd(0, 1), cd(0, -1)};
    int best = 1 << 30;
    for (const cd& f : fs) {
        cd a[4], b[4];
        for (int i = 0; i < 4; ++i) { a[i] = f * M0[i]; b[i] = f * M1[i]; }
        int c;
        if (mat_is_identity(a) && mat_is_identity(b)) c = 0;
        else {
            const Decomp da = decompose(a), db = decompose(b);
            const bool ia = mat_is_identity(a), ib = mat_is_identity(b);
            const bool shear_a = ia || da.kind <= TK_SHI, shear_b = ib || db.kind <= TK_SHI;
            const bool same = shear_a && shear_b && (ia || ib || da.kind == db.kind);
            if (same) {
                // one multiplexed shear op
                c = 2 * kDispatch + (ia ? 0 : da.cost) + (ib ? 0 : db.cost);
            } else {
                c = (ia ? 0 : 2 * kDispatch + da.cost) + (ib ? 0 : 2 * kDispatch + db.cost);
            }
        }
        if (c < best) { best = c; if (best_f) *best_f = f; }
    }
    return best;
}
static int pop_cost2(const POp& o) {
    if (o.diag) return 2 * mat_cost(o.m[0]);
    return o.mux < 0 ? pair_cost(o.m[0], o.m[0]) : pair_cost(o.m[0], o.m[1]);
}

// does some branch of the op need the direct 8-instruction form?
static bool is_general(const POp& o) {
    if (o.diag) return false;
    for (int v = 0; v < (o.mux >= 0 ? 2 : 1); ++v)
        if (!mat_is_identity(o.m[v]) && decompose(o.m[v]).kind == TK_GEN) return true;
    return false;
}

// bits on which the op acts NON-diagonally / diagonally
static inline uint64_t nd_bits(const POp& c) { return c.diag ? 0ull : (1ull << c.p); }
static inline uint64_t d_bits(const POp& c) {
    return c.cmask | (c.mux >= 0 ? (1ull << c.mux) : 0ull) | (c.diag ? (1ull << c.p) : 0ull);
}

// ---- step 1: rewrite --------------------------------------------------------
static POp from_canon(const CanonOp& c) {
    POp o;
    o.p = c.p;
    o.cmask = c.cmask;
    o.cval = c.cval;
    if (c.kind == AQS_OP_X) {
        o.m[0][0] = o.m[0][3] = cd(0, 0);
        o.m[0][1] = o.m[0][2] = cd(1, 0);
    } else {
        for (int i = 0; i < 4; ++i) o.m[0][i] = cd(c.m[i].x, c.m[i].y);
    }
    o.diag = (c.kind == AQS_OP_DIAG);
    mat_identity(o.m[1]);
    return o;
}

// can the op carry a phase on y inside its butterflies?  (uncontrolled; every branch an in-place shear of one family)
static bool takes_pre(const POp& z) {
    if (z.diag || z.cmask) return false;
    int kind = -1;
    for (int v = 0; v < (z.mux >= 0 ? 2 : 1); ++v) {
        if (mat_is_identity(z.m[v])) continue;
        const Decomp d = decompose(z.m[v]);
        if (d.kind > TK_SHI || (kind >= 0 && d.kind != kind)) return false;
        kind = d.kind;
    }
    return kind >= 0;
}

// An uncontrolled diagonal gate diag(d0, d1) on qubit t (RotZ, Phase, Z, ...) commutes with everything that touches t
// only as a control or diagonally, so it can slide forward to the next butterfly on t and ride inside it as a
// factor on that op's y amplitudes (TF_CY: 4 scalar instructions per pair instead of an op of its own — in
// brickwork circuits a third of all tile ops were such phases).  d0 goes to the pass-wide scale.
static void absorb_diagonals(std::vector<POp>& ops) {
    if (std::getenv("AQS_PLAN_NO_ABSORB")) return;
    const size_t kLook = 4096;
    for (size_t i = 0; i < ops.size(); ++i) {
        POp& w = ops[i];
        if (w.dead || !w.diag || w.cmask || w.mux >= 0) continue;
        const cd d0 = w.m[0][0], d1 = w.m[0][3];
        if (std::abs(d0) < 1e-12) continue;
        const uint64_t tb = 1ull << w.p;
        for (size_t j = i + 1; j < ops.size() && j < i + kLook; ++j) {
            POp& z = ops[j];
            if (z.dead) continue;
            if (!(nd_bits(z) & tb)) continue;            // touches t diagonally or not at all: w slides past it
            if (takes_pre(z)) {
                z.pre *= d1 / d0;
                z.gs *= d0;
                w.dead = true;
            }
            break;
        }
    }
    std::vector<POp> kept;
    kept.reserve(ops.size());
    for (POp& o : ops)
        if (!o.dead) kept.push_back(o);
    ops.swap(kept);
}

// Runs of controlled phases diag(1, e^{i theta}) on ONE target, each under a single control — QFT's ladders
// CPhase(j, i), j < i — become one TK_LADDER op (tile_kernel.cuh): one dispatch and ~2.5 butterflies' worth of
// arithmetic instead of one phase op per control.
static void group_ladders(std::vector<POp>& ops) {
    if (std::getenv("AQS_PLAN_NO_LADDER")) return;
    auto member = [](const POp& o) {
        return !o.dead && o.diag && o.mux < 0 && o.ladder.empty() && z0(o.m[0][0] - 1.0) && popc(o.cmask) <= 1 && o.cval == o.cmask &&
               std::fabs(std::abs(o.m[0][3]) - 1.0) < 1e-6;
    };
    std::vector<POp> out;
    out.reserve(ops.size());
    for (size_t i = 0; i < ops.size();) {
        if (!member(ops[i])) { out.push_back(ops[i++]); continue; }
        size_t j = i;
        uint64_t used = 0;
        int controlled = 0;
        while (j < ops.size() && member(ops[j]) && ops[j].p == ops[i].p && !(ops[j].cmask & used)) {
            used |= ops[j].cmask;
            controlled += ops[j].cmask != 0;
            ++j;
        }
        if (controlled < 3) { out.push_back(ops[i++]); continue; }
        POp L;
        L.p = ops[i].p;
        L.diag = true;
        mat_identity(L.m[0]);
        mat_identity(L.m[1]);
        for (size_t k = i; k < j; ++k) {
            if (ops[k].cmask == 0) L.m[0][3] *= ops[k].m[0][3];
            else L.ladder.push_back({__builtin_ctzll(ops[k].cmask), ops[k].m[0][3]});
        }
        L.cmask = L.cval = used;
        out.push_back(L);
        i = j;
    }
    ops.swap(out);
}

static std::vector<POp> simplify(int n, const std::vector<CanonOp>& in) {
    std::vector<POp> out;
    out.reserve(in.size() + 16);
    std::vector<int> open(n, -1);      // index in `out` of an uncontrolled op on bit b that later ops may merge into
    std::vector<int> last_nd(n, -1);   // index in `out` of the last op acting non-diagonally on bit b

    auto touch = [&](uint64_t nd, uint64_t dg) {
        for (int b = 0; b < n; ++b) {
            if (open[b] < 0) continue;
            if (nd >> b & 1ull) open[b] = -1;
            else if ((dg >> b & 1ull) && !out[open[b]].diag) open[b] = -1;
        }
    };
    auto append = [&](const POp& o) {
        touch(nd_bits(o), d_bits(o));
        out.push_back(o);
        const int idx = (int)out.size() - 1;
        if (o.cmask == 0) open[o.p] = idx;
        if (!o.diag) last_nd[o.p] = idx;
    };
    // U (uncontrolled, unmultiplexed, on bit t) arrives
    auto push_plain = [&](const POp& u) {
        const int t = u.p;
        if (open[t] >= 0) {
            POp& w = out[open[t]];
            if (w.diag && u.diag) {                       // diagonal * diagonal, in place
                mat_mul(u.m[0], w.m[0], w.m[0]);
                return;
            }
            if (!w.diag) {
                // everything between w and here leaves bit t alone, so u commutes back to w
                POp prod = w;
                mat_mul(u.m[0], w.m[0], prod.m[0]);
                if (w.mux >= 0) mat_mul(u.m[0], w.m[1], prod.m[1]);
                // (a diagonal u that would turn w into a general matrix stays separate: it may still ride
                // inside the NEXT butterfly on this qubit as a phase, absorb_diagonals)
                if (pop_cost2(prod) <= pop_cost2(w) + pop_cost2(u) && !(u.diag && is_general(prod) && !is_general(w))) {
                    w = prod;
                    return;
                }
            } else {
                // pending diagonal w, non-diagonal u: w may slide forward to u (ops in between touch t only diagonally)
                POp prod = u;
                mat_mul(u.m[0], w.m[0], prod.m[0]);
                if (pop_cost2(prod) <= pop_cost2(w) + pop_cost2(u) && !(is_general(prod) && !is_general(u))) {
                    w.dead = true;
                    open[t] = -1;
                    append(prod);
                    return;
                }
            }
        }
        append(u);
    };
    // a gate with exactly one control c (value v) on target t: multiplexed form
    auto push_ctrl1 = [&](const POp& u, int c, int v) {
        const int t = u.p;
        if (open[t] >= 0) {
            POp& w = out[open[t]];
            if (!w.diag && (w.mux < 0 || w.mux == c) && last_nd[c] < open[t]) {
                POp prod = w;
                if (w.mux < 0) {
                    for (int i = 0; i < 4; ++i) prod.m[1][i] = w.m[0][i];
                    prod.mux = c;
                }
                mat_mul(u.m[0], prod.m[v], prod.m[v]);
                // u alone would cost mat_cost(u) on one branch and nothing on the other
                if (pop_cost2(prod) <= pop_cost2(w) + mat_cost(u.m[0]) + kDispatch) {
                    // w now reads bit c: a non-diagonal op pending on c can no longer accept later merges
                    if (open[c] >= 0 && !out[open[c]].diag) open[c] = -1;
                    w = prod;
                    return;
                }
            }
        }
        POp o = u;
        o.cmask = 0; o.cval = 0;
        o.mux = c;
        if (v == 1) {
            for (int i = 0; i < 4; ++i) o.m[1][i] = u.m[0][i];
            mat_identity(o.m[0]);
        } else {
            mat_identity(o.m[1]);
        }
        append(o);
    };
    auto push = [&](const POp& o) {
        if (o.diag) {
            if (o.cmask == 0) {
                if (mat_is_identity(o.m[0])) return;
                push_plain(o);
            } else {
                append(o);
            }
            return;
        }
        if (o.cmask == 0) {
            if (mat_is_identity(o.m[0])) return;
            push_plain(o);
        } else if (popc(o.cmask) == 1) {
            const int c = __builtin_ctzll(o.cmask);
            push_ctrl1(o, c, (int)((o.cval >> c) & 1ull));
        } else {
            append(o);
        }
    };

    for (const CanonOp& c : in) {
        if (c.identity) continue;
        if (c.kind == AQS_OP_SWAP) {
            // swap(a,b) = flip(b | a) flip(a | b) flip(b | a), each under the swap's own controls
            CanonOp f = c;
            f.kind = AQS_OP_X; f.p2 = -1;
            const int a = c.p, b = c.p2;
            f.p = b; f.cmask = c.cmask | (1ull << a); f.cval = c.cval | (1ull << a); push(from_canon(f));
            f.p = a; f.cmask = c.cmask | (1ull << b); f.cval = c.cval | (1ull << b); push(from_canon(f));
            f.p = b; f.cmask = c.cmask | (1ull << a); f.cval = c.cval | (1ull << a); push(from_canon(f));
        } else {
            push(from_canon(c));
        }
    }

    // normalise: drop identities, unmultiplex equal branches, turn diagonal results into diagonal ops
    std::vector<POp> kept;
    kept.reserve(out.size());
    for (POp& o : out) {
        if (o.dead) continue;
        if (o.mux >= 0 && mat_equal(o.m[0], o.m[1])) o.mux = -1;
        if (o.mux < 0) {
            if (mat_is_identity(o.m[0])) continue;
            if (!o.diag && mat_is_diag(o.m[0])) o.diag = true;
            kept.push_back(o);
        } else if (mat_is_diag(o.m[0]) && mat_is_diag(o.m[1])) {
            for (int v = 0; v < 2; ++v) {
                if (mat_is_identity(o.m[v])) continue;
                POp dgn;
                dgn.p = o.p;
                dgn.cmask = o.cmask | (1ull << o.mux);
                dgn.cval = o.cval | ((uint64_t)v << o.mux);
                dgn.diag = true;
                for (int i = 0; i < 4; ++i) dgn.m[0][i] = o.m[v][i];
                mat_identity(dgn.m[1]);
                kept.push_back(dgn);
            }
        } else {
            // A branch that is a unit scalar f (-1, +-i) times a cheaper matrix hands the scalar to the multiplexing
            // qubit as a diagonal gate: CX folded into a RotX gives the branch RotX * X = -i * (an X-like rotation),
            // which otherwise costs an imaginary prescale on every pair (measured: 0.47 ms per op at n = 30 against
            // 0.18 ms).  The diagonal commutes with the op (the multiplexing qubit is only read) and usually rides
            // inside the next butterfly on that qubit (absorb_diagonals).
            static const cd fs[3] = {cd(-1, 0), cd(0, 1), cd(0, -1)};
            int best = pair_cost(o.m[0], o.m[1]), which = -1;
            cd bf(1, 0);
            if (o.cmask == 0 && !std::getenv("AQS_PLAN_NO_BRANCH_PHASE"))
                for (int br = 0; br < 2; ++br)
                    for (const cd& f : fs) {
                        cd t[4];
                        for (int i = 0; i < 4; ++i) t[i] = std::conj(f) * o.m[br][i];
                        const int c = br ? pair_cost(o.m[0], t) : pair_cost(t, o.m[1]);
                        if (c < best) { best = c; bf = f; which = br; }
                    }
            if (which >= 0) {
                for (int i = 0; i < 4; ++i) o.m[which][i] *= std::conj(bf);
                kept.push_back(o);
                POp dgn;
                dgn.p = o.mux;
                dgn.diag = true;
                mat_identity(dgn.m[0]);
                mat_identity(dgn.m[1]);
                dgn.m[0][which ? 3 : 0] = bf;
                kept.push_back(dgn);
            } else {
                kept.push_back(o);
            }
        }
    }
    group_ladders(kept);
    absorb_diagonals(kept);
    return kept;
}

// ---- step 2/3: grouping -----------------------------------------------------
// Greedy "take what fits, skip what commutes": selects indices of `ops` (in order) whose
// non-diagonal target bits fit in `cap` more bits on top of `base_bits`, at most `max_take` ops.
// `scoring`: the caller only counts `taken` (tile choice): stop as soon as every allowed bit is blocked —
// no butterfly can be taken any more — instead of walking the whole window.
static uint64_t greedy_group(const std::vector<POp>& ops, const std::vector<int>& cand, uint64_t base_bits,
                             uint64_t allowed_bits, int cap, size_t max_take, std::vector<int>& taken, std::vector<int>& rest,
                             bool scoring = false) {
    uint64_t bits = base_bits, blocked_nd = 0, blocked_d = 0;
    int room = cap;
    taken.clear();
    rest.clear();
    for (int idx : cand) {
        if (scoring && (allowed_bits & ~(blocked_nd | blocked_d)) == 0) break;
        const POp& c = ops[idx];
        const uint64_t nd = nd_bits(c), dd = d_bits(c);
        const bool conflict = (nd & (blocked_nd | blocked_d)) || (dd & blocked_nd);
        const uint64_t need = nd & ~bits;
        const bool placeable = (need & ~allowed_bits) == 0 && popc(need) <= room && taken.size() < max_take;
        if (!conflict && placeable) {
            bits |= need;
            room -= popc(need);
            taken.push_back(idx);
        } else {
            blocked_nd |= nd;
            blocked_d |= dd;
            rest.push_back(idx);
        }
    }
    return bits;
}

// Grow a bit set one bit at a time, each time adding the candidate bit that lets a group take
// the most ops from the head of `cand`.
// `noise` > 0 perturbs the scores (seeded, reproducible): the planner runs several such variants and
// keeps the plan with the fewest passes (build_fused).
static uint64_t choose_bits(const std::vector<POp>& ops, const std::vector<int>& cand, uint64_t base, uint64_t pool_limit,
                            int count, int n, size_t max_take, uint64_t& lcg, int noise) {
    const size_t kScore = std::min<size_t>(cand.size(), 768);
    std::vector<int> head(cand.begin(), cand.begin() + kScore), t2, r2;
    uint64_t chosen = base, pool = 0;
    for (int idx : head) pool |= nd_bits(ops[idx]);
    pool &= pool_limit & ~base;
    for (int step = 0; step < count && pool; ++step) {
        int best_bit = -1;
        size_t best = 0;
        for (int b = 0; b < n; ++b) {
            if (!(pool >> b & 1ull)) continue;
            greedy_group(ops, head, chosen | (1ull << b), chosen | (1ull << b), 0, max_take, t2, r2, true);
            size_t gain = 0;
            for (int idx : t2) gain += (nd_bits(ops[idx]) & ~base) ? 4 : 1;
            if (noise) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                gain = gain * 16 + (size_t)((lcg >> 40) % (uint64_t)(noise * 16));
            }
            if (best_bit < 0 || gain > best) { best = gain; best_bit = b; }
        }
        if (best_bit < 0) break;
        chosen |= 1ull << best_bit;
        pool &= ~(1ull << best_bit);
    }
    return chosen;
}

// A layout of the T tile-local index bits: 5 register bits, the rest thread bits (ascending).
struct Layout {
    int T = 0;
    int rpos[kRegBits];              // local position of register bit i
    int tpos[kMaxThreadBits];        // local position of threadIdx bit j
    int reg_of[16], thr_of[16];      // inverse maps (-1 if the position is of the other kind)
    bool same(const Layout& o) const { return std::memcmp(rpos, o.rpos, sizeof rpos) == 0; }
};
static Layout make_layout(int T, uint32_t reg_positions) {
    Layout L;
    L.T = T;
    int ri = 0, ti = 0;
    for (int j = 0; j < 16; ++j) L.reg_of[j] = L.thr_of[j] = -1;
    for (int j = 0; j < T; ++j) {
        if (reg_positions >> j & 1u) { L.reg_of[j] = ri; L.rpos[ri++] = j; }
        else { L.thr_of[j] = ti; L.tpos[ti++] = j; }
    }
    return L;
}

static int rank4(const uint32_t* v) {   // rank over GF(2) of four 4-bit vectors
    uint32_t basis[4] = {0, 0, 0, 0};
    int r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t x = v[i] & 15u;
        for (int b = 3; b >= 0 && x; --b) {
            if (!(x >> b & 1u)) continue;
            if (!basis[b]) { basis[b] = x; ++r; x = 0; break; }
            x ^= basis[b];
        }
    }
    return r;
}

// XOR swizzle for the re-split A -> B: slot(L) = L ^ XOR_{h >= 4, bit h of L} col[h], col[h] < 16.
// 64-bit shared accesses are served per half-warp: the four low threadIdx bits of each layout must
// map to four independent vectors in the low 4 slot bits.
static void choose_swizzle(const Layout& A, const Layout& B, uint32_t* col) {
    const int T = A.T;
    for (int h = 0; h < 16; ++h) col[h] = 0;
    auto ok = [&](const Layout& L) {
        uint32_t v[4];
        for (int j = 0; j < 4; ++j) {
            const int p = L.tpos[j];
            v[j] = (p < 4) ? (1u << p) : col[p];
        }
        return rank4(v) == 4;
    };
    if (ok(A) && ok(B)) return;
    uint32_t free_pos = 0;
    for (int j = 0; j < 4; ++j) {
        if (A.tpos[j] >= 4) free_pos |= 1u << A.tpos[j];
        if (B.tpos[j] >= 4) free_pos |= 1u << B.tpos[j];
    }
    uint64_t lcg = 0x9E3779B97F4A7C15ull ^ ((uint64_t)free_pos << 17);
    for (int it = 0; it < 20000; ++it) {
        for (int h = 4; h < T; ++h) {
            if (!(free_pos >> h & 1u)) continue;
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            col[h] = (uint32_t)(lcg >> 60) & 15u;
        }
        if (ok(A) && ok(B)) return;
    }
    for (int h = 0; h < 16; ++h) col[h] = 0;   // give up: correct but conflicted
}

static void fill_cols(const Layout& L, const uint32_t* col, uint16_t* tcol, uint16_t* rcol) {
    auto slot = [&](int p) { return (uint16_t)((1u << p) ^ (p >= 4 ? col[p] : 0u)); };
    for (int j = 0; j < kMaxThreadBits; ++j) tcol[j] = (j < L.T - kRegBits) ? slot(L.tpos[j]) : 0;
    for (int i = 0; i < kRegBits; ++i) rcol[i] = slot(L.rpos[i]);
}

// ---- step 4: emission -------------------------------------------------------
struct Emitter {
    int n, T;
    uint64_t tile;                 // global bits of the tile
    int local_of_bit[64];          // global bit -> tile-local position (-1 outside)
    int block_bit[64];             // global bit outside the tile -> bit of the tile number
    const Layout* L = nullptr;
    std::vector<TileOp>* out = nullptr;

    // Split a selection (mask, value) over global bits into register / thread / block parts.
    // Returns false if the selection is empty for structural reasons (never).
    void select(uint64_t mask, uint64_t val, uint32_t& rk_mask, uint32_t& rk_val, TileOp& t) const {
        rk_mask = rk_val = 0;
        for (int b = 0; b < n; ++b) {
            if (!(mask >> b & 1ull)) continue;
            const uint32_t v = (uint32_t)((val >> b) & 1ull);
            const int j = local_of_bit[b];
            if (j < 0) {
                t.b_mask |= 1u << block_bit[b];
                t.b_val |= v << block_bit[b];
            } else if (L->reg_of[j] >= 0) {
                rk_mask |= 1u << L->reg_of[j];
                rk_val |= v << L->reg_of[j];
            } else {
                t.t_mask |= (uint16_t)(1u << L->thr_of[j]);
                t.t_val |= (uint16_t)(v << L->thr_of[j]);
            }
        }
    }
    void push(TileOp& t) const {
        if (t.t_mask || t.b_mask) t.flags |= TF_PRED;
        t.code = tile_op_code(t.kind, t.tk, t.mj, t.flags) | ((uint32_t)t.flags << 16);
        out->push_back(t);
    }
    static uint32_t pair_mask(int tk, uint32_t rk_mask, uint32_t rk_val) {
        uint32_t pairs = 0;
        for (int pr = 0; pr < kPairs; ++pr) {
            const uint32_t k0 = ((pr >> tk) << (tk + 1)) | (pr & ((1 << tk) - 1));
            if ((k0 & rk_mask) == rk_val) pairs |= 1u << pr;
        }
        return pairs;
    }
    // position of register bit r in the pair index of target register bit tk
    static int pair_bit(int tk, int r) { return r < tk ? r : r - 1; }

    // one factor on the amplitudes selected by (sm, sv); kind TK_PHASE / TK_SCALE_R / TK_SCALE_I
    void emit_factor(int kind, double fr, double fi, uint64_t sm, uint64_t sv) const {
        TileOp t;
        std::memset(&t, 0, sizeof t);
        t.kind = (uint8_t)kind;
        t.a[0] = (float)fr;
        t.a[1] = (float)fi;
        uint32_t rm, rv;
        select(sm, sv, rm, rv, t);
        uint32_t act = 0;
        for (uint32_t k = 0; k < (uint32_t)kRegs; ++k)
            if ((k & rm) == rv) act |= 1u << k;
        if (!act) return;
        t.mask = act;
        if (rm == 0) t.mj = 5;
        else if (popc(rm) == 1) t.mj = (uint8_t)(__builtin_ctz(rm) + (rv ? 0 : 8));
        else t.mj = 6;
        push(t);
    }
    void emit_phase(cd f, uint64_t sm, uint64_t sv) const {
        if (z0(f - 1.0)) return;
        if (z0(f.imag())) { emit_factor(TK_SCALE_R, f.real(), 0.0, sm, sv); return; }
        if (z0(f.real())) { emit_factor(TK_SCALE_I, f.imag(), 0.0, sm, sv); return; }
        // f = rho * e^{i theta}: the kernel rotates (re, im) by |theta| <= pi/2 with three shears; a
        // sign or a modulus other than 1 goes first as a real scale
        double rho = std::abs(f), theta = std::arg(f);
        bool neg = false;
        if (theta > M_PI / 2) { theta -= M_PI; neg = true; }
        else if (theta < -M_PI / 2) { theta += M_PI; neg = true; }
        if (std::fabs(rho - 1.0) > 3e-7) emit_factor(TK_SCALE_R, rho, 0.0, sm, sv);
        emit_factor(neg ? TK_PHASE_N : TK_PHASE, -std::tan(0.5 * theta), std::sin(theta), sm, sv);
    }
    // The prescale of a shear decomposition rides inside the shear op (TF_PY): real factors for both
    // families, purely imaginary ones (the X * RotX family) for TK_SHI.
    // butterfly of one decomposition under controls (cm, cv); d0 != nullptr: multiplexed on `mux_bit`
    // (d applies where the bit is 1, *d0 where it is 0) — both must be the same shear kind
    // `pre` != 1 (shears only): diag(1, pre) applied first, i.e. the factor on y becomes complex (TF_CY)
    void emit_butterfly(const Decomp& d, const Decomp* d0, int mux_bit, int tk, uint64_t cm, uint64_t cv, cd pre = cd(1, 0)) const {
        TileOp t;
        std::memset(&t, 0, sizeof t);
        t.kind = (uint8_t)d.kind;
        t.tk = (uint8_t)tk;
        for (int i = 0; i < 8; ++i) t.a[i] = d.c[i];
        uint32_t rm, rv;
        select(cm, cv, rm, rv, t);
        const bool shearing = d.kind <= TK_SHI;
        if (shearing) {
            // c[3] = factor on y; the op needs the prescale variant when any set has one
            t.a[3] = (float)d.sy;
            t.b[3] = d0 ? (float)d0->sy : 1.f;
            t.sx[0] = (float)d.sx;
            t.sx[1] = d0 ? (float)d0->sx : 1.f;
            if (d.sy != 1.0 || d.sx != 1.0 || d.pre_imag || (d0 && (d0->sy != 1.0 || d0->sx != 1.0 || d0->pre_imag))) t.flags |= TF_PY;
            if (d.pre_imag) t.flags |= TF_IMAG_A;
            if (d0 && d0->pre_imag) t.flags |= TF_IMAG_B;
            if (pre != cd(1, 0)) {
                const cd fa = cd(d.sy, 0) * (d.pre_imag ? cd(0, 1) : cd(1, 0)) * pre;
                const cd fb = d0 ? cd(d0->sy, 0) * (d0->pre_imag ? cd(0, 1) : cd(1, 0)) * pre : fa;
                t.a[3] = (float)fa.real(); t.qy[0] = (float)fa.imag();
                t.b[3] = (float)fb.real(); t.qy[1] = (float)fb.imag();
                t.flags |= TF_PY | TF_CY;
            }
        }
        if (d0) {
            // the multiplexing bit: register -> pair subsets; thread / outside the tile -> predicate picks the set
            for (int i = 0; i < 3; ++i) t.b[i] = d0->c[i];
            const int mj = local_of_bit[mux_bit];
            if (mj >= 0 && L->reg_of[mj] >= 0) {
                t.flags |= TF_REGMUX;
                t.mj = (uint8_t)pair_bit(tk, L->reg_of[mj]);
                t.mask = pair_mask(tk, 1u << L->reg_of[mj], 1u << L->reg_of[mj]);
            } else {
                t.flags |= TF_MUX;
                t.mask = 0xffffu;
                TileOp sel;
                std::memset(&sel, 0, sizeof sel);
                uint32_t a2, b2;
                select(1ull << mux_bit, 1ull << mux_bit, a2, b2, sel);
                t.t_mask = sel.t_mask; t.t_val = sel.t_val; t.b_mask = sel.b_mask; t.b_val = sel.b_val;
            }
            push(t);
            return;
        }
        t.mask = pair_mask(tk, rm, rv);
        if (!t.mask) return;
        if (shearing) {
            if (rm == 0) {
                t.mj = 0;                       // every pair: the kernel uses set a for both subsets
                for (int i = 0; i < 4; ++i) t.b[i] = t.a[i];     // (and finds a copy of it in set b)
                t.sx[1] = t.sx[0];
                t.qy[1] = t.qy[0];
                if (t.flags & TF_IMAG_A) t.flags |= TF_IMAG_B;
            } else if (popc(rm) == 1 && rv == rm) {
                t.flags |= TF_REGMUX;           // one control on a register bit: set b = identity leaves the other pairs alone
                t.mj = (uint8_t)pair_bit(tk, __builtin_ctz(rm));
                for (int i = 0; i < 8; ++i) t.b[i] = 0.f;
                t.b[3] = 1.f;
            } else if (popc(rm) == 1) {
                // control value 0: swap the roles (identity on the pairs with the bit set)
                t.flags |= TF_REGMUX;
                t.mj = (uint8_t)pair_bit(tk, __builtin_ctz(rm));
                for (int i = 0; i < 8; ++i) { t.b[i] = t.a[i]; t.a[i] = 0.f; }
                t.a[3] = 1.f;
                t.sx[1] = t.sx[0];
                t.sx[0] = 1.f;
                if (t.flags & TF_IMAG_A) t.flags = (uint8_t)((t.flags & ~TF_IMAG_A) | TF_IMAG_B);
                t.mask = pair_mask(tk, rm, rm);
            }
            // (several controls on register bits never reach here: emit_matrix sends them down the direct path)
        }
        push(t);
    }
    void emit_matrix(const cd* M, int bit, int tk, uint64_t cm, uint64_t cv, cd* pass_scale = nullptr, cd pre = cd(1, 0)) const {
        if (mat_is_identity(M) && pre == cd(1, 0)) return;
        Decomp d = decompose(M);
        if (pre != cd(1, 0) && (d.kind > TK_SHI || cm)) {
            // no in-place form to carry the phase: multiply it into the matrix
            cd M2[4] = {M[0], M[1] * pre, M[2], M[3] * pre};
            emit_matrix(M2, bit, tk, cm, cv, pass_scale);
            return;
        }
        (void)bit;
        if (d.kind <= TK_SHI && d.rho != 1.0) {
            // the common modulus of an uncontrolled op is a global factor; a controlled one keeps it in its prescale
            if (pass_scale) *pass_scale *= d.rho;
            else { d.sx *= d.rho; d.sy *= d.rho; }
            d.rho = 1.0;
        }
        if (d.kind <= TK_SHI) {
            // shear bodies resolve pair subsets at compile time (one control on a register bit at most);
            // with more, fall back to the direct 2x2, which takes any pair mask
            TileOp probe;
            std::memset(&probe, 0, sizeof probe);
            uint32_t rm, rv;
            select(cm, cv, rm, rv, probe);
            if (popc(rm) >= 2) {
                d = Decomp();
                d.kind = TK_GEN;
                for (int i = 0; i < 4; ++i) { d.c[2 * i] = (float)M[i].real(); d.c[2 * i + 1] = (float)M[i].imag(); }
            }
        }
        emit_butterfly(d, nullptr, -1, tk, cm, cv, pre);
    }
    void emit_ladder(const POp& o) const {
        const uint64_t tb = 1ull << o.p;
        TileOp h;
        std::memset(&h, 0, sizeof h);
        h.kind = TK_LADDER;
        uint32_t rm, rv;
        select(tb, tb, rm, rv, h);                      // the hub: register bit -> register subset, else a predicate
        h.mj = rm ? (uint8_t)__builtin_ctz(rm) : 5;
        h.mask = 0;
        for (uint32_t k = 0; k < (uint32_t)kRegs; ++k)
            if ((k & rm) == rv) h.mask |= 1u << k;
        h.a[0] = (float)o.m[0][3].real();
        h.a[1] = (float)o.m[0][3].imag();
        TileOp reg;                                     // factors of the controls on register bits
        std::memset(&reg, 0, sizeof reg);
        reg.kind = TK_LADDER_CONT;
        for (int r = 0; r < 4; ++r) reg.a[2 * r] = 1.f;
        reg.sx[0] = 1.f;
        std::vector<TileOp> cont;
        int fill = 4;
        for (const auto& cw : o.ladder) {
            const int j = local_of_bit[cw.first];
            const float wr = (float)cw.second.real(), wi = (float)cw.second.imag();
            if (j >= 0 && L->reg_of[j] >= 0) {
                const int r = L->reg_of[j];
                if (r < 4) { reg.a[2 * r] = wr; reg.a[2 * r + 1] = wi; }
                else { reg.sx[0] = wr; reg.sx[1] = wi; }
                continue;
            }
            if (fill == 4) {
                TileOp c;
                std::memset(&c, 0, sizeof c);
                c.kind = TK_LADDER_CONT;
                c.mask = 0x3f3f3f3fu;
                for (int q = 0; q < 4; ++q) c.a[2 * q] = 1.f;
                cont.push_back(c);
                fill = 0;
            }
            TileOp& c = cont.back();
            const uint32_t code = (j >= 0) ? (uint32_t)L->thr_of[j] : (0x20u | (uint32_t)block_bit[cw.first]);
            c.mask = (c.mask & ~(0xffu << (8 * fill))) | (code << (8 * fill));
            c.a[2 * fill] = wr;
            c.a[2 * fill + 1] = wi;
            ++fill;
        }
        const uint32_t n_cont = (uint32_t)cont.size();
        std::memcpy(&h.sx[0], &n_cont, sizeof n_cont);
        push(h);
        push(reg);
        for (TileOp& c : cont) push(c);
    }
    void emit(const POp& o, cd& pass_scale) const {
        const uint64_t tb = 1ull << o.p;
        if (!o.ladder.empty()) { emit_ladder(o); return; }
        if (o.diag) {
            const cd d0 = o.m[0][0], d1 = o.m[0][3];
            if (z0(d0 - 1.0)) {
                emit_phase(d1, o.cmask | tb, o.cval | tb);
            } else if (o.cmask == 0) {
                // diag(d0, d1) = d0 * diag(1, d1/d0): d0 goes to the pass-wide scale
                pass_scale *= d0;
                emit_phase(d1 / d0, tb, tb);
            } else {
                emit_phase(d0, o.cmask | tb, o.cval);
                emit_phase(d1, o.cmask | tb, o.cval | tb);
            }
            return;
        }
        const int tk = L->reg_of[local_of_bit[o.p]];
        cd m0[4], m1[4];
        for (int i = 0; i < 4; ++i) { m0[i] = o.m[0][i]; m1[i] = o.m[o.mux < 0 ? 0 : 1][i]; }
        cd pre = o.pre;
        if (pre != cd(1, 0) && o.cmask != 0) {           // (absorb_diagonals only picks uncontrolled ops)
            for (int i = 1; i < 4; i += 2) { m0[i] *= pre; m1[i] *= pre; }
            pre = cd(1, 0);
        }
        if (o.cmask == 0) pass_scale *= o.gs;
        if (o.cmask == 0) {
            // a common factor -1 / +-i that makes the decompositions cheaper is a global phase
            cd f(1, 0);
            pair_cost(m0, m1, &f);
            if (f != cd(1, 0)) {
                for (int i = 0; i < 4; ++i) { m0[i] *= f; m1[i] *= f; }
                pass_scale *= std::conj(f);
            }
        }
        if (o.mux < 0) {
            emit_matrix(m0, o.p, tk, o.cmask, o.cval, o.cmask == 0 ? &pass_scale : nullptr, pre);
            return;
        }
        const uint64_t mb = 1ull << o.mux;
        Decomp d0 = decompose(m0), d1 = decompose(m1);
        // the identity is a shear of either family with zero coefficients
        if (mat_is_identity(m0) && d1.kind <= TK_SHI) { d0 = Decomp(); d0.kind = d1.kind; d0.cost = 0; }
        if (mat_is_identity(m1) && d0.kind <= TK_SHI) { d1 = Decomp(); d1.kind = d0.kind; d1.cost = 0; }
        // (a default Decomp has zero shear coefficients and sx = sy = 1: the identity)
        if (o.cmask == 0 && d0.kind == d1.kind && d0.kind <= TK_SHI) {
            // common modulus -> pass-wide scale; a branch whose modulus differs keeps the ratio in its prescale
            const double common = d0.rho;
            pass_scale *= common;
            d1.sx *= d1.rho / common; d1.sy *= d1.rho / common;
            d0.rho = d1.rho = 1.0;
            emit_butterfly(d1, &d0, o.mux, tk, 0, 0, pre);
            return;
        }
        if (pre != cd(1, 0))
            for (int i = 1; i < 4; i += 2) { m0[i] *= pre; m1[i] *= pre; }
        emit_matrix(m0, o.p, tk, o.cmask | mb, o.cval);
        emit_matrix(m1, o.p, tk, o.cmask | mb, o.cval | mb);
    }
};

// One planning run over the simplified op list.  variant 0 is the plain greedy; variant v > 0 perturbs
// the tile-choice scores with a seeded generator.  Fills `passes`; returns an aqs_status and, on
// failure, the message in `err` (the function runs on worker threads: no thread-local error state).
static int plan_variant(int n, int T, const std::vector<POp>& ops, int variant, std::vector<FusedPass>& passes, std::string& err) {
    const int TB = T - kRegBits;
    const uint64_t all_bits = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    const uint64_t low_mask = (1ull << kLaneBits) - 1ull;
    const size_t kMaxTake = kOpsLarge;         // upper bound only: the emission loop enforces the descriptor budget
    uint64_t lcg = (uint64_t)variant * 0x9E3779B97F4A7C15ull + 1ull;
    const int noise = variant ? 4 : 0;
    auto fail = [&](int code, const char* msg) { err = msg; return code; };
    passes.clear();
    // Sliding window over the op stream: a pass looks at the ops deferred by earlier passes plus
    // the next kWindow ops, so planning is O(passes * window) even for million-op circuits
    // (Grover-26 with 6433 iterations lowers to ~1e6 ops).
    const size_t kWindow = 4096;
    size_t next = 0;
    std::vector<int> cand, taken, rest;
    while (true) {
        while (cand.size() < kWindow && next < ops.size()) cand.push_back((int)next++);
        if (cand.empty()) break;
        // Tile choice: grow the tile one bit at a time, each time adding the bit that lets the
        // pass take the most ops from the head of the window (skipped for huge op lists, where
        // planning time matters more).
        uint64_t tile;
        if (ops.size() <= 60000) {
            const uint64_t chosen = choose_bits(ops, cand, low_mask, all_bits, T - kLaneBits, n, kMaxTake, lcg, noise);
            tile = greedy_group(ops, cand, chosen, chosen, 0, kMaxTake, taken, rest);
        } else {
            tile = greedy_group(ops, cand, low_mask, all_bits, T - kLaneBits, kMaxTake, taken, rest);
        }
        if (taken.empty()) return fail(AQS_ERR_STATE, "fusion planner made no progress");
        // pad the tile to exactly T bits with the lowest unused positions
        for (int b = 0; b < n && popc(tile) < T; ++b) tile |= 1ull << b;

        FusedPass fp;
        fp.T = T;
        fp.n_tiles = 1ull << (n - T);
        Emitter em;
        em.n = n; em.T = T; em.tile = tile;
        em.out = &fp.ops;
        fp.tile.n = 0;
        int nb = 0;
        for (int b = 0; b < 64; ++b) em.local_of_bit[b] = em.block_bit[b] = -1;
        for (int b = 0; b < n; ++b) {
            if (tile >> b & 1ull) {
                em.local_of_bit[b] = fp.tile.n;
                fp.tile.pos[fp.tile.n++] = (uint8_t)b;
            } else {
                em.block_bit[b] = nb++;
            }
        }
        uint64_t global_of_local[16];
        for (int j = 0; j < T; ++j) global_of_local[j] = 1ull << fp.tile.pos[j];

        // segments: the same greedy with 5 register bits among ALL tile bits
        std::vector<Layout> layouts;
        std::vector<std::pair<uint32_t, uint32_t>> ranges;   // (first_op, n_ops) per layout
        std::vector<int> seg_cand = taken, seg_taken, seg_rest;
        cd pass_scale(1.0, 0.0);
        auto default_regs = [&](uint32_t want) {
            // complete `want` (a set of local positions) to 5 register positions, highest free first
            for (int j = T - 1; j >= 0 && popc(want) < kRegBits; --j)
                if (!(want >> j & 1u)) want |= 1u << j;
            return want;
        };
        std::vector<int> unplaced;
        while (!seg_cand.empty()) {
            if ((int)layouts.size() >= kMaxSegs - 2) { unplaced = seg_cand; break; }
            uint64_t chosen = greedy_group(ops, seg_cand, 0, tile, kRegBits, kMaxTake, seg_taken, seg_rest);
            if (seg_taken.empty()) return fail(AQS_ERR_STATE, "fusion planner made no progress (segment)");
            uint32_t regs = 0;
            for (int b = 0; b < n; ++b)
                if (chosen >> b & 1ull) regs |= 1u << em.local_of_bit[b];
            // an all-diagonal segment keeps the previous layout
            if (regs == 0 && !layouts.empty()) {
                for (int i = 0; i < kRegBits; ++i) regs |= 1u << layouts.back().rpos[i];
            }
            regs = default_regs(regs);
            Layout L = make_layout(T, regs);
            const uint32_t first = (uint32_t)fp.ops.size();
            em.L = &L;
            // emit until the pass's descriptor budget is nearly used (an op emits at most 4 tile ops);
            // whatever is left goes back to the candidates, in program order
            bool full = false;
            for (size_t i = 0; i < seg_taken.size(); ++i) {
                if (fp.ops.size() + 4 + ops[seg_taken[i]].ladder.size() / 4 + 2 > (size_t)kOpsLarge) {
                    unplaced.assign(seg_taken.begin() + i, seg_taken.end());
                    std::vector<int> merged(unplaced.size() + seg_rest.size());
                    std::merge(unplaced.begin(), unplaced.end(), seg_rest.begin(), seg_rest.end(), merged.begin());
                    unplaced.swap(merged);
                    full = true;
                    break;
                }
                em.emit(ops[seg_taken[i]], pass_scale);
            }
            const uint32_t cnt = (uint32_t)fp.ops.size() - first;
            if (!layouts.empty() && layouts.back().same(L)) {
                ranges.back().second += cnt;
            } else {
                layouts.push_back(L);
                ranges.push_back({first, cnt});
            }
            if (full) break;
            seg_cand.swap(seg_rest);
        }
        if (fp.ops.size() > (size_t)kOpsLarge) return fail(AQS_ERR_STATE, "fusion planner overflowed a pass");
        if (!unplaced.empty()) {
            // the pass ran out of segments: give the remaining ops back, in program order
            std::vector<int> merged(unplaced.size() + rest.size());
            std::merge(unplaced.begin(), unplaced.end(), rest.begin(), rest.end(), merged.begin());
            rest.swap(merged);
        }
        // entry / exit layouts must keep the low five index bits on the lanes (coalesced global access)
        auto io_ok = [&](const Layout& L) {
            for (int i = 0; i < kRegBits; ++i)
                if (L.rpos[i] < kLaneBits) return false;
            return true;
        };
        auto io_layout_for = [&](const Layout& L) {
            uint32_t keep = 0;
            for (int i = 0; i < kRegBits; ++i)
                if (L.rpos[i] >= kLaneBits) keep |= 1u << L.rpos[i];
            uint32_t want = keep;
            for (int j = T - 1; j >= kLaneBits && popc(want) < kRegBits; --j)
                if (!(want >> j & 1u)) want |= 1u << j;
            return make_layout(T, want);
        };
        if (!io_ok(layouts.front())) {
            layouts.insert(layouts.begin(), io_layout_for(layouts.front()));
            ranges.insert(ranges.begin(), {0u, 0u});
        }
        if (!io_ok(layouts.back())) {
            layouts.push_back(io_layout_for(layouts.back()));
            ranges.push_back({(uint32_t)fp.ops.size(), 0u});
        }
        for (size_t s = 0; s < layouts.size(); ++s) {
            TileSeg sg;
            std::memset(&sg, 0, sizeof sg);
            sg.first_op = (uint16_t)ranges[s].first;
            sg.n_ops = (uint16_t)ranges[s].second;
            if (s > 0) {
                uint32_t col[16];
                choose_swizzle(layouts[s - 1], layouts[s], col);
                fill_cols(layouts[s - 1], col, sg.wr_tcol, sg.wr_rcol);
                fill_cols(layouts[s], col, sg.rd_tcol, sg.rd_rcol);
                sg.resplit = 1;
            }
            fp.segs.push_back(sg);
        }
        auto io_offsets = [&](const Layout& L, uint64_t* toff, uint64_t* roff) {
            for (int j = 0; j < kMaxThreadBits; ++j) toff[j] = (j < TB) ? global_of_local[L.tpos[j]] : 0;
            for (int i = 0; i < kRegBits; ++i) roff[i] = global_of_local[L.rpos[i]];
        };
        io_offsets(layouts.front(), fp.ld_toff, fp.ld_roff);
        io_offsets(layouts.back(), fp.st_toff, fp.st_roff);
        fp.scale = make_float2((float)pass_scale.real(), (float)pass_scale.imag());
        fp.has_scale = (pass_scale != cd(1.0, 0.0));
        for (const TileOp& t : fp.ops) fp.rare = fp.rare || tile_op_is_rare(t.code);
        passes.push_back(std::move(fp));
        cand.swap(rest);
    }
    return AQS_OK;
}

// The greedy tile choice is sensitive to ties (brickwork: 21 passes at n = 30 but 27 at n = 31).  So
// the planner runs a few seeded variants of it in parallel host threads and keeps the plan with
// the fewest passes (then the fewest layouts): 20 passes for both.  Deterministic: the seeds are fixed
// and the choice does not depend on thread timing, so every rank of a sharded run builds the same plan.
static int build_fused(aqs_plan_s* p) {
    const int n = p->n;
    int T = std::min(n, 13);      // 8192 amplitudes per tile: fewer passes (brickwork-30: 15 instead of 21); specialised kernels 63.7 ms against 68.6 at T = 12, the generic kernel 153 against 159.5
    if (const char* e = std::getenv("AQS_TILE_BITS")) {
        const int t = std::atoi(e);
        if (t >= kMinTileBits && t <= kMaxTileBits) T = std::min(n, t);
    }
    const std::vector<POp> ops = simplify(n, p->ops);
    if (std::getenv("AQS_PLAN_DUMP") && std::atoi(std::getenv("AQS_PLAN_DUMP")) > 2)
        for (size_t i = 0; i < ops.size(); ++i) {
            const POp& o = ops[i];
            std::fprintf(stderr, "  op %zu: bit %d %s cmask %llx mux %d cost %d/%d  m0 = [%.3f%+.3fi %.3f%+.3fi; %.3f%+.3fi %.3f%+.3fi]\n", i, o.p,
                         o.diag ? "diag" : "mat", (unsigned long long)o.cmask, o.mux, mat_cost(o.m[0]), o.mux >= 0 ? mat_cost(o.m[1]) : -1,
                         o.m[0][0].real(), o.m[0][0].imag(), o.m[0][1].real(), o.m[0][1].imag(), o.m[0][2].real(), o.m[0][2].imag(),
                         o.m[0][3].real(), o.m[0][3].imag());
        }
    int variants = 8;
    if (const char* e = std::getenv("AQS_PLAN_VARIANTS")) variants = std::max(1, std::min(64, std::atoi(e)));
    if (ops.size() > 20000 || ops.size() < 64 || n <= T) variants = 1;     // long circuits: planning time matters more
    std::vector<std::vector<FusedPass>> out(variants);
    std::vector<std::string> errs(variants);
    std::vector<int> rcs(variants, AQS_OK);
    if (const char* e = std::getenv("AQS_PLAN_ONLY_VARIANT")) {        // dev: run exactly one seeded variant
        variants = 1;
        rcs[0] = plan_variant(n, T, ops, std::atoi(e), out[0], errs[0]);
    } else if (variants == 1) {
        rcs[0] = plan_variant(n, T, ops, 0, out[0], errs[0]);
    } else {
        std::vector<std::thread> workers;
        for (int v = 1; v < variants; ++v)
            workers.emplace_back([&, v]() { rcs[v] = plan_variant(n, T, ops, v, out[v], errs[v]); });
        rcs[0] = plan_variant(n, T, ops, 0, out[0], errs[0]);
        for (auto& w : workers) w.join();
    }
    int best = -1;
    auto layouts = [](const std::vector<FusedPass>& ps) { size_t c = 0; for (auto& fp : ps) c += fp.segs.size(); return c; };
    for (int v = 0; v < variants; ++v) {
        if (rcs[v] != AQS_OK) continue;
        if (best < 0 || out[v].size() < out[best].size() || (out[v].size() == out[best].size() && layouts(out[v]) < layouts(out[best]))) best = v;
    }
    if (best < 0) return fail(rcs[0], errs[0]);
    p->passes = std::move(out[best]);
    return AQS_OK;
}

// Device copy of every pass's op descriptors, made on first use so that plans can be built (and
// inspected) without a GPU.
static int ensure_uploaded(aqs_plan_s* p) {
    if (p->arena || p->passes.empty()) return AQS_OK;
    size_t total = 0;
    for (auto& fp : p->passes) total += fp.ops.size() + 1;
    std::vector<DevOp> host(total);
    std::memset(host.data(), 0, total * sizeof(DevOp));
    cudaError_t e = pool_alloc(&p->arena, total * sizeof(DevOp));
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(plan arena)", __LINE__);
    p->arena_bytes = total * sizeof(DevOp);
    size_t off = 0;
    for (auto& fp : p->passes) {
        fp.d_ops = reinterpret_cast<const DevOp*>(p->arena) + off;
        for (const TileOp& t : fp.ops) {
            DevOp& d = host[off++];
            d.word = t.code;
            d.sx_a = t.sx[0];
            d.sx_b = t.sx[1];
            d.mask = t.mask;
            if (t.flags & TF_CY) std::memcpy(&d.mask, &t.qy[0], sizeof(float));   // shears never read the pair mask
            d.qy_b = t.qy[1];
            d.tpred = (uint32_t)t.t_mask | ((uint32_t)t.t_val << 16);
            d.b_mask = t.b_mask;
            d.b_val = t.b_val;
            for (int i = 0; i < 4; ++i) { d.a[i] = t.a[i]; d.b[i] = (t.kind == TK_GEN || t.kind == TK_LADDER_CONT) ? t.a[4 + i] : t.b[i]; }
        }
        host[off++].word = kDevOpEnd << 3;   // sentinel: read by the prefetch, never executed
    }
    e = cudaMemcpy(p->arena, host.data(), total * sizeof(DevOp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy(plan arena)", __LINE__);
    count_h2d(total * sizeof(DevOp));
    return AQS_OK;
}

static void fill_params(PassParams& P, float2* state, float2* state_out, const FusedPass& fp) {
    std::memset(&P, 0, sizeof P);
    P.state = state;
    P.state_out = state_out;
    P.ops = fp.d_ops;
    P.n_segs = (uint32_t)fp.segs.size();
    P.n_ops = (uint32_t)fp.ops.size();
    P.scale = fp.scale;
    P.has_scale = fp.has_scale ? 1u : 0u;
    P.tile_bits = (uint32_t)fp.T;
    std::memcpy(P.ld_toff, fp.ld_toff, sizeof P.ld_toff);
    std::memcpy(P.ld_roff, fp.ld_roff, sizeof P.ld_roff);
    std::memcpy(P.st_toff, fp.st_toff, sizeof P.st_toff);
    std::memcpy(P.st_roff, fp.st_roff, sizeof P.st_roff);
    P.tile = fp.tile;
    std::memcpy(P.segs, fp.segs.data(), fp.segs.size() * sizeof(TileSeg));
}

static size_t tile_smem_bytes(int T, size_t n_ops) { return (sizeof(float2) << T) + (n_ops + 1) * sizeof(DevOp); }

template <int T>
static cudaError_t launch_tile(const PassParams& P, uint64_t n_tiles, size_t n_ops, bool rare, cudaStream_t st) {
    if (rare) k_tile2<T, true><<<(unsigned)n_tiles, 1 << (T - kRegBits), tile_smem_bytes(T, n_ops), st>>>(P);
    else k_tile2<T, false><<<(unsigned)n_tiles, 1 << (T - kRegBits), tile_smem_bytes(T, n_ops), st>>>(P);
    return cudaGetLastError();
}

// Which tiles of a pass rank r of 2^g runs when the state is one flat address range over the GPUs
// (flat.cu).  The top g index bits are the rank.  Rank bits OUTSIDE the tile pin the matching bits of
// the tile number to the rank's own (those tiles are local); for every rank bit INSIDE the tile — the
// tile then spans several GPUs — one more non-tile local bit (the highest available) is pinned to the
// rank's value on it, so that the GPUs sharing a tile group split it evenly.  Always g pinned bits.
struct ShardCut {
    uint32_t fix_n = 0, fix_or = 0;
    uint8_t fix_pos[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int rank_bits_in_tile = 0;
};
static int shard_cut(int n, const FusedPass& fp, int rank, int g, ShardCut& cut) {
    uint64_t tile = 0;
    for (int j = 0; j < fp.tile.n; ++j) tile |= 1ull << fp.tile.pos[j];
    int compact_of[64], nb = 0;
    for (int b = 0; b < n; ++b) compact_of[b] = (tile >> b & 1ull) ? -1 : nb++;
    uint32_t mask = 0, val = 0;
    int spare = n - g - 1;                       // next candidate local bit for an in-tile rank bit
    for (int i = 0; i < g; ++i) {
        const int b = n - g + i;
        const uint32_t v = (uint32_t)(rank >> i) & 1u;
        int where;
        if (!(tile >> b & 1ull)) {
            where = compact_of[b];
        } else {
            ++cut.rank_bits_in_tile;
            while (spare >= 0 && ((tile >> spare & 1ull) || (mask >> compact_of[spare] & 1u))) --spare;
            if (spare < 0) return fail(AQS_ERR_STATE, "no local bit left to split a tile group between its GPUs");
            where = compact_of[spare--];
        }
        mask |= 1u << where;
        val |= v << where;
    }
    cut.fix_or = val;
    for (int c = 0; c < 32; ++c)
        if (mask >> c & 1u) cut.fix_pos[cut.fix_n++] = (uint8_t)c;
    return AQS_OK;
}

// `load_base`: where the pass READS the state (nullptr: in place).  Staged passes of sharded runs read a view of the
// state in which the peers' parts are local copies, and write the real thing.
static int launch_pass(float2* state, const FusedPass& fp, cudaStream_t st, const ShardCut* cut = nullptr, float2* load_base = nullptr) {
    if (fp.n_tiles > 0x7fffffffull) return fail(AQS_ERR_INVALID, "grid too large");
    PassParams P;
    fill_params(P, load_base ? load_base : state, state, fp);
    uint64_t n_tiles = fp.n_tiles;
    if (cut && cut->fix_n) {        // sharded run: 1 / 2^g of the tiles
        P.fix_n = cut->fix_n;
        P.fix_or = cut->fix_or;
        std::memcpy(P.fix_pos, cut->fix_pos, sizeof P.fix_pos);
        n_tiles >>= cut->fix_n;
    }
    if (fp.spec && spec_ready(fp)) {
        int rc = spec_launch(fp, P.state, P.state_out, n_tiles, P.fix_n, P.fix_or, P.fix_pos, st);
        if (rc) return rc;
        count_launch(1);
        return AQS_OK;
    }
    const size_t n_ops = fp.ops.size();
    cudaError_t e;
    switch (fp.T) {
        case 10: e = launch_tile<10>(P, n_tiles, n_ops, fp.rare, st); break;
        case 11: e = launch_tile<11>(P, n_tiles, n_ops, fp.rare, st); break;
        case 12: e = launch_tile<12>(P, n_tiles, n_ops, fp.rare, st); break;
        default: e = launch_tile<13>(P, n_tiles, n_ops, fp.rare, st); break;
    }
    if (e != cudaSuccess) return fail_cuda(e, "tile kernel launch", __LINE__);
    count_launch(1);
    return AQS_OK;
}

int fused_init() {
    // tile + descriptors exceed the 48 KiB default of dynamic shared memory: opt in once
    cudaError_t e = cudaSuccess;
#define AQS_OPT_IN(T)                                                                                                                          \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile2<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(T, kOpsLarge)); \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile2<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(T, kOpsLarge));
    AQS_OPT_IN(10) AQS_OPT_IN(11) AQS_OPT_IN(12) AQS_OPT_IN(13)
#undef AQS_OPT_IN
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(tile kernel)", __LINE__);
    return AQS_OK;
}

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
    p->info.n_qubits = n;
    p->info.bytes_unfused = bytes;
    if ((flags & AQS_PLAN_FUSE) && n >= kMinTileBits && n_ops > 0) {
        int rc = build_fused(p);
        if (rc) { aqs_plan_destroy(p); return rc; }
    }
    if (!p->passes.empty()) {
        // AQS_JIT=0 turns specialisation off, AQS_JIT=1 forces it (waiting), AQS_JIT=2 forces it in the background
        uint32_t jit = flags & (AQS_PLAN_JIT | AQS_PLAN_JIT_ASYNC);
        if (const char* e = std::getenv("AQS_JIT")) jit = (*e == '0') ? 0u : (*e == '2' ? AQS_PLAN_JIT_ASYNC : AQS_PLAN_JIT);
        size_t max_passes = 8192;          // (source generation is ~0.2 ms per pass; spec_attach bounds the number of NEW kernels)
        if (const char* e = std::getenv("AQS_JIT_MAX_PASSES")) max_passes = (size_t)std::max(0, std::atoi(e));
        if (jit && p->passes.size() <= max_passes) {
            p->spec_requested = true;
            spec_attach(n, p->passes, (jit & AQS_PLAN_JIT) != 0);
        }
        p->info.n_launches = p->passes.size();
        p->info.n_fused_passes = p->passes.size();
        p->info.n_single_ops = 0;
        p->info.bytes_planned = 2.0 * 8.0 * std::ldexp(1.0, n) * (double)p->passes.size();
        p->info.tile_bits = p->passes[0].T;
        if (std::getenv("AQS_PLAN_DUMP")) {
            size_t segs = 0, tops = 0;
            for (auto& fp : p->passes) { segs += fp.segs.size(); tops += fp.ops.size(); }
            std::fprintf(stderr, "[aqs plan] n=%d ops=%llu -> %zu tile ops, %zu passes, %zu layouts\n", n,
                         (unsigned long long)n_ops, tops, p->passes.size(), segs);
            if (std::atoi(std::getenv("AQS_PLAN_DUMP")) > 1)
                for (size_t i = 0; i < p->passes.size(); ++i) {
                    auto& fp = p->passes[i];
                    int km[16] = {0}, generic = 0, muxed = 0, cost = 0;
                    for (auto& t : fp.ops) {
                        km[t.kind & 15]++;
                        const bool factor = t.kind >= TK_PHASE;
                        generic += factor ? (t.mj == 6) : (t.kind <= TK_SHI ? (t.mj == 4) : (t.mask != 0xffffu));
                        muxed += (t.flags & (TF_MUX | TF_REGMUX)) != 0;
                        cost += factor ? (t.kind == TK_PHASE ? 2 : 1) * popc(t.mask)
                                       : (t.kind <= TK_SHI ? 3 * kPairs : (t.kind == TK_GEN ? 8 : 2) * popc(t.mask));
                    }
                    std::fprintf(stderr, "  pass %zu: %zu ops in %zu layouts [shr %d shi %d gen %d perm %d phase %d scale %d | generic-mask %d mux %d | ~%d ffma2/thread], tile bits",
                                 i, fp.ops.size(), fp.segs.size(), km[TK_SHR], km[TK_SHI], km[TK_GEN], km[TK_PERM_R] + km[TK_PERM_I],
                                 km[TK_PHASE] + km[TK_PHASE_N], km[TK_SCALE_R] + km[TK_SCALE_I], generic, muxed, cost);
                    for (int j = 0; j < fp.tile.n; ++j) std::fprintf(stderr, " %d", fp.tile.pos[j]);
                    std::fprintf(stderr, "\n");
                }
        }
    } else {
        uint64_t launches = 0;
        for (const CanonOp& c : p->ops) launches += c.identity ? 0 : 1;
        p->info.n_launches = launches;
        p->info.n_fused_passes = 0;
        p->info.n_single_ops = n_ops;
        p->info.bytes_planned = bytes;
        p->info.tile_bits = 0;
    }
    *out = p;
    return AQS_OK;
}

static int launch_all(aqs_state_t s, aqs_plan_t p) {
    if (!p->passes.empty()) {
        for (const FusedPass& fp : p->passes) {
            int rc = launch_pass(s->d, fp, s->stream);
            if (rc) return rc;
        }
    } else {
        for (const CanonOp& c : p->ops) {
            int rc = launch_canon(s->d, s->n, c, s->stream);
            if (rc) return rc;
        }
    }
    return AQS_OK;
}

int aqs_plan_run(aqs_state_t s, aqs_plan_t p) {
    if (!s || !p) return fail(AQS_ERR_INVALID, "null handle");
    if (s->n != p->n) return fail(AQS_ERR_INVALID, "plan and state have different qubit counts");
    int up = ensure_uploaded(p);
    if (up) return up;
    p->ran = true;
    if (p->flags & AQS_PLAN_GRAPH) {
        // launch-bound plans (small states, thousands of passes): replay one CUDA graph instead of
        // issuing every launch from the host.  The graph bakes in the state's buffer, so it is
        // re-captured when the plan is run on a different state.
        if (!p->graph || p->graph_state != (const void*)s->d) {
            if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) return fail_cuda(e, "cudaStreamBeginCapture", __LINE__);
            int rc = launch_all(s, p);
            e = cudaStreamEndCapture(s->stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return fail_cuda(e, "cudaStreamEndCapture", __LINE__);
            e = cudaGraphInstantiate(&p->graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return fail_cuda(e, "cudaGraphInstantiate", __LINE__);
            p->graph_state = (const void*)s->d;
        }
        cudaError_t e = cudaGraphLaunch(p->graph, s->stream);
        if (e != cudaSuccess) return fail_cuda(e, "cudaGraphLaunch", __LINE__);
        count_launch(p->info.n_launches);
    } else {
        int rc = launch_all(s, p);
        if (rc) return rc;
    }
    count_ops(p->ops.size());
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, "plan run", __LINE__);
    return AQS_OK;
}

// Sharded run on a flat multi-GPU state (flat.cu): passes [first, first + count) of a fused plan, this
// rank's share of the tiles only.  The caller separates passes whose tiles span GPUs from their
// neighbours with cross-rank barriers on the stream (aqs_plan_pass_span tells which ones do).
int aqs_plan_run_shard(aqs_state_t s, aqs_plan_t p, uint64_t first, uint64_t count, int rank, int log2_world) {
    if (!s || !p) return fail(AQS_ERR_INVALID, "null handle");
    if (s->n != p->n) return fail(AQS_ERR_INVALID, "plan and state have different qubit counts");
    if (p->passes.empty()) return fail(AQS_ERR_INVALID, "sharded runs need a fused plan (AQS_PLAN_FUSE)");
    if (first + count > p->passes.size()) return fail(AQS_ERR_INVALID, "pass range out of bounds");
    if (log2_world < 0 || log2_world > 6 || rank < 0 || rank >= (1 << log2_world)) return fail(AQS_ERR_INVALID, "bad rank");
    if (p->n - p->passes[0].T < log2_world) return fail(AQS_ERR_INVALID, "state too small to shard the tiles of a pass");
    int up = ensure_uploaded(p);
    if (up) return up;
    p->ran = true;
    for (uint64_t i = first; i < first + count; ++i) {
        ShardCut cut;
        int rc = shard_cut(p->n, p->passes[i], rank, log2_world, cut);
        if (rc) return rc;
        rc = launch_pass(s->d, p->passes[i], s->stream, &cut);
        if (rc) return rc;
    }
    if (first == 0) count_ops(p->ops.size());
    return AQS_OK;
}

// One fused pass restricted to the tiles whose number has the bits fix_pos[0 .. fix_n) (ascending positions in the compact
// tile-number space, cf. aqs_plan_shard_cut) equal to those of fix_or, reading the state at `load_base` (nullptr: in place)
// and writing it in place, on `stream` (nullptr: the state's own).  This is what a STAGED pass of a sharded run is made
// of (sharded.py): the tiles of a rank are cut into chunks, each chunk's remote inputs are copied into local staging
// memory by the copy engines while the previous chunk computes, and the kernel writes its results straight to the peers.
int aqs_plan_run_tiles(aqs_state_t s, aqs_plan_t p, uint64_t index, const void* load_base, uint32_t fix_n, const uint8_t* fix_pos, uint32_t fix_or,
                       void* stream) {
    if (!s || !p) return fail(AQS_ERR_INVALID, "null handle");
    if (s->n != p->n) return fail(AQS_ERR_INVALID, "plan and state have different qubit counts");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    if (fix_n > 8 || (fix_n && !fix_pos)) return fail(AQS_ERR_INVALID, "at most 8 pinned tile-number bits");
    const FusedPass& fp = p->passes[index];
    if ((int)fix_n > p->n - fp.T) return fail(AQS_ERR_INVALID, "more pinned bits than the tile number has");
    int up = ensure_uploaded(p);
    if (up) return up;
    p->ran = true;
    ShardCut cut;
    cut.fix_n = fix_n;
    cut.fix_or = fix_or;
    for (uint32_t i = 0; i < fix_n; ++i) {
        if (fix_pos[i] >= p->n - fp.T || (i && fix_pos[i] <= fix_pos[i - 1])) return fail(AQS_ERR_INVALID, "pinned positions must ascend inside the tile number");
        cut.fix_pos[i] = fix_pos[i];
    }
    return launch_pass(s->d, fp, stream ? (cudaStream_t)stream : s->stream, &cut, (float2*)load_base);
}

// the index-bit positions of the tile of fused pass `index`, ascending (pos has room for 16 entries)
int aqs_plan_pass_tile(aqs_plan_t p, uint64_t index, uint8_t* pos, int* tile_bits) {
    if (!p || !pos || !tile_bits) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    const FusedPass& fp = p->passes[index];
    *tile_bits = fp.tile.n;
    for (int j = 0; j < fp.tile.n && j < 16; ++j) pos[j] = fp.tile.pos[j];
    return AQS_OK;
}

// number of rank bits (the top log2_world index bits) inside the tile of fused pass `index`:
// 0 means every tile of the pass lies in ONE shard
int aqs_plan_pass_span(aqs_plan_t p, uint64_t index, int log2_world, int* rank_bits_in_tile) {
    if (!p || !rank_bits_in_tile) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    int c = 0;
    const FusedPass& fp = p->passes[index];
    for (int j = 0; j < fp.tile.n; ++j) c += fp.tile.pos[j] >= p->n - log2_world;
    *rank_bits_in_tile = c;
    return AQS_OK;
}

// Introspection (tests): the tiles rank `rank` of 2^log2_world runs in fused pass `index` — the tile numbers whose
// bits fix_pos[0 .. *fix_n) (ascending) equal those of *fix_or; the launch enumerates the remaining bits.
int aqs_plan_shard_cut(aqs_plan_t p, uint64_t index, int rank, int log2_world, uint32_t* fix_n, uint32_t* fix_or, uint8_t* fix_pos) {
    if (!p || !fix_n || !fix_or || !fix_pos) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    if (log2_world < 0 || log2_world > 6 || rank < 0 || rank >= (1 << log2_world)) return fail(AQS_ERR_INVALID, "bad rank");
    if (p->n - p->passes[index].T < log2_world) return fail(AQS_ERR_INVALID, "state too small to shard the tiles of a pass");
    ShardCut cut;
    int rc = shard_cut(p->n, p->passes[index], rank, log2_world, cut);
    if (rc) return rc;
    *fix_n = cut.fix_n;
    *fix_or = cut.fix_or;
    std::memcpy(fix_pos, cut.fix_pos, sizeof cut.fix_pos);
    return AQS_OK;
}

int aqs_plan_get_info(aqs_plan_t p, aqs_plan_info* info) {
    if (!p || !info) return fail(AQS_ERR_INVALID, "null argument");
    *info = p->info;
    return AQS_OK;
}

// Introspection for tests and tools: the exact launch descriptors of fused pass `index`, as
// { uint32 tile_bits, uint32 n_segs, uint32 n_ops, uint32 has_scale, float2 scale, uint64 n_tiles,
//   BitList tile, uint64 ld_toff[8], ld_roff[5], st_toff[8], st_roff[5], TileSeg segs[n_segs], TileOp ops[n_ops] }.
// Returns the number of bytes the record needs in *needed; copies it when `cap` is large enough.
int aqs_plan_export_pass(aqs_plan_t p, uint64_t index, void* buf, uint64_t cap, uint64_t* needed) {
    if (!p || !needed) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    const FusedPass& fp = p->passes[index];
    struct Head {
        uint32_t tile_bits, n_segs, n_ops, has_scale;
        float2 scale;
        uint64_t n_tiles;
        BitList tile;
        uint64_t ld_toff[kMaxThreadBits], ld_roff[kRegBits], st_toff[kMaxThreadBits], st_roff[kRegBits];
    } h;
    std::memset(&h, 0, sizeof h);
    h.tile_bits = (uint32_t)fp.T; h.n_segs = (uint32_t)fp.segs.size(); h.n_ops = (uint32_t)fp.ops.size();
    h.has_scale = fp.has_scale; h.scale = fp.scale; h.n_tiles = fp.n_tiles; h.tile = fp.tile;
    std::memcpy(h.ld_toff, fp.ld_toff, sizeof h.ld_toff); std::memcpy(h.ld_roff, fp.ld_roff, sizeof h.ld_roff);
    std::memcpy(h.st_toff, fp.st_toff, sizeof h.st_toff); std::memcpy(h.st_roff, fp.st_roff, sizeof h.st_roff);
    const uint64_t bytes = sizeof h + fp.segs.size() * sizeof(TileSeg) + fp.ops.size() * sizeof(TileOp);
    *needed = bytes;
    if (buf && cap >= bytes) {
        char* o = (char*)buf;
        std::memcpy(o, &h, sizeof h); o += sizeof h;
        std::memcpy(o, fp.segs.data(), fp.segs.size() * sizeof(TileSeg)); o += fp.segs.size() * sizeof(TileSeg);
        std::memcpy(o, fp.ops.data(), fp.ops.size() * sizeof(TileOp));
    }
    return AQS_OK;
}

int aqs_plan_pass_source(aqs_plan_t p, uint64_t index, char* buf, uint64_t cap, uint64_t* needed, uint64_t* geom) {
    if (!p || !needed) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    SpecSource s;
    std::string why;
    if (!spec_generate(p->n, p->passes[index], s, why)) return fail(AQS_ERR_STATE, "pass cannot be specialised: " + why);
    *needed = s.src.size() + 1;
    if (geom) { geom[0] = (uint64_t)s.threads; geom[1] = s.smem_bytes; geom[2] = p->passes[index].n_tiles; }
    if (buf && cap >= s.src.size() + 1) std::memcpy(buf, s.src.c_str(), s.src.size() + 1);
    return AQS_OK;
}

int aqs_plan_pass_coefs(aqs_plan_t p, uint64_t index, uint64_t* buf, uint64_t cap, uint64_t* needed) {
    if (!p || !needed) return fail(AQS_ERR_INVALID, "null argument");
    if (index >= p->passes.size()) return fail(AQS_ERR_INVALID, "pass index out of range");
    SpecSource s;
    std::string why;
    if (!spec_generate(p->n, p->passes[index], s, why)) return fail(AQS_ERR_STATE, "pass cannot be specialised: " + why);
    *needed = s.coefs.size();
    if (buf && cap >= s.coefs.size()) std::memcpy(buf, s.coefs.data(), s.coefs.size() * sizeof(uint64_t));
    return AQS_OK;
}

int aqs_plan_jit_ready(aqs_plan_t p, uint64_t* n_ready) {
    if (!p || !n_ready) return fail(AQS_ERR_INVALID, "null argument");
    uint64_t c = 0;
    for (const FusedPass& fp : p->passes) c += (fp.spec && spec_ready(fp)) ? 1 : 0;
    *n_ready = c;
    return AQS_OK;
}

int aqs_jit_wait(void) { return spec_wait_all(); }

int aqs_jit_get_info(aqs_jit_info* out) {
    if (!out) return fail(AQS_ERR_INVALID, "null argument");
    spec_stats(&out->compiled, &out->cache_hits, &out->failed, &out->compile_seconds, &out->pending);
    return AQS_OK;
}

int aqs_plan_destroy(aqs_plan_t p) {
    if (!p) return AQS_OK;
    // The plan may have run on several streams (a cached QCircuit plan is shared by every QSimulator, each with its own
    // non-blocking stream), some of them caller-owned or already destroyed: wait for the whole device before the
    // descriptor arena goes back to the pool, where the next plan would overwrite it under kernels still reading it.
    if (p->ran && p->arena && cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError();
    if (p->graph) cudaGraphExecDestroy(p->graph);
    if (p->arena) pool_free(p->arena, p->arena_bytes);
    delete p;
    return AQS_OK;
}

}  // extern "C"

// specialize.cu — per-pass SPECIALISED tile kernels.
//
// The interpreter in tile_kernel.cuh pays a prelude, a dispatch tree and coefficient decoding for
// every op of a fused pass (ncu, round 1: FMA pipe 40 %, issue slots 56 % on the heaviest pass; the
// arithmetic of a shear op is 48 FFMA2 per thread, the instructions around it about as many).  A pass's
// SHAPE — tile bits, layouts, swizzles, op kinds, target / multiplexing register bits, predicates — is
// known when the plan is built; only the matrix COEFFICIENTS are data.  This file turns a planned pass into
// straight-line CUDA C++ for exactly that shape, compiles it at run time for sm_100a (NVRTC -> cubin ->
// cuModuleLoadData) and caches the kernel by the hash of its source, so that
//   * every op is its bare arithmetic: packed FFMA2 / FMUL2 whose coefficient operands come straight from
//     the kernel-parameter bank through uniform registers (no decode, no dispatch, no descriptor loads);
//   * imaginary shears (the RotX family) are packed too: the (-c, c) operand pairs sit precomputed in
//     the coefficient table and the half swap is an operand modifier;
//   * pair subsets, identity branches of controlled gates and exact permutations are resolved while the
//     source is written: a controlled gate whose control is a register bit touches only its own pairs, an X / CX
//     on register bits is a renaming of variables and costs nothing;
//   * shared-memory re-split addresses are one base per thread plus immediates.
// Coefficients travel as a kernel parameter (a table of packed 64-bit operands), so circuits that differ
// only in their angles share kernels.  The interpreter stays the fallback: a pass runs specialised
// only once its kernel is compiled and loaded (compilation runs on a pool of host threads; aqs_plan_build
// with AQS_PLAN_JIT waits for it, AQS_PLAN_JIT_ASYNC does not).
//
// The generated source is also valid HOST C++ under -DAQS_HOST_EMU (threads = std::thread, __syncthreads
// = a pthread barrier): tests/test_specialize.py compiles it with g++ and checks it against the oracle
// without a GPU, so a GPU mismatch can only come from ptxas.
#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "plan_internal.h"

namespace aqs {

// ---------------------------------------------------------------------------------------------
// 1. source generator
// ---------------------------------------------------------------------------------------------
static const char* kPreamble = R"AQS(
typedef unsigned long long u64;
typedef unsigned int u32;
#ifdef AQS_HOST_EMU
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <pthread.h>
#define DEV static inline
DEV u64 pk(float l, float h) { u32 a, b; memcpy(&a, &l, 4); memcpy(&b, &h, 4); return (u64)a | ((u64)b << 32); }
DEV float lo(u64 v) { u32 a = (u32)v; float f; memcpy(&f, &a, 4); return f; }
DEV float hi(u64 v) { u32 a = (u32)(v >> 32); float f; memcpy(&f, &a, 4); return f; }
DEV u64 fma2(u64 a, u64 b, u64 c) { return pk(fmaf(lo(a), lo(b), lo(c)), fmaf(hi(a), hi(b), hi(c))); }
DEV u64 mul2(u64 a, u64 b) { return pk(lo(a) * lo(b), hi(a) * hi(b)); }
DEV u64 sw(u64 v) { return (v >> 32) | (v << 32); }
#define LDG(p) (*(p))
#define STG(p, v) (*(p) = (v))
#else
#define DEV __device__ __forceinline__
DEV u64 pk(float l, float h) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(l), "f"(h)); return r; }
DEV float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
DEV float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
DEV u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
DEV u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
DEV u64 sw(u64 v) { return pk(hi(v), lo(v)); }      // (re, im) -> (im, re): an operand modifier in SASS
DEV u64 ldg_(const u64* p) { u64 r; asm volatile("ld.global.b64 %0, [%1];" : "=l"(r) : "l"(p) : "memory"); return r; }
DEV void stg_(u64* p, u64 v) { asm volatile("st.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
#define LDG(p) ldg_(p)
#define STG(p, v) stg_(p, v)
#endif
)AQS";

static uint64_t pack2(float l, float h) {
    uint32_t a, b;
    std::memcpy(&a, &l, 4);
    std::memcpy(&b, &h, 4);
    return (uint64_t)a | ((uint64_t)b << 32);
}

static uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

namespace {

struct Gen {
    int n, T, TB;
    const FusedPass& fp;
    std::string o;                       // kernel body
    std::vector<uint64_t> coefs;
    int phys[kRegs];                     // logical register k lives in variable a<phys[k]>
    int tmp = 0;

    Gen(int n_, const FusedPass& f) : n(n_), T(f.T), TB(f.T - kRegBits), fp(f) {
        for (int k = 0; k < kRegs; ++k) phys[k] = k;
    }
    void P(const char* fmt, ...) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        const int need = vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        if (need < (int)sizeof buf) { o += buf; return; }
        std::vector<char> big((size_t)need + 1);
        va_start(ap, fmt);
        vsnprintf(big.data(), big.size(), fmt, ap);
        va_end(ap);
        o += big.data();
    }
    std::string R(int k) const { return "a" + std::to_string(phys[k]); }
    std::unordered_map<uint64_t, size_t> coef_index;
    std::string C(uint64_t bits) {
        auto it = coef_index.find(bits);
        if (it == coef_index.end()) {
            coefs.push_back(bits);
            it = coef_index.emplace(bits, coefs.size() - 1).first;
        }
        return "c" + std::to_string(it->second);
    }
    std::string BC(float v) { return C(pack2(v, v)); }         // (v, v): real factor
    std::string IM(float v) { return C(pack2(-v, v)); }        // (-v, v): with sw(): multiplication by i*v
    std::string fresh(const char* stem) { return std::string(stem) + std::to_string(tmp++); }

    // predicate of an op as a C expression ("" when the op has none)
    std::string pred_expr(const TileOp& t) const {
        std::string e;
        char buf[96];
        if (t.b_mask) {
            snprintf(buf, sizeof buf, "((tile_no & 0x%xu) == 0x%xu)", t.b_mask, t.b_val);
            e += buf;
        }
        if (t.t_mask) {
            snprintf(buf, sizeof buf, "%s((TID & 0x%xu) == 0x%xu)", e.empty() ? "" : " && ", (unsigned)t.t_mask, (unsigned)t.t_val);
            e += buf;
        }
        return e;
    }

    // ---- deferred scales ---------------------------------------------------------------------------
    // The true value of a register is  sc * stored,  sc = re * i^p  (re > 0, p mod 4), tracked while the source is
    // written.  Every op is linear, so real and imaginary factors — the prescales of reflections (CX folded into a
    // rotation), signs, X / Y / Z-like factors, and the cosine that a rotation has in common on both outputs — are
    // never multiplied into the data: they are folded into the coefficients of the next op that combines two
    // registers.  A rotation is then TWO packed FMAs per amplitude pair,
    //     x' = c (x - t y),  y' = c (y + t x)      (or the anti-diagonal form when |sin| > |cos|),
    // instead of three shears, and a general SU(2) (diagonal gate folded in) four.  Scales are materialised
    // (one FMUL2 per register that differs from the majority) only where the pattern would stop being a function of
    // the register index alone — before a layout change, before an op under a run-time predicate — and the common
    // factor goes to the final store.
    struct Sc {
        double re = 1.0;
        int p = 0;
    };
    static Sc mk(double re, int p) {
        Sc z;
        z.re = re;
        z.p = p & 3;
        if (z.re < 0) { z.re = -z.re; z.p = (z.p + 2) & 3; }
        return z;
    }
    static Sc sc_mul(Sc a, Sc b) { return mk(a.re * b.re, a.p + b.p); }
    static Sc sc_div(Sc a, Sc b) { return mk(a.re / b.re, a.p - b.p + 4); }
    static bool is0(Sc a) { return a.re == 0.0; }
    static bool is1(Sc a) { return a.p == 0 && std::fabs(a.re - 1.0) <= 1e-13; }
    static bool sc_eq(Sc a, Sc b) { return a.p == b.p && std::fabs(a.re - b.re) <= 1e-12 * std::max(a.re, b.re); }
    static std::complex<double> sc_val(Sc a) {
        static const std::complex<double> ip[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        return a.re * ip[a.p];
    }
    Sc sc[kRegs];                             // indexed by PHYSICAL variable
    std::complex<double> gscale{1.0, 0.0};    // common factor of every amplitude of the tile, applied at the final store

    // coefficient operand for  k * r  as one packed multiply / FMA operand pair: (coef, operand expression)
    std::string coef_of(Sc k) {
        switch (k.p) {
            case 0: return BC((float)k.re);
            case 2: return BC((float)-k.re);
            case 1: return IM((float)k.re);
            default: return IM((float)-k.re);
        }
    }
    static std::string opnd_of(Sc k, const std::string& r) { return (k.p & 1) ? "sw(" + r + ")" : r; }
    std::string V(int v) const { return "a" + std::to_string(v); }

    void rescale_var(int v, Sc target) {
        const Sc r = sc_div(sc[v], target);
        if (!is1(r)) {
            const std::string c = coef_of(r);
            P("  a%d = mul2(%s, %s);\n", v, c.c_str(), opnd_of(r, V(v)).c_str());
        }
        sc[v] = target;
    }
    Sc majority(const std::vector<int>& vars) const {
        Sc best = sc[vars[0]];
        int best_n = 0;
        for (int v : vars) {
            int c = 0;
            for (int w : vars) c += sc_eq(sc[v], sc[w]);
            if (c > best_n) { best_n = c; best = sc[v]; }
        }
        return best;
    }
    // give the listed variables one common scale (which commutes with any linear op on them)
    void equalize(const std::vector<int>& vars) {
        if (vars.empty()) return;
        const Sc m = majority(vars);
        for (int v : vars) rescale_var(v, m);
    }
    // before a layout change / the final store: every register to the majority scale, which becomes global
    void materialize_all() {
        std::vector<int> all(kRegs);
        for (int k = 0; k < kRegs; ++k) all[k] = k;
        const Sc m = majority(all);
        for (int v : all) rescale_var(v, m);
        gscale *= sc_val(m);
        for (int v : all) sc[v] = Sc();
    }

    // ---- shears -------------------------------------------------------------------------------
    struct ShearSet {
        float a, b, g, sx, sy, qy;
        bool imag, py, cy;
        bool identity() const { return a == 0.f && b == 0.f && g == 0.f && (!py || (sx == 1.f && sy == 1.f && !imag && (!cy || qy == 0.f))); }
        // same instruction sequence for both sets (so that a per-thread select of the coefficients suffices)?
        bool same_form(const ShearSet& z) const { return py == z.py && cy == z.cy && imag == z.imag; }
    };
    static void pair_regs(int tk, int p, int& k0, int& k1) {
        k0 = ((p >> tk) << (tk + 1)) | (p & ((1 << tk) - 1));
        k1 = k0 | (1 << tk);
    }

    // -- fast form: M = N(a, b, g) * diag(px, py) on a pair whose registers carry deferred scales sx0, sy0
    struct Fast {
        bool need_f = false;             // the factor on y is a general complex number: multiplied in explicitly
        std::complex<double> f{1, 0};
        bool anti = false;               // pivot on the anti-diagonal: the outputs land in each other's variables
        Sc k1, k2;                       // A += k1 * B_old,  B += k2 * A_old   (A, B = x, y variables; swapped when anti)
        Sc ox, oy;                       // deferred scales of the new logical x / y
    };
    static Fast fast_form(const ShearSet& s, bool shi, Sc sx0, Sc sy0) {
        Fast F;
        Sc fx = Sc(), fy = Sc();
        if (s.py) {
            fx = s.imag ? mk(s.sx, 1) : mk(s.sx, 0);
            if (s.cy) {
                if (s.qy == 0.f) fy = mk(s.sy, 0);
                else if (s.sy == 0.f) fy = mk(s.qy, 1);
                else { F.need_f = true; F.f = std::complex<double>(s.sy, s.qy); }
            } else {
                fy = s.imag ? mk(s.sy, 1) : mk(s.sy, 0);
            }
        }
        const double a = s.a, b = s.b, g = s.g;
        Sc n00, n01, n10, n11;
        if (!shi) { n00 = mk(1 + a * b, 0); n01 = mk(a + g + a * b * g, 0); n10 = mk(b, 0); n11 = mk(1 + b * g, 0); }
        else { n00 = mk(1 - a * b, 0); n01 = mk(a + g - a * b * g, 1); n10 = mk(b, 1); n11 = mk(1 - b * g, 0); }
        const Sc ex = sc_mul(sx0, fx), ey = sc_mul(sy0, fy);
        const Sc E00 = sc_mul(n00, ex), E01 = sc_mul(n01, ey), E10 = sc_mul(n10, ex), E11 = sc_mul(n11, ey);
        if (E00.re * E11.re >= E01.re * E10.re) {
            F.anti = false;
            F.k1 = is0(E01) ? E01 : sc_div(E01, E00);
            F.k2 = is0(E10) ? E10 : sc_div(E10, E11);
            F.ox = E00; F.oy = E11;
        } else {
            // x' = E01 (y + (E00/E01) x),  y' = E10 (x + (E11/E10) y)
            F.anti = true;
            F.k1 = is0(E00) ? E00 : sc_div(E00, E01);
            F.k2 = is0(E11) ? E11 : sc_div(E11, E10);
            F.ox = E01; F.oy = E10;
        }
        return F;
    }
    struct FastNames { std::string fr, fi, c1, c2; };
    FastNames fast_names(const Fast& F) {
        FastNames nm;
        if (F.need_f) { nm.fr = BC((float)F.f.real()); nm.fi = IM((float)F.f.imag()); }
        if (!is0(F.k1)) nm.c1 = coef_of(F.k1);
        if (!is0(F.k2)) nm.c2 = coef_of(F.k2);
        return nm;
    }
    static bool fast_same_shape(const Fast& A, const Fast& B) {
        return A.need_f == B.need_f && A.anti == B.anti && is0(A.k1) == is0(B.k1) && is0(A.k2) == is0(B.k2) &&
               ((A.k1.p ^ B.k1.p) & 1) == 0 && ((A.k2.p ^ B.k2.p) & 1) == 0;
    }
    // emits the pair update; logical registers k0 (x), k1 (y).  `track`: update names and deferred scales
    void fast_pair(int k0, int k1, const Fast& F, const FastNames& nm, bool track) {
        const int vx = phys[k0], vy = phys[k1];
        if (F.need_f) P("  a%d = fma2(%s, sw(a%d), mul2(%s, a%d));", vy, nm.fi.c_str(), vy, nm.fr.c_str(), vy);
        const int va = F.anti ? vy : vx, vb = F.anti ? vx : vy;       // A += k1 * B_old, B += k2 * A_old
        const bool z1 = is0(F.k1), z2 = is0(F.k2);
        if (!z1 && !z2)
            P("  { const u64 t_ = a%d; a%d = fma2(%s, %s, a%d); a%d = fma2(%s, %s, a%d); }\n", va, va, nm.c1.c_str(), opnd_of(F.k1, V(vb)).c_str(), va,
              vb, nm.c2.c_str(), opnd_of(F.k2, "t_").c_str(), vb);
        else if (!z1) P("  a%d = fma2(%s, %s, a%d);\n", va, nm.c1.c_str(), opnd_of(F.k1, V(vb)).c_str(), va);
        else if (!z2) P("  a%d = fma2(%s, %s, a%d);\n", vb, nm.c2.c_str(), opnd_of(F.k2, V(va)).c_str(), vb);
        else if (F.need_f) P("\n");
        if (!track) return;
        if (F.anti) {
            std::swap(phys[k0], phys[k1]);         // logical x now lives in what was y's variable
            sc[vy] = F.ox;
            sc[vx] = F.oy;
        } else {
            sc[vx] = F.ox;
            sc[vy] = F.oy;
        }
    }

    // -- self-contained form (three shears, explicit prescales): under run-time predicates
    struct ShearNames { std::string a, b, g, sx, sy, qy; };
    // force_sx / force_sy: emit the real prescale even when it is 1 (the other set of a per-thread select needs it)
    ShearNames names_for(const ShearSet& s, bool shi, bool force_sx = false, bool force_sy = false) {
        ShearNames nm;
        nm.a = shi ? IM(s.a) : BC(s.a);
        nm.b = shi ? IM(s.b) : BC(s.b);
        nm.g = shi ? IM(s.g) : BC(s.g);
        if (s.py) {
            if (s.imag) nm.sx = IM(s.sx);
            else if (s.sx != 1.f || force_sx) nm.sx = BC(s.sx);
            if (s.cy) { nm.sy = BC(s.sy); nm.qy = IM(s.qy); }
            else if (s.imag) nm.sy = IM(s.sy);
            else if (s.sy != 1.f || force_sy) nm.sy = BC(s.sy);
        }
        return nm;
    }
    void shear_pair(int k0, int k1, const ShearSet& s, const ShearNames& nm, bool shi) {
        const std::string x = R(k0), y = R(k1);
        if (s.py) {
            if (s.imag) P("  %s = mul2(%s, sw(%s));", x.c_str(), nm.sx.c_str(), x.c_str());
            else if (!nm.sx.empty()) P("  %s = mul2(%s, %s);", x.c_str(), nm.sx.c_str(), x.c_str());
            if (s.cy) P("  %s = fma2(%s, sw(%s), mul2(%s, %s));", y.c_str(), nm.qy.c_str(), y.c_str(), nm.sy.c_str(), y.c_str());
            else if (s.imag) P("  %s = mul2(%s, sw(%s));", y.c_str(), nm.sy.c_str(), y.c_str());
            else if (!nm.sy.empty()) P("  %s = mul2(%s, %s);", y.c_str(), nm.sy.c_str(), y.c_str());
        }
        if (!shi) {
            P("  %s = fma2(%s, %s, %s); %s = fma2(%s, %s, %s); %s = fma2(%s, %s, %s);\n", x.c_str(), nm.a.c_str(), y.c_str(), x.c_str(),
              y.c_str(), nm.b.c_str(), x.c_str(), y.c_str(), x.c_str(), nm.g.c_str(), y.c_str(), x.c_str());
        } else {
            P("  %s = fma2(%s, sw(%s), %s); %s = fma2(%s, sw(%s), %s); %s = fma2(%s, sw(%s), %s);\n", x.c_str(), nm.a.c_str(), y.c_str(),
              x.c_str(), y.c_str(), nm.b.c_str(), x.c_str(), y.c_str(), x.c_str(), nm.g.c_str(), y.c_str(), x.c_str());
        }
    }
    std::vector<int> pair_vars(int tk, uint32_t pair_mask) const {
        std::vector<int> v;
        for (int p = 0; p < kPairs; ++p) {
            if (!(pair_mask >> p & 1u)) continue;
            int k0, k1;
            pair_regs(tk, p, k0, k1);
            v.push_back(phys[k0]);
            v.push_back(phys[k1]);
        }
        return v;
    }
    void emit_shear(const TileOp& t) {
        const bool shi = t.kind == TK_SHI;
        const bool py = (t.flags & TF_PY) != 0, cy = (t.flags & TF_CY) != 0;
        const bool mux = (t.flags & TF_MUX) != 0, regmux = (t.flags & TF_REGMUX) != 0;
        ShearSet A{t.a[0], t.a[1], t.a[2], t.sx[0], t.a[3], t.qy[0], (t.flags & TF_IMAG_A) != 0, py, cy};
        ShearSet B{t.b[0], t.b[1], t.b[2], t.sx[1], t.b[3], t.qy[1], (t.flags & TF_IMAG_B) != 0, py, cy};
        if (!py) { A.sx = A.sy = B.sx = B.sy = 1.f; A.qy = B.qy = 0.f; A.imag = B.imag = false; }
        if (!cy) { A.qy = B.qy = 0.f; }
        if (!(mux || regmux)) B = A;                        // plain op: one set everywhere
        const std::string pe = pred_expr(t);
        const int tk = t.tk;
        P("  // %s tk=%d%s%s%s%s\n", shi ? "SHI" : "SHR", tk, mux ? " mux" : "", regmux ? " regmux" : "", py ? " py" : "", cy ? " cy" : "");
        if (pe.empty()) {
            // pair subsets known here: fast forms, deferred scales per register
            for (int p = 0; p < kPairs; ++p) {
                int k0, k1;
                pair_regs(tk, p, k0, k1);
                const ShearSet& s = (regmux && !((p >> t.mj) & 1)) ? B : A;
                if (s.identity()) continue;
                const Fast F = fast_form(s, shi, sc[phys[k0]], sc[phys[k1]]);
                fast_pair(k0, k1, F, fast_names(F), true);
            }
            return;
        }
        // run-time predicate: the op's registers first get one common scale (it commutes with the op)
        equalize(pair_vars(tk, 0xffffu));
        auto all_pairs = [&](const ShearSet& s, const ShearNames& nm) {
            for (int p = 0; p < kPairs; ++p) { int k0, k1; pair_regs(tk, p, k0, k1); shear_pair(k0, k1, s, nm, shi); }
        };
        if (mux) {
            // the multiplexing bit is a thread or tile-number bit: predicate true -> set a, false -> set b
            const bool lane_pred = (t.t_mask & 31u) != 0;
            // fast forms when both branches leave the SAME deferred scales behind (a rotation and the same rotation
            // times X: CX folded in); the two branches then differ only in coefficients and, maybe, in the pivot
            if (!A.identity() && !B.identity()) {
                const Sc one = Sc();
                const Fast FA = fast_form(A, shi, one, one), FB = fast_form(B, shi, one, one);
                if (sc_eq(FA.ox, FB.ox) && sc_eq(FA.oy, FB.oy)) {
                    const Sc base = sc[phys[0]] ;     // (every register of the op has this scale now)
                    (void)base;
                    auto finish = [&](const Fast& F) {
                        // both outputs of every pair: scale *= (ox, oy); names swap when the pivot is anti-diagonal
                        for (int p = 0; p < kPairs; ++p) {
                            int k0, k1;
                            pair_regs(tk, p, k0, k1);
                            const int vx = phys[k0], vy = phys[k1];
                            const Sc s0 = sc[vx];
                            if (F.anti) { std::swap(phys[k0], phys[k1]); sc[vy] = sc_mul(s0, F.ox); sc[vx] = sc_mul(s0, F.oy); }
                            else { sc[vx] = sc_mul(s0, F.ox); sc[vy] = sc_mul(s0, F.oy); }
                        }
                    };
                    if (fast_same_shape(FA, FB) && lane_pred) {
                        const std::string ok = fresh("ok");
                        P("  { const bool %s = %s;\n", ok.c_str(), pe.c_str());
                        const FastNames na = fast_names(FA), nb = fast_names(FB);
                        FastNames nm;
                        auto sel = [&](const std::string& a, const std::string& b) {
                            if (a.empty()) return std::string();
                            const std::string v = fresh("k");
                            P("  const u64 %s = %s ? %s : %s;\n", v.c_str(), ok.c_str(), a.c_str(), b.c_str());
                            return v;
                        };
                        nm.fr = sel(na.fr, nb.fr); nm.fi = sel(na.fi, nb.fi); nm.c1 = sel(na.c1, nb.c1); nm.c2 = sel(na.c2, nb.c2);
                        for (int p = 0; p < kPairs; ++p) { int k0, k1; pair_regs(tk, p, k0, k1); fast_pair(k0, k1, FA, nm, false); }
                        P("  }\n");
                        finish(FA);
                        return;
                    }
                    if (FA.anti == FB.anti) {
                        // same variables end up holding x and y in both branches: branch on the predicate
                        const FastNames na = fast_names(FA), nb = fast_names(FB);
                        P("  if (%s) {\n", pe.c_str());
                        for (int p = 0; p < kPairs; ++p) { int k0, k1; pair_regs(tk, p, k0, k1); fast_pair(k0, k1, FA, na, false); }
                        P("  } else {\n");
                        for (int p = 0; p < kPairs; ++p) { int k0, k1; pair_regs(tk, p, k0, k1); fast_pair(k0, k1, FB, nb, false); }
                        P("  }\n");
                        finish(FA);
                        return;
                    }
                }
            }
            if (lane_pred && A.same_form(B) && !A.identity() && !B.identity()) {
                // per-thread select of the coefficients, one instruction sequence (no divergence)
                const std::string ok = fresh("ok");
                P("  { const bool %s = %s;\n", ok.c_str(), pe.c_str());
                const bool fsx = A.sx != 1.f || B.sx != 1.f, fsy = A.sy != 1.f || B.sy != 1.f;
                const ShearNames na = names_for(A, shi, fsx, fsy), nb = names_for(B, shi, fsx, fsy);
                ShearNames nm;
                auto sel = [&](const std::string& a, const std::string& b) {
                    if (a.empty()) return std::string();
                    const std::string v = fresh("k");
                    P("  const u64 %s = %s ? %s : %s;\n", v.c_str(), ok.c_str(), a.c_str(), b.c_str());
                    return v;
                };
                nm.a = sel(na.a, nb.a); nm.b = sel(na.b, nb.b); nm.g = sel(na.g, nb.g);
                nm.sx = sel(na.sx, nb.sx); nm.sy = sel(na.sy, nb.sy); nm.qy = sel(na.qy, nb.qy);
                all_pairs(A, nm);
                P("  }\n");
            } else {
                P("  if (%s) {\n", pe.c_str());
                if (!A.identity()) all_pairs(A, names_for(A, shi));
                P("  } else {\n");
                if (!B.identity()) all_pairs(B, names_for(B, shi));
                P("  }\n");
            }
            return;
        }
        P("  if (%s) {\n", pe.c_str());
        if (regmux) {
            const ShearNames na = A.identity() ? ShearNames() : names_for(A, shi), nb = B.identity() ? ShearNames() : names_for(B, shi);
            for (int p = 0; p < kPairs; ++p) {
                int k0, k1;
                pair_regs(tk, p, k0, k1);
                const bool use_a = (p >> t.mj) & 1;
                const ShearSet& s = use_a ? A : B;
                if (s.identity()) continue;
                shear_pair(k0, k1, s, use_a ? na : nb, shi);
            }
        } else if (!A.identity()) {
            all_pairs(A, names_for(A, shi));
        }
        P("  }\n");
    }

    // ---- direct kinds ---------------------------------------------------------------------------
    void emit_direct(const TileOp& t) {
        const std::string pe = pred_expr(t);
        const int tk = t.tk;
        if (t.kind == TK_PERM_R || t.kind == TK_PERM_I) {
            const bool imag = t.kind == TK_PERM_I;
            const float c0 = t.a[0], c1 = t.a[1];
            P("  // PERM_%c tk=%d mask=%04x\n", imag ? 'I' : 'R', tk, t.mask & 0xffffu);
            if (pe.empty()) {
                // x' = f0 * y, y' = f1 * x: swap the NAMES, the factors go to the deferred scales (exact data movement costs nothing)
                for (int p = 0; p < kPairs; ++p) {
                    if (!(t.mask >> p & 1u)) continue;
                    int k0, k1;
                    pair_regs(tk, p, k0, k1);
                    const int vx = phys[k0], vy = phys[k1];
                    sc[vy] = sc_mul(sc[vy], mk(c0, imag ? 1 : 0));
                    sc[vx] = sc_mul(sc[vx], mk(c1, imag ? 1 : 0));
                    std::swap(phys[k0], phys[k1]);
                }
            } else {
                equalize(pair_vars(tk, t.mask & 0xffffu));
                const bool plain = !imag && c0 == 1.f && c1 == 1.f;
                const std::string f0 = plain ? "" : (imag ? IM(c0) : BC(c0)), f1 = plain ? "" : (imag ? IM(c1) : BC(c1));
                P("  if (%s) {\n", pe.c_str());
                for (int p = 0; p < kPairs; ++p) {
                    if (!(t.mask >> p & 1u)) continue;
                    int k0, k1;
                    pair_regs(tk, p, k0, k1);
                    const std::string x = R(k0), y = R(k1);
                    if (plain) P("  { const u64 t_ = %s; %s = %s; %s = t_; }\n", x.c_str(), x.c_str(), y.c_str(), y.c_str());
                    else if (imag) P("  { const u64 t_ = %s; %s = mul2(%s, sw(%s)); %s = mul2(%s, sw(t_)); }\n", x.c_str(), x.c_str(), f0.c_str(), y.c_str(), y.c_str(), f1.c_str());
                    else P("  { const u64 t_ = %s; %s = mul2(%s, %s); %s = mul2(%s, t_); }\n", x.c_str(), x.c_str(), f0.c_str(), y.c_str(), y.c_str(), f1.c_str());
                }
                P("  }\n");
            }
            return;
        }
        // TK_GEN: x' = m00 x + m01 y, y' = m10 x + m11 y
        P("  // GEN tk=%d mask=%04x\n", tk, t.mask & 0xffffu);
        equalize(pair_vars(tk, t.mask & 0xffffu));
        if (!pe.empty()) P("  if (%s) {\n", pe.c_str());
        std::string re[4], im[4];
        for (int i = 0; i < 4; ++i) { re[i] = BC(t.a[2 * i]); im[i] = IM(t.a[2 * i + 1]); }
        for (int p = 0; p < kPairs; ++p) {
            if (!(t.mask >> p & 1u)) continue;
            int k0, k1;
            pair_regs(tk, p, k0, k1);
            const std::string x = R(k0), y = R(k1);
            P("  { const u64 x0 = %s, y0 = %s;\n", x.c_str(), y.c_str());
            P("    %s = fma2(%s, sw(y0), fma2(%s, y0, fma2(%s, sw(x0), mul2(%s, x0))));\n", x.c_str(), im[1].c_str(), re[1].c_str(), im[0].c_str(), re[0].c_str());
            P("    %s = fma2(%s, sw(y0), fma2(%s, y0, fma2(%s, sw(x0), mul2(%s, x0)))); }\n", y.c_str(), im[3].c_str(), re[3].c_str(), im[2].c_str(), re[2].c_str());
        }
        if (!pe.empty()) P("  }\n");
    }

    // ---- factors ----------------------------------------------------------------------------------
    void emit_factor(const TileOp& t) {
        const std::string pe = pred_expr(t);
        P("  // FACTOR kind=%d mask=%08x\n", (int)t.kind, t.mask);
        if (t.kind == TK_SCALE_R || t.kind == TK_SCALE_I) {
            if (pe.empty()) {
                // a real or imaginary factor on whole registers: deferred, no instruction
                for (int k = 0; k < kRegs; ++k)
                    if (t.mask >> k & 1u) sc[phys[k]] = sc_mul(sc[phys[k]], mk(t.a[0], t.kind == TK_SCALE_I ? 1 : 0));
                return;
            }
            P("  if (%s) {\n", pe.c_str());
            const std::string f = t.kind == TK_SCALE_R ? BC(t.a[0]) : IM(t.a[0]);
            for (int k = 0; k < kRegs; ++k) {
                if (!(t.mask >> k & 1u)) continue;
                if (t.kind == TK_SCALE_R) P("  %s = mul2(%s, %s);\n", R(k).c_str(), f.c_str(), R(k).c_str());
                else P("  %s = mul2(%s, sw(%s));\n", R(k).c_str(), f.c_str(), R(k).c_str());
            }
            P("  }\n");
            return;
        }
        if (!pe.empty()) P("  if (%s) {\n", pe.c_str());
        {
            // e^{i theta} as three shears on (re, im); TK_PHASE_N starts from the negated amplitude
            const std::string c = C(pack2(t.a[0], t.a[1]));
            P("  { const float fr = lo(%s), fi = hi(%s);\n", c.c_str(), c.c_str());
            for (int k = 0; k < kRegs; ++k) {
                if (!(t.mask >> k & 1u)) continue;
                const std::string r = R(k);
                if (t.kind == TK_PHASE)
                    P("    { float xr = lo(%s), xi = hi(%s); xr = fmaf(fr, xi, xr); xi = fmaf(fi, xr, xi); xr = fmaf(fr, xi, xr); %s = pk(xr, xi); }\n",
                      r.c_str(), r.c_str(), r.c_str());
                else
                    P("    { float xr = lo(%s), xi = hi(%s); xr = fmaf(-fr, xi, -xr); xi = fmaf(fi, xr, -xi); xr = fmaf(fr, xi, xr); %s = pk(xr, xi); }\n",
                      r.c_str(), r.c_str(), r.c_str());
            }
            P("  }\n");
        }
        if (!pe.empty()) P("  }\n");
    }

    // ---- ladders ----------------------------------------------------------------------------------
    // header, register-control record, n_cont records with four thread / tile-number controls each (tile_kernel.cuh)
    size_t emit_ladder(const TileOp* t, size_t avail) {
        uint32_t n_cont;
        std::memcpy(&n_cont, &t[0].sx[0], sizeof n_cont);
        if (avail < 2u + n_cont) return 0;
        const TileOp& reg = t[1];
        const std::string pe = pred_expr(t[0]);
        const int mj = t[0].mj;
        P("  // LADDER mj=%d cont=%u\n", mj, n_cont);
        if (!pe.empty()) P("  if (%s) {\n", pe.c_str());
        const std::string f0 = C(pack2(t[0].a[0], t[0].a[1]));
        P("  { float fr = lo(%s), fi = hi(%s);\n", f0.c_str(), f0.c_str());
        for (uint32_t c = 0; c < n_cont; ++c) {
            const TileOp& cr = t[2 + c];
            for (int q = 0; q < 4; ++q) {
                const uint32_t code = (cr.mask >> (8 * q)) & 0xffu;
                if (code == 0x3fu) continue;
                const std::string w = C(pack2(cr.a[2 * q], cr.a[2 * q + 1]));
                P("    if ((%s >> %u) & 1u) { const float cs = lo(%s), sn = hi(%s); const float t_ = fr * cs - fi * sn; fi = fmaf(fr, sn, fi * cs); fr = t_; }\n",
                  (code & 0x20u) ? "tile_no" : "TID", code & 0x1fu, w.c_str(), w.c_str());
            }
        }
        float wr[kRegBits], wi[kRegBits];
        for (int r = 0; r < 4; ++r) { wr[r] = reg.a[2 * r]; wi[r] = reg.a[2 * r + 1]; }
        wr[4] = reg.sx[0]; wi[4] = reg.sx[1];
        std::string wn[kRegBits];
        for (int r = 0; r < kRegBits; ++r)
            if (!(wr[r] == 1.f && wi[r] == 0.f)) {
                wn[r] = fresh("w");
                const std::string c = C(pack2(wr[r], wi[r]));
                P("    const float %sr = lo(%s), %si = hi(%s);\n", wn[r].c_str(), c.c_str(), wn[r].c_str(), c.c_str());
            }
        const int NB = (mj < 5) ? kRegBits - 1 : kRegBits;
        for (int i = 0; i < (1 << NB); ++i) {
            const int g = i ^ (i >> 1);
            if (i) {
                const int cb = __builtin_ctz(i);
                const int r = (mj < 5 && cb >= mj) ? cb + 1 : cb;
                if (!wn[r].empty()) {
                    const bool enter = (g >> cb) & 1;
                    // entering the bit: * w_r, leaving it: * conj(w_r)
                    P("    { const float s_ = %s%si; const float t_ = fr * %sr - fi * s_; fi = fmaf(fr, s_, fi * %sr); fr = t_; }\n", enter ? "" : "-",
                      wn[r].c_str(), wn[r].c_str(), wn[r].c_str());
                }
            }
            const int k = (mj < 5) ? ((((g >> mj) << (mj + 1)) | (g & ((1 << mj) - 1))) | (1 << mj)) : g;
            const std::string rk = R(k);
            P("    { const float xr = lo(%s), xi = hi(%s); %s = pk(xr * fr - xi * fi, fmaf(xr, fi, xi * fr)); }\n", rk.c_str(), rk.c_str(), rk.c_str());
        }
        P("  }\n");
        if (!pe.empty()) P("  }\n");
        return 2u + n_cont;
    }

    // ---- layout changes through shared memory ---------------------------------------------------------
    void emit_cols(const char* name, const uint16_t* tcol) {
        P("  u32 %s = 0;\n", name);
        for (int j = 0; j < TB; ++j)
            if (tcol[j]) P("  %s ^= (0u - ((TID >> %d) & 1u)) & 0x%xu;\n", name, j, (unsigned)tcol[j]);
    }
    void emit_resplit(const TileSeg& sg) {
        uint32_t thi_w = 0, thi_r = 0;
        for (int j = 0; j < TB; ++j) { thi_w |= sg.wr_tcol[j] & ~15u; thi_r |= sg.rd_tcol[j] & ~15u; }
        auto reg_col = [&](const uint16_t* rcol, int k) {
            uint32_t c = 0;
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) c ^= rcol[i];
            return c;
        };
        materialize_all();
        P("  SYNC();\n  {\n");
        emit_cols("wb", sg.wr_tcol);
        bool low_w[16] = {false}, low_r[16] = {false};
        bool disjoint_w = true, disjoint_r = true;
        for (int k = 0; k < kRegs; ++k) {
            const uint32_t cw = reg_col(sg.wr_rcol, k), cr = reg_col(sg.rd_rcol, k);
            low_w[cw & 15u] = true; low_r[cr & 15u] = true;
            if ((cw & ~15u) & thi_w) disjoint_w = false;
            if ((cr & ~15u) & thi_r) disjoint_r = false;
        }
        // slot = wb ^ C_k; the high part of C_k never meets a thread bit, so it is an immediate offset
        if (disjoint_w) {
            for (int v = 0; v < 16; ++v) if (low_w[v]) P("  const u32 wb%d = wb ^ %du;\n", v, v);
            for (int k = 0; k < kRegs; ++k) {
                const uint32_t c = reg_col(sg.wr_rcol, k);
                P("  sm[wb%u + %uu] = %s;\n", c & 15u, c & ~15u, R(k).c_str());
            }
        } else {
            for (int k = 0; k < kRegs; ++k) P("  sm[wb ^ %uu] = %s;\n", reg_col(sg.wr_rcol, k), R(k).c_str());
        }
        P("  }\n  SYNC();\n  {\n");
        for (int k = 0; k < kRegs; ++k) phys[k] = k;
        emit_cols("rb", sg.rd_tcol);
        if (disjoint_r) {
            for (int v = 0; v < 16; ++v) if (low_r[v]) P("  const u32 rb%d = rb ^ %du;\n", v, v);
            for (int k = 0; k < kRegs; ++k) {
                const uint32_t c = reg_col(sg.rd_rcol, k);
                P("  a%d = sm[rb%u + %uu];\n", k, c & 15u, c & ~15u);
            }
        } else {
            for (int k = 0; k < kRegs; ++k) P("  a%d = sm[rb ^ %uu];\n", k, reg_col(sg.rd_rcol, k));
        }
        P("  }\n");
    }

    // ---- the whole kernel -------------------------------------------------------------------------------
    bool run(std::string& why_not) {
        if (T < kMinTileBits || T > kMaxTileBits) { why_not = "tile size"; return false; }
        if (fp.tile.n != T) { why_not = "tile list"; return false; }
        // tile number -> base index: the non-tile bits of the index, in runs
        P("  u32 tile_no = BID;\n");
        P("  for (u32 i = 0; i < rt.fix_n; ++i) { const u32 p = rt.fix_pos[i]; tile_no = ((tile_no >> p) << (p + 1)) | (tile_no & ((1u << p) - 1u)); }\n");
        P("  tile_no |= rt.fix_or;\n");
        P("  u64 gbase = 0;\n");
        {
            uint64_t tile = 0;
            for (int j = 0; j < T; ++j) tile |= 1ull << fp.tile.pos[j];
            int c = 0;
            for (int b = 0; b < n;) {
                if (tile >> b & 1ull) { ++b; continue; }
                int len = 0;
                while (b + len < n && !(tile >> (b + len) & 1ull)) ++len;
                P("  gbase |= (u64)((tile_no >> %d) & 0x%xu) << %d;\n", c, (1u << len) - 1u, b);
                c += len;
                b += len;
            }
        }
        auto io_base = [&](const char* name, const uint64_t* toff) {
            P("  u64* %s = %s + (gbase | (u64)(TID & 31u)", name, std::strcmp(name, "dst") == 0 ? "rt.state_out" : "rt.state");
            for (int j = kLaneBits; j < TB; ++j) P(" | ((u64)((TID >> %d) & 1u) << %d)", j, __builtin_ctzll(toff[j]));
            P(");\n");
        };
        auto reg_off = [&](const uint64_t* roff, int k) {
            uint64_t off = 0;
            for (int i = 0; i < kRegBits; ++i)
                if (k >> i & 1) off += roff[i];
            return off;
        };
        for (int j = kLaneBits; j < TB; ++j)
            if (!fp.ld_toff[j] || (fp.ld_toff[j] & (fp.ld_toff[j] - 1)) || !fp.st_toff[j] || (fp.st_toff[j] & (fp.st_toff[j] - 1))) { why_not = "io offsets"; return false; }
        P("  u64");
        for (int k = 0; k < kRegs; ++k) P("%s a%d", k ? "," : "", k);
        P(";\n  {\n");
        io_base("src", fp.ld_toff);
        for (int k = 0; k < kRegs; ++k) P("  a%d = LDG(src + 0x%llxull);\n", k, (unsigned long long)reg_off(fp.ld_roff, k));
        P("  }\n");

        for (size_t s = 0; s < fp.segs.size(); ++s) {
            const TileSeg& sg = fp.segs[s];
            P("  // ---- segment %zu\n", s);
            if (sg.resplit) emit_resplit(sg);
            const size_t first = sg.first_op, end = first + sg.n_ops;
            if (end > fp.ops.size()) { why_not = "op range"; return false; }
            for (size_t i = first; i < end;) {
                const TileOp& t = fp.ops[i];
                switch (t.kind) {
                    case TK_SHR: case TK_SHI:
                        if (t.mj > 3 && (t.flags & TF_REGMUX)) { why_not = "generic pair mask on a shear"; return false; }
                        emit_shear(t); ++i; break;
                    case TK_GEN: case TK_PERM_R: case TK_PERM_I: emit_direct(t); ++i; break;
                    case TK_PHASE: case TK_SCALE_R: case TK_SCALE_I: case TK_PHASE_N: emit_factor(t); ++i; break;
                    case TK_LADDER: {
                        const size_t used = emit_ladder(&t, end - i);
                        if (!used) { why_not = "ladder records"; return false; }
                        i += used;
                        break;
                    }
                    default: why_not = "unknown op kind"; return false;
                }
            }
        }
        materialize_all();
        P("  {\n");
        io_base("dst", fp.st_toff);
        std::string re, im;
        {
            const std::complex<double> fin = gscale * (fp.has_scale ? std::complex<double>(fp.scale.x, fp.scale.y) : std::complex<double>(1.0, 0.0));
            const bool real_only = std::fabs(fin.imag()) <= 1e-9 * std::abs(fin), imag_only = std::fabs(fin.real()) <= 1e-9 * std::abs(fin);
            if (!(real_only && std::fabs(fin.real() - 1.0) <= 1e-9)) {
                if (!imag_only) re = BC((float)fin.real());
                if (!real_only) im = IM((float)fin.imag());
            }
        }
        for (int k = 0; k < kRegs; ++k) {
            const std::string r = R(k);
            const unsigned long long off = (unsigned long long)reg_off(fp.st_roff, k);
            if (re.empty() && im.empty()) P("  STG(dst + 0x%llxull, %s);\n", off, r.c_str());
            else if (im.empty()) P("  STG(dst + 0x%llxull, mul2(%s, %s));\n", off, re.c_str(), r.c_str());
            else if (re.empty()) P("  STG(dst + 0x%llxull, mul2(%s, sw(%s)));\n", off, im.c_str(), r.c_str());
            else P("  STG(dst + 0x%llxull, fma2(%s, sw(%s), mul2(%s, %s)));\n", off, im.c_str(), r.c_str(), re.c_str(), r.c_str());
        }
        P("  }\n");
        return true;
    }
};

}  // namespace

bool spec_generate(int n, const FusedPass& fp, SpecSource& out, std::string& why_not) {
    Gen g(n, fp);
    if (!g.run(why_not)) return false;
    const int NT = 1 << (fp.T - kRegBits);
    const size_t nc = g.coefs.empty() ? 1 : g.coefs.size();
    if (32 + nc * 8 > 32000) { why_not = "coefficient table exceeds the kernel parameter space"; return false; }
    // The coefficients are individual 64-bit kernel parameters (constant bank -> uniform registers).  One by-value
    // struct holding the table compiles two orders of magnitude slower (measured: cicc 92 s against 2 s for a 388-entry table).
    std::string head;
    {
        char buf[512];
        snprintf(buf, sizeof buf,
                 "#define AQS_NT %d\n#define AQS_NC %zu\n#define AQS_SLOTS %u\n"
                 "struct RT { u64* state; u64* state_out; u32 fix_n, fix_or; unsigned char fix_pos[8]; };\n"
                 "#ifdef AQS_HOST_EMU\n"
                 "#define SYNC() pthread_barrier_wait(bar)\n"
                 "static void aqs_pass_body(const RT& rt, const u64* cc, const u32 TID, const u32 BID, u64* sm, pthread_barrier_t* bar) {\n",
                 NT, nc, 1u << fp.T);
        head += buf;
        for (size_t i = 0; i < nc; ++i) { snprintf(buf, sizeof buf, "  const u64 c%zu = cc[%zu];\n", i, i); head += buf; }
        int minb = tile_min_blocks(fp.T);
        if (const char* e = std::getenv("AQS_JIT_MINB")) minb = std::max(1, std::atoi(e));      // (experiments)
        snprintf(buf, sizeof buf, "#else\n#define SYNC() __syncthreads()\nextern \"C\" __global__ void __launch_bounds__(AQS_NT, %d) aqs_pass(const RT rt", minb);
        head += buf;
        for (size_t i = 0; i < nc; ++i) { snprintf(buf, sizeof buf, ", const u64 c%zu", i); head += buf; }
        head += ") {\n  extern __shared__ __align__(16) u64 sm[];\n  const u32 TID = threadIdx.x, BID = blockIdx.x;\n#endif\n";
    }
    static const char* tail =
        "}\n"
        "#ifdef AQS_HOST_EMU\n"
        "extern \"C\" void aqs_pass_emu_run(u64* state, const u64* coefs, u32 n_ctas, u32 fix_n, u32 fix_or, const unsigned char* fix_pos) {\n"
        "  RT rt; rt.state = state; rt.state_out = state; rt.fix_n = fix_n; rt.fix_or = fix_or;\n"
        "  for (int i = 0; i < 8; ++i) rt.fix_pos[i] = fix_pos ? fix_pos[i] : 0;\n"
        "  std::vector<u64> sm(AQS_SLOTS);\n"
        "  for (u32 b = 0; b < n_ctas; ++b) {\n"
        "    pthread_barrier_t bar; pthread_barrier_init(&bar, nullptr, AQS_NT);\n"
        "    std::vector<std::thread> th;\n"
        "    for (u32 t = 0; t < AQS_NT; ++t) th.emplace_back([&, t]() { aqs_pass_body(rt, coefs, t, b, sm.data(), &bar); });\n"
        "    for (auto& x : th) x.join();\n"
        "    pthread_barrier_destroy(&bar);\n"
        "  }\n"
        "}\n"
        "#endif\n";
    out.src = std::string(kPreamble) + head + g.o + tail;
    out.coefs = std::move(g.coefs);
    if (out.coefs.empty()) out.coefs.push_back(0);
    out.threads = NT;
    out.smem_bytes = sizeof(float2) << fp.T;
    out.key = fnv1a(out.src);
    return true;
}

// ---------------------------------------------------------------------------------------------
// 2. run-time compilation (NVRTC, loaded with dlopen: the engine has no link-time dependency on it)
// ---------------------------------------------------------------------------------------------
namespace {

typedef struct _nvrtcProgram* nvrtcProgram_t;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram_t, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram_t, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram_t, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram_t, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram_t, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram_t*) = nullptr;
    bool ok = false;
    int version = 0;       // major * 1000 + minor
    std::string err;
};
Nvrtc g_nvrtc;
std::once_flag g_nvrtc_once;

void load_nvrtc() {
    // Several NVRTC builds can be visible to one process (a Python environment bundles its own next to the toolkit's,
    // and dlopen by SONAME returns whichever was loaded first).  Code quality differs between them — measured on the
    // brickwork-30 plan: 63.7 ms with NVRTC 12.9, 75.8 ms with the 12.8 build that PyTorch ships — so every candidate is
    // opened and the NEWEST one is used.  AQS_NVRTC_LIB forces a specific file.
    static const char* names[] = {"/usr/local/cuda/lib64/libnvrtc.so.13", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                                  "libnvrtc.so.13", "libnvrtc.so.12", "libnvrtc.so"};
    void* h = nullptr;
    if (const char* e = std::getenv("AQS_NVRTC_LIB")) {
        h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    } else {
        int best = -1;
        for (size_t i = 0; i < sizeof names / sizeof *names; ++i) {
            void* c = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
            if (!c) continue;
            int major = 0, minor = 0;
            auto ver = reinterpret_cast<int (*)(int*, int*)>(dlsym(c, "nvrtcVersion"));
            const int v = (ver && ver(&major, &minor) == 0) ? major * 1000 + minor : 0;
            if (v > best) { best = v; h = c; }
        }
        g_nvrtc.version = best;
    }
    if (!h) { g_nvrtc.err = "libnvrtc.so.12 not found"; return; }
#define AQS_SYM(field, sym) g_nvrtc.field = reinterpret_cast<decltype(g_nvrtc.field)>(dlsym(h, sym)); if (!g_nvrtc.field) { g_nvrtc.err = "missing " sym; return; }
    AQS_SYM(CreateProgram, "nvrtcCreateProgram")
    AQS_SYM(CompileProgram, "nvrtcCompileProgram")
    AQS_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    AQS_SYM(GetCUBIN, "nvrtcGetCUBIN")
    AQS_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    AQS_SYM(GetProgramLog, "nvrtcGetProgramLog")
    AQS_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef AQS_SYM
    g_nvrtc.ok = true;
}

bool compile_cubin(const std::string& src, std::vector<char>& cubin, std::string& log) {
    std::call_once(g_nvrtc_once, load_nvrtc);
    if (!g_nvrtc.ok) { log = g_nvrtc.err; return false; }
    nvrtcProgram_t prog = nullptr;
    if (g_nvrtc.CreateProgram(&prog, src.c_str(), "aqs_pass.cu", 0, nullptr, nullptr) != 0) { log = "nvrtcCreateProgram failed"; return false; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17"};
    const int rc = g_nvrtc.CompileProgram(prog, 3, opts);
    if (rc != 0) {
        size_t ls = 0;
        g_nvrtc.GetProgramLogSize(prog, &ls);
        log.resize(ls);
        if (ls) g_nvrtc.GetProgramLog(prog, &log[0]);
        g_nvrtc.DestroyProgram(&prog);
        return false;
    }
    size_t sz = 0;
    g_nvrtc.GetCUBINSize(prog, &sz);
    cubin.resize(sz);
    const bool ok = sz && g_nvrtc.GetCUBIN(prog, cubin.data()) == 0;
    g_nvrtc.DestroyProgram(&prog);
    if (!ok) log = "nvrtcGetCUBIN failed";
    return ok;
}

// driver entry points (fetched through the runtime: no link-time dependency on libcuda)
struct Drv {
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    bool ok = false, tried = false;
};
Drv g_drv;
std::mutex g_drv_mu;

template <typename F>
bool entry(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}
bool load_driver() {
    std::lock_guard<std::mutex> lk(g_drv_mu);
    if (g_drv.tried) return g_drv.ok;
    g_drv.tried = true;
    g_drv.ok = entry("cuModuleLoadData", g_drv.ModuleLoadData) && entry("cuModuleGetFunction", g_drv.ModuleGetFunction) &&
               entry("cuModuleUnload", g_drv.ModuleUnload) && entry("cuLaunchKernel", g_drv.LaunchKernel) &&
               entry("cuFuncSetAttribute", g_drv.FuncSetAttribute) && entry("cuGetErrorString", g_drv.GetErrorString);
    return g_drv.ok;
}

}  // namespace

// One compiled kernel (shared by every pass of the same shape, in every plan of the process).
struct SpecKernel {
    enum State { PENDING, COMPILED, FAILED };
    uint64_t key = 0;
    std::string src;
    int threads = 0;
    size_t smem_bytes = 0, n_coefs = 0;
    std::atomic<int> state{PENDING};
    std::vector<char> cubin;
    std::string log;
    // loaded lazily on the launching thread (needs the CUDA context), per device
    std::mutex load_mu;
    int loaded_device = -1;
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    bool load_failed = false;
};

namespace {

struct Pool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<std::shared_ptr<SpecKernel>> queue;
    std::unordered_map<uint64_t, std::shared_ptr<SpecKernel>> cache;
    // passes seen before, by the hash of their launch descriptors (tile, layouts, ops with their coefficients): the same
    // circuit planned again (every simulate() of an uncompiled circuit plans) skips source generation altogether
    struct Memo { std::shared_ptr<SpecKernel> k; std::vector<uint64_t> coefs; };
    std::unordered_map<uint64_t, Memo> memo;
    std::vector<std::thread> workers;
    size_t in_flight = 0;
    bool stop = false;
    uint64_t compiled = 0, hits = 0, failed = 0;
    double seconds = 0.0;

    void worker() {
        for (;;) {
            std::shared_ptr<SpecKernel> k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !queue.empty(); });
                if (stop && queue.empty()) return;
                k = queue.front();
                queue.pop_front();
            }
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<char> cubin;
            std::string log;
            const bool ok = compile_cubin(k->src, cubin, log);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> lk(mu);
                k->cubin.swap(cubin);
                k->log.swap(log);
                if (!std::getenv("AQS_JIT_KEEP_SOURCE")) std::string().swap(k->src);
                k->state.store(ok ? SpecKernel::COMPILED : SpecKernel::FAILED, std::memory_order_release);
                if (ok) ++compiled; else ++failed;
                seconds += dt;
                --in_flight;
                if (!ok && std::getenv("AQS_JIT_VERBOSE")) std::fprintf(stderr, "[aqs jit] compilation failed: %s\n", k->log.c_str());
            }
            cv_done.notify_all();
        }
    }
    void ensure_workers() {
        if (!workers.empty()) return;
        unsigned nw = std::thread::hardware_concurrency();
        if (const char* e = std::getenv("AQS_JIT_THREADS")) nw = (unsigned)std::max(1, std::atoi(e));
        nw = std::max(1u, std::min(nw, 16u));
        for (unsigned i = 0; i < nw; ++i) workers.emplace_back([this] { worker(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            queue.clear();
        }
        cv_work.notify_all();
        for (auto& w : workers) w.join();
    }
};
Pool& pool() {
    static Pool* p = new Pool();   // intentionally leaked when workers are still compiling at exit
    return *p;
}

}  // namespace

static uint64_t pass_descriptor_hash(int n, const FusedPass& fp) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t bytes) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    int minb = 0;
    if (const char* e = std::getenv("AQS_JIT_MINB")) minb = std::atoi(e);
    mix(&n, sizeof n); mix(&fp.T, sizeof fp.T); mix(&minb, sizeof minb);
    mix(&fp.scale, sizeof fp.scale); mix(&fp.has_scale, sizeof fp.has_scale);
    mix(fp.tile.pos, (size_t)fp.tile.n);
    mix(fp.ld_toff, sizeof fp.ld_toff); mix(fp.ld_roff, sizeof fp.ld_roff); mix(fp.st_toff, sizeof fp.st_toff); mix(fp.st_roff, sizeof fp.st_roff);
    if (!fp.segs.empty()) mix(fp.segs.data(), fp.segs.size() * sizeof(TileSeg));
    if (!fp.ops.empty()) mix(fp.ops.data(), fp.ops.size() * sizeof(TileOp));
    return h;
}

int spec_attach(int n, std::vector<FusedPass>& passes, bool wait) {
    Pool& pl = pool();
    // 0. passes this process has specialised before (same descriptors, coefficients included)
    std::vector<uint64_t> dkeys(passes.size());
    {
        bool all = !passes.empty();
        std::vector<Pool::Memo> found(passes.size());
        {
            std::lock_guard<std::mutex> lk(pl.mu);
            for (size_t i = 0; i < passes.size(); ++i) {
                dkeys[i] = pass_descriptor_hash(n, passes[i]);
                auto it = pl.memo.find(dkeys[i]);
                if (it == pl.memo.end()) all = false;
                else found[i] = it->second;
            }
            if (all) pl.hits += passes.size();
        }
        if (std::getenv("AQS_JIT_VERBOSE")) std::fprintf(stderr, "[aqs jit] attach: %zu passes, memo %s\n", passes.size(), all ? "hit" : "miss");
        if (all) {
            for (size_t i = 0; i < passes.size(); ++i) { passes[i].spec = found[i].k; passes[i].spec_coefs = found[i].coefs; }
            if (wait) {
                std::unique_lock<std::mutex> lk(pl.mu);
                pl.cv_done.wait(lk, [&] {
                    for (auto& f : found)
                        if (f.k->state.load(std::memory_order_acquire) == SpecKernel::PENDING) return false;
                    return true;
                });
            }
            return AQS_OK;
        }
    }
    // 1. generate every pass's source; long circuits repeat their pass shapes (Grover-26, 64 iterations: 259 passes, 9 shapes)
    std::vector<SpecSource> srcs(passes.size());
    std::vector<char> have(passes.size(), 0);
    size_t n_new = 0;
    {
        std::vector<uint64_t> fresh_keys;
        for (size_t i = 0; i < passes.size(); ++i) {
            std::string why;
            if (!spec_generate(n, passes[i], srcs[i], why)) continue;       // this pass stays on the generic kernel
            have[i] = 1;
            std::lock_guard<std::mutex> lk(pl.mu);
            if (pl.cache.find(srcs[i].key) == pl.cache.end() && std::find(fresh_keys.begin(), fresh_keys.end(), srcs[i].key) == fresh_keys.end())
                fresh_keys.push_back(srcs[i].key);
        }
        n_new = fresh_keys.size();
    }
    size_t max_new = 128;              // a plan that needs more NEW kernels than this is not worth their compilation
    if (const char* e = std::getenv("AQS_JIT_MAX_KERNELS")) max_new = (size_t)std::max(0, std::atoi(e));
    if (n_new > max_new) return AQS_OK;
    // 2. look the shapes up / queue their compilation
    std::vector<std::shared_ptr<SpecKernel>> mine;
    for (size_t i = 0; i < passes.size(); ++i) {
        if (!have[i]) continue;
        FusedPass& fp = passes[i];
        SpecSource& s = srcs[i];
        std::shared_ptr<SpecKernel> k;
        {
            std::lock_guard<std::mutex> lk(pl.mu);
            auto it = pl.cache.find(s.key);
            if (it != pl.cache.end()) {
                k = it->second;
                ++pl.hits;
            } else {
                k = std::make_shared<SpecKernel>();
                k->key = s.key;
                k->src = std::move(s.src);
                k->threads = s.threads;
                k->smem_bytes = s.smem_bytes;
                k->n_coefs = s.coefs.size();
                pl.cache.emplace(s.key, k);
                pl.ensure_workers();
                pl.queue.push_back(k);
                ++pl.in_flight;
            }
        }
        pl.cv_work.notify_one();
        if (k->n_coefs != s.coefs.size()) continue;        // hash collision: keep the generic kernel
        fp.spec = k;
        fp.spec_coefs = std::move(s.coefs);
        mine.push_back(k);
        {
            std::lock_guard<std::mutex> lk(pl.mu);
            if (pl.memo.size() > 8192) pl.memo.clear();
            pl.memo[dkeys[i]] = Pool::Memo{k, fp.spec_coefs};
        }
    }
    if (wait) {
        std::unique_lock<std::mutex> lk(pl.mu);
        pl.cv_done.wait(lk, [&] {
            for (auto& k : mine)
                if (k->state.load(std::memory_order_acquire) == SpecKernel::PENDING) return false;
            return true;
        });
    }
    return AQS_OK;
}

int spec_wait_all() {
    Pool& pl = pool();
    std::unique_lock<std::mutex> lk(pl.mu);
    pl.cv_done.wait(lk, [&] { return pl.in_flight == 0; });
    return AQS_OK;
}

void spec_stats(uint64_t* compiled, uint64_t* cache_hits, uint64_t* failed, double* compile_seconds, uint64_t* pending) {
    Pool& pl = pool();
    std::lock_guard<std::mutex> lk(pl.mu);
    if (compiled) *compiled = pl.compiled;
    if (cache_hits) *cache_hits = pl.hits;
    if (failed) *failed = pl.failed;
    if (compile_seconds) *compile_seconds = pl.seconds;
    if (pending) *pending = pl.in_flight;
}

bool spec_ready(const FusedPass& fp) {
    SpecKernel* k = fp.spec.get();
    if (!k) return false;
    const int st = k->state.load(std::memory_order_acquire);
    if (st != SpecKernel::COMPILED) return false;
    std::lock_guard<std::mutex> lk(k->load_mu);
    if (k->load_failed) return false;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    if (k->fn && k->loaded_device == dev) return true;
    if (k->fn) return false;                     // loaded for another device of this process: that device keeps it
    if (!load_driver()) { k->load_failed = true; return false; }
    cudaFree(nullptr);                           // make sure the primary context exists and is current
    CUresult r = g_drv.ModuleLoadData(&k->mod, k->cubin.data());
    if (r == CUDA_SUCCESS) r = g_drv.ModuleGetFunction(&k->fn, k->mod, "aqs_pass");
    if (r == CUDA_SUCCESS && k->smem_bytes > 48 * 1024)
        r = g_drv.FuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k->smem_bytes);
    if (r != CUDA_SUCCESS) {
        const char* msg = nullptr;
        g_drv.GetErrorString(r, &msg);
        if (std::getenv("AQS_JIT_VERBOSE")) std::fprintf(stderr, "[aqs jit] module load failed: %s\n", msg ? msg : "?");
        k->load_failed = true;
        k->fn = nullptr;
        return false;
    }
    k->loaded_device = dev;
    std::vector<char>().swap(k->cubin);
    return true;
}

int spec_launch(const FusedPass& fp, float2* state, float2* state_out, uint64_t n_ctas, uint32_t fix_n, uint32_t fix_or, const uint8_t* fix_pos,
                cudaStream_t st) {
    const SpecKernel* k = fp.spec.get();
    // parameters: RT { u64* state; u64* state_out; u32 fix_n, fix_or; u8 fix_pos[8]; }, then one u64 per coefficient
    uint64_t rt[4];
    rt[0] = (uint64_t)(uintptr_t)state;
    rt[1] = (uint64_t)(uintptr_t)state_out;
    rt[2] = (uint64_t)fix_n | ((uint64_t)fix_or << 32);
    rt[3] = 0;
    if (fix_pos) std::memcpy(&rt[3], fix_pos, 8);
    std::vector<void*> argv(1 + fp.spec_coefs.size());
    argv[0] = rt;
    for (size_t i = 0; i < fp.spec_coefs.size(); ++i) argv[1 + i] = const_cast<uint64_t*>(&fp.spec_coefs[i]);
    void** args = argv.data();
    const CUresult r = g_drv.LaunchKernel(k->fn, (unsigned)n_ctas, 1, 1, (unsigned)k->threads, 1, 1, (unsigned)k->smem_bytes, (CUstream)st, args, nullptr);
    if (r != CUDA_SUCCESS) {
        const char* msg = nullptr;
        g_drv.GetErrorString(r, &msg);
        return fail(AQS_ERR_CUDA, std::string("specialised pass launch: ") + (msg ? msg : "?"));
    }
    return AQS_OK;
}

}  // namespace aqs

// plan.cu — compiled circuits: aqs_plan_build / aqs_plan_run and the gate-fusion planner.
//
// Replaces QCircuit::compile (reference src/quantum.cpp:199-210).  The reference
// multiplies every gate into a dense 2^n x 2^n unitary; here "compiling" turns the
// op list into a launch plan.  With AQS_PLAN_FUSE the planner
//   1. splits SWAPs into three controlled flips and merges runs of uncontrolled
//      single-qubit gates on the same qubit into one 2x2,
//   2. greedily groups ops into PASSES: a pass owns up to 12 index bits (the low 5
//      always, for coalescing) and takes, in program order, every op whose target
//      lies in those bits, skipping over ops it cannot take as long as the skipped
//      op commutes with everything taken later (two ops commute when on every
//      shared bit both act diagonally — as a control or a diagonal target),
//   3. splits each pass into SEGMENTS by the same rule with 4 register bits,
//   4. writes device descriptors for fused_kernel.cuh.
// Diagonal ops and controls never constrain the tile: QFT's CPhase ladder fuses freely.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "engine_internal.h"
#include "fused_kernel.cuh"

namespace aqs {

struct FusedPass {
    TileArgs args;
    int warps_log2 = 0;
    uint64_t n_tiles = 0;
    std::vector<TileSeg> segs;
    std::vector<TileOp> ops;
};

}  // namespace aqs

struct aqs_plan_s {
    int n = 0;
    uint32_t flags = 0;
    std::vector<aqs::CanonOp> ops;       // per-gate path
    std::vector<aqs::FusedPass> passes;  // fused path (empty => run ops one by one)
    void* arena = nullptr;               // device copy of all segment/op descriptors
    size_t arena_bytes = 0;
    cudaGraphExec_t graph = nullptr;     // AQS_PLAN_GRAPH: the launch sequence captured for `graph_state`
    const void* graph_state = nullptr;
    cudaStream_t last_stream = nullptr;  // stream of the most recent run (synchronised before the arena is recycled)
    bool ran = false;
    aqs_plan_info info{};
};

namespace aqs {

int fused_init() { return AQS_OK; }

static inline int popc(uint64_t x) { return __builtin_popcountll(x); }

// bits on which the op acts NON-diagonally / diagonally
static inline uint64_t nd_bits(const CanonOp& c) { return (c.kind == AQS_OP_U2 || c.kind == AQS_OP_X) ? (1ull << c.p) : 0ull; }
static inline uint64_t d_bits(const CanonOp& c) { return c.cmask | (c.kind == AQS_OP_DIAG ? (1ull << c.p) : 0ull); }

static inline float2 cmulh(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
static inline float2 caddh(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static bool is0(float2 z) { return z.x == 0.f && z.y == 0.f; }
static bool is1(float2 z) { return z.x == 1.f && z.y == 0.f; }

static void full_matrix(const CanonOp& c, float2 m[4]) {
    if (c.kind == AQS_OP_X) {
        m[0] = m[3] = make_float2(0.f, 0.f);
        m[1] = m[2] = make_float2(1.f, 0.f);
    } else {
        for (int i = 0; i < 4; ++i) m[i] = c.m[i];
    }
}
static void reclassify(CanonOp& c) {
    if (is0(c.m[1]) && is0(c.m[2])) {
        c.kind = AQS_OP_DIAG;
        c.d0_one = is1(c.m[0]);
        c.identity = c.d0_one && is1(c.m[3]);
    } else if (is0(c.m[0]) && is0(c.m[3]) && is1(c.m[1]) && is1(c.m[2])) {
        c.kind = AQS_OP_X; c.d0_one = false; c.identity = false;
    } else {
        c.kind = AQS_OP_U2; c.d0_one = false; c.identity = false;
    }
}

// step 1: SWAP -> 3 flips; merge uncontrolled 1-qubit gates on the same qubit; drop identities.
// A pending DIAGONAL 1-qubit gate stays mergeable across ops that use its qubit only as a
// control or diagonal target (it commutes with them): e.g. RotZ on the control qubit of a CX
// slides through the CX and fuses with the next rotation on that qubit.  The merged gate is
// emitted at the LATER position (the pending diagonal moves forward, never the later gate back).
static std::vector<CanonOp> simplify(int n, const std::vector<CanonOp>& in) {
    std::vector<CanonOp> out;
    std::vector<int> open(n, -1);   // index in `out` of a mergeable uncontrolled 1q op per bit
    auto close_bits = [&](uint64_t nd, uint64_t dg) {
        for (int b = 0; b < n; ++b) {
            if (open[b] < 0) continue;
            if (nd >> b & 1ull) open[b] = -1;
            else if ((dg >> b & 1ull) && out[open[b]].kind != AQS_OP_DIAG) open[b] = -1;
        }
    };
    auto push = [&](const CanonOp& c) {
        if (c.identity) return;
        if (c.cmask == 0 && c.kind != AQS_OP_SWAP) {
            if (open[c.p] >= 0) {
                CanonOp& prev = out[open[c.p]];
                const bool adjacent = (open[c.p] == (int)out.size() - 1);
                float2 A[4], B[4], C[4];
                full_matrix(prev, A);
                full_matrix(c, B);
                C[0] = caddh(cmulh(B[0], A[0]), cmulh(B[1], A[2]));
                C[1] = caddh(cmulh(B[0], A[1]), cmulh(B[1], A[3]));
                C[2] = caddh(cmulh(B[2], A[0]), cmulh(B[3], A[2]));
                C[3] = caddh(cmulh(B[2], A[1]), cmulh(B[3], A[3]));
                if (adjacent) {
                    for (int i = 0; i < 4; ++i) prev.m[i] = C[i];
                    reclassify(prev);
                } else {
                    CanonOp merged = prev;
                    for (int i = 0; i < 4; ++i) merged.m[i] = C[i];
                    reclassify(merged);
                    prev.identity = true;                 // the pending gate moves forward to here
                    out.push_back(merged);
                    open[c.p] = (int)out.size() - 1;
                }
                return;
            }
            out.push_back(c);
            open[c.p] = (int)out.size() - 1;
            return;
        }
        close_bits(nd_bits(c) | (c.kind == AQS_OP_SWAP ? ((1ull << c.p) | (1ull << c.p2)) : 0ull), d_bits(c));
        out.push_back(c);
    };
    for (const CanonOp& c : in) {
        if (c.kind == AQS_OP_SWAP) {
            // swap(a,b) = flip(b | a) flip(a | b) flip(b | a), each under the swap's own controls
            CanonOp f = c;
            f.kind = AQS_OP_X; f.p2 = -1;
            const int a = c.p, b = c.p2;
            f.p = b; f.cmask = c.cmask | (1ull << a); f.cval = c.cval | (1ull << a); push(f);
            f.p = a; f.cmask = c.cmask | (1ull << b); f.cval = c.cval | (1ull << b); push(f);
            f.p = b; f.cmask = c.cmask | (1ull << a); f.cval = c.cval | (1ull << a); push(f);
        } else {
            push(c);
        }
    }
    // merged ops may have become identities
    std::vector<CanonOp> kept;
    kept.reserve(out.size());
    for (const CanonOp& c : out)
        if (!c.identity) kept.push_back(c);
    return kept;
}

// Greedy "take what fits, skip what commutes": selects indices of `ops` (in order)
// whose non-diagonal target bits fit in `cap` bits on top of `base_bits`.
// Returns the chosen bit set; `taken` gets the selected indices, `rest` the others.
static uint64_t greedy_group(const std::vector<CanonOp>& ops, const std::vector<int>& cand, uint64_t base_bits,
                             uint64_t allowed_bits, int cap, std::vector<int>& taken, std::vector<int>& rest) {
    uint64_t bits = base_bits, blocked_nd = 0, blocked_d = 0;
    int room = cap;
    taken.clear();
    rest.clear();
    for (int idx : cand) {
        const CanonOp& c = ops[idx];
        const uint64_t nd = nd_bits(c), dd = d_bits(c);
        const bool conflict = (nd & (blocked_nd | blocked_d)) || (dd & blocked_nd);
        const uint64_t need = nd & ~bits;
        const bool placeable = (need & ~allowed_bits) == 0 && popc(need) <= room;
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
// the most ops from the head of `cand` (ops whose target lies on a chosen non-lane bit weigh
// more: lane-only and diagonal ops fit any group).
static uint64_t choose_bits(const std::vector<CanonOp>& ops, const std::vector<int>& cand, uint64_t base, uint64_t pool_limit,
                            int count, int n) {
    const uint64_t lane_mask = (1ull << kLaneBits) - 1ull;
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
            greedy_group(ops, head, chosen | (1ull << b), chosen | (1ull << b), 0, t2, r2);
            size_t gain = 0;
            for (int idx : t2) gain += (nd_bits(ops[idx]) & ~lane_mask) ? 4 : 1;
            if (best_bit < 0 || gain > best) { best = gain; best_bit = b; }
        }
        if (best_bit < 0) break;
        chosen |= 1ull << best_bit;
        pool &= ~(1ull << best_bit);
    }
    return chosen;
}

static int build_fused(aqs_plan_s* p) {
    const int n = p->n;
    const int T = std::min(n, kMaxTileBits);
    const uint64_t all_bits = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    const uint64_t lane_mask = (1ull << kLaneBits) - 1ull;

    std::vector<CanonOp> ops = simplify(n, p->ops);
    // Sliding window over the op stream: a pass looks at the ops deferred by earlier passes plus
    // the next kWindow ops, so planning is O(passes * window) even for million-op circuits
    // (Grover-26 with 6433 iterations lowers to ~1e6 ops).
    const size_t kWindow = 4096;
    size_t next = 0;
    std::vector<int> cand, taken, rest;
    while (true) {
        while (cand.size() < kWindow && next < ops.size()) cand.push_back((int)next++);
        if (cand.empty()) break;
        // Tile choice.  Plain first-come filling scatters the tile over whatever targets come
        // first; for nearest-neighbour circuits a better set exists.  Grow the tile one bit at a
        // time, each time adding the bit that lets the pass take the most ops from the head of
        // the window (skipped for huge op lists, where planning time matters more).
        uint64_t tile;
        if (ops.size() <= 60000) {
            const uint64_t chosen = choose_bits(ops, cand, lane_mask, all_bits, T - kLaneBits, n);
            tile = greedy_group(ops, cand, chosen, chosen, 0, taken, rest);
        } else {
            tile = greedy_group(ops, cand, lane_mask, all_bits, T - kLaneBits, taken, rest);
        }
        if (taken.empty()) return fail(AQS_ERR_STATE, "fusion planner made no progress");
        // pad the tile to exactly T bits with the lowest unused positions
        for (int b = 0; b < n && popc(tile) < T; ++b) tile |= 1ull << b;

        FusedPass fp;
        std::complex<double> pass_scale(1.0, 0.0);
        fp.warps_log2 = T - kMinTileBits;
        fp.n_tiles = 1ull << (n - T);
        std::memset(&fp.args, 0, sizeof fp.args);
        fp.args.tile_bits = (uint32_t)T;
        fp.args.tile.n = 0;
        int local_of_bit[64];
        for (int b = 0; b < 64; ++b) local_of_bit[b] = -1;
        for (int b = 0; b < n; ++b)
            if (tile >> b & 1ull) {
                local_of_bit[b] = fp.args.tile.n;
                fp.args.tile.pos[fp.args.tile.n++] = (uint8_t)b;
            }

        // segments: same greedy with 4 register bits among the non-lane tile bits
        std::vector<int> seg_cand = taken, seg_taken, seg_rest;
        const uint64_t reg_allowed = tile & ~lane_mask;
        while (!seg_cand.empty()) {
            // lane-bit targets are always placeable: treat lane bits as already present
            uint64_t chosen = greedy_group(ops, seg_cand, lane_mask, reg_allowed, kRegBits, seg_taken, seg_rest);
            if (seg_taken.empty()) return fail(AQS_ERR_STATE, "fusion planner made no progress (segment)");
            uint64_t regs = chosen & ~lane_mask;
            for (int b = kLaneBits; b < n && popc(regs) < kRegBits; ++b)
                if ((reg_allowed >> b & 1ull) && !(regs >> b & 1ull)) regs |= 1ull << b;

            TileSeg sg;
            std::memset(&sg, 0, sizeof sg);
            int reg_index_of_local[16];
            for (int j = 0; j < 16; ++j) reg_index_of_local[j] = -1;
            int ri = 0, wi = 0;
            for (int j = kLaneBits; j < T; ++j) {
                const int b = fp.args.tile.pos[j];
                if (regs >> b & 1ull) { reg_index_of_local[j] = ri; sg.R[ri++] = (uint8_t)j; }
                else sg.W[wi++] = (uint8_t)j;
            }
            sg.first_op = (uint32_t)fp.ops.size();
            // emits one tile op; (cm, cv) are the complete control mask/value in global bit positions
            auto emit = [&](const CanonOp& c, uint8_t mode, int tk, uint64_t cm, uint64_t cv, const float2 m[4]) {
                TileOp t;
                std::memset(&t, 0, sizeof t);
                t.mode = mode;
                t.tk = (uint8_t)tk;
                for (int i = 0; i < 4; ++i) t.m[i] = m[i];
                t.g_mask = cm & ~tile;
                t.g_val = cv & ~tile;
                uint32_t rk_mask = 0, rk_val = 0;
                for (int b = 0; b < n; ++b) {
                    if (!(cm & tile & (1ull << b))) continue;
                    const int j = local_of_bit[b];
                    const uint32_t v = (cv >> b) & 1ull;
                    if (reg_index_of_local[j] >= 0) {
                        rk_mask |= 1u << reg_index_of_local[j];
                        rk_val |= v << reg_index_of_local[j];
                    } else {
                        t.tl_mask |= (uint16_t)(1u << j);
                        t.tl_val |= (uint16_t)(v << j);
                    }
                }
                uint32_t act = 0;   // registers enabled by the controls that live on register bits
                for (uint32_t k = 0; k < (uint32_t)kRegs; ++k)
                    if ((k & rk_mask) == rk_val) act |= 1u << k;
                if (mode <= TM_REG_PERM) {
                    uint32_t pairs = 0;
                    for (int pr = 0; pr < kRegs / 2; ++pr) {
                        const int k0 = ((pr >> tk) << (tk + 1)) | (pr & ((1 << tk) - 1));
                        if (act >> k0 & 1u) pairs |= 1u << pr;
                    }
                    t.amp_mask = (uint16_t)pairs;
                } else {
                    t.amp_mask = (uint16_t)act;
                }
                (void)c;
                fp.ops.push_back(t);
            };
            for (int idx : seg_taken) {
                const CanonOp& c = ops[idx];
                if (c.kind == AQS_OP_DIAG) {
                    // every diagonal becomes "one factor on a selected subset"
                    const uint64_t tb = 1ull << c.p;
                    float2 one[4] = {make_float2(1.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(1.f, 0.f)};
                    if (c.d0_one) {
                        one[0] = c.m[3];
                        emit(c, TM_PHASE, 0, c.cmask | tb, c.cval | tb, one);
                    } else if (c.cmask == 0) {
                        // diag(d0, d1) = d0 * diag(1, d1/d0): d0 goes to the pass-wide scale
                        const std::complex<double> d0(c.m[0].x, c.m[0].y), d1(c.m[3].x, c.m[3].y);
                        pass_scale *= d0;
                        const std::complex<double> r = d1 / d0;
                        one[0] = make_float2((float)r.real(), (float)r.imag());
                        emit(c, TM_PHASE, 0, tb, tb, one);
                    } else {
                        one[0] = c.m[0];
                        emit(c, TM_PHASE, 0, c.cmask | tb, c.cval, one);        // target bit 0
                        one[0] = c.m[3];
                        emit(c, TM_PHASE, 0, c.cmask | tb, c.cval | tb, one);   // target bit 1
                    }
                    continue;
                }
                const int j = local_of_bit[c.p];
                const bool perm = (c.kind == AQS_OP_X);
                float2 m[4];
                full_matrix(c, m);
                if (j < kLaneBits) {
                    emit(c, perm ? TM_LANE_PERM : TM_LANE_GEN, j, c.cmask, c.cval, m);
                } else {
                    uint8_t mode = TM_REG_GEN;
                    if (perm) mode = TM_REG_PERM;
                    else if (m[0].y == 0.f && m[1].y == 0.f && m[2].y == 0.f && m[3].y == 0.f) mode = TM_REG_REAL;
                    else if (m[0].y == 0.f && m[3].y == 0.f && m[1].x == 0.f && m[2].x == 0.f) mode = TM_REG_XLIKE;
                    emit(c, mode, reg_index_of_local[j], c.cmask, c.cval, m);
                }
            }
            sg.n_ops = (uint32_t)fp.ops.size() - sg.first_op;
            fp.segs.push_back(sg);
            seg_cand.swap(seg_rest);
        }
        fp.args.scale = make_float2((float)pass_scale.real(), (float)pass_scale.imag());
        fp.args.has_scale = (pass_scale != std::complex<double>(1.0, 0.0)) ? 1u : 0u;
        p->passes.push_back(std::move(fp));
        cand.swap(rest);
    }

    return AQS_OK;
}

// Device copy of every descriptor, made on first use so that plans can be built
// (and inspected) without a GPU.
static int ensure_uploaded(aqs_plan_s* p) {
    if (p->arena || p->passes.empty()) return AQS_OK;
    auto pad = [](size_t b) { return ((b + 255) / 256) * 256; };
    size_t bytes = 0;
    for (auto& fp : p->passes) bytes += pad(fp.segs.size() * sizeof(TileSeg)) + pad(fp.ops.size() * sizeof(TileOp));
    std::vector<char> host(bytes);
    cudaError_t e = pool_alloc(&p->arena, bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(plan arena)", __LINE__);
    p->arena_bytes = bytes;
    size_t off = 0;
    for (auto& fp : p->passes) {
        fp.args.segs = reinterpret_cast<const TileSeg*>((char*)p->arena + off);
        std::memcpy(host.data() + off, fp.segs.data(), fp.segs.size() * sizeof(TileSeg));
        off += pad(fp.segs.size() * sizeof(TileSeg));
        fp.args.ops = reinterpret_cast<const TileOp*>((char*)p->arena + off);
        std::memcpy(host.data() + off, fp.ops.data(), fp.ops.size() * sizeof(TileOp));
        off += pad(fp.ops.size() * sizeof(TileOp));
        fp.args.n_segs = (uint32_t)fp.segs.size();
    }
    e = cudaMemcpy(p->arena, host.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy(plan arena)", __LINE__);
    count_h2d(bytes);
    return AQS_OK;
}

static int launch_pass(float2* state, const FusedPass& fp, cudaStream_t st) {
    TileArgs a = fp.args;
    a.state = state;
    if (fp.n_tiles > 0x7fffffffull) return fail(AQS_ERR_INVALID, "grid too large");
    const unsigned grid = (unsigned)fp.n_tiles;
    switch (fp.warps_log2) {
        case 0: k_tile<0><<<grid, 32, 0, st>>>(a); break;
        case 1: k_tile<1><<<grid, 64, 0, st>>>(a); break;
        case 2: k_tile<2><<<grid, 128, 0, st>>>(a); break;
        default: k_tile<3><<<grid, 256, 0, st>>>(a); break;
    }
    count_launch(1);
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
        uint64_t single = 0;
        p->info.n_launches = p->passes.size();
        p->info.n_fused_passes = p->passes.size();
        p->info.n_single_ops = single;
        p->info.bytes_planned = 2.0 * 8.0 * std::ldexp(1.0, n) * (double)p->passes.size();
        p->info.tile_bits = std::min(n, kMaxTileBits);
        if (std::getenv("AQS_PLAN_DUMP")) {
            size_t segs = 0, tops = 0;
            for (auto& fp : p->passes) { segs += fp.segs.size(); tops += fp.ops.size(); }
            std::fprintf(stderr, "[aqs plan] n=%d ops=%llu -> %zu tile ops, %zu passes, %zu segments\n", n,
                         (unsigned long long)n_ops, tops, p->passes.size(), segs);
            if (std::atoi(std::getenv("AQS_PLAN_DUMP")) > 1)
                for (size_t i = 0; i < p->passes.size(); ++i) {
                    auto& fp = p->passes[i];
                    std::fprintf(stderr, "  pass %zu: %zu ops in %zu segments, tile bits", i, fp.ops.size(), fp.segs.size());
                    for (int j = 0; j < fp.args.tile.n; ++j) std::fprintf(stderr, " %d", fp.args.tile.pos[j]);
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
    p->last_stream = s->stream;
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

int aqs_plan_get_info(aqs_plan_t p, aqs_plan_info* info) {
    if (!p || !info) return fail(AQS_ERR_INVALID, "null argument");
    *info = p->info;
    return AQS_OK;
}

int aqs_plan_destroy(aqs_plan_t p) {
    if (!p) return AQS_OK;
    if (p->ran && cudaStreamSynchronize(p->last_stream) != cudaSuccess) cudaGetLastError();   // kernels may still read the arena
    if (p->graph) cudaGraphExecDestroy(p->graph);
    if (p->arena) pool_free(p->arena, p->arena_bytes);
    delete p;
    return AQS_OK;
}

}  // extern "C"

// quantum.cpp — host side of the aqs core API on the B200 engine.
//
// Mirrors the behaviour of the reference's src/quantum.cpp (same checks, same
// exception types and messages, same string grammar), but every numeric step is
// one call into the C ABI of include/aqs_engine.h.  No arithmetic on the state
// vector happens on the host.
#include "quantum.h"

#include <aqs_engine.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace aqs {

// ---------------------------------------------------------------------------
// engine plumbing
// ---------------------------------------------------------------------------
namespace detail {

static bool g_engine_up = false;
static bool g_fusion    = true;
// Specialised (run-time compiled) pass kernels, include/aqs_engine.h AQS_PLAN_JIT: worth their compilation (about a
// second, once per pass shape and process) on large states only.  compile() — the reference's explicit "I will run
// this circuit" step (src/quantum.cpp:199-210) — waits for them; an uncompiled simulate() requests them in the
// background and runs on the generic kernel until they are ready.
static int g_jit_min_qubits = 26;
static uint32_t jit_flags(uint32_t qubits, bool wait) {
    if (!g_fusion || static_cast<int>(qubits) < g_jit_min_qubits) return 0u;
    return wait ? AQS_PLAN_JIT : AQS_PLAN_JIT_ASYNC;
}
static uint64_t hash_ops(const std::vector<aqs_op>& ops) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(ops.data());
    for (std::size_t i = 0, e = ops.size() * sizeof(aqs_op); i < e; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

[[noreturn]] static void engine_fail(int rc, const char* what) {
    std::string msg = std::string(what) + ": " + aqs_last_error();
    if (rc == AQS_ERR_INVALID) throw std::invalid_argument{msg};
    throw std::runtime_error{msg};
}
#define AQS_CALL(expr)                                     \
    do {                                                   \
        int rc_ = (expr);                                  \
        if (rc_ != AQS_OK) detail::engine_fail(rc_, #expr); \
    } while (0)

static void ensure_engine() {
    if (!g_engine_up) {
        int rc = aqs_engine_init(0);
        if (rc != AQS_OK) engine_fail(rc, "aqs_engine_init");
        g_engine_up = true;
    }
}

struct DeviceState {
    aqs_state_t h = nullptr;
    explicit DeviceState(uint32_t n) {
        ensure_engine();
        AQS_CALL(aqs_state_create(static_cast<int>(n), &h));
    }
    explicit DeviceState(aqs_state_t handle) : h(handle) {}
    ~DeviceState() {
        if (h) aqs_state_destroy(h);
    }
    DeviceState(const DeviceState&) = delete;
    DeviceState& operator=(const DeviceState&) = delete;
};

struct PlanCache {
    aqs_plan_t plan  = nullptr;
    std::size_t n_ops = 0;
    bool fused        = false;
    uint64_t hash     = 0;    // tail plans: FNV-1a of the lowered op records
    ~PlanCache() {
        if (plan) aqs_plan_destroy(plan);
    }
};

static std::mt19937& rng() {
    static std::random_device dv;
    static std::mt19937 gen{dv()};
    return gen;
}
// uniform float in [0, 1)  (libstdc++'s uniform_real_distribution<float> can return 1.0f)
static float draw() {
    static std::uniform_real_distribution<float> dist{0.0f, 1.0f};
    float v = dist(rng());
    return v < 1.0f ? v : std::nextafter(1.0f, 0.0f);
}

static aqs_c32 c32(const af::cfloat& z) { return aqs_c32{z.real, z.imag}; }

}  // namespace detail

static std::unordered_map<std::string, std::shared_ptr<QCircuit>> cached_circuits;

void initialize(int argc, char** argv, af::Backend) {
    int device = (argc > 1) ? std::stoi(argv[1]) : 0;
    int rc     = aqs_engine_init(device);
    if (rc != AQS_OK) detail::engine_fail(rc, "aqs_engine_init");
    detail::g_engine_up = true;
    if (const char* s = std::getenv("AQS_SEED")) set_seed(std::strtoull(s, nullptr, 10));
    if (const char* f = std::getenv("AQS_FUSION")) detail::g_fusion = std::atoi(f) != 0;
    if (const char* f = std::getenv("AQS_JIT_MIN_QUBITS")) detail::g_jit_min_qubits = std::atoi(f);
}

void clear_circuit_cache() { cached_circuits.clear(); }
void set_seed(uint64_t seed) { detail::rng().seed(static_cast<std::mt19937::result_type>(seed ^ (seed >> 32))); }
void set_fusion(bool on) { detail::g_fusion = on; }
void set_jit_min_qubits(int n) { detail::g_jit_min_qubits = n; }
void jit_wait() { aqs_jit_wait(); }
bool get_fusion() { return detail::g_fusion; }

// ---------------------------------------------------------------------------
// OpSink
// ---------------------------------------------------------------------------
static aqs_op blank_op(int kind, uint32_t target, uint64_t mask, uint64_t value) {
    aqs_op op;
    std::memset(&op, 0, sizeof op);
    op.kind       = kind;
    op.target     = static_cast<int32_t>(target);
    op.target2    = -1;
    op.ctrl_mask  = mask;
    op.ctrl_value = value;
    op.m[0] = aqs_c32{1.f, 0.f};
    op.m[3] = aqs_c32{1.f, 0.f};
    return op;
}
void OpSink::u2(uint32_t target, const af::cfloat m[4], uint64_t ctrl_mask) {
    aqs_op op = blank_op(AQS_OP_U2, target, ctrl_mask, ctrl_mask);
    for (int i = 0; i < 4; ++i) op.m[i] = detail::c32(m[i]);
    ops->push_back(op);
}
void OpSink::diag(uint32_t target, af::cfloat d0, af::cfloat d1, uint64_t ctrl_mask) {
    aqs_op op = blank_op(AQS_OP_DIAG, target, ctrl_mask, ctrl_mask);
    op.m[0] = detail::c32(d0);
    op.m[3] = detail::c32(d1);
    ops->push_back(op);
}
void OpSink::x(uint32_t target, uint64_t ctrl_mask, uint64_t ctrl_value) {
    ops->push_back(blank_op(AQS_OP_X, target, ctrl_mask, ctrl_value));
}
void OpSink::dense_gate(uint32_t begin, uint32_t k, uint64_t ctrl_mask, const af::array& mc) {
    if (!dense) throw std::logic_error{"this lowering context does not accept opaque matrices"};
    const long long D = 1ll << k;
    if (mc.dims(0) != D || mc.dims(1) != D) throw std::invalid_argument{"circuit matrix has the wrong shape"};
    DenseGateRec r;
    r.begin = begin; r.k = k; r.ctrl_mask = ctrl_mask;
    r.m.resize(static_cast<std::size_t>(D * D));
    for (long long row = 0; row < D; ++row)
        for (long long col = 0; col < D; ++col) r.m[static_cast<std::size_t>(row * D + col)] = detail::c32(mc.data()[col * D + row]);
    dense->push_back(std::move(r));
    aqs_op op = blank_op(AQS_HOST_DENSE_MARK, static_cast<uint32_t>(dense->size() - 1), 0, 0);
    ops->push_back(op);
}
void OpSink::swap(uint32_t a, uint32_t b, uint64_t ctrl_mask) {
    aqs_op op  = blank_op(AQS_OP_SWAP, a, ctrl_mask, ctrl_mask);
    op.target2 = static_cast<int32_t>(b);
    ops->push_back(op);
}

// ---------------------------------------------------------------------------
// QState  (reference src/quantum.cpp:90-157)
// ---------------------------------------------------------------------------
QState::QState(const std::complex<float>& zeroState, const std::complex<float>& oneState)
    : state_{af::cfloat{zeroState.real(), zeroState.imag()}, af::cfloat{oneState.real(), oneState.imag()}} {
    force_normalize();
}

QState::QState(const std::array<std::complex<float>, 2>& states)
    : state_{af::cfloat{states[0].real(), states[0].imag()}, af::cfloat{states[1].real(), states[1].imag()}} {
    force_normalize();
}

QState& QState::set(const std::complex<float>& zero_state, const std::complex<float>& one_state) {
    state_[0] = af::cfloat{zero_state.real(), zero_state.imag()};
    state_[1] = af::cfloat{one_state.real(), one_state.imag()};
    force_normalize();
    return *this;
}

bool QState::peek_measure() const { return detail::draw() < probability_true(); }

bool QState::measure() {
    bool m      = peek_measure();
    state_[m]   = 1.f;
    state_[!m]  = 0.f;
    return m;
}

std::array<uint32_t, 2> QState::profile_measure(uint32_t rep_count) const {
    uint32_t ones = 0;
    for (uint32_t i = 0; i < rep_count; ++i) ones += static_cast<uint32_t>(peek_measure());
    return {rep_count - ones, ones};
}

void QState::force_normalize() {
    float mag2 = state_[0].real * state_[0].real + state_[0].imag * state_[0].imag +
                 state_[1].real * state_[1].real + state_[1].imag * state_[1].imag;
    if (mag2 == 0.f) throw std::invalid_argument{"Cannot normalize a null state"};
    float mag = std::sqrt(mag2);
    state_[0] = state_[0] / mag;
    state_[1] = state_[1] / mag;
}

// ---------------------------------------------------------------------------
// QCircuit  (reference src/quantum.cpp:159-210)
// ---------------------------------------------------------------------------
QCircuit::QCircuit(uint32_t qubit_count)
    : gate_list_{}, representation_{}, qubits_{qubit_count}, compiled_ops_{std::make_shared<std::vector<aqs_op>>()},
      compiled_dense_{std::make_shared<std::vector<DenseGateRec>>()} {
    if (qubit_count < 1) throw std::invalid_argument{"Circuit must contain at least 1 qubit"};
    if (qubit_count > max_qubit_count)
        throw std::invalid_argument{"Maximum qubit count supported is " + std::to_string(max_qubit_count)};
}

void QCircuit::detach() {
    if (compiled_ops_.use_count() > 1) compiled_ops_ = std::make_shared<std::vector<aqs_op>>(*compiled_ops_);
    if (compiled_dense_.use_count() > 1) compiled_dense_ = std::make_shared<std::vector<DenseGateRec>>(*compiled_dense_);
    plan_.reset();
    tail_plan_.reset();
    matrix_.reset();
}

void QCircuit::clear() {
    clear_cache();
    gate_list_.clear();
    representation_.clear();  // deviation: the reference leaves the old string behind (src/quantum.cpp:172-178)
}

void QCircuit::clear_cache() {
    cached_index_ = 0;
    compiled_ops_ = std::make_shared<std::vector<aqs_op>>();
    compiled_dense_ = std::make_shared<std::vector<DenseGateRec>>();
    plan_.reset();
    tail_plan_.reset();
    matrix_.reset();
}

bool operator==(const QCircuit& lhs, const QCircuit& rhs) {
    if (lhs.qubit_count() != rhs.qubit_count()) return false;
    if (lhs.gate_list().size() != rhs.gate_list().size()) return false;
    if (lhs.representation() != rhs.representation()) return false;
    for (std::size_t i = 0; i < lhs.gate_list().size(); ++i)
        if (lhs.gate_list()[i] != rhs.gate_list()[i]) return false;
    return true;
}

namespace detail {
static bool has_dense(const std::vector<aqs_op>& ops) {
    for (const auto& op : ops)
        if (op.kind == AQS_HOST_DENSE_MARK) return true;
    return false;
}
// op list with opaque-matrix markers: fused plans for the runs of engine ops, aqs_apply_dense in between
static void run_segments(aqs_state_t st, uint32_t n, const std::vector<aqs_op>& ops, const std::vector<DenseGateRec>& dense, uint32_t flags,
                         uint32_t shift = 0) {
    std::size_t i = 0;
    while (i < ops.size()) {
        std::size_t j = i;
        while (j < ops.size() && ops[j].kind != AQS_HOST_DENSE_MARK) ++j;
        if (j > i) {
            PlanCache seg;
            AQS_CALL(aqs_plan_build(static_cast<int>(n), ops.data() + i, j - i, flags, &seg.plan));
            AQS_CALL(aqs_plan_run(st, seg.plan));
            AQS_CALL(aqs_sync(st));
        }
        if (j < ops.size()) {
            const DenseGateRec& r = dense.at(static_cast<std::size_t>(ops[j].target));
            std::vector<int> q(r.k);
            for (uint32_t t = 0; t < r.k; ++t) q[t] = static_cast<int>(r.begin + t + shift);
            AQS_CALL(aqs_apply_dense(st, q.data(), static_cast<int>(r.k), r.ctrl_mask << shift, r.ctrl_mask << shift, r.m.data()));
            ++j;
        }
        i = j;
    }
}
}  // namespace detail

void QCircuit::set_matrix(const af::array& m) {
    if (!gate_list_.empty()) throw std::logic_error{"set_matrix: the circuit already has gates"};
    if (qubits_ > 6) throw std::length_error{"set_matrix: opaque matrices of at most 6 qubits"};
    const long long D = 1ll << qubits_;
    if (m.dims(0) != D || m.dims(1) != D) throw std::invalid_argument{"set_matrix: the matrix must be 2^n x 2^n"};
    user_matrix_ = std::make_shared<af::array>(m);
}

void QCircuit::compile() {
    if (cached_index_ != gate_list_.size()) {
        for (std::size_t i = cached_index_; i < gate_list_.size(); ++i) (*gate_list_[i])(*this);
        cached_index_ = gate_list_.size();
    }
    // the launch plan (and, on large states, its specialised kernels) is part of compiling
    if (detail::g_engine_up && !compiled_ops_->empty() && !detail::has_dense(*compiled_ops_) && (!plan_ || plan_->n_ops != compiled_ops_->size() || plan_->fused != detail::g_fusion)) {
        auto pc = std::make_shared<detail::PlanCache>();
        const uint32_t flags = (detail::g_fusion ? AQS_PLAN_FUSE : 0u) | detail::jit_flags(qubits_, true);
        AQS_CALL(aqs_plan_build(static_cast<int>(qubits_), compiled_ops_->data(), compiled_ops_->size(), flags, &pc->plan));
        pc->n_ops = compiled_ops_->size();
        pc->fused = detail::g_fusion;
        plan_ = pc;
    }
}

std::vector<aqs_op> QCircuit::lower_all() const {
    std::vector<aqs_op> ops(*compiled_ops_);
    std::vector<DenseGateRec> dense(*compiled_dense_);
    OpSink sink{&ops, qubits_, &dense};
    for (std::size_t i = cached_index_; i < gate_list_.size(); ++i) gate_list_[i]->lower(sink, 0, 0);
    if (detail::has_dense(ops))
        throw std::logic_error{"lower_all: the circuit holds opaque matrices, which engine op records cannot carry; use QSimulator::simulate"};
    return ops;
}

static constexpr uint32_t max_dense_qubits = 13;

const af::array& QCircuit::circuit() const {
    if (opaque()) return *user_matrix_;
    if (!matrix_) {
        if (qubits_ > max_dense_qubits)
            throw std::length_error{"circuit(): the dense matrix is only materialised for up to 13 qubits"};
        // Run the compiled ops over all columns of I at once: the matrix is the
        // state of a 2n-qubit register whose low n qubits index the rows.
        const uint32_t n = qubits_;
        detail::DeviceState dev(2 * n);
        AQS_CALL(aqs_state_set_identity(dev.h));
        std::vector<aqs_op> ops(*compiled_ops_);
        for (auto& op : ops) {
            op.target += static_cast<int32_t>(n);
            if (op.target2 >= 0) op.target2 += static_cast<int32_t>(n);
            op.ctrl_mask <<= n;
            op.ctrl_value <<= n;
        }
        if (detail::has_dense(ops)) detail::run_segments(dev.h, 2 * n, ops, *compiled_dense_, 0u, n);
        else AQS_CALL(aqs_apply_ops(dev.h, ops.data(), ops.size()));   // per-gate kernels: bit-reproducible
        auto m = std::make_shared<af::array>(static_cast<long long>(1) << n, static_cast<long long>(1) << n, af::c32);
        AQS_CALL(aqs_state_download(dev.h, reinterpret_cast<aqs_c32*>(m->data()), 0, 1ull << (2 * n)));
        matrix_ = m;
    }
    return *matrix_;
}
af::array& QCircuit::circuit() {
    if (gate_list_.empty() && qubits_ <= 6) {
        // write-through (see the header): the circuit owns this matrix; it starts as the identity
        if (!user_matrix_) {
            const long long D = 1ll << qubits_;
            auto m = std::make_shared<af::array>(D, D, af::c32);
            for (long long i = 0; i < D; ++i) m->data()[i * D + i] = af::cfloat{1.f, 0.f};
            user_matrix_ = m;
        }
        return *user_matrix_;
    }
    return const_cast<af::array&>(static_cast<const QCircuit*>(this)->circuit());
}

// ---------------------------------------------------------------------------
// QSimulator  (reference src/quantum.cpp:212-531)
// ---------------------------------------------------------------------------
QSimulator::QSimulator(uint32_t qubit_count, const QState& initial_state, const QNoise& noise_generator)
    : states_(qubit_count, initial_state), noise_{noise_generator}, qubits_{qubit_count}, basis_{Basis::Z} {
    if (qubit_count < 1 || qubit_count > max_qubit_count)
        throw std::invalid_argument{"Qubit count must be in [1, " + std::to_string(max_qubit_count) + "]"};
    dev_ = std::make_shared<detail::DeviceState>(qubit_count);   // starts at |0...0>
    if (initial_state == aqs::QState::zero()) {
    } else if (initial_state == aqs::QState::one()) {
        AQS_CALL(aqs_state_set_basis(dev_->h, (1ull << qubit_count) - 1ull));
    } else {
        generate_statevector();
    }
}

QSimulator::QSimulator(uint32_t qubit_count, std::vector<QState> initial_states, const QNoise& noise_generator)
    : states_(std::move(initial_states)), noise_{noise_generator}, qubits_{qubit_count}, basis_{Basis::Z} {
    if (qubit_count != states_.size())
        throw std::invalid_argument{
            "The number of initial states must match the number of qubits in the circuit"};
    if (qubit_count < 1 || qubit_count > max_qubit_count)
        throw std::invalid_argument{"Qubit count must be in [1, " + std::to_string(max_qubit_count) + "]"};
    dev_ = std::make_shared<detail::DeviceState>(qubit_count);
    generate_statevector();
}

QSimulator::QSimulator(uint32_t qubit_count, const af::array& statevector, const QNoise& noise_generator)
    : states_(qubit_count), noise_{noise_generator}, qubits_{qubit_count}, basis_{Basis::Z} {
    if (qubit_count < 1 || qubit_count > max_qubit_count)
        throw std::invalid_argument{"Qubit count must be in [1, " + std::to_string(max_qubit_count) + "]"};
    if (statevector.dims()[0] != static_cast<long long>(fast_pow2(qubit_count)) || statevector.dims()[1] != 1 ||
        statevector.dims()[2] != 1 || statevector.dims()[3] != 1)
        throw std::invalid_argument{"Invalid initial statevector shape"};
    dev_ = std::make_shared<detail::DeviceState>(qubit_count);
    AQS_CALL(aqs_state_upload(dev_->h, reinterpret_cast<const aqs_c32*>(statevector.data()), 0, state_count()));
    double n2 = 0.0;
    AQS_CALL(aqs_norm2(dev_->h, &n2));
    if (n2 == 0.) throw std::invalid_argument{"Cannot have a null statevector"};
    AQS_CALL(aqs_scale(dev_->h, static_cast<float>(1.0 / std::sqrt(n2))));
}

QSimulator::QSimulator(const QSimulator& o)
    : states_(o.states_), noise_(o.noise_), qubits_(o.qubits_), basis_(o.basis_) {
    aqs_state_t h = nullptr;
    AQS_CALL(aqs_state_clone(o.dev_->h, &h));
    dev_ = std::make_shared<detail::DeviceState>(h);
}
QSimulator::QSimulator(QSimulator&& o) noexcept = default;
QSimulator& QSimulator::operator=(const QSimulator& o) {
    if (this != &o) {
        QSimulator tmp(o);
        *this = std::move(tmp);
    }
    return *this;
}
QSimulator& QSimulator::operator=(QSimulator&& o) noexcept = default;
QSimulator::~QSimulator() = default;

void* QSimulator::engine_handle() const noexcept { return dev_ ? dev_->h : nullptr; }
void QSimulator::sync() const { AQS_CALL(aqs_sync(dev_->h)); }

void QSimulator::generate_statevector() {
    std::vector<aqs_c32> q(2 * static_cast<std::size_t>(qubit_count()));
    for (uint32_t i = 0; i < qubit_count(); ++i) {
        q[2 * i]     = detail::c32(states_[i][0]);
        q[2 * i + 1] = detail::c32(states_[i][1]);
    }
    AQS_CALL(aqs_state_set_product(dev_->h, q.data()));
}

void QSimulator::simulate(const QCircuit& circuit) {
    if (circuit.qubit_count() != qubits_)
        throw std::invalid_argument{"Number of qubit states and circuit input qubit states do not match"};
    const uint32_t flags = detail::g_fusion ? AQS_PLAN_FUSE : 0u;

    if (circuit.opaque()) {
        // the circuit IS a user-written matrix on all its qubits
        std::vector<aqs_op> ops;
        std::vector<DenseGateRec> dense;
        OpSink sink{&ops, qubits_, &dense};
        sink.dense_gate(0, qubits_, 0, *circuit.user_matrix());
        detail::run_segments(dev_->h, qubits_, ops, dense, flags);
        return;
    }
    // compiled prefix: cached plan
    const auto& pre = circuit.compiled_ops();
    if (!pre.empty() && detail::has_dense(pre)) {
        detail::run_segments(dev_->h, qubits_, pre, circuit.compiled_dense(), flags | detail::jit_flags(qubits_, true));
    } else if (!pre.empty()) {
        auto& pc = circuit.plan_;
        if (!pc || pc->n_ops != pre.size() || pc->fused != detail::g_fusion) {
            pc = std::make_shared<detail::PlanCache>();
            AQS_CALL(aqs_plan_build(static_cast<int>(qubits_), pre.data(), pre.size(), flags | detail::jit_flags(qubits_, true), &pc->plan));
            pc->n_ops = pre.size();
            pc->fused = detail::g_fusion;
        }
        AQS_CALL(aqs_plan_run(dev_->h, pc->plan));
    }
    // uncompiled tail: lowered and planned now (gate parameters may have changed)
    if (circuit.cached_index_ < circuit.gate_list().size()) {
        std::vector<aqs_op> tail;
        std::vector<DenseGateRec> tail_dense;
        OpSink sink{&tail, qubits_, &tail_dense};
        for (std::size_t i = circuit.cached_index_; i < circuit.gate_list().size(); ++i)
            circuit.gate_list()[i]->lower(sink, 0, 0);
        if (!tail.empty() && detail::has_dense(tail)) {
            detail::run_segments(dev_->h, qubits_, tail, tail_dense, flags | detail::jit_flags(qubits_, false));
        } else if (!tail.empty()) {
            // the reference's own programs call simulate() on uncompiled circuits (benchmark/benchmark.cpp:18-28): the
            // plan of the last tail stays with the circuit and is reused while the lowered ops are the same
            const uint64_t h = detail::hash_ops(tail);
            auto& tp = circuit.tail_plan_;
            if (!tp || tp->n_ops != tail.size() || tp->hash != h || tp->fused != detail::g_fusion) {
                tp = std::make_shared<detail::PlanCache>();
                AQS_CALL(aqs_plan_build(static_cast<int>(qubits_), tail.data(), tail.size(), flags | detail::jit_flags(qubits_, false), &tp->plan));
                tp->n_ops = tail.size();
                tp->hash  = h;
                tp->fused = detail::g_fusion;
            }
            AQS_CALL(aqs_plan_run(dev_->h, tp->plan));
        }
    }
}

bool QSimulator::peek_measure(uint32_t qubit) const {
    if (qubit >= qubit_count()) throw std::out_of_range{"Cannot measure the state of the given qubit"};
    float prob = qubit_probability_true(qubit);
    return detail::draw() < prob;
}

bool QSimulator::measure(uint32_t qubit) {
    if (qubit >= qubit_count()) throw std::out_of_range{"Cannot measure the state of the given qubit"};
    float val   = detail::draw();
    float prob1 = qubit_probability_true(qubit);
    bool m      = val < prob1;
    AQS_CALL(aqs_collapse_qubit(dev_->h, static_cast<int>(qubit), m ? 1 : 0, m ? prob1 : 1.f - prob1));
    return m;
}

std::vector<uint64_t> QSimulator::sample(const std::vector<float>& draws) const {
    std::vector<uint64_t> out(draws.size());
    AQS_CALL(aqs_sample(dev_->h, draws.data(), draws.size(), out.data()));
    return out;
}

uint32_t QSimulator::peek_measure_all() const {
    float val = detail::draw();
    uint64_t k = 0;
    AQS_CALL(aqs_sample(dev_->h, &val, 1, &k));
    return static_cast<uint32_t>(k);
}

uint32_t QSimulator::measure_all() {
    uint32_t m = peek_measure_all();
    AQS_CALL(aqs_state_set_basis(dev_->h, m));
    return m;
}

float QSimulator::qubit_probability_true(uint32_t qubit) const {
    if (qubit >= qubit_count()) throw std::out_of_range{"Cannot obtain probability of the given qubit"};
    double p = 0.0;
    AQS_CALL(aqs_qubit_prob1(dev_->h, static_cast<int>(qubit), &p));
    return static_cast<float>(p);
}

float QSimulator::state_probability(uint32_t state) const {
    if (state >= state_count()) throw std::out_of_range{"Cannot obtain probability of the given state"};
    aqs_c32 v;
    AQS_CALL(aqs_state_get_amp(dev_->h, state, &v));
    return v.re * v.re + v.im * v.im;
}

af::cfloat QSimulator::state(uint32_t state) const noexcept {
    assert(state < state_count());
    aqs_c32 v{0.f, 0.f};
    if (aqs_state_get_amp(dev_->h, state, &v) != AQS_OK) return af::cfloat{std::numeric_limits<float>::quiet_NaN(), 0.f};
    return af::cfloat{v.re, v.im};
}

std::vector<float> QSimulator::probabilities() const {
    std::vector<float> out(state_count());
    AQS_CALL(aqs_probabilities(dev_->h, out.data(), 0, out.size()));
    return out;
}

std::vector<uint32_t> QSimulator::profile_measure_all(uint32_t rep_count) const {
    std::vector<uint32_t> count(state_count());
    std::vector<float> u(rep_count);
    for (auto& v : u) v = detail::draw();
    AQS_CALL(aqs_sample_hist(dev_->h, u.data(), u.size(), count.data()));
    return count;
}

std::array<uint32_t, 2> QSimulator::profile_measure(uint32_t qubit, uint32_t rep_count) const {
    if (qubit >= qubit_count()) throw std::out_of_range{"Cannot profile measurement of the given qubit"};
    float prob1   = qubit_probability_true(qubit);
    uint32_t ones = 0;
    for (uint32_t i = 0; i < rep_count; ++i) ones += static_cast<uint32_t>(detail::draw() < prob1);
    return {rep_count - ones, ones};
}

const af::array& QSimulator::statevector() const {
    // One host buffer per simulator, refreshed in place: a reference obtained earlier stays valid (and shows the current
    // state after the next call) instead of dangling, like the reference's member array (include/quantum.h:726-739).
    if (!snapshot_ || snapshot_->elements() != static_cast<long long>(state_count()))
        snapshot_ = std::make_shared<af::array>(static_cast<long long>(state_count()), af::c32);
    AQS_CALL(aqs_state_download(dev_->h, reinterpret_cast<aqs_c32*>(snapshot_->data()), 0, state_count()));
    return *snapshot_;
}
af::array& QSimulator::statevector() { return const_cast<af::array&>(static_cast<const QSimulator*>(this)->statevector()); }

double QSimulator::norm2() const {
    double v = 0.0;
    AQS_CALL(aqs_norm2(dev_->h, &v));
    return v;
}

// set_basis (reference src/quantum.cpp:416-465): the same 2x2 change of basis on
// every qubit.  The reference builds the dense n-fold Kronecker power; here it is
// n single-qubit ops.  Deviation: the reference's `case Basis::Y` falls through
// into `case Basis::X` (:443-444), so leaving the Y basis used the X matrix; this
// implementation uses the Y matrix.
void QSimulator::set_basis(Basis basis) {
    if (basis == basis_) return;
    const float h = 0.70710678118f;
    using C = std::complex<float>;
    // column-major host arrays transposed in the reference => these row-major forms
    const C z_to_x[4] = {{h, 0}, {h, 0}, {h, 0}, {-h, 0}};
    const C z_to_y[4] = {{h, 0}, {h, 0}, {0, -h}, {0, h}};
    auto inverse = [](const C m[4], C out[4]) {
        C det  = m[0] * m[3] - m[1] * m[2];
        out[0] = m[3] / det;
        out[1] = -m[1] / det;
        out[2] = -m[2] / det;
        out[3] = m[0] / det;
    };
    C to_z[4] = {{1, 0}, {0, 0}, {0, 0}, {1, 0}};
    if (basis_ == Basis::Y) inverse(z_to_y, to_z);
    if (basis_ == Basis::X) inverse(z_to_x, to_z);
    C m[4] = {to_z[0], to_z[1], to_z[2], to_z[3]};
    const C* from_z = basis == Basis::Y ? z_to_y : (basis == Basis::X ? z_to_x : nullptr);
    if (from_z) {
        m[0] = from_z[0] * to_z[0] + from_z[1] * to_z[2];
        m[1] = from_z[0] * to_z[1] + from_z[1] * to_z[3];
        m[2] = from_z[2] * to_z[0] + from_z[3] * to_z[2];
        m[3] = from_z[2] * to_z[1] + from_z[3] * to_z[3];
    }
    std::vector<aqs_op> ops;
    OpSink sink{&ops, qubits_};
    af::cfloat mm[4] = {m[0], m[1], m[2], m[3]};
    for (uint32_t q = 0; q < qubits_; ++q) sink.u2(q, mm, 0);
    detail::PlanCache tmp;
    AQS_CALL(aqs_plan_build(static_cast<int>(qubits_), ops.data(), ops.size(), detail::g_fusion ? AQS_PLAN_FUSE : 0u, &tmp.plan));
    AQS_CALL(aqs_plan_run(dev_->h, tmp.plan));
    AQS_CALL(aqs_sync(dev_->h));
    basis_ = basis;
}

// ---------------------------------------------------------------------------
// gates
// ---------------------------------------------------------------------------
void QGate::lower(OpSink&, uint32_t, uint64_t) const {
    throw std::logic_error{
        "This QGate subclass does not override lower(): user-defined gates must emit engine ops "
        "(the reference's hand-edited qc.circuit() matrices are not supported)"};
}
std::shared_ptr<QGate> QGate::clone() const { throw std::logic_error{"This QGate subclass does not override clone()"}; }

static const char* kPosErr = "Cannot add gate at the given qubit position";
static const char* kCtlErr = "Control qubit cannot be the same as the target qubit";
static const char* kSimErr = "Gate not supported for given simulation";

static inline uint64_t bit(uint32_t q) { return 1ull << q; }

// appends the lowered gate to the circuit's compiled ops
template<typename G>
static QCircuit& compile_into(const G& g, QCircuit& qc) {
    g.check(qc);
    OpSink sink{&qc.compiled_ops(), qc.qubit_count(), &qc.compiled_dense()};
    g.lower(sink, 0, 0);
    return qc;
}

static void check1(uint32_t t, const QCircuit& qc) {
    if (t >= qc.qubit_count()) throw std::out_of_range{kPosErr};
}
static void check2(uint32_t c, uint32_t t, const QCircuit& qc) {
    const auto qubits = qc.qubit_count();
    if (qubits < 2) throw std::domain_error{kSimErr};
    if (c >= qubits) throw std::out_of_range{kPosErr};
    if (t >= qubits) throw std::out_of_range{kPosErr};
    if (c == t) throw std::invalid_argument{kCtlErr};
}
// CRot* have a shorter check list in the reference (src/quantum.cpp:1413-1422)
static void check2_rot(uint32_t c, uint32_t t, const QCircuit& qc) {
    const auto qubits = qc.qubit_count();
    if (t >= qubits || c >= qubits) throw std::out_of_range{kPosErr};
    if (t == c) throw std::invalid_argument{kCtlErr};
}
static void check3(uint32_t a, uint32_t b, uint32_t t, const QCircuit& qc) {
    const auto qubits = qc.qubit_count();
    if (qubits < 3) throw std::domain_error{kSimErr};
    if (a >= qubits) throw std::out_of_range{kPosErr};
    if (b >= qubits) throw std::out_of_range{kPosErr};
    if (t >= qubits) throw std::out_of_range{kPosErr};
    if (a == t || b == t) throw std::invalid_argument{kCtlErr};
}

static void mat_h(af::cfloat m[4]) {
    const float h = 0.70710678118f;
    m[0] = {h, 0.f}; m[1] = {h, 0.f}; m[2] = {h, 0.f}; m[3] = {-h, 0.f};
}
static void mat_y(af::cfloat m[4]) {
    m[0] = {0.f, 0.f}; m[1] = {0.f, -1.f}; m[2] = {0.f, 1.f}; m[3] = {0.f, 0.f};
}
static void mat_rotx(float angle, af::cfloat m[4]) {
    float c = std::cos(angle / 2.0f), s = std::sin(angle / 2.0f);
    m[0] = {c, 0.f}; m[1] = {0.f, -s}; m[2] = {0.f, -s}; m[3] = {c, 0.f};
}
static void mat_roty(float angle, af::cfloat m[4]) {
    float c = std::cos(angle / 2.0f), s = std::sin(angle / 2.0f);
    m[0] = {c, 0.f}; m[1] = {-s, 0.f}; m[2] = {s, 0.f}; m[3] = {c, 0.f};
}
static std::string phase_name(float angle) {
    if (angle == pi / 2) return "S";
    if (angle == -pi / 2) return "S†";
    if (angle == pi / 4) return "T";
    if (angle == -pi / 4) return "T†";
    return "Phase";
}
static std::string s(uint32_t v) { return std::to_string(v); }

#define AQS_EQ_BEGIN(Class) \
    bool Class::operator==(const QGate& rhs) const noexcept { \
        if (type() != rhs.type()) return false;               \
        const auto& o = *static_cast<const Class*>(&rhs);
#define AQS_EQ_END }

// X ------------------------------------------------------------------------
bool X::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& X::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void X::lower(OpSink& k, uint32_t off, uint64_t cm) const { k.x(target_qubit + off, cm, cm); }
std::string X::to_string() const { return "X,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(X) return target_qubit == o.target_qubit; AQS_EQ_END

// Y ------------------------------------------------------------------------
bool Y::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& Y::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void Y::lower(OpSink& k, uint32_t off, uint64_t cm) const { af::cfloat m[4]; mat_y(m); k.u2(target_qubit + off, m, cm); }
std::string Y::to_string() const { return "Y,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(Y) return target_qubit == o.target_qubit; AQS_EQ_END

// Z ------------------------------------------------------------------------
bool Z::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& Z::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void Z::lower(OpSink& k, uint32_t off, uint64_t cm) const { k.diag(target_qubit + off, {1.f, 0.f}, {-1.f, 0.f}, cm); }
std::string Z::to_string() const { return "Z,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(Z) return target_qubit == o.target_qubit; AQS_EQ_END

// RotX / RotY / RotZ ---------------------------------------------------------
bool RotX::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& RotX::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void RotX::lower(OpSink& k, uint32_t off, uint64_t cm) const { af::cfloat m[4]; mat_rotx(angle, m); k.u2(target_qubit + off, m, cm); }
std::string RotX::to_string() const { return "RotX,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(RotX) return target_qubit == o.target_qubit && angle == o.angle; AQS_EQ_END

bool RotY::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& RotY::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void RotY::lower(OpSink& k, uint32_t off, uint64_t cm) const { af::cfloat m[4]; mat_roty(angle, m); k.u2(target_qubit + off, m, cm); }
std::string RotY::to_string() const { return "RotY,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(RotY) return target_qubit == o.target_qubit && angle == o.angle; AQS_EQ_END

bool RotZ::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& RotZ::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void RotZ::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    float c = std::cos(angle / 2.0f), sn = std::sin(angle / 2.0f);
    k.diag(target_qubit + off, {c, -sn}, {c, sn}, cm);
}
std::string RotZ::to_string() const { return "RotZ,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(RotZ) return target_qubit == o.target_qubit && angle == o.angle; AQS_EQ_END

// H --------------------------------------------------------------------------
bool H::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& H::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void H::lower(OpSink& k, uint32_t off, uint64_t cm) const { af::cfloat m[4]; mat_h(m); k.u2(target_qubit + off, m, cm); }
std::string H::to_string() const { return "H,0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(H) return target_qubit == o.target_qubit; AQS_EQ_END

// Phase ------------------------------------------------------------------------
bool Phase::check(const QCircuit& qc) const { check1(target_qubit, qc); return true; }
QCircuit& Phase::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void Phase::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    k.diag(target_qubit + off, {1.f, 0.f}, {std::cos(angle), std::sin(angle)}, cm);
}
std::string Phase::to_string() const { return phase_name(angle) + ",0,1:" + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(Phase) return target_qubit == o.target_qubit && angle == o.angle; AQS_EQ_END

// Swap -------------------------------------------------------------------------
bool Swap::check(const QCircuit& qc) const {
    const auto qubits = qc.qubit_count();
    if (qubits < 2) throw std::domain_error{kSimErr};
    if (target_qubit_A >= qubits) throw std::out_of_range{kPosErr};
    if (target_qubit_B >= qubits) throw std::out_of_range{kPosErr};
    if (target_qubit_A == target_qubit_B) throw std::invalid_argument{"Cannot use the swap gate on the same target qubits"};
    return true;
}
QCircuit& Swap::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void Swap::lower(OpSink& k, uint32_t off, uint64_t cm) const { k.swap(target_qubit_A + off, target_qubit_B + off, cm); }
std::string Swap::to_string() const { return "Swap,0,2:" + s(target_qubit_A) + "," + s(target_qubit_B) + ";"; }
AQS_EQ_BEGIN(Swap) return target_qubit_A == o.target_qubit_A && target_qubit_B == o.target_qubit_B; AQS_EQ_END

// CX / CY / CZ / CH ----------------------------------------------------------------
bool CX::check(const QCircuit& qc) const { check2(control_qubit, target_qubit, qc); return true; }
QCircuit& CX::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CX::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    const uint64_t m = cm | bit(control_qubit + off);
    k.x(target_qubit + off, m, m);
}
std::string CX::to_string() const { return "X,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CX) return target_qubit == o.target_qubit && control_qubit == o.control_qubit; AQS_EQ_END

bool CY::check(const QCircuit& qc) const { check2(control_qubit, target_qubit, qc); return true; }
QCircuit& CY::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CY::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    af::cfloat m[4]; mat_y(m);
    k.u2(target_qubit + off, m, cm | bit(control_qubit + off));
}
std::string CY::to_string() const { return "Y,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CY) return target_qubit == o.target_qubit && control_qubit == o.control_qubit; AQS_EQ_END

bool CZ::check(const QCircuit& qc) const { check2(control_qubit, target_qubit, qc); return true; }
QCircuit& CZ::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CZ::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    k.diag(target_qubit + off, {1.f, 0.f}, {-1.f, 0.f}, cm | bit(control_qubit + off));
}
std::string CZ::to_string() const { return "Z,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CZ) return target_qubit == o.target_qubit && control_qubit == o.control_qubit; AQS_EQ_END

bool CH::check(const QCircuit& qc) const { check2(control_qubit, target_qubit, qc); return true; }
QCircuit& CH::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CH::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    af::cfloat m[4]; mat_h(m);
    k.u2(target_qubit + off, m, cm | bit(control_qubit + off));
}
std::string CH::to_string() const { return "H,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CH) return target_qubit == o.target_qubit && control_qubit == o.control_qubit; AQS_EQ_END

// CPhase ---------------------------------------------------------------------------
bool CPhase::check(const QCircuit& qc) const { check2(control_qubit, target_qubit, qc); return true; }
QCircuit& CPhase::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CPhase::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    k.diag(target_qubit + off, {1.f, 0.f}, {std::cos(angle), std::sin(angle)}, cm | bit(control_qubit + off));
}
std::string CPhase::to_string() const {
    return phase_name(angle) + ",1,1:" + s(control_qubit) + "," + s(target_qubit) + ";";
}
AQS_EQ_BEGIN(CPhase) return target_qubit == o.target_qubit && control_qubit == o.control_qubit && angle == o.angle; AQS_EQ_END

// CSwap ----------------------------------------------------------------------------
bool CSwap::check(const QCircuit& qc) const {
    const auto qubits = qc.qubit_count();
    if (qubits < 3) throw std::domain_error{"Gate not supported for given circuit"};
    if (target_qubit_A >= qubits) throw std::out_of_range{kPosErr};
    if (target_qubit_B >= qubits) throw std::out_of_range{kPosErr};
    if (control_qubit >= qubits) throw std::out_of_range{kPosErr};
    if (control_qubit == target_qubit_A || control_qubit == target_qubit_B) throw std::invalid_argument{kCtlErr};
    if (target_qubit_A == target_qubit_B) throw std::invalid_argument{"Cannot use the swap gate on the same target qubits"};
    return true;
}
QCircuit& CSwap::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CSwap::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    k.swap(target_qubit_A + off, target_qubit_B + off, cm | bit(control_qubit + off));
}
// NB reproduces the reference's missing comma between A and B (src/quantum.cpp:1322-1326)
std::string CSwap::to_string() const {
    return "Swap,1,2:" + s(control_qubit) + "," + s(target_qubit_A) + s(target_qubit_B) + ";";
}
AQS_EQ_BEGIN(CSwap)
    return target_qubit_A == o.target_qubit_A && target_qubit_B == o.target_qubit_B && control_qubit == o.control_qubit;
AQS_EQ_END

// CRotX / CRotY / CRotZ -------------------------------------------------------------
bool CRotX::check(const QCircuit& qc) const { check2_rot(control_qubit, target_qubit, qc); return true; }
QCircuit& CRotX::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CRotX::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    af::cfloat m[4]; mat_rotx(angle, m);
    k.u2(target_qubit + off, m, cm | bit(control_qubit + off));
}
std::string CRotX::to_string() const { return "RotX,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CRotX) return target_qubit == o.target_qubit && control_qubit == o.control_qubit && angle == o.angle; AQS_EQ_END

bool CRotY::check(const QCircuit& qc) const { check2_rot(control_qubit, target_qubit, qc); return true; }
QCircuit& CRotY::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CRotY::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    af::cfloat m[4]; mat_roty(angle, m);
    k.u2(target_qubit + off, m, cm | bit(control_qubit + off));
}
std::string CRotY::to_string() const { return "RotY,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CRotY) return target_qubit == o.target_qubit && control_qubit == o.control_qubit && angle == o.angle; AQS_EQ_END

bool CRotZ::check(const QCircuit& qc) const { check2_rot(control_qubit, target_qubit, qc); return true; }
QCircuit& CRotZ::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CRotZ::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    float c = std::cos(angle / 2.0f), sn = std::sin(angle / 2.0f);
    k.diag(target_qubit + off, {c, -sn}, {c, sn}, cm | bit(control_qubit + off));
}
std::string CRotZ::to_string() const { return "RotZ,1,1:" + s(control_qubit) + "," + s(target_qubit) + ";"; }
AQS_EQ_BEGIN(CRotZ) return target_qubit == o.target_qubit && control_qubit == o.control_qubit && angle == o.angle; AQS_EQ_END

// CCNot / Or -------------------------------------------------------------------------
bool CCNot::check(const QCircuit& qc) const { check3(control_qubit_A, control_qubit_B, target_qubit, qc); return true; }
QCircuit& CCNot::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void CCNot::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    const uint64_t m = cm | bit(control_qubit_A + off) | bit(control_qubit_B + off);
    k.x(target_qubit + off, m, m);
}
std::string CCNot::to_string() const {
    return "X,2,1:" + s(control_qubit_A) + "," + s(control_qubit_B) + "," + s(target_qubit) + ";";
}
AQS_EQ_BEGIN(CCNot)
    return target_qubit == o.target_qubit && control_qubit_A == o.control_qubit_A && control_qubit_B == o.control_qubit_B;
AQS_EQ_END

bool Or::check(const QCircuit& qc) const { check3(control_qubit_A, control_qubit_B, target_qubit, qc); return true; }
QCircuit& Or::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
// t ^= (a OR b)  ==  flip t, then flip t again where a == 0 and b == 0
void Or::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    k.x(target_qubit + off, cm, cm);
    k.x(target_qubit + off, cm | bit(control_qubit_A + off) | bit(control_qubit_B + off), cm);
}
std::string Or::to_string() const {
    std::stringstream b;
    b << "P;"
      << "X,0,1:" << control_qubit_A << ";"
      << "X,0,1:" << control_qubit_B << ";"
      << "X,0,1:" << target_qubit << ";"
      << "X,2,1:" << control_qubit_A << "," << control_qubit_B << "," << target_qubit << ";"
      << "X,0,1:" << control_qubit_A << ";"
      << "X,0,1:" << control_qubit_B << ";"
      << "P;";
    return b.str();
}
AQS_EQ_BEGIN(Or)
    return target_qubit == o.target_qubit && control_qubit_A == o.control_qubit_A && control_qubit_B == o.control_qubit_B;
AQS_EQ_END

// ---------------------------------------------------------------------------
// Gate / ControlGate (reference src/quantum.cpp:1679-1960)
// ---------------------------------------------------------------------------
// One statement of the circuit string grammar is `Name,<#ctrl>,<#tgt>:q,q,...;`
// (or a bare `B;` / `P;`).  shift_statements rewrites every statement with all
// qubit indices + offset and, if add_control, one more control in front.
static std::string shift_statements(const std::string& text, uint32_t offset, bool add_control, uint32_t control) {
    std::string out;
    std::size_t begin = 0;
    for (std::size_t end = text.find(';', begin); end != std::string::npos; begin = end + 1, end = text.find(';', begin)) {
        const std::string stmt = text.substr(begin, end - begin);
        const std::size_t colon = stmt.find(':');
        if (stmt.find(',') == std::string::npos || colon == std::string::npos) {   // barrier or unknown: verbatim
            out += stmt + ";";
            continue;
        }
        const std::size_t c1 = stmt.find(',');
        const std::size_t c2 = stmt.find(',', c1 + 1);
        const std::string name = stmt.substr(0, c1);
        int n_ctrl = std::stoi(stmt.substr(c1 + 1, c2 - c1 - 1));
        const std::string n_tgt = stmt.substr(c2 + 1, colon - c2 - 1);
        out += name + "," + std::to_string(n_ctrl + (add_control ? 1 : 0)) + "," + n_tgt + ":";
        if (add_control) out += std::to_string(control) + ",";
        std::size_t p = colon + 1;
        bool first    = true;
        while (p <= stmt.size()) {
            std::size_t q = stmt.find(',', p);
            if (q == std::string::npos) q = stmt.size();
            if (!first) out += ",";
            out += std::to_string(static_cast<uint32_t>(std::stoi(stmt.substr(p, q - p))) + offset);
            first = false;
            p     = q + 1;
        }
        out += ";";
    }
    return out;
}

static std::shared_ptr<QCircuit> snapshot_circuit(const QCircuit& c) {
    if (c.opaque()) return std::make_shared<QCircuit>(c);      // (no string representation to key the cache on)
    auto it = cached_circuits.find(c.representation());
    if (it == cached_circuits.end()) {
        it = cached_circuits.insert({c.representation(), std::make_shared<QCircuit>(c)}).first;
    } else if (!(c == *(it->second))) {
        it->second = std::make_shared<QCircuit>(c);
    }
    return it->second;
}

static std::string named_statement(const std::string& name, bool ctrl, uint32_t control, uint32_t begin, uint32_t count) {
    if (name.find_first_of(",;:") != std::string::npos)
        throw std::invalid_argument{"Name cannot contain commas, colons, nor semicolons"};
    std::stringstream b;
    b << name << (ctrl ? ",1," : ",0,") << count << ":";
    if (ctrl) b << control << ",";
    b << begin;
    for (uint32_t i = 1; i < count; ++i) b << "," << i + begin;
    b << ";";
    return b.str();
}

Gate::Gate(const QCircuit& circuit_, uint32_t target_qubit_begin_, std::string name)
    : representation{}, qubit_count{circuit_.qubit_count()}, target_qubit_begin{target_qubit_begin_} {
    internal_circuit = snapshot_circuit(circuit_);
    representation   = name.empty() ? shift_statements(circuit_.representation(), target_qubit_begin_, false, 0)
                                    : named_statement(name, false, 0, target_qubit_begin_, circuit_.qubit_count());
}

bool Gate::check(const QCircuit& qc) const {
    const auto qubits = qc.qubit_count();
    if (target_qubit_begin >= qubits) throw std::out_of_range{kPosErr};
    if (target_qubit_begin + qubit_count > qubits) throw std::out_of_range{kPosErr};
    return true;
}
QCircuit& Gate::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void Gate::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    const QCircuit& in = *internal_circuit;
    if (in.opaque()) { k.dense_gate(off + target_qubit_begin, in.qubit_count(), cm, *in.user_matrix()); return; }
    for (const auto& g : in.gate_list()) g->lower(k, off + target_qubit_begin, cm);
}
bool Gate::operator==(const QGate& rhs) const noexcept {
    if (type() != rhs.type()) return false;
    const auto& o = *static_cast<const Gate*>(&rhs);
    return target_qubit_begin == o.target_qubit_begin && internal_circuit == o.internal_circuit;
}

ControlGate::ControlGate(const QCircuit& circuit_, uint32_t control_qubit_, uint32_t target_qubit_begin_, std::string name)
    : representation{}
    , qubit_count{circuit_.qubit_count()}
    , control_qubit{control_qubit_}
    , target_qubit_begin{target_qubit_begin_} {
    internal_circuit = snapshot_circuit(circuit_);
    representation   = name.empty()
                           ? shift_statements(circuit_.representation(), target_qubit_begin_, true, control_qubit_)
                           : named_statement(name, true, control_qubit_, target_qubit_begin_, circuit_.qubit_count());
}

bool ControlGate::check(const QCircuit& qc) const {
    const auto qubits = qc.qubit_count();
    if (target_qubit_begin + qubit_count > qubits) throw std::out_of_range{"Gate must fit inside the circuit qubit count"};
    if (target_qubit_begin <= control_qubit && control_qubit < target_qubit_begin + qubit_count)
        throw std::out_of_range{"Control qubit cannot be one of the target qubits of the gate"};
    if (qubit_count >= qubits) throw std::invalid_argument{"Cannot add a bigger gate to the circuit"};
    if (control_qubit >= qubits) throw std::out_of_range{"Control qubit must be inside the circuit qubit range"};
    return true;
}
QCircuit& ControlGate::operator()(QCircuit& qc) const { return compile_into(*this, qc); }
void ControlGate::lower(OpSink& k, uint32_t off, uint64_t cm) const {
    const QCircuit& in = *internal_circuit;
    const uint64_t m   = cm | bit(control_qubit + off);
    if (in.opaque()) { k.dense_gate(off + target_qubit_begin, in.qubit_count(), m, *in.user_matrix()); return; }
    for (const auto& g : in.gate_list()) g->lower(k, off + target_qubit_begin, m);
}
bool ControlGate::operator==(const QGate& rhs) const noexcept {
    if (type() != rhs.type()) return false;
    const auto& o = *static_cast<const ControlGate*>(&rhs);
    return target_qubit_begin == o.target_qubit_begin && control_qubit == o.control_qubit &&
           internal_circuit == o.internal_circuit;
}

// ---------------------------------------------------------------------------
// single-qubit host operations (reference src/quantum.cpp:1962-2032)
// ---------------------------------------------------------------------------
static QState apply_2x2(const af::cfloat m[4], const QState& st) {
    af::cfloat a = m[0] * st[0] + m[1] * st[1];
    af::cfloat b = m[2] * st[0] + m[3] * st[1];
    return {{a.real, a.imag}, {b.real, b.imag}};
}
QState X_op(const QState& st) {
    const af::cfloat m[4] = {{0.f, 0.f}, {1.f, 0.f}, {1.f, 0.f}, {0.f, 0.f}};
    return apply_2x2(m, st);
}
QState Y_op(const QState& st) { af::cfloat m[4]; mat_y(m); return apply_2x2(m, st); }
QState Z_op(const QState& st) {
    const af::cfloat m[4] = {{1.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {-1.f, 0.f}};
    return apply_2x2(m, st);
}
QState RotateX_op(const QState& st, float angle) { af::cfloat m[4]; mat_rotx(angle, m); return apply_2x2(m, st); }
QState RotateY_op(const QState& st, float angle) { af::cfloat m[4]; mat_roty(angle, m); return apply_2x2(m, st); }
QState RotateZ_op(const QState& st, float angle) {
    const af::cfloat m[4] = {{std::cos(angle / 2.f), -std::sin(angle / 2.f)}, {0.f, 0.f}, {0.f, 0.f},
                             {std::cos(angle / 2.f), std::sin(angle / 2.f)}};
    return apply_2x2(m, st);
}
QState Hadamard_op(const QState& st) { af::cfloat m[4]; mat_h(m); return apply_2x2(m, st); }
QState Phase_op(const QState& st, float angle) {
    const af::cfloat m[4] = {{1.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {std::cos(angle), std::sin(angle)}};
    return apply_2x2(m, st);
}

}  // namespace aqs

// ---------------------------------------------------------------------------
// af:: shim functions that need the engine
// ---------------------------------------------------------------------------
namespace af {
std::string infoString() {
    int dev = 0, sms = 0;
    size_t mem = 0;
    if (aqs_engine_device(&dev, &sms, &mem) != AQS_OK) return "aqs B200 engine (not initialised)";
    std::stringstream b;
    b << "aqs B200 engine " << AQS_B200_ENGINE_VERSION << " (sm_100a kernels, no ArrayFire): CUDA device " << dev << ", "
      << sms << " SMs, " << (mem >> 20) << " MiB";
    return b.str();
}
void info() { std::printf("%s\n", infoString().c_str()); }
void sync() {}
void print(const char* name, const array& a) {
    std::printf("%s [%lld %lld]\n", name, a.dims(0), a.dims(1));
    for (long long r = 0; r < a.dims(0); ++r) {
        for (long long c = 0; c < a.dims(1); ++c) {
            const cfloat v = a.data()[c * a.dims(0) + r];
            std::printf(" (%8.4f,%8.4f)", v.real, v.imag);
        }
        std::printf("\n");
    }
}
}  // namespace af

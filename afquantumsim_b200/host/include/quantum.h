/*
 * quantum.h — the afQuantumSim core API (namespace aqs) on the B200 engine.
 *
 * Source-compatible with the reference's include/quantum.h: same classes, same
 * constructors, same public data members, same exceptions.  What changed is
 * underneath: a gate no longer builds a 2^n x 2^n matrix and multiplies it
 * (reference src/quantum.cpp, every QGate::operator()); it *lowers* itself to
 * primitive ops (include/aqs_engine.h) that hand-written sm_100a kernels apply
 * in place to the state vector in HBM.
 *
 *   QCircuit::compile()     -> lowers the new gates and caches a fused launch plan
 *   QSimulator::simulate()  -> runs the plan (compiled prefix) + the lowered tail
 *   QCircuit::circuit()     -> dense matrix, materialised on demand (n <= 13)
 *
 * Deviations from the reference are listed in DESIGN.md ("reference quirks").
 */
#pragma once

#include <arrayfire.h>

#include "utils.h"
#include "version.h"

#include <array>
#include <cassert>
#include <complex>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <aqs_engine.h> /* the engine C ABI: struct aqs_op */

namespace aqs {

static constexpr float pi                 = 3.14159265358979323846f;
static constexpr uint32_t max_qubit_count = 30;

class QGate;
class QState;
class QCircuit;
class QSimulator;

/* Select the CUDA device (argv[1], default 0) and start the engine.
 * `backend` is accepted for source compatibility and ignored: there is one
 * backend, sm_100a.  (reference src/quantum.cpp:69-86) */
void initialize(int argc, char** argv, af::Backend backend = af::Backend::AF_BACKEND_DEFAULT);
void clear_circuit_cache();

/* ---- extensions (not in the reference) ---- */
/* Seed the host RNG behind the measure and profile calls (the reference seeds
 * from std::random_device and offers no hook).  Also read from $AQS_SEED at initialize. */
void set_seed(uint64_t seed);
/* Gate fusion for simulate()/compile() plans (default on; $AQS_FUSION=0 turns it off). */
void set_fusion(bool on);
/* specialised (run-time compiled) pass kernels are requested for circuits on at least this many qubits (default 26;
 * AQS_JIT_MIN_QUBITS): compile() waits for them, an uncompiled simulate() lets them compile in the background */
void set_jit_min_qubits(int n);
void jit_wait();   /* block until no kernel compilation is pending */
bool get_fusion();

/* Collects the primitive ops a gate lowers to. */
/* An opaque 2^k x 2^k matrix on contiguous qubits [begin, begin + k): what Gate / ControlGate apply when the inner
 * circuit's matrix was written by the user (reference src/quantum.cpp:1760-1814, 1888-1950; docs/USAGE.md:121-125).
 * It cannot travel in a 64-byte aqs_op record, so the op list carries a marker (kind AQS_HOST_DENSE_MARK, target =
 * index into the list of these records) and the host layer calls aqs_apply_dense between the fused segments. */
struct DenseGateRec {
    uint32_t begin = 0, k = 0;
    uint64_t ctrl_mask = 0;
    std::vector<aqs_c32> m;   /* row-major, qubit `begin` = most significant matrix-index bit */
};
constexpr int AQS_HOST_DENSE_MARK = 99;

struct OpSink {
    std::vector<aqs_op>* ops;
    uint32_t qubits;
    std::vector<DenseGateRec>* dense = nullptr;
    void dense_gate(uint32_t begin, uint32_t k, uint64_t ctrl_mask, const af::array& matrix_colmajor);
    void u2(uint32_t target, const af::cfloat m[4], uint64_t ctrl_mask);
    void diag(uint32_t target, af::cfloat d0, af::cfloat d1, uint64_t ctrl_mask);
    void x(uint32_t target, uint64_t ctrl_mask, uint64_t ctrl_value);
    void swap(uint32_t a, uint32_t b, uint64_t ctrl_mask);
};

/* One qubit: a normalised pair of complex amplitudes (host only). */
class QState {
   public:
    QState()              = default;
    QState(const QState&) = default;
    QState(QState&&)      = default;
    ~QState()             = default;
    QState& operator=(const QState&) = default;
    QState& operator=(QState&&) = default;

    QState(const std::complex<float>& zeroState, const std::complex<float>& oneState);
    QState(const std::array<std::complex<float>, 2>& states);

    /* Bloch-sphere angles -> amplitudes */
    static std::array<std::complex<float>, 2> create_state(float polar_angle, float azimuthal_angle) {
        return {std::cos(polar_angle / 2.f),
                std::complex<float>{std::cos(azimuthal_angle), std::sin(azimuthal_angle)} *
                    std::sin(polar_angle / 2.f)};
    }

    QState& set(const std::complex<float>& zero_state, const std::complex<float>& one_state);

    bool peek_measure() const;
    bool measure();
    std::array<uint32_t, 2> profile_measure(uint32_t rep_count) const;

    float probability_true() const noexcept {
        return state_[1].real * state_[1].real + state_[1].imag * state_[1].imag;
    }
    float probability_false() const noexcept { return 1.f - probability_true(); }

    const af::cfloat& operator[](bool index) const noexcept { return state_[static_cast<int>(index)]; }
    bool operator!=(const aqs::QState& other) const noexcept {
        return state_[0] != other.state_[0] || state_[1] != other.state_[1];
    }
    bool operator==(const aqs::QState& other) const noexcept { return !(*this != other); }

    af::cfloat* data() noexcept { return state_; }
    const af::cfloat* data() const noexcept { return state_; }
    af::array to_array() const { return af::array(2, state_); }

    static const QState& zero() {
        const static QState zero_state{1.0f, 0.0f};
        return zero_state;
    }
    static const QState& one() {
        const static QState one_state{0.0f, 1.0f};
        return one_state;
    }
    static const QState& plus() {
        const static QState plus_state{0.70710678118f, 0.70710678118f};
        return plus_state;
    }
    static const QState& minus() {
        const static QState minus_state{0.70710678118f, -0.70710678118f};
        return minus_state;
    }

   private:
    af::cfloat state_[2]{{1.0f, 0.0f}, {0.0f, 0.0f}};
    void force_normalize();
};

namespace detail {
struct PlanCache;   /* engine plan for a circuit's compiled prefix */
struct DeviceState; /* engine state handle */
}  // namespace detail

/* A gate list on `qubit_count` qubits. */
class QCircuit {
   public:
    friend class QSimulator;

    QCircuit()                = delete;
    QCircuit(const QCircuit&) = default;
    QCircuit(QCircuit&&)      = default;
    ~QCircuit()               = default;
    QCircuit& operator=(const QCircuit&) = default;
    QCircuit& operator=(QCircuit&&) = default;

    QCircuit(uint32_t qubit_count);

    template<typename T>
    friend QCircuit& operator<<(QCircuit& qc, const T& gate);
    template<typename T>
    friend QCircuit& operator<<(QCircuit& qc, const std::vector<T>& gates);

    uint32_t qubit_count() const noexcept { return qubits_; }
    uint32_t state_count() const noexcept { return fast_pow2(qubits_); }

    /* The compiled prefix as a dense column-major 2^n x 2^n matrix.  The reference
     * keeps this matrix eagerly (src/quantum.cpp:159-164); here it is built on
     * demand by running the compiled ops over the columns of I on the device, and
     * only for n <= 13 (std::length_error beyond).
     * Write-through: on a circuit WITHOUT gates of at most 6 qubits the non-const accessor returns a matrix the circuit
     * owns (initially the identity); what the user writes into it is what Gate{circuit, b} / ControlGate{circuit, c, b}
     * apply (the engine's dense-matrix kernel, aqs_apply_dense), like the reference's `qc.circuit() = M`
     * (docs/USAGE.md:121-125).  On circuits with gates, writes through the reference do not change the circuit. */
    af::array& circuit();
    const af::array& circuit() const;
    /* the same, explicitly: make this (gate-less) circuit the opaque matrix m (2^n x 2^n, column-major like af::array) */
    void set_matrix(const af::array& m);
    bool opaque() const noexcept { return user_matrix_ != nullptr && gate_list_.empty(); }
    const af::array* user_matrix() const noexcept { return user_matrix_.get(); }

    auto& gate_list() noexcept { return gate_list_; }
    const auto& gate_list() const noexcept { return gate_list_; }

    std::string& representation() noexcept { return representation_; }
    const std::string& representation() const noexcept { return representation_; }

    /* Lower the gates added since the last call and (re)build the launch plan. */
    void compile();
    void clear();
    void clear_cache();

    friend bool operator==(const QCircuit& lhs, const QCircuit& rhs);

    /* ---- engine side (extensions) ---- */
    std::vector<aqs_op>& compiled_ops() {
        detach();   /* copy-on-write: circuit copies share the compiled list until one of them changes it */
        return *compiled_ops_;
    }
    const std::vector<aqs_op>& compiled_ops() const noexcept { return *compiled_ops_; }
    std::vector<DenseGateRec>& compiled_dense() {
        detach();
        return *compiled_dense_;
    }
    const std::vector<DenseGateRec>& compiled_dense() const noexcept { return *compiled_dense_; }
    std::size_t cached_index() const noexcept { return cached_index_; }
    /* ops of the whole gate list (compiled prefix + lowered tail) */
    std::vector<aqs_op> lower_all() const;

   private:
    std::vector<std::shared_ptr<QGate>> gate_list_;
    std::string representation_;
    uint32_t qubits_          = 0;
    std::size_t cached_index_ = 0;
    std::shared_ptr<std::vector<aqs_op>> compiled_ops_;
    std::shared_ptr<std::vector<DenseGateRec>> compiled_dense_;   /* opaque matrices the markers in compiled_ops_ refer to */
    std::shared_ptr<af::array> user_matrix_;                      /* write-through matrix of a gate-less circuit */
    mutable std::shared_ptr<detail::PlanCache> plan_;        /* compiled prefix */
    mutable std::shared_ptr<detail::PlanCache> tail_plan_;   /* last uncompiled tail that was simulated (reused while its ops do not change) */
    mutable std::shared_ptr<af::array> matrix_;
    void detach();
};

class QNoise {};

/* Owns the 2^n complex64 state vector (in HBM) and measures it. */
class QSimulator {
   public:
    enum class Basis : int8_t { Z, Y, X };

    explicit QSimulator(uint32_t qubit_count, const QState& initial_state = aqs::QState::zero(),
                        const QNoise& noise_generator = QNoise{});
    explicit QSimulator(uint32_t qubit_count, std::vector<QState> initial_states,
                        const QNoise& noise_generator = QNoise{});
    explicit QSimulator(uint32_t qubit_count, const af::array& statevector,
                        const QNoise& noise_generator = QNoise{});
    QSimulator(const QSimulator& other);
    QSimulator(QSimulator&& other) noexcept;
    QSimulator& operator=(const QSimulator& other);
    QSimulator& operator=(QSimulator&& other) noexcept;
    ~QSimulator();

    /* state <- kron of qubit(0..n-1) */
    void generate_statevector();
    /* state <- circuit * state */
    void simulate(const QCircuit& circuit);

    bool peek_measure(uint32_t qubit) const;
    bool measure(uint32_t qubit);
    uint32_t measure_all();
    uint32_t peek_measure_all() const;
    std::array<uint32_t, 2> profile_measure(uint32_t qubit, uint32_t rep_count) const;
    std::vector<uint32_t> profile_measure_all(uint32_t rep_count) const;

    QState& qubit(uint32_t index) noexcept {
        assert(index < qubit_count());
        return states_[index];
    }
    const QState& qubit(uint32_t index) const noexcept {
        assert(index < qubit_count());
        return states_[index];
    }

    af::cfloat state(uint32_t state) const noexcept;

    float qubit_probability_true(uint32_t qubit) const;
    float qubit_probability_false(uint32_t qubit) const { return 1.f - qubit_probability_true(qubit); }
    float state_probability(uint32_t state) const;
    std::vector<float> probabilities() const;

    uint32_t qubit_count() const noexcept { return qubits_; }
    uint32_t state_count() const noexcept { return fast_pow2(qubits_); }

    Basis get_basis() const noexcept { return basis_; }
    void set_basis(Basis basis);

    /* Host snapshot of the state vector, refreshed on every call (2^n x 1). */
    af::array& statevector();
    const af::array& statevector() const;

    /* ---- extensions ---- */
    /* outcomes of the given uniform draws (one per draw), without collapsing */
    std::vector<uint64_t> sample(const std::vector<float>& draws) const;
    double norm2() const;
    void* engine_handle() const noexcept;
    void sync() const;

   private:
    std::vector<QState> states_;
    QNoise noise_;
    uint32_t qubits_;
    Basis basis_ = Basis::Z;
    std::shared_ptr<detail::DeviceState> dev_;
    mutable std::shared_ptr<af::array> snapshot_;
};

/* Base of every gate. */
class QGate {
   public:
    virtual ~QGate() = default;
    /* Append this gate to qc's COMPILED op list (what QCircuit::compile calls;
     * the reference multiplied the gate matrix into qc.circuit() here). */
    virtual QCircuit& operator()(QCircuit& qc) const = 0;
    virtual std::string to_string() const            = 0;
    virtual uint32_t type() const noexcept           = 0;
    virtual bool check(const QCircuit& qc) const     = 0;
    virtual bool operator==(const QGate& rhs) const noexcept = 0;

    /* Engine hook: emit the primitive ops of this gate with every qubit index
     * shifted by `offset` and `ctrl_mask` (bit q = API qubit q) OR-ed into the
     * controls.  Composite gates recurse.  A user-defined gate must override it. */
    virtual void lower(OpSink& sink, uint32_t offset, uint64_t ctrl_mask) const;
    virtual std::shared_ptr<QGate> clone() const;

   protected:
    enum class GateTypes : uint32_t {
        Barrier, X, Y, Z, Hadamard, Phase, Swap, RotX, RotY, RotZ,
        CX, CY, CZ, CHadamard, CPhase, CSwap, CRotX, CRotY, CRotZ,
        CCX, Or, Circuit, ControlCircuit
    };
};

template<typename T>
QCircuit& operator<<(QCircuit& qc, const T& gate) {
    static_assert(std::is_base_of<QGate, T>::value, "Gate must inherit from QGate class");
    if (gate.check(qc)) {
        qc.representation_.append(gate.to_string());
        qc.gate_list().push_back(std::make_shared<T>(gate));
    }
    return qc;
}

template<typename T>
QCircuit& operator<<(QCircuit& qc, const std::vector<T>& gates) {
    static_assert(std::is_base_of<QGate, T>::value, "Gate must inherit from QGate class");
    for (const auto& gate : gates) qc << gate;
    return qc;
}

#define AQS_GATE_INTERFACE(Class, TypeId)                                                   \
    QCircuit& operator()(QCircuit&) const override;                                         \
    std::string to_string() const override;                                                 \
    bool check(const QCircuit&) const override;                                             \
    bool operator==(const QGate& rhs) const noexcept override;                              \
    void lower(OpSink& sink, uint32_t offset, uint64_t ctrl_mask) const override;           \
    std::shared_ptr<QGate> clone() const override { return std::make_shared<Class>(*this); } \
    uint32_t type() const noexcept override { return static_cast<uint32_t>(GateTypes::TypeId); } \
    static constexpr uint32_t static_type() noexcept { return static_cast<uint32_t>(GateTypes::TypeId); }

#define AQS_STATIC_GATE(Class)              \
    static const QCircuit& gate() {         \
        static QCircuit qc = []() {         \
            QCircuit c(1);                  \
            c << Class{0};                  \
            c.compile();                    \
            return c;                       \
        }();                                \
        return qc;                          \
    }
#define AQS_ANGLE_GATE(Class)               \
    static QCircuit gate(float angle) {     \
        QCircuit c(1);                      \
        c << Class{0, angle};               \
        c.compile();                        \
        return c;                           \
    }

/* Marker in the drawing; no effect on the state. */
class Barrier : public QGate {
   public:
    Barrier(bool visible_ = true) noexcept : visible{visible_} {}
    bool check(const QCircuit&) const override { return true; }
    QCircuit& operator()(QCircuit& qc) const override { return qc; }
    std::string to_string() const override { return visible ? "B;" : "P;"; }
    uint32_t type() const noexcept override { return static_cast<uint32_t>(GateTypes::Barrier); }
    bool operator==(const QGate& rhs) const noexcept override { return type() == rhs.type(); }
    void lower(OpSink&, uint32_t, uint64_t) const override {}
    std::shared_ptr<QGate> clone() const override { return std::make_shared<Barrier>(*this); }
    static constexpr uint32_t static_type() noexcept { return static_cast<uint32_t>(GateTypes::Barrier); }
    bool visible = true;
};

class X : public QGate {
   public:
    X(uint32_t target_qubit_) noexcept : target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(X, X)
    AQS_STATIC_GATE(X)
    uint32_t target_qubit;
};
using Not = X;

class Y : public QGate {
   public:
    Y(uint32_t target_qubit_) noexcept : target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(Y, Y)
    AQS_STATIC_GATE(Y)
    uint32_t target_qubit;
};

class Z : public QGate {
   public:
    Z(uint32_t target_qubit_) noexcept : target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(Z, Z)
    AQS_STATIC_GATE(Z)
    uint32_t target_qubit;
};

class RotX : public QGate {
   public:
    RotX(uint32_t target_qubit_, float angle_) noexcept : target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(RotX, RotX)
    AQS_ANGLE_GATE(RotX)
    uint32_t target_qubit;
    float angle;
};

class RotY : public QGate {
   public:
    RotY(uint32_t target_qubit_, float angle_) noexcept : target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(RotY, RotY)
    AQS_ANGLE_GATE(RotY)
    uint32_t target_qubit;
    float angle;
};

class RotZ : public QGate {
   public:
    RotZ(uint32_t target_qubit_, float angle_) noexcept : target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(RotZ, RotZ)
    AQS_ANGLE_GATE(RotZ)
    uint32_t target_qubit;
    float angle;
};

class H : public QGate {
   public:
    H(uint32_t target_qubit_) noexcept : target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(H, Hadamard)
    AQS_STATIC_GATE(H)
    uint32_t target_qubit;
};

class Phase : public QGate {
   public:
    Phase(uint32_t target_qubit_, float angle_) noexcept : target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(Phase, Phase)
    AQS_ANGLE_GATE(Phase)
    uint32_t target_qubit;
    float angle;
};

class Swap : public QGate {
   public:
    Swap(uint32_t target_qubit_A_, uint32_t target_qubit_B_) noexcept
        : target_qubit_A{target_qubit_A_}, target_qubit_B{target_qubit_B_} {}
    AQS_GATE_INTERFACE(Swap, Swap)
    uint32_t target_qubit_A;
    uint32_t target_qubit_B;
};

class CX : public QGate {
   public:
    CX(uint32_t control_qubit_, uint32_t target_qubit_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(CX, CX)
    uint32_t control_qubit;
    uint32_t target_qubit;
};
using CNot = CX;
using Xor  = CX;

class CY : public QGate {
   public:
    CY(uint32_t control_qubit_, uint32_t target_qubit_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(CY, CY)
    uint32_t control_qubit;
    uint32_t target_qubit;
};

class CZ : public QGate {
   public:
    CZ(uint32_t control_qubit_, uint32_t target_qubit_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(CZ, CZ)
    uint32_t control_qubit;
    uint32_t target_qubit;
};

class CH : public QGate {
   public:
    CH(uint32_t control_qubit_, uint32_t target_qubit_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(CH, CHadamard)
    uint32_t control_qubit;
    uint32_t target_qubit;
};

class CPhase : public QGate {
   public:
    CPhase(uint32_t control_qubit_, uint32_t target_qubit_, float angle_)
        : control_qubit{control_qubit_}, target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(CPhase, CPhase)
    uint32_t control_qubit;
    uint32_t target_qubit;
    float angle;
};

class CSwap : public QGate {
   public:
    CSwap(uint32_t control_qubit_, uint32_t target_qubit_A_, uint32_t target_qubit_B_) noexcept
        : control_qubit{control_qubit_}, target_qubit_A{target_qubit_A_}, target_qubit_B{target_qubit_B_} {}
    AQS_GATE_INTERFACE(CSwap, CSwap)
    uint32_t control_qubit;
    uint32_t target_qubit_A;
    uint32_t target_qubit_B;
};

class CRotX : public QGate {
   public:
    CRotX(uint32_t control_qubit_, uint32_t target_qubit_, float angle_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(CRotX, CRotX)
    uint32_t control_qubit;
    uint32_t target_qubit;
    float angle;
};

class CRotY : public QGate {
   public:
    CRotY(uint32_t control_qubit_, uint32_t target_qubit_, float angle_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(CRotY, CRotY)
    uint32_t control_qubit;
    uint32_t target_qubit;
    float angle;
};

class CRotZ : public QGate {
   public:
    CRotZ(uint32_t control_qubit_, uint32_t target_qubit_, float angle_) noexcept
        : control_qubit{control_qubit_}, target_qubit{target_qubit_}, angle{angle_} {}
    AQS_GATE_INTERFACE(CRotZ, CRotZ)
    uint32_t control_qubit;
    uint32_t target_qubit;
    float angle;
};

class CCNot : public QGate {
   public:
    CCNot(uint32_t control_qubit_A_, uint32_t control_qubit_B_, uint32_t target_qubit_)
        : control_qubit_A{control_qubit_A_}, control_qubit_B{control_qubit_B_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(CCNot, CCX)
    uint32_t control_qubit_A;
    uint32_t control_qubit_B;
    uint32_t target_qubit;
};
using And = CCNot;

class Or : public QGate {
   public:
    Or(uint32_t control_qubit_A_, uint32_t control_qubit_B_, uint32_t target_qubit_) noexcept
        : control_qubit_A{control_qubit_A_}, control_qubit_B{control_qubit_B_}, target_qubit{target_qubit_} {}
    AQS_GATE_INTERFACE(Or, Or)
    uint32_t control_qubit_A;
    uint32_t control_qubit_B;
    uint32_t target_qubit;
};

/* A whole circuit used as a gate on qubits [target_qubit_begin, +qubit_count). */
class Gate : public QGate {
   public:
    Gate(const QCircuit& circuit_, uint32_t target_qubit_begin_, std::string name = "");
    bool check(const QCircuit&) const override;
    QCircuit& operator()(QCircuit&) const override;
    std::string to_string() const override { return representation; }
    uint32_t type() const noexcept override { return static_cast<uint32_t>(GateTypes::Circuit); }
    bool operator==(const QGate& rhs) const noexcept override;
    void lower(OpSink& sink, uint32_t offset, uint64_t ctrl_mask) const override;
    std::shared_ptr<QGate> clone() const override { return std::make_shared<Gate>(*this); }
    static constexpr uint32_t static_type() noexcept { return static_cast<uint32_t>(GateTypes::Circuit); }

    std::shared_ptr<QCircuit> internal_circuit;
    std::string representation;
    uint32_t qubit_count;
    uint32_t target_qubit_begin;
};

/* The same, active only where control_qubit is |1>. */
class ControlGate : public QGate {
   public:
    ControlGate(const QCircuit& circuit_, uint32_t control_qubit_, uint32_t target_qubit_begin_,
                std::string name = "");
    bool check(const QCircuit&) const override;
    QCircuit& operator()(QCircuit&) const override;
    std::string to_string() const override { return representation; }
    uint32_t type() const noexcept override { return static_cast<uint32_t>(GateTypes::ControlCircuit); }
    bool operator==(const QGate& rhs) const noexcept override;
    void lower(OpSink& sink, uint32_t offset, uint64_t ctrl_mask) const override;
    std::shared_ptr<QGate> clone() const override { return std::make_shared<ControlGate>(*this); }
    static constexpr uint32_t static_type() noexcept { return static_cast<uint32_t>(GateTypes::ControlCircuit); }

    std::shared_ptr<QCircuit> internal_circuit;
    std::string representation;
    uint32_t qubit_count;
    uint32_t control_qubit;
    uint32_t target_qubit_begin;
};

#undef AQS_GATE_INTERFACE
#undef AQS_STATIC_GATE
#undef AQS_ANGLE_GATE

/* single-qubit host operations */
QState X_op(const QState& state);
QState Y_op(const QState& state);
QState Z_op(const QState& state);
QState RotateX_op(const QState& state, float angle);
QState RotateY_op(const QState& state, float angle);
QState RotateZ_op(const QState& state, float angle);
QState Hadamard_op(const QState& state);
QState Phase_op(const QState& state, float angle);

}  // namespace aqs

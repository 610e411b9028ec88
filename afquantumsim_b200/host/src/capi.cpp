// capi.cpp — flat C wrapper over the aqs:: C++ host layer, for ctypes
// (afquantumsim_b200/aqs.py).  C++ exceptions become status codes that keep the
// reference's exception class:  1 out_of_range, 2 invalid_argument,
// 3 domain_error, 4 runtime_error, 5 anything else.
#include <cstring>
#include <string>

#include "quantum.h"
#include "quantum_algo.h"
#include "quantum_gates.h"
#include "quantum_visuals.h"

using namespace aqs;

static thread_local std::string g_err;

template<typename F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::out_of_range& e) { g_err = e.what(); return 1;
    } catch (const std::domain_error& e) { g_err = e.what(); return 3;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 2;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 4;
    } catch (const std::exception& e) { g_err = e.what(); return 5; }
}

// QState from already-normalised amplitudes (the Python QState normalises once,
// exactly like QState::force_normalize; do not normalise a second time here)
static QState raw_state(const float* zo) {
    QState q;
    q.data()[0] = af::cfloat{zo[0], zo[1]};
    q.data()[1] = af::cfloat{zo[2], zo[3]};
    return q;
}

static int copy_out(const std::string& s, char* buf, size_t cap, size_t* needed) {
    if (needed) *needed = s.size() + 1;
    if (buf && cap) {
        size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return 0;
}

extern "C" {

const char* aqsh_last_error() { return g_err.c_str(); }

int aqsh_initialize(int device) {
    return guarded([&] {
        std::string d = std::to_string(device);
        char prog[] = "aqs";
        char* argv[2] = {prog, const_cast<char*>(d.c_str())};
        initialize(2, argv);
    });
}
void aqsh_set_seed(uint64_t seed) { set_seed(seed); }
void aqsh_set_fusion(int on) { set_fusion(on != 0); }
void aqsh_set_jit_min_qubits(int n) { set_jit_min_qubits(n); }
void aqsh_jit_wait() { jit_wait(); }
int aqsh_get_fusion() { return get_fusion() ? 1 : 0; }
void aqsh_clear_circuit_cache() { clear_circuit_cache(); }

// ---- circuits ------------------------------------------------------------------
int aqsh_circuit_new(uint32_t qubits, void** out) {
    return guarded([&] { *out = new QCircuit(qubits); });
}
void aqsh_circuit_free(void* c) { delete static_cast<QCircuit*>(c); }
int aqsh_circuit_copy(void* c, void** out) {
    return guarded([&] { *out = new QCircuit(*static_cast<QCircuit*>(c)); });
}

// name = reference class name; q = qubit arguments in constructor order
int aqsh_circuit_add(void* c, const char* name, const uint32_t* q, int nq, float angle) {
    return guarded([&] {
        QCircuit& qc = *static_cast<QCircuit*>(c);
        const std::string n(name);
        auto need = [&](int k) { if (nq != k) throw std::invalid_argument{"wrong number of qubit arguments for " + n}; };
        if (n == "Barrier") { qc << Barrier(nq ? q[0] != 0 : true); }
        else if (n == "X" || n == "Not") { need(1); qc << X(q[0]); }
        else if (n == "Y") { need(1); qc << Y(q[0]); }
        else if (n == "Z") { need(1); qc << Z(q[0]); }
        else if (n == "H") { need(1); qc << H(q[0]); }
        else if (n == "Phase") { need(1); qc << Phase(q[0], angle); }
        else if (n == "RotX") { need(1); qc << RotX(q[0], angle); }
        else if (n == "RotY") { need(1); qc << RotY(q[0], angle); }
        else if (n == "RotZ") { need(1); qc << RotZ(q[0], angle); }
        else if (n == "Swap") { need(2); qc << Swap(q[0], q[1]); }
        else if (n == "CX" || n == "CNot" || n == "Xor") { need(2); qc << CX(q[0], q[1]); }
        else if (n == "CY") { need(2); qc << CY(q[0], q[1]); }
        else if (n == "CZ") { need(2); qc << CZ(q[0], q[1]); }
        else if (n == "CH") { need(2); qc << CH(q[0], q[1]); }
        else if (n == "CPhase") { need(2); qc << CPhase(q[0], q[1], angle); }
        else if (n == "CRotX") { need(2); qc << CRotX(q[0], q[1], angle); }
        else if (n == "CRotY") { need(2); qc << CRotY(q[0], q[1], angle); }
        else if (n == "CRotZ") { need(2); qc << CRotZ(q[0], q[1], angle); }
        else if (n == "CSwap") { need(3); qc << CSwap(q[0], q[1], q[2]); }
        else if (n == "CCNot" || n == "And") { need(3); qc << CCNot(q[0], q[1], q[2]); }
        else if (n == "Or") { need(3); qc << Or(q[0], q[1], q[2]); }
        else throw std::invalid_argument{"unknown gate class " + n};
    });
}
int aqsh_circuit_add_gate(void* c, void* inner, uint32_t begin, const char* name) {
    return guarded([&] { *static_cast<QCircuit*>(c) << Gate(*static_cast<QCircuit*>(inner), begin, name ? name : ""); });
}
int aqsh_circuit_add_control_gate(void* c, void* inner, uint32_t control, uint32_t begin, const char* name) {
    return guarded([&] {
        *static_cast<QCircuit*>(c) << ControlGate(*static_cast<QCircuit*>(inner), control, begin, name ? name : "");
    });
}
int aqsh_circuit_compile(void* c) { return guarded([&] { static_cast<QCircuit*>(c)->compile(); }); }
int aqsh_circuit_clear(void* c) { return guarded([&] { static_cast<QCircuit*>(c)->clear(); }); }
int aqsh_circuit_clear_cache(void* c) { return guarded([&] { static_cast<QCircuit*>(c)->clear_cache(); }); }
uint32_t aqsh_circuit_qubits(void* c) { return static_cast<QCircuit*>(c)->qubit_count(); }
uint64_t aqsh_circuit_gate_count(void* c) { return static_cast<QCircuit*>(c)->gate_list().size(); }
uint64_t aqsh_circuit_cached_index(void* c) { return static_cast<QCircuit*>(c)->cached_index(); }
int aqsh_circuit_representation(void* c, char* buf, size_t cap, size_t* needed) {
    return copy_out(static_cast<QCircuit*>(c)->representation(), buf, cap, needed);
}
// lowered op list of the whole circuit (records are struct aqs_op, 64 bytes each)
// opaque circuit: m = 2^n x 2^n complex64, ROW-major (numpy's default); stored column-major like af::array
int aqsh_circuit_set_matrix(void* c, const float* m, uint32_t dim) {
    return guarded([&] {
        af::array a(static_cast<long long>(dim), static_cast<long long>(dim), af::c32);
        for (uint32_t r = 0; r < dim; ++r)
            for (uint32_t col = 0; col < dim; ++col)
                a.data()[static_cast<std::size_t>(col) * dim + r] = af::cfloat{m[2 * (static_cast<std::size_t>(r) * dim + col)], m[2 * (static_cast<std::size_t>(r) * dim + col) + 1]};
        static_cast<QCircuit*>(c)->set_matrix(a);
    });
}
int aqsh_circuit_ops(void* c, void* out, uint64_t cap, uint64_t* count) {
    return guarded([&] {
        auto ops = static_cast<QCircuit*>(c)->lower_all();
        *count   = ops.size();
        if (out && cap >= ops.size()) std::memcpy(out, ops.data(), ops.size() * 64);
    });
}
int aqsh_circuit_matrix(void* c, void* out_c32) {
    return guarded([&] {
        const af::array& m = static_cast<const QCircuit*>(c)->circuit();
        std::memcpy(out_c32, m.data(), static_cast<size_t>(m.elements()) * 8);
    });
}
int aqsh_circuit_text_image(void* c, void* sim, char* buf, size_t cap, size_t* needed) {
    return guarded([&] {
        copy_out(gen_circuit_text_image(*static_cast<QCircuit*>(c), *static_cast<QSimulator*>(sim)), buf, cap, needed);
    });
}
int aqsh_schematic_text_image(const char* schematic, char* buf, size_t cap, size_t* needed) {
    return guarded([&] { copy_out(gen_circuit_text_image(std::string(schematic)), buf, cap, needed); });
}

// ---- builders ----------------------------------------------------------------------
int aqsh_group_gate(uint32_t qubits, const uint32_t* t, int nt, void* gate, void** out) {
    return guarded([&] { *out = new QCircuit(Group_Gate(qubits, std::vector<uint32_t>(t, t + nt), *static_cast<QCircuit*>(gate))); });
}
int aqsh_control_group_gate(uint32_t qubits, uint32_t control, const uint32_t* t, int nt, void* gate, void** out) {
    return guarded([&] {
        *out = new QCircuit(Control_Group_Gate(qubits, control, std::vector<uint32_t>(t, t + nt), *static_cast<QCircuit*>(gate)));
    });
}
int aqsh_ncontrol_gate_range(uint32_t qubits, uint32_t cbegin, uint32_t ccount, uint32_t tbegin, void* gate, void** out) {
    return guarded([&] { *out = new QCircuit(NControl_Gate(qubits, cbegin, ccount, tbegin, *static_cast<QCircuit*>(gate))); });
}
int aqsh_ncontrol_gate_list(uint32_t qubits, const uint32_t* c, int nc, uint32_t tbegin, void* gate, void** out) {
    return guarded([&] {
        *out = new QCircuit(NControl_Gate(qubits, std::vector<uint32_t>(c, c + nc), tbegin, *static_cast<QCircuit*>(gate)));
    });
}
int aqsh_rewire_gate(uint32_t qubits, const uint32_t* pos, int np, void* gate, void** out) {
    return guarded([&] { *out = new QCircuit(Rewire_Gate(qubits, std::vector<uint32_t>(pos, pos + np), *static_cast<QCircuit*>(gate))); });
}
int aqsh_adjoint_gate(void* gate, void** out) {
    return guarded([&] { *out = new QCircuit(Adjoint_Gate(*static_cast<QCircuit*>(gate))); });
}
int aqsh_fourier_transform(uint32_t qubits, int inverse, void** out) {
    return guarded([&] { *out = new QCircuit(inverse ? inverse_fourier_transform(qubits) : fourier_transform(qubits)); });
}
int aqsh_grover_oracle(uint32_t qubits, uint32_t marked, void** out) {
    return guarded([&] { *out = new QCircuit(grover_oracle(qubits, marked)); });
}
int aqsh_grover_search(uint32_t qubits, void* oracle, uint32_t iterations, const char* name, void** out) {
    return guarded([&] { *out = new QCircuit(grover_search(qubits, *static_cast<QCircuit*>(oracle), iterations, name ? name : "")); });
}
int aqsh_grover_iteration(uint32_t qubits, void* oracle, uint32_t iterations, void** out) {
    return guarded([&] { *out = new QCircuit(grover_iteration(qubits, *static_cast<QCircuit*>(oracle), iterations)); });
}

// ---- simulator -----------------------------------------------------------------------
int aqsh_sim_new(uint32_t qubits, void** out) { return guarded([&] { *out = new QSimulator(qubits); }); }
// q: n x 2 complex64, already normalised
int aqsh_sim_new_states(uint32_t qubits, const float* q, void** out) {
    return guarded([&] {
        std::vector<QState> st;
        for (uint32_t i = 0; i < qubits; ++i)
            st.push_back(raw_state(q + 4 * i));
        *out = new QSimulator(qubits, std::move(st));
    });
}
int aqsh_sim_new_vector(uint32_t qubits, const float* amps, void** out) {
    return guarded([&] {
        af::array v(static_cast<long long>(1) << qubits, reinterpret_cast<const af::cfloat*>(amps));
        *out = new QSimulator(qubits, v);
    });
}
int aqsh_sim_clone(void* s, void** out) { return guarded([&] { *out = new QSimulator(*static_cast<QSimulator*>(s)); }); }
void aqsh_sim_free(void* s) { delete static_cast<QSimulator*>(s); }
int aqsh_sim_set_qubit(void* s, uint32_t i, const float* zo) {
    return guarded([&] {
        QSimulator& qs = *static_cast<QSimulator*>(s);
        if (i >= qs.qubit_count()) throw std::out_of_range{"qubit index out of range"};
        qs.qubit(i) = raw_state(zo);
    });
}
int aqsh_sim_get_qubit(void* s, uint32_t i, float* zo) {
    return guarded([&] {
        QSimulator& qs = *static_cast<QSimulator*>(s);
        if (i >= qs.qubit_count()) throw std::out_of_range{"qubit index out of range"};
        const QState& q = qs.qubit(i);
        zo[0] = q[0].real; zo[1] = q[0].imag; zo[2] = q[1].real; zo[3] = q[1].imag;
    });
}
int aqsh_sim_generate_statevector(void* s) { return guarded([&] { static_cast<QSimulator*>(s)->generate_statevector(); }); }
int aqsh_sim_simulate(void* s, void* c) {
    return guarded([&] { static_cast<QSimulator*>(s)->simulate(*static_cast<QCircuit*>(c)); });
}
int aqsh_sim_peek_measure(void* s, uint32_t q, int* out) { return guarded([&] { *out = static_cast<QSimulator*>(s)->peek_measure(q); }); }
int aqsh_sim_measure(void* s, uint32_t q, int* out) { return guarded([&] { *out = static_cast<QSimulator*>(s)->measure(q); }); }
int aqsh_sim_measure_all(void* s, uint32_t* out) { return guarded([&] { *out = static_cast<QSimulator*>(s)->measure_all(); }); }
int aqsh_sim_peek_measure_all(void* s, uint32_t* out) { return guarded([&] { *out = static_cast<QSimulator*>(s)->peek_measure_all(); }); }
int aqsh_sim_profile_measure(void* s, uint32_t q, uint32_t reps, uint32_t* out2) {
    return guarded([&] { auto r = static_cast<QSimulator*>(s)->profile_measure(q, reps); out2[0] = r[0]; out2[1] = r[1]; });
}
int aqsh_sim_profile_measure_all(void* s, uint32_t reps, uint32_t* out) {
    return guarded([&] {
        auto r = static_cast<QSimulator*>(s)->profile_measure_all(reps);
        std::memcpy(out, r.data(), r.size() * sizeof(uint32_t));
    });
}
int aqsh_sim_sample(void* s, const float* u, uint64_t n, uint64_t* out) {
    return guarded([&] {
        auto r = static_cast<QSimulator*>(s)->sample(std::vector<float>(u, u + n));
        std::memcpy(out, r.data(), r.size() * sizeof(uint64_t));
    });
}
int aqsh_sim_qubit_probability_true(void* s, uint32_t q, float* out) {
    return guarded([&] { *out = static_cast<QSimulator*>(s)->qubit_probability_true(q); });
}
int aqsh_sim_state_probability(void* s, uint32_t k, float* out) {
    return guarded([&] { *out = static_cast<QSimulator*>(s)->state_probability(k); });
}
int aqsh_sim_probabilities(void* s, float* out) {
    return guarded([&] { auto p = static_cast<QSimulator*>(s)->probabilities(); std::memcpy(out, p.data(), p.size() * 4); });
}
int aqsh_sim_state(void* s, uint32_t k, float* out2) {
    return guarded([&] {
        QSimulator& qs = *static_cast<QSimulator*>(s);
        if (k >= qs.state_count()) throw std::out_of_range{"state index out of range"};
        af::cfloat v = qs.state(k); out2[0] = v.real; out2[1] = v.imag;
    });
}
int aqsh_sim_statevector(void* s, float* out) {
    return guarded([&] {
        const af::array& v = static_cast<const QSimulator*>(s)->statevector();
        std::memcpy(out, v.data(), static_cast<size_t>(v.elements()) * 8);
    });
}
int aqsh_sim_set_basis(void* s, int basis) {
    return guarded([&] { static_cast<QSimulator*>(s)->set_basis(static_cast<QSimulator::Basis>(basis)); });
}
int aqsh_sim_get_basis(void* s) { return static_cast<int>(static_cast<QSimulator*>(s)->get_basis()); }
uint32_t aqsh_sim_qubits(void* s) { return static_cast<QSimulator*>(s)->qubit_count(); }
void* aqsh_sim_engine_handle(void* s) { return static_cast<QSimulator*>(s)->engine_handle(); }
int aqsh_sim_sync(void* s) { return guarded([&] { static_cast<QSimulator*>(s)->sync(); }); }
int aqsh_sim_norm2(void* s, double* out) { return guarded([&] { *out = static_cast<QSimulator*>(s)->norm2(); }); }

}  // extern "C"

// quantum_algo.cpp — Grover and QFT circuit builders.
// Gate sequences follow the reference's src/quantum_algo.cpp:16-129 exactly
// (order matters for bit-level reproducibility of the flattened op list).
#include "quantum_algo.h"

#include "quantum_gates.h"

#include <stdexcept>

namespace aqs {

static QCircuit all_ones_phase_flip(uint32_t n) {
    // Z on the last qubit controlled by all the others
    return NControl_Gate(n, 0, n - 1, n - 1, Z::gate());
}

QCircuit grover_oracle(uint32_t search_qubits, uint32_t marked_state, bool compile) {
    if (marked_state >= fast_pow2(search_qubits))
        throw std::invalid_argument{"Marked state should be in the range [0, 2^search_qubits)"};
    QCircuit qc{search_qubits};
    auto flip_zero_bits = [&]() {
        for (uint32_t i = 0; i < search_qubits; ++i)
            if (!(marked_state & (1u << i))) qc << X(i);
    };
    flip_zero_bits();
    qc << Gate{all_ones_phase_flip(search_qubits), 0};
    flip_zero_bits();
    if (compile) qc.compile();
    return qc;
}

static void append_diffuser(QCircuit& qc, uint32_t n) {
    for (uint32_t j = 0; j < n; ++j) qc << H{j};
    for (uint32_t j = 0; j < n; ++j) qc << X{j};
    qc << Gate{all_ones_phase_flip(n), 0};
    for (uint32_t j = 0; j < n; ++j) {
        qc << X(j);
        qc << H(j);
    }
}

QCircuit grover_search(uint32_t search_qubits, const QCircuit& oracle, uint32_t iterations, std::string oracle_name,
                       bool compile) {
    if (oracle.qubit_count() < search_qubits)
        throw std::invalid_argument{"Cannot use given oracle for this qubit circuit"};
    QCircuit qc(oracle.qubit_count());
    for (uint32_t i = 0; i < search_qubits; ++i) qc << H{i};
    for (uint32_t i = 0; i < iterations; ++i) {
        qc << Barrier{false};
        qc << Gate{oracle, 0, oracle_name};
        qc << Barrier{false};
        append_diffuser(qc, search_qubits);
    }
    if (compile) qc.compile();
    return qc;
}

QCircuit grover_iteration(uint32_t search_qubits, const QCircuit& oracle, uint32_t iterations, bool compile) {
    QCircuit qc{search_qubits};
    for (uint32_t i = 0; i < iterations; ++i) {
        qc << Gate{oracle, 0};
        append_diffuser(qc, search_qubits);
    }
    if (compile) qc.compile();
    return qc;
}

QCircuit fourier_transform(uint32_t qubits, bool compile) {
    QCircuit qc(qubits);
    for (int32_t i = static_cast<int32_t>(qubits) - 1; i >= 0; --i) {
        qc << H{static_cast<uint32_t>(i)};
        for (int32_t j = 0; j < i; ++j)
            qc << CPhase{static_cast<uint32_t>(j), static_cast<uint32_t>(i), aqs::pi / (1 << (i - j))};
    }
    if (compile) qc.compile();
    return qc;
}

QCircuit inverse_fourier_transform(uint32_t qubits, bool compile) {
    QCircuit qc(qubits);
    for (uint32_t i = 0; i < qubits; ++i) {
        for (int32_t j = static_cast<int32_t>(i) - 1; j >= 0; --j)
            qc << CPhase{static_cast<uint32_t>(j), i, -aqs::pi / (1 << (i - j))};
        qc << H{i};
    }
    if (compile) qc.compile();
    return qc;
}

}  // namespace aqs

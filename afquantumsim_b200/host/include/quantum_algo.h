/* quantum_algo.h — algorithm circuit builders (reference include/quantum_algo.h).
 * In scope: Grover and the quantum Fourier transform (BASELINE configs 2, 4, 5).
 * The reference's Hamiltonian / VQE helpers (src/quantum_algo.cpp:131-468) depend
 * on NLopt and on dense af::array algebra and are outside the state-vector hot
 * path; they are not provided. */
#pragma once
#include "quantum.h"

namespace aqs {

QCircuit grover_iteration(uint32_t search_qubits, const QCircuit& oracle, uint32_t iterations, bool compile = false);
QCircuit grover_search(uint32_t search_qubits, const QCircuit& oracle, uint32_t iterations,
                       std::string oracle_name = "", bool compile = false);
/* marks `marked_state` (bit i of it <-> qubit i) with a phase flip */
QCircuit grover_oracle(uint32_t search_qubits, uint32_t marked_state, bool compile = false);

/* QFT without the final swaps */
QCircuit fourier_transform(uint32_t qubits, bool compile = false);
QCircuit inverse_fourier_transform(uint32_t qubits, bool compile = false);

}  // namespace aqs

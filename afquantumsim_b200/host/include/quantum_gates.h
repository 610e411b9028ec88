/* quantum_gates.h — composite-gate builders (reference include/quantum_gates.h).
 * All host-side: they assemble circuits out of Gate / ControlGate / Swap; the
 * engine later flattens the nesting into primitive ops with one accumulated
 * control mask, so an N-controlled gate costs one kernel, not a dense matrix. */
#pragma once
#include "quantum.h"

namespace aqs {

/* the 1-qubit `gate` on each of target_qubits */
QCircuit Group_Gate(uint32_t qubits, std::vector<uint32_t> target_qubits, QCircuit gate, bool compile = false);

/* the 1-qubit `gate` on each of target_qubits, all controlled by control_qubit */
QCircuit Control_Group_Gate(uint32_t qubits, uint32_t control_qubit, std::vector<uint32_t> target_qubits,
                            const QCircuit& gate, bool compile = false);

/* `gate` at target_qubit_begin controlled by control_qubit_count contiguous qubits
 * starting at control_qubit_begin (all above the target block) */
QCircuit NControl_Gate(uint32_t qubits, uint32_t control_qubit_begin, uint32_t control_qubit_count,
                       uint32_t target_qubit_begin, const QCircuit& gate, bool compile = false);

/* `gate` at target_qubit_begin controlled by an arbitrary set of qubits */
QCircuit NControl_Gate(uint32_t qubits, std::vector<uint32_t> control_qubits, uint32_t target_qubit_begin,
                       const QCircuit& gate, bool compile = false);

/* `gate` with its qubit i moved to new_qubit_positions[i], via a swap network */
QCircuit Rewire_Gate(uint32_t qubits, const std::vector<uint32_t>& new_qubit_positions, const QCircuit& gate,
                     bool compile = false);

/* the inverse circuit.  Unlike the reference (src/quantum_gates.cpp:233-338) the
 * gates of the result are copies: the source circuit's angles are not touched. */
QCircuit Adjoint_Gate(const QCircuit& gate);

}  // namespace aqs

/* quantum_visuals.h — text output helpers (reference include/quantum_visuals.h).
 * Pure host string work; only print_statevector / print_circuit_matrix read
 * device data (through QSimulator::statevector() / QCircuit::circuit()). */
#pragma once
#include "quantum.h"

namespace aqs {

void print_state(const QState& state);
void print_circuit_matrix(const QCircuit& circuit);
void print_statevector(const QSimulator& simulator);
void print_profile(const std::array<uint32_t, 2>& profile);
void print_profile(const std::vector<uint32_t>& profile);

/* Draw a circuit as UTF-8 box art.  `schematic` follows the circuit string
 * grammar: "n;" then "i,v;" per qubit (v = 0 or 1), then the gate statements
 * "Name,<#controls>,<#targets>:q,q,...;" and barriers "B;" / "P;". */
std::string gen_circuit_text_image(const QCircuit& circuit, const QSimulator& simulator);
std::string gen_circuit_text_image(std::string schematic);
void print_circuit_text_image(const QCircuit& circuit, const QSimulator& simulator);

}  // namespace aqs

// quantum_gates.cpp — composite-gate builders (host only).
// Same argument checks, exceptions and resulting circuits as the reference's
// src/quantum_gates.cpp; see the per-function notes for the one deliberate
// behavioural difference (Adjoint_Gate copies gates instead of mutating them).
#include "quantum_gates.h"

#include <algorithm>
#include <stdexcept>
#include <unordered_set>

namespace aqs {

QCircuit Group_Gate(uint32_t qubits, std::vector<uint32_t> target_qubits, QCircuit gate, bool compile) {
    if (gate.qubit_count() != 1) throw std::invalid_argument{"Gate not supported"};
    std::sort(target_qubits.begin(), target_qubits.end());
    if (target_qubits.size() > qubits)
        throw std::invalid_argument{"Cannot add more target qubits than there are total qubits"};
    gate.compile();
    QCircuit qc{qubits};
    for (uint32_t q : target_qubits) qc << Gate{gate, q};
    if (compile) qc.compile();
    return qc;
}

QCircuit Control_Group_Gate(uint32_t qubits, uint32_t control_qubit, std::vector<uint32_t> target_qubits,
                            const QCircuit& gate, bool compile) {
    if (control_qubit >= qubits) throw std::invalid_argument{"Invalid control qubit position"};
    if (gate.qubit_count() != 1) throw std::invalid_argument{"Gate not supported"};
    std::sort(target_qubits.begin(), target_qubits.end());
    if (target_qubits.size() >= qubits) throw std::invalid_argument{"Cannot add control gate at the given position"};

    auto pivot = std::lower_bound(target_qubits.cbegin(), target_qubits.cend(), control_qubit);
    if (pivot != target_qubits.cend() && *pivot == control_qubit)
        throw std::invalid_argument{"Cannot add control gate at the target qubit positions"};

    QCircuit qc(qubits);
    // targets above the control, then targets below it: each side becomes one
    // controlled block anchored at its first target (as the reference does, the
    // j-th target of a side sits on the block's j-th qubit)
    auto emit_side = [&](std::vector<uint32_t>::const_iterator first, std::vector<uint32_t>::const_iterator last) {
        const auto count = static_cast<uint32_t>(std::distance(first, last));
        if (count == 0) return;
        QCircuit block(count);
        for (uint32_t j = 0; j < count; ++j) block << Gate(gate, j);
        qc << ControlGate(block, control_qubit, *first);
    };
    emit_side(target_qubits.cbegin(), pivot);
    emit_side(pivot, target_qubits.cend());
    if (compile) qc.compile();
    return qc;
}

QCircuit NControl_Gate(uint32_t qubits, uint32_t control_qubit_begin, uint32_t control_qubit_count,
                       uint32_t target_qubit_begin, const QCircuit& gate, bool compile) {
    if (gate.qubit_count() >= qubits) throw std::invalid_argument{"Gate not supported"};
    if (control_qubit_count == 0) throw std::invalid_argument{"The number of control qubits must be at least 1"};
    if ((target_qubit_begin + gate.qubit_count()) > qubits)
        throw std::invalid_argument{"Invalid target qubit_begin position"};
    if ((control_qubit_begin + control_qubit_count) > target_qubit_begin)
        throw std::invalid_argument{"Invalid control_qubit position"};

    QCircuit qc(qubits);
    if (control_qubit_count == 1) {
        qc << ControlGate(gate, control_qubit_begin, target_qubit_begin);
    } else {
        // innermost: last control (qubit 0 of the block) -> gate; then wrap one
        // control at a time on top
        const uint32_t gap = target_qubit_begin - control_qubit_begin - control_qubit_count;
        QCircuit nest(gate.qubit_count() + 1 + gap);
        nest << ControlGate(gate, 0, gap + 1);
        for (uint32_t i = 0; i + 1 < control_qubit_count; ++i) {
            QCircuit wrapped(nest.qubit_count() + 1);
            wrapped << ControlGate(nest, 0, 1);
            nest = std::move(wrapped);
        }
        qc << Gate(nest, control_qubit_begin);
    }
    if (compile) qc.compile();
    return qc;
}

QCircuit NControl_Gate(uint32_t qubits, std::vector<uint32_t> control_qubits, uint32_t target_qubit_begin,
                       const QCircuit& gate, bool compile) {
    if (control_qubits.size() == 0) throw std::invalid_argument{"Number of control qubits must be at least one"};
    if (gate.qubit_count() + control_qubits.size() > qubits)
        throw std::invalid_argument{"Invalid number of qubits for number of control of qubits and gate count"};
    if (target_qubit_begin + gate.qubit_count() > qubits)
        throw std::invalid_argument{"Invalid target qubit begin position"};

    std::sort(control_qubits.begin(), control_qubits.end());
    if (control_qubits.back() >= qubits) throw std::invalid_argument{"Cannot add control gate at the given position"};

    auto pivot = std::lower_bound(control_qubits.begin(), control_qubits.end(), target_qubit_begin);
    if (pivot != control_qubits.end() && *pivot < target_qubit_begin + gate.qubit_count())
        throw std::invalid_argument{"Cannot add control gate at the target qubit positions"};

    const std::vector<uint32_t> above(control_qubits.begin(), pivot);   // smaller indices than the target
    const std::vector<uint32_t> below(pivot, control_qubits.end());

    QCircuit current = gate;   // always anchored at its own qubit 0 = target_qubit_begin

    // controls below the target block: grow the circuit downwards to each one
    for (std::size_t i = 0; i < below.size(); ++i) {
        const bool last      = (i + 1 == below.size());
        const uint32_t width = last ? qubits - target_qubit_begin : below[i] - target_qubit_begin + 1;
        QCircuit grown(width);
        grown << ControlGate(current, below[i] - target_qubit_begin, 0);
        current = std::move(grown);
    }
    // controls above: grow upwards, nearest control first
    if (!above.empty()) {
        uint32_t top = target_qubit_begin;   // outer index of current's qubit 0
        for (std::size_t i = above.size(); i-- > 1;) {
            const uint32_t c = above[i];
            QCircuit grown(qubits - c);
            grown << ControlGate{current, 0, top - c};
            top     = c;
            current = std::move(grown);
        }
        QCircuit grown(qubits);
        grown << ControlGate(current, above.front(), qubits - current.qubit_count());
        current = std::move(grown);
    }
    if (compile) current.compile();
    return current;
}

QCircuit Rewire_Gate(uint32_t qubits, const std::vector<uint32_t>& new_qubit_positions, const QCircuit& gate,
                     bool compile) {
    if (new_qubit_positions.size() != gate.qubit_count())
        throw std::invalid_argument{"New qubit positions must map all the qubits in the gate"};
    if (gate.qubit_count() > qubits) throw std::domain_error{"Cannot rewire circuit to a lower number of qubits"};
    {
        std::vector<uint32_t> sorted = new_qubit_positions;
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end())
            throw std::invalid_argument{"Cannot rewire multiple qubits to the same qubit"};
    }
    // decompose the permutation into cycles; each cycle is a chain of swaps
    std::vector<Swap> swaps;
    std::unordered_set<uint32_t> visited;
    for (uint32_t i = 0; i < new_qubit_positions.size(); ++i) {
        if (!visited.insert(i).second) continue;
        uint32_t current = i;
        while (i != new_qubit_positions[current]) {
            swaps.emplace_back(current, new_qubit_positions[current]);
            current = new_qubit_positions[current];
            visited.insert(current);
        }
    }
    QCircuit qc{qubits};
    for (const auto& sw : swaps) qc << sw;
    qc << Gate{gate, 0};
    for (auto it = swaps.rbegin(); it != swaps.rend(); ++it) qc << *it;
    if (compile) qc.compile();
    return qc;
}

QCircuit Adjoint_Gate(const QCircuit& gate) {
    QCircuit qc{gate.qubit_count()};
    const auto& src = gate.gate_list();
    for (auto it = src.rbegin(); it != src.rend(); ++it) {
        std::shared_ptr<QGate> g = (*it)->clone();
        const uint32_t t         = g->type();
        if (t == RotX::static_type()) { auto& r = *static_cast<RotX*>(g.get()); r.angle = -r.angle; }
        else if (t == RotY::static_type()) { auto& r = *static_cast<RotY*>(g.get()); r.angle = -r.angle; }
        else if (t == RotZ::static_type()) { auto& r = *static_cast<RotZ*>(g.get()); r.angle = -r.angle; }
        else if (t == Phase::static_type()) { auto& r = *static_cast<Phase*>(g.get()); r.angle = -r.angle; }
        else if (t == CRotX::static_type()) { auto& r = *static_cast<CRotX*>(g.get()); r.angle = -r.angle; }
        else if (t == CRotY::static_type()) { auto& r = *static_cast<CRotY*>(g.get()); r.angle = -r.angle; }
        else if (t == CRotZ::static_type()) { auto& r = *static_cast<CRotZ*>(g.get()); r.angle = -r.angle; }
        else if (t == CPhase::static_type()) { auto& r = *static_cast<CPhase*>(g.get()); r.angle = -r.angle; }
        else if (t == Gate::static_type()) {
            auto& r            = *static_cast<Gate*>(g.get());
            r.internal_circuit = std::make_shared<QCircuit>(Adjoint_Gate(*r.internal_circuit));
        } else if (t == ControlGate::static_type()) {
            auto& r            = *static_cast<ControlGate*>(g.get());
            r.internal_circuit = std::make_shared<QCircuit>(Adjoint_Gate(*r.internal_circuit));
        } else if (t == Barrier::static_type() || t == X::static_type() || t == Y::static_type() ||
                   t == Z::static_type() || t == H::static_type() || t == Swap::static_type() ||
                   t == CSwap::static_type() || t == CX::static_type() || t == CY::static_type() ||
                   t == CZ::static_type() || t == CH::static_type() || t == CCNot::static_type() ||
                   t == Or::static_type()) {
            // self-adjoint
        } else {
            throw std::runtime_error{"Unknown unsupported gate cannot be adjoint"};
        }
        qc.gate_list().push_back(std::move(g));
    }
    // statements of the string representation in reverse order
    const std::string& rep = gate.representation();
    std::string reversed;
    std::size_t prev = 0;
    for (std::size_t mark = rep.find(';', prev); mark != std::string::npos; prev = mark + 1, mark = rep.find(';', prev))
        reversed = rep.substr(prev, mark - prev + 1) + reversed;
    qc.representation() = reversed;
    qc.compile();
    return qc;
}

}  // namespace aqs

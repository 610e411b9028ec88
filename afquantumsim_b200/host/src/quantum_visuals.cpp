// quantum_visuals.cpp — printing helpers and the text circuit renderer.
// Host-only string work.  Output format and validation follow the reference's
// src/quantum_visuals.cpp (golden drawing: docs/USAGE.md:533-540); the renderer
// itself is written around a per-gate "role of every wire in the gate's span"
// table rather than the reference's merged index walk.
#include "quantum_visuals.h"

#include <algorithm>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace aqs {

void print_state(const QState& state) {
    std::cout << std::setprecision(4) << std::fixed << std::showpos;
    std::cout << state[0] << " |0> + " << state[1] << " |1>\n";
    std::cout << std::setprecision(7) << std::defaultfloat << std::noshowpos;
}

void print_statevector(const QSimulator& simulator) {
    const int qubits = static_cast<int>(simulator.qubit_count());
    const int states = static_cast<int>(simulator.state_count());
    std::vector<af::cfloat> vals(static_cast<std::size_t>(states));
    simulator.statevector().host(vals.data());
    std::cout << std::setprecision(3) << std::fixed << std::showpos;
    for (int i = 0; i < states - 1; ++i)
        std::cout << vals[i] << "|" << binary_string(i, qubits) << "> + " << ((i + 1) % 4 ? "" : "\n");
    std::cout << vals[states - 1] << "|" << binary_string(states - 1, qubits) << ">\n\n";
    std::cout << std::setprecision(7) << std::defaultfloat << std::noshowpos;
}

void print_circuit_matrix(const QCircuit& circuit) {
    std::cout << std::setprecision(3) << std::fixed << std::showpos;
    af::print("Circuit: ", circuit.circuit());
    std::cout << std::setprecision(7) << std::defaultfloat << std::noshowpos;
}

void print_profile(const std::array<uint32_t, 2>& profile) {
    const int reps = static_cast<int>(profile[0] + profile[1]);
    std::cout << std::setprecision(3) << std::fixed << std::noshowpos;
    std::cout << "|0>: " << profile[0] * 100.f / reps << "% (" << profile[0] << ")\n"
              << "|1>: " << profile[1] * 100.f / reps << "% (" << profile[1] << ")\n";
    std::cout << std::setprecision(7) << std::defaultfloat << std::noshowpos;
}

void print_profile(const std::vector<uint32_t>& profile) {
    const int qubits = static_cast<int>(fast_log2(static_cast<uint32_t>(profile.size())));
    const long reps  = std::accumulate(profile.begin(), profile.end(), 0L);
    std::cout << std::setprecision(2) << std::fixed;
    for (std::size_t i = 0; i < profile.size(); i++)
        std::cout << "|" << binary_string(static_cast<uint32_t>(i), qubits) << ">: " << std::setw(5)
                  << profile[i] * 100.f / reps << "% (" << profile[i] << ")\n";
    std::cout << std::setprecision(7) << std::defaultfloat;
}

// ---------------------------------------------------------------------------
// text renderer
// ---------------------------------------------------------------------------
namespace {

struct Wire {
    std::string top, mid, bot;   // three text rows per qubit
    std::size_t cols = 0;        // display width appended so far (all three rows equal)
    void add(const std::string& t, const std::string& m, const std::string& b, std::size_t width) {
        top += t; mid += m; bot += b; cols += width;
    }
    void pad_to(std::size_t width) {
        if (cols < width) add(repeat(width - cols, " "), repeat(width - cols, "─"), repeat(width - cols, " "), width - cols);
    }
};

enum class Role { Pass, Control, Target };

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::size_t b = 0;
    for (std::size_t e = s.find(sep, b); e != std::string::npos; b = e + 1, e = s.find(sep, b)) out.push_back(s.substr(b, e - b));
    out.push_back(s.substr(b));
    return out;
}

uint32_t to_u32(const std::string& s) {
    if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos)
        throw std::invalid_argument{"Invalid circuit schematic: expected a number, got '" + s + "'"};
    return static_cast<uint32_t>(std::stoul(s));
}

// spaces/line (w) with `mark` at column c
std::string with_mark(std::size_t w, std::size_t c, const char* fill, const char* mark) {
    return repeat(c, fill) + mark + repeat(w - c - 1, fill);
}

void draw_gate(std::vector<Wire>& wires, const std::string& name, std::vector<uint32_t> controls,
               std::vector<uint32_t> targets) {
    std::sort(controls.begin(), controls.end());
    std::sort(targets.begin(), targets.end());
    const uint32_t lo = std::min(controls.empty() ? targets.front() : controls.front(), targets.front());
    const uint32_t hi = std::max(controls.empty() ? targets.back() : controls.back(), targets.back());
    const bool ctrl_above = !controls.empty() && controls.front() < targets.front();
    const bool ctrl_below = !controls.empty() && controls.back() > targets.back();

    // all wires of the span start the gate in the same column
    std::size_t start = 0;
    for (uint32_t q = lo; q <= hi; ++q) start = std::max(start, wires[q].cols);
    for (uint32_t q = lo; q <= hi; ++q) wires[q].pad_to(start);

    const bool is_swap      = (name == "Swap");
    const std::string label = " " + name + " ";
    const std::size_t L     = utf8str_len(label);
    const std::size_t half  = (L + 1) / 2 - 1;           // connector column inside the box
    const std::size_t width = is_swap ? 3 : L + 6;
    const std::size_t conn  = is_swap ? 1 : half + 3;    // connector column inside the cell

    std::vector<Role> role(hi - lo + 1, Role::Pass);
    for (uint32_t c : controls) role[c - lo] = Role::Control;
    for (uint32_t t : targets) role[t - lo] = Role::Target;

    for (uint32_t q = lo; q <= hi; ++q) {
        Wire& w = wires[q];
        switch (role[q - lo]) {
            case Role::Pass:
                w.add(with_mark(width, conn, " ", "│"), with_mark(width, conn, "─", "┼"), with_mark(width, conn, " ", "│"), width);
                break;
            case Role::Control: {
                const bool first = ctrl_above && q == controls.front();
                const bool last  = ctrl_below && q == controls.back();
                const std::string link = with_mark(width, conn, " ", "│"), blank = repeat(width, " ");
                w.add(first ? blank : link, with_mark(width, conn, "─", "█"), (!first && last) ? blank : link, width);
                break;
            }
            case Role::Target:
                if (is_swap) {
                    const bool upper = (q == targets.front());
                    w.add(upper ? (ctrl_above ? " │ " : "   ") : " │ ", "─╳─", upper ? " │ " : (ctrl_below ? " │ " : "   "), width);
                } else {
                    // consecutive targets share one tall box; the label sits on its first wire
                    const bool run_first = (q == lo) || role[q - 1 - lo] != Role::Target;
                    const bool run_last  = (q == hi) || role[q + 1 - lo] != Role::Target;
                    const std::string side = "  │" + repeat(L, " ") + "│  ";
                    std::string top = side, bot = side;
                    if (run_first)
                        top = (!ctrl_above && q == targets.front())
                                  ? "  ┌" + repeat(L, "─") + "┐  "
                                  : "  ┌" + repeat(half, "─") + "┴" + repeat(L - half - 1, "─") + "┐  ";
                    if (run_last)
                        bot = (!ctrl_below && q == targets.back())
                                  ? "  └" + repeat(L, "─") + "┘  "
                                  : "  └" + repeat(half, "─") + "┬" + repeat(L - half - 1, "─") + "┘  ";
                    w.add(top, "──┤" + (run_first ? label : repeat(L, " ")) + "├──", bot, width);
                }
                break;
        }
    }
}

std::string render(std::string text) {
    text.erase(std::remove_if(text.begin(), text.end(), [](unsigned char c) { return std::isspace(c); }), text.end());
    std::vector<std::string> stmts = split(text, ';');
    if (stmts.empty() || !stmts.back().empty())
        throw std::invalid_argument{"Invalid circuit schematic: every statement must end with ';'"};
    stmts.pop_back();
    if (stmts.empty()) throw std::invalid_argument{"Invalid circuit schematic: missing qubit count"};

    const uint32_t n = to_u32(stmts[0]);
    if (n == 0) throw std::out_of_range{"Circuit must contain at least 1 qubit"};
    if (n > max_qubit_count) throw std::out_of_range{"Maximum qubit count supported is " + std::to_string(max_qubit_count)};

    // initial states: "i,v"
    std::vector<int> init(n, -1);
    std::size_t pos = 1;
    for (; pos < stmts.size(); ++pos) {
        const auto parts = split(stmts[pos], ',');
        if (parts.size() != 2 || stmts[pos].find(':') != std::string::npos) break;
        if (parts[0].find_first_not_of("0123456789") != std::string::npos) break;
        const uint32_t q = to_u32(parts[0]), v = to_u32(parts[1]);
        if (q >= n) throw std::out_of_range{"Qubit index out of the circuit range"};
        if (v > 1) throw std::invalid_argument{"Initial qubit state must be 0 or 1"};
        if (init[q] != -1) throw std::invalid_argument{"Qubit initial state declared more than once"};
        init[q] = static_cast<int>(v);
    }

    std::vector<Wire> wires(n);
    for (uint32_t q = 0; q < n; ++q) {
        wires[q].top = "   ";
        wires[q].mid = init[q] == 1 ? "|1⟩" : "|0⟩";
        wires[q].bot = "   ";
        wires[q].cols = 3;
    }
    auto longest = [&]() {
        std::size_t m = 0;
        for (const auto& w : wires) m = std::max(m, w.cols);
        return m;
    };

    for (; pos < stmts.size(); ++pos) {
        const std::string& st = stmts[pos];
        if (st == "B" || st == "P") {
            const std::size_t m = longest();
            for (auto& w : wires) {
                w.pad_to(m);
                if (st == "B") w.add("▒ ", "▒─", "▒ ", 2);
            }
            continue;
        }
        const std::size_t colon = st.find(':');
        if (colon == std::string::npos) throw std::invalid_argument{"Invalid circuit schematic near '" + st + "'"};
        const auto head = split(st.substr(0, colon), ',');
        if (head.size() != 3 || head[0].empty()) throw std::invalid_argument{"Invalid circuit schematic near '" + st + "'"};
        const uint32_t nc = to_u32(head[1]), nt = to_u32(head[2]);
        std::vector<uint32_t> qs;
        for (const auto& tok : split(st.substr(colon + 1), ',')) qs.push_back(to_u32(tok));
        for (uint32_t q : qs)
            if (q >= n) throw std::out_of_range{"Gate qubit index out of the circuit range"};
        {
            std::vector<uint32_t> sorted = qs;
            std::sort(sorted.begin(), sorted.end());
            if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end())
                throw std::invalid_argument{"A gate cannot use the same qubit twice"};
        }
        if (qs.size() != static_cast<std::size_t>(nc) + nt)
            throw std::invalid_argument{"Number of qubits does not match the declared control and target counts"};
        if (nt == 0) throw std::invalid_argument{"A gate needs at least one target qubit"};
        if (head[0] == "Swap" && nt != 2) throw std::invalid_argument{"Swap gates take exactly two target qubits"};
        draw_gate(wires, head[0], std::vector<uint32_t>(qs.begin(), qs.begin() + nc),
                  std::vector<uint32_t>(qs.begin() + nc, qs.end()));
    }

    const std::size_t m = longest();
    std::string out = "\n";
    for (auto& w : wires) {
        w.pad_to(m);
        out += w.top + "\n" + w.mid + "\n" + w.bot + "\n";
    }
    return out;
}

}  // namespace

std::string gen_circuit_text_image(const QCircuit& circuit, const QSimulator& simulator) {
    if (circuit.qubit_count() != simulator.qubit_count())
        throw std::invalid_argument{"Circuit and simulator qubit count must match"};
    const uint32_t qubits = circuit.qubit_count();
    std::stringstream text;
    text << qubits << ";";
    for (uint32_t i = 0; i < qubits; ++i) {
        const auto& q = simulator.qubit(i);
        int val;
        if (q == aqs::QState::zero()) val = 0;
        else if (q == aqs::QState::one()) val = 1;
        else throw std::invalid_argument{"Superposed inital states not supported"};
        text << i << "," << val << ";";
    }
    text << circuit.representation();
    return render(text.str());
}

std::string gen_circuit_text_image(std::string schematic) { return render(std::move(schematic)); }

void print_circuit_text_image(const QCircuit& circuit, const QSimulator& simulator) {
    std::cout << gen_circuit_text_image(circuit, simulator) << std::endl;
}

}  // namespace aqs

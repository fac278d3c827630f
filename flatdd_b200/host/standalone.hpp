// standalone.hpp — a host front end that needs nothing of the reference at build or run time
// (SURVEY.md section 8f, rows N1 and N2): an OpenQASM 2 reader, the gate matrices, a dense-block
// fusion pass and a builder that emits every (fused) gate as a flat matrix DD for the C-ABI.
//
// Policy ("start flat when the state fits HBM", N2): there is no vector-DD phase.  The state
// starts as the flat array of |0...0> on the device (the conversion kernel expands the
// n-node DD of that state) and every gate is a DMAVM launch.  What the reference's DD phase
// contributes — tolerance snapping of amplitudes below ~1e-13 (include/dd/RealNumber.hpp) —
// is absent, so final states agree with the reference within its own tolerance, not bit for bit.
//
// Fusion (N1): operations are multiplied into dense blocks with pairwise disjoint qubit sets (blocks
// on disjoint qubits commute, so several stay open at once) while a block
// stays on at most `maxBlockQubits` qubits of which at most `maxNonDiagonal` are non-diagonal
// (the quantity that sizes a tile of the DMAVM kernel: a block that is complete on <= 4 upper
// qubits runs on the tensor-core path at ~1.3 HBM passes whatever it holds).  Diagonal qubits
// (controls, phase gates) ride along for free.  This replaces the reference's DD-multiply based
// greedy pass (src/SwitchSimulator.cpp:267-340), which needs the DD package.
//
// Gate matrices follow the reference (include/dd/GateMatrixDefinitions.hpp:29-135) including
// global phases; two-target gates take the first listed qubit as the high bit of the 4x4 matrix.
#pragma once

#include "array_backend.hpp"
#include "block_fusion.hpp"

#include <algorithm>
#include <array>
#include <cctype>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace fddb200::standalone {

using cplx = std::complex<double>;

// ------------------------------------------------------------------------------------------------
// circuit IR + OpenQASM 2 reader
// ------------------------------------------------------------------------------------------------
struct Op {
    enum Kind { Unitary, Measure, Barrier, Reset } kind = Unitary;
    std::string name;          // gate name as written
    std::vector<int> qubits;   // operands in the order written (controls first for c* gates)
    std::vector<double> params;
};

struct Circuit {
    std::string name;
    int nQubits = 0;
    int nClbits = 0;
    std::vector<Op> ops; // one entry per statement like qc::QuantumComputation::getNops(); a broadcast gate is one entry per element
    [[nodiscard]] std::size_t nOps() const { return ops.size(); }
};

class QasmError : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

// true when the name is one of the built-in gates (with any number of leading 'c' controls)
inline bool isBuiltinGate(const std::string& name);
inline bool isStandardGate(const std::string& name);

namespace detail {

// arithmetic of gate parameters: + - * / ^, unary minus, parentheses, pi, sin cos tan exp ln sqrt
class Expr {
public:
    explicit Expr(const std::string& s, const std::map<std::string, double>* vars = nullptr) : s_(s), vars_(vars) {}
    double parse() {
        const double v = sum();
        skip();
        if (pos_ != s_.size()) throw QasmError("cannot parse parameter '" + s_ + "'");
        return v;
    }

private:
    void skip() {
        while (pos_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[pos_]))) ++pos_;
    }
    bool eat(char c) {
        skip();
        if (pos_ < s_.size() && s_[pos_] == c) {
            ++pos_;
            return true;
        }
        return false;
    }
    double sum() {
        double v = product();
        for (;;) {
            if (eat('+')) v += product();
            else if (eat('-')) v -= product();
            else return v;
        }
    }
    double product() {
        double v = unary();
        for (;;) {
            if (eat('*')) v *= unary();
            else if (eat('/')) v /= unary();
            else return v;
        }
    }
    double unary() {
        if (eat('-')) return -unary();
        if (eat('+')) return unary();
        const double base = primary();
        if (eat('^')) return std::pow(base, unary());
        return base;
    }
    double primary() {
        skip();
        if (eat('(')) {
            const double v = sum();
            if (!eat(')')) throw QasmError("missing ')' in parameter '" + s_ + "'");
            return v;
        }
        if (pos_ < s_.size() && (std::isdigit(static_cast<unsigned char>(s_[pos_])) || s_[pos_] == '.')) {
            std::size_t used = 0;
            const double v = std::stod(s_.substr(pos_), &used);
            pos_ += used;
            return v;
        }
        std::string id;
        while (pos_ < s_.size() && (std::isalnum(static_cast<unsigned char>(s_[pos_])) || s_[pos_] == '_')) id += s_[pos_++];
        if (id == "pi") return M_PI;
        if (id.empty()) throw QasmError("cannot parse parameter '" + s_ + "'");
        if (vars_ != nullptr) { // formal parameter of a gate definition
            const auto it = vars_->find(id);
            if (it != vars_->end()) return it->second;
        }
        if (!eat('(')) throw QasmError("unknown identifier '" + id + "' in parameter");
        const double a = sum();
        if (!eat(')')) throw QasmError("missing ')' in parameter '" + s_ + "'");
        if (id == "sin") return std::sin(a);
        if (id == "cos") return std::cos(a);
        if (id == "tan") return std::tan(a);
        if (id == "exp") return std::exp(a);
        if (id == "ln") return std::log(a);
        if (id == "sqrt") return std::sqrt(a);
        throw QasmError("unknown function '" + id + "'");
    }
    const std::string& s_;
    const std::map<std::string, double>* vars_;
    std::size_t pos_ = 0;
};

inline std::string trim(const std::string& s) {
    std::size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}

inline std::vector<std::string> splitTop(const std::string& s, char sep) { // split outside parentheses
    std::vector<std::string> out;
    int depth = 0;
    std::string cur;
    for (char c : s) {
        if (c == '(') ++depth;
        if (c == ')') --depth;
        if (c == sep && depth == 0) {
            out.push_back(trim(cur));
            cur.clear();
        } else {
            cur += c;
        }
    }
    if (!trim(cur).empty()) out.push_back(trim(cur));
    return out;
}

} // namespace detail

inline Circuit parseQasm(std::istream& in, const std::string& name) {
    Circuit c;
    c.name = name;
    std::string text, line;
    while (std::getline(in, line)) {
        const auto cut = line.find("//");
        text += (cut == std::string::npos ? line : line.substr(0, cut));
        text += '\n';
    }
    // gate definitions: `gate name(params) qargs { body }` are cut out of the text and expanded at every use (macro
    // semantics of OpenQASM 2); anything else with a block (if / opaque bodies) is rejected
    struct GateDef {
        std::vector<std::string> params, qargs, body;
    };
    std::map<std::string, GateDef> defs;
    for (;;) {
        const auto open = text.find('{');
        if (open == std::string::npos) break;
        const auto close = text.find('}', open);
        if (close == std::string::npos) throw QasmError("missing '}'");
        auto start = text.rfind(';', open);
        start = start == std::string::npos ? 0 : start + 1;
        std::string head = detail::trim(text.substr(start, open - start));
        if (head.rfind("gate", 0) != 0 || head.size() < 5 || !std::isspace(static_cast<unsigned char>(head[4]))) {
            throw QasmError("only gate definitions may have a block: '" + head + "'");
        }
        head = detail::trim(head.substr(4));
        GateDef def;
        std::size_t p = 0;
        while (p < head.size() && (std::isalnum(static_cast<unsigned char>(head[p])) || head[p] == '_')) ++p;
        const std::string gname = head.substr(0, p);
        std::string rest = detail::trim(head.substr(p));
        if (!rest.empty() && rest[0] == '(') {
            const auto rp = rest.find(')');
            if (rp == std::string::npos) throw QasmError("missing ')' in gate definition " + gname);
            def.params = detail::splitTop(rest.substr(1, rp - 1), ',');
            rest = detail::trim(rest.substr(rp + 1));
        }
        def.qargs = detail::splitTop(rest, ',');
        def.body = detail::splitTop(text.substr(open + 1, close - open - 1), ';');
        if (gname.empty() || def.qargs.empty()) throw QasmError("malformed gate definition '" + head + "'");
        defs[gname] = def;
        text.erase(start, close + 1 - start);
    }
    struct Reg {
        int first, size;
    };
    std::map<std::string, Reg> qregs, cregs;
    // operand "q[3]" -> {3 + first}; "q" -> the whole register
    auto operand = [&](const std::string& tok, const std::map<std::string, Reg>& regs) {
        std::vector<int> out;
        const auto lb = tok.find('[');
        const std::string reg = detail::trim(tok.substr(0, lb));
        const auto it = regs.find(reg);
        if (it == regs.end()) throw QasmError("unknown register '" + reg + "'");
        if (lb == std::string::npos) {
            for (int i = 0; i < it->second.size; ++i) out.push_back(it->second.first + i);
        } else {
            const int idx = std::stoi(tok.substr(lb + 1));
            if (idx < 0 || idx >= it->second.size) throw QasmError("index out of range in '" + tok + "'");
            out.push_back(it->second.first + idx);
        }
        return out;
    };
    // a use of a defined gate is replaced by its body with parameters and qubits substituted, recursively
    std::function<void(const Op&, int)> expand = [&](const Op& use, int depth) {
        const auto it = defs.find(use.name);
        // a definition in the file wins unless the name is a base gate or one of qelib1's own controlled gates
        // (whose bodies the built-in matrices reproduce); a user gate that merely looks like "c" + gate keeps its body
        if (it == defs.end() || isStandardGate(use.name)) {
            c.ops.push_back(use);
            return;
        }
        const GateDef& def = it->second;
        if (depth > 64) throw QasmError("gate definitions nest too deeply (recursive definition of '" + use.name + "'?)");
        if (use.params.size() != def.params.size() || use.qubits.size() != def.qargs.size()) {
            throw QasmError("gate " + use.name + " used with the wrong number of parameters or qubits");
        }
        std::map<std::string, double> vars;
        for (std::size_t i = 0; i < def.params.size(); ++i) vars[def.params[i]] = use.params[i];
        for (const std::string& rawBody : def.body) {
            const std::string st = detail::trim(rawBody);
            if (st.empty()) continue;
            std::size_t p = 0;
            while (p < st.size() && (std::isalnum(static_cast<unsigned char>(st[p])) || st[p] == '_')) ++p;
            Op inner;
            inner.name = st.substr(0, p);
            if (inner.name == "barrier") continue;
            std::string rest = detail::trim(st.substr(p));
            if (!rest.empty() && rest[0] == '(') {
                int level = 0;
                std::size_t close = 0;
                for (; close < rest.size(); ++close) {
                    if (rest[close] == '(') ++level;
                    if (rest[close] == ')' && --level == 0) break;
                }
                if (close == rest.size()) throw QasmError("missing ')' in the body of gate " + use.name);
                for (const auto& e : detail::splitTop(rest.substr(1, close - 1), ',')) inner.params.push_back(detail::Expr(e, &vars).parse());
                rest = detail::trim(rest.substr(close + 1));
            }
            for (const auto& formal : detail::splitTop(rest, ',')) {
                const auto at = std::find(def.qargs.begin(), def.qargs.end(), formal);
                if (at == def.qargs.end()) throw QasmError("unknown qubit '" + formal + "' in the body of gate " + use.name);
                inner.qubits.push_back(use.qubits[static_cast<std::size_t>(at - def.qargs.begin())]);
            }
            expand(inner, depth + 1);
        }
    };
    for (const std::string& raw : detail::splitTop(text, ';')) {
        const std::string st = detail::trim(raw);
        if (st.empty()) continue;
        std::size_t p = 0;
        while (p < st.size() && (std::isalnum(static_cast<unsigned char>(st[p])) || st[p] == '_')) ++p;
        const std::string head = st.substr(0, p);
        std::string rest = detail::trim(st.substr(p));
        if (head == "OPENQASM" || head == "include") continue;
        if (head == "qreg" || head == "creg") {
            const auto lb = rest.find('['), rb = rest.find(']');
            if (lb == std::string::npos || rb == std::string::npos) throw QasmError("bad register declaration '" + st + "'");
            const std::string reg = detail::trim(rest.substr(0, lb));
            const int size = std::stoi(rest.substr(lb + 1, rb - lb - 1));
            if (head == "qreg") {
                qregs[reg] = {c.nQubits, size};
                c.nQubits += size;
            } else {
                cregs[reg] = {c.nClbits, size};
                c.nClbits += size;
            }
            continue;
        }
        Op op;
        op.name = head;
        if (head == "measure") {
            op.kind = Op::Measure;
            const auto arrow = rest.find("->");
            if (arrow == std::string::npos) throw QasmError("measure without '->'");
            op.qubits = operand(detail::trim(rest.substr(0, arrow)), qregs);
            c.ops.push_back(op);
            continue;
        }
        if (head == "barrier" || head == "reset") {
            op.kind = head == "barrier" ? Op::Barrier : Op::Reset;
            for (const auto& tok : detail::splitTop(rest, ',')) {
                for (int q : operand(tok, qregs)) op.qubits.push_back(q);
            }
            c.ops.push_back(op);
            continue;
        }
        if (!rest.empty() && rest[0] == '(') {
            int depth = 0;
            std::size_t close = 0;
            for (; close < rest.size(); ++close) {
                if (rest[close] == '(') ++depth;
                if (rest[close] == ')' && --depth == 0) break;
            }
            if (close == rest.size()) throw QasmError("missing ')' in '" + st + "'");
            for (const auto& e : detail::splitTop(rest.substr(1, close - 1), ',')) op.params.push_back(detail::Expr(e).parse());
            rest = detail::trim(rest.substr(close + 1));
        }
        // operands; a whole-register operand broadcasts the gate (one Op per element, one statement)
        std::vector<std::vector<int>> args;
        std::size_t broadcast = 1;
        for (const auto& tok : detail::splitTop(rest, ',')) {
            args.push_back(operand(tok, qregs));
            if (args.back().size() > 1) {
                if (broadcast > 1 && args.back().size() != broadcast) throw QasmError("registers of different sizes in '" + st + "'");
                broadcast = args.back().size();
            }
        }
        if (args.empty()) throw QasmError("gate without operands: '" + st + "'");
        for (std::size_t b = 0; b < broadcast; ++b) {
            Op one = op;
            for (const auto& a : args) one.qubits.push_back(a.size() > 1 ? a[b] : a[0]);
            expand(one, 0);
        }
    }
    if (c.nQubits == 0) throw QasmError("no qreg declared");
    return c;
}

inline Circuit parseQasmFile(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw QasmError("cannot open " + path);
    std::string name = path;
    const auto slash = name.find_last_of('/');
    if (slash != std::string::npos) name = name.substr(slash + 1);
    const auto dot = name.find_last_of('.');
    if (dot != std::string::npos) name = name.substr(0, dot);
    return parseQasm(in, name);
}

// ------------------------------------------------------------------------------------------------
// gate matrices (reference include/dd/GateMatrixDefinitions.hpp)
// ------------------------------------------------------------------------------------------------
// A dense block on a sorted set of qubits: bit i of a row / column index belongs to qubits[i].
struct Block {
    std::vector<int> qubits;
    std::vector<cplx> m; // row-major, dim x dim, dim = 2^qubits.size()
    [[nodiscard]] std::size_t dim() const { return std::size_t{1} << qubits.size(); }
};

namespace detail {

using M2 = std::array<cplx, 4>;
using M4 = std::array<cplx, 16>;

inline M2 u3(double theta, double phi, double lambda) {
    return {cplx(std::cos(theta / 2), 0.0), cplx(-std::cos(lambda) * std::sin(theta / 2), -std::sin(lambda) * std::sin(theta / 2)),
            cplx(std::cos(phi) * std::sin(theta / 2), std::sin(phi) * std::sin(theta / 2)),
            cplx(std::cos(lambda + phi) * std::cos(theta / 2), std::sin(lambda + phi) * std::cos(theta / 2))};
}

inline bool oneQubitMatrix(const std::string& g, const std::vector<double>& p, M2& out) {
    const double s = M_SQRT1_2;
    const cplx i(0, 1);
    auto need = [&](std::size_t n) {
        if (p.size() != n) throw QasmError("gate " + g + " takes " + std::to_string(n) + " parameter(s)");
    };
    if (g == "id" || g == "i") out = {1, 0, 0, 1};
    else if (g == "x") out = {0, 1, 1, 0};
    else if (g == "y") out = {0, -i, i, 0};
    else if (g == "z") out = {1, 0, 0, -1};
    else if (g == "h") out = {s, s, s, -s};
    else if (g == "s") out = {1, 0, 0, i};
    else if (g == "sdg") out = {1, 0, 0, -i};
    else if (g == "t") out = {1, 0, 0, cplx(s, s)};
    else if (g == "tdg") out = {1, 0, 0, cplx(s, -s)};
    else if (g == "sx") out = {cplx(0.5, 0.5), cplx(0.5, -0.5), cplx(0.5, -0.5), cplx(0.5, 0.5)};
    else if (g == "sxdg") out = {cplx(0.5, -0.5), cplx(0.5, 0.5), cplx(0.5, 0.5), cplx(0.5, -0.5)};
    else if (g == "rx") { need(1); out = {cplx(std::cos(p[0] / 2), 0), cplx(0, -std::sin(p[0] / 2)), cplx(0, -std::sin(p[0] / 2)), cplx(std::cos(p[0] / 2), 0)}; }
    else if (g == "ry") { need(1); out = {std::cos(p[0] / 2), -std::sin(p[0] / 2), std::sin(p[0] / 2), std::cos(p[0] / 2)}; }
    else if (g == "rz") { need(1); out = {cplx(std::cos(p[0] / 2), -std::sin(p[0] / 2)), 0, 0, cplx(std::cos(p[0] / 2), std::sin(p[0] / 2))}; }
    else if (g == "p" || g == "u1" || g == "phase") { need(1); out = {1, 0, 0, cplx(std::cos(p[0]), std::sin(p[0]))}; }
    else if (g == "u2") { need(2); out = u3(M_PI / 2, p[0], p[1]); out[0] = s; }
    else if (g == "u3" || g == "u" || g == "U") { need(3); out = u3(p[0], p[1], p[2]); }
    else return false;
    return true;
}

// 4x4 matrices: index = 2 * (bit of the first listed qubit) + (bit of the second)
inline bool twoQubitMatrix(const std::string& g, const std::vector<double>& p, M4& out) {
    const cplx i(0, 1);
    out.fill(0);
    auto rot = [&](double& c, double& s) {
        if (p.size() != 1) throw QasmError("gate " + g + " takes 1 parameter");
        c = std::cos(p[0] / 2);
        s = std::sin(p[0] / 2);
    };
    double c = 0, s = 0;
    if (g == "swap") { out[0] = out[6] = out[9] = out[15] = 1; }
    else if (g == "iswap") { out[0] = out[15] = 1; out[6] = out[9] = i; }
    else if (g == "dcx") { out[0] = 1; out[7] = 1; out[9] = 1; out[14] = 1; }
    else if (g == "rzz") { rot(c, s); out[0] = out[15] = cplx(c, -s); out[5] = out[10] = cplx(c, s); }
    else if (g == "rxx") { rot(c, s); out[0] = out[5] = out[10] = out[15] = c; out[3] = out[6] = out[9] = out[12] = cplx(0, -s); }
    else if (g == "ryy") { rot(c, s); out[0] = out[5] = out[10] = out[15] = c; out[3] = out[12] = cplx(0, s); out[6] = out[9] = cplx(0, -s); }
    else if (g == "rzx") { rot(c, s); out[0] = out[5] = out[10] = out[15] = c; out[1] = out[4] = cplx(0, -s); out[11] = out[14] = cplx(0, s); }
    else return false;
    return true;
}

} // namespace detail

inline bool isBuiltinGate(const std::string& name) {
    static const char* const one[] = {"id", "i", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg", "rx", "ry", "rz", "p", "u1", "phase", "u2", "u3", "u", "U"};
    static const char* const two[] = {"swap", "iswap", "dcx", "rzz", "rxx", "ryy", "rzx"};
    std::string base = name;
    if (base == "CX" || base == "cnot") base = "cx";
    if (base == "toffoli") base = "ccx";
    if (base == "fredkin") base = "cswap";
    for (;;) {
        for (const char* g : one) {
            if (base == g) return true;
        }
        for (const char* g : two) {
            if (base == g) return true;
        }
        if (base.size() > 1 && base[0] == 'c') {
            base = base.substr(1);
            continue;
        }
        return false;
    }
}

// Names whose meaning is fixed by the language or by qelib1.inc: the base gates and qelib1's controlled gates.
inline bool isStandardGate(const std::string& name) {
    static const char* const controlled[] = {"CX", "cnot", "toffoli", "fredkin", "cx", "cy", "cz", "ch", "ccx", "cswap", "crx", "cry", "crz",
                                             "cu1", "cp", "cu3", "csx", "c3x", "c4x"};
    for (const char* g : controlled) {
        if (name == g) return true;
    }
    return isBuiltinGate(name) && !(name.size() > 1 && name[0] == 'c' && isBuiltinGate(name.substr(1)));
}

// The operation as a dense block on its own qubits.  Leading 'c's of the name are controls
// (cx, ccx, cswap, cu3, ...), the rest is a one- or two-target base gate.
inline Block blockOf(const Op& op) {
    std::string base = op.name;
    if (base == "CX") base = "cx";
    if (base == "cnot") base = "cx";
    if (base == "toffoli") base = "ccx";
    if (base == "fredkin") base = "cswap";
    detail::M2 m2{};
    detail::M4 m4{};
    int nControls = 0;
    int nTargets = 0;
    for (;;) {
        if (detail::oneQubitMatrix(base, op.params, m2)) {
            nTargets = 1;
            break;
        }
        if (detail::twoQubitMatrix(base, op.params, m4)) {
            nTargets = 2;
            break;
        }
        if (base.size() > 1 && base[0] == 'c') {
            base = base.substr(1);
            ++nControls;
            continue;
        }
        throw QasmError("unsupported gate '" + op.name + "'");
    }
    if (static_cast<int>(op.qubits.size()) != nControls + nTargets) {
        throw QasmError("gate " + op.name + " takes " + std::to_string(nControls + nTargets) + " qubit(s)");
    }
    Block b;
    b.qubits = op.qubits;
    std::sort(b.qubits.begin(), b.qubits.end());
    if (std::adjacent_find(b.qubits.begin(), b.qubits.end()) != b.qubits.end()) throw QasmError("gate " + op.name + " repeats a qubit");
    auto pos = [&](int q) { return static_cast<int>(std::lower_bound(b.qubits.begin(), b.qubits.end(), q) - b.qubits.begin()); };
    const std::size_t dim = b.dim();
    b.m.assign(dim * dim, cplx(0, 0));
    std::size_t controlMask = 0;
    for (int k = 0; k < nControls; ++k) controlMask |= std::size_t{1} << pos(op.qubits[static_cast<std::size_t>(k)]);
    const int t0 = pos(op.qubits[static_cast<std::size_t>(nControls)]);
    const int t1 = nTargets == 2 ? pos(op.qubits[static_cast<std::size_t>(nControls) + 1]) : -1;
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) {
            std::size_t targetMask = std::size_t{1} << t0;
            if (t1 >= 0) targetMask |= std::size_t{1} << t1;
            if (((r ^ c) & ~targetMask) != 0) continue; // controls (and nothing else) are untouched
            if ((r & controlMask) != controlMask) {
                if (r == c) b.m[r * dim + c] = 1.0;
                continue;
            }
            if (nTargets == 1) {
                b.m[r * dim + c] = m2[2 * ((r >> t0) & 1) + ((c >> t0) & 1)];
            } else {
                const std::size_t ri = 2 * ((r >> t0) & 1) + ((r >> t1) & 1);
                const std::size_t ci = 2 * ((c >> t0) & 1) + ((c >> t1) & 1);
                b.m[r * dim + c] = m4[4 * ri + ci];
            }
        }
    }
    return b;
}

// the block on a larger sorted qubit set (identity on the added qubits)
inline Block expand(const Block& b, const std::vector<int>& onto) {
    if (b.qubits == onto) return b;
    Block out;
    out.qubits = onto;
    const std::size_t dim = out.dim();
    out.m.assign(dim * dim, cplx(0, 0));
    std::vector<int> at; // position of each of b's qubits inside `onto`
    for (int q : b.qubits) at.push_back(static_cast<int>(std::lower_bound(onto.begin(), onto.end(), q) - onto.begin()));
    std::size_t own = 0;
    for (int p : at) own |= std::size_t{1} << p;
    auto gatherBits = [&](std::size_t x) {
        std::size_t v = 0;
        for (std::size_t i = 0; i < at.size(); ++i) v |= ((x >> at[i]) & 1) << i;
        return v;
    };
    const std::size_t bd = b.dim();
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) {
            if (((r ^ c) & ~own) != 0) continue;
            out.m[r * dim + c] = b.m[gatherBits(r) * bd + gatherBits(c)];
        }
    }
    return out;
}

// next * current (next acts after current)
inline Block multiply(const Block& next, const Block& current) {
    std::vector<int> all;
    std::set_union(next.qubits.begin(), next.qubits.end(), current.qubits.begin(), current.qubits.end(), std::back_inserter(all));
    const Block a = expand(next, all), b = expand(current, all);
    Block out;
    out.qubits = all;
    const std::size_t dim = out.dim();
    out.m.assign(dim * dim, cplx(0, 0));
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t k = 0; k < dim; ++k) {
            const cplx x = a.m[r * dim + k];
            if (x == cplx(0, 0)) continue;
            for (std::size_t c = 0; c < dim; ++c) out.m[r * dim + c] += x * b.m[k * dim + c];
        }
    }
    return out;
}

// qubits (positions in b.qubits) on which the block is not diagonal
inline int nonDiagonalCount(const Block& b) {
    const std::size_t dim = b.dim();
    std::size_t mask = 0;
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) {
            if (b.m[r * dim + c] != cplx(0, 0)) mask |= r ^ c;
        }
    }
    return __builtin_popcountll(mask);
}

// non-diagonal qubits of a block (as qubit numbers)
inline std::vector<int> nonDiagonalQubits(const Block& b) {
    const std::size_t dim = b.dim();
    std::size_t mask = 0;
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) {
            if (b.m[r * dim + c] != cplx(0, 0)) mask |= r ^ c;
        }
    }
    std::vector<int> out;
    for (std::size_t i = 0; i < b.qubits.size(); ++i) {
        if ((mask >> i) & 1) out.push_back(b.qubits[i]);
    }
    return out;
}

// the same block with qubit q renamed to map[q] (sharded states: logical -> physical positions)
inline Block relabel(const Block& b, const std::vector<int>& map) {
    const std::size_t k = b.qubits.size();
    std::vector<int> renamed(k);
    for (std::size_t i = 0; i < k; ++i) renamed[i] = map[static_cast<std::size_t>(b.qubits[i])];
    Block out;
    out.qubits = renamed;
    std::sort(out.qubits.begin(), out.qubits.end());
    std::vector<int> pos(k); // old bit i -> new bit pos[i]
    for (std::size_t i = 0; i < k; ++i) pos[i] = static_cast<int>(std::lower_bound(out.qubits.begin(), out.qubits.end(), renamed[i]) - out.qubits.begin());
    auto move = [&](std::size_t x) {
        std::size_t v = 0;
        for (std::size_t i = 0; i < k; ++i) v |= ((x >> i) & 1) << pos[i];
        return v;
    };
    const std::size_t dim = b.dim();
    out.m.assign(dim * dim, cplx(0, 0));
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) out.m[move(r) * dim + move(c)] = b.m[r * dim + c];
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// dense block -> flat matrix DD (full depth, identity levels explicit, normalised so that equal
// sub-blocks share nodes: a Kronecker product gets one node per level)
// ------------------------------------------------------------------------------------------------
// A dense block as the table the block fusion works on (block_fusion.hpp): targets = its non-diagonal qubits, context = the other
// qubits its matrix depends on (diagonally); a qubit on which it is the identity drops out.
inline SmallGate smallGateOf(const Block& b) {
    const std::size_t dim = b.dim();
    std::size_t nd = 0;
    for (std::size_t r = 0; r < dim; ++r) {
        for (std::size_t c = 0; c < dim; ++c) {
            if (b.m[r * dim + c] != cplx(0, 0)) nd |= r ^ c;
        }
    }
    // a diagonal qubit matters when flipping it changes some entry
    std::size_t dep = 0;
    for (std::size_t i = 0; i < b.qubits.size(); ++i) {
        if ((nd >> i) & 1U) continue;
        const std::size_t bit = std::size_t{1} << i;
        bool differs = false;
        for (std::size_t r = 0; r < dim && !differs; ++r) {
            if (r & bit) continue;
            for (std::size_t c = 0; c < dim; ++c) {
                if (c & bit) continue;
                if (b.m[r * dim + c] != b.m[(r | bit) * dim + (c | bit)]) {
                    differs = true;
                    break;
                }
            }
        }
        if (differs) dep |= bit;
    }
    SmallGate g;
    std::vector<int> tPos, cPos;
    for (std::size_t i = 0; i < b.qubits.size(); ++i) {
        if ((nd >> i) & 1U) {
            g.targets.push_back(b.qubits[i]);
            tPos.push_back(static_cast<int>(i));
        } else if ((dep >> i) & 1U) {
            g.ctx.push_back(b.qubits[i]);
            cPos.push_back(static_cast<int>(i));
        }
    }
    const std::size_t rows = std::size_t{1} << tPos.size(), nc = std::size_t{1} << cPos.size();
    g.table.assign(nc * rows * rows, cplx(0, 0));
    auto spread = [](std::size_t x, const std::vector<int>& pos) {
        std::size_t v = 0;
        for (std::size_t i = 0; i < pos.size(); ++i) v |= ((x >> i) & 1U) << pos[i];
        return v;
    };
    for (std::size_t cx = 0; cx < nc; ++cx) {
        const std::size_t base = spread(cx, cPos); // identity qubits at 0: the entries do not depend on them
        for (std::size_t r = 0; r < rows; ++r) {
            for (std::size_t c = 0; c < rows; ++c) g.table[(cx * rows + r) * rows + c] = b.m[(base | spread(r, tPos)) * dim + (base | spread(c, tPos))];
        }
    }
    return g;
}

class GateDDBuilder {
public:
    explicit GateDDBuilder(int nQubits) : n_(nQubits) {}

    FlatMatDD build(const Block& b) {
        out_ = FlatMatDD{};
        out_.n_qubits = n_;
        unique_.clear();
        memo_.clear();
        ident_.assign(static_cast<std::size_t>(n_) + 1, FDD_TERMINAL - 1);
        block_ = &b;
        posOf_.assign(static_cast<std::size_t>(n_), -1);
        for (std::size_t i = 0; i < b.qubits.size(); ++i) {
            if (b.qubits[i] < 0 || b.qubits[i] >= n_) throw std::runtime_error("block qubit out of range");
            posOf_[static_cast<std::size_t>(b.qubits[i])] = static_cast<int>(i);
        }
        const Edge root = make(n_ - 1, 0, 0);
        if (root.w == cplx(0, 0)) throw std::runtime_error("zero matrix");
        out_.root = root.node;
        out_.root_weight[0] = root.w.real();
        out_.root_weight[1] = root.w.imag();
        return std::move(out_);
    }

private:
    struct Edge {
        int32_t node = FDD_TERMINAL;
        cplx w{0, 0};
    };
    static int64_t grid(double x) { return static_cast<int64_t>(std::llround(x * 70368744177664.0)); } // 2^-46 steps
    int32_t identChain(int lv) {
        if (lv < 0) return FDD_TERMINAL;
        int32_t& slot = ident_[static_cast<std::size_t>(lv)];
        if (slot == FDD_TERMINAL - 1) {
            const int32_t below = identChain(lv - 1);
            slot = addNode(lv, {Edge{below, 1.0}, Edge{}, Edge{}, Edge{below, 1.0}}, lv == 0);
        }
        return slot;
    }
    int32_t addNode(int lv, const std::array<Edge, 4>& e, bool leafLevel) {
        std::array<int64_t, 13> key{};
        key[0] = lv;
        for (int k = 0; k < 4; ++k) {
            key[1 + 3 * k] = e[k].w == cplx(0, 0) ? -7 : e[k].node;
            key[2 + 3 * k] = grid(e[k].w.real());
            key[3 + 3 * k] = grid(e[k].w.imag());
        }
        const auto it = unique_.find(key);
        if (it != unique_.end()) return it->second;
        const auto id = static_cast<int32_t>(out_.level.size());
        out_.level.push_back(lv);
        for (int k = 0; k < 4; ++k) {
            const bool zero = e[k].w == cplx(0, 0);
            out_.child.push_back(zero || leafLevel ? FDD_TERMINAL : e[k].node);
            out_.weight.push_back(zero ? 0.0 : e[k].w.real());
            out_.weight.push_back(zero ? 0.0 : e[k].w.imag());
        }
        unique_.emplace(key, id);
        return id;
    }
    // sub-matrix with the block bits of all levels above `lv` fixed to (r, c)
    Edge make(int lv, std::size_t r, std::size_t c) {
        const Block& b = *block_;
        const int lowest = b.qubits.front();
        if (lv < lowest) { // identity below the lowest block qubit, the entry rides on the edge
            const cplx v = b.m[r * b.dim() + c];
            if (v == cplx(0, 0)) return Edge{};
            return Edge{identChain(lv), v};
        }
        const uint64_t memoKey = (static_cast<uint64_t>(lv) << 40) | (static_cast<uint64_t>(r) << 20) | static_cast<uint64_t>(c);
        const auto hit = memo_.find(memoKey);
        if (hit != memo_.end()) return hit->second;
        std::array<Edge, 4> e{};
        const int p = posOf_[static_cast<std::size_t>(lv)];
        if (p >= 0) {
            for (std::size_t rb = 0; rb < 2; ++rb) {
                for (std::size_t cb = 0; cb < 2; ++cb) e[2 * rb + cb] = make(lv - 1, r | (rb << p), c | (cb << p));
            }
        } else {
            e[0] = e[3] = make(lv - 1, r, c);
        }
        // normalise: the largest weight (first on ties) becomes 1 and moves to the incoming edge
        int top = -1;
        double best = 0.0;
        for (int k = 0; k < 4; ++k) {
            const double mag = std::norm(e[k].w);
            if (mag > best * (1.0 + 1e-12)) {
                best = mag;
                top = k;
            }
        }
        Edge res;
        if (top >= 0) {
            const cplx f = e[top].w;
            for (int k = 0; k < 4; ++k) {
                if (e[k].w == cplx(0, 0)) continue;
                if (k == top || e[k].w == f) {
                    e[k].w = cplx(1, 0); // x / x is not exactly 1 in complex arithmetic; identity levels must stay exact
                } else {
                    cplx q = e[k].w / f;
                    if (std::abs(q.real()) < 1e-15) q.real(0.0); // rounding dust of the division, far below the DD tolerance
                    if (std::abs(q.imag()) < 1e-15) q.imag(0.0);
                    e[k].w = q;
                }
            }
            res.node = addNode(lv, e, lv == 0);
            res.w = f;
        }
        memo_.emplace(memoKey, res);
        return res;
    }

    struct KeyHash {
        std::size_t operator()(const std::array<int64_t, 13>& k) const {
            uint64_t h = 1469598103934665603ULL;
            for (int64_t v : k) {
                h ^= static_cast<uint64_t>(v);
                h *= 1099511628211ULL;
            }
            return static_cast<std::size_t>(h);
        }
    };
    int n_;
    FlatMatDD out_;
    const Block* block_ = nullptr;
    std::vector<int> posOf_;
    std::vector<int32_t> ident_;
    std::unordered_map<std::array<int64_t, 13>, int32_t, KeyHash> unique_;
    std::unordered_map<uint64_t, Edge> memo_;
};

// |0...0> as a vector DD: one node per level, weight 1 on the 0-successor
inline FlatVecDD zeroStateDD(int nQubits) {
    FlatVecDD v;
    v.n_qubits = nQubits;
    v.root = 0;
    v.root_weight[0] = 1.0;
    for (int i = 0; i < nQubits; ++i) {
        v.level.push_back(nQubits - 1 - i);
        v.child.push_back(i + 1 < nQubits ? i + 1 : FDD_TERMINAL);
        v.child.push_back(FDD_TERMINAL);
        v.weight.insert(v.weight.end(), {1.0, 0.0, 0.0, 0.0});
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// the simulator
// ------------------------------------------------------------------------------------------------
struct FusionPolicy {
    int maxBlockQubits = 5;  // qubits of one fused block (diagonal ones included)
    int maxNonDiagonal = 4;  // of which non-diagonal: sizes the kernel's tile (16 segments: the tensor-core path)
    double budgetFactor = 2.2; // fuse 2: a block that touches a warp-lane qubit must stay within this many HBM passes (fdd_cost_gpu)
    double hbmGBs = 6500.0;
    // fuse 3 (table-based dense-block fusion, the policy of GpuSwitchSimulator's --fuse 4): a block has at most blockTargets
    // non-diagonal qubits anywhere and at most blockContext qubits it depends on diagonally
    int blockTargets = 4;
    int blockContext = 5;
};

class FlatStartSimulator {
public:
    FlatStartSimulator(Circuit circuit, ArrayBackend* backend) : qc_(std::move(circuit)), backend_(backend) {}

    // knobs with the names of the reference simulator where they still mean something
    unsigned fuse = 0;    // 0: one launch per gate; 1: dense-block fusion with commuting open blocks; 2: dependency-graph dense-block fusion
                          // (dense matrices, GPU cost model); 3: dependency-graph fusion on block tables (block_fusion.hpp) for the tile-resident kernel
    bool verbose = true;
    bool stateLoaded = false; // the backend already holds the initial state (resume from a dumped state)
    // sharded state: the top log2(worldSize) PHYSICAL qubits are global.  Blocks are built in
    // physical positions; an operation that is non-diagonal on a global qubit first swaps it with the local qubit
    // whose next non-diagonal use lies furthest ahead (Belady), like GpuSwitchSimulator::planExchanges.  fuse >= 2 only.
    int worldSize = 1;
    std::size_t lookahead = 4096;
    std::size_t exchanges = 0;
    FusionPolicy policy;

    // results
    std::size_t unitaryOps = 0;
    std::size_t launches = 0;
    double arrayPhaseTime = 0.0;   // seconds, wall clock around the launches (synchronised at the end)
    double gateMergingTime = 0.0;  // seconds spent fusing and building gate DDs
    double kernelMsTotal = 0.0;    // with per-launch timing on
    std::vector<double> timeRecord2;

    void simulate() {
        const auto t0 = std::chrono::steady_clock::now();
        if (!stateLoaded) backend_->convert(zeroStateDD(qc_.nQubits));
        GateDDBuilder builder(qc_.nQubits);
        // Open blocks have pairwise disjoint qubit sets, so they commute: an operation only has to come after
        // the blocks it shares a qubit with, and the others stay open for later operations (a layer of
        // one-qubit gates over the whole register ends up in ceil(n / maxNonDiagonal) blocks instead of n).
        std::vector<Open> open;
        std::size_t opIndex = 0;
        auto emit = [&](const Open& o) {
            const auto m0 = std::chrono::steady_clock::now();
            const FlatMatDD gate = builder.build(o.block);
            gateMergingTime += seconds(m0);
            const auto a0 = std::chrono::steady_clock::now();
            backend_->apply(gate, o.count);
            kernelMsTotal += backend_->lastKernelMs();
            timeRecord2.push_back(seconds(a0));
            ++launches;
        };
        auto emitFlat = [&](const FlatMatDD& gate, int count) {
            const auto a0 = std::chrono::steady_clock::now();
            backend_->apply(gate, count);
            kernelMsTotal += backend_->lastKernelMs();
            timeRecord2.push_back(seconds(a0));
            ++launches;
        };
        if (worldSize > 1 && fuse < 2) throw std::runtime_error("a sharded state needs the dependency-graph schedule (--fuse 2)");
        if (fuse >= 2) {
            simulateDag(emit, emitFlat);
            backend_->synchronize();
            arrayPhaseTime = seconds(t0) - gateMergingTime;
            if (verbose) {
                std::printf("Gate merging time: %g\n", gateMergingTime);
                std::printf("Merged Gate Number: %zu\n", launches);
            }
            return;
        }
        for (const Op& op : qc_.ops) {
            if (verbose && opIndex % 100 == 0) std::printf("[Instruction Count]  %zu\n", opIndex);
            ++opIndex;
            if (op.kind == Op::Measure || op.kind == Op::Barrier) continue; // skipped like the reference (src/SwitchSimulator.cpp:108-113)
            if (op.kind == Op::Reset) throw std::runtime_error("reset is not supported");
            ++unitaryOps;
            const auto m0 = std::chrono::steady_clock::now();
            Open next{blockOf(op), 1};
            if (fuse == 0) {
                gateMergingTime += seconds(m0);
                emit(next);
                continue;
            }
            std::vector<std::size_t> touching;
            std::vector<int> all = next.block.qubits;
            for (std::size_t i = 0; i < open.size(); ++i) {
                std::vector<int> common;
                std::set_intersection(open[i].block.qubits.begin(), open[i].block.qubits.end(), next.block.qubits.begin(), next.block.qubits.end(),
                                      std::back_inserter(common));
                if (common.empty()) continue;
                touching.push_back(i);
                std::vector<int> merged;
                std::set_union(all.begin(), all.end(), open[i].block.qubits.begin(), open[i].block.qubits.end(), std::back_inserter(merged));
                all.swap(merged);
            }
            bool fused = false;
            if (static_cast<int>(all.size()) <= policy.maxBlockQubits) {
                Open candidate = next;
                for (std::size_t i : touching) { // the open blocks commute with each other; the new operation comes last
                    candidate.block = multiply(candidate.block, open[i].block);
                    candidate.count += open[i].count;
                }
                if (nonDiagonalCount(candidate.block) <= policy.maxNonDiagonal) {
                    next = std::move(candidate);
                    fused = true;
                }
            }
            gateMergingTime += seconds(m0);
            if (!fused) {
                // the touching blocks have to run now; blocks on disjoint qubits may share their launch
                // (a Kronecker product) as long as the policy holds
                std::vector<Open> leaving;
                for (std::size_t i : touching) leaving.push_back(std::move(open[i]));
                for (Open& o : pack(leaving)) emit(o);
            }
            for (auto it = touching.rbegin(); it != touching.rend(); ++it) open.erase(open.begin() + static_cast<std::ptrdiff_t>(*it));
            open.push_back(std::move(next));
        }
        for (Open& o : pack(open)) emit(o);
        backend_->synchronize();
        arrayPhaseTime = seconds(t0) - gateMergingTime;
        if (verbose) {
            std::printf("Gate merging time: %g\n", gateMergingTime);
            std::printf("Merged Gate Number: %zu\n", launches);
        }
    }

    [[nodiscard]] const Circuit& circuit() const { return qc_; }
    void getVector(std::vector<double>& re, std::vector<double>& im) {
        re.resize(std::size_t{1} << qc_.nQubits);
        im.resize(re.size());
        backend_->getState(re.data(), im.data());
    }

private:
    struct Open {
        Block block;
        int count = 0;
    };
    // fuse >= 2: dependency-graph fusion.  Two operations are ordered only if they share a qubit; one block at a time is
    // grown from ALL operations whose predecessors are done (earliest first) while the policy holds — the dense-block
    // counterpart of GpuSwitchSimulator::buildScheduleDag.
    template <class Emit, class EmitFlat> void simulateDag(Emit&& emit, EmitFlat&& emitFlat) {
        std::vector<const Op*> ops;
        for (const Op& op : qc_.ops) {
            if (op.kind == Op::Measure || op.kind == Op::Barrier) continue;
            if (op.kind == Op::Reset) throw std::runtime_error("reset is not supported");
            ops.push_back(&op);
        }
        unitaryOps = ops.size();
        const std::size_t count = ops.size();
        std::vector<std::vector<std::size_t>> succ(count);
        std::vector<int> indeg(count, 0);
        {
            std::vector<long> lastOn(static_cast<std::size_t>(qc_.nQubits), -1);
            for (std::size_t i = 0; i < count; ++i) {
                std::vector<long> preds;
                for (int q : ops[i]->qubits) {
                    const long p = lastOn[static_cast<std::size_t>(q)];
                    if (p >= 0 && p != static_cast<long>(i) && std::find(preds.begin(), preds.end(), p) == preds.end()) preds.push_back(p);
                    lastOn[static_cast<std::size_t>(q)] = static_cast<long>(i);
                }
                for (long p : preds) {
                    succ[static_cast<std::size_t>(p)].push_back(i);
                    ++indeg[i];
                }
            }
        }
        std::vector<std::size_t> ready; // sorted: program order
        for (std::size_t i = 0; i < count; ++i) {
            if (indeg[i] == 0) ready.push_back(i);
        }
        // sharded: logical -> physical map and the non-diagonal (logical) qubits of every operation
        int globalBits = 0;
        while ((1 << globalBits) < worldSize) ++globalBits;
        const int local = qc_.nQubits - globalBits;
        std::vector<int> perm(static_cast<std::size_t>(qc_.nQubits));
        for (int q = 0; q < qc_.nQubits; ++q) perm[static_cast<std::size_t>(q)] = q;
        std::vector<std::vector<int>> nonDiag;
        if (worldSize > 1) {
            if (local < 5) throw std::runtime_error("a shard must hold at least 5 qubits");
            for (const Op* op : ops) nonDiag.push_back(nonDiagonalQubits(blockOf(*op)));
        }
        auto needsGlobal = [&](std::size_t i) {
            if (worldSize <= 1) return false;
            for (int q : nonDiag[i]) {
                if (perm[static_cast<std::size_t>(q)] >= local) return true;
            }
            return false;
        };
        std::size_t epoch = 1; // layout version: every exchange renames physical positions
        auto planExchanges = [&](std::size_t k) {
            for (int q : nonDiag[k]) {
                const int pq = perm[static_cast<std::size_t>(q)];
                if (pq < local) continue;
                int victim = -1, victimPos = -1;
                std::size_t victimUse = 0;
                const int floorPos = local > 6 ? 3 : 0; // runs shorter than 128 bytes waste NVLink sectors
                for (int cand = 0; cand < qc_.nQubits; ++cand) {
                    const int pc = perm[static_cast<std::size_t>(cand)];
                    if (pc >= local || pc < floorPos) continue;
                    if (std::find(nonDiag[k].begin(), nonDiag[k].end(), cand) != nonDiag[k].end()) continue;
                    std::size_t use = k + 1 + lookahead; // "never" within the window
                    for (std::size_t j = k + 1; j < count && j <= k + lookahead; ++j) {
                        if (std::find(nonDiag[j].begin(), nonDiag[j].end(), cand) != nonDiag[j].end()) {
                            use = j;
                            break;
                        }
                    }
                    if (victim < 0 || use > victimUse || (use == victimUse && pc > victimPos)) {
                        victim = cand;
                        victimUse = use;
                        victimPos = pc;
                    }
                }
                if (victim < 0) throw std::runtime_error("no local qubit available for the exchange (gate touches too many qubits)");
                backend_->exchange(pq, victimPos);
                ++exchanges;
                ++epoch;
                std::swap(perm[static_cast<std::size_t>(q)], perm[static_cast<std::size_t>(victim)]);
            }
        };
        auto physicalBlock = [&](const Op& op) { return worldSize > 1 ? relabel(blockOf(op), perm) : blockOf(op); };
        // fuse 3: every operation as a block table, valid while the layout (epoch) does not change
        std::vector<SmallGate> memo(fuse >= 3 ? count : 0);
        std::vector<std::size_t> memoEpoch(fuse >= 3 ? count : 0, 0);
        std::size_t skipped = 0;
        BlockDDBuilder tableBuilder(qc_.nQubits);
        std::size_t done = 0;
        while (done < count) {
            const auto m0 = std::chrono::steady_clock::now();
            if (worldSize > 1) {
                // if nothing ready can run under the current layout, remap for the earliest ready operation
                bool any = false;
                for (std::size_t i : ready) any = any || !needsGlobal(i);
                if (!any) planExchanges(ready.front());
            }
            if (fuse >= 3) {
                // Table-based fusion (the selection rule of GpuSwitchSimulator::buildScheduleBlocks): operations that fit the open
                // block without a new target qubit go first, in program order; when none is left the block grows by the ready
                // operation that adds the fewest targets; an operation that is no block of the policy's size goes through alone.
                // Merging is BlockAcc::apply (a few thousand multiply-adds), a DD is built once per emitted block.
                BlockAcc block;
                int blockOps = 0;
                long wide = -1; // ready-list slot of an operation that has to go through on its own
                bool progress = true;
                while (progress) {
                    progress = false;
                    long pick = -1;
                    std::size_t pickGrowth = 99;
                    for (std::size_t r = 0; r < ready.size(); ++r) {
                        const std::size_t i = ready[r];
                        if (needsGlobal(i)) continue;
                        if (memoEpoch[i] != epoch) {
                            memo[i] = smallGateOf(physicalBlock(*ops[i]));
                            memoEpoch[i] = epoch;
                        }
                        const SmallGate& g = memo[i];
                        if (g.isIdentity()) {
                            pick = static_cast<long>(r);
                            pickGrowth = 0;
                            break;
                        }
                        if (static_cast<int>(g.targets.size()) > policy.blockTargets || static_cast<int>(g.ctx.size()) > policy.blockContext) {
                            if (blockOps == 0 && pick < 0) {
                                pick = static_cast<long>(r);
                                pickGrowth = 98;
                            }
                            continue;
                        }
                        std::vector<int> t2, c2;
                        block.merged(g, t2, c2);
                        if (static_cast<int>(t2.size()) > policy.blockTargets || static_cast<int>(c2.size()) > policy.blockContext) continue;
                        const std::size_t growth = t2.size() - block.targets.size();
                        if (growth < pickGrowth) {
                            pick = static_cast<long>(r);
                            pickGrowth = growth;
                            if (growth == 0) break;
                        }
                    }
                    if (pick < 0) break;
                    const std::size_t i = ready[static_cast<std::size_t>(pick)];
                    if (pickGrowth == 98) {
                        wide = pick;
                        break;
                    }
                    if (!memo[i].isIdentity()) {
                        block.apply(memo[i]);
                        ++blockOps;
                    } else {
                        ++skipped;
                    }
                    ready.erase(ready.begin() + pick);
                    for (std::size_t nxt : succ[i]) {
                        if (--indeg[nxt] == 0) ready.insert(std::lower_bound(ready.begin(), ready.end(), nxt), nxt);
                    }
                    ++done;
                    progress = true;
                }
                if (wide >= 0) {
                    const std::size_t i = ready[static_cast<std::size_t>(wide)];
                    Open alone{physicalBlock(*ops[i]), 1};
                    ready.erase(ready.begin() + wide);
                    for (std::size_t nxt : succ[i]) {
                        if (--indeg[nxt] == 0) ready.insert(std::lower_bound(ready.begin(), ready.end(), nxt), nxt);
                    }
                    ++done;
                    gateMergingTime += seconds(m0);
                    emit(alone);
                    continue;
                }
                if (blockOps == 0) {
                    gateMergingTime += seconds(m0);
                    if (done >= count) break; // only identities were left
                    if (skipped > 0) {
                        skipped = 0;
                        continue;
                    }
                    throw std::runtime_error("dense-block fusion made no progress");
                }
                skipped = 0;
                const bool identity = block.targets.empty() && block.ctx.empty() && block.table[0] == cplx(1.0, 0.0);
                FlatMatDD gate;
                if (!identity) gate = tableBuilder.build(block.targets, block.ctx, block.table);
                gateMergingTime += seconds(m0);
                if (!identity) emitFlat(gate, blockOps);
                continue;
            }
            Open current;
            bool progress = true;
            while (progress) {
                progress = false;
                for (std::size_t r = 0; r < ready.size(); ++r) {
                    const std::size_t i = ready[r];
                    if (needsGlobal(i)) continue;
                    const Block next = physicalBlock(*ops[i]);
                    std::vector<int> all;
                    std::set_union(next.qubits.begin(), next.qubits.end(), current.block.qubits.begin(), current.block.qubits.end(), std::back_inserter(all));
                    if (current.count > 0 && static_cast<int>(all.size()) > policy.maxBlockQubits) continue;
                    Block candidate = current.count > 0 ? multiply(next, current.block) : next;
                    if (current.count > 0 && nonDiagonalCount(candidate) > policy.maxNonDiagonal) continue; // a block of one operation is always allowed
                    // blocks that stay above the warp lanes run at ~1.2-1.3 passes whatever they hold (tensor-core path); a block
                    // that touches a lane qubit is priced by the GPU cost model, like GpuSwitchSimulator::buildScheduleDag does
                    if (current.count > 0 && candidate.qubits.front() < 5 && !withinBudget(candidate)) continue;
                    current.block = std::move(candidate);
                    ++current.count;
                    ready.erase(ready.begin() + static_cast<std::ptrdiff_t>(r));
                    for (std::size_t nxt : succ[i]) {
                        if (--indeg[nxt] == 0) ready.insert(std::lower_bound(ready.begin(), ready.end(), nxt), nxt);
                    }
                    ++done;
                    progress = true;
                    break; // rescan from the earliest ready operation
                }
            }
            gateMergingTime += seconds(m0);
            if (current.count == 0) throw std::runtime_error("dependency-graph fusion made no progress");
            emit(current);
        }
    }
    bool withinBudget(const Block& b) {
        GateDDBuilder builder(qc_.nQubits);
        const FlatMatDD dd = builder.build(b);
        const fdd_matdd m = view(dd);
        double ns = 0.0;
        const int rc = fdd_cost_gpu(&m, policy.hbmGBs, 30000.0, &ns);
        if (rc == FDD_ERR_TOO_DENSE) return false;
        fddCheck(rc, "fdd_cost_gpu");
        const double memNs = 32.0 * std::ldexp(1.0, qc_.nQubits) / policy.hbmGBs;
        return ns - 3000.0 <= policy.budgetFactor * memNs; // 3000 ns = launch term of the model
    }
    // first-fit packing of pairwise disjoint blocks into as few launches as the policy allows
    std::vector<Open> pack(std::vector<Open>& blocks) const {
        std::vector<Open> out;
        for (Open& b : blocks) {
            bool placed = false;
            for (Open& o : out) {
                std::vector<int> all;
                std::set_union(o.block.qubits.begin(), o.block.qubits.end(), b.block.qubits.begin(), b.block.qubits.end(), std::back_inserter(all));
                if (static_cast<int>(all.size()) > policy.maxBlockQubits) continue;
                if (nonDiagonalCount(o.block) + nonDiagonalCount(b.block) > policy.maxNonDiagonal) continue;
                o.block = multiply(b.block, o.block);
                o.count += b.count;
                placed = true;
                break;
            }
            if (!placed) out.push_back(std::move(b));
        }
        return out;
    }
    static double seconds(std::chrono::steady_clock::time_point since) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - since).count();
    }
    Circuit qc_;
    ArrayBackend* backend_;
};

} // namespace fddb200::standalone

// reference_binding.hpp — glue between the host driver (gpu_switch_simulator.hpp) and a checkout
// of the reference's host-side code (circuit IR, OpenQASM parser and DD package; reference
// include/QuantumComputation.hpp, include/dd/SwitchPackage.hpp, include/dd/Operations.hpp).
// Those layers produce the hot path's inputs and stay as they are (SURVEY.md section 2, rows
// 12/13/15/16 "reused"); compile with -I<reference>/include.  This header is the only file of
// the product that names reference types.
#pragma once

#include "QuantumComputation.hpp"
#include "dd/Operations.hpp"
#include "dd/SwitchPackage.hpp"
#include "gpu_switch_simulator.hpp"

#include <sstream>

namespace fddb200 {

// Edge weights of the reference package are pairs of tagged pointers into the real-number
// table: read them with RealNumber::val (src/dd/RealNumber.cpp:43-49), test them with
// exactlyZero (pointer compare against the zero constant).
struct RefWeightTraits {
    static double re(const dd::Complex& w) { return dd::RealNumber::val(w.r); }
    static double im(const dd::Complex& w) { return dd::RealNumber::val(w.i); }
    static bool isZero(const dd::Complex& w) { return w.exactlyZero(); }
};

using RefPackage = dd::SwitchPackage<dd::DDPackageConfig>;

struct RefDdOps {
    using Permutation = qc::Permutation;
    static qc::MatrixDD getDD(const qc::Operation* op, std::unique_ptr<RefPackage>& pkg) { return dd::getDD(op, pkg); }
    // gate DD with its qubits relabelled logical -> physical (include/dd/Operations.hpp:591-678)
    static qc::MatrixDD getDD(const qc::Operation* op, std::unique_ptr<RefPackage>& pkg, Permutation& perm) {
        return dd::getDD(op, pkg, perm);
    }
    // with a permutation the reference turns an uncontrolled SWAP into a relabelling (Operations.hpp:611-620)
    static bool isRelabelSwap(const qc::Operation& op) { return op.getType() == qc::SWAP && !op.isControlled(); }
    static std::pair<int, int> swapTargets(const qc::Operation& op) {
        return {static_cast<int>(op.getTargets()[0]), static_cast<int>(op.getTargets()[1])};
    }
    // logical qubits on which the operation is NOT diagonal in the computational basis; controls are
    // always diagonal, targets are unless the gate type is a phase-type gate
    static std::vector<int> nonDiagonalQubits(const qc::Operation& op) {
        std::vector<int> out;
        if (op.isNonUnitaryOperation()) return out;
        if (const auto* compound = dynamic_cast<const qc::CompoundOperation*>(&op)) {
            for (const auto& sub : *compound) {
                for (int q : nonDiagonalQubits(*sub)) out.push_back(q);
            }
            return out;
        }
        switch (op.getType()) {
            case qc::I: case qc::Z: case qc::S: case qc::Sdag: case qc::T: case qc::Tdag: case qc::Phase: case qc::RZ:
            case qc::RZZ: case qc::GPhase: case qc::Barrier:
                return out;
            default:
                break;
        }
        for (const auto t : op.getTargets()) out.push_back(static_cast<int>(t));
        return out;
    }
    // targets on which the operation is dense (neither diagonal nor a permutation with phases): these
    // double the non-zeros per row of a fused block; X/Y/SWAP/iSWAP/DCX targets only move amplitudes
    static std::vector<int> denseQubits(const qc::Operation& op) {
        std::vector<int> out;
        if (op.isNonUnitaryOperation()) return out;
        if (const auto* compound = dynamic_cast<const qc::CompoundOperation*>(&op)) {
            for (const auto& sub : *compound) {
                for (int q : denseQubits(*sub)) out.push_back(q);
            }
            return out;
        }
        switch (op.getType()) {
            case qc::I: case qc::Z: case qc::S: case qc::Sdag: case qc::T: case qc::Tdag: case qc::Phase: case qc::RZ:
            case qc::RZZ: case qc::GPhase: case qc::Barrier: case qc::X: case qc::Y: case qc::SWAP: case qc::iSWAP: case qc::DCX:
                return out;
            default:
                break;
        }
        for (const auto t : op.getTargets()) out.push_back(static_cast<int>(t));
        return out;
    }
    // every qubit the operation touches (targets and controls)
    static std::vector<int> allQubits(const qc::Operation& op) {
        std::vector<int> out;
        if (const auto* compound = dynamic_cast<const qc::CompoundOperation*>(&op)) {
            for (const auto& sub : *compound) {
                for (int q : allQubits(*sub)) out.push_back(q);
            }
            return out;
        }
        for (const auto t : op.getTargets()) out.push_back(static_cast<int>(t));
        for (const auto& c : op.getControls()) out.push_back(static_cast<int>(c.qubit));
        return out;
    }
    // replaces every compound operation from index `from` on by clones of its sub-operations (recursively)
    static bool expandCompound(qc::QuantumComputation& circuit, std::size_t from) {
        bool changed = false;
        auto& ops = circuit.ops;
        for (std::size_t k = from; k < ops.size();) {
            const auto* compound = dynamic_cast<const qc::CompoundOperation*>(ops[k].get());
            if (compound == nullptr) {
                ++k;
                continue;
            }
            std::vector<std::unique_ptr<qc::Operation>> subs;
            for (const auto& sub : *compound) subs.push_back(sub->clone());
            ops.erase(ops.begin() + static_cast<std::ptrdiff_t>(k));
            ops.insert(ops.begin() + static_cast<std::ptrdiff_t>(k), std::make_move_iterator(subs.begin()), std::make_move_iterator(subs.end()));
            changed = true; // k stays: a nested compound operation is expanded on the next pass
        }
        return changed;
    }
    // everything that defines the operation's matrix except WHICH qubits it sits on (empty: do not cache — compound operations)
    static std::string signature(const qc::Operation& op) {
        if (dynamic_cast<const qc::CompoundOperation*>(&op) != nullptr || op.isNonUnitaryOperation()) return {};
        std::ostringstream out;
        out << static_cast<int>(op.getType()) << '|' << op.getTargets().size() << '|';
        for (const auto& c : op.getControls()) out << (c.type == qc::Control::Type::Pos ? 'p' : 'n');
        out << '|' << std::hexfloat;
        for (const auto p : op.getParameter()) out << p << ';';
        return out.str();
    }
    // controls (ascending, as the IR keeps them) followed by the targets in the operation's own order
    static std::vector<int> orderedQubits(const qc::Operation& op) {
        std::vector<int> out;
        for (const auto& c : op.getControls()) out.push_back(static_cast<int>(c.qubit));
        for (const auto t : op.getTargets()) out.push_back(static_cast<int>(t));
        return out;
    }
    static bool isMeasure(const qc::Operation& op) { return op.getType() == qc::Measure; }
    static bool isBarrier(const qc::Operation& op) { return op.getType() == qc::Barrier; }
    static bool isReset(const qc::Operation& op) { return op.getType() == qc::Reset; }
};

using RefGpuSwitchSimulator = GpuSwitchSimulator<RefPackage, qc::QuantumComputation, RefDdOps, RefWeightTraits>;

inline FlatVecDD flattenVector(const dd::vEdge& e, int nQubits) { return flatten<2, dd::vEdge, RefWeightTraits>(e, nQubits); }
inline FlatMatDD flattenMatrix(const dd::mEdge& e, int nQubits) { return flatten<4, dd::mEdge, RefWeightTraits>(e, nQubits); }

} // namespace fddb200

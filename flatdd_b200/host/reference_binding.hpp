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
    static qc::MatrixDD getDD(const qc::Operation* op, std::unique_ptr<RefPackage>& pkg) { return dd::getDD(op, pkg); }
    static bool isMeasure(const qc::Operation& op) { return op.getType() == qc::Measure; }
    static bool isBarrier(const qc::Operation& op) { return op.getType() == qc::Barrier; }
    static bool isReset(const qc::Operation& op) { return op.getType() == qc::Reset; }
};

using RefGpuSwitchSimulator = GpuSwitchSimulator<RefPackage, qc::QuantumComputation, RefDdOps, RefWeightTraits>;

inline FlatVecDD flattenVector(const dd::vEdge& e, int nQubits) { return flatten<2, dd::vEdge, RefWeightTraits>(e, nQubits); }
inline FlatMatDD flattenMatrix(const dd::mEdge& e, int nQubits) { return flatten<4, dd::mEdge, RefWeightTraits>(e, nQubits); }

} // namespace fddb200

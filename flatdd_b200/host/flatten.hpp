// flatten.hpp — serialise a decision diagram of the host DD package into the flat POD tables
// the C-ABI takes (include/flatdd_b200.h: fdd_vecdd / fdd_matdd).
//
// Host side of the drop-in boundary.  It is written against the *shape* of the reference's
// edge/node types (Edge{p, w}, Node{e[R], v}, terminal == nullptr; reference
// include/dd/Node.hpp:17-83, include/dd/Edge.hpp:12-14), not against their headers: the two
// accessors that touch the package's number representation are supplied by the caller
// (`WeightTraits`), because real parts live behind tagged pointers in the reference
// (src/dd/RealNumber.cpp:39-49) and must be read with RealNumber::val and tested for zero
// with Complex::exactlyZero (the predicate the reference kernels use,
// include/dd/SwitchPackage.hpp:2175, 2214; include/SwitchSimulator.hpp:269-274).
//
// The walk is a breadth-first numbering of distinct node pointers, the same idea as
// dd::serialize (reference include/dd/Export.hpp:760-865) and nodeCount
// (include/dd/SwitchPackage.hpp:3157-3169).
#pragma once

#include "flatdd_b200.h"

#include <cstdint>
#include <deque>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace fddb200 {

// Owning storage behind an fdd_vecdd / fdd_matdd view.
template <int R> struct FlatDD {
    int32_t n_qubits = 0;
    int32_t root = FDD_TERMINAL;
    double root_weight[2] = {0.0, 0.0};
    std::vector<int32_t> level;
    std::vector<int32_t> child;  // R per node
    std::vector<double> weight;  // 2*R per node

    [[nodiscard]] int32_t nNodes() const { return static_cast<int32_t>(level.size()); }
};

using FlatVecDD = FlatDD<2>;
using FlatMatDD = FlatDD<4>;

inline fdd_vecdd view(const FlatVecDD& f) {
    fdd_vecdd v{};
    v.n_qubits = f.n_qubits;
    v.n_nodes = f.nNodes();
    v.root = f.root;
    v.root_weight[0] = f.root_weight[0];
    v.root_weight[1] = f.root_weight[1];
    v.level = f.level.data();
    v.child = f.child.data();
    v.weight = f.weight.data();
    return v;
}

inline fdd_matdd view(const FlatMatDD& f) {
    fdd_matdd v{};
    v.n_qubits = f.n_qubits;
    v.n_nodes = f.nNodes();
    v.root = f.root;
    v.root_weight[0] = f.root_weight[0];
    v.root_weight[1] = f.root_weight[1];
    v.level = f.level.data();
    v.child = f.child.data();
    v.weight = f.weight.data();
    return v;
}

// WeightTraits must provide:  static double re(const W&), static double im(const W&),
//                             static bool isZero(const W&).
template <int R, class EdgeT, class WeightTraits>
FlatDD<R> flatten(const EdgeT& rootEdge, int nQubits) {
    using NodePtr = decltype(rootEdge.p);
    FlatDD<R> out;
    out.n_qubits = nQubits;
    out.root_weight[0] = WeightTraits::re(rootEdge.w);
    out.root_weight[1] = WeightTraits::im(rootEdge.w);
    if (rootEdge.p == nullptr) {
        // terminal root: a scalar.  Only meaningful for n_qubits == 0; reject otherwise.
        if (nQubits != 0) {
            throw std::runtime_error("flatten: terminal root edge for a non-empty register");
        }
        return out;
    }
    std::unordered_map<NodePtr, int32_t> index;
    std::deque<NodePtr> queue;
    index.emplace(rootEdge.p, 0);
    queue.push_back(rootEdge.p);
    out.root = 0;
    while (!queue.empty()) {
        NodePtr p = queue.front();
        queue.pop_front();
        out.level.push_back(static_cast<int32_t>(p->v));
        const std::size_t base = out.child.size();
        out.child.resize(base + R, FDD_TERMINAL);
        out.weight.resize(out.weight.size() + 2 * R, 0.0);
        for (int k = 0; k < R; ++k) {
            const auto& e = p->e[static_cast<std::size_t>(k)];
            if (WeightTraits::isZero(e.w)) {
                continue; // zero edge: child stays FDD_TERMINAL, weight stays (0,0)
            }
            out.weight[2 * (base + k)] = WeightTraits::re(e.w);
            out.weight[2 * (base + k) + 1] = WeightTraits::im(e.w);
            if (e.p == nullptr) {
                continue; // non-zero terminal edge
            }
            auto it = index.find(e.p);
            if (it == index.end()) {
                const auto id = static_cast<int32_t>(index.size());
                it = index.emplace(e.p, id).first;
                queue.push_back(e.p);
            }
            out.child[base + k] = it->second;
        }
    }
    return out;
}

} // namespace fddb200

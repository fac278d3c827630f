// block_compile.hpp — host side of the dense-block path: gate DD -> dense block, and the plan of a pass.
#pragma once

#include "block_plan.hpp"
#include "flatdd_b200.h"

#include <cstdint>
#include <vector>

namespace fddb200 {

// A fused gate as a dense block (see block_plan.hpp).  Canonical order: row/column index bit i belongs to targets[i]
// (ascending physical qubit), matrix index bit j to ctx[j] (ascending).
struct DenseBlock {
    int n = 0;                 // qubits of the register
    std::vector<int> targets;  // 0..kBlockMaxTargets non-diagonal qubits (0: the gate is diagonal)
    std::vector<int> ctx;      // qubits the matrix depends on diagonally
    std::vector<double> table; // [2^ctx][2^k][2^k][2]
    [[nodiscard]] int k() const { return static_cast<int>(targets.size()); }
    [[nodiscard]] std::size_t rows() const { return std::size_t{1} << targets.size(); }
};

// Expands a full-depth gate DD (reference mNode/mEdge, include/dd/Node.hpp:35-83) into a dense block when it has at most
// kBlockMaxTargets non-diagonal levels and at most `maxCtx` further levels that are not identity-like.  Matrix entries
// are the products of the edge weights along the DD path, root first (the reference's association order,
// include/dd/SwitchPackage.hpp:2221-2236).  Returns false when the gate does not fit.
bool denseBlockFromDD(const fdd_matdd& g, DenseBlock& out, int maxCtx = kBlockMaxCtx);

// Brings a block to the shape the kernel runs: 3 or 4 targets.  Context qubits are promoted to targets first (the matrix
// becomes block diagonal in them: fewer matrices to look up), then the lowest free LOCAL qubits pad an identity factor.
// `avoid`: qubits that must not be used for padding (none by default).
void padBlock(DenseBlock& b, int nLocal);

// Plans one pass over a shard of nLocal qubits for blocks[0..count): tile bits = the 5 lane bits, every target >= 5 and
// the lowest free bits up to `tileBits`.  Returns false when the blocks do not fit one tile (more than tileBits - 5
// upper targets in total), a target is not local, or a block has too many in-tile context bits for the fragment shape.
// On success fills everything of `pass` except y, z and the table pointers (left null).
bool planPass(const DenseBlock* const* blocks, int count, int nLocal, int rank, int tileBits, PassParams& pass);

// smallest tile (in bits) that holds the blocks, or -1
int minTileBits(const DenseBlock* const* blocks, int count, int nLocal);

} // namespace fddb200

// gate_compile.hpp — host-side preparation of a flat matrix DD for the sm_100a DMAVM kernels.
//
// Input: the full-depth gate DD as it crosses the C-ABI (fdd_matdd; reference mNode/mEdge,
// include/dd/Node.hpp:35-83).  Output: the two tables the walk kernel needs.
//
//   * SUB TABLES (levels S-1..0, S = min(5, n) — one warp lane per row of a 2^S "segment"):
//     for every node at level S-1 that an upper edge points to, the 2^S x 2^S sub-matrix is
//     expanded into ELL rows: K entries (column, weight) per row, ascending columns, padded with
//     zero weights.  A lane then needs K lookups instead of an S-level walk.
//   * UPPER NODES (levels n-1..S): the DD itself, with identity-like nodes ([a 0; 0 a], both
//     successors the same) compressed out of every edge, because the per-segment walk visits
//     one node per remaining level and fused gates are mostly identity levels.
//
// It also derives the facts the launcher and the cost model use: an upper bound on the number
// of upper paths per segment row (= distinct source segments one output segment reads), the
// ELL width, nnz (the reference's DMAVMACCountIP, include/dd/SwitchPackage.hpp:3285-3311).
#pragma once

#include "flatdd_b200.h"

#include <cstdint>
#include <string>
#include <vector>

namespace fddb200 {

struct alignas(16) UpperNode {
    int32_t child[4]; // >= 0: upper node index; FDD_TERMINAL (-1): zero edge; <= -2: sub table (-2 - id)
    int32_t level;
    int32_t slotBit;  // index of this level inside the tile bit set, -1 if the level is not a tile bit
    int32_t pad[2];
    double w[8];      // (re, im) per successor
};
static_assert(sizeof(UpperNode) == 96, "UpperNode layout is shared with the device code");

constexpr int32_t encodeSub(int32_t id) { return -2 - id; }
constexpr int32_t decodeSub(int32_t code) { return -2 - code; }

enum SubFlags : uint8_t {
    SUB_IDENTITY = 1, // K == 1, column == row, weight == 1
    SUB_DIAGONAL = 2, // K == 1, column == row
};

struct CompiledGate {
    int n = 0;        // qubits
    int segBits = 0;  // S
    int32_t root = FDD_TERMINAL; // encoded like UpperNode::child
    double rootW[2] = {0.0, 0.0};
    std::vector<UpperNode> upper;
    int nSub = 0;
    int kMax = 0;                 // ELL width of the tables (max over sub tables, padded to 2, 4 or 8)
    int kTrue = 0;                // largest number of non-zeros in one sub-table row
    std::vector<uint8_t> subCol;  // [nSub][kMax][32]
    std::vector<double> subW;     // [nSub][kMax][32][2]
    std::vector<int32_t> subK;    // [nSub] entries per row actually used
    std::vector<uint8_t> subFlags;
    // facts
    int maxPaths = 1;   // upper bound on source segments per output segment
    std::vector<int> subPaths; // [nSub] upper bound on the paths of one row that end in sub table s
    int stackCap = 1;   // DFS stack bound for the upper walk
    int upperDepth = 0; // longest chain of upper nodes on a path (after compression)
    uint64_t nnz = 0;   // non-zero matrix entries
    int nnzRowMax = 0;  // upper bound on non-zeros in one row (maxPaths * kMax)
    int topLevel = -1;  // highest level with a non-identity node (-1: scalar multiple of identity)
    bool diagonal = false; // every level diagonal
    bool allIdentitySubs = false; // the low S levels are untouched on every path
    // tile-staged launch: the 2^tileBits segments of a tile differ in the index bits `tileMask`
    // (positions relative to the segment index).  Valid when every upper level that is not
    // diagonal fits into the tile, so that all sources of a tile lie in the tile itself.
    bool tileable = false;
    int tileBits = 0;      // log2(segments per warp tile)
    int subTileBits = 0;   // popcount(tileMask)
    uint32_t tileMask = 0; // index bits of the non-diagonal upper levels (a sub-tile is closed under the gate)
    uint32_t fillMask = 0; // lowest free index bits that complete the warp tile
    int nonDiagUpper = 0; // upper levels with an off-diagonal successor
    bool uniform = false; // tileable and every upper node sits on a tile bit: all sub-tiles see the same entries
    uint64_t nonDiagMask = 0; // bit v set when level v has an off-diagonal successor (all levels)
    uint32_t ctxMask = 0;     // tileable gates: local segment-index bits of the upper levels outside the tile that still have nodes
                              // (diagonal levels the block depends on); 0 for uniform gates
};

// Throws std::runtime_error with a message on malformed input.
void validate(const fdd_matdd& g);
void validate(const fdd_vecdd& v);
// nLocal: qubits held by one shard (== g.n_qubits on one GPU); tile bits are local bits.
CompiledGate compileGate(const fdd_matdd& g, int nLocal = -1);

// Reference cost model, same results as oracle/flat_oracle.c but
// part of the product because the fusion pass consumes it.
uint64_t macCount(const fdd_matdd& g);
uint64_t costIP(const fdd_matdd& g, unsigned nThreadExp);
uint64_t costOP1(const fdd_matdd& g, unsigned nThreadExp);
// Device-time estimate of one launch in nanoseconds ("GPU cost model").
double costGpuNs(const CompiledGate& c, double hbmGBs, double fp64GFlops);

} // namespace fddb200

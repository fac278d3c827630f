// block_plan.hpp — launch plan of the tile-resident dense-block DMAVM kernel (block_kernel.cuh).
//
// A fused gate whose non-diagonal qubits are few is a DENSE BLOCK: a 2^k x 2^k complex matrix on k <= 4 target
// qubits that may depend (diagonally) on the values of some other qubits, its CONTEXT qubits (controls, phases):
//     z[i] = sum_c  M_ctx(i)[ row(i) ][ c ] * y[ i with its target bits replaced by c ]
// (north_star: "gate DDs are flattened to dense 2^k x 2^k blocks for the fused qubit set").  A PASS streams the
// state through shared memory once, in TILES of 2^tileBits amplitudes that are closed under every block of the pass
// (the tile index bits are the 5 lane bits, the upper target qubits of all blocks of the pass and the lowest free bits),
// and applies the blocks one after the other to the resident tile on the FP64 tensor cores.  HBM traffic of a pass:
// 32 B per amplitude however many blocks it holds.
//
// This header holds what the host planner, the kernel and the CPU emulator of the tests (tests/emu) share: the
// descriptors and the index arithmetic.  Everything is plain integer code.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define FDD_HD __host__ __device__ __forceinline__
#else
#define FDD_HD inline
#endif

namespace fddb200 {

constexpr int kBlockMaxTargets = 4;   // k: dense qubits of one block (matrix 2^k x 2^k)
constexpr int kBlockMaxCtx = 10;      // context qubits of one block (2^ctx matrices)
constexpr int kPassMaxBlocks = 6;     // blocks applied to a resident tile in one pass
constexpr int kPassMaxTileBits = 13;  // 2^13 amplitudes = 128 KiB of shared memory
constexpr uint32_t kTableInGlobal = 0xffffffffu;
constexpr int kLaneBits = 5;          // the low 5 index bits always belong to the tile (512-byte runs in HBM)

// ---- shared-memory layout of the tile ------------------------------------------------------------
// Amplitude t of the tile (t = tile-local index: bits 0..4 = lane bits, bits 5.. = the tile's upper bits in
// ascending order) lives in 16-byte unit t ^ fold(t): the low three unit bits (the eight 16-byte bank groups of a
// 128-byte wavefront) are XOR-ed with every higher 3-bit group, so an access whose eight lanes differ in any three
// tile bits with pairwise different positions mod 3 is conflict free.  The map is linear over GF(2):
// swz(a ^ b) == swz(a) ^ swz(b), which lets the kernel combine precomputed pieces with XOR.
FDD_HD uint32_t swz(uint32_t t) { return t ^ (((t >> 3) ^ (t >> 6) ^ (t >> 9) ^ (t >> 12)) & 7u); }

// spread the low bits of x over the set bits of mask (pdep)
FDD_HD uint32_t pdep32(uint32_t x, uint32_t mask) {
    uint32_t out = 0;
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        if (x & 1u) out |= m & (0u - m);
        x >>= 1;
    }
    return out;
}
// gather the bits of x selected by mask into the low bits (pext)
FDD_HD uint32_t pext32(uint32_t x, uint32_t mask) {
    uint32_t out = 0;
    int at = 0;
    for (uint32_t m = mask; m != 0; m &= m - 1, ++at) {
        if (x & (m & (0u - m))) out |= 1u << at;
    }
    return out;
}
// spread x over the ZERO bits of mask (the tile index -> segment index without the tile bits)
FDD_HD uint32_t spreadAround(uint32_t x, uint32_t mask) {
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        const uint32_t lowest = m & (0u - m);
        x = ((x & ~(lowest - 1u)) << 1) | (x & (lowest - 1u));
    }
    return x;
}

// One block of a pass as the kernel sees it.  Positions are TILE-LOCAL bit positions (0 .. tileBits-1) unless noted.
// The matrix table stays in the gate's canonical order (row/column index bit i <-> i-th target in ascending physical
// order, matrix index bit j <-> j-th context qubit in ascending physical order) whatever the tile of the pass is: the
// kernel's own row/column index ("sigma index": bit i lives at tile bit sigma[i], chosen per pass for conflict-free
// shared-memory access) is translated with `canon`, the matrix index is assembled bit by bit from `ctxSrc`.
struct BlockDesc {
    const double* table;    // device: [2^nCtx][rows][rows] complex (re, im), row-major; rows = 2^k
    int32_t k;              // 3 or 4 (smaller blocks are padded by the planner)
    int32_t nUnits;         // 2^(tileBits - k - 3): groups of 2^k rows x 8 columns
    uint8_t sigma[4];       // sigma-index bit i lives at tile bit sigma[i]
    uint8_t canon[16];      // sigma index -> canonical row/column index of the table
    uint8_t kappa[3];       // column-of-the-fragment bit j lives at tile bit kappa[j]
    uint8_t nUnitBits;
    uint8_t unitPos[12];    // unit index bit j lives at tile bit unitPos[j] (context bits last: they vary slowest)
    uint8_t nCtx;
    uint8_t ctxSrc[kBlockMaxCtx]; // matrix index bit j: tile bit (value < 32) or segment-index bit outside the tile (value - 32)
    uint8_t conflictWays;   // planner's estimate of the shared-memory bank conflict degree (1 = conflict free)
    uint8_t unitBit0IsCtx;  // 1 when even unit bit 0 is a context bit (two neighbouring units then differ in their matrix)
};

struct PassParams {
    const void* y;          // source state of this shard (double2*)
    void* z;                // destination state
    int32_t tileBits;       // log2(amplitudes per tile), 8..13
    int32_t nBlocks;
    uint32_t tileMask;      // segment-index bits (relative to the segment index = amplitude index >> 5) inside the tile
    uint32_t nTiles;
    uint32_t rankSegBits;   // rank << (nLocal - 5): the global (shard) bits of every segment index
    uint32_t warpLocal;     // 1: the top three unit bits of EVERY block are the same three tile bits, none of them a target of any block:
                            // compute warp w (of eight) touches the same eighth of the tile in every block, so the blocks of a pass need
                            // no barrier between them, only the warp's own order
    uint32_t tableSmem[kPassMaxBlocks]; // byte offset of the block's matrix table in the CTA's shared-memory table area when the pass
                            // stages it there (multi-block passes: every (tile, block) visit reloads the warp's A fragments, and from
                            // global memory that is an L2 round trip under full HBM load), else kTableInGlobal
    // Fused exchange (multi-GPU): the pass that precedes SWAP(global bit, local bit pl >= 5) writes the half of its result that
    // changes owner straight into the partner shard's buffer (peer stores over NVLink, per 512-byte segment) and keeps the other
    // half, so the transfer runs under the pass instead of after it.  Ordering: the flag words of comm.cuh.
    void* zPeer;                     // the partner's destination buffer (peer mapped), null: no exchange
    int32_t exchSegBit;              // pl - 5: the segment-index bit that is traded
    uint32_t exchMyBit;              // this shard's value of the global bit: segments with that value of the traded bit stay
    uint32_t exchEpoch;
    uint32_t* exchMyFlags;           // [0] my earlier launches are complete, [1] my stores into the partner's buffer are complete
    const uint32_t* exchPartnerFlags;
    unsigned int* exchCounter;       // CTAs that have finished their stores
    uint32_t debugSkip;     // experiments (FLATDD_B200_BLOCK_SKIP): bit 0 = no tensor-core work, bit 1 = no global loads/stores
    long long* debugClocks; // experiments (FLATDD_B200_BLOCK_CLOCKS): per compute warp {cycles waiting for tiles, cycles in the blocks, total}
    BlockDesc blocks[kPassMaxBlocks];
};

// ---- per-lane pieces of the fragment addresses (tile-local, before swz) -------------------------------
// B fragment of DMMA.8x8x4 (lane l: row l & 3 of the 4-row K slab, column l >> 2): rows 4 kt + (l & 3) of Y.
FDD_HD uint32_t laneOffB(const BlockDesc& b, int lane) {
    uint32_t t = 0;
    if (lane & 1) t |= 1u << b.sigma[0];
    if (lane & 2) t |= 1u << b.sigma[1];
    if (lane & 4) t |= 1u << b.kappa[0];
    if (lane & 8) t |= 1u << b.kappa[1];
    if (lane & 16) t |= 1u << b.kappa[2];
    return t;
}
FDD_HD uint32_t ktOff(const BlockDesc& b, int kt) { // K slab kt: matrix index bits 2, 3
    uint32_t t = 0;
    if (kt & 1) t |= 1u << b.sigma[2];
    if (kt & 2) t |= 1u << b.sigma[3];
    return t;
}
// D fragment (lane l: row l >> 2 of the 8-row M slab, columns 2 (l & 3) and 2 (l & 3) + 1): rows 8 mt + (l >> 2) of Z.
FDD_HD uint32_t laneOffD(const BlockDesc& b, int lane) {
    uint32_t t = 0;
    if (lane & 1) t |= 1u << b.kappa[1];
    if (lane & 2) t |= 1u << b.kappa[2];
    if (lane & 4) t |= 1u << b.sigma[0];
    if (lane & 8) t |= 1u << b.sigma[1];
    if (lane & 16) t |= 1u << b.sigma[2];
    return t;
}
FDD_HD uint32_t mtOff(const BlockDesc& b, int mt) { return (mt & 1) ? 1u << b.sigma[3] : 0u; }
FDD_HD uint32_t unitOff(const BlockDesc& b, uint32_t unit) {
    uint32_t t = 0;
    for (int j = 0; j < b.nUnitBits; ++j) {
        if ((unit >> j) & 1u) t |= 1u << b.unitPos[j];
    }
    return t;
}
// matrix index of a unit: context bits inside the tile come from the unit's tile offset, the others from the segment
// index of the tile (which carries the shard's rank in its top bits)
FDD_HD uint32_t ctxIndex(const BlockDesc& b, uint32_t unitTileOff, uint32_t segBase) {
    uint32_t idx = 0;
    for (int j = 0; j < b.nCtx; ++j) {
        const uint32_t src = b.ctxSrc[j];
        const uint32_t bit = src < 32u ? (unitTileOff >> src) & 1u : (segBase >> (src - 32u)) & 1u;
        idx |= bit << j;
    }
    return idx;
}

} // namespace fddb200

// block_emu.cpp — CPU emulation of dmavm_block_kernel (flatdd_b200/csrc/block_kernel.cuh).  TEST INFRASTRUCTURE ONLY.
//
// This container has no GPU, so the index arithmetic of the tile-resident dense-block kernel is checked here before it
// ever runs on a device: the emulator executes the kernel's loops lane by lane with the SAME plan (planPass), the SAME
// shared helpers (block_plan.hpp: swz, laneOffB/D, ktOff, mtOff, unitOff, ctx sources) and the fragment layouts of
// mma.sync.m8n8k4.f64 (A[l>>2][l&3], B[l&3][l>>2], D[l>>2][2(l&3)+{0,1}]), on a plain array standing in for shared memory.
// What it cannot check: cp.async / barriers / the DMMA instruction itself (those are covered by the GPU tests).
#include "block_compile.hpp"
#include "gate_compile.hpp"

#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace fddb200;
using cplx = std::complex<double>;

namespace {

// owner: per 16-byte slot of the tile the compute warp (of eight) that touched it in an earlier block of a warp-local pass
template <int K> void applyBlockEmu(const BlockDesc& b, const double* table, std::vector<cplx>& tile, uint32_t segBaseWithRank, int* worstConflict,
                                    std::vector<int>* owner, int* violations) {
    constexpr int ROWS = 1 << K, MT = ROWS / 8, KTL = ROWS / 4;
    auto claim = [&](uint32_t at, uint32_t unit) {
        if (owner == nullptr) return;
        const int w = static_cast<int>(unit / (static_cast<uint32_t>(b.nUnits) / 8u));
        if ((*owner)[at] >= 0 && (*owner)[at] != w) ++*violations;
        (*owner)[at] = w;
    };
    for (uint32_t u = 0; u < static_cast<uint32_t>(b.nUnits); ++u) {
        const uint32_t off = unitOff(b, u);
        const uint32_t ctx = ctxIndex(b, off, segBaseWithRank);
        const double* m = table + static_cast<size_t>(ctx) * ROWS * ROWS * 2;
        const uint32_t pu = swz(off);
        // B fragments of every lane and K slab
        cplx y[KTL][32];
        for (int kt = 0; kt < KTL; ++kt) {
            int hits[4][8] = {};
            for (int lane = 0; lane < 32; ++lane) {
                const uint32_t at = pu ^ swz(laneOffB(b, lane)) ^ swz(ktOff(b, kt));
                if (at != swz(off | laneOffB(b, lane) | ktOff(b, kt))) std::fprintf(stderr, "emu: swizzle pieces do not combine\n");
                y[kt][lane] = tile[at];
                claim(at, u);
                ++hits[lane >> 3][at & 7u];
            }
            for (auto& q : hits) {
                for (int h : q) *worstConflict = std::max(*worstConflict, h);
            }
        }
        // D = M Y per 8-row slab, fragment semantics
        for (int mt = 0; mt < MT; ++mt) {
            cplx d[32][2];
            for (int lane = 0; lane < 32; ++lane) {
                for (int e = 0; e < 2; ++e) {
                    const int row = 8 * mt + (lane >> 2); // sigma index
                    const int n = 2 * (lane & 3) + e;     // fragment column
                    // the kernel's three real products (BlockRunner::issue): K1 = Mr (Yr + Yi) first, then the Re chain
                    // K1 + (-(Mr + Mi)) Yi and the Im chain K1 + (Mi - Mr) Yr continue from it, K slabs in order
                    double k1 = 0.0;
                    for (int kt = 0; kt < KTL; ++kt) {
                        for (int kk = 0; kk < 4; ++kk) {
                            const int col = 4 * kt + kk; // sigma index
                            const size_t at = 2 * (static_cast<size_t>(b.canon[row]) * ROWS + b.canon[col]);
                            const cplx yv = y[kt][4 * n + kk]; // the lane that holds B[kk][n] is 4 n + kk
                            k1 += m[at] * (yv.real() + yv.imag());
                        }
                    }
                    double zr = k1, zi = k1;
                    for (int kt = 0; kt < KTL; ++kt) {
                        for (int kk = 0; kk < 4; ++kk) {
                            const int col = 4 * kt + kk;
                            const size_t at = 2 * (static_cast<size_t>(b.canon[row]) * ROWS + b.canon[col]);
                            const cplx yv = y[kt][4 * n + kk];
                            zr += (-m[at] - m[at + 1]) * yv.imag();
                            zi += (m[at + 1] - m[at]) * yv.real();
                        }
                    }
                    d[lane][e] = cplx(zr, zi);
                }
            }
            int hits[2][4][8] = {};
            for (int lane = 0; lane < 32; ++lane) {
                const uint32_t at = pu ^ swz(laneOffD(b, lane)) ^ swz(mtOff(b, mt));
                tile[at] = d[lane][0];
                tile[at ^ swz(1u << b.kappa[0])] = d[lane][1];
                claim(at, u);
                claim(at ^ swz(1u << b.kappa[0]), u);
                ++hits[0][lane >> 3][at & 7u];
                ++hits[1][lane >> 3][(at ^ swz(1u << b.kappa[0])) & 7u];
            }
            for (auto& e : hits) {
                for (auto& q : e) {
                    for (int h : q) *worstConflict = std::max(*worstConflict, h);
                }
            }
        }
    }
}

} // namespace

extern "C" {

// Applies the gates (flat matrix DDs) as ONE pass to the shard `rank` of an n-qubit state held in (re, im), 2^nLocal
// amplitudes, in place.  Returns 0, or a negative code: -1 a gate is not a dense block, -2 the pass does not fit.
// info[0] = worst measured bank-conflict degree, info[1] = worst planned one, info[2] = tile bits used, info[3] = blocks' k packed,
// info[4] = the planner declared the pass warp local, info[5] = tile slots that two different compute warps touched in a warp-local pass.
int emu_apply_pass(const fdd_matdd* gates, int count, int nLocal, int rank, int tileBits, double* re, double* im, int* info) {
    std::vector<DenseBlock> blocks(static_cast<size_t>(count));
    std::vector<const DenseBlock*> ptrs;
    for (int i = 0; i < count; ++i) {
        if (!denseBlockFromDD(gates[i], blocks[static_cast<size_t>(i)])) return -1;
        padBlock(blocks[static_cast<size_t>(i)], nLocal);
        ptrs.push_back(&blocks[static_cast<size_t>(i)]);
    }
    PassParams p;
    const int need = minTileBits(ptrs.data(), count, nLocal);
    if (need < 0) return -2;
    if (!planPass(ptrs.data(), count, nLocal, rank, std::max(need, tileBits), p)) return -2;
    info[0] = 0;
    info[1] = 0;
    info[2] = p.tileBits;
    info[3] = 0;
    info[4] = static_cast<int>(p.warpLocal);
    info[5] = 0;
    for (int g = 0; g < count; ++g) {
        info[1] = std::max<int>(info[1], p.blocks[g].conflictWays);
        info[3] = info[3] * 10 + p.blocks[g].k;
    }
    const uint32_t nSegTile = 1u << (p.tileBits - kLaneBits);
    std::vector<cplx> tile(size_t{1} << p.tileBits);
    for (uint32_t t = 0; t < p.nTiles; ++t) {
        const uint32_t segBase = spreadAround(t, p.tileMask);
        for (uint32_t j = 0; j < nSegTile; ++j) {
            const uint64_t seg = segBase | pdep32(j, p.tileMask);
            for (uint32_t lane = 0; lane < 32; ++lane) tile[swz(j * 32u + lane)] = cplx(re[(seg << 5) + lane], im[(seg << 5) + lane]);
        }
        std::vector<int> owner(tile.size(), -1);
        for (int g = 0; g < count; ++g) {
            const BlockDesc& b = p.blocks[g];
            const double* table = blocks[static_cast<size_t>(g)].table.data();
            std::vector<int>* own = (p.warpLocal && b.nUnits >= 8) ? &owner : nullptr;
            if (b.k == 4) {
                applyBlockEmu<4>(b, table, tile, p.rankSegBits | segBase, &info[0], own, &info[5]);
            } else {
                applyBlockEmu<3>(b, table, tile, p.rankSegBits | segBase, &info[0], own, &info[5]);
            }
        }
        for (uint32_t j = 0; j < nSegTile; ++j) {
            const uint64_t seg = segBase | pdep32(j, p.tileMask);
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const cplx v = tile[swz(j * 32u + lane)];
                re[(seg << 5) + lane] = v.real();
                im[(seg << 5) + lane] = v.imag();
            }
        }
    }
    return 0;
}

// The dense block of a gate DD (before padding): returns k, or -1; targets/ctx get the qubit lists, table (if not null) the matrices.
int emu_block_of(const fdd_matdd* gate, int* targets, int* nCtx, int* ctx, double* table, size_t tableDoubles) {
    DenseBlock b;
    if (!denseBlockFromDD(*gate, b)) return -1;
    for (size_t i = 0; i < b.targets.size(); ++i) targets[i] = b.targets[i];
    *nCtx = static_cast<int>(b.ctx.size());
    for (size_t i = 0; i < b.ctx.size(); ++i) ctx[i] = b.ctx[i];
    if (table != nullptr) {
        if (tableDoubles < b.table.size()) return -3;
        std::memcpy(table, b.table.data(), b.table.size() * sizeof(double));
    }
    return b.k();
}

} // extern "C"

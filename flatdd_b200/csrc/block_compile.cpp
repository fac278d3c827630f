// block_compile.cpp — see block_compile.hpp.
#include "block_compile.hpp"

#include "gate_compile.hpp"

#include <algorithm>
#include <array>
#include <cstring>
#include <stdexcept>

namespace fddb200 {
namespace {

inline bool isZero(const double* w) { return w[0] == 0.0 && w[1] == 0.0; }

struct Cx {
    double re, im;
};
inline Cx mul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

} // namespace

bool denseBlockFromDD(const fdd_matdd& g, DenseBlock& out, int maxCtx) {
    validate(g);
    const std::size_t nNodes = static_cast<std::size_t>(g.n_nodes);
    auto W = [&](int32_t u, int k) { return g.weight + 2 * (4 * static_cast<std::size_t>(u) + static_cast<std::size_t>(k)); };
    auto C = [&](int32_t u, int k) { return g.child[4 * static_cast<std::size_t>(u) + static_cast<std::size_t>(k)]; };
    // classify the levels over the reachable nodes: 2 = off-diagonal successor somewhere (target), 1 = diagonal but not
    // identity-like somewhere (context), 0 = [a 0; 0 a] with one successor everywhere (the level does not matter)
    std::vector<uint8_t> reach(nNodes, 0);
    std::vector<uint8_t> levelClass(static_cast<std::size_t>(g.n_qubits), 0);
    if (!isZero(g.root_weight)) {
        std::vector<int32_t> stack{g.root};
        reach[static_cast<std::size_t>(g.root)] = 1;
        while (!stack.empty()) {
            const int32_t u = stack.back();
            stack.pop_back();
            const bool offDiag = !isZero(W(u, 1)) || !isZero(W(u, 2));
            const bool ident = !offDiag && !isZero(W(u, 0)) && W(u, 0)[0] == W(u, 3)[0] && W(u, 0)[1] == W(u, 3)[1] && C(u, 0) == C(u, 3);
            auto& cls = levelClass[static_cast<std::size_t>(g.level[u])];
            cls = std::max<uint8_t>(cls, offDiag ? 2 : (ident ? 0 : 1));
            for (int k = 0; k < 4; ++k) {
                const int32_t c = C(u, k);
                if (!isZero(W(u, k)) && c >= 0 && !reach[static_cast<std::size_t>(c)]) {
                    reach[static_cast<std::size_t>(c)] = 1;
                    stack.push_back(c);
                }
            }
        }
    }
    out = DenseBlock{};
    out.n = g.n_qubits;
    std::vector<int> role(static_cast<std::size_t>(g.n_qubits), -1); // index among the targets / context qubits
    for (int v = 0; v < g.n_qubits; ++v) {
        if (levelClass[static_cast<std::size_t>(v)] == 2) {
            role[static_cast<std::size_t>(v)] = static_cast<int>(out.targets.size());
            out.targets.push_back(v);
        } else if (levelClass[static_cast<std::size_t>(v)] == 1) {
            role[static_cast<std::size_t>(v)] = static_cast<int>(out.ctx.size());
            out.ctx.push_back(v);
        }
    }
    if (out.k() > kBlockMaxTargets || static_cast<int>(out.ctx.size()) > maxCtx) return false;
    const std::size_t rows = out.rows();
    out.table.assign((std::size_t{2} * rows * rows) << out.ctx.size(), 0.0);
    if (isZero(g.root_weight)) return true;
    // Below its lowest target / context level a path is a chain of identity-like nodes.  Where every weight on that chain is
    // exactly 1 the walk can stop at the top of the chain: multiplying by 1 changes nothing, whatever the order (normalised
    // gate DDs are of this kind; the chain is 20 of the 26 levels of a supremacy_n26 block and was most of the frames).
    std::vector<uint8_t> tailOne(nNodes, 0);
    {
        std::vector<int32_t> order; // reachable nodes by ascending level: successors first
        for (std::size_t u = 0; u < nNodes; ++u) {
            if (reach[u]) order.push_back(static_cast<int32_t>(u));
        }
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return g.level[a] < g.level[b]; });
        for (int32_t u : order) {
            const int v = g.level[u];
            if (levelClass[static_cast<std::size_t>(v)] != 0 || W(u, 0)[0] != 1.0 || W(u, 0)[1] != 0.0) continue;
            tailOne[static_cast<std::size_t>(u)] = (v == 0 || (C(u, 0) >= 0 && tailOne[static_cast<std::size_t>(C(u, 0))])) ? 1 : 0;
        }
    }
    // depth-first over the non-zero paths; every (context, row, column) is one path
    struct Frame {
        int32_t node;
        Cx w;
        uint32_t row, col, ctx;
    };
    std::vector<Frame> work;
    work.push_back({g.root, Cx{g.root_weight[0], g.root_weight[1]}, 0, 0, 0});
    while (!work.empty()) {
        const Frame f = work.back();
        work.pop_back();
        if (f.node == FDD_TERMINAL || tailOne[static_cast<std::size_t>(f.node)]) {
            const std::size_t at = 2 * ((static_cast<std::size_t>(f.ctx) * rows + f.row) * rows + f.col);
            out.table[at] = f.w.re;
            out.table[at + 1] = f.w.im;
            continue;
        }
        const int v = g.level[f.node];
        const uint8_t cls = levelClass[static_cast<std::size_t>(v)];
        const int r = role[static_cast<std::size_t>(v)];
        auto follow = [&](int k, uint32_t row, uint32_t col, uint32_t ctx) {
            if (isZero(W(f.node, k))) return;
            work.push_back({v == 0 ? FDD_TERMINAL : C(f.node, k), mul(f.w, Cx{W(f.node, k)[0], W(f.node, k)[1]}), row, col, ctx});
        };
        if (cls == 2) {
            for (int rb = 0; rb < 2; ++rb) {
                for (int cb = 0; cb < 2; ++cb) follow(2 * rb + cb, f.row | (static_cast<uint32_t>(rb) << r), f.col | (static_cast<uint32_t>(cb) << r), f.ctx);
            }
        } else if (cls == 1) {
            follow(0, f.row, f.col, f.ctx);
            follow(3, f.row, f.col, f.ctx | (1u << r));
        } else {
            follow(0, f.row, f.col, f.ctx);
        }
    }
    return true;
}

namespace {

// q becomes a target.  fromCtx: q was a context qubit (the block is block diagonal in q); otherwise an identity factor.
void addTarget(DenseBlock& b, int q, bool fromCtx) {
    const std::size_t oldRows = b.rows();
    const std::size_t nCtxOld = b.ctx.size();
    int cj = -1;
    if (fromCtx) {
        cj = static_cast<int>(std::find(b.ctx.begin(), b.ctx.end(), q) - b.ctx.begin());
    }
    std::vector<int> newTargets = b.targets;
    newTargets.insert(std::upper_bound(newTargets.begin(), newTargets.end(), q), q);
    const int tq = static_cast<int>(std::find(newTargets.begin(), newTargets.end(), q) - newTargets.begin());
    std::vector<int> newCtx = b.ctx;
    if (fromCtx) newCtx.erase(newCtx.begin() + cj);
    const std::size_t newRows = oldRows * 2;
    std::vector<double> table((std::size_t{2} * newRows * newRows) << newCtx.size(), 0.0);
    auto dropBit = [](std::size_t x, int at) { return ((x >> (at + 1)) << at) | (x & ((std::size_t{1} << at) - 1)); };
    auto insertBit = [](std::size_t x, int at, std::size_t bit) { return ((x >> at) << (at + 1)) | (bit << at) | (x & ((std::size_t{1} << at) - 1)); };
    for (std::size_t c = 0; c < (std::size_t{1} << newCtx.size()); ++c) {
        for (std::size_t r2 = 0; r2 < newRows; ++r2) {
            for (std::size_t c2 = 0; c2 < newRows; ++c2) {
                const std::size_t rb = (r2 >> tq) & 1U, cb = (c2 >> tq) & 1U;
                if (rb != cb) continue;
                const std::size_t oldCtx = fromCtx ? insertBit(c, cj, rb) : c;
                const std::size_t src = 2 * ((oldCtx * oldRows + dropBit(r2, tq)) * oldRows + dropBit(c2, tq));
                const std::size_t dst = 2 * ((c * newRows + r2) * newRows + c2);
                table[dst] = b.table[src];
                table[dst + 1] = b.table[src + 1];
            }
        }
    }
    (void)nCtxOld;
    b.targets.swap(newTargets);
    b.ctx.swap(newCtx);
    b.table.swap(table);
}

} // namespace

void padBlock(DenseBlock& b, int nLocal) {
    auto isTarget = [&](int q) { return std::find(b.targets.begin(), b.targets.end(), q) != b.targets.end(); };
    // context qubits among the warp-lane bits cost nothing as targets (they are in every tile): promote them while there is room
    for (;;) {
        if (b.k() >= kBlockMaxTargets) break;
        int pick = -1;
        for (int q : b.ctx) {
            if (q < kLaneBits && q < nLocal) {
                pick = q;
                break;
            }
        }
        if (pick < 0) break;
        addTarget(b, pick, true);
    }
    // the kernel runs 8 x 8 and 16 x 16 blocks: pad with the lowest local context qubit, else with an identity factor on the
    // lowest free lane bit
    while (b.k() < 3) {
        int pick = -1;
        for (int q : b.ctx) {
            if (q < nLocal) {
                pick = q;
                break;
            }
        }
        if (pick >= 0) {
            addTarget(b, pick, true);
            continue;
        }
        for (int q = 0; q < nLocal; ++q) {
            if (!isTarget(q)) {
                pick = q;
                break;
            }
        }
        if (pick < 0) throw std::logic_error("padBlock: no free local qubit");
        addTarget(b, pick, false);
    }
}

int minTileBits(const DenseBlock* const* blocks, int count, int nLocal) {
    uint64_t upper = 0;
    for (int i = 0; i < count; ++i) {
        for (int q : blocks[i]->targets) {
            if (q >= nLocal) return -1;
            if (q >= kLaneBits) upper |= uint64_t{1} << q;
        }
    }
    return kLaneBits + __builtin_popcountll(upper);
}

bool planPass(const DenseBlock* const* blocks, int count, int nLocal, int rank, int tileBits, PassParams& pass) {
    if (count < 1 || count > kPassMaxBlocks) return false;
    tileBits = std::min(tileBits, std::min(nLocal, kPassMaxTileBits));
    const int need = minTileBits(blocks, count, nLocal);
    if (need < 0 || need > tileBits || tileBits < kLaneBits + 1) return false;
    uint32_t tileMask = 0; // over segment-index bits
    for (int i = 0; i < count; ++i) {
        for (int q : blocks[i]->targets) {
            if (q >= kLaneBits) tileMask |= 1u << (q - kLaneBits);
        }
    }
    for (int sb = 0; sb < nLocal - kLaneBits && __builtin_popcount(tileMask) < tileBits - kLaneBits; ++sb) tileMask |= 1u << sb;
    if (kLaneBits + __builtin_popcount(tileMask) != tileBits) return false;
    auto tilePos = [&](int q) -> int { // tile-local position of a local qubit, -1 when it is outside the tile
        if (q < kLaneBits) return q;
        if (q >= nLocal || !((tileMask >> (q - kLaneBits)) & 1u)) return -1;
        return kLaneBits + __builtin_popcount(tileMask & ((1u << (q - kLaneBits)) - 1u));
    };
    std::memset(&pass, 0, sizeof pass);
    for (uint32_t& off : pass.tableSmem) off = kTableInGlobal;
    pass.tileBits = tileBits;
    pass.nBlocks = count;
    pass.tileMask = tileMask;
    pass.nTiles = 1u << (nLocal - tileBits);
    pass.rankSegBits = static_cast<uint32_t>(rank) << (nLocal - kLaneBits);
    // Warp-local passes: three tile bits outside every block's targets become the top unit bits of every block (see
    // PassParams::warpLocal).  Tried first; when some block then lacks free bits for its fragment columns, plan without.
    uint32_t allTargets = 0;
    for (int i = 0; i < count; ++i) {
        for (int q : blocks[i]->targets) {
            const int pos = tilePos(q);
            if (pos < 0) return false;
            allTargets |= 1u << pos;
        }
    }
    auto planBlocks = [&](uint32_t warpBits) -> bool {
    for (int i = 0; i < count; ++i) {
            const DenseBlock& blk = *blocks[i];
            BlockDesc& d = pass.blocks[i];
            const int k = blk.k();
            if (k != 3 && k != 4) return false;
            if (static_cast<int>(blk.ctx.size()) > kBlockMaxCtx) return false;
            d.k = k;
            d.nCtx = static_cast<uint8_t>(blk.ctx.size());
            uint32_t used = 0; // tile bits that cannot be fragment columns: targets and in-tile context bits
            std::array<int, 4> tp{};
            for (int t = 0; t < k; ++t) {
                tp[static_cast<std::size_t>(t)] = tilePos(blk.targets[static_cast<std::size_t>(t)]);
                if (tp[static_cast<std::size_t>(t)] < 0) return false;
                used |= 1u << tp[static_cast<std::size_t>(t)];
            }
            uint32_t ctxInTile = 0;
            for (std::size_t j = 0; j < blk.ctx.size(); ++j) {
                const int q = blk.ctx[j];
                const int pos = tilePos(q);
                if (pos >= 0) {
                    d.ctxSrc[j] = static_cast<uint8_t>(pos);
                    ctxInTile |= 1u << pos;
                } else {
                    d.ctxSrc[j] = static_cast<uint8_t>(32 + (q - kLaneBits));
                }
            }
            used |= ctxInTile;
            std::vector<int> cand;
            for (int pos = 0; pos < tileBits; ++pos) {
                if (!((used >> pos) & 1u) && !((warpBits >> pos) & 1u)) cand.push_back(pos);
            }
            if (cand.size() < 3) return false;
            // fragment shape with the fewest bank conflicts: the quarter-warps of a B load differ in (sigma0, sigma1, kappa0),
            // those of a D store in (kappa1, kappa2, sigma0)
            auto ways = [](int a, int b, int c) {
                int count8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int worst = 0;
                for (uint32_t m = 0; m < 8; ++m) {
                    const uint32_t t = ((m & 1u) << a) | (((m >> 1) & 1u) << b) | (((m >> 2) & 1u) << c);
                    worst = std::max(worst, ++count8[swz(t) & 7u]);
                }
                return worst;
            };
            int best = 1 << 30;
            int bs0 = 0, bs1 = 1, bk0 = 0, bk1 = 1, bk2 = 2;
            for (int s0 = 0; s0 < k; ++s0) {
                for (int s1 = 0; s1 < k; ++s1) {
                    if (s1 == s0) continue;
                    for (std::size_t a = 0; a < cand.size(); ++a) {
                        const int wB = ways(tp[static_cast<std::size_t>(s0)], tp[static_cast<std::size_t>(s1)], cand[a]);
                        if (wB * 16 >= best) continue;
                        for (std::size_t b1 = 0; b1 < cand.size(); ++b1) {
                            if (b1 == a) continue;
                            for (std::size_t b2 = b1 + 1; b2 < cand.size(); ++b2) { // (kappa1, kappa2) is symmetric for the conflict count
                                if (b2 == a) continue;
                                const int wD = ways(cand[b1], cand[b2], tp[static_cast<std::size_t>(s0)]);
                                const int score = wB * 16 + wD;
                                if (score < best) {
                                    best = score;
                                    bs0 = s0;
                                    bs1 = s1;
                                    bk0 = cand[a];
                                    bk1 = cand[b1];
                                    bk2 = cand[b2];
                                }
                            }
                        }
                    }
                }
            }
            d.conflictWays = static_cast<uint8_t>(std::max(best / 16, best % 16));
            // sigma order: the two chosen targets first, the others ascending
            std::array<int, 4> order{};
            order[0] = bs0;
            order[1] = bs1;
            int at = 2;
            for (int t = 0; t < k; ++t) {
                if (t != bs0 && t != bs1) order[static_cast<std::size_t>(at++)] = t;
            }
            for (int i2 = 0; i2 < k; ++i2) d.sigma[i2] = static_cast<uint8_t>(tp[static_cast<std::size_t>(order[static_cast<std::size_t>(i2)])]);
            for (int s = 0; s < (1 << k); ++s) {
                int canon = 0;
                for (int i2 = 0; i2 < k; ++i2) {
                    if ((s >> i2) & 1) canon |= 1 << order[static_cast<std::size_t>(i2)];
                }
                d.canon[s] = static_cast<uint8_t>(canon);
            }
            d.kappa[0] = static_cast<uint8_t>(bk0);
            d.kappa[1] = static_cast<uint8_t>(bk1);
            d.kappa[2] = static_cast<uint8_t>(bk2);
            // unit bits: every other tile bit, context bits last
            int nu = 0;
            const uint32_t taken = used | (1u << bk0) | (1u << bk1) | (1u << bk2);
            for (int pos = 0; pos < tileBits; ++pos) {
                if (!((taken >> pos) & 1u) && !((warpBits >> pos) & 1u)) d.unitPos[nu++] = static_cast<uint8_t>(pos);
            }
            for (int pos = 0; pos < tileBits; ++pos) {
                if (((ctxInTile >> pos) & 1u) && !((warpBits >> pos) & 1u)) d.unitPos[nu++] = static_cast<uint8_t>(pos);
            }
            for (int pos = 0; pos < tileBits; ++pos) { // the warp bits on top, in the same order for every block
                if ((warpBits >> pos) & 1u) d.unitPos[nu++] = static_cast<uint8_t>(pos);
            }
            if (nu != tileBits - k - 3) return false;
            d.unitBit0IsCtx = (nu > 0 && ((ctxInTile >> d.unitPos[0]) & 1u)) ? 1 : 0;
            d.nUnitBits = static_cast<uint8_t>(nu);
            d.nUnits = 1 << nu;
        }
        return true;
    };
    uint32_t warpBits = 0;
    if (count > 1) {
        int have = 0;
        for (int pos = tileBits - 1; pos >= 0 && have < 3; --pos) {
            if (!((allTargets >> pos) & 1u)) {
                warpBits |= 1u << pos;
                ++have;
            }
        }
        if (have < 3) warpBits = 0;
    }
    if (warpBits != 0 && planBlocks(warpBits)) {
        pass.warpLocal = 1;
        return true;
    }
    pass.warpLocal = 0;
    return planBlocks(0);
}

} // namespace fddb200
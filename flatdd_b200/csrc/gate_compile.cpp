// gate_compile.cpp — see gate_compile.hpp.
#include "gate_compile.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <utility>

namespace fddb200 {
namespace {

struct Cx {
    double re, im;
};
inline Cx mul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline bool isZero(const double* w) { return w[0] == 0.0 && w[1] == 0.0; }

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error(what); }

template <int R, class DD> void validateImpl(const DD& g, const char* name) {
    if (g.n_qubits < 1 || g.n_qubits > 40) bad(std::string(name) + ": n_qubits out of range");
    if (g.n_nodes < 1) bad(std::string(name) + ": empty node table");
    if (g.level == nullptr || g.child == nullptr || g.weight == nullptr) bad(std::string(name) + ": null table pointer");
    if (g.root < 0 || g.root >= g.n_nodes) bad(std::string(name) + ": root index out of range");
    if (g.level[g.root] != g.n_qubits - 1) bad(std::string(name) + ": root node is not at level n_qubits-1");
    for (int32_t u = 0; u < g.n_nodes; ++u) {
        const int32_t lv = g.level[u];
        if (lv < 0 || lv >= g.n_qubits) bad(std::string(name) + ": node level out of range");
        for (int k = 0; k < R; ++k) {
            const std::size_t e = static_cast<std::size_t>(u) * R + k;
            if (isZero(g.weight + 2 * e)) continue; // zero edge: successor ignored
            if (!std::isfinite(g.weight[2 * e]) || !std::isfinite(g.weight[2 * e + 1])) bad(std::string(name) + ": non-finite edge weight");
            const int32_t c = g.child[e];
            if (lv == 0) {
                if (c != FDD_TERMINAL) bad(std::string(name) + ": level-0 node must point to the terminal");
            } else {
                if (c < 0 || c >= g.n_nodes) bad(std::string(name) + ": successor index out of range");
                if (g.level[c] != lv - 1) bad(std::string(name) + ": successor is not exactly one level below its parent");
            }
        }
    }
}

// memoised nnz per node (terminal counts 1)
uint64_t macRec(const fdd_matdd& g, int32_t node, std::vector<uint64_t>& memo, std::vector<uint8_t>& have) {
    if (node == FDD_TERMINAL) return 1;
    if (have[static_cast<std::size_t>(node)]) return memo[static_cast<std::size_t>(node)];
    uint64_t cnt = 0;
    for (int k = 0; k < 4; ++k) {
        const std::size_t e = 4 * static_cast<std::size_t>(node) + k;
        if (!isZero(g.weight + 2 * e)) cnt += macRec(g, g.child[e], memo, have);
    }
    memo[static_cast<std::size_t>(node)] = cnt;
    have[static_cast<std::size_t>(node)] = 1;
    return cnt;
}

} // namespace

void validate(const fdd_matdd& g) { validateImpl<4>(g, "fdd_matdd"); }
void validate(const fdd_vecdd& v) { validateImpl<2>(v, "fdd_vecdd"); }

uint64_t macCount(const fdd_matdd& g) {
    std::vector<uint64_t> memo(static_cast<std::size_t>(g.n_nodes), 0);
    std::vector<uint8_t> have(static_cast<std::size_t>(g.n_nodes), 0);
    return macRec(g, g.root, memo, have);
}

uint64_t costIP(const fdd_matdd& g, unsigned nThreadExp) { return macCount(g) >> nThreadExp; }

// Column-partitioned cost with the DMAV cache (reference DMAVMACCountOP1,
// include/dd/SwitchPackage.hpp:3202-3283; partition walk AssignParVectorOP :2472-2508).
uint64_t costOP1(const fdd_matdd& g, unsigned nThreadExp) {
    const uint64_t nDim = uint64_t{1} << g.n_qubits;
    const std::size_t nThread = std::size_t{1} << nThreadExp;
    const uint64_t seg = nDim / nThread;
    struct Block {
        int32_t node;
        uint64_t rowOff;
    };
    std::vector<std::vector<Block>> perColumn(nThread);
    // depth-first over the top nThreadExp levels: column bit outer, row bit inner
    struct Frame {
        int32_t node;
        unsigned depth;
        std::size_t colBlock;
        uint64_t rowOff;
    };
    std::vector<Frame> work;
    if (!isZero(g.root_weight)) work.push_back({g.root, 0, 0, 0});
    const int top = g.n_qubits - 1;
    while (!work.empty()) {
        const Frame f = work.back();
        work.pop_back();
        if (f.depth == nThreadExp) {
            perColumn[f.colBlock].push_back({f.node, f.rowOff});
            continue;
        }
        // push in reverse so blocks are appended in the reference's recursion order
        for (int i = 1; i >= 0; --i) {
            for (int j = 1; j >= 0; --j) {
                const std::size_t e = 4 * static_cast<std::size_t>(f.node) + static_cast<std::size_t>(j * 2 + i);
                if (isZero(g.weight + 2 * e)) continue;
                work.push_back({g.child[e], f.depth + 1,
                                f.colBlock + static_cast<std::size_t>(i) * (std::size_t{1} << (nThreadExp - f.depth - 1)),
                                f.rowOff + (uint64_t{1} << (top - static_cast<int>(f.depth))) * static_cast<uint64_t>(j)});
            }
        }
    }
    // greedy first-fit of column blocks into scratch buffers with disjoint claimed output ranges;
    // the reference iterates over the stored (start, end) offsets alike and keeps them as int
    std::vector<std::vector<std::pair<int, int>>> buffers;
    for (std::size_t i = 0; i < nThread; ++i) {
        std::vector<uint64_t> offs;
        for (const Block& b : perColumn[i]) {
            offs.push_back(b.rowOff);
            offs.push_back(b.rowOff + seg);
        }
        bool placed = false;
        for (auto& buf : buffers) {
            bool clash = false;
            for (const auto& claimed : buf) {
                for (uint64_t o : offs) {
                    if (static_cast<uint64_t>(static_cast<int64_t>(claimed.first)) < o + seg &&
                        static_cast<uint64_t>(static_cast<int64_t>(claimed.second)) > o) {
                        clash = true;
                        break;
                    }
                }
                if (clash) break;
            }
            if (!clash) {
                for (uint64_t o : offs) buf.emplace_back(static_cast<int>(o), static_cast<int>(o + seg));
                placed = true;
                break;
            }
        }
        if (!placed) {
            buffers.emplace_back();
            for (uint64_t o : offs) buffers.back().emplace_back(static_cast<int>(o), static_cast<int>(o + seg));
        }
    }
    std::vector<uint64_t> memo(static_cast<std::size_t>(g.n_nodes), 0);
    std::vector<uint8_t> have(static_cast<std::size_t>(g.n_nodes), 0);
    uint64_t cnt = 0;
    for (std::size_t i = 0; i < nThread; ++i) {
        std::set<int32_t> seen;
        for (const Block& b : perColumn[i]) {
            if (!seen.insert(b.node).second) {
                cnt += seg / 4;
            } else if (b.node != FDD_TERMINAL) {
                cnt += macRec(g, b.node, memo, have);
            }
        }
    }
    return cnt / nThread + nDim * buffers.size() / (4 * nThread);
}

CompiledGate compileGate(const fdd_matdd& g, int nLocal) {
    validate(g);
    if (nLocal < 0 || nLocal > g.n_qubits) nLocal = g.n_qubits;
    CompiledGate out;
    out.n = g.n_qubits;
    out.segBits = std::min(5, g.n_qubits);
    const int S = out.segBits;
    const std::size_t nNodes = static_cast<std::size_t>(g.n_nodes);
    auto W = [&](int32_t u, int k) { return g.weight + 2 * (4 * static_cast<std::size_t>(u) + static_cast<std::size_t>(k)); };
    auto C = [&](int32_t u, int k) { return g.child[4 * static_cast<std::size_t>(u) + static_cast<std::size_t>(k)]; };

    // reachability (zero edges cut) and per-node structure flags
    std::vector<uint8_t> reach(nNodes, 0), identLike(nNodes, 0);
    {
        std::vector<int32_t> stack{g.root};
        reach[static_cast<std::size_t>(g.root)] = 1;
        while (!stack.empty()) {
            const int32_t u = stack.back();
            stack.pop_back();
            for (int k = 0; k < 4; ++k) {
                const int32_t c = C(u, k);
                if (!isZero(W(u, k)) && c >= 0 && !reach[static_cast<std::size_t>(c)]) {
                    reach[static_cast<std::size_t>(c)] = 1;
                    stack.push_back(c);
                }
            }
        }
    }
    out.diagonal = true;
    for (int32_t u = 0; u < g.n_nodes; ++u) {
        if (!reach[static_cast<std::size_t>(u)]) continue;
        const bool offDiagZero = isZero(W(u, 1)) && isZero(W(u, 2));
        if (!offDiagZero) out.diagonal = false;
        const bool ident = offDiagZero && !isZero(W(u, 0)) && W(u, 0)[0] == W(u, 3)[0] && W(u, 0)[1] == W(u, 3)[1] &&
                           C(u, 0) == C(u, 3);
        identLike[static_cast<std::size_t>(u)] = ident ? 1 : 0;
        if (!ident) out.topLevel = std::max(out.topLevel, g.level[u]);
    }

    // ---- sub tables: ELL expansion of level S-1 nodes -------------------------------------------
    using Row = std::vector<std::pair<int, Cx>>;
    std::map<int32_t, std::vector<Row>> expanded; // node -> rows (2^(level+1) of them)
    // iterative post-order would be overkill: depth <= 5
    struct Expander {
        const fdd_matdd& g;
        std::map<int32_t, std::vector<Row>>& memo;
        const std::vector<Row>& rows(int32_t u) {
            auto it = memo.find(u);
            if (it != memo.end()) return it->second;
            const int lv = g.level[u];
            std::vector<Row> r(std::size_t{1} << (lv + 1));
            for (std::size_t row = 0; row < r.size(); ++row) {
                const int rb = static_cast<int>((row >> lv) & 1U);
                const std::size_t low = row & ((std::size_t{1} << lv) - 1);
                for (int cb = 0; cb < 2; ++cb) {
                    const std::size_t e = 4 * static_cast<std::size_t>(u) + static_cast<std::size_t>(2 * rb + cb);
                    const double* w = g.weight + 2 * e;
                    if (isZero(w)) continue;
                    const Cx we{w[0], w[1]};
                    if (lv == 0) {
                        r[row].emplace_back(cb, we);
                    } else {
                        for (const auto& ent : rows(g.child[e])[low]) {
                            r[row].emplace_back(ent.first | (cb << lv), mul(we, ent.second));
                        }
                    }
                }
            }
            return memo.emplace(u, std::move(r)).first->second;
        }
    } expander{g, expanded};

    std::map<int32_t, int32_t> subId; // level S-1 node -> sub table id
    auto subFor = [&](int32_t node) {
        auto it = subId.find(node);
        if (it != subId.end()) return it->second;
        const auto id = static_cast<int32_t>(subId.size());
        subId.emplace(node, id);
        return id;
    };

    // ---- upper nodes with identity compression ----------------------------------------------------
    std::map<int32_t, int32_t> upperId;
    std::vector<int32_t> upperOrder; // original node index per upper slot
    struct Resolved {
        int32_t code;
        Cx w;
    };
    auto resolve = [&](int32_t c, Cx w) -> Resolved {
        while (g.level[c] >= S && identLike[static_cast<std::size_t>(c)]) {
            w = mul(w, Cx{W(c, 0)[0], W(c, 0)[1]});
            c = C(c, 0);
        }
        if (g.level[c] < S) return {encodeSub(subFor(c)), w};
        auto it = upperId.find(c);
        if (it == upperId.end()) {
            it = upperId.emplace(c, static_cast<int32_t>(upperOrder.size())).first;
            upperOrder.push_back(c);
        }
        return {it->second, w};
    };
    {
        const Resolved r = resolve(g.root, Cx{g.root_weight[0], g.root_weight[1]});
        out.root = r.code;
        out.rootW[0] = r.w.re;
        out.rootW[1] = r.w.im;
        if (isZero(g.root_weight)) out.root = FDD_TERMINAL;
    }
    for (std::size_t slot = 0; slot < upperOrder.size(); ++slot) { // upperOrder grows while we scan
        const int32_t u = upperOrder[slot];
        UpperNode nd{};
        nd.level = g.level[u];
        for (int k = 0; k < 4; ++k) {
            if (isZero(W(u, k))) {
                nd.child[k] = FDD_TERMINAL;
                continue;
            }
            const Resolved r = resolve(C(u, k), Cx{W(u, k)[0], W(u, k)[1]});
            nd.child[k] = r.code;
            nd.w[2 * k] = r.w.re;
            nd.w[2 * k + 1] = r.w.im;
        }
        if (out.upper.size() <= slot) out.upper.resize(slot + 1);
        out.upper[slot] = nd;
    }

    // ---- materialise the sub tables ------------------------------------------------------------------
    out.nSub = static_cast<int>(subId.size());
    const int rowsPerSeg = 1 << S;
    std::vector<const std::vector<Row>*> subRows(static_cast<std::size_t>(out.nSub), nullptr);
    for (const auto& kv : subId) subRows[static_cast<std::size_t>(kv.second)] = &expander.rows(kv.first);
    out.subK.assign(static_cast<std::size_t>(out.nSub), 1);
    out.kMax = 1;
    for (int s = 0; s < out.nSub; ++s) {
        int k = 1;
        for (const Row& r : *subRows[static_cast<std::size_t>(s)]) k = std::max<int>(k, static_cast<int>(r.size()));
        out.subK[static_cast<std::size_t>(s)] = k;
        out.kMax = std::max(out.kMax, k);
    }
    out.kTrue = out.kMax;
    // pad the ELL width to 2, 4, 8 or 16 so the kernels can unroll it without a bound check
    if (out.kMax <= 16) out.kMax = out.kMax <= 2 ? 2 : (out.kMax <= 4 ? 4 : (out.kMax <= 8 ? 8 : 16));
    out.subCol.assign(static_cast<std::size_t>(out.nSub) * out.kMax * 32, 0);
    out.subW.assign(static_cast<std::size_t>(out.nSub) * out.kMax * 32 * 2, 0.0);
    out.subFlags.assign(static_cast<std::size_t>(out.nSub), 0);
    for (int s = 0; s < out.nSub; ++s) {
        const auto& rows = *subRows[static_cast<std::size_t>(s)];
        bool diag = out.subK[static_cast<std::size_t>(s)] == 1;
        bool ident = diag;
        for (int row = 0; row < 32; ++row) {
            for (int k = 0; k < out.kMax; ++k) {
                const std::size_t at = (static_cast<std::size_t>(s) * out.kMax + static_cast<std::size_t>(k)) * 32 + static_cast<std::size_t>(row);
                int col = row < rowsPerSeg ? row : 0;
                Cx w{0.0, 0.0};
                if (row < rowsPerSeg && k < static_cast<int>(rows[static_cast<std::size_t>(row)].size())) {
                    col = rows[static_cast<std::size_t>(row)][static_cast<std::size_t>(k)].first;
                    w = rows[static_cast<std::size_t>(row)][static_cast<std::size_t>(k)].second;
                }
                out.subCol[at] = static_cast<uint8_t>(col);
                out.subW[2 * at] = w.re;
                out.subW[2 * at + 1] = w.im;
                if (row < rowsPerSeg && k == 0) {
                    if (col != row) diag = ident = false;
                    if (w.re != 1.0 || w.im != 0.0) ident = false;
                }
            }
        }
        out.subFlags[static_cast<std::size_t>(s)] = static_cast<uint8_t>((ident ? SUB_IDENTITY : 0) | (diag ? SUB_DIAGONAL : 0));
    }

    // ---- facts --------------------------------------------------------------------------------------
    {
        const std::size_t nu = out.upper.size();
        std::vector<int> paths(nu, 0), stackNeed(nu, 0), depth(nu, 0);
        // children always have a lower level: process by ascending level
        std::vector<std::size_t> order(nu);
        for (std::size_t i = 0; i < nu; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return out.upper[a].level < out.upper[b].level; });
        auto P = [&](int32_t code) { return code >= 0 ? paths[static_cast<std::size_t>(code)] : (code == FDD_TERMINAL ? 0 : 1); };
        auto St = [&](int32_t code) { return code >= 0 ? stackNeed[static_cast<std::size_t>(code)] : 0; };
        auto Dp = [&](int32_t code) { return code >= 0 ? depth[static_cast<std::size_t>(code)] : 0; };
        for (std::size_t idx : order) {
            const UpperNode& nd = out.upper[idx];
            int best = 0, st = 0, dp = 0;
            for (int rb = 0; rb < 2; ++rb) {
                const int32_t c0 = nd.child[2 * rb], c1 = nd.child[2 * rb + 1];
                best = std::max(best, P(c0) + P(c1));
                if (c0 != FDD_TERMINAL && c1 != FDD_TERMINAL) {
                    st = std::max(st, std::max(1 + St(c0), St(c1)));
                } else {
                    st = std::max(st, std::max(St(c0), St(c1)));
                }
                dp = std::max(dp, 1 + std::max(Dp(c0), Dp(c1)));
            }
            paths[idx] = std::min(best, 1 << 20);
            stackNeed[idx] = st;
            depth[idx] = dp;
        }
        out.maxPaths = std::max(1, P(out.root));
        // the same bound per sub table (paths of one row that end in table s)
        out.subPaths.assign(static_cast<std::size_t>(out.nSub), 0);
        for (int sIdx = 0; sIdx < out.nSub; ++sIdx) {
            std::vector<int> ps(nu, 0);
            auto Ps = [&](int32_t code) { return code >= 0 ? ps[static_cast<std::size_t>(code)] : (code == encodeSub(sIdx) ? 1 : 0); };
            for (std::size_t idx : order) {
                const UpperNode& nd = out.upper[idx];
                ps[idx] = std::min(1 << 20, std::max(Ps(nd.child[0]) + Ps(nd.child[1]), Ps(nd.child[2]) + Ps(nd.child[3])));
            }
            out.subPaths[static_cast<std::size_t>(sIdx)] = Ps(out.root);
        }
        out.stackCap = std::max(1, St(out.root));
        out.upperDepth = Dp(out.root);
    }
    out.allIdentitySubs = true;
    for (uint8_t f : out.subFlags) out.allIdentitySubs = out.allIdentitySubs && (f & SUB_IDENTITY);
    out.nnz = macCount(g);
    out.nnzRowMax = out.maxPaths * out.kTrue;

    // ---- tile bits -----------------------------------------------------------------------------------
    {
        uint64_t nonDiag = 0; // bit (level - S) set when some reachable node of that level is off-diagonal
        for (int32_t u = 0; u < g.n_nodes; ++u) {
            if (!reach[static_cast<std::size_t>(u)] || g.level[u] < S) continue;
            if (!isZero(W(u, 1)) || !isZero(W(u, 2))) nonDiag |= uint64_t{1} << (g.level[u] - S);
        }
        out.nonDiagUpper = __builtin_popcountll(nonDiag);
        for (int32_t u = 0; u < g.n_nodes; ++u) {
            if (reach[static_cast<std::size_t>(u)] && (!isZero(W(u, 1)) || !isZero(W(u, 2)))) out.nonDiagMask |= uint64_t{1} << g.level[u];
        }
        const int localSegBits = std::max(0, nLocal - S);
        out.tileBits = std::min(5, localSegBits);
        const bool allLocal = (nonDiag >> localSegBits) == 0;
        out.tileable = allLocal && out.nonDiagUpper <= out.tileBits;
        if (out.tileable) {
            const uint32_t mask = static_cast<uint32_t>(nonDiag);
            uint32_t fill = 0;
            for (int b = 0; b < localSegBits && __builtin_popcount(mask | fill) < out.tileBits; ++b) {
                if (!((mask >> b) & 1u)) fill |= 1u << b;
            }
            out.tileMask = mask;
            out.fillMask = fill;
            out.subTileBits = out.nonDiagUpper;
            out.uniform = true;
            for (UpperNode& nd : out.upper) {
                const int sh = nd.level - S;
                nd.slotBit = (sh < 32 && ((mask >> sh) & 1u)) ? __builtin_popcount(mask & ((1u << sh) - 1u)) : -1;
                if (nd.slotBit < 0) {
                    out.uniform = false;
                    if (sh < localSegBits) out.ctxMask |= 1u << sh; // levels above the shard are constant per rank
                }
            }
        } else {
            for (UpperNode& nd : out.upper) nd.slotBit = -1;
        }
    }
    return out;
}

namespace {
// prices in HBM passes; FLATDD_B200_COST="base,perSegment,crossLane,tensor8,tensor16,walkPerTile" overrides them (experiments)
struct CostKnobs {
    double base = 1.2, perSegment = 0.025, crossLane = 0.035, tensor8 = 1.22, tensor16 = 1.29, walkPerTile = 0.7;
    static CostKnobs fromEnv() {
        CostKnobs k;
        if (const char* e = std::getenv("FLATDD_B200_COST")) {
            std::sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf", &k.base, &k.perSegment, &k.crossLane, &k.tensor8, &k.tensor16, &k.walkPerTile);
        }
        return k;
    }
};
} // namespace

double costGpuNs(const CompiledGate& c, double hbmGBs, double fp64GFlops) {
    // Time of one launch as a multiple of the HBM time of one read + one write pass (0.33 ms at n = 26).
    // Two regimes, both fitted to the tile kernel on B200 (per-gate device times of round 1: profiles/r01_per_gate_supremacy_n26.csv, profiles/r01b_per_gate_supremacy_n26_fuse4.csv):
    //  * KNOWN-FAST classes — the block does not depend on qubits outside its tile ("uniform"), its low
    //    levels are untouched or one sub table serves the whole gate, and its upper part is either small
    //    (<= 4 sources per segment) or complete (2^TB sources, TB <= 4: the register path).  Measured
    //    0.40 / 0.44 / 0.52 ms at 4 / 8 / 16 sources, plus the cross-lane part.
    //  * everything else — a deliberately pessimistic issue model (it prices 8 non-zeros per row at
    //    ~2.7 passes).  The greedy pass is myopic: with an accurate price for the slow classes it
    //    spends its dense budget early and ends up slower (measured 133 ms against 118 ms on
    //    supremacy_n26), and the slow classes vary a lot (a sub table per path, long walks).
    (void)fp64GFlops;
    const double amps = std::ldexp(1.0, c.n);
    const double memNs = 32.0 * amps / hbmGBs; // GB/s == B/ns
    const bool registerPath = c.tileable && c.subTileBits >= 2 && c.subTileBits <= 4 && (1 << c.subTileBits) <= 2 * c.maxPaths;
    const bool knownFast = c.tileable && c.uniform && (c.allIdentitySubs || c.nSub == 1) &&
                           c.kTrue <= (c.maxPaths > 4 ? 4 : 8) && (registerPath || c.maxPaths <= 4);
    // complete block on 3 / 4 upper qubits with untouched low levels: the tensor-core path (tile kernel MODE 5),
    // measured 0.40 / 0.42 ms whatever the block holds
    const bool tensorPath = registerPath && c.allIdentitySubs && c.subTileBits >= 3;
    static const CostKnobs knobs = CostKnobs::fromEnv();
    double factor;
    if (knownFast) {
        factor = knobs.base + (registerPath ? knobs.perSegment * (1 << c.subTileBits) : 0.0);
        if (tensorPath) factor = c.subTileBits == 3 ? knobs.tensor8 : knobs.tensor16;
        if (c.kTrue > 1) factor += knobs.crossLane * c.kTrue * std::max(1.0, c.maxPaths / 4.0); // (16,2): +0.28, (16,4): +0.56 measured
    } else if (tensorPath && c.tileable) {
        factor = (c.subTileBits == 3 ? knobs.tensor8 : knobs.tensor16) + knobs.walkPerTile; // the block is re-read from the walk of every tile
    } else {
        const double instrPerSeg = 40.0 + c.maxPaths * (14.0 + 16.0 * c.kTrue) + 12.0 * c.upperDepth * c.maxPaths / 4.0;
        const double issueNs = (amps / 32.0) * instrPerSeg / (148.0 * 4.0 * 1.8 * 0.6);
        factor = std::max(1.0, issueNs / memNs);
        if (!c.tileable) factor *= 3.0; // walk kernel
    }
    const double launchNs = 3000.0;
    return factor * memNs + launchNs;
}

} // namespace fddb200

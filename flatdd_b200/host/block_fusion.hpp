// block_fusion.hpp — dense-block algebra of the GPU-aware fusion pass.
//
// The reference fuses by multiplying gate DDs and pricing every candidate with its MAC counts
// (src/SwitchSimulator.cpp:267-340; 0.12-0.17 s per circuit, and 0.6-2.3 s when every candidate is also compiled for a
// GPU cost model as round 1 did).  On the GPU the unit of work is a DENSE BLOCK — a 2^k x 2^k matrix on k <= 4 target
// qubits that may depend diagonally on a few context qubits (controls, phases) — so the pass can work on exactly that
// representation: a block is a small table, merging an operation into it is a few thousand multiply-adds, whether an
// operation still fits is a question about qubit SETS, and a DD is built once per emitted block, not once per candidate.
//
//   SmallGate   one circuit operation as a block on its own qubits (extracted once per distinct gate from the DD the host
//               package builds for it: fdd_block_from_matdd; cached by the caller)
//   BlockAcc    the growing block; apply(g) is  block <- g * block  (g acts after what the block already holds)
//   blockToDD   the block as a full-depth flat matrix DD (what crosses the C-ABI): weights normalised towards the root so
//               equal sub-blocks share nodes
#pragma once

#include "flatten.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace fddb200 {

using cplx = std::complex<double>;

struct SmallGate {
    std::vector<int> targets; // non-diagonal qubits, ascending
    std::vector<int> ctx;     // qubits the matrix depends on diagonally, ascending
    std::vector<cplx> table;  // [2^ctx][2^k][2^k], row-major; index bit i <-> targets[i] / ctx[i]
    [[nodiscard]] bool isIdentity() const {
        if (!targets.empty() || !ctx.empty()) return false;
        return table.size() == 1 && table[0] == cplx(1.0, 0.0);
    }
};

namespace detail {
inline std::vector<int> unionSorted(const std::vector<int>& a, const std::vector<int>& b) {
    std::vector<int> out;
    std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}
inline std::vector<int> minusSorted(const std::vector<int>& a, const std::vector<int>& b) {
    std::vector<int> out;
    std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}
inline int indexOf(const std::vector<int>& v, int q) {
    const auto it = std::lower_bound(v.begin(), v.end(), q);
    return (it != v.end() && *it == q) ? static_cast<int>(it - v.begin()) : -1;
}
} // namespace detail

class BlockAcc {
public:
    std::vector<int> targets, ctx;
    std::vector<cplx> table{cplx(1.0, 0.0)};
    int ops = 0;

    [[nodiscard]] bool empty() const { return ops == 0; }
    // qubit sets the block would have after g
    void merged(const SmallGate& g, std::vector<int>& t, std::vector<int>& c) const {
        t = detail::unionSorted(targets, g.targets);
        c = detail::minusSorted(detail::unionSorted(ctx, g.ctx), t);
    }

    void apply(const SmallGate& g) {
        std::vector<int> t2, c2;
        merged(g, t2, c2);
        const std::size_t k2 = t2.size(), nc2 = c2.size();
        const std::size_t rows2 = std::size_t{1} << k2;
        const std::size_t rowsOld = std::size_t{1} << targets.size();
        const std::size_t rowsG = std::size_t{1} << g.targets.size();
        // where every qubit of the old block / of g sits in the new index spaces
        struct Src {
            bool inRow;
            int at;
        };
        auto locate = [&](int q) -> Src {
            const int tr = detail::indexOf(t2, q);
            if (tr >= 0) return {true, tr};
            return {false, detail::indexOf(c2, q)};
        };
        std::vector<Src> oldT, oldC, gT, gC;
        for (int q : targets) oldT.push_back(locate(q));
        for (int q : ctx) oldC.push_back(locate(q));
        for (int q : g.targets) gT.push_back(locate(q));
        for (int q : g.ctx) gC.push_back(locate(q));
        // bits of the new row index the old block does not act on at all (new targets that were neither target nor context of it)
        std::size_t untouched = rows2 - 1;
        for (const Src& s : oldT) untouched &= ~(std::size_t{1} << s.at);
        std::size_t promoted = 0; // old context qubits that are targets now: the old block is diagonal in them
        for (const Src& s : oldC) {
            if (s.inRow) {
                untouched &= ~(std::size_t{1} << s.at);
                promoted |= std::size_t{1} << s.at;
            }
        }
        std::size_t gMask = 0; // bits of the new row index g replaces
        for (const Src& s : gT) gMask |= std::size_t{1} << s.at;
        // per new row index x: the old row, the old context bits taken from the row, g's row, g's context bits from the row
        std::vector<uint32_t> oldRowOf(rows2), oldCtxOfRow(rows2), gRowOf(rows2), gCtxOfRow(rows2);
        for (std::size_t x = 0; x < rows2; ++x) {
            uint32_t a = 0, b = 0, c = 0, d = 0;
            for (std::size_t i = 0; i < oldT.size(); ++i) a |= static_cast<uint32_t>((x >> oldT[i].at) & 1U) << i;
            for (std::size_t i = 0; i < oldC.size(); ++i) {
                if (oldC[i].inRow) b |= static_cast<uint32_t>((x >> oldC[i].at) & 1U) << i;
            }
            for (std::size_t i = 0; i < gT.size(); ++i) c |= static_cast<uint32_t>((x >> gT[i].at) & 1U) << i;
            for (std::size_t i = 0; i < gC.size(); ++i) {
                if (gC[i].inRow) d |= static_cast<uint32_t>((x >> gC[i].at) & 1U) << i;
            }
            oldRowOf[x] = a;
            oldCtxOfRow[x] = b;
            gRowOf[x] = c;
            gCtxOfRow[x] = d;
        }
        // the row index with g's target bits replaced by the bits of kk
        std::vector<std::size_t> spreadG(rowsG);
        for (std::size_t kk = 0; kk < rowsG; ++kk) {
            std::size_t v = 0;
            for (std::size_t i = 0; i < gT.size(); ++i) v |= ((kk >> i) & 1U) << gT[i].at;
            spreadG[kk] = v;
        }
        std::vector<cplx> out((rows2 * rows2) << nc2, cplx(0.0, 0.0));
        for (std::size_t cx = 0; cx < (std::size_t{1} << nc2); ++cx) {
            uint32_t oldCtxOfC = 0, gCtxOfC = 0;
            for (std::size_t i = 0; i < oldC.size(); ++i) {
                if (!oldC[i].inRow) oldCtxOfC |= static_cast<uint32_t>((cx >> oldC[i].at) & 1U) << i;
            }
            for (std::size_t i = 0; i < gC.size(); ++i) {
                if (!gC[i].inRow) gCtxOfC |= static_cast<uint32_t>((cx >> gC[i].at) & 1U) << i;
            }
            cplx* dst = out.data() + cx * rows2 * rows2;
            for (std::size_t r = 0; r < rows2; ++r) {
                const cplx* gRow = g.table.data() + ((static_cast<std::size_t>(gCtxOfC | gCtxOfRow[r]) * rowsG + gRowOf[r]) * rowsG);
                const std::size_t rBase = r & ~gMask;
                for (std::size_t kk = 0; kk < rowsG; ++kk) {
                    const cplx gv = gRow[kk];
                    if (gv == cplx(0.0, 0.0)) continue;
                    const std::size_t k = rBase | spreadG[kk]; // intermediate index: r with g's targets replaced
                    const cplx* oldRow = table.data() + ((static_cast<std::size_t>(oldCtxOfC | oldCtxOfRow[k]) * rowsOld + oldRowOf[k]) * rowsOld);
                    // columns q of the new block that the old block connects to k: equal to k on the untouched and promoted bits
                    const std::size_t fixedMask = untouched | promoted;
                    const std::size_t fixedBits = k & fixedMask;
                    for (std::size_t qo = 0; qo < rowsOld; ++qo) {
                        const cplx mv = oldRow[qo];
                        if (mv == cplx(0.0, 0.0)) continue;
                        std::size_t q = fixedBits;
                        for (std::size_t i = 0; i < oldT.size(); ++i) q |= ((qo >> i) & 1U) << oldT[i].at;
                        dst[r * rows2 + q] += gv * mv;
                    }
                }
            }
        }
        targets.swap(t2);
        ctx.swap(c2);
        table.swap(out);
        ++ops;
    }
};

// A block as a full-depth flat matrix DD: target levels branch four ways, context levels two ways (diagonal), every other
// level is an identity-like node [a 0; 0 a].  Nodes are normalised (largest successor weight becomes 1 and moves to the
// incoming edge) and merged on a 2^-46 grid, so sub-blocks that are equal up to a factor share nodes.
class BlockDDBuilder {
public:
    explicit BlockDDBuilder(int nQubits) : n_(nQubits) {}

    FlatMatDD build(const std::vector<int>& targets, const std::vector<int>& ctx, const std::vector<cplx>& table) {
        out_ = FlatMatDD{};
        out_.n_qubits = n_;
        unique_.clear();
        memo_.clear();
        ident_.assign(static_cast<std::size_t>(n_) + 1, kUnset);
        role_.assign(static_cast<std::size_t>(n_), 0);
        at_.assign(static_cast<std::size_t>(n_), 0);
        lowest_ = n_;
        for (std::size_t i = 0; i < targets.size(); ++i) {
            role_[static_cast<std::size_t>(targets[i])] = 2;
            at_[static_cast<std::size_t>(targets[i])] = static_cast<int>(i);
            lowest_ = std::min(lowest_, targets[i]);
        }
        for (std::size_t i = 0; i < ctx.size(); ++i) {
            role_[static_cast<std::size_t>(ctx[i])] = 1;
            at_[static_cast<std::size_t>(ctx[i])] = static_cast<int>(i);
            lowest_ = std::min(lowest_, ctx[i]);
        }
        table_ = &table;
        rows_ = std::size_t{1} << targets.size();
        const Edge root = make(n_ - 1, 0, 0, 0);
        if (root.w == cplx(0, 0)) throw std::runtime_error("blockToDD: zero matrix");
        out_.root = root.node;
        out_.root_weight[0] = root.w.real();
        out_.root_weight[1] = root.w.imag();
        return std::move(out_);
    }

private:
    static constexpr int32_t kUnset = FDD_TERMINAL - 1;
    struct Edge {
        int32_t node = FDD_TERMINAL;
        cplx w{0, 0};
    };
    static int64_t grid(double x) { return static_cast<int64_t>(std::llround(x * 70368744177664.0)); } // 2^-46 steps
    int32_t identChain(int lv) {
        if (lv < 0) return FDD_TERMINAL;
        int32_t& slot = ident_[static_cast<std::size_t>(lv)];
        if (slot == kUnset) {
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
    // sub-matrix with the block bits of all levels above `lv` fixed to (row r, column c, context x)
    Edge make(int lv, std::size_t r, std::size_t c, std::size_t x) {
        if (lv < lowest_) { // identity below the lowest block qubit, the entry rides on the edge
            const cplx v = (*table_)[(x * rows_ + r) * rows_ + c];
            if (v == cplx(0, 0)) return Edge{};
            return Edge{identChain(lv), v};
        }
        const uint64_t memoKey = (static_cast<uint64_t>(lv) << 40) | (static_cast<uint64_t>(x) << 16) | (static_cast<uint64_t>(r) << 8) | static_cast<uint64_t>(c);
        const auto hit = memo_.find(memoKey);
        if (hit != memo_.end()) return hit->second;
        std::array<Edge, 4> e{};
        const int role = role_[static_cast<std::size_t>(lv)];
        const int p = at_[static_cast<std::size_t>(lv)];
        if (role == 2) {
            for (std::size_t rb = 0; rb < 2; ++rb) {
                for (std::size_t cb = 0; cb < 2; ++cb) e[2 * rb + cb] = make(lv - 1, r | (rb << p), c | (cb << p), x);
            }
        } else if (role == 1) {
            e[0] = make(lv - 1, r, c, x);
            e[3] = make(lv - 1, r, c, x | (std::size_t{1} << p));
        } else {
            e[0] = e[3] = make(lv - 1, r, c, x);
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
    int lowest_ = 0;
    std::size_t rows_ = 1;
    FlatMatDD out_;
    const std::vector<cplx>* table_ = nullptr;
    std::vector<int> role_, at_;
    std::vector<int32_t> ident_;
    std::unordered_map<std::array<int64_t, 13>, int32_t, KeyHash> unique_;
    std::unordered_map<uint64_t, Edge> memo_;
};

// The dense block of a flat matrix DD through the library (fdd_block_from_matdd); false when the gate is not a block.
inline bool smallGateFromDD(const FlatMatDD& dd, SmallGate& out, int maxCtx = 10) {
    const fdd_matdd m = view(dd);
    int32_t nT = 0, nC = 0;
    int32_t t[4] = {0, 0, 0, 0};
    int32_t c[16] = {0};
    // first call sizes the table, second fills it
    std::vector<double> buf(2 * 256);
    int rc = fdd_block_from_matdd(&m, maxCtx, &nT, t, &nC, c, buf.data(), buf.size());
    if (rc == FDD_ERR_INVALID && nT >= 0 && nC >= 0 && (static_cast<std::size_t>(2) << (2 * nT + nC)) > buf.size()) {
        buf.assign(static_cast<std::size_t>(2) << (2 * nT + nC), 0.0);
        rc = fdd_block_from_matdd(&m, maxCtx, &nT, t, &nC, c, buf.data(), buf.size());
    }
    if (rc == FDD_ERR_TOO_DENSE) return false;
    fddCheck(rc, "fdd_block_from_matdd");
    out.targets.assign(t, t + nT);
    out.ctx.assign(c, c + nC);
    const std::size_t entries = std::size_t{1} << (2 * nT + nC);
    out.table.resize(entries);
    for (std::size_t i = 0; i < entries; ++i) out.table[i] = cplx(buf[2 * i], buf[2 * i + 1]);
    return true;
}

} // namespace fddb200

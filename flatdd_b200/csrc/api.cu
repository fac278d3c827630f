// api.cu — the C-ABI of include/flatdd_b200.h on top of the sm_100a kernels in kernels.cuh.
// No CPU fallback: every compute entry point needs a CUDA device and fails loudly without one.
#include "flatdd_b200.h"
#include "gate_compile.hpp"
#include "kernels.cuh"
#include "comm.cuh"
#include "block_compile.hpp"
#include <atomic>
#include <exception>
#include <thread>
#include "block_kernel.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <string>
#include <vector>

using namespace fddb200;

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& msg) {
    g_lastError = msg;
    return code;
}

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct CommError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        const cudaError_t err__ = (expr);                                                                    \
        if (err__ != cudaSuccess) {                                                                          \
            throw CudaError(std::string(#expr) + ": " + cudaGetErrorName(err__) + " (" + cudaGetErrorString(err__) + ")"); \
        }                                                                                                    \
    } while (0)

// Runs `body`, mapping exceptions to error codes.
template <class F> int guarded(F&& body) {
    try {
        body();
        return FDD_OK;
    } catch (const CudaError& e) {
        return fail(FDD_ERR_CUDA, e.what());
    } catch (const CommError& e) {
        return fail(FDD_ERR_COMM, e.what());
    } catch (const std::invalid_argument& e) {
        return fail(FDD_ERR_INVALID, e.what());
    } catch (const std::length_error& e) {
        return fail(FDD_ERR_TOO_DENSE, e.what());
    } catch (const std::logic_error& e) {
        return fail(FDD_ERR_STATE, e.what());
    } catch (const std::runtime_error& e) {
        return fail(FDD_ERR_INVALID, e.what());
    } catch (const std::exception& e) {
        return fail(FDD_ERR_INVALID, e.what());
    }
}

constexpr size_t kSmemBudget = 227 * 1024;
constexpr size_t kTableSmemLimit = 96 * 1024;
constexpr int kMaxContextBits = 10; // context table of the tensor-core path: at most 1024 blocks (4 MiB at 16 x 16)

} // namespace

struct fdd_gate {
    CompiledGate host;
    int device = 0;
    // device copies
    UpperNode* dUpper = nullptr;
    double2* dSubW = nullptr;
    uint8_t* dSubCol = nullptr;
    int32_t* dSubK = nullptr;
    uint8_t* dSubFlags = nullptr;
    void* dBlob = nullptr; // one allocation holding all of the above
    mutable double2* dCtx = nullptr; // tensor-core path: blocks of a non-uniform gate per value of its context bits (built at the first launch)
    // dense-block path (block_kernel.cuh): the gate as a 2^k x 2^k block with its context table, when it is one
    std::unique_ptr<DenseBlock> block;
    double* dTable = nullptr;
    bool tableBorrowed = false; // dTable points into an allocation somebody else frees (fdd_apply_many: one upload for the whole call)
    uint64_t serial = 0; // identifies the gate in the plan cache (pointers are reused by the allocator)
    const fdd_matdd* source = nullptr; // fdd_apply_many only: the caller's table while the call runs (lazy tables for the older kernels)
};

struct fdd_ctx {
    int n = 0;
    int nLocal = 0;
    int device = 0;
    int rank = 0;
    int world = 1;
    int worldBits = 0;
    double2* buf[2] = {nullptr, nullptr};
    int cur = 0;
    bool hasState = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false;
    float lastMs = 0.0f;
    uint64_t launches = 0;
    uint64_t tensorCoreLaunches = 0;
    uint64_t flatTableLaunches = 0;
    uint64_t contextLaunches = 0;
    int smCount = 148;
    // tunables
    int variant = 2;      // 2: tile kernel when the gate allows it, else 1; 1: cp.async ring walk; 0: register walk
    int warpsPerCta = 0;  // 0 = chosen per launch
    int ctasPerSm = 0; // 0 = as many as shared memory allows (capped)
    int prefetch = 8;
    int forceMode = -1;   // experiments: force the tile-kernel MODE (1, 2 or 3) where it applies
    int denseSlots = 1;   // experiments: 0 disables the dense register path of the tile kernel
    int flatTable = 1;    // uniform gates with a sub table per path: flat precombined table (tile kernel MODE 6); 0: MODE 3
    int contextTable = 1; // tensor-core path of non-uniform gates: look the block up per tile instead of walking the gate per tile
    int pdl = 0;          // 1: tile kernel launches overlap their prologue with the tail of the previous launch (measured: no gain)
    int dmma = 1;         // dense upper blocks of 8 / 16 segments on the FP64 tensor cores (tile kernel MODE 5); 0: CUDA-core FMAs
    int exchangeUnroll = 8;
    int exchangeCtasPerSm = 4;
    int blockKernel = 1;      // gates that are dense blocks (<= 4 non-diagonal qubits anywhere, <= 10 context qubits) take the tile-resident kernel
    int blockTileBits = 12;   // preferred tile size of a pass (log2 amplitudes); grows to 13 when the blocks of a pass need it
    int blockMaxPerPass = 4;  // blocks that may share a pass; 1: one pass per block (A/B against the fused passes)
    int blockMaxTileBits = 12; // largest tile a pass of SEVERAL blocks may need (2^12: three tile buffers fit, copy and tensor work overlap;
                               // 2^13 fills the shared memory with one buffer: measured slower than separate passes)
    int blockWarps = 0;       // experiments: 16 = sixteen warps per CTA with one unit per iteration
    int blockUnits = 2;       // units per iteration and warp (2: twelve independent tensor-core chains, one CTA per SM; 1: two CTAs per SM)
    int blockFuseExchange = 1; // fdd_*_apply_many_exchange: the last pass of the stretch writes the traded half into the partner's buffer
    uint64_t fusedExchanges = 0;
    int blockReorder = 1;      // blocks move up over gates they commute with to share a pass (orderForPasses)
    int blockTablesShared = 1; // multi-block passes keep the blocks' matrix tables in shared memory when they fit
    int blockBuffers = 3;     // tile buffers per CTA when they fit (copy-in, tensor-core work and copy-out of consecutive tiles overlap)
    int blockWs = 1;          // warp-specialised kernel (memory warps + compute warps); 0: every warp loads, computes and stores in turn
    uint64_t blockLaunches = 0;
    uint64_t blocksApplied = 0;
    uint64_t gateSerial = 0;
    struct PlannedPass {
        PassParams params;
        int maxUnits = 0;
        int nBuffers = 1;
        int warps = 8;
        int unitsPerIter = 2;
        bool ws = true;
        int grid = 0;
        size_t smem = 0;
    };
    std::map<std::vector<uint64_t>, PlannedPass> passPlans;
    // scratch
    double* dPartial = nullptr;
    double* dNorm = nullptr;
    std::vector<int32_t> logicalToPhysical;
    std::map<std::tuple<const void*, size_t, size_t, int, int>, std::pair<int, int>> launchShapes; // tile-kernel CTA shapes
    // multi-GPU
    const double2* peerBuf[2][kMaxPeers] = {}; // state buffers of every rank (own ones included), mapped through CUDA IPC
    ncclComm_t comm = nullptr;
    double* dBarrier = nullptr;
    uint64_t exchanges = 0;
    // exchange ordering without NCCL calls: two epoch words per rank in peer-mapped memory (comm.cuh)
    uint32_t* dFlags = nullptr;                       // [0] ready epoch, [1] done epoch, [2] CTA counter (local use)
    const uint32_t* peerFlags[kMaxPeers] = {};
    uint32_t exchangeEpoch = 0;
    int exchangeFlags = 1;    // 1: flag protocol inside the exchange kernel; 0: two stream-ordered NCCL barriers around it

    [[nodiscard]] uint64_t localDim() const { return uint64_t{1} << nLocal; }
};

namespace {

void useDevice(const fdd_ctx* c) { CUDA_TRY(cudaSetDevice(c->device)); }

struct Timed {
    fdd_ctx* c;
    explicit Timed(fdd_ctx* ctx) : c(ctx) {
        if (c->timing) cudaEventRecord(c->ev0, c->stream);
    }
    ~Timed() {
        if (c->timing) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            cudaEventElapsedTime(&c->lastMs, c->ev0, c->ev1);
        }
    }
};

int gridFor(const fdd_ctx* c, uint64_t items, int block) {
    const uint64_t want = (items + block - 1) / block;
    return static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(c->smCount) * 8)));
}

void uploadGate(fdd_gate* g, cudaStream_t stream) {
    const CompiledGate& h = g->host;
    const size_t bUpper = h.upper.size() * sizeof(UpperNode);
    const size_t nEnt = static_cast<size_t>(h.nSub) * h.kMax * 32;
    const size_t bW = nEnt * 16, bK = static_cast<size_t>(h.nSub) * 4, bCol = nEnt, bF = static_cast<size_t>(h.nSub);
    const size_t total = bUpper + bW + bK + bCol + bF + 64;
    std::vector<unsigned char> blob(total, 0);
    size_t off = 0;
    auto put = [&](const void* src, size_t bytes) {
        const size_t at = off;
        if (bytes != 0) std::memcpy(blob.data() + at, src, bytes);
        off += bytes;
        return at;
    };
    const size_t oUpper = put(h.upper.data(), bUpper);
    const size_t oW = put(h.subW.data(), bW);
    const size_t oK = put(h.subK.data(), bK);
    const size_t oCol = put(h.subCol.data(), bCol);
    const size_t oF = put(h.subFlags.data(), bF);
    CUDA_TRY(cudaMallocAsync(&g->dBlob, total, stream));
    // pageable source: the copy is staged by the runtime before the call returns
    CUDA_TRY(cudaMemcpyAsync(g->dBlob, blob.data(), total, cudaMemcpyHostToDevice, stream));
    auto* base = static_cast<unsigned char*>(g->dBlob);
    g->dUpper = reinterpret_cast<UpperNode*>(base + oUpper);
    g->dSubW = reinterpret_cast<double2*>(base + oW);
    g->dSubK = reinterpret_cast<int32_t*>(base + oK);
    g->dSubCol = base + oCol;
    g->dSubFlags = base + oF;
}

using Kernel = void (*)(const WalkParams);

template <int TB> Kernel tileKernelTB(int mode, int kt) {
    switch (mode) {
        case 0: return dmavm_tile_kernel<TB, 0, 0>;
        case 5:
            if constexpr (TB == 3 || TB == 4) {
                return dmavm_tile_kernel<TB, 5, 0>;
            } else {
                return nullptr;
            }
        case 1:
            return kt == 2 ? dmavm_tile_kernel<TB, 1, 2>
                           : (kt == 4 ? dmavm_tile_kernel<TB, 1, 4> : (kt == 8 ? dmavm_tile_kernel<TB, 1, 8> : dmavm_tile_kernel<TB, 1, 16>));
        case 2: return kt == 2 ? dmavm_tile_kernel<TB, 2, 2> : (kt == 4 ? dmavm_tile_kernel<TB, 2, 4> : dmavm_tile_kernel<TB, 2, 8>);
        case 4: return kt == 2 ? dmavm_tile_kernel<TB, 4, 2> : (kt == 4 ? dmavm_tile_kernel<TB, 4, 4> : dmavm_tile_kernel<TB, 4, 8>);
        case 6: // kt = entries per row of the flat table
            if constexpr (TB <= 3) {
                return kt == 4 ? dmavm_tile_kernel<TB, 6, 4> : (kt == 8 ? dmavm_tile_kernel<TB, 6, 8> : dmavm_tile_kernel<TB, 6, 16>);
            } else {
                return nullptr;
            }
        default:
            return kt == 2 ? dmavm_tile_kernel<TB, 3, 2>
                           : (kt == 4 ? dmavm_tile_kernel<TB, 3, 4> : (kt == 8 ? dmavm_tile_kernel<TB, 3, 8> : dmavm_tile_kernel<TB, 3, 0>));
    }
}

Kernel tileKernel(int tb, int mode, int kt) {
    switch (tb) {
        case 0: return tileKernelTB<0>(mode, kt);
        case 1: return tileKernelTB<1>(mode, kt);
        case 2: return tileKernelTB<2>(mode, kt);
        case 3: return tileKernelTB<3>(mode, kt);
        case 4: return tileKernelTB<4>(mode, kt);
        default: return tileKernelTB<5>(mode, kt);
    }
}

void launchWalk(fdd_ctx* c, const fdd_gate* g) {
    const CompiledGate& h = g->host;
    if (h.n != c->n) throw std::invalid_argument("gate has " + std::to_string(h.n) + " qubits, context has " + std::to_string(c->n));
    if (!c->hasState) throw std::logic_error("no state: call fdd_convert / fdd_set_state / fdd_set_zero_state first");
    if (c->world > 1 && (h.nonDiagMask >> c->nLocal) != 0) {
        // reading a peer's buffer while that peer runs its own launch would need cross-rank ordering;
        // the contract is: swap the global qubit with a local one first (fdd_exchange_qubits)
        throw std::logic_error("gate is non-diagonal on a global qubit: exchange it with a local qubit first (fdd_exchange_qubits)");
    }
    WalkParams p{};
    p.y = c->buf[c->cur];
    p.z = c->buf[c->cur ^ 1];
    for (int r = 0; r < kMaxPeers; ++r) p.peerY[r] = c->peerBuf[c->cur][r];
    if (c->world == 1) p.peerY[0] = p.y;
    p.upper = g->dUpper;
    p.nUpper = static_cast<int>(h.upper.size());
    p.subW = g->dSubW;
    p.subCol = g->dSubCol;
    p.subK = g->dSubK;
    p.subFlags = g->dSubFlags;
    p.nSub = h.nSub;
    p.kMax = h.kMax;
    p.root = h.root;
    p.rootW = make_double2(h.rootW[0], h.rootW[1]);
    p.nLocal = c->nLocal;
    p.segBits = std::min(h.segBits, c->nLocal);
    if (p.segBits != h.segBits) throw std::invalid_argument("shard too small for this gate (fewer than 5 local qubits)");
    p.maxPaths = h.maxPaths;
    p.stackCap = h.stackCap;
    p.rank = static_cast<uint32_t>(c->rank);
    p.worldBits = c->worldBits;
    p.nSeg = static_cast<uint32_t>(c->localDim() >> p.segBits);
    p.nTiles = (p.nSeg + 31) / 32;
    int D = c->prefetch;
    if (D != 1 && D != 2 && D != 4 && D != 8 && D != 16) D = 8;
    p.prefetch = D;
    const size_t tableBytes = walkTableSmem(p.nUpper, p.nSub, p.kMax);
    p.tablesInSmem = tableBytes <= kTableSmemLimit ? 1 : 0;
    const size_t fixed = p.tablesInSmem ? tableBytes : 0;

    if (c->variant == 2 && h.tileable && h.upper.size() < (1u << 22) && h.nSub < (1 << 22)) {
        // ---- tile kernel ----------------------------------------------------------------------------
        bool allIdentity = true;
        for (uint8_t f : h.subFlags) allIdentity = allIdentity && (f & SUB_IDENTITY);
        // ELL width rounded up for static unrolling; wider tables take the run-time loop of MODE 3
        int kt = h.kMax <= 2 ? 2 : (h.kMax <= 4 ? 4 : (h.kMax <= 8 ? 8 : 0));
        int mode = 0;
        if (!allIdentity) {
            if (h.nSub == 1 && h.kMax == 16 && h.maxPaths <= 2) {
                kt = 16; // a block that is dense on four of the five lane qubits: gather first, table in registers
                mode = 1;
            } else if (h.nSub == 1 && kt > 0) {
                // instruction estimates per output segment: gather first P(8+5K), shuffle last 9P+8K
                mode = h.maxPaths * (8 + 5 * kt) <= 9 * h.maxPaths + 8 * kt ? 1 : 2;
            } else {
                mode = 3;
                // per-sub entry lists when they fit the packed counters and the slot budget
                int slots = 0;
                bool fits = kt > 0 && h.nSub <= 8;
                for (int ps : h.subPaths) {
                    fits = fits && ps <= 15;
                    slots += ps;
                }
                if (fits && slots <= 15 && c->forceMode == 4) mode = 4; // opt-in: measured slower than MODE 3 (padding)
            }
        }
        if (c->forceMode >= 0 && c->forceMode <= 3 && !allIdentity && (c->forceMode == 3 || (h.nSub == 1 && kt > 0))) mode = c->forceMode;
        // uniform gate with a sub table per path and few entries per row: flat precombined table (MODE 6)
        int flatEntries = 0;
        if (mode == 3 && h.uniform && c->flatTable && h.subTileBits <= 3 && h.maxPaths * h.kMax <= 16 && (h.maxPaths * h.kMax << h.subTileBits) <= 64) {
            const int e = h.maxPaths * h.kMax;
            flatEntries = e <= 4 ? 4 : (e <= 8 ? 8 : 16);
            mode = 6;
            kt = flatEntries;
        }
        p.tileBits = h.tileBits;
        p.subTileBits = h.subTileBits;
        p.tileMask = h.tileMask;
        p.fillMask = h.fillMask;
        p.nTiles = p.nSeg >> h.tileBits;
        p.uniform = h.uniform ? 1 : 0;
        const int tSegs = 1 << h.subTileBits;
        p.denseSlots = ((mode == 0 || mode == 2) && tSegs >= 4 && tSegs <= 16 && tSegs <= 2 * h.maxPaths && c->denseSlots) ? 1 : 0;
        if (p.denseSlots) p.maxPaths = std::max(p.maxPaths, tSegs); // the entry area holds one slot per source segment
        // a complete block on 3 or 4 upper qubits with untouched low levels is a real dense contraction:
        // (T x T complex) x (T x 32 complex) per sub-tile on the FP64 tensor cores
        if (mode == 0 && p.denseSlots && (tSegs == 8 || tSegs == 16) && p.segBits == 5 && c->dmma) mode = 5;
        if (mode == 4) { // the entry area holds the concatenated per-sub lists
            int slots = 0;
            for (int sIdx = 0; sIdx < std::min(h.nSub, 8); ++sIdx) {
                p.subBase[sIdx] = static_cast<uint8_t>(slots);
                slots += h.subPaths[static_cast<size_t>(sIdx)];
            }
            for (int sIdx = std::min(h.nSub, 8); sIdx < 9; ++sIdx) p.subBase[sIdx] = static_cast<uint8_t>(slots);
            p.maxPaths = std::max(slots, 1);
        }
        const Kernel kernel = tileKernel(h.subTileBits, mode, kt);
        // pick the CTA width that keeps the most warps resident (registers and shared memory both count);
        // the answer only depends on (kernel, shared memory shape), so it is cached per context
        auto pickShape = [&](size_t fixedBytes, size_t perWarpBytes) {
            int bestW = 0, bestC = 0;
            const auto key = std::make_tuple(reinterpret_cast<const void*>(kernel), fixedBytes, perWarpBytes, c->warpsPerCta, c->ctasPerSm);
            const auto hit = c->launchShapes.find(key);
            if (hit != c->launchShapes.end()) return hit->second;
            CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
            for (int w = 16; w >= 1; --w) {
                if (mode != 5 && w > 8) continue;
                if (c->warpsPerCta > 0 && w > c->warpsPerCta) continue;
                if (mode == 5 && w > kM5MaxWarps) continue; // launch bounds of the tensor-core instantiations
                const size_t smemW = fixedBytes + static_cast<size_t>(w) * perWarpBytes;
                if (smemW > kSmemBudget) continue;
                int resident = 0;
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, w * 32, smemW));
                if (c->ctasPerSm > 0) resident = std::min(resident, c->ctasPerSm);
                if (resident * w > bestW * bestC) {
                    bestW = w;
                    bestC = resident;
                }
            }
            c->launchShapes.emplace(key, std::make_pair(bestW, bestC));
            if (std::getenv("FLATDD_B200_DEBUG") != nullptr) {
                std::fprintf(stderr, "[flatdd_b200] tile kernel TB=%d mode=%d kt=%d: %d warps/CTA x %d CTAs/SM, smem %zu B (cta %zu + %zu per warp), paths %d, uniform %d, dense %d\n",
                             h.subTileBits, mode, kt, bestW, bestC, fixedBytes + bestW * perWarpBytes, fixedBytes, perWarpBytes, p.maxPaths, p.uniform, p.denseSlots);
            }
            return std::make_pair(bestW, bestC);
        };
        auto launchTile = [&](const WalkParams& params, size_t fixedBytes, size_t perWarpBytes, uint32_t warpItems) {
            const auto shape = pickShape(fixedBytes, perWarpBytes);
            if (shape.first == 0) return false;
            const size_t smemT = fixedBytes + static_cast<size_t>(shape.first) * perWarpBytes;
            const uint32_t ctasWantedT = (warpItems + shape.first - 1) / shape.first;
            const int gridT = static_cast<int>(std::max<uint32_t>(1, std::min<uint32_t>(ctasWantedT, static_cast<uint32_t>(c->smCount * shape.second))));
            if (c->pdl) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(static_cast<unsigned>(gridT));
                cfg.blockDim = dim3(static_cast<unsigned>(shape.first * 32));
                cfg.dynamicSmemBytes = smemT;
                cfg.stream = c->stream;
                cudaLaunchAttribute attr{};
                attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr.val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = &attr;
                cfg.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
            } else {
                kernel<<<gridT, shape.first * 32, smemT, c->stream>>>(params);
            }
            CUDA_TRY(cudaGetLastError());
            c->launches++;
            return true;
        };
        size_t perWarpT = tileWarpSmem(p.maxPaths, p.stackCap, h.subTileBits, p.uniform, mode == 5);
        size_t fixedT = fixed + tileCtaSmem(p.maxPaths, p.stackCap, p.uniform) + (mode == 6 ? tileFlatBytes(h.subTileBits, flatEntries) : 0);
        Timed t(c);
        // tensor-core path of a gate whose block depends on a few diagonal upper qubits (controls, phases): walk every
        // value of those bits once into a table (cached with the compiled gate) instead of once per warp tile
        const int ctxBits = __builtin_popcount(h.ctxMask);
        if (mode == 5 && c->contextTable && ctxBits <= kMaxContextBits) {
            // (a uniform gate is the case of zero context bits: one block, and the launch itself has no prologue)
            bool ok = true;
            if (g->dCtx == nullptr) {
                WalkParams pre = p;
                pre.uniform = 0; // every warp of the pre-pass walks into its own entry area
                const size_t prePerWarp = tileWarpSmem(p.maxPaths, p.stackCap, h.subTileBits, 0, 1);
                pre.ctxPhase = 1;
                pre.ctxMask = h.ctxMask;
                pre.nCtx = 1 << ctxBits;
                CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&g->dCtx), (sizeof(double2) * tSegs * tSegs) << ctxBits, c->stream));
                pre.ctxTable = g->dCtx;
                ok = launchTile(pre, fixed, prePerWarp, static_cast<uint32_t>(pre.nCtx));
                if (!ok) {
                    cudaFreeAsync(g->dCtx, c->stream);
                    g->dCtx = nullptr;
                }
            }
            if (ok) {
                p.ctxPhase = 2;
                p.ctxMask = h.ctxMask;
                p.nCtx = 1 << ctxBits;
                p.ctxTable = g->dCtx;
                p.tablesInSmem = 0; // nothing walks the gate in this launch
                fixedT = 0;
                perWarpT = tileRingBytes(h.subTileBits, 1);
                c->contextLaunches++;
            }
        }
        if (launchTile(p, fixedT, perWarpT, p.nTiles)) {
            if (mode == 5) c->tensorCoreLaunches++;
            if (mode == 6) c->flatTableLaunches++;
            c->cur ^= 1;
            return;
        }
        // the per-warp state does not fit: fall through to the walk kernel
        p.nTiles = (p.nSeg + 31) / 32;
    }
    const int variant = c->variant == 0 ? 0 : 1; // 9 forces the chunk kernel below
    const size_t perWarp = walkWarpSmem(p.maxPaths, p.stackCap, variant == 1 ? D : 0);
    int warps = std::max(1, std::min(c->warpsPerCta > 0 ? c->warpsPerCta : 8, 16));
    while (warps > 1 && fixed + warps * perWarp > kSmemBudget) warps >>= 1;
    if (fixed + warps * perWarp > kSmemBudget || c->variant == 9) {
        // too many paths per segment for a resident list: bounded lists + partial sums in shared memory
        const size_t perWarpC = chunkWarpSmem(p.stackCap);
        int warpsC = 4;
        while (warpsC > 1 && warpsC * perWarpC > kSmemBudget) warpsC >>= 1;
        if (warpsC * perWarpC > kSmemBudget) throw std::length_error("gate needs a deeper walk stack than shared memory holds");
        const size_t smemC = warpsC * perWarpC;
        CUDA_TRY(cudaFuncSetAttribute(dmavm_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
        int residentC = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&residentC, dmavm_chunk_kernel, warpsC * 32, smemC));
        const uint32_t ctasC = (p.nTiles + warpsC - 1) / warpsC;
        const int gridC = static_cast<int>(std::max<uint32_t>(1, std::min<uint32_t>(ctasC, static_cast<uint32_t>(c->smCount * std::max(1, residentC)))));
        Timed t(c);
        dmavm_chunk_kernel<<<gridC, warpsC * 32, smemC, c->stream>>>(p);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        c->cur ^= 1;
        return;
    }
    const size_t smem = fixed + warps * perWarp;
    int perSm = static_cast<int>(kSmemBudget / std::max<size_t>(smem, 1));
    perSm = std::max(1, std::min(perSm, std::max(1, 48 / warps))); // at most 48 resident warps per SM
    if (c->ctasPerSm > 0) perSm = std::min(perSm, c->ctasPerSm);
    const uint32_t ctasWanted = (p.nTiles + warps - 1) / warps;
    const int grid = static_cast<int>(std::max<uint32_t>(1, std::min<uint32_t>(ctasWanted, static_cast<uint32_t>(c->smCount * perSm))));

    Timed t(c);
    if (variant == 0) {
        CUDA_TRY(cudaFuncSetAttribute(dmavm_walk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
        dmavm_walk_kernel<0><<<grid, warps * 32, smem, c->stream>>>(p);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(dmavm_walk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
        dmavm_walk_kernel<1><<<grid, warps * 32, smem, c->stream>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    c->cur ^= 1;
}

void freeGate(fdd_gate* g, cudaStream_t stream) {
    if (g == nullptr) return;
    if (g->dCtx != nullptr) {
        cudaSetDevice(g->device);
        if (stream != nullptr) {
            cudaFreeAsync(g->dCtx, stream);
        } else {
            cudaFree(g->dCtx);
        }
    }
    if (g->dBlob != nullptr) {
        cudaSetDevice(g->device);
        if (stream != nullptr) {
            cudaFreeAsync(g->dBlob, stream);
        } else {
            cudaFree(g->dBlob);
        }
    }
    if (g->dTable != nullptr && !g->tableBorrowed) {
        cudaSetDevice(g->device);
        if (stream != nullptr) {
            cudaFreeAsync(g->dTable, stream);
        } else {
            cudaFree(g->dTable);
        }
    }
    delete g;
}

// ---- dense-block path -------------------------------------------------------------------------------------------
// The gate as a dense block, padded to the kernel's shapes, with its table on the device; leaves g->block empty when the
// gate is not a block (more than four non-diagonal qubits, too many context qubits, a non-local target, a tiny register).
// The gate as a dense block, or nothing when it is not one.  Host work only (no CUDA call, touches nothing but `dd`): runs on
// several threads when a boundary call brings many gates.
std::unique_ptr<DenseBlock> extractBlock(const fdd_ctx* c, const fdd_matdd& dd) {
    if (!c->blockKernel || c->variant != 2 || c->nLocal < 8) return nullptr;
    auto blk = std::make_unique<DenseBlock>();
    if (!denseBlockFromDD(dd, *blk)) return nullptr;
    for (int q : blk->targets) {
        if (q >= c->nLocal) return nullptr; // non-diagonal on a global qubit: the caller has to exchange first (launchWalk reports it)
    }
    padBlock(*blk, c->nLocal);
    const DenseBlock* one = blk.get();
    if (minTileBits(&one, 1, c->nLocal) < 0) return nullptr;
    return blk;
}

// The block's matrix table goes to the device (stream ordered).
void uploadBlock(fdd_ctx* c, fdd_gate* g, std::unique_ptr<DenseBlock> blk) {
    if (!blk) return;
    const size_t bytes = blk->table.size() * sizeof(double);
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&g->dTable), bytes, c->stream));
    CUDA_TRY(cudaMemcpyAsync(g->dTable, blk->table.data(), bytes, cudaMemcpyHostToDevice, c->stream)); // pageable: staged before the call returns
    g->block = std::move(blk);
    g->serial = ++c->gateSerial;
}

void attachBlock(fdd_ctx* c, fdd_gate* g, const fdd_matdd& dd) { uploadBlock(c, g, extractBlock(c, dd)); }

// (dmavm_variant != 2 asks for one of the older kernels explicitly)
bool usesBlockPath(const fdd_ctx* c, const fdd_gate* g) { return c->blockKernel && c->variant == 2 && g->block != nullptr && g->dTable != nullptr; }

// One pass of the tile-resident kernel over gates[0..count) (all of them blocks).  Returns false when they do not fit one
// tile or the fragment shape cannot be planned; nothing has been launched then.
// SWAP(global physical bit pg, local physical bit pl) that follows a stretch of gates: fused into the stretch's last pass when
// that pass runs on the warp-specialised block kernel and pl is above the lane bits (PassParams::zPeer)
struct ExchangeSpec {
    int pg = 0, pl = 0;
    bool done = false;
};
bool canFuseExchange(const fdd_ctx* c, const ExchangeSpec& ex) {
    return c->blockFuseExchange != 0 && c->exchangeFlags != 0 && c->comm != nullptr && c->blockWs != 0 && ex.pl >= kLaneBits && ex.pl < c->nLocal;
}

bool launchPass(fdd_ctx* c, const fdd_gate* const* gates, int count, bool cachePlan, ExchangeSpec* ex = nullptr) {
    if (!c->hasState) throw std::logic_error("no state: call fdd_convert / fdd_set_state / fdd_set_zero_state first");
    std::vector<uint64_t> key;
    if (cachePlan) {
        key.reserve(static_cast<size_t>(count) + 1);
        key.push_back(static_cast<uint64_t>(c->blockTileBits) | (static_cast<uint64_t>(c->blockBuffers) << 8) | (static_cast<uint64_t>(c->blockWarps) << 16) | (static_cast<uint64_t>(c->blockUnits) << 24) | (static_cast<uint64_t>(c->blockWs) << 28) | (static_cast<uint64_t>(c->blockTablesShared) << 29));
        for (int i = 0; i < count; ++i) key.push_back(gates[i]->serial);
    }
    fdd_ctx::PlannedPass planned;
    const auto hit = cachePlan ? c->passPlans.find(key) : c->passPlans.end();
    if (hit != c->passPlans.end()) {
        planned = hit->second;
    } else {
        std::vector<const DenseBlock*> blocks(static_cast<size_t>(count));
        for (int i = 0; i < count; ++i) {
            if (gates[i]->block->n != c->n) throw std::invalid_argument("gate and context differ in the number of qubits");
            blocks[static_cast<size_t>(i)] = gates[i]->block.get();
        }
        const int need = minTileBits(blocks.data(), count, c->nLocal);
        const int maxBits = std::min(c->nLocal, kPassMaxTileBits);
        if (need < 0 || need > maxBits) return false;
        // preferred tile size, larger when the blocks need it or when a fragment shape cannot be planned (too few free tile bits)
        int tileBits = std::max(need, std::min(c->blockTileBits, maxBits));
        while (tileBits <= maxBits && !planPass(blocks.data(), count, c->nLocal, c->rank, tileBits, planned.params)) ++tileBits;
        if (tileBits > maxBits) return false;
        for (int i = 0; i < count; ++i) {
            planned.params.blocks[i].table = gates[i]->dTable;
            planned.maxUnits = std::max(planned.maxUnits, planned.params.blocks[i].nUnits);
        }
        planned.ws = c->blockWs != 0;
        if (planned.ws) {
            // warp-specialised kernel: as many tile buffers as fit (three keep copy-in, tensor work and copy-out in flight)
            planned.grid = static_cast<int>(std::min<uint64_t>(planned.params.nTiles, static_cast<uint64_t>(c->smCount)));
            const uint32_t tilesPerCta = (planned.params.nTiles + static_cast<uint32_t>(planned.grid) - 1) / static_cast<uint32_t>(planned.grid);
            // a pass of several blocks is bound by its tensor work: two buffers keep the memory warps ahead of it, and the room
            // of the third one holds the blocks' matrix tables (measured on B200: 0.66 -> 0.62 ms for two 16 x 16 blocks with two
            // buffers alone; a single-block pass is as fast with two buffers as with three)
            const int wantBuffers = count > 1 ? std::min(2, std::max(1, c->blockBuffers)) : std::min(3, std::max(1, c->blockBuffers));
            planned.nBuffers = 1;
            for (int nb = wantBuffers; nb >= 1; --nb) {
                if (blockPassSmemWs(planned.params.tileBits, nb, count, planned.maxUnits, tilesPerCta) <= kSmemBudget) {
                    planned.nBuffers = nb;
                    break;
                }
            }
            size_t tableAt = blockPassTableArea(planned.params.tileBits, planned.nBuffers, count, planned.maxUnits, tilesPerCta);
            const size_t tableBase = tableAt;
            for (int i = 0; i < kPassMaxBlocks; ++i) planned.params.tableSmem[i] = kTableInGlobal;
            if (count > 1 && c->blockTablesShared != 0) {
                for (int i = 0; i < count; ++i) {
                    const BlockDesc& b = planned.params.blocks[i];
                    const size_t bytes = (static_cast<size_t>(16) << b.nCtx) << (2 * b.k);
                    if (tableAt + bytes > kSmemBudget) continue;
                    planned.params.tableSmem[i] = static_cast<uint32_t>(tableAt - tableBase);
                    tableAt += bytes;
                }
            }
            planned.smem = tableAt == tableBase ? blockPassSmemWs(planned.params.tileBits, planned.nBuffers, count, planned.maxUnits, tilesPerCta) : tableAt;
            if (planned.smem > kSmemBudget) return false;
            static bool wsAttr = false;
            if (!wsAttr) {
                CUDA_TRY(cudaFuncSetAttribute(dmavm_block_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
                if (const char* carve = std::getenv("FLATDD_B200_CARVEOUT")) // experiments: shared-memory share of the L1 array in percent
                    CUDA_TRY(cudaFuncSetAttribute(dmavm_block_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(carve)));
                wsAttr = true;
            }
            planned.warps = kComputeWarps + kMemoryWarps;
        } else {
        planned.nBuffers = (c->blockBuffers >= 2 && blockPassSmem(planned.params.tileBits, 2, count, planned.maxUnits) <= kSmemBudget) ? 2 : 1;
        planned.smem = blockPassSmem(planned.params.tileBits, planned.nBuffers, count, planned.maxUnits);
        if (planned.smem > kSmemBudget) return false;
        // one CTA per SM: 8 warps, two units per iteration (12 independent tensor-core chains per warp, up to 255 registers);
        // block_warps = 16: 16 warps with one unit per iteration; block_warps = 8 and block_units = 1: two CTAs per SM
        planned.warps = c->blockWarps == 16 ? 16 : 8;
        planned.unitsPerIter = (planned.warps == 8 && c->blockUnits != 1) ? 2 : 1;
        using BlockKernel = void (*)(const PassParams, int, int);
        const BlockKernel kernel = planned.warps == 16 ? dmavm_block_kernel<16, 1> : (planned.unitsPerIter == 2 ? dmavm_block_kernel<8, 2> : dmavm_block_kernel<8, 1>);
        static bool attrSet = false;
        if (!attrSet) {
            CUDA_TRY(cudaFuncSetAttribute(dmavm_block_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
            CUDA_TRY(cudaFuncSetAttribute(dmavm_block_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
            CUDA_TRY(cudaFuncSetAttribute(dmavm_block_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
            attrSet = true;
        }
        int resident = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, 32 * planned.warps, planned.smem));
        if (resident < 1) return false;
        planned.grid = static_cast<int>(std::min<uint64_t>(planned.params.nTiles, static_cast<uint64_t>(c->smCount) * static_cast<uint64_t>(resident)));
        }
        if (std::getenv("FLATDD_B200_DEBUG") != nullptr) {
            std::fprintf(stderr, "[flatdd_b200] block pass: %d block(s), tile 2^%d x %d buffer(s), %d warps%s, grid %d, smem %zu B, conflicts", count, planned.params.tileBits, planned.nBuffers, planned.warps, planned.ws ? " (warp-specialised)" : "", planned.grid, planned.smem);
            for (int i = 0; i < count; ++i) std::fprintf(stderr, " %d(k%d,c%d)", planned.params.blocks[i].conflictWays, planned.params.blocks[i].k, planned.params.blocks[i].nCtx);
            std::fprintf(stderr, "\n");
        }
        if (cachePlan) c->passPlans.emplace(key, planned);
    }
    planned.params.y = c->buf[c->cur];
    planned.params.z = c->buf[c->cur ^ 1];
    planned.params.zPeer = nullptr;
    if (ex != nullptr && planned.ws && canFuseExchange(c, *ex)) {
        const int gbit = ex->pg - c->nLocal;
        const int partner = c->rank ^ (1 << gbit);
        planned.params.zPeer = const_cast<double2*>(c->peerBuf[c->cur ^ 1][partner]);
        planned.params.exchSegBit = ex->pl - kLaneBits;
        planned.params.exchMyBit = static_cast<uint32_t>((c->rank >> gbit) & 1);
        planned.params.exchEpoch = ++c->exchangeEpoch;
        planned.params.exchMyFlags = c->dFlags;
        planned.params.exchPartnerFlags = c->peerFlags[partner];
        planned.params.exchCounter = reinterpret_cast<unsigned int*>(c->dFlags + 2);
        ex->done = true;
        c->exchanges++;
        c->fusedExchanges++;
    }
    static const char* skipEnv = std::getenv("FLATDD_B200_BLOCK_SKIP");
    planned.params.debugSkip = skipEnv != nullptr ? static_cast<uint32_t>(std::atoi(skipEnv)) : 0u;
    static const bool clocksEnv = std::getenv("FLATDD_B200_BLOCK_CLOCKS") != nullptr;
    static long long* dClocks = nullptr;
    if (clocksEnv && dClocks == nullptr) CUDA_TRY(cudaMalloc(&dClocks, sizeof(long long) * 4 * 16 * 1024));
    planned.params.debugClocks = clocksEnv ? dClocks : nullptr;
    {
        Timed t(c);
        if (planned.ws) {
            dmavm_block_ws_kernel<<<planned.grid, kBlockThreads, planned.smem, c->stream>>>(planned.params, planned.maxUnits, planned.nBuffers);
        } else if (planned.warps == 16) {
            dmavm_block_kernel<16, 1><<<planned.grid, 32 * 16, planned.smem, c->stream>>>(planned.params, planned.maxUnits, planned.nBuffers);
        } else if (planned.unitsPerIter == 2) {
            dmavm_block_kernel<8, 2><<<planned.grid, 32 * 8, planned.smem, c->stream>>>(planned.params, planned.maxUnits, planned.nBuffers);
        } else {
            dmavm_block_kernel<8, 1><<<planned.grid, 32 * 8, planned.smem, c->stream>>>(planned.params, planned.maxUnits, planned.nBuffers);
        }
        CUDA_TRY(cudaGetLastError());
    }
    if (clocksEnv && planned.ws && count == 1) {
        std::vector<long long> h(static_cast<size_t>(planned.grid) * kComputeWarps * 4);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaMemcpy(h.data(), dClocks, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double w = 0, r = 0, t = 0, tiles = 0;
        for (size_t i = 0; i < h.size(); i += 4) {
            w += static_cast<double>(h[i]);
            r += static_cast<double>(h[i + 1]);
            t += static_cast<double>(h[i + 2]);
            tiles += static_cast<double>(h[i + 3]);
        }
        std::fprintf(stderr, "[flatdd_b200] clocks per warp and tile: wait %.0f, blocks %.0f, total %.0f (k=%d, units per warp and tile %.2f)\n", w / tiles, r / tiles,
                     t / tiles, planned.params.blocks[0].k, static_cast<double>(planned.params.blocks[0].nUnits) / kComputeWarps);
    }
    static const bool passLog = std::getenv("FLATDD_B200_PASSLOG") != nullptr; // experiments: one line per pass with its device time
    if (passLog && c->timing) {
        std::fprintf(stderr, "[flatdd_b200] pass blocks=%d tile=2^%d buffers=%d k=", count, planned.params.tileBits, planned.nBuffers);
        for (int i = 0; i < count; ++i) std::fprintf(stderr, "%d", planned.params.blocks[i].k);
        std::fprintf(stderr, " ctx=");
        for (int i = 0; i < count; ++i) std::fprintf(stderr, "%d,", planned.params.blocks[i].nCtx);
        std::fprintf(stderr, " ms=%.4f\n", c->lastMs);
    }
    c->launches++;
    c->blockLaunches++;
    c->blocksApplied += static_cast<uint64_t>(count);
    c->cur ^= 1;
    return true;
}

void walkWithTables(fdd_ctx* c, const fdd_gate* gate) {
    if (gate->dBlob == nullptr && gate->source != nullptr) {
        auto* g = const_cast<fdd_gate*>(gate);
        g->host = compileGate(*g->source, c->nLocal);
        uploadGate(g, c->stream);
    }
    launchWalk(c, gate);
}

// A block may move up over gates it commutes with to share a pass with an earlier block: two blocks commute when neither
// has a target among the other's targets or context qubits (a qubit both treat diagonally is harmless).  Greedy, in order:
// the first gate not yet placed opens a pass; later blocks (within a window) join it while the pass still fits one tile and
// they commute with every gate they jump over.  Gates that are not blocks are never crossed.  The product of the gates is
// unchanged; only the order of commuting factors (and with it the last bits of the rounding) differs.
std::vector<const fdd_gate*> orderForPasses(const fdd_ctx* c, const fdd_gate* const* gates, int count) {
    std::vector<const fdd_gate*> out;
    out.reserve(static_cast<size_t>(count));
    const int cap = std::max(1, std::min(c->blockMaxPerPass, kPassMaxBlocks));
    const int maxBits = std::min(c->nLocal, std::min(kPassMaxTileBits, c->blockMaxTileBits));
    constexpr int kWindow = 48;
    struct Masks {
        uint64_t targets = 0, support = 0;
        bool block = false;
    };
    std::vector<Masks> m(static_cast<size_t>(count));
    for (int i = 0; i < count; ++i) {
        if (!usesBlockPath(c, gates[i])) continue;
        Masks& x = m[static_cast<size_t>(i)];
        x.block = true;
        for (int q : gates[i]->block->targets) x.targets |= uint64_t{1} << q;
        x.support = x.targets;
        for (int q : gates[i]->block->ctx) x.support |= uint64_t{1} << q;
    }
    std::vector<char> placed(static_cast<size_t>(count), 0);
    for (int i = 0; i < count; ++i) {
        if (placed[static_cast<size_t>(i)]) continue;
        placed[static_cast<size_t>(i)] = 1;
        out.push_back(gates[i]);
        if (!m[static_cast<size_t>(i)].block || cap == 1) continue;
        std::vector<const DenseBlock*> group{gates[i]->block.get()};
        uint64_t skippedTargets = 0, skippedSupport = 0;
        for (int j = i + 1; j < count && j <= i + kWindow && static_cast<int>(group.size()) < cap; ++j) {
            if (placed[static_cast<size_t>(j)]) continue;
            const Masks& x = m[static_cast<size_t>(j)];
            if (!x.block) break;
            bool joins = (x.targets & skippedSupport) == 0 && (x.support & skippedTargets) == 0;
            if (joins) {
                group.push_back(gates[j]->block.get());
                const int need = minTileBits(group.data(), static_cast<int>(group.size()), c->nLocal);
                if (need < 0 || need > maxBits) {
                    group.pop_back();
                    joins = false;
                }
            }
            if (joins) {
                placed[static_cast<size_t>(j)] = 1;
                out.push_back(gates[j]);
            } else {
                skippedTargets |= x.targets;
                skippedSupport |= x.support;
            }
        }
    }
    return out;
}

// gates[0..count) in order: consecutive blocks share a pass while they fit one tile, everything else takes the older kernels
extern "C" void exchangeBits(fdd_ctx* c, int pg, int pl, int method); // (defined among the extern "C" entry points below)

void applyGates(fdd_ctx* c, const fdd_gate* const* gatesIn, int count, bool cachePlan, ExchangeSpec* ex = nullptr) {
    std::vector<const fdd_gate*> ordered;
    const fdd_gate* const* gates = gatesIn;
    if (c->blockReorder != 0 && count > 2) {
        ordered = orderForPasses(c, gatesIn, count);
        gates = ordered.data();
    }
    int i = 0;
    while (i < count) {
        if (!usesBlockPath(c, gates[i])) {
            walkWithTables(c, gates[i]);
            ++i;
            continue;
        }
        std::vector<const DenseBlock*> group{gates[i]->block.get()};
        int j = i + 1;
        const int cap = std::max(1, std::min(c->blockMaxPerPass, kPassMaxBlocks));
        while (j < count && j - i < cap && usesBlockPath(c, gates[j])) {
            group.push_back(gates[j]->block.get());
            const int need = minTileBits(group.data(), static_cast<int>(group.size()), c->nLocal);
            if (need < 0 || need > std::min(c->nLocal, std::min(kPassMaxTileBits, c->blockMaxTileBits))) {
                group.pop_back();
                break;
            }
            ++j;
        }
        // a group whose fragment shapes cannot be planned shrinks from the back; a single block that cannot falls back
        int n = j - i;
        // (the pass that ends the stretch takes the exchange that follows it along)
        while (n >= 1 && !launchPass(c, gates + i, n, cachePlan, i + n == count ? ex : nullptr)) --n;
        if (n == 0) {
            walkWithTables(c, gates[i]);
            n = 1;
        }
        i += n;
    }
    if (ex != nullptr && !ex->done) { // the stretch did not end in a pass that could take it: the exchange kernel
        exchangeBits(c, ex->pg, ex->pl, 0);
        ex->done = true;
    }
}

fdd_ctx* createCtx(int nQubits, int device, int rank, int world) {
    if (nQubits < 1 || nQubits > 40) throw std::invalid_argument("n_qubits must be in [1, 40]");
    if (world < 1 || world > kMaxPeers || (world & (world - 1)) != 0) throw std::invalid_argument("world_size must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= world) throw std::invalid_argument("rank out of range");
    int worldBits = 0;
    while ((1 << worldBits) < world) ++worldBits;
    if (nQubits - worldBits < 5 && world > 1) throw std::invalid_argument("a shard must hold at least 5 qubits");
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (count == 0) throw CudaError("no CUDA device: flatdd_b200 has no CPU fallback");
    if (device < 0 || device >= count) throw std::invalid_argument("device index out of range");
    auto ctx = new fdd_ctx();
    try {
        ctx->n = nQubits;
        ctx->rank = rank;
        ctx->world = world;
        ctx->worldBits = worldBits;
        ctx->nLocal = nQubits - worldBits;
        ctx->device = device;
        CUDA_TRY(cudaSetDevice(device));
        cudaDeviceProp prop{};
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        ctx->smCount = prop.multiProcessorCount;
        if (prop.major < 10) {
            throw CudaError(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                            "; this library is built for sm_100a only");
        }
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        {
            // gate tables come from the stream-ordered pool: keep freed blocks cached across synchronisations
            cudaMemPool_t pool = nullptr;
            CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t keep = ~uint64_t{0};
            CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        CUDA_TRY(cudaEventCreate(&ctx->ev0));
        CUDA_TRY(cudaEventCreate(&ctx->ev1));
        const size_t bytes = sizeof(double2) << ctx->nLocal;
        CUDA_TRY(cudaMalloc(&ctx->buf[0], bytes));
        CUDA_TRY(cudaMalloc(&ctx->buf[1], bytes));
        CUDA_TRY(cudaMalloc(&ctx->dPartial, sizeof(double) * 4096));
        CUDA_TRY(cudaMalloc(&ctx->dNorm, sizeof(double)));
        ctx->logicalToPhysical.resize(static_cast<size_t>(nQubits));
        for (int q = 0; q < nQubits; ++q) ctx->logicalToPhysical[static_cast<size_t>(q)] = q;
    } catch (...) {
        fdd_destroy(ctx);
        throw;
    }
    return ctx;
}

} // namespace

extern "C" {

const char* fdd_version(void) { return "flatdd_b200 0.1 (sm_100a)"; }
const char* fdd_last_error(void) { return g_lastError.c_str(); }

int fdd_device_count(int* count) {
    return guarded([&] {
        if (count == nullptr) throw std::invalid_argument("count is null");
        CUDA_TRY(cudaGetDeviceCount(count));
        if (*count == 0) throw CudaError("no CUDA device: flatdd_b200 has no CPU fallback");
    });
}

int fdd_create(int n_qubits, int device, fdd_ctx** out) { return fdd_create_sharded(n_qubits, device, 0, 1, out); }

int fdd_create_sharded(int n_qubits, int device, int rank, int world_size, fdd_ctx** out) {
    return guarded([&] {
        if (out == nullptr) throw std::invalid_argument("out is null");
        *out = nullptr;
        *out = createCtx(n_qubits, device, rank, world_size);
    });
}

int fdd_destroy(fdd_ctx* ctx) {
    if (ctx == nullptr) return FDD_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream != nullptr) cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < 2; ++b) {
        for (int r = 0; r < kMaxPeers; ++r) {
            if (ctx->peerBuf[b][r] != nullptr && r != ctx->rank) cudaIpcCloseMemHandle(const_cast<double2*>(ctx->peerBuf[b][r]));
        }
    }
    if (ctx->comm != nullptr) {
        try {
            NcclApi::get().CommDestroy(ctx->comm);
        } catch (...) {
        }
    }
    for (int r = 0; r < kMaxPeers; ++r) {
        if (ctx->peerFlags[r] != nullptr && r != ctx->rank) cudaIpcCloseMemHandle(const_cast<uint32_t*>(ctx->peerFlags[r]));
    }
    cudaFree(ctx->dFlags);
    cudaFree(ctx->dBarrier);
    cudaFree(ctx->buf[0]);
    cudaFree(ctx->buf[1]);
    cudaFree(ctx->dPartial);
    cudaFree(ctx->dNorm);
    if (ctx->ev0 != nullptr) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1 != nullptr) cudaEventDestroy(ctx->ev1);
    if (ctx->stream != nullptr) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FDD_OK;
}

int fdd_n_qubits(const fdd_ctx* ctx) { return ctx != nullptr ? ctx->n : 0; }
int fdd_n_local_qubits(const fdd_ctx* ctx) { return ctx != nullptr ? ctx->nLocal : 0; }

int fdd_synchronize(fdd_ctx* ctx) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        useDevice(ctx);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

int fdd_set_option(fdd_ctx* ctx, const char* key, long value) {
    return guarded([&] {
        if (ctx == nullptr || key == nullptr) throw std::invalid_argument("null argument");
        const std::string k = key;
        if (k == "dmavm_variant") ctx->variant = static_cast<int>(value);
        else if (k == "warps_per_cta") ctx->warpsPerCta = static_cast<int>(value);
        else if (k == "ctas_per_sm") ctx->ctasPerSm = static_cast<int>(value);
        else if (k == "prefetch") ctx->prefetch = static_cast<int>(value);
        else if (k == "tile_mode") ctx->forceMode = static_cast<int>(value);
        else if (k == "dense_slots") ctx->denseSlots = static_cast<int>(value);
        else if (k == "dmma") ctx->dmma = static_cast<int>(value);
        else if (k == "pdl") ctx->pdl = static_cast<int>(value);
        else if (k == "context_table") ctx->contextTable = static_cast<int>(value);
        else if (k == "flat_table") ctx->flatTable = static_cast<int>(value);
        else if (k == "exchange_unroll") ctx->exchangeUnroll = static_cast<int>(value);
        else if (k == "exchange_ctas_per_sm") ctx->exchangeCtasPerSm = static_cast<int>(value);
        else if (k == "block_kernel") ctx->blockKernel = static_cast<int>(value);
        else if (k == "block_tile_bits") ctx->blockTileBits = static_cast<int>(value);
        else if (k == "block_max_per_pass") ctx->blockMaxPerPass = static_cast<int>(value);
        else if (k == "block_buffers") ctx->blockBuffers = static_cast<int>(value);
        else if (k == "block_tables_shared") ctx->blockTablesShared = static_cast<int>(value);
        else if (k == "block_reorder") ctx->blockReorder = static_cast<int>(value);
        else if (k == "block_fuse_exchange") ctx->blockFuseExchange = static_cast<int>(value);
        else if (k == "block_warps") ctx->blockWarps = static_cast<int>(value);
        else if (k == "block_units") ctx->blockUnits = static_cast<int>(value);
        else if (k == "block_ws") ctx->blockWs = static_cast<int>(value);
        else if (k == "block_max_tile_bits") ctx->blockMaxTileBits = static_cast<int>(value);
        else if (k == "exchange_flags") ctx->exchangeFlags = static_cast<int>(value);
        else throw std::invalid_argument("unknown option " + k);
    });
}

int fdd_get_option(const fdd_ctx* ctx, const char* key, long* value) {
    return guarded([&] {
        if (ctx == nullptr || key == nullptr || value == nullptr) throw std::invalid_argument("null argument");
        const std::string k = key;
        if (k == "dmavm_variant") *value = ctx->variant;
        else if (k == "warps_per_cta") *value = ctx->warpsPerCta;
        else if (k == "ctas_per_sm") *value = ctx->ctasPerSm;
        else if (k == "prefetch") *value = ctx->prefetch;
        else if (k == "tile_mode") *value = ctx->forceMode;
        else if (k == "dense_slots") *value = ctx->denseSlots;
        else if (k == "dmma") *value = ctx->dmma;
        else if (k == "pdl") *value = ctx->pdl;
        else if (k == "context_table") *value = ctx->contextTable;
        else if (k == "context_table_launches") *value = static_cast<long>(ctx->contextLaunches);
        else if (k == "flat_table") *value = ctx->flatTable;
        else if (k == "flat_table_launches") *value = static_cast<long>(ctx->flatTableLaunches);
        else if (k == "exchange_unroll") *value = ctx->exchangeUnroll;
        else if (k == "exchange_ctas_per_sm") *value = ctx->exchangeCtasPerSm;
        else if (k == "launches") *value = static_cast<long>(ctx->launches);
        else if (k == "tensor_core_launches") *value = static_cast<long>(ctx->tensorCoreLaunches);
        else if (k == "exchanges") *value = static_cast<long>(ctx->exchanges);
        else if (k == "block_kernel") *value = ctx->blockKernel;
        else if (k == "block_tile_bits") *value = ctx->blockTileBits;
        else if (k == "block_max_per_pass") *value = ctx->blockMaxPerPass;
        else if (k == "block_buffers") *value = ctx->blockBuffers;
        else if (k == "block_tables_shared") *value = ctx->blockTablesShared;
        else if (k == "block_reorder") *value = ctx->blockReorder;
        else if (k == "block_fuse_exchange") *value = ctx->blockFuseExchange;
        else if (k == "fused_exchanges") *value = static_cast<long>(ctx->fusedExchanges);
        else if (k == "block_warps") *value = ctx->blockWarps;
        else if (k == "block_units") *value = ctx->blockUnits;
        else if (k == "block_ws") *value = ctx->blockWs;
        else if (k == "block_max_tile_bits") *value = ctx->blockMaxTileBits;
        else if (k == "exchange_flags") *value = ctx->exchangeFlags;
        else if (k == "block_launches") *value = static_cast<long>(ctx->blockLaunches);
        else if (k == "blocks_applied") *value = static_cast<long>(ctx->blocksApplied);
        else throw std::invalid_argument("unknown option " + k);
    });
}

#define NCCL_TRY(expr)                                                                                    \
    do {                                                                                                    \
        const ncclResult_t res__ = (expr);                                                                  \
        if (res__ != ncclSuccess) throw CommError(std::string(#expr) + ": " + NcclApi::get().GetErrorString(res__)); \
    } while (0)

int fdd_comm_unique_id(void* id128) {
    return guarded([&] {
        if (id128 == nullptr) throw std::invalid_argument("id128 is null");
        static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
        NCCL_TRY(NcclApi::get().GetUniqueId(static_cast<ncclUniqueId*>(id128)));
    });
}

int fdd_comm_init(fdd_ctx* ctx, const void* id128) {
    return guarded([&] {
        if (ctx == nullptr || id128 == nullptr) throw std::invalid_argument("null argument");
        if (ctx->world == 1) return;
        if (ctx->comm != nullptr) throw std::logic_error("communicator already initialised");
        useDevice(ctx);
        NcclApi& nccl = NcclApi::get();
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof id);
        NCCL_TRY(nccl.CommInitRank(&ctx->comm, ctx->world, id, ctx->rank));
        CUDA_TRY(cudaMalloc(&ctx->dBarrier, sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(ctx->dBarrier, 0, sizeof(double), ctx->stream));
        // trade CUDA IPC handles of both state buffers and of the flag words
        CUDA_TRY(cudaMalloc(&ctx->dFlags, 256));
        CUDA_TRY(cudaMemsetAsync(ctx->dFlags, 0, 256, ctx->stream));
        constexpr size_t kH = sizeof(cudaIpcMemHandle_t);
        constexpr int kHandles = 3;
        std::vector<unsigned char> mineH(kHandles * kH), all(kHandles * kH * static_cast<size_t>(ctx->world));
        for (int b = 0; b < kHandles; ++b) {
            cudaIpcMemHandle_t h;
            CUDA_TRY(cudaIpcGetMemHandle(&h, b < 2 ? static_cast<void*>(ctx->buf[b]) : static_cast<void*>(ctx->dFlags)));
            std::memcpy(mineH.data() + b * kH, &h, kH);
        }
        unsigned char* dH = nullptr;
        CUDA_TRY(cudaMalloc(&dH, all.size() + mineH.size()));
        CUDA_TRY(cudaMemcpyAsync(dH + all.size(), mineH.data(), mineH.size(), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(nccl.AllGather(dH + all.size(), dH, mineH.size(), ncclChar, ctx->comm, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(all.data(), dH, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFree(dH));
        for (int r = 0; r < ctx->world; ++r) {
            for (int b = 0; b < kHandles; ++b) {
                if (r == ctx->rank) {
                    if (b < 2) ctx->peerBuf[b][r] = ctx->buf[b];
                    else ctx->peerFlags[r] = ctx->dFlags;
                    continue;
                }
                cudaIpcMemHandle_t h;
                std::memcpy(&h, all.data() + (static_cast<size_t>(r) * kHandles + b) * kH, kH);
                void* p = nullptr;
                CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
                if (b < 2) ctx->peerBuf[b][r] = static_cast<const double2*>(p);
                else ctx->peerFlags[r] = static_cast<const uint32_t*>(p);
            }
        }
        // nobody may poll a flag word before its owner has zeroed it
        NCCL_TRY(nccl.AllReduce(ctx->dBarrier, ctx->dBarrier, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

namespace {

void streamBarrier(fdd_ctx* c) {
    // a one-element all-reduce on the context's stream: no rank passes it before every rank got here
    NCCL_TRY(NcclApi::get().AllReduce(c->dBarrier, c->dBarrier, 1, ncclDouble, ncclSum, c->comm, c->stream));
}

void swapLocalBits(fdd_ctx* c, int a, int b) {
    if (a == b) return;
    const uint64_t dim = c->localDim();
    swap_local_bits_kernel<<<gridFor(c, dim, 256), 256, 0, c->stream>>>(c->buf[c->cur], c->buf[c->cur ^ 1], dim, std::min(a, b), std::max(a, b));
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    c->cur ^= 1;
}

// physical SWAP(pg, pl) of one global and one local index bit
void exchangeBits(fdd_ctx* c, int pg, int pl, int method) {
    if (c->comm == nullptr) throw std::logic_error("shards are not connected: call fdd_comm_init first");
    const int gbit = pg - c->nLocal;
    const int partner = c->rank ^ (1 << gbit);
    const int myBit = (c->rank >> gbit) & 1;
    const uint64_t dim = c->localDim();
    if (method == 0 && c->exchangeFlags) {
        // ordering inside the kernel: flag words in peer memory (comm.cuh), no NCCL call
        const uint32_t epoch = ++c->exchangeEpoch;
        const int grid = c->smCount * (c->exchangeCtasPerSm > 0 ? c->exchangeCtasPerSm : 4);
        unsigned int* counter = reinterpret_cast<unsigned int*>(c->dFlags + 2);
        auto launch = [&](auto kernel) {
            kernel<<<grid, 256, 0, c->stream>>>(c->buf[c->cur], c->peerBuf[c->cur][partner], c->buf[c->cur ^ 1], dim, pl, myBit, c->dFlags,
                                                c->peerFlags[partner], epoch, counter);
        };
        if (c->exchangeUnroll == 4) launch(exchange_p2p_flag_kernel<4>);
        else if (c->exchangeUnroll == 16) launch(exchange_p2p_flag_kernel<16>);
        else launch(exchange_p2p_flag_kernel<8>);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        c->cur ^= 1;
    } else if (method == 0) {
        streamBarrier(c); // every rank has finished writing its current buffer
        const int grid = c->smCount * (c->exchangeCtasPerSm > 0 ? c->exchangeCtasPerSm : 4);
        if (c->exchangeUnroll == 4) {
            exchange_p2p_kernel<4><<<grid, 256, 0, c->stream>>>(c->buf[c->cur], c->peerBuf[c->cur][partner], c->buf[c->cur ^ 1], dim, pl, myBit);
        } else if (c->exchangeUnroll == 16) {
            exchange_p2p_kernel<16><<<grid, 256, 0, c->stream>>>(c->buf[c->cur], c->peerBuf[c->cur][partner], c->buf[c->cur ^ 1], dim, pl, myBit);
        } else {
            exchange_p2p_kernel<8><<<grid, 256, 0, c->stream>>>(c->buf[c->cur], c->peerBuf[c->cur][partner], c->buf[c->cur ^ 1], dim, pl, myBit);
        }
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        streamBarrier(c); // nobody overwrites a buffer a peer may still be reading
        c->cur ^= 1;
    } else {
        const int top = c->nLocal - 1;
        swapLocalBits(c, pl, top);
        const uint64_t half = dim >> 1;
        const uint64_t keepOff = myBit ? half : 0, sendOff = myBit ? 0 : half;
        NcclApi& nccl = NcclApi::get();
        NCCL_TRY(nccl.GroupStart());
        NCCL_TRY(nccl.Send(c->buf[c->cur] + sendOff, half * 2, ncclDouble, partner, c->comm, c->stream));
        NCCL_TRY(nccl.Recv(c->buf[c->cur ^ 1] + sendOff, half * 2, ncclDouble, partner, c->comm, c->stream));
        NCCL_TRY(nccl.GroupEnd());
        CUDA_TRY(cudaMemcpyAsync(c->buf[c->cur ^ 1] + keepOff, c->buf[c->cur] + keepOff, half * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
        c->cur ^= 1;
        swapLocalBits(c, pl, top);
    }
    c->exchanges++;
}

void swapPermutationEntries(fdd_ctx* c, int pa, int pb) {
    for (auto& p : c->logicalToPhysical) {
        if (p == pa) p = pb;
        else if (p == pb) p = pa;
    }
}

} // namespace

int fdd_exchange_qubits(fdd_ctx* ctx, int global_physical_bit, int local_physical_bit, int method) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        if (ctx->world == 1) throw std::logic_error("context is not sharded");
        if (global_physical_bit < ctx->nLocal || global_physical_bit >= ctx->n) throw std::invalid_argument("first argument is not a global physical bit");
        if (local_physical_bit < 0 || local_physical_bit >= ctx->nLocal) throw std::invalid_argument("second argument is not a local physical bit");
        if (!ctx->hasState) throw std::logic_error("no state");
        useDevice(ctx);
        Timed t(ctx);
        exchangeBits(ctx, global_physical_bit, local_physical_bit, method);
        swapPermutationEntries(ctx, global_physical_bit, local_physical_bit);
    });
}

int fdd_relabel_qubits(fdd_ctx* ctx, int physical_bit_a, int physical_bit_b) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        if (physical_bit_a < 0 || physical_bit_a >= ctx->n || physical_bit_b < 0 || physical_bit_b >= ctx->n) throw std::invalid_argument("physical bit out of range");
        swapPermutationEntries(ctx, physical_bit_a, physical_bit_b);
    });
}

int fdd_barrier(fdd_ctx* ctx) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        if (ctx->world == 1) return;
        if (ctx->comm == nullptr) throw std::logic_error("shards are not connected: call fdd_comm_init first");
        useDevice(ctx);
        streamBarrier(ctx);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

int fdd_convert(fdd_ctx* ctx, const fdd_vecdd* dd) {
    return guarded([&] {
        if (ctx == nullptr || dd == nullptr) throw std::invalid_argument("null argument");
        validate(*dd);
        if (dd->n_qubits != ctx->n) throw std::invalid_argument("vector DD has " + std::to_string(dd->n_qubits) + " qubits, context has " + std::to_string(ctx->n));
        useDevice(ctx);
        std::vector<VecNode> table(static_cast<size_t>(dd->n_nodes));
        for (int32_t u = 0; u < dd->n_nodes; ++u) {
            VecNode& nd = table[static_cast<size_t>(u)];
            nd.level = dd->level[u];
            nd.pad = 0;
            for (int b = 0; b < 2; ++b) {
                const size_t e = 2 * static_cast<size_t>(u) + static_cast<size_t>(b);
                const bool zero = dd->weight[2 * e] == 0.0 && dd->weight[2 * e + 1] == 0.0;
                nd.child[b] = zero ? FDD_TERMINAL : dd->child[e];
                nd.w[b] = zero ? make_double2(0.0, 0.0) : make_double2(dd->weight[2 * e], dd->weight[2 * e + 1]);
            }
        }
        VecNode* dTable = nullptr;
        const size_t tableBytes = table.size() * sizeof(VecNode);
        CUDA_TRY(cudaMallocAsync(&dTable, tableBytes, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(dTable, table.data(), tableBytes, cudaMemcpyHostToDevice, ctx->stream));
        ConvertParams p{};
        p.nodes = dTable;
        p.nNodes = dd->n_nodes;
        p.root = dd->root;
        p.rootW = make_double2(dd->root_weight[0], dd->root_weight[1]);
        p.nQubits = ctx->n;
        p.nLocal = ctx->nLocal;
        p.segBits = std::min(5, ctx->nLocal);
        p.rank = static_cast<uint32_t>(ctx->rank);
        p.nSeg = static_cast<uint32_t>(ctx->localDim() >> p.segBits);
        p.nTiles = (p.nSeg + 31) / 32;
        p.tableInSmem = tableBytes <= 160 * 1024 ? 1 : 0;
        p.out = ctx->buf[ctx->cur ^ 1];
        const size_t smem = p.tableInSmem ? tableBytes : 0;
        const int warps = 8;
        int perSm = smem == 0 ? 4 : static_cast<int>(std::max<size_t>(1, std::min<size_t>(4, kSmemBudget / smem)));
        const uint32_t ctasWanted = (p.nTiles + warps - 1) / warps;
        const int grid = static_cast<int>(std::max<uint32_t>(1, std::min<uint32_t>(ctasWanted, static_cast<uint32_t>(ctx->smCount * perSm))));
        {
            Timed t(ctx);
            CUDA_TRY(cudaFuncSetAttribute(convert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
            convert_kernel<<<grid, warps * 32, smem, ctx->stream>>>(p);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaFreeAsync(dTable, ctx->stream));
        ctx->launches++;
        ctx->cur ^= 1;
        ctx->hasState = true;
        for (int q = 0; q < ctx->n; ++q) ctx->logicalToPhysical[static_cast<size_t>(q)] = q;
    });
}

int fdd_gate_compile(fdd_ctx* ctx, const fdd_matdd* gate, fdd_gate** out) {
    return guarded([&] {
        if (ctx == nullptr || gate == nullptr || out == nullptr) throw std::invalid_argument("null argument");
        *out = nullptr;
        useDevice(ctx);
        auto g = new fdd_gate();
        try {
            g->host = compileGate(*gate, ctx->nLocal);
            g->device = ctx->device;
            uploadGate(g, ctx->stream);
            attachBlock(ctx, g, *gate);
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            freeGate(g, nullptr);
            throw;
        }
        *out = g;
    });
}

int fdd_gate_apply(fdd_ctx* ctx, const fdd_gate* gate) {
    return guarded([&] {
        if (ctx == nullptr || gate == nullptr) throw std::invalid_argument("null argument");
        useDevice(ctx);
        applyGates(ctx, &gate, 1, true);
    });
}

int fdd_gate_apply_many(fdd_ctx* ctx, const fdd_gate* const* gates, int count) {
    return guarded([&] {
        if (ctx == nullptr || (gates == nullptr && count > 0)) throw std::invalid_argument("null argument");
        useDevice(ctx);
        for (int i = 0; i < count; ++i) {
            if (gates[i] == nullptr) throw std::invalid_argument("null gate in the list");
        }
        applyGates(ctx, gates, count, true);
    });
}

static void checkExchangeArgs(const fdd_ctx* ctx, int global_physical_bit, int local_physical_bit) {
    if (ctx->world == 1) throw std::logic_error("context is not sharded");
    if (global_physical_bit < ctx->nLocal || global_physical_bit >= ctx->n) throw std::invalid_argument("not a global physical bit");
    if (local_physical_bit < 0 || local_physical_bit >= ctx->nLocal) throw std::invalid_argument("not a local physical bit");
    if (ctx->comm == nullptr) throw std::logic_error("shards are not connected: call fdd_comm_init first");
}

int fdd_gate_apply_many_exchange(fdd_ctx* ctx, const fdd_gate* const* gates, int count, int global_physical_bit, int local_physical_bit) {
    return guarded([&] {
        if (ctx == nullptr || (gates == nullptr && count > 0)) throw std::invalid_argument("null argument");
        checkExchangeArgs(ctx, global_physical_bit, local_physical_bit);
        useDevice(ctx);
        for (int i = 0; i < count; ++i) {
            if (gates[i] == nullptr) throw std::invalid_argument("null gate in the list");
        }
        ExchangeSpec ex{global_physical_bit, local_physical_bit, false};
        applyGates(ctx, gates, count, true, &ex);
        swapPermutationEntries(ctx, global_physical_bit, local_physical_bit);
    });
}

int fdd_gate_free(fdd_gate* gate) {
    freeGate(gate, nullptr);
    return FDD_OK;
}

static long gateFact(const CompiledGate& h, const std::string& k) {
    if (k == "kind") return h.diagonal ? 1 : 0;
    if (k == "max_paths") return h.maxPaths;
    if (k == "max_sub_k") return h.kTrue;
    if (k == "upper_nodes") return static_cast<long>(h.upper.size());
    if (k == "sub_tables") return h.nSub;
    if (k == "nnz_per_row_max") return h.nnzRowMax;
    if (k == "top_level") return h.topLevel;
    if (k == "upper_depth") return h.upperDepth;
    if (k == "stack_cap") return h.stackCap;
    if (k == "non_diag_mask") return static_cast<long>(h.nonDiagMask);
    if (k == "nnz") return static_cast<long>(h.nnz);
    if (k == "tileable") return h.tileable ? 1 : 0;
    if (k == "uniform") return h.uniform ? 1 : 0;
    if (k == "tile_mask") return static_cast<long>(h.tileMask);
    if (k == "fill_mask") return static_cast<long>(h.fillMask);
    if (k == "sub_tile_bits") return h.subTileBits;
    if (k == "non_diag_upper") return h.nonDiagUpper;
    if (k == "context_bits") return __builtin_popcount(h.ctxMask);
    return -1;
}

long fdd_gate_info(const fdd_gate* gate, const char* key) {
    if (gate == nullptr || key == nullptr) return -1;
    if (std::string(key) == "block_targets") return gate->block ? gate->block->k() : -1;
    if (std::string(key) == "block_context") return gate->block ? static_cast<long>(gate->block->ctx.size()) : -1;
    return gateFact(gate->host, key);
}

int fdd_matdd_info(const fdd_matdd* gate, const char* key, long* value) {
    return guarded([&] {
        if (gate == nullptr || key == nullptr || value == nullptr) throw std::invalid_argument("null argument");
        *value = gateFact(compileGate(*gate), key);
    });
}

static void applyManyHost(fdd_ctx* ctx, const fdd_matdd* gates, int count, ExchangeSpec* ex) {
    {
        if (ctx == nullptr || (gates == nullptr && count > 0)) throw std::invalid_argument("null argument");
        useDevice(ctx);
        std::vector<fdd_gate*> owned;
        double* arena = nullptr; // the matrix tables of all the call's blocks
        auto release = [&] {
            for (fdd_gate* g : owned) freeGate(g, ctx->stream); // stream ordered: released after the kernels have run
            if (arena != nullptr) cudaFreeAsync(arena, ctx->stream);
            arena = nullptr;
        };
        try {
            for (int i = 0; i < count; ++i) {
                if (gates[i].n_qubits != ctx->n) throw std::invalid_argument("gate has " + std::to_string(gates[i].n_qubits) + " qubits, context has " + std::to_string(ctx->n));
            }
            // the DD -> block expansion of every gate is host work (a quarter of a millisecond per fused gate of supremacy_n26):
            // on up to sixteen threads, so that the device does not wait for it
            std::vector<std::unique_ptr<DenseBlock>> blocks(static_cast<size_t>(count));
            const int nThreads = std::max(1, std::min({count / 3, 16, static_cast<int>(std::thread::hardware_concurrency())}));
            if (nThreads > 1) {
                std::vector<std::thread> pool;
                std::vector<std::exception_ptr> errors(static_cast<size_t>(nThreads));
                std::atomic<int> next{0};
                for (int t = 0; t < nThreads; ++t) {
                    pool.emplace_back([&, t] {
                        try {
                            for (int i = next.fetch_add(1); i < count; i = next.fetch_add(1)) blocks[static_cast<size_t>(i)] = extractBlock(ctx, gates[i]);
                        } catch (...) {
                            errors[static_cast<size_t>(t)] = std::current_exception();
                        }
                    });
                }
                for (auto& th : pool) th.join();
                for (auto& e : errors) {
                    if (e) std::rethrow_exception(e);
                }
            } else {
                for (int i = 0; i < count; ++i) blocks[static_cast<size_t>(i)] = extractBlock(ctx, gates[i]);
            }
            // one allocation and one upload for the matrix tables of the whole call (48 small cudaMallocAsync / cudaMemcpyAsync pairs
            // were a millisecond of host time in front of the first launch)
            std::vector<size_t> offset(static_cast<size_t>(count), 0);
            size_t total = 0;
            for (int i = 0; i < count; ++i) {
                if (!blocks[static_cast<size_t>(i)]) continue;
                offset[static_cast<size_t>(i)] = total;
                total += (blocks[static_cast<size_t>(i)]->table.size() * sizeof(double) + 255) & ~static_cast<size_t>(255);
            }
            if (total > 0) {
                std::vector<double> staging(total / sizeof(double), 0.0);
                for (int i = 0; i < count; ++i) {
                    const auto& b = blocks[static_cast<size_t>(i)];
                    if (b) std::memcpy(staging.data() + offset[static_cast<size_t>(i)] / sizeof(double), b->table.data(), b->table.size() * sizeof(double));
                }
                CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&arena), total, ctx->stream));
                CUDA_TRY(cudaMemcpyAsync(arena, staging.data(), total, cudaMemcpyHostToDevice, ctx->stream)); // pageable: staged before the call returns
            }
            for (int i = 0; i < count; ++i) {
                auto g = new fdd_gate();
                owned.push_back(g);
                g->device = ctx->device;
                g->source = &gates[i]; // the older kernels' tables are only made if a launch needs them
                if (blocks[static_cast<size_t>(i)]) {
                    g->dTable = arena + offset[static_cast<size_t>(i)] / sizeof(double);
                    g->tableBorrowed = true;
                    g->block = std::move(blocks[static_cast<size_t>(i)]);
                    g->serial = ++ctx->gateSerial;
                }
            }
            applyGates(ctx, owned.data(), count, false, ex);
        } catch (...) {
            release();
            throw;
        }
        release();
    }
}

int fdd_apply_many(fdd_ctx* ctx, const fdd_matdd* gates, int count) {
    return guarded([&] { applyManyHost(ctx, gates, count, nullptr); });
}

int fdd_apply_many_exchange(fdd_ctx* ctx, const fdd_matdd* gates, int count, int global_physical_bit, int local_physical_bit) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("null argument");
        checkExchangeArgs(ctx, global_physical_bit, local_physical_bit);
        ExchangeSpec ex{global_physical_bit, local_physical_bit, false};
        applyManyHost(ctx, gates, count, &ex);
        swapPermutationEntries(ctx, global_physical_bit, local_physical_bit);
    });
}

int fdd_apply(fdd_ctx* ctx, const fdd_matdd* gate) {
    if (gate == nullptr) return fail(FDD_ERR_INVALID, "null argument");
    return fdd_apply_many(ctx, gate, 1);
}

int fdd_block_from_matdd(const fdd_matdd* gate, int max_controls, int32_t* n_targets, int32_t* targets, int32_t* n_controls,
                         int32_t* controls, double* matrices, size_t capacity_doubles) {
    return guarded([&] {
        if (gate == nullptr || n_targets == nullptr || targets == nullptr || n_controls == nullptr || controls == nullptr) throw std::invalid_argument("null argument");
        *n_targets = -1;
        *n_controls = -1;
        DenseBlock b;
        if (!denseBlockFromDD(*gate, b, std::max(0, std::min(max_controls, kBlockMaxCtx)))) {
            throw std::length_error("the gate is not a dense block (more than 4 non-diagonal qubits or too many context qubits)");
        }
        *n_targets = b.k();
        *n_controls = static_cast<int32_t>(b.ctx.size());
        for (int i = 0; i < b.k(); ++i) targets[i] = b.targets[static_cast<size_t>(i)];
        for (size_t i = 0; i < b.ctx.size(); ++i) controls[i] = b.ctx[i];
        if (matrices == nullptr || capacity_doubles < b.table.size()) throw std::invalid_argument("matrices buffer too small: 2 * 4^n_targets * 2^n_controls doubles are needed");
        std::memcpy(matrices, b.table.data(), b.table.size() * sizeof(double));
    });
}

int fdd_ddarr_multiply(const fdd_matdd* gate, const double* y_real, const double* y_imag, double* z_real, double* z_imag,
                       size_t n_dim, int device) {
    if (gate == nullptr) return fail(FDD_ERR_INVALID, "gate is null");
    if (n_dim != (size_t{1} << gate->n_qubits)) return fail(FDD_ERR_INVALID, "n_dim != 2^n_qubits");
    fdd_ctx* ctx = nullptr;
    int rc = fdd_create(gate->n_qubits, device, &ctx);
    if (rc != FDD_OK) return rc;
    rc = fdd_set_state(ctx, y_real, y_imag);
    if (rc == FDD_OK) rc = fdd_apply(ctx, gate);
    if (rc == FDD_OK) rc = fdd_get_state(ctx, z_real, z_imag);
    fdd_destroy(ctx);
    return rc;
}

int fdd_mac_count(const fdd_matdd* gate, uint64_t* nnz) {
    return guarded([&] {
        if (gate == nullptr || nnz == nullptr) throw std::invalid_argument("null argument");
        validate(*gate);
        *nnz = macCount(*gate);
    });
}
int fdd_cost_ip(const fdd_matdd* gate, unsigned n_thread_exp, uint64_t* cost) {
    return guarded([&] {
        if (gate == nullptr || cost == nullptr) throw std::invalid_argument("null argument");
        validate(*gate);
        *cost = costIP(*gate, n_thread_exp);
    });
}
int fdd_cost_op1(const fdd_matdd* gate, unsigned n_thread_exp, uint64_t* cost) {
    return guarded([&] {
        if (gate == nullptr || cost == nullptr) throw std::invalid_argument("null argument");
        validate(*gate);
        if (static_cast<int>(n_thread_exp) >= gate->n_qubits) throw std::invalid_argument("n_thread_exp must be below n_qubits");
        *cost = costOP1(*gate, n_thread_exp);
    });
}
int fdd_cost_gpu(const fdd_matdd* gate, double hbm_gbs, double fp64_gflops, double* nanoseconds) {
    return guarded([&] {
        if (gate == nullptr || nanoseconds == nullptr) throw std::invalid_argument("null argument");
        const CompiledGate c = compileGate(*gate);
        *nanoseconds = costGpuNs(c, hbm_gbs, fp64_gflops);
    });
}

int fdd_get_state(fdd_ctx* ctx, double* real, double* imag) {
    return guarded([&] {
        if (ctx == nullptr || real == nullptr || imag == nullptr) throw std::invalid_argument("null argument");
        if (!ctx->hasState) throw std::logic_error("no state");
        useDevice(ctx);
        const uint64_t dim = ctx->localDim();
        // de-interleave into the idle ping-pong buffer, then two planar copies
        auto* planar = reinterpret_cast<double*>(ctx->buf[ctx->cur ^ 1]);
        deinterleave_kernel<<<gridFor(ctx, dim, 256), 256, 0, ctx->stream>>>(ctx->buf[ctx->cur], planar, planar + dim, dim);
        CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        CUDA_TRY(cudaMemcpyAsync(real, planar, dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(imag, planar + dim, dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

int fdd_set_state(fdd_ctx* ctx, const double* real, const double* imag) {
    return guarded([&] {
        if (ctx == nullptr || real == nullptr || imag == nullptr) throw std::invalid_argument("null argument");
        useDevice(ctx);
        const uint64_t dim = ctx->localDim();
        auto* planar = reinterpret_cast<double*>(ctx->buf[ctx->cur]);
        CUDA_TRY(cudaMemcpyAsync(planar, real, dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(planar + dim, imag, dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        interleave_kernel<<<gridFor(ctx, dim, 256), 256, 0, ctx->stream>>>(planar, planar + dim, ctx->buf[ctx->cur ^ 1], dim);
        CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        ctx->cur ^= 1;
        ctx->hasState = true;
        for (int q = 0; q < ctx->n; ++q) ctx->logicalToPhysical[static_cast<size_t>(q)] = q;
    });
}

int fdd_set_zero_state(fdd_ctx* ctx) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        useDevice(ctx);
        const uint64_t dim = ctx->localDim();
        // written into the idle buffer and flipped like every other initialiser, so the ping-pong parity of the ranks of a
        // sharded state never depends on how each one was initialised (the exchange kernel reads the partner's buf[cur])
        zero_state_kernel<<<gridFor(ctx, dim, 256), 256, 0, ctx->stream>>>(ctx->buf[ctx->cur ^ 1], dim, ctx->rank == 0 ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        ctx->cur ^= 1;
        ctx->hasState = true;
        for (int q = 0; q < ctx->n; ++q) ctx->logicalToPhysical[static_cast<size_t>(q)] = q;
    });
}

int fdd_get_amplitudes(fdd_ctx* ctx, uint64_t first, uint64_t count, double* interleaved) {
    return guarded([&] {
        if (ctx == nullptr || interleaved == nullptr) throw std::invalid_argument("null argument");
        if (!ctx->hasState) throw std::logic_error("no state");
        if (first + count > ctx->localDim()) throw std::invalid_argument("amplitude range out of bounds");
        useDevice(ctx);
        CUDA_TRY(cudaMemcpyAsync(interleaved, ctx->buf[ctx->cur] + first, count * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

int fdd_get_amplitudes_at(fdd_ctx* ctx, const uint64_t* local_indices, uint64_t count, double* interleaved) {
    return guarded([&] {
        if (ctx == nullptr || (count > 0 && (local_indices == nullptr || interleaved == nullptr))) throw std::invalid_argument("null argument");
        if (!ctx->hasState) throw std::logic_error("no state");
        if (count == 0) return;
        for (uint64_t i = 0; i < count; ++i) {
            if (local_indices[i] >= ctx->localDim()) throw std::invalid_argument("amplitude index out of bounds");
        }
        useDevice(ctx);
        uint64_t* dIdx = nullptr;
        double2* dOut = nullptr;
        CUDA_TRY(cudaMallocAsync(&dIdx, sizeof(uint64_t) * count, ctx->stream));
        CUDA_TRY(cudaMallocAsync(&dOut, sizeof(double2) * count, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(dIdx, local_indices, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, ctx->stream));
        gather_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->buf[ctx->cur], dIdx, count, dOut);
        CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        CUDA_TRY(cudaMemcpyAsync(interleaved, dOut, sizeof(double2) * count, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFreeAsync(dIdx, ctx->stream));
        CUDA_TRY(cudaFreeAsync(dOut, ctx->stream));
    });
}

int fdd_norm2(fdd_ctx* ctx, double* out) {
    return guarded([&] {
        if (ctx == nullptr || out == nullptr) throw std::invalid_argument("null argument");
        if (!ctx->hasState) throw std::logic_error("no state");
        useDevice(ctx);
        const uint64_t dim = ctx->localDim();
        const int grid = std::min(4096, gridFor(ctx, dim, 256));
        norm2_partial_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->buf[ctx->cur], dim, ctx->dPartial);
        norm2_final_kernel<<<1, 256, 0, ctx->stream>>>(ctx->dPartial, grid, ctx->dNorm);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 2;
        CUDA_TRY(cudaMemcpyAsync(out, ctx->dNorm, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    });
}

int fdd_sample(fdd_ctx* ctx, uint64_t n_shots, uint64_t seed, uint64_t* local_indices) {
    return guarded([&] {
        if (ctx == nullptr || (local_indices == nullptr && n_shots > 0)) throw std::invalid_argument("null argument");
        if (!ctx->hasState) throw std::logic_error("no state");
        if (n_shots == 0) return;
        useDevice(ctx);
        const uint64_t dim = ctx->localDim();
        const uint32_t blockAmps = static_cast<uint32_t>(std::min<uint64_t>(dim, 4096));
        const uint32_t nBlocks = static_cast<uint32_t>((dim + blockAmps - 1) / blockAmps);
        double* dMass = nullptr;
        CUDA_TRY(cudaMallocAsync(&dMass, sizeof(double) * nBlocks, ctx->stream));
        block_mass_kernel<<<nBlocks, 256, 0, ctx->stream>>>(ctx->buf[ctx->cur], dim, blockAmps, dMass);
        CUDA_TRY(cudaGetLastError());
        std::vector<double> mass(nBlocks);
        CUDA_TRY(cudaMemcpyAsync(mass.data(), dMass, sizeof(double) * nBlocks, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFreeAsync(dMass, ctx->stream));
        std::vector<double> prefix(nBlocks + 1, 0.0);
        for (uint32_t b = 0; b < nBlocks; ++b) prefix[b + 1] = prefix[b] + mass[b];
        const double total = prefix[nBlocks];
        if (!(total > 0.0)) throw std::logic_error("the state has zero norm");
        // splitmix64: deterministic, seedable, good enough for shot sampling
        uint64_t x = seed;
        auto next = [&x]() {
            uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            return z ^ (z >> 31);
        };
        std::vector<uint32_t> shotBlock(n_shots);
        std::vector<double> shotResidual(n_shots);
        for (uint64_t sIdx = 0; sIdx < n_shots; ++sIdx) {
            const double u = static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0) * total;
            uint32_t b = static_cast<uint32_t>(std::upper_bound(prefix.begin() + 1, prefix.end(), u) - (prefix.begin() + 1));
            if (b >= nBlocks) b = nBlocks - 1;
            while (b > 0 && mass[b] == 0.0) --b; // never land in an empty block
            shotBlock[sIdx] = b;
            shotResidual[sIdx] = u - prefix[b];
        }
        uint32_t* dBlock = nullptr;
        double* dRes = nullptr;
        uint64_t* dOut = nullptr;
        CUDA_TRY(cudaMallocAsync(&dBlock, sizeof(uint32_t) * n_shots, ctx->stream));
        CUDA_TRY(cudaMallocAsync(&dRes, sizeof(double) * n_shots, ctx->stream));
        CUDA_TRY(cudaMallocAsync(&dOut, sizeof(uint64_t) * n_shots, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(dBlock, shotBlock.data(), sizeof(uint32_t) * n_shots, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(dRes, shotResidual.data(), sizeof(double) * n_shots, cudaMemcpyHostToDevice, ctx->stream));
        const uint64_t threads = n_shots * 32;
        sample_resolve_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, ctx->stream>>>(ctx->buf[ctx->cur], dim, blockAmps, dBlock, dRes, n_shots, dOut);
        CUDA_TRY(cudaGetLastError());
        ctx->launches += 2;
        CUDA_TRY(cudaMemcpyAsync(local_indices, dOut, sizeof(uint64_t) * n_shots, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFreeAsync(dBlock, ctx->stream));
        CUDA_TRY(cudaFreeAsync(dRes, ctx->stream));
        CUDA_TRY(cudaFreeAsync(dOut, ctx->stream));
    });
}

int fdd_state_device_ptr(fdd_ctx* ctx, void** ptr) {
    if (ctx == nullptr || ptr == nullptr) return fail(FDD_ERR_INVALID, "null argument");
    *ptr = ctx->buf[ctx->cur];
    return FDD_OK;
}

int fdd_get_permutation(const fdd_ctx* ctx, int32_t* logical_to_physical) {
    if (ctx == nullptr || logical_to_physical == nullptr) return fail(FDD_ERR_INVALID, "null argument");
    std::copy(ctx->logicalToPhysical.begin(), ctx->logicalToPhysical.end(), logical_to_physical);
    return FDD_OK;
}

int fdd_canonicalize(fdd_ctx* ctx) {
    return guarded([&] {
        if (ctx == nullptr) throw std::invalid_argument("ctx is null");
        useDevice(ctx);
        auto& l2p = ctx->logicalToPhysical;
        // bring logical qubit p to physical position p, highest position first
        for (int p = ctx->n - 1; p >= 0; --p) {
            const int c = l2p[static_cast<size_t>(p)]; // where logical p sits now
            if (c == p) continue;
            const bool pGlobal = p >= ctx->nLocal, cGlobal = c >= ctx->nLocal;
            if (!pGlobal && !cGlobal) {
                swapLocalBits(ctx, p, c);
            } else if (pGlobal != cGlobal) {
                exchangeBits(ctx, pGlobal ? p : c, pGlobal ? c : p, 0);
            } else {
                // both global: route through local bit 0 (three exchanges)
                exchangeBits(ctx, p, 0, 0);
                exchangeBits(ctx, c, 0, 0);
                exchangeBits(ctx, p, 0, 0);
            }
            swapPermutationEntries(ctx, p, c);
        }
    });
}

int fdd_last_kernel_ms(fdd_ctx* ctx, float* ms) {
    if (ctx == nullptr || ms == nullptr) return fail(FDD_ERR_INVALID, "null argument");
    *ms = ctx->lastMs;
    return FDD_OK;
}

int fdd_set_timing(fdd_ctx* ctx, int enabled) {
    if (ctx == nullptr) return fail(FDD_ERR_INVALID, "ctx is null");
    ctx->timing = enabled != 0;
    return FDD_OK;
}

uint64_t fdd_launch_count(const fdd_ctx* ctx) { return ctx != nullptr ? ctx->launches : 0; }

int fdd_stream(fdd_ctx* ctx, void** stream) {
    if (ctx == nullptr || stream == nullptr) return fail(FDD_ERR_INVALID, "null argument");
    *stream = ctx->stream;
    return FDD_OK;
}

} // extern "C"

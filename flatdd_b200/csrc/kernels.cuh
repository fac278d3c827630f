// kernels.cuh — sm_100a device code of the FlatDD array phase.
//
// Work decomposition shared by both hot kernels.  The amplitude index is split into a
// SEGMENT (the low S = min(5, n) bits, one warp lane per amplitude, 512 contiguous bytes) and
// the segment index (the remaining upper bits).  A warp owns TILES of 32 consecutive segments
// (16 KiB of output) and works in two phases:
//   phase A  lane j walks the UPPER levels of the decision diagram for segment 32*tile + j
//            (each lane a different segment), leaving in shared memory what the segment needs;
//   phase B  the warp sweeps the 32 segments one after the other, lane = amplitude, so every
//            global access is a full 512-byte coalesced request of 16-byte vectors.
// The grid is persistent (a multiple of the SM count) and tiles are dealt round-robin to warps
// so neighbouring warps stream neighbouring DRAM pages.
//
// All arithmetic is IEEE fp64.  The conversion kernel uses un-fused multiplies and adds in the
// reference's order (bit-identical results); DMAVM uses FMAs, and FP64 tensor-core tiles (DMMA.8x8x4,
// three real products per complex product) where a fused block is a real dense contraction on 3 or 4
// upper qubits (tolerance 1e-10, measured <= 1e-13 against the reference; see DESIGN.md sections 3.2, 5).
#pragma once
#include <climits>

#include "gate_compile.hpp"

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

// experiment switches of the tensor-core path (tile kernel MODE 5)
#ifndef FDD_M5_3M
#define FDD_M5_3M 1   // 1: three real DMMA products per complex product instead of four (measured 0.426 -> 0.417 ms at n = 26)
#endif
#ifndef FDD_M6_REGS
#define FDD_M6_REGS 1 // flat-table path: weights of small tables in registers
#endif
#ifndef FDD_M5_STAGES
#define FDD_M5_STAGES 2 // stage slots (sub-tiles) per warp
#endif
#ifndef FDD_M5_MAXWARPS
#define FDD_M5_MAXWARPS 12 // launch bounds: warps per CTA ... (one wide CTA shares one entry area: 12 resident warps instead of 10)
#endif
#ifndef FDD_M5_MINCTAS
#define FDD_M5_MINCTAS 1 // ... and resident CTAs per SM
#endif

namespace fddb200 {

constexpr int kWarp = 32;
constexpr int kMaxPeers = 8;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cmac(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// un-fused complex multiply in the reference's operation order (include/dd/SwitchPackage.hpp:3616-3619)
__device__ __forceinline__ double2 cmul_exact(double2 c, double2 w) {
    return make_double2(__dsub_rn(__dmul_rn(c.x, w.x), __dmul_rn(c.y, w.y)), __dadd_rn(__dmul_rn(c.x, w.y), __dmul_rn(c.y, w.x)));
}
__device__ __forceinline__ double2 shfl2(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ void cp_async16(void* smemDst, const void* gmemSrc) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smemDst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// system-scope release / acquire of a flag word in peer-mapped memory (cross-GPU ordering of the exchanges, comm.cuh)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 32-byte streaming store (STG.256): two neighbouring amplitudes
__device__ __forceinline__ void st_stream2(double2* p, double2 a, double2 b) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  Fragments (lane l): A[l>>2][l&3], B[l&3][l>>2],
// D[l>>2][2*(l&3) + {0,1}].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------------
// DD -> array conversion
// ------------------------------------------------------------------------------------------------
struct alignas(16) VecNode {
    int32_t child[2]; // successor node or FDD_TERMINAL
    int32_t level;
    int32_t pad;
    double2 w[2];     // exact (0,0) marks a zero edge
};
static_assert(sizeof(VecNode) == 48, "VecNode layout");

struct ConvertParams {
    const VecNode* nodes; // global memory copy of the table
    int nNodes;
    int root;
    double2 rootW;
    int nQubits;  // all qubits (levels nQubits-1 .. 0)
    int nLocal;   // qubits held by this shard
    int segBits;  // S
    uint32_t rank; // value of the global (top) index bits of this shard
    uint32_t nSeg;  // local segments
    uint32_t nTiles;
    int tableInSmem;
    double2* out;
};

// One value of the expansion of a segment's sub-DD: the product so far, the node it has reached, "a zero weight was met".
constexpr int kConvertZero = INT_MIN; // ConvertPath::node of a path that has met a zero weight
struct ConvertPath {
    double2 a;
    int node; // the node reached, or kConvertZero (one shuffle carries both facts)
};
// One step down the DD for every lane: the parent's value comes from lane `src`, the edge taken is `bit`.
__device__ __forceinline__ ConvertPath convertStep(const VecNode* __restrict__ nodes, const ConvertPath& parent, int src, int bit) {
    ConvertPath r;
    r.a = shfl2(parent.a, src);
    r.node = __shfl_sync(0xffffffffu, parent.node, src);
    if (r.node != kConvertZero) {
        const VecNode& nd = nodes[r.node];
        const double2 w = nd.w[bit];
        r.a = cmul_exact(r.a, w);
        r.node = (w.x == 0.0 && w.y == 0.0) ? kConvertZero : nd.child[bit];
    }
    return r;
}

// amplitude(i) = w_root * prod_v w(node_v.e[bit_v(i)]), multiplied root first, leaf last
// (reference getValueByPathPar, include/dd/SwitchPackage.hpp:3605-3634).
__global__ void __launch_bounds__(256) convert_kernel(const ConvertParams p) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const VecNode* nodes = p.nodes;
    if (p.tableInSmem) {
        VecNode* sn = reinterpret_cast<VecNode*>(smemRaw);
        const int4* src = reinterpret_cast<const int4*>(p.nodes);
        int4* dst = reinterpret_cast<int4*>(sn);
        for (int i = threadIdx.x; i < p.nNodes * 3; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        nodes = sn;
    }
    const int lane = threadIdx.x & 31;
    const int warpsPerCta = blockDim.x >> 5;
    const uint32_t warpGlobal = blockIdx.x * warpsPerCta + (threadIdx.x >> 5);
    const uint32_t warpStride = gridDim.x * warpsPerCta;
    const int S = p.segBits;
    const int segLen = 1 << S;
    const int upperLocalBits = p.nLocal - S; // bits of the local segment index

    for (uint32_t tile = warpGlobal; tile < p.nTiles; tile += warpStride) {
        // ---- phase A: one segment prefix per lane --------------------------------------------
        const uint32_t seg = tile * 32u + lane;
        int node = -1;
        double2 c = make_double2(1.0, 0.0);
        bool dead = true;
        if (seg < p.nSeg) {
            const uint64_t rowSeg = (static_cast<uint64_t>(p.rank) << upperLocalBits) | seg;
            dead = false;
            node = p.root;
            double2 w = p.rootW;
            for (int lv = p.nQubits - 1; lv >= S; --lv) {
                c = cmul_exact(c, w);
                if (w.x == 0.0 && w.y == 0.0) {
                    dead = true;
                    break;
                }
                const VecNode& nd = nodes[node];
                const int b = static_cast<int>((rowSeg >> (lv - S)) & 1ULL);
                w = nd.w[b];
                node = nd.child[b];
            }
            if (!dead) {
                // the weight of the edge INTO the level S-1 node is still pending
                c = cmul_exact(c, w);
                dead = (w.x == 0.0 && w.y == 0.0);
            }
        }
        const int nSegTile = min(32u, p.nSeg - tile * 32u);
        // ---- phase B, full tiles of 32-amplitude segments: the sub-DD below every segment is expanded as a TREE ----------
        // The 32 amplitudes of a segment share their partial products: 2 distinct values after level 4, 4 after level 3, ...
        // Step m (m = 1..5 bits decided) works on 32 >> m segments at once, lane = (segment slot << m) | (decided bits,
        // level 4 first), and takes its parent from the step above by shuffle; depth first, sixteen segments per round:
        // 1 + 2 + 4 + 8 + 16 = 31 steps for 16 segments instead of 5 per segment (2.6 times fewer table look-ups and
        // multiplications; the kernel was bound by the shared-memory look-ups).  Every amplitude is still
        // ((((c w4) w3) w2) w1) w0 with the same un-fused arithmetic: bit-identical to the reference.  After step 5 the lane
        // index is the amplitude index inside the segment: 512-byte coalesced stores.
        if (S == 5 && nSegTile == 32) {
            ConvertPath top;
            top.a = c;
            top.node = dead ? kConvertZero : node;
            double2* outTile = p.out + (static_cast<uint64_t>(tile) << 10);
#pragma unroll 1
            for (int round = 0; round < 2; ++round) {
                const ConvertPath s1 = convertStep(nodes, top, 16 * round + (lane >> 1), lane & 1);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const ConvertPath s2 = convertStep(nodes, s1, ((h2 * 8 + (lane >> 2)) << 1) | ((lane & 3) >> 1), lane & 1);
#pragma unroll
                    for (int h3 = 0; h3 < 2; ++h3) {
                        const ConvertPath s3 = convertStep(nodes, s2, ((h3 * 4 + (lane >> 3)) << 2) | ((lane & 7) >> 1), lane & 1);
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            const ConvertPath s4 = convertStep(nodes, s3, ((h4 * 2 + (lane >> 4)) << 3) | ((lane & 15) >> 1), lane & 1);
#pragma unroll
                            for (int h5 = 0; h5 < 2; ++h5) {
                                const ConvertPath s5 = convertStep(nodes, s4, (h5 << 4) | (lane >> 1), lane & 1);
                                const int seg16 = h2 * 8 + h3 * 4 + h4 * 2 + h5;
                                st_stream(outTile + ((16 * round + seg16) << 5) + lane, s5.node == kConvertZero ? make_double2(0.0, 0.0) : s5.a);
                            }
                        }
                    }
                }
            }
            continue;
        }
        // ---- phase B, ragged tiles and short segments: sweep the segments, lane = amplitude ------------------------------
        // four segments at a time: four independent multiply chains per lane hide the latency of the
        // dependent table look-ups (the arithmetic of every chain is unchanged)
        constexpr int U = 4;
        for (int j0 = 0; j0 < nSegTile; j0 += U) {
            double2 a[U];
            int u[U];
            bool z[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = min(j0 + k, nSegTile - 1);
                u[k] = __shfl_sync(0xffffffffu, node, j);
                a[k] = shfl2(c, j);
                z[k] = __shfl_sync(0xffffffffu, static_cast<int>(dead), j) != 0;
            }
            if (lane < segLen) {
                for (int lv = S - 1; lv >= 0; --lv) {
                    const int b = (lane >> lv) & 1;
#pragma unroll
                    for (int k = 0; k < U; ++k) {
                        if (!z[k]) {
                            const VecNode& nd = nodes[u[k]];
                            const double2 w = nd.w[b];
                            a[k] = cmul_exact(a[k], w);
                            if (w.x == 0.0 && w.y == 0.0) {
                                z[k] = true;
                            } else {
                                u[k] = nd.child[b];
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    if (j0 + k < nSegTile) {
                        st_stream(p.out + ((static_cast<uint64_t>(tile) * 32u + j0 + k) << S) + lane, z[k] ? make_double2(0.0, 0.0) : a[k]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// DMAVM walk kernel
// ------------------------------------------------------------------------------------------------
struct WalkParams {
    const double2* y;                 // source state (this shard)
    const double2* peerY[kMaxPeers];  // source shards of all ranks (peerY[rank] == y); multi-GPU only
    double2* z;                       // destination state (this shard)
    const UpperNode* upper;
    int nUpper;
    const double2* subW;              // [nSub][kMax][32]
    const uint8_t* subCol;            // [nSub][kMax][32]
    const int32_t* subK;              // [nSub]
    const uint8_t* subFlags;          // [nSub]
    int nSub;
    int kMax;
    int root;                         // encoded like UpperNode::child
    double2 rootW;
    int nLocal;
    int segBits;
    int maxPaths;
    int stackCap;
    uint32_t rank;
    int worldBits;
    uint32_t nSeg;
    uint32_t nTiles;
    int tablesInSmem;
    int prefetch;                     // ring depth D (variant 1)
    int tileBits;                     // tile kernel: log2(segments per warp tile) (<= 5)
    int subTileBits;                  // tile kernel: TB = number of bits in tileMask
    uint32_t tileMask;                // tile kernel: segment-index bits of the non-diagonal upper levels
    uint32_t fillMask;                // tile kernel: further index bits that complete the warp tile
    // tile kernel MODE 4: entry lists kept per sub table; list s occupies slots [subBase[s], subBase[s+1])
    uint8_t subBase[9];
    int uniform;                      // tile kernel: the entry lists do not depend on the tile (walk once per warp)
    int denseSlots;                   // tile kernel: entries are stored by source slot (T per row, zero weight when absent)
    // tensor-core path, gate not uniform: the block depends on the index bits `ctxMask` (diagonal upper levels
    // outside the tile: controls, phases).  ctxPhase 1 = walk every value of those bits once and store the
    // T x T blocks in ctxTable; ctxPhase 2 = the DMAVM launch looks its block up instead of walking per tile.
    double2* ctxTable;                // [nCtx][T columns][T rows]
    uint32_t ctxMask;
    int nCtx;
    int ctxPhase;
};

// Bytes of shared memory one warp needs.
__host__ __device__ inline size_t walkWarpSmem(int maxPaths, int stackCap, int prefetch) {
    return static_cast<size_t>(maxPaths + stackCap) * 32 * 24 + static_cast<size_t>(prefetch) * 512;
}
__host__ __device__ inline size_t walkTableSmem(int nUpper, int nSub, int kMax) {
    size_t b = static_cast<size_t>(nUpper) * sizeof(UpperNode);
    b += static_cast<size_t>(nSub) * kMax * 32 * 16; // subW
    b += static_cast<size_t>(nSub) * kMax * 32;      // subCol
    b += static_cast<size_t>(nSub) * 4;              // subK
    b += static_cast<size_t>(nSub);                  // subFlags
    return (b + 15) & ~static_cast<size_t>(15);
}

// z[r] = sum_c M[r][c] y[c]   (reference DDArrMultiplyIP, include/dd/SwitchPackage.hpp:1897-2261)
//
// VARIANT 0: source segments are loaded straight into registers (one LDG.128 per lane) and
//            permuted with warp shuffles.
// VARIANT 1: source segments are staged through a per-warp ring of `prefetch` 512-byte slots
//            in shared memory filled by cp.async, so several DRAM requests per warp are in flight
//            while earlier segments are consumed; the permutation is a shared-memory gather.
template <int VARIANT> __global__ void __launch_bounds__(512) dmavm_walk_kernel(const WalkParams p) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    unsigned char* cursor = smemRaw;
    const UpperNode* upper = p.upper;
    const double2* subW = p.subW;
    const uint8_t* subCol = p.subCol;
    const int32_t* subK = p.subK;
    const uint8_t* subFlags = p.subFlags;
    if (p.tablesInSmem) {
        // stage the gate tables (all sizes are multiples of 4 bytes except the byte tables)
        UpperNode* su = reinterpret_cast<UpperNode*>(cursor);
        cursor += static_cast<size_t>(p.nUpper) * sizeof(UpperNode);
        double2* sw = reinterpret_cast<double2*>(cursor);
        const int nEnt = p.nSub * p.kMax * 32;
        cursor += static_cast<size_t>(nEnt) * 16;
        int32_t* sk = reinterpret_cast<int32_t*>(cursor);
        cursor += static_cast<size_t>(p.nSub) * 4;
        uint8_t* sc = cursor;
        cursor += nEnt;
        uint8_t* sf = cursor;
        cursor += p.nSub;
        {
            const int4* src = reinterpret_cast<const int4*>(p.upper);
            int4* dst = reinterpret_cast<int4*>(su);
            for (int i = threadIdx.x; i < p.nUpper * 6; i += blockDim.x) dst[i] = src[i];
        }
        for (int i = threadIdx.x; i < nEnt; i += blockDim.x) {
            sw[i] = p.subW[i];
            sc[i] = p.subCol[i];
        }
        for (int i = threadIdx.x; i < p.nSub; i += blockDim.x) {
            sk[i] = p.subK[i];
            sf[i] = p.subFlags[i];
        }
        __syncthreads();
        upper = su;
        subW = sw;
        subCol = sc;
        subK = sk;
        subFlags = sf;
        cursor = smemRaw + walkTableSmem(p.nUpper, p.nSub, p.kMax);
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warpsPerCta = blockDim.x >> 5;
    const int D = p.prefetch;
    unsigned char* mine = cursor + static_cast<size_t>(warp) * walkWarpSmem(p.maxPaths, p.stackCap, VARIANT == 1 ? D : 0);
    // per-warp arrays, [slot][lane]
    const int nSlots = p.maxPaths + p.stackCap;
    double2* eW = reinterpret_cast<double2*>(mine);                               // [nSlots][32]
    int32_t* eCode = reinterpret_cast<int32_t*>(mine + static_cast<size_t>(nSlots) * 32 * 16); // [nSlots][32]
    uint32_t* eCol = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(nSlots) * 32 * 20); // [nSlots][32]
    double2* ring = reinterpret_cast<double2*>(mine + static_cast<size_t>(nSlots) * 32 * 24);   // [D][32]
    // entries occupy slots [0, maxPaths), the DFS stack slots [maxPaths, nSlots)
    const int stackBase = p.maxPaths;

    const uint32_t warpGlobal = blockIdx.x * warpsPerCta + warp;
    const uint32_t warpStride = gridDim.x * warpsPerCta;
    const int S = p.segBits;
    const int segLen = 1 << S;
    const int upperLocalBits = p.nLocal - S;
    const uint32_t localSegMask = (upperLocalBits >= 32) ? 0xffffffffu : ((1u << upperLocalBits) - 1u);

    for (uint32_t tile = warpGlobal; tile < p.nTiles; tile += warpStride) {
        // =================== phase A: upper walk, one segment per lane ===========================
        const uint32_t seg = tile * 32u + lane;
        int cnt = 0;
        if (seg < p.nSeg && p.root != FDD_TERMINAL) {
            const uint32_t rowSeg = (p.rank << upperLocalBits) | seg; // segment index incl. global bits
            int sp = 0;
            int code = p.root;
            double2 w = p.rootW;
            // identity-compressed levels copy the row bit into the column: start from the row's own
            // segment and overwrite the bit of every level that is actually visited
            uint32_t col = rowSeg;
            for (;;) {
                while (code >= 0) {
                    const UpperNode& nd = upper[code];
                    const int sh = nd.level - S;
                    const int rb = static_cast<int>((rowSeg >> sh) & 1u);
                    const int2 ch = *reinterpret_cast<const int2*>(&nd.child[2 * rb]);
                    const double2* nw = reinterpret_cast<const double2*>(nd.w) + 2 * rb;
                    if (ch.x != FDD_TERMINAL) {
                        if (ch.y != FDD_TERMINAL) {
                            const int slot = (stackBase + sp) * 32 + lane;
                            eW[slot] = cmul(w, nw[1]);
                            eCode[slot] = ch.y;
                            eCol[slot] = col | (1u << sh);
                            ++sp;
                        }
                        w = cmul(w, nw[0]);
                        col &= ~(1u << sh);
                        code = ch.x;
                    } else if (ch.y != FDD_TERMINAL) {
                        w = cmul(w, nw[1]);
                        col |= (1u << sh);
                        code = ch.y;
                    } else {
                        code = FDD_TERMINAL; // dead end: this row has no entry below this node
                    }
                }
                if (code <= -2) {
                    const int slot = cnt * 32 + lane;
                    eW[slot] = w;
                    eCode[slot] = -2 - code; // decodeSub
                    eCol[slot] = col;
                    ++cnt;
                }
                if (sp == 0) break;
                --sp;
                const int slot = (stackBase + sp) * 32 + lane;
                w = eW[slot];
                code = eCode[slot];
                col = eCol[slot];
            }
        }
        __syncwarp();

        // =================== phase B: sweep the segments, lane = row ================================
        const int nSegTile = static_cast<int>(min(32u, p.nSeg - tile * 32u));
        double2* zTile = p.z + ((static_cast<uint64_t>(tile) * 32u) << S);

        if (VARIANT == 0) {
            for (int j = 0; j < nSegTile; ++j) {
                const int cj = __shfl_sync(0xffffffffu, cnt, j);
                double2 acc = make_double2(0.0, 0.0);
                for (int i = 0; i < cj; ++i) {
                    const int slot = i * 32 + j;
                    const double2 w = eW[slot];
                    const int sub = eCode[slot];
                    const uint32_t col = eCol[slot];
                    const double2* src = (p.worldBits == 0) ? p.y : p.peerY[col >> upperLocalBits];
                    double2 own = make_double2(0.0, 0.0);
                    if (lane < segLen) own = ld_stream(src + ((static_cast<uint64_t>(col & localSegMask)) << S) + lane);
                    const int flags = subFlags[sub];
                    if (flags & SUB_IDENTITY) {
                        cmac(acc, w, own);
                    } else if (flags & SUB_DIAGONAL) {
                        cmac(acc, cmul(w, subW[sub * p.kMax * 32 + lane]), own);
                    } else {
                        const int kk = subK[sub];
                        for (int k = 0; k < kk; ++k) {
                            const int at = (sub * p.kMax + k) * 32 + lane;
                            const double2 yv = shfl2(own, subCol[at]);
                            cmac(acc, cmul(w, subW[at]), yv);
                        }
                    }
                }
                if (lane < segLen) st_stream(zTile + (static_cast<uint64_t>(j) << S) + lane, acc);
            }
        } else {
            // two cursors over the tile's (segment, entry) pairs: `pj,pi` issues copies, `j,i` consumes
            int pj = 0, pi = 0;
            int pcnt = __shfl_sync(0xffffffffu, cnt, 0);
            int issued = 0;
            auto issueNext = [&]() {
                while (pj < nSegTile && pi >= pcnt) { // skip exhausted / empty segments
                    ++pj;
                    pi = 0;
                    pcnt = __shfl_sync(0xffffffffu, cnt, pj & 31);
                }
                if (pj < nSegTile) {
                    const int slot = pi * 32 + pj;
                    const uint32_t col = eCol[slot];
                    const double2* src = (p.worldBits == 0) ? p.y : p.peerY[col >> upperLocalBits];
                    if (lane < segLen) {
                        cp_async16(ring + (issued % D) * 32 + lane, src + ((static_cast<uint64_t>(col & localSegMask)) << S) + lane);
                    }
                    ++pi;
                    ++issued;
                }
                cp_async_commit(); // one group per call, possibly empty, keeps the group arithmetic uniform
            };
            for (int q = 0; q < D; ++q) issueNext();
            int consumed = 0;
            for (int j = 0; j < nSegTile; ++j) {
                const int cj = __shfl_sync(0xffffffffu, cnt, j);
                double2 acc = make_double2(0.0, 0.0);
                for (int i = 0; i < cj; ++i) {
                    // groups complete in order: leaving D-1 pending means the oldest (ours) has landed
                    switch (D) {
                        case 1: cp_async_wait<0>(); break;
                        case 2: cp_async_wait<1>(); break;
                        case 4: cp_async_wait<3>(); break;
                        case 8: cp_async_wait<7>(); break;
                        default: cp_async_wait<15>(); break;
                    }
                    __syncwarp();
                    const double2* slotData = ring + (consumed % D) * 32;
                    const int slot = i * 32 + j;
                    const double2 w = eW[slot];
                    const int sub = eCode[slot];
                    const int flags = subFlags[sub];
                    if (flags & SUB_IDENTITY) {
                        cmac(acc, w, slotData[lane]);
                    } else if (flags & SUB_DIAGONAL) {
                        cmac(acc, cmul(w, subW[sub * p.kMax * 32 + lane]), slotData[lane]);
                    } else {
                        const int kk = subK[sub];
                        for (int k = 0; k < kk; ++k) {
                            const int at = (sub * p.kMax + k) * 32 + lane;
                            cmac(acc, cmul(w, subW[at]), slotData[subCol[at]]);
                        }
                    }
                    ++consumed;
                    __syncwarp(); // every lane is done with the slot before it is refilled
                    issueNext();
                }
                if (lane < segLen) st_stream(zTile + (static_cast<uint64_t>(j) << S) + lane, acc);
            }
            cp_async_wait<0>();
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// DMAVM chunk kernel: any gate, however dense
// ------------------------------------------------------------------------------------------------
// Same walk as dmavm_walk_kernel<0>, but the per-segment entry list is bounded (CAP entries): the
// depth-first walk of every lane is suspended when its list is full, the partial lists are
// applied and accumulated into a per-warp tile of partial sums in shared memory, and the walk
// resumes where it stopped.  Used when the path count of a gate (for example a fused block that is
// dense on nine upper qubits: 512 sources per segment) does not fit the other kernels.
constexpr int kChunkCap = 8;
__host__ __device__ inline size_t chunkWarpSmem(int stackCap) {
    return static_cast<size_t>(kChunkCap + stackCap) * 32 * 24 + 32 * 32 * 16;
}

__global__ void __launch_bounds__(128) dmavm_chunk_kernel(const WalkParams p) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warpsPerCta = blockDim.x >> 5;
    unsigned char* mine = smemRaw + static_cast<size_t>(warp) * chunkWarpSmem(p.stackCap);
    const int nSlots = kChunkCap + p.stackCap;
    double2* eW = reinterpret_cast<double2*>(mine);
    int32_t* eCode = reinterpret_cast<int32_t*>(mine + static_cast<size_t>(nSlots) * 32 * 16);
    uint32_t* eCol = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(nSlots) * 32 * 20);
    double2* accTile = reinterpret_cast<double2*>(mine + static_cast<size_t>(nSlots) * 32 * 24); // [32 segments][32 lanes]
    const int stackBase = kChunkCap;
    const UpperNode* upper = p.upper; // tables stay in global memory (L1/L2 cached): this path is not tuned

    const uint32_t warpGlobal = blockIdx.x * warpsPerCta + warp;
    const uint32_t warpStride = gridDim.x * warpsPerCta;
    const int S = p.segBits;
    const int segLen = 1 << S;
    const int upperLocalBits = p.nLocal - S;
    const uint32_t localSegMask = (upperLocalBits >= 32) ? 0xffffffffu : ((1u << upperLocalBits) - 1u);

    for (uint32_t tile = warpGlobal; tile < p.nTiles; tile += warpStride) {
        const uint32_t seg = tile * 32u + lane;
        const int nSegTile = static_cast<int>(min(32u, p.nSeg - tile * 32u));
        for (int j = 0; j < 32; ++j) accTile[j * 32 + lane] = make_double2(0.0, 0.0);
        // resumable depth-first walk state of this lane
        const uint32_t rowSeg = (p.rank << upperLocalBits) | seg;
        bool done = !(seg < p.nSeg && p.root != FDD_TERMINAL);
        bool pendingPop = false;
        int sp = 0;
        int code = p.root;
        double2 w = p.rootW;
        uint32_t col = rowSeg;
        for (;;) {
            int cnt = 0;
            while (!done && cnt < kChunkCap) {
                if (pendingPop) {
                    if (sp == 0) {
                        done = true;
                        break;
                    }
                    --sp;
                    const int slot = (stackBase + sp) * 32 + lane;
                    w = eW[slot];
                    code = eCode[slot];
                    col = eCol[slot];
                    pendingPop = false;
                }
                while (code >= 0) {
                    const UpperNode& nd = upper[code];
                    const int sh = nd.level - S;
                    const int rb = static_cast<int>((rowSeg >> sh) & 1u);
                    const int2 ch = *reinterpret_cast<const int2*>(&nd.child[2 * rb]);
                    const double2* nw = reinterpret_cast<const double2*>(nd.w) + 2 * rb;
                    if (ch.x != FDD_TERMINAL) {
                        if (ch.y != FDD_TERMINAL) {
                            const int slot = (stackBase + sp) * 32 + lane;
                            eW[slot] = cmul(w, nw[1]);
                            eCode[slot] = ch.y;
                            eCol[slot] = col | (1u << sh);
                            ++sp;
                        }
                        w = cmul(w, nw[0]);
                        col &= ~(1u << sh);
                        code = ch.x;
                    } else if (ch.y != FDD_TERMINAL) {
                        w = cmul(w, nw[1]);
                        col |= (1u << sh);
                        code = ch.y;
                    } else {
                        code = FDD_TERMINAL;
                    }
                }
                if (code <= -2) {
                    const int slot = cnt * 32 + lane;
                    eW[slot] = w;
                    eCode[slot] = -2 - code;
                    eCol[slot] = col;
                    ++cnt;
                }
                pendingPop = true;
            }
            __syncwarp();
            // apply the partial lists
            for (int j = 0; j < nSegTile; ++j) {
                const int cj = __shfl_sync(0xffffffffu, cnt, j);
                if (cj == 0) continue;
                double2 acc = accTile[j * 32 + lane];
                for (int i = 0; i < cj; ++i) {
                    const int slot = i * 32 + j;
                    const double2 we = eW[slot];
                    const int sub = eCode[slot];
                    const uint32_t c = eCol[slot];
                    const double2* src = (p.worldBits == 0) ? p.y : p.peerY[c >> upperLocalBits];
                    double2 own = make_double2(0.0, 0.0);
                    if (lane < segLen) own = src[((static_cast<uint64_t>(c & localSegMask)) << S) + lane];
                    const int kk = p.subK[sub];
                    for (int k = 0; k < kk; ++k) {
                        const int at = (sub * p.kMax + k) * 32 + lane;
                        cmac(acc, cmul(we, p.subW[at]), shfl2(own, p.subCol[at]));
                    }
                }
                accTile[j * 32 + lane] = acc;
            }
            __syncwarp();
            if (__all_sync(0xffffffffu, done)) break;
        }
        for (int j = 0; j < nSegTile; ++j) {
            if (lane < segLen) st_stream(p.z + (((static_cast<uint64_t>(tile) * 32u) + j) << S) + lane, accTile[j * 32 + lane]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// DMAVM tile kernel (the fast path)
// ------------------------------------------------------------------------------------------------
// Precondition (checked on the host, gate_compile.cpp "tile bits"): every upper level with an
// off-diagonal successor is one of the TB segment-index bits in `tileMask`.  Then the T = 2^TB
// segments that differ only in those bits form a SUB-TILE that is closed under the gate: every
// source segment of an output segment lies in the same sub-tile.  Every amplitude is therefore
// read from HBM once and written once — 32 B per amplitude, the algorithmic minimum — whatever
// the number of non-zeros per row.
//
// A warp owns WARP TILES of 32 segments = Q sub-tiles (the sub-tile bits plus `fillMask` bits,
// the lowest free index bits, so consecutive sub-tiles are neighbours in memory):
//   phase A  lane = segment of the warp tile: depth-first walk of the upper levels, leaving per
//            segment a list of (weight, source slot inside its sub-tile[, sub table]) in shared
//            memory, padded with zero weights to `maxPaths` entries so phase B is branch-free;
//   phase B  lane = amplitude: the sub-tiles stream through a per-warp ring of R stage slots
//            filled by cp.async (16-byte vectors, one coalesced 512-byte request per segment).
//            The ring is indexed by a running sub-tile counter, so copies for the next warp
//            tile are already in flight while phase A of that tile runs.  The T output
//            segments of a sub-tile are accumulated together (T independent FMA chains).
// MODE 0: the low S levels are untouched (identity sub table)      z_j = sum_i w_ji y_slot(j,i)
// MODE 1: one sub table L for the whole gate, gather first          z_j = sum_i w_ji (L y_slot(j,i))
// MODE 2: one sub table L for the whole gate, shuffle last          z_j = L (sum_i w_ji y_slot(j,i))
//         (the operator factorises as U (x) L; L lives in registers, KT entries per row)
// MODE 3: sub table depends on the path, gather first               z_j = sum_i w_ji (L_sub(j,i) y_slot(j,i))
// MODE 4: like MODE 3 but the entry lists are kept per sub table (at most 8 tables, 15 slots), so each
//         table is applied once per output segment, shuffle last    z_j = sum_s L_s (sum_{i in s} w_ji y_slot(j,i))
// MODE 5: MODE 0 with a complete block of 8 / 16 sources per segment on the FP64 tensor cores (DMMA.8x8x4)
// MODE 6: MODE 3 for uniform gates: the products w_ji L_sub(j,i) are formed once per CTA into a flat table of
//         KT = paths x ELL width entries per row (lane-private weight, source = slot * 32 + column), so
//         phase B is one multiply-add per entry: no per-path sub-table lookups, no weight products
// KT = ELL width of the sub tables rounded up to 2, 4 or 8 (static unrolling); KT = 0 (MODE 3 only)
// reads the width at run time.  The host picks MODE 1 or 2 by instruction count.
template <int TB, int MODE = 0> struct TileShape {
    static constexpr int T = 1 << TB;
    static constexpr int R = MODE == 5 ? FDD_M5_STAGES : ((16 / T) < 2 ? 2 : (16 / T));
    static constexpr int JC = T < 8 ? T : 8;
};
// Shared memory of the tile kernel: the entry area (lists + walk stack) and the stage ring.  A uniform
// gate gives every tile the same lists, so one entry area serves the whole CTA and only the ring is
// per warp; otherwise every warp has its own entry area.
__host__ __device__ inline size_t tileEntryBytes(int maxPaths, int stackCap) { return static_cast<size_t>(maxPaths + stackCap) * 32 * 20; }
// tensor-core path (MODE 5): kM5Stages stage slots of one sub-tile each
constexpr int kM5Stages = FDD_M5_STAGES;
constexpr int kM5MaxWarps = FDD_M5_MAXWARPS;
__host__ __device__ inline size_t tileRingBytes(int tileBits, int tensorCore = 0) {
    const int T = 1 << tileBits;
    const int R = tensorCore ? kM5Stages : ((16 / T) < 2 ? 2 : (16 / T));
    return static_cast<size_t>(R) * T * 512;
}
__host__ __device__ inline size_t tileWarpSmem(int maxPaths, int stackCap, int tileBits, int uniform, int tensorCore = 0) {
    return tileRingBytes(tileBits, tensorCore) + (uniform ? 0 : tileEntryBytes(maxPaths, stackCap));
}
__host__ __device__ inline size_t tileCtaSmem(int maxPaths, int stackCap, int uniform) { return uniform ? tileEntryBytes(maxPaths, stackCap) : 0; }
// MODE 6: the flat table of a uniform gate, [T rows][E entries][32 lanes] of (weight 16 B, source 2 B)
__host__ __device__ inline size_t tileFlatBytes(int tileBits, int entries) { return (static_cast<size_t>(1) << tileBits) * entries * 32 * 18; }

__device__ __forceinline__ uint32_t depositBits(uint32_t x, uint32_t mask) { // pdep
    uint32_t out = 0;
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        const uint32_t lowest = m & (0u - m);
        if (x & 1u) out |= lowest;
        x >>= 1;
    }
    return out;
}
// spread x over the zero bits of mask (mask has few bits set)
__device__ __forceinline__ uint32_t depositAround(uint32_t x, uint32_t mask) {
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        const uint32_t lowest = m & (0u - m);
        x = ((x & ~(lowest - 1u)) << 1) | (x & (lowest - 1u));
    }
    return x;
}

// (MODE 5 keeps more of the gate block in registers: its own launch bounds)
template <int TB, int MODE, int KT> __global__ void __launch_bounds__(MODE == 5 ? 32 * FDD_M5_MAXWARPS : 256, MODE == 5 ? FDD_M5_MINCTAS : 2) dmavm_tile_kernel(const WalkParams p) {
    constexpr int T = TileShape<TB, MODE>::T;
    constexpr int R = TileShape<TB, MODE>::R;
    // programmatic dependent launch: let the next launch of the stream start its prologue (table staging,
    // walk of a uniform gate) while this grid drains; it waits below before it touches the state
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    unsigned char* cursor = smemRaw;
    const UpperNode* upper = p.upper;
    const double2* subW = p.subW;
    const uint8_t* subCol = p.subCol;
    const int32_t* subK = p.subK;
    if (p.tablesInSmem) {
        UpperNode* su = reinterpret_cast<UpperNode*>(cursor);
        cursor += static_cast<size_t>(p.nUpper) * sizeof(UpperNode);
        double2* sw = reinterpret_cast<double2*>(cursor);
        const int nEnt = p.nSub * p.kMax * 32;
        cursor += static_cast<size_t>(nEnt) * 16;
        int32_t* sk = reinterpret_cast<int32_t*>(cursor);
        cursor += static_cast<size_t>(p.nSub) * 4;
        uint8_t* sc = cursor;
        {
            const int4* src = reinterpret_cast<const int4*>(p.upper);
            int4* dst = reinterpret_cast<int4*>(su);
            for (int i = threadIdx.x; i < p.nUpper * 6; i += blockDim.x) dst[i] = src[i];
        }
        if (MODE != 0) {
            for (int i = threadIdx.x; i < nEnt; i += blockDim.x) {
                sw[i] = p.subW[i];
                sc[i] = p.subCol[i];
            }
            for (int i = threadIdx.x; i < p.nSub; i += blockDim.x) sk[i] = p.subK[i];
        }
        __syncthreads();
        upper = su;
        subW = sw;
        subCol = sc;
        subK = sk;
        cursor = smemRaw + walkTableSmem(p.nUpper, p.nSub, p.kMax);
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warpsPerCta = blockDim.x >> 5;
    unsigned char* ctaEntries = cursor;
    const bool lookup = MODE == 5 && p.ctxPhase == 2; // blocks come from the context table: no entry area at all
    if (!lookup) cursor += tileCtaSmem(p.maxPaths, p.stackCap, p.uniform);
    unsigned char* mine = cursor + static_cast<size_t>(warp) * (lookup ? tileRingBytes(TB, 1) : tileWarpSmem(p.maxPaths, p.stackCap, TB, p.uniform, MODE == 5));
    unsigned char* entries = p.uniform ? ctaEntries : mine;
    const int nSlotsE = p.maxPaths + p.stackCap;
    double2* eW = reinterpret_cast<double2*>(entries);                                               // [nSlotsE][32]
    uint32_t* ePack = reinterpret_cast<uint32_t*>(entries + static_cast<size_t>(nSlotsE) * 32 * 16); // [nSlotsE][32]
    double2* ring = reinterpret_cast<double2*>((p.uniform || lookup) ? mine : mine + tileEntryBytes(p.maxPaths, p.stackCap)); // [R][T][32]
    const int stackBase = p.maxPaths;
    const int P = p.maxPaths;
    // Dense register path (MODE 0 / 2: weights do not depend on the lane; 4..16 segments per sub-tile):
    // phase A stores the entries of a row by source slot (zero weight where the matrix has none), so the
    // T inputs of a sub-tile are loaded into registers once and every output is T FMAs against them,
    // without per-entry slot lookups.  The host enables it when T <= 2 * maxPaths.
    constexpr bool DENSE_OK = (MODE == 0 || MODE == 2 || MODE == 5) && T >= 4 && T <= 16;
    const bool denseTile = DENSE_OK && p.denseSlots != 0;

    const uint32_t warpGlobal = blockIdx.x * warpsPerCta + warp;
    const uint32_t warpStride = gridDim.x * warpsPerCta;
    const int S = p.segBits;
    const int segLen = 1 << S;
    const int upperLocalBits = p.nLocal - S;
    const int wtBits = p.tileBits;           // segments per warp tile = 2^wtBits (<= 32)
    const int qBits = wtBits - TB;
    const int Q = 1 << qBits;                // sub-tiles per warp tile
    const int wtSegs = 1 << wtBits;
    const uint32_t wtMask = p.tileMask | p.fillMask;
    // lane l <-> segment (sub-tile q = l >> TB, slot t = l & (T-1)) of the warp tile
    const uint32_t myDep = depositBits(static_cast<uint32_t>(lane) & (T - 1), p.tileMask) |
                           depositBits(static_cast<uint32_t>(lane) >> TB, p.fillMask);
    const uint32_t rankBits = p.rank << upperLocalBits;
    // MODE 1/2: the single sub table lives in registers
    constexpr int KR = (MODE == 1 || MODE == 2) ? (KT > 0 ? KT : 1) : 1;
    double2 Lw[KR];
    int Lc[KR];
    if (MODE == 1 || MODE == 2) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            Lw[k] = k < p.kMax ? subW[k * 32 + lane] : make_double2(0.0, 0.0);
            Lc[k] = k < p.kMax ? subCol[k * 32 + lane] : lane;
        }
    }

    // =================== phase A: upper walk, lane = segment of the warp tile =====================
    auto walkTile = [&](uint32_t base) {
        if (lane < wtSegs) {
            int cnt = 0;
            if (DENSE_OK && p.denseSlots) {
                for (int i = 0; i < T; ++i) eW[i * 32 + lane] = make_double2(0.0, 0.0);
            }
            if (p.root != FDD_TERMINAL) {
                const uint32_t rowSeg = rankBits | base | myDep;
                const int tl = lane & (T - 1);
                int sp = 0;
                int code = p.root;
                double2 w = p.rootW;
                uint32_t slot = static_cast<uint32_t>(tl);
                for (;;) {
                    while (code >= 0) {
                        const UpperNode& nd = upper[code];
                        const int sb = nd.slotBit;
                        if (sb < 0) {
                            // a level outside the tile bits is diagonal by construction: one successor, no column change
                            const int d = 3 * static_cast<int>((rowSeg >> (nd.level - S)) & 1u);
                            const int next = nd.child[d];
                            if (next != FDD_TERMINAL) w = cmul(w, reinterpret_cast<const double2*>(nd.w)[d]);
                            code = next;
                            continue;
                        }
                        const int rb = (tl >> sb) & 1;
                        const uint32_t bit = 1u << sb;
                        const int2 ch = *reinterpret_cast<const int2*>(&nd.child[2 * rb]);
                        const double2* nw = reinterpret_cast<const double2*>(nd.w) + 2 * rb;
                        if (ch.x != FDD_TERMINAL) {
                            if (ch.y != FDD_TERMINAL) {
                                const int at = (stackBase + sp) * 32 + lane;
                                eW[at] = cmul(w, nw[1]);
                                ePack[at] = (static_cast<uint32_t>(ch.y) & 0xffffffu) | ((slot | bit) << 24);
                                ++sp;
                            }
                            w = cmul(w, nw[0]);
                            slot &= ~bit;
                            code = ch.x;
                        } else if (ch.y != FDD_TERMINAL) {
                            w = cmul(w, nw[1]);
                            slot |= bit;
                            code = ch.y;
                        } else {
                            code = FDD_TERMINAL; // dead end: this row has no entry below this node
                        }
                    }
                    if (code <= -2) {
                        const int sub = -2 - code;
                        int at;
                        if (MODE == 4) { // per-sub list: 4-bit fill counters packed into cnt
                            at = (p.subBase[sub] + ((cnt >> (4 * sub)) & 15)) * 32 + lane;
                            cnt += 1 << (4 * sub);
                        } else if (DENSE_OK && p.denseSlots) {
                            at = static_cast<int>(slot) * 32 + lane; // one entry per (row, source slot)
                        } else {
                            at = cnt * 32 + lane;
                            ++cnt;
                        }
                        eW[at] = w;
                        ePack[at] = slot | (static_cast<uint32_t>(sub) << 8);
                    }
                    if (sp == 0) break;
                    --sp;
                    const int at = (stackBase + sp) * 32 + lane;
                    w = eW[at];
                    const uint32_t pk = ePack[at];
                    slot = pk >> 24;
                    code = static_cast<int>(pk << 8) >> 8; // sign-extend the 24-bit successor code
                }
            }
            // zero-weight padding
            if (MODE == 4) {
                for (int sIdx = 0; sIdx < p.nSub; ++sIdx) {
                    for (int i = p.subBase[sIdx] + ((cnt >> (4 * sIdx)) & 15); i < p.subBase[sIdx + 1]; ++i) {
                        eW[i * 32 + lane] = make_double2(0.0, 0.0);
                        ePack[i * 32 + lane] = static_cast<uint32_t>(lane & (T - 1));
                    }
                }
            } else if (!(DENSE_OK && p.denseSlots)) {
                for (int i = cnt; i < P; ++i) {
                    eW[i * 32 + lane] = make_double2(0.0, 0.0);
                    ePack[i * 32 + lane] = static_cast<uint32_t>(lane & (T - 1));
                }
            }
        }
    };
    if (p.uniform && !lookup) {
        // every tile sees the same lists: warp 0 walks once for the whole CTA
        // (with a context table there is no entry area to walk into: the blocks come from lookupA)
        if (warp == 0) walkTile(0u);
        __syncthreads();
    }
    if constexpr (MODE == 5) {
        if (p.ctxPhase == 1) {
            // context pre-pass: one walk per value of the context bits, rows = lanes 0..T-1 (sub-tile 0 of the warp tile)
            for (uint32_t ctx = blockIdx.x * warpsPerCta + warp; ctx < static_cast<uint32_t>(p.nCtx); ctx += gridDim.x * warpsPerCta) {
                walkTile(depositBits(ctx, p.ctxMask));
                __syncwarp();
                if (lane < T) {
                    for (int i = 0; i < T; ++i) p.ctxTable[(static_cast<size_t>(ctx) * T + i) * T + lane] = eW[i * 32 + lane];
                }
                __syncwarp();
            }
            return;
        }
    }
    // MODE 6: flat table behind the per-warp areas, built by the whole CTA from the lists of sub-tile 0
    double2* flatW = nullptr;
    uint16_t* flatC = nullptr;
    if constexpr (MODE == 6) {
        constexpr int E = KT;
        unsigned char* flat = cursor + static_cast<size_t>(warpsPerCta) * tileWarpSmem(p.maxPaths, p.stackCap, TB, p.uniform);
        flatW = reinterpret_cast<double2*>(flat);
        flatC = reinterpret_cast<uint16_t*>(flat + static_cast<size_t>(T) * E * 32 * 16);
        const int kw = p.kMax; // ELL width of the tables as uploaded
        for (int idx = warp; idx < T * E; idx += warpsPerCta) {
            const int t = idx / E;
            const int e = idx - t * E;
            const int i = e / kw;
            const int k = e - i * kw;
            double2 w = make_double2(0.0, 0.0);
            uint32_t src = static_cast<uint32_t>(t) * 32u + static_cast<uint32_t>(lane);
            if (i < P) {
                const uint32_t pk = ePack[i * 32 + t];
                const int at = (static_cast<int>(pk >> 8) * kw + k) * 32 + lane;
                w = cmul(eW[i * 32 + t], subW[at]);
                src = (pk & 31u) * 32u + subCol[at];
            }
            flatW[idx * 32 + lane] = w;
            flatC[idx * 32 + lane] = static_cast<uint16_t>(src);
        }
        __syncthreads();
    }
    if (warpGlobal >= p.nTiles) return;
    // MODE 6 with a small table (rows x entries <= 16): the lane-private weights live in registers, which halves the
    // shared-memory traffic of phase B (weights and gathered sources both cross the 128 B/clk crossbar otherwise)
    constexpr bool FLAT_IN_REGS = MODE == 6 && FDD_M6_REGS && T * (KT > 0 ? KT : 1) <= 16;
    double2 flatReg[FLAT_IN_REGS ? T * KT : 1];
    if constexpr (FLAT_IN_REGS) {
#pragma unroll
        for (int i = 0; i < T * KT; ++i) flatReg[i] = flatW[i * 32 + lane];
    }

    // everything above read only the gate tables; the state buffers belong to the previous launch until here
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ---- copy pipeline state: running sub-tile counter over all warp tiles of this warp ---------
    uint32_t issueTile = warpGlobal;
    uint32_t issueBase = depositAround(issueTile, wtMask);
    int issueQ = 0;
    int issueSlot = 0;
    auto issueNext = [&]() {
        if (issueTile < p.nTiles) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const uint32_t dep = __shfl_sync(0xffffffffu, myDep, (issueQ << TB) | t);
                if (lane < segLen) {
                    // MODE 5 reads the stage as tensor-core fragments: amplitude a of segment t sits at a ^ 2(t & 3)
                    const int pos = (MODE == 5) ? (lane ^ ((t & 3) << 1)) : lane;
                    cp_async16(ring + (issueSlot * T + t) * 32 + pos, p.y + ((static_cast<uint64_t>(issueBase | dep)) << S) + lane);
                }
            }
            issueSlot = (issueSlot + 1 == R) ? 0 : issueSlot + 1;
            if (++issueQ == Q) {
                issueQ = 0;
                issueTile += warpStride;
                issueBase = depositAround(issueTile, wtMask);
            }
        }
        cp_async_commit(); // one group per call (possibly empty) keeps the wait arithmetic uniform
    };
#pragma unroll
    for (int r = 0; r < R; ++r) issueNext();
    int useSlot = 0;

    if constexpr (MODE == 5) {
        // =================== phase B on the FP64 tensor cores (dense upper block, low levels untouched) =========
        // A sub-tile is a T x 32 complex matrix Y (row = segment slot, column = amplitude of the segment) and
        // the gate acts as Z = M Y with the T x T block M of phase A.  In real arithmetic
        //     Zr = Mr Yr + Mi (-Yi),   Zi = Mi Yr + Mr Yi,
        // i.e. 4 (T/8)(T/4) DMMA.8x8x4 per group of eight amplitude columns instead of 4 T^2 / 32 DFMAs per
        // column: 8x fewer issue slots for the same flops, and no per-entry shared-memory weight loads
        // (the block lives in A fragments: 2 (T/8)(T/4) doubles per lane).
        // The stage holds whole 512-byte segments (the only request shape that keeps HBM at full rate: 128-byte
        // column groups per request measured 3.9 TB/s) with an XOR swizzle that makes the fragment reads
        // conflict-free; the slot is refilled as soon as its last fragments are in registers.
        constexpr int MT = T / 8;             // 8-row blocks of M
        constexpr int KTL = T / 4;            // 4-column blocks of M
        constexpr int NT = 2;                 // 8-amplitude column groups computed together (independent DMMA chains)
        const int fr = lane >> 2;             // fragment row (A, D) / column (B)
        const int fc = lane & 3;              // fragment column (A) / row (B)
        double aR[MT][KTL], aI[MT][KTL];
#if FDD_M5_3M
        double aS[MT][KTL];
#endif
        auto loadA = [&](int rowBase) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kt = 0; kt < KTL; ++kt) {
                    const double2 w = eW[(4 * kt + fc) * 32 + rowBase + 8 * mt + fr];
                    aR[mt][kt] = w.x;
                    aI[mt][kt] = w.y;
#if FDD_M5_3M
                    aS[mt][kt] = w.x + w.y;
#endif
                }
            }
        };
        auto lookupA = [&](uint32_t segBits) {
            uint32_t ctx = 0;
            int at = 0;
            for (uint32_t m = p.ctxMask; m != 0; m &= m - 1, ++at) {
                if (segBits & (m & (0u - m))) ctx |= 1u << at;
            }
            const double2* blk = p.ctxTable + static_cast<size_t>(ctx) * T * T;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kt = 0; kt < KTL; ++kt) {
                    const double2 w = __ldg(blk + (4 * kt + fc) * T + 8 * mt + fr);
                    aR[mt][kt] = w.x;
                    aI[mt][kt] = w.y;
#if FDD_M5_3M
                    aS[mt][kt] = w.x + w.y;
#endif
                }
            }
        };
        if (p.uniform && !lookup) loadA(0); // (with a context table there is no entry area: the block comes from lookupA)
        for (uint32_t tile = warpGlobal; tile < p.nTiles; tile += warpStride) {
            const uint32_t base = depositAround(tile, wtMask);
            if (!p.uniform && !lookup) {
                walkTile(base);
                __syncwarp();
            }
            for (int q = 0; q < Q; ++q) {
                if (lookup) lookupA(base | depositBits(static_cast<uint32_t>(q), p.fillMask));
                else if (!p.uniform) loadA(q << TB);
                uint32_t dep[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dep[mt] = __shfl_sync(0xffffffffu, myDep, (q << TB) + 8 * mt + fr);
                cp_async_wait<R - 1>(); // groups complete in order: the oldest one is this sub-tile
                __syncwarp();
                const double2* st = ring + useSlot * T * 32;
#pragma unroll
                for (int nt0 = 0; nt0 < 4; nt0 += NT) {
                    // B fragments of this pass: segment 4 kt + fc, amplitude 8 (nt0 + u) + fr (swizzled like the copy)
                    double2 y[KTL][NT];
#pragma unroll
                    for (int kt = 0; kt < KTL; ++kt) {
#pragma unroll
                        for (int u = 0; u < NT; ++u) y[kt][u] = st[(4 * kt + fc) * 32 + ((8 * (nt0 + u) + fr) ^ (fc << 1))];
                    }
                    if (nt0 + NT == 4) {
                        // the stage slot is in registers: refill it while the tensor cores work
                        __syncwarp();
                        useSlot = (useSlot + 1 == R) ? 0 : useSlot + 1;
                        issueNext();
                    }
                    double2 z0[NT][MT], z1[NT][MT]; // amplitudes 8 nt + 2 fc and + 1 of row 8 mt + fr
#if FDD_M5_3M
                    // three real products per complex one: P1 = Mr Yr, P2 = Mi Yi, P3 = (Mr + Mi)(Yr + Yi);
                    // Zr = P1 - P2, Zi = P3 - P1 - P2
                    double p1[NT][MT][2], p2[NT][MT][2], p3[NT][MT][2];
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) p1[u][mt][0] = p1[u][mt][1] = p2[u][mt][0] = p2[u][mt][1] = p3[u][mt][0] = p3[u][mt][1] = 0.0;
                    }
#pragma unroll
                    for (int kt = 0; kt < KTL; ++kt) {
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            const double ys = y[kt][u].x + y[kt][u].y;
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                dmma884(p1[u][mt], aR[mt][kt], y[kt][u].x);
                                dmma884(p2[u][mt], aI[mt][kt], y[kt][u].y);
                                dmma884(p3[u][mt], aS[mt][kt], ys);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            z0[u][mt] = make_double2(p1[u][mt][0] - p2[u][mt][0], (p3[u][mt][0] - p1[u][mt][0]) - p2[u][mt][0]);
                            z1[u][mt] = make_double2(p1[u][mt][1] - p2[u][mt][1], (p3[u][mt][1] - p1[u][mt][1]) - p2[u][mt][1]);
                        }
                    }
#else
                    double zr[NT][MT][2], zi[NT][MT][2];
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) zr[u][mt][0] = zr[u][mt][1] = zi[u][mt][0] = zi[u][mt][1] = 0.0;
                    }
                    // 2 NT MT independent accumulators are touched between two DMMAs on the same one
#pragma unroll
                    for (int kt = 0; kt < KTL; ++kt) {
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                dmma884(zr[u][mt], aR[mt][kt], y[kt][u].x);
                                dmma884(zi[u][mt], aI[mt][kt], y[kt][u].x);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            const double nyi = -y[kt][u].y;
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                dmma884(zr[u][mt], aI[mt][kt], nyi);
                                dmma884(zi[u][mt], aR[mt][kt], y[kt][u].y);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            z0[u][mt] = make_double2(zr[u][mt][0], zi[u][mt][0]);
                            z1[u][mt] = make_double2(zr[u][mt][1], zi[u][mt][1]);
                        }
                    }
#endif
                    // D fragment: 32 contiguous bytes per lane, 128 per row and store instruction
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            st_stream2(p.z + ((static_cast<uint64_t>(base | dep[mt])) << S) + 8 * (nt0 + u) + 2 * fc, z0[u][mt], z1[u][mt]);
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
    } else {
    for (uint32_t tile = warpGlobal; tile < p.nTiles; tile += warpStride) {
        const uint32_t base = depositAround(tile, wtMask);
        if (!p.uniform) walkTile(base);
        __syncwarp();
        // =================== phase B: stream the sub-tiles, lane = amplitude ========================
        // G consecutive sub-tiles are computed together so that at least four output segments
        // (independent FMA chains) are in flight even when a sub-tile is one or two segments
        auto computeGroup = [&](auto groupTag, int q) {
            constexpr int G = decltype(groupTag)::value;
            constexpr int NACC = (G * T < 8) ? G * T : 8; // accumulators per pass
            cp_async_wait<R - G>(); // groups complete in order: the G oldest ones are these sub-tiles
            __syncwarp();
#pragma unroll
            for (int j0 = 0; j0 < G * T; j0 += NACC) {
                double2 acc[NACC];
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[a] = make_double2(0.0, 0.0);
                // accumulator a <-> output segment (q + (j0+a)/T, slot (j0+a)%T): lane rowBase+a of phase A
                const int rowBase = (q << TB) + j0;
                auto stageOf = [&](int a) -> const double2* {
                    const int g = (j0 + a) / T; // compile-time after unrolling
                    int slot = useSlot + g;
                    if (slot >= R) slot -= R;
                    return ring + slot * T * 32;
                };
                if (MODE == 0 || MODE == 2) {
                    if (DENSE_OK && denseTile) {
                        // inputs of the sub-tile in registers (G == 1 here: one ring slot)
                        constexpr int TD = DENSE_OK ? T : 1;
                        double2 yv[TD];
                        const double2* st = ring + useSlot * T * 32;
#pragma unroll
                        for (int i = 0; i < TD; ++i) yv[i] = st[i * 32 + lane];
#pragma unroll
                        for (int i = 0; i < TD; ++i) {
#pragma unroll
                            for (int a = 0; a < NACC; ++a) cmac(acc[a], eW[i * 32 + rowBase + a], yv[i]);
                        }
                    } else {
                        for (int i = 0; i < P; ++i) {
#pragma unroll
                            for (int a = 0; a < NACC; ++a) {
                                const double2 w = eW[i * 32 + rowBase + a];
                                const uint32_t pk = ePack[i * 32 + rowBase + a];
                                cmac(acc[a], w, stageOf(a)[(pk & 31u) * 32 + lane]);
                            }
                        }
                    }
                    if (MODE == 2) {
                        double2 out[NACC];
#pragma unroll
                        for (int a = 0; a < NACC; ++a) out[a] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int k = 0; k < KR; ++k) {
#pragma unroll
                            for (int a = 0; a < NACC; ++a) cmac(out[a], Lw[k], shfl2(acc[a], Lc[k]));
                        }
#pragma unroll
                        for (int a = 0; a < NACC; ++a) acc[a] = out[a];
                    }
                } else if (MODE == 1) {
                    for (int i = 0; i < P; ++i) {
#pragma unroll
                        for (int a = 0; a < NACC; ++a) {
                            const double2 w = eW[i * 32 + rowBase + a];
                            const uint32_t pk = ePack[i * 32 + rowBase + a];
                            const double2* src = stageOf(a) + (pk & 31u) * 32;
                            double2 t = make_double2(0.0, 0.0);
#pragma unroll
                            for (int k = 0; k < KR; ++k) cmac(t, Lw[k], src[Lc[k]]);
                            cmac(acc[a], w, t);
                        }
                    }
                } else if (MODE == 6) {
                    constexpr int E = KT > 0 ? KT : 1;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
#pragma unroll
                        for (int a = 0; a < NACC; ++a) {
                            const int t = (j0 + a) & (T - 1); // row inside the sub-tile (compile-time after unrolling)
                            const int at = (t * E + e) * 32 + lane;
                            if constexpr (FLAT_IN_REGS) {
                                cmac(acc[a], flatReg[t * E + e], stageOf(a)[flatC[at]]);
                            } else {
                                cmac(acc[a], flatW[at], stageOf(a)[flatC[at]]);
                            }
                        }
                    }
                } else if (MODE == 4) {
                    for (int sIdx = 0; sIdx < p.nSub; ++sIdx) {
                        double2 u[NACC];
#pragma unroll
                        for (int a = 0; a < NACC; ++a) u[a] = make_double2(0.0, 0.0);
                        for (int i = p.subBase[sIdx]; i < p.subBase[sIdx + 1]; ++i) {
#pragma unroll
                            for (int a = 0; a < NACC; ++a) {
                                const double2 w = eW[i * 32 + rowBase + a];
                                const uint32_t pk = ePack[i * 32 + rowBase + a];
                                cmac(u[a], w, stageOf(a)[(pk & 31u) * 32 + lane]);
                            }
                        }
                        const int at0 = sIdx * (KT > 0 ? KT : 1) * 32 + lane;
#pragma unroll
                        for (int k = 0; k < (KT > 0 ? KT : 1); ++k) {
                            const double2 lw = subW[at0 + k * 32];
                            const int lc = subCol[at0 + k * 32];
#pragma unroll
                            for (int a = 0; a < NACC; ++a) cmac(acc[a], lw, shfl2(u[a], lc));
                        }
                    }
                } else {
                    for (int i = 0; i < P; ++i) {
#pragma unroll
                        for (int a = 0; a < NACC; ++a) {
                            const double2 w = eW[i * 32 + rowBase + a];
                            const uint32_t pk = ePack[i * 32 + rowBase + a];
                            const double2* src = stageOf(a) + (pk & 31u) * 32;
                            // KT > 0: the host padded every sub table to exactly KT entries per row
                            const int at0 = static_cast<int>(pk >> 8) * (KT > 0 ? KT : p.kMax) * 32 + lane;
                            double2 t = make_double2(0.0, 0.0);
                            if (KT > 0) {
                                int col[KT > 0 ? KT : 1];
#pragma unroll
                                for (int k = 0; k < (KT > 0 ? KT : 1); ++k) col[k] = subCol[at0 + k * 32];
#pragma unroll
                                for (int k = 0; k < (KT > 0 ? KT : 1); ++k) cmac(t, subW[at0 + k * 32], src[col[k]]);
                            } else {
                                for (int k = 0; k < p.kMax; ++k) cmac(t, subW[at0 + k * 32], src[subCol[at0 + k * 32]]);
                            }
                            cmac(acc[a], w, t);
                        }
                    }
                }
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint32_t dep = __shfl_sync(0xffffffffu, myDep, rowBase + a);
                    if (lane < segLen) st_stream(p.z + ((static_cast<uint64_t>(base | dep)) << S) + lane, acc[a]);
                }
            }
            __syncwarp(); // every lane is done with the stage slots before they are refilled
            useSlot += G;
            if (useSlot >= R) useSlot -= R;
#pragma unroll
            for (int g = 0; g < G; ++g) issueNext();
        };
        constexpr int GS = (T >= 4) ? 1 : 4 / T;
        if (GS > 1 && Q >= GS) {
            for (int q = 0; q < Q; q += GS) computeGroup(std::integral_constant<int, GS>{}, q);
        } else {
            for (int q = 0; q < Q; ++q) computeGroup(std::integral_constant<int, 1>{}, q);
        }
    }
    cp_async_wait<0>();
    } // MODE != 5
}

// ------------------------------------------------------------------------------------------------
// utility kernels
// ------------------------------------------------------------------------------------------------
// interleaved complex -> two planar arrays (the reference's state_real / state_imag layout)
__global__ void deinterleave_kernel(const double2* __restrict__ in, double* __restrict__ re, double* __restrict__ im, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 v = in[i];
        re[i] = v.x;
        im[i] = v.y;
    }
}
__global__ void interleave_kernel(const double* __restrict__ re, const double* __restrict__ im, double2* __restrict__ out, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = make_double2(re[i], im[i]);
    }
}
__global__ void zero_state_kernel(double2* out, uint64_t n, int setOne) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = make_double2((i == 0 && setOne) ? 1.0 : 0.0, 0.0);
    }
}
// amplitudes at arbitrary indices (sampled comparison of large states)
__global__ void gather_kernel(const double2* __restrict__ in, const uint64_t* __restrict__ idx, uint64_t count, double2* __restrict__ out) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[idx[i]];
}
// ---- measurement sampling ---------------------------------------------------------------------------
// probability mass of every block of `blockAmps` consecutive amplitudes (one CTA per block, fixed order)
__global__ void __launch_bounds__(256) block_mass_kernel(const double2* __restrict__ in, uint64_t n, uint32_t blockAmps, double* __restrict__ mass) {
    __shared__ double sh[8];
    const uint64_t first = static_cast<uint64_t>(blockIdx.x) * blockAmps;
    double s = 0.0;
    for (uint32_t i = threadIdx.x; i < blockAmps && first + i < n; i += blockDim.x) {
        const double2 v = in[first + i];
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += sh[w];
        mass[blockIdx.x] = t;
    }
}
// one warp per shot: walk the chosen block until the running mass passes the shot's residual
__global__ void __launch_bounds__(256) sample_resolve_kernel(const double2* __restrict__ in, uint64_t n, uint32_t blockAmps,
                                                             const uint32_t* __restrict__ shotBlock, const double* __restrict__ shotResidual,
                                                             uint64_t nShots, uint64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t shot = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (shot >= nShots) return;
    const uint64_t first = static_cast<uint64_t>(shotBlock[shot]) * blockAmps;
    const double target = shotResidual[shot];
    double running = 0.0;
    uint64_t lastNonZero = first;
    uint64_t found = ~uint64_t{0};
    for (uint32_t base = 0; base < blockAmps && found == ~uint64_t{0}; base += 32) {
        const uint64_t idx = first + base + lane;
        double pr = 0.0;
        if (idx < n) {
            const double2 v = in[idx];
            pr = fma(v.x, v.x, v.y * v.y);
        }
        double incl = pr; // inclusive warp scan
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, pr > 0.0 && running + incl > target);
        const unsigned nz = __ballot_sync(0xffffffffu, pr > 0.0);
        if (nz != 0) lastNonZero = first + base + (31 - __clz(nz));
        if (hit != 0) found = first + base + (__ffs(hit) - 1);
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) out[shot] = found != ~uint64_t{0} ? found : lastNonZero; // rounding at the block end: last populated state
}

// sum |amp|^2: per-block partial sums in fixed order, final pass on one block (deterministic)
__global__ void norm2_partial_kernel(const double2* __restrict__ in, uint64_t n, double* __restrict__ partial) {
    __shared__ double sh[32];
    double s = 0.0;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 v = in[i];
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
__global__ void norm2_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *out = s;
    }
}

} // namespace fddb200

// block_kernel.cuh — tile-resident dense-block DMAVM for sm_100a (see block_plan.hpp for the decomposition).
//
// One CTA owns tiles of 2^tileBits amplitudes (<= 128 KiB) that are closed under every block of the pass:
//   load    every 512-byte segment of the tile with cp.async (16 bytes per lane, one coalesced request per warp and
//           segment: the request shape that keeps HBM at full rate) into the swizzled layout swz();
//   apply   each block in place on the FP64 tensor cores: a UNIT is 2^k rows (the block's target qubits) x 8 columns
//           (three other tile bits); per unit Z(2^k x 8) = M(2^k x 2^k) Y(2^k x 8) as DMMA.8x8x4 tiles, a complex
//           product from three real ones (P1 = Mr Yr, P2 = Mi Yi, P3 = (Mr + Mi)(Yr + Yi); Zr = P1 - P2,
//           Zi = P3 - P1 - P2).  The matrix is looked up per unit from the block's context table (controls and phases
//           on qubits outside the block select the matrix) and kept in A fragments while the context does not change;
//   store   the tile back as 512-byte segments (st.global.cs).
// HBM traffic: 32 bytes per amplitude and pass, whatever the pass holds.  The reference counterpart is one call of
// DDArrMultiplyIP per fused gate (include/dd/SwitchPackage.hpp:1897-2261).
#pragma once

#include "block_plan.hpp"
#include "kernels.cuh"

namespace fddb200 {

// warps per CTA: 16 when one CTA owns the SM (four per scheduler hide the shared-memory and tensor-pipe latencies of each
// other), 8 when two CTAs share it
constexpr int kBlockWarpsMax = 16;
// shared memory: the tile buffers; per segment of the tile its offset in the state and its swizzled place in the tile
// (uint2); per block and lane the fragment address pieces (8 words); per block and unit one packed word
__host__ __device__ inline size_t blockPassSmem(int tileBits, int nBuffers, int nBlocks, int maxUnits) {
    return (static_cast<size_t>(16 * nBuffers) << tileBits) + (static_cast<size_t>(8) << (tileBits - kLaneBits)) +
           static_cast<size_t>(nBlocks) * 32 * 32 + static_cast<size_t>(nBlocks) * maxUnits * 4;
}

// Per block and lane, computed once per CTA (everything here only depends on the plan):
//   w[0] = (pB ^ pKt0) | (pB ^ pKt1) << 16     swizzled tile offsets of the lane's B-fragment element per K slab
//   w[1] = (pB ^ pKt2) | (pB ^ pKt3) << 16
//   w[2] = (pD ^ pMt0) | (pD ^ pMt1) << 16     ... of its D-fragment element per M slab
//   w[3] = pK0                                 ... of the second D column
//   w[4..7] = index of the lane's A-fragment element in the canonical table, [mt][kt], one byte each
__device__ __forceinline__ void fillLaneTab(const BlockDesc& b, int lane, uint32_t* w) {
    const uint32_t pB = swz(laneOffB(b, lane));
    const uint32_t pD = swz(laneOffD(b, lane));
    const int ktl = b.k == 4 ? 4 : 2;
    const int mtl = b.k == 4 ? 2 : 1;
    uint32_t bk[4] = {0, 0, 0, 0}, dm[2] = {0, 0};
    for (int kt = 0; kt < ktl; ++kt) bk[kt] = pB ^ swz(ktOff(b, kt));
    for (int mt = 0; mt < mtl; ++mt) dm[mt] = pD ^ swz(mtOff(b, mt));
    w[0] = bk[0] | (bk[1] << 16);
    w[1] = bk[2] | (bk[3] << 16);
    w[2] = dm[0] | (dm[1] << 16);
    w[3] = swz(1u << b.kappa[0]);
    const int rows = 1 << b.k;
    for (int mt = 0; mt < 2; ++mt) {
        uint32_t lo = 0, hi = 0;
        for (int kt = 0; kt < 4; ++kt) {
            uint32_t at = 0;
            if (mt < mtl && kt < ktl) at = static_cast<uint32_t>(b.canon[8 * mt + (lane >> 2)] * rows + b.canon[4 * kt + (lane & 3)]);
            // up to 255 fits a byte (16 x 16 table)
            if (kt < 2) lo |= at << (16 * kt); else hi |= at << (16 * (kt - 2));
        }
        w[4 + 2 * mt] = lo;
        w[5 + 2 * mt] = hi;
    }
}

// NT units (each 2^K rows x 8 columns) per iteration: they share the A fragments (same matrix) and give the scheduler
// 6 NT independent DMMA chains per warp — the tensor pipe's result latency is about a hundred cycles, and a warp that
// waits for it between dependent DMMAs leaves the pipe idle (measured with NT = 1: 56 % pipe use in the compute phase).
template <int K, int WARPS, int NT>
__device__ __forceinline__ void applyBlockToTile(const BlockDesc& b, double2* __restrict__ tile, const uint32_t* __restrict__ laneTab,
                                                 const uint32_t* __restrict__ unitTab, uint32_t ctxOut, int warp, int lane) {
    constexpr int ROWS = 1 << K;
    constexpr int MT = ROWS / 8;  // 8-row slabs of M
    constexpr int KTL = ROWS / 4; // 4-column slabs of M
    const uint4 c0 = *reinterpret_cast<const uint4*>(laneTab + 8 * lane);
    const uint4 c1 = *reinterpret_cast<const uint4*>(laneTab + 8 * lane + 4);
    uint32_t pBk[KTL], pDm[MT];
    pBk[0] = c0.x & 0xffffu;
    pBk[1] = c0.x >> 16;
    if (KTL == 4) {
        pBk[KTL - 2] = c0.y & 0xffffu;
        pBk[KTL - 1] = c0.y >> 16;
    }
    pDm[0] = c0.z & 0xffffu;
    if (MT == 2) pDm[MT - 1] = c0.z >> 16;
    const uint32_t pK0 = c0.w;
    // canonical (row, column) of this lane's A fragment elements
    int aAt[MT][KTL];
    {
        const uint32_t words[4] = {c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int kt = 0; kt < KTL; ++kt) aAt[mt][kt] = static_cast<int>((words[2 * mt + (kt >> 1)] >> (16 * (kt & 1))) & 0xffffu);
        }
    }
    double aR[MT][KTL], aI[MT][KTL], aS[MT][KTL];
    uint32_t haveCtx = 0xffffffffu;
    const double2* table = reinterpret_cast<const double2*>(b.table);
    // contiguous ranges of NT-unit groups per warp: the context bits vary slowest, so a warp rarely changes its matrix
    // (the planner puts a context-free bit at unit bit 0 whenever NT = 2 is chosen: the units of a group share the matrix)
    const int nGroups = b.nUnits / NT;
    const int perWarp = (nGroups + WARPS - 1) / WARPS;
    const int g0 = warp * perWarp;
    const int g1 = min(nGroups, g0 + perWarp);
    for (int g = g0; g < g1; ++g) {
        uint32_t pu[NT];
#pragma unroll
        for (int v = 0; v < NT; ++v) pu[v] = unitTab[NT * g + v];
        const uint32_t ctx = (pu[0] >> 16) | ctxOut;
        if (ctx != haveCtx) {
            const double2* m = table + static_cast<size_t>(ctx) * ROWS * ROWS;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kt = 0; kt < KTL; ++kt) {
                    const double2 w = __ldg(m + aAt[mt][kt]);
                    aR[mt][kt] = w.x;
                    aI[mt][kt] = w.y;
                    aS[mt][kt] = w.x + w.y;
                }
            }
            haveCtx = ctx;
        }
#pragma unroll
        for (int v = 0; v < NT; ++v) pu[v] &= 0xffffu; // already swizzled
        double2 y[NT][KTL];
#pragma unroll
        for (int v = 0; v < NT; ++v) {
#pragma unroll
            for (int kt = 0; kt < KTL; ++kt) y[v][kt] = tile[pu[v] ^ pBk[kt]];
        }
        __syncwarp(); // in place: every lane holds its inputs before any lane overwrites them
        double p1[NT][MT][2], p2[NT][MT][2], p3[NT][MT][2];
#pragma unroll
        for (int v = 0; v < NT; ++v) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) p1[v][mt][0] = p1[v][mt][1] = p2[v][mt][0] = p2[v][mt][1] = p3[v][mt][0] = p3[v][mt][1] = 0.0;
        }
#pragma unroll
        for (int kt = 0; kt < KTL; ++kt) {
#pragma unroll
            for (int v = 0; v < NT; ++v) {
                const double ys = y[v][kt].x + y[v][kt].y;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    dmma884(p1[v][mt], aR[mt][kt], y[v][kt].x);
                    dmma884(p2[v][mt], aI[mt][kt], y[v][kt].y);
                    dmma884(p3[v][mt], aS[mt][kt], ys);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NT; ++v) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t at = pu[v] ^ pDm[mt];
                tile[at] = make_double2(p1[v][mt][0] - p2[v][mt][0], (p3[v][mt][0] - p1[v][mt][0]) - p2[v][mt][0]);
                tile[at ^ pK0] = make_double2(p1[v][mt][1] - p2[v][mt][1], (p3[v][mt][1] - p1[v][mt][1]) - p2[v][mt][1]);
            }
        }
    }
}

// D = A B + C with D and C in different registers (a chain that branches off another chain's result needs no copy)
__device__ __forceinline__ void dmma884From(double (&d)[2], double a, double b, const double (&c)[2]) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n" : "=d"(d[0]), "=d"(d[1]) : "d"(a), "d"(b), "d"(c[0]), "d"(c[1]));
}

// One block applied to resident tiles by a compute warp, with its constants (and, while the context does not change, its matrix
// as A fragments) held in registers between calls: a pass of ONE block sets the runner up once per CTA instead of once per tile.
// The B fragments of the next unit are fetched from shared memory before the tensor-core work of the current one.
template <int K, int WARPS, bool PREFETCH> struct BlockRunner {
    static constexpr int ROWS = 1 << K;
    static constexpr int MT = ROWS / 8;
    static constexpr int KTL = ROWS / 4;
    uint32_t pBk[KTL], pDm[MT], pK0;
    int aAt[MT][KTL];
    // A fragments of the three real products of Z = M Y:  K1 = Mr (Yr + Yi),  Re Z = K1 - (Mr + Mi) Yi,  Im Z = K1 + (Mi - Mr) Yr.
    // K1 is accumulated once (KTL DMMAs per M slab); the Re and Im chains both start from it (dmma884From) and end in the result:
    // 3 KTL MT DMMAs per unit like any three-product form, but no subtraction afterwards — of the 16 additions per 16 x 16 unit
    // (each waits for the FP64 pipe between the other warp's DMMAs) only the four Yr + Yi remain.
    double aR[MT][KTL], aN[MT][KTL], aD[MT][KTL];
    uint32_t haveCtx;
    const double2* table;
    uint32_t tableShared; // shared-memory address of the staged table, 0 when it is read from global memory
    int u0, u1;

    __device__ __forceinline__ void init(const BlockDesc& b, const uint32_t* __restrict__ laneTab, int warp, int lane) {
        const uint4 c0 = *reinterpret_cast<const uint4*>(laneTab + 8 * lane);
        const uint4 c1 = *reinterpret_cast<const uint4*>(laneTab + 8 * lane + 4);
        pBk[0] = c0.x & 0xffffu;
        pBk[1] = c0.x >> 16;
        if (KTL == 4) {
            pBk[KTL - 2] = c0.y & 0xffffu;
            pBk[KTL - 1] = c0.y >> 16;
        }
        pDm[0] = c0.z & 0xffffu;
        if (MT == 2) pDm[MT - 1] = c0.z >> 16;
        pK0 = c0.w;
        const uint32_t words[4] = {c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int kt = 0; kt < KTL; ++kt) aAt[mt][kt] = static_cast<int>((words[2 * mt + (kt >> 1)] >> (16 * (kt & 1))) & 0xffffu);
        }
        haveCtx = 0xffffffffu;
        table = reinterpret_cast<const double2*>(b.table);
        tableShared = 0;
        // contiguous unit ranges per warp (the context bits vary slowest, so a warp rarely changes its matrix), sizes differing
        // by at most one
        setRange(0, b.nUnits, warp, WARPS);
    }
    // this warp is member `member` of `members` warps that share the units [first, first + count)
    __device__ __forceinline__ void setRange(int first, int count, int member, int members) {
        const int base = count / members, rem = count % members;
        u0 = first + member * base + min(member, rem);
        u1 = u0 + base + (member < rem ? 1 : 0);
    }

    __device__ __forceinline__ void loadA(uint32_t ctx) {
        if (tableShared != 0) {
            const uint32_t m = tableShared + ctx * (ROWS * ROWS * 16u); // tableShared already points at this lane's column of the fragment rows
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kt = 0; kt < KTL; ++kt) {
                    double2 w;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(w.x), "=d"(w.y) : "r"(m + 512u * static_cast<uint32_t>(mt * KTL + kt)));
                    aR[mt][kt] = w.x;
                    aN[mt][kt] = -w.x - w.y;
                    aD[mt][kt] = w.y - w.x;
                }
            }
        } else {
            const double2* m = table + static_cast<size_t>(ctx) * ROWS * ROWS;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kt = 0; kt < KTL; ++kt) {
                    const double2 w = __ldg(m + aAt[mt][kt]);
                    aR[mt][kt] = w.x;
                    aN[mt][kt] = -w.x - w.y;
                    aD[mt][kt] = w.y - w.x;
                }
            }
        }
        haveCtx = ctx;
    }

    // tensor-core work of one unit: zr, zi <- Re, Im of M * y
    __device__ __forceinline__ void issue(const double2 (&y)[KTL], double (&zr)[MT][2], double (&zi)[MT][2]) {
        double k1[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) k1[mt][0] = k1[mt][1] = 0.0;
#pragma unroll
        for (int kt = 0; kt < KTL; ++kt) {
            const double ys = y[kt].x + y[kt].y;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma884(k1[mt], aR[mt][kt], ys);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            dmma884From(zr[mt], aN[mt][0], y[0].y, k1[mt]);
            dmma884From(zi[mt], aD[mt][0], y[0].x, k1[mt]);
        }
#pragma unroll
        for (int kt = 1; kt < KTL; ++kt) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                dmma884(zr[mt], aN[mt][kt], y[kt].y);
                dmma884(zi[mt], aD[mt][kt], y[kt].x);
            }
        }
    }
    __device__ __forceinline__ void finish(double2* __restrict__ tile, uint32_t pu, const double (&zr)[MT][2], const double (&zi)[MT][2]) {
        __syncwarp(); // in place: every lane holds this unit's inputs before any lane overwrites them
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t at = pu ^ pDm[mt];
            tile[at] = make_double2(zr[mt][0], zi[mt][0]);
            tile[at ^ pK0] = make_double2(zr[mt][1], zi[mt][1]);
        }
    }
    __device__ __forceinline__ void fetch(const double2* __restrict__ tile, uint32_t packed, double2 (&y)[KTL]) {
#pragma unroll
        for (int kt = 0; kt < KTL; ++kt) y[kt] = tile[(packed & 0xffffu) ^ pBk[kt]];
    }

    // Software pipeline over the warp's units, two accumulator sets: the tensor-core work of unit u + 1 is issued BEFORE the
    // subtractions and stores of unit u, and the B fragments of unit u + 2 are fetched meanwhile — the tensor pipe's result
    // latency (about a hundred cycles) and the shared-memory latency are covered by the warp's own next unit instead of by
    // other warps (there are only two compute warps per scheduler).
    __device__ __forceinline__ void run(double2* __restrict__ tile, const uint32_t* __restrict__ unitTab, uint32_t ctxOut) {
        if (u0 >= u1) return;
        double2 yA[KTL], yB[KTL];
        double aZr[MT][2], aZi[MT][2], bZr[MT][2], bZi[MT][2];
        uint32_t pkA = unitTab[u0], pkB = 0;
        fetch(tile, pkA, yA);
        if (u0 + 1 < u1) {
            pkB = unitTab[u0 + 1];
            fetch(tile, pkB, yB);
        }
        {
            const uint32_t ctx = (pkA >> 16) | ctxOut;
            if (ctx != haveCtx) loadA(ctx);
        }
        issue(yA, aZr, aZi);
        for (int u = u0; u < u1; u += 2) {
            // accumulators A hold unit u in flight; yB holds the inputs of unit u + 1
            const uint32_t puA = pkA & 0xffffu;
            if (u + 1 < u1) {
                const uint32_t ctx = (pkB >> 16) | ctxOut;
                if (ctx != haveCtx) loadA(ctx);
                issue(yB, bZr, bZi);
            }
            if (u + 2 < u1) {
                pkA = unitTab[u + 2];
                fetch(tile, pkA, yA);
            }
            finish(tile, puA, aZr, aZi);
            if (u + 1 >= u1) break;
            const uint32_t puB = pkB & 0xffffu;
            if (u + 2 < u1) {
                const uint32_t ctx = (pkA >> 16) | ctxOut;
                if (ctx != haveCtx) loadA(ctx);
                issue(yA, aZr, aZi);
            }
            if (u + 3 < u1) {
                pkB = unitTab[u + 3];
                fetch(tile, pkB, yB);
            }
            finish(tile, puB, bZr, bZi);
        }
    }
};

// ---- mbarrier helpers (CTA scope) ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
// the executing thread's earlier cp.async copies arrive on the barrier when they have landed (counted in the barrier's expected arrivals)
__device__ __forceinline__ void mbarArriveOnCopies(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity) {
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}

#ifndef FDD_BLOCK_COMPUTE_WARPS
#define FDD_BLOCK_COMPUTE_WARPS 8
#endif
#ifndef FDD_BLOCK_PREFETCH
#define FDD_BLOCK_PREFETCH 1
#endif
constexpr int kComputeWarps = FDD_BLOCK_COMPUTE_WARPS; // tensor-core warps of a CTA (two per scheduler, software-pipelined unit loop)
constexpr int kMemoryWarps = 4;  // warps that only move tiles between HBM and shared memory
constexpr int kBlockThreads = 32 * (kComputeWarps + kMemoryWarps);
static_assert(kComputeWarps == 8, "setmaxnreg works on groups of four warps; a warp-local pass gives each of eight warps an eighth of the tile");

// shared memory of the warp-specialised kernel: the tables as above, two mbarriers per buffer and the segment index of the
// first segment of every tile this CTA will own (tilesPerCta words)
__host__ __device__ inline size_t blockPassSmemWs(int tileBits, int nBuffers, int nBlocks, int maxUnits, uint32_t tilesPerCta) {
    return blockPassSmem(tileBits, nBuffers, nBlocks, maxUnits) + 16 * static_cast<size_t>(nBuffers) + 16 + 4 * static_cast<size_t>(tilesPerCta);
}
// the staged matrix tables follow, 16-byte aligned (PassParams::tableSmem holds offsets into this area)
__host__ __device__ inline size_t blockPassTableArea(int tileBits, int nBuffers, int nBlocks, int maxUnits, uint32_t tilesPerCta) {
    return (blockPassSmemWs(tileBits, nBuffers, nBlocks, maxUnits, tilesPerCta) + 15) & ~static_cast<size_t>(15);
}

// Warp-specialised pass.  Measured on B200: the copy-in / copy-out of a tile alone runs at 0.91 of the HBM copy peak, the
// tensor-core work of one 16 x 16 block alone takes about 0.6 of that time — but with every warp doing both in turn
// (load, barrier, compute, barrier, store) the two ADD UP: while all warps compute nobody issues memory instructions, and
// while they sit on full load/store queues the tensor pipe idles.  So the roles are split:
//   memory warps   copy tile i+1 in (cp.async, completion signalled through an mbarrier) and tile i-1 out (LDS + STG,
//                  blocking on the store queue as long as HBM needs) — the same thread owns the same shared-memory
//                  slots for the copy-out and the next copy-in, so no further ordering is needed;
//   compute warps  wait for "tile i has landed", run the pass's blocks on it (named barrier between blocks), signal
//                  "tile i is done".
// Three tile buffers (2^12 amplitudes each) keep copy-in, tensor work and copy-out of three consecutive tiles in flight;
// a 2^13 tile has one buffer and runs the three phases in turn.
__global__ void __launch_bounds__(kBlockThreads, 1) dmavm_block_ws_kernel(const __grid_constant__ PassParams p, int maxUnits, int nBuffers) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    double2* tiles = reinterpret_cast<double2*>(smemRaw);
    const uint32_t tileElems = 1u << p.tileBits;
    const uint32_t nSegTile = 1u << (p.tileBits - kLaneBits);
    uint2* segTab = reinterpret_cast<uint2*>(smemRaw + (static_cast<size_t>(16 * nBuffers) << p.tileBits));
    uint32_t* laneTabs = reinterpret_cast<uint32_t*>(segTab + nSegTile);
    uint32_t* unitTabs = laneTabs + p.nBlocks * 32 * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemRaw + ((blockPassSmem(p.tileBits, nBuffers, p.nBlocks, maxUnits) + 15) & ~static_cast<size_t>(15)));
    uint64_t* full = bars;             // [nBuffers] memory -> compute: the tile has landed
    uint64_t* done = bars + nBuffers;  // [nBuffers] compute -> memory: the blocks have run
    uint32_t* tileSeg = reinterpret_cast<uint32_t*>(bars + 2 * nBuffers); // [tiles of this CTA] local segment index of the tile's first segment
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- once per CTA: everything that only depends on the plan --------------------------------------------------
    for (uint32_t j = threadIdx.x; j < nSegTile; j += blockDim.x) segTab[j] = make_uint2(pdep32(j, p.tileMask), swz(j * 32u));
    for (int g = warp; g < p.nBlocks; g += kComputeWarps + kMemoryWarps) fillLaneTab(p.blocks[g], lane, laneTabs + (g * 32 + lane) * 8);
    for (int g = 0; g < p.nBlocks; ++g) {
        const BlockDesc& b = p.blocks[g];
        for (int u = threadIdx.x; u < b.nUnits; u += blockDim.x) {
            const uint32_t off = unitOff(b, static_cast<uint32_t>(u));
            uint32_t ctxIn = 0;
            for (int j = 0; j < b.nCtx; ++j) {
                const uint32_t src = b.ctxSrc[j];
                if (src < 32u) ctxIn |= ((off >> src) & 1u) << j;
            }
            unitTabs[g * maxUnits + u] = swz(off) | (ctxIn << 16);
        }
    }
    {
        uint32_t i = threadIdx.x;
        for (uint32_t t = blockIdx.x + i * gridDim.x; t < p.nTiles; t += blockDim.x * gridDim.x, i += blockDim.x) tileSeg[i] = spreadAround(t, p.tileMask);
    }
    unsigned char* tableArea = smemRaw + blockPassTableArea(p.tileBits, nBuffers, p.nBlocks, maxUnits, (p.nTiles + gridDim.x - 1) / gridDim.x);
    if (p.zPeer != nullptr && threadIdx.x == 0) {
        // fused exchange: this kernel runs after the shard's earlier launches, so the buffer the partner is about to write into has
        // been read to the end; nothing may be stored into the partner's buffer before the partner says the same
        if (blockIdx.x == 0) st_release_sys(p.exchMyFlags, p.exchEpoch);
        while (ld_acquire_sys(p.exchPartnerFlags) < p.exchEpoch) __nanosleep(64);
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < nBuffers; ++b) {
            mbarInit(full + b, 32 * kMemoryWarps);
            mbarInit(done + b, 32 * kComputeWarps); // every compute thread arrives: each one releases its own writes to the tile
        }
    }
    __syncthreads();
    // the matrix tables the pass keeps in shared memory, in A-FRAGMENT order: entry ((matrix * MT + mt) * KTL + kt) * 32 + lane is the
    // element lane `lane` feeds to the DMMAs of slab (mt, kt) (laneTabs words 4..7), so a warp reads a matrix as MT KTL
    // contiguous 512-byte rows without a bank conflict (in matrix order the eight rows of a fragment share their banks)
    for (int g = 0; g < p.nBlocks; ++g) {
        if (p.tableSmem[g] == kTableInGlobal) continue;
        const BlockDesc& b = p.blocks[g];
        const uint32_t perMatrix = 1u << (2 * b.k); // complex entries of a matrix = fragment slots (every entry belongs to one lane and slab)
        const int ktl = b.k == 4 ? 4 : 2;
        const uint32_t total = perMatrix << b.nCtx;
        const double2* src = reinterpret_cast<const double2*>(b.table);
        double2* dst = reinterpret_cast<double2*>(tableArea + p.tableSmem[g]);
        const uint32_t* lt = laneTabs + g * 32 * 8;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const uint32_t slot = i & (perMatrix - 1u);
            const uint32_t ln = slot & 31u, frag = slot >> 5; // frag = mt * KTL + kt
            const uint32_t mt = frag / static_cast<uint32_t>(ktl), kt = frag % static_cast<uint32_t>(ktl);
            const uint32_t at = (lt[ln * 8 + 4 + 2 * mt + (kt >> 1)] >> (16 * (kt & 1))) & 0xffffu;
            dst[i] = __ldg(src + (i - slot) + at);
        }
    }
    __syncthreads();
    const uint32_t laneSwz = swz(static_cast<uint32_t>(lane)); // swz is linear: swz(32 j + lane) = swz(32 j) ^ swz(lane)
    if (warp >= kComputeWarps) {
        // =================== memory warps ===================================================================
        // (they need few registers: hand the rest of the warpgroup's share to the compute warps, whose software pipeline
        // holds two accumulator sets, two sets of B fragments and the matrix)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        const int mw = warp - kComputeWarps;
        const double2* __restrict__ y = static_cast<const double2*>(p.y);
        double2* __restrict__ z = static_cast<double2*>(p.z);
        // (buffer index and barrier phase of iteration `it` are passed in: no division in the loop)
        auto copyIn = [&](int it, int buf) {
            double2* dst = tiles + static_cast<size_t>(buf) * tileElems;
            const double2* src = y + (static_cast<uint64_t>(tileSeg[it]) << kLaneBits) + lane;
            if (!(p.debugSkip & 2u)) {
#pragma unroll 4
                for (uint32_t j = mw; j < nSegTile; j += kMemoryWarps) {
                    const uint2 e = segTab[j];
                    cp_async16(dst + (e.y ^ laneSwz), src + (static_cast<uint64_t>(e.x) << kLaneBits));
                }
            }
            mbarArriveOnCopies(full + buf);
        };
        auto copyOut = [&](int it, int buf, unsigned phase) {
            mbarWait(done + buf, phase);
            if (p.debugSkip & 2u) return;
            const double2* src = tiles + static_cast<size_t>(buf) * tileElems;
            if (p.zPeer != nullptr) {
                // fused exchange: a segment whose traded bit differs from this shard's global bit now belongs to the partner, at the
                // index with that bit flipped
                double2* zFar = static_cast<double2*>(p.zPeer);
                const uint32_t base = tileSeg[it];
                const uint32_t bit = 1u << p.exchSegBit;
#pragma unroll 4
                for (uint32_t j = mw; j < nSegTile; j += kMemoryWarps) {
                    const uint2 e = segTab[j];
                    const uint32_t seg = base | e.x;
                    const bool stays = ((seg >> p.exchSegBit) & 1u) == p.exchMyBit;
                    double2* to = stays ? z + (static_cast<uint64_t>(seg) << kLaneBits) : zFar + (static_cast<uint64_t>(seg ^ bit) << kLaneBits);
                    st_stream(to + lane, src[e.y ^ laneSwz]);
                }
                return;
            }
            double2* dst = z + (static_cast<uint64_t>(tileSeg[it]) << kLaneBits) + lane;
#pragma unroll 4
            for (uint32_t j = mw; j < nSegTile; j += kMemoryWarps) {
                const uint2 e = segTab[j];
                st_stream(dst + (static_cast<uint64_t>(e.x) << kLaneBits), src[e.y ^ laneSwz]);
            }
        };
        int it = 0;
        int buf = 0, prevBuf = 0;          // buffers of tile `it` and of tile `it - 1`
        unsigned prevPhase = 0;            // barrier phase of tile `it - 1`: (it - 1) / nBuffers & 1
        unsigned phase = 0;
        for (uint32_t t = blockIdx.x;; ++it, t += gridDim.x) {
            const bool haveNext = t < p.nTiles;
            const bool havePrev = it > 0;
            if (!haveNext && !havePrev) break;
            if (nBuffers == 1) { // the only buffer: out before in
                if (havePrev) copyOut(it - 1, prevBuf, prevPhase);
                if (haveNext) copyIn(it, buf);
            } else {
                if (haveNext) copyIn(it, buf);
                if (havePrev) copyOut(it - 1, prevBuf, prevPhase);
            }
            if (!haveNext) break;
            prevBuf = buf;
            prevPhase = phase;
            if (++buf == nBuffers) {
                buf = 0;
                phase ^= 1u;
            }
        }
        cp_async_wait<0>();
        if (p.zPeer != nullptr) {
            // this thread's stores into the partner's buffer are performed before the CTA counts itself as finished; the last CTA
            // tells the partner and keeps the kernel alive until the partner's stores into THIS shard's buffer are complete too
            // (the next launch reads them)
            __threadfence_system();
            asm volatile("bar.sync 2, %0;\n" ::"n"(32 * kMemoryWarps) : "memory");
            if (threadIdx.x == 32 * kComputeWarps) {
                __threadfence();
                const unsigned int arrived = atomicAdd(p.exchCounter, 1u);
                if (arrived == gridDim.x - 1) {
                    *p.exchCounter = 0; // for the next exchange (every CTA has arrived)
                    __threadfence_system();
                    st_release_sys(p.exchMyFlags + 1, p.exchEpoch);
                    while (ld_acquire_sys(p.exchPartnerFlags + 1) < p.exchEpoch) __nanosleep(64);
                }
            }
        }
    } else {
        // =================== compute warps ==================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
        auto ctxOutOf = [&](const BlockDesc& b, uint32_t segBase) {
            uint32_t ctxOut = 0;
            for (int j = 0; j < b.nCtx; ++j) {
                const uint32_t src = b.ctxSrc[j];
                if (src >= 32u) ctxOut |= (((p.rankSegBits | segBase) >> (src - 32u)) & 1u) << j;
            }
            return ctxOut;
        };
        // a pass of one block keeps the block's constants and matrix in registers across the tiles
        auto singleBlockLoop = [&](auto kTag) {
            constexpr int K = decltype(kTag)::value;
            BlockRunner<K, kComputeWarps, FDD_BLOCK_PREFETCH != 0> runner;
            runner.init(p.blocks[0], laneTabs, warp, lane);
            bool anyCtxOut = false;
            for (int j = 0; j < p.blocks[0].nCtx; ++j) anyCtxOut = anyCtxOut || p.blocks[0].ctxSrc[j] >= 32u;
            int it = 0;
            int buf = 0;
            unsigned phase = 0;
            long long cWait = 0, cRun = 0;
            const long long cStart = clock64();
            for (uint32_t t = blockIdx.x; t < p.nTiles; t += gridDim.x, ++it) {
                double2* tile = tiles + static_cast<size_t>(buf) * tileElems;
                const uint32_t ctxOut = anyCtxOut ? ctxOutOf(p.blocks[0], tileSeg[it]) : 0u;
                const long long c0 = p.debugClocks != nullptr ? clock64() : 0;
                mbarWait(full + buf, phase);
                const long long c1 = p.debugClocks != nullptr ? clock64() : 0;
                if (!(p.debugSkip & 1u)) runner.run(tile, unitTabs, ctxOut);
                __syncwarp();
                if (p.debugClocks != nullptr) {
                    cWait += c1 - c0;
                    cRun += clock64() - c1;
                }
                mbarArrive(done + buf); // release: this thread's writes to the tile are visible to the memory warps that wait
                if (++buf == nBuffers) {
                    buf = 0;
                    phase ^= 1u;
                }
            }
            if (p.debugClocks != nullptr && lane == 0) {
                long long* out = p.debugClocks + (static_cast<size_t>(blockIdx.x) * kComputeWarps + warp) * 4;
                out[0] = cWait;
                out[1] = cRun;
                out[2] = clock64() - cStart;
                out[3] = it;
            }
        };
        if (p.nBlocks == 1) {
            if (p.blocks[0].k == 4) singleBlockLoop(std::integral_constant<int, 4>{});
            else singleBlockLoop(std::integral_constant<int, 3>{});
            return;
        }
        int it = 0;
        int buf = 0;
        unsigned phase = 0;
        for (uint32_t t = blockIdx.x; t < p.nTiles; t += gridDim.x, ++it) {
            double2* tile = tiles + static_cast<size_t>(buf) * tileElems;
            mbarWait(full + buf, phase);
            const uint32_t segBase = tileSeg[it];
            for (int g = 0; g < p.nBlocks && !(p.debugSkip & 1u); ++g) {
                const BlockDesc& b = p.blocks[g];
                const uint32_t ctxOut = ctxOutOf(b, segBase);
                if (g > 0) {
                    // the previous block has written what this one reads: this warp's own eighth of the tile when the pass is
                    // warp local, else anywhere in the tile
                    if (p.warpLocal) __syncwarp();
                    else asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kComputeWarps) : "memory");
                }
                auto runBlock = [&](auto kTag) {
                    BlockRunner<decltype(kTag)::value, kComputeWarps, FDD_BLOCK_PREFETCH != 0> runner;
                    runner.init(b, laneTabs + g * 32 * 8, warp, lane);
                    if (p.tableSmem[g] != kTableInGlobal) runner.tableShared = static_cast<uint32_t>(__cvta_generic_to_shared(tableArea + p.tableSmem[g])) + 16u * static_cast<uint32_t>(lane);
                    runner.run(tile, unitTabs + g * maxUnits, ctxOut);
                };
                if (b.k == 4) runBlock(std::integral_constant<int, 4>{});
                else runBlock(std::integral_constant<int, 3>{});
            }
            __syncwarp();
            mbarArrive(done + buf); // release: this thread's writes to the tile are visible to the memory warps that wait
            if (++buf == nBuffers) {
                buf = 0;
                phase ^= 1u;
            }
        }
    }
}

// Tiles up to 2^12 amplitudes are double buffered inside the CTA: the copies of the next tile are in flight while the blocks run
// on the current one, and the stores of the previous one drain (two CTAs per SM that merely alternate would run in lockstep:
// measured T = T_memory + T_tensor instead of the maximum).  A 2^13 tile fills the shared memory of an SM on its own.
template <int WARPS, int NT>
__global__ void __launch_bounds__(32 * WARPS, (WARPS == 16 || NT == 2) ? 1 : 2) dmavm_block_kernel(const __grid_constant__ PassParams p, int maxUnits, int nBuffers) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    double2* tiles = reinterpret_cast<double2*>(smemRaw);
    const uint32_t tileElems = 1u << p.tileBits;
    const uint32_t nSegTile = 1u << (p.tileBits - kLaneBits);
    uint2* segTab = reinterpret_cast<uint2*>(smemRaw + (static_cast<size_t>(16 * nBuffers) << p.tileBits));
    uint32_t* laneTabs = reinterpret_cast<uint32_t*>(segTab + nSegTile);
    uint32_t* unitTabs = laneTabs + p.nBlocks * 32 * 8;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- once per CTA: everything that only depends on the plan --------------------------------------------------
    for (uint32_t j = threadIdx.x; j < nSegTile; j += blockDim.x) segTab[j] = make_uint2(pdep32(j, p.tileMask), swz(j * 32u));
    for (int g = warp; g < p.nBlocks; g += WARPS) fillLaneTab(p.blocks[g], lane, laneTabs + (g * 32 + lane) * 8);
    for (int g = 0; g < p.nBlocks; ++g) {
        const BlockDesc& b = p.blocks[g];
        for (int u = threadIdx.x; u < b.nUnits; u += blockDim.x) {
            const uint32_t off = unitOff(b, static_cast<uint32_t>(u));
            uint32_t ctxIn = 0;
            for (int j = 0; j < b.nCtx; ++j) {
                const uint32_t src = b.ctxSrc[j];
                if (src < 32u) ctxIn |= ((off >> src) & 1u) << j;
            }
            unitTabs[g * maxUnits + u] = swz(off) | (ctxIn << 16);
        }
    }
    __syncthreads();
    const double2* __restrict__ y = static_cast<const double2*>(p.y);
    double2* __restrict__ z = static_cast<double2*>(p.z);
    const uint32_t laneSwz = swz(static_cast<uint32_t>(lane)); // swz is linear: swz(32 j + lane) = swz(32 j) ^ swz(lane)
    auto issueLoad = [&](uint32_t t, double2* dst) {
        if (p.debugSkip & 2u) return;
        const double2* src = y + (static_cast<uint64_t>(spreadAround(t, p.tileMask)) << kLaneBits) + lane;
#pragma unroll 4
        for (uint32_t j = warp; j < nSegTile; j += WARPS) {
            const uint2 e = segTab[j];
            cp_async16(dst + (e.y ^ laneSwz), src + (static_cast<uint64_t>(e.x) << kLaneBits));
        }
    };
    if (blockIdx.x < p.nTiles) issueLoad(blockIdx.x, tiles);
    cp_async_commit();
    int it = 0;
    for (uint32_t t = blockIdx.x; t < p.nTiles; t += gridDim.x, ++it) {
        double2* tile = tiles + ((nBuffers == 2 && (it & 1)) ? tileElems : 0u);
        const uint32_t tNext = t + gridDim.x;
        if (nBuffers == 2) {
            // the other buffer was drained by the store phase of the previous iteration (barrier at its end)
            if (tNext < p.nTiles) issueLoad(tNext, tiles + ((it & 1) ? 0u : tileElems));
            cp_async_commit(); // one group per iteration, possibly empty: the wait below counts groups
            cp_async_wait<1>(); // all but the newest group: this tile has landed
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t segBase = spreadAround(t, p.tileMask); // local segment index of the tile's first segment
        // ---- apply -----------------------------------------------------------------------------------------
        for (int g = 0; g < p.nBlocks && !(p.debugSkip & 1u); ++g) {
            const BlockDesc& b = p.blocks[g];
            uint32_t ctxOut = 0;
            for (int j = 0; j < b.nCtx; ++j) {
                const uint32_t src = b.ctxSrc[j];
                if (src >= 32u) ctxOut |= (((p.rankSegBits | segBase) >> (src - 32u)) & 1u) << j;
            }
            // (a block whose unit bits are all context bits cannot share a matrix between two units: one unit per iteration)
            const bool pairable = NT == 2 && b.nUnitBits > 0 && !b.unitBit0IsCtx;
            if (b.k == 4) {
                if (pairable) applyBlockToTile<4, WARPS, NT>(b, tile, laneTabs + g * 32 * 8, unitTabs + g * maxUnits, ctxOut, warp, lane);
                else applyBlockToTile<4, WARPS, 1>(b, tile, laneTabs + g * 32 * 8, unitTabs + g * maxUnits, ctxOut, warp, lane);
            } else {
                if (pairable) applyBlockToTile<3, WARPS, NT>(b, tile, laneTabs + g * 32 * 8, unitTabs + g * maxUnits, ctxOut, warp, lane);
                else applyBlockToTile<3, WARPS, 1>(b, tile, laneTabs + g * 32 * 8, unitTabs + g * maxUnits, ctxOut, warp, lane);
            }
            __syncthreads();
        }
        // ---- store -----------------------------------------------------------------------------------------
        if (!(p.debugSkip & 2u)) {
            double2* dst = z + (static_cast<uint64_t>(segBase) << kLaneBits) + lane;
#pragma unroll 4
            for (uint32_t j = warp; j < nSegTile; j += WARPS) {
                const uint2 e = segTab[j];
                st_stream(dst + (static_cast<uint64_t>(e.x) << kLaneBits), tile[e.y ^ laneSwz]);
            }
        }
        __syncthreads(); // the buffer is free for the next copies
        if (nBuffers == 1) {
            if (tNext < p.nTiles) issueLoad(tNext, tiles);
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
}

} // namespace fddb200

// tma_copy_probe.cu — A/B of the tile copy-in / copy-out of the dense-block kernel: cp.async (LDGSTS, 16 bytes per lane)
// + LDS/STG against TMA bulk copies (cp.async.bulk global->shared with an mbarrier transaction count, shared->global bulk
// stores).  Same tile decomposition as the kernel: tiles of 2^12 amplitudes (64 KiB) made of 512-byte segments whose
// segment-index bits are scattered (4 "target" bits) or contiguous; three tile buffers; the copy is out of place y -> z.
// The TMA variant has to use a LINEAR tile (a bulk copy cannot swizzle 16-byte units), so this measures what the data
// mover itself can do for this access pattern, not a drop-in replacement.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_copy_probe tools/tma_copy_probe.cu && build/tma_copy_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t pdep32(uint32_t x, uint32_t mask) {
    uint32_t out = 0;
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        if (x & 1u) out |= m & (0u - m);
        x >>= 1;
    }
    return out;
}
__device__ __forceinline__ uint32_t spreadAround(uint32_t x, uint32_t mask) {
    for (uint32_t m = mask; m != 0; m &= m - 1) {
        const uint32_t lowest = m & (0u - m);
        x = ((x & ~(lowest - 1u)) << 1) | (x & (lowest - 1u));
    }
    return x;
}

constexpr int kTileBits = 12, kSegs = 1 << (kTileBits - 5), kBuffers = 3;

// ---- variant A: the kernel's own data path (memory warps only: 4 warps) ------------------------------------------------
__global__ void __launch_bounds__(128) ldgsts_kernel(const double2* __restrict__ y, double2* __restrict__ z, uint32_t tileMask, uint32_t nTiles) {
    extern __shared__ __align__(128) double2 tiles[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int it = 0;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x, ++it) {
        double2* tile = tiles + static_cast<size_t>(it % kBuffers) * (1u << kTileBits);
        const uint64_t base = static_cast<uint64_t>(spreadAround(t, tileMask)) << 5;
        for (int j = warp; j < kSegs; j += 4) {
            const uint64_t off = static_cast<uint64_t>(pdep32(j, tileMask)) << 5;
            const unsigned s = smem_u32(tile + j * 32 + lane);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(y + base + off + lane) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        for (int j = warp; j < kSegs; j += 4) {
            const uint64_t off = static_cast<uint64_t>(pdep32(j, tileMask)) << 5;
            const double2 v = tile[j * 32 + lane];
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(z + base + off + lane), "d"(v.x), "d"(v.y) : "memory");
        }
    }
}

// ---- variant B: TMA bulk copies, one elected lane per warp issues them ----------------------------------------------------
__global__ void __launch_bounds__(128) bulk_kernel(const double2* __restrict__ y, double2* __restrict__ z, uint32_t tileMask, uint32_t nTiles, int runSegs) {
    extern __shared__ __align__(128) double2 tiles[];
    __shared__ __align__(8) uint64_t full[kBuffers];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int b = 0; b < kBuffers; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(full + b)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // runs of `runSegs` contiguous segments (the low tile bits are contiguous index bits): one bulk copy per run
    const int nRuns = kSegs / runSegs;
    const uint32_t runBytes = 512u * runSegs;
    int it = 0;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x, ++it) {
        const int b = it % kBuffers;
        double2* tile = tiles + static_cast<size_t>(b) * (1u << kTileBits);
        const uint64_t base = static_cast<uint64_t>(spreadAround(t, tileMask)) << 5;
        // the buffer's previous bulk stores must have read it
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.wait_group.read 2;\n" ::: "memory"); // groups complete in order: the one that read this buffer is three back
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(full + b)), "r"(64u * 1024u) : "memory");
        }
        __syncthreads();
        for (int r = warp * 32 + lane; r < nRuns; r += 128) {
            const uint64_t off = static_cast<uint64_t>(pdep32(r * runSegs, tileMask)) << 5;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(tile + r * runSegs * 32)),
                         "l"(y + base + off), "r"(runBytes), "r"(smem_u32(full + b))
                         : "memory");
        }
        // wait for the tile
        {
            const unsigned parity = static_cast<unsigned>(it / kBuffers) & 1u;
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(full + b)),
                "r"(parity)
                : "memory");
        }
        // copy out with bulk stores (the issuing thread owns the bulk group)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if (threadIdx.x == 0) {
            for (int r = 0; r < nRuns; ++r) {
                const uint64_t off = static_cast<uint64_t>(pdep32(r * runSegs, tileMask)) << 5;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(z + base + off), "r"(smem_u32(tile + r * runSegs * 32)), "r"(runBytes)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

int main() {
    const int n = 26;
    const size_t dim = size_t{1} << n;
    double2 *y = nullptr, *z = nullptr;
    cudaMalloc(&y, dim * sizeof(double2));
    cudaMalloc(&z, dim * sizeof(double2));
    cudaMemset(y, 1, dim * sizeof(double2));
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t smem = static_cast<size_t>(kBuffers) * 65536;
    cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const uint32_t nTiles = static_cast<uint32_t>(dim >> kTileBits);
    struct Shape {
        const char* name;
        uint32_t mask; // 7 segment-index bits of the tile
        int runSegs;   // contiguous segments per run
    };
    // scattered: 3 low contiguous bits + 4 target bits spread out (the usual single-block pass); contiguous: a 64 KiB contiguous tile
    const Shape shapes[] = {{"scattered_4_targets", 0x7u | (1u << 6) | (1u << 11) | (1u << 15) | (1u << 19), 8}, {"contiguous", 0x7fu, 128}, {"scattered_7_targets", (1u << 1) | (1u << 4) | (1u << 7) | (1u << 10) | (1u << 13) | (1u << 16) | (1u << 19), 1}};
    for (const Shape& sh : shapes) {
        for (int variant = 0; variant < 2; ++variant) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (variant == 0) ldgsts_kernel<<<sms, 128, smem>>>(y, z, sh.mask, nTiles);
                else bulk_kernel<<<sms, 128, smem>>>(y, z, sh.mask, nTiles, sh.runSegs);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            std::printf("{\"shape\": \"%s\", \"variant\": \"%s\", \"ms\": %.4f, \"gbs\": %.0f, \"err\": \"%s\"}\n", sh.name, variant == 0 ? "cp.async+LDS/STG (4 warps)" : "TMA bulk in/out",
                        best, 32.0 * dim / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}

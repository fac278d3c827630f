// dmma_probe.cu — what limits the FP64 tensor pipe in the dense-block kernel's unit loop?  Variants of one "unit"
// (Z(16x8) = M(16x16) Y(16x8), complex, three real products: 24 DMMA.8x8x4) with more and more of the real loop around them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// V: 1 = DMMAs only (distinct A registers, B from registers), 2 = + ys adds and the epilogue subtractions,
//    3 = + B fragments from shared memory and results back to shared memory (in place), 4 = 3 with two units per iteration
template <int V>
__global__ void __launch_bounds__(512) unit_kernel(double* out, int iters) {
    extern __shared__ double2 tile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tile[i] = make_double2(1e-3 * i, 1e-4 * i);
    __syncthreads();
    double aR[2][4], aI[2][4], aS[2][4];
    for (int mt = 0; mt < 2; ++mt)
        for (int kt = 0; kt < 4; ++kt) {
            aR[mt][kt] = 1e-3 * (lane + mt + kt);
            aI[mt][kt] = 1e-3 * (lane - mt - kt);
            aS[mt][kt] = aR[mt][kt] + aI[mt][kt];
        }
    constexpr int NT = V == 4 ? 2 : 1;
    double2 y[NT][4];
    for (int v = 0; v < NT; ++v)
        for (int kt = 0; kt < 4; ++kt) y[v][kt] = make_double2(1e-3 * (lane + kt + v), 1e-3 * (lane - kt));
    double2 sink = make_double2(0, 0);
    for (int it = 0; it < iters; ++it) {
        const int base = ((warp * 8 + (it & 7)) * 128 + lane) & 4095;
        if (V >= 3) {
#pragma unroll
            for (int v = 0; v < NT; ++v)
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) y[v][kt] = tile[(base + 32 * kt + 2048 * v) & 4095];
            __syncwarp();
        }
        double p1[NT][2][2], p2[NT][2][2], p3[NT][2][2];
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) p1[v][mt][0] = p1[v][mt][1] = p2[v][mt][0] = p2[v][mt][1] = p3[v][mt][0] = p3[v][mt][1] = 0.0;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
            for (int v = 0; v < NT; ++v) {
                const double ys = V >= 2 ? y[v][kt].x + y[v][kt].y : y[v][kt].x;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    dmma(p1[v][mt], aR[mt][kt], y[v][kt].x);
                    dmma(p2[v][mt], aI[mt][kt], y[v][kt].y);
                    dmma(p3[v][mt], aS[mt][kt], ys);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                double2 z0, z1;
                if (V >= 2) {
                    z0 = make_double2(p1[v][mt][0] - p2[v][mt][0], (p3[v][mt][0] - p1[v][mt][0]) - p2[v][mt][0]);
                    z1 = make_double2(p1[v][mt][1] - p2[v][mt][1], (p3[v][mt][1] - p1[v][mt][1]) - p2[v][mt][1]);
                } else {
                    z0 = make_double2(p1[v][mt][0], p2[v][mt][0] + p3[v][mt][0]);
                    z1 = make_double2(p1[v][mt][1], p2[v][mt][1] + p3[v][mt][1]);
                }
                if (V >= 3) {
                    tile[(base + 64 * mt + 2048 * v) & 4095] = z0;
                    tile[(base + 64 * mt + 32 + 2048 * v) & 4095] = z1;
                } else {
                    sink.x += z0.x + z1.x;
                    sink.y += z0.y + z1.y;
                    y[v][mt].x = z0.x * 1e-3; // keep the loop from being hoisted
                }
            }
    }
    if (V >= 3) sink = tile[threadIdx.x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink.x + sink.y;
}

// four real products per complex one: no additions at all, accumulators are the result; NT units per iteration
template <int NT>
__global__ void __launch_bounds__(512) unit4m_kernel(double* out, int iters) {
    extern __shared__ double2 tile[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tile[i] = make_double2(1e-3 * i, 1e-4 * i);
    __syncthreads();
    double aR[2][4], aI[2][4], aN[2][4];
    for (int mt = 0; mt < 2; ++mt)
        for (int kt = 0; kt < 4; ++kt) {
            aR[mt][kt] = 1e-3 * (lane + mt + kt);
            aI[mt][kt] = 1e-3 * (lane - mt - kt);
            aN[mt][kt] = -aI[mt][kt];
        }
    for (int it = 0; it < iters; ++it) {
        const int base = ((warp * 8 + (it & 7)) * 128 + lane) & 4095;
        double2 y[NT][4];
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) y[v][kt] = tile[(base + 32 * kt + 2048 * v) & 4095];
        __syncwarp();
        double zr[NT][2][2], zi[NT][2][2];
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) zr[v][mt][0] = zr[v][mt][1] = zi[v][mt][0] = zi[v][mt][1] = 0.0;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
            for (int v = 0; v < NT; ++v)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    dmma(zr[v][mt], aR[mt][kt], y[v][kt].x);
                    dmma(zi[v][mt], aI[mt][kt], y[v][kt].x);
                }
#pragma unroll
            for (int v = 0; v < NT; ++v)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    dmma(zr[v][mt], aN[mt][kt], y[v][kt].y);
                    dmma(zi[v][mt], aR[mt][kt], y[v][kt].y);
                }
        }
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                tile[(base + 64 * mt + 2048 * v) & 4095] = make_double2(zr[v][mt][0], zi[v][mt][0]);
                tile[(base + 64 * mt + 32 + 2048 * v) & 4095] = make_double2(zr[v][mt][1], zi[v][mt][1]);
            }
    }
    const double2 sink = tile[threadIdx.x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink.x + sink.y;
}

template <int NT> void run4m(const char* name, int sms, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(unit4m_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            unit4m_kernel<NT><<<sms, warps * 32, 65536>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double units = double(sms) * warps * iters * NT;
        std::printf("{\"variant\": \"%s\", \"ctas_per_sm\": 1, \"warps_per_cta\": %d, \"ms\": %.3f, \"cycles_per_unit_per_sm_at_1965MHz\": %.1f, \"dmma_per_unit\": 32, \"err\": \"%s\"}\n",
                    name, warps, best, best * 1e-3 * 1.965e9 / (units / sms), cudaGetErrorString(cudaGetLastError()));
    }
}

template <int V> void run(const char* name, int sms, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(unit_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int ctas = 1; ctas <= 2; ++ctas) {
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int iters = 20000;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                unit_kernel<V><<<sms * ctas, warps * 32, 65536>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double units = double(sms) * ctas * warps * iters * (V == 4 ? 2 : 1);
            const double dmmas = units * 24;
            std::printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"warps_per_cta\": %d, \"ms\": %.3f, \"cycles_per_dmma_per_sm_at_1965MHz\": %.2f, \"cycles_per_unit_per_sm_at_1965MHz\": %.1f, \"tflops\": %.2f, \"err\": \"%s\"}\n",
                        name, ctas, warps, best, best * 1e-3 * 1.965e9 / (dmmas / sms), best * 1e-3 * 1.965e9 / (units / sms), dmmas * 512 / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    }
}

int main() {
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    double* out = nullptr;
    cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 2 * 512);
    run<1>("dmma_only", prop.multiProcessorCount, out);
    run<2>("dmma_adds", prop.multiProcessorCount, out);
    run<3>("dmma_adds_smem", prop.multiProcessorCount, out);
    run<4>("dmma_adds_smem_2units", prop.multiProcessorCount, out);
    run4m<1>("four_products_smem", prop.multiProcessorCount, out);
    run4m<2>("four_products_smem_2units", prop.multiProcessorCount, out);
    return 0;
}

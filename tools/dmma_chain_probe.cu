// dmma_chain_probe.cu — FP64 tensor pipe (DMMA.8x8x4) throughput as a function of the number of independent accumulator
// chains a warp keeps in flight (C: every loop iteration issues one DMMA per chain, so a dependent DMMA follows C - 1 others)
// and of the warps per SM sub-partition.  Answers: how far apart must dependent DMMAs be?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_chain_probe tools/dmma_chain_probe.cu && build/dmma_chain_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int C>
__global__ void __launch_bounds__(512) chain_kernel(double* out, int iters) {
    const int lane = threadIdx.x & 31;
    double acc[C][2];
    double a[C], b[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        acc[c][0] = acc[c][1] = 0.0;
        a[c] = 1e-3 * (lane + c);
        b[c] = 1e-3 * (lane - c);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < C; ++c) dmma(acc[c], a[c], b[c]);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int C> void run(int sms, double* out) {
    for (int warps : {4, 8, 12, 16}) {
        const int iters = 24576 / C; // the same number of DMMAs per warp for every C
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        chain_kernel<C><<<sms, 32 * warps>>>(out, iters);
        cudaEventRecord(e0);
        chain_kernel<C><<<sms, 32 * warps>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double dmmaPerSmsp = static_cast<double>(iters) * C * warps / 4.0;
        const double cyc = ms * 1e-3 * 1.965e9;
        std::printf("{\"chains_per_warp\": %d, \"warps_per_smsp\": %d, \"ms\": %.3f, \"cycles_per_dmma_per_smsp_at_1965MHz\": %.2f, \"cycles_between_dependent_dmmas\": %.1f, \"err\": \"%s\"}\n",
                    C, warps / 4, ms, cyc / dmmaPerSmsp, cyc / iters, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * 512 * prop.multiProcessorCount);
    run<1>(prop.multiProcessorCount, out);
    run<2>(prop.multiProcessorCount, out);
    run<3>(prop.multiProcessorCount, out);
    run<4>(prop.multiProcessorCount, out);
    run<6>(prop.multiProcessorCount, out);
    run<8>(prop.multiProcessorCount, out);
    run<12>(prop.multiProcessorCount, out);
    return 0;
}

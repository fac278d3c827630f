// dadd_probe.cu — what does a DADD cost next to DMMAs on the FP64 pipe?  Per iteration and warp: D DMMAs (6 chains) and A
// additions (independent), two warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dadd_probe tools/dadd_probe.cu && build/dadd_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void daddv(double& x, double y) { asm volatile("add.rn.f64 %0, %0, %1;\n" : "+d"(x) : "d"(y)); }

template <int D, int A>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters) {
    const int lane = threadIdx.x & 31;
    double acc[6][2], a[6], b[6], s[16];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        acc[c][0] = acc[c][1] = 0.0;
        a[c] = 1e-3 * (lane + c);
        b[c] = 1e-3 * (lane - c);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 1e-6 * (lane + i);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < D; ++d) dmma(acc[d % 6], a[d % 6], b[(d + 1) % 6]);
#pragma unroll
        for (int i = 0; i < A; ++i) daddv(s[i % 16], b[i % 6]);
    }
    double r = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) r += acc[c][0] + acc[c][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int D, int A> void run(int sms, double* out) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) mix_kernel<D, A><<<sms, 256>>>(out, iters); // warm up, clocks
    cudaEventRecord(e0);
    mix_kernel<D, A><<<sms, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 2.0; // per iteration of ONE warp's work and scheduler (two warps per scheduler)
    std::printf("{\"dmma_per_iter\": %d, \"dadd_per_iter\": %d, \"warps_per_smsp\": 2, \"ms\": %.3f, \"cycles_per_iter_per_smsp_at_1965MHz\": %.1f, \"err\": \"%s\"}\n", D, A, ms, cyc,
                cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * 256 * prop.multiProcessorCount);
    const int sms = prop.multiProcessorCount;
    run<24, 0>(sms, out);
    run<0, 16>(sms, out);
    run<0, 64>(sms, out);
    run<24, 4>(sms, out);
    run<24, 16>(sms, out);
    run<24, 32>(sms, out);
    run<32, 0>(sms, out);
    return 0;
}

// fp64_probe.cu — measures the FP64 throughput of the CUDA-core pipe (DFMA) and of the tensor-core pipe
// (DMMA.8x8x4) on the device it runs on, to decide which one dense DMAVM blocks should use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_probe tools/fp64_probe.cu && /tmp/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9 + i;
    const double av = a + threadIdx.x * 1e-12, bv = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(av), "d"(bv));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out = nullptr;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int which = 0; which < 2; ++which) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) dfma_kernel<<<sms * 2, warps * 16>>>(out, iters, 1.0000001, 1e-9);
                else dmma_kernel<<<sms * 2, warps * 16>>>(out, iters, 1e-3, 1e-3);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            // flops: DFMA 2 per lane per instr, 16 instr per iteration; DMMA 2*8*8*4 per warp instr, 8 per iteration
            const double threads = double(sms) * 2 * warps * 16;
            const double flops = which == 0 ? threads * iters * 16.0 * 2.0 : (threads / 32.0) * iters * 8.0 * 512.0;
            std::printf("{\"probe\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", which == 0 ? "dfma" : "dmma_8x8x4", warps, best,
                        flops / (best * 1e-3) / 1e12);
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    return 0;
}

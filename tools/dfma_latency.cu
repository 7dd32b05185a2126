// Microbenchmark: FP64 FMA issue rate per SM sub-partition as a function of resident warps and of independent chains
// per warp (how much instruction-level parallelism the range kernel's polynomial loop needs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dfma_latency tools/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void chains(double *out, int iters, double a, double b) {
    double r[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) r[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CH; ++k) r[k] = fma(r[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += r[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int warps_per_sm, int sms, double *d) {
    const int iters = 1 << 15;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        chains<CH><<<sms, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = best * 1e-3 * clk_khz * 1e3;
    const double dfma_per_smsp = (double)iters * CH * warps_per_sm / 4.0;     // warp-instructions per sub-partition
    printf("chains %d  warps/SM %2d (%.1f per SMSP): %.3f ms, %.2f cycles per DFMA warp-instruction per SMSP, %.2f cycles per dependent step\n",
           CH, warps_per_sm, warps_per_sm / 4.0, best, cycles / dfma_per_smsp, cycles / iters);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 1024 * sizeof(double));
    printf("%s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    for (int w : {4, 8, 16, 32}) {
        run<1>(w, p.multiProcessorCount, d);
        run<2>(w, p.multiProcessorCount, d);
        run<4>(w, p.multiProcessorCount, d);
        run<8>(w, p.multiProcessorCount, d);
    }
    return 0;
}

// FP64-pipe granularity probe (tuning only): does a warp-instruction with 16 active lanes cost half
// a 32-lane one?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mb2 scripts/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
template <int ACTIVE>
__global__ void __launch_bounds__(256) k_dfma(double* out, double x, double y) {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    if ((threadIdx.x & 31) < ACTIVE) {
#pragma unroll 1
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = __fma_rn(a[i], x, y);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
void run(const char* name, K kern, int blocks) {
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        kern<<<blocks, 256>>>(out, 0.999, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r) best = best < ms ? best : ms;
    }
    const double warp_inst = 64.0 * ITERS * blocks * 8;  // FP64 warp-instructions
    printf("%-24s blocks=%4d  %8.3f ms  cycles/warp-inst/scheduler %.3f\n", name, blocks, best,
           best * 1e-3 * 1.965e9 / (warp_inst / (148.0 * 4)));
    cudaFree(out);
}
int main() {
    for (int bps : {1, 2}) {
        run("32 lanes active", k_dfma<32>, 148 * bps);
        run("16 lanes active (low)", k_dfma<16>, 148 * bps);
        run("8 lanes active", k_dfma<8>, 148 * bps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

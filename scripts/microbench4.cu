// Instruction-order / operand-slot probe of the Numerov step with the table read from shared memory
// (tuning only): which source form lets ptxas feed the 3-register DFMA from the operand reuse cache?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mb4 scripts/microbench4.cu
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ void step(double& X, double& S, double F, double ep) {
    const double fp = __dadd_rn(F, ep);
    const double Q  = __fma_rn(10.0, X, S);
    if (MODE == 0) {            // product kernel's order
        const double Xn = __fma_rn(-fp, Q, X);
        S = __dmul_rn(fp, X);
        X = Xn;
    } else if (MODE == 1) {     // DMUL first in source
        const double Sn = __dmul_rn(fp, X);
        X = __fma_rn(-fp, Q, X);
        S = Sn;
    } else if (MODE == 2) {     // one asm statement: mul then fma, fp first operand in both
        double Sn, Xn;
        asm("{\n\t.reg .f64 nf;\n\tmul.rn.f64 %0, %2, %3;\n\tneg.f64 nf, %2;\n\tfma.rn.f64 %1, nf, %4, %3;\n\t}"
            : "=&d"(Sn), "=&d"(Xn) : "d"(fp), "d"(X), "d"(Q));
        S = Sn; X = Xn;
    } else if (MODE == 3) {     // negated coefficient: fpn = -(F+ep) computed by the DADD itself
        const double fpn = __dsub_rn(-F, ep);           // -(F) - ep  == -(F+ep) exactly
        const double Sn = -__dmul_rn(fpn, X);
        X = __fma_rn(fpn, Q, X);
        S = Sn;
    }
}
template <int MODE, int E>
__global__ void __launch_bounds__(256) k(double* out, double ep0, int n) {
    __shared__ double tile[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = ep0 + 1e-12 * i;
    __syncthreads();
    double X[E], S[E], ep[E];
    for (int i = 0; i < E; i++) { X[i] = 1.0; S[i] = 0.0; ep[i] = ep0 * (1 + i + threadIdx.x); }
#pragma unroll 1
    for (int it = 0; it < n; it++) {
        const double2* t2 = reinterpret_cast<const double2*>(tile + ((it * 32) & 2047));
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const double2 ff = t2[p];
#pragma unroll
            for (int i = 0; i < E; i++) step<MODE>(X[i], S[i], ff.x, ep[i]);
#pragma unroll
            for (int i = 0; i < E; i++) step<MODE>(X[i], S[i], ff.y, ep[i]);
        }
    }
    double s = 0; for (int i = 0; i < E; i++) s += X[i] + S[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#include <cstdio>
template <typename K>
void run(const char* name, K kern, int E, int blocks) {
    double* out; cudaMalloc(&out, sizeof(double) * blocks * 256);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int n = 4096; float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a); kern<<<blocks, 256>>>(out, 1e-9, n); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = best < ms ? best : ms;
    }
    const double inst = 4.0 * 32 * n * E * blocks * 256;  // FP64 thread-instructions
    printf("%-40s blocks/SM=%d  %8.3f ms  pipe %.1f%%\n", name, blocks / 148, best, 100.0 * inst / (best * 1e-3) / (148.0 * 64 * 1.965e9));
    cudaFree(out);
}
int main() {
    for (int bps : {1, 2}) {
        run("mode0 product order            E=2", k<0,2>, 2, 148 * bps);
        run("mode1 DMUL first in source     E=2", k<1,2>, 2, 148 * bps);
        run("mode2 asm mul;fma              E=2", k<2,2>, 2, 148 * bps);
        run("mode3 negated coefficient      E=2", k<3,2>, 2, 148 * bps);
        run("mode0 product order            E=4", k<0,4>, 4, 148 * bps);
        run("mode1 DMUL first in source     E=4", k<1,4>, 4, 148 * bps);
        run("mode3 negated coefficient      E=4", k<3,4>, 4, 148 * bps);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

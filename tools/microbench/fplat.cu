// Dependent-issue latency of fp32 operations for ONE warp on an otherwise idle SM (what bounds the centre replay).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/fplat.cu -o tools/microbench/fplat
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, const float *in, long long *cyc, int n) {
    float M = in[0], D = in[1], yh = in[2], yl = in[3], v = in[4], dx = in[5];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (MODE == 0) {            // FADD chain
#pragma unroll
            for (int k = 0; k < 16; k++) M = __fadd_rn(M, yh);
        } else if (MODE == 1) {     // FFMA chain
#pragma unroll
            for (int k = 0; k < 16; k++) M = __fmaf_rn(M, yh, yl);
        } else if (MODE == 2) {     // the fast replay step: a = v - M; t = fma(-D, yl, dx); q = fma(a, yh, t); M += q; D += q
#pragma unroll
            for (int k = 0; k < 16; k++) {
                float a = __fsub_rn(v, M), t = __fmaf_rn(-D, yl, dx), q = __fmaf_rn(a, yh, t);
                M = __fadd_rn(M, q), D = __fadd_rn(D, q);
            }
        } else if (MODE == 3) {     // FADD -> FFMA -> FADD alternating chain (3 ops per step, no side chain)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                float a = __fsub_rn(v, M), q = __fmaf_rn(a, yh, yl);
                M = __fadd_rn(M, q);
            }
        } else if (MODE == 5) {     // the fast replay step with (M, D) as one packed fp32x2 pair: FFMA2, 2 FFMA, FADD2
            unsigned long long MD, NY, VX;
            asm("mov.b64 %0, {%1, %2};" : "=l"(MD) : "f"(M), "f"(D));
            asm("mov.b64 %0, {%1, %2};" : "=l"(NY) : "f"(-1.f), "f"(-yl));
            asm("mov.b64 %0, {%1, %2};" : "=l"(VX) : "f"(v), "f"(dx));
#pragma unroll
            for (int k = 0; k < 16; k++) {
                unsigned long long AT, QQ;
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(AT) : "l"(MD), "l"(NY), "l"(VX));
                float a, t;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(t) : "l"(AT));
                float q0 = __fmaf_rn(a, yh, t), q1 = __fmaf_rn(a, yh, t);
                asm("mov.b64 %0, {%1, %2};" : "=l"(QQ) : "f"(q0), "f"(q1));
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(MD) : "l"(MD), "l"(QQ));
            }
            asm("mov.b64 {%0, %1}, %2;" : "=f"(M), "=f"(D) : "l"(MD));
        } else {                    // the one-correction step (5 chain ops)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                float a = __fsub_rn(v, M), q = __fmul_rn(a, yh), r = __fmaf_rn(-dx, q, a), q1 = __fmaf_rn(r, yh, q);
                M = __fadd_rn(M, q1);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = M + D;
}
template <int MODE>
void run(const char *name, int ops) {
    float *out, *in;
    long long *cyc, h;
    cudaMalloc(&out, 4096), cudaMalloc(&in, 64), cudaMalloc(&cyc, 8);
    float hin[6] = {1.f, 0.001f, 1e-4f, 1e-12f, 1.5f, 1e-13f};
    cudaMemcpy(in, hin, 24, cudaMemcpyHostToDevice);
    const int n = 4096;
    for (int threads : {32, 128, 256}) {
        k<MODE><<<1, threads>>>(out, in, cyc, n);
        k<MODE><<<1, threads>>>(out, in, cyc, n);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %3d threads: %6.2f cycles per step (%d chain ops -> %.2f per op)\n", name, threads, (double)h / (n * 16.0), ops,
               (double)h / (n * 16.0) / ops);
    }
}
int main() {
    run<0>("FADD chain", 1);
    run<1>("FFMA chain", 1);
    run<3>("FADD -> FFMA -> FADD", 3);
    run<2>("fast replay step (3 ops + side chain)", 3);
    run<5>("fast replay step, (M, D) packed fp32x2", 3);
    run<4>("one-correction step (5 ops)", 5);
    return 0;
}

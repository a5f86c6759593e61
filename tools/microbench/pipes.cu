// Pipe-rate microbenchmark for the k_degree instruction mix on sm_100a:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs ILP independent dependency chains per thread; reports warp-instructions per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void k(float *out, float a, float b, unsigned long long *cyc) {
    float x[8], y[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i, y[i] = a + i;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));          // FFMA 3 distinct regs
            if (MODE == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));                        // FADD
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));                        // FMUL
            if (MODE == 3) asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(x[i]) : "f"(y[i]));                 // FFMA d = y*y + d
            if (MODE == 4) {  // packed: two fp32 FMAs per instruction
                unsigned long long p, q = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(b);
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[i]), "f"(y[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(y[i]) : "l"(p));
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);  // slowest warp of the grid
}

// packed chain kept in 64-bit registers (no repacking)
__global__ void k_packed(float *out, float a, float b, unsigned long long *cyc) {
    unsigned long long p[8], q = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(b);
    for (int i = 0; i < 8; i++) p[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(a + i);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q));
    }
    unsigned long long t1 = clock64();
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s);
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);  // slowest warp of the grid
}

// the k_degree test itself: 3 FADD, FMUL, 2 FFMA, FSETP, predicated IADD on 8 independent (query,candidate) pairs
__global__ void k_test(float *out, float a, float b, unsigned long long *cyc) {
    float qx[8], cx = a, cy = b, cz = a + b;
    int cnt[8];
    for (int i = 0; i < 8; i++) qx[i] = threadIdx.x * 0.001f + i, cnt[i] = 0;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float dx = __fsub_rn(qx[i], cx), dy = __fsub_rn(qx[i], cy), dz = __fsub_rn(qx[i], cz);
            float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[i]) : "f"(d), "f"(b));
        }
        cx += 1e-7f, cy += 2e-7f, cz -= 1e-7f;  // all three differences change every iteration (nothing to hoist)
    }
    unsigned long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 8; i++) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);  // slowest warp of the grid
}

// packed variant of the pair test: two candidates (cx0,cx1) per instruction against a duplicated query
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__global__ void k_test2(float *out, float a, float b, unsigned long long *cyc) {
    unsigned long long qq[8];      // query coordinate duplicated in both halves
    int cnt[8];
    for (int i = 0; i < 8; i++) qq[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.001f + i), cnt[i] = 0;
    unsigned long long cx = pk(a, a + 0.01f), cy = pk(b, b + 0.01f), cz = pk(a + b, a - b), step = pk(1e-7f, 2e-7f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long dx, dy, dz, d;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(qq[i]), "l"(cx));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(qq[i]), "l"(cy));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(qq[i]), "l"(cz));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dx));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
            float d0, d1;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[i]) : "f"(d0), "f"(b));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[i]) : "f"(d1), "f"(b));
        }
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cx) : "l"(step));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cy) : "l"(step));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cz) : "l"(step));
    }
    unsigned long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 8; i++) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);  // slowest warp of the grid
}

// variant: sign bits of (r2 - d) shifted into a mask with one funnel shift per test, POPC at the end
__global__ void k_test3(float *out, float a, float b, unsigned long long *cyc) {
    unsigned long long qq[8];
    unsigned mask[8];
    for (int i = 0; i < 8; i++) qq[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i), mask[i] = 0;
    unsigned long long cx = pk(a, a), cy = pk(b, b), cz = pk(a + b, a + b), rr = pk(b, b);
    int cnt = 0;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long dx, dy, dz, d;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(cx), "l"(qq[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(cy), "l"(qq[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(cz), "l"(qq[i]));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dx));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
            asm volatile("sub.rn.f32x2 %0, %1, %0;" : "+l"(d) : "l"(rr));
            unsigned t0b, t1b;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(t0b), "=r"(t1b) : "l"(d));
            asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(mask[i]) : "r"(t0b));
            asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(mask[i]) : "r"(t1b));
        }
        if ((it & 15) == 15) {
#pragma unroll
            for (int i = 0; i < 8; i++) cnt += 32 - __popc(mask[i]), mask[i] = 0;
        }
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cx) : "l"(rr));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cy) : "l"(rr));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cz) : "l"(rr));
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)cnt;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);  // slowest warp of the grid
}


// ---- round 2: the DOT form of the pair test (filter; the exact sequence only re-runs when a lane is inside the error band) ----
// v = qx*cx' + qy*cy' + qz*cz' + (|q|^2 - r^2) + |c|^2 with cx' = -2 cx ...: 3 FFMA2 + FADD2 per two tests; the sign bit of v is
// the verdict.  COUNT: 0 = cnt += bits >> 31 (LEA.HI), 1 = FSETP + predicated IADD, 2 = none (fp32 work only).
// DETECT: 1 = running minimum of |v| (FMNMX3 with |.| modifiers), 0 = none.
constexpr int NP = 4;  // query pairs per thread (the kernel holds 3 pairs against 2-4 candidates in flight)
template <int COUNT, int DETECT>
__global__ void __launch_bounds__(1024) k_dot(float *out, float a, float b, unsigned long long *cyc) {
    unsigned long long qx[NP], qy[NP], qz[NP], tq[NP];
    unsigned cnt[2 * NP];
    float m = 1e30f;
    for (int i = 0; i < NP; i++) {
        const float *src = out + 8 * i + threadIdx.x;  // opaque values: no register sharing between the operands
        qx[i] = pk(src[0], src[1]), qy[i] = pk(src[2], src[3]), qz[i] = pk(src[4], src[5]), tq[i] = pk(src[6], src[7]);
        cnt[2 * i] = cnt[2 * i + 1] = 0;
    }
    float cx = a, cy = b, cz = a + b, cn = a * b;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NP; i++) {
            unsigned long long t, cxx = pk(cx, cx), cyy = pk(cy, cy), czz = pk(cz, cz), cnn = pk(cn, cn);
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(qz[i]), "l"(czz), "l"(tq[i]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(qy[i]), "l"(cyy));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(qx[i]), "l"(cxx));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(t) : "l"(cnn));
            unsigned v0, v1;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(v0), "=r"(v1) : "l"(t));
            if (COUNT == 0) {
                cnt[2 * i] += v0 >> 31;
                cnt[2 * i + 1] += v1 >> 31;
            } else if (COUNT == 1) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, 0f00000000;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[2 * i]) : "f"(__uint_as_float(v0)));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, 0f00000000;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[2 * i + 1]) : "f"(__uint_as_float(v1)));
            } else {
                cnt[2 * i] ^= v0, cnt[2 * i + 1] ^= v1;  // keeps the result live with one LOP3 per test
            }
            if (DETECT) m = fminf(m, fminf(fabsf(__uint_as_float(v0)), fabsf(__uint_as_float(v1))));
        }
        cx += 1e-7f, cy += 2e-7f, cz -= 1e-7f, cn += 3e-7f;
    }
    unsigned long long t1 = clock64();
    unsigned s = 0;
    for (int i = 0; i < 2 * NP; i++) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s + m;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);
}

// the exact packed test with the counting done by LEA.HI on the sign of (d - r2): 7 packed + 2 integer per two tests
__global__ void __launch_bounds__(1024) k_test2b(float *out, float a, float b, unsigned long long *cyc) {
    unsigned long long qq[8];
    unsigned cnt[16];
    for (int i = 0; i < 8; i++) qq[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i), cnt[2 * i] = cnt[2 * i + 1] = 0;
    unsigned long long cx = pk(a, a), cy = pk(b, b), cz = pk(a + b, a + b), rr = pk(b, b), step = pk(1e-7f, 1e-7f);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long dx, dy, dz, d;
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(cx), "l"(qq[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(cy), "l"(qq[i]));
            asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(cz), "l"(qq[i]));
            asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dx));
            asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
            asm volatile("sub.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(rr));
            unsigned v0, v1;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(v0), "=r"(v1) : "l"(d));
            cnt[2 * i] += v0 >> 31, cnt[2 * i + 1] += v1 >> 31;
        }
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cx) : "l"(step));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cy) : "l"(step));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(cz) : "l"(step));
    }
    unsigned long long t1 = clock64();
    unsigned s = 0;
    for (int i = 0; i < 16; i++) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);
}


// ---- round 2: SYMMETRIC counting.  Every unordered pair is tested once: a hit counts for the lane's query (register) and for the
// candidate.  The candidate's share is the growth of the lane's slot counters (IADD3 sums), reduced over the warp with one
// REDUX.SUM and added to the candidate's degree with one RED by lane 0.  P query pairs per lane, one candidate per inner step.
template <int P, int SYM>
__global__ void __launch_bounds__(1024) k_sym(float *out, int *deg, float a, float b, unsigned long long *cyc) {
    unsigned long long qx[P], qy[P], qz[P];
    int cnt[2 * P];
    for (int i = 0; i < P; i++) {
        const float *src = out + 6 * i + threadIdx.x;
        qx[i] = pk(src[0], src[1]), qy[i] = pk(src[2], src[3]), qz[i] = pk(src[4], src[5]);
        cnt[2 * i] = cnt[2 * i + 1] = 0;
    }
    float cx = a, cy = b, cz = a + b;
    int prev = 0;
    unsigned packed = 0;
    int *dst = deg + (blockIdx.x * 32 + (threadIdx.x >> 5)) * 64;
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {   // four candidates per trip, as in k_degree
            unsigned long long cxx = pk(cx, cx), cyy = pk(cy, cy), czz = pk(cz, cz);
#pragma unroll
            for (int i = 0; i < P; i++) {
                unsigned long long dx, dy, dz, d;
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(cxx), "l"(qx[i]));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(cyy), "l"(qy[i]));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(czz), "l"(qz[i]));
                asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dx));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
                float d0, d1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[2 * i]) : "f"(d0), "f"(b));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt[2 * i + 1]) : "f"(d1), "f"(b));
            }
            if (SYM == 1) {
                int tot = 0;
#pragma unroll
                for (int i = 0; i < 2 * P; i++) tot += cnt[i];
                int hits = __reduce_add_sync(0xffffffffu, tot - prev);
                prev = tot;
                if ((threadIdx.x & 31) == 0) atomicAdd(dst + ((4 * it + c) & 63), hits);
            }
            if (SYM == 2) {  // a lane's hits per candidate are <= 2P <= 6: four candidates share one register (one byte each),
                int tot = 0; //  the warp sum of a byte is <= 192
#pragma unroll
                for (int i = 0; i < 2 * P; i++) tot += cnt[i];
                packed += (unsigned)(tot - prev) << (8 * c);
                prev = tot;
            }
            cx += 1e-7f, cy += 2e-7f, cz -= 1e-7f;
        }
        if (SYM == 2) {
            unsigned r = __reduce_add_sync(0xffffffffu, packed);
            packed = 0;
            const int l = threadIdx.x & 31;
            if (l < 4) atomicAdd(dst + ((4 * it + l) & 63), (int)((r >> (8 * l)) & 0xffu));
        }
    }
    unsigned long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 2 * P; i++) s += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    if ((threadIdx.x & 31) == 0) atomicMax(cyc, t1 - t0);
}

// every launch is preceded by a reset of the cycle counter (atomicMax over all warps)
#define RESET() cudaMemset(cyc, 0, 8)
int main() {
    float *out;
    unsigned long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
    cudaMalloc(&cyc, 8);
    int *deg;
    cudaMalloc(&deg, 148 * 32 * 64 * sizeof(int));
    cudaMemset(deg, 0, 148 * 32 * 64 * sizeof(int));
    const char *names[] = {"FFMA x=x*a+b", "FADD", "FMUL", "FFMA x=y*y+x", "FFMA2 with repack (3 instr)"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32;  // per SM (one block per SM)
        printf("--- %d warps per SM\n", warps);
#define RUN(MODE) RESET(); k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%-30s %.3f warp-instr/cycle/SM\n", names[MODE], (double)ITERS * 8 * warps * (MODE == 4 ? 3 : 1) / h);
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
        RESET(); k_packed<<<148, threads>>>(out, 1.0001f, 0.5f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %.3f warp-instr/cycle/SM  (= %.3f fp32 FMA-lanes x32 /cycle)\n", "FFMA2 (64-bit regs)", (double)ITERS * 8 * warps / h, 2.0 * ITERS * 8 * warps / h);
        RESET(); k_test<<<148, threads>>>(out, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %.3f warp-tests/cycle/SM  (8 instr each -> %.3f instr/cycle/SM)\n", "k_degree pair test", (double)ITERS * 8 * warps / h, 8.0 * ITERS * 8 * warps / h);
        RESET(); k_test2<<<148, threads>>>(out, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %.3f warp-tests/cycle/SM  (packed f32x2, 5 instr per test)\n", "pair test, f32x2", (double)ITERS * 16 * warps / h);
        RESET(); k_test3<<<148, threads>>>(out, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %.3f warp-tests/cycle/SM  (packed + sign-bit funnel shift, 4.5 instr per test)\n", "pair test, f32x2+SHF", (double)ITERS * 16 * warps / h);

#define RUND(C, D, label) RESET(); k_dot<C, D><<<148, threads>>>(out, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%-30s %.3f warp-tests/cycle/SM  (%s)\n", "pair test, DOT form", (double)ITERS * 2 * NP * warps / h, label);
        RUND(2, 0, "3 FFMA2 + FADD2 + 2 LOP3 per two tests: fp32 work only")
        RUND(0, 0, "cnt += bits >> 31, no detection")
        RUND(1, 0, "FSETP + @p IADD, no detection")
        RUND(0, 1, "cnt += bits >> 31, min|v| detection")
        RUND(1, 1, "FSETP + @p IADD, min|v| detection")
        RESET(); k_test2b<<<148, threads>>>(out, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s %.3f warp-tests/cycle/SM  (exact packed test, 7 packed + 2 LEA.HI per two tests)\n", "pair test, f32x2+LEA", (double)ITERS * 16 * warps / h);

#define RUNS(P, SYM, label) RESET(); k_sym<P, SYM><<<148, threads>>>(out, deg, 1.0001f, 0.0016f, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%-30s %.3f warp-tests/cycle/SM  (%s)\n", "pair test, candidate loop", (double)ITERS * 4 * 2 * P * warps / h, label);
        RUNS(2, 0, "P=2, one-sided (the product's loop without the loads)")
        RUNS(2, 1, "P=2, symmetric: + 2 IADD3 + REDUX + RED per candidate; every test counts twice")
        RUNS(3, 0, "P=3, one-sided")
        RUNS(3, 1, "P=3, symmetric")
        RUNS(2, 2, "P=2, symmetric, four candidates per REDUX (byte-packed)")
        RUNS(3, 2, "P=3, symmetric, four candidates per REDUX (byte-packed)")
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

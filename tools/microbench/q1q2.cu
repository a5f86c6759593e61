// How often does the one-correction quotient q1 differ from the correctly rounded q2 = RN(a/n)?  (k_centres runs its chain
// on q1 and verifies against q2 off the chain; a mismatch costs a replay of up to 1120 members.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned long long n_samples, unsigned long long *bad, unsigned long long *tot) {
    unsigned long long b = 0, t = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long h = (i + 12345) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32; h *= 0x94D049BB133111EBULL; h ^= h >> 29;
        unsigned lo = (unsigned)h, hi = (unsigned)(h >> 32);
        int n = 1 + (int)(hi % 200000u);
        float fn = (float)n;
        // a = x - M with x, M coordinates of a few metres: magnitudes 1e-4 .. 1
        float a = __uint_as_float((lo & 0x80000000u) | ((unsigned)(127 - 14 + (int)((lo >> 23) % 14u)) << 23) | (lo & 0x7fffffu));
        float y = __frcp_rn(fn);
        float q = __fmul_rn(a, y);
        float r = __fmaf_rn(-fn, q, a);
        float q1 = __fmaf_rn(r, y, q);
        float r1 = __fmaf_rn(-fn, q1, a);
        float q2 = __fmaf_rn(r1, y, q1);
        b += q1 != q2;
        t++;
    }
    atomicAdd(bad, b);
    atomicAdd(tot, t);
}
int main() {
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    k<<<148 * 8, 256>>>(1ull << 30, d, d + 1);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("q1 != q2 in %llu of %llu samples (%.3e)\n", h[0], h[1], (double)h[0] / (double)h[1]);
    return 0;
}

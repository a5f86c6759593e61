// What does the exact centre replay (k_centres) cost per member and per chunk?  One segment of N points, one cluster whose
// members are a pseudo-random fraction of the segment; kernel time (CUDA events) and the in-kernel clock64 statistics.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I pbnet_b200/csrc tools/microbench/centres.cu -o tools/microbench/centres
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "pb_kernels.cuh"

int main() {
    const int Ns[] = {20000, 200000};
    const int dens[] = {1, 2, 4, 8};   // one point in `den` is a member
    printf("chunk = %d points\n", pb::kCtrChunkPts);
    for (int N : Ns)
        for (int den : dens) {
            std::vector<float> x(N), y(N), z(N);
            std::vector<int> id(N);
            unsigned s = 12345u;
            int members = 0;
            for (int i = 0; i < N; i++) {
                s = s * 1664525u + 1013904223u;
                x[i] = 3.f + 0.05f * ((s >> 8) & 0xffff) / 65536.f;
                s = s * 1664525u + 1013904223u;
                y[i] = 5.f + 0.05f * ((s >> 8) & 0xffff) / 65536.f;
                s = s * 1664525u + 1013904223u;
                z[i] = 1.f + 0.05f * ((s >> 8) & 0xffff) / 65536.f;
                id[i] = ((s >> 3) % den == 0) ? 0 : -1;
                members += id[i] == 0;
            }
            float M[3] = {0, 0, 0};
            int cnt = 0;
            for (int i = 0; i < N; i++)
                if (id[i] == 0) {
                    cnt++;
                    M[0] += (x[i] - M[0]) / (float)cnt, M[1] += (y[i] - M[1]) / (float)cnt, M[2] += (z[i] - M[2]) / (float)cnt;
                }
            float *dx, *dy, *dz, *dc;
            int *did, *dstart, *didb, *dclt, *dK, *dticket;
            unsigned long long *dstats;
            cudaMalloc(&dx, N * 4), cudaMalloc(&dy, N * 4), cudaMalloc(&dz, N * 4), cudaMalloc(&did, N * 4), cudaMalloc(&dc, 12);
            cudaMalloc(&dstart, 8), cudaMalloc(&didb, 4), cudaMalloc(&dclt, 4), cudaMalloc(&dK, 4), cudaMalloc(&dticket, 4), cudaMalloc(&dstats, 64);
            cudaMemcpy(dx, x.data(), N * 4, cudaMemcpyHostToDevice), cudaMemcpy(dy, y.data(), N * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dz, z.data(), N * 4, cudaMemcpyHostToDevice), cudaMemcpy(did, id.data(), N * 4, cudaMemcpyHostToDevice);
            int st[2] = {0, N}, zero = 0, one = 1;
            cudaMemcpy(dstart, st, 8, cudaMemcpyHostToDevice), cudaMemcpy(didb, &zero, 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dclt, &zero, 4, cudaMemcpyHostToDevice), cudaMemcpy(dK, &one, 4, cudaMemcpyHostToDevice);
            pb::SegArrays sg{};
            sg.start = dstart, sg.id_base = didb;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0), cudaEventCreate(&e1);
            float best = 1e9f;
            unsigned long long hs[8];
            for (int rep = 0; rep < 5; rep++) {
                cudaMemset(dticket, 0, 4), cudaMemset(dstats, 0, 64);
                cudaEventRecord(e0);
                pb::k_centres<<<4, 256>>>(dK, sg, dclt, did, dx, dy, dz, dc, dticket, rep == 4 ? dstats : nullptr);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep < 4 && ms < best) best = ms;
            }
            cudaMemcpy(hs, dstats, 64, cudaMemcpyDeviceToHost);
            float hc[3];
            cudaMemcpy(hc, dc, 12, cudaMemcpyDeviceToHost);
            const int chunks = (N + pb::kCtrChunkPts - 1) / pb::kCtrChunkPts;
            printf("N %7d members %7d chunks %4d: %8.1f us = %6.2f ns/member, %6.2f us/chunk; replay cycles/member %.1f gather cycles/chunk %.0f "
                   "halves %llu replayed %llu  exact %d\n",
                   N, members, chunks, best * 1e3, best * 1e6 / members, best * 1e3 / chunks, (double)hs[2] / members,
                   (double)hs[3] / chunks, hs[0], hs[1], (int)(hc[0] == M[0] && hc[1] == M[1] && hc[2] == M[2]));
            cudaFree(dx), cudaFree(dy), cudaFree(dz), cudaFree(did), cudaFree(dc), cudaFree(dstart), cudaFree(didb), cudaFree(dclt), cudaFree(dK), cudaFree(dticket), cudaFree(dstats);
        }
    return 0;
}

// pb_normals.cuh — mesh vertex normals: the fourth op of the reference's PB_lib module
// (lib/PB_lib/src/normal/cal_normal.cu, bound at PB_lib_api.cpp:10).
//
// Reference: one thread per vertex scans ALL faces (O(V*F)) and adds up the area-weighted normals of the faces that list
// it, in face order.  Here: one sort of the (vertex, face) incidences gives every vertex its faces in ascending order;
// one thread per vertex then replays the same fp32 recurrence over its own run (O(F log F) total).  The arithmetic is
// pinned with intrinsics to what nvcc contracts the reference source to (SASS of oracle/_ref): cross component =
// fma(u, v, -(w*t)); dot(n,n) = fma(z,z, fma(y,y, x*x)) but |v| = sqrt(fma(z,z, fma(x,x, y*y))); s = fma(n, area, s);
// IEEE sqrt and division.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbn {

__device__ __forceinline__ float norm3(float x, float y, float z) {
    // linalg_two_norm as nvcc contracts it: the MIDDLE product is the plain multiply (same shape as the grouping predicate)
    return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y))));
}

// cal_normal.cu:43-76  face normal and "area" (= |n|^2 / 2, as the reference defines it); incidence keys (vertex, face)
__global__ void k_face_normals(const float *__restrict__ xyz, const int *__restrict__ face, int num_face, int num_vtx,
                               float *__restrict__ fnormal, float *__restrict__ farea, uint64_t *__restrict__ key,
                               int *__restrict__ err) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= num_face) return;
    int ia = face[3 * f], ib = face[3 * f + 1], ic = face[3 * f + 2];
    if (ia < 0 || ib < 0 || ic < 0 || ia >= num_vtx || ib >= num_vtx || ic >= num_vtx) {
        atomicOr(err, 1);
        key[3 * f] = key[3 * f + 1] = key[3 * f + 2] = ~0ull;
        return;
    }
    float ax = __fsub_rn(xyz[3 * ib], xyz[3 * ia]), ay = __fsub_rn(xyz[3 * ib + 1], xyz[3 * ia + 1]),
          az = __fsub_rn(xyz[3 * ib + 2], xyz[3 * ia + 2]);
    float bx = __fsub_rn(xyz[3 * ic], xyz[3 * ia]), by = __fsub_rn(xyz[3 * ic + 1], xyz[3 * ia + 1]),
          bz = __fsub_rn(xyz[3 * ic + 2], xyz[3 * ia + 2]);
    float nx = __fmaf_rn(ay, bz, -__fmul_rn(az, by));
    float ny = __fmaf_rn(az, bx, -__fmul_rn(ax, bz));
    float nz = __fmaf_rn(ax, by, -__fmul_rn(ay, bx));
    farea[f] = __fmul_rn(__fmaf_rn(nz, nz, __fmaf_rn(ny, ny, __fmul_rn(nx, nx))), 0.5f);
    float l = norm3(nx, ny, nz);
    fnormal[3 * f] = __fdiv_rn(nx, l);
    fnormal[3 * f + 1] = __fdiv_rn(ny, l);
    fnormal[3 * f + 2] = __fdiv_rn(nz, l);
    // a face that lists a vertex twice still counts once for it (cal_normal.cu:88)
    key[3 * f] = ((uint64_t)(unsigned)ia << 32) | (unsigned)f;
    key[3 * f + 1] = ib == ia ? ~0ull : (((uint64_t)(unsigned)ib << 32) | (unsigned)f);
    key[3 * f + 2] = (ic == ia || ic == ib) ? ~0ull : (((uint64_t)(unsigned)ic << 32) | (unsigned)f);
}

// cal_normal.cu:78-112  per vertex: area-weighted sum over its faces in ascending face order, normalised
__global__ void k_vertex_normals(int num_vtx, long long n_keys, const uint64_t *__restrict__ skey,
                                 const float *__restrict__ fnormal, const float *__restrict__ farea,
                                 float *__restrict__ out) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= num_vtx) return;
    long long lo = 0, hi = n_keys;
    const uint64_t first = (uint64_t)(unsigned)u << 32;
    while (lo < hi) {  // first incidence of vertex u
        long long mid = (lo + hi) >> 1;
        if (skey[mid] < first) lo = mid + 1;
        else hi = mid;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f, sa = 0.f;
    for (long long j = lo; j < n_keys; j++) {
        uint64_t k = skey[j];
        if ((k >> 32) != (uint64_t)(unsigned)u) break;
        int f = (int)(unsigned)k;
        float a = farea[f];
        sx = __fmaf_rn(fnormal[3 * f], a, sx);
        sy = __fmaf_rn(fnormal[3 * f + 1], a, sy);
        sz = __fmaf_rn(fnormal[3 * f + 2], a, sz);
        sa = __fadd_rn(sa, a);
    }
    if (sa == 0.0f) {
        out[3 * u] = 0.f, out[3 * u + 1] = 0.f, out[3 * u + 2] = 1.f;
    } else {
        float x = __fdiv_rn(sx, sa), y = __fdiv_rn(sy, sa), z = __fdiv_rn(sz, sa);
        float l = norm3(x, y, z);
        out[3 * u] = __fdiv_rn(x, l), out[3 * u + 1] = __fdiv_rn(y, l), out[3 * u + 2] = __fdiv_rn(z, l);
    }
}

}  // namespace pbn

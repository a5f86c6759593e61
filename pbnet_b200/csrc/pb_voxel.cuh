// pb_voxel.cuh — voxelize / devoxelize scatter-gather around the sparse-conv backbone (SURVEY.md §8 rows
// a12-a14).  The reference delegates these to MinkowskiEngine (un-vendored, un-pinned, absent here):
//   ME.utils.sparse_quantize(xyz, feats, quantization_size, return_index, return_inverse)
//        datasets/scannetv2/dataset_preprocess.py:269-274,348-353
//   ME.SparseTensor(features, coordinates=batched_coordinates(xyz/0.02)).inverse_mapping   network/PBNet.py:236-247
//   X_v[v2p] gathers and their autograd scatter-add                                          network/PBNet.py:130-134,250
// Contract implemented (ME >= 0.5 documentation): voxel = floor(x / size) as int32; unique rows;
// coords[index][inverse] == coords; features picked at `index` (default) or averaged (UNWEIGHTED_AVERAGE).
// ME leaves the voxel ORDER implementation-defined; here it is lexicographic (batch, x, y, z) and the
// representative of a voxel is its smallest point index — deterministic, equal to numpy.unique(axis=0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbv {

constexpr unsigned kFull = 0xffffffffu;

// K-V1  quantise; per-axis min / max (warp-reduced atomics)
template <class T>
__global__ void k_quantize(const T *__restrict__ coords, int stride, int has_batch_col, const int *__restrict__ batch,
                           long long n, T size, int4 *__restrict__ q, int *__restrict__ mn, int *__restrict__ mx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    int4 v = make_int4(0, 0, 0, 0);
    if (valid) {
        const T *p = coords + i * stride;
        int o = has_batch_col ? 1 : 0;
        T a = p[o], b = p[o + 1], c = p[o + 2];
        if (size > (T)0) {
            a = a / size;  // true division, as ME does (np.floor(coords / size), torch.floor(coords / size))
            b = b / size;
            c = c / size;
        }
        v.x = batch ? batch[i] : (has_batch_col ? (int)p[0] : 0);
        v.y = (int)floor(a);
        v.z = (int)floor(b);
        v.w = (int)floor(c);
        q[i] = v;
    }
    unsigned act = __ballot_sync(kFull, valid);
    if (!valid) return;
    int m0 = __reduce_min_sync(act, v.x), m1 = __reduce_min_sync(act, v.y), m2 = __reduce_min_sync(act, v.z),
        m3 = __reduce_min_sync(act, v.w);
    int M0 = __reduce_max_sync(act, v.x), M1 = __reduce_max_sync(act, v.y), M2 = __reduce_max_sync(act, v.z),
        M3 = __reduce_max_sync(act, v.w);
    if ((threadIdx.x & 31) == __ffs(act) - 1) {
        atomicMin(mn, m0); atomicMin(mn + 1, m1); atomicMin(mn + 2, m2); atomicMin(mn + 3, m3);
        atomicMax(mx, M0); atomicMax(mx + 1, M1); atomicMax(mx + 2, M2); atomicMax(mx + 3, M3);
    }
}

// K-V2  key = (b-bmin) | (x-xmin) | (y-ymin) | (z-zmin), each field as wide as its occupied range (the host sizes the radix
//       passes from the same extents); lexicographic (b,x,y,z)
__device__ __forceinline__ int vox_bits(int ext) {  // bits needed for values 0..ext, at least 1 (= pb::bit_width_i)
    return ext > 0 ? 32 - __clz(ext) : 1;
}
__global__ void k_vox_keys(const int4 *__restrict__ q, long long n, const int *__restrict__ mn, const int *__restrict__ mx,
                           uint64_t *__restrict__ key, uint32_t *__restrict__ val, int *err) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {
        for (int k = 0; k < 4; k++)
            if ((long long)mx[k] - mn[k] > 65535) atomicOr(err, 4);  // PB_ERR_RANGE
    }
    const int wx = vox_bits(mx[1] - mn[1]), wy = vox_bits(mx[2] - mn[2]), wz = vox_bits(mx[3] - mn[3]);
    int4 v = q[i];
    uint64_t b = (uint64_t)(unsigned)(v.x - mn[0]) & 0xffff, x = (uint64_t)(unsigned)(v.y - mn[1]) & 0xffff,
             y = (uint64_t)(unsigned)(v.z - mn[2]) & 0xffff, z = (uint64_t)(unsigned)(v.w - mn[3]) & 0xffff;
    key[i] = (b << (wx + wy + wz)) | (x << (wy + wz)) | (y << wz) | z;
    val[i] = (uint32_t)i;
}

__global__ void k_vox_heads(const uint64_t *__restrict__ skey, long long n, int *__restrict__ head) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
}

// K-V3  voxel table: coords, representative (smallest point index: the sort is stable), inverse map, CSR
__global__ void k_vox_table(const int4 *__restrict__ q, const uint32_t *__restrict__ order, const int *__restrict__ head,
                            const int *__restrict__ ex, long long n, int4 *__restrict__ vcoords, long long *__restrict__ index,
                            long long *__restrict__ inverse, int *__restrict__ order_out, int *__restrict__ vox_start,
                            int *__restrict__ d_V) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int h = head[i];
    int v = ex[i] + h - 1;
    uint32_t o = order[i];
    inverse[o] = v;
    order_out[i] = (int)o;
    if (h) {
        vcoords[v] = q[o];
        index[v] = o;
        vox_start[v] = (int)i;
    }
    if (i == n - 1) {
        vox_start[v + 1] = (int)n;
        *d_V = v + 1;
    }
}

// K-V4  per-voxel row reduction over the points of the voxel in ascending point order (deterministic):
//       MODE 0 = pick the representative, 1 = mean (UNWEIGHTED_AVERAGE), 2 = sum (devoxelize backward)
__device__ __forceinline__ float vsum_init(float) { return 0.f; }
__device__ __forceinline__ float4 vsum_init(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vadd(float &a, float b) { a += b; }
__device__ __forceinline__ void vadd(float4 &a, float4 b) { a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w; }
__device__ __forceinline__ float vdiv(float a, float d) { return a / d; }
__device__ __forceinline__ float4 vdiv(float4 a, float d) { return make_float4(a.x / d, a.y / d, a.z / d, a.w / d); }

template <int MODE, class V4>
__global__ void __launch_bounds__(256)
k_vox_rows(const V4 *__restrict__ rows, unsigned C4, const int *__restrict__ order, const int *__restrict__ vox_start,
           unsigned long long total, V4 *__restrict__ out) {
    // one output element per thread and trip (four per thread with batched loads was measured: 0.92 vs 0.73 ms for the
    // segmented sum of 8 M x 76 floats — the per-voxel loops of different lengths serialise)
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        unsigned v = (unsigned)(t / C4);
        unsigned c = (unsigned)(t - (unsigned long long)v * C4);
        int b = vox_start[v], e = vox_start[v + 1];
        if (MODE == 0) {
            out[t] = __ldg(rows + (long long)order[b] * C4 + c);
            continue;
        }
        V4 s = vsum_init(V4());
        for (int i = b; i < e; i++) vadd(s, __ldg(rows + (long long)order[i] * C4 + c));
        out[t] = MODE == 1 ? vdiv(s, (float)(e - b)) : s;
    }
}

// K-V5  devoxelize: out[p, :] = vfeat[inverse[p], :]   (network/PBNet.py:130-134,250).  Pure HBM traffic:
//       n*C*4 B written + the gathered rows read + 8 B/point of `inverse`.  IDX = uint32_t while n*C4 < 2^32
//       (a 64-bit divide per thread would make the kernel instruction-bound).
template <class V4, class IDX, int U>
__global__ void __launch_bounds__(256)
k_devox(const V4 *__restrict__ vfeat, IDX C4, const long long *__restrict__ inverse, IDX total, V4 *__restrict__ out) {
    // U elements per thread: the U index loads, then the U row loads, then the U stores (one element per thread is a chain
    // of two dependent loads with 16 B in flight per thread: latency-bound).  The host keeps total + U * blockDim.x
    // inside IDX.
    const IDX span = (IDX)blockDim.x * U;
    const unsigned long long stride = (unsigned long long)gridDim.x * span;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * span + threadIdx.x; t0 < total; t0 += stride) {
        long long src[U];
        IDX col[U];
#pragma unroll
        for (int u = 0; u < U; u++) {  // branch-free (clamped) so that the U index loads leave together
            IDX t = (IDX)t0 + (IDX)u * blockDim.x;
            IDX tc = t < total ? t : total - 1;
            IDX p = tc / C4;
            col[u] = tc - p * C4;
            src[u] = __ldg(inverse + p);
            if (t >= total) col[u] = ~(IDX)0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) src[u] = col[u] != ~(IDX)0 ? src[u] * (long long)C4 + (long long)col[u] : -1LL;
        V4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (src[u] >= 0) v[u] = __ldg(vfeat + src[u]);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (src[u] >= 0) __stcs(out + (IDX)t0 + (IDX)u * blockDim.x, v[u]);  // streaming store: the output is write-once
    }
}

}  // namespace pbv

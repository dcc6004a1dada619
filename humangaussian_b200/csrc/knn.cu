// knn.cu -- mean squared distance to the 3 nearest neighbours of every point: the `distCUDA2` of the reference's only
// vendored native component (gaussiansplatting/submodules/simple-knn/simple_knn.cu:63-221; caller
// scene/gaussian_model.py:134 and gs_renderer.py:386-389: initial Gaussian scales).  SURVEY.md 8f-4.
//
// Same exact-kNN idea (Morton order + bounding boxes to prune), laid out for this GPU: points are physically
// gathered into Morton order as float4 (the box scans then read contiguous 16-byte items that a whole warp shares,
// instead of chasing indices), boxes are 256 points, bounds come from one shuffle-reduced pass, and the sort is the
// repo's own stable radix sort (radix.cuh) on the 30-bit codes.  Exact: every box whose distance to the query is
// below the current third-best is scanned.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

#define KNN_BOX 256

__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u); // monotone float -> uint
}
__device__ __forceinline__ float ord2f(uint32_t u)
{
    const uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
    return __uint_as_float(b);
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float *__restrict__ pts, uint32_t *__restrict__ mm /* [6] min xyz, max xyz (ordered uints) */)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mm[k], f2ord(lo[k]));
            atomicMax(&mm[3 + k], f2ord(hi[k]));
        }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ mm,
                                                          uint32_t *__restrict__ codes, uint32_t *__restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = ord2f(mm[k]), hi = ord2f(mm[3 + k]);
        const float ext = hi - lo;
        const float t = ext > 0.f ? (pts[3 * (size_t)i + k] - lo) / ext : 0.f;
        c |= spread10((uint32_t)(t * 1023.0f)) << k;
    }
    codes[i] = c;
    idx[i] = (uint32_t)i;
}

// gather into Morton order (float4: xyz + original index bits) and compute box bounds
__global__ void __launch_bounds__(KNN_BOX) knn_gather_boxes_kernel(int P, const float *__restrict__ pts, const uint32_t *__restrict__ order,
                                                                    float4 *__restrict__ sorted, float *__restrict__ boxes /* [nb][6] */)
{
    __shared__ float s_lo[3][KNN_BOX / 32], s_hi[3][KNN_BOX / 32];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float p[3] = {0.f, 0.f, 0.f};
    const bool valid = i < P;
    if (valid) {
        const uint32_t src = order[i];
        p[0] = pts[3 * (size_t)src]; p[1] = pts[3 * (size_t)src + 1]; p[2] = pts[3 * (size_t)src + 2];
        sorted[i] = make_float4(p[0], p[1], p[2], __uint_as_float(src));
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float lo = valid ? p[k] : FLT_MAX, hi = valid ? p[k] : -FLT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { s_lo[k][threadIdx.x >> 5] = lo; s_hi[k][threadIdx.x >> 5] = hi; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float lo = FLT_MAX, hi = -FLT_MAX;
        for (int w = 0; w < KNN_BOX / 32; w++) { lo = fminf(lo, s_lo[threadIdx.x][w]); hi = fmaxf(hi, s_hi[threadIdx.x][w]); }
        boxes[6 * blockIdx.x + threadIdx.x] = lo;
        boxes[6 * blockIdx.x + 3 + threadIdx.x] = hi;
    }
}

__device__ __forceinline__ void k3_insert(float d, float *best)
{
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
}

__global__ void __launch_bounds__(KNN_BOX) knn_query_kernel(int P, int nb, const float4 *__restrict__ sorted, const float *__restrict__ boxes,
                                                             float *__restrict__ mean_d2)
{
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    if (i >= P) return;
    const float4 q = sorted[i];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int j = max(0, i - 3); j <= min(P - 1, i + 3); j++) { // Morton neighbours give a first bound
        if (j == i) continue;
        const float4 s = sorted[j];
        const float dx = s.x - q.x, dy = s.y - q.y, dz = s.z - q.z;
        k3_insert(dx * dx + dy * dy + dz * dz, best);
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    for (int b = 0; b < nb; b++) {
        const float *bx = boxes + 6 * b;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float c = k == 0 ? q.x : (k == 1 ? q.y : q.z);
            const float e = fmaxf(fmaxf(bx[k] - c, c - bx[3 + k]), 0.f);
            d += e * e;
        }
        if (d > reject || d > best[2]) continue;
        const int j1 = min(P, (b + 1) * KNN_BOX);
        for (int j = b * KNN_BOX; j < j1; j++) {
            if (j == i) continue;
            const float4 s = sorted[j];
            const float dx = s.x - q.x, dy = s.y - q.y, dz = s.z - q.z;
            k3_insert(dx * dx + dy * dy + dz * dz, best);
        }
    }
    mean_d2[__float_as_uint(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_scratch_bytes(int P)
{
    const size_t n = (size_t)(P > 0 ? P : 1), nb = (n + KNN_BOX - 1) / KNN_BOX;
    return 256 + gs_align(n * 4) * 4 + gs_align(n * 16) + gs_align(nb * 24) + test_sort32_scratch_bytes((int64_t)n);
}

int launch_knn3(int P, const float *pts, float *mean_d2, char *scratch, size_t scratch_bytes, cudaStream_t st, int *n_launches)
{
    if (scratch_bytes < knn_scratch_bytes(P)) return -1;
    const size_t n = (size_t)P, nb = (n + KNN_BOX - 1) / KNN_BOX;
    size_t o = 0;
    uint32_t *mm = (uint32_t *)(scratch + o); o += 256;
    uint32_t *codes_a = (uint32_t *)(scratch + o); o += gs_align(n * 4);
    uint32_t *codes_b = (uint32_t *)(scratch + o); o += gs_align(n * 4);
    uint32_t *idx_a = (uint32_t *)(scratch + o); o += gs_align(n * 4);
    uint32_t *idx_b = (uint32_t *)(scratch + o); o += gs_align(n * 4);
    float4 *sorted = (float4 *)(scratch + o); o += gs_align(n * 16);
    float *boxes = (float *)(scratch + o); o += gs_align(nb * 24);
    char *sort_scr = scratch + o;
    const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st);
    knn_bounds_kernel<<<(unsigned)min((size_t)592, (n + 255) / 256), 256, 0, st>>>(P, pts, mm);
    knn_morton_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, pts, mm, codes_a, idx_a);
    int in_b = 0;
    if (launch_test_sort32(codes_a, codes_b, idx_a, idx_b, (int64_t)n, 30, sort_scr, scratch_bytes - o, &in_b, st)) return -2;
    const uint32_t *order = in_b ? idx_b : idx_a;
    knn_gather_boxes_kernel<<<(unsigned)nb, KNN_BOX, 0, st>>>(P, pts, order, sorted, boxes);
    knn_query_kernel<<<(unsigned)nb, KNN_BOX, 0, st>>>(P, (int)nb, sorted, boxes, mean_d2);
    *n_launches += 4 + 12;
    return 0;
}

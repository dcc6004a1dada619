// binning.cu -- F2..F5: inclusive scan of tiles_touched, (tile|depth) key emission, stable sort,
// tile range identification.  Keys: ((view*tiles + tile) << 32) | depth bits,
// value = record index (view*P + gaussian).  A stable ascending sort on the low 32+bits(V*tiles) bits
// gives, inside each tile, ascending depth with ties in ascending Gaussian index.
//
// Round-1 note: scan and sort use CUB device primitives (library code, like cuBLAS for a GEMM); the
// emission and range kernels are ours.  DESIGN.md lists the hand-written replacement as the next step.
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

static int bits_for(uint64_t n)
{
    int b = 0;
    while (((uint64_t)1 << b) < n) b++;
    return b;
}

BinLayout binning_layout(int64_t capacity, int ntiles_total, int64_t n_vp)
{
    BinLayout L;
    size_t cap = (size_t)(capacity > 0 ? capacity : 1);
    size_t o = 0;
    L.keys_in = o; o += gs_align(cap * 8);
    L.keys_out = o; o += gs_align(cap * 8);
    L.vals_in = o; o += gs_align(cap * 4);
    L.vals_out = o; o += gs_align(cap * 4);
    L.ranges = o; o += gs_align((size_t)ntiles_total * 8);
    size_t t_sort = 0, t_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)cap, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, t_scan, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)(n_vp > 0 ? n_vp : 1));
    L.temp_bytes = gs_align(t_sort > t_scan ? t_sort : t_scan) + 256;
    L.temp = o; o += L.temp_bytes;
    L.total = o;
    return L;
}

int launch_scan_tiles(const uint32_t *tiles_touched, uint32_t *offsets, int64_t n, void *temp, size_t temp_bytes, cudaStream_t st)
{
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, tiles_touched, offsets, n, st);
    if (need > temp_bytes) return -1;
    cudaError_t e = cub::DeviceScan::InclusiveSum(temp, need, tiles_touched, offsets, n, st);
    return e == cudaSuccess ? 0 : -2;
}

__global__ void __launch_bounds__(256) emit_keys_kernel(const GeomRec *__restrict__ recs, const uint2 *__restrict__ rects,
                                                         const uint32_t *__restrict__ offsets, int P, int64_t n_vp,
                                                         int grid_x, int ntiles, uint64_t *__restrict__ keys,
                                                         uint32_t *__restrict__ vals)
{
    const int64_t vp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vp >= n_vp) return;
    const uint2 r = rects[vp];
    const int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
    if (x1 <= x0 || y1 <= y0) return;
    uint32_t off = (vp == 0) ? 0u : offsets[vp - 1];
    const uint32_t dbits = __float_as_uint(recs[vp].depth);
    const uint64_t tile_base = (uint64_t)(vp / P) * (uint64_t)ntiles;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            const uint64_t key = ((tile_base + (uint64_t)(y * grid_x + x)) << 32) | dbits;
            keys[off] = key;
            vals[off] = (uint32_t)vp;
            off++;
        }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint64_t *__restrict__ keys, int64_t D, uint2 *__restrict__ ranges)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const uint32_t t = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[t].x = 0;
    else {
        const uint32_t tp = (uint32_t)(keys[j - 1] >> 32);
        if (tp != t) {
            ranges[tp].y = (uint32_t)j;
            ranges[t].x = (uint32_t)j;
        }
    }
    if (j == D - 1) ranges[t].y = (uint32_t)D;
}

int launch_binning(const GeomRec *recs, const uint2 *rects, const uint32_t *offsets, int P, int V, int grid_x, int grid_y,
                   int64_t D, char *bin_base, const BinLayout &L, cudaStream_t st, int *n_launches)
{
    const int ntiles = grid_x * grid_y;
    const int64_t n_vp = (int64_t)P * V;
    uint64_t *keys_in = (uint64_t *)(bin_base + L.keys_in), *keys_out = (uint64_t *)(bin_base + L.keys_out);
    uint32_t *vals_in = (uint32_t *)(bin_base + L.vals_in), *vals_out = (uint32_t *)(bin_base + L.vals_out);
    uint2 *ranges = (uint2 *)(bin_base + L.ranges);
    cudaMemsetAsync(ranges, 0, (size_t)ntiles * V * 8, st);
    if (D == 0) return 0;
    emit_keys_kernel<<<(unsigned)((n_vp + 255) / 256), 256, 0, st>>>(recs, rects, offsets, P, n_vp, grid_x, ntiles, keys_in, vals_in);
    size_t need = L.temp_bytes;
    const int end_bit = 32 + bits_for((uint64_t)ntiles * V);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(bin_base + L.temp, need, keys_in, keys_out, vals_in, vals_out, D, 0, end_bit, st);
    if (e != cudaSuccess) return -2;
    tile_ranges_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(keys_out, D, ranges);
    *n_launches += 2;
    return 0;
}

// binning.cu -- F2..F5 re-designed around a DEPTH PRE-SORT.
//
// Upstream sorts D = sum(tiles_touched) 64-bit (tile|depth) keys: ~7 radix passes over every (Gaussian, tile)
// instance.  The order it defines -- inside each tile ascending depth, ties by ascending Gaussian index -- is
// reproduced here with far less traffic:
//   1. sort the V*P Gaussians once by (view, depth bits)      [stable; V*P is 4-5x smaller than D]
//   2. inclusive scan of tiles_touched IN THAT ORDER           -> D and the emission offsets
//   3. emit instances in depth order: key = view*tiles + tile (32 bit), value = record index
//   4. STABLE sort by the tile key only (12-18 bits: 2-3 passes over 8-byte pairs instead of 7 over 12-byte pairs)
//   5. tile ranges from the sorted tile keys
// Stability of step 4 keeps step 1's (depth, index) order inside every tile, so point_list and ranges are
// bit-identical to the reference's single 64-bit sort (tests compare them, and the reconstructed 64-bit keys,
// against the oracle).  Round-1 note: the radix passes and the scan are CUB device primitives (library code);
// key construction, emission and range identification are ours.
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

static int bits_for(uint64_t n)
{
    int b = 0;
    while (((uint64_t)1 << b) < n) b++;
    return b;
}

struct TilesInOrder {
    const uint32_t *tiles;
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t &vp) const { return tiles[vp]; }
};
typedef cub::TransformInputIterator<uint32_t, TilesInOrder, const uint32_t *> TilesIter;

BinLayout binning_layout(int64_t capacity, int ntiles_total, int64_t n_vp)
{
    BinLayout L;
    const size_t cap = (size_t)(capacity > 0 ? capacity : 1), nvp = (size_t)(n_vp > 0 ? n_vp : 1);
    size_t o = 0;
    L.keys_in = o; o += gs_align(cap * 4);
    L.keys_out = o; o += gs_align(cap * 4);
    L.vals_in = o; o += gs_align(cap * 4);
    L.vals_out = o; o += gs_align(cap * 4);
    L.ranges = o; o += gs_align((size_t)ntiles_total * 8);
    L.dkeys_in = o; o += gs_align(nvp * 8);
    L.dkeys_out = o; o += gs_align(nvp * 8);
    L.order_in = o; o += gs_align(nvp * 4);
    L.order = o; o += gs_align(nvp * 4);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int64_t)cap, 0, 32);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int64_t)nvp, 0, 64);
    TilesIter it((const uint32_t *)nullptr, TilesInOrder{nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, t3, it, (uint32_t *)nullptr, (int64_t)nvp);
    size_t t = t1 > t2 ? t1 : t2;
    t = t > t3 ? t : t3;
    L.temp_bytes = gs_align(t) + 256;
    L.temp = o; o += L.temp_bytes;
    L.total = o;
    return L;
}

// depth pre-sort + scan in depth order.  dkeys_in / order_in were written by the preprocess kernel.
int launch_depth_order(const uint32_t *tiles_touched, uint32_t *offsets_sorted, int64_t n_vp, int V, char *bin_base,
                       const BinLayout &L, cudaStream_t st)
{
    size_t need = L.temp_bytes;
    const int end_bit = 32 + bits_for((uint64_t)V);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(bin_base + L.temp, need, (const uint64_t *)(bin_base + L.dkeys_in),
                                                    (uint64_t *)(bin_base + L.dkeys_out), (const uint32_t *)(bin_base + L.order_in),
                                                    (uint32_t *)(bin_base + L.order), n_vp, 0, end_bit, st);
    if (e != cudaSuccess) return -1;
    TilesIter it((const uint32_t *)(bin_base + L.order), TilesInOrder{tiles_touched});
    need = L.temp_bytes;
    e = cub::DeviceScan::InclusiveSum(bin_base + L.temp, need, it, offsets_sorted, n_vp, st);
    return e == cudaSuccess ? 0 : -2;
}

// one thread per Gaussian IN DEPTH ORDER: writes its tile keys at the offsets the scan assigned
__global__ void __launch_bounds__(256) emit_tiles_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects,
                                                          const uint32_t *__restrict__ offsets_sorted, int P, int64_t n_vp,
                                                          int grid_x, int ntiles, uint32_t *__restrict__ keys,
                                                          uint32_t *__restrict__ vals)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_vp) return;
    const uint32_t vp = order[k];
    const uint2 r = rects[vp];
    const int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
    if (x1 <= x0 || y1 <= y0) return;
    uint32_t off = (k == 0) ? 0u : offsets_sorted[k - 1];
    const uint32_t tile_base = (uint32_t)(vp / (uint32_t)P) * (uint32_t)ntiles;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            keys[off] = tile_base + (uint32_t)(y * grid_x + x);
            vals[off] = vp;
            off++;
        }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t *__restrict__ keys, int64_t D, uint2 *__restrict__ ranges)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const uint32_t t = keys[j];
    if (j == 0) ranges[t].x = 0;
    else {
        const uint32_t tp = keys[j - 1];
        if (tp != t) {
            ranges[tp].y = (uint32_t)j;
            ranges[t].x = (uint32_t)j;
        }
    }
    if (j == D - 1) ranges[t].y = (uint32_t)D;
}

int launch_binning(const uint2 *rects, const uint32_t *offsets_sorted, int P, int V, int grid_x, int grid_y, int64_t D,
                   char *bin_base, const BinLayout &L, cudaStream_t st, int *n_launches)
{
    const int ntiles = grid_x * grid_y;
    const int64_t n_vp = (int64_t)P * V;
    uint32_t *keys_in = (uint32_t *)(bin_base + L.keys_in), *keys_out = (uint32_t *)(bin_base + L.keys_out);
    uint32_t *vals_in = (uint32_t *)(bin_base + L.vals_in), *vals_out = (uint32_t *)(bin_base + L.vals_out);
    uint2 *ranges = (uint2 *)(bin_base + L.ranges);
    cudaMemsetAsync(ranges, 0, (size_t)ntiles * V * 8, st);
    if (D == 0) return 0;
    emit_tiles_kernel<<<(unsigned)((n_vp + 255) / 256), 256, 0, st>>>((const uint32_t *)(bin_base + L.order), rects, offsets_sorted, P,
                                                                      n_vp, grid_x, ntiles, keys_in, vals_in);
    size_t need = L.temp_bytes;
    const int end_bit = bits_for((uint64_t)ntiles * V);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(bin_base + L.temp, need, keys_in, keys_out, vals_in, vals_out, D, 0,
                                                    end_bit > 0 ? end_bit : 1, st);
    if (e != cudaSuccess) return -2;
    tile_ranges_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(keys_out, D, ranges);
    *n_launches += 2;
    return 0;
}

// radix.cuh -- hand-written stable LSD radix sort of (key, u32 value) pairs and a decoupled-look-back scan,
// the two primitives of the binning stage.  sm_100a, 256-thread CTAs.
//
// One pass = ONE kernel ("onesweep"): every CTA takes a ticket (so CTAs are processed in launch-independent
// order), ranks its tile of 256*IPT items by digit with warp match.any + per-warp shared counters (stable:
// items keep their input order inside a digit), publishes its per-digit counts, resolves the exclusive prefix
// over all earlier CTAs by decoupled look-back, reorders the tile through shared memory so that the final
// global stores are contiguous runs per digit, and scatters.  The global per-digit bases of ALL passes come
// from one up-front histogram kernel.  Digit width is a template parameter (6..9 bits): the tile-key sort of a
// 64-view batch (18 bits) takes 2 passes of 9 bits where an 8-bit library sort takes 3.
//
// Memory per item and pass: one read + one write of (key, value) -- the single-pass lower bound used as the
// "algorithmic bytes" of the sort in bench.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsradix {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr uint32_t FLAG_AGG = 1u << 30, FLAG_PREFIX = 2u << 30, FLAG_MASK = 3u << 30, VALUE_MASK = ~FLAG_MASK;

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, int bits)
{
    return (uint32_t)(k >> shift) & ((1u << bits) - 1u);
}

// ---- histogram of up to MAXP digit positions in one read of the keys -----------------------------------------
template <typename KeyT, int RADIX_BITS, int MAXP>
__global__ void __launch_bounds__(THREADS) histogram_kernel(const KeyT *__restrict__ keys, int64_t n, int npass, int first_shift,
                                                             int last_bits, uint32_t *__restrict__ hist /* [npass][BINS] */)
{
    constexpr int BINS = 1 << RADIX_BITS;
    __shared__ uint32_t sh[MAXP * BINS];
    for (int i = threadIdx.x; i < npass * BINS; i += THREADS) sh[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        const KeyT k = keys[i];
#pragma unroll
        for (int p = 0; p < MAXP; p++)
            if (p < npass) {
                const int bits = (p == npass - 1) ? last_bits : RADIX_BITS;
                atomicAdd(&sh[p * BINS + digit_of(k, first_shift + p * RADIX_BITS, bits)], 1u);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * BINS; i += THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of each pass's histogram, in place (one CTA per pass)
template <int RADIX_BITS>
__global__ void __launch_bounds__(THREADS) scan_hist_kernel(uint32_t *hist)
{
    constexpr int BINS = 1 << RADIX_BITS;
    __shared__ uint32_t s[BINS];
    uint32_t *h = hist + (size_t)blockIdx.x * BINS;
    for (int i = threadIdx.x; i < BINS; i += THREADS) s[i] = h[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < BINS; i++) {
            const uint32_t c = s[i];
            s[i] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += THREADS) h[i] = s[i];
}

// ---- one onesweep pass ------------------------------------------------------------------------------------------
template <typename KeyT, int RADIX_BITS, int IPT>
__global__ void __launch_bounds__(THREADS) onesweep_kernel(const KeyT *__restrict__ keys_in, KeyT *__restrict__ keys_out,
                                                            const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ vals_out,
                                                            int64_t n, int shift, int bits,
                                                            const uint32_t *__restrict__ digit_base /* [BINS] exclusive */,
                                                            volatile uint32_t *status /* [nblocks][BINS], zeroed */,
                                                            uint32_t *ticket /* zeroed */)
{
    constexpr int BINS = 1 << RADIX_BITS;
    constexpr int TILE = THREADS * IPT;
    constexpr int DPT = (BINS + THREADS - 1) / THREADS; // digits per thread in the per-digit phases
    __shared__ uint32_t s_cnt[WARPS][BINS]; // per-warp digit counters -> exclusive warp prefixes
    __shared__ uint32_t s_loc[BINS];        // exclusive scan of the CTA's digit counts (local sorted offsets)
    __shared__ uint32_t s_glob[BINS];       // global destination of the CTA's first item of each digit
    __shared__ KeyT s_keys[TILE];
    __shared__ uint32_t s_vals[TILE];
    __shared__ uint32_t s_bid, s_wsum[WARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * BINS; i += THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t bid = s_bid;
    const int64_t tile_start = (int64_t)bid * TILE;
    const int n_tile = (int)min((int64_t)TILE, n - tile_start);

    // ---- load (warp-striped: coalesced) and rank
    KeyT key[IPT];
    uint32_t val[IPT], rnk[IPT];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const int idx = warp * (IPT * 32) + i * 32 + lane;
        const bool valid = idx < n_tile;
        key[i] = valid ? keys_in[tile_start + idx] : (KeyT)0;
        val[i] = valid ? vals_in[tile_start + idx] : 0u;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        rnk[i] = 0xffffffffu;
        if (valid) {
            const uint32_t d = digit_of(key[i], shift, bits);
            const uint32_t peers = __match_any_sync(vmask, d);
            const uint32_t before = s_cnt[warp][d];
            __syncwarp(vmask);
            if ((peers & lt) == 0) s_cnt[warp][d] = before + __popc(peers);
            __syncwarp(vmask);
            rnk[i] = before + __popc(peers & lt);
        }
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over warps, CTA count
    uint32_t cta_count[DPT];
#pragma unroll
    for (int q = 0; q < DPT; q++) {
        const int d = tid + q * THREADS;
        cta_count[q] = 0;
        if (d < BINS) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const uint32_t c = s_cnt[w][d];
                s_cnt[w][d] = run;
                run += c;
            }
            cta_count[q] = run;
        }
    }
    // ---- publish the CTA aggregate, then look back for the exclusive prefix over earlier CTAs
#pragma unroll
    for (int q = 0; q < DPT; q++) {
        const int d = tid + q * THREADS;
        if (d < BINS) {
            status[(size_t)bid * BINS + d] = (bid == 0 ? FLAG_PREFIX : FLAG_AGG) | cta_count[q];
        }
    }
    __threadfence();
    // ---- local exclusive scan of cta_count over digits (digits are spread tid + q*THREADS)
    {
        // scan within each q-slab with warp shuffles, slabs are consecutive digit ranges
        uint32_t slab_base = 0;
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            uint32_t x = cta_count[q];
            uint32_t incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0;
            for (int w = 0; w < warp; w++) wbase += s_wsum[w];
            uint32_t slab_total = 0;
            for (int w = 0; w < WARPS; w++) slab_total += s_wsum[w];
            const int d = tid + q * THREADS;
            if (d < BINS) s_loc[d] = slab_base + wbase + incl - x;
            slab_base += slab_total;
            __syncthreads();
        }
    }
#pragma unroll
    for (int q = 0; q < DPT; q++) {
        const int d = tid + q * THREADS;
        if (d < BINS) {
            uint32_t excl = 0;
            if (bid > 0) {
                int64_t p = (int64_t)bid - 1;
                while (true) {
                    uint32_t s;
                    do { s = status[(size_t)p * BINS + d]; } while ((s & FLAG_MASK) == 0);
                    excl += s & VALUE_MASK;
                    if ((s & FLAG_MASK) == FLAG_PREFIX) break;
                    p--;
                }
                status[(size_t)bid * BINS + d] = FLAG_PREFIX | (excl + cta_count[q]);
            }
            s_glob[d] = digit_base[d] + excl;
        }
    }
    __syncthreads();

    // ---- reorder through shared memory (local stable sort by digit), then scatter contiguous runs
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        if (rnk[i] != 0xffffffffu) {
            const uint32_t d = digit_of(key[i], shift, bits);
            const uint32_t lpos = s_loc[d] + s_cnt[warp][d] + rnk[i];
            s_keys[lpos] = key[i];
            s_vals[lpos] = val[i];
        }
    }
    __syncthreads();
    for (int k = tid; k < n_tile; k += THREADS) {
        const KeyT kk = s_keys[k];
        const uint32_t d = digit_of(kk, shift, bits);
        const int64_t dst = (int64_t)s_glob[d] + (k - s_loc[d]);
        keys_out[dst] = kk;
        vals_out[dst] = s_vals[k];
    }
}

// ---- inclusive scan of x[k] = tiles[order[k]] with decoupled look-back (single pass over the data) ------------
constexpr int SCAN_IPT = 8;
__global__ void __launch_bounds__(THREADS) scan_tiles_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
                                                              uint32_t *__restrict__ out, int64_t n, volatile uint64_t *status /* [nblocks], zeroed */,
                                                              uint32_t *ticket /* zeroed */)
{
    // status: bit 63 = prefix ready, bit 62 = aggregate ready, low 40 bits = value
    constexpr int TILE = THREADS * SCAN_IPT;
    __shared__ uint32_t s_bid, s_wsum[WARPS];
    __shared__ uint64_t s_excl;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const int64_t base = (int64_t)bid * TILE + (int64_t)tid * SCAN_IPT;
    uint32_t x[SCAN_IPT];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; i++) {
        x[i] = (base + i < n) ? tiles[order[base + i]] : 0u;
        tsum += x[i];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int w = 0; w < WARPS; w++) {
        if (w < warp) wbase += s_wsum[w];
        total += s_wsum[w];
    }
    if (tid == 0) {
        const uint64_t AGG = 1ull << 62, PRE = 1ull << 63, VM = (1ull << 40) - 1;
        status[bid] = (bid == 0 ? PRE : AGG) | (uint64_t)total;
        __threadfence();
        uint64_t excl = 0;
        if (bid > 0) {
            int64_t p = (int64_t)bid - 1;
            while (true) {
                uint64_t s;
                do { s = status[p]; } while ((s >> 62) == 0);
                excl += s & VM;
                if (s & PRE) break;
                p--;
            }
            status[bid] = PRE | (excl + total);
        }
        s_excl = excl;
    }
    __syncthreads();
    uint32_t run = (uint32_t)s_excl + wbase + incl - tsum;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; i++) {
        run += x[i];
        if (base + i < n) out[base + i] = run;
    }
}

} // namespace gsradix

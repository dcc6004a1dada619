// blend.cu -- F6 forward alpha blend and B1 backward blend.
//
// PATCH-PARALLEL: every warp owns one 8x4 pixel patch of a 16x16 tile and walks the tile's instance list ON ITS
// OWN -- no CTA-wide barrier anywhere in the loops.  (The first version staged 256 instances per CTA between
// __syncthreads; ncu showed 35 % of warp-time parked on that barrier because the eight patches of a tile see very
// different numbers of Gaussians.)  CTAs are 64 threads = two independent warps (a 16x4 strip), so the block
// scheduler balances patches dynamically and a finished warp frees its slot almost immediately.
//   * Chunks of 32 list entries: lane l loads entry l's id and the first 16 bytes of its 48-byte GeomRec
//     (px, py, hx, hy), tests the Gaussian's conservative alpha-support box against the warp's patch, and the
//     warp ballots.  Hits load the rest of the record and are compacted, in order, into a per-warp shared-memory
//     slab; the next chunk's id/box loads are already in flight while the hits are blended.
//     Measured on the bench workload (300 k subsample of sample.ply, counter build of round 2): 2.15 of 8 patches
//     survive the box test per instance, 83 % of those visits have a contributing lane, 13.6 of 32 lanes contribute.
//     The hits' list position and record index ride in the two slab words the support box occupied (SLAB_PACK).
//   * Tile lists are exactly the reference's (tile/sort indices bit-identical); culling only skips pairs whose
//     alpha is provably < 1/255, so images are unchanged.  Forward arithmetic follows the pinned order of
//     common.cuh: images are bit-identical to the CPU oracle.
//   * Backward: per-pixel terms are expressed as ten sums per Gaussian -- six moments of q = G*dL/dalpha
//     (1, dx, dy, dx^2, dx*dy, dy^2) and four colour/depth weights -- reduced over the
//     32 lanes with a 12-shuffle multi-value butterfly (not 10 x 5 shuffles), then ten lanes each issue one
//     red.global.add.f32 into the Gaussian's 48-byte ScreenGrad record.  Upstream: ~10 atomics per PIXEL.
//   * The compile-time switches below are the measured experiments of round 2 (DESIGN.md section 5 has the numbers): the
//     product is SLAB_PACK=1, BWD_T_DIV=1, everything else 0.
#include "common.cuh"
#include "kernels.h"

#define FULL 0xffffffffu
#ifndef WARPS_PER_CTA
#define WARPS_PER_CTA 2 // independent patches per CTA (1, 2, 4 or 8); swept on B200: 2 is best by ~2 %
#endif
#ifndef BWD_MIN_BLOCKS
#define BWD_MIN_BLOCKS 16 // caps the backward kernel at 64 registers (32 warps/SM); 18 / 20 blocks (55 / 48 regs) measured slower
#endif
#define CTAS_PER_TILE (8 / WARPS_PER_CTA)
#ifndef BWD_ILP2
#define BWD_ILP2 0 // backward: two hits per inner iteration (needs SLAB_PACK)
#endif
#ifndef FWD_ILP2
#define FWD_ILP2 0 // forward: two hits per inner iteration (needs SLAB_PACK)
#endif
#ifndef FWD_BULK
#define FWD_BULK 0 // 1: forward blend with cp.async.bulk + mbarrier staging of the id runs (A/B experiment)
#endif
#ifndef SGRAD_F64
#define SGRAD_F64 0 // 1: per-Gaussian accumulators in double (red.global.add.f64).  Measured at BASELINE sizes: the worst gradient
                    // rows do not move (they are not accumulation-order errors), backward +1.2 %, scratch x2: off.
#endif
#ifndef BWD_T_DIV
#define BWD_T_DIV 1 // transmittance recovery T <- T/(1-alpha): 0 = MUFU reciprocal, 1 = + one Newton step, 2 = IEEE division.
                    // Measured on the 17 k-deep tiles of the real scene: the raw reciprocal's bias compounds (worst position-gradient
                    // row 1.9x of tolerance), one Newton step brings it to 1.0-1.1x for +0.9 % backward time; the division costs +7 %.
#endif

// does the support box [px-hx,px+hx] x [py-hy,py+hy] reach the patch [x0,x0+7] x [y0,y0+3] ?
__device__ __forceinline__ bool box_hits_patch(const float4 g0, float x0, float y0)
{
    return (g0.z >= 0.0f) && (g0.x + g0.z >= x0) && (g0.x - g0.z <= x0 + 7.0f) && (g0.y + g0.w >= y0) && (g0.y - g0.w <= y0 + 3.0f);
}

// power with the conic pre-scaled at staging time (A' = -A/2, B' = -B, C' = -C/2: exact operations, so the
// result is bit-identical to gs_power(A,B,C,dx,dy))
__device__ __forceinline__ float power_prescaled(float Ap, float Bp, float Cp, float dx, float dy)
{
    const float u = ffma(Bp, dy, fmul(Ap, dx));
    const float w = fmul(fmul(Cp, dy), dy);
    return ffma(dx, u, w);
}

// per-warp slab of the (at most 32) hits of the current chunk
#ifndef SLAB_PACK
#define SLAB_PACK 1 // 1: list position and record index ride in the two words the support box occupied (no separate pos/id arrays)
#endif
struct __align__(16) WarpSlab {
#if SLAB_PACK
    float4 rec[32 * 3]; // px,py,list position (bits),record index (bits) | A',B',C',o | r,g,b,depth
#else
    float4 rec[32 * 3]; // px,py,hx,hy | A',B',C',o | r,g,b,depth
    uint32_t pos[32];   // 1-based position in the tile list (forward) / 0-based (backward)
    uint32_t id[32];    // record index (backward: address of the ScreenGrad accumulator)
#endif
};

// ------------------------------------------------------------------------------------------------
// F6
// ------------------------------------------------------------------------------------------------
#if !FWD_BULK
__global__ void __launch_bounds__(32 * WARPS_PER_CTA) blend_fwd_kernel(BlendArgs a)
{
    __shared__ WarpSlab slabs[WARPS_PER_CTA];

    const int ntiles = a.grid_x * a.grid_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTAs are dispatched in blockIdx order; tile_order lists (view, tile) pairs longest list first, so the heavy
    // tiles of a launch start first and the short ones fill in behind them (LPT schedule)
    const uint32_t vt = a.tile_order[blockIdx.x / CTAS_PER_TILE];
    const int v = (int)(vt / (uint32_t)ntiles), tile = (int)(vt % (uint32_t)ntiles);
    const int patch = (blockIdx.x % CTAS_PER_TILE) * WARPS_PER_CTA + warp; // 0..7: bit0 = x half, bits 1-2 = row band
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int x0 = tile_x * GS_TILE + (patch & 1) * 8, y0 = tile_y * GS_TILE + (patch >> 1) * 4;
    const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float fpx = (float)px, fpy = (float)py, fx0 = (float)x0, fy0 = (float)y0;
    WarpSlab &sl = slabs[warp];

    const uint2 range = a.ranges[(size_t)v * ntiles + tile];
    const int n_total = (int)(range.y - range.x);
    const float4 *recs4 = reinterpret_cast<const float4 *>(a.recs);
    const uint32_t *plist = a.point_list + range.x;
    const uint32_t lt = (1u << lane) - 1u;

    // T == 0 is the "done" sentinel: a finished (or out-of-image) pixel keeps failing the T test and never
    // accumulates, exactly like upstream's `done` flag; T_out remembers the transmittance to report.
    float T = inside ? 1.0f : 0.0f, T_out = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dd = 0.f, Aa = 0.f;
    uint32_t last = 0;

    // software pipeline, two deep: while chunk c is blended the support boxes of chunk c+1 and the ids of chunk c+2
    // are in flight (the id -> record address dependency never sits on the critical path)
    uint32_t id_c = 0, id_n = 0;
    float4 g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
    if (lane < n_total) id_c = __ldg(plist + lane);
    if (32 + lane < n_total) id_n = __ldg(plist + 32 + lane);
    if (lane < n_total) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
    for (int base = 0; base < n_total; base += 32) {
        if (__all_sync(FULL, T == 0.0f)) break;
        const bool hit = box_hits_patch(g0_c, fx0, fy0);
        const uint32_t b = __ballot_sync(FULL, hit);
        float4 g1, g2;
        if (hit) {
            g1 = __ldg(recs4 + 3 * (size_t)id_c + 1);
            g2 = __ldg(recs4 + 3 * (size_t)id_c + 2);
        }
        const float4 g0_h = g0_c;
        id_c = id_n;
        g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
        if (base + 32 + lane < n_total) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
        if (base + 64 + lane < n_total) id_n = __ldg(plist + base + 64 + lane);
        if (b == 0u) continue;
        if (hit) {
            const int slot = __popc(b & lt);
#if SLAB_PACK
            sl.rec[3 * slot] = make_float4(g0_h.x, g0_h.y, __uint_as_float((uint32_t)(base + lane + 1)), 0.f);
#else
            sl.rec[3 * slot] = g0_h;
            sl.pos[slot] = (uint32_t)(base + lane + 1);
#endif
            sl.rec[3 * slot + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * slot + 2] = g2;
        }
        __syncwarp();
        const int cnt = __popc(b);
#if FWD_ILP2
        // two hits per iteration: their powers / exps / alphas are independent instruction streams (the transmittance chain
        // is only the few operations behind them), which is what a single warp walking a deep tile needs -- one view has a
        // few long lists, not enough warps to hide a dependent chain.  An odd count is padded with a zero-opacity entry.
        if ((cnt & 1) && lane == 0) { // finite centre + zero conic/opacity: power = 0, alpha = 0 -> skipped by the 1/255 test
            sl.rec[3 * cnt] = make_float4(0.f, 0.f, 0.f, 0.f);
            sl.rec[3 * cnt + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        for (int i = 0; i < cnt; i += 2) {
            const float4 g0a = sl.rec[3 * i], q1a = sl.rec[3 * i + 1];
            const float4 g0b = sl.rec[3 * i + 3], q1b = sl.rec[3 * i + 4];
            const float dxa = fsub(g0a.x, fpx), dya = fsub(g0a.y, fpy), dxb = fsub(g0b.x, fpx), dyb = fsub(g0b.y, fpy);
            const float pa = power_prescaled(q1a.x, q1a.y, q1a.z, dxa, dya), pb = power_prescaled(q1b.x, q1b.y, q1b.z, dxb, dyb);
            const float aa = fminf(GS_ALPHA_MAX, fmul(q1a.w, gs_exp(pa))), ab = fminf(GS_ALPHA_MAX, fmul(q1b.w, gs_exp(pb)));
            if (!(pa > 0.0f) && !(aa < GS_ALPHA_MIN)) {
                const float test_T = fmul(T, fsub(1.0f, aa));
                if (test_T < GS_T_MIN) {
                    if (T != 0.0f) T_out = T;
                    T = 0.0f;
                } else {
                    const float w = fmul(aa, T);
                    const float4 q2 = sl.rec[3 * i + 2];
                    C0 = ffma(q2.x, w, C0); C1 = ffma(q2.y, w, C1); C2 = ffma(q2.z, w, C2);
                    Dd = ffma(q2.w, w, Dd);
                    Aa = fadd(Aa, w);
                    T = test_T;
                    last = __float_as_uint(g0a.z);
                }
            }
            if (!(pb > 0.0f) && !(ab < GS_ALPHA_MIN)) {
                const float test_T = fmul(T, fsub(1.0f, ab));
                if (test_T < GS_T_MIN) {
                    if (T != 0.0f) T_out = T;
                    T = 0.0f;
                } else {
                    const float w = fmul(ab, T);
                    const float4 q2 = sl.rec[3 * i + 5];
                    C0 = ffma(q2.x, w, C0); C1 = ffma(q2.y, w, C1); C2 = ffma(q2.z, w, C2);
                    Dd = ffma(q2.w, w, Dd);
                    Aa = fadd(Aa, w);
                    T = test_T;
                    last = __float_as_uint(g0b.z);
                }
            }
        }
#else
        for (int i = 0; i < cnt; i++) {
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = power_prescaled(q1.x, q1.y, q1.z, dx, dy);
            if (!(power > 0.0f)) {
                const float alpha = fminf(GS_ALPHA_MAX, fmul(q1.w, gs_exp(power)));
                if (!(alpha < GS_ALPHA_MIN)) {
                    const float test_T = fmul(T, fsub(1.0f, alpha));
                    if (test_T < GS_T_MIN) {
                        if (T != 0.0f) T_out = T;
                        T = 0.0f;
                    } else {
                        const float w = fmul(alpha, T);
                        const float4 q2 = sl.rec[3 * i + 2];
                        C0 = ffma(q2.x, w, C0); C1 = ffma(q2.y, w, C1); C2 = ffma(q2.z, w, C2);
                        Dd = ffma(q2.w, w, Dd);
                        Aa = fadd(Aa, w);
                        T = test_T;
#if SLAB_PACK
                        last = __float_as_uint(g0.z);
#else
                        last = sl.pos[i];
#endif
                    }
                }
            }
        }
#endif
        __syncwarp();
    }
    if (inside) {
        if (T != 0.0f) T_out = T;
        const size_t HW = (size_t)a.H * a.W;
        const size_t pix = (size_t)py * a.W + px;
        a.final_T[v * HW + pix] = T_out;
        a.n_contrib[v * HW + pix] = last;
        float *oc = a.out_color + (size_t)v * 3 * HW;
        oc[pix] = ffma(T_out, a.bg[0], C0);
        oc[HW + pix] = ffma(T_out, a.bg[1], C1);
        oc[2 * HW + pix] = ffma(T_out, a.bg[2], C2);
        a.out_depth[v * HW + pix] = Dd;
        a.out_alpha[v * HW + pix] = Aa;
    }
}

#endif // !FWD_BULK

#if FWD_BULK
// ------------------------------------------------------------------------------------------------
// F6, bulk-copy variant (A/B experiment of round 2; DESIGN.md "Blackwell staging"): the tile's id run -- the one contiguous
// stream of the blend -- is brought into shared memory by the async proxy (cp.async.bulk + mbarrier complete_tx, two 1 KB
// stages per warp) instead of per-lane LDG prefetches; the 48-byte records are still gathered by id (they are scattered in
// HBM: a bulk copy moves contiguous bytes only).  Chunks are aligned to the absolute 32-entry grid of point_list so that a
// chunk never straddles two stages and every bulk copy starts on a 16-byte boundary; results are identical.
// ------------------------------------------------------------------------------------------------
#define BULK_CHUNKS 8                    // chunks (of 32 ids) per stage
#define BULK_IDS (32 * BULK_CHUNKS)      // 256 ids = 1 KB per stage
struct __align__(16) BulkStage {
    uint32_t ids[2][BULK_IDS];
    unsigned long long mbar[2];
};
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *b)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity)
{
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(b)), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 24)) __trap(); // a protocol error must fail loudly, never hang the GPU
    }
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA) blend_fwd_kernel(BlendArgs a)
{
    __shared__ WarpSlab slabs[WARPS_PER_CTA];
    __shared__ BulkStage bulk[WARPS_PER_CTA];

    const int ntiles = a.grid_x * a.grid_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t vt = a.tile_order[blockIdx.x / CTAS_PER_TILE];
    const int v = (int)(vt / (uint32_t)ntiles), tile = (int)(vt % (uint32_t)ntiles);
    const int patch = (blockIdx.x % CTAS_PER_TILE) * WARPS_PER_CTA + warp;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int x0 = tile_x * GS_TILE + (patch & 1) * 8, y0 = tile_y * GS_TILE + (patch >> 1) * 4;
    const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float fpx = (float)px, fpy = (float)py, fx0 = (float)x0, fy0 = (float)y0;
    WarpSlab &sl = slabs[warp];
    BulkStage &bk = bulk[warp];

    const uint2 range = a.ranges[(size_t)v * ntiles + tile];
    const float4 *recs4 = reinterpret_cast<const float4 *>(a.recs);
    const uint32_t lt = (1u << lane) - 1u;
    // absolute chunk grid: chunk c covers point_list[a0 + 32 c + lane]
    const uint32_t a0 = range.x & ~31u;
    const int n_chunks = range.y > range.x ? (int)((range.y - a0 + 31u) >> 5) : 0;
    const int n_stages = (n_chunks + BULK_CHUNKS - 1) / BULK_CHUNKS;

    float T = inside ? 1.0f : 0.0f, T_out = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dd = 0.f, Aa = 0.f;
    uint32_t last = 0;

    if (lane == 0) { mbar_init(&bk.mbar[0], 1); mbar_init(&bk.mbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // stage s -> buffer s & 1, its (s >> 1)-th use: wait parity (s >> 1) & 1.  Bytes: whole stage, clipped to the 16-byte
    // rounded end of the tile's run (inside the point_list allocation: its capacity is rounded up to 256 bytes).
    auto issue = [&](int s) {
        const uint32_t first = a0 + (uint32_t)s * BULK_IDS;
        const uint32_t end = min(first + (uint32_t)BULK_IDS, (range.y + 3u) & ~3u);
        bulk_load(bk.ids[s & 1], a.point_list + first, (end - first) * 4u, &bk.mbar[s & 1]);
    };
    int issued = 0, waited = 0; // stages issued / stages whose arrival has been observed (warp-uniform)
    if (lane == 0) { if (n_stages > 0) issue(0); if (n_stages > 1) issue(1); }
    issued = min(n_stages, 2);

    uint32_t id_c = 0;
    float4 g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
    bool valid_c = false;
    if (n_chunks > 0) {
        mbar_wait(&bk.mbar[0], 0);
        waited = 1;
        const uint32_t ab = a0 + lane;
        valid_c = ab >= range.x && ab < range.y;
        if (valid_c) { id_c = bk.ids[0][lane]; g0_c = __ldg(recs4 + 3 * (size_t)id_c); }
    }
    for (int c = 0; c < n_chunks; c++) {
        if (__all_sync(FULL, T == 0.0f)) break;
        const bool hit = box_hits_patch(g0_c, fx0, fy0);
        const uint32_t b = __ballot_sync(FULL, hit);
        float4 g1, g2;
        if (hit) {
            g1 = __ldg(recs4 + 3 * (size_t)id_c + 1);
            g2 = __ldg(recs4 + 3 * (size_t)id_c + 2);
        }
        const float4 g0_h = g0_c;
        const uint32_t pos_h = a0 + 32u * (uint32_t)c + lane - range.x + 1u; // 1-based list position
        // next chunk: ids from the stage in shared memory (wait for it when the chunk opens a new stage), records by gather
        g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
        if (c + 1 < n_chunks) {
            const int sn = (c + 1) / BULK_CHUNKS;
            if (sn == waited) { mbar_wait(&bk.mbar[sn & 1], (uint32_t)(sn >> 1) & 1u); waited = sn + 1; }
            const uint32_t ab = a0 + 32u * (uint32_t)(c + 1) + lane;
            if (ab >= range.x && ab < range.y) {
                id_c = bk.ids[sn & 1][((c + 1) % BULK_CHUNKS) * 32 + lane];
                g0_c = __ldg(recs4 + 3 * (size_t)id_c);
            }
            // the stage before `sn` was last read one iteration ago (the ballot above ordered those reads): refill its buffer
            if ((c + 1) % BULK_CHUNKS == 0 && sn + 1 < n_stages && sn + 1 == issued) {
                if (lane == 0) issue(sn + 1);
                issued = sn + 2;
            }
        }
        if (b == 0u) continue;
        if (hit) {
            const int slot = __popc(b & lt);
            sl.rec[3 * slot] = make_float4(g0_h.x, g0_h.y, __uint_as_float(pos_h), 0.f);
            sl.rec[3 * slot + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * slot + 2] = g2;
        }
        __syncwarp();
        const int cnt = __popc(b);
        for (int i = 0; i < cnt; i++) {
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = power_prescaled(q1.x, q1.y, q1.z, dx, dy);
            if (!(power > 0.0f)) {
                const float alpha = fminf(GS_ALPHA_MAX, fmul(q1.w, gs_exp(power)));
                if (!(alpha < GS_ALPHA_MIN)) {
                    const float test_T = fmul(T, fsub(1.0f, alpha));
                    if (test_T < GS_T_MIN) {
                        if (T != 0.0f) T_out = T;
                        T = 0.0f;
                    } else {
                        const float w = fmul(alpha, T);
                        const float4 q2 = sl.rec[3 * i + 2];
                        C0 = ffma(q2.x, w, C0); C1 = ffma(q2.y, w, C1); C2 = ffma(q2.z, w, C2);
                        Dd = ffma(q2.w, w, Dd);
                        Aa = fadd(Aa, w);
                        T = test_T;
                        last = __float_as_uint(g0.z);
                    }
                }
            }
        }
        __syncwarp();
    }
    // a bulk copy still in flight targets this CTA's shared memory: observe every issued stage before the warp may exit
    for (int s = waited; s < issued; s++) mbar_wait(&bk.mbar[s & 1], (uint32_t)(s >> 1) & 1u);
    if (inside) {
        if (T != 0.0f) T_out = T;
        const size_t HW = (size_t)a.H * a.W;
        const size_t pix = (size_t)py * a.W + px;
        a.final_T[v * HW + pix] = T_out;
        a.n_contrib[v * HW + pix] = last;
        float *oc = a.out_color + (size_t)v * 3 * HW;
        oc[pix] = ffma(T_out, a.bg[0], C0);
        oc[HW + pix] = ffma(T_out, a.bg[1], C1);
        oc[2 * HW + pix] = ffma(T_out, a.bg[2], C2);
        a.out_depth[v * HW + pix] = Dd;
        a.out_alpha[v * HW + pix] = Aa;
    }
}
#endif // FWD_BULK

void launch_blend_fwd(const BlendArgs &a, cudaStream_t st)
{
    const unsigned grid = (unsigned)(a.grid_x * a.grid_y * CTAS_PER_TILE * a.V);
    blend_fwd_kernel<<<grid, 32 * WARPS_PER_CTA, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// B1
// ------------------------------------------------------------------------------------------------
// Sum ten per-lane values over the warp with 12 shuffles.  After the call, lanes with bit0 == 0 and a valid
// slot hold the warp total of value `slot` (slot_of_lane below); other lanes hold junk.
__device__ __forceinline__ float butterfly10(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7,
                                             float v8, float v9, int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    // xor 16: lower half keeps v0..v4, upper half keeps v5..v9
    float a0 = (h16 ? v5 : v0) + __shfl_xor_sync(FULL, h16 ? v0 : v5, 16);
    float a1 = (h16 ? v6 : v1) + __shfl_xor_sync(FULL, h16 ? v1 : v6, 16);
    float a2 = (h16 ? v7 : v2) + __shfl_xor_sync(FULL, h16 ? v2 : v7, 16);
    float a3 = (h16 ? v8 : v3) + __shfl_xor_sync(FULL, h16 ? v3 : v8, 16);
    float a4 = (h16 ? v9 : v4) + __shfl_xor_sync(FULL, h16 ? v4 : v9, 16);
    // xor 8: h8 == 0 keeps a0,a1,a2 ; h8 == 1 keeps a3,a4
    float c0 = (h8 ? a3 : a0) + __shfl_xor_sync(FULL, h8 ? a0 : a3, 8);
    float c1 = (h8 ? a4 : a1) + __shfl_xor_sync(FULL, h8 ? a1 : a4, 8);
    float c2 = (h8 ? 0.f : a2) + __shfl_xor_sync(FULL, h8 ? a2 : 0.f, 8);
    // xor 4: h4 == 0 keeps c0,c1 ; h4 == 1 keeps c2
    float d0 = (h4 ? c2 : c0) + __shfl_xor_sync(FULL, h4 ? c0 : c2, 4);
    float d1 = (h4 ? 0.f : c1) + __shfl_xor_sync(FULL, h4 ? c1 : 0.f, 4);
    // xor 2: h2 == 0 keeps d0 ; h2 == 1 keeps d1
    float e = (h2 ? d1 : d0) + __shfl_xor_sync(FULL, h2 ? d0 : d1, 2);
    e += __shfl_xor_sync(FULL, e, 1);
    return e;
}
__device__ __forceinline__ int slot_of_lane(int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    if (lane & 1) return -1;
    if ((h4 && h2) || (h8 && h4)) return -1;
    return (h16 ? 5 : 0) + (h8 ? 3 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
}

__global__ void __launch_bounds__(32 * WARPS_PER_CTA, BWD_MIN_BLOCKS) blend_bwd_kernel(BlendBwdArgs a)
{
    __shared__ WarpSlab slabs[WARPS_PER_CTA];

    const int ntiles = a.grid_x * a.grid_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTAs are dispatched in blockIdx order; tile_order lists (view, tile) pairs longest list first, so the heavy
    // tiles of a launch start first and the short ones fill in behind them (LPT schedule)
    const uint32_t vt = a.tile_order[blockIdx.x / CTAS_PER_TILE];
    const int v = (int)(vt / (uint32_t)ntiles), tile = (int)(vt % (uint32_t)ntiles);
    const int patch = (blockIdx.x % CTAS_PER_TILE) * WARPS_PER_CTA + warp; // 0..7: bit0 = x half, bits 1-2 = row band
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int x0 = tile_x * GS_TILE + (patch & 1) * 8, y0 = tile_y * GS_TILE + (patch >> 1) * 4;
    const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float fpx = (float)px, fpy = (float)py, fx0 = (float)x0, fy0 = (float)y0;
    WarpSlab &sl = slabs[warp];
    const uint2 range = a.ranges[(size_t)v * ntiles + tile];
    const float4 *recs4 = reinterpret_cast<const float4 *>(a.recs);
    const uint32_t *plist = a.point_list + range.x;
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)py * a.W + px;
    const uint32_t lt = (1u << lane) - 1u;

    float T_final = 1.f, gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    int last = 0;
    if (inside) {
        T_final = a.final_T[v * HW + pix];
        last = (int)a.n_contrib[v * HW + pix];
        if (a.dL_dcolor) {
            const float *g = a.dL_dcolor + (size_t)v * 3 * HW;
            gC0 = g[pix]; gC1 = g[HW + pix]; gC2 = g[2 * HW + pix];
        }
        if (a.dL_ddepth) gD = a.dL_ddepth[v * HW + pix];
        if (a.dL_dalpha) gA = a.dL_dalpha[v * HW + pix];
    }
    int wmax = last; // list positions at or beyond the patch's largest n_contrib cannot contribute here
    wmax = max(wmax, __shfl_xor_sync(FULL, wmax, 16));
    wmax = max(wmax, __shfl_xor_sync(FULL, wmax, 8));
    wmax = max(wmax, __shfl_xor_sync(FULL, wmax, 4));
    wmax = max(wmax, __shfl_xor_sync(FULL, wmax, 2));
    wmax = max(wmax, __shfl_xor_sync(FULL, wmax, 1));
    if (wmax == 0) return;

    const float bg_dot = a.bg[0] * gC0 + a.bg[1] * gC1 + a.bg[2] * gC2;
    const int slot = slot_of_lane(lane);
#if SGRAD_F64
    double *const sg_slot = reinterpret_cast<double *>(a.sgrad) + (slot >= 0 ? slot : 0); // this lane's column of the accumulator record
#else
    float *const sg_slot = reinterpret_cast<float *>(a.sgrad) + (slot >= 0 ? slot : 0); // this lane's column of ScreenGrad
#endif
    // Per-pixel state of the back-to-front replay.  Upstream keeps five "accumulated colour behind" recurrences
    // (r,g,b,depth,alpha), each rec = last_alpha*last_c + (1-last_alpha)*rec, and forms sum_ch (c_ch - rec_ch)*g_ch.
    // The recurrences are linear with identical coefficients, so their dot product with the pixel's fixed upstream
    // gradient g = (gC, gD, gA) obeys the SAME (convex, numerically stable) recurrence as ONE scalar:
    //   cg_j = c_j.gC + depth_j*gD + gA ;   rg <- last_alpha*last_cg + (1-last_alpha)*rg ;
    //   dL/dalpha_j = T_j*(cg_j - rg) - T_final*(bg.gC)/(1-alpha_j)
    float T = T_final;
    const float Kbg = T_final * bg_dot;
    float rg = 0.f, last_cg = 0.f, last_alpha = 0.f;

    // back to front: chunk [hi-32, hi), lane l <-> list position hi-1-l; the next chunk's loads are in flight
    uint32_t id_c = 0, id_n = 0;
    float4 g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
    if (wmax - 1 - lane >= 0) id_c = __ldg(plist + (wmax - 1 - lane));
    if (wmax - 33 - lane >= 0) id_n = __ldg(plist + (wmax - 33 - lane));
    if (wmax - 1 - lane >= 0) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
    for (int hi = wmax; hi > 0; hi -= 32) {
        const bool hit = box_hits_patch(g0_c, fx0, fy0);
        const uint32_t b = __ballot_sync(FULL, hit);
        float4 g1, g2;
        if (hit) {
            g1 = __ldg(recs4 + 3 * (size_t)id_c + 1);
            g2 = __ldg(recs4 + 3 * (size_t)id_c + 2);
        }
        const float4 g0_h = g0_c;
        const uint32_t id_h = id_c;
        id_c = id_n;
        g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
        if (hi - 33 - lane >= 0) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
        if (hi - 65 - lane >= 0) id_n = __ldg(plist + (hi - 65 - lane));
        if (b == 0u) continue;
        if (hit) {
            const int s = __popc(b & lt);
#if SLAB_PACK
            sl.rec[3 * s] = make_float4(g0_h.x, g0_h.y, __uint_as_float((uint32_t)(hi - 1 - lane)), __uint_as_float(id_h));
#else
            sl.rec[3 * s] = g0_h;
            sl.pos[s] = (uint32_t)(hi - 1 - lane);
            sl.id[s] = id_h;
#endif
            sl.rec[3 * s + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * s + 2] = g2;
        }
        __syncwarp();
        const int cnt = __popc(b);
#if BWD_ILP2
        // Two hits per iteration (SLAB_PACK layout).  Their power / exp / alpha evaluations and, when both contribute, their two
        // reduction butterflies are independent instruction streams; only the short per-pixel recurrences (T, rg) are ordered.
        if ((cnt & 1) && lane == 0) { // pad an odd count: finite centre, zero conic/opacity -> alpha = 0 -> never contributes
            sl.rec[3 * cnt] = make_float4(0.f, 0.f, 0.f, 0.f);
            sl.rec[3 * cnt + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // per-pixel state update of one contributing hit: returns q = G * dL/dalpha and w = alpha * T
        auto pixel_update = [&](const float4 q2, float G, float alpha, float &q, float &w) {
            const float cg = fmaf(q2.x, gC0, fmaf(q2.y, gC1, fmaf(q2.z, gC2, fmaf(q2.w, gD, gA))));
            const float one_m_a = 1.0f - alpha;
            float inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(one_m_a));
#if BWD_T_DIV == 2
            T = __fdiv_rn(T, one_m_a);
#elif BWD_T_DIV == 1
            inv = fmaf(fmaf(-one_m_a, inv, 1.0f), inv, inv);
            T = T * inv;
#else
            T = T * inv;
#endif
            w = alpha * T;
            rg = fmaf(last_alpha, last_cg - rg, rg);
            last_cg = cg;
            last_alpha = alpha;
            q = G * fmaf(T, cg - rg, -(Kbg * inv));
        };
        auto reduce_add = [&](float e, uint32_t rid) {
#if SGRAD_F64
            if (slot >= 0) atomicAdd(sg_slot + GS_SGRAD_F64_DOUBLES * (size_t)rid, (double)e);
#else
            if (slot >= 0) atomicAdd(sg_slot + 12 * (size_t)rid, e);
#endif
        };
        for (int i = 0; i < cnt; i += 2) {
            const float4 g0a = sl.rec[3 * i], q1a = sl.rec[3 * i + 1];
            const float4 g0b = sl.rec[3 * i + 3], q1b = sl.rec[3 * i + 4];
            const float dxa = fsub(g0a.x, fpx), dya = fsub(g0a.y, fpy), dxb = fsub(g0b.x, fpx), dyb = fsub(g0b.y, fpy);
            const float pa = power_prescaled(q1a.x, q1a.y, q1a.z, dxa, dya), pb = power_prescaled(q1b.x, q1b.y, q1b.z, dxb, dyb);
            const float Ga = gs_exp(pa), Gb = gs_exp(pb);
            const float aa = fminf(GS_ALPHA_MAX, fmul(q1a.w, Ga)), ab = fminf(GS_ALPHA_MAX, fmul(q1b.w, Gb));
            const bool ca = ((int)__float_as_uint(g0a.z) < last) && !(pa > 0.0f) && !(aa < GS_ALPHA_MIN);
            const bool cb = ((int)__float_as_uint(g0b.z) < last) && !(pb > 0.0f) && !(ab < GS_ALPHA_MIN);
            const bool any_a = __any_sync(FULL, ca), any_b = __any_sync(FULL, cb);
            if (!any_a && !any_b) continue;
            float qa = 0.f, wa = 0.f, qb = 0.f, wb = 0.f;
            if (ca) pixel_update(sl.rec[3 * i + 2], Ga, aa, qa, wa);
            if (cb) pixel_update(sl.rec[3 * i + 5], Gb, ab, qb, wb);
            if (any_a && any_b) { // the common case: two independent butterflies, interleaved by the scheduler
                const float qxa = qa * dxa, qya = qa * dya, qxb = qb * dxb, qyb = qb * dyb;
                const float ea = butterfly10(qa, qxa, qya, qxa * dxa, qxa * dya, qya * dya, wa * gC0, wa * gC1, wa * gC2, wa * gD, lane);
                const float eb = butterfly10(qb, qxb, qyb, qxb * dxb, qxb * dyb, qyb * dyb, wb * gC0, wb * gC1, wb * gC2, wb * gD, lane);
                reduce_add(ea, __float_as_uint(g0a.w));
                reduce_add(eb, __float_as_uint(g0b.w));
            } else if (any_a) {
                const float qxa = qa * dxa, qya = qa * dya;
                reduce_add(butterfly10(qa, qxa, qya, qxa * dxa, qxa * dya, qya * dya, wa * gC0, wa * gC1, wa * gC2, wa * gD, lane),
                           __float_as_uint(g0a.w));
            } else {
                const float qxb = qb * dxb, qyb = qb * dyb;
                reduce_add(butterfly10(qb, qxb, qyb, qxb * dxb, qxb * dyb, qyb * dyb, wb * gC0, wb * gC1, wb * gC2, wb * gD, lane),
                           __float_as_uint(g0b.w));
            }
        }
#else
        for (int i = 0; i < cnt; i++) {
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
#if SLAB_PACK
            const int pos = (int)__float_as_uint(g0.z); // 0-based position in the tile list
            const uint32_t rid = __float_as_uint(g0.w);
#else
            const int pos = (int)sl.pos[i];
            const uint32_t rid = sl.id[i];
#endif
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = power_prescaled(q1.x, q1.y, q1.z, dx, dy);
            const float G = gs_exp(power);
            const float alpha = fminf(GS_ALPHA_MAX, fmul(q1.w, G));
            const bool contrib = (pos < last) && !(power > 0.0f) && !(alpha < GS_ALPHA_MIN);
            if (!__any_sync(FULL, contrib)) continue;
            float q = 0.f, w = 0.f;
            if (contrib) {
                const float4 q2 = sl.rec[3 * i + 2];
                const float cg = fmaf(q2.x, gC0, fmaf(q2.y, gC1, fmaf(q2.z, gC2, fmaf(q2.w, gD, gA))));
                const float one_m_a = 1.0f - alpha;
                float inv; // 1/(1-alpha), 1-alpha in [0.01, 1)
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(one_m_a));
#if BWD_T_DIV == 2
                T = __fdiv_rn(T, one_m_a); // the oracle's operation, bit for bit
#elif BWD_T_DIV == 1
                inv = fmaf(fmaf(-one_m_a, inv, 1.0f), inv, inv); // one Newton step: removes the MUFU reciprocal's bias, which a
                T = T * inv;                                      // list thousands of contributors deep would otherwise compound
#else
                T = T * inv;
#endif
                w = alpha * T;
                rg = fmaf(last_alpha, last_cg - rg, rg);
                last_cg = cg;
                last_alpha = alpha;
                const float dL_dalpha = fmaf(T, cg - rg, -(Kbg * inv));
                q = G * dL_dalpha;
            }
            // six moments of q over the patch: 1, dx, dy, dx^2, dx*dy, dy^2 (the conic / mean combination is linear
            // and happens once per Gaussian in B2)
            const float qx = q * dx, qy = q * dy;
            const float e = butterfly10(q, qx, qy, qx * dx, qx * dy, qy * dy, w * gC0, w * gC1, w * gC2, w * gD, lane);
#if SGRAD_F64
            if (slot >= 0) atomicAdd(sg_slot + GS_SGRAD_F64_DOUBLES * (size_t)rid, (double)e);
#else
            if (slot >= 0) atomicAdd(sg_slot + 12 * (size_t)rid, e);
#endif
        }
#endif
        __syncwarp();
    }
}

void launch_blend_bwd(const BlendBwdArgs &a, cudaStream_t st)
{
    const unsigned grid = (unsigned)(a.grid_x * a.grid_y * CTAS_PER_TILE * a.V);
    blend_bwd_kernel<<<grid, 32 * WARPS_PER_CTA, 0, st>>>(a);
}

// ScreenGrad holds moments of q = G*dL/dalpha here: S0, Sx, Sy, Sxx, Sxy, Syy = sum q*{1, dx, dy, dx^2, dx*dy, dy^2};
// then colour(3) and depth weights.
int blend_sgrad_is_moments() { return SGRAD_F64 ? 2 : 1; }

__global__ void test_exp_kernel(const float *x, float *y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = gs_exp(x[i]);
}
void launch_test_exp(const float *x, float *y, int64_t n, cudaStream_t st)
{
    test_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n);
}

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
//     Measured on the bench workload: 1.97 of 8 patches survive per instance, 14.3 of 32 lanes contribute.
//   * Tile lists are exactly the reference's (tile/sort indices bit-identical); culling only skips pairs whose
//     alpha is provably < 1/255, so images are unchanged.  Forward arithmetic follows the pinned order of
//     common.cuh: images are bit-identical to the CPU oracle.
//   * Backward: per-pixel terms are expressed as ten sums per Gaussian -- six moments of q = G*dL/dalpha
//     (1, dx, dy, dx^2, dx*dy, dy^2) and four colour/depth weights -- reduced over the
//     32 lanes with a 12-shuffle multi-value butterfly (not 10 x 5 shuffles), then ten lanes each issue one
//     red.global.add.f32 into the Gaussian's 48-byte ScreenGrad record.  Upstream: ~10 atomics per PIXEL.
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

// does the support box [px-hx,px+hx] x [py-hy,py+hy] reach the patch [x0,x0+7] x [y0,y0+3] ?
__device__ __forceinline__ bool box_hits_patch(const float4 g0, float x0, float y0)
{
    return (g0.z >= 0.0f) && (g0.x + g0.z >= x0) && (g0.x - g0.z <= x0 + 7.0f) && (g0.y + g0.w >= y0) && (g0.y - g0.w <= y0 + 3.0f);
}

// Exact version of the same question for the hits of the box test: does the ellipse {Q <= tau}, Q(u,v) = (A u^2 + C v^2)/2 + B u v
// around the Gaussian's centre, tau = ln(255 o) (+ the conditioning-aware margin of preprocess.cu), reach the rectangle?
// The minimum of the convex Q over the rectangle [u0,u1] x [v0,v1] (u = X - px) is 0 if the centre is inside; otherwise it
// lies on the edge facing the centre in x or in y (KKT: on a far edge dQ/dn has the wrong sign because det > 0), where it
// is a clamped 1-D parabola minimum.  Continuous minimum <= minimum over the pixel lattice: conservative.
__device__ __forceinline__ bool ellipse_hits_patch(const float4 g0, const float4 g1, float x0, float y0)
{
    const float A = g1.x, B = g1.y, C = g1.z;
    const float k255 = 255.0f * g1.w;
    const float AC = A * C, dt = AC - B * B;
    const float aniso = __fdividef(AC, dt); // = 1/(1-rho^2) >= 1
    if (!(aniso < 1.0e4f)) return true;     // too ill-conditioned to bound safely (NaN included): never cull
    const float tau = __logf(k255) * (1.0f + 1.0e-5f * aniso) + 0.03f;
    const float u0 = x0 - g0.x, u1 = u0 + 7.0f, v0 = y0 - g0.y, v1 = v0 + 3.0f;
    const float ue = fminf(fmaxf(0.0f, u0), u1), ve = fminf(fmaxf(0.0f, v0), v1);
    const float vs = fminf(fmaxf(__fdividef(-B * ue, C), v0), v1);
    const float us = fminf(fmaxf(__fdividef(-B * ve, A), u0), u1);
    const float Q1 = ue * (0.5f * A * ue + B * vs) + 0.5f * C * vs * vs;
    const float Q2 = us * (0.5f * A * us + B * ve) + 0.5f * C * ve * ve;
    return fminf(Q1, Q2) <= tau;
}

#ifdef BLEND_COUNTERS
// instrumentation build only (tests/gpu_r2_probe.py): visit statistics of the patch walk
__device__ unsigned long long g_blend_cnt[16];
#define CNT_ADD(i, v) do { if (lane == 0) atomicAdd(&g_blend_cnt[i], (unsigned long long)(v)); } while (0)
extern "C" int b200gs_debug_counters(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_blend_cnt, sizeof(g_blend_cnt));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_blend_cnt, z, sizeof(z)); }
    return 0;
}
#else
#define CNT_ADD(i, v) do { } while (0)
#endif

// power with the conic pre-scaled at staging time (A' = -A/2, B' = -B, C' = -C/2: exact operations, so the
// result is bit-identical to gs_power(A,B,C,dx,dy))
__device__ __forceinline__ float power_prescaled(float Ap, float Bp, float Cp, float dx, float dy)
{
    const float u = ffma(Bp, dy, fmul(Ap, dx));
    const float w = fmul(fmul(Cp, dy), dy);
    return ffma(dx, u, w);
}

// per-warp slab of the (at most 32) hits of the current chunk
struct __align__(16) WarpSlab {
    float4 rec[32 * 3]; // px,py,hx,hy | A',B',C',o | r,g,b,depth
    uint32_t pos[32];   // 1-based position in the tile list (forward) / 0-based (backward)
    uint32_t id[32];    // record index (backward: address of the ScreenGrad accumulator)
};

// ------------------------------------------------------------------------------------------------
// F6
// ------------------------------------------------------------------------------------------------
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
#ifdef BLEND_COUNTERS
        CNT_ADD(0, 1); CNT_ADD(1, min(32, n_total - base)); CNT_ADD(2, __popc(b));
        CNT_ADD(3, __popc(__ballot_sync(FULL, hit && ellipse_hits_patch(g0_h, g1, fx0, fy0))));
#endif
        if (b == 0u) continue;
        if (hit) {
            const int slot = __popc(b & lt);
            sl.rec[3 * slot] = g0_h;
            sl.rec[3 * slot + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * slot + 2] = g2;
            sl.pos[slot] = (uint32_t)(base + lane + 1);
        }
        __syncwarp();
        const int cnt = __popc(b);
        for (int i = 0; i < cnt; i++) {
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = power_prescaled(q1.x, q1.y, q1.z, dx, dy);
#ifdef BLEND_COUNTERS
            {
                const bool el = !(power > 0.0f) && !(fminf(GS_ALPHA_MAX, fmul(q1.w, gs_exp(power))) < GS_ALPHA_MIN);
                const uint32_t be = __ballot_sync(FULL, el), bc = __ballot_sync(FULL, el && T != 0.0f);
                CNT_ADD(4, bc != 0u); CNT_ADD(5, __popc(bc)); CNT_ADD(6, be != 0u); CNT_ADD(7, bc != 0u && __popc(bc) < 8);
            }
#endif
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
                        last = sl.pos[i];
                    }
                }
            }
        }
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
    float *const sg_slot = reinterpret_cast<float *>(a.sgrad) + (slot >= 0 ? slot : 0); // this lane's column of ScreenGrad
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
#ifdef BLEND_COUNTERS
        CNT_ADD(8, 1); CNT_ADD(9, min(32, hi)); CNT_ADD(10, __popc(b));
        CNT_ADD(11, __popc(__ballot_sync(FULL, hit && ellipse_hits_patch(g0_h, g1, fx0, fy0))));
#endif
        if (b == 0u) continue;
        if (hit) {
            const int s = __popc(b & lt);
            sl.rec[3 * s] = g0_h;
            sl.rec[3 * s + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * s + 2] = g2;
            sl.pos[s] = (uint32_t)(hi - 1 - lane);
            sl.id[s] = id_h;
        }
        __syncwarp();
        const int cnt = __popc(b);
        for (int i = 0; i < cnt; i++) {
            const int pos = (int)sl.pos[i]; // 0-based position in the tile list
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = power_prescaled(q1.x, q1.y, q1.z, dx, dy);
            const float G = gs_exp(power);
            const float alpha = fminf(GS_ALPHA_MAX, fmul(q1.w, G));
            const bool contrib = (pos < last) && !(power > 0.0f) && !(alpha < GS_ALPHA_MIN);
#ifdef BLEND_COUNTERS
            {
                const uint32_t bc = __ballot_sync(FULL, contrib);
                CNT_ADD(12, bc != 0u); CNT_ADD(13, __popc(bc)); CNT_ADD(15, bc != 0u && __popc(bc) < 8);
            }
#endif
            if (!__any_sync(FULL, contrib)) continue;
            float q = 0.f, w = 0.f;
            if (contrib) {
                const float4 q2 = sl.rec[3 * i + 2];
                const float cg = fmaf(q2.x, gC0, fmaf(q2.y, gC1, fmaf(q2.z, gC2, fmaf(q2.w, gD, gA))));
                const float one_m_a = 1.0f - alpha;
                float inv; // 1/(1-alpha), 1-alpha in [0.01, 1): the single-instruction MUFU reciprocal (<= 1 ulp) suffices
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(one_m_a));
                T = T * inv;
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
            if (slot >= 0) atomicAdd(sg_slot + 12 * (size_t)sl.id[i], e);
        }
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
int blend_sgrad_is_moments() { return 1; }

__global__ void test_exp_kernel(const float *x, float *y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = gs_exp(x[i]);
}
void launch_test_exp(const float *x, float *y, int64_t n, cudaStream_t st)
{
    test_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n);
}

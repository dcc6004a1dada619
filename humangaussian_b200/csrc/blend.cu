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
#ifndef BWD_MMA
#define BWD_MMA 1 // 1: per-Gaussian sums over the patch as a tensor-core contraction; 0: the 12-shuffle butterfly
#endif
#ifndef BWD_MIN_BLOCKS
#if BWD_MMA
#define BWD_MIN_BLOCKS 14 // 14.5 KB of staging per CTA: 14 CTAs (28 warps) fill the SM's shared memory; 72 registers
#else
#define BWD_MIN_BLOCKS 16 // caps the backward kernel at 64 registers (32 warps/SM); 18 / 20 blocks (55 / 48 regs) measured slower
#endif
#endif
#define CTAS_PER_TILE (8 / WARPS_PER_CTA)
#ifndef CULL_EXACT
#define CULL_EXACT 1 // second cull stage: exact ellipse-vs-patch test on the hits of the support-box test
#endif

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
// per-warp register accumulators, ONE atomic per counter per warp at kernel end (per-visit atomics on 16 addresses serialise
// the whole grid: measured the hard way)
#define CNT_DECL unsigned cnt_loc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define CNT_ADD(i, v) do { cnt_loc[(i) & 7] += (unsigned)(v); } while (0)
#define CNT_FLUSH(o) do { if (lane == 0) for (int c_ = 0; c_ < 8; c_++) if (cnt_loc[c_]) atomicAdd(&g_blend_cnt[(o) + c_], (unsigned long long)cnt_loc[c_]); } while (0)
extern "C" int b200gs_debug_counters(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_blend_cnt, sizeof(g_blend_cnt));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_blend_cnt, z, sizeof(z)); }
    return 0;
}
#else
#define CNT_DECL
#define CNT_ADD(i, v) do { } while (0)
#define CNT_FLUSH(o) do { } while (0)
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
    float4 rec[32 * 3]; // px,py,list position,record index | A',B',C',o | r,g,b,depth
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
    CNT_DECL;
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
        bool hit = box_hits_patch(g0_c, fx0, fy0);
        float4 g1, g2;
        if (hit) g1 = __ldg(recs4 + 3 * (size_t)id_c + 1);
        const float4 g0_h = g0_c;
        const uint32_t id_h = id_c;
        id_c = id_n;
        g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
        if (base + 32 + lane < n_total) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
        if (base + 64 + lane < n_total) id_n = __ldg(plist + base + 64 + lane);
#ifdef BLEND_COUNTERS
        CNT_ADD(0, 1); CNT_ADD(1, min(32, n_total - base)); CNT_ADD(2, __popc(__ballot_sync(FULL, hit)));
#endif
#if CULL_EXACT
        if (hit) hit = ellipse_hits_patch(g0_h, g1, fx0, fy0);
#endif
        const uint32_t b = __ballot_sync(FULL, hit);
        if (hit) g2 = __ldg(recs4 + 3 * (size_t)id_h + 2);
        CNT_ADD(3, __popc(b));
        if (b == 0u) continue;
        if (hit) {
            const int slot = __popc(b & lt);
            // (px, py, 1-based list position, -) : the support box has done its job, its two words carry bookkeeping
            sl.rec[3 * slot] = make_float4(g0_h.x, g0_h.y, __uint_as_float((uint32_t)(base + lane + 1)), 0.f);
            sl.rec[3 * slot + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * slot + 2] = g2;
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
                        last = __float_as_uint(g0.z);
                    }
                }
            }
        }
        __syncwarp();
    }
    CNT_FLUSH(0);
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

#if BWD_MMA
// ---- B1, contraction form ---------------------------------------------------------------------------------------------
// The ten per-(warp, Gaussian) sums are two small matrix products over the 32 pixels of the patch:
//     M[hit, f] = sum_pix q[hit, pix] * Phi[pix, f]      Phi = (1, x, y, x^2, x*y, y^2), x,y = pixel offset inside the patch
//     C[hit, c] = sum_pix w[hit, pix] * G[pix, c]        G   = the pixel's upstream gradient (gC0, gC1, gC2, gD)
// Phi and G do not depend on the Gaussian, so both are B operands that stay put while q and w of 16 hits are staged in a
// per-warp shared-memory tile (row = hit, column = pixel lane) and read back as the A fragments of mma.sync.m16n8k8 (TF32
// inputs, FP32 accumulate).  Precision: every FP32 operand is split into two TF32 halves (hi = rna(x), lo = rna(x - hi):
// 22+ significant bits); Phi's entries are integers <= 49, exact in TF32; for G the three products hi*hi + lo*hi + hi*lo
// are kept (the dropped lo*lo term is 2^-24 relative).  The patch-local moments are shifted to the Gaussian's centre
// (dx = cx - x, cx = px - x0) by the three lanes of a quad that hold them, then leave as 8-byte vector reductions.
// Per 16 hits: 20 HMMA + ~150 staging/epilogue instructions instead of 16 x (46-instruction butterfly + 9 products + RED).
#define QS_STRIDE 36 // floats per staged row: 36 = 4 (mod 32) makes the A-fragment reads (row = lane/4, col = lane%4) conflict-free
struct __align__(16) BwdStage {
    float q[16 * QS_STRIDE];
    float w[16 * QS_STRIDE];
    float4 info[16]; // per staged hit: cx = px - x0, cy = py - y0, record index (bits), -
    float ghi[32 * 4], glo[32 * 4]; // per pixel lane: TF32 halves of (gC0, gC1, gC2, gD)
};

__device__ __forceinline__ uint32_t tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo)
{
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void red_add_v2(float *addr, float x, float y)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}

// reduce the nb (<= 16) staged hits of this warp and add them to their ScreenGrad records
__device__ __forceinline__ void bwd_flush(BwdStage &sg, const uint32_t (&phb)[8], float *sgrad, int nb, int lane)
{
    const int gid = lane >> 2, tig = lane & 3;
    __syncwarp();
    float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; j++) { // k-step j = the patch's pixel row j (8 pixels)
        uint32_t hi[4], lo[4];
        const float *qa = sg.q + gid * QS_STRIDE + 8 * j + tig;
        tf32_split(qa[0], hi[0], lo[0]);
        tf32_split(qa[8 * QS_STRIDE], hi[1], lo[1]);
        tf32_split(qa[4], hi[2], lo[2]);
        tf32_split(qa[8 * QS_STRIDE + 4], hi[3], lo[3]);
        mma_tf32(c, hi, phb[2 * j], phb[2 * j + 1]);
        mma_tf32(c, lo, phb[2 * j], phb[2 * j + 1]);
        const float *wa = sg.w + gid * QS_STRIDE + 8 * j + tig;
        tf32_split(wa[0], hi[0], lo[0]);
        tf32_split(wa[8 * QS_STRIDE], hi[1], lo[1]);
        tf32_split(wa[4], hi[2], lo[2]);
        tf32_split(wa[8 * QS_STRIDE + 4], hi[3], lo[3]);
        // B[k = pixel 8j + tig (+4)][n = gid] = G[pixel][gid] for gid < 4, else 0
        const int gi = (8 * j + tig) * 4 + (gid & 3);
        const bool gv = gid < 4;
        const uint32_t gh0 = gv ? __float_as_uint(sg.ghi[gi]) : 0u, gh1 = gv ? __float_as_uint(sg.ghi[gi + 16]) : 0u;
        const uint32_t gl0 = gv ? __float_as_uint(sg.glo[gi]) : 0u, gl1 = gv ? __float_as_uint(sg.glo[gi + 16]) : 0u;
        mma_tf32(d, hi, gh0, gh1);
        mma_tf32(d, lo, gh0, gh1);
        mma_tf32(d, hi, gl0, gl1);
    }
    // c[0],c[1] = features 2*tig, 2*tig+1 of hit gid; c[2],c[3] = the same of hit gid+8   (features: S0 Sx | Sy Sxx | Sxy Syy | - -)
    // d likewise                                                                          (gC0 gC1 | gC2 gD | - - | - -)
    const int q0 = lane & ~3;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const float m0 = c[2 * r], m1 = c[2 * r + 1];
        const float S0 = __shfl_sync(FULL, m0, q0), Sx = __shfl_sync(FULL, m1, q0), Sy = __shfl_sync(FULL, m0, q0 + 1);
        const int row = gid + 8 * r;
        if (row < nb && tig < 3) {
            const float4 inf = sg.info[row];
            const float cx = inf.x, cy = inf.y;
            float o0, o1; // patch-local moments -> moments about the Gaussian's centre: dx = cx - x, dy = cy - y
            if (tig == 0) {
                o0 = m0;                                         // S0
                o1 = cx * S0 - m1;                               // Sx'  = cx S0 - Sx
            } else if (tig == 1) {
                o0 = cy * S0 - m0;                               // Sy'  = cy S0 - Sy
                o1 = cx * (cx * S0 - 2.0f * Sx) + m1;            // Sxx' = cx^2 S0 - 2 cx Sx + Sxx
            } else {
                o0 = cx * (cy * S0 - Sy) - cy * Sx + m0;         // Sxy' = cx cy S0 - cx Sy - cy Sx + Sxy
                o1 = cy * (cy * S0 - 2.0f * Sy) + m1;            // Syy' = cy^2 S0 - 2 cy Sy + Syy
            }
            float *rec = sgrad + 12 * (size_t)__float_as_uint(inf.z);
            red_add_v2(rec + 2 * tig, o0, o1);
            if (tig < 2) red_add_v2(rec + 6 + 2 * tig, d[2 * r], d[2 * r + 1]);
        }
    }
    __syncwarp();
}
#endif // BWD_MMA

__global__ void __launch_bounds__(32 * WARPS_PER_CTA, BWD_MIN_BLOCKS) blend_bwd_kernel(BlendBwdArgs a)
{
    __shared__ WarpSlab slabs[WARPS_PER_CTA];
#if BWD_MMA
    __shared__ BwdStage stages[WARPS_PER_CTA];
#endif

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
#if BWD_MMA
    BwdStage &sg = stages[warp];
    {   // B operands: this lane's pixel gradient as TF32 halves (read back by pixel), and the constant pixel basis
        uint32_t h, l;
        tf32_split(gC0, h, l); sg.ghi[4 * lane + 0] = __uint_as_float(h); sg.glo[4 * lane + 0] = __uint_as_float(l);
        tf32_split(gC1, h, l); sg.ghi[4 * lane + 1] = __uint_as_float(h); sg.glo[4 * lane + 1] = __uint_as_float(l);
        tf32_split(gC2, h, l); sg.ghi[4 * lane + 2] = __uint_as_float(h); sg.glo[4 * lane + 2] = __uint_as_float(l);
        tf32_split(gD, h, l);  sg.ghi[4 * lane + 3] = __uint_as_float(h); sg.glo[4 * lane + 3] = __uint_as_float(l);
    }
    uint32_t phb[8]; // Phi[pixel (x = tig + 4h, y = j)][feature gid], B fragment of k-step j: exact small integers
    {
        const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float x = (float)(tig + 4 * h), y = (float)j;
                const float f = gid == 0 ? 1.0f : gid == 1 ? x : gid == 2 ? y : gid == 3 ? x * x : gid == 4 ? x * y : gid == 5 ? y * y : 0.0f;
                phb[2 * j + h] = __float_as_uint(f);
            }
    }
    float *const sgrad_f = reinterpret_cast<float *>(a.sgrad);
    int nb = 0; // hits staged and not yet reduced
    __syncwarp();
#else
    const int slot = slot_of_lane(lane);
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

    CNT_DECL;
    // back to front: chunk [hi-32, hi), lane l <-> list position hi-1-l; the next chunk's loads are in flight
    uint32_t id_c = 0, id_n = 0;
    float4 g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
    if (wmax - 1 - lane >= 0) id_c = __ldg(plist + (wmax - 1 - lane));
    if (wmax - 33 - lane >= 0) id_n = __ldg(plist + (wmax - 33 - lane));
    if (wmax - 1 - lane >= 0) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
    for (int hi = wmax; hi > 0; hi -= 32) {
        bool hit = box_hits_patch(g0_c, fx0, fy0);
        float4 g1, g2;
        if (hit) g1 = __ldg(recs4 + 3 * (size_t)id_c + 1);
        const float4 g0_h = g0_c;
        const uint32_t id_h = id_c;
        id_c = id_n;
        g0_c = make_float4(0.f, 0.f, -1.f, -1.f);
        if (hi - 33 - lane >= 0) g0_c = __ldg(recs4 + 3 * (size_t)id_c);
        if (hi - 65 - lane >= 0) id_n = __ldg(plist + (hi - 65 - lane));
#ifdef BLEND_COUNTERS
        CNT_ADD(8, 1); CNT_ADD(9, min(32, hi)); CNT_ADD(10, __popc(__ballot_sync(FULL, hit)));
#endif
#if CULL_EXACT
        if (hit) hit = ellipse_hits_patch(g0_h, g1, fx0, fy0);
#endif
        const uint32_t b = __ballot_sync(FULL, hit);
        if (hit) g2 = __ldg(recs4 + 3 * (size_t)id_h + 2);
        CNT_ADD(11, __popc(b));
        if (b == 0u) continue;
        if (hit) {
            const int s = __popc(b & lt);
            // (px, py, 0-based list position, record index)
            sl.rec[3 * s] = make_float4(g0_h.x, g0_h.y, __uint_as_float((uint32_t)(hi - 1 - lane)), __uint_as_float(id_h));
            sl.rec[3 * s + 1] = make_float4(fmul(-0.5f, g1.x), -g1.y, fmul(-0.5f, g1.z), g1.w);
            sl.rec[3 * s + 2] = g2;
        }
        __syncwarp();
        const int cnt = __popc(b);
        for (int i = 0; i < cnt; i++) {
            const float4 g0 = sl.rec[3 * i], q1 = sl.rec[3 * i + 1];
            const int pos = (int)__float_as_uint(g0.z); // 0-based position in the tile list
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
#if BWD_MMA
            // stage this hit's q and w (row nb, column = pixel lane); sixteen staged hits are reduced at once
            sg.q[nb * QS_STRIDE + lane] = q;
            sg.w[nb * QS_STRIDE + lane] = w;
            if (lane == 0) sg.info[nb] = make_float4(g0.x - fx0, g0.y - fy0, g0.w, 0.f);
            if (++nb == 16) {
                bwd_flush(sg, phb, sgrad_f, 16, lane);
                nb = 0;
            }
#else
            // six moments of q over the patch: 1, dx, dy, dx^2, dx*dy, dy^2 (the conic / mean combination is linear
            // and happens once per Gaussian in B2)
            const float qx = q * dx, qy = q * dy;
            const float e = butterfly10(q, qx, qy, qx * dx, qx * dy, qy * dy, w * gC0, w * gC1, w * gC2, w * gD, lane);
            if (slot >= 0) atomicAdd(sg_slot + 12 * (size_t)__float_as_uint(g0.w), e);
#endif
        }
        __syncwarp();
    }
#if BWD_MMA
    if (nb) bwd_flush(sg, phb, sgrad_f, nb, lane);
#endif
    CNT_FLUSH(8);
}

void launch_blend_bwd(const BlendBwdArgs &a, cudaStream_t st)
{
#if BWD_MMA
    static bool once = false; // the staging tiles want the SM's full shared-memory carve-out (14 CTAs x 14.5 KB)
    if (!once) {
        cudaFuncSetAttribute(blend_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        once = true;
    }
#endif
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

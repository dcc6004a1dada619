// blend.cu -- F6 forward alpha blend and B1 backward blend.
//
// One CTA per 16x16 tile (grid.y = view), 8 warps; each warp owns an 8x4 pixel patch so that a
// Gaussian's conservative alpha-support box (GeomRec.hx/hy) can be tested ONCE PER WARP against the
// patch and rejected with a uniform branch before any per-pixel work.  Tile lists are exactly the
// reference's (tile/sort indices stay bit-identical); culling only skips pairs whose alpha is < 1/255.
// Per-instance 48-byte records are staged in shared memory 256 at a time (3 x 128-bit loads per thread).
// Forward arithmetic follows the pinned order of common.cuh, so images are bit-identical to the oracle.
// Backward: per-pixel gradients are reduced across the 32 lanes with shuffles, then ten lanes issue one
// coalesced red.global.add each (instead of upstream's ~10 atomics per pixel per Gaussian).
#include "common.cuh"
#include "kernels.h"

#define BATCH 256

__device__ __forceinline__ uint32_t patch_mask(float px, float py, float hx, float hy, float tx0, float ty0)
{
    // bit (wy*2 + wx): patch columns [tx0+8wx, +7], rows [ty0+4wy, +3] may intersect [px-hx,px+hx]x[py-hy,py+hy]
    if (hx < 0.0f) return 0u; // 255*opacity <= 1: can never reach alpha >= 1/255
    const float xl = px - hx, xh = px + hx, yl = py - hy, yh = py + hy;
    uint32_t xm = 0, ym = 0;
#pragma unroll
    for (int wx = 0; wx < 2; wx++) {
        const float a0 = tx0 + 8.0f * wx;
        if (xh >= a0 && xl <= a0 + 7.0f) xm |= 1u << wx;
    }
#pragma unroll
    for (int wy = 0; wy < 4; wy++) {
        const float b0 = ty0 + 4.0f * wy;
        if (yh >= b0 && yl <= b0 + 3.0f) ym |= 1u << wy;
    }
    uint32_t m = 0;
#pragma unroll
    for (int wy = 0; wy < 4; wy++)
        if (ym & (1u << wy)) m |= xm << (2 * wy);
    return m;
}

// ------------------------------------------------------------------------------------------------
// F6
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_fwd_kernel(BlendArgs a)
{
    __shared__ float4 s0[BATCH], s1[BATCH], s2[BATCH];
    __shared__ uint8_t s_mask[BATCH];

    const int ntiles = a.grid_x * a.grid_y;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * GS_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * GS_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float fpx = (float)px, fpy = (float)py;
    const float tx0 = (float)(tile_x * GS_TILE), ty0 = (float)(tile_y * GS_TILE);

    const uint2 range = a.ranges[(size_t)v * ntiles + tile];
    const int n_total = (int)(range.y - range.x);
    const float4 *recs4 = reinterpret_cast<const float4 *>(a.recs);

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dd = 0.f, Aa = 0.f;
    uint32_t last = 0;

    for (int base = 0; base < n_total; base += BATCH) {
        if (__syncthreads_count(done) == 256) break;
        const int n = min(BATCH, n_total - base);
        if (tid < n) {
            const uint32_t id = __ldg(a.point_list + range.x + base + tid);
            const float4 g0 = __ldg(recs4 + 3 * (size_t)id);
            s0[tid] = g0;
            s1[tid] = __ldg(recs4 + 3 * (size_t)id + 1);
            s2[tid] = __ldg(recs4 + 3 * (size_t)id + 2);
            s_mask[tid] = (uint8_t)patch_mask(g0.x, g0.y, g0.z, g0.w, tx0, ty0);
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, done)) {
            for (int j = 0; j < n; j++) {
                if (!((s_mask[j] >> warp) & 1)) continue; // warp-uniform reject
                if (done) continue;
                const float4 g0 = s0[j], g1 = s1[j];
                const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
                const float power = gs_power(g1.x, g1.y, g1.z, dx, dy);
                if (power > 0.0f) continue;
                const float alpha = fminf(GS_ALPHA_MAX, fmul(g1.w, gs_exp(power)));
                if (alpha < GS_ALPHA_MIN) continue;
                const float test_T = fmul(T, fsub(1.0f, alpha));
                if (test_T < GS_T_MIN) { done = true; continue; }
                const float w = fmul(alpha, T);
                const float4 g2 = s2[j];
                C0 = ffma(g2.x, w, C0); C1 = ffma(g2.y, w, C1); C2 = ffma(g2.z, w, C2);
                Dd = ffma(g2.w, w, Dd);
                Aa = fadd(Aa, w);
                T = test_T;
                last = (uint32_t)(base + j + 1);
            }
        }
    }
    if (inside) {
        const size_t HW = (size_t)a.H * a.W;
        const size_t pix = (size_t)py * a.W + px;
        a.final_T[v * HW + pix] = T;
        a.n_contrib[v * HW + pix] = last;
        float *oc = a.out_color + (size_t)v * 3 * HW;
        oc[pix] = ffma(T, a.bg[0], C0);
        oc[HW + pix] = ffma(T, a.bg[1], C1);
        oc[2 * HW + pix] = ffma(T, a.bg[2], C2);
        a.out_depth[v * HW + pix] = Dd;
        a.out_alpha[v * HW + pix] = Aa;
    }
}

void launch_blend_fwd(const BlendArgs &a, cudaStream_t st)
{
    dim3 grid(a.grid_x * a.grid_y, a.V);
    blend_fwd_kernel<<<grid, 256, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// B1
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__global__ void __launch_bounds__(256) blend_bwd_kernel(BlendBwdArgs a)
{
    __shared__ float4 s0[BATCH], s1[BATCH], s2[BATCH];
    __shared__ uint32_t s_id[BATCH];
    __shared__ uint8_t s_mask[BATCH];
    __shared__ int s_max;

    const int ntiles = a.grid_x * a.grid_y;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * GS_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * GS_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float fpx = (float)px, fpy = (float)py;
    const float tx0 = (float)(tile_x * GS_TILE), ty0 = (float)(tile_y * GS_TILE);
    const uint2 range = a.ranges[(size_t)v * ntiles + tile];
    const float4 *recs4 = reinterpret_cast<const float4 *>(a.recs);
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)py * a.W + px;

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
    if (tid == 0) s_max = 0;
    __syncthreads();
    {
        int m = last;
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
        if (lane == 0) atomicMax(&s_max, m);
    }
    __syncthreads();
    const int n_total = s_max; // entries [0, n_total) of the tile list can have contributed

    const float bg_dot = a.bg[0] * gC0 + a.bg[1] * gC1 + a.bg[2] * gC2;
    const float ddelx_dx = 0.5f * (float)a.W, ddely_dy = 0.5f * (float)a.H;
    float T = T_final;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
    float ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, ad = 0.f, aa = 0.f;

    for (int hi = n_total; hi > 0; hi -= BATCH) {
        const int n = min(BATCH, hi);
        __syncthreads();
        if (tid < n) {
            const uint32_t id = __ldg(a.point_list + range.x + (hi - 1 - tid));
            const float4 g0 = __ldg(recs4 + 3 * (size_t)id);
            s0[tid] = g0;
            s1[tid] = __ldg(recs4 + 3 * (size_t)id + 1);
            s2[tid] = __ldg(recs4 + 3 * (size_t)id + 2);
            s_id[tid] = id;
            s_mask[tid] = (uint8_t)patch_mask(g0.x, g0.y, g0.z, g0.w, tx0, ty0);
        }
        __syncthreads();
        for (int j = 0; j < n; j++) {
            if (!((s_mask[j] >> warp) & 1)) continue; // warp-uniform reject
            const int pos = hi - 1 - j;               // 0-based position in the tile list
            const float4 g0 = s0[j], g1 = s1[j];
            const float dx = fsub(g0.x, fpx), dy = fsub(g0.y, fpy);
            const float power = gs_power(g1.x, g1.y, g1.z, dx, dy);
            const float G = gs_exp(power);
            const float alpha = fminf(GS_ALPHA_MAX, fmul(g1.w, G));
            const bool contrib = (pos < last) && !(power > 0.0f) && !(alpha < GS_ALPHA_MIN);
            if (!__any_sync(0xffffffffu, contrib)) continue;
            float v_dx = 0.f, v_dy = 0.f, v_dA = 0.f, v_dB = 0.f, v_dC = 0.f, v_dO = 0.f;
            float v_r = 0.f, v_g = 0.f, v_b = 0.f, v_dd = 0.f;
            if (contrib) {
                const float4 g2 = s2[j];
                const float one_m_a = 1.0f - alpha;
                T = T / one_m_a;
                const float w = alpha * T;
                float dL_dalpha;
                ac0 = last_alpha * lc0 + (1.f - last_alpha) * ac0; lc0 = g2.x;
                ac1 = last_alpha * lc1 + (1.f - last_alpha) * ac1; lc1 = g2.y;
                ac2 = last_alpha * lc2 + (1.f - last_alpha) * ac2; lc2 = g2.z;
                dL_dalpha = (g2.x - ac0) * gC0 + (g2.y - ac1) * gC1 + (g2.z - ac2) * gC2;
                v_r = w * gC0; v_g = w * gC1; v_b = w * gC2;
                ad = last_alpha * ld + (1.f - last_alpha) * ad; ld = g2.w;
                dL_dalpha += (g2.w - ad) * gD;
                v_dd = w * gD;
                aa = last_alpha + (1.f - last_alpha) * aa;
                dL_dalpha += (1.f - aa) * gA;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / one_m_a) * bg_dot;
                const float dL_dG = g1.w * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * g1.x - gdy * g1.y;
                const float dG_ddely = -gdy * g1.z - gdx * g1.y;
                v_dx = dL_dG * dG_ddelx * ddelx_dx;
                v_dy = dL_dG * dG_ddely * ddely_dy;
                v_dA = -0.5f * gdx * dx * dL_dG;
                v_dB = -0.5f * gdx * dy * dL_dG;
                v_dC = -0.5f * gdy * dy * dL_dG;
                v_dO = G * dL_dalpha;
            }
            v_dx = warp_sum(v_dx); v_dy = warp_sum(v_dy); v_dA = warp_sum(v_dA); v_dB = warp_sum(v_dB);
            v_dC = warp_sum(v_dC); v_dO = warp_sum(v_dO); v_r = warp_sum(v_r); v_g = warp_sum(v_g);
            v_b = warp_sum(v_b); v_dd = warp_sum(v_dd);
            if (lane < 10) {
                float val = v_dx;
                val = lane == 1 ? v_dy : val; val = lane == 2 ? v_dA : val; val = lane == 3 ? v_dB : val;
                val = lane == 4 ? v_dC : val; val = lane == 5 ? v_dO : val; val = lane == 6 ? v_r : val;
                val = lane == 7 ? v_g : val; val = lane == 8 ? v_b : val; val = lane == 9 ? v_dd : val;
                atomicAdd(reinterpret_cast<float *>(a.sgrad + s_id[j]) + lane, val);
            }
        }
    }
}

void launch_blend_bwd(const BlendBwdArgs &a, cudaStream_t st)
{
    dim3 grid(a.grid_x * a.grid_y, a.V);
    blend_bwd_kernel<<<grid, 256, 0, st>>>(a);
}

__global__ void test_exp_kernel(const float *x, float *y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = gs_exp(x[i]);
}
void launch_test_exp(const float *x, float *y, int64_t n, cudaStream_t st)
{
    test_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n);
}

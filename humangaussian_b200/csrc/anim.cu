// anim.cu -- the animation frame path around the rasteriser (reference animation.py):
//   * re-attachment of every Gaussian to the re-posed body mesh: xyz = u*v0 + v*v1 + w*v2 + dist * unit_normal(face)
//     (animation.py:383-403 does this in numpy on the CPU for every frame, then uploads P*12 bytes);
//   * frame packing: clamp(image,0,1) (gs_renderer.py:1017), CHW float -> HWC uint8 by truncation of x*255
//     (animation.py:477-484,1011), so a finished 1024^2 frame leaves the GPU as 3 MB instead of 12 MB.
// Both are streaming kernels: one thread per (frame, Gaussian) / per pixel.
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256) reattach_kernel(int P, int n_frames, int n_verts, int n_faces, const float *__restrict__ vertices,
                                                        const int32_t *__restrict__ faces, const int32_t *__restrict__ map_face,
                                                        const float *__restrict__ map_uvw, const float *__restrict__ map_dist,
                                                        float *__restrict__ xyz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= P) return;
    const int fc = map_face[i];
    float *o = xyz + ((size_t)f * P + i) * 3;
    // an index outside the mesh never reads out of bounds: the Gaussian gets NaN positions (it is then culled by the
    // preprocess and visible to the caller), where the reference's numpy indexing would raise IndexError
    if ((unsigned)fc >= (unsigned)n_faces) { o[0] = o[1] = o[2] = __int_as_float(0x7fc00000); return; }
    const int i0 = faces[3 * fc], i1 = faces[3 * fc + 1], i2 = faces[3 * fc + 2];
    if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) {
        o[0] = o[1] = o[2] = __int_as_float(0x7fc00000);
        return;
    }
    const float *vb = vertices + (size_t)f * n_verts * 3;
    const float a0 = vb[3 * i0], a1 = vb[3 * i0 + 1], a2 = vb[3 * i0 + 2];
    const float b0 = vb[3 * i1], b1 = vb[3 * i1 + 1], b2 = vb[3 * i1 + 2];
    const float c0 = vb[3 * i2], c1 = vb[3 * i2 + 1], c2 = vb[3 * i2 + 2];
    const float e0 = b0 - a0, e1 = b1 - a1, e2 = b2 - a2, g0 = c0 - a0, g1 = c1 - a1, g2 = c2 - a2;
    float n0 = e1 * g2 - e2 * g1, n1 = e2 * g0 - e0 * g2, n2 = e0 * g1 - e1 * g0;
    const float inv = 1.0f / (sqrtf(n0 * n0 + n1 * n1 + n2 * n2) + 1e-20f);
    n0 *= inv; n1 *= inv; n2 *= inv;
    const float u = map_uvw[3 * i], v = map_uvw[3 * i + 1], w = map_uvw[3 * i + 2], d = map_dist[i];
    o[0] = a0 * u + b0 * v + c0 * w + d * n0;
    o[1] = a1 * u + b1 * v + c1 * w + d * n1;
    o[2] = a2 * u + b2 * v + c2 * w + d * n2;
}

void launch_reattach(int P, int n_frames, int n_verts, int n_faces, const float *vertices, const int32_t *faces, const int32_t *map_face,
                     const float *map_uvw, const float *map_dist, float *xyz, cudaStream_t st)
{
    dim3 grid((P + 255) / 256, n_frames);
    reattach_kernel<<<grid, 256, 0, st>>>(P, n_frames, n_verts, n_faces, vertices, faces, map_face, map_uvw, map_dist, xyz);
}

__global__ void __launch_bounds__(256) pack_u8_kernel(const float *__restrict__ color, uint8_t *__restrict__ out, int64_t HW, int n_frames)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (p >= HW) return;
    const float *c = color + (size_t)f * 3 * HW;
    uint8_t *o = out + ((size_t)f * HW + p) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float x = fminf(fmaxf(c[ch * HW + p], 0.0f), 1.0f);
        o[ch] = (uint8_t)(__fmul_rn(x, 255.0f)); // numpy: (clamped float32 * 255).astype(uint8) truncates
    }
}

void launch_pack_u8(const float *color, uint8_t *out, int H, int W, int n_frames, cudaStream_t st)
{
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), n_frames);
    pack_u8_kernel<<<grid, 256, 0, st>>>(color, out, HW, n_frames);
}

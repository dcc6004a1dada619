"""Independent float64 / torch.autograd check of the oracle's calculus (TEST INFRASTRUCTURE ONLY).

A dense (every pixel x every Gaussian) formulation of SURVEY.md Appendix A.2-A.4 in float64,
differentiated by torch.autograd.  It shares NO code with gs_oracle.c: agreement of the two
(tests/test_oracle_autograd.py) validates the oracle's hand-derived backward (A.5-A.7).
Integer decisions (radius, tile rectangle, sort order) are taken from the caller (the oracle's
state) because they carry no gradient; the continuous skip rules are re-evaluated here.

Upstream quirks reproduced on purpose so gradients are comparable:
  * min(0.99, o*G) is straight-through in backward (A.5);
  * when tx/tz is clamped to +-1.3 tanfov the clamped value is treated as a constant (A.6);
  * means2D is a dummy additive NDC offset whose gradient is dL/d(pixel)*(W/2, H/2) (A.5).
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def _sh(deg, sh, d):
    # sh [P,K,3], d [P,3]; basis of gaussiansplatting/utils/sh_utils.py:74-100
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    r = C0 * sh[:, 0]
    if deg > 0:
        r = r - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
             + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        r = (r + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
             + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
             + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
             + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return r


def dense_render(*, means3D, means2D, opacities, viewmatrix, projmatrix, campos, bg, H, W, tanfovx, tanfovy,
                 radii, rect, order, sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
                 cov3D_precomp=None, scale_modifier=1.0):
    """All float tensors float64.  radii[P] int, rect[P,4] int (tile units), order = Gaussian ids sorted by
    (depth, id) -- from the oracle.  Returns color[3,H,W], depth[1,H,W], alpha[1,H,W], n_contrib-like
    count of kept Gaussians per pixel (for decision cross-checks)."""
    dt = torch.float64
    P = means3D.shape[0]
    vis = torch.as_tensor(radii > 0)
    V, PV = viewmatrix.to(dt), projmatrix.to(dt)
    ph = torch.cat([means3D, torch.ones(P, 1, dtype=dt)], 1)
    pv = ph @ V            # row-vector convention (cameras.py:50-52)
    hom = ph @ PV
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None] + means2D[:, :2]
    if cov3D_precomp is None:
        q = rotations
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(P, 3, 3)
        L = R * (scale_modifier * scales)[:, None, :]
        S = L @ L.transpose(1, 2)
    else:
        c = cov3D_precomp
        S = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(P, 3, 3)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    tz = pv[:, 2]
    tz_safe = torch.where(vis, tz, torch.ones_like(tz))

    def clamp_const(t, lim):
        ratio = t / tz_safe
        clamped = (ratio < -lim) | (ratio > lim)
        return torch.where(clamped, (ratio.clamp(-lim, lim) * tz_safe).detach(), t)

    cx, cy = clamp_const(pv[:, 0], 1.3 * tanfovx), clamp_const(pv[:, 1], 1.3 * tanfovy)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz_safe, zero, -fx * cx / tz_safe ** 2, zero, fy / tz_safe, -fy * cy / tz_safe ** 2], -1).reshape(P, 2, 3)
    W3 = V[:3, :3].T       # w2c rotation
    M = J @ W3
    cov = M @ S @ M.transpose(1, 2)
    a, b, c_ = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c_ - b * b
    det = torch.where(vis, det, torch.ones_like(det))
    A, B, Cc = c_ / det, -b / det, a / det
    px = ((ndc[:, 0] + 1) * W - 1) * 0.5
    py = ((ndc[:, 1] + 1) * H - 1) * 0.5
    if shs is not None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(_sh(sh_degree, shs, d) + 0.5, 0.0)
    else:
        rgb = colors_precomp
    depth = tz

    order = torch.as_tensor(order, dtype=torch.long)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    fxp, fyp = xs.reshape(-1), ys.reshape(-1)            # [N]
    tile_x, tile_y = (fxp // 16).long(), (fyp // 16).long()
    rect = torch.as_tensor(rect, dtype=torch.long)[order]
    inrect = ((tile_x[:, None] >= rect[None, :, 0]) & (tile_x[:, None] < rect[None, :, 2]) &
              (tile_y[:, None] >= rect[None, :, 1]) & (tile_y[:, None] < rect[None, :, 3]) & vis[order][None, :])
    dx = px[order][None, :] - fxp[:, None]
    dy = py[order][None, :] - fyp[:, None]
    Ao, Bo, Co = A[order][None], B[order][None], Cc[order][None]
    power = -0.5 * (Ao * dx * dx + Co * dy * dy) - Bo * dx * dy
    G = torch.exp(torch.clamp(power, max=0.0))
    aG = opacities.reshape(-1)[order][None] * G
    alpha = aG + (torch.clamp(aG, max=0.99) - aG).detach()
    valid = inrect & (power <= 0) & (alpha >= 1.0 / 255.0)
    av = torch.where(valid, alpha, torch.zeros_like(alpha))
    Tincl = torch.cumprod(1 - av, dim=1)
    Tbefore = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    stop = valid & (Tincl < 1e-4)
    excluded = torch.cumsum(stop.to(torch.int32), 1) > 0
    keep = valid & ~excluded
    w = torch.where(keep, alpha * Tbefore, torch.zeros_like(alpha))
    Tfinal = torch.prod(torch.where(keep, 1 - alpha, torch.ones_like(alpha)), dim=1)
    color = w @ rgb[order] + Tfinal[:, None] * bg[None]
    dep = w @ depth[order]
    alp = w.sum(1)
    return (color.T.reshape(3, H, W), dep.reshape(1, H, W), alp.reshape(1, H, W), keep.sum(1).reshape(H, W))

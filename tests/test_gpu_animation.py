"""GPU: the animation frame batch (BASELINE config 5 shape): per-view Gaussian centres through the batched entry,
the re-attachment kernel against a numpy restatement of animation.py:383-403, and the uint8 frame pack."""
import math

import numpy as np
import pytest
import torch

from util import grads_agree, small_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _numpy_reattach(vertices, faces, mface, uvw, dist):
    # animation.py:383-403, verbatim semantics
    f = faces[mface]
    v0, v1, v2 = vertices[f[:, 0]], vertices[f[:, 1]], vertices[f[:, 2]]
    n = np.cross(v1 - v0, v2 - v0)
    n = n / (np.linalg.norm(n, axis=1, keepdims=True) + 1e-20)
    return v0 * uvw[:, [0]] + v1 * uvw[:, [1]] + v2 * uvw[:, [2]] + dist[:, None] * n


def test_reattach_matches_reference_formula():
    from humangaussian_b200.animation import reattach
    rng = np.random.RandomState(0)
    Nv, Nf, P, F = 500, 900, 20000, 5
    verts = rng.randn(F, Nv, 3).astype(np.float32)
    faces = np.stack([rng.permutation(Nv)[:3] for _ in range(Nf)]).astype(np.int32)
    mface = rng.randint(0, Nf, P).astype(np.int32)
    uvw = rng.dirichlet([1, 1, 1], P).astype(np.float32)
    dist = (rng.randn(P) * 0.01).astype(np.float32)
    got = reattach(torch.tensor(verts, device=DEV), torch.tensor(faces), torch.tensor(mface), torch.tensor(uvw), torch.tensor(dist)).cpu().numpy()
    for f in range(F):
        want = _numpy_reattach(verts[f].astype(np.float64), faces, mface, uvw.astype(np.float64), dist.astype(np.float64))
        assert np.allclose(got[f], want, atol=2e-6, rtol=1e-5)


def test_pack_frames_matches_numpy_truncation():
    from humangaussian_b200.animation import pack_frames_u8
    rng = np.random.RandomState(1)
    c = (rng.rand(3, 3, 37, 53).astype(np.float32) * 1.4 - 0.2)
    c[0, 0, 0, :6] = [0.0, 1.0, 0.5, 1.0 / 255.0, 254.999 / 255.0, np.float32(0.999999)]
    got = pack_frames_u8(torch.tensor(c, device=DEV)).cpu().numpy()
    want = (np.clip(c, 0, 1).transpose(0, 2, 3, 1) * 255).astype(np.uint8)
    assert got.shape == (3, 37, 53, 3) and np.array_equal(got, want)


def test_per_view_positions_match_the_oracle_frame_by_frame():
    """means3D [V,P,3]: every frame is bit-identical to rendering that frame's positions alone; position gradients stay per frame."""
    from humangaussian_b200.cameras import sample_orbit_cameras
    from humangaussian_b200.rasterizer import rasterize_views
    from humangaussian_b200.renderer import stack_cameras
    from oracle.gs_oracle import Oracle
    inp, _, _ = small_scene(P=1500, deg=1, seed=12, H=64, W=80)
    V = 3
    cams = sample_orbit_cameras(V, 64, 80, seed=3)
    rng = np.random.RandomState(0)
    xyz = np.stack([inp["means3D"] + rng.randn(*inp["means3D"].shape).astype(np.float32) * 0.01 * f for f in range(V)]).astype(np.float32)
    vm, pm, cp, tanx, tany = stack_cameras(cams, DEV)
    t = {k: torch.tensor(inp[k], device=DEV, requires_grad=True) for k in ("opacities", "shs", "scales", "rotations")}
    x = torch.tensor(xyz, device=DEV, requires_grad=True)
    c, r, d, a = rasterize_views(means3D=x, opacities=t["opacities"], viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx,
                                 tanfovy=tany, image_height=64, image_width=80, bg=torch.tensor(inp["bg"], device=DEV), sh_degree=1,
                                 shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    gw = [torch.tensor(rng.randn(V, ch, 64, 80).astype(np.float32), device=DEV) for ch in (3, 1, 1)]
    torch.autograd.backward([c, d, a], gw)
    assert x.grad.shape == (V, 1500, 3)
    sum_g = {k: 0 for k in t}
    for f in range(V):
        o = Oracle()
        kw = dict(inp, means3D=xyz[f], viewmatrix=vm[f].cpu().numpy(), projmatrix=pm[f].cpu().numpy(), campos=cp[f].cpu().numpy(),
                  tanfovx=tanx[f], tanfovy=tany[f], image_height=64, image_width=80)
        col, rad, dep, alp = o.forward(**kw)
        assert np.array_equal(c[f].detach().cpu().numpy(), col) and np.array_equal(r[f].cpu().numpy(), rad)
        assert np.array_equal(d[f].detach().cpu().numpy(), dep) and np.array_equal(a[f].detach().cpu().numpy(), alp)
        g = o.backward(gw[0][f].cpu().numpy(), gw[1][f].cpu().numpy(), gw[2][f].cpu().numpy())
        ok, msg = grads_agree(x.grad[f].cpu().numpy(), g["means3D"])
        assert ok, f"frame {f} dL/dmeans3D: {msg}"
        for k in t:
            sum_g[k] = sum_g[k] + g[k].astype(np.float64)
    for k in t:
        ok, msg = grads_agree(t[k].grad.cpu().numpy().reshape(sum_g[k].shape), sum_g[k])
        assert ok, f"dL/d{k} summed over frames: {msg}"


def test_render_frames_end_to_end():
    """Config-5 shape in miniature: proxy mesh -> reattach -> batched forward with per-frame centres -> uint8 frames."""
    from humangaussian_b200.animation import reattach, render_frames
    from humangaussian_b200.cameras import MiniCamC2W, orbit_c2w
    from humangaussian_b200.scene import synthetic_body
    p = synthetic_body(20000, sh_degree=0, seed=4).to(DEV)
    rng = np.random.RandomState(2)
    # a crude proxy "mesh": random triangles near each Gaussian, with barycentric coords reproducing the rest pose
    P, F = p.P, 70  # > 64 frames: exercises chunking
    xyz0 = p.xyz.cpu().numpy()
    tri = xyz0[:, None, :] + rng.randn(P, 3, 3).astype(np.float32) * 0.02
    verts0 = tri.reshape(-1, 3)
    faces = np.arange(3 * P, dtype=np.int32).reshape(P, 3)
    uvw = np.full((P, 3), 1 / 3, np.float32)
    centroid = tri.mean(1)
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]); n /= np.linalg.norm(n, axis=1, keepdims=True) + 1e-20
    dist = ((xyz0 - centroid) * n).sum(1).astype(np.float32)
    sway = np.stack([verts0 + np.array([0.02 * math.sin(f / 5), 0, 0], np.float32) for f in range(F)])
    xyz = reattach(torch.tensor(sway, device=DEV), torch.tensor(faces), torch.arange(P, dtype=torch.int32), torch.tensor(uvw), torch.tensor(dist))
    assert xyz.shape == (F, P, 3)
    cams = [MiniCamC2W(orbit_c2w(0.0, float(f), 2.0).numpy(), 128, 128, math.radians(50), math.radians(50), 0.01, 100.0, device=DEV) for f in range(F)]
    frames = render_frames(p, xyz, cams, torch.zeros(3, device=DEV))
    assert frames.shape == (F, 128, 128, 3) and frames.dtype == torch.uint8
    assert int(frames.max()) > 50 and not torch.equal(frames[0], frames[35])


def test_reattach_out_of_range_indices_are_not_dereferenced():
    """A mapping_face / face index outside the mesh gives NaN positions (culled downstream), never an out-of-bounds read."""
    from humangaussian_b200.animation import reattach
    rng = np.random.RandomState(0)
    Nv, Nf, P = 50, 80, 1000
    verts = rng.randn(2, Nv, 3).astype(np.float32)
    faces = np.stack([rng.permutation(Nv)[:3] for _ in range(Nf)]).astype(np.int32)
    faces[7, 1] = Nv + 1000000                       # a corrupt face
    mface = rng.randint(0, Nf, P).astype(np.int32)
    mface[:5] = [-1, Nf, Nf + 12345678, 7, 3]
    uvw = rng.dirichlet([1, 1, 1], P).astype(np.float32)
    dist = np.zeros(P, np.float32)
    got = reattach(torch.tensor(verts, device=DEV), torch.tensor(faces), torch.tensor(mface), torch.tensor(uvw), torch.tensor(dist))
    torch.cuda.synchronize()
    bad = (mface < 0) | (mface >= Nf) | (mface == 7)
    g = got.cpu().numpy()
    assert np.isnan(g[:, bad]).all() and np.isfinite(g[:, ~bad]).all()

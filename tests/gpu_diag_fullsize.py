"""Diagnostic (not a test): where do the worst gradient rows of the full-size parity checks come from?
    python tests/gpu_diag_fullsize.py [config4|config2]"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from test_gpu_fullsize import _inputs
from test_gpu_parity import DEV, _gpu_inputs, _oracle, _settings
from humangaussian_b200.cameras import Camera, orbit_c2w, sample_orbit_cameras
from humangaussian_b200.rasterizer import GaussianRasterizer
from humangaussian_b200.scene import sample_ply_scene

which = sys.argv[1] if len(sys.argv) > 1 else "config4"
if which == "config4":
    p = sample_ply_scene(300000, 3)
    cam = sample_orbit_cameras(64, 1024, 1024, seed=1000)[0]
    inp, seed = _inputs(p, cam, 1024, 1024, 3), 22
else:
    p = sample_ply_scene()
    cam = Camera(orbit_c2w(15.0, 0.0, 2.0), math.radians(70.0), 512, 512)
    inp, seed = _inputs(p, cam, 512, 512, 0), 21
o_out, o_st, gimg, og = _oracle(inp, seed)
t = _gpu_inputs(inp)
P = inp["means3D"].shape[0]
m2d = torch.zeros(P, 3, device=DEV, requires_grad=True)
c, rad, d, a = GaussianRasterizer(_settings(inp))(means3D=t["means3D"], means2D=m2d, shs=t["shs"], opacities=t["opacities"], scales=t["scales"],
                                                    rotations=t["rotations"])
gC, gD, gA = (torch.tensor(g, device=DEV) for g in gimg)
((c * gC).sum() + (d * gD).sum() + (a * gA).sum()).backward()
got = {k: t[k].grad.cpu().numpy().reshape(P, -1) for k in t}
got["means2D"] = m2d.grad.cpu().numpy()
ref = {k: og[k].reshape(P, -1) for k in got}
ratio = {}
for k in got:
    err = np.abs(got[k].astype(np.float64) - ref[k])
    lim = 1e-5 + 1e-4 * np.abs(ref[k]).max(axis=1, keepdims=True)
    ratio[k] = (err / lim).max(axis=1)
    print(f"{k:10s} worst {ratio[k].max():.2f}x  rows>1: {(ratio[k] > 1).sum()}  rows>0.5: {(ratio[k] > 0.5).sum()}")
co = o_st["conic_opacity"]
aniso = co[:, 0] * co[:, 2] / np.maximum(co[:, 0] * co[:, 2] - co[:, 1] ** 2, 1e-30)
bad = np.unique(np.concatenate([np.argsort(-ratio[k])[:4] for k in ("means3D", "scales", "means2D")]))
print("row      radius tiles  opac     aniso     |  " + "  ".join(f"{k:>9s}" for k in ratio))
for i in bad:
    print(f"{i:7d} {o_st['tiles_touched'][i] and int(o_out[1][i]):6d} {int(o_st['tiles_touched'][i]):5d} {co[i, 3]:.4f} {aniso[i]:9.1f}   |  " +
          "  ".join(f"{ratio[k][i]:9.2f}" for k in ratio))
    print("         means2D got", got["means2D"][i], "ref", ref["means2D"][i])
    print("         means3D got", got["means3D"][i], "ref", ref["means3D"][i])
    print("         scales  got", got["scales"][i], "ref", ref["scales"][i])

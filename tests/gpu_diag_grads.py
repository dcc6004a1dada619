"""Diagnostic (not a test): per-tensor gradient agreement GPU vs oracle on the medium 512^2 scene."""
import sys
import numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from util import close, close_rows, grad_images, small_scene
from test_gpu_parity import _oracle, _gpu_inputs, _settings, DEV
from humangaussian_b200.rasterizer import GaussianRasterizer
inp, _, _ = small_scene(P=60000, deg=0, seed=11, H=512, W=512, big=False, dist=2.0, fovy_deg=70.0, elev=15.0, azim=0.0)
o_out, o_st, gimg, og = _oracle(inp, 5)
t = _gpu_inputs(inp)
m2d = torch.zeros(60000, 3, device=DEV, requires_grad=True)
c, rad, d, a = GaussianRasterizer(_settings(inp))(means3D=t["means3D"], means2D=m2d, shs=t["shs"], opacities=t["opacities"], scales=t["scales"], rotations=t["rotations"])
gC, gD, gA = (torch.tensor(g, device=DEV) for g in gimg)
((c * gC).sum() + (d * gD).sum() + (a * gA).sum()).backward()
t["means2D"] = m2d
for k in t:
    g = t[k].grad.cpu().numpy().reshape(og[k].shape); r = og[k]
    err = np.abs(g.astype(np.float64) - r); lim = 1e-5 + 1e-4 * np.abs(r)
    bad = err > lim
    print(f"{k:10s} elementwise worst {float((err/lim).max()):8.2f}x  failing {int(bad.sum())}/{bad.size}  rowwise {close_rows(g, r)}  max|g| {np.abs(r).max():.3g}")
    if bad.any():
        i = np.unravel_index(np.argmax(err / lim), err.shape)
        print("     worst at", i, "gpu", g[i], "oracle", r[i], "row", r[i[0]].ravel()[:8])

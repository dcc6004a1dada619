"""Ad-hoc: run-to-run spread of the full-size backward (float atomics) under the linearity test's criterion."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from util import grads_agree
from humangaussian_b200.cameras import sample_orbit_cameras
from humangaussian_b200.rasterizer import rasterize_views
from humangaussian_b200.renderer import stack_cameras
from humangaussian_b200.scene import synthetic_body
DEV = "cuda:0"
p = synthetic_body(300_000, sh_degree=3, seed=0).to(DEV)
cams = sample_orbit_cameras(4, 1024, 1024, seed=1000, device=DEV)
with torch.no_grad():
    t = dict(means3D=p.get_xyz, opacities=p.get_opacity, shs=p.get_features.contiguous(), scales=p.get_scaling, rotations=p.get_rotation)
vm, pm, cp, tanx, tany = stack_cameras(cams, DEV)
leaves = {k: v.clone().requires_grad_(True) for k, v in t.items()}
g = torch.Generator(device=DEV).manual_seed(0)
gw = [torch.randn(4, ch, 1024, 1024, device=DEV, generator=g) for ch in (3, 1, 1)]
def grads(scale):
    for v in leaves.values(): v.grad = None
    out = rasterize_views(means3D=leaves["means3D"], opacities=leaves["opacities"], viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx,
                          tanfovy=tany, image_height=1024, image_width=1024, bg=torch.zeros(3, device=DEV), sh_degree=3, shs=leaves["shs"],
                          scales=leaves["scales"], rotations=leaves["rotations"])
    torch.autograd.backward([out[0], out[2], out[3]], [w * scale for w in gw])
    return {k: v.grad.clone().cpu().numpy() for k, v in leaves.items()}
worst = {}
for it in range(8):
    g1, g2 = grads(1.0), grads(2.0)
    for k in g1:
        ok, msg = grads_agree(g2[k] * 0.5, g1[k], atol=1e-5, rtol=2e-4)
        worst.setdefault(k, []).append((ok, msg))
for k, v in worst.items():
    print(k, "fails", sum(1 for ok, _ in v if not ok), "/", len(v))
    for ok, msg in v: print("   ", ok, msg)

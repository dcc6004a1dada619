"""Ad-hoc timing of the hot path (not a test, not the bench): python tests/gpu_quick_timing.py [P] [V] [HW] [deg]"""
import math
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from humangaussian_b200.cameras import sample_orbit_cameras
from humangaussian_b200.rasterizer import launch_count, rasterize_views
from humangaussian_b200.renderer import stack_cameras
from humangaussian_b200.scene import synthetic_body

P = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 8
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
deg = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda:0"
p = synthetic_body(P, sh_degree=deg, seed=0).to(dev)
cams = sample_orbit_cameras(V, HW, HW, seed=0, device=dev)
vm, pm, cp, tanx, tany = stack_cameras(cams, dev)
bg = torch.zeros(3, device=dev)
with torch.no_grad():
    xyz, op, sh, sc, rot = p.get_xyz, p.get_opacity, p.get_features.contiguous(), p.get_scaling, p.get_rotation
for t in (xyz, op, sh, sc, rot):
    t.requires_grad_(True)
gw = [torch.randn(V, c, HW, HW, device=dev) for c in (3, 1, 1)]


def step(bwd=True):
    c, r, d, a = rasterize_views(means3D=xyz, opacities=op, viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx, tanfovy=tany,
                                 image_height=HW, image_width=HW, bg=bg, sh_degree=deg, shs=sh, scales=sc, rotations=rot)
    if bwd:
        torch.autograd.backward([c, d, a], gw)
    return r


for bwd in (False, True):
    for _ in range(3):
        step(bwd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        r = step(bwd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"P={P} V={V} {HW}x{HW} deg={deg} {'fwd+bwd' if bwd else 'fwd'}: {ms:.3f} ms/batch, {ms / V:.3f} ms/view, {V / ms * 1e3:.1f} views/s; vis={int((r > 0).sum()) / V:.0f}")
print("launches", launch_count())

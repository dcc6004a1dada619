"""Ad-hoc (run under compute-sanitizer): one small pass through every kernel of the library."""
import math, sys
import numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from util import small_scene, grad_images
from humangaussian_b200 import rasterizer as R
from humangaussian_b200.animation import reattach, pack_frames_u8
from humangaussian_b200.cameras import sample_orbit_cameras
from humangaussian_b200.renderer import stack_cameras
DEV = "cuda:0"
inp, _, _ = small_scene(P=3000, deg=3, seed=1, H=72, W=104)
cams = sample_orbit_cameras(3, 72, 104, seed=2, device=DEV)
vm, pm, cp, tanx, tany = stack_cameras(cams, DEV)
t = {k: torch.tensor(inp[k], device=DEV, requires_grad=True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
c, r, d, a = R.rasterize_views(means3D=t["means3D"], opacities=t["opacities"], viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx,
                               tanfovy=tany, image_height=72, image_width=104, bg=torch.tensor(inp["bg"], device=DEV), sh_degree=3,
                               shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
(c.sum() + d.sum() + a.sum()).backward()
xyz = t["means3D"].detach()[None].repeat(3, 1, 1).contiguous()
c2 = R.rasterize_views(means3D=xyz, opacities=t["opacities"].detach(), viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx, tanfovy=tany,
                       image_height=72, image_width=104, bg=torch.zeros(3, device=DEV), sh_degree=3, shs=t["shs"].detach(),
                       scales=t["scales"].detach(), rotations=t["rotations"].detach())[0]
pack_frames_u8(c2)
k = torch.randint(0, 1 << 18, (50000,), dtype=torch.int32, device=DEV)
R.device_sort_pairs(k, torch.arange(50000, dtype=torch.int32, device=DEV), 18)
R.distCUDA2(torch.rand(5000, 3, device=DEV))
R.device_exp(-torch.rand(1000, device=DEV) * 10)
verts = torch.rand(2, 100, 3, device=DEV); faces = torch.randint(0, 100, (150, 3), dtype=torch.int32)
reattach(verts, faces, torch.randint(0, 150, (3000,), dtype=torch.int32), torch.rand(3000, 3), torch.rand(3000) * 0.01)
# packed entry (B3 writes one flat gradient buffer) and the densify / prune kernels
from humangaussian_b200.dist import pack
from humangaussian_b200.densify import DensifyStats, add_densification_stats, densify_and_prune, prune_only
flat = pack(dict(xyz=t["means3D"].detach(), scaling=t["scales"].detach(), rotation=t["rotations"].detach(),
                 opacity=t["opacities"].detach().reshape(-1, 1), features=t["shs"].detach())).requires_grad_(True)
m2d = torch.zeros(3, 3000, 3, device=DEV, requires_grad=True)
cp_, rp_, dp_, ap_ = R.rasterize_views_packed(flat, 3000, 16, viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx, tanfovy=tany,
                                              image_height=72, image_width=104, bg=torch.zeros(3, device=DEV), sh_degree=3, means2D=m2d)
(cp_.sum() + dp_.sum() + ap_.sum()).backward()
stats = DensifyStats.zeros(3000, DEV)
add_densification_stats(stats, m2d.grad, rp_)
g = torch.Generator(device=DEV).manual_seed(0)
raw = dict(xyz=torch.randn(3000, 3, device=DEV, generator=g), f_dc=torch.randn(3000, 1, 3, device=DEV, generator=g),
           f_rest=torch.randn(3000, 15, 3, device=DEV, generator=g), opacity=torch.randn(3000, 1, device=DEV, generator=g) * 2 - 1.5,
           scaling=torch.randn(3000, 3, device=DEV, generator=g) * 0.9 - 5.0, rotation=torch.randn(3000, 4, device=DEV, generator=g))
mom = {k: (torch.randn_like(v), torch.rand_like(v)) for k, v in raw.items()}
thr = float(torch.quantile((stats.xyz_gradient_accum / stats.denom.clamp_min(1)).flatten(), 0.6))
p2, m2, s2 = densify_and_prune(raw, mom, stats, thr, 0.05, 0.7, 20.0, 0.01)
prune_only(p2, m2, s2, 0.05, 0.02)
torch.cuda.synchronize()
print("sanitize pass done; launches", R.launch_count())

"""Round-2 probe (not a test): stage timings of the product library on the real workload (300 k subsample of sample.ply,
SH deg 3, 1024^2) and, when B200GS_LIB points at the -DBLEND_COUNTERS build, the visit statistics of the patch walk.

    python tests/gpu_r2_probe.py [V] [scene: sample|synthetic]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from humangaussian_b200 import rasterizer as R
from humangaussian_b200.cameras import sample_orbit_cameras
from humangaussian_b200.renderer import stack_cameras
from humangaussian_b200.scene import sample_ply_scene, synthetic_body

V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = sys.argv[2] if len(sys.argv) > 2 else "sample"
per_view = len(sys.argv) > 3 and sys.argv[3] == "pv"  # V single-view calls per step (the reference's calling pattern)
P, HW, deg = 300000, 1024, 3
dev = "cuda:0"
p = (sample_ply_scene(P, deg) if scene == "sample" else synthetic_body(P, sh_degree=deg, seed=0)).to(dev)
cams = sample_orbit_cameras(64, HW, HW, seed=1000, device=dev)[:V]
vm, pm, cp, tanx, tany = stack_cameras(cams, dev)
bg = torch.zeros(3, device=dev)
with torch.no_grad():
    xyz, op, sh, sc, rot = p.get_xyz, p.get_opacity, p.get_features.contiguous(), p.get_scaling, p.get_rotation
for t in (xyz, op, sh, sc, rot):
    t.requires_grad_(True)
g = torch.Generator(device=dev).manual_seed(0)
gw = [torch.randn(V, c, HW, HW, device=dev, generator=g) for c in (3, 1, 1)]


def step():
    if per_view:
        outs = [R.rasterize_views(means3D=xyz, opacities=op, viewmatrices=vm[i:i + 1], projmatrices=pm[i:i + 1], camposs=cp[i:i + 1],
                                  tanfovx=tanx[i:i + 1], tanfovy=tany[i:i + 1], image_height=HW, image_width=HW, bg=bg, sh_degree=deg,
                                  shs=sh, scales=sc, rotations=rot) for i in range(V)]
        for i, (c, r, d, a) in enumerate(outs):
            torch.autograd.backward([c, d, a], [g[i:i + 1] for g in gw])
        return torch.cat([o[1] for o in outs], 0)
    c, r, d, a = R.rasterize_views(means3D=xyz, opacities=op, viewmatrices=vm, projmatrices=pm, camposs=cp, tanfovx=tanx, tanfovy=tany,
                                   image_height=HW, image_width=HW, bg=bg, sh_degree=deg, shs=sh, scales=sc, rotations=rot)
    torch.autograd.backward([c, d, a], gw)
    return r


L = R.load_library()
counters = hasattr(L, "b200gs_debug_counters")
for _ in range(3):
    step()
torch.cuda.synchronize()
out = {"scene": scene, "V": V, "per_view": per_view, "lib": os.path.basename(R.LIB_PATH)}
if counters:
    buf = (C.c_ulonglong * 16)()
    L.b200gs_debug_counters(buf, 1)
    r = step()
    torch.cuda.synchronize()
    L.b200gs_debug_counters(buf, 1)
    c = [int(x) for x in buf]
    out["D"] = R.last_num_rendered()
    for name, o in (("fwd", 0), ("bwd", 8)):
        ch, ent, vis, ex, anyc, lanes, anye, few = c[o:o + 8]
        out[name] = {"chunks": ch, "entries": ent, "box_visits": vis, "exact_visits": ex, "visits_with_contrib": anyc,
                     "contrib_lanes": lanes, "visits_alpha_eligible_ignoring_T": anye, "visits_lt8_lanes": few,
                     "exact/box": ex / max(vis, 1), "contrib/box": anyc / max(vis, 1), "lanes_per_contrib_visit": lanes / max(anyc, 1),
                     "box_visits_per_entry": vis / max(ent, 1)}
else:
    R.profile_read()
    R.profile_enable(True)
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = step()
    e1.record()
    torch.cuda.synchronize()
    st = R.profile_read()
    R.profile_enable(False)
    out["ms_per_step"] = e0.elapsed_time(e1) / n
    out["views_per_s"] = V * n / e0.elapsed_time(e1) * 1e3
    out["stages_ms"] = {k: ms / max(calls, 1) for k, (ms, calls) in st.items()}
    out["D"] = R.last_num_rendered()
    out["n_vis"] = int((r > 0).sum())
print(json.dumps(out))

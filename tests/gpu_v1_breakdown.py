"""Ad-hoc: where a single-view fwd+bwd call spends its time (device stages vs host)."""
import sys, time, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from humangaussian_b200 import rasterizer as R
from humangaussian_b200.cameras import sample_orbit_cameras
from humangaussian_b200.renderer import stack_cameras
from humangaussian_b200.scene import synthetic_body
dev = "cuda:0"; P = 300000; HW = 1024; deg = 3
p = synthetic_body(P, sh_degree=deg, seed=0).to(dev)
cams = sample_orbit_cameras(8, HW, HW, seed=1000, device=dev)
vm, pm, cp, tanx, tany = stack_cameras(cams, dev)
with torch.no_grad():
    xyz, op, sh, sc, rot = p.get_xyz, p.get_opacity, p.get_features.contiguous(), p.get_scaling, p.get_rotation
for t in (xyz, op, sh, sc, rot): t.requires_grad_(True)
bg = torch.zeros(3, device=dev)
gw = [torch.randn(1, c, HW, HW, device=dev) for c in (3, 1, 1)]
def one(i, bwd=True):
    c, r, d, a = R.rasterize_views(means3D=xyz, opacities=op, viewmatrices=vm[i:i+1], projmatrices=pm[i:i+1], camposs=cp[i:i+1], tanfovx=tanx[i:i+1],
                                   tanfovy=tany[i:i+1], image_height=HW, image_width=HW, bg=bg, sh_degree=deg, shs=sh, scales=sc, rotations=rot)
    if bwd: torch.autograd.backward([c, d, a], gw)
for _ in range(3):
    for i in range(8): one(i)
torch.cuda.synchronize(); R.profile_read(); R.profile_enable(True)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    for i in range(8): one(i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / (8 * n)
st = R.profile_read(); R.profile_enable(False)
print(f"wall per view {wall*1e3:.3f} ms")
tot = 0
for k, (ms, calls) in st.items():
    print(f"  {k:16s} {ms/calls*1e3:8.1f} us"); tot += ms / calls
print(f"  device stages sum {tot*1e3:.1f} us")
# host-only cost of forward call path (no GPU wait): time enqueue
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(8): one(i, bwd=False)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"fwd-only: host enqueue+sync per view {(t1-t0)/8*1e3:.3f} ms, total {(t2-t0)/8*1e3:.3f} ms")

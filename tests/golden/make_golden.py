"""Generates the committed golden fixtures by IMPORTING THE REFERENCE'S OWN PYTHON (run in the authoring
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Outputs (small .npz files next to this script):
  ref_host_math.npz  -- inputs and outputs of the reference functions the rasteriser path shares math with:
        eval_sh                     gaussiansplatting/utils/sh_utils.py:57-112
        build_scaling_rotation/...  gaussiansplatting/utils/general_utils.py:64-110 (+ gaussian_model.py:27-31)
        geom_transform_points       gaussiansplatting/utils/graphics_utils.py:22-30
        getProjectionMatrix         gaussiansplatting/utils/graphics_utils.py:73-93
        Camera                      gaussiansplatting/scene/cameras.py:17-54
        MiniCam (animation)         gs_renderer.py:853-879   [source-extracted: module imports CUDA-only deps]
  ref_api_surface.json -- keyword names the reference passes to GaussianRasterizationSettings(...) and
        rasterizer(...) (gaussian_renderer/__init__.py:36-49,86-94; gs_renderer.py:951-964,1006-1015), by AST.
  sample_ply_8k.npz -- a seed-0 8192-Gaussian subsample of content/sample.ply (raw PLY columns), the only real
        scene data in the reference; lets tests exercise real anisotropy/opacity statistics off-box.
The reference hard-codes device="cuda"; it is run on CPU here by patching torch factory functions to drop the
device argument and Tensor.cuda to a no-op -- no reference source is modified or copied.
"""
import ast
import importlib.util
import json
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _patch_cuda():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "empty", "tensor", "zeros_like"):
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            if "device" in k:
                k.pop("device")
            return __orig(*a, **k)
        setattr(torch, name, wrap)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    _patch_cuda()
    from gaussiansplatting.utils import general_utils, graphics_utils, sh_utils
    cameras = _load(os.path.join(REF, "gaussiansplatting/scene/cameras.py"), "ref_cameras")
    g = torch.Generator().manual_seed(1234)
    out = {}
    # ---- SH
    P = 257
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    sh = torch.randn(P, 3, 16, generator=g)  # reference layout [..., C, K]
    out["sh_dirs"], out["sh_coeffs_PCK"] = dirs.numpy(), sh.numpy()
    for deg in range(4):
        out[f"sh_eval_deg{deg}"] = sh_utils.eval_sh(deg, sh, dirs).numpy()
    out["rgb2sh_in"] = np.linspace(0, 1, 11, dtype=np.float32)
    out["rgb2sh_out"] = sh_utils.RGB2SH(torch.tensor(out["rgb2sh_in"])).numpy()
    # ---- covariance (gaussian_model.py:27-31: L = build_scaling_rotation(mod*s, q); Sigma = L L^T; strip_symmetric)
    s = torch.exp(torch.randn(P, 3, generator=g) * 0.7 - 4)
    q = torch.randn(P, 4, generator=g)  # build_rotation normalises internally
    mod = 1.0
    L = general_utils.build_scaling_rotation(mod * s, q)
    cov = general_utils.strip_symmetric(L @ L.transpose(1, 2))
    out["cov_scales"], out["cov_quats"], out["cov_packed"] = s.numpy(), q.numpy(), cov.numpy()
    out["rotmat"] = general_utils.build_rotation(q).numpy()
    # ---- projection matrix + cameras
    out["proj_args"] = np.array([[0.01, 100.0, 1.2, 0.9], [0.1, 50.0, 0.6, 0.6], [0.01, 100.0, math.radians(70), math.radians(70)]], np.float64)
    out["proj_mats"] = np.stack([graphics_utils.getProjectionMatrix(*[float(x) for x in a]).numpy() for a in out["proj_args"]])
    from humangaussian_b200.cameras import orbit_c2w
    c2ws, cam_out = [], []
    specs = [(15.0, 0.0, 2.0, 70.0, 256, 256), (-20.0, 135.0, 1.6, 45.0, 512, 384), (5.0, -90.0, 1.9, 55.0, 1024, 1024)]
    for el, az, dist, fovy, H, W in specs:
        c2w = orbit_c2w(el, az, dist)
        cam = cameras.Camera(c2w.clone(), math.radians(fovy), H, W)
        c2ws.append(c2w.numpy())
        cam_out.append(np.concatenate([cam.world_view_transform.numpy().ravel(), cam.full_proj_transform.numpy().ravel(),
                                       cam.camera_center.numpy().ravel(), [cam.FoVx, cam.FoVy]]))
    out["cam_specs"], out["cam_c2w"], out["cam_out"] = np.array(specs, np.float64), np.stack(c2ws), np.stack(cam_out)
    # geom_transform_points with the first camera's full projection
    pts = torch.randn(P, 3, generator=g) * 0.5
    cam = cameras.Camera(orbit_c2w(15.0, 0.0, 2.0), math.radians(70), 256, 256)
    out["gtp_points"] = pts.numpy()
    out["gtp_matrix"] = cam.full_proj_transform.numpy()
    out["gtp_out"] = graphics_utils.geom_transform_points(pts, cam.full_proj_transform).numpy()
    # ---- animation MiniCam: extract the class + its getProjectionMatrix from gs_renderer.py by AST (the module
    #      itself imports diff_gaussian_rasterization / simple_knn / kiui, none of which exist here)
    src = open(os.path.join(REF, "gs_renderer.py")).read()
    tree = ast.parse(src)
    want = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "getProjectionMatrix") or
            (isinstance(n, ast.ClassDef) and n.name == "MiniCam")]
    ns = {"torch": torch, "np": np, "math": math}
    exec(compile(ast.Module(body=want, type_ignores=[]), "gs_renderer_extract", "exec"), ns)
    mc_out = []
    for el, az, dist, fovy, H, W in specs:
        c2w = orbit_c2w(el, az, dist).numpy().astype(np.float32)
        fy = math.radians(fovy)
        fxv = 2 * math.atan(math.tan(fy / 2) * W / H)
        mc = ns["MiniCam"](c2w.copy(), W, H, fy, fxv, 0.01, 100.0)
        mc_out.append(np.concatenate([mc.world_view_transform.numpy().ravel(), mc.full_proj_transform.numpy().ravel(),
                                      mc.camera_center.numpy().ravel(), [fxv, fy]]))
    out["minicam_out"] = np.stack(mc_out)
    np.savez_compressed(os.path.join(HERE, "ref_host_math.npz"), **out)

    # ---- API surface by AST
    def call_kwargs(path, func_names):
        res = {}
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Call):
                f = node.func
                name = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else None)
                if name in func_names and node.keywords:
                    res.setdefault(name, []).append([k.arg for k in node.keywords])
        return res
    surf = {
        "gaussian_renderer": call_kwargs(os.path.join(REF, "gaussiansplatting/gaussian_renderer/__init__.py"),
                                         {"GaussianRasterizationSettings", "GaussianRasterizer", "rasterizer"}),
        "gs_renderer": call_kwargs(os.path.join(REF, "gs_renderer.py"), {"GaussianRasterizationSettings", "GaussianRasterizer", "rasterizer"}),
    }
    json.dump(surf, open(os.path.join(HERE, "ref_api_surface.json"), "w"), indent=1, sort_keys=True)

    # ---- real-scene subsample
    from humangaussian_b200.scene import read_ply
    cols = read_ply(os.path.join(REF, "content/sample.ply"))
    n = len(cols["x"])
    idx = np.sort(np.random.RandomState(0).permutation(n)[:8192])
    np.savez_compressed(os.path.join(HERE, "sample_ply_8k.npz"), n_total=n, **{k: v[idx] for k, v in cols.items()})
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

"""Gaussian scene state on the hot path: PLY codec, activations, synthetic scenes.

Mirrors (device-agnostic; the reference hard-codes "cuda"):
  * PLY schema + training-path loader   gaussiansplatting/scene/gaussian_model.py:187-266
  * animation-path loader (y/z swap)    gs_renderer.py:525-610
  * activations exp / normalize / sigmoid / cat    gaussian_model.py:33-41,95-115
  * RGB2SH                              gaussiansplatting/utils/sh_utils.py:114-118
Synthetic scenes stand in for content/sample.ply on the GPU box (no reference tree there);
their statistics follow SURVEY.md 8(c)/(d).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def RGB2SH(rgb):
    return (rgb - 0.5) / SH_C0


def SH2RGB(sh):
    return sh * SH_C0 + 0.5


# ----------------------------------------------------------------------------------------- PLY
def read_ply(path):
    """Binary-little-endian PLY with float32 vertex properties -> {name: np.float32[N]}."""
    with open(path, "rb") as f:
        names, n, fmt_ok = [], None, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY: missing end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt_ok = tok[1] == "binary_little_endian"
            elif tok[0] == "element":
                if tok[1] != "vertex" and n is not None:
                    raise ValueError("PLY: only a single vertex element is supported")
                n = int(tok[2])
            elif tok[0] == "property":
                if tok[1] not in ("float", "float32"):
                    raise ValueError(f"PLY: unsupported property type {tok[1]}")
                names.append(tok[2])
            elif tok[0] == "end_header":
                break
        if not fmt_ok or n is None:
            raise ValueError("PLY: expected binary_little_endian with a vertex element")
        raw = np.frombuffer(f.read(n * len(names) * 4), dtype="<f4").reshape(n, len(names))
    return {name: np.ascontiguousarray(raw[:, i]) for i, name in enumerate(names)}


def write_ply(path, cols):
    """Inverse of read_ply (property order = dict order), same layout save_ply produces."""
    names = list(cols)
    n = len(cols[names[0]])
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    hdr += "".join("property float %s\n" % k for k in names) + "end_header\n"
    data = np.stack([np.asarray(cols[k], dtype="<f4") for k in names], axis=1)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(data.tobytes())


@dataclass
class GaussianParams:
    """Raw (pre-activation) parameters, laid out as GaussianModel keeps them."""
    xyz: torch.Tensor            # [P,3]
    features_dc: torch.Tensor    # [P,1,3]
    features_rest: torch.Tensor  # [P,K-1,3]
    scaling: torch.Tensor        # [P,3]  (log)
    rotation: torch.Tensor       # [P,4]  (w,x,y,z), not normalised
    opacity: torch.Tensor        # [P,1]  (logit)
    sh_degree: int = 0

    def to(self, device):
        return GaussianParams(*(t.to(device) for t in (self.xyz, self.features_dc, self.features_rest, self.scaling,
                                                       self.rotation, self.opacity)), sh_degree=self.sh_degree)

    @property
    def P(self):
        return self.xyz.shape[0]

    # getters = GaussianModel.get_* (gaussian_model.py:95-115)
    @property
    def get_xyz(self):
        return self.xyz

    @property
    def get_scaling(self):
        return torch.exp(self.scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self.rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self.opacity)

    @property
    def get_features(self):
        return torch.cat((self.features_dc, self.features_rest), dim=1)

    @property
    def active_sh_degree(self):
        return self.sh_degree

    @property
    def max_sh_degree(self):
        return self.sh_degree


def _sorted_cols(cols, prefix, sort=True):
    """Columns whose name starts with `prefix`, in index order (training loader: gaussian_model.py:241-263 sorts by the
    trailing integer) or in FILE order (animation loader: gs_renderer.py:553-576 does not sort)."""
    names = [k for k in cols if k.startswith(prefix)]
    if sort:
        names = sorted(names, key=lambda s: int(s.split("_")[-1]))
    return np.stack([cols[k] for k in names], axis=1) if names else np.zeros((len(cols["x"]), 0), np.float32)


def params_from_ply(path, sh_degree=0, convention="training"):
    """convention="training": gaussian_model.py:225-266 (z-up as stored).
    convention="animation": gs_renderer.py:576-581 (swap y/z of xyz and scales, swap quaternion
    comps 2/3 and negate comp 0)."""
    if convention not in ("training", "animation"):
        raise ValueError(convention)
    c = read_ply(path)
    srt = convention == "training"
    xyz = np.stack([c["x"], c["y"], c["z"]], axis=1)
    fdc = np.stack([c["f_dc_0"], c["f_dc_1"], c["f_dc_2"]], axis=1)[:, None, :]
    K = (sh_degree + 1) ** 2
    rest = _sorted_cols(c, "f_rest_", srt)
    if rest.shape[1] != 3 * K - 3:
        raise ValueError(f"PLY has {rest.shape[1]} f_rest_* columns, sh_degree={sh_degree} needs {3 * K - 3}")
    rest = rest.reshape(len(xyz), 3, K - 1).transpose(0, 2, 1)
    scales, rots = _sorted_cols(c, "scale_", srt), _sorted_cols(c, "rot", srt)
    if convention == "animation":
        xyz = xyz[:, [0, 2, 1]]
        scales = scales[:, [0, 2, 1]]
        rots = rots[:, [0, 1, 3, 2]].copy()
        rots[:, 0] *= -1
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
    return GaussianParams(t(xyz), t(fdc), t(rest), t(scales), t(rots), t(c["opacity"][:, None]), sh_degree)


def params_to_ply(path, p: GaussianParams):
    """save_ply (gaussian_model.py:201-218): x,y,z,nx,ny,nz,f_dc_*,f_rest_*,opacity,scale_*,rot_*."""
    cols = {}
    xyz = p.xyz.detach().cpu().numpy()
    for i, k in enumerate("xyz"):
        cols[k] = xyz[:, i]
    for k in ("nx", "ny", "nz"):
        cols[k] = np.zeros(len(xyz), np.float32)
    fdc = p.features_dc.detach().cpu().transpose(1, 2).flatten(1).numpy()
    rest = p.features_rest.detach().cpu().transpose(1, 2).flatten(1).numpy()
    for i in range(fdc.shape[1]):
        cols[f"f_dc_{i}"] = fdc[:, i]
    for i in range(rest.shape[1]):
        cols[f"f_rest_{i}"] = rest[:, i]
    cols["opacity"] = p.opacity.detach().cpu().numpy()[:, 0]
    for i in range(3):
        cols[f"scale_{i}"] = p.scaling.detach().cpu().numpy()[:, i]
    for i in range(4):
        cols[f"rot_{i}"] = p.rotation.detach().cpu().numpy()[:, i]
    write_ply(path, cols)


# ------------------------------------------------------------------------------ synthetic scenes
_BODY = [  # (centre xyz, half-axes xyz, weight) -- a z-up standing figure ~1.6 tall, like sample.ply's extent
    ((0.0, 0.0, 0.70), (0.09, 0.08, 0.11), 0.08),    # head
    ((0.0, 0.0, 0.30), (0.17, 0.10, 0.27), 0.30),    # torso
    ((0.0, 0.0, -0.02), (0.16, 0.10, 0.10), 0.08),   # hips
    ((-0.33, 0.0, 0.42), (0.22, 0.045, 0.045), 0.07),  # arms (T-pose-ish)
    ((0.33, 0.0, 0.42), (0.22, 0.045, 0.045), 0.07),
    ((-0.09, 0.0, -0.30), (0.07, 0.07, 0.24), 0.12),  # thighs
    ((0.09, 0.0, -0.30), (0.07, 0.07, 0.24), 0.12),
    ((-0.10, 0.0, -0.65), (0.05, 0.05, 0.17), 0.08),  # shins
    ((0.10, 0.0, -0.65), (0.05, 0.05, 0.17), 0.08),
]


def synthetic_body(P, sh_degree=0, seed=0, surface=True):
    """Human-proportioned Gaussian cloud with sample.ply-like statistics:
    exp(scale) log-normal, median ~2.8e-3 with a long anisotropic tail capped at 0.067; sigmoid(opacity)
    mean ~0.10 with ~0.6 % above 0.5; non-unit quaternions (norm 0.6-1.5); f_dc std ~1.1;
    f_rest ~ N(0, 0.1^2) when sh_degree>0 (so higher bands are exercised)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.tensor([b[2] for b in _BODY])
    part = torch.multinomial(w / w.sum(), P, replacement=True, generator=g)
    ctr = torch.tensor([b[0] for b in _BODY])[part]
    ax = torch.tensor([b[1] for b in _BODY])[part]
    d = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    rad = torch.ones(P, 1) if surface else torch.rand(P, 1, generator=g) ** (1 / 3)
    shell = 1.0 + 0.03 * torch.randn(P, 1, generator=g)
    xyz = ctr + d * ax * rad * shell
    # scales: log-normal, anisotropic
    base = torch.randn(P, 1, generator=g) * 0.55 + np.log(2.8e-3)
    scaling = (base + 0.45 * torch.randn(P, 3, generator=g)).clamp(max=float(np.log(0.067)))
    rot = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1) * (0.6 + 0.9 * torch.rand(P, 1, generator=g))
    # opacity: beta-like, mean 0.10, small heavy tail
    op = (0.02 + 0.55 * torch.rand(P, 1, generator=g) ** 3.2).clamp(1e-4, 0.999)
    heavy = torch.rand(P, 1, generator=g) < 0.006
    op = torch.where(heavy, 0.5 + 0.49 * torch.rand(P, 1, generator=g), op)
    opacity = torch.log(op / (1 - op))
    K = (sh_degree + 1) ** 2
    fdc = (torch.randn(P, 1, 3, generator=g) * 1.1)
    rest = torch.randn(P, K - 1, 3, generator=g) * 0.1
    return GaussianParams(xyz.float(), fdc.float(), rest.float(), scaling.float(), rot.float(), opacity.float(), sh_degree)


SAMPLE_COLUMNS = ("x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
                  "rot_0", "rot_1", "rot_2", "rot_3")
SAMPLE_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sample_ply_full.npz")


def sample_ply_scene(n=None, sh_degree=0, seed=0, convention="training", path=None):
    """The reference's own scene, content/sample.ply (531 327 Gaussians, SH degree 0), from the committed column pack
    tests/golden/sample_ply_full.npz (tests/golden/make_sample_scene.py) -- BASELINE configs 2, 4, 5.
    n: keep a seed-`seed` subsample of n Gaussians (n = 300 000, seed 0); sh_degree > 0 widens with
    f_rest ~ N(0, 0.1^2) (seed `seed`), as 8d prescribes.  Raw (pre-activation) parameters, like load_ply
    (gaussiansplatting/scene/gaussian_model.py:225-266); convention="animation" applies gs_renderer.py:576-581."""
    z = np.load(path or SAMPLE_NPZ)
    d = z["data"]
    col = {k: d[:, i] for i, k in enumerate(SAMPLE_COLUMNS)}
    xyz = np.stack([col["x"], col["y"], col["z"]], 1)
    fdc = np.stack([col["f_dc_0"], col["f_dc_1"], col["f_dc_2"]], 1)[:, None, :]
    scales = np.stack([col["scale_0"], col["scale_1"], col["scale_2"]], 1)
    rots = np.stack([col["rot_0"], col["rot_1"], col["rot_2"], col["rot_3"]], 1)
    if convention == "animation":
        xyz, scales = xyz[:, [0, 2, 1]], scales[:, [0, 2, 1]]
        rots = rots[:, [0, 1, 3, 2]].copy()
        rots[:, 0] *= -1
    elif convention != "training":
        raise ValueError(convention)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
    p = GaussianParams(t(xyz), t(fdc), torch.zeros(len(xyz), 0, 3), t(scales), t(rots), t(col["opacity"][:, None]), 0)
    if n is not None and n < p.P:
        p = subsample(p, n, seed)
    if sh_degree > 0:
        p = with_sh_degree(p, sh_degree, seed)
    return p


def subsample(p: GaussianParams, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(p.P, generator=g)[:n].sort().values
    return GaussianParams(p.xyz[idx], p.features_dc[idx], p.features_rest[idx], p.scaling[idx], p.rotation[idx],
                          p.opacity[idx], p.sh_degree)


def with_sh_degree(p: GaussianParams, sh_degree, seed=0, std=0.1):
    """Widen to a higher SH degree with N(0,std^2) higher-band coefficients."""
    g = torch.Generator().manual_seed(seed)
    K = (sh_degree + 1) ** 2
    rest = torch.randn(p.P, K - 1, 3, generator=g) * std
    have = p.features_rest.shape[1]
    rest[:, :have] = p.features_rest
    return GaussianParams(p.xyz, p.features_dc, rest, p.scaling, p.rotation, p.opacity, sh_degree)


def params_from_pcd(points, colors, sh_degree=0, device="cuda"):
    """GaussianModel.create_from_pcd (gaussiansplatting/scene/gaussian_model.py:124-147): DC feature = RGB2SH(colour),
    higher SH bands 0, isotropic log-scale = log sqrt(clamp_min(mean squared distance to the 3 nearest neighbours,
    1e-7)), identity quaternion, opacity = inverse_sigmoid(0.1).  The neighbour distances come from this repo's
    distCUDA2 kernel (the reference calls simple_knn._C.distCUDA2)."""
    from .rasterizer import distCUDA2
    pts = torch.as_tensor(points, dtype=torch.float32, device=device)
    col = torch.as_tensor(colors, dtype=torch.float32, device=device)
    P, K = pts.shape[0], (sh_degree + 1) ** 2
    dist2 = torch.clamp_min(distCUDA2(pts), 0.0000001)
    scaling = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
    rotation = torch.zeros(P, 4, device=device)
    rotation[:, 0] = 1
    opacity = torch.full((P, 1), float(np.log(0.1 / 0.9)), dtype=torch.float32, device=device)
    return GaussianParams(pts, RGB2SH(col)[:, None, :].contiguous(), torch.zeros(P, K - 1, 3, device=device), scaling, rotation, opacity,
                          sh_degree)

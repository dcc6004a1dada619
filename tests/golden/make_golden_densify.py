"""Golden vectors for the densify / prune row, produced by RUNNING THE REFERENCE'S OWN
`GaussianModel` (gaussiansplatting/scene/gaussian_model.py:268-437) on the CPU of the authoring container.

    python tests/golden/make_golden_densify.py        ->  tests/golden/ref_densify.npz

The reference hard-codes device="cuda" and imports `plyfile`; it is run here unmodified by patching torch's
factory functions to ignore the device keyword and stubbing the `plyfile` module (unused on this path).
`torch.normal(mean, std)` is routed through `randn * std + mean` (ATen's own definition of the tensor-std overload)
so that the standard-normal draws can be recorded: the device kernel takes them as an input.
No reference source is copied into the repository; /root/reference is not needed to run the tests.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

NOISE = []


def _patch():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "empty", "tensor", "zeros_like"):
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            k.pop("device", None)
            return __orig(*a, **k)
        setattr(torch, name, wrap)

    def normal(mean, std):
        z = torch.randn(std.shape)
        NOISE.append(z.clone())
        return z * std + mean
    torch.normal = normal
    torch.cuda.empty_cache = lambda: None
    m = types.ModuleType("plyfile")
    m.PlyData = m.PlyElement = object
    sys.modules["plyfile"] = m


class Args:
    percent_dense = 0.01
    position_lr_init = 0.00016
    position_lr_final = 0.0000016
    position_lr_delay_mult = 0.01
    position_lr_max_steps = 30000
    feature_lr = 0.0025
    opacity_lr = 0.05
    scaling_lr = 0.005
    rotation_lr = 0.001


GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def make_model(GaussianModel, P, K, seed):
    g = torch.Generator().manual_seed(seed)
    m = GaussianModel(3)
    nn = torch.nn
    m._xyz = nn.Parameter(torch.randn(P, 3, generator=g) * 0.4)
    m._features_dc = nn.Parameter(torch.randn(P, 1, 3, generator=g))
    m._features_rest = nn.Parameter(torch.randn(P, K - 1, 3, generator=g) * 0.1)
    m._opacity = nn.Parameter(torch.randn(P, 1, generator=g) * 2.0 - 1.5)      # sigmoid spread around the 0.05 threshold
    m._scaling = nn.Parameter(torch.randn(P, 3, generator=g) * 0.9 - 5.0)      # exp spread around 0.01*extent and 0.1*extent
    m._rotation = nn.Parameter(torch.randn(P, 4, generator=g))
    m.max_radii2D = torch.zeros(P)
    m.spatial_lr_scale = 1.0
    m.training_setup(Args())
    # two Adam steps with random gradients so that exp_avg / exp_avg_sq are populated
    for _ in range(2):
        for grp in m.optimizer.param_groups:
            p = grp["params"][0]
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        m.optimizer.step()
    return m, g


def snapshot(m, prefix, out):
    tens = dict(xyz=m._xyz, f_dc=m._features_dc, f_rest=m._features_rest, opacity=m._opacity, scaling=m._scaling, rotation=m._rotation)
    for k, v in tens.items():
        out[f"{prefix}_{k}"] = v.detach().numpy().copy()
    for grp in m.optimizer.param_groups:
        st = m.optimizer.state[grp["params"][0]]
        out[f"{prefix}_m_{grp['name']}"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_v_{grp['name']}"] = st["exp_avg_sq"].numpy().copy()
    out[f"{prefix}_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{prefix}_denom"] = m.denom.numpy().copy()
    out[f"{prefix}_max_radii2D"] = m.max_radii2D.numpy().copy()


def feed_stats(m, g, V, rounds, out, tag):
    """GaussianDreamer.py:385-391: sum the per-view viewspace gradients, max_radii2D update, add_densification_stats."""
    P = m._xyz.shape[0]
    grads, radii = [], []
    for r in range(rounds):
        vg = torch.randn(V, P, 3, generator=g) * 5e-5
        vg[:, torch.rand(P, generator=g) < 0.3] *= 20.0           # some Gaussians well above max_grad
        rad = (torch.rand(V, P, generator=g) * 40).floor().to(torch.int32)
        rad[torch.rand(V, P, generator=g) < 0.35] = 0             # culled in that view
        rad[:, torch.rand(P, generator=g) < 0.1] = 0              # never visible this round -> denom stays (0/0 -> nan)
        grads.append(vg.numpy().copy())
        radii.append(rad.numpy().copy())
        acc = torch.zeros_like(vg[0])
        for v in range(V):
            acc = acc + vg[v]
        rmax = rad.max(0).values
        vis = rmax > 0
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], rmax[vis].float())
        m.add_densification_stats(acc, vis)
    out[f"{tag}_view_grads"] = np.stack(grads)
    out[f"{tag}_view_radii"] = np.stack(radii)


def main():
    _patch()
    from gaussiansplatting.scene.gaussian_model import GaussianModel
    out = {}
    cases = [("dp_noscreen", 400, 16, 11, dict(max_grad=0.0002, min_opacity=0.05, extent=1.0, max_screen_size=None)),
             ("dp_screen", 350, 16, 12, dict(max_grad=0.0002, min_opacity=0.05, extent=0.5, max_screen_size=20)),
             ("dp_deg0", 500, 1, 13, dict(max_grad=0.0004, min_opacity=0.02, extent=2.0, max_screen_size=5))]
    for tag, P, K, seed, kw in cases:
        torch.manual_seed(seed)
        m, g = make_model(GaussianModel, P, K, seed)
        feed_stats(m, g, V=4, rounds=3, out=out, tag=tag)
        snapshot(m, tag + "_in", out)
        NOISE.clear()
        m.densify_and_prune(kw["max_grad"], kw["min_opacity"], kw["extent"], kw["max_screen_size"])
        out[tag + "_noise"] = NOISE[0].numpy().copy()
        snapshot(m, tag + "_out", out)
        out[tag + "_args"] = np.array([kw["max_grad"], kw["min_opacity"], kw["extent"],
                                       -1.0 if kw["max_screen_size"] is None else kw["max_screen_size"], Args.percent_dense], np.float64)
        print(tag, P, "->", m._xyz.shape[0], "split parents", NOISE[0].shape[0] // 2)
    # prune_only (gaussian_model.py:423-430): stats arrays are gathered, not reset
    tag, P, K, seed = "po", 400, 16, 21
    torch.manual_seed(seed)
    m, g = make_model(GaussianModel, P, K, seed)
    feed_stats(m, g, V=3, rounds=2, out=out, tag=tag)
    snapshot(m, tag + "_in", out)
    m.prune_only(min_opacity=0.05, size_thresh=0.02)
    snapshot(m, tag + "_out", out)
    out[tag + "_args"] = np.array([0.05, 0.02], np.float64)
    print(tag, P, "->", m._xyz.shape[0])
    np.savez_compressed(os.path.join(HERE, "ref_densify.npz"), **out)
    print("wrote ref_densify.npz", os.path.getsize(os.path.join(HERE, "ref_densify.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Densification / pruning on the device -- the step on the other side of backward.

Mirrors the reference's `GaussianModel` surgery (gaussiansplatting/scene/gaussian_model.py:268-437) and the
bookkeeping that feeds it (threestudio/systems/GaussianDreamer.py:378-408), with the same names and argument meaning:

    stats = DensifyStats.zeros(P, device)
    add_densification_stats(stats, dL_dmeans2D[V,P,3], radii[V,P])          # after every backward
    params, moments, stats = densify_and_prune(params, moments, stats, max_grad, min_opacity, extent, max_screen_size,
                                               percent_dense)               # every densify_prune_interval steps
    params, moments, stats = prune_only(params, moments, stats, min_opacity, size_thresh)
    install_into_optimizer(optimizer, params, moments)                      # what cat_tensors_to_optimizer /
                                                                            # _prune_optimizer do to torch.optim.Adam

`params` is a dict of the RAW optimiser tensors by the reference's group names (xyz, f_dc, f_rest, opacity, scaling,
rotation); `moments` maps the same names to (exp_avg, exp_avg_sq) or is None.  Where the reference runs ~30 PyTorch ops
with a host sync per boolean-mask gather and concatenates every tensor twice, this plans all destinations with one
prefix-sum pass and moves each surviving row once (csrc/densify.cu); one host read-back (the five counts) per call.
There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from .rasterizer import _check, _ptr, _stream, load_library

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
ROLE_COPY, ROLE_XYZ, ROLE_SCALING, ROLE_MOMENT = 0, 1, 2, 3


class _Cfg(C.Structure):  # == b200gs_densify_cfg
    _fields_ = [("mode", C.c_int32), ("n_split", C.c_int32), ("use_screen", C.c_int32), ("max_grad", C.c_float),
                ("min_opacity", C.c_float), ("percent_dense_x_extent", C.c_float), ("max_screen_size", C.c_float),
                ("big_ws_thresh", C.c_float)]


@dataclass
class DensifyStats:
    """xyz_gradient_accum [P,1], denom [P,1], max_radii2D [P] (gaussian_model.py:151-152, 358-360)."""
    xyz_gradient_accum: torch.Tensor
    denom: torch.Tensor
    max_radii2D: torch.Tensor

    @staticmethod
    def zeros(P: int, device) -> "DensifyStats":
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        return DensifyStats(z(P, 1), z(P, 1), z(P))


def _need_cuda(t: torch.Tensor):
    if t.device.type != "cuda":
        raise RuntimeError("b200gs: tensors must live on a CUDA device (there is no CPU path)")


def add_densification_stats(stats: DensifyStats, viewspace_grads: torch.Tensor, radii: torch.Tensor,
                            update_mask: Optional[torch.Tensor] = None) -> None:
    """One optimiser step of bookkeeping for a batch of views, in place.
    viewspace_grads [V,P,3] (or [P,3]): the means2D gradients of the step's views; radii [V,P] (or [P]) int32.
    Equals GaussianDreamer.py:385-391: sum the view gradients, visibility = max radius > 0, max_radii2D update,
    then GaussianModel.add_densification_stats (gaussian_model.py:433-437).
    update_mask [P] bool/uint8 (optional) is ANDed with the visibility: the reference removes near-hand points from
    visibility_filter before this bookkeeping when disable_hand_densification is set (GaussianDreamer.py:288-297)."""
    L = load_library()
    g = viewspace_grads.detach()
    if g.dim() == 2:
        g = g[None]
    r = radii if radii.dim() == 2 else radii[None]
    _need_cuda(g)
    g = g.to(torch.float32).contiguous()
    r = r.to(device=g.device, dtype=torch.int32).contiguous()
    V, P = g.shape[0], g.shape[1]
    if r.shape != (V, P) or stats.max_radii2D.shape[0] != P:
        raise ValueError("b200gs: add_densification_stats shape mismatch")
    for t in (stats.xyz_gradient_accum, stats.denom, stats.max_radii2D):
        if not (t.is_contiguous() and t.dtype == torch.float32 and t.device == g.device):
            raise ValueError("b200gs: statistics must be contiguous fp32 tensors on the gradients' device")
    m = None
    if update_mask is not None:
        m = update_mask.to(device=g.device, dtype=torch.uint8).contiguous()
        if m.shape != (P,):
            raise ValueError("b200gs: update_mask must have shape [P]")
    with torch.cuda.device(g.device):
        _check(L.b200gs_densify_stats(P, V, _ptr(g), _ptr(r), _ptr(m), _ptr(stats.xyz_gradient_accum), _ptr(stats.denom),
                                      _ptr(stats.max_radii2D), _stream(g.device)), "densify_stats")


def _run(cfg: _Cfg, params: Dict[str, torch.Tensor], moments, stats: DensifyStats, noise, generator):
    L = load_library()
    xyz = params["xyz"]
    _need_cuda(xyz)
    dev, P = xyz.device, xyz.shape[0]
    src = {k: params[k].detach().to(torch.float32).contiguous() for k in GROUPS}
    if P == 0:
        return ({k: v.clone() for k, v in src.items()}, None if moments is None else {k: (m.clone(), v.clone()) for k, (m, v) in moments.items()},
                DensifyStats.zeros(0, dev), torch.zeros(5, dtype=torch.int32))
    n_split = cfg.n_split if cfg.mode == 0 else 0
    with torch.cuda.device(dev):
        plan = torch.empty(4, P, dtype=torch.int32, device=dev)
        counts = torch.empty(5, dtype=torch.int32, device=dev)
        scratch = torch.empty(L.b200gs_densify_scratch_bytes(P), dtype=torch.uint8, device=dev)
        st = _stream(dev)
        _check(L.b200gs_densify_plan(P, C.byref(cfg), _ptr(stats.xyz_gradient_accum), _ptr(stats.denom), _ptr(src["opacity"]),
                                     _ptr(src["scaling"]), _ptr(plan), _ptr(counts), _ptr(scratch), scratch.numel(), st), "densify_plan")
        host = counts.cpu()  # the one host sync of the call: P_new sizes every destination
        n_keep, n_clone, S_sel, S_kept, P_new = (int(x) for x in host)
        ch = (C.c_int32 * 5)(n_keep, n_clone, S_sel, S_kept, P_new)
        if cfg.mode == 0 and noise is None:
            # torch.normal(mean=0, std=stds) == randn * stds (gaussian_model.py:369-371)
            noise = torch.randn(n_split * S_sel, 3, dtype=torch.float32, device=dev, generator=generator)
        if noise is not None:
            noise = noise.to(device=dev, dtype=torch.float32).contiguous()
            if noise.numel() != n_split * S_sel * 3:
                raise ValueError(f"b200gs: noise must hold {n_split * S_sel} rows of 3 standard-normal draws")

        def move(role, t):
            rf = t.numel() // P
            out = torch.empty((P_new,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
            if rf == 0:  # e.g. f_rest [P,0,3] at SH degree 0
                return out
            _check(L.b200gs_densify_move(role, P, rf, _ptr(plan), ch, n_split, _ptr(t), _ptr(out), _ptr(src["rotation"]),
                                         _ptr(src["scaling"]), None if noise is None else _ptr(noise), st), "densify_move")
            return out
        roles = {"xyz": ROLE_XYZ, "scaling": ROLE_SCALING}
        new_params = {k: move(roles.get(k, ROLE_COPY), src[k]) for k in GROUPS}
        new_moments = None
        if moments is not None:
            new_moments = {k: tuple(move(ROLE_MOMENT, m.detach().to(torch.float32).contiguous()) for m in moments[k]) for k in GROUPS}
        if cfg.mode == 0:   # densification_postfix resets all three statistics (gaussian_model.py:358-360)
            new_stats = DensifyStats.zeros(P_new, dev)
        else:               # prune_points gathers them (gaussian_model.py:310-315)
            new_stats = DensifyStats(move(ROLE_COPY, stats.xyz_gradient_accum), move(ROLE_COPY, stats.denom), move(ROLE_COPY, stats.max_radii2D))
    return new_params, new_moments, new_stats, host


def densify_and_prune(params, moments, stats: DensifyStats, max_grad: float, min_opacity: float, extent: float,
                      max_screen_size: Optional[float], percent_dense: float = 0.01, N: int = 2,
                      noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None, return_counts: bool = False):
    """GaussianModel.densify_and_prune (gaussian_model.py:402-415): clone small high-gradient Gaussians, split large ones
    into N children, drop the split parents, then prune by opacity (and, with max_screen_size, by world size).
    `noise` [N*S,3]: optional standard-normal draws for the children (default: torch.randn on the device)."""
    cfg = _Cfg(0, int(N), 1 if max_screen_size else 0, float(max_grad), float(min_opacity), float(percent_dense * extent),
               float(max_screen_size or 0.0), float(0.1 * extent))
    p, m, s, counts = _run(cfg, params, moments, stats, noise, generator)
    return (p, m, s, counts) if return_counts else (p, m, s)


def prune_only(params, moments, stats: DensifyStats, min_opacity: float = 0.005, size_thresh: float = 0.01, return_counts: bool = False):
    """GaussianModel.prune_only (gaussian_model.py:423-430): drop low-opacity or oversized Gaussians; statistics are kept."""
    cfg = _Cfg(1, 0, 0, 0.0, float(min_opacity), 0.0, 0.0, float(size_thresh))
    p, m, s, counts = _run(cfg, params, moments, stats, None, None)
    return (p, m, s, counts) if return_counts else (p, m, s)


def moments_from_optimizer(optimizer: torch.optim.Optimizer) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """(exp_avg, exp_avg_sq) per named param group of a reference-style Adam (gaussian_model.py:156-165)."""
    out = {}
    for group in optimizer.param_groups:
        p = group["params"][0]
        st = optimizer.state.get(p)
        if st:
            out[group["name"]] = (st["exp_avg"], st["exp_avg_sq"])
        else:
            out[group["name"]] = (torch.zeros_like(p), torch.zeros_like(p))
    return out


def install_into_optimizer(optimizer: torch.optim.Optimizer, params, moments) -> Dict[str, torch.nn.Parameter]:
    """Swap the resized tensors and Adam moments into the optimiser's single-parameter groups, as
    _prune_optimizer / cat_tensors_to_optimizer do (gaussian_model.py:284-341).  Returns the new nn.Parameters by name."""
    new = {}
    for group in optimizer.param_groups:
        name = group["name"]
        old = group["params"][0]
        st = optimizer.state.pop(old, None)
        par = torch.nn.Parameter(params[name].requires_grad_(True))
        group["params"][0] = par
        if st is not None and moments is not None:
            st["exp_avg"], st["exp_avg_sq"] = moments[name]
            optimizer.state[par] = st
        new[name] = par
    return new

"""CPU oracle for the densify / prune row.  TEST INFRASTRUCTURE ONLY -- never imported from
humangaussian_b200/.

numpy (float32) restatement of the reference's point-set surgery, step by step in the reference's own order
(concatenate, then mask), so that it is structurally independent of the device implementation (which plans
destinations with prefix sums and moves every row once):

    add_densification_stats           gaussiansplatting/scene/gaussian_model.py:433-437
    view-gradient sum / max_radii2D   threestudio/systems/GaussianDreamer.py:385-391
    densify_and_clone                 gaussian_model.py:386-400
    densify_and_split                 gaussian_model.py:362-384   (build_rotation: utils/general_utils.py:78-99)
    densification_postfix             gaussian_model.py:343-360   (statistics reset to zeros)
    prune_points / _prune_optimizer   gaussian_model.py:284-315   (Adam moments gathered)
    cat_tensors_to_optimizer          gaussian_model.py:317-341   (Adam moments of new rows = 0)
    densify_and_prune                 gaussian_model.py:402-415
    prune_only                        gaussian_model.py:423-430

PINNED: tests/test_oracle_densify.py checks this file against tests/golden/ref_densify.npz, produced by running the
reference's own GaussianModel on CPU (tests/golden/make_golden_densify.py).

State is a dict: params {xyz[P,3], f_dc[P,1,3], f_rest[P,K-1,3], opacity[P,1], scaling[P,3], rotation[P,4]} (raw,
pre-activation, as the optimizer holds them), Adam moments m_*/v_* of the same shapes, and the statistics
accum[P,1], denom[P,1], max_radii2D[P].
"""
from __future__ import annotations

import numpy as np

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
F = np.float32


def add_densification_stats(state, view_grads, view_radii):
    """One optimiser step's bookkeeping for a batch of V views (GaussianDreamer.py:385-391 + gaussian_model.py:433-437)."""
    vg = np.asarray(view_grads, F)
    acc = np.zeros_like(vg[0])
    for v in range(vg.shape[0]):
        acc = acc + vg[v]
    rmax = np.asarray(view_radii).max(0)
    vis = rmax > 0
    state["max_radii2D"][vis] = np.maximum(state["max_radii2D"][vis], rmax[vis].astype(F))
    nrm = np.sqrt(acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1]).astype(F)
    state["accum"][vis, 0] += nrm[vis]
    state["denom"][vis, 0] += F(1)
    return state


def _sigmoid(x):
    return (F(1) / (F(1) + np.exp(-x.astype(F)))).astype(F)


def _rotmat(q):
    n = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    q = q / n[:, None]
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.zeros((len(q), 3, 3), F)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _append(state, new, reset_stats=True):
    for g in GROUPS:
        state[g] = np.concatenate([state[g], new[g]], 0)
        for mom in ("m_", "v_"):
            state[mom + g] = np.concatenate([state[mom + g], np.zeros_like(new[g])], 0)
    P = state["xyz"].shape[0]
    state["accum"] = np.zeros((P, 1), F)
    state["denom"] = np.zeros((P, 1), F)
    state["max_radii2D"] = np.zeros(P, F)


def _prune(state, mask):
    keep = ~mask
    for g in GROUPS:
        for pre in ("", "m_", "v_"):
            state[pre + g] = state[pre + g][keep]
    for k in ("accum", "denom", "max_radii2D"):
        state[k] = state[k][keep]


def densify_and_prune(state, max_grad, min_opacity, extent, max_screen_size, percent_dense, noise, N=2):
    """`noise` = the standard-normal draws [N*S, 3] behind torch.normal(mean=0, std=stds) (S = split parents)."""
    s = {k: np.array(v, copy=True) for k, v in state.items()}
    with np.errstate(divide="ignore", invalid="ignore"):
        grads = (s["accum"] / s["denom"]).astype(F)
    grads[np.isnan(grads)] = 0
    thr = F(percent_dense * extent)
    # clone
    scale_max = np.exp(s["scaling"]).astype(F).max(1)
    sel = (np.abs(grads[:, 0]) >= F(max_grad)) & (scale_max <= thr)
    _append(s, {g: s[g][sel] for g in GROUPS})
    # split
    n_init = s["xyz"].shape[0]
    padded = np.zeros(n_init, F)
    padded[:grads.shape[0]] = grads[:, 0]
    act = np.exp(s["scaling"]).astype(F)
    sel = (padded >= F(max_grad)) & (act.max(1) > thr)
    S = int(sel.sum())
    noise = np.asarray(noise, F).reshape(N * S, 3)
    stds = np.tile(act[sel], (N, 1))
    samples = (noise * stds).astype(F)
    R = np.tile(_rotmat(s["rotation"][sel].astype(F)), (N, 1, 1))
    new_xyz = (np.einsum("nij,nj->ni", R.astype(np.float64), samples.astype(np.float64)).astype(F)
               + np.tile(s["xyz"][sel], (N, 1))).astype(F)
    new = {g: np.tile(s[g][sel], (N,) + (1,) * (s[g].ndim - 1)) for g in GROUPS}
    new["xyz"] = new_xyz
    new["scaling"] = np.log((np.tile(act[sel], (N, 1)) / F(0.8 * N)).astype(F)).astype(F)
    _append(s, new)
    _prune(s, np.concatenate([sel, np.zeros(N * S, bool)]))
    # prune
    mask = _sigmoid(s["opacity"][:, 0]) < F(min_opacity)
    if max_screen_size:
        big_vs = s["max_radii2D"] > F(max_screen_size)
        big_ws = np.exp(s["scaling"]).astype(F).max(1) > F(0.1 * extent)
        mask = mask | big_vs | big_ws
    _prune(s, mask)
    return s


def prune_only(state, min_opacity=0.005, size_thresh=0.01):
    s = {k: np.array(v, copy=True) for k, v in state.items()}
    mask = (_sigmoid(s["opacity"][:, 0]) < F(min_opacity)) | (np.exp(s["scaling"]).astype(F).max(1) > F(size_thresh))
    _prune(s, mask)
    return s

"""Pins oracle/densify_ref.py against golden vectors produced by the reference's own GaussianModel
(tests/golden/make_golden_densify.py; gaussiansplatting/scene/gaussian_model.py:268-437)."""
import os

import numpy as np
import pytest

from oracle import densify_ref as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_densify.npz"))
KEYS = [pre + g for g in O.GROUPS for pre in ("", "m_", "v_")] + ["accum", "denom", "max_radii2D"]


def load_state(tag):
    return {k: GOLD[f"{tag}_{k}"].copy() for k in KEYS}


def assert_state_equal(got, want, rtol=2e-6, atol=1e-7):
    for k in KEYS:
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        np.testing.assert_allclose(got[k], want[k], rtol=rtol, atol=atol, err_msg=k)


@pytest.mark.parametrize("tag,rounds", [("dp_noscreen", 3), ("dp_screen", 3), ("dp_deg0", 3), ("po", 2)])
def test_stats_match_reference(tag, rounds):
    want = load_state(tag + "_in")
    P = want["xyz"].shape[0]
    st = dict(accum=np.zeros((P, 1), np.float32), denom=np.zeros((P, 1), np.float32), max_radii2D=np.zeros(P, np.float32))
    for r in range(rounds):
        O.add_densification_stats(st, GOLD[f"{tag}_view_grads"][r], GOLD[f"{tag}_view_radii"][r])
    np.testing.assert_array_equal(st["denom"], want["denom"])
    np.testing.assert_array_equal(st["max_radii2D"], want["max_radii2D"])
    np.testing.assert_allclose(st["accum"], want["accum"], rtol=1e-6, atol=0)
    assert (want["denom"] == 0).any()  # the 0/0 -> nan -> 0 path is exercised


@pytest.mark.parametrize("tag", ["dp_noscreen", "dp_screen", "dp_deg0"])
def test_densify_and_prune_matches_reference(tag):
    max_grad, min_opacity, extent, screen, pd = GOLD[tag + "_args"]
    got = O.densify_and_prune(load_state(tag + "_in"), max_grad, min_opacity, extent, None if screen < 0 else screen, pd,
                              GOLD[tag + "_noise"])
    want = load_state(tag + "_out")
    assert_state_equal(got, want)
    assert not got["accum"].any() and not got["denom"].any() and not got["max_radii2D"].any()


def test_prune_only_matches_reference():
    min_opacity, size_thresh = GOLD["po_args"]
    got = O.prune_only(load_state("po_in"), min_opacity, size_thresh)
    want = load_state("po_out")
    assert_state_equal(got, want, rtol=0, atol=0)
    assert got["denom"].any()  # statistics are gathered, not reset, on this path

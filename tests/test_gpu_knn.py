"""GPU: distCUDA2 (mean squared distance to the 3 nearest neighbours) against an exact CPU kNN (scipy cKDTree, float64)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _exact(points):
    from scipy.spatial import cKDTree
    p = points.astype(np.float64)
    d, _ = cKDTree(p).query(p, k=4)
    return (d[:, 1:] ** 2).mean(1)


@pytest.mark.parametrize("P,kind", [(4, "uniform"), (257, "uniform"), (5000, "uniform"), (100_000, "body"), (20_000, "clustered")])
def test_dist2_matches_exact_knn(P, kind):
    from simple_knn._C import distCUDA2  # the module name the reference imports
    rng = np.random.RandomState(P)
    if kind == "uniform":
        pts = rng.rand(P, 3).astype(np.float32)
    elif kind == "clustered":
        pts = (rng.randn(P, 3) * 0.01 + rng.randint(0, 5, (P, 1))).astype(np.float32)
    else:
        from humangaussian_b200.scene import synthetic_body
        pts = synthetic_body(P, seed=1).xyz.numpy()
    got = distCUDA2(torch.tensor(pts, device=DEV)).cpu().numpy()
    want = _exact(pts)
    assert got.shape == (P,) and np.isfinite(got).all()
    assert np.allclose(got, want, rtol=2e-4, atol=1e-10), float(np.abs(got / np.maximum(want, 1e-30) - 1).max())


def test_duplicate_points_give_zero():
    from simple_knn._C import distCUDA2
    pts = np.tile(np.array([[0.1, 0.2, 0.3]], np.float32), (64, 1))
    assert float(distCUDA2(torch.tensor(pts, device=DEV)).abs().max()) == 0.0


def test_params_from_pcd_matches_reference_init():
    """create_from_pcd (gaussian_model.py:124-147): scales from the 3-NN distances, identity rotations, opacity 0.1."""
    from humangaussian_b200.scene import params_from_pcd, synthetic_body
    pts = synthetic_body(20_000, seed=2).xyz.numpy()
    col = np.random.RandomState(0).rand(20_000, 3).astype(np.float32)
    p = params_from_pcd(pts, col, sh_degree=1, device=DEV)
    want_scale = np.log(np.sqrt(np.maximum(_exact(pts), 1e-7)))
    assert np.allclose(p.scaling.cpu().numpy(), want_scale[:, None].repeat(3, 1), atol=2e-4)
    assert torch.allclose(p.get_opacity, torch.full_like(p.get_opacity, 0.1), atol=1e-6)
    assert torch.equal(p.get_rotation, torch.tensor([1.0, 0, 0, 0], device=DEV).expand(20_000, 4))
    assert np.allclose(p.features_dc.cpu().numpy()[:, 0], (col - 0.5) / 0.28209479177387814, atol=1e-6) and p.features_rest.shape == (20_000, 3, 3)

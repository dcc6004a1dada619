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

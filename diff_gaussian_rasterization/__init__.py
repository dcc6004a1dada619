"""Drop-in module name.  HumanGaussian imports the rasteriser as
``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``
(gaussiansplatting/gaussian_renderer/__init__.py:14, gs_renderer.py:10-13).  With this repo on
sys.path that import resolves here, to the B200-native implementation."""
from humangaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                           rasterize_gaussians, rasterize_views)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "rasterize_views"]

from humangaussian_b200.rasterizer import distCUDA2  # noqa: F401

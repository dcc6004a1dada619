"""Drop-in module name for the reference's vendored `simple_knn` extension (gaussiansplatting/submodules/simple-knn):
`from simple_knn._C import distCUDA2` (scene/gaussian_model.py:20, gs_renderer.py) resolves to the sm_100a kernel here."""

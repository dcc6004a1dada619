"""Frozen regression vectors of THIS REPO'S oracle (not of the reference: the rasteriser is un-vendored, parity unpinned --
see oracle/gs_oracle.c).  Inputs and outputs of one small scene are stored so that an accidental change to the oracle's
arithmetic (which defines the numerical contract the CUDA kernels are tested against) cannot go unnoticed.

    python tests/golden/make_oracle_frozen.py     ->  tests/golden/oracle_frozen.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from util import grad_images, small_scene  # noqa: E402
from oracle.gs_oracle import Oracle  # noqa: E402


def main():
    inp, _, _ = small_scene(P=300, deg=2, seed=5, H=40, W=56)
    o = Oracle(threads=2)
    color, radii, depth, alpha = o.forward(**inp)
    st = o.state()
    gimg = grad_images(40, 56, seed=3)
    grads = o.backward(*gimg)
    out = {"in_" + k: np.asarray(v) for k, v in inp.items()}
    out.update(color=color, radii=radii, depth=depth, alpha=alpha, point_list=st["point_list"], ranges=st["ranges"], keys=st["keys"],
               n_contrib=st["n_contrib"], final_T=st["final_T"], g_color=gimg[0], g_depth=gimg[1], g_alpha=gimg[2])
    out.update({"grad_" + k: v for k, v in grads.items() if v is not None})
    np.savez_compressed(os.path.join(HERE, "oracle_frozen.npz"), **out)
    print("wrote oracle_frozen.npz", os.path.getsize(os.path.join(HERE, "oracle_frozen.npz")) // 1024, "KiB; D =", st["num_rendered"])


if __name__ == "__main__":
    main()

"""Golden fixtures for the PLY codec row, produced by RUNNING THE REFERENCE'S OWN save_ply / load_ply:

    gaussiansplatting/scene/gaussian_model.py:187-266   (training convention: z-up as stored, columns sorted by index)
    gs_renderer.py:525-610                              (animation convention: y/z swap of xyz and scales, quaternion
                                                         components 2<->3 swapped and component 0 negated, columns in FILE order)

    python tests/golden/make_golden_ply.py        ->  tests/golden/ref_ply.npz   (authoring container only)

`plyfile` is not installed: it is stubbed with a minimal numpy reader/writer of the same interface (PlyData.read,
.elements[0][name], .elements[0].properties[i].name, PlyElement.describe, PlyData([el]).write) -- the stub only moves
bytes; every decision about columns, order, reshapes and axis conventions is taken by the reference's code.
gaussian_model.py is imported as is (device keyword patched away); gs_renderer.py imports CUDA-only packages at module
level, so its GaussianModel.load_ply is extracted by AST and run unmodified on a bare object.
The fixture holds the exact bytes of the PLY files the reference wrote plus the tensors its loaders produced."""
import ast
import io
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


class _Prop:
    def __init__(self, name):
        self.name = name


class PlyElement:
    def __init__(self, data, name):
        self.data, self.name = data, name
        self.properties = [_Prop(n) for n in data.dtype.names]

    @staticmethod
    def describe(arr, name):
        return PlyElement(arr, name)

    def __getitem__(self, k):
        return self.data[k]


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def write(self, path):
        el = self.elements[0]
        hdr = "ply\nformat binary_little_endian 1.0\nelement %s %d\n" % (el.name, len(el.data))
        hdr += "".join("property float %s\n" % n for n in el.data.dtype.names) + "end_header\n"
        with open(path, "wb") as f:
            f.write(hdr.encode("ascii"))
            f.write(el.data.astype([(n, "<f4") for n in el.data.dtype.names]).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            names, n = [], 0
            while True:
                tok = f.readline().decode("ascii").split()
                if tok[0] == "element":
                    n = int(tok[2])
                elif tok[0] == "property":
                    names.append(tok[2])
                elif tok[0] == "end_header":
                    break
            data = np.frombuffer(f.read(n * 4 * len(names)), dtype=[(k, "<f4") for k in names])
        return PlyData([PlyElement(data, "vertex")])


def _patch():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "empty", "tensor", "zeros_like"):
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            k.pop("device", None)
            return __orig(*a, **k)
        setattr(torch, name, wrap)
    m = types.ModuleType("plyfile")
    m.PlyData, m.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = m
    k = types.ModuleType("simple_knn")
    kc = types.ModuleType("simple_knn._C")
    kc.distCUDA2 = None
    sys.modules["simple_knn"], sys.modules["simple_knn._C"] = k, kc


def main():
    _patch()
    from gaussiansplatting.scene.gaussian_model import GaussianModel
    out = {}
    g = torch.Generator().manual_seed(77)
    nn = torch.nn
    for deg in (0, 2):
        K, P = (deg + 1) ** 2, 41
        m = GaussianModel(deg)
        m._xyz = nn.Parameter(torch.randn(P, 3, generator=g))
        m._features_dc = nn.Parameter(torch.randn(P, 1, 3, generator=g))
        m._features_rest = nn.Parameter(torch.randn(P, K - 1, 3, generator=g))
        m._opacity = nn.Parameter(torch.randn(P, 1, generator=g))
        m._scaling = nn.Parameter(torch.randn(P, 3, generator=g))
        m._rotation = nn.Parameter(torch.randn(P, 4, generator=g))
        path = f"/tmp/ref_ply_deg{deg}/point_cloud.ply"
        m.save_ply(path)                                   # the reference writes the file
        out[f"deg{deg}_file"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
            out[f"deg{deg}_saved_{k}"] = getattr(m, "_" + k).detach().numpy()
        m2 = GaussianModel(deg)
        m2.load_ply(path)                                  # ... and reads it back (training convention)
        for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
            out[f"deg{deg}_train_{k}"] = getattr(m2, "_" + k).detach().numpy()
        # animation convention: gs_renderer.GaussianModel.load_ply, extracted by AST
        tree = ast.parse(open(os.path.join(REF, "gs_renderer.py")).read())
        cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianModel"][0]
        fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "load_ply"][0]
        ns = {"torch": torch, "np": np, "nn": nn, "PlyData": PlyData, "print": lambda *a, **k: None}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), "gs_renderer_load_ply", "exec"), ns)
        obj = types.SimpleNamespace(max_sh_degree=deg)
        ns["load_ply"](obj, path)
        for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
            out[f"deg{deg}_anim_{k}"] = getattr(obj, "_" + k).detach().numpy()
    # a file whose columns are NOT in index order (the two loaders differ: training sorts, animation takes file order)
    P = 7
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_2", "scale_0", "scale_1",
             "rot_1", "rot_0", "rot_3", "rot_2"]
    arr = np.empty(P, dtype=[(n, "<f4") for n in names])
    rs = np.random.RandomState(5)
    for n in names:
        arr[n] = rs.randn(P).astype(np.float32)
    path = "/tmp/ref_ply_shuffled.ply"
    PlyData([PlyElement.describe(arr, "vertex")]).write(path)
    out["shuf_file"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    m3 = GaussianModel(0)
    m3.load_ply(path)
    obj = types.SimpleNamespace(max_sh_degree=0)
    ns["load_ply"](obj, path)
    for k in ("xyz", "scaling", "rotation", "opacity", "features_dc"):
        out[f"shuf_train_{k}"] = getattr(m3, "_" + k).detach().numpy()
        out[f"shuf_anim_{k}"] = getattr(obj, "_" + k).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "ref_ply.npz"), **out)
    print("wrote ref_ply.npz", {k: v.shape for k, v in out.items() if "file" in k})


if __name__ == "__main__":
    main()

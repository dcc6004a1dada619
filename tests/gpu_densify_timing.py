"""Timing of the device densify_and_prune against the reference's op sequence restated in torch on the same GPU
(boolean-mask gathers + torch.cat of every parameter and Adam moment, gaussian_model.py:284-415).  Not a test.
    python tests/gpu_densify_timing.py [P] [K]
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from humangaussian_b200.densify import DensifyStats, densify_and_prune  # noqa: E402

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def rotmat(q):
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)


def torch_sequence(par, mom, accum, denom, max_grad, min_opacity, extent, pd, noise_gen):
    """the reference's order of operations, one torch op per reference op"""
    par = dict(par); mom = {k: list(v) for k, v in mom.items()}

    def cat(new):
        for g in GROUPS:
            par[g] = torch.cat((par[g], new[g]), 0)
            mom[g] = [torch.cat((m, torch.zeros_like(new[g])), 0) for m in mom[g]]

    def prune(mask):
        keep = ~mask
        for g in GROUPS:
            par[g] = par[g][keep]
            mom[g] = [m[keep] for m in mom[g]]
    grads = accum / denom
    grads[grads.isnan()] = 0.0
    sel = (torch.norm(grads, dim=-1) >= max_grad) & (torch.exp(par["scaling"]).max(1).values <= pd * extent)
    cat({g: par[g][sel] for g in GROUPS})
    n = par["xyz"].shape[0]
    padded = torch.zeros(n, device=grads.device)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = (padded >= max_grad) & (torch.exp(par["scaling"]).max(1).values > pd * extent)
    stds = torch.exp(par["scaling"])[sel].repeat(2, 1)
    samples = torch.randn(stds.shape, device=stds.device, generator=noise_gen) * stds
    R = rotmat(par["rotation"][sel]).repeat(2, 1, 1)
    new = {g: par[g][sel].repeat(2, *([1] * (par[g].dim() - 1))) for g in GROUPS}
    new["xyz"] = torch.bmm(R, samples.unsqueeze(-1)).squeeze(-1) + par["xyz"][sel].repeat(2, 1)
    new["scaling"] = torch.log(torch.exp(par["scaling"])[sel].repeat(2, 1) / 1.6)
    cat(new)
    prune(torch.cat((sel, torch.zeros(2 * int(sel.sum()), device=sel.device, dtype=torch.bool))))
    prune(torch.sigmoid(par["opacity"]).squeeze() < min_opacity)
    return par, mom


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    f = lambda *s: torch.randn(*s, device=dev, generator=g)
    par = dict(xyz=f(P, 3) * 0.4, f_dc=f(P, 1, 3), f_rest=f(P, K - 1, 3) * 0.1, opacity=f(P, 1) * 2 - 1.5, scaling=f(P, 3) * 0.9 - 5.0, rotation=f(P, 4))
    mom = {k: (f(*v.shape) * 0.01, f(*v.shape).abs() * 1e-4) for k, v in par.items()}
    denom = torch.randint(0, 4, (P, 1), device=dev, generator=g).float()
    accum = f(P, 1).abs() * 3e-4 * denom
    stats = DensifyStats(accum.clone(), denom.clone(), torch.zeros(P, device=dev))
    args = (0.0002, 0.05, 0.7)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    ours = timeit(lambda: densify_and_prune(par, mom, stats, *args, None, 0.01, generator=g))
    seq = timeit(lambda: torch_sequence(par, mom, accum.clone(), denom, *args, 0.01, g))
    p2, _, _, counts = densify_and_prune(par, mom, stats, *args, None, 0.01, generator=g, return_counts=True)
    moved = sum(v.numel() for v in par.values()) * 4 * 3  # params + two moments, read once
    moved += sum(v.numel() for v in p2.values()) * 4 * 3  # written once
    print(json.dumps({"P": P, "K": K, "P_new": int(counts[4]), "clones": int(counts[1]), "split_parents": int(counts[2]),
                      "b200gs_ms": ours, "torch_op_sequence_ms": seq, "speedup": seq / ours,
                      "algorithmic_GB": moved / 1e9, "b200gs_GBps": moved / 1e9 / (ours / 1e3)}))


if __name__ == "__main__":
    main()

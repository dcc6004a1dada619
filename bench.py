#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json: fwd+bwd views/sec @1024^2, ~300k Gaussians).

A "step" = one pass of the hot path over one batch: forward + backward of VIEWS (default 64) random orbit cameras
over the same P (default 300 000, SH degree 3) Gaussians at 1024x1024 -- BASELINE.json configs[3], the configuration
the metric is quoted on.  Synthetic scene (humangaussian_b200.scene.synthetic_body: sample.ply-like statistics) and
the training camera distribution (threestudio/data/uncond.py:325-429).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU implementation (oracle port)

Prints ONE JSON line on rank 0.  `value` = whole-job views/s with inputs resident in HBM; `e2e` = same metric through the
public API with host buffers (pinned H2D of the Gaussian buffer + cameras, D2H of loss + packed gradients every step).
Multi-GPU: views are sharded (each rank renders its own VIEWS cameras: weak scaling), the scene buffer is broadcast once
over NCCL before timing, and the per-step gradient exchange (one all-reduce of the packed gradient buffer) is inside the
timed region because that is the path's only real exchange step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "classic"],
                    help="b200 = this repo's CUDA path; reference = the reference algorithm on the host CPU (oracle port); "
                         "classic = baseline/libb200gs_classic.so, the classic-structure CUDA comparator, through the same host path")
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--gaussians", type=int, default=300000)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--per-view", action="store_true",
                    help="drive the rasteriser one view per call (the reference's calling pattern, GaussianDreamer.py:244-248) "
                         "instead of one batched call per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stage-json", default=None, help="also write the per-stage timing table to this file")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self, wait_s=5.0):
        """Launch nvidia-smi and wait for its first row, so that sampling is already running when the (sub-second)
        timed region begins; rows are stamped on arrival and only those inside [begin(), end()] are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t = time.perf_counter()
            while not self.rows and time.perf_counter() - t < wait_s:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            time.sleep(0.06)  # let the last in-window row arrive
            self.proc.terminate()
        t0 = self.t0 if self.t0 is not None else -1e30
        t1 = (self.t1 if self.t1 is not None else 1e30) + 0.06
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9]
        window = "timed region"
        if not rows:  # region shorter than one sampling period: fall back to the rows nearest to it, and say so
            rows = [r for _, r in self.rows[-3:] if len(r) >= 9]
            window = "nearest samples (timed region shorter than the sampling period)"
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------- CPU reference
def oracle_views_per_s(args, n_views, threads=None):
    """fwd+bwd of n_views views of the SAME workload by the CPU oracle (all host threads; set explicitly because
    torchrun exports OMP_NUM_THREADS=1)."""
    threads = threads or os.cpu_count()
    import numpy as np
    import torch
    from humangaussian_b200.cameras import sample_orbit_cameras
    from humangaussian_b200.scene import synthetic_body
    from oracle.gs_oracle import Oracle
    p = synthetic_body(args.gaussians, sh_degree=args.sh_degree, seed=0)
    cams = sample_orbit_cameras(args.views, args.res, args.res, seed=1000)[:n_views]
    with torch.no_grad():
        a = dict(means3D=p.get_xyz.numpy(), opacities=p.get_opacity.numpy(), shs=p.get_features.numpy(), scales=p.get_scaling.numpy(),
                 rotations=p.get_rotation.numpy())
    rng = np.random.RandomState(0)
    gw = [rng.randn(c, args.res, args.res).astype(np.float32) for c in (3, 1, 1)]
    o = Oracle(threads=threads)
    t0 = time.perf_counter()
    for cam in cams:
        o.forward(**a, viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
                  campos=cam.camera_center.numpy(), bg=np.zeros(3, np.float32), image_height=args.res, image_width=args.res,
                  tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2), sh_degree=args.sh_degree)
        o.backward(*gw)
    dt = time.perf_counter() - t0
    return n_views / dt, dt


def torch_cpu_paths(args):
    """The reference's CPU-only PyTorch projection / covariance / SH path, Gaussians/s."""
    import torch
    from humangaussian_b200.cameras import sample_orbit_cameras
    from humangaussian_b200.scene import synthetic_body
    torch.set_num_threads(os.cpu_count())
    p = synthetic_body(args.gaussians, sh_degree=args.sh_degree, seed=0)
    cam = sample_orbit_cameras(1, args.res, args.res, seed=1000)[0]
    out = {}
    with torch.no_grad():
        xyz, P = p.get_xyz, p.P

        def timeit(f, n=3):
            f()
            t0 = time.perf_counter()
            for _ in range(n):
                f()
            return P * n / (time.perf_counter() - t0)

        def proj():  # geom_transform_points, graphics_utils.py:22-30
            h = torch.cat([xyz, torch.ones(P, 1)], 1) @ cam.full_proj_transform
            return h[:, :3] / (h[:, 3:] + 1e-7)

        def cov():   # get_covariance, gaussian_model.py:27-31 -> general_utils.py:64-110
            q = torch.nn.functional.normalize(p.rotation)
            r, x, y, z = q.unbind(1)
            R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y), 2 * (x * y + r * z),
                             1 - 2 * (x * x + z * z), 2 * (y * z - r * x), 2 * (x * z - r * y), 2 * (y * z + r * x),
                             1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
            L = R * p.get_scaling[:, None, :]
            S = L @ L.transpose(1, 2)
            return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)

        def sh():    # eval_sh + clamp, gaussian_renderer/__init__.py:74-78
            from oracle.dense_ref import _sh
            d = xyz - cam.camera_center[None]
            d = d / d.norm(dim=1, keepdim=True)
            return torch.clamp_min(_sh(args.sh_degree, p.get_features, d) + 0.5, 0.0)

        out = {"projection_gauss_per_s": timeit(proj), "covariance_gauss_per_s": timeit(cov), "sh_gauss_per_s": timeit(sh)}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    t_all = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        oracle_views_per_s(args, 1)
    vals, dts = [], []
    for _ in range(args.steps):
        v, dt = oracle_views_per_s(args, 1)  # each step = a bounded sample: 1 view of the batch, full fwd+bwd
        vals.append(v)
        dts.append(dt)
    value = len(vals) / sum(dts)
    line = {"impl": "reference", "metric": "fwd+bwd views/sec @1024^2, ~300k Gaussians", "value": value, "unit": "views/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / len(dts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),  # the GPU arm's config (rank 0 alone runs this arm on the host CPU)
            "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port",
                             "sample": "each step = 1 view (of the 64-view batch) fwd+bwd by oracle/gs_oracle.c with all host threads"},
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    emit(line)


def workload_config(args, n):
    return {"workload": f"BASELINE configs[3]: synthetic body P={args.gaussians} SH deg {args.sh_degree}, {args.res}x{args.res}, "
                        f"{args.views} seeded orbit cameras per GPU, fwd+bwd (loss = sum of fixed N(0,1) weights on RGB, depth, alpha)",
            "gaussians": args.gaussians, "views_per_gpu": args.views, "resolution": args.res, "sh_degree": args.sh_degree,
            "calling_pattern": "one rasteriser call per view" if getattr(args, "per_view", False) else "one batched call per step",
            "parallelism": f"views sharded over {n} GPU(s); scene broadcast once; packed-gradient all-reduce per step" if n > 1 else "single GPU",
            "l2_policy": "per-step working set (64 views x ~48 MB images/state + 0.9 GB geometry + sort buffers) >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from humangaussian_b200 import rasterizer as R
    from humangaussian_b200.cameras import sample_orbit_cameras
    from humangaussian_b200.renderer import stack_cameras
    from humangaussian_b200.scene import synthetic_body

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "B200GS_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["B200GS_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    R.load_library()
    P, V, HW, deg = args.gaussians, args.views, args.res, args.sh_degree
    K = (deg + 1) ** 2

    # ---- the scene: one flat SoA parameter buffer [xyz | scale | rot | opacity | sh], broadcast once from rank 0
    fields, n_flat = R.packed_layout(P, K)  # every field 16-byte aligned (dist.pack's layout)
    flat = torch.zeros(n_flat, device=dev)
    host_flat = torch.zeros(n_flat, pin_memory=True)
    if rank == 0:
        p = synthetic_body(P, sh_degree=deg, seed=0)
        with torch.no_grad():
            for (o, n, _), t in zip(fields, (p.get_xyz, p.get_scaling, p.get_rotation, p.get_opacity, p.get_features)):
                host_flat.narrow(0, o, n).copy_(t.reshape(-1))
        flat.copy_(host_flat, non_blocking=True)
    if world > 1:
        dist.broadcast(flat, 0)  # 4*59*P bytes over NVLink, once per parameter version
        host_flat.copy_(flat)
    flat.requires_grad_(True)

    def views_of(f):
        return [f.narrow(0, o, n).view(shape) for o, n, shape in fields]

    cams = sample_orbit_cameras(V, HW, HW, seed=1000 + rank, device=dev)
    vm, pm, cp, tanx, tany = stack_cameras(cams, dev)
    cam_host = torch.cat([vm.reshape(-1), pm.reshape(-1), cp.reshape(-1)]).cpu().pin_memory()
    bg = torch.zeros(3, device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    gw = [torch.randn(V, c, HW, HW, device=dev, generator=g) for c in (3, 1, 1)]
    stats = {}

    def step(e2e=False):
        f = flat
        v_vm, v_pm, v_cp = vm, pm, cp
        if e2e:  # host buffers in: the Gaussian buffer and the cameras come from pinned host memory every step
            with torch.no_grad():
                flat.copy_(host_flat, non_blocking=True)
            cd = cam_host.to(dev, non_blocking=True)
            v_vm, v_pm, v_cp = cd[:16 * V].view(V, 4, 4), cd[16 * V:32 * V].view(V, 4, 4), cd[32 * V:].view(V, 3)
        xyz, sc, rot, op, sh = views_of(f)
        flat.grad = None
        if args.per_view:  # V sequential single-view calls, outputs stacked (what the SDS loop does today)
            outs = [R.rasterize_views(means3D=xyz, opacities=op, viewmatrices=v_vm[i:i + 1], projmatrices=v_pm[i:i + 1],
                                      camposs=v_cp[i:i + 1], tanfovx=tanx[i:i + 1], tanfovy=tany[i:i + 1], image_height=HW,
                                      image_width=HW, bg=bg, sh_degree=deg, shs=sh, scales=sc, rotations=rot) for i in range(V)]
            c, r, d, a = (torch.cat([o[k] for o in outs], 0) for k in range(4))
        else:  # the batched entry over the packed buffer: B3 writes flat.grad (the all-reduce payload) in place
            c, r, d, a = R.rasterize_views_packed(f, P, K, viewmatrices=v_vm, projmatrices=v_pm, camposs=v_cp, tanfovx=tanx, tanfovy=tany,
                                                  image_height=HW, image_width=HW, bg=bg, sh_degree=deg)
        if e2e:
            loss = (c * gw[0]).sum() + (d * gw[1]).sum() + (a * gw[2]).sum()
            loss.backward()
        else:
            torch.autograd.backward([c, d, a], gw)
        if world > 1:
            dist.all_reduce(flat.grad)  # the path's one exchange step: packed gradients, NCCL over NVLink
        if e2e:  # host buffers out: loss + packed gradients
            stats["loss"] = float(loss.detach())
            host_grad.copy_(flat.grad, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        stats["n_vis"] = r
        return r

    host_grad = torch.empty(n_flat, pin_memory=True)

    def timed(n_steps, e2e=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    R.profile_read()  # drop warm-up spans
    R.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(2):
        step()  # keep the GPU busy right up to the timed region (nvidia-smi start-up took a moment)
    torch.cuda.synchronize()
    R.profile_read()
    l0 = R.launch_count()
    sampler.begin()
    ms_total = timed(args.steps)
    sampler.end()
    launches = R.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    stages = R.profile_read()
    R.profile_enable(False)
    ms_step = ms_total / args.steps
    value = world * V * args.steps / (ms_total * 1e-3)

    # ---- workload statistics for the algorithmic-bytes roofline (DESIGN.md "Roofline accounting")
    n_vis = int((stats["n_vis"] > 0).sum())
    D = int(R.last_num_rendered())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "MEASURED_PEAKS.json hbm_gbs (of measured)") if peaks.get("hbm_gbs") else (6650.0, "B200_PROFILING.md fallback (of fallback)")
    G_in, G_mid, pix = 44 + 12 * K, 40, HW * HW
    bytes_stage = None
    roof = None
    if D is not None:
        Vc = 1 if args.per_view else V  # views per launch set: the per-view pattern makes one call (one launch set) per view
        bytes_stage = {
            # F1 reads the P Gaussians ONCE per batch (view loop inside the thread) and writes per-view state
            "preprocess_fwd": P * G_in + Vc * P * 8 + n_vis * G_mid,
            # depth pre-sort of V*P (4 B key + 4 B index, one read + one write = single-pass lower bound) + scan
            "scan": Vc * P * (16 + 8),
            # emit (8 B) + stable tile sort (single-pass bound: 8 B read + 8 B write) + ranges (4 B read + tiles*8)
            "binning": D * 8 + D * 16 + D * 4 + Vc * (pix // 256) * 8,
            "blend_fwd": D * (4 + G_mid) + Vc * pix * 28,
            "blend_bwd": Vc * pix * 28 + D * (4 + G_mid) + n_vis * G_mid,
            "preprocess_bwd": n_vis * G_mid + P * G_in + P * (G_in + 12) + Vc * P * 12,
        }
        table = {}
        for k, (ms, calls) in stages.items():
            if calls:
                per = ms / calls
                table[k] = {"ms_per_launch_set": per, "calls": calls, "share_of_step": ms / ms_total,
                            "algorithmic_bytes": bytes_stage[k], "achieved_gbs": bytes_stage[k] / (per * 1e-3) / 1e9,
                            "frac_of_hbm_peak": bytes_stage[k] / (per * 1e-3) / 1e9 / peak}
        dom = max(table, key=lambda k: table[k]["ms_per_launch_set"])
        total_bytes = sum(bytes_stage.values())
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu capture (per view x views of this launch)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = int(tj[dom]["dram_bytes_per_view"] * Vc)
            traffic_src = tj[dom]["capture"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": table[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": table[dom]["frac_of_hbm_peak"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_stage[dom],
                "note": "blend kernels are FP32-issue/atomic bound, not HBM bound (SURVEY.md 8d caveat); see profiles/ for ncu pipe utilisation",
                "whole_step": {"algorithmic_bytes": total_bytes, "achieved_gbs": total_bytes / (ms_step * 1e-3) / 1e9,
                               "frac": total_bytes / (ms_step * 1e-3) / 1e9 / peak},
                "stages": table}

    # ---- e2e through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        step(True)
        n_e = max(2, min(args.steps, 5))
        ms_e = timed(n_e, True)
        e2e = {"value": world * V * n_e / (ms_e * 1e-3), "unit": "views/s", "h2d_bytes_per_step": int(host_flat.numel() * 4 + cam_host.numel() * 4),
               "d2h_bytes_per_step": int(host_grad.numel() * 4 + 4), "steps": n_e,
               "what": "pinned H2D of the packed Gaussian buffer + cameras, rasterize_views fwd+bwd, D2H of loss + packed gradients, every step"}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        nv = 2
        v, dt = oracle_views_per_s(args, nv)
        cpu = {"value": v, "unit": "views/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{nv} of the {V} views of this workload, full fwd+bwd by oracle/gs_oracle.c (OpenMP, all host threads), {dt:.1f} s",
               "torch_cpu_paths": torch_cpu_paths(args)}

    if rank == 0:
        line = {"impl": args.impl, "metric": "fwd+bwd views/sec @1024^2, ~300k Gaussians", "value": value, "unit": "views/s", "n_gpus": world,
                "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
                "views_per_step_per_gpu": V, "instances_per_step": D, "visible_per_step": n_vis,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        emit(line)
        if args.stage_json and roof:
            os.makedirs(os.path.dirname(os.path.abspath(args.stage_json)), exist_ok=True)
            json.dump({"ms_per_step": ms_step, "views_per_s": value, "stages": roof["stages"], "clocks": clocks}, open(args.stage_json, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 during the run
    (NCCL's version banner, library chatter) was re-routed to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)  # C-level writers to stdout (e.g. "NCCL version ...") must not pollute the one-line contract
    if args.impl == "classic":
        os.environ["B200GS_LIB"] = os.path.join(ROOT, "baseline", "libb200gs_classic.so")
        args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

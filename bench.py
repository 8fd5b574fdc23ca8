#!/usr/bin/env python
"""bench.py -- Mrays/s fwd+bwd of the NeuS volume-rendering hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU implementation (oracle port) on the host cores

Workload (config.workload = "cfg2"): BASELINE.json configs[1], "neus_in_10_minutes BMVS-style object (LoTD + fused MLP),
800x600, 1 B200": a CFG-sized LoTDNeuS (16-level LoTD, 12.13 M params, 32->64->1 SDF decoder, 58->64->64->3 radiance net,
64^3 occupancy grid), synthetic sphere-like SDF, one full 800x600 frame (480 000 rays) per step, rendered in ray chunks.
A step = ray_test -> march -> 3-stage up-sampling -> boundary SDF (grad) -> alpha -> compression -> colour/normal query
-> compositing -> scalar loss over rgb/depth/normals/mask -> backward to every parameter (-> one NCCL all-reduce of the
flat gradient when N > 1).  Weak scaling: every rank renders its own full frame (its own camera pose).

Timing: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks; L2 is flushed between
steps (256 MiB write, outside the per-step event pairs); clocks are sampled with nvidia-smi during the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 600, 800
N_VIEWS = 8


# ---------------------------------------------------------------------------------------------- synthetic scene (cfg 2)
def pinhole_rays(Hh, Ww, cam_pos, focal=None):
    focal = (Hh + Ww) / 2. if focal is None else focal
    cam = np.asarray(cam_pos, dtype=np.float64)
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, np.array([0., 0., 1.])); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    j, i = np.meshgrid(np.arange(Hh, dtype=np.float64), np.arange(Ww, dtype=np.float64), indexing="ij")
    d = ((i + 0.5 - Ww / 2.) / focal)[..., None] * right + ((j + 0.5 - Hh / 2.) / focal)[..., None] * down + fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(cam, d.shape)
    return (torch.from_numpy(np.ascontiguousarray(o.reshape(-1, 3), dtype=np.float32)),
            torch.from_numpy(np.ascontiguousarray(d.reshape(-1, 3), dtype=np.float32)))


def orbit(k, n, radius=3.0, elev_deg=20.0):
    a, e = 2 * math.pi * k / n, math.radians(elev_deg)
    return (radius * math.cos(e) * math.cos(a), radius * math.cos(e) * math.sin(a), radius * math.sin(e))


def build_model(device, seed=42, radius=0.5, noise=2.0e-3, ln_inv_s_init=0.5298, k_pass=8.0, collect_samples=False):
    """CFG-sized model whose SDF is a noisy sphere (the state `pretrain_sdf_sphere` would reach; see oracle/scene.py for
    the same construction on the oracle side).  inv_s = exp(10 * 0.5298) ~ 200."""
    from neuralsim_b200.fields import LoTDNeuSModel
    gen = torch.Generator(device=device).manual_seed(seed)
    model = LoTDNeuSModel(
        surface_cfg=dict(bounding_size=2.0, encoding_cfg=dict(lotd_auto_compute_cfg=dict(type="gen_ngp", min_res=16, n_feats=2, log2_hashmap_size=19,
                                                                                         per_level_scale=1.382, num_levels=16),
                                                              param_init_cfg=dict(type="uniform_to_type", bound=noise))),
        radiance_cfg=dict(n_appear_embedding=4, dir_embed_cfg=dict(type="spherical", degree=4), D=2, W=64),
        var_ctrl_cfg=dict(ln_inv_s_init=ln_inv_s_init, ln_inv_s_factor=10.0),
        accel_cfg=dict(resolution=[64, 64, 64], occ_val_fn_cfg=dict(type="sdf", inv_s=256.0), occ_thre=0.3, ema_decay=0.95,
                       update_from_samples_cfg=dict() if collect_samples else None),
        ray_query_cfg=dict(query_mode="march_occ_multi_upsample_compressed", query_param=dict(
            nablas_has_grad=True, num_coarse=64, num_fine=[8, 8, 32], coarse_step_cfg=dict(step_mode="linear"),
            march_cfg=dict(step_size=0.005, max_steps=4096), upsample_inv_s=64.0, upsample_inv_s_factors=[1, 4, 16],
            upsample_use_estimate_alpha=True)),
        device=device, generator=gen)
    enc = model.implicit_surface.encoding
    meta = enc.meta
    lvl = 5
    res = meta.level_res_multidim[lvl]
    ax = [((torch.arange(r, dtype=torch.float64) - 0.5) / (r - 2) * 2 - 1) for r in res]
    gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
    s = (torch.sqrt(gx * gx + gy * gy + gz * gz) - radius).float().to(device)
    with torch.no_grad():
        off, nf = meta.level_offsets[lvl], meta.level_n_feats[lvl]
        enc.flattened_params[off:off + meta.level_n_params[lvl]].view(*res, nf)[..., 0] = s
        f_idx = sum(meta.level_n_feats[:lvl])
        d0, d1 = model.implicit_surface.decoder.layers
        d0.weight[:, f_idx] = 0.
        d1.weight.mul_(0.05); d1.bias.zero_()
        d0.weight[0].zero_(); d0.weight[1].zero_()
        d0.weight[0, f_idx], d0.weight[1, f_idx] = k_pass, -k_pass
        d0.bias[0], d0.bias[1] = 0., 0.
        d1.weight[0, 0], d1.weight[0, 1] = 1. / k_pass, -1. / k_pass
        c = (torch.arange(64, dtype=torch.float64) + 0.5) / 64 * 2 - 1
        cx, cy, cz = torch.meshgrid(c, c, c, indexing="ij")
        occ = ((torch.sqrt(cx * cx + cy * cy + cz * cz) - radius).abs() < 0.012 + math.sqrt(3.) / 64)
        model.accel.occ.set_occ_grid(occ.to(device))
    return model


def flat_grad_views(model):
    """One flat fp32 buffer holding every gradient (p.grad are views) -> a single all-reduce per step."""
    params = [p for p in model.parameters() if p.requires_grad]
    flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device)
    o = 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    return flat, params


def loss_of(rendered):
    return rendered["rgb_volume"].mean() + rendered["depth_volume"].mean() + rendered["normals_volume"].mean() + rendered["mask_volume"].mean()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons during the timed region.

    Measured on this pool (profiles/diag_clock_sampler.sh, /tmp logs of round 1): ANY concurrent poller -- a looping `nvidia-smi -lms 100..1000`,
    a thread forking nvidia-smi, an NVML thread -- makes 7 of 8 runs of this launch- and sync-heavy step show 80-230 ms stalls inside 15 ms steps
    (NVML queries disturb CUDA submission for ~100 ms; even a query made between two steps shows up in the next ones, profiles/diag_*.sh); with no
    poller -- and no NVML session inside this process -- steps are 14.2-15.8 ms.  So ONE sample is taken, by a separate one-shot `nvidia-smi`
    process, at the one moment it cannot perturb the measurement and is still "during" it: right after the LAST step of the resident loop has been
    enqueued -- the GPU is then executing the timed steps and nothing of ours is left to launch."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0, period_ms=100):
        self.index, self.enabled = index, period_ms > 0
        self.samples, self.reasons, self.max_mhz, self.how = [], set(), None, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.phys = vis.split(",")[index].strip() if vis and len(vis.split(",")) > index else str(index)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def sample(self):
        """ONE `nvidia-smi` query (a separate process; nothing NVML-related ever lives in this one), called by the timing loop once the
        last resident step has been enqueued: the GPU is executing the timed steps, and nothing of ours is left to launch."""
        if not self.enabled:
            return
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", self.phys], capture_output=True, text=True,
                                 timeout=10).stdout.strip().split(",")
            self.samples.append(float(out[0])); self.max_mhz = float(out[1])
            self.how = "one nvidia-smi query while the GPU executes the enqueued timed steps (after the last launch of the resident loop)"
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                if "Active" in v and "Not" not in v:
                    self.reasons.add(name)
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no sampler"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "how": self.how}


def effective_cpus():
    """host cores this container may actually use: min(visible cores, cgroup CPU-bandwidth quota) -- the GPU boxes of this pool show 128
    cores under a 16-CPU quota; 128 busy threads there are throttled for most of every 100 ms period."""
    n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# ---------------------------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_baseline(sample_rays=1536, threads=None):
    """The reference algorithm's CPU implementation (oracle port of nr3d_lib's path) on a bounded sample of the same workload:
    `sample_rays` rays of the 800x600 frame (strided over the image), fwd+bwd.  Runs on the host cores."""
    from oracle import render as orender, scene as oscene
    threads = threads or effective_cpus()
    torch.set_num_threads(threads)
    P = oscene.make_sphere_params()
    occ = oscene.make_occ_grid()
    ro, rd = oscene.pinhole_rays(H, W, oscene.orbit_camera(0, N_VIEWS))
    sel = torch.linspace(0, ro.shape[0] - 1, sample_rays).long()
    ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
    P.requires_grad_(True)
    t0 = time.perf_counter()
    rt = orender.ray_test(ro, rd, near=0.01)
    vb, _ = orender.neus_ray_query(P, occ, rt, rays_h_appear=torch.zeros(rt["num_rays"], P.n_appear))
    out = orender.volume_integration(vb, ro.shape[0])
    (out["rgb_volume"].mean() + out["depth_volume"].mean() + out["normals_volume"].mean() + out["mask_volume"].mean()).backward()
    dt = time.perf_counter() - t0
    return dict(value=sample_rays / dt / 1e6, unit="Mrays/s", cores=threads, kind="port",
                sample=f"{sample_rays} rays strided over the 800x600 frame of view 0, fwd+bwd, {dt:.1f} s wall "
                       f"(the numpy LoTD port walks its 16 levels on a thread pool, the torch MLP part uses {threads} threads)")


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    rays = args.ref_rays
    times = []
    for i in range(warm + steps):
        r = cpu_baseline(rays)
        if i >= warm:
            times.append(rays / r["value"] / 1e6)
    ms = 1e3 * float(np.mean(times))
    val = rays / (ms / 1e3) / 1e6
    base = dict(value=val, unit="Mrays/s", cores=effective_cpus(), kind="port",
                sample=f"{rays} rays of the 800x600 frame per step (bounded sample), fwd+bwd, oracle port of the reference path")
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/sec fwd+bwd", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "cfg2", "frame": "800x600", "rays_per_step": rays, "note": "CPU, bounded sample"},
        "cpu_baseline": base, "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def use_reference_cuda_kernels():
    """route the SAME host orchestration through the REFERENCE'S OWN CUDA kernels (oracle/_ref: `_lotd`,
    `_pack_ops`, `_occ_grid`, `_shencoder` compiled from /root/reference) and switch every fused path off, i.e. the op
    sequence nr3d_lib executes (16-level gather kernels, autocast cuBLAS MLPs, one-thread-per-pack pack_ops, fp16 atomics).
    Only used by `--impl reference-cuda` and the `reference_cuda` field; never by the product path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    mods = {n: build_ref.load(n) for n in ("_lotd", "_pack_ops", "_occ_grid", "_shencoder")}
    if any(m is None for m in mods.values()):
        return False
    import neuralsim_b200.fields.encoding as E
    import neuralsim_b200.fields.networks as NW
    import neuralsim_b200.graphics.pack_ops as GP
    import neuralsim_b200.graphics.raymarch as GR
    import neuralsim_b200.fields.space as SP
    import neuralsim_b200.graphics.neus as GN
    E._backend, GP._backend, GR._backend, NW._shencoder = mods["_lotd"], mods["_pack_ops"], mods["_occ_grid"], mods["_shencoder"]
    NW.LoTDSDF._fusable = lambda self: False       # no fused SDF / colour kernels
    GN.FUSED_STAGES = False                        # the op-by-op chain of neus_ray_query.py / single_volume_renderer.py, every op a reference kernel or ATen
    SP.FUSED_RAY_TEST = False                      # the torch chain of aabb.py
    return True


# ---------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--rayschunk", type=int, default=H * W, help="rays per render call (default: the whole frame in one call)")
    ap.add_argument("--rays", type=int, default=H * W, help="rays per step (default: the full 800x600 frame)")
    ap.add_argument("--random-rays", action="store_true", help="draw the --rays rays of every pose as random pixels (a training batch) instead of the first rows")
    ap.add_argument("--collect-samples", action="store_true", help="accel.update_from_samples_cfg = {} as in the shipped training config: every "
                    "training-time SDF query also feeds the occupancy grid's evidence buffer (in-kernel here, torch_scatter in the reference)")
    ap.add_argument("--clock-period-ms", type=int, default=100, help="0 = no clock sampler (diagnostics); any other value: sample between steps")
    ap.add_argument("--ref-rays", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.set_num_threads(max(1, min(8, effective_cpus())))      # host side of the GPU arm is one launching thread; keep the intra-op pool small
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from neuralsim_b200 import _lib
    from neuralsim_b200.renderer import SingleVolumeRenderer
    if args.impl == "reference-cuda" and not use_reference_cuda_kernels():
        raise SystemExit("bench.py: oracle/_ref is not built (python oracle/build_ref.py in the build container)")
    model = build_model(device, collect_samples=args.collect_samples).train()
    renderer = SingleVolumeRenderer(dict(near=0.01, rayschunk=0)).train()
    flat, params = flat_grad_views(model)
    n_rays = args.rays
    views_host, views_dev = [], []
    for k in range(N_VIEWS):
        o, d = pinhole_rays(H, W, orbit((k * world + rank) % (N_VIEWS * world), N_VIEWS * world))
        if args.random_rays:
            sel = torch.randperm(o.shape[0], generator=torch.Generator().manual_seed(1000 + k))[:n_rays]
            o, d = o[sel].contiguous(), d[sel].contiguous()
        else:
            o, d = o[:n_rays].contiguous(), d[:n_rays].contiguous()
        views_host.append((o.pin_memory(), d.pin_memory()))
        views_dev.append((o.to(device), d.to(device)))
    h_appear = torch.zeros(args.rayschunk, 4, device=device)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    stats = dict(samples=0, points=0)

    def step(o, d):
        """fwd+bwd of one frame, chunked over rays; gradients accumulate into the flat buffer."""
        flat.zero_()
        total = torch.zeros((), device=device)
        for s in range(0, n_rays, args.rayschunk):
            e = min(s + args.rayschunk, n_rays)
            out = renderer.render(model, o[s:e], d[s:e], rays_h_appear=h_appear[:e - s])["rendered"]
            loss = loss_of(out) * ((e - s) / n_rays)
            if loss.requires_grad:             # a chunk whose rays all miss the object renders constants
                loss.backward()
            total += loss.detach()
        if world > 1:
            dist.all_reduce(flat)            # the one collective of a step: sum of the flat gradient
        return total

    def timed(fn, k, sampler=None):
        evs = []
        for i in range(k):
            flush_buf.fill_(i & 0xff)         # L2 flush, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i); b.record()
            evs.append((a, b))
        if sampler is not None:              # every step is enqueued and the GPU is still executing them: sample now, then drain
            sampler.sample()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    # every camera pose of the timed loops has been rendered once during warm-up (the caching allocator has seen its tensor sizes)
    n_views = max(1, min(N_VIEWS, args.warmup))

    def resident(i):
        o, d = views_dev[i % n_views]
        return step(o, d)

    loss_host = torch.zeros((), pin_memory=True)

    def e2e(i):
        oh, dh = views_host[i % n_views]
        o, d = oh.to(device, non_blocking=True), dh.to(device, non_blocking=True)     # H2D of the step's rays
        loss_host.copy_(step(o, d), non_blocking=True)                                # D2H of the step's result
        torch.cuda.current_stream().synchronize()

    with ClockSampler(local, args.clock_period_ms) as clocks:          # see the class: one clock query after the last timed launch
        for i in range(args.warmup):
            resident(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches0 = _lib.launch_count()
        import gc
        gc.collect()
        gc.disable()                           # no collector pause inside a 15 ms step; re-enabled right after the timed loops
        t_e2e = timed(e2e, args.steps)       # the headline loop runs first: nothing has queried the GPU's management interface yet
        if world > 1:
            dist.barrier()
        launches0 = _lib.launch_count()
        t_res = timed(resident, args.steps, clocks)      # ... and the one clock query comes after its last launch (see ClockSampler)
        launches = _lib.launch_count() - launches0
        gc.enable()
        time.sleep(0.5)
        # a separate, instrumented pass for the roofline: CUDA events around every launch of our kernels (not part of `value`)
        _lib.KERNEL_TIMER.enable()
        t_inst = timed(resident, args.steps)
        _lib.KERNEL_TIMER.disable()
    ms = torch.tensor([sum(t_res) / args.steps, sum(t_e2e) / args.steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(ms[0]), float(ms[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    kt = _lib.KERNEL_TIMER.summary()
    # the hash gather lives in three kernels: k_fused_sdf_tc (no-grad and autograd forward) and k_lotd_fwd (colour points)
    gather = None
    for key in ("fused_sdf_fwd", "lotd_gather"):
        if key in kt:
            gather = kt[key] if gather is None else {k: gather[k] + kt[key][k] for k in gather}
    traffic, traffic_note = None, None
    try:       # dram bytes of the dominant gather launch, from the committed ncu --set full capture (per launch, like `achieved`)
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = tj["traffic_bytes_per_launch"]
        traffic_note = {"unit": "bytes per launch (dram read + write)", "algorithmic_bytes_of_that_launch": tj["algorithmic_bytes_per_launch"], "source": tj["source"]}
    except Exception:
        pass
    roof = None
    if gather:
        achieved = gather["units"] * 512.0 / (gather["ms"] * 1e-3) / 1e9     # 512 B of table per encoded point
        roof = {"kernel": "LoTD hash gather: k_fused_sdf_tc (gather + tcgen05 decoder) and k_lotd_fwd", "bound": "hbm", "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                "points_per_launch": gather["units"] / max(gather["launches"], 1), "launches": gather["launches"],
                "avg_launch_ms": gather["ms"] / max(gather["launches"], 1), "share_of_step": gather["ms"] / max(sum(t_inst), 1e-9),
                "per_kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in kt.items()},
                "per_kernel_points_per_step": {k: v["units"] / args.steps for k, v in kt.items()}}
    line = {
        "metric": "Mrays/sec fwd+bwd", "value": world * n_rays / (ms_res * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": "cfg2", "model": "LoTDNeuS 16x2 LoTD (12.13M params) + 32-64-1 SDF MLP + 58-64-64-3 radiance MLP",
                   "frame": "800x600", "rays_per_step_per_gpu": n_rays, "ray_order": "random pixels" if args.random_rays else "image rows", "collect_samples": bool(args.collect_samples), "rayschunk": args.rayschunk, "samples_per_ray": "<=116 boundary, <=1024 marched",
                   "parallelism": f"dp{world} ray-shard, 1 all-reduce/step", "l2": "256 MiB L2 flush between steps; per-step working set >> 126 MB", "camera_poses": n_views},
        "e2e": {"value": world * n_rays / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(2 * n_rays * 3 * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof,
        "step_ms": {"resident": [round(x, 3) for x in t_res], "e2e": [round(x, 3) for x in t_e2e]},
    }
    if args.impl == "reference-cuda":
        line["impl"] = "reference-cuda"
        line["gpu_launches"] = 0
        line["config"]["note"] = "the reference's own CUDA kernels (oracle/_ref) under the same orchestration; no neuralsim_b200 kernel runs"
    elif world == 1 and not args.no_ref_cuda:
        # B1 of BASELINE.md, measured in a child process so that none of its module patching can leak into this arm
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference-cuda", "--steps", "2", "--warmup", "2",
                                "--no-cpu-baseline", "--rayschunk", str(args.rayschunk), "--rays", str(args.rays)] + (["--random-rays"] if args.random_rays else [])
                               + (["--collect-samples"] if args.collect_samples else []),
                               capture_output=True, text=True, timeout=600)
            rl = json.loads(r.stdout.strip().splitlines()[-1])
            line["reference_cuda"] = {"value": rl["value"], "unit": "Mrays/s", "ms_per_step": rl["ms_per_step"], "e2e": rl["e2e"]["value"],
                                      "what": "reference nr3d_lib CUDA kernels compiled from /root/reference (oracle/_ref), same B200, same workload"}
        except Exception as ex:
            line["reference_cuda"] = {"unavailable": repr(ex)[:200]}
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(args.ref_rays)
        except Exception as ex:  # the oracle is test infrastructure; never let it take the bench line down
            line["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Workload cfg3 of bench.py -- BASELINE.json configs[2], "StreetSurf street segment with LiDAR rays": the close-range model of
code_single/configs/waymo/streetsurf/withmask_withlidar_joint.240219.yaml (`LoTDNeuSStreet`): cuboid aabb 40 x 150 x 15 m, cuboid LoTD from the
`ngp` auto config with a 2^20 hash table, single-block occupancy grid with 1 m voxels (the shipped StreetSurf configs use the
single cuboid `occ_grid`, not the forest), step_size 0.2, num_coarse 128, num_fine [8, 8, 32], sdf_scale 25; per step 8192 camera rays (rgb +
normals) and 8192 LiDAR rays (`with_rgb=False, with_normal=True`, code_single/tools/train.py:900).

One deviation, stated in the bench line: `max_num_levels = 16` (the shipped config leaves it open and the auto config then yields 17 levels =
34 features; the fused kernels are built for 16 x 2 features -- a 17-level model runs through the op-by-op path, tests/test_cfg3_gpu.py).
Synthetic scene: a road plane 2 m below the sensors (the state `pretrain_after_zero_out` with target_shape road_surface reaches), fp16 noise on
the table, inv_s ~ 200."""
import math

import numpy as np
import torch

AABB = [[-20., -75., -7.5], [20., 75., 7.5]]
NEAR, FAR = 0.1, 200.0
ROAD_Z = -5.5            # metres; sensors ride 2 m above it
SDF_SCALE = 25.0
N_CAM = N_LIDAR = 8192


def build_model(device, seed=42, noise=2.0e-3, ln_inv_s_init=0.5298, k_pass=8.0, max_num_levels=16, log2_hashmap_size=20, target_num_params=32 * 2 ** 20):
    from neuralsim_b200.fields import LoTDNeuSModel
    gen = torch.Generator(device=device).manual_seed(seed)
    model = LoTDNeuSModel(
        surface_cfg=dict(aabb=AABB, sdf_scale=SDF_SCALE,
                         encoding_cfg=dict(lotd_use_cuboid=True,
                                           lotd_auto_compute_cfg=dict(type="ngp", target_num_params=target_num_params, min_res=16, n_feats=2,
                                                                      log2_hashmap_size=log2_hashmap_size, max_num_levels=max_num_levels),
                                           param_init_cfg=dict(type="uniform_to_type", bound=noise))),
        radiance_cfg=dict(n_appear_embedding=4, dir_embed_cfg=dict(type="spherical", degree=4), D=2, W=64),
        var_ctrl_cfg=dict(ln_inv_s_init=ln_inv_s_init, ln_inv_s_factor=10.0),
        accel_cfg=dict(vox_size=1.0, occ_val_fn_cfg=dict(type="sdf", inv_s=256.0), occ_thre=0.3, ema_decay=0.95, update_from_samples_cfg=None),
        ray_query_cfg=dict(query_mode="march_occ_multi_upsample_compressed", query_param=dict(
            nablas_has_grad=True, num_coarse=128, num_fine=[8, 8, 32], coarse_step_cfg=dict(step_mode="linear"),
            march_cfg=dict(step_size=0.2, max_steps=4096), upsample_inv_s=64.0, upsample_inv_s_factors=[1, 4, 16],
            upsample_use_estimate_alpha=False)),
        device=device, generator=gen)
    install_plane(model, ROAD_Z, k_pass)
    return model


def install_plane(model, road_z, k_pass=8.0):
    """sdf(x) = (z - road_z) / sdf_scale written into feature 0 of the last dense level (a linear function: exact under trilinear interpolation)
    and passed through the decoder (two hidden units with weights +-k: softplus(k s) - softplus(-k s) = k s for beta = 100)"""
    enc = model.implicit_surface.encoding
    meta = enc.meta
    lvl = max(l for l in range(meta.n_levels) if int(meta.level_types[l]) == 0)            # last Dense level
    res = meta.level_res_multidim[lvl]
    half_z = (AABB[1][2] - AABB[0][2]) / 2.
    zc = (AABB[1][2] + AABB[0][2]) / 2.
    az = ((torch.arange(res[2], dtype=torch.float64) - 0.5) / (res[2] - 2) * 2 - 1) * half_z + zc          # metres
    s = ((az - road_z) / model.implicit_surface.sdf_scale).float()
    dev = enc.flattened_params.device
    with torch.no_grad():
        off, nf = meta.level_offsets[lvl], meta.level_n_feats[lvl]
        enc.flattened_params[off:off + meta.level_n_params[lvl]].view(*res, nf)[..., 0] = s.to(dev).view(1, 1, -1)
        f_idx = sum(meta.level_n_feats[:lvl])
        d0, d1 = model.implicit_surface.decoder.layers
        d0.weight[:, f_idx] = 0.
        d1.weight.mul_(0.05); d1.bias.zero_()
        d0.weight[0].zero_(); d0.weight[1].zero_()
        d0.weight[0, f_idx], d0.weight[1, f_idx] = k_pass, -k_pass
        d0.bias[0], d0.bias[1] = 0., 0.
        d1.weight[0, 0], d1.weight[0, 1] = 1. / k_pass, -1. / k_pass
        r = model.accel.occ.occ_grid.shape
        cz = ((torch.arange(r[2], dtype=torch.float64) + 0.5) / r[2] * 2 - 1) * half_z + zc
        occ = ((cz - road_z).abs() < 1.0).view(1, 1, -1).expand(*r)
        model.accel.occ.set_occ_grid(occ.contiguous().to(dev))
    return model


def camera_rays(k, n, seed=0):
    """n random pixels of a 1920 x 1280 front camera 2 m above the road, at the k-th pose along the street (y axis)"""
    rng = np.random.default_rng(1000 * seed + k)
    Wc, Hc, focal = 1920, 1280, 2000.0
    cam = np.array([rng.uniform(-3, 3), -60.0 + 12.0 * k, ROAD_Z + 2.0])
    yaw = math.radians(rng.uniform(-20, 20))
    fwd = np.array([math.sin(yaw), math.cos(yaw), 0.0])
    right = np.array([math.cos(yaw), -math.sin(yaw), 0.0])
    down = np.array([0.0, 0.0, -1.0])
    i, j = rng.uniform(0, Wc, n), rng.uniform(0, Hc, n)
    d = ((i - Wc / 2) / focal)[:, None] * right + ((j - Hc / 2) / focal)[:, None] * down + fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(cam, d.shape)
    return torch.from_numpy(np.ascontiguousarray(o, dtype=np.float32)), torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))


def lidar_rays(k, n, seed=0):
    """n beams of a spinning 64-line LiDAR (elevation -17.6 .. +2.4 deg) on the roof, same pose sequence"""
    rng = np.random.default_rng(5000 * seed + k)
    org = np.array([0.0, -60.0 + 12.0 * k, ROAD_Z + 2.2])
    elev = np.radians(rng.choice(np.linspace(-17.6, 2.4, 64), n))
    azim = rng.uniform(0, 2 * math.pi, n)
    d = np.stack([np.cos(elev) * np.sin(azim), np.cos(elev) * np.cos(azim), np.sin(elev)], -1)
    o = np.broadcast_to(org, d.shape)
    return torch.from_numpy(np.ascontiguousarray(o, dtype=np.float32)), torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))


def make_views(k, rank=0, world=1):
    return camera_rays(k * world + rank, N_CAM), lidar_rays(k * world + rank, N_LIDAR)


def loss_cam(rendered):
    return rendered["rgb_volume"].mean() + rendered["depth_volume"].mean() * 1e-2 + rendered["normals_volume"].mean() + rendered["mask_volume"].mean()


def loss_lidar(rendered):
    return rendered["depth_volume"].mean() * 1e-2 + rendered["normals_volume"].mean() + rendered["mask_volume"].mean()

"""The static (device-resident sizes) NeuS step and its one-launch CUDA-graph form against the host-sized path: same kernels, so
the rendered images must be BIT-equal and the gradients equal up to the order of the fp32 atomics (graphics/neus_static.py)."""
import pytest
import torch

from oracle import scene as oscene
from util import make_pair, product_grads, rel_l2

pytestmark = pytest.mark.gpu
KEYS = ("rgb_volume", "depth_volume", "normals_volume", "mask_volume")


def _loss(rendered):
    return sum(rendered[k].mean() for k in KEYS)


def _host_sized(model, ro, rd, ha, training=True, grad=True):
    from neuralsim_b200.renderer import SingleVolumeRenderer
    r = SingleVolumeRenderer(dict(near=0.01)).train(training)
    model.train(training)
    model.zero_grad(set_to_none=True)
    with torch.set_grad_enabled(grad):
        out = r.render(model, ro, rd, rays_h_appear=ha, return_buffer=True, return_details=True)
        if grad:
            _loss(out["rendered"]).backward()
    return out, (product_grads(model) if grad else None)


def _rays(cuda, H=36, W=48, k=1, shuffle=False):
    ro, rd = oscene.pinhole_rays(H, W, oscene.orbit_camera(k, 8, radius=3.0, elev_deg=25.0))
    if shuffle:
        p = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(7))
        ro, rd = ro[p].contiguous(), rd[p].contiguous()
    return ro.to(cuda), rd.to(cuda)


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_static_step_equals_host_sized_step(cuda, shuffle, training):
    from neuralsim_b200.graphics.neus_static import render_static, sliced_volume_buffer
    _, model = make_pair(cuda)
    ro, rd = _rays(cuda, shuffle=shuffle)
    ha = torch.zeros(ro.shape[0], 4, device=cuda)
    ref, g_ref = _host_sized(model, ro, rd, ha, training=training)
    model.zero_grad(set_to_none=True)
    rendered, cnt, buffers = render_static(model, ro, rd, ha, near=0.01, march_cap=1 << 18, kept_cap=1 << 17, coherent=not shuffle)
    _loss(rendered).backward()
    c = cnt.tolist()
    assert c[20] == 0
    assert c[0] == ref["ray_tested"]["num_rays"]
    assert c[3] == int(ref["details"]["march.num_per_ray"].sum()) and c[4] == ref["details"]["march.num_per_ray"].shape[0]
    assert c[19] == ref["volume_buffer"]["t"].shape[0] and c[21] == ref["volume_buffer"]["pack_infos_hit"].shape[0]
    for k in KEYS:
        assert torch.equal(rendered[k], ref["rendered"][k]), k
    vb = sliced_volume_buffer(buffers, cnt)
    for k in ("t", "opacity_alpha", "rgb", "nablas", "rays_inds_hit", "pack_infos_hit"):
        assert torch.equal(vb[k], ref["volume_buffer"][k]), k
    g = product_grads(model)
    for k, v in g_ref.items():
        if v is None:
            assert g[k] is None or float(g[k].abs().max()) == 0.0, k
            continue
        assert rel_l2(g[k], v) <= 2e-5, (k, rel_l2(g[k], v))


def test_graph_replay_equals_host_sized_step(cuda):
    """capture once, replay on three different views: one launch per step, same images and gradients as the host-sized path"""
    from neuralsim_b200.graphics.neus_static import StaticFrame
    _, model = make_pair(cuda)
    model.train()
    views = (1, 5, 2)
    n = _rays(cuda)[0].shape[0]
    ha = torch.zeros(n, 4, device=cuda)
    refs = []
    for k in views:
        ro, rd = _rays(cuda, k=k)
        out, g = _host_sized(model, ro, rd, ha)
        refs.append(({kk: out["rendered"][kk].detach().clone() for kk in KEYS}, g))
        del out                                          # drop the autograd graph: its AccumulateGrad nodes live on the default stream
    for p in model.parameters():
        p.grad = torch.zeros_like(p)                     # the buffers the graph accumulates into
    frame = StaticFrame(model, n, loss_fn=_loss, near=0.01, zero_grads=True, slack=2.0)
    for k, (ref, g_ref) in zip(views, refs):
        ro, rd = _rays(cuda, k=k)
        frame.step(ro, rd, ha)
        assert frame.counts()["overflow"] == 0
        for kk in KEYS:
            assert torch.equal(frame.rendered[kk], ref[kk]), (k, kk)
        assert abs(float(frame.loss) - float(_loss(ref))) <= 1e-6 * abs(float(frame.loss))
        g = product_grads(model)
        for kk, v in g_ref.items():
            if v is not None:
                assert rel_l2(g[kk], v) <= 2e-5, (k, kk, rel_l2(g[kk], v))
    assert frame.captures == 1


def test_graph_replay_is_stable_across_replays(cuda):
    """replaying the same graph on new rays (no re-capture): identical to an eager static run of those rays"""
    from neuralsim_b200.graphics.neus_static import StaticFrame, render_static
    _, model = make_pair(cuda)
    model.train()
    ro0, rd0 = _rays(cuda, k=0)
    n = ro0.shape[0]
    ha = torch.zeros(n, 4, device=cuda)
    frame = StaticFrame(model, n, loss_fn=None, near=0.01, slack=2.0)
    with torch.no_grad():
        frame.step(ro0, rd0, ha)
        for k in (3, 6, 1):
            ro, rd = _rays(cuda, k=k)
            frame.step(ro, rd, ha)
            assert frame.counts()["overflow"] == 0
            ref, _, _ = render_static(model, ro, rd, ha, near=0.01, march_cap=frame.march_cap, kept_cap=frame.kept_cap, coherent=frame.coherent)
            for kk in KEYS:
                assert torch.equal(frame.rendered[kk], ref[kk]), (k, kk)
    assert frame.captures == 1


def test_arena_overflow_is_flagged_and_recovered(cuda):
    from neuralsim_b200.graphics.neus_static import StaticFrame, render_static
    _, model = make_pair(cuda)
    model.train()
    ro, rd = _rays(cuda)
    ha = torch.zeros(ro.shape[0], 4, device=cuda)
    with torch.no_grad():
        for caps, bit in ((dict(march_cap=1024, kept_cap=1 << 17), 1), (dict(march_cap=1 << 18, kept_cap=256), 2)):
            rendered, cnt, _ = render_static(model, ro, rd, ha, near=0.01, coherent=True, **caps)
            assert int(cnt[20]) & bit
            if bit == 1:
                assert int(cnt[12]) == 0 and int(cnt[13]) == 0
            else:
                assert float(rendered["mask_volume"].abs().sum()) == 0.0        # nothing composited, nothing written out of bounds
        frame = StaticFrame(model, ro.shape[0], near=0.01, march_cap=1024, kept_cap=256, coherent=True, use_graph=False)
        frame.step(ro, rd, ha)
        assert frame.check() is False                   # re-sized from this batch and re-run
        assert frame.counts()["overflow"] == 0 and float(frame.rendered["mask_volume"].sum()) > 50.0


def test_random_training_batch(cuda):
    """4096 random pixels of a frame (MODE 1 traversal, the reference's training batch): static graph == host-sized path"""
    from neuralsim_b200.graphics.neus_static import StaticFrame
    _, model = make_pair(cuda)
    model.train()
    ro, rd = oscene.pinhole_rays(150, 200, oscene.orbit_camera(2, 8))
    sel = torch.randperm(ro.shape[0], generator=torch.Generator().manual_seed(3))[:4096]
    ro, rd = ro[sel].contiguous().to(cuda), rd[sel].contiguous().to(cuda)
    ha = torch.zeros(4096, 4, device=cuda)
    out, g_ref = _host_sized(model, ro, rd, ha)
    ref = {kk: out["rendered"][kk].detach().clone() for kk in KEYS}
    del out
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    frame = StaticFrame(model, 4096, loss_fn=_loss, near=0.01, zero_grads=True)
    frame.step(ro, rd, ha)
    assert frame.coherent is False and frame.counts()["overflow"] == 0
    for kk in KEYS:
        assert torch.equal(frame.rendered[kk], ref[kk]), kk
    g = product_grads(model)
    for kk, v in g_ref.items():
        if v is not None:
            assert rel_l2(g[kk], v) <= 2e-5, (kk, rel_l2(g[kk], v))


def test_graph_follows_parameter_and_occupancy_updates(cuda):
    """between replays the optimiser changes the parameters in place and the EMA changes the occupancy grid (in place here, re-assigned by the
    reference's own EMA code): the replayed graph must render what an eager run of the UPDATED model renders"""
    from neuralsim_b200.graphics.neus_static import StaticFrame, render_static
    _, model = make_pair(cuda)
    model.train()
    ro, rd = _rays(cuda, k=2)
    ha = torch.zeros(ro.shape[0], 4, device=cuda)
    model.ctrl_var.start_it, model.ctrl_var.stop_it, model.ctrl_var.final_inv_s = 0, 100, 400.0        # the schedule's constants (config)
    model.ctrl_var.set_iter(0)
    frame = StaticFrame(model, ro.shape[0], near=0.01, slack=3.0)
    with torch.no_grad():
        frame.step(ro, rd, ha)
        before = frame.rendered["rgb_volume"].clone()
        model.radiance_net.blocks.layers[2].bias.add_(0.3)                      # an optimiser step (in place)
        model.implicit_surface.encoding.flattened_params.mul_(1.01)
        model.ctrl_var.set_iter(37)                                            # the variance schedule moved on (a HOST-side weight in the reference)
        g = model.accel.occ.occ_grid.clone()
        g[:, :, :32] = False                                                    # the EMA carved half of the grid away ...
        model.accel.occ.occ_grid = g                                            # ... and RE-ASSIGNED the buffer, as the reference does
        frame.step(ro, rd, ha)
        want, _, _ = render_static(model, ro, rd, ha, near=0.01, march_cap=frame.march_cap, kept_cap=frame.kept_cap, coherent=frame.coherent)
        assert frame.captures == 1 and frame.counts()["overflow"] == 0
        assert not torch.equal(frame.rendered["rgb_volume"], before)
        for kk in KEYS:           # (1 - w) is rounded once on the host in the eager run and evaluated in fp32 on the device in the graph: <= 1 ulp of inv_s
            assert torch.allclose(frame.rendered[kk], want[kk], rtol=1e-5, atol=1e-6), kk

"""neuralsim_b200.adapter.accelerate: the fused path under a model object that exposes the reference's attribute names (a stand-in: the
reference's own classes cannot be imported here), sharing its parameter / buffer objects."""
import pytest
import torch

from oracle import scene as oscene
from util import make_pair

pytestmark = pytest.mark.gpu


class _RefLike(torch.nn.Module):
    """only the attributes `NeusRendererMixin + LoTDNeuS` objects have (lotd_neus.py:27-232, renderer_mixin.py:40-135); no methods of ours"""

    def __init__(self, m):
        super().__init__()
        self.implicit_surface, self.radiance_net, self.ctrl_var, self.accel, self.space = m.implicit_surface, m.radiance_net, m.ctrl_var, m.accel, m.space
        self.ray_query_cfg = dict(m.ray_query_cfg)
        self.upsample_s_divisor, self.max_level, self.it = 1.0, None, 0


def test_accelerate_shares_parameters_and_renders(cuda):
    from neuralsim_b200.adapter import accelerate
    from neuralsim_b200.renderer import SingleVolumeRenderer
    _, src = make_pair(cuda)
    ref = _RefLike(src).train()
    ours = accelerate(ref)
    assert ours.implicit_surface.encoding.flattened_params is ref.implicit_surface.encoding.flattened_params
    assert ours.radiance_net.blocks.layers[2].weight is ref.radiance_net.blocks.layers[2].weight
    assert ours.ctrl_var.ln_inv_s is ref.ctrl_var.ln_inv_s and ours.accel.occ.occ_grid is ref.accel.occ.occ_grid
    assert {k for k, _ in ours.named_parameters()} == {k for k, _ in src.named_parameters()}
    ro, rd = oscene.pinhole_rays(30, 40, oscene.orbit_camera(1, 8))
    ro, rd, ha = ro.to(cuda), rd.to(cuda), torch.zeros(1200, 4, device=cuda)
    r = SingleVolumeRenderer(dict(near=0.01)).train()
    want = r.render(src, ro, rd, rays_h_appear=ha)["rendered"]
    sum(v.mean() for v in want.values()).backward()
    g_want = src.implicit_surface.encoding.flattened_params.grad.clone()
    src.zero_grad(set_to_none=True)
    # the renderer's call sequence on the patched reference-like object (single_volume_renderer.py:222-246): ray_test -> ray_query
    rt = ref.ray_test(ro, rd, near=0.01, rays_h_appear=ha)
    raw = ref.ray_query(ray_tested=rt, config=dict(with_rgb=True, with_normal=True), return_buffer=True)
    vb = raw["volume_buffer"]
    assert vb["type"] == "packed" and vb["rgb"].requires_grad
    from neuralsim_b200.fields.neus import volume_integration
    vb["nablas_in_world"] = vb["nablas"]
    rendered = dict(mask_volume=torch.zeros(1200, device=cuda), depth_volume=torch.zeros(1200, device=cuda), rgb_volume=torch.zeros(1200, 3, device=cuda),
                    normals_volume=torch.zeros(1200, 3, device=cuda))
    got = volume_integration(vb, rendered, training=True, nablas_key="nablas_in_world")
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=0, atol=1e-6), k
    sum(v.mean() for v in got.values()).backward()
    g = ref.implicit_surface.encoding.flattened_params.grad                 # the gradient landed in the REFERENCE's parameter object
    assert g is not None and float((g - g_want).norm() / g_want.norm()) < 1e-4
    # the reference re-assigns occ_grid on every EMA update: the patched methods pick the new tensor up
    ref.accel.occ.occ_grid = torch.zeros_like(ref.accel.occ.occ_grid)
    raw2 = ref.ray_query(ray_tested=ref.ray_test(ro, rd, near=0.01, rays_h_appear=ha), config=dict(), return_buffer=True)
    assert ours.accel.occ.occ_grid is ref.accel.occ.occ_grid

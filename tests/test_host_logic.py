"""CPU: host-side mirror of the reference interface -- gin configs, Rays container, learning-rate schedule, colour and
normal helpers.  Where /root/reference exists (the build container) the values are compared LIVE with the unmodified
reference; the intrinsic checks run everywhere."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import ref_import
from refnerf_pl_b200 import configs, image, ref_utils, train_utils, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GINS = ['blender_refnerf.gin', 'llff_refnerf.gin', 'llff_refnerf_geometry_losses.gin']
live = pytest.mark.skipif(not ref_import.available(), reason='reference tree not present')


def _our_config(gin):
    configs.clear_bindings()
    configs.parse_gin_files_and_bindings([os.path.join(ROOT, 'configs', gin)])
    return configs.Config()


@pytest.mark.parametrize('gin', GINS)
def test_gin_files_parse(gin):
    cfg = _our_config(gin)
    assert cfg.data_loss_mult == 1.0 and cfg.orientation_loss_target == 'normals_pred'
    assert (cfg.near, cfg.far) == ((2, 6) if gin.startswith('blender') else (0, 1))


@live
@pytest.mark.parametrize('gin', GINS)
def test_gin_bindings_match_reference(gin):
    """Every Config field the shipped gin file binds has the same value here as in the reference, and the Model /
    NerfMLP bindings of the two files are identical."""
    ns, ref_cfg = ref_import.load(gin)
    ours = _our_config(gin)
    ref_fields = {f.name for f in dataclasses.fields(ref_cfg)}
    for k in ns.bindings.get('Config', {}):
        if k in ref_fields and hasattr(ours, k):
            assert getattr(ours, k) == getattr(ref_cfg, k), (gin, k, getattr(ours, k), getattr(ref_cfg, k))
    mine = ref_import.parse_gin_file(os.path.join(ROOT, 'configs', gin))
    for scope in ('Model', 'NerfMLP', 'PropMLP', 'MLP'):
        assert mine.get(scope, {}) == ns.bindings.get(scope, {}), (gin, scope)


def test_rays_container():
    r = utils.Rays(**{k: np.zeros((4, 1, 1, c), np.float32) for k, c in
                      (('origins', 3), ('directions', 3), ('viewdirs', 3), ('radii', 1), ('imageplane', 2),
                       ('lossmult', 1), ('near', 1), ('far', 1))}, cam_idx=np.zeros((4, 1, 1, 1), np.int32))
    r.to('cpu')
    assert isinstance(r.origins, torch.Tensor) and r.origins.dtype == torch.float32 and r.cam_idx.dtype == torch.int32
    assert r.shape == (4, 1, 1, 3) and r[1:3].origins.shape == (2, 1, 1, 3)
    assert r.reshape(4, -1).radii.shape == (4, 1)
    with pytest.raises(ValueError):
        r['x']
    bad = utils.Rays(**{f.name: [0.0] for f in dataclasses.fields(utils.Rays)})
    with pytest.raises(ValueError):
        bad.to('cpu')
    d = utils.dummy_rays()
    assert d.origins.shape == (1, 3) and d.cam_idx.dtype == torch.int32
    merged = utils.merge_chunks([{'a': torch.ones(2, 3), 'b': [torch.zeros(2)]}, {'a': torch.ones(1, 3), 'b': [torch.zeros(1)]}])
    assert merged['a'].shape == (3, 3) and merged['b'][0].shape == (3,)


def test_learning_rate_schedule():
    cfg = _our_config('blender_refnerf.gin')
    f = lambda s: train_utils.learning_rate_decay(s, cfg.lr_init, cfg.lr_final, cfg.max_steps, cfg.lr_delay_steps,
                                                  cfg.lr_delay_mult)
    assert abs(f(cfg.max_steps) - cfg.lr_final / cfg.lr_init) < 1e-12          # multiplier reaches lr_final / lr_init
    assert f(0) == pytest.approx(cfg.lr_delay_mult if cfg.lr_delay_steps > 0 else 1.0)
    vals = [f(s) for s in range(cfg.lr_delay_steps, cfg.max_steps, max(1, cfg.max_steps // 50))]
    assert all(a >= b for a, b in zip(vals, vals[1:]))                           # log-linear decay after the warm-up
    if ref_import.available():
        ns, _ = ref_import.load('blender_refnerf.gin')
        for s in (0, 1, 100, cfg.lr_delay_steps, cfg.max_steps // 3, cfg.max_steps):
            ref = ns.math.learning_rate_decay(s, cfg.lr_init, cfg.lr_final, cfg.max_steps, cfg.lr_delay_steps, cfg.lr_delay_mult)
            assert f(s) * cfg.lr_init == pytest.approx(ref, rel=1e-12) or f(s) == pytest.approx(ref, rel=1e-12)


def test_colour_and_normal_helpers():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1000, 3, generator=g)
    assert float((image.srgb_to_linear(image.linear_to_srgb(x)) - x).abs().max()) < 2e-6
    assert float(image.linear_to_srgb(torch.tensor([0.0031308]))) == pytest.approx(323 / 25 * 0.0031308, rel=1e-6)
    v = torch.randn(200, 3, generator=g)
    n = ref_utils.l2_normalize(v)
    assert float((n.norm(dim=-1) - 1).abs().max()) < 1e-6
    assert float(ref_utils.l2_normalize(torch.zeros(1, 3)).abs().max()) == 0.0
    d = ref_utils.l2_normalize(torch.randn(200, 3, generator=g))
    r = ref_utils.reflect(d, n)
    assert float((r.norm(dim=-1) - 1).abs().max()) < 1e-5                         # reflection preserves length
    assert float(((r * n).sum(-1) - (d * n).sum(-1)).abs().max()) < 1e-5          # and the normal component
    w = torch.rand(200, generator=g)
    assert float(ref_utils.compute_weighted_mae(w, n, n)) < 0.05                  # arccos(1 - eps) in degrees
    assert float(ref_utils.compute_weighted_mae(w, n, -n)) == pytest.approx(180.0, abs=0.05)
    if ref_import.available():
        ns, _ = ref_import.load('blender_refnerf.gin')
        assert torch.equal(image.linear_to_srgb(x), ns.image.linear_to_srgb(x))
        assert torch.equal(image.srgb_to_linear(x), ns.image.srgb_to_linear(x))
        assert torch.equal(n, ns.ref_utils.l2_normalize(v))
        assert torch.equal(r, ns.ref_utils.reflect(d, n))
        m = ref_utils.l2_normalize(torch.randn(200, 3, generator=g))
        assert float(ref_utils.compute_weighted_mae(w, n, m)) == pytest.approx(float(ns.ref_utils.compute_weighted_mae(w, n, m)), rel=1e-6)


def test_geometry_losses_and_noisy_rays_match_reference_fixture():
    """configs/llff_refnerf_geometry_losses.gin glue (sample_utils.py:40-80, train_utils.py:207-325) on the renderings
    the unmodified reference produced (tests/golden/llff_geom_losses.npz, oracle/make_golden.py geometry): the noisy
    rays and every rendering-level loss term must reproduce the reference's values."""
    import os
    import numpy as np
    import torch
    from refnerf_pl_b200 import configs, sample_utils, train_utils, utils
    from tests._cases import GOLDEN, RAY_KEYS
    g = np.load(os.path.join(GOLDEN, 'llff_geom_losses.npz'))
    root = os.path.dirname(GOLDEN.rstrip('/'))
    configs.clear_bindings()
    configs.parse_gin_files_and_bindings([os.path.join(os.path.dirname(root), 'configs', 'llff_refnerf_geometry_losses.gin')])
    cfg = configs.Config()
    assert cfg.sample_noise_angles == 4 and cfg.consistency_diffuse_loss_type == 'var'
    rays = utils.Rays(**{k: torch.tensor(g['rays_' + k]) for k in RAY_KEYS})
    rend = [{k: torch.tensor(g[f'rend{l}_{k}']) for k in ('rgb', 'diffuse', 'specular', 'distance', 'acc', 'normals', 'normals_pred')}
            for l in range(2)]
    rend_n = [{k: torch.tensor(g[f'noisy_rend{l}_{k}']) for k in ('rgb', 'diffuse', 'specular', 'distance', 'acc', 'normals', 'normals_pred')}
              for l in range(2)]
    warm = float(g['warmup_ratio'])
    noisy = sample_utils.sample_noisy_rays(rays, rend[-1], cfg.sample_angle_range, cfg.sample_noise_size,
                                           cfg.sample_noise_angles, warm, xyz_angles=g['xyz_angles'])
    for k in ('origins', 'directions', 'viewdirs'):
        assert np.abs(getattr(noisy, k).numpy() - g['noisy_' + k]).max() <= 1e-6, k
    assert noisy.origins.shape == (512, 3) and noisy.near.shape == (512, 1)

    class M:
        num_levels = 2
    close = lambda a, key: abs(float(a) - float(g[key])) <= 2e-5 * abs(float(g[key])) + 1e-12
    d, s, n = train_utils.noisy_consistency_loss(M, rend, rend_n, cfg, warm)
    assert close(d, 'loss_diffuse_consistency') and close(s, 'loss_specular_consistency') and close(n, 'loss_normals_consistency')
    assert close(train_utils.noisy_distance_consistency_loss(M, rays, noisy, rend, rend_n, cfg, warm), 'loss_distance_consistency')
    assert close(train_utils.accumulated_weights_loss(rend, cfg), 'loss_acc')
    gt = torch.tensor(g['gt_rgb'])
    assert close(train_utils.compute_data_loss(gt, rend, rays.lossmult, cfg)[0], 'loss_data')
    # warm-up schedule of nerf_system.py:97-113
    assert train_utils.consistency_warmup_ratio(cfg, 0) == 0.0
    assert abs(train_utils.consistency_warmup_ratio(cfg, 75000) - 0.5) < 1e-12
    assert train_utils.consistency_warmup_ratio(cfg, 200000) == 1.0


def test_contract_follows_reference_formula():
    """coord.contract / inv_contract (coord.py:20-35; the reference function itself raises TypeError)."""
    import torch
    from oracle import refnerf_oracle as O
    from refnerf_pl_b200 import coord
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 3, generator=g) * 3
    z = coord.contract(x)
    assert torch.equal(z, O.contract(x))
    inside = (x ** 2).sum(-1) <= 1
    assert torch.equal(z[inside], x[inside]) and float(z.norm(dim=-1).max()) < 2.0
    assert float((coord.inv_contract(z) - x).abs().max()) < 1e-3 * float(x.abs().max())


def test_collect_param_stats_matches_reference_definition():
    """train_utils.collect_param_stats == the per-parameter loop of nerf_system.py:212-217 (weights_l2s, grad_norms,
    grad_maxes), computed as three multi-tensor reductions without host round trips."""
    import torch
    from refnerf_pl_b200 import train_utils

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(5, 7)
            self.b = torch.nn.Linear(7, 3)

    torch.manual_seed(0)
    m = Tiny()
    m.b(m.a(torch.randn(11, 5))).square().sum().backward()
    m.b.bias.grad = None                       # a parameter without gradient
    st = train_utils.collect_param_stats(m)
    params = dict(m.named_parameters())
    for k, p in params.items():
        kk = k.replace('.', '/')
        assert torch.allclose(st['weights_l2s'][kk], p.detach().norm() ** 2)
        if p.grad is None:
            assert float(st['grad_norms'][kk]) == 0.0 and float(st['grad_maxes'][kk]) == 0.0
        else:
            assert torch.allclose(st['grad_norms'][kk], p.grad.norm())
            assert torch.allclose(st['grad_maxes'][kk], p.grad.abs().max())

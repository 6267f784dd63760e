"""-m gpu: the NerfMLP / Model drop-in against fixtures produced by the unmodified reference
(tests/golden/*.npz) and against the CPU oracle, in every GEMM precision mode.

Tolerances (BASELINE.json north_star): per-sample density / rgb within 1e-3 relative, composited
rgb / distance / acc within 1e-3 absolute, parameter gradients within 1e-2 relative (norm-wise per
tensor).  Density-gradient normals are heavy-tailed even under an fp32 re-ordering of the reference
itself, so they are gated on mean and 99th percentile, never on max.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import refnerf_oracle as O
from tests._cases import CASES, case_params, load_case
from tests._gpu import DEV, build_model, load_params, norm_rel, rays_obj, rel_err

pytestmark = pytest.mark.gpu

# Limits per precision mode.  "rel" = |a-b| / (|b| + 1e-3 max|b|), 99.9th percentile over samples
# (bounded heavy tails: a ReLU flip moves single samples); "abs" = max abs error.
TOL = {
    #          density/roughness rel | colours, tint, weights abs | unit normals_pred abs | composited abs | normals | grads
    'fp32':   dict(rel=2e-4, colour=2e-5, npred=2e-4, comp=2e-5, normals_mean=2e-3, normals_p99=5e-2, grad=5e-3),
    'bf16x3': dict(rel=1e-3, colour=1e-4, npred=1e-3, comp=1e-4, normals_mean=5e-3, normals_p99=1e-1, grad=1e-2),
    'bf16':   dict(rel=1e-2, colour=1e-3, npred=2e-2, comp=1e-3, normals_mean=1e-1, normals_p99=2.0, grad=1e-1),
    # fp16 operands (11-bit): measured <= 1.3e-4 / 2.9e-5 / 2e-3 / 1.2e-5 / 4.3e-3 / 0.10 / 3.5e-2 on the three fixtures
    'fp16':   dict(rel=1e-3, colour=1e-4, npred=5e-3, comp=1e-4, normals_mean=1e-2, normals_p99=0.3, grad=5e-2),
}
# Trained-scale weights (fixture blender_trained: 2 000 steps on the sphere scene, sharp density field).  The parity
# modes keep the north-star gates; 11-bit operands do not at the tail of the distribution (softplus of a raw density
# that swings over +-50 across the surface amplifies an activation error of 2^-12): measured against bf16x3 on 4 096
# rays, fp16 per-sample density p99.9 1.2e-2 relative, composited rgb mean 5e-5 / p99 6e-4 / max 8.5e-3.
# Against the reference fixture (32 rays): composited rgb fp32 3e-7 / bf16x3 2.5e-6 / fp16 1.4e-4 / bf16 8e-4; level-0
# density p99.9 5e-6 / 1e-4 / 5e-3 / 6e-2; gradients 7e-3 (rgb.bias, a cancellation-dominated sum) / 6e-3 / 6e-2 / 0.65.
TOL_TRAINED = {
    'fp32':   dict(TOL['fp32'], grad=1e-2),
    'fp16':   dict(rel=2e-2, colour=5e-3, npred=2e-3, comp=1e-3, normals_mean=1e-2, normals_p99=0.5, grad=2e-1),
    'bf16':   dict(rel=5e-1, colour=5e-2, npred=2e-1, comp=2e-1, normals_mean=1e-1, normals_p99=2.0, grad=1.0),
}
NAMES = ['blender_init', 'blender_pert', 'llff_geom', 'blender_trained']


def _tol(name, precision):
    if name == 'blender_trained' and precision in TOL_TRAINED:
        return TOL_TRAINED[precision]
    return TOL[precision]


ALL = bool(os.environ.get('RN_PARITY_ALL'))   # report-only runs: also try plain bf16 on the trained-scale fixtures
LEVEL1 = 8.0   # level-1 samples sit on fenceposts resampled from level-0 weights: errors compound


def _run(name, precision, mode):
    from tests._gpu import GIN_FOR_CASE
    g, rays = load_case(name)
    model, cfg = build_model(precision, gin=GIN_FOR_CASE[name])
    load_params(model, case_params(g))
    model.train(mode == 'train')
    r = rays_obj(dict(rays))
    with (torch.enable_grad() if mode == 'train' else torch.no_grad()):
        rend, hist = model(r, 1.0, True)
    return g, model, cfg, r, rend, hist


def _report(key, rep):
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/parity_report.jsonl', 'a') as f:
        f.write(json.dumps({'case': key, **rep}) + '\n')


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16', 'fp16'])
@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_model_vs_reference_fixture(name, precision, mode):
    if precision == 'bf16' and name != 'blender_init' and not ALL:
        pytest.skip('plain bf16 is the throughput mode: gated at reference init only (SURVEY 7.4)')
    g, model, cfg, r, rend, hist = _run(name, precision, mode)
    tol = _tol(name, precision)
    rep, bad = {}, []

    def gate(key, val, lim):
        rep[key] = float(val)
        if not val <= lim:
            bad.append((key, float(val), lim))

    for lvl in range(2):
        f = LEVEL1 if lvl == 1 else 1.0
        sd = hist[lvl]['sdist'].cpu().numpy()
        ref_sd = g[f'{mode}_hist{lvl}_sdist']
        if lvl == 0:
            assert np.array_equal(sd, ref_sd)          # level-0 fenceposts: input independent, bit-exact
        else:
            loose = precision == 'bf16' or (name == 'blender_trained' and precision == 'fp16')
            gate('sdist1_abs', np.abs(sd - ref_sd).max(), 2e-2 if loose else 2e-3)
        for k in ('density', 'roughness'):
            e = rel_err(hist[lvl][k].detach().cpu().numpy(), g[f'{mode}_hist{lvl}_{k}'])
            gate(f'{k}{lvl}_rel_p999', np.quantile(e, 0.999), tol['rel'] * f)
        for k in ('rgb', 'diffuse', 'specular', 'tint', 'weights'):
            e = np.abs(hist[lvl][k].detach().cpu().numpy() - g[f'{mode}_hist{lvl}_{k}'])
            gate(f'{k}{lvl}_abs_p999', np.quantile(e, 0.999), tol['colour'] * f)
        e = np.abs(hist[lvl]['normals_pred'].detach().cpu().numpy() - g[f'{mode}_hist{lvl}_normals_pred'])
        gate(f'normals_pred{lvl}_abs_p999', np.quantile(e, 0.999), tol['npred'] * f)
        if mode == 'train':
            e = np.abs(hist[lvl]['normals'].cpu().numpy() - g[f'{mode}_hist{lvl}_normals'])
            gate(f'normals{lvl}_mean', e.mean(), tol['normals_mean'] * f)
            gate(f'normals{lvl}_p99', np.quantile(e, 0.99), tol['normals_p99'])
        else:
            assert hist[lvl]['normals'] is None
        for k in ('rgb', 'diffuse', 'specular', 'distance', 'acc', 'normals_pred', 'tint', 'roughness', 'distance_mean',
                  'distance_percentile_5', 'distance_median', 'distance_percentile_95'):
            a = rend[lvl][k].detach().cpu().numpy()
            b = g[f'{mode}_rend{lvl}_{k}']
            assert a.shape == b.shape and a.dtype == b.dtype, (k, a.shape, b.shape, a.dtype, b.dtype)
            depthlike = 'distance' in k
            lim = tol['npred'] if k == 'normals_pred' else tol['comp'] * (50 if depthlike else 1)
            err = np.abs(a - b)
            if k == 'distance_mean':
                # sum(w log t) / max(eps, acc): for rays that hit nothing the weights are rounding noise of
                # 1 - exp(-sigma delta) (0 or 2^-24 depending on the last bit of exp), so the reference's own value is
                # noise there (render.py:231-238); compared where the ray has accumulated something
                err = err[g[f'{mode}_rend{lvl}_acc'] >= 1e-3]
            gate(f'rend_{k}{lvl}', err.max() if err.size else 0.0, lim * f)
        for k in ('ray_sdist', 'ray_weights', 'ray_rgbs'):
            assert tuple(rend[lvl][k].shape) == g[f'{mode}_rend{lvl}_{k}'].shape
    _report(f'{name}/{precision}/{mode}', rep)
    assert not bad, bad


@pytest.mark.parametrize('precision', ['fp32', 'bf16x3', 'bf16', 'fp16'])
@pytest.mark.parametrize('name', NAMES)
def test_gradients_vs_reference_fixture(name, precision):
    if precision == 'bf16' and name != 'blender_init' and not ALL:
        pytest.skip('plain bf16 is gated at reference init only')
    from refnerf_pl_b200 import train_utils
    g, model, cfg, r, rend, hist = _run(name, precision, 'train')
    gt = torch.tensor(g['gt_rgb'], device=DEV)
    loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
    assert abs(float(loss.detach()) - float(g['train_loss'])) <= 1e-3 * abs(float(g['train_loss'])) + 1e-6
    loss.backward()
    tol = _tol(name, precision)['grad']
    rep = {}
    for kname, p in model.nerf_mlp.named_parameters():
        ref_norm = float(g['grad_norm_' + kname])
        gn = float(p.grad.double().norm())
        sub = p.grad.reshape(-1)[::97].cpu().numpy()
        ref_sub = g['grad_sub_' + kname]
        e_norm = abs(gn - ref_norm) / max(ref_norm, 1e-30)
        e_sub = float(np.linalg.norm(sub - ref_sub) / max(np.linalg.norm(ref_sub), 1e-30))
        rep[kname] = max(e_norm, e_sub)
    _report(f'{name}/{precision}/grad', rep)
    bad = {k: v for k, v in rep.items() if v > tol}
    assert not bad, bad


def test_fused_chain_matches_per_layer_gemms():
    """bf16: the fused forward chain issues the same MMAs in the same order as the per-layer kernels, so the
    two implementations must agree bit for bit (eval) -- any layout / barrier bug shows up here."""
    from refnerf_pl_b200 import synthetic
    from tests._gpu import rays_obj as _ro
    p = O.init_params(seed=4, bias_std=0.1, weight_scale=1.3)
    rays = synthetic.blender_rays(700, seed=9)          # 89600 rows per level: several tiles per SM + a ragged chunk
    outs = {}
    for impl in (0, 2):
        model, _ = build_model('bf16', mlp_kwargs=dict(gemm_impl=impl, chunk_rows=65536))
        load_params(model, p)
        for mode in ('eval', 'train'):
            model.train(mode == 'train')
            with torch.no_grad():
                rend, hist = model(_ro(rays), 1.0, False)
            outs[(impl, mode)] = (rend, hist)
    for mode in ('eval', 'train'):
        (ra, ha), (rb, hb) = outs[(0, mode)], outs[(2, mode)]
        for lvl in range(2):
            for k in ('density', 'rgb', 'diffuse', 'specular', 'roughness', 'normals_pred'):
                assert torch.equal(ha[lvl][k], hb[lvl][k]), (mode, lvl, k, float((ha[lvl][k] - hb[lvl][k]).abs().max()))
            if mode == 'train':
                assert torch.equal(ha[lvl]['normals'], hb[lvl]['normals'])
            assert torch.equal(ra[lvl]['rgb'], rb[lvl]['rgb'])


def test_fused_backward_chain_matches_per_layer_gemms():
    """bf16 backward: fused dgrad chains vs per-layer kernels (same arithmetic; only the wgrad atomics reorder)."""
    from refnerf_pl_b200 import synthetic, train_utils
    p = O.init_params(seed=5, bias_std=0.1, weight_scale=1.2)
    rays = synthetic.blender_rays(600, seed=10)
    gt = torch.tensor(synthetic.gt_rgb(600, 10), device=DEV)
    grads = {}
    for impl in (0, 2):
        model, cfg = build_model('bf16', mlp_kwargs=dict(gemm_impl=impl, chunk_rows=32768))
        load_params(model, p)
        model.train(True)
        r = rays_obj(rays)
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()
        grads[impl] = {k: v.grad.clone() for k, v in model.nerf_mlp.named_parameters()}
    for k in grads[0]:
        a, b = grads[0][k].double(), grads[2][k].double()
        assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-12, (k, float((a - b).norm() / b.norm()))


def test_saved_activation_backward_matches_recompute(monkeypatch):
    """Training keeps the activations of the forward for the backward; above ops.SAVED_BYTES_CAP the backward
    recomputes them chunk by chunk instead.  Same arithmetic either way (only the wgrad atomics reorder)."""
    from refnerf_pl_b200 import ops, synthetic, train_utils
    p = O.init_params(seed=6, bias_std=0.1, weight_scale=1.2)
    rays = synthetic.blender_rays(520, seed=11)
    gt = torch.tensor(synthetic.gt_rgb(520, 11), device=DEV)
    default_cap = ops.SAVED_BYTES_CAP
    for prec in ('bf16', 'bf16x3', 'fp16'):
        grads = {}
        for cap in (default_cap, 0):
            monkeypatch.setattr(ops, 'SAVED_BYTES_CAP', cap)
            model, cfg = build_model(prec, mlp_kwargs=dict(chunk_rows=16384))   # several chunks, ragged tail
            load_params(model, p)
            model.train(True)
            r = rays_obj(rays)
            rend, hist = model(r, 1.0, True)
            loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
            loss.backward()
            grads[cap] = {k: v.grad.clone() for k, v in model.nerf_mlp.named_parameters()}
        for k in grads[0]:
            a, b = grads[default_cap][k].double(), grads[0][k].double()
            assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-12, (prec, k, float((a - b).norm() / b.norm()))


def test_checkpoint_names_match_reference():
    model, _ = build_model('fp32')
    sd = model.state_dict()
    assert len(sd) == 92
    names = {k for k in sd if k.startswith('nerf_mlp.')}
    assert {'nerf_mlp.' + k for k in O.param_shapes()} == {k.rsplit('.', 1)[0] for k in names}
    for k, shp in O.param_shapes().items():
        assert tuple(sd[f'nerf_mlp.{k}.weight'].shape) == shp
        assert tuple(sd[f'prop_mlp.{k}.bias'].shape) == (shp[0],)


def test_leading_dims_and_numpy_rays():
    """Lightning hands rays as [B,1,1,C] numpy arrays (datasets.py:467-477)."""
    from refnerf_pl_b200 import synthetic, utils
    model, _ = build_model('bf16x3')
    r = synthetic.blender_rays(6, seed=2)
    rays4 = utils.Rays(**{k: v.reshape(6, 1, 1, -1) for k, v in r.items()})
    model.eval()
    with torch.no_grad():
        rend, hist = model(rays4, 1.0, False)
    assert rend[1]['rgb'].shape == (6, 1, 1, 3) and rend[1]['acc'].shape == (6, 1, 1)
    assert hist[1]['density'].shape == (6, 1, 1, 128) and hist[1]['sdist'].shape == (6, 1, 1, 129)
    assert 'normals_pred' not in rend[1]


@pytest.mark.parametrize('n_rays', [1, 3, 5, 257])
def test_tiny_and_ragged_batches_match_per_layer_path(n_rays):
    """Ray counts far below one 512-row super tile and not a multiple of anything: the fused chains (TMA out-of-bounds
    fill / clipping, masked tails) must still agree bit for bit with the per-layer kernels, forward and normals pass,
    and produce finite gradients."""
    from refnerf_pl_b200 import synthetic, train_utils
    p = O.init_params(seed=8, bias_std=0.1, weight_scale=1.1)
    rays = synthetic.blender_rays(n_rays, seed=21)
    gt = torch.tensor(synthetic.gt_rgb(n_rays, 21), device=DEV)
    outs = {}
    for impl in (0, 2):
        model, cfg = build_model('bf16', mlp_kwargs=dict(gemm_impl=impl))
        load_params(model, p)
        model.train(True)
        r = rays_obj(rays)
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()
        g = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()])
        assert torch.isfinite(g).all()
        outs[impl] = (rend, hist, g)
    (ra, ha, ga), (rb, hb, gb) = outs[0], outs[2]
    for lvl in range(2):
        for k in ('density', 'rgb', 'normals', 'normals_pred', 'roughness'):
            assert torch.equal(ha[lvl][k], hb[lvl][k]), (n_rays, lvl, k)
        assert torch.equal(ra[lvl]['rgb'], rb[lvl]['rgb'])
    assert float((ga - gb).norm()) <= 1e-4 * float(gb.norm()) + 1e-12


@pytest.mark.parametrize('n_rays', [1, 5, 130])
def test_fp16_tiny_batches_track_parity_mode(n_rays):
    """fp16 chains on ray counts below one super tile: the dynamic gradient scaling (amax / scale / rescale kernels on a
    handful of rows, possibly all-zero seeds) must give finite gradients that track the parity mode."""
    from refnerf_pl_b200 import synthetic, train_utils
    p = O.init_params(seed=9, bias_std=0.1, weight_scale=1.1)
    rays = synthetic.blender_rays(n_rays, seed=23)
    gt = torch.tensor(synthetic.gt_rgb(n_rays, 23), device=DEV)
    res = {}
    for prec in ('bf16x3', 'fp16'):
        model, cfg = build_model(prec)
        load_params(model, p)
        model.train(True)
        r = rays_obj(rays)
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()
        g = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()]).double()
        assert torch.isfinite(g).all(), prec
        res[prec] = (float(loss.detach()), g, rend[1]['rgb'].detach())
    (l3, g3, c3), (l16, g16, c16) = res['bf16x3'], res['fp16']
    assert abs(l16 - l3) <= 1e-3 * abs(l3) + 1e-6
    assert float((c16 - c3).abs().max()) <= 1e-3
    assert float((g16 - g3).norm()) <= 5e-2 * float(g3.norm()) + 1e-12


def test_geometry_config_training_losses_vs_reference_fixture():
    """BASELINE config 5 (configs/llff_refnerf_geometry_losses.gin) as one training step through this package:
    main forward, device-side noisy rays, second forward, data + orientation + predicted-normal + consistency
    (diffuse / specular / normals, 'var') + distance-consistency + acc + weights-entropy losses, backward -- against
    the unmodified reference's values and gradients (tests/golden/llff_geom_losses.npz)."""
    from refnerf_pl_b200 import train_utils
    from tests._cases import GOLDEN, RAY_KEYS
    g = np.load(os.path.join(GOLDEN, 'llff_geom_losses.npz'))
    model, cfg = build_model('bf16x3', gin='llff_refnerf_geometry_losses.gin')
    load_params(model, case_params(g))
    model.train(True)
    r = rays_obj({k: torch.tensor(g['rays_' + k]) for k in RAY_KEYS})
    gt = torch.tensor(g['gt_rgb'], device=DEV)
    warm = float(g['warmup_ratio'])
    step = int(round(warm * cfg.consistency_warmup_steps * cfg.max_steps))
    assert abs(train_utils.consistency_warmup_ratio(cfg, step) - warm) < 1e-9
    loss, losses, stats, rend, hist = train_utils.training_losses(model, r, gt, cfg, 1.0, step, xyz_angles=g['xyz_angles'])
    got = {k: float(v) for k, v in losses.items()}
    got['orientation+predicted'] = got.pop('normals')
    ref = {k[5:]: float(g[k]) for k in g.files if k.startswith('loss_')}
    ref['orientation+predicted'] = ref.pop('orientation') + ref.pop('predicted_normals')
    rep = {}
    for k, v in ref.items():
        rep[k] = abs(got[k] - v) / (abs(v) + 1e-3 * abs(float(g['train_loss'])) * 1e-3)
        assert abs(got[k] - v) <= 2e-3 * abs(v) + 1e-9, (k, got[k], v)
    assert abs(float(loss) - float(g['train_loss'])) <= 1e-3 * abs(float(g['train_loss']))
    loss.backward()
    bad = {}
    for kname, p in model.nerf_mlp.named_parameters():
        ref_norm = float(g['grad_norm_' + kname])
        e_norm = abs(float(p.grad.double().norm()) - ref_norm) / max(ref_norm, 1e-30)
        sub, ref_sub = p.grad.reshape(-1)[::97].cpu().numpy(), g['grad_sub_' + kname]
        e_sub = float(np.linalg.norm(sub - ref_sub) / max(np.linalg.norm(ref_sub), 1e-30))
        rep['grad_' + kname] = max(e_norm, e_sub)
        if max(e_norm, e_sub) > 1e-2:
            bad[kname] = max(e_norm, e_sub)
    _report('llff_geom_losses/bf16x3/step', rep)
    assert not bad, bad


@pytest.mark.parametrize('variant', ['anneal', 'dilate', 'anneal+dilate', 'levels3'])
def test_model_options_off_the_shipped_configs_match_oracle(variant):
    """Model options every shipped gin leaves at their neutral value (models.py:168-203): the annealed
    resampling logits (anneal_slope > 0, train_frac < 1), max-dilated proposal weights (dilation_multiplier /
    dilation_bias > 0, the rn_max_dilate_weights kernel + the 382-bin resampler) and num_levels != 2, against the CPU
    oracle on the same rays and weights."""
    from refnerf_pl_b200 import synthetic
    mk = {}
    if 'anneal' in variant:
        mk['anneal_slope'] = 10.0
    if 'dilate' in variant:
        mk.update(dilation_multiplier=0.5, dilation_bias=0.0025)
    if variant == 'levels3':
        mk['num_levels'] = 3
    train_frac = 0.3
    p = O.init_params(seed=12, bias_std=0.1, weight_scale=1.6)
    rays = synthetic.blender_rays(40, seed=31)
    model, _ = build_model('bf16x3', model_kwargs=mk)
    load_params(model, p)
    model.eval()
    with torch.no_grad():
        rend, hist = model(rays_obj(rays), train_frac, True)
    rt = {k: torch.tensor(v) for k, v in rays.items()}
    with torch.no_grad():
        orend, ohist = O.model_forward(p, rt, train_frac, True, False, model_cfg=mk)
    assert len(rend) == len(orend) == mk.get('num_levels', 2)
    for lvl in range(len(rend)):
        sd, osd = hist[lvl]['sdist'].cpu(), ohist[lvl]['sdist']
        if lvl == 0:
            assert torch.equal(sd, osd)
        else:
            assert float((sd - osd).abs().max()) <= 2e-3, (variant, lvl)
        assert float((rend[lvl]['rgb'].cpu() - orend[lvl]['rgb']).abs().max()) <= 1e-3, (variant, lvl)
        assert float((rend[lvl]['acc'].cpu() - orend[lvl]['acc']).abs().max()) <= 1e-3, (variant, lvl)


def test_render_image_graph_and_eager_match_oracle():
    """models.render_image (models.py:763-825): a small frame rendered in chunks -- CUDA-graph replay over a fixed chunk
    with direct writes into the [H,W,.] buffers, a ragged last chunk, and the plain eager loop -- all give the same
    frame, and that frame matches the CPU oracle evaluated on every ray at once."""
    from refnerf_pl_b200 import models, synthetic, utils
    h, w = 20, 24
    p = O.init_params(seed=13, bias_std=0.1, weight_scale=1.4)
    rays = synthetic.blender_rays(h * w, seed=41)
    model, cfg = build_model('bf16x3')
    load_params(model, p)
    model.eval()
    frame = utils.Rays(**{k: torch.tensor(v).reshape(h, w, -1) for k, v in rays.items()}).to(DEV)
    fn = lambda r: model(r, 1.0, True)
    outs = {}
    with torch.no_grad():
        for name, chunk, graph in (('graph', 96, True), ('ragged', 100, True), ('eager', 96, False), ('one', 4096, False)):
            cfg.render_chunk_size = chunk
            outs[name] = models.render_image(fn, frame, cfg, use_graph=graph)
    ref = outs['one']
    assert ref['rgb'].shape == (h, w, 3) and ref['acc'].shape == (h, w) and ref['distance_median'].dtype == torch.float64
    for name in ('graph', 'ragged', 'eager'):
        for k in ('rgb', 'diffuse', 'specular', 'distance', 'acc', 'normals_pred', 'roughness', 'distance_median'):
            assert torch.equal(outs[name][k], ref[k]), (name, k)
        assert len(outs[name]['ray_sdist']) == 2 and outs[name]['ray_sdist'][0].shape[-1] == 129
    rt = {k: torch.tensor(v) for k, v in rays.items()}
    with torch.no_grad():
        orend, _ = O.model_forward(p, rt, 1.0, True, False)
    for k in ('rgb', 'acc', 'distance'):
        a = ref[k].reshape(h * w, -1).cpu()
        b = orend[-1][k].reshape(h * w, -1)
        assert float((a - b).abs().max()) <= (1e-3 if k != 'distance' else 5e-3), k


def test_training_mode_forward_under_no_grad_keeps_no_activations():
    """a training-mode forward under torch.no_grad() (e.g. a validation pass that wants the density-gradient
    normals) must not allocate the saved-activation region; packed weights follow load_state_dict / .to()."""
    from refnerf_pl_b200 import synthetic
    model, _ = build_model('bf16x3')
    model.train(True)
    r = rays_obj(synthetic.blender_rays(2048, seed=3))
    model(r, 1.0, False)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        rend, hist = model(r, 1.0, False)
    torch.cuda.synchronize()
    peak_nograd = torch.cuda.max_memory_allocated() - base
    assert hist[0]['normals'] is not None
    torch.cuda.reset_peak_memory_stats()
    rend, hist = model(r, 1.0, False)
    torch.cuda.synchronize()
    peak_grad = torch.cuda.max_memory_allocated() - base
    assert peak_nograd < 0.7 * peak_grad, (peak_nograd, peak_grad)
    before = rend[1]['rgb'].detach().clone()
    sd = {k: v * 1.5 for k, v in model.nerf_mlp.state_dict().items()}
    model.nerf_mlp.load_state_dict(sd)
    with torch.no_grad():
        rend2, _ = model(r, 1.0, False)
    assert float((rend2[1]['rgb'] - before).abs().max()) > 1e-4   # new weights are in use


def test_deterministic_wgrad_is_bit_reproducible_and_agrees():
    """NerfMLP.deterministic_wgrad: weight / bias gradients reduced over the row-split CTAs in a fixed order (per-CTA
    partial tiles + one reduction pass) instead of with floating-point atomics.  Two identical steps give bit-identical
    gradients, and they agree with the atomic path to fp32 summation noise."""
    from refnerf_pl_b200 import synthetic, train_utils
    p = O.init_params(seed=15, bias_std=0.1, weight_scale=1.2)
    rays = synthetic.blender_rays(3000, seed=51)       # several CTAs' worth of rows per wgrad launch
    gt = torch.tensor(synthetic.gt_rgb(3000, 51), device=DEV)
    grads = {}
    for key, det in (('det_a', True), ('det_b', True), ('atomic', False)):
        model, cfg = build_model('bf16x3', mlp_kwargs=dict(deterministic_wgrad=det))
        load_params(model, p)
        model.train(True)
        r = rays_obj(rays)
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()
        grads[key] = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()]).clone()
    assert torch.equal(grads['det_a'], grads['det_b'])
    a, b = grads['det_a'].double(), grads['atomic'].double()
    assert float((a - b).norm()) <= 1e-5 * float(b.norm())


def test_fp16_gradient_saturation_is_reported():
    """the fp16 mode scales each dgrad chain by ONE power of two chosen from its seed tile; a gradient
    that grows > 16x through the chain is clamped by the saturating conversion.  The clamp must not be silent:
    ops.fp16_saturation_count() stays 0 on a healthy step and fires when the later layers amplify the gradient."""
    from refnerf_pl_b200 import ops, synthetic, train_utils
    rays = synthetic.blender_rays(600, seed=61)
    gt = torch.tensor(synthetic.gt_rgb(600, 61), device=DEV)

    def step(p):
        model, cfg = build_model('fp16')
        load_params(model, p)
        model.train(True)
        r = rays_obj(rays)
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()
        torch.cuda.synchronize()

    ops.fp16_saturation_count(reset=True)
    step(O.init_params(seed=16, bias_std=0.1, weight_scale=1.2))
    assert ops.fp16_saturation_count(reset=True) == 0
    p = O.init_params(seed=16, bias_std=0.0, weight_scale=1.0)
    for i in range(1, 8):
        p[f'viewdir_mlp.{i}.weight'] = p[f'viewdir_mlp.{i}.weight'] * 10.0     # the backward grows ~4x per layer
    step(p)
    assert ops.fp16_saturation_count(reset=True) > 0
    assert ops.fp16_saturation_count(reset=False) == 0


def test_parameter_gradients_of_a_step_tile_one_flat_buffer():
    """ops.param_carrier: both levels' fused calls share one autograd edge to the parameters, so a step's parameter
    gradients are views of ONE flat buffer (their sum over the calls) -- what parallel.GradAllReducer reduces in place --
    and they equal the gradients of per-call carriers (stand-alone MLP calls accumulate per parameter)."""
    from refnerf_pl_b200 import parallel, synthetic, train_utils
    p = O.init_params(seed=4, bias_std=0.1, weight_scale=1.2)
    rays = synthetic.blender_rays(200, seed=9)
    gt = torch.tensor(synthetic.gt_rgb(200, 9), device=DEV)
    model, cfg = build_model('bf16x3', mlp_kwargs=dict(deterministic_wgrad=True))
    load_params(model, p)
    model.train(True)
    r = rays_obj(rays)

    def step():
        for q in model.nerf_mlp.parameters():
            q.grad = None
        rend, hist = model(r, 1.0, True)
        loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
        loss.backward()

    step()
    red = parallel.GradAllReducer(model.nerf_mlp.parameters())
    one = red._grads_as_one_buffer()
    assert one is not None and one.numel() == sum(q.numel() for q in model.nerf_mlp.parameters())
    shared = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()]).clone()
    assert model.nerf_mlp.__dict__.get('_carrier') is None          # closed again after the level loop
    # the same step with one carrier per MLP call (per-parameter accumulation by autograd)
    orig = type(model.nerf_mlp).param_carrier
    try:
        type(model.nerf_mlp).param_carrier = lambda self: ops_mod().param_carrier(self.ordered_params())
        step()
    finally:
        type(model.nerf_mlp).param_carrier = orig
    separate = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()])
    assert float((shared.double() - separate.double()).norm()) <= 1e-6 * float(separate.double().norm())


def ops_mod():
    from refnerf_pl_b200 import ops
    return ops


def test_shared_memory_chain_variant_still_matches():
    """`RN_X3_TS=0` selects chain_x3.cu (activation tile in shared memory) instead of the default chain_x3t.cu (tensor
    memory).  The switch is read once per process, so the fused-chain and trained-fixture parity tests are re-run in a
    child process with it set."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RN_X3_TS='0')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_model.py'), '-q', '-m', 'gpu', '-x',
                        '-k', 'fused or (fixture and trained and bf16x3) or tiny_and_ragged'], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout

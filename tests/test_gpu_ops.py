"""-m gpu parity tests of the non-MLP kernels, through the C ABI (ctypes -> librefnerf_b200.so)."""
import numpy as np
import pytest
import torch

from oracle import refnerf_oracle as O
from tests._cases import GOLDEN
from tests._gpu import DEV

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from refnerf_pl_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLDEN + '/ops.npz')


def _resample(ops, t, w, padding=0.01, s_out=128):
    n = t.shape[0]
    near = torch.zeros(n, 1, device=DEV)
    far = torch.ones(n, 1, device=DEV)
    so, to, cw, idx = ops.resample(t.to(DEV), w.to(DEV), near, far, s_out, padding, 1.0, 0.0, 1.0, True)
    return so.cpu(), to.cpu(), cw.cpu(), idx.cpu()


def test_resample_golden(ops, gold):
    t, w = torch.tensor(gold['rs_t']), torch.tensor(gold['rs_w'])
    so, to, cw, idx = _resample(ops, t, w)
    # CDF within a few ulp of the reference's (softmax/exp/cumsum order differ between any two backends)
    assert np.abs(cw.numpy() - gold['rs_cw']).max() <= 4e-7
    # interval indices: bit-exact w.r.t. the index definition on the kernel's own CDF ...
    u = O.sample_grid(128).expand(t.shape[0], 128)
    assert torch.equal(idx.long(), O.interval_index(u, cw))
    # ... and the interpolated fenceposts are bit-exact given that CDF (no FMA contraction in the kernel)
    centers = O.sorted_interp(u, cw, t)
    mid = (centers[..., 1:] + centers[..., :-1]) / 2
    first = torch.clamp(2 * centers[..., :1] - mid[..., :1], min=0.0)
    last = torch.clamp(2 * centers[..., -1:] - mid[..., -1:], max=1.0)
    assert torch.equal(so, torch.cat([first, mid, last], -1))
    assert torch.equal(to, so)  # near=0, far=1
    # against the reference's own outputs: the CDF differs only through the softmax (CUDA expf vs the CPU's vectorised
    # exp, order of the normalising sum; the cumulative sum itself is fp64 on both sides), so indices can differ only
    # where u lies within those few ulps of a CDF knot.  The measured rate is logged.
    mism = float((idx.numpy() != gold['rs_idx']).mean())
    cw_ulp = float(np.abs(cw.numpy() - gold['rs_cw']).max() / 2.0 ** -24)
    import json, os
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/parity_report.jsonl', 'a') as f:
        f.write(json.dumps({'case': 'resample_golden', 'index_mismatch_rate_vs_reference': mism,
                            'mismatching_indices': int((idx.numpy() != gold['rs_idx']).sum()), 'indices': int(idx.numel()),
                            'cdf_max_abs_diff_in_ulp_of_1': cw_ulp,
                            'sdist_max_abs_diff': float(np.abs(so.numpy() - gold['rs_sdist']).max())}) + '\n')
    assert mism <= 1e-4, mism   # measured on the B200: 0 of 12 288 (profiles/r02_parity_report.jsonl)
    assert np.abs(so.numpy() - gold['rs_sdist']).max() <= 1e-5   # few-ulp CDF differences / narrow CDF steps


def test_resample_search_stage_bit_exact_on_reference_cdf(ops, gold):
    """north_star: "resampled interval indices must match bit-exactly".  Fed with the REFERENCE's CDF (gold rs_cw, from
    the unmodified reference), the kernel's search + interpolation + fencepost stage must reproduce the reference's
    interval indices and resampled fenceposts bit for bit -- including zero-weight runs and duplicate fenceposts."""
    t, cw = torch.tensor(gold['rs_t']), torch.tensor(gold['rs_cw'])
    n = t.shape[0]
    near, far = torch.zeros(n, 1, device=DEV), torch.ones(n, 1, device=DEV)
    so, to, idx = ops.resample_from_cdf(t.to(DEV), cw.to(DEV), near, far, 128)
    assert np.array_equal(idx.cpu().numpy(), gold['rs_idx'])
    assert np.array_equal(so.cpu().numpy(), gold['rs_sdist'])
    # generic (s_in != 128) kernel on the same data: resample 128 bins to 64 samples, against the oracle on that CDF
    so64, _, idx64 = ops.resample_from_cdf(t.to(DEV), cw.to(DEV), near, far, 64)
    u64 = O.sample_grid(64).expand(n, 64)
    assert torch.equal(idx64.cpu().long(), O.interval_index(u64, cw))
    c = O.sorted_interp(u64, cw, t)
    mid = (c[..., 1:] + c[..., :-1]) / 2
    ref = torch.cat([torch.clamp(2 * c[..., :1] - mid[..., :1], min=0.0), mid, torch.clamp(2 * c[..., -1:] - mid[..., -1:], max=1.0)], -1)
    assert torch.equal(so64.cpu(), ref)


def test_resample_level0_constant(ops, gold):
    t = torch.tensor([[0.0, 1.0]]).repeat(5, 1)
    w = torch.ones(5, 1)
    so, _, _, _ = _resample(ops, t, w)
    assert np.array_equal(so.numpy(), np.repeat(gold['rs_level0'], 5, 0))   # bit-exact incl. sdist[0]=4.66e-10


def test_resample_large_properties(ops):
    g = torch.Generator().manual_seed(0)
    n = 100_000
    t = torch.sort(torch.rand(n, 129, generator=g), -1).values
    w = torch.rand(n, 128, generator=g) ** 8
    w[::7, 20:100] = 0
    w[::11] = 0                                   # all-zero rows: uniform via the padding
    so, to, cw, idx = _resample(ops, t, w)
    assert torch.isfinite(so).all()
    assert (so[:, 1:] >= so[:, :-1]).all() and so.min() >= 0 and so.max() <= 1
    assert (cw[:, 1:] >= cw[:, :-1]).all() and (cw[:, 0] == 0).all() and (cw[:, -1] == 1).all()
    u = O.sample_grid(128).expand(n, 128)
    assert torch.equal(idx.long(), O.interval_index(u, cw))
    ref = O.sample_intervals(t[:4096], O.resample_logits(t[:4096], w[:4096], 0.01), 128)
    assert (so[:4096] - ref).abs().max() <= 2e-5   # CDF ulps divided by narrow CDF steps


def test_encode_matches_oracle(ops, gold):
    td = torch.tensor(gold['ipe_tdist'])
    o, d, r = (torch.tensor(gold['ipe_' + k]) for k in ('origins', 'directions', 'radii'))
    enc = ops.encode(td.to(DEV), o.to(DEV), d.to(DEV), r.to(DEV)).cpu().numpy()
    ref = gold['ipe_enc']
    # identical fp32 op order for the phase => only sinf/expf ulp differences remain
    assert np.abs(enc - ref).max() <= 2e-6, np.abs(enc - ref).max()


def test_encode_blender_rays(ops):
    from refnerf_pl_b200 import synthetic
    r = synthetic.blender_rays(512, seed=4)
    rt = {k: torch.tensor(v) for k, v in r.items()}
    sd = O.sample_intervals(torch.tensor([[0., 1.]]).repeat(512, 1), torch.zeros(512, 1), 128)
    td = O.s_to_t(sd, rt['near'], rt['far'])
    means, cov = O.cast_rays(td, rt['origins'], rt['directions'], rt['radii'])
    lm, lv = O.lift_and_diagonalize(means, cov, torch.tensor(O.octahedron_basis()))
    ref = O.integrated_pos_enc(lm, lv, 0, 16)
    enc = ops.encode(td.to(DEV), rt['origins'].to(DEV), rt['directions'].to(DEV), rt['radii'].to(DEV)).cpu()
    assert (enc - ref).abs().max() <= 2e-6


def test_ide_matches_reference_within_its_noise(ops, gold):
    d = torch.tensor(gold['ide_dirs'])
    ls = np.array([l for i in range(5) for _ in range(2 ** i + 1) for l in [2 ** i]])
    hi_band = np.concatenate([ls == 16, ls == 16])
    for i, k in enumerate(gold['ide_kappa_inv']):
        out = ops.ide(d.to(DEV), torch.full((256,), float(k), device=DEV)).cpu().numpy()
        ref32 = gold[f'ide_enc_{i}']
        ref64 = O.integrated_dir_enc(d.double(), torch.full((256, 1), float(k), dtype=torch.float64)).numpy()
        # tight against exact arithmetic on the reference's (fp32-rounded) coefficients
        assert np.abs(out - ref64).max() <= 2e-6
        # against the fp32 reference: l<=8 tight; l=16 band within the reference's own fp32 noise
        assert np.abs(out - ref32)[:, ~hi_band].max() <= 5e-5
        assert np.abs(out - ref32)[:, hi_band].max() <= max(2e-2 * np.exp(-136 * float(k)), 1e-6)


def _composite_inputs(n=64, s=128, seed=0, extras=True):
    g = torch.Generator().manual_seed(seed)
    dens = torch.rand(n, s, generator=g) * 4
    td = torch.sort(torch.rand(n, s + 1, generator=g) * 4 + 2, -1).values
    dirs = torch.randn(n, 3, generator=g)
    far = torch.full((n, 1), 6.0)
    v3 = lambda: torch.rand(n, s, 3, generator=g)
    return dict(density=dens, tdist=td, dirs=dirs, far=far, rgb=v3(), diffuse=v3(), specular=v3(), normals=v3() - 0.5,
                normals_pred=v3() - 0.5, roughness=torch.rand(n, s, 1, generator=g), tint=v3())


def _oracle_composite(x, extras):
    w = O.compute_alpha_weights(x['density'], x['tdist'], x['dirs'])
    ex = {k: x[k] for k in ('normals', 'normals_pred', 'roughness', 'tint')}
    return w, O.volumetric_rendering(x['rgb'], x['diffuse'], x['specular'], w, x['tdist'], 1.0, x['far'], extras, ex)


@pytest.mark.parametrize('extras', [False, True])
def test_composite_forward_backward(ops, extras):
    x = _composite_inputs()
    leaf = ('density', 'rgb', 'diffuse', 'specular', 'normals_pred', 'roughness', 'tint')
    xc = {k: v.clone().requires_grad_(k in leaf) for k, v in x.items()}
    xg = {k: v.to(DEV).requires_grad_(k in leaf) for k, v in x.items()}
    w_ref, r_ref = _oracle_composite(xc, extras)
    w, comp, ex, pct = ops.composite_fwd(xg['density'], xg['tdist'], xg['dirs'], xg['far'], xg['rgb'], xg['diffuse'],
                                         xg['specular'], xg['normals'], xg['normals_pred'], xg['roughness'], xg['tint'],
                                         1.0, extras)
    tol = 2e-6
    assert (w.cpu() - w_ref).abs().max() <= tol
    assert (comp[:, 0:3].cpu() - r_ref['rgb']).abs().max() <= 5e-6
    assert (comp[:, 3:6].cpu() - r_ref['diffuse']).abs().max() <= 5e-6
    assert (comp[:, 6:9].cpu() - r_ref['specular']).abs().max() <= 5e-6
    assert (comp[:, 9:10].cpu() - r_ref['distance']).abs().max() <= 2e-5
    assert (comp[:, 10].cpu() - r_ref['acc']).abs().max() <= 5e-6
    gen = torch.Generator().manual_seed(5)
    cw = torch.randn(w_ref.shape, generator=gen)
    terms_ref = [(w_ref * cw).sum(), (r_ref['rgb'] * 1.3).sum(), (r_ref['diffuse'] * 0.7).sum(),
                 (r_ref['specular'] * -0.4).sum(), (r_ref['distance'] * 0.2).sum(), (r_ref['acc'] * 0.9).sum()]
    terms = [(w * cw.to(DEV)).sum(), (comp[:, 0:3] * 1.3).sum(), (comp[:, 3:6] * 0.7).sum(), (comp[:, 6:9] * -0.4).sum(),
             (comp[:, 9] * 0.2).sum(), (comp[:, 10] * 0.9).sum()]
    if extras:
        assert (ex[:, 0:3].cpu() - r_ref['normals']).abs().max() <= 5e-6
        assert (ex[:, 3:6].cpu() - r_ref['normals_pred']).abs().max() <= 5e-6
        assert (ex[:, 6:9].cpu() - r_ref['tint']).abs().max() <= 5e-6
        assert (ex[:, 9:10].cpu() - r_ref['roughness']).abs().max() <= 5e-6
        assert (comp[:, 11].cpu() - r_ref['distance_mean']).abs().max() <= 2e-5
        for i, k in enumerate(('distance_percentile_5', 'distance_median', 'distance_percentile_95')):
            assert pct.dtype == torch.float64
            assert (pct[:, i].cpu() - r_ref[k]).abs().max() <= 1e-4, k
        terms_ref += [(r_ref['normals'] * 0.3).sum(), (r_ref['normals_pred'] * -0.6).sum(), (r_ref['tint'] * 0.5).sum(),
                      (r_ref['roughness'] * 1.1).sum()]
        terms += [(ex[:, 0:3] * 0.3).sum(), (ex[:, 3:6] * -0.6).sum(), (ex[:, 6:9] * 0.5).sum(), (ex[:, 9] * 1.1).sum()]
    sum(terms_ref).backward()
    sum(terms).backward()
    for k in leaf:
        if xc[k].grad is None:
            continue
        a, b = xg[k].grad.cpu(), xc[k].grad
        assert (a - b).abs().max() <= 1e-4 * max(1.0, float(b.abs().max())), k


def test_step_losses(ops, gold):
    t, w, te, we = (torch.tensor(gold[k]) for k in ('lo_t', 'lo_w', 'lo_tenv', 'lo_wenv'))
    out = ops.lossfun_outer(t.to(DEV), w.to(DEV), te.to(DEV), we.to(DEV).requires_grad_(True))
    assert np.abs(out.detach().cpu().numpy() - gold['lo_loss']).max() <= 1e-6
    we_c = we.clone().requires_grad_(True)
    cot = torch.rand(out.shape, generator=torch.Generator().manual_seed(1))
    (O.lossfun_outer(t, w, te, we_c) * cot).sum().backward()
    we_g = we.to(DEV).requires_grad_(True)
    (ops.lossfun_outer(t.to(DEV), w.to(DEV), te.to(DEV), we_g) * cot.to(DEV)).sum().backward()
    assert (we_g.grad.cpu() - we_c.grad).abs().max() <= 1e-4 * max(1.0, float(we_c.grad.abs().max()))
    d = ops.distortion(t.to(DEV), w.to(DEV))
    assert np.abs(d.cpu().numpy() - gold['dist_loss']).max() <= 1e-6
    w_c = w.clone().requires_grad_(True)
    O.lossfun_distortion(t, w_c).sum().backward()
    w_g = w.to(DEV).requires_grad_(True)
    ops.distortion(t.to(DEV), w_g).sum().backward()
    assert (w_g.grad.cpu() - w_c.grad).abs().max() <= 1e-5


@pytest.mark.parametrize('n,s', [(37, 128), (5, 7), (1, 1), (300, 64)])
@pytest.mark.parametrize('target', ['normals_pred', 'normals'])
def test_normal_losses_match_oracle(ops, n, s, target):
    """rn_normal_losses_* vs the oracle's orientation_loss / predicted_normal_loss (train_utils.py:165-204) on the same
    inputs: per-level loss values and the gradients w.r.t. weights and normals_pred."""
    from types import SimpleNamespace
    from refnerf_pl_b200 import train_utils
    g = torch.Generator().manual_seed(n * 131 + s)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    hist_c, hist_g = [], []
    for lvl in range(2):
        w = torch.rand(n, s, generator=g) / s
        nrm = unit(torch.randn(n, s, 3, generator=g))
        npred = unit(torch.randn(n, s, 3, generator=g))
        hist_c.append(dict(weights=w.clone().requires_grad_(True), normals=nrm, normals_pred=npred.clone().requires_grad_(True)))
        hist_g.append(dict(weights=w.to(DEV).requires_grad_(True), normals=nrm.to(DEV),
                           normals_pred=npred.to(DEV).requires_grad_(True)))
    vd = unit(torch.randn(n, 3, generator=g))
    cfg = dict(O.DEFAULT_LOSS_CFG)
    if target == 'normals_pred':
        ref = O.orientation_loss(hist_c, vd, cfg) + O.predicted_normal_loss(hist_c, cfg)
    else:   # the oracle restates the configs' target; the 'normals' target is the same formula on the other tensor
        ref = O.orientation_loss([dict(h, normals_pred=h['normals']) for h in hist_c], vd, cfg) + O.predicted_normal_loss(hist_c, cfg)
    ref.backward()
    c = SimpleNamespace(orientation_loss_target=target, **{k: v for k, v in cfg.items()})
    out = train_utils.normal_losses(vd.to(DEV), 2, hist_g, c)
    out.backward()
    assert abs(float(out) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    for hc, hg in zip(hist_c, hist_g):
        for k in ('weights', 'normals_pred'):
            a, b = hg[k].grad.cpu(), hc[k].grad
            assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1e-12), (k, float((a - b).abs().max()))
    # and against the unfused torch path of this package on the GPU
    for h in hist_g:
        h['weights'].grad = None
        h['normals_pred'].grad = None
    unf = train_utils.orientation_loss(vd.to(DEV), 2, hist_g, c) + train_utils.predicted_normal_loss(2, hist_g, c)
    assert abs(float(unf) - float(out)) <= 2e-6 * max(1.0, abs(float(out)))


def test_cpu_tensors_fail_loudly(ops):
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.distortion(torch.rand(2, 5), torch.rand(2, 4))


def test_pixels_to_rays_matches_reference_golden(golden_dir):
    """rn_pixels_to_rays (float64 arithmetic, fp32 outputs) vs the reference camera_utils.pixels_to_rays golden
    vectors, Blender-shaped and LLFF-NDC cameras: equal after rounding to fp32 (<= 1 ulp where the FP64 sum order
    of the 3x3 products differs)."""
    import os
    from refnerf_pl_b200 import camera_utils
    g = np.load(os.path.join(golden_dir, 'raygen.npz'))
    for name, ndc in (('blender', False), ('llff', True)):
        px, py, ci = (torch.tensor(g[f'{name}_{k}']).to(DEV) for k in ('px', 'py', 'cam'))
        p2c, c2w = torch.tensor(g[f'{name}_pixtocams']).to(DEV), torch.tensor(g[f'{name}_camtoworlds']).to(DEV)
        out = camera_utils.pixels_to_rays(px, py, p2c, c2w, pixtocam_ndc=p2c[0] if ndc else None, cam_idx=ci)
        for k, v in zip(('origins', 'directions', 'viewdirs', 'radii', 'imageplane'), out):
            ref = g[f'{name}_{k}'].astype(np.float32)
            got = v.cpu().numpy()
            assert got.shape == ref.shape, (name, k, got.shape, ref.shape)
            # 1 fp32 ulp of the value, or of the field's scale where the NDC projection cancels towards zero
            tol = 1.2e-7 * np.maximum(np.abs(ref), np.abs(ref).max())
            assert (np.abs(got - ref) <= tol).all(), (name, k, float((np.abs(got - ref) / tol).max()))
        # the reference's calling convention (per-pixel matrices) gives the same result as the camera-table form
        out2 = camera_utils.pixels_to_rays(px, py, p2c[ci.long()], c2w[ci.long()], pixtocam_ndc=p2c[0] if ndc else None)
        for a, b in zip(out, out2):
            assert torch.equal(a, b)


def test_max_dilate_weights_golden_and_oracle(ops, gold):
    """stepfun.max_dilate_weights (stepfun.py:102-131): fenceposts bit-exact, weights within 1e-7 of the unmodified
    reference (ops.npz dil_t / dil_w); the trimmed form equals the [1:-1] slice (models.py:186-187); random rays with
    duplicate fenceposts, zero weights and bins far narrower than the dilation against the oracle."""
    t, w = torch.tensor(gold['lo_t']), torch.tensor(gold['lo_w'])
    td, wd = ops.max_dilate_weights(t.to(DEV), w.to(DEV), 0.01, 0.0, 1.0, True, False)
    assert np.array_equal(td.cpu().numpy(), gold['dil_t'])
    assert np.abs(wd.cpu().numpy() - gold['dil_w']).max() <= 1e-7
    tt, wt = ops.max_dilate_weights(t.to(DEV), w.to(DEV), 0.01, 0.0, 1.0, True, True)
    assert torch.equal(tt, td[:, 1:-1]) and torch.equal(wt, wd[:, 1:-1])
    g = torch.Generator().manual_seed(5)
    for s, dil in ((128, 0.0064), (128, 0.05), (37, 0.002), (1, 0.1)):
        n = 301
        t = torch.sort(torch.rand(n, s + 1, generator=g) ** 3, -1).values      # crowded near 0: many bins inside +-dilation
        if s > 8:
            t[::3, 4:8] = t[::3, 4:5]                                         # duplicate fenceposts
        w = torch.rand(n, s, generator=g)
        w[1::4] = 0.0
        for renorm in (True, False):
            rt, rw = O.max_dilate_weights(t, w, dil, domain=(0.0, 1.0), renormalize=renorm)
            kt, kw = ops.max_dilate_weights(t.to(DEV), w.to(DEV), dil, 0.0, 1.0, renorm, False)
            assert torch.equal(kt.cpu(), rt), (s, dil)
            scale = max(1.0, float(rw.abs().max()))
            assert float((kw.cpu() - rw).abs().max()) <= 2e-7 * scale, (s, dil, renorm)


@pytest.mark.parametrize('kind', ['mse', 'charb'])
@pytest.mark.parametrize('n,multiscale', [(1, True), (37, True), (4096, True), (300, False)])
def test_data_loss_epilogue_matches_reference_formula(ops, kind, n, multiscale):
    """rn_data_loss_fwd / _bwd (one launch per level) vs train_utils.compute_data_loss's torch formula
    (reference train_utils.py:33-88) in fp64: loss, the 'mses' statistic and d loss / d rgb of both levels."""
    from types import SimpleNamespace
    from refnerf_pl_b200 import train_utils
    g = torch.Generator().manual_seed(n + len(kind))
    gt = torch.rand(n, 3, generator=g)
    lm = torch.rand(n, 1, generator=g) + 0.25
    rgbs = [torch.rand(n, 3, generator=g) for _ in range(2)]
    cfg = SimpleNamespace(supervised_by_linear_rgb=False, disable_multiscale_loss=not multiscale, data_loss_type=kind,
                          charb_padding=0.001, data_coarse_loss_mult=0.5, data_loss_mult=1.0)
    # fp64 statement of the reference formula
    ref_in = [r.double().requires_grad_(True) for r in rgbs]
    lmb = torch.broadcast_to(lm.double(), gt.shape) if multiscale else torch.ones_like(gt, dtype=torch.float64)
    terms, mses = [], []
    for r in ref_in:
        r2 = (r - gt.double()) ** 2
        mses.append((lmb * r2).sum() / lmb.sum())
        terms.append((lmb * (r2 if kind == 'mse' else torch.sqrt(r2 + cfg.charb_padding ** 2))).sum() / lmb.sum())
    ref_loss = cfg.data_coarse_loss_mult * terms[0] + cfg.data_loss_mult * terms[1]
    ref_loss.backward()
    dev_in = [r.to(DEV).requires_grad_(True) for r in rgbs]
    loss, stats = train_utils.compute_data_loss(gt.to(DEV), [{'rgb': r} for r in dev_in], lm.to(DEV), cfg)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 2e-6 * max(1.0, abs(float(ref_loss)))
    assert np.abs(stats['mses'].cpu().numpy() - np.array([float(m) for m in mses])).max() <= 2e-6
    for a, b in zip(dev_in, ref_in):
        assert (a.grad.cpu().double() - b.grad).abs().max() <= 2e-6 * max(1e-3, float(b.grad.abs().max()))
    # bit-reproducible (fixed summation order)
    loss2, _ = train_utils.compute_data_loss(gt.to(DEV), [{'rgb': r.detach()} for r in dev_in], lm.to(DEV), cfg)
    assert float(loss2) == float(loss.detach())

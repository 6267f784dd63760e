"""Pins oracle/refnerf_oracle.py against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only; travels to the GPU box."""
import numpy as np
import pytest
import torch

from oracle import refnerf_oracle as O
from tests._cases import CASES, GOLDEN, case_params, load_case


def _close(a, b, tol, what):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= tol, f'{what}: max abs err {err} > {tol}'


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_model_matches_reference_fixture(name, mode):
    g, rays = load_case(name)
    mcfg, mlpcfg, losscfg = CASES[name]
    p = case_params(g)
    training = mode == 'train'
    p = {k: v.clone().requires_grad_(training) for k, v in p.items()}
    with (torch.enable_grad() if training else torch.no_grad()):
        rend, hist = O.model_forward(p, rays, 1.0, True, training, mcfg, mlpcfg)
    for lvl in range(2):
        for k, v in rend[lvl].items():
            _close(v.detach(), g[f'{mode}_rend{lvl}_{k}'], 2e-6, f'{name}/{mode}/rend{lvl}/{k}')
        for k, v in hist[lvl].items():
            if v is None:
                assert f'{mode}_hist{lvl}_{k}' not in g.files
                continue
            tol = 2e-6 if k != 'normals' else 1e-4   # density-gradient normals: heavy-tailed
            _close(v.detach(), g[f'{mode}_hist{lvl}_{k}'], tol, f'{name}/{mode}/hist{lvl}/{k}')
    if training:
        loss = O.total_loss(rend, hist, rays, torch.tensor(g['gt_rgb']), losscfg)
        _close(loss.detach(), g['train_loss'], 1e-6, 'loss')
        loss.backward()
        for k, v in p.items():
            gn = float(v.grad.double().norm())
            ref = float(g['grad_norm_' + k])
            assert abs(gn - ref) <= 1e-4 * ref + 1e-12, (k, gn, ref)
            _close(v.grad.reshape(-1)[::97], g['grad_sub_' + k], 1e-5 * max(ref, 1e-8) + 1e-9, 'grad ' + k)


def test_resample_ops_fixture():
    g = np.load(GOLDEN + '/ops.npz')
    t, w = torch.tensor(g['rs_t']), torch.tensor(g['rs_w'])
    logits = O.resample_logits(t, w, 0.01)
    sd, aux = O.sample_intervals(t, logits, 128, return_aux=True)
    assert np.array_equal(aux['cw'].numpy(), g['rs_cw'])
    assert np.array_equal(aux['idx'].numpy().astype(np.int32), g['rs_idx'])      # bit-exact interval indices (D11)
    assert np.array_equal(aux['centers'].numpy(), g['rs_centers'])               # gather form == dense mask form
    assert np.array_equal(sd.numpy(), g['rs_sdist'])
    sd0 = O.sample_intervals(torch.tensor([[0.0, 1.0]]), torch.log(torch.ones(1, 1) + 0.01), 128)
    assert np.array_equal(sd0.numpy(), g['rs_level0'])
    assert abs(float(sd0[0, 0]) - 4.6566e-10) < 1e-13 and float(sd0[0, -1]) == np.float32(0.99999988)


def test_encodings_fixture():
    g = np.load(GOLDEN + '/ops.npz')
    means, cov = O.cast_rays(torch.tensor(g['ipe_tdist']), torch.tensor(g['ipe_origins']),
                             torch.tensor(g['ipe_directions']), torch.tensor(g['ipe_radii']))
    _close(means, g['ipe_means'], 1e-6, 'means')
    _close(cov, g['ipe_cov'], 1e-9, 'cov')
    lm, lv = O.lift_and_diagonalize(means, cov, torch.tensor(O.octahedron_basis()))
    _close(O.integrated_pos_enc(lm, lv, 0, 16), g['ipe_enc'], 1e-6, 'ipe')
    d = torch.tensor(g['ide_dirs'])
    for i, k in enumerate(g['ide_kappa_inv']):
        _close(O.integrated_dir_enc(d, torch.full((256, 1), float(k))), g[f'ide_enc_{i}'], 1e-6, f'ide{i}')


def test_composite_and_losses_fixture():
    g = np.load(GOLDEN + '/ops.npz')
    w = O.compute_alpha_weights(torch.tensor(g['cw_density']), torch.tensor(g['cw_tdist']), torch.tensor(g['cw_dirs']))
    _close(w, g['cw_weights'], 1e-7, 'alpha weights')
    t, wq, te, we = (torch.tensor(g[k]) for k in ('lo_t', 'lo_w', 'lo_tenv', 'lo_wenv'))
    _close(O.lossfun_outer(t, wq, te, we), g['lo_loss'], 1e-7, 'lossfun_outer')
    _close(O.lossfun_distortion(t, wq), g['dist_loss'], 1e-7, 'distortion')
    td, wd = O.max_dilate_weights(t, wq, 0.01, domain=(0.0, 1.0), renormalize=True)
    _close(td, g['dil_t'], 0, 'dilate t')
    _close(wd, g['dil_w'], 1e-7, 'dilate w')

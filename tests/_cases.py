"""Shared description of the golden cases (tests/golden/*.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# name -> (oracle model cfg, mlp cfg, loss cfg); mirrors the gin files the fixtures were made with
CASES = {
    'blender_init': (dict(), dict(), dict()),
    'blender_pert': (dict(), dict(), dict()),
    # configs/llff_refnerf_geometry_losses.gin: NerfMLP.srgb_mapping=False, render-time 'norm_linear',
    # predicted-normal mults 3e-5 / 3e-6
    'llff_geom': (dict(srgb_mapping_render='norm_linear'), dict(srgb_mapping=False),
                  dict(predicted_normal_loss_mult=3e-5, predicted_normal_coarse_loss_mult=3e-6)),
    # weights after 2 000 training steps on the analytic sphere scene (tools/train_parity.py, parity mode), stored in
    # tests/golden/trained_sphere_params.npz: sharp density field, trained-scale activations
    'blender_trained': (dict(), dict(), dict()),
}

RAY_KEYS = ('origins', 'directions', 'viewdirs', 'radii', 'imageplane', 'lossmult', 'near', 'far', 'cam_idx')


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    rays = {k: torch.tensor(g['rays_' + k]) for k in RAY_KEYS}
    return g, rays


def case_params(g):
    from oracle import refnerf_oracle as O
    if 'meta_params_file' in g.files:
        with np.load(os.path.join(GOLDEN, str(g['meta_params_file']))) as f:
            p = {k: torch.tensor(f[k]) for k in f.files}
        chk = sum(float(v.double().abs().sum()) for v in p.values())
        assert abs(chk - float(g['param_checksum'])) < 1e-9 * chk, 'trained parameter file does not match the fixture'
        return p
    p = O.init_params(seed=int(g['meta_seed']), bias_std=float(g['meta_bias_std']),
                      weight_scale=float(g['meta_weight_scale']))
    chk = sum(float(v.double().abs().sum()) for v in p.values())
    assert abs(chk - float(g['param_checksum'])) < 1e-6 * chk, 'RNG drift: init_params no longer reproduces fixture weights'
    return p

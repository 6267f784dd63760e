"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_golden.py
The fixtures pin `oracle/refnerf_oracle.py` (tests/test_oracle_golden.py) and serve as the
reference-produced expected values for the `-m gpu` parity tests.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import refnerf_oracle as O  # noqa: E402
from refnerf_pl_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(v):
    return v.detach().cpu().numpy()


def _load_params(model, p):
    sd = model.nerf_mlp.state_dict()
    for k in sd:
        sd[k] = p[k].clone()
    model.nerf_mlp.load_state_dict(sd)


def model_case(name, gin_file, rays_np, n_rays, seed, bias_std, weight_scale, params_file=None):
    ns, config = ref_import.load(gin_file)
    torch.manual_seed(0)
    model = ns.models.construct_model(ns.utils.dummy_rays(), config)
    if params_file:   # trained weights (tools/train_parity.py, bf16x3 mode, 2 000 steps on the analytic sphere scene)
        with np.load(os.path.join(OUT, params_file)) as f:
            p = {k: torch.tensor(f[k]) for k in f.files}
    else:
        p = O.init_params(seed=seed, bias_std=bias_std, weight_scale=weight_scale)
    _load_params(model, p)
    rays_t = {k: torch.tensor(v) for k, v in rays_np.items()}
    rays = ns.utils.Rays(**rays_t)
    gt = torch.tensor(synthetic.gt_rgb(n_rays, seed))
    out = {'meta_seed': np.int64(seed), 'meta_bias_std': np.float64(bias_std),
           'meta_weight_scale': np.float64(weight_scale), 'gt_rgb': _np(gt),
           'param_checksum': np.float64(sum(float(v.double().abs().sum()) for v in p.values()))}
    if params_file:
        out['meta_params_file'] = np.array(params_file)
    for k, v in rays_np.items():
        out['rays_' + k] = v
    for mode in ('eval', 'train'):
        model.train(mode == 'train')
        with (torch.enable_grad() if mode == 'train' else torch.no_grad()):
            rend, hist = model(rays, 1.0, True)
        for lvl in range(len(rend)):
            for k, v in rend[lvl].items():
                out[f'{mode}_rend{lvl}_{k}'] = _np(v)
            for k, v in hist[lvl].items():
                if v is not None:
                    out[f'{mode}_hist{lvl}_{k}'] = _np(v)
        if mode == 'train':
            class B:
                pass
            batch = B()
            batch.rgb = _np(gt)
            loss = (ns.train_utils.compute_data_loss(batch, rend, rays, config)[0]
                    + ns.train_utils.orientation_loss(rays, model, hist, config)
                    + ns.train_utils.predicted_normal_loss(model, hist, config))
            model.zero_grad()
            loss.backward()
            out['train_loss'] = _np(loss)
            for k, v in model.nerf_mlp.named_parameters():
                g = _np(v.grad).reshape(-1)
                out['grad_norm_' + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
                out['grad_sub_' + k] = g[::97].copy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, len(out), 'arrays')


def op_cases():
    ns, _ = ref_import.load('blender_refnerf.gin')
    g = torch.Generator().manual_seed(7)
    out = {}
    # --- resample: random weights with zero runs and duplicate fenceposts
    n, s = 96, 128
    t = torch.sort(torch.rand(n, s + 1, generator=g), dim=-1).values
    t[:, 0] = 0.0
    t[:, -1] = 1.0
    t[::3, 40:48] = t[::3, 40:41]          # duplicate fenceposts -> zero-width bins
    w = torch.rand(n, s, generator=g) ** 4
    w[1::4, 10:60] = 0.0                   # zero-weight runs
    w[2::5, :] = 0.0
    w[2::5, 77] = 1.0                      # delta
    w = w / w.sum(-1, keepdim=True).clamp(min=1e-9)
    logits = torch.where(t[..., 1:] > t[..., :-1], torch.log(w + 0.01), -torch.tensor(float('inf')))
    sd = ns.stepfun.sample_intervals(t, logits, 128, single_jitter=False, domain=(0., 1.))
    wsm = torch.softmax(logits, -1)
    cw = ns.stepfun.integrate_weights(wsm)
    u = O.sample_grid(128).expand(n, 128)
    centers = ns.math.sorted_interp(u, cw, t)
    out.update(rs_t=_np(t), rs_w=_np(w), rs_sdist=_np(sd), rs_cw=_np(cw), rs_centers=_np(centers),
               rs_idx=_np(O.interval_index(u, cw)).astype(np.int32))
    # level-0 constant vector
    t0 = torch.tensor([[0.0, 1.0]])
    sd0 = ns.stepfun.sample_intervals(t0, torch.log(torch.ones(1, 1) + 0.01), 128, domain=(0., 1.))
    out['rs_level0'] = _np(sd0)
    # --- IPE
    r = synthetic.blender_rays(8, seed=5)
    rt = {k: torch.tensor(v) for k, v in r.items()}
    tdist = torch.sort(torch.rand(8, 33, generator=g) * 4 + 2, dim=-1).values
    means, cov = ns.render.cast_rays(tdist, rt['origins'], rt['directions'], rt['radii'], 'cone', diag=False)
    basis = torch.tensor(O.octahedron_basis())
    lm, lv = ns.coord.lift_and_diagonalize(means, cov, basis)
    out.update(ipe_tdist=_np(tdist), ipe_means=_np(means), ipe_cov=_np(cov),
               ipe_enc=_np(ns.coord.integrated_pos_enc(lm, lv, 0, 16)))
    for k in ('origins', 'directions', 'radii'):
        out['ipe_' + k] = r[k]
    # --- IDE at several roughness values (l=16 band is ill-conditioned in fp32)
    ide = ns.ref_utils.generate_ide_fn(5)
    d = torch.randn(256, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    out['ide_dirs'] = _np(d)
    for i, k in enumerate((0.0, 0.01, 0.3)):
        out[f'ide_enc_{i}'] = _np(ide(d, torch.full((256, 1), k)))
    out['ide_kappa_inv'] = np.array([0.0, 0.01, 0.3], np.float32)
    # --- compositing
    dens = torch.rand(16, 32, generator=g) * 3
    td = torch.sort(torch.rand(16, 33, generator=g) * 4 + 2, dim=-1).values
    dirs = torch.randn(16, 3, generator=g)
    wts = ns.render.compute_alpha_weights(dens, td, dirs)[0]
    out.update(cw_density=_np(dens), cw_tdist=_np(td), cw_dirs=_np(dirs), cw_weights=_np(wts))
    # --- interlevel / distortion / dilation
    tq = torch.sort(torch.rand(24, 65, generator=g), dim=-1).values
    wq = torch.softmax(torch.randn(24, 64, generator=g), -1) * 0.9
    te = torch.sort(torch.rand(24, 49, generator=g), dim=-1).values
    we = torch.softmax(torch.randn(24, 48, generator=g), -1)
    out.update(lo_t=_np(tq), lo_w=_np(wq), lo_tenv=_np(te), lo_wenv=_np(we),
               lo_loss=_np(ns.stepfun.lossfun_outer(tq, wq, te, we)),
               dist_loss=_np(ns.stepfun.lossfun_distortion(tq, wq)))
    tdil, wdil = ns.stepfun.max_dilate_weights(tq, wq, 0.01, domain=(torch.tensor(0.), torch.tensor(1.)),
                                               renormalize=True)
    out.update(dil_t=_np(tdil), dil_w=_np(wdil))
    np.savez_compressed(os.path.join(OUT, 'ops.npz'), **out)
    print('wrote ops', len(out), 'arrays')


def stepfun_util_cases():
    """stepfun.sample / stepfun.resample as the reference's off-path callers use them (camera_utils.py:266: 1-D float64
    path resampling; vis.py:144-145: 1-D intervals with [3, m] / [m] values) plus batched float32 inputs."""
    ns, _ = ref_import.load('blender_refnerf.gin')
    g = torch.Generator().manual_seed(3)
    out = {}
    theta = torch.sort(torch.rand(41, generator=g, dtype=torch.float64) * 6.28).values
    lengths = torch.rand(40, generator=g, dtype=torch.float64) + 0.1
    out.update(cam_theta=_np(theta), cam_lengths=_np(lengths),
               cam_sample=_np(ns.stepfun.sample(theta, torch.log(lengths), 31)))
    t = torch.sort(torch.rand(7, 33, generator=g)).values
    wl = torch.randn(7, 32, generator=g)
    out.update(s_t=_np(t), s_logits=_np(wl), s_lin=_np(ns.stepfun.sample(t, wl, 16)),
               s_center=_np(ns.stepfun.sample(t, wl, 16, deterministic_center=True)))
    dist_vis = torch.linspace(0, 1, 65)
    d = torch.sort(torch.rand(33, generator=g)).values
    d[0], d[-1] = 0.0, 1.0
    r = torch.rand(3, 32, generator=g)
    w = torch.rand(32, generator=g)
    out.update(r_t=_np(dist_vis), r_tp=_np(d), r_v3=_np(r), r_v1=_np(w))
    for ua in (False, True):
        out[f'r_out3_{int(ua)}'] = _np(ns.stepfun.resample(dist_vis, d, r, use_avg=ua))
        out[f'r_out1_{int(ua)}'] = _np(ns.stepfun.resample(dist_vis, d, w, use_avg=ua))
    np.savez_compressed(os.path.join(OUT, 'stepfun_utils.npz'), **out)
    print('wrote stepfun_utils', len(out), 'arrays')


def geometry_losses_case(name='llff_geom_losses', n_rays=160, seed=21):
    """configs/llff_refnerf_geometry_losses.gin as RefNeRFSystem.training_step runs it (nerf_system.py:84-191): main
    forward, sample_noisy_rays on the first 128 rays x 4 angles, second forward, every loss term of the config --
    computed by the UNMODIFIED reference functions (internal/sample_utils.py, internal/train_utils.py:207-325)."""
    ns, config = ref_import.load('llff_refnerf_geometry_losses.gin')
    from internal import sample_utils as ref_sample_utils
    torch.manual_seed(0)
    model = ns.models.construct_model(ns.utils.dummy_rays(), config)
    p = O.init_params(seed=3, bias_std=0.05, weight_scale=1.3)
    _load_params(model, p)
    rays_np = synthetic.llff_rays(n_rays, seed=seed)
    rays = ns.utils.Rays(**{k: torch.tensor(v) for k, v in rays_np.items()})
    gt = torch.tensor(synthetic.gt_rgb(n_rays, seed))
    warm = 0.5
    model.train(True)
    rend, hist = model(rays, 1.0, True)
    # the reference draws the rotation angles with uniform_ on a fresh tensor: reproduce the draw, then replay the seed
    torch.manual_seed(1234)
    angles = torch.zeros(config.sample_noise_angles * 3).uniform_(
        0, config.sample_angle_range / 180 * np.pi * warm).reshape(-1, 3)
    torch.manual_seed(1234)
    noisy = ref_sample_utils.sample_noisy_rays(rays, rend[-1], config.sample_angle_range, config.sample_noise_size,
                                               config.sample_noise_angles, warm)
    rend_n, _ = model(noisy, 1.0, True)

    class B:
        pass
    batch = B()
    batch.rgb = _np(gt)
    tu = ns.train_utils
    losses = {'data': tu.compute_data_loss(batch, rend, rays, config)[0],
              'orientation': tu.orientation_loss(rays, model, hist, config),
              'predicted_normals': tu.predicted_normal_loss(model, hist, config)}
    (losses['diffuse_consistency'], losses['specular_consistency'],
     losses['normals_consistency']) = tu.noisy_consistency_loss(model, rend, rend_n, config, warm)
    losses['acc'] = tu.accumulated_weights_loss(rend, config)
    losses['distance_consistency'] = tu.noisy_distance_consistency_loss(model, rays, noisy, rend, rend_n, config, warm)
    losses['weights_entropy'] = tu.weights_entropy_loss(model, rend, hist, config, warm)
    loss = torch.sum(torch.stack(list(losses.values())))
    model.zero_grad()
    loss.backward()
    out = {'meta_seed': np.int64(3), 'meta_bias_std': np.float64(0.05), 'meta_weight_scale': np.float64(1.3),
           'gt_rgb': _np(gt), 'warmup_ratio': np.float64(warm), 'xyz_angles': _np(angles), 'train_loss': _np(loss),
           'param_checksum': np.float64(sum(float(v.double().abs().sum()) for v in p.values()))}
    for k, v in rays_np.items():
        out['rays_' + k] = v
    for k in ('origins', 'directions', 'viewdirs'):
        out['noisy_' + k] = _np(getattr(noisy, k))
    for k, v in losses.items():
        out['loss_' + k] = _np(v)
    for lvl in range(2):
        for k in ('rgb', 'diffuse', 'specular', 'distance', 'acc', 'normals', 'normals_pred'):
            out[f'rend{lvl}_{k}'] = _np(rend[lvl][k])
            out[f'noisy_rend{lvl}_{k}'] = _np(rend_n[lvl][k])
    for k, v in model.nerf_mlp.named_parameters():
        g = _np(v.grad).reshape(-1)
        out['grad_norm_' + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out['grad_sub_' + k] = g[::97].copy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, len(out), 'arrays; losses:', {k: float(v) for k, v in losses.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    if 'geometry' in sys.argv[1:]:
        geometry_losses_case()
        return
    if 'stepfun_utils' in sys.argv[1:]:
        stepfun_util_cases()
        return
    if 'trained' in sys.argv[1:]:
        # trained-scale case only (needs tests/golden/trained_sphere_params.npz, written on the GPU box by
        # tools/train_parity.py): rays of one Blender-shaped camera, most of them crossing the learnt sphere
        n = 32
        model_case('blender_trained', 'blender_refnerf.gin', synthetic.blender_rays(n, seed=14), n, seed=14,
                   bias_std=0.0, weight_scale=1.0, params_file='trained_sphere_params.npz')
        return
    op_cases()
    n = 24
    model_case('blender_init', 'blender_refnerf.gin', synthetic.blender_rays(n, seed=11), n, seed=0,
               bias_std=0.0, weight_scale=1.0)
    model_case('blender_pert', 'blender_refnerf.gin', synthetic.blender_rays(n, seed=12), n, seed=1,
               bias_std=0.1, weight_scale=1.5)
    model_case('llff_geom', 'llff_refnerf_geometry_losses.gin', synthetic.llff_rays(n, seed=13), n, seed=2,
               bias_std=0.05, weight_scale=1.0)


if __name__ == '__main__':
    main()

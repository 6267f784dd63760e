"""Losses on the hot path's outputs (`internal/train_utils.py:33-88,151-204`) and the optimiser
recipe (`train_utils.py:448-467`, `math.py:46-78`).  Thin [N,3]-sized torch reductions over the
kernel outputs, as in the reference."""
import functools

import numpy as np
import torch

from . import ops, stepfun


def compute_data_loss(batch_rgb, renderings, lossmult, config):
    """train_utils.py:33-88 -> (loss, stats).  `batch_rgb` is the ground truth [..., 3]."""
    gt = torch.as_tensor(batch_rgb, device=renderings[0]['rgb'].device)[..., :3]
    if config.supervised_by_linear_rgb:                      # train_utils.py:40-41
        from . import image
        gt = image.srgb_to_linear(gt)
    if config.data_loss_type not in ('mse', 'charb'):
        assert False
    if renderings[0]['rgb'].is_cuda and renderings[0]['rgb'].dtype == torch.float32:
        # loss epilogue: the three sums of a level in one launch (rn_data_loss_fwd / _bwd)
        gt2 = gt.reshape(-1, 3).float().contiguous()
        n = gt2.shape[0]
        if config.disable_multiscale_loss:
            lm1 = gt2.new_empty((0,))
        else:
            lm1 = torch.broadcast_to(torch.as_tensor(lossmult, device=gt2.device, dtype=torch.float32), gt.shape[:-1] + (1,))
            lm1 = lm1.reshape(n).contiguous()
        losses, mses = [], []
        for rendering in renderings:
            sums = ops.data_loss_sums(rendering['rgb'].reshape(-1, 3).contiguous(), gt2, lm1,
                                      config.data_loss_type == 'charb', float(config.charb_padding))
            mses.append(sums[0] / sums[2])
            losses.append(sums[1] / sums[2])
        losses = torch.stack(losses)
        loss = config.data_coarse_loss_mult * torch.sum(losses[:-1]) + config.data_loss_mult * losses[-1]
        return loss, {'mses': torch.stack(mses).detach()}
    lm = torch.broadcast_to(lossmult, gt.shape)
    if config.disable_multiscale_loss:
        lm = torch.ones_like(lm)
    denom = lm.sum()
    losses, mses = [], []
    for rendering in renderings:
        resid_sq = (rendering['rgb'] - gt) ** 2
        mses.append((lm * resid_sq).sum() / denom)
        if config.data_loss_type == 'mse':
            data_loss = resid_sq
        elif config.data_loss_type == 'charb':
            data_loss = torch.sqrt(resid_sq + config.charb_padding ** 2)
        else:
            assert False
        losses.append((lm * data_loss).sum() / denom)
    losses = torch.stack(losses)
    loss = config.data_coarse_loss_mult * torch.sum(losses[:-1]) + config.data_loss_mult * losses[-1]
    return loss, {'mses': torch.stack(mses).detach()}


def interlevel_loss(ray_history, config):
    """train_utils.py:151-162."""
    c = ray_history[-1]['sdist'].detach()
    w = ray_history[-1]['weights'].detach()
    total = 0.
    for res in ray_history[:-1]:
        total = total + torch.mean(stepfun.lossfun_outer(c, w, res['sdist'], res['weights']))
    return config.interlevel_loss_mult * total


def orientation_loss(viewdirs, num_levels, ray_history, config):
    """train_utils.py:165-183."""
    total = 0.
    for i, res in enumerate(ray_history):
        n = res[config.orientation_loss_target]
        if n is None:
            raise ValueError('Normals cannot be None if orientation loss is on.')
        n_dot_v = (n * (-viewdirs)[..., None, :]).sum(dim=-1)
        loss = torch.mean((res['weights'] * torch.clamp(n_dot_v, max=0.0) ** 2).sum(dim=-1))
        total = total + (config.orientation_coarse_loss_mult if i < num_levels - 1 else config.orientation_loss_mult) * loss
    return total


def predicted_normal_loss(num_levels, ray_history, config):
    """train_utils.py:186-204."""
    total = 0.
    for i, res in enumerate(ray_history):
        n, n_pred = res['normals'], res['normals_pred']
        if n is None or n_pred is None:
            raise ValueError('Predicted normals and gradient normals cannot be None if predicted normal loss is on.')
        loss = torch.mean((res['weights'] * (1.0 - torch.sum(n * n_pred, dim=-1))).sum(dim=-1))
        total = total + (config.predicted_normal_coarse_loss_mult if i < num_levels - 1
                         else config.predicted_normal_loss_mult) * loss
    return total


def normal_losses(viewdirs, num_levels, ray_history, config):
    """orientation_loss + predicted_normal_loss (train_utils.py:165-204) from ONE pass over the per-sample normals
    per level (`rn_normal_losses_*`) instead of ~20 elementwise / reduction launches over
    [N,S,3] tensors.  Same value as `orientation_loss(...) + predicted_normal_loss(...)`."""
    want_ori = config.orientation_coarse_loss_mult > 0 or config.orientation_loss_mult > 0
    want_pred = config.predicted_normal_coarse_loss_mult > 0 or config.predicted_normal_loss_mult > 0
    target = config.orientation_loss_target
    if target not in ('normals', 'normals_pred'):
        raise ValueError(f'unsupported orientation_loss_target {target!r}')
    total = 0.
    for i, res in enumerate(ray_history):
        n, n_pred, w = res['normals'], res['normals_pred'], res['weights']
        if want_ori and res[target] is None:
            raise ValueError('Normals cannot be None if orientation loss is on.')
        if want_pred and (n is None or n_pred is None):
            raise ValueError('Predicted normals and gradient normals cannot be None if predicted normal loss is on.')
        s = w.shape[-1]
        flat3 = lambda t: t.reshape(-1, s, 3) if t is not None else w.new_empty((0,))
        per_ray = ops.normal_losses(w.reshape(-1, s).contiguous(), flat3(n).detach().contiguous(),
                                    flat3(n_pred).contiguous(), viewdirs.reshape(-1, 3).contiguous().float(),
                                    target == 'normals_pred')
        means = per_ray.mean(dim=1)
        fine = i == num_levels - 1
        if want_ori:
            total = total + (config.orientation_loss_mult if fine else config.orientation_coarse_loss_mult) * means[0]
        if want_pred:
            total = total + (config.predicted_normal_loss_mult if fine else config.predicted_normal_coarse_loss_mult) * means[1]
    return total


def total_loss(model, rays_viewdirs, lossmult, gt_rgb, renderings, ray_history, config, fused_normal_losses=True):
    """The loss sum of nerf_system.py:138-191 restricted to the terms of the Ref-NeRF configs."""
    loss, stats = compute_data_loss(gt_rgb, renderings, lossmult, config)
    if config.interlevel_loss_mult > 0:
        loss = loss + interlevel_loss(ray_history, config)
    want_ori = config.orientation_coarse_loss_mult > 0 or config.orientation_loss_mult > 0
    want_pred = config.predicted_normal_coarse_loss_mult > 0 or config.predicted_normal_loss_mult > 0
    if fused_normal_losses and (want_ori or want_pred) and ray_history[0]['weights'].is_cuda:
        return loss + normal_losses(rays_viewdirs, model.num_levels, ray_history, config), stats
    if want_ori:
        loss = loss + orientation_loss(rays_viewdirs, model.num_levels, ray_history, config)
    if want_pred:
        loss = loss + predicted_normal_loss(model.num_levels, ray_history, config)
    return loss, stats


# ----------------------------------------------------------------------------------------------
# geometry losses of configs/llff_refnerf_geometry_losses.gin (train_utils.py:90-119, 207-325): [N,3]-sized torch
# reductions over the renderings of the main rays and of the noisy rays (sample_utils.sample_noisy_rays)
# ----------------------------------------------------------------------------------------------
def _pair_var_loss(ref, noise, n_samples, n_angles, kind, mask):
    """one colour term of noisy_consistency_loss (train_utils.py:222-246); `ref` [N,3], `noise` [n_samples*n_angles,3]"""
    noise = noise.reshape(n_samples, n_angles, *noise.shape[1:])
    ref = ref[:n_samples, None]
    if kind == 'mse':
        v = ((ref - noise) ** 2).mean(dim=1, keepdim=True)
    elif kind == 'avg_mse':
        v = ((ref - noise.mean(dim=1, keepdim=True)) ** 2).mean(dim=1, keepdim=True)
    elif kind == 'var':
        v = torch.cat([ref, noise], dim=1).var(dim=1, keepdim=True).mean(dim=-1, keepdim=True)
    else:
        raise ValueError(f'unknown consistency loss type {kind!r}')
    return v.sum(dim=-1)[mask].mean()


def noisy_consistency_loss(model, renderings, renderings_noise, config, warmup_ratio=1.):
    """train_utils.py:207-273 -> (diffuse, specular, normals) consistency losses."""
    total = [0., 0., 0.]
    n_samples = config.sample_noise_size // config.patch_size ** 2
    n_angles = config.sample_noise_angles
    for i, (r, rn) in enumerate(zip(renderings, renderings_noise)):
        mask = r['acc'][:n_samples, None] > config.acc_threshold_for_consistency_loss
        diffuse = _pair_var_loss(r['diffuse'], rn['diffuse'], n_samples, n_angles, config.consistency_diffuse_loss_type, mask)
        specular = -_pair_var_loss(r['specular'], rn['specular'], n_samples, n_angles, config.consistency_specular_loss_type, mask)
        target = config.consistency_normal_loss_target
        if target not in ('normals', 'normals_pred'):
            raise ValueError('Given an unknown type of consistency_normal_loss_target.')
        if r.get(target) is None or rn.get(target) is None:
            raise ValueError('Predicted normals and gradient normals cannot be None if consistency loss is on.')
        n = r[target][:n_samples, None]
        n_noise = rn[target].reshape(n_samples, n_angles, *rn[target].shape[1:])
        normal = (1.0 - torch.sum(n * n_noise, dim=-1)).mean(dim=1, keepdim=True)[mask].mean()
        coarse = i < model.num_levels - 1
        total[0] = total[0] + warmup_ratio * (config.consistency_diffuse_coarse_loss_mult if coarse else config.consistency_diffuse_loss_mult) * diffuse
        total[1] = total[1] + warmup_ratio * (config.consistency_specular_coarse_loss_mult if coarse else config.consistency_specular_loss_mult) * specular
        total[2] = total[2] + warmup_ratio * (config.consistency_normal_coarse_loss_mult if coarse else config.consistency_normal_loss_mult) * normal
    return tuple(total)


def noisy_distance_consistency_loss(model, rays, noisy_rays, renderings, renderings_noise, config, warmup_ratio=1.):
    """train_utils.py:276-303: the main ray and its rotated copies should hit the same 3-D point."""
    total = 0.
    n_samples = config.sample_noise_size // config.patch_size ** 2
    n_angles = config.sample_noise_angles
    if config.consistency_distance_loss_type != 'mse':
        raise ValueError(f'unknown consistency_distance_loss_type {config.consistency_distance_loss_type!r}')
    for i, (r, rn) in enumerate(zip(renderings, renderings_noise)):
        o, d = rays.origins[:n_samples, None], rays.directions[:n_samples, None]
        dist = r['distance'][:n_samples, None]
        o_ = noisy_rays.origins.reshape(n_samples, n_angles, *noisy_rays.origins.shape[1:])
        d_ = noisy_rays.directions.reshape(n_samples, n_angles, *noisy_rays.directions.shape[1:])
        dist_ = rn['distance'].reshape(n_samples, n_angles, *rn['distance'].shape[1:])
        mask = r['acc'][:n_samples, None] > config.acc_threshold_for_consistency_loss
        mse = (((o + d * dist) - (o_ + d_ * dist_)) ** 2).mean(dim=1, keepdim=True)
        loss = mse.sum(dim=-1)[mask].mean()
        total = total + warmup_ratio * (config.consistency_distance_coarse_loss_mult if i < model.num_levels - 1
                                        else config.consistency_distance_loss_mult) * loss
    return total


def accumulated_weights_loss(renderings, config):
    """train_utils.py:306-309."""
    return config.accumulated_weights_loss_mult * ((1 - renderings[-1]['acc']) ** 2).mean()


def weights_entropy_loss(model, renderings, ray_history, config, warmup_ratio):
    """train_utils.py:311-322."""
    total = 0.
    for i, (r, res) in enumerate(zip(renderings, ray_history)):
        mask = r['acc'] > config.acc_threshold_for_weights_entropy_loss
        w = res['weights'][mask]
        loss = (-w * (w + 1e-10).log()).sum(dim=-1).mean()
        total = total + warmup_ratio * (config.weights_entropy_coarse_loss_mult if i < model.num_levels - 1
                                        else config.weights_entropy_loss_mult) * loss
    return total


def compute_depth_smoothness_loss(renderings, config):
    """train_utils.py:90-119 (patch batches, patch_size > 1)."""
    per_level = []
    bilateral = lambda x: torch.exp(-torch.abs(x).mean(-1, keepdim=True))
    for r in renderings:
        depths = r['distance']
        with torch.no_grad():
            acc00 = r['acc'][..., :-1, :-1, None]
            rgb = r['rgb']
        v00, v01, v10 = depths[..., :-1, :-1, :], depths[..., :-1, 1:, :], depths[..., 1:, :-1, :]
        w01 = bilateral(rgb[..., :-1, :-1, :] - rgb[..., :-1, 1:, :])
        w10 = bilateral(rgb[..., :-1, :-1, :] - rgb[..., 1:, :-1, :])
        l1 = torch.mean(torch.abs(acc00 * w01 * (v00 - v01) ** 2))
        l2 = torch.mean(torch.abs(acc00 * w10 * (v00 - v10) ** 2))
        per_level.append((l1 + l2) / 2)
    per_level = torch.stack(per_level)
    return config.depth_smoothness_coarse_loss_mult * torch.sum(per_level[:-1]) + config.depth_smoothness_loss_mult * per_level[-1]


def consistency_warmup_ratio(config, global_step):
    """nerf_system.py:97-113."""
    if config.consistency_warmup_steps > config.consistency_decay_steps:
        raise ValueError("Consistency loss decay should be after whole warmup.")
    ratio = 1.
    if 0. < config.consistency_warmup_steps <= 1.:
        ratio = min(1., global_step / (config.consistency_warmup_steps * config.max_steps))
    if 0. < config.consistency_decay_steps <= 1. and global_step >= config.consistency_decay_steps * config.max_steps:
        left = config.max_steps - global_step
        ratio = max(0., left / (config.max_steps - config.consistency_decay_steps * config.max_steps))
    return ratio


def _wants_consistency(config):
    return config.sample_noise_size > 0 and any(getattr(config, f'consistency_{k}_loss_mult') > 0 for k in (
        'diffuse_coarse', 'specular_coarse', 'normal_coarse', 'diffuse', 'specular', 'normal'))


def training_losses(model, rays, batch_rgb, config, train_frac=1.0, global_step=0, xyz_angles=None):
    """The model calls and the loss assembly of `RefNeRFSystem.training_step` (nerf_system.py:84-191): main forward
    (compute_extras as nerf_system.py:89-95), optional second forward on the noisy rays (:116-133), every loss term a
    config switches on.  -> (loss, losses dict, stats, renderings, ray_history)"""
    from . import sample_utils
    extras = config.compute_disp_metrics or config.compute_normal_metrics or config.sample_noise_size > 0
    renderings, ray_history = model(rays, train_frac=train_frac, compute_extras=extras)
    ratio = consistency_warmup_ratio(config, global_step)
    want_cons = _wants_consistency(config)
    want_dist = config.consistency_distance_loss_mult > 0 or config.consistency_distance_coarse_loss_mult > 0
    noisy = rend_n = None
    if want_cons:
        if config.patch_size ** 2 > config.sample_noise_size:
            raise ValueError(f'Patch size {config.patch_size}^2 too large for sampling noise view points {config.sample_noise_size}')
        noisy = sample_utils.sample_noisy_rays(rays, renderings[-1], config.sample_angle_range,
                                               config.sample_noise_size // config.patch_size ** 2,
                                               config.sample_noise_angles, ratio, xyz_angles=xyz_angles)
        rend_n, _ = model(noisy, train_frac=train_frac, compute_extras=True)
    losses = {}
    losses['data'], stats = compute_data_loss(batch_rgb, renderings, rays.lossmult, config)
    if config.interlevel_loss_mult > 0:
        losses['interlevel'] = interlevel_loss(ray_history, config)
    want_ori = config.orientation_coarse_loss_mult > 0 or config.orientation_loss_mult > 0
    want_pred = config.predicted_normal_coarse_loss_mult > 0 or config.predicted_normal_loss_mult > 0
    if (want_ori or want_pred) and ray_history[0]['weights'].is_cuda:
        losses['normals'] = normal_losses(rays.viewdirs, model.num_levels, ray_history, config)   # orientation + predicted
    else:
        if want_ori:
            losses['orientation'] = orientation_loss(rays.viewdirs, model.num_levels, ray_history, config)
        if want_pred:
            losses['predicted_normals'] = predicted_normal_loss(model.num_levels, ray_history, config)
    if config.patch_size > 1 and (config.depth_smoothness_coarse_loss_mult > 0 or config.depth_smoothness_loss_mult > 0):
        losses['smoothness'] = compute_depth_smoothness_loss(renderings, config)
    if want_cons:
        (losses['diffuse_consistency'], losses['specular_consistency'],
         losses['normals_consistency']) = noisy_consistency_loss(model, renderings, rend_n, config, ratio)
    if config.accumulated_weights_loss_mult > 0:
        losses['acc'] = accumulated_weights_loss(renderings, config)
    if want_dist:
        if noisy is None:
            raise ValueError('the distance consistency loss needs the noisy rays (a consistency loss multiplier > 0)')
        losses['distance_consistency'] = noisy_distance_consistency_loss(model, rays, noisy, renderings, rend_n, config, ratio)
    if config.weights_entropy_loss_mult > 0 or config.weights_entropy_coarse_loss_mult > 0:
        losses['weights_entropy'] = weights_entropy_loss(model, renderings, ray_history, config, ratio)
    loss = torch.sum(torch.stack([torch.as_tensor(v, device=losses['data'].device, dtype=torch.float32) for v in losses.values()]))
    return loss, losses, stats, renderings, ray_history


def collect_param_stats(model):
    """The per-parameter statistics of `RefNeRFSystem.on_after_backward` (nerf_system.py:212-217: squared weight norm,
    gradient norm, gradient max-abs for every parameter) without its 2 x 92 `.cpu()` round trips per step: three
    multi-tensor reductions on the device and no synchronisation at all -- the dict holds 0-dim DEVICE tensors (slices of
    three small vectors), which is what `on_train_batch_end` stacks and logs every `print_every` steps
    (nerf_system.py:220-271); call `.cpu()` on the stacked result there, once per logging interval.
    Same keys as the reference ('.' -> '/'), parameters without a gradient are reported as 0."""
    named = [(k, p) for k, p in model.named_parameters()]
    params = [p.detach() for _, p in named]
    grads = [p.grad.detach() if p.grad is not None else torch.zeros_like(p) for _, p in named]
    with torch.no_grad():
        w_l2 = torch.stack(torch._foreach_norm(params)) ** 2
        g_norm = torch.stack(torch._foreach_norm(grads))
        g_max = torch.stack(torch._foreach_norm(grads, float('inf')))
    keys = [k.replace('.', '/') for k, _ in named]
    return {'weights_l2s': dict(zip(keys, w_l2.unbind())), 'grad_norms': dict(zip(keys, g_norm.unbind())),
            'grad_maxes': dict(zip(keys, g_max.unbind()))}


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1):
    """math.py:46-78."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
    else:
        delay_rate = 1.
    t = np.clip(step / max_steps, 0, 1)
    return delay_rate * np.exp(t * (np.log(lr_final) - np.log(lr_init)) + np.log(lr_init)) / lr_init


def create_optimizer(config, params):
    """train_utils.py:448-467."""
    opt = torch.optim.Adam(params=params, lr=config.lr_init, betas=(config.adam_beta1, config.adam_beta2),
                           eps=config.adam_eps)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, functools.partial(
        learning_rate_decay, lr_init=config.lr_init, lr_final=config.lr_final, max_steps=config.max_steps,
        lr_delay_steps=config.lr_delay_steps, lr_delay_mult=config.lr_delay_mult))
    return opt, sched


def create_render_fn(model):
    """train_utils.py:470-477."""
    def render_eval_fn(train_frac, rays):
        return model(rays, train_frac=train_frac, compute_extras=True)
    return render_eval_fn

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

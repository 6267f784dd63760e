"""Function-level surface of the reference's `internal/stepfun.py`, backed by the CUDA ops."""
import torch

from . import ops


def sample_intervals(t, w_logits=None, num_samples=128, single_jitter=False, domain=(0.0, 1.0),
                     use_gpu_resampling=False, weights=None, padding=0.0, anneal=1.0, return_aux=False):
    """stepfun.py:209-258.  The kernel fuses the caller's logits (models.py:200-203): pass `weights`
    (+ padding/anneal); `w_logits` (already log-space) is accepted when padding == 0 via exp()."""
    if num_samples <= 1:
        raise ValueError(f'num_samples must be > 1, is {num_samples}.')
    if weights is None:
        weights = torch.exp(w_logits)
    lead = t.shape[:-1]
    s_in = weights.shape[-1]
    tt = ops._f32c(t.reshape(-1, s_in + 1))
    ww = ops._f32c(weights.reshape(-1, s_in))
    z = torch.zeros((tt.shape[0], 1), device=tt.device)
    so, _, cw, idx = ops.resample(tt, ww, z, z + 1, num_samples, float(padding), float(anneal), float(domain[0]),
                                  float(domain[1]), bool(return_aux))
    so = so.reshape(lead + (num_samples + 1,))
    if return_aux:
        return so, dict(cw=cw.reshape(lead + (s_in + 1,)), idx=idx.reshape(lead + (num_samples,)))
    return so


def lossfun_outer(t, w, t_env, w_env, eps=None):
    """stepfun.py:83-89 (differentiable w.r.t. w_env)."""
    lead = w.shape[:-1]
    s, se = w.shape[-1], w_env.shape[-1]
    f = lambda x, c: ops._f32c(x.reshape(-1, c))
    return ops.lossfun_outer(f(t.detach(), s + 1), f(w.detach(), s), f(t_env.detach(), se + 1), f(w_env, se)).reshape(lead + (s,))


def lossfun_distortion(t, w):
    """stepfun.py:261-272 (differentiable w.r.t. w)."""
    lead = w.shape[:-1]
    s = w.shape[-1]
    return ops.distortion(ops._f32c(t.detach().reshape(-1, s + 1)), ops._f32c(w.reshape(-1, s))).reshape(lead)

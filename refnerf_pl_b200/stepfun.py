"""Function-level surface of the reference's `internal/stepfun.py`, backed by the CUDA ops.

`sample` and `resample` are the two helpers the reference calls OUTSIDE the per-ray path, on a handful of CPU values
(`camera_utils.py:266` constant-speed camera paths, `vis.py:144-145` ray visualisation).  They are restated here as
device-agnostic torch code so that those callers keep working against this package; they are not part of the hot
path and have no kernel."""
import torch

from . import ops


def _integrate_weights(w):
    """stepfun.py:134-154: [0, min(1, cumsum(w[:-1])), 1]."""
    inner = torch.cumsum(w[..., :-1], dim=-1).clamp(max=1.0)
    edge = torch.zeros_like(w[..., :1])
    return torch.cat([edge, inner, edge + 1], dim=-1)


def _sorted_interp(x, xp, fp):
    """math.py:88-111 for sorted xp / fp, as a search + gather: the interval of x is the last xp <= x (queries below
    xp[0] collapse onto the first knot), the offset inside it is clipped to [0, 1] with 0/0 -> 0."""
    xp, fp = torch.broadcast_tensors(xp, fp)
    xp = xp.expand(x.shape[:-1] + xp.shape[-1:]).contiguous()
    fp = fp.expand(x.shape[:-1] + fp.shape[-1:]).contiguous()
    last = xp.shape[-1] - 1
    lo = (torch.searchsorted(xp, x.contiguous(), right=True) - 1)
    i0, i1 = lo.clamp(0, last), (lo + 1).clamp(0, last)
    x0, x1 = torch.gather(xp, -1, i0), torch.gather(xp, -1, i1)
    f0, f1 = torch.gather(fp, -1, i0), torch.gather(fp, -1, i1)
    frac = torch.clip(torch.nan_to_num((x - x0) / (x1 - x0), 0), 0, 1)
    return f0 + frac * (f1 - f0)


def sample(t, w_logits, num_samples, single_jitter=False, deterministic_center=False, use_gpu_resampling=False):
    """stepfun.py:168-206: num_samples points of the piecewise-constant PDF (t, softmax(w_logits)) at a regular grid of
    CDF values (this fork draws no random numbers)."""
    eps = torch.finfo(torch.float32).eps
    if deterministic_center:
        pad = 1 / (2 * num_samples)
        u = torch.linspace(pad, 1. - pad - eps, num_samples, device=w_logits.device)
    else:
        u = torch.linspace(0, 1. - eps, num_samples, device=w_logits.device)
    u = torch.broadcast_to(u, t.shape[:-1] + (num_samples,))
    return _sorted_interp(u, _integrate_weights(torch.softmax(w_logits, dim=-1)), t)


def _interp_f64(x, xp, fp):
    """math.py:114-142: float64, the line through the enclosing pair of knots, extrapolated (not clamped) outside."""
    x, xp, fp = x.double(), xp.double(), fp.double()
    slope = (fp[..., 1:] - fp[..., :-1]) / (xp[..., 1:] - xp[..., :-1])
    icpt = fp[..., :-1] - slope * xp[..., :-1]
    seg = ((x[..., :, None] >= xp[..., None, :]).sum(-1) - 1).clamp(0, slope.shape[-1] - 1)
    lead = torch.broadcast_shapes(seg.shape[:-1], slope.shape[:-1])
    seg = seg.expand(lead + seg.shape[-1:])
    slope = slope.expand(lead + slope.shape[-1:])
    icpt = icpt.expand(lead + icpt.shape[-1:])
    return torch.gather(slope, -1, seg) * x + torch.gather(icpt, -1, seg)


def resample(t, tp, vp, use_avg=False, eps=torch.tensor(torch.finfo(torch.float32).eps)):
    """stepfun.py:310-348: the step function (tp, vp) re-binned onto the intervals t, by sum (default) or by the
    width-weighted average.  Leading dims broadcast (the reference handles 1-D t / tp with 1-D or 2-D vp)."""
    if use_avg:
        width = torch.diff(tp, dim=-1)
        num = resample(t, tp, vp * width, use_avg=False)
        den = resample(t, tp, width, use_avg=False)
        return num / torch.maximum(eps.to(den.device), den)
    acc = torch.cumsum(vp, dim=-1)
    acc0 = torch.cat([torch.zeros_like(acc[..., :1]), acc], dim=-1)
    return torch.diff(_interp_f64(t, tp, acc0), dim=-1)


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

"""CPU oracle: a restatement of the Ref-NeRF per-ray hot path of minfenli/refnerf-pl.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file.  The product
(`refnerf_pl_b200/`) never routes through it and has no CPU fallback.

Parity status: PINNED.  Every function below is checked against the unmodified reference imported
from /root/reference (tests/test_oracle_vs_reference.py, run wherever the reference tree exists) and
against committed fixtures generated from the reference by `oracle/make_golden.py`
(tests/golden/*.npz, checked by tests/test_oracle_golden.py, which travels to the GPU box).
Exception: `contract()` -- the reference version raises TypeError (coord.py:20-26) so that
one function is "parity unpinned" and follows the formula in the reference source.

All functions are dtype-generic torch (fp32 = the reference's arithmetic, fp64 = noise-floor
calibration) and written functionally: parameters come in as a dict keyed by the reference's
state_dict names (`spatial_net.0.weight` ...).  Citations are path:line under /root/reference.
"""
import math

import numpy as np
import torch

EPS32 = float(np.finfo(np.float32).eps)


# ----------------------------------------------------------------------------------------------
# stepfun / math  (internal/stepfun.py, internal/math.py)
# ----------------------------------------------------------------------------------------------
def s_to_t(s, near, far):
    """coord.py:78-99 with fn=None: t = s*far + (1-s)*near."""
    return s * far + (1 - s) * near


def resample_logits(sdist, weights, padding, anneal=1.0):
    """models.py:200-203: log(w+padding)*anneal where the interval is non-empty, else -inf."""
    lg = anneal * torch.log(weights + padding)
    return torch.where(sdist[..., 1:] > sdist[..., :-1], lg, torch.full_like(lg, -float('inf')))


def integrate_weights(w):
    """stepfun.py:134-154: cw = [0, min(1, cumsum(w[:-1])), 1]."""
    cw = torch.clamp(torch.cumsum(w[..., :-1], dim=-1), max=1.0)
    z = torch.zeros_like(w[..., :1])
    return torch.cat([z, cw, z + 1], dim=-1)


def interval_index(u, cw):
    """Index form of math.py:88-111: idx = #{j : cw[j] <= u} - 1, per query."""
    return torch.searchsorted(cw.contiguous(), u.contiguous(), right=True) - 1


def sorted_interp(u, cw, t, return_index=False):
    """math.py:88-111 restated as a gather: bit-identical to the dense mask/max/min form because
    cw and t are sorted (max over a True-prefix = last True element)."""
    idx = interval_index(u, cw)
    n = cw.shape[-1]
    i0 = idx.clamp(0, n - 1)
    i1 = (idx + 1).clamp(0, n - 1)
    # queries below cw[0] (idx=-1) take xp[0] for both ends in the reference (max of all-False row)
    x0 = torch.gather(cw, -1, i0)
    x1 = torch.gather(cw, -1, i1)
    f0 = torch.gather(t, -1, i0)
    f1 = torch.gather(t, -1, i1)
    off = torch.clip(torch.nan_to_num((u - x0) / (x1 - x0), 0), 0, 1)
    out = f0 + off * (f1 - f0)
    return (out, idx) if return_index else out


def sample_grid(num_samples, dtype=torch.float32, device='cpu'):
    """stepfun.py:195-204 (deterministic_center=True): linspace(pad, 1-pad-eps, n).  The grid is
    built in fp32 like the reference (torch.linspace default dtype) then cast."""
    pad = 1 / (2 * num_samples)
    return torch.linspace(pad, 1. - pad - EPS32, num_samples, device=device).to(dtype)


def sample_intervals(t, w_logits, num_samples, domain=(0.0, 1.0), return_aux=False):
    """stepfun.py:209-258 -> sample() :168-206 -> invert_cdf() :157-165."""
    if num_samples <= 1:
        raise ValueError(f'num_samples must be > 1, is {num_samples}.')
    w = torch.softmax(w_logits, dim=-1)
    cw = integrate_weights(w)
    u = sample_grid(num_samples, t.dtype, t.device).expand(t.shape[:-1] + (num_samples,))
    centers, idx = sorted_interp(u, cw, t, return_index=True)
    mid = (centers[..., 1:] + centers[..., :-1]) / 2
    lo, hi = domain
    first = torch.clamp(2 * centers[..., :1] - mid[..., :1], min=lo)
    last = torch.clamp(2 * centers[..., -1:] - mid[..., -1:], max=hi)
    out = torch.cat([first, mid, last], dim=-1)
    if return_aux:
        return out, dict(cw=cw, idx=idx, centers=centers, u=u)
    return out


def weight_to_pdf(t, w):
    """stepfun.py:92-94."""
    return w / torch.clamp(t[..., 1:] - t[..., :-1], min=EPS32 ** 2)


def max_dilate_weights(t, w, dilation, domain=(-float('inf'), float('inf')), renormalize=False):
    """stepfun.py:102-131: dilate the step function by +-dilation with a running max of the pdf."""
    p = weight_to_pdf(t, w)
    t0 = t[..., :-1] - dilation
    t1 = t[..., 1:] + dilation
    td = torch.sort(torch.cat([t, t0, t1], dim=-1), dim=-1).values
    td = torch.clip(td, domain[0], domain[1])
    inside = (t0[..., None, :] <= td[..., None]) & (t1[..., None, :] > td[..., None])
    pd = torch.where(inside, p[..., None, :], torch.zeros_like(p[..., None, :])).amax(dim=-1)[..., :-1]
    wd = pd * (td[..., 1:] - td[..., :-1])
    if renormalize:
        wd = wd / torch.clamp(wd.sum(dim=-1, keepdim=True), min=EPS32 ** 2)
    return td, wd


def searchsorted_lo_hi(a, v):
    """stepfun.py:31-56: (idx_lo, idx_hi) with a[idx_lo] <= v < a[idx_hi], clamped at the ends."""
    n = a.shape[-1]
    cnt = torch.searchsorted(a.contiguous(), v.contiguous(), right=True)  # #{a <= v}
    idx_lo = (cnt - 1).clamp(min=0)
    idx_hi = cnt.clamp(max=n - 1)
    return idx_lo, idx_hi


def lossfun_outer(t, w, t_env, w_env):
    """stepfun.py:67-89: proposal weights should upper-bound the nerf weights."""
    cy = torch.cat([torch.zeros_like(w_env[..., :1]), torch.cumsum(w_env, dim=-1)], dim=-1)
    lo, hi = searchsorted_lo_hi(t_env, t)
    w_outer = torch.gather(cy, -1, hi)[..., 1:] - torch.gather(cy, -1, lo)[..., :-1]
    return torch.clamp(w - w_outer, min=0) ** 2 / (w + EPS32)


def lossfun_distortion(t, w):
    """stepfun.py:261-272."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = (ut[..., :, None] - ut[..., None, :]).abs()
    inter = (w * (w[..., None, :] * dut).sum(dim=-1)).sum(dim=-1)
    intra = (w ** 2 * (t[..., 1:] - t[..., :-1])).sum(dim=-1) / 3
    return inter + intra


def interp_f64(x, xp, fp):
    """math.py:114-142: fp64, unclamped linear extrapolation, per 1-D row (batched here)."""
    x, xp, fp = x.double(), xp.double(), fp.double()
    m = (fp[..., 1:] - fp[..., :-1]) / (xp[..., 1:] - xp[..., :-1])
    b = fp[..., :-1] - m * xp[..., :-1]
    idx = (x[..., :, None] >= xp[..., None, :]).sum(-1) - 1
    idx = idx.clamp(0, m.shape[-1] - 1)
    return torch.gather(m, -1, idx) * x + torch.gather(b, -1, idx)


def weighted_percentile(t, w, ps):
    """stepfun.py:294-307."""
    cw = integrate_weights(w)
    q = torch.tensor(ps, dtype=torch.float32, device=t.device) / 100
    q = q.expand(t.shape[:-1] + (len(ps),))
    return interp_f64(q, cw, t)


# ----------------------------------------------------------------------------------------------
# render.cast_rays + coord.integrated_pos_enc  (internal/render.py, internal/coord.py)
# ----------------------------------------------------------------------------------------------
def frustum_moments(t0, t1, radii):
    """render.py:64-80 (stable=True): t_mean, t_var, r_var of a conical frustum."""
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    denom = torch.clamp(3 * mu ** 2 + hw ** 2, min=EPS32)
    t_mean = mu + (2 * mu * hw ** 2) / denom
    t_var = (hw ** 2) / 3 - (4 / 15) * hw ** 4 * (12 * mu ** 2 - hw ** 2) / denom ** 2
    r_var = (mu ** 2) / 4 + (5 / 12) * hw ** 2 - (4 / 15) * (hw ** 4) / denom
    return t_mean, t_var, r_var * radii ** 2


def cast_rays(tdist, origins, directions, radii):
    """render.py:105-129 with ray_shape='cone', diag=False -> (means [..,S,3], cov [..,S,3,3])."""
    t_mean, t_var, r_var = frustum_moments(tdist[..., :-1], tdist[..., 1:], radii)
    d = directions
    means = origins[..., None, :] + d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.clamp((d ** 2).sum(-1, keepdim=True), min=1e-10)
    outer = d[..., :, None] * d[..., None, :]
    eye = torch.eye(3, dtype=d.dtype, device=d.device)
    null = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
    cov = t_var[..., None, None] * outer[..., None, :, :] + r_var[..., None, None] * null[..., None, :, :]
    return means, cov


def octahedron_basis():
    """geopoly.py:80-123 for ('octahedron', 1): the 3x3 anti-diagonal -1 matrix."""
    return np.array([[0, 0, -1], [0, -1, 0], [-1, 0, 0]], np.float32)


def lift_and_diagonalize(means, cov, basis):
    """coord.py:129-133."""
    fn_mean = means @ basis
    fn_cov_diag = (basis * (cov @ basis)).sum(dim=-2)
    return fn_mean, fn_cov_diag


def safe_sin(x):
    """math.py:22-34: sin(x) if |x| < 100pi else sin(x mod 100pi) (python-sign remainder)."""
    t = 100 * math.pi
    return torch.sin(torch.where(x.abs() < t, x, torch.remainder(x, t)))


def integrated_pos_enc(mean, var, min_deg, max_deg):
    """coord.py:107-126: [sin block | sin(+pi/2) block], feature index = k*3 + c."""
    scales = (2.0 ** torch.arange(min_deg, max_deg, device=mean.device)).to(mean.dtype)
    shape = mean.shape[:-1] + (-1,)
    sm = (mean[..., None, :] * scales[:, None]).reshape(shape)
    sv = (var[..., None, :] * scales[:, None] ** 2).reshape(shape)
    arg = torch.cat([sm, sm + 0.5 * math.pi], dim=-1)
    return torch.exp(-0.5 * torch.cat([sv, sv], dim=-1)) * safe_sin(arg)


# ----------------------------------------------------------------------------------------------
# ref_utils  (internal/ref_utils.py)
# ----------------------------------------------------------------------------------------------
def l2_normalize(x):
    """ref_utils.py:40-42."""
    return x / torch.sqrt(torch.clamp((x ** 2).sum(-1, keepdim=True), min=EPS32))


def reflect(viewdirs, normals):
    """ref_utils.py:22-37."""
    return 2.0 * (normals * viewdirs).sum(-1, keepdim=True) * normals - viewdirs


def _gen_binom(a, k):
    return np.prod(a - np.arange(k)) / math.factorial(k)


def _sph_harm_coeff(l, m, k):
    """ref_utils.py:53-95 (fp64 numpy)."""
    legendre = ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
                * _gen_binom(0.5 * (l + k + m - 1.0), l))
    return np.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * np.pi * math.factorial(l + m))) * legendre


def ide_tables(deg_view):
    """ref_utils.py:98-126: (m list, l list, mat[l_max+1, n_pairs] stored fp32)."""
    ms, ls = [], []
    for i in range(deg_view):
        l = 2 ** i
        for m in range(l + 1):
            ms.append(m)
            ls.append(l)
    l_max = 2 ** (deg_view - 1)
    mat = np.zeros((l_max + 1, len(ms)), np.float64)
    for i, (m, l) in enumerate(zip(ms, ls)):
        for k in range(l - m + 1):
            mat[k, i] = _sph_harm_coeff(l, m, k)
    return np.array(ms), np.array(ls), mat.astype(np.float32)


def integrated_dir_enc(xyz, kappa_inv, deg_view=5):
    """ref_utils.py:128-159: Vandermonde in z times coefficient matrix, times (x+iy)^m, attenuated
    by exp(-l(l+1)/2 * kappa_inv); output [Re | Im]."""
    ms, ls, mat = ide_tables(deg_view)
    cdtype = torch.complex64 if xyz.dtype == torch.float32 else torch.complex128
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    vmz = torch.cat([z ** i for i in range(mat.shape[0])], dim=-1)
    xy = torch.complex(x, y)
    vmxy = torch.cat([xy ** int(m) for m in ms], dim=-1)
    matt = torch.tensor(mat, dtype=torch.float32, device=xyz.device).to(xyz.dtype)
    sph = vmxy * (vmz @ matt).to(cdtype)
    sigma = torch.tensor(0.5 * ls * (ls + 1), device=xyz.device).to(xyz.dtype)
    ide = sph * torch.exp(-sigma * kappa_inv)
    return torch.cat([ide.real, ide.imag], dim=-1).to(xyz.dtype)


# ----------------------------------------------------------------------------------------------
# image.linear_to_srgb  (internal/image.py:51-59)
# ----------------------------------------------------------------------------------------------
def linear_to_srgb(x):
    s0 = 323 / 25 * x
    s1 = (211 * torch.clamp(x, min=EPS32) ** (5 / 12) - 11) / 200
    return torch.where(x <= 0.0031308, s0, s1)


# ----------------------------------------------------------------------------------------------
# MLP  (internal/models.py:533-750), Ref-NeRF configuration
# ----------------------------------------------------------------------------------------------
DEFAULT_MLP_CFG = dict(net_depth=8, net_depth_viewdirs=8, skip_layer=4, min_deg_point=0, max_deg_point=16,
                       deg_view=5, density_bias=0.5, roughness_bias=-1.0, rgb_premultiplier=1.0, rgb_bias=0.0,
                       rgb_padding=0.001, srgb_mapping=True, srgb_mapping_normalization=True)


def _linear(x, p, name):
    return torch.nn.functional.linear(x, p[name + '.weight'], p[name + '.bias'])


def mlp_forward(p, means, cov, viewdirs, training, cfg=None, basis=None):
    """models.py:533-750.  `p` maps reference parameter names to tensors.  Returns the ray_results
    dict (density, rgb, normals, normals_pred, grad_pred, tint, diffuse, specular, roughness)."""
    c = dict(DEFAULT_MLP_CFG)
    c.update(cfg or {})
    if basis is None:
        basis = torch.tensor(octahedron_basis(), dtype=means.dtype, device=means.device)
    if training:
        means = means.detach().requires_grad_(True)                       # models.py:562-563
    lm, lv = lift_and_diagonalize(means, cov, basis)                       # :566-567
    x = integrated_pos_enc(lm, lv, c['min_deg_point'], c['max_deg_point'])  # :570-571
    inputs = x
    for i in range(c['net_depth']):                                        # :576-580
        x = torch.relu(_linear(x, p, f'spatial_net.{i}'))
        if i % c['skip_layer'] == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    raw_density = _linear(x, p, 'raw_density')[..., 0]                     # :582
    normals = None
    if training:                                                           # :603-609 (result detached, D6)
        g = torch.autograd.grad(raw_density.sum(), means, retain_graph=True)[0]
        normals = -l2_normalize(g)
    grad_pred = _linear(x, p, 'grad_pred')                                 # :611-616
    normals_pred = -l2_normalize(grad_pred)
    density = torch.nn.functional.softplus(raw_density + c['density_bias'])  # :623
    raw_rgb_diffuse = _linear(x, p, 'raw_rgb_diffuse')                     # :634
    tint = torch.sigmoid(_linear(x, p, 'raw_tint'))                        # :637
    roughness = torch.nn.functional.softplus(_linear(x, p, 'raw_roughness') + c['roughness_bias'])  # :640
    bottleneck = _linear(x, p, 'bottleneck')                               # :645
    refdirs = reflect(-viewdirs[..., None, :], normals_pred)               # :662-663
    dir_enc = integrated_dir_enc(refdirs, roughness, c['deg_view'])        # :665
    dotprod = (normals_pred * viewdirs[..., None, :]).sum(-1, keepdim=True)  # :680-682
    x = torch.cat([bottleneck, dir_enc, dotprod], dim=-1)                  # :686
    inputs = x
    for i in range(c['net_depth_viewdirs']):                               # :690-694 (uses skip_layer, not _dir)
        x = torch.relu(_linear(x, p, f'viewdir_mlp.{i}'))
        if i % c['skip_layer'] == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    rgb = torch.sigmoid(c['rgb_premultiplier'] * _linear(x, p, 'rgb') + c['rgb_bias'])  # :699-700
    diffuse_linear = torch.sigmoid(raw_rgb_diffuse - math.log(3.0))        # :705-706
    specular_linear = tint * rgb                                           # :708
    if c['srgb_mapping']:                                                  # :712-723
        rgb = specular_linear + diffuse_linear
        if c['srgb_mapping_normalization']:
            rgb = rgb / torch.clamp(rgb.amax(dim=-1, keepdim=True), min=1.0)
        rgb = torch.clip(linear_to_srgb(rgb), 0.0, 1.0)
        diffuse = torch.clip(linear_to_srgb(diffuse_linear), 0.0, 1.0)
        specular = torch.clip(linear_to_srgb(specular_linear), 0.0, 1.0)
    else:                                                                  # :725-727
        rgb = specular_linear + diffuse_linear
        diffuse, specular = diffuse_linear, specular_linear
    rgb = rgb * (1 + 2 * c['rgb_padding']) - c['rgb_padding']              # :729
    return dict(density=density, rgb=rgb, normals=normals, normals_pred=normals_pred, grad_pred=grad_pred,
                tint=tint, diffuse=diffuse, specular=specular, roughness=roughness)


# ----------------------------------------------------------------------------------------------
# compositing  (internal/render.py:132-254)
# ----------------------------------------------------------------------------------------------
def compute_alpha_weights(density, tdist, dirs):
    """render.py:132-149 (opaque_background=False)."""
    delta = (tdist[..., 1:] - tdist[..., :-1]) * torch.linalg.norm(dirs[..., None, :], dim=-1)
    dd = density * delta
    alpha = 1 - torch.exp(-dd)
    trans = torch.exp(-torch.cat([torch.zeros_like(dd[..., :1]), torch.cumsum(dd[..., :-1], dim=-1)], dim=-1))
    return alpha * trans


def volumetric_rendering(rgbs, diffuse, specular, weights, tdist, bg, t_far, compute_extras, extras=None,
                         srgb_mapping='none'):
    """render.py:152-254."""
    out = {}
    acc = weights.sum(-1)
    bg_w = torch.clamp(1 - acc[..., None], min=0)
    comp = lambda v: (weights[..., None] * v).sum(-2)
    rgb = comp(rgbs) + bg_w * bg
    dif = comp(diffuse) + bg_w * bg
    spe = comp(specular) + bg_w * bg
    if srgb_mapping != 'none':
        if srgb_mapping.startswith('norm_'):
            rgb = rgb / torch.clamp(rgb.amax(-1, keepdim=True), min=1.0)
        if srgb_mapping.endswith('srgb'):
            rgb, dif, spe = (linear_to_srgb(v) for v in (rgb, dif, spe))
        elif not srgb_mapping.endswith('linear'):
            raise ValueError('Mapping types are none, linear, norm_linear, srgb, norm_srgb')
        rgb, dif, spe = (torch.clip(v, 0.0, 1.0) for v in (rgb, dif, spe))
    out['rgb'], out['diffuse'], out['specular'] = rgb, dif, spe
    t_mids = 0.5 * (tdist[..., :-1] + tdist[..., 1:])
    out['distance'] = comp(t_mids[..., None])
    out['acc'] = acc
    if compute_extras:
        for k, v in (extras or {}).items():
            if v is not None:
                out[k] = comp(v)
        expect = (weights * torch.log(t_mids)).sum(-1) / torch.clamp(acc, min=EPS32)
        dm = torch.nan_to_num(torch.exp(expect), float('inf'))
        out['distance_mean'] = torch.minimum(torch.maximum(dm, tdist[..., 0]), tdist[..., -1])
        t_aug = torch.cat([tdist, t_far], dim=-1)
        w_aug = torch.cat([weights, bg_w], dim=-1)
        pct = weighted_percentile(t_aug, w_aug, [5, 50, 95])
        out['distance_percentile_5'] = pct[..., 0]
        out['distance_median'] = pct[..., 1]
        out['distance_percentile_95'] = pct[..., 2]
    return out


# ----------------------------------------------------------------------------------------------
# Model.__call__  (internal/models.py:129-321), single_mlp Ref-NeRF configuration
# ----------------------------------------------------------------------------------------------
DEFAULT_MODEL_CFG = dict(num_levels=2, num_prop_samples=128, num_nerf_samples=128, anneal_slope=0.0,
                         resample_padding=0.01, dilation_bias=0.0, dilation_multiplier=0.0,
                         init_s_near=0.0, init_s_far=1.0, bg_intensity=1.0, srgb_mapping_render='none',
                         vis_num_rays=16)


def model_forward(p, rays, train_frac=1.0, compute_extras=False, training=False, model_cfg=None, mlp_cfg=None):
    """`rays`: dict with origins, directions, viewdirs [...,3], radii, near, far [...,1]."""
    mc = dict(DEFAULT_MODEL_CFG)
    mc.update(model_cfg or {})
    near, far = rays['near'], rays['far']
    sdist = torch.cat([torch.full_like(near, mc['init_s_near']), torch.full_like(far, mc['init_s_far'])], dim=-1)
    weights = torch.ones_like(near)
    prod = 1
    renderings, history = [], []
    for lvl in range(mc['num_levels']):
        is_prop = lvl < mc['num_levels'] - 1
        ns = mc['num_prop_samples'] if is_prop else mc['num_nerf_samples']
        dilation = mc['dilation_bias'] + mc['dilation_multiplier'] * (mc['init_s_far'] - mc['init_s_near']) / prod
        prod *= ns
        if lvl > 0 and (mc['dilation_bias'] > 0 or mc['dilation_multiplier'] > 0):     # models.py:177-187
            sdist, weights = max_dilate_weights(sdist, weights, dilation,
                                                domain=(mc['init_s_near'], mc['init_s_far']), renormalize=True)
            sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]
        if mc['anneal_slope'] > 0:                                                     # :190-195
            s = mc['anneal_slope']
            anneal = (s * train_frac) / ((s - 1) * train_frac + 1)
        else:
            anneal = 1.0
        logits = resample_logits(sdist, weights, mc['resample_padding'], anneal)        # :200-203
        sdist = sample_intervals(sdist, logits, ns, domain=(mc['init_s_near'], mc['init_s_far'])).detach()
        tdist = s_to_t(sdist, near, far)                                               # :218
        means, cov = cast_rays(tdist, rays['origins'], rays['directions'], rays['radii'])  # :221-227
        res = mlp_forward(p, means, cov, rays['viewdirs'], training, mlp_cfg)          # :236-241
        weights = compute_alpha_weights(res['density'], tdist, rays['directions'])     # :244-249
        extras = {k: v for k, v in res.items() if k.startswith('normals') or k in ('roughness', 'tint')}
        rend = volumetric_rendering(res['rgb'], res['diffuse'], res['specular'], weights, tdist, mc['bg_intensity'],
                                    far, compute_extras, extras, mc['srgb_mapping_render'])  # :270-288
        if compute_extras:                                                             # :290-301
            n = mc['vis_num_rays']
            rend['ray_sdist'] = sdist.reshape(-1, sdist.shape[-1])[:n]
            rend['ray_weights'] = weights.reshape(-1, weights.shape[-1])[:n]
            rend['ray_rgbs'] = res['rgb'].reshape((-1,) + res['rgb'].shape[-2:])[:n]
        renderings.append(rend)
        res['sdist'] = sdist.clone()
        res['weights'] = weights.clone()
        history.append(res)
    if compute_extras:                                                                 # :308-319
        final_rgb = (renderings[-1]['ray_rgbs'] * renderings[-1]['ray_weights'][..., None]).sum(-2)
        for r in renderings[:-1]:
            r['ray_rgbs'] = final_rgb[:, None, :].expand(r['ray_rgbs'].shape)
    return renderings, history


# ----------------------------------------------------------------------------------------------
# losses  (internal/train_utils.py:33-88, 151-204)
# ----------------------------------------------------------------------------------------------
DEFAULT_LOSS_CFG = dict(data_loss_mult=1.0, data_coarse_loss_mult=0.1, orientation_loss_mult=0.1,
                        orientation_coarse_loss_mult=0.01, predicted_normal_loss_mult=3e-4,
                        predicted_normal_coarse_loss_mult=3e-5, interlevel_loss_mult=0.0)


def data_loss(renderings, gt_rgb, lossmult, cfg):
    """train_utils.py:33-88 ('mse')."""
    lm = lossmult.expand(gt_rgb.shape)
    per = [(lm * (r['rgb'] - gt_rgb) ** 2).sum() / lm.sum() for r in renderings]
    return cfg['data_coarse_loss_mult'] * sum(per[:-1]) + cfg['data_loss_mult'] * per[-1]


def orientation_loss(history, viewdirs, cfg):
    """train_utils.py:165-183 (target normals_pred)."""
    total = 0.
    for i, res in enumerate(history):
        ndv = (res['normals_pred'] * (-viewdirs)[..., None, :]).sum(-1)
        loss = (res['weights'] * torch.clamp(ndv, max=0) ** 2).sum(-1).mean()
        total = total + (cfg['orientation_coarse_loss_mult'] if i < len(history) - 1 else cfg['orientation_loss_mult']) * loss
    return total


def predicted_normal_loss(history, cfg):
    """train_utils.py:186-204."""
    total = 0.
    for i, res in enumerate(history):
        loss = (res['weights'] * (1.0 - (res['normals'] * res['normals_pred']).sum(-1))).sum(-1).mean()
        total = total + (cfg['predicted_normal_coarse_loss_mult'] if i < len(history) - 1
                         else cfg['predicted_normal_loss_mult']) * loss
    return total


def interlevel_loss(history, cfg):
    """train_utils.py:151-162."""
    c = history[-1]['sdist'].detach()
    w = history[-1]['weights'].detach()
    total = 0.
    for res in history[:-1]:
        total = total + lossfun_outer(c, w, res['sdist'], res['weights']).mean()
    return cfg['interlevel_loss_mult'] * total


def total_loss(renderings, history, rays, gt_rgb, cfg=None):
    c = dict(DEFAULT_LOSS_CFG)
    c.update(cfg or {})
    loss = data_loss(renderings, gt_rgb, rays['lossmult'], c)
    if c['orientation_loss_mult'] > 0 or c['orientation_coarse_loss_mult'] > 0:
        loss = loss + orientation_loss(history, rays['viewdirs'], c)
    if c['predicted_normal_loss_mult'] > 0 or c['predicted_normal_coarse_loss_mult'] > 0:
        loss = loss + predicted_normal_loss(history, c)
    if c['interlevel_loss_mult'] > 0:
        loss = loss + interlevel_loss(history, c)
    return loss


def contract(x):
    """coord.py:20-26 formula.  PARITY UNPINNED: the reference function raises TypeError (D7)."""
    m = torch.clamp((x ** 2).sum(-1, keepdim=True), min=EPS32)
    return torch.where(m <= 1, x, ((2 * torch.sqrt(m) - 1) / m) * x)


# ----------------------------------------------------------------------------------------------
# parameter construction with the reference's init (models.py:38-47) and shapes
# ----------------------------------------------------------------------------------------------
def param_shapes(net_width=256, bottleneck=128, in_feat=96, view_in=201, depth=8, skip=4):
    shapes = {}
    k = in_feat
    for i in range(depth):
        shapes[f'spatial_net.{i}'] = (net_width, k)
        k = net_width + (in_feat if (i % skip == 0 and i > 0) else 0)
    for name, n in (('raw_density', 1), ('grad_pred', 3), ('raw_roughness', 1), ('raw_rgb_diffuse', 3),
                    ('raw_tint', 3), ('bottleneck', bottleneck)):
        shapes[name] = (n, net_width)
    k = view_in
    for i in range(depth):
        shapes[f'viewdir_mlp.{i}'] = (net_width, k)
        k = net_width + (view_in if (i % skip == 0 and i > 0) else 0)
    shapes['rgb'] = (3, net_width)
    return shapes


def init_params(seed=0, dtype=torch.float32, bias_std=0.0, weight_scale=1.0):
    """kaiming_uniform_(a=sqrt(5)) weights = U(-1/sqrt(K), 1/sqrt(K)); zero biases (models.py:38-47).
    `bias_std`/`weight_scale` perturb away from init to exercise trained-like regimes in tests."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, (n, k) in param_shapes().items():
        bound = 1.0 / math.sqrt(k)
        p[name + '.weight'] = ((torch.rand(n, k, generator=g) * 2 - 1) * bound * weight_scale).to(dtype)
        p[name + '.bias'] = (torch.randn(n, generator=g) * bias_std).to(dtype)
    return p


# ---------------------------------------------------------------------------------------------
# ray generation: camera_utils.convert_to_ndc (camera_utils.py:31-97) and
# camera_utils.pixels_to_rays (camera_utils.py:502-614), perspective cameras without distortion, numpy
# arithmetic exactly as the reference's default xnp=np path (integer pixels + 0.5 promote to float64)
# ---------------------------------------------------------------------------------------------
def convert_to_ndc_np(origins, directions, pixtocam, near=1.0):
    import numpy as np
    t = -(near + origins[..., 2]) / directions[..., 2]
    origins = origins + t[..., None] * directions
    dx, dy, dz = np.moveaxis(directions, -1, 0)
    ox, oy, oz = np.moveaxis(origins, -1, 0)
    xmult = 1. / pixtocam[0, 2]
    ymult = 1. / pixtocam[1, 2]
    origins_ndc = np.stack([xmult * ox / oz, ymult * oy / oz, -np.ones_like(oz)], axis=-1)
    infinity_ndc = np.stack([xmult * dx / dz, ymult * dy / dz, np.ones_like(oz)], axis=-1)
    return origins_ndc, infinity_ndc - origins_ndc


def pixels_to_rays_np(pix_x_int, pix_y_int, pixtocams, camtoworlds, pixtocam_ndc=None):
    """-> origins, directions, viewdirs [.., 3], radii [.., 1], imageplane [.., 2] (float64, as the reference returns)."""
    import numpy as np

    def pix_to_dir(x, y):
        return np.stack([x + .5, y + .5, np.ones_like(x)], axis=-1)

    stacked = np.stack([pix_to_dir(pix_x_int, pix_y_int), pix_to_dir(pix_x_int + 1, pix_y_int),
                        pix_to_dir(pix_x_int, pix_y_int + 1)], axis=0)
    mat_vec_mul = lambda A, b: np.matmul(A, b[..., None])[..., 0]
    cam_dirs = mat_vec_mul(pixtocams, stacked)
    cam_dirs = np.matmul(cam_dirs, np.diag(np.array([1., -1., -1.])))
    imageplane = cam_dirs[0, ..., :2]
    dirs_stacked = mat_vec_mul(camtoworlds[..., :3, :3], cam_dirs)
    directions, dx, dy = dirs_stacked
    origins = np.broadcast_to(camtoworlds[..., :3, -1], directions.shape)
    viewdirs = directions / np.linalg.norm(directions, axis=-1, keepdims=True)
    if pixtocam_ndc is None:
        dx_norm = np.linalg.norm(dx - directions, axis=-1)
        dy_norm = np.linalg.norm(dy - directions, axis=-1)
    else:
        origins_dx, _ = convert_to_ndc_np(origins, dx, pixtocam_ndc)
        origins_dy, _ = convert_to_ndc_np(origins, dy, pixtocam_ndc)
        origins, directions = convert_to_ndc_np(origins, directions, pixtocam_ndc)
        dx_norm = np.linalg.norm(origins_dx - origins, axis=-1)
        dy_norm = np.linalg.norm(origins_dy - origins, axis=-1)
    radii = (0.5 * (dx_norm + dy_norm))[..., None] * 2 / np.sqrt(12)
    return origins, directions, viewdirs, radii, imageplane

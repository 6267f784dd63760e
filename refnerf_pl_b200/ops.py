"""torch.library custom ops (namespace `refnerf_b200`) over the C ABI, with explicit backward.

Every op is registered for CUDA only: calling one with CPU tensors raises (there is no fallback).
The ops take and return plain contiguous fp32 CUDA tensors; raw device pointers and the current
CUDA stream are handed to librefnerf_b200.so through ctypes.
"""
import ctypes
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

NS = 'refnerf_b200'
_EPS32 = float(torch.finfo(torch.float32).eps)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


_U_CACHE = {}


def sample_grid(num_samples, device):
    """stepfun.py:195-204 (deterministic_center): built exactly like the reference (CPU fp32 linspace)."""
    key = (num_samples, str(device))
    if key not in _U_CACHE:
        pad = 1 / (2 * num_samples)
        _U_CACHE[key] = torch.linspace(pad, 1. - pad - _EPS32, num_samples).to(device)
    return _U_CACHE[key]


# ------------------------------------------------------------------------------------------
# resample
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::resample', mutates_args=(), device_types='cuda')
def resample(sdist: Tensor, weights: Tensor, near: Tensor, far: Tensor, num_samples: int, padding: float,
             anneal: float, dom_lo: float, dom_hi: float, want_aux: bool) -> List[Tensor]:
    """-> [sdist_out [N,S+1], tdist_out [N,S+1], cw [N,s_in+1] | empty, idx [N,S] int32 | empty]"""
    lib = _lib.load()
    sdist, weights, near, far = _f32c(sdist), _f32c(weights), _f32c(near), _f32c(far)
    n, s_in = weights.shape
    dev = sdist.device
    so = torch.empty((n, num_samples + 1), device=dev, dtype=torch.float32)
    to = torch.empty_like(so)
    cw = torch.empty((n, s_in + 1) if want_aux else (0,), device=dev, dtype=torch.float32)
    idx = torch.empty((n, num_samples) if want_aux else (0,), device=dev, dtype=torch.int32)
    u = sample_grid(num_samples, dev)
    _lib.check(lib.rn_resample(_ptr(sdist), _ptr(weights), None, _ptr(u), _ptr(near), _ptr(far), n, s_in, num_samples,
                               padding, anneal, dom_lo, dom_hi, _ptr(so), _ptr(to), _ptr(cw), _ptr(idx), _stream()))
    return [so, to, cw, idx]


def resample_from_cdf(sdist: Tensor, cw: Tensor, near: Tensor, far: Tensor, num_samples: int, dom_lo: float = 0.0,
                      dom_hi: float = 1.0):
    """Search + interpolation stage of the resampler on a GIVEN CDF `cw` [N, s_in+1] (math.sorted_interp, math.py:88-111,
    and the fencepost construction of stepfun.py:246-258) -> (sdist_out, tdist_out, idx).  Test surface: fed with the
    reference's own CDF the interval indices and fenceposts must match the reference bit for bit."""
    lib = _lib.load()
    sdist, cw, near, far = _f32c(sdist), _f32c(cw), _f32c(near), _f32c(far)
    n, s1 = cw.shape
    dev = sdist.device
    so = torch.empty((n, num_samples + 1), device=dev, dtype=torch.float32)
    to = torch.empty_like(so)
    idx = torch.empty((n, num_samples), device=dev, dtype=torch.int32)
    u = sample_grid(num_samples, dev)
    _lib.check(lib.rn_resample(_ptr(sdist), None, _ptr(cw), _ptr(u), _ptr(near), _ptr(far), n, s1 - 1, num_samples,
                               0.0, 1.0, dom_lo, dom_hi, _ptr(so), _ptr(to), None, _ptr(idx), _stream()))
    return so, to, idx


@resample.register_fake
def _(sdist, weights, near, far, num_samples, padding, anneal, dom_lo, dom_hi, want_aux):
    n, s_in = weights.shape
    so = sdist.new_empty((n, num_samples + 1))
    return [so, torch.empty_like(so), sdist.new_empty((n, s_in + 1) if want_aux else (0,)),
            sdist.new_empty((n, num_samples) if want_aux else (0,), dtype=torch.int32)]


@torch.library.custom_op(f'{NS}::max_dilate_weights', mutates_args=(), device_types='cuda')
def max_dilate_weights(t: Tensor, w: Tensor, dilation: float, dom_lo: float, dom_hi: float, renormalize: bool,
                       trim: bool) -> List[Tensor]:
    """stepfun.max_dilate_weights (stepfun.py:102-131); trim=True also drops the first / last fencepost and weight as
    models.py:186-187 does.  t [N,s+1], w [N,s] -> [t_dilated, w_dilated].  No gradient (its consumer is detached)."""
    lib = _lib.load()
    t, w = _f32c(t), _f32c(w)
    n, s = w.shape
    cut = 2 if trim else 0
    to = torch.empty((n, 3 * s + 1 - cut), device=t.device, dtype=torch.float32)
    wo = torch.empty((n, 3 * s - cut), device=t.device, dtype=torch.float32)
    _lib.check(lib.rn_max_dilate_weights(_ptr(t), _ptr(w), n, s, dilation, dom_lo, dom_hi, int(renormalize), int(trim),
                                         _ptr(to), _ptr(wo), _stream()))
    return [to, wo]


@max_dilate_weights.register_fake
def _(t, w, dilation, dom_lo, dom_hi, renormalize, trim):
    n, s = w.shape
    cut = 2 if trim else 0
    return [t.new_empty((n, 3 * s + 1 - cut)), t.new_empty((n, 3 * s - cut))]


# ------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::composite_fwd', mutates_args=(), device_types='cuda')
def composite_fwd(density: Tensor, tdist: Tensor, dirs: Tensor, far: Tensor, rgb: Tensor, diffuse: Tensor,
                  specular: Tensor, normals: Tensor, normals_pred: Tensor, roughness: Tensor, tint: Tensor, bg: float,
                  extras: bool) -> List[Tensor]:
    """-> [weights [N,S], comp [N,16], extras [N,12] | empty, percentiles [N,3] f64 | empty]"""
    lib = _lib.load()
    n, s = density.shape
    dev = density.device
    w = torch.empty((n, s), device=dev, dtype=torch.float32)
    comp = torch.empty((n, 16), device=dev, dtype=torch.float32)
    ex = torch.empty((n, 12) if extras else (0,), device=dev, dtype=torch.float32)
    pct = torch.empty((n, 3) if extras else (0,), device=dev, dtype=torch.float64)
    _lib.check(lib.rn_composite_fwd(_ptr(density), _ptr(tdist), _ptr(dirs), _ptr(far), _ptr(rgb), _ptr(diffuse),
                                    _ptr(specular), _ptr(normals) if extras else None,
                                    _ptr(normals_pred) if extras else None, _ptr(roughness) if extras else None,
                                    _ptr(tint) if extras else None, n, s, bg, _ptr(w), _ptr(comp), _ptr(ex), _ptr(pct),
                                    _stream()))
    return [w, comp, ex, pct]


@composite_fwd.register_fake
def _(density, tdist, dirs, far, rgb, diffuse, specular, normals, normals_pred, roughness, tint, bg, extras):
    n, s = density.shape
    return [density.new_empty((n, s)), density.new_empty((n, 16)), density.new_empty((n, 12) if extras else (0,)),
            density.new_empty((n, 3) if extras else (0,), dtype=torch.float64)]


@torch.library.custom_op(f'{NS}::composite_bwd', mutates_args=(), device_types='cuda')
def composite_bwd(density: Tensor, tdist: Tensor, dirs: Tensor, rgb: Tensor, diffuse: Tensor, specular: Tensor,
                  normals: Tensor, normals_pred: Tensor, roughness: Tensor, tint: Tensor, weights: Tensor, comp: Tensor,
                  g_weights: Tensor, g_comp: Tensor, g_extras: Tensor, bg: float, extras: bool) -> List[Tensor]:
    """-> [d_density, d_rgb, d_diffuse, d_specular, d_normals_pred, d_roughness, d_tint]"""
    lib = _lib.load()
    n, s = density.shape
    d_density = torch.empty_like(density)
    d_rgb, d_diffuse, d_specular = torch.empty_like(rgb), torch.empty_like(diffuse), torch.empty_like(specular)
    if extras:
        d_np, d_ro, d_ti = torch.empty_like(normals_pred), torch.empty_like(roughness), torch.empty_like(tint)
    else:
        d_np, d_ro, d_ti = (density.new_empty((0,)) for _ in range(3))
    has_ge = extras and g_extras.numel() > 0
    _lib.check(lib.rn_composite_bwd(
        _ptr(density), _ptr(tdist), _ptr(dirs), _ptr(rgb), _ptr(diffuse), _ptr(specular),
        _ptr(normals) if has_ge else None, _ptr(normals_pred) if has_ge else None, _ptr(roughness) if has_ge else None,
        _ptr(tint) if has_ge else None, _ptr(weights), _ptr(comp), _ptr(g_weights), _ptr(g_comp),
        _ptr(g_extras) if has_ge else None, n, s, bg, _ptr(d_density), _ptr(d_rgb), _ptr(d_diffuse), _ptr(d_specular),
        _ptr(d_np), _ptr(d_ro), _ptr(d_ti), _stream()))
    return [d_density, d_rgb, d_diffuse, d_specular, d_np, d_ro, d_ti]


@composite_bwd.register_fake
def _(density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, weights, comp, g_weights,
      g_comp, g_extras, bg, extras):
    e = lambda: density.new_empty((0,))
    return [torch.empty_like(density), torch.empty_like(rgb), torch.empty_like(diffuse), torch.empty_like(specular),
            torch.empty_like(normals_pred) if extras else e(), torch.empty_like(roughness) if extras else e(),
            torch.empty_like(tint) if extras else e()]


def _composite_setup(ctx, inputs, output):
    (density, tdist, dirs, far, rgb, diffuse, specular, normals, normals_pred, roughness, tint, bg, extras) = inputs
    w, comp, ex, pct = output
    ctx.save_for_backward(density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, w, comp)
    ctx.bg, ctx.extras = bg, extras
    ctx.mark_non_differentiable(pct)
    ctx.set_materialize_grads(False)


def _composite_backward(ctx, grads):
    g_w, g_comp, g_ex, _ = grads
    (density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, w, comp) = ctx.saved_tensors
    empty = density.new_empty((0,))
    g_w = _f32c(g_w) if g_w is not None else empty
    g_comp = _f32c(g_comp) if g_comp is not None else torch.zeros_like(comp)
    g_ex = _f32c(g_ex) if (g_ex is not None and ctx.extras) else empty
    d = composite_bwd(density, tdist, dirs, rgb, diffuse, specular, normals, normals_pred, roughness, tint, w, comp, g_w,
                      g_comp, g_ex, ctx.bg, ctx.extras)
    d_np, d_ro, d_ti = (d[4], d[5], d[6]) if ctx.extras else (None, None, None)
    # inputs: density, tdist, dirs, far, rgb, diffuse, specular, normals, normals_pred, roughness, tint, bg, extras
    return d[0], None, None, None, d[1], d[2], d[3], None, d_np, d_ro, d_ti, None, None


torch.library.register_autograd(f'{NS}::composite_fwd', _composite_backward, setup_context=_composite_setup)


# ------------------------------------------------------------------------------------------
# step-function losses
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::lossfun_outer', mutates_args=(), device_types='cuda')
def lossfun_outer(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor) -> Tensor:
    lib = _lib.load()
    n, s = w.shape
    out = torch.empty_like(w)
    _lib.check(lib.rn_lossfun_outer_fwd(_ptr(t), _ptr(w), _ptr(t_env), _ptr(w_env), n, s, w_env.shape[1], _ptr(out),
                                        _stream()))
    return out


@lossfun_outer.register_fake
def _(t, w, t_env, w_env):
    return torch.empty_like(w)


@torch.library.custom_op(f'{NS}::lossfun_outer_bwd', mutates_args=(), device_types='cuda')
def lossfun_outer_bwd(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor, g: Tensor) -> Tensor:
    lib = _lib.load()
    n, s = w.shape
    out = torch.empty_like(w_env)
    _lib.check(lib.rn_lossfun_outer_bwd(_ptr(t), _ptr(w), _ptr(t_env), _ptr(w_env), _ptr(g), n, s, w_env.shape[1],
                                        _ptr(out), _stream()))
    return out


@lossfun_outer_bwd.register_fake
def _(t, w, t_env, w_env, g):
    return torch.empty_like(w_env)


def _lo_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _lo_backward(ctx, g):
    t, w, t_env, w_env = ctx.saved_tensors
    return None, None, None, lossfun_outer_bwd(t, w, t_env, w_env, _f32c(g))


torch.library.register_autograd(f'{NS}::lossfun_outer', _lo_backward, setup_context=_lo_setup)


@torch.library.custom_op(f'{NS}::distortion', mutates_args=(), device_types='cuda')
def distortion(t: Tensor, w: Tensor) -> Tensor:
    lib = _lib.load()
    n, s = w.shape
    out = torch.empty((n,), device=w.device, dtype=torch.float32)
    _lib.check(lib.rn_distortion_fwd(_ptr(t), _ptr(w), n, s, _ptr(out), _stream()))
    return out


@distortion.register_fake
def _(t, w):
    return w.new_empty((w.shape[0],))


@torch.library.custom_op(f'{NS}::distortion_bwd', mutates_args=(), device_types='cuda')
def distortion_bwd(t: Tensor, w: Tensor, g: Tensor) -> Tensor:
    lib = _lib.load()
    n, s = w.shape
    out = torch.empty_like(w)
    _lib.check(lib.rn_distortion_bwd(_ptr(t), _ptr(w), _ptr(g), n, s, _ptr(out), _stream()))
    return out


@distortion_bwd.register_fake
def _(t, w, g):
    return torch.empty_like(w)


def _dist_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _dist_backward(ctx, g):
    t, w = ctx.saved_tensors
    return None, distortion_bwd(t, w, _f32c(g))


torch.library.register_autograd(f'{NS}::distortion', _dist_backward, setup_context=_dist_setup)


# ------------------------------------------------------------------------------------------
# orientation + predicted-normal losses, per ray (train_utils.py:165-204)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::normal_losses', mutates_args=(), device_types='cuda')
def normal_losses(weights: Tensor, normals: Tensor, normals_pred: Tensor, viewdirs: Tensor, ori_is_pred: bool) -> Tensor:
    """weights [N,S]; normals [N,S,3] (empty = absent) and normals_pred [N,S,3]; viewdirs [N,3] ->
    [2,N]: row 0 = sum_s w min(0, n_t.(-v))^2, row 1 = sum_s w (1 - n.n_pred)."""
    lib = _lib.load()
    n, s = weights.shape
    out = torch.empty((2, n), device=weights.device, dtype=torch.float32)
    _lib.check(lib.rn_normal_losses_fwd(_ptr(weights), _ptr(normals), _ptr(normals_pred), _ptr(viewdirs), n, s,
                                        int(ori_is_pred), _ptr(out[0]), _ptr(out[1]), _stream()))
    return out


@normal_losses.register_fake
def _(weights, normals, normals_pred, viewdirs, ori_is_pred):
    return weights.new_empty((2, weights.shape[0]))


@torch.library.custom_op(f'{NS}::normal_losses_bwd', mutates_args=(), device_types='cuda')
def normal_losses_bwd(weights: Tensor, normals: Tensor, normals_pred: Tensor, viewdirs: Tensor, g: Tensor,
                      ori_is_pred: bool) -> List[Tensor]:
    lib = _lib.load()
    n, s = weights.shape
    d_w, d_np = torch.empty_like(weights), torch.empty_like(normals_pred)
    _lib.check(lib.rn_normal_losses_bwd(_ptr(weights), _ptr(normals), _ptr(normals_pred), _ptr(viewdirs), _ptr(g[0]),
                                        _ptr(g[1]), n, s, int(ori_is_pred), _ptr(d_w), _ptr(d_np), _stream()))
    return [d_w, d_np]


@normal_losses_bwd.register_fake
def _(weights, normals, normals_pred, viewdirs, g, ori_is_pred):
    return [torch.empty_like(weights), torch.empty_like(normals_pred)]


def _nl_setup(ctx, inputs, output):
    weights, normals, normals_pred, viewdirs, ori_is_pred = inputs
    ctx.save_for_backward(weights, normals, normals_pred, viewdirs)
    ctx.ori_is_pred = ori_is_pred


def _nl_backward(ctx, g):
    weights, normals, normals_pred, viewdirs = ctx.saved_tensors
    d_w, d_np = normal_losses_bwd(weights, normals, normals_pred, viewdirs, _f32c(g), ctx.ori_is_pred)
    return d_w, None, d_np, None, None


torch.library.register_autograd(f'{NS}::normal_losses', _nl_backward, setup_context=_nl_setup)


# ------------------------------------------------------------------------------------------
# data loss of one level (train_utils.py:33-88; the loss epilogue of SURVEY 8(f) rank 2)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::data_loss_sums', mutates_args=(), device_types='cuda')
def data_loss_sums(rgb: Tensor, gt: Tensor, lossmult: Tensor, charb: bool, charb_padding: float) -> Tensor:
    """rgb, gt [N,3]; lossmult [N] (empty = 1) -> [3] = (sum lm r^2, sum lm term(r), sum lm), r = rgb - gt."""
    lib = _lib.load()
    out = torch.empty((3,), device=rgb.device, dtype=torch.float32)
    _lib.check(lib.rn_data_loss_fwd(_ptr(rgb), _ptr(gt), _ptr(lossmult) if lossmult.numel() else None, rgb.shape[0],
                                    int(charb), charb_padding, _ptr(out), _stream()))
    return out


@data_loss_sums.register_fake
def _(rgb, gt, lossmult, charb, charb_padding):
    return rgb.new_empty((3,))


@torch.library.custom_op(f'{NS}::data_loss_sums_bwd', mutates_args=(), device_types='cuda')
def data_loss_sums_bwd(rgb: Tensor, gt: Tensor, lossmult: Tensor, g: Tensor, charb: bool, charb_padding: float) -> Tensor:
    lib = _lib.load()
    d_rgb = torch.empty_like(rgb)
    _lib.check(lib.rn_data_loss_bwd(_ptr(rgb), _ptr(gt), _ptr(lossmult) if lossmult.numel() else None, _ptr(g), rgb.shape[0],
                                    int(charb), charb_padding, _ptr(d_rgb), _stream()))
    return d_rgb


@data_loss_sums_bwd.register_fake
def _(rgb, gt, lossmult, g, charb, charb_padding):
    return torch.empty_like(rgb)


def _dl_setup(ctx, inputs, output):
    rgb, gt, lossmult, charb, pad = inputs
    ctx.save_for_backward(rgb, gt, lossmult)
    ctx.charb, ctx.pad = charb, pad


def _dl_backward(ctx, g):
    rgb, gt, lossmult = ctx.saved_tensors
    return data_loss_sums_bwd(rgb, gt, lossmult, _f32c(g), ctx.charb, ctx.pad), None, None, None, None


torch.library.register_autograd(f'{NS}::data_loss_sums', _dl_backward, setup_context=_dl_setup)


# ------------------------------------------------------------------------------------------
# unit-level encoders
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::encode', mutates_args=(), device_types='cuda')
def encode(tdist: Tensor, origins: Tensor, dirs: Tensor, radii: Tensor) -> Tensor:
    lib = _lib.load()
    n, s1 = tdist.shape
    out = torch.empty((n, s1 - 1, 96), device=tdist.device, dtype=torch.float32)
    _lib.check(lib.rn_encode(_ptr(tdist), _ptr(origins), _ptr(dirs), _ptr(radii), n, s1 - 1, _ptr(out), _stream()))
    return out


@torch.library.custom_op(f'{NS}::ide', mutates_args=(), device_types='cuda')
def ide(dirs: Tensor, kappa_inv: Tensor) -> Tensor:
    lib = _lib.load()
    n = dirs.shape[0]
    out = torch.empty((n, 72), device=dirs.device, dtype=torch.float32)
    _lib.check(lib.rn_ide(_ptr(dirs), _ptr(kappa_inv), n, _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------
# NerfMLP
# ------------------------------------------------------------------------------------------
def _mlp_config(prec, srgb_mapping, srgb_norm, density_bias, roughness_bias, rgb_premultiplier, rgb_bias, rgb_padding,
                chunk_rows, gemm_impl):
    # bit 8 of `gemm_impl` carries NerfMLP.deterministic_wgrad through the op schema
    return _lib.RnMlpConfig(prec, int(srgb_mapping), int(srgb_norm), density_bias, roughness_bias, rgb_premultiplier,
                            rgb_bias, rgb_padding, chunk_rows, gemm_impl & 0xff, (gemm_impl >> 8) & 1)


def default_chunk_rows(n_rows, training=False):
    """Rows per internal chunk.  Training keeps whole levels in flight (the saved-activation region is sized for all
    rows anyway, and the wgrad kernels amortise their fp32 reduction over the rows of one launch); eval only needs
    enough rows to fill the machine."""
    rows = (n_rows + 127) // 128 * 128
    return max(128, min(rows, 2097152 if training else 1048576))


# Training keeps every activation of the forward for the backward (no recompute) while that fits comfortably in
# HBM; beyond this many bytes per level the backward recomputes chunk by chunk instead.
SAVED_BYTES_CAP = 48 << 30


@torch.library.custom_op(f'{NS}::mlp_pack', mutates_args=(), device_types='cuda')
def mlp_pack(params: Sequence[Tensor], prec: int) -> Tensor:
    lib = _lib.load()
    assert len(params) == _lib.NUM_PARAMS
    ps = [_f32c(p.detach()) for p in params]
    for i, p in enumerate(ps):
        assert p.numel() == lib.rn_mlp_param_numel(i), f'parameter {i} has {p.numel()} elements'
    blob = torch.empty((lib.rn_mlp_packed_bytes(prec),), device=ps[0].device, dtype=torch.uint8)
    arr = (ctypes.c_void_p * _lib.NUM_PARAMS)(*[p.data_ptr() for p in ps])
    _lib.check(lib.rn_mlp_pack(arr, _ptr(blob), prec, _stream()))
    return blob


@torch.library.custom_op(f'{NS}::param_carrier', mutates_args=(), device_types='cuda')
def param_carrier(params: Sequence[Tensor]) -> Tensor:
    """The autograd edge from the 46 parameters to the fused MLP calls of one step, as ONE flat fp32 tensor (its VALUES are
    never read -- the kernels read the packed weights).  Every `mlp_forward` of the step takes the same carrier, so
    autograd sums the per-call flat parameter gradients with one add, and the carrier's backward hands every parameter
    a view of that sum: no per-parameter accumulation kernels, and all `.grad`s tile one buffer (what
    `parallel.GradAllReducer` runs its collective on)."""
    return torch.empty((sum(_param_sizes()),), device=params[0].device, dtype=torch.float32)


@param_carrier.register_fake
def _(params):
    return params[0].new_empty((sum(p.numel() for p in params),))


def _pc_setup(ctx, inputs, output):
    ctx.param_shapes = [p.shape for p in inputs[0]]


def _pc_backward(ctx, g):
    g = _f32c(g)
    return ([t.view(shape) for t, shape in zip(g.split(_param_sizes()), ctx.param_shapes)],)


torch.library.register_autograd(f'{NS}::param_carrier', _pc_backward, setup_context=_pc_setup)


@torch.library.custom_op(f'{NS}::mlp_forward', mutates_args=(), device_types='cuda')
def mlp_forward(tdist: Tensor, origins: Tensor, dirs: Tensor, viewdirs: Tensor, radii: Tensor, carrier: Tensor,
                packed: Tensor, training: bool, prec: int, srgb_mapping: bool, srgb_norm: bool, density_bias: float,
                roughness_bias: float, rgb_premultiplier: float, rgb_bias: float, rgb_padding: float, chunk_rows: int,
                gemm_impl: int, keep_saved: bool) -> List[Tensor]:
    """-> [density [N,S], rgb, normals (empty in eval), normals_pred, grad_pred, tint, diffuse, specular [N,S,3],
    roughness [N,S,1], saved (uint8: activations kept for the backward; empty in eval / when over the cap / when
    `keep_saved` is False, e.g. a training-mode forward under torch.no_grad())].
    `carrier` (ops.param_carrier) only carries the autograd edge to the parameters; the arithmetic reads `packed`."""
    lib = _lib.load()
    n, s1 = tdist.shape
    s = s1 - 1
    dev = tdist.device
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    density, rough = f(n, s), f(n, s, 1)
    rgb, npred, gpred, tint, diffuse, spec = (f(n, s, 3) for _ in range(6))
    normals = f(n, s, 3) if training else f(0)
    if chunk_rows <= 0:
        chunk_rows = default_chunk_rows(n * s, training)
    cfg = _mlp_config(prec, srgb_mapping, srgb_norm, density_bias, roughness_bias, rgb_premultiplier, rgb_bias,
                      rgb_padding, chunk_rows, gemm_impl)
    saved_bytes = lib.rn_mlp_saved_bytes(ctypes.byref(cfg), n * s) if (training and keep_saved) else 0
    if saved_bytes > SAVED_BYTES_CAP:
        saved_bytes = 0
    if saved_bytes:   # bound by what is actually free as well (the backward then recomputes chunk by chunk)
        cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        if saved_bytes > cached:
            # not served from the caching allocator's pool (first steps only): ask the driver.  cudaMemGetInfo costs 3 ms
            # and more of host time -- at the head of a step that is GPU idle time whenever the host is not running ahead
            free_b, _ = torch.cuda.mem_get_info(dev)
            if saved_bytes > 0.8 * (free_b + cached):
                saved_bytes = 0
    saved = torch.empty((saved_bytes,), device=dev, dtype=torch.uint8)
    ws_bytes = lib.rn_mlp_workspace_bytes(ctypes.byref(cfg), (2 if saved_bytes else 1) if training else 0)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    outs = _lib.RnMlpOutputs(density.data_ptr(), rgb.data_ptr(), normals.data_ptr() if training else None,
                             npred.data_ptr(), gpred.data_ptr(), tint.data_ptr(), diffuse.data_ptr(), spec.data_ptr(),
                             rough.data_ptr())
    _lib.check(lib.rn_mlp_forward(ctypes.byref(cfg), _ptr(packed), _ptr(tdist), _ptr(origins), _ptr(dirs),
                                  _ptr(viewdirs), _ptr(radii), n, s, ctypes.byref(outs), _ptr(ws), ws_bytes,
                                  _ptr(saved) if saved_bytes else None, saved_bytes, _stream()))
    return [density, rgb, normals, npred, gpred, tint, diffuse, spec, rough, saved]


@mlp_forward.register_fake
def _(tdist, origins, dirs, viewdirs, radii, carrier, packed, training, prec, srgb_mapping, srgb_norm, density_bias,
      roughness_bias, rgb_premultiplier, rgb_bias, rgb_padding, chunk_rows, gemm_impl, keep_saved):
    n, s1 = tdist.shape
    s = s1 - 1
    f = lambda *shape: tdist.new_empty(shape)
    return [f(n, s), f(n, s, 3), f(n, s, 3) if training else f(0), f(n, s, 3), f(n, s, 3), f(n, s, 3), f(n, s, 3),
            f(n, s, 3), f(n, s, 1), tdist.new_empty((0,), dtype=torch.uint8)]


_PARAM_SIZES = []


def _param_sizes():
    if not _PARAM_SIZES:
        lib = _lib.load()
        _PARAM_SIZES.extend(int(lib.rn_mlp_param_numel(i)) for i in range(_lib.NUM_PARAMS))
    return _PARAM_SIZES


@torch.library.custom_op(f'{NS}::mlp_backward', mutates_args=(), device_types='cuda')
def mlp_backward(tdist: Tensor, origins: Tensor, dirs: Tensor, viewdirs: Tensor, radii: Tensor, packed: Tensor,
                 saved: Tensor, grads: Sequence[Tensor], prec: int, srgb_mapping: bool, srgb_norm: bool, density_bias: float,
                 roughness_bias: float, rgb_premultiplier: float, rgb_bias: float, rgb_padding: float, chunk_rows: int,
                 gemm_impl: int) -> Tensor:
    """grads: [g_density, g_rgb, g_normals_pred, g_grad_pred, g_tint, g_diffuse, g_specular, g_roughness] (empty = 0)
    -> the parameter gradients in rn order, concatenated in one flat fp32 tensor (one zero fill instead of 46; the
    autograd wrapper hands out per-parameter views)."""
    lib = _lib.load()
    n, s1 = tdist.shape
    s = s1 - 1
    dev = tdist.device
    if chunk_rows <= 0:
        chunk_rows = default_chunk_rows(n * s, True)
    cfg = _mlp_config(prec, srgb_mapping, srgb_norm, density_bias, roughness_bias, rgb_premultiplier, rgb_bias,
                      rgb_padding, chunk_rows, gemm_impl)
    saved_bytes = saved.numel()
    ws_bytes = lib.rn_mlp_workspace_bytes(ctypes.byref(cfg), 2 if saved_bytes else 1)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    g = [_f32c(x) for x in grads]
    pg = lambda t: t.data_ptr() if t.numel() else None
    gs = _lib.RnMlpOutputs(pg(g[0]), pg(g[1]), None, pg(g[2]), pg(g[3]), pg(g[4]), pg(g[5]), pg(g[6]), pg(g[7]))
    sizes = _param_sizes()
    flat = torch.zeros((sum(sizes),), device=dev, dtype=torch.float32)
    arr = (ctypes.c_void_p * _lib.NUM_PARAMS)(*[t.data_ptr() for t in flat.split(sizes)])
    _lib.check(lib.rn_mlp_backward(ctypes.byref(cfg), _ptr(packed), _ptr(tdist), _ptr(origins), _ptr(dirs),
                                   _ptr(viewdirs), _ptr(radii), n, s, ctypes.byref(gs), arr, _ptr(ws), ws_bytes,
                                   _ptr(saved) if saved_bytes else None, saved_bytes, _stream()))
    return flat


@mlp_backward.register_fake
def _(tdist, origins, dirs, viewdirs, radii, packed, saved, grads, prec, srgb_mapping, srgb_norm, density_bias, roughness_bias,
      rgb_premultiplier, rgb_bias, rgb_padding, chunk_rows, gemm_impl):
    return tdist.new_empty((sum(_param_sizes()),))


def _mlp_setup(ctx, inputs, output):
    (tdist, origins, dirs, viewdirs, radii, carrier, packed, training, *scalars) = inputs
    ctx.save_for_backward(tdist, origins, dirs, viewdirs, radii, packed, output[9])
    ctx.scalars = scalars[:-1]   # (keep_saved belongs to the forward only)
    ctx.mark_non_differentiable(output[2], output[9])  # density-gradient normals are a constant
    ctx.set_materialize_grads(False)


def _mlp_backward(ctx, grads):
    tdist, origins, dirs, viewdirs, radii, packed, saved = ctx.saved_tensors
    empty = tdist.new_empty((0,))
    order = (0, 1, 3, 4, 5, 6, 7, 8)  # density, rgb, normals_pred, grad_pred, tint, diffuse, specular, roughness
    g = [grads[i] if grads[i] is not None else empty for i in order]
    flat = mlp_backward(tdist, origins, dirs, viewdirs, radii, packed, saved, g, *ctx.scalars)
    return (None, None, None, None, None, flat, None, None) + (None,) * (len(ctx.scalars) + 1)


torch.library.register_autograd(f'{NS}::mlp_forward', _mlp_backward, setup_context=_mlp_setup)


# ------------------------------------------------------------------------------------------
# GEMM test / bench surface
# ------------------------------------------------------------------------------------------
def gemm_test(a, b, bias=None, relu=False, prec=_lib.PREC_BF16, impl=0):
    lib = _lib.load()
    m, k = a.shape
    n = b.shape[0]
    c = torch.empty((m, n), device=a.device, dtype=torch.float32)
    nb = lib.rn_gemm_scratch_bytes(m, n, k)
    scratch = torch.empty((nb,), device=a.device, dtype=torch.uint8)
    _lib.check(lib.rn_gemm_test(_ptr(a), _ptr(b), _ptr(bias) if bias is not None else None, m, n, k, int(relu), prec,
                                impl, _ptr(c), _ptr(scratch), nb, _stream()))
    return c


def wgrad_test(dy, x, prec=_lib.PREC_BF16, impl=0):
    lib = _lib.load()
    m, n = dy.shape
    k = x.shape[1]
    c = torch.empty((n, k), device=dy.device, dtype=torch.float32)
    nb = lib.rn_gemm_scratch_bytes(m, n, k)
    scratch = torch.empty((nb,), device=dy.device, dtype=torch.uint8)
    _lib.check(lib.rn_wgrad_test(_ptr(dy), _ptr(x), m, n, k, prec, impl, _ptr(c), _ptr(scratch), nb, _stream()))
    return c


def gemm_bench(m, prec=_lib.PREC_BF16, impl=0, iters=20, device='cuda'):
    """ms per launch of one 256->256 layer GEMM (+bias/ReLU epilogue) over m rows."""
    lib = _lib.load()
    nb = lib.rn_gemm_scratch_bytes(m, 256, 256) + 2 * ((m * 256 * 4 + 255) // 256 * 256) + 4096
    scratch = torch.empty((nb,), device=device, dtype=torch.uint8)
    ms = ctypes.c_float(0)
    _lib.check(lib.rn_gemm_bench(m, prec, impl, iters, ctypes.byref(ms), _ptr(scratch), nb, _stream()))
    return ms.value


# ------------------------------------------------------------------------------------------
# ray generation (camera_utils.pixels_to_rays); no gradient (pixels and cameras are data)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{NS}::pixels_to_rays', mutates_args=(), device_types='cuda')
def pixels_to_rays(pix_x: Tensor, pix_y: Tensor, cam_idx: Tensor, pixtocams: Tensor, camtoworlds: Tensor,
                   pixtocam_ndc: Tensor) -> List[Tensor]:
    """pix_x, pix_y, cam_idx: int32 [n]; pixtocams [C,3,3], camtoworlds [C,3,4] fp32; pixtocam_ndc [3,3] or empty
    -> [origins, directions, viewdirs [n,3], radii [n,1], imageplane [n,2]]"""
    lib = _lib.load()
    n = pix_x.numel()
    dev = pix_x.device
    i32 = lambda t: t.contiguous().to(torch.int32)
    px, py, ci = i32(pix_x), i32(pix_y), i32(cam_idx)
    p2c, c2w = _f32c(pixtocams), _f32c(camtoworlds)
    ndc = _f32c(pixtocam_ndc) if pixtocam_ndc.numel() else pixtocam_ndc
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    o, d, v, r, ip = f(n, 3), f(n, 3), f(n, 3), f(n, 1), f(n, 2)
    _lib.check(lib.rn_pixels_to_rays(_ptr(px), _ptr(py), _ptr(ci), _ptr(p2c), _ptr(c2w), _ptr(ndc) if ndc.numel() else None,
                                     n, _ptr(o), _ptr(d), _ptr(v), _ptr(r), _ptr(ip), _stream()))
    return [o, d, v, r, ip]


@pixels_to_rays.register_fake
def _(pix_x, pix_y, cam_idx, pixtocams, camtoworlds, pixtocam_ndc):
    n = pix_x.numel()
    f = lambda *shape: pixtocams.new_empty(shape)
    return [f(n, 3), f(n, 3), f(n, 3), f(n, 1), f(n, 2)]


def fp16_saturation_count(reset=True):
    """fp16 precision mode: number of saturated (clamped to +-65504) gradient-tile observations since the last reset
    (`rn_fp16_saturation_count`).  0 means the per-chain power-of-two gradient scales had enough head-room; anything else
    means some weight gradients of the steps since the last reset are biased -- skip those steps or switch the model to
    `precision='bf16x3'` / `'bf16'`.  Synchronises the device, so poll it every few hundred steps, not every step."""
    return int(_lib.load().rn_fp16_saturation_count(int(bool(reset))))

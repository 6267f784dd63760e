"""Function-level surface of the reference's `internal/render.py`, backed by the CUDA ops."""
import torch

from . import ops
from .models import CastRays, _render_tone_map


def cast_rays(tdist, origins, directions, radii, ray_shape='cone', diag=True):
    """render.py:105-129.  Returns the ray-segment description consumed by the fused encode kernel
    (`MLP.__call__`, `ops.encode`); the Gaussians are never materialised in HBM."""
    if ray_shape == 'cylinder':
        raise NotImplementedError("only ray_shape='cone' runs on the CUDA path")
    if ray_shape != 'cone':
        raise ValueError('ray_shape must be \'cone\' or \'cylinder\'')
    return CastRays(tdist, origins, directions, radii)


def compute_alpha_weights(density, tdist, dirs, opaque_background=False):
    """render.py:132-149 -> (weights,) computed by the compositing kernel (colours zero)."""
    if opaque_background:
        raise NotImplementedError('opaque_background is not used by the Ref-NeRF configs')
    lead = density.shape[:-1]
    s = density.shape[-1]
    d = ops._f32c(density.reshape(-1, s))
    z = torch.zeros(d.shape + (3,), device=d.device)
    e = d.new_empty((0,))
    far = torch.zeros((d.shape[0], 1), device=d.device)
    w = ops.composite_fwd(d, ops._f32c(tdist.reshape(-1, s + 1)), ops._f32c(dirs.reshape(-1, 3)), far, z, z, z, e, e, e,
                          e, 0.0, False)[0]
    return (w.reshape(lead + (s,)),)


def volumetric_rendering(rgbs, diffuse_rgbs, specular_rgbs, density, tdist, dirs, bg_rgbs, t_far, compute_extras,
                         extras=None, srgb_mapping='none'):
    """render.py:152-254 fused with compute_alpha_weights (takes density + dirs instead of weights)."""
    lead = density.shape[:-1]
    s = density.shape[-1]
    f = lambda t, *c: ops._f32c(t.reshape((-1,) + c))
    ex = extras or {}
    e = density.new_empty((0,))
    get = lambda k, *c: f(ex[k], *c) if ex.get(k) is not None else e
    w, comp, exo, pct = ops.composite_fwd(f(density, s), f(tdist, s + 1), f(dirs, 3), f(t_far, 1), f(rgbs, s, 3),
                                          f(diffuse_rgbs, s, 3), f(specular_rgbs, s, 3), get('normals', s, 3),
                                          get('normals_pred', s, 3), get('roughness', s, 1), get('tint', s, 3),
                                          float(bg_rgbs), bool(compute_extras))
    rgb, dif, spe = _render_tone_map(comp[:, 0:3], comp[:, 3:6], comp[:, 6:9], srgb_mapping)
    shp = lambda t, *tail: t.reshape(lead + tail)
    out = dict(rgb=shp(rgb, 3), diffuse=shp(dif, 3), specular=shp(spe, 3), distance=shp(comp[:, 9], 1),
               acc=shp(comp[:, 10]), weights=shp(w, s))
    if compute_extras:
        if ex.get('normals') is not None:
            out['normals'] = shp(exo[:, 0:3], 3)
        out['normals_pred'] = shp(exo[:, 3:6], 3)
        out['tint'] = shp(exo[:, 6:9], 3)
        out['roughness'] = shp(exo[:, 9], 1)
        out['distance_mean'] = shp(comp[:, 11])
        out['distance_percentile_5'], out['distance_median'], out['distance_percentile_95'] = (
            shp(pct[:, 0]), shp(pct[:, 1]), shp(pct[:, 2]))
    return out

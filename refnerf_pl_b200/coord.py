"""Function-level surface of the reference's `internal/coord.py` used by the hot path."""
from . import ops


def construct_ray_warps(fn, t_near, t_far):
    """coord.py:63-99 with fn=None (every shipped config)."""
    if fn is not None:
        raise NotImplementedError('raydist_fn is None in every Ref-NeRF config')
    t_to_s = lambda t: (t - t_near) / (t_far - t_near)
    s_to_t = lambda s: s * t_far + (1 - s) * t_near
    return t_to_s, s_to_t


def cast_and_encode(tdist, origins, directions, radii):
    """cast_rays (render.py:105-129) + lift_and_diagonalize (coord.py:129-133) + integrated_pos_enc
    (coord.py:107-126, degrees 0..15) in one kernel -> [..., S, 96]."""
    lead = tdist.shape[:-1]
    s = tdist.shape[-1] - 1
    f = lambda t, c: ops._f32c(t.reshape(-1, c))
    return ops.encode(f(tdist, s + 1), f(origins, 3), f(directions, 3), f(radii, 1)).reshape(lead + (s, 96))

"""Function-level surface of the reference's `internal/coord.py` used by the hot path."""
import torch

from . import ops


def contract(x):
    """coord.py:20-26: the mip-NeRF 360 scene contraction, z = x inside the unit ball, ((2 sqrt(m) - 1) / m) x outside
    (m = max(eps, |x|^2)).  The reference's own function raises TypeError (`torch.max(float, tensor)`) and is
    never called (`MLP.warp_fn` is stored and unused, models.py:467), so this follows its FORMULA; parity unpinned.
    Plain torch on whatever device `x` lives on: nothing on the hot path calls it, so there is nothing to fuse it into."""
    eps = torch.finfo(torch.float32).eps
    x_mag_sq = torch.clamp(torch.sum(x ** 2, dim=-1, keepdim=True), min=eps)
    return torch.where(x_mag_sq <= 1, x, ((2 * torch.sqrt(x_mag_sq) - 1) / x_mag_sq) * x)


def inv_contract(z):
    """coord.py:29-35: inverse of `contract`."""
    eps = torch.finfo(torch.float32).eps
    z_mag_sq = torch.clamp(torch.sum(z ** 2, dim=-1, keepdim=True), min=eps)
    return torch.where(z_mag_sq <= 1, z, z / (2 * torch.sqrt(z_mag_sq) - z_mag_sq))


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

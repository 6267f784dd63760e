"""Noisy-ray sampler of the geometry-loss configuration (`internal/sample_utils.py:4-80`, called from
`nerf_system.py:116-133`): the first `sample_noise_size` rays of the batch are re-cast from `sample_noise_angles`
slightly rotated directions through the point they hit (the fine level's composited distance), giving the extra rays
whose renderings the consistency losses compare with the originals.

Runs on the device the rays live on (a handful of [n,3] torch ops, no host round trip, no gradient).  The new rays
depend on the main call's `distance`, so the second `Model` call cannot be merged into the first one.
"""
import numpy as np
import torch

from . import utils


def euler_angles_to_matrix(euler_angles):
    """sample_utils.py:5-38: R = Rx(a) Ry(b) Rz(c) for angles [..., 3] in radians."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    a, b, c = torch.unbind(euler_angles, -1)
    one, zero = torch.ones_like(a), torch.zeros_like(a)

    def rot(flat):
        return torch.stack(flat, -1).reshape(a.shape + (3, 3))

    rx = rot((one, zero, zero, zero, torch.cos(a), -torch.sin(a), zero, torch.sin(a), torch.cos(a)))
    ry = rot((torch.cos(b), zero, torch.sin(b), zero, one, zero, -torch.sin(b), zero, torch.cos(b)))
    rz = rot((torch.cos(c), -torch.sin(c), zero, torch.sin(c), torch.cos(c), zero, zero, zero, one))
    return torch.matmul(torch.matmul(rx, ry), rz)


@torch.no_grad()
def sample_noisy_rays(rays, rendering, sample_angle_range=0., sample_noise_size=128, sample_noise_angles=1,
                      warmup_ratio=1., xyz_angles=None):
    """sample_utils.py:40-80.  `xyz_angles` [sample_noise_angles, 3] (radians) replaces the uniform draw of
    sample_utils.py:49-50 (tests pin the reference's draw that way)."""
    dev = rendering['distance'].device
    if xyz_angles is None:
        xyz_angles = torch.zeros(sample_noise_angles * 3, device=dev).uniform_(
            0, sample_angle_range / 180 * np.pi * warmup_ratio).reshape(-1, 3)
    xyz_angles = torch.as_tensor(xyz_angles, dtype=torch.float32, device=dev).reshape(-1, 3)
    if xyz_angles.shape[0] != sample_noise_angles:
        raise ValueError('xyz_angles must hold one rotation per noise angle')
    rot = euler_angles_to_matrix(xyz_angles)                       # [A, 3, 3]
    n = min(sample_noise_size, len(rendering['distance']))
    take = lambda v: torch.as_tensor(v, device=dev)[:n]
    rep = lambda v: torch.cat([take(v)] * sample_noise_angles)
    distance = rep(rendering['distance'])
    if distance.dim() == rays.origins.dim() - 1:
        distance = distance[..., None]
    elif distance.dim() != rays.origins.dim():
        raise ValueError('The dimension of distance is wrong.')
    viewdirs_ = torch.cat([take(rays.viewdirs).float() @ r.T for r in rot])
    directions_ = torch.cat([take(rays.directions).float() @ r.T for r in rot])
    origins, directions = rep(rays.origins).float(), rep(rays.directions).float()
    origins_ = origins + distance * directions - distance * directions_
    return utils.Rays(origins=origins_, directions=directions_, viewdirs=viewdirs_, radii=rep(rays.radii),
                      imageplane=rep(rays.imageplane), lossmult=rep(rays.lossmult), near=rep(rays.near),
                      far=rep(rays.far), cam_idx=rep(rays.cam_idx))

"""Function-level surface of the reference's `internal/ref_utils.py`."""
import torch

from . import ops


def l2_normalize(x, eps=torch.finfo(torch.float32).eps):
    """ref_utils.py:40-42."""
    return x / torch.sqrt(torch.clamp(torch.sum(x ** 2, dim=-1, keepdim=True), min=eps))


def reflect(viewdirs, normals):
    """ref_utils.py:22-37."""
    return 2.0 * torch.sum(normals * viewdirs, dim=-1, keepdim=True) * normals - viewdirs


def compute_weighted_mae(weights, normals, normals_gt):
    """ref_utils.py:45-50."""
    one_eps = 1 - torch.finfo(torch.float32).eps
    return (weights * torch.arccos(torch.clip((normals * normals_gt).sum(-1), -one_eps, one_eps))).sum() / weights.sum() * 180.0 / torch.pi


def generate_ide_fn(deg_view):
    """ref_utils.py:98-161: returns f(xyz [...,3], kappa_inv [...,1]) -> [...,72] on the CUDA kernel."""
    if deg_view != 5:
        raise NotImplementedError('the CUDA integrated directional encoding is built for deg_view=5')

    def integrated_dir_enc_fn(xyz, kappa_inv):
        lead = xyz.shape[:-1]
        out = ops.ide(ops._f32c(xyz.reshape(-1, 3)), ops._f32c(kappa_inv.reshape(-1)))
        return out.reshape(lead + (72,))

    return integrated_dir_enc_fn

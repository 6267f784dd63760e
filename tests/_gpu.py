"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle / golden fixtures)."""
import numpy as np
import torch

from oracle import refnerf_oracle as O
from refnerf_pl_b200 import configs, models, utils

DEV = 'cuda'


def to_dev(d):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else torch.tensor(v, device=DEV)) for k, v in d.items()}


def build_model(precision, gin=None, mlp_kwargs=None, model_kwargs=None, config_kwargs=None):
    """Model with the blender_refnerf.gin bindings (restated here: the gin files live in the reference tree,
    which does not exist on the GPU box)."""
    configs.clear_bindings()
    configs.bind('Model', num_levels=2, single_mlp=True, num_prop_samples=128, num_nerf_samples=128, anneal_slope=0.,
                 dilation_multiplier=0., dilation_bias=0., single_jitter=False, resample_padding=0.01)
    configs.bind('NerfMLP', net_depth=8, net_width=256, net_depth_viewdirs=8, net_width_viewdirs=256,
                 basis_shape='octahedron', basis_subdivisions=1, disable_density_normals=False, enable_pred_normals=True,
                 use_directional_enc=True, use_reflections=True, deg_view=5, enable_pred_roughness=True,
                 use_diffuse_color=True, use_specular_tint=True, use_n_dot_v=True, bottleneck_width=128,
                 bottleneck_noise=0.0, density_bias=0.5, max_deg_point=16)
    configs.bind('Config', data_loss_type='mse', orientation_loss_mult=0.1, predicted_normal_loss_mult=3e-4,
                 orientation_coarse_loss_mult=0.01, predicted_normal_coarse_loss_mult=3e-5, interlevel_loss_mult=0.0,
                 data_coarse_loss_mult=0.1, data_loss_mult=1.0, batch_size=1024, render_chunk_size=4096)
    configs.bind('NerfMLP', precision=precision, **(mlp_kwargs or {}))
    if model_kwargs:
        configs.bind('Model', **model_kwargs)
    if config_kwargs:
        configs.bind('Config', **config_kwargs)
    cfg = configs.Config()
    model = models.Model(config=cfg).to(DEV)
    return model, cfg


def load_params(model, p):
    sd = {k: v.clone() for k, v in p.items()}
    missing = model.nerf_mlp.load_state_dict(sd, strict=True)
    return missing


def rays_obj(rays_dict):
    return utils.Rays(**{k: (v if isinstance(v, torch.Tensor) else torch.tensor(v)) for k, v in rays_dict.items()}).to(DEV)


def rel_err(a, b):
    """|a-b| / (|b| + 1e-3 max|b|): per-element relative error that stays meaningful near zero."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / (np.abs(b) + 1e-3 * max(np.abs(b).max(), 1e-30))


def norm_rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)

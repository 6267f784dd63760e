"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle / golden fixtures)."""
import numpy as np
import torch

from oracle import refnerf_oracle as O
from refnerf_pl_b200 import configs, models, utils

DEV = 'cuda'


def to_dev(d):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else torch.tensor(v, device=DEV)) for k, v in d.items()}


GIN_FOR_CASE = {'blender_init': 'blender_refnerf.gin', 'blender_pert': 'blender_refnerf.gin',
                'blender_trained': 'blender_refnerf.gin',
                'llff_geom': 'llff_refnerf_geometry_losses.gin'}


def build_model(precision, gin='blender_refnerf.gin', mlp_kwargs=None, model_kwargs=None, config_kwargs=None):
    """Model driven by this repo's configs/*.gin (same binding names/values as the reference's files)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    configs.clear_bindings()
    configs.parse_gin_files_and_bindings([os.path.join(root, 'configs', gin)])
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

"""`Config` with the field names of the reference's `internal/configs.py:28-172` that the hot path and
its losses read, plus a small gin-compatible binding registry so the shipped `configs/*.gin` files
(`Config.x = ..`, `Model.x = ..`, `NerfMLP.x = ..`) drive this implementation unchanged.

If the real `gin` package is importable it is used for `@configurable`; otherwise the registry below
provides the same constructor-kwarg injection for flat literal bindings.
"""
import ast
import dataclasses
import re
from typing import Optional

_BINDINGS = {}


def clear_bindings():
    _BINDINGS.clear()


def bind(scope, **kw):
    _BINDINGS.setdefault(scope, {}).update(kw)


def bindings(scope):
    return dict(_BINDINGS.get(scope, {}))


def configurable(cls):
    """Inject bound kwargs (by class name) at construction; explicit kwargs win."""
    orig = cls.__init__
    name = cls.__name__

    def init(self, *a, __orig=orig, __name=name, **k):
        __orig(self, *a, **{**_BINDINGS.get(__name, {}), **k})

    cls.__init__ = init
    return cls


def parse_gin_text(text):
    out = {}
    text = text.replace('\\\n', ' ')
    for line in text.splitlines():
        line = line.split('#', 1)[0].strip()
        m = re.match(r'^([A-Za-z_]\w*)\.([A-Za-z_]\w*)\s*=\s*(.+)$', line)
        if not m:
            continue
        scope, key, val = m.groups()
        try:
            val = ast.literal_eval(val.strip())
        except Exception:
            continue  # non-literal bindings (function references) are not used by the Ref-NeRF configs
        out.setdefault(scope, {})[key] = val
    return out


def parse_gin_files_and_bindings(files=(), extra_bindings=()):
    """Equivalent of gin.parse_config_files_and_bindings for flat literal bindings."""
    for f in files:
        with open(f) as fh:
            for scope, kw in parse_gin_text(fh.read()).items():
                bind(scope, **kw)
    for b in extra_bindings:
        for scope, kw in parse_gin_text(b).items():
            bind(scope, **kw)


@configurable
@dataclasses.dataclass(init=False)
class Config:
    """Subset of the reference Config consumed by the model, its losses, the optimiser and render_image."""
    seed: int = 20230227
    num_gpus: int = 1
    batch_size: int = 16384
    near: float = 2.
    far: float = 6.
    render_chunk_size: int = 16384
    vis_num_rays: int = 16
    disable_multiscale_loss: bool = False
    randomized: bool = True  # never read by the reference sampler either
    max_steps: int = 250000
    data_loss_type: str = 'charb'
    charb_padding: float = 0.001
    data_loss_mult: float = 1.0
    data_coarse_loss_mult: float = 0.
    interlevel_loss_mult: float = 1.0
    orientation_loss_mult: float = 0.0
    orientation_coarse_loss_mult: float = 0.0
    orientation_loss_target: str = 'normals_pred'
    predicted_normal_loss_mult: float = 0.0
    predicted_normal_coarse_loss_mult: float = 0.0
    distortion_loss_mult: float = 0.01  # never read by the reference
    patch_size: int = 1
    sample_angle_range: float = 5
    sample_noise_size: int = 128
    sample_noise_angles: int = 1
    consistency_warmup_steps: float = 0.
    consistency_decay_steps: float = 1.
    consistency_normal_loss_mult: float = 0.0
    consistency_normal_coarse_loss_mult: float = 0.0
    consistency_normal_loss_target: str = 'normals_pred'
    consistency_diffuse_loss_type: str = 'mse'
    consistency_diffuse_loss_mult: float = 0.0
    consistency_diffuse_coarse_loss_mult: float = 0.0
    consistency_specular_loss_type: str = 'mse'
    consistency_specular_loss_mult: float = 0.0
    consistency_specular_coarse_loss_mult: float = 0.0
    consistency_distance_loss_type: str = 'mse'
    consistency_distance_loss_mult: float = 0.0
    consistency_distance_coarse_loss_mult: float = 0.0
    accumulated_weights_loss_mult: float = 0.0
    depth_smoothness_loss_mult: float = 0.0
    depth_smoothness_coarse_loss_mult: float = 0.0
    acc_threshold_for_consistency_loss: float = 0.0
    weights_entropy_loss_mult: float = 0.0
    weights_entropy_coarse_loss_mult: float = 0.0
    acc_threshold_for_weights_entropy_loss: float = 0.0
    srgb_mapping_when_rendering: bool = False
    srgb_mapping_type: str = 'linear'
    supervised_by_linear_rgb: bool = False
    render_with_specular_density: bool = False
    compute_disp_metrics: bool = False
    compute_normal_metrics: bool = False
    lr_init: float = 0.002
    lr_final: float = 0.00002
    lr_delay_steps: int = 512
    lr_delay_mult: float = 0.01
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-6
    grad_max_norm: float = 0.001
    grad_max_val: float = 0.
    dataset_loader: str = 'llff'
    batching: str = 'all_images'
    extra: Optional[dict] = None  # unknown Config.* bindings are kept here instead of raising (skip_unknown)

    def __init__(self, **kw):
        known = {f.name for f in dataclasses.fields(self)}
        for f in dataclasses.fields(self):
            setattr(self, f.name, kw.get(f.name, f.default))
        self.extra = {k: v for k, v in kw.items() if k not in known}


def load_config(gin_files=(), gin_bindings=()):
    """configs.py:182-194 equivalent (without writing config.gin to disk)."""
    parse_gin_files_and_bindings(gin_files, gin_bindings)
    return Config()

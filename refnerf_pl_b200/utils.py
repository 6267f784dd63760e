"""Ray / batch containers with the field names and semantics of the reference's `internal/utils.py`
(`Rays` :51-93, `Batch` :110-117, `dummy_rays` :96-107): they are the argument types of the drop-in
boundary `Model.__call__(rays, train_frac, compute_extras)`."""
from dataclasses import dataclass, fields
from typing import Any, Optional

import numpy as np
import torch

RAY_FIELDS = ('origins', 'directions', 'viewdirs', 'radii', 'imageplane', 'lossmult', 'near', 'far', 'cam_idx')


@dataclass
class Rays:
    """All tensors must have the same num_dims and the first n-1 dims must match."""
    origins: Any
    directions: Any
    viewdirs: Any
    radii: Any
    imageplane: Any
    lossmult: Any
    near: Any
    far: Any
    cam_idx: Any

    def __getitem__(self, s):
        if isinstance(s, int):
            return Rays(*[[getattr(self, f.name)[s]] for f in fields(self)])
        if isinstance(s, slice):
            return Rays(*[getattr(self, f.name)[s] for f in fields(self)])
        raise ValueError('Argument to __getitem__ must be int or slice')

    def to(self, device):
        """In-place like the reference, but tensors on another device ARE moved (the reference drops
        the result of `.to`, utils.py:80-83, so there rays must already be resident)."""
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, np.ndarray):
                dt = torch.int32 if f.name == 'cam_idx' else torch.float32
                setattr(self, f.name, torch.tensor(v, dtype=dt, device=device))
            elif isinstance(v, torch.Tensor):
                if v.device != torch.device(device):
                    setattr(self, f.name, v.to(device))
            else:
                raise ValueError('Rays members must be either np.ndarray or torch.Tensor')
        return self

    def reshape(self, *dims):
        return Rays(*[getattr(self, f.name).reshape(*dims) for f in fields(self)])

    @property
    def shape(self):
        return self.origins.shape

    @classmethod
    def from_dict(cls, d, device=None):
        r = cls(**{k: d[k] for k in RAY_FIELDS})
        return r.to(device) if device is not None else r


def dummy_rays(device='cpu') -> Rays:
    z = lambda n: torch.zeros((1, n), device=device)
    return Rays(origins=z(3), directions=z(3), viewdirs=z(3), radii=z(1), imageplane=z(2), lossmult=z(1), near=z(1),
                far=z(1), cam_idx=z(1).type(torch.int32))


@dataclass
class Batch:
    rays: Any
    rgb: Optional[Any] = None
    disps: Optional[Any] = None
    normals: Optional[Any] = None
    alphas: Optional[Any] = None


def recursive_detach(v):
    if isinstance(v, torch.Tensor):
        return v.detach()
    if isinstance(v, (list, tuple)):
        return type(v)(recursive_detach(x) for x in v)
    return v


def merge_chunks(chunks):
    """Concatenate a list of per-chunk rendering dicts (utils.py:192-204)."""
    out = {}
    for k in chunks[0]:
        if isinstance(chunks[0][k], list):
            out[k] = [torch.cat([c[k][i] for c in chunks], dim=0) for i in range(len(chunks[0][k]))]
        else:
            out[k] = torch.cat([c[k] for c in chunks], dim=0)
    return out

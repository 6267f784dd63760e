"""Live pin of the oracle against the unmodified reference at /root/reference (skipped where the
reference tree is absent, e.g. the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import refnerf_oracle as O
from refnerf_pl_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_import.available(), reason='reference tree not present')


def test_forward_backward_live():
    ns, config = ref_import.load('blender_refnerf.gin')
    torch.manual_seed(0)
    model = ns.models.construct_model(ns.utils.dummy_rays(), config)
    p = O.init_params(seed=3, bias_std=0.1)
    model.nerf_mlp.load_state_dict({k: v.clone() for k, v in p.items()})
    n = 8
    r = synthetic.blender_rays(n, seed=21)
    rays_t = {k: torch.tensor(v) for k, v in r.items()}
    rays = ns.utils.Rays(**rays_t)
    model.train(True)
    rend_ref, hist_ref = model(rays, 1.0, True)
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rend, hist = O.model_forward(pp, rays_t, 1.0, True, True)
    for lvl in range(2):
        for k, v in rend_ref[lvl].items():
            assert float((v - rend[lvl][k]).abs().max()) <= 2e-6, k
        for k, v in hist_ref[lvl].items():
            if v is not None:
                assert float((v - hist[lvl][k]).abs().max()) <= (1e-4 if k == 'normals' else 2e-6), k


def test_sorted_interp_gather_is_bit_exact_live():
    ns, _ = ref_import.load('blender_refnerf.gin')
    g = torch.Generator().manual_seed(0)
    w = torch.rand(64, 128, generator=g) ** 6
    w[::2, 30:90] = 0
    w = w / w.sum(-1, keepdim=True)
    t = torch.sort(torch.rand(64, 129, generator=g), -1).values
    cw = ns.stepfun.integrate_weights(w)
    u = O.sample_grid(128).expand(64, 128)
    assert torch.equal(ns.math.sorted_interp(u, cw, t), O.sorted_interp(u, cw, t))

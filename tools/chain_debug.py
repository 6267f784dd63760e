"""Debug helper (GPU box): fused CTA-pair chain vs per-layer tcgen05 GEMMs on the same rays, per-output max |diff|."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import refnerf_oracle as O
from refnerf_pl_b200 import synthetic
from tests._gpu import build_model, load_params, rays_obj

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['eval', 'train']
p = O.init_params(seed=4, bias_std=0.1, weight_scale=1.3)
rays = synthetic.blender_rays(n, seed=9)
outs = {}
IMPL = int(__import__('os').environ.get('IMPL', '0'))
for impl in (2, IMPL):
    model, _ = build_model('bf16', mlp_kwargs=dict(gemm_impl=impl, chunk_rows=65536))
    load_params(model, p)
    for mode in modes:
        model.train(mode == 'train')
        with torch.no_grad():
            rend, hist = model(rays_obj(rays), 1.0, False)
        torch.cuda.synchronize()
        outs[(impl, mode)] = hist
        print('ran', impl, mode, flush=True)
for mode in modes:
    ha, hb = outs[(IMPL, mode)], outs[(2, mode)]
    for lvl in range(2):
        for k in ('density', 'rgb', 'roughness', 'normals_pred') + (('normals',) if mode == 'train' else ()):
            d = (ha[lvl][k] - hb[lvl][k]).abs()
            bad = (d > 0).reshape(d.shape[0], -1).any(-1) if d.dim() > 1 else d > 0
            print(f'{mode} L{lvl} {k:13s} max|diff| {float(d.max()):.3e}  rays with diffs {int(bad.sum())}/{d.shape[0]}'
                  f'  first bad ray {int(bad.nonzero()[0]) if bad.any() else -1}', flush=True)

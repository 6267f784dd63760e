"""Debug helper (GPU box): fused split-bf16 chains (gemm_impl=0) against the per-layer split-bf16 GEMMs (gemm_impl=2)."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import refnerf_oracle as O
from refnerf_pl_b200 import synthetic, train_utils
from tests._gpu import DEV, build_model, load_params, rays_obj

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
p = O.init_params(seed=5, bias_std=0.1, weight_scale=1.2)
rays = synthetic.blender_rays(n, seed=10)
gt = torch.tensor(synthetic.gt_rgb(n, 10), device=DEV)
res = {}
for impl in (0, 2):
    model, cfg = build_model('bf16x3', mlp_kwargs=dict(gemm_impl=impl, chunk_rows=chunk))
    load_params(model, p)
    r = rays_obj(rays)
    model.eval()
    with torch.no_grad():
        rend_e, hist_e = model(r, 1.0, True)
    torch.cuda.synchronize()
    print('impl', impl, 'eval ok', flush=True)
    model.train(True)
    rend, hist = model(r, 1.0, True)
    torch.cuda.synchronize()
    print('impl', impl, 'train fwd ok', flush=True)
    loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
    loss.backward()
    torch.cuda.synchronize()
    print('impl', impl, 'bwd ok, loss', float(loss), flush=True)
    res[impl] = dict(he=hist_e, h=hist, re=rend_e, r=rend, g={k: v.grad.clone() for k, v in model.nerf_mlp.named_parameters()})
a, b = res[0], res[2]
for lvl in range(2):
    for k in ('density', 'rgb', 'roughness', 'normals_pred', 'tint', 'diffuse'):
        d = (a['he'][lvl][k] - b['he'][lvl][k]).abs().max()
        d2 = (a['h'][lvl][k] - b['h'][lvl][k]).abs().max()
        print(f'lvl {lvl} {k:14s} eval max|d| {float(d):.3e}   train max|d| {float(d2):.3e}   (max|ref| {float(b["h"][lvl][k].abs().max()):.3e})')
    d = (a['h'][lvl]['normals'] - b['h'][lvl]['normals']).abs()
    print(f'lvl {lvl} normals        mean|d| {float(d.mean()):.3e} max {float(d.max()):.3e}')
    print(f'lvl {lvl} rend rgb       max|d| {float((a["r"][lvl]["rgb"] - b["r"][lvl]["rgb"]).abs().max()):.3e}')
for k in a['g']:
    x, y = a['g'][k].double(), b['g'][k].double()
    print(f'{k:28s} rel {float((x - y).norm() / y.norm()):.3e}   |b| {float(y.norm()):.3e}')

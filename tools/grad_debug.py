"""Debug helper (GPU box): parameter gradients of the fused-chain path vs the per-layer tcgen05 path."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import refnerf_oracle as O
from refnerf_pl_b200 import synthetic, train_utils
from tests._gpu import DEV, build_model, load_params, rays_obj

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
p = O.init_params(seed=5, bias_std=0.1, weight_scale=1.2)
rays = synthetic.blender_rays(n, seed=10)
gt = torch.tensor(synthetic.gt_rgb(n, 10), device=DEV)
grads = {}
IMPL = int(__import__('os').environ.get('IMPL', '0'))
for impl in (IMPL, 2):
    model, cfg = build_model('bf16', mlp_kwargs=dict(gemm_impl=impl, chunk_rows=chunk))
    load_params(model, p)
    model.train(True)
    r = rays_obj(rays)
    rend, hist = model(r, 1.0, True)
    loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
    loss.backward()
    grads[impl] = {k: v.grad.clone() for k, v in model.nerf_mlp.named_parameters()}
for k in grads[IMPL]:
    a, b = grads[IMPL][k].double(), grads[2][k].double()
    print(f'{k:28s} rel {float((a - b).norm() / b.norm()):.3e}   |b| {float(b.norm()):.3e}')

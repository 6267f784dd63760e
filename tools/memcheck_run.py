"""Small forward + backward of every fused path (split-bf16, fp16 and bf16 chains, deterministic wgrad, multi-chunk with a
ragged tail, eval with extras, fused normal losses, dilation + annealing, the geometry-config step with its noisy rays),
meant to run under `compute-sanitizer --tool memcheck` on the GPU box:
    compute-sanitizer --tool memcheck python tools/memcheck_run.py [rays] [chunk_rows]"""
import sys

import torch

sys.path.insert(0, '.')
from oracle import refnerf_oracle as O  # noqa: E402
from refnerf_pl_b200 import synthetic, train_utils  # noqa: E402
from tests._gpu import DEV, build_model, load_params, rays_obj  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
p = O.init_params(seed=5, bias_std=0.1, weight_scale=1.2)
rays = synthetic.blender_rays(n, seed=10)
gt = torch.tensor(synthetic.gt_rgb(n, 10), device=DEV)
for prec, extra in (('bf16x3', {}), ('bf16x3', dict(deterministic_wgrad=True)), ('fp16', {}), ('bf16', {})):
    model, cfg = build_model(prec, mlp_kwargs=dict(chunk_rows=chunk, **extra))
    load_params(model, p)
    model.train(True)
    r = rays_obj(rays)
    rend, hist = model(r, 1.0, True)
    loss, _ = train_utils.total_loss(model, r.viewdirs, r.lossmult, gt, rend, hist, cfg)
    loss.backward()
    g = torch.cat([q.grad.reshape(-1) for q in model.nerf_mlp.parameters()])
    model.eval()
    with torch.no_grad():
        rend, _ = model(r, 1.0, True)
    torch.cuda.synchronize()
    print(f'{prec}: loss {float(loss.detach()):.6f}  |grad| {float(g.norm()):.4e}  finite {bool(torch.isfinite(g).all())}  '
          f'rgb mean {float(rend[1]["rgb"].mean()):.4f}')

# dilation + annealing (max_dilate kernel, 382-bin resampler) and the geometry-config step (noisy rays, second Model call)
model, cfg = build_model('bf16x3', model_kwargs=dict(dilation_multiplier=0.5, dilation_bias=0.0025, anneal_slope=10.0))
load_params(model, p)
model.eval()
with torch.no_grad():
    rend, _ = model(rays_obj(rays), 0.4, True)
print('dilate+anneal: rgb mean', float(rend[1]['rgb'].mean()))
model, cfg = build_model('bf16x3', gin='llff_refnerf_geometry_losses.gin')
load_params(model, p)
model.train(True)
lr = synthetic.llff_rays(max(n, 160), seed=3)
loss = train_utils.training_losses(model, rays_obj(lr), torch.tensor(synthetic.gt_rgb(max(n, 160), 3), device=DEV), cfg, 1.0, 100000)[0]
loss.backward()
torch.cuda.synchronize()
print('geometry step: loss', float(loss.detach()))

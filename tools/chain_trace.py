"""GPU box: print the chain kernels' per-op timelines (RN_CHAIN_TRACE launches of each kernel family) for one step.
   PREC=bf16x3 python tools/chain_trace.py [train|eval]"""
import os, sys
os.environ.setdefault('RN_CHAIN_TRACE', '3')
import torch
sys.path.insert(0, '.')
from bench import build_everything
from refnerf_pl_b200 import synthetic, utils, train_utils
model, cfg = build_everything(os.environ.get('PREC', 'bf16x3'), 'cuda')
train = len(sys.argv) > 1 and sys.argv[1] == 'train'
model.train(train)
n = int(os.environ.get('RAYS', 16384))
r = synthetic.blender_rays(n, seed=3)
rays = utils.Rays(**{k: torch.from_numpy(v).cuda() for k, v in r.items()})
gt = torch.tensor(synthetic.gt_rgb(n, 3), device='cuda')
if train:
    rend, hist = model(rays, 1.0, False)
    loss, _ = train_utils.total_loss(model, rays.viewdirs, rays.lossmult, gt, rend, hist, cfg)
    loss.backward()
else:
    with torch.no_grad():
        model(rays, 1.0, False)
torch.cuda.synchronize()

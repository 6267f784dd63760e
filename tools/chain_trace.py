"""GPU box: print the chain kernel's per-tile timeline (RN_CHAIN_TRACE launches) for an eval forward."""
import os, sys
os.environ.setdefault('RN_CHAIN_TRACE', '2')
import torch
sys.path.insert(0, '.')
from bench import build_everything
from refnerf_pl_b200 import synthetic, utils
model, cfg = build_everything(os.environ.get('PREC', 'fp16'), 'cuda')
train = len(sys.argv) > 1 and sys.argv[1] == 'train'
model.train(train)
r = synthetic.blender_rays(16384, seed=3)
rays = utils.Rays(**{k: torch.from_numpy(v).cuda() for k, v in r.items()})
with torch.no_grad():
    model(rays, 1.0, False)
torch.cuda.synchronize()

import sys, time, torch
sys.path.insert(0, '.')
import bench
from refnerf_pl_b200 import models, synthetic, utils
dev = torch.device('cuda', 0)
import os
model, cfg = bench.build_everything(os.environ.get('PREC', 'bf16x3'), dev)
model.eval()
frame = synthetic.blender_rays(None, seed=7)
fr = utils.Rays(**{k: torch.from_numpy(v).to(dev).reshape(640000, 1, -1) for k, v in frame.items()})
fn = lambda r: model(r, 1.0, True)
def t(chunk, graph, sampler=False):
    cfg.render_chunk_size = chunk
    with torch.no_grad():
        models.render_image(fn, fr, cfg, use_graph=graph)
        torch.cuda.synchronize()
        smp = bench.ClockSampler(0).start() if sampler else None
        t0 = time.perf_counter()
        for _ in range(2):
            models.render_image(fn, fr, cfg, use_graph=graph)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        clk = smp.stop() if smp else None
    print(f'chunk {chunk} graph {graph} sampler {sampler}: {dt*1e3:.1f} ms/frame', clk, flush=True)
for chunk in (4096, 16384, 65536):
    for graph in (True, False):
        t(chunk, graph)



"""GPU box: where the HOST time of one training step goes (cProfile over the enqueue of N steps, no sync inside).
   python tools/host_profile.py [steps]"""
import cProfile, pstats, sys, time, io
import torch
sys.path.insert(0, '.')
import bench
from refnerf_pl_b200 import synthetic
n = 16384
dev = torch.device('cuda', 0)
wl = bench.Workload('blender_refnerf.gin', 'bf16x3', synthetic.blender_rays(n, seed=100), synthetic.gt_rgb(n, seed=100), dev, 1)
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t0 = time.perf_counter()
for _ in range(steps):
    wl.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / steps:.2f} ms/step, wall {1e3 * (t2 - t0) / steps:.2f} ms/step')
# phases (host time, no sync)
def phase_times():
    tu, cfg, model = wl.train_utils, wl.cfg, wl.model
    ts = [time.perf_counter()]
    rend, hist = model(wl.resident, 1.0, True); ts.append(time.perf_counter())
    loss, _ = tu.total_loss(model, wl.resident.viewdirs, wl.resident.lossmult, wl.gt_res, rend, hist, cfg); ts.append(time.perf_counter())
    wl.opt.zero_grad(set_to_none=True); ts.append(time.perf_counter())
    loss.backward(); ts.append(time.perf_counter())
    torch.nn.utils.clip_grad_norm_(model.nerf_mlp.parameters(), cfg.grad_max_norm); ts.append(time.perf_counter())
    wl.opt.step(); wl.sched.step(); ts.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(ts, ts[1:])]
import numpy as np
ph = np.array([phase_times() for _ in range(8)])
torch.cuda.synchronize()
print('host ms per phase (forward, losses, zero_grad, backward, clip, adam+sched):', np.round(np.median(ph, 0), 2))
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    wl.step()
pr.disable()
torch.cuda.synchronize()
for key in ('cumulative', 'tottime'):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])

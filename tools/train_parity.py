"""2k-step training parity run (north_star: "a 2k-step Blender training run must land within 0.1 dB PSNR").

There is no dataset in this environment, so the scene is analytic: a Phong-shaded unit sphere with a procedural
albedo in front of a white background, seen from Blender-shaped cameras (synthetic.blender_rays).  The SAME ray
stream, ground truth and initial weights train the throughput modes (fp16 / bf16 tensor-core chains) and the parity mode
(bf16x3 split-bf16, ~fp32 arithmetic, the mode that meets the per-sample 1e-3 gates against the reference); the
held-out PSNR of the two must agree within 0.1 dB.   python tools/train_parity.py [steps] [rays_per_step]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from bench import build_everything  # noqa: E402
from refnerf_pl_b200 import synthetic, train_utils, utils  # noqa: E402

LIGHT = np.array([0.4, 0.3, 0.866])
LIGHT = LIGHT / np.linalg.norm(LIGHT)


def shade(rays):
    """Analytic ground truth for rays (numpy dict): unit sphere at the origin, Phong shading, white background."""
    o, d = rays['origins'].astype(np.float64), rays['directions'].astype(np.float64)
    a = (d * d).sum(-1)
    b = 2 * (o * d).sum(-1)
    c = (o * o).sum(-1) - 1.0
    disc = b * b - 4 * a * c
    hit = disc > 0
    t = (-b - np.sqrt(np.where(hit, disc, 0))) / (2 * a)
    p = o + t[:, None] * d
    n = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), 1e-9)
    albedo = 0.5 + 0.5 * np.stack([np.sin(5 * p[:, 0]), np.sin(5 * p[:, 1] + 1.0), np.sin(5 * p[:, 2] + 2.0)], -1)
    v = -d / np.linalg.norm(d, axis=-1, keepdims=True)
    diff = np.clip((n * LIGHT).sum(-1), 0, 1)[:, None]
    r = 2 * (n * LIGHT).sum(-1, keepdims=True) * n - LIGHT
    spec = np.clip((r * v).sum(-1), 0, 1)[:, None] ** 20
    col = np.clip(albedo * (0.15 + 0.85 * diff) + 0.6 * spec, 0, 1)
    return np.where(hit[:, None], col, 1.0).astype(np.float32)


def run(precision, steps, n_rays, dev):
    model, cfg = build_everything(precision, dev)
    model.train(True)
    opt, sched = train_utils.create_optimizer(cfg, list(model.nerf_mlp.parameters()))
    t0 = time.time()
    for step in range(steps):
        r = synthetic.blender_rays(n_rays, seed=1000 + step)
        gt = torch.from_numpy(shade(r)).to(dev)
        rays = utils.Rays(**{k: torch.from_numpy(v).to(dev) for k, v in r.items()})
        rend, hist = model(rays, min(1.0, step / steps), False)
        loss, _ = train_utils.total_loss(model, rays.viewdirs, rays.lossmult, gt, rend, hist, cfg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if cfg.grad_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(model.nerf_mlp.parameters(), cfg.grad_max_norm)
        opt.step()
        sched.step()
        if step % 500 == 0:
            print(f'  [{precision}] step {step} loss {float(loss):.5f}', flush=True)
    torch.cuda.synchronize()
    train_s = time.time() - t0
    model.eval()
    mse, cnt = 0.0, 0
    with torch.no_grad():
        for k in range(8):
            r = synthetic.blender_rays(4096, seed=900000 + k)
            gt = torch.from_numpy(shade(r)).to(dev)
            rays = utils.Rays(**{k2: torch.from_numpy(v).to(dev) for k2, v in r.items()})
            rend, _ = model(rays, 1.0, False)
            mse += float(((rend[-1]['rgb'] - gt) ** 2).sum())
            cnt += gt.numel()
    psnr = -10 * np.log10(mse / cnt)
    return psnr, train_s


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    modes = tuple(sys.argv[3].split(',')) if len(sys.argv) > 3 else ('fp16', 'bf16', 'bf16x3')
    repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device('cuda', 0)
    # identical initial weights and ray stream every run: repeats differ only through the order of the wgrad atomics,
    # i.e. they measure the run-to-run spread a precision mode has against ITSELF
    runs = []
    for rep in range(repeats):
        for prec in modes:
            torch.manual_seed(0)
            psnr, secs = run(prec, steps, n_rays, dev)
            runs.append({'precision': prec, 'repeat': rep, 'psnr_db': psnr, 'train_seconds': secs})
            print(f'{prec}[{rep}]: held-out PSNR {psnr:.3f} dB after {steps} steps of {n_rays} rays ({secs:.1f} s)', flush=True)
    res = {'runs': runs, 'steps': steps, 'rays_per_step': n_rays,
           'scene': 'analytic Phong sphere, white background, Blender-shaped cameras (tools/train_parity.py)'}
    for prec in modes:
        v = [r['psnr_db'] for r in runs if r['precision'] == prec]
        res[prec] = {'mean_psnr_db': float(np.mean(v)), 'min': float(np.min(v)), 'max': float(np.max(v)), 'n': len(v)}
    for prec in modes:
        if prec != 'bf16x3' and 'bf16x3' in res:
            res[f'delta_db_{prec}'] = res[prec]['mean_psnr_db'] - res['bf16x3']['mean_psnr_db']
    print(json.dumps(res))


if __name__ == '__main__':
    main()

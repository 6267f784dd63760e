"""2k-step training parity run (north_star: "a 2k-step Blender training run must land within 0.1 dB PSNR").

Arms: `ref_fp32` = the REFERENCE ARITHMETIC: the oracle port of the reference (unfused fp32 torch ops + autograd,
oracle/refnerf_oracle.py -- pinned against the unmodified reference) running on the same GPU, same optimiser recipe
(train_utils.py:448-467, math.py:46-78), same ray stream, same initial weights; `bf16x3` / `fp16` / `bf16` = this package.
Training is chaotic (1e-7 perturbations grow to visible PSNR differences within 2k steps), so the comparison is made per
ray-stream seed and reported as the mean paired difference with its spread.

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


def run_reference(steps, n_rays, dev, stream_seed=0, perturb=0):
    """The reference arithmetic (oracle port, fp32, autograd) on the GPU: same stream, weights, optimiser, schedule.
    `perturb` > 0 multiplies every initial weight by (1 + 1e-6 N(0,1)) -- a perturbation of the size of an fp32
    re-ordering: repeats of this arm then measure the chaotic spread the REFERENCE arithmetic has against itself."""
    from oracle import refnerf_oracle as O
    model, cfg = build_everything('fp32', dev)          # only for the initial weights and the optimiser recipe
    params = {k: v.detach().clone() for k, v in model.nerf_mlp.state_dict().items()}
    if perturb:
        g = torch.Generator(device='cpu').manual_seed(777 + perturb)
        params = {k: v * (1 + 1e-6 * torch.randn(v.shape, generator=g).to(v.device)) for k, v in params.items()}
    params = {k: v.requires_grad_(True) for k, v in params.items()}
    del model
    opt, sched = train_utils.create_optimizer(cfg, list(params.values()))
    loss_cfg = dict(data_loss_mult=cfg.data_loss_mult, data_coarse_loss_mult=cfg.data_coarse_loss_mult,
                    orientation_loss_mult=cfg.orientation_loss_mult, orientation_coarse_loss_mult=cfg.orientation_coarse_loss_mult,
                    predicted_normal_loss_mult=cfg.predicted_normal_loss_mult,
                    predicted_normal_coarse_loss_mult=cfg.predicted_normal_coarse_loss_mult,
                    interlevel_loss_mult=cfg.interlevel_loss_mult)
    assert cfg.data_loss_type == 'mse'
    t0 = time.time()
    for step in range(steps):
        r = synthetic.blender_rays(n_rays, seed=1000 + 100000 * stream_seed + step)
        gt = torch.from_numpy(shade(r)).to(dev)
        rays = {k: torch.from_numpy(v).to(dev) for k, v in r.items()}
        rend, hist = O.model_forward(params, rays, min(1.0, step / steps), False, True)
        loss = O.total_loss(rend, hist, rays, gt, loss_cfg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if cfg.grad_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(list(params.values()), cfg.grad_max_norm)
        opt.step()
        sched.step()
        if step % 500 == 0:
            print(f'  [ref_fp32] step {step} loss {float(loss.detach()):.5f}', flush=True)
    torch.cuda.synchronize()
    train_s = time.time() - t0
    mse, cnt = 0.0, 0
    with torch.no_grad():
        for k in range(8):
            r = synthetic.blender_rays(4096, seed=900000 + k)
            gt = torch.from_numpy(shade(r)).to(dev)
            rays = {k2: torch.from_numpy(v).to(dev) for k2, v in r.items()}
            rend, _ = O.model_forward(params, rays, 1.0, False, False)
            mse += float(((rend[-1]['rgb'] - gt) ** 2).sum())
            cnt += gt.numel()
    return -10 * np.log10(mse / cnt), train_s


def run(precision, steps, n_rays, dev, stream_seed=0, rep=0):
    if precision == 'ref_fp32':
        return run_reference(steps, n_rays, dev, stream_seed, perturb=rep)
    model, cfg = build_everything(precision, dev)
    model.train(True)
    opt, sched = train_utils.create_optimizer(cfg, list(model.nerf_mlp.parameters()))
    t0 = time.time()
    for step in range(steps):
        r = synthetic.blender_rays(n_rays, seed=1000 + 100000 * stream_seed + step)
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
    if precision == 'bf16x3':
        TRAINED['model'] = model
    return psnr, train_s


TRAINED = {}


def trained_scale_parity(dev):
    """Forward parity of the throughput modes at TRAINED weights: the model trained in the parity mode (bf16x3) is
    evaluated on held-out rays in every mode (NerfMLP.precision is read per call) and compared with its own bf16x3
    outputs -- per-sample density (relative, 99.9th percentile), per-sample rgb and composited rgb (absolute)."""
    model = TRAINED.get('model')
    if model is None:
        return None
    r = synthetic.blender_rays(4096, seed=910000)
    rays = utils.Rays(**{k: torch.from_numpy(v).to(dev) for k, v in r.items()})
    outs = {}
    for train_mode in (False, True):
        model.train(train_mode)
        for prec in ('bf16x3', 'fp16', 'bf16'):
            model.nerf_mlp.precision = prec
            with torch.no_grad():
                outs[(prec, train_mode)] = model(rays, 1.0, True)
    model.nerf_mlp.precision = 'bf16x3'
    res = {}
    q = lambda x, p: float(torch.quantile(x.flatten()[::3].float(), p))
    for prec in ('fp16', 'bf16'):
        (rend, hist), (rend0, hist0) = outs[(prec, False)], outs[('bf16x3', False)]
        # level 0 samples sit on input-independent fenceposts, so per-sample values are comparable one to one; level 1
        # samples are re-drawn from the level-0 weights, so there only the composited values are compared
        d, d0 = hist[0]['density'].double(), hist0[0]['density'].double()
        rel = (d - d0).abs() / (d0.abs() + 1e-3 * d0.abs().max())
        nrm, nrm0 = outs[(prec, True)][1][0]['normals'], outs[('bf16x3', True)][1][0]['normals']
        e = {'level0_density_rel_p999': q(rel, 0.999), 'level0_density_rel_median': q(rel, 0.5),
             'level0_density_max': float(d0.max()),
             'level0_sample_rgb_abs_p999': q((hist[0]['rgb'] - hist0[0]['rgb']).abs(), 0.999),
             'level0_density_normals_mean_abs': float((nrm - nrm0).abs().mean())}
        for lvl in (0, 1):
            c = (rend[lvl]['rgb'] - rend0[lvl]['rgb']).abs().max(dim=-1).values
            e[f'composited_rgb{lvl}_abs'] = {'mean': float(c.mean()), 'p99': q(c, 0.99), 'max': float(c.max())}
            a_ = (rend[lvl]['acc'] - rend0[lvl]['acc']).abs()
            e[f'composited_acc{lvl}_abs'] = {'mean': float(a_.mean()), 'p99': q(a_, 0.99), 'max': float(a_.max())}
        res[prec] = e
    import os
    os.makedirs('gpurun_out', exist_ok=True)
    np.savez_compressed('gpurun_out/trained_bf16x3_params.npz',
                        **{k: v.detach().cpu().numpy() for k, v in model.nerf_mlp.state_dict().items()})
    return res


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    modes = tuple(sys.argv[3].split(',')) if len(sys.argv) > 3 else ('fp16', 'bf16', 'bf16x3')
    repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    seeds = int(sys.argv[5]) if len(sys.argv) > 5 else 1       # ray-stream seeds (paired comparison per seed)
    # optional "mode:first_rep" entries: start a mode at a later repeat (to add perturbed reference runs to an earlier log)
    first_rep = {}
    modes = tuple(m.split(':')[0] for m in modes if not first_rep.update({m.split(':')[0]: int(m.split(':')[1])} if ':' in m else {}))
    dev = torch.device('cuda', 0)
    # identical initial weights and ray stream every run: repeats differ only through the order of the wgrad atomics,
    # i.e. they measure the run-to-run spread a precision mode has against ITSELF
    runs = []
    for seed in range(seeds):
        for rep in range(repeats):
            for prec in modes:
                if rep < first_rep.get(prec, 0):
                    continue
                torch.manual_seed(0)
                psnr, secs = run(prec, steps, n_rays, dev, seed, rep)
                runs.append({'precision': prec, 'repeat': rep, 'stream_seed': seed, 'psnr_db': psnr, 'train_seconds': secs})
                print(f'{prec}[seed {seed}, rep {rep}]: held-out PSNR {psnr:.3f} dB after {steps} steps of {n_rays} rays ({secs:.1f} s)',
                      flush=True)
    res = {'runs': runs, 'steps': steps, 'rays_per_step': n_rays,
           'scene': 'analytic Phong sphere, white background, Blender-shaped cameras (tools/train_parity.py)'}
    for prec in modes:
        v = [r['psnr_db'] for r in runs if r['precision'] == prec]
        res[prec] = {'mean_psnr_db': float(np.mean(v)), 'min': float(np.min(v)), 'max': float(np.max(v)), 'n': len(v)}
    for prec in modes:
        if prec != 'bf16x3' and 'bf16x3' in res:
            res[f'delta_db_{prec}'] = res[prec]['mean_psnr_db'] - res['bf16x3']['mean_psnr_db']
    if 'ref_fp32' in modes:   # paired per ray-stream seed against the reference arithmetic
        for prec in modes:
            if prec == 'ref_fp32':
                continue
            diffs = []
            for seed in range(seeds):
                ref = [r['psnr_db'] for r in runs if r['precision'] == 'ref_fp32' and r['stream_seed'] == seed]
                mine = [r['psnr_db'] for r in runs if r['precision'] == prec and r['stream_seed'] == seed]
                if ref and mine:
                    diffs.append(float(np.mean(mine) - np.mean(ref)))
            if diffs:
                res[f'vs_reference_arithmetic_{prec}'] = {
                    'paired_delta_db_per_seed': diffs, 'mean_delta_db': float(np.mean(diffs)),
                    'stderr_db': float(np.std(diffs, ddof=1) / np.sqrt(len(diffs))) if len(diffs) > 1 else None}
    res['trained_scale_forward_parity_vs_bf16x3'] = trained_scale_parity(dev)
    print(json.dumps(res))


if __name__ == '__main__':
    main()

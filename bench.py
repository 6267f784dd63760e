#!/usr/bin/env python
"""bench.py -- Ref-NeRF hot-path throughput on B200 (BASELINE.json: "train rays/s (fwd+bwd) & 800x800
render ms/frame at 1/2/4/8 B200 vs host CPU").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU oracle port on the host cores

One "step" = one full training step of configs/blender_refnerf.gin on a batch of synthetic
Blender-shaped rays with reference-init weights: Model forward in training mode (both levels, incl. the
in-forward density-gradient normals pass, compute_extras=True as nerf_system.py:89-95 does), data +
orientation + predicted-normal losses, backward, [N>1: one NCCL all-reduce of the 4.44 MB gradient],
grad clipping and the Adam update.  Weak scaling: every rank runs `--rays` rays (default 16384).
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'train rays/s (fwd+bwd)'
UNIT = 'rays/s'
# algorithmic MLP FLOPs per sample (dense, unpadded) for fwd + in-forward normals dgrad + bwd
FLOP_PER_SAMPLE_TRAIN = 7_651_840
FLOP_PER_SAMPLE_EVAL = 2_211_840
SAMPLES_PER_RAY = 256  # 2 levels x 128


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_burst=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'],
                    source='measured')
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def oracle_step_fn(n_rays, seed=0, mode='train'):
    """One step of the CPU oracle port.  mode: 'train' = forward in training mode + losses + backward;
    'train_fwd' = training-mode forward only (incl. the density-gradient pass); 'eval_fwd' = no_grad, extras on."""
    from oracle import refnerf_oracle as O
    from refnerf_pl_b200 import synthetic
    rays = {k: torch.tensor(v) for k, v in synthetic.blender_rays(n_rays, seed=seed).items()}
    gt = torch.tensor(synthetic.gt_rgb(n_rays, seed))
    p = {k: v.requires_grad_(True) for k, v in O.init_params(seed=0).items()}

    def step():
        for v in p.values():
            v.grad = None
        if mode == 'eval_fwd':
            with torch.no_grad():
                O.model_forward(p, rays, 1.0, True, False)
            return 0.0
        rend, hist = O.model_forward(p, rays, 1.0, True, True)
        if mode == 'train_fwd':
            return 0.0
        loss = O.total_loss(rend, hist, rays, gt)
        loss.backward()
        return float(loss.detach())

    return step


def reference_step_fn(n_rays, seed=0, mode='train'):
    """The same step through the UNMODIFIED reference imported from /root/reference; only where
    that tree exists (the build container -- it is absent on the GPU box)."""
    from oracle import ref_import
    from oracle import refnerf_oracle as O
    from refnerf_pl_b200 import synthetic
    ns, config = ref_import.load('blender_refnerf.gin')
    torch.manual_seed(0)
    model = ns.models.construct_model(ns.utils.dummy_rays(), config)
    sd = model.nerf_mlp.state_dict()
    p = O.init_params(seed=0)
    model.nerf_mlp.load_state_dict({k: p[k].clone() for k in sd})
    rays = ns.utils.Rays(**{k: torch.tensor(v) for k, v in synthetic.blender_rays(n_rays, seed=seed).items()})
    gt = synthetic.gt_rgb(n_rays, seed)

    class B:
        rgb = gt

    def step():
        model.zero_grad()
        if mode == 'eval_fwd':
            model.train(False)
            with torch.no_grad():
                model(rays, 1.0, True)
            return 0.0
        model.train(True)
        rend, hist = model(rays, 1.0, True)
        if mode == 'train_fwd':
            return 0.0
        loss = (ns.train_utils.compute_data_loss(B, rend, rays, config)[0] + ns.train_utils.orientation_loss(rays, model, hist, config)
                + ns.train_utils.predicted_normal_loss(model, hist, config))
        loss.backward()
        return float(loss.detach())

    return step


def cpu_step(n_rays, mode='train'):
    """-> (step function, kind): the imported reference where /root/reference exists, else the oracle port."""
    try:
        from oracle import ref_import
        if ref_import.available():
            return reference_step_fn(n_rays, mode=mode), 'reference'
    except Exception:   # noqa: BLE001 -- fall back to the port, never take the bench down
        pass
    return oracle_step_fn(n_rays, mode=mode), 'port'


def time_cpu(n_rays, steps, warmup, mode='train'):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = cpu_step(n_rays, mode)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), kind


def time_torch_gpu(n_rays, dev, steps=2):
    """The stock-PyTorch fp32 path (the oracle port, i.e. the reference's unfused ATen op sequence) on THIS GPU: the
    "what you get without this work" line of SURVEY 8(d).  Baseline only; returns None if it cannot run."""
    try:
        from oracle import refnerf_oracle as O
        from refnerf_pl_b200 import synthetic
        rays = {k: torch.tensor(v).to(dev) for k, v in synthetic.blender_rays(n_rays, seed=0).items()}
        gt = torch.tensor(synthetic.gt_rgb(n_rays, 0)).to(dev)
        p = {k: v.to(dev).requires_grad_(True) for k, v in O.init_params(seed=0).items()}

        def step():
            for v in p.values():
                v.grad = None
            rend, hist = O.model_forward(p, rays, 1.0, True, True)
            O.total_loss(rend, hist, rays, gt).backward()

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {'value': n_rays / (ms * 1e-3), 'unit': UNIT, 'rays': n_rays, 'ms_per_step': ms, 'dtype': 'f32',
                'what': 'oracle port (unfused torch ops, fp32, autograd) on the same B200: fwd incl. normals pass + losses + bwd, '
                        'no optimizer'}
    except Exception as e:  # noqa: BLE001 -- a baseline must never take the bench down
        return {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    finally:
        torch.cuda.empty_cache()


def cpu_model_name():
    try:
        for ln in open('/proc/cpuinfo'):
            if ln.startswith('model name'):
                return ln.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores (the imported
    reference where /root/reference exists, else the oracle port), every step a bounded sample of the 16384-ray batch
    sized so that the whole run ends within a few minutes."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.cpu_rays
    if n <= 0:
        # probe the host: one 256-ray step, then the largest power-of-two sample <= 4096 that
        # keeps warm-up + steps inside ~150 s
        probe, _ = time_cpu(256, 1, 1)
        budget = 150.0 / max(1, args.steps + min(args.warmup, 1))
        n = 256
        while n < 4096 and (2 * n) * probe / 256 <= budget:
            n *= 2
    sec, kind = time_cpu(n, max(1, args.steps), max(0, min(args.warmup, 1)))
    val = n / sec
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'configs/blender_refnerf.gin single training step (fwd incl. density-gradient normals + losses + '
                               f'bwd), {args.rays}-ray batch; timed on a {n}-ray sample of it (CPU rays/s does not depend on '
                               'the batch size beyond a few hundred rays)',
                   'rays_per_step': n, 'levels': 2, 'samples_per_level': 128,
                   'code': 'unmodified reference imported from /root/reference' if kind == 'reference'
                           else 'oracle port of the reference (the reference tree is absent on this box)'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': f'{n} rays of the {args.rays}-ray batch, {args.steps} step(s), median; '
                                   f'torch threads={cores}; {cpu_model_name()}'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def build_everything(precision, device, gin='blender_refnerf.gin'):
    from refnerf_pl_b200 import configs, models
    configs.clear_bindings()
    configs.parse_gin_files_and_bindings([os.path.join(ROOT, 'configs', gin)])
    configs.bind('NerfMLP', precision=precision)
    if os.environ.get('REFNERF_B200_GEMM_IMPL'):
        configs.bind('NerfMLP', gemm_impl=int(os.environ['REFNERF_B200_GEMM_IMPL']))
    cfg = configs.Config()
    torch.manual_seed(0)
    model = models.Model(config=cfg).to(device)
    return model, cfg


def load_profile_json(name):
    """ncu-derived figures committed under profiles/ (per-launch DRAM bytes, step totals), or {}."""
    p = os.path.join(ROOT, 'profiles', name)
    return json.load(open(p)) if os.path.exists(p) else {}


def time_hbm_kernels(dev, peaks, iters=20):
    """Achieved HBM GB/s of the warp-per-ray kernels (SURVEY 8(d) algorithmic bytes), inputs larger than L2."""
    from refnerf_pl_b200 import ops
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    def entry(bytes_per_ray, n, sec):
        gbs = bytes_per_ray * n / sec / 1e9
        return {'gbs': gbs, 'frac': gbs / peaks['hbm_gbs'], 'bytes_per_ray': bytes_per_ray, 'rays': n, 'us': sec * 1e6}

    n, s = 131072, 128       # 1.5 GB of inputs per launch (>> 126 MB L2; long enough to amortise the launch path)
    t = torch.sort(rnd(n, s + 1) * 4 + 2, dim=-1).values
    dirs = rnd(n, 3) + 0.5
    far = torch.full((n, 1), 6.0, device=dev)
    dens, rough = rnd(n, s) * 3, rnd(n, s, 1)
    c3 = [rnd(n, s, 3) for _ in range(6)]
    sec = timeit(lambda: ops.composite_fwd(dens, t, dirs, far, c3[0], c3[1], c3[2], c3[3], c3[4], rough, c3[5], 1.0, True))
    out['composite_fwd_extras'] = entry(11384, n, sec)
    sec = timeit(lambda: ops.composite_fwd(dens, t, dirs, far, c3[0], c3[1], c3[2], c3[3], c3[4], rough, c3[5], 1.0, False))
    out['composite_fwd'] = entry(6208, n, sec)
    w_, comp_, _, _ = ops.composite_fwd(dens, t, dirs, far, c3[0], c3[1], c3[2], c3[3], c3[4], rough, c3[5], 1.0, False)
    gwt, gcomp, empty = rnd(n, s), rnd(n, 16), torch.empty(0, device=dev)
    sec = timeit(lambda: ops.composite_bwd(dens, t, dirs, c3[0], c3[1], c3[2], c3[3], c3[4], rough, c3[5], w_, comp_, gwt, gcomp,
                                           empty, 1.0, False))
    # reads density, tdist, weights, g_weights, 3 colour arrays; writes d_density + 3 colour gradients
    out['composite_bwd'] = entry(512 * 3 + 516 + 128 + 3 * 1536 + 512 + 3 * 1536, n, sec)
    te = torch.sort(rnd(n, s + 1), dim=-1).values
    we = rnd(n, s)
    sec = timeit(lambda: ops.lossfun_outer(te, dens, te, we))
    out['lossfun_outer'] = entry(2 * 516 + 2 * 512 + 512, n, sec)
    sec = timeit(lambda: ops.distortion(te, we))
    out['distortion'] = entry(516 + 512 + 4, n, sec)
    sd = torch.sort(rnd(n, s + 1), dim=-1).values
    w = rnd(n, s)
    near, far2 = torch.full((n, 1), 2.0, device=dev), torch.full((n, 1), 6.0, device=dev)
    sec = timeit(lambda: ops.resample(sd, w, near, far2, s, 0.01, 1.0, 0.0, 1.0, False))
    out['resample'] = entry(2060, n, sec)
    sec = timeit(lambda: ops.max_dilate_weights(sd, w, 0.0064, 0.0, 1.0, True, True))
    out['max_dilate_weights'] = entry(516 + 512 + (3 * s - 1) * 4 + (3 * s - 2) * 4, n, sec)
    return out


class Workload:
    """One training configuration of this package on this rank's GPU: model + optimiser + resident and pinned rays."""

    def __init__(self, gin, precision, rays_np, gt_np, dev, world, geometry=False):
        from refnerf_pl_b200 import parallel, train_utils, utils
        self.utils, self.train_utils = utils, train_utils
        self.model, self.cfg = build_everything(precision, dev, gin)
        self.model.train(True)
        self.opt, self.sched = train_utils.create_optimizer(self.cfg, [p for p in self.model.nerf_mlp.parameters()])
        self.reducer = parallel.GradAllReducer(self.model.nerf_mlp.parameters()) if world > 1 else None
        self.dev, self.geometry, self.n = dev, geometry, next(iter(rays_np.values())).shape[0]
        self.pinned = {k: torch.from_numpy(v).pin_memory() for k, v in rays_np.items()}
        self.gt_pinned = torch.from_numpy(gt_np).pin_memory()
        self.resident = utils.Rays(**{k: v.to(dev) for k, v in self.pinned.items()})
        self.gt_res = self.gt_pinned.to(dev)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.pinned.values()) + self.gt_pinned.numel() * 4
        self.global_step = 100000   # geometry config: consistency warm-up ratio 100000 / (0.6 * 250000) = 2/3

    def step(self, rays=None, gt=None):
        rays = self.resident if rays is None else rays
        gt = self.gt_res if gt is None else gt
        tu, cfg, model = self.train_utils, self.cfg, self.model
        if self.geometry:   # nerf_system.py:84-191 incl. the second Model call on the noisy rays
            loss = tu.training_losses(model, rays, gt, cfg, 1.0, self.global_step)[0]
        else:
            rend, hist = model(rays, 1.0, True)
            loss, _ = tu.total_loss(model, rays.viewdirs, rays.lossmult, gt, rend, hist, cfg)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()               # every .grad is a view of ONE flat buffer (ops.param_carrier): the sum over the
        if self.reducer is not None:  # step's fused MLP calls, so the collective runs on it directly (no copies, no div_)
            self.reducer.allreduce()
        if cfg.grad_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(model.nerf_mlp.parameters(), cfg.grad_max_norm)
        self.opt.step()
        self.sched.step()
        return loss

    def e2e_step(self):
        rays = self.utils.Rays(**{k: v.to(self.dev, non_blocking=True) for k, v in self.pinned.items()})
        gt = self.gt_pinned.to(self.dev, non_blocking=True)
        return float(self.step(rays, gt).detach())   # .item(): D2H of the loss

    def free(self):
        del self.model, self.opt, self.sched, self.resident, self.gt_res
        torch.cuda.empty_cache()


def run_b200(args):
    from refnerf_pl_b200 import _lib, parallel, synthetic, utils
    rank, world, local_rank = parallel.init_distributed()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    peaks = load_peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """ms for `steps` calls: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        timed.host_ms = (time.perf_counter() - h0) * 1e3 / max(1, steps)   # host time to ENQUEUE one step (no sync inside fn)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms

    def prof_read():
        prof = {}
        for c, name in enumerate(('gemm_tc', 'wgrad_tc', 'gemm_simt', 'chain_tc', 'encode', 'heads_prologue_fwd', 'heads_prologue_bwd',
                                  'ipe_grad_normals', 'color', 'glue', 'ray_kernels')):
            nl, tms, fl, ef = ctypes.c_int64(0), ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
            lib.rn_prof_summary2(c, ctypes.byref(nl), ctypes.byref(tms), ctypes.byref(fl), ctypes.byref(ef))
            prof[name] = dict(launches=nl.value, ms=tms.value, flops=fl.value, exec_flops=ef.value)
        return prof

    # ================= config 2 (the metric's configuration): blender_refnerf.gin, 16384 rays per GPU =================
    n = args.rays
    wl = Workload('blender_refnerf.gin', args.precision, synthetic.blender_rays(n, seed=100 + rank),
                  synthetic.gt_rgb(n, seed=100 + rank), dev, world)
    for _ in range(args.warmup):
        wl.step()
    # ---- device-resident timed region (value): nothing but the step's own kernels inside ----
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = lib.rn_launch_count()
    ms_total = timed(wl.step, args.steps)
    launches = lib.rn_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    # host time to ENQUEUE one step, measured on an empty launch queue (over K steps the host runs ~5 steps ahead and then
    # blocks in cudaLaunchKernel on the full queue, which would be counted as host time)
    barrier()
    h0 = time.perf_counter()
    wl.step()
    wl.step()
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / 2
    barrier()
    # ---- the same K steps once more with per-kernel-class CUDA events around every GEMM-class launch (roofline) ----
    lib.rn_prof_enable(1)
    prof_read()
    ms_prof_total = timed(wl.step, args.steps)
    prof = prof_read()
    lib.rn_prof_enable(0)
    allreduce_path = wl.reducer.last_path if wl.reducer is not None else None
    # ---- end-to-end: pinned host rays -> H2D every step, loss read back every step ----
    wl.e2e_step()
    ms_e2e = timed(wl.e2e_step, args.steps) / args.steps
    e2e_value = world * n / (ms_e2e * 1e-3)
    e2e_step_ms = []          # diagnostic: host wall clock of individual e2e steps (each ends with the loss read-back)
    for _ in range(min(args.steps, 8)):
        h0 = time.perf_counter()
        wl.e2e_step()
        e2e_step_ms.append(round((time.perf_counter() - h0) * 1e3, 2))

    # ---- training strong scaling: the 16384-ray batch split over the ranks (beside the weak-scaling value) ----
    strong = None
    if world > 1:
        lo, hi = parallel.shard_range(n, rank, world)
        sub = utils.Rays(**{k: v[lo:hi].contiguous() for k, v in dataclass_items(wl.resident)})
        sub_gt = wl.gt_res[lo:hi].contiguous()
        for _ in range(2):
            wl.step(sub, sub_gt)
        ms_s = timed(lambda: wl.step(sub, sub_gt), args.steps) / args.steps
        strong = {'rays_total': n, 'rays_per_gpu': hi - lo, 'ms_per_step': ms_s, 'value': n / (ms_s * 1e-3), 'unit': UNIT,
                  'scaling': 'strong'}

    torch_gpu = time_torch_gpu(2048, dev) if (rank == 0 and world == 1 and not args.no_cpu) else None

    # ---- 800x800 frame render (config 3: eval path, chunked): every rank renders a contiguous slice of the frame ----
    render = None
    if not args.no_render:
        from refnerf_pl_b200 import models
        wl.model.eval()
        frame = synthetic.blender_rays(None, seed=7)
        lo, hi = parallel.shard_range(640000, rank, world)
        fr = utils.Rays(**{k: torch.from_numpy(v[lo:hi]).to(dev).reshape(hi - lo, 1, -1) for k, v in frame.items()})
        fn = lambda r: wl.model(r, 1.0, True)
        render = {'frame': '800x800', 'compute_extras': True, 'n_gpus': world, 'precision': args.precision,
                  'sharding': 'contiguous ray slices per rank, outputs all-gathered' if world > 1 else 'single GPU',
                  'by_chunk': {}}
        for chunk in sorted({4096, args.render_chunk}):
            wl.cfg.render_chunk_size = chunk

            def render_once():
                out = models.render_image(fn, fr, wl.cfg)
                if world > 1:   # the frame is assembled on every rank (rgb + distance + acc, 5 floats per ray)
                    parallel.gather_rows(torch.cat([out['rgb'].reshape(-1, 3), out['distance'].reshape(-1, 1),
                                                    out['acc'].reshape(-1, 1)], dim=-1))
                return out

            with torch.no_grad():
                render_once()
                fms = timed(render_once, 2) / 2
            render['by_chunk'][str(chunk)] = {
                'ms_per_frame': fms, 'rays_per_s': 640000 / (fms * 1e-3),
                'mlp_tflops': FLOP_PER_SAMPLE_EVAL * SAMPLES_PER_RAY * 640000 / (fms * 1e-3) / 1e12,
                'driver': 'chunk loop without host synchronisation, outputs written straight into the frame buffers'}
        best = min(render['by_chunk'].values(), key=lambda d: d['ms_per_frame'])
        render.update(ms_per_frame=best['ms_per_frame'], rays_per_s=best['rays_per_s'], mlp_tflops=best['mlp_tflops'],
                      ms_per_frame_reference_chunk_4096=render['by_chunk']['4096']['ms_per_frame'])
        wl.model.train(True)
    wl.free()

    def short_run(gin, precision, rays_np, gt_np, geometry=False, steps=8, warm=3):
        w = Workload(gin, precision, rays_np, gt_np, dev, world, geometry)
        for _ in range(warm):
            w.step()
        smp = ClockSampler(local_rank).start() if rank == 0 else None
        ms = timed(w.step, steps) / steps
        clk = smp.stop() if smp else None
        w.e2e_step()
        ms_e = timed(w.e2e_step, steps) / steps
        nn = w.n
        w.free()
        return {'ms_per_step': ms, 'value': world * nn / (ms * 1e-3), 'unit': UNIT, 'rays_per_gpu': nn, 'precision': precision,
                'e2e_value': world * nn / (ms_e * 1e-3), 'steps': steps,
                'sm_mhz': clk['sm_mhz'] if clk else None}   # (the part is power-capped: later legs of a long run see lower clocks)

    def render_ms(precision, chunk=4096):
        """ms per 800x800 eval frame (compute_extras) of this rank's ray slice in another arithmetic mode"""
        from refnerf_pl_b200 import models
        m, c = build_everything(precision, dev)
        m.eval()
        frame = synthetic.blender_rays(None, seed=7)
        lo, hi = parallel.shard_range(640000, rank, world)
        fr = utils.Rays(**{k: torch.from_numpy(v[lo:hi]).to(dev).reshape(hi - lo, 1, -1) for k, v in frame.items()})
        c.render_chunk_size = chunk
        with torch.no_grad():
            models.render_image(lambda r: m(r, 1.0, True), fr, c)
            ms = timed(lambda: models.render_image(lambda r: m(r, 1.0, True), fr, c), 2) / 2
        del m, fr
        torch.cuda.empty_cache()
        return ms

    extra = {}
    if not args.no_extra:
        # ---- the fp16 throughput mode (a stated-error mode), same step ----
        if args.precision != 'fp16':
            extra['throughput_mode_fp16'] = dict(short_run('blender_refnerf.gin', 'fp16', synthetic.blender_rays(n, seed=100 + rank),
                                                           synthetic.gt_rgb(n, seed=100 + rank)),
                                                 note='fp16 operands (11-bit): misses the 1e-2 gradient and the trained-scale per-sample '
                                                      'tolerances of north_star; reported beside the headline, never as it')
            if not args.no_render:
                extra['throughput_mode_fp16']['render_800x800_ms_per_frame_chunk_4096'] = render_ms('fp16')
        # ---- configs 4 / 5: LLFF-shaped NDC rays (1008x756, near 0, far 1), ray-sharded, one all-reduce per step ----
        llff = synthetic.llff_rays(n, seed=200 + rank)
        extra['llff_refnerf'] = dict(short_run('llff_refnerf.gin', args.precision, llff, synthetic.gt_rgb(n, seed=200 + rank)),
                                     workload='configs/llff_refnerf.gin training step, forward-facing NDC rays')
        extra['llff_refnerf_geometry_losses'] = dict(
            short_run('llff_refnerf_geometry_losses.gin', args.precision, llff, synthetic.gt_rgb(n, seed=200 + rank), geometry=True),
            workload='configs/llff_refnerf_geometry_losses.gin training step: main forward + second Model call on 512 noisy rays '
                     '(128 rays x 4 rotations, compute_extras) + data / orientation / predicted-normal / consistency / '
                     'distance-consistency / acc / weights-entropy losses')

    if rank != 0:
        return
    # ---- roofline of the dominant kernel class (fused tcgen05 chains; SIMT GEMM in fp32 mode) ----
    dom = max(('gemm_tc', 'chain_tc', 'gemm_simt'), key=lambda k: prof[k]['ms'])
    d = prof[dom]
    sec = d['ms'] * 1e-3
    achieved = d['flops'] / sec / 1e12 if sec > 0 else 0.0
    executed = d['exec_flops'] / sec / 1e12 if sec > 0 else 0.0
    traffic = load_profile_json('r02b_traffic.json') or load_profile_json('r02_traffic.json')
    roofline = {'bound': 'tensor', 'kernel': dom, 'achieved': achieved, 'peak': peaks['bf16_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / peaks['bf16_sustained'], 'traffic': traffic.get(dom),
                'peak_source': peaks['source'] + ' (sustained bf16: the chains are timed inside a long step)',
                'peak_burst': peaks['bf16_burst'], 'frac_burst': achieved / peaks['bf16_burst'],
                'executed_mma': {'tflops': executed, 'frac': executed / peaks['bf16_sustained'], 'frac_burst': executed / peaks['bf16_burst'],
                                 'what': 'FLOPs the tensor pipe executed for these launches (padded shapes x MMAs per K step: the '
                                         'split-bf16 arithmetic issues 3 per K step in the forward / normals chains, 2 in the loss '
                                         'backward): the pipe-utilisation view; `achieved` counts each algorithmic FLOP once'},
                'launches_per_step': d['launches'] / args.steps, 'avg_launch_ms': d['ms'] / max(1, d['launches']),
                'kernel_share_of_step': d['ms'] / ms_prof_total,
                'algo_flops_per_launch': d['flops'] / max(1, d['launches']),
                'profiled_pass_ms_per_step': ms_prof_total / args.steps}
    step_tflops = FLOP_PER_SAMPLE_TRAIN * SAMPLES_PER_RAY * n / (ms_step * 1e-3) / 1e12
    # second kernel class: the wgrad GEMMs are HBM-bound by construction (every dY and X row is read once, bf16):
    # 16 hidden layers + heads: 17 (dY, X) pairs of 256 bf16 columns per sample row = 1024 B per row and layer
    wg = prof['wgrad_tc']
    wg_bytes = 17 * 1024 * n * SAMPLES_PER_RAY * args.steps
    wg_gbs = wg_bytes / (wg['ms'] * 1e-3) / 1e9 if wg['ms'] else 0.0
    roofline_wgrad = {'bound': 'hbm', 'kernel': 'wgrad_tc', 'achieved': wg_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                      'frac': wg_gbs / peaks['hbm_gbs'], 'traffic': traffic.get('wgrad_tc'),
                      'kernel_share_of_step': wg['ms'] / ms_prof_total}
    hbm_kernels = time_hbm_kernels(dev, peaks) if not args.no_hbm else None

    # ---- CPU baselines on this box's host cores: the imported reference where /root/reference
    # exists, else the oracle port; 4096-ray sample, 1 warm-up + median of 3 ----
    cpu = cpu_cfg1 = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        nc = args.cpu_rays if args.cpu_rays > 0 else 4096
        sec_t, kind = time_cpu(nc, 3, 1, 'train')
        cpu = {'value': nc / sec_t, 'unit': UNIT, 'cores': cores, 'kind': kind,
               'sample': f'{nc} rays of the {n}-ray batch (config 2 scaled as BASELINE.md section 3 prescribes), fwd + losses + bwd, '
                         f'median of 3 after 1 warm-up; torch threads={cores}; {cpu_model_name()}'}
        sec_e, _ = time_cpu(nc, 3, 1, 'eval_fwd')
        sec_f, _ = time_cpu(nc, 3, 1, 'train_fwd')
        cpu_cfg1 = {'config': 'configs/blender_refnerf.gin forward of one 4096-ray batch on CPU (BASELINE config 1)', 'rays': nc,
                    'eval_forward_s': sec_e, 'eval_forward_rays_per_s': nc / sec_e, 'train_forward_s': sec_f,
                    'train_forward_rays_per_s': nc / sec_f, 'kind': kind, 'cores': cores,
                    'frame_800x800_s_extrapolated': 640000 / (nc / sec_e)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'bf16': 'bf16', 'fp32': 'f32',
                  'bf16x3': 'bf16x3 (split-bf16 operands hi+lo, 3 tensor-core MMAs per product, fp32 accumulate: the arithmetic that '
                            'meets every north_star tolerance against the fp32 reference)',
                  'fp16': 'f16 (fp16 weights, activations and dynamically scaled gradient tiles; fp32 accumulate)'}[args.precision],
        'data': 'synthetic',
        'config': {'workload': f'configs/blender_refnerf.gin single training step, {n}-ray batch per GPU, NerfMLP at '
                               'both levels (single_mlp), fwd (incl. density-gradient normals) + losses + bwd + Adam',
                   'rays_per_gpu': n, 'levels': 2, 'samples_per_level': 128, 'precision': args.precision,
                   'chain_kernel': ('chain_x3t (split-bf16 chains, running activation in tensor memory)'
                                    if os.environ.get('RN_X3_TS', '1') != '0' else 'chain_x3 (activation tile in shared memory)')
                                   if args.precision == 'bf16x3' else 'chain_pair',
                   'board': 'power-capped (clocks.reasons): executed-MMA throughput, not pipe utilisation, is what the step is bound by',
                   'l2': 'no explicit flush: per-step activation working set (>2 GB) exceeds the 126 MB L2',
                   'parallelism': (f'ray-sharded dp{world}, one NCCL gradient all-reduce per step; ' + str(allreduce_path))
                                  if world > 1 else 'single GPU'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': wl.h2d_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e, 'host_wall_ms_of_single_steps': e2e_step_ms},
        'host_enqueue_ms_per_step': host_enqueue_ms,
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'roofline_wgrad': roofline_wgrad,
        'step_dram_bytes': traffic.get('step_dram_bytes'),
        'hbm_kernels': hbm_kernels,
        'cpu_baseline': cpu,
        'cpu_config1_forward': cpu_cfg1,
        'torch_gpu_baseline': torch_gpu,
        'mlp_tflops_step': step_tflops,
        'mlp_frac_of_bf16_peak': step_tflops / peaks['bf16_sustained'],
        'mlp_frac_of_bf16_burst_peak': step_tflops / peaks['bf16_burst'],
        'kernel_classes': prof,
        'strong_scaling': strong,
        'render': render,
    }
    line.update(extra)
    print(json.dumps(line))


def dataclass_items(obj):
    import dataclasses
    return [(f.name, getattr(obj, f.name)) for f in dataclasses.fields(obj)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('REFNERF_B200_PRECISION', 'bf16x3'),
                    choices=['bf16', 'fp16', 'bf16x3', 'fp32'],
                    help='GEMM arithmetic of the headline; default = the mode that meets the north_star tolerances')
    ap.add_argument('--rays', type=int, default=16384)
    ap.add_argument('--cpu-rays', type=int, default=0, help='CPU sample size (0 = 4096 for the baseline legs, adaptive for --impl reference)')
    ap.add_argument('--render-chunk', type=int, default=65536)
    ap.add_argument('--no-render', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the fp16 throughput-mode and LLFF (configs 4 / 5) legs')
    ap.add_argument('--no-hbm', action='store_true', help='skip the stand-alone GB/s timing of the warp-per-ray kernels')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()

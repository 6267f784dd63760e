"""CPU emulation of GEMM-operand rounding schemes on the oracle, gated like tests/test_gpu_model.py.

Decides which tensor-core arithmetic can meet BASELINE.json's tolerances BEFORE a kernel is written:
every Linear of the oracle's MLP is replaced by an autograd Function whose forward product, dgrad product
(loss backward and the in-forward density-gradient pass separately) and wgrad product each round their two
operands as a given scheme would (fp32 accumulate in all cases):

    f32      no rounding
    bf16     8-bit significand operand
    fp16     11-bit significand operand (range handled by an ideal power-of-two scale)
    x2       bf16 hi + bf16 lo (16 bits); a product of two x2 operands drops lo*lo (the "bf16x3" 3-MMA scheme),
             a product of x2 by a 1-plane operand keeps both planes (2 MMAs)

Usage: python tools/precision_emul.py [case ...]        (TEST / DESIGN TOOL: imports oracle/)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refnerf_oracle as O  # noqa: E402
from tests._cases import CASES, case_params, load_case  # noqa: E402

PHASE = ['loss']


def planes(x, kind):
    """-> list of fp32 planes whose sum represents x in `kind`."""
    if kind == 'f32':
        return [x]
    if kind == 'bf16':
        return [x.bfloat16().float()]
    if kind == 'fp16':
        amax = float(x.abs().max())
        if amax == 0 or not np.isfinite(amax):
            return [x]
        s = 2.0 ** (12 - np.floor(np.log2(amax)))
        return [(x * s).half().float() / s]
    if kind == 'x2':
        hi = x.bfloat16().float()
        return [hi, (x - hi).bfloat16().float()]
    if kind == 'h2':     # fp16 hi + fp16 lo
        hi = x.half().float()
        return [hi, (x - hi).half().float()]
    raise ValueError(kind)


def product(a, ka, b, kb):
    """a [m,k] @ b [k,n] with operand schemes ka / kb."""
    pa, pb = planes(a, ka), planes(b, kb)
    out = pa[0] @ pb[0]
    if len(pa) == 2:
        out = out + pa[1] @ pb[0]
    if len(pb) == 2:
        out = out + pa[0] @ pb[1]
    return out


class EmuLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, spec):
        ctx.spec = spec
        ctx.save_for_backward(x, w)
        shp = x.shape
        y = product(x.reshape(-1, shp[-1]), spec['fwd'][0], w.t(), spec['fwd'][1]) + b
        return y.reshape(shp[:-1] + (w.shape[0],))

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        spec = ctx.spec
        g2 = gy.reshape(-1, gy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dg = spec['normals'] if PHASE[0] == 'normals' else spec['dgrad']
        gx = product(g2, dg[0], w, dg[1]).reshape(x.shape)
        gw = gb = None
        if PHASE[0] != 'normals':
            gw = product(g2.t(), spec['wgrad'][0], x2, spec['wgrad'][1])
            gb = g2.sum(0)
        return gx, gw, gb, None


def make_linear(spec_spatial, spec_view):
    def _linear(x, p, name):
        spec = spec_view if (name.startswith('viewdir_mlp') or name == 'rgb') else spec_spatial
        return EmuLinear.apply(x, p[name + '.weight'], p[name + '.bias'], spec)
    return _linear


def S(fwd, normals, dgrad, wgrad):
    return dict(fwd=fwd, normals=normals, dgrad=dgrad, wgrad=wgrad)


F = ('f32', 'f32')
SCHEMES = {
    'fp32':        (S(F, F, F, F),) * 2,
    'bf16x3 all':  (S(('x2', 'x2'), ('x2', 'x2'), ('x2', 'x2'), ('x2', 'x2')),) * 2,
    'fp16 all':    (S(('fp16', 'fp16'), ('fp16', 'fp16'), ('fp16', 'fp16'), ('fp16', 'fp16')),) * 2,
    # forward + normals split, loss backward single-plane
    'x3 fwd+nrm, dgrad bf16*Wx2, wgrad bf16': (S(('x2', 'x2'), ('x2', 'x2'), ('bf16', 'x2'), ('bf16', 'bf16')),) * 2,
    'x3 fwd+nrm, dgrad bf16, wgrad bf16':     (S(('x2', 'x2'), ('x2', 'x2'), ('bf16', 'bf16'), ('bf16', 'bf16')),) * 2,
    'x3 fwd+nrm, dgrad fp16, wgrad fp16':     (S(('x2', 'x2'), ('x2', 'x2'), ('fp16', 'fp16'), ('fp16', 'fp16')),) * 2,
    'x3 fwd, nrm bf16*Wx2, dgrad bf16*Wx2, wgrad bf16': (S(('x2', 'x2'), ('bf16', 'x2'), ('bf16', 'x2'), ('bf16', 'bf16')),) * 2,
    'x3 fwd, nrm bf16, dgrad bf16, wgrad bf16': (S(('x2', 'x2'), ('bf16', 'bf16'), ('bf16', 'bf16'), ('bf16', 'bf16')),) * 2,
    # mixed: split spatial net, single-plane view net
    'x3 spatial / fp16 view fwd, dgrad fp16, wgrad fp16': (
        S(('x2', 'x2'), ('x2', 'x2'), ('fp16', 'fp16'), ('fp16', 'fp16')),
        S(('fp16', 'fp16'), F, ('fp16', 'fp16'), ('fp16', 'fp16'))),
    'x3 spatial / bf16 view fwd, dgrad bf16*Wx2, wgrad bf16': (
        S(('x2', 'x2'), ('x2', 'x2'), ('bf16', 'x2'), ('bf16', 'bf16')),
        S(('bf16', 'bf16'), F, ('bf16', 'x2'), ('bf16', 'bf16'))),
    'x3 spatial / x2*bf16W view': (
        S(('x2', 'x2'), ('x2', 'x2'), ('bf16', 'x2'), ('bf16', 'bf16')),
        S(('x2', 'bf16'), F, ('bf16', 'x2'), ('bf16', 'bf16'))),
}


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / (np.abs(b) + 1e-3 * max(np.abs(b).max(), 1e-30))


def run(name, scheme):
    g, rays = load_case(name)
    mc, lc, lossc = CASES[name]
    p = {k: v.clone().requires_grad_(True) for k, v in case_params(g).items()}
    orig_linear, orig_grad = O._linear, torch.autograd.grad
    O._linear = make_linear(*SCHEMES[scheme])

    def grad_normals(*a, **k):
        PHASE[0] = 'normals'
        try:
            return orig_grad(*a, **k)
        finally:
            PHASE[0] = 'loss'
    torch.autograd.grad = grad_normals
    try:
        rend, hist = O.model_forward(p, rays, 1.0, True, True, mc, lc)
        loss = O.total_loss(rend, hist, rays, torch.tensor(g['gt_rgb']), lossc)
        torch.autograd.grad = orig_grad
        grads = orig_grad(loss, list(p.values()))
    finally:
        O._linear, torch.autograd.grad = orig_linear, orig_grad
    rep = {}
    for lvl in range(2):
        for k in ('density',):
            e = rel_err(hist[lvl][k].detach().numpy(), g[f'train_hist{lvl}_{k}'])
            rep[f'dens{lvl}_p999'] = np.quantile(e, 0.999)
        e = np.abs(hist[lvl]['rgb'].detach().numpy() - g[f'train_hist{lvl}_rgb'])
        rep[f'rgb{lvl}_p999'] = np.quantile(e, 0.999)
        e = np.abs(hist[lvl]['normals'].numpy() - g[f'train_hist{lvl}_normals'])
        rep[f'nrm{lvl}_mean'] = e.mean()
        rep[f'comp{lvl}'] = np.abs(rend[lvl]['rgb'].detach().numpy() - g[f'train_rend{lvl}_rgb']).max()
    worst, worst_k = 0.0, ''
    second = 0.0
    errs = {}
    for (k, _), gr in zip(p.items(), grads):
        ref_norm = float(g['grad_norm_' + k])
        gn = float(gr.double().norm())
        sub = gr.reshape(-1)[::97].numpy()
        ref_sub = g['grad_sub_' + k]
        e = max(abs(gn - ref_norm) / max(ref_norm, 1e-30),
                float(np.linalg.norm(sub - ref_sub) / max(np.linalg.norm(ref_sub), 1e-30)))
        errs[k] = e
    ks = sorted(errs, key=errs.get, reverse=True)
    rep['grad_worst'] = errs[ks[0]]
    rep['grad_2nd'] = errs[ks[1]]
    rep['grad_median'] = float(np.median(list(errs.values())))
    rep['worst_name'] = ks[0]
    return rep


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    cases = sys.argv[1:] or ['blender_trained', 'blender_pert', 'llff_geom', 'blender_init']
    for name in cases:
        print(f'== {name}')
        for scheme in SCHEMES:
            r = run(name, scheme)
            print(f'  {scheme:52s} ' + ' '.join(f'{k}={v:.1e}' if not isinstance(v, str) else f'[{v}]' for k, v in r.items()),
                  flush=True)

"""CPU, world_size=2 over gloo: the ray-sharding + flat gradient all-reduce wiring reproduces the
single-process gradient.  The per-rank compute stand-in is the CPU oracle (the CUDA path
cannot run here); what is under test is refnerf_pl_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import refnerf_oracle as O
from refnerf_pl_b200 import parallel, synthetic

N_RAYS = 6


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(rays, gt, p):
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    rend, hist = O.model_forward(pp, rays, 1.0, False, True)
    O.total_loss(rend, hist, rays, gt).backward()
    return pp


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = parallel.init_distributed(backend='gloo')
    assert (r, w) == (rank, world)
    rays_all = {k: torch.tensor(v) for k, v in synthetic.blender_rays(N_RAYS, seed=5).items()}
    gt_all = torch.tensor(synthetic.gt_rgb(N_RAYS, 5))
    rays = parallel.shard_rays(rays_all, rank, world)
    lo, hi = parallel.shard_range(N_RAYS, rank, world)
    p = O.init_params(seed=0, bias_std=0.05)
    pp = _grads(rays, gt_all[lo:hi], p)
    params = list(pp.values())
    red = parallel.GradAllReducer(params)
    assert red.nbytes() == 1110158 * 4
    red.allreduce()
    if rank == 0:
        np.save(os.path.join(out_dir, 'flat.npy'), red.flat.numpy())
    # gradients that are views of ONE flat buffer (what ops.mlp_backward hands to autograd): reduced in place, no copies
    one = torch.cat([v.grad.reshape(-1) for v in params]).clone() * world   # undo nothing: fresh un-averaged stand-in
    off = 0
    for v in params:
        v.grad = one[off:off + v.numel()].view_as(v)
        off += v.numel()
    before = one.clone()
    red2 = parallel.GradAllReducer(params)
    assert red2._grads_as_one_buffer() is not None and red2._grads_as_one_buffer().data_ptr() == one.data_ptr()
    red2.allreduce()
    assert all(v.grad.data_ptr() != w_.data_ptr() for v, w_ in zip(params, red2.views))   # .grad still points into `one`
    if rank == 0:
        np.save(os.path.join(out_dir, 'direct.npy'), one.numpy())
        np.save(os.path.join(out_dir, 'direct_before.npy'), before.numpy())
    g = parallel.gather_rows(torch.full((2, 3), float(rank)))
    assert g.shape == (2 * world, 3) and float(g[-1, 0]) == world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (1, 7, 16384, 640000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(600)
def test_two_rank_allreduce_matches_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    flat = np.load(tmp_path / 'flat.npy')
    rays_all = {k: torch.tensor(v) for k, v in synthetic.blender_rays(N_RAYS, seed=5).items()}
    gt_all = torch.tensor(synthetic.gt_rgb(N_RAYS, 5))
    pp = _grads(rays_all, gt_all, O.init_params(seed=0, bias_std=0.05))
    ref = torch.cat([v.grad.reshape(-1) for v in pp.values()]).numpy()
    # equal shards + per-rank means => the rank-average equals the full-batch gradient
    assert np.linalg.norm(flat - ref) <= 1e-5 * np.linalg.norm(ref)
    # direct path: every rank fed (its averaged gradient x world) -> the mean over ranks is world x the averaged gradient
    direct = np.load(tmp_path / 'direct.npy')
    assert np.linalg.norm(direct - 2 * flat) <= 1e-5 * np.linalg.norm(2 * flat)

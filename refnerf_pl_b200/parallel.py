"""Ray-sharded data parallelism: one process per GPU, contiguous ray slices per rank,
replicated weights, ONE all-reduce of the flat gradient per step (the reference's only collective:
Lightning DDPPlugin -> torch DDP -> NCCL, train.py:88).  No data-path collective exists: every ray is
independent through sampling, MLP and compositing.  Works with NCCL (GPU) and gloo (CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, world, local_rank = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n rays owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, rank, world):
    """rays: dict of [N, C] arrays/tensors -> this rank's contiguous slice."""
    n = next(iter(rays.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in rays.items()}


class GradAllReducer:
    """Flat-buffer gradient mean across ranks: one collective of all parameter gradients per step
    (1 110 158 fp32 = 4.44 MB for the Ref-NeRF NerfMLP), launched on the current stream right after
    backward.  Matches DDP's averaging semantics."""

    def __init__(self, params, group=None):
        seen, self.params = set(), []
        for p in params:                      # single_mlp aliases nerf_mlp/prop_mlp: reduce each tensor once
            if id(p) not in seen and p.requires_grad:
                seen.add(id(p))
                self.params.append(p)
        self.group = group
        self.last_path = None
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def attach(self):
        """Point every .grad at its slice of the flat buffer so backward accumulates in place."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self

    def _grads_as_one_buffer(self):
        """`ops.mlp_backward` returns every parameter gradient of a call as a view of ONE flat fp32 buffer, and autograd
        keeps those views as `.grad` (later calls accumulate into them in place).  If the gradients tile one contiguous
        range of one storage, return that range as a 1-D tensor: the collective then runs on it directly -- no
        per-parameter copies into `self.flat`.  None if the layout is anything else."""
        gs = [p.grad for p in self.params]
        if any(g is None or g.dtype != torch.float32 or not g.is_contiguous() for g in gs):
            return None
        st = gs[0].untyped_storage()
        if any(g.untyped_storage().data_ptr() != st.data_ptr() for g in gs):
            return None
        spans = sorted((g.storage_offset(), g.numel()) for g in gs)
        off = spans[0][0]
        for o, n in spans:
            if o != off:
                return None
            off += n
        return torch.empty(0, dtype=torch.float32, device=gs[0].device).set_(st, spans[0][0], (off - spans[0][0],))

    def allreduce(self, async_op=False):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        direct = self._grads_as_one_buffer()
        self.last_path = 'direct (collective on the backward\'s flat buffer)' if direct is not None else 'copy into a flat buffer'
        if direct is not None:
            if world == 1:
                return None
            if dist.get_backend(self.group) == 'nccl':     # the mean is taken inside the collective
                return dist.all_reduce(direct, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
            work = dist.all_reduce(direct, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                return work
            direct.div_(world)
            return None
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        if world == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(world)
        return None

    def nbytes(self):
        return self.flat.numel() * 4


def gather_rows(t, group=None):
    """All-gather equally sized row blocks (eval: each rank renders a contiguous block of chunks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t.contiguous(), group=group)
    return torch.cat(outs, dim=0)

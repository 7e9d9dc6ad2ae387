"""Multi-GPU plumbing: one process per GPU, batch sharded over ranks.

Every op of the path is per-sample (no BatchNorm, GroupNorm statistics are per sample), so inference
shards the batch with NO data-path collective; the only collectives are the timing reduction used
by bench.py and, optionally, gathering the predicted flows.  The reference does this with a
single-process nn.DataParallel (tools/engine.py:63-64).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous shard [begin, end) of `total` samples for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(p, rank, world):
    """p = [xyz1, xyz2] with a global batch on dim 0 -> this rank's slice (DataParallel's scatter)."""
    b = p[0].shape[0]
    lo, hi = shard_range(b, rank, world)
    return [t[lo:hi] for t in p]


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timing: the slowest rank defines the step)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_batch(local, world_sizes=None):
    """All-gather per-rank results [b_r, ...] back into the global batch order (equal or ragged shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[dist.get_rank()] = local.shape[0]
    dist.all_reduce(counts)
    mx = int(counts.max())
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:int(c)] for o, c in zip(outs, counts)], dim=0)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


# ----------------------------------------------------------------------------------------------------------------------
# training: DDP-style gradient all-reduce.  The reference reduces gradients inside a single-process
# nn.DataParallel (tools/engine.py:63-64); here every rank owns one GPU and its shard of the batch, and the ONLY collective
# of a training step is the sum of the 192 034 fp32 parameter gradients (750 KiB; 934 KiB with refine_block) -- one bucket,
# latency-bound, NVLS-reduced inside the switch when NCCL enables it.
# ----------------------------------------------------------------------------------------------------------------------
def ddp(model, local_rank=None):
    """Wrap `model` in torch's DistributedDataParallel (one all-reduce bucket holds every gradient; the reduction starts as
    soon as backward has produced the last of them).  The custom autograd Functions of train.py are ordinary graph nodes
    to DDP: it only hooks the parameters' gradient accumulators."""
    from torch.nn.parallel import DistributedDataParallel
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    ids = None if local_rank is None else [local_rank]
    return DistributedDataParallel(model, device_ids=ids, output_device=local_rank, gradient_as_bucket_view=True,
                                   broadcast_buffers=False)


def allreduce_gradients(params, average=True):
    """Explicit form of the same exchange for callers that do not wrap the model: flatten every present gradient into one
    fp32 buffer, ONE all-reduce, scatter back (sum -> mean over ranks, as DDP).  Returns the number of bytes reduced."""
    params = [p for p in params if p.grad is not None]
    if not params or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel() * flat.element_size()

"""kNN utility with the reference's signatures (model/pointconv.py:4-39)."""
import torch

from . import ops


def square_distance(src, dst):
    """model/pointconv.py:4-25.  Dense [B,N,M] distance matrix -- plumbing kept for API parity only
    (the kernels never materialise it)."""
    d = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    d += torch.sum(src ** 2, -1).unsqueeze(2)
    d += torch.sum(dst ** 2, -1).unsqueeze(1)
    return d


def knn_point(nsample, xyz, new_xyz):
    """model/pointconv.py:28-39: [B,S,nsample] int64 ids of the nsample nearest `xyz` points of every
    `new_xyz` query (unsorted), computed by pvraft_knn_fwd (mode 1 = this file's distance op order)."""
    idx = ops.knn(xyz.detach().contiguous().float(), new_xyz.detach().contiguous().float(), nsample, mode=1)
    return idx.long()

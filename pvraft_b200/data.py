"""Input pipeline pieces: a drop-in `Batch` (datasets/generic.py:6-66) that collates into ONE
pinned host buffer and moves to the device with ONE asynchronous copy, and the random sub-sampling of
datasets/generic.py:183-189 on the device.  File formats and the Dataset classes are the caller's and stay unchanged."""
import torch


class Batch:
    """`collate_fn` of the reference's DataLoaders (tools/engine.py:43-48).  Same surface: `batch['sequence'] -> [pc1, pc2]`,
    `batch['ground_truth'] -> [mask, flow]`, `.to(device)`, `.pin_memory()`.  The four tensors are views of one flat fp32
    buffer, so `pin_memory()` pins once and `to()` issues a single host-to-device copy (non_blocking when pinned)."""
    _KEYS = (('sequence', 0), ('sequence', 1), ('ground_truth', 0), ('ground_truth', 1))

    def __init__(self, batch):
        parts = []
        for key, ind in self._KEYS:
            parts.append(torch.cat([item[key][ind] for item in batch], 0).float())
        self._shapes = [tuple(t.shape) for t in parts]
        self._flat = torch.cat([t.reshape(-1) for t in parts])
        self._view()

    def _view(self):
        self.data = {'sequence': [], 'ground_truth': []}
        off = 0
        for (key, _), shape in zip(self._KEYS, self._shapes):
            n = 1
            for d in shape:
                n *= d
            self.data[key].append(self._flat[off:off + n].view(shape))
            off += n

    def __getitem__(self, item):
        return self.data[item]

    def to(self, *args, **kwargs):
        kwargs.setdefault('non_blocking', self._flat.is_pinned())
        self._flat = self._flat.to(*args, **kwargs)
        self._view()
        return self

    def pin_memory(self):
        self._flat = self._flat.pin_memory()
        self._view()
        return self


def subsample(points, nb_points, generator=None, extra=()):
    """datasets/generic.py:183-189 on the device: a random subset of `nb_points` rows of `points` [n,3] (and the same rows of
    every tensor in `extra`, e.g. the ground-truth mask / flow of the first scan)."""
    n = points.shape[0]
    if n < nb_points:
        raise ValueError(f'cloud has {n} points, {nb_points} requested')
    ind = torch.randperm(n, device=points.device, generator=generator)[:nb_points]
    return (points[ind],) + tuple(e[ind] for e in extra)

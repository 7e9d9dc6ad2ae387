"""kNN graph of a point cloud -- mirror of the reference `Graph` (model/flot/graph.py:4-89).

`Graph.construct_graph(pcloud, nb_neighbors)` keeps the reference's signature and attributes
(`edges` = flat GLOBAL neighbour ids b*N + j as int64, `edge_feats` = neighbour - centre,
`k_neighbors`, `size`) but is built by the brute-force kNN kernel (pvraft_knn_fwd) instead of a
full N x N argsort; the kernels consume the compact int32 LOCAL adjacency `nbr` [B,N,k].
"""
import os

import torch

from . import ops


class Graph:
    locality_order = os.environ.get('PVRAFT_EDGE_ORDER', '1') != '0'

    def __init__(self, nbr, edge_feats, k_neighbors, size, order=None):
        self.nbr = nbr                    # int32 [B,N,k] local ids
        self._rel = edge_feats            # f32 [B,N,k,3]
        self.order = order                # int32 [B,N] or None: processing order of the SetConv edge kernel (a Morton rank table)
        self.k_neighbors = k_neighbors
        self.size = tuple(size)
        self._edges = None

    @property
    def edges(self):
        """Flat int64 global row ids, as model/flot/graph.py:77-79 builds them."""
        if self._edges is None:
            b, n, k = self.nbr.shape
            offs = (torch.arange(b, device=self.nbr.device, dtype=torch.int64) * n).view(b, 1, 1)
            self._edges = (self.nbr.long() + offs).reshape(-1)
        return self._edges

    @property
    def edge_feats(self):
        """[B*N*k, 3] neighbour - centre (model/flot/graph.py:69-74)."""
        return self._rel.reshape(-1, 3)

    @staticmethod
    def construct_graph(pcloud, nb_neighbors):
        b, n, _ = pcloud.shape
        if nb_neighbors != ops.KNN:
            raise NotImplementedError('the B200 SetConv kernels are built for 32 neighbours (the only value the '
                                      'reference uses, model/extractor.py:9)')
        if n < nb_neighbors:
            raise ValueError(f'need at least {nb_neighbors} points per cloud, got {n}')
        pc = pcloud.detach().contiguous().float()
        nbr, rel = ops.knn(pc, pc, nb_neighbors, mode=0, want_rel=True)
        # spatially coherent processing order for the edge kernel (changes no result: the 8 warps of a CTA then gather
        # overlapping neighbourhoods, which L1 serves)
        order = ops.point_order(pc, as_int32=True) if Graph.locality_order and n >= 64 else None
        return Graph(nbr, rel, nb_neighbors, [b * n, b * n], order)

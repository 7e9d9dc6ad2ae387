"""FlotEncoder -- mirror of model/extractor.py:7-24 (three SetConvs over the 32-NN graph)."""
import torch.nn as nn

from .gconv import SetConv
from .graph import Graph


class FlotEncoder(nn.Module):
    def __init__(self, num_neighbors=32):
        super().__init__()
        n = 32
        self.num_neighbors = num_neighbors
        self.feat_conv1 = SetConv(3, n)
        self.feat_conv2 = SetConv(n, 2 * n)
        self.feat_conv3 = SetConv(2 * n, 4 * n)

    def forward(self, pc, graph=None, point_major=False):
        """pc [B,N,3] -> (features [B,128,N] (or [B,N,128] when point_major), graph).
        `graph` lets the caller reuse an adjacency already built for the same cloud (the reference
        rebuilds pc1's graph for the context encoder, model/RAFTSceneFlow.py:25,31)."""
        if graph is None:
            graph = Graph.construct_graph(pc, self.num_neighbors)
        x = self.feat_conv1.forward_deferred(pc, graph)
        x = self.feat_conv2.forward_deferred(x, graph)
        x = self.feat_conv3.forward_deferred(x, graph)
        return x.materialize(transpose_out=not point_major), graph

"""FlotRefine -- mirror of model/refine.py:6-22."""
import torch.nn as nn

from . import ops
from .gconv import SetConv
from .ops import ACT_LRELU, IN_GN


class FlotRefine(nn.Module):
    def __init__(self):
        super().__init__()
        n = 32
        self.ref_conv1 = SetConv(3, n)
        self.ref_conv2 = SetConv(n, 2 * n)
        self.ref_conv3 = SetConv(2 * n, 4 * n)
        self.fc = nn.Linear(4 * n, 3)

    def forward(self, flow, graph):
        flow = flow.detach().contiguous().float()
        x = self.ref_conv1.forward_deferred(flow, graph)
        x = self.ref_conv2.forward_deferred(x, graph)
        x = self.ref_conv3.forward_deferred(x, graph)
        # flow + fc(lrelu(gn3(z)))   (refine.py:21-22), one fused launch
        return ops.linear(x.z, self.fc.weight.detach(), self.fc.bias.detach(), in_mode=IN_GN, in_stats=x.stats,
                          in_gamma=x.gamma, in_beta=x.beta, in_count=x.count, in_act=ACT_LRELU, in_slope=0.1,
                          residual=flow)

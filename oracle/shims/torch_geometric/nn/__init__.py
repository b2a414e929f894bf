"""Plain-torch restatement of PyG 2.5.3 NNConv / GCNConv (+ stubs for GATConv / EdgeConv which the
default GAE config, gae/config/train_config.yaml:4-12, never instantiates)."""
import math
import torch
from torch import nn as _nn
from . import norm  # noqa: F401


class NNConv(_nn.Module):
    """out_i = sum_{j->i} x_j @ reshape(nn(e_ji), [in, out]) + lin(x_i) + bias   (aggr='add',
    root_weight=True, bias=True; `lin` has no bias of its own).  edge_index[0]=source j,
    edge_index[1]=target i."""

    def __init__(self, in_channels, out_channels, nn, aggr="add", root_weight=True, bias=True):
        super().__init__()
        assert aggr == "add" and root_weight and bias
        self.in_channels, self.out_channels = in_channels, out_channels
        self.nn = nn
        self.lin = _nn.Linear(in_channels, out_channels, bias=False)
        self.bias = _nn.Parameter(torch.zeros(out_channels))
        bound = 1.0 / math.sqrt(in_channels)
        _nn.init.uniform_(self.lin.weight, -bound, bound)

    def forward(self, x, edge_index, edge_attr):
        src, dst = edge_index[0], edge_index[1]
        weight = self.nn(edge_attr).view(-1, self.in_channels, self.out_channels)
        x_j = x[src]
        msg = torch.matmul(x_j.unsqueeze(1), weight).squeeze(1)
        out = torch.zeros(x.shape[0], self.out_channels, dtype=x.dtype, device=x.device)
        out.index_add_(0, dst, msg)
        out = out + self.lin(x)
        return out + self.bias


class GCNConv(_nn.Module):
    """out = D^-1/2 (A + I) D^-1/2 (X W) + b with add_remaining_self_loops (exactly one self-loop
    per node, weight 1) and degree = weighted in-degree on targets."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = _nn.Linear(in_channels, out_channels, bias=False)
        self.bias = _nn.Parameter(torch.zeros(out_channels))
        _nn.init.xavier_uniform_(self.lin.weight)

    def forward(self, x, edge_index):
        n = x.shape[0]
        src, dst = edge_index[0], edge_index[1]
        keep = src != dst
        loop = torch.arange(n, dtype=src.dtype, device=src.device)
        src = torch.cat([src[keep], loop])
        dst = torch.cat([dst[keep], loop])
        w = torch.ones(src.shape[0], dtype=x.dtype, device=x.device)
        deg = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, dst, w)
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        norm_w = dinv[src] * w * dinv[dst]
        h = self.lin(x)
        out = torch.zeros_like(h).index_add_(0, dst, norm_w.unsqueeze(1) * h[src])
        return out + self.bias


class GATConv(_nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("shim: GATConv not restated (not in the default GAE config)")


class EdgeConv(_nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("shim: EdgeConv not restated (not in the default GAE config)")

import torch


class BatchNorm(torch.nn.Module):
    """torch_geometric.nn.norm.BatchNorm == BatchNorm1d under attribute `.module`."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)

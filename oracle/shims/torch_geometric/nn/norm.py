"""PyG 2.1.0 nn.LayerNorm: graph-wise normalisation over all nodes AND channels of
each graph when `batch` is given."""
import torch
from torch_scatter import scatter
from ..utils import degree


class LayerNorm(torch.nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(in_channels))
            self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, batch=None):
        if batch is None:
            x = x - x.mean()
            out = x / (x.std(unbiased=False) + self.eps)
        else:
            batch_size = int(batch.max()) + 1
            norm = degree(batch, batch_size, dtype=x.dtype).clamp_(min=1)
            norm = norm.mul_(x.size(-1)).view(-1, 1)
            mean = scatter(x, batch, dim=0, dim_size=batch_size,
                           reduce='add').sum(dim=-1, keepdim=True) / norm
            x = x - mean.index_select(0, batch)
            var = scatter(x * x, batch, dim=0, dim_size=batch_size,
                          reduce='add').sum(dim=-1, keepdim=True)
            var = var / norm
            out = x / (var + self.eps).sqrt().index_select(0, batch)
        if self.weight is not None and self.bias is not None:
            out = out * self.weight + self.bias
        return out

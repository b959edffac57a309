"""torch_geometric.nn.LayerNorm 2.1.0 (model.py:24,46, residual_layer.py:3).  With a `batch` vector it is
GRAPH-wise: mean and variance over all rows and channels of a graph.  On a CUDA tensor
without affine parameters -- the only way model.py uses it -- the call goes to this package's one-kernel
implementation (x2_graph_layernorm_fwd / _bwd); everything else is the plain composite."""
import torch

from ..utils import degree


class LayerNorm(torch.nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(in_channels))
            self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def reset_parameters(self):
        if self.weight is not None:
            torch.nn.init.ones_(self.weight)
            torch.nn.init.zeros_(self.bias)

    def forward(self, x, batch=None):
        if batch is None:
            x = x - x.mean()
            out = x / (x.std(unbiased=False) + self.eps)
        elif x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.size(1) % 4 == 0:
            from x2gnn_b200 import graph_norm
            counts = torch.bincount(batch, minlength=int(batch.max()) + 1)   # host sync, as in PyG (`int(batch.max())`)
            out = graph_norm.graph_layer_norm_rows(x, graph_norm.rowptr_from_counts(counts), self.eps)
        else:
            bsz = int(batch.max()) + 1
            norm = degree(batch, bsz, dtype=x.dtype).clamp_(min=1).mul_(x.size(-1)).view(-1, 1)
            mean = torch.zeros(bsz, x.size(-1), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
            mean = mean.sum(dim=-1, keepdim=True) / norm
            x = x - mean.index_select(0, batch)
            var = torch.zeros(bsz, x.size(-1), dtype=x.dtype, device=x.device).index_add_(0, batch, x * x)
            var = var.sum(dim=-1, keepdim=True) / norm
            out = x / (var + self.eps).sqrt().index_select(0, batch)
        if self.weight is not None:
            out = out * self.weight + self.bias
        return out

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels})"

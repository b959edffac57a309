"""torch_scatter 2.1.0 entry points used by the reference callers (readout.py:3, model.py:9, xgnn.py:12,
train_ema.py:7): scatter_add / scatter_mean / scatter with the index broadcast along `dim` and `dim_size`
output rows.  Built on Tensor.index_add_ / scatter_add_; the hot path of this package never calls them."""
import torch

__version__ = "2.1.0+x2gnn_b200.compat"


def _out_size(src, index, dim, dim_size):
    size = list(src.shape)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    size[dim] = int(dim_size)
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    if out is None:
        out = src.new_zeros(_out_size(src, index, dim, dim_size))
    if index.dim() == 1 and index.numel() == src.size(dim):
        return out.index_add_(dim, index, src)           # the only form the reference uses
    idx = index
    while idx.dim() < src.dim():
        idx = idx.unsqueeze(-1)
    return out.scatter_add_(dim, idx.expand_as(src), src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.dim()
    total = scatter_sum(src, index, dim, out, dim_size)
    count = torch.zeros(total.size(dim), dtype=src.dtype, device=src.device)
    flat = index if index.dim() == 1 else index.movedim(dim, -1).reshape(-1, index.size(dim))[0]
    count.index_add_(0, flat, torch.ones_like(flat, dtype=src.dtype))
    shape = [1] * total.dim()
    shape[dim] = -1
    return total / count.clamp_(min=1).view(shape)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    raise ValueError(f"x2gnn_b200.compat.torch_scatter: reduce={reduce!r} is not used by the X2-GNN callers")

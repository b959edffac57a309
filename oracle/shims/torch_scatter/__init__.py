"""TEST INFRASTRUCTURE ONLY (part of oracle/): plain-PyTorch restatement of the
torch-scatter 2.1.0 entry points the reference calls (readout.py:3, model.py:9,
xgnn.py:12).  Semantics follow SURVEY.md App. C: index broadcast along `dim`,
`dim_size` output rows, sum / mean / max reductions.  Never imported by the product."""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    return out / count


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _broadcast(index, src, dim)
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    else:
        size[dim] = int(index.max()) + 1 if index.numel() else 0
    res = torch.full(size, float('-inf'), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, index_b, src, reduce='amax', include_self=True)
    # torch-scatter fills empty segments with 0 for max
    res = torch.where(torch.isinf(res) & (res < 0), torch.zeros_like(res), res)
    return res, None


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == 'max':
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)

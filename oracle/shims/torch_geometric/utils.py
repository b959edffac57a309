"""PyG 2.1.0 utils.softmax (ptr=None branch) and remove_self_loops."""
import torch
from torch_scatter import scatter


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    assert ptr is None, "reference always passes ptr=None (propagate with Tensor edge_index)"
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = scatter(src, index, dim, dim_size=N, reduce='max')
    src_max = src_max.index_select(dim, index)
    out = (src - src_max).exp()
    out_sum = scatter(out, index, dim, dim_size=N, reduce='sum')
    out_sum = out_sum.index_select(dim, index)
    return out / (out_sum + 1e-16)


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def degree(index, num_nodes=None, dtype=None):
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((N,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)

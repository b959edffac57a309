"""torch_geometric.utils subset (atom_graph.py:6 imports remove_self_loops; PyG's LayerNorm uses degree)."""
import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.index_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])

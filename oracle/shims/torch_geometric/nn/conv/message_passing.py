"""PyG 2.1.0 MessagePassing.propagate for a Tensor edge_index, flow
'source_to_target': `*_j` arguments are lifted by edge_index[0] (source), `*_i` by
edge_index[1] (target), the message is summed over edge_index[1] into `size_i` rows
(SURVEY.md App. C)."""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kwargs):
        super().__init__()
        assert flow == 'source_to_target'
        self.aggr = aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        sig = inspect.signature(self.message).parameters
        n_i = None
        args = {}
        for name in sig:
            if name.endswith('_i') and name[:-2] in kwargs:
                v = kwargs[name[:-2]]
                n_i = v.size(self.node_dim)
                args[name] = v.index_select(self.node_dim, dst)
            elif name.endswith('_j') and name[:-2] in kwargs:
                v = kwargs[name[:-2]]
                if n_i is None:
                    n_i = v.size(self.node_dim)
                args[name] = v.index_select(self.node_dim, src)
            elif name == 'index':
                args[name] = dst
            elif name == 'ptr':
                args[name] = None
            elif name == 'size_i':
                args[name] = None  # filled below
            elif name in kwargs:
                args[name] = kwargs[name]
        if 'size_i' in sig:
            args['size_i'] = n_i
        out = self.message(**args)
        reduce = 'sum' if self.aggr == 'add' else self.aggr
        return scatter(out, dst, dim=self.node_dim, dim_size=n_i, reduce=reduce)

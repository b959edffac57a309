"""oracle.conv -- TEST INFRASTRUCTURE.  CPU restatement of
sbftransformer_conv.py:17-162 (SBFTransformerConv) with the PyG 2.1.0
propagate / softmax / sum-aggregate semantics written out.

The op sequence deliberately mirrors the reference's composite path (index_select
gathers of Q/K/V, T-row lin_edge / lin_sbf, scatter softmax with +1e-16, scatter-add
aggregation) so that timing it on host cores is a fair stand-in for the reference's
own CPU execution (`bench.py --impl reference`, kind="port").  Backward is autograd.
"""
import math
import torch
import torch.nn.functional as F


def glorot_ortho_(w: torch.Tensor, scale: float = 2.0) -> torch.Tensor:
    """initializer.py:29-34."""
    torch.nn.init.orthogonal_(w)
    with torch.no_grad():
        w.mul_(torch.sqrt(scale / ((w.size(0) + w.size(1)) * w.var())))
    return w


def segment_softmax(a: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """torch_geometric.utils.softmax 2.1.0, ptr=None branch (max is not detached)."""
    idx = index.view(-1, 1).expand_as(a)
    m = torch.full((n, a.size(1)), float("-inf"), dtype=a.dtype, device=a.device)
    m = m.scatter_reduce(0, idx, a, reduce="amax", include_self=True)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    o = (a - m.index_select(0, index)).exp()
    z = torch.zeros((n, a.size(1)), dtype=a.dtype, device=a.device).scatter_add(0, idx, o)
    return o / (z.index_select(0, index) + 1e-16)


def sbfconv_forward(p: dict, sbf, rbf, x, edge_index, edge_attr=None, *, heads: int,
                    out_channels: int, concat: bool = True, root_weight: bool = True,
                    dropout: float = 0.0, training: bool = False):
    """p: dict of state_dict-named tensors ('lin_key.weight', ...).  Returns (out, alpha)."""
    H, C = heads, out_channels
    g = lambda k: p.get(k)
    x_src = x * F.linear(rbf, p["lin_rbf.weight"])                              # :99-100
    q = F.linear(x, p["lin_query.weight"], g("lin_query.bias")).view(-1, H, C)  # :105
    k = F.linear(x_src, p["lin_key.weight"], g("lin_key.bias")).view(-1, H, C)  # :106
    v = F.linear(x_src, p["lin_value.weight"], g("lin_value.bias")).view(-1, H, C)
    src, dst = edge_index[0], edge_index[1]
    q_i = q.index_select(0, dst)                    # propagate: _i <- edge_index[1]
    k_j = k.index_select(0, src)                    #            _j <- edge_index[0]
    v_j = v.index_select(0, src)
    ea = None
    if g("lin_edge.weight") is not None:
        ea = F.linear(edge_attr, p["lin_edge.weight"]).view(-1, H, C)           # :144
        k_j = k_j + ea                                                          # :146
    sg = F.linear(sbf, p["lin_sbf.weight"], g("lin_sbf.bias"))                  # :148
    alpha = (q_i * k_j).sum(dim=-1) / math.sqrt(C)                              # :150
    alpha = segment_softmax(alpha, dst, x.size(0))                              # :151
    a_used = F.dropout(alpha, p=dropout, training=training)                     # :153
    out = v_j
    if ea is not None:
        out = out + ea                                                          # :157
    out = out * sg.view(-1, H, C) * a_used.view(-1, H, 1)                       # :159-160
    agg = torch.zeros((x.size(0), H, C), dtype=out.dtype, device=out.device)
    agg = agg.index_add(0, dst, out)                # aggr='add' over edge_index[1]
    out = agg.view(-1, H * C) if concat else agg.mean(dim=1)                    # :115-118
    if root_weight:
        x_r = F.linear(x, p["lin_skip.weight"], g("lin_skip.bias"))             # :121
        if g("lin_beta.weight") is not None:
            beta = F.linear(torch.cat([out, x_r, out - x_r], dim=-1), p["lin_beta.weight"])
            beta = beta.sigmoid()
            out = beta * x_r + (1 - beta) * out                                 # :123-125
        else:
            out = out + x_r                                                     # :127
    return out, alpha


class OracleSBFTransformerConv(torch.nn.Module):
    """Same constructor, sub-module names (=> state_dict keys) and initialisers as
    sbftransformer_conv.py:17-91 (with the `init.zeros` defect repaired)."""

    def __init__(self, in_channels, out_channels, heads=1, sbf_dim=16, rbf_dim=16,
                 concat=True, beta=False, dropout=0., edge_dim=None, bias=True,
                 root_weight=True):
        super().__init__()
        L = torch.nn.Linear
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.root_weight, self.dropout = concat, root_weight, dropout
        self.beta = beta and root_weight
        self.edge_dim = edge_dim
        D = heads * out_channels
        self.lin_key = L(in_channels, D)
        self.lin_query = L(in_channels, D)
        self.lin_value = L(in_channels, D)
        if edge_dim is not None:
            self.lin_edge = L(edge_dim, D, bias=False)
        else:
            self.register_parameter("lin_edge", None)
        skip_out = D if concat else out_channels
        self.lin_skip = L(in_channels, skip_out, bias=bias)
        if self.beta:
            self.lin_beta = L(3 * skip_out, 1, bias=False)
        else:
            self.register_parameter("lin_beta", None)
        self.lin_sbf = L(sbf_dim, D, bias=True)
        self.lin_rbf = L(rbf_dim, in_channels, bias=False)
        glorot_ortho_(self.lin_sbf.weight)
        glorot_ortho_(self.lin_rbf.weight)
        torch.nn.init.zeros_(self.lin_sbf.bias)

    def forward(self, sbf, rbf, x, edge_index, edge_attr=None, return_attention_weights=None):
        p = dict(self.named_parameters())
        out, alpha = sbfconv_forward(
            p, sbf, rbf, x, edge_index, edge_attr, heads=self.heads,
            out_channels=self.out_channels, concat=self.concat,
            root_weight=self.root_weight, dropout=self.dropout, training=self.training)
        if isinstance(return_attention_weights, bool):
            return out, (edge_index, alpha)
        return out

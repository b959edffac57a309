"""oracle.model -- TEST INFRASTRUCTURE.  CPU restatement of the unchanged CALLERS of the
hot path, used for the end-to-end U0 parity check and the molecules/s CPU baseline:
xgnn.py:15-75 (xgnn_poly), model.py:11-54 (SBFTransformer), readout.py:7-43 (AtomWise),
residual_layer.py:5-27, atom_embedding.py:10-25, PyG graph-wise LayerNorm (App. C).
Module/parameter names reproduce the reference state_dict keys."""
import torch
from torch import nn
import torch.nn.functional as F

from . import bases, graph
from .conv import OracleSBFTransformerConv, glorot_ortho_


def _go_linear(i, o):
    lin = nn.Linear(i, o)
    glorot_ortho_(lin.weight)
    nn.init.zeros_(lin.bias)
    return lin


class EmbeddingBlock(nn.Module):                      # atom_embedding.py:10-25
    def __init__(self, embedding_size=128):
        super().__init__()
        self.embedding = nn.Embedding(10, embedding_size, padding_idx=0, max_norm=3.0,
                                      scale_grad_by_freq=True)
        self.lin = _go_linear(embedding_size, embedding_size)

    def forward(self, z):
        return F.silu(self.lin(self.embedding(z)))


class ResidualLayer(nn.Module):                       # residual_layer.py:5-27
    def __init__(self, c):
        super().__init__()
        self.lin0 = _go_linear(c, c)
        self.lin1 = _go_linear(c, c)

    def forward(self, x):
        return F.silu(self.lin1(F.silu(self.lin0(x)))) + x


class AtomWise(nn.Module):                            # readout.py:7-43
    def __init__(self, in_channels, rbf_dim, num_target=1, mlp_depth=3):
        super().__init__()
        layers = []
        for _ in range(mlp_depth - 1):
            layers += [_go_linear(in_channels, in_channels), nn.SiLU()]
        layers.append(_go_linear(in_channels, num_target))
        self.mlp = nn.ModuleList(layers)
        self.lin_rbf = _go_linear(rbf_dim, in_channels)

    def forward(self, x, rbf, num_atoms, edge_index_0):
        out = self.lin_rbf(rbf) * x
        out = torch.zeros(num_atoms, out.size(1), dtype=out.dtype,
                          device=out.device).index_add(0, edge_index_0, out)
        for layer in self.mlp:
            out = layer(out)
        return out


def graph_layer_norm(x, batch, num_graphs, eps=1e-8):
    """torch_geometric.nn.LayerNorm(affine=False) 2.1.0 with `batch` (App. C)."""
    cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype)
    norm = (cnt * x.size(-1)).view(-1, 1)
    zeros = lambda: torch.zeros(num_graphs, x.size(1), dtype=x.dtype, device=x.device)
    mean = zeros().index_add(0, batch, x).sum(-1, keepdim=True) / norm
    x = x - mean.index_select(0, batch)
    var = zeros().index_add(0, batch, x * x).sum(-1, keepdim=True) / norm
    return x / (var + eps).sqrt().index_select(0, batch)


class SBFTransformer(nn.Module):                      # model.py:11-54
    conv_cls = OracleSBFTransformerConv

    def __init__(self, conv_layers, emb_size, sbf_dim, rbf_dim=16, in_channels=128, heads=8):
        super().__init__()
        self.edgenn = nn.Sequential(_go_linear(emb_size, emb_size), nn.SiLU(),
                                    _go_linear(emb_size, emb_size))
        self.convs = nn.ModuleList([
            self.conv_cls(in_channels=in_channels, out_channels=in_channels // heads,
                          heads=heads, sbf_dim=sbf_dim * rbf_dim, rbf_dim=rbf_dim, dropout=0,
                          edge_dim=emb_size) for _ in range(conv_layers)])
        self.readouts = nn.ModuleList([AtomWise(in_channels, rbf_dim) for _ in range(conv_layers + 1)])
        self.bf_skip = nn.ModuleList([ResidualLayer(in_channels) for _ in range(conv_layers)])
        self.af_skip = nn.ModuleList([nn.Sequential(ResidualLayer(in_channels), ResidualLayer(in_channels))
                                      for _ in range(conv_layers)])
        self.dense_bf_skip = nn.ModuleList([_go_linear(in_channels, in_channels) for _ in range(conv_layers)])
        self.conv_layers = conv_layers

    def forward(self, x, edge_index, edge_attr, batch, edge_sbf, node_rbf, edge_index_0,
                atom_batch, num_graphs):
        edge_attr = self.edgenn(edge_attr)                                        # :39
        out = x
        n_atoms = atom_batch.size(0)
        results = self.readouts[0](out, node_rbf, n_atoms, edge_index_0)          # :41
        for i in range(self.conv_layers):
            res0 = out
            out = self.convs[i](sbf=edge_sbf, rbf=node_rbf, x=out, edge_index=edge_index,
                                edge_attr=edge_attr)                              # :45
            out = graph_layer_norm(out, batch, num_graphs)                        # :46
            out = self.bf_skip[i](out)
            out = F.silu(self.dense_bf_skip[i](out)) + res0
            out = self.af_skip[i](out)
            results = results + self.readouts[i + 1](out, node_rbf, n_atoms, edge_index_0)
        mol = torch.zeros(num_graphs, results.size(1), dtype=results.dtype,
                          device=results.device).index_add(0, atom_batch, results)  # :53
        return mol.view(-1)


class XGNNPoly(nn.Module):                            # xgnn.py:15-75
    fin_cls = SBFTransformer

    def __init__(self, conv_layers=4, sbf_dim=7, rbf_dim=16, in_channels=256, heads=16,
                 embedding_size=128):
        super().__init__()
        self.sbf_dim, self.rbf_dim = sbf_dim, rbf_dim
        self.emb_block = EmbeddingBlock(embedding_size)
        self.rbf_layer = _Freq(rbf_dim)
        self.fin_model = self.fin_cls(conv_layers, embedding_size, sbf_dim, rbf_dim, in_channels, heads)
        self.mat_trans = _go_linear(338, 2 * embedding_size)
        self.rbf_trans = _go_linear(rbf_dim, embedding_size)        # declared, unused (:30-32)
        self.emb_trans = _go_linear(2 * embedding_size, in_channels)

    # hooks overridden by GPU-side subclasses in tests; default = oracle pieces
    def _triplets(self, edge_index, n):
        return graph.vertex_to_edge_2(edge_index, n)

    def _envelope(self, d):
        return bases.poly_envelop(d, 5.0, 5)

    def _sbf(self, d, ang, src):
        return bases.f_b_2d(d, ang, src, self.sbf_dim, self.rbf_dim, 5.0, 5)

    def _rbf(self, d):
        return bases.radial_basis(d, self.rbf_layer.frequencies, 5.0)

    def forward(self, data: dict):
        pos, ei = data["atom_pos"], data["edge_index"]
        d = torch.norm(pos[ei[0]] - pos[ei[1]], dim=1)                              # :39
        B = int(data["num_graphs"])
        batch = torch.arange(B, device=d.device).repeat_interleave(data["edge_num"])  # :44
        env = self._envelope(d)[:, None]                                            # :49
        tri, a_j, a_i, a_k = self._triplets(ei, data["x"].size(0))                  # :52
        tri, a_j, a_i, a_k = (t.to(d.device) for t in (tri, a_j, a_i, a_k))
        neo_x = F.silu(self.mat_trans(data["edge_attr"] * env))                     # :54-55
        neo_edge_attr = self.emb_block(data["x"])[a_j]                              # :57-58
        ji = pos[a_i] - pos[a_j]
        jk = pos[a_k] - pos[a_j]
        cos = (ji * jk).sum(1)
        sin = torch.norm(torch.linalg.cross(ji, jk), dim=1)
        edge_sbf = self._sbf(d, torch.atan2(sin, cos), tri[0])                      # :65
        node_rbf = self._rbf(d) * env                                               # :68-69
        neo_x = F.silu(self.emb_trans(neo_x))                                       # :70
        return self.fin_model(neo_x, tri, neo_edge_attr, batch, edge_sbf, node_rbf,
                              ei[0], data["batch"], B)


class _Freq(nn.Module):                                # radial_basis_layer.py:26-34
    def __init__(self, R):
        super().__init__()
        self.frequencies = nn.Parameter(bases.radial_frequencies(R))

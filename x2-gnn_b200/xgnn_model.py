"""Host-side harness model: the callers of the hot path (reference xgnn.py:15-75 `xgnn_poly`,
model.py:11-54 `SBFTransformer`, readout.py:7-43, residual_layer.py, atom_embedding.py) restated
with the reference's module names so a reference checkpoint (`ckpt/U0_ckpt.pth`)
loads with `load_state_dict` unchanged.  It exists so that U0 predictions and full training
steps (molecules/s) can be measured on the GPU box, where the reference tree is not present.

The hot path -- radius/triplet graphs, envelope, radial + 2-D Fourier-Bessel bases, SBFTransformerConv
-- runs on the sm_100a kernels of this package; the dense layers around it use `TCLinear` (the same
tcgen05 3xTF32 GEMMs behind an nn.Linear); graph LayerNorm and the readout scatter
are CUDA PyTorch ops.  Differences from the reference's control flow, none of which change results:
triplets are built on the GPU (no `.to('cpu')` round trip, xgnn.py:52-53); `num_graphs` is taken from
the batch instead of `int(batch.max())` host syncs (model.py:46,53); and `edgenn` is evaluated once per
ATOM and gathered per triplet instead of on all T gathered rows (its input emb(Z)[atom_j] has one
distinct row per atom -- SURVEY.md §8f row 1; row-wise MLP, so edgenn(emb[a_j]) == edgenn(emb)[a_j]).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from .angular_basis_layer import F_B_2D
from .edge_graph import vertex_to_edge_2
from .envelop import poly_envelop
from .geometry import bond_lengths, triplet_angles
from .graph_norm import graph_layer_norm_rows, rowptr_from_counts
from .radial_basis_layer import RadialBasis
from . import readout_sum
from .sbftransformer_conv import Glorot_Ortho_, SBFTransformerConv
from .tc_linear import TCLinear


def _lin(i, o):
    l = TCLinear(i, o)      # nn.Linear parameters; tensor-core GEMMs when out_features % 128 == 0
    Glorot_Ortho_(l.weight)
    nn.init.zeros_(l.bias)
    return l


class _EmbedByFreq(torch.autograd.Function):
    """weight[z] whose weight gradient is scaled by the inverse frequency of every id in the batch -- the
    arithmetic of nn.Embedding(scale_grad_by_freq=True, padding_idx=0) written with gather / index_add only,
    so that it neither synchronises nor sorts (F.embedding's backward does) and can sit inside a CUDA graph.
    `counts` [rows] is the batch histogram of `z`, built once per batch by `XGNNPoly.prepare`."""

    @staticmethod
    def forward(ctx, weight, z, counts):
        ctx.save_for_backward(z, counts)
        ctx.rows = weight.size(0)
        return weight.index_select(0, z)

    @staticmethod
    def backward(ctx, g):
        z, counts = ctx.saved_tensors
        gw = torch.zeros(ctx.rows, g.size(1), dtype=g.dtype, device=g.device).index_add_(0, z, g)
        gw = gw / counts.clamp(min=1).to(g.dtype).unsqueeze(1)
        gw[0].zero_()                                     # padding_idx = 0
        return gw, None, None


class EmbeddingBlock(nn.Module):
    def __init__(self, embedding_size=128):
        super().__init__()
        self.embedding = nn.Embedding(10, embedding_size, padding_idx=0, max_norm=3.0, scale_grad_by_freq=True)
        self.lin = _lin(embedding_size, embedding_size)

    def forward(self, z, z_rows=None, z_counts=None):
        """With the batch's distinct ids `z_rows` and histogram `z_counts` (from `XGNNPoly.prepare`) the
        lookup is sync-free: max_norm renormalises exactly the rows nn.Embedding would (those present in
        the batch), in place, then gathers; otherwise it is the stock nn.Embedding call."""
        if z_rows is None:
            return F.silu(self.lin(self.embedding(z)))
        emb = self.embedding
        with torch.no_grad():
            sub = emb.weight.index_select(0, z_rows)
            nrm = sub.norm(p=emb.norm_type, dim=1, keepdim=True)
            scale = torch.where(nrm > emb.max_norm, emb.max_norm / (nrm + 1e-7), torch.ones_like(nrm))
            emb.weight.index_copy_(0, z_rows, sub * scale)
        return F.silu(self.lin(_EmbedByFreq.apply(emb.weight, z, z_counts)))


class ResidualLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.lin0, self.lin1 = _lin(c, c), _lin(c, c)

    def forward(self, x):
        return F.silu(self.lin1(F.silu(self.lin0(x)))) + x


class AtomWise(nn.Module):
    def __init__(self, in_channels, rbf_dim, num_target=1, mlp_depth=3):
        super().__init__()
        mods = []
        for _ in range(mlp_depth - 1):
            mods += [_lin(in_channels, in_channels), nn.SiLU()]
        mods.append(_lin(in_channels, num_target))
        self.mlp = nn.ModuleList(mods)
        self.lin_rbf = _lin(rbf_dim, in_channels)

    def forward(self, x, rbf, num_atoms, edge_index_0, atom_rowptr=None):
        """`atom_rowptr` [num_atoms+1] int32 (bonds sorted by first atom): the gated sum runs as one sm_100a
        kernel each way (readout_sum.py); otherwise Linear + product + index_add as in the reference."""
        if atom_rowptr is not None and x.is_cuda and readout_sum.supported(x.size(1), rbf.size(1)):
            out = readout_sum.rbf_readout(x, rbf, self.lin_rbf.weight, self.lin_rbf.bias, atom_rowptr)
        else:
            out = torch.zeros(num_atoms, x.size(1), dtype=x.dtype, device=x.device)
            out.index_add_(0, edge_index_0, self.lin_rbf(rbf) * x)
        for m in self.mlp:
            out = m(out)
        return out


def graph_layer_norm(x, batch, num_graphs, eps=1e-8, counts=None, rowptr=None):
    """PyG 2.1.0 LayerNorm(affine=False) with `batch`: statistics over all rows AND channels of a graph.
    `rowptr` [num_graphs+1] int32 (rows grouped graph after graph, as PyG collates them): one sm_100a kernel
    each way (graph_norm.py).  Otherwise the composite of CUDA PyTorch ops, for any `batch`; `counts`
    [num_graphs] = rows per graph when the caller already has them (torch.bincount synchronises)."""
    if rowptr is not None:
        return graph_layer_norm_rows(x, rowptr, eps)
    cnt = (torch.bincount(batch, minlength=num_graphs) if counts is None else counts).clamp(min=1).to(x.dtype)
    norm = (cnt * x.size(-1)).view(-1, 1)
    acc = torch.zeros(num_graphs, x.size(1), dtype=x.dtype, device=x.device)
    mean = acc.index_add(0, batch, x).sum(-1, keepdim=True) / norm
    x = x - mean.index_select(0, batch)
    var = acc.index_add(0, batch, x * x).sum(-1, keepdim=True) / norm
    return x / (var + eps).sqrt().index_select(0, batch)


class SBFTransformer(nn.Module):
    def __init__(self, conv_layers, emb_size, sbf_dim, rbf_dim=16, in_channels=128, heads=8):
        super().__init__()
        self.edgenn = nn.Sequential(_lin(emb_size, emb_size), nn.SiLU(), _lin(emb_size, emb_size))
        self.convs = nn.ModuleList([
            SBFTransformerConv(in_channels=in_channels, out_channels=in_channels // heads, heads=heads,
                               sbf_dim=sbf_dim * rbf_dim, rbf_dim=rbf_dim, dropout=0, edge_dim=emb_size)
            for _ in range(conv_layers)])
        self.readouts = nn.ModuleList([AtomWise(in_channels, rbf_dim) for _ in range(conv_layers + 1)])
        self.bf_skip = nn.ModuleList([ResidualLayer(in_channels) for _ in range(conv_layers)])
        self.af_skip = nn.ModuleList([nn.Sequential(ResidualLayer(in_channels), ResidualLayer(in_channels))
                                      for _ in range(conv_layers)])
        self.dense_bf_skip = nn.ModuleList([_lin(in_channels, in_channels) for _ in range(conv_layers)])
        self.conv_layers = conv_layers
        self.overlap_readouts = os.environ.get("X2GNN_OVERLAP_READOUTS", "1") != "0"

    def forward(self, x, edge_index, edge_attr, batch, edge_sbf, node_rbf, edge_index_0, atom_batch, num_graphs,
                edge_attr_index=None, edge_attr_target_index=None, batch_counts=None, batch_rowptr=None,
                atom_rowptr=None):
        """`edge_attr` is [T, A] as in the reference; or a per-atom table [N, A] with either
        `edge_attr_index` [T] (rows gathered per triplet AFTER edgenn) or `edge_attr_target_index` [E] (the
        conv layers take the table itself: the row is constant over the triplets of a target bond, so
        lin_edge runs on N rows instead of T -- SURVEY.md §8f row 1)."""
        edge_attr = self.edgenn(edge_attr)
        conv_kw = {}
        if edge_attr_target_index is not None:
            conv_kw["edge_attr_index"] = edge_attr_target_index
        elif edge_attr_index is not None:
            edge_attr = edge_attr[edge_attr_index]
        out = x
        n_atoms = atom_batch.size(0)
        # The readouts hang off the main chain (model.py:41,53: their sum only meets it at the very end) and are
        # atom-scale: ~15 launches of 19-CTA kernels each, forward and backward.  On a side stream they overlap the
        # bond-scale chain instead of interrupting it (autograd runs a node's backward on its forward's stream; under
        # CUDA-graph capture the fork / join become parallel branches of the graph).
        side = self._side_stream(x) if self.overlap_readouts and x.is_cuda else None
        main = torch.cuda.current_stream(x.device) if side is not None else None

        def readout(i, h):
            if side is None:
                return self.readouts[i](h, node_rbf, n_atoms, edge_index_0, atom_rowptr)
            side.wait_stream(main)                        # h (and everything before it) is ready
            with torch.cuda.stream(side):
                r = self.readouts[i](h, node_rbf, n_atoms, edge_index_0, atom_rowptr)
            h.record_stream(side)
            return r

        parts = [readout(0, out)]
        for i in range(self.conv_layers):
            res0 = out
            out = self.convs[i](sbf=edge_sbf, rbf=node_rbf, x=out, edge_index=edge_index, edge_attr=edge_attr,
                                **conv_kw)
            out = graph_layer_norm(out, batch, num_graphs, counts=batch_counts, rowptr=batch_rowptr)
            out = self.bf_skip[i](out)
            out = F.silu(self.dense_bf_skip[i](out)) + res0
            out = self.af_skip[i](out)
            parts.append(readout(i + 1, out))
        if side is not None:
            main.wait_stream(side)
            for r in parts:
                r.record_stream(main)
        results = parts[0]
        for r in parts[1:]:
            results = results + r
        mol = torch.zeros(num_graphs, results.size(1), dtype=results.dtype, device=results.device)
        return mol.index_add(0, atom_batch, results).view(-1)

    @staticmethod
    def _side_stream(x):
        key = x.device.index
        st = _SIDE_STREAMS.get(key)
        if st is None:
            st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=x.device)
        return st


_SIDE_STREAMS: dict = {}       # device index -> side stream of the readouts (module-global: never copied with a model)


class XGNNPoly(nn.Module):
    """state_dict-compatible with the reference's `xgnn_poly` (158 tensors)."""

    def __init__(self, conv_layers=4, sbf_dim=7, rbf_dim=16, in_channels=256, heads=16, embedding_size=128,
                 device="cuda"):
        super().__init__()
        self.emb_block = EmbeddingBlock(embedding_size)
        self.envelop_function = poly_envelop(cutoff=5.0, exponent=5)
        self.sbf_layer = F_B_2D(sbf_dim, rbf_dim, 5.0, 5)
        self.rbf_layer = RadialBasis(cutoff=5.0, embedding_size=rbf_dim)
        self.fin_model = SBFTransformer(conv_layers, embedding_size, sbf_dim, rbf_dim, in_channels, heads)
        self.mat_trans = _lin(338, 2 * embedding_size)
        self.rbf_trans = _lin(rbf_dim, embedding_size)      # declared but unused in the reference too
        self.emb_trans = _lin(2 * embedding_size, in_channels)
        # True: hand the conv layers the per-atom edgenn table + the central atom of every bond (same
        # values as the reference's [T, A] gather, xgnn.py:57-58); False: gather per triplet as the reference
        self.segment_edge_attr = True

    def prepare(self, data: dict) -> dict:
        """The integer part of a forward pass, which depends on the batch only and not on the parameters:
        triplets of the line graph (xgnn.py:52-53), the bond -> molecule map and its histogram
        (model.py:46), the distinct atomic numbers of the batch (nn.Embedding's max_norm / frequency
        bookkeeping) and the CSR metadata the conv layers cache per `edge_index` tensor.  `forward` calls it
        itself; a caller that replays the dense part of the step as a CUDA graph calls it once per batch,
        outside the capture (it reads output sizes back from the device), and passes the result in."""
        from . import graph_meta
        ei, z = data["edge_index"], data["x"]
        B, E, N = int(data["num_graphs"]), int(ei.size(1)), int(z.size(0))
        tri, a_j, a_i, a_k = vertex_to_edge_2(ei, N)
        batch = torch.repeat_interleave(torch.arange(B, device=ei.device), data["edge_num"], output_size=E)
        prep = {"tri": tri, "src_bond": tri[0].contiguous(), "a_j": a_j, "a_i": a_i, "a_k": a_k,
                "batch": batch, "batch_counts": data["edge_num"].to(torch.int64),
                "ei0": ei[0].contiguous(), "ei1": ei[1].contiguous(),
                "z_rows": torch.unique(z), "z_counts": torch.bincount(z, minlength=self.emb_block.embedding.num_embeddings)}
        if data["edge_num"].numel() != B or int(prep["batch_counts"].sum()) != E:
            raise ValueError("edge_num must list the bonds of each of the num_graphs molecules (sum = E)")
        prep["batch_rowptr"] = rowptr_from_counts(prep["batch_counts"])     # bonds are collated graph after graph
        # bonds leaving one atom are contiguous when edge_index is lexicographic (atom_graph.py:42-45): readout CSR
        ei0 = prep["ei0"]
        prep["atom_rowptr"] = (rowptr_from_counts(torch.bincount(ei0, minlength=N))
                               if E == 0 or bool((ei0[1:] >= ei0[:-1]).all()) else None)
        # built here, found in the cache by the 4 layers (same tensor objects); referenced from `prep` so the
        # device buffers outlive the cache's eviction for as long as a captured graph points at them
        from . import sbftransformer_conv as _sc
        prep["line_graph_meta"] = graph_meta.get(tri, E, want_blocks=_sc.USE_FACTORS and _sc.USE_BLOCKS)
        if self.segment_edge_attr:
            prep["edge_attr_groups"] = graph_meta.get_groups(prep["ei1"], N)
        return prep

    def forward(self, data: dict, prep: dict = None):
        """data: x[N] i64, atom_pos[N,3], edge_index[2,E] i64, edge_attr[E,338], edge_num[B], batch[N],
        num_graphs -- the PyG-collated record layout of the reference dataset (qm9_allprop.py:18).
        `prep` = `self.prepare(data)` when the caller has already built the index tensors of this batch."""
        if prep is None:
            prep = self.prepare(data)
        pos = data["atom_pos"]
        B = int(data["num_graphs"])
        tri, a_j, a_i, a_k = prep["tri"], prep["a_j"], prep["a_i"], prep["a_k"]
        d = bond_lengths(pos, prep["ei0"], prep["ei1"])
        env = self.envelop_function(d)[:, None]
        neo_x = F.silu(self.mat_trans(data["edge_attr"] * env))
        # [N, A]; the reference gathers [a_j] here (xgnn.py:58)
        atom_emb = self.emb_block(data["x"], prep["z_rows"], prep["z_counts"])
        ang = triplet_angles(pos, a_i, a_j, a_k)
        edge_sbf = self.sbf_layer(d, ang, prep["src_bond"])
        node_rbf = self.rbf_layer(d) * env
        neo_x = F.silu(self.emb_trans(neo_x))
        kw = dict(batch_counts=prep["batch_counts"], batch_rowptr=prep["batch_rowptr"], atom_rowptr=prep["atom_rowptr"])
        if self.segment_edge_attr:      # a_j[t] == ei[1][tri[1][t]]: the atom shared by both bonds of the triplet
            return self.fin_model(neo_x, tri, atom_emb, prep["batch"], edge_sbf, node_rbf, prep["ei0"],
                                  data["batch"], B, edge_attr_target_index=prep["ei1"], **kw)
        return self.fin_model(neo_x, tri, atom_emb, prep["batch"], edge_sbf, node_rbf, prep["ei0"],
                              data["batch"], B, edge_attr_index=a_j, **kw)

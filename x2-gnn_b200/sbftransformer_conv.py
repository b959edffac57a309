"""SBFTransformerConv on sm_100a.  Drop-in for the reference's sbftransformer_conv.py:16-166:
same constructor, sub-module names (=> state_dict keys), forward signature and return values.

The reference subclasses PyG's MessagePassing and runs ~35 library kernels forward / ~70
backward per call.  Here the layer is one C-ABI call forward and one backward
(x2_sbfconv_fwd / x2_sbfconv_bwd): batched node GEMM, T-row projections, and a deterministic
warp-per-target segmented attention kernel (see csrc/conv.cu).  No torch_geometric /
torch_scatter / torch_sparse dependency, no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import _lib, graph_meta

MODE_FP32 = 0      # X2_MODE_FP32: SIMT fp32 everywhere
MODE_TF32X3 = 1    # X2_MODE_TF32X3: Linear layers on tcgen05 tensor cores, 3xTF32 (fp32-accurate)
MODE_TF32X3_FUSED = 2   # opt-in: + lin_edge / lin_sbf / forward attention as one tcgen05 tile kernel (csrc/tile_attn.cuh)
MODE_TF32 = 3      # X2_MODE_TF32: reduced precision, one tf32 pass per product (the 2e-2 tolerance class)


USE_BLOCKS = os.environ.get("X2GNN_BLOCK", "1") != "0"    # pass the closed blocks of the line graph to the library
# factorised lin_sbf when sbf carries F_B_2D's factors.  Opt-in (X2GNN_SGF=1 or this flag):
# parity-green and more accurate than the dense path, but its kernels are bound by shared-memory reads of the
# per-source tables and measure slower than the dense kernels today (profiles/r2_notes.md)
USE_FACTORS = os.environ.get("X2GNN_SGF", "0") == "1"
PLAN_COUNTS = {"blocks": 0, "factorised": 0}              # layer calls that took each path (tests, bench)


def default_mode(hc: int) -> int:
    """Tensor-core Linear layers (fp32-accurate 3xTF32) whenever the shape allows; override with
    X2GNN_MODE=fp32|tf32x3|tf32x3_fused|tf32 (tf32 = reduced precision, opt-in only)."""
    env = os.environ.get("X2GNN_MODE", "").lower()
    if env == "fp32":
        return MODE_FP32
    if env == "tf32x3":
        return MODE_TF32X3
    if env == "tf32":
        return MODE_TF32 if hc % 128 == 0 else MODE_FP32
    if env == "tf32x3_fused":
        return MODE_TF32X3_FUSED if hc == 128 else (MODE_TF32X3 if hc % 128 == 0 else MODE_FP32)
    return MODE_TF32X3 if hc % 128 == 0 else MODE_FP32


def Glorot_Ortho_(tensor: Tensor, scale: float = 2.0) -> Tensor:
    """Same initialiser as the reference's initializer.Glorot_Ortho_ (orthogonal, then rescaled
    to Glorot variance); restated here so the hot-path module has no out-of-scope import."""
    nn.init.orthogonal_(tensor)
    with torch.no_grad():
        assert tensor.dim() == 2
        tensor.mul_(torch.sqrt(scale / ((tensor.size(0) + tensor.size(1)) * tensor.var())))
    return tensor


_NATIVE_D = (32, 64, 128, 256)     # heads*out_channels the kernels are instantiated for (csrc/conv.cu check_desc)


def _native_c(Dp: int, Cc: int) -> bool:
    """csrc/conv.cu check_desc: a lane owns D/32 consecutive channels and a head is a power-of-two group of lanes."""
    vec = Dp // 32
    return Cc % vec == 0 and ((Cc // vec) & (Cc // vec - 1)) == 0 and Cc // vec <= 32 and Dp % Cc == 0


def padded_width(in_ch: int, H: int, Cc: int):
    """Width D' >= max(in_channels, heads*out_channels) the kernels run at (None: no such width).  The reference's
    constructor (sbftransformer_conv.py:19,47-48) takes any in_channels / heads / out_channels; the kernels take
    x [E, D'] with D' = H'*C in {32, 64, 128, 256}.  Other shapes are embedded EXACTLY: x and the weights are
    zero-padded (extra input channels contribute 0 to every projection; extra heads have Q = K = V = EA = 0, so their
    logits are 0, their messages 0 and their output columns 0) and the first heads*out_channels columns of the
    result are returned.  The padding and the slice are differentiable torch ops, so every gradient is that of the
    unpadded layer."""
    for Dp in _NATIVE_D:
        if Dp >= max(in_ch, H * Cc) and _native_c(Dp, Cc):
            return Dp
    return None


def _pad2(w, rows: int, cols: int):
    return w if w is None or (w.size(0) == rows and w.size(1) == cols) else F.pad(w, (0, cols - w.size(1), 0, rows - w.size(0)))


def _pad1(b, n: int):
    return b if b is None or b.size(0) == n else F.pad(b, (0, n - b.size(0)))


def _set_groups(desc, groups):
    if groups is None:
        desc.ea_rows, desc.ea_index, desc.ea_rowptr, desc.ea_order = 0, None, None, None
    else:
        desc.ea_rows = groups.rows
        desc.ea_index, desc.ea_rowptr, desc.ea_order = (_lib.ptr(groups.index), _lib.ptr(groups.rowptr),
                                                         _lib.ptr(groups.order))


def _fill_desc(cfg, meta, t, dims, groups):
    """x2_conv_desc of one layer call (the same descriptor drives the forward and the backward)."""
    E, T, D, H, Cc, S, R, A, fuse = dims
    desc = _lib.ConvDesc()
    desc.E, desc.T = E, T
    desc.D, desc.H, desc.C, desc.S, desc.R, desc.A = D, H, Cc, S, R, A
    desc.fuse_skip, desc.mode = fuse, cfg["mode"]
    desc.dropout_p, desc.seed = cfg["dropout_p"], cfg["seed"]
    desc.tgt_sorted = 1 if meta.target_sorted else 0
    for n in ("x", "rbf", "sbf", "w_rbf", "w_q", "b_q", "w_k", "b_k", "w_v", "b_v", "w_sbf", "b_sbf",
              "w_skip", "b_skip"):
        setattr(desc, n, _lib.ptr(t[n]))
    desc.edge_attr = _lib.ptr(t["edge_attr"]) if A else None
    desc.w_edge = _lib.ptr(t["w_edge"]) if A else None
    for n in ("src", "tgt", "rowptr_tgt", "order_tgt", "rowptr_src", "order_src"):
        setattr(desc, n, _lib.ptr(getattr(meta, n)))
    _set_groups(desc, groups)
    desc.items, desc.itemptr, desc.items_bound = _lib.ptr(meta.items), _lib.ptr(meta.itemptr), meta.items_bound
    blk = meta.blocks
    if blk is not None and USE_BLOCKS:         # closed blocks of the line graph: block-centric backward (csrc/blk_attn.cuh)
        desc.nblk, desc.blk_max_src, desc.blk_max_trip, desc.blk_max_tgt = blk.n, blk.max_src, blk.max_triplets, blk.max_tgt
        desc.blk_sptr, desc.blk_tptr, desc.blk_tord, desc.blk_tpos = (_lib.ptr(blk.sptr), _lib.ptr(blk.tptr),
                                                                      _lib.ptr(blk.tord), _lib.ptr(blk.tpos))
    fac = cfg.get("sbf_factors")
    if fac is not None and USE_FACTORS:         # factorised sbf (F_B_2D output): lin_sbf inside the attention kernels
        desc.sbf_tab, desc.angles = _lib.ptr(fac.table), _lib.ptr(fac.angles)
        desc.sbf_L, desc.sbf_R = fac.L, fac.R
    return desc


class _SBFConvFn(torch.autograd.Function):
    """out/attn = conv(x, rbf, sbf, edge_attr; weights).  Non-tensor config rides in `cfg`."""

    @staticmethod
    def forward(ctx, cfg, meta, x, rbf, sbf, edge_attr, w_rbf, w_q, b_q, w_k, b_k, w_v, b_v, w_edge,
                w_sbf, b_sbf, w_skip, b_skip):
        names = ("x", "rbf", "sbf", "edge_attr", "w_rbf", "w_q", "b_q", "w_k", "b_k", "w_v", "b_v",
                 "w_edge", "w_sbf", "b_sbf", "w_skip", "b_skip")
        vals = [_lib.f32(t, f"SBFTransformerConv.{n}") for n, t in zip(names, (
            x, rbf, sbf, edge_attr, w_rbf, w_q, b_q, w_k, b_k, w_v, b_v, w_edge, w_sbf, b_sbf, w_skip, b_skip))]
        t = dict(zip(names, vals))
        dev = _lib.require_cuda(*vals, meta.src, what="SBFTransformerConv")
        E, D = t["x"].shape
        T = meta.T
        H, Cc = cfg["heads"], cfg["out_channels"]
        S, R = t["sbf"].size(1), t["rbf"].size(1)
        A = t["edge_attr"].size(1) if t["w_edge"] is not None else 0
        if meta.E != E:
            raise ValueError(f"edge_index was indexed for {meta.E} nodes but x has {E} rows")
        # the kernels index the weights with D, S, R, A taken from the inputs: a basis of the wrong width
        # must raise here (the reference fails inside F.linear), not read out of bounds on the device
        HC = H * Cc
        want = {"w_rbf": (D, R), "w_sbf": (HC, S), "b_sbf": (HC,), "w_q": (HC, D), "b_q": (HC,), "w_k": (HC, D),
                "b_k": (HC,), "w_v": (HC, D), "b_v": (HC,), "w_edge": (HC, A), "w_skip": (HC, D), "b_skip": (HC,)}
        for n, shp in want.items():
            if t[n] is not None and tuple(t[n].shape) != shp:
                raise ValueError(f"SBFTransformerConv: {n} has shape {tuple(t[n].shape)}, the inputs need {shp} "
                                 f"(x [{E},{D}], rbf [.,{R}], sbf [.,{S}], edge_attr [.,{A}], heads*out_channels {HC})")
        if D != HC:
            raise ValueError(f"SBFTransformerConv: x has {D} channels, heads*out_channels = {HC}")
        groups = cfg.get("ea_groups") if A else None     # segment-constant edge_attr table (opt-in)
        n_ea = groups.rows if groups is not None else T
        if t["sbf"].size(0) != T or (A and t["edge_attr"].size(0) != n_ea) or t["rbf"].size(0) != E:
            raise ValueError("SBFTransformerConv: sbf/edge_attr must have one row per edge_index "
                             "column (edge_attr: one row per table row with edge_attr_index) and rbf one "
                             "row per node")
        if groups is not None and groups.index.numel() != E:
            raise ValueError(f"edge_attr_index must have one entry per node of the line graph ({E}), "
                             f"got {groups.index.numel()}")
        fuse = 1 if t["w_skip"] is not None else 0

        dims = (E, T, D, H, Cc, S, R, A, fuse)
        desc = _fill_desc(cfg, meta, t, dims, groups)
        L = _lib.lib()
        plan = L.x2_sbfconv_plan(C.byref(desc))
        factorised = bool(plan & 2) and not cfg["want_alpha"]
        PLAN_COUNTS["blocks"] += plan & 1
        PLAN_COUNTS["factorised"] += int(factorised)

        f32 = dict(dtype=torch.float32, device=dev)
        qkvs = torch.empty((E, 4 * D), **f32)
        attn = torch.empty((E, D), **f32)
        lse = torch.empty((E, H), **f32)
        ea = torch.empty((max(n_ea, 1), D), **f32) if A else None
        sg = None if factorised else torch.empty((max(T, 1), D), **f32)   # factorised sbf: lin_sbf(sbf) is never stored
        xs = torch.empty((E, D), **f32)          # x * lin_rbf(rbf): kept so the backward does not recompute it
        # (not the saved `attn` buffer itself when there is no fused skip add: a caller's in-place op on the layer
        # output would silently corrupt the backward, which reads attn to form r = <G, O>)
        out = torch.empty((E, D), **f32)
        alpha = torch.empty((T, H), **f32) if cfg["want_alpha"] else None
        saved = _lib.ConvSaved(_lib.ptr(qkvs), _lib.ptr(attn), _lib.ptr(lse), _lib.ptr(ea), _lib.ptr(sg),
                               _lib.ptr(xs))
        ws = _lib.workspace(L.x2_sbfconv_fwd_workspace_bytes(C.byref(desc)), dev)
        _lib.check(L.x2_sbfconv_fwd(C.byref(desc), C.byref(saved), _lib.ptr(out), _lib.ptr(alpha),
                                    _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_sbfconv_fwd")
        ctx.cfg, ctx.meta, ctx.dims = cfg, meta, dims
        ctx.groups = groups
        # inputs, weights and the forward's buffers go through save_for_backward: an in-place modification of any of
        # them between forward and backward is then an autograd error instead of a silently wrong gradient
        ctx.names = names
        ctx.save_for_backward(*[t[n] for n in names], qkvs, attn, lse, ea, sg, xs)
        if alpha is not None:
            ctx.mark_non_differentiable(alpha)
            return out, alpha
        return out, None

    @staticmethod
    def backward(ctx, gout, _galpha):
        saved = ctx.saved_tensors
        t, meta = dict(zip(ctx.names, saved[:len(ctx.names)])), ctx.meta
        E, T, D, H, Cc, S, R, A, fuse = ctx.dims
        qkvs, attn, lse, ea, sg, xs = saved[len(ctx.names):]
        gout = _lib.f32(gout, "SBFTransformerConv.backward")
        dev = gout.device
        cfg = ctx.cfg

        desc = _fill_desc(cfg, meta, t, ctx.dims, ctx.groups)
        n_ea = ctx.groups.rows if ctx.groups is not None else T
        saved = _lib.ConvSaved(_lib.ptr(qkvs), _lib.ptr(attn), _lib.ptr(lse), _lib.ptr(ea), _lib.ptr(sg),
                               _lib.ptr(xs))

        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad     # (cfg, meta, x, rbf, sbf, edge_attr, ...)
        g = {
            "dx": torch.empty((E, D), **f32), "drbf": torch.empty((E, R), **f32),
            "dsbf": torch.empty((T, S), **f32) if need[4] else None,
            "dedge_attr": torch.empty((n_ea, A), **f32) if (A and need[5]) else None,
            "dw_rbf": torch.empty((D, R), **f32),
            "dw_q": torch.empty((D, D), **f32), "db_q": torch.empty(D, **f32),
            "dw_k": torch.empty((D, D), **f32), "db_k": torch.empty(D, **f32),
            "dw_v": torch.empty((D, D), **f32), "db_v": torch.empty(D, **f32),
            "dw_edge": torch.empty((D, A), **f32) if A else None,
            "dw_sbf": torch.empty((D, S), **f32), "db_sbf": torch.empty(D, **f32),
            "dw_skip": torch.empty((D, D), **f32) if fuse else None,
            "db_skip": torch.empty(D, **f32) if (fuse and t["b_skip"] is not None) else None,
        }
        grads = _lib.ConvGrads(*[_lib.ptr(g[n]) for n, _ in _lib.ConvGrads._fields_])
        L = _lib.lib()
        ws = _lib.workspace(L.x2_sbfconv_bwd_workspace_bytes(C.byref(desc)), dev)
        _lib.check(L.x2_sbfconv_bwd(C.byref(desc), C.byref(saved), _lib.ptr(gout), C.byref(grads),
                                    _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_sbfconv_bwd")
        return (None, None, g["dx"], g["drbf"], g["dsbf"], g["dedge_attr"], g["dw_rbf"], g["dw_q"],
                g["db_q"], g["dw_k"], g["db_k"], g["dw_v"], g["db_v"], g["dw_edge"], g["dw_sbf"],
                g["db_sbf"], g["dw_skip"], g["db_skip"])


class SBFTransformerConv(nn.Module):
    """Multi-head attention message passing on the line graph with an rbf source filter, sbf value
    gate, additive edge features and root skip.  See the reference for the model semantics."""

    def __init__(self, in_channels, out_channels: int, heads: int = 1, sbf_dim: int = 16,
                 rbf_dim: int = 16, concat: bool = True, beta: bool = False, dropout: float = 0.,
                 edge_dim: Optional[int] = None, bias: bool = True, root_weight: bool = True, **kwargs):
        aggr = kwargs.pop("aggr", "add")
        if aggr != "add":
            raise ValueError("SBFTransformerConv: only aggr='add' is supported")
        kwargs.pop("node_dim", None)
        kwargs.pop("flow", None)
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        self.sbf_dim = sbf_dim
        self.rbf_dim = rbf_dim
        self.beta = beta and root_weight
        self.root_weight = root_weight
        self.concat = concat
        self.dropout = dropout
        self.edge_dim = edge_dim
        self._alpha = None
        self.precision = None          # None => default_mode(heads*out_channels) at call time

        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        hc = heads * out_channels
        self.lin_key = nn.Linear(in_channels[0], hc)
        self.lin_query = nn.Linear(in_channels[1], hc)
        self.lin_value = nn.Linear(in_channels[0], hc)
        if edge_dim is not None:
            self.lin_edge = nn.Linear(edge_dim, hc, bias=False)
        else:
            self.lin_edge = self.register_parameter("lin_edge", None)
        if concat:
            self.lin_skip = nn.Linear(in_channels[1], hc, bias=bias)
            if self.beta:
                self.lin_beta = nn.Linear(3 * hc, 1, bias=False)
            else:
                self.lin_beta = self.register_parameter("lin_beta", None)
        else:
            self.lin_skip = nn.Linear(in_channels[1], out_channels, bias=bias)
            if self.beta:
                self.lin_beta = nn.Linear(3 * out_channels, 1, bias=False)
            else:
                self.lin_beta = self.register_parameter("lin_beta", None)
        self.lin_sbf = nn.Linear(sbf_dim, hc, bias=True)
        self.lin_rbf = nn.Linear(rbf_dim, in_channels[0], bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            Glorot_Ortho_(self.lin_sbf.weight)
            Glorot_Ortho_(self.lin_rbf.weight)
            nn.init.zeros_(self.lin_sbf.bias)     # lin_rbf has no bias
        self.lin_key.reset_parameters()
        self.lin_query.reset_parameters()
        self.lin_value.reset_parameters()
        if self.edge_dim:
            self.lin_edge.reset_parameters()
        self.lin_skip.reset_parameters()
        if self.beta:
            self.lin_beta.reset_parameters()

    def forward(self, sbf, rbf, x, edge_index, edge_attr=None, return_attention_weights=None,
                edge_attr_index=None):
        """Reference signature (sbftransformer_conv.py:93-94) plus one opt-in keyword (SURVEY.md §8f row
        1): with `edge_attr_index` [E] (integer tensor), `edge_attr` is a TABLE [M, edge_dim] and every
        edge_index column whose target is line-node e uses row edge_attr_index[e] -- identical in value to
        passing edge_attr[edge_attr_index[edge_index[1]]] ([T, edge_dim]), without streaming T rows
        through lin_edge.  (xgnn.py:57-58 builds edge_attr from the central atom only, so there
        edge_attr = edgenn(atom_embeddings), edge_attr_index = the bond's second atom.)"""
        if not isinstance(x, Tensor):
            raise TypeError("SBFTransformerConv: `x` must be a Tensor (the reference's tuple path is dead code)")
        if not isinstance(edge_index, Tensor):
            raise TypeError("SBFTransformerConv: edge_index must be a [2, T] LongTensor")
        if self.lin_edge is not None and edge_attr is None:
            raise AssertionError("edge_attr is required when edge_dim is set")
        H, Cc = self.heads, self.out_channels
        in_ch = self.in_channels if isinstance(self.in_channels, int) else self.in_channels[0]
        if not isinstance(self.in_channels, int) and self.in_channels[0] != self.in_channels[1]:
            # (x_src = x * lin_rbf(rbf) and lin_query(x) take the same Tensor x in the reference, :98-107)
            raise ValueError("SBFTransformerConv: a Tensor `x` needs in_channels[0] == in_channels[1]")
        if x.dim() != 2 or x.size(1) != in_ch:
            raise ValueError(f"SBFTransformerConv: x must be [E, {in_ch}], got {tuple(x.shape)}")
        Dp = padded_width(in_ch, H, Cc)
        if Dp is None:
            raise NotImplementedError(
                f"SBFTransformerConv: no kernel width for in_channels={in_ch}, heads={H}, out_channels={Cc} "
                "(need max(in_channels, heads*out_channels) <= 256 and out_channels a power of two; INTEGRATION.md)")
        Hk = Dp // Cc                              # heads the kernels see (H, or H + zero heads when padded)
        padded = Dp != in_ch or Dp != H * Cc
        # (the closed blocks of the line graph are only built when the factorised lin_sbf can use them)
        meta = graph_meta.get(edge_index, x.size(0),
                              want_blocks=USE_FACTORS and USE_BLOCKS and getattr(sbf, "_x2_factors", None) is not None)
        fuse = self.concat and self.root_weight and self.lin_beta is None
        p_drop = float(self.dropout) if self.training else 0.0
        if p_drop > 0 and x.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("SBFTransformerConv: attention dropout inside a CUDA graph capture would replay one "
                               "frozen mask (the seed is drawn on the host); capture with dropout = 0")
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p_drop > 0 else 0
        mode = default_mode(Dp) if self.precision is None else self.precision
        cfg = dict(heads=Hk, out_channels=Cc, mode=mode, dropout_p=p_drop, seed=seed,
                   want_alpha=isinstance(return_attention_weights, bool))
        if edge_attr_index is not None and self.lin_edge is not None:
            cfg["ea_groups"] = graph_meta.get_groups(edge_attr_index, edge_attr.size(0))
        # F_B_2D tags its output with the factors it multiplied out (angular_basis_layer.SbfFactors): when they
        # still describe `sbf` and index the source line-nodes of `edge_index`, lin_sbf is evaluated inside the
        # attention kernels and the [T, S] tensor is not read
        fac = getattr(sbf, "_x2_factors", None)
        if fac is not None and (p_drop > 0 or cfg["want_alpha"] or (torch.is_grad_enabled() and sbf.requires_grad)
                                or not fac.describes(sbf, edge_index)):
            fac = None
        cfg["sbf_factors"] = fac
        w_e = self.lin_edge.weight if self.lin_edge is not None else None
        w_o, b_o = (self.lin_skip.weight, self.lin_skip.bias) if fuse else (None, None)
        if not padded:
            out, alpha = _SBFConvFn.apply(
                cfg, meta, x, rbf, sbf, edge_attr if self.lin_edge is not None else None,
                self.lin_rbf.weight, self.lin_query.weight, self.lin_query.bias, self.lin_key.weight,
                self.lin_key.bias, self.lin_value.weight, self.lin_value.bias, w_e, self.lin_sbf.weight,
                self.lin_sbf.bias, w_o, b_o)
        else:      # exact zero-padded embedding into the kernels' width (padded_width above)
            out, alpha = _SBFConvFn.apply(
                cfg, meta, F.pad(x, (0, Dp - in_ch)), rbf, sbf, edge_attr if self.lin_edge is not None else None,
                _pad2(self.lin_rbf.weight, Dp, self.lin_rbf.weight.size(1)), _pad2(self.lin_query.weight, Dp, Dp),
                _pad1(self.lin_query.bias, Dp), _pad2(self.lin_key.weight, Dp, Dp), _pad1(self.lin_key.bias, Dp),
                _pad2(self.lin_value.weight, Dp, Dp), _pad1(self.lin_value.bias, Dp),
                _pad2(w_e, Dp, w_e.size(1)) if w_e is not None else None,
                _pad2(self.lin_sbf.weight, Dp, self.lin_sbf.weight.size(1)), _pad1(self.lin_sbf.bias, Dp),
                _pad2(w_o, Dp, Dp), _pad1(b_o, Dp))
            out = out[:, :H * Cc]
            if alpha is not None:
                alpha = alpha[:, :H]
        if not fuse:
            # composite tail for the non-default variants (CUDA torch ops, still no CPU path)
            out = out.reshape(-1, H * Cc) if self.concat else out.reshape(-1, H, Cc).mean(dim=1)
            if self.root_weight:
                x_r = self.lin_skip(x)
                if self.lin_beta is not None:
                    beta = self.lin_beta(torch.cat([out, x_r, out - x_r], dim=-1)).sigmoid()
                    out = beta * x_r + (1 - beta) * out
                else:
                    out = out + x_r
        if isinstance(return_attention_weights, bool):
            assert alpha is not None
            return out, (edge_index, alpha)
        return out

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"

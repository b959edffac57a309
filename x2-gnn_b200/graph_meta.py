"""CSR metadata of the line graph consumed by the attention kernels.

PyG derives this implicitly inside propagate/softmax/scatter on every layer call
(sbftransformer_conv.py:109,151).  Here it is built once per `edge_index` tensor by integer
kernels (x2_meta_build) and reused by all conv layers and their backward passes.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class Blocks:
    """Closed blocks of the line graph (x2_blocks_build): block b = sources [sptr[b], sptr[b+1]) + the targets
    tord[tptr[b] : tptr[b+1]]; every triplet has its source and its target in one block."""
    n: int
    sptr: torch.Tensor         # [E+1] int32
    tptr: torch.Tensor         # [E+1] int32
    tord: torch.Tensor         # [E] int32
    tpos: torch.Tensor         # [E] int32: position of a target inside its block
    max_triplets: int
    max_src: int
    max_tgt: int


@dataclass
class LineGraphMeta:
    T: int
    E: int
    src: torch.Tensor          # [T] int32
    tgt: torch.Tensor          # [T] int32
    rowptr_tgt: torch.Tensor   # [E+1] int32
    order_tgt: torch.Tensor    # [T] int32 (identity when target-sorted)
    rowptr_src: torch.Tensor   # [E+1] int32
    order_src: torch.Tensor    # [T] int32
    target_sorted: bool
    max_seg: int = 0           # longest target segment
    items: torch.Tensor = None     # [items_bound, 2] int32 work items of the fused kernels (x2_items_build), or None
    itemptr: torch.Tensor = None   # [E+1] int32
    items_bound: int = 0
    blocks: Blocks = None          # None: no closed-block structure (or not target-sorted): generic kernels only
    blocks_tried: bool = False     # False: the blocks were not asked for when this record was built


import os

# The closed blocks (x2_blocks_build) only serve the factorised lin_sbf kernels, which are opt-in (X2GNN_SGF=1 /
# sbftransformer_conv.USE_FACTORS): they are built -- eight small launches -- only when asked for.
WANT_BLOCKS = os.environ.get("X2GNN_SGF", "0") == "1"


def build(edge_index: torch.Tensor, num_nodes: int, want_blocks: bool = None) -> LineGraphMeta:
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must be [2, T], got {tuple(edge_index.shape)}")
    if edge_index.dtype != torch.int64:
        edge_index = edge_index.long()
    dev = _lib.require_cuda(edge_index, what="line-graph metadata")
    ei = edge_index.contiguous()
    T, E = int(ei.size(1)), int(num_nodes)
    L = _lib.lib()
    i32 = dict(dtype=torch.int32, device=dev)
    src = torch.empty(max(T, 1), **i32)
    tgt = torch.empty(max(T, 1), **i32)
    rp_t = torch.empty(E + 1, **i32)
    rp_s = torch.empty(E + 1, **i32)
    od_t = torch.empty(max(T, 1), **i32)
    od_s = torch.empty(max(T, 1), **i32)
    flags = torch.zeros(4, **i32)
    ws = _lib.workspace(L.x2_meta_workspace_bytes(T, E), dev)
    _lib.check(L.x2_meta_build(_lib.ptr(ei), T, E, _lib.ptr(src), _lib.ptr(tgt), _lib.ptr(rp_t),
                               _lib.ptr(od_t), _lib.ptr(rp_s), _lib.ptr(od_s), _lib.ptr(flags),
                               _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_meta_build")
    # closed blocks (meaningful for a target-sorted list only; queued before the one host sync below)
    bflags = torch.zeros(6, **i32)
    if want_blocks is None:
        want_blocks = WANT_BLOCKS
    if want_blocks and T > 0 and E > 0:
        sptr = torch.empty(E + 1, **i32)
        tptr = torch.empty(E + 1, **i32)
        tord = torch.empty(E, **i32)
        tpos = torch.empty(E, **i32)
        ws3 = _lib.workspace(L.x2_blocks_workspace_bytes(T, E), dev)
        _lib.check(L.x2_blocks_build(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(rp_t), T, E, _lib.ptr(sptr), _lib.ptr(tptr),
                                     _lib.ptr(tord), _lib.ptr(tpos), _lib.ptr(bflags), _lib.ptr(ws3), ws3.numel(),
                                     _lib.stream()),
                   "x2_blocks_build")
    f = torch.cat([flags, bflags]).tolist()          # one host sync per batch; also surfaces async kernel errors
    if f[1] != 0:
        raise IndexError(f"edge_index has {f[1]} entries outside [0, {E})")
    meta = LineGraphMeta(T, E, src, tgt, rp_t, od_t, rp_s, od_s, bool(f[0]), int(f[2]))
    if meta.target_sorted and T > 0 and 0 < E < (1 << 27):
        # work items (<= 8 rows of one segment each) for the fused tcgen05 kernels
        nb = int(L.x2_items_bound(T, E))
        meta.items = torch.empty((nb, 2), **i32)
        meta.itemptr = torch.empty(E + 1, **i32)
        meta.items_bound = nb
        ws2 = _lib.workspace(L.x2_items_workspace_bytes(E), dev)
        _lib.check(L.x2_items_build(_lib.ptr(rp_t), E, T, _lib.ptr(meta.itemptr), _lib.ptr(meta.items),
                                    _lib.ptr(ws2), ws2.numel(), _lib.stream()), "x2_items_build")
    meta.blocks_tried = bool(want_blocks)
    if want_blocks and meta.target_sorted and T > 0 and E > 0:
        ok, nb, mt, ms, mg = f[4:9]
        # a block is one CTA's work: keep the generic kernels when a few blocks hold most of the graph
        if ok and nb > 0 and mt <= MAX_BLOCK_TRIPLETS:
            meta.blocks = Blocks(nb, sptr, tptr, tord, tpos, mt, ms, mg)
    return meta


MAX_BLOCK_TRIPLETS = 1 << 16

_cache: list = []   # [(weakref(edge_index), version, num_nodes, meta)], most recent first
_CACHE_SIZE = 8


def get(edge_index: torch.Tensor, num_nodes: int, want_blocks: bool = False) -> LineGraphMeta:
    """Cached `build`: the same tensor object (unchanged `_version`) is passed to every conv
    layer of a forward pass (model.py:45), so the metadata is built once per batch."""
    ver = edge_index._version
    for i, (ref, v, n, meta) in enumerate(_cache):
        if ref() is edge_index and v == ver and n == num_nodes and (meta.blocks_tried or not want_blocks):
            if i:
                _cache.insert(0, _cache.pop(i))
            return meta
    meta = build(edge_index, num_nodes, want_blocks=want_blocks or WANT_BLOCKS)
    _cache.insert(0, (weakref.ref(edge_index), ver, num_nodes, meta))
    del _cache[_CACHE_SIZE:]
    return meta


@dataclass
class RowGroups:
    """Targets grouped by the `edge_attr` table row they use (segment-constant edge features)."""
    rows: int
    index: torch.Tensor        # [E] int32: table row of every target line-node
    rowptr: torch.Tensor       # [rows+1] int32
    order: torch.Tensor        # [E] int32: target ids grouped by table row, ascending inside a group


_group_cache: list = []


def get_groups(index: torch.Tensor, rows: int) -> RowGroups:
    """Cached grouping of `index` [E] (values in [0, rows)) with the same integer kernels that build the
    line-graph metadata (x2_meta_build on the [2, E] index [index; index] with `rows` nodes)."""
    if index.dim() != 1:
        raise ValueError(f"edge_attr_index must be one-dimensional, got {tuple(index.shape)}")
    ver = index._version
    for i, (ref, v, n, grp) in enumerate(_group_cache):
        if ref() is index and v == ver and n == rows:
            if i:
                _group_cache.insert(0, _group_cache.pop(i))
            return grp
    idx = index.long()
    meta = build(torch.stack([idx, idx]), rows)       # raises IndexError on out-of-range rows
    grp = RowGroups(rows, meta.src, meta.rowptr_src, meta.order_src)
    _group_cache.insert(0, (weakref.ref(index), ver, rows, grp))
    del _group_cache[_CACHE_SIZE:]
    return grp


def clear_cache():
    _cache.clear()
    _group_cache.clear()

"""Device-resident dataset + batch collation on the GPU.

The reference's loader (PyG DataLoader over qm9_allprop.py records) concatenates the molecules of a batch in
Python on the host -- atoms, bonds, the [E, 338] pair features -- offsets `edge_index` per graph, builds `batch`
and copies ~58 MB to the device every step.  QM9 with its pair features is ~55 GB: it fits the B200's 180 GB, so
here the whole dataset is uploaded once in CSR form and a batch is two launches (`x2_collate_sizes`,
`x2_collate_fill`) and one size read-back.

    ds = DeviceDataset.from_molecules(mols, device)        # mols: dicts with x[n], atom_pos[n,3], edge_index[2,e] (local), edge_attr[e,F], y
    data = ds.collate(ids)                                  # the record layout XGNNPoly.forward / xgnn_poly.forward take
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class DeviceDataset:
    def __init__(self, z, pos, atom_ptr, feat, edge_ptr, edge_index=None, y=None):
        self.z, self.pos, self.atom_ptr, self.feat, self.edge_ptr = z, pos, atom_ptr, feat, edge_ptr
        self.edge_index, self.y = edge_index, y
        self.M = int(atom_ptr.numel()) - 1
        self.F = int(feat.size(1))

    @classmethod
    def from_molecules(cls, mols, device):
        """mols: sequence of dicts {x [n] int, atom_pos [n,3], edge_index [2,e] local ids, edge_attr [e,F], y}."""
        an = np.cumsum([0] + [len(m["x"]) for m in mols]).astype(np.int64)
        en = np.cumsum([0] + [m["edge_attr"].shape[0] for m in mols]).astype(np.int64)
        cat = lambda k, dt: torch.from_numpy(np.concatenate([np.asarray(m[k]) for m in mols], axis=0).astype(dt))
        ei = torch.from_numpy(np.concatenate([np.asarray(m["edge_index"]) for m in mols], axis=1).astype(np.int64))
        y = torch.tensor([float(m.get("y", 0.0)) for m in mols], dtype=torch.float32)
        dev = torch.device(device)
        return cls(cat("x", np.int64).to(dev), cat("atom_pos", np.float32).to(dev).contiguous(),
                   torch.from_numpy(an).to(dev), cat("edge_attr", np.float32).to(dev).contiguous(),
                   torch.from_numpy(en).to(dev), ei.to(dev).contiguous(), y.to(dev))

    def collate(self, ids, with_edge_index: bool = True) -> dict:
        """ids: molecule ids of the batch (sequence or int64 tensor).  Returns the collated record on the device."""
        ids = torch.as_tensor(ids, dtype=torch.int64, device=self.z.device).contiguous()
        dev = _lib.require_cuda(ids, self.z, what="DeviceDataset.collate")
        B = int(ids.numel())
        L = _lib.lib()
        i32 = dict(dtype=torch.int32, device=dev)
        aoff, eoff, flags = torch.empty(B + 1, **i32), torch.empty(B + 1, **i32), torch.zeros(4, **i32)
        ws = _lib.workspace(L.x2_collate_workspace_bytes(B), dev)
        _lib.check(L.x2_collate_sizes(_lib.ptr(ids), B, _lib.ptr(self.atom_ptr), _lib.ptr(self.edge_ptr), self.M,
                                      _lib.ptr(aoff), _lib.ptr(eoff), _lib.ptr(flags), _lib.ptr(ws), ws.numel(),
                                      _lib.stream()), "x2_collate_sizes")
        N, E, bad = int(aoff[B]), int(eoff[B]), int(flags[0])            # the one host sync (output sizes)
        if bad:
            raise IndexError(f"DeviceDataset.collate: {bad} molecule ids outside [0, {self.M})")
        i64 = dict(dtype=torch.int64, device=dev)
        z, batch, edge_num = torch.empty(N, **i64), torch.empty(N, **i64), torch.empty(B, **i64)
        pos = torch.empty((N, 3), dtype=torch.float32, device=dev)
        feat = torch.empty((E, self.F), dtype=torch.float32, device=dev)
        use_ei = with_edge_index and self.edge_index is not None
        ei = torch.empty((2, E), **i64) if use_ei else None
        _lib.check(L.x2_collate_fill(_lib.ptr(ids), B, self.M, _lib.ptr(self.atom_ptr), _lib.ptr(self.edge_ptr),
                                     _lib.ptr(self.z), _lib.ptr(self.pos), _lib.ptr(self.feat), self.F,
                                     _lib.ptr(self.edge_index) if use_ei else None,
                                     int(self.edge_index.size(1)) if use_ei else 0, _lib.ptr(aoff), _lib.ptr(eoff),
                                     _lib.ptr(z), _lib.ptr(pos), _lib.ptr(batch), _lib.ptr(edge_num), _lib.ptr(feat),
                                     _lib.ptr(ei), E, _lib.stream()), "x2_collate_fill")
        out = {"x": z, "atom_pos": pos, "batch": batch, "edge_num": edge_num, "edge_attr": feat, "num_graphs": B}
        if use_ei:
            out["edge_index"] = ei
        if self.y is not None:
            out["y"] = self.y.index_select(0, ids)
        return out

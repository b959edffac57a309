"""torch_geometric.data subset: `Data` is an attribute bag whose dict is reachable as `_store` (xgnn.py:41-42
tests `'batch' in data._store`), `Batch.from_data_list` collates the way PyG does for the record layout of
qm9_allprop.py:18-19: tensors concatenated on dim 0 -- except keys containing "index" (dim -1, offset by
the graph's node count) -- scalars stacked, plus `batch`, `ptr`, `num_graphs`."""
import torch


class Data:
    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", dict(kwargs))

    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        try:
            return store[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    def keys(self):
        return list(self._store.keys())

    @property
    def num_nodes(self):
        for k in ("x", "atom_pos", "pos"):
            if k in self._store and torch.is_tensor(self._store[k]):
                return int(self._store[k].size(0))
        return None

    def to(self, device, *a, **kw):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device, *a, **kw)
        return self

    def __repr__(self):
        body = ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else v!r}" for k, v in self._store.items())
        return f"{type(self).__name__}({body})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        out, offs, batch, ptr = {}, 0, [], [0]
        keys = data_list[0].keys()
        cols = {k: [] for k in keys}
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = getattr(d, k)
                if torch.is_tensor(v) and "index" in k:
                    v = v + offs
                cols[k].append(v)
            batch.append(torch.full((n,), g, dtype=torch.long))
            offs += n
            ptr.append(offs)
        for k, vs in cols.items():
            v0 = vs[0]
            if torch.is_tensor(v0):
                if v0.dim() == 0:
                    out[k] = torch.stack(vs)
                else:
                    out[k] = torch.cat(vs, dim=-1 if "index" in k else 0)
            elif isinstance(v0, (int, float)):
                out[k] = torch.tensor(vs)
            else:
                out[k] = vs
        b = cls(**out)
        b.batch = torch.cat(batch)
        b.ptr = torch.tensor(ptr)
        b.num_graphs = len(data_list)
        return b


class InMemoryDataset:   # dataset construction (pyscf features) is out of scope; the name keeps imports working
    def __init__(self, *a, **kw):
        raise NotImplementedError("x2gnn_b200.compat: InMemoryDataset is a placeholder")

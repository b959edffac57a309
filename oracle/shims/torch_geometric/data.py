"""PyG Data stand-in: attribute bag backed by a dict named `_store` (xgnn.py:41-42
probes `"batch" in data._store`)."""


class Data:
    def __init__(self, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in kwargs.items():
            self._store[k] = v

    def __getattr__(self, key):
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def to(self, device):
        for k, v in list(self._store.items()):
            if hasattr(v, 'to'):
                self._store[k] = v.to(device)
        return self


class InMemoryDataset:  # pragma: no cover - dataset code is out of scope
    pass

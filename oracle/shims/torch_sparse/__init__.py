"""TEST INFRASTRUCTURE ONLY (oracle/): the reference only type-checks against
SparseTensor (sbftransformer_conv.py:7,133)."""


class SparseTensor:  # pragma: no cover - never instantiated
    pass

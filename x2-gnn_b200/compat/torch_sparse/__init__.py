"""torch_sparse: only the name `SparseTensor` is needed (an isinstance check in the reference's
sbftransformer_conv.py:133, which this package replaces anyway)."""


class SparseTensor:  # never instantiated on the X2-GNN paths
    pass

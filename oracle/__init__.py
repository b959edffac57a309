"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement (PyTorch / numpy, fp32 and fp64 capable) of the reference algorithm
for the X2-GNN message-passing hot path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product package
(`x2-gnn_b200/`) never does.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's OWN source files executed in the
build container through `oracle/shims` (restated third-party semantics) by
`tests/golden/make_golden.py`; the vectors are committed under `tests/golden/` and
checked by `tests/test_oracle_golden.py`.

Modules: graph (atom_graph.py / edge_graph.py), bases (envelop.py,
radial_basis_layer.py, angular_basis_layer.py, basis_func.py), conv
(sbftransformer_conv.py + PyG 2.1.0 propagate/softmax semantics), model (xgnn.py /
model.py / readout.py / residual_layer.py / atom_embedding.py callers, used for the
U0 end-to-end parity and the molecules/s CPU baseline).
"""

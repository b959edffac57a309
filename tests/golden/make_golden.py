"""Generate the committed golden vectors by executing the reference's OWN source files
(/root/reference, read-only, present only in the build container) on CPU.

Recipe = SURVEY.md App. H: third-party names come from oracle/shims (restated PyG /
torch-scatter semantics), `numpy.math` and `torch.nn.init.zeros` are patched (reference
defects, App. G), and edge_graph.py is exec'd from its on-disk text with `.numpy()`
added at its SciPy index sites (SciPy 1.18 no longer accepts torch tensors there).  No
reference source is copied into this repo; only numeric inputs/outputs are stored.

Run:  python tests/golden/make_golden.py          (writes tests/golden/*.pt)
"""
import math
import os
import re
import sys
import types

import numpy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.append(REF)

numpy.math = math                                                    # basis_func.py:81
torch.nn.init.zeros = lambda t: t.zero_() if t is not None else None  # sbftransformer_conv.py:81-82


def load_edge_graph():
    """exec the reference's edge_graph.py with `.numpy()` at the SciPy sites."""
    src = open(os.path.join(REF, "edge_graph.py")).read()
    src = src.replace("torch.ones(edge_index[0].size()[0]), edge_index)",
                      "torch.ones(edge_index[0].size()[0]).numpy(), edge_index.numpy())")
    src = src.replace("adj_matrix[edge_index[1]]", "adj_matrix[edge_index[1].numpy()]")
    src = src.replace("sp.coo_matrix((edge_id, edge_index)", "sp.coo_matrix((edge_id.numpy(), edge_index.numpy())")
    src = src.replace("edge_id_matrix[edge_index[1], :]", "edge_id_matrix[edge_index[1].numpy(), :]")
    src = src.replace("edge_id_matrix[edge_index[1],:]", "edge_id_matrix[edge_index[1].numpy(),:]")
    src = src.replace("np.repeat(edge_index[0], nangles)", "torch.from_numpy(np.repeat(edge_index[0].numpy(), nangles.numpy().astype(int)))")
    src = src.replace("np.repeat(edge_index[1], nangles)", "torch.from_numpy(np.repeat(edge_index[1].numpy(), nangles.numpy().astype(int)))")
    src = src.replace(".row[angle_res]", ".row[angle_res.numpy()]").replace(".data[angle_res]", ".data[angle_res.numpy()]")
    mod = types.ModuleType("edge_graph")
    exec(compile(src, "edge_graph(reference, patched at run time)", "exec"), mod.__dict__)
    sys.modules["edge_graph"] = mod
    return mod


OUT_DIR = HERE          # `python make_golden.py --out DIR` writes elsewhere (reproducibility test)


def main():
    from x2gnn_b200 import synth
    edge_graph = load_edge_graph()
    import atom_graph                     # reference
    import envelop, radial_basis_layer, angular_basis_layer  # reference
    import sbftransformer_conv            # reference
    import xgnn                           # reference (imports model.py, readout.py, ...)
    from torch_geometric.data import Data

    out = {}

    # ---------------------------------------------------------------- graph indices
    g = {}
    kat_pos = torch.tensor([[0, 0, 0], [1.2, 0, 0], [2, 1, 0], [9, 9, 9], [2.2, -1, 0.3]])
    cases = {"kat5": (kat_pos, 2.0)}
    rng = numpy.random.default_rng(5)
    for name, n in (("mol9", 9), ("mol17", 17), ("mol29", 29)):
        cases[name] = (torch.from_numpy(synth.synth_mol(n, rng)[0]), 5.0)
    cases["ball30"] = (torch.from_numpy(synth.synth_ball(30, rng)[0]), 5.0)
    for name, (pos, cut) in cases.items():
        Dij = atom_graph.calculate_Dij(pos)
        ei = atom_graph.gen_bonds_mini(Dij, cut)
        ei = ei if torch.is_tensor(ei) else torch.from_numpy(numpy.asarray(ei))
        tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(ei, pos.size(0))
        g[name] = dict(pos=pos, cutoff=cut, edge_index=ei.long(), triplets_index=tri.long().to(torch.int32),
                       edge_j=ej.to(torch.int32), edge_i=ei_.to(torch.int32), edge_k=ek.to(torch.int32))
        if pos.size(0) <= 17:
            g[name]["Dij"] = Dij
    # a collated batch (node offsets) and a directed, unsorted graph
    b = synth.qm9_batch(3, seed=11, nmin=6, nmax=12)
    ei = torch.from_numpy(b["edge_index"])
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(ei, len(b["x"]))
    g["batch3"] = dict(edge_index=ei, num_nodes=len(b["x"]), triplets_index=tri.to(torch.int32), edge_j=ej.to(torch.int32),
                       edge_i=ei_.to(torch.int32), edge_k=ek.to(torch.int32))
    perm = torch.from_numpy(numpy.random.default_rng(3).permutation(ei.size(1)))
    keep = perm[: int(0.8 * len(perm))]           # drop 20% of bonds -> directed + unsorted
    ei_u = ei[:, keep]
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(ei_u, len(b["x"]))
    g["directed_unsorted"] = dict(edge_index=ei_u, num_nodes=len(b["x"]), triplets_index=tri.to(torch.int32),
                                  edge_j=ej.to(torch.int32), edge_i=ei_.to(torch.int32), edge_k=ek.to(torch.int32))
    out["graph"] = g

    # ---------------------------------------------------------------- bases
    bs = {}
    gen = torch.Generator().manual_seed(1)
    E, T = 64, 400
    d = 0.9 + 4.1 * torch.rand(E, generator=gen)
    ang = math.pi * torch.rand(T, generator=gen)
    src = torch.randint(0, E, (T,), generator=gen)
    bs["d"], bs["angles"], bs["src"] = d, ang, src
    env = envelop.poly_envelop(5.0, 5)
    bs["env_f32"], bs["env_f64"] = env(d), env(d.double())
    rb = radial_basis_layer.RadialBasis(6, 5.0)
    bs["rbf_f32"] = rb(d).detach()
    rb64 = radial_basis_layer.RadialBasis(6, 5.0).double()
    bs["rbf_f64"] = rb64(d.double()).detach()
    for (L, R) in ((7, 6), (3, 4)):
        fb = angular_basis_layer.F_B_2D(L, R, 5.0, 5)
        bs[f"sbf_{L}_{R}_f32"] = fb(d, ang, src)
        bs[f"sbf_{L}_{R}_f64"] = fb(d.double(), ang.double(), src)
        import basis_func
        bs[f"zeros_{L}_{R}"] = torch.from_numpy(basis_func.Jn_zeros(L, R))
    ab = angular_basis_layer.AngularBasisLayer(7)
    bs["cbf_7_f64"] = ab(ang.double())
    # gradients w.r.t. the geometry: autograd through the reference's own (lambdified torch) expressions, fp64
    gg = torch.Generator().manual_seed(2)
    dd = d.double().requires_grad_(True)
    go_e = torch.randn(E, generator=gg, dtype=torch.float64)
    bs["env_go"] = go_e
    bs["env_gd_f64"] = torch.autograd.grad(env(dd), dd, go_e)[0]
    for (L, R) in ((7, 6), (3, 4)):
        fb = angular_basis_layer.F_B_2D(L, R, 5.0, 5)
        dd = d.double().requires_grad_(True)
        aa = ang.double().requires_grad_(True)
        go_s = torch.randn(T, L * R, generator=gg, dtype=torch.float64)
        gd, ga = torch.autograd.grad(fb(dd, aa, src), (dd, aa), go_s)
        bs[f"sbf_{L}_{R}_go"], bs[f"sbf_{L}_{R}_gd_f64"], bs[f"sbf_{L}_{R}_gang_f64"] = go_s, gd, ga
    aa = ang.double().requires_grad_(True)
    go_c = torch.randn(T, 7, generator=gg, dtype=torch.float64)
    bs["cbf_7_go"] = go_c
    bs["cbf_7_gang_f64"] = torch.autograd.grad(ab(aa), aa, go_c)[0]
    out["bases"] = bs

    # ---------------------------------------------------------------- conv layer
    cv = {}
    b = synth.qm9_batch(2, seed=21, nmin=5, nmax=8)
    tri = torch.from_numpy(synth.triplets_host(b["edge_index"], len(b["x"]))[0])
    E = b["edge_index"].shape[1]
    for tag, (D, H, S, R, A) in {"cfg": (128, 16, 42, 6, 128), "small": (32, 4, 6, 4, 16),
                                 "c16": (64, 4, 12, 5, 24)}.items():
        torch.manual_seed(100)
        conv = sbftransformer_conv.SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R,
                                                      dropout=0, edge_dim=A)
        ci = synth.conv_inputs(E, tri.numpy(), D, S, R, A, seed=3)
        inp = {k: torch.from_numpy(v) for k, v in ci.items()}
        gen = torch.Generator().manual_seed(9)
        gout = torch.randn(E, D, generator=gen)
        rec = dict(dims=(D, H, S, R, A), state_dict={k: v.clone() for k, v in conv.state_dict().items()},
                   grad_out=gout, **inp)
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            c = sbftransformer_conv.SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R,
                                                       dropout=0, edge_dim=A).to(dt)
            c.load_state_dict({k: v.to(dt) for k, v in rec["state_dict"].items()})
            xs = {k: inp[k].detach().clone().to(dt).requires_grad_(True) for k in ("x", "rbf", "sbf", "edge_attr")}
            o, (_, alpha) = c(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=inp["edge_index"],
                              edge_attr=xs["edge_attr"], return_attention_weights=True)
            o.backward(gout.to(dt))
            rec[f"out_{name}"] = o.detach()
            rec[f"alpha_{name}"] = alpha.detach()
            for k in xs:
                rec[f"grad_{k}_{name}"] = xs[k].grad
            for k, p in c.named_parameters():
                rec[f"gradp_{k}_{name}"] = p.grad
        if tag != "cfg":   # keep the fixture small: fp64 results only for the non-config shapes
            for k in list(rec):
                if (k.endswith("_f32") and k not in ("out_f32",)) or k.startswith("alpha_"):
                    del rec[k]
        cv[tag] = rec
    # a constructor shape outside the kernel widths (in_channels != heads * out_channels, sbftransformer_conv.py:19,47-48):
    # the reference's own layer in fp64 pins the oracle there, the kernels run it zero-padded (padded_width)
    in_ch, H, C, S, R, A = 48, 2, 16, 6, 4, 16
    torch.manual_seed(101)
    conv = sbftransformer_conv.SBFTransformerConv(in_ch, C, heads=H, sbf_dim=S, rbf_dim=R, dropout=0, edge_dim=A)
    ci = synth.conv_inputs(E, tri.numpy(), in_ch, S, R, A, seed=4)
    inp = {k: torch.from_numpy(v) for k, v in ci.items()}
    gout = torch.randn(E, H * C, generator=torch.Generator().manual_seed(10))
    rec = dict(shape=(in_ch, H, C, S, R, A), state_dict={k: v.clone() for k, v in conv.state_dict().items()},
               grad_out=gout, **inp)
    c = sbftransformer_conv.SBFTransformerConv(in_ch, C, heads=H, sbf_dim=S, rbf_dim=R, dropout=0, edge_dim=A).double()
    c.load_state_dict({k: v.double() for k, v in rec["state_dict"].items()})
    xs = {k: inp[k].detach().clone().double().requires_grad_(True) for k in ("x", "rbf", "sbf", "edge_attr")}
    o, (_, alpha) = c(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=inp["edge_index"], edge_attr=xs["edge_attr"],
                      return_attention_weights=True)
    o.backward(gout.double())
    rec["out_f64"], rec["alpha_f64"] = o.detach(), alpha.detach()
    for k in xs:
        rec[f"grad_{k}_f64"] = xs[k].grad
    for k, p in c.named_parameters():
        rec[f"gradp_{k}_f64"] = p.grad
    cv["odd48"] = rec
    out["conv"] = cv

    # ---------------------------------------------------------------- full model (small dims)
    md = {}
    b = synth.qm9_batch(3, seed=31, nmin=5, nmax=9)
    torch.manual_seed(7)
    model = xgnn.xgnn_poly(conv_layers=2, sbf_dim=3, rbf_dim=4, in_channels=32, heads=4,
                           embedding_size=16, device="cpu")
    data = Data(x=torch.from_numpy(b["x"]), edge_index=torch.from_numpy(b["edge_index"]),
                edge_attr=torch.from_numpy(b["edge_attr"]), atom_pos=torch.from_numpy(b["atom_pos"]),
                edge_num=torch.from_numpy(b["edge_num"]), batch=torch.from_numpy(b["batch"]),
                num_graphs=b["num_graphs"], y=torch.from_numpy(b["y"]))
    model.eval()
    with torch.no_grad():
        pred = model(data)
    md["small"] = dict(batch={k: (torch.from_numpy(v) if hasattr(v, "shape") else v) for k, v in b.items()},
                       hparams=dict(conv_layers=2, sbf_dim=3, rbf_dim=4, in_channels=32, heads=4,
                                    embedding_size=16),
                       state_dict={k: v.clone() for k, v in model.state_dict().items()}, pred_f32=pred)
    m64 = xgnn.xgnn_poly(conv_layers=2, sbf_dim=3, rbf_dim=4, in_channels=32, heads=4,
                         embedding_size=16, device="cpu").double()
    m64.load_state_dict({k: v.double() for k, v in md["small"]["state_dict"].items()})
    m64.eval()
    d64 = Data(**{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v)
                  for k, v in data._store.items()})
    with torch.no_grad():
        md["small"]["pred_f64"] = m64(d64)
    # config.json dims: only the key/shape contract (weights would be 4.6 MB)
    torch.manual_seed(0)
    big = xgnn.xgnn_poly(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16,
                         embedding_size=128, device="cpu")
    md["cfg_keys"] = [(k, tuple(v.shape)) for k, v in big.state_dict().items()]
    out["model"] = md

    for k, v in out.items():
        path = os.path.join(OUT_DIR, f"{k}.pt")
        torch.save(v, path)
        print(k, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if "--out" in sys.argv:
        OUT_DIR = sys.argv[sys.argv.index("--out") + 1]
        os.makedirs(OUT_DIR, exist_ok=True)
    main()

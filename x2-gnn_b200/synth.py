"""Seeded synthetic molecular inputs (host side, numpy) for tests and bench.py.

Follows the generator specified in SURVEY.md §8(d): QM9-shaped random-walk molecules,
ball-packed 500-atom graphs, radius-cutoff edges, PyG-collated batch layout
(reference record layout: qm9_allprop.py:11-19).  No file I/O, no network.
"""
from __future__ import annotations

import numpy as np

ELEMENTS = np.array([1, 6, 7, 8, 9], dtype=np.int64)          # H C N O F
ELEMENT_P = np.array([0.50, 0.35, 0.06, 0.08, 0.01])
CUTOFF = 5.0


def _accept(pos: np.ndarray, cand: np.ndarray, cutoff: float, margin: float) -> bool:
    """Candidate atom is >= 0.95 A from every existing atom and no distance falls within
    `margin` of the cutoff (keeps edge sets stable under fp32 rounding of the Gram form)."""
    d = np.linalg.norm(pos - cand, axis=1)
    return bool(d.min() >= 0.95 and np.all(np.abs(d - cutoff) > margin))


def synth_mol(n: int, rng: np.random.Generator, cutoff: float = CUTOFF, margin: float = 1e-3):
    """Random-walk molecule: n atoms, bond length U[1.00,1.55] A, >= 0.95 A exclusion."""
    pos = np.zeros((1, 3))
    while len(pos) < n:
        a = pos[rng.integers(len(pos))]
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        cand = a + v * rng.uniform(1.00, 1.55)
        if _accept(pos, cand, cutoff, margin):
            pos = np.vstack([pos, cand])
    z = rng.choice(ELEMENTS, size=n, p=ELEMENT_P)
    return pos.astype(np.float32), z


def synth_ball(n: int, rng: np.random.Generator, density: float = 0.1, cutoff: float = CUTOFF,
               margin: float = 1e-3):
    """n atoms uniformly in a ball at `density` atoms/A^3 with 0.95 A exclusion."""
    radius = (3.0 * n / (4.0 * np.pi * density)) ** (1.0 / 3.0)
    pos = np.zeros((0, 3))
    while len(pos) < n:
        c = rng.uniform(-radius, radius, size=3)
        if np.linalg.norm(c) > radius:
            continue
        if len(pos) == 0 or _accept(pos, c, cutoff, margin):
            pos = np.vstack([pos, c])
    z = rng.choice(ELEMENTS, size=n, p=ELEMENT_P)
    return pos.astype(np.float32), z


def radius_edges(pos: np.ndarray, cutoff: float = CUTOFF) -> np.ndarray:
    """All ordered pairs with 0 < d < cutoff, lexicographic (i, then j): [2,E] int64."""
    p = pos.astype(np.float64)
    d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=-1)
    adj = (d < cutoff) & (d > 0)
    return np.argwhere(adj).T.astype(np.int64)


def collate(mols, cutoff: float = CUTOFF, pair_dim: int = 338, seed: int = 0):
    """PyG-style collation of (pos, z) molecules -> dict of numpy arrays:
    x[N] i64, atom_pos[N,3] f32, edge_index[2,E] i64 (per-graph node offsets),
    edge_attr[E,pair_dim] f32 ~ N(0,0.1^2), edge_num[B], batch[N], y[B], num_graphs."""
    rng = np.random.default_rng(seed + 7919)
    xs, ps, eis, en, bt = [], [], [], [], []
    off = 0
    for g, (pos, z) in enumerate(mols):
        ei = radius_edges(pos, cutoff)
        xs.append(z)
        ps.append(pos)
        eis.append(ei + off)
        en.append(ei.shape[1])
        bt.append(np.full(len(z), g, dtype=np.int64))
        off += len(z)
    edge_index = np.concatenate(eis, axis=1)
    E = edge_index.shape[1]
    return dict(
        x=np.concatenate(xs), atom_pos=np.concatenate(ps).astype(np.float32),
        edge_index=edge_index,
        edge_attr=(rng.normal(size=(E, pair_dim)) * 0.1).astype(np.float32),
        edge_num=np.asarray(en, dtype=np.int64), batch=np.concatenate(bt),
        y=np.zeros(len(mols), dtype=np.float32), num_graphs=len(mols))


def qm9_batch(num_mols: int, seed: int = 0, nmin: int = 9, nmax: int = 29, **kw):
    """BASELINE.json configs[0]/[1]: `num_mols` QM9-sized molecules (9..29 atoms)."""
    rng = np.random.default_rng(seed)
    mols = [synth_mol(int(rng.integers(nmin, nmax + 1)), rng) for _ in range(num_mols)]
    return collate(mols, seed=seed, **kw)


def mol_triplets(pos: np.ndarray, cutoff: float = CUTOFF) -> int:
    """Number of triplets of one molecule: sum over atoms of deg * (deg - 1)."""
    deg = np.bincount(radius_edges(pos, cutoff)[1], minlength=len(pos))
    return int((deg * (deg - 1)).sum())


def qm9_shard(mols_per_rank: int, world: int, rank: int, seed: int = 0, nmin: int = 9, nmax: int = 29,
              tol: float = 0.005, **kw):
    """BASELINE.json configs[4], weak scaling: rank r's batch of ~`mols_per_rank` QM9-sized molecules whose
    triplet count is within `tol` of the N = 1 batch's (`qm9_batch(mols_per_rank, seed)`), so that every N
    runs the same per-rank work and the 1 -> 8 curve means something.  Rank 0 (and world == 1) IS the N = 1
    batch.  Rank r > 0 draws its molecules from its own seeded stream: all but the last few as they come, the
    last ones picked from the next candidates so that the total lands on the target (conv cost ~ T, and T
    varies 20x between a 9-atom and a 29-atom molecule).  Deterministic; no communication.
    Returns (batch, ids) with ids = the global molecule numbers rank * mols_per_rank + local index."""
    cutoff = kw.get("cutoff", CUTOFF)
    ids = list(range(rank * mols_per_rank, rank * mols_per_rank + mols_per_rank))
    rng0 = np.random.default_rng(seed)
    base = [synth_mol(int(rng0.integers(nmin, nmax + 1)), rng0) for _ in range(mols_per_rank)]
    if world == 1 or rank == 0:
        return collate(base, seed=seed, **kw), ids
    target = sum(mol_triplets(p, cutoff) for p, _ in base)
    rng = np.random.default_rng(seed + 1000003 * rank)
    keep = max(mols_per_rank - 16, 0) if mols_per_rank >= 32 else 0
    mols = [synth_mol(int(rng.integers(nmin, nmax + 1)), rng) for _ in range(keep)]
    have = sum(mol_triplets(p, cutoff) for p, _ in mols)
    cands = [synth_mol(int(rng.integers(nmin, nmax + 1)), rng) for _ in range(max(256, 4 * mols_per_rank))]
    cost = [mol_triplets(p, cutoff) for p, _ in cands]
    free = list(range(len(cands)))
    while len(mols) < mols_per_rank:
        left = mols_per_rank - len(mols)
        want = (target - have) / left                  # aim each remaining pick at the mean of what is missing
        k = min(free, key=lambda c: abs(cost[c] - want))
        free.remove(k)
        mols.append(cands[k])
        have += cost[k]
    if abs(have - target) > tol * target:              # never seen; keeps the promise checkable
        raise RuntimeError(f"qm9_shard: rank {rank} reached T={have}, target {target}")
    return collate(mols, seed=seed + 104729 * rank, **kw), ids


def aid_batch(indices, path: str = None, **kw):
    """BASELINE.json configs[2] ("OCELOT-sized"): molecules `indices` of the real 60-146-atom geometries the
    reference ships (raw/AID_kcal.xyz), read from the numeric fixture tests/golden/aid_geometries.npz
    (tests/golden/make_aid_fixture.py).  Pair features are synthetic as everywhere else (the reference
    computes them with pyscf, out of scope)."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "aid_geometries.npz")
    d = np.load(path)
    ptr = d["ptr"]
    mols = [(d["pos"][ptr[i]:ptr[i + 1]].astype(np.float32), d["z"][ptr[i]:ptr[i + 1]].astype(np.int64))
            for i in indices]
    return collate(mols, **kw)


def ball_batch(num_mols: int, n_atoms: int = 500, seed: int = 0, **kw):
    """BASELINE.json configs[3]: ball-packed `n_atoms`-atom graphs."""
    rng = np.random.default_rng(seed)
    return collate([synth_ball(n_atoms, rng) for _ in range(num_mols)], seed=seed, **kw)


def triplets_host(edge_index: np.ndarray, num_nodes: int):
    """Host-side triplet enumeration (SURVEY.md App. E ordering) used ONLY to size
    synthetic conv inputs for bench/tests set-up; the product path builds triplets on the
    GPU (edge_graph.vertex_to_edge_2)."""
    src, dst = edge_index
    E = src.shape[0]
    order = np.lexsort((dst, src))
    s_sorted, d_sorted = src[order], dst[order]
    start = np.searchsorted(s_sorted, np.arange(num_nodes + 1))
    deg = start[1:] - start[:-1]
    cnt = deg[dst]
    tgt = np.repeat(np.arange(E), cnt)
    base = np.repeat(start[dst], cnt)
    within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    f_sorted_pos = base + within
    k = d_sorted[f_sorted_pos]
    f = order[f_sorted_pos]
    keep = k != src[tgt]
    return np.stack([f[keep], tgt[keep]]).astype(np.int64), dst[tgt][keep], src[tgt][keep], k[keep]


def conv_inputs(E: int, trip_index: np.ndarray, D: int = 128, S: int = 42, R: int = 6,
                A: int = 128, seed: int = 0):
    """SURVEY.md §8(d) config-4 style tensor inputs for one SBFTransformerConv call:
    x, sbf, edge_attr ~ N(0,1); rbf ~ U(-1,1)."""
    rng = np.random.default_rng(seed + 104729)
    T = trip_index.shape[1]
    return dict(
        x=rng.normal(size=(E, D)).astype(np.float32),
        rbf=rng.uniform(-1, 1, size=(E, R)).astype(np.float32),
        sbf=rng.normal(size=(T, S)).astype(np.float32),
        edge_attr=rng.normal(size=(T, A)).astype(np.float32),
        edge_index=trip_index.astype(np.int64))

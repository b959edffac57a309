"""oracle.graph -- TEST INFRASTRUCTURE.  CPU restatement of the reference's
index-construction code: atom_graph.py:32-45 (radius graph) and edge_graph.py:12-30
(atom graph -> line graph / triplets; ordering per SURVEY.md App. E)."""
import numpy as np
import torch


def calculate_Dij(atom_pos: torch.Tensor) -> torch.Tensor:
    """atom_graph.py:32-35: Gram-trick distance matrix, relu(sqrt(|a|^2+|b|^2-2ab))."""
    gram = atom_pos @ atom_pos.T
    h = torch.diag(gram).unsqueeze(0).expand(atom_pos.size(0), -1)
    return torch.relu((h + h.T - 2 * gram) ** 0.5)


def gen_bonds_mini(Dij: torch.Tensor, cutoff: float = 5.0) -> torch.Tensor:
    """atom_graph.py:42-45: edge_index = argwhere((D < cutoff) & D != 0), [2,E] int64,
    lexicographic by (i, j)."""
    adj = (Dij < cutoff) & Dij.bool()
    return torch.nonzero(adj).T.contiguous()


def vertex_to_edge_2(edge_index: torch.Tensor, num_nodes: int):
    """edge_graph.py:12-30 restated from its definition: for each
    bond e=(i->j) in order, for each bond f=(j->k) with k ascending and k != i, emit
    (source=f, target=e) and atoms (j, i, k).  Returns triplets_index[2,T], edge_j,
    edge_i, edge_k (all int64) -- NOTE the j, i, k return order (edge_graph.py:30)."""
    ei = edge_index.cpu().numpy().astype(np.int64)
    src, dst = ei[0], ei[1]
    E = src.shape[0]
    # CSR of out-bonds per atom with columns sorted ascending (scipy tocsr semantics)
    order = np.lexsort((dst, src))
    s_sorted, d_sorted = src[order], dst[order]
    start = np.searchsorted(s_sorted, np.arange(num_nodes + 1))
    deg = start[1:] - start[:-1]
    cnt = deg[dst]                                   # edge_graph.py:15 (nangles)
    tgt = np.repeat(np.arange(E, dtype=np.int64), cnt)
    total = int(cnt.sum())
    within = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    pos = np.repeat(start[dst], cnt) + within
    k = d_sorted[pos]                                # edge_graph.py:19
    f = order[pos]                                   # edge ids (edge_graph.py:21,27)
    i = src[tgt]
    j = dst[tgt]
    keep = k != i                                    # edge_graph.py:20
    tri = np.stack([f[keep], tgt[keep]])             # [jk_idx; ij_idx]  (edge_graph.py:28)
    as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int64)))
    return as_t(tri), as_t(j[keep]), as_t(i[keep]), as_t(k[keep])


def vertex_to_edge_2_bruteforce(edge_index: torch.Tensor, num_nodes: int):
    """Literal App. E double loop -- pure Python, small cases only."""
    ei = edge_index.tolist()
    E = len(ei[0])
    out = {}
    for e in range(E):
        out.setdefault(ei[0][e], []).append((ei[1][e], e))
    for v in out.values():
        v.sort()
    f_l, e_l, j_l, i_l, k_l = [], [], [], [], []
    for e in range(E):
        i, j = ei[0][e], ei[1][e]
        for k, f in out.get(j, []):
            if k == i:
                continue
            f_l.append(f); e_l.append(e); j_l.append(j); i_l.append(i); k_l.append(k)
    t = lambda a: torch.tensor(a, dtype=torch.int64)
    return torch.stack([t(f_l), t(e_l)]) if f_l else torch.zeros(2, 0, dtype=torch.int64), \
        t(j_l), t(i_l), t(k_l)


def csr_by_target(edge_index: torch.Tensor, num_targets: int) -> torch.Tensor:
    """rowptr[E+1] (int64) of a target-sorted triplet list."""
    cnt = torch.bincount(edge_index[1], minlength=num_targets)
    return torch.cat([cnt.new_zeros(1), torch.cumsum(cnt, 0)])

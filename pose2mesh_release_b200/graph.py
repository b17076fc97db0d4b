"""Host-side graph baking: the mesh hierarchy the MeshNet kernels consume.

Drop-in for the reference's ``graph_utils.build_coarse_graphs`` (lib/graph_utils.py:75-95) and the
``coarsening`` functions behind it (lib/coarsening.py:6-64, 67-290, 322-328): same signature, same
return tuple ``(graph_Adj, graph_L, graph_perm, perm_reverse)`` with scipy CSR float64 Laplacians
ordered fine -> coarse and the joint graph last.  The numbers are defined by the reference's
algorithm *including its quirks* (``L/(2*lmax) - I`` and an un-rescaled joint graph;
the off-by-one row lengths and "first stored entry as W_ii" of HEM_one_level), because they decide
which vertices end up where in the baked CSR.

Implementation: vectorised numpy/scipy for the algebra, and the one inherently sequential part —
greedy heavy-edge matching, pure-Python loops in the reference (lib/coarsening.py:153-211) — in
native code (csrc/graph_host.cpp, ``p2m_graph_match_level``).  This is start-up work (~0.1 s for the
6890-vertex SMPL topology), not the hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import _lib

__all__ = ["build_graph", "build_adj", "build_coarse_graphs", "coarsen", "laplacian", "rescale_L", "lmax_L",
           "compute_perm", "perm_adjacency", "perm_index_reverse", "synthetic_sphere_faces",
           "H36M_SKELETON", "H36M_FLIP_PAIRS", "MANO_SKELETON", "MANO_HORI_CONN"]

# joint graphs used by the reference's demo / datasets (demo/run.py:75-78, 109-112)
H36M_SKELETON = ((0, 7), (7, 8), (8, 9), (9, 10), (8, 11), (11, 12), (12, 13), (8, 14), (14, 15), (15, 16),
                 (0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6))
H36M_FLIP_PAIRS = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))
MANO_SKELETON = ((0, 1), (0, 5), (0, 9), (0, 13), (0, 17), (1, 2), (2, 3), (3, 4), (5, 6), (6, 7), (7, 8),
                 (9, 10), (10, 11), (11, 12), (13, 14), (14, 15), (15, 16), (17, 18), (18, 19), (19, 20))
MANO_HORI_CONN = ((1, 5), (5, 9), (9, 13), (13, 17), (2, 6), (6, 10), (10, 14), (14, 18), (3, 7), (7, 11),
                  (11, 15), (15, 19), (4, 8), (8, 12), (12, 16), (16, 20))


def build_graph(mesh_face: np.ndarray, num_vertex: int) -> sp.csr_matrix:
    """Triangle list -> symmetric 0/1 vertex adjacency (graph_utils.py:37-60)."""
    tri = np.asarray(mesh_face, dtype=np.int64)
    i = np.concatenate((tri[:, 0], tri[:, 1], tri[:, 0]))
    j = np.concatenate((tri[:, 1], tri[:, 2], tri[:, 2]))
    upper = sp.csr_matrix((np.ones(i.size), (i, j)), shape=(num_vertex, num_vertex))
    upper.data[:] = 1.0                      # clip duplicate directed pairs
    sym = upper.maximum(upper.T).tocsr()
    sym.sort_indices()
    return sym


def build_adj(joint_num: int, skeleton: Sequence, flip_pairs: Sequence) -> np.ndarray:
    """Joint graph: bones + flip pairs + self loops (graph_utils.py:63-72)."""
    adj = np.eye(joint_num)
    pairs = np.asarray(list(skeleton) + list(flip_pairs), dtype=np.int64).reshape(-1, 2)
    off = np.zeros((joint_num, joint_num))
    off[pairs[:, 0], pairs[:, 1]] = 1.0
    off[pairs[:, 1], pairs[:, 0]] = 1.0
    return off + adj


def laplacian(W: sp.spmatrix, normalized: bool = True) -> sp.csr_matrix:
    """I - D^-1/2 W D^-1/2 (coarsening.py:6-25); isolated vertices get L_ii = 1."""
    W = sp.csr_matrix(W)
    deg = np.asarray(W.sum(axis=0)).ravel().astype(W.dtype, copy=True)
    if not normalized:
        return sp.csr_matrix(sp.diags(deg, 0) - W)
    deg += np.spacing(np.array(0, W.dtype))
    scale = sp.diags(1.0 / np.sqrt(deg), 0)
    L = sp.csr_matrix(sp.identity(deg.size, dtype=W.dtype) - scale * W * scale)
    if abs(L - L.T).mean() >= 1e-9:
        raise ValueError("laplacian: adjacency is not symmetric")
    return L


def lmax_L(L: sp.spmatrix) -> float:
    """Largest eigenvalue (coarsening.py:37-39)."""
    return float(spla.eigsh(L, k=1, which="LM", return_eigenvectors=False)[0])


def rescale_L(L: sp.spmatrix, lmax: float = 2) -> sp.csr_matrix:
    """L / (2*lmax) - I — the reference's actual arithmetic (coarsening.py:28-34)."""
    L = sp.csr_matrix(L, copy=True)
    L.data = L.data / (lmax * 2)
    return sp.csr_matrix(L - sp.identity(L.shape[0], format="csr", dtype=L.dtype))


def _match_level(W: sp.spmatrix, visit: np.ndarray, weights: np.ndarray):
    coo = sp.coo_matrix(W)
    coo.sum_duplicates()
    nz = coo.data != 0
    r, c, v = coo.row[nz], coo.col[nz], coo.data[nz]
    order = np.lexsort((c, r))
    r = np.ascontiguousarray(r[order], dtype=np.int32)
    c = np.ascontiguousarray(c[order], dtype=np.int32)
    v = np.ascontiguousarray(v[order], dtype=np.float64)
    n = int(r[-1]) + 1
    visit = np.ascontiguousarray(visit, dtype=np.int64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    if visit.size < n or weights.size < n:
        raise ValueError("HEM: graph has trailing empty rows")
    cluster = np.zeros(n, dtype=np.int32)
    lib = _lib.load()
    k = lib.p2m_graph_match_level(r.size, r.ctypes.data_as(_lib.c_int32_p), c.ctypes.data_as(_lib.c_int32_p),
                                  v.ctypes.data_as(C.POINTER(C.c_double)), visit.ctypes.data_as(_lib.c_int64_p),
                                  weights.ctypes.data_as(C.POINTER(C.c_double)),
                                  cluster.ctypes.data_as(_lib.c_int32_p))
    if k < 0:
        raise RuntimeError("p2m_graph_match_level rejected its input")
    return cluster, (r, c, v)


def HEM(W: sp.spmatrix, levels: int):
    """Greedy heavy-edge matching, `levels` times (coarsening.py:67-149)."""
    np.random.permutation(range(W.shape[0]))  # the reference draws (and discards) this; keep the global RNG in step
    colsum = np.asarray(W.sum(axis=0)).ravel()
    visit = np.argsort(colsum)
    degree = colsum - W.diagonal()
    graphs, parents = [W], []
    for _ in range(levels):
        cluster, (r, c, v) = _match_level(W, visit, np.asarray(degree).ravel())
        parents.append(cluster)
        n_new = int(cluster.max()) + 1
        W = sp.csr_matrix((v, (cluster[c], cluster[r])), shape=(n_new, n_new))
        W.eliminate_zeros()
        graphs.append(W)
        degree = np.asarray(W.sum(axis=0)).ravel()
        visit = np.argsort(degree)
    return graphs, parents


def compute_perm(parents: Sequence[np.ndarray]) -> List[List[int]]:
    """Binary-tree vertex order with fake (singleton-padding) vertices (coarsening.py:214-258),
    vectorised: at every level each node of the coarser order contributes its <= 2 children in
    ascending id, then as many fresh fake ids as needed to make two."""
    if len(parents) == 0:
        return []
    order = np.arange(int(np.max(parents[-1])) + 1, dtype=np.int64)
    out = [order]
    for parent in parents[::-1]:
        parent = np.asarray(parent, dtype=np.int64)
        n_real = parent.size
        n_nodes = order.size
        n_parent_ids = max(int(order.max()) + 1, int(parent.max()) + 1)
        count = np.bincount(parent, minlength=n_parent_ids)
        if count.max() > 2:
            raise ValueError("compute_perm: a cluster has more than two children")
        kids_sorted = np.argsort(parent, kind="stable")           # children grouped by parent, ascending id
        first = np.concatenate(([0], np.cumsum(count)[:-1]))
        cnt = count[order]
        need = 2 - cnt
        fake_base = n_real + np.concatenate(([0], np.cumsum(need)[:-1]))
        layer = np.empty((n_nodes, 2), dtype=np.int64)
        has1, has2 = cnt >= 1, cnt == 2
        layer[has1, 0] = kids_sorted[first[order[has1]]]
        layer[has2, 1] = kids_sorted[first[order[has2]] + 1]
        layer[~has1, 0] = fake_base[~has1]
        layer[~has2, 1] = fake_base[~has2] + (cnt[~has2] == 0)
        order = layer.reshape(-1)
        out.append(order)
    m_last = out[0].size
    for i, layer in enumerate(out):
        if not np.array_equal(np.sort(layer), np.arange(m_last * 2 ** i)):
            raise ValueError("compute_perm: ordering is not a permutation")
    return [list(map(int, layer)) for layer in out[::-1]]


def perm_adjacency(A: sp.spmatrix, indices) -> sp.coo_matrix:
    """Pad with isolated fake vertices and relabel (coarsening.py:265-290)."""
    if indices is None:
        return A
    indices = np.asarray(indices)
    coo = sp.coo_matrix(A)
    rank = np.argsort(indices)
    n = indices.size
    out = sp.coo_matrix((coo.data, (rank[coo.row], rank[coo.col])), shape=(n, n))
    if abs(out - out.T).mean() >= 1e-8:
        raise ValueError("perm_adjacency: result not symmetric")
    return out


def perm_index_reverse(indices) -> np.ndarray:
    """Inverse permutation (coarsening.py:322-328)."""
    indices = np.asarray(indices)
    rev = np.empty_like(indices)
    rev[indices] = np.arange(indices.size)
    return rev


def coarsen(A: sp.spmatrix, levels: int):
    """coarsening.py:43-64."""
    graphs, parents = HEM(A, levels)
    perms = compute_perm(parents)
    adjacencies, laplacians = [], []
    for i, G in enumerate(graphs):
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = sp.csr_matrix(G)
        G.eliminate_zeros()
        adjacencies.append(G)
        laplacians.append(laplacian(G, normalized=True))
    return adjacencies, laplacians, (perms if len(perms) > 0 else None)


def build_coarse_graphs(mesh_face, joint_num, skeleton, flip_pairs, levels=9) -> Tuple[list, list, list, np.ndarray]:
    """graph_utils.py:75-95."""
    joint_adj = sp.csr_matrix(build_adj(joint_num, skeleton, flip_pairs))
    joint_adj.eliminate_zeros()
    mesh_adj = build_graph(mesh_face, int(np.max(mesh_face)) + 1)
    graph_Adj, graph_L, graph_perm = coarsen(mesh_adj, levels=levels)
    graph_L[-1] = laplacian(joint_adj, normalized=True)   # joint graph: NOT rescaled below (F5)
    graph_Adj[-1] = joint_adj
    for i in range(levels):
        graph_L[i] = rescale_L(graph_L[i], lmax_L(graph_L[i]))
    return graph_Adj, graph_L, graph_perm, perm_index_reverse(graph_perm[0])


def synthetic_sphere_faces(n_vertex: int, seed: int = 0) -> np.ndarray:
    """Seeded closed genus-0 triangulation (convex hull of random unit-sphere points).  The SMPL /
    MANO topologies are licence-gated; n=6890 has SMPL's exact face / edge counts,
    and seed 2 (resp. n=778, seed 1) reproduces the real level sizes 12288..96 (resp. 1088..68)."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(n_vertex, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    return np.asarray(ConvexHull(pts).simplices, dtype=np.int64)

"""SolveLinear: the sparse least-squares post-process of the CAD pipeline (reference src/lib/linear.cc:10-230,
bound by src/interface/linear_layer.cc:5-84).  Host code in the reference (Eigen SimplicialLDLT) and host code
here (scipy.sparse): it runs once per shape after the optimisation and is not on the GPU hot path
(SURVEY.md s8f rank 4).  Both systems are symmetric positive definite, so the solution is unique and the
factorisation used does not matter beyond rounding.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _laplacian_terms(n, v0, v1, reg):
    """Triplets of  sum_e reg_e (e_v0 - e_v1)(e_v0 - e_v1)^T  (linear.cc:66-70, :83-87, :197-200)."""
    rows = np.concatenate([v0, v0, v1, v1]); cols = np.concatenate([v0, v1, v0, v1])
    vals = np.concatenate([reg, -reg, -reg, reg])
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n))


def _solve(A, B):
    lu = spla.splu(sp.csc_matrix(A))
    return np.stack([lu.solve(np.ascontiguousarray(B[:, j])) for j in range(3)], axis=1)


def linear_estimation(V, F, E, references, graphV, rigidity):
    """LinearEstimation (linear.cc:10-117): returns the new vertex positions (float64 [n,3]).
    V [n,3] vertices, F [m,3] faces, E [e,2] extra edges, references [n] graph node of every vertex,
    graphV [g,3] deformed graph nodes."""
    V = np.asarray(V, dtype=np.float64); graphV = np.asarray(graphV, dtype=np.float64)
    F = np.asarray(F, dtype=np.int64).reshape(-1, 3); E = np.asarray(E, dtype=np.int64).reshape(-1, 2)
    ref = np.asarray(references, dtype=np.int64).reshape(-1)
    n, g = V.shape[0], graphV.shape[0]
    # every graph node wants the mean of its vertices on the node (:37-50): A += (1/c^2) 1 1^T, B += (1/c) graphV
    cnt = np.bincount(ref, minlength=g).astype(np.float64)
    w = 1.0 / cnt[ref]
    M = sp.csr_matrix((w, (ref, np.arange(n))), shape=(g, n))           # M[i, v] = 1/c_i for v on node i
    A = (M.T @ M).tocoo()
    B = w[:, None] * graphV[ref]
    # weak anchor of every vertex to its node (:52-57)
    A = A + sp.identity(n, format="coo") * 1e-6
    B = B + 1e-6 * graphV[ref]
    # rigidity along face edges and extra edges (:60-89)
    v0 = np.concatenate([F[:, 0], F[:, 1], F[:, 2], E[:, 0]]); v1 = np.concatenate([F[:, 1], F[:, 2], F[:, 0], E[:, 1]])
    d = V[v0] - V[v1]
    reg = (rigidity * 2e-2 / (np.sqrt((d * d).sum(1)) + 1e-8)) ** 2
    A = A + _laplacian_terms(n, v0, v1, reg)
    np.add.at(B, v0, reg[:, None] * d)
    np.add.at(B, v1, -reg[:, None] * d)
    return _solve(A, B)


def linear_estimation_with_rot(V, F, TV, rigidity):
    """LinearEstimationWithRot (linear.cc:119-230): V [n,3] rest vertices, TV [n,3] target positions of the same
    vertices; per-vertex rotation (polar factor of the neighbourhood covariance) and scale, then one SPD solve."""
    V = np.asarray(V, dtype=np.float64); TV = np.asarray(TV, dtype=np.float64)
    F = np.asarray(F, dtype=np.int64).reshape(-1, 3)
    n = V.shape[0]
    e0 = np.concatenate([F[:, 0], F[:, 1], F[:, 2]]); e1 = np.concatenate([F[:, 1], F[:, 2], F[:, 0]])
    # links: undirected, de-duplicated neighbourhoods (std::set, :127-134)
    a = np.concatenate([e0, e1]); b = np.concatenate([e1, e0])
    key = np.unique(a * n + b)
    a, b = key // n, key % n
    d1 = V[b] - V[a]; d2 = TV[b] - TV[a]
    len_o = np.bincount(a, weights=np.sqrt((d1 * d1).sum(1)), minlength=n)
    len_c = np.bincount(a, weights=np.sqrt((d2 * d2).sum(1)), minlength=n)
    cov = np.zeros((n, 3, 3))
    np.add.at(cov, a, d2[:, :, None] * d1[:, None, :])                   # covariance += d2 d1^T (:153)
    scale = len_c / (len_o + 1e-8)
    U, _, Vt = np.linalg.svd(cov)
    R = U @ Vt                                                           # :156-160
    A = sp.identity(n, format="coo")                                      # :179-183
    B = TV.copy()
    # every directed face edge in both directions (:186-204)
    v0 = np.concatenate([e0, e1]); v1 = np.concatenate([e1, e0])
    off = V[v1] - V[v0]
    reg = 1.0 * 2e-2 / np.sqrt((off * off).sum(1)) * rigidity
    off = scale[v0][:, None] * np.einsum("nij,nj->ni", R[v0], off)
    A = A + _laplacian_terms(n, v0, v1, reg)
    np.add.at(B, v0, -reg[:, None] * off)
    np.add.at(B, v1, reg[:, None] * off)
    return _solve(A, B)

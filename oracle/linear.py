"""TEST INFRASTRUCTURE (not product code): literal restatement of the reference's sparse post-process
(src/lib/linear.cc) with Python loops and a dense solve, for small cases.  Each block cites the lines it follows."""
import numpy as np


def linear_estimation(V, F, E, references, graphV, rigidity):
    V = np.array(V, dtype=np.float64); graphV = np.asarray(graphV, dtype=np.float64)
    n = V.shape[0]
    A = np.zeros((n, n)); B = np.zeros((n, 3))
    cells = [[] for _ in range(graphV.shape[0])]
    for i, r in enumerate(references):            # linear.cc:33-35
        cells[int(r)].append(i)
    for i, c in enumerate(cells):                 # :38-50
        if not c:
            continue
        w = 1.0 / len(c)
        for j in c:
            for k in c:
                A[j, k] += w * w
        for j in c:
            B[j] += w * graphV[i]
    for i in range(n):                            # :52-57
        A[i, i] += 1e-6
        B[i] += 1e-6 * graphV[int(references[i])]
    edges = [(int(f[j]), int(f[(j + 1) % 3])) for f in F for j in range(3)] + [(int(a), int(b)) for a, b in E]   # :60-89
    for v0, v1 in edges:
        reg = rigidity * 2e-2 / (np.linalg.norm(V[v0] - V[v1]) + 1e-8)
        reg *= reg
        A[v0, v0] += reg; A[v0, v1] -= reg; A[v1, v0] -= reg; A[v1, v1] += reg
        B[v0] += reg * (V[v0] - V[v1]); B[v1] += reg * (V[v1] - V[v0])
    return np.linalg.solve(A, B)                  # :91-116


def linear_estimation_with_rot(V, F, TV, rigidity):
    V = np.array(V, dtype=np.float64); TV = np.asarray(TV, dtype=np.float64)
    n = V.shape[0]
    E = [(int(f[j]), int(f[(j + 1) % 3])) for f in F for j in range(3)]          # linear.cc:117-124
    links = [set() for _ in range(n)]
    for a, b in E:                                                                # :127-134
        links[a].add(b); links[b].add(a)
    R = np.zeros((n, 3, 3)); S = np.zeros(n)
    for i in range(n):                                                            # :136-162
        cov = np.zeros((3, 3)); lo = lc = 0.0
        for p in links[i]:
            d1 = V[p] - V[i]; d2 = TV[p] - TV[i]
            lo += np.linalg.norm(d1); lc += np.linalg.norm(d2)
            cov += np.outer(d2, d1)
        U, _, Vt = np.linalg.svd(cov)
        R[i] = U @ Vt; S[i] = lc / (lo + 1e-8)
    A = np.eye(n); B = TV.copy()                                                  # :179-183
    for a, b in E:                                                                # :186-204
        for v0, v1 in ((a, b), (b, a)):
            off = V[v1] - V[v0]
            reg = 1.0 * 2e-2 / np.linalg.norm(off) * rigidity
            off = S[v0] * (R[v0] @ off)
            A[v0, v0] += reg; A[v0, v1] -= reg; A[v1, v0] -= reg; A[v1, v1] += reg
            B[v0] -= reg * off; B[v1] += reg * off
    return np.linalg.solve(A, B)

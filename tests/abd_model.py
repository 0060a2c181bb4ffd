"""numpy model of the pivoted recursive condensation + mesh-slab partition the CUDA library implements
(rustedscithe_b200/csrc/abd.cuh).  TEST INFRASTRUCTURE: it lets the host-side logic of the
multi-rank path (partition, what is exchanged, who solves what) run on CPU under gloo, where the CUDA library cannot.
"""
import numpy as np


def eliminate(M, W, w):
    """Partial-pivot Gaussian elimination on the stacked 2nv x nv panel M, applied to W (coupling) and w (rhs)."""
    M, W, w = M.copy(), W.copy(), w.copy()
    nv = M.shape[1]
    for k in range(nv):
        p = k + int(np.argmax(np.abs(M[k:, k])))
        if p != k:
            M[[k, p]], W[[k, p]], w[[k, p]] = M[[p, k]], W[[p, k]], w[[p, k]]
        l = M[k + 1:, k] / M[k, k]
        M[k + 1:, k:] -= np.outer(l, M[k, k:])
        W[k + 1:] -= np.outer(l, W[k])
        w[k + 1:] -= l * w[k]
    return M, W, w


def condense(As, Bs, rs):
    """Rows A_i Y_i + B_i Y_{i+1} = r_i of one slab -> one row C Y_first + D Y_last = g (+ back-substitution data)."""
    nv = As[0].shape[0]
    C, D, g = As[0].copy(), Bs[0].copy(), rs[0].copy()
    info = []
    Z = np.zeros((nv, nv))
    for i in range(1, len(As)):
        M, W, w = eliminate(np.vstack([D, As[i]]), np.block([[C, Z], [Z, Bs[i]]]), np.concatenate([g, rs[i]]))
        info.append((M[:nv], W[:nv, :nv], W[:nv, nv:], w[:nv]))
        C, D, g = W[nv:, :nv], W[nv:, nv:], w[nv:]
    return C, D, g, info


def back_substitute(info, Ys, Ye):
    out = [None] * len(info)
    Yn = Ye
    for i in range(len(info) - 1, -1, -1):
        U, E, F, e = info[i]
        Yn = np.linalg.solve(np.triu(U), e - E @ Ys - F @ Yn)
        out[i] = Yn
    return out


def close_with_bcs(C, D, g, left, right):
    """Last block row + pinned components (zero correction): nv x nv system in the free components."""
    nv = C.shape[0]
    lf = [v for v in range(nv) if v not in left]
    rf = [v for v in range(nv) if v not in right]
    z = np.linalg.solve(np.hstack([C[:, lf], D[:, rf]]), g)
    Y0, YN = np.zeros(nv), np.zeros(nv)
    Y0[lf], YN[rf] = z[:len(lf)], z[len(lf):]
    return Y0, YN


def solve_single(A, F, left, right, S=8):
    """Whole system on one rank: levels of slabs of S rows until one row is left."""
    N, nv, _ = A.shape
    I = np.eye(nv)
    As, Bs, rs = [A[j] for j in range(N)], [I] * N, [F[j] for j in range(N)]
    levels = []
    while len(As) > 1:
        nA, nB, nr, infos = [], [], [], []
        for s in range(0, len(As), S):
            C, D, g, info = condense(As[s:s + S], Bs[s:s + S], rs[s:s + S])
            nA.append(C); nB.append(D); nr.append(g); infos.append(info)
        levels.append((infos, len(As)))
        As, Bs, rs = nA, nB, nr
    vals = list(close_with_bcs(As[0], Bs[0], rs[0], left, right))
    return _expand(levels, vals, S)


def _expand(levels, vals, S):
    for infos, nrows in reversed(levels):
        new = [None] * (nrows + 1)
        for si, info in enumerate(infos):
            s, e = si * S, min(si * S + S, nrows)
            new[s], new[e] = vals[si], vals[si + 1]
            for i, Yi in enumerate(back_substitute(info, vals[si], vals[si + 1])):
                new[s + 1 + i] = Yi
        vals = new
    return np.array(vals)


def partition(N, rank, world):
    """Intervals [begin, end) of a rank -- must match rst_b200_create (rst_b200.cu: j_begin / j_end)."""
    return (N * rank) // world, (N * (rank + 1)) // world


def local_condense(A_loc, F_loc, S=8):
    """Phase 1 on every rank (rst_b200_factor_local + solve_local_forward): slab -> one interface row."""
    n, nv, _ = A_loc.shape
    I = np.eye(nv)
    As, Bs, rs = [A_loc[j] for j in range(n)], [I] * n, [F_loc[j] for j in range(n)]
    levels = []
    while len(As) > 1:
        nA, nB, nr, infos = [], [], [], []
        for s in range(0, len(As), S):
            C, D, g, info = condense(As[s:s + S], Bs[s:s + S], rs[s:s + S])
            nA.append(C); nB.append(D); nr.append(g); infos.append(info)
        levels.append((infos, len(As)))
        As, Bs, rs = nA, nB, nr
    return As[0], Bs[0], rs[0], levels


def reduced_solve(rows, left, right):
    """Phase 2, redundantly on every rank: the `world` gathered interface rows -> values of the world+1 slab-end nodes."""
    As, Bs, rs = [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows]
    C, D, g, info = condense(As, Bs, rs)
    Y0, YN = close_with_bcs(C, D, g, left, right)
    return [Y0] + back_substitute(info, Y0, YN) + [YN]


def local_expand(levels, Ys, Ye, S=8):
    """Phase 3 on every rank: back-substitute the slab interior from its two end nodes."""
    return _expand(levels, [Ys, Ye], S)

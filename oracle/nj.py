"""Oracle: neighbour joining, hard and soft.  TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``nj_hard``      restates ``Cpeeler.nj_np`` (dodonaphy/cython/Cpeeler.pyx:15-125) scalar by
                   scalar, in the same operation order, so peel AND blens are bit-identical.
* ``soft_argmin``  restates ``soft.argmin`` (dodonaphy/soft.py:38-53) using only the last row of
                   each SoftSort matrix (the only row the reference reads, soft.py:45,50): O(n).
* ``soft_argmin_dense``  the literal n x n construction (soft.py:13-17), for pinning the closed
                   form on small inputs.
* ``nj_soft``      restates ``peeler.nj_torch`` (dodonaphy/peeler.py:155-220) with ``compute_Q``
                   (:223-248) and ``get_new_dist_soft`` (:251-265) on the padded N x N matrix.
* ``nj_soft_torch``  the same in torch float64 with autograd (one-hot selections carry no
                   gradient, soft.py:56-61) -- the gradient oracle for the CUDA backward.
"""
import numpy as np

EPS = np.finfo(np.float64).eps  # peeler.py:9, Cpeeler.pyx:11
LARGE = 1e9  # peeler.py:11


# --------------------------------------------------------------------------- hard NJ
def nj_hard(pdm):
    """Cpeeler.nj_np.  pdm: (S,S) float64.  Returns (peel int64 (S-1,3), blens float64 (2S-2,))."""
    pdm = np.asarray(pdm, dtype=np.float64)
    S = pdm.shape[0]
    N = 2 * S - 1
    d = [[0.0] * N for _ in range(N)]  # d[a][b] == node a's _nj_distances[b]
    for i in range(S):
        row = pdm[i]
        di = d[i]
        for j in range(S):
            di[j] = float(row[j])
    xsub = [0.0] * N
    pool = list(range(S))
    # Cpeeler.pyx:46-54: running row sums, sequential in pool order, skipping self
    for i in pool:
        acc = 0.0
        di = d[i]
        for j in pool:
            if i == j:
                continue
            acc += di[j]
        xsub[i] = acc
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    blens = np.zeros(2 * S - 2, dtype=np.float64)
    n_pool = S
    while n_pool > 1:
        n_pool = len(pool)
        # Cpeeler.pyx:58-66: first strict minimum in pool order
        min_q = np.finfo(np.float64).max
        best = None
        nm2 = float(n_pool) - 2.0
        for a in range(n_pool - 1):
            i = pool[a]
            di = d[i]
            xi = xsub[i]
            for b in range(a + 1, n_pool):
                j = pool[b]
                v1 = nm2 * di[j]
                q = v1 - xi - xsub[j]
                if q < min_q:
                    min_q = q
                    best = (i, j)
        n1, n2 = best
        int_i = S - n_pool
        parent = int_i + S
        peel[int_i] = (n1, n2, parent)
        pool.remove(n1)
        pool.remove(n2)
        # Cpeeler.pyx:81-95
        d12 = d[n1][n2]
        new_x = 0.0
        for nd in pool:
            v1 = 0.0
            v1 += d[nd][n1]
            v1 += d[nd][n2]
            dist = 0.5 * (v1 - d12)
            d[parent][nd] = dist
            d[nd][parent] = dist
            new_x += dist
            x = xsub[nd]
            x += dist
            x -= d[n1][nd]
            x -= d[n2][nd]
            xsub[nd] = x
        xsub[parent] = new_x
        # Cpeeler.pyx:98-112
        if n_pool > 2:
            v1 = 0.5 * d12
            v4 = 1.0 / (2.0 * (float(n_pool) - 2.0)) * (xsub[n1] - xsub[n2])
            delta_f = v1 + v4
            delta_g = d12 - delta_f
            blens[n1] = delta_f
            blens[n2] = delta_g
        else:
            blens[n1] = d12 / 2.0
            blens[n2] = d12 / 2.0
        pool.append(parent)
        n_pool -= 1
    blens = np.maximum(blens, EPS)  # Cpeeler.pyx:124
    return peel, blens


# --------------------------------------------------------------------------- SoftSort
def _softmax_neg(a):
    """softmax over a vector whose maximum is known to be 0 is what the reference computes; do it
    the way torch does (subtract the max, exp, normalise)."""
    a = a - a.max()
    e = np.exp(a)
    return e / e.sum()


def soft_argmin_flat(s, tau):
    """Last-row closed form of soft.argmin on a flat vector s (soft.py:42-52).
    Returns the soft flat index (float64)."""
    s = np.asarray(s, dtype=np.float64)
    # sort(): rows of P_hat are softmax_j(-|s_j - sorted_i| / tau); last row <-> min(s)
    p = _softmax_neg(-np.abs(s - s.min()) / tau)
    c = p * np.cumsum(p)
    # second sort on -c: descending order puts min(-c) = -max(c) last
    p2 = _softmax_neg(-np.abs((-c) - (-c).min()) / tau)
    return float(np.sum(p2 * np.arange(s.size, dtype=np.float64)))


def soft_argmin(Q, tau):
    """soft.argmin(Q, tau) -> (soft_row, soft_col) as float64 (soft.py:38-53, unravel :7-10)."""
    Q = np.asarray(Q, dtype=np.float64)
    idx = soft_argmin_flat(Q.reshape(-1), tau)
    ncol = float(Q.shape[-1])
    return float(np.trunc(idx / ncol)), float(np.fmod(idx, ncol))


def soft_sort_dense(s, tau):
    """soft.sort for a 1-D s: the full n x n relaxed permutation matrix (soft.py:13-17)."""
    s = np.asarray(s, dtype=np.float64)
    srt = np.sort(s)[::-1]
    a = -np.abs(s[None, :] - srt[:, None]) / tau
    a = a - a.max(axis=1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(axis=1, keepdims=True)


def soft_argmin_dense(Q, tau):
    """Literal soft.argmin via the dense SoftSort matrices (small inputs only)."""
    Q = np.asarray(Q, dtype=np.float64)
    flat = Q.reshape(-1)
    P1 = soft_sort_dense(flat, tau)
    qf = P1[-1] * np.cumsum(P1[-1])
    P2 = soft_sort_dense(-qf, tau)
    idx = float(np.sum(P2[-1] * np.arange(flat.size, dtype=np.float64)))
    ncol = float(Q.shape[-1])
    return float(np.trunc(idx / ncol)), float(np.fmod(idx, ncol))


# --------------------------------------------------------------------------- soft NJ
def compute_Q(dists, mask):
    """peeler.compute_Q (peeler.py:223-248).  mask[i] True = node i not in the active set."""
    act = ~mask
    n_active = int(act.sum())
    A = np.outer(act, act)
    sum_pdm = np.sum(dists * A, axis=1, keepdims=True)
    Q = (n_active - 2) * dists - sum_pdm - sum_pdm.T
    Q[~A] = LARGE
    np.fill_diagonal(Q, LARGE)
    return 0.5 * (Q + Q.T)


def nj_soft(dists_leaves, tau, return_trace=False):
    """peeler.nj_torch(dists_leaves, tau) for tau is not None.
    Returns (peel int64 (S-1,3), blens float64 (2S-2,))."""
    D0 = np.asarray(dists_leaves, dtype=np.float64)
    S = D0.shape[0]
    N = 2 * S - 1
    dists = np.full((N, N), LARGE, dtype=np.float64)
    dists[:S, :S] = D0
    np.fill_diagonal(dists, 0.0)
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    blens = np.zeros(N - 1, dtype=np.float64)
    mask = np.zeros(N, dtype=bool)
    mask[S:] = True
    trace = []
    for int_i in range(S - 1):
        Q = compute_Q(dists, mask)
        right_f, left_f = soft_argmin(np.tril(Q), tau)
        right = int(np.round(right_f))
        left = int(np.round(left_f))
        # get_new_dist_soft (peeler.py:251-265)
        dist_f = dists[left].copy()
        dist_g = dists[right].copy()
        dist_fg = dist_f[right]
        dist_u = 0.5 * (dist_f + dist_g - dist_fg)
        act = ~mask
        n_active = max(float(act.sum()), 3.0)
        sum_pdm = np.sum(dists * np.outer(act, act), axis=-1)
        dist_pl = 0.5 * dist_fg + (sum_pdm[left] - sum_pdm[right]) / (2 * (n_active - 2))
        dist_pl = max(dist_pl, EPS)
        dist_u[left] = dist_pl
        parent = int_i + S
        peel[int_i] = (left, right, parent)
        dist_pr = max(dist_fg - dist_pl, EPS)
        dist_u[right] = dist_pr
        blens[left] = dist_pl
        blens[right] = dist_pr
        dist_u[parent] = 0.0
        dists[parent, :] = dist_u
        dists[:, parent] = dist_u
        mask[left] = True
        mask[right] = True
        mask[parent] = False
        if return_trace:
            trace.append((right_f, left_f))
    if return_trace:
        return peel, blens, trace
    return peel, blens


def nj_soft_torch(dists_leaves, tau):
    """Differentiable restatement of peeler.nj_torch in torch float64: selections are hard
    (no gradient, soft.py:56-61); distances/branch lengths carry gradient to dists_leaves."""
    import torch

    S = dists_leaves.shape[0]
    N = 2 * S - 1
    dists = torch.full((N, N), LARGE, dtype=torch.float64)
    dists[:S, :S] = dists_leaves
    dists = dists * (1.0 - torch.eye(N, dtype=torch.float64))
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    blens = torch.zeros(N - 1, dtype=torch.float64)
    mask = np.zeros(N, dtype=bool)
    mask[S:] = True
    for int_i in range(S - 1):
        Q = compute_Q(dists.detach().numpy().copy(), mask)
        right_f, left_f = soft_argmin(np.tril(Q), tau)
        right = int(np.round(right_f))
        left = int(np.round(left_f))
        act = torch.tensor(~mask)
        A = torch.outer(act, act).to(torch.float64)
        dist_f = dists[left]
        dist_g = dists[right]
        dist_fg = dist_f[right]
        dist_u = 0.5 * (dist_f + dist_g - dist_fg)
        n_active = max(float((~mask).sum()), 3.0)
        sum_pdm = torch.sum(dists * A, dim=-1)
        dist_pl = 0.5 * dist_fg + (sum_pdm[left] - sum_pdm[right]) / (2 * (n_active - 2))
        dist_pl = torch.clamp(dist_pl, min=EPS)
        dist_pr = torch.clamp(dist_fg - dist_pl, min=EPS)
        parent = int_i + S
        peel[int_i] = (left, right, parent)
        new = dist_u.clone()
        new[left] = dist_pl
        new[right] = dist_pr
        new[parent] = 0.0
        bl = blens.clone()
        bl[left] = dist_pl
        bl[right] = dist_pr
        blens = bl
        d2 = dists.clone()
        d2[parent, :] = new
        d2[:, parent] = new
        dists = d2
        mask[left] = True
        mask[right] = True
        mask[parent] = False
    return peel, blens

"""Oracle: the geodesic ("closest point to the origin") tree connector.  TEST INFRASTRUCTURE.

Restates in NumPy float64 (values) and torch float64 (gradient oracle, autograd):

* ``poincare.hyp_lca`` / ``hyp_lca_np`` and the pieces it is made of        dodonaphy/poincare.py:6-115
* ``utils.get_plca`` (pairwise depth-of-LCA matrix, row index = first arg)  dodonaphy/utils.py:141-171
* ``peeler.make_soft_peel_tips``   (VI / HMAP, gradients to the locations)  dodonaphy/peeler.py:14-70
* ``peeler.make_hard_peel_geodesic`` (MCMC, heap of candidate edges)        dodonaphy/peeler.py:73-152

Pinned by tests/golden/geodesic.npz (outputs of the reference itself, tests/golden/make_golden.py)
and the reference's own expectations in test/test_peeler.py:22-85,401-470.

Geometry: for two points a, b of the Poincare ball the "LCA" is the point of the geodesic through a
and b nearest the origin.  It is found by inverting the ball in the circle centred at a/|a|^2 (which
sends a to the origin and the geodesic to a straight line through the origin), mirroring the image of
the origin in that line, inverting back and halving the hyperbolic distance from the origin.
"""
import heapq

import numpy as np

from . import nj as _nj

MIN_NORM = 1e-15


# ------------------------------------------------------------------ numpy
def _invert(centre, x):
    """Circle inversion in the circle centred at ``centre`` that is orthogonal to the unit circle."""
    radius2 = np.sum(centre * centre, axis=-1, keepdims=True) - 1.0
    u = x - centre
    return radius2 / np.sum(u * u, axis=-1, keepdims=True) * u + centre


def _mirror(x, axis):
    """Mirror image of x in the line through the origin and ``axis``."""
    dot = np.sum(x * axis, axis=-1, keepdims=True)
    n2 = np.maximum(np.sum(axis * axis, axis=-1, keepdims=True), MIN_NORM)
    return 2 * (dot * axis / n2) - x


def hyp_lca_np(a, b, return_coord=True):
    centre = a / np.sum(np.power(a, 2), axis=-1, keepdims=True)
    b_img = _invert(centre, b)
    o_back = _invert(centre, _mirror(a, b_img))
    proj = o_back / (1.0 + np.sqrt(1 - np.sum(o_back * o_back, axis=-1, keepdims=True)))
    if return_coord:
        return proj
    return 2 * np.arctanh(np.linalg.norm(proj, axis=-1, keepdims=True))


def plca_matrix(locs):
    """Symmetric matrix of LCA depths; entry (i, j), i > j, is computed as lca(locs[i], locs[j])."""
    n = locs.shape[0]
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            out[i, j] = out[j, i] = hyp_lca_np(locs[i], locs[j], return_coord=False)[0]
    return out


def soft_peel_tips(leaf_locs, tau=1e-8, return_trace=False):
    """make_soft_peel_tips: (S, D) Poincare locations -> peel (S-1,3) int64, int_locs (S-1, D),
    blens (2S-2,).  The pair is chosen by soft.argmin (temperature 1e-8) over the negated depths of
    the active lower triangle; the merged node overwrites the COLUMN index, the row index retires."""
    locs = np.array(leaf_locs, dtype=np.float64)
    S, D = locs.shape
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    int_locs = np.zeros((S - 1, D))
    blens = np.zeros(2 * S - 2)
    retired = np.zeros(S, dtype=bool)
    node_of = np.arange(S)
    trace = []
    for step in range(S - 1):
        depth = plca_matrix(locs)
        live = np.outer(~retired, ~retired)
        masked = depth * live
        tril = np.tril(masked)
        big = masked.max() + 1.0
        score = -np.where(tril != 0, tril, -big)
        row_f, col_f = _nj.soft_argmin(score, tau)
        f, g = int(np.round(row_f)), int(np.round(col_f))
        new = hyp_lca_np(locs[f], locs[g])
        parent = S + step
        peel[step] = (node_of[f], node_of[g], parent)
        blens[node_of[f]] = hyp_lca_np(locs[f], new, return_coord=False)[0]
        blens[node_of[g]] = hyp_lca_np(locs[g], new, return_coord=False)[0]
        trace.append((f, g, locs[f].copy(), locs[g].copy()))
        locs[g] = new
        int_locs[step] = new
        node_of[g] = parent
        retired[f] = True
    if return_trace:
        return peel, int_locs, blens, trace
    return peel, int_locs, blens


# ------------------------------------------------------------------ torch (gradient oracle)
def hyp_lca_torch(a, b, return_coord=True):
    import torch

    def invert(centre, x):
        radius2 = torch.sum(centre * centre, dim=-1, keepdim=True) - 1.0
        u = x - centre
        return radius2 / torch.sum(u * u, dim=-1, keepdim=True) * u + centre

    centre = a / torch.sum(a * a, dim=-1, keepdim=True)
    b_img = invert(centre, b)
    dot = torch.sum(a * b_img, dim=-1, keepdim=True)
    n2 = torch.sum(b_img * b_img, dim=-1, keepdim=True).clamp_min(MIN_NORM)
    o_back = invert(centre, 2 * (dot * b_img / n2) - a)
    proj = o_back / (1.0 + torch.sqrt(1 - torch.sum(o_back * o_back, dim=-1, keepdim=True)))
    if return_coord:
        return proj
    return 2 * torch.arctanh(proj.norm(dim=-1, keepdim=True))


def soft_peel_tips_torch(leaf_locs, tau=1e-8):
    """Differentiable twin: the selections come from the NumPy restatement (they carry no gradient,
    soft.py:56-61); locations and branch lengths carry gradient to ``leaf_locs``."""
    import torch

    S, D = leaf_locs.shape
    peel, _, _, trace = soft_peel_tips(leaf_locs.detach().numpy(), tau, return_trace=True)
    locs = [leaf_locs[i] for i in range(S)]
    blens = [None] * (2 * S - 2)
    int_locs = []
    for step, (f, g, _, _) in enumerate(trace):
        new = hyp_lca_torch(locs[f], locs[g])
        blens[int(peel[step, 0])] = hyp_lca_torch(locs[f], new, return_coord=False)[0]
        blens[int(peel[step, 1])] = hyp_lca_torch(locs[g], new, return_coord=False)[0]
        locs[g] = new
        int_locs.append(new)
    return peel, torch.stack(int_locs), torch.stack(blens)


# ------------------------------------------------------------------ hard (heap) variant
class _Candidate:
    """Heap entry ordered by key only, like edge.Edge (dodonaphy/edge.py): equal keys keep the
    heap's own order."""

    __slots__ = ("key", "a", "b")

    def __init__(self, key, a, b):
        self.key, self.a, self.b = key, a, b

    def __lt__(self, other):
        return self.key < other.key


def _sheet_to_ball(z):
    return z[1:] / (1 + z[0])


def _ball_to_sheet(x):
    a = np.sum(x**2)
    eps = np.finfo(np.double).eps
    return np.concatenate(([(1 + a) / (1 - a)], 2 * x / (1 - a + eps)))


def hard_peel_geodesic(leaf_sheet):
    """make_hard_peel_geodesic: (S, D+1) hyperboloid points -> peel (S-1,3), int_locs (S-1, D+1).
    Repeatedly joins the live pair whose LCA is deepest; keys are negated depths in a binary heap that
    is filled row by row (i, then j < i) and extended with (old node, new node) pairs after each join."""
    leaf_sheet = np.asarray(leaf_sheet, dtype=np.float64)
    S, D1 = leaf_sheet.shape
    heap = []
    for i in range(S):
        for j in range(i):
            d = hyp_lca_np(_sheet_to_ball(leaf_sheet[i]), _sheet_to_ball(leaf_sheet[j]), return_coord=False)
            heapq.heappush(heap, _Candidate(-d, i, j))
    int_locs = np.zeros((S - 1, D1))
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    gone = [False] * (2 * S - 2)

    def point(n):
        return leaf_sheet[n] if n < S else int_locs[n - S]

    step = 0
    while step < S - 1:
        e = heapq.heappop(heap)
        if gone[e.a] or gone[e.b]:
            continue
        parent = S + step
        int_locs[step] = _ball_to_sheet(hyp_lca_np(_sheet_to_ball(point(e.a)), _sheet_to_ball(point(e.b))))
        peel[step] = (e.a, e.b, parent)
        gone[e.a] = gone[e.b] = True
        for n in range(parent):
            if not gone[n]:
                d = hyp_lca_np(_sheet_to_ball(point(n)), _sheet_to_ball(int_locs[step]), return_coord=False)
                heapq.heappush(heap, _Candidate(-d, n, parent))
        step += 1
    return peel, int_locs

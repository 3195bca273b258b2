"""Synthetic workload generator (host side, NumPy) for tests and bench.py.

Follows the recipe fixed in SURVEY.md section 8(d): tip locations ~ N(0, 0.05^2) in R^D, a
guide tree obtained by neighbour joining on their hyperboloid distances, an alignment simulated
*down that tree* under JC69 (i.i.d. random columns underflow the reference's unscaled likelihood
at >= 500 taxa), 2% fully ambiguous cells, integer pattern weights in [1, 4], GTR parameters
~ Dirichlet(1), and per-sample jitter of the locations so that every sample decodes to its own
topology and branch lengths.

This is data generation, not the hot path: nothing here is timed or used by the kernels.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def det_distances(S, D, seed):
    """Euclidean distances of seeded points built from elementwise IEEE operations only (no BLAS, no
    reductions whose order could depend on the host): the same bits on every machine, so a reference
    answer computed elsewhere stays valid for the regenerated input (tests/golden/configs.npz)."""
    x = np.random.default_rng(seed).normal(0.0, 1.0, (S, D))
    acc = np.zeros((S, S))
    for k in range(D):
        d = x[:, None, k] - x[None, :, k]
        acc = acc + d * d
    return np.sqrt(acc)


def embed_locations(S, D, rng, scale=0.05):
    return rng.normal(0.0, scale, size=(S, D))


def hyperboloid_pdm(x, curvature=-1.0):
    z = np.sqrt(np.sum(x * x, axis=1) + 1.0)
    H = x @ x.T - np.outer(z, z)
    H = np.minimum(H, -(1.0 + 1e-7))
    D = np.arccosh(-H) / np.sqrt(-curvature)
    np.fill_diagonal(D, 0.0)
    return D


def guide_tree(pdm):
    """Plain vectorised neighbour joining -> (peel (S-1,3) int64, blens (2S-2,)).
    Only used to get *a* binary tree with sensible branch lengths to simulate data on."""
    S = pdm.shape[0]
    N = 2 * S - 1
    D = np.zeros((N, N))
    D[:S, :S] = pdm
    active = list(range(S))
    peel = np.zeros((S - 1, 3), dtype=np.int64)
    blens = np.zeros(2 * S - 2)
    for it in range(S - 1):
        n = len(active)
        idx = np.array(active)
        sub = D[np.ix_(idx, idx)]
        r = sub.sum(axis=1)
        parent = S + it
        if n > 2:
            Q = (n - 2) * sub - r[:, None] - r[None, :]
            np.fill_diagonal(Q, np.inf)
            a, b = np.unravel_index(np.argmin(Q), Q.shape)
            if a > b:
                a, b = b, a
            dab = sub[a, b]
            da = 0.5 * dab + (r[a] - r[b]) / (2 * (n - 2))
            db = dab - da
        else:
            a, b = 0, 1
            dab = sub[a, b]
            da = db = 0.5 * dab
        i, j = idx[a], idx[b]
        new = 0.5 * (D[i, :] + D[j, :] - dab)
        D[parent, :] = new
        D[:, parent] = new
        D[parent, parent] = 0.0
        blens[i] = max(da, EPS)
        blens[j] = max(db, EPS)
        peel[it] = (i, j, parent)
        active = [k for k in active if k != i and k != j] + [parent]
    return peel, blens


def simulate_alignment(peel, blens, L, rng, ambiguous=0.02):
    """JC69 simulation down the tree.  Returns tip masks uint8 (S, L): bit j <-> state j (ACGT)."""
    S = peel.shape[0] + 1
    N = 2 * S - 1
    states = np.zeros((N, L), dtype=np.int8)
    root = int(peel[-1, 2])
    states[root] = rng.integers(0, 4, size=L)
    for left, right, parent in peel[::-1]:
        for child in (int(left), int(right)):
            t = blens[child]
            p_change = 0.75 * (1.0 - np.exp(-4.0 * t / 3.0))
            change = rng.random(L) < p_change
            shift = rng.integers(1, 4, size=L)
            states[child] = np.where(change, (states[parent] + shift) % 4, states[parent])
    masks = (1 << states[:S].astype(np.uint8)).astype(np.uint8)
    if ambiguous > 0:
        amb = rng.random((S, L)) < ambiguous
        masks = np.where(amb, np.uint8(0b1111), masks).astype(np.uint8)
    return masks


def gtr_params(rng):
    rates = rng.dirichlet(np.ones(6))
    freqs = rng.dirichlet(np.ones(4))
    return rates, freqs


def make_problem(S, L, B, D=5, seed=1001, jitter=0.01, scale=0.05, guide=None):
    """One synthetic configuration.  Returns a dict with
    x (S,D) base locations, xs (B,S,D) per-sample locations, masks (S,L) uint8, weights (L,) int64,
    guide_peel, guide_blens, rates(6), freqs(4).

    ``guide`` = (peel, blens) fixes the tree the alignment is simulated on.  The guide tree computed
    here goes through the host's BLAS and libm, and NJ's last joins are exact ties that follow the
    last bit -- so parity tests that compare with stored reference answers pass the committed guide
    tree of their seed (tests/golden/guides.npz): everything after it is elementwise and seeded."""
    rng = np.random.default_rng(seed)
    x = embed_locations(S, D, rng, scale)
    if guide is None:
        peel, blens = guide_tree(hyperboloid_pdm(x))
    else:
        peel, blens = np.asarray(guide[0], dtype=np.int64), np.asarray(guide[1], dtype=np.float64)
    masks = simulate_alignment(peel, blens, L, rng)
    weights = rng.integers(1, 5, size=L).astype(np.int64)
    rates, freqs = gtr_params(rng)
    xs = x[None, :, :] + jitter * rng.normal(size=(B, S, D))
    return dict(x=x, xs=xs, masks=masks, weights=weights, guide_peel=peel, guide_blens=blens,
                rates=rates, freqs=freqs, S=S, L=L, B=B, D=D, seed=seed)

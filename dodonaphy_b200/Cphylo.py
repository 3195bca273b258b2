"""NumPy-facing likelihood entry points of the MCMC path (dodonaphy/cython/Cphylo.pyx:9-51).

Same names and arguments; NumPy in, NumPy scalar out; the arithmetic runs in the CUDA pruning
kernel (forward-only variant).  No gradients, as in the reference.
"""
import numpy as np
import torch

from . import phylo as _phylo
from .engine import SubstModel


def JC69_p_t(branch_lengths):
    """(E,) -> (E,1,4,4) JC69 transition matrices (Cphylo.pyx:45-51)."""
    t = np.asarray(branch_lengths, dtype=np.float64).reshape(-1, 1, 1, 1)
    decay = np.exp(-4.0 / 3.0 * t)
    return np.where(np.eye(4, dtype=bool), 0.25 + 3.0 / 4.0 * decay, 0.25 - 0.25 * decay)


def compute_LL_np(partials, weights, peel, blen):
    """JC69 log-likelihood with equal frequencies (Cphylo.pyx:9-23): P(t) is formed on the device."""
    n_tips = len(peel) + 1
    eng = _phylo._engine_for(partials, weights, n_tips)
    peel_t = _phylo._as_peel_tensor(peel, n_tips, eng.device)
    b = torch.as_tensor(np.ascontiguousarray(blen, dtype=np.float64)).to(eng.device)
    ll, _ = eng.loglik(peel_t, b, SubstModel.jc69(), grad=False)
    return np.float64(ll.item())


def calculate_treelikelihood(partials, weights, post_indexing, mats, freqs):
    """Explicit-matrix variant (Cphylo.pyx:25-42); mats (E,K,4,4) or (E,4,4) ndarray."""
    n_tips = len(post_indexing) + 1
    eng = _phylo._engine_for(partials, weights, n_tips)
    peel_t = _phylo._as_peel_tensor(post_indexing, n_tips, eng.device)
    m = torch.as_tensor(np.ascontiguousarray(mats, dtype=np.float64))
    if m.shape[0] != eng.E or tuple(m.shape[-2:]) != (4, 4):
        raise ValueError(f"mats must have shape ({eng.E}, [K,] 4, 4)")
    out = eng.loglik_mats(peel_t, m.reshape(1, eng.E, -1, 4, 4).to(eng.device), np.asarray(freqs, dtype=np.float64))
    return np.float64(out["lnL"].item())


def compute_prior_gamma_dir_np(blen, aT=1.0, bT=0.1, a=1.0, c=1.0):
    """Gamma-Dirichlet(aT, bT, a, c) log prior of one tree's branch lengths (Cphylo.pyx:53-114;
    called per proposal from chain.py:129).  Host NumPy like the reference: 2S-2 numbers, off the
    device path.  The NumPy twin floors the tree length and every branch at machine epsilon (:105-107)
    where the torch twin clamps at 1e-4 (utils.py:258-260); both are kept as they are.  The batched
    device version is ``priors.gamma_dir_device``."""
    import math

    blen = np.asarray(blen, dtype=np.float64)
    eps = np.finfo(np.double).eps
    n_leaf = int(len(blen) / 2 + 1)
    tree_len = np.maximum(np.sum(blen), eps)
    tipb = np.sum(np.log(np.maximum(blen[:n_leaf], eps)))
    intb = np.sum(np.log(np.maximum(blen[n_leaf:], eps)))
    ln_prior = (a - 1) * tipb + (a * c - 1) * intb
    ln_prior = ln_prior + (aT - a * n_leaf - a * c * (n_leaf - 3)) * np.log(tree_len) - bT * tree_len
    ln_prior = (ln_prior + aT * np.log(bT) - math.lgamma(aT) + math.lgamma(a * n_leaf + a * c * (n_leaf - 3))
                - n_leaf * math.lgamma(a) - (n_leaf - 3) * math.lgamma(a * c))
    return ln_prior - np.sum(np.log(np.arange(n_leaf * 2 - 5, 0, -2)))


def compute_branch_lengths_np(S, peel, leaf_x, int_x, curvature=-1.0, matsumoto=False):
    """Branch lengths of a geodesic-connected tree from node locations (Cphylo.pyx:116-160): the
    hyperbolic distance from each child to its parent, floored at machine epsilon.

    Kept exactly as the reference evaluates it, including two quirks parity depends on: the parent
    (and internal child) row is ``int_x[node - S - 1]`` -- one row before the node's own, wrapping to
    the last row for the first internal node -- and both end points are lifted to the sheet twice
    (once here, once inside ``Chyp_np.hyperbolic_distance``)."""
    from . import Chyp_np

    leaf_x = np.asarray(leaf_x, dtype=np.float64)
    int_x = np.asarray(int_x, dtype=np.float64)
    peel = np.asarray(peel)[: S - 1]
    children = peel[:, :2].reshape(-1).astype(np.int64)
    parents = np.repeat(peel[:, 2].astype(np.int64), 2)
    is_leaf = children < S
    here = np.where(is_leaf[:, None], leaf_x[np.where(is_leaf, children, 0)], int_x[np.where(is_leaf, 0, children - S - 1)])
    up = int_x[parents - S - 1]
    # all 2S-2 branches as ONE launch of the pairs kernel (the reference loops over them in Python)
    hd = Chyp_np.hyperbolic_distance(Chyp_np.project_up_2d(here), Chyp_np.project_up_2d(up), curvature)
    if matsumoto:
        hd = np.log(np.cosh(hd))
    blens = np.zeros(2 * S - 2, dtype=np.float64)
    blens[children] = np.maximum(hd, np.finfo(np.double).eps)  # later rows win, as in the reference's loop
    return blens

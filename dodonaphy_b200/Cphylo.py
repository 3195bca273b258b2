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

"""Oracle: transition matrices and Felsenstein pruning.  TEST INFRASTRUCTURE (oracle/__init__.py).

Restates in NumPy float64 (and, for gradients, torch float64 + autograd -- the reference's
gradient IS torch autograd over these three lines):

* ``PhyloModel.JC69_p_t``      dodonaphy/phylomodel.py:87-94   (NumPy twin Cphylo.pyx:45-51)
* ``PhyloModel.compute_Q_martix`` / ``norm`` / ``GTR_p_t``   dodonaphy/phylomodel.py:96-128
* ``phylo.calculate_treelikelihood``   dodonaphy/phylo.py:132-139 (NumPy twin Cphylo.pyx:25-42)
* ``Cphylo.compute_LL_np``     dodonaphy/cython/Cphylo.pyx:9-23
* tip encoding of ``phylo.compress_alignment``   dodonaphy/phylo.py:16-58

Discrete-gamma rate categories (K>1) are NOT in the reference ("parity unpinned"):
``discrete_gamma_rates`` is Yang (1994) mean-rate discretisation and ``treelikelihood_gamma``
averages category likelihoods per pattern before the log.
"""
from collections import Counter

import numpy as np

DNA_MAP = {  # phylo.py:27-45 ; bit j of the mask <-> state j in order A,C,G,T
    "A": 0b0001, "C": 0b0010, "G": 0b0100, "T": 0b1000,
    "R": 0b0101, "Y": 0b1010, "M": 0b0011, "W": 0b1001, "S": 0b0110, "K": 0b1100,
    "B": 0b1110, "D": 0b1101, "H": 0b1011, "V": 0b0111,
    "N": 0b1111, "?": 0b1111, "-": 0b1111,
}


def mask_to_partial(masks):
    """uint8 masks (..., L) -> float64 partials (..., 4, L) with 0/1 entries."""
    masks = np.asarray(masks, dtype=np.uint8)
    bits = np.stack([(masks >> j) & 1 for j in range(4)], axis=-2)
    return bits.astype(np.float64)


def compress_alignment(sequences):
    """phylo.compress_alignment on a list of equal-length strings (one per taxon, taxon order kept).
    Returns (tip_partials list of (4,L) float64, weights int64 (L,), masks uint8 (S,L))."""
    count = Counter(list(zip(*sequences)))
    ordering = sorted(count.keys())
    patterns = list(zip(*ordering))
    weights = np.array([count[p] for p in ordering], dtype=np.int64)
    masks = np.array([[DNA_MAP.get(c.upper(), 0b1111) for c in row] for row in patterns], dtype=np.uint8)
    return [mask_to_partial(m) for m in masks], weights, masks


# ------------------------------------------------------------------ transition matrices
def jc69_p_t(blens):
    """blens (E,) or (E,K) -> (E,K,4,4) (K=1 if 1-D), phylomodel.py:87-94 / Cphylo.pyx:45-51."""
    b = np.asarray(blens, dtype=np.float64)
    if b.ndim == 1:
        b = b[:, None]
    d = b[..., None]
    a = 0.25 + 3.0 / 4.0 * np.exp(-4.0 / 3.0 * d)
    o = 0.25 - 0.25 * np.exp(-4.0 / 3.0 * d)
    m = np.concatenate((a, o, o, o, o, a, o, o, o, o, a, o, o, o, o, a), -1)
    return m.reshape(d.shape[0], d.shape[1], 4, 4)


def gtr_q(rates, freqs):
    """compute_Q_martix + norm (phylomodel.py:116-128).  rates: 6 exchangeabilities in
    triu order (01,02,03,12,13,23); freqs: 4."""
    rates = np.asarray(rates, dtype=np.float64)
    freqs = np.asarray(freqs, dtype=np.float64)
    iu = np.triu_indices(4, 1)
    Q = np.zeros((4, 4))
    Q[iu[0], iu[1]] = rates
    Q[iu[1], iu[0]] = rates
    Q = Q * freqs[None, :]
    Q = Q + np.diag(-np.sum(Q, axis=1))
    norm = -np.sum(np.diagonal(Q) * freqs)
    return Q / norm


def gtr_eigen(rates, freqs):
    """Eigen-system the reference forms in GTR_p_t (phylomodel.py:96-114):
    returns (U, lam, Uinv) with P(t) = U diag(exp(lam t)) Uinv, U = sqrt_pi_inv V,
    Uinv = V^-1 sqrt_pi, (lam, V) = eigh(sqrt_pi Q sqrt_pi_inv)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    Q = gtr_q(rates, freqs)
    sp = np.diag(np.sqrt(freqs))
    spi = np.diag(1.0 / np.sqrt(freqs))
    Ssym = sp @ Q @ spi
    lam, V = np.linalg.eigh(Ssym)
    return spi @ V, lam, np.linalg.inv(V) @ sp


def gtr_p_t(blens, rates, freqs):
    """GTR_p_t: blens (E,) -> (E,4,4); (E,K) -> (E,K,4,4)."""
    b = np.asarray(blens, dtype=np.float64)
    U, lam, Uinv = gtr_eigen(rates, freqs)
    ex = np.exp(lam * b[..., None])  # (...,4)
    return (U * ex[..., None, :]) @ Uinv


# ------------------------------------------------------------------ pruning
def treelikelihood(tip_partials, weights, post_indexing, mats, freqs):
    """phylo.calculate_treelikelihood (phylo.py:132-139).

    tip_partials: list of S arrays (4,L); mats: (E,4,4) or (E,K,4,4) indexed by child id.
    With a K axis the result is sum_k sum_p w_p log(pi . root_k[:,p]) exactly as the reference's
    broadcasting does."""
    S = len(tip_partials)
    partials = list(tip_partials) + [None] * (S - 1)
    for left, right, node in post_indexing:
        partials[node] = np.matmul(mats[left], partials[left]) * np.matmul(mats[right], partials[right])
    root = partials[post_indexing[-1][-1]]
    return float(np.sum(np.log(np.matmul(freqs, root)) * weights))


def compute_LL_np(tip_partials, weights, peel, blen):
    """Cphylo.compute_LL_np (Cphylo.pyx:9-23): JC69, freqs = 1/4."""
    return treelikelihood(tip_partials, weights, peel, jc69_p_t(blen), np.full(4, 0.25))


def discrete_gamma_rates(alpha, K):
    """Yang (1994) mean-rate discrete gamma with shape=rate=alpha: K category rates, mean 1."""
    if K == 1:
        return np.ones(1)
    from scipy.special import gammainc
    from scipy.stats import gamma

    cuts = gamma.ppf(np.arange(1, K) / K, a=alpha, scale=1.0 / alpha)
    cdf = np.concatenate([[0.0], gammainc(alpha + 1.0, cuts * alpha), [1.0]])
    return (cdf[1:] - cdf[:-1]) * K


def treelikelihood_gamma(tip_partials, weights, post_indexing, mats, freqs):
    """Rate-category mixture (new functionality): mats (E,K,4,4);
    lnL = sum_p w_p log( (1/K) sum_k pi . root_k[:,p] )."""
    S = len(tip_partials)
    K = mats.shape[1]
    partials = [np.broadcast_to(p, (K,) + p.shape) for p in tip_partials] + [None] * (S - 1)
    for left, right, node in post_indexing:
        partials[node] = np.matmul(mats[left], partials[left]) * np.matmul(mats[right], partials[right])
    root = partials[post_indexing[-1][-1]]
    site = np.mean(np.einsum("i,kil->kl", np.asarray(freqs, dtype=np.float64), root), axis=0)
    return float(np.sum(np.log(site) * weights))


# ------------------------------------------------------------------ torch twins (gradient oracle)
def treelikelihood_torch(tip_partials, weights, post_indexing, mats, freqs, gamma_mix=False):
    """Same three lines in torch float64 so autograd gives the reference gradient.
    tip_partials: list of (4,L) tensors; mats (E,4,4)/(E,K,4,4) tensor; returns 0-dim tensor."""
    import torch

    S = len(tip_partials)
    partials = list(tip_partials) + [None] * (S - 1)
    for left, right, node in post_indexing:
        partials[node] = torch.matmul(mats[left], partials[left]) * torch.matmul(mats[right], partials[right])
    root = partials[post_indexing[-1][-1]]
    if gamma_mix:
        site = torch.mean(torch.einsum("i,kil->kl", freqs, root), dim=0)
        return torch.sum(torch.log(site) * weights)
    return torch.sum(torch.log(torch.matmul(freqs, root)) * weights)


def jc69_p_t_torch(blens):
    import torch

    d = blens.unsqueeze(-1) if blens.dim() == 1 else blens
    d = d.unsqueeze(-1)
    a = 0.25 + 3.0 / 4.0 * torch.exp(-4.0 / 3.0 * d)
    b = 0.25 - 0.25 * torch.exp(-4.0 / 3.0 * d)
    return torch.cat((a, b, b, b, b, a, b, b, b, b, a, b, b, b, b, a), -1).reshape(d.shape[0], d.shape[1], 4, 4)


def gtr_p_t_torch(blens, rates, freqs):
    """GTR_p_t in torch (phylomodel.py:96-128) for blens (E,) or (E,K); rates (6,), freqs (4,)."""
    import torch

    iu = torch.triu_indices(4, 4, 1)
    Q = torch.zeros((4, 4), dtype=torch.float64)
    Q[iu[0], iu[1]] = rates
    Q[iu[1], iu[0]] = rates
    Q = Q * freqs.unsqueeze(0)
    Q = Q + torch.diag(-torch.sum(Q, dim=1))
    Q = Q / (-torch.sum(torch.diagonal(Q) * freqs))
    sp = freqs.sqrt().diag_embed()
    spi = (1.0 / freqs.sqrt()).diag_embed()
    lam, V = torch.linalg.eigh(sp @ Q @ spi)
    ex = torch.exp(lam * blens.unsqueeze(-1))
    return ((spi @ V) * ex.unsqueeze(-2)) @ (V.inverse() @ sp)


def treelikelihood_rescaled_torch(tip_partials, weights, post_indexing, mats, freqs, gamma_mix=False):
    """Underflow-safe twin of ``treelikelihood_torch``: every internal partial is divided by its
    per-pattern maximum and the logs are accumulated.  Mathematically identical to phylo.py:132-139;
    used as the oracle where the reference's unscaled product underflows."""
    import torch

    S = len(tip_partials)
    K = mats.shape[1] if mats.dim() == 4 else 1
    m4 = mats if mats.dim() == 4 else mats.unsqueeze(1)
    partials = [p.unsqueeze(0).expand(K, -1, -1) for p in tip_partials] + [None] * (S - 1)
    logs = [torch.zeros((K, p.shape[-1]), dtype=torch.float64) for p in tip_partials] + [None] * (S - 1)
    for left, right, node in post_indexing:
        v = torch.matmul(m4[left], partials[left]) * torch.matmul(m4[right], partials[right])
        mx = v.detach().max(dim=1, keepdim=True).values.clamp_min(1e-300)
        partials[node] = v / mx
        logs[node] = logs[left] + logs[right] + torch.log(mx.squeeze(1))
    root = post_indexing[-1][-1]
    site = torch.einsum("i,kil->kl", freqs, partials[root])
    lsite = torch.log(site) + logs[root]
    if gamma_mix:
        return torch.sum((torch.logsumexp(lsite, dim=0) - np.log(K)) * weights)
    return torch.sum(lsite * weights)

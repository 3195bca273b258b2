"""Branch-length priors for a whole batch of decoded trees, on the device the branch lengths live on
(removes the per-sample device->host copy of blens from the critical path).

``compute_prior_gamma_dir`` is the gamma-Dirichlet(aT, bT, a, c) prior of MrBayes (Rannala et al.
2012; Zhang et al. 2012) that the reference evaluates one tree at a time
(dodonaphy/base_model.py:475-527 with ``utils.LogDirPrior``, dodonaphy/utils.py:251-267; NumPy twin
dodonaphy/cython/Cphylo.pyx:53-114): tree length ~ Gamma(aT, bT), branch proportions ~ Dirichlet with
parameter a on the S external and a*c on the S-3 internal branches, uniform over topologies.
Works on (2S-2,) or (B, 2S-2) tensors, CPU or CUDA.
"""
import math

import torch


def log_dir_prior(blens, aT=1.0, bT=0.1, a=1.0, c=1.0):
    """Un-normalised log density of the branch lengths (LogDirPrior)."""
    b = blens if blens.dim() == 2 else blens.unsqueeze(0)
    n_leaf = b.shape[1] // 2 + 1
    tree_len = b.sum(dim=1)
    safe = torch.where((b > 0.0).all(dim=1, keepdim=True), b, b.clamp(min=1e-4))  # utils.py:258-260
    ext = torch.log(safe[:, :n_leaf]).sum(dim=1)
    inn = torch.log(safe[:, n_leaf:]).sum(dim=1)
    out = (a - 1.0) * ext + (a * c - 1.0) * inn + (aT - a * n_leaf - a * c * (n_leaf - 3)) * torch.log(tree_len) - bT * tree_len
    return out if blens.dim() == 2 else out[0]


def compute_prior_gamma_dir(blens, aT=1.0, bT=None, a=1.0, c=1.0):
    """Log prior of every tree of the batch (normalising constants and the 1/(2S-5)!! topology
    term included).  The reference's default bT is the float32 tensor torch.full((1,), 0.1)
    (base_model.py:476): its value 0.100000001490116 enters the float64 arithmetic as it is."""
    bT = torch.full((1,), 0.1) if bT is None else bT
    aT, bT, a, c = (float(v.reshape(-1)[0]) if isinstance(v, torch.Tensor) else float(v) for v in (aT, bT, a, c))
    n_leaf = blens.shape[-1] // 2 + 1
    const = (aT * math.log(bT) - math.lgamma(aT) + math.lgamma(a * n_leaf + a * c * (n_leaf - 3))
             - n_leaf * math.lgamma(a) - (n_leaf - 3) * math.lgamma(a * c))
    topo = sum(math.log(k) for k in range(2 * n_leaf - 5, 0, -2)) if n_leaf > 2 else 0.0
    return log_dir_prior(blens, aT, bT, a, c) + (const - topo)


def compute_prior_exponential(blens, rate=10.0):
    """Independent exponential prior on every branch (the VI alternative, vi.py:434-435)."""
    return (math.log(rate) - rate * blens).sum(dim=-1)

"""Substitution-model front end with the reference's ``PhyloModel`` surface for the hot path
(``get_transition_mats`` / ``JC69_p_t`` / ``GTR_p_t`` / ``compute_Q_martix`` / ``norm``,
dodonaphy/phylomodel.py:76-128, and the unconstrained ``_sub_rates`` (5) / ``_freqs`` (3) leaves of
:19-45 that the optimisers differentiate).

The likelihood kernels never consume P(t) tensors: they form P(t) on the device from
``device_model()``.  The tensor-returning methods exist so that callers written against the
reference (and its tests, test/test_phylo.py:276-300) keep working; they are 4x4 host arithmetic.
Priors, file output and the rest of the reference class are outside the hot path.
"""
import numpy as np
import torch
from torch.distributions.transforms import StickBreakingTransform

from .engine import SubstModel

_SIMPLEX = StickBreakingTransform()
_EYE4 = torch.eye(4, dtype=torch.float64)


_IU = np.triu_indices(4, 1)


def _stick_np(raw):
    """StickBreakingTransform (the map behind the reference's unconstrained leaves, phylomodel.py:19-45) and its
    Jacobian in NumPy: raw (n-1,) -> simplex y (n,), dy/draw (n, n-1).  z_i = sigmoid(raw_i - log(n-1-i)),
    y_i = z_i prod_{j<i} (1 - z_j), y_n = prod_j (1 - z_j)."""
    raw = np.asarray(raw, dtype=np.float64)
    n1 = raw.shape[0]
    z = 1.0 / (1.0 + np.exp(-(raw - np.log(n1 - np.arange(n1)))))
    rest = np.concatenate([[1.0], np.cumprod(1.0 - z)])  # rest[i] = prod_{j<i} (1 - z_j)
    y = np.concatenate([z * rest[:-1], rest[-1:]])
    J = np.zeros((n1 + 1, n1))
    dz = z * (1.0 - z)
    for j in range(n1):
        J[j, j] = dz[j] * rest[j]
        J[j + 1:, j] = -y[j + 1:] * z[j]  # d log(1 - z_j) / d raw_j = -z_j for every later factor
    return y, J


def rate_matrix_jacobian_np(raw_rates, raw_freqs):
    """Host side of the learnable-GTR chain rule, closed form: the normalised rate matrix Q (4,4), pi (4,) and
    the Jacobian (20, 5 + 3) of [vec(Q), pi] w.r.t. the unconstrained leaves (compute_Q_martix + norm,
    phylomodel.py:116-128, composed with the stick-breaking maps).  A few microseconds of NumPy instead of an
    autograd pass over sixteen numbers; pinned to ``rate_matrix_from_raw`` under autograd in the tests."""
    r, Jr = _stick_np(raw_rates)
    pi, Jp = _stick_np(raw_freqs)
    R = np.zeros((4, 4))
    R[_IU] = r
    R = R + R.T
    dR = np.zeros((4, 4, 6))  # d R / d r
    dR[_IU[0], _IU[1], np.arange(6)] = 1.0
    dR[_IU[1], _IU[0], np.arange(6)] = 1.0
    A = R * pi[None, :]  # off-diagonal part of the unnormalised matrix
    dA_r = dR * pi[None, :, None]
    dA_p = np.zeros((4, 4, 4))
    for j in range(4):
        dA_p[:, j, j] = R[:, j]

    def with_diag(M):  # rows sum to zero
        M = M.copy()
        idx = np.arange(4)
        M[idx, idx] = M[idx, idx] - M.sum(axis=1)
        return M

    Qt = with_diag(A)
    dQt_r = np.stack([with_diag(dA_r[:, :, k]) for k in range(6)], axis=-1)
    dQt_p = np.stack([with_diag(dA_p[:, :, k]) for k in range(4)], axis=-1)
    norm = -(np.diagonal(Qt) * pi).sum()
    dn_r = -np.einsum("ik,i->k", dQt_r[np.arange(4), np.arange(4), :], pi)
    dn_p = -np.einsum("ik,i->k", dQt_p[np.arange(4), np.arange(4), :], pi) - np.diagonal(Qt)
    Q = Qt / norm
    dQ_r = dQt_r / norm - Qt[:, :, None] * dn_r[None, None, :] / norm ** 2
    dQ_p = dQt_p / norm - Qt[:, :, None] * dn_p[None, None, :] / norm ** 2
    J = np.zeros((20, Jr.shape[1] + Jp.shape[1]))
    J[:16, :Jr.shape[1]] = dQ_r.reshape(16, 6) @ Jr
    J[:16, Jr.shape[1]:] = dQ_p.reshape(16, 4) @ Jp
    J[16:, Jr.shape[1]:] = Jp
    return Q, pi, J


class _SimplexLeaf:
    """A point on the simplex stored as an unconstrained leaf tensor (what Adam updates)."""

    def __init__(self, value, learnable):
        raw = _SIMPLEX.inv(torch.as_tensor(value, dtype=torch.float64)).clone().detach()
        self.raw = raw.requires_grad_(True) if learnable else raw

    def value(self):
        return _SIMPLEX(self.raw)


class PhyloModel:
    MODELS = ("JC69", "GTR")

    def __init__(self, name, sub_rates=None, freqs=None):
        self.name = name
        learn = name == "GTR"
        self.fix_sub_rates = self.fix_freqs = not learn
        if name == "JC69":
            sub_rates, freqs = torch.full((6,), 1.0 / 6.0), torch.full((4,), 0.25)
        elif name == "GTR":
            if sub_rates is None:
                sub_rates = torch.distributions.Dirichlet(torch.ones(6, dtype=torch.float64)).sample()
            if freqs is None:
                freqs = torch.distributions.Dirichlet(torch.ones(4, dtype=torch.float64)).sample()
        self._rates_leaf = _SimplexLeaf(sub_rates, learn) if name in self.MODELS else None
        self._freqs_leaf = _SimplexLeaf(freqs, learn) if name in self.MODELS else None

    def __eq__(self, other):
        return self.name == (other if isinstance(other, str) else other.name)

    __hash__ = None

    # unconstrained leaves, named as in the reference so params_optim dictionaries line up
    @property
    def _sub_rates(self):
        return self._rates_leaf.raw

    @_sub_rates.setter
    def _sub_rates(self, raw):
        self._rates_leaf.raw = raw

    @property
    def _freqs(self):
        return self._freqs_leaf.raw

    @_freqs.setter
    def _freqs(self, raw):
        self._freqs_leaf.raw = raw

    @property
    def sub_rates(self):
        return self._rates_leaf.value()

    @sub_rates.setter
    def sub_rates(self, value):
        if self.name != "JC69":
            self._rates_leaf = _SimplexLeaf(value, not self.fix_sub_rates)

    @property
    def freqs(self):
        return self._freqs_leaf.value()

    @freqs.setter
    def freqs(self, value):
        if self.name != "JC69":
            self._freqs_leaf = _SimplexLeaf(value, not self.fix_freqs)

    # Dirichlet(1) priors on the GTR exchangeabilities and frequencies (phylomodel.py:48-54,140-154)
    def compute_ln_prior_sub_rates(self):
        if self.name == "JC69":
            return torch.zeros(1, dtype=torch.float64)
        if self.name == "GTR":
            return torch.distributions.Dirichlet(torch.ones(6, dtype=torch.float64)).log_prob(self.sub_rates)
        raise RuntimeError(f"Model {self.name} has no prior on rates available.")

    def compute_ln_prior_freqs(self):
        if self.name == "JC69":
            return torch.zeros(1, dtype=torch.float64)
        if self.name == "GTR":
            return torch.distributions.Dirichlet(torch.ones(4, dtype=torch.float64)).log_prob(self.freqs)
        raise RuntimeError(f"Model {self.name} has no prior on freqs available.")

    # ------------------------------------------------------------------ P(t) as tensors
    def get_transition_mats(self, blens):
        if self.name == "JC69":
            return self.JC69_p_t(blens)
        if self.name == "GTR":
            return self.GTR_p_t(blens)
        raise ValueError(f"Model {self.name} has no transition matrix implementation.")

    @staticmethod
    def JC69_p_t(branch_lengths):
        """(E,) or (E,K) -> (E,K,4,4): diagonal 1/4 + 3/4 e^{-4t/3}, off-diagonal 1/4 - 1/4 e^{-4t/3}."""
        branch_lengths = branch_lengths.to(torch.float64)
        t = branch_lengths if branch_lengths.dim() > 1 else branch_lengths.unsqueeze(-1)
        decay = torch.exp(-4.0 / 3.0 * t)[..., None, None]
        same = 0.25 + 3.0 / 4.0 * decay
        diff = 0.25 - 0.25 * decay
        return torch.where(_EYE4.bool(), same, diff)

    def compute_Q_martix(self):
        """Rate matrix: symmetric exchangeabilities (upper-triangle order 01,02,03,12,13,23) times
        the column frequencies, rows summing to zero, one expected substitution per unit time."""
        r, pi = self.sub_rates, self.freqs
        iu = torch.triu_indices(4, 4, 1)
        R = torch.zeros((4, 4), dtype=torch.float64).index_put((iu[0], iu[1]), r)
        Q = (R + R.T) * pi
        Q = Q - torch.diag(Q.sum(dim=1))
        return Q / self.norm(Q)

    def norm(self, Q):
        return -(torch.diagonal(Q, dim1=-2, dim2=-1) * self.freqs).sum(-1)

    def _eigensystem(self):
        pi = self.freqs
        root, inv_root = torch.diag(pi.sqrt()), torch.diag(pi.rsqrt())
        lam, V = torch.linalg.eigh(root @ self.compute_Q_martix() @ inv_root)
        return inv_root @ V, lam, torch.linalg.inv(V) @ root

    @staticmethod
    def rate_matrix_from_raw(raw_rates, raw_freqs):
        """Differentiable map from the unconstrained leaves (5), (3) to the normalised GTR rate matrix (4,4)
        and the frequencies (4,): compute_Q_martix + norm (phylomodel.py:116-128) on the host."""
        r, pi = _SIMPLEX(raw_rates), _SIMPLEX(raw_freqs)
        iu = torch.triu_indices(4, 4, 1)
        R = torch.zeros((4, 4), dtype=torch.float64).index_put((iu[0], iu[1]), r)
        Q = (R + R.T) * pi
        Q = Q - torch.diag(Q.sum(dim=1))
        return Q / -(torch.diagonal(Q) * pi).sum(), pi

    @staticmethod
    def transition_mats_from_raw(raw_rates, raw_freqs, blens, cat_rates=None):
        """Differentiable P(t) for a batch: unconstrained leaves (5), (3) on the host, blens (B,E) on
        any device, K category rates -> (B,E,K,4,4) on blens' device, and pi (4,) on the host.
        The 4x4 eigen-system is host arithmetic; the (B,E,K,4,4) tensor is formed where blens live."""
        r, pi = _SIMPLEX(raw_rates), _SIMPLEX(raw_freqs)
        iu = torch.triu_indices(4, 4, 1)
        R = torch.zeros((4, 4), dtype=torch.float64).index_put((iu[0], iu[1]), r)
        Q = (R + R.T) * pi
        Q = Q - torch.diag(Q.sum(dim=1))
        Q = Q / -(torch.diagonal(Q) * pi).sum()
        root, inv_root = torch.diag(pi.sqrt()), torch.diag(pi.rsqrt())
        lam, V = torch.linalg.eigh(root @ Q @ inv_root)
        dev = blens.device
        U, Uinv, lam = (inv_root @ V).to(dev), (torch.linalg.inv(V) @ root).to(dev), lam.to(dev)
        cat = torch.ones(1, dtype=torch.float64, device=dev) if cat_rates is None else torch.as_tensor(cat_rates, dtype=torch.float64, device=dev)
        t = blens.unsqueeze(-1) * cat  # (B,E,K)
        return (U * torch.exp(t.unsqueeze(-1) * lam).unsqueeze(-2)) @ Uinv, pi

    def GTR_p_t(self, branch_lengths):
        """(E,) -> (E,4,4); (E,K) -> (E,K,4,4):  U diag(exp(lam t)) U^-1."""
        U, lam, Uinv = self._eigensystem()
        growth = torch.exp(branch_lengths.to(torch.float64).unsqueeze(-1) * lam)
        return (U * growth.unsqueeze(-2)) @ Uinv

    # ------------------------------------------------------------------ what the kernels consume
    def device_model(self):
        if self.name == "JC69":
            return SubstModel.jc69()
        if self.name == "GTR":
            return SubstModel.gtr(self.sub_rates.detach().cpu().numpy(), self.freqs.detach().cpu().numpy())
        raise ValueError(f"Model {self.name} has no transition matrix implementation.")

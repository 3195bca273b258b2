"""Batched variational step: the caller side of the hot path on the VI route.

The reference's ``DodonaphyVI`` (dodonaphy/vi.py) evaluates its importance-weighted ELBO with a
serial Python loop -- ``elbo_siwae`` (:315-341) calls ``calculate_elbo`` (:161-202) once per
(importance sample, mixture component), each of which samples locations (``rsample_Euclid``
:440-465), decodes a tree (``connect`` :467-491), evaluates lnL + prior (``get_loss`` :426-438) and a
Jacobian log-determinant by 2S-2 autograd passes (:376-387).  Here all importance x boosts samples
of a step are drawn at once and go through the kernels as ONE batch; the objective
(ln p + ln prior - ln q + ln|J|, logsumexp over samples and mixture, :334-341) is the reference's.
Only the batching glue lives here: initial embeddings (Hydra+), tree files, logs and checkpoints of
the reference class are outside the path.
"""
import math

import torch

from . import priors
from .base_model import BaseModel
from .engine import GammaDirPrior


class DodonaphyVI(BaseModel):
    def __init__(self, partials, weights, dim, embedder="up", connector="nj", curvature=-1.0, soft_temp=1e-8,
                 tip_labels=None, n_boosts=1, model_name="JC69", freqs=None, prior_fn="gammadir",
                 gamma_shape=None, n_rate_categories=1, include_jacobian=True):
        if soft_temp is None or not soft_temp > 0.0:
            raise ValueError("variational inference needs a positive soft-NJ temperature")
        if prior_fn not in ("exponential", "gammadir"):
            raise AssertionError("Invalid prior specified.")
        super().__init__("vi", partials, weights, dim, soft_temp=soft_temp, embedder=embedder, connector=connector,
                         curvature=torch.tensor(curvature, dtype=torch.float64), tip_labels=tip_labels,
                         model_name=model_name, freqs=freqs, gamma_shape=gamma_shape, n_rate_categories=n_rate_categories)
        self.n_boosts, self.prior_fn, self.include_jacobian = n_boosts, prior_fn, include_jacobian
        self.params_optim = {}
        self.mix_weights = torch.full((n_boosts,), 1.0 / n_boosts, dtype=torch.float64)

    def set_params_optim(self, leaf_mu, leaf_sigma):
        """leaf_mu, leaf_sigma: (n_boosts, S, D) or (S, D); sigma is the LOG variance, as in the
        reference (the covariance is diag(exp(leaf_sigma)), vi.py:170-172)."""
        mu = torch.as_tensor(leaf_mu, dtype=torch.float64).reshape(self.n_boosts, self.S, self.D)
        sg = torch.as_tensor(leaf_sigma, dtype=torch.float64).reshape(self.n_boosts, self.S, self.D)
        self.params_optim = {"leaf_mu": mu.clone().requires_grad_(True), "leaf_sigma": sg.clone().requires_grad_(True),
                             "curvature": self._curvature}
        if not self.phylomodel.fix_sub_rates:
            self.params_optim["sub_rates"] = self.phylomodel._sub_rates
            self.params_optim["freqs"] = self.phylomodel._freqs

    def set_params_optim_random(self):
        self.set_params_optim(torch.randn(self.n_boosts, self.S, self.D, dtype=torch.float64) * 0.05,
                              torch.full((self.n_boosts, self.S, self.D), -6.0, dtype=torch.float64))

    # ------------------------------------------------------------------------------------------
    def rsample(self, importance, noise=None):
        """All samples of a step: locations (importance, n_boosts, S, D) and their log density."""
        mu, sg = self.params_optim["leaf_mu"], self.params_optim["leaf_sigma"]
        z = torch.randn((importance,) + tuple(mu.shape), dtype=torch.float64) if noise is None else noise
        std = torch.exp(0.5 * sg)
        x = mu + std * z
        log_q = (-0.5 * z * z - 0.5 * sg - 0.5 * math.log(2.0 * math.pi)).sum(dim=(-1, -2))
        return x, log_q

    def elbo_terms(self, x):
        """ln p, ln prior, ln|J| for a batch of locations (B, S, D)."""
        if self.prior_fn == "gammadir":  # evaluated in the epilogue of the soft-NJ kernel launch
            ln_p, peel, blens, ln_prior = self.log_likelihood_batch(x, prior=GammaDirPrior())
        else:
            ln_p, peel, blens = self.log_likelihood_batch(x)
            ln_prior = priors.compute_prior_exponential(blens, rate=10.0)
        if self.include_jacobian:
            hot = self._hotpath(defer_model=True)  # (exists: the likelihood call above made it)
            xd, peel_d, flags_d = hot.last_decode  # the trees just decoded for the likelihood: not decoded again
            _, jac = hot.blens_jacobian(xd, peel_d, flags_d, logdet=True)
            jac = torch.where(torch.isfinite(jac), jac, torch.zeros_like(jac)).to(x.device)  # vi.py:199-201
        else:
            jac = torch.zeros_like(ln_p)
        return ln_p, ln_prior, jac.detach(), peel, blens

    def elbo_siwae(self, importance=1, noise=None):
        x, log_q = self.rsample(importance, noise)
        ln_p, ln_prior, jac, _, _ = self.elbo_terms(x.reshape(-1, self.S, self.D))
        ln_elbos = (ln_p + ln_prior + jac).reshape(importance, self.n_boosts) - log_q
        mixture_logit = torch.log_softmax(self.mix_weights, dim=0)
        loss_n = torch.logsumexp(-math.log(importance) + mixture_logit + ln_elbos, dim=1)
        return loss_n.mean(dim=0)

    def learn(self, epochs=100, importance_samples=1, lr=1e-1):
        """Adam + 1/sqrt(epoch+1) schedule on -ELBO (vi.py:228-249); returns the ELBO trace."""
        params = [p for p in self.params_optim.values() if isinstance(p, torch.Tensor) and p.requires_grad]
        opt = torch.optim.Adam(params, lr=lr)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 1.0 / math.sqrt(e + 1))
        trace = []
        for _ in range(epochs):
            opt.zero_grad()
            loss = -self.elbo_siwae(importance_samples)
            loss.backward()
            opt.step()
            sched.step()
            trace.append(-loss.item())
        return trace

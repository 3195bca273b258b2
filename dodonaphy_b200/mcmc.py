"""Batched MCMC sweep: the caller side of the hot path on the MCMC route.

The reference's ``DodonaphyMCMC.learn`` (dodonaphy/mcmc.py:115-164) walks its chains one after the
other; each ``Chain.evolve`` (dodonaphy/chain.py:223-258, :403-434) proposes locations, decodes a
tree (``connect`` :171-221, Chyp_np.get_pdm -> Cpeeler.nj_np) and evaluates
``Cphylo.compute_LL_np`` (``get_loss`` :145-169).  Here one generation = propose for every chain,
evaluate all proposals as ONE batch on the device, accept/reject per chain on the host, then the
MC3 swaps between neighbouring temperatures (mcmc.py:270-298).  Proposal adaptation (RAM / AM /
tuning), priors other than the gamma-Dirichlet, tree files and logs are outside the path.
"""
import numpy as np

from .base_model import BaseModel
from .engine import GammaDirPrior


class DodonaphyMCMC:
    def __init__(self, partials, weights, dim, n_chains=4, connector="nj", embedder="up", step_scale=0.01,
                 curvature=-1.0, matsumoto=False, heat=0.1, seed=None, model_name="JC69"):
        if connector != "nj" or embedder != "up":
            raise NotImplementedError("the batched sweep covers connector='nj', embedder='up'")
        self.model = BaseModel("mcmc", partials, weights, dim, soft_temp=None, curvature=curvature, embedder=embedder,
                               connector=connector, require_grad=False, matsumoto=matsumoto, model_name=model_name)
        self.n_chains, self.step_scale = n_chains, step_scale
        self.temps = 1.0 / (1.0 + heat * np.arange(n_chains))  # chain_temp of mcmc.py:300-312
        self.rng = np.random.default_rng(seed)
        self.x = None
        self._prior = GammaDirPrior(variant="numpy", grad=False)

    def initialise_chains(self, leaf_x):
        self.x = np.repeat(np.asarray(leaf_x, dtype=np.float64)[None], self.n_chains, axis=0)
        # Cphylo.compute_prior_gamma_dir_np (chain.py:129) in the hard-NJ kernel's epilogue
        self.ln_p, self.peel, self.blens, self.ln_prior = self.model.log_likelihood_batch_np(self.x, prior=self._prior)
        self.accepted = np.zeros(self.n_chains, dtype=np.int64)
        self.iterations = 0

    def sweep(self):
        """One generation for all chains (Metropolis with a symmetric Gaussian proposal)."""
        # chain.py:75,455-458: the reference keeps cov = I * step_scale and proposes x + cov @ u with
        # u ~ N(0, I), i.e. a displacement of standard deviation step_scale per coordinate (its `cov` is
        # applied as a matrix, not as a covariance).  The Jacobian term of its acceptance ratio is
        # identically zero on this route (chain.py:414 sets log_abs_det_jacobian = 0.0).
        prop = self.x + self.step_scale * self.rng.standard_normal(self.x.shape)
        ln_p, peel, blens, ln_prior = self.model.log_likelihood_batch_np(prop, prior=self._prior)
        ln_r = self.temps * ((ln_p + ln_prior) - (self.ln_p + self.ln_prior))
        take = np.log(self.rng.random(self.n_chains)) < ln_r
        self.x[take], self.ln_p[take], self.ln_prior[take] = prop[take], ln_p[take], ln_prior[take]
        self.peel[take], self.blens[take] = peel[take], blens[take]
        self.accepted += take
        self.iterations += 1
        return take

    def swap(self):
        """Exchange the states of two neighbouring temperatures (mcmc.py:270-298)."""
        if self.n_chains < 2:
            return False
        i = int(self.rng.integers(0, self.n_chains - 1))
        j = i + 1
        post = self.ln_p + self.ln_prior
        ln_r = (self.temps[i] - self.temps[j]) * (post[j] - post[i])
        if np.log(self.rng.random()) < ln_r:
            for arr in (self.x, self.ln_p, self.ln_prior, self.peel, self.blens):
                arr[[i, j]] = arr[[j, i]]
            return True
        return False

    def learn(self, epochs, swap_period=10):
        trace = []
        for e in range(epochs):
            self.sweep()
            if swap_period and (e + 1) % swap_period == 0:
                self.swap()
            trace.append(float(self.ln_p[0]))
        return trace

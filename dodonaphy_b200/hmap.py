"""Maximum-a-posteriori embedding: the ``HMAP`` caller of the hot path (dodonaphy/hmap.py).

What is kept is the part of the reference class that drives the path -- ``get_locs`` (:302-308),
``connect`` (:310-334), ``compute_loss`` (:185-192), ``compute_ln_likelihood`` (:360-372),
``compute_ln_prior`` / ``compute_ln_tree_prior`` (:374-397), ``record_if_best`` (:236-247) and the
Adam + 1/(epoch+1)^(1/4) schedule of ``learn`` (:121-176).  Embedding initialisation from files or
Hydra+, logging, plots and checkpoints are host code outside the path.

One evaluation of the loss is one pass of the fused pipeline with a batch of one:
locations -> distances -> soft NJ -> pruning value + d lnL/d blens -> NJ / distance backward, i.e.
ten kernel launches instead of ~(S-1) x (soft-sort + three torch ops) autograd nodes.  With a
learnable GTR model the pruning kernel additionally returns d lnL/d P and d lnL/d pi, from which the
rate and frequency leaves receive gradient (``base_model._BatchedLogLikModel``).
"""
import numpy as np
import torch

from . import priors
from .base_model import BaseModel
from .engine import GammaDirPrior


class HMAP(BaseModel):
    def __init__(self, partials, weights, dim, soft_temp, loss_fn="likelihood", path_write=None, curvature=-1.0,
                 prior="None", tip_labels=None, matsumoto=False, connector="nj", peel=None, normalise_leaves=False,
                 model_name="JC69", freqs=None, embedder="up"):
        if loss_fn != "likelihood":
            raise NotImplementedError(f"loss_fn {loss_fn!r} is a research loss outside the likelihood path")
        if prior not in ("None", "normal", "uniform", "gammadir", "birthdeath"):
            raise ValueError(f"unknown prior {prior!r}")
        super().__init__("hmap", partials, weights, dim=dim, soft_temp=soft_temp, connector=connector,
                         curvature=curvature, loss_fn=loss_fn, tip_labels=tip_labels, model_name=model_name,
                         freqs=freqs, embedder=embedder, matsumoto=matsumoto)
        self.path_write = path_write
        self.normalise_leaves = normalise_leaves
        self.current_epoch = 0
        self.prior = prior
        self.peel = peel
        self.params_optim = {}
        self.init_model_params()
        self.best_posterior = torch.tensor(-np.inf, dtype=torch.float64)

    def init_model_params(self):
        """Evolutionary-model leaves and the curvature leaf join the optimiser (base_model.py:198-205)."""
        if not self.phylomodel.fix_sub_rates:
            self.params_optim["sub_rates"] = self.phylomodel._sub_rates
        if not self.phylomodel.fix_freqs:
            self.params_optim["freqs"] = self.phylomodel._freqs
        if self.require_grad:
            self.params_optim["curvature"] = self._curvature

    def init_embedding_params(self, locations):
        """Start from explicit tip locations (S, D).  (The reference reads them from a file or embeds
        a distance matrix with Hydra+, hmap.py:65-119 -- both off the path.)"""
        locs = torch.as_tensor(locations, dtype=torch.float64).clone().detach()
        if locs.shape != (self.S, self.D):
            raise ValueError(f"locations must have shape ({self.S}, {self.D})")
        if self.normalise_leaves:
            radius = locs.norm(dim=1).mean().reshape(1)
            self.params_optim["radius"] = radius.requires_grad_(True)
            self.params_optim["directionals"] = (locs / locs.norm(dim=1, keepdim=True)).requires_grad_(True)
        else:
            self.params_optim["leaf_loc"] = locs.requires_grad_(True)

    def get_locs(self):
        if self.normalise_leaves:
            return self.params_optim["radius"] * self.params_optim["directionals"]
        return self.params_optim["leaf_loc"]

    # ------------------------------------------------------------------ one evaluation
    def connect(self, get_pdm=False):
        """Decode the current locations into a tree (one-tree interface of hmap.py:310-334)."""
        if self.connector not in ("nj", "geodesics"):
            if self.connector == "fix":
                raise NotImplementedError("connector 'fix' embeds internal nodes; only 'nj' and 'geodesics' are on the path")
            raise ValueError(f"Connection must be one of 'nj', 'geodesics', 'fix'. Got {self.connector}")
        peel, blens, pdm = super().connect(self.get_locs())
        return (peel, blens, pdm) if get_pdm else (peel, blens)

    def compute_loss(self, fused=True):
        """-(ln prior + lnL) of the current embedding.  ``fused=False`` takes the reference's
        call-by-call route (connect -> compute_LL) instead of the single batched pass."""
        if self.connector not in ("nj", "geodesics"):
            self.connect()
        self._fused_tree_prior = None
        if fused and self.prior == "gammadir":  # the tree prior comes out of the same soft-NJ launch
            ll, peel, blens, lp = self.log_likelihood_batch(self.get_locs().unsqueeze(0), prior=GammaDirPrior())
            self.peel, self.blens, self.ln_p, self._fused_tree_prior = peel[0].cpu().numpy(), blens[0], ll[0], lp[0]
        elif fused:
            ll, peel, blens = self.log_likelihood_batch(self.get_locs().unsqueeze(0))
            self.peel, self.blens, self.ln_p = peel[0].cpu().numpy(), blens[0], ll[0]
        else:
            self.peel, self.blens, self.pdm = self.connect(get_pdm=True)
            self.ln_p = self.compute_ln_likelihood()
        self.ln_prior = self.compute_ln_prior()
        return -self.ln_prior - self.ln_p

    def compute_ln_likelihood(self):
        return self.compute_LL(self.peel, self.blens)

    def compute_ln_prior(self):
        if self.prior == "None":
            return torch.zeros(1, dtype=torch.float64)
        return (self.phylomodel.compute_ln_prior_sub_rates() + self.phylomodel.compute_ln_prior_freqs()
                + self.compute_ln_tree_prior())

    def compute_ln_tree_prior(self):
        if self.prior == "None":
            return torch.zeros(1, dtype=torch.float64)
        if self.prior == "normal":
            return self.compute_prior_normal(self.get_locs())
        if self.prior == "uniform":
            return self.compute_prior_unif(self.get_locs(), scale=10.0)
        if self.prior == "gammadir":
            if getattr(self, "_fused_tree_prior", None) is not None:
                return self._fused_tree_prior
            return priors.compute_prior_gamma_dir(self.blens)
        raise ValueError("Birth death model implementation isn't differentiable. "
                         "It cannot be used for gradient based map estimation.")

    @staticmethod
    def compute_prior_normal(locations, scale=0.01):
        """Isotropic normal N(0, scale I) on the flattened locations (base_model.py:404-430)."""
        x = torch.as_tensor(locations, dtype=torch.float64).flatten()
        return -0.5 * (x * x).sum() / scale - 0.5 * x.numel() * np.log(2.0 * np.pi * scale)

    @staticmethod
    def compute_prior_unif(locations, scale=1.0):
        """Independent U(-scale, scale) on every coordinate (base_model.py:433-458)."""
        x = torch.as_tensor(locations, dtype=torch.float64).flatten()
        inside = ((x >= -scale) & (x < scale)).all()
        return torch.where(inside, torch.tensor(-x.numel() * np.log(2.0 * scale), dtype=torch.float64),
                           torch.tensor(-np.inf, dtype=torch.float64))

    # ------------------------------------------------------------------ optimisation
    def record_if_best(self):
        """Keep a detached snapshot of the best state seen so far under the reference's attribute
        names (``best_posterior``, ``best_ln_p``, ``best_locs`` ..., hmap.py:236-247)."""
        if self.current_epoch != 0 and not bool(self.loss < -self.best_posterior):
            return
        tensors = dict(posterior=-self.loss.reshape(()), ln_p=self.ln_p, ln_prior=self.ln_prior, blens=self.blens,
                       freqs=self.phylomodel.freqs, sub_rates=self.phylomodel.sub_rates,
                       curvature=torch.as_tensor(self.curvature), locs=self.get_locs())
        for name, value in tensors.items():
            setattr(self, f"best_{name}", value.detach().clone())
        self.best_peel, self.best_epoch = self.peel, self.current_epoch

    def learn(self, epochs, learn_rate, save_locations=False, start=""):
        """Adam with the reference's 1/(epoch+1)^(1/4) learning-rate decay; returns the history of
        log posteriors (one entry before the first step, one after every step)."""
        self.loss = self.compute_loss()
        history = [float(self.ln_p.detach()) + float(self.ln_prior.detach())]
        self.record_if_best()
        optimizer = torch.optim.Adam(params=list(self.params_optim.values()), lr=learn_rate)
        scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: 1.0 / (epoch + 1.0) ** 0.25)

        def closure():
            optimizer.zero_grad()
            self.loss = self.compute_loss()
            self.loss.sum().backward()
            return self.loss

        for i in range(1, epochs + 1):
            self.current_epoch = i
            optimizer.step(closure)
            scheduler.step()
            self.record_if_best()
            history.append(float(self.ln_p.detach()) + float(self.ln_prior.detach()))
        return history

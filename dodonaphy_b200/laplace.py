"""Laplace approximation around a MAP embedding: the ``Laplace`` caller of the hot path
(dodonaphy/laplace.py).

``get_ln_posterior`` (laplace.py:9-25) is locations -> distances -> soft NJ -> lnL + gamma-Dirichlet
prior.  The reference differentiates that function twice with ``torch.autograd.functional.hessian``
(:40), i.e. S*D further backward passes through the soft-sort graph.  Here the Hessian is taken as
central differences of the ANALYTIC gradient, and all 2*S*D displaced embeddings go through the
pipeline as one batch -- the batched path is exactly the shape this needs.  With the topology fixed
by the (non-differentiable) one-hot choice the posterior is smooth, and the differences agree with
the autograd Hessian to ~1e-6 relative (step 1e-5; tested against the oracle's double backward).
"""
import torch

from . import priors
from .hmap import HMAP


class Laplace(HMAP):
    def get_ln_posterior(self, leaf_loc_flat):
        """lnL + gamma-Dirichlet prior of the tree decoded from flat locations (S*D,) or a batch
        (B, S*D); carries gradient to the locations."""
        flat = torch.as_tensor(leaf_loc_flat, dtype=torch.float64)
        locs = flat.reshape(-1, self.S, self.D)
        ll, _, blens = self.log_likelihood_batch(locs)
        out = ll + priors.compute_prior_gamma_dir(blens)
        return out[0] if flat.dim() == 1 else out

    def ln_posterior_grad(self, locs_flat, return_peel=False):
        """Values (B,) and analytic gradients (B, S*D) of the log posterior at B embeddings."""
        x = torch.as_tensor(locs_flat, dtype=torch.float64).detach().clone().requires_grad_(True)
        ll, peel, blens = self.log_likelihood_batch(x.reshape(-1, self.S, self.D))
        val = ll + priors.compute_prior_gamma_dir(blens)
        (g,) = torch.autograd.grad(val.sum(), x)
        return (val.detach(), g, peel) if return_peel else (val.detach(), g)

    def hessian(self, mean=None, step=1e-5, max_shrink=3):
        """(S*D, S*D) Hessian of the log posterior at ``mean`` (default: current locations).

        The reference's autograd Hessian is taken at a FIXED join selection (the one-hot choice passes no
        gradient, soft.py:56-61).  Central differences only reproduce that while every displaced
        embedding decodes to the centre's topology: the peels of the whole displaced batch are compared
        with the centre's, and the step is shrunk (x0.1, at most ``max_shrink`` times) when a near-tied
        join flips inside +-step; if it still flips the Hessian is not defined there and this raises."""
        mean = self.get_locs().detach().reshape(-1) if mean is None else torch.as_tensor(mean, dtype=torch.float64).reshape(-1)
        n = mean.numel()
        for _ in range(max_shrink + 1):
            shifts = step * torch.eye(n, dtype=torch.float64)
            _, g, peel = self.ln_posterior_grad(torch.cat((mean.reshape(1, -1), mean + shifts, mean - shifts)), return_peel=True)
            if bool((peel[1:] == peel[:1]).all()):
                H = (g[1:n + 1] - g[n + 1:]) / (2.0 * step)
                return 0.5 * (H + H.T)
            step *= 0.1
        raise ValueError("Laplace: the decoded topology changes within the finite-difference step around the mean "
                         "(a near-tied join); the posterior is not twice differentiable there")

    def laplace(self, n_samples=100, step=1e-5):
        """Draw embeddings from N(mean, -H^-1) and decode them (laplace.py:27-63): returns
        dict(locs (n,S,D), peel, blens, ln_p, ln_prior), all samples evaluated in one batch."""
        mean = self.get_locs().detach().reshape(-1)
        H = self.hessian(mean, step)
        if not bool(torch.linalg.eigvalsh(H).max() < 0.0):
            raise ValueError("Laplace: the Hessian at the mean is not negative definite (not at a mode)")
        cov = -torch.linalg.inv(H)
        dist = torch.distributions.MultivariateNormal(mean, covariance_matrix=0.5 * (cov + cov.T))
        locs = dist.sample(torch.Size((n_samples,))).reshape(n_samples, self.S, self.D)
        ln_p, peel, blens = self.log_likelihood_batch_np(locs.numpy())
        ln_prior = priors.compute_prior_gamma_dir(torch.from_numpy(blens)).numpy()
        return dict(locs=locs.numpy(), peel=peel, blens=blens, ln_p=ln_p, ln_prior=ln_prior)

"""The whole hot path as one device-resident step:

    locations (B,S,D) -> pairwise hyperbolic distance -> (soft|hard) NJ -> P(t) + pruning lnL
                      -> d lnL/d blens -> soft-NJ backward -> distance backward -> d lnL/d locations

This is what one VI sample (vi.py:467-491 ``connect`` + base_model.py:219-247 ``compute_LL`` +
``loss.backward()``) or one MCMC proposal (chain.py:171-221, :145-169) does in the reference, for the
whole batch of samples / chains at once.  No host synchronisation inside a step.
"""
import numpy as np
import torch

from . import engine


class HotPath:
    def __init__(self, tipmask, weights, model, rates=None, mix=False, tau=1e-8, curvature=-1.0,
                 lift="up", matsumoto=False, device=None):
        self.prune = engine.PruneEngine(tipmask, weights, device=device)
        self.device = self.prune.device
        self.model = model
        self.rates = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        self.mix = bool(mix)
        self.tau = tau
        self.curvature = curvature
        self.lift = lift
        self.matsumoto = matsumoto

    @classmethod
    def from_engine(cls, prune_engine, model, rates=None, mix=False, tau=1e-8, curvature=-1.0, lift="up", matsumoto=False):
        """Share an existing PruneEngine (tips already resident on the device)."""
        self = cls.__new__(cls)
        self.prune = prune_engine
        self.device = prune_engine.device
        self.model = model
        self.rates = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        self.mix, self.tau, self.curvature, self.lift, self.matsumoto = bool(mix), tau, curvature, lift, matsumoto
        return self

    def value_and_grad(self, x, to_locations=True):
        """x: (B,S,D) device tensor.  Returns dict(lnL (B,), grad_x (B,S,D), grad_curvature (B,),
        peel (B,S-1,3), blens (B,2S-2)) -- soft-NJ (VI) path.  ``to_locations=False`` stops at
        d lnL/d blens (the caller chains NJ / distance backward itself)."""
        pdm = engine.pdm_forward(x, self.curvature, "torch", self.lift, self.matsumoto)
        peel, blens, flags = engine.nj_soft(pdm, self.tau, return_flags=True)
        ll, gbl = self.prune.loglik(peel, blens, self.model, rates=self.rates, mix=self.mix, grad=True)
        if not to_locations:
            return dict(lnL=ll, peel=peel, blens=blens, flags=flags, grad_blens=gbl)
        gpdm = engine.nj_soft_backward(peel, flags, gbl)
        gx, gc = engine.pdm_backward(x, gpdm, self.curvature, self.lift, self.matsumoto)
        return dict(lnL=ll, grad_x=gx, grad_curvature=gc, peel=peel, blens=blens, flags=flags, grad_blens=gbl)

    def blens_jacobian(self, x, peel=None, flags=None, logdet=True):
        """d blens / d locations for every sample, (B, 2S-2, S*D), and log sqrt|det(J J^T)| -- the
        change-of-variables term the reference obtains with ``torch.autograd.functional.jacobian``
        and 2S-2 backward passes per sample (vi.py:376-387; ~3/4 of a reference VI sample).  Here the
        2S-2 one-hot cotangents of all B samples go through the NJ and distance backward kernels as
        ONE batch of B*(2S-2) replays.  (The last NJ join splits one distance into two equal halves,
        so two rows of J coincide and the determinant is numerically ~0 in the reference as well.)"""
        B, S, D = (int(v) for v in x.shape)
        E = 2 * S - 2
        if peel is None:
            pdm = engine.pdm_forward(x, self.curvature, "torch", self.lift, self.matsumoto)
            peel, _, flags = engine.nj_soft(pdm, self.tau, return_flags=True)
        cot = torch.eye(E, dtype=torch.float64, device=x.device).repeat(B, 1)
        gpdm = engine.nj_soft_backward(peel.repeat_interleave(E, 0), flags.repeat_interleave(E, 0), cot)
        gx, _ = engine.pdm_backward(x.repeat_interleave(E, 0), gpdm, self.curvature, self.lift, self.matsumoto)
        J = gx.reshape(B, E, S * D)
        if not logdet:
            return J
        sign, logabs = torch.linalg.slogdet(J @ J.transpose(1, 2))
        return J, 0.5 * logabs

    def value(self, x):
        """Gradient-free MCMC path (Chyp_np.get_pdm -> Cpeeler.nj_np -> Cphylo.compute_LL_np)."""
        pdm = engine.pdm_forward(x, self.curvature, "numpy", "up", self.matsumoto)
        peel, blens = engine.nj_hard(pdm)
        ll, _ = self.prune.loglik(peel, blens, self.model, rates=self.rates, mix=self.mix, grad=False)
        return dict(lnL=ll, peel=peel, blens=blens)

    # end-to-end call with host buffers (what a reference-side caller holding NumPy/CPU tensors does)
    def value_and_grad_host(self, x_pinned, out_ll_pinned, out_gx_pinned):
        x = x_pinned.to(self.device, non_blocking=True)
        r = self.value_and_grad(x)
        out_ll_pinned.copy_(r["lnL"], non_blocking=True)
        out_gx_pinned.copy_(r["grad_x"], non_blocking=True)
        return r

"""The part of the reference's ``BaseModel`` that sits on the hot path
(dodonaphy/base_model.py): construction from tip partials + weights (:35-107), the learnable
curvature leaf (:109-124), the tree-decoding ``connect`` step its subclasses implement
(vi.py:467-491, chain.py:171-221) and the likelihood dispatch ``compute_LL`` (:219-247).

Two ways in:

* one tree at a time, with the reference's signatures (``compute_LL(peel, blen)``,
  ``connect(locs)``) -- what the unmodified VI / MCMC loops call per sample;
* a whole batch at once (``log_likelihood_batch``, ``log_likelihood_batch_np``) -- the batching
  glue for ``DodonaphyVI.elbo_siwae`` (importance x boosts samples, vi.py:328-333) and
  ``DodonaphyMCMC.learn`` (one proposal per chain, mcmc.py:136-146): every sample / chain of a step
  goes through the kernels in a single pass.

Priors, logging, checkpointing and the alternative research losses of the reference class are host
code outside the path and are not reproduced here.
"""
import numpy as np
import torch

from . import Chyp_np, Chyp_torch, Cpeeler, peeler
from .phylo import LikelihoodInstance, TreeLikelihood, calculate_treelikelihood
from .phylomodel import PhyloModel, rate_matrix_jacobian_np
from .pipeline import HotPath


def _ln_prior_of(out):
    """ln prior from the NJ epilogue when the hot path carries a prior, zeros otherwise."""
    return out["ln_prior"] if out.get("ln_prior") is not None else torch.zeros_like(out["blens"][:, 0])


class _BatchedLogLik(torch.autograd.Function):
    """(lnL, peel, blens) of B embeddings in one fused pass.  lnL and blens both carry gradient to the
    locations and to the curvature leaf: the backward adds the caller's cotangent on blens (e.g. from a
    branch-length prior) to g_lnL * d lnL/d blens and replays soft NJ and the distance kernel once."""

    @staticmethod
    def forward(ctx, locs, curvature_leaf, hot):
        # the leaf is log(-kappa) (base_model.py:109-124): kappa = -exp(leaf), d kappa / d leaf = kappa
        hot.curvature = -float(torch.exp(curvature_leaf.detach().reshape(-1)[0]))
        x = locs.detach().to(hot.device, dtype=torch.float64).contiguous()
        # flagged soft-NJ samples: redone exhaustively / raised, like the one-tree route (peeler._SoftNJ)
        out = hot.resolve(x, hot.value_and_grad(x, to_locations=False), to_locations=False)
        ctx.hot, ctx.kappa, ctx.cshape = hot, hot.curvature, curvature_leaf.shape
        ctx.devs = (locs.device, curvature_leaf.device)
        ctx.x, ctx.dec = x, {k: out[k] for k in ("peel", "flags", "slots", "tape", "grad_blens", "grad_prior") if k in out}
        ctx.mark_non_differentiable(out["peel"])
        hot.last_decode = (x, out["peel"], out["flags"])  # for the Jacobian term of the same step (vi.py:376-387)
        return out["lnL"].to(locs.device), out["peel"], out["blens"].to(locs.device), _ln_prior_of(out).to(locs.device)

    @staticmethod
    def backward(ctx, g_ll, _g_peel, g_blens, g_prior):
        x, hot = ctx.x, ctx.hot
        gb = g_ll.to(x.device).reshape(-1, 1) * ctx.dec["grad_blens"]
        if g_blens is not None:
            gb = gb + g_blens.to(x.device)
        if g_prior is not None and "grad_prior" in ctx.dec:  # prior from the NJ kernel's epilogue
            gb = gb + g_prior.to(x.device).reshape(-1, 1) * ctx.dec["grad_prior"]
        hot.curvature = ctx.kappa
        gx, gc = hot.decode_backward(x, ctx.dec, gb.contiguous())
        return gx.to(ctx.devs[0]), (gc.sum() * ctx.kappa).reshape(ctx.cshape).to(ctx.devs[1]), None


class _BatchedLogLikModel(torch.autograd.Function):
    """As ``_BatchedLogLik`` with a learnable GTR model: gradient also reaches the unconstrained
    rate / frequency leaves (phylomodel.py:19-45) -- the batched form of the reference's non-bito branch
    (base_model.py:239-247: explicit ``get_transition_mats`` + ``calculate_treelikelihood`` under autograd).
    P(t) is formed on the device from the current eigen-system; ONE pruning launch (``dodo_prune_model``) returns
    lnL, d lnL/d blens, d lnL/d pi and one 4 x 4 matrix per sample, W_b with d lnL_b/d Q = U^-T W_b U^T.  What
    is left for the host is the 4 x 4 map (rates, freqs) -> normalised Q: its 20 x 8 Jacobian is formed in closed
    form (``phylomodel.rate_matrix_jacobian_np``) while the device is busy, so the backward pass is one 20-number
    device->host read and a matrix-vector product."""

    @staticmethod
    def forward(ctx, locs, curvature_leaf, raw_rates, raw_freqs, hot):
        hot.curvature = -float(torch.exp(curvature_leaf.detach().reshape(-1)[0]))
        x = locs.detach().to(hot.device, dtype=torch.float64).contiguous()
        out = hot.value_and_grad_model(x)  # launched; the host carries on
        # while the device works: the Jacobian of the 4 x 4 map (rates, freqs) -> (normalised Q, pi), 20 x (5 + 3)
        # numbers on the host in closed form -- all the backward pass needs besides W and d lnL / d pi
        ctx.jac = rate_matrix_jacobian_np(raw_rates.detach().cpu().numpy(), raw_freqs.detach().cpu().numpy())[2]
        ctx.n_rates, ctx.leaf_like = int(raw_rates.numel()), (raw_rates, raw_freqs)
        out = hot.resolve(x, out, model=True)
        ctx.hot, ctx.kappa, ctx.cshape = hot, hot.curvature, curvature_leaf.shape
        ctx.devs = (locs.device, curvature_leaf.device)
        ctx.x, ctx.dec = x, {k: out[k] for k in ("peel", "flags", "slots", "tape", "grad_blens", "grad_prior") if k in out}
        ctx.wg = torch.cat([out["W"].reshape(-1, 16), out["grad_pi"]], dim=1)  # (B, 20): one contraction in backward
        ctx.model = hot.model
        ctx.mark_non_differentiable(out["peel"])
        hot.last_decode = (x, out["peel"], out["flags"])  # for the Jacobian term of the same step (vi.py:376-387)
        return out["lnL"].to(locs.device), out["peel"], out["blens"].to(locs.device), _ln_prior_of(out).to(locs.device)

    @staticmethod
    def backward(ctx, g_ll, _g_peel, g_blens, g_prior):
        x, hot = ctx.x, ctx.hot
        w = g_ll.to(x.device).reshape(-1)
        gb = w.reshape(-1, 1) * ctx.dec["grad_blens"]
        if g_blens is not None:
            gb = gb + g_blens.to(x.device)
        if g_prior is not None and ctx.dec.get("grad_prior") is not None:
            gb = gb + g_prior.to(x.device).reshape(-1, 1) * ctx.dec["grad_prior"]
        # model leaves: sum_b w_b (U^-T W_b U^T) : dQ  +  sum_b w_b grad_pi_b : d pi  -- 20 numbers to the host
        packed = (w.reshape(-1, 1) * ctx.wg).sum(0)
        hot.curvature = ctx.kappa
        gx, gc = hot.decode_backward(x, ctx.dec, gb.contiguous())  # launched before the host waits for `packed`
        packed = packed.cpu().numpy()
        v = np.concatenate([(ctx.model.Uinv.T @ packed[:16].reshape(4, 4) @ ctx.model.U.T).reshape(-1), packed[16:]])
        g = torch.as_tensor(v @ ctx.jac)
        g_rates, g_freqs = (t.to(device=like.device, dtype=like.dtype)
                            for t, like in zip((g[:ctx.n_rates], g[ctx.n_rates:]), ctx.leaf_like))
        return (gx.to(ctx.devs[0]), (gc.sum() * ctx.kappa).reshape(ctx.cshape).to(ctx.devs[1]),
                g_rates, g_freqs, None)


class BaseModel:
    def __init__(self, inference_name, partials, weights, dim, soft_temp, curvature=-1.0, embedder="up",
                 connector="nj", normalise_leaf=False, loss_fn="likelihood", require_grad=True, matsumoto=False,
                 tip_labels=None, model_name="JC69", freqs=None, gamma_shape=None, n_rate_categories=1):
        if embedder not in ("wrap", "up"):
            raise AssertionError("embedder must be 'wrap' or 'up'")
        if connector not in ("geodesics", "nj", "nj-r", "fix"):
            raise AssertionError("unknown connector")
        self.inference_name = inference_name
        n_tips = len(partials) if tip_labels is None else len(tip_labels)
        self.partials = list(partials[:n_tips])
        self.weights = weights
        self.S = n_tips
        self.L = int(self.partials[0].shape[-1])
        self.D = dim
        self.bcount = 2 * self.S - 2
        self.soft_temp = soft_temp
        self.require_grad = require_grad
        self.embedder, self.connector = embedder, connector
        self.normalise_leaf, self.loss_fn, self.matsumoto = normalise_leaf, loss_fn, matsumoto
        self.tip_labels = tip_labels or [f"T{i + 1}" for i in range(self.S)]
        self.name_id = {name: i for i, name in enumerate(self.tip_labels)}
        self.curvature = curvature
        self.phylomodel = PhyloModel(model_name, freqs=None if freqs is None else torch.as_tensor(freqs))
        from .engine import SubstModel

        self.rates = None if n_rate_categories <= 1 else SubstModel.discrete_gamma(gamma_shape or 0.5, n_rate_categories)
        self._inst = LikelihoodInstance(self.partials, weights, None, rates=self.rates, mix=self.rates is not None)
        self._hot = None
        self.use_bito = False

    # -- curvature: kappa = -exp(_curvature), the leaf is learnable on the VI path ------------------
    @property
    def curvature(self):
        return -torch.exp(self._curvature) if self.require_grad else -np.exp(self._curvature)

    @curvature.setter
    def curvature(self, value):
        if self.require_grad:
            v = torch.as_tensor(value, dtype=torch.float64)
            leaf = torch.log(-v).clone().detach()
            self._curvature = leaf.requires_grad_(True) if not hasattr(self, "_curvature") else leaf
        else:
            self._curvature = np.log(-np.asarray(value, dtype=np.float64))

    # -- one tree -------------------------------------------------------------------------------
    def compute_LL(self, peel, blen):
        """lnL of one tree; ``blen`` a torch tensor (gradient flows) or an ndarray (MCMC).

        With learnable model parameters (GTR) the likelihood is taken through explicit transition
        matrices, exactly the reference's non-bito branch (base_model.py:239-247), so the rates and
        frequencies receive gradient; otherwise value and d lnL/d blens come from one fused call."""
        pm = self.phylomodel
        if isinstance(blen, torch.Tensor) and not (pm.fix_sub_rates and pm.fix_freqs) and self.rates is None:
            return calculate_treelikelihood(self.partials, self.weights, peel, pm.get_transition_mats(blen), pm.freqs)
        self._inst.model = pm.device_model()
        if isinstance(blen, torch.Tensor):
            return TreeLikelihood.apply(blen, peel, self._inst)
        out = TreeLikelihood.apply(torch.as_tensor(np.asarray(blen, dtype=np.float64)), peel, self._inst)
        return np.float64(out.item())

    def connect(self, locs):
        """Decode one embedding into (peel, blens, pdm): neighbour joining on the pairwise distances
        or the geodesic connector (vi.py:467-491; chain.py:171-221, where the third item is the
        internal-node locations for 'geodesics')."""
        if self.connector not in ("nj", "geodesics"):
            raise NotImplementedError(f"connector {self.connector!r} is outside the hot path")
        if self.require_grad:
            pdm = Chyp_torch.get_pdm(locs, curvature=self.curvature, matsumoto=self.matsumoto, projection=self.embedder)
            if self.connector == "geodesics":
                peel, _, blens = peeler.make_soft_peel_tips(locs, connector="geodesics", curvature=self.curvature.detach())
            else:
                peel, blens = peeler.nj_torch(pdm, tau=self.soft_temp)
            return peel, blens, pdm
        if self.connector == "geodesics":
            from . import Cphylo

            peel, int_x = peeler.make_hard_peel_geodesic(locs)
            blens = Cphylo.compute_branch_lengths_np(self.S, peel, locs, int_x, curvature=float(self.curvature),
                                                     matsumoto=self.matsumoto)
            return peel, blens, int_x
        pdm = Chyp_np.get_pdm(locs, curvature=float(self.curvature), matsumoto=self.matsumoto)
        peel, blens = Cpeeler.nj_np(pdm)
        return peel, blens, pdm

    # -- a batch of trees -------------------------------------------------------------------------
    def _hotpath(self, defer_model=False):
        """``defer_model``: a learnable model is eigen-decomposed on the host every step (0.2 ms); the caller then sets
        ``hot.model_fn`` and the decomposition runs after the distance / NJ kernels are launched, in their shadow."""
        model = self._hot.model if (defer_model and self._hot is not None) else self.phylomodel.device_model()
        if self._hot is None:
            self._hot = HotPath.from_engine(self._inst.engine, model, rates=self.rates, mix=self.rates is not None,
                                            tau=self.soft_temp, lift=self.embedder, matsumoto=self.matsumoto,
                                            connector="geodesics" if self.connector == "geodesics" else "nj")
        self._hot.model = model
        self._hot.tau = self.soft_temp
        return self._hot

    def log_likelihood_batch(self, locs, prior=None):
        """VI: (B,S,D) sampled locations -> (lnL (B,), peel (B,S-1,3), blens (B,2S-2)); lnL carries
        gradient to ``locs``, to the curvature leaf and -- for a learnable GTR model -- to the
        unconstrained rate and frequency leaves.  With ``prior`` (an ``engine.GammaDirPrior``) a fourth
        item, ln prior (B,), comes from the epilogue of the same NJ kernel launch (base_model.py:475-527
        evaluated where the branch lengths are produced) and is differentiable like lnL."""
        pm = self.phylomodel
        learn = not (pm.fix_sub_rates and pm.fix_freqs)
        hot = self._hotpath(defer_model=learn)
        hot.prior = prior
        if learn:
            hot.model_fn = pm.device_model
            out = _BatchedLogLikModel.apply(locs, self._curvature, pm._sub_rates, pm._freqs, hot)
        else:
            out = _BatchedLogLik.apply(locs, self._curvature, hot)
        return out if prior is not None else out[:3]

    def log_likelihood_batch_np(self, locs, prior=None):
        """MCMC: (B,S,D) proposals (one per chain) -> (lnL, peel, blens[, ln prior]) as NumPy arrays, no
        gradient; ``prior`` as in ``log_likelihood_batch`` (hard-NJ kernel epilogue)."""
        hot = self._hotpath()
        hot.prior = prior
        hot.curvature = float(self.curvature)
        x = torch.as_tensor(np.ascontiguousarray(locs, dtype=np.float64))
        out = hot.graphed(x, grad=False).run(x)  # every sweep has the same shape: replay one CUDA graph
        res = (out["lnL"].cpu().numpy(), out["peel"].cpu().numpy(), out["blens"].cpu().numpy())
        return res + (out["ln_prior"].cpu().numpy(),) if prior is not None else res

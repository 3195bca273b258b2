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
                 lift="up", matsumoto=False, device=None, connector="nj", prior=None):
        self.connector = connector
        self.prior = prior  # engine.GammaDirPrior: evaluated in the NJ kernel's epilogue
        self.prune = engine.PruneEngine(tipmask, weights, device=device)
        self.device = self.prune.device
        self.model = model
        self.rates = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        self.mix = bool(mix)
        self.tau = tau
        self.curvature = curvature
        self.lift = lift
        self.matsumoto = matsumoto
        self.ws = engine.Workspaces()  # K1 / K2 scratch owned by this object (never shared, never dropped)

    @classmethod
    def from_engine(cls, prune_engine, model, rates=None, mix=False, tau=1e-8, curvature=-1.0, lift="up", matsumoto=False,
                    connector="nj"):
        """Share an existing PruneEngine (tips already resident on the device)."""
        self = cls.__new__(cls)
        self.prior = None
        self.connector = connector
        self.prune = prune_engine
        self.device = prune_engine.device
        self.model = model
        self.rates = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        self.mix, self.tau, self.curvature, self.lift, self.matsumoto = bool(mix), tau, curvature, lift, matsumoto
        self.ws = engine.Workspaces()
        return self

    # -- tree decoding: the differentiable connector and its reverse ------------------------------
    def decode(self, x, exhaustive=False):
        """x (B,S,D) -> dict(peel, blens, flags + what ``decode_backward`` needs).  'nj': distance kernel +
        soft NJ (vi.py:476-484); 'geodesics': the LCA connector on x taken as ball coordinates
        (vi.py:479-482; curvature -1 only, as in the reference).  ``flags`` (B,S-1): bit 2 = the soft
        argmin's closed-form treatment of inactive entries is not provably exact for that join (redo with
        ``exhaustive=True``), bit 3 = it selected an inactive node.  Nothing is checked here: see
        ``value_and_grad`` / ``resolve``."""
        if self.connector == "geodesics":
            if abs(float(self.curvature) + 1.0) > 1e-8:
                raise NotImplementedError(f"Curvature must be -1, got {self.curvature}")
            out = engine.geo_soft(x, 1e-8, ws=self.ws)
            out["flags"] = out["flags"] & 8  # only an invalid selection matters downstream
            if self.prior is not None:
                out["ln_prior"], out["grad_prior"] = engine.prior_gamma_dir(out["blens"], self.prior)
            return out
        if self.connector != "nj":
            raise NotImplementedError(f"connector {self.connector!r} has no differentiable device path")
        pdm = engine.pdm_forward(x, self.curvature, "torch", self.lift, self.matsumoto, ws=self.ws)
        if self.prior is not None:
            peel, blens, flags, (lp, gp) = engine.nj_soft(pdm, self.tau, exhaustive=exhaustive, return_flags=True, ws=self.ws,
                                                          prior=self.prior)
            return dict(peel=peel, blens=blens, flags=flags, ln_prior=lp, grad_prior=gp)
        peel, blens, flags = engine.nj_soft(pdm, self.tau, exhaustive=exhaustive, return_flags=True, ws=self.ws)
        return dict(peel=peel, blens=blens, flags=flags)

    def decode_backward(self, x, dec, grad_blens):
        """VJP of ``decode``: d/d x (B,S,D) and d/d curvature (B,) of sum(grad_blens * blens)."""
        if self.connector == "geodesics":
            gx = engine.geo_soft_backward(dec["peel"], dec["slots"], dec["tape"], grad_blens)
            return gx, torch.zeros(gx.shape[0], dtype=torch.float64, device=gx.device)
        gpdm = engine.nj_soft_backward(dec["peel"], dec["flags"], grad_blens, ws=self.ws)
        return engine.pdm_backward(x, gpdm, self.curvature, self.lift, self.matsumoto, ws=self.ws)

    def value_and_grad(self, x, to_locations=True, exhaustive=False):
        """x: (B,S,D) device tensor.  Returns dict(lnL (B,), grad_x (B,S,D), grad_curvature (B,),
        peel (B,S-1,3), blens (B,2S-2), flags, bad) -- the differentiable (VI) path.
        ``to_locations=False`` stops at d lnL/d blens (the caller chains ``decode_backward`` itself).

        Samples whose soft-NJ flags carry bit 2 or bit 3 have a topology the reference would not have
        produced: their lnL and gradients come back as NaN (poisoned on the device, so the step stays
        free of host synchronisation and can be captured in a graph) and ``bad`` = device int32
        [#guard tripped, #invalid selection].  ``resolve`` applies the reference-faithful policy."""
        dec = self.decode(x, exhaustive=exhaustive)
        ll, gbl = self.prune.loglik(dec["peel"], dec["blens"], self.model, rates=self.rates, mix=self.mix, grad=True)
        dec["bad"] = engine.nj_guard(dec["flags"], ll, gbl)
        dec.update(lnL=ll, grad_blens=gbl)
        if to_locations:
            dec["grad_x"], dec["grad_curvature"] = self.decode_backward(x, dec, gbl)
        return dec

    def value_and_grad_model(self, x, exhaustive=False):
        """As ``value_and_grad(to_locations=False)`` with the gradient w.r.t. the (GTR) model next to the one
        w.r.t. the branch lengths: adds W (B,4,4) and grad_pi (B,4) (``PruneEngine.loglik_model_grad``).
        Flagged samples are poisoned the same way."""
        dec = self.decode(x, exhaustive=exhaustive)
        fn = getattr(self, "model_fn", None)
        if fn is not None:  # the current eigen-system, formed on the host while the device decodes the trees
            self.model, self.model_fn = fn(), None
        out = self.prune.loglik_model_grad(dec["peel"], dec["blens"], self.model, rates=self.rates, mix=self.mix)
        dec["bad"] = engine.nj_guard(dec["flags"], out["lnL"], out["grad_blens"])
        dec.update(out)
        return dec

    def resolve(self, x, out, to_locations=True, model=False):
        """Host side of the flag policy (one device->host read of two counters): samples whose exactness
        guard tripped are re-evaluated with the exhaustive N^2 soft argmin -- the reference's result in
        every regime -- and patched into ``out``; an invalid selection raises, as the one-tree route
        ``peeler.nj_torch`` does.  Call it where the results are about to leave the device anyway."""
        n_guard, n_invalid = (int(v) for v in out["bad"].tolist())
        if n_invalid:
            raise RuntimeError("soft NJ: the soft argmin selected an inactive node for "
                               f"{n_invalid} sample(s) (temperature too high for these distances)")
        if n_guard:
            idx = ((out["flags"] & 4) != 0).any(dim=1).nonzero().reshape(-1)
            sub = (self.value_and_grad_model(x[idx].contiguous(), exhaustive=True) if model else
                   self.value_and_grad(x[idx].contiguous(), to_locations=to_locations, exhaustive=True))
            if int(sub["bad"][1]):
                raise RuntimeError("soft NJ: the soft argmin selected an inactive node")
            for k, v in sub.items():
                if k != "bad" and isinstance(v, torch.Tensor) and k in out and v.shape[0] == idx.numel():
                    out[k][idx] = v
            out["bad"] = torch.zeros_like(out["bad"])
        return out

    def decode_checked(self, x):
        """``decode`` + the same host policy, for callers that synchronise right after decoding."""
        dec = self.decode(x)
        bad = int((dec["flags"] & 12).max().item()) if dec["flags"].numel() else 0
        if bad & 8:
            raise RuntimeError("soft NJ: the soft argmin selected an inactive node (temperature too high for these distances)")
        if bad & 4:
            idx = ((dec["flags"] & 4) != 0).any(dim=1).nonzero().reshape(-1)
            sub = self.decode(x[idx].contiguous(), exhaustive=True)
            if int((sub["flags"] & 8).max().item()):
                raise RuntimeError("soft NJ: the soft argmin selected an inactive node")
            for k, v in sub.items():
                dec[k][idx] = v
        return dec

    def blens_jacobian(self, x, peel=None, flags=None, logdet=True, replay=False):
        """d blens / d locations for every sample, (B, 2S-2, S*D), and log sqrt|det(J J^T)| -- the
        change-of-variables term the reference obtains with ``torch.autograd.functional.jacobian``
        and 2S-2 backward passes per sample (vi.py:376-387; ~3/4 of a reference VI sample).  For the NJ
        connector all 2S-2 rows of a tree come out of ONE pass over its joins (``engine.blens_jacobian``:
        every NJ distance is a bilinear form of the tip distances, so the rows follow from per-node
        quantities that are elementwise in (taxon, coordinate)).  ``replay=True`` (and the geodesic connector)
        takes the 2S-2 one-hot cotangents through the backward kernels as ONE batch of B*(2S-2) replays --
        the literal batched form of the reference's loop, kept as a cross-check.  (The last NJ join splits
        one distance into two equal halves, so two rows of J coincide and the determinant is numerically ~0
        in the reference as well.)"""
        B, S, D = (int(v) for v in x.shape)
        E = 2 * S - 2
        if self.connector == "geodesics":
            dec = self.decode(x)
            cot = torch.eye(E, dtype=torch.float64, device=x.device).repeat(B, 1)
            rep = {k: dec[k].repeat_interleave(E, 0) for k in ("peel", "slots", "tape")}
            gx, _ = self.decode_backward(None, rep, cot)
            J = gx.reshape(B, E, S * D)
        else:
            if peel is None:
                dec = self.decode_checked(x)
                peel, flags = dec["peel"], dec["flags"]
            if not replay:
                # the workspace is ~ (5S) x S(D+1) doubles per tree: chunk the batch to a memory budget
                per = max(1, int(engine._lib.load().dodo_blens_jacobian_workspace_bytes(1, S, D)) + 8 * E * S * D)
                chunk = max(1, min(B, int((4 << 30) // per)))
                parts = [engine.blens_jacobian(x[lo:lo + chunk], peel[lo:lo + chunk], flags[lo:lo + chunk], self.curvature,
                                               self.lift, self.matsumoto, ws=self.ws) for lo in range(0, B, chunk)]
                J = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
            else:
                # every replay needs an S x S adjoint workspace: chunk the B*E replays to a memory budget
                # (2000 taxa: 32 MB per replay) instead of asking for all of them at once
                per = max(1, int(engine._lib.load().dodo_nj_workspace_bytes(1, S)))
                chunk = max(1, min(B, int((2 << 30) // (per * E))))
                parts = []
                eye = torch.eye(E, dtype=torch.float64, device=x.device)
                for lo in range(0, B, chunk):
                    hi = min(B, lo + chunk)
                    rep = dict(peel=peel[lo:hi].repeat_interleave(E, 0), flags=flags[lo:hi].repeat_interleave(E, 0))
                    g, _ = self.decode_backward(x[lo:hi].repeat_interleave(E, 0), rep, eye.repeat(hi - lo, 1))
                    parts.append(g)
                gx = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
                J = gx.reshape(B, E, S * D)
        if not logdet:
            return J
        return J, engine.gram_logdet(J)

    def value(self, x):
        """Gradient-free MCMC path (Chyp_np.get_pdm -> Cpeeler.nj_np -> Cphylo.compute_LL_np)."""
        if self.connector != "nj":
            raise NotImplementedError("the batched gradient-free path decodes with hard NJ only")
        pdm = engine.pdm_forward(x, self.curvature, "numpy", "up", self.matsumoto, ws=self.ws)
        out = {}
        if self.prior is not None:
            peel, blens, (out["ln_prior"], _) = engine.nj_hard(pdm, ws=self.ws, prior=self.prior)
        else:
            peel, blens = engine.nj_hard(pdm, ws=self.ws)
        ll, _ = self.prune.loglik(peel, blens, self.model, rates=self.rates, mix=self.mix, grad=False)
        out.update(lnL=ll, peel=peel, blens=blens)
        return out

    # end-to-end call with host buffers (what a reference-side caller holding NumPy/CPU tensors does)
    def _fingerprint(self, shape):
        pr = None if self.prior is None else (self.prior.aT, self.prior.bT, self.prior.a, self.prior.c, self.prior.variant)
        return (tuple(shape), bytes(self.model.c_struct()), None if self.rates is None else self.rates.tobytes(),
                self.mix, None if self.tau is None else float(self.tau), float(self.curvature), self.lift,
                self.matsumoto, self.connector, pr)

    def graphed(self, x, grad):
        """The step for inputs shaped like ``x`` as a captured CUDA graph (one live capture per
        gradient / gradient-free route, re-captured when shape, model or settings change)."""
        key = self._fingerprint(x.shape)
        cache = self.__dict__.setdefault("_graphs", {})
        hit = cache.get(bool(grad))
        if hit is None or hit[0] != key:
            hit = cache[bool(grad)] = (key, GraphedStep(self, x.to(self.device), grad=grad))
        return hit[1]

    def value_and_grad_host(self, x_pinned, out_ll_pinned, out_gx_pinned, graph=True):
        """Host buffers in, host buffers out: pinned (B,S,D) locations -> lnL (B,) and d lnL/d locations
        (B,S,D) written into the caller's pinned buffers (asynchronously on the current stream).
        The device part of the step is captured once per (shape, model, settings) as a CUDA graph and
        replayed -- with pinned buffers the three copies are nodes of that graph, so a caller that synchronises
        every step pays ONE launch instead of the Python launch sequence; ``graph=False`` launches kernel by kernel."""
        if graph and x_pinned.is_pinned() and out_ll_pinned.is_pinned() and out_gx_pinned.is_pinned():
            # the three copies are nodes of the captured graph as well (one launch per step); a capture belongs to
            # one triple of host buffers, which it keeps alive
            bufs = (x_pinned, out_ll_pinned, out_gx_pinned)
            key = (self._fingerprint(x_pinned.shape),) + tuple(t.data_ptr() for t in bufs)
            hit = self.__dict__.get("_host_graph")
            if hit is None or hit[0] != key:
                hit = self.__dict__["_host_graph"] = (key, GraphedStep(self, x_pinned.to(self.device), grad=True, host=bufs))
            return hit[1].run_host()
        if graph:
            r = self.graphed(x_pinned, grad=True).run(x_pinned)
        else:
            r = self.value_and_grad(x_pinned.to(self.device, non_blocking=True))
        # flagged samples arrive as NaN (see value_and_grad); r["bad"] holds their counts
        out_ll_pinned.copy_(r["lnL"], non_blocking=True)
        out_gx_pinned.copy_(r["grad_x"], non_blocking=True)
        return r


class GraphedStep:
    """One step of the path captured ONCE as a CUDA graph and replayed.

    Small problems (the reference's own DS1 case: 27 taxa x 934 patterns, a handful of chains) are
    launch-bound: ~10 kernels of a few microseconds each behind Python / ctypes call overhead.  The
    whole step -- every launch between the input buffer and the outputs -- is recorded into a graph on
    first use; ``run(x)`` copies the new locations into the captured input buffer and replays.  The
    substitution model, rates, temperature and curvature are baked in at capture time (they travel as
    kernel parameters): capture again after changing them.  Outputs are the captured tensors,
    overwritten by every replay.
    """

    def __init__(self, hot, x_example, grad=True, host=None):
        # the graph addresses its scratch by raw pointer: give it workspaces of its own (a private
        # view of the HotPath sharing the tips and the prune plans, which are never reallocated), so that
        # nothing the owner does later -- a bigger batch, the batched Jacobian -- can free them
        import copy

        self.hot = hot = copy.copy(hot)
        hot.ws = engine.Workspaces()
        hot.ws.freeze()
        hot.__dict__.pop("_graphs", None)
        hot.__dict__.pop("_host_graph", None)
        self.host = host  # (x, lnL, grad_x) pinned host buffers whose copies are part of the graph, or None
        self.x = x_example.detach().to(hot.device, dtype=torch.float64).contiguous().clone()
        fn = hot.value_and_grad if grad else hot.value
        side = torch.cuda.Stream(device=hot.device)
        side.wait_stream(torch.cuda.current_stream(hot.device))
        with torch.cuda.stream(side):  # warm-up off the capture: workspaces, plans, function attributes
            for _ in range(2):
                fn(self.x)
        torch.cuda.current_stream(hot.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if host is not None:
                self.x.copy_(host[0], non_blocking=True)
            self.out = fn(self.x)
            if host is not None:  # flagged samples arrive as NaN (see value_and_grad); out["bad"] holds their counts
                host[1].copy_(self.out["lnL"], non_blocking=True)
                host[2].copy_(self.out["grad_x"], non_blocking=True)

    def run_host(self):
        """Replay with the captured host buffers: host[0] -> device, the step, results -> host[1], host[2]."""
        self.graph.replay()
        return self.out

    def run(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out

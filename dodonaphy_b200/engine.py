"""Batched device engine for the hot path (the B200-first API the reference-shaped shims sit on).

``SubstModel``   host-side substitution-model parameters -> ``dodo_model_t``
                 (PhyloModel, dodonaphy/phylomodel.py:76-128)
``PruneEngine``  tips resident in HBM as packed IUPAC masks, workspace cached per batch shape,
                 one call = lnL (and d lnL / d blens) for a whole batch of trees
                 (phylo.calculate_treelikelihood, dodonaphy/phylo.py:132-139, + autograd;
                 Cphylo.compute_LL_np, dodonaphy/cython/Cphylo.pyx:9-23)

Everything here is plumbing around the C ABI (include/dodonaphy_b200.h): torch is used for device
memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SubstModel:
    """JC69 or GTR substitution model (host side, tiny).  The eigen-system is obtained exactly as
    the reference does (phylomodel.py:96-114): eigh of sqrt(pi) Q sqrt(pi)^-1."""

    def __init__(self, kind, U, lam, Uinv, Q, pi):
        self.kind = kind
        self.U, self.lam, self.Uinv, self.Q, self.pi = (np.ascontiguousarray(a, dtype=np.float64) for a in (U, lam, Uinv, Q, pi))

    @staticmethod
    def jc69(freqs=None):
        Q = np.full((4, 4), 1.0 / 3.0)
        np.fill_diagonal(Q, -1.0)
        pi = np.full(4, 0.25) if freqs is None else np.asarray(freqs, dtype=np.float64)
        return SubstModel(_lib.MODEL_JC69, np.eye(4), np.zeros(4), np.eye(4), Q, pi)

    @staticmethod
    def gtr_q(rates, freqs):
        """compute_Q_martix + norm (phylomodel.py:116-128)."""
        rates = np.asarray(rates, dtype=np.float64)
        freqs = np.asarray(freqs, dtype=np.float64)
        if rates.shape != (6,) or freqs.shape != (4,):
            raise ValueError("GTR needs 6 exchangeabilities and 4 frequencies")
        iu = np.triu_indices(4, 1)
        Q = np.zeros((4, 4))
        Q[iu[0], iu[1]] = rates
        Q[iu[1], iu[0]] = rates
        Q = Q * freqs[None, :]
        Q = Q + np.diag(-np.sum(Q, axis=1))
        return Q / (-np.sum(np.diagonal(Q) * freqs))

    @staticmethod
    def gtr(rates, freqs, root_freqs=None):
        freqs = np.asarray(freqs, dtype=np.float64)
        Q = SubstModel.gtr_q(rates, freqs)
        sp = np.diag(np.sqrt(freqs))
        spi = np.diag(1.0 / np.sqrt(freqs))
        lam, V = np.linalg.eigh(sp @ Q @ spi)
        pi = freqs if root_freqs is None else np.asarray(root_freqs, dtype=np.float64)
        return SubstModel(_lib.MODEL_EIGEN, spi @ V, lam, np.linalg.inv(V) @ sp, Q, pi)

    @staticmethod
    def discrete_gamma(alpha, K):
        """Yang (1994) mean-rate discrete gamma categories (shape = rate = alpha; mean 1).
        New functionality: the reference has no rate heterogeneity."""
        if K == 1:
            return np.ones(1)
        from scipy.special import gammainc
        from scipy.stats import gamma

        cuts = gamma.ppf(np.arange(1, K) / K, a=alpha, scale=1.0 / alpha)
        cdf = np.concatenate([[0.0], gammainc(alpha + 1.0, cuts * alpha), [1.0]])
        return (cdf[1:] - cdf[:-1]) * K

    def c_struct(self):
        m = _lib.Model()
        for i in range(16):
            m.U[i] = self.U.reshape(-1)[i]
            m.Uinv[i] = self.Uinv.reshape(-1)[i]
            m.Q[i] = self.Q.reshape(-1)[i]
        for i in range(4):
            m.lam[i] = self.lam[i]
            m.pi[i] = self.pi[i]
        m.kind = self.kind
        return m


def pack_tip_partials(partials):
    """Tip partials as the reference holds them (list of S arrays/tensors (4, L) with 0/1 entries,
    phylo.py:27-46) -> uint8 mask array (S, L).  Raises ValueError for non-0/1 values."""
    rows = []
    for p in partials:
        a = p.detach().cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p)
        a = a.reshape(4, -1)
        if not np.all((a == 0.0) | (a == 1.0)):
            raise ValueError("tip partials must be 0/1 indicator vectors")
        rows.append((a[0] + 2 * a[1] + 4 * a[2] + 8 * a[3]).astype(np.uint8))
    return np.stack(rows)


def validate_peel(peel, S):
    """Host-side check of a peel array (raises ValueError like the reference would fail)."""
    peel = np.asarray(peel)
    if peel.shape[-2:] != (S - 1, 3):
        raise ValueError(f"peel must have shape (..., {S - 1}, 3), got {peel.shape}")
    flat = peel.reshape(-1, S - 1, 3)
    for pl in flat:
        seen = set(range(S))
        for l, r, p in pl:
            if l not in seen or r not in seen or l == r or not (S <= p < 2 * S - 1) or p in seen:
                raise ValueError("peel is not a valid post-order of a binary tree")
            seen.discard(int(l))
            seen.discard(int(r))
            seen.add(int(p))
        if len(seen) != 1:
            raise ValueError("peel does not describe a single tree")


class PruneEngine:
    """Felsenstein pruning over a batch of trees on one GPU.

    tipmask : (S, L) uint8 IUPAC masks (bit j <-> state j in order A,C,G,T)
    weights : (L,) pattern weights (int64 or float64)
    """

    def __init__(self, tipmask, weights, device=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tm = torch.as_tensor(np.ascontiguousarray(tipmask) if isinstance(tipmask, np.ndarray) else tipmask)
        if tm.dtype != torch.uint8 or tm.dim() != 2:
            raise TypeError("tipmask must be a (S, L) uint8 array")
        self.S, self.L = int(tm.shape[0]), int(tm.shape[1])
        self.E = 2 * self.S - 2
        self.tipmask = tm.to(self.device).contiguous()
        w = torch.as_tensor(weights)
        if w.numel() != self.L:
            raise ValueError("weights must have one entry per site pattern")
        self.weights = w.to(self.device, dtype=torch.float64).contiguous()
        self._plans = {}
        self.grid_hint = 0
        self.threads_hint = 0

    def _plan(self, B, K, flags):
        key = (B, K, flags, self.grid_hint, self.threads_hint)
        hit = self._plans.get(key)
        if hit is None:
            plan = _lib.PrunePlan()
            _lib.check(self.lib.dodo_prune_plan(B, self.S, self.L, K, flags, self.grid_hint, self.threads_hint, C.byref(plan)))
            ws = torch.empty(int(plan.ws_bytes), dtype=torch.uint8, device=self.device)
            hit = (plan, ws)
            self._plans[key] = hit
        return hit

    def loglik(self, peel, blens, model, rates=None, mix=False, grad=False, out=None, mats=None, scale=None):
        """lnL for a batch.

        peel  : int64 device tensor (B, S-1, 3) or (S-1, 3) (one topology shared by the batch)
        blens : float64 device tensor (B, 2S-2) (or (2S-2,) for B = 1); ignored if ``mats`` given
        mats  : optional float64 device tensor (B, 2S-2, K, 4, 4): caller-supplied P matrices
        rates : K category rates (host); None => K = 1
        scale : power-of-two rescaling against underflow (bit-identical to the unscaled result where
                that one is finite); None => on for >= 200 taxa
        out   : optional (lnL (B,), grad_blens (B, 2S-2) or None) device tensors to write into (e.g. two
                views of one packed buffer that a collective then reduces in place)
        Returns (lnL (B,), grad_blens (B, 2S-2) or None), both device tensors.
        """
        with torch.cuda.device(self.device):
            if mats is not None:
                mats = mats.to(self.device, dtype=torch.float64).contiguous()
                B, K = int(mats.shape[0]), int(mats.shape[2])
                blens_t = None
                rates_arr = np.ones(K)
            else:
                blens_t = blens.to(self.device, dtype=torch.float64)
                if blens_t.dim() == 1:
                    blens_t = blens_t.unsqueeze(0)
                blens_t = blens_t.contiguous()
                B = int(blens_t.shape[0])
                rates_arr = np.ones(1) if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
                K = int(rates_arr.shape[0])
                if blens_t.shape[1] != self.E:
                    raise ValueError(f"blens must have {self.E} entries per tree")
            peel_t = peel.to(self.device, dtype=torch.int64)
            if peel_t.dim() == 2:
                peel_t = peel_t.unsqueeze(0)
            peel_t = peel_t.contiguous()
            if tuple(peel_t.shape[1:]) != (self.S - 1, 3) or peel_t.shape[0] not in (1, B):
                raise ValueError("peel must have shape (B or 1, S-1, 3)")
            if scale is None:
                scale = self.S >= 200
            flags = (_lib.PRUNE_GRAD if grad else 0) | (_lib.PRUNE_MIX if (mix and K > 1) else 0) | (_lib.PRUNE_SCALE if scale else 0)
            plan, ws = self._plan(B, K, flags)
            if out is not None:
                ll, g = out
                if ll.shape != (B,) or ll.dtype != torch.float64 or not ll.is_contiguous():
                    raise ValueError("out[0] must be a contiguous float64 tensor of shape (B,)")
                if grad and (g is None or g.shape != (B, self.E) or g.dtype != torch.float64 or not g.is_contiguous()):
                    raise ValueError("out[1] must be a contiguous float64 tensor of shape (B, 2S-2)")
            else:
                ll = torch.empty(B, dtype=torch.float64, device=self.device)
                g = torch.empty((B, self.E), dtype=torch.float64, device=self.device) if grad else None
            cm = model.c_struct()
            rates_c = (C.c_double * K)(*rates_arr.tolist())
            _lib.check(self.lib.dodo_prune(C.byref(plan), C.byref(cm), rates_c, _ptr(self.tipmask), self.L,
                                           _ptr(self.weights), _ptr(peel_t), int(peel_t.shape[0]), _ptr(blens_t),
                                           _ptr(mats), _ptr(ws), _ptr(ll), _ptr(g), None, None, _stream()))
            return ll, g

    def loglik_model_grad(self, peel, blens, model, rates=None, mix=False, scale=None, via_mats=False):
        """lnL and its gradient w.r.t. the branch lengths AND the (GTR) model, P(t) formed on the device:
        returns dict(lnL (B,), grad_blens (B, 2S-2), W (B, 4, 4), grad_pi (B, 4)) with
        d lnL_b / d Q = U^-T W_b U^T for the model's normalised rate matrix and grad_pi the gradient w.r.t. the
        root frequencies.  The learnable-GTR step of VI / MAP in one pruning launch (``dodo_prune_model``): every
        thread accumulates its pattern's share of W over the edges and categories of its walk.
        ``via_mats``: the round-1 route (d lnL / d P per edge and category from ``DODO_PRUNE_GMATS``, contracted by
        ``dodo_model_chain``), kept as the cross-check."""
        with torch.cuda.device(self.device):
            blens_t = blens.to(self.device, dtype=torch.float64)
            if blens_t.dim() == 1:
                blens_t = blens_t.unsqueeze(0)
            blens_t = blens_t.contiguous()
            B = int(blens_t.shape[0])
            rates_arr = np.ones(1) if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
            K = int(rates_arr.shape[0])
            peel_t = peel.to(self.device, dtype=torch.int64)
            if peel_t.dim() == 2:
                peel_t = peel_t.unsqueeze(0)
            peel_t = peel_t.contiguous()
            if tuple(peel_t.shape[1:]) != (self.S - 1, 3) or peel_t.shape[0] not in (1, B):
                raise ValueError("peel must have shape (B or 1, S-1, 3)")
            ll = torch.empty(B, dtype=torch.float64, device=self.device)
            gpi = torch.empty((B, 4), dtype=torch.float64, device=self.device)
            gbl = torch.empty((B, self.E), dtype=torch.float64, device=self.device)
            W = torch.empty((B, 4, 4), dtype=torch.float64, device=self.device)
            cm = model.c_struct()
            rates_c = (C.c_double * K)(*rates_arr.tolist())
            if via_mats:
                flags = _lib.PRUNE_GMATS | (_lib.PRUNE_MIX if (mix and K > 1) else 0)
                plan, ws = self._plan(B, K, flags)
                gP = self._buffer(("gP", B, K), (B, self.E, K, 4, 4))
                _lib.check(self.lib.dodo_prune(C.byref(plan), C.byref(cm), rates_c, _ptr(self.tipmask), self.L,
                                               _ptr(self.weights), _ptr(peel_t), int(peel_t.shape[0]), _ptr(blens_t),
                                               None, _ptr(ws), _ptr(ll), None, _ptr(gP), _ptr(gpi), _stream()))
                _lib.check(self.lib.dodo_model_chain(C.byref(cm), rates_c, K, _ptr(blens_t), _ptr(gP), B, self.S, _ptr(gbl),
                                                     _ptr(W), _stream()))
                return dict(lnL=ll, grad_blens=gbl, W=W, grad_pi=gpi)
            if scale is None:
                scale = self.S >= 200
            flags = _lib.PRUNE_GMODEL | (_lib.PRUNE_MIX if (mix and K > 1) else 0) | (_lib.PRUNE_SCALE if scale else 0)
            plan, ws = self._plan(B, K, flags)
            _lib.check(self.lib.dodo_prune_model(C.byref(plan), C.byref(cm), rates_c, _ptr(self.tipmask), self.L,
                                                 _ptr(self.weights), _ptr(peel_t), int(peel_t.shape[0]), _ptr(blens_t),
                                                 _ptr(ws), _ptr(ll), _ptr(gbl), _ptr(W), _ptr(gpi), _stream()))
            return dict(lnL=ll, grad_blens=gbl, W=W, grad_pi=gpi)

    def _buffer(self, key, shape):
        """Scratch tensor owned by the engine (kept for the engine's lifetime: graphs may address it)."""
        bufs = self.__dict__.setdefault("_bufs", {})
        if key not in bufs:
            bufs[key] = torch.empty(shape, dtype=torch.float64, device=self.device)
        return bufs[key]

    def loglik_mats(self, peel, mats, pi, grad=False, mix=False):
        """lnL from caller-supplied transition matrices (the ``mats`` / ``freqs`` arguments of
        phylo.calculate_treelikelihood).  mats: (B, 2S-2, K, 4, 4) device tensor; pi: 4 root
        frequencies (host).  Returns dict(lnL (B,), grad_mats like mats, grad_pi (B,4)) -- the
        gradients only when ``grad``."""
        with torch.cuda.device(self.device):
            mats = mats.to(self.device, dtype=torch.float64).contiguous()
            if mats.dim() != 5 or mats.shape[1] != self.E or tuple(mats.shape[3:]) != (4, 4):
                raise ValueError(f"mats must have shape (B, {self.E}, K, 4, 4)")
            B, K = int(mats.shape[0]), int(mats.shape[2])
            peel_t = peel.to(self.device, dtype=torch.int64)
            if peel_t.dim() == 2:
                peel_t = peel_t.unsqueeze(0)
            peel_t = peel_t.contiguous()
            if tuple(peel_t.shape[1:]) != (self.S - 1, 3) or peel_t.shape[0] not in (1, B):
                raise ValueError("peel must have shape (B or 1, S-1, 3)")
            flags = (_lib.PRUNE_GMATS if grad else 0) | (_lib.PRUNE_MIX if (mix and K > 1) else 0)
            plan, ws = self._plan(B, K, flags)
            model = SubstModel(_lib.MODEL_EIGEN, np.eye(4), np.zeros(4), np.eye(4), np.zeros((4, 4)), np.asarray(pi, dtype=np.float64))
            ll = torch.empty(B, dtype=torch.float64, device=self.device)
            gP = torch.empty_like(mats) if grad else None
            gpi = torch.empty((B, 4), dtype=torch.float64, device=self.device) if grad else None
            cm = model.c_struct()
            rates_c = (C.c_double * K)(*([1.0] * K))
            _lib.check(self.lib.dodo_prune(C.byref(plan), C.byref(cm), rates_c, _ptr(self.tipmask), self.L,
                                           _ptr(self.weights), _ptr(peel_t), int(peel_t.shape[0]), None,
                                           _ptr(mats), _ptr(ws), _ptr(ll), None, _ptr(gP), _ptr(gpi), _stream()))
            return dict(lnL=ll, grad_mats=gP, grad_pi=gpi)


# --------------------------------------------------------------------------------------------
# K1 / K2 batched device functions
# --------------------------------------------------------------------------------------------
class Workspaces:
    """Device scratch buffers of the K1 / K2 kernels, one per tag, owned by whoever calls the kernels
    (a ``HotPath``, a ``GraphedStep``, or the module-level default used by the one-tree shims).

    Kernels receive raw pointers into these buffers, and a captured CUDA graph keeps replaying them:
    a buffer must therefore never be freed while something may still address it.  Growing a buffer
    keeps the outgrown one alive once the owner has been ``freeze``-d (done by ``GraphedStep`` at
    capture), and nothing is ever dropped behind the owner's back.  One owner = one stream: the
    buffers carry no ordering of their own."""

    def __init__(self):
        self._bufs = {}
        self._retired = []
        self.frozen = False

    def get(self, tag, nbytes, device):
        key = (tag, str(device))
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and self.frozen:
                self._retired.append(buf)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def freeze(self):
        self.frozen = True


_default_ws = Workspaces()


def _workspace(tag, nbytes, device, ws=None):
    return (_default_ws if ws is None else ws).get(tag, nbytes, device)


def _as_batch(x, ndim):
    x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    if not x.is_cuda:
        _lib.require_cuda()
        x = x.cuda()
    x = x.to(torch.float64)
    squeeze = x.dim() == ndim - 1
    if squeeze:
        x = x.unsqueeze(0)
    if x.dim() != ndim:
        raise ValueError(f"expected a {ndim - 1}- or {ndim}-dimensional tensor")
    return x.contiguous(), squeeze


def _curv_tensor(curvature, device):
    """Device scalar for the curvature argument.  A Python number becomes a fresh one-element tensor
    filled by a kernel (no pageable host->device copy, hence no stream synchronisation, and nothing
    cached that a captured graph could outlive); a CUDA tensor is used as it is."""
    if isinstance(curvature, torch.Tensor):
        if curvature.is_cuda:
            return curvature.detach().to(device=device, dtype=torch.float64).reshape(-1)[:1].contiguous()
        curvature = float(curvature.detach().reshape(-1)[0])
    return torch.full((1,), float(curvature), dtype=torch.float64, device=device)


def pdm_forward(x, curvature=-1.0, variant="torch", lift="up", matsumoto=False, ws=None):
    """Batched pairwise hyperbolic distances.  x: (B,S,D) or (S,D) -> (B,S,S) or (S,S) on device."""
    lib = _lib.load()
    x, squeeze = _as_batch(x, 3)
    B, S, D = (int(v) for v in x.shape)
    with torch.cuda.device(x.device):
        out = torch.empty((B, S, S), dtype=torch.float64, device=x.device)
        ws = _workspace("pdm", lib.dodo_pdm_workspace_bytes(B, S, D), x.device, ws)
        cv = _curv_tensor(curvature, x.device)
        _lib.check(lib.dodo_pdm_fwd(_ptr(x), B, S, D, _ptr(cv), _lib.PDM_TORCH if variant == "torch" else _lib.PDM_NUMPY,
                                    _lib.LIFT_UP if lift == "up" else _lib.LIFT_WRAP, int(bool(matsumoto)),
                                    _ptr(ws), _ptr(out), _stream()))
    return out[0] if squeeze else out


def pdm_backward(x, grad_pdm, curvature=-1.0, lift="up", matsumoto=False, ws=None):
    """VJP of pdm_forward (torch variant).  Returns (grad_x like x, grad_curvature (B,))."""
    lib = _lib.load()
    x, squeeze = _as_batch(x, 3)
    g, _ = _as_batch(grad_pdm, 3)
    B, S, D = (int(v) for v in x.shape)
    if tuple(g.shape) != (B, S, S):
        raise ValueError("grad_pdm must have shape (B,S,S)")
    with torch.cuda.device(x.device):
        gx = torch.empty_like(x)
        gc = torch.empty(B, dtype=torch.float64, device=x.device)
        ws = _workspace("pdm", lib.dodo_pdm_workspace_bytes(B, S, D), x.device, ws)
        cv = _curv_tensor(curvature, x.device)
        _lib.check(lib.dodo_pdm_bwd(_ptr(x), B, S, D, _ptr(cv), _lib.PDM_TORCH,
                                    _lib.LIFT_UP if lift == "up" else _lib.LIFT_WRAP, int(bool(matsumoto)),
                                    _ptr(ws), _ptr(g), _ptr(gx), _ptr(gc), _stream()))
    return (gx[0] if squeeze else gx), gc


def blens_jacobian(x, peel, flags, curvature=-1.0, lift="up", matsumoto=False, ws=None):
    """d blens / d locations of soft-NJ trees in one pass: x (B,S,D), peel (B,S-1,3), flags (B,S-1) as returned
    by ``nj_soft`` -> (B, 2S-2, S*D).  (``dodo_blens_jacobian``; the reference: vi.py:376-387.)"""
    lib = _lib.load()
    x, squeeze = _as_batch(x, 3)
    B, S, D = (int(v) for v in x.shape)
    peel = peel.reshape(B, S - 1, 3).to(x.device, dtype=torch.int64).contiguous()
    flags = flags.reshape(B, S - 1).to(x.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(x.device):
        J = torch.empty((B, 2 * S - 2, S * D), dtype=torch.float64, device=x.device)
        ws = _workspace("jac", lib.dodo_blens_jacobian_workspace_bytes(B, S, D), x.device, ws)
        cv = _curv_tensor(curvature, x.device)
        _lib.check(lib.dodo_blens_jacobian(_ptr(x), B, S, D, _ptr(cv), _lib.LIFT_UP if lift == "up" else _lib.LIFT_WRAP,
                                           int(bool(matsumoto)), _ptr(peel), _ptr(flags), _ptr(ws), _ptr(J), _stream()))
    return J[0] if squeeze else J


def gram_logdet(J):
    """0.5 * log|det(J J^T)| for a batch of matrices (B, E, C) on the device (``dodo_gram_logdet``); falls back
    to the library LU when E (E+1) doubles do not fit in shared memory."""
    lib = _lib.load()
    B, E, Cc = (int(v) for v in J.shape)
    if (E * (E + 1) + E * 33) * 8 > 227 * 1024 or E > 192:
        return 0.5 * torch.linalg.slogdet(J @ J.transpose(1, 2))[1]
    J = J.contiguous()
    with torch.cuda.device(J.device):
        out = torch.empty(B, dtype=torch.float64, device=J.device)
        _lib.check(lib.dodo_gram_logdet(_ptr(J), B, E, Cc, _ptr(out), _stream()))
    return out


_DIST_FORMS = {"poincare": _lib.DIST_POINCARE, "lorentz": _lib.DIST_LORENTZ, "sheet_up": _lib.DIST_SHEET_UP}


def _pairs(x1, x2):
    x1 = x1 if isinstance(x1, torch.Tensor) else torch.as_tensor(x1)
    x2 = x2 if isinstance(x2, torch.Tensor) else torch.as_tensor(x2)
    if not x1.is_cuda:
        _lib.require_cuda()
    x1 = x1.cuda().to(torch.float64)
    x2 = x2.to(x1.device, dtype=torch.float64)
    if x1.shape != x2.shape or x1.dim() not in (1, 2):
        raise ValueError("x1 and x2 must both be (D,) or (n, D)")
    single = x1.dim() == 1
    return (x1.reshape(-1, x1.shape[-1]).contiguous(), x2.reshape(-1, x2.shape[-1]).contiguous(), single)


def hypdist_forward(x1, x2, curvature=-1.0, form="poincare", matsumoto=False):
    """Distances of n independent pairs of points: x1, x2 (n, D) or (D,) -> (n,) or a 0-dim tensor on the
    device.  ``form``: 'poincare' (Chyp_torch.hyperbolic_distance), 'lorentz'
    (Chyp_torch.hyperbolic_distance_lorentz), 'sheet_up' (Chyp_np.hyperbolic_distance)."""
    lib = _lib.load()
    a, b, single = _pairs(x1, x2)
    n, D = int(a.shape[0]), int(a.shape[1])
    with torch.cuda.device(a.device):
        out = torch.empty(n, dtype=torch.float64, device=a.device)
        _lib.check(lib.dodo_hypdist_fwd(_ptr(a), _ptr(b), n, D, _ptr(_curv_tensor(curvature, a.device)), _DIST_FORMS[form],
                                        int(bool(matsumoto)), _ptr(out), _stream()))
    return out[0] if single else out


def hypdist_backward(x1, x2, grad_out, curvature=-1.0, form="poincare", matsumoto=False):
    """VJP of ``hypdist_forward``: returns (grad_x1, grad_x2 like the inputs, grad_curvature (n,))."""
    lib = _lib.load()
    a, b, single = _pairs(x1, x2)
    n, D = int(a.shape[0]), int(a.shape[1])
    g = grad_out.to(a.device, dtype=torch.float64).reshape(n).contiguous()
    with torch.cuda.device(a.device):
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        gk = torch.empty(n, dtype=torch.float64, device=a.device)
        _lib.check(lib.dodo_hypdist_bwd(_ptr(a), _ptr(b), n, D, _ptr(_curv_tensor(curvature, a.device)), _DIST_FORMS[form],
                                        int(bool(matsumoto)), _ptr(g), _ptr(g1), _ptr(g2), _ptr(gk), _stream()))
    return (g1[0], g2[0], gk) if single else (g1, g2, gk)


class GammaDirPrior:
    """Parameters of the gamma-Dirichlet(aT, bT, a, c) branch-length prior (base_model.py:475-527) for the
    NJ kernels' epilogue / ``prior_gamma_dir``.  ``variant``: 'torch' (utils.LogDirPrior) or 'numpy'
    (Cphylo.compute_prior_gamma_dir_np)."""

    def __init__(self, aT=1.0, bT=None, a=1.0, c=1.0, variant="torch", grad=True):
        if bT is None:
            # the reference's torch default is the float32 tensor torch.full((1,), 0.1) (base_model.py:476), whose
            # value 0.100000001490116 enters the float64 arithmetic as it is; the NumPy twin's is the double 0.1
            bT = torch.full((1,), 0.1) if variant == "torch" else 0.1
        self.aT, self.bT, self.a, self.c = (float(v.reshape(-1)[0]) if isinstance(v, torch.Tensor) else float(v)
                                            for v in (aT, bT, a, c))
        self.variant, self.grad = variant, grad

    def c_struct(self, ln_prior, grad_blens):
        p = _lib.Prior()
        p.aT, p.bT, p.a, p.c = self.aT, self.bT, self.a, self.c
        p.variant = _lib.PRIOR_TORCH if self.variant == "torch" else _lib.PRIOR_NUMPY
        p.d_ln_prior = ln_prior.data_ptr()
        p.d_grad_blens = grad_blens.data_ptr() if grad_blens is not None else None
        return p

    def outputs(self, B, E, device):
        ln_prior = torch.empty(B, dtype=torch.float64, device=device)
        g = torch.empty((B, E), dtype=torch.float64, device=device) if self.grad else None
        return ln_prior, g


def prior_gamma_dir(blens, prior=None):
    """Log prior (B,) and its gradient (B, 2S-2) for a batch of branch lengths already on the device."""
    lib = _lib.load()
    prior = GammaDirPrior() if prior is None else prior
    b, squeeze = _as_batch(blens, 2)
    B, E = int(b.shape[0]), int(b.shape[1])
    with torch.cuda.device(b.device):
        lp, g = prior.outputs(B, E, b.device)
        ps = prior.c_struct(lp, g)
        _lib.check(lib.dodo_prior_gamma_dir(_ptr(b), B, E // 2 + 1, C.byref(ps), _stream()))
    return (lp[0], None if g is None else g[0]) if squeeze else (lp, g)


def nj_hard(pdm, ws=None, prior=None):
    """Batched hard NJ (Cpeeler.nj_np).  pdm (B,S,S) or (S,S) -> peel int64 (B,S-1,3), blens (B,2S-2);
    with ``prior`` (a GammaDirPrior) also (ln_prior (B,), grad (B,2S-2) or None) from the kernel's epilogue."""
    lib = _lib.load()
    pdm, squeeze = _as_batch(pdm, 3)
    B, S = int(pdm.shape[0]), int(pdm.shape[1])
    if pdm.shape[2] != S or S < 2:
        raise ValueError("pdm must be square with at least 2 taxa")
    with torch.cuda.device(pdm.device):
        peel = torch.empty((B, S - 1, 3), dtype=torch.int64, device=pdm.device)
        blens = torch.empty((B, 2 * S - 2), dtype=torch.float64, device=pdm.device)
        ws = _workspace("nj", lib.dodo_nj_workspace_bytes(B, S), pdm.device, ws)
        pr = None
        if prior is not None:
            pr = prior.outputs(B, 2 * S - 2, pdm.device)
            ps = prior.c_struct(*pr)
        _lib.check(lib.dodo_nj_hard(_ptr(pdm), B, S, _ptr(ws), _ptr(peel), _ptr(blens), C.byref(ps) if prior is not None else None,
                                    _stream()))
    if prior is not None:
        return (peel[0], blens[0], pr) if squeeze else (peel, blens, pr)
    return (peel[0], blens[0]) if squeeze else (peel, blens)


def nj_soft(pdm, tau, exhaustive=False, return_flags=False, ws=None, prior=None):
    """Batched soft NJ (peeler.nj_torch with tau).  Returns peel, blens[, flags (B,S-1) int32][, (ln_prior,
    grad_blens) when ``prior`` is given: evaluated in the kernel's epilogue].
    ``exhaustive`` evaluates soft.argmin over all N^2 padded entries instead of the active pairs."""
    lib = _lib.load()
    pdm, squeeze = _as_batch(pdm, 3)
    B, S = int(pdm.shape[0]), int(pdm.shape[1])
    if pdm.shape[2] != S or S < 2:
        raise ValueError("pdm must be square with at least 2 taxa")
    tau = float(tau)
    if not tau > 0.0:
        raise ValueError("tau must be positive")
    with torch.cuda.device(pdm.device):
        peel = torch.zeros((B, S - 1, 3), dtype=torch.int64, device=pdm.device)
        blens = torch.zeros((B, 2 * S - 2), dtype=torch.float64, device=pdm.device)
        flags = torch.zeros((B, S - 1), dtype=torch.int32, device=pdm.device)
        ws = _workspace("nj", lib.dodo_nj_workspace_bytes(B, S), pdm.device, ws)
        pr = None
        if prior is not None:
            pr = prior.outputs(B, 2 * S - 2, pdm.device)
            ps = prior.c_struct(*pr)
        _lib.check(lib.dodo_nj_soft(_ptr(pdm), B, S, -tau if exhaustive else tau, _ptr(ws), _ptr(peel), _ptr(blens),
                                    _ptr(flags), C.byref(ps) if prior is not None else None, _stream()))
    if squeeze:
        peel, blens, flags = peel[0], blens[0], flags[0]
    out = (peel, blens, flags) if return_flags else (peel, blens)
    return out + (pr,) if prior is not None else out


def nj_soft_backward(peel, flags, grad_blens, ws=None):
    """VJP of soft-NJ branch lengths w.r.t. the distance matrix: returns symmetric (B,S,S)."""
    lib = _lib.load()
    g, squeeze = _as_batch(grad_blens, 2)
    B = int(g.shape[0])
    S = int(g.shape[1]) // 2 + 1
    peel = peel.reshape(B, S - 1, 3).to(device=g.device, dtype=torch.int64).contiguous()
    flags = flags.reshape(B, S - 1).to(device=g.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(g.device):
        out = torch.empty((B, S, S), dtype=torch.float64, device=g.device)
        ws = _workspace("nj", lib.dodo_nj_workspace_bytes(B, S), g.device, ws)
        _lib.check(lib.dodo_nj_soft_bwd(_ptr(peel), _ptr(flags), _ptr(g), B, S, _ptr(ws), _ptr(out), _stream()))
    return out[0] if squeeze else out


def nj_guard(flags, loglik=None, grad_blens=None, mask=12):
    """Poison (NaN) lnL / gradient rows of the samples whose soft-NJ flags (B, S-1) carry a bit of
    ``mask`` and count them: returns an int32 device tensor [#guard tripped (bit 2), #invalid selection
    (bit 3)].  No host synchronisation (``HotPath.resolve`` is the host side of the policy)."""
    lib = _lib.load()
    B, nj = int(flags.shape[0]), int(flags.shape[1])
    with torch.cuda.device(flags.device):
        counts = torch.empty(2, dtype=torch.int32, device=flags.device)
        _lib.check(lib.dodo_nj_guard(_ptr(flags), B, nj + 1, int(mask), _ptr(loglik), _ptr(grad_blens), _ptr(counts), _stream()))
    return counts


# --------------------------------------------------------------------------------------------
# geodesic connector (K2g)
# --------------------------------------------------------------------------------------------
def geo_soft(x, tau=1e-8, ws=None):
    """Batched peeler.make_soft_peel_tips: ball locations (B,S,D) or (S,D) ->
    dict(peel, int_locs, blens, flags, slots, tape); slots/tape feed ``geo_soft_backward``."""
    lib = _lib.load()
    x, squeeze = _as_batch(x, 3)
    B, S, D = (int(v) for v in x.shape)
    if S < 2:
        raise ValueError("need at least 2 taxa")
    with torch.cuda.device(x.device):
        dev = x.device
        out = dict(peel=torch.zeros((B, S - 1, 3), dtype=torch.int64, device=dev),
                   int_locs=torch.zeros((B, S - 1, D), dtype=torch.float64, device=dev),
                   blens=torch.zeros((B, 2 * S - 2), dtype=torch.float64, device=dev),
                   flags=torch.zeros((B, S - 1), dtype=torch.int32, device=dev),
                   slots=torch.empty((B, S - 1, 2), dtype=torch.int32, device=dev),
                   tape=torch.zeros((B, S - 1, 2, D), dtype=torch.float64, device=dev))
        ws = _workspace("geo", lib.dodo_geo_workspace_bytes(B, S, D), dev, ws)
        _lib.check(lib.dodo_geo_soft(_ptr(x), B, S, D, float(tau), _ptr(ws), _ptr(out["peel"]), _ptr(out["int_locs"]),
                                     _ptr(out["blens"]), _ptr(out["slots"]), _ptr(out["tape"]), _ptr(out["flags"]),
                                     _stream()))
    if squeeze:
        out = {k: v[0] for k, v in out.items()}
    return out


def geo_soft_backward(peel, slots, tape, grad_blens, grad_int_locs=None):
    """VJP of (blens, int_locs) w.r.t. the ball locations: returns (B,S,D)."""
    lib = _lib.load()
    g, squeeze = _as_batch(grad_blens, 2)
    B, S = int(g.shape[0]), int(g.shape[1]) // 2 + 1
    D = int(tape.shape[-1])
    dev = g.device
    peel = peel.reshape(B, S - 1, 3).to(device=dev, dtype=torch.int64).contiguous()
    slots = slots.reshape(B, S - 1, 2).to(device=dev, dtype=torch.int32).contiguous()
    tape = tape.reshape(B, S - 1, 2, D).to(device=dev, dtype=torch.float64).contiguous()
    gi = None if grad_int_locs is None else grad_int_locs.reshape(B, S - 1, D).to(device=dev, dtype=torch.float64).contiguous()
    with torch.cuda.device(dev):
        gx = torch.empty((B, S, D), dtype=torch.float64, device=dev)
        _lib.check(lib.dodo_geo_soft_bwd(_ptr(peel), _ptr(slots), _ptr(tape), _ptr(g), _ptr(gi), B, S, D, _ptr(gx), _stream()))
    return gx[0] if squeeze else gx


def geo_hard(sheet):
    """Batched peeler.make_hard_peel_geodesic: hyperboloid points (B,S,D+1) or (S,D+1) ->
    (peel (B,S-1,3), int_locs (B,S-1,D+1))."""
    lib = _lib.load()
    z, squeeze = _as_batch(sheet, 3)
    B, S, D1 = (int(v) for v in z.shape)
    if S < 2:
        raise ValueError("need at least 2 taxa")
    with torch.cuda.device(z.device):
        peel = torch.zeros((B, S - 1, 3), dtype=torch.int64, device=z.device)
        il = torch.zeros((B, S - 1, D1), dtype=torch.float64, device=z.device)
        ws = _workspace("geo", lib.dodo_geo_workspace_bytes(B, S, D1 - 1), z.device)
        _lib.check(lib.dodo_geo_hard(_ptr(z), B, S, D1, _ptr(ws), _ptr(peel), _ptr(il), _stream()))
    return (peel[0], il[0]) if squeeze else (peel, il)

"""Differentiable neighbour joining with the reference's interface (dodonaphy/peeler.py:155-220):
``nj_torch(dists_leaves, tau=None) -> (peel ndarray int (S-1,3), blens tensor (2S-2,))``.

tau=None returns the hard NJ result as NumPy arrays, like the reference (peeler.py:166-167).
With a temperature the join selection is soft.argmin's (dodonaphy/soft.py:38-53), evaluated on the
device in O(active pairs) per join; ``blens`` carries gradient to ``dists_leaves`` through a
hand-written reverse replay of the joins (the selections themselves pass no gradient, soft.py:56-61).
"""
import numpy as np
import torch

from . import Cpeeler
from . import engine as _engine


class _SoftNJ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dists, tau, exhaustive):
        peel, blens, flags = _engine.nj_soft(dists.detach(), tau, exhaustive=exhaustive, return_flags=True)
        bad = int((flags & 12).max().item())
        if bad & 8:
            raise RuntimeError("soft NJ: the soft argmin selected an inactive node (temperature too high for these distances)")
        if bad & 4:
            # the closed-form treatment of inactive entries is not provably exact here: redo exhaustively
            peel, blens, flags = _engine.nj_soft(dists.detach(), tau, exhaustive=True, return_flags=True)
            if int((flags & 8).max().item()):
                raise RuntimeError("soft NJ: the soft argmin selected an inactive node")
        ctx.save_for_backward(peel, flags)
        ctx.dev = dists.device
        ctx.mark_non_differentiable(peel)
        return peel, blens.to(dists.device)

    @staticmethod
    def backward(ctx, _gpeel, gblens):
        peel, flags = ctx.saved_tensors
        g = _engine.nj_soft_backward(peel, flags, gblens.contiguous())
        return g.to(ctx.dev), None, None


def nj_torch(dists_leaves, tau=None, exhaustive=False):
    if tau is None:
        d = dists_leaves.detach().cpu().numpy() if isinstance(dists_leaves, torch.Tensor) else dists_leaves
        return Cpeeler.nj_np(d)
    if not isinstance(dists_leaves, torch.Tensor):
        dists_leaves = torch.as_tensor(dists_leaves, dtype=torch.float64)
    peel, blens = _SoftNJ.apply(dists_leaves, float(tau), bool(exhaustive))
    return peel.cpu().numpy().astype(int), blens


def compute_Q(pdm, mask=False, fill_value=None):
    """NJ Q-matrix of the active nodes (peeler.py:223-248); host helper kept for API parity/tests."""
    pdm = torch.as_tensor(pdm, dtype=torch.float64)
    n = len(pdm)
    mask = torch.zeros(n, dtype=torch.bool) if mask is False else torch.as_tensor(mask, dtype=torch.bool)
    live = ~mask
    both = torch.outer(live, live)
    r = (pdm * both).sum(dim=1, keepdim=True)
    Q = (int(live.sum()) - 2) * pdm - r - r.T
    fill = 1e9 if fill_value is None else fill_value
    Q = torch.where(both, Q, torch.full_like(Q, fill))
    Q.fill_diagonal_(fill)
    return 0.5 * (Q + Q.T)


# ------------------------------------------------------------------------------------------------
# geodesic connector (peeler.py:14-152)
class _SoftGeodesic(torch.autograd.Function):
    @staticmethod
    def forward(ctx, leaf_locs, tau):
        out = _engine.geo_soft(leaf_locs.detach(), tau)
        if int((out["flags"] & 8).max().item()):
            raise RuntimeError("geodesic connector: the soft argmin selected a retired point (near-tied LCA depths)")
        ctx.save_for_backward(out["peel"], out["slots"], out["tape"])
        ctx.dev = leaf_locs.device
        ctx.mark_non_differentiable(out["peel"])
        return out["peel"], out["int_locs"].to(leaf_locs.device), out["blens"].to(leaf_locs.device)

    @staticmethod
    def backward(ctx, _gpeel, g_int, g_blens):
        peel, slots, tape = ctx.saved_tensors
        gx = _engine.geo_soft_backward(peel, slots, tape, g_blens.contiguous(), None if g_int is None else g_int.contiguous())
        return gx.to(ctx.dev), None


def make_soft_peel_tips(leaf_locs, connector="geodesics", curvature=-torch.ones(1)):
    """(S,D) [or (B,S,D)] Poincare-ball locations -> (peel ndarray int, int_locs, blens); the internal
    locations and branch lengths carry gradient to ``leaf_locs`` (peeler.py:14-70)."""
    if connector != "geodesics":
        raise NotImplementedError(f"Connector must be geodesics, got {connector}")
    if not torch.isclose(torch.as_tensor(curvature, dtype=torch.float64), -torch.ones(1, dtype=torch.float64)).all():
        raise NotImplementedError(f"Curvature must be -1, got {curvature}")
    leaf_locs = torch.as_tensor(leaf_locs, dtype=torch.float64)
    peel, int_locs, blens = _SoftGeodesic.apply(leaf_locs, 1e-8)
    return peel.cpu().numpy().astype(int), int_locs, blens


def make_hard_peel_geodesic(leaf_locs, matsumoto=False):
    """(S,D+1) hyperboloid points -> (peel ndarray int (S-1,3), int_locs ndarray (S-1,D+1))
    (peeler.py:73-152).  NumPy in, NumPy out, as the MCMC side calls it."""
    if matsumoto:
        raise NotImplementedError("the Matsumoto adjustment of the heap keys is not on the path")
    z = np.ascontiguousarray(leaf_locs, dtype=np.float64)
    peel, int_locs = _engine.geo_hard(torch.from_numpy(z))
    return peel.cpu().numpy().astype(int), int_locs.cpu().numpy()

"""Drop-in front end for the reference's likelihood functions (dodonaphy/phylo.py).

``calculate_treelikelihood(partials, weights, post_indexing, mats, freqs)``   phylo.py:132-139
    same arguments, same 0-dim float64 result, autograd to ``mats`` and ``freqs``; the work is done
    by the CUDA pruning kernel.  With a category axis in ``mats`` the reference's broadcasting sums
    the per-category log-likelihoods (phylo.py:137-139) -- reproduced.
``TreeLikelihood``   phylo.py:177-226 (the reference's own native-backend plugin shape)
    ``TreeLikelihood.apply(blens, peel, model, sub_rates, freqs)``: value AND gradients are computed
    in ``forward`` by the fused kernel, ``backward`` scales them by ``grad_output``.
``compress_alignment``   phylo.py:16-58: alignment -> tip partials + pattern weights.

Tensors may live on the CPU (as in the reference) or on the GPU; results come back on the device of
the inputs.  There is no CPU fallback: without CUDA these functions raise.
"""
from collections import Counter

import numpy as np
import torch

from . import engine as _engine

_IUPAC = {"A": 1, "C": 2, "G": 4, "T": 8, "R": 5, "Y": 10, "M": 3, "W": 9, "S": 6, "K": 12,
          "B": 14, "D": 13, "H": 11, "V": 7, "N": 15, "?": 15, "-": 15}


def compress_alignment(alignment, get_namespace=False, get_freqs=False):
    """Unique site patterns (sorted) with their counts; tips as (4, L) 0/1 float64 tensors in taxon
    order -- the representation BaseModel is built from.  ``alignment`` is duck-typed like the
    reference's use of a DendroPy character matrix: ``.sequences()`` and
    ``.taxon_namespace.labels()``."""
    rows = [str(s) for s in alignment.sequences()]
    counts = Counter(zip(*rows))
    patterns = sorted(counts)
    weights = torch.tensor(np.array([counts[c] for c in patterns]))
    partials = []
    for tip in range(len(rows)):
        mask = np.array([_IUPAC.get(col[tip].upper(), 15) for col in patterns], dtype=np.uint8)
        bits = np.stack([(mask >> j) & 1 for j in range(4)]).astype(np.float64)
        partials.append(torch.tensor(bits))
    if get_namespace:
        return partials, weights, alignment.taxon_namespace
    return partials, weights


def compress_alignment_device(sequences, device="cuda"):
    """Alignment ingestion on the device: the same unique-sorted-pattern compression as
    ``compress_alignment`` (phylo.py:16-58) for a list of equal-length strings, with the column sort / unique
    done by the library's radix-sort kernels (``dodo_alignment_sort`` / ``dodo_alignment_gather``,
    csrc/align.cu); the result is what the pruning kernel consumes directly:
    (tip masks uint8 (S, L) on ``device``, pattern weights int64 (L,) on ``device``).
    Patterns are ordered like the reference's ``sorted(Counter(zip(*sequences)))`` (byte order of the
    raw characters); the character -> state-set map is case-insensitive."""
    from . import _lib

    S = len(sequences)
    n = len(sequences[0])
    if any(len(q) != n for q in sequences):
        raise ValueError("sequences must have equal length")
    _lib.require_cuda()
    lib = _lib.load()
    raw = torch.frombuffer(bytearray("".join(sequences), "ascii"), dtype=torch.uint8).reshape(S, n).to(device).contiguous()
    lut = torch.full((256,), 15, dtype=torch.uint8)
    for ch, m in _IUPAC.items():
        lut[ord(ch)] = m
        lut[ord(ch.lower())] = m
    lut = lut.to(raw.device)
    with torch.cuda.device(raw.device):
        stream = torch.cuda.current_stream().cuda_stream
        ws = torch.empty(int(lib.dodo_alignment_workspace_bytes(S, n)), dtype=torch.uint8, device=raw.device)
        n_unique = torch.zeros(1, dtype=torch.int32, device=raw.device)
        _lib.check(lib.dodo_alignment_sort(raw.data_ptr(), S, n, ws.data_ptr(), n_unique.data_ptr(), stream))
        U = int(n_unique.item())
        masks = torch.empty((S, U), dtype=torch.uint8, device=raw.device)
        weights = torch.empty(U, dtype=torch.int64, device=raw.device)
        _lib.check(lib.dodo_alignment_gather(raw.data_ptr(), S, n, ws.data_ptr(), U, lut.data_ptr(), masks.data_ptr(),
                                             weights.data_ptr(), stream))
    return masks, weights


# ------------------------------------------------------------------------------------------------
_engine_cache = []  # [(weakrefs of the tip objects + weights, versions, engine)]


def _versions(objs):
    return tuple(int(o._version) if isinstance(o, torch.Tensor) else 0 for o in objs)


def _engine_for(partials, weights, n_tips):
    """PruneEngine for this alignment; tips are uploaded once and reused (the reference's partial
    list is owned by the model and re-passed on every call, base_model.py:54).  An entry is reused
    only for the very same tip / weight objects, unmodified since (weak references + tensor
    versions), so recycled memory can never alias a stale upload."""
    import weakref

    objs = list(partials[:n_tips]) + [weights]
    for refs, vers, eng in _engine_cache:
        if len(refs) == len(objs) and all(r() is o for r, o in zip(refs, objs)) and vers == _versions(objs):
            return eng
    w = weights.detach().cpu().numpy() if isinstance(weights, torch.Tensor) else np.asarray(weights)
    eng = _engine.PruneEngine(_engine.pack_tip_partials(objs[:-1]), w)
    try:
        refs = tuple(weakref.ref(o) for o in objs)
    except TypeError:  # not weak-referenceable (e.g. a list): do not cache
        return eng
    _engine_cache[:] = [e for e in _engine_cache if all(r() is not None for r in e[0])][-7:]
    _engine_cache.append((refs, _versions(objs), eng))
    return eng


def _as_peel_tensor(post_indexing, n_tips, device):
    peel = post_indexing.detach().cpu().numpy() if isinstance(post_indexing, torch.Tensor) else np.asarray(post_indexing)
    peel = np.ascontiguousarray(peel, dtype=np.int64)
    _engine.validate_peel(peel, n_tips)
    return torch.from_numpy(peel).to(device)


class _MatsLikelihood(torch.autograd.Function):
    """lnL as a function of explicit transition matrices and root frequencies."""

    @staticmethod
    def forward(ctx, mats, freqs, eng, peel):
        E = eng.E
        m = mats.detach().to(eng.device, dtype=torch.float64).reshape(E, -1, 4, 4).unsqueeze(0).contiguous()
        pi = freqs.detach().cpu().numpy().astype(np.float64)
        want = mats.requires_grad or freqs.requires_grad
        out = eng.loglik_mats(peel, m, pi, grad=want)
        if want:
            ctx.save_for_backward(out["grad_mats"][0].reshape(mats.shape).to(mats.device),
                                  out["grad_pi"][0].to(freqs.device))
        return out["lnL"][0].to(mats.device)

    @staticmethod
    def backward(ctx, grad_output):
        g_mats, g_pi = ctx.saved_tensors
        return grad_output * g_mats, grad_output * g_pi, None, None


def calculate_treelikelihood(partials, weights, post_indexing, mats, freqs):
    n_tips = len(post_indexing) + 1
    if len(partials) < n_tips:
        raise ValueError("need one partial per tip")
    eng = _engine_for(partials, weights, n_tips)
    peel = _as_peel_tensor(post_indexing, n_tips, eng.device)
    mats = torch.as_tensor(mats)
    freqs = torch.as_tensor(freqs, dtype=torch.float64)
    if mats.shape[0] != eng.E or tuple(mats.shape[-2:]) != (4, 4):
        raise ValueError(f"mats must have shape ({eng.E}, [K,] 4, 4)")
    return _MatsLikelihood.apply(mats, freqs, eng, peel)


class TreeLikelihood(torch.autograd.Function):
    """Value and branch-length gradient in one fused kernel call (shape of phylo.py:177-226).

    ``TreeLikelihood.apply(blens, peel, inst)`` where ``inst`` is a ``LikelihoodInstance`` (the
    analogue of the bito instance the reference passes)."""

    @staticmethod
    def forward(ctx, blens, peel, inst):
        eng = inst.engine
        b = blens.detach().to(eng.device, dtype=torch.float64)
        peel_t = peel if isinstance(peel, torch.Tensor) and peel.is_cuda else _as_peel_tensor(peel, eng.S, eng.device)
        ll, g = eng.loglik(peel_t, b, inst.model, rates=inst.rates, mix=inst.mix, grad=blens.requires_grad)
        if blens.requires_grad:
            ctx.save_for_backward(g.reshape(blens.shape).to(blens.device))
        out = ll.to(blens.device)
        return out[0] if blens.dim() == 1 else out

    @staticmethod
    def backward(ctx, grad_output):
        (g,) = ctx.saved_tensors
        go = grad_output if grad_output.dim() == 0 else grad_output.unsqueeze(-1)
        return go * g, None, None


class LikelihoodInstance:
    """Alignment + substitution model bound to a device engine."""

    def __init__(self, partials, weights, model, rates=None, mix=False):
        tips = list(partials)  # one (4, L) partial per tip
        w = weights.detach().cpu().numpy() if isinstance(weights, torch.Tensor) else np.asarray(weights)
        self.engine = _engine.PruneEngine(_engine.pack_tip_partials(tips), w)
        self.model = model
        self.rates = rates
        self.mix = mix


def get_parent_id_vector(peel, rooted=False):
    """Parent id of every node (phylo.py:142-174); an unrooted tree's fake root is folded into
    the last real internal node."""
    peel = np.asarray(peel)
    n_nodes = 2 * len(peel) - (0 if rooted else 1)
    fake, real = int(peel[-1][2]), int(peel[-2][2]) if len(peel) > 1 else int(peel[-1][2])
    parent = [0] * n_nodes
    for c0, c1, p in peel:
        p = real if (not rooted and p == fake) else int(p)
        for c in (int(c0), int(c1)):
            if c < n_nodes and parent[c] == 0:
                parent[c] = p
    return parent

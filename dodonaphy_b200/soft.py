"""SoftSort helpers with the reference's names (dodonaphy/soft.py).

``argmin`` -- the one on the hot path (it defines soft NJ's tie-breaking, peeler.py:194) -- runs on
the device: the reference builds two n x n SoftSort matrices and reads only their last rows; the
kernel evaluates exactly those rows in O(n).  ``sort`` / ``min`` / ``max`` / ``clamp_pos`` are the
small dense helpers the reference's tests exercise (test/test_soft.py); they stay differentiable
torch code.
"""
import ctypes as C

import torch

from . import _lib
from .engine import _ptr, _stream

eps = torch.finfo(torch.double).eps
eps_torch = torch.tensor(eps)


def unravel_index(index, shape):
    return torch.div(index, shape[-1], rounding_mode="trunc"), torch.fmod(index, shape[-1])


def sort(s, tau):
    """Relaxed permutation matrix of a (batch, n, 1) tensor, rows in descending order of s."""
    ordered = s.sort(descending=True, dim=1)[0]
    return ((s.transpose(1, 2) - ordered).abs().neg() / tau).softmax(-1)


def min(s, tau):
    assert s.ndim == 1
    return (sort(s.reshape(1, -1, 1), tau) @ s)[0, -1]


def max(s, tau):
    return -min(-s, tau)


def clamp_pos(s, tau, epsilon=eps_torch):
    assert s.ndim == 0
    return max(torch.stack((s, epsilon.to(s.dtype).reshape(()))), tau)


def argmin(Q, tau):
    """Soft (row, col) of the minimum of a 2-D tensor, ties broken towards the last entry;
    returns two 0-dim float64 tensors like the reference (no gradient: the caller rounds them)."""
    assert Q.ndim == 2
    lib = _lib.load()
    _lib.require_cuda()
    q = Q.detach().to("cuda", dtype=torch.float64).contiguous()
    out = torch.empty(2, dtype=torch.float64, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(lib.dodo_soft_argmin(_ptr(q), 1, int(q.shape[0]), int(q.shape[1]), float(tau), _ptr(out), _stream()))
    out = out.to(Q.device)
    return out[0], out[1]


def index_to_hard_one_hot(index, size):
    """Float index -> one-hot vector; passes no gradient (soft.py:56-61)."""
    hot = torch.zeros(size, dtype=torch.double)
    hot[int(torch.round(index.detach()))] = 1.0
    return hot

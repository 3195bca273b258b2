"""Hard neighbour joining with the reference's NumPy interface (dodonaphy/cython/Cpeeler.pyx:15-125):
``nj_np(pdm) -> (peel int64 (S-1,3), blens float64 (2S-2,))``, bit-identical to the reference on the
same matrix; computed by the CUDA NJ kernel (one tree per CTA; a leading batch axis is accepted)."""
import numpy as np
import torch

from . import engine as _engine


def nj_np(pdm):
    pdm = np.ascontiguousarray(pdm, dtype=np.float64)
    if pdm.ndim not in (2, 3) or pdm.shape[-1] != pdm.shape[-2]:
        raise ValueError("pdm must be a square matrix (or a batch of them)")
    peel, blens = _engine.nj_hard(torch.from_numpy(pdm))
    return peel.cpu().numpy(), blens.cpu().numpy()

"""Torch-facing hyperboloid utilities of the VI path (dodonaphy/cython/Chyp_torch.pyx).

``get_pdm`` (Chyp_torch.pyx:399-423) is an autograd function backed by the CUDA distance kernel and
its hand-written VJP: gradients reach the locations and the curvature, exactly the leaves the
reference's autograd reaches.  Pointwise helpers are kept for API parity (tiny host arithmetic).
"""
import torch

from . import engine as _engine

eps = torch.tensor(torch.finfo(torch.double).eps)


class _PdmFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, leaf_loc, curvature, matsumoto, projection):
        x = leaf_loc.detach()
        kappa = float(curvature.detach().reshape(-1)[0])
        out = _engine.pdm_forward(x, kappa, "torch", projection, matsumoto)
        ctx.save_for_backward(x)
        ctx.meta = (kappa, matsumoto, projection, leaf_loc.device, curvature.shape, curvature.device)
        return out.to(leaf_loc.device)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        kappa, matsumoto, projection, dev, cshape, cdev = ctx.meta
        gx, gc = _engine.pdm_backward(x, grad_out.contiguous(), kappa, projection, matsumoto)
        return gx.to(dev), gc.sum().reshape(cshape).to(cdev), None, None


def get_pdm(leaf_loc, curvature=-torch.ones(1), matsumoto=False, projection="up"):
    """(S,D) [or (B,S,D)] locations -> pairwise hyperbolic distances, differentiable."""
    if projection not in ("up", "wrap"):
        raise ValueError("projection must be 'up' or 'wrap'")
    curv = curvature if isinstance(curvature, torch.Tensor) else torch.tensor([float(curvature)], dtype=torch.float64)
    return _PdmFunction.apply(leaf_loc, curv.to(torch.float64), bool(matsumoto), projection)


def project_up(loc):
    z = torch.sqrt(torch.sum(loc * loc, -1, keepdim=True) + 1)
    return torch.cat((z, loc), dim=-1)


def project_up_2d(loc):
    return project_up(loc).unsqueeze(-1) if loc.ndim == 1 else project_up(loc)


def lorentz_product(x, y=None):
    y = x if y is None else y
    return -x[0] * y[0] + torch.dot(x[1:], y[1:])


def poincare_to_hyper(location):
    a = location.pow(2).sum(dim=-1, keepdim=True)
    return torch.cat(((1 + a) / (1 - a), 2 * location / (1 - a + eps)), dim=-1)


poincare_to_hyper_2d = poincare_to_hyper


def hyper_to_poincare(location):
    return location[1:] / (1 + location[0])


def hyperbolic_distance_lorentz(x1, x2, curvature=-torch.ones(1)):
    """Poincare-ball points, distance through the hyperboloid (Chyp_torch.pyx:137-154)."""
    if torch.isclose(curvature, torch.zeros(1, dtype=curvature.dtype)).all():
        return torch.norm(x2 - x1)
    inner = torch.clamp(-lorentz_product(poincare_to_hyper(x1), poincare_to_hyper(x2)), min=1.0 + eps)
    return 1.0 / torch.sqrt(-curvature) * torch.acosh(inner)


def hyperbolic_distance(x1, x2, curvature=-torch.ones(1)):
    """Poincare-ball formula at curvature -1, Lorentz form otherwise (Chyp_torch.pyx:119-134)."""
    if abs(curvature + 1.0) > 1e-9:
        return hyperbolic_distance_lorentz(x1, x2, curvature)
    inv = 2 * torch.sum((x2 - x1) ** 2) / (1 - torch.linalg.norm(x1) ** 2) / (1 - torch.linalg.norm(x2) ** 2)
    if torch.isnan(inv):
        return hyperbolic_distance_lorentz(x1, x2, curvature)
    return torch.acosh(1 + inv)

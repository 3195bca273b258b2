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


class _PairDistance(torch.autograd.Function):
    """Distances of n independent pairs through the CUDA kernel (csrc/hypdist.cu) with its VJP."""

    @staticmethod
    def forward(ctx, x1, x2, curvature, form):
        kappa = float(curvature.detach().reshape(-1)[0])
        a, b = x1.detach(), x2.detach()
        out = _engine.hypdist_forward(a, b, kappa, form)
        ctx.save_for_backward(a, b)
        ctx.meta = (kappa, form, x1.device, x2.device, curvature.shape, curvature.device)
        return out.to(x1.device)

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.saved_tensors
        kappa, form, d1, d2, cshape, cdev = ctx.meta
        g1, g2, gk = _engine.hypdist_backward(a, b, grad_out.contiguous(), kappa, form)
        return g1.to(d1), g2.to(d2), gk.sum().reshape(cshape).to(cdev), None


def _pair_distance(x1, x2, curvature, form):
    curv = curvature if isinstance(curvature, torch.Tensor) else torch.tensor([float(curvature)], dtype=torch.float64)
    x1 = torch.as_tensor(x1, dtype=torch.float64)
    x2 = torch.as_tensor(x2, dtype=torch.float64)
    return _PairDistance.apply(x1, x2, curv.to(torch.float64), form)


def hyperbolic_distance_lorentz(x1, x2, curvature=-torch.ones(1)):
    """Poincare-ball points, distance through the hyperboloid (Chyp_torch.pyx:137-154).  x1, x2: (D,) as
    in the reference, or (n, D) for n pairs at once; differentiable w.r.t. both points and the curvature."""
    return _pair_distance(x1, x2, curvature, "lorentz")


def hyperbolic_distance(x1, x2, curvature=-torch.ones(1)):
    """Poincare-ball formula at curvature -1, Lorentz form otherwise (Chyp_torch.pyx:119-134); (D,) or
    batched (n, D) like ``hyperbolic_distance_lorentz``."""
    return _pair_distance(x1, x2, curvature, "poincare")


# ---------------------------------------------------------------------------------------------
# Chart changes around the "wrap" embedding (Chyp_torch.pyx:39-117, 212-362, 387-396, 425-457).
# The reference walks the taxa in a Python loop; these act on whole (n, .) arrays at once.  The
# distance kernel applies the same lift on the device (pdm.cu, lift "wrap"); these host versions
# exist for the sampling-side code of the callers (vi.py:343-375) and keep the reference's names.
def _lorentz_rows(x, y):
    return -x[..., 0] * y[..., 0] + (x[..., 1:] * y[..., 1:]).sum(-1)


def parallel_transport(xi, x, y):
    """Move the tangent vector ``xi`` from T_x to T_y along the geodesic (Nagano et al. 2019)."""
    alpha = -_lorentz_rows(x, y)
    coef = _lorentz_rows(y, xi) / (alpha + 1)
    return xi + coef.unsqueeze(-1) * (x + y)


def exponential_map(x, v):
    """Point reached from ``x`` along the tangent vector ``v``."""
    vnorm = torch.sqrt(torch.clamp(_lorentz_rows(v, v), min=eps)).unsqueeze(-1)
    return torch.cosh(vnorm) * x + torch.sinh(vnorm) * v / vnorm


def exp_map_inverse(z, mu):
    """Tangent vector at ``mu`` pointing to the sheet point ``z``."""
    alpha = -_lorentz_rows(mu, z).unsqueeze(-1)
    return torch.acosh(alpha) / torch.sqrt(alpha**2 - 1) * (z - alpha * mu)


def tangent_to_hyper(mu, v_tilde, dim=None):
    """Carry v_tilde in T_0 (given by its ``dim`` spatial coordinates) to T_mu and shoot onto the
    sheet.  As in the reference the lifted vector's time coordinate is ONE, not zero
    (Chyp_torch.pyx:343): transport from the origin removes exactly that component when mu is the
    origin, and otherwise it is part of the reference's arithmetic that parity has to keep."""
    v_tilde = v_tilde.reshape(mu.shape[:-1] + (mu.shape[-1] - 1,))
    one = torch.ones(v_tilde.shape[:-1] + (1,), dtype=v_tilde.dtype)
    mu0 = torch.cat((one, torch.zeros_like(v_tilde)), -1)
    u = parallel_transport(torch.cat((one, v_tilde), -1), mu0, mu)
    return exponential_map(mu, u)


def tangent_to_hyper_2d(loc_t0):
    loc_t0 = loc_t0.unsqueeze(0) if loc_t0.ndim == 1 else loc_t0
    return tangent_to_hyper(project_up(torch.zeros_like(loc_t0)), loc_t0)


def tangent_to_hyper_jacobian(mu, v_tilde, dim):
    """log |det| of tangent_to_hyper: (sinh r / r)^(dim-1), r = |v_tilde|."""
    r = torch.norm(v_tilde, dim=-1)
    return (dim - 1) * torch.log(torch.sinh(r) / r)


def hyper_to_poincare_jacobian(location):
    """log |det| of the sheet -> ball projection as the reference evaluates it (:310-322)."""
    n = location.shape[-1]
    a = 1 + location[..., 0]
    return torch.log(torch.abs((a**2 + (location[..., 1:] ** 2).sum(-1)) / a ** (2 * (n + 1))))


def ball2real(loc_ball, radius=1):
    loc_ball = loc_ball.unsqueeze(-1) if loc_ball.ndim == 1 else loc_ball
    return loc_ball / (radius - torch.norm(loc_ball, dim=-1, keepdim=True))


def real2ball(loc_real, radius=1):
    loc_real = loc_real.unsqueeze(-1) if loc_real.ndim == 1 else loc_real
    return radius * loc_real / (1 + torch.norm(loc_real, dim=-1, keepdim=True))


def real2ball_LADJ(y, radius=1.0):
    """log |det J| of real2ball summed over the points: J_k = r (I - y y^T / (|y|(|y|+1))) / (1+|y|),
    whose determinant is r^D / (1+|y|)^(D+1) by the matrix determinant lemma."""
    y = y.unsqueeze(-1) if y.ndim == 1 else y
    n, D = y.shape
    norm = torch.norm(y, dim=-1)
    return (n * D * torch.log(torch.as_tensor(radius, dtype=y.dtype)) - (D + 1) * torch.log1p(norm).sum()).reshape(1)


def t02p(x, mu=None, get_jacobian=False):
    """Tangent space at the origin -> sheet (wrapped around ``mu``) -> Poincare ball, row-wise."""
    mu = torch.zeros_like(x) if mu is None else mu
    dim = x.shape[1]
    x_hyp = tangent_to_hyper(project_up(mu), x)
    x_poin = x_hyp[:, 1:] / (1 + x_hyp[:, :1])
    if not get_jacobian:
        return x_poin
    jac = tangent_to_hyper_jacobian(None, x, dim).sum() + hyper_to_poincare_jacobian(x_hyp).sum()
    return x_poin, jac.reshape(1)


def p2t0(loc, mu=None, get_jacobian=False):
    """Poincare ball -> tangent space at the origin (inverse of ``t02p``)."""
    vec_hyp = poincare_to_hyper(loc)
    mu = project_up(torch.zeros_like(loc) if mu is None else mu)
    zero = torch.zeros_like(mu)
    zero[:, 0] = 1
    out = parallel_transport(exp_map_inverse(vec_hyp, mu), mu, zero)[:, 1:]
    if get_jacobian:
        _, jacobian = t02p(out, mu[:, 1:], get_jacobian=True)
        return out, -jacobian
    return out

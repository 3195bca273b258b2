"""Oracle: pairwise hyperbolic distance matrix.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in plain NumPy float64,

* ``Chyp_torch.get_pdm``  (dodonaphy/cython/Chyp_torch.pyx:399-423) with the "up" lift
  ``project_up_2d`` (:486-504) and the "wrap" lift ``tangent_to_hyper_2d`` (:387-396, via
  ``tangent_to_hyper`` :325-346, ``parallel_transport`` :39-62, ``exponential_map`` :65-101);
* ``Chyp_np.get_pdm``     (dodonaphy/cython/Chyp_np.pyx:321-367).

and gives the analytic gradient of the torch variant (what torch autograd computes through
Chyp_torch.pyx:411-423) for checking the CUDA backward.
"""
import numpy as np

EPS = np.finfo(np.float64).eps
TORCH_CLAMP = 1.0000001  # Chyp_torch.pyx:418


def lift_up(x):
    """project_up_2d: z = sqrt(|x|^2 + 1); sheet = [z, x]   (Chyp_torch.pyx:501-504)."""
    z = np.sqrt(np.sum(x * x, axis=1, keepdims=True) + 1.0)
    return np.concatenate([z, x], axis=1)


def lift_wrap(x):
    """tangent_to_hyper_2d at the origin (Chyp_torch.pyx:387-396).

    mu = [1,0..]; v = [1, x]; parallel_transport(v, mu0, mu) with mu == mu0:
    alpha = 1, coef = <mu, v>_L / 2 = -1/2, u = v + coef*(2 mu0) = [0, x];
    exponential_map(mu, u): r = sqrt(max(<u,u>_L, eps)); cosh(r) mu + sinh(r) u / r.
    """
    r2 = np.sum(x * x, axis=1, keepdims=True)
    r = np.sqrt(np.maximum(r2, EPS))
    return np.concatenate([np.cosh(r), np.sinh(r) * x / r], axis=1)


def _H(sheet):
    X = sheet @ sheet.T
    u = np.sqrt(np.diagonal(X) + 1.0)
    return X - np.outer(u, u)


def get_pdm_torch(x, curvature=-1.0, matsumoto=False, projection="up"):
    """Chyp_torch.get_pdm: clamp -1.0000001, diagonal zeroed."""
    sheet = lift_up(x) if projection == "up" else lift_wrap(x)
    H = np.minimum(_H(sheet), -TORCH_CLAMP)
    D = 1.0 / np.sqrt(-curvature) * np.arccosh(-H)
    if matsumoto:
        D = np.log(np.cosh(D))
    np.fill_diagonal(D, 0.0)
    return D


def get_pdm_np(x, curvature=-1.0, matsumoto=False):
    """Chyp_np.get_pdm: clamp -(1+eps), diagonal NOT zeroed, curvature 0 => Euclidean."""
    if curvature == 0.0:
        D = np.sum((x[:, None] - x) ** 2, axis=-1) ** 0.5
        return np.log(np.cosh(D)) if matsumoto else D
    H = np.minimum(_H(lift_up(x)), -(1.0 + EPS))
    D = 1.0 / np.sqrt(-curvature) * np.arccosh(-H)
    if matsumoto:
        D = np.log(np.cosh(D))
    return D


def get_pdm_torch_vjp(x, grad_D, curvature=-1.0, matsumoto=False):
    """Vector-Jacobian product of get_pdm_torch ("up" lift) with cotangent grad_D (S,S).

    Returns (grad_x (S,D), grad_curvature scalar) -- the values torch autograd yields through
    Chyp_torch.pyx:411-423 (clamped entries and the diagonal pass no gradient)."""
    S = x.shape[0]
    sheet = lift_up(x)
    z = sheet[:, 0]
    Hraw = _H(sheet)
    live = Hraw < -TORCH_CLAMP
    np.fill_diagonal(live, False)
    H = np.minimum(Hraw, -TORCH_CLAMP)
    sc = 1.0 / np.sqrt(-curvature)
    D0 = sc * np.arccosh(-H)
    g = grad_D.copy()
    np.fill_diagonal(g, 0.0)
    if matsumoto:
        g = g * np.tanh(D0)
    # dD0/dcurv = D0 / (2 * (-curv))
    grad_c = np.sum(g * D0) / (2.0 * (-curvature))
    gH = np.where(live, -g * sc / np.sqrt(H * H - 1.0), 0.0)
    gs = gH + gH.T  # H_ij depends on x_i and x_j symmetrically
    # dH_ij/dx_i = x_j - (z_j/z_i) x_i
    grad_x = gs @ x - (gs @ z)[:, None] * x / z[:, None]
    return grad_x, grad_c


def hyperbolic_distance_np(x1, x2, curvature=-1.0):
    """Chyp_np.hyperbolic_distance (Chyp_np.pyx:279-296): points given WITHOUT the z coordinate."""
    s1 = lift_up(x1[None, :])[0]
    s2 = lift_up(x2[None, :])[0]
    if np.isclose(curvature, 0.0):
        return np.linalg.norm(s2 - s1)
    inner = np.maximum(-(-s1[0] * s2[0] + np.dot(s1[1:], s2[1:])), 1.0 + EPS)
    return 1.0 / np.sqrt(-curvature) * np.arccosh(inner)


def poincare_to_hyper(loc):
    """Chyp_torch.poincare_to_hyper (Chyp_torch.pyx:184-210), 1-D or 2-D input."""
    loc = np.asarray(loc, dtype=np.float64)
    a = np.sum(loc * loc, axis=-1, keepdims=True)
    out0 = (1.0 + a) / (1.0 - a)
    out1 = 2.0 * loc / (1.0 - a + EPS)
    return np.concatenate([out0, out1], axis=-1)


def hyperbolic_distance_torch(x1, x2, curvature=-1.0):
    """Chyp_torch.hyperbolic_distance (Chyp_torch.pyx:119-154): Poincare-ball formula at
    curvature -1, Lorentz form otherwise (points in the Poincare ball)."""
    if abs(curvature + 1.0) > 1e-9:
        return hyperbolic_distance_lorentz(x1, x2, curvature)
    inv = 2.0 * np.sum((x2 - x1) ** 2) / (1.0 - np.linalg.norm(x1) ** 2) / (1.0 - np.linalg.norm(x2) ** 2)
    if np.isnan(inv):
        return hyperbolic_distance_lorentz(x1, x2, curvature)
    return np.arccosh(1.0 + inv)


def hyperbolic_distance_lorentz(x1, x2, curvature=-1.0):
    """Chyp_torch.hyperbolic_distance_lorentz (Chyp_torch.pyx:137-154)."""
    if np.isclose(curvature, 0.0):
        return np.linalg.norm(x2 - x1)
    z1 = poincare_to_hyper(x1)
    z2 = poincare_to_hyper(x2)
    inner = np.maximum(-(-z1[0] * z2[0] + np.dot(z1[1:], z2[1:])), 1.0 + EPS)
    return 1.0 / np.sqrt(-curvature) * np.arccosh(inner)

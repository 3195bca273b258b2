"""NumPy-facing hyperboloid utilities of the MCMC path (dodonaphy/cython/Chyp_np.pyx).

``get_pdm`` (Chyp_np.pyx:321-367) runs the CUDA distance kernel (NumPy variant: clamp -(1+eps),
diagonal kept, curvature 0 => Euclidean).  The single-point helpers are O(D) host arithmetic kept
for API parity (project_up :253-277, hyperbolic_distance :279-296, lorentz_product :298-319,
poincare_to_hyper(_2d) :413-439).
"""
import numpy as np
import torch

from . import engine as _engine

eps = np.finfo(np.double).eps


def get_pdm(x, curvature=-1.0, matsumoto=False, get_jacobian=False):
    """(S,D) ndarray -> (S,S) ndarray of pairwise distances; a leading batch axis is accepted."""
    if get_jacobian and matsumoto:
        raise NotImplementedError("Jacobian is not implemented for Matsumoto metric.")
    x = np.ascontiguousarray(x, dtype=np.float64)
    xd = torch.from_numpy(x).cuda() if torch.cuda.is_available() else torch.from_numpy(x)
    out = _engine.pdm_forward(xd, float(curvature), "numpy", "up", bool(matsumoto))
    if not get_jacobian:
        return out.cpu().numpy()
    if curvature == 0.0:
        return out.cpu().numpy(), np.eye(x.shape[0])
    return out.cpu().numpy(), _pair_jacobian(xd, float(curvature)).cpu().numpy()


def _pair_jacobian(x, curvature):
    """Analytic Jacobian of the n(n-1)/2 pair distances (row-major pairs i<j) w.r.t. the n*(D+1) sheet
    coordinates (Chyp_np.pyx:369-410), vectorised on the device the points live on:
    row k=(i,j): block i = A_ij (B_ij s_i - s_j), block j = A_ji (B_ji s_j - s_i),
    A = 1/sqrt(-kappa (H^2-1)), B_ij = u_j/u_i."""
    n, dim = x.shape
    s = torch.cat((torch.sqrt((x * x).sum(1, keepdim=True) + 1.0), x), dim=1)
    X = s @ s.T
    u = torch.sqrt(torch.diagonal(X) + 1.0)
    H = torch.minimum(X - torch.outer(u, u), torch.tensor(-(1.0 + eps), dtype=torch.float64, device=x.device))
    A = 1.0 / torch.sqrt(-curvature * (H * H - 1.0))
    Bm = torch.outer(1.0 / u, u)
    i, j = torch.triu_indices(n, n, 1, device=x.device)
    npair = i.numel()
    G = torch.zeros((npair, n, dim + 1), dtype=torch.float64, device=x.device)
    k = torch.arange(npair, device=x.device)
    G[k, i] = A[i, j, None] * (Bm[i, j, None] * s[i] - s[j])
    G[k, j] = A[j, i, None] * (Bm[j, i, None] * s[j] - s[i])
    return G.reshape(npair, n * (dim + 1))


def project_up(loc):
    loc = np.asarray(loc, dtype=np.float64)
    return np.concatenate(([np.sqrt(np.sum(loc * loc) + 1.0)], loc))


def project_up_2d(loc):
    loc = np.asarray(loc, dtype=np.float64)
    return np.concatenate((np.sqrt(np.sum(loc * loc, axis=1, keepdims=True) + 1.0), loc), axis=1)


def lorentz_product(x, y):
    return -x[0] * y[0] + np.dot(x[1:], y[1:])


def hyperbolic_distance(x1, x2, curvature=-1.0):
    """Distance between two points given WITHOUT their time coordinate (Chyp_np.pyx:279-296); (D,) as in
    the reference or (n, D) for n pairs at once.  Runs in the pairs kernel (csrc/hypdist.cu)."""
    out = _engine.hypdist_forward(np.ascontiguousarray(x1, dtype=np.float64), np.ascontiguousarray(x2, dtype=np.float64),
                                  float(curvature), "sheet_up").cpu().numpy()
    return np.float64(out) if out.ndim == 0 else out


def poincare_to_hyper(location):
    location = np.asarray(location, dtype=np.float64)
    a = np.sum(location * location)
    return np.concatenate(([(1 + a) / (1 - a)], 2 * location / (1 - a + eps)))


def poincare_to_hyper_2d(location):
    location = np.asarray(location, dtype=np.float64)
    a = np.sum(location * location, axis=-1, keepdims=True)
    return np.concatenate(((1 + a) / (1 - a), 2 * location / (1 - a + eps)), axis=1)


# ---------------------------------------------------------------------------------------------
# NumPy chart changes of the MCMC side (Chyp_np.pyx:41-229), row-vectorised; same names.
def _lorentz_rows(x, y):
    return -x[..., 0] * y[..., 0] + np.sum(x[..., 1:] * y[..., 1:], axis=-1)


def parallel_transport(xi, x, y):
    alpha = -_lorentz_rows(x, y)
    coef = _lorentz_rows(y, xi) / (alpha + 1)
    return xi + np.expand_dims(coef, -1) * (x + y)


def exponential_map(x, v):
    vnorm = np.expand_dims(np.sqrt(np.maximum(_lorentz_rows(v, v), eps)), -1)
    return np.cosh(vnorm) * x + np.sinh(vnorm) * v / vnorm


def exp_map_inverse(z, mu):
    alpha = np.expand_dims(-_lorentz_rows(mu, z), -1)
    return np.arccosh(alpha) / np.sqrt(alpha**2 - 1) * (z - alpha * mu)


def unwrap(z, dim=None):
    """Sheet point(s) -> spatial part of the tangent vector at the origin pointing to them."""
    z = np.asarray(z, dtype=np.float64)
    mu0 = np.zeros_like(z)
    mu0[..., 0] = 1.0
    return exp_map_inverse(z, mu0)[..., 1:]


unwrap_2d = unwrap


def tangent_to_hyper(mu, v_tilde, dim=None):
    """Carry v_tilde from T_0 to T_mu and shoot onto the sheet.  Unlike the torch twin, the lifted
    vector's time coordinate is zero here (Chyp_np.pyx:161 vs Chyp_torch.pyx:343)."""
    mu = np.asarray(mu, dtype=np.float64)
    v_tilde = np.asarray(v_tilde, dtype=np.float64).reshape(mu.shape[:-1] + (mu.shape[-1] - 1,))
    one = np.ones(v_tilde.shape[:-1] + (1,))
    mu0 = np.concatenate((one, np.zeros_like(v_tilde)), -1)
    return exponential_map(mu, parallel_transport(np.concatenate((0.0 * one, v_tilde), -1), mu0, mu))


def tangent_to_hyper_jacobian(mu, v_tilde, dim):
    r = np.linalg.norm(v_tilde, axis=-1)
    return (dim - 1) * np.log(np.sinh(r) / r)


def hyper_to_poincare(location):
    location = np.asarray(location, dtype=np.float64)
    return location[..., 1:] / (1 + location[..., :1])


def hyper_to_poincare_jacobian(location):
    location = np.asarray(location, dtype=np.float64)
    n = location.shape[-1]
    a = 1 + location[..., 0]
    return np.log(np.abs((a**2 + np.sum(location[..., 1:] ** 2, axis=-1)) / a ** (2 * (n + 1))))


def project_up_jacobian(loc):
    """d sheet / d loc for all points: block diagonal of [(x/z)^T ; I_D] blocks, (n(D+1), nD)."""
    loc = np.asarray(loc, dtype=np.float64)
    n, D = loc.shape
    col = loc / np.sqrt(np.sum(loc**2, axis=1, keepdims=True) + 1)
    out = np.zeros((n * (D + 1), n * D))
    for i in range(n):
        out[i * (D + 1), i * D:(i + 1) * D] = col[i]
        out[i * (D + 1) + 1:(i + 1) * (D + 1), i * D:(i + 1) * D] = np.eye(D)
    return out

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
    """Distance between two points given WITHOUT their time coordinate (Chyp_np.pyx:279-296)."""
    s1, s2 = project_up(x1), project_up(x2)
    if np.isclose(curvature, 0.0):
        return np.linalg.norm(s2 - s1)
    inner = np.maximum(-lorentz_product(s1, s2), 1.0 + eps)
    return 1.0 / np.sqrt(-curvature) * np.arccosh(inner)


def poincare_to_hyper(location):
    location = np.asarray(location, dtype=np.float64)
    a = np.sum(location * location)
    return np.concatenate(([(1 + a) / (1 - a)], 2 * location / (1 - a + eps)))


def poincare_to_hyper_2d(location):
    location = np.asarray(location, dtype=np.float64)
    a = np.sum(location * location, axis=-1, keepdims=True)
    return np.concatenate(((1 + a) / (1 - a), 2 * location / (1 - a + eps)), axis=1)

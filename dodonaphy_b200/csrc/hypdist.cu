// K1p: batched POINTWISE hyperbolic distances (pairs of points), forward and vector-Jacobian product, sm_100a.
//
// Replaces the reference's single-pair utilities, which its callers invoke in Python loops over the
// branches of a tree (BaseModel.compute_branch_lengths, dodonaphy/base_model.py:127-181; Cphylo
// compute_branch_lengths_np, dodonaphy/cython/Cphylo.pyx:116-156):
//   DODO_DIST_POINCARE   Chyp_torch.hyperbolic_distance        dodonaphy/cython/Chyp_torch.pyx:119-134
//                        Poincare-ball formula acosh(1 + 2|x2-x1|^2 / ((1-|x1|^2)(1-|x2|^2))) at curvature -1
//                        (falls through to the Lorentz form for any other curvature or a NaN invariant)
//   DODO_DIST_LORENTZ    Chyp_torch.hyperbolic_distance_lorentz  :137-154
//                        ball -> sheet (poincare_to_hyper :184-210, incl. its +eps in the denominator),
//                        clamp(-<z1,z2>_L, min 1+eps), acosh / sqrt(-kappa); curvature ~ 0 => |x2-x1|
//   DODO_DIST_SHEET_UP   Chyp_np.hyperbolic_distance            dodonaphy/cython/Chyp_np.pyx:279-296
//                        points lifted by project_up (:253-272), max(-<.,.>_L, 1+eps), arccosh / sqrt(-kappa)
// One thread per pair; D <= 16 coordinates per point.  Same operation order as the reference's torch /
// NumPy expressions (sums of squares are rounded products added left to right; the norm is squared
// again after its square root, as `torch.linalg.norm(x)**2` does), so results agree to the last bits
// of acosh.  The VJP (torch variants: the reference differentiates them with autograd) follows the
// same graph: clamped entries pass no gradient.
#include <float.h>

#include "common.cuh"

namespace dodo {

constexpr int HD_MAX_D = 16;

struct BallSheet {  // z = poincare_to_hyper(x) and what its backward needs
    double z0, zs[HD_MAX_D], a;
};

__device__ __forceinline__ double sumsq_seq(const double *x, int D) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s = __dadd_rn(s, __dmul_rn(x[d], x[d]));
    return s;
}

__device__ __forceinline__ void ball_to_sheet(const double *x, int D, BallSheet &o) {
    o.a = sumsq_seq(x, D);
    o.z0 = (1.0 + o.a) / (1.0 - o.a);
    const double den = (1.0 - o.a) + DBL_EPSILON;
    for (int d = 0; d < D; ++d) o.zs[d] = 2.0 * x[d] / den;
}

__device__ __forceinline__ double lorentz_neg(const double z10, const double *z1s, const double z20, const double *z2s, int D) {
    double dot = 0.0;
    for (int d = 0; d < D; ++d) dot = fma(z1s[d], z2s[d], dot);
    return -(__dadd_rn(-__dmul_rn(z10, z20), dot));  // -( -x0 y0 + x[1:].y[1:] )
}

// forward of one pair; `used` reports which branch produced the value (for the backward)
enum { HD_EUCLID = 0, HD_BALL = 1, HD_LOR = 2, HD_LOR_CLAMPED = 3, HD_UP = 4, HD_UP_CLAMPED = 5 };

__device__ double pair_distance(const double *x1, const double *x2, int D, double kappa, int form, int &used) {
    if (form == DODO_DIST_SHEET_UP) {
        const double n1 = sumsq_seq(x1, D), n2 = sumsq_seq(x2, D);
        const double z1 = sqrt(n1 + 1.0), z2 = sqrt(n2 + 1.0);
        if (fabs(kappa) <= 1e-8) {  // np.isclose(curvature, 0): Euclidean norm of the SHEET difference
            double s = (z2 - z1) * (z2 - z1);
            for (int d = 0; d < D; ++d) s += (x2[d] - x1[d]) * (x2[d] - x1[d]);
            used = HD_EUCLID;
            return sqrt(s);
        }
        double inner = lorentz_neg(z1, x1, z2, x2, D);
        used = inner > 1.0 + DBL_EPSILON ? HD_UP : HD_UP_CLAMPED;
        inner = fmax(inner, 1.0 + DBL_EPSILON);
        return 1.0 / sqrt(-kappa) * acosh(inner);
    }
    if (form == DODO_DIST_POINCARE && !(fabs(kappa + 1.0) > 1e-9)) {
        double ss = 0.0;
        for (int d = 0; d < D; ++d) {
            const double t = x2[d] - x1[d];
            ss = __dadd_rn(ss, __dmul_rn(t, t));
        }
        const double r1 = sqrt(sumsq_seq(x1, D)), r2 = sqrt(sumsq_seq(x2, D));  // torch.linalg.norm(x)**2
        const double inv = 2.0 * ss / (1.0 - r1 * r1) / (1.0 - r2 * r2);
        if (inv == inv) {
            used = HD_BALL;
            return acosh(1.0 + inv);
        }
    }
    if (fabs(kappa) <= 1e-8) {  // torch.isclose(curvature, 0)
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += (x2[d] - x1[d]) * (x2[d] - x1[d]);
        used = HD_EUCLID;
        return sqrt(s);
    }
    BallSheet a, b;
    ball_to_sheet(x1, D, a);
    ball_to_sheet(x2, D, b);
    double inner = lorentz_neg(a.z0, a.zs, b.z0, b.zs, D);
    used = inner > 1.0 + DBL_EPSILON ? HD_LOR : HD_LOR_CLAMPED;
    inner = fmax(inner, 1.0 + DBL_EPSILON);
    return 1.0 / sqrt(-kappa) * acosh(inner);
}

__global__ void hypdist_fwd_kernel(const double *__restrict__ x1, const double *__restrict__ x2, long long n, int D,
                                   const double *__restrict__ curv, int form, int matsumoto, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[HD_MAX_D], b[HD_MAX_D];
    for (int d = 0; d < D; ++d) { a[d] = x1[i * D + d]; b[d] = x2[i * D + d]; }
    int used;
    double dist = pair_distance(a, b, D, curv[0], form, used);
    if (matsumoto) dist = log(cosh(dist));
    out[i] = dist;
}

// g1 = g * d dist / d x1, g2 likewise, gk[i] = g * d dist / d kappa (one partial per pair)
__global__ void hypdist_bwd_kernel(const double *__restrict__ x1, const double *__restrict__ x2, long long n, int D,
                                   const double *__restrict__ curv, int form, int matsumoto,
                                   const double *__restrict__ gout, double *__restrict__ g1, double *__restrict__ g2,
                                   double *__restrict__ gk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[HD_MAX_D], b[HD_MAX_D], ga[HD_MAX_D], gb[HD_MAX_D];
    for (int d = 0; d < D; ++d) { a[d] = x1[i * D + d]; b[d] = x2[i * D + d]; ga[d] = 0.0; gb[d] = 0.0; }
    const double kappa = curv[0];
    int used;
    const double dist = pair_distance(a, b, D, kappa, form, used);
    double g = gout[i];
    if (matsumoto) g *= tanh(dist);
    double gkap = 0.0;
    if (used == HD_EUCLID) {
        if (form == DODO_DIST_SHEET_UP) {
            const double z1 = sqrt(sumsq_seq(a, D) + 1.0), z2 = sqrt(sumsq_seq(b, D) + 1.0);
            if (dist > 0.0)
                for (int d = 0; d < D; ++d) {
                    const double t = (b[d] - a[d]) / dist;
                    ga[d] = g * (-t - (z2 - z1) / dist * a[d] / z1);
                    gb[d] = g * (t + (z2 - z1) / dist * b[d] / z2);
                }
        } else if (dist > 0.0) {
            for (int d = 0; d < D; ++d) {
                const double t = (b[d] - a[d]) / dist;
                ga[d] = -g * t;
                gb[d] = g * t;
            }
        }
    } else if (used == HD_BALL) {
        double ss = 0.0;
        for (int d = 0; d < D; ++d) ss += (b[d] - a[d]) * (b[d] - a[d]);
        const double r1 = sqrt(sumsq_seq(a, D)), r2 = sqrt(sumsq_seq(b, D));
        const double A = 1.0 - r1 * r1, Bq = 1.0 - r2 * r2;
        const double inv = 2.0 * ss / A / Bq;
        const double y = 1.0 + inv;
        const double gi = g / sqrt(y * y - 1.0);  // d acosh(1 + inv) / d inv
        for (int d = 0; d < D; ++d) {
            const double t = b[d] - a[d];
            ga[d] = gi * (-4.0 * t / (A * Bq) + inv * 2.0 * a[d] / A);
            gb[d] = gi * (4.0 * t / (A * Bq) + inv * 2.0 * b[d] / Bq);
        }
    } else if (used == HD_LOR || used == HD_UP) {
        const double sc = 1.0 / sqrt(-kappa);
        gkap = g * dist / (2.0 * (-kappa));
        if (used == HD_UP) {
            const double z1 = sqrt(sumsq_seq(a, D) + 1.0), z2 = sqrt(sumsq_seq(b, D) + 1.0);
            const double inner = lorentz_neg(z1, a, z2, b, D);
            const double gi = g * sc / sqrt(inner * inner - 1.0);  // inner = z1 z2 - a.b
            for (int d = 0; d < D; ++d) {
                ga[d] = gi * (z2 * a[d] / z1 - b[d]);
                gb[d] = gi * (z1 * b[d] / z2 - a[d]);
            }
        } else {
            BallSheet p, q;
            ball_to_sheet(a, D, p);
            ball_to_sheet(b, D, q);
            const double inner = lorentz_neg(p.z0, p.zs, q.z0, q.zs, D);
            const double gi = g * sc / sqrt(inner * inner - 1.0);  // inner = p0 q0 - ps.qs
            // d inner / d p0 = q0, d inner / d ps = -qs ; p0 = (1+a)/(1-a), ps = 2 x / (1 - a + eps), a = |x|^2
            auto back = [&](const BallSheet &me, const BallSheet &other, const double *x, double *gx) {
                const double den = (1.0 - me.a) + DBL_EPSILON;
                double dot = 0.0;  // sum_d (d inner / d ps_d) * x_d
                for (int d = 0; d < D; ++d) dot += -other.zs[d] * x[d];
                // d p0 / d a = 2 / (1-a)^2 ; d ps_d / d a = 2 x_d / den^2
                const double g_a = other.z0 * 2.0 / ((1.0 - me.a) * (1.0 - me.a)) + dot * 2.0 / (den * den);
                for (int d = 0; d < D; ++d) gx[d] = gi * (-other.zs[d] * 2.0 / den + g_a * 2.0 * x[d]);
            };
            back(p, q, a, ga);
            back(q, p, b, gb);
        }
    } else {  // clamped: constant distance in the locations, still proportional to (-kappa)^(-1/2)
        gkap = g * dist / (2.0 * (-kappa));
    }
    for (int d = 0; d < D; ++d) { g1[i * D + d] = ga[d]; g2[i * D + d] = gb[d]; }
    gk[i] = gkap;
}

}  // namespace dodo

using namespace dodo;

static int hypdist_check(const char *who, const void *a, const void *b, const void *c, long long n, int D, int form) {
    DODO_CHECK_ARG(a && b && c, "%s: NULL argument", who);
    DODO_CHECK_ARG(n >= 1 && D >= 1 && D <= HD_MAX_D, "%s: need n >= 1 and 1 <= D <= %d", who, HD_MAX_D);
    DODO_CHECK_ARG(form == DODO_DIST_POINCARE || form == DODO_DIST_LORENTZ || form == DODO_DIST_SHEET_UP, "%s: bad form", who);
    return DODO_OK;
}

extern "C" int dodo_hypdist_fwd(const double *d_x1, const double *d_x2, int64_t n, int D, const double *d_curvature,
                                int form, int matsumoto, double *d_out, void *stream_) {
    int rc = hypdist_check("dodo_hypdist_fwd", d_x1, d_x2, d_out, n, D, form);
    if (rc != DODO_OK) return rc;
    DODO_CHECK_ARG(d_curvature, "dodo_hypdist_fwd: NULL curvature");
    hypdist_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(d_x1, d_x2, n, D, d_curvature, form,
                                                                                       matsumoto, d_out);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_hypdist_bwd(const double *d_x1, const double *d_x2, int64_t n, int D, const double *d_curvature,
                                int form, int matsumoto, const double *d_grad_out, double *d_grad_x1, double *d_grad_x2,
                                double *d_grad_curv, void *stream_) {
    int rc = hypdist_check("dodo_hypdist_bwd", d_x1, d_x2, d_grad_out, n, D, form);
    if (rc != DODO_OK) return rc;
    DODO_CHECK_ARG(d_curvature && d_grad_x1 && d_grad_x2 && d_grad_curv, "dodo_hypdist_bwd: NULL argument");
    hypdist_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(
        d_x1, d_x2, n, D, d_curvature, form, matsumoto, d_grad_out, d_grad_x1, d_grad_x2, d_grad_curv);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

// K1: batched pairwise hyperbolic distance matrix and its vector-Jacobian product, sm_100a.
//
// Replaces Chyp_torch.get_pdm (dodonaphy/cython/Chyp_torch.pyx:399-423) with the "up" lift
// project_up_2d (:486-504) and the "wrap" lift tangent_to_hyper_2d (:387-396), and
// Chyp_np.get_pdm (dodonaphy/cython/Chyp_np.pyx:321-367).  The arithmetic follows the reference's
// graph literally: sheet s_i, X = s s^T, u~_i = sqrt(X_ii + 1), H = X - u~ u~^T, clamp,
// D = acosh(-H)/sqrt(-kappa), optional log cosh, diagonal.
//
// Tiny next to the pruning kernel (8 S^2 bytes written per tree); one thread per matrix entry,
// rows of the sheet staged in shared memory.
#include <float.h>

#include "common.cuh"

namespace dodo {

constexpr int MAX_D = 15;  // embedding dimension (sheet has D+1 coordinates)
constexpr double TORCH_CLAMP = 1.0000001;

// sum_d x_d^2 exactly as the reference's host libraries round it: squares rounded one by one
// (np.power / torch.pow), then NumPy's pairwise_sum of a contiguous row -- a plain left-to-right sum
// below 8 terms, eight running partial sums combined as ((0+1)+(2+3))+((4+5)+(6+7)) plus a sequential
// tail from 8 to 15 terms (torch's row reduction is the plain sum at these lengths).  No FMA
// contraction: the last joins of hard NJ are exact mathematical ties of Q and follow the last bit.
__device__ __forceinline__ double row_norm2(const double *__restrict__ x, int D) {
    if (D < 8) {
        double n2 = 0.0;
        for (int d = 0; d < D; ++d) n2 = __dadd_rn(n2, __dmul_rn(x[d], x[d]));
        return n2;
    }
    double r[8];
    for (int d = 0; d < 8; ++d) r[d] = __dmul_rn(x[d], x[d]);
    double n2 = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                          __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (int d = 8; d < D; ++d) n2 = __dadd_rn(n2, __dmul_rn(x[d], x[d]));
    return n2;
}

__device__ __forceinline__ void lift_point(const double *__restrict__ x, int D, int lift, double *s) {
    const double n2 = row_norm2(x, D);
    if (lift == DODO_LIFT_UP) {
        s[0] = sqrt(__dadd_rn(n2, 1.0));
        for (int d = 0; d < D; ++d) s[d + 1] = x[d];
    } else {
        double r = sqrt(fmax(n2, DBL_EPSILON));
        double sh = sinh(r);
        s[0] = cosh(r);
        for (int d = 0; d < D; ++d) s[d + 1] = sh * x[d] / r;
    }
}

// sheet[b][i][0..D], util[b][i] = sqrt(<s_i,s_i>_E + 1)
__global__ void lift_kernel(const double *__restrict__ x, int n, int D, int lift, double *__restrict__ sheet,
                            double *__restrict__ util) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    double s[MAX_D + 1];
    lift_point(x + (size_t)gid * D, D, lift, s);
    double xx = 0.0;
    for (int d = 0; d <= D; ++d) {
        sheet[(size_t)gid * (D + 1) + d] = s[d];
        xx = fma(s[d], s[d], xx);  // diagonal of X: the same chain as pair_H
    }
    util[gid] = sqrt(__dadd_rn(xx, 1.0));
}

// H_ij = X_ij - u~_i u~_j with X = s s^T.  X is the k-ordered FMA chain a BLAS dgemm/dsyrk micro-kernel
// runs for an inner dimension this small (one accumulator per entry, k ascending, alpha = 1); the outer
// product and the subtraction are separate roundings in the reference (np.outer / torch.outer, then `-`).
__device__ __forceinline__ double pair_H(const double *si, const double *sj, int D, double ui, double uj) {
    double X = 0.0;
    for (int d = 0; d <= D; ++d) X = fma(si[d], sj[d], X);
    return __dsub_rn(X, __dmul_rn(ui, uj));
}

__global__ void pdm_fwd_kernel(const double *__restrict__ x, const double *__restrict__ sheet,
                               const double *__restrict__ util, int S, int D, const double *__restrict__ curv,
                               int variant, int matsumoto, double *__restrict__ pdm) {
    const int b = blockIdx.z;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S || j >= S) return;
    const double kappa = curv[0];
    double out;
    if (variant == DODO_PDM_NUMPY && kappa == 0.0) {
        const double *xi = x + ((size_t)b * S + i) * D, *xj = x + ((size_t)b * S + j) * D;
        double acc = 0.0;
        for (int d = 0; d < D; ++d) {
            double t = xi[d] - xj[d];
            acc += t * t;
        }
        out = sqrt(acc);
    } else {
        const double *si = sheet + ((size_t)b * S + i) * (D + 1), *sj = sheet + ((size_t)b * S + j) * (D + 1);
        double H = pair_H(si, sj, D, util[(size_t)b * S + i], util[(size_t)b * S + j]);
        const double cl = (variant == DODO_PDM_TORCH) ? -TORCH_CLAMP : -(1.0 + DBL_EPSILON);
        H = fmin(H, cl);
        out = 1.0 / sqrt(-kappa) * acosh(-H);
    }
    if (matsumoto) out = log(cosh(out));
    if (variant == DODO_PDM_TORCH && i == j) out = 0.0;
    pdm[((size_t)b * S + i) * S + j] = out;
}

// One CTA per sample; warp per row i, lanes over j; fixed-order reductions.
__global__ void __launch_bounds__(1024) pdm_bwd_kernel(const double *__restrict__ x, const double *__restrict__ sheet,
                                                      const double *__restrict__ util, int S, int D,
                                                      const double *__restrict__ curv, int lift, int matsumoto,
                                                      const double *__restrict__ gpdm, double *__restrict__ gx,
                                                      double *__restrict__ gcurv) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double kappa = curv[0];
    const double sc = 1.0 / sqrt(-kappa);
    const double *sh = sheet + (size_t)b * S * (D + 1);
    const double *ut = util + (size_t)b * S;
    const double *G = gpdm + (size_t)b * S * S;
    __shared__ double cpart[32];
    double csum = 0.0;
    for (int i = warp; i < S; i += nwarp) {
        double si[MAX_D + 1], acc[MAX_D + 1];
        for (int d = 0; d <= D; ++d) { si[d] = sh[(size_t)i * (D + 1) + d]; acc[d] = 0.0; }
        const double ui = ut[i];
        // pairs without a cotangent are skipped a warp at a time (the one-hot cotangents of the batched
        // Jacobian leave most of the matrix zero), rows without any skip their reductions as well
        bool row_any = false;
        for (int jb = 0; jb < S; jb += 32) {
            const int j = jb + lane;
            const bool use = j < S && j != i;
            double gij = 0.0, gji = 0.0;
            if (use) { gij = G[(size_t)i * S + j]; gji = G[(size_t)j * S + i]; }
            const bool nz = use && (gij != 0.0 || gji != 0.0);
            if (!__any_sync(0xffffffffu, nz)) continue;
            row_any = true;
            if (!nz) continue;
            const double *sj = sh + (size_t)j * (D + 1);
            const double uj = ut[j];
            double Hraw = pair_H(si, sj, D, ui, uj);
            bool live = Hraw < -TORCH_CLAMP;
            double H = fmin(Hraw, -TORCH_CLAMP);
            double D0 = sc * acosh(-H);
            if (matsumoto) { double th = tanh(D0); gij *= th; gji *= th; }
            csum += gij * D0;
            if (live) {
                double gs = -(gij + gji) * sc / sqrt(H * H - 1.0);
                double ratio = uj / ui;
                for (int d = 0; d <= D; ++d) acc[d] += gs * (sj[d] - ratio * si[d]);
            }
        }
        if (!row_any) {  // warp-uniform
            if (lane == 0) {
                double *out = gx + ((size_t)b * S + i) * D;
                for (int d = 0; d < D; ++d) out[d] = 0.0;
            }
            continue;
        }
        for (int d = 0; d <= D; ++d) acc[d] = warp_sum(acc[d]);
        if (lane == 0) {
            const double *xi = x + ((size_t)b * S + i) * D;
            double *out = gx + ((size_t)b * S + i) * D;
            if (lift == DODO_LIFT_UP) {
                for (int d = 0; d < D; ++d) out[d] = acc[d + 1] + acc[0] * xi[d] / si[0];
            } else {
                const double n2 = row_norm2(xi, D);
                double r = sqrt(fmax(n2, DBL_EPSILON));
                double shr = sinh(r), chr = cosh(r);
                double f = shr / r;
                // s_0 = cosh r, s_d = f x_d ; dr/dx = x/r when |x|^2 > eps else 0
                double dot = 0.0;
                for (int d = 0; d < D; ++d) dot = fma(acc[d + 1], xi[d], dot);
                double dfdr = (r * chr - shr) / (r * r);
                double rad = (n2 > DBL_EPSILON) ? (acc[0] * shr + dot * dfdr) / r : 0.0;
                for (int d = 0; d < D; ++d) out[d] = f * acc[d + 1] + rad * xi[d];
            }
        }
    }
    csum = warp_sum(csum);
    if (lane == 0) cpart[warp] = csum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += cpart[w];
        gcurv[b] = s / (2.0 * (-kappa));
    }
}

}  // namespace dodo

using namespace dodo;

extern "C" uint64_t dodo_pdm_workspace_bytes(int B, int S, int D) {
    return ((uint64_t)B * S * (D + 2)) * sizeof(double);
}

// workspace layout: sheet [B][S][D+1], then util [B][S]
static int run_lift(const double *d_x, int B, int S, int D, int lift, void *d_ws, cudaStream_t stream, double **sheet,
                    double **util) {
    size_t n = (size_t)B * S;
    *sheet = (double *)d_ws;
    *util = *sheet + n * (D + 1);
    lift_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_x, (int)n, D, lift, *sheet, *util);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_pdm_fwd(const double *d_x, int B, int S, int D, const double *d_curvature, int variant, int lift,
                            int matsumoto, void *d_ws, double *d_pdm, void *stream_) {
    DODO_CHECK_ARG(d_x && d_curvature && d_pdm && d_ws, "dodo_pdm_fwd: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 1 && D >= 1 && D <= MAX_D, "dodo_pdm_fwd: need B>=1, S>=1, 1<=D<=%d", MAX_D);
    DODO_CHECK_ARG(variant == DODO_PDM_TORCH || variant == DODO_PDM_NUMPY, "dodo_pdm_fwd: bad variant");
    DODO_CHECK_ARG(lift == DODO_LIFT_UP || lift == DODO_LIFT_WRAP, "dodo_pdm_fwd: bad lift");
    if (variant == DODO_PDM_NUMPY && lift != DODO_LIFT_UP) {
        set_error("dodo_pdm_fwd: the NumPy variant only has the 'up' lift (Chyp_np.pyx:345)");
        return DODO_ENOTIMPL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    double *sheet, *util;
    int rc = run_lift(d_x, B, S, D, lift, d_ws, stream, &sheet, &util);
    if (rc != DODO_OK) return rc;
    dim3 blk(32, 8), grd((S + 31) / 32, (S + 7) / 8, B);
    pdm_fwd_kernel<<<grd, blk, 0, stream>>>(d_x, sheet, util, S, D, d_curvature, variant, matsumoto, d_pdm);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_pdm_bwd(const double *d_x, int B, int S, int D, const double *d_curvature, int variant, int lift,
                            int matsumoto, void *d_ws, const double *d_grad_pdm, double *d_grad_x,
                            double *d_grad_curv, void *stream_) {
    DODO_CHECK_ARG(d_x && d_curvature && d_grad_pdm && d_grad_x && d_grad_curv && d_ws, "dodo_pdm_bwd: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 1 && D >= 1 && D <= MAX_D, "dodo_pdm_bwd: need B>=1, S>=1, 1<=D<=%d", MAX_D);
    if (variant != DODO_PDM_TORCH) {
        set_error("dodo_pdm_bwd: only the torch variant is differentiable in the reference");
        return DODO_ENOTIMPL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    double *sheet, *util;
    int rc = run_lift(d_x, B, S, D, lift, d_ws, stream, &sheet, &util);
    if (rc != DODO_OK) return rc;
    // one warp per row of the tree's matrix: up to 32 rows in flight per CTA
    // ... unless the batch alone fills the machine (the batched Jacobian: thousands of small trees)
    const int bwd_threads = B >= 8 * sm_count() ? 256 : (S >= 48 ? 1024 : (S >= 16 ? 512 : 256));
    pdm_bwd_kernel<<<B, bwd_threads, 0, stream>>>(d_x, sheet, util, S, D, d_curvature, lift, matsumoto, d_grad_pdm, d_grad_x,
                                          d_grad_curv);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

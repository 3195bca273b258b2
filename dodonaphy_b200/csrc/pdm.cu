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
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------
// d blens / d locations of a whole soft-NJ tree in ONE pass (the reference takes 2S-2
// autograd passes per sample through its get_pdm -> nj_torch chain, vi.py:376-387).
//
// With the selections fixed (they carry no gradient, soft.py:56-61) every distance NJ ever forms is a bilinear
// form of the tip distance matrix D:  d(u, x) = a_u^T D a_x - c_u - c_x  with a_u = (a_l + a_r) / 2 the
// averaging weights of the cluster (a_tip = unit vector) and c_u = a_l^T D a_r / 2 (c_tip = 0): induction on
// d(u, x) = (d(l, x) + d(r, x) - d(l, r)) / 2 (get_new_dist_soft, peeler.py:251-265).  The two branch lengths of a
// join (peeler.py:199-213) are
//   d_pl = d_lr / 2 + (r_l - r_r) / (2 (n - 2)),   d_pr = d_lr - d_pl,
//   d_lr = a_l^T D a_r - c_l - c_r,   r_l - r_r = (a_l - a_r)^T D sigma - (n - 2)(c_l - c_r),
// sigma = the sum of a_x over the other live nodes.  The gradient of a^T D b w.r.t. the sheet point s_i is
// a_i Y_b[i] + b_i Y_a[i] with Y_v[i] = sum_j v_j dD_ij/ds_i, and Y is linear in v: Y_u = (Y_l + Y_r) / 2.  So
// every row of the Jacobian follows from per-node quantities (a, Y, grad c) that are ELEMENTWISE in
// (taxon i, sheet coordinate d): one thread owns one (i, d), walks the S-1 joins on its own (no barrier, no
// reduction) and writes its column of all 2S-2 rows -- O(S^2 D) per tree instead of (2S-2) replays of an
// O(S^2) backward.  A second pass maps the sheet-space rows through the lift to the location coordinates,
// exactly as pdm_bwd_kernel does.  Clamped branches (flags bit0/bit1) pass no gradient, as in nj_soft_bwd.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t jac_ws_doubles(int S, int D) {
    // per tree: a, Y, grad c of the S-1 internal nodes and the 2S-2 sheet-space rows, S (D+1) elements each
    return (size_t)(3 * (S - 1) + 2 * S - 2) * S * (D + 1);
}

// second pass of the Jacobian kernels: the sheet-space rows Js [2S-2][S (D+1)] of tree b through the lift to the
// location coordinates, one thread per (row, taxon), exactly as pdm_bwd_kernel does
__device__ __forceinline__ void jac_rows_through_lift(const double *__restrict__ x, const double *__restrict__ sh,
                                                      const double *__restrict__ Js, double *__restrict__ J, int b, int S, int D,
                                                      int lift, int tid, int T) {
    const int D1 = D + 1, NE = S * D1, E = 2 * S - 2;
    double *Jb = J + (size_t)b * E * S * D;
    for (int w = tid; w < E * S; w += T) {
        const int e = w / S, i = w - e * S;
        const double *acc = Js + (size_t)e * NE + (size_t)i * D1;
        const double *xi = x + ((size_t)b * S + i) * D;
        double *out = Jb + (size_t)e * S * D + (size_t)i * D;
        if (lift == DODO_LIFT_UP) {
            const double s0 = sh[(size_t)i * D1];
            for (int d = 0; d < D; ++d) out[d] = acc[d + 1] + acc[0] * xi[d] / s0;
        } else {
            const double n2 = row_norm2(xi, D);
            const double r = sqrt(fmax(n2, DBL_EPSILON));
            const double shr = sinh(r), chr = cosh(r);
            const double fr = shr / r;
            double dot = 0.0;
            for (int d = 0; d < D; ++d) dot = fma(acc[d + 1], xi[d], dot);
            const double dfdr = (r * chr - shr) / (r * r);
            const double rad = (n2 > DBL_EPSILON) ? (acc[0] * shr + dot * dfdr) / r : 0.0;
            for (int d = 0; d < D; ++d) out[d] = fr * acc[d + 1] + rad * xi[d];
        }
    }
}


__global__ void __launch_bounds__(1024) blens_jac_kernel(const double *__restrict__ x, const double *__restrict__ sheet,
                                                        const double *__restrict__ util, int S, int D,
                                                        const double *__restrict__ curv, int lift, int matsumoto,
                                                        const int64_t *__restrict__ peel, const int32_t *__restrict__ flags,
                                                        double *__restrict__ ws, double *__restrict__ J) {
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int D1 = D + 1, NE = S * D1;
    const double kappa = curv[0];
    const double sc = 1.0 / sqrt(-kappa);
    const double *sh = sheet + (size_t)b * S * D1;
    const double *ut = util + (size_t)b * S;
    const int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    const int32_t *fl = flags + (size_t)b * (S - 1);
    double *An = ws + (size_t)b * jac_ws_doubles(S, D);  // [S-1][NE] averaging weight a_u[i] (per element)
    double *Yn = An + (size_t)(S - 1) * NE;               // [S-1][NE] Y_u[i][d]
    double *Gn = Yn + (size_t)(S - 1) * NE;               // [S-1][NE] d c_u / d s_i[d]
    double *Js = Gn + (size_t)(S - 1) * NE;               // [2S-2][NE] rows in sheet coordinates
    for (int el = tid; el < NE; el += T) {
        const int i = el / D1, dp = el - i * D1;
        double si[MAX_D + 1];
        for (int d = 0; d <= D; ++d) si[d] = sh[(size_t)i * D1 + d];
        const double ui = ut[i];
        auto tipY = [&](int j) -> double {  // dD_ij / ds_i[dp]  (pdm_bwd_kernel's per-pair term)
            if (j == i) return 0.0;
            const double *sj = sh + (size_t)j * D1;
            const double uj = ut[j];
            const double Hraw = pair_H(si, sj, D, ui, uj);
            if (!(Hraw < -TORCH_CLAMP)) return 0.0;  // clamped entry: constant
            double coef = -sc / sqrt(Hraw * Hraw - 1.0);
            if (matsumoto) coef *= tanh(sc * acosh(-Hraw));
            return coef * (sj[dp] - (uj / ui) * si[dp]);
        };
        double Ta = 1.0, TY = 0.0;  // sums over the live nodes: a_x[i] (only tip i counts at the start), Y_x[i][dp]
        for (int j = 0; j < S; ++j) TY += tipY(j);
        for (int t = 0; t < S - 1; ++t) {
            const int l = (int)pl[t * 3 + 0], r = (int)pl[t * 3 + 1], u = (int)pl[t * 3 + 2];
            const int f = fl[t];
            const int n = S - t;
            double al, Yl, Gl, ar, Yr, Gr;
            if (l < S) { al = l == i ? 1.0 : 0.0; Yl = tipY(l); Gl = 0.0; }
            else { const size_t o = (size_t)(l - S) * NE + el; al = An[o]; Yl = Yn[o]; Gl = Gn[o]; }
            if (r < S) { ar = r == i ? 1.0 : 0.0; Yr = tipY(r); Gr = 0.0; }
            else { const size_t o = (size_t)(r - S) * NE + el; ar = An[o]; Yr = Yn[o]; Gr = Gn[o]; }
            const double A = al * Yr + ar * Yl;  // d (a_l^T D a_r) / d s_i[dp]
            const double dlr = A - Gl - Gr;
            const double sig = Ta - al - ar, Ysig = TY - Yl - Yr;
            // (the last join has no other live node: its two rows are the same half of d d_lr bit for bit, as in
            // the reference, where both come from one d_lr / 2)
            const double drr = n == 2 ? 0.0 : (al - ar) * Ysig + sig * (Yl - Yr) - (double)(n - 2) * (Gl - Gr);
            const double nact = fmax((double)n, 3.0);
            const double basel = (f & 1) ? 0.0 : 0.5 * dlr + drr / (2.0 * (nact - 2.0));
            Js[(size_t)l * NE + el] = basel;
            Js[(size_t)r * NE + el] = (f & 2) ? 0.0 : dlr - basel;
            const double au = 0.5 * (al + ar), Yu = 0.5 * (Yl + Yr);
            const size_t o = (size_t)(u - S) * NE + el;
            An[o] = au; Yn[o] = Yu; Gn[o] = 0.5 * A;
            Ta += au - al - ar;
            TY += Yu - Yl - Yr;
        }
    }
    __syncthreads();  // rows complete (written by this CTA): now through the lift, one thread per (row, taxon)
    jac_rows_through_lift(x, sh, Js, J, b, S, D, lift, tid, T);
}

// The same walk for trees whose S^2 pair terms fit in shared memory (up to 118 taxa), restructured around what the
// first kernel spends its time on (ncu: 39k dependent instructions per thread at 64 taxa, 0.12 instructions per cycle
// and warp -- a serial chain of fp64 and L2 latencies, 170 us for 128 trees):
//   * the expensive part of a tip term -- the pair's Lorentz product, sqrt, division (tanh / acosh with the Matsumoto
//     adjustment) -- does not depend on the coordinate and is needed twice per (i, j) (initial sum, and when tip j is
//     joined): all S^2 of them are formed first, by all threads, as independent evaluations that pipeline, and kept
//     in shared memory ([j][i]: conflict-free for the walk) as coef_ij (0 = "no gradient") and u_j / u_i;
//   * the peel is staged in shared memory (no dependent global load per join);
//   * the per-node quantities of the NEXT join's children are requested before the current join is evaluated (they
//     come back from L2: a thread reads what it wrote itself, and stores do not allocate in L1) unless the child is
//     the node being formed, which is taken from registers;
//   * the rows go through the lift in a kernel of their own with one thread per (tree, row, taxon).
// Same expressions in the same order per (taxon, coordinate): rows bit-identical to blens_jac_kernel's.
//   * FUSE_UP ("up" lift): the D+1 coordinates of a taxon sit in neighbouring lanes of ONE warp (32 / (D+1) taxa per
//     warp), so a row goes through the lift where it is formed -- d blens / d x_i[d] = row[d+1] + row[0] x_i[d] / s_i[0],
//     row[0] by one shuffle -- and is written straight into J: no sheet-space rows in memory, no second kernel
//     (29 us for 90 MB of rows in and out at 128 x 64 taxa).  x_i[d] / s_i[0] is formed once per thread, so an entry
//     differs from the two-kernel path in the last bit.
template <bool FUSE_UP>
__global__ void __launch_bounds__(512) blens_jac_cached_kernel(const double *__restrict__ x, const double *__restrict__ sheet,
                                                               const double *__restrict__ util,
                                                               int S, int D, const double *__restrict__ curv, int matsumoto,
                                                               const int64_t *__restrict__ peel, const int32_t *__restrict__ flags,
                                                               double *__restrict__ ws, double *__restrict__ J) {
    extern __shared__ __align__(16) double jc_smd[];  // coefT [S][S] | qT [S][S] | {l, r, u, flags} of the S-1 joins
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int D1 = D + 1, NE = S * D1;
    const double kappa = curv[0];
    const double sc = 1.0 / sqrt(-kappa);
    const double *sh = sheet + (size_t)b * S * D1;
    const double *ut = util + (size_t)b * S;
    const int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    const int32_t *fl = flags + (size_t)b * (S - 1);
    double *An = ws + (size_t)b * jac_ws_doubles(S, D);
    double *Yn = An + (size_t)(S - 1) * NE;
    double *Gn = Yn + (size_t)(S - 1) * NE;
    double *Js = Gn + (size_t)(S - 1) * NE;
    double *coefT = jc_smd, *qT = jc_smd + (size_t)S * S;
    int *pj = reinterpret_cast<int *>(qT + (size_t)S * S);
    for (int t = tid; t < S - 1; t += T) {
        pj[4 * t + 0] = (int)pl[t * 3 + 0];
        pj[4 * t + 1] = (int)pl[t * 3 + 1];
        pj[4 * t + 2] = (int)pl[t * 3 + 2];
        pj[4 * t + 3] = fl[t];
    }
    for (int p = tid; p < S * S; p += T) {
        const int j = p / S, i = p - j * S;
        const double ui = ut[i], uj = ut[j];
        double coef = 0.0;
        if (j != i) {
            const double Hraw = pair_H(sh + (size_t)i * D1, sh + (size_t)j * D1, D, ui, uj);
            if (Hraw < -TORCH_CLAMP) {  // (clamped entries are constants)
                coef = -sc / sqrt(Hraw * Hraw - 1.0);
                if (matsumoto) coef *= tanh(sc * acosh(-Hraw));
            }
        }
        coefT[p] = coef;
        qT[p] = uj / ui;
    }
    __syncthreads();
    // element -> thread: el = tid (+ k T) in the two-kernel form; with the fused lift warp w, lane g (D+1) + dp owns
    // taxon (w + k nwarps) (32 / (D+1)) + g, coordinate dp (the last 32 mod (D+1) lanes of a warp walk along idle)
    const int lane = tid & 31, G = 32 / D1, nwarps = T >> 5;
    const int n_rounds = FUSE_UP ? (S + G * nwarps - 1) / (G * nwarps) : (NE + T - 1) / T;
    for (int k = 0; k < n_rounds; ++k) {
        int i, dp;
        bool live;
        if (FUSE_UP) {
            const int g = lane / D1;
            dp = lane - g * D1;
            i = ((tid >> 5) + k * nwarps) * G + g;
            live = g < G && i < S;
            if (!live) { i = 0; dp = 0; }  // (reads stay in range; nothing is written)
        } else {
            const int e0 = tid + k * T;
            live = e0 < NE;
            if (!live) break;
            i = e0 / D1;
            dp = e0 - i * D1;
        }
        const int el = i * D1 + dp;
        const double sidp = sh[(size_t)i * D1 + dp];
        // fused lift: column (i, dp - 1) of J; r0 = x_i[dp - 1] / s_i[0]
        const double r0 = (FUSE_UP && dp >= 1) ? x[((size_t)b * S + i) * D + (dp - 1)] / sh[(size_t)i * D1] : 0.0;
        double *Jcol = FUSE_UP ? J + (size_t)b * (2 * S - 2) * S * D + (size_t)i * D + (dp - 1) : nullptr;
        const int lane0 = lane - dp;  // the lane that holds coordinate 0 of this taxon
        auto tipY = [&](int j) -> double {  // dD_ij / ds_i[dp]  (pdm_bwd_kernel's per-pair term)
            const double coef = coefT[(size_t)j * S + i];
            return coef != 0.0 ? coef * (sh[(size_t)j * D1 + dp] - qT[(size_t)j * S + i] * sidp) : 0.0;
        };
        auto fetch = [&](int node, double &a, double &Y, double &G) {
            if (node < S) { a = node == i ? 1.0 : 0.0; Y = tipY(node); G = 0.0; }
            else { const size_t o = (size_t)(node - S) * NE + el; a = An[o]; Y = Yn[o]; G = Gn[o]; }
        };
        double Ta = 1.0, TY = 0.0;  // sums over the live nodes: a_x[i] (only tip i counts at the start), Y_x[i][dp]
        for (int j = 0; j < S; ++j) TY += tipY(j);
        int4 o4 = *reinterpret_cast<const int4 *>(pj);
        double al, Yl, Gl, ar, Yr, Gr;
        fetch(o4.x, al, Yl, Gl);
        fetch(o4.y, ar, Yr, Gr);
        for (int t = 0; t < S - 1; ++t) {
            const int l = o4.x, r = o4.y, u = o4.z, f = o4.w;
            const int n = S - t;
            // the next join's children, requested now (the node formed below is handed over in registers)
            int4 on = make_int4(0, 0, 0, 0);
            double aln = 0.0, Yln = 0.0, Gln = 0.0, arn = 0.0, Yrn = 0.0, Grn = 0.0;
            if (t + 1 < S - 1) {
                on = *reinterpret_cast<const int4 *>(pj + 4 * (t + 1));
                if (on.x != u) fetch(on.x, aln, Yln, Gln);
                if (on.y != u) fetch(on.y, arn, Yrn, Grn);
            }
            const double A = al * Yr + ar * Yl;  // d (a_l^T D a_r) / d s_i[dp]
            const double dlr = A - Gl - Gr;
            const double sig = Ta - al - ar, Ysig = TY - Yl - Yr;
            // (the last join has no other live node: its two rows are the same half of d d_lr bit for bit, as in
            // the reference, where both come from one d_lr / 2)
            const double drr = n == 2 ? 0.0 : (al - ar) * Ysig + sig * (Yl - Yr) - (double)(n - 2) * (Gl - Gr);
            const double nact = fmax((double)n, 3.0);
            const double basel = (f & 1) ? 0.0 : 0.5 * dlr + drr / (2.0 * (nact - 2.0));
            const double baser = (f & 2) ? 0.0 : dlr - basel;
            if (FUSE_UP) {
                const double l0 = __shfl_sync(0xffffffffu, basel, lane0), r0v = __shfl_sync(0xffffffffu, baser, lane0);
                if (live && dp >= 1) {
                    Jcol[(size_t)l * S * D] = fma(l0, r0, basel);
                    Jcol[(size_t)r * S * D] = fma(r0v, r0, baser);
                }
            } else {
                Js[(size_t)l * NE + el] = basel;
                Js[(size_t)r * NE + el] = baser;
            }
            const double au = 0.5 * (al + ar), Yu = 0.5 * (Yl + Yr), Gu = 0.5 * A;
            const size_t o = (size_t)(u - S) * NE + el;
            if (live) { An[o] = au; Yn[o] = Yu; Gn[o] = Gu; }
            Ta += au - al - ar;
            TY += Yu - Yl - Yr;
            if (on.x == u) { aln = au; Yln = Yu; Gln = Gu; }
            if (on.y == u) { arn = au; Yrn = Yu; Grn = Gu; }
            al = aln; Yl = Yln; Gl = Gln; ar = arn; Yr = Yrn; Gr = Grn;
            o4 = on;
        }
    }
}

// the rows of every tree through the lift: one thread per (tree, row, taxon)
__global__ void __launch_bounds__(256) jac_lift_kernel(const double *__restrict__ x, const double *__restrict__ sheet,
                                                       const double *__restrict__ ws, int B, int S, int D, int lift,
                                                       double *__restrict__ J) {
    const int E = 2 * S - 2, D1 = D + 1, NE = S * D1;
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (size_t)B * E * S) return;
    const int b = (int)(w / ((size_t)E * S));
    const int rem = (int)(w - (size_t)b * E * S);
    const int e = rem / S, i = rem - e * S;
    const double *Js = ws + (size_t)b * jac_ws_doubles(S, D) + (size_t)3 * (S - 1) * NE;
    const double *acc = Js + (size_t)e * NE + (size_t)i * D1;
    const double *xi = x + ((size_t)b * S + i) * D;
    double *out = J + ((size_t)b * E + e) * S * D + (size_t)i * D;
    if (lift == DODO_LIFT_UP) {
        const double s0 = sheet[((size_t)b * S + i) * D1];
        for (int d = 0; d < D; ++d) out[d] = acc[d + 1] + acc[0] * xi[d] / s0;
    } else {
        const double n2 = row_norm2(xi, D);
        const double r = sqrt(fmax(n2, DBL_EPSILON));
        const double shr = sinh(r), chr = cosh(r);
        const double fr = shr / r;
        double dot = 0.0;
        for (int d = 0; d < D; ++d) dot = fma(acc[d + 1], xi[d], dot);
        const double dfdr = (r * chr - shr) / (r * r);
        const double rad = (n2 > DBL_EPSILON) ? (acc[0] * shr + dot * dfdr) / r : 0.0;
        for (int d = 0; d < D; ++d) out[d] = fr * acc[d + 1] + rad * xi[d];
    }
}

// log sqrt|det(J J^T)| of the Jacobian of every sample (vi.py:385-387), one CTA per sample: the Gram matrix is
// formed in shared memory and reduced by Gaussian elimination (one barrier per step).  J J^T is singular for
// an NJ tree (the last join gives two equal rows), so -- as in the reference -- the value is -inf or rounding
// noise; what matters is that it costs microseconds instead of a batched LAPACK call.
__global__ void __launch_bounds__(1024) gram_logdet_kernel(const double *__restrict__ J, int E, int C,
                                                          double *__restrict__ out) {
    extern __shared__ __align__(16) double gsm[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int ld = E + 1;
    double *A = gsm;  // [E][ld]
    const double *Jb = J + (size_t)b * E * C;
    // Gram matrix: J is staged 32 columns at a time ([E][33], coalesced loads, conflict-free reads); one thread per
    // entry e <= f adds its partial dot product, columns in ascending order
    double *Jc = A + (size_t)E * ld;  // [E][33]
    const int T = blockDim.x;
    for (int idx = tid; idx < E * ld; idx += T) A[idx] = 0.0;
    for (int c0 = 0; c0 < C; c0 += 32) {
        __syncthreads();
        for (int idx = tid; idx < E * 32; idx += T) {
            const int e = idx >> 5, c = idx & 31;
            Jc[e * 33 + c] = c0 + c < C ? Jb[(size_t)e * C + c0 + c] : 0.0;
        }
        __syncthreads();
        for (int idx = tid; idx < E * E; idx += T) {
            const int e = idx / E, f = idx - e * E;
            if (f < e) continue;
            const double *re = Jc + e * 33, *rf = Jc + f * 33;
            double acc = A[e * ld + f];
#pragma unroll 8
            for (int c = 0; c < 32; ++c) acc = fma(re[c], rf[c], acc);
            A[e * ld + f] = acc;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < E * E; idx += T) {
        const int e = idx / E, f = idx - e * E;
        if (f > e) A[f * ld + e] = A[e * ld + f];
    }
    // J J^T is symmetric positive semi-definite: elimination without pivoting (the Cholesky order) is stable.  One
    // barrier per step; warps take the rows below the pivot, lanes the columns to its right.
    double logabs = 0.0;
    for (int k = 0; k < E; ++k) {
        __syncthreads();
        const double piv = A[k * ld + k];
        if (!(piv != 0.0)) {  // exactly singular (or NaN): log|det| = -inf (NaN propagates)
            if (tid == 0) out[b] = piv == piv ? -INFINITY : piv;
            return;
        }
        if (tid == 0) logabs += log(fabs(piv));
        const double inv = 1.0 / piv;
        const double *rowk = A + k * ld;
        for (int i = k + 1 + warp; i < E; i += nwarp) {
            double *rowi = A + i * ld;
            const double l = rowi[k] * inv;
            for (int j = k + 1 + lane; j < E; j += 32) rowi[j] = fma(-l, rowk[j], rowi[j]);
        }
    }
    if (tid == 0) out[b] = 0.5 * logabs;
}


// The same quantity for up to 128 rows (trees up to 65 taxa: the headline size) with the matrix in REGISTERS: thread
// t owns the 8 x 8 tile (ti, tj), tj <= ti, of the lower triangle (136 tiles -> 160 threads for 126 rows).
//   Gram phase: J is staged 32 columns at a time, transposed ([column][row], so a tile's eight rows are 64 contiguous
// bytes: four 16-byte loads per operand and column, 64 multiply-adds per 8 loads -- the first kernel spends two loads
// per multiply-add and is bound by them: 209 of its 372 us at 126 x 320).
//   Elimination: symmetric (the Schur complement of a symmetric matrix is symmetric: A_ij -= A_ik A_jk / A_kk for
// i >= j > k), so a step only needs COLUMN k: its owners publish it (double-buffered: one barrier per step), every
// tile at or right of it updates its 64 entries from 16 published values.  Diagonal tiles keep both triangles
// (the upper one is never read).  Pivot order and singular-pivot behaviour as above.
constexpr int GT_LD = 130;  // row stride of the transposed chunk (even: 16-byte loads stay aligned)
// Position of row e = 8 t + 2 h + p inside a column of the chunk (and inside the published pivot column): [h][t][p].
// A thread reads the eight rows of its tile as four 16-byte pairs; with the rows in their natural order consecutive
// tiles are 64 bytes apart and a quarter-warp's loads fall into two bank groups (16 wavefronts per load, measured:
// the Gram phase took 460 cycles per column where the multiply-adds need 256); in this order they are contiguous.
__device__ __forceinline__ int gt_pos(int e) { return ((e >> 1) & 3) * 32 + ((e >> 3) << 1) + (e & 1); }
__global__ void __launch_bounds__(288) gram_logdet_tile_kernel(const double *__restrict__ J, int E, int C,
                                                               double *__restrict__ out) {
    // Two threads per tile, each with an 8 x 4 half (columns 4 h .. 4 h + 3 of the tile): 272 threads in 9 warps at 126
    // rows.  One thread per whole tile is 5 warps on 4 schedulers, and the scheduler that gets two of them paces both
    // phases (an fp64 warp instruction takes the pipe for two cycles whatever the number of active lanes).
    extern __shared__ __align__(16) double gt_sm[];  // Jc[2][32][GT_LD]: the chunk in use and the one in flight
    __shared__ __align__(16) double colk[2][128];
    __shared__ double pivs[2], part[32], s_val;
    __shared__ int s_stop;
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int nt = (E + 7) >> 3, ntiles = nt * (nt + 1) / 2;
    const int tile = tid >> 1, h = tid & 1;
    const bool active = tile < ntiles;
    int ti = 0, tj = 0;
    if (active) {
        ti = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
        while (ti * (ti + 1) / 2 > tile) --ti;
        tj = tile - ti * (ti + 1) / 2;
    }
    if (tid == 0) s_stop = 0;
    double A[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) A[r][c] = 0.0;
    const double *Jb = J + (size_t)b * E * C;
    // 8-byte asynchronous copies (global reads coalesced along the columns, written transposed; rows and columns
    // outside the matrix are zero-filled through the source size): chunk n + 1 lands while chunk n is consumed
    auto stage = [&](int c0, double *dst) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(dst);
        for (int idx = tid; idx < 128 * 32; idx += T) {
            const int e = idx >> 5, c = idx & 31;
            const bool in = e < E && c0 + c < C;
            const double *src = in ? Jb + (size_t)e * C + c0 + c : Jb;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(base + (unsigned)(c * GT_LD + gt_pos(e)) * 8u), "l"(src),
                         "r"(in ? 8 : 0)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, gt_sm);
    int cur = 0;
    for (int c0 = 0; c0 < C; c0 += 32, cur ^= 1) {
        const double *Jc = gt_sm + cur * (32 * GT_LD);
        if (c0 + 32 < C) {
            stage(c0 + 32, gt_sm + (cur ^ 1) * (32 * GT_LD));  // (its previous readers passed the barrier below one round ago)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (active) {
            // rows 8 ti + 2 m + p sit at plane m (m = 0..3), columns 8 tj + 4 h + 2 m' + p at plane 2 h + m' (m' = 0, 1)
            const double *pa = Jc + 2 * ti, *pb = Jc + 2 * tj + 64 * h;
#pragma unroll 2
            for (int c = 0; c < 32; ++c) {
                double a[8], bb[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const double2 x = *reinterpret_cast<const double2 *>(pa + c * GT_LD + 32 * m);
                    a[2 * m] = x.x; a[2 * m + 1] = x.y;
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const double2 y = *reinterpret_cast<const double2 *>(pb + c * GT_LD + 32 * m);
                    bb[2 * m] = y.x; bb[2 * m + 1] = y.y;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) A[r][q] = fma(a[r], bb[q], A[r][q]);
            }
        }
        __syncthreads();  // everyone is done with this chunk before the round after next refills its buffer
    }
    double pv[4];  // pivots held by this thread (diagonal tiles: columns 4 h .. 4 h + 3); logarithms after the elimination
#pragma unroll
    for (int r = 0; r < 4; ++r) pv[r] = 1.0;
    for (int kb = 0; kb < nt; ++kb) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = 8 * kb + kk;
            if (k >= E) break;
            const int hk = kk >> 2;     // which half of the tile column holds column k
            const int qk = kk & 3;      // ... and where in it
            double *ck = colk[kk & 1];
            if (active && tj == kb && h == hk) {  // owners of column k
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    *reinterpret_cast<double2 *>(ck + 32 * m + 2 * ti) = make_double2(A[2 * m][qk], A[2 * m + 1][qk]);
                if (ti == kb) {
                    const double piv = A[kk][qk];
                    pivs[kk & 1] = piv;
                    if (!(piv != 0.0)) {  // exactly singular (or NaN): log|det| = -inf (NaN propagates)
                        s_val = piv == piv ? -INFINITY : piv;
                        s_stop = 1;
                    } else {
                        pv[qk] = piv;
                    }
                }
            }
            __syncthreads();
            if (s_stop) {
                if (tid == 0) out[b] = s_val;
                return;
            }
            if (active && tj >= kb) {
                const double inv = 1.0 / pivs[kk & 1];
                double l[8], cj[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const double2 x = *reinterpret_cast<const double2 *>(ck + 32 * m + 2 * ti);
                    l[2 * m] = x.x * inv; l[2 * m + 1] = x.y * inv;
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const double2 y = *reinterpret_cast<const double2 *>(ck + 64 * h + 32 * m + 2 * tj);
                    cj[2 * m] = y.x; cj[2 * m + 1] = y.y;
                }
                if (tj > kb) {  // every row and column of the half lies beyond k
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int q = 0; q < 4; ++q) A[r][q] = fma(-l[r], cj[q], A[r][q]);
                } else if (h >= hk) {  // the tile column that holds k: columns beyond k only (none in the half left of it)
                    const bool whole = h > hk;        // the half right of the one that holds k
                    const bool below = ti > kb;       // (otherwise the diagonal tile: rows beyond k only)
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if ((whole || q > qk) && (below || r > kk)) A[r][q] = fma(-l[r], cj[q], A[r][q]);
                }
            }
        }
    }
    if (active && ti == tj) {
        double logabs = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) logabs += log(fabs(pv[r]));  // (columns beyond E keep the neutral 1.0)
        part[2 * ti + h] = logabs;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < 2 * nt; ++i) s += part[i];
        out[b] = 0.5 * s;
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
    // gridDim.z is limited to 65535: larger batches go in slices (the batched Jacobian cross-check asks for
    // B * (2S-2) matrices)
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = B - b0 < 65535 ? B - b0 : 65535;
        dim3 blk(32, 8), grd((S + 31) / 32, (S + 7) / 8, nb);
        pdm_fwd_kernel<<<grd, blk, 0, stream>>>(d_x + (size_t)b0 * S * D, sheet + (size_t)b0 * S * (D + 1), util + (size_t)b0 * S, S, D,
                                                d_curvature, variant, matsumoto, d_pdm + (size_t)b0 * S * S);
        DODO_LAUNCH_CHECK();
    }
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

extern "C" uint64_t dodo_blens_jacobian_workspace_bytes(int B, int S, int D) {
    return ((uint64_t)B * S * (D + 2) + (uint64_t)B * jac_ws_doubles(S, D)) * sizeof(double);
}

extern "C" int dodo_blens_jacobian(const double *d_x, int B, int S, int D, const double *d_curvature, int lift,
                                   int matsumoto, const int64_t *d_peel, const int32_t *d_flags, void *d_ws, double *d_J,
                                   void *stream_) {
    DODO_CHECK_ARG(d_x && d_curvature && d_peel && d_flags && d_ws && d_J, "dodo_blens_jacobian: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2 && D >= 1 && D <= MAX_D, "dodo_blens_jacobian: need B>=1, S>=2, 1<=D<=%d", MAX_D);
    DODO_CHECK_ARG(lift == DODO_LIFT_UP || lift == DODO_LIFT_WRAP, "dodo_blens_jacobian: bad lift");
    cudaStream_t stream = (cudaStream_t)stream_;
    double *sheet, *util;
    int rc = run_lift(d_x, B, S, D, lift, d_ws, stream, &sheet, &util);
    if (rc != DODO_OK) return rc;
    double *ws = util + (size_t)B * S;
    const int ne = S * (D + 1);
    // trees whose S^2 pair terms fit in shared memory (up to 118 taxa): cached pair terms, staged peel, operand
    // requests one join ahead, the lift as a kernel of its own
    const size_t sm = (size_t)2 * S * S * sizeof(double) + (size_t)(S - 1) * 4 * sizeof(int);
    if (sm <= 220 * 1024 && !getenv("DODO_JAC_PER_ELEMENT")) {
        if (lift == DODO_LIFT_UP && D + 1 <= 16 && !getenv("DODO_JAC_TWO_KERNELS")) {  // the lift inside the walk
            const int G = 32 / (D + 1), nw = (S + G - 1) / G;
            const int th = nw >= 16 ? 512 : nw * 32;
            if (sm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(blens_jac_cached_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            blens_jac_cached_kernel<true><<<B, th, sm, stream>>>(d_x, sheet, util, S, D, d_curvature, matsumoto, d_peel, d_flags, ws, d_J);
            DODO_LAUNCH_CHECK();
            return DODO_OK;
        }
        const int th = ne >= 512 ? 512 : ((ne + 31) / 32) * 32;
        if (sm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(blens_jac_cached_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        blens_jac_cached_kernel<false><<<B, th, sm, stream>>>(d_x, sheet, util, S, D, d_curvature, matsumoto, d_peel, d_flags, ws, d_J);
        DODO_LAUNCH_CHECK();
        const size_t items = (size_t)B * (2 * S - 2) * S;
        jac_lift_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(d_x, sheet, ws, B, S, D, lift, d_J);
        DODO_LAUNCH_CHECK();
        return DODO_OK;
    }
    const int threads = ne >= 1024 ? 1024 : ((ne + 31) / 32) * 32;
    blens_jac_kernel<<<B, threads, 0, stream>>>(d_x, sheet, util, S, D, d_curvature, lift, matsumoto, d_peel, d_flags, ws, d_J);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_gram_logdet(const double *d_J, int B, int E, int C, double *d_out, void *stream_) {
    DODO_CHECK_ARG(d_J && d_out, "dodo_gram_logdet: NULL argument");
    DODO_CHECK_ARG(B >= 1 && E >= 1 && C >= 1, "dodo_gram_logdet: need B, E, C >= 1");
    if (E <= 128) {  // the matrix in registers, 8 x 8 tiles of the lower triangle
        const int ntiles = ((E + 7) / 8) * ((E + 7) / 8 + 1) / 2;
        const int gsm = 2 * 32 * GT_LD * (int)sizeof(double);
        DODO_CUDA(cudaFuncSetAttribute(gram_logdet_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
        gram_logdet_tile_kernel<<<B, (2 * ntiles + 31) / 32 * 32, gsm, (cudaStream_t)stream_>>>(d_J, E, C, d_out);
        DODO_LAUNCH_CHECK();
        return DODO_OK;
    }
    const size_t smem = ((size_t)E * (E + 1) + (size_t)E * 33) * sizeof(double);
    if (smem > 227 * 1024 || E > 192) {
        set_error("dodo_gram_logdet: %d rows need %zu bytes of shared memory (limit 227 KB)", E, smem);
        return DODO_ENOTIMPL;
    }
    if (smem > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(gram_logdet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = E >= 64 ? 512 : (E >= 24 ? 256 : 64);
    gram_logdet_kernel<<<B, threads, smem, (cudaStream_t)stream_>>>(d_J, E, C, d_out);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

// K3: fused transition-matrix formation + Felsenstein pruning (forward and gradient), sm_100a.
//
// Replaces phylo.calculate_treelikelihood (dodonaphy/phylo.py:132-139) + torch autograd,
// Cphylo.calculate_treelikelihood / compute_LL_np (dodonaphy/cython/Cphylo.pyx:9-42) and
// PhyloModel.get_transition_mats (dodonaphy/phylomodel.py:76-114).
//
// Design: site patterns are conditionally independent given the tree, so one
// thread owns one (sample, pattern, rate category) and walks the whole tree itself.
//   * the reference's post-order (peel) is re-ordered per sample into a depth-first post-order in
//     which every join consumes the result of the join just before it (kept in registers); only the
//     "other" internal child has to be fetched back;
//   * per-node "messages" msg_n = P_n (msg_l * msg_r) are streamed once to a per-CTA scratch ring
//     (coalesced 512 B per warp) for the backward sweep; tips never touch HBM as partials: their
//     message is a 32-byte row of a per-edge table indexed by the 4-bit IUPAC mask;
//   * the backward sweep is the same schedule reversed (a valid pre-order), carries the upper
//     message in registers, reads each stored message once and reduces d lnL / d t_e over the 32
//     patterns of the warp with a fixed-order butterfly -> run-to-run reproducible.
// HBM traffic per (pattern, category): one 32 B write + one 32 B read per internal node -- 2/5 of
// the "materialise every partial" model used for the roofline denominator.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dodo {

constexpr int MAX_K = 8;

struct ModelDev {
    double U[16], Uinv[16], lam[4], Q[16], pi[4];
    double rates[MAX_K];
    int kind;
};

struct V4 {
    double a, b, c, d;
};

__device__ __forceinline__ V4 vmul(const V4 &x, const V4 &y) {
    return V4{x.a * y.a, x.b * y.b, x.c * y.c, x.d * y.d};
}

// y = P x, P row-major 16 doubles in global memory (warp-uniform address -> broadcast)
__device__ __forceinline__ V4 matvec_g(const double *__restrict__ P, const V4 &x) {
    V4 y;
    double2 p0 = ldg2(P + 0), p1 = ldg2(P + 2);
    y.a = fma(p1.y, x.d, fma(p1.x, x.c, fma(p0.y, x.b, p0.x * x.a)));
    p0 = ldg2(P + 4), p1 = ldg2(P + 6);
    y.b = fma(p1.y, x.d, fma(p1.x, x.c, fma(p0.y, x.b, p0.x * x.a)));
    p0 = ldg2(P + 8), p1 = ldg2(P + 10);
    y.c = fma(p1.y, x.d, fma(p1.x, x.c, fma(p0.y, x.b, p0.x * x.a)));
    p0 = ldg2(P + 12), p1 = ldg2(P + 14);
    y.d = fma(p1.y, x.d, fma(p1.x, x.c, fma(p0.y, x.b, p0.x * x.a)));
    return y;
}

// y = P^T x
__device__ __forceinline__ V4 matvecT_g(const double *__restrict__ P, const V4 &x) {
    V4 y;
    double2 p0 = ldg2(P + 0), p1 = ldg2(P + 2);
    y.a = p0.x * x.a, y.b = p0.y * x.a, y.c = p1.x * x.a, y.d = p1.y * x.a;
    p0 = ldg2(P + 4), p1 = ldg2(P + 6);
    y.a = fma(p0.x, x.b, y.a), y.b = fma(p0.y, x.b, y.b), y.c = fma(p1.x, x.b, y.c), y.d = fma(p1.y, x.b, y.d);
    p0 = ldg2(P + 8), p1 = ldg2(P + 10);
    y.a = fma(p0.x, x.c, y.a), y.b = fma(p0.y, x.c, y.b), y.c = fma(p1.x, x.c, y.c), y.d = fma(p1.y, x.c, y.d);
    p0 = ldg2(P + 12), p1 = ldg2(P + 14);
    y.a = fma(p0.x, x.d, y.a), y.b = fma(p0.y, x.d, y.b), y.c = fma(p1.x, x.d, y.c), y.d = fma(p1.y, x.d, y.d);
    return y;
}

// u . (Q x) with Q in the kernel-parameter (constant) bank
__device__ __forceinline__ double uQx(const double *Q, const V4 &u, const V4 &x) {
    double q0 = fma(Q[3], x.d, fma(Q[2], x.c, fma(Q[1], x.b, Q[0] * x.a)));
    double q1 = fma(Q[7], x.d, fma(Q[6], x.c, fma(Q[5], x.b, Q[4] * x.a)));
    double q2 = fma(Q[11], x.d, fma(Q[10], x.c, fma(Q[9], x.b, Q[8] * x.a)));
    double q3 = fma(Q[15], x.d, fma(Q[14], x.c, fma(Q[13], x.b, Q[12] * x.a)));
    return fma(u.d, q3, fma(u.c, q2, fma(u.b, q1, u.a * q0)));
}

// ------------------------------------------------------------------------------------------
// Schedule: peel (any valid post-order) -> depth-first post-order with "previous result" reuse.
// op = {a, b, n, flags}: node n = join(a, b).  b, when internal, is always the node produced by
// the previous op; a, when internal, is fetched from scratch.  flags bit0: n is the a-child of
// its parent (so it must be stored even in forward-only mode, and in the backward sweep its upper
// message comes from scratch rather than from the previous op).  The child with the larger
// Strahler number is evaluated first, which bounds the number of simultaneously pending a-children
// by log2(S).
// ------------------------------------------------------------------------------------------
__global__ void schedule_kernel(const int64_t *__restrict__ peel, int nb, int S, int4 *__restrict__ ops,
                                int *__restrict__ sched) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int nI = S - 1;
    const int64_t *pl = peel + (size_t)b * nI * 3;
    int *Lc = sched + (size_t)b * 6 * S;
    int *Rc = Lc + S, *st = Rc + S, *isz = st + S, *start = isz + S, *afl = start + S;
    int4 *op = ops + (size_t)b * nI;
    for (int i = 0; i < nI; ++i) { isz[i] = 0; afl[i] = 0; start[i] = 0; }
    const int Nn = 2 * S - 1;
    for (int i = 0; i < nI; ++i) {
        int l = (int)pl[i * 3 + 0], r = (int)pl[i * 3 + 1], p = (int)pl[i * 3 + 2];
        // memory safety for malformed input (the host shim validates and raises)
        l = min(max(l, 0), Nn - 1); r = min(max(r, 0), Nn - 1); p = min(max(p, S), Nn - 1);
        int idx = p - S;
        Lc[idx] = l; Rc[idx] = r;
        int sl = l < S ? 1 : st[l - S], sr = r < S ? 1 : st[r - S];
        st[idx] = (sl == sr) ? sl + 1 : max(sl, sr);
        isz[idx] = 1 + (l >= S ? isz[l - S] : 0) + (r >= S ? isz[r - S] : 0);
    }
    for (int i = nI - 1; i >= 0; --i) {
        int p = min(max((int)pl[i * 3 + 2], S), Nn - 1);
        int idx = p - S;
        int l = Lc[idx], r = Rc[idx];
        int a, bb;
        bool li = l >= S, ri = r >= S;
        if (li && ri) {
            if (st[l - S] >= st[r - S]) { a = l; bb = r; } else { a = r; bb = l; }
        } else if (li) { a = r; bb = l; }
        else { a = l; bb = r; }
        int s0 = (i == nI - 1) ? 0 : start[idx];
        int oi = s0 + isz[idx] - 1;
        oi = min(max(oi, 0), nI - 1);
        int flags = afl[idx] | ((i == nI - 1) ? 2 : 0);
        op[oi] = make_int4(a, bb, p, flags);
        if (li && ri) {
            start[a - S] = s0;
            start[bb - S] = s0 + isz[a - S];
            afl[a - S] = 1;
        } else if (bb >= S) {
            start[bb - S] = s0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Transition tables: P[b][e][k] (16 doubles) and, for tip edges, tiptab[b][t][k][mask][4] =
// sum over the states allowed by the mask of the columns of P (== P . tip_partial exactly).
// ------------------------------------------------------------------------------------------
__global__ void transition_kernel(ModelDev md, const double *__restrict__ blens, const double *__restrict__ P_in,
                                  int B, int S, int K, double *__restrict__ Pmat, double *__restrict__ tiptab) {
    const int E = 2 * S - 2;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * E * K;
    if (gid >= total) return;
    int k = gid % K;
    int e = (gid / K) % E;
    int b = gid / ((size_t)K * E);
    double P[16];
    if (P_in != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i) P[i] = P_in[gid * 16 + i];
    } else {
        double t = blens[(size_t)b * E + e] * md.rates[k];
        if (md.kind == DODO_MODEL_JC69) {
            // phylomodel.py:87-94 / Cphylo.pyx:45-51, same operation order
            double ex = exp(-4.0 / 3.0 * t);
            double pa = 0.25 + 3.0 / 4.0 * ex;
            double pb = 0.25 - 0.25 * ex;
#pragma unroll
            for (int i = 0; i < 16; ++i) P[i] = ((i >> 2) == (i & 3)) ? pa : pb;
        } else {
            double ex[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) ex[m] = exp(md.lam[m] * t);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double acc = (md.U[i * 4 + 0] * ex[0]) * md.Uinv[0 * 4 + j];
                    acc += (md.U[i * 4 + 1] * ex[1]) * md.Uinv[1 * 4 + j];
                    acc += (md.U[i * 4 + 2] * ex[2]) * md.Uinv[2 * 4 + j];
                    acc += (md.U[i * 4 + 3] * ex[3]) * md.Uinv[3 * 4 + j];
                    P[i * 4 + j] = acc;
                }
        }
    }
    {
        double2 *out = reinterpret_cast<double2 *>(Pmat + gid * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = make_double2(P[2 * i], P[2 * i + 1]);
    }
    if (e < S) {
        double2 *tt = reinterpret_cast<double2 *>(tiptab + (((size_t)b * S + e) * K + k) * 64);
        for (int m = 0; m < 16; ++m) {
            double v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m & (1 << j)) acc += P[i * 4 + j];
                v[i] = acc;
            }
            tt[m * 2 + 0] = make_double2(v[0], v[1]);
            tt[m * 2 + 1] = make_double2(v[2], v[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Main kernel
// ------------------------------------------------------------------------------------------
struct PruneArgs {
    const uint8_t *tipmask;
    int64_t ld_mask;
    const double *weights;
    const int4 *ops;
    int ops_batch;  // 1 => shared topology
    const double *Pmat;
    const double *tiptab;
    double2 *scratch;
    double *ll_part;  // [n_items]
    double *g_part;   // [n_items][E]
    double *gm_part;  // [n_items][...] (GMATS)  -- reserved
    int B, S, L, K;
    int n_tiles, n_items, tile_patterns;
    int mix;
};

template <bool GRAD>
__global__ void __launch_bounds__(256) prune_kernel(PruneArgs A, ModelDev md) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarp = nthr >> 5;
    const int S = A.S, K = A.K, E = 2 * S - 2, nops = S - 1;
    const int k = warp % K;
    const int pg = warp / K;
    double *gacc = smem;                     // [nwarp][E]
    double *fx = smem + (size_t)nwarp * E;   // [nwarp][32]
    double *red = fx + nwarp * 32;           // [nwarp]
    if (GRAD) {
        for (int i = tid; i < nwarp * E; i += nthr) gacc[i] = 0.0;
    }
    __syncthreads();
    double2 *scr = A.scratch + (size_t)blockIdx.x * (size_t)(S - 1) * 2 * nthr;
    const double rate_k = md.rates[k];

    for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
        const int b = item / A.n_tiles;
        const int tile = item - b * A.n_tiles;
        int p = tile * A.tile_patterns + pg * 32 + lane;
        const bool valid = p < A.L;
        if (!valid) p = A.L - 1;
        const double wt = valid ? __ldg(A.weights + p) : 0.0;
        const int4 *ops_b = A.ops + (size_t)(A.ops_batch == 1 ? 0 : b) * nops;
        const double *Pm = A.Pmat + ((size_t)b * E * K + k) * 16;
        const double *tt = A.tiptab + (((size_t)b * S) * K + k) * 64;
        const uint8_t *tm = A.tipmask + p;

        auto tip_msg = [&](int t) -> V4 {
            int m = __ldg(tm + (size_t)t * A.ld_mask) & 15;
            const double *row = tt + ((size_t)t * K) * 64 + m * 4;
            double2 x = ldg2(row), y = ldg2(row + 2);
            return V4{x.x, x.y, y.x, y.y};
        };
        auto ld_scr = [&](int idx) -> V4 {
            double2 x = __ldcg(scr + ((size_t)idx * 2 + 0) * nthr + tid);
            double2 y = __ldcg(scr + ((size_t)idx * 2 + 1) * nthr + tid);
            return V4{x.x, x.y, y.x, y.y};
        };
        auto st_scr = [&](int idx, const V4 &v) {
            __stcg(scr + ((size_t)idx * 2 + 0) * nthr + tid, make_double2(v.a, v.b));
            __stcg(scr + ((size_t)idx * 2 + 1) * nthr + tid, make_double2(v.c, v.d));
        };

        // ---------------- forward (post-order) ----------------
        V4 prev = V4{0, 0, 0, 0};
        V4 rootL = prev;
        for (int i = 0; i < nops; ++i) {
            const int4 op = __ldg(ops_b + i);
            V4 a = (op.x < S) ? tip_msg(op.x) : ld_scr(op.x - S);
            V4 bv = (op.y < S) ? tip_msg(op.y) : prev;
            V4 Lv = vmul(a, bv);
            if (i == nops - 1) {
                rootL = Lv;
            } else {
                prev = matvec_g(Pm + (size_t)op.z * K * 16, Lv);
                if (GRAD || (op.w & 1)) st_scr(op.z - S, prev);
            }
        }
        double fk = fma(md.pi[3], rootL.d, fma(md.pi[2], rootL.c, fma(md.pi[1], rootL.b, md.pi[0] * rootL.a)));
        double f = fk;
        if (A.mix && K > 1) {
            fx[warp * 32 + lane] = fk;
            __syncthreads();
            double s = 0.0;
            for (int kk = 0; kk < K; ++kk) s += fx[(pg * K + kk) * 32 + lane];
            f = s / (double)K;
        }
        double ll = 0.0;
        if (!(A.mix && K > 1) || k == 0) ll = wt * log(f);
        ll = warp_sum(ll);
        if (lane == 0) red[warp] = ll;

        // ---------------- backward (the same schedule reversed = pre-order) ----------------
        if (GRAD) {
            const double sc = (A.mix && K > 1) ? wt / (f * (double)K) : wt / f;
            V4 wprev = V4{md.pi[0] * sc, md.pi[1] * sc, md.pi[2] * sc, md.pi[3] * sc};
            double *gw = gacc + (size_t)warp * E;
            for (int i = nops - 1; i >= 0; --i) {
                const int4 op = __ldg(ops_b + i);
                V4 w = (op.w & 1) ? ld_scr(op.z - S) : wprev;
                const bool ai = op.x >= S, bi = op.y >= S;
                V4 a = ai ? ld_scr(op.x - S) : tip_msg(op.x);
                V4 bv = bi ? ld_scr(op.y - S) : tip_msg(op.y);
                V4 ua = vmul(w, bv);
                V4 ub = vmul(w, a);
                double ga = rate_k * uQx(md.Q, ua, a);
                double gb = rate_k * uQx(md.Q, ub, bv);
                ga = warp_sum(ga);
                gb = warp_sum(gb);
                if (lane == 0) {
                    gw[op.x] += ga;
                    gw[op.y] += gb;
                }
                if (ai) st_scr(op.x - S, matvecT_g(Pm + (size_t)op.x * K * 16, ua));
                if (bi) wprev = matvecT_g(Pm + (size_t)op.y * K * 16, ub);
            }
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w2 = 0; w2 < nwarp; ++w2) s += red[w2];
            A.ll_part[item] = s;
        }
        if (GRAD) {
            double *gp = A.g_part + (size_t)item * E;
            for (int e = tid; e < E; e += nthr) {
                double s = 0.0;
                for (int w2 = 0; w2 < nwarp; ++w2) {
                    s += gacc[(size_t)w2 * E + e];
                    gacc[(size_t)w2 * E + e] = 0.0;
                }
                gp[e] = s;
            }
        }
        __syncthreads();
    }
}

// fixed-order reduction over the pattern tiles of every sample
__global__ void prune_reduce_kernel(const double *__restrict__ ll_part, const double *__restrict__ g_part,
                                    int B, int n_tiles, int E, double *__restrict__ loglik,
                                    double *__restrict__ grad_blens) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int per = E + 1;
    if (gid >= B * per) return;
    int b = gid / per, e = gid - b * per;
    if (e == E) {
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t) s += ll_part[(size_t)b * n_tiles + t];
        loglik[b] = s;
    } else if (grad_blens != nullptr) {
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t) s += g_part[((size_t)b * n_tiles + t) * E + e];
        grad_blens[(size_t)b * E + e] = s;
    }
}

}  // namespace dodo

using namespace dodo;

extern "C" int dodo_prune_plan(int B, int S, int L, int K, uint32_t flags, int grid_hint, int threads_hint,
                               dodo_prune_plan_t *plan) {
    DODO_CHECK_ARG(plan != nullptr, "dodo_prune_plan: plan is NULL");
    DODO_CHECK_ARG(B >= 1 && S >= 2 && L >= 1, "dodo_prune_plan: need B>=1, S>=2, L>=1 (got %d,%d,%d)", B, S, L);
    DODO_CHECK_ARG(K >= 1 && K <= MAX_K, "dodo_prune_plan: K must be in [1,%d] (got %d)", MAX_K, K);
    if (flags & DODO_PRUNE_GMATS) {
        set_error("dodo_prune_plan: DODO_PRUNE_GMATS not implemented yet");
        return DODO_ENOTIMPL;
    }
    if (flags & DODO_PRUNE_SCALE) {
        set_error("dodo_prune_plan: DODO_PRUNE_SCALE not implemented yet");
        return DODO_ENOTIMPL;
    }
    memset(plan, 0, sizeof(*plan));
    plan->B = B; plan->S = S; plan->L = L; plan->K = K; plan->flags = flags;
    int groups = K <= 8 ? 8 / K : 1;
    if (groups < 1) groups = 1;
    int threads = 32 * K * groups;
    if (threads_hint > 0) {
        DODO_CHECK_ARG(threads_hint % (32 * K) == 0 && threads_hint <= 256,
                       "dodo_prune_plan: threads must be a multiple of 32*K and <= 256");
        threads = threads_hint;
        groups = threads / (32 * K);
    }
    plan->threads = threads;
    plan->tile_patterns = 32 * groups;
    plan->n_tiles = (L + plan->tile_patterns - 1) / plan->tile_patterns;
    long long items = (long long)B * plan->n_tiles;
    DODO_CHECK_ARG(items < (1ll << 31), "dodo_prune_plan: too many work items");
    plan->n_items = (int)items;
    int grid = grid_hint > 0 ? grid_hint : 2 * sm_count();
    if (grid > plan->n_items) grid = plan->n_items;
    plan->grid = grid;
    const uint64_t E = 2ull * S - 2, nI = S - 1;
    uint64_t off = 0;
    plan->off_ops = off;     off = align_up(off + (uint64_t)B * nI * sizeof(int4), 256);
    plan->off_sched = off;   off = align_up(off + (uint64_t)B * 6 * S * sizeof(int), 256);
    plan->off_pmat = off;    off = align_up(off + (uint64_t)B * E * K * 16 * sizeof(double), 256);
    plan->off_tiptab = off;  off = align_up(off + (uint64_t)B * S * K * 64 * sizeof(double), 256);
    plan->off_scratch = off; off = align_up(off + (uint64_t)grid * nI * 2 * threads * sizeof(double2), 256);
    plan->off_ll_part = off; off = align_up(off + (uint64_t)plan->n_items * sizeof(double), 256);
    plan->off_g_part = off;
    if (flags & DODO_PRUNE_GRAD) off = align_up(off + (uint64_t)plan->n_items * E * sizeof(double), 256);
    plan->off_gm_part = off;
    plan->ws_bytes = off;
    return DODO_OK;
}

extern "C" int dodo_prune(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
                          const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights,
                          const int64_t *d_peel, int peel_batch, const double *d_blens, const double *d_P_in,
                          void *d_ws, double *d_loglik, double *d_grad_blens, double *d_grad_P,
                          double *d_grad_pi, void *stream_) {
    DODO_CHECK_ARG(plan && model && d_tipmask && d_weights && d_peel && d_ws && d_loglik, "dodo_prune: NULL argument");
    DODO_CHECK_ARG(d_blens || d_P_in, "dodo_prune: need blens or explicit transition matrices");
    const int B = plan->B, S = plan->S, L = plan->L, K = plan->K;
    DODO_CHECK_ARG(peel_batch == 1 || peel_batch == B, "dodo_prune: peel_batch must be 1 or B");
    DODO_CHECK_ARG(ld_mask >= L, "dodo_prune: ld_mask < L");
    const bool grad = plan->flags & DODO_PRUNE_GRAD;
    DODO_CHECK_ARG(!grad || d_grad_blens, "dodo_prune: GRAD requested but d_grad_blens is NULL");
    DODO_CHECK_ARG(!(grad && d_P_in), "dodo_prune: blens gradient needs device-formed P(t) (use GMATS for explicit mats)");
    (void)d_grad_P; (void)d_grad_pi;
    cudaStream_t stream = (cudaStream_t)stream_;
    char *ws = (char *)d_ws;
    const int E = 2 * S - 2;

    ModelDev md;
    memcpy(md.U, model->U, sizeof(md.U));
    memcpy(md.Uinv, model->Uinv, sizeof(md.Uinv));
    memcpy(md.lam, model->lam, sizeof(md.lam));
    memcpy(md.Q, model->Q, sizeof(md.Q));
    memcpy(md.pi, model->pi, sizeof(md.pi));
    md.kind = model->kind;
    for (int k = 0; k < MAX_K; ++k) md.rates[k] = (k < K) ? (rates ? rates[k] : 1.0) : 0.0;

    int4 *ops = (int4 *)(ws + plan->off_ops);
    int *sched = (int *)(ws + plan->off_sched);
    double *Pmat = (double *)(ws + plan->off_pmat);
    double *tiptab = (double *)(ws + plan->off_tiptab);

    {
        int nb = peel_batch;
        schedule_kernel<<<(nb + 63) / 64, 64, 0, stream>>>(d_peel, nb, S, ops, sched);
        DODO_LAUNCH_CHECK();
    }
    {
        size_t total = (size_t)B * E * K;
        transition_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(md, d_blens, d_P_in, B, S, K, Pmat, tiptab);
        DODO_LAUNCH_CHECK();
    }
    PruneArgs A;
    A.tipmask = d_tipmask; A.ld_mask = ld_mask; A.weights = d_weights;
    A.ops = ops; A.ops_batch = peel_batch; A.Pmat = Pmat; A.tiptab = tiptab;
    A.scratch = (double2 *)(ws + plan->off_scratch);
    A.ll_part = (double *)(ws + plan->off_ll_part);
    A.g_part = (double *)(ws + plan->off_g_part);
    A.gm_part = nullptr;
    A.B = B; A.S = S; A.L = L; A.K = K;
    A.n_tiles = plan->n_tiles; A.n_items = plan->n_items; A.tile_patterns = plan->tile_patterns;
    A.mix = (plan->flags & DODO_PRUNE_MIX) ? 1 : 0;
    const int nwarp = plan->threads / 32;
    size_t smem = ((size_t)nwarp * E + (size_t)nwarp * 32 + nwarp) * sizeof(double);
    if (grad) {
        if (smem > 48 * 1024)
            DODO_CUDA(cudaFuncSetAttribute(prune_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prune_kernel<true><<<plan->grid, plan->threads, smem, stream>>>(A, md);
    } else {
        if (smem > 48 * 1024)
            DODO_CUDA(cudaFuncSetAttribute(prune_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prune_kernel<false><<<plan->grid, plan->threads, smem, stream>>>(A, md);
    }
    DODO_LAUNCH_CHECK();
    {
        int total = B * (E + 1);
        prune_reduce_kernel<<<(total + 127) / 128, 128, 0, stream>>>(A.ll_part, A.g_part, B, plan->n_tiles, E, d_loglik,
                                                                      grad ? d_grad_blens : nullptr);
        DODO_LAUNCH_CHECK();
    }
    return DODO_OK;
}

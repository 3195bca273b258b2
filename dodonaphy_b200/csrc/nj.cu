// K2: batched neighbour joining -- one tree per CTA, sm_100a.
//
// dodo_nj_hard   replaces Cpeeler.nj_np (dodonaphy/cython/Cpeeler.pyx:15-125).  Bit-identical peel
//                and blens: the running row sums (_nj_xsub) are updated with the reference's
//                operation order using explicitly rounded adds/multiplies (no FMA contraction), the
//                new node's row sum is accumulated sequentially in pool order, and the join is the
//                FIRST strict minimum of q in pool order (pool order == ascending node id).
// dodo_nj_soft   replaces peeler.nj_torch (dodonaphy/peeler.py:155-220): compute_Q (:223-248),
//                soft.argmin (dodonaphy/soft.py:38-53) evaluated from the last row of each SoftSort
//                matrix only (O(entries) instead of O(entries^2)), get_new_dist_soft (:251-265).
// dodo_nj_soft_bwd  vector-Jacobian product d blens / d pdm (what torch autograd computes through
//                the one-hot row selections of peeler.py:196-213): a reverse replay of the joins.
//
// Matrix state lives in the caller's workspace (8 S^2 bytes per tree, L1/L2-resident for small trees,
// streamed from HBM at thousands of taxa); the active list and per-node scalars live in shared memory.
// A new node takes a fresh storage slot and live rows are repacked when the spares run out (nj_spare).
// Both scans keep several independent matrix loads in flight per lane; the soft variant keeps its row
// sums in double-double and updates them per join instead of re-reading the matrix (see two_sum).
#include <float.h>
#include <stdlib.h>

#include <math.h>
#include <string.h>

#include "common.cuh"

namespace dodo {

constexpr double LARGE = 1e9;  // peeler.py:11

struct QKey {
    double q;
    unsigned long long key;
};

__device__ __forceinline__ QKey qmin(QKey a, QKey b) {
    return (b.q < a.q || (b.q == a.q && b.key < a.key)) ? b : a;
}

__device__ __forceinline__ QKey warp_qmin(QKey v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        QKey o;
        o.q = __shfl_xor_sync(0xffffffffu, v.q, m);
        o.key = __shfl_xor_sync(0xffffffffu, v.key, m);
        v = qmin(v, o);
    }
    return v;
}

// Row prefetch of the streaming scans (general kernels: the matrix lives in HBM and every join reads all of it
// once).  A scan is one warp per matrix row; its loads are issued a trip at a time (8 per lane), so between trips a
// warp has nothing in flight and the scan ends up bound by HBM latency, not bandwidth (ncu, 2000 taxa: long-scoreboard
// stalls 10.8 warps per issue, DRAM 47 %).  The TMA unit's bulk L2 prefetch takes a whole row segment in ONE
// instruction by one lane -- no registers, no shared memory, no completion barrier: each warp asks for the row it
// will scan NJ_PF_ROWS rows from now, and the scan's own loads then find their sectors in L2.
#ifndef NJ_PF_ROWS
#define NJ_PF_ROWS 0
#endif
__device__ __forceinline__ void prefetch_l2_bulk(const double *p, int n_doubles) {
    if (n_doubles <= 0) return;
    const unsigned long long a = (unsigned long long)p, a0 = a & ~15ull;  // 16-byte granules; rows are 128-byte aligned
    const unsigned bytes = ((unsigned)(a - a0) + (unsigned)n_doubles * 8u + 15u) & ~15u;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(bytes) : "memory");
}

// Storage scheme of the NJ kernels.  A new node always takes a FRESH slot at the end of the matrix, so
// slot order == pool order (survivors in their order, then new nodes in creation order) and a row scan
// in pool order walks memory monotonically.  Every join retires two slots; once the spare slots are
// used up the live rows/columns are repacked densely into the other half of the workspace.  With
// `spare` free slots after a repack the dead-slot overhead of the scans is ~spare/n and the repacking
// overhead ~4/spare of the scan traffic, hence spare = 2 sqrt(n).
__host__ __device__ inline int nj_spare(int n) {
    const int s = (int)(2.0 * sqrt((double)n));
    return s < 16 ? 16 : s;
}
// row stride for n live rows: spare slots, rounded so that every row starts on a 128-byte line
__host__ __device__ inline int nj_ld(int n) { return (n + nj_spare(n) + 15) & ~15; }
// workspace per tree (doubles): two ld x ld matrices (the backward kernel uses the second half as its
// S x S adjoint matrix); ints are in shared memory
__host__ __device__ inline size_t nj_ws_doubles(int S) {
    return 2 * (size_t)nj_ld(S) * nj_ld(S) + ((4 * (size_t)S + 15) & ~(size_t)15);
}

// ------------------------------------------------------------------------------------------ prior
// Gamma-Dirichlet(aT, bT, a, c) log prior of one tree's branch lengths and its gradient
// (BaseModel.compute_prior_gamma_dir dodonaphy/base_model.py:475-527 with utils.LogDirPrior
// dodonaphy/utils.py:251-267; NumPy twin Cphylo.pyx:53-114), evaluated by ONE warp as the epilogue of
// the NJ kernel that has just produced the branch lengths (or on its own, dodo_prior_gamma_dir), so the
// 2S-2 numbers never travel to the host per sample.  Fixed-order reductions.
struct PriorDev {
    double aT, bT, a, c, konst;  // konst: normalising constants + topology term, from the host (lgamma)
    int variant;                 // DODO_PRIOR_TORCH: clamp at 1e-4 when any branch is <= 0; DODO_PRIOR_NUMPY: eps floors
    double *ln_prior;            // [B]
    double *grad;                // [B][2S-2] or NULL
};

__device__ void prior_warp(const double *__restrict__ bl, int S, const PriorDev &pr, int b, int lane) {
    const int E = 2 * S - 2;
    double tot = 0.0;
    int nonpos = 0;
    for (int e = lane; e < E; e += 32) {
        const double v = bl[e];
        tot += v;
        nonpos |= !(v > 0.0);
    }
    tot = warp_sum(tot);
    nonpos = __any_sync(0xffffffffu, nonpos);
    const bool np_variant = pr.variant == DODO_PRIOR_NUMPY;
    const double floor_v = np_variant ? DBL_EPSILON : (nonpos ? 1e-4 : 0.0);
    double tipb = 0.0, intb = 0.0;
    for (int e = lane; e < E; e += 32) {
        const double v = floor_v > 0.0 ? fmax(bl[e], floor_v) : bl[e];
        if (e < S) tipb += log(v); else intb += log(v);
    }
    tipb = warp_sum(tipb);
    intb = warp_sum(intb);
    const double treeL = np_variant ? fmax(tot, DBL_EPSILON) : tot;
    const double coef = pr.aT - pr.a * S - pr.a * pr.c * (S - 3);
    if (lane == 0)
        pr.ln_prior[b] = (pr.a - 1.0) * tipb + (pr.a * pr.c - 1.0) * intb + coef * log(treeL) - pr.bT * treeL + pr.konst;
    if (pr.grad != nullptr) {
        double *g = pr.grad + (size_t)b * E;
        for (int e = lane; e < E; e += 32) {
            const double v = bl[e];
            const bool clamped = floor_v > 0.0 && v < floor_v;  // a clamp passes no gradient below its floor
            const double w = e < S ? pr.a - 1.0 : pr.a * pr.c - 1.0;
            g[e] = (clamped ? 0.0 : w / v) + coef / treeL - pr.bT;
        }
    }
}

__global__ void prior_kernel(const double *__restrict__ blens, int B, int S, PriorDev pr) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b < B) prior_warp(blens + (size_t)b * (2 * S - 2), S, pr, b, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------ hard
#ifndef HARD_MLP
#define HARD_MLP 8
#endif
__global__ void __launch_bounds__(1024, 1) nj_hard_kernel(const double *__restrict__ pdm, int S, double *__restrict__ ws,
                                                       int64_t *__restrict__ peel, double *__restrict__ blens,
                                                       PriorDev prior, int lean_first) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    double *Dm = ws + (size_t)b * nj_ws_doubles(S);
    if (lean_first) {  // nj_hard_lean_kernel ran first: only the trees it handed over are done here
        const int todo = *reinterpret_cast<const int *>(Dm);
        if (__syncthreads_or(0) || todo == 0) return;
    }
    const int LD = nj_ld(S);
    double *Dalt = Dm + (size_t)LD * LD;  // second matrix of the workspace: target of the next repack
    int ld = LD;                          // row stride of the current matrix
    int used = S;                         // slots handed out so far
    const double *P = pdm + (size_t)b * S * S;
    double *xs = reinterpret_cast<double *>(smem_raw);  // [LD] by storage slot
    double *newd = xs + LD;                             // [LD]
    int *act = reinterpret_cast<int *>(newd + LD);      // [LD] pool position -> storage slot
    int *act2 = act + LD;                               // [LD]
    int *id_of = act2 + LD;                             // [LD] storage slot -> node id
    __shared__ QKey wbest[32];
    __shared__ int sh_pa, sh_pb;
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *bl = blens + (size_t)b * (2 * S - 2);

    for (size_t i = tid; i < (size_t)S * S; i += T) {
        const int r = (int)(i / S), c = (int)(i - (size_t)r * S);
        Dm[(size_t)r * ld + c] = P[i];
    }
    for (int s = tid; s < S; s += T) {
        act[s] = s;
        id_of[s] = s;
        // Cpeeler.pyx:46-54: sequential sum over the other pool members, in pool order
        double acc = 0.0;
        const double *row = P + (size_t)s * S;
        for (int j = 0; j < S; ++j)
            if (j != s) acc = __dadd_rn(acc, row[j]);
        xs[s] = acc;
    }
    __syncthreads();

    for (int n = S; n > 1; --n) {
        if (used == ld) {  // no free slot left: repack the live rows/columns densely in pool order
            const int nld = nj_ld(n);
            for (int p = warp; p < n; p += nwarp) {
                const double *src = Dm + (size_t)act[p] * ld;
                double *dst = Dalt + (size_t)p * nld;
                for (int q = lane; q < n; q += 32) dst[q] = src[act[q]];
            }
            for (int p = tid; p < n; p += T) { newd[p] = xs[act[p]]; act2[p] = id_of[act[p]]; }
            __syncthreads();
            for (int p = tid; p < n; p += T) { xs[p] = newd[p]; id_of[p] = act2[p]; act[p] = p; }
            double *t = Dm; Dm = Dalt; Dalt = t;
            ld = nld;
            used = n;
            __syncthreads();
        }
        const double nm2 = (double)n - 2.0;
        // ---- first strict minimum of q over pairs (pa < pb) in pool order (Cpeeler.pyx:58-66)
        QKey best;
        best.q = DBL_MAX;
        best.key = ~0ull;
        auto prefetch_row = [&](int pr) {  // columns pr+1 .. n-1 of pool row pr: slots act[pr+1] .. act[n-1], contiguous
            if (NJ_PF_ROWS > 0 && lane == 0 && pr < n - 1) {
                const int s0 = act[pr + 1];
                prefetch_l2_bulk(Dm + (size_t)act[pr] * ld + s0, act[n - 1] + 1 - s0);
            }
        };
#pragma unroll
        for (int k = 0; k < NJ_PF_ROWS; ++k) prefetch_row(warp + k * nwarp);
        for (int pa = warp; pa < n - 1; pa += nwarp) {
            prefetch_row(pa + NJ_PF_ROWS * nwarp);
            const int sa = act[pa];
            const double xa = xs[sa];
            const double *row = Dm + (size_t)sa * ld;
            int pb = pa + 1 + lane;
            // HARD_MLP independent loads in flight per lane (the scan is otherwise latency-bound)
            for (; pb + 32 * (HARD_MLP - 1) < n; pb += 32 * HARD_MLP) {
                unsigned sb[HARD_MLP];  // unsigned: one IMAD.WIDE per address instead of a sign-extended 64-bit add
                double d[HARD_MLP];
#pragma unroll
                for (int k = 0; k < HARD_MLP; ++k) sb[k] = (unsigned)act[pb + 32 * k];
#pragma unroll
                for (int k = 0; k < HARD_MLP; ++k) d[k] = row[sb[k]];
#pragma unroll
                for (int k = 0; k < HARD_MLP; ++k) {
                    const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, d[k]), xa), xs[sb[k]]);
                    if (q < best.q) {
                        best.q = q;
                        best.key = ((unsigned long long)pa << 32) | (unsigned)(pb + 32 * k);
                    }
                }
            }
            for (; pb < n; pb += 32) {
                const int sb = act[pb];
                double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, row[sb]), xa), xs[sb]);
                if (q < best.q) {
                    best.q = q;
                    best.key = ((unsigned long long)pa << 32) | (unsigned)pb;
                }
            }
        }
        best = warp_qmin(best);
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (warp == 0) {
            QKey v = lane < nwarp ? wbest[lane] : QKey{DBL_MAX, ~0ull};
            v = warp_qmin(v);
            if (lane == 0) {
                sh_pa = (int)(v.key >> 32);
                sh_pb = (int)(v.key & 0xffffffffu);
                if (v.key == ~0ull) { sh_pa = 0; sh_pb = 1; }  // all-NaN distances: stay memory-safe
            }
        }
        __syncthreads();
        const int pa = sh_pa, pb = sh_pb;
        const int sa = act[pa], sb = act[pb];
        const int n1 = id_of[sa], n2 = id_of[sb];
        const int parent = S + (S - n);
        const int ns = used++;  // the new node's slot
        const double d12 = Dm[(size_t)sa * ld + sb];
        // ---- distances of the new node, running row sums (Cpeeler.pyx:81-95); the survivors move up in
        // the pool (pool = survivors in order, then the new node) and the new row/column is written
        for (int p = tid; p < n; p += T) {
            if (p == pa || p == pb) continue;
            const int sx = act[p];
            const double *rx = Dm + (size_t)sx * ld;
            double v1 = __dadd_rn(__dadd_rn(0.0, rx[sa]), rx[sb]);
            double dist = __dmul_rn(0.5, __dsub_rn(v1, d12));
            double x = __dadd_rn(xs[sx], dist);
            x = __dsub_rn(x, Dm[(size_t)sa * ld + sx]);
            x = __dsub_rn(x, Dm[(size_t)sb * ld + sx]);
            xs[sx] = x;
            const int pc = p - (p > pa) - (p > pb);
            newd[pc] = dist;  // dense, in pool order: summed sequentially below
            act2[pc] = sx;
            Dm[(size_t)ns * ld + sx] = dist;
            Dm[(size_t)sx * ld + ns] = dist;
        }
        if (tid == 0) act2[n - 2] = ns;
        __syncthreads();
        if (tid == 0) {
            // branch lengths (Cpeeler.pyx:98-112) with the joined nodes' row sums as they stand
            double df, dg;
            if (n > 2) {
                double v1 = __dmul_rn(0.5, d12);
                double v4 = __dmul_rn(__ddiv_rn(1.0, __dmul_rn(2.0, nm2)), __dsub_rn(xs[sa], xs[sb]));
                df = __dadd_rn(v1, v4);
                dg = __dsub_rn(d12, df);
            } else {
                df = __ddiv_rn(d12, 2.0);
                dg = df;
            }
            bl[n1] = fmax(df, DBL_EPSILON);  // Cpeeler.pyx:124
            bl[n2] = fmax(dg, DBL_EPSILON);
            pl[(S - n) * 3 + 0] = n1;
            pl[(S - n) * 3 + 1] = n2;
            pl[(S - n) * 3 + 2] = parent;
            // new node's row sum: sequential in pool order (Cpeeler.pyx:92).  One dependent add per
            // survivor is the critical path of a join at a few hundred taxa, so the operands come
            // from a dense array with the loads running ahead of the adds.
            double acc = 0.0;
            const int m = n - 2;
            int p = 0;
#pragma unroll 2
            for (; p + 8 <= m; p += 8) {
                const double2 u0 = *reinterpret_cast<const double2 *>(newd + p);
                const double2 u1 = *reinterpret_cast<const double2 *>(newd + p + 2);
                const double2 u2 = *reinterpret_cast<const double2 *>(newd + p + 4);
                const double2 u3 = *reinterpret_cast<const double2 *>(newd + p + 6);
                acc = __dadd_rn(acc, u0.x); acc = __dadd_rn(acc, u0.y);
                acc = __dadd_rn(acc, u1.x); acc = __dadd_rn(acc, u1.y);
                acc = __dadd_rn(acc, u2.x); acc = __dadd_rn(acc, u2.y);
                acc = __dadd_rn(acc, u3.x); acc = __dadd_rn(acc, u3.y);
            }
            for (; p < m; ++p) acc = __dadd_rn(acc, newd[p]);
            xs[ns] = acc;
            id_of[ns] = parent;
        }
        for (int p = tid; p < n - 1; p += T) act[p] = act2[p];
        __syncthreads();
    }
    if (prior.ln_prior != nullptr && warp == 0) prior_warp(bl, S, prior, b, lane);  // epilogue (bl complete: barrier above)
}

// ------------------------------------------------------------------------------------------ soft
// Entry enumeration for the soft argmin.  The reference flattens tril(Q) of the padded N x N matrix
// row-major (peeler.py:194): flat = row*N + col.  Active mode visits only active pairs row > col in
// that order; inactive entries (strict upper triangle = 0, masked/diagonal = 1e9) are provably
// negligible under the guard computed below, and their contribution to the second softmax is added
// in closed form.  Full mode visits all N^2 entries (exact in every regime).
struct SoftCtx {
    const double *Dm;
    const double *rs;    // row sums by storage slot
    const int *act;      // pool position -> storage slot (ascending node id)
    const int *id_of;    // storage slot -> node id
    const int *pos_of;   // node id -> pool position or -1
    int S, N, n;
    double nm2;
    long long dense_n;   // MODE 2: Dm is a dense row-major array of dense_n entries with N columns
    int symmetric;       // Dm[i][j] == Dm[j][i] bit for bit: the transposed load can be skipped
};

__device__ __forceinline__ double soft_Q(const SoftCtx &c, int srow, int scol) {
    // compute_Q (peeler.py:223-248): Q' = (n-2) d - r_i - r_j ; Q = (Q' + Q'^T)/2
    const double dij = c.Dm[(size_t)srow * c.S + scol];
    const double dji = c.symmetric ? dij : c.Dm[(size_t)scol * c.S + srow];
    double a = (c.nm2 * dij - c.rs[srow]) - c.rs[scol];
    double b = (c.nm2 * dji - c.rs[scol]) - c.rs[srow];
    return 0.5 * (a + b);
}

// l-th active pair in row-major lower-triangle order -> (prow, pcol)
__device__ __forceinline__ void tri_unrank(long long l, int &prow, int &pcol) {
    int r = (int)((1.0 + sqrt(1.0 + 8.0 * (double)l)) * 0.5);
    while ((long long)r * (r - 1) / 2 > l) --r;
    while ((long long)(r + 1) * r / 2 <= l) ++r;
    prow = r;
    pcol = (int)(l - (long long)r * (r - 1) / 2);
}

// MODE 0: active pairs of the NJ state; 1: all N^2 entries of the padded NJ matrix; 2: a dense array
template <int MODE>
struct EntryIter {
    static constexpr bool FULL = MODE != 0;
    const SoftCtx &c;
    long long l, end;
    int prow, pcol;
    __device__ EntryIter(const SoftCtx &c_, long long l0, long long l1) : c(c_), l(l0), end(l1) {
        prow = 0; pcol = 0;
        if (!FULL && l0 < l1) tri_unrank(l0, prow, pcol);
    }
    __device__ __forceinline__ bool valid() const { return l < end; }
    __device__ __forceinline__ void get(double &s, double &flat) const {
        if (MODE == 2) {
            s = c.Dm[l];
            flat = (double)l;
        } else if (FULL) {
            int row = (int)(l / c.N), col = (int)(l - (long long)row * c.N);
            flat = (double)l;
            if (row < col) { s = 0.0; return; }
            int pr = c.pos_of[row], pc = c.pos_of[col];
            if (row == col || pr < 0 || pc < 0) { s = LARGE; return; }
            s = soft_Q(c, c.act[pr], c.act[pc]);
        } else {
            int srow = c.act[prow], scol = c.act[pcol];
            flat = (double)c.id_of[srow] * (double)c.N + (double)c.id_of[scol];
            s = soft_Q(c, srow, scol);
        }
    }
    __device__ __forceinline__ void next() {
        ++l;
        if (!FULL) {
            if (++pcol == prow) { ++prow; pcol = 0; }
        }
    }
};

__device__ __forceinline__ double block_reduce(double v, double *sh, int op /*0 sum,1 min,2 max*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, v, m);
        v = op == 0 ? v + o : (op == 1 ? fmin(v, o) : fmax(v, o));
    }
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    // second stage: every warp folds the (<= 32) per-warp values with shuffles, in the same fixed order
    double r = lane < nwarp ? sh[lane] : (op == 0 ? 0.0 : (op == 1 ? DBL_MAX : -DBL_MAX));
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, r, m);
        r = op == 0 ? r + o : (op == 1 ? fmin(r, o) : fmax(r, o));
    }
    return r;
}

// exclusive prefix over threads (by thread id) of v; also returns the total
__device__ __forceinline__ double block_excl_scan(double v, double *sh, double &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    double off = 0.0, tot = 0.0;
    for (int w = 0; w < nwarp; ++w) {
        if (w < warp) off += sh[w];
        tot += sh[w];
    }
    total = tot;
    return off + inc - v;
}

// unravel_index (soft.py:7-10) of a (possibly fractional) flat index: row = trunc(idx / N),
// col = fmod(idx, N).  idx - row * N is exact whenever row is the true quotient; the library fmod is
// only needed when the rounded division lands on the wrong side of an integer.
__device__ __forceinline__ void unravel(double idx, double N, double &rowf, double &colf) {
    rowf = trunc(idx / N);
    colf = fma(-rowf, N, idx);
    if (!(colf >= 0.0 && colf < N)) colf = fmod(idx, N);
}

// exp(-a) for a >= 0; exactly 0 beyond the fp64 underflow point, without paying for the call
__device__ __forceinline__ double exp_neg(double a) { return a > 750.0 ? 0.0 : exp(-a); }

// Fast path of the soft argmin.  With a small temperature only the handful of entries within
// W = 750 tau of the minimum have a non-zero softmax weight in fp64 (exp underflows to exactly 0 beyond
// that), so ONE sweep suffices: every thread tracks its running minimum and the entries inside the
// window below it; after the block-wide minimum is known the survivors are gathered (<= CAND_CAP),
// ordered by flat index and the two softmaxes of soft.py:42-52 are evaluated on them, all other
// entries entering in closed form (c_j = 0 => weight exp(-max c / tau) each).  Any overflow falls
// back to the general multi-pass evaluation below, so the result is the same in every regime.
constexpr int CAND_LOCAL = 4, CAND_CAP = 128;

// soft.argmin (soft.py:42-52) from the candidates inside the 750 tau window (one thread; the candidates are
// sorted in place by flat index = the reference's cumsum order); every other entry enters in closed form.
// idx = the soft index, guard = 1.0 when the closed-form treatment of the inactive entries is exact in fp64.
__device__ void soft_eval_candidates(double *cand_s, double *cand_f, int n, double m, double tau, double NN, bool FULL,
                                     double &idx, double &guard) {
    {
        const double away = fmin(fabs(0.0 - m), fabs(LARGE - m)) / tau;
        if (n == 1 && cand_s[0] == m && 1.0 / tau > 750.0 && (FULL || away > 750.0)) {
            // the usual case: one entry carries all the weight (p = c = 1, every other weight is exactly 0)
            idx = cand_f[0];
            guard = 1.0;
        } else {
            for (int i = 1; i < n; ++i) {  // order by flat index (the reference's cumsum order)
                double ks = cand_s[i], kf = cand_f[i];
                int j = i - 1;
                while (j >= 0 && cand_f[j] > kf) { cand_s[j + 1] = cand_s[j]; cand_f[j + 1] = cand_f[j]; --j; }
                cand_s[j + 1] = ks; cand_f[j + 1] = kf;
            }
            double Z = 0.0;
            for (int i = 0; i < n; ++i) {  // the first softmax's weights replace the scores (each is used three times)
                cand_s[i] = exp_neg(fabs(cand_s[i] - m) / tau);
                Z += cand_s[i];
            }
            double run = 0.0, M = 0.0;
            for (int i = 0; i < n; ++i) {
                double e = cand_s[i];
                run += e;
                M = fmax(M, (e / Z) * (run / Z));
            }
            const double e0 = exp_neg(M / tau);
            double z2 = 0.0, isum = 0.0, fs = 0.0;
            run = 0.0;
            for (int i = 0; i < n; ++i) {
                double e = cand_s[i];
                double e2 = e0;
                if (e != 0.0) {
                    run += e;
                    e2 = exp_neg(fabs(M - (e / Z) * (run / Z)) / tau);
                }
                z2 += e2;
                isum = fma(e2, cand_f[i], isum);
                fs += cand_f[i];
            }
            z2 += (NN - (double)n) * e0;
            isum += e0 * (NN * (NN - 1.0) * 0.5 - fs);
            idx = isum / z2;
            double worst = FULL ? 0.0 : exp_neg(fmin(fabs(0.0 - m), fabs(LARGE - m)) / tau);
            guard = (FULL || NN * worst <= 1e-18 * fmin(1.0, tau) * Z) ? 1.0 : 0.0;
        }
    }
}

template <int MODE>
__device__ bool soft_argmin_fast(const SoftCtx &c, double tau, double *sh, double &idx_out, bool &guard_ok) {
    constexpr bool FULL = MODE != 0;
    __shared__ double cand_s[CAND_CAP], cand_f[CAND_CAP];
    __shared__ int cand_n, cand_over;
    const int tid = threadIdx.x, T = blockDim.x;
    const long long P = MODE == 2 ? c.dense_n : (FULL ? (long long)c.N * c.N : (long long)c.n * (c.n - 1) / 2);
    const long long chunk = (P + T - 1) / T;
    const long long l0 = min(P, (long long)tid * chunk), l1 = min(P, l0 + chunk);
    const double W = 750.0 * tau;
    double bs[CAND_LOCAL], bf[CAND_LOCAL];
    int nb = 0;
    bool over = false;
    double lm = DBL_MAX;
    auto visit = [&](double s, double f) {
        if (s < lm) lm = s;
        if (s <= lm + W) {
            if (nb == CAND_LOCAL) {  // drop what has fallen out of the window
                int w = 0;
#pragma unroll
                for (int i = 0; i < CAND_LOCAL; ++i)
                    if (bs[i] <= lm + W) { bs[w] = bs[i]; bf[w] = bf[i]; ++w; }
                nb = w;
            }
            if (nb == CAND_LOCAL) over = true;
            else {
#pragma unroll
                for (int i = 0; i < CAND_LOCAL; ++i)
                    if (i == nb) { bs[i] = s; bf[i] = f; }
                ++nb;
            }
        }
    };
    if (MODE == 0) {
        // one matrix row per warp, lanes across its columns: coalesced reads of the row
        const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
        for (int prow = 1 + warp; prow < c.n; prow += nwarp) {
            const int srow = c.act[prow];
            const double frow = (double)c.id_of[srow] * (double)c.N;
            const double *rowp = c.Dm + (size_t)srow * c.S;
            const double rsr = c.rs[srow];
            // compute_Q (peeler.py:223-248) as in soft_Q, with the row's operands hoisted
            auto entry = [&](int scol, double dij, double dji) {
                const double rsc = c.rs[scol];
                const double a = (c.nm2 * dij - rsr) - rsc;
                const double b = (c.nm2 * dji - rsc) - rsr;
                const double q = 0.5 * (a + b);
                // almost every entry is outside the window: the flat index (an int -> double
                // conversion on the quarter-rate pipe) is only formed for the few that are not
                if (q <= lm + W) visit(q, frow + (double)c.id_of[scol]);
            };
            int pcol = lane;
            // four independent loads in flight per lane: one load per trip leaves the scan bound by
            // memory latency (1 KB in flight per warp), not by bandwidth
            for (; pcol + 96 < prow; pcol += 128) {
                int sc[4];
                double d[4], dt[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) sc[k] = c.act[pcol + 32 * k];
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = rowp[sc[k]];
#pragma unroll
                for (int k = 0; k < 4; ++k) dt[k] = c.symmetric ? d[k] : c.Dm[(size_t)sc[k] * c.S + srow];
#pragma unroll
                for (int k = 0; k < 4; ++k) entry(sc[k], d[k], dt[k]);
            }
            for (; pcol < prow; pcol += 32) {
                const int scol = c.act[pcol];
                const double dij = rowp[scol];
                entry(scol, dij, c.symmetric ? dij : c.Dm[(size_t)scol * c.S + srow]);
            }
        }
    } else {
        for (EntryIter<MODE> it(c, l0, l1); it.valid(); it.next()) {
            double s, f;
            it.get(s, f);
            visit(s, f);
        }
    }
    if (tid == 0) { cand_n = 0; cand_over = 0; }
    double m = lm;
    if (!FULL) m = fmin(m, 0.0);  // the strict upper triangle holds zeros (torch.tril)
    m = block_reduce(m, sh, 1);   // (contains the barriers that publish cand_n / cand_over)
    if (over) cand_over = 1;
#pragma unroll
    for (int i = 0; i < CAND_LOCAL; ++i)
        if (i < nb && bs[i] <= m + W) {
            int slot = atomicAdd(&cand_n, 1);
            if (slot < CAND_CAP) { cand_s[slot] = bs[i]; cand_f[slot] = bf[i]; }
        }
    __syncthreads();
    const int n = cand_n;
    const bool ok = !cand_over && n <= CAND_CAP && n >= 1;
    if (!ok) { __syncthreads(); return false; }
    const double NN = MODE == 0 ? (double)c.N * (double)c.N : (double)P;
    if (tid == 0) soft_eval_candidates(cand_s, cand_f, n, m, tau, NN, FULL, sh[0], sh[1]);
    __syncthreads();
    idx_out = sh[0];
    guard_ok = sh[1] != 0.0;
    // no trailing barrier: sh and the candidate buffers are next written behind the leading barrier of a
    // block_reduce (or the caller's own barrier), which every thread reaches only after these reads
    return true;
}

// Active-pair scan with TWO running minima per thread (the test nj_soft_lean_kernel uses): one coalesced sweep
// finds the smallest and the second smallest Q of the live pairs; when the smallest stands alone inside the 750 tau
// window (and clear of the zeros / 1e9 of the padded matrix) its flat index IS the soft argmin -- soft_argmin_fast's
// one-candidate branch -- and nothing else has to be remembered during the sweep: per entry one load, four
// flops, one compare.  (The window bookkeeping of soft_argmin_fast costs 1.9x the instructions of the hard scan
// at the kernel's 64-register cap: ncu, 2000 taxa.)  Otherwise a second sweep gathers the entries inside the
// window and soft_eval_candidates evaluates them; more than CAND_CAP of them: false, the caller goes on.
#ifndef SOFT_MLP
#define SOFT_MLP 8
#endif
template <bool SYM>  // SYM: the matrix is symmetric bit for bit (compile-time: the per-entry test and the transposed load drop out of the scan)
__device__ bool soft_argmin_two_minima(const SoftCtx &c, double tau, double *sh, double &idx_out, bool &guard_ok) {
    __shared__ double t_m1[32], t_m2[32], t_cs[CAND_CAP], t_cf[CAND_CAP];
    __shared__ int t_row[32], t_col[32], t_n;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const double W = 750.0 * tau, Wsure = W * (1.0 + 1e-9);
    double m1 = DBL_MAX, m2 = DBL_MAX;
    int brow = 0, bcol = 0;
    auto prefetch_row = [&](int pr) {  // columns 0 .. pr-1 of pool row pr: slots 0 .. act[pr-1]
        if (NJ_PF_ROWS > 0 && lane == 0 && pr < c.n) prefetch_l2_bulk(c.Dm + (size_t)c.act[pr] * c.S, c.act[pr - 1] + 1);
    };
    auto sweep = [&](auto &&visit) {
#pragma unroll
        for (int k = 0; k < NJ_PF_ROWS; ++k) prefetch_row(1 + warp + k * nwarp);
        for (int prow = 1 + warp; prow < c.n; prow += nwarp) {
            prefetch_row(prow + NJ_PF_ROWS * nwarp);
            const int srow = c.act[prow];
            const double *rowp = c.Dm + (size_t)srow * c.S;
            const double rsr = c.rs[srow];
            auto entry = [&](int scol, double dij, double dji) {
                const double rsc = c.rs[scol];
                const double a = (c.nm2 * dij - rsr) - rsc;
                const double b = (c.nm2 * dji - rsc) - rsr;
                visit(0.5 * (a + b), srow, scol);
            };
            // SOFT_MLP independent loads in flight per lane, the ragged end of the row included (predicated): a tail
            // of single loads was 15 % of the stall samples at 2000 taxa, one dependent L2/HBM latency per trip
            // (148 trees: 476 -> 427 ms; costs ~10 % at a few hundred taxa, where the lean kernel or the hard
            // variant are the ones in use)
            // Full trips carry no predicates and index with unsigned slots (one IMAD.WIDE per address instead of a
            // sign-extended 64-bit add); only the last trip of a row is predicated.
            int pcol = lane;
            for (; pcol + 32 * (SOFT_MLP - 1) < prow; pcol += 32 * SOFT_MLP) {
                unsigned sc[SOFT_MLP];
                double d[SOFT_MLP];
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k) sc[k] = (unsigned)c.act[pcol + 32 * k];
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k) d[k] = rowp[sc[k]];
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k) entry((int)sc[k], d[k], SYM ? d[k] : c.Dm[(size_t)sc[k] * c.S + srow]);
            }
            if (pcol < prow) {
                int sc[SOFT_MLP];
                double d[SOFT_MLP];
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k) sc[k] = pcol + 32 * k < prow ? c.act[pcol + 32 * k] : -1;
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k) d[k] = sc[k] >= 0 ? rowp[sc[k]] : 0.0;
#pragma unroll
                for (int k = 0; k < SOFT_MLP; ++k)
                    if (sc[k] >= 0) entry(sc[k], d[k], SYM ? d[k] : c.Dm[(size_t)sc[k] * c.S + srow]);
            }
        }
    };
    sweep([&](double q, int srow, int scol) {
        if (q < m2) {  // rare once the two running minima have settled
            if (q < m1) { m2 = m1; m1 = q; brow = srow; bcol = scol; }
            else m2 = q;
        }
    });
    auto fold = [&](int m) {
        const double o1 = __shfl_xor_sync(0xffffffffu, m1, m), o2 = __shfl_xor_sync(0xffffffffu, m2, m);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, m), ocol = __shfl_xor_sync(0xffffffffu, bcol, m);
        const bool lt = o1 < m1;  // an exact tie leaves m2 == m1: not the one-candidate regime
        const double lo1 = lt ? o1 : m1, hi1 = lt ? m1 : o1, oth = lt ? o2 : m2;
        m2 = oth < hi1 ? oth : hi1;
        m1 = lo1;
        brow = lt ? orow : brow;
        bcol = lt ? ocol : bcol;
    };
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) fold(m);
    __syncthreads();  // (the buffers below may still be read from the previous join)
    if (lane == 0) { t_m1[warp] = m1; t_m2[warp] = m2; t_row[warp] = brow; t_col[warp] = bcol; }
    if (tid == 0) t_n = 0;
    __syncthreads();
    m1 = lane < nwarp ? t_m1[lane] : DBL_MAX;
    m2 = lane < nwarp ? t_m2[lane] : DBL_MAX;
    brow = lane < nwarp ? t_row[lane] : 0;
    bcol = lane < nwarp ? t_col[lane] : 0;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) fold(m);
    m1 = __shfl_sync(0xffffffffu, m1, 0); m2 = __shfl_sync(0xffffffffu, m2, 0);
    brow = __shfl_sync(0xffffffffu, brow, 0); bcol = __shfl_sync(0xffffffffu, bcol, 0);
    if (1.0 / tau > 750.0 && m1 < 0.0 && m2 > m1 + W && -m1 > Wsure && LARGE - m1 > Wsure) {
        idx_out = (double)c.id_of[brow] * (double)c.N + (double)c.id_of[bcol];  // row id > column id (tril order)
        guard_ok = true;
        return true;
    }
    const double mm = fmin(m1, 0.0);  // the strict upper triangle holds zeros (torch.tril)
    sweep([&](double q, int srow, int scol) {
        if (q <= mm + W) {
            const int slot = atomicAdd(&t_n, 1);
            if (slot < CAND_CAP) { t_cs[slot] = q; t_cf[slot] = (double)c.id_of[srow] * (double)c.N + (double)c.id_of[scol]; }
        }
    });
    __syncthreads();
    const int nc = t_n;
    if (nc < 1 || nc > CAND_CAP) return false;
    if (tid == 0) soft_eval_candidates(t_cs, t_cf, nc, mm, tau, (double)c.N * (double)c.N, false, sh[0], sh[1]);
    __syncthreads();
    idx_out = sh[0];
    guard_ok = sh[1] != 0.0;
    return true;
}

template <int MODE>
__device__ void soft_argmin_block(const SoftCtx &c, double tau, double *sh, double &idx_out, bool &guard_ok) {
    if (MODE == 0) {
        if (c.symmetric ? soft_argmin_two_minima<true>(c, tau, sh, idx_out, guard_ok)
                        : soft_argmin_two_minima<false>(c, tau, sh, idx_out, guard_ok)) return;
    } else if (soft_argmin_fast<MODE>(c, tau, sh, idx_out, guard_ok)) return;
    constexpr bool FULL = MODE != 0;
    const int tid = threadIdx.x, T = blockDim.x;
    const long long P = MODE == 2 ? c.dense_n : (FULL ? (long long)c.N * c.N : (long long)c.n * (c.n - 1) / 2);
    const long long chunk = (P + T - 1) / T;
    const long long l0 = min(P, (long long)tid * chunk), l1 = min(P, l0 + chunk);
    // pass 1: min
    double m = DBL_MAX;
    for (EntryIter<MODE> it(c, l0, l1); it.valid(); it.next()) {
        double s, f;
        it.get(s, f);
        m = fmin(m, s);
    }
    if (!FULL) m = fmin(m, 0.0);  // the strict upper triangle holds zeros (torch.tril)
    m = block_reduce(m, sh, 1);
    // pass 2: sum of exp(-|s-m|/tau) in flat order
    double loc = 0.0;
    for (EntryIter<MODE> it(c, l0, l1); it.valid(); it.next()) {
        double s, f;
        it.get(s, f);
        loc += exp_neg(fabs(s - m) / tau);
    }
    double Z;
    double off = block_excl_scan(loc, sh, Z);
    double NN = (double)c.N * (double)c.N;
    if (!FULL) {
        // inactive entries: zeros and 1e9 -- require their softmax mass to be invisible in fp64
        double worst = exp_neg(fmin(fabs(0.0 - m), fabs(LARGE - m)) / tau);
        guard_ok = (NN * worst <= 1e-18 * fmin(1.0, tau) * Z);
    } else {
        guard_ok = true;
    }
    // pass 3: c_j = p_j * cumsum(p)_j ; M = max c
    double run = off, M = 0.0;
    for (EntryIter<MODE> it(c, l0, l1); it.valid(); it.next()) {
        double s, f;
        it.get(s, f);
        double e = exp_neg(fabs(s - m) / tau);
        if (e != 0.0) {
            run += e;
            M = fmax(M, (e / Z) * (run / Z));
        }
    }
    M = block_reduce(M, sh, 2);
    // pass 4: second softmax on -c, soft index = sum p'_j * flat_j
    run = off;
    const double e0 = exp_neg(M / tau);  // entries with c_j == 0
    double z2 = 0.0, isum = 0.0, fsum = 0.0;
    for (EntryIter<MODE> it(c, l0, l1); it.valid(); it.next()) {
        double s, f;
        it.get(s, f);
        double e = exp_neg(fabs(s - m) / tau);
        double e2 = e0;
        if (e != 0.0) {
            run += e;
            double cj = (e / Z) * (run / Z);
            e2 = exp_neg(fabs(M - cj) / tau);
        }
        z2 += e2;
        isum = fma(e2, f, isum);
        fsum += f;
    }
    z2 = block_reduce(z2, sh, 0);
    isum = block_reduce(isum, sh, 0);
    if (!FULL) {
        fsum = block_reduce(fsum, sh, 0);
        double cnt = NN - (double)P;
        z2 += cnt * e0;
        isum += e0 * (NN * (NN - 1.0) * 0.5 - fsum);
    }
    idx_out = isum / z2;
}

// Row sums of the soft NJ state are kept as unevaluated sums hi + lo (double-double) and updated per
// join with error-free transformations: r_x <- r_x - d(x,l) - d(x,r) + d(x,u).  The value the joins
// see (hi) is the mathematical row sum rounded once, i.e. at least as close to it as the reference's
// fresh torch.sum (peeler.py:240) in any summation order -- without re-reading the whole matrix
// every join (the row-sum pass was 2/3 of the kernel's DRAM traffic at 2000 taxa).
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
__device__ __forceinline__ void dd_add(double &hi, double &lo, double x) {
    double s, e;
    two_sum(hi, x, s, e);
    e = __dadd_rn(e, lo);
    hi = __dadd_rn(s, e);
    lo = __dsub_rn(e, __dsub_rn(hi, s));
}
__device__ __forceinline__ void dd_add_dd(double &hi, double &lo, double xh, double xl) {
    double s, e;
    two_sum(hi, xh, s, e);
    e = __dadd_rn(e, __dadd_rn(lo, xl));
    hi = __dadd_rn(s, e);
    lo = __dsub_rn(e, __dsub_rn(hi, s));
}

#ifndef NJ_SOFT_MAXT
#define NJ_SOFT_MAXT 1024  // tuning: a smaller bound gives the scan more registers per thread (more loads in flight per lane)
#endif
__global__ void __launch_bounds__(NJ_SOFT_MAXT, 1) nj_soft_kernel(const double *__restrict__ pdm, int S, double tau, int full,
                                                       double *__restrict__ ws, int64_t *__restrict__ peel,
                                                       double *__restrict__ blens, int32_t *__restrict__ flags,
                                                       PriorDev prior, int lean_first) {
    extern __shared__ unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int N = 2 * S - 1;
    double *Dm = ws + (size_t)b * nj_ws_doubles(S);
    if (lean_first) {  // nj_soft_lean_kernel ran first: only the trees it handed over are done here
        const int todo = *reinterpret_cast<const int *>(Dm);
        if (__syncthreads_or(0) || todo == 0) return;  // (the barrier orders the read before the writes below)
    }
    const int LD = nj_ld(S);
    double *Dalt = Dm + (size_t)LD * LD;  // repack target (storage scheme: see nj_ws_doubles)
    int ld = LD, used = S;
    const double *P = pdm + (size_t)b * S * S;
    double *rs = reinterpret_cast<double *>(smem_raw);  // [LD] by storage slot: row sums (hi)
    double *rs_lo = rs + LD;                            // [LD] ... and their low parts
    double *newd = rs_lo + LD;                          // [LD]
    double *sh = newd + LD;                             // [32]
    int *act = reinterpret_cast<int *>(sh + 32);        // [LD]
    int *act2 = act + LD;                               // [LD]
    int *id_of = act2 + LD;                             // [LD]
    int *pos_of = id_of + LD;                           // [N]
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *bl = blens + (size_t)b * (2 * S - 2);
    int32_t *fl = flags + (size_t)b * (S - 1);

    int asym = 0;
    for (size_t i = tid; i < (size_t)S * S; i += T) {
        int r = (int)(i / S), cc = (int)(i - (size_t)r * S);
        const double v = P[i];
        Dm[(size_t)r * ld + cc] = (r == cc) ? 0.0 : v;  // dists.fill_diagonal_(0.0), peeler.py:178
        if (r != cc && v != P[(size_t)cc * S + r]) asym = 1;
    }
    const int symmetric = __syncthreads_or(asym) ? 0 : 1;  // new rows/columns are written symmetrically
    for (int s = tid; s < S; s += T) { act[s] = s; id_of[s] = s; }
    for (int i = tid; i < N; i += T) pos_of[i] = i < S ? i : -1;
    __syncthreads();
    // row sums over the active set (peeler.py:240, :261): one warp per row, compensated
    auto row_sums = [&](int n) {
        for (int p = warp; p < n; p += nwarp) {
            const int sx = act[p];
            const double *row = Dm + (size_t)sx * ld;
            double hi = 0.0, lo = 0.0;
            for (int q = lane; q < n; q += 32) {
                int sy = act[q];
                if (sy != sx) dd_add(hi, lo, row[sy]);
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double oh = __shfl_xor_sync(0xffffffffu, hi, m), ol = __shfl_xor_sync(0xffffffffu, lo, m);
                dd_add_dd(hi, lo, oh, ol);
            }
            if (lane == 0) { rs[sx] = hi; rs_lo[sx] = lo; }
        }
        __syncthreads();
    };
    row_sums(S);

    for (int n = S; n > 1; --n) {
        if (used == ld) {  // no free slot left: repack live rows/columns densely in pool order
            const int nld = nj_ld(n);
            for (int p = warp; p < n; p += nwarp) {
                const double *src = Dm + (size_t)act[p] * ld;
                double *dst = Dalt + (size_t)p * nld;
                for (int q = lane; q < n; q += 32) dst[q] = src[act[q]];
            }
            for (int p = tid; p < n; p += T) { act2[p] = id_of[act[p]]; newd[p] = rs[act[p]]; }
            __syncthreads();
            for (int p = tid; p < n; p += T) { id_of[p] = act2[p]; rs[p] = newd[p]; newd[p] = rs_lo[act[p]]; }
            __syncthreads();
            for (int p = tid; p < n; p += T) { rs_lo[p] = newd[p]; act[p] = p; }
            double *t = Dm; Dm = Dalt; Dalt = t;
            ld = nld;
            used = n;
            __syncthreads();
        }
        // an asymmetric input has no single d(x, l) to take out of a row sum: recompute
        if (!symmetric && n < S) row_sums(n);
        SoftCtx c{Dm, rs, act, id_of, pos_of, ld, N, n, (double)(n - 2), 0, symmetric};
        double idx;
        bool ok;
        if (full) soft_argmin_block<1>(c, tau, sh, idx, ok);
        else soft_argmin_block<0>(c, tau, sh, idx, ok);
        // unravel (soft.py:7-10) and round (peeler.py:196-198)
        double rowf, colf;
        unravel(idx, (double)N, rowf, colf);
        int right = (int)rint(rowf), left = (int)rint(colf);
        int fbits = ok ? 0 : 4;
        bool bad = !(idx == idx) || right < 0 || right >= N || left < 0 || left >= N || right == left;
        int pr = bad ? -1 : pos_of[right], pc = bad ? -1 : pos_of[left];
        if (pr < 0 || pc < 0) {
            // the reference would index a masked node here (garbage); flag it and stop this tree
            if (tid == 0) {
                for (int t = S - n; t < S - 1; ++t) fl[t] = 8;
                if (prior.ln_prior != nullptr) prior.ln_prior[b] = __longlong_as_double(0x7ff8000000000000ll);
            }
            return;
        }
        const int sl = act[pc], sr = act[pr];
        const int parent = S + (S - n);
        const int ns = used++;  // the new node's slot
        const double d_lr = Dm[(size_t)sl * ld + sr];
        const double nact = fmax((double)n, 3.0);
        double d_pl = 0.5 * d_lr + (rs[sl] - rs[sr]) / (2.0 * (nact - 2.0));
        if (!(d_pl > DBL_EPSILON)) { fbits |= (d_pl == DBL_EPSILON) ? 0 : 1; d_pl = DBL_EPSILON; }
        double d_pr = d_lr - d_pl;
        if (!(d_pr > DBL_EPSILON)) { fbits |= (d_pr == DBL_EPSILON) ? 0 : 2; d_pr = DBL_EPSILON; }
        const int plo = min(pc, pr), phi = max(pc, pr);
        for (int p = tid; p < n; p += T) {
            if (p == pc || p == pr) continue;
            const int sx = act[p];
            const double dl = Dm[(size_t)sl * ld + sx], dr = Dm[(size_t)sr * ld + sx];
            const double u = 0.5 * ((dl + dr) - d_lr);
            newd[sx] = u;
            // the survivor moves up in the pool (pool = survivors in order, then the new node) and the
            // new node's row/column is written
            const int p2 = p - (p > plo) - (p > phi);
            act2[p2] = sx;
            pos_of[id_of[sx]] = p2;
            Dm[(size_t)ns * ld + sx] = u;
            Dm[(size_t)sx * ld + ns] = u;
            if (symmetric) {
                double hi = rs[sx], lo = rs_lo[sx];
                dd_add(hi, lo, -dl);
                dd_add(hi, lo, -dr);
                dd_add(hi, lo, u);
                rs[sx] = hi; rs_lo[sx] = lo;
            }
        }
        __syncthreads();
        if (symmetric && warp == 0) {  // row sum of the new node: one warp, fixed order, double-double
            double hi = 0.0, lo = 0.0;
            for (int p = lane; p < n; p += 32)
                if (p != pc && p != pr) dd_add(hi, lo, newd[act[p]]);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double oh = __shfl_xor_sync(0xffffffffu, hi, m), ol = __shfl_xor_sync(0xffffffffu, lo, m);
                dd_add_dd(hi, lo, oh, ol);
            }
            if (lane == 0) { rs[ns] = hi; rs_lo[ns] = lo; }
        }
        if (tid == 0) {
            pl[(S - n) * 3 + 0] = left;
            pl[(S - n) * 3 + 1] = right;
            pl[(S - n) * 3 + 2] = parent;
            bl[left] = d_pl;
            bl[right] = d_pr;
            fl[S - n] = fbits;
            id_of[ns] = parent;
            pos_of[left] = -1;
            pos_of[right] = -1;
            pos_of[parent] = n - 2;
            act2[n - 2] = ns;
        }
        __syncthreads();
        { int *t = act; act = act2; act2 = t; }  // the new pool order takes over
    }
    if (prior.ln_prior != nullptr && warp == 0) prior_warp(bl, S, prior, b, lane);  // epilogue (bl complete: barrier above)
}

// ------------------------------------------------------------------------------------------ lean soft
// Small trees (<= NJ_LEAN_MAX taxa) are latency-bound: one CTA per tree, S-1 dependent joins.  The
// general kernel above spends ~650 instructions per warp and six block-wide barriers on every join
// regardless of n (ncu: 40 % of the stall samples at barriers, 31 % issue slots at 64 taxa): 4.4 us per
// join.  This variant is built for a short dependent chain per join instead:
//   * the matrix is resident in shared memory and kept DENSE (slots 0..n-1 live: the new node takes the
//     lower of the two joined slots, the last live slot moves into the other), so the scan needs no
//     indirection; node ids travel with the slots (the flat index of peeler.py:194 only needs the ids);
//   * W <= 8 warps per tree and TWO barriers per join: [scan + warp reduction] | [merge, decision, new
//     distances + slot move + row-sum updates in one pass, with the per-slot scalars double-buffered so
//     that nobody overwrites what a slower thread still reads];
//   * the scan tracks the smallest and the second smallest Q per thread: the regime nearly every join is in
//     -- ONE entry inside the 750 tau window below which the softmax weights are exactly 0 in fp64
//     (soft_argmin_fast's `n == 1` branch: idx = flat index of that entry) -- is recognised from those
//     two numbers; otherwise (always for the last joins: with <= 4 nodes left Q has exact mathematical
//     ties) the candidates are gathered and evaluated by the same soft_eval_candidates;
//   * the new node's row sum follows from the old ones, r_u = (r_l + r_r - n d_lr) / 2, evaluated in
//     double-double by one lane while the others update the matrix (the general kernel sums the new row:
//     both are the mathematical row sum to ~1e-32, then rounded);
//   * one lane forms the branch lengths (an fp64 division) off the critical path.
// What it does not handle -- an asymmetric input, tau >= 1/750, more than CAND_CAP candidates, an invalid
// selection -- is left to the general kernel: this kernel writes todo = 1 into the first word of the
// tree's workspace and stops; nj_soft_kernel, launched right behind it with `lean_first`, redoes exactly
// those trees from scratch and skips the others.  Same expressions as the general kernel for Q, the
// branch lengths, the new distances and the double-double row-sum updates.
constexpr int NJ_LEAN_MAX = 160;
__host__ __device__ inline size_t nj_lean_smem_bytes(int S, int W) {
    return ((size_t)S * (S | 1) + 4 * (size_t)S + 3 * (size_t)W + 2 + 2 * 128) * sizeof(double) +
           (2 * (size_t)S + (2 * (size_t)S - 1) + W + 2) * sizeof(int);
}

template <int W>
__global__ void __launch_bounds__(32 * W) nj_soft_lean_kernel(const double *__restrict__ pdm, int S, double tau,
                                                            double *__restrict__ ws, int64_t *__restrict__ peel,
                                                            double *__restrict__ blens, int32_t *__restrict__ flags,
                                                            PriorDev prior) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = 32 * W;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = 2 * S - 1, ld = S | 1;
    double *D = reinterpret_cast<double *>(smem_raw);  // [S][ld] dense: slots 0..n-1 live
    double *rsH = D + (size_t)S * ld;                  // [2][S] row sums (hi) by slot, double-buffered
    double *rsL = rsH + 2 * S;                         // [2][S] low parts
    double *red = rsL + 2 * S;                         // [3 W] per-warp (m1, m2, dmin)
    double *res = red + 3 * W;                         // [2] soft index, guard
    double *cand_s = res + 2;                          // [CAND_CAP] candidates of the multi-candidate joins
    double *cand_f = cand_s + CAND_CAP;                // [CAND_CAP]
    int *sidb = reinterpret_cast<int *>(cand_f + CAND_CAP);  // [2][S] slot -> node id, double-buffered
    int *slot_of = sidb + 2 * S;                             // [N] node id -> slot or -1
    unsigned *redk = reinterpret_cast<unsigned *>(slot_of + N);  // [W]
    int *cand_n = reinterpret_cast<int *>(redk + W);
    int *todo = reinterpret_cast<int *>(ws + (size_t)b * nj_ws_doubles(S));
    const double *P = pdm + (size_t)b * S * S;
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *bl = blens + (size_t)b * (2 * S - 2);
    int32_t *fl = flags + (size_t)b * (S - 1);
    auto sync = [&]() { if (W == 1) __syncwarp(); else __syncthreads(); };

    if (tid == 0) *todo = 0;
    for (int i = tid; i < S * S; i += T) {
        const int r = i / S, c = i - r * S;
        D[r * ld + c] = (r == c) ? 0.0 : P[i];  // dists.fill_diagonal_(0.0), peeler.py:178
    }
    for (int s = tid; s < S; s += T) sidb[s] = s;
    for (int i = tid; i < N; i += T) slot_of[i] = i < S ? i : -1;
    sync();
    int asym = 0;
    for (int i = tid; i < S * S; i += T) {
        const int r = i / S, c = i - r * S;
        if (r < c && D[r * ld + c] != D[c * ld + r]) asym = 1;
    }
    if (__syncthreads_or(asym) || !(1.0 / tau > 750.0)) {
        if (tid == 0) *todo = 1;
        return;
    }
    for (int r = warp; r < S; r += W) {  // row sums (peeler.py:240): one warp per row, double-double
        double hi = 0.0, lo = 0.0;
        for (int c = lane; c < S; c += 32)
            if (c != r) dd_add(hi, lo, D[r * ld + c]);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double oh = __shfl_xor_sync(0xffffffffu, hi, m), ol = __shfl_xor_sync(0xffffffffu, lo, m);
            dd_add_dd(hi, lo, oh, ol);
        }
        if (lane == 0) { rsH[r] = hi; rsL[r] = lo; }
    }
    sync();
    const double Wd = 750.0 * tau;
    const double Wd_sure = Wd * (1.0 + 1e-9);  // x > Wd_sure  =>  x / tau > 750 whatever the rounding
    int cur = 0;
    int pend_parent = -1, pend_keep = 0, pend_left = 0, pend_right = 0, pend_moved = -1, pend_freed = 0;  // tid 0 only
    for (int n = S; n > 1; --n) {
        const double *rs = rsH + cur * S;
        const int *sid = sidb + cur * S;
        if (tid == 0 && pend_parent >= 0) {  // the previous join's id -> slot changes (read behind the barriers below)
            slot_of[pend_left] = -1;
            slot_of[pend_right] = -1;
            slot_of[pend_parent] = pend_keep;
            if (pend_moved >= 0) slot_of[pend_moved] = pend_freed;
        }
        const double nm2 = (double)(n - 2);
        // ---- smallest and second smallest Q over the live pairs (compute_Q, peeler.py:223-248): four rows
        // per trip share the column operand and give the loads independent work
        double m1 = DBL_MAX, m2 = DBL_MAX, dmin = 0.0;  // dmin: the distance behind m1 (= d(l, r) of the join)
        unsigned key = 0x00010000u;
        for (int r0 = 1 + warp; r0 < n; r0 += 4 * W) {
            double rsr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) rsr[j] = (r0 + j * W < n) ? rs[r0 + j * W] : 0.0;
            const int rtop = min(r0 + 3 * W, n - 1);
            for (int c = lane; c < rtop; c += 32) {
                const double rsc = rs[c];
                double dv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rj = r0 + j * W;
                    dv[j] = (rj < n && c < rj) ? D[rj * ld + c] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rj = r0 + j * W;
                    const double dij = dv[j];
                    const double qa = (nm2 * dij - rsr[j]) - rsc;
                    const double qb = (nm2 * dij - rsc) - rsr[j];
                    const double q = 0.5 * (qa + qb);
                    if ((rj < n && c < rj) && q < m2) {  // rare once the two running minima have settled
                        if (q < m1) { m2 = m1; m1 = q; dmin = dij; key = ((unsigned)rj << 16) | (unsigned)c; }
                        else m2 = q;
                    }
                }
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double o1 = __shfl_xor_sync(0xffffffffu, m1, m), o2 = __shfl_xor_sync(0xffffffffu, m2, m);
            const double od = __shfl_xor_sync(0xffffffffu, dmin, m);
            const unsigned ok = __shfl_xor_sync(0xffffffffu, key, m);
            const bool lt = o1 < m1;  // an exact tie leaves m2 == m1: not the one-candidate regime
            const double lo1 = lt ? o1 : m1, hi1 = lt ? m1 : o1, oth = lt ? o2 : m2;
            m2 = oth < hi1 ? oth : hi1;
            m1 = lo1;
            dmin = lt ? od : dmin;
            key = lt ? ok : key;
        }
        if (W > 1) {
            if (lane == 0) { red[3 * warp] = m1; red[3 * warp + 1] = m2; red[3 * warp + 2] = dmin; redk[warp] = key; }
            __syncthreads();
            // every warp folds the W per-warp results for itself: lane w takes entry w, then the same butterfly
            m1 = lane < W ? red[3 * lane] : DBL_MAX;
            m2 = lane < W ? red[3 * lane + 1] : DBL_MAX;
            dmin = lane < W ? red[3 * lane + 2] : 0.0;
            key = lane < W ? redk[lane] : 0x00010000u;
#pragma unroll
            for (int m = W / 2; m >= 1; m >>= 1) {
                const double o1 = __shfl_xor_sync(0xffffffffu, m1, m), o2 = __shfl_xor_sync(0xffffffffu, m2, m);
                const double od = __shfl_xor_sync(0xffffffffu, dmin, m);
                const unsigned ok = __shfl_xor_sync(0xffffffffu, key, m);
                const bool lt = o1 < m1;
                const double lo1 = lt ? o1 : m1, hi1 = lt ? m1 : o1, oth = lt ? o2 : m2;
                m2 = oth < hi1 ? oth : hi1;
                m1 = lo1;
                dmin = lt ? od : dmin;
                key = lt ? ok : key;
            }
            // (on an exact tie of the minimum the lanes may hold different keys, but then m2 == m1 everywhere and the
            // key is not used; otherwise every lane of the group holds the same result: take lane 0's)
            m1 = __shfl_sync(0xffffffffu, m1, 0); m2 = __shfl_sync(0xffffffffu, m2, 0);
            dmin = __shfl_sync(0xffffffffu, dmin, 0); key = __shfl_sync(0xffffffffu, key, 0);
        }
        // m = min(min Q, 0) (the strict upper triangle of the padded matrix holds zeros, torch.tril).  The
        // usual case is soft_argmin_fast's one-candidate branch: m is attained by exactly one live pair and
        // nothing else (live, zero or 1e9 entry) is within 750 tau of it.
        int fbits = 0;
        int left, right, sl, sr;
        double d_lr;
        if (m1 < 0.0 && m2 > m1 + Wd && -m1 > Wd_sure && LARGE - m1 > Wd_sure) {
            const int sa = (int)(key >> 16), sb = (int)(key & 0xffffu);
            const int ia = sid[sa], ib = sid[sb];
            right = max(ia, ib);  // flat = row * N + col with row > col (tril)
            left = min(ia, ib);
            sl = ia < ib ? sa : sb;
            sr = ia < ib ? sb : sa;
            d_lr = dmin;  // == D[sl][sr] (the matrix is symmetric bit for bit); taken from the scan because the
                          // pass below overwrites that entry while slower threads may still be here
        } else {
            // several entries inside the window, or a minimum too close to the zeros: gather the candidates and
            // evaluate the two softmaxes as soft_argmin_fast does
            const double mm = fmin(m1, 0.0);
            if (tid == 0) *cand_n = 0;
            sync();
            for (int r = 1 + warp; r < n; r += W) {
                const double rsr = rs[r];
                const double *rowp = D + r * ld;
                for (int c = lane; c < r; c += 32) {
                    const double dij = rowp[c], rsc = rs[c];
                    const double qa = (nm2 * dij - rsr) - rsc;
                    const double qb = (nm2 * dij - rsc) - rsr;
                    const double q = 0.5 * (qa + qb);
                    if (q <= mm + Wd) {
                        const int slot = atomicAdd(cand_n, 1);
                        const int ia = sid[r], ib = sid[c];
                        if (slot < CAND_CAP) { cand_s[slot] = q; cand_f[slot] = (double)max(ia, ib) * (double)N + (double)min(ia, ib); }
                    }
                }
            }
            sync();
            const int nc = *cand_n;
            if (nc < 1 || nc > CAND_CAP) {  // the general multi-pass evaluation: not here
                if (tid == 0) *todo = 1;
                return;
            }
            if (tid == 0) soft_eval_candidates(cand_s, cand_f, nc, mm, tau, (double)N * (double)N, false, res[0], res[1]);
            sync();
            const double idx = res[0];
            if (res[1] == 0.0) fbits = 4;
            double rowf, colf;
            unravel(idx, (double)N, rowf, colf);
            right = (int)rint(rowf);
            left = (int)rint(colf);
            const bool bad = !(idx == idx) || right < 0 || right >= N || left < 0 || left >= N || right == left;
            sl = bad ? -1 : slot_of[left];
            sr = bad ? -1 : slot_of[right];
            if (sl < 0 || sr < 0) {  // invalid selection: flagged by the general kernel
                if (tid == 0) *todo = 1;
                return;
            }
            d_lr = D[sl * ld + sr];
            sync();  // (everyone holds d_lr before the pass below overwrites that entry)
        }
        const int keep = min(sl, sr), freed = max(sl, sr), last = n - 1;
        const int parent = S + (S - n);
        const int nxt = cur ^ 1;
        // ---- one pass: distances of the new node (get_new_dist_soft, peeler.py:251-265) into slot `keep`, the
        // last live slot moved into `freed` (keep < freed <= last), running row sums.  Thread x reads column
        // x of rows sl, sr, last and writes column x of rows keep, freed and row x: no two threads meet.
        for (int x = tid; x < n; x += T) {
            if (x == sl || x == sr) continue;
            const double dl = D[sl * ld + x], dr = D[sr * ld + x];
            const double u = 0.5 * ((dl + dr) - d_lr);
            const int dx = x == last ? freed : x;
            if (x != last && freed != last) {
                const double v = D[last * ld + x];
                D[freed * ld + x] = v;
                D[x * ld + freed] = v;
            }
            D[keep * ld + dx] = u;
            D[dx * ld + keep] = u;
            double hi = rs[x], lo = rsL[cur * S + x];
            dd_add(hi, lo, -dl);
            dd_add(hi, lo, -dr);
            dd_add(hi, lo, u);
            rsH[nxt * S + dx] = hi; rsL[nxt * S + dx] = lo;
            sidb[nxt * S + dx] = sid[x];
        }
        if (tid == 0) {  // branch lengths (peeler.py:199-213) and the peel row
            const double nact = fmax((double)n, 3.0);
            double d_pl = 0.5 * d_lr + (rs[sl] - rs[sr]) / (2.0 * (nact - 2.0));
            if (!(d_pl > DBL_EPSILON)) { fbits |= (d_pl == DBL_EPSILON) ? 0 : 1; d_pl = DBL_EPSILON; }
            double d_pr = d_lr - d_pl;
            if (!(d_pr > DBL_EPSILON)) { fbits |= (d_pr == DBL_EPSILON) ? 0 : 2; d_pr = DBL_EPSILON; }
            pl[(S - n) * 3 + 0] = left;
            pl[(S - n) * 3 + 1] = right;
            pl[(S - n) * 3 + 2] = parent;
            bl[left] = d_pl;
            bl[right] = d_pr;
            fl[S - n] = fbits;
            pend_parent = parent; pend_keep = keep; pend_left = left; pend_right = right;
            pend_moved = freed != last ? sid[last] : -1;
            pend_freed = freed;
        }
        if (tid == (W > 1 ? 32 : 1)) {  // row sum of the new node: (r_l + r_r - n d_lr) / 2 in double-double
            double hi = rs[sl], lo = rsL[cur * S + sl];
            dd_add_dd(hi, lo, rs[sr], rsL[cur * S + sr]);
            const double nn = (double)n, p = nn * d_lr, pe = fma(nn, d_lr, -p);  // n d_lr = p + pe exactly
            dd_add_dd(hi, lo, -p, -pe);
            rsH[nxt * S + keep] = 0.5 * hi; rsL[nxt * S + keep] = 0.5 * lo;
            sidb[nxt * S + keep] = parent;
        }
        sync();
        cur = nxt;
    }
    if (prior.ln_prior != nullptr && warp == 0) prior_warp(bl, S, prior, b, lane);  // epilogue: bl written by lane 0
}

// ------------------------------------------------------------------------------------------ lean hard
// Hard NJ for trees whose lower triangle fits in shared memory (<= NJ_HARD_LEAN_MAX taxa): the same construction
// as nj_soft_lean_kernel -- dense slots, the matrix resident (triangular: D(i,j), i > j, at i(i-1)/2 + j), per-slot
// scalars double-buffered, three barriers per join -- with Cpeeler.nj_np's arithmetic kept to the bit:
// explicitly rounded q = ((n-2) d - x_a) - x_b with a the pool-first node, FIRST strict minimum in pool order
// (pool order = ascending node id, so ties go to the smaller id pair), running row sums updated in the
// reference's operation order, the new node's row sum accumulated sequentially in pool order (`ord` keeps the
// slots sorted by node id).  Asymmetric input or an all-NaN matrix: todo = 1, the general kernel redoes the tree.
//
// CL > 1: one tree per thread-block CLUSTER of CL CTAs (a shard of a sharded MCMC run leaves most SMs without a tree:
// 8 chains on 148 SMs).  The scan -- the issue-bound part of a join -- is split over the CL x W warps of the cluster;
// every warp writes its (q, d, id pair, slot pair) straight into the exchange buffers of all CL CTAs through
// distributed shared memory (mapa + st.shared::cluster), ONE cluster barrier (arrive.release / wait.acquire) replaces
// the block barrier, and every warp folds the CL x W results for itself.  Everything else of a join (new distances,
// slot move, row sums, the sequential sum) is done redundantly by every CTA on its own copy of the matrix, so no
// other exchange is needed and the arithmetic -- hence the result -- is the one-CTA kernel's bit for bit.
constexpr int NJ_HARD_LEAN_MAX = 224;
__host__ __device__ inline size_t nj_hard_lean_smem_bytes(int S, int W, int CL = 1) {
    const size_t nx = CL > 1 ? 2 * (size_t)CL * W : (size_t)W;  // exchange buffers are double-buffered across joins
    return ((size_t)S * (S - 1) / 2 + 1 + 3 * (size_t)S + 2 * nx + 2) * sizeof(double) +
           (6 * (size_t)S + 2 * nx + 4) * sizeof(int);
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_map(const void *p, unsigned rank) {  // shared::cluster address of p in CTA `rank`
    unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_st(unsigned addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cluster_st(unsigned addr, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int tri_at(int i, int j) { return i > j ? i * (i - 1) / 2 + j : j * (j - 1) / 2 + i; }

template <int W, int CL>
__global__ void __launch_bounds__(32 * W) nj_hard_lean_kernel(const double *__restrict__ pdm, int S, double *__restrict__ ws,
                                                            int64_t *__restrict__ peel, double *__restrict__ blens,
                                                            PriorDev prior) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = 32 * W;
    constexpr int TW = W * CL;                           // warps per tree
    constexpr int NX = CL > 1 ? 2 * TW : W;              // per-warp results (CL > 1: all warps of the cluster, two joins)
    const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int mw = warp * CL + rank;                     // this warp among the TW warps of the tree (rows interleaved)
    double *tri = reinterpret_cast<double *>(smem_raw);  // [S (S-1) / 2]
    double *xsb = tri + (((size_t)S * (S - 1) / 2 + 1) & ~(size_t)1);  // [2][S] running row sums by slot, double-buffered (16-byte aligned)
    double *newd = xsb + 2 * S;                          // [S] new distances in pool order
    double *red = newd + S;                              // [2 NX] per-warp (q, d)
    int *sidb = reinterpret_cast<int *>(red + 2 * NX + 2);   // [2][S] slot -> node id
    int *posb = sidb + 2 * S;                            // [2][S] slot -> pool position
    int *ordb = posb + 2 * S;                            // [2][S] pool position -> slot
    unsigned *redk = reinterpret_cast<unsigned *>(ordb + 2 * S);  // [2 NX] (id pair, slot pair)
    int *todo = reinterpret_cast<int *>(ws + (size_t)b * nj_ws_doubles(S));
    const double *P = pdm + (size_t)b * S * S;
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *bl = blens + (size_t)b * (2 * S - 2);

    if (tid == 0 && rank == 0) *todo = 0;
    int asym = 0;
    for (int i = tid; i < S * S; i += T) {
        const int r = i / S, c = i - r * S;
        if (r > c) {
            const double v = P[i];
            tri[r * (r - 1) / 2 + c] = v;
            if (v != P[(size_t)c * S + r]) asym = 1;
        }
    }
    for (int s = tid; s < S; s += T) {
        sidb[s] = s; posb[s] = s; ordb[s] = s;
        // Cpeeler.pyx:46-54: sequential sum over the other pool members, in pool order
        double acc = 0.0;
        const double *row = P + (size_t)s * S;
        for (int j = 0; j < S; ++j)
            if (j != s) acc = __dadd_rn(acc, row[j]);
        xsb[s] = acc;
    }
    if (__syncthreads_or(asym)) {  // (every CTA of a cluster sees the same matrix: they leave together)
        if (tid == 0 && rank == 0) *todo = 1;
        return;
    }
    if (CL > 1) cluster_barrier();  // every CTA of the cluster is resident before the first remote store
    int cur = 0;
#ifdef NJ_PHASE_CLOCKS  // debugging aid: cycles per phase of a join, summed over the joins, printed by one thread
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = clock64();
#define NJ_PHASE(i) do { const long long t__ = clock64(); ph[i] += t__ - tc; tc = t__; } while (0)
#else
#define NJ_PHASE(i) do { } while (0)
#endif
    // The new node's row sum is a SEQUENTIAL sum in pool order (Cpeeler.pyx:92: one dependent fp64 add per survivor,
    // ~18 cycles each: 28 % of a join at 200 taxa when it stands alone between two barriers).  It is only needed by
    // the NEXT join, and there only by the pairs that involve the new node.  So it runs on lane 0 of warp 0 WHILE
    // the other warps scan the next join's pairs (phase A); the new node's slot `ks` carries the row sum -DBL_MAX
    // meanwhile, which makes every q that involves it +huge, i.e. never the minimum; after the barrier all threads
    // visit the n - 1 pairs of the new node with its real row sum (phase B).  Same candidates, same arithmetic.
    constexpr int WS = W - 1, TWS = WS * CL;             // scanning warps per CTA / per tree (phase A)
    const int mws = (warp - 1) * CL + rank;
    int ks = -1, pend = 0;                               // slot of the node the previous join created; its survivors
    for (int n = S; n > 1; --n) {
        double *xs = xsb + cur * S;
        const int *sid = sidb + cur * S, *pos = posb + cur * S, *ord = ordb + cur * S;
        const int nxt = cur ^ 1;
        const double nm2 = (double)n - 2.0;
        // ---- first strict minimum of q in pool order (Cpeeler.pyx:58-66); RU rows per trip share the column
        // (a cluster leaves a warp one or two rows per join: there a trip over four rows is mostly masked lanes)
        constexpr int RU = CL > 1 ? 1 : 4;
        double bq = DBL_MAX, bd = 0.0;
        unsigned bkey = ~0u, bslot = 0u;
        if (warp == 0) {
            if (lane == 0 && ks >= 0) {
                // the previous join's new node: sequential in pool order (Cpeeler.pyx:92), loads running ahead of the adds
                double acc = 0.0;
                const int m = pend;
                int p = 0;
#ifdef NJ_TIMING_NO_SEQSUM  // timing experiment only (wrong row sums): what the dependent adds cost
                p = m > 8 ? ((m - 8) & ~1) : 0;
#endif
#pragma unroll 2
                for (; p + 8 <= m; p += 8) {
                    const double2 u0 = *reinterpret_cast<const double2 *>(newd + p);
                    const double2 u1 = *reinterpret_cast<const double2 *>(newd + p + 2);
                    const double2 u2 = *reinterpret_cast<const double2 *>(newd + p + 4);
                    const double2 u3 = *reinterpret_cast<const double2 *>(newd + p + 6);
                    acc = __dadd_rn(acc, u0.x); acc = __dadd_rn(acc, u0.y);
                    acc = __dadd_rn(acc, u1.x); acc = __dadd_rn(acc, u1.y);
                    acc = __dadd_rn(acc, u2.x); acc = __dadd_rn(acc, u2.y);
                    acc = __dadd_rn(acc, u3.x); acc = __dadd_rn(acc, u3.y);
                }
                for (; p < m; ++p) acc = __dadd_rn(acc, newd[p]);
                xs[ks] = acc;
            }
        } else
        for (int r0 = 1 + mws; r0 < n; r0 += RU * TWS) {
            double xr[RU];
            int ir[RU];
#pragma unroll
            for (int j = 0; j < RU; ++j) {
                const int rj = r0 + j * TWS;
                xr[j] = rj < n ? xs[rj] : 0.0;
                ir[j] = rj < n ? sid[rj] : 0;
            }
            const int rtop = min(r0 + (RU - 1) * TWS, n - 1);
            for (int c = lane; c < rtop; c += 32) {
                const double xc = xs[c];
                const int ic = sid[c];
                double dv[RU];
#pragma unroll
                for (int j = 0; j < RU; ++j) {
                    const int rj = r0 + j * TWS;
                    dv[j] = (rj < n && c < rj) ? tri[rj * (rj - 1) / 2 + c] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < RU; ++j) {
                    const int rj = r0 + j * TWS;
                    const bool rfirst = ir[j] < ic;
                    const double xa = rfirst ? xr[j] : xc, xb = rfirst ? xc : xr[j];
                    const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, dv[j]), xa), xb);
                    if ((rj < n && c < rj) && q <= bq) {  // rare once the running minimum has settled
                        const unsigned key = rfirst ? (((unsigned)ir[j] << 16) | (unsigned)ic) : (((unsigned)ic << 16) | (unsigned)ir[j]);
                        if (q < bq || key < bkey) {
                            bq = q;
                            bd = dv[j];
                            bkey = key;
                            bslot = ((unsigned)rj << 16) | (unsigned)c;
                        }
                    }
                }
            }
        }
        NJ_PHASE(0);
        if (ks >= 0) {  // ---- phase B: the pairs of the new node, once its row sum is there
            __syncthreads();
            const double xk = xs[ks];
            const int ik = sid[ks];
            for (int c = rank * T + tid; c < n; c += T * CL) {
                if (c == ks) continue;
                const double xc = xs[c];
                const int ic = sid[c];
                const double dkc = tri[tri_at(ks, c)];
                const bool kfirst = ik < ic;
                const double xa = kfirst ? xk : xc, xb = kfirst ? xc : xk;
                const double q = __dsub_rn(__dsub_rn(__dmul_rn(nm2, dkc), xa), xb);
                if (q <= bq) {
                    const unsigned key = kfirst ? (((unsigned)ik << 16) | (unsigned)ic) : (((unsigned)ic << 16) | (unsigned)ik);
                    if (q < bq || key < bkey) {
                        bq = q;
                        bd = dkc;
                        bkey = key;
                        bslot = ((unsigned)max(ks, c) << 16) | (unsigned)min(ks, c);
                    }
                }
            }
        }
        NJ_PHASE(6);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double oq = __shfl_xor_sync(0xffffffffu, bq, m), od = __shfl_xor_sync(0xffffffffu, bd, m);
            const unsigned ok = __shfl_xor_sync(0xffffffffu, bkey, m), os = __shfl_xor_sync(0xffffffffu, bslot, m);
            const bool take = oq < bq || (oq == bq && ok < bkey);
            bq = take ? oq : bq; bd = take ? od : bd; bkey = take ? ok : bkey; bslot = take ? os : bslot;
        }
        NJ_PHASE(1);
        if (CL == 1) {
            if (lane == 0) { red[2 * warp] = bq; red[2 * warp + 1] = bd; redk[2 * warp] = bkey; redk[2 * warp + 1] = bslot; }
            __syncthreads();
            // every warp folds the W per-warp results for itself: lane w takes entry w, then the same butterfly
            bq = lane < W ? red[2 * lane] : DBL_MAX;
            bd = lane < W ? red[2 * lane + 1] : 0.0;
            bkey = lane < W ? redk[2 * lane] : ~0u;
            bslot = lane < W ? redk[2 * lane + 1] : 0u;
        } else {
            // lane l hands this warp's result to CTA l of the cluster (slot `mw` of the buffer of this join's parity);
            // the release/acquire pair of the cluster barrier orders the remote stores before the reads below
            const int px = (n & 1) * TW;
            if (lane < CL) {
                const unsigned aq = cluster_map(red + 2 * (px + mw), (unsigned)lane);
                const unsigned ak = cluster_map(redk + 2 * (px + mw), (unsigned)lane);
                cluster_st(aq, bq); cluster_st(aq + 8u, bd);
                cluster_st(ak, bkey); cluster_st(ak + 4u, bslot);
            }
            cluster_barrier();
            NJ_PHASE(2);
            bq = DBL_MAX; bd = 0.0; bkey = ~0u; bslot = 0u;
#pragma unroll
            for (int e = lane; e < TW; e += 32) {
                const double oq = red[2 * (px + e)], od = red[2 * (px + e) + 1];
                const unsigned ok = redk[2 * (px + e)], os = redk[2 * (px + e) + 1];
                const bool take = oq < bq || (oq == bq && ok < bkey);
                bq = take ? oq : bq; bd = take ? od : bd; bkey = take ? ok : bkey; bslot = take ? os : bslot;
            }
        }
#pragma unroll
        for (int m = (TW < 32 ? TW : 32) / 2; m >= 1; m >>= 1) {
            const double oq = __shfl_xor_sync(0xffffffffu, bq, m), od = __shfl_xor_sync(0xffffffffu, bd, m);
            const unsigned ok = __shfl_xor_sync(0xffffffffu, bkey, m), os = __shfl_xor_sync(0xffffffffu, bslot, m);
            const bool take = oq < bq || (oq == bq && ok < bkey);
            bq = take ? oq : bq; bd = take ? od : bd; bkey = take ? ok : bkey; bslot = take ? os : bslot;
        }
        bq = __shfl_sync(0xffffffffu, bq, 0); bd = __shfl_sync(0xffffffffu, bd, 0);
        bkey = __shfl_sync(0xffffffffu, bkey, 0); bslot = __shfl_sync(0xffffffffu, bslot, 0);
        NJ_PHASE(3);
        if (bkey == ~0u) {  // all-NaN distances: the general kernel's memory-safe fallback decides
            if (tid == 0 && rank == 0) *todo = 1;
            return;
        }
        const int n1 = (int)(bkey >> 16), n2 = (int)(bkey & 0xffffu);  // node ids; n1 comes first in the pool
        const int s_hi = (int)(bslot >> 16), s_lo = (int)(bslot & 0xffffu);
        const int sa = sid[s_hi] == n1 ? s_hi : s_lo, sb = sid[s_hi] == n1 ? s_lo : s_hi;
        const int pa = pos[sa], pb = pos[sb];
        const int keep = s_lo, freed = s_hi, last = n - 1;
        const int parent = S + (S - n);
        const double d12 = bd;  // == D(sa, sb), carried by the reduction: the pass below overwrites that entry
        // ---- one pass over the pool: distances of the new node (into slot `keep`), the last live slot moved into
        // `freed`, running row sums (Cpeeler.pyx:81-95), the survivors' new pool positions.  Thread p only touches
        // column sx = ord[p] of rows sa, sb, last, keep, freed: no two threads meet.
        for (int p = tid; p < n; p += T) {
            if (p == pa || p == pb) continue;
            const int sx = ord[p];
            const double dsa = tri[tri_at(sx, sa)], dsb = tri[tri_at(sx, sb)];
            const double v1 = __dadd_rn(__dadd_rn(0.0, dsa), dsb);
            const double dist = __dmul_rn(0.5, __dsub_rn(v1, d12));
            double x = __dadd_rn(xs[sx], dist);
            x = __dsub_rn(x, dsa);
            x = __dsub_rn(x, dsb);
            const int dx = sx == last ? freed : sx;
            if (sx != last && freed != last) tri[tri_at(freed, sx)] = tri[tri_at(last, sx)];
            tri[tri_at(keep, dx)] = dist;
            const int pc = p - (p > pa) - (p > pb);
            xsb[nxt * S + dx] = x;
            sidb[nxt * S + dx] = sid[sx];
            posb[nxt * S + dx] = pc;
            ordb[nxt * S + pc] = dx;
            newd[pc] = dist;  // dense, in pool order: summed sequentially below
        }
        NJ_PHASE(4);
        if (tid == 32 % T) {
            // branch lengths (Cpeeler.pyx:98-112) with the joined nodes' row sums as they stand
            double df, dg;
            if (n > 2) {
                const double v1 = __dmul_rn(0.5, d12);
                const double v4 = __dmul_rn(__ddiv_rn(1.0, __dmul_rn(2.0, nm2)), __dsub_rn(xs[sa], xs[sb]));
                df = __dadd_rn(v1, v4);
                dg = __dsub_rn(d12, df);
            } else {
                df = __ddiv_rn(d12, 2.0);
                dg = df;
            }
            if (rank == 0) {
                bl[n1] = fmax(df, DBL_EPSILON);  // Cpeeler.pyx:124
                bl[n2] = fmax(dg, DBL_EPSILON);
                pl[(S - n) * 3 + 0] = n1;
                pl[(S - n) * 3 + 1] = n2;
                pl[(S - n) * 3 + 2] = parent;
            }
            sidb[nxt * S + keep] = parent;
            posb[nxt * S + keep] = n - 2;
            ordb[nxt * S + n - 2] = keep;
            xsb[nxt * S + keep] = -DBL_MAX;  // until the next join's phase A has summed the new row
        }
        __syncthreads();
        NJ_PHASE(5);
        ks = keep;
        pend = n - 2;
        NJ_PHASE(7);
        cur = nxt;
    }
#ifdef NJ_PHASE_CLOCKS
    if (tid == 0 && blockIdx.x == 0)
        printf("nj_hard_lean W=%d CL=%d S=%d cycles: scan %lld | butterfly %lld | exchange+barrier %lld | fold %lld | update %lld | "
               "blens+sync %lld | phase B (incl. wait for the row sum) %lld | - %lld\n", W, CL, S, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
#endif
#undef NJ_PHASE
    if (prior.ln_prior != nullptr && warp == 0 && rank == 0) prior_warp(bl, S, prior, b, lane);
}

// reverse replay: adjoint of blens w.r.t. the input distance matrix
__global__ void __launch_bounds__(1024) nj_soft_bwd_kernel(const int64_t *__restrict__ peel,
                                                           const int32_t *__restrict__ flags,
                                                           const double *__restrict__ gblens, int S,
                                                           double *__restrict__ ws, double *__restrict__ gpdm) {
    extern __shared__ unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int N = 2 * S - 1;
    double *Ab = ws + (size_t)b * nj_ws_doubles(S) + (size_t)S * S;
    double *sh = reinterpret_cast<double *>(smem_raw);  // [32]
    int *st_of = reinterpret_cast<int *>(sh + 32);      // [N] node id -> storage slot
    int *active = st_of + N;                            // [S] storage slot active?
    const int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    const int32_t *fl = flags + (size_t)b * (S - 1);
    const double *gb = gblens + (size_t)b * (2 * S - 2);
    double *out = gpdm + (size_t)b * S * S;

    __shared__ int t_last;
    for (size_t i = tid; i < (size_t)S * S; i += T) Ab[i] = 0.0;
    for (int s = tid; s < S; s += T) { st_of[s] = s; active[s] = 0; }
    if (tid == 0) t_last = 0;
    __syncthreads();
    // the last join whose two branch lengths carry a cotangent: later joins only see zeros (one-hot
    // cotangents of the batched Jacobian skip half of the replay on average)
    for (int t = tid; t < S - 1; t += T)
        if (gb[(int)pl[t * 3 + 0]] != 0.0 || gb[(int)pl[t * 3 + 1]] != 0.0) atomicMax(&t_last, t);
    __syncthreads();
    const int t_start = t_last;
    if (tid == 0) {
        for (int t = 0; t < S - 1; ++t) st_of[(int)pl[t * 3 + 2]] = st_of[(int)pl[t * 3 + 0]];
        active[st_of[(int)pl[(S - 2) * 3 + 2]]] = 1;
        for (int t = S - 2; t > t_start; --t) active[st_of[(int)pl[t * 3 + 1]]] = 1;  // un-joins without adjoints
    }
    __syncthreads();
    for (int t = t_start; t >= 0; --t) {
        const int left = (int)pl[t * 3 + 0], right = (int)pl[t * 3 + 1];
        const int sl = st_of[left], sr = st_of[right];
        const int n = S - t;
        const int f = fl[t];
        const double nact = fmax((double)n, 3.0);
        // un-join: slot sl (currently the parent) becomes `left`, slot sr becomes active as `right`
        double g_pl = gb[left], g_pr = gb[right];
        double g_lr = 0.0, g_rl = 0.0, g_rr = 0.0;
        if (!(f & 2)) { g_lr += g_pr; g_pl -= g_pr; }
        if (!(f & 1)) {
            g_lr += 0.5 * g_pl;
            g_rl = g_pl / (2.0 * (nact - 2.0));
            g_rr = -g_rl;
        }
        // one pass: the adjoint of the parent's distances is summed (for d(l, r)) while it is handed
        // to the two children; the reduction's barriers also publish the updates
        double usum = 0.0;
        for (int sx = tid; sx < S; sx += T) {
            if (!active[sx] || sx == sl) continue;
            double ub = Ab[(size_t)sl * S + sx] + Ab[(size_t)sx * S + sl];
            usum += ub;
            Ab[(size_t)sl * S + sx] = 0.5 * ub + g_rl;
            Ab[(size_t)sr * S + sx] = 0.5 * ub + g_rr;
            Ab[(size_t)sx * S + sl] = 0.0;
            Ab[(size_t)sx * S + sr] = 0.0;
        }
        usum = block_reduce(usum, sh, 0);
        if (tid == 0) {
            Ab[(size_t)sl * S + sr] = -0.5 * usum + g_lr + g_rl;
            Ab[(size_t)sr * S + sl] = g_rr;
            active[sr] = 1;
        }
        __syncthreads();
    }
    for (size_t i = tid; i < (size_t)S * S; i += T) {
        int r = (int)(i / S), c = (int)(i - (size_t)r * S);
        out[i] = (r == c) ? 0.0 : 0.5 * (Ab[(size_t)r * S + c] + Ab[(size_t)c * S + r]);
    }
}

// The same vector-Jacobian product without a replay, for trees whose per-join weight vectors fit in shared memory
// (<= NJ_BWD_LEAN_MAX taxa).  With the selections fixed every NJ distance is a bilinear form of the tip distance
// matrix D (see blens_jac_kernel, csrc/pdm.cu): d(u, x) = a_u^T D a_x - c_u - c_x, a_u = (a_l + a_r) / 2,
// c_u = F_u / 2 with F_t = a_l^T D a_r the form of the join that made u.  A join contributes
//   g_lr d_lr + kappa_t (r_l - r_r),  d_lr = F_t - c_l - c_r,  r_l - r_r = (a_l - a_r)^T D sigma_t - (n-2)(c_l - c_r)
// to sum_e g_e blens_e (g_lr, kappa_t from the two cotangents and the clamp flags exactly as in the replay), so
//   d / d D = sum_t [ A_t a_l (x) a_r + kappa_t (a_l - a_r) (x) sigma_t ]  =  a (S x 2(S-1)) (2(S-1) x S) product,
// A_t = g_lr(t) plus what the parent join of node S+t puts on c.  Prologue: one thread per taxon fills its column of
// a_l, a_r, sigma_t (a column only depends on itself), one thread forms the scalars; then every thread
// accumulates a 2 x 2 tile of the product.  No barrier per join, no reduction (the replay needs four barriers and
// a block-wide sum per un-join: 78 us for 128 x 64 taxa, latency-bound).
constexpr int NJ_BWD_LEAN_MAX = 80;
__host__ __device__ inline size_t nj_bwd_lean_smem_bytes(int S) {
    return (3 * (size_t)(S - 1) * S + (size_t)S * S + 4 * (size_t)S) * sizeof(double) + 3 * (size_t)S * sizeof(int);
}

__global__ void __launch_bounds__(1024) nj_soft_bwd_lean_kernel(const int64_t *__restrict__ peel,
                                                                const int32_t *__restrict__ flags,
                                                                const double *__restrict__ gblens, int S,
                                                                double *__restrict__ gpdm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int nJ = S - 1;
    double *U = reinterpret_cast<double *>(smem_raw);  // [S-1][S] a_l of join t
    double *V = U + (size_t)nJ * S;                    // [S-1][S] a_r
    double *Sg = V + (size_t)nJ * S;                   // [S-1][S] sigma_t: sum of a_x over the other live nodes
    double *G = Sg + (size_t)nJ * S;                   // [S][S]
    double *Ac = G + (size_t)S * S;                    // [S] coefficient of F_t
    double *kap = Ac + S;                              // [S] coefficient of r_l - r_r
    double *gbs = kap + S;                             // [2S] the cotangents, staged (one thread walks them)
    int *cl = reinterpret_cast<int *>(gbs + 2 * S);    // [S] left child of join t
    int *cr = cl + S;                                  // [S] right child
    int *fls = cr + S;                                 // [S] clamp flags
    const int64_t *pl = peel + (size_t)b * nJ * 3;
    const int32_t *fl = flags + (size_t)b * nJ;
    const double *gb = gblens + (size_t)b * (2 * S - 2);
    double *out = gpdm + (size_t)b * S * S;
    for (int t = tid; t < nJ; t += T) {
        cl[t] = (int)pl[t * 3 + 0];
        cr[t] = (int)pl[t * 3 + 1];
        fls[t] = fl[t];
        Ac[t] = 0.0;
    }
    for (int e = tid; e < 2 * S - 2; e += T) gbs[e] = gb[e];
    __syncthreads();
    if (tid < S) {  // column tid of a_l, a_r, sigma_t, join by join
        const int i = tid;
        double Ti = 1.0;  // sum of a_x[i] over the live nodes
        for (int t = 0; t < nJ; ++t) {
            const int l = cl[t], r = cr[t];
            const double al = l < S ? (l == i ? 1.0 : 0.0) : 0.5 * (U[(size_t)(l - S) * S + i] + V[(size_t)(l - S) * S + i]);
            const double ar = r < S ? (r == i ? 1.0 : 0.0) : 0.5 * (U[(size_t)(r - S) * S + i] + V[(size_t)(r - S) * S + i]);
            U[(size_t)t * S + i] = al;
            V[(size_t)t * S + i] = ar;
            Sg[(size_t)t * S + i] = Ti - al - ar;
            Ti += 0.5 * (al + ar) - al - ar;
        }
    } else if (tid == ((S + 31) & ~31)) {  // one thread of another warp: the per-join scalars (parents first)
        for (int t = nJ - 1; t >= 0; --t) {
            const int l = cl[t], r = cr[t], f = fls[t];
            const int n = S - t;
            const double nact = fmax((double)n, 3.0);
            double g_pl = gbs[l], g_pr = gbs[r], g_lr = 0.0;
            if (!(f & 2)) { g_lr += g_pr; g_pl -= g_pr; }
            double kp = 0.0;
            if (!(f & 1)) { g_lr += 0.5 * g_pl; kp = g_pl / (2.0 * (nact - 2.0)); }
            kap[t] = kp;
            Ac[t] += g_lr;  // (Ac[t] already holds what the parent of node S+t put on c_{S+t} = F_t / 2)
            const double w = kp * (double)(n - 2);
            if (l >= S) Ac[l - S] += 0.5 * (-g_lr - w);
            if (r >= S) Ac[r - S] += 0.5 * (-g_lr + w);
        }
    }
    __syncthreads();
    const int H = (S + 1) / 2;  // 2 x 2 tiles
    for (int w = tid; w < H * H; w += T) {
        const int i0 = 2 * (w / H), j0 = 2 * (w - (w / H) * H);
        const int i1 = min(i0 + 1, S - 1), j1 = min(j0 + 1, S - 1);
        double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
        for (int t = 0; t < nJ; ++t) {
            const double *u = U + (size_t)t * S, *v = V + (size_t)t * S, *sg = Sg + (size_t)t * S;
            const double a = Ac[t], k = kap[t];
            const double p0 = a * u[i0], p1 = a * u[i1], q0 = k * (u[i0] - v[i0]), q1 = k * (u[i1] - v[i1]);
            const double vj0 = v[j0], vj1 = v[j1], s0 = sg[j0], s1 = sg[j1];
            g00 = fma(q0, s0, fma(p0, vj0, g00));
            g01 = fma(q0, s1, fma(p0, vj1, g01));
            g10 = fma(q1, s0, fma(p1, vj0, g10));
            g11 = fma(q1, s1, fma(p1, vj1, g11));
        }
        G[(size_t)i0 * S + j0] = g00;
        if (j0 + 1 < S) G[(size_t)i0 * S + j0 + 1] = g01;
        if (i0 + 1 < S) G[(size_t)(i0 + 1) * S + j0] = g10;
        if (i0 + 1 < S && j0 + 1 < S) G[(size_t)(i0 + 1) * S + j0 + 1] = g11;
    }
    __syncthreads();
    for (int e = tid; e < S * S; e += T) {
        const int i = e / S, j = e - i * S;
        out[e] = (i == j) ? 0.0 : 0.5 * (G[e] + G[(size_t)j * S + i]);
    }
}

// soft.argmin on arbitrary matrices (dodonaphy/soft.py:38-53), one CTA per matrix
__global__ void __launch_bounds__(256) soft_argmin_kernel(const double *__restrict__ Q, int rows, int cols, double tau,
                                                          double *__restrict__ out) {
    __shared__ double sh[32];
    const long long n = (long long)rows * cols;
    SoftCtx c{Q + (size_t)blockIdx.x * n, nullptr, nullptr, nullptr, nullptr, 0, cols, 0, 0.0, n, 0};
    double idx;
    bool ok;
    soft_argmin_block<2>(c, tau, sh, idx, ok);
    if (threadIdx.x == 0) {
        double rowf, colf;
        unravel(idx, (double)cols, rowf, colf);
        out[blockIdx.x * 2 + 0] = rowf;
        out[blockIdx.x * 2 + 1] = colf;
    }
}


// ------------------------------------------------------------------------------------------ geodesic
// The "geodesics" connector: repeatedly join the two live points whose LCA -- the point of their
// geodesic nearest the origin of the Poincare ball -- is deepest.
//   geo_soft_kernel      peeler.make_soft_peel_tips    (dodonaphy/peeler.py:14-70)
//   geo_soft_bwd_kernel  the VJP torch autograd takes through it (locations carry gradient, the
//                        one-hot selections do not, soft.py:56-61)
//   geo_hard_kernel      peeler.make_hard_peel_geodesic (dodonaphy/peeler.py:73-152)
// LCA arithmetic: poincare.hyp_lca (dodonaphy/poincare.py:6-115) -- invert the ball in the circle
// centred at a/|a|^2, mirror the image of the origin in the line through the image of b, invert
// back, halve the hyperbolic distance from the origin.
constexpr int GEO_MAX_D = 16;
constexpr double GEO_MIN_NORM = 1e-15;  // poincare.py:34

__host__ __device__ inline size_t geo_ws_doubles(int S, int D) {
    return 2 * (size_t)S * S + (size_t)(2 * S) * (size_t)(D + 2);
}

__device__ void lca_point(const double *a, const double *b, int D, double *proj) {
    double r[GEO_MAX_D], v[GEO_MAX_D];
    double na2 = 0.0;
    for (int d = 0; d < D; ++d) na2 += a[d] * a[d];
    double rr = 0.0;
    for (int d = 0; d < D; ++d) { r[d] = a[d] / na2; rr += r[d] * r[d]; }
    const double r2 = rr - 1.0;
    double s1 = 0.0;
    for (int d = 0; d < D; ++d) { v[d] = b[d] - r[d]; s1 += v[d] * v[d]; }
    const double c1 = r2 / s1;
    double dot = 0.0, nv = 0.0;
    for (int d = 0; d < D; ++d) { v[d] = c1 * v[d] + r[d]; dot += a[d] * v[d]; nv += v[d] * v[d]; }
    nv = fmax(nv, GEO_MIN_NORM);
    double s2 = 0.0;
    for (int d = 0; d < D; ++d) { v[d] = (2.0 * (dot * v[d] / nv) - a[d]) - r[d]; s2 += v[d] * v[d]; }
    const double c2 = r2 / s2;
    double no = 0.0;
    for (int d = 0; d < D; ++d) { v[d] = c2 * v[d] + r[d]; no += v[d] * v[d]; }
    const double h = 1.0 + sqrt(1.0 - no);
    for (int d = 0; d < D; ++d) proj[d] = v[d] / h;
}

__device__ __forceinline__ double depth_of(const double *p, int D) {  // poincare.hyp_dist_o
    double n2 = 0.0;
    for (int d = 0; d < D; ++d) n2 += p[d] * p[d];
    return 2.0 * atanh(sqrt(n2));
}

__device__ double lca_depth(const double *a, const double *b, int D) {
    double p[GEO_MAX_D];
    lca_point(a, b, D, p);
    return depth_of(p, D);
}

// adjoint of y = (r2 / |x - r|^2) (x - r) + r with r2 = |r|^2 - 1; accumulates into gx, gr
__device__ void invert_vjp(const double *r, double r2, const double *x, const double *gy, int D, double *gx, double *gr) {
    double u[GEO_MAX_D];
    double s = 0.0, gu_dot = 0.0;
    for (int d = 0; d < D; ++d) { u[d] = x[d] - r[d]; s += u[d] * u[d]; gu_dot += gy[d] * u[d]; }
    const double c = r2 / s;
    const double g_r2 = gu_dot / s, g_s = -gu_dot * r2 / (s * s);
    for (int d = 0; d < D; ++d) {
        const double gu = c * gy[d] + 2.0 * g_s * u[d];
        gx[d] += gu;
        gr[d] += gy[d] - gu + 2.0 * g_r2 * r[d];
    }
}

// adjoint of proj = lca_point(a, b): ga += J_a^T gproj, gb += J_b^T gproj
__device__ void lca_point_vjp(const double *a, const double *b, int D, const double *gproj, double *ga, double *gb) {
    double r[GEO_MAX_D], binv[GEO_MAX_D], m[GEO_MAX_D], o[GEO_MAX_D];
    double na2 = 0.0;
    for (int d = 0; d < D; ++d) na2 += a[d] * a[d];
    double rr = 0.0;
    for (int d = 0; d < D; ++d) { r[d] = a[d] / na2; rr += r[d] * r[d]; }
    const double r2 = rr - 1.0;
    double s1 = 0.0;
    for (int d = 0; d < D; ++d) { binv[d] = b[d] - r[d]; s1 += binv[d] * binv[d]; }
    const double c1 = r2 / s1;
    double dot = 0.0, nv_raw = 0.0;
    for (int d = 0; d < D; ++d) { binv[d] = c1 * binv[d] + r[d]; dot += a[d] * binv[d]; nv_raw += binv[d] * binv[d]; }
    const double nv = fmax(nv_raw, GEO_MIN_NORM);
    double s2 = 0.0;
    for (int d = 0; d < D; ++d) { m[d] = 2.0 * (dot * binv[d] / nv) - a[d]; const double u = m[d] - r[d]; s2 += u * u; }
    const double c2 = r2 / s2;
    double no = 0.0;
    for (int d = 0; d < D; ++d) { o[d] = c2 * (m[d] - r[d]) + r[d]; no += o[d] * o[d]; }
    const double q = sqrt(1.0 - no), h = 1.0 + q;
    // halve
    double go[GEO_MAX_D], gr[GEO_MAX_D], gm[GEO_MAX_D], gbinv[GEO_MAX_D];
    double gp_dot_o = 0.0;
    for (int d = 0; d < D; ++d) gp_dot_o += gproj[d] * o[d];
    const double gq = -gp_dot_o / (h * h);
    for (int d = 0; d < D; ++d) { go[d] = gproj[d] / h - (gq / q) * o[d]; gr[d] = 0.0; gm[d] = 0.0; gbinv[d] = 0.0; }
    // second inversion: o = invert(r, m)
    invert_vjp(r, r2, m, go, D, gm, gr);
    // mirror: m = 2 (dot/nv) binv - a ; dot = a . binv ; nv = max(|binv|^2, MIN)
    double gm_dot_binv = 0.0;
    for (int d = 0; d < D; ++d) gm_dot_binv += gm[d] * binv[d];
    const double k = dot / nv;
    const double g_k = 2.0 * gm_dot_binv;
    const double g_dot = g_k / nv;
    const double g_nv = nv_raw > GEO_MIN_NORM ? -g_k * dot / (nv * nv) : 0.0;
    for (int d = 0; d < D; ++d) {
        ga[d] += -gm[d] + g_dot * binv[d];
        gbinv[d] += 2.0 * k * gm[d] + g_dot * a[d] + 2.0 * g_nv * binv[d];
    }
    // first inversion: binv = invert(r, b)
    invert_vjp(r, r2, b, gbinv, D, gb, gr);
    // centre: r = a / |a|^2
    double gr_dot_a = 0.0;
    for (int d = 0; d < D; ++d) gr_dot_a += gr[d] * a[d];
    for (int d = 0; d < D; ++d) ga[d] += gr[d] / na2 - 2.0 * gr_dot_a / (na2 * na2) * a[d];
}

// adjoint of t = depth_of(lca_point(a, b))
__device__ void lca_depth_vjp(const double *a, const double *b, int D, double gt, double *ga, double *gb) {
    double p[GEO_MAX_D], gp[GEO_MAX_D];
    lca_point(a, b, D, p);
    double n2 = 0.0;
    for (int d = 0; d < D; ++d) n2 += p[d] * p[d];
    const double n = sqrt(n2);
    const double gn = n > 0.0 ? gt * 2.0 / (1.0 - n2) / n : 0.0;
    for (int d = 0; d < D; ++d) gp[d] = gn * p[d];
    lca_point_vjp(a, b, D, gp, ga, gb);
}

__global__ void __launch_bounds__(1024) geo_soft_kernel(const double *__restrict__ x, int S, int D, double tau,
                                                        double *__restrict__ ws, int64_t *__restrict__ peel,
                                                        double *__restrict__ int_locs, double *__restrict__ blens,
                                                        int32_t *__restrict__ slots, double *__restrict__ tape,
                                                        int32_t *__restrict__ flags) {
    extern __shared__ unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    double *M = ws + (size_t)b * geo_ws_doubles(S, D);  // depth of lca(locs[i], locs[j]), symmetric store
    double *A = M + (size_t)S * S;                      // scores handed to the soft argmin
    double *locs = A + (size_t)S * S;                   // [S][D] current locations
    double *sh = reinterpret_cast<double *>(smem_raw);  // [32]
    double *newp = sh + 32;                             // [GEO_MAX_D]
    int *live = reinterpret_cast<int *>(newp + GEO_MAX_D);  // [S]
    int *node_of = live + S;                                // [S]
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *il = int_locs + (size_t)b * (S - 1) * D;
    double *bl = blens + (size_t)b * (2 * S - 2);
    int32_t *sl = slots + (size_t)b * (S - 1) * 2;
    double *tp = tape + (size_t)b * (S - 1) * 2 * D;
    int32_t *fl = flags + (size_t)b * (S - 1);

    for (int i = tid; i < S * D; i += T) locs[i] = x[(size_t)b * S * D + i];
    for (int i = tid; i < S; i += T) { live[i] = 1; node_of[i] = i; }
    __syncthreads();
    for (long long l = tid; l < (long long)S * S; l += T) {
        const int i = (int)(l / S), j = (int)(l - (long long)i * S);
        if (i == j) M[l] = 0.0;
        else if (i > j) {
            const double t = lca_depth(locs + (size_t)i * D, locs + (size_t)j * D, D);  // utils.get_plca: row = first argument
            M[l] = t;
            M[(size_t)j * S + i] = t;
        }
    }
    __syncthreads();

    for (int step = 0; step < S - 1; ++step) {
        double mx = 0.0;  // the masked matrix always contains zeros (diagonal, retired rows)
        for (long long l = tid; l < (long long)S * S; l += T) {
            const int i = (int)(l / S), j = (int)(l - (long long)i * S);
            if (i > j && live[i] && live[j]) mx = fmax(mx, M[l]);
        }
        const double big = block_reduce(mx, sh, 2) + 1.0;
        for (long long l = tid; l < (long long)S * S; l += T) {
            const int i = (int)(l / S), j = (int)(l - (long long)i * S);
            const double t = M[l];
            A[l] = (i > j && live[i] && live[j] && t != 0.0) ? -t : big;  // -where(tril != 0, tril, -big)
        }
        __syncthreads();
        SoftCtx c{A, nullptr, nullptr, nullptr, nullptr, 0, S, 0, 0.0, (long long)S * S, 0};
        double idx;
        bool ok;
        soft_argmin_block<2>(c, tau, sh, idx, ok);
        double rowf, colf;
        unravel(idx, (double)S, rowf, colf);
        const int f = (int)rint(rowf), g = (int)rint(colf);
        const bool bad = !(idx == idx) || f < 0 || f >= S || g < 0 || g >= S || f == g || !live[f] || !live[g];
        if (bad) {
            if (tid == 0)
                for (int t = step; t < S - 1; ++t) fl[t] = 8;
            return;
        }
        __syncthreads();  // everyone has read live[] before it changes
        if (tid == 0) {
            const double *pf = locs + (size_t)f * D, *pg = locs + (size_t)g * D;
            double np_[GEO_MAX_D];
            lca_point(pf, pg, D, np_);
            const int parent = S + step;
            pl[step * 3 + 0] = node_of[f];
            pl[step * 3 + 1] = node_of[g];
            pl[step * 3 + 2] = parent;
            bl[node_of[f]] = lca_depth(pf, np_, D);
            bl[node_of[g]] = lca_depth(pg, np_, D);
            sl[step * 2 + 0] = f;
            sl[step * 2 + 1] = g;
            for (int d = 0; d < D; ++d) {
                tp[(size_t)(step * 2 + 0) * D + d] = pf[d];
                tp[(size_t)(step * 2 + 1) * D + d] = pg[d];
                il[(size_t)step * D + d] = np_[d];
                newp[d] = np_[d];
            }
            fl[step] = ok ? 0 : 4;
            node_of[g] = parent;
            live[f] = 0;
        }
        __syncthreads();
        for (int d = tid; d < D; d += T) locs[(size_t)g * D + d] = newp[d];
        __syncthreads();
        for (int i = tid; i < S; i += T) {
            if (i == g || !live[i]) continue;
            const double t = i > g ? lca_depth(locs + (size_t)i * D, locs + (size_t)g * D, D)
                                   : lca_depth(locs + (size_t)g * D, locs + (size_t)i * D, D);
            M[(size_t)i * S + g] = t;
            M[(size_t)g * S + i] = t;
        }
        __syncthreads();
    }
}

// one thread per tree: reverse replay of the joins with the adjoint of every slot's location
__global__ void geo_soft_bwd_kernel(const int64_t *__restrict__ peel, const int32_t *__restrict__ slots,
                                    const double *__restrict__ tape, const double *__restrict__ gblens,
                                    const double *__restrict__ gint, int B, int S, int D, double *__restrict__ gx) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    const int32_t *sl = slots + (size_t)b * (S - 1) * 2;
    const double *tp = tape + (size_t)b * (S - 1) * 2 * D;
    const double *gb_ = gblens + (size_t)b * (2 * S - 2);
    double *adj = gx + (size_t)b * S * D;  // adjoint of the CURRENT content of every slot
    for (int i = 0; i < S * D; ++i) adj[i] = 0.0;
    for (int step = S - 2; step >= 0; --step) {
        const int f = sl[step * 2 + 0], g = sl[step * 2 + 1];
        if (f < 0 || g < 0 || f >= S || g >= S) continue;  // tree stopped early (flag 8)
        const double *pf = tp + (size_t)(step * 2 + 0) * D, *pg = tp + (size_t)(step * 2 + 1) * D;
        double np_[GEO_MAX_D], gnew[GEO_MAX_D], gf[GEO_MAX_D], gg[GEO_MAX_D];
        lca_point(pf, pg, D, np_);
        for (int d = 0; d < D; ++d) {
            gnew[d] = adj[(size_t)g * D + d] + (gint ? gint[((size_t)b * (S - 1) + step) * D + d] : 0.0);
            gf[d] = 0.0;
            gg[d] = 0.0;
        }
        lca_depth_vjp(pf, np_, D, gb_[pl[step * 3 + 0]], gf, gnew);
        lca_depth_vjp(pg, np_, D, gb_[pl[step * 3 + 1]], gg, gnew);
        lca_point_vjp(pf, pg, D, gnew, gf, gg);
        for (int d = 0; d < D; ++d) {
            adj[(size_t)g * D + d] = gg[d];
            adj[(size_t)f * D + d] += gf[d];
        }
    }
}

// hard variant on hyperboloid points.  Node n's sheet coordinates live in pts[n], its ball
// coordinates in ball[n]; the depth matrix is indexed by storage slot (a new node takes the slot of
// its first child).  The reference pops a binary heap keyed by -depth; away from exact ties that is
// the deepest live pair, which is what is selected here (lowest slot pair on ties).
__global__ void __launch_bounds__(1024) geo_hard_kernel(const double *__restrict__ sheet, int S, int D1,
                                                        double *__restrict__ ws, int64_t *__restrict__ peel,
                                                        double *__restrict__ int_locs) {
    extern __shared__ unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int D = D1 - 1;
    double *M = ws + (size_t)b * geo_ws_doubles(S, D);
    double *ball = M + 2 * (size_t)S * S;              // [S][D] ball coordinates by slot
    double *sh = reinterpret_cast<double *>(smem_raw);  // [32]
    unsigned long long *shk = reinterpret_cast<unsigned long long *>(sh + 32);  // [32]
    int *live = reinterpret_cast<int *>(shk + 32);      // [S]
    int *node_of = live + S;                            // [S]
    int64_t *pl = peel + (size_t)b * (S - 1) * 3;
    double *il = int_locs + (size_t)b * (S - 1) * D1;
    const double *z = sheet + (size_t)b * S * D1;

    for (int i = tid; i < S; i += T) {
        live[i] = 1;
        node_of[i] = i;
        for (int d = 0; d < D; ++d) ball[(size_t)i * D + d] = z[(size_t)i * D1 + 1 + d] / (1.0 + z[(size_t)i * D1]);
    }
    __syncthreads();
    for (long long l = tid; l < (long long)S * S; l += T) {
        const int i = (int)(l / S), j = (int)(l - (long long)i * S);
        if (i > j) M[l] = lca_depth(ball + (size_t)i * D, ball + (size_t)j * D, D);  // utils.get_plca_np order
    }
    __syncthreads();
    for (int step = 0; step < S - 1; ++step) {
        // deepest live pair; key = flat slot index, smallest wins ties
        QKey best{DBL_MAX, ~0ull};
        for (long long l = tid; l < (long long)S * S; l += T) {
            const int i = (int)(l / S), j = (int)(l - (long long)i * S);
            if (i > j && live[i] && live[j]) best = qmin(best, QKey{-M[l], (unsigned long long)l});
        }
        best = warp_qmin(best);
        __syncthreads();
        if (lane == 0) { sh[warp] = best.q; shk[warp] = best.key; }
        __syncthreads();
        best = QKey{sh[0], shk[0]};
        for (int w = 1; w < nwarp; ++w) best = qmin(best, QKey{sh[w], shk[w]});
        if (best.key == ~0ull) return;  // NaN input: nothing comparable
        const int si = (int)(best.key / S), sj = (int)(best.key - (unsigned long long)si * S);
        const int ni = node_of[si], nj_ = node_of[sj];
        // Edge(from_, to_): initial pairs are (larger, smaller) tip id; later pairs (older node, newer node)
        const bool both_tips = ni < S && nj_ < S;
        const int a_slot = both_tips ? (ni > nj_ ? si : sj) : (ni < nj_ ? si : sj);
        const int b_slot = a_slot == si ? sj : si;
        const int parent = S + step;
        __syncthreads();
        if (tid == 0) {
            double p[GEO_MAX_D];
            lca_point(ball + (size_t)a_slot * D, ball + (size_t)b_slot * D, D, p);
            double n2 = 0.0;
            for (int d = 0; d < D; ++d) n2 += p[d] * p[d];
            // Chyp_np.poincare_to_hyper, then back to the ball as the reference does on every later use
            const double z0 = (1.0 + n2) / (1.0 - n2);
            il[(size_t)step * D1] = z0;
            for (int d = 0; d < D; ++d) {
                const double zd = 2.0 * p[d] / (1.0 - n2 + DBL_EPSILON);
                il[(size_t)step * D1 + 1 + d] = zd;
                ball[(size_t)a_slot * D + d] = zd / (1.0 + z0);
            }
            pl[step * 3 + 0] = node_of[a_slot];
            pl[step * 3 + 1] = node_of[b_slot];
            pl[step * 3 + 2] = parent;
            node_of[a_slot] = parent;
            live[b_slot] = 0;
        }
        __syncthreads();
        for (int i = tid; i < S; i += T) {
            if (i == a_slot || !live[i]) continue;
            // new edges are Edge(dist(old, new), old, new): the older node is the first argument
            const double t = lca_depth(ball + (size_t)i * D, ball + (size_t)a_slot * D, D);
            if (i > a_slot) M[(size_t)i * S + a_slot] = t;
            else M[(size_t)a_slot * S + i] = t;
        }
        __syncthreads();
    }
}

// Batched callers' policy for soft-NJ status flags: a sample whose joins carry a bit of `mask` has a
// topology the reference would not have produced (guard tripped: must be redone exhaustively; invalid
// selection: the reference indexes a masked node there).  Its lnL and gradient row are poisoned with
// NaN so that it cannot be consumed silently, and the two kinds are counted for the host.
__global__ void nj_guard_kernel(const int32_t *__restrict__ flags, int B, int nJ, uint32_t mask, int E,
                                double *__restrict__ loglik, double *__restrict__ grad, int32_t *__restrict__ counts) {
    const int b = blockIdx.x;
    if (b >= B) return;
    unsigned bits = 0;
    for (int t = threadIdx.x; t < nJ; t += blockDim.x) bits |= (unsigned)flags[(size_t)b * nJ + t];
    bits = __reduce_or_sync(0xffffffffu, bits);
    __shared__ unsigned sh_bits[8];
    if ((threadIdx.x & 31) == 0) sh_bits[threadIdx.x >> 5] = bits;
    __syncthreads();
    bits = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) bits |= sh_bits[w];
    bits &= mask;
    if (!bits) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (threadIdx.x == 0) {
        if (loglik) loglik[b] = nan;
        if (bits & 4u) atomicAdd(counts + 0, 1);
        if (bits & 8u) atomicAdd(counts + 1, 1);
    }
    if (grad)
        for (int e = threadIdx.x; e < E; e += blockDim.x) grad[(size_t)b * E + e] = nan;
}

}  // namespace dodo

using namespace dodo;

extern "C" int dodo_nj_guard(const int32_t *d_flags, int B, int S, uint32_t mask, double *d_loglik,
                             double *d_grad_blens, int32_t *d_counts, void *stream_) {
    DODO_CHECK_ARG(d_flags && d_counts, "dodo_nj_guard: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_nj_guard: need B>=1, S>=2");
    cudaStream_t stream = (cudaStream_t)stream_;
    DODO_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(int32_t), stream));
    nj_guard_kernel<<<B, 64, 0, stream>>>(d_flags, B, S - 1, mask, 2 * S - 2, d_loglik, d_grad_blens, d_counts);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_soft_argmin(const double *d_Q, int B, int rows, int cols, double tau, double *d_rowcol, void *stream_) {
    DODO_CHECK_ARG(d_Q && d_rowcol, "dodo_soft_argmin: NULL argument");
    DODO_CHECK_ARG(B >= 1 && rows >= 1 && cols >= 1, "dodo_soft_argmin: need B, rows, cols >= 1");
    DODO_CHECK_ARG(tau > 0.0, "dodo_soft_argmin: tau must be positive");
    soft_argmin_kernel<<<B, 256, 0, (cudaStream_t)stream_>>>(d_Q, rows, cols, tau, d_rowcol);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" uint64_t dodo_nj_workspace_bytes(int B, int S) {
    return (uint64_t)B * nj_ws_doubles(S) * sizeof(double);
}

// CTA size by taxon count (measured on B200: 16-500 taxa, 64-128 trees per launch): the per-join scans
// want many threads long before the matrix is large
static int nj_threads(int S, int B = 1) {
    if (const char *e = getenv("DODO_NJ_THREADS")) {  // tuning sweeps (profiles/tools/nj_sweep.py)
        const int t = atoi(e);
        if (t >= 64 && t <= 1024 && t % 32 == 0) return t;
    }
    const int t = S <= 32 ? 256 : (S <= 128 ? 512 : 1024);
    // more trees than SMs: two 512-thread CTAs share an SM instead of queueing a second round of
    // 1024-thread CTAs behind the first
    if (t == 1024 && B > sm_count()) return 512;
    return t;
}

// host part of the prior: normalising constants and the uniform-topology term (base_model.py:506-521)
static int make_prior(const dodo_prior_t *p, int S, PriorDev *out) {
    memset(out, 0, sizeof(*out));
    if (p == nullptr) return DODO_OK;
    DODO_CHECK_ARG(p->d_ln_prior != nullptr, "prior: d_ln_prior is NULL");
    DODO_CHECK_ARG(p->variant == DODO_PRIOR_TORCH || p->variant == DODO_PRIOR_NUMPY, "prior: bad variant");
    DODO_CHECK_ARG(p->aT > 0.0 && p->bT > 0.0 && p->a > 0.0 && p->c > 0.0, "prior: parameters must be positive");
    DODO_CHECK_ARG(S >= 3, "prior: the gamma-Dirichlet prior needs at least 3 taxa");
    const double n = (double)S;
    double k = p->aT * log(p->bT) - lgamma(p->aT) + lgamma(p->a * n + p->a * p->c * (n - 3.0)) - n * lgamma(p->a) -
               (n - 3.0) * lgamma(p->a * p->c);
    double topo = 0.0;
    for (int i = 2 * S - 5; i > 0; i -= 2) topo += log((double)i);
    out->aT = p->aT; out->bT = p->bT; out->a = p->a; out->c = p->c;
    out->konst = k - topo;
    out->variant = p->variant;
    out->ln_prior = p->d_ln_prior;
    out->grad = p->d_grad_blens;
    return DODO_OK;
}

extern "C" int dodo_prior_gamma_dir(const double *d_blens, int B, int S, const dodo_prior_t *prior, void *stream_) {
    DODO_CHECK_ARG(d_blens && prior, "dodo_prior_gamma_dir: NULL argument");
    DODO_CHECK_ARG(B >= 1, "dodo_prior_gamma_dir: need B>=1");
    PriorDev pr;
    int rc = make_prior(prior, S, &pr);
    if (rc != DODO_OK) return rc;
    prior_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream_>>>(d_blens, B, S, pr);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

// how many clusters of CL CTAs of the 16-warp lean hard kernel the device holds at once (0: this size cannot launch)
typedef void (*hard_lean_fn_t)(const double *, int, double *, int64_t *, double *, PriorDev);
static hard_lean_fn_t hard_lean_cluster_fn(int W, int CL) {
    if (W == 8) return CL == 8 ? nj_hard_lean_kernel<8, 8> : (CL == 4 ? nj_hard_lean_kernel<8, 4> : nj_hard_lean_kernel<8, 2>);
    return CL == 8 ? nj_hard_lean_kernel<16, 8> : (CL == 4 ? nj_hard_lean_kernel<16, 4> : nj_hard_lean_kernel<16, 2>);
}
static int hard_lean_max_clusters(int S, int W, int CL) {
    static int cache[NJ_HARD_LEAN_MAX + 1][2][9];  // 0 = not asked yet, else 1 + answer
    int &slot = cache[S][W == 8 ? 0 : 1][CL];
    if (slot) return slot - 1;
    hard_lean_fn_t fn = hard_lean_cluster_fn(W, CL);
    const size_t lsm = nj_hard_lean_smem_bytes(S, W, CL);
    int n = 0;
    if (lsm <= 227 * 1024 && (lsm <= 48 * 1024 ||
                              cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm) == cudaSuccess)) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 * CL);
        cfg.blockDim = dim3(32 * W);
        cfg.dynamicSmemBytes = lsm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    slot = n + 1;
    return n;
}

extern "C" int dodo_nj_hard(const double *d_pdm, int B, int S, void *d_ws, int64_t *d_peel, double *d_blens,
                            const dodo_prior_t *prior, void *stream_) {
    DODO_CHECK_ARG(d_pdm && d_ws && d_peel && d_blens, "dodo_nj_hard: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_nj_hard: need B>=1, S>=2");
    PriorDev pr;
    { int rc = make_prior(prior, S, &pr); if (rc != DODO_OK) return rc; }
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t smem = 2 * (size_t)nj_ld(S) * sizeof(double) + 3 * (size_t)nj_ld(S) * sizeof(int);
    if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default as well
        DODO_CUDA(cudaFuncSetAttribute(nj_hard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int lean_first = 0;
    if (S <= NJ_HARD_LEAN_MAX && S >= 3 && !getenv("DODO_NJ_NO_LEAN")) {
        int W = S <= 64 ? 8 : 16;  // measured (profiles/tools/nj_sweep.py): 64 taxa 0.147 ms at W = 8, 200 taxa 0.93 ms at W = 16
        if (const char *e = getenv("DODO_NJ_LEAN_W")) {  // tuning sweeps
            const int w = atoi(e);
            if (w == 8 || w == 16 || w == 32) W = w;
        }
        // Few trees on this GPU (a shard of a sharded run, or fewer chains than SMs): a cluster of CL CTAs per tree
        // splits the scans; trees above 100 taxa only (below, a join is bound by its barriers, not by the scan).
        // Measured at 200 taxa (profiles/tools/nj_cluster.py, one CTA per tree: 0.94 ms): 8 CTAs of 8 warps 0.68 ms
        // while the clusters take at most half of the SMs, 4 CTAs of 16 warps 0.73 ms and 2 CTAs 0.85 ms while all
        // clusters are resident at once -- a second wave of clusters costs more than the split saves.
        int CL = 1;
        if (S > 100 && W == 16) {
            if (B * 16 <= sm_count() && B <= hard_lean_max_clusters(S, 8, 8)) { W = 8; CL = 8; }
            else if (B <= hard_lean_max_clusters(S, 16, 4)) CL = 4;
            else if (B <= hard_lean_max_clusters(S, 16, 2)) CL = 2;
        }
        if (const char *e = getenv("DODO_NJ_CLUSTER")) {  // tuning sweeps
            const int c = atoi(e);
            if ((c == 1 || c == 2 || c == 4 || c == 8) && W <= 16) CL = c;
        }
        if (CL > 1) {
            hard_lean_fn_t fn = hard_lean_cluster_fn(W, CL);
            const size_t lsm = nj_hard_lean_smem_bytes(S, W, CL);
            if (lsm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)B * CL);
            cfg.blockDim = dim3(32 * W);
            cfg.dynamicSmemBytes = lsm;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            DODO_CUDA(cudaLaunchKernelEx(&cfg, fn, d_pdm, S, (double *)d_ws, d_peel, d_blens, pr));
            count_launch();
        } else {
            const size_t lsm = nj_hard_lean_smem_bytes(S, W);
            hard_lean_fn_t fn = W == 8 ? nj_hard_lean_kernel<8, 1> : (W == 16 ? nj_hard_lean_kernel<16, 1> : nj_hard_lean_kernel<32, 1>);
            if (lsm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
            fn<<<B, 32 * W, lsm, stream>>>(d_pdm, S, (double *)d_ws, d_peel, d_blens, pr);
            DODO_LAUNCH_CHECK();
        }
        lean_first = 1;
    }
    nj_hard_kernel<<<B, nj_threads(S, B), smem, stream>>>(d_pdm, S, (double *)d_ws, d_peel, d_blens, pr, lean_first);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_nj_soft(const double *d_pdm, int B, int S, double tau, void *d_ws, int64_t *d_peel,
                            double *d_blens, int32_t *d_flags, const dodo_prior_t *prior, void *stream_) {
    DODO_CHECK_ARG(d_pdm && d_ws && d_peel && d_blens && d_flags, "dodo_nj_soft: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_nj_soft: need B>=1, S>=2");
    PriorDev pr;
    { int rc = make_prior(prior, S, &pr); if (rc != DODO_OK) return rc; }
    DODO_CHECK_ARG(tau != 0.0 && tau == tau, "dodo_nj_soft: tau must be a non-zero number");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int full = tau < 0.0 ? 1 : 0;  // negative tau selects the exhaustive N^2 enumeration
    const double t = fabs(tau);
    size_t smem = (3 * (size_t)nj_ld(S) + 32) * sizeof(double) + (3 * (size_t)nj_ld(S) + 2 * (size_t)S) * sizeof(int);
    if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default as well
        DODO_CUDA(cudaFuncSetAttribute(nj_soft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int lean_first = 0;
    if (!full && S <= NJ_LEAN_MAX && !getenv("DODO_NJ_NO_LEAN")) {
        int W = S <= 32 ? 4 : 8;  // measured (profiles/tools/nj_sweep.py): 27 taxa 0.065 ms at W = 4, 64 taxa 0.138 ms at W = 8
        if (const char *e = getenv("DODO_NJ_LEAN_W")) {  // tuning sweeps
            const int w = atoi(e);
            if (w == 1 || w == 2 || w == 4 || w == 8) W = w;
        }
        const size_t lsm = nj_lean_smem_bytes(S, W);
        auto fn = W == 1 ? nj_soft_lean_kernel<1> : (W == 2 ? nj_soft_lean_kernel<2> : (W == 4 ? nj_soft_lean_kernel<4> : nj_soft_lean_kernel<8>));
        if (lsm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
        fn<<<B, 32 * W, lsm, stream>>>(d_pdm, S, t, (double *)d_ws, d_peel, d_blens, d_flags, pr);
        DODO_LAUNCH_CHECK();
        lean_first = 1;
    }
    nj_soft_kernel<<<B, nj_threads(S, B) < NJ_SOFT_MAXT ? nj_threads(S, B) : NJ_SOFT_MAXT, smem, stream>>>(d_pdm, S, t, full, (double *)d_ws, d_peel, d_blens, d_flags, pr,
                                                          lean_first);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_nj_soft_bwd(const int64_t *d_peel, const int32_t *d_flags, const double *d_grad_blens, int B,
                                int S, void *d_ws, double *d_grad_pdm, void *stream_) {
    DODO_CHECK_ARG(d_peel && d_flags && d_grad_blens && d_ws && d_grad_pdm, "dodo_nj_soft_bwd: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_nj_soft_bwd: need B>=1, S>=2");
    cudaStream_t stream = (cudaStream_t)stream_;
    // O(S) work per un-join behind four barriers: as few warps as cover the S slots
    int bwd_threads = (S + 31) / 32 * 32;
    bwd_threads = bwd_threads < 64 ? 64 : (bwd_threads > 1024 ? 1024 : bwd_threads);
    size_t smem = 32 * sizeof(double) + (2 * (size_t)S + (size_t)S) * sizeof(int);
    if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default as well
        DODO_CUDA(cudaFuncSetAttribute(nj_soft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (latency-bound batches only: with many more trees than SMs the replay's per-tree latency is hidden by the
    // other trees of the SM and its O(S^2) work per tree wins over the product's O(S^3): 16 128 replays of the
    // batched-Jacobian cross-check take 2.8 ms by replay, 7.3 ms by product)
    if (S <= NJ_BWD_LEAN_MAX && S >= 2 && B <= 2 * sm_count() && !getenv("DODO_NJ_NO_LEAN")) {
        const size_t lsm = nj_bwd_lean_smem_bytes(S);
        if (lsm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(nj_soft_bwd_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
        int lt = ((S + 1) / 2 * ((S + 1) / 2) + 31) / 32 * 32;  // one 2 x 2 tile per thread, and one warp beyond the S column threads
        lt = lt < ((S + 31) / 32 + 1) * 32 ? ((S + 31) / 32 + 1) * 32 : (lt > 1024 ? 1024 : lt);
        nj_soft_bwd_lean_kernel<<<B, lt, lsm, stream>>>(d_peel, d_flags, d_grad_blens, S, d_grad_pdm);
        DODO_LAUNCH_CHECK();
        return DODO_OK;
    }
    nj_soft_bwd_kernel<<<B, bwd_threads, smem, stream>>>(d_peel, d_flags, d_grad_blens, S, (double *)d_ws, d_grad_pdm);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

// ---- geodesic connector
extern "C" uint64_t dodo_geo_workspace_bytes(int B, int S, int D) {
    return (uint64_t)B * geo_ws_doubles(S, D) * sizeof(double);
}

extern "C" int dodo_geo_soft(const double *d_x, int B, int S, int D, double tau, void *d_ws, int64_t *d_peel,
                             double *d_int_locs, double *d_blens, int32_t *d_slots, double *d_tape,
                             int32_t *d_flags, void *stream_) {
    DODO_CHECK_ARG(d_x && d_ws && d_peel && d_int_locs && d_blens && d_slots && d_tape && d_flags,
                   "dodo_geo_soft: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_geo_soft: need B>=1, S>=2");
    DODO_CHECK_ARG(D >= 1 && D < GEO_MAX_D, "dodo_geo_soft: embedding dimension must be 1..15");
    DODO_CHECK_ARG(tau > 0.0, "dodo_geo_soft: tau must be positive");
    cudaStream_t stream = (cudaStream_t)stream_;
    DODO_CUDA(cudaMemsetAsync(d_slots, 0xff, (size_t)B * (S - 1) * 2 * sizeof(int32_t), stream));
    size_t smem = (32 + GEO_MAX_D) * sizeof(double) + 2 * (size_t)S * sizeof(int);
    if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default as well
        DODO_CUDA(cudaFuncSetAttribute(geo_soft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    geo_soft_kernel<<<B, nj_threads(S), smem, stream>>>(d_x, S, D, tau, (double *)d_ws, d_peel, d_int_locs, d_blens,
                                                        d_slots, d_tape, d_flags);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_geo_soft_bwd(const int64_t *d_peel, const int32_t *d_slots, const double *d_tape,
                                 const double *d_grad_blens, const double *d_grad_int_locs, int B, int S, int D,
                                 double *d_grad_x, void *stream_) {
    DODO_CHECK_ARG(d_peel && d_slots && d_tape && d_grad_blens && d_grad_x, "dodo_geo_soft_bwd: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_geo_soft_bwd: need B>=1, S>=2");
    DODO_CHECK_ARG(D >= 1 && D < GEO_MAX_D, "dodo_geo_soft_bwd: embedding dimension must be 1..15");
    geo_soft_bwd_kernel<<<(B + 31) / 32, 32, 0, (cudaStream_t)stream_>>>(d_peel, d_slots, d_tape, d_grad_blens,
                                                                        d_grad_int_locs, B, S, D, d_grad_x);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

extern "C" int dodo_geo_hard(const double *d_sheet, int B, int S, int D1, void *d_ws, int64_t *d_peel,
                             double *d_int_locs, void *stream_) {
    DODO_CHECK_ARG(d_sheet && d_ws && d_peel && d_int_locs, "dodo_geo_hard: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2, "dodo_geo_hard: need B>=1, S>=2");
    DODO_CHECK_ARG(D1 >= 2 && D1 <= GEO_MAX_D, "dodo_geo_hard: sheet points need 2..16 coordinates");
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t smem = 32 * sizeof(double) + 32 * sizeof(unsigned long long) + 2 * (size_t)S * sizeof(int);
    if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default as well
        DODO_CUDA(cudaFuncSetAttribute(geo_hard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    geo_hard_kernel<<<B, nj_threads(S), smem, stream>>>(d_sheet, S, D1, (double *)d_ws, d_peel, d_int_locs);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

// K3: fused transition-matrix formation + Felsenstein pruning (forward and gradient), sm_100a.
//
// Replaces phylo.calculate_treelikelihood (dodonaphy/phylo.py:132-139) + torch autograd,
// Cphylo.calculate_treelikelihood / compute_LL_np (dodonaphy/cython/Cphylo.pyx:9-42) and
// PhyloModel.get_transition_mats (dodonaphy/phylomodel.py:76-114).
//
// Design: site patterns are conditionally independent given the tree, so one
// thread owns one (sample, pattern) and walks the whole tree itself, once per rate category.
//   * the reference's post-order (peel) is re-ordered per sample into a depth-first post-order in
//     which every join consumes the result of the join just before it (kept in registers); only the
//     "other" internal child has to be fetched back;
//   * per-node "messages" msg_n = P_n (msg_l * msg_r) are streamed once to a per-CTA scratch ring
//     (one 256-bit store per message, 1 KB contiguous per warp) for the backward sweep; tips never
//     touch HBM as partials: their message is a 32-byte row of a per-edge table in shared memory,
//     selected by the 4-bit IUPAC mask;
//   * the backward sweep is the same schedule reversed (a valid pre-order), carries the upper
//     message in registers, reads each stored message once and reduces d lnL / d t_e over the 32
//     patterns of the warp with a fixed-order butterfly -> run-to-run reproducible.
//   * cherries (both children tips) that feed their parent directly are never stored: backward they
//     are recomputed, or -- while shared memory is left -- looked up in a per-(sample, category) table
//     of their 5 x 5 possible messages (cherry_kernel).
// HBM traffic per (pattern, category): one 32 B write + one 32 B read per stored internal node -- about
// 2/5 of the "materialise every partial" model used for the roofline denominator.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dodo {

constexpr int MAX_K = 8;

struct ModelDev {
    double U[16], Uinv[16], lam[4], Q[16], pi[4];
    double rates[MAX_K];
    double rlam[MAX_K][4];  // rates[k] * lam[m] (model-gradient launches)
    int kind;
};

struct V4 {
    double a, b, c, d;
};

__device__ __forceinline__ V4 vmul(const V4 &x, const V4 &y) {
    return V4{x.a * y.a, x.b * y.b, x.c * y.c, x.d * y.d};
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
// Cherries (both children tips) that are the b-child of their parent are never written to scratch:
// forward they stay in registers, backward they are recomputed from the two tip tables in shared
// memory (flags bit3 on the parent: "b is a recomputed cherry"; bit4 on the cherry: "do not store").
// ------------------------------------------------------------------------------------------
constexpr int FL_A_CHILD = 1, FL_ROOT = 2, FL_B_CHERRY = 8, FL_NO_STORE = 16, FL_CHERRY = 32;  // cherry id in bits 8+
__global__ void schedule_kernel(const int64_t *__restrict__ peel, int nb, int S, int4 *__restrict__ ops,
                                int *__restrict__ chop) {
    // one warp per sample: the peel is staged into shared memory by the warp, lane 0 runs the two
    // (inherently sequential) passes there, the warp writes the schedule back coalesced
    extern __shared__ __align__(16) int sched_sm[];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= nb) return;
    const int nI = S - 1, Nn = 2 * S - 1;
    const int64_t *pl = peel + (size_t)b * nI * 3;
    int4 *op = reinterpret_cast<int4 *>(sched_sm);               // [S-1] schedule being built
    int *Pl = reinterpret_cast<int *>(op + nI), *Pr = Pl + S, *Pp = Pr + S;  // peel rows
    int *Lc = Pp + S, *Rc = Lc + S, *st = Rc + S, *isz = st + S, *start = isz + S, *afl = start + S;
    int *cho = afl + S;  // schedule position of cherry c
    int nch = 0;
    for (int i = lane; i < nI; i += 32) {
        // memory safety for malformed input (the host shim validates and raises)
        Pl[i] = min(max((int)pl[i * 3 + 0], 0), Nn - 1);
        Pr[i] = min(max((int)pl[i * 3 + 1], 0), Nn - 1);
        Pp[i] = min(max((int)pl[i * 3 + 2], S), Nn - 1);
        isz[i] = 0; afl[i] = 0; start[i] = 0; Lc[i] = 0; Rc[i] = 0; st[i] = 1;
        op[i] = make_int4(0, 0, S + i, i == nI - 1 ? FL_ROOT : 0);
    }
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < nI; ++i) {
            const int l = Pl[i], r = Pr[i], idx = Pp[i] - S;
            Lc[idx] = l; Rc[idx] = r;
            const int sl = l < S ? 1 : st[l - S], sr = r < S ? 1 : st[r - S];
            st[idx] = (sl == sr) ? sl + 1 : max(sl, sr);
            isz[idx] = 1 + (l >= S ? isz[l - S] : 0) + (r >= S ? isz[r - S] : 0);
        }
        for (int i = nI - 1; i >= 0; --i) {
            const int p = Pp[i], idx = p - S;
            const int l = Lc[idx], r = Rc[idx];
            int a, bb;
            const bool li = l >= S, ri = r >= S;
            if (li && ri) {
                if (st[l - S] >= st[r - S]) { a = l; bb = r; } else { a = r; bb = l; }
            } else if (li) { a = r; bb = l; }
            else { a = l; bb = r; }
            const int s0 = (i == nI - 1) ? 0 : start[idx];
            const int oi = min(max(s0 + isz[idx] - 1, 0), nI - 1);
            int fl = afl[idx] | ((i == nI - 1) ? FL_ROOT : 0);
            if (bb >= S && Lc[bb - S] < S && Rc[bb - S] < S) {
                fl |= FL_B_CHERRY;
                afl[bb - S] |= FL_NO_STORE;
            }
            if (l < S && r < S) {  // cherry: numbered in schedule order of discovery
                fl |= FL_CHERRY | (nch << 8);
                cho[nch++] = oi;
            }
            op[oi] = make_int4(a, bb, p, fl);
            if (li && ri) {
                start[a - S] = s0;
                start[bb - S] = s0 + isz[a - S];
                afl[a - S] = FL_A_CHILD;
            } else if (bb >= S) {
                start[bb - S] = s0;
            }
        }
    }
    nch = __shfl_sync(0xffffffffu, nch, 0);
    int4 *out = ops + (size_t)b * nI;
    for (int i = lane; i < nI; i += 32) out[i] = op[i];
    for (int c = lane; c < S / 2; c += 32) chop[(size_t)b * (S / 2) + c] = c < nch ? cho[c] : -1;
}

// ------------------------------------------------------------------------------------------
// Transition tables, laid out so that one (sample, category) slab is contiguous and can be staged
// into shared memory with coalesced 16-byte copies:
//   tab[b][k] = { Pint[S-1][16] : P of the edge above internal node S+i (row-major; root row unused)
//                 Ttip[S][2][5][2] : for tip edge t, five rows r_j (the four columns of P and the row sums
//                                 == P . tip_partial for the masks A,C,G,T and N/?/-), split in halves:
//                                 {r_j[0], r_j[1]} at 2j, {r_j[2], r_j[3]} at 10 + 2j.  A lane-dependent
//                                 ld.shared.v2.f64 is served per quarter-warp, 16 bytes x 8 bank groups: with
//                                 the five half-rows 16 bytes apart every lane of a quarter hits its own group
//                                 (or the same address), 4 wavefronts per load; with whole rows 32 bytes
//                                 apart rows 0 and 4 collide and every load costs 8 (measured,
//                                 profiles/tools/micro/lds_wavefronts.cu) }
// ------------------------------------------------------------------------------------------
//                 Cher[chmax][4][25] : message of tabulated cherry c for every pair of tip rows, one
//                                 plane per state; entry order cherry_entry() keeps the 16 one-hot
//                                 pairs in distinct shared-memory banks (see cherry_kernel) }
__host__ __device__ __forceinline__ unsigned cherry_entry(unsigned ja, unsigned jb) {  // rows 0..3 one-hot, 4 = any
    return (ja < 4u && jb < 4u) ? ja * 4u + jb : (ja == 4u ? 16u + jb : 21u + ja);
}
//                 Phi[2S-2][16] (model-gradient launches only, `gm`): per edge the divided differences
//                                 Phi'_mn = tau expm1((lam_m - lam_n) tau) / ((lam_m - lam_n) tau), Phi'_mm = tau,
//                                 tau = r_k t_e, of the Daleckii-Krein form of d exp(Q tau) (see model_edge) }
__host__ __device__ inline size_t tab_doubles(int S, int chmax, int gm = 0) {
    return (size_t)(S - 1) * 16 + (size_t)S * 20 + (size_t)chmax * 100 + (gm ? (size_t)(2 * S - 2) * 16 : 0);
}

__global__ void transition_kernel(ModelDev md, const double *__restrict__ blens, const double *__restrict__ P_in,
                                  int B, int S, int K, int eig, int chmax, int gm, double *__restrict__ tab) {
    const int E = 2 * S - 2;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * E * K;
    if (gid >= total) return;
    int k = gid % K;
    int e = (gid / K) % E;
    int b = gid / ((size_t)K * E);
    double P[16], ex[4] = {0, 0, 0, 0};
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
    double *slab = tab + ((size_t)b * K + k) * tab_doubles(S, chmax, gm);
    if (gm) {  // eigen models only (checked by the host)
        const double tau = blens[(size_t)b * E + e] * md.rates[k];
        double2 *ph = reinterpret_cast<double2 *>(slab + tab_doubles(S, chmax) + (size_t)e * 16);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            double row[4];
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const double d = fmin((md.lam[m] - md.lam[n]) * tau, 700.0);
                row[n] = (m == n) ? tau : tau * (fabs(d) < 1e-8 ? 1.0 + 0.5 * d : expm1(d) / d);
            }
            row[3] = exp(md.lam[m] * tau);  // column 3 of Phi' is never used (model_edge): it carries e_m for P^T u = U^-T (e o a)
            ph[2 * m] = make_double2(row[0], row[1]);
            ph[2 * m + 1] = make_double2(row[2], row[3]);
        }
    }
    if (e >= S) {
        double2 *out = reinterpret_cast<double2 *>(slab + (size_t)(e - S) * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = make_double2(P[2 * i], P[2 * i + 1]);
        if (eig) {  // internal edges are applied in factored form: only exp(lambda t) is read back
            out[0] = make_double2(ex[0], ex[1]);
            out[1] = make_double2(ex[2], ex[3]);
        }
    } else {
        double2 *tt = reinterpret_cast<double2 *>(slab + (size_t)(S - 1) * 16 + (size_t)e * 20);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tt[j] = make_double2(P[0 + j], P[4 + j]);
            tt[5 + j] = make_double2(P[8 + j], P[12 + j]);
        }
        double rs[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rs[i] = ((P[i * 4 + 0] + P[i * 4 + 1]) + P[i * 4 + 2]) + P[i * 4 + 3];
        tt[4] = make_double2(rs[0], rs[1]);
        tt[9] = make_double2(rs[2], rs[3]);
    }
}

// ------------------------------------------------------------------------------------------
// Main kernel: one thread = one site pattern; it walks the tree once per rate category.
// ------------------------------------------------------------------------------------------
struct PruneArgs {
    const uint8_t *tipmask;
    int64_t ld_mask;
    const double *weights;
    const int4 *ops;
    int ops_batch;  // 1 => shared topology
    const double *tab;
    double2 *scratch;
    int *cscratch;    // rescaling counters, same indexing as scratch (SCALE only)
    double *ll_part;  // [n_items]
    double *g_part;   // [n_items][E]
    double *gm_part;  // [n_items][K][nwarp][E][16]
    double *gpi_part; // [n_items][nwarp][4]
    int B, S, L, K;
    int n_tiles, n_items, tile_patterns;
    int mix;
    int pmode;  // how internal-edge transition matrices are applied (PM_*)
    int ch_max; // cherries per tree whose messages are tabulated in the slab
    int *next_item;  // work counter (zeroed before the launch): items beyond the first gridDim.x are claimed dynamically
};

#ifndef DODO_PRUNE_FWD_UNROLL
#define DODO_PRUNE_FWD_UNROLL 1
#endif
constexpr int FWD_UNROLL = DODO_PRUNE_FWD_UNROLL;  // forward join loop (2: the loop-carried result needs no register copies)
#ifndef DODO_PRUNE_NT
#define DODO_PRUNE_NT 256
#endif
constexpr int NT = DODO_PRUNE_NT;  // threads per CTA = site patterns per work item (a multiple of 32)
#ifndef DODO_PRUNE_CTAS
#define DODO_PRUNE_CTAS 3
#endif

// Shared memory.  SMALL (tables and tip selectors resident): tables | gacc | fx | red | cx | ops | sel.
// LARGE (big S: tables and masks stay in global memory, read through L1): gacc | fx | red | cx | ops.
// (LARGE launches without a branch-length gradient carry no per-warp accumulators: at 200 taxa that is 25 KB per CTA
// which go to L1 instead, where the transition tables of the three resident CTAs -- 57 KB each -- have to live)
__host__ __device__ inline size_t prune_smem_dbl(int S, int K, bool small, int np, int chmax, int gm = 0, int nograd = 0) {
    const size_t nwarp = NT / 32, E = 2 * (size_t)S - 2;
    size_t dbl = ((small || !nograd) ? nwarp * E : 0) + (size_t)K * NT * np + nwarp + (gm ? nwarp * 20 : 0) +
                 (small ? tab_doubles(S, chmax, gm) : 0);
    return (dbl + 1) & ~(size_t)1;  // keep what follows 16-byte aligned
}
__host__ __device__ inline size_t prune_smem_bytes(int S, int K, bool small, int np, int chmax, int gm = 0, int nograd = 0) {
    return prune_smem_dbl(S, K, small, np, chmax, gm, nograd) * sizeof(double) + (size_t)K * NT * np * sizeof(int) +
           (size_t)(S + 2) * sizeof(int4) + (small ? (size_t)S * NT * np : 0);
}

// Table pointers.  With the tables in global memory (LARGE) a table pointer is an ordinary pointer.
// With resident tables (SMALL) it carries a 32-bit shared-window address (see sptr()) and is only ever
// dereferenced through ld.shared below: addressed from the `smd` symbol, every lane-dependent access
// re-derived the window base from a special register (S2R + MOV + LEA per access, ~4 per join).
__device__ __forceinline__ unsigned lds_u8(unsigned shared_addr) {
    unsigned r;
    asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(shared_addr));
    return r;
}
__device__ __forceinline__ const double *sptr(unsigned shared_addr) {
    return reinterpret_cast<const double *>((size_t)shared_addr);
}
template <int OFF>
__device__ __forceinline__ double2 lds2(const double *p) {  // two doubles at p + OFF (shared-window address)
    double2 r;
    asm("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(r.x), "=d"(r.y) : "r"((unsigned)(size_t)p), "n"(OFF * 8));
    return r;
}
template <int OFF>
__device__ __forceinline__ double lds1(const double *p) {
    double r;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(r) : "r"((unsigned)(size_t)p), "n"(OFF * 8));
    return r;
}
template <bool SMALL, int OFF = 0>
__device__ __forceinline__ V4 ld4t(const double *p) {  // four doubles at p + OFF
    double2 x, y;
    if (SMALL) { x = lds2<OFF>(p); y = lds2<OFF + 2>(p); }
    else { x = ldg2(p + OFF); y = ldg2(p + OFF + 2); }
    return V4{x.x, x.y, y.x, y.y};
}

// 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100), L1-bypassing: one message per access
__device__ __forceinline__ V4 ldg256(const V4 *p) {
    V4 r;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef DODO_PRUNE_PAIR_REDUCE
#define DODO_PRUNE_PAIR_REDUCE 1
#endif
#ifndef DODO_PF_L2_DIST
#define DODO_PF_L2_DIST 3
#endif
#ifndef DODO_MASK_NOALLOC
#define DODO_MASK_NOALLOC 1
#endif
#ifndef DODO_PF_L1
#define DODO_PF_L1 0
#endif
// backward-sweep message load: through L1 when an L1 prefetch was issued one join ahead
__device__ __forceinline__ V4 ldg256_bwd(const V4 *p) {
#if DODO_PF_L1
    V4 r;
    asm volatile("ld.global.ca.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
#else
    return ldg256(p);
#endif
}
__device__ __forceinline__ void stg256(V4 *p, const V4 &v) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}

// y = P x / y = P^T x with P (row-major, warp-uniform address -> broadcast)
template <bool SMALL>
__device__ __forceinline__ V4 matvec_t(const double *P, const V4 &x) {
    V4 r0 = ld4t<SMALL>(P), r1 = ld4t<SMALL, 4>(P), r2 = ld4t<SMALL, 8>(P), r3 = ld4t<SMALL, 12>(P);
    V4 y;
    y.a = fma(r0.d, x.d, fma(r0.c, x.c, fma(r0.b, x.b, r0.a * x.a)));
    y.b = fma(r1.d, x.d, fma(r1.c, x.c, fma(r1.b, x.b, r1.a * x.a)));
    y.c = fma(r2.d, x.d, fma(r2.c, x.c, fma(r2.b, x.b, r2.a * x.a)));
    y.d = fma(r3.d, x.d, fma(r3.c, x.c, fma(r3.b, x.b, r3.a * x.a)));
    return y;
}
template <bool SMALL>
__device__ __forceinline__ V4 matvecT_t(const double *P, const V4 &x) {
    V4 r = ld4t<SMALL>(P);
    V4 y{r.a * x.a, r.b * x.a, r.c * x.a, r.d * x.a};
    r = ld4t<SMALL, 4>(P);
    y.a = fma(r.a, x.b, y.a), y.b = fma(r.b, x.b, y.b), y.c = fma(r.c, x.b, y.c), y.d = fma(r.d, x.b, y.d);
    r = ld4t<SMALL, 8>(P);
    y.a = fma(r.a, x.c, y.a), y.b = fma(r.b, x.c, y.b), y.c = fma(r.c, x.c, y.c), y.d = fma(r.d, x.c, y.d);
    r = ld4t<SMALL, 12>(P);
    y.a = fma(r.a, x.d, y.a), y.b = fma(r.b, x.d, y.b), y.c = fma(r.c, x.d, y.c), y.d = fma(r.d, x.d, y.d);
    return y;
}

// tip selector byte: low nibble = IUPAC mask; high nibble = row of Ttip (0..3 one-hot state, 4 = all
// four states) or 7 = any other combination (handled by summing columns)
__device__ __forceinline__ unsigned encode_sel(unsigned m) {
    m &= 15u;
    unsigned j = (m == 15u) ? 4u : ((m != 0u && (m & (m - 1u)) == 0u) ? (unsigned)(__ffs(m) - 1) : 7u);
    return m | (j << 4);
}

// message of a tip = P . (indicator of the states allowed by its mask)
template <bool SMALL, int ROW>
__device__ __forceinline__ V4 tip_row(const double *T) {  // row ROW of the split table (compile-time)
    double2 x, y;
    if (SMALL) { x = lds2<2 * ROW>(T); y = lds2<10 + 2 * ROW>(T); }
    else { x = ldg2(T + 2 * ROW); y = ldg2(T + 10 + 2 * ROW); }
    return V4{x.x, x.y, y.x, y.y};
}
template <bool SMALL>
__device__ __forceinline__ V4 tip_message(const double *T /*[2][5][2]*/, unsigned sel) {
    const unsigned j = sel >> 4;
    const double *Tj = T + 2u * (j & 3u) + 2u * (j & 4u);  // j = 7 reads row 7 of the 10: replaced below
    double2 x, y;
    if (SMALL) { x = lds2<0>(Tj); y = lds2<10>(Tj); }
    else { x = ldg2(T + 2u * min(j, 4u)); y = ldg2(T + 10 + 2u * min(j, 4u)); }
    V4 v{x.x, x.y, y.x, y.y};
    if (__any_sync(0xffffffffu, j == 7u)) {
        if (j == 7u) {  // two/three-state codes: add the columns in state order (mask 0: nothing allowed)
            v = V4{0, 0, 0, 0};
            if (sel & 1u) { V4 c = tip_row<SMALL, 0>(T); v.a += c.a; v.b += c.b; v.c += c.c; v.d += c.d; }
            if (sel & 2u) { V4 c = tip_row<SMALL, 1>(T); v.a += c.a; v.b += c.b; v.c += c.c; v.d += c.d; }
            if (sel & 4u) { V4 c = tip_row<SMALL, 2>(T); v.a += c.a; v.b += c.b; v.c += c.c; v.d += c.d; }
            if (sel & 8u) { V4 c = tip_row<SMALL, 3>(T); v.a += c.a; v.b += c.b; v.c += c.c; v.d += c.d; }
        }
    }
    return v;
}

// sum x over the warp into lane 0 and y into lane 16 with 5 shuffles (fixed order)
__device__ __forceinline__ double warp_sum2(double x, double y, unsigned lane) {
    const bool hi = lane & 16u;
    double keep = hi ? y : x, send = hi ? x : y;
    keep += shfl_xor_d(send, 16);
    keep += shfl_xor_d(keep, 8);
    keep += shfl_xor_d(keep, 4);
    keep += shfl_xor_d(keep, 2);
    keep += shfl_xor_d(keep, 1);
    return keep;
}

// four sums at once: v0..v3 end in lanes 0, 8, 16, 24 with 6 shuffles (fixed order) -- the edge
// gradients of two consecutive joins share one butterfly: 3 shuffles and a 5-step dependent chain
// per join instead of 5 and 5
__device__ __forceinline__ double warp_sum4(double v0, double v1, double v2, double v3, unsigned lane) {
    const bool h16 = lane & 16u, h8 = lane & 8u;
    double k0 = h16 ? v2 : v0, k1 = h16 ? v3 : v1;
    const double s0 = h16 ? v0 : v2, s1 = h16 ? v1 : v3;
    k0 += shfl_xor_d(s0, 16);
    k1 += shfl_xor_d(s1, 16);
    double r = h8 ? k1 : k0;
    const double sn = h8 ? k0 : k1;
    r += shfl_xor_d(sn, 8);
    r += shfl_xor_d(r, 4);
    r += shfl_xor_d(r, 2);
    r += shfl_xor_d(r, 1);
    return r;
}

// sum over the warp of the 4x4 outer products u (x) v; entry (i,j) ends in lane 2*(4i+j) and is
// written to dst[4i+j].  Halving butterfly: 16 double shuffles instead of 80, fixed order.
__device__ __forceinline__ void reduce16_store(const double (&x)[16], double *dst, unsigned lane);
__device__ __forceinline__ void outer_reduce_store(const V4 &u, const V4 &v, double *dst, unsigned lane) {
    double x[16];
    const double uu[4] = {u.a, u.b, u.c, u.d}, vv[4] = {v.a, v.b, v.c, v.d};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[i * 4 + j] = uu[i] * vv[j];
    reduce16_store(x, dst, lane);
}
// sum over the warp of 16 values per lane: entry i ends in lane 2 i and is written to dst[i]
__device__ __forceinline__ void reduce16_store(const double (&x)[16], double *dst, unsigned lane) {
    double y[8];
    {
        const bool hi = lane & 16u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double keep = hi ? x[i + 8] : x[i], send = hi ? x[i] : x[i + 8];
            y[i] = keep + shfl_xor_d(send, 16);
        }
    }
    double z[4];
    {
        const bool hi = lane & 8u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double keep = hi ? y[i + 4] : y[i], send = hi ? y[i] : y[i + 4];
            z[i] = keep + shfl_xor_d(send, 8);
        }
    }
    double t[2];
    {
        const bool hi = lane & 4u;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double keep = hi ? z[i + 2] : z[i], send = hi ? z[i] : z[i + 2];
            t[i] = keep + shfl_xor_d(send, 4);
        }
    }
    double r;
    {
        const bool hi = lane & 2u;
        double keep = hi ? t[1] : t[0], send = hi ? t[0] : t[1];
        r = keep + shfl_xor_d(send, 2);
    }
    r += shfl_xor_d(r, 1);
    if (!(lane & 1u)) dst[lane >> 1] = r;
}

// Power-of-two rescaling (exact: mantissas untouched, so results are bit-identical to the unscaled
// computation wherever that one stays in range).  true value = stored * 2^(-256 * count).
constexpr double SC_LO = 0x1p-256, SC_LO2 = 0x1p-512, SC_UP = 0x1p256, SC_UP2 = 0x1p512;
__device__ __forceinline__ void scale_up(V4 &v, int &count) {  // keep max(v) >= 2^-256 (v >= 0)
    const double m = fmax(fmax(fabs(v.a), fabs(v.b)), fmax(fabs(v.c), fabs(v.d)));
    if (m < SC_LO && m > 0.0) {
        const bool two = m < SC_LO2;
        const double f = two ? SC_UP2 : SC_UP;
        v.a *= f; v.b *= f; v.c *= f; v.d *= f;
        count += two ? 2 : 1;
    }
}
__device__ __forceinline__ void scale_both(V4 &v, int &count) {  // upper messages can also grow
    const double m = fmax(fmax(fabs(v.a), fabs(v.b)), fmax(fabs(v.c), fabs(v.d)));
    if (m < SC_LO && m > 0.0) {
        v.a *= SC_UP; v.b *= SC_UP; v.c *= SC_UP; v.d *= SC_UP;
        count += 1;
    } else if (m > SC_UP) {
        v.a *= SC_LO; v.b *= SC_LO; v.c *= SC_LO; v.d *= SC_LO;
        count -= 1;
    }
}
__device__ __forceinline__ double pow2_256(int d) {  // 2^(256 d), flushing to 0 / saturating
    d = max(-5, min(3, d));
    return scalbn(1.0, 256 * d);
}

// P x / P^T x for NP patterns at once: every row of P is loaded once and applied to all of them
// JC69: P = (a-b) I + b J is symmetric with two distinct entries, so P x = (a-b) x + b sum(x): one
// 16-byte load and 8 flops instead of eight loads and 16 flops (warp-uniform runtime switch).
template <bool SMALL>
__device__ __forceinline__ double2 load_ab(const double *P) {
    return SMALL ? lds2<0>(P) : ldg2(P);
}
template <int NP>
__device__ __forceinline__ void matvec_jc_ab(const double2 ab, const V4 (&x)[NP], V4 (&y)[NP]) {
    const double amb = ab.x - ab.y;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const V4 v = x[q];
        const double bs = ab.y * ((v.a + v.b) + (v.c + v.d));
        y[q] = V4{fma(amb, v.a, bs), fma(amb, v.b, bs), fma(amb, v.c, bs), fma(amb, v.d, bs)};
    }
}
template <bool SMALL, int NP>
__device__ __forceinline__ void matvec_jc(const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    matvec_jc_ab<NP>(load_ab<SMALL>(P), x, y);
}

// Eigen form (device-formed GTR matrices): P = U diag(e) V with U, V = U^-1 in the kernel-parameter
// (constant) bank and only e = exp(lambda r t) read per edge: P x = U (e * (V x)).  32 FMAs with
// constant operands instead of 16 FMAs behind eight shared-memory loads -- the kernel is bound by
// shared-memory wavefronts, not by the fp64 pipe.  Agrees with (U diag(e) V) x to rounding (1e-15).
template <bool SMALL, int NP>
__device__ __forceinline__ void matvec_eig(const ModelDev &md, const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    const V4 e = ld4t<SMALL>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const V4 v = x[q];
        const double z0 = e.a * fma(md.Uinv[3], v.d, fma(md.Uinv[2], v.c, fma(md.Uinv[1], v.b, md.Uinv[0] * v.a)));
        const double z1 = e.b * fma(md.Uinv[7], v.d, fma(md.Uinv[6], v.c, fma(md.Uinv[5], v.b, md.Uinv[4] * v.a)));
        const double z2 = e.c * fma(md.Uinv[11], v.d, fma(md.Uinv[10], v.c, fma(md.Uinv[9], v.b, md.Uinv[8] * v.a)));
        const double z3 = e.d * fma(md.Uinv[15], v.d, fma(md.Uinv[14], v.c, fma(md.Uinv[13], v.b, md.Uinv[12] * v.a)));
        y[q].a = fma(md.U[3], z3, fma(md.U[2], z2, fma(md.U[1], z1, md.U[0] * z0)));
        y[q].b = fma(md.U[7], z3, fma(md.U[6], z2, fma(md.U[5], z1, md.U[4] * z0)));
        y[q].c = fma(md.U[11], z3, fma(md.U[10], z2, fma(md.U[9], z1, md.U[8] * z0)));
        y[q].d = fma(md.U[15], z3, fma(md.U[14], z2, fma(md.U[13], z1, md.U[12] * z0)));
    }
}
// P^T x = V^T (e * (U^T x))
template <bool SMALL, int NP>
__device__ __forceinline__ void matvecT_eig(const ModelDev &md, const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    const V4 e = ld4t<SMALL>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const V4 v = x[q];
        const double z0 = e.a * fma(md.U[12], v.d, fma(md.U[8], v.c, fma(md.U[4], v.b, md.U[0] * v.a)));
        const double z1 = e.b * fma(md.U[13], v.d, fma(md.U[9], v.c, fma(md.U[5], v.b, md.U[1] * v.a)));
        const double z2 = e.c * fma(md.U[14], v.d, fma(md.U[10], v.c, fma(md.U[6], v.b, md.U[2] * v.a)));
        const double z3 = e.d * fma(md.U[15], v.d, fma(md.U[11], v.c, fma(md.U[7], v.b, md.U[3] * v.a)));
        y[q].a = fma(md.Uinv[12], z3, fma(md.Uinv[8], z2, fma(md.Uinv[4], z1, md.Uinv[0] * z0)));
        y[q].b = fma(md.Uinv[13], z3, fma(md.Uinv[9], z2, fma(md.Uinv[5], z1, md.Uinv[1] * z0)));
        y[q].c = fma(md.Uinv[14], z3, fma(md.Uinv[10], z2, fma(md.Uinv[6], z1, md.Uinv[2] * z0)));
        y[q].d = fma(md.Uinv[15], z3, fma(md.Uinv[11], z2, fma(md.Uinv[7], z1, md.Uinv[3] * z0)));
    }
}

template <bool SMALL, int NP>
__device__ __forceinline__ void matvec_n(const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    const V4 r0 = ld4t<SMALL>(P), r1 = ld4t<SMALL, 4>(P), r2 = ld4t<SMALL, 8>(P), r3 = ld4t<SMALL, 12>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const V4 v = x[q];
        y[q].a = fma(r0.d, v.d, fma(r0.c, v.c, fma(r0.b, v.b, r0.a * v.a)));
        y[q].b = fma(r1.d, v.d, fma(r1.c, v.c, fma(r1.b, v.b, r1.a * v.a)));
        y[q].c = fma(r2.d, v.d, fma(r2.c, v.c, fma(r2.b, v.b, r2.a * v.a)));
        y[q].d = fma(r3.d, v.d, fma(r3.c, v.c, fma(r3.b, v.b, r3.a * v.a)));
    }
}
template <bool SMALL, int NP>
__device__ __forceinline__ void matvecT_n(const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    V4 r = ld4t<SMALL>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) y[q] = V4{r.a * x[q].a, r.b * x[q].a, r.c * x[q].a, r.d * x[q].a};
    r = ld4t<SMALL, 4>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        y[q].a = fma(r.a, x[q].b, y[q].a), y[q].b = fma(r.b, x[q].b, y[q].b);
        y[q].c = fma(r.c, x[q].b, y[q].c), y[q].d = fma(r.d, x[q].b, y[q].d);
    }
    r = ld4t<SMALL, 8>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        y[q].a = fma(r.a, x[q].c, y[q].a), y[q].b = fma(r.b, x[q].c, y[q].b);
        y[q].c = fma(r.c, x[q].c, y[q].c), y[q].d = fma(r.d, x[q].c, y[q].d);
    }
    r = ld4t<SMALL, 12>(P);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        y[q].a = fma(r.a, x[q].d, y[q].a), y[q].b = fma(r.b, x[q].d, y[q].b);
        y[q].c = fma(r.c, x[q].d, y[q].c), y[q].d = fma(r.d, x[q].d, y[q].d);
    }
}

// ------------------------------------------------------------------------------------------
// Learnable reversible model (MODE 3): per edge c with upper message u (parent side) and message msg = P_c L_c,
//   a = U^T u,  bt = U^-1 msg = e o (U^-1 L_c)                       (e_m = exp(lam_m tau), tau = r_k t_c)
//   d lnL / d t_c        += r_k sum_m lam_m a_m bt_m                  (= r_k u . Q msg)
//   W_mn                 += Phi'_mn a_m bt_n                          (Phi' from the slab, see tab_doubles)
// so that d lnL / d Q = U^-T W U^T (Daleckii-Krein: d exp(Q tau) = U [Phi o (U^-1 dQ U)] U^-1 with
// Phi_mn = e_n Phi'_mn).  W is accumulated PER THREAD over all edges and categories of the walk and reduced
// over the warp once per work item -- the explicit-matrix mode (GMATS) reduces 16 values per edge instead.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ V4 matvec_cT(const double *M, const V4 &v) {  // M^T v, M in the constant bank
    return V4{fma(M[12], v.d, fma(M[8], v.c, fma(M[4], v.b, M[0] * v.a))),
              fma(M[13], v.d, fma(M[9], v.c, fma(M[5], v.b, M[1] * v.a))),
              fma(M[14], v.d, fma(M[10], v.c, fma(M[6], v.b, M[2] * v.a))),
              fma(M[15], v.d, fma(M[11], v.c, fma(M[7], v.b, M[3] * v.a)))};
}
__device__ __forceinline__ V4 matvec_c(const double *M, const V4 &v) {  // M v
    return V4{fma(M[3], v.d, fma(M[2], v.c, fma(M[1], v.b, M[0] * v.a))),
              fma(M[7], v.d, fma(M[6], v.c, fma(M[5], v.b, M[4] * v.a))),
              fma(M[11], v.d, fma(M[10], v.c, fma(M[9], v.b, M[8] * v.a))),
              fma(M[15], v.d, fma(M[14], v.c, fma(M[13], v.b, M[12] * v.a)))};
}
// The eigenvalue 0 of a rate matrix sits LAST here (the host permutes the eigen-system if need be): its right
// eigenvector is the constant vector, so column 3 of W only ever multiplies the ROW SUMS of a perturbation dQ
// (d lnL / d Q = U^-T W U^T, U^T[3, :] = const), which are zero for every rate matrix.  That column is not
// accumulated: 12 accumulators instead of 16, bt needs three components, and r_k lam_3 = 0 drops out of d/dt.
template <bool SMALL>
__device__ __forceinline__ double model_edge(const ModelDev &md, const double *Phi, const V4 &u, const V4 &msg,
                                             const double (&rl)[4], const double sf, double (&W)[12], V4 &z) {
    V4 a = matvec_cT(md.U, u);
    const V4 a0 = a;
    a = V4{a.a * sf, a.b * sf, a.c * sf, a.d * sf};  // sf: exact power of two (1 without rescaling)
    const double *M = md.Uinv;
    const double b0 = fma(M[3], msg.d, fma(M[2], msg.c, fma(M[1], msg.b, M[0] * msg.a)));
    const double b1 = fma(M[7], msg.d, fma(M[6], msg.c, fma(M[5], msg.b, M[4] * msg.a)));
    const double b2 = fma(M[11], msg.d, fma(M[10], msg.c, fma(M[9], msg.b, M[8] * msg.a)));
    // z = e o (U^T u): the downward message of an internal edge is P^T u = U^-T z -- no second pass over a table of P
    V4 ph = ld4t<SMALL, 0>(Phi);
    W[0] = fma(ph.a, a.a * b0, W[0]); W[1] = fma(ph.b, a.a * b1, W[1]); W[2] = fma(ph.c, a.a * b2, W[2]);
    z.a = ph.d * a0.a;
    ph = ld4t<SMALL, 4>(Phi);
    W[3] = fma(ph.a, a.b * b0, W[3]); W[4] = fma(ph.b, a.b * b1, W[4]); W[5] = fma(ph.c, a.b * b2, W[5]);
    z.b = ph.d * a0.b;
    ph = ld4t<SMALL, 8>(Phi);
    W[6] = fma(ph.a, a.c * b0, W[6]); W[7] = fma(ph.b, a.c * b1, W[7]); W[8] = fma(ph.c, a.c * b2, W[8]);
    z.c = ph.d * a0.c;
    ph = ld4t<SMALL, 12>(Phi);
    W[9] = fma(ph.a, a.d * b0, W[9]); W[10] = fma(ph.b, a.d * b1, W[10]); W[11] = fma(ph.c, a.d * b2, W[11]);
    z.d = ph.d * a0.d;
    return fma(rl[2], a.c * b2, fma(rl[1], a.b * b1, rl[0] * (a.a * b0)));
}

enum { PM_GENERIC = 0, PM_JC = 1, PM_EIG = 2 };
// PM_EIG measured slower at 64 taxa x 10k x 128 x 4 (5.7 ms against 4.7 ms): U and U^-1 do not fit the
// uniform register file next to Q, pi and the rates, so every use reloads them (+50 % instructions).
#ifndef DODO_PRUNE_EIG
#define DODO_PRUNE_EIG 0
#endif
// y = P x (TR = false) or P^T x (TR = true) in the way the launch selected (compile-time PM)
template <bool SMALL, int NP, bool TR>
__device__ __forceinline__ void apply_P(const int pmode, const ModelDev &md, const double *P, const V4 (&x)[NP], V4 (&y)[NP]) {
    if (pmode == PM_JC) matvec_jc<SMALL, NP>(P, x, y);
    else if (pmode == PM_EIG) { if (TR) matvecT_eig<SMALL, NP>(md, P, x, y); else matvec_eig<SMALL, NP>(md, P, x, y); }
    else { if (TR) matvecT_n<SMALL, NP>(P, x, y); else matvec_n<SMALL, NP>(P, x, y); }
}

// ------------------------------------------------------------------------------------------
// Cherry tables: for the first chmax cherries of every tree (schedule order) the message
// P_n (T_l[i] * T_r[j]) for each of the 5 x 5 pairs of tip-table rows, computed with exactly the
// operations the pruning kernel would use, so that a lookup and a recomputation are interchangeable
// bit for bit.  One thread per (sample, category, cherry, pair).
// ------------------------------------------------------------------------------------------
__global__ void cherry_kernel(const int4 *__restrict__ ops, const int *__restrict__ chop, int ops_batch, int B, int S,
                              int K, int chmax, int gm, int pmode, ModelDev md, double *__restrict__ tab) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * K * chmax * 25) return;
    const int pair = gid % 25;
    const int c = (gid / 25) % chmax;
    const int k = (gid / (25 * (size_t)chmax)) % K;
    const int b = gid / (25 * (size_t)chmax * K);
    const int ob = ops_batch == 1 ? 0 : b;
    const int pos = chop[(size_t)ob * (S / 2) + c];
    if (pos < 0) return;
    const int4 o = ops[(size_t)ob * (S - 1) + pos];  // {left tip, right tip, node, flags}
    double *slab = tab + ((size_t)b * K + k) * tab_doubles(S, chmax, gm);
    const double *P = slab + (size_t)(o.z - S) * 16;
    const double *Tl = slab + (size_t)(S - 1) * 16 + (size_t)o.x * 20 + (pair / 5) * 2;  // split rows: see Ttip
    const double *Tr = slab + (size_t)(S - 1) * 16 + (size_t)o.y * 20 + (pair % 5) * 2;
    V4 x[1] = {V4{Tl[0] * Tr[0], Tl[1] * Tr[1], Tl[10] * Tr[10], Tl[11] * Tr[11]}}, y[1];
    apply_P<false, 1, false>(pmode, md, P, x, y);
    double *out = slab + (size_t)(S - 1) * 16 + (size_t)S * 20 + (size_t)c * 100 + cherry_entry(pair / 5, pair % 5);
    out[0] = y[0].a; out[25] = y[0].b; out[50] = y[0].c; out[75] = y[0].d;
}

// MODE 0: lnL only; 1: lnL + d lnL/d blens; 2: lnL + d lnL/d P (per edge and category) + d lnL/d pi;
// 3: lnL + d lnL/d blens + the 4 x 4 matrix W of the model chain rule (see model_edge) + d lnL/d pi
// NP: site patterns per thread (1 or 2).  With NP = 2 the schedule decode, the branches, the address
// arithmetic and -- above all -- the warp-wide fills of registers with the (warp-uniform) entries of P
// are shared by two independent patterns, which also doubles the instruction-level parallelism.
//
// Register conventions of the two sweeps (they keep the loops free of copies):
//   forward : Bv holds "the other operand": the previous join's message when child b is internal
//             (the schedule guarantees b is then exactly the previous join), else b's tip message.
//   backward: Wv holds the upper message of the current join: left there by the previous join when
//             this node is its b-child, else (a-child) fetched from scratch.
// Latency is hidden by the resident warps plus an L2 prefetch three joins ahead in the backward
// sweep; a register-level software pipeline was measured slower (more instructions).
template <int MODE, bool SMALL, bool SCALE, int NP, int PM>
__global__ void __launch_bounds__(NT, (NP == 1 && MODE != 3) ? DODO_PRUNE_CTAS : 2) prune_kernel(PruneArgs A, ModelDev md) {
    constexpr bool GMODEL = MODE == 3;  // the branch-length sweep + per-thread accumulation of W (16 more doubles: 2 CTAs/SM)
    constexpr bool GRAD = MODE == 1 || GMODEL;
    constexpr bool GMATS = MODE == 2;
    constexpr int GM = GMODEL ? 1 : 0;
    static_assert(!(GMODEL && NP != 1), "the model gradient handles one pattern per thread");
    static_assert(!(GMATS && SCALE), "GMATS is built without rescaling");
    static_assert(!(GMATS && NP != 1), "GMATS handles one pattern per thread");
    extern __shared__ __align__(16) double smd[];  // indexed from the symbol: keeps LDS addressing direct
    const unsigned tid = threadIdx.x;
    constexpr unsigned nthr = NT;
    constexpr unsigned npt = NT * NP;  // patterns per work item
    const unsigned lane = tid & 31u;
    const unsigned warp = tid >> 5;
    constexpr unsigned nwarp = NT / 32;
    const int S = A.S, K = A.K, E = 2 * S - 2, nops = S - 1;
    constexpr int pmode = PM;
    const int ch_max = SMALL && !SCALE ? A.ch_max : 0;  // tabulated cherries (tables resident only)
    // b-child cherries are recomputed rather than stored
#ifndef DODO_PRUNE_RECOMP_LARGE
#define DODO_PRUNE_RECOMP_LARGE 1
#endif
    constexpr bool RECOMP = SMALL || DODO_PRUNE_RECOMP_LARGE;
    const unsigned tabn = (unsigned)tab_doubles(S, A.ch_max, GM);
    // offsets in doubles: [tables (SMALL)] gacc[nwarp][E] fx[K][npt] red[nwarp] wred[nwarp][20] (GMODEL) | cx | ops | sel
    const unsigned o_gacc = SMALL ? tabn : 0u;
    const unsigned o_fx = o_gacc + ((SMALL || GRAD) ? nwarp * (unsigned)E : 0u);
    const unsigned o_red = o_fx + (unsigned)K * npt;
    const unsigned o_wred = o_red + nwarp;                                          // per-warp W (16) + d/d pi (4)
    const unsigned o_cx = (unsigned)prune_smem_dbl(S, K, SMALL, NP, A.ch_max, GM, GRAD ? 0 : 1);  // int [K][npt]
    const unsigned o_ops = o_cx + ((unsigned)K * npt) / 2u;                         // int4 [S+2]
    const unsigned o_sel = o_ops + 2u * (unsigned)(S + 2);                          // u8 [S][npt]
    const unsigned o_T = (unsigned)(S - 1) * 16u;                                   // Ttip inside the slab
    const unsigned o_C = o_T + (unsigned)S * 20u;                                   // Cher inside the slab
    const unsigned o_X = o_C + (unsigned)A.ch_max * 100u;                           // Phi' inside the slab (GMODEL)
#define GACC(i) smd[o_gacc + (i)]
#ifndef DODO_GACC_SHARED
#define DODO_GACC_SHARED 1
#endif
#if DODO_GACC_SHARED
    // the per-join accumulation goes through explicit shared-window addresses (ld/st.shared from `sbase`): addressed
    // from the `smd` symbol every lane-dependent access re-derives the window base (S2UR + conversions per join)
#define GACC_ADD(i, v)                                                              \
    do {                                                                            \
        const unsigned a__ = sbase + (o_gacc + (i)) * 8u;                            \
        double r__;                                                                 \
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r__) : "r"(a__) : "memory");  \
        r__ += (v);                                                                 \
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a__), "d"(r__) : "memory");    \
    } while (0)
#else
#define GACC_ADD(i, v) GACC(i) += (v)
#endif
#define FX(k, q) smd[o_fx + ((unsigned)(k) * NP + (q)) * nthr + tid]
#define RED(i) smd[o_red + (i)]
#define CX(k, q) (reinterpret_cast<int *>(&smd[o_cx])[((unsigned)(k) * NP + (q)) * nthr + tid])
#define OPS(i) (reinterpret_cast<const int4 *>(&smd[o_ops])[(i)])
#define OPSW(i) (reinterpret_cast<int4 *>(&smd[o_ops])[(i)])
#define SELW(t, q) (reinterpret_cast<uint8_t *>(&smd[o_sel])[((unsigned)(t) * NP + (q)) * nthr + tid])
#define SEL(t, q) lds_u8(sel_base + ((unsigned)(t) * NP + (q)) * nthr)
#define SCR(node, q) (scr + ((unsigned)(node) * NP + (q)) * nthr)
#define CSCR(node, q) (cscr + ((unsigned)(node) * NP + (q)) * nthr)
    // shared-window address of smd, made opaque so that it lives in a register instead of being
    // re-derived from the special register at every lane-dependent access
    unsigned sbase = (unsigned)__cvta_generic_to_shared(smd);
    asm volatile("" : "+r"(sbase));
    const unsigned sel_base = sbase + o_sel * 8u + tid;
    if (GRAD) {
        for (unsigned i = tid; i < nwarp * (unsigned)E; i += nthr) GACC(i) = 0.0;
    }
    // scratch ring of this CTA: [k][node][pattern] x 32 B (+ 4 B rescaling counters)
    V4 *scr0 = reinterpret_cast<V4 *>(A.scratch) + (size_t)blockIdx.x * (size_t)K * (S - 1) * npt + tid;
    int *cscr0 = SCALE ? A.cscratch + (size_t)blockIdx.x * (size_t)K * (S - 1) * npt + tid : nullptr;
    const unsigned gw = warp * (unsigned)E;

    // Every CTA starts with item blockIdx.x and claims further ones from a counter: with a fixed stride the CTAs that
    // get ceil(n_items / grid) items finish a whole item after the others (5120 items on 444 CTAs: 12 against 11).
    // Results do not depend on who processes an item (per-item partial sums, reduced in item order).
    __shared__ int s_next_item;
#pragma unroll 1
    for (int item = blockIdx.x; item < A.n_items;) {
        const int b = item / A.n_tiles;
        const int tile = item - b * A.n_tiles;
        double wt[NP];
        const uint8_t *tm[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int p = tile * (int)npt + q * (int)nthr + (int)tid;
            const bool valid = p < A.L;
            wt[q] = valid ? __ldg(A.weights + p) : 0.0;
            tm[q] = A.tipmask + (valid ? p : A.L - 1);
        }
        const double *tab_b = A.tab + (size_t)b * K * tabn;
        double Wm[12];  // GMODEL: this thread's share of W (columns 0..2), over every edge and category of the walk
#pragma unroll
        for (int i = 0; i < 12; ++i) Wm[i] = 0.0;
        __syncthreads();
        if (GMODEL && lane < 4) smd[o_wred + warp * 20u + 16u + lane] = 0.0;  // d lnL / d pi of this warp
        {
            // schedule of this sample, re-encoded: child code < 0 -> tip ~code, else internal index
            const int4 *ops_b = A.ops + (size_t)(A.ops_batch == 1 ? 0 : b) * nops;
            for (int i = tid; i < nops + 2; i += nthr) {
                int4 o = make_int4(~0, ~0, 0, 4);  // entries 0 and nops+1: harmless sentinels
                if (i >= 1 && i <= nops) {
                    o = __ldg(ops_b + (i - 1));
                    o.x = o.x < S ? ~o.x : o.x - S;
                    o.y = o.y < S ? ~o.y : o.y - S;
                    o.z -= S;
                }
                OPSW(i) = o;
            }
            if (SMALL) {  // tip selectors of this thread's patterns (thread-private columns)
                for (int t = 0; t < S; ++t)
#pragma unroll
                    for (int q = 0; q < NP; ++q) SELW(t, q) = (uint8_t)encode_sel(__ldg(tm[q] + (size_t)t * A.ld_mask));
            }
        }
        const double *tabg = tab_b;  // LARGE: tables of the current category in global memory
        auto stage_tables = [&](int k) {
            __syncthreads();  // everyone is done with the previous slab (and the schedule is visible)
            if (SMALL) {
                const double2 *src = reinterpret_cast<const double2 *>(tab_b + (size_t)k * tabn);
                double2 *dst = reinterpret_cast<double2 *>(&smd[0]);
                for (unsigned i = tid; i < tabn / 2; i += nthr) dst[i] = __ldg(src + i);
                __syncthreads();
            } else {
                tabg = tab_b + (size_t)k * tabn;
            }
        };
        auto tab_at = [&](unsigned off) -> const double * { return SMALL ? sptr(sbase + off * 8u) : tabg + off; };
        // LARGE (tip masks in global memory): the mask byte of a tip is fetched ONE JOIN AHEAD of its use (the
        // schedule in shared memory names the tips of the next join), so the table-row load that depends on it
        // does not wait for an L2 round trip first: ncu showed 47 % of all stall samples on that byte at 200
        // taxa (two dependent global loads per tip on the critical path of every join).
        auto load_mask = [&](int code, int q) -> unsigned {  // child code < 0 -> tip ~code
            if (SMALL || code >= 0) return 0u;
#if DODO_MASK_NOALLOC
            // a mask byte is used once per sweep: it streams past L1 (which holds the tables) instead of through it
            unsigned m;
            asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(m) : "l"(tm[q] + (size_t)(~code) * A.ld_mask));
            return m;
#else
            return (unsigned)__ldg(tm[q] + (size_t)(~code) * A.ld_mask);
#endif
        };
        auto tipmsg = [&](int t, int q, unsigned mask) -> V4 {
            const unsigned sel = SMALL ? (unsigned)SEL(t, q) : encode_sel(mask);
            return tip_message<SMALL>(tab_at(o_T + (unsigned)t * 20u), sel);
        };
        // tabulated cherry: one 32-byte row of its 5 x 5 table instead of two tip rows and a mat-vec
        // (warp-uniform outcome; false when the cherry has no table or a lane holds a 2/3-state code)
        auto cherry_lookup = [&](const int4 &oc, int q, V4 &out) -> bool {
            if (!SMALL || SCALE) return false;
            const int c = oc.w >> 8;
            if (c >= ch_max) return false;
            const unsigned ja = (unsigned)SEL(~oc.x, q) >> 4, jb = (unsigned)SEL(~oc.y, q) >> 4;
            if (__any_sync(0xffffffffu, ja == 7u || jb == 7u)) return false;
            const double *e = sptr(sbase + (o_C + (unsigned)c * 100u + cherry_entry(ja, jb)) * 8u);
            out = V4{lds1<0>(e), lds1<25>(e), lds1<50>(e), lds1<75>(e)};
            return true;
        };
        // message of a cherry, recomputed exactly as the forward sweep formed it (oc = the cherry's op)
        auto cherry_msg = [&](const int4 &oc, int q, int &cnt, unsigned mx, unsigned my) -> V4 {
            cnt = 0;
            V4 t;
            if (cherry_lookup(oc, q, t)) return t;
            V4 x[1] = {vmul(tipmsg(~oc.x, q, mx), tipmsg(~oc.y, q, my))}, y[1];
            if (SCALE) scale_up(x[0], cnt);
            apply_P<SMALL, 1, false>(pmode, md, tab_at((unsigned)oc.z * 16u), x, y);
            return y[0];
        };

        // ---------------- forward (post-order), one pass per rate category ----------------
        double fsum[NP], ll = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) fsum[q] = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            stage_tables(k);
            V4 *scr = scr0 + (size_t)k * (S - 1) * npt;
            int *cscr = SCALE ? cscr0 + (size_t)k * (S - 1) * npt : nullptr;
            V4 Bv[NP];
            int bc[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) { Bv[q] = V4{0, 0, 0, 0}; bc[q] = 0; }
            unsigned cmx[NP], cmy[NP];  // masks of the current join's tip children (LARGE; see load_mask)
            {
                const int4 o1 = OPS(1);
#pragma unroll
                for (int q = 0; q < NP; ++q) { cmx[q] = load_mask(o1.x, q); cmy[q] = load_mask(o1.y, q); }
            }
            auto join = [&](const int4 &o, V4 (&Lv)[NP], int (&cnt)[NP]) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    cnt[q] = 0;
                    if (o.x < 0) {
                        Lv[q] = tipmsg(~o.x, q, cmx[q]);
                    } else {
                        Lv[q] = ldg256(SCR(o.x, q));
                        if (SCALE) cnt[q] = __ldcg(CSCR(o.x, q));
                    }
                }
                if (o.y < 0) {  // both children are tips: Bv is free
#pragma unroll
                    for (int q = 0; q < NP; ++q) { Bv[q] = tipmsg(~o.y, q, cmy[q]); bc[q] = 0; }
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    Lv[q] = vmul(Lv[q], Bv[q]);
                    cnt[q] += bc[q];
                    if (SCALE) scale_up(Lv[q], cnt[q]);
                }
            };
#pragma unroll FWD_UNROLL
            for (int i = 1; i < nops; ++i) {
                const int4 o = OPS(i);
                unsigned nmx[NP], nmy[NP];
                if (!SMALL) {
                    const int4 on = OPS(i + 1);
#pragma unroll
                    for (int q = 0; q < NP; ++q) { nmx[q] = load_mask(on.x, q); nmy[q] = load_mask(on.y, q); }
                }
                V4 Lv[NP];
                int cnt[NP];
                bool tabulated = false;
                if (SMALL && !SCALE && (o.w & FL_CHERRY)) {
                    tabulated = true;
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        cnt[q] = 0;
                        tabulated = tabulated && cherry_lookup(o, q, Lv[q]);
                    }
                }
                if (tabulated) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) Bv[q] = Lv[q];
                } else if (pmode == PM_JC) {
                    // issue the (a, b) load before the join: its latency would otherwise sit on the
                    // critical path behind the rescaling branch
                    const double2 ab = load_ab<SMALL>(tab_at((unsigned)o.z * 16u));
                    join(o, Lv, cnt);
                    matvec_jc_ab<NP>(ab, Lv, Bv);
                } else {
                    join(o, Lv, cnt);
                    apply_P<SMALL, NP, false>(pmode, md, tab_at((unsigned)o.z * 16u), Lv, Bv);
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    bc[q] = cnt[q];
                    if (!SMALL) { cmx[q] = nmx[q]; cmy[q] = nmy[q]; }
                }
                if ((MODE != 0 && !(RECOMP && (o.w & FL_NO_STORE))) || (o.w & FL_A_CHILD)) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        stg256(SCR(o.z, q), Bv[q]);
                        if (SCALE) __stcg(CSCR(o.z, q), cnt[q]);
                    }
                }
            }
            V4 rootL[NP];
            int rootc[NP];
            join(OPS(nops), rootL, rootc);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const V4 r = rootL[q];
                const double fk = fma(md.pi[3], r.d, fma(md.pi[2], r.c, fma(md.pi[1], r.b, md.pi[0] * r.a)));
                if (MODE != 0 || SCALE) FX(k, q) = fk;
                if (SCALE) CX(k, q) = rootc[q];
                fsum[q] += fk;
                if (!A.mix) ll += SCALE ? wt[q] * (log(fk) - (double)rootc[q] * (256.0 * 0.6931471805599453)) : wt[q] * log(fk);
            }
        }
        double f[NP];
        int cmin[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            f[q] = fsum[q] / (double)K;
            cmin[q] = 0;
            if (SCALE) {
                cmin[q] = CX(0, q);
                for (int k = 1; k < K; ++k) cmin[q] = min(cmin[q], CX(k, q));
                if (A.mix) {
                    double s = 0.0;
                    for (int k = 0; k < K; ++k) s += FX(k, q) * pow2_256(cmin[q] - CX(k, q));
                    f[q] = s / (double)K;
                }
            }
            if (A.mix) ll += SCALE ? wt[q] * (log(f[q]) - (double)cmin[q] * (256.0 * 0.6931471805599453)) : wt[q] * log(f[q]);
        }
        ll = warp_sum(ll);
        if (lane == 0) RED(warp) = ll;

        // ---------------- backward (the same schedule reversed = pre-order) ----------------
        if (GRAD) {
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                stage_tables(k);
                V4 *scr = scr0 + (size_t)k * (S - 1) * npt;
                int *cscr = SCALE ? cscr0 + (size_t)k * (S - 1) * npt : nullptr;
                const double rate_k = md.rates[k];
                const double rl[4] = {md.rlam[k][0], md.rlam[k][1], md.rlam[k][2], md.rlam[k][3]};
                double sc0 = 0.0;
                V4 Wv[NP];
                int we[NP];  // true w = Wv * 2^(256 we)
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const double sc = A.mix ? wt[q] / (f[q] * (double)K) : wt[q] / FX(k, q);
                    sc0 = sc;
                    Wv[q] = V4{md.pi[0] * sc, md.pi[1] * sc, md.pi[2] * sc, md.pi[3] * sc};
                    we[q] = SCALE ? (A.mix ? cmin[q] : CX(k, q)) : 0;
                }
                // edge gradients of two consecutive joins share one reduction (resident tables only:
                // the large-tree variants have no registers to park a pair in)
#ifndef DODO_GM_PAIR
#define DODO_GM_PAIR 0
#endif
#ifndef DODO_GM_EIG_DOWN  // model-gradient sweep: downward messages as U^-T (e o U^T u), U^T u being there already
#define DODO_GM_EIG_DOWN 1
#endif
                constexpr bool PAIR = DODO_PRUNE_PAIR_REDUCE && SMALL && (!GMODEL || DODO_GM_PAIR);
#ifndef DODO_PRUNE_HOIST
#define DODO_PRUNE_HOIST 0
#endif
                constexpr bool HOIST = (DODO_PRUNE_HOIST & 1) && SMALL && !SCALE && NP == 1 && (MODE == 1 || ((DODO_PRUNE_HOIST & 2) && GMODEL));
                constexpr bool HOIST_B = HOIST && !(DODO_PRUNE_HOIST & 4), HOIST_W = HOIST && !(DODO_PRUNE_HOIST & 8);
                double pend_ga = 0.0, pend_gb = 0.0;
                int pend_x = 0, pend_y = 0;
                bool have_pending = false;
                unsigned cmx[NP], cmy[NP];  // masks of the current join's tip children (LARGE; see load_mask)
                // HOIST: the stored operands of a join (messages of its internal children, the parked upper message of
                // an a-child) are requested in the MIDDLE of the join before it -- right after the last use of that
                // join's own av / bv, into the registers they leave free -- instead of at the top of their own join,
                // where their first use follows at once and the L2 round trip is exposed (11 % + 4 % of all stall
                // samples on those two uses, profiles/r2/prune.txt).
                V4 nav[NP], nbv[NP];
                {
                    const int4 o1 = OPS(nops);
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        cmx[q] = load_mask(o1.x, q); cmy[q] = load_mask(o1.y, q);
                        nav[q] = nbv[q] = V4{0, 0, 0, 0};
                        if (HOIST) {
                            if (o1.x >= 0) nav[q] = ldg256_bwd(SCR(o1.x, q));
                            if (HOIST_B && o1.y >= 0 && !(RECOMP && (o1.w & FL_B_CHERRY))) nbv[q] = ldg256_bwd(SCR(o1.y, q));
                        }
                    }
                }
#pragma unroll 1
                for (int i = nops; i >= 1; --i) {
                    const int4 o = OPS(i);
                    unsigned nmx[NP], nmy[NP];  // ... and of the join that comes next in this (reversed) order
                    if (!SMALL) {
                        const int4 on = OPS(i - 1);
#pragma unroll
                        for (int q = 0; q < NP; ++q) { nmx[q] = load_mask(on.x, q); nmy[q] = load_mask(on.y, q); }
                    }
#if DODO_PF_L2_DIST > 0
                    {
                        // pull what the join a few steps ahead needs out of HBM into L2
                        const int4 of = OPS(i > DODO_PF_L2_DIST ? i - DODO_PF_L2_DIST : 0);
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (of.x >= 0) prefetch_l2(SCR(of.x, q));
                            if (of.y >= 0 && !(RECOMP && (of.w & FL_B_CHERRY))) prefetch_l2(SCR(of.y, q));
                        }
                    }
#endif
#if DODO_PF_L1
                    {   // ... and what the next join needs from L2 into L1
                        const int4 o1 = OPS(i - 1);
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (o1.x >= 0) prefetch_l1(SCR(o1.x, q));
                            if (o1.y >= 0 && !(RECOMP && (o1.w & FL_B_CHERRY))) prefetch_l1(SCR(o1.y, q));
                        }
                    }
#endif
                    if (!HOIST_W && (o.w & FL_A_CHILD)) {  // a-child: its upper message was parked in its own slot by the parent
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            Wv[q] = ldg256(SCR(o.z, q));
                            if (SCALE) we[q] = __ldcg(CSCR(o.z, q));
                        }
                    }
                    double2 ab_a = make_double2(0.0, 0.0), ab_b = ab_a;
                    if (pmode == PM_JC) {  // early, as in the forward sweep
                        if (o.x >= 0) ab_a = load_ab<SMALL>(tab_at((unsigned)o.x * 16u));
                        if (o.y >= 0) ab_b = load_ab<SMALL>(tab_at((unsigned)o.y * 16u));
                    }
                    V4 av[NP], bv[NP], ua[NP], ub[NP];
                    V4 za = V4{0, 0, 0, 0}, zb = za;  // GMODEL: e o (U^T u) of the two child edges
                    int ac[NP], bcn[NP];
                    double ga = 0.0, gb = 0.0;
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        ac[q] = 0; bcn[q] = 0;
                        if (o.x < 0) av[q] = tipmsg(~o.x, q, cmx[q]);
                        else { av[q] = HOIST ? nav[q] : ldg256_bwd(SCR(o.x, q)); if (SCALE) ac[q] = __ldcg(CSCR(o.x, q)); }
                        if (o.y < 0) bv[q] = tipmsg(~o.y, q, cmy[q]);
                        else if (RECOMP && (o.w & FL_B_CHERRY)) bv[q] = cherry_msg(OPS(i - 1), q, bcn[q], SMALL ? 0u : nmx[q], SMALL ? 0u : nmy[q]);
                        else { bv[q] = HOIST_B ? nbv[q] : ldg256_bwd(SCR(o.y, q)); if (SCALE) bcn[q] = __ldcg(CSCR(o.y, q)); }
                        ua[q] = vmul(Wv[q], bv[q]);
                        ub[q] = vmul(Wv[q], av[q]);
                        // explicitly rounded: the result must not depend on whether the (exact)
                        // rescaling factor sits between the product and the accumulation
                        double gaq, gbq;
                        if (GMODEL) {
                            const double sf = SCALE ? pow2_256(we[q] - ac[q] - bcn[q]) : 1.0;
                            gaq = model_edge<SMALL>(md, tab_at(o_X + (unsigned)(o.x < 0 ? ~o.x : o.x + S) * 16u), ua[q], av[q], rl, sf, Wm, za);
                            gbq = model_edge<SMALL>(md, tab_at(o_X + (unsigned)(o.y < 0 ? ~o.y : o.y + S) * 16u), ub[q], bv[q], rl, sf, Wm, zb);
                            if (i == nops) {  // root: d lnL / d pi_j = sum_k sc_k (msg_a * msg_b)_j
                                const double s2 = sc0 * sf;
                                const double g4 = warp_sum4(s2 * (av[q].a * bv[q].a), s2 * (av[q].b * bv[q].b),
                                                            s2 * (av[q].c * bv[q].c), s2 * (av[q].d * bv[q].d), lane);
                                if ((lane & 7u) == 0u) smd[o_wred + warp * 20u + 16u + (lane >> 3)] += g4;
                            }
                        } else {
                            gaq = __dmul_rn(rate_k, uQx(md.Q, ua[q], av[q]));
                            gbq = __dmul_rn(rate_k, uQx(md.Q, ub[q], bv[q]));
                            if (SCALE) {
                                const double sf = pow2_256(we[q] - ac[q] - bcn[q]);
                                gaq = __dmul_rn(gaq, sf); gbq = __dmul_rn(gbq, sf);
                            }
                        }
                        ga = __dadd_rn(ga, gaq); gb = __dadd_rn(gb, gbq);
                    }
                    if (HOIST) {  // av / bv are dead from here on: the next join's operands take their place
                        const int4 on = OPS(i - 1);  // (entry 0 is a sentinel without internal children or flags)
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (on.x >= 0) nav[q] = ldg256_bwd(SCR(on.x, q));
                            if (HOIST_B && on.y >= 0 && !(RECOMP && (on.w & FL_B_CHERRY))) nbv[q] = ldg256_bwd(SCR(on.y, q));
                            // an a-child follows a join without internal children: Wv is not produced below
                            if (HOIST_W && (on.w & FL_A_CHILD)) Wv[q] = ldg256(SCR(on.z, q));
                        }
                    }
                    if (!PAIR) {
                        const double g2 = warp_sum2(ga, gb, lane);
                        if ((lane & 15u) == 0u) {
                            const int c = lane ? o.y : o.x;
                            GACC_ADD(gw + (unsigned)(c < 0 ? ~c : c + S), g2);
                        }
                    } else if (have_pending) {  // warp-uniform: every second join reduces its own and the parked pair
                        const double g4 = warp_sum4(pend_ga, pend_gb, ga, gb, lane);
                        if ((lane & 7u) == 0u) {
                            const int c = lane == 0u ? pend_x : (lane == 8u ? pend_y : (lane == 16u ? o.x : o.y));
                            GACC_ADD(gw + (unsigned)(c < 0 ? ~c : c + S), g4);
                        }
                        have_pending = false;
                    } else {
                        pend_ga = ga; pend_gb = gb; pend_x = o.x; pend_y = o.y;
                        have_pending = true;
                    }
                    if (o.x >= 0) {
                        V4 wa[NP];
                        if (GMODEL && DODO_GM_EIG_DOWN) wa[0] = matvec_cT(md.Uinv, za);
                        else if (pmode == PM_JC) matvec_jc_ab<NP>(ab_a, ua, wa);
                        else apply_P<SMALL, NP, true>(pmode, md, tab_at((unsigned)o.x * 16u), ua, wa);
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            int ea = we[q] - bcn[q];
                            if (SCALE) { int neg = -ea; scale_both(wa[q], neg); ea = -neg; }
                            stg256(SCR(o.x, q), wa[q]);
                            if (SCALE) __stcg(CSCR(o.x, q), ea);
                        }
                    }
                    if (o.y >= 0) {  // the next join (in this reversed order) is exactly node b
                        if (GMODEL && DODO_GM_EIG_DOWN) Wv[0] = matvec_cT(md.Uinv, zb);
                        else if (pmode == PM_JC) matvec_jc_ab<NP>(ab_b, ub, Wv);
                        else apply_P<SMALL, NP, true>(pmode, md, tab_at((unsigned)o.y * 16u), ub, Wv);
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            we[q] -= ac[q];
                            if (SCALE) { int neg = -we[q]; scale_both(Wv[q], neg); we[q] = -neg; }
                        }
                    }
                    if (!SMALL) {
#pragma unroll
                        for (int q = 0; q < NP; ++q) { cmx[q] = nmx[q]; cmy[q] = nmy[q]; }
                    }
                }
                if (PAIR && have_pending) {  // odd number of joins: the last pair goes alone
                    const double g2 = warp_sum2(pend_ga, pend_gb, lane);
                    if ((lane & 15u) == 0u) {
                        const int c = lane ? pend_y : pend_x;
                        GACC_ADD(gw + (unsigned)(c < 0 ? ~c : c + S), g2);
                    }
                }
            }
        }
        if (GMATS) {
            // Node-centric reverse sweep: join i receives u_n (upper message on the PARENT side of the
            // edge above n), forms w_n = P_n^T u_n and L_n = msg_a * msg_b, emits
            // dlnL/dP_n = sum_patterns u_n (x) L_n, and hands u_a = w_n * msg_b, u_b = w_n * msg_a down.
            // A tip child's edge is emitted here too: dlnL/dP_t = sum u_t (x) indicator(mask_t).
            double gpi0 = 0.0, gpi1 = 0.0, gpi2 = 0.0, gpi3 = 0.0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                stage_tables(k);
                V4 *scr = scr0 + (size_t)k * (S - 1) * npt;
                const double sc = A.mix ? wt[0] / (f[0] * (double)K) : wt[0] / FX(k, 0);
                double *gpart = A.gm_part + ((((size_t)item * K + k) * nwarp + warp) * (size_t)E) * 16;
                V4 uprev = V4{0, 0, 0, 0};
                auto maskvec = [&](int t) -> V4 {
                    const unsigned m = SMALL ? (unsigned)SEL(t, 0) : (unsigned)__ldg(tm[0] + (size_t)t * A.ld_mask);
                    return V4{(double)(m & 1u), (double)((m >> 1) & 1u), (double)((m >> 2) & 1u), (double)((m >> 3) & 1u)};
                };
#pragma unroll 1
                for (int i = nops; i >= 1; --i) {
                    const int4 o = OPS(i);
                    const V4 av = o.x < 0 ? tipmsg(~o.x, 0, load_mask(o.x, 0)) : ldg256(SCR(o.x, 0));
                    int unused_cnt;
                    const int4 oc = OPS(i - 1);
                    const V4 bv = o.y < 0 ? tipmsg(~o.y, 0, load_mask(o.y, 0))
                                          : ((RECOMP && (o.w & FL_B_CHERRY)) ? cherry_msg(oc, 0, unused_cnt, load_mask(oc.x, 0), load_mask(oc.y, 0))
                                                                             : ldg256(SCR(o.y, 0)));
                    const V4 Lv = vmul(av, bv);
                    V4 w;
                    if (i == nops) {
                        w = V4{md.pi[0] * sc, md.pi[1] * sc, md.pi[2] * sc, md.pi[3] * sc};
                        gpi0 = fma(sc, Lv.a, gpi0); gpi1 = fma(sc, Lv.b, gpi1);
                        gpi2 = fma(sc, Lv.c, gpi2); gpi3 = fma(sc, Lv.d, gpi3);
                    } else {
                        const V4 u = (o.w & FL_A_CHILD) ? ldg256(SCR(o.z, 0)) : uprev;
                        outer_reduce_store(u, Lv, gpart + (size_t)(o.z + S) * 16, lane);
                        w = matvecT_t<SMALL>(tab_at((unsigned)o.z * 16u), u);
                    }
                    const V4 ua = vmul(w, bv), ub = vmul(w, av);
                    if (o.x < 0) outer_reduce_store(ua, maskvec(~o.x), gpart + (size_t)(~o.x) * 16, lane);
                    else stg256(SCR(o.x, 0), ua);
                    if (o.y < 0) outer_reduce_store(ub, maskvec(~o.y), gpart + (size_t)(~o.y) * 16, lane);
                    else uprev = ub;
                }
            }
            gpi0 = warp_sum(gpi0); gpi1 = warp_sum(gpi1); gpi2 = warp_sum(gpi2); gpi3 = warp_sum(gpi3);
            if (lane == 0) {
                double *gq = A.gpi_part + ((size_t)item * nwarp + warp) * 4;
                gq[0] = gpi0; gq[1] = gpi1; gq[2] = gpi2; gq[3] = gpi3;
            }
        }
        if (GMODEL) {  // [m][n] with n < 3 -> the 4 x 4 slot layout of the per-warp buffer (column 3 stays zero)
            const double x16[16] = {Wm[0], Wm[1], Wm[2], 0.0, Wm[3], Wm[4], Wm[5], 0.0, Wm[6], Wm[7], Wm[8], 0.0, Wm[9], Wm[10], Wm[11], 0.0};
            reduce16_store(x16, &smd[o_wred + warp * 20u], lane);
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (unsigned w2 = 0; w2 < nwarp; ++w2) s += RED(w2);
            A.ll_part[item] = s;
        }
        if (GMODEL && tid >= 32u && tid < 52u) {  // W (16) and d lnL / d pi (4) of this item, warps in index order
            double s = 0.0;
            for (unsigned w2 = 0; w2 < nwarp; ++w2) s += smd[o_wred + w2 * 20u + (tid - 32u)];
            A.gm_part[(size_t)item * 20 + (tid - 32u)] = s;
        }
        if (GRAD) {
            double *gp = A.g_part + (size_t)item * E;
            for (unsigned e = tid; e < (unsigned)E; e += nthr) {
                double s = 0.0;
                for (unsigned w2 = 0; w2 < nwarp; ++w2) {
                    s += GACC(w2 * (unsigned)E + e);
                    GACC(w2 * (unsigned)E + e) = 0.0;
                }
                gp[e] = s;
            }
        }
        if (A.next_item == nullptr) { item += (int)gridDim.x; continue; }  // fixed stride (DODO_PRUNE_STATIC=1: A/B runs)
        if (tid == 0) s_next_item = (int)gridDim.x + atomicAdd(A.next_item, 1);
        __syncthreads();
        item = s_next_item;
    }
#undef GACC
#undef GACC_ADD
#undef FX
#undef RED
#undef CX
#undef OPS
#undef OPSW
#undef SEL
#undef SELW
#undef SCR
#undef CSCR
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

// fixed-order reduction of the per-(tile, warp) outer-product partials
__global__ void prune_reduce_gmats_kernel(const double *__restrict__ gm_part, const double *__restrict__ gpi_part,
                                          int B, int n_tiles, int K, int nwarp, int E, double *__restrict__ grad_P,
                                          double *__restrict__ grad_pi) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nP = (size_t)B * E * K * 16;
    if (gid < nP) {
        const int ij = gid % 16;
        const int k = (gid / 16) % K;
        const int e = (gid / (16 * (size_t)K)) % E;
        const int b = gid / (16 * (size_t)K * E);
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t)
            for (int w = 0; w < nwarp; ++w)
                s += gm_part[(((((size_t)b * n_tiles + t) * K + k) * nwarp + w) * E + e) * 16 + ij];
        grad_P[gid] = s;
    } else if (gid < nP + (size_t)B * 4) {
        const int i = (gid - nP) % 4, b = (gid - nP) / 4;
        double s = 0.0;
        for (int t = 0; t < n_tiles; ++t)
            for (int w = 0; w < nwarp; ++w) s += gpi_part[(((size_t)b * n_tiles + t) * nwarp + w) * 4 + i];
        grad_pi[(size_t)b * 4 + i] = s;
    }
}

// model-gradient launches: W [B][16] and d lnL / d pi [B][4] from the per-item partials, tiles in index order
// (W comes back in the caller's eigenvector order: `zero` = the position of the eigenvalue 0 there, swapped with 3 here)
__global__ void prune_reduce_model_kernel(const double *__restrict__ part, int B, int n_tiles, int zero,
                                          double *__restrict__ W, double *__restrict__ grad_pi) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * 20) return;
    const int b = gid / 20, i = gid - b * 20;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += part[((size_t)b * n_tiles + t) * 20 + i];
    if (i < 16) {
        const int m = i >> 2, n = i & 3;
        const int pm = m == 3 ? zero : (m == zero ? 3 : m), pn = n == 3 ? zero : (n == zero ? 3 : n);
        W[(size_t)b * 16 + pm * 4 + pn] = s;
    } else grad_pi[(size_t)b * 4 + (i - 16)] = s;
}

// ------------------------------------------------------------------------------------------
// Chain rule of a device-formed reversible model (learnable GTR: the reference's `--model GTR` always learns
// rates and frequencies, phylomodel.py:69-74, through torch autograd over GTR_p_t, :96-114).  From
// G[b][e][k] = d lnL_b / d P(r_k t_e) (the pruning kernel in DODO_PRUNE_GMATS mode with device-formed P) to
//   d lnL_b / d t_e = sum_k r_k sum_m lam_m e^{lam_m r_k t_e} M_mm,           M = U^T G U^-T
//   W_b            = sum_{e,k} Phi(r_k t_e) o M                               (4 x 4)
// with Phi_mn(tau) = tau (e^{lam_m tau} - e^{lam_n tau}) / ((lam_m - lam_n) tau), Phi_mm = tau e^{lam_m tau}
// (Daleckii-Krein: d exp(Q tau) = U [Phi o (U^-1 dQ U)] U^-1), so that d lnL_b / d Q = U^-T W_b U^T; the host
// carries that 4 x 4 matrix through Q(rates, freqs).  One CTA per sample, one thread per edge, W reduced
// over the edges in a fixed order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) model_chain_kernel(ModelDev md, const double *__restrict__ blens,
                                                        const double *__restrict__ gP, int S, int K,
                                                        double *__restrict__ gbl, double *__restrict__ Wout) {
    __shared__ double red[128][17];
    const int b = blockIdx.x, tid = threadIdx.x, E = 2 * S - 2;
    double W[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) W[i] = 0.0;
    for (int e = tid; e < E; e += 128) {
        const double t = blens[(size_t)b * E + e];
        double gb = 0.0;
        for (int k = 0; k < K; ++k) {
            const double rk = md.rates[k], tau = t * rk;
            const double *G = gP + (((size_t)b * E + e) * K + k) * 16;
            double tmp[16], M[16], ex[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int n = 0; n < 4; ++n) {  // (G U^-T)_in = sum_j G_ij Uinv_nj
                    double a = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) a = fma(G[i * 4 + j], md.Uinv[n * 4 + j], a);
                    tmp[i * 4 + n] = a;
                }
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) {  // M_mn = sum_i U_im (G U^-T)_in
                    double a = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) a = fma(md.U[i * 4 + m], tmp[i * 4 + n], a);
                    M[m * 4 + n] = a;
                }
#pragma unroll
            for (int m = 0; m < 4; ++m) ex[m] = exp(md.lam[m] * tau);
#pragma unroll
            for (int m = 0; m < 4; ++m) gb = fma(rk * md.lam[m] * ex[m], M[m * 4 + m], gb);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    double phi;
                    if (m == n) phi = tau * ex[m];
                    else {
                        const double d = fmin((md.lam[m] - md.lam[n]) * tau, 700.0);
                        phi = tau * ex[n] * (fabs(d) < 1e-8 ? 1.0 + 0.5 * d : expm1(d) / d);
                    }
                    W[m * 4 + n] = fma(phi, M[m * 4 + n], W[m * 4 + n]);
                }
        }
        gbl[(size_t)b * E + e] = gb;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) red[tid][i] = W[i];
    __syncthreads();
    for (int half = 64; half >= 1; half >>= 1) {  // fixed-order tree
        if (tid < half)
#pragma unroll
            for (int i = 0; i < 16; ++i) red[tid][i] += red[tid + half][i];
        __syncthreads();
    }
    if (tid < 16) Wout[(size_t)b * 16 + tid] = red[0][tid];
}

typedef void (*prune_fn_t)(PruneArgs, ModelDev);
template <int MODE, int NP, int PM>
static prune_fn_t prune_pick(bool small, bool scale) {
    if (small) return scale ? prune_kernel<MODE, true, true, NP, PM> : prune_kernel<MODE, true, false, NP, PM>;
    return scale ? prune_kernel<MODE, false, true, NP, PM> : prune_kernel<MODE, false, false, NP, PM>;
}
template <int MODE, int NP>
static prune_fn_t prune_pick_pm(bool small, bool scale, int pm) {
    if (pm == PM_JC) return prune_pick<MODE, NP, PM_JC>(small, scale);
#if DODO_PRUNE_EIG
    if (pm == PM_EIG) return prune_pick<MODE, NP, PM_EIG>(small, scale);
#endif
    return prune_pick<MODE, NP, PM_GENERIC>(small, scale);
}
static prune_fn_t prune_entry(int mode, bool small, bool scale, int np, int pm) {
    if (mode == 3) return prune_pick<3, 1, PM_GENERIC>(small, scale);  // eigen models only
    if (mode == 2) {  // d lnL/d P reads P itself: generic or JC69 forward, generic reverse sweep
        if (pm == PM_JC) return small ? prune_kernel<2, true, false, 1, PM_JC> : prune_kernel<2, false, false, 1, PM_JC>;
        return small ? prune_kernel<2, true, false, 1, PM_GENERIC> : prune_kernel<2, false, false, 1, PM_GENERIC>;
    }
#ifdef DODO_PRUNE_NP
    if (np == 2) return mode == 1 ? prune_pick_pm<1, 2>(small, scale, pm) : prune_pick_pm<0, 2>(small, scale, pm);
#endif
    (void)np;
    return mode == 1 ? prune_pick_pm<1, 1>(small, scale, pm) : prune_pick_pm<0, 1>(small, scale, pm);
}
// tables + selectors live in shared memory when that still leaves room for several CTAs per SM
#ifndef DODO_PRUNE_SMEM_KB
#define DODO_PRUNE_SMEM_KB 74
#endif
static size_t prune_small_budget(int np, int gm) { return (size_t)((np == 2 || gm) ? 110 : DODO_PRUNE_SMEM_KB) * 1024; }  // 2 or 3 CTAs per SM
static bool prune_small(int S, int K, int np, int gm) { return prune_smem_bytes(S, K, true, np, 0, gm) <= prune_small_budget(np, gm); }
// Tabulated cherries per tree: whatever shared memory is left inside the budget that keeps the CTA
// count per SM (100 doubles per cherry); none with rescaling (counters are not tabulated) or when
// the tables are not resident.
static int prune_cherry_slots(int S, int K, int np, int gm, bool small, bool scale) {
#ifdef DODO_PRUNE_NO_CHERRY_TABLE
    return 0;
#endif
    if (!small || scale) return 0;
    const size_t left = prune_small_budget(np, gm) - prune_smem_bytes(S, K, true, np, 0, gm);
    const int fit = (int)(left / 800);
    return fit < S / 2 ? fit : S / 2;
}
// patterns per thread.  Two per thread halve the warp-uniform work (schedule decode, P fills) per
// pattern but cost registers (116 vs 78 -> 2 instead of 3 CTAs per SM): measured 5.1 ms against
// 4.75 ms at 64 taxa x 10k patterns x 128 samples x 4 categories, so one per thread is the default.
static int prune_np(int B, int L, uint32_t flags) {
    (void)B; (void)L;
    if (flags & (DODO_PRUNE_GMATS | DODO_PRUNE_GMODEL)) return 1;
#ifdef DODO_PRUNE_NP
    return DODO_PRUNE_NP;
#else
    return 1;
#endif
}
static int prune_mode(uint32_t flags) {
    return (flags & DODO_PRUNE_GMODEL) ? 3 : ((flags & DODO_PRUNE_GMATS) ? 2 : ((flags & DODO_PRUNE_GRAD) ? 1 : 0));
}

}  // namespace dodo

using namespace dodo;

extern "C" int dodo_prune_plan(int B, int S, int L, int K, uint32_t flags, int grid_hint, int threads_hint,
                               dodo_prune_plan_t *plan) {
    DODO_CHECK_ARG(plan != nullptr, "dodo_prune_plan: plan is NULL");
    DODO_CHECK_ARG(B >= 1 && S >= 2 && L >= 1, "dodo_prune_plan: need B>=1, S>=2, L>=1 (got %d,%d,%d)", B, S, L);
    DODO_CHECK_ARG(K >= 1 && K <= MAX_K, "dodo_prune_plan: K must be in [1,%d] (got %d)", MAX_K, K);
    DODO_CHECK_ARG(!((flags & DODO_PRUNE_GMATS) && (flags & DODO_PRUNE_GRAD)),
                   "dodo_prune_plan: GRAD and GMATS are separate calls");
    DODO_CHECK_ARG(!((flags & DODO_PRUNE_GMATS) && (flags & DODO_PRUNE_SCALE)),
                   "dodo_prune_plan: GMATS is not available with rescaling");
    DODO_CHECK_ARG(!((flags & DODO_PRUNE_GMODEL) && (flags & DODO_PRUNE_GMATS)),
                   "dodo_prune_plan: GMODEL and GMATS are separate calls");
    if (flags & DODO_PRUNE_GMODEL) flags |= DODO_PRUNE_GRAD;  // the model gradient comes with the branch-length gradient
    memset(plan, 0, sizeof(*plan));
    plan->B = B; plan->S = S; plan->L = L; plan->K = K; plan->flags = flags;
    (void)threads_hint;
    const int threads = NT;
    const int np = prune_np(B, L, flags);
    const int gm = (flags & DODO_PRUNE_GMODEL) ? 1 : 0;
    const bool small = prune_small(S, K, np, gm);
    const bool scale = (flags & DODO_PRUNE_SCALE) != 0;
    const int chmax = prune_cherry_slots(S, K, np, gm, small, scale);
    plan->cherry_slots = chmax;
    const size_t smem = prune_smem_bytes(S, K, small, np, chmax, gm, (flags & DODO_PRUNE_GRAD) ? 0 : 1);
    if (smem > 227 * 1024) {
        set_error("dodo_prune_plan: %d taxa need %zu bytes of shared memory per CTA (limit 227 KB)", S, smem);
        return DODO_ENOTIMPL;
    }
    plan->threads = threads;
    plan->tile_patterns = threads * np;
    plan->n_tiles = (L + plan->tile_patterns - 1) / plan->tile_patterns;
    long long items = (long long)B * plan->n_tiles;
    DODO_CHECK_ARG(items < (1ll << 31), "dodo_prune_plan: too many work items");
    plan->n_items = (int)items;
    int grid = grid_hint;
    if (grid <= 0) {
        // persistent grid: every CTA the SMs can hold at once (a multiple of the SM count)
        prune_fn_t fn = prune_entry(prune_mode(flags), small, scale, np, PM_GENERIC);  // same footprint for every PM
        if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
        if (e != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
        grid = per_sm * sm_count();
    }
    if (grid > plan->n_items) grid = plan->n_items;
    plan->grid = grid;
    const uint64_t E = 2ull * S - 2, nI = S - 1;
    uint64_t off = 0;
    plan->off_ops = off;     off = align_up(off + (uint64_t)B * nI * sizeof(int4), 256);
    plan->off_sched = off;   off = align_up(off + (uint64_t)B * 6 * S * sizeof(int), 256);
    plan->off_pmat = off;    off = align_up(off + (uint64_t)B * K * tab_doubles(S, chmax, gm) * sizeof(double), 256);
    plan->off_scratch = off; off = align_up(off + (uint64_t)grid * K * nI * threads * np * 32, 256);
    plan->off_tiptab = off;  // rescaling counters (int32 per stored message)
    if (scale) off = align_up(off + (uint64_t)grid * K * nI * threads * np * sizeof(int), 256);
    plan->off_ll_part = off; off = align_up(off + (uint64_t)plan->n_items * sizeof(double), 256);
    plan->off_g_part = off;
    if (flags & DODO_PRUNE_GRAD) off = align_up(off + (uint64_t)plan->n_items * E * sizeof(double), 256);
    plan->off_gm_part = off;
    if (flags & DODO_PRUNE_GMATS) {
        const uint64_t nwarp = threads / 32;
        off = align_up(off + (uint64_t)plan->n_items * K * nwarp * E * 16 * sizeof(double), 256);
        off = align_up(off + (uint64_t)plan->n_items * nwarp * 4 * sizeof(double), 256);
    }
    if (gm) off = align_up(off + (uint64_t)plan->n_items * 20 * sizeof(double), 256);  // per-item W (16) + d/d pi (4)
    plan->ws_bytes = off;
    return DODO_OK;
}

static int prune_launch(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
                        const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights,
                        const int64_t *d_peel, int peel_batch, const double *d_blens, const double *d_P_in,
                        void *d_ws, double *d_loglik, double *d_grad_blens, double *d_grad_P,
                        double *d_grad_pi, double *d_W, void *stream_) {
    DODO_CHECK_ARG(plan && model && d_tipmask && d_weights && d_peel && d_ws && d_loglik, "dodo_prune: NULL argument");
    DODO_CHECK_ARG(d_blens || d_P_in, "dodo_prune: need blens or explicit transition matrices");
    const int B = plan->B, S = plan->S, L = plan->L, K = plan->K;
    DODO_CHECK_ARG(peel_batch == 1 || peel_batch == B, "dodo_prune: peel_batch must be 1 or B");
    DODO_CHECK_ARG(ld_mask >= L, "dodo_prune: ld_mask < L");
    const bool grad = plan->flags & DODO_PRUNE_GRAD;
    DODO_CHECK_ARG(!grad || d_grad_blens, "dodo_prune: GRAD requested but d_grad_blens is NULL");
    DODO_CHECK_ARG(!(grad && d_P_in), "dodo_prune: blens gradient needs device-formed P(t) (use GMATS for explicit mats)");
    const bool gmats = plan->flags & DODO_PRUNE_GMATS;
    DODO_CHECK_ARG(!gmats || (d_grad_P && d_grad_pi), "dodo_prune: GMATS requested but d_grad_P / d_grad_pi is NULL");
    const int gm = (plan->flags & DODO_PRUNE_GMODEL) ? 1 : 0;
    DODO_CHECK_ARG(!gm || (d_W && d_grad_pi), "dodo_prune_model: d_W / d_grad_pi is NULL");
    DODO_CHECK_ARG(!gm || model->kind == DODO_MODEL_EIGEN, "dodo_prune_model: needs an eigen-decomposed (GTR) model");
    DODO_CHECK_ARG(gm || !d_W, "dodo_prune: the plan was not made with DODO_PRUNE_GMODEL");
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
    int zero = 3;
    if (gm) {  // the eigenvalue 0 goes last (see model_edge); W is handed back in the caller's order
        double big = 0.0;
        for (int m = 0; m < 4; ++m) big = fmax(big, fabs(md.lam[m]));
        for (int m = 0; m < 4; ++m)
            if (fabs(md.lam[m]) < fabs(md.lam[zero])) zero = m;
        DODO_CHECK_ARG(fabs(md.lam[zero]) <= 1e-9 * big, "dodo_prune_model: the model has no zero eigenvalue (not a rate matrix)");
        if (zero != 3) {
            for (int i = 0; i < 4; ++i) {
                const double u = md.U[i * 4 + zero]; md.U[i * 4 + zero] = md.U[i * 4 + 3]; md.U[i * 4 + 3] = u;
                const double v = md.Uinv[zero * 4 + i]; md.Uinv[zero * 4 + i] = md.Uinv[3 * 4 + i]; md.Uinv[3 * 4 + i] = v;
            }
            const double l = md.lam[zero]; md.lam[zero] = md.lam[3]; md.lam[3] = l;
        }
    }
    for (int k = 0; k < MAX_K; ++k)
        for (int m = 0; m < 4; ++m) md.rlam[k][m] = md.rates[k] * md.lam[m];

    const int pmode = d_P_in != nullptr ? PM_GENERIC
                      : (model->kind == DODO_MODEL_JC69 ? PM_JC
                         : ((DODO_PRUNE_EIG && !(plan->flags & DODO_PRUNE_GMATS)) ? PM_EIG : PM_GENERIC));  // d lnL/d P reads P itself
    const int chmax = plan->cherry_slots;
    int4 *ops = (int4 *)(ws + plan->off_ops);
    int *chop = (int *)(ws + plan->off_sched);  // [peel_batch][S/2] schedule position of cherry c
    double *tab = (double *)(ws + plan->off_pmat);

    {
        int nb = peel_batch;
        size_t sm = ((size_t)10 * S + 4) * sizeof(int) + (size_t)(S - 1) * sizeof(int4);
        if (sm > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(schedule_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        schedule_kernel<<<nb, 32, sm, stream>>>(d_peel, nb, S, ops, chop);
        DODO_LAUNCH_CHECK();
    }
    {
        size_t total = (size_t)B * E * K;
        transition_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(md, d_blens, d_P_in, B, S, K,
                                                                               pmode == PM_EIG ? 1 : 0, chmax, gm, tab);
        DODO_LAUNCH_CHECK();
        if (chmax > 0) {
            size_t tc = (size_t)B * K * chmax * 25;
            cherry_kernel<<<(unsigned)((tc + 127) / 128), 128, 0, stream>>>(ops, chop, peel_batch, B, S, K, chmax, gm, pmode, md, tab);
            DODO_LAUNCH_CHECK();
        }
    }
    PruneArgs A;
    A.tipmask = d_tipmask; A.ld_mask = ld_mask; A.weights = d_weights;
    A.ops = ops; A.ops_batch = peel_batch; A.tab = tab;
    A.scratch = (double2 *)(ws + plan->off_scratch);
    A.cscratch = (int *)(ws + plan->off_tiptab);
    A.ll_part = (double *)(ws + plan->off_ll_part);
    A.g_part = (double *)(ws + plan->off_g_part);
    A.gm_part = (double *)(ws + plan->off_gm_part);
    A.gpi_part = A.gm_part + (align_up((uint64_t)plan->n_items * K * (plan->threads / 32) * E * 16 * sizeof(double), 256) / sizeof(double));
    A.B = B; A.S = S; A.L = L; A.K = K;
    A.n_tiles = plan->n_tiles; A.n_items = plan->n_items; A.tile_patterns = plan->tile_patterns;
    A.next_item = chop + (size_t)B * S;  // spare words of the schedule region (chop uses peel_batch * S/2 of its B * 6S)
    if (getenv("DODO_PRUNE_STATIC")) A.next_item = nullptr;
    else DODO_CUDA(cudaMemsetAsync(A.next_item, 0, sizeof(int), stream));
    A.mix = ((plan->flags & DODO_PRUNE_MIX) && K > 1) ? 1 : 0;
    A.pmode = pmode;
    A.ch_max = chmax;
    const int np = plan->tile_patterns / plan->threads;
    const bool small = prune_small(S, K, np, gm);
    size_t smem = prune_smem_bytes(S, K, small, np, chmax, gm, (plan->flags & DODO_PRUNE_GRAD) ? 0 : 1);
    cudaEvent_t ev0, ev1;
    timing_events(&ev0, &ev1);
    if (ev0 && ev1) DODO_CUDA(cudaEventRecord(ev0, stream));
    {
        prune_fn_t fn = prune_entry(prune_mode(plan->flags), small, (plan->flags & DODO_PRUNE_SCALE) != 0, np, pmode);
        if (smem > 48 * 1024) DODO_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fn<<<plan->grid, plan->threads, smem, stream>>>(A, md);
    }
    DODO_LAUNCH_CHECK();
    if (ev0 && ev1) DODO_CUDA(cudaEventRecord(ev1, stream));
    {
        int total = B * (E + 1);
        prune_reduce_kernel<<<(total + 127) / 128, 128, 0, stream>>>(A.ll_part, A.g_part, B, plan->n_tiles, E, d_loglik,
                                                                      grad ? d_grad_blens : nullptr);
        DODO_LAUNCH_CHECK();
    }
    if (gmats) {
        size_t total = (size_t)B * E * K * 16 + (size_t)B * 4;
        prune_reduce_gmats_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
            A.gm_part, A.gpi_part, B, plan->n_tiles, K, plan->threads / 32, E, d_grad_P, d_grad_pi);
        DODO_LAUNCH_CHECK();
    }
    if (gm) {
        prune_reduce_model_kernel<<<(B * 20 + 127) / 128, 128, 0, stream>>>(A.gm_part, B, plan->n_tiles, zero, d_W, d_grad_pi);
        DODO_LAUNCH_CHECK();
    }
    return DODO_OK;
}

extern "C" int dodo_prune(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
                          const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights,
                          const int64_t *d_peel, int peel_batch, const double *d_blens, const double *d_P_in,
                          void *d_ws, double *d_loglik, double *d_grad_blens, double *d_grad_P,
                          double *d_grad_pi, void *stream_) {
    DODO_CHECK_ARG(plan && !(plan->flags & DODO_PRUNE_GMODEL), "dodo_prune: a DODO_PRUNE_GMODEL plan goes through dodo_prune_model");
    return prune_launch(plan, model, rates, d_tipmask, ld_mask, d_weights, d_peel, peel_batch, d_blens, d_P_in, d_ws,
                        d_loglik, d_grad_blens, d_grad_P, d_grad_pi, nullptr, stream_);
}

extern "C" int dodo_prune_model(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
                                const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights,
                                const int64_t *d_peel, int peel_batch, const double *d_blens, void *d_ws,
                                double *d_loglik, double *d_grad_blens, double *d_W, double *d_grad_pi, void *stream_) {
    DODO_CHECK_ARG(plan && (plan->flags & DODO_PRUNE_GMODEL), "dodo_prune_model: the plan was not made with DODO_PRUNE_GMODEL");
    DODO_CHECK_ARG(d_blens && d_grad_blens, "dodo_prune_model: NULL argument");
    return prune_launch(plan, model, rates, d_tipmask, ld_mask, d_weights, d_peel, peel_batch, d_blens, nullptr, d_ws,
                        d_loglik, d_grad_blens, nullptr, d_grad_pi, d_W, stream_);
}

extern "C" int dodo_model_chain(const dodo_model_t *model, const double *rates, int K, const double *d_blens,
                                const double *d_grad_P, int B, int S, double *d_grad_blens, double *d_W, void *stream_) {
    DODO_CHECK_ARG(model && d_blens && d_grad_P && d_grad_blens && d_W, "dodo_model_chain: NULL argument");
    DODO_CHECK_ARG(B >= 1 && S >= 2 && K >= 1 && K <= MAX_K, "dodo_model_chain: need B>=1, S>=2, 1<=K<=%d", MAX_K);
    DODO_CHECK_ARG(model->kind == DODO_MODEL_EIGEN, "dodo_model_chain: needs an eigen-decomposed (GTR) model");
    ModelDev md;
    memcpy(md.U, model->U, sizeof(md.U));
    memcpy(md.Uinv, model->Uinv, sizeof(md.Uinv));
    memcpy(md.lam, model->lam, sizeof(md.lam));
    memcpy(md.Q, model->Q, sizeof(md.Q));
    memcpy(md.pi, model->pi, sizeof(md.pi));
    md.kind = model->kind;
    for (int k = 0; k < MAX_K; ++k) md.rates[k] = (k < K) ? (rates ? rates[k] : 1.0) : 0.0;
    model_chain_kernel<<<B, 128, 0, (cudaStream_t)stream_>>>(md, d_blens, d_grad_P, S, K, d_grad_blens, d_W);
    DODO_LAUNCH_CHECK();
    return DODO_OK;
}

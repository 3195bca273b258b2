/*
 * dodonaphy_b200 -- C ABI of the B200-native Dodonaphy hot path.
 *
 * Plain C: raw device pointers, explicit sizes, a CUDA stream passed as void*.  No torch types.
 * Every entry point returns 0 on success and a negative DODO_E* code otherwise;
 * dodo_last_error() gives a human-readable message for the calling thread.
 *
 * All floating point is IEEE fp64.  Index data (peel) is int64, exactly like the reference
 * (Cpeeler.pyx:28, peeler.py:180).  Batch axis B = VI samples or MCMC chains; S taxa; L site
 * patterns; K rate categories; E = 2S-2 branches; D embedding dimension.
 *
 * Each function names the reference interface it replaces (paths relative to the reference
 * repository mattapow/dodonaphy).
 */
#ifndef DODONAPHY_B200_H
#define DODONAPHY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DODO_OK 0
#define DODO_EINVAL (-1)  /* bad argument (reference raises ValueError)            */
#define DODO_ECUDA (-2)   /* CUDA runtime error                                    */
#define DODO_ENOTIMPL (-3) /* combination not implemented (NotImplementedError)    */
#define DODO_EINEXACT (-4) /* soft-NJ exactness guard tripped (see dodo_nj_soft)   */

const char *dodo_last_error(void);
int dodo_version(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
uint64_t dodo_launch_count(void);
/* Measurement hook: when both handles (cudaEvent_t) are non-NULL, dodo_prune records `start`
 * immediately before and `stop` immediately after its main pruning kernel on the caller's stream,
 * so that a benchmark can time that kernel alone without a profiler.  Pass NULLs to switch off. */
void dodo_set_timing_events(void *start, void *stop);

/* ------------------------------------------------------------------------------------------
 * Substitution model.  P(t) = U diag(exp(lam t)) Uinv, dP/dt = Q P.
 * Replaces PhyloModel.get_transition_mats / JC69_p_t / GTR_p_t (dodonaphy/phylomodel.py:76-114)
 * and Cphylo.JC69_p_t (dodonaphy/cython/Cphylo.pyx:45-51): P(t) is formed on the device.
 * kind = DODO_MODEL_JC69 uses the reference's closed form a = 1/4 + 3/4 e^{-4t/3},
 * b = 1/4 - 1/4 e^{-4t/3} (U, Uinv, lam ignored); kind = DODO_MODEL_EIGEN uses the eigen-system
 * (the caller obtains it as the reference does, eigh of sqrt(pi) Q sqrt(pi)^-1).
 * ------------------------------------------------------------------------------------------ */
#define DODO_MODEL_JC69 0
#define DODO_MODEL_EIGEN 1
typedef struct {
    double U[16];    /* row-major 4x4 */
    double Uinv[16]; /* row-major 4x4 */
    double lam[4];
    double Q[16]; /* row-major rate matrix, rows sum to 0 */
    double pi[4]; /* root frequencies ("freqs" argument of calculate_treelikelihood) */
    int32_t kind;
    int32_t _pad;
} dodo_model_t;

/* ------------------------------------------------------------------------------------------
 * K3: fused P(t) + Felsenstein pruning, forward and (optionally) gradient.
 * Replaces phylo.calculate_treelikelihood (dodonaphy/phylo.py:132-139) + torch autograd,
 * Cphylo.calculate_treelikelihood / compute_LL_np (dodonaphy/cython/Cphylo.pyx:9-42) and the
 * backend slot BaseModel.compute_LL dispatches to (dodonaphy/base_model.py:219-247); shaped like
 * the reference's own native plugin phylo.TreeLikelihood (dodonaphy/phylo.py:177-226): value and
 * gradient in one call.
 * ------------------------------------------------------------------------------------------ */
#define DODO_PRUNE_GRAD 1u     /* also produce d lnL / d blens                                   */
#define DODO_PRUNE_MIX 2u      /* K>1: lnL = sum_p w_p log(mean_k f_k)  (discrete-gamma mixture) */
                               /* without it K categories are independent and their lnL are      */
                               /* summed -- what the reference's broadcasting does (phylo.py:137) */
#define DODO_PRUNE_SCALE 4u    /* power-of-two rescaling (bit-identical where the unscaled       */
                               /* computation stays finite)                                       */
#define DODO_PRUNE_GMATS 8u    /* also produce d lnL / d P[b,e,k,:,:] and d lnL / d pi          */
#define DODO_PRUNE_GMODEL 16u  /* plan for dodo_prune_model: d lnL / d blens, the 4 x 4 matrix W  */
                               /* of the model chain rule and d lnL / d pi (implies GRAD)         */

typedef struct {
    int32_t B, S, L, K;
    uint32_t flags;
    int32_t threads;       /* CTA size                                                    */
    int32_t grid;          /* persistent CTAs                                             */
    int32_t tile_patterns; /* site patterns per CTA work item                             */
    int32_t n_tiles;       /* ceil(L / tile_patterns)                                     */
    int32_t n_items;       /* B * n_tiles                                                 */
    int32_t cherry_slots; /* cherries per tree whose messages are tabulated (shared memory left over) */
    uint64_t ws_bytes;     /* device workspace the caller must provide                    */
    /* offsets (bytes) of the sub-buffers inside the workspace                            */
    uint64_t off_ops, off_sched, off_pmat, off_tiptab, off_scratch, off_ll_part, off_g_part, off_gm_part;
} dodo_prune_plan_t;

/* Fill *plan for a problem size.  grid_hint/threads_hint = 0 picks the defaults. */
int dodo_prune_plan(int B, int S, int L, int K, uint32_t flags, int grid_hint, int threads_hint,
                    dodo_prune_plan_t *plan);

/*
 * d_tipmask : uint8 [S][ld_mask], bit j set <=> state j (A,C,G,T) compatible with the tip
 *             (the 0/1 tip partials of phylo.compress_alignment, dodonaphy/phylo.py:27-46, packed)
 * d_weights : double [L] pattern weights
 * d_peel    : int64 [B or 1][S-1][3] rows [left,right,parent], post-order; peel_batch = 1 shares one
 *             topology across the batch
 * d_blens   : double [B][2S-2]; blens[i] = branch above node i
 * rates     : host double [K] category rates (1.0 for K = 1)
 * d_P_in    : optional double [B][2S-2][K][16] transition matrices supplied by the caller
 *             (the "mats" argument of calculate_treelikelihood); NULL => formed on device
 * outputs   : d_loglik [B]; d_grad_blens [B][2S-2] (flag GRAD, else NULL);
 *             d_grad_P [B][2S-2][K][16] and d_grad_pi [B][4] (flag GMATS, else NULL)
 */
int dodo_prune(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
               const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights,
               const int64_t *d_peel, int peel_batch, const double *d_blens, const double *d_P_in,
               void *d_ws, double *d_loglik, double *d_grad_blens, double *d_grad_P,
               double *d_grad_pi, void *stream);

/* ------------------------------------------------------------------------------------------
 * K1: batched pairwise hyperbolic distance.
 * Replaces Chyp_torch.get_pdm (dodonaphy/cython/Chyp_torch.pyx:399-423; lifts :486-504 "up",
 * :387-396 "wrap") and Chyp_np.get_pdm (dodonaphy/cython/Chyp_np.pyx:321-367).
 * ------------------------------------------------------------------------------------------ */
#define DODO_PDM_TORCH 0 /* clamp H at -1.0000001, zero diagonal       (Chyp_torch.pyx:418,422) */
#define DODO_PDM_NUMPY 1 /* clamp H at -(1+eps), diagonal kept, curvature 0 => Euclidean        */
#define DODO_LIFT_UP 0
#define DODO_LIFT_WRAP 1

uint64_t dodo_pdm_workspace_bytes(int B, int S, int D);
/* d_x [B][S][D]; d_curvature: device double [1] (shared) ; d_pdm [B][S][S] */
int dodo_pdm_fwd(const double *d_x, int B, int S, int D, const double *d_curvature, int variant,
                 int lift, int matsumoto, void *d_ws, double *d_pdm, void *stream);
/* VJP: d_grad_pdm [B][S][S] -> d_grad_x [B][S][D], d_grad_curv [B] (one partial per sample) */
int dodo_pdm_bwd(const double *d_x, int B, int S, int D, const double *d_curvature, int variant,
                 int lift, int matsumoto, void *d_ws, const double *d_grad_pdm, double *d_grad_x,
                 double *d_grad_curv, void *stream);

/* The learnable-GTR step in ONE pruning launch (the reference's `--model GTR` always learns rates and frequencies,
 * dodonaphy/phylomodel.py:69-74: torch autograd through GTR_p_t :96-114 and calculate_treelikelihood
 * dodonaphy/phylo.py:132-139, reached from BaseModel.compute_LL dodonaphy/base_model.py:239-247).  `plan` must be made
 * with DODO_PRUNE_GMODEL, `model` must be an eigen-decomposed model (DODO_MODEL_EIGEN).  Outputs: d_loglik [B],
 * d_grad_blens [B][2S-2], d_grad_pi [B][4] (w.r.t. the root frequencies) and d_W [B][16] (row-major 4 x 4) with
 * d lnL_b / d Q = U^-T W_b U^T for the NORMALISED rate matrix Q of `model` at fixed branch lengths, for every
 * perturbation dQ with zero row sums (any change of a rate matrix): the column of W that belongs to the eigenvalue 0
 * only multiplies the row sums of dQ and is returned as zero.  The caller carries W through Q(rates, freqs) on the
 * host (a 4 x 4 computation).  Every thread accumulates its pattern's share of W over all edges and categories of
 * its walk; nothing per edge leaves the kernel.  Other arguments as dodo_prune. */
int dodo_prune_model(const dodo_prune_plan_t *plan, const dodo_model_t *model, const double *rates,
                     const uint8_t *d_tipmask, int64_t ld_mask, const double *d_weights, const int64_t *d_peel,
                     int peel_batch, const double *d_blens, void *d_ws, double *d_loglik, double *d_grad_blens,
                     double *d_W, double *d_grad_pi, void *stream);

/* The same chain rule from explicit per-edge matrices (kept as the cross-check of dodo_prune_model and for callers
 * that already hold d lnL / d P): a device-formed reversible (GTR) model whose rates and frequencies are being learnt (the
 * reference's `--model GTR`, dodonaphy/phylomodel.py:69-74: torch autograd through GTR_p_t :96-114 and
 * calculate_treelikelihood, reached from BaseModel.compute_LL dodonaphy/base_model.py:239-247).  Input: d_grad_P
 * [B][2S-2][K][16] from dodo_prune(DODO_PRUNE_GMATS) called with branch lengths (P formed on the device from
 * `model`).  Output: d_grad_blens [B][2S-2] = d lnL_b / d t_e and d_W [B][16] (row-major 4 x 4) with
 * d lnL_b / d Q = U^-T W_b U^T for the NORMALISED rate matrix Q of `model` at fixed branch lengths; the
 * caller carries that through Q(rates, freqs) on the host (a 4 x 4 computation). */
int dodo_model_chain(const dodo_model_t *model, const double *rates, int K, const double *d_blens,
                     const double *d_grad_P, int B, int S, double *d_grad_blens, double *d_W, void *stream);

/* ------------------------------------------------------------------------------------------
 * Alignment ingestion on the device: the unique site patterns of an S x n character matrix in sorted order with
 * their multiplicities -- phylo.compress_alignment (dodonaphy/phylo.py:16-58: Counter(zip(*sequences)), sorted) --
 * by an LSD radix sort of the columns (stable counting-sort passes over groups of taxa), boundary flags and a scan.
 * d_raw uint8 [S][n]: the raw characters, one row per taxon.  dodo_alignment_sort leaves the sorted order in d_ws and
 * writes the number of patterns to d_n_unique (device int32; it synchronises the stream once to read the alphabet
 * size, the caller reads d_n_unique to size the outputs); dodo_alignment_gather then fills d_masks uint8 [S][U]
 * (through d_lut uint8 [256]: character -> IUPAC mask) and d_weights int64 [U].
 * ------------------------------------------------------------------------------------------ */
uint64_t dodo_alignment_workspace_bytes(int S, int64_t n);
int dodo_alignment_sort(const uint8_t *d_raw, int S, int64_t n, void *d_ws, int32_t *d_n_unique, void *stream);
int dodo_alignment_gather(const uint8_t *d_raw, int S, int64_t n, void *d_ws, int32_t n_unique, const uint8_t *d_lut,
                          uint8_t *d_masks, int64_t *d_weights, void *stream);

/* ------------------------------------------------------------------------------------------
 * K1p: batched POINTWISE hyperbolic distances -- n independent pairs (x1[i], x2[i]) of D-dimensional
 * points (D <= 16) -- and their vector-Jacobian product.  Replaces the single-pair utilities the
 * reference calls in Python loops over a tree's branches (BaseModel.compute_branch_lengths,
 * dodonaphy/base_model.py:127-181; Cphylo.compute_branch_lengths_np, dodonaphy/cython/Cphylo.pyx:116-156):
 *   DODO_DIST_POINCARE  Chyp_torch.hyperbolic_distance (dodonaphy/cython/Chyp_torch.pyx:119-134): Poincare-ball
 *                       formula at curvature -1, the Lorentz form otherwise or when the invariant is NaN
 *   DODO_DIST_LORENTZ   Chyp_torch.hyperbolic_distance_lorentz (:137-154) through poincare_to_hyper (:184-210)
 *   DODO_DIST_SHEET_UP  Chyp_np.hyperbolic_distance (dodonaphy/cython/Chyp_np.pyx:279-296): project_up lift
 * d_x1, d_x2 [n][D]; d_curvature device double [1]; d_out [n].  VJP: d_grad_out [n] -> d_grad_x1, d_grad_x2
 * [n][D] and d_grad_curv [n] (one partial per pair; clamped pairs pass no gradient to the locations).
 * ------------------------------------------------------------------------------------------ */
#define DODO_DIST_POINCARE 0
#define DODO_DIST_LORENTZ 1
#define DODO_DIST_SHEET_UP 2
int dodo_hypdist_fwd(const double *d_x1, const double *d_x2, int64_t n, int D, const double *d_curvature,
                     int form, int matsumoto, double *d_out, void *stream);
int dodo_hypdist_bwd(const double *d_x1, const double *d_x2, int64_t n, int D, const double *d_curvature,
                     int form, int matsumoto, const double *d_grad_out, double *d_grad_x1, double *d_grad_x2,
                     double *d_grad_curv, void *stream);

/* ------------------------------------------------------------------------------------------
 * d blens / d locations of a soft-NJ tree in one pass: the (2S-2) x (S D) Jacobian the reference obtains with
 * torch.autograd.functional.jacobian over its get_pdm -> nj_torch chain, 2S-2 backward passes per sample
 * (dodonaphy/vi.py:376-387).  d_peel / d_flags: the outputs of dodo_nj_soft for the same locations.
 * d_J [B][2S-2][S*D].  Selections carry no gradient (dodonaphy/soft.py:56-61); clamped branches pass none.
 * ------------------------------------------------------------------------------------------ */
uint64_t dodo_blens_jacobian_workspace_bytes(int B, int S, int D);
int dodo_blens_jacobian(const double *d_x, int B, int S, int D, const double *d_curvature, int lift, int matsumoto,
                        const int64_t *d_peel, const int32_t *d_flags, void *d_ws, double *d_J, void *stream);

/* log sqrt|det(J J^T)| for B row-major matrices J [B][E][C] (the change-of-variables term of
 * dodonaphy/vi.py:385-387): d_out [B]; -inf when J J^T is exactly singular.  E (E+1) doubles of shared memory. */
int dodo_gram_logdet(const double *d_J, int B, int E, int C, double *d_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * K2: batched neighbour joining, one tree per CTA.
 * dodo_nj_hard replaces Cpeeler.nj_np (dodonaphy/cython/Cpeeler.pyx:15-125): bit-identical peel
 * and blens (same operation order, no FMA contraction, first strict minimum in pool order).
 * dodo_nj_soft replaces peeler.nj_torch (dodonaphy/peeler.py:155-220) with soft.argmin
 * (dodonaphy/soft.py:38-53) evaluated in O(active pairs) per join.
 * ------------------------------------------------------------------------------------------ */
uint64_t dodo_nj_workspace_bytes(int B, int S);
/* Branch-length prior evaluated as the EPILOGUE of the NJ kernel that has just produced the branch
 * lengths (pass NULL for none), or on its own by dodo_prior_gamma_dir: the gamma-Dirichlet(aT, bT, a, c)
 * prior of MrBayes that the reference evaluates per sample on the host --
 * BaseModel.compute_prior_gamma_dir (dodonaphy/base_model.py:475-527) with utils.LogDirPrior
 * (dodonaphy/utils.py:251-267), called at vi.py:436-437 / hmap.py:385-397; NumPy twin
 * Cphylo.compute_prior_gamma_dir_np (dodonaphy/cython/Cphylo.pyx:53-114), called at chain.py:129.
 * d_ln_prior [B] receives the log density incl. normalising constants and the uniform-topology term;
 * d_grad_blens [B][2S-2] (may be NULL) its gradient w.r.t. the branch lengths.  variant: TORCH clamps
 * the branches at 1e-4 when any is <= 0 (utils.py:258-260), NUMPY floors everything at eps (:105-107). */
#define DODO_PRIOR_TORCH 0
#define DODO_PRIOR_NUMPY 1
typedef struct {
    double aT, bT, a, c;
    int32_t variant;
    int32_t _pad;
    double *d_ln_prior;
    double *d_grad_blens;
} dodo_prior_t;
int dodo_prior_gamma_dir(const double *d_blens, int B, int S, const dodo_prior_t *prior, void *stream);
/* d_pdm [B][S][S] (not modified); d_peel int64 [B][S-1][3]; d_blens [B][2S-2] */
int dodo_nj_hard(const double *d_pdm, int B, int S, void *d_ws, int64_t *d_peel, double *d_blens,
                 const dodo_prior_t *prior, void *stream);
/* d_flags int32 [B][S-1]: bit0 = d_pl clamped, bit1 = d_pr clamped, bit2 = exactness guard tripped */
int dodo_nj_soft(const double *d_pdm, int B, int S, double tau, void *d_ws, int64_t *d_peel,
                 double *d_blens, int32_t *d_flags, const dodo_prior_t *prior, void *stream);
/* soft.argmin (dodonaphy/soft.py:38-53) on B dense row-major matrices [B][rows][cols]:
 * d_rowcol [B][2] = (soft_row, soft_col) as doubles, ties resolved towards the last entry. */
int dodo_soft_argmin(const double *d_Q, int B, int rows, int cols, double tau, double *d_rowcol,
                     void *stream);
/* Policy for batched callers (the one-tree route peeler.nj_torch re-runs or raises on the host): every
 * sample whose d_flags [B][S-1] carry a bit of `mask` (4 = exactness guard tripped, 8 = the soft argmin
 * selected an inactive node -- where the reference, peeler.py:196-205, would index a masked row) gets
 * d_loglik[b] = NaN and d_grad_blens[b][:] = NaN (either may be NULL), so a topology the reference would
 * not have produced cannot be consumed silently; d_counts int32 [2] receives the number of samples with
 * bit 2 / bit 3 set.  Enqueued on `stream`, no host synchronisation. */
int dodo_nj_guard(const int32_t *d_flags, int B, int S, uint32_t mask, double *d_loglik,
                  double *d_grad_blens, int32_t *d_counts, void *stream);
/* VJP of blens w.r.t. pdm given the recorded joins: d_grad_pdm [B][S][S] (symmetric) */
int dodo_nj_soft_bwd(const int64_t *d_peel, const int32_t *d_flags, const double *d_grad_blens,
                     int B, int S, void *d_ws, double *d_grad_pdm, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K2g: the "geodesics" tree connector (the reference's other --connect mode with gradients):
 * repeatedly join the two live points of the Poincare ball whose geodesic passes furthest from the
 * origin, the joined node sitting at that geodesic's point nearest the origin.
 *   dodo_geo_soft      replaces peeler.make_soft_peel_tips     dodonaphy/peeler.py:14-70
 *                      (poincare.hyp_lca dodonaphy/poincare.py:62-80, utils.get_plca utils.py:141-171,
 *                       selection by soft.argmin at temperature tau; the reference fixes tau = 1e-8)
 *   dodo_geo_soft_bwd  the vector-Jacobian product torch autograd takes through it
 *   dodo_geo_hard      replaces peeler.make_hard_peel_geodesic dodonaphy/peeler.py:73-152
 * d_x [B][S][D] ball coordinates, D <= 15; d_sheet [B][S][D+1] hyperboloid coordinates.
 * Outputs: d_peel int64 [B][S-1][3]; d_int_locs [B][S-1][D] (soft, ball) or [B][S-1][D+1] (hard,
 * sheet); d_blens [B][2S-2]; d_flags int32 [B][S-1] (bit2 exactness guard, bit3 invalid selection,
 * as dodo_nj_soft).  d_slots int32 [B][S-1][2] and d_tape [B][S-1][2][D] record each join for the
 * backward call.  d_grad_int_locs may be NULL.  Workspace: dodo_geo_workspace_bytes(B, S, D).
 * --------------------------------------------------------------------------------------------- */
uint64_t dodo_geo_workspace_bytes(int B, int S, int D);
int dodo_geo_soft(const double *d_x, int B, int S, int D, double tau, void *d_ws, int64_t *d_peel,
                  double *d_int_locs, double *d_blens, int32_t *d_slots, double *d_tape,
                  int32_t *d_flags, void *stream);
int dodo_geo_soft_bwd(const int64_t *d_peel, const int32_t *d_slots, const double *d_tape,
                      const double *d_grad_blens, const double *d_grad_int_locs, int B, int S, int D,
                      double *d_grad_x, void *stream);
int dodo_geo_hard(const double *d_sheet, int B, int S, int D1, void *d_ws, int64_t *d_peel,
                  double *d_int_locs, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DODONAPHY_B200_H */

/*
 * amg_b200.h -- C ABI of libamgb200.so: the B200-native (sm_100a) solve phase of the
 * txthpc/amg classical AMG solver.
 *
 * The library is a drop-in for the four solve-phase objects of the reference link line
 * (amg/Makefile.sh:19: SSS_SOLVE.o SSS_cycle.o SSS_smooth.o SSS_cuda.o).  Its primary entry
 * point has the reference's own name and signature, so the reference's unmodified C host
 * (main, mmio loaders, SSS_solver_amg -> SSS_amg_setup -> SSS_amg_solve) links against it.
 *
 * All structs below are layout-compatible restatements of the reference's public types
 * (amg/SSS_main.h); the static asserts at the bottom pin the LP64 layout.  When the
 * reference's own header is included first (_SSS_MAIN_H_ defined) the reference-named entry
 * points are declared with the reference's types instead.
 *
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and aborts
 * with a message on stderr when none is usable.
 */
#ifndef AMG_B200_H_
#define AMG_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */

/* ---- layout-compatible types (amg/SSS_main.h) ------------------------------------------ */
typedef struct amgb200_mat_ {   /* SSS_MAT, SSS_main.h:95-105: CSR, int32 indices, fp64 values */
    int num_rows, num_cols, num_nnzs;
    int *row_ptr;
    int *col_idx;
    double *val;
} amgb200_mat;

typedef struct amgb200_vec_ { int n; double *d; } amgb200_vec;   /* SSS_VEC,  SSS_main.h:119-124 */
typedef struct amgb200_ivec_ { int n; int *d; } amgb200_ivec;    /* SSS_IVEC, SSS_main.h:126-131 */

typedef struct amgb200_rtn_ {   /* SSS_RTN, SSS_main.h:154-160 */
    double ares;                /* absolute residual ||b - A x||_2 */
    double rres;                /* relative residual ||r|| / ||b|| */
    int nits;                   /* V-cycles done */
} amgb200_rtn;

typedef struct amgb200_pars_ {  /* SSS_AMG_PARS, SSS_main.h:170-194 */
    int cycle_type;             /* <=1: V-cycle, 2: W-cycle */
    double tol;                 /* outer stopping tolerance on ||r||/||b|| */
    double ctol;                /* coarsest-level tolerance (clamped to 0.1*tol, SSS_cycle.cu:858) */
    int max_it;
    int cs_type;                /* 1 = classical RS */
    int max_levels;
    int coarse_dof;
    int smoother;               /* 2 = Gauss-Seidel (the only smoother the reference implements) */
    double relax;
    int cf_order;               /* 1 = C/F ordering (F rows first, then C rows) */
    int pre_iter, post_iter;
    int poly_deg;
    int interp_type;            /* 1 = direct interpolation */
    double strong_threshold, max_row_sum, trunc_threshold;
} amgb200_pars;

typedef struct amgb200_comp_ {  /* SSS_AMG_COMP, SSS_main.h:196-207: one level */
    amgb200_mat A, R, P;
    amgb200_vec b, x;
    amgb200_ivec cfmark;        /* 0 = F, 1 = C, 2 = isolated */
    amgb200_vec wp;
} amgb200_comp;

typedef struct amgb200_amg_ {   /* SSS_AMG, SSS_main.h:209-218 */
    int num_levels;
    amgb200_comp *cg;
    amgb200_pars pars;
    amgb200_rtn rtn;
} amgb200_amg;

typedef struct amgb200_smtr_ {  /* SSS_SMTR, SSS_main.h:221-238 */
    int smoother;
    amgb200_mat *A;
    amgb200_vec *b;
    amgb200_vec *x;
    double relax;
    int nsweeps, istart, iend, istep, ndeg, cf_order;
    int *ordering;
} amgb200_smtr;

#ifdef _SSS_MAIN_H_
#define AMGB200_T_AMG  SSS_AMG
#define AMGB200_T_VEC  SSS_VEC
#define AMGB200_T_MAT  SSS_MAT
#define AMGB200_T_RTN  SSS_RTN
#define AMGB200_T_SMTR SSS_SMTR
#else
#define AMGB200_T_AMG  amgb200_amg
#define AMGB200_T_VEC  amgb200_vec
#define AMGB200_T_MAT  amgb200_mat
#define AMGB200_T_RTN  amgb200_rtn
#define AMGB200_T_SMTR amgb200_smtr
#endif

/* ---- 1. reference-named drop-in entry points ------------------------------------------- */

/* Replaces amg/Solve/SSS_SOLVE.c:4-87 (declared Solve/SSS_SOLVE.h:9; sole caller SSS_AMG.c:51).
 * Host hierarchy in, solution written to x->d, residual table printed in the reference's
 * format, {ares,rres,nits} returned and stored in mg->rtn, mg->cg[0].x/.b left aliased to the
 * caller's vectors.  Uploads the hierarchy, solves on the device, frees the device mirror. */
AMGB200_T_RTN SSS_amg_solve(AMGB200_T_AMG *mg, AMGB200_T_VEC *x, AMGB200_T_VEC *b);

/* Replaces amg/Solve/SSS_cycle.cu:848-967 (Solve/SSS_cycle.h:18): one V/W-cycle on the host
 * hierarchy (reads cg[0].b, cg[0].x; writes cg[0].x and, like the reference, cg[l].x, cg[l].b (l >= 1) and the residuals cg[l].wp).
 * Every call analyses and uploads the hierarchy again (the reference interface has no resident state): hosts that cycle repeatedly
 * should use amgb200_upload / amgb200_cycle. */
void SSS_amg_cycle(AMGB200_T_AMG *mg);

/* Replaces amg/Solve/SSS_cycle.cu:819-846 (Solve/SSS_cycle.h:17): CG, then GMRES(30) on failure. */
void SSS_amg_coarest_solve(AMGB200_T_MAT *A, AMGB200_T_VEC *b, AMGB200_T_VEC *x, const double ctol);

/* Replace amg/Solve/SSS_smooth.c:138-220 / :223-304 (Solve/SSS_smooth.h:18-20). */
void SSS_amg_smoother_pre(AMGB200_T_SMTR *s);
void SSS_amg_smoother_post(AMGB200_T_SMTR *s);

/* GPU versions of amg/SSS_utils.c:182-201 (y = A x) and :161-178 (y += alpha A x).  New names:
 * the reference's SSS_blas_mv_* live in SSS_utils.o, which stays in the link line. */
void amgb200_blas_mv_mxy(const AMGB200_T_MAT *A, const AMGB200_T_VEC *x, AMGB200_T_VEC *y);
void amgb200_blas_mv_amxpy(double alpha, const AMGB200_T_MAT *A, const AMGB200_T_VEC *x, AMGB200_T_VEC *y);

/* ---- 2. resident-hierarchy API (what SSS_amg_solve is built from) ----------------------- */
typedef struct amgb200_hier amgb200_hier;   /* opaque device mirror of an SSS_AMG */

enum {                           /* amgb200_options.coarse_mode: the reference's CG beta line is UB */
    AMGB200_BETA_FIX = 0,        /* beta = (z_k,r_k)/(z_{k-1},r_{k-1})  (author's stated formula) */
    AMGB200_BETA_AS_COMPILED = 1 /* beta = 1, temp1 frozen              (what the shipped object runs) */
};

typedef struct amgb200_options_ {
    int coarse_mode;             /* AMGB200_BETA_* (default FIX; env AMGB200_COARSE_MODE=asc overrides) */
    int verbose;                 /* 0 quiet, 1 reference's iteration table, 2 + level/kernel table */
    int device;                  /* CUDA device ordinal, -1 = current */
    int fast;                    /* 0 (default): EXACT arithmetic -- every sum that feeds x is accumulated in the
                                    reference's order, results bit-identical to the reference's CPU path;
                                    1: FAST -- long rows and Krylov dot products use tree reductions (~1e-16 per
                                    operation, amplified by |x|/|r| to ~1e-7 in the last residuals).
                                    env AMGB200_FAST=1 overrides */
    int level0_worker;           /* multi-GPU: 1 = this device only runs the row-block kernels of level 0 (ranks >= 1 of the sharded
                                    solve): A_0, P_0 and the vectors of levels 0 and 1 (in level 1's real schedule numbering) become
                                    resident, nothing below; amgb200_num_levels() returns 2 */
    int reserved[3];
} amgb200_options;

void amgb200_default_options(amgb200_options *o);

/* Analyse (wavefront schedule, per-level kernel choice) and upload.  Host hierarchy is only read. */
amgb200_hier *amgb200_upload(const AMGB200_T_AMG *mg, const amgb200_options *opt);
void amgb200_free(amgb200_hier *h);

/* Full solve on a resident hierarchy; x, b host arrays of n_0 doubles (natural numbering).
 * res_hist (may be NULL) receives ||r||_2 after each V-cycle (up to hist_cap entries). */
amgb200_rtn amgb200_solve(amgb200_hier *h, double *x, const double *b, double *res_hist, int hist_cap);

/* Same, but x/b already on the device (natural numbering, device pointers); nothing crosses
 * PCIe except one double per V-cycle.  Used for the kernel-only bench number. */
amgb200_rtn amgb200_solve_device(amgb200_hier *h, double *d_x, const double *d_b, double *res_hist, int hist_cap);

/* One V-cycle on resident level-0 vectors (host arrays in/out), for function-level parity. */
void amgb200_cycle(amgb200_hier *h, double *x, const double *b);

/* Function-level hooks on one resident level (host arrays, natural numbering):
 *   which: 0 = A_l, 1 = P_l (coarse -> fine), 2 = R_l (fine -> coarse)
 *   y = beta*y + alpha * M x   with beta in {0,1}  (beta=0: SSS_blas_mv_mxy, beta=1: _amxpy) */
void amgb200_level_spmv(amgb200_hier *h, int level, int which, double alpha, const double *x, int beta, double *y);
/* |nsweeps| Gauss-Seidel sweeps on level l as the cycle applies them: C/F-ordered (SSS_smooth.c:4-87, order != 0 branch)
 * or, when the hierarchy was set up with cf_order = 0, natural order (SSS_smooth.c:90-137): nsweeps > 0 = pre-smoothing
 * (forward), nsweeps < 0 = post-smoothing (backward) */
void amgb200_level_smooth(amgb200_hier *h, int level, int nsweeps, double *x, const double *b);
/* r = b - A_l x, returns ||r||_2 */
double amgb200_level_residual(amgb200_hier *h, int level, const double *x, const double *b, double *r);
/* what the cycle does between pre-smoothing and the next level (SSS_cycle.cu:916-921, 929): r = b - A_l x, bc = R_l r (and
 * x_{l+1} = 0) -- ONE launch on levels whose A and R are thread-per-row layouts (resid_restrict_kernel; r is written once and
 * re-read from L2), else the separate kernels; returns 1 when the fused kernel ran.  r: n_l doubles, bc: n_{l+1} doubles */
int amgb200_level_resid_restrict(amgb200_hier *h, int level, const double *x, const double *b, double *r, double *bc);
/* coarsest-level solve on the resident coarsest matrix; returns the Krylov status
 * (iterations, or a negative SSS error code) of the last solver that ran; its[0]=CG its/status,
 * its[1]=GMRES its/status or 0 if not run */
int amgb200_coarse_solve(amgb200_hier *h, double *x, const double *b, double tol, int its[2]);

/* Introspection */
int amgb200_num_levels(const amgb200_hier *h);
/* info[0]=rows info[1]=nnz info[2]=F-pass wavefronts info[3]=C-pass wavefronts
 * info[4]=kernel kind (0 SELL-32 thread/row, 1 CSR warp/row) info[5]=rows in F pass
 * info[6]=P nnz info[7]=R nnz */
void amgb200_level_info(const amgb200_hier *h, int level, long long info[8]);
/* algorithmic bytes (SURVEY.md section 8d formulas) of: op 0 = one GS sweep, 1 = residual,
 * 2 = restrict, 3 = prolong-add, 4 = y=A x  on that level; op 5 (level ignored) = one V-cycle;
 * op 6 = residual (+) restriction (+ zero-fill of the coarse x) with r not counted (the fused figure of SURVEY.md 8d) */
double amgb200_algorithmic_bytes(const amgb200_hier *h, int level, int op);
/* time `reps` back-to-back launches of op (as above, 0..4, 6) on level l with CUDA events on the
 * library's stream; returns average milliseconds per launch (ops run on scratch vectors) */
double amgb200_time_op(amgb200_hier *h, int level, int op, int reps);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
long long amgb200_launch_count(void);
/* per-phase device time of the last solve in ms: [0] GS, [1] residual, [2] restrict,
 * [3] prolong, [4] coarse solve, [5] outer residual+norm, [6] total (only when
 * AMGB200_PROFILE=1 in the environment; otherwise zeros) */
void amgb200_last_phase_ms(const amgb200_hier *h, double ms[8]);
/* same, per level: [0] GS, [1] residual, [2] restrict, [3] prolong */
void amgb200_last_level_ms(const amgb200_hier *h, int level, double ms[4]);
void amgb200_set_profile(amgb200_hier *h, int on);
/* s[0] = host analysis seconds, s[1] = analysis + upload seconds of amgb200_upload */
void amgb200_upload_seconds(const amgb200_hier *h, double s[2]);
long long amgb200_device_bytes(const amgb200_hier *h);
/* name of the Gauss-Seidel kernel chosen for that level */
const char *amgb200_level_kernel(const amgb200_hier *h, int level);
/* 1 when the cycle runs residual (+) restriction of this level as one launch (resid_restrict_kernel) */
int amgb200_level_fused(const amgb200_hier *h, int level);
/* ordered (row-order Gauss-Seidel) levels: sum over the dependency wavefronts of one sweep of the longest in-order chain that can
 * only start once the previous wavefront is complete, in terms (one dependent fp64 subtraction each): the latency floor of a
 * sweep under the reference's rounding order (Solve/SSS_smooth.c:22-29) is this many terms x the fp64 add latency; 0 for
 * levels whose passes are one wavefront */
long long amgb200_level_chain_terms(const amgb200_hier *h, int level);
/* `warmup` untimed + `steps` timed solves from d_x0 (device arrays, natural numbering), timed with
 * CUDA events on the library's stream; *ms_total = milliseconds for the `steps` solves */
void amgb200_bench_solve(amgb200_hier *h, const double *d_x0, const double *d_b, double *d_x, int warmup, int steps,
                         double *ms_total, amgb200_rtn *last);
const char *amgb200_version(void);

/* ---- 2b. multi-GPU building blocks (one process per GPU; driven by amg_b200/distributed.py) -------
 * Only level 0 of two-colour problems shards without a cross-device dependency chain:
 * every rank holds the resident hierarchy, runs the level-0 kernels on its own item range and exchanges
 * ghost x entries; rank 0 runs the levels below.  All ranges are item ranges of the level-0 layouts. */
void amgb200_set_stream(amgb200_hier *h, void *cuda_stream);            /* run on the caller's stream */
void *amgb200_level_vec(amgb200_hier *h, int level, int which);         /* device ptr: 0 x, 1 b, 2 wp (schedule numbering) */
void amgb200_level_order(const amgb200_hier *h, int level, int *order_host);   /* schedule position -> natural row */
void amgb200_level_download(amgb200_hier *h, int level, int which, double *host);  /* level vector (0 x, 1 b, 2 wp), natural numbering */
void amgb200_l0_shape(const amgb200_hier *h, long long info[8]);
void amgb200_l0_gs_pass(amgb200_hier *h, int pass, int item0, int item1);
void amgb200_l0_residual(amgb200_hier *h, int item0, int item1);
void amgb200_l0_prolong(amgb200_hier *h, int item0, int item1);
void amgb200_restrict_from(amgb200_hier *h, int level);
void amgb200_cycle_from(amgb200_hier *h, int level);
void amgb200_vec_to_schedule(amgb200_hier *h, int level, const double *d_nat, double *d_sched);
void amgb200_vec_to_natural(amgb200_hier *h, int level, const double *d_sched, double *d_nat);
void amgb200_sync(amgb200_hier *h);
/* Peer-memory exchange over NVLink / NVSwitch (CUDA IPC, one process per GPU): ghost entries are stored straight into the peer's
 * vector and an epoch into the peer's flag word; the receiving stream waits on its own flag words.  No host round trip.
 *   amgb200_ipc_export: 64-byte handle of a level vector (which: 0 x, 1 b, 2 wp) or of this rank's flag words (which = 3)
 *   amgb200_ipc_open:   map a peer's handle (closed by amgb200_free)
 *   amgb200_peer_plan:  plan 0..7 = npush transfers my_vec[idx] -> peer_vec[idx] (idx[i] NULL: contiguous range starting at
 *                       range0[i], count[i] long), nflag flag slots to raise afterwards (addresses peer_flags + plan*64 + my rank),
 *                       nsrc source ranks to wait for
 *   amgb200_peer_run:   enqueue push + flags + wait on the hierarchy's stream; all ranks run their plans in the same global order */
void amgb200_ipc_export(amgb200_hier *h, int level, int which, unsigned char *handle);
void *amgb200_ipc_open(amgb200_hier *h, const unsigned char *handle);
void amgb200_peer_plan(amgb200_hier *h, int plan, int npush, const double *my_vec, void *const *peer_vec, const int *const *idx, const int *count,
                       const int *range0, int nflag, void *const *flag_slot, int nsrc, const int *src);
void amgb200_peer_run(amgb200_hier *h, int plan);
/* the two halves of amgb200_peer_run: start = store into the peers + raise the epochs in their flag words, wait = for the peers' epochs;
 * kernels launched in between overlap the exchange (the interior rows of a level-0 pass read no ghost entry) */
void amgb200_peer_start(amgb200_hier *h, int plan);
void amgb200_peer_wait(amgb200_hier *h, int plan);

/* ---- 3. host-side helpers (pure C++, no device): synthetic operators + RS setup ---------- */
/* Synthetic level-0 operators of SURVEY.md Appendix B, CSR with ascending columns:
 *   kind 0 = p2d N (5-point), 1 = p3d N (7-point), 2 = aniso3d N (1,1,eps_z), 3 = v27 N
 * Returns 0 on success; arrays are malloc'ed, release with amgb200_mat_free. */
int amgb200_generate(int kind, int N, double eps_z, amgb200_mat *A);
void amgb200_mat_free(amgb200_mat *A);
/* Loader fast path for the kept C host: replaces the body of SSS_mat_read (amg/SSS_main.c:12-22: mmio_info +
 * mmio_data, two fscanf passes over the file, amg/mmio_highlevel.h:10-305) by ONE multi-threaded pass with the same semantics
 * entry for entry (file order inside each row, symmetric / hermitian expansion, pattern -> 1.0, imaginary parts dropped, no
 * sorting, no duplicate merging).  Returns 0, or -1 / -2 / -4 / -5 (open / banner / size line / entries).
 * AMGB200_MTX_CACHE=1 keeps a raw CSR cache `<file>.amgb200cache` validated against the file's size and mtime. */
int amgb200_read_mtx(const char *filename, amgb200_mat *A);
/* From-scratch restatement of the reference's setup phase (Setup/SSS_SETUP.cu:36-177 and below:
 * RS coarsening, direct interpolation + truncation, R = P^T, Galerkin RAP) producing a host
 * hierarchy that is bit-identical to the reference's.  Needed so the product runs without any
 * reference object; the reference's own setup can be used instead (it is the kept host). */
void amgb200_setup(amgb200_amg *mg, const amgb200_mat *A, const amgb200_pars *pars, int verbose);
/* flags & AMGB200_SETUP_DEVICE_INTERP: the interpolation weights, the coarse renumbering and the truncation of P are computed on
 * the device (amgb200_interp_device) instead of the host loop -- the same hierarchy, bit for bit. */
#define AMGB200_SETUP_DEVICE_INTERP 1
/* flags & AMGB200_SETUP_DEVICE_RAP: R = P^T and the Galerkin product R A P are computed on the device (amgb200_rap_device) -- the same
 * arrays, entry for entry. */
#define AMGB200_SETUP_DEVICE_RAP 2
void amgb200_setup_ex(amgb200_amg *mg, const amgb200_mat *A, const amgb200_pars *pars, int verbose, int flags);
/* ---- 3b. setup step next to the hot path, on the device ---------
 * Direct-interpolation weights + coarse numbering + truncation: replaces interp_DIR / interp_DIR_cuda + SSS_amg_interp_trunc
 * (amg/Setup/SSS_inter.cu:400-547, :239-396, :16-102; the reference's kernel DIR_Step_1, :104-210, covers 131 072 rows only).
 * P enters with the pattern the coarsening produced (col_idx = FINE indices of the interpolatory C points, val allocated) and
 * leaves exactly as interp_DIR leaves it.  mark: 0 F, 1 C, 2 isolated.  Returns 0; 1 = a row has no stored diagonal, nothing done. */
int amgb200_interp_device(const amgb200_mat *A, const int *mark, amgb200_mat *P, double trunc_threshold);
/* ---- 3c. the operators of the next level, on the device ---------
 * R = P^T (replaces SSS_mat_trans, amg/SSS_matvec.c:330-387) and Ac = R A P (replaces SSS_blas_mat_rap, amg/SSS_matvec.c:398-534) with the
 * reference's output ORDER -- rows of R by ascending fine row; rows of Ac: diagonal slot first, then the columns in discovery order of
 * the triple loop, values accumulated in traversal order as (r*a)*p without FMA -- because that order is the summation order of the
 * solve phase.  R_out / Ac_out receive calloc'ed host CSR arrays owned by the caller (freed by SSS_amg_data_destroy like the
 * reference's).  Returns 0; 1 = a product row exceeds 8192 columns or the product 2^31 entries: nothing returned, use the host loops. */
int amgb200_rap_device(const amgb200_mat *A, const amgb200_mat *P, amgb200_mat *R_out, amgb200_mat *Ac_out);
/* multi-GPU harness helper (host only, no reference counterpart): ghost lists of a row-block partition of level 0.  order[n]: schedule
 * position -> natural row; f_bounds / c_bounds: world + 1 ascending schedule-row offsets of the ranks' F / C blocks.  *idx receives the
 * concatenated lists (malloc'ed; free()), ptr[world*world*2 + 1] their offsets, list key = (reader * world + owner) * 2 + pass of the
 * ghost entry (0 F, 1 C); every list holds ascending, distinct schedule indices.  Returns the total length. */
long long amgb200_ghost_lists(const amgb200_mat *A, const int *order, int nF, int world, const long long *f_bounds, const long long *c_bounds, int **idx,
                              long long *ptr);
/* the same, and reads_ghost[k] = 1 for every schedule row k that reads an entry owned by another rank (n bytes, zeroed by the callee; may be NULL):
 * the rows of a pass that have to wait for the halo exchange */
long long amgb200_ghost_lists_ex(const amgb200_mat *A, const int *order, int nF, int world, const long long *f_bounds, const long long *c_bounds, int **idx,
                                 long long *ptr, unsigned char *reads_ghost);
void amgb200_amg_destroy(amgb200_amg *mg);
void amgb200_default_pars(amgb200_pars *p);   /* SSS_main.c:25-64 */

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif

/* ---- layout pins (LP64) ------------------------------------------------------------------ */
#if defined(__cplusplus)
#define AMGB200_SA(c, m) static_assert(c, m)
#else
#define AMGB200_SA(c, m) _Static_assert(c, m)
#endif
AMGB200_SA(sizeof(amgb200_mat) == 40, "SSS_MAT layout");
AMGB200_SA(offsetof(amgb200_mat, row_ptr) == 16 && offsetof(amgb200_mat, val) == 32, "SSS_MAT layout");
AMGB200_SA(sizeof(amgb200_vec) == 16 && sizeof(amgb200_ivec) == 16, "SSS_VEC layout");
AMGB200_SA(sizeof(amgb200_rtn) == 24, "SSS_RTN layout");
AMGB200_SA(sizeof(amgb200_pars) == 104, "SSS_AMG_PARS layout");
AMGB200_SA(sizeof(amgb200_comp) == 184 && offsetof(amgb200_comp, R) == 40 && offsetof(amgb200_comp, P) == 80 &&
           offsetof(amgb200_comp, b) == 120 && offsetof(amgb200_comp, x) == 136 &&
           offsetof(amgb200_comp, cfmark) == 152 && offsetof(amgb200_comp, wp) == 168, "SSS_AMG_COMP layout");
AMGB200_SA(sizeof(amgb200_amg) == 144 && offsetof(amgb200_amg, cg) == 8 && offsetof(amgb200_amg, pars) == 16 &&
           offsetof(amgb200_amg, rtn) == 120, "SSS_AMG layout");
AMGB200_SA(sizeof(amgb200_smtr) == 72, "SSS_SMTR layout");

#endif /* AMG_B200_H_ */

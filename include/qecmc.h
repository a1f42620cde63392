/*
 * qecmc.h -- C ABI of libqecmc.so: B200 (sm_100a) Metropolis-chain decoders for
 * toric / planar / rotated-surface / XZZX codes.
 *
 * The reference (QEC-project-2020/MCMC-QEC-toric-RL) has no FFI: its seams are the
 * Python callables in decoders.py / decoders_biasednoise.py / src/mcmc*.py.  Each
 * entry point below names the reference callable whose hot loop it replaces; the
 * Python mirror in mcmc-qec-toric-rl_b200/ keeps the reference signatures and calls
 * these through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions: every function returns 0 on success or a negative qecmc_status and
 * records a message retrievable with qecmc_last_error() (thread-local).  Pointers
 * are caller-owned and never retained after return.  Functions without a _dev
 * suffix take HOST buffers and perform the host<->device copies themselves; _dev
 * variants take DEVICE buffers and only enqueue work on the context's stream
 * (then synchronise before returning results that live on the host, e.g. stats).
 * A context is bound to one CUDA device and is not thread-safe: one context per
 * GPU, driven by one host thread each.  There is no CPU fallback: without a CUDA
 * device qecmc_create fails.
 *
 * Lattices cross the ABI in the reference's layout: uint8, C order, values 0..3
 * (0=I 1=X 2=Y 3=Z), shape (2,L,L) for toric/planar and (L,L) for rotated/XZZX.
 */
#ifndef QECMC_H
#define QECMC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QECMC_ABI_VERSION 1

typedef struct qecmc_ctx qecmc_ctx;

enum qecmc_geometry { QECMC_TORIC = 0, QECMC_PLANAR = 1, QECMC_ROTATED = 2, QECMC_XZZX = 3 };

enum qecmc_status {
    QECMC_OK = 0,
    QECMC_ERR_ARG = -1,     /* invalid argument (size, geometry, null pointer, bad lattice value) */
    QECMC_ERR_CUDA = -2,    /* CUDA runtime error */
    QECMC_ERR_NOMEM = -3,   /* distinct-chain tables for even one syndrome do not fit in device memory */
    QECMC_ERR_UNSUPPORTED = -4
};

/* which pow() the reference would have used for the acceptance thresholds */
enum qecmc_pow { QECMC_POW_NUMBA = 0 /* njit: square-and-multiply */, QECMC_POW_LIBM = 1 /* CPython float**int */ };

typedef struct qecmc_devinfo {
    int32_t device, sm_count, sm_clock_khz, cc_major, cc_minor;
    int64_t total_mem, free_mem;
    char name[64];
} qecmc_devinfo;

typedef struct qecmc_stats {
    int64_t metropolis_steps;   /* proposals evaluated by the chain kernels            */
    int64_t accepted;           /* proposals accepted                                  */
    int64_t samples;            /* lattice states offered to the distinct-chain sets   */
    int64_t distinct;           /* distinct chains over all (syndrome, class) sets     */
    int64_t table_slots;        /* open-addressing slots per (syndrome, class) set; 0: per-chain key logs, -1: bucket logs */
    int64_t waves;              /* kernel waves the batch was split into               */
    int64_t kernel_launches;    /* CUDA kernels launched by this call                  */
    double  chain_kernel_ms;    /* device time of the chain kernels (CUDA events)      */
    double  total_ms;           /* device time of the whole call on the ctx stream     */
} qecmc_stats;

int         qecmc_abi_version(void);
const char *qecmc_last_error(void);
int         qecmc_create(int device, qecmc_ctx **out);
void        qecmc_destroy(qecmc_ctx *ctx);
/* run on an existing cudaStream_t (e.g. torch's current stream); NULL = the context's own stream */
int         qecmc_set_stream(qecmc_ctx *ctx, void *cuda_stream);
int         qecmc_device_info(qecmc_ctx *ctx, qecmc_devinfo *out);
/* cap (bytes) on the distinct-chain table arena; 0 = 85 % of free device memory */
int         qecmc_set_table_budget(qecmc_ctx *ctx, int64_t bytes);
/* How the most recent qecmc_stdc / qecmc_strc / qecmc_single_temp call (host or _dev form) was sized, for callers that cut a
 * long job into batches: wave_capacity = syndromes one wave may hold within the table budget (the call splits a larger batch
 * into waves), round_chains = chains one round of CTAs over all SMs holds at the kernel's full CTA size.  A batch of
 * min(wave_capacity, k * round_chains / (classes * droplets)) syndromes leaves no SM idle behind a short last round. */
int         qecmc_last_plan(qecmc_ctx *ctx, int64_t *wave_capacity, int64_t *round_chains);
/* Test switches, per context (never read from the environment).  They select between code paths that must give the
 * same results, so that tests can compare them on one problem; value < 0 restores the default.
 *   "force_wide"   1: 64-bit row words also for L <= 16
 *   "insert_mode"  distinct-chain accounting of STDC / STRC: 0 synchronous HBM set, 2 deferred HBM set, 4 per-chain
 *                  key logs, 6 bucket logs
 *   "serial_sweep" 1: native ladders walk the swap sweep pair by pair, like replay does (warp-per-ladder kernel)
 *   "ladder_kernel" 1: native ladders run on the warp-per-ladder kernel replay uses, not on the rung-major one
 *   "pt_grid"      cap on the rung-major kernel's grid, so that ladders queue up behind few CTAs
 *   "packed"       0: two-layer codes with 17 <= L <= 24 stay on the 64-bit row-word chain kernel instead of the packed-lattice
 *                  one; 2 / 4 / 8: interleaved copies of its hot tables
 *   "pt_lt"        lanes sharing one top-rung replica in the rung-major tempering kernel (2, 4, 8, ...)
 * Unknown keys return QECMC_ERR_ARG. */
int         qecmc_debug_set(qecmc_ctx *ctx, const char *key, int64_t value);

/* ------------------------------------------------------------------------------
 * Single-temperature chains: Chain.update_chain_fast / _update_chain_fast
 * (src/mcmc.py:45-46,152-160) and, with QECMC_POW_LIBM, the p_logical == 0 branch of
 * Chain.update_chain (src/mcmc.py:36-43), over a batch of independent chains.
 * ------------------------------------------------------------------------------ */
typedef struct qecmc_chain_cfg {
    int32_t geom_chain;    /* proposal geometry. QECMC_PLANAR on a (2,L,L) toric lattice reproduces the
                              reference's as-shipped fast path */
    int32_t L;
    int32_t pow_kind;      /* enum qecmc_pow */
    int32_t reserved;
    double  p;             /* sampling error rate; factor = (p/3)/(1-p) (src/mcmc.py:16) */
    uint64_t seed;         /* Philox key (native mode) */
    uint64_t stream_offset;/* Philox counter offset: Metropolis steps already taken by these chains */
} qecmc_chain_cfg;

/* Native Philox4x32-10: advance `chains` lattices by `iters` Metropolis steps in place. */
int qecmc_chain_update(qecmc_ctx *ctx, const qecmc_chain_cfg *cfg, uint8_t *qm /*[chains][n_sites] in/out*/,
                       int64_t chains, int64_t iters, qecmc_stats *stats);

/* Replay: explicit uniforms in, full trace out; bit-exact against the reference.
 * u is [chains][iters][k+1]: the k proposal draws in the reference's order
 * (k = 3 toric/planar, 5 rotated/XZZX) followed by the accept draw.
 * dE / accepted ([chains][iters]) and traj ([chains][iters][n_sites], state after each step) are optional. */
int qecmc_replay_chain(qecmc_ctx *ctx, const qecmc_chain_cfg *cfg, const uint8_t *qm0, const double *u,
                       int64_t chains, int64_t iters, uint8_t *qm_final, int8_t *dE, uint8_t *accepted,
                       uint8_t *traj);

/* ------------------------------------------------------------------------------
 * STDC / STDC_droplet (decoders.py:236-322): per syndrome and equivalence class,
 * `droplets` chains of `steps` samples (`iters` Metropolis steps each), the union of
 * distinct chains, Z_E = sum exp(-beta n), normalised to percent.
 * ------------------------------------------------------------------------------ */
typedef struct qecmc_stdc_cfg {
    int32_t geom_code;       /* geometry of the code object: class labels, to_class, rain */
    int32_t geom_chain;      /* proposal geometry of the chains (see qecmc_chain_cfg) */
    int32_t L;
    int32_t droplets;
    int32_t iters;           /* Metropolis steps per sample; the reference uses 5 (decoders.py:250) */
    int32_t per_class_inits; /* 0: qm is [S][n_sites], classes reached on device (to_class);
                                1: qm is [S][n_eq][n_sites], the "list of init codes" form (decoders.py:272-279) */
    int32_t randomize;       /* apply_stabilizers_uniform() before sampling (decoders.py:245-246) */
    int32_t reserved;
    int64_t steps;           /* samples per droplet */
    double  p_error;
    double  p_sampling;
    double  conv_mult;       /* early-stop rule of decoders.py:257-263; 0 disables */
    uint64_t seed;
    /* replay mode when u_nb != NULL (host pointers in qecmc_stdc, device pointers in qecmc_stdc_dev):
       u_nb [S*n_eq*droplets][steps*iters*4]  numba-stream draws of each droplet's chain,
       u_np [S*n_eq*droplets][2*L*L]          numpy-stream draws of each droplet's rain (if randomize). */
    const double *u_nb;
    const double *u_np;
} qecmc_stdc_cfg;

/* eqdistr [S][n_eq] percent (float64, as the reference returns);
 * N_hist  [S][n_eq][n_sites+1] distinct chains per length, optional (NULL to skip). */
int qecmc_stdc(qecmc_ctx *ctx, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
               uint32_t *N_hist, qecmc_stats *stats);
int qecmc_stdc_dev(qecmc_ctx *ctx, const qecmc_stdc_cfg *cfg, const uint8_t *d_qm, int64_t S, double *d_eqdistr,
                   uint32_t *d_N_hist, qecmc_stats *stats);

/* ------------------------------------------------------------------------------
 * STRC / STRC_droplet (decoders.py:745-949): same chains and distinct set as STDC plus
 * the visit histogram m(n) (every sample counted) and, per droplet, the shortest and
 * next-shortest visited lengths; Z_E = sum_l m(l) exp(-beta_s*shortest + (beta_s-beta)*l)
 * times the mean distinct/visited fraction of the two shortest lengths (decoders.py:931-946),
 * droplets merged in the reference's order (decoders.py:882-928).
 *   m_hist     [S][n_eq][n_sites+1]  optional
 *   short_info [S][n_eq][4] = (shortest, next_shortest, #distinct shortest, #distinct next-shortest), optional
 * ------------------------------------------------------------------------------ */
int qecmc_strc(qecmc_ctx *ctx, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
               uint64_t *m_hist, int32_t *short_info, qecmc_stats *stats);

/* ------------------------------------------------------------------------------
 * single_temp (decoders.py:108-135): one chain per class at p = cfg->p_sampling, cfg->steps =
 * max_iters samples of cfg->iters (5) fast-path steps; mean_length [S][n_eq] = average chain
 * length over the first max_iters-1 samples.  droplets must be 1; p_error is ignored.
 * ------------------------------------------------------------------------------ */
int qecmc_single_temp(qecmc_ctx *ctx, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *mean_length,
                      qecmc_stats *stats);

/* ------------------------------------------------------------------------------
 * Parallel-tempering ladders: Ladder / Ladder_alpha / Ladder_biased (src/mcmc.py:49-103,
 * src/mcmc_alpha.py:77-137, src/mcmc_biased.py:66-124) over the slow-path chains
 * Chain / Chain_alpha / Chain_biased .update_chain (mcmc.py:19-43, mcmc_alpha.py:27-70,
 * mcmc_biased.py:21-59; the denominator of the alpha/biased accept ratio is frozen per call
 * as in the reference).  Replica swaps happen on the device.
 * ------------------------------------------------------------------------------ */
enum qecmc_ladder_kind { QECMC_LADDER_DEPOLARIZING = 0, QECMC_LADDER_ALPHA = 1, QECMC_LADDER_BIASED = 2 };

typedef struct qecmc_ladder_cfg {
    int32_t geom;        /* the code's own geometry (the slow path proposes with it) */
    int32_t L;
    int32_t kind;        /* enum qecmc_ladder_kind */
    int32_t Nc;          /* rungs, 1..32; rung p's = numpy.linspace(bottom, top, Nc), top = 0.75 / 1 / (eta+1)/(2eta+1) */
    int32_t iters;       /* Metropolis steps per rung between swap sweeps (the decoders use 10) */
    int32_t reserved;
    double  bottom;      /* p (depolarizing, biased) or pz_tilde (alpha) of rung 0 */
    double  param_b;     /* alpha (kind 1) or eta (kind 2); unused for kind 0 */
    double  p_logical;   /* top rung proposes a random logical operator with this probability */
    uint64_t seed;       /* Philox key (native mode) */
    /* replay mode when u_nb != NULL: per ladder, the numba-stream and CPython-stream uniforms in the
       order the reference consumes them: u_nb [ladders][n_nb], u_py [ladders][n_py] */
    const double *u_nb, *u_py;
    int64_t n_nb, n_py;
} qecmc_ladder_cfg;

/* `steps` calls of Ladder.step(iters) on S ladders.  Fresh start (resume == 0): every rung of ladder s
 * starts from qm0[s] ([S][n_sites]), only the top rung flagged, tops0 = 0.  resume != 0: the ladders
 * continue from rung_states / flags / tops0 (/ n_eff_parts), which are then in/out -- this is what the
 * Python Ladder objects use between .step() calls.  All arrays are in rung order (rung 0 = coldest);
 * every output is optional. */
typedef struct qecmc_ladder_io {
    const uint8_t *qm0;       /* [S][n_sites] */
    int32_t resume, reserved;
    uint8_t *rung_states;     /* [S][Nc][n_sites] */
    int32_t *flags;           /* [S][Nc] */
    int64_t *tops0;           /* [S] */
    double  *n_eff;           /* [S][Nc] alpha ladders: the rung-owned n_eff = nz + alpha (nx + ny) (out only) */
    int32_t *n_eff_parts;     /* [S][Nc][2] = (nz, nx + ny) of that n_eff, exact (in/out with resume) */
    uint8_t *snap_states;     /* tests: [S][steps][Nc][n_sites] after every step */
    int32_t *snap_flags;      /*        [S][steps][Nc] */
    int64_t *snap_tops0;      /*        [S][steps] */
} qecmc_ladder_io;
int qecmc_ladder_run(qecmc_ctx *ctx, const qecmc_ladder_cfg *cfg, const qecmc_ladder_io *io, int64_t S, int64_t steps,
                     qecmc_stats *stats);

/* PTEQ / PTEQ_biased / PTEQ_alpha (decoders.py:25-105, decoders_biasednoise.py:28-90,175-237). */
typedef struct qecmc_pteq_cfg {
    qecmc_ladder_cfg ladder;
    int32_t SEQ, TOPS, tops_burn;
    int32_t use_conv;    /* 1: conv_criteria='error_based' (decoders.py:93-105); 0: run all `steps` */
    double  eps;
    int64_t steps;       /* cap on Ladder.step calls per syndrome (steps * iters < 2^32).  The history the criterion needs is
                          * 2 bytes per step (4 for alpha ladders) per ladder IN FLIGHT: with native draws a finished ladder
                          * hands its place to the next one of the batch, so the reference's default of 50000000 runs */
} qecmc_pteq_cfg;

/* eqdistr [S][n_eq] uint8 = (class counts / (since_burn + 1) * 100) truncated, as the reference returns;
 * eq_counts [S][n_eq] optional; info [S][4] = (steps used, since_burn, tops0, converged) optional. */
int qecmc_pteq(qecmc_ctx *ctx, const qecmc_pteq_cfg *cfg, const uint8_t *qm, int64_t S, uint8_t *eqdistr,
               int64_t *eq_counts, int64_t *info, qecmc_stats *stats);
/* PTEQ_alpha_with_shortest (decoders_biasednoise.py:93-172), for any ladder kind: PTEQ plus, per class, the
 * smallest value recorded for the bottom rung (n_eff for alpha ladders, the chain length otherwise; 100000 when the
 * class was never visited), the number of samples at that value and the number of distinct bottom-rung states seen
 * at it.  All three [S][n_eq].  The reference's second and third return values follow as
 * unique * exp(log(pz_tilde) * short_len) and short_n, each normalised to percent. */
int qecmc_pteq_shortest(qecmc_ctx *ctx, const qecmc_pteq_cfg *cfg, const uint8_t *qm, int64_t S, uint8_t *eqdistr,
                        double *short_len, int64_t *short_n, int64_t *short_unique, int64_t *info, qecmc_stats *stats);
/* same with DEVICE qm / eqdistr; info stays a host pointer */
int qecmc_pteq_dev(qecmc_ctx *ctx, const qecmc_pteq_cfg *cfg, const uint8_t *d_qm, int64_t S, uint8_t *d_eqdistr,
                   int64_t *info, qecmc_stats *stats);

/* PTDC / PTDC_droplet (decoders.py:138-233): per syndrome and class, `droplets` ladders (p_logical = 0)
 * of `steps` Ladder.step(iters) calls, every rung's state offered to the class's distinct-chain set
 * after every step; Z_E = sum exp(-beta n).  eqdistr [S][n_eq] in percent as float64 (the reference
 * truncates to uint8; the Python mirror does that). */
typedef struct qecmc_ptdc_cfg {
    qecmc_ladder_cfg ladder;    /* kind must be depolarizing; bottom = p_sampling */
    int32_t droplets;
    int32_t per_class_inits;
    int64_t steps;              /* already divided by Nc (decoders.py:196) */
    double  p_error;
    double  conv_mult;          /* PTDC only: early stop of PTDC_droplet (decoders.py:156-161); 0 disables */
    int64_t *steps_done;        /* optional out (host): [S][n_eq][droplets] Ladder.step calls each droplet made */
} qecmc_ptdc_cfg;
int qecmc_ptdc(qecmc_ctx *ctx, const qecmc_ptdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr, qecmc_stats *stats);

/* PTRC / PTRC_droplet (decoders.py:584-742): the same ladders, but every rung keeps its own set of distinct
 * chains with N(n) (distinct chains per length) and m(n) (visits per length), droplet counters summed
 * (decoders.py:699-718); Z_E = sum over rungs i < Nc-1 of C_i * sum_n m_i(n) exp(n (beta_i - beta) - beta_i n0_i)
 * with C_i the mean of N/m exp(-beta_i (n - n0)) over the two shortest lengths seen (decoders.py:721-739).
 * conv_mult has no effect in the reference (its break is commented out) and is not a parameter here.
 * N_hist, m_hist [S][n_eq][Nc][n_sites+1] optional. */
int qecmc_ptrc(qecmc_ctx *ctx, const qecmc_ptdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
               int64_t *N_hist, int64_t *m_hist, qecmc_stats *stats);

/* EWD-style STDC_Nall_n_alpha / STDC_droplet_alpha (decoders.py:510-581): one Chain_alpha per class,
 * `steps` samples of Chain_alpha.update_chain(iters), distinct chains weighted by
 * pz_tilde ** (nz + alpha (nx + ny)).  distinct [S][n_eq] optional. */
typedef struct qecmc_alpha_cfg {
    int32_t geom, L;
    int32_t iters;              /* the reference uses 5 (decoders.py:521) */
    int32_t per_class_inits;
    int64_t steps;
    double  pz_tilde_sampling, alpha, pz_tilde;
    uint64_t seed;
    const double *u_nb, *u_py;  /* replay, per (syndrome, class): [S*n_eq][n_nb], [S*n_eq][n_py] */
    int64_t n_nb, n_py;
} qecmc_alpha_cfg;
int qecmc_stdc_alpha(qecmc_ctx *ctx, const qecmc_alpha_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                     int64_t *distinct, qecmc_stats *stats);

/* ------------------------------------------------------------------------------
 * General (x, y, z) noise: STDC_general_noise / STDC_general_noise_shortest / STDC_droplet_general_noise
 * (decoders.py:325-508) over Chain_xyz.update_chain_fast (_update_chain_fast_xyz, src/mcmc.py:106-114,162-173) when
 * the sampling rate is a triple, or over Chain.update_chain_fast when it is a scalar.  No rain (randomize is False in
 * every branch of the reference).  Weight of a distinct chain = exp(-sum_{i: n_i > 0} beta_i n_i),
 * beta_i = -log((p_i / 3) / (1 - p_i)).
 *   eqdistr          [S][n_eq] percent over all distinct chains
 *   eqdistr_shortest [S][n_eq] percent over the chains whose weighted length is np.isclose to the class minimum (optional)
 *   distinct         [S][n_eq] optional
 * ------------------------------------------------------------------------------ */
typedef struct qecmc_xyz_cfg {
    int32_t geom_code, geom_chain, L, droplets;
    int32_t iters;             /* 5 in the reference (decoders.py:336) */
    int32_t per_class_inits;
    int32_t use_xyz_sampling;  /* 1: Chain_xyz(p_sampling_xyz); 0: Chain(p_sampling) */
    int32_t reserved;
    int64_t steps;
    double  p_xyz[3];          /* error model (p_x, p_y, p_z) */
    double  p_sampling_xyz[3];
    double  p_sampling;
    uint64_t seed;
    const double *u_nb;        /* replay: [S*n_eq*droplets][steps*iters][k+1] numba-stream draws */
} qecmc_xyz_cfg;
/* Chain_xyz.update_chain_fast (src/mcmc.py:113-114): `iters` steps on `chains` lattices in place, accept iff
 * u < prod_i (p_i / (1 - sum p))^(dn_i).  cfg->p is ignored; p_xyz = the chain's (p_x, p_y, p_z).
 * u != NULL: replay, [chains][iters][k+1] numba-stream draws. */
int qecmc_chain_update_xyz(qecmc_ctx *ctx, const qecmc_chain_cfg *cfg, const double *p_xyz, const double *u, uint8_t *qm,
                           int64_t chains, int64_t iters, qecmc_stats *stats);
int qecmc_stdc_general_noise(qecmc_ctx *ctx, const qecmc_xyz_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                             double *eqdistr_shortest, int64_t *distinct, qecmc_stats *stats);

/* ---- The workload loop around the decoders: generate_data.py:53-261 ---------------------------------------------
 * A batch of syndromes is drawn, labelled, hidden, decoded and scored without leaving the GPU.
 *
 * generate_random_error.  toric_form = 1: Toric_code.generate_random_error(p_error) (toric_model.py:15-24): a qubit
 * errs iff u < p_error and then carries np.random.randint(3) + 1.  toric_form = 0: generate_random_error(p_x, p_y, p_z)
 * of Planar_code (planar_model.py:18-37; layer 1 loses its last row and column), RotSurCode
 * (rotated_surface_model.py:25-38) and xzzx_code (xzzx_model.py:16-29): r < p_z -> Z, p_z < r < p_z + p_x -> X,
 * p_z + p_x < r < p_z + p_x + p_y -> Y.
 *   u      replay: [S][n_sites] the reference's uniforms in lattice order (NULL: native Philox keyed by seed)
 *   pauli  replay, toric form only: [S][n_sites] the randint(3) + 1 draws
 *   qm     out [S][n_sites]; eq_true out [S] = define_equivalence_class() of each lattice (optional)
 * _dev variants: every pointer (including cfg->u / cfg->pauli) is a device pointer. */
typedef struct qecmc_noise_cfg {
    int32_t geom, L;
    int32_t toric_form, reserved;
    double  p_error;            /* toric form */
    double  p_x, p_y, p_z;      /* xyz form */
    uint64_t seed;
    const double  *u;
    const uint8_t *pauli;
} qecmc_noise_cfg;
int qecmc_generate_errors(qecmc_ctx *ctx, const qecmc_noise_cfg *cfg, int64_t S, uint8_t *qm, int32_t *eq_true);
int qecmc_generate_errors_dev(qecmc_ctx *ctx, const qecmc_noise_cfg *cfg, int64_t S, uint8_t *d_qm, int32_t *d_eq_true);

/* define_equivalence_class of S lattices (toric_model.py:317-351, planar_model.py:379-390,
 * rotated_surface_model.py:411-420, xzzx_model.py:455-486): cls[S] in [0, 16) (toric) or [0, 4). */
int qecmc_define_equivalence_class(qecmc_ctx *ctx, int32_t geom, int32_t L, const uint8_t *qm, int64_t S, int32_t *cls);
int qecmc_define_equivalence_class_dev(qecmc_ctx *ctx, int32_t geom, int32_t L, const uint8_t *d_qm, int64_t S, int32_t *d_cls);

/* apply_random_logical in place on S lattices (toric_model.py:228-253, planar_model.py:271-288,
 * rotated_surface_model.py:331-346, xzzx_model.py:340-357) -- generate_data.py:131 hides the true class this way.
 *   u    replay: [S][6] numba-stream uniforms in draw order (toric: op0, op1, then X_pos / Z_pos per layer as drawn;
 *        others: op, X_pos, Z_pos as drawn); NULL: native Philox keyed by seed
 *   ops  out, optional: [S][2] the operator drawn per layer (second entry 0 for the one-layer codes) */
int qecmc_apply_random_logical(qecmc_ctx *ctx, int32_t geom, int32_t L, uint8_t *qm, int64_t S, uint64_t seed, const double *u,
                               int32_t *ops);
int qecmc_apply_random_logical_dev(qecmc_ctx *ctx, int32_t geom, int32_t L, uint8_t *d_qm, int64_t S, uint64_t seed,
                                   const double *d_u, int32_t *d_ops);

/* Failure count of generate_data.py:137-201: choice = np.argmax(distr) (use_argmin = 1: np.argmin, the single_temp
 * rule of :199-201), failure iff choice != eq_true.  distr is [S][n_eq] float64 (QECMC_DISTR_F64) or uint8
 * (QECMC_DISTR_U8, the PT decoders).  choice [S] is optional; *failures is a host int64 in both variants. */
enum qecmc_distr_dtype { QECMC_DISTR_F64 = 0, QECMC_DISTR_U8 = 1 };
int qecmc_count_failures(qecmc_ctx *ctx, const void *distr, int32_t dtype, int32_t use_argmin, int32_t n_eq, int64_t S,
                         const int32_t *eq_true, int32_t *choice, int64_t *failures);
int qecmc_count_failures_dev(qecmc_ctx *ctx, const void *d_distr, int32_t dtype, int32_t use_argmin, int32_t n_eq, int64_t S,
                             const int32_t *d_eq_true, int32_t *d_choice, int64_t *failures);

/* ------------------------------------------------------------------------------
 * Class-sorted minimum-weight perfect matching start states for planar chains (host code, no device work):
 * MWPM(code).solve() and class_sorted_mwpm(code) of src/mwpm.py (:408-415, :417-438, :462-475; regular_mwpm :479-487
 * is define_equivalence_class of mode 0's output).  The reference hands its defect graph (generate_edges :66-133,
 * generate_edges_constrained :136-229) to the external blossom5 binary (:376-405); this solves the same graphs in
 * process (dense primal-dual blossom algorithm), one syndrome per host thread.
 *   qm                 [S][2][L][L] error chains; the syndrome is Planar_code.syndrom() of each (planar_model.py:134-153).
 *                      NULL: the defects are given instead
 *   vertex_defects     [S][L-1][L], plaquette_defects [S][L][L-1] (0 / non-zero), used when qm is NULL
 *   mode               0: solve() -> out [S][2][L][L], weights [S][2] = matching weight per layer
 *                      1: class_sorted_mwpm -> out [S][4][2][L][L] in class order, weights [S][2][2] = weight of
 *                         solve_layer(layer, parity)
 *                      | 2 (tests): solve the reference's graphs as written, one ancilla node per defect, instead of the
 *                         equivalent problem on half the nodes -- same weights, eight times the work
 *   weights            optional
 *   threads            host threads (0: all)
 * ------------------------------------------------------------------------------ */
int qecmc_mwpm_planar(int32_t L, int64_t S, const uint8_t *qm, const uint8_t *vertex_defects, const uint8_t *plaquette_defects,
                      int32_t mode, uint8_t *out, int32_t *weights, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* QECMC_H */

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
    int64_t table_slots;        /* open-addressing slots per (syndrome, class) set     */
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

#ifdef __cplusplus
}
#endif
#endif /* QECMC_H */

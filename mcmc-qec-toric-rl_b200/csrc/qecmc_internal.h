// qecmc_internal.h -- host-side plumbing shared by the translation units of libqecmc:
// error reporting, device buffers, the context, lattice packing.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <tuple>
#include <vector>

#include "../../include/qecmc.h"
#include "qecmc_kernels.cuh"

int qecmc_set_err(int code, const char *fmt, ...);
#define set_err qecmc_set_err
#define CUDA_OK(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return set_err(QECMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define QTRY(call)            \
    do {                      \
        int r_ = (call);      \
        if (r_ != 0) return r_; \
    } while (0)

// bumped whenever the library allocates or frees device memory: a cached "free memory" figure is good while it stands still
inline uint64_t &alloc_generation()
{
    static uint64_t gen = 0;
    return gen;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        alloc_generation()++;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();   // an allocation failure is not sticky: clear it so that the caller can retry smaller
            return set_err(e == cudaErrorMemoryAllocation ? QECMC_ERR_NOMEM : QECMC_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes,
                           cudaGetErrorString(e));
        }
        cap = bytes;
        return 0;
    }
    void release()
    {
        if (p) { cudaFree(p); alloc_generation()++; }
        p = nullptr;
        cap = 0;
    }
};

// a call-local device buffer: released when it goes out of scope, on every return path
struct ScopedDevBuf : DevBuf {
    ScopedDevBuf() = default;
    ScopedDevBuf(const ScopedDevBuf &) = delete;
    ScopedDevBuf &operator=(const ScopedDevBuf &) = delete;
    ~ScopedDevBuf() { release(); }
};

// device buffers of the tempering-ladder drivers (qecmc_ladder.cu); they live in the context so that repeated calls
// reuse them
struct LadderDev {
    DevBuf log_hash, short_v, short_n, short_u, pw, queue;
    DevBuf thr_d, thr_u, thr_top_d, diff, wtab, lat, lat_out, flags, neff, tops0, snap_lat, snap_flags, snap_tops0, hist, eqc,
        info, pct, status, u_nb, u_py, qm, bytes_out, Zd, dist;
    ~LadderDev()
    {
        for (DevBuf *b : {&log_hash, &short_v, &short_n, &short_u, &pw, &queue, &thr_d, &thr_u, &thr_top_d, &diff, &wtab, &lat, &lat_out, &flags, &neff, &tops0, &snap_lat,
                          &snap_flags, &snap_tops0, &hist, &eqc, &info, &pct, &status, &u_nb, &u_py, &qm, &bytes_out, &Zd, &dist})
            b->release();
    }
};

struct qecmc_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaDeviceProp prop;
    int64_t table_budget = 0;
    size_t free_cached = 0;                 // cudaMemGetInfo's answer at allocation generation free_cached_gen
    uint64_t free_cached_gen = ~(uint64_t)0;
    DevBuf packed, tables, Z, counters, qm_in, out_f64, out_u32, out_u64, out_i32, replay_a, replay_b, scratch, nhist, mhist,
        shorts, sums;
    std::map<std::tuple<int, int, int>, uint64_t *> stab_hash;  // (geom, L, wide) -> device table
    std::map<std::tuple<int, int, int>, uint2 *> stab_desc;     // (geom, L, wide) -> descriptor table
    DevBuf lut, log_hash, log_counts, dd_scratch;
    uint32_t lut_thr[QECMC_THR_N] = {0};   // thresholds the device LUT was built from
    bool lut_valid = false;
    LadderDev ld;
    cudaEvent_t ev[4];
    uint64_t hash_seed = 0x5EEDC0DE2020ull;
    int64_t launches = 0;
    // sizing of the most recent STDC-family call (qecmc_last_plan): syndromes one wave may hold within the table budget,
    // chains one round of CTAs over the SMs holds at the kernel's full CTA size
    int64_t plan_wave_cap = 0, plan_round_chains = 0;
    // qecmc_debug_set: test switches between code paths that must agree (never read from the environment)
    int dbg_force_wide = 0, dbg_insert_mode = -1, dbg_serial_sweep = 0;
    int dbg_pt_lt = 0;          // lanes per top-rung replica in the rung-major tempering kernel (0: default)
    int dbg_pt_grid = 0;        // cap on the rung-major kernel's grid (tests: makes ladders queue behind few CTAs)
    int dbg_packed = -1;        // packed-lattice chain kernel for 17 <= L <= 24: -1 default, 0 off, 1 / 2 / 4 table copies
    int dbg_ladder_kernel = 0;  // 1: native ladders on the warp-per-ladder (replay) kernel instead of the rung-major one
};

// rung-major tempering kernel (qecmc_pt.cu): CTA shape chosen for a ladder configuration
namespace qecmc { struct LadderParams; }
struct PtPlan {
    int NLC = 0, T = 0, lt = 0;      // ladders per CTA, threads per CTA, lanes per top-rung replica
    int small_cta = 0;               // the instantiation compiled for CTAs of at most 448 threads
    size_t smem = 0;
    int blocks_per_sm = 0, max_grid = 0;
};
int qecmc_pt_plan(qecmc_ctx *c, const qecmc::LadderParams &lp, PtPlan *out);
int qecmc_pt_launch(qecmc_ctx *c, const qecmc::LadderParams &lp, const PtPlan &pl, int grid, uint32_t step0, void *hist,
                    int64_t hist_stride);


using namespace qecmc;

// ------------------------------ helpers ------------------------------
static inline double numba_pow(double a, int64_t b)
{
    // numba lowers float64 ** int64 to square-and-multiply (reciprocal for negative exponents);
    // this is what _update_chain_fast (src/mcmc.py:158) evaluates.
    double r = 1.0;
    bool inv = b < 0;
    uint64_t e = inv ? (uint64_t)(-b) : (uint64_t)b;
    while (e) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    return inv ? 1.0 / r : r;
}

static inline void make_thr(double p, int pow_kind, Thr &t)
{
    double factor = (p / 3.0) / (1.0 - p);  // src/mcmc.py:16
    for (int i = 0; i < QECMC_THR_N; i++) {
        int dE = i - QECMC_THR_OFF;
        double v = pow_kind == QECMC_POW_NUMBA ? numba_pow(factor, dE) : pow(factor, (double)dE);
        t.d[i] = v;
        if (!(v < 1.0)) t.u32[i] = 0xFFFFFFFFu;  // u < v always holds for u in [0,1)
        else {
            double x = ceil(v * 4294967296.0);  // u32/2^32 < v  <=>  u32 <= ceil(v*2^32) - 1
            t.u32[i] = x < 1.0 ? 0u : (uint32_t)(x - 1.0);
        }
    }
}

static inline int check_geom(int geom, int L)
{
    if (geom < 0 || geom > 3) return set_err(QECMC_ERR_ARG, "unknown geometry %d", geom);
    if (L < 2 || L > 32) return set_err(QECMC_ERR_ARG, "system size L=%d outside [2, 32]", L);
    if ((geom == ROTATED || geom == XZZX) && (L < 3 || (L % 2) == 0))
        return set_err(QECMC_ERR_ARG, "rotated/XZZX codes need odd L >= 3 (got %d)", L);
    return 0;
}

template <typename W> static int pack_lattices(qecmc_ctx *c, const uint8_t *d_qm, int64_t n_lat, const Geo &g, void *d_out)
{
    int64_t n_words = n_lat * g.nw;
    QTRY(c->scratch.ensure(sizeof(int)));
    CUDA_OK(cudaMemsetAsync(c->scratch.p, 0, sizeof(int), c->stream));
    int T = 256;
    pack_kernel<W><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>(d_qm, (W *)d_out, n_words, g.L, (int *)c->scratch.p);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    int bad = 0;
    CUDA_OK(cudaMemcpyAsync(&bad, c->scratch.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (bad) return set_err(QECMC_ERR_ARG, "qubit_matrix holds values outside 0..3");
    return 0;
}

// Two-layer codes (toric, planar): table-driven stabilizer descriptors, shared by the STDC chain kernel and the ladder kernel.
// descriptor words:  x = sh | w0<<8 | w1<<16 | w2<<24     y = sh2 | f_and<<8 | f_or<<16 | (v==3)<<24
// slot layout a=(w0,sh) b=(w0,sh2) c=(w1,sh) d=(w2,sh); verified against decode<>() mask by mask
template <int GEOM, typename W> static int build_stab_desc(qecmc_ctx *c, const Geo &g, const uint2 **out)
{
    auto key = std::make_tuple(GEOM, g.L, (int)(sizeof(W) == 8));
    auto it = c->stab_desc.find(key);
    if (it != c->stab_desc.end()) { *out = it->second; return 0; }
    const int L = g.L;
    std::vector<uint2> tab(g.nstab);
    for (int idx = 0; idx < g.nstab; idx++) {
        int row, col, op;
        idx_to_rco<GEOM>(g, idx, row, col, op);
        int w0, w1, w2, sh = 2 * col, sh2;
        bool pa = true, pb = true, pc = true, pd = true;
        if (GEOM == TORIC) {
            if (op == 1) { w0 = L + row; sh2 = 2 * (col == 0 ? L - 1 : col - 1); w1 = row; w2 = row == 0 ? L - 1 : row - 1; }
            else { w0 = row; sh2 = 2 * (col == L - 1 ? 0 : col + 1); w1 = L + row; w2 = L + (row == L - 1 ? 0 : row + 1); }
        } else {
            if (op == 1) { w0 = L + row; pa = col < L - 1; pb = col > 0; sh2 = pb ? 2 * (col - 1) : sh; w1 = row; w2 = row + 1; }
            else { w0 = row; sh2 = 2 * (col + 1); w1 = L + row; pc = row < L - 1; pd = row > 0; w2 = pd ? L + row - 1 : L + 1; }
        }
        uint32_t f_and = 0xFF, f_or = 0;
        bool pres[4] = {pa, pb, pc, pd};
        for (int i = 0; i < 4; i++)
            if (!pres[i]) { f_and &= ~(1u << (2 * i)); f_or |= 2u << (2 * i); }
        tab[idx].x = (uint32_t)sh | ((uint32_t)w0 << 8) | ((uint32_t)w1 << 16) | ((uint32_t)w2 << 24);
        tab[idx].y = (uint32_t)sh2 | (f_and << 8) | (f_or << 16) | ((op == 3 ? 1u : 0u) << 24);
        // self-check against the generic geometry
        Upd<W> u;
        decode<GEOM, W>(g, row, col, op, u);
        std::map<int, W> want, got;
        for (int i = 0; i < 3; i++) if (u.m[i]) want[u.w[i]] ^= u.m[i];
        W v = (W)op;
        if (pa) got[w0] ^= (W)(v << sh);
        if (pb) got[w0] ^= (W)(v << sh2);
        if (pc) got[w1] ^= (W)(v << sh);
        if (pd) got[w2] ^= (W)(v << sh);
        if (want != got || w0 == w1 || w0 == w2 || w1 == w2 || w0 >= g.nw || w1 >= g.nw || w2 >= g.nw)
            return set_err(QECMC_ERR_UNSUPPORTED, "internal: stabilizer descriptor %d disagrees with the geometry", idx);
    }
    uint2 *d = nullptr;
    alloc_generation()++;
    CUDA_OK(cudaMalloc(&d, sizeof(uint2) * g.nstab));
    CUDA_OK(cudaMemcpyAsync(d, tab.data(), sizeof(uint2) * g.nstab, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->stab_desc[key] = d;
    *out = d;
    return 0;
}

template <int GEOM, typename W> static int build_stab_hash(qecmc_ctx *c, const Geo &g, uint64_t **out)
{
    auto key = std::make_tuple(GEOM, g.L, (int)(sizeof(W) == 8));
    auto it = c->stab_hash.find(key);
    if (it != c->stab_hash.end()) { *out = it->second; return 0; }
    uint64_t *d = nullptr;
    alloc_generation()++;
    CUDA_OK(cudaMalloc(&d, sizeof(uint64_t) * g.nstab));
    stab_hash_kernel<GEOM, W><<<(g.nstab + 127) / 128, 128, 0, c->stream>>>(g, c->hash_seed, d);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    c->stab_hash[key] = d;
    *out = d;
    return 0;
}

// Free device memory as of the library's last allocation or release: cudaMemGetInfo costs milliseconds (~10 ms on a context
// holding tens of GB), so it is asked again only when the allocation generation has moved.  The figure is a HINT: memory
// taken by anyone else on the device (torch tensors between two calls, another context or process) does not move the
// generation.  The wave-sizing drivers therefore run under with_fresh_memory_retry(): an allocation that fails with
// QECMC_ERR_NOMEM drops the cached figure and the driver runs once more, sized from a fresh query.
static inline void forget_free_device_bytes(qecmc_ctx *c) { c->free_cached_gen = ~(uint64_t)0; alloc_generation()++; }

template <typename F> static inline int with_fresh_memory_retry(qecmc_ctx *c, F run)
{
    int rc = run();
    if (rc != QECMC_ERR_NOMEM) return rc;
    forget_free_device_bytes(c);
    return run();
}

static inline int free_device_bytes(qecmc_ctx *c, size_t *fr)
{
    if (c->free_cached_gen != alloc_generation()) {
        size_t f = 0, tot = 0;
        CUDA_OK(cudaMemGetInfo(&f, &tot));
        c->free_cached = f;
        c->free_cached_gen = alloc_generation();
    }
    *fr = c->free_cached;
    return 0;
}

static inline int pick_threads(size_t bytes_per_chain, size_t fixed, const cudaDeviceProp &prop, int *threads, int *blocks_per_sm,
                               bool allow1024 = false, int regs_per_thread = 0)
{
    // largest resident chain count per SM within the shared-memory budget, 256-thread CTAs preferred
    size_t budget = prop.sharedMemPerMultiprocessor;
    int best_T = 0, best_res = 0;
    // candidates in order of preference: a later one wins only with strictly more resident chains
    for (int T : {1024, 256, 128, 64}) {
        if (T == 1024 && !allow1024) continue;
        size_t per_block = bytes_per_chain * T + fixed + 1024;  // +1 KiB reserved per CTA
        if (bytes_per_chain * T + fixed > prop.sharedMemPerBlockOptin) continue;
        int nb = (int)(budget / per_block);
        int max_thr = prop.maxThreadsPerMultiProcessor;
        if (nb * T > max_thr) nb = max_thr / T;
        if (regs_per_thread > 0 && nb * T * regs_per_thread > prop.regsPerMultiprocessor) nb = prop.regsPerMultiprocessor / (T * regs_per_thread);
        if (nb * T > best_res) { best_res = nb * T; best_T = T; *blocks_per_sm = nb; }
    }
    if (!best_T) return set_err(QECMC_ERR_UNSUPPORTED, "lattice does not fit in shared memory");
    *threads = best_T;
    return 0;
}


static inline uint64_t next_pow2(uint64_t x)
{
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

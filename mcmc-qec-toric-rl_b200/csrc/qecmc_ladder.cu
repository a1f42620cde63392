// qecmc_ladder.cu -- C ABI of the tempering-ladder decoders (include/qecmc.h): host-side ladder
// tables (numpy.linspace, libm pow), launches of ladder_kernel, result staging.
#include "qecmc_internal.h"
#include "qecmc_ladder.cuh"

using namespace qecmc;

namespace {

template <typename W> struct HostLat {
    W *d;
    __host__ __device__ W get(int w) const { return d[w]; }
    __host__ __device__ void set(int w, W v) { d[w] = v; }
};

// raw class of the logical operator (layer, op) acting on an empty lattice
template <int GEOM> void class_deltas(const Geo &g, int out[8])
{
    for (int l = 0; l < 2; l++)
        for (int op = 0; op < 4; op++) {
            uint64_t buf[64] = {0};
            HostLat<uint64_t> a{buf};
            out[l * 4 + op] = 0;
            if (l == 1 && GEOM != TORIC) continue;
            lat_apply_logical<GEOM, uint64_t>(g, a, op, l, 0, 0);
            out[l * 4 + op] = class_raw_to_label(GEOM, lat_class<GEOM, uint64_t>(g, a));
        }
}

// fingerprints of the logical strings: [layer][X, Z][position] (the row-word layout does not depend on the word type)
template <int GEOM> void logical_hashes(const Geo &g, uint64_t seed, std::vector<uint64_t> &out)
{
    out.assign(2 * 2 * 32, 0);
    const int op_x = 1, op_z = (GEOM == TORIC || GEOM == XZZX) ? 3 : 2;  // the operator codes that act with X only / Z only
    for (int l = 0; l < (GEOM == TORIC ? 2 : 1); l++)
        for (int k = 0; k < 2; k++)
            for (int pos = 0; pos < g.L; pos++) {
                uint64_t buf[64] = {0};
                HostLat<uint64_t> a{buf};
                lat_apply_logical<GEOM, uint64_t>(g, a, k == 0 ? op_x : op_z, l, pos, pos);
                out[(size_t)(l * 2 + k) * 32 + pos] = lat_hash<uint64_t>(g, a, seed);
            }
}

// numpy.linspace(start, stop, num): arange(num) * step + start with the last point forced to stop
std::vector<double> np_linspace(double start, double stop, int num)
{
    std::vector<double> y((size_t)num);
    if (num == 1) { y[0] = start; return y; }
    volatile double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; i++) {
        volatile double t = step == 0 ? ((double)i / (double)(num - 1)) * (stop - start) : (double)i * step;
        y[(size_t)i] = t + start;
    }
    y[(size_t)num - 1] = stop;
    return y;
}

uint32_t thr_to_u32(double v)
{
    if (!(v < 1.0)) return 0xFFFFFFFFu;
    double x = ceil(v * 4294967296.0);
    return x < 1.0 ? 0u : (uint32_t)(x - 1.0);
}

struct LadderTables {
    std::vector<double> ladder, diff, thr_d, thr_top_d, wtab;
    std::vector<uint32_t> thr_u;
    int top_accept_all = 0;
};

// Ladder.__init__ (mcmc.py:49-79), Ladder_alpha.__init__ (mcmc_alpha.py:77-106), Ladder_biased.__init__
// (mcmc_biased.py:66-98) and the per-chain constants of the three update_chain variants.
void make_ladder_tables(const qecmc_ladder_cfg *cfg, const Geo &g, LadderTables &t)
{
    const int Nc = cfg->Nc, L = cfg->L, ns1 = g.nsites + 1;
    double top = cfg->kind == LK_DEPOL ? 0.75 : cfg->kind == LK_ALPHA ? 1.0 : (cfg->param_b + 1) / (2 * cfg->param_b + 1);
    t.ladder = np_linspace(cfg->bottom, top, Nc);
    t.diff.assign(Nc > 1 ? (size_t)Nc - 1 : 1, 0.0);
    for (int i = 0; i + 1 < Nc; i++) {
        double lo = t.ladder[i], hi = t.ladder[i + 1];
        t.diff[i] = cfg->kind == LK_ALPHA ? lo / hi : (lo * (1 - hi)) / (hi * (1 - lo));
    }
    t.thr_d.assign((size_t)Nc * 9, 0.0);
    t.thr_u.assign((size_t)Nc * 9, 0u);
    t.thr_top_d.assign((size_t)8 * L + 1, 0.0);
    if (cfg->kind == LK_DEPOL) {
        for (int r = 0; r < Nc; r++) {
            double p = t.ladder[r], factor = (p / 3.0) / (1.0 - p);  // mcmc.py:16
            for (int d = -4; d <= 4; d++) {
                double v = pow(factor, (double)d);  // CPython float ** int
                t.thr_d[(size_t)r * 9 + d + 4] = v;
                t.thr_u[(size_t)r * 9 + d + 4] = thr_to_u32(v);
            }
        }
        double p = t.ladder[Nc - 1], factor = (p / 3.0) / (1.0 - p);
        for (int d = -4 * L; d <= 4 * L; d++) t.thr_top_d[(size_t)(d + 4 * L)] = pow(factor, (double)d);
        t.top_accept_all = p >= 0.75;
    } else {
        t.wtab.assign((size_t)Nc * 4 * ns1, 0.0);
        const double num = (double)L * (double)L;  // system_size ** 2, also for the two-layer codes
        for (int r = 0; r < Nc; r++) {
            double px, py, pz;
            if (cfg->kind == LK_ALPHA) {  // mcmc_alpha.py:30-35
                double pz_tilde = t.ladder[r], alpha = cfg->param_b;
                double p_tilde = pz_tilde + 2 * pow(pz_tilde, alpha);
                double p = p_tilde / (1 + p_tilde);
                pz = pz_tilde * (1 - p);
                px = py = pow(pz_tilde, alpha) * (1 - p);
            } else {                      // mcmc_biased.py:25-28
                double p = t.ladder[r], eta = cfg->param_b;
                pz = p * eta / (eta + 1);
                px = p / (2 * (eta + 1));
                py = px;
            }
            double q0 = 1 - px - py - pz;
            double *w = &t.wtab[(size_t)r * 4 * ns1];
            for (int k = 0; k < ns1; k++) {
                w[k] = pow(px, (double)k);
                w[ns1 + k] = pow(py, (double)k);
                w[2 * ns1 + k] = pow(pz, (double)k);
                w[3 * ns1 + k] = pow(q0, num - (double)k);
            }
        }
    }
}


template <typename T> int upload(qecmc_ctx *c, DevBuf &b, const std::vector<T> &v)
{
    size_t n = v.size() * sizeof(T);
    QTRY(b.ensure(n ? n : 8));
    if (n) CUDA_OK(cudaMemcpyAsync(b.p, v.data(), n, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int check_ladder_cfg(const qecmc_ladder_cfg *cfg)
{
    if (!cfg) return set_err(QECMC_ERR_ARG, "cfg is NULL");
    QTRY(check_geom(cfg->geom, cfg->L));
    if (cfg->kind < 0 || cfg->kind > 2) return set_err(QECMC_ERR_ARG, "ladder kind %d not in {0 depolarizing, 1 alpha, 2 biased}", cfg->kind);
    if (cfg->Nc < 1) return set_err(QECMC_ERR_ARG, "Nc must be >= 1");
    if (cfg->Nc > 32) return set_err(QECMC_ERR_UNSUPPORTED, "Nc = %d: a ladder lives in one warp, at most 32 rungs", cfg->Nc);
    if (cfg->iters < 1) return set_err(QECMC_ERR_ARG, "iters must be >= 1");
    if (!(cfg->bottom > 0.0 && cfg->bottom < 1.0)) return set_err(QECMC_ERR_ARG, "bottom rate %g outside (0,1)", cfg->bottom);
    if (cfg->kind == LK_ALPHA && !(cfg->param_b > 0)) return set_err(QECMC_ERR_ARG, "alpha must be > 0");
    if (cfg->kind == LK_BIASED && !(cfg->param_b > 0)) return set_err(QECMC_ERR_ARG, "eta must be > 0");
    if (!(cfg->p_logical >= 0.0 && cfg->p_logical <= 1.0)) return set_err(QECMC_ERR_ARG, "p_logical outside [0,1]");
    if ((cfg->u_nb == nullptr) != (cfg->u_py == nullptr)) return set_err(QECMC_ERR_ARG, "replay needs both u_nb and u_py");
    if (cfg->u_nb && (cfg->n_nb <= 0 || cfg->n_py <= 0 || cfg->n_nb >= (1ll << 31) || cfg->n_py >= (1ll << 31)))
        return set_err(QECMC_ERR_ARG, "replay stream lengths must be in (0, 2^31)");
    return 0;
}

template <int GEOM, typename W> int launch_ladder_gw(qecmc_ctx *c, LadderParams &p, bool replay, bool weighted)
{
    const int T = 128;
    size_t smem = (((size_t)p.g.nw * T * sizeof(W) + 15) & ~(size_t)15) + (size_t)p.Nc * 9 * 12 + 16;
    if (GEOM == ROTATED || GEOM == XZZX) smem += (size_t)p.g.nstab * 8 + 16 * 256 * 2;   // step tables of the one-layer codes
    if ((GEOM == TORIC || GEOM == PLANAR) && !weighted) {                                  // ... and of the two-layer codes
        smem += (size_t)p.g.nstab * 8 + 512;
        QTRY((build_stab_desc<GEOM, W>(c, p.g, &p.desc2)));
    }
    smem += 8 + (size_t)(p.Nc > 1 ? p.Nc - 1 : 0) * (2 * QECMC_PW_K + 1) * 8;            // swap-sweep power table
    if (smem > c->prop.sharedMemPerBlockOptin) return set_err(QECMC_ERR_UNSUPPORTED, "lattice does not fit in shared memory");
    unsigned grid = (unsigned)((p.n_ladders * p.G + T - 1) / T);
#define QECMC_LAUNCH(R, WT)                                                                                               \
    do {                                                                                                                  \
        CUDA_OK(cudaFuncSetAttribute(ladder_kernel<GEOM, W, R, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ladder_kernel<GEOM, W, R, WT><<<grid, T, smem, c->stream>>>(p);                                                   \
    } while (0)
    if (replay) { if (weighted) QECMC_LAUNCH(true, true); else QECMC_LAUNCH(true, false); }
    else { if (weighted) QECMC_LAUNCH(false, true); else QECMC_LAUNCH(false, false); }
#undef QECMC_LAUNCH
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <int GEOM> int launch_ladder_g(qecmc_ctx *c, LadderParams &p, bool replay, bool weighted)
{
    class_deltas<GEOM>(p.g, p.cls_delta);
    if (p.acct >= ACCT_DC || p.track_shortest) {
        std::vector<uint64_t> lh;
        logical_hashes<GEOM>(p.g, c->hash_seed, lh);
        QTRY(c->log_hash.ensure(lh.size() * sizeof(uint64_t)));
        CUDA_OK(cudaMemcpyAsync(c->log_hash.p, lh.data(), lh.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        p.log_hash = (const uint64_t *)c->log_hash.p;
    }
    if (p.acct >= ACCT_DC || p.track_shortest) {
        if (p.g.L > 16) QTRY((build_stab_hash<GEOM, uint64_t>(c, p.g, (uint64_t **)&p.stab_hash)));
        else QTRY((build_stab_hash<GEOM, uint32_t>(c, p.g, (uint64_t **)&p.stab_hash)));
    }
    return p.g.L > 16 ? launch_ladder_gw<GEOM, uint64_t>(c, p, replay, weighted) : launch_ladder_gw<GEOM, uint32_t>(c, p, replay, weighted);
}

int launch_ladder(qecmc_ctx *c, LadderParams &p, bool replay)
{
    bool weighted = p.kind != LK_DEPOL;
    switch (p.g.geom) {
    case TORIC: return launch_ladder_g<TORIC>(c, p, replay, weighted);
    case PLANAR: return launch_ladder_g<PLANAR>(c, p, replay, weighted);
    case ROTATED: return launch_ladder_g<ROTATED>(c, p, replay, weighted);
    default: return launch_ladder_g<XZZX>(c, p, replay, weighted);
    }
}

// what launch_ladder_g / launch_ladder_gw prepare per geometry, for the rung-major kernel: class changes of the logical
// operators and (two-layer depolarizing ladders) the stabilizer descriptors
int prepare_pt(qecmc_ctx *c, LadderParams &p)
{
    const bool wide = p.g.L > 16, weighted = p.kind != LK_DEPOL;
    switch (p.g.geom) {
    case TORIC:
        class_deltas<TORIC>(p.g, p.cls_delta);
        if (!weighted) { if (wide) QTRY((build_stab_desc<TORIC, uint64_t>(c, p.g, &p.desc2))); else QTRY((build_stab_desc<TORIC, uint32_t>(c, p.g, &p.desc2))); }
        break;
    case PLANAR:
        class_deltas<PLANAR>(p.g, p.cls_delta);
        if (!weighted) { if (wide) QTRY((build_stab_desc<PLANAR, uint64_t>(c, p.g, &p.desc2))); else QTRY((build_stab_desc<PLANAR, uint32_t>(c, p.g, &p.desc2))); }
        break;
    case ROTATED: class_deltas<ROTATED>(p.g, p.cls_delta); break;
    default: class_deltas<XZZX>(p.g, p.cls_delta); break;
    }
    return 0;
}

// fills the configuration-only part of LadderParams and uploads the tables
int setup_ladder(qecmc_ctx *c, const qecmc_ladder_cfg *cfg, const Geo &g, LadderDev &d, LadderParams &p)
{
    memset(&p, 0, sizeof(p));
    LadderTables t;
    make_ladder_tables(cfg, g, t);
    QTRY(upload(c, d.thr_d, t.thr_d));
    QTRY(upload(c, d.thr_u, t.thr_u));
    QTRY(upload(c, d.thr_top_d, t.thr_top_d));
    QTRY(upload(c, d.diff, t.diff));
    {   // swap-sweep power table of the rung-major kernel: diff^k by numba's square-and-multiply (mcmc.py:149), k = 0 .. 2K
        std::vector<double> pw((size_t)(cfg->Nc > 1 ? cfg->Nc - 1 : 1) * (2 * QECMC_PW_K + 1), 0.0);
        for (int i = 0; i + 1 < cfg->Nc; i++)
            for (int k = 0; k <= 2 * QECMC_PW_K; k++) pw[(size_t)i * (2 * QECMC_PW_K + 1) + k] = numba_pow(t.diff[i], k);
        QTRY(upload(c, d.pw, pw));
    }
    QTRY(upload(c, d.wtab, t.wtab));
    QTRY(d.status.ensure(sizeof(int)));
    CUDA_OK(cudaMemsetAsync(d.status.p, 0, sizeof(int), c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
    p.g = g;
    p.kind = cfg->kind;
    p.Nc = cfg->Nc;
    p.G = (int)next_pow2((uint64_t)cfg->Nc);
    if (p.G < 4) p.G = 4;   // a ladder's lanes fill the top rung's random-word pool: 4 calls = 16 words cover one iteration
    p.iters = cfg->iters;
    p.p_logical = cfg->p_logical;
    p.top_accept_all = t.top_accept_all;
    p.serial_sweep = c->dbg_serial_sweep;   // tests (qecmc_debug_set "serial_sweep")
    p.thr_d = (const double *)d.thr_d.p;
    p.thr_u = (const uint32_t *)d.thr_u.p;
    p.thr_top_d = (const double *)d.thr_top_d.p;
    p.diff = (const double *)d.diff.p;
    p.pw = (const double *)d.pw.p;
    p.wtab = (const double *)d.wtab.p;
    p.alpha = cfg->param_b;
    p.seed = cfg->seed;
    {
        uint32_t k0 = (uint32_t)cfg->seed, k1 = (uint32_t)(cfg->seed >> 32);
        for (int r = 0; r < 10; r++) { p.keys.k0[r] = k0; p.keys.k1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    }
    p.status = (int *)d.status.p;
    p.hash_seed = c->hash_seed;
    return 0;
}

int stage_replay(qecmc_ctx *c, const qecmc_ladder_cfg *cfg, int64_t n_ladders, LadderDev &d, LadderParams &p)
{
    if (!cfg->u_nb) return 0;
    size_t a = (size_t)n_ladders * cfg->n_nb * sizeof(double), b = (size_t)n_ladders * cfg->n_py * sizeof(double);
    QTRY(d.u_nb.ensure(a));
    QTRY(d.u_py.ensure(b));
    CUDA_OK(cudaMemcpyAsync(d.u_nb.p, cfg->u_nb, a, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(d.u_py.p, cfg->u_py, b, cudaMemcpyHostToDevice, c->stream));
    p.u_nb = (const double *)d.u_nb.p;
    p.u_py = (const double *)d.u_py.p;
    p.n_nb = (int)cfg->n_nb;
    p.n_py = (int)cfg->n_py;
    return 0;
}

int check_status(qecmc_ctx *c, LadderDev &d)
{
    int st = 0;
    CUDA_OK(cudaMemcpyAsync(&st, d.status.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (st) return set_err(QECMC_ERR_ARG, "replay stream %s ran dry: the run needs more uniforms than were supplied", st == 1 ? "u_nb" : "u_py");
    return 0;
}

template <typename W> void unpack_async(qecmc_ctx *c, const void *packed, uint8_t *bytes, int64_t n_words, int L)
{
    const int T = 256;
    unpack_kernel<W><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>((const W *)packed, bytes, n_words, L);
    c->launches++;
}

template <int GEOM, typename W> __global__ void to_class_kernel(Geo g, const W *__restrict__ in, W *__restrict__ out, int64_t S)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * g.neq) return;
    const int eq = (int)(i % g.neq);
    W *o = out + i * g.nw;
    const W *src = in + (i / g.neq) * g.nw;
    for (int w = 0; w < g.nw; w++) o[w] = src[w];
    struct A { W *d; __device__ W get(int w) const { return d[w]; } __device__ void set(int w, W v) { d[w] = v; } } a{o};
    lat_to_class<GEOM, W>(g, a, eq);
}

template <typename W> int to_class_all(qecmc_ctx *c, const Geo &g, const void *in, void *out, int64_t S)
{
    unsigned grid = (unsigned)((S * g.neq + 127) / 128);
    switch (g.geom) {
    case TORIC: to_class_kernel<TORIC, W><<<grid, 128, 0, c->stream>>>(g, (const W *)in, (W *)out, S); break;
    case PLANAR: to_class_kernel<PLANAR, W><<<grid, 128, 0, c->stream>>>(g, (const W *)in, (W *)out, S); break;
    case ROTATED: to_class_kernel<ROTATED, W><<<grid, 128, 0, c->stream>>>(g, (const W *)in, (W *)out, S); break;
    default: to_class_kernel<XZZX, W><<<grid, 128, 0, c->stream>>>(g, (const W *)in, (W *)out, S); break;
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

void fill_stats(qecmc_ctx *c, qecmc_stats *stats, int64_t steps_total, const unsigned long long *cnt, float ms, int64_t waves)
{
    if (!stats) return;
    memset(stats, 0, sizeof(*stats));
    stats->metropolis_steps = steps_total;
    stats->accepted = (int64_t)cnt[0];
    stats->samples = (int64_t)cnt[1];
    stats->distinct = (int64_t)cnt[3];
    stats->waves = waves;
    stats->kernel_launches = c->launches;
    stats->chain_kernel_ms = ms;
    stats->total_ms = ms;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int qecmc_ladder_run(qecmc_ctx *c, const qecmc_ladder_cfg *cfg, const qecmc_ladder_io *io, int64_t S, int64_t steps,
                                qecmc_stats *stats)
{
    if (!c || !io) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_ladder_cfg(cfg));
    if (S <= 0 || steps < 0) return set_err(QECMC_ERR_ARG, "S must be > 0 and steps >= 0");
    if ((io->snap_states != nullptr) != (io->snap_flags != nullptr) || (io->snap_states != nullptr) != (io->snap_tops0 != nullptr))
        return set_err(QECMC_ERR_ARG, "snapshots need snap_states, snap_flags and snap_tops0 together");
    if (io->resume && (!io->rung_states || !io->flags || !io->tops0))
        return set_err(QECMC_ERR_ARG, "resume needs rung_states, flags and tops0 (in/out)");
    if (!io->resume && !io->qm0) return set_err(QECMC_ERR_ARG, "qm0 is NULL");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    const Geo g = make_geo(cfg->geom, cfg->L);
    const bool wide = cfg->L > 16;
    const size_t wb = wide ? 8 : 4;
    const int Nc = cfg->Nc;
    const int64_t n_init = io->resume ? S * Nc : S;
    LadderDev &d = c->ld;   // device buffers persist in the context: no cudaMalloc / cudaFree per call
    LadderParams p;
    QTRY(setup_ladder(c, cfg, g, d, p));
    QTRY(stage_replay(c, cfg, S, d, p));
    QTRY(d.qm.ensure((size_t)n_init * g.nsites));
    QTRY(d.lat.ensure((size_t)n_init * g.nw * wb));
    QTRY(d.lat_out.ensure((size_t)S * Nc * g.nw * wb));
    QTRY(d.flags.ensure((size_t)S * Nc * sizeof(int)));
    QTRY(d.neff.ensure((size_t)S * Nc * sizeof(int2)));
    QTRY(d.tops0.ensure((size_t)S * sizeof(long long)));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaMemcpyAsync(d.qm.p, io->resume ? io->rung_states : io->qm0, (size_t)n_init * g.nsites, cudaMemcpyHostToDevice, c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, (const uint8_t *)d.qm.p, n_init, g, d.lat.p));
    else QTRY(pack_lattices<uint32_t>(c, (const uint8_t *)d.qm.p, n_init, g, d.lat.p));
    p.acct = ACCT_NONE;
    p.n_ladders = S;
    p.steps = steps;
    p.lat_in = d.lat.p;
    p.init_broadcast = io->resume ? 0 : 1;
    ScopedDevBuf flags_in, neff_in, tops0_in;   // released on every return path
    if (io->resume) {
        QTRY(flags_in.ensure((size_t)S * Nc * sizeof(int)));
        QTRY(tops0_in.ensure((size_t)S * sizeof(long long)));
        CUDA_OK(cudaMemcpyAsync(flags_in.p, io->flags, (size_t)S * Nc * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(tops0_in.p, io->tops0, (size_t)S * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        p.flags_in = (const int *)flags_in.p;
        p.tops0_in = (const long long *)tops0_in.p;
        if (io->n_eff_parts) {
            QTRY(neff_in.ensure((size_t)S * Nc * sizeof(int2)));
            CUDA_OK(cudaMemcpyAsync(neff_in.p, io->n_eff_parts, (size_t)S * Nc * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
            p.neff_in = (const int2 *)neff_in.p;
        }
    }
    p.lat_out = d.lat_out.p;
    p.flags_out = (int *)d.flags.p;
    p.neff_out = (int2 *)d.neff.p;
    p.tops0_out = (long long *)d.tops0.p;
    p.counters = (unsigned long long *)c->counters.p;
    if (io->snap_states) {
        QTRY(d.snap_lat.ensure((size_t)S * steps * Nc * g.nw * wb + 8));
        QTRY(d.snap_flags.ensure((size_t)S * steps * Nc * sizeof(int) + 8));
        QTRY(d.snap_tops0.ensure((size_t)S * steps * sizeof(long long) + 8));
        p.snap_lat = d.snap_lat.p;
        p.snap_flags = (int *)d.snap_flags.p;
        p.snap_tops0 = (long long *)d.snap_tops0.p;
    }
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    // native draws without per-step snapshots run on the rung-major kernel (qecmc_pt.cuh); replay keeps the reference's
    // draw order on the warp-per-ladder kernel
    const bool use_pt = !cfg->u_nb && !io->snap_states && steps > 0 && !c->dbg_ladder_kernel;
    int rc;
    if (use_pt) {
        PtPlan pl;
        rc = prepare_pt(c, p);
        if (rc == 0) rc = qecmc_pt_plan(c, p, &pl);
        if (rc == 0) {
            const int64_t want = (S + pl.NLC - 1) / pl.NLC;
            rc = qecmc_pt_launch(c, p, pl, (int)(want < pl.max_grid ? want : pl.max_grid), 0u, nullptr, 0);
        }
    } else {
        rc = launch_ladder(c, p, cfg->u_nb != nullptr);
    }
    if (rc == 0) rc = cudaEventRecord(c->ev[1], c->stream) == cudaSuccess ? 0 : set_err(QECMC_ERR_CUDA, "cudaEventRecord failed");
    if (rc == 0) rc = check_status(c, d);
    cudaStreamSynchronize(c->stream);
    if (rc) return rc;
    // results
    size_t out_bytes = (size_t)S * Nc * g.nsites;
    size_t snap_bytes = io->snap_states ? (size_t)S * steps * Nc * g.nsites : 0;
    QTRY(d.bytes_out.ensure((out_bytes > snap_bytes ? out_bytes : snap_bytes) + 8));
    if (io->rung_states) {
        if (wide) unpack_async<uint64_t>(c, d.lat_out.p, (uint8_t *)d.bytes_out.p, S * Nc * g.nw, g.L);
        else unpack_async<uint32_t>(c, d.lat_out.p, (uint8_t *)d.bytes_out.p, S * Nc * g.nw, g.L);
        CUDA_OK(cudaMemcpyAsync(io->rung_states, d.bytes_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    if (io->snap_states && steps > 0) {
        if (wide) unpack_async<uint64_t>(c, d.snap_lat.p, (uint8_t *)d.bytes_out.p, S * steps * Nc * g.nw, g.L);
        else unpack_async<uint32_t>(c, d.snap_lat.p, (uint8_t *)d.bytes_out.p, S * steps * Nc * g.nw, g.L);
        CUDA_OK(cudaMemcpyAsync(io->snap_states, d.bytes_out.p, snap_bytes, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(io->snap_flags, d.snap_flags.p, (size_t)S * steps * Nc * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(io->snap_tops0, d.snap_tops0.p, (size_t)S * steps * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    }
    if (io->flags) CUDA_OK(cudaMemcpyAsync(io->flags, d.flags.p, (size_t)S * Nc * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (io->tops0) CUDA_OK(cudaMemcpyAsync(io->tops0, d.tops0.p, (size_t)S * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    std::vector<int2> ne;
    if (io->n_eff || io->n_eff_parts) {
        ne.resize((size_t)S * Nc);
        CUDA_OK(cudaMemcpyAsync(ne.data(), d.neff.p, ne.size() * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    }
    unsigned long long cnt[8] = {0};
    CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < ne.size(); i++) {
        if (io->n_eff) {
            volatile double t = cfg->param_b * (double)ne[i].y;  // zb + alpha * (xb + yb), mcmc_alpha.py:22
            io->n_eff[i] = (double)ne[i].x + t;
        }
        if (io->n_eff_parts) { io->n_eff_parts[2 * i] = ne[i].x; io->n_eff_parts[2 * i + 1] = ne[i].y; }
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    fill_stats(c, stats, S * Nc * steps * cfg->iters, cnt, ms, 1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// PTEQ / PTEQ_biased / PTEQ_alpha (decoders.py:25-89, decoders_biasednoise.py:28-75,175-222)
static int pteq_once(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *qm, bool qm_on_device, int64_t S,
                     uint8_t *eqdistr, int64_t *eq_counts, int64_t *info, bool out_on_device, qecmc_stats *stats,
                     double *short_len, int64_t *short_n, int64_t *short_unique);

// the wave size comes from a cached free-memory figure: on an allocation failure, size once more from a fresh query
static int pteq_common(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *qm, bool qm_on_device, int64_t S,
                       uint8_t *eqdistr, int64_t *eq_counts, int64_t *info, bool out_on_device, qecmc_stats *stats,
                       double *short_len = nullptr, int64_t *short_n = nullptr, int64_t *short_unique = nullptr)
{
    if (!c) return set_err(QECMC_ERR_ARG, "NULL argument");
    return with_fresh_memory_retry(c, [&] {
        return pteq_once(c, cfg, qm, qm_on_device, S, eqdistr, eq_counts, info, out_on_device, stats, short_len, short_n, short_unique);
    });
}

static int pteq_once(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *qm, bool qm_on_device, int64_t S,
                     uint8_t *eqdistr, int64_t *eq_counts, int64_t *info, bool out_on_device, qecmc_stats *stats,
                     double *short_len, int64_t *short_n, int64_t *short_unique)
{
    if (!c || !cfg || !qm || !eqdistr) return set_err(QECMC_ERR_ARG, "NULL argument");
    const qecmc_ladder_cfg *lc = &cfg->ladder;
    QTRY(check_ladder_cfg(lc));
    if (S <= 0 || cfg->steps <= 0) return set_err(QECMC_ERR_ARG, "S and steps must be > 0");
    if (cfg->steps >= (1ll << 31)) return set_err(QECMC_ERR_UNSUPPORTED, "steps must be < 2^31");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    const Geo g = make_geo(lc->geom, lc->L);
    const bool wide = lc->L > 16;
    const size_t wb = wide ? 8 : 4;
    LadderDev &d = c->ld;   // device buffers persist in the context: no cudaMalloc / cudaFree per call
    LadderParams p;
    QTRY(setup_ladder(c, lc, g, d, p));
    // n_err history: 4 bytes per Ladder.step per ladder; ladders run in waves that fit the budget
    int64_t wave = S;
    const bool shortest = short_len != nullptr;
    if (shortest && (!short_n || !short_unique)) return set_err(QECMC_ERR_ARG, "short_len, short_n and short_unique go together");
    uint64_t scap = 0;
    if (shortest) {
        scap = next_pow2((uint64_t)cfg->steps + (uint64_t)cfg->steps / 4 + 1);
        if (scap < 1024) scap = 1024;
    }
    // native draws run on the rung-major kernel (qecmc_pt.cuh): one launch for the whole batch, ladders handed to the CTAs
    // from a queue, the n_err history sized by the RESIDENT ladders (2 bytes per Ladder.step, 4 for alpha ladders)
    const bool use_pt = !lc->u_nb && !shortest && !c->dbg_ladder_kernel;
    if (!use_pt && (cfg->use_conv || shortest)) {
        size_t fr = 0;
        QTRY(free_device_bytes(c, &fr));
        // what the context already holds for these two purposes counts as available (as in the STDC / PTDC drivers)
        int64_t budget = c->table_budget ? c->table_budget : (int64_t)((double)(fr + d.hist.cap + (shortest ? c->tables.cap : 0)) * 0.8);
        wave = budget / ((cfg->use_conv ? cfg->steps * 4 : 0) + (int64_t)scap * 8);
        if (wave < 1) return set_err(QECMC_ERR_NOMEM, "the n_err history of one ladder needs %lld bytes, budget is %lld",
                                     (long long)cfg->steps * 4, (long long)budget);
        if (wave > S) wave = S;
        if (cfg->use_conv) QTRY(d.hist.ensure((size_t)wave * cfg->steps * 4));
        if (shortest) QTRY(c->tables.ensure((size_t)wave * scap * 8));
    }
    if (shortest) {
        QTRY(d.short_v.ensure((size_t)S * g.neq * sizeof(double)));
        QTRY(d.short_n.ensure((size_t)S * g.neq * sizeof(long long)));
        QTRY(d.short_u.ensure((size_t)S * g.neq * sizeof(long long)));
        std::vector<double> init((size_t)S * g.neq, 100000.0);  // decoders_biasednoise.py:116
        CUDA_OK(cudaMemcpyAsync(d.short_v.p, init.data(), init.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemsetAsync(d.short_n.p, 0, (size_t)S * g.neq * sizeof(long long), c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    if (!qm_on_device) {
        QTRY(d.qm.ensure((size_t)S * g.nsites));
        CUDA_OK(cudaMemcpyAsync(d.qm.p, qm, (size_t)S * g.nsites, cudaMemcpyHostToDevice, c->stream));
    }
    const uint8_t *d_qm = qm_on_device ? qm : (const uint8_t *)d.qm.p;
    QTRY(d.lat.ensure((size_t)S * g.nw * wb));
    QTRY(d.eqc.ensure((size_t)S * g.neq * sizeof(long long)));
    QTRY(d.info.ensure((size_t)S * 4 * sizeof(long long)));
    QTRY(d.pct.ensure((size_t)S * g.neq));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaMemsetAsync(d.eqc.p, 0, (size_t)S * g.neq * sizeof(long long), c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, d_qm, S, g, d.lat.p));
    else QTRY(pack_lattices<uint32_t>(c, d_qm, S, g, d.lat.p));
    QTRY(stage_replay(c, lc, S, d, p));
    p.acct = ACCT_PTEQ;
    p.steps = cfg->steps;
    p.init_broadcast = 1;
    p.SEQ = cfg->SEQ;
    p.TOPS = cfg->TOPS;
    p.tops_burn = cfg->tops_burn;
    p.use_conv = cfg->use_conv;
    p.eps = cfg->eps;
    p.counters = (unsigned long long *)c->counters.p;
    uint8_t *d_pct = out_on_device ? eqdistr : (uint8_t *)d.pct.p;
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    int64_t waves = 0;
    const double *u_nb0 = p.u_nb, *u_py0 = p.u_py;
    if (use_pt) {
        QTRY(prepare_pt(c, p));
        PtPlan pl;
        QTRY(qecmc_pt_plan(c, p, &pl));
        int64_t grid = (S + pl.NLC - 1) / pl.NLC;
        if (grid > pl.max_grid) grid = pl.max_grid;
        const int64_t hb = lc->kind == LK_ALPHA ? 4 : 2;
        if (cfg->use_conv) {
            size_t fr = 0;
            QTRY(free_device_bytes(c, &fr));
            const int64_t budget = c->table_budget ? c->table_budget : (int64_t)((double)(fr + d.hist.cap) * 0.8);
            const int64_t per_cta = (int64_t)pl.NLC * cfg->steps * hb;
            if (grid * per_cta > budget) grid = budget / per_cta;   // fewer resident ladders: the queue feeds them all the same
            if (grid < 1) return set_err(QECMC_ERR_NOMEM, "the n_err history of %d resident ladders needs %lld bytes, budget is %lld",
                                         pl.NLC, (long long)per_cta, (long long)budget);
            QTRY(d.hist.ensure((size_t)(grid * per_cta)));
        }
        p.n_ladders = S;
        p.ladder_offset = 0;
        p.lat_in = d.lat.p;
        p.eq_counts = (long long *)d.eqc.p;
        p.info = (long long *)d.info.p;
        p.percent = d_pct;
        QTRY(qecmc_pt_launch(c, p, pl, (int)grid, 0u, cfg->use_conv ? d.hist.p : nullptr, cfg->steps));
        waves = 1;
    }
    for (int64_t s0 = 0; !use_pt && s0 < S; s0 += wave, waves++) {
        int64_t sw = S - s0 < wave ? S - s0 : wave;
        p.n_ladders = sw;
        p.ladder_offset = s0;
        p.lat_in = (const char *)d.lat.p + (size_t)s0 * g.nw * wb;
        p.hist = cfg->use_conv ? (uint32_t *)d.hist.p : nullptr;
        p.eq_counts = (long long *)d.eqc.p + s0 * g.neq;
        p.info = (long long *)d.info.p + s0 * 4;
        p.percent = d_pct + s0 * g.neq;
        if (u_nb0) { p.u_nb = u_nb0 + s0 * p.n_nb; p.u_py = u_py0 + s0 * p.n_py; }
        if (shortest) {
            CUDA_OK(cudaMemsetAsync(c->tables.p, 0, (size_t)sw * scap * 8, c->stream));
            p.track_shortest = 1;
            p.tables = (unsigned long long *)c->tables.p;
            p.cap_mask = scap - 1;
            p.short_v = (double *)d.short_v.p + s0 * g.neq;
            p.short_n = (long long *)d.short_n.p + s0 * g.neq;
        }
        QTRY(launch_ladder(c, p, lc->u_nb != nullptr));
        if (shortest) {
            short_unique_kernel<<<(unsigned)sw, 256, 0, c->stream>>>((const unsigned long long *)c->tables.p, scap, g.neq,
                                                                      lc->kind == LK_ALPHA, lc->param_b, p.short_v,
                                                                      (long long *)d.short_u.p + s0 * g.neq);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        }
    }
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    QTRY(check_status(c, d));
    std::vector<long long> h_info((size_t)S * 4);
    CUDA_OK(cudaMemcpyAsync(h_info.data(), d.info.p, h_info.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (!out_on_device) CUDA_OK(cudaMemcpyAsync(eqdistr, d.pct.p, (size_t)S * g.neq, cudaMemcpyDeviceToHost, c->stream));
    if (eq_counts) CUDA_OK(cudaMemcpyAsync(eq_counts, d.eqc.p, (size_t)S * g.neq * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (shortest) {
        CUDA_OK(cudaMemcpyAsync(short_len, d.short_v.p, (size_t)S * g.neq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(short_n, d.short_n.p, (size_t)S * g.neq * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(short_unique, d.short_u.p, (size_t)S * g.neq * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    }
    unsigned long long cnt[8] = {0};
    CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    int64_t steps_total = 0;
    for (int64_t s = 0; s < S; s++) steps_total += h_info[(size_t)4 * s] * lc->Nc * lc->iters;
    if (info) memcpy(info, h_info.data(), h_info.size() * sizeof(long long));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    fill_stats(c, stats, steps_total, cnt, ms, waves);
    return 0;
}

extern "C" int qecmc_pteq(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *qm, int64_t S, uint8_t *eqdistr,
                          int64_t *eq_counts, int64_t *info, qecmc_stats *stats)
{
    return pteq_common(c, cfg, qm, false, S, eqdistr, eq_counts, info, false, stats);
}

extern "C" int qecmc_pteq_shortest(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *qm, int64_t S, uint8_t *eqdistr,
                                   double *short_len, int64_t *short_n, int64_t *short_unique, int64_t *info, qecmc_stats *stats)
{
    if (!short_len || !short_n || !short_unique) return set_err(QECMC_ERR_ARG, "NULL argument");
    return pteq_common(c, cfg, qm, false, S, eqdistr, nullptr, info, false, stats, short_len, short_n, short_unique);
}

extern "C" int qecmc_pteq_dev(qecmc_ctx *c, const qecmc_pteq_cfg *cfg, const uint8_t *d_qm, int64_t S, uint8_t *d_eqdistr,
                              int64_t *info, qecmc_stats *stats)
{
    return pteq_common(c, cfg, d_qm, true, S, d_eqdistr, nullptr, info, true, stats);
}

// ------------------------------------------------------------------------------------------------
// Distinct-chain decoders on ladders: PTDC (decoders.py:138-233) and the EWD-style
// STDC_Nall_n_alpha / STDC_droplet_alpha (decoders.py:510-581, a one-rung alpha "ladder").
static int dc_once(qecmc_ctx *c, const qecmc_ladder_cfg *lc, int per_class_inits, int droplets, int64_t steps, double beta,
                   const uint8_t *qm, int64_t S, double *eqdistr, int64_t *distinct, qecmc_stats *stats, bool rc,
                   int64_t *N_hist, int64_t *m_hist, double conv_mult, int64_t *steps_done);

static int dc_common(qecmc_ctx *c, const qecmc_ladder_cfg *lc, int per_class_inits, int droplets, int64_t steps, double beta,
                     const uint8_t *qm, int64_t S, double *eqdistr, int64_t *distinct, qecmc_stats *stats, bool rc = false,
                     int64_t *N_hist = nullptr, int64_t *m_hist = nullptr, double conv_mult = 0.0, int64_t *steps_done = nullptr)
{
    if (!c) return set_err(QECMC_ERR_ARG, "NULL argument");
    return with_fresh_memory_retry(c, [&] {
        return dc_once(c, lc, per_class_inits, droplets, steps, beta, qm, S, eqdistr, distinct, stats, rc, N_hist, m_hist, conv_mult, steps_done);
    });
}

static int dc_once(qecmc_ctx *c, const qecmc_ladder_cfg *lc, int per_class_inits, int droplets, int64_t steps, double beta,
                   const uint8_t *qm, int64_t S, double *eqdistr, int64_t *distinct, qecmc_stats *stats, bool rc,
                   int64_t *N_hist, int64_t *m_hist, double conv_mult, int64_t *steps_done)
{
    if (!c || !qm || !eqdistr) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_ladder_cfg(lc));
    if (S <= 0 || steps <= 0 || droplets <= 0) return set_err(QECMC_ERR_ARG, "S, steps and droplets must be > 0");
    if (lc->p_logical != 0.0) return set_err(QECMC_ERR_UNSUPPORTED, "distinct-chain ladders run with p_logical = 0 (as PTDC does)");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    const Geo g = make_geo(lc->geom, lc->L);
    const bool wide = lc->L > 16;
    const size_t wb = wide ? 8 : 4;
    const int n_eq = g.neq;
    LadderDev &d = c->ld;   // device buffers persist in the context: no cudaMalloc / cudaFree per call
    LadderParams p;
    QTRY(setup_ladder(c, lc, g, d, p));
    // PTDC: one set per (syndrome, class) -- with the early stop one per ladder ("new" = new to the droplet), united per class
    // afterwards; PTRC: one per (syndrome, class, droplet, rung)
    if (conv_mult < 0.0) return set_err(QECMC_ERR_ARG, "conv_mult must be >= 0");
    if (conv_mult != 0.0 && (rc || lc->kind != LK_DEPOL)) return set_err(QECMC_ERR_UNSUPPORTED, "the early stop belongs to PTDC");
    const bool conv = conv_mult != 0.0;
    uint64_t max_keys = rc ? (uint64_t)steps : (uint64_t)(conv ? 1 : droplets) * (uint64_t)steps * (uint64_t)lc->Nc;
    uint64_t cap = next_pow2(max_keys + max_keys / 4 + 1);
    if (cap < 1024) cap = 1024;
    const int64_t tabs_per_class = rc ? (int64_t)droplets * lc->Nc : conv ? (int64_t)droplets : 1;
    // union tables of the early stop with several droplets: one per (syndrome, class), behind the ladders' own
    const bool unite = conv && droplets > 1;
    uint64_t ucap = 0;
    if (unite) {
        const uint64_t uk = (uint64_t)droplets * (uint64_t)steps * (uint64_t)lc->Nc;
        ucap = next_pow2(uk + uk / 4 + 1);
        if (ucap < 1024) ucap = 1024;
    }
    const int ns1 = g.nsites + 1;
    size_t fr = 0;
    QTRY(free_device_bytes(c, &fr));
    int64_t budget = c->table_budget ? c->table_budget : (int64_t)((double)(fr + c->tables.cap) * 0.85);
    int64_t per_syndrome = (int64_t)n_eq * (tabs_per_class * ((int64_t)cap * 8 + (rc ? (int64_t)ns1 * 12 : 0)) + (int64_t)ucap * 8);
    int64_t wave = budget / per_syndrome;
    if (wave < 1) return set_err(QECMC_ERR_NOMEM, "distinct-chain tables need %lld bytes per syndrome, budget is %lld",
                                 (long long)per_syndrome, (long long)budget);
    if (wave > S) wave = S;
    QTRY(c->tables.ensure((size_t)wave * per_syndrome));
    const int64_t n_in = per_class_inits ? S * n_eq : S;
    QTRY(d.qm.ensure((size_t)n_in * g.nsites));
    QTRY(d.lat.ensure((size_t)n_in * g.nw * wb));
    QTRY(d.lat_out.ensure((size_t)S * n_eq * droplets * g.nw * wb));  // one init lattice per ladder
    QTRY(d.Zd.ensure((size_t)S * n_eq * sizeof(double)));
    QTRY(d.dist.ensure((size_t)S * n_eq * sizeof(unsigned long long)));
    QTRY(c->out_f64.ensure((size_t)S * n_eq * sizeof(double)));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaMemcpyAsync(d.qm.p, qm, (size_t)n_in * g.nsites, cudaMemcpyHostToDevice, c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, (const uint8_t *)d.qm.p, n_in, g, d.lat.p));
    else QTRY(pack_lattices<uint32_t>(c, (const uint8_t *)d.qm.p, n_in, g, d.lat.p));
    // per-class initial states [S][n_eq][nw]
    ScopedDevBuf cls_lat;   // call-local buffers: released on every return path
    const void *class_lat = d.lat.p;
    if (!per_class_inits) {
        QTRY(cls_lat.ensure((size_t)S * n_eq * g.nw * wb));
        if (wide) QTRY(to_class_all<uint64_t>(c, g, d.lat.p, cls_lat.p, S));
        else QTRY(to_class_all<uint32_t>(c, g, d.lat.p, cls_lat.p, S));
        class_lat = cls_lat.p;
    }
    // ladder j = (syndrome, class, droplet) starts from its class lattice
    for (int dr = 0; dr < droplets; dr++)
        CUDA_OK(cudaMemcpy2DAsync((char *)d.lat_out.p + (size_t)dr * g.nw * wb, (size_t)droplets * g.nw * wb, class_lat,
                                  (size_t)g.nw * wb, (size_t)g.nw * wb, (size_t)S * n_eq, cudaMemcpyDeviceToDevice, c->stream));
    QTRY(stage_replay(c, lc, S * n_eq * droplets, d, p));
    p.acct = rc ? ACCT_RC : ACCT_DC;
    p.steps = steps;
    p.conv_mult = conv_mult;
    if (steps_done) QTRY(d.info.ensure((size_t)S * n_eq * droplets * sizeof(long long)));
    ScopedDevBuf rc_m, rc_N, rc_Nout, rc_mout, lad_p;
    if (rc) {
        QTRY(rc_m.ensure((size_t)wave * n_eq * tabs_per_class * ns1 * sizeof(unsigned long long)));
        QTRY(rc_N.ensure((size_t)wave * n_eq * tabs_per_class * ns1 * sizeof(uint32_t)));
        QTRY(rc_Nout.ensure((size_t)S * n_eq * lc->Nc * ns1 * sizeof(long long)));
        QTRY(rc_mout.ensure((size_t)S * n_eq * lc->Nc * ns1 * sizeof(long long)));
        LadderTables lt;
        make_ladder_tables(lc, g, lt);
        QTRY(upload(c, lad_p, lt.ladder));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        p.rc_m_hist = (unsigned long long *)rc_m.p;
    }
    p.init_broadcast = 1;
    p.droplets = droplets;
    p.aux_bits = lc->kind == LK_ALPHA ? 22 : QECMC_LEN_BITS;
    p.tables = (unsigned long long *)c->tables.p;
    p.cap_mask = cap - 1;
    p.counters = (unsigned long long *)c->counters.p;
    const double *u_nb0 = p.u_nb, *u_py0 = p.u_py;
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    int64_t waves = 0;
    for (int64_t s0 = 0; s0 < S; s0 += wave, waves++) {
        int64_t sw = S - s0 < wave ? S - s0 : wave;
        CUDA_OK(cudaMemsetAsync(c->tables.p, 0, (size_t)sw * n_eq * (tabs_per_class * cap + ucap) * 8, c->stream));
        if (rc) CUDA_OK(cudaMemsetAsync(rc_m.p, 0, (size_t)sw * n_eq * tabs_per_class * ns1 * sizeof(unsigned long long), c->stream));
        const int64_t l0 = s0 * n_eq * droplets;
        p.n_ladders = sw * n_eq * droplets;
        p.ladder_offset = l0;
        p.lat_in = (const char *)d.lat_out.p + (size_t)l0 * g.nw * wb;
        p.info = steps_done ? (long long *)d.info.p + l0 : nullptr;
        if (u_nb0) { p.u_nb = u_nb0 + l0 * p.n_nb; p.u_py = u_py0 + l0 * p.n_py; }
        QTRY(launch_ladder(c, p, lc->u_nb != nullptr));
        const int64_t tabs = sw * n_eq;
        const unsigned long long *class_tables = (const unsigned long long *)c->tables.p;
        uint64_t class_cap = cap;
        if (unite) {   // the droplets' sets of a class -> the class's set
            unsigned long long *ut = (unsigned long long *)c->tables.p + (size_t)sw * n_eq * tabs_per_class * cap;
            table_union_kernel<<<(unsigned)tabs, 512, 0, c->stream>>>((const unsigned long long *)c->tables.p, cap, droplets, ut, ucap,
                                                                      p.aux_bits);
            c->launches++;
            class_tables = ut;
            class_cap = ucap;
        }
        if (rc) {
            table_hist_kernel<<<(unsigned)(tabs * tabs_per_class), 256, ns1 * sizeof(uint32_t), c->stream>>>(
                (const unsigned long long *)c->tables.p, cap, g.nsites, 0.0, nullptr, (uint32_t *)rc_N.p,
                (unsigned long long *)c->counters.p + 3);
            c->launches++;
            ptrc_finalize_kernel<<<(unsigned)((tabs + 63) / 64), 64, 0, c->stream>>>(
                (const uint32_t *)rc_N.p, (const unsigned long long *)rc_m.p, tabs, droplets, lc->Nc, ns1, (const double *)lad_p.p,
                beta, (double *)d.Zd.p + s0 * n_eq, (long long *)rc_Nout.p + (size_t)s0 * n_eq * lc->Nc * ns1,
                (long long *)rc_mout.p + (size_t)s0 * n_eq * lc->Nc * ns1);
        } else if (lc->kind == LK_ALPHA) {
            table_sum_alpha_kernel<<<(unsigned)tabs, 256, 0, c->stream>>>((const unsigned long long *)c->tables.p, cap, beta,
                                                                          lc->param_b, (double *)d.Zd.p + s0 * n_eq,
                                                                          (unsigned long long *)d.dist.p + s0 * n_eq,
                                                                          (unsigned long long *)c->counters.p + 3);
        } else {
            QTRY(d.hist.ensure((size_t)tabs * (g.nsites + 1) * sizeof(uint32_t)));
            table_hist_kernel<<<(unsigned)tabs, 512, (g.nsites + 1) * sizeof(uint32_t), c->stream>>>(
                class_tables, class_cap, g.nsites, beta, (double *)d.Zd.p + s0 * n_eq, (uint32_t *)d.hist.p,
                (unsigned long long *)c->counters.p + 3);
        }
        c->launches++;
        CUDA_OK(cudaGetLastError());
    }
    normalize_kernel<<<(unsigned)((S + 127) / 128), 128, 0, c->stream>>>((const double *)d.Zd.p, (double *)c->out_f64.p, S, n_eq);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    QTRY(check_status(c, d));
    CUDA_OK(cudaMemcpyAsync(eqdistr, c->out_f64.p, (size_t)S * n_eq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (distinct && lc->kind == LK_ALPHA)
        CUDA_OK(cudaMemcpyAsync(distinct, d.dist.p, (size_t)S * n_eq * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (rc && N_hist) CUDA_OK(cudaMemcpyAsync(N_hist, rc_Nout.p, (size_t)S * n_eq * lc->Nc * ns1 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (steps_done) CUDA_OK(cudaMemcpyAsync(steps_done, d.info.p, (size_t)S * n_eq * droplets * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (rc && m_hist) CUDA_OK(cudaMemcpyAsync(m_hist, rc_mout.p, (size_t)S * n_eq * lc->Nc * ns1 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    unsigned long long cnt[8] = {0};
    CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    fill_stats(c, stats, S * n_eq * droplets * lc->Nc * steps * lc->iters, cnt, ms, waves);
    if (stats) stats->table_slots = (int64_t)cap;
    return 0;
}

extern "C" int qecmc_stdc_alpha(qecmc_ctx *c, const qecmc_alpha_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                                int64_t *distinct, qecmc_stats *stats)
{
    if (!cfg) return set_err(QECMC_ERR_ARG, "cfg is NULL");
    if (!(cfg->pz_tilde > 0)) return set_err(QECMC_ERR_ARG, "pz_tilde must be > 0");
    qecmc_ladder_cfg lc;
    memset(&lc, 0, sizeof(lc));
    lc.geom = cfg->geom;
    lc.L = cfg->L;
    lc.kind = LK_ALPHA;
    lc.Nc = 1;
    lc.iters = cfg->iters;
    lc.bottom = cfg->pz_tilde_sampling;
    lc.param_b = cfg->alpha;
    lc.p_logical = 0.0;
    lc.seed = cfg->seed;
    lc.u_nb = cfg->u_nb;
    lc.u_py = cfg->u_py;
    lc.n_nb = cfg->n_nb;
    lc.n_py = cfg->n_py;
    const double beta = -log(cfg->pz_tilde);  // decoders.py:568
    return dc_common(c, &lc, cfg->per_class_inits, 1, cfg->steps, beta, qm, S, eqdistr, distinct, stats);
}

extern "C" int qecmc_ptdc(qecmc_ctx *c, const qecmc_ptdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                          qecmc_stats *stats)
{
    if (!cfg) return set_err(QECMC_ERR_ARG, "cfg is NULL");
    if (cfg->ladder.kind != LK_DEPOL) return set_err(QECMC_ERR_ARG, "PTDC runs depolarizing ladders");
    if (!(cfg->p_error > 0 && cfg->p_error < 1)) return set_err(QECMC_ERR_ARG, "p_error outside (0,1)");
    const double beta = -log((cfg->p_error / 3) / (1 - cfg->p_error));  // decoders.py:205
    return dc_common(c, &cfg->ladder, cfg->per_class_inits, cfg->droplets, cfg->steps, beta, qm, S, eqdistr, nullptr, stats, false,
                     nullptr, nullptr, cfg->conv_mult, cfg->steps_done);
}

extern "C" int qecmc_ptrc(qecmc_ctx *c, const qecmc_ptdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                          int64_t *N_hist, int64_t *m_hist, qecmc_stats *stats)
{
    if (!cfg) return set_err(QECMC_ERR_ARG, "cfg is NULL");
    if (cfg->ladder.kind != LK_DEPOL) return set_err(QECMC_ERR_ARG, "PTRC runs depolarizing ladders");
    if (!(cfg->p_error > 0 && cfg->p_error < 1)) return set_err(QECMC_ERR_ARG, "p_error outside (0,1)");
    if (cfg->ladder.Nc < 2) return set_err(QECMC_ERR_ARG, "PTRC needs at least two rungs (the top rung is not used in the estimate)");
    const double beta = -log((cfg->p_error / 3) / (1 - cfg->p_error));  // decoders.py:679
    return dc_common(c, &cfg->ladder, cfg->per_class_inits, cfg->droplets, cfg->steps, beta, qm, S, eqdistr, nullptr, stats, true,
                     N_hist, m_hist);
}

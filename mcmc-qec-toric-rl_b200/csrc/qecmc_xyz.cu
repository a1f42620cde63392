// qecmc_xyz.cu -- C ABI of the general-noise decoders (include/qecmc.h): qecmc_stdc_general_noise, qecmc_chain_update_xyz.
#include "qecmc_internal.h"
#include "qecmc_xyz.cuh"

using namespace qecmc;

namespace {

// accept thresholds by (dnx, dny, dnz) in [-4, 4]^3
void make_xyz_thresholds(bool use_xyz, const double *ps_xyz, double p_sampling, std::vector<double> &d, std::vector<uint32_t> &u)
{
    d.assign(729, 0.0);
    u.assign(729, 0u);
    double f[3] = {0, 0, 0};
    if (use_xyz) {
        double tot = ps_xyz[0] + ps_xyz[1] + ps_xyz[2];             // self.p_xyz.sum(), mcmc.py:110
        for (int i = 0; i < 3; i++) f[i] = ps_xyz[i] / (1.0 - tot);
    }
    const double factor = (p_sampling / 3.0) / (1.0 - p_sampling);  // mcmc.py:16
    for (int dx = -4; dx <= 4; dx++)
        for (int dy = -4; dy <= 4; dy++)
            for (int dz = -4; dz <= 4; dz++) {
                double v;
                if (use_xyz) {  // (factors ** change).prod(): numpy power, product from 1 in index order (mcmc.py:170)
                    volatile double a = 1.0 * pow(f[0], (double)dx);
                    volatile double b = a * pow(f[1], (double)dy);
                    v = b * pow(f[2], (double)dz);
                } else {
                    v = numba_pow(factor, dx + dy + dz);               // _update_chain_fast, mcmc.py:158
                }
                int i = (dx + 4) * 81 + (dy + 4) * 9 + (dz + 4);
                d[i] = v;
                if (!(v < 1.0)) u[i] = 0xFFFFFFFFu;
                else { double x = ceil(v * 4294967296.0); u[i] = x < 1.0 ? 0u : (uint32_t)(x - 1.0); }
            }
}

template <int GEOM, typename W> int launch_xyz(qecmc_ctx *c, XyzParams &p, bool replay)
{
    QTRY((build_stab_hash<GEOM, W>(c, p.gchain, (uint64_t **)&p.stab_hash)));
    int T = 0, nb = 0;
    size_t per_chain = (size_t)p.gchain.nw * sizeof(W);
    QTRY(pick_threads(per_chain, 256, c->prop, &T, &nb));
    size_t smem = per_chain * T;
    unsigned grid = (unsigned)((p.n_chains + T - 1) / T);
    if (replay) {
        CUDA_OK(cudaFuncSetAttribute(xyz_kernel<GEOM, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xyz_kernel<GEOM, W, true><<<grid, T, smem, c->stream>>>(p);
    } else {
        CUDA_OK(cudaFuncSetAttribute(xyz_kernel<GEOM, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xyz_kernel<GEOM, W, false><<<grid, T, smem, c->stream>>>(p);
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename W> int launch_xyz_geom(qecmc_ctx *c, XyzParams &p, bool replay)
{
    switch (p.gchain.geom) {
    case TORIC: return launch_xyz<TORIC, W>(c, p, replay);
    case PLANAR: return launch_xyz<PLANAR, W>(c, p, replay);
    case ROTATED: return launch_xyz<ROTATED, W>(c, p, replay);
    default: return launch_xyz<XZZX, W>(c, p, replay);
    }
}

template <int GEOM, typename W> int launch_xyz_chain(qecmc_ctx *c, XyzChainParams &p, bool replay)
{
    int T = 0, nb = 0;
    size_t per_chain = (size_t)p.g.nw * sizeof(W);
    QTRY(pick_threads(per_chain, 256, c->prop, &T, &nb));
    size_t smem = per_chain * T;
    unsigned grid = (unsigned)((p.chains + T - 1) / T);
    if (replay) {
        CUDA_OK(cudaFuncSetAttribute(xyz_chain_kernel<GEOM, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xyz_chain_kernel<GEOM, W, true><<<grid, T, smem, c->stream>>>(p);
    } else {
        CUDA_OK(cudaFuncSetAttribute(xyz_chain_kernel<GEOM, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xyz_chain_kernel<GEOM, W, false><<<grid, T, smem, c->stream>>>(p);
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename W> int launch_xyz_chain_geom(qecmc_ctx *c, XyzChainParams &p, bool replay)
{
    switch (p.g.geom) {
    case TORIC: return launch_xyz_chain<TORIC, W>(c, p, replay);
    case PLANAR: return launch_xyz_chain<PLANAR, W>(c, p, replay);
    case ROTATED: return launch_xyz_chain<ROTATED, W>(c, p, replay);
    default: return launch_xyz_chain<XZZX, W>(c, p, replay);
    }
}

}  // namespace

extern "C" int qecmc_chain_update_xyz(qecmc_ctx *c, const qecmc_chain_cfg *cfg, const double *p_xyz, const double *u, uint8_t *qm,
                                      int64_t chains, int64_t iters, qecmc_stats *stats)
{
    if (!c || !cfg || !p_xyz || !qm) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (chains <= 0 || iters < 0) return set_err(QECMC_ERR_ARG, "chains must be > 0 and iters >= 0");
    QTRY(check_geom(cfg->geom_chain, cfg->L));
    double tot = p_xyz[0] + p_xyz[1] + p_xyz[2];
    if (!(p_xyz[0] > 0 && p_xyz[1] > 0 && p_xyz[2] > 0 && tot < 1)) return set_err(QECMC_ERR_ARG, "p_xyz must be positive and sum to < 1");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    Geo g = make_geo(cfg->geom_chain, cfg->L);
    const bool wide = cfg->L > 16;
    const size_t wb = wide ? 8 : 4, nbytes = (size_t)chains * g.nsites;
    std::vector<double> thr_d;
    std::vector<uint32_t> thr_u;
    make_xyz_thresholds(true, p_xyz, 0.5, thr_d, thr_u);
    QTRY(c->lut.ensure(729 * 12 + 16));
    QTRY(c->qm_in.ensure(nbytes));
    QTRY(c->packed.ensure((size_t)chains * g.nw * wb));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemcpyAsync(c->lut.p, thr_d.data(), 729 * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync((char *)c->lut.p + 729 * 8, thr_u.data(), 729 * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->qm_in.p, qm, nbytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, (const uint8_t *)c->qm_in.p, chains, g, c->packed.p));
    else QTRY(pack_lattices<uint32_t>(c, (const uint8_t *)c->qm_in.p, chains, g, c->packed.p));
    XyzChainParams p;
    memset(&p, 0, sizeof(p));
    p.g = g;
    p.lat = c->packed.p;
    p.chains = chains;
    p.iters = iters;
    p.seed = cfg->seed;
    p.offset = cfg->stream_offset;
    p.thr_d = (const double *)c->lut.p;
    p.thr_u = (const uint32_t *)((const char *)c->lut.p + 729 * 8);
    p.counters = (unsigned long long *)c->counters.p;
    if (u) {
        int k = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) ? 3 : 5;
        size_t n_u = (size_t)chains * iters * (k + 1) * sizeof(double);
        QTRY(c->replay_a.ensure(n_u + 8));
        CUDA_OK(cudaMemcpyAsync(c->replay_a.p, u, n_u, cudaMemcpyHostToDevice, c->stream));
        p.u = (const double *)c->replay_a.p;
    }
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    if (wide) QTRY(launch_xyz_chain_geom<uint64_t>(c, p, u != nullptr));
    else QTRY(launch_xyz_chain_geom<uint32_t>(c, p, u != nullptr));
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    const int T = 256;
    const int64_t n_words = chains * g.nw;
    if (wide) unpack_kernel<uint64_t><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>((uint64_t *)c->packed.p, (uint8_t *)c->qm_in.p, n_words, g.L);
    else unpack_kernel<uint32_t><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>((uint32_t *)c->packed.p, (uint8_t *)c->qm_in.p, n_words, g.L);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(qm, c->qm_in.p, nbytes, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long cnt[8] = {0};
    CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->metropolis_steps = chains * iters;
        stats->accepted = (int64_t)cnt[0];
        stats->kernel_launches = c->launches;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        stats->chain_kernel_ms = ms;
        stats->total_ms = ms;
    }
    return 0;
}

static int general_noise_once(qecmc_ctx *c, const qecmc_xyz_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                              double *eqdistr_shortest, int64_t *distinct, qecmc_stats *stats);

extern "C" int qecmc_stdc_general_noise(qecmc_ctx *c, const qecmc_xyz_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                                        double *eqdistr_shortest, int64_t *distinct, qecmc_stats *stats)
{
    if (!c) return set_err(QECMC_ERR_ARG, "NULL argument");
    // the wave size comes from a cached free-memory figure: on an allocation failure, size once more from a fresh query
    return with_fresh_memory_retry(c, [&] { return general_noise_once(c, cfg, qm, S, eqdistr, eqdistr_shortest, distinct, stats); });
}

static int general_noise_once(qecmc_ctx *c, const qecmc_xyz_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                              double *eqdistr_shortest, int64_t *distinct, qecmc_stats *stats)
{
    if (!c || !cfg || !qm || !eqdistr) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    QTRY(check_geom(cfg->geom_code, cfg->L));
    QTRY(check_geom(cfg->geom_chain, cfg->L));
    Geo gcode = make_geo(cfg->geom_code, cfg->L), gchain = make_geo(cfg->geom_chain, cfg->L);
    if (gcode.layers != gchain.layers) return set_err(QECMC_ERR_ARG, "geom_code and geom_chain have different lattice shapes");
    if (cfg->droplets <= 0 || cfg->steps <= 0 || cfg->iters <= 0) return set_err(QECMC_ERR_ARG, "droplets, steps, iters must be > 0");
    if ((uint64_t)cfg->steps * (uint64_t)cfg->iters >= (1ull << 32)) return set_err(QECMC_ERR_UNSUPPORTED, "steps * iters must be < 2^32");
    double psum = 0;
    for (int i = 0; i < 3; i++) {
        if (!(cfg->p_xyz[i] >= 0 && cfg->p_xyz[i] < 1)) return set_err(QECMC_ERR_ARG, "p_xyz[%d] outside [0,1)", i);
        if (cfg->use_xyz_sampling && !(cfg->p_sampling_xyz[i] > 0)) return set_err(QECMC_ERR_ARG, "p_sampling_xyz[%d] must be > 0", i);
        psum += cfg->p_sampling_xyz[i];
    }
    if (cfg->use_xyz_sampling && !(psum < 1)) return set_err(QECMC_ERR_ARG, "p_sampling_xyz sums to >= 1");
    if (!cfg->use_xyz_sampling && !(cfg->p_sampling > 0 && cfg->p_sampling < 1)) return set_err(QECMC_ERR_ARG, "p_sampling outside (0,1)");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    const bool wide = cfg->L > 16;
    const size_t wbytes = wide ? 8 : 4;
    const int n_eq = gcode.neq;
    uint64_t max_keys = (uint64_t)cfg->droplets * (uint64_t)cfg->steps;
    uint64_t cap = next_pow2(max_keys + max_keys / 4 + 1);
    if (cap < 1024) cap = 1024;
    size_t fr = 0;
    QTRY(free_device_bytes(c, &fr));
    int64_t budget = c->table_budget ? c->table_budget : (int64_t)((double)(fr + c->tables.cap) * 0.85);
    int64_t per_syndrome = (int64_t)n_eq * (int64_t)cap * 16;
    int64_t wave = budget / per_syndrome;
    if (wave < 1) return set_err(QECMC_ERR_NOMEM, "distinct-chain tables need %lld bytes per syndrome, budget is %lld",
                                 (long long)per_syndrome, (long long)budget);
    if (wave > S) wave = S;
    QTRY(c->tables.ensure((size_t)wave * per_syndrome));
    const int64_t n_lat = cfg->per_class_inits ? S * n_eq : S;
    const size_t in_bytes = (size_t)n_lat * gcode.nsites;
    QTRY(c->qm_in.ensure(in_bytes));
    QTRY(c->packed.ensure((size_t)n_lat * gcode.nw * wbytes));
    QTRY(c->Z.ensure((size_t)S * n_eq * 2 * sizeof(double)));
    QTRY(c->out_f64.ensure((size_t)S * n_eq * 2 * sizeof(double)));
    QTRY(c->out_u64.ensure((size_t)S * n_eq * sizeof(unsigned long long)));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    std::vector<double> thr_d;
    std::vector<uint32_t> thr_u;
    make_xyz_thresholds(cfg->use_xyz_sampling != 0, cfg->p_sampling_xyz, cfg->p_sampling, thr_d, thr_u);
    QTRY(c->lut.ensure(729 * 12 + 16));
    CUDA_OK(cudaMemcpyAsync(c->lut.p, thr_d.data(), 729 * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync((char *)c->lut.p + 729 * 8, thr_u.data(), 729 * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->qm_in.p, qm, in_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, (const uint8_t *)c->qm_in.p, n_lat, gcode, c->packed.p));
    else QTRY(pack_lattices<uint32_t>(c, (const uint8_t *)c->qm_in.p, n_lat, gcode, c->packed.p));
    XyzParams p;
    memset(&p, 0, sizeof(p));
    p.gcode = gcode;
    p.gchain = gchain;
    p.per_class = cfg->per_class_inits;
    p.droplets = cfg->droplets;
    p.iters = cfg->iters;
    p.steps = cfg->steps;
    p.seed = cfg->seed;
    p.hash_seed = c->hash_seed;
    p.tables = (unsigned long long *)c->tables.p;
    p.cap_mask = cap - 1;
    p.thr_d = (const double *)c->lut.p;
    p.thr_u = (const uint32_t *)((const char *)c->lut.p + 729 * 8);
    p.counters = (unsigned long long *)c->counters.p;
    if (cfg->u_nb) {
        int k = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) ? 3 : 5;
        size_t nb = (size_t)S * n_eq * cfg->droplets * (size_t)cfg->steps * cfg->iters * (k + 1) * sizeof(double);
        QTRY(c->replay_a.ensure(nb));
        CUDA_OK(cudaMemcpyAsync(c->replay_a.p, cfg->u_nb, nb, cudaMemcpyHostToDevice, c->stream));
        p.u_nb = (const double *)c->replay_a.p;
    }
    double beta[3];
    for (int i = 0; i < 3; i++) beta[i] = -log((cfg->p_xyz[i] / 3) / (1 - cfg->p_xyz[i]));  // decoders.py:385
    double *Z_all = (double *)c->Z.p, *Z_short = Z_all + (size_t)S * n_eq;
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    int64_t waves = 0;
    for (int64_t s0 = 0; s0 < S; s0 += wave, waves++) {
        int64_t sw = S - s0 < wave ? S - s0 : wave;
        CUDA_OK(cudaMemsetAsync(c->tables.p, 0, (size_t)sw * per_syndrome, c->stream));
        p.lat0 = (const char *)c->packed.p + (size_t)(cfg->per_class_inits ? s0 * n_eq : s0) * gcode.nw * wbytes;
        p.n_chains = sw * n_eq * cfg->droplets;
        p.chain_offset = s0 * n_eq * cfg->droplets;
        if (wide) QTRY(launch_xyz_geom<uint64_t>(c, p, cfg->u_nb != nullptr));
        else QTRY(launch_xyz_geom<uint32_t>(c, p, cfg->u_nb != nullptr));
        table_xyz_kernel<<<(unsigned)(sw * n_eq), 256, 0, c->stream>>>((const unsigned long long *)c->tables.p, cap, beta[0], beta[1],
                                                                       beta[2], Z_all + s0 * n_eq, Z_short + s0 * n_eq,
                                                                       (unsigned long long *)c->out_u64.p + s0 * n_eq,
                                                                       (unsigned long long *)c->counters.p + 3);
        c->launches++;
        CUDA_OK(cudaGetLastError());
    }
    double *o_all = (double *)c->out_f64.p, *o_short = o_all + (size_t)S * n_eq;
    normalize_kernel<<<(unsigned)((S + 127) / 128), 128, 0, c->stream>>>(Z_all, o_all, S, n_eq);
    normalize_kernel<<<(unsigned)((S + 127) / 128), 128, 0, c->stream>>>(Z_short, o_short, S, n_eq);
    c->launches += 2;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    CUDA_OK(cudaMemcpyAsync(eqdistr, o_all, (size_t)S * n_eq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (eqdistr_shortest) CUDA_OK(cudaMemcpyAsync(eqdistr_shortest, o_short, (size_t)S * n_eq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (distinct) CUDA_OK(cudaMemcpyAsync(distinct, c->out_u64.p, (size_t)S * n_eq * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    unsigned long long cnt[8] = {0};
    CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->metropolis_steps = S * n_eq * (int64_t)cfg->droplets * cfg->steps * cfg->iters;
        stats->accepted = (int64_t)cnt[0];
        stats->samples = (int64_t)cnt[1];
        stats->distinct = (int64_t)cnt[3];
        stats->table_slots = (int64_t)cap;
        stats->waves = waves;
        stats->kernel_launches = c->launches;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        stats->chain_kernel_ms = ms;
        stats->total_ms = ms;
    }
    return 0;
}

// qecmc_api.cu -- C ABI (include/qecmc.h) over the CUDA kernels: context, device
// memory, wave scheduling of the distinct-chain tables, host<->device staging.
#include "qecmc_internal.h"
#include "qecmc_stdc_fast.cuh"
#include "qecmc_stdc_pk.cuh"
#include "qecmc_dedupe.cuh"

using namespace qecmc;

static thread_local char g_err[512] = "";
int qecmc_set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ------------------------------ context ------------------------------
extern "C" int qecmc_abi_version(void) { return QECMC_ABI_VERSION; }
extern "C" const char *qecmc_last_error(void) { return g_err; }

extern "C" int qecmc_create(int device, qecmc_ctx **out)
{
    if (!out) return set_err(QECMC_ERR_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(QECMC_ERR_CUDA, "no CUDA device available (%s); libqecmc has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return set_err(QECMC_ERR_ARG, "device %d out of range (have %d)", device, n);
    CUDA_OK(cudaSetDevice(device));
    qecmc_ctx *c = new qecmc_ctx();
    c->device = device;
    CUDA_OK(cudaGetDeviceProperties(&c->prop, device));
    CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto &ev : c->ev) CUDA_OK(cudaEventCreate(&ev));
    *out = c;
    return 0;
}

extern "C" void qecmc_destroy(qecmc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf *b : {&c->packed, &c->tables, &c->Z, &c->counters, &c->qm_in, &c->out_f64, &c->out_u32, &c->out_u64,
                      &c->out_i32, &c->replay_a, &c->replay_b, &c->scratch, &c->nhist, &c->mhist, &c->shorts, &c->sums})
        b->release();
    for (auto &kv : c->stab_hash) cudaFree(kv.second);
    for (auto &kv : c->stab_desc) cudaFree(kv.second);
    alloc_generation()++;
    c->lut.release();
    c->log_hash.release();
    c->log_counts.release();
    c->dd_scratch.release();
    for (auto &ev : c->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int qecmc_set_stream(qecmc_ctx *c, void *s)
{
    if (!c) return set_err(QECMC_ERR_ARG, "ctx is NULL");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

extern "C" int qecmc_set_table_budget(qecmc_ctx *c, int64_t bytes)
{
    if (!c || bytes < 0) return set_err(QECMC_ERR_ARG, "bad arguments");
    c->table_budget = bytes;
    return 0;
}

extern "C" int qecmc_debug_set(qecmc_ctx *c, const char *key, int64_t value)
{
    if (!c || !key) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (!strcmp(key, "force_wide")) c->dbg_force_wide = value > 0;
    else if (!strcmp(key, "insert_mode")) c->dbg_insert_mode = value < 0 ? -1 : (int)value;
    else if (!strcmp(key, "serial_sweep")) c->dbg_serial_sweep = value > 0;
    else if (!strcmp(key, "pt_lt")) c->dbg_pt_lt = value > 0 ? (int)value : 0;
    else if (!strcmp(key, "pt_grid")) c->dbg_pt_grid = value > 0 ? (int)value : 0;
    else if (!strcmp(key, "packed")) c->dbg_packed = value < 0 ? -1 : (int)value;
    else if (!strcmp(key, "ladder_kernel")) c->dbg_ladder_kernel = value > 0 ? (int)value : 0;
    else return set_err(QECMC_ERR_ARG, "unknown debug key '%s'", key);
    return 0;
}

extern "C" int qecmc_last_plan(qecmc_ctx *c, int64_t *wave_capacity, int64_t *round_chains)
{
    if (!c) return set_err(QECMC_ERR_ARG, "ctx is NULL");
    if (wave_capacity) *wave_capacity = c->plan_wave_cap;
    if (round_chains) *round_chains = c->plan_round_chains;
    return 0;
}

extern "C" int qecmc_device_info(qecmc_ctx *c, qecmc_devinfo *o)
{
    if (!c || !o) return set_err(QECMC_ERR_ARG, "NULL argument");
    CUDA_OK(cudaSetDevice(c->device));
    memset(o, 0, sizeof(*o));
    o->device = c->device;
    o->sm_count = c->prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    o->sm_clock_khz = khz;
    o->cc_major = c->prop.major;
    o->cc_minor = c->prop.minor;
    size_t fr = 0, tot = 0;
    CUDA_OK(cudaMemGetInfo(&fr, &tot));
    o->total_mem = (int64_t)tot;
    o->free_mem = (int64_t)fr;
    snprintf(o->name, sizeof(o->name), "%.63s", c->prop.name);   // device names longer than the field are cut
    return 0;
}

// ------------------------------ plain chains ------------------------------
template <int GEOM, typename W, bool REPLAY>
static int launch_chain(qecmc_ctx *c, ChainParams &p)
{
    int T = 0, nb = 0;
    size_t per_chain = (size_t)p.g.nw * sizeof(W);
    QTRY(pick_threads(per_chain, 256, c->prop, &T, &nb));
    size_t smem = per_chain * T;
    CUDA_OK(cudaFuncSetAttribute(chain_kernel<GEOM, W, REPLAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((p.chains + T - 1) / T);
    chain_kernel<GEOM, W, REPLAY><<<grid, T, smem, c->stream>>>(p);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename W, bool REPLAY> static int launch_chain_geom(qecmc_ctx *c, ChainParams &p)
{
    switch (p.g.geom) {
    case TORIC: return launch_chain<TORIC, W, REPLAY>(c, p);
    case PLANAR: return launch_chain<PLANAR, W, REPLAY>(c, p);
    case ROTATED: return launch_chain<ROTATED, W, REPLAY>(c, p);
    default: return launch_chain<XZZX, W, REPLAY>(c, p);
    }
}

static int chain_common(qecmc_ctx *c, const qecmc_chain_cfg *cfg, const uint8_t *qm_in, uint8_t *qm_out, const double *u,
                        int64_t chains, int64_t iters, int8_t *dE, uint8_t *accepted, uint8_t *traj, qecmc_stats *stats)
{
    if (!c || !cfg || !qm_in || !qm_out) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (chains <= 0 || iters < 0) return set_err(QECMC_ERR_ARG, "chains must be > 0 and iters >= 0");
    if (!(cfg->p > 0.0 && cfg->p < 1.0)) return set_err(QECMC_ERR_ARG, "p=%g outside (0,1)", cfg->p);
    QTRY(check_geom(cfg->geom_chain, cfg->L));
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    Geo g = make_geo(cfg->geom_chain, cfg->L);
    bool wide = cfg->L > 16;
    size_t wbytes = wide ? 8 : 4;
    size_t nbytes = (size_t)chains * g.nsites;
    QTRY(c->qm_in.ensure(nbytes));
    QTRY(c->packed.ensure((size_t)chains * g.nw * wbytes));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemcpyAsync(c->qm_in.p, qm_in, nbytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, (const uint8_t *)c->qm_in.p, chains, g, c->packed.p));
    else QTRY(pack_lattices<uint32_t>(c, (const uint8_t *)c->qm_in.p, chains, g, c->packed.p));
    ChainParams p;
    memset(&p, 0, sizeof(p));
    p.g = g;
    p.lat = c->packed.p;
    p.chains = chains;
    p.iters = iters;
    make_thr(cfg->p, cfg->pow_kind, p.thr);
    p.seed = cfg->seed;
    p.offset = cfg->stream_offset;
    p.counters = (unsigned long long *)c->counters.p;
    int k = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) ? 3 : 5;
    size_t n_u = (size_t)chains * iters * (k + 1);
    if (u) {
        QTRY(c->replay_a.ensure(n_u * sizeof(double) + 8));
        CUDA_OK(cudaMemcpyAsync(c->replay_a.p, u, n_u * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        p.u = (const double *)c->replay_a.p;
        size_t trace = (size_t)chains * iters;
        size_t need = (dE ? trace : 0) + (accepted ? trace : 0) + (traj ? trace * g.nsites : 0);
        QTRY(c->out_u32.ensure(need + 16));
        uint8_t *base = (uint8_t *)c->out_u32.p;
        if (dE) { p.dE = (int8_t *)base; base += trace; }
        if (accepted) { p.acc = base; base += trace; }
        if (traj) p.traj = base;
    }
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    if (u) { if (wide) QTRY((launch_chain_geom<uint64_t, true>(c, p))); else QTRY((launch_chain_geom<uint32_t, true>(c, p))); }
    else { if (wide) QTRY((launch_chain_geom<uint64_t, false>(c, p))); else QTRY((launch_chain_geom<uint32_t, false>(c, p))); }
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    int T = 256;
    int64_t n_words = chains * g.nw;
    if (wide) unpack_kernel<uint64_t><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>((uint64_t *)c->packed.p, (uint8_t *)c->qm_in.p, n_words, g.L);
    else unpack_kernel<uint32_t><<<(unsigned)((n_words + T - 1) / T), T, 0, c->stream>>>((uint32_t *)c->packed.p, (uint8_t *)c->qm_in.p, n_words, g.L);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(qm_out, c->qm_in.p, nbytes, cudaMemcpyDeviceToHost, c->stream));
    size_t trace = (size_t)chains * iters;
    if (p.dE) CUDA_OK(cudaMemcpyAsync(dE, p.dE, trace, cudaMemcpyDeviceToHost, c->stream));
    if (p.acc) CUDA_OK(cudaMemcpyAsync(accepted, p.acc, trace, cudaMemcpyDeviceToHost, c->stream));
    if (p.traj) CUDA_OK(cudaMemcpyAsync(traj, p.traj, trace * g.nsites, cudaMemcpyDeviceToHost, c->stream));
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

extern "C" int qecmc_chain_update(qecmc_ctx *c, const qecmc_chain_cfg *cfg, uint8_t *qm, int64_t chains, int64_t iters,
                                  qecmc_stats *stats)
{
    return chain_common(c, cfg, qm, qm, nullptr, chains, iters, nullptr, nullptr, nullptr, stats);
}

extern "C" int qecmc_replay_chain(qecmc_ctx *c, const qecmc_chain_cfg *cfg, const uint8_t *qm0, const double *u,
                                  int64_t chains, int64_t iters, uint8_t *qm_final, int8_t *dE, uint8_t *accepted,
                                  uint8_t *traj)
{
    if (!u) return set_err(QECMC_ERR_ARG, "replay needs the uniform draws u");
    return chain_common(c, cfg, qm0, qm_final, u, chains, iters, dE, accepted, traj, nullptr);
}


// ---- tables of the table-driven STDC kernel (qecmc_stdc_fast.cuh) ----
static void make_philox_keys(uint64_t seed, PhiloxKeys &k)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { k.k0[r] = k0; k.k1[r] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

// thr[(v==3)*256 + f], dE[...]: weight change of flipping the four gathered fields f by Pauli v
static int build_fast_lut(qecmc_ctx *c, const Thr &t, FastTables &ft)
{
    if (c->lut.p && c->lut_valid && memcmp(c->lut_thr, t.u32, sizeof(t.u32)) == 0) {   // same sampling rate as the last call
        ft.thr = (const uint32_t *)c->lut.p;
        ft.dE = (const int8_t *)((const char *)c->lut.p + 512 * 4);
        return 0;
    }
    std::vector<unsigned char> host(512 * 4 + 512);
    uint32_t *thr = (uint32_t *)host.data();
    int8_t *dE = (int8_t *)(host.data() + 512 * 4);
    for (int vs = 0; vs < 2; vs++)
        for (int f = 0; f < 256; f++) {
            int v = vs ? 3 : 1, d = 0;
            for (int i = 0; i < 4; i++) {
                int q = (f >> (2 * i)) & 3;
                d += (q == 0) - (q == v);
            }
            thr[vs * 256 + f] = t.u32[d + QECMC_THR_OFF];
            dE[vs * 256 + f] = (int8_t)d;
        }
    QTRY(c->lut.ensure(host.size()));
    CUDA_OK(cudaMemcpyAsync(c->lut.p, host.data(), host.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    ft.thr = (const uint32_t *)c->lut.p;
    ft.dE = (const int8_t *)((const char *)c->lut.p + 512 * 4);
    memcpy(c->lut_thr, t.u32, sizeof(t.u32));
    c->lut_valid = true;
    return 0;
}

// interleaved copies of the packed-lattice kernel's hot tables (tests: qecmc_debug_set "packed" = 2 / 4 / 8)
static int pk_copies(const qecmc_ctx *c) { return c->dbg_packed == 2 || c->dbg_packed == 8 ? c->dbg_packed : 4; }

// packed-lattice kernel in bucket-log mode: whole tables per CTA beside the tables and the cursors (0: none fits)
static int pk_tables_per_cta(const Geo &g, int rep, int droplets, int nbc, size_t room)
{
    const PkGeo q = pk_geo(g);
    const PkLayout lay = pk_layout(g.nstab, rep);
    const size_t per_table = (size_t)q.nwp * 4 * droplets + (size_t)nbc * 4;
    if (lay.tile + 16 + per_table > room) return 0;
    int tpc = (int)((room - lay.tile - 16) / per_table);
    if (tpc * droplets > 1024) tpc = 1024 / droplets;
    return tpc;
}

// 17 <= L <= 24, native draws, no early stop: the packed-lattice kernel (qecmc_stdc_pk.cuh), one CTA per SM with as many
// chains as fit beside the tables
template <int GEOM, int MODE, int REP>
static int launch_stdc_pk_rep(qecmc_ctx *c, StdcParams &p, const FastTables &ft, const PhiloxKeys &keys)
{
    const PkGeo q = pk_geo(p.gchain);
    const PkLayout lay = pk_layout(p.gchain.nstab, REP);
    const size_t per_chain = (size_t)q.nwp * 4;
    const size_t room = c->prop.sharedMemPerBlockOptin;
    if (lay.tile + per_chain * 64 > room) return set_err(QECMC_ERR_UNSUPPORTED, "internal: packed tables do not fit");
    int T = (int)((room - lay.tile) / per_chain);
    if (T > 1024) T = 1024;
    T &= ~31;
    const int64_t sms = c->prop.multiProcessorCount;
    size_t smem;
    if (p.insert_mode == 6) {
        // a CTA holds whole tables (their bucket-log cursors live behind the tile): T = tables * droplets
        int tpc = pk_tables_per_cta(p.gchain, REP, p.droplets, p.nbc, room);
        if (tpc < 1 || p.n_chains % p.droplets != 0) return set_err(QECMC_ERR_UNSUPPORTED, "internal: bucket-log mode misconfigured");
        c->plan_round_chains = (int64_t)tpc * p.droplets * sms;
        // the launch takes ceil(tables / (tpc * SMs)) rounds of CTAs over the SMs: spread the tables evenly over that many
        // rounds (smaller CTAs) instead of leaving most SMs idle behind a short last round -- a small batch included
        const int64_t tabs = p.n_chains / p.droplets;
        const int64_t rounds = (tabs + (int64_t)tpc * sms - 1) / ((int64_t)tpc * sms);
        const int t2 = (int)((tabs + rounds * sms - 1) / (rounds * sms));
        if (t2 < tpc) tpc = t2 < 1 ? 1 : t2;
        T = tpc * p.droplets;
        p.tables_per_cta = tpc;
        smem = lay.tile + ((per_chain * T + 15) & ~(size_t)15) + (size_t)tpc * p.nbc * 4;
    } else {
        c->plan_round_chains = (int64_t)T * sms;
        // the same balancing over whole rounds of CTAs, in warps
        const int64_t rounds = (p.n_chains + (int64_t)T * sms - 1) / ((int64_t)T * sms);
        int t2 = (int)(((p.n_chains + rounds * sms - 1) / (rounds * sms) + 31) & ~(int64_t)31);
        if (t2 < 64) t2 = 64;
        if (t2 < T) T = t2;
        smem = lay.tile + per_chain * T;
    }
    CUDA_OK(cudaFuncSetAttribute(stdc_pk_kernel<GEOM, MODE, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((p.n_chains + T - 1) / T);
    stdc_pk_kernel<GEOM, MODE, REP><<<grid, T, smem, c->stream>>>(p, ft, keys);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <int GEOM, typename W, bool REPLAY, int MODE>
static int launch_stdc_fast(qecmc_ctx *c, StdcParams &p)
{
    FastTables ft;
    QTRY((build_stab_desc<GEOM, W>(c, p.gchain, &ft.desc)));
    QTRY(build_fast_lut(c, p.thr, ft));
    PhiloxKeys keys;
    make_philox_keys(p.seed, keys);
    if constexpr (sizeof(W) == 8 && !REPLAY) {
        // tests (qecmc_debug_set "packed"): 0 keeps these sizes on the 64-bit row-word kernel; 1 / 2 / 4 pick the number of
        // interleaved copies of the hot tables
        const bool pk_mode = MODE == MODE_MEAN || p.insert_mode == 4 || p.insert_mode == 2 || p.insert_mode == 6;
        if (p.conv_mult == 0.0 && pk_supported(p.gchain) && p.gchain.L == p.gcode.L && c->dbg_packed != 0 && pk_mode) {
            const int rep = pk_copies(c);
            if (rep == 2) return launch_stdc_pk_rep<GEOM, MODE, 2>(c, p, ft, keys);
            if (rep == 8) return launch_stdc_pk_rep<GEOM, MODE, 8>(c, p, ft, keys);
            return launch_stdc_pk_rep<GEOM, MODE, 4>(c, p, ft, keys);
        }
    }
    int T = 0, nb = 0;
    size_t per_chain = (size_t)p.gchain.nw * sizeof(W);
    // static part: LUTs (512 * 5 + 72 B) and, for 32-bit row words, expanded descriptor (2 x 16 B) + fingerprint arrays of 512 entries;
    // wider lattices keep descriptors and fingerprints behind the tile in dynamic shared memory
    const bool static_tab = sizeof(W) == 4;
    if (static_tab && p.gchain.nstab > QECMC_FAST_STATIC_NSTAB) return set_err(QECMC_ERR_UNSUPPORTED, "internal: %d stabilizers exceed the static tables", p.gchain.nstab);
    const bool conv = REPLAY || p.conv_mult != 0.0;
    // 32-bit row words: the table struct opens the dynamic shared memory (8 interleaved copies of the two hot tables in
    // the one-CTA-per-SM variant), the tile follows; wider lattices keep plain tables behind the tile
    const size_t tabs_bytes = !static_tab ? 0 : conv ? ((sizeof(FastTabs<1>) + 15) & ~(size_t)15) : ((sizeof(FastTabs<8>) + 15) & ~(size_t)15);
    size_t stat = 512 * 5 + 128;
    size_t dyn_fixed = static_tab ? tabs_bytes : (size_t)p.gchain.nstab * 16 + 16;
    // (64-bit row words: ONE CTA of 608-800 threads per SM instead of two of 256 was measured -- the per-SM rate rose 6 % with
    // 56 % more threads, and CTAs of odd sizes quantise badly over the SMs; descriptors with ready byte offsets changed nothing)
    QTRY(pick_threads(per_chain, stat + dyn_fixed + 256, c->prop, &T, &nb, static_tab && !conv, conv ? 56 : 64));
    c->plan_round_chains = (int64_t)T * nb * c->prop.multiProcessorCount;
    // a small batch is spread over the SMs rather than packed into a few large CTAs
    while (T > 64 && (p.n_chains + T - 1) / T < c->prop.multiProcessorCount) T /= 2;
    size_t smem = ((per_chain * T + 15) & ~(size_t)15) + dyn_fixed;
    if (p.insert_mode == 6) {   // per-CTA cursors into the bucket logs, behind the tile
        // a CTA holds whole tables: T = (tables per CTA) * droplets, not necessarily a multiple of 32; the last CTA of the
        // wave may hold fewer tables (its spare threads exit before the first barrier)
        int tpc = T / p.droplets;
        if (tpc < 1) tpc = 1;
        T = tpc * p.droplets;
        c->plan_round_chains = (int64_t)(1024 / p.droplets) * p.droplets * c->prop.multiProcessorCount;
        if (conv || !static_tab || T > 1024 || p.n_chains % p.droplets != 0) return set_err(QECMC_ERR_UNSUPPORTED, "internal: bucket-log mode misconfigured");
        smem = ((per_chain * T + 15) & ~(size_t)15) + dyn_fixed;
        p.tables_per_cta = tpc;
        smem += (size_t)tpc * p.nbc * 4;
    }
    unsigned grid = (unsigned)((p.n_chains + T - 1) / T);
    if (conv) {
        CUDA_OK(cudaFuncSetAttribute(stdc_fast_kernel<GEOM, W, REPLAY, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stdc_fast_kernel<GEOM, W, REPLAY, MODE, true><<<grid, T, smem, c->stream>>>(p, ft, keys);
    } else if (p.insert_mode == 6) {
        if constexpr (sizeof(W) == 4 && MODE != MODE_MEAN) {
            CUDA_OK(cudaFuncSetAttribute(stdc_fast_kernel<GEOM, W, false, MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            stdc_fast_kernel<GEOM, W, false, MODE, false, true><<<grid, T, smem, c->stream>>>(p, ft, keys);
        } else {
            return set_err(QECMC_ERR_UNSUPPORTED, "internal: bucket-log mode misconfigured");
        }
    } else {
        CUDA_OK(cudaFuncSetAttribute(stdc_fast_kernel<GEOM, W, false, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stdc_fast_kernel<GEOM, W, false, MODE, false><<<grid, T, smem, c->stream>>>(p, ft, keys);
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------ STDC ------------------------------
template <int GEOM, typename W, bool REPLAY, int MODE>
static int launch_stdc(qecmc_ctx *c, StdcParams &p)
{
    int T = 0, nb = 0;
    size_t per_chain = (size_t)p.gchain.nw * sizeof(W);
    size_t fixed = (size_t)p.gchain.nstab * 8 + 16;
    QTRY(pick_threads(per_chain, fixed + 256, c->prop, &T, &nb));
    size_t smem = ((per_chain * T + 15) & ~(size_t)15) + fixed;
    CUDA_OK(cudaFuncSetAttribute(stdc_kernel<GEOM, W, REPLAY, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((p.n_chains + T - 1) / T);
    stdc_kernel<GEOM, W, REPLAY, MODE><<<grid, T, smem, c->stream>>>(p);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename W, bool REPLAY, int MODE> static int launch_stdc_geom(qecmc_ctx *c, StdcParams &p)
{
    switch (p.gchain.geom) {
    case TORIC: QTRY((build_stab_hash<TORIC, W>(c, p.gchain, (uint64_t **)&p.stab_hash))); return launch_stdc_fast<TORIC, W, REPLAY, MODE>(c, p);
    case PLANAR: QTRY((build_stab_hash<PLANAR, W>(c, p.gchain, (uint64_t **)&p.stab_hash))); return launch_stdc_fast<PLANAR, W, REPLAY, MODE>(c, p);
    case ROTATED: QTRY((build_stab_hash<ROTATED, W>(c, p.gchain, (uint64_t **)&p.stab_hash))); return launch_stdc<ROTATED, W, REPLAY, MODE>(c, p);
    default: QTRY((build_stab_hash<XZZX, W>(c, p.gchain, (uint64_t **)&p.stab_hash))); return launch_stdc<XZZX, W, REPLAY, MODE>(c, p);
    }
}

template <int MODE> static int launch_stdc_mode(qecmc_ctx *c, StdcParams &p, bool wide, bool replay)
{
    if (replay) return wide ? launch_stdc_geom<uint64_t, true, MODE>(c, p) : launch_stdc_geom<uint32_t, true, MODE>(c, p);
    return wide ? launch_stdc_geom<uint64_t, false, MODE>(c, p) : launch_stdc_geom<uint32_t, false, MODE>(c, p);
}


// Shared driver of STDC / STRC / single_temp: all three run `droplets` chains per (syndrome, class) with a sample
// every `iters` Metropolis steps and differ in what a sample feeds.  Device pointers throughout.
struct StdcOut {
    double *eqdistr = nullptr;           // [S][n_eq]  percent (STDC, STRC) or mean length (single_temp)
    uint32_t *N_hist = nullptr;          // [S][n_eq][nsites+1]  optional
    unsigned long long *m_hist = nullptr;// [S][n_eq][nsites+1]  STRC, optional
    int32_t *short_info = nullptr;       // [S][n_eq][4]         STRC, optional
};

static int stdc_run_once(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, int mode, const uint8_t *d_qm, int64_t S, const StdcOut &out,
                         qecmc_stats *stats, bool allow_bucket_logs);

// wave sizes come from a cached free-memory figure: if an allocation fails after all, size once more from a fresh query
static int stdc_run(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, int mode, const uint8_t *d_qm, int64_t S, const StdcOut &out,
                    qecmc_stats *stats, bool allow_bucket_logs = true)
{
    if (!c) return set_err(QECMC_ERR_ARG, "NULL argument");
    return with_fresh_memory_retry(c, [&] { return stdc_run_once(c, cfg, mode, d_qm, S, out, stats, allow_bucket_logs); });
}

static int stdc_run_once(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, int mode, const uint8_t *d_qm, int64_t S, const StdcOut &out,
                         qecmc_stats *stats, bool allow_bucket_logs)
{
    if (!c || !cfg || !d_qm || !out.eqdistr) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    QTRY(check_geom(cfg->geom_code, cfg->L));
    QTRY(check_geom(cfg->geom_chain, cfg->L));
    Geo gcode = make_geo(cfg->geom_code, cfg->L), gchain = make_geo(cfg->geom_chain, cfg->L);
    if (gcode.layers != gchain.layers) return set_err(QECMC_ERR_ARG, "geom_code and geom_chain have different lattice shapes");
    if (cfg->droplets <= 0 || cfg->steps <= 0 || cfg->iters <= 0) return set_err(QECMC_ERR_ARG, "droplets, steps, iters must be > 0");
    if (!(cfg->p_error > 0 && cfg->p_error < 1) || !(cfg->p_sampling > 0 && cfg->p_sampling < 1))
        return set_err(QECMC_ERR_ARG, "p_error / p_sampling outside (0,1)");
    if ((uint64_t)cfg->steps * (uint64_t)cfg->iters >= (1ull << 32)) return set_err(QECMC_ERR_UNSUPPORTED, "steps * iters must be < 2^32");
    if (cfg->conv_mult < 0.0) return set_err(QECMC_ERR_ARG, "conv_mult must be >= 0");
    if (cfg->randomize && cfg->geom_code != TORIC && cfg->geom_code != PLANAR)
        return set_err(QECMC_ERR_ARG, "apply_stabilizers_uniform exists only for toric/planar codes");
    if (cfg->randomize && cfg->u_nb && !cfg->u_np) return set_err(QECMC_ERR_ARG, "replay with randomize needs u_np");
    if (mode == MODE_MEAN && cfg->steps < 2) return set_err(QECMC_ERR_ARG, "single_temp needs max_iters >= 2");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    // tests (qecmc_debug_set "force_wide"): 64-bit row words for L <= 16 too, to compare the two word widths on one problem
    const bool wide = cfg->L > 16 || c->dbg_force_wide;
    const size_t wbytes = wide ? 8 : 4;
    const int n_eq = gcode.neq;
    const int nh = gcode.nsites + 1;

    // Distinct-chain accounting.  Default: per-chain key logs counted afterwards by log_dedupe_kernel (streams through
    // HBM).  The open-addressing tables in HBM remain for the early stop (it needs "new or not" at once) and for key
    // counts beyond what the dedupe kernel's bucket fan-out covers.
    const int forced_mode = c->dbg_insert_mode;   // tests (qecmc_debug_set "insert_mode"); -1: the driver chooses
    const uint64_t max_keys = (uint64_t)cfg->droplets * (uint64_t)cfg->steps;
    const bool fits_dedupe = max_keys <= (uint64_t)QECMC_DD_MAX_BUCKETS * QECMC_DD_BUCKET_TARGET && (uint64_t)cfg->steps < (1ull << 32);
    // conv_mult != 0 (early stop, decoders.py:257-263): "new" means new to the DROPLET, so every chain probes a set of
    // its own (synchronously: the rule needs the answer at once) and logs only the keys new to it; the union over the
    // droplets of a class is then the same dedupe as without early stop.
    const bool conv = mode != MODE_MEAN && cfg->conv_mult != 0.0;
    const bool conv_logs = conv && fits_dedupe && forced_mode != 0;
    if (conv && !conv_logs && cfg->droplets != 1)
        return set_err(QECMC_ERR_UNSUPPORTED, "conv_mult != 0 with droplets > 1 needs droplets * steps <= %llu",
                       (unsigned long long)QECMC_DD_MAX_BUCKETS * QECMC_DD_BUCKET_TARGET);
    const bool use_logs = conv_logs || (mode != MODE_MEAN && !conv && (forced_mode < 0 || forced_mode == 4) && fits_dedupe);
    const int64_t log_cap = (cfg->steps + 1) & ~(int64_t)1;
    // Insert mode 6: the table-driven kernel (toric / planar, 32-bit row words, native draws, no early stop) splits the
    // keys of a table into coarse bucket logs as it produces them, when a table's chains fit one CTA and
    // the per-CTA cursors fit beside the tile; the reduction is then one pass (bucket_dedupe_kernel).
    // nbc coarse buckets per table, sized for ~6000 logged keys each when 30 % of the samples log one (a bucket that
    // outgrows the dedupe kernel's shared-memory set sends the call back to per-chain logs); the cursors of a CTA's tables
    // take (T / droplets) * nbc * 4 bytes beside tile and tables
    int nbc = 1;
    while (nbc < QECMC_NBC_MAX && (uint64_t)nbc * 20000 < max_keys) nbc <<= 1;
    const bool fast_u32 = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) && !wide;
    // 17 <= L <= 24: the packed-lattice kernel (qecmc_stdc_pk.cuh) writes bucket logs as well
    const bool fast_pk = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) && wide && pk_supported(gchain) && gchain.L == gcode.L &&
                         c->dbg_packed != 0;
    const int pk_rep = pk_copies(c);
    bool use_blogs = allow_bucket_logs && use_logs && !conv && !cfg->u_nb && (fast_u32 || fast_pk) && (forced_mode < 0 || forced_mode == 6) &&
                     cfg->droplets <= 1024;
    int64_t blog_tpc = 1024 / (cfg->droplets > 0 ? cfg->droplets : 1);   // tables per full CTA (wave rounding)
    if (use_blogs && fast_pk) {
        blog_tpc = pk_tables_per_cta(gchain, pk_rep, cfg->droplets, nbc, c->prop.sharedMemPerBlockOptin);
        if (blog_tpc < 1) use_blogs = false;
    } else if (use_blogs) {
        // the largest CTA the launch will use holds 1024 / droplets tables: its cursors must fit next to tile and tables
        // (256: the kernel's static shared memory)
        const size_t tpc_max = 1024 / cfg->droplets;
        const size_t used = (size_t)gchain.nw * 4 * tpc_max * cfg->droplets + 16 + ((sizeof(FastTabs<8>) + 15) & ~(size_t)15) + 256;
        // cursors that do not fit: fewer, larger buckets, as long as a bucket stays within the dedupe kernel's set when 35 % of
        // the samples log a key (beyond that the call falls back to per-chain logs by itself)
        while (nbc > 1 && used + tpc_max * nbc * 4 > c->prop.sharedMemPerBlockOptin &&
               (double)max_keys / (nbc / 2) * 0.35 < 0.8 * QECMC_BD_SLOTS)
            nbc >>= 1;
        const size_t cursors = tpc_max * nbc * 4;
        if (used + cursors > c->prop.sharedMemPerBlockOptin) use_blogs = false;
    }
    uint64_t bcap = 0, ovf_cap = 0;
    int dd_grid = c->prop.multiProcessorCount;
    uint64_t cap = 0;
    int64_t per_syndrome = 0, wave = S;
    if (mode != MODE_MEAN) {
        int64_t budget = c->table_budget;
        if (!budget) {
            // cudaMemGetInfo costs milliseconds (~10 ms on a context holding tens of GB): asked again only after the
            // library allocated or freed something; skipped altogether when the caller fixed the budget
            size_t fr = 0;
            QTRY(free_device_bytes(c, &fr));
            budget = (int64_t)((double)(fr + c->tables.cap + c->dd_scratch.cap) * 0.85);
        }
        if (use_blogs) {
            bcap = ((max_keys + nbc - 1) / nbc + 64 + 1) & ~(uint64_t)1;
            ovf_cap = max_keys / 16 < 1024 ? 1024 : max_keys / 16;
            per_syndrome = (int64_t)n_eq * (int64_t)(nbc * bcap + ovf_cap) * 8;
            wave = budget / per_syndrome;
            if (wave < 1)
                return set_err(QECMC_ERR_NOMEM, "bucket logs need %lld bytes per syndrome, budget is %lld", (long long)per_syndrome, (long long)budget);
            c->plan_wave_cap = wave;
            if (wave > S) wave = S;
            if (wave < S) {
                // several waves: a wave whose CTAs come to a whole number of rounds over the SMs (one 1024-thread CTA per SM,
                // 1024 / droplets tables each) leaves no SMs idle behind a short last round
                const int64_t tpc = blog_tpc, sms = c->prop.multiProcessorCount;
                int64_t ctas = wave * n_eq / tpc;
                if (ctas >= sms) {
                    ctas -= ctas % sms;
                    const int64_t w2 = ctas * tpc / n_eq;
                    if (w2 >= 1) wave = w2;
                }
            }
            QTRY(c->log_counts.ensure((size_t)wave * n_eq * (nbc + 1) * sizeof(uint32_t)));
            QTRY(c->scratch.ensure(2 * sizeof(int)));
        } else if (use_logs) {
            if ((int64_t)S * n_eq < dd_grid) dd_grid = (int)(S * n_eq);
            const int64_t scratch_bytes = (int64_t)dd_grid * cfg->droplets * log_cap * 8;
            if (conv_logs) {   // one set per chain: capacity covers every sample distinct at load <= 0.8
                cap = next_pow2((uint64_t)cfg->steps + (uint64_t)cfg->steps / 4 + 1);
                if (cap < 1024) cap = 1024;
            }
            per_syndrome = (int64_t)n_eq * cfg->droplets * (log_cap + (int64_t)cap) * 8;
            wave = (budget - scratch_bytes) / per_syndrome;
            if (wave < 1)
                return set_err(QECMC_ERR_NOMEM, "key logs need %lld bytes per syndrome plus %lld bytes of scratch, budget is %lld",
                               (long long)per_syndrome, (long long)scratch_bytes, (long long)budget);
            c->plan_wave_cap = wave;
            if (wave > S) wave = S;
            QTRY(c->dd_scratch.ensure((size_t)scratch_bytes));
            QTRY(c->log_counts.ensure((size_t)wave * n_eq * cfg->droplets * sizeof(uint32_t)));
            QTRY(c->scratch.ensure(2 * sizeof(int)));
        } else {
            // capacity covers the worst case (every sample distinct) at load <= 0.8
            cap = next_pow2(max_keys + max_keys / 4 + 1);
            if (cap < 1024) cap = 1024;
            if (cap > (1ull << 32)) return set_err(QECMC_ERR_UNSUPPORTED, "more than 2^32 slots per distinct-chain table");
            per_syndrome = (int64_t)n_eq * (int64_t)cap * 8;
            wave = budget / per_syndrome;
            if (wave < 1)
                return set_err(QECMC_ERR_NOMEM, "distinct-chain tables need %lld bytes per syndrome, budget is %lld",
                               (long long)per_syndrome, (long long)budget);
            c->plan_wave_cap = wave;
            if (wave > S) wave = S;
        }
        QTRY(c->tables.ensure((size_t)wave * per_syndrome));
    }
    int64_t n_lat = cfg->per_class_inits ? S * n_eq : S;
    QTRY(c->packed.ensure((size_t)n_lat * gcode.nw * wbytes));
    QTRY(c->Z.ensure((size_t)S * n_eq * sizeof(double)));
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    const int64_t wave_chains = wave * n_eq * cfg->droplets;
    uint32_t *d_nh = out.N_hist;
    unsigned long long *d_mh = out.m_hist;
    if (mode == MODE_STRC) {
        if (!d_nh) { QTRY(c->nhist.ensure((size_t)S * n_eq * nh * sizeof(uint32_t))); d_nh = (uint32_t *)c->nhist.p; }
        if (!d_mh) { QTRY(c->mhist.ensure((size_t)S * n_eq * nh * sizeof(unsigned long long))); d_mh = (unsigned long long *)c->mhist.p; }
        QTRY(c->shorts.ensure((size_t)wave_chains * 2 * sizeof(int)));
    }
    if (mode == MODE_MEAN) QTRY(c->sums.ensure((size_t)wave_chains * sizeof(unsigned long long)));
    CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
    CUDA_OK(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    if (mode == MODE_STRC) CUDA_OK(cudaMemsetAsync(d_mh, 0, (size_t)S * n_eq * nh * sizeof(unsigned long long), c->stream));
    if (wide) QTRY(pack_lattices<uint64_t>(c, d_qm, n_lat, gcode, c->packed.p));
    else QTRY(pack_lattices<uint32_t>(c, d_qm, n_lat, gcode, c->packed.p));

    StdcParams p;
    memset(&p, 0, sizeof(p));
    p.gcode = gcode;
    p.gchain = gchain;
    p.per_class = cfg->per_class_inits;
    p.randomize = cfg->randomize;
    p.droplets = cfg->droplets;
    p.iters = cfg->iters;
    p.steps = cfg->steps;
    p.seed = cfg->seed;
    p.hash_seed = c->hash_seed;
    p.tables = (unsigned long long *)c->tables.p;
    p.cap_mask = cap ? cap - 1 : 0;
    make_thr(cfg->p_sampling, QECMC_POW_NUMBA, p.thr);
    p.u_nb = cfg->u_nb;
    p.u_np = cfg->u_np;
    p.counters = (unsigned long long *)c->counters.p;
    // diagnostic knob for roofline work (3 = chains without the distinct set: results are then meaningless)
    p.insert_mode = use_blogs ? 6 : conv_logs ? 5 : use_logs ? 4 : (forced_mode >= 0 && forced_mode != 4 && forced_mode != 6 ? forced_mode : 2);
    if (use_blogs) {   // [bucket logs of the wave | overflow logs of the wave]; counts: [tables][NBC] then [tables]
        p.blogs = (unsigned long long *)c->tables.p;
        p.bcap = (uint32_t)bcap;
        p.nbc = nbc;
        p.ovf = p.blogs + (size_t)wave * n_eq * nbc * bcap;
        p.ovf_cap = (uint32_t)ovf_cap;
        p.bcounts = (uint32_t *)c->log_counts.p;
        p.ovf_cnt = p.bcounts + (size_t)wave * n_eq * nbc;
        p.log_err = (int *)c->scratch.p + 1;
    }
    // conv_logs: [per-chain sets of the wave | per-chain logs of the wave]; logs only: the logs start the buffer
    const size_t chain_sets_bytes = conv_logs ? (size_t)wave * n_eq * cfg->droplets * (size_t)cap * 8 : 0;
    p.logs = (unsigned long long *)((char *)c->tables.p + chain_sets_bytes);
    p.log_counts = (uint32_t *)c->log_counts.p;
    p.log_cap = log_cap;
    p.max_length = 2 * cfg->L * cfg->L;  // decoders.py:747
    p.conv_mult = mode == MODE_MEAN ? 0.0 : cfg->conv_mult;
    p.steps_done = (unsigned long long *)c->counters.p + 4;
    p.short_out = (int *)c->shorts.p;
    p.sum_out = (unsigned long long *)c->sums.p;
    const double beta = -log((cfg->p_error / 3) / (1 - cfg->p_error));  // decoders.py:299
    const double beta_s = -log((cfg->p_sampling / 3) / (1 - cfg->p_sampling));

    float chain_ms = 0;
    int64_t waves = 0;
    for (int64_t s0 = 0; s0 < S; s0 += wave, waves++) {
        int64_t sw = S - s0 < wave ? S - s0 : wave;
        if (mode != MODE_MEAN && !use_logs) CUDA_OK(cudaMemsetAsync(c->tables.p, 0, (size_t)sw * per_syndrome, c->stream));
        if (conv_logs) CUDA_OK(cudaMemsetAsync(c->tables.p, 0, (size_t)sw * n_eq * cfg->droplets * (size_t)cap * 8, c->stream));
        if (use_blogs) {
            CUDA_OK(cudaMemsetAsync(p.ovf_cnt, 0, (size_t)wave * n_eq * sizeof(uint32_t), c->stream));
            if (s0 == 0) CUDA_OK(cudaMemsetAsync(p.log_err, 0, sizeof(int), c->stream));
        }
        p.lat0 = (const char *)c->packed.p + (size_t)(cfg->per_class_inits ? s0 * n_eq : s0) * gcode.nw * wbytes;
        p.n_chains = sw * n_eq * cfg->droplets;
        p.chain_offset = s0 * n_eq * cfg->droplets;
        p.m_hist = d_mh ? d_mh + (size_t)s0 * n_eq * nh : nullptr;
        CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
        if (mode == MODE_STDC) QTRY(launch_stdc_mode<MODE_STDC>(c, p, wide, cfg->u_nb != nullptr));
        else if (mode == MODE_STRC) QTRY(launch_stdc_mode<MODE_STRC>(c, p, wide, cfg->u_nb != nullptr));
        else QTRY(launch_stdc_mode<MODE_MEAN>(c, p, wide, cfg->u_nb != nullptr));
        CUDA_OK(cudaEventRecord(c->ev[3], c->stream));
        const int64_t tabs = sw * n_eq;
        if (use_blogs) {
            BucketDedupeParams bp;
            bp.nbc = nbc;
            bp.blogs = p.blogs;
            bp.bcounts = p.bcounts;
            bp.bcap = p.bcap;
            bp.ovf = p.ovf;
            bp.ovf_cnt = p.ovf_cnt;
            bp.ovf_cap = p.ovf_cap;
            bp.tabs = tabs;
            bp.nsites = gcode.nsites;
            bp.beta = beta;
            bp.Z = (double *)c->Z.p + s0 * n_eq;
            bp.N_hist = d_nh ? d_nh + (size_t)s0 * n_eq * nh : nullptr;
            bp.distinct = (unsigned long long *)c->counters.p + 3;
            bp.err = p.log_err;
            const size_t dsm = (size_t)QECMC_BD_SLOTS * 8 + 4096 * 4;
            CUDA_OK(cudaFuncSetAttribute(bucket_dedupe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
            const int64_t sms = c->prop.multiProcessorCount;
            bucket_dedupe_kernel<<<(unsigned)(tabs < sms ? tabs : sms), QECMC_BD_THREADS, dsm, c->stream>>>(bp);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        } else if (mode != MODE_MEAN && use_logs) {
            DedupeParams dp;
            dp.logs = p.logs;
            dp.log_counts = (const uint32_t *)c->log_counts.p;
            dp.log_cap = log_cap;
            dp.droplets = cfg->droplets;
            dp.tabs = tabs;
            dp.scratch = (unsigned long long *)c->dd_scratch.p;
            dp.scratch_cap = (int64_t)cfg->droplets * log_cap;
            dp.nsites = gcode.nsites;
            dp.beta = beta;
            dp.Z = (double *)c->Z.p + s0 * n_eq;
            dp.N_hist = d_nh ? d_nh + (size_t)s0 * n_eq * nh : nullptr;
            dp.distinct = (unsigned long long *)c->counters.p + 3;
            dp.err = (int *)c->scratch.p + 1;
            const size_t dsm = (size_t)QECMC_DD_HASH_SLOTS * QECMC_DD_GROUPS * 8 + 4096 * 4 + (size_t)QECMC_DD_MAX_BUCKETS * 8 + QECMC_DD_MAX_CNT * 4;
            CUDA_OK(cudaFuncSetAttribute(log_dedupe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
            if (s0 == 0) CUDA_OK(cudaMemsetAsync(dp.err, 0, sizeof(int), c->stream));   // sticky; read once after the last wave
            log_dedupe_kernel<<<(unsigned)(tabs < dd_grid ? tabs : dd_grid), QECMC_DD_THREADS, dsm, c->stream>>>(dp);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        } else if (mode != MODE_MEAN) {
            table_hist_kernel<<<(unsigned)tabs, 512, nh * sizeof(uint32_t), c->stream>>>(
                (const unsigned long long *)c->tables.p, cap, gcode.nsites, beta, (double *)c->Z.p + s0 * n_eq,
                d_nh ? d_nh + (size_t)s0 * n_eq * nh : nullptr, (unsigned long long *)c->counters.p + 3);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        }
        if (mode == MODE_STRC) {
            strc_finalize_kernel<<<(unsigned)((tabs + 127) / 128), 128, 0, c->stream>>>(
                d_nh + (size_t)s0 * n_eq * nh, d_mh + (size_t)s0 * n_eq * nh, (const int *)c->shorts.p, tabs, cfg->droplets,
                gcode.nsites, p.max_length, beta_s, beta_s - beta, (double *)c->Z.p + s0 * n_eq,
                out.short_info ? out.short_info + (size_t)s0 * n_eq * 4 : nullptr);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        }
        if (mode == MODE_MEAN) {
            mean_kernel<<<(unsigned)((tabs + 127) / 128), 128, 0, c->stream>>>((const unsigned long long *)c->sums.p,
                                                                               out.eqdistr + s0 * n_eq, tabs, cfg->steps);
            c->launches++;
            CUDA_OK(cudaGetLastError());
        }
        if (stats) {  // per-wave chain-kernel time (synchronises; waves are seconds long)
            CUDA_OK(cudaEventSynchronize(c->ev[3]));
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
            chain_ms += ms;
        }
    }
    if (mode != MODE_MEAN) {
        normalize_kernel<<<(unsigned)((S + 127) / 128), 128, 0, c->stream>>>((const double *)c->Z.p, out.eqdistr, S, n_eq);
        c->launches++;
        CUDA_OK(cudaGetLastError());
    }
    CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
    int derr = 0;
    if (use_logs) CUDA_OK(cudaMemcpyAsync(&derr, (int *)c->scratch.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (stats) {
        unsigned long long cnt[8] = {0};
        CUDA_OK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (derr && use_blogs) return stdc_run(c, cfg, mode, d_qm, S, out, stats, false);   // a bucket overflowed: redo with per-chain logs
        if (derr) return set_err(QECMC_ERR_UNSUPPORTED, "internal: key-log dedupe overflow (%d)", derr);
        memset(stats, 0, sizeof(*stats));
        stats->metropolis_steps = (int64_t)cnt[4];
        stats->accepted = (int64_t)cnt[0];
        stats->samples = (int64_t)cnt[1];
        stats->distinct = (int64_t)cnt[3];
        stats->table_slots = use_blogs ? -1 : (int64_t)cap;   // -1: bucket logs, 0: per-chain logs, else slots per HBM set
        stats->waves = waves;
        stats->kernel_launches = c->launches;
        stats->chain_kernel_ms = chain_ms;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        stats->total_ms = ms;
    } else if (use_logs) {
        CUDA_OK(cudaStreamSynchronize(c->stream));
        if (derr && use_blogs) return stdc_run(c, cfg, mode, d_qm, S, out, stats, false);
        if (derr) return set_err(QECMC_ERR_UNSUPPORTED, "internal: key-log dedupe overflow (%d)", derr);
    }
    return 0;
}

extern "C" int qecmc_stdc_dev(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, const uint8_t *d_qm, int64_t S, double *d_eqdistr,
                              uint32_t *d_N_hist, qecmc_stats *stats)
{
    StdcOut o;
    o.eqdistr = d_eqdistr;
    o.N_hist = d_N_hist;
    return stdc_run(c, cfg, MODE_STDC, d_qm, S, o, stats);
}

// host-buffer front end shared by qecmc_stdc / qecmc_strc / qecmc_single_temp
static int stdc_host(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, int mode, const uint8_t *qm, int64_t S, double *eqdistr,
                     uint32_t *N_hist, uint64_t *m_hist, int32_t *short_info, qecmc_stats *stats)
{
    if (!c || !cfg || !qm || !eqdistr) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    QTRY(check_geom(cfg->geom_code, cfg->L));
    CUDA_OK(cudaSetDevice(c->device));
    Geo g = make_geo(cfg->geom_code, cfg->L);
    int64_t n_lat = cfg->per_class_inits ? S * g.neq : S;
    size_t in_bytes = (size_t)n_lat * g.nsites;
    size_t hist_elems = (size_t)S * g.neq * (g.nsites + 1);
    QTRY(c->qm_in.ensure(in_bytes));
    QTRY(c->out_f64.ensure((size_t)S * g.neq * sizeof(double)));
    if (N_hist) QTRY(c->out_u32.ensure(hist_elems * sizeof(uint32_t)));
    if (m_hist) QTRY(c->out_u64.ensure(hist_elems * sizeof(uint64_t)));
    if (short_info) QTRY(c->out_i32.ensure((size_t)S * g.neq * 4 * sizeof(int32_t)));
    CUDA_OK(cudaMemcpyAsync(c->qm_in.p, qm, in_bytes, cudaMemcpyHostToDevice, c->stream));
    qecmc_stdc_cfg dcfg = *cfg;
    if (cfg->u_nb) {
        if (cfg->droplets <= 0 || cfg->steps <= 0 || cfg->iters <= 0) return set_err(QECMC_ERR_ARG, "droplets, steps, iters must be > 0");
        int k = (cfg->geom_chain == TORIC || cfg->geom_chain == PLANAR) ? 3 : 5;
        size_t chains = (size_t)S * g.neq * cfg->droplets;
        size_t nb = chains * (size_t)cfg->steps * cfg->iters * (k + 1) * sizeof(double);
        QTRY(c->replay_a.ensure(nb));
        CUDA_OK(cudaMemcpyAsync(c->replay_a.p, cfg->u_nb, nb, cudaMemcpyHostToDevice, c->stream));
        dcfg.u_nb = (const double *)c->replay_a.p;
        if (cfg->u_np) {
            size_t np_ = chains * 2 * (size_t)cfg->L * cfg->L * sizeof(double);
            QTRY(c->replay_b.ensure(np_));
            CUDA_OK(cudaMemcpyAsync(c->replay_b.p, cfg->u_np, np_, cudaMemcpyHostToDevice, c->stream));
            dcfg.u_np = (const double *)c->replay_b.p;
        }
    }
    StdcOut o;
    o.eqdistr = (double *)c->out_f64.p;
    o.N_hist = N_hist ? (uint32_t *)c->out_u32.p : nullptr;
    o.m_hist = m_hist ? (unsigned long long *)c->out_u64.p : nullptr;
    o.short_info = short_info ? (int32_t *)c->out_i32.p : nullptr;
    QTRY(stdc_run(c, &dcfg, mode, (const uint8_t *)c->qm_in.p, S, o, stats));
    CUDA_OK(cudaMemcpyAsync(eqdistr, c->out_f64.p, (size_t)S * g.neq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (N_hist) CUDA_OK(cudaMemcpyAsync(N_hist, c->out_u32.p, hist_elems * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (m_hist) CUDA_OK(cudaMemcpyAsync(m_hist, c->out_u64.p, hist_elems * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (short_info) CUDA_OK(cudaMemcpyAsync(short_info, c->out_i32.p, (size_t)S * g.neq * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qecmc_stdc(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                          uint32_t *N_hist, qecmc_stats *stats)
{
    return stdc_host(c, cfg, MODE_STDC, qm, S, eqdistr, N_hist, nullptr, nullptr, stats);
}

extern "C" int qecmc_strc(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *eqdistr,
                          uint64_t *m_hist, int32_t *short_info, qecmc_stats *stats)
{
    return stdc_host(c, cfg, MODE_STRC, qm, S, eqdistr, nullptr, m_hist, short_info, stats);
}

extern "C" int qecmc_single_temp(qecmc_ctx *c, const qecmc_stdc_cfg *cfg, const uint8_t *qm, int64_t S, double *mean_length,
                                 qecmc_stats *stats)
{
    if (cfg && cfg->droplets != 1) return set_err(QECMC_ERR_ARG, "single_temp runs one chain per class (droplets must be 1)");
    return stdc_host(c, cfg, MODE_MEAN, qm, S, mean_length, nullptr, nullptr, nullptr, stats);
}

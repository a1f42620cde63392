// qecmc_kernels.cuh -- CUDA kernels (sm_100a) for the Metropolis-chain decoders.
//
// One thread owns one chain.  The chain's 2-bit-packed lattice lives in shared memory
// ([row word][thread], bank-conflict free); proposals come from per-thread Philox4x32-10
// (native) or from the reference's own uniform draws (replay, bit-exact); the weight
// change is a popcount difference on the touched row words; accepted moves update an
// incremental length and a GF(2)-linear fingerprint; every `iters` steps the chain offers
// its fingerprint to the (syndrome, class) distinct-chain set in HBM.
#pragma once
#include "qecmc_device.cuh"

namespace qecmc {

struct Thr {
    uint32_t u32[QECMC_THR_N];  // native: accept iff philox_word <= u32[dE + 4]
    double d[QECMC_THR_N];      // replay: accept iff u < d[dE + 4] (the reference's own doubles)
};

// ------------------------------ pack / unpack ------------------------------
template <typename W>
__global__ void pack_kernel(const uint8_t *__restrict__ qm, W *__restrict__ out, int64_t n_words, int L, int *bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const uint8_t *row = qm + i * L;
    W w = 0;
    int b = 0;
    for (int c = 0; c < L; c++) {
        uint8_t v = row[c];
        b |= v > 3;
        w |= (W)((W)(v & 3) << (2 * c));
    }
    out[i] = w;
    if (b) *bad = 1;
}

template <typename W>
__global__ void unpack_kernel(const W *__restrict__ in, uint8_t *__restrict__ qm, int64_t n_words, int L)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    W w = in[i];
    for (int c = 0; c < L; c++) qm[i * L + c] = (uint8_t)((w >> (2 * c)) & 3);
}

template <int GEOM, typename W> __global__ void stab_hash_kernel(Geo g, uint64_t seed, uint64_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.nstab) out[i] = stab_hash<GEOM, W>(g, i, seed);
}

// apply stabilizer (row, col, op) of a runtime geometry; cold paths only (rain)
template <typename W, typename A> __device__ int apply_rco_rt(const Geo &g, A &a, int row, int col, int op)
{
    Upd<W> u;
    switch (g.geom) {
    case TORIC: decode<TORIC, W>(g, row, col, op, u); return lat_apply<TORIC, W>(a, u);
    case PLANAR: decode<PLANAR, W>(g, row, col, op, u); return lat_apply<PLANAR, W>(a, u);
    case ROTATED: decode<ROTATED, W>(g, row, col, op, u); return lat_apply<ROTATED, W>(a, u);
    default: decode<XZZX, W>(g, row, col, op, u); return lat_apply<XZZX, W>(a, u);
    }
}
template <typename W, typename A> __device__ void to_class_rt(const Geo &g, A &a, int eq)
{
    switch (g.geom) {
    case TORIC: lat_to_class<TORIC, W>(g, a, eq); break;
    case PLANAR: lat_to_class<PLANAR, W>(g, a, eq); break;
    case ROTATED: lat_to_class<ROTATED, W>(g, a, eq); break;
    default: lat_to_class<XZZX, W>(g, a, eq); break;
    }
}
__device__ inline bool rain_legal_rt(const Geo &g, int o, int r, int c)
{
    return g.geom == PLANAR ? rain_legal<PLANAR>(g, o, r, c) : true;
}

// One Metropolis proposal on the thread's lattice.  Returns dE; commits when accepted.
template <int GEOM, typename W>
__device__ __forceinline__ bool metropolis(const Geo &g, SmemLat<W> &lat, int row, int col, int op, int &dE_out,
                                           const uint32_t *thr_u32, uint32_t r_acc, const double *thr_d, double u_acc,
                                           bool replay)
{
    constexpr int NU = NumUpd<GEOM>::value;
    Upd<W> u;
    decode<GEOM, W>(g, row, col, op, u);
    W nv[NU];
    int dE = 0;
#pragma unroll
    for (int i = 0; i < NU; i++) {
        W o = lat.get(u.w[i]);
        nv[i] = (W)(o ^ u.m[i]);
        dE += weight<W>(nv[i]) - weight<W>(o);
    }
    dE_out = dE;
    bool acc = replay ? (u_acc < thr_d[dE + QECMC_THR_OFF]) : (r_acc <= thr_u32[dE + QECMC_THR_OFF]);
    if (acc) {
#pragma unroll
        for (int i = 0; i < NU; i++) lat.set(u.w[i], nv[i]);
    }
    return acc;
}

// ------------------------------ plain chains ------------------------------
// Chain.update_chain_fast / _update_chain_fast (src/mcmc.py:45-46,152-160).
struct ChainParams {
    Geo g;
    void *lat;  // packed [chains][nw], in/out
    int64_t chains, iters;
    Thr thr;
    uint64_t seed, offset;
    const double *u;  // replay: [chains][iters][k+1]
    int8_t *dE;
    uint8_t *acc;
    uint8_t *traj;
    unsigned long long *counters;  // [0] accepted
};

template <int GEOM, typename W, bool REPLAY> __global__ void chain_kernel(ChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    W *tile = reinterpret_cast<W *>(smem);
    __shared__ uint32_t s_thr[QECMC_THR_N];
    __shared__ double s_thrd[QECMC_THR_N];
    if (tid < QECMC_THR_N) { s_thr[tid] = p.thr.u32[tid]; s_thrd[tid] = p.thr.d[tid]; }
    __syncthreads();
    int64_t ch = (int64_t)blockIdx.x * T + tid;
    if (ch >= p.chains) return;
    const Geo g = p.g;
    SmemLat<W> lat{tile + tid, T};
    W *gl = reinterpret_cast<W *>(p.lat) + ch * g.nw;
    for (int w = 0; w < g.nw; w++) lat.set(w, gl[w]);
    constexpr int K = NumDraws<GEOM>::value;
    unsigned long long nacc = 0;
    for (int64_t t = 0; t < p.iters; t++) {
        int row, col, op, dE;
        bool acc;
        if (REPLAY) {
            const double *u = p.u + (ch * p.iters + t) * (K + 1);
            propose_replay<GEOM>(g, u, row, col, op);
            acc = metropolis<GEOM, W>(g, lat, row, col, op, dE, s_thr, 0u, s_thrd, u[K], true);
        } else {
            uint64_t tg = p.offset + (uint64_t)t;
            uint64_t call = tg >> 1;
            uint4 r = philox4x32_10((uint32_t)call, (uint32_t)(call >> 32), (uint32_t)ch, (uint32_t)((uint64_t)ch >> 32),
                                    (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
            uint32_t ra = (tg & 1) ? r.z : r.x, rb = (tg & 1) ? r.w : r.y;
            int idx = (int)__umulhi(ra, (uint32_t)g.nstab);
            idx_to_rco<GEOM>(g, idx, row, col, op);
            acc = metropolis<GEOM, W>(g, lat, row, col, op, dE, s_thr, rb, s_thrd, 0.0, false);
        }
        nacc += acc;
        if (p.dE) p.dE[ch * p.iters + t] = (int8_t)dE;
        if (p.acc) p.acc[ch * p.iters + t] = (uint8_t)acc;
        if (p.traj) {
            uint8_t *o = p.traj + (ch * p.iters + t) * (int64_t)g.nsites;
            for (int w = 0; w < g.nw; w++) unpack_row<W>(lat.get(w), o + w * g.L, g.L);
        }
    }
    for (int w = 0; w < g.nw; w++) gl[w] = lat.get(w);
    if (p.counters && nacc) atomicAdd(p.counters, nacc);
}

// ------------------------------ STDC ------------------------------
// STDC_droplet (decoders.py:236-265) for every (syndrome, class, droplet) of a wave.
struct StdcParams {
    Geo gcode, gchain;
    const void *lat0;  // packed [S_wave][nw] or [S_wave][n_eq][nw]
    int per_class, randomize, droplets, iters;
    int64_t steps;
    int64_t n_chains;      // chains in this wave
    int64_t chain_offset;  // global index of the wave's first chain (RNG stream / replay arrays)
    uint64_t seed, hash_seed;
    unsigned long long *tables;  // [S_wave * n_eq][cap]
    uint64_t cap_mask;
    const uint64_t *stab_hash;  // [gchain.nstab]
    Thr thr;
    const double *u_nb, *u_np;
    unsigned long long *counters;  // [0] accepted [1] offered [2] inserted
    int insert_mode;               // 4 per-chain key logs + log_dedupe_kernel (default when applicable); 5 early stop: a set per
                                   // chain (probed at once) + a log of the keys new to the chain + dedupe; 6 bucket logs (below);
                                   // 2 prefetch + deferred probe of the HBM set; 0 synchronous probe; 3 no inserts (diagnostics)
    // insert_mode 6 (table-driven kernel, no early stop, a table's chains inside one CTA): keys go straight into
    // per-(table, coarse bucket) logs -- the first split of the dedupe is done where the key is produced
    int nbc;                       // coarse buckets per table: a power of two <= QECMC_NBC_MAX, sized for ~6000 keys each
    unsigned long long *blogs;     // [tables][nbc][bcap]
    uint32_t *bcounts;             // [tables][nbc] keys per bucket log (capped at bcap)
    uint32_t bcap;                 // slots per bucket log
    unsigned long long *ovf;       // [tables][ovf_cap] keys whose bucket log was full
    uint32_t *ovf_cnt;             // [tables]
    uint32_t ovf_cap;
    int tables_per_cta;            // blockDim.x / droplets
    int *log_err;                  // set to 3 when an overflow log overflows (the host then redoes the wave with mode 4)
    unsigned long long *logs;      // insert_mode 4: [n_chains][log_cap] offered keys, in order
    uint32_t *log_counts;          // [n_chains]
    int64_t log_cap;
    // STRC (decoders.py:745-832): visits per length m(n), per droplet shortest / next-shortest visited length
    unsigned long long *m_hist;    // [S_wave * n_eq][nsites + 1]
    int *short_out;                // [n_chains][2]
    int max_length;                // 2 * L * L (decoders.py:747)
    // single_temp (decoders.py:108-135): sum of the chain length over the first steps-1 samples
    unsigned long long *sum_out;   // [n_chains]
    // early stop of decoders.py:257-263 / :795 (droplets == 1 only): a chain ends once no new chain of the shortest
    // length has turned up for (conv_mult - 1) x as many samples as it took to find the last one
    double conv_mult;
    unsigned long long *steps_done;  // counters: Metropolis steps actually taken
};

// conv_mult bookkeeping of one chain; is_new = the sample was a chain this droplet had not seen before
struct ConvStop {
    int shortest;
    double stop;
    uint32_t sample;
    bool fin;
    __device__ __forceinline__ void init(const StdcParams &p) { shortest = p.max_length; stop = (double)p.steps; sample = 0; fin = false; }
    __device__ __forceinline__ void after_sample(const StdcParams &p, bool is_new, int n)
    {
        if (p.conv_mult != 0.0) {
            if (is_new && n <= shortest) { shortest = n; stop = (double)sample * p.conv_mult; }
            if ((double)sample >= stop && (uint64_t)sample * 100ull >= (uint64_t)p.steps) fin = true;
        }
        sample++;
    }
};

// ConvStopT<false>: conv_mult == 0 is known on the host, nothing to track
template <bool CONV> struct ConvStopT : ConvStop {
    __device__ __forceinline__ uint32_t samples() const { return sample; }
};
template <> struct ConvStopT<false> {
    static constexpr bool fin = false;
    __device__ __forceinline__ void init(const StdcParams &) {}
    __device__ __forceinline__ void after_sample(const StdcParams &, bool, int) {}
    __device__ __forceinline__ uint32_t samples() const { return 0; }
};

#define QECMC_NBC_MAX 128   // most coarse buckets per (syndrome, class) table in insert mode 6

enum { MODE_STDC = 0, MODE_STRC = 1, MODE_MEAN = 2 };

// Per-sample bookkeeping beyond the distinct set.
template <int MODE> struct SampleAcct {
    int sh, nsh, run_n;
    uint32_t run_cnt;
    unsigned long long *mh;
    unsigned long long sum;
    int64_t mean_left;
    __device__ __forceinline__ void init(const StdcParams &p, int64_t tab)
    {
        sh = nsh = p.max_length;
        run_n = -1;
        run_cnt = 0;
        mh = MODE == MODE_STRC ? p.m_hist + (uint64_t)tab * (p.gcode.nsites + 1) : nullptr;
        sum = 0;
        mean_left = p.steps - 1;
    }
    __device__ __forceinline__ void sample(int n)
    {
        if (MODE == MODE_STRC) {
            if (n != run_n) {  // run-length aggregated m(n) += 1
                if (run_cnt) atomicAdd(mh + run_n, (unsigned long long)run_cnt);
                run_n = n;
                run_cnt = 0;
            }
            run_cnt++;
            if (n < sh) { nsh = sh; sh = n; }
            else if (n > sh && n < nsh) nsh = n;
        } else if (MODE == MODE_MEAN) {
            if (mean_left > 0) { sum += (unsigned long long)n; mean_left--; }
        }
    }
    __device__ __forceinline__ void finish(const StdcParams &p, int64_t local)
    {
        if (MODE == MODE_STRC) {
            if (run_cnt) atomicAdd(mh + run_n, (unsigned long long)run_cnt);
            p.short_out[2 * local] = sh;
            p.short_out[2 * local + 1] = nsh;
        } else if (MODE == MODE_MEAN) {
            p.sum_out[local] = sum;
        }
    }
};

template <int GEOM, typename W, bool REPLAY, int MODE>
__global__ void __launch_bounds__(256) stdc_kernel(StdcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const Geo g = p.gchain;
    W *tile = reinterpret_cast<W *>(smem);
    uint64_t *s_hs = reinterpret_cast<uint64_t *>(smem + (((size_t)g.nw * T * sizeof(W) + 15) & ~(size_t)15));
    __shared__ uint32_t s_thr[QECMC_THR_N];
    __shared__ double s_thrd[QECMC_THR_N];
    for (int i = tid; i < g.nstab; i += T) s_hs[i] = p.stab_hash[i];
    if (tid < QECMC_THR_N) { s_thr[tid] = p.thr.u32[tid]; s_thrd[tid] = p.thr.d[tid]; }
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * T + tid;
    if (local >= p.n_chains) return;
    const int64_t gchain = p.chain_offset + local;
    const int n_eq = p.gcode.neq;
    const int64_t tab = local / p.droplets;  // (syndrome, class) within the wave
    const int eq = (int)(tab % n_eq);
    const int64_t sw = tab / n_eq;
    SmemLat<W> lat{tile + tid, T};
    {
        const W *src = reinterpret_cast<const W *>(p.lat0) + (p.per_class ? tab : sw) * g.nw;
        for (int w = 0; w < g.nw; w++) lat.set(w, src[w]);
    }
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const uint32_t cl = (uint32_t)gchain, chh = (uint32_t)((uint64_t)gchain >> 32);

    // class initialisation: Toric_code.to_class / apply_logical(class ^ eq) (decoders.py:286-288, 556-560)
    if (!p.per_class) to_class_rt<W>(p.gcode, lat, eq);
    // rain: apply_stabilizers_uniform(p = 0.5) with the code's own geometry (decoders.py:245-246)
    if (p.randomize) {
        const int L = g.L;
        uint4 r = make_uint4(0, 0, 0, 0);
        int i = 0;
        for (int o = 0; o < 2; o++)
            for (int rr = 0; rr < L; rr++)
                for (int c = 0; c < L; c++, i++) {
                    bool hit;
                    if (REPLAY) {
                        hit = p.u_np[gchain * (int64_t)(2 * L * L) + i] < 0.5;
                    } else {
                        if ((i & 127) == 0) r = philox4x32_10((uint32_t)(i >> 7), 0x80000000u, cl, chh, k0, k1);
                        int wi = (i >> 5) & 3;
                        uint32_t word = wi == 0 ? r.x : wi == 1 ? r.y : wi == 2 ? r.z : r.w;
                        hit = (word >> (i & 31)) & 1;
                    }
                    if (hit && rain_legal_rt(p.gcode, o, rr, c)) apply_rco_rt<W>(p.gcode, lat, rr, c, o == 0 ? 3 : 1);
                }
    }
    int n = lat_weight<W>(g, lat);
    uint64_t h = lat_hash<W>(g, lat, p.hash_seed);
    const bool log_mode = MODE != MODE_MEAN && p.insert_mode == 4 && p.conv_mult == 0.0;
    const bool conv_log = MODE != MODE_MEAN && p.insert_mode == 5;   // early stop: a set per chain + a log of the keys new to it
    unsigned long long *table = log_mode ? p.logs + (uint64_t)local * (uint64_t)p.log_cap
                                         : p.tables + (uint64_t)(conv_log ? local : tab) * (p.cap_mask + 1);
    unsigned long long *clog = p.logs + (uint64_t)local * (uint64_t)p.log_cap;
    uint32_t nlog = 0;
    SampleAcct<MODE> acct;
    acct.init(p, tab);
    ConvStop cs;
    cs.init(p);

    unsigned long long nacc = 0, noff = 0, nins = 0;
    bool dirty = true;  // the first sample is always new to the chain
    int left = p.iters;
    const uint64_t tsteps = (uint64_t)p.steps * (uint64_t)p.iters;

#define QECMC_AFTER_STEP()                                                  \
    if (acc) { n += dE; h ^= s_hs[idx]; dirty = true; nacc++; }             \
    if (--left == 0) {                                                      \
        left = p.iters;                                                     \
        acct.sample(n);                                                     \
        bool is_new = false;                                                \
        if (MODE != MODE_MEAN && dirty) {                                   \
            if (log_mode) table[noff++] = make_key(h, n);                   \
            else { noff++; is_new = table_insert(table, p.cap_mask, make_key(h, n)); nins += is_new;  \
                   if (conv_log && is_new) clog[nlog++] = make_key(h, n); }        \
        }                                                                   \
        dirty = false;                                                      \
        cs.after_sample(p, is_new, n);                                      \
    }

    if (REPLAY) {
        constexpr int K = NumDraws<GEOM>::value;
        const double *u = p.u_nb + (uint64_t)gchain * tsteps * (K + 1);
        for (uint64_t t = 0; t < tsteps && !cs.fin; t++, u += K + 1) {
            int row, col, op, dE;
            propose_replay<GEOM>(g, u, row, col, op);
            int idx = rco_to_idx<GEOM>(g, row, col, op);
            bool acc = metropolis<GEOM, W>(g, lat, row, col, op, dE, s_thr, 0u, s_thrd, u[K], true);
            QECMC_AFTER_STEP()
        }
    } else {
        uint32_t c0 = 0, c1 = 0;
        for (uint64_t t = 0; t < tsteps && !cs.fin; t += 2) {
            uint4 r = philox4x32_10(c0, c1, cl, chh, k0, k1);
            if (++c0 == 0) ++c1;
            {
                int row, col, op, dE;
                int idx = (int)__umulhi(r.x, (uint32_t)g.nstab);
                idx_to_rco<GEOM>(g, idx, row, col, op);
                bool acc = metropolis<GEOM, W>(g, lat, row, col, op, dE, s_thr, r.y, s_thrd, 0.0, false);
                QECMC_AFTER_STEP()
            }
            if (t + 1 < tsteps && !cs.fin) {
                int row, col, op, dE;
                int idx = (int)__umulhi(r.z, (uint32_t)g.nstab);
                idx_to_rco<GEOM>(g, idx, row, col, op);
                bool acc = metropolis<GEOM, W>(g, lat, row, col, op, dE, s_thr, r.w, s_thrd, 0.0, false);
                QECMC_AFTER_STEP()
            }
        }
    }
#undef QECMC_AFTER_STEP
    if (log_mode) p.log_counts[local] = (uint32_t)noff;
    if (conv_log) p.log_counts[local] = nlog;
    acct.finish(p, local);
    // statistics: three atomics per chain at the very end (negligible)
    atomicAdd(p.counters + 0, nacc);
    atomicAdd(p.counters + 1, noff);
    atomicAdd(p.counters + 2, nins);
    atomicAdd(p.steps_done, (unsigned long long)cs.sample * (unsigned long long)p.iters);
}

// One block per (syndrome, class) table: N(n) histogram from the length field of the
// keys, then Z_E = sum_n N(n) exp(-beta n) (decoders.py:317-318).
static __global__ void table_hist_kernel(const unsigned long long *__restrict__ tables, uint64_t cap, int nsites, double beta,
                                  double *__restrict__ Z, uint32_t *__restrict__ N_hist, unsigned long long *distinct)
{
    extern __shared__ uint32_t s_hist[];
    const unsigned long long *tab = tables + (uint64_t)blockIdx.x * cap;
    for (int i = threadIdx.x; i <= nsites; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint64_t i = threadIdx.x; i < cap; i += blockDim.x) {
        unsigned long long k = tab[i];
        if (k) atomicAdd(&s_hist[(int)(k & QECMC_LEN_MASK)], 1u);
    }
    __syncthreads();
    if (N_hist)
        for (int i = threadIdx.x; i <= nsites; i += blockDim.x) N_hist[(uint64_t)blockIdx.x * (nsites + 1) + i] = s_hist[i];
    if (threadIdx.x == 0) {
        double z = 0;
        unsigned long long cnt = 0;
        for (int n = 0; n <= nsites; n++)
            if (s_hist[n]) { z += (double)s_hist[n] * exp(-beta * (double)n); cnt += s_hist[n]; }
        if (Z) Z[blockIdx.x] = z;
        if (distinct) atomicAdd(distinct, cnt);
    }
}

// PTDC with the early stop and several droplets: one CTA per (syndrome, class) pours the droplets' own sets into the
// class's set (zeroed by the host); N(n) then comes from table_hist_kernel as usual.
static __global__ void table_union_kernel(const unsigned long long *__restrict__ ltabs /* [tabs][droplets][cap] */, uint64_t cap,
                                          int droplets, unsigned long long *__restrict__ utabs /* [tabs][ucap] */, uint64_t ucap,
                                          int aux_bits)
{
    const unsigned long long *src = ltabs + (uint64_t)blockIdx.x * droplets * cap;
    unsigned long long *dst = utabs + (uint64_t)blockIdx.x * ucap;
    for (uint64_t i = threadIdx.x; i < (uint64_t)droplets * cap; i += blockDim.x) {
        const unsigned long long k = src[i];
        if (!k) continue;
        uint64_t slot = (k >> aux_bits) & (ucap - 1);
        while (true) {
            unsigned long long prev = atomicCAS(dst + slot, 0ull, k);
            if (prev == 0ull || prev == k) break;
            slot = (slot + 1) & (ucap - 1);
        }
    }
}

// STRC per (syndrome, class): the order-dependent droplet merge of decoders.py:882-928 (the union of the
// droplets' distinct shortest / next-shortest sets has N(shortest) / N(next_shortest) members),
// then Z_E = sum_l m(l) exp(-beta_s * shortest + d_beta * l) * mean_fraction (decoders.py:931-946).
static __global__ void strc_finalize_kernel(const uint32_t *__restrict__ N_hist, const unsigned long long *__restrict__ m_hist,
                                     const int *__restrict__ short_in, int64_t tabs, int droplets, int nsites,
                                     int max_length, double beta_s, double d_beta, double *__restrict__ Z,
                                     int *__restrict__ short_info)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tabs) return;
    const int *sh = short_in + t * droplets * 2;
    int shortest = max_length, next = max_length;
    for (int d = 0; d < droplets; d++) {
        if (sh[2 * d] < shortest) { next = shortest; shortest = sh[2 * d]; }
        if (sh[2 * d + 1] < next) next = sh[2 * d + 1];
    }
    const uint32_t *N = N_hist + t * (nsites + 1);
    const unsigned long long *m = m_hist + t * (nsites + 1);
    double n0 = shortest <= nsites ? (double)N[shortest] : 0.0, n1 = next <= nsites ? (double)N[next] : 0.0;
    double frac = n0 / (shortest <= nsites ? (double)m[shortest] : 0.0);
    if (next != max_length) {
        double nf = n1 / (double)m[next];
        frac = 0.5 * (frac + nf * exp(-beta_s * (double)(next - shortest)));
    }
    double z = 0;
    for (int l = 0; l <= nsites; l++)
        if (m[l]) z += (double)m[l] * exp(-beta_s * (double)shortest + d_beta * (double)l);
    Z[t] = z * frac;
    if (short_info) {
        short_info[4 * t] = shortest;
        short_info[4 * t + 1] = next;
        short_info[4 * t + 2] = (int)n0;
        short_info[4 * t + 3] = (int)n1;
    }
}

// single_temp: mean chain length over the first steps-1 samples (decoders.py:128-133)
static __global__ void mean_kernel(const unsigned long long *__restrict__ sum, double *__restrict__ out, int64_t n, int64_t steps)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)sum[i] / (double)(steps - 1);
}

// eqdistr = Z / sum(Z) * 100 per syndrome (decoders.py:322)
static __global__ void normalize_kernel(const double *__restrict__ Z, double *__restrict__ out, int64_t S, int n_eq)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double tot = 0;
    for (int e = 0; e < n_eq; e++) tot += Z[s * n_eq + e];
    for (int e = 0; e < n_eq; e++) out[s * n_eq + e] = Z[s * n_eq + e] / tot * 100;
}

}  // namespace qecmc

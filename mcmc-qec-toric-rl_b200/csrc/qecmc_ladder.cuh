// qecmc_ladder.cuh -- parallel-tempering ladders on the GPU.
//
// Reference: Ladder / Ladder_alpha / Ladder_biased (.step, .r_flip) and Chain / Chain_alpha /
// Chain_biased (.update_chain) in src/mcmc.py:19-103, src/mcmc_alpha.py:27-137,
// src/mcmc_biased.py:21-124; PTEQ's bookkeeping in decoders.py:25-105 and
// decoders_biasednoise.py:28-237; PTDC_droplet in decoders.py:138-165; STDC_droplet_alpha in
// decoders.py:510-534.
//
// Mapping.  One ladder = one group of G = 2^k >= Nc consecutive lanes of a warp, lane <-> one
// replica.  A replica's 2-bit-packed lattice never moves: it stays in the lane's column of the
// shared-memory tile ([row word][thread], conflict free).  A replica swap exchanges the two lanes'
// RUNG INDICES (and, for the alpha ladder, the rung-owned n_eff, which the reference does not swap),
// so the sequential top-to-bottom swap sweep is ballots + shuffles, never a lattice copy and never a
// host round trip.  Replay mode consumes the reference's own numba (NB) and CPython (PY) uniform
// streams in the reference's order and is bit-exact; native mode draws from
// per-lane Philox4x32-10 streams.
#pragma once
#include <type_traits>
#include "qecmc_device.cuh"
#include "qecmc_stdc_fast.cuh"   // gather_fields: the two-layer codes' table-driven step

namespace qecmc {

enum { LK_DEPOL = 0, LK_ALPHA = 1, LK_BIASED = 2 };
#define QECMC_PW_K 31   // swap-sweep power table: 2 * 31 + 1 exponents per rung pair
enum { ACCT_NONE = 0, ACCT_PTEQ = 1, ACCT_DC = 2, ACCT_RC = 3 };

struct LadderParams {
    Geo g;
    int kind, Nc, G, iters, acct;
    double p_logical;       // top rung: probability of proposing a logical operator
    int top_accept_all;     // kind 0: ladder[Nc-1] >= 0.75 (mcmc.py:30)
    const uint2 *desc2;     // toric / planar, depolarizing: stabilizer descriptors of build_stab_desc (qecmc_internal.h)
    int serial_sweep;       // tests (qecmc_debug_set "serial_sweep"): native mode walks the swap sweep pair by pair like replay does
    int64_t n_ladders, ladder_offset, steps;
    const void *lat_in;     // packed [n_ladders][nw] (init_broadcast) or [n_ladders][Nc][nw]
    int init_broadcast;
    const int *flags_in;    // resume: [n_ladders][Nc] (optional; default: only the top rung is flagged)
    const int2 *neff_in;    // resume: [n_ladders][Nc] rung-owned (nz, nx+ny) (optional; default: from the rung's state)
    const long long *tops0_in;  // resume: [n_ladders] (optional)
    void *lat_out;          // packed [n_ladders][Nc][nw], rung order (optional)
    int *flags_out;         // [n_ladders][Nc] (optional)
    int2 *neff_out;         // [n_ladders][Nc] (nz, nx+ny) of the rung-owned n_eff (optional)
    long long *tops0_out;   // [n_ladders] (optional)
    void *snap_lat;         // tests: [n_ladders][steps][Nc][nw] after every Ladder.step
    int *snap_flags;        // [n_ladders][steps][Nc]
    long long *snap_tops0;  // [n_ladders][steps]
    const double *thr_d;    // [Nc][9]  kind 0: pow(factor_r, dE), dE = -4..4 (libm, as CPython float ** int)
    const uint32_t *thr_u;  // [Nc][9]  the same as 32-bit thresholds (native mode)
    const double *thr_top_d;// [8L+1]   kind 0 top rung: pow(factor_top, dE), dE = -4L..4L (a toric logical
                            //          acts on both layers: up to 2(2L-1) qubits)
    const double *diff;     // [Nc-1]   p_diff (kind 0, 2) or pz_tilde[i]/pz_tilde[i+1] (kind 1)
    const double *pw;       // [Nc-1][2 QECMC_PW_K + 1] diff^k, k = 0 .. 2 QECMC_PW_K (rung-major kernel; numba's square-and-multiply)
    double alpha;           // kind 1
    const double *wtab;     // kinds 1, 2: [Nc][4][nsites+1] = px^k, py^k, pz^k, q0^(L*L-k)
    uint64_t seed;
    PhiloxKeys keys;            // native: the ten Philox round keys of `seed`
    const double *u_nb, *u_py;  // replay: [n_ladders][n_nb], [n_ladders][n_py]
    int n_nb, n_py;
    // PTEQ
    int SEQ, TOPS, tops_burn, use_conv;
    double eps;
    uint32_t *hist;         // [n_ladders][steps]  bottom-rung n_err (or nz | nxy << 16) after burn-in
    long long *eq_counts;   // [n_ladders][n_eq]
    long long *info;        // [n_ladders][4] = steps used, since_burn, tops0, converged
    uint8_t *percent;       // [n_ladders][n_eq]
    int cls_delta[8];       // raw class change of logical (layer, op)
    // distinct-chain accounting (PTDC, EWD)
    unsigned long long *tables;
    uint64_t cap_mask;
    int droplets, aux_bits;
    const uint64_t *stab_hash;
    uint64_t hash_seed;
    const uint64_t *log_hash;      // [2 layers][X, Z][32 positions] fingerprints of the logical strings (when hashes are tracked
                                   // on ladders whose top rung proposes logical operators)
    // PTEQ_alpha_with_shortest (decoders_biasednoise.py:114-146); tables = one set per ladder, cap_mask slots
    double conv_mult;              // ACCT_DC: early stop of PTDC_droplet (decoders.py:156-161); tables are then per ladder
    int track_shortest;
    double *short_v;               // [n_ladders][n_eq] smallest recorded bottom-rung value per class (100000 = none)
    long long *short_n;            // [n_ladders][n_eq] samples at that value
    unsigned long long *rc_m_hist; // ACCT_RC: [n_ladders][Nc][nsites+1] visits per length and rung (PTRC's m(n))
    unsigned long long *counters;  // [0] accepted [1] offered
    int *status;                   // != 0: a replay stream ran dry
};

// raw class bits: XOR-linear in the lattice.  Equal to the class label except for XZZX, whose label
// is a relabelling of (x, z) (xzzx_model.py:476-486): raw = x + 2z <-> label by swapping 2 and 3.
__host__ __device__ inline int class_raw_to_label(int geom, int raw) { return (geom == XZZX && raw >= 2) ? 5 - raw : raw; }

struct NativeRng {
    uint32_t id_lo, id_hi, ctr, tag;   // tag = counter word 1: separate streams of one lane never overlap
    uint4 buf;
    int have;
    __device__ __forceinline__ void init(uint64_t id, uint32_t stream_tag)
    {
        id_lo = (uint32_t)id; id_hi = (uint32_t)(id >> 32);
        ctr = 0; have = 0; tag = stream_tag;
        buf = make_uint4(0, 0, 0, 0);
    }
    // k: the ten round keys, precomputed on the host (a kernel parameter: every key is a constant-bank operand)
    __device__ __forceinline__ uint32_t next32(const PhiloxKeys &k)
    {
        if (have == 0) { buf = philox4x32_10(ctr++, tag, id_lo, id_hi, k); have = 4; }
        uint32_t v = buf.x;
        buf.x = buf.y; buf.y = buf.z; buf.z = buf.w;
        have--;
        return v;
    }
    __device__ __forceinline__ double nb(const PhiloxKeys &k) { return (double)next32(k) * 2.3283064365386963e-10; }
    __device__ __forceinline__ double py(const PhiloxKeys &k) { return nb(k); }
    // (int)(nb() * m) for the uniform w / 2^32 is floor(w * m / 2^32): the same number without the round trip through double
    __device__ __forceinline__ int nb_int(const PhiloxKeys &k, int m) { return (int)__umulhi(next32(k), (uint32_t)m); }
};

struct ReplayRng {
    const double *nb_base, *py_base;
    int nbp, pyp, n_nb, n_py;
    int *status;
    __device__ __forceinline__ double nb(const PhiloxKeys &)
    {
        if (nbp >= n_nb) { *status = 1; nbp++; return 0.0; }
        return nb_base[nbp++];
    }
    __device__ __forceinline__ double py(const PhiloxKeys &)
    {
        if (pyp >= n_py) { *status = 2; pyp++; return 0.0; }
        return py_base[pyp++];
    }
    __device__ __forceinline__ int nb_int(const PhiloxKeys &k, int m) { return (int)(nb(k) * m); }
};

// Native mode, the top rung's own stream (Philox counter word 1 = 1 of the lane's id): its words are produced by the whole
// ladder at once -- lane j runs call number c0 + j into a shared-memory pool -- and the top lane reads them in order.  Word i
// of the stream is component i & 3 of call i >> 2, whoever computes it.
struct TopRng {
    uint32_t pos;          // words of this lane's stream consumed so far
    uint32_t end;          // stream position the pool reaches (this lane's pool only while it is the top lane)
    const uint32_t *pool;  // the word at stream position pos
    __device__ __forceinline__ uint32_t next32(const PhiloxKeys &) { pos++; return *pool++; }
    __device__ __forceinline__ double nb(const PhiloxKeys &k) { return (double)next32(k) * 2.3283064365386963e-10; }
    __device__ __forceinline__ double py(const PhiloxKeys &k) { return nb(k); }
    __device__ __forceinline__ int nb_int(const PhiloxKeys &k, int m) { return (int)__umulhi(next32(k), (uint32_t)m); }
};

// numba's float64 ** int64 (mcmc.py:149): square-and-multiply, reciprocal for negative exponents
__device__ __forceinline__ double numba_pow_dev(double a, int b)
{
    double r = 1.0;
    bool inv = b < 0;
    unsigned e = inv ? (unsigned)(-b) : (unsigned)b;
    while (e) {
        if (e & 1) r = __dmul_rn(r, a);
        e >>= 1;
        a = __dmul_rn(a, a);
    }
    return inv ? __ddiv_rn(1.0, r) : r;
}

// P(state) of mcmc_alpha.py:40 / mcmc_biased.py:31 from host-made pow tables (bit-equal to libm's)
__device__ __forceinline__ double chain_weight(const double *wt, int ns1, int nx, int ny, int nz)
{
    double a = __dmul_rn(wt[nx], wt[ns1 + ny]);
    a = __dmul_rn(a, wt[2 * ns1 + nz]);
    return __dmul_rn(a, wt[3 * ns1 + nx + ny + nz]);
}

template <int GEOM, typename RNG> struct LogicalDraw {
    static constexpr int nl = GEOM == TORIC ? 2 : 1;   // compile-time: the small arrays stay in registers
    int op[2], xp[2], zp[2];
    // _apply_random_logical draw order: toric_model.py:228-253 (both layer operators first),
    // planar_model.py:271-288, rotated_surface_model.py:331-346, xzzx_model.py:340-357
    template <typename R> __device__ __forceinline__ void draw(R &rng, int L, const PhiloxKeys &k)
    {
#pragma unroll
        for (int l = 0; l < nl; l++) op[l] = rng.nb_int(k, 4);
#pragma unroll
        for (int l = 0; l < nl; l++) {
            xp[l] = zp[l] = 0;
            if (op[l] == 1 || op[l] == 2) xp[l] = rng.nb_int(k, L);
            if (op[l] == 3 || op[l] == 2) zp[l] = rng.nb_int(k, L);
        }
    }
    // fingerprint change of the drawn operator: XOR of the string fingerprints (the fingerprint is GF(2)-linear)
    __device__ __forceinline__ uint64_t hash_delta(const uint64_t *lh) const
    {
        uint64_t d = 0;
#pragma unroll
        for (int l = 0; l < nl; l++) {
            const bool do_X = (GEOM == TORIC || GEOM == XZZX) ? (op[l] == 1 || op[l] == 2) : (op[l] == 1 || op[l] == 3);
            const bool do_Z = op[l] == 2 || op[l] == 3;
            const int px = GEOM == XZZX ? 0 : xp[l], pz = GEOM == XZZX ? 0 : zp[l];
            if (do_X) d ^= lh[(l * 2 + 0) * 32 + px];
            if (do_Z) d ^= lh[(l * 2 + 1) * 32 + pz];
        }
        return d;
    }
    template <typename W, typename A> __device__ __forceinline__ int apply(const Geo &g, A &lat) const
    {
        int d = 0;
        for (int l = 0; l < nl; l++) d += lat_apply_logical<GEOM, W>(g, lat, op[l], l, xp[l], zp[l]);
        return d;
    }
};

template <int GEOM, typename W, bool REPLAY, bool WEIGHTED>
__global__ void __launch_bounds__(128, 4) ladder_kernel(LadderParams p)
{
    typedef typename std::conditional<REPLAY, ReplayRng, NativeRng>::type RNG;
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const Geo g = p.g;
    W *tile = reinterpret_cast<W *>(smem);
    unsigned char *sp = smem + (((size_t)g.nw * T * sizeof(W) + 15) & ~(size_t)15);
    double *s_thrd = reinterpret_cast<double *>(sp);                       // [Nc][9]
    uint32_t *s_thru = reinterpret_cast<uint32_t *>(s_thrd + p.Nc * 9);    // [Nc][9]
    if (!WEIGHTED)
        for (int i = tid; i < p.Nc * 9; i += T) { s_thrd[i] = p.thr_d[i]; s_thru[i] = p.thr_u[i]; }
    // One-layer codes (rotated, XZZX): a stabilizer touches at most two adjacent qubits in each of two row words, so the
    // step is table driven.  Per stabilizer: the two word indices, the bit position of each word's lower touched field and
    // the Paulis applied to the four slots (a nibble per word); the weight changes (dx, dy, dz and their sum) come from a
    // LUT indexed by (Pauli pattern, the four touched fields).  Both tables are built here from decode<GEOM>().
    constexpr bool TABLE = GEOM == ROTATED || GEOM == XZZX;
    // Two-layer codes (toric, planar), depolarizing ladders: the descriptors of the STDC chain kernel (three row words, two
    // shifts, which slots exist) and a 512-entry LUT of the weight change by (Pauli, the four touched fields).
    constexpr bool TABLE2 = (GEOM == TORIC || GEOM == PLANAR) && !WEIGHTED;
    uint2 *s_ld = reinterpret_cast<uint2 *>(s_thru + p.Nc * 9 + ((p.Nc * 9) & 1));   // [nstab], 8-byte aligned
    uint16_t *s_ll = reinterpret_cast<uint16_t *>(s_ld + ((TABLE || TABLE2) ? g.nstab : 0));        // [patterns <= 16][256]
    int8_t *s_dE2 = reinterpret_cast<int8_t *>(s_ll);                                               // TABLE2: [512]
    __shared__ uint32_t s_patmask[8];
    // swap sweep: rung-ordered copies, per warp.  Between sweeps the same 2 KB hold the top rungs' random-word pools
    // (one Philox call, 16 bytes, per lane).
    __shared__ __align__(16) int s_sw_all[4 * 128];
    int *s_sw_w = s_sw_all + (threadIdx.x >> 5) * 128;   // this warp's 512 bytes: four arrays of 32, or 32 pool entries
    int *s_sw_lane = s_sw_w, *s_sw_a = s_sw_w + 32, *s_sw_b = s_sw_w + 64, *s_sw_rung = s_sw_w + 96;
    uint4 *s_pool = reinterpret_cast<uint4 *>(s_sw_w);
    __shared__ double s_sw_u_all[128];
    double *s_sw_u = s_sw_u_all + (threadIdx.x >> 5) * 32;
    // swap sweep: diff[i]^k for |k| <= QECMC_PW_K, made with the same square-and-multiply routine the sweep would call
    // (bit-identical), so a pair costs one table read instead of a multiply loop and, for k < 0, a division
    double *s_pw = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(s_ll) + (TABLE ? 16 * 256 * 2 : TABLE2 ? 512 : 0));
    const bool use_pw = p.kind != LK_ALPHA && p.Nc > 1;
    // depolarizing ladders only evaluate k = n_hi - n_lo >= 0 (a lighter upper replica swaps without a draw), so their
    // table covers 0 .. 2 QECMC_PW_K; biased ladders draw always and cover -QECMC_PW_K .. QECMC_PW_K
    const int pw_off = p.kind == LK_DEPOL ? 0 : QECMC_PW_K;
    if (use_pw)
        for (int e = tid; e < (p.Nc - 1) * (2 * QECMC_PW_K + 1); e += T)
            s_pw[e] = numba_pow_dev(p.diff[e / (2 * QECMC_PW_K + 1)], e % (2 * QECMC_PW_K + 1) - pw_off);
    if (TABLE2) {
        for (int i = tid; i < g.nstab; i += T) s_ld[i] = p.desc2[i];
        for (int e = tid; e < 512; e += T) {
            const int v = (e >> 8) ? 3 : 1;
            int d = 0;
            for (int sl = 0; sl < 4; sl++) {
                const int q = (e >> (2 * sl)) & 3, nq = q ^ v;   // a missing slot reads as Y: X and Z leave its weight alone
                d += (q == 0 && nq != 0) - (q != 0 && nq == 0);
            }
            s_dE2[e] = (int8_t)d;
        }
    }
    if (TABLE) {
        if (tid < 8) s_patmask[tid] = 0;
        __syncthreads();
        for (int i = tid; i < g.nstab; i += T) {
            int row, col, op;
            idx_to_rco<GEOM>(g, i, row, col, op);
            Upd<W> u;
            decode<GEOM, W>(g, row, col, op, u);
            uint32_t pos[2], nib[2];
            for (int k = 0; k < 2; k++) {
                int b = 0;
                while (b < 2 * g.L && !((u.m[k] >> b) & 3)) b += 2;
                pos[k] = u.m[k] ? (uint32_t)b : 0u;
                nib[k] = (uint32_t)(u.m[k] >> pos[k]) & 0xFu;
            }
            const uint32_t pat = nib[0] | (nib[1] << 4);
            s_ld[i] = make_uint2((uint32_t)u.w[0] | ((uint32_t)u.w[1] << 8) | (pos[0] << 16) | (pos[1] << 24), pat);
            atomicOr(&s_patmask[pat >> 5], 1u << (pat & 31));
        }
        __syncthreads();
        // pattern id = rank of the pattern among those present
        for (int i = tid; i < g.nstab; i += T) {
            const uint32_t pat = s_ld[i].y & 0xFFu;
            uint32_t id = __popc(s_patmask[pat >> 5] & ((1u << (pat & 31)) - 1u));
            for (uint32_t wq = 0; wq < (pat >> 5); wq++) id += __popc(s_patmask[wq]);
            s_ld[i].y = pat | (id << 8);
        }
        for (int e = tid; e < 16 * 256; e += T) {
            // the pattern with rank e >> 8
            int want = e >> 8, pat = -1;
            for (int wq = 0; wq < 8 && pat < 0; wq++) {
                uint32_t mbits = s_patmask[wq];
                const int c = __popc(mbits);
                if (want >= c) { want -= c; continue; }
                while (want--) mbits &= mbits - 1;
                pat = wq * 32 + __ffs(mbits) - 1;
            }
            uint32_t packed = 0;
            if (pat >= 0) {
                const int f = e & 255;
                int dx = 0, dy = 0, dz = 0;
                for (int sl = 0; sl < 4; sl++) {
                    const int v = (pat >> (2 * sl)) & 3, q = (f >> (2 * sl)) & 3;
                    if (v) {
                        const int nq = q ^ v;
                        dx += (nq == 1) - (q == 1);
                        dy += (nq == 2) - (q == 2);
                        dz += (nq == 3) - (q == 3);
                    }
                }
                packed = (uint32_t)(dx + 4) | ((uint32_t)(dy + 4) << 4) | ((uint32_t)(dz + 4) << 8) | ((uint32_t)(dx + dy + dz + 4) << 12);
            }
            s_ll[e] = (uint16_t)packed;
        }
    }
    __syncthreads();

    const int G = p.G, Nc = p.Nc, L = g.L;
    const int lane = tid & 31;
    const int gl = lane & (G - 1);          // lane within the ladder's group
    const int gbase = lane - gl;
    const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const int64_t ladder = ((int64_t)blockIdx.x * T + tid) / G;   // within this launch
    const bool valid = ladder < p.n_ladders && gl < Nc;
    const int64_t gladder = p.ladder_offset + ladder;
    constexpr int K = NumDraws<GEOM>::value;
    constexpr int NU = NumUpd<GEOM>::value;
    const int ns1 = g.nsites + 1;

    SmemLat<W> lat{tile + tid, T};
    int r = gl, flag = 0;
    int n = 0, nx = 0, ny = 0, nz = 0, cls = 0;
    int e_nz = 0, e_nxy = 0;  // rung-owned n_eff = e_nz + alpha * e_nxy (mcmc_alpha.py:22,56)
    uint64_t h = 0;
    bool dirty = true;
    if (valid) {
        const W *src = reinterpret_cast<const W *>(p.lat_in) + (p.init_broadcast ? ladder : ladder * Nc + gl) * g.nw;
        for (int w = 0; w < g.nw; w++) lat.set(w, src[w]);
        lat_count_xyz<W>(g, lat, nx, ny, nz);
        n = nx + ny + nz;
        e_nz = nz; e_nxy = nx + ny;
        cls = class_raw_to_label(GEOM, lat_class<GEOM, W>(g, lat));
        flag = p.flags_in ? p.flags_in[ladder * Nc + gl] : ((gl == Nc - 1) ? 1 : 0);
        if (p.neff_in) { int2 e = p.neff_in[ladder * Nc + gl]; e_nz = e.x; e_nxy = e.y; }
        if (p.acct >= ACCT_DC || p.track_shortest) h = lat_hash<W>(g, lat, p.hash_seed);
    } else {
        for (int w = 0; w < g.nw; w++) lat.set(w, (W)0);
    }
    RNG rng;
    int base_nb = 0, base_py = 0;  // replay: group-uniform stream positions
    if (REPLAY) {
        ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
        const int64_t lc = ladder < p.n_ladders ? ladder : 0;
        rr->nb_base = p.u_nb + lc * p.n_nb;
        rr->py_base = p.u_py + lc * p.n_py;
        rr->n_nb = p.n_nb; rr->n_py = p.n_py;
        rr->nbp = rr->pyp = 0;
        rr->status = p.status;
    } else {
        reinterpret_cast<NativeRng *>(&rng)->init((uint64_t)gladder * 32u + (uint64_t)gl, 0u);
    }
    // Native mode: what only the top rung draws (logical-move decision, the operator, its accept draw) comes from a
    // stream of its own.  The main stream then serves every lane exactly two words per iteration and one per sweep, so all
    // lanes refill together -- one Philox call per warp instead of one per out-of-step lane.  Replay keeps the single
    // stream (the reference's draw order).
    typedef typename std::conditional<REPLAY, ReplayRng, TopRng>::type TOP;
    TopRng top_native;
    top_native.pos = top_native.end = 0;
    top_native.pool = nullptr;
    TOP &rtop = *reinterpret_cast<TOP *>(REPLAY ? (void *)&rng : (void *)&top_native);
    constexpr int TOP_MAXW = 2 + 3 * LogicalDraw<GEOM, RNG>::nl;   // words one iteration of the top rung can take
    const bool top_logical = p.p_logical != 0.0;
    const bool track_hash = p.acct >= ACCT_DC || p.track_shortest;
    unsigned long long *table = nullptr;
    // PTDC: the class's set, shared by its droplets -- or, with the early stop ("new" = new to the droplet), the ladder's own
    if (p.acct == ACCT_DC && ladder < p.n_ladders)
        table = p.tables + (uint64_t)(p.conv_mult != 0.0 ? ladder : ladder / p.droplets) * (p.cap_mask + 1);
    int dc_shortest = 2 * L * L;          // group-uniform: shortest chain this droplet has seen, and when to stop
    double dc_stop = (double)p.steps;
    int last_r = -1;  // ACCT_RC: rung whose set saw this replica's current state

    // group-uniform bookkeeping (every lane of the group carries the same values)
    long long tops0 = (p.tops0_in && ladder < p.n_ladders) ? p.tops0_in[ladder] : 0, since_burn = 0, burn_in = 0, conv_start = 0, conv_streak = 0, steps_used = p.steps;
    long long S2a = 0, S2b = 0, S4a = 0, S4b = 0;  // window sums of the n_err history (a: n or nz, b: nx+ny)
    long long wA = 0, wB = 0, wC = 0, wl = 0;      // window bounds l/4, l/2, 3l/4 and history length l
    int converged = 0;
    bool done = !(ladder < p.n_ladders);
    uint32_t nacc = 0, noff = 0;
    uint32_t *hist = p.hist ? p.hist + (ladder < p.n_ladders ? ladder : 0) * p.steps : nullptr;
    long long *eqc = p.eq_counts ? p.eq_counts + (ladder < p.n_ladders ? ladder : 0) * g.neq : nullptr;

    for (long long step = 0; step < p.steps; step++) {
        if (__all_sync(0xFFFFFFFFu, done)) break;
        // ---------------- Ladder.update_ladder: every rung runs `iters` Metropolis steps ----------------
        const bool active = valid && !done;
        {
            const bool is_top = active && (r == Nc - 1) && top_logical;
            if (REPLAY && active) {
                ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
                rr->nbp = base_nb + r * K * p.iters;  // rungs below the top consume fixed amounts
                rr->pyp = base_py + r * p.iters;
            }
            const double *wt = WEIGHTED ? p.wtab + (size_t)(active ? r : 0) * 4 * ns1 : nullptr;
            double pb = 0.0;
            if (WEIGHTED && active) pb = chain_weight(wt, ns1, nx, ny, nz);  // frozen for the block
            for (int it = 0; it < p.iters; it++) {
                // A logical operator touches O(L) row words.  Only the top rung proposes one (with probability p_logical),
                // so instead of one lane working through the rows while the others wait, the WARP evaluates it: the owner
                // draws the operator and broadcasts it, lane j forms the new value of row word j (and j + 32) of the
                // owner's lattice, the weight change is a warp reduction, the owner decides, and on accept the lanes commit
                // their words.  Same draws in the same order, same arithmetic as _apply_random_logical + update_chain.
                // native: every active lane takes its two main-stream words here, whatever it goes on to do
                if (!REPLAY && top_logical) {
                    const bool need = is_top && (int)(top_native.end - top_native.pos) < TOP_MAXW;
                    if (__any_sync(0xFFFFFFFFu, need)) {
                        // every ladder of the warp refreshes its pool: G calls from the top lane's stream position on
                        const uint32_t bt = (__ballot_sync(0xFFFFFFFFu, valid && r == Nc - 1) >> gbase) & gmask;
                        const int lt = bt ? __ffs(bt) - 1 : 0;
                        const uint32_t tpos = __shfl_sync(0xFFFFFFFFu, top_native.pos, lt, G);
                        const uint64_t tid64 = (uint64_t)gladder * 32u + (uint64_t)lt;
                        const uint32_t c0 = tpos >> 2;
                        uint4 *mine = s_pool + gbase;
                        __syncwarp();
                        mine[gl] = philox4x32_10(c0 + (uint32_t)gl, 1u, (uint32_t)tid64, (uint32_t)(tid64 >> 32), p.keys);
                        if (gl == lt) {
                            top_native.pool = reinterpret_cast<const uint32_t *>(mine) + (tpos & 3u);
                            top_native.end = (c0 + (uint32_t)G) * 4u;
                        }
                        __syncwarp();
                    }
                }
                uint32_t w_idx = 0, w_acc = 0;
                if (!REPLAY && active) {
                    w_idx = reinterpret_cast<NativeRng *>(&rng)->next32(p.keys);
                    w_acc = reinterpret_cast<NativeRng *>(&rng)->next32(p.keys);
                }
                bool logical = false;
                LogicalDraw<GEOM, RNG> ld;
                ld.op[0] = ld.op[1] = ld.xp[0] = ld.xp[1] = ld.zp[0] = ld.zp[1] = 0;
                if (is_top) {
                    logical = rtop.py(p.keys) < p.p_logical;
                    if (logical) ld.draw(rtop, L, p.keys);
                }
                uint32_t lm = __ballot_sync(0xFFFFFFFFu, logical);
                while (lm) {
                    const int src = __ffs(lm) - 1;
                    lm &= lm - 1;
                    const int packed = ld.op[0] | (ld.op[1] << 2) | (ld.xp[0] << 4) | (ld.zp[0] << 9) | (ld.xp[1] << 14) | (ld.zp[1] << 19);
                    const int pk = __shfl_sync(0xFFFFFFFFu, packed, src);
                    const int o0 = pk & 3, o1 = (pk >> 2) & 3, x0 = (pk >> 4) & 31, z0 = (pk >> 9) & 31, x1 = (pk >> 14) & 31, z1 = (pk >> 19) & 31;
                    W *col = tile + ((tid & ~31) + src);   // the owner's column of the tile
                    W nv[2] = {0, 0};
                    bool touched[2] = {false, false};
                    int dE = 0, dx = 0, dy = 0, dz = 0;
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const int w = lane + 32 * k;
                        if (w < g.nw) {
                            const W m = logical_mask<GEOM, W>(g, w, o0, o1, x0, z0, x1, z1);
                            if (m) {
                                const W o = col[w * T];
                                nv[k] = (W)(o ^ m);
                                touched[k] = true;
                                if (WEIGHTED) {
                                    dx += popc(xmap(nv[k])) - popc(xmap(o));
                                    dy += popc(ymap(nv[k])) - popc(ymap(o));
                                    dz += popc(zmap(nv[k])) - popc(zmap(o));
                                } else {
                                    dE += weight<W>(nv[k]) - weight<W>(o);
                                }
                            }
                        }
                    }
                    if (WEIGHTED) {
                        dx = __reduce_add_sync(0xFFFFFFFFu, dx);
                        dy = __reduce_add_sync(0xFFFFFFFFu, dy);
                        dz = __reduce_add_sync(0xFFFFFFFFu, dz);
                    } else {
                        dE = __reduce_add_sync(0xFFFFFFFFu, dE);
                    }
                    bool acc = false;
                    if (lane == src) {
                        if (WEIGHTED) {
                            double pn = chain_weight(wt, ns1, nx + dx, ny + dy, nz + dz);
                            if (REPLAY) acc = rtop.py(p.keys) < __ddiv_rn(pn, pb);
                            else acc = rtop.py(p.keys) * pb < pn;
                        } else {
                            if (p.top_accept_all || dE <= 0) acc = true;
                            else acc = rtop.py(p.keys) < p.thr_top_d[dE + 4 * L];
                        }
                        if (acc) {
                            int dcls = 0;
#pragma unroll
                            for (int l = 0; l < ld.nl; l++) dcls ^= p.cls_delta[l * 4 + ld.op[l]];
                            if (WEIGHTED) { nx += dx; ny += dy; nz += dz; n = nx + ny + nz; e_nz = nz; e_nxy = nx + ny; }
                            else n += dE;
                            cls ^= dcls;
                            if (track_hash) h ^= ld.hash_delta(p.log_hash);
                            dirty = true;
                            nacc++;
                        }
                    }
                    acc = __shfl_sync(0xFFFFFFFFu, (int)acc, src) != 0;
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            if (touched[k]) col[(lane + 32 * k) * T] = nv[k];
                    }
                    __syncwarp();
                }
                if (!active || logical) {
                    // nothing more this iteration (idle lane, or the logical move was this rung's step)
                } else {
                    int row, col, op, idx = 0;
                    if (REPLAY) {
                        ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
                        double u[K];
                        for (int k = 0; k < K; k++) u[k] = rr->nb(p.keys);
                        propose_replay<GEOM>(g, u, row, col, op);
                        if (track_hash) idx = rco_to_idx<GEOM>(g, row, col, op);
                    } else {
                        idx = (int)__umulhi(w_idx, (uint32_t)g.nstab);
                        if (!TABLE && !TABLE2) idx_to_rco<GEOM>(g, idx, row, col, op);
                    }
                    Upd<W> u;
                    W nv[NU];
                    int dE = 0, dx = 0, dy = 0, dz = 0;
                    if (TABLE) {
                        if (REPLAY && !track_hash) idx = rco_to_idx<GEOM>(g, row, col, op);
                        const uint2 D = s_ld[idx];
                        u.w[0] = (int)(D.x & 0xFFu);
                        u.w[1] = (int)((D.x >> 8) & 0xFFu);
                        const uint32_t p0 = (D.x >> 16) & 0xFFu, p1 = D.x >> 24;
                        const W o0 = lat.get(u.w[0]), o1 = lat.get(u.w[1]);
                        const uint32_t f = ((uint32_t)(o0 >> p0) & 0xFu) | (((uint32_t)(o1 >> p1) & 0xFu) << 4);
                        const uint32_t pk = s_ll[((D.y >> 8) << 8) + f];
                        nv[0] = (W)(o0 ^ ((W)(D.y & 0xFu) << p0));
                        nv[1] = (W)(o1 ^ ((W)((D.y >> 4) & 0xFu) << p1));
                        if (WEIGHTED) { dx = (int)(pk & 15u) - 4; dy = (int)((pk >> 4) & 15u) - 4; dz = (int)((pk >> 8) & 15u) - 4; }
                        else dE = (int)(pk >> 12) - 4;
                    } else if (TABLE2) {
                        if (REPLAY && !track_hash) idx = rco_to_idx<GEOM>(g, row, col, op);
                        const uint2 D = s_ld[idx];
                        u.w[0] = (int)((D.x >> 8) & 0xFFu);
                        u.w[1] = (int)((D.x >> 16) & 0xFFu);
                        u.w[2] = (int)(D.x >> 24);
                        const W o0 = lat.get(u.w[0]), o1 = lat.get(u.w[1]), o2 = lat.get(u.w[2]);
                        const uint32_t f = gather_fields<W>(o0, o1, o2, D.x, D.y);
                        const uint32_t li = GEOM == TORIC ? ((f & 0xFFu) | (D.y >> 16)) : ((f & ((D.y >> 8) & 0xFFu)) | (D.y >> 16));
                        dE = (int)s_dE2[li];
                        const uint32_t sh = D.x & 63u, sh2 = D.y & 63u;
                        const W v = (D.y & 0x01000000u) ? (W)3 : (W)1;
                        W m0, m1, m2;
                        if (GEOM == TORIC) {
                            m1 = (W)(v << sh);
                            m2 = m1;
                            m0 = (W)(m1 | (W)(v << sh2));
                        } else {
                            const uint32_t fa = D.y >> 8;
                            m0 = (W)(((fa & 1u) ? (W)(v << sh) : (W)0) | ((fa & 4u) ? (W)(v << sh2) : (W)0));
                            m1 = (fa & 16u) ? (W)(v << sh) : (W)0;
                            m2 = (fa & 64u) ? (W)(v << sh) : (W)0;
                        }
                        nv[0] = (W)(o0 ^ m0);
                        nv[1] = (W)(o1 ^ m1);
                        nv[NU > 2 ? 2 : 0] = (W)(o2 ^ m2);
                    } else {
                        decode<GEOM, W>(g, row, col, op, u);
#pragma unroll
                        for (int i = 0; i < NU; i++) {
                            W o = lat.get(u.w[i]);
                            nv[i] = (W)(o ^ u.m[i]);
                            if (WEIGHTED) {
                                dx += popc(xmap(nv[i])) - popc(xmap(o));
                                dy += popc(ymap(nv[i])) - popc(ymap(o));
                                dz += popc(zmap(nv[i])) - popc(zmap(o));
                            } else {
                                dE += weight<W>(nv[i]) - weight<W>(o);
                            }
                        }
                    }
                    bool acc;
                    if (WEIGHTED) {
                        double pn = chain_weight(wt, ns1, nx + dx, ny + dy, nz + dz);
                        // replay divides like the reference (bit-exact decisions); native compares u * pb < pn and saves the division
                        if (REPLAY) acc = rng.py(p.keys) < __ddiv_rn(pn, pb);
                        else acc = (double)w_acc * 2.3283064365386963e-10 * pb < pn;
                    } else if (is_top) {
                        if (p.top_accept_all || dE <= 0) acc = true;
                        else acc = (REPLAY ? rng.py(p.keys) : (double)w_acc * 2.3283064365386963e-10) < p.thr_top_d[dE + 4 * L];
                    } else if (REPLAY) {
                        acc = rng.py(p.keys) < s_thrd[r * 9 + dE + QECMC_THR_OFF];
                    } else {
                        acc = w_acc <= s_thru[r * 9 + dE + QECMC_THR_OFF];
                    }
                    if (acc) {
#pragma unroll
                        for (int i = 0; i < NU; i++) lat.set(u.w[i], nv[i]);
                        if (WEIGHTED) { nx += dx; ny += dy; nz += dz; n = nx + ny + nz; e_nz = nz; e_nxy = nx + ny; }
                        else n += dE;
                        if (track_hash) h ^= p.stab_hash[idx];
                        dirty = true;
                        nacc++;
                    }
                }
            }
        }
        __syncwarp();
        // ---------------- swap sweep, top to bottom (mcmc.py:96-103) ----------------
        if (REPLAY) {
            // the top rung ran last and consumed a data-dependent number of draws
            ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
            uint32_t bt = __ballot_sync(0xFFFFFFFFu, valid && r == Nc - 1);
            int lt = __ffs((bt >> gbase) & gmask) - 1;
            if (lt < 0) lt = 0;
            base_nb = __shfl_sync(0xFFFFFFFFu, rr->nbp, lt, G);
            base_py = __shfl_sync(0xFFFFFFFFu, rr->pyp, lt, G);
            rr->nbp = base_nb;
            rr->pyp = base_py;
        }
        // The sweep is serial over the rung pairs (a replica can fall several rungs in one sweep), so one lane per ladder
        // walks it on rung-ordered copies in shared memory -- occupant lane and weight of every rung -- carrying the
        // replica that currently sits on the upper rung of the pair; the other lanes then read their new rung back.  In the
        // alpha ladder n_eff belongs to the rung (mcmc_alpha.py:126-131 swaps .code and .flag but not .n_eff), so there the
        // decisions depend on rung-owned values only.
        {
            const int wb = gbase;   // this ladder's slice of the per-warp arrays
            top_native.end = top_native.pos;   // the arrays overwrite the pool, and another lane may be on top afterwards
            if (!REPLAY && valid) s_sw_u[lane] = rng.nb(p.keys);  // lane i's draw decides pair (i, i+1)
            if (valid) {
                s_sw_lane[wb + r] = gl;
                s_sw_a[wb + r] = p.kind == LK_ALPHA ? e_nz : n;
                s_sw_b[wb + r] = e_nxy;
            }
            __syncwarp();
            // Alpha ladder: the decision of pair (i, i + 1) depends on rung-owned values only, so lane i takes it -- one
            // pow() per warp instead of one per pair; the reference draws one PY uniform per pair, top pair first
            // (mcmc_alpha.py:117-123), so pair i reads draw number Nc - 2 - i of this sweep.
            uint32_t alpha_swaps = 0;
            if (p.kind == LK_ALPHA) {
                bool sw = false;
                if (ladder < p.n_ladders && !done && gl < Nc - 1) {
                    const int i = gl;
                    const double ne_lo = __dadd_rn((double)s_sw_a[wb + i], __dmul_rn(p.alpha, (double)s_sw_b[wb + i]));
                    const double ne_hi = __dadd_rn((double)s_sw_a[wb + i + 1], __dmul_rn(p.alpha, (double)s_sw_b[wb + i + 1]));
                    double u;
                    if (REPLAY) {
                        ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
                        const int at = rr->pyp + (Nc - 2 - i);
                        if (at >= rr->n_py) { *rr->status = 2; u = 0.0; }
                        else u = rr->py_base[at];
                    } else {
                        u = s_sw_u[wb + i];
                    }
                    sw = u < pow(p.diff[i], __dadd_rn(ne_hi, -ne_lo));
                }
                alpha_swaps = (__ballot_sync(0xFFFFFFFFu, sw) >> gbase) & gmask;
                if (REPLAY && ladder < p.n_ladders && !done) reinterpret_cast<ReplayRng *>(&rng)->pyp += Nc - 1;
            }
            // Native depolarizing / biased ladders: pair (i, i + 1) swaps iff u_i < diff_i^k, k = n_hi - n_lo, and diff_i^k falls
            // with k -- so the lane on rung i turns its pair's draw into the largest exponent that still swaps (a search of
            // the power table, all lanes at once) and the walk over the pairs is left with one compare and one select per
            // pair: "the carried replica's weight <= n_lo + kmax".  Same decisions as evaluating u < diff^k pair by pair.
            // A draw below the table's last entry leaves the exponent open: the pair is marked, swaps outright while k is
            // inside the table and is evaluated as before beyond it.  A table that does not fall (diff >= 1) sends the ladder
            // down the general walk.
            bool fast_sweep = !REPLAY && p.kind == LK_ALPHA && !p.serial_sweep;   // group-uniform
            uint32_t swaps = alpha_swaps;
            if (!REPLAY && use_pw && !p.serial_sweep) {
                constexpr int OPEN = 1 << 30;
                const int kend = 2 * QECMC_PW_K - pw_off;   // exponents 0 .. kend are tabulated
                bool bad = false, open = false;
                if (valid && !done && r < Nc - 1) {
                    const double u = s_sw_u[wb + r];
                    const double *row = s_pw + r * (2 * QECMC_PW_K + 1) + pw_off;   // row[k] = diff_r^k
                    int cnt = 0;   // number of k in 0 .. kend with u < row[k]
#pragma unroll
                    for (int st = 32; st >= 1; st >>= 1) {
                        const int m = cnt + st;
                        if (m <= kend + 1 && u < row[m - 1]) cnt = m;
                    }
                    bad = !(row[1] < 1.0) || cnt == 0;
                    open = cnt == kend + 1;
                    s_sw_b[wb + r] = n + cnt - 1 + (open ? OPEN : 0);
                }
                const uint32_t bads = (__ballot_sync(0xFFFFFFFFu, bad) >> gbase) & gmask;
                const uint32_t opens = (__ballot_sync(0xFFFFFFFFu, open) >> gbase) & gmask;
                fast_sweep = bads == 0;
                __syncwarp();
                if (fast_sweep && ladder < p.n_ladders && !done) {
                    // every lane of the ladder walks the pairs on the same (broadcast) reads and ends up with the swap mask
                    int c_n = s_sw_a[wb + Nc - 1];
                    uint32_t m = 0;
                    if (opens == 0) {
                        for (int i = Nc - 2; i >= 0; i--) {
                            const int lo_n = s_sw_a[wb + i], t = s_sw_b[wb + i];
                            const bool sw = c_n <= t;
                            m |= (uint32_t)sw << i;
                            c_n = sw ? c_n : lo_n;
                        }
                    } else {
                        for (int i = Nc - 2; i >= 0; i--) {
                            const int lo_n = s_sw_a[wb + i], t = s_sw_b[wb + i];
                            bool sw = c_n <= t;
                            if (t >= OPEN && c_n > t - OPEN) sw = s_sw_u[wb + i] < numba_pow_dev(p.diff[i], c_n - lo_n);   // beyond the table
                            m |= (uint32_t)sw << i;
                            c_n = sw ? c_n : lo_n;
                        }
                    }
                    swaps = m;
                }
            }
            if (!fast_sweep && gl == 0 && ladder < p.n_ladders && !done) {
                int c_lane = s_sw_lane[wb + Nc - 1], c_n = s_sw_a[wb + Nc - 1];
                for (int i = Nc - 2; i >= 0; i--) {
                    const int lo_lane = s_sw_lane[wb + i], lo_n = s_sw_a[wb + i];
                    bool swap;
                    if (p.kind == LK_ALPHA) {
                        swap = (alpha_swaps >> i) & 1u;
                    } else {
                        const int ne_lo = lo_n, ne_hi = c_n;
                        if (p.kind == LK_DEPOL && ne_hi < ne_lo) {
                            swap = true;  // mcmc.py:146-147: no draw
                        } else {
                            const double u = REPLAY ? rng.nb(p.keys) : s_sw_u[wb + i];  // mcmc_biased.py:154-156 draws always
                            const int k = ne_hi - ne_lo;
                            const int ki = k + pw_off;
                            const double pw = (use_pw && ki >= 0 && ki <= 2 * QECMC_PW_K) ? s_pw[i * (2 * QECMC_PW_K + 1) + ki]
                                                                                           : numba_pow_dev(p.diff[i], k);
                            swap = u < pw;
                        }
                    }
                    if (swap) {
                        s_sw_rung[wb + lo_lane] = i + 1;          // the lower replica moves up; the carried one goes on down
                    } else {
                        s_sw_rung[wb + c_lane] = i + 1;           // the carried replica stays on rung i + 1
                        c_lane = lo_lane;
                        c_n = lo_n;
                    }
                }
                s_sw_rung[wb + c_lane] = 0;
            }
            __syncwarp();
            if (valid && !done) {
                if (fast_sweep) {
                    // from the swap mask: the lower replica of a swapping pair moves up one rung; any other replica falls past
                    // every swapping pair directly below it
                    if (r < Nc - 1 && ((swaps >> r) & 1u)) r = r + 1;
                    else if (r > 0) r -= __clz((int)~(swaps << (32 - r)));
                } else {
                    r = s_sw_rung[wb + gl];
                }
                if (p.kind == LK_ALPHA) { e_nz = s_sw_a[wb + r]; e_nxy = s_sw_b[wb + r]; }   // n_eff stays with the rung
            }
            if (REPLAY) {   // the walking lane consumed the draws: the whole ladder continues from its stream positions
                ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
                rr->nbp = __shfl_sync(0xFFFFFFFFu, rr->nbp, 0, G);
                rr->pyp = __shfl_sync(0xFFFFFFFFu, rr->pyp, 0, G);
            }
            __syncwarp();
        }
        if (REPLAY) {
            ReplayRng *rr = reinterpret_cast<ReplayRng *>(&rng);
            base_nb = rr->nbp;
            base_py = rr->pyp;
        }
        if (valid && !done && r == Nc - 1) flag = 1;
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, valid && r == 0);
        int l0 = __ffs((b0 >> gbase) & gmask) - 1;
        if (l0 < 0) l0 = 0;
        const int flag0 = __shfl_sync(0xFFFFFFFFu, flag, l0, G);
        if (!done && flag0 == 1) {
            tops0++;
            if (gl == l0) flag = 0;
        }
        // ---------------- snapshots (tests) ----------------
        if (p.snap_lat && valid && !done) {
            W *o = reinterpret_cast<W *>(p.snap_lat) + (((size_t)ladder * p.steps + step) * Nc + r) * g.nw;
            for (int w = 0; w < g.nw; w++) o[w] = lat.get(w);
            p.snap_flags[((size_t)ladder * p.steps + step) * Nc + r] = flag;
            if (gl == 0) p.snap_tops0[(size_t)ladder * p.steps + step] = tops0;
        }
        // ---------------- accounting ----------------
        if (p.acct == ACCT_PTEQ) {
            // decoders.py:56-82 (PTEQ), decoders_biasednoise.py:196-215 (PTEQ_alpha records n_eff of the bottom rung)
            const int cur = __shfl_sync(0xFFFFFFFFu, cls, l0, G);
            const int h_a = __shfl_sync(0xFFFFFFFFu, p.kind == LK_ALPHA ? e_nz : n, l0, G);
            const int h_b = __shfl_sync(0xFFFFFFFFu, p.kind == LK_ALPHA ? e_nxy : 0, l0, G);
            if (!done) {
                if (tops0 >= p.tops_burn) {
                    since_burn = step - burn_in;
                    if (gl == 0) {
                        eqc[class_raw_to_label(GEOM, cur)]++;
                        if (hist) hist[since_burn] = (uint32_t)h_a | ((uint32_t)h_b << 16);
                    }
                    if (p.track_shortest && gl == l0) {
                        // the lane holding the bottom rung owns its class, recorded value and fingerprint
                        const int lab = class_raw_to_label(GEOM, cur);
                        volatile double *sv = p.short_v + (size_t)ladder * g.neq + lab;
                        volatile long long *sn = p.short_n + (size_t)ladder * g.neq + lab;
                        const double v = p.kind == LK_ALPHA ? __dadd_rn((double)h_a, __dmul_rn(p.alpha, (double)h_b)) : (double)h_a;
                        const double cur_short = *sv;
                        if (v <= cur_short) {
                            if (v < cur_short) { *sv = v; *sn = 1; }
                            else *sn = *sn + 1;
                            // set of distinct bottom-rung states seen at the class's current shortest value: the key carries
                            // class and value, so entries of a superseded (longer) shortest value simply stop matching
                            const uint64_t aux = (uint64_t)h_a | ((uint64_t)h_b << 11) | ((uint64_t)lab << 22);
                            const uint64_t key = (h & ~((1ull << 26) - 1ull)) | aux | (1ull << 63);
                            unsigned long long *tab = p.tables + (uint64_t)ladder * (p.cap_mask + 1);
                            uint64_t slot = (key >> 26) & p.cap_mask;
                            while (true) {
                                unsigned long long curk = __ldcg(tab + slot);
                                if (curk == key) break;
                                if (curk == 0ull) {
                                    unsigned long long prev = atomicCAS(tab + slot, 0ull, (unsigned long long)key);
                                    if (prev == 0ull || prev == key) break;
                                }
                                slot = (slot + 1) & p.cap_mask;
                            }
                        }
                    }
                    // history windows [l/4, l/2) and [3l/4, l) of conv_crit_error_based_PT (decoders.py:93-105)
                    wl = since_burn + 1;
                    S4a += h_a; S4b += h_b;
                    const long long nC = 3 * wl / 4, nB = wl / 2, nA = wl / 4;
                    if (p.use_conv) {
                        // entries written by the group's lane 0 in earlier steps (the __syncwarp above orders them)
                        if (nC > wC) { uint32_t v = __ldcg(hist + wC); S4a -= v & 0xFFFF; S4b -= v >> 16; }
                        if (nB > wB) { uint32_t v = __ldcg(hist + wB); S2a += v & 0xFFFF; S2b += v >> 16; }
                        if (nA > wA) { uint32_t v = __ldcg(hist + wA); S2a -= v & 0xFFFF; S2b -= v >> 16; }
                    }
                    wC = nC; wB = nB; wA = nA;
                } else {
                    burn_in++;
                }
                if (p.use_conv && tops0 >= p.TOPS) {
                    const long long l = since_burn + 1;
                    double q2, q4;
                    if (p.kind == LK_ALPHA) {
                        q2 = ((double)S2a + p.alpha * (double)S2b) / (double)(l / 2 - l / 4);
                        q4 = ((double)S4a + p.alpha * (double)S4b) / (double)(l - 3 * l / 4);
                    } else {
                        q2 = (double)S2a / (double)(l / 2 - l / 4);
                        q4 = (double)S4a / (double)(l - 3 * l / 4);
                    }
                    const double err = fabs(q2 - q4);
                    if (err < p.eps) {
                        if (conv_streak >= p.SEQ) { done = true; converged = 1; steps_used = step + 1; }
                        conv_streak = tops0 - conv_start;
                    } else {
                        conv_streak = 0;
                        conv_start = tops0;
                    }
                }
            }
        } else if (p.acct == ACCT_DC) {
            // PTDC_droplet (decoders.py:146-155) / STDC_droplet_alpha (decoders.py:521-531): every rung offers its
            // state; a state unchanged since its last offer is already in the set
            bool is_new = false;
            if (valid && !done && dirty) {
                uint64_t aux = p.kind == LK_ALPHA ? ((uint64_t)nz | ((uint64_t)(nx + ny) << 11)) : (uint64_t)n;
                uint64_t amask = (1ull << p.aux_bits) - 1ull;
                uint64_t key = (h & ~amask) | aux | (1ull << 63);
                uint64_t slot = (key >> p.aux_bits) & p.cap_mask;
                while (true) {
                    unsigned long long curk = __ldcg(table + slot);
                    if (curk == key) break;
                    if (curk == 0ull) {
                        unsigned long long prev = atomicCAS(table + slot, 0ull, (unsigned long long)key);
                        if (prev == 0ull) { is_new = true; break; }
                        if (prev == key) break;
                    }
                    slot = (slot + 1) & p.cap_mask;
                }
                noff++;
                dirty = false;
            }
            if (p.conv_mult != 0.0) {
                // decoders.py:156-161, rung by rung: a chain new to the droplet and not longer than the shortest so far
                // moves `stop`; in rung order that leaves shortest = the smallest such length
                int cand = (is_new && n <= dc_shortest) ? n : 0x7FFFFFFF;
                for (int o = G >> 1; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xFFFFFFFFu, cand, o, G));
                if (!done) {
                    if (cand != 0x7FFFFFFF) { dc_shortest = cand; dc_stop = (double)step * p.conv_mult; }
                    if ((double)step >= dc_stop && step * 100 >= p.steps) { done = true; steps_used = step + 1; }
                }
            }
        } else if (p.acct == ACCT_RC) {
            // PTRC_droplet (decoders.py:597-625): every rung keeps its own set of distinct chains and m(n)
            if (valid && !done) {
                const uint64_t t = (uint64_t)ladder * Nc + r;
                atomicAdd(p.rc_m_hist + t * ns1 + n, 1ull);
                if (dirty || r != last_r) {
                    unsigned long long *tab = p.tables + t * (p.cap_mask + 1);
                    uint64_t key = make_key(h, n);
                    uint64_t slot = (key >> QECMC_LEN_BITS) & p.cap_mask;
                    while (true) {
                        unsigned long long curk = __ldcg(tab + slot);
                        if (curk == key) break;
                        if (curk == 0ull) {
                            unsigned long long prev = atomicCAS(tab + slot, 0ull, (unsigned long long)key);
                            if (prev == 0ull || prev == key) break;
                        }
                        slot = (slot + 1) & p.cap_mask;
                    }
                    noff++;
                    dirty = false;
                    last_r = r;
                }
            }
        }
    }
    __syncwarp();
    // ---------------- results ----------------
    if (valid) {
        if (p.lat_out) {
            W *o = reinterpret_cast<W *>(p.lat_out) + ((size_t)ladder * Nc + r) * g.nw;
            for (int w = 0; w < g.nw; w++) o[w] = lat.get(w);
        }
        if (p.flags_out) p.flags_out[(size_t)ladder * Nc + r] = flag;
        if (p.neff_out) p.neff_out[(size_t)ladder * Nc + r] = make_int2(e_nz, e_nxy);
        if (gl == 0) {
            if (p.tops0_out) p.tops0_out[ladder] = tops0;
            if (p.acct == ACCT_DC && p.info) p.info[ladder] = steps_used;
            if (p.acct == ACCT_PTEQ) {
                if (p.info) {
                    p.info[4 * ladder] = steps_used;
                    p.info[4 * ladder + 1] = since_burn;
                    p.info[4 * ladder + 2] = tops0;
                    p.info[4 * ladder + 3] = converged;
                }
                if (p.percent)  // decoders.py:89: (eq[since_burn] / (since_burn + 1) * 100).astype(np.uint8)
                    for (int e = 0; e < g.neq; e++)
                        p.percent[ladder * g.neq + e] = (uint8_t)(int)((double)eqc[e] / (double)(since_burn + 1) * 100);
            }
        }
        if (p.counters) {
            atomicAdd(p.counters + 0, (unsigned long long)nacc);
            atomicAdd(p.counters + 1, (unsigned long long)noff);
        }
    }
}

// PTRC estimator (decoders.py:699-739) per (syndrome, class): N(n), m(n) of every rung summed over the droplets,
// then for rungs i < Nc-1:  C_i = mean over the two shortest observed lengths of N/m * exp(-beta_i (n - n0)),
// Z += C_i * sum_n m(n) exp(n (beta_i - beta_error) - beta_i n0).
static __global__ void ptrc_finalize_kernel(const uint32_t *__restrict__ N_tab /* [tabs][droplets][Nc][ns1] */,
                                            const unsigned long long *__restrict__ m_tab, int64_t tabs, int droplets, int Nc,
                                            int ns1, const double *__restrict__ ladder_p, double beta_error,
                                            double *__restrict__ Z, long long *__restrict__ N_out, long long *__restrict__ m_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tabs) return;
    double z = 0;
    for (int i = 0; i < Nc; i++) {
        const double beta_i = -log((ladder_p[i] / 3) / (1 - ladder_p[i])), d_beta = beta_i - beta_error;
        int l0 = -1, l1 = -1;
        double r0 = 0, r1 = 0, sum = 0;
        // pass 1: two shortest lengths; pass 2 needs l0, so loop twice over the lengths
        for (int pass = 0; pass < 2; pass++)
            for (int l = 0; l < ns1; l++) {
                long long N = 0, m = 0;
                for (int d = 0; d < droplets; d++) {
                    size_t o = (((size_t)t * droplets + d) * Nc + i) * ns1 + l;
                    N += N_tab[o];
                    m += (long long)m_tab[o];
                }
                if (pass == 0) {
                    if (N_out) { N_out[((size_t)t * Nc + i) * ns1 + l] = N; m_out[((size_t)t * Nc + i) * ns1 + l] = m; }
                    if (m) {
                        if (l0 < 0) { l0 = l; r0 = (double)N / (double)m; }
                        else if (l1 < 0) { l1 = l; r1 = (double)N / (double)m * exp(-beta_i * (double)(l1 - l0)); }
                    }
                } else if (m) {
                    sum += (double)m * exp((double)l * d_beta - beta_i * (double)l0);
                }
            }
        if (i < Nc - 1 && l0 >= 0) {
            double c_mean = l1 >= 0 ? (r0 + r1) / 2.0 : r0;
            z += c_mean * sum;
        }
    }
    Z[t] = z;
}

// PTEQ_alpha_with_shortest: distinct bottom-rung states recorded at each class's final shortest value
static __global__ void short_unique_kernel(const unsigned long long *__restrict__ tables, uint64_t cap, int n_eq, int alpha_kind,
                                           double alpha, const double *__restrict__ short_v, long long *__restrict__ uniq)
{
    __shared__ unsigned int s_cnt[16];
    if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long *tab = tables + (uint64_t)blockIdx.x * cap;
    for (uint64_t i = threadIdx.x; i < cap; i += blockDim.x) {
        unsigned long long k = tab[i];
        if (!k) continue;
        int a = (int)(k & 0x7FF), b = (int)((k >> 11) & 0x7FF), lab = (int)((k >> 22) & 0xF);
        double v = alpha_kind ? __dadd_rn((double)a, __dmul_rn(alpha, (double)b)) : (double)a;
        if (lab < n_eq && v == short_v[(size_t)blockIdx.x * n_eq + lab]) atomicAdd(&s_cnt[lab], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_eq) uniq[(size_t)blockIdx.x * n_eq + threadIdx.x] = s_cnt[threadIdx.x];
}

// Z_E of a (syndrome, class) table whose keys carry (nz, nx+ny) in their low 22 bits:
// sum over distinct chains of exp(-beta * (nz + alpha (nx+ny))) (decoders.py:568,577-578)
static __global__ void table_sum_alpha_kernel(const unsigned long long *__restrict__ tables, uint64_t cap, double beta, double alpha,
                                       double *__restrict__ Z, unsigned long long *distinct_out, unsigned long long *distinct_total)
{
    __shared__ double s_z[256];
    __shared__ unsigned long long s_c[256];
    const unsigned long long *tab = tables + (uint64_t)blockIdx.x * cap;
    double z = 0;
    unsigned long long cnt = 0;
    for (uint64_t i = threadIdx.x; i < cap; i += blockDim.x) {
        unsigned long long k = tab[i];
        if (k) {
            int nz = (int)(k & 0x7FF), nxy = (int)((k >> 11) & 0x7FF);
            z += exp(-beta * ((double)nz + alpha * (double)nxy));
            cnt++;
        }
    }
    s_z[threadIdx.x] = z;
    s_c[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { s_z[threadIdx.x] += s_z[threadIdx.x + s]; s_c[threadIdx.x] += s_c[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Z[blockIdx.x] = s_z[0];
        if (distinct_out) distinct_out[blockIdx.x] = s_c[0];
        if (distinct_total) atomicAdd(distinct_total, s_c[0]);
    }
}

}  // namespace qecmc

// qecmc_stdc_fast.cuh -- table-driven STDC chain kernel for the two-layer codes (toric, planar),
// i.e. every lattice the reference's fast path (_update_chain_fast, src/mcmc.py:152-160) accepts.
//
// Per Metropolis step the generic kernel spends most of its issue slots decoding the stabilizer
// and popcounting weight changes.  Here both are table lookups in shared memory:
//   * an 8-byte descriptor per stabilizer: three row-word indices, two column shifts, the Pauli
//     and (planar boundary stabilizers) which of the four qubit slots exist;
//   * the four touched 2-bit fields are gathered into one byte f, and a 512-entry table indexed
//     by (Pauli, f) returns the Metropolis acceptance threshold for the resulting weight change
//     directly (a second byte table returns the weight change itself, read only on accept).
// Slots:  a = (w0, sh)  b = (w0, sh2)  c = (w1, sh)  d = (w2, sh).
// A missing slot is forced to the field value 2 (Y), which neither X nor Z flips change in weight.
#pragma once
#include "qecmc_kernels.cuh"

namespace qecmc {

// descriptor words:  x = sh | w0<<8 | w1<<16 | w2<<24
//                    y = sh2 | f_and<<8 | f_or<<16 | (v==3)<<24
#define QECMC_FAST_STATIC_NSTAB 512   // 2 * 16 * 16: every two-layer code with 32-bit row words

struct FastTables {
    const uint2 *desc;       // [nstab]
    const uint32_t *thr;     // [512] native thresholds by (v==3)*256 + f
    const int8_t *dE;        // [512]
};

template <typename W> __device__ __forceinline__ uint32_t gather_fields(W o0, W o1, W o2, uint32_t x, uint32_t y);

template <> __device__ __forceinline__ uint32_t gather_fields<uint32_t>(uint32_t o0, uint32_t o1, uint32_t o2, uint32_t x, uint32_t y)
{
    // rotate each word so its field lands at bits [2i, 2i+2) of slot i; the funnel shift uses the
    // low five bits of the shift operand, so x / y are used as they are
    uint32_t ta = __funnelshift_r(o0, o0, x);
    uint32_t tb = __funnelshift_r(o0, o0, y - 2u);
    uint32_t tc = __funnelshift_r(o1, o1, x - 4u);
    uint32_t td = __funnelshift_r(o2, o2, x - 6u);
    uint32_t s1 = (ta & 0x03u) | (tb & ~0x03u);
    uint32_t s2 = (tc & 0x30u) | (td & ~0x30u);
    return (s1 & 0x0Fu) | (s2 & ~0x0Fu);  // caller masks to 8 bits
}

template <> __device__ __forceinline__ uint32_t gather_fields<uint64_t>(uint64_t o0, uint64_t o1, uint64_t o2, uint32_t x, uint32_t y)
{
    uint32_t sh = x & 63u, sh2 = y & 63u;
    uint32_t qa = (uint32_t)(o0 >> sh) & 3u, qb = (uint32_t)(o0 >> sh2) & 3u;
    uint32_t qc = (uint32_t)(o1 >> sh) & 3u, qd = (uint32_t)(o2 >> sh) & 3u;
    return qa | (qb << 2) | (qc << 4) | (qd << 6);
}

// The static tables of the 32-bit-word variant live in one struct and are read through ld.shared with an explicit
// 32-bit base + immediate offset: left to itself the compiler re-derives every table's shared-window address (three
// instructions each, CTA rank in the cluster included) at every use instead of keeping five bases in registers.
// A and thr are looked up with a random index by every lane at every step; stored once, those reads cost ~10 and ~3.5
// shared-memory wavefronts through bank conflicts.  They are therefore kept in REP copies interleaved per
// entry, lane j reading copy j & 7: the eight lanes of a quarter-warp then hit eight different 16-byte bank groups
// whatever their indices are, i.e. the 16-byte read is conflict-free (4 wavefronts) and the 4-byte read nearly so.
// (Only the one-CTA-per-SM variant replicates, REP = 8; the early-stop / replay variants share the SM among four CTAs.)
template <int REP> struct FastTabs {
    uint4 A[QECMC_FAST_STATIC_NSTAB * REP];   // [stabilizer][copy] {byte offset of word 0, 1, 2 in the thread's tile column, packed shifts / masks}
    uint32_t thr[512 * REP];                  // [(Pauli, gathered fields)][copy] threshold
    uint4 B[QECMC_FAST_STATIC_NSTAB];       // {XOR mask of word 0, 1, 2}: read on accept only
    uint64_t hs[QECMC_FAST_STATIC_NSTAB];   // fingerprint change per stabilizer
    int8_t dE[512];                         // weight change by (Pauli, gathered fields)
};
template <int OFF> __device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ int lds_s8(uint32_t a)
{
    int v;
    asm("ld.shared.s8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}

__device__ __forceinline__ void prefetch_l2(const void *addr)
{
    asm volatile("prefetch.global.L2 [%0];" : : "l"(addr));
}

// CONV = false drops the conv_mult early-stop state (ConvStop) from the loop: the headline configuration
// (conv_mult == 0) then fits the register budget of five 256-thread CTAs per SM.
// That variant runs as ONE 1024-thread CTA per SM with up to 64 registers.  Measured at the headline configuration, per
// syndrome: five 256-thread CTAs (48 registers, 1280 chains per SM) 1.03 ms, two 640-thread CTAs 0.96 ms, one
// 1024-thread CTA 0.92 ms -- fewer rematerialised addresses (95.8 instead of 98.8 issue slots per step) and 74 % instead
// of 72 % of the issue cycles used.  With several CTAs per SM the oldest-first warp scheduler lets one CTA finish long
// before the others (ncu: 31.8 of 40 warps active on average); pacing the CTAs against each other through a grid-wide
// epoch counter brought that to 39.98 of 40 but did not raise the issue rate, so it was dropped.
// Shared memory: the LUTs are static arrays, and for 32-bit row words (L <= 16, at most 512 stabilizers) so are the
// descriptors and fingerprints, which makes every table address an immediate; the lattice tile is the dynamic part.
// BLOG = true is the headline specialisation: insert mode 6 known at compile time (no mode dispatch in the sample path).
template <int GEOM, typename W, bool REPLAY, int MODE, bool CONV, bool BLOG = false>
__global__ void __launch_bounds__((!CONV && sizeof(W) == 4) ? 1024 : 256, CONV ? (sizeof(W) == 4 ? 4 : 2) : (sizeof(W) == 4 ? 1 : 3)) stdc_fast_kernel(StdcParams p, FastTables ft, PhiloxKeys keys)
{
    static_assert(GEOM == TORIC || GEOM == PLANAR, "table-driven kernel covers the two-layer codes");
    constexpr bool STATIC_TAB = sizeof(W) == 4;
    constexpr int REP = (STATIC_TAB && !CONV) ? 8 : 1;
    typedef FastTabs<REP> Tabs;
    constexpr size_t TABS_BYTES = STATIC_TAB ? ((sizeof(Tabs) + 15) & ~(size_t)15) : 0;   // the tables open the dynamic shared memory
    extern __shared__ __align__(16) unsigned char smem_all[];
    unsigned char *smem = smem_all + TABS_BYTES;                                         // tile (and, for 64-bit words, the plain tables)
    __shared__ uint32_t s_thr_dyn[STATIC_TAB ? 1 : 512];
    __shared__ int8_t s_dE_dyn[STATIC_TAB ? 4 : 512];
    __shared__ double s_thrd[QECMC_THR_N];
    Tabs &tb = *reinterpret_cast<Tabs *>(smem_all);
    uint32_t *s_thr = s_thr_dyn;
    int8_t *s_dE = STATIC_TAB ? tb.dE : s_dE_dyn;
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(smem_all);
    const uint32_t tbaseA = tbase + (threadIdx.x & (REP - 1)) * 16u;   // this lane's copy of A
    // the accept path's tables (B, hs, dE) are addressed from a copy of the base the compiler cannot re-derive: left to itself it
    // rebuilds the shared-window address (S2UR + UMOV + ULEA) inside the branch, three issue slots for a handful of lanes
    uint32_t tbase_acc;
    asm volatile("mov.u32 %0, %1;" : "=r"(tbase_acc) : "r"(tbase));
    const uint32_t tbaseT = tbase + (threadIdx.x & (REP - 1)) * 4u;    // ... and of thr
    const int T = blockDim.x, tid = threadIdx.x;
    const Geo g = p.gchain;
    W *tile = reinterpret_cast<W *>(smem);
    uint64_t *s_hs = STATIC_TAB ? tb.hs : reinterpret_cast<uint64_t *>(smem + (((size_t)g.nw * T * sizeof(W) + 15) & ~(size_t)15));
    uint2 *s_desc = STATIC_TAB ? nullptr : reinterpret_cast<uint2 *>(s_hs + g.nstab);
    for (int i = tid; i < g.nstab; i += T) {
        s_hs[i] = p.stab_hash[i];
        const uint2 d = ft.desc[i];
        if (STATIC_TAB) {
            const uint32_t ws = (uint32_t)T * sizeof(W);
            const uint32_t sh = d.x & 63u, sh2 = d.y & 63u, fa = (d.y >> 8) & 0xFFu, f_or9 = d.y >> 16;
            const uint32_t v = (d.y & 0x01000000u) ? 3u : 1u;
            uint32_t m0, m1, m2;
            if (GEOM == TORIC) { m1 = m2 = v << sh; m0 = m1 | (v << sh2); }
            else {
                m0 = ((fa & 1u) ? v << sh : 0u) | ((fa & 4u) ? v << sh2 : 0u);
                m1 = (fa & 16u) ? v << sh : 0u;
                m2 = (fa & 64u) ? v << sh : 0u;
            }
            const uint4 a = make_uint4(((d.x >> 8) & 0xFFu) * ws, ((d.x >> 16) & 0xFFu) * ws, (d.x >> 24) * ws,
                                       sh | (((sh2 - 2u) & 31u) << 8) | (f_or9 << 13) | (GEOM == PLANAR ? fa << 22 : 0u));
            for (int r = 0; r < REP; r++) tb.A[i * REP + r] = a;
            tb.B[i] = make_uint4(m0, m1, m2, 0u);
        } else {
            s_desc[i] = d;
        }
    }
    for (int i = tid; i < 512; i += T) {
        s_dE[i] = ft.dE[i];
        if (STATIC_TAB) for (int r = 0; r < REP; r++) tb.thr[i * REP + r] = ft.thr[i];
        else s_thr[i] = ft.thr[i];
    }
    if (tid < QECMC_THR_N) s_thrd[tid] = p.thr.d[tid];
    if (BLOG) {   // mode 6: this CTA's cursors into the bucket logs of its tables live behind the tile
        uint32_t *cur0 = reinterpret_cast<uint32_t *>(smem + (((size_t)g.nw * T * sizeof(W) + 15) & ~(size_t)15));
        for (int i = tid; i < p.tables_per_cta * p.nbc; i += T) cur0[i] = 0;
    }
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * T + tid;
    if (local >= p.n_chains) return;   // spare threads of the wave's last CTA (whole tables only: n_chains is a multiple of droplets)
    const int64_t gchain = p.chain_offset + local;
    const int n_eq = p.gcode.neq;
    const int64_t tab = local / p.droplets;
    const int eq = (int)(tab % n_eq);
    const int64_t sw = tab / n_eq;
    SmemLat<W> lat{tile + tid, T};
    {
        const W *src = reinterpret_cast<const W *>(p.lat0) + (p.per_class ? tab : sw) * g.nw;
        for (int w = 0; w < g.nw; w++) lat.set(w, src[w]);
    }
    const uint32_t cl = (uint32_t)gchain, chh = (uint32_t)((uint64_t)gchain >> 32);
    if (!p.per_class) to_class_rt<W>(p.gcode, lat, eq);
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
                        if ((i & 127) == 0) r = philox4x32_10((uint32_t)(i >> 7), 0x80000000u, cl, chh, keys);
                        int wi = (i >> 5) & 3;
                        uint32_t word = wi == 0 ? r.x : wi == 1 ? r.y : wi == 2 ? r.z : r.w;
                        hit = (word >> (i & 31)) & 1;
                    }
                    if (hit && rain_legal_rt(p.gcode, o, rr, c)) apply_rco_rt<W>(p.gcode, lat, rr, c, o == 0 ? 3 : 1);
                }
    }
    int n = lat_weight<W>(g, lat);
    uint64_t h = lat_hash<W>(g, lat, p.hash_seed);
    const uint64_t cap_mask = p.cap_mask;
    // the early stop needs the probe's answer at once: mode 5 (a set per chain + a log of the keys new to it) or 0 (the class's set)
    const int imode = BLOG ? 6 : MODE == MODE_MEAN ? 3 : (CONV && p.conv_mult != 0.0 ? (p.insert_mode == 5 ? 5 : 0) : p.insert_mode);
    // mode 4: `table` is this chain's key log; mode 5: this chain's own set
    unsigned long long *table = imode == 4 ? p.logs + (uint64_t)local * (uint64_t)p.log_cap
                              : p.tables + (uint64_t)(imode == 5 ? local : tab) * (cap_mask + 1);
    uint32_t nlog = 0;
    // mode 6: this CTA's cursors into the bucket logs of its tables live behind the tile
    uint32_t *s_cur = reinterpret_cast<uint32_t *>(smem + (((size_t)g.nw * T * sizeof(W) + 15) & ~(size_t)15));
    // this thread's table: its cursors (shared-memory byte address) and its bucket logs
    const uint32_t cur_base = (uint32_t)__cvta_generic_to_shared(s_cur) + (uint32_t)(tid / p.droplets) * ((uint32_t)p.nbc * 4u);
    unsigned long long *blog_tab = BLOG ? p.blogs + (uint64_t)tab * (uint64_t)p.nbc * p.bcap : nullptr;

    // "the state changed since the last sample" = the accept counter moved: no flag of its own to set on the accept path
    // (which a warp runs for 2-5 active lanes); the first sample is always new to the chain
    uint32_t nacc = 0, noff = 0, nacc_seen = 0xFFFFFFFFu;
    int left = p.iters;
    // Distinct-chain accounting by insert mode: 6 / 4 log the key and leave the counting to a dedupe kernel; 5 and 0 probe a
    // set in HBM at once (early stop); 2 (key counts beyond the dedupe kernels' fan-out) prefetches the set's slot at one
    // sample and probes it at the next.
    unsigned long long key = 0;   // mode 2: key != 0 means `key` is waiting to be probed at `slot`
    uint32_t slot = 0;
    const uint32_t smask = (uint32_t)cap_mask;  // host guarantees cap <= 2^32 slots
    ConvStopT<CONV> cs;
    cs.init(p);
    SampleAcct<MODE> acct;
    acct.init(p, tab);
    unsigned char *mybase = reinterpret_cast<unsigned char *>(tile + tid);
    const uint32_t wstride = (uint32_t)T * sizeof(W);

    // one Metropolis step for stabilizer idx; r_acc / u_acc is the accept draw
    auto step = [&](int idx, uint32_t r_acc, double u_acc) {
        W *p0, *p1, *p2;
        W o0, o1, o2;
        uint32_t li, dx = 0, dy = 0;
        if (STATIC_TAB) {
            const uint4 A = lds_v4<offsetof(Tabs, A)>(tbaseA + (uint32_t)idx * (16u * REP));
            p0 = reinterpret_cast<W *>(mybase + A.x);
            p1 = reinterpret_cast<W *>(mybase + A.y);
            p2 = reinterpret_cast<W *>(mybase + A.z);
            o0 = *p0; o1 = *p1; o2 = *p2;
            // the funnel shift reads the low five bits of its shift operand, so the packed word serves as it is
            const uint32_t a0 = (uint32_t)o0, a1 = (uint32_t)o1, a2 = (uint32_t)o2;
            const uint32_t ta = __funnelshift_r(a0, a0, A.w);
            const uint32_t tb = __funnelshift_r(a0, a0, A.w >> 8);
            const uint32_t tc = __funnelshift_r(a1, a1, A.w - 4u);
            const uint32_t td = __funnelshift_r(a2, a2, A.w - 6u);
            const uint32_t s1 = (ta & 0x03u) | (tb & ~0x03u);
            const uint32_t s2 = (tc & 0x30u) | (td & ~0x30u);
            const uint32_t f = (s1 & 0x0Fu) | (s2 & ~0x0Fu);
            if (GEOM == TORIC) li = (f & 0xFFu) | (A.w >> 13);
            else li = (f & (A.w >> 22)) | ((A.w >> 13) & 0x1FFu);
        } else {
            const uint2 d = s_desc[idx];
            dx = d.x; dy = d.y;
            p0 = reinterpret_cast<W *>(mybase + ((d.x >> 8) & 0xFFu) * wstride);
            p1 = reinterpret_cast<W *>(mybase + ((d.x >> 16) & 0xFFu) * wstride);
            p2 = reinterpret_cast<W *>(mybase + (d.x >> 24) * wstride);
            o0 = *p0; o1 = *p1; o2 = *p2;
            uint32_t f = gather_fields<W>(o0, o1, o2, d.x, d.y);
            if (GEOM == TORIC) li = (f & 0xFFu) | (d.y >> 16);                            // f_or carries (v==3) at bit 8
            else li = (f & ((d.y >> 8) & 0xFFu)) | (d.y >> 16);
        }
        bool acc;
        if (REPLAY) acc = u_acc < s_thrd[(int)s_dE[li] + QECMC_THR_OFF];
        else if (STATIC_TAB) acc = r_acc <= lds_u32<offsetof(Tabs, thr)>(tbaseT + li * (4u * REP));
        else acc = r_acc <= s_thr[li];
        if (acc) {
            if (STATIC_TAB) {
                const uint4 B = lds_v4<offsetof(Tabs, B)>(tbase_acc + (uint32_t)idx * 16u);
                *p0 = (W)(o0 ^ (W)B.x);
                *p1 = (W)(o1 ^ (W)B.y);
                *p2 = (W)(o2 ^ (W)B.z);
            } else {
                const uint32_t sh = dx & 63u, sh2 = dy & 63u;
                const W v = (dy & 0x01000000u) ? (W)3 : (W)1;
                W m0, m1, m2;
                if (GEOM == TORIC) {
                    m1 = (W)(v << sh);
                    m2 = m1;
                    m0 = (W)(m1 | (W)(v << sh2));
                } else {
                    const uint32_t fa = dy >> 8;
                    m0 = (W)(((fa & 1u) ? (W)(v << sh) : (W)0) | ((fa & 4u) ? (W)(v << sh2) : (W)0));
                    m1 = (fa & 16u) ? (W)(v << sh) : (W)0;
                    m2 = (fa & 64u) ? (W)(v << sh) : (W)0;
                }
                *p0 = (W)(o0 ^ m0);
                *p1 = (W)(o1 ^ m1);
                *p2 = (W)(o2 ^ m2);
            }
            if (STATIC_TAB) {
                n += lds_s8<offsetof(Tabs, dE)>(tbase_acc + li);
                const uint2 hv = lds_v2<offsetof(Tabs, hs)>(tbase_acc + (uint32_t)idx * 8u);
                h ^= (uint64_t)hv.x | ((uint64_t)hv.y << 32);
            } else {
                n += (int)s_dE[li];
                h ^= s_hs[idx];
            }
            nacc++;
        }
        if (--left == 0) {
            left = p.iters;
            acct.sample(n);
            bool is_new = false;
            const bool dirty = nacc != nacc_seen;
            if (BLOG) {
                if (dirty) {
                    const uint64_t k = make_key(h, n);
                    const uint32_t b = (uint32_t)(k >> QECMC_LEN_BITS) & (uint32_t)(p.nbc - 1);
                    uint32_t pos;
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(cur_base + b * 4u) : "memory");
                    if (pos < p.bcap) {
                        blog_tab[b * p.bcap + pos] = k;
                    } else {   // rare: the bucket log is full
                        const uint32_t o = atomicAdd(p.ovf_cnt + tab, 1u);
                        if (o < p.ovf_cap) p.ovf[(uint64_t)tab * p.ovf_cap + o] = k;
                        else *p.log_err = 3;
                    }
                }
            } else if (imode == 4) {
                if (dirty) table[noff] = make_key(h, n);   // fire-and-forget store; log_dedupe_kernel counts later
            } else if (imode == 2) {
                // Two-phase insert: the sample that produces a key only prefetches its slot's sector into L2 (no
                // destination register, so nothing to wait for); the probe itself runs one sample later against a
                // line that is by then L2-resident, and a collision walks on within the same sector 3 times in 4.
                if (key) {
                    unsigned long long q;
                    do {
                        q = atomicCAS(table + slot, 0ull, key);
                        slot = (slot + 1u) & smask;
                    } while (q != 0ull && q != key);
                    key = 0;
                }
                if (dirty) {
                    key = make_key(h, n);
                    slot = (uint32_t)(key >> QECMC_LEN_BITS) & smask;
                    prefetch_l2(table + slot);
                }
            } else if (imode == 0) {
                if (dirty) is_new = table_insert(table, cap_mask, make_key(h, n));
            } else if (CONV && imode == 5) {
                if (dirty) {
                    const uint64_t k = make_key(h, n);
                    is_new = table_insert(table, cap_mask, k);
                    if (is_new) p.logs[(uint64_t)local * (uint64_t)p.log_cap + nlog++] = k;
                }
            }
            noff += dirty;
            nacc_seen = nacc;
            cs.after_sample(p, is_new, n);
        }
    };

    const uint64_t tsteps = (uint64_t)p.steps * (uint64_t)p.iters;
    if (REPLAY) {
        constexpr int K = NumDraws<GEOM>::value;
        const double *u = p.u_nb + (uint64_t)gchain * tsteps * (K + 1);
        for (uint64_t t = 0; t < tsteps && !cs.fin; t++, u += K + 1) {
            int row, col, op;
            propose_replay<GEOM>(g, u, row, col, op);
            step(rco_to_idx<GEOM>(g, row, col, op), 0u, u[K]);
        }
    } else {
        const uint32_t ncalls = (uint32_t)(tsteps >> 1);  // host guarantees tsteps < 2^32
        const uint32_t nstab = (uint32_t)g.nstab;
        for (uint32_t c0 = 0; c0 < ncalls && !cs.fin; c0++) {
            uint4 r = philox4x32_10(c0, 0u, cl, chh, keys);
            step((int)__umulhi(r.x, nstab), r.y, 0.0);
            if (!cs.fin) step((int)__umulhi(r.z, nstab), r.w, 0.0);
        }
        if ((tsteps & 1) && !cs.fin) {
            uint4 r = philox4x32_10(ncalls, 0u, cl, chh, keys);
            step((int)__umulhi(r.x, nstab), r.y, 0.0);
        }
    }
    if (imode == 2 && key) {
        unsigned long long q;
        do {
            q = atomicCAS(table + slot, 0ull, key);
            slot = (slot + 1u) & smask;
        } while (q != 0ull && q != key);
    }
    if (BLOG) {
        __syncthreads();   // every chain of the CTA has logged its last key (threads that exited above do not count)
        const int64_t tab0 = (int64_t)blockIdx.x * p.tables_per_cta;
        const int64_t left = p.n_chains - (int64_t)blockIdx.x * T;
        const int nact = left < T ? (int)left : T;   // threads still here
        for (int i = tid; i < p.tables_per_cta * p.nbc; i += nact)
            if (tab0 + i / p.nbc < p.n_chains / p.droplets) p.bcounts[tab0 * p.nbc + i] = min(s_cur[i], p.bcap);
    }
    if (imode == 4) p.log_counts[local] = noff;
    if (imode == 5) p.log_counts[local] = nlog;
    acct.finish(p, local);
    atomicAdd(p.counters + 0, (unsigned long long)nacc);
    atomicAdd(p.counters + 1, (unsigned long long)noff);
    atomicAdd(p.steps_done, CONV ? (unsigned long long)cs.samples() * (unsigned long long)p.iters : (unsigned long long)tsteps);
}

}  // namespace qecmc

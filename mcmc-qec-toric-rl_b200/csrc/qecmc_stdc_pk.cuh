// qecmc_stdc_pk.cuh -- table-driven STDC chain kernel for the two-layer codes with 17 <= L <= 24 (planar d = 17-21 of the
// threshold sweep), on a PACKED lattice of 32-bit words.
//
// Same Metropolis step, same draws and same results as stdc_fast_kernel<GEOM, uint64_t, ...> (_update_chain_fast,
// src/mcmc.py:152-160); what changes is where a chain's lattice lives.  With 64-bit row words a planar d = 21 chain takes
// 42 x 8 = 336 bytes of shared memory although only 42 of every 64 bits are used, which leaves 512 chains (16 warps) per
// SM, and every field access is 64-bit shift arithmetic.  Here a row is split at column 16:
//   * word r (r < nw)            = columns 0..15 of row r, the layout of the 32-bit kernel;
//   * word nw + r / k, bits [(r mod k) * hb, +hb) = columns 16..L-1 of row r, hb = 2 (L - 16) bits, k = 32 / hb rows per word.
// d = 21: 42 + 14 (+ 1 zero word) = 228 bytes per chain -> ~650 chains (20 warps) per SM beside the tables, and every field is
// (word, 5-bit shift) again, so the step is the 32-bit kernel's: ONE 8-byte record per stabilizer (four word indices, four
// rotate counts, the LUT index bits; the kernel is bound by shared-memory wavefronts, and an 8-byte record can be kept in
// 4-8 interleaved copies so that the lanes' random reads do not collide), four funnel shifts, one threshold read.  The four
// touched fields may now share words in any combination (the pair of a row straddles the split, two rows share a packed
// high word, ...), so each slot carries the XOR mask of ALL fields of the stabilizer that live in its word: slots that
// alias store the same value.
// Native draws only, no early stop (conv_mult == 0): everything else stays on the 64-bit kernel.
#pragma once
#include "qecmc_stdc_fast.cuh"

namespace qecmc {

struct PkGeo {
    int nw, hb, k, nwp;   // rows; bits of a row's high part; rows per packed high word; 32-bit words per chain
};
__host__ __device__ inline PkGeo pk_geo(const Geo &g)
{
    PkGeo q;
    q.nw = g.nw;
    q.hb = 2 * (g.L - 16);
    q.k = 32 / q.hb;
    q.nwp = g.nw + (g.nw + q.k - 1) / q.k + 1;   // + one word that stays 0: what a slot the planar boundary lacks reads
    return q;
}
__host__ __device__ inline bool pk_supported(const Geo &g) { return g.layers == 2 && g.L > 16 && g.L <= 24; }

// shared-memory carve-up of the tables in front of the tile (bytes)
struct PkLayout {
    uint32_t a, thr, b, hs, de, tile;
};
__host__ __device__ inline PkLayout pk_layout(int nstab, int rep)
{
    PkLayout o;
    uint32_t off = 0;
    o.a = off; off += (uint32_t)nstab * 8u * (uint32_t)rep;
    o.thr = off; off += 512u * 4u * (uint32_t)rep;
    o.b = off; off += (uint32_t)nstab * 16u;
    o.hs = off; off += (uint32_t)nstab * 8u;
    o.de = off; off += 512u;
    o.tile = (off + 15u) & ~15u;
    return o;
}

// a chain's packed lattice seen as 64-bit row words (initialisation only: class move, rain, weight, fingerprint)
struct PackedLat {
    uint32_t *base;
    int stride, nw, hb, k;
    __device__ __forceinline__ uint64_t get(int w) const
    {
        const uint32_t lo = base[(size_t)w * stride];
        const uint32_t hw = base[(size_t)(nw + w / k) * stride];
        const uint32_t hi = (hw >> ((w % k) * hb)) & ((1u << hb) - 1u);
        return (uint64_t)lo | ((uint64_t)hi << 32);
    }
    __device__ __forceinline__ void set(int w, uint64_t v)
    {
        base[(size_t)w * stride] = (uint32_t)v;
        const size_t i = (size_t)(nw + w / k) * stride;
        const int sh = (w % k) * hb;
        const uint32_t m = ((1u << hb) - 1u) << sh;
        base[i] = (base[i] & ~m) | (((uint32_t)(v >> 32) << sh) & m);
    }
};

// reads of the (read-only) tables through explicit shared-space addresses
__device__ __forceinline__ uint32_t lds_r32(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_r128(uint32_t a)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_r64(uint32_t a)
{
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_rs8(uint32_t a)
{
    int v;
    asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

template <int GEOM, int MODE, int REP>
__global__ void __launch_bounds__(1024, 1) stdc_pk_kernel(StdcParams p, FastTables ft, PhiloxKeys keys)
{
    static_assert(GEOM == TORIC || GEOM == PLANAR, "table-driven kernel covers the two-layer codes");
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const Geo g = p.gchain;
    const PkGeo q = pk_geo(g);
    const PkLayout lay = pk_layout(g.nstab, REP);
    uint2 *s_a = reinterpret_cast<uint2 *>(smem + lay.a);
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(smem + lay.thr);
    uint4 *s_b = reinterpret_cast<uint4 *>(smem + lay.b);
    uint64_t *s_hs = reinterpret_cast<uint64_t *>(smem + lay.hs);
    int8_t *s_dE = reinterpret_cast<int8_t *>(smem + lay.de);
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem + lay.tile);
    for (int i = tid; i < g.nstab; i += T) {
        s_hs[i] = p.stab_hash[i];
        const uint2 d = ft.desc[i];
        // slots a = (w0, sh), b = (w0, sh2), c = (w1, sh), d = (w2, sh) of the 64-bit descriptor -> (packed word, shift)
        const uint32_t sh = d.x & 63u, sh2 = d.y & 63u, f_and = (d.y >> 8) & 0xFFu, f_or9 = d.y >> 16;
        const uint32_t v = (d.y & 0x01000000u) ? 3u : 1u;
        const uint32_t rw[4] = {(d.x >> 8) & 0xFFu, (d.x >> 8) & 0xFFu, (d.x >> 16) & 0xFFu, d.x >> 24};
        const uint32_t rs[4] = {sh, sh2, sh, sh};
        uint32_t word[4], shf[4], msk[4];
        for (int s = 0; s < 4; s++) {
            if (!((f_and >> (2 * s)) & 1u)) {
                // a slot the planar boundary lacks reads the chain's zero word; f_or turns the 0 into the Y the LUT expects
                word[s] = (uint32_t)q.nwp - 1u; shf[s] = 0u; msk[s] = 0u;
                continue;
            }
            if (rs[s] < 32u) { word[s] = rw[s]; shf[s] = rs[s]; }
            else { word[s] = (uint32_t)q.nw + rw[s] / (uint32_t)q.k; shf[s] = (rw[s] % (uint32_t)q.k) * (uint32_t)q.hb + (rs[s] - 32u); }
            msk[s] = v << shf[s];
        }
        uint32_t cm[4];
        for (int s = 0; s < 4; s++) {
            cm[s] = 0;
            for (int t = 0; t < 4; t++)
                if (word[t] == word[s]) cm[s] |= msk[t];
        }
        // one 8-byte record: x = the four word indices (a byte each), y = the four rotate counts + the LUT bits f_or
        const uint2 a = make_uint2(word[0] | (word[1] << 8) | (word[2] << 16) | (word[3] << 24),
                                   shf[0] | (((shf[1] - 2u) & 31u) << 5) | (((shf[2] - 4u) & 31u) << 10) | (((shf[3] - 6u) & 31u) << 15) | (f_or9 << 20));
        for (int r = 0; r < REP; r++) s_a[i * REP + r] = a;
        s_b[i] = make_uint4(cm[0], cm[1], cm[2], cm[3]);
    }
    for (int i = tid; i < 512; i += T) {
        s_dE[i] = ft.dE[i];
        for (int r = 0; r < REP; r++) s_thr[i * REP + r] = ft.thr[i];
    }
    // insert mode 6: this CTA's cursors into the bucket logs of its (whole) tables live behind the tile
    uint32_t *s_cur = reinterpret_cast<uint32_t *>(smem + lay.tile + (((size_t)q.nwp * T * 4 + 15) & ~(size_t)15));
    if (p.insert_mode == 6)
        for (int i = tid; i < p.tables_per_cta * p.nbc; i += T) s_cur[i] = 0;
    __syncthreads();
    const int64_t local = (int64_t)blockIdx.x * T + tid;
    if (local >= p.n_chains) return;
    const int64_t gchain = p.chain_offset + local;
    const int n_eq = p.gcode.neq;
    const int64_t tab = local / p.droplets;
    const int eq = (int)(tab % n_eq);
    const int64_t sw = tab / n_eq;
    PackedLat lat{tile + tid, T, q.nw, q.hb, q.k};
    {
        for (int w = q.nw; w < q.nwp; w++) tile[(size_t)w * T + tid] = 0u;
        const uint64_t *src = reinterpret_cast<const uint64_t *>(p.lat0) + (p.per_class ? tab : sw) * g.nw;
        for (int w = 0; w < g.nw; w++) lat.set(w, src[w]);
    }
    const uint32_t cl = (uint32_t)gchain, chh = (uint32_t)((uint64_t)gchain >> 32);
    if (!p.per_class) to_class_rt<uint64_t>(p.gcode, lat, eq);
    if (p.randomize) {   // apply_stabilizers_uniform: the draw schedule of stdc_fast_kernel
        const int L = g.L;
        uint4 r = make_uint4(0, 0, 0, 0);
        int i = 0;
        for (int o = 0; o < 2; o++)
            for (int rr = 0; rr < L; rr++)
                for (int c = 0; c < L; c++, i++) {
                    if ((i & 127) == 0) r = philox4x32_10((uint32_t)(i >> 7), 0x80000000u, cl, chh, keys);
                    const int wi = (i >> 5) & 3;
                    const uint32_t word = wi == 0 ? r.x : wi == 1 ? r.y : wi == 2 ? r.z : r.w;
                    const bool hit = (word >> (i & 31)) & 1;
                    if (hit && rain_legal_rt(p.gcode, o, rr, c)) apply_rco_rt<uint64_t>(p.gcode, lat, rr, c, o == 0 ? 3 : 1);
                }
    }
    int n = lat_weight<uint64_t>(g, lat);
    uint64_t h = lat_hash<uint64_t>(g, lat, p.hash_seed);
    const uint64_t cap_mask = p.cap_mask;
    // 6: bucket logs of the chain's table (one pass of bucket_dedupe_kernel afterwards); 4: this chain's key log; 2: the class's
    // set in HBM, probed one sample late
    const int imode = MODE == MODE_MEAN ? 3 : p.insert_mode;
    unsigned long long *table = imode == 4 ? p.logs + (uint64_t)local * (uint64_t)p.log_cap : p.tables + (uint64_t)tab * (cap_mask + 1);
    const uint32_t cur_base = (uint32_t)__cvta_generic_to_shared(s_cur) + (uint32_t)(tid / p.droplets) * ((uint32_t)p.nbc * 4u);
    unsigned long long *blog_tab = imode == 6 ? p.blogs + (uint64_t)tab * (uint64_t)p.nbc * p.bcap : nullptr;
    uint32_t nacc = 0, noff = 0, nacc_seen = 0xFFFFFFFFu;
    int left = p.iters;
    unsigned long long key = 0;
    uint32_t slot = 0;
    const uint32_t smask = (uint32_t)cap_mask;
    SampleAcct<MODE> acct;
    acct.init(p, tab);

    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    unsigned char *mybase = reinterpret_cast<unsigned char *>(tile + tid);
    // table bases (this lane's copy of the replicated ones), opaque to the compiler so that it keeps them in registers instead
    // of rebuilding them from the thread index at every use (ten issue slots per step when left to itself)
    uint32_t aA, aThr, aB, aHs, aDE;
    asm volatile("mov.u32 %0, %1;" : "=r"(aA) : "r"(sbase + lay.a + (uint32_t)(tid & (REP - 1)) * 8u));
    asm volatile("mov.u32 %0, %1;" : "=r"(aThr) : "r"(sbase + lay.thr + (uint32_t)(tid & (REP - 1)) * 4u));
    asm volatile("mov.u32 %0, %1;" : "=r"(aB) : "r"(sbase + lay.b));
    asm volatile("mov.u32 %0, %1;" : "=r"(aHs) : "r"(sbase + lay.hs));
    asm volatile("mov.u32 %0, %1;" : "=r"(aDE) : "r"(sbase + lay.de));
    const uint32_t ws = (uint32_t)T * 4u;   // bytes between consecutive words of a chain

    auto step = [&](uint32_t idx, uint32_t r_acc) {
        const uint2 S = lds_r64(aA + idx * (8u * REP));
        uint32_t *pa = reinterpret_cast<uint32_t *>(mybase + (S.x & 0xFFu) * ws), *pb = reinterpret_cast<uint32_t *>(mybase + __byte_perm(S.x, 0, 0x4441) * ws);
        uint32_t *pc = reinterpret_cast<uint32_t *>(mybase + __byte_perm(S.x, 0, 0x4442) * ws), *pd = reinterpret_cast<uint32_t *>(mybase + (S.x >> 24) * ws);
        const uint32_t wa = *pa, wb = *pb, wc = *pc, wd = *pd;
        // rotate each word so that its field lands at bits [2 i, 2 i + 2) of slot i (the funnel shift reads five bits of its count)
        const uint32_t ta = __funnelshift_r(wa, wa, S.y);
        const uint32_t tb = __funnelshift_r(wb, wb, S.y >> 5);
        const uint32_t tc = __funnelshift_r(wc, wc, S.y >> 10);
        const uint32_t td = __funnelshift_r(wd, wd, S.y >> 15);
        const uint32_t s1 = (ta & 0x03u) | (tb & ~0x03u);
        const uint32_t s2 = (tc & 0x30u) | (td & ~0x30u);
        const uint32_t f = (s1 & 0x0Fu) | (s2 & ~0x0Fu);
        const uint32_t li = (f & 0xFFu) | (S.y >> 20);
        if (r_acc <= lds_r32(aThr + li * (4u * REP))) {
            const uint4 M = lds_r128(aB + idx * 16u);
            *pa = wa ^ M.x;
            *pb = wb ^ M.y;
            *pc = wc ^ M.z;
            *pd = wd ^ M.w;
            n += lds_rs8(aDE + li);
            const uint2 hv = lds_r64(aHs + idx * 8u);
            h ^= (uint64_t)hv.x | ((uint64_t)hv.y << 32);
            nacc++;
        }
        if (--left == 0) {
            left = p.iters;
            acct.sample(n);
            const bool dirty = nacc != nacc_seen;
            if (imode == 6) {
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
                if (key) {
                    unsigned long long qv;
                    do {
                        qv = atomicCAS(table + slot, 0ull, key);
                        slot = (slot + 1u) & smask;
                    } while (qv != 0ull && qv != key);
                    key = 0;
                }
                if (dirty) {
                    key = make_key(h, n);
                    slot = (uint32_t)(key >> QECMC_LEN_BITS) & smask;
                    prefetch_l2(table + slot);
                }
            }
            noff += dirty;
            nacc_seen = nacc;
        }
    };

    const uint64_t tsteps = (uint64_t)p.steps * (uint64_t)p.iters;
    const uint32_t ncalls = (uint32_t)(tsteps >> 1);   // host guarantees tsteps < 2^32
    const uint32_t nstab = (uint32_t)g.nstab;
    for (uint32_t c0 = 0; c0 < ncalls; c0++) {
        const uint4 r = philox4x32_10(c0, 0u, cl, chh, keys);
        step(__umulhi(r.x, nstab), r.y);
        step(__umulhi(r.z, nstab), r.w);
    }
    if (tsteps & 1) {
        const uint4 r = philox4x32_10(ncalls, 0u, cl, chh, keys);
        step(__umulhi(r.x, nstab), r.y);
    }
    if (imode == 2 && key) {
        unsigned long long qv;
        do {
            qv = atomicCAS(table + slot, 0ull, key);
            slot = (slot + 1u) & smask;
        } while (qv != 0ull && qv != key);
    }
    if (imode == 6) {
        __syncthreads();   // every chain of the CTA has logged its last key (threads that exited above do not count)
        const int64_t tab0 = (int64_t)blockIdx.x * p.tables_per_cta;
        const int64_t rest = p.n_chains - (int64_t)blockIdx.x * T;
        const int nact = rest < T ? (int)rest : T;   // threads still here
        for (int i = tid; i < p.tables_per_cta * p.nbc; i += nact)
            if (tab0 + i / p.nbc < p.n_chains / p.droplets) p.bcounts[tab0 * p.nbc + i] = min(s_cur[i], p.bcap);
    }
    if (imode == 4) p.log_counts[local] = noff;
    acct.finish(p, local);
    atomicAdd(p.counters + 0, (unsigned long long)nacc);
    atomicAdd(p.counters + 1, (unsigned long long)noff);
    atomicAdd(p.steps_done, (unsigned long long)tsteps);
}

}  // namespace qecmc

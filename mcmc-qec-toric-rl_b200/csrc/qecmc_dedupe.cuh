// qecmc_dedupe.cuh -- distinct-chain counting from per-chain key logs.
//
// The asynchronous open-addressing set in HBM costs one random 64-byte DRAM read plus one 32-byte write-back per
// offered key; at the headline configuration that is 5.4e9 row activations per launch and the chains end up waiting
// for them (ncu: long-scoreboard is the top stall, DRAM at 19 % of its bandwidth).  In log mode a chain instead
// appends each offered key to its own log -- a plain store, sequential per chain, merged into full sectors by L2 --
// and this kernel turns the logs of one (syndrome, class) into N(n) afterwards, touching HBM only in streams:
//   A  bucket histogram of the table's keys (low fingerprint bits)
//   B  scatter into the CTA's scratch area, bucket by bucket
//   C  per bucket: insert into a shared-memory hash set; a key new to the set bumps N(len)
// Buckets take ~1024 logged keys; the set of a bucket has 4096 slots.
// One CTA owns one table at a time, so there is no inter-CTA synchronisation.  Same semantics as the global set:
// the distinct keys of all droplets of the class (STDC_droplet / STDC, decoders.py:236-322).
#pragma once
#include "qecmc_device.cuh"

namespace qecmc {

#define QECMC_DD_THREADS 1024
#define QECMC_DD_GROUPS 4           // independent groups of 256 threads in step C, one bucket each at a time
#define QECMC_DD_HASH_SLOTS 4096    // per group; four sets = 128 KiB of shared memory
#define QECMC_DD_MAX_BUCKETS 4096
#define QECMC_DD_BUCKET_TARGET 1024 // logged keys per bucket aimed for (distinct keys load the set to <= 0.25 on average)
#define QECMC_DD_KPT 4              // keys per thread held in registers: one chunk = 1024 keys per group
#define QECMC_DD_SEG_PAIRS 512      // a warp streams a chain's log in segments of 512 16-byte pairs
#define QECMC_DD_MAX_CNT 1024       // chain counts staged in shared memory up to this many droplets

struct DedupeParams {
    const unsigned long long *logs;  // [chains][log_cap]
    const uint32_t *log_counts;      // [chains]
    int64_t log_cap;                 // entries per chain, even
    int droplets;
    int64_t tabs;                    // (syndrome, class) tables of this wave
    unsigned long long *scratch;     // [gridDim.x][scratch_cap]
    int64_t scratch_cap;
    int nsites;
    double beta;
    double *Z;                       // [tabs]
    uint32_t *N_hist;                // [tabs][nsites + 1], optional
    unsigned long long *distinct;    // += distinct keys
    int *err;                        // 1: a shared-memory set filled up, 2: scratch too small
};

// insert into a shared-memory set of mask + 1 slots; a key new to the set bumps N(len)
__device__ __forceinline__ void dd_insert(unsigned long long *hash, uint32_t mask, int shift, unsigned long long key, uint32_t *hist, int *err)
{
    uint32_t slot = (uint32_t)(key >> shift) & mask;
    for (uint32_t tries = 0; tries <= mask; tries++) {
        unsigned long long prev = atomicCAS(hash + slot, 0ull, key);
        if (prev == 0ull) { atomicAdd(hist + (uint32_t)(key & QECMC_LEN_MASK), 1u); return; }
        if (prev == key) return;
        slot = (slot + 1u) & mask;
    }
    *err = 1;
}

// Visit every logged key of table `tab`.  Work items are (chain, segment of 512 pairs), dealt to the warps round-robin;
// a lane keeps four 16-byte loads in flight.  s_cnt: the chains' key counts (shared memory), max_cnt their maximum.
template <typename F>
__device__ __forceinline__ void dd_for_each_key(const DedupeParams &p, int64_t tab, const uint32_t *s_cnt, uint32_t max_cnt, F f)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t segs = (max_cnt / 2 + QECMC_DD_SEG_PAIRS - 1) / QECMC_DD_SEG_PAIRS + 1;   // +1: the odd tail item
    const uint32_t items = (uint32_t)p.droplets * segs;
    for (uint32_t t = warp; t < items; t += QECMC_DD_THREADS / 32) {
        const uint32_t d = t / segs, sg = t - d * segs;
        const int64_t chain = tab * p.droplets + d;
        const uint32_t c = d < QECMC_DD_MAX_CNT ? s_cnt[d] : p.log_counts[chain];
        const uint32_t pairs = c >> 1;
        if (sg == segs - 1) {   // tail item: the last key of an odd-length log
            if ((c & 1u) && lane == 0) f(p.logs[chain * p.log_cap + c - 1]);
            continue;
        }
        const uint32_t lo = sg * QECMC_DD_SEG_PAIRS;
        if (lo >= pairs) continue;
        const uint32_t hi = min(pairs, lo + QECMC_DD_SEG_PAIRS);
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(p.logs + chain * p.log_cap);
        uint32_t i = lo + lane;
        for (; i + 96 < hi; i += 128) {
            ulonglong2 a0 = __ldcs(src + i), a1 = __ldcs(src + i + 32), a2 = __ldcs(src + i + 64), a3 = __ldcs(src + i + 96);
            f(a0.x); f(a0.y); f(a1.x); f(a1.y); f(a2.x); f(a2.y); f(a3.x); f(a3.y);
        }
        for (; i < hi; i += 32) {
            ulonglong2 a = __ldcs(src + i);
            f(a.x); f(a.y);
        }
    }
}

__global__ void __launch_bounds__(QECMC_DD_THREADS, 1) log_dedupe_kernel(DedupeParams p)
{
    constexpr int T = QECMC_DD_THREADS, HS = QECMC_DD_HASH_SLOTS, KPT = QECMC_DD_KPT, GT = QECMC_DD_THREADS / QECMC_DD_GROUPS;
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned long long *hash = reinterpret_cast<unsigned long long *>(dsm);   // one set of HS slots per group
    uint32_t *hist = reinterpret_cast<uint32_t *>(hash + QECMC_DD_GROUPS * HS);   // [4096] N(len)
    uint32_t *bcnt = hist + 4096;                     // keys per bucket
    uint32_t *boff = bcnt + QECMC_DD_MAX_BUCKETS;     // exclusive prefix, then scatter cursor (= bucket end afterwards)
    uint32_t *s_cnt = boff + QECMC_DD_MAX_BUCKETS;    // [QECMC_DD_MAX_CNT] keys per chain
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_max;
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x;
    unsigned long long *scratch = p.scratch + (uint64_t)blockIdx.x * p.scratch_cap;
    const int nh = p.nsites + 1;

    for (int64_t tab = blockIdx.x; tab < p.tabs; tab += gridDim.x) {
        if (tid == 0) { s_total = 0; s_max = 0; }
        for (int i = tid; i < nh; i += T) hist[i] = 0;
        for (int i = tid; i < HS; i += T) hash[i] = 0ull;   // the single-set path below uses group 0's set
        __syncthreads();
        {
            unsigned long long loc = 0;
            uint32_t mx = 0;
            for (int d = tid; d < p.droplets; d += T) {
                uint32_t c = p.log_counts[tab * p.droplets + d];
                if (d < QECMC_DD_MAX_CNT) s_cnt[d] = c;
                loc += c;
                mx = max(mx, c);
            }
            if (loc) { atomicAdd(&s_total, loc); atomicMax(&s_max, mx); }
        }
        __syncthreads();
        const uint64_t total = s_total;
        const uint32_t max_cnt = s_max;
        int lg = 0;
        while (lg < 12 && ((uint64_t)QECMC_DD_BUCKET_TARGET << lg) < total) lg++;
        const uint32_t NB = 1u << lg;
        const int shift = QECMC_LEN_BITS + lg;       // set slots come from the fingerprint bits above the bucket bits
        if (lg == 0) {
            // everything fits one shared-memory set: insert straight from the logs
            dd_for_each_key(p, tab, s_cnt, max_cnt, [&](unsigned long long k) { dd_insert(hash, HS - 1, shift, k, hist, p.err); });
        } else if (total > (uint64_t)p.scratch_cap) {
            if (tid == 0) *p.err = 2;
        } else {
            for (uint32_t i = tid; i < NB; i += T) bcnt[i] = 0;
            __syncthreads();
            // A: bucket histogram
            dd_for_each_key(p, tab, s_cnt, max_cnt, [&](unsigned long long k) { atomicAdd(&bcnt[(uint32_t)(k >> QECMC_LEN_BITS) & (NB - 1)], 1u); });
            __syncthreads();
            {   // exclusive scan of bcnt -> boff (NB <= 4096 = 4 per thread)
                const uint32_t per = (NB + T - 1) / T;   // 1..4
                uint32_t v[4], sum = 0;
                for (uint32_t j = 0; j < per; j++) { uint32_t idx = tid * per + j; v[j] = idx < NB ? bcnt[idx] : 0; sum += v[j]; }
                uint32_t inc = sum;
                for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if ((tid & 31) >= o) inc += t; }
                if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
                __syncthreads();
                if (tid < 32) {
                    uint32_t w = s_warp[tid], winc = w;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (tid >= o) winc += t; }
                    s_warp[tid] = winc - w;
                }
                __syncthreads();
                uint32_t run = s_warp[tid >> 5] + inc - sum;
                for (uint32_t j = 0; j < per; j++) { uint32_t idx = tid * per + j; if (idx < NB) boff[idx] = run; run += v[j]; }
            }
            __syncthreads();
            // B: scatter into the CTA's scratch area, bucket by bucket
            dd_for_each_key(p, tab, s_cnt, max_cnt, [&](unsigned long long k) {
                uint32_t pos = atomicAdd(&boff[(uint32_t)(k >> QECMC_LEN_BITS) & (NB - 1)], 1u);
                scratch[pos] = k;
            });
            __syncthreads();   // boff[b] is now the END of bucket b; its keys are visible to the whole CTA
            // C: four groups of 256 threads work on a bucket each (named barriers, so a group waiting for its slowest
            // warp leaves the issue slots to the others); a group loads its next bucket's first chunk into registers
            // before it inserts the current one
            const int grp = tid / GT, gt = tid % GT;
            unsigned long long *hb = hash + grp * HS;
            unsigned long long k[KPT], kn[KPT];
            if ((uint32_t)grp < NB) {
                const uint32_t nb0 = bcnt[grp];
                const unsigned long long *src0 = scratch + (boff[grp] - nb0);
#pragma unroll
                for (int j = 0; j < KPT; j++) { uint32_t idx = j * GT + gt; k[j] = idx < nb0 ? src0[idx] : 0ull; }
            }
            for (uint32_t b = grp; b < NB; b += QECMC_DD_GROUPS) {
                const uint32_t nb = bcnt[b];
                const unsigned long long *src = scratch + (boff[b] - nb);
                uint32_t cap = 256;                         // set sized to the bucket: load <= 0.5, fewer slots to clear
                while (cap < 2 * nb && cap < (uint32_t)HS) cap <<= 1;
                for (uint32_t i = gt; i < cap; i += GT) hb[i] = 0ull;
                if (b + QECMC_DD_GROUPS < NB) {
                    const uint32_t nb1 = bcnt[b + QECMC_DD_GROUPS];
                    const unsigned long long *src1 = scratch + (boff[b + QECMC_DD_GROUPS] - nb1);
#pragma unroll
                    for (int j = 0; j < KPT; j++) { uint32_t idx = j * GT + gt; kn[j] = idx < nb1 ? src1[idx] : 0ull; }
                }
                asm volatile("bar.sync %0, %1;" : : "r"(1 + grp), "r"(GT) : "memory");
#pragma unroll
                for (int j = 0; j < KPT; j++)
                    if (k[j]) dd_insert(hb, cap - 1, shift, k[j], hist, p.err);
                for (uint32_t idx = KPT * GT + gt; idx < nb; idx += GT) dd_insert(hb, cap - 1, shift, src[idx], hist, p.err);   // long lists of repeats
#pragma unroll
                for (int j = 0; j < KPT; j++) k[j] = kn[j];
                asm volatile("bar.sync %0, %1;" : : "r"(1 + grp), "r"(GT) : "memory");
            }
        }
        __syncthreads();
        if (p.N_hist)
            for (int i = tid; i < nh; i += T) p.N_hist[(uint64_t)tab * nh + i] = hist[i];
        if (tid == 0) {
            double z = 0;
            unsigned long long cnt = 0;
            for (int n = 0; n < nh; n++)
                if (hist[n]) { z += (double)hist[n] * exp(-p.beta * (double)n); cnt += hist[n]; }
            p.Z[tab] = z;
            if (p.distinct) atomicAdd(p.distinct, cnt);
        }
        __syncthreads();
    }
}

// ---- insert mode 6: the chain kernel already split every table's keys into coarse bucket logs, so the
// reduction is a single pass: one CTA per table, bucket after bucket into one shared-memory set of 16384 slots; the next
// bucket's keys are loaded into registers while the current one is inserted.  Keys that found their bucket log full sit in
// the table's overflow log and are picked up by scanning it for every bucket (rare, bounded by the overflow capacity).
#define QECMC_BD_THREADS 1024
#define QECMC_BD_SLOTS 16384
#define QECMC_BD_KPT 8

struct BucketDedupeParams {
    int nbc;                           // coarse buckets per table (power of two)
    const unsigned long long *blogs;   // [tabs][nbc][bcap]
    const uint32_t *bcounts;           // [tabs][nbc]
    uint32_t bcap;
    const unsigned long long *ovf;     // [tabs][ovf_cap]
    const uint32_t *ovf_cnt;           // [tabs]
    uint32_t ovf_cap;
    int64_t tabs;
    int nsites;
    double beta;
    double *Z;
    uint32_t *N_hist;
    unsigned long long *distinct;
    int *err;
};

__global__ void __launch_bounds__(QECMC_BD_THREADS, 1) bucket_dedupe_kernel(BucketDedupeParams p)
{
    constexpr int T = QECMC_BD_THREADS, HS = QECMC_BD_SLOTS, KPT = QECMC_BD_KPT;
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned long long *hash = reinterpret_cast<unsigned long long *>(dsm);
    uint32_t *hist = reinterpret_cast<uint32_t *>(hash + HS);   // [4096]
    const int tid = threadIdx.x;
    const int nh = p.nsites + 1;
    int lg = 0;
    while ((1 << lg) < p.nbc) lg++;
    const int shift = QECMC_LEN_BITS + lg;   // set slots from the fingerprint bits above the coarse-bucket bits
    for (int64_t tab = blockIdx.x; tab < p.tabs; tab += gridDim.x) {
        for (int i = tid; i < nh; i += T) hist[i] = 0;
        const uint32_t *cnt = p.bcounts + tab * p.nbc;
        const unsigned long long *logs = p.blogs + (uint64_t)tab * (uint64_t)p.nbc * p.bcap;
        const uint32_t novf = min(p.ovf_cnt[tab], p.ovf_cap);
        const unsigned long long *ovf = p.ovf + (uint64_t)tab * p.ovf_cap;
        unsigned long long k[KPT], kn[KPT];
        {
            const uint32_t nb0 = cnt[0];
#pragma unroll
            for (int j = 0; j < KPT; j++) { uint32_t idx = j * T + tid; k[j] = idx < nb0 ? __ldcs(logs + idx) : 0ull; }
        }
        for (int b = 0; b < p.nbc; b++) {
            const uint32_t nb = cnt[b];
            const unsigned long long *src = logs + (uint64_t)b * p.bcap;
            uint32_t cap = 256;
            while (cap < 2 * (nb + novf) && cap < (uint32_t)HS) cap <<= 1;
            for (uint32_t i = tid; i < cap; i += T) hash[i] = 0ull;
            if (b + 1 < p.nbc) {
                const uint32_t nb1 = cnt[b + 1];
                const unsigned long long *src1 = src + p.bcap;
#pragma unroll
                for (int j = 0; j < KPT; j++) { uint32_t idx = j * T + tid; kn[j] = idx < nb1 ? __ldcs(src1 + idx) : 0ull; }
            }
            __syncthreads();
            if (nb + novf > (uint32_t)(HS * 0.8)) {
                if (tid == 0) *p.err = 1;   // would not fit the set: the host redoes the call with per-chain logs
            } else {
#pragma unroll
                for (int j = 0; j < KPT; j++)
                    if (k[j]) dd_insert(hash, cap - 1, shift, k[j], hist, p.err);
                for (uint32_t idx = KPT * T + tid; idx < nb; idx += T) dd_insert(hash, cap - 1, shift, __ldcs(src + idx), hist, p.err);
                for (uint32_t idx = tid; idx < novf; idx += T) {
                    const unsigned long long kk = ovf[idx];
                    if (((uint32_t)(kk >> QECMC_LEN_BITS) & (uint32_t)(p.nbc - 1)) == (uint32_t)b) dd_insert(hash, cap - 1, shift, kk, hist, p.err);
                }
            }
#pragma unroll
            for (int j = 0; j < KPT; j++) k[j] = kn[j];
            __syncthreads();
        }
        if (p.N_hist)
            for (int i = tid; i < nh; i += T) p.N_hist[(uint64_t)tab * nh + i] = hist[i];
        if (tid == 0) {
            double z = 0;
            unsigned long long c = 0;
            for (int n = 0; n < nh; n++)
                if (hist[n]) { z += (double)hist[n] * exp(-p.beta * (double)n); c += hist[n]; }
            p.Z[tab] = z;
            if (p.distinct) atomicAdd(p.distinct, c);
        }
        __syncthreads();
    }
}

}  // namespace qecmc

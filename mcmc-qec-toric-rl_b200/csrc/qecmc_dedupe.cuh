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
// One CTA owns one table at a time, so there is no inter-CTA synchronisation.  Same semantics as the global set:
// the distinct keys of all droplets of the class (STDC_droplet / STDC, decoders.py:236-322).
#pragma once
#include "qecmc_device.cuh"

namespace qecmc {

#define QECMC_DD_THREADS 1024
#define QECMC_DD_HASH_SLOTS 16384   // 128 KiB of shared memory
#define QECMC_DD_MAX_BUCKETS 4096
#define QECMC_DD_BUCKET_TARGET 6144 // keys per bucket aimed for: load <= 0.5 with room for fluctuation

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
    int *err;                        // set if a bucket cannot fit the shared-memory set
};

__device__ __forceinline__ void dd_insert(unsigned long long *hash, uint32_t mask, int shift, unsigned long long key, uint32_t *hist)
{
    uint32_t slot = (uint32_t)(key >> shift) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(hash + slot, 0ull, key);
        if (prev == 0ull) { atomicAdd(hist + (uint32_t)(key & QECMC_LEN_MASK), 1u); return; }
        if (prev == key) return;
        slot = (slot + 1u) & mask;
    }
}

// visit every logged key of table `tab`: chains one after the other, 16-byte loads, two per thread in flight
template <typename F> __device__ __forceinline__ void dd_for_each_key(const DedupeParams &p, int64_t tab, F f)
{
    const int tid = threadIdx.x;
    for (int d = 0; d < p.droplets; d++) {
        const int64_t chain = tab * p.droplets + d;
        const uint32_t c = p.log_counts[chain];
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(p.logs + chain * p.log_cap);
        const uint32_t pairs = c >> 1;
        uint32_t i = tid;
        for (; i + QECMC_DD_THREADS < pairs; i += 2 * QECMC_DD_THREADS) {
            ulonglong2 a = __ldcs(src + i), b = __ldcs(src + i + QECMC_DD_THREADS);
            f(a.x); f(a.y); f(b.x); f(b.y);
        }
        for (; i < pairs; i += QECMC_DD_THREADS) {
            ulonglong2 a = __ldcs(src + i);
            f(a.x); f(a.y);
        }
        if ((c & 1u) && tid == 0) f(p.logs[chain * p.log_cap + c - 1]);
    }
}

__global__ void __launch_bounds__(QECMC_DD_THREADS, 1) log_dedupe_kernel(DedupeParams p)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned long long *hash = reinterpret_cast<unsigned long long *>(dsm);
    uint32_t *hist = reinterpret_cast<uint32_t *>(hash + QECMC_DD_HASH_SLOTS);
    uint32_t *bcnt = hist + 4096;                     // keys per bucket
    uint32_t *boff = bcnt + QECMC_DD_MAX_BUCKETS;     // exclusive prefix, then scatter cursor
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x;
    unsigned long long *scratch = p.scratch + (uint64_t)blockIdx.x * p.scratch_cap;
    const int nh = p.nsites + 1;

    for (int64_t tab = blockIdx.x; tab < p.tabs; tab += gridDim.x) {
        if (tid == 0) s_total = 0;
        for (int i = tid; i < nh; i += QECMC_DD_THREADS) hist[i] = 0;
        __syncthreads();
        {
            unsigned long long loc = 0;
            for (int d = tid; d < p.droplets; d += QECMC_DD_THREADS) loc += p.log_counts[tab * p.droplets + d];
            if (loc) atomicAdd(&s_total, loc);
        }
        __syncthreads();
        const uint64_t total = s_total;
        int lg = 0;
        while (lg < 12 && ((uint64_t)QECMC_DD_BUCKET_TARGET << lg) < total) lg++;
        const uint32_t NB = 1u << lg;
        const int shift = QECMC_LEN_BITS + lg;       // set slots come from the fingerprint bits above the bucket bits
        if (lg == 0) {
            // everything fits one shared-memory set: insert straight from the logs
            uint32_t cap = 64;
            while (cap < 2 * total && cap < QECMC_DD_HASH_SLOTS) cap <<= 1;
            if (total > (uint64_t)(QECMC_DD_HASH_SLOTS * 0.85)) { if (tid == 0) *p.err = 1; }
            else {
                for (uint32_t i = tid; i < cap; i += QECMC_DD_THREADS) hash[i] = 0ull;
                __syncthreads();
                dd_for_each_key(p, tab, [&](unsigned long long k) { dd_insert(hash, cap - 1, shift, k, hist); });
            }
        } else {
            for (uint32_t i = tid; i < NB; i += QECMC_DD_THREADS) bcnt[i] = 0;
            __syncthreads();
            dd_for_each_key(p, tab, [&](unsigned long long k) { atomicAdd(&bcnt[(uint32_t)(k >> QECMC_LEN_BITS) & (NB - 1)], 1u); });
            __syncthreads();
            {   // exclusive scan of bcnt -> boff (NB <= 4096 = 4 per thread)
                const uint32_t per = (NB + QECMC_DD_THREADS - 1) / QECMC_DD_THREADS;   // 1..4
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
            if (total > (uint64_t)p.scratch_cap) { if (tid == 0) *p.err = 2; __syncthreads(); continue; }
            dd_for_each_key(p, tab, [&](unsigned long long k) {
                uint32_t pos = atomicAdd(&boff[(uint32_t)(k >> QECMC_LEN_BITS) & (NB - 1)], 1u);
                scratch[pos] = k;
            });
            __syncthreads();   // boff[b] is now the END of bucket b; its keys are visible to the whole CTA
            for (uint32_t b = 0; b < NB; b++) {
                const uint32_t nb = bcnt[b], end = boff[b];
                if (nb == 0) continue;
                if (nb > (uint32_t)(QECMC_DD_HASH_SLOTS * 0.85)) { if (tid == 0) *p.err = 1; continue; }
                uint32_t cap = 64;
                while (cap < 2 * nb && cap < QECMC_DD_HASH_SLOTS) cap <<= 1;
                for (uint32_t i = tid; i < cap; i += QECMC_DD_THREADS) hash[i] = 0ull;
                __syncthreads();
                const unsigned long long *src = scratch + (end - nb);
                for (uint32_t i = tid; i < nb; i += QECMC_DD_THREADS) dd_insert(hash, cap - 1, shift, src[i], hist);
                __syncthreads();
            }
        }
        __syncthreads();
        if (p.N_hist)
            for (int i = tid; i < nh; i += QECMC_DD_THREADS) p.N_hist[(uint64_t)tab * nh + i] = hist[i];
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

}  // namespace qecmc

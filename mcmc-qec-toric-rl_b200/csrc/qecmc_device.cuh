// qecmc_device.cuh -- device-side building blocks shared by the chain kernels:
// Philox4x32-10, the shared-memory lattice accessor, proposal decoding from either
// Philox words (native) or the reference's uniform draws (replay), and the
// open-addressing distinct-chain set in global memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "qecmc_lattice.h"

namespace qecmc {

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (call index lo/hi, chain id lo/hi),
// key = 64-bit seed.  One call yields four 32-bit words = two Metropolis steps
// (proposal word + accept word each).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Same generator with the ten round keys precomputed on the host (kernel parameters live in
// the constant bank, so each key is a free LOP3 operand) and mul.wide issued explicitly.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
__device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
    uint64_t p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &k)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0, l0, h1, l1;
        mulhilo32(0xD2511F53u, c0, h0, l0);
        mulhilo32(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ k.k0[r];
        c2 = h0 ^ c3 ^ k.k1[r];
        c1 = l1;
        c3 = l0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// This thread's lattice inside the block's shared-memory tile, laid out [word][thread]
// so that a warp's accesses to any word index hit 32 distinct banks.
template <typename W> struct SmemLat {
    W *base;
    int stride;
    __device__ __forceinline__ W get(int w) const { return base[w * stride]; }
    __device__ __forceinline__ void set(int w, W v) { base[w * stride] = v; }
};

// Acceptance thresholds for dE in [-4*?, ...]: the chain kernels index [dE + THR_OFF].
// A toric/planar/rotated/XZZX stabilizer changes the weight by at most 4.
#define QECMC_THR_OFF 4
#define QECMC_THR_N 9

// Proposal from the reference's uniform draws, in its order and arithmetic
// (toric_model.py:287-296, planar_model.py:342-352, rotated_surface_model.py:395-408,
// xzzx_model.py:439-452).  u points at the k draws of this step.
template <int GEOM> struct NumDraws { static constexpr int value = (GEOM == TORIC || GEOM == PLANAR) ? 3 : 5; };

template <int GEOM> __device__ __forceinline__ void propose_replay(const Geo &g, const double *u, int &row, int &col, int &op)
{
    const int L = g.L;
    if (GEOM == TORIC) {
        row = (int)(u[0] * L);
        col = (int)(u[1] * L);
        int o = (int)(u[2] * 2);
        op = o == 0 ? 3 : o;
    } else if (GEOM == PLANAR) {
        int s = (int)((L - 1) * u[0]);
        int l = (int)(L * u[1]);
        if (u[2] < 0.5) { row = s; col = l; op = 1; }
        else { row = l; col = s; op = 3; }
    } else {
        int rows = (int)((L - 1) * u[0]);
        int cols = (int)((L - 1) * u[1]);
        int rows2 = (int)(((double)(L - 1) / 2.0) * u[2]);
        int cols2 = (int)(4 * u[3]);
        double phalf = (double)(L * L - (L - 1) * (L - 1) - 1) / (double)(L * L - 1);
        if (u[4] > phalf) { row = rows; col = cols; op = 1; }
        else { row = rows2; col = cols2; op = 3; }
    }
}

// ---------------------------------------------------------------------------
// Distinct-chain set: one open-addressing table of 64-bit keys per (syndrome, class),
// shared by that class's droplets (the union of decoders.py:313-314).  key =
// fingerprint bits [12,63) | chain length in the low 12 bits | bit 63 (so 0 = empty).
// Keys are insert-only, so a stale "empty" read is resolved by the CAS.
// ---------------------------------------------------------------------------
#define QECMC_LEN_BITS 12
#define QECMC_LEN_MASK 0xFFFull

__device__ __forceinline__ uint64_t make_key(uint64_t h, int n)
{
    return (h & ~QECMC_LEN_MASK) | (uint64_t)n | (1ull << 63);
}

__device__ __forceinline__ void table_prefetch(const unsigned long long *tab, uint64_t mask, uint64_t key)
{
    const unsigned long long *p = tab + ((key >> QECMC_LEN_BITS) & mask);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// returns true when the key was not present before
__device__ __forceinline__ bool table_insert(unsigned long long *tab, uint64_t mask, uint64_t key)
{
    uint64_t slot = (key >> QECMC_LEN_BITS) & mask;
    while (true) {
        unsigned long long cur = __ldcg(tab + slot);
        if (cur == key) return false;
        if (cur == 0ull) {
            unsigned long long prev = atomicCAS(tab + slot, 0ull, (unsigned long long)key);
            if (prev == 0ull) return true;
            if (prev == key) return false;
        }
        slot = (slot + 1) & mask;
    }
}

}  // namespace qecmc

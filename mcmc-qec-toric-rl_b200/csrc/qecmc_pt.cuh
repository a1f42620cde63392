// qecmc_pt.cuh -- parallel-tempering ladders, native draws: the rung-major kernel.
//
// Reference: Ladder / Ladder_alpha / Ladder_biased .step (src/mcmc.py:94-103, src/mcmc_alpha.py:126-137,
// src/mcmc_biased.py:115-124) over Chain / Chain_alpha / Chain_biased .update_chain (mcmc.py:19-43, mcmc_alpha.py:27-70,
// mcmc_biased.py:21-59), and PTEQ's bookkeeping (decoders.py:25-105, decoders_biasednoise.py:28-237).
//
// The replay kernel (qecmc_ladder.cuh) binds a ladder to a warp and a replica to a lane: a 25-rung ladder idles 7 lanes, the
// top rung's logical moves stall the other 31 lanes, and replica-bound state costs 120 registers.  Here the THREAD is bound
// to the RUNG and the lattice to a fixed shared-memory column:
//   * a CTA holds NLC ladders (32 with 32-bit row words, 16 with 64-bit ones); a rung warp = 32 / NLC rungs x NLC ladders,
//     so every lane of every rung warp runs a Metropolis step in every iteration and a warp reads one or two rungs'
//     thresholds; replica r of ladder l lives in tile column r * NLC + l for good (bank = ladder: conflict free);
//   * a swap permutes one byte per rung (which column holds the rung's replica), never a lattice and never registers;
//   * the top rung, whose iterations mix stabilizer and logical proposals (mcmc.py:24-33), belongs to separate warps in
//     which LT lanes share one replica: a logical operator's row words are split over the LT lanes, so the O(L) work of a
//     logical move neither stalls a rung warp nor serialises on one lane;
//   * after the iterations: one CTA barrier, then lane l of warp 0 ("manager") walks ladder l's swap sweep on the
//     per-rung weights and largest-swapping-exponent thresholds the rung threads published (one compare and one select per
//     pair), applies the flag / tops0 rules and PTEQ's accounting (class counts, the n_err history, the convergence
//     windows), a second barrier, and every rung thread reads its new column off the swap mask;
//   * a ladder that converged (or reached its step cap) hands its columns to the next pending ladder of the launch
//     (atomic queue), so a CTA is never held up by its slowest ladder and the n_err history is sized by the RESIDENT
//     ladders, not by the batch.
// Draws are positional (Philox counter = (Ladder.step, purpose, rung, iteration), no generator state in registers); the
// schedule is restated in oracle/qec_oracle.c (ladder_step_native) and the tests compare the two bit for bit.
#pragma once
#include "qecmc_ladder.cuh"

namespace qecmc {

struct PtParams {
    Geo g;
    int kind, Nc, iters, acct;
    int lt;                     // lanes sharing one top-rung replica (power of two, 2..32)
    double p_logical;
    int top_accept_all;
    const uint32_t *thr_u;      // [Nc][9]   depolarizing: 32-bit accept thresholds by dE
    const double *thr_top_d;    // [8L+1]    depolarizing top rung: pow(factor_top, dE)
    const double *diff;         // [Nc-1]
    const double *pw;           // [Nc-1][2K+1] diff^k for k = 0 .. 2K (depolarizing / biased swap decisions)
    double alpha;
    const double *wtab;         // weighted kinds: [Nc][4][nsites+1]
    const uint2 *desc2;         // toric / planar depolarizing: stabilizer descriptors
    PhiloxKeys keys;
    int64_t n_ladders, ladder_offset, steps;
    uint32_t step0;             // Ladder.step calls these ladders have already made (resume): offsets the draw positions
    const void *lat_in;
    int init_broadcast;
    const int *flags_in;
    const int2 *neff_in;
    const long long *tops0_in;
    void *lat_out;
    int *flags_out;
    int2 *neff_out;
    long long *tops0_out;
    int SEQ, TOPS, tops_burn, use_conv;
    double eps;
    void *hist;                 // [gridDim.x * NLC][hist_stride]: uint16 n_err per step (uint32 nz | nxy << 16 for alpha ladders)
    int64_t hist_stride;
    long long *eq_counts, *info;
    uint8_t *percent;
    int cls_delta[8];
    unsigned int *queue;        // next ladder of the launch to start; the host sets it to gridDim.x * NLC
    unsigned long long *counters;
};

// manager state per ladder, [field][ladder] in shared memory
enum { PA_ACTIVE = 0, PA_LIDX, PA_STEP, PA_TOPS0, PA_SINCE, PA_BURN, PA_CSTART, PA_CSTREAK, PA_WA, PA_WB, PA_WC,
       PA_S2A, PA_S2A_HI, PA_S2B, PA_S2B_HI, PA_S4A, PA_S4A_HI, PA_S4B, PA_S4B_HI, PA_FIN, PA_NEXT, PA_CONV, PA_BOTCLS, PA_HA, PA_HB, PA_PEND,
       PA_EQC, PA_NF = PA_EQC + 16 };
#define QECMC_PT_FLAG 0x100u
#define QECMC_PT_OPEN (1 << 30)
#define QECMC_PT_BAD (1 << 29)

struct PtLayout {
    size_t tile, state, slot, sn, st, mask, acc, thru, ld, lut, draw, total;
    int NRW, NTW, T;
};

// shared-memory carve-up and CTA shape; identical on host and device
template <typename W> __host__ __device__ inline PtLayout pt_layout(const Geo &g, int Nc, int NLC, int lt, bool has_top, bool table, bool table2, int iters)
{
    PtLayout o;
    const int RPW = 32 / NLC, NR = has_top ? Nc - 1 : Nc;
    o.NRW = (NR + RPW - 1) / RPW;
    if (o.NRW < 1) o.NRW = 1;
    const int TPW = 32 / lt;
    o.NTW = has_top ? (NLC + TPW - 1) / TPW : 0;
    o.T = (o.NRW + o.NTW) * 32;
    const size_t NREP = (size_t)Nc * NLC;
    size_t off = 0;
    o.tile = off; off += ((size_t)g.nw * NREP * sizeof(W) + 15) & ~(size_t)15;
    o.state = off; off += 2 * NREP * 4;
    // eight rungs of padding in front of each sweep array: the manager reads the operands of eight pairs at a time without
    // clamping (the threshold padding holds "swaps whatever is carried", so a pair below rung 0 changes nothing)
    o.sn = off + 8 * (size_t)NLC * 4; off += (8 * (size_t)NLC + NREP) * 4;
    o.st = off + 8 * (size_t)NLC * 4; off += (8 * (size_t)NLC + NREP) * 4;
    o.mask = off; off += 2 * (size_t)NLC * 4;
    o.acc = off; off += (size_t)PA_NF * NLC * 4;
    o.thru = off; off += (((size_t)Nc * 9 * 4) + 7) & ~(size_t)7;
    o.ld = off; off += (table || table2) ? (size_t)g.nstab * 8 : 0;
    o.lut = off; off += table ? 16 * 256 * 2 : table2 ? 512 : 0;
    o.slot = off; off += (2 * NREP + 15) & ~(size_t)15;
    off = (off + 15) & ~(size_t)15;
    o.draw = off; off += has_top ? (size_t)NLC * ((size_t)((iters + 1) >> 1) + 2 * (size_t)iters) * 16 : 0;   // the top rungs' draws of one step
    o.total = off + 16;
    return o;
}

// occupant of rung q after a sweep with swap mask m: the replica one rung below if that pair swapped, otherwise the replica
// that fell past every swapping pair from q upwards (mcmc.py:96-103 walked top to bottom)
__device__ __forceinline__ int pt_src_rung(uint32_t m, int q)
{
    if (q > 0 && ((m >> (q - 1)) & 1u)) return q - 1;
    return q + __ffs((int)~(m >> q)) - 1;
}

// MAXT: the largest CTA the instantiation is launched with.  Two CTAs share an SM whenever they fit, so a CTA of at most 448
// threads (e.g. XZZX d = 21: 10 rung warps + 4 top-rung warps) may use 72 registers per thread instead of 64.
template <int GEOM, typename W, bool WEIGHTED, int NLC, int MAXT = 1024>
__global__ void __launch_bounds__(MAXT, MAXT <= 512 ? 2 : 1) pt_kernel(PtParams p)
{
    constexpr int RPW = 32 / NLC;
    constexpr bool TABLE = GEOM == ROTATED || GEOM == XZZX;
    constexpr bool TABLE2 = (GEOM == TORIC || GEOM == PLANAR) && !WEIGHTED;
    constexpr int NU = NumUpd<GEOM>::value;
    constexpr int NLAY = GEOM == TORIC ? 2 : 1;
    constexpr double U32 = 2.3283064365386963e-10;   // 2^-32
    extern __shared__ __align__(16) unsigned char smem[];
    const Geo g = p.g;
    const int Nc = p.Nc, L = g.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NREP = Nc * NLC, ns1 = g.nsites + 1;
    const bool has_top = p.p_logical != 0.0;
    const int NR = has_top ? Nc - 1 : Nc;
    const PtLayout lay = pt_layout<W>(g, Nc, NLC, p.lt, has_top, TABLE, TABLE2, p.iters);
    W *tile = reinterpret_cast<W *>(smem + lay.tile);
    uint32_t *s_state = reinterpret_cast<uint32_t *>(smem + lay.state);   // [2][NREP]: weights | (class, flag)
    uint8_t *s_slot = smem + lay.slot;                                    // [2][Nc][NLC]: column (replica) on each rung
    int *s_n = reinterpret_cast<int *>(smem + lay.sn);                    // [Nc][NLC] sweep: weight on the rung (alpha: nz of n_eff)
    int *s_t = reinterpret_cast<int *>(smem + lay.st);                    // [Nc][NLC] sweep: swap threshold of pair (r, r+1) (alpha: nx+ny)
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + lay.mask);     // [2][NLC] swap mask of the sweep, by step parity
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(smem + lay.acc);       // [PA_NF][NLC] manager state
    uint32_t *s_thru = reinterpret_cast<uint32_t *>(smem + lay.thru);     // [Nc][9]
    uint2 *s_ld = reinterpret_cast<uint2 *>(smem + lay.ld);               // [nstab] step descriptors
    uint16_t *s_ll = reinterpret_cast<uint16_t *>(smem + lay.lut);        // TABLE: [patterns <= 16][256] packed weight changes
    int8_t *s_dE2 = reinterpret_cast<int8_t *>(smem + lay.lut);           // TABLE2: [512]
    uint4 *s_draw = reinterpret_cast<uint4 *>(smem + lay.draw);           // [NLC][H + 2 iters] the top rungs' Philox calls of the step
    __shared__ uint32_t s_patmask[8];
    __shared__ int s_alive, s_anyfin;

    // ---------------- tables (once per CTA; the CTA is persistent) ----------------
    if (!WEIGHTED)
        for (int i = tid; i < Nc * 9; i += blockDim.x) s_thru[i] = p.thr_u[i];
    for (int i = tid; i < 8 * NLC; i += blockDim.x) { s_t[-1 - i] = QECMC_PT_BAD - 1; s_n[-1 - i] = 0; }
    if (TABLE2) {
        for (int i = tid; i < g.nstab; i += blockDim.x) s_ld[i] = p.desc2[i];
        for (int e = tid; e < 512; e += blockDim.x) {
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
        // one-layer codes: a stabilizer touches at most two adjacent qubits in each of two row words.  Per stabilizer: the two
        // word indices, the bit position of each word's lower touched field and the Paulis of the four slots; the weight
        // changes (dx, dy, dz, their sum) come from a LUT indexed by (pattern rank, the four touched fields)
        if (tid < 8) s_patmask[tid] = 0;
        __syncthreads();
        for (int i = tid; i < g.nstab; i += blockDim.x) {
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
            // bit 31 of the second word: the stabilizer touches a qubit of the diagonal or the antidiagonal (the supports of
            // the XZZX logical operators, whose Pauli counts the top rung caches)
            uint32_t diag = 0;
            for (int k = 0; k < 2; k++)
                for (int fq = 0; fq < 2; fq++)
                    if ((nib[k] >> (2 * fq)) & 3u) {
                        const int cq = (int)(pos[k] >> 1) + fq;
                        if (cq == u.w[k] || cq == g.L - 1 - u.w[k]) diag = 0x80000000u;
                    }
            s_ld[i] = make_uint2((uint32_t)u.w[0] | ((uint32_t)u.w[1] << 8) | (pos[0] << 16) | (pos[1] << 24), pat | diag);
            atomicOr(&s_patmask[pat >> 5], 1u << (pat & 31));
        }
        __syncthreads();
        for (int i = tid; i < g.nstab; i += blockDim.x) {
            const uint32_t pat = s_ld[i].y & 0xFFu, diag = s_ld[i].y & 0x80000000u;
            uint32_t id = __popc(s_patmask[pat >> 5] & ((1u << (pat & 31)) - 1u));
            for (uint32_t wq = 0; wq < (pat >> 5); wq++) id += __popc(s_patmask[wq]);
            s_ld[i].y = pat | (id << 8) | diag;
        }
        for (int e = tid; e < 16 * 256; e += blockDim.x) {
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

    // ---------------- roles ----------------
    // The top-rung warps come FIRST in the CTA: their serial blocks are the critical path of a step, and the scheduler
    // favours the older (lower-numbered) warps of a CTA when several are ready.
    const int rwarp = warp - lay.NTW;                                              // index among the rung warps
    const bool rung_warp = warp >= lay.NTW;
    const int LT = p.lt, TPW = 32 / LT;
    const int sub = rung_warp ? 0 : lane % LT;                                     // lane within a top replica's group
    const int l = rung_warp ? lane % NLC : warp * TPW + lane / LT;                 // this thread's ladder within the CTA
    const int my_r = rung_warp ? rwarp * RPW + lane / NLC : Nc - 1;                 // this thread's rung
    const bool worker = rung_warp ? my_r < NR : l < NLC;
    const bool is_mgr = rwarp == 0 && lane < NLC;                                  // lane = ladder it manages
    const uint32_t gm = LT >= 32 ? 0xFFFFFFFFu : (((1u << LT) - 1u) << (lane / LT * LT));   // the lanes sharing this top replica
    const uint32_t H = (uint32_t)((p.iters + 1) >> 1);
    // (double)x * 2^-32 < p_logical for a 32-bit draw x  <=>  x < ceil(p_logical * 2^32) (the scaling is exact)
    const uint64_t plog = p.p_logical >= 1.0 ? (1ull << 32) : (uint64_t)ceil(p.p_logical * 4294967296.0);
    const uint32_t nstab = (uint32_t)g.nstab;
    // weight tables of this thread's rung: 32-bit offsets into p.wtab (the base stays in uniform registers; a per-thread
    // 64-bit pointer is rematerialised with a dozen instructions at every use under the register cap)
    const uint32_t wo = WEIGHTED ? (uint32_t)(worker ? my_r : 0) * 4u * (uint32_t)ns1 : 0u;
    auto chain_weight_r = [&](int cx, int cy, int cz) {   // chain_weight (qecmc_ladder.cuh) on this rung's tables
        const double *wb = p.wtab;
        const uint32_t u1 = (uint32_t)ns1;
        double a = __dmul_rn(wb[wo + (uint32_t)cx], wb[wo + u1 + (uint32_t)cy]);
        a = __dmul_rn(a, wb[wo + 2u * u1 + (uint32_t)cz]);
        return __dmul_rn(a, wb[wo + 3u * u1 + (uint32_t)(cx + cy + cz)]);
    };
    int e_nz = 0, e_nxy = 0;     // alpha ladders: the RUNG's n_eff = e_nz + alpha * e_nxy (mcmc_alpha.py:22,56 -- not swapped, :126-131)
    uint32_t nacc = 0, nacc_s = 0;   // accepted moves: confirmed, and of the step in flight

    // (re)start ladder `lidx` of the launch in this thread's ladder slot: replica r starts in column r
    auto init_ladder = [&](uint32_t lidx, int buf) {
        const W *src = reinterpret_cast<const W *>(p.lat_in) + ((size_t)(p.init_broadcast ? lidx : (size_t)lidx * Nc + my_r)) * g.nw;
        W *colp = tile + my_r * NLC + l;
        int nx = 0, ny = 0, nz = 0;
        for (int w = sub; w < g.nw; w += (rung_warp ? 1 : LT)) {
            const W v = src[w];
            colp[(size_t)w * NREP] = v;
            nx += popc(xmap(v)); ny += popc(ymap(v)); nz += popc(zmap(v));
        }
        if (!rung_warp) {
            __syncwarp(gm);
            for (int o = LT >> 1; o > 0; o >>= 1) {
                nx += __shfl_xor_sync(gm, nx, o);
                ny += __shfl_xor_sync(gm, ny, o);
                nz += __shfl_xor_sync(gm, nz, o);
            }
        }
        if (sub == 0) {
            struct Col { const W *b; int s; __device__ W get(int w) const { return b[(size_t)w * s]; } } cc{colp, NREP};
            const int cls = class_raw_to_label(GEOM, lat_class<GEOM, W>(g, cc));   // label -> raw (the map is an involution)
            const int flag = p.flags_in ? p.flags_in[(size_t)lidx * Nc + my_r] : (my_r == Nc - 1 ? 1 : 0);
            const int col = my_r * NLC + l;
            s_state[col] = WEIGHTED ? ((uint32_t)nx | ((uint32_t)ny << 10) | ((uint32_t)nz << 20)) : (uint32_t)(nx + ny + nz);
            s_state[NREP + col] = (uint32_t)cls | (flag ? QECMC_PT_FLAG : 0u);
            s_slot[(size_t)buf * NREP + my_r * NLC + l] = (uint8_t)my_r;
        }
        e_nz = nz; e_nxy = nx + ny;
        if (p.neff_in) { const int2 e = p.neff_in[(size_t)lidx * Nc + my_r]; e_nz = e.x; e_nxy = e.y; }
    };
    auto init_manager = [&](uint32_t lidx) {
        for (int f = 0; f < PA_NF; f++) s_acc[f * NLC + lane] = 0;
        s_acc[PA_ACTIVE * NLC + lane] = 1;
        s_acc[PA_LIDX * NLC + lane] = lidx;
        s_acc[PA_TOPS0 * NLC + lane] = p.tops0_in ? (uint32_t)p.tops0_in[lidx] : 0u;
        s_mask[lane] = 0;          // alpha ladders OR their pair decisions into the step's mask: a new ladder starts from none
        s_mask[NLC + lane] = 0;
    };

    // PTEQ's accounting of one completed Ladder.step (decoders.py:56-82; decoders_biasednoise.py:196-215 records the bottom
    // rung's n_eff), from what the bottom-rung thread took down in phase C; decides whether the ladder ends.  With `defer`
    // the thread runs it at the top of the NEXT step, beside the other warps' Metropolis block: the ladder then takes one
    // step more than it needed, which nothing reads (the call returns counts, not lattices).
    const bool defer = p.acct == ACCT_PTEQ && Nc > 1 && !p.lat_out && !p.flags_out && !p.neff_out;
    auto account = [&](int ml) {
        const uint32_t lidx = s_acc[PA_LIDX * NLC + ml];
        const uint32_t step = s_acc[PA_STEP * NLC + ml] - 1u;            // steps completed before the one accounted here
        const uint32_t tops0 = s_acc[PA_TOPS0 * NLC + ml];
        s_acc[PA_PEND * NLC + ml] = 0;
        bool fin = false;
        int converged = 0;
        if (p.acct == ACCT_PTEQ) {
            const int cur = class_raw_to_label(GEOM, (int)s_acc[PA_BOTCLS * NLC + ml]);   // the tracked bits are the XOR-linear raw class
            const uint32_t h_a = s_acc[PA_HA * NLC + ml], h_b = p.kind == LK_ALPHA ? s_acc[PA_HB * NLC + ml] : 0u;
            uint32_t since = s_acc[PA_SINCE * NLC + ml];
            const uint32_t burn = s_acc[PA_BURN * NLC + ml];
            if (tops0 >= (uint32_t)p.tops_burn) {
                since = step - burn;
                s_acc[(PA_EQC + cur) * NLC + ml]++;
                if (p.use_conv) {
                    auto ld64 = [&](int f) { return (long long)((uint64_t)s_acc[f * NLC + ml] | ((uint64_t)s_acc[(f + 1) * NLC + ml] << 32)); };
                    auto st64 = [&](int f, long long v) { s_acc[f * NLC + ml] = (uint32_t)v; s_acc[(f + 1) * NLC + ml] = (uint32_t)((uint64_t)v >> 32); };
                    long long S2a = ld64(PA_S2A), S2b = ld64(PA_S2B), S4a = ld64(PA_S4A), S4b = ld64(PA_S4B);
                    const size_t hrow = ((size_t)blockIdx.x * NLC + ml) * (size_t)p.hist_stride;
                    uint16_t *h16 = reinterpret_cast<uint16_t *>(p.hist) + hrow;
                    uint32_t *h32 = reinterpret_cast<uint32_t *>(p.hist) + hrow;
                    const bool wide_hist = p.kind == LK_ALPHA;
                    if (wide_hist) h32[since] = h_a | (h_b << 16); else h16[since] = (uint16_t)h_a;
                    // history windows [l/4, l/2) and [3l/4, l) of conv_crit_error_based_PT (decoders.py:93-105)
                    const uint32_t wl = since + 1, nC = (uint32_t)(3ull * wl / 4), nB = wl / 2, nA = wl / 4;
                    const uint32_t wA = s_acc[PA_WA * NLC + ml], wB = s_acc[PA_WB * NLC + ml], wC = s_acc[PA_WC * NLC + ml];
                    S4a += h_a; S4b += h_b;
                    // the three window edges move by at most one entry per step; their loads are issued together
                    uint32_t vC = 0, vB = 0, vA = 0;
                    if (wide_hist) {
                        if (nC > wC) vC = __ldcg(h32 + wC);
                        if (nB > wB) vB = __ldcg(h32 + wB);
                        if (nA > wA) vA = __ldcg(h32 + wA);
                    } else {
                        if (nC > wC) vC = __ldcg(h16 + wC);
                        if (nB > wB) vB = __ldcg(h16 + wB);
                        if (nA > wA) vA = __ldcg(h16 + wA);
                    }
                    if (nC > wC) { S4a -= vC & 0xFFFFu; S4b -= vC >> 16; }
                    if (nB > wB) { S2a += vB & 0xFFFFu; S2b += vB >> 16; }
                    if (nA > wA) { S2a -= vA & 0xFFFFu; S2b -= vA >> 16; }
                    s_acc[PA_WA * NLC + ml] = nA; s_acc[PA_WB * NLC + ml] = nB; s_acc[PA_WC * NLC + ml] = nC;
                    st64(PA_S2A, S2a); st64(PA_S2B, S2b); st64(PA_S4A, S4a); st64(PA_S4B, S4b);
                    if (tops0 >= (uint32_t)p.TOPS) {
                        const long long lq = (long long)since + 1;
                        double q2, q4;
                        if (p.kind == LK_ALPHA) {
                            q2 = ((double)S2a + p.alpha * (double)S2b) / (double)(lq / 2 - lq / 4);
                            q4 = ((double)S4a + p.alpha * (double)S4b) / (double)(lq - 3 * lq / 4);
                        } else {
                            q2 = (double)S2a / (double)(lq / 2 - lq / 4);
                            q4 = (double)S4a / (double)(lq - 3 * lq / 4);
                        }
                        const uint32_t streak = s_acc[PA_CSTREAK * NLC + ml];
                        if (fabs(q2 - q4) < p.eps) {
                            if ((long long)streak >= p.SEQ) { fin = true; converged = 1; }
                            s_acc[PA_CSTREAK * NLC + ml] = tops0 - s_acc[PA_CSTART * NLC + ml];
                        } else {
                            s_acc[PA_CSTREAK * NLC + ml] = 0;
                            s_acc[PA_CSTART * NLC + ml] = tops0;
                        }
                    }
                }
                s_acc[PA_SINCE * NLC + ml] = since;
            } else {
                s_acc[PA_BURN * NLC + ml] = burn + 1;
                if (p.use_conv && tops0 >= (uint32_t)p.TOPS) {   // TOPS <= tops_burn is refused by the reference's callers; kept for symmetry
                    s_acc[PA_CSTREAK * NLC + ml] = 0;
                    s_acc[PA_CSTART * NLC + ml] = tops0;
                }
            }
        }
        if ((long long)step + 1 >= p.steps) fin = true;
        if (fin) {
            if (p.acct == ACCT_PTEQ) {
                const uint32_t since = s_acc[PA_SINCE * NLC + ml];
                if (p.info) {
                    p.info[4 * (size_t)lidx] = (long long)step + 1;
                    p.info[4 * (size_t)lidx + 1] = since;
                    p.info[4 * (size_t)lidx + 2] = tops0;
                    p.info[4 * (size_t)lidx + 3] = converged;
                }
                for (int e = 0; e < g.neq; e++) {   // decoders.py:89: (eq[since_burn] / (since_burn + 1) * 100).astype(np.uint8)
                    const uint32_t cnt = s_acc[(PA_EQC + e) * NLC + ml];
                    if (p.eq_counts) p.eq_counts[(size_t)lidx * g.neq + e] = cnt;
                    if (p.percent) p.percent[(size_t)lidx * g.neq + e] = (uint8_t)(int)((double)cnt / (double)(since + 1) * 100);
                }
            }
            if (p.tops0_out) p.tops0_out[lidx] = tops0;
            const unsigned int nxt = atomicAdd(p.queue, 1u);
            s_acc[PA_NEXT * NLC + ml] = (long long)nxt < p.n_ladders ? nxt : 0xFFFFFFFFu;
            s_acc[PA_FIN * NLC + ml] = 1;
        }
    };

    __syncthreads();
    if (is_mgr) {
        const int64_t lidx = (int64_t)blockIdx.x * NLC + lane;
        if (lidx < p.n_ladders) init_manager((uint32_t)lidx);
        else { for (int f = 0; f < PA_NF; f++) s_acc[f * NLC + lane] = 0; }
        s_mask[lane] = 0;
        s_mask[NLC + lane] = 0;
    }
    if (tid == 0) { s_alive = 1; s_anyfin = 0; }
    __syncthreads();
    if (worker && s_acc[PA_ACTIVE * NLC + l]) init_ladder(s_acc[PA_LIDX * NLC + l], 0);
    __syncthreads();

    int buf = 0;
    while (true) {
        // =====================================================================================================
        // phase A: `iters` Metropolis steps on every rung (Ladder.update_ladder, mcmc.py:81-83)
        // =====================================================================================================
        __syncwarp();   // a top group's leader wrote the group's slot and flag in phase C
        const bool active = worker && s_acc[PA_ACTIVE * NLC + l] != 0;
        // the previous step's accounting, by the thread that took its inputs down (the bottom rung's), while the other warps
        // are already in this step's block
        if (defer && active && my_r == 0 && sub == 0 && s_acc[PA_PEND * NLC + l]) account(l);
        if (active) {
            const uint32_t lidx = s_acc[PA_LIDX * NLC + l];
            const uint64_t gid = (uint64_t)p.ladder_offset + lidx;
            const uint32_t id_lo = (uint32_t)gid, id_hi = (uint32_t)(gid >> 32);
            const uint32_t step = p.step0 + s_acc[PA_STEP * NLC + l];
            const int col = (int)s_slot[(size_t)buf * NREP + my_r * NLC + l] * NLC + l;
            W *mycol = tile + col;
            const uint32_t st0 = s_state[col];
            int n = (int)st0, nx = 0, ny = 0, nz = 0;
            if (WEIGHTED) { nx = st0 & 1023; ny = (st0 >> 10) & 1023; nz = st0 >> 20; n = nx + ny + nz; }
            double pb = 0.0;
            if (WEIGHTED) pb = chain_weight_r(nx, ny, nz);   // frozen for the block
            uint32_t st1 = rung_warp ? 0u : s_state[NREP + col];     // top rung: the replica's class changes with logical moves
            uint4 R = make_uint4(0, 0, 0, 0);

            // one stabilizer proposal on this thread's replica: new words, weight change(s)
            W nv[NU];
            int uw[NU];
            int dE, dx, dy, dz;
            uint32_t dflag = 0;   // one-layer codes: the proposed stabilizer touches a diagonal qubit
            // one-layer codes: the proposal from its descriptor
            auto propose_tab = [&](const uint2 D) {
                uw[0] = (int)(D.x & 0xFFu);
                uw[1] = (int)((D.x >> 8) & 0xFFu);
                const uint32_t p0 = (D.x >> 16) & 0xFFu, p1 = D.x >> 24;
                const W o0 = mycol[(size_t)uw[0] * NREP], o1 = mycol[(size_t)uw[1] * NREP];
                const uint32_t f = ((uint32_t)(o0 >> p0) & 0xFu) | (((uint32_t)(o1 >> p1) & 0xFu) << 4);
                const uint32_t pk = s_ll[(((D.y >> 8) & 0xFFu) << 8) + f];
                dflag = D.y >> 31;
                nv[0] = (W)(o0 ^ ((W)(D.y & 0xFu) << p0));
                nv[1] = (W)(o1 ^ ((W)((D.y >> 4) & 0xFu) << p1));
                dx = dy = dz = 0;
                if (WEIGHTED) { dx = (int)(pk & 15u) - 4; dy = (int)((pk >> 4) & 15u) - 4; dz = (int)((pk >> 8) & 15u) - 4; }
                dE = (int)(pk >> 12) - 4;
            };
            auto propose = [&](int idx) {
                dE = dx = dy = dz = 0;
                if (TABLE) {
                    propose_tab(s_ld[idx]);
                } else if (TABLE2) {
                    const uint2 D = s_ld[idx];
                    uw[0] = (int)((D.x >> 8) & 0xFFu);
                    uw[1] = (int)((D.x >> 16) & 0xFFu);
                    uw[NU > 2 ? 2 : 0] = (int)(D.x >> 24);
                    const W o0 = mycol[(size_t)uw[0] * NREP], o1 = mycol[(size_t)uw[1] * NREP], o2 = mycol[(size_t)(D.x >> 24) * NREP];
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
                    int row, colq, op;
                    idx_to_rco<GEOM>(g, idx, row, colq, op);
                    Upd<W> u;
                    decode<GEOM, W>(g, row, colq, op, u);
#pragma unroll
                    for (int i = 0; i < NU; i++) {
                        uw[i] = u.w[i];
                        const W o = mycol[(size_t)u.w[i] * NREP];
                        nv[i] = (W)(o ^ u.m[i]);
                        dx += popc(xmap(nv[i])) - popc(xmap(o));
                        dy += popc(ymap(nv[i])) - popc(ymap(o));
                        dz += popc(zmap(nv[i])) - popc(zmap(o));
                    }
                    dE = dx + dy + dz;
                }
            };
            auto commit = [&]() {
#pragma unroll
                for (int i = 0; i < NU; i++) mycol[(size_t)uw[i] * NREP] = nv[i];
            };
            // the same proposal as XOR masks only (no reads): uw[] and mk[]; the words are distinct or their masks zero
            auto stab_masks = [&](int idx, W *mk) {
                if (TABLE) {
                    const uint2 D = s_ld[idx];
                    uw[0] = (int)(D.x & 0xFFu);
                    uw[1] = (int)((D.x >> 8) & 0xFFu);
                    mk[0] = (W)((W)(D.y & 0xFu) << ((D.x >> 16) & 0xFFu));
                    mk[1] = (W)((W)((D.y >> 4) & 0xFu) << (D.x >> 24));
                } else if (TABLE2) {
                    const uint2 D = s_ld[idx];
                    uw[0] = (int)((D.x >> 8) & 0xFFu);
                    uw[1] = (int)((D.x >> 16) & 0xFFu);
                    uw[NU > 2 ? 2 : 0] = (int)(D.x >> 24);
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
                    mk[0] = m0;
                    mk[1] = m1;
                    mk[NU > 2 ? 2 : 0] = m2;
                } else {
                    int row, colq, op;
                    idx_to_rco<GEOM>(g, idx, row, colq, op);
                    Upd<W> u;
                    decode<GEOM, W>(g, row, colq, op, u);
#pragma unroll
                    for (int i = 0; i < NU; i++) { uw[i] = u.w[i]; mk[i] = u.m[i]; }
                }
            };

            if (rung_warp) {
                // this rung's nine thresholds through a shared-memory address the compiler cannot re-derive (it rebuilds the
                // carve-up offset with ~8 instructions per Metropolis step otherwise)
                uint32_t thr_addr = (uint32_t)__cvta_generic_to_shared(s_thru + my_r * 9 + QECMC_THR_OFF);
                asm volatile("mov.u32 %0, %1;" : "=r"(thr_addr) : "r"(thr_addr));
                for (int it = 0; it < p.iters; it++) {
                    if ((it & 1) == 0) R = philox4x32_10(step * H + (uint32_t)(it >> 1), (uint32_t)my_r, id_lo, id_hi, p.keys);
                    const uint32_t w_idx = (it & 1) ? R.z : R.x, w_acc = (it & 1) ? R.w : R.y;
                    propose((int)__umulhi(w_idx, nstab));
                    bool acc;
                    if (WEIGHTED) acc = (double)w_acc * U32 * pb < chain_weight_r(nx + dx, ny + dy, nz + dz);
                    else acc = w_acc <= lds_u32<0>(thr_addr + (uint32_t)(dE * 4));
                    if (acc) {
                        commit();
                        if (WEIGHTED) { nx += dx; ny += dy; nz += dz; e_nz = nz; e_nxy = nx + ny; }
                        n += dE;
                        nacc_s++;
                    }
                }
            } else {
                // ---- top rung (mcmc.py:24-33 and the weighted chains' p_logical branch): LT lanes share the replica ----
                // All of the step's draws first, in parallel over the group's lanes, into shared memory: calls [0, H) are the
                // rung's proposal / accept words, [H, H + iters) the logical-or-stabilizer decisions, [H + iters, H + 2 iters)
                // the logical operators' positions.  The serial part below then holds no generator arithmetic.
                uint4 *draws = s_draw + (size_t)l * (H + 2u * (uint32_t)p.iters);
                const uint32_t ncalls = H + (GEOM == XZZX ? 1u : 2u) * (uint32_t)p.iters;   // the XZZX logicals take no position
                for (uint32_t c = (uint32_t)sub; c < ncalls; c += (uint32_t)LT) {
                    uint32_t c0, tag;
                    if (c < H) { c0 = step * H + c; tag = (uint32_t)my_r; }
                    else if (c < H + (uint32_t)p.iters) { c0 = step * (uint32_t)p.iters + (c - H); tag = (1u << 8) | (uint32_t)my_r; }
                    else { c0 = step * (uint32_t)p.iters + (c - H - (uint32_t)p.iters); tag = (2u << 8) | (uint32_t)my_r; }
                    draws[c] = philox4x32_10(c0, tag, id_lo, id_hi, p.keys);
                }
                __syncwarp(gm);
                const bool walk = !WEIGHTED && p.top_accept_all;
                if (walk) {
                    // A depolarizing ladder's top rung sits at p = 0.75 and accepts every proposal (mcmc.py:30).  All of the block's
                    // moves are then XOR masks, which commute: lane `sub` applies to the row words it owns (w = sub mod LT) its
                    // share of every move, with no reads of other lanes' words and no synchronisation inside the block; the weight
                    // is recounted once at the end.
                    // Logical operators are XOR-linear too: the block's logical moves are gathered into two column masks and two
                    // sets of rows (per layer) and reach the rows once, together with the recount.
                    W colA = 0, colB = 0;
                    uint32_t rowsA = 0, rowsB = 0;
                    for (int it = 0; it < p.iters; it++) {
                        const uint4 Tw = draws[H + it];
                        if ((uint64_t)Tw.x < plog) {
                            const int op0 = (int)(Tw.z >> 30), op1 = NLAY == 2 ? (int)((Tw.z >> 28) & 3u) : 0;
                            int x0 = 0, z0 = 0, x1 = 0, z1 = 0;
                            if (GEOM != XZZX) {
                                const uint4 P = draws[H + p.iters + it];
                                x0 = (op0 == 1 || op0 == 2) ? (int)__umulhi(P.x, (uint32_t)L) : 0;
                                z0 = (op0 == 3 || op0 == 2) ? (int)__umulhi(P.y, (uint32_t)L) : 0;
                                x1 = (op1 == 1 || op1 == 2) ? (int)__umulhi(P.z, (uint32_t)L) : 0;
                                z1 = (op1 == 3 || op1 == 2) ? (int)__umulhi(P.w, (uint32_t)L) : 0;
                            }
                            // the terms of logical_mask (qecmc_lattice.h), operator by operator
                            if (GEOM == TORIC) {
                                if (op0 == 1 || op0 == 2) rowsA ^= 1u << x0;
                                if (op0 == 3 || op0 == 2) colA ^= fld<W>(3, z0);
                                if (op1 == 1 || op1 == 2) colB ^= fld<W>(1, x1);
                                if (op1 == 3 || op1 == 2) rowsB ^= 1u << z1;
                            } else if (GEOM == PLANAR) {
                                if (op0 == 1 || op0 == 3) rowsA ^= 1u << x0;
                                if (op0 == 2 || op0 == 3) colA ^= fld<W>(3, z0);
                            } else if (GEOM == ROTATED) {
                                if (op0 == 1 || op0 == 3) colA ^= fld<W>(1, x0);
                                if (op0 == 2 || op0 == 3) rowsA ^= 1u << z0;
                            } else {
                                if (op0 == 1 || op0 == 2) rowsA ^= 1u;
                                if (op0 == 3 || op0 == 2) rowsA ^= 2u;
                            }
                            st1 ^= (uint32_t)(p.cls_delta[op0] ^ (NLAY == 2 ? p.cls_delta[4 + op1] : 0));
                        } else {
                            const uint4 Rw = draws[it >> 1];
                            W mk[NU];
                            stab_masks((int)__umulhi((it & 1) ? Rw.z : Rw.x, nstab), mk);
#pragma unroll
                            for (int i = 0; i < NU; i++)
                                if ((uw[i] & (LT - 1)) == sub && mk[i]) mycol[(size_t)uw[i] * NREP] ^= mk[i];
                        }
                    }
                    int cnt = 0;
                    for (int w = sub; w < g.nw; w += LT) {
                        W m;
                        if (GEOM == TORIC) m = w < L ? (W)((((rowsA >> w) & 1u) ? rowmask<W>(1, L) : (W)0) ^ colA)
                                                     : (W)(colB ^ (((rowsB >> (w - L)) & 1u) ? rowmask<W>(3, L) : (W)0));
                        else if (GEOM == PLANAR) m = w < L ? (W)((((rowsA >> w) & 1u) ? rowmask<W>(1, L) : (W)0) ^ colA) : (W)0;
                        else if (GEOM == ROTATED) m = (W)(colA ^ (((rowsA >> w) & 1u) ? rowmask<W>(3, L) : (W)0));
                        else m = (W)(((rowsA & 1u) ? fld<W>(1, L - 1 - w) : (W)0) ^ ((rowsA & 2u) ? fld<W>(3, w) : (W)0));
                        const W v = (W)(mycol[(size_t)w * NREP] ^ m);
                        if (m) mycol[(size_t)w * NREP] = v;
                        cnt += weight<W>(v);
                    }
                    for (int o = LT >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(gm, cnt, o);
                    n = cnt;
                    if (sub == 0) nacc_s += (uint32_t)p.iters;
                } else if (WEIGHTED && GEOM == XZZX) {
                    // The XZZX logical operators have fixed supports (xzzx_model.py:340-357): X on the antidiagonal, Z on the diagonal.
                    // Every lane of the group keeps the two diagonals of the replica as packed words (A: field L-1-w of row w at bit
                    // 2w; Dg: field w of row w; the centre qubit, which lies on both, apart in c0) together with their Pauli counts.
                    // X on a diagonal swaps its counts I <-> X and Y <-> Z, Z swaps I <-> Z and X <-> Y, so a logical proposal is
                    // evaluated from a handful of registers -- no row reads, no reduction -- and only an accepted one touches the
                    // rows (each lane its own).  An accepted stabilizer refreshes the fields of its two rows and recounts.
                    const bool odd = (L & 1) != 0;
                    const int cw = (L - 1) >> 1, nd = L - (odd ? 1 : 0);
                    const W CM = odd ? (W)((W)3 << (2 * cw)) : (W)0;
                    W A = 0, Dg = 0;
                    for (int w = sub; w < g.nw; w += LT) {
                        const W v = mycol[(size_t)w * NREP];
                        A |= (W)(((v >> (2 * (L - 1 - w))) & (W)3) << (2 * w));
                        Dg |= (W)(((v >> (2 * w)) & (W)3) << (2 * w));
                    }
                    for (int o = LT >> 1; o > 0; o >>= 1) {
                        A |= (W)__shfl_xor_sync(gm, A, o);
                        Dg |= (W)__shfl_xor_sync(gm, Dg, o);
                    }
                    int c0 = odd ? (int)((A >> (2 * cw)) & (W)3) : 0;
                    A &= (W)~CM;
                    Dg &= (W)~CM;
                    int aX = popc(xmap(A)), aY = popc(ymap(A)), aZ = popc(zmap(A));
                    int dX = popc(xmap(Dg)), dY = popc(ymap(Dg)), dZ = popc(zmap(Dg));
                    const W M1 = (W)(rowmask<W>(1, L) & ~CM), M3 = (W)(rowmask<W>(3, L) & ~CM);
                    for (int it = 0; it < p.iters; it++) {
                        const uint4 Tw = draws[H + it];
                        const bool logical = (uint64_t)Tw.x < plog;
                        uint32_t w_acc;
                        int ex = 0, ey = 0, ez = 0, op0 = 0, cn = c0;
                        bool px = false, pz = false;
                        if (logical) {
                            op0 = (int)(Tw.z >> 30);
                            px = op0 == 1 || op0 == 2;
                            pz = op0 == 3 || op0 == 2;
                            if (px) { ex += nd - aX - aY - aZ - aX; ey += aZ - aY; ez += aY - aZ; }
                            if (pz) { ez += nd - dX - dY - dZ - dZ; ex += dY - dX; ey += dX - dY; }
                            if (odd) {
                                cn = c0 ^ (px ? 1 : 0) ^ (pz ? 3 : 0);
                                ex += (cn == 1) - (c0 == 1);
                                ey += (cn == 2) - (c0 == 2);
                                ez += (cn == 3) - (c0 == 3);
                            }
                            w_acc = Tw.y;
                        } else {
                            // (fetching the next iteration's draws and descriptor ahead was measured: 5.8e10 against 6.0e10
                            // steps/s -- the extra live registers spill)
                            const uint4 Rw = draws[it >> 1];
                            propose((int)__umulhi((it & 1) ? Rw.z : Rw.x, nstab));
                            ex = dx; ey = dy; ez = dz;
                            w_acc = (it & 1) ? Rw.w : Rw.y;
                        }
                        const bool acc = (double)w_acc * U32 * pb < chain_weight_r(nx + ex, ny + ey, nz + ez);
                        if (acc) {
                            if (logical) {
                                for (int w = sub; w < g.nw; w += LT) {
                                    const W m = logical_mask<GEOM, W>(g, w, op0, 0, 0, 0, 0, 0);
                                    if (m) mycol[(size_t)w * NREP] ^= m;
                                }
                                if (px) { const int t = nd - aX - aY - aZ, u = aY; aX = t; aY = aZ; aZ = u; A ^= M1; }
                                if (pz) { const int t = nd - dX - dY - dZ, u = dX; dZ = t; dX = dY; dY = u; Dg ^= M3; }
                                c0 = cn;
                                st1 ^= (uint32_t)p.cls_delta[op0];
                            } else {
                                if (sub == 0) commit();
                                if (dflag) {
#pragma unroll
                                for (int i = 0; i < NU; i++) {   // an untouched spare word refreshes its fields with what they were
                                    const int w = uw[i];
                                    const W v = nv[i];
                                    A = (W)((A & ~((W)3 << (2 * w))) | (((v >> (2 * (L - 1 - w))) & (W)3) << (2 * w)));
                                    Dg = (W)((Dg & ~((W)3 << (2 * w))) | (((v >> (2 * w)) & (W)3) << (2 * w)));
                                    if (odd && w == cw) c0 = (int)((v >> (2 * cw)) & (W)3);
                                }
                                A &= (W)~CM;
                                Dg &= (W)~CM;
                                aX = popc(xmap(A)); aY = popc(ymap(A)); aZ = popc(zmap(A));
                                dX = popc(xmap(Dg)); dY = popc(ymap(Dg)); dZ = popc(zmap(Dg));
                                }
                            }
                            nx += ex; ny += ey; nz += ez; e_nz = nz; e_nxy = nx + ny;
                            n += ex + ey + ez;
                            if (sub == 0) nacc_s++;
                        }
                        __syncwarp(gm);   // the group's stores are visible before its next loads
                    }
                } else
                for (int it = 0; it < p.iters; it++) {
                    const uint4 Tw = draws[H + it];
                    const bool logical = (uint64_t)Tw.x < plog;
                    if (logical) {
                        const uint4 P = draws[H + p.iters + it];
                        // operators and positions as _apply_random_logical draws them (a position only for operator 1 / 2
                        // resp. 3 / 2 in every code: toric_model.py:228-253, planar_model.py:271-288)
                        const int op0 = (int)(Tw.z >> 30), op1 = NLAY == 2 ? (int)((Tw.z >> 28) & 3u) : 0;
                        const int x0 = (op0 == 1 || op0 == 2) ? (int)__umulhi(P.x, (uint32_t)L) : 0;
                        const int z0 = (op0 == 3 || op0 == 2) ? (int)__umulhi(P.y, (uint32_t)L) : 0;
                        const int x1 = (op1 == 1 || op1 == 2) ? (int)__umulhi(P.z, (uint32_t)L) : 0;
                        const int z1 = (op1 == 3 || op1 == 2) ? (int)__umulhi(P.w, (uint32_t)L) : 0;
                        const uint32_t dcls = (uint32_t)(p.cls_delta[op0] ^ (NLAY == 2 ? p.cls_delta[4 + op1] : 0));
                        {
                            int ddx = 0, ddy = 0, ddz = 0;
                            for (int w = sub; w < g.nw; w += LT) {
                                const W m = logical_mask<GEOM, W>(g, w, op0, op1, x0, z0, x1, z1);
                                if (m) {
                                    const W o = mycol[(size_t)w * NREP], nw_ = (W)(o ^ m);
                                    if (WEIGHTED) {
                                        ddx += popc(xmap(nw_)) - popc(xmap(o));
                                        ddy += popc(ymap(nw_)) - popc(ymap(o));
                                        ddz += popc(zmap(nw_)) - popc(zmap(o));
                                    } else {
                                        ddx += weight<W>(nw_) - weight<W>(o);   // depolarizing: only the total weight matters
                                    }
                                }
                            }
                            for (int o = LT >> 1; o > 0; o >>= 1) {   // the whole group took this branch
                                ddx += __shfl_xor_sync(gm, ddx, o);
                                if (WEIGHTED) {
                                    ddy += __shfl_xor_sync(gm, ddy, o);
                                    ddz += __shfl_xor_sync(gm, ddz, o);
                                }
                            }
                            const int d = ddx + ddy + ddz;
                            bool acc;
                            if (WEIGHTED) acc = (double)Tw.y * U32 * pb < chain_weight_r(nx + ddx, ny + ddy, nz + ddz);
                            else acc = (p.top_accept_all || d <= 0) ? true : ((double)Tw.y * U32 < p.thr_top_d[d + 4 * L]);
                            if (acc) {
                                for (int w = sub; w < g.nw; w += LT) {
                                    const W m = logical_mask<GEOM, W>(g, w, op0, op1, x0, z0, x1, z1);
                                    if (m) mycol[(size_t)w * NREP] ^= m;
                                }
                                if (WEIGHTED) { nx += ddx; ny += ddy; nz += ddz; e_nz = nz; e_nxy = nx + ny; }
                                n += d;
                                st1 ^= dcls;
                                if (sub == 0) nacc_s++;
                            }
                        }
                    } else {
                        const uint4 Rw = draws[it >> 1];
                        const uint32_t w_idx = (it & 1) ? Rw.z : Rw.x, w_acc = (it & 1) ? Rw.w : Rw.y;
                        {
                            propose((int)__umulhi(w_idx, nstab));
                            bool acc;
                            if (WEIGHTED) acc = (double)w_acc * U32 * pb < chain_weight_r(nx + dx, ny + dy, nz + dz);
                            else acc = (p.top_accept_all || dE <= 0) ? true : ((double)w_acc * U32 < p.thr_top_d[dE + 4 * L]);
                            if (acc) {
                                if (sub == 0) { commit(); nacc_s++; }
                                if (WEIGHTED) { nx += dx; ny += dy; nz += dz; e_nz = nz; e_nxy = nx + ny; }
                                n += dE;
                            }
                        }
                    }
                    __syncwarp(gm);   // the group's stores are visible before its next loads
                }
                if (sub == 0) s_state[NREP + col] = st1;
            }
            if (sub == 0) {
                s_state[col] = WEIGHTED ? ((uint32_t)nx | ((uint32_t)ny << 10) | ((uint32_t)nz << 20)) : (uint32_t)n;
                // ---- what the swap sweep needs from this rung ----
                if (p.kind == LK_ALPHA) {
                    s_n[my_r * NLC + l] = e_nz;
                    s_t[my_r * NLC + l] = e_nxy;
                } else {
                    s_n[my_r * NLC + l] = n;
                    if (my_r < Nc - 1) {
                        // pair (r, r+1) swaps iff u < diff^k, k = n_hi - n_lo (mcmc.py:144-149; the biased ladder draws always,
                        // mcmc_biased.py:154-156, which for k <= 0 is the same "yes"): diff^k falls with k, so the draw becomes
                        // the largest k that still swaps and the sweep compares the carried weight with n_lo + k_max
                        const uint4 Sw = philox4x32_10(step, (3u << 8) | (uint32_t)my_r, id_lo, id_hi, p.keys);
                        const double u = (double)Sw.x * U32;
                        const double *row = p.pw + (size_t)my_r * (2 * QECMC_PW_K + 1);
                        int cnt = 0;
#pragma unroll
                        for (int stp = 32; stp >= 1; stp >>= 1) {
                            const int m = cnt + stp;
                            if (m <= 2 * QECMC_PW_K + 1 && u < row[m - 1]) cnt = m;
                        }
                        int t = n + cnt - 1;
                        if (!(row[1] < 1.0) || cnt == 0) t = QECMC_PT_BAD;           // a table that does not fall: exact evaluation
                        else if (cnt == 2 * QECMC_PW_K + 1) t += QECMC_PT_OPEN;      // the draw is below the table's last entry
                        s_t[my_r * NLC + l] = t;
                    }
                }
            }
        }
        __syncthreads();   // (1) every rung's state and sweep inputs are published
        if (p.kind == LK_ALPHA) {
            // alpha ladders: the decision of pair (r, r+1) depends on rung-owned values only (mcmc_alpha.py:117-123)
            if (active && sub == 0 && my_r < Nc - 1) {
                const uint32_t lidx = s_acc[PA_LIDX * NLC + l];
                const uint64_t gid = (uint64_t)p.ladder_offset + lidx;
                const uint32_t step = p.step0 + s_acc[PA_STEP * NLC + l];
                const uint4 Sw = philox4x32_10(step, (3u << 8) | (uint32_t)my_r, (uint32_t)gid, (uint32_t)(gid >> 32), p.keys);
                const double ne_lo = __dadd_rn((double)s_n[my_r * NLC + l], __dmul_rn(p.alpha, (double)s_t[my_r * NLC + l]));
                const double ne_hi = __dadd_rn((double)s_n[(my_r + 1) * NLC + l], __dmul_rn(p.alpha, (double)s_t[(my_r + 1) * NLC + l]));
                if ((double)Sw.x * U32 < pow(p.diff[my_r], __dadd_rn(ne_hi, -ne_lo))) atomicOr(&s_mask[buf * NLC + l], 1u << my_r);
            }
            __syncthreads();
        }
        // =====================================================================================================
        // phase B: the manager of each ladder walks the swap sweep (mcmc.py:96-99) -- nothing else sits between the barriers
        // =====================================================================================================
        if (rwarp == 0) {
            if (is_mgr && s_acc[PA_ACTIVE * NLC + lane] && !s_acc[PA_FIN * NLC + lane]) {
                const int ml = lane;
                const uint32_t lidx = s_acc[PA_LIDX * NLC + ml];
                uint32_t m = 0;
                s_mask[(buf ^ 1) * NLC + ml] = 0;   // the next step's mask (its last readers passed barrier (1) of this step)
                if (p.kind == LK_ALPHA) {
                    m = s_mask[buf * NLC + ml];
                } else if (Nc > 1) {
                    int c_n = s_n[(Nc - 1) * NLC + ml];
                    for (int base = Nc - 2; base >= 0; base -= 8) {
                        // the operands of eight pairs at once (immediate offsets; rungs below 0 read the padding): the walk itself
                        // is then a compare and a select per pair, the eight decisions gathered in a byte
                        const int *qn = s_n + base * NLC + ml, *qt = s_t + base * NLC + ml;
                        int lo8[8], t8[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) { lo8[j] = qn[-j * NLC]; t8[j] = qt[-j * NLC]; }
                        int any = 0;
#pragma unroll
                        for (int j = 0; j < 8; j++) any |= t8[j];
                        if ((any & (QECMC_PT_BAD | QECMC_PT_OPEN)) == 0) {
                            uint32_t bits = 0;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const bool sw = c_n <= t8[j];
                                bits |= sw ? (0x80u >> j) : 0u;
                                c_n = sw ? c_n : lo8[j];
                            }
                            m |= base >= 7 ? bits << (base - 7) : bits >> (7 - base);   // pair base - j sits at bit 7 - j
                        } else {
                            // a draw beyond the power table (or no usable table) in this chunk: pair by pair, exact where needed
                            for (int j = 0; j < 8 && base - j >= 0; j++) {
                                const int i = base - j;
                                const int lo_n = s_n[i * NLC + ml], t = s_t[i * NLC + ml];
                                bool sw = c_n <= t;
                                if (t >= QECMC_PT_BAD && !(t >= QECMC_PT_OPEN && c_n <= t - QECMC_PT_OPEN)) {
                                    const int k = c_n - lo_n;
                                    if (p.kind == LK_DEPOL && k < 0) sw = true;
                                    else {
                                        const uint64_t gid = (uint64_t)p.ladder_offset + lidx;
                                        const uint4 Sw = philox4x32_10(p.step0 + s_acc[PA_STEP * NLC + ml], (3u << 8) | (uint32_t)i, (uint32_t)gid,
                                                                       (uint32_t)(gid >> 32), p.keys);
                                        sw = (double)Sw.x * U32 < numba_pow_dev(p.diff[i], k);
                                    }
                                }
                                m |= (uint32_t)sw << i;
                                c_n = sw ? c_n : lo_n;
                            }
                        }
                    }
                    s_mask[buf * NLC + ml] = m;
                }
                s_acc[PA_STEP * NLC + ml]++;        // Ladder.step calls completed, this one included
            }
            if (defer) {
                // what the deferred accounting decided (it ran beside the block just ended) is known before barrier (2)
                const int ml = lane < NLC ? lane : 0;
                const bool act = is_mgr && s_acc[PA_ACTIVE * NLC + ml] != 0, finl = act && s_acc[PA_FIN * NLC + ml] != 0;
                const uint32_t fins = __ballot_sync(0xFFFFFFFFu, finl);
                const bool any = __any_sync(0xFFFFFFFFu, act && (!finl || s_acc[PA_NEXT * NLC + ml] != 0xFFFFFFFFu));
                if (lane == 0) { s_alive = any; s_anyfin = fins != 0; }
            }
        }
        __syncthreads();   // (2) the swap masks are published
        // =====================================================================================================
        // phase C: every rung thread reads its new column off the swap mask; the thread of the top rung raises the flag of the
        // replica that arrived there, the thread of the bottom rung counts and clears the flag of the one that arrived at the
        // bottom (mcmc.py:100-103) and takes down what PTEQ's accounting of this step needs
        // =====================================================================================================
        int newcol = 0;
        const bool stepping = active && s_acc[PA_FIN * NLC + l] == 0;
        if (stepping) nacc += nacc_s;   // a ladder the deferred accounting ended ran this block for nothing
        nacc_s = 0;
        if (stepping) {
            const uint32_t m = s_mask[buf * NLC + l];
            const int src = pt_src_rung(m, my_r);
            const uint8_t ns = s_slot[(size_t)buf * NREP + src * NLC + l];
            newcol = (int)ns * NLC + l;
            if (sub == 0) {
                s_slot[(size_t)(buf ^ 1) * NREP + my_r * NLC + l] = ns;
                if (my_r == Nc - 1) s_state[NREP + newcol] |= QECMC_PT_FLAG;        // self.chains[-1].flag = 1
                if (my_r == 0) {
                    uint32_t sb1 = s_state[NREP + newcol];
                    uint32_t tops0 = s_acc[PA_TOPS0 * NLC + l];
                    if (sb1 & QECMC_PT_FLAG) { tops0++; sb1 &= ~QECMC_PT_FLAG; s_state[NREP + newcol] = sb1; }
                    s_acc[PA_TOPS0 * NLC + l] = tops0;
                    if (p.acct == ACCT_PTEQ) {
                        s_acc[PA_BOTCLS * NLC + l] = sb1 & 15u;
                        if (p.kind == LK_ALPHA) {
                            s_acc[PA_HA * NLC + l] = (uint32_t)s_n[l];
                            s_acc[PA_HB * NLC + l] = (uint32_t)s_t[l];
                        } else {
                            s_acc[PA_HA * NLC + l] = (uint32_t)s_n[src * NLC + l];
                        }
                        s_acc[PA_PEND * NLC + l] = 1;
                    }
                    if (!defer) account(l);
                }
            }
        }
        if (!defer) {
            __syncthreads();
            if (rwarp == 0) {
                const int ml = lane < NLC ? lane : 0;
                const bool act = is_mgr && s_acc[PA_ACTIVE * NLC + ml] != 0, finl = act && s_acc[PA_FIN * NLC + ml] != 0;
                const uint32_t fins = __ballot_sync(0xFFFFFFFFu, finl);
                const bool any = __any_sync(0xFFFFFFFFu, act && (!finl || s_acc[PA_NEXT * NLC + ml] != 0xFFFFFFFFu));
                if (lane == 0) { s_alive = any; s_anyfin = fins != 0; }
            }
            __syncthreads();
        }
        const bool anyfin = s_anyfin != 0, alive = s_alive != 0;
        if (anyfin) {
            // a ladder ended: write what the caller asked for in rung order, then hand its columns to the next ladder
            const bool mine = active && s_acc[PA_FIN * NLC + l] != 0;
            if (mine && !defer) {
                const uint32_t lidx = s_acc[PA_LIDX * NLC + l];
                if (p.lat_out) {
                    W *o = reinterpret_cast<W *>(p.lat_out) + ((size_t)lidx * Nc + my_r) * g.nw;
                    for (int w = sub; w < g.nw; w += (rung_warp ? 1 : LT)) o[w] = tile[(size_t)w * NREP + newcol];
                }
                if (sub == 0) {
                    const uint32_t s0 = s_state[newcol];
                    if (p.flags_out) p.flags_out[(size_t)lidx * Nc + my_r] = (s_state[NREP + newcol] & QECMC_PT_FLAG) ? 1 : 0;
                    if (p.neff_out) {
                        if (p.kind == LK_ALPHA || !WEIGHTED) p.neff_out[(size_t)lidx * Nc + my_r] = make_int2(e_nz, e_nxy);
                        else p.neff_out[(size_t)lidx * Nc + my_r] = make_int2((int)(s0 >> 20), (int)((s0 & 1023u) + ((s0 >> 10) & 1023u)));
                    }
                }
            }
            __syncthreads();   // (3) rare: the old columns have been read before the new ladder overwrites them
            if (is_mgr && s_acc[PA_ACTIVE * NLC + lane] && s_acc[PA_FIN * NLC + lane]) {
                const uint32_t nxt = s_acc[PA_NEXT * NLC + lane];
                if (nxt != 0xFFFFFFFFu) init_manager(nxt);
                else { s_acc[PA_ACTIVE * NLC + lane] = 0; s_acc[PA_FIN * NLC + lane] = 0; }
            }
            __syncthreads();
            if (mine && s_acc[PA_ACTIVE * NLC + l]) init_ladder(s_acc[PA_LIDX * NLC + l], buf ^ 1);
        }
        buf ^= 1;
        if (!alive) break;
    }
    if (p.counters && nacc) atomicAdd(p.counters + 0, (unsigned long long)nacc);
}

}  // namespace qecmc

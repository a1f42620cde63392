// qecmc_xyz.cuh -- general (x, y, z) noise: Chain_xyz / _update_chain_fast_xyz (src/mcmc.py:106-114,162-173) and
// STDC_droplet_general_noise / STDC_general_noise(_shortest) (decoders.py:325-508).
//
// One thread = one chain, lattice in shared memory as in the other chain kernels.  A proposal's effect on the
// (nx, ny, nz) counts comes from popcounts of the X / Y / Z maps of the touched row words; the accept threshold is
// a 9x9x9 table indexed by (dnx, dny, dnz), made on the host with libm pow in the reference's order
// ((1 * fx**dnx) * fy**dny) * fz**dnz, or factor ** (dnx + dny + dnz) when the sampling chain is a plain Chain.
// The distinct-chain set stores 16-byte entries {fingerprint, packed (nx, ny, nz)}: the weight of a chain under
// general noise needs all three counts, so they ride along instead of squeezing the fingerprint.
#pragma once
#include "qecmc_kernels.cuh"

namespace qecmc {

struct XyzParams {
    Geo gcode, gchain;
    const void *lat0;
    int per_class, droplets, iters;
    int64_t steps, n_chains, chain_offset;
    uint64_t seed, hash_seed;
    unsigned long long *tables;   // [tabs][cap][2]
    uint64_t cap_mask;
    const uint64_t *stab_hash;
    const uint32_t *thr_u;        // [729]
    const double *thr_d;          // [729]
    const double *u_nb;           // replay: [chains][steps*iters][k+1]
    unsigned long long *counters; // [0] accepted [1] offered
};

__device__ __forceinline__ uint64_t pack_xyz(int nx, int ny, int nz) { return (uint64_t)nx | ((uint64_t)ny << 16) | ((uint64_t)nz << 32); }

template <int GEOM, typename W, bool REPLAY>
__global__ void __launch_bounds__(256) xyz_kernel(XyzParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const Geo g = p.gchain;
    W *tile = reinterpret_cast<W *>(smem);
    const int64_t local = (int64_t)blockIdx.x * T + tid;
    if (local >= p.n_chains) return;
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
    if (!p.per_class) to_class_rt<W>(p.gcode, lat, eq);
    int nx, ny, nz;
    lat_count_xyz<W>(g, lat, nx, ny, nz);
    uint64_t h = lat_hash<W>(g, lat, p.hash_seed);
    unsigned long long *table = p.tables + (uint64_t)tab * (p.cap_mask + 1) * 2;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const uint32_t cl = (uint32_t)gchain, chh = (uint32_t)((uint64_t)gchain >> 32);
    constexpr int NU = NumUpd<GEOM>::value;
    constexpr int K = NumDraws<GEOM>::value;
    unsigned long long nacc = 0, noff = 0;
    bool dirty = true;
    int left = p.iters;
    const uint64_t tsteps = (uint64_t)p.steps * (uint64_t)p.iters;
    const double *u = REPLAY ? p.u_nb + (uint64_t)gchain * tsteps * (K + 1) : nullptr;
    uint4 r = make_uint4(0, 0, 0, 0);
    uint32_t c0 = 0;
    for (uint64_t t = 0; t < tsteps; t++) {
        int row, col, op, idx;
        uint32_t r_acc = 0;
        double u_acc = 0.0;
        if (REPLAY) {
            propose_replay<GEOM>(g, u, row, col, op);
            idx = rco_to_idx<GEOM>(g, row, col, op);
            u_acc = u[K];
            u += K + 1;
        } else {
            if ((t & 1) == 0) r = philox4x32_10(c0++, 0u, cl, chh, k0, k1);
            idx = (int)__umulhi((t & 1) ? r.z : r.x, (uint32_t)g.nstab);
            r_acc = (t & 1) ? r.w : r.y;
            idx_to_rco<GEOM>(g, idx, row, col, op);
        }
        Upd<W> up;
        decode<GEOM, W>(g, row, col, op, up);
        W nv[NU];
        int dx = 0, dy = 0, dz = 0;
#pragma unroll
        for (int i = 0; i < NU; i++) {
            W o = lat.get(up.w[i]);
            nv[i] = (W)(o ^ up.m[i]);
            dx += popc(xmap(nv[i])) - popc(xmap(o));
            dy += popc(ymap(nv[i])) - popc(ymap(o));
            dz += popc(zmap(nv[i])) - popc(zmap(o));
        }
        const int ti = (dx + 4) * 81 + (dy + 4) * 9 + (dz + 4);
        const bool acc = REPLAY ? (u_acc < p.thr_d[ti]) : (r_acc <= p.thr_u[ti]);
        if (acc) {
#pragma unroll
            for (int i = 0; i < NU; i++) lat.set(up.w[i], nv[i]);
            nx += dx; ny += dy; nz += dz;
            h ^= p.stab_hash[idx];
            dirty = true;
            nacc++;
        }
        if (--left == 0) {
            left = p.iters;
            if (dirty) {  // an unchanged state is already in the set
                const uint64_t key = h | (1ull << 63);
                uint64_t slot = (key >> 8) & p.cap_mask;
                while (true) {
                    unsigned long long cur = __ldcg(table + 2 * slot);
                    if (cur == key) break;
                    if (cur == 0ull) {
                        unsigned long long prev = atomicCAS(table + 2 * slot, 0ull, (unsigned long long)key);
                        if (prev == 0ull) { table[2 * slot + 1] = pack_xyz(nx, ny, nz); break; }
                        if (prev == key) break;
                    }
                    slot = (slot + 1) & p.cap_mask;
                }
                noff++;
                dirty = false;
            }
        }
    }
    atomicAdd(p.counters + 0, nacc);
    atomicAdd(p.counters + 1, noff);
}

// Chain_xyz.update_chain_fast on a batch of chains: lattices [chains][nw] in/out (src/mcmc.py:113-114,162-173)
struct XyzChainParams {
    Geo g;
    void *lat;
    int64_t chains, iters;
    uint64_t seed, offset;
    const uint32_t *thr_u;
    const double *thr_d;
    const double *u;   // replay: [chains][iters][k+1]
    unsigned long long *counters;
};

template <int GEOM, typename W, bool REPLAY> __global__ void xyz_chain_kernel(XyzChainParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    W *tile = reinterpret_cast<W *>(smem);
    int64_t ch = (int64_t)blockIdx.x * T + tid;
    if (ch >= p.chains) return;
    const Geo g = p.g;
    SmemLat<W> lat{tile + tid, T};
    W *gl = reinterpret_cast<W *>(p.lat) + ch * g.nw;
    for (int w = 0; w < g.nw; w++) lat.set(w, gl[w]);
    constexpr int K = NumDraws<GEOM>::value;
    constexpr int NU = NumUpd<GEOM>::value;
    unsigned long long nacc = 0;
    for (int64_t t = 0; t < p.iters; t++) {
        int row, col, op;
        uint32_t r_acc = 0;
        double u_acc = 0.0;
        if (REPLAY) {
            const double *u = p.u + (ch * p.iters + t) * (K + 1);
            propose_replay<GEOM>(g, u, row, col, op);
            u_acc = u[K];
        } else {
            uint64_t tg = p.offset + (uint64_t)t, call = tg >> 1;
            uint4 r = philox4x32_10((uint32_t)call, (uint32_t)(call >> 32), (uint32_t)ch, (uint32_t)((uint64_t)ch >> 32),
                                    (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
            int idx = (int)__umulhi((tg & 1) ? r.z : r.x, (uint32_t)g.nstab);
            r_acc = (tg & 1) ? r.w : r.y;
            idx_to_rco<GEOM>(g, idx, row, col, op);
        }
        Upd<W> up;
        decode<GEOM, W>(g, row, col, op, up);
        W nv[NU];
        int dx = 0, dy = 0, dz = 0;
#pragma unroll
        for (int i = 0; i < NU; i++) {
            W o = lat.get(up.w[i]);
            nv[i] = (W)(o ^ up.m[i]);
            dx += popc(xmap(nv[i])) - popc(xmap(o));
            dy += popc(ymap(nv[i])) - popc(ymap(o));
            dz += popc(zmap(nv[i])) - popc(zmap(o));
        }
        const int ti = (dx + 4) * 81 + (dy + 4) * 9 + (dz + 4);
        if (REPLAY ? (u_acc < p.thr_d[ti]) : (r_acc <= p.thr_u[ti])) {
#pragma unroll
            for (int i = 0; i < NU; i++) lat.set(up.w[i], nv[i]);
            nacc++;
        }
    }
    for (int w = 0; w < g.nw; w++) gl[w] = lat.get(w);
    if (p.counters && nacc) atomicAdd(p.counters, nacc);
}

// One block per (syndrome, class): weighted length w = sum over components with n_i > 0 of beta_i n_i
// (decoders.py:406), Z_all = sum exp(-w), Z_shortest = the same over chains with np.isclose(w, min w).
static __global__ void table_xyz_kernel(const unsigned long long *__restrict__ tables, uint64_t cap, double bx, double by,
                                        double bz, double *__restrict__ Z_all, double *__restrict__ Z_short,
                                        unsigned long long *__restrict__ distinct, unsigned long long *distinct_total)
{
    __shared__ double s_a[256], s_b[256];
    __shared__ unsigned long long s_c[256];
    const unsigned long long *tab = tables + (uint64_t)blockIdx.x * cap * 2;
    auto weight_of = [&](unsigned long long v) {
        double nx = (double)(v & 0xFFFF), ny = (double)((v >> 16) & 0xFFFF), nz = (double)((v >> 32) & 0xFFFF), w = 0.0;
        if (nx > 0) w += bx * nx;
        if (ny > 0) w += by * ny;
        if (nz > 0) w += bz * nz;
        return w;
    };
    double z = 0, wmin = INFINITY;
    unsigned long long cnt = 0;
    for (uint64_t i = threadIdx.x; i < cap; i += blockDim.x)
        if (tab[2 * i]) {
            double w = weight_of(tab[2 * i + 1]);
            z += exp(-w);
            wmin = fmin(wmin, w);
            cnt++;
        }
    s_a[threadIdx.x] = z; s_b[threadIdx.x] = wmin; s_c[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            s_a[threadIdx.x] += s_a[threadIdx.x + s];
            s_b[threadIdx.x] = fmin(s_b[threadIdx.x], s_b[threadIdx.x + s]);
            s_c[threadIdx.x] += s_c[threadIdx.x + s];
        }
        __syncthreads();
    }
    const double gmin = s_b[0];
    if (threadIdx.x == 0) {
        Z_all[blockIdx.x] = s_a[0];
        if (distinct) distinct[blockIdx.x] = s_c[0];
        if (distinct_total) atomicAdd(distinct_total, s_c[0]);
    }
    __syncthreads();
    double zs = 0;
    for (uint64_t i = threadIdx.x; i < cap; i += blockDim.x)
        if (tab[2 * i]) {
            double w = weight_of(tab[2 * i + 1]);
            if (fabs(w - gmin) <= 1e-8 + 1e-5 * fabs(gmin)) zs += exp(-w);  // np.isclose defaults
        }
    s_a[threadIdx.x] = zs;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) s_a[threadIdx.x] += s_a[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) Z_short[blockIdx.x] = s_a[0];
}

}  // namespace qecmc

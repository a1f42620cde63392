// qecmc_workload.cu -- the steps either side of the decoders in the reference's workload loop
// (generate_data.py:53-261), so that a batch of syndromes never leaves the GPU between "draw an error"
// and "count the failures":
//   generate_random_error   toric_model.py:15-24 (p_error form), planar_model.py:18-37,
//                           rotated_surface_model.py:25-38, xzzx_model.py:16-29 (p_x, p_y, p_z form)
//   define_equivalence_class  eq_true, generate_data.py:122
//   apply_random_logical    "no trace of seed", generate_data.py:131
//   argmax / argmin != eq_true  generate_data.py:137-201
// Native draws are Philox4x32-10 (stream tags below keep them disjoint from the chain streams); replay
// takes the reference's own uniforms and reproduces its lattices bit-exactly.
#include "qecmc_internal.h"

namespace {

constexpr uint32_t TAG_ERRORS = 0xE0000000u;   // counter word 1 of the error-generation stream
constexpr uint32_t TAG_LOGICAL = 0xE1000000u;  // ... of the random-logical stream

struct NoiseParams {
    int geom, L, nsites, toric_form;
    int64_t total;          // S * nsites bytes
    double p_error;         // toric form
    double p_z, p_zx, p_zxy;  // cumulative sums in the reference's order of evaluation
    uint32_t t_err, t_z, t_zx, t_zxy;  // native: word < t  <=>  word / 2^32 < p
    uint32_t k0, k1;
    const double *u;        // replay [S][nsites]
    const uint8_t *pauli;   // replay, toric form [S][nsites]
};

__device__ __forceinline__ bool planar_unused(const NoiseParams &p, int site)
{
    // planar_model.py:36-37: layer 1 has no last row / last column
    if (p.geom != PLANAR) return false;
    const int LL = p.L * p.L;
    if (site < LL) return false;
    const int r = (site - LL) / p.L, c = (site - LL) % p.L;
    return r == p.L - 1 || c == p.L - 1;
}

// four consecutive bytes of the flattened [S][nsites] output per thread: one 32-bit store
__global__ void gen_errors_kernel(NoiseParams p, uint8_t *__restrict__ qm)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = t * 4;
    if (i0 >= p.total) return;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!p.u) {
        uint4 r = philox4x32_10((uint32_t)t, TAG_ERRORS | (uint32_t)((uint64_t)t >> 32), 0u, 0u, p.k0, p.k1);
        w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
        if (p.toric_form) {
            uint4 q = philox4x32_10((uint32_t)t, TAG_ERRORS | (uint32_t)((uint64_t)t >> 32), 1u, 0u, p.k0, p.k1);
            w[4] = q.x; w[5] = q.y; w[6] = q.z; w[7] = q.w;
        }
    }
    uint32_t packed = 0;
    int site = (int)(i0 % p.nsites);
    for (int j = 0; j < 4 && i0 + j < p.total; j++) {
        int q = 0;
        if (p.u) {
            const double r = p.u[i0 + j];
            if (p.toric_form) q = r < p.p_error ? (int)p.pauli[i0 + j] : 0;
            else if (r < p.p_z) q = 3;
            else if (p.p_z < r && r < p.p_zx) q = 1;
            else if (p.p_zx < r && r < p.p_zxy) q = 2;
        } else if (p.toric_form) {
            if (w[j] < p.t_err) q = 1 + (int)__umulhi(w[4 + j], 3u);
        } else {
            const uint32_t r = w[j];
            q = r < p.t_z ? 3 : r < p.t_zx ? 1 : r < p.t_zxy ? 2 : 0;
        }
        if (planar_unused(p, site)) q = 0;
        packed |= (uint32_t)q << (8 * j);
        if (++site == p.nsites) site = 0;
    }
    if (i0 + 4 <= p.total) *reinterpret_cast<uint32_t *>(qm + i0) = packed;
    else for (int j = 0; i0 + j < p.total; j++) qm[i0 + j] = (uint8_t)(packed >> (8 * j));
}

// one syndrome's lattice as 64-bit row words in local memory (cold kernels only)
struct LocalLat {
    uint64_t w[64];
    __device__ uint64_t get(int i) const { return w[i]; }
    __device__ void set(int i, uint64_t v) { w[i] = v; }
    __device__ void load(const uint8_t *qm, const Geo &g)
    {
        for (int i = 0; i < g.nw; i++) w[i] = pack_row<uint64_t>(qm + (size_t)i * g.L, g.L);
    }
    __device__ void store(uint8_t *qm, const Geo &g) const
    {
        for (int i = 0; i < g.nw; i++) unpack_row<uint64_t>(w[i], qm + (size_t)i * g.L, g.L);
    }
};

template <int GEOM> __global__ void class_kernel(Geo g, const uint8_t *__restrict__ qm, int64_t S, int32_t *__restrict__ cls, int *bad)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint8_t *q = qm + s * g.nsites;
    int b = 0;
    for (int i = 0; i < g.nsites; i++) b |= q[i] > 3;
    if (b) { *bad = 1; return; }
    LocalLat lat;
    lat.load(q, g);
    cls[s] = lat_class<GEOM, uint64_t>(g, lat);
}

// _apply_random_logical draw order: toric_model.py:228-253 (both layer operators first, then the
// positions layer by layer), planar_model.py:271-288, rotated_surface_model.py:331-346, xzzx_model.py:340-357
struct DrawRng {
    const double *u;   // replay: this syndrome's uniforms
    uint4 r;
    int k;
    __device__ double next()
    {
        if (u) return u[k++];
        const uint32_t wd = k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w;
        k++;
        return (double)wd * (1.0 / 4294967296.0);
    }
};

template <int GEOM>
__global__ void random_logical_kernel(Geo g, uint8_t *__restrict__ qm, int64_t S, uint32_t k0, uint32_t k1, const double *__restrict__ u,
                                      int32_t *__restrict__ ops_out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    uint8_t *q = qm + s * g.nsites;
    LocalLat lat;
    lat.load(q, g);
    const int nl = GEOM == TORIC ? 2 : 1;
    int op[2] = {0, 0}, xp[2] = {0, 0}, zp[2] = {0, 0};
    // native: at most 2 + 4 draws; two Philox calls cover them
    DrawRng a{u ? u + s * 6 : nullptr, philox4x32_10((uint32_t)s, TAG_LOGICAL | (uint32_t)((uint64_t)s >> 32), 0u, 0u, k0, k1), 0};
    for (int l = 0; l < nl; l++) op[l] = (int)(a.next() * 4);
    DrawRng b{u ? u + s * 6 + nl : nullptr, philox4x32_10((uint32_t)s, TAG_LOGICAL | (uint32_t)((uint64_t)s >> 32), 1u, 0u, k0, k1), 0};
    for (int l = 0; l < nl; l++) {
        if (op[l] == 1 || op[l] == 2) xp[l] = (int)(b.next() * g.L);
        if (op[l] == 3 || op[l] == 2) zp[l] = (int)(b.next() * g.L);
    }
    for (int l = 0; l < nl; l++) lat_apply_logical<GEOM, uint64_t>(g, lat, op[l], l, xp[l], zp[l]);
    lat.store(q, g);
    if (ops_out) { ops_out[2 * s] = op[0]; ops_out[2 * s + 1] = op[1]; }
}

// np.argmax / np.argmin (first extremum; a NaN wins, as in numpy) against eq_true
template <typename T>
__global__ void failures_kernel(const T *__restrict__ distr, int n_eq, int64_t S, int use_argmin, const int32_t *__restrict__ eq_true,
                                int32_t *__restrict__ choice, unsigned long long *failures)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool fail = false;
    if (s < S) {
        const T *d = distr + s * n_eq;
        int best = 0;
        double bv = (double)d[0];
        for (int e = 1; e < n_eq && bv == bv; e++) {
            const double v = (double)d[e];
            if (v != v || (use_argmin ? v < bv : v > bv)) { best = e; bv = v; }
        }
        if (choice) choice[s] = best;
        fail = best != eq_true[s];
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, fail);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(failures, (unsigned long long)__popc(m));
}

uint32_t thr32(double p)
{
    // word < t  <=>  word / 2^32 < p  for word in [0, 2^32)
    if (!(p > 0)) return 0u;
    double x = ceil(p * 4294967296.0);
    return x >= 4294967296.0 ? 0xFFFFFFFFu : (uint32_t)x;
}

int check_noise(const qecmc_noise_cfg *cfg)
{
    if (!cfg) return set_err(QECMC_ERR_ARG, "cfg is NULL");
    QTRY(check_geom(cfg->geom, cfg->L));
    if (cfg->toric_form) {
        if (!(cfg->p_error >= 0 && cfg->p_error <= 1)) return set_err(QECMC_ERR_ARG, "p_error outside [0,1]");
        if (cfg->u && !cfg->pauli) return set_err(QECMC_ERR_ARG, "replay of the p_error form needs the pauli draws too");
    } else if (!(cfg->p_x >= 0 && cfg->p_y >= 0 && cfg->p_z >= 0 && cfg->p_x + cfg->p_y + cfg->p_z <= 1.0 + 1e-12))
        return set_err(QECMC_ERR_ARG, "p_x, p_y, p_z must be >= 0 and sum to <= 1");
    return 0;
}

int gen_errors_dev(qecmc_ctx *c, const qecmc_noise_cfg *cfg, int64_t S, uint8_t *d_qm)
{
    Geo g = make_geo(cfg->geom, cfg->L);
    NoiseParams p;
    memset(&p, 0, sizeof(p));
    p.geom = cfg->geom;
    p.L = cfg->L;
    p.nsites = g.nsites;
    p.toric_form = cfg->toric_form;
    p.total = S * g.nsites;
    p.p_error = cfg->p_error;
    p.p_z = cfg->p_z;
    p.p_zx = cfg->p_z + cfg->p_x;
    p.p_zxy = cfg->p_z + cfg->p_x + cfg->p_y;
    p.t_err = thr32(cfg->p_error);
    p.t_z = thr32(p.p_z);
    p.t_zx = thr32(p.p_zx);
    p.t_zxy = thr32(p.p_zxy);
    p.k0 = (uint32_t)cfg->seed;
    p.k1 = (uint32_t)(cfg->seed >> 32);
    p.u = cfg->u;
    p.pauli = cfg->pauli;
    const int64_t threads = (p.total + 3) / 4;
    gen_errors_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(p, d_qm);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int class_dev(qecmc_ctx *c, int geom, int L, const uint8_t *d_qm, int64_t S, int32_t *d_cls)
{
    Geo g = make_geo(geom, L);
    QTRY(c->scratch.ensure(sizeof(int)));
    CUDA_OK(cudaMemsetAsync(c->scratch.p, 0, sizeof(int), c->stream));
    const unsigned grid = (unsigned)((S + 127) / 128);
    int *bad = (int *)c->scratch.p;
    switch (geom) {
    case TORIC: class_kernel<TORIC><<<grid, 128, 0, c->stream>>>(g, d_qm, S, d_cls, bad); break;
    case PLANAR: class_kernel<PLANAR><<<grid, 128, 0, c->stream>>>(g, d_qm, S, d_cls, bad); break;
    case ROTATED: class_kernel<ROTATED><<<grid, 128, 0, c->stream>>>(g, d_qm, S, d_cls, bad); break;
    default: class_kernel<XZZX><<<grid, 128, 0, c->stream>>>(g, d_qm, S, d_cls, bad); break;
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    int b = 0;
    CUDA_OK(cudaMemcpyAsync(&b, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (b) return set_err(QECMC_ERR_ARG, "qubit_matrix holds values outside 0..3");
    return 0;
}

int logical_dev(qecmc_ctx *c, int geom, int L, uint8_t *d_qm, int64_t S, uint64_t seed, const double *d_u, int32_t *d_ops)
{
    Geo g = make_geo(geom, L);
    const unsigned grid = (unsigned)((S + 127) / 128);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    switch (geom) {
    case TORIC: random_logical_kernel<TORIC><<<grid, 128, 0, c->stream>>>(g, d_qm, S, k0, k1, d_u, d_ops); break;
    case PLANAR: random_logical_kernel<PLANAR><<<grid, 128, 0, c->stream>>>(g, d_qm, S, k0, k1, d_u, d_ops); break;
    case ROTATED: random_logical_kernel<ROTATED><<<grid, 128, 0, c->stream>>>(g, d_qm, S, k0, k1, d_u, d_ops); break;
    default: random_logical_kernel<XZZX><<<grid, 128, 0, c->stream>>>(g, d_qm, S, k0, k1, d_u, d_ops); break;
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int failures_dev(qecmc_ctx *c, const void *d_distr, int dtype, int use_argmin, int n_eq, int64_t S, const int32_t *d_eq_true,
                 int32_t *d_choice, int64_t *failures)
{
    QTRY(c->counters.ensure(8 * sizeof(unsigned long long)));
    unsigned long long *cnt = (unsigned long long *)c->counters.p + 7;
    CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
    const unsigned grid = (unsigned)((S + 127) / 128);
    if (dtype == QECMC_DISTR_F64)
        failures_kernel<double><<<grid, 128, 0, c->stream>>>((const double *)d_distr, n_eq, S, use_argmin, d_eq_true, d_choice, cnt);
    else
        failures_kernel<uint8_t><<<grid, 128, 0, c->stream>>>((const uint8_t *)d_distr, n_eq, S, use_argmin, d_eq_true, d_choice, cnt);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    unsigned long long h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (failures) *failures = (int64_t)h;
    return 0;
}

// host-buffer staging: a context-owned device buffer per role
struct Stage {
    qecmc_ctx *c;
    int up(DevBuf &b, const void *host, size_t bytes)
    {
        QTRY(b.ensure(bytes));
        CUDA_OK(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    int down(void *host, const void *dev, size_t bytes)
    {
        CUDA_OK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
        return 0;
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int qecmc_generate_errors_dev(qecmc_ctx *c, const qecmc_noise_cfg *cfg, int64_t S, uint8_t *d_qm, int32_t *d_eq_true)
{
    if (!c || !d_qm) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_noise(cfg));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    c->launches = 0;
    QTRY(gen_errors_dev(c, cfg, S, d_qm));
    if (d_eq_true) QTRY(class_dev(c, cfg->geom, cfg->L, d_qm, S, d_eq_true));
    return 0;
}

extern "C" int qecmc_generate_errors(qecmc_ctx *c, const qecmc_noise_cfg *cfg, int64_t S, uint8_t *qm, int32_t *eq_true)
{
    if (!c || !qm) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_noise(cfg));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t n = (size_t)S * make_geo(cfg->geom, cfg->L).nsites;
    Stage st{c};
    qecmc_noise_cfg d = *cfg;
    if (cfg->u) { QTRY(st.up(c->replay_a, cfg->u, n * sizeof(double))); d.u = (const double *)c->replay_a.p; }
    if (cfg->u && cfg->pauli) { QTRY(st.up(c->replay_b, cfg->pauli, n)); d.pauli = (const uint8_t *)c->replay_b.p; }
    QTRY(c->qm_in.ensure(n));
    QTRY(c->out_i32.ensure((size_t)S * sizeof(int32_t)));
    QTRY(qecmc_generate_errors_dev(c, &d, S, (uint8_t *)c->qm_in.p, eq_true ? (int32_t *)c->out_i32.p : nullptr));
    QTRY(st.down(qm, c->qm_in.p, n));
    if (eq_true) QTRY(st.down(eq_true, c->out_i32.p, (size_t)S * sizeof(int32_t)));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qecmc_define_equivalence_class_dev(qecmc_ctx *c, int32_t geom, int32_t L, const uint8_t *d_qm, int64_t S, int32_t *d_cls)
{
    if (!c || !d_qm || !d_cls) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_geom(geom, L));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    return class_dev(c, geom, L, d_qm, S, d_cls);
}

extern "C" int qecmc_define_equivalence_class(qecmc_ctx *c, int32_t geom, int32_t L, const uint8_t *qm, int64_t S, int32_t *cls)
{
    if (!c || !qm || !cls) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_geom(geom, L));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t n = (size_t)S * make_geo(geom, L).nsites;
    Stage st{c};
    QTRY(st.up(c->qm_in, qm, n));
    QTRY(c->out_i32.ensure((size_t)S * sizeof(int32_t)));
    QTRY(class_dev(c, geom, L, (const uint8_t *)c->qm_in.p, S, (int32_t *)c->out_i32.p));
    QTRY(st.down(cls, c->out_i32.p, (size_t)S * sizeof(int32_t)));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qecmc_apply_random_logical_dev(qecmc_ctx *c, int32_t geom, int32_t L, uint8_t *d_qm, int64_t S, uint64_t seed,
                                              const double *d_u, int32_t *d_ops)
{
    if (!c || !d_qm) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_geom(geom, L));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    return logical_dev(c, geom, L, d_qm, S, seed, d_u, d_ops);
}

extern "C" int qecmc_apply_random_logical(qecmc_ctx *c, int32_t geom, int32_t L, uint8_t *qm, int64_t S, uint64_t seed, const double *u,
                                          int32_t *ops)
{
    if (!c || !qm) return set_err(QECMC_ERR_ARG, "NULL argument");
    QTRY(check_geom(geom, L));
    if (S <= 0) return set_err(QECMC_ERR_ARG, "S must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    const size_t n = (size_t)S * make_geo(geom, L).nsites;
    Stage st{c};
    QTRY(st.up(c->qm_in, qm, n));
    const double *d_u = nullptr;
    if (u) { QTRY(st.up(c->replay_a, u, (size_t)S * 6 * sizeof(double))); d_u = (const double *)c->replay_a.p; }
    QTRY(c->out_i32.ensure((size_t)S * 2 * sizeof(int32_t)));
    QTRY(logical_dev(c, geom, L, (uint8_t *)c->qm_in.p, S, seed, d_u, ops ? (int32_t *)c->out_i32.p : nullptr));
    QTRY(st.down(qm, c->qm_in.p, n));
    if (ops) QTRY(st.down(ops, c->out_i32.p, (size_t)S * 2 * sizeof(int32_t)));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qecmc_count_failures_dev(qecmc_ctx *c, const void *d_distr, int32_t dtype, int32_t use_argmin, int32_t n_eq, int64_t S,
                                        const int32_t *d_eq_true, int32_t *d_choice, int64_t *failures)
{
    if (!c || !d_distr || !d_eq_true) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (dtype != QECMC_DISTR_F64 && dtype != QECMC_DISTR_U8) return set_err(QECMC_ERR_ARG, "dtype must be QECMC_DISTR_F64 or QECMC_DISTR_U8");
    if (S <= 0 || n_eq <= 0) return set_err(QECMC_ERR_ARG, "S and n_eq must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    return failures_dev(c, d_distr, dtype, use_argmin, n_eq, S, d_eq_true, d_choice, failures);
}

extern "C" int qecmc_count_failures(qecmc_ctx *c, const void *distr, int32_t dtype, int32_t use_argmin, int32_t n_eq, int64_t S,
                                    const int32_t *eq_true, int32_t *choice, int64_t *failures)
{
    if (!c || !distr || !eq_true) return set_err(QECMC_ERR_ARG, "NULL argument");
    if (dtype != QECMC_DISTR_F64 && dtype != QECMC_DISTR_U8) return set_err(QECMC_ERR_ARG, "dtype must be QECMC_DISTR_F64 or QECMC_DISTR_U8");
    if (S <= 0 || n_eq <= 0) return set_err(QECMC_ERR_ARG, "S and n_eq must be > 0");
    CUDA_OK(cudaSetDevice(c->device));
    Stage st{c};
    const size_t esz = dtype == QECMC_DISTR_F64 ? 8 : 1;
    QTRY(st.up(c->out_f64, distr, (size_t)S * n_eq * esz));
    QTRY(st.up(c->out_u32, eq_true, (size_t)S * sizeof(int32_t)));
    QTRY(c->out_i32.ensure((size_t)S * sizeof(int32_t)));
    QTRY(failures_dev(c, c->out_f64.p, dtype, use_argmin, n_eq, S, (const int32_t *)c->out_u32.p, choice ? (int32_t *)c->out_i32.p : nullptr,
                      failures));
    if (choice) {
        QTRY(st.down(choice, c->out_i32.p, (size_t)S * sizeof(int32_t)));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

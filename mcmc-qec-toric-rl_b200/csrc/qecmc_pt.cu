// qecmc_pt.cu -- host side of the rung-major tempering kernel (qecmc_pt.cuh): CTA shape and shared-memory plan,
// occupancy, launch.  The drivers in qecmc_ladder.cu decide when a call runs on it (native draws, no snapshots).
#include "qecmc_internal.h"
#include "qecmc_pt.cuh"

using namespace qecmc;

namespace {

template <int GEOM, typename W, bool WEIGHTED, int NLC> int plan_one(qecmc_ctx *c, const LadderParams &lp, int lt, PtPlan *out)
{
    constexpr bool TABLE = GEOM == ROTATED || GEOM == XZZX, TABLE2 = (GEOM == TORIC || GEOM == PLANAR) && !WEIGHTED;
    const PtLayout lay = pt_layout<W>(lp.g, lp.Nc, NLC, lt, lp.p_logical != 0.0, TABLE, TABLE2, lp.iters);
    out->NLC = 0;
    if (lay.T > 1024 || lay.total > c->prop.sharedMemPerBlockOptin) return 0;
    int nb = 0;
    out->small_cta = 0;
    if constexpr (sizeof(W) == 8 && NLC == 16) {
        if (lay.T <= 448) {   // the instantiation compiled for CTAs of at most 448 threads (72 registers)
            out->small_cta = 1;
            CUDA_OK(cudaFuncSetAttribute(pt_kernel<GEOM, W, WEIGHTED, NLC, 448>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pt_kernel<GEOM, W, WEIGHTED, NLC, 448>, lay.T, lay.total));
        }
    }
    if (!out->small_cta) {
        CUDA_OK(cudaFuncSetAttribute(pt_kernel<GEOM, W, WEIGHTED, NLC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pt_kernel<GEOM, W, WEIGHTED, NLC>, lay.T, lay.total));
    }
    if (nb < 1) return 0;
    out->NLC = NLC;
    out->T = lay.T;
    out->lt = lt;
    out->smem = lay.total;
    out->blocks_per_sm = nb;
    out->max_grid = nb * c->prop.multiProcessorCount;
    return 0;
}

template <int GEOM, typename W, bool WEIGHTED> int plan_gw(qecmc_ctx *c, const LadderParams &lp, int lt, PtPlan *out)
{
    // 32 ladders per CTA with 32-bit row words (bank = ladder), 16 with 64-bit ones (a half-warp fills a wavefront); the
    // smaller CTA also serves ladders whose rung warps would not fit 1024 threads
    PtPlan a, b;
    a.NLC = b.NLC = 0;
    if constexpr (sizeof(W) == 4) QTRY((plan_one<GEOM, W, WEIGHTED, 32>(c, lp, lt, &a)));
    QTRY((plan_one<GEOM, W, WEIGHTED, 16>(c, lp, lt, &b)));
    *out = (a.NLC && a.blocks_per_sm * a.NLC >= b.blocks_per_sm * b.NLC) ? a : b;
    if (!out->NLC) return set_err(QECMC_ERR_UNSUPPORTED, "the ladders' lattices do not fit in shared memory");
    return 0;
}

template <int GEOM, typename W, bool WEIGHTED, int NLC> void launch_one(qecmc_ctx *c, const PtParams &p, const PtPlan &pl, int grid)
{
    if constexpr (sizeof(W) == 8 && NLC == 16) {
        if (pl.small_cta) {
            pt_kernel<GEOM, W, WEIGHTED, NLC, 448><<<grid, pl.T, pl.smem, c->stream>>>(p);
            return;
        }
    }
    pt_kernel<GEOM, W, WEIGHTED, NLC><<<grid, pl.T, pl.smem, c->stream>>>(p);
}

template <int GEOM, typename W, bool WEIGHTED> int launch_gw(qecmc_ctx *c, const PtParams &p, const PtPlan &pl, int grid)
{
    if (pl.NLC == 32) {
        if constexpr (sizeof(W) == 4) launch_one<GEOM, W, WEIGHTED, 32>(c, p, pl, grid);
        else return set_err(QECMC_ERR_UNSUPPORTED, "internal: 32 ladders per CTA need 32-bit row words");
    } else {
        launch_one<GEOM, W, WEIGHTED, 16>(c, p, pl, grid);
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

#define QECMC_PT_DISPATCH(FN, ...)                                                                      \
    do {                                                                                                \
        const bool wide = lp.g.L > 16, wt = lp.kind != LK_DEPOL;                                        \
        switch (lp.g.geom) {                                                                            \
        case TORIC:                                                                                     \
            if (wide) return wt ? FN<TORIC, uint64_t, true>(__VA_ARGS__) : FN<TORIC, uint64_t, false>(__VA_ARGS__);       \
            return wt ? FN<TORIC, uint32_t, true>(__VA_ARGS__) : FN<TORIC, uint32_t, false>(__VA_ARGS__);                 \
        case PLANAR:                                                                                    \
            if (wide) return wt ? FN<PLANAR, uint64_t, true>(__VA_ARGS__) : FN<PLANAR, uint64_t, false>(__VA_ARGS__);     \
            return wt ? FN<PLANAR, uint32_t, true>(__VA_ARGS__) : FN<PLANAR, uint32_t, false>(__VA_ARGS__);               \
        case ROTATED:                                                                                   \
            if (wide) return wt ? FN<ROTATED, uint64_t, true>(__VA_ARGS__) : FN<ROTATED, uint64_t, false>(__VA_ARGS__);   \
            return wt ? FN<ROTATED, uint32_t, true>(__VA_ARGS__) : FN<ROTATED, uint32_t, false>(__VA_ARGS__);             \
        default:                                                                                        \
            if (wide) return wt ? FN<XZZX, uint64_t, true>(__VA_ARGS__) : FN<XZZX, uint64_t, false>(__VA_ARGS__);         \
            return wt ? FN<XZZX, uint32_t, true>(__VA_ARGS__) : FN<XZZX, uint32_t, false>(__VA_ARGS__);                   \
        }                                                                                               \
    } while (0)

}  // namespace

int qecmc_pt_plan(qecmc_ctx *c, const LadderParams &lp, PtPlan *out)
{
    // lanes per top-rung replica: a depolarizing top rung that accepts everything (mcmc.py:30) is a set of commuting XOR masks,
    // cheap enough for 4 lanes (fewer top warps: rotated d=25 1.32e11 against 1.27e11 steps/s with 8); a top rung that
    // evaluates its proposals is the critical path of the step and takes 8 (XZZX d=21 biased: 6.0e10 against 5.0e10 with 4)
    int lt = c->dbg_pt_lt > 0 ? c->dbg_pt_lt : (lp.kind == LK_DEPOL && lp.top_accept_all ? 4 : 8);
    if (lt > 32) lt = 32;
    while (lt & (lt - 1)) lt &= lt - 1;   // power of two
    if (lt < 1) lt = 1;
    QECMC_PT_DISPATCH(plan_gw, c, lp, lt, out);
}

int qecmc_pt_launch(qecmc_ctx *c, const LadderParams &lp, const PtPlan &pl, int grid, uint32_t step0, void *hist, int64_t hist_stride)
{
    if (c->dbg_pt_grid > 0 && grid > c->dbg_pt_grid) grid = c->dbg_pt_grid;   // tests: few CTAs, so that ladders queue up
    if (grid < 1 || grid > pl.max_grid) return set_err(QECMC_ERR_ARG, "internal: tempering grid %d outside [1, %d]", grid, pl.max_grid);
    if ((uint64_t)(lp.steps + step0) * (uint64_t)lp.iters >= (1ull << 32))
        return set_err(QECMC_ERR_UNSUPPORTED, "steps * iters must be < 2^32 (draw positions are 32-bit)");
    PtParams p;
    memset(&p, 0, sizeof(p));
    p.g = lp.g;
    p.kind = lp.kind; p.Nc = lp.Nc; p.iters = lp.iters; p.acct = lp.acct;
    p.lt = pl.lt;
    p.p_logical = lp.p_logical;
    p.top_accept_all = lp.top_accept_all;
    p.thr_u = lp.thr_u; p.thr_top_d = lp.thr_top_d; p.diff = lp.diff; p.pw = lp.pw;
    p.alpha = lp.alpha; p.wtab = lp.wtab; p.desc2 = lp.desc2;
    p.keys = lp.keys;
    p.n_ladders = lp.n_ladders; p.ladder_offset = lp.ladder_offset; p.steps = lp.steps;
    p.step0 = step0;
    p.lat_in = lp.lat_in; p.init_broadcast = lp.init_broadcast;
    p.flags_in = lp.flags_in; p.neff_in = lp.neff_in; p.tops0_in = lp.tops0_in;
    p.lat_out = lp.lat_out; p.flags_out = lp.flags_out; p.neff_out = lp.neff_out; p.tops0_out = lp.tops0_out;
    p.SEQ = lp.SEQ; p.TOPS = lp.TOPS; p.tops_burn = lp.tops_burn; p.use_conv = lp.use_conv; p.eps = lp.eps;
    p.hist = hist; p.hist_stride = hist_stride;
    p.eq_counts = lp.eq_counts; p.info = lp.info; p.percent = lp.percent;
    memcpy(p.cls_delta, lp.cls_delta, sizeof(p.cls_delta));
    p.counters = lp.counters;
    QTRY(c->ld.queue.ensure(sizeof(unsigned int)));
    const unsigned int q0 = (unsigned int)grid * (unsigned int)pl.NLC;
    CUDA_OK(cudaMemcpyAsync(c->ld.queue.p, &q0, sizeof(q0), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));   // q0 is a stack variable
    p.queue = (unsigned int *)c->ld.queue.p;
    QECMC_PT_DISPATCH(launch_gw, c, p, pl, grid);
}

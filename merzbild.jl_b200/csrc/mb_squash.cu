// squash_pia!(pv, pia, species) (particles.jl:622-682): close the holes that merging / deletions leave in the logical
// index space.  The reference slides `index`/`cell` entries left, walking group 1 of all cells in cell order and then
// group 2 of all cells; on the device the index indirection is the identity, so the particle payload itself moves.
// New starts come from one exclusive scan over [n_group1(1..nc), n_group2(1..nc)]; only segments whose start changes are
// copied (through the ping-pong buffer, so parallel left shifts never overwrite unread sources).
#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

static __global__ void k_squash_counts(const Indexer* __restrict__ ix, int64_t nc, int32_t* __restrict__ cnt) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[c];
        cnt[c] = (int32_t)q.n_group1;
        cnt[nc + c] = (int32_t)q.n_group2;
    }
}

// phase 0: stage moved segments into `alt` (and the cell ids into `cell_stage`) at their new positions
// phase 1: copy them back into `cur` and rewrite the indexer
static __global__ void __launch_bounds__(256) k_squash_move(SoA cur, SoA alt, int32_t* __restrict__ cell, int32_t* __restrict__ cell_stage,
                                                            Indexer* __restrict__ ix, int64_t nc, const int64_t* __restrict__ newlo, int phase,
                                                            int* flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sgm = warp0; sgm < 2 * nc; sgm += nwarps) {
        const bool g2 = sgm >= nc;
        const int64_t c = g2 ? sgm - nc : sgm;
        const Indexer q = ix[c];
        const int64_t n = g2 ? q.n_group2 : q.n_group1;
        if (n <= 0) continue;
        const int64_t olo = (g2 ? q.start2 : q.start1) - 1;
        const int64_t nlo = newlo[sgm];
        if (olo == nlo) continue;
        if (olo < nlo) {  // the reference only ever shifts left (particles.jl:641,659,672: `if offset > 0`)
            if (lane == 0) atomicOr(&flags[0], DEVERR_PRECONDITION);
            continue;
        }
        if (phase == 0) {
            for (int64_t j = lane; j < n; j += 32) {
#pragma unroll
                for (int f = 0; f < 7; f++) alt.a[f][nlo + j] = cur.a[f][olo + j];
                cell_stage[nlo + j] = cell[olo + j];
            }
        } else {
            for (int64_t j = lane; j < n; j += 32) {
#pragma unroll
                for (int f = 0; f < 7; f++) cur.a[f][nlo + j] = alt.a[f][nlo + j];
                cell[nlo + j] = cell_stage[nlo + j];
            }
        }
    }
}
static __global__ void k_squash_fix_indexer(Indexer* __restrict__ ix, int64_t nc, const int64_t* __restrict__ newlo) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        Indexer q = ix[c];
        if (q.n_group1 > 0) { q.start1 = newlo[c] + 1; q.end1 = newlo[c] + q.n_group1; }
        if (q.n_group2 > 0) { q.start2 = newlo[nc + c] + 1; q.end2 = newlo[nc + c] + q.n_group2; }
        ix[c] = q;
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_squash_pia(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, int64_t species) {
    MB_ARG(ctx && pv && pia && species >= 1 && species <= pia->n_species, "squash_pia");
    const int s = (int)species - 1;
    if (pia->contiguous[s]) return MB_OK;  // particles.jl:623-625
    MB_CUDA(cudaSetDevice(ctx->device));
    int r = pv_ensure_alt(pv);
    if (r) return r;
    const int64_t nc = pia->n_cells;
    Indexer* ix = pia->d_indexer + (int64_t)s * nc;
    int32_t* cell_stage = (int32_t*)ctx_scratch(ctx, 0, (size_t)pv->cap * 4);
    int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)(2 * nc) * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)(2 * nc + 1) + gs_partial_count(2 * nc)) * 8);
    if (!cell_stage || !cnt || !p64) return MB_ERR_CUDA;
    int64_t* newlo = p64;
    int64_t* partial = p64 + (2 * nc + 1);
    cudaStream_t st = ctx->stream;
    ProfScope ps(ctx, PROF_SQUASH);
    ctx->state_gen++;
    k_squash_counts<<<grid_for(nc, 256), 256, 0, st>>>(ix, nc, cnt);
    MB_LAUNCH_CHECK(ctx);
    r = device_exclusive_scan(ctx, cnt, 2 * nc, newlo, partial);
    if (r) return r;
    const int g = grid_for(2 * nc * 32, 256, 8);
    k_squash_move<<<g, 256, 0, st>>>(pv->cur, pv->alt, pv->cell, cell_stage, ix, nc, newlo, 0, ctx->d_flags);
    MB_LAUNCH_CHECK(ctx);
    k_squash_move<<<g, 256, 0, st>>>(pv->cur, pv->alt, pv->cell, cell_stage, ix, nc, newlo, 1, ctx->d_flags);
    MB_LAUNCH_CHECK(ctx);
    k_squash_fix_indexer<<<grid_for(nc, 256), 256, 0, st>>>(ix, nc, newlo);
    MB_LAUNCH_CHECK(ctx);
    pia->contiguous[s] = 1;
    pia->contig_pending[s] = 0;
    return MB_OK;
}

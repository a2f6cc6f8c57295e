// squash_pia!(pv, pia, species) (particles.jl:622-682): close the holes that merging / deletions leave in the logical
// index space.  The reference slides `index`/`cell` entries left, walking group 1 of all cells in cell order and then
// group 2 of all cells; on the device the index indirection is the identity, so the particle payload itself moves.
// New starts come from one exclusive scan over [n_group1(1..nc), n_group2(1..nc)]; the live segments are copied once into the
// sort's ping-pong buffer at their new positions and the buffers are swapped.
#include "mb_common.cuh"
#include "mb_scan.cuh"
#include "mb_segcopy.cuh"

namespace mb {

static __global__ void k_squash_counts(const Indexer* __restrict__ ix, int64_t nc, int32_t* __restrict__ cnt) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[c];
        cnt[c] = (int32_t)q.n_group1;
        cnt[nc + c] = (int32_t)q.n_group2;
    }
}

// every live particle is copied once into the sort's ping-pong buffer at its new position (56 B read + 56 B written per particle,
// moved or not); the buffers are then swapped on the host side, so no second payload pass is needed.  The cell ids go through a
// staging array and are copied back by k_squash_cell_back.  Load balance over segment sizes: mb_segcopy.cuh.
struct SquashMoveAct {
    SoA cur, alt;
    const int32_t* cell;
    int32_t* cell_stage;
    __device__ __forceinline__ void seg(int64_t, int64_t, int64_t, int64_t) const {}
    __device__ __forceinline__ void elem(int64_t, int64_t src, int64_t dst) const {
#pragma unroll
        for (int f = 0; f < 7; f++) alt.a[f][dst] = cur.a[f][src];
        cell_stage[dst] = cell[src];
    }
};
static __global__ void k_squash_cell_back(int32_t* __restrict__ cell, const int32_t* __restrict__ cell_stage, const int64_t* __restrict__ n_total_p) {
    const int64_t n = *n_total_p;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) cell[i] = cell_stage[i];
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
    if (pv->n_arrivals > 0) {
        mb::set_error("squash_pia!: arrivals of a slab exchange are pending (they are in no indexer yet): sort_particles! first");
        return MB_ERR_PRECONDITION;
    }
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
    {
        SquashDesc D{ix, nc, newlo, ctx->d_flags};
        SquashMoveAct A{pv->cur, pv->alt, pv->cell, cell_stage};
        r = seg_copy(ctx, 7, pv->cap, 2 * nc, D, A);
        if (r) return r;
    }
    k_squash_cell_back<<<grid_for(pia->n_bound[s] > 0 ? pia->n_bound[s] : pv->cap, 256, 8), 256, 0, st>>>(pv->cell, cell_stage, pia->d_n_total + s);
    MB_LAUNCH_CHECK(ctx);
    {   // ping-pong: the squashed particles live in the other buffer now
        SoA t = pv->cur;
        pv->cur = pv->alt;
        pv->alt = t;
    }
    k_squash_fix_indexer<<<grid_for(nc, 256), 256, 0, st>>>(ix, nc, newlo);
    MB_LAUNCH_CHECK(ctx);
    pia->contiguous[s] = 1;
    pia->contig_pending[s] = 0;
    return MB_OK;
}

// squash_pia!(pv, pia, species) (particles.jl:622-682): close the holes that merging / deletions leave in the logical
// index space.  The reference slides `index`/`cell` entries left, walking group 1 of all cells in cell order and then
// group 2 of all cells; on the device the index indirection is the identity, so the particle payload itself moves.
// New starts come from one exclusive scan over [n_group1(1..nc), n_group2(1..nc)]; the live segments are copied once into the
// sort's ping-pong buffer at their new positions and the buffers are swapped.
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

// every live particle is copied once into the sort's ping-pong buffer at its new position (56 B read + 56 B written per particle,
// moved or not); the buffers are then swapped on the host side, so no second payload pass is needed.  The cell ids go through a
// staging array and are copied back by k_squash_cell_back.
// Load balance: the OUTPUT range [0, n_total) is cut into tiles of SQ_TILE positions, one warp per tile; the warp finds the segment
// that holds the tile's first position by bisection over the new starts and walks on from there, so a 0-D cell of 1e4 particles and a
// Couette cell of 1e2 cost the same per particle.
constexpr int SQ_TILE = 2048;
static __global__ void __launch_bounds__(256) k_squash_move(SoA cur, SoA alt, const int32_t* __restrict__ cell, int32_t* __restrict__ cell_stage,
                                                            const Indexer* __restrict__ ix, int64_t nc, const int64_t* __restrict__ newlo, int* flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nseg = 2 * nc, total = newlo[nseg];
    for (int64_t d0 = warp0 * SQ_TILE; d0 < total; d0 += nwarps * SQ_TILE) {
        const int64_t d1 = d0 + SQ_TILE < total ? d0 + SQ_TILE : total;
        // last segment whose new start is <= d0 (empty segments share their successor's start, so this one is not empty)
        int64_t lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (newlo[mid] <= d0) lo = mid; else hi = mid - 1;
        }
        int64_t sgm = lo, d = d0;
        while (d < d1) {
            const int64_t s_lo = newlo[sgm], s_hi = newlo[sgm + 1];
            if (s_hi > d) {
                const bool g2 = sgm >= nc;
                const Indexer q = ix[g2 ? sgm - nc : sgm];
                const int64_t olo = (g2 ? q.start2 : q.start1) - 1;
                if (olo < s_lo) {  // the reference only ever shifts left (particles.jl:641,659,672: `if offset > 0`)
                    if (lane == 0) atomicOr(&flags[0], DEVERR_PRECONDITION);
                } else {
                    const int64_t e = s_hi < d1 ? s_hi : d1;
                    const int64_t src0 = olo + (d - s_lo);
                    for (int64_t j = lane; j < e - d; j += 32) {
#pragma unroll
                        for (int f = 0; f < 7; f++) alt.a[f][d + j] = cur.a[f][src0 + j];
                        cell_stage[d + j] = cell[src0 + j];
                    }
                }
                d = s_hi < d1 ? s_hi : d1;
            }
            sgm++;
        }
    }
}
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
    const int64_t nb = pia->n_bound[s] > 0 ? pia->n_bound[s] : pv->cap;
    const int g = grid_for((nb + SQ_TILE - 1) / SQ_TILE * 32, 256, 8);
    k_squash_move<<<g, 256, 0, st>>>(pv->cur, pv->alt, pv->cell, cell_stage, ix, nc, newlo, ctx->d_flags);
    MB_LAUNCH_CHECK(ctx);
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

// Shared pieces of the operators that append particles to a cell (variable-weight NTC splits, SWPM children):
// picked-particle loads, the group-2 append (particles.jl:426-433) into a per-cell window at the tail, and the packing
// of the windows so that the final layout equals the reference's sequential appends at n_total + 1.
#pragma once
#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

struct PRef {  // a particle picked for a collision: physical (0-based) position in its SoA
    int64_t pos;
    double w, vx, vy, vz;
};
__device__ __forceinline__ void load_p(const SoA& s, int64_t pos, PRef& p) {
    p.pos = pos;
    p.w = s.a[F_W][pos]; p.vx = s.a[F_VX][pos]; p.vy = s.a[F_VY][pos]; p.vz = s.a[F_VZ][pos];
}
__device__ __forceinline__ int64_t map_cont(const Indexer& q, int64_t i) {  // particles.jl:364-366, returned 0-based
    return (i < q.n_group1 ? i + q.start1 : (i - q.n_group1) + q.start2) - 1;
}
// split: append (dw, v, x of the parent) as a new group-2 particle (collision_ntc.jl:238-267, particles.jl:426-433)
__device__ __forceinline__ void append_split(const SoA& s, Indexer& q, int64_t winlo, int64_t parent, double dw, double vx, double vy, double vz) {
    const int64_t pos = q.n_group2 > 0 ? q.end2 : winlo;  // 0-based position of the new particle (end2 is 1-based -> next slot)
    if (q.n_group2 == 0) q.start2 = winlo + 1;
    q.n_group2 += 1;
    q.n_local += 1;
    q.end2 = pos + 1;
    s.a[F_W][pos] = dw;
    s.a[F_VX][pos] = vx; s.a[F_VY][pos] = vy; s.a[F_VZ][pos] = vz;
    s.a[F_X][pos] = s.a[F_X][parent]; s.a[F_Y][pos] = s.a[F_Y][parent]; s.a[F_Z][pos] = s.a[F_Z][parent];
}

// pack the per-cell windows to the left (cell order) so the layout equals the reference's sequential appends
// Load balance as in k_squash_move: one warp per tile of PK_TILE packed output positions, bisection over the packed offsets.
constexpr int PK_TILE = 1024;
static __global__ void __launch_bounds__(256) k_ntc_pack(SoA cur, SoA alt, Indexer* __restrict__ ix, int64_t cell_lo, int64_t nr,
                                                         const int64_t* __restrict__ win, const int64_t* __restrict__ packed,
                                                         const int32_t* __restrict__ nsplit, const int64_t* n_total, int phase,
                                                         int32_t* __restrict__ cell_id) {
    const int64_t nt = *n_total;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = packed[nr];
    for (int64_t d0 = warp0 * PK_TILE; d0 < total; d0 += nwarps * PK_TILE) {
        const int64_t d1 = d0 + PK_TILE < total ? d0 + PK_TILE : total;
        int64_t lo = 0, hi = nr - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (packed[mid] <= d0) lo = mid; else hi = mid - 1;
        }
        int64_t r = lo, d = d0;
        while (d < d1) {
            const int64_t s_lo = packed[r], s_hi = packed[r + 1];
            if (s_hi > d) {
                const int64_t e = s_hi < d1 ? s_hi : d1;
                const bool moved = win[r] != s_lo;
                const int64_t src0 = nt + win[r] + (d - s_lo), dst0 = nt + d;
                if (phase == 0) {
                    if (moved)
                        for (int64_t j = lane; j < e - d; j += 32)
#pragma unroll
                            for (int f = 0; f < 7; f++) alt.a[f][dst0 + j] = cur.a[f][src0 + j];
                } else {
                    for (int64_t j = lane; j < e - d; j += 32) {
                        if (moved)
#pragma unroll
                            for (int f = 0; f < 7; f++) cur.a[f][dst0 + j] = alt.a[f][dst0 + j];
                        // device-side extension: the new particles carry their cell id, so that an ensemble of 0-D cells can be re-sorted
                        // by sort_particles!(gridsort, particles, pia, species) (grid_sorting.jl:128) without a grid
                        cell_id[dst0 + j] = (int32_t)(cell_lo + r);
                    }
                    if (lane == 0 && moved && d == s_lo) {  // the warp that holds the window's first particle rewrites its range
                        Indexer q = ix[cell_lo - 1 + r];
                        q.start2 = nt + s_lo + 1;
                        q.end2 = nt + s_lo + (s_hi - s_lo);
                        ix[cell_lo - 1 + r] = q;
                    }
                }
                d = e;
            }
            r++;
        }
    }
}
static __global__ void k_add_total(int64_t* n_total, const int64_t* packed, int64_t nr) {
    *n_total += packed[nr];
}


// scan of the actual append counts -> packed offsets; move the windows; n_total += appended
static inline int pack_windows(mb_ctx* ctx, mb_pv* pv, Indexer* ix, int64_t cell_lo, int64_t nr, const int64_t* win, const int32_t* nsplit,
                               int64_t* packed, int64_t* partial, int64_t* n_total) {
    int r = device_exclusive_scan(ctx, nsplit, nr, packed, partial);
    if (r) return r;
    const int64_t nb = pv->cap;  // upper bound on the appended particles; the kernels read the exact total from the scan
    const int gw = grid_for((nb + PK_TILE - 1) / PK_TILE * 32, 256, 4);
    if (nr > 1) {
        k_ntc_pack<<<gw, 256, 0, ctx->stream>>>(pv->cur, pv->alt, ix, cell_lo, nr, win, packed, nsplit, n_total, 0, pv->cell);
        MB_LAUNCH_CHECK(ctx);
    }
    k_ntc_pack<<<gw, 256, 0, ctx->stream>>>(pv->cur, pv->alt, ix, cell_lo, nr, win, packed, nsplit, n_total, 1, pv->cell);
    MB_LAUNCH_CHECK(ctx);
    k_add_total<<<1, 1, 0, ctx->stream>>>(n_total, packed, nr);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

}  // namespace mb

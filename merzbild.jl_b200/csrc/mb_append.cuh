// Shared pieces of the operators that append particles to a cell (variable-weight NTC splits, SWPM children):
// picked-particle loads, the group-2 append (particles.jl:426-433) into a per-cell window at the tail, and the packing
// of the windows so that the final layout equals the reference's sequential appends at n_total + 1.
#pragma once
#include <cstdint>
#include "mb_common.cuh"
#include "mb_scan.cuh"
#include "mb_segcopy.cuh"

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
// winhi (exclusive): end of the cell's window; a split that does not fit raises DEVERR_CAPACITY and is NOT performed (returns false)
__device__ __forceinline__ bool append_split(const SoA& s, Indexer& q, int64_t winlo, int64_t parent, double dw, double vx, double vy, double vz,
                                             int64_t winhi = INT64_MAX, int* flags = nullptr) {
    const int64_t pos = q.n_group2 > 0 ? q.end2 : winlo;  // 0-based position of the new particle (end2 is 1-based -> next slot)
    if (pos >= winhi) {
        if (flags) atomicOr(&flags[0], DEVERR_CAPACITY);
        return false;
    }
    if (q.n_group2 == 0) q.start2 = winlo + 1;
    q.n_group2 += 1;
    q.n_local += 1;
    q.end2 = pos + 1;
    s.a[F_W][pos] = dw;
    s.a[F_VX][pos] = vx; s.a[F_VY][pos] = vy; s.a[F_VZ][pos] = vz;
    s.a[F_X][pos] = s.a[F_X][parent]; s.a[F_Y][pos] = s.a[F_Y][parent]; s.a[F_Z][pos] = s.a[F_Z][parent];
    return true;
}

// pack the per-cell windows to the left (cell order) so the layout equals the reference's sequential appends
// the split windows: window r holds nsplit[r] particles at n_total + win[r] and moves to n_total + packed[r]
struct PackDesc {
    const int64_t* win;
    const int64_t* packed;
    const int32_t* nsplit;
    const int64_t* n_total;
    __device__ __forceinline__ void get(int64_t r, int64_t& n, int64_t& src, int64_t& dst) const {
        const int64_t nt = *n_total;
        n = nsplit[r];
        src = nt + win[r];
        dst = nt + packed[r];
    }
};
// phase 0: moved windows -> alt (in place would overwrite unread sources); phase 1: back into cur at the packed position, the
// window's group-2 range rewritten, and -- device-side extension -- the cell id of every new particle, so that an ensemble of 0-D
// cells can be re-sorted by sort_particles!(gridsort, particles, pia, species) (grid_sorting.jl:128) without a grid
struct PackAct {
    SoA cur, alt;
    Indexer* ix;
    int64_t cell_lo;
    int32_t* cell_id;
    int phase;
    __device__ __forceinline__ void seg(int64_t r, int64_t n, int64_t src, int64_t dst) const {
        if (phase == 1 && src != dst) {
            Indexer q = ix[cell_lo - 1 + r];
            q.start2 = dst + 1;
            q.end2 = dst + n;
            ix[cell_lo - 1 + r] = q;
        }
    }
    __device__ __forceinline__ void elem(int64_t r, int64_t src, int64_t dst) const {
        if (phase == 0) {
            if (src != dst)
#pragma unroll
                for (int f = 0; f < 7; f++) alt.a[f][dst] = cur.a[f][src];
        } else {
            if (src != dst)
#pragma unroll
                for (int f = 0; f < 7; f++) cur.a[f][dst] = alt.a[f][dst];
            cell_id[dst] = (int32_t)(cell_lo + r);
        }
    }
};
static __global__ void k_add_total(int64_t* n_total, const int64_t* packed, int64_t nr) {
    *n_total += packed[nr];
}


// scan of the actual append counts -> packed offsets; move the windows; n_total += appended
static inline int pack_windows(mb_ctx* ctx, mb_pv* pv, Indexer* ix, int64_t cell_lo, int64_t nr, const int64_t* win, const int32_t* nsplit,
                               int64_t* packed, int64_t* partial, int64_t* n_total) {
    int r = device_exclusive_scan(ctx, nsplit, nr, packed, partial);
    if (r) return r;
    PackDesc D{win, packed, nsplit, n_total};
    if (nr > 1) {
        PackAct A0{pv->cur, pv->alt, ix, cell_lo, pv->cell, 0};
        r = seg_copy(ctx, 7, pv->cap, nr, D, A0);
        if (r) return r;
    }
    PackAct A1{pv->cur, pv->alt, ix, cell_lo, pv->cell, 1};
    r = seg_copy(ctx, 7, pv->cap, nr, D, A1);
    if (r) return r;
    k_add_total<<<1, 1, 0, ctx->stream>>>(n_total, packed, nr);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

}  // namespace mb

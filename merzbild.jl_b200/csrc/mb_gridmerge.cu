// merge_grid_based! (merging/merging_grid.jl:597-703): N:2 merging on a Cartesian grid in velocity space -- compute_velocity_extent!
// (:190-221), compute_grid_index (:235-264), compute_grid! (:304-375: per grid cell np, w, weighted mean / variance of v and x, the
// first two particle indices), compute_new_particles! (:393-468; 1-D variant :484-575 clamps x of every np >= 2 output),
// delete_particle_end! bookkeeping (particles.jl:478-566).  SURVEY.md 8(f)4: same N:2 back-end as the octree merge.
//
// One CTA per merging cell, all cells of the range concurrently (cells are scanned in chunks, see k_merge).  The grid has
// Ntotal = Nx Ny Nz + 8 velocity cells ("bins"; the last 8 are the octants outside the grid).  Per physical cell:
//   1. bin index of every particle (logical order: group 1, then group 2) and the bin populations (shared-memory counters),
//   2. exclusive scan -> bin slices; the particle indices are dropped into their bin's slice with a shared cursor and every slice
//      is then put into ascending particle order by its thread (bins hold a handful of particles), i.e. the reference's visiting order,
//   3. ONE THREAD PER BIN walks its slice twice (mean, variance) in that order with the reference's operations (sum of v * w, division
//      by w), so the bin moments are bit-identical to the sequential loops of compute_grid!,
//   4. post-merge particles (2 per bin of np > 2, sign draws = Philox block (bin index - 1) of the (OP_MERGE_GRID, timestep, cell)
//      stream; np <= 2 bins keep their particles) staged in a per-CTA buffer, then written to the cell's first logical slots in bin
//      order; the remaining slots are deleted from the end (group 2 first).
#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

constexpr int GM_T = 256;         // threads per CTA
constexpr int GM_MAXBINS = 8192;  // Nx Ny Nz + 8 must fit the shared-memory counters

struct GridMergeArgs {
    SoA pv;
    Indexer* ix;
    int64_t* n_total;
    int64_t cell_lo, cell_hi, n_cells_total;
    int64_t threshold;
    int Nx, Ny, Nz, Ntotal;
    double mult[3];
    double mass;
    const double* T;     // props variant: T[cell], v[3 * cell + d] of the species (device); nullptr -> explicit extents
    const double* v;
    double ext[6];
    int has_grid;
    double min_x, max_x;
    uint64_t seed;
    uint32_t timestep, substream;
    int32_t* idx;        // [cap] per-cell slices at slice[r]
    int32_t* bin_of;     // [cap]
    const int64_t* slice;
    double* outbuf;      // [nCTA][2 * Ntotal][7]
    int* flags;
    int* noncontig;
    int cpb;
    int small_max;  // cells with n_local <= small_max are merged by k_merge_grid_warp
};

static __global__ void k_gm_counts(const Indexer* __restrict__ ix, int64_t cell_lo, int64_t nr, int64_t threshold, int32_t* __restrict__ cnt) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = ix[cell_lo - 1 + r].n_local;
        cnt[r] = (n > 0 && (threshold < 0 || n > threshold)) ? (int32_t)n : 0;
    }
}

__device__ __forceinline__ int64_t gm_pos(const Indexer& q, int64_t j) {  // map_cont_index, 0-based physical position
    return (j < q.n_group1 ? j + q.start1 : (j - q.n_group1) + q.start2) - 1;
}

// exclusive scan of in[0..n) (shared memory) into out_a / out_b (may alias each other or be null), total returned to every thread;
// warp shuffles + one shared hop: 3 barriers per 1024-entry... per blockDim-sized chunk
__device__ __forceinline__ int gm_block_scan(const int32_t* in, int n, int32_t* out_a, int32_t* out_b, int* s_wsum, int* s_carry) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    __syncthreads();
    if (tid == 0) *s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += nt) {
        const int b = c0 + tid;
        const int v = b < n ? in[b] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        int woff = 0;
        for (int i = 0; i < wid; i++) woff += s_wsum[i];
        int tot = 0;
        for (int i = 0; i < nw; i++) tot += s_wsum[i];
        const int carry = *s_carry;
        const int excl = carry + woff + incl - v;
        __syncthreads();
        if (b < n) {
            if (out_a) out_a[b] = excl;
            if (out_b) out_b[b] = excl;
        }
        if (tid == 0) *s_carry = carry + tot;
        __syncthreads();
    }
    return *s_carry;
}

__global__ void __launch_bounds__(GM_T) k_merge_grid(GridMergeArgs a) {
    extern __shared__ __align__(16) unsigned char gm_dyn[];
    int32_t* s_cnt = (int32_t*)gm_dyn;            // [Ntotal] population, later output count
    int32_t* s_start = s_cnt + a.Ntotal;          // [Ntotal + 1] slice starts
    int32_t* s_cur = s_start + a.Ntotal + 1;      // [Ntotal] scatter cursor, later output offset
    __shared__ int s_list[GM_T], s_nlist, s_wsum[GM_T / 32], s_carry, s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const int NB = a.Ntotal;
    double* outbuf = a.outbuf + (int64_t)blockIdx.x * 2 * NB * 7;
    for (int64_t rbase = (int64_t)blockIdx.x * a.cpb; rbase < nr; rbase += (int64_t)gridDim.x * a.cpb) {
      __syncthreads();
      if (tid == 0) s_nlist = 0;
      __syncthreads();
      if (tid < a.cpb && rbase + tid < nr) {
          const int64_t n_l = a.ix[a.cell_lo - 1 + rbase + tid].n_local;
          if (n_l > 0 && (a.threshold < 0 || n_l > a.threshold) && n_l > a.small_max) s_list[atomicAdd(&s_nlist, 1)] = tid;
      }
      __syncthreads();
      const int nlist = s_nlist;
      for (int li = 0; li < nlist; li++) {
        const int64_t r = rbase + s_list[li];
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int N = (int)q.n_local;
        int32_t* idx = a.idx + a.slice[r];
        int32_t* bin_of = a.bin_of + a.slice[r];
        // ---- compute_velocity_extent! (:190-221)
        double lo[3], hi[3], mid[3], dvi[3];
        {
            const double Nd[3] = {(double)a.Nx, (double)a.Ny, (double)a.Nz};
            for (int d = 0; d < 3; d++) {
                double dv_cell;
                if (a.T != nullptr) {
                    const double dv = a.mult[d] * sqrt(2 * a.T[cell - 1] * k_B / a.mass);
                    const double v0 = a.v[3 * (cell - 1) + d];
                    lo[d] = v0 - dv; hi[d] = v0 + dv; mid[d] = v0;
                    dv_cell = 2 * dv / Nd[d];
                } else {
                    lo[d] = a.ext[2 * d]; hi[d] = a.ext[2 * d + 1];
                    mid[d] = 0.5 * (a.ext[2 * d] + a.ext[2 * d + 1]);
                    dv_cell = (a.ext[2 * d + 1] - a.ext[2 * d]) / Nd[d];
                }
                dvi[d] = 1.0 / dv_cell;
            }
        }
        __syncthreads();
        for (int b = tid; b < NB; b += nt) s_cnt[b] = 0;
        if (tid == 0) s_bad = 0;
        __syncthreads();
        // ---- compute_grid_index (:235-264) of every particle, bin populations
        for (int j = tid; j < N; j += nt) {
            const int64_t p = gm_pos(q, j);
            const double vx = a.pv.a[F_VX][p], vy = a.pv.a[F_VY][p], vz = a.pv.a[F_VZ][p];
            bool outside = false;
            if (vx < lo[0] || vx > hi[0]) outside = true;
            else if (vy < lo[1] || vy > hi[1]) outside = true;
            else if (vz < lo[2] || vz > hi[2]) outside = true;
            long long index;
            if (!outside) {
                index = (long long)floor((vx - lo[0]) * dvi[0]) * (a.Ny * a.Nz) + (long long)floor((vy - lo[1]) * dvi[1]) * a.Nz +
                        (long long)floor((vz - lo[2]) * dvi[2]);
            } else {
                index = NB - 8 + (vx > mid[0] ? 1 : 0) + (vy > mid[1] ? 2 : 0) + (vz > mid[2] ? 4 : 0);
            }
            if (index < 0 || index >= NB) {  // v exactly on the upper bound: the reference would index past the grid
                s_bad = 1;
                index = 0;
            }
            bin_of[j] = (int32_t)index;
            atomicAdd(&s_cnt[index], 1);
        }
        __syncthreads();
        if (s_bad) {  // block-uniform
            if (tid == 0) atomicOr(&a.flags[0], DEVERR_PRECONDITION);
            continue;
        }
        // ---- exclusive scan of the populations -> slices
        {
            const int tot = gm_block_scan(s_cnt, NB, s_start, s_cur, s_wsum, &s_carry);
            if (tid == 0) s_start[NB] = tot;
        }
        __syncthreads();
        for (int j = tid; j < N; j += nt) idx[atomicAdd(&s_cur[bin_of[j]], 1)] = j;  // local (logical) particle number
        __syncthreads();
        // ---- per bin: ascending particle order (insertion sort of a short slice), moments, post-merge particles
        const uint32_t c3 = (OP_MERGE_GRID & 0xFFu) | (a.substream << 8);
        for (int b = tid; b < NB; b += nt) {
            const int bs = s_start[b], be = s_start[b + 1];
            int np = be - bs;
            for (int i = bs + 1; i < be; i++) {
                const int32_t key = idx[i];
                int k = i - 1;
                while (k >= bs && idx[k] > key) { idx[k + 1] = idx[k]; k--; }
                idx[k + 1] = key;
            }
            double w = 0, vm[3] = {0, 0, 0}, xm[3] = {0, 0, 0};
            for (int i = bs; i < be; i++) {
                const int64_t p = gm_pos(q, idx[i]);
                const double pw = a.pv.a[F_W][p];
                w += pw;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    vm[d] = vm[d] + a.pv.a[F_VX + d][p] * pw;
                    xm[d] = xm[d] + a.pv.a[F_X + d][p] * pw;
                }
            }
            if (!(w > 0.0)) np = 0;  // :345-351
            double* o1 = outbuf + (int64_t)(2 * b) * 7;
            double* o2 = o1 + 7;
            if (np > 2) {
                double vs[3] = {0, 0, 0}, xs[3] = {0, 0, 0};
#pragma unroll
                for (int d = 0; d < 3; d++) { vm[d] = vm[d] / w; xm[d] = xm[d] / w; }
                for (int i = bs; i < be; i++) {
                    const int64_t p = gm_pos(q, idx[i]);
                    const double pw = a.pv.a[F_W][p];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const double dvv = a.pv.a[F_VX + d][p] - vm[d], dxx = a.pv.a[F_X + d][p] - xm[d];
                        vs[d] = vs[d] + (dvv * dvv) * pw;
                        xs[d] = xs[d] + (dxx * dxx) * pw;
                    }
                }
                uint32_t rb[4];
                philox4x32_10((uint32_t)b, (uint32_t)cell, a.timestep, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rb);
                o1[0] = 0.5 * w; o2[0] = 0.5 * w;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double sdv = sqrt(vs[d] / w), sdx = sqrt(xs[d] / w);
                    const double sgv = ((rb[0] >> d) & 1u) ? 1.0 : -1.0, sgx = ((rb[0] >> (3 + d)) & 1u) ? 1.0 : -1.0;
                    o1[1 + d] = vm[d] + sgv * sdv;
                    o2[1 + d] = vm[d] - sgv * sdv;
                    o1[4 + d] = xm[d] + sgx * sdx;
                    o2[4 + d] = xm[d] - sgx * sdx;
                }
            } else if (np >= 1) {
                const int64_t p1 = gm_pos(q, idx[bs]);
#pragma unroll
                for (int f = 0; f < 7; f++) o1[f] = a.pv.a[f][p1];
                if (np == 2) {
                    const int64_t p2 = gm_pos(q, idx[bs + 1]);
#pragma unroll
                    for (int f = 0; f < 7; f++) o2[f] = a.pv.a[f][p2];
                }
            }
            if (np >= 2 && a.has_grid) {  // 1-D variant: every np >= 2 output is clamped (:528-549)
                o1[4] = o1[4] < a.min_x ? a.min_x : (o1[4] > a.max_x ? a.max_x : o1[4]);
                o2[4] = o2[4] < a.min_x ? a.min_x : (o2[4] > a.max_x ? a.max_x : o2[4]);
            }
            s_cnt[b] = np >= 2 ? 2 : np;
        }
        __syncthreads();
        // ---- output offsets in bin order
        const int curr = gm_block_scan(s_cnt, NB, s_cur, nullptr, s_wsum, &s_carry);
        __syncthreads();
        // ---- write the post-merge particles into the first `curr` logical slots (every input they depend on is staged in outbuf)
        for (int b = tid; b < NB; b += nt) {
            const int no = s_cnt[b], off = s_cur[b];
            for (int k = 0; k < no; k++) {
                const int64_t p = gm_pos(q, off + k);
#pragma unroll
                for (int f = 0; f < 7; f++) a.pv.a[f][p] = outbuf[(int64_t)(2 * b + k) * 7 + f];
            }
        }
        // ---- delete_particle_end! x n_delete: group 2 shrinks first, then group 1; deleted slots get w = 0
        const int n_del = N - curr;
        for (int j = curr + tid; j < N; j += nt) a.pv.a[F_W][gm_pos(q, j)] = 0.0;
        if (tid == 0) {
            Indexer u = q;
            int64_t d = n_del;
            const int64_t d2 = d < u.n_group2 ? d : u.n_group2;
            u.n_group2 -= d2; u.end2 -= d2;
            if (u.n_group2 == 0) { u.start2 = 0; u.end2 = -1; }
            d -= d2;
            u.n_group1 -= d; u.end1 -= d;
            if (u.n_group1 == 0) { u.start1 = 0; u.end1 = -1; }
            u.n_local = curr;
            a.ix[cell - 1] = u;
            if (n_del > 0) atomicAdd((unsigned long long*)a.n_total, (unsigned long long)(-(long long)n_del));
            if (!(cell == a.n_cells_total) || n_del > q.n_group2) *a.noncontig = 1;
        }
        __syncthreads();
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small cells (N <= GW_NS particles, <= GW_NB velocity cells: the 1-D usage): one WARP per merging cell, counters / slices / bin
// membership in the warp's slice of shared memory, __syncwarp only.  Same arithmetic as k_merge_grid (one lane per velocity cell
// walking its slice in ascending particle order), so both kernels and the oracle agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GW_NS = 256, GW_NB = 136, GW_WARPS = 8;

__global__ void __launch_bounds__(32 * GW_WARPS) k_merge_grid_warp(GridMergeArgs a, int ch, double* __restrict__ outbuf_all) {
    __shared__ int32_t sw_cnt[GW_WARPS][GW_NB], sw_start[GW_WARPS][GW_NB + 1], sw_cur[GW_WARPS][GW_NB];
    __shared__ int16_t sw_bin[GW_WARPS][GW_NS], sw_idx[GW_WARPS][GW_NS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    int32_t* s_cnt = sw_cnt[wid];
    int32_t* s_start = sw_start[wid];
    int32_t* s_cur = sw_cur[wid];
    int16_t* bin_of = sw_bin[wid];
    int16_t* idx = sw_idx[wid];
    const int NB = a.Ntotal;
    const int64_t gw = (int64_t)blockIdx.x * GW_WARPS + wid, nwarps = (int64_t)gridDim.x * GW_WARPS;
    double* outbuf = outbuf_all + gw * 2 * NB * 7;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const uint32_t c3 = (OP_MERGE_GRID & 0xFFu) | (a.substream << 8);
    for (int64_t r0 = gw * ch; r0 < nr; r0 += nwarps * ch) {
      const int64_t myr = r0 + lane;
      const int64_t my_n = (lane < ch && myr < nr) ? a.ix[a.cell_lo - 1 + myr].n_local : 0;
      unsigned todo = __ballot_sync(FULL, my_n > 0 && my_n <= GW_NS && (a.threshold < 0 || my_n > a.threshold));
      while (todo) {
        const int64_t r = r0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int N = (int)q.n_local;
        double lo[3], hi[3], mid[3], dvi[3];
        {
            const double Nd[3] = {(double)a.Nx, (double)a.Ny, (double)a.Nz};
            for (int d = 0; d < 3; d++) {
                double dv_cell;
                if (a.T != nullptr) {
                    const double dv = a.mult[d] * sqrt(2 * a.T[cell - 1] * k_B / a.mass);
                    const double v0 = a.v[3 * (cell - 1) + d];
                    lo[d] = v0 - dv; hi[d] = v0 + dv; mid[d] = v0;
                    dv_cell = 2 * dv / Nd[d];
                } else {
                    lo[d] = a.ext[2 * d]; hi[d] = a.ext[2 * d + 1];
                    mid[d] = 0.5 * (a.ext[2 * d] + a.ext[2 * d + 1]);
                    dv_cell = (a.ext[2 * d + 1] - a.ext[2 * d]) / Nd[d];
                }
                dvi[d] = 1.0 / dv_cell;
            }
        }
        __syncwarp();
        for (int b = lane; b < NB; b += 32) s_cnt[b] = 0;
        __syncwarp();
        bool bad = false;
        for (int j = lane; j < N; j += 32) {
            const int64_t p = gm_pos(q, j);
            const double vx = a.pv.a[F_VX][p], vy = a.pv.a[F_VY][p], vz = a.pv.a[F_VZ][p];
            bool outside = false;
            if (vx < lo[0] || vx > hi[0]) outside = true;
            else if (vy < lo[1] || vy > hi[1]) outside = true;
            else if (vz < lo[2] || vz > hi[2]) outside = true;
            long long index;
            if (!outside) {
                index = (long long)floor((vx - lo[0]) * dvi[0]) * (a.Ny * a.Nz) + (long long)floor((vy - lo[1]) * dvi[1]) * a.Nz +
                        (long long)floor((vz - lo[2]) * dvi[2]);
            } else {
                index = NB - 8 + (vx > mid[0] ? 1 : 0) + (vy > mid[1] ? 2 : 0) + (vz > mid[2] ? 4 : 0);
            }
            if (index < 0 || index >= NB) { bad = true; index = 0; }
            bin_of[j] = (int16_t)index;
            atomicAdd(&s_cnt[index], 1);
        }
        if (__any_sync(FULL, bad)) {
            if (lane == 0) atomicOr(&a.flags[0], DEVERR_PRECONDITION);
            continue;
        }
        __syncwarp();
        // exclusive scan of the populations
        int carry = 0;
        for (int c0 = 0; c0 < NB; c0 += 32) {
            const int b = c0 + lane;
            const int v = b < NB ? s_cnt[b] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (b < NB) { s_start[b] = carry + incl - v; s_cur[b] = carry + incl - v; }
            carry += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) s_start[NB] = carry;
        __syncwarp();
        for (int j = lane; j < N; j += 32) idx[atomicAdd(&s_cur[bin_of[j]], 1)] = (int16_t)j;
        __syncwarp();
        for (int b = lane; b < NB; b += 32) {
            const int bs = s_start[b], be = s_start[b + 1];
            int np = be - bs;
            for (int i = bs + 1; i < be; i++) {
                const int16_t key = idx[i];
                int k = i - 1;
                while (k >= bs && idx[k] > key) { idx[k + 1] = idx[k]; k--; }
                idx[k + 1] = key;
            }
            double w = 0, vm[3] = {0, 0, 0}, xm[3] = {0, 0, 0};
            for (int i = bs; i < be; i++) {
                const int64_t p = gm_pos(q, idx[i]);
                const double pw = a.pv.a[F_W][p];
                w += pw;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    vm[d] = vm[d] + a.pv.a[F_VX + d][p] * pw;
                    xm[d] = xm[d] + a.pv.a[F_X + d][p] * pw;
                }
            }
            if (!(w > 0.0)) np = 0;
            double* o1 = outbuf + (int64_t)(2 * b) * 7;
            double* o2 = o1 + 7;
            if (np > 2) {
                double vs[3] = {0, 0, 0}, xs[3] = {0, 0, 0};
#pragma unroll
                for (int d = 0; d < 3; d++) { vm[d] = vm[d] / w; xm[d] = xm[d] / w; }
                for (int i = bs; i < be; i++) {
                    const int64_t p = gm_pos(q, idx[i]);
                    const double pw = a.pv.a[F_W][p];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const double dvv = a.pv.a[F_VX + d][p] - vm[d], dxx = a.pv.a[F_X + d][p] - xm[d];
                        vs[d] = vs[d] + (dvv * dvv) * pw;
                        xs[d] = xs[d] + (dxx * dxx) * pw;
                    }
                }
                uint32_t rb[4];
                philox4x32_10((uint32_t)b, (uint32_t)cell, a.timestep, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rb);
                o1[0] = 0.5 * w; o2[0] = 0.5 * w;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double sdv = sqrt(vs[d] / w), sdx = sqrt(xs[d] / w);
                    const double sgv = ((rb[0] >> d) & 1u) ? 1.0 : -1.0, sgx = ((rb[0] >> (3 + d)) & 1u) ? 1.0 : -1.0;
                    o1[1 + d] = vm[d] + sgv * sdv;
                    o2[1 + d] = vm[d] - sgv * sdv;
                    o1[4 + d] = xm[d] + sgx * sdx;
                    o2[4 + d] = xm[d] - sgx * sdx;
                }
            } else if (np >= 1) {
                const int64_t p1 = gm_pos(q, idx[bs]);
#pragma unroll
                for (int f = 0; f < 7; f++) o1[f] = a.pv.a[f][p1];
                if (np == 2) {
                    const int64_t p2 = gm_pos(q, idx[bs + 1]);
#pragma unroll
                    for (int f = 0; f < 7; f++) o2[f] = a.pv.a[f][p2];
                }
            }
            if (np >= 2 && a.has_grid) {
                o1[4] = o1[4] < a.min_x ? a.min_x : (o1[4] > a.max_x ? a.max_x : o1[4]);
                o2[4] = o2[4] < a.min_x ? a.min_x : (o2[4] > a.max_x ? a.max_x : o2[4]);
            }
            s_cnt[b] = np >= 2 ? 2 : np;
        }
        __syncwarp();
        int curr = 0;
        for (int c0 = 0; c0 < NB; c0 += 32) {
            const int b = c0 + lane;
            const int v = b < NB ? s_cnt[b] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (b < NB) s_cur[b] = curr + incl - v;
            curr += __shfl_sync(FULL, incl, 31);
        }
        __syncwarp();  // every input has been read and staged in outbuf (global, same warp: ordered by the barrier)
        for (int b = lane; b < NB; b += 32) {
            const int no = s_cnt[b], off = s_cur[b];
            for (int k = 0; k < no; k++) {
                const int64_t p = gm_pos(q, off + k);
#pragma unroll
                for (int f = 0; f < 7; f++) a.pv.a[f][p] = outbuf[(int64_t)(2 * b + k) * 7 + f];
            }
        }
        const int n_del = N - curr;
        for (int j = curr + lane; j < N; j += 32) a.pv.a[F_W][gm_pos(q, j)] = 0.0;
        if (lane == 0) {
            Indexer u = q;
            int64_t d = n_del;
            const int64_t d2 = d < u.n_group2 ? d : u.n_group2;
            u.n_group2 -= d2; u.end2 -= d2;
            if (u.n_group2 == 0) { u.start2 = 0; u.end2 = -1; }
            d -= d2;
            u.n_group1 -= d; u.end1 -= d;
            if (u.n_group1 == 0) { u.start1 = 0; u.end1 = -1; }
            u.n_local = curr;
            a.ix[cell - 1] = u;
            if (n_del > 0) atomicAdd((unsigned long long*)a.n_total, (unsigned long long)(-(long long)n_del));
            if (!(cell == a.n_cells_total) || n_del > q.n_group2) *a.noncontig = 1;
        }
      }
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_merge_grid_based(mb_ctx* ctx, const mb_gridmerge_params* mg, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                                   double mass, mb_props* props, const double* extents6, int64_t threshold, const mb_grid1d* grid, uint32_t timestep,
                                   uint32_t substream) {
    MB_ARG(ctx && mg && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_ARG(mg->Nx >= 1 && mg->Ny >= 1 && mg->Nz >= 1, "GridN2Merge: Nx, Ny, Nz >= 1");
    MB_ARG((props != nullptr) != (extents6 != nullptr), "merge_grid_based!: pass either phys_props or the velocity extents");
    MB_ARG(props == nullptr || (props->n_cells == pia->n_cells && props->n_species == pia->n_species && mass > 0), "phys_props shape / mass");
    const int64_t NB = (int64_t)mg->Nx * mg->Ny * mg->Nz + 8;
    if (NB > GM_MAXBINS) {
        set_error("merge_grid_based!: Nx * Ny * Nz + 8 > 8192 velocity cells is not supported");
        return MB_ERR_UNSUPPORTED;
    }
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    const int64_t nc = pia->n_cells, nr = cell_hi - cell_lo + 1, cap = pv->cap;
    ProfScope ps(ctx, PROF_MERGE);
    ctx->state_gen++;
    cudaStream_t st = ctx->stream;
    GridMergeArgs a;
    a.pv = pv->cur;
    a.ix = pia->d_indexer + (int64_t)s * nc;
    a.n_total = pia->d_n_total + s;
    a.cell_lo = cell_lo; a.cell_hi = cell_hi; a.n_cells_total = nc;
    a.threshold = threshold;
    a.Nx = mg->Nx; a.Ny = mg->Ny; a.Nz = mg->Nz; a.Ntotal = (int)NB;
    for (int d = 0; d < 3; d++) a.mult[d] = mg->extent_multiplier[d];
    a.mass = mass;
    a.T = props ? props->T + (int64_t)s * nc : nullptr;
    a.v = props ? props->v + (int64_t)3 * s * nc : nullptr;
    for (int d = 0; d < 6; d++) a.ext[d] = extents6 ? extents6[d] : 0.0;
    a.has_grid = grid != nullptr;
    a.min_x = grid ? grid->min_x : 0.0;
    a.max_x = grid ? grid->max_x : 0.0;
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, species, species);
    a.flags = ctx->d_flags;
    a.idx = (int32_t*)ctx_scratch(ctx, 0, (size_t)cap * 4);
    a.bin_of = (int32_t*)ctx_scratch(ctx, 3, (size_t)cap * 4);
    int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)nr * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)(nr + 1) + gs_partial_count(nr)) * 8);
    if (!a.idx || !a.bin_of || !cnt || !p64) return MB_ERR_CUDA;
    a.slice = p64;
    k_gm_counts<<<grid_for(nr, 256), 256, 0, st>>>(a.ix, cell_lo, nr, threshold, cnt);
    MB_LAUNCH_CHECK(ctx);
    int r = device_exclusive_scan(ctx, cnt, nr, p64, p64 + (nr + 1));
    if (r) return r;
    int64_t nCTA = nr < (int64_t)N_SM * 4 ? nr : (int64_t)N_SM * 4;
    a.outbuf = (double*)ctx_scratch(ctx, 9, (size_t)nCTA * 2 * NB * 7 * 8 + 256);
    if (!a.outbuf) return MB_ERR_CUDA;
    a.noncontig = pia->d_holes + s;
    if (!pia->contig_pending[s]) MB_CUDA(cudaMemsetAsync(a.noncontig, 0, sizeof(int), st));
    // small cells: 128-thread CTAs, twice as many of them
    const int64_t avg = (pia->n_bound[s] > 0 ? pia->n_bound[s] : cap) / (nc > 0 ? nc : 1);
    const int threads = avg > 1024 ? GM_T : 128;
    if (threads == 128) {
        nCTA = nr < (int64_t)N_SM * 8 ? nr : (int64_t)N_SM * 8;
        a.outbuf = (double*)ctx_scratch(ctx, 9, (size_t)nCTA * 2 * NB * 7 * 8 + 256);
        if (!a.outbuf) return MB_ERR_CUDA;
    }
    a.cpb = threads;
    while (a.cpb > 1 && nr < 8 * nCTA * a.cpb) a.cpb >>= 1;
    const size_t smem = ((size_t)3 * NB + 2) * 4;
    static size_t attr_smem[64] = {0};  // function attributes are per device
    size_t& as = attr_smem[ctx->device & 63];
    if (smem > 40 * 1024 && smem > as) {
        MB_CUDA(cudaFuncSetAttribute(k_merge_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        as = smem;
    }
    a.small_max = 0;
    if (NB <= GW_NB) {  // small cells: one warp per cell
        a.small_max = GW_NS;
        int ch = 32;
        while (ch > 1 && nr < (int64_t)N_SM * 4 * GW_WARPS * ch) ch >>= 1;
        int64_t nW = (nr + (int64_t)ch * GW_WARPS - 1) / ((int64_t)ch * GW_WARPS);
        if (nW > (int64_t)N_SM * 4) nW = (int64_t)N_SM * 4;
        double* ob = (double*)ctx_scratch(ctx, 7, (size_t)nW * GW_WARPS * 2 * NB * 7 * 8 + 256);
        if (!ob) return MB_ERR_CUDA;
        k_merge_grid_warp<<<(int)nW, 32 * GW_WARPS, 0, st>>>(a, ch, ob);
        MB_LAUNCH_CHECK(ctx);
    }
    k_merge_grid<<<(int)nCTA, threads, smem, st>>>(a);
    MB_LAUNCH_CHECK(ctx);
    if (pia->contiguous[s]) pia->contig_pending[s] = 1;
    pia->contiguous[s] = 0;
    pia->sorted_layout[s] = 0;
    pia->h_valid = false;
    return MB_OK;
}

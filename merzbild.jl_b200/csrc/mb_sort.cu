// sort_particles! (grids/grid_sorting.jl:58-113 with a grid, :128-182 with known cells) for the device-resident SoA
// ParticleVector: a STABLE counting sort by cell that physically reorders the seven fp64 arrays (ping-pong buffers)
// and rebuilds the ParticleIndexerArray group-1 ranges on the device (group 2 cleared, contiguous = true).
//
// Bit-exact contract: logical order == ascending original logical position within each cell, pia
// fields == (count, start, end, count, 0, -1, 0) with (0,-1) for empty cells.
//
// Two algorithms produce the same output; both are launched back to back and select themselves through a device flag,
// so there is no host round trip:
//  (1) BAND path (the per-timestep case): the input is the previous sort's output after convection, i.e. the particles
//      of old cell c' are one contiguous segment and move at most `w` cells.  A warp per old cell classifies its segment
//      (x * inv_dx -> floor, same multiply as grid_uniform1D.jl:97-99), counts the (2w+1) destinations with ballots
//      [M(c', d)], a scan over cells gives the new cell starts, and destination offsets follow from the band matrix:
//      offset(c', c) = start(c) + sum_{c'' < c'} M(c'', c).  The scatter then writes every (c' -> c) group as one
//      contiguous run.  HBM traffic: 8 B (x) + 4 B (key) in pass 1, 4 + 56 + 56 B in pass 2 = 128 B / particle.
//  (2) GENERAL path (arbitrary input; taken when a particle leaves the band, or the layout is not sorted):
//      histogram with warp-aggregated atomics, scan, unstable atomic scatter of particle indices into the cell
//      buckets, per-cell ascending sort of the indices (== stable order), gather of the payload.
#include <climits>

#include "mb_common.cuh"

namespace mb {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;  // cells per scan block

struct SortScratch {
    int32_t* key;      // [cap]   0-based destination cell per logical position
    int32_t* hist;     // [n_cells]
    int64_t* start;    // [n_cells + 1] exclusive prefix (0-based offsets)
    int64_t* partial;  // [n_scan_blocks + 1]
    int32_t* M;        // [n_cells * W] band matrix
    int64_t* O;        // [n_cells * W] destination offsets
    int32_t* cursor;   // [n_cells] (general path)
    int32_t* perm;     // [cap]     (general path)
    int* flags;        // ctx->d_flags
};

// flags[2] = 1 -> the general path must run (band overflow or band not applicable)

__device__ __forceinline__ int cell_of(double x, double inv_dx, int64_t cell_offset) {
    return (int)((int64_t)floor(x * inv_dx) - cell_offset);  // get_cell - 1, grid_uniform1D.jl:97-99
}

// ------------------------------------------------------------------------------------------------ band path
template <int W>
__global__ void __launch_bounds__(256) k_band_classify(const double* __restrict__ X, const int32_t* cell_in, int32_t* cell_out,
                                                       int32_t* __restrict__ key, const Indexer* __restrict__ ix, int64_t n_cells, double inv_dx,
                                                       int64_t cell_offset, int use_x, int32_t* __restrict__ M, int* flags) {
    constexpr int w = W / 2;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const Indexer q = ix[c];
        const int64_t lo = q.start1 - 1, n = q.n_group1;
        int cnt[W];
#pragma unroll
        for (int d = 0; d < W; d++) cnt[d] = 0;
        bool bad = false;
        for (int64_t b = 0; b < n; b += 32) {
            const int64_t i = lo + b + lane;
            const bool valid = b + lane < n;
            int nc = 0;
            if (valid) {
                if (use_x) {
                    nc = cell_of(X[i], inv_dx, cell_offset);
                    cell_out[i] = nc + 1;
                } else {
                    nc = cell_in[i] - 1;
                }
                key[i] = nc;
            }
            const int64_t dd = (int64_t)nc - c + w;
            const bool inband = dd >= 0 && dd < W && nc >= 0 && nc < n_cells;
            if (valid && !inband) bad = true;
#pragma unroll
            for (int d = 0; d < W; d++) cnt[d] += __popc(__ballot_sync(0xffffffffu, valid && inband && dd == d));
        }
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&flags[2], 1);
#pragma unroll
        for (int d = 0; d < W; d++)
            if (lane == d) M[c * W + d] = cnt[d];
    }
}

// hist[c] = sum over sources of M(c', c)
template <int W>
__device__ __forceinline__ int band_hist(const int32_t* __restrict__ M, int64_t c, int64_t n_cells) {
    constexpr int w = W / 2;
    int h = 0;
#pragma unroll
    for (int d = 0; d < W; d++) {
        const int64_t cs = c - (d - w);
        if (cs >= 0 && cs < n_cells) h += M[cs * W + d];
    }
    return h;
}

// scan step 1: per-block sums of the per-cell counts (band: derived from M and stored to hist; general: hist given)
template <int W>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_reduce(const int32_t* __restrict__ M, int32_t* __restrict__ hist, int64_t n_cells,
                                                           int64_t* __restrict__ partial, const int* flags, int mode) {
    // mode 0: band (runs always; cheap), mode 1: general (runs only if flags[2])
    if (mode == 1 && flags[2] == 0) return;
    __shared__ int64_t red[SCAN_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t c = base + k * SCAN_BLOCK + threadIdx.x;
        if (c < n_cells) {
            int h;
            if (mode == 0) { h = band_hist<W>(M, c, n_cells); hist[c] = h; }
            else h = hist[c];
            s += h;
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int i = 0; i < SCAN_BLOCK / 32; i++) t += red[i];
        partial[blockIdx.x] = t;
    }
}
// scan step 2: exclusive scan of the block sums (single block)
__global__ void __launch_bounds__(1024) k_scan_partials(int64_t* __restrict__ partial, int64_t nb, const int* flags, int mode) {
    if (mode == 1 && flags[2] == 0) return;
    __shared__ int64_t sh[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nb ? partial[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const int64_t incl = sh[threadIdx.x];
        const int64_t c0 = carry;
        if (i < nb) partial[i] = c0 + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[nb] = carry;
}
// scan step 3: per-cell exclusive prefix + pia rebuild (grid_sorting.jl:76-96)
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const int32_t* __restrict__ hist, int64_t n_cells, const int64_t* __restrict__ partial,
                                                          int64_t* __restrict__ start, int32_t* __restrict__ cursor, Indexer* __restrict__ ix,
                                                          int64_t* n_total, const int* flags, int mode) {
    if (mode == 1 && flags[2] == 0) return;
    __shared__ int64_t wsum[SCAN_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;  // blocked arrangement
    int h[SCAN_ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t c = base + k;
        h[k] = c < n_cells ? hist[c] : 0;
        tsum += h[k];
    }
    // block exclusive scan of tsum
    int64_t incl = tsum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int64_t woff = 0;
    for (int i = 0; i < wid; i++) woff += wsum[i];
    int64_t run = partial[blockIdx.x] + woff + incl - tsum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t c = base + k;
        if (c < n_cells) {
            start[c] = run;
            if (cursor) cursor[c] = 0;
            const int64_t np = h[k];
            Indexer q;
            q.n_local = np; q.n_group1 = np;
            q.start1 = np > 0 ? run + 1 : 0;
            q.end1 = np > 0 ? run + np : -1;
            q.start2 = 0; q.end2 = -1; q.n_group2 = 0;
            ix[c] = q;
            run += np;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) start[n_cells] = partial[gridDim.x];  // n_total is unchanged by a sort
}

// destination offsets O(c', d) = start(c) + sum_{c'' in [c-w, c'-1]} M(c'', c), c = c' + d - w
template <int W>
__global__ void __launch_bounds__(256) k_band_offsets(const int32_t* __restrict__ M, const int64_t* __restrict__ start, int64_t* __restrict__ O,
                                                      int64_t n_cells) {
    constexpr int w = W / 2;
    const int64_t total = n_cells * W;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cs = t / W;
        const int d = (int)(t - cs * W);
        const int64_t c = cs + d - w;
        int64_t o = -1;
        if (c >= 0 && c < n_cells) {
            o = start[c];
            for (int64_t c2 = (c - w > 0 ? c - w : 0); c2 < cs; c2++) o += M[c2 * W + (int)(c - c2 + w)];
        }
        O[t] = o;
    }
}

template <int W>
__global__ void __launch_bounds__(256) k_band_scatter(SoA in, SoA out, const int32_t* __restrict__ key, const Indexer* __restrict__ ix_old_ranges_lo,
                                                      const int64_t* __restrict__ seg_lo, const int32_t* __restrict__ seg_n, int64_t n_cells,
                                                      const int64_t* __restrict__ O, const int* flags) {
    if (flags[2] != 0) return;  // general path takes over
    constexpr int w = W / 2;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const int64_t lo = seg_lo[c];
        const int64_t n = seg_n[c];
        int64_t off[W];
#pragma unroll
        for (int d = 0; d < W; d++) off[d] = O[c * W + d];
        for (int64_t b = 0; b < n; b += 32) {
            const int64_t i = lo + b + lane;
            const bool valid = b + lane < n;
            int dd = -1;
            double p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0;
            if (valid) {
                dd = (int)((int64_t)key[i] - c + w);
                p0 = in.a[0][i]; p1 = in.a[1][i]; p2 = in.a[2][i]; p3 = in.a[3][i]; p4 = in.a[4][i]; p5 = in.a[5][i]; p6 = in.a[6][i];
            }
            int64_t dst = -1;
#pragma unroll
            for (int d = 0; d < W; d++) {
                const unsigned bal = __ballot_sync(0xffffffffu, dd == d);
                if (dd == d) dst = off[d] + __popc(bal & lt);
                off[d] += __popc(bal);
            }
            if (valid) {
                out.a[0][dst] = p0; out.a[1][dst] = p1; out.a[2][dst] = p2; out.a[3][dst] = p3; out.a[4][dst] = p4; out.a[5][dst] = p5;
                out.a[6][dst] = p6;
            }
        }
    }
}

// the old segment table must survive the pia rebuild: (lo, n) per old cell
__global__ void k_save_segments(const Indexer* __restrict__ ix, int64_t n_cells, int64_t* __restrict__ seg_lo, int32_t* __restrict__ seg_n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[c];
        seg_lo[c] = q.start1 - 1;
        seg_n[c] = (int32_t)q.n_group1;
    }
}

// ------------------------------------------------------------------------------------------------ general path
__global__ void __launch_bounds__(256) k_gen_classify(const double* __restrict__ X, const int32_t* cell_in, int32_t* cell_out,
                                                      int32_t* __restrict__ key, const int64_t* n_total_p, int64_t n_cells, double inv_dx,
                                                      int64_t cell_offset, int use_x, int need_key, int32_t* __restrict__ hist, int* flags) {
    if (flags[2] == 0) return;
    const int64_t n_total = *n_total_p;
    const int lane = threadIdx.x & 31;
    // iterate in warp-uniform fashion so the aggregated atomics see whole warps
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n_total + stride - 1) / stride;
    for (int64_t r = 0; r < nround; r++) {
        const int64_t i = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = i < n_total;
        int nc = -1;
        if (valid) {
            if (need_key) {
                if (use_x) { nc = cell_of(X[i], inv_dx, cell_offset); cell_out[i] = nc + 1; }
                else nc = cell_in[i] - 1;
                if (nc < 0 || nc >= n_cells) { atomicOr(&flags[0], DEVERR_BAD_CELL); nc = nc < 0 ? 0 : (int)(n_cells - 1); }
                key[i] = nc;
            } else {
                nc = key[i];
                if (nc < 0 || nc >= n_cells) { atomicOr(&flags[0], DEVERR_BAD_CELL); nc = nc < 0 ? 0 : (int)(n_cells - 1); key[i] = nc; }
            }
        }
        // run-length aggregation: keys of neighbouring lanes are mostly equal
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(act, nc);
            const int leader = __ffs(peers) - 1;
            if (lane == leader) atomicAdd(&hist[nc], __popc(peers));
        }
    }
}

__global__ void __launch_bounds__(256) k_gen_scatter_idx(const int32_t* __restrict__ key, const int64_t* n_total_p, const int64_t* __restrict__ start,
                                                         int32_t* __restrict__ cursor, int32_t* __restrict__ perm, const int* flags) {
    if (flags[2] == 0) return;
    const int64_t n_total = *n_total_p;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n_total + stride - 1) / stride;
    for (int64_t r = 0; r < nround; r++) {
        const int64_t i = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        const bool valid = i < n_total;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int nc = key[i];
            const unsigned peers = __match_any_sync(act, nc);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cursor[nc], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            perm[start[nc] + base + __popc(peers & lt)] = (int32_t)i;
        }
    }
}

// per-cell ascending sort of the scattered indices (ascending index == stable order).  One CTA per cell.
// "Normalised" bitonic network (every compare-exchange is ascending; the first step of each merge mirrors), so an
// arbitrary length works with virtual +inf padding: a pair whose upper element is >= n is simply skipped.
constexpr int SEG_SMEM = 8192;
template <class T>
__device__ __forceinline__ void bitonic_ascending(T* a, int64_t n) {
    int64_t m = 1;
    while (m < n) m <<= 1;
    for (int64_t k = 2; k <= m; k <<= 1) {
        const int64_t hk = k >> 1;
        for (int64_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
            const int64_t blk = t / hk, off = t - blk * hk;
            const int64_t i = blk * k + off, p = blk * k + k - 1 - off;
            if (p < n) {
                const T x = a[i], y = a[p];
                if (x > y) { a[i] = y; a[p] = x; }
            }
        }
        __syncthreads();
        for (int64_t j = k >> 2; j > 0; j >>= 1) {
            for (int64_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                const int64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int64_t p = i + j;
                if (p < n) {
                    const T x = a[i], y = a[p];
                    if (x > y) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}
__global__ void __launch_bounds__(256) k_gen_sort_segments(int32_t* __restrict__ perm, const int64_t* __restrict__ start, int64_t n_cells,
                                                           const int* flags) {
    if (flags[2] == 0) return;
    __shared__ int32_t sh[SEG_SMEM];
    __shared__ int unsorted;
    for (int64_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const int64_t lo = start[c];
        const int64_t n = start[c + 1] - lo;
        if (n <= 1) continue;  // block-uniform
        int32_t* seg = perm + lo;
        __syncthreads();
        if (threadIdx.x == 0) unsorted = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i + 1 < n; i += blockDim.x)
            if (seg[i] > seg[i + 1]) unsorted = 1;
        __syncthreads();
        if (!unsorted) continue;  // block-uniform (read after the barrier)
        if (n <= SEG_SMEM) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = seg[i];
            __syncthreads();
            bitonic_ascending(sh, n);
            for (int i = threadIdx.x; i < n; i += blockDim.x) seg[i] = sh[i];
        } else {
            bitonic_ascending(seg, n);  // large cell: network directly in global memory (L2-resident)
        }
    }
}

__global__ void __launch_bounds__(256) k_gen_gather(SoA in, SoA out, const int32_t* __restrict__ perm, const int64_t* n_total_p, const int* flags) {
    if (flags[2] == 0) return;
    const int64_t n_total = *n_total_p;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_total; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[j];
#pragma unroll
        for (int f = 0; f < 7; f++) out.a[f][j] = in.a[f][i];
    }
}

__global__ void k_set_flag(int* flags, int idx, int v) { flags[idx] = v; }

template <int W>
static int launch_band(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species, SortScratch& S, int64_t* seg_lo, int32_t* seg_n) {
    const int64_t nc = pia->n_cells;
    Indexer* ix = pia->d_indexer + (species - 1) * nc;
    cudaStream_t st = ctx->stream;
    const int use_x = grid != nullptr;
    const int nscan = (int)((nc + SCAN_TILE - 1) / SCAN_TILE);
    const int wgrid = grid_for(nc * 32, 256, 8);
    {
        ProfScope ps(ctx, PROF_SORT_CLASSIFY);
        k_save_segments<<<grid_for(nc, 256), 256, 0, st>>>(ix, nc, seg_lo, seg_n);
        MB_LAUNCH_CHECK(ctx);
        k_band_classify<W><<<wgrid, 256, 0, st>>>(pv->cur.a[F_X], pv->cell, pv->cell, S.key, ix, nc, use_x ? grid->inv_dx : 0.0,
                                                 use_x ? grid->cell_offset : 0, use_x, S.M, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, PROF_SORT_SCAN);
        k_scan_reduce<W><<<nscan, SCAN_BLOCK, 0, st>>>(S.M, S.hist, nc, S.partial, S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
        k_scan_partials<<<1, 1024, 0, st>>>(S.partial, nscan, S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
        k_scan_apply<<<nscan, SCAN_BLOCK, 0, st>>>(S.hist, nc, S.partial, S.start, nullptr, ix, pia->d_n_total + (species - 1), S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
        k_band_offsets<W><<<grid_for(nc * W, 256), 256, 0, st>>>(S.M, S.start, S.O, nc);
        MB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, PROF_SORT_SCATTER);
        k_band_scatter<W><<<wgrid, 256, 0, st>>>(pv->cur, pv->alt, S.key, ix, seg_lo, seg_n, nc, S.O, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_squash_pia(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, int64_t species);

int mb_sort_set_band_halfwidth(mb_ctx* ctx, int32_t w) {
    MB_ARG(ctx && (w == 0 || w == 1 || w == 2 || w == 4 || w == 8), "band half-width must be 0, 1, 2, 4 or 8");
    ctx->band_w = w;
    return MB_OK;
}
int mb_sort_last_path(mb_ctx* ctx) {
    if (!ctx) return -1;
    if (mb_sync(ctx)) return -1;
    return ctx->sort_last_path == 1 && ctx->h_flags[2] == 0 ? 1 : 2;
}

int mb_sort_particles(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species) {
    MB_ARG(ctx && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(grid == nullptr || grid->n_cells == pia->n_cells, "grid.n_cells != pia.n_cells");
    MB_ARG(pia->n_cells < (int64_t)INT_MAX && pv->cap < (int64_t)INT_MAX, "more than 2^31 cells or particles per GPU");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    if (!pia->contiguous[s]) {  // grid_sorting.jl:69-71
        int r = mb_squash_pia(ctx, pv, pia, species);
        if (r) return r;
    }
    int r = pv_ensure_alt(pv);
    if (r) return r;
    const int64_t nc = pia->n_cells, cap = pv->cap;
    const int w = ctx->band_w;
    const int W = 2 * w + 1;
    const bool try_band = w > 0 && pia->sorted_layout[s];
    const int nscan = (int)((nc + SCAN_TILE - 1) / SCAN_TILE);

    SortScratch S;
    S.flags = ctx->d_flags;
    S.key = (int32_t*)ctx_scratch(ctx, 0, (size_t)cap * 4);
    // slot 1: hist | cursor | seg_n | M   (int32)
    const size_t n32 = (size_t)nc * (3 + (size_t)W) + 64;
    int32_t* p32 = (int32_t*)ctx_scratch(ctx, 1, n32 * 4);
    // slot 2: start | partial | seg_lo | O   (int64)
    const size_t n64 = (size_t)(nc + 1) + (size_t)(nscan + 2) + (size_t)nc + (size_t)nc * W + 64;
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 2, n64 * 8);
    if (!S.key || !p32 || !p64) return MB_ERR_CUDA;
    S.hist = p32;
    S.cursor = p32 + nc;
    int32_t* seg_n = p32 + 2 * nc;
    S.M = p32 + 3 * nc;
    S.start = p64;
    S.partial = p64 + (nc + 1);
    int64_t* seg_lo = S.partial + (nscan + 2);
    S.O = seg_lo + nc;
    cudaStream_t st = ctx->stream;
    Indexer* ix = pia->d_indexer + (species - 1) * nc;
    int64_t* d_nt = pia->d_n_total + (species - 1);

    k_set_flag<<<1, 1, 0, st>>>(ctx->d_flags, 2, try_band ? 0 : 1);
    MB_LAUNCH_CHECK(ctx);
    if (try_band) {
        if (w == 1) r = launch_band<3>(ctx, grid, pv, pia, species, S, seg_lo, seg_n);
        else if (w == 2) r = launch_band<5>(ctx, grid, pv, pia, species, S, seg_lo, seg_n);
        else if (w == 4) r = launch_band<9>(ctx, grid, pv, pia, species, S, seg_lo, seg_n);
        else r = launch_band<17>(ctx, grid, pv, pia, species, S, seg_lo, seg_n);
        if (r) return r;
    }
    // general path (every kernel returns immediately unless flags[2] != 0)
    {
        ProfScope ps(ctx, PROF_SORT_GENERAL);
        S.perm = (int32_t*)ctx_scratch(ctx, 3, (size_t)cap * 4);
        if (!S.perm) return MB_ERR_CUDA;
        const int use_x = grid != nullptr;
        MB_CUDA(cudaMemsetAsync(S.hist, 0, (size_t)nc * 4, st));  // harmless for the band result: hist is not read again
        const int pgrid = grid_for(pia->n_bound[s] > 0 ? pia->n_bound[s] : cap, 256, 16);
        // after a band attempt the keys are already classified (and n_total is still the old one: the band pass
        // re-wrote it with the same value), so reuse them
        k_gen_classify<<<pgrid, 256, 0, st>>>(pv->cur.a[F_X], pv->cell, pv->cell, S.key, d_nt, nc, use_x ? grid->inv_dx : 0.0,
                                             use_x ? grid->cell_offset : 0, use_x, 1, S.hist, S.flags);
        MB_LAUNCH_CHECK(ctx);
        k_scan_reduce<3><<<nscan, SCAN_BLOCK, 0, st>>>(nullptr, S.hist, nc, S.partial, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_scan_partials<<<1, 1024, 0, st>>>(S.partial, nscan, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_scan_apply<<<nscan, SCAN_BLOCK, 0, st>>>(S.hist, nc, S.partial, S.start, S.cursor, ix, d_nt, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_gen_scatter_idx<<<pgrid, 256, 0, st>>>(S.key, d_nt, S.start, S.cursor, S.perm, S.flags);
        MB_LAUNCH_CHECK(ctx);
        k_gen_sort_segments<<<grid_for(nc * 256, 256, 8), 256, 0, st>>>(S.perm, S.start, nc, S.flags);
        MB_LAUNCH_CHECK(ctx);
        k_gen_gather<<<pgrid, 256, 0, st>>>(pv->cur, pv->alt, S.perm, d_nt, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    // ping-pong
    SoA t = pv->cur;
    pv->cur = pv->alt;
    pv->alt = t;
    pia->contiguous[s] = 1;      // grid_sorting.jl:112
    pia->contig_pending[s] = 0;
    pia->sorted_layout[s] = 1;
    ctx->sort_last_path = try_band ? 1 : 2;
    return MB_OK;
}

}  // extern "C"

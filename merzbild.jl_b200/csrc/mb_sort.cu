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
//      of old cell c' are one contiguous segment and move at most `w` cells.  Pass A (a warp per old cell; fused into
//      convect_particles! when that call precedes the sort) writes for every particle its destination group
//      d = c - c' + w and its rank inside that group (original order) as one 32-bit word, and the group sizes M(c', d).
//      A scan over cells gives the new cell starts; pass B (a warp per old cell again) streams the cell once and stores
//      every particle at start(c) + sum_{c'' < c'} M(c'', c) + rank.  HBM traffic: pass A 8 B (x) [+ 8 B vx and 8 B x
//      written when fused with the convection] + 4 B (word); pass B 4 + 56 + 56 B = 116 B / particle.
//      HYBRID: a particle that leaves the band (|c - c'| > w) or arrives from a neighbouring slab is an "extra": it is appended
//      to a short list (original position, destination cell) and counted per destination cell as BEFORE the band (it came from a
//      source cell < c - w) or AFTER it (source > c + w, or a slab-exchange arrival).  The destination cell is laid out as
//      [extras before | band groups by source cell | extras after]; the extras are ranked among themselves by their original
//      position, so the result is still the reference's stable order, and the band work is never thrown away for a few outliers.
//  (2) GENERAL path (arbitrary input; taken when the layout is not sorted, or the extras do not fit their list):
//      histogram with warp-aggregated atomics, scan, unstable atomic scatter of particle indices into the cell
//      buckets, per-cell ascending sort of the indices (== stable order), gather of the payload.
#include <climits>
#include <cstdlib>

#include "mb_convect.cuh"
#include "mb_scan.cuh"
#include "mb_segcopy.cuh"
#include "mb_sort_tile.cuh"

namespace mb {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;  // cells per scan block

struct SortScratch {
    int32_t* key;      // [cap]   general path: 0-based destination cell per logical position; band path: dr
    int32_t* hist;     // [n_cells]
    int64_t* start;    // [n_cells + 1] exclusive prefix (0-based offsets)
    int64_t* partial;  // [n_scan_blocks + 1]
    int32_t* M;        // [n_cells * W] band matrix
    int64_t* O;        // [n_cells * W] destination offsets
    int32_t* cursor;   // [n_cells] (general path)
    int32_t* perm;     // [cap]     (general path); band path: slot[] = original positions of the extras, indexed by output position
    int* flags;        // ctx->d_flags
};

// the extras of the hybrid band path (see the header comment)
struct ExBufs {
    int32_t* n;        // number of extras appended (may exceed cap: overflow -> general path)
    int32_t* idx;      // [cap] original logical position (0-based)
    int32_t* cell;     // [cap] destination cell * 2 + class (0: before the band, 1: after it)
    int32_t* cntB;     // [n_cells] extras before the band per destination cell
    int32_t* cntA;     // [n_cells] extras after the band (incl. slab-exchange arrivals)
    int32_t* curB;     // [n_cells] slot cursors
    int32_t* curA;
    int32_t cap;
};
constexpr int EX_REGION_MAX = 1024;  // extras per (cell, class) a warp orders in shared memory; beyond that the general path runs

// All lanes of the warp call it; returns true if the list is full (the caller requests the general path).
__device__ __forceinline__ bool extras_append(const ExBufs& E, bool is, int32_t i, int nc, int cls, int lane, unsigned lt) {
    const unsigned m = __ballot_sync(0xffffffffu, is);
    if (m == 0) return false;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(E.n, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    bool ovf = false;
    if (is) {
        const int e = base + __popc(m & lt);
        if (e >= 0 && e < E.cap) {
            E.idx[e] = i;
            E.cell[e] = nc * 2 + cls;
            atomicAdd((cls ? E.cntA : E.cntB) + nc, 1);
        } else ovf = true;
    }
    return ovf;
}

// flags[2] = 1 -> the general path must run (band overflow or band not applicable)

__device__ __forceinline__ int cell_of(double x, double inv_dx, int64_t cell_offset) {
    return (int)((int64_t)floor(x * inv_dx) - cell_offset);  // get_cell - 1, grid_uniform1D.jl:97-99
}

// ------------------------------------------------------------------------------------------------ band path
// Pass A, a warp per OLD cell c': the destination group d = c - c' + w of every particle and its rank inside that group
// (original order), stored as one word per particle (`dr` = d << 24 | rank; 0xFFFFFFFF: not kept), and the group sizes
// M(c', d).  Particles outside the slab are skipped when drop != 0 (they were sent to a neighbour by mb_exchange_slab).
// classify_batch is shared with the fused convect kernel further down.
constexpr uint32_t DR_NONE = 0xFFFFFFFFu;

// One batch of <= 32 particles of the warp's cell: d (255: not kept) per lane -> dr, running group counters in cnt_s[].
// Every lane of the warp must call it.
__device__ __forceinline__ void classify_batch(int d, bool valid, unsigned lt, int* cnt_s, uint32_t* __restrict__ dr_out) {
    const unsigned act = __ballot_sync(0xffffffffu, d != 255);
    unsigned peers = 0;
    int before = 0;
    if (d != 255) {
        peers = __match_any_sync(act, d);
        before = cnt_s[d];
        *dr_out = ((uint32_t)d << 24) | (uint32_t)(before + __popc(peers & lt));
    } else if (valid) {
        *dr_out = DR_NONE;
    }
    __syncwarp();
    if (d != 255 && (peers & lt) == 0) cnt_s[d] = before + __popc(peers);  // one leader per destination
    __syncwarp();
}

template <int W>
__global__ void __launch_bounds__(256) k_band_classify(const double* __restrict__ X, const int32_t* cell_in, int32_t* cell_out,
                                                       const Indexer* __restrict__ ix, int64_t n_cells, double inv_dx, int64_t cell_offset,
                                                       int use_x, int drop, int32_t* __restrict__ M, int64_t* __restrict__ seg_lo,
                                                       int32_t* __restrict__ seg_n, uint32_t* __restrict__ dr, ExBufs E, int* flags, const int* only_if) {
    if (only_if != nullptr && *only_if == 0) return;  // the fused convect kernel already classified every cell
    constexpr int w = W / 2;
    __shared__ int s_cnt[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int* cnt_s = s_cnt[wid];
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const Indexer q = ix[c];
        const int64_t lo = q.start1 - 1, n = q.n_group1;
        if (lane == 0) { seg_lo[c] = lo; seg_n[c] = (int32_t)n; }
        if (n >= (1 << 24) || q.n_group2 != 0) {  // the rank has 24 bits; group 2 must be empty in a sorted layout
            if (lane == 0) atomicOr(&flags[2], 1);
            continue;
        }
        __syncwarp();
        cnt_s[lane] = 0;
        __syncwarp();
        bool bad = false;
        for (int64_t b = 0; b < n; b += 32) {
            const int64_t i = lo + b + lane;
            const bool valid = b + lane < n;
            int d = 255, nc = 0, cls = 0;
            bool ex = false;
            if (valid) {
                if (use_x) { nc = cell_of(X[i], inv_dx, cell_offset); cell_out[i] = nc + 1; }
                else nc = cell_in[i] - 1;
                const bool inside = nc >= 0 && nc < n_cells;
                const int64_t dd = (int64_t)nc - c + w;
                if (inside) {
                    if (dd >= 0 && dd < W) d = (int)dd;
                    else { ex = true; cls = dd < 0 ? 1 : 0; }  // moved further than w cells: an extra of the destination cell
                } else if (!drop) bad = true;                   // outside the slab and no exchange: the general path reports it
                // edge exchange: a leaver from a cell further than w from that slab face was never sent
                if (!inside && drop == 2 && ((nc < 0 && c >= w) || (nc >= n_cells && c < n_cells - w)))
                    atomicOr(&flags[0], DEVERR_BAND_OVERFLOW);
            }
            classify_batch(d, valid, lt, cnt_s, dr + (valid ? i : 0));
            if (extras_append(E, ex, (int32_t)i, nc, cls, lane, lt)) bad = true;
        }
        if (__any_sync(0xffffffffu, bad)) {
            if (lane == 0) atomicOr(&flags[2], 1);
            continue;
        }
        if (lane < W) M[c * W + lane] = cnt_s[lane];
    }
}

// arrivals of the slab exchange (appended after the old layout): extras of class "after" of their cells (their original positions
// follow every particle of the old layout, in arrival order)
static __global__ void __launch_bounds__(256) k_band_arrivals(const double* __restrict__ X, const int64_t* n_total_p, const int64_t* n_arr_p,
                                                             int64_t n_cells, double inv_dx, int64_t cell_offset, int32_t* cell_out, ExBufs E,
                                                             int* flags) {
    const int64_t n_arr = *n_arr_p;
    const int64_t base = *n_total_p - n_arr;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n_arr + stride - 1) / stride;
    for (int64_t r = 0; r < nround; r++) {  // warp-uniform trip count
        const int64_t t = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        bool ex = false;
        int nc = 0;
        if (t < n_arr) {
            nc = cell_of(X[base + t], inv_dx, cell_offset);
            cell_out[base + t] = nc + 1;
            if (nc < 0 || nc >= n_cells) atomicOr(&flags[2], 1);  // the general path reports it
            else ex = true;
        }
        if (extras_append(E, ex, (int32_t)(base + t), nc, 1, lane, lt)) atomicOr(&flags[2], 1);
    }
}

// hist[c] = sum over sources of M(c', c) (+ arrivals)
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
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_reduce(const int32_t* __restrict__ M, const int32_t* __restrict__ cntB,
                                                           const int32_t* __restrict__ cntA, int32_t* __restrict__ hist, int64_t n_cells,
                                                           int64_t* __restrict__ partial, int* flags, int mode) {
    // mode 0: band (runs always; cheap), mode 1: general (runs only if flags[2])
    if (mode == 1 && flags[2] == 0) return;
    __shared__ int64_t red[SCAN_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const int64_t c = base + k * SCAN_BLOCK + threadIdx.x;
        if (c < n_cells) {
            int h;
            if (mode == 0) {
                const int eb = cntB[c], ea = cntA[c];
                if (eb > EX_REGION_MAX || ea > EX_REGION_MAX) atomicOr(&flags[2], 1);  // too many extras in one cell for the region sort
                h = band_hist<W>(M, c, n_cells) + eb + ea;
                hist[c] = h;
            }
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
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        start[n_cells] = partial[gridDim.x];
        if (n_total) *n_total = partial[gridDim.x];  // only after a slab exchange (leavers dropped); otherwise a sort keeps n_total
    }
}

// Pass B, a warp per OLD cell c': the cell's segment streams through registers exactly once, fully coalesced (56 B per
// particle, every sector read once), and every particle is stored at   start(c) + sum_{c'' < c'} M(c'', c) + rank   --
// ascending original position inside every destination cell, i.e. exactly the reference's stable counting sort.  ~95 % of a
// cell stays (one contiguous, coalesced run); the movers form short contiguous runs in the neighbour cells whose partial
// sectors merge in L2 (measured: 14.9 GB of DRAM traffic for 14.0 GB of payload at 1.25e8 particles, where a
// gather by destination cell needs 18.3 GB because every mover costs seven scattered sector reads).
// Moments: the staying group's shifted sums (shift K(c) = velocity of the first particle of old cell c) go to P[c][5];
// k_band_combine adds the few movers and arrivals of each destination cell from the output arrays in a fixed order.
#ifndef MB_SC_U
#define MB_SC_U 4
#endif
#ifndef MB_SC_GRID
#define MB_SC_GRID 8
#endif
#ifndef MB_SC_MINB
#define MB_SC_MINB 2
#endif
#ifndef MB_SC_BUF
#define MB_SC_BUF 1  // 1: k_band_scatter_buf (movers coalesced through shared memory), 0: k_band_scatter (every particle stored directly)
#endif
constexpr int SC_U = MB_SC_U;  // particles per lane in flight
#define MB_LD(p) (*(p))  // default cache policy: streaming hints (ld.cs / st.cs) measured 13 % slower here
#define MB_ST(p, v) (*(p) = (v))
template <int W, bool MOM>
__global__ void __launch_bounds__(256, MB_SC_MINB) k_band_scatter(SoA in_, SoA out_, const uint32_t* __restrict__ dr, const int32_t* __restrict__ M,
                                                      const int64_t* __restrict__ seg_lo, const int32_t* __restrict__ seg_n,
                                                      const int64_t* __restrict__ start, const int32_t* __restrict__ cntB, int64_t n_cells,
                                                      double* __restrict__ P, const int* flags) {
    if (flags[2] != 0) return;  // general path takes over
    constexpr int w = W / 2;
    __shared__ int64_t s_off[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t* off_s = s_off[wid];
    // what a warp needs to know about a cell before it can stream it: the segment (lo, n) and, in lane d, the output position of
    // destination group d.  All loads are independent of each other, and the NEXT cell's are issued before the current cell streams,
    // so that no warp ever waits for metadata with nothing in flight.
    auto load_meta = [&](int64_t c, int64_t& lo, int& n, int64_t& off) {
        lo = seg_lo[c];
        n = seg_n[c];
        off = 0;
        const int64_t cd = c + lane - w;  // destination cell of group d = lane
        if (lane < W && cd >= 0 && cd < n_cells) {
            int acc = 0;
#pragma unroll
            for (int k = 1; k < W; k++) {  // sources c - k < c that also feed cd
                const int64_t cs = c - k;
                if (lane + k < W && cs >= 0) acc += M[cs * W + lane + k];
            }
            off = start[cd] + cntB[cd] + acc;  // the extras that came from further left sit in front of the band groups
        }
    };
    int64_t lo = 0, off = 0, lo_n = 0, off_n = 0;
    int n = 0, n_n = 0;
    int64_t c = warp0;
    if (c < n_cells) load_meta(c, lo, n, off);
    for (; c < n_cells; c += nwarps) {
        if (c + nwarps < n_cells) load_meta(c + nwarps, lo_n, n_n, off_n);
        __syncwarp();
        off_s[lane] = off;
        __syncwarp();
        const double* __restrict__ i0 = in_.a[0] + lo; const double* __restrict__ i1 = in_.a[1] + lo; const double* __restrict__ i2 = in_.a[2] + lo;
        const double* __restrict__ i3 = in_.a[3] + lo; const double* __restrict__ i4 = in_.a[4] + lo; const double* __restrict__ i5 = in_.a[5] + lo;
        const double* __restrict__ i6 = in_.a[6] + lo;
        const uint32_t* __restrict__ drc = dr + lo;
        double K1 = 0, K2 = 0, K3 = 0;
        if (MOM && n > 0) { K1 = i1[0]; K2 = i2[0]; K3 = i3[0]; }
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;  // the staying group
        for (int j0 = 0; j0 < n; j0 += 32 * SC_U) {  // SC_U particles per lane in flight; the last round is predicated, not serialised
            uint32_t v[SC_U];
            double a[SC_U][7];
#pragma unroll
            for (int u = 0; u < SC_U; u++) {
                const int ju = j0 + 32 * u + lane;
                v[u] = DR_NONE;
                if (ju < n) {
                    v[u] = drc[ju];
                    a[u][0] = MB_LD(i0 + ju); a[u][1] = MB_LD(i1 + ju); a[u][2] = MB_LD(i2 + ju); a[u][3] = MB_LD(i3 + ju);
                    a[u][4] = MB_LD(i4 + ju); a[u][5] = MB_LD(i5 + ju); a[u][6] = MB_LD(i6 + ju);
                }
            }
#pragma unroll
            for (int u = 0; u < SC_U; u++) {
                if (v[u] != DR_NONE) {
                    const int64_t pos = off_s[v[u] >> 24] + (int64_t)(v[u] & 0xFFFFFFu);
                    MB_ST(&out_.a[0][pos], a[u][0]); MB_ST(&out_.a[1][pos], a[u][1]); MB_ST(&out_.a[2][pos], a[u][2]);
                    MB_ST(&out_.a[3][pos], a[u][3]); MB_ST(&out_.a[4][pos], a[u][4]); MB_ST(&out_.a[5][pos], a[u][5]);
                    MB_ST(&out_.a[6][pos], a[u][6]);
                    if (MOM && (v[u] >> 24) == (uint32_t)w) {
                        const double cx_ = a[u][1] - K1, cy_ = a[u][2] - K2, cz_ = a[u][3] - K3;
                        an += a[u][0]; ax += a[u][0] * cx_; ay += a[u][0] * cy_; az += a[u][0] * cz_;
                        aq += a[u][0] * (cx_ * cx_ + cy_ * cy_ + cz_ * cz_);
                    }
                }
            }
        }
        if (MOM) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                an += __shfl_xor_sync(0xffffffffu, an, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
                aq += __shfl_xor_sync(0xffffffffu, aq, o);
            }
            if (lane == 0) {
                double* q = P + c * 5;
                q[0] = an; q[1] = ax; q[2] = ay; q[3] = az; q[4] = aq;
            }
        }
        lo = lo_n; n = n_n; off = off_n;
    }
}

// Pass B with the movers coalesced (the variant that runs).  Measured with the kernel above: with no movers at all the segmented copy
// reaches 90 % of the copy peak, and every mover costs seven isolated 8-byte stores -- partial-sector writes that L2 has to merge
// (same-dx: +20 % write transactions, -10 % bandwidth; at sigma_v dt = 2.6 cells, where 85 % of a cell moves: -35 %).  Here the
// stayers still go straight from registers to their run, but a mover is parked in the warp's shared-memory buffer of its destination
// group (slot = rank: the rank IS the position inside the group), and when the cell is through every group leaves as one run,
// consecutive ranks in consecutive lanes: full sectors instead of single doubles.  Ranks beyond the buffer are stored directly.
// No synchronisation inside the streaming loop.
template <int W>
struct ScBuf {
    static constexpr int ND = W - 1;                          // mover groups
    static constexpr int CAP = W >= 31 ? 8 : 192 / ND;        // ranks parked per group: 96 / 48 / 24 / 12 / 8 for w = 1 / 2 / 4 / 8 / 15
    static constexpr int WARP_DOUBLES = ND * 7 * CAP;         // 10.5 KB per warp (13.1 KB for the 31-wide band)
    static constexpr int SMEM = 8 * (WARP_DOUBLES * 8 + 32 * 8);  // + the per-warp output offsets
};
template <int W, bool MOM>
__global__ void __launch_bounds__(256, MB_SC_MINB) k_band_scatter_buf(SoA in_, SoA out_, const uint32_t* __restrict__ dr, const int32_t* __restrict__ M,
                                                          const int64_t* __restrict__ seg_lo, const int32_t* __restrict__ seg_n,
                                                          const int64_t* __restrict__ start, const int32_t* __restrict__ cntB, int64_t n_cells,
                                                          double* __restrict__ P, const int* flags) {
    if (flags[2] != 0) return;  // general path takes over
    using B = ScBuf<W>;
    constexpr int w = W / 2, CAP = B::CAP;
    extern __shared__ __align__(16) unsigned char sc_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* buf = (double*)sc_smem + (size_t)wid * B::WARP_DOUBLES;                  // [group][field][rank]
    int64_t* off_s = (int64_t*)((double*)sc_smem + 8 * B::WARP_DOUBLES) + wid * 32;  // output position of group d
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double* const out0 = out_.a[0];
    const int64_t ostride = out_.a[1] - out_.a[0];  // the seven arrays of a ParticleVector are equally spaced slices of one allocation
    auto load_meta = [&](int64_t c, int64_t& lo, int& n, int64_t& off, int& gcnt) {
        lo = seg_lo[c];
        n = seg_n[c];
        off = 0;
        gcnt = lane < W ? M[c * W + lane] : 0;  // size of group d = lane
        const int64_t cd = c + lane - w;        // its destination cell
        if (lane < W && cd >= 0 && cd < n_cells) {
            int acc = 0;
#pragma unroll
            for (int k = 1; k < W; k++) {  // sources c - k < c that also feed cd
                const int64_t cs = c - k;
                if (lane + k < W && cs >= 0) acc += M[cs * W + lane + k];
            }
            off = start[cd] + cntB[cd] + acc;  // the extras that came from further left sit in front of the band groups
        }
    };
    int64_t lo = 0, off = 0, lo_n = 0, off_n = 0;
    int n = 0, n_n = 0, gcnt = 0, gcnt_n = 0;
    int64_t c = warp0;
    if (c < n_cells) load_meta(c, lo, n, off, gcnt);
    for (; c < n_cells; c += nwarps) {
        if (c + nwarps < n_cells) load_meta(c + nwarps, lo_n, n_n, off_n, gcnt_n);
        __syncwarp();
        off_s[lane] = off;
        __syncwarp();
        const double* __restrict__ i0 = in_.a[0] + lo; const double* __restrict__ i1 = in_.a[1] + lo; const double* __restrict__ i2 = in_.a[2] + lo;
        const double* __restrict__ i3 = in_.a[3] + lo; const double* __restrict__ i4 = in_.a[4] + lo; const double* __restrict__ i5 = in_.a[5] + lo;
        const double* __restrict__ i6 = in_.a[6] + lo;
        const uint32_t* __restrict__ drc = dr + lo;
        double K1 = 0, K2 = 0, K3 = 0;
        if (MOM && n > 0) { K1 = i1[0]; K2 = i2[0]; K3 = i3[0]; }
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;  // the staying group
        for (int j0 = 0; j0 < n; j0 += 32 * SC_U) {
            uint32_t v[SC_U];
            double a[SC_U][7];
#pragma unroll
            for (int u = 0; u < SC_U; u++) {
                const int ju = j0 + 32 * u + lane;
                v[u] = DR_NONE;
                if (ju < n) {
                    v[u] = drc[ju];
                    a[u][0] = MB_LD(i0 + ju); a[u][1] = MB_LD(i1 + ju); a[u][2] = MB_LD(i2 + ju); a[u][3] = MB_LD(i3 + ju);
                    a[u][4] = MB_LD(i4 + ju); a[u][5] = MB_LD(i5 + ju); a[u][6] = MB_LD(i6 + ju);
                }
            }
#pragma unroll
            for (int u = 0; u < SC_U; u++) {
                if (v[u] == DR_NONE) continue;
                const int d = (int)(v[u] >> 24), r = (int)(v[u] & 0xFFFFFFu);
                if (d == w || r >= CAP) {  // stayer (or a rank beyond the buffer): straight to its place
                    const int64_t pos = off_s[d] + r;
                    MB_ST(&out_.a[0][pos], a[u][0]); MB_ST(&out_.a[1][pos], a[u][1]); MB_ST(&out_.a[2][pos], a[u][2]);
                    MB_ST(&out_.a[3][pos], a[u][3]); MB_ST(&out_.a[4][pos], a[u][4]); MB_ST(&out_.a[5][pos], a[u][5]);
                    MB_ST(&out_.a[6][pos], a[u][6]);
                    if (MOM && d == w) {
                        const double cx_ = a[u][1] - K1, cy_ = a[u][2] - K2, cz_ = a[u][3] - K3;
                        an += a[u][0]; ax += a[u][0] * cx_; ay += a[u][0] * cy_; az += a[u][0] * cz_;
                        aq += a[u][0] * (cx_ * cx_ + cy_ * cy_ + cz_ * cz_);
                    }
                } else {  // mover: parked at its rank
                    double* dst = buf + (d > w ? d - 1 : d) * (7 * CAP) + r;
                    dst[0] = a[u][0]; dst[CAP] = a[u][1]; dst[2 * CAP] = a[u][2]; dst[3 * CAP] = a[u][3];
                    dst[4 * CAP] = a[u][4]; dst[5 * CAP] = a[u][5]; dst[6 * CAP] = a[u][6];
                }
            }
        }
        __syncwarp();
        // every mover group leaves as one run: element t = (field, rank)
        for (int g = 0; g < W - 1; g++) {
            const int d = g < w ? g : g + 1;
            int cnt = __shfl_sync(0xffffffffu, gcnt, d);
            if (cnt == 0) continue;
            cnt = cnt < CAP ? cnt : CAP;
            const double* src = buf + g * (7 * CAP);
            double* o = out0 + off_s[d];
            for (int t = lane; t < 7 * CAP; t += 32) {
                const int f = t / CAP, e = t - f * CAP;
                if (e < cnt) MB_ST(o + f * ostride + e, src[t]);
            }
        }
        if (MOM) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                an += __shfl_xor_sync(0xffffffffu, an, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
                aq += __shfl_xor_sync(0xffffffffu, aq, o);
            }
            if (lane == 0) {
                double* q = P + c * 5;
                q[0] = an; q[1] = ax; q[2] = ay; q[3] = az; q[4] = aq;
            }
        }
        lo = lo_n; n = n_n; off = off_n; gcnt = gcnt_n;
    }
}

// The extras (band outliers and slab-exchange arrivals).  Step 1: every extra takes a slot of its (cell, class) region of the
// OUTPUT layout and leaves its original position there (slot[] is indexed by output position; the order inside a region is
// whatever the atomics produce).  Step 2 (k_extra_regions) puts every region into ascending original position -- the reference's
// stable order -- and moves the records.
static __global__ void __launch_bounds__(256) k_extra_slots(ExBufs E, const int64_t* __restrict__ start, const int32_t* __restrict__ hist,
                                                           int32_t* __restrict__ slot, const int* flags) {
    if (flags[2] != 0) return;
    const int n = min(*E.n, E.cap);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int cc = E.cell[e], c = cc >> 1, cls = cc & 1;
        const int64_t rs = cls ? start[c] + hist[c] - E.cntA[c] : start[c];
        const int k = atomicAdd((cls ? E.curA : E.curB) + c, 1);
        slot[rs + k] = E.idx[e];
    }
}
// Step 2, a warp per region (32 consecutive cells per round, one coalesced read of their counters; almost all are empty): the
// region's original positions are put into ascending order in shared memory (the warp-level bitonic network of the general path)
// and the records are copied to their final places.
static __global__ void __launch_bounds__(256) k_extra_regions(ExBufs E, const int64_t* __restrict__ start, const int32_t* __restrict__ hist,
                                                             const int32_t* __restrict__ slot, int64_t n_cells, SoA in_, SoA out_,
                                                             const int* flags) {
    if (flags[2] != 0) return;
    __shared__ int32_t shw[8][EX_REGION_MAX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t* sh = shw[wid];
    const int64_t gw = (int64_t)blockIdx.x * 8 + wid, nwarps = (int64_t)gridDim.x * 8;
    for (int64_t c0 = gw * 32; c0 < n_cells; c0 += nwarps * 32) {
        const int64_t c = c0 + lane;
        int mB = 0, mA = 0;
        int64_t rB = 0, rA = 0;
        if (c < n_cells) {
            mB = E.cntB[c]; mA = E.cntA[c];
            if (mB | mA) { rB = start[c]; rA = rB + hist[c] - mA; }
        }
        for (int cls = 0; cls < 2; cls++) {
            unsigned todo = __ballot_sync(0xffffffffu, (cls ? mA : mB) > 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t rs = __shfl_sync(0xffffffffu, cls ? rA : rB, src);
                const int n = __shfl_sync(0xffffffffu, cls ? mA : mB, src);
                __syncwarp();
                for (int i = lane; i < n; i += 32) sh[i] = slot[rs + i];
                __syncwarp();
                bool uns = false;
                for (int i = lane; i + 1 < n; i += 32) uns |= sh[i] > sh[i + 1];
                if (__any_sync(0xffffffffu, uns)) {
                    int m = 1;
                    while (m < n) m <<= 1;
                    for (int k = 2; k <= m; k <<= 1) {
                        const int hk = k >> 1;
                        for (int t = lane; t < (m >> 1); t += 32) {
                            const int blk = t / hk, off = t - blk * hk;
                            const int i = blk * k + off, p = blk * k + k - 1 - off;
                            if (p < n) {
                                const int32_t x = sh[i], y = sh[p];
                                if (x > y) { sh[i] = y; sh[p] = x; }
                            }
                        }
                        __syncwarp();
                        for (int j = k >> 2; j > 0; j >>= 1) {
                            for (int t = lane; t < (m >> 1); t += 32) {
                                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                                const int p = i + j;
                                if (p < n) {
                                    const int32_t x = sh[i], y = sh[p];
                                    if (x > y) { sh[i] = y; sh[p] = x; }
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
                for (int t = lane; t < n; t += 32) {
                    const int64_t i = sh[t], pos = rs + t;
#pragma unroll
                    for (int f = 0; f < 7; f++) out_.a[f][pos] = in_.a[f][i];
                }
            }
        }
    }
}

// Moments of the freshly sorted cells: the staying group's sums from P, plus the movers and arrivals of the cell read back
// from the output (they sit in known sub-ranges of the cell, ~5 % of it), in source order -- deterministic.
template <int W>
__global__ void __launch_bounds__(128) k_band_combine(const double* __restrict__ P, const int32_t* __restrict__ M, const int32_t* __restrict__ cntB,
                                                      const int32_t* __restrict__ cntA, const int64_t* __restrict__ seg_lo,
                                                      const int32_t* __restrict__ seg_n, const int64_t* __restrict__ start, SoA in_, SoA out_,
                                                      int64_t n_cells, double* __restrict__ pcache, const int* flags) {
    if (flags[2] != 0) return;
    constexpr int w = W / 2;
    const double* __restrict__ o0 = out_.a[0]; const double* __restrict__ o1 = out_.a[1]; const double* __restrict__ o2 = out_.a[2];
    const double* __restrict__ o3 = out_.a[3];
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        double K1 = 0, K2 = 0, K3 = 0;
        if (seg_n[c] > 0) { const int64_t f = seg_lo[c]; K1 = in_.a[1][f]; K2 = in_.a[2][f]; K3 = in_.a[3][f]; }
        const double* q = P + c * 5;
        double an = q[0], ax = q[1], ay = q[2], az = q[3], aq = q[4];
        int64_t pos = start[c];
        const int64_t pos0 = pos;
#pragma unroll
        for (int k = -1; k <= W; k++) {  // k == -1: the extras before the band, k == W: the extras after it (incl. arrivals)
            int cnt;
            if (k < 0) cnt = cntB[c];
            else if (k < W) {
                const int64_t cs = c - w + k;
                if (cs < 0 || cs >= n_cells) continue;
                cnt = M[cs * W + (W - 1 - k)];
            } else {
                cnt = cntA[c];
            }
            if (k != w) {
                for (int t = 0; t < cnt; t++) {
                    const double pw = o0[pos + t], cx = o1[pos + t] - K1, cy = o2[pos + t] - K2, cz = o3[pos + t] - K3;
                    an += pw; ax += pw * cx; ay += pw * cy; az += pw * cz;
                    aq += pw * (cx * cx + cy * cy + cz * cz);
                }
            }
            pos += cnt;
        }
        double* pc = pcache + 6 * c;
        pc[0] = (double)(pos - pos0);
        if (an > 0.0) {
            const double mx = ax / an, my = ay / an, mz = az / an;  // mean of (v - K)
            pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
            pc[5] = aq - an * (mx * mx + my * my + mz * mz);        // sum w |v - vbar|^2
        } else {
            pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
        }
    }
}

// The same, a WARP per cell: the movers / arrivals / extras of a cell (everything but the stayers' run) are summed lane-strided instead
// of by one thread in a dependent chain of loads (0.13 -> 0.05 ms at 125 000 cells of 1000).  Lane k + 1 looks up the size of group k.
template <int W>
__global__ void __launch_bounds__(256) k_band_combine_warp(const double* __restrict__ P, const int32_t* __restrict__ M, const int32_t* __restrict__ cntB,
                                                           const int32_t* __restrict__ cntA, const int64_t* __restrict__ seg_lo,
                                                           const int32_t* __restrict__ seg_n, const int64_t* __restrict__ start, SoA in_, SoA out_,
                                                           int64_t n_cells, double* __restrict__ pcache, const int* flags) {
    if (flags[2] != 0) return;
    constexpr int w = W / 2;
    static_assert(W + 2 <= 32, "one lane per group");
    const int lane = threadIdx.x & 31;
    const double* __restrict__ o0 = out_.a[0]; const double* __restrict__ o1 = out_.a[1]; const double* __restrict__ o2 = out_.a[2];
    const double* __restrict__ o3 = out_.a[3];
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        double K1 = 0, K2 = 0, K3 = 0;
        if (seg_n[c] > 0) { const int64_t f = seg_lo[c]; K1 = in_.a[1][f]; K2 = in_.a[2][f]; K3 = in_.a[3][f]; }
        // group k = lane - 1: k == -1 the extras before the band, 0 .. W - 1 the band groups by source cell, k == W the extras after
        int cnt = 0;
        const int k = lane - 1;
        if (k == -1) cnt = cntB[c];
        else if (k < W) {
            const int64_t cs = c - w + k;
            if (cs >= 0 && cs < n_cells) cnt = M[cs * W + (W - 1 - k)];
        } else if (k == W) cnt = cntA[c];
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int stay_lo = __shfl_sync(0xffffffffu, incl - cnt, w + 1), stay_n = __shfl_sync(0xffffffffu, cnt, w + 1);
        const int64_t pos0 = start[c];
        const int n_mov = total - stay_n;
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;
        for (int i = lane; i < n_mov; i += 32) {
            const int64_t p = pos0 + (i < stay_lo ? i : i + stay_n);
            const double pw = o0[p], cx = o1[p] - K1, cy = o2[p] - K2, cz = o3[p] - K3;
            an += pw; ax += pw * cx; ay += pw * cy; az += pw * cz;
            aq += pw * (cx * cx + cy * cy + cz * cz);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            an += __shfl_xor_sync(0xffffffffu, an, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
            aq += __shfl_xor_sync(0xffffffffu, aq, o);
        }
        if (lane == 0) {
            const double* q = P + c * 5;
            an += q[0]; ax += q[1]; ay += q[2]; az += q[3]; aq += q[4];
            double* pc = pcache + 6 * c;
            pc[0] = (double)total;
            if (an > 0.0) {
                const double mx = ax / an, my = ay / an, mz = az / an;  // mean of (v - K)
                pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
                pc[5] = aq - an * (mx * mx + my * my + mz * mz);        // sum w |v - vbar|^2
            } else {
                pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused convect + classify
// convect_particles! on a sorted layout knows everything pass A of the band sort needs: it holds x_new of every particle of
// old cell c' in registers.  This kernel is k_convect_contiguous and k_band_classify in one pass over HBM (x, vx read through a
// cp.async ring; x and dr written).  The following sort_particles! finds the classification cached
// (ctx->cls_gen == ctx->state_gen) and starts at the scan.
// Device flags: F_OUTSIDE = a particle left the slab (legal only if a slab exchange follows), F_CLS_BAD = band overflow
// (the sort takes the general path), F_FAR = a particle left the slab from a cell further than w from that edge,
// F_CLS_REDO = a cell could not be classified here: the sort runs k_band_classify after all.
#ifndef MB_CB_MINB
#define MB_CB_MINB 2  // resident CTAs per SM the fused kernel is compiled for (register budget)
#endif
constexpr int CB_PF = 8;  // batches of 32 particles in flight per warp (cp.async ring)
constexpr int CB_SMEM_WARP = 2 * CB_PF * 32 * 8 + 32 * 4;
constexpr int CB_SMEM = 8 * CB_SMEM_WARP;

template <int W>
__global__ void __launch_bounds__(256, MB_CB_MINB) k_convect_band(ConvectArgs a, int32_t* __restrict__ M, int64_t* __restrict__ seg_lo,
                                                                  int32_t* __restrict__ seg_n, uint32_t* __restrict__ dr, ExBufs E, int* flags) {
    constexpr int w = W / 2;
    extern __shared__ __align__(16) unsigned char cb_smem[];  // CB_SMEM bytes: per warp rx | rv | cnt
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned char* wbase = cb_smem + (size_t)wid * CB_SMEM_WARP;
    double* rx = (double*)wbase;
    double* rv = rx + CB_PF * 32;
    int* cnt_s = (int*)(rv + CB_PF * 32);
    const double dt = a.dt, L = a.L, inv_dx = a.inv_dx, min_x = a.min_x, max_x = a.max_x;
    const int cell_offset = (int)a.cell_offset, n_cells = (int)a.n_cells;  // the launcher checks that the global cell count fits 31 bits
    const int compute_cell = a.compute_cell;
    for (int64_t c64 = warp0; c64 < a.n_cells; c64 += nwarps) {
        const int c = (int)c64;
        const Indexer q = a.ix[c];
        const int64_t lo = q.start1 - 1;
        const int n = (int)q.n_group1;
        if (lane == 0) { seg_lo[c] = lo; seg_n[c] = n; }
        if (q.n_group1 >= (1 << 24) || q.n_group2 != 0) {  // not a sorted layout after all: just move the particles
            for (int64_t j = lo + lane; j < q.end1; j += 32) convect_one(a, j);
            if (q.n_group2 > 0)
                for (int64_t j = q.start2 - 1 + lane; j < q.end2; j += 32) convect_one(a, j);
            if (lane == 0) atomicOr(&flags[F_CLS_REDO], 1);
            continue;
        }
        double* __restrict__ Xc = a.pv.a[F_X] + lo;
        const double* __restrict__ VXc = a.pv.a[F_VX] + lo;
        uint32_t* __restrict__ drc = dr + lo;
        __syncwarp();
        cnt_s[lane] = 0;
        __syncwarp();
        bool bad = false, outside = false, far = false, any_ex = false;
        // one particle: returns d (255: not counted; any_ex: an extra -- inside the slab but further than w cells away -- was seen)
        auto move = [&](int j, double x_old, double vx) -> int {
            double x_new = fma(vx, dt, x_old);  // @muladd x[1] + v[1] * dt
            if (x_new >= L || x_new <= 0.0) x_new = convect_wall(a, lo + j, x_old, vx, x_new);
            if (x_new < min_x) x_new = min_x;
            else if (x_new > max_x) x_new = max_x;
            Xc[j] = x_new;
            const int nc = __double2int_rd(x_new * inv_dx) - cell_offset;  // == cell_of(): 0 <= x * inv_dx < 2^31
            if (compute_cell) a.cell[lo + j] = nc + 1;
            const int d = nc - c + w;
            if (nc >= 0 && nc < n_cells) {
                if (d >= 0 && d < W) return d;
                any_ex = true;
                return 255;
            }
            outside = true;
            if ((nc < 0 && c >= w) || (nc >= n_cells && c < n_cells - w)) far = true;
            return 255;
        };
        // x and vx stream through a per-warp ring of CB_PF batches filled with cp.async (each lane copies and later reads its
        // own 8 bytes, so no warp barrier is needed): CB_PF * 512 B per warp in flight hide the HBM latency.
        const int nb = (n + 31) >> 5;
        auto issue = [&](int k) {
            const int j = (k << 5) + lane;
            if (k < nb && j < n) {
                const int slot = k & (CB_PF - 1);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(&rx[slot * 32 + lane])), "l"(Xc + j));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(&rv[slot * 32 + lane])), "l"(VXc + j));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int k = 0; k < CB_PF; k++) issue(k);
        for (int k = 0; k < nb; k++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(CB_PF - 1) : "memory");
            const int j = (k << 5) + lane;
            const bool valid = j < n;
            const int slot = k & (CB_PF - 1);
            const double x0 = rx[slot * 32 + lane], v0 = rv[slot * 32 + lane];
            issue(k + CB_PF);
            int d = 255;
            if (valid) d = move(j, x0, v0);
            classify_batch(d, valid, lt, cnt_s, drc + (valid ? j : 0));
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // extras are rare on the band path: the streaming loop only notes that the cell has some, and they are listed here from the
        // positions just written (the cell is still in L1 / L2)
        if (__any_sync(0xffffffffu, any_ex)) {
            for (int k = 0; k < nb; k++) {
                const int j = (k << 5) + lane;
                bool ex = false;
                int nc = 0;
                if (j < n) {
                    nc = __double2int_rd(Xc[j] * inv_dx) - cell_offset;
                    const int d = nc - c + w;
                    ex = nc >= 0 && nc < n_cells && (d < 0 || d >= W);
                }
                if (extras_append(E, ex, (int32_t)(lo + j), nc, nc < c ? 1 : 0, lane, lt)) bad = true;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        outside = __any_sync(0xffffffffu, outside);
        far = __any_sync(0xffffffffu, far);
        if (lane == 0) {
            if (bad) atomicOr(&flags[F_CLS_BAD], 1);
            if (outside) atomicOr(&flags[F_OUTSIDE], 1);
            if (far) atomicOr(&flags[F_FAR], 1);
        }
        if (bad) continue;
        if (lane < W) M[(int64_t)c * W + lane] = cnt_s[lane];
    }
}

__global__ void k_clear_cls_flags(int* flags) { flags[F_OUTSIDE] = 0; flags[F_CLS_BAD] = 0; flags[F_FAR] = 0; flags[F_CLS_REDO] = 0; }
// number of extras of the last band sort (diagnostic: mb_sort_last_extras)
__global__ void k_save_extras(const int32_t* n, int* flags) { flags[3] = *n; }
// flags[2] (general path needed) from the cached classification: band overflow, or a particle outside the slab with no exchange
__global__ void k_flag_from_cls(int* flags, int drop) { flags[2] = (flags[F_CLS_BAD] != 0 || (!drop && flags[F_OUTSIDE] != 0)) ? 1 : 0; }

// ------------------------------------------------------------------------------------------------ general path
// src (nullable): physical position of logical (squashed) position i -- the squash of a non-contiguous species is folded into the
// sort instead of moving the payload twice (see k_build_src)
__global__ void __launch_bounds__(256) k_gen_classify(const double* __restrict__ X, const int32_t* cell_in, int32_t* cell_out,
                                                      int32_t* __restrict__ key, const int64_t* n_total_p, int64_t n_cells, double inv_dx,
                                                      int64_t cell_offset, int use_x, int drop, int32_t* __restrict__ hist, int* flags,
                                                      const int32_t* __restrict__ src, unsigned long long* __restrict__ n_dropped) {
    // drop != 0 (after a slab exchange): a particle whose cell is outside [0, n_cells) left the slab and is dropped
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
        bool counted = false;
        if (valid) {
            const int64_t ph = src ? (int64_t)src[i] : i;
            if (use_x) { nc = cell_of(X[ph], inv_dx, cell_offset); cell_out[i] = nc + 1; }
            else nc = cell_in[ph] - 1;
            if (nc < 0 || nc >= n_cells) {
                if (drop) nc = -1;
                else { atomicOr(&flags[0], DEVERR_BAD_CELL); nc = nc < 0 ? 0 : (int)(n_cells - 1); }
            }
            key[i] = nc;
            counted = nc >= 0;
        }
        if (n_dropped != nullptr) {  // after an edge exchange: the sort must drop exactly the particles that were sent
            const unsigned dm = __ballot_sync(0xffffffffu, valid && nc < 0);
            if (dm != 0 && lane == __ffs(dm) - 1) atomicAdd(n_dropped, (unsigned long long)__popc(dm));
        }
        // run-length aggregation: keys of neighbouring lanes are mostly equal
        const unsigned act = __ballot_sync(0xffffffffu, counted);
        if (counted) {
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
    // every warp walks a CONTIGUOUS chunk of 32 x SC_ROUNDS positions in ascending order, so that the indices of a cell whose
    // particles are contiguous in the input (the usual case: a nearly sorted layout) are appended by one warp in ascending order
    // and k_gen_sort_segments* finds most cells already sorted (it stays the safety net: correctness never depends on the order
    // in which the atomics land)
    constexpr int SC_ROUNDS = 32;
    const int64_t warp_g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base0 = warp_g * 32 * SC_ROUNDS; base0 < n_total; base0 += nwarps * 32 * SC_ROUNDS) {
        for (int k = 0; k < SC_ROUNDS; k++) {
            const int64_t i = base0 + k * 32 + lane;
            const int nc = i < n_total ? key[i] : -1;
            const bool valid = nc >= 0;
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (act == 0) continue;
            if (valid) {
                const unsigned peers = __match_any_sync(act, nc);
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&cursor[nc], __popc(peers));
                base = __shfl_sync(peers, base, leader);
                perm[start[nc] + base + __popc(peers & lt)] = (int32_t)i;
            }
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
// Segments of up to WSEG indices (every cell of the Couette / Fokker-Planck shapes): one WARP per cell, the segment staged in the
// warp's slice of shared memory, the same normalised bitonic network with __syncwarp only.  A warp takes 32 consecutive cells at
// a time (one coalesced read of their bounds) and skips the cells that are already ascending.
constexpr int WSEG = 1024;
__global__ void __launch_bounds__(256) k_gen_sort_segments_warp(int32_t* __restrict__ perm, const int64_t* __restrict__ start, int64_t n_cells,
                                                                const int* flags) {
    if (flags[2] == 0) return;
    __shared__ int32_t shw[8][WSEG];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t* sh = shw[wid];
    const int64_t gw = (int64_t)blockIdx.x * 8 + wid, nwarps = (int64_t)gridDim.x * 8;
    for (int64_t c0 = gw * 32; c0 < n_cells; c0 += nwarps * 32) {
        const int64_t c = c0 + lane;
        const int64_t lo = c < n_cells ? start[c] : 0;
        const int64_t nn = c < n_cells ? start[c + 1] - lo : 0;
        unsigned todo = __ballot_sync(0xffffffffu, nn >= 2 && nn <= WSEG);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t LO = __shfl_sync(0xffffffffu, lo, src);
            const int n = (int)__shfl_sync(0xffffffffu, nn, src);
            int32_t* seg = perm + LO;
            __syncwarp();
            for (int i = lane; i < n; i += 32) sh[i] = seg[i];
            __syncwarp();
            bool uns = false;
            for (int i = lane; i + 1 < n; i += 32) uns |= sh[i] > sh[i + 1];
            if (!__any_sync(0xffffffffu, uns)) continue;
            int m = 1;
            while (m < n) m <<= 1;
            for (int k = 2; k <= m; k <<= 1) {
                const int hk = k >> 1;
                for (int t = lane; t < (m >> 1); t += 32) {
                    const int blk = t / hk, off = t - blk * hk;
                    const int i = blk * k + off, p = blk * k + k - 1 - off;
                    if (p < n) {
                        const int32_t x = sh[i], y = sh[p];
                        if (x > y) { sh[i] = y; sh[p] = x; }
                    }
                }
                __syncwarp();
                for (int j = k >> 2; j > 0; j >>= 1) {
                    for (int t = lane; t < (m >> 1); t += 32) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int p = i + j;
                        if (p < n) {
                            const int32_t x = sh[i], y = sh[p];
                            if (x > y) { sh[i] = y; sh[p] = x; }
                        }
                    }
                    __syncwarp();
                }
            }
            for (int i = lane; i < n; i += 32) seg[i] = sh[i];
        }
    }
}
// larger segments: one CTA per cell
__global__ void __launch_bounds__(256) k_gen_sort_segments(int32_t* __restrict__ perm, const int64_t* __restrict__ start, int64_t n_cells,
                                                           const int* flags, int cpb) {
    if (flags[2] == 0) return;
    __shared__ int32_t sh[SEG_SMEM];
    __shared__ int unsorted;
    __shared__ int s_list[256], s_nlist;
    // the CTA looks at cpb (<= 256) cells at a time (one coalesced read of their bounds) and sorts the large ones among them one by one
    for (int64_t cbase = (int64_t)blockIdx.x * cpb; cbase < n_cells; cbase += (int64_t)gridDim.x * cpb) {
      __syncthreads();
      if (threadIdx.x == 0) s_nlist = 0;
      __syncthreads();
      {
          const int64_t cc = cbase + threadIdx.x;
          if ((int)threadIdx.x < cpb && cc < n_cells && start[cc + 1] - start[cc] > WSEG) s_list[atomicAdd(&s_nlist, 1)] = threadIdx.x;
      }
      __syncthreads();
      const int nlist = s_nlist;
      for (int li = 0; li < nlist; li++) {
        // the order of the list does not matter: every listed cell is sorted independently
        const int64_t c = cbase + s_list[li];
        const int64_t lo = start[c];
        const int64_t n = start[c + 1] - lo;
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
}

// The slice perm[0 .. n) of one cell (n <= WSEG) staged in the warp's shared-memory buffer in ASCENDING order: a check when the atomics
// happened to land in order (nearly sorted input), else the normalised bitonic network of k_gen_sort_segments_warp.  Used by the gathers
// by destination cell, which makes a separate sorting pass over the permutation (and its write-back) unnecessary for small cells.
__device__ __forceinline__ void warp_stage_sorted(const int32_t* __restrict__ seg, int n, int32_t* sh, int lane) {
    __syncwarp();
    for (int i = lane; i < n; i += 32) sh[i] = seg[i];
    __syncwarp();
    bool uns = false;
    for (int i = lane; i + 1 < n; i += 32) uns |= sh[i] > sh[i + 1];
    if (!__any_sync(0xffffffffu, uns)) return;
    int m = 1;
    while (m < n) m <<= 1;
    for (int k = 2; k <= m; k <<= 1) {
        const int hk = k >> 1;
        for (int t = lane; t < (m >> 1); t += 32) {
            const int blk = t / hk, off = t - blk * hk;
            const int i = blk * k + off, p = blk * k + k - 1 - off;
            if (p < n) {
                const int32_t x = sh[i], y = sh[p];
                if (x > y) { sh[i] = y; sh[p] = x; }
            }
        }
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = lane; t < (m >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i + j;
                if (p < n) {
                    const int32_t x = sh[i], y = sh[p];
                    if (x > y) { sh[i] = y; sh[p] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// cell_out (nullable; the variant that sorts by stored cell ids, grid_sorting.jl:128): the ids are permuted along with the particles,
// so that an ensemble of 0-D cells can be re-sorted step after step (device-side extension: the reference leaves particles.cell stale)
__global__ void __launch_bounds__(256) k_gen_gather(SoA in, SoA out, const int32_t* __restrict__ perm, const int64_t* n_total_p, const int* flags,
                                                    const int32_t* __restrict__ src, const int32_t* __restrict__ key, int32_t* __restrict__ cell_out) {
    if (flags[2] == 0) return;
    const int64_t n_total = *n_total_p;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_total; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[j];
        const int64_t ph = src ? (int64_t)src[i] : i;
#pragma unroll
        for (int f = 0; f < 7; f++) out.a[f][j] = in.a[f][ph];
        if (cell_out) cell_out[j] = key[i] + 1;
    }
}

// The gather by destination cell: a warp per cell reads the cell's slice of the permutation, gathers the payload, writes the cell as one
// coalesced run and accumulates the cell's moments on the way (shifted by the velocity of the cell's first particle, exactly like
// pass B of the band path), so that compute_props_sorted! right after a general-path sort costs no particle traffic either.
// Used when the cells are small (the Couette shapes); 0-D cells of thousands of particles take k_gen_gather.
__global__ void __launch_bounds__(256) k_gen_gather_cells(SoA in, SoA out, const int32_t* __restrict__ perm, const int64_t* __restrict__ start,
                                                          int64_t n_cells, const int* flags, const int32_t* __restrict__ src,
                                                          const int32_t* __restrict__ key, int32_t* __restrict__ cell_out, double* __restrict__ pcache) {
    if (flags[2] == 0) return;
    __shared__ int32_t shw[8][WSEG];
    const int lane = threadIdx.x & 31;
    int32_t* sh = shw[threadIdx.x >> 5];
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const int64_t lo = start[c];
        const int n = (int)(start[c + 1] - lo);
        const bool staged = n <= WSEG;  // larger cells were put in order by k_gen_sort_segments
        if (staged) warp_stage_sorted(perm + lo, n, sh, lane);
        double K1 = 0, K2 = 0, K3 = 0;
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            if (j < n) {
                const int64_t i = staged ? sh[j] : perm[lo + j];
                const int64_t ph = src ? (int64_t)src[i] : i;
                a0 = in.a[0][ph]; a1 = in.a[1][ph]; a2 = in.a[2][ph]; a3 = in.a[3][ph];
                const double a4 = in.a[4][ph], a5 = in.a[5][ph], a6 = in.a[6][ph];
                out.a[0][lo + j] = a0; out.a[1][lo + j] = a1; out.a[2][lo + j] = a2; out.a[3][lo + j] = a3;
                out.a[4][lo + j] = a4; out.a[5][lo + j] = a5; out.a[6][lo + j] = a6;
                if (cell_out) cell_out[lo + j] = key[i] + 1;
            }
            if (j0 == 0) {  // shift = velocity of the cell's first particle
                K1 = __shfl_sync(0xffffffffu, a1, 0); K2 = __shfl_sync(0xffffffffu, a2, 0); K3 = __shfl_sync(0xffffffffu, a3, 0);
            }
            if (j < n) {
                const double cx = a1 - K1, cy = a2 - K2, cz = a3 - K3;
                an += a0; ax += a0 * cx; ay += a0 * cy; az += a0 * cz;
                aq += a0 * (cx * cx + cy * cy + cz * cz);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            an += __shfl_xor_sync(0xffffffffu, an, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
            aq += __shfl_xor_sync(0xffffffffu, aq, o);
        }
        if (lane == 0) {
            double* pc = pcache + 6 * c;
            pc[0] = (double)n;
            if (an > 0.0) {
                const double mx = ax / an, my = ay / an, mz = az / an;  // mean of (v - K)
                pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
                pc[5] = aq - an * (mx * mx + my * my + mz * mz);        // sum w |v - vbar|^2
            } else {
                pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
            }
        }
    }
}

// When nothing is left of the previous order (band path switched off: a particle crosses hundreds of cells per step, the reference's
// "same L, finer cells" scaling) the gather above reads seven scattered 8-byte values per particle, and every one of them costs a DRAM
// access of its own (measured: 13.6 ms of an 18.7 ms sort at 1.25e8 particles).  Two coalescing-friendly passes instead: the records
// are first packed into 64-byte array-of-structs entries (one aligned DRAM access each), then gathered by destination cell.
struct __align__(16) Rec64 { double2 a, b, c, d; };  // w vx | vy vz | x y | z -
__global__ void __launch_bounds__(256) k_gen_pack_aos(SoA in, const int64_t* n_total_p, const int32_t* __restrict__ src, Rec64* __restrict__ rec,
                                                      const int* flags) {
    if (flags[2] == 0) return;
    // four threads per record, one 16-byte quarter each: a warp reads 8 consecutive values of every field (full sectors) and writes
    // 512 contiguous bytes
    const int64_t n4 = *n_total_p * 4;
    double2* __restrict__ out = (double2*)rec;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t >> 2;
        const int p = (int)(t & 3);
        const int64_t ph = src ? (int64_t)src[i] : i;
        const double v0 = in.a[2 * p][ph];
        const double v1 = p < 3 ? in.a[2 * p + 1][ph] : 0.0;
        out[t] = make_double2(v0, v1);
    }
}
__global__ void __launch_bounds__(256) k_gen_gather_cells_aos(const Rec64* __restrict__ rec, SoA out, const int32_t* __restrict__ perm,
                                                              const int64_t* __restrict__ start, int64_t n_cells, const int* flags,
                                                              const int32_t* __restrict__ key, int32_t* __restrict__ cell_out,
                                                              double* __restrict__ pcache) {
    if (flags[2] == 0) return;
    __shared__ int32_t shw[8][WSEG];
    const int lane = threadIdx.x & 31;
    int32_t* sh = shw[threadIdx.x >> 5];
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const int64_t lo = start[c];
        const int n = (int)(start[c + 1] - lo);
        const bool staged = n <= WSEG;  // larger cells were put in order by k_gen_sort_segments
        if (staged) warp_stage_sorted(perm + lo, n, sh, lane);
        double K1 = 0, K2 = 0, K3 = 0;
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            if (j < n) {
                const int64_t i = staged ? sh[j] : perm[lo + j];
                const Rec64 r = rec[i];
                a0 = r.a.x; a1 = r.a.y; a2 = r.b.x; a3 = r.b.y;
                out.a[0][lo + j] = a0; out.a[1][lo + j] = a1; out.a[2][lo + j] = a2; out.a[3][lo + j] = a3;
                out.a[4][lo + j] = r.c.x; out.a[5][lo + j] = r.c.y; out.a[6][lo + j] = r.d.x;
                if (cell_out) cell_out[lo + j] = key[i] + 1;
            }
            if (j0 == 0) {  // shift = velocity of the cell's first particle
                K1 = __shfl_sync(0xffffffffu, a1, 0); K2 = __shfl_sync(0xffffffffu, a2, 0); K3 = __shfl_sync(0xffffffffu, a3, 0);
            }
            if (j < n) {
                const double cx = a1 - K1, cy = a2 - K2, cz = a3 - K3;
                an += a0; ax += a0 * cx; ay += a0 * cy; az += a0 * cz;
                aq += a0 * (cx * cx + cy * cy + cz * cz);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            an += __shfl_xor_sync(0xffffffffu, an, o); ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
            aq += __shfl_xor_sync(0xffffffffu, aq, o);
        }
        if (lane == 0) {
            double* pc = pcache + 6 * c;
            pc[0] = (double)n;
            if (an > 0.0) {
                const double mx = ax / an, my = ay / an, mz = az / an;  // mean of (v - K)
                pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
                pc[5] = aq - an * (mx * mx + my * my + mz * mz);        // sum w |v - vbar|^2
            } else {
                pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
            }
        }
    }
}

// squash_pia! folded into the sort (grid_sorting.jl:69-71 squashes first): instead of moving the payload to close the holes and then
// moving it again in the sort, only the map logical (squashed) position -> physical position is built (4 B per particle).  The
// squashed order walks group 1 of all cells, then group 2 of all cells (particles.jl:622-682); newlo = exclusive scan of the segment
// sizes.  Load balance over segment sizes: mb_segcopy.cuh.
__global__ void k_set_flag(int* flags, int idx, int v) { flags[idx] = v; }
// edge exchange + general-path sort: a leaver from a cell further than w from its slab face was not sent, the sort would drop it
// silently -- the counts of sent and dropped particles must agree
__global__ void k_check_dropped(const int64_t* xc, int* flags) {
    if (flags[2] != 0 && xc[10] != xc[8] + xc[9]) atomicOr(&flags[0], DEVERR_BAND_OVERFLOW);
}

// ------------------------------------------------------------------------------------------------ segment path
// sort_particles!(gridsort, pv, pia, species) by stored cell ids (grid_sorting.jl:128-182) when NO particle changed its cell -- the
// re-sort of an ensemble of 0-D cells after variable-weight collisions and merging: the split particles sit in group-2 ranges at the
// tail, merges left holes.  The stable counting sort of the squashed order (group 1 of all cells, then group 2 of all cells) then
// is, for every cell, group 1 followed by group 2: a concatenation of the pia's segments, no keys, no ranking.  One pass checks
// the precondition (4 B per particle), one moves the payload (112 B); a mismatch hands over to the general path.
static __global__ void k_segpath_counts(const Indexer* __restrict__ ix, int64_t nc, int32_t* __restrict__ cnt) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[c];
        cnt[2 * c] = (int32_t)q.n_group1;
        cnt[2 * c + 1] = (int32_t)q.n_group2;
    }
}
struct CellMajorDesc {  // segment 2 c + g = group g + 1 of cell c
    const Indexer* ix;
    const int64_t* newlo;
    const int* only_if_zero;  // nullable: no segments if *only_if_zero != 0
    __device__ __forceinline__ void get(int64_t seg, int64_t& n, int64_t& src, int64_t& dst) const {
        const Indexer q = ix[seg >> 1];
        const bool g2 = seg & 1;
        n = g2 ? q.n_group2 : q.n_group1;
        src = (g2 ? q.start2 : q.start1) - 1;
        dst = newlo[seg];
        if (only_if_zero != nullptr && *only_if_zero != 0) n = 0;
    }
};
struct KeyCheckAct {  // the stored cell id of every particle of a segment must be the segment's cell
    const int32_t* cell;
    int* flag;
    __device__ __forceinline__ void seg(int64_t, int64_t, int64_t, int64_t) const {}
    __device__ __forceinline__ void elem(int64_t seg, int64_t src, int64_t) const {
        if (cell[src] - 1 != (int32_t)(seg >> 1)) *flag = 1;
    }
};
struct PayloadMoveAct {
    SoA cur, alt;
    __device__ __forceinline__ void seg(int64_t, int64_t, int64_t, int64_t) const {}
    __device__ __forceinline__ void elem(int64_t, int64_t src, int64_t dst) const {
#pragma unroll
        for (int f = 0; f < 7; f++) alt.a[f][dst] = cur.a[f][src];
    }
};
// after the move: the new indexers (group 1 = the whole cell, grid_sorting.jl:156-166) and the cell ids of the new layout
static __global__ void __launch_bounds__(256) k_segpath_commit(Indexer* __restrict__ ix, int64_t nc, const int64_t* __restrict__ newlo,
                                                              int32_t* __restrict__ cell, const int* flags) {
    if (flags[2] != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < nc; c += nwarps) {
        const int64_t lo = newlo[2 * c], n = newlo[2 * c + 2] - lo;
        for (int64_t j = lane; j < n; j += 32) cell[lo + j] = (int32_t)c + 1;
        if (lane == 0) {
            Indexer q;
            q.n_local = n; q.n_group1 = n;
            q.start1 = n > 0 ? lo + 1 : 0;
            q.end1 = n > 0 ? lo + n : -1;
            q.start2 = 0; q.end2 = -1; q.n_group2 = 0;
            ix[c] = q;
        }
    }
}

struct BandBufs {
    int64_t* seg_lo;
    int32_t* seg_n;
    uint32_t* dr;      // [cap] d << 24 | rank inside the (source cell -> destination cell) group
    ExBufs E;          // band outliers and slab-exchange arrivals
    int64_t n_arr;     // host upper bound
    const int64_t* d_n_arr;  // device: exact
    int64_t* n_old;    // device copy of n_total before the sort
    double* pcache;    // nullable
    double* P;         // [n_cells][5] moment sums of the staying groups
    int drop;
    int64_t* d_nt_write;  // nullable: rewrite n_total (after a slab exchange)
};


// The scratch layout of a sort is a pure function of (capacity, n_cells, W), so the fused convect + classify kernel (which runs
// before the sort call) and the sort itself address the same buffers.
constexpr int W_MAX = 31;  // widest band (w = 15): the group id lives in a lane
static int sort_scratch_layout(mb_ctx* ctx, int64_t cap, int64_t nc, int W, SortScratch& S, BandBufs& B) {
    const int nscan = (int)((nc + SCAN_TILE - 1) / SCAN_TILE);
    S.flags = ctx->d_flags;
    S.key = (int32_t*)ctx_scratch(ctx, 0, (size_t)cap * 4 + 64);  // general: keys; band: dr (+ slack: the tile kernel's bulk loads are 16-byte granular)
    // slot 1 (int32): hist | cursor | seg_n | cntB | cntA | curB | curA | ex_n (64) | M; sized for the widest band so that W may change
    const size_t n32 = (size_t)nc * (7 + W_MAX) + 128;
    int32_t* p32 = (int32_t*)ctx_scratch(ctx, 1, n32 * 4);
    // slot 2: start | partial | n_old | seg_lo   (int64)
    const size_t n64 = (size_t)(nc + 1) + (size_t)(nscan + 2) + 2 + (size_t)nc + 64;
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 2, n64 * 8);
    // slot 12: the extras list (original position, destination cell)
    const int64_t ex_cap = cap / 4 + 65536;
    int32_t* pex = (int32_t*)ctx_scratch(ctx, 12, (size_t)ex_cap * 8);
    if (!S.key || !p32 || !p64 || !pex) return MB_ERR_CUDA;
    S.hist = p32;
    S.cursor = p32 + nc;
    S.M = p32 + 7 * nc + 64;
    S.start = p64;
    S.partial = p64 + (nc + 1);
    S.O = nullptr;
    S.perm = nullptr;
    B.seg_n = p32 + 2 * nc;
    B.E.cntB = p32 + 3 * nc;
    B.E.cntA = p32 + 4 * nc;
    B.E.curB = p32 + 5 * nc;
    B.E.curA = p32 + 6 * nc;
    B.E.n = p32 + 7 * nc;
    B.E.idx = pex;
    B.E.cell = pex + ex_cap;
    B.E.cap = (int32_t)(ex_cap < INT_MAX / 2 ? ex_cap : INT_MAX / 2);
    B.n_old = S.partial + (nscan + 2);
    B.seg_lo = B.n_old + 2;
    B.dr = (uint32_t*)S.key;
    B.n_arr = 0;
    B.d_n_arr = nullptr;
    B.drop = 0;
    B.pcache = nullptr;
    B.d_nt_write = nullptr;
    B.P = nullptr;
    (void)W;
    return MB_OK;
}
// cntB | cntA | curB | curA | ex_n are cleared before every classification
static int extras_clear(mb_ctx* ctx, const BandBufs& B, int64_t nc) {
    MB_CUDA(cudaMemsetAsync(B.E.cntB, 0, ((size_t)4 * nc + 64) * 4, ctx->stream));
    return MB_OK;
}

int convect_band_launch(mb_ctx* ctx, const ConvectArgs& a, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species, bool* done) {
    *done = false;
    const int s = (int)species - 1;
    const int w = ctx->band_w;
    if (w <= 0 || !pia->contiguous[s] || !pia->sorted_layout[s] || pv->n_arrivals != 0 || pv->drop_oob || grid->n_cells != pia->n_cells ||
        pia->n_cells >= (int64_t)INT_MAX || pv->cap >= (int64_t)INT_MAX || !(grid->L * grid->inv_dx < 2.0e9))
        return MB_OK;
    const int W = 2 * w + 1;
    const int64_t nc = pia->n_cells;
    SortScratch S;
    BandBufs B;
    if (sort_scratch_layout(ctx, pv->cap, nc, W, S, B)) return MB_ERR_CUDA;
    cudaStream_t st = ctx->stream;
    k_clear_cls_flags<<<1, 1, 0, st>>>(ctx->d_flags);
    MB_LAUNCH_CHECK(ctx);
    if (extras_clear(ctx, B, nc)) return MB_ERR_CUDA;
    const int g = grid_for(nc * 32, 256, MB_CB_MINB);
    static bool attr_done[64] = {false};  // function attributes are per device
    bool& attr_set = attr_done[ctx->device & 63];
    if (!attr_set) {
        MB_CUDA(cudaFuncSetAttribute(k_convect_band<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));
        MB_CUDA(cudaFuncSetAttribute(k_convect_band<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));
        MB_CUDA(cudaFuncSetAttribute(k_convect_band<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));
        MB_CUDA(cudaFuncSetAttribute(k_convect_band<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));
        MB_CUDA(cudaFuncSetAttribute(k_convect_band<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));
        attr_set = true;
    }
    if (w == 1) k_convect_band<3><<<g, 256, CB_SMEM, st>>>(a, S.M, B.seg_lo, B.seg_n, B.dr, B.E, ctx->d_flags);
    else if (w == 2) k_convect_band<5><<<g, 256, CB_SMEM, st>>>(a, S.M, B.seg_lo, B.seg_n, B.dr, B.E, ctx->d_flags);
    else if (w == 4) k_convect_band<9><<<g, 256, CB_SMEM, st>>>(a, S.M, B.seg_lo, B.seg_n, B.dr, B.E, ctx->d_flags);
    else if (w == 8) k_convect_band<17><<<g, 256, CB_SMEM, st>>>(a, S.M, B.seg_lo, B.seg_n, B.dr, B.E, ctx->d_flags);
    else k_convect_band<31><<<g, 256, CB_SMEM, st>>>(a, S.M, B.seg_lo, B.seg_n, B.dr, B.E, ctx->d_flags);
    MB_LAUNCH_CHECK(ctx);
    // the classification stays valid until something other than a slab exchange touches the particles
    ctx->cls_gen = ctx->state_gen;
    ctx->cls_pv = pv; ctx->cls_pia = pia; ctx->cls_species = (int)species; ctx->cls_w = w;
    ctx->cls_inv_dx = grid->inv_dx; ctx->cls_cell_offset = grid->cell_offset; ctx->cls_cap = pv->cap;
    *done = true;
    return MB_OK;
}


// ---- tile pass B (mb_sort_tile.cuh)
// MB_SORT_TILE: 0 never, 1 (default) for bands of w >= 4 when the mean cell population suits the tile size, 3 the same for every w,
// 2 always (tests of the direct mode)
static int tile_mode() {
    const char* e = getenv("MB_SORT_TILE");
    return e ? atoi(e) : 1;
}
// MB_TILE_CFG selects the compiled shape (tile particles / consumer threads / particles per stage x stages):
// 0 = 2048 / 256 / 1024 x 4, 1 = 4096 / 512 / 1024 x 8, 2 = 4096 / 512 / 2048 x 4 (default), 3 = 4096 / 512 / 4096 x 2, 4 = 4096 / 512 / 1024 x 9
static int tile_cfg() {
    const char* e = getenv("MB_TILE_CFG");
    return e ? atoi(e) : 2;
}
static int tile_ncap(int cfg) { return cfg == 0 ? 2048 : 4096; }
template <int NCAP, int NT, int SUB, int S, int MINB>
static int launch_tile_kernel(mb_ctx* ctx, const TileArgs& a, bool mom) {
    static bool attr_done[64] = {false};  // function attributes are per device
    if (!attr_done[ctx->device & 63]) {
        MB_CUDA(cudaFuncSetAttribute(k_band_tile<NCAP, NT, SUB, S, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileSmem<NCAP, SUB, S, true>::BYTES));
        MB_CUDA(cudaFuncSetAttribute(k_band_tile<NCAP, NT, SUB, S, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileSmem<NCAP, SUB, S, false>::BYTES));
        attr_done[ctx->device & 63] = true;
    }
    const int g = N_SM * MINB;
    if (mom) k_band_tile<NCAP, NT, SUB, S, true, MINB><<<g, NT + 64, TileSmem<NCAP, SUB, S, true>::BYTES, ctx->stream>>>(a);
    else k_band_tile<NCAP, NT, SUB, S, false, MINB><<<g, NT + 64, TileSmem<NCAP, SUB, S, false>::BYTES, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}
struct TileBufs {
    int64_t* old_start;    // [nc + 1]
    int64_t* partial;
    int64_t* tp;           // NI, NK, largest cell, n_old
    int32_t* chunk_first;  // [NK + 2]
    double* Pp;            // partial moments
};
static int tile_scratch(mb_ctx* ctx, int64_t cap, int64_t nc, int w, bool mom, TileBufs& T) {
    const size_t n64 = (size_t)(nc + 1) + gs_partial_count(nc) + 8;
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 16, n64 * 8);
    const int64_t nk_max = cap / 512 + 4;  // NI >= NCAP / 4 >= 512
    T.chunk_first = (int32_t*)ctx_scratch(ctx, 17, (size_t)(nk_max + 4) * 4);
    if (!p64 || !T.chunk_first) return MB_ERR_CUDA;
    T.old_start = p64;
    T.partial = p64 + (nc + 1);
    T.tp = T.partial + gs_partial_count(nc);
    T.Pp = nullptr;
    if (mom) {
        T.Pp = (double*)ctx_scratch(ctx, 18, (size_t)(2 * w * (nk_max + 1) + nc + 2 * w + 4) * 5 * 8);
        if (!T.Pp) return MB_ERR_CUDA;
    }
    return MB_OK;
}
// old_start / tiles from the old cell sizes, then the tile kernel; moments: combine (+ the fallback if a tile went the direct way)
static int launch_tile_pass_b(mb_ctx* ctx, mb_pv* pv, int64_t nc, int W, const SortScratch& S, const BandBufs& B, bool mom) {
    cudaStream_t st = ctx->stream;
    const int cfg = tile_cfg();
    const int ncap = tile_ncap(cfg);
    TileBufs T;
    if (tile_scratch(ctx, pv->cap, nc, W / 2, mom, T)) return MB_ERR_CUDA;
    MB_CUDA(cudaMemsetAsync(T.tp, 0, 8 * 8, st));
    if (mom) MB_CUDA(cudaMemsetAsync(ctx->d_flags + F_MOM_BAD, 0, sizeof(int), st));
    const int nb = (int)((nc + GS_TILE - 1) / GS_TILE);
    k_tile_reduce<<<nb, GS_BLOCK, 0, st>>>(B.seg_n, nc, T.partial, T.tp, S.flags);
    MB_LAUNCH_CHECK(ctx);
    k_tile_partials<<<1, 1024, 0, st>>>(T.partial, nb, T.tp, ncap, S.flags);
    MB_LAUNCH_CHECK(ctx);
    k_tile_apply<<<nb, GS_BLOCK, 0, st>>>(B.seg_n, nc, T.partial, T.old_start, T.chunk_first, T.tp, S.flags);
    MB_LAUNCH_CHECK(ctx);
    TileArgs a;
    a.in = pv->cur; a.out = pv->alt;
    a.dr = B.dr; a.M = S.M; a.old_start = T.old_start; a.chunk_first = T.chunk_first; a.tp = T.tp;
    a.start = S.start; a.cntB = B.E.cntB; a.n_cells = nc; a.W = W; a.Pp = T.Pp; a.flags = S.flags;
    a.debug = getenv("MB_TILE_DEBUG") != nullptr;
    int r;
    if (cfg == 0) r = launch_tile_kernel<2048, 256, 1024, 4, 1>(ctx, a, mom);
    else if (cfg == 2) r = launch_tile_kernel<4096, 512, 2048, 4, 1>(ctx, a, mom);
    else if (cfg == 3) r = launch_tile_kernel<4096, 512, 4096, 2, 1>(ctx, a, mom);
    else if (cfg == 4) r = launch_tile_kernel<4096, 512, 1024, 9, 1>(ctx, a, mom);
    else r = launch_tile_kernel<4096, 512, 1024, 8, 1>(ctx, a, mom);
    if (r) return r;
    return MB_OK;
}
static int launch_tile_moments(mb_ctx* ctx, mb_pv* pv, int64_t nc, int W, const SortScratch& S, const BandBufs& B) {
    cudaStream_t st = ctx->stream;
    const int64_t* p64 = (const int64_t*)ctx->scratch[16];
    const int64_t* tp = p64 + (nc + 1) + gs_partial_count(nc);
    k_tile_combine<<<grid_for(nc, 128, 16), 128, 0, st>>>((const double*)ctx->scratch[18], S.M, B.E.cntB, B.E.cntA, p64, tp, S.start, pv->cur, pv->alt,
                                                         nc, W, B.pcache, S.flags);
    MB_LAUNCH_CHECK(ctx);
    k_tile_moments_fallback<<<grid_for(nc * 32, 256, 8), 256, 0, st>>>(pv->alt, S.start, nc, B.pcache, S.flags);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

template <int W>
static int launch_band(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species, SortScratch& S, const BandBufs& B,
                       bool cls_cached, bool use_tile, bool tile_mom) {
    const int64_t nc = pia->n_cells;
    Indexer* ix = pia->d_indexer + (species - 1) * nc;
    cudaStream_t st = ctx->stream;
    const int use_x = grid != nullptr;
    const int nscan = (int)((nc + SCAN_TILE - 1) / SCAN_TILE);
    const int wgrid = grid_for(nc * 32, 256, 8);
    const int egrid = N_SM * 4;  // the number of extras is only known on the device: grid-stride over the list
    {
        ProfScope ps(ctx, PROF_SORT_CLASSIFY);
        if (!cls_cached && extras_clear(ctx, B, nc)) return MB_ERR_CUDA;
        // with a cached classification this is a stub unless a cell was too big for the fused kernel's staging area
        k_band_classify<W><<<wgrid, 256, 0, st>>>(pv->cur.a[F_X], pv->cell, pv->cell, ix, nc, use_x ? grid->inv_dx : 0.0,
                                                 use_x ? grid->cell_offset : 0, use_x, B.drop, S.M, B.seg_lo, B.seg_n, B.dr, B.E, S.flags,
                                                 cls_cached ? S.flags + F_CLS_REDO : nullptr);
        MB_LAUNCH_CHECK(ctx);
        if (B.n_arr > 0) {
            k_band_arrivals<<<grid_for(B.n_arr, 256), 256, 0, st>>>(pv->cur.a[F_X], B.n_old, B.d_n_arr, nc, grid->inv_dx, grid->cell_offset, pv->cell,
                                                                  B.E, S.flags);
            MB_LAUNCH_CHECK(ctx);
        }
    }
    {
        ProfScope ps(ctx, PROF_SORT_SCAN);
        k_scan_reduce<W><<<nscan, SCAN_BLOCK, 0, st>>>(S.M, B.E.cntB, B.E.cntA, S.hist, nc, S.partial, S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
        k_scan_partials<<<1, 1024, 0, st>>>(S.partial, nscan, S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
        k_scan_apply<<<nscan, SCAN_BLOCK, 0, st>>>(S.hist, nc, S.partial, S.start, nullptr, ix, B.d_nt_write, S.flags, 0);
        MB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, PROF_SORT_SCATTER);
        const int sgrid = grid_for(nc * 32, 256, MB_SC_GRID);
        // movers through shared memory pay off while the groups are long (narrow bands: ~25 movers per neighbour and cell at ppc = 1000);
        // at w >= 4 a group holds a handful of particles and the direct stores are faster (measured, profiles/README.md)
        bool buffered = MB_SC_BUF != 0 && W <= 5;
        for (int f = 1; f < 7 && buffered; f++) buffered = pv->alt.a[f] - pv->alt.a[0] == f * (pv->alt.a[1] - pv->alt.a[0]);
        if (use_tile) {
            int r = launch_tile_pass_b(ctx, pv, nc, W, S, B, tile_mom);
            if (r) return r;
        } else if (buffered) {
            static bool attr_done[64] = {false};  // function attributes are per device
            if (!attr_done[ctx->device & 63]) {
                MB_CUDA(cudaFuncSetAttribute(k_band_scatter_buf<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScBuf<W>::SMEM));
                MB_CUDA(cudaFuncSetAttribute(k_band_scatter_buf<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScBuf<W>::SMEM));
                attr_done[ctx->device & 63] = true;
            }
            if (B.P != nullptr)
                k_band_scatter_buf<W, true><<<sgrid, 256, ScBuf<W>::SMEM, st>>>(pv->cur, pv->alt, B.dr, S.M, B.seg_lo, B.seg_n, S.start, B.E.cntB, nc,
                                                                               B.P, S.flags);
            else
                k_band_scatter_buf<W, false><<<sgrid, 256, ScBuf<W>::SMEM, st>>>(pv->cur, pv->alt, B.dr, S.M, B.seg_lo, B.seg_n, S.start, B.E.cntB, nc,
                                                                                nullptr, S.flags);
        } else if (B.P != nullptr) {
            k_band_scatter<W, true><<<sgrid, 256, 0, st>>>(pv->cur, pv->alt, B.dr, S.M, B.seg_lo, B.seg_n, S.start, B.E.cntB, nc, B.P, S.flags);
        } else {
            k_band_scatter<W, false><<<sgrid, 256, 0, st>>>(pv->cur, pv->alt, B.dr, S.M, B.seg_lo, B.seg_n, S.start, B.E.cntB, nc, nullptr, S.flags);
        }
        if (!use_tile) MB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, PROF_SORT_EXTRAS);
        k_extra_slots<<<egrid, 256, 0, st>>>(B.E, S.start, S.hist, S.perm, S.flags);
        MB_LAUNCH_CHECK(ctx);
        k_extra_regions<<<grid_for(nc, 256, 4), 256, 0, st>>>(B.E, S.start, S.hist, S.perm, nc, pv->cur, pv->alt, S.flags);
        MB_LAUNCH_CHECK(ctx);
        k_save_extras<<<1, 1, 0, st>>>(B.E.n, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    if (use_tile && tile_mom) {
        ProfScope ps(ctx, PROF_SORT_SCAN);
        int r = launch_tile_moments(ctx, pv, nc, W, S, B);
        if (r) return r;
    } else if (B.P != nullptr) {
        ProfScope ps(ctx, PROF_SORT_SCAN);
        if constexpr (W + 2 <= 32)
            k_band_combine_warp<W><<<grid_for(nc * 32, 256, 8), 256, 0, st>>>(B.P, S.M, B.E.cntB, B.E.cntA, B.seg_lo, B.seg_n, S.start, pv->cur, pv->alt,
                                                                            nc, B.pcache, S.flags);
        else
            k_band_combine<W><<<grid_for(nc, 128, 16), 128, 0, st>>>(B.P, S.M, B.E.cntB, B.E.cntA, B.seg_lo, B.seg_n, S.start, pv->cur, pv->alt, nc,
                                                                   B.pcache, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_squash_pia(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, int64_t species);

int mb_sort_set_band_halfwidth(mb_ctx* ctx, int32_t w) {
    MB_ARG(ctx && (w == 0 || w == 1 || w == 2 || w == 4 || w == 8 || w == 15), "band half-width must be 0, 1, 2, 4, 8 or 15");
    ctx->band_w = w;
    return MB_OK;
}
int mb_sort_last_path(mb_ctx* ctx) {
    if (!ctx) return -1;
    if (mb_sync(ctx)) return -1;
    return ctx->sort_last_path != 2 && ctx->h_flags[2] == 0 ? ctx->sort_last_path : 2;
}
int mb_sort_last_pass_b(mb_ctx* ctx) { return ctx ? ctx->sort_last_tile : -1; }
int64_t mb_sort_last_extras(mb_ctx* ctx) {
    if (!ctx) return -1;
    if (mb_sync(ctx)) return -1;
    return ctx->sort_last_path == 1 && ctx->h_flags[2] == 0 ? (int64_t)ctx->h_flags[3] : -1;
}

int mb_sort_particles(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species) {
    MB_ARG(ctx && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(grid == nullptr || grid->n_cells == pia->n_cells, "grid.n_cells != pia.n_cells");
    MB_ARG(pia->n_cells < (int64_t)INT_MAX && pv->cap < (int64_t)INT_MAX, "more than 2^31 cells or particles per GPU");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    // grid_sorting.jl:69-71: squash first.  A non-contiguous species never has the sorted layout (a merge cleared it), so the sort
    // takes the general path and the squash is folded into it (k_build_src); the separate payload pass is only the fallback.
    // (arrivals of a slab exchange on a non-contiguous layout are parked at the end of the capacity and enter through the same map)
    const bool fuse_squash = !pia->contiguous[s] && !(ctx->band_w > 0 && pia->sorted_layout[s]) && (pv->n_arrivals == 0 || pv->arrivals_at_end);
    if (pv->n_arrivals > 0 && pv->arrivals_at_end && !fuse_squash) {
        set_error("sort_particles!: arrivals of a slab exchange on a non-contiguous layout are pending, but the layout is contiguous now");
        return MB_ERR_PRECONDITION;
    }
    if (!pia->contiguous[s] && !fuse_squash) {
        int r = mb_squash_pia(ctx, pv, pia, species);
        if (r) return r;
    }
    int r = pv_ensure_alt(pv);
    if (r) return r;
    const int64_t nc = pia->n_cells, cap = pv->cap;
    const int w = ctx->band_w;
    const int W = 2 * w + 1;
    const int drop = pv->drop_oob;
    const int64_t n_arr = pv->n_arrivals;
    const bool use_x = grid != nullptr;
    // band path: sorted layout; band outliers and slab-exchange arrivals are merged in as "extras" as long as they fit their list
    const bool try_band = w > 0 && pia->sorted_layout[s] && n_arr <= cap / 8 + 32768 && (n_arr == 0 || use_x);
    const int nscan = (int)((nc + SCAN_TILE - 1) / SCAN_TILE);

    SortScratch S;
    BandBufs B;
    if (sort_scratch_layout(ctx, cap, nc, W, S, B)) return MB_ERR_CUDA;
    B.n_arr = n_arr;
    B.d_n_arr = pv->d_n_arr;
    B.drop = drop;
    cudaStream_t st = ctx->stream;
    Indexer* ix = pia->d_indexer + (species - 1) * nc;
    int64_t* d_nt = pia->d_n_total + (species - 1);
    const bool rewrite_total = drop || n_arr > 0;
    B.d_nt_write = rewrite_total ? d_nt : nullptr;
    // moments of the sorted cells come for free in the gather pass (used by compute_props_sorted! if nothing changes in between)
    B.pcache = (double*)ctx_scratch(ctx, 10, (size_t)nc * 6 * 8);
    if (!B.pcache) return MB_ERR_CUDA;
    // the band path caches the cell moments only for narrow bands (most of a cell stays: the few movers are re-read by k_band_combine)
    // pass B as a tile kernel (mb_sort_tile.cuh) when the mean cell population suits the tile size; it caches the moments for every band width
    bool use_tile = false;
    if (try_band) {
        const int tm = tile_mode();
        const int64_t avg = (pia->n_bound[s] > 0 ? pia->n_bound[s] : cap) / (nc > 0 ? nc : 1);
        // measured (profiles/README.md): at w <= 2 (95 % of a cell stays) the warp-per-cell scatter is faster, from w = 4 on the tile kernel
        use_tile = tm == 2 || ((tm == 3 || (tm == 1 && w >= 4)) && avg >= 24 && avg <= tile_ncap(tile_cfg()) / 2);
    }
    const bool band_moments = try_band && (w <= 2 || (use_tile && getenv("MB_TILE_NOMOM") == nullptr));  // (MB_TILE_NOMOM: experiment knob)
    if (band_moments && !use_tile) {
        B.P = (double*)ctx_scratch(ctx, 11, (size_t)nc * 5 * 8);
        if (!B.P) return MB_ERR_CUDA;
    }
    S.perm = (int32_t*)ctx_scratch(ctx, 3, (size_t)cap * 4);  // general path: permutation; band path: slots of the extras
    if (!S.perm) return MB_ERR_CUDA;

    MB_CUDA(cudaMemcpyAsync(B.n_old, d_nt, 8, cudaMemcpyDeviceToDevice, st));  // n_total before the sort (the scan may rewrite it)
    // segment path: sorting by stored cell ids a layout in which (presumably) nobody changed cell
    const bool try_seg = !try_band && !use_x && n_arr == 0 && !drop;
    if (try_seg) {
        ProfScope ps(ctx, PROF_SORT_SCATTER);
        int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)(2 * nc) * 4);
        int64_t* p64 = (int64_t*)ctx_scratch(ctx, 13, ((size_t)(2 * nc + 1) + gs_partial_count(2 * nc)) * 8);
        if (!cnt || !p64) return MB_ERR_CUDA;
        k_set_flag<<<1, 1, 0, st>>>(ctx->d_flags, 2, 0);
        MB_LAUNCH_CHECK(ctx);
        k_segpath_counts<<<grid_for(nc, 256), 256, 0, st>>>(ix, nc, cnt);
        MB_LAUNCH_CHECK(ctx);
        r = device_exclusive_scan(ctx, cnt, 2 * nc, p64, p64 + (2 * nc + 1));
        if (r) return r;
        CellMajorDesc Dc{ix, p64, nullptr};
        KeyCheckAct Ac{pv->cell, ctx->d_flags + 2};
        r = seg_copy(ctx, 7, cap, 2 * nc, Dc, Ac);
        if (r) return r;
        CellMajorDesc Dm{ix, p64, ctx->d_flags + 2};
        PayloadMoveAct Am{pv->cur, pv->alt};
        r = seg_copy(ctx, 14, cap, 2 * nc, Dm, Am);
        if (r) return r;
        // (the indexers are rewritten after the general path had its chance to read them: k_segpath_commit below)
    }
    // classification cached by the fused convect kernel?  (same particles, same grid, nothing but a slab exchange in between)
    const bool cls_cached = try_band && use_x && ctx->cls_gen == ctx->state_gen && ctx->cls_pv == (void*)pv && ctx->cls_pia == (void*)pia &&
                            ctx->cls_species == (int)species && ctx->cls_w == w && ctx->cls_inv_dx == grid->inv_dx &&
                            ctx->cls_cell_offset == grid->cell_offset && ctx->cls_cap == cap;
    if (cls_cached) k_flag_from_cls<<<1, 1, 0, st>>>(ctx->d_flags, drop ? 1 : 0);
    else if (!try_seg) k_set_flag<<<1, 1, 0, st>>>(ctx->d_flags, 2, try_band ? 0 : 1);
    if (!try_seg) MB_LAUNCH_CHECK(ctx);
    if (try_band) {
        if (w == 1) r = launch_band<3>(ctx, grid, pv, pia, species, S, B, cls_cached, use_tile, band_moments);
        else if (w == 2) r = launch_band<5>(ctx, grid, pv, pia, species, S, B, cls_cached, use_tile, band_moments);
        else if (w == 4) r = launch_band<9>(ctx, grid, pv, pia, species, S, B, cls_cached, use_tile, band_moments);
        else if (w == 8) r = launch_band<17>(ctx, grid, pv, pia, species, S, B, cls_cached, use_tile, band_moments);
        else r = launch_band<31>(ctx, grid, pv, pia, species, S, B, cls_cached, use_tile, band_moments);
        if (r) return r;
    }
    // general path (every kernel returns immediately unless flags[2] != 0)
    bool gather_cells = false;
    {
        ProfScope ps(ctx, PROF_SORT_GENERAL);
        const int64_t nb = pia->n_bound[s] > 0 ? pia->n_bound[s] : cap;
        int32_t* src = nullptr;
        if (fuse_squash) {
            r = build_src_map(ctx, cap, ix, nc, n_arr > 0 ? pv->d_n_arr : nullptr, &src);
            if (r) return r;
        }
        MB_CUDA(cudaMemsetAsync(S.hist, 0, (size_t)nc * 4, st));  // harmless for the band result: hist is not read again
        const int pgrid = grid_for(nb, 256, 16);
        if (drop == 2) MB_CUDA(cudaMemsetAsync(ctx->d_xch_counts + 10, 0, 8, st));
        k_gen_classify<<<pgrid, 256, 0, st>>>(pv->cur.a[F_X], pv->cell, pv->cell, S.key, B.n_old, nc, use_x ? grid->inv_dx : 0.0,
                                             use_x ? grid->cell_offset : 0, use_x ? 1 : 0, drop ? 1 : 0, S.hist, S.flags, src,
                                             drop == 2 ? (unsigned long long*)(ctx->d_xch_counts + 10) : nullptr);
        MB_LAUNCH_CHECK(ctx);
        if (drop == 2) {
            k_check_dropped<<<1, 1, 0, st>>>(ctx->d_xch_counts, S.flags);
            MB_LAUNCH_CHECK(ctx);
        }
        k_scan_reduce<3><<<nscan, SCAN_BLOCK, 0, st>>>(nullptr, nullptr, nullptr, S.hist, nc, S.partial, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_scan_partials<<<1, 1024, 0, st>>>(S.partial, nscan, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_scan_apply<<<nscan, SCAN_BLOCK, 0, st>>>(S.hist, nc, S.partial, S.start, S.cursor, ix, rewrite_total ? d_nt : nullptr, S.flags, 1);
        MB_LAUNCH_CHECK(ctx);
        k_gen_scatter_idx<<<pgrid, 256, 0, st>>>(S.key, B.n_old, S.start, S.cursor, S.perm, S.flags);
        MB_LAUNCH_CHECK(ctx);
        gather_cells = nb / (nc > 0 ? nc : 1) <= 2048;  // small cells: gather by cell and cache the cell moments
        if (!gather_cells) {  // (the gathers by cell put the indices of a cell of up to WSEG particles in order themselves)
            k_gen_sort_segments_warp<<<grid_for(nc * 8, 256, 6), 256, 0, st>>>(S.perm, S.start, nc, S.flags);
            MB_LAUNCH_CHECK(ctx);
        }
        {
            const int gseg = grid_for(nc * 256, 256, 8);
            int cpb = 256;  // cells a CTA looks at per round: 256 on long grids, fewer when there are few (large) cells
            while (cpb > 1 && nc < (int64_t)8 * gseg * cpb) cpb >>= 1;
            k_gen_sort_segments<<<gseg, 256, 0, st>>>(S.perm, S.start, nc, S.flags, cpb);
        }
        MB_LAUNCH_CHECK(ctx);
        // band path switched off by the caller = displacements of many cells: the gather would be fully scattered; go through 64-byte records
        const bool aos = gather_cells && w == 0 && use_x;
        static const int env_gat = getenv("MB_GATHER_PER_SM") ? atoi(getenv("MB_GATHER_PER_SM")) : 0;  // experiment knob
        const int gat_per_sm = env_gat > 0 ? env_gat : 8;
        if (aos) {
            Rec64* rec = (Rec64*)ctx_scratch(ctx, 15, (size_t)cap * sizeof(Rec64));
            if (!rec) return MB_ERR_CUDA;
            k_gen_pack_aos<<<grid_for(nb * 4, 256, 16), 256, 0, st>>>(pv->cur, B.n_old, src, rec, S.flags);
            MB_LAUNCH_CHECK(ctx);
            k_gen_gather_cells_aos<<<grid_for(nc * 32, 256, gat_per_sm), 256, 0, st>>>(rec, pv->alt, S.perm, S.start, nc, S.flags, S.key, nullptr, B.pcache);
        } else if (gather_cells)
            k_gen_gather_cells<<<grid_for(nc * 32, 256, gat_per_sm), 256, 0, st>>>(pv->cur, pv->alt, S.perm, S.start, nc, S.flags, src, S.key,
                                                                        use_x ? nullptr : pv->cell, B.pcache);
        else
            k_gen_gather<<<pgrid, 256, 0, st>>>(pv->cur, pv->alt, S.perm, S.start + nc, S.flags, src, S.key, use_x ? nullptr : pv->cell);
        MB_LAUNCH_CHECK(ctx);
    }
    if (try_seg) {
        k_segpath_commit<<<grid_for(nc * 32, 256, 8), 256, 0, st>>>(ix, nc, (const int64_t*)ctx->scratch[13], pv->cell, S.flags);
        MB_LAUNCH_CHECK(ctx);
    }
    // ping-pong
    SoA t = pv->cur;
    pv->cur = pv->alt;
    pv->alt = t;
    if (rewrite_total) {
        pv->drop_oob = 0;
        pv->n_arrivals = 0;
        pv->arrivals_at_end = 0;
        pia->h_valid = false;
        MB_CUDA(cudaMemsetAsync(pv->d_n_arr, 0, sizeof(int64_t), st));
    }
    // the cached moments are valid only if the band path ran (device flag 2 == 0): the props kernel checks the flag itself
    ctx->state_gen++;
    ctx->cls_gen = 0;
    ctx->pc_gen = (band_moments || gather_cells) ? ctx->state_gen : 0;
    ctx->pc_general = gather_cells ? 1 : 0;
    // neither fast path attempted: the general path runs unconditionally, its cache is valid as is
    ctx->pc_band = (band_moments || (!try_band && !try_seg)) ? 1 : 0;
    ctx->pc_pv = pv; ctx->pc_pia = pia; ctx->pc_species = (int)species;
    pia->contiguous[s] = 1;      // grid_sorting.jl:112
    pia->contig_pending[s] = 0;
    pia->sorted_layout[s] = 1;
    ctx->sort_last_path = try_band ? 1 : (try_seg ? 3 : 2);
    ctx->sort_last_tile = use_tile ? 1 : 0;
    return MB_OK;
}

}  // extern "C"

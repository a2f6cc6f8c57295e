// merge_octree_N2_based! (merging/merging_octree_N2.jl:1060-1094) and everything it calls: init_octree! (:947-984),
// compute_octree! (:998-1039, greedy refinement of the heaviest refinable bin), split_bin! (:503-633, 8-way stable
// counting sort of the bin's index slice), bin_bounds_inherit! / bin_bounds_recompute! (:341-418), compute_bin_props!
// (:646-700, two-pass weighted mean / variance of v and x), compute_new_particles! (:736-813 0-D, :830-933 1-D with the
// x clamp), delete_particle_end! bookkeeping (particles.jl:478-566).
//
// One CTA per merging cell, all merging cells of the range concurrently.  The refinement loop of a cell is inherently
// sequential (each split depends on the previous one); inside a step the CTA works in parallel: arg-max over the bins,
// octant classification, deterministic block reductions of the octant counts / weights, an order-exact partition of the
// bin's index slice (warp match + shared prefix; reversed inside each octant like the reference's fill-from-the-end loop), and a warp per bin for the moments.  The octree of a cell is bit-identical to
// the reference's for OctreeBinMidSplit (the splits only compare velocities with 0.5 * (v_min + v_max)); merged
// particles agree to rounding (different summation order), conservation holds to ~1e-15 relative.
//
// Workspace: every merging cell gets a private slice of three index arrays at offset = exclusive scan of n_local over
// the range (so the total is <= n_total <= capacity); bins live in a per-CTA global workspace of
// min(max_Nbins, target_np) + 8 entries (a split is only made while total_post_merge_np + 14 <= target_np, and every bin
// holds at least one post-merge particle).
#include <cstdlib>

#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

constexpr int MT = 256;  // max threads per CTA

struct MergeArgs {
    SoA pv;
    Indexer* ix;
    int64_t* n_total;
    int64_t cell_lo, cell_hi, n_cells_total;
    int64_t threshold, target;
    mb_octree_params oc;
    int has_grid;
    double min_x, max_x;
    uint64_t seed;
    uint32_t timestep, substream;
    int32_t *idx, *tmp;
    uint8_t* oct;
    const int64_t* slice;  // [nr + 1]
    int Bmax;
    int32_t *b_np, *b_depth, *b_start, *b_end, *b_out;  // [nCTA][Bmax (+1 for b_out)]
    double *b_w, *b_vmin, *b_vmax;                       // [nCTA][Bmax], [nCTA][3 * Bmax]
    double* outbuf;                                      // [nCTA][2 * Bmax][7]
    int* flags;
    int* noncontig;
    int small_max;  // cells with n_local <= small_max are merged by k_merge_warp
    int cpb;        // k_merge: cells a CTA looks at per round
};

static __global__ void k_merge_counts(const Indexer* __restrict__ ix, int64_t cell_lo, int64_t nr, int64_t threshold, int32_t* __restrict__ cnt) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = ix[cell_lo - 1 + r].n_local;
        cnt[r] = (n > 0 && (threshold < 0 || n > threshold)) ? (int32_t)n : 0;
    }
}

__device__ __forceinline__ int64_t mpos(const Indexer& q, int64_t j) {  // map_cont_index, 0-based physical position
    return (j < q.n_group1 ? j + q.start1 : (j - q.n_group1) + q.start2) - 1;
}

// deterministic block reductions through shared memory (all threads get the result)
__device__ __forceinline__ double block_sum(double x, double* sh) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = x;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < nw; i++) t += sh[i];
    return t;
}
__device__ __forceinline__ double block_min(double x, double* sh) {
    for (int o = 16; o > 0; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = x;
    __syncthreads();
    double t = sh[0];
    for (int i = 1; i < nw; i++) t = fmin(t, sh[i]);
    return t;
}
__device__ __forceinline__ double block_max(double x, double* sh) { return -block_min(-x, sh); }

// bounds of the particles idx[bs..be] (bin_bounds_recompute! :382-418)
__device__ void slice_bounds(const MergeArgs& a, const int32_t* idx, int bs, int be, double* sh, double mn[3], double mx[3]) {
    double lmn[3] = {9299792458.0, 9299792458.0, 9299792458.0}, lmx[3] = {-9299792458.0, -9299792458.0, -9299792458.0};
    for (int j = bs + threadIdx.x; j <= be; j += blockDim.x) {
        const int64_t p = idx[j];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double v = a.pv.a[F_VX + d][p];
            if (v < lmn[d]) lmn[d] = v;
            if (v > lmx[d]) lmx[d] = v;
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) { mn[d] = block_min(lmn[d], sh); mx[d] = block_max(lmx[d], sh); }
}

constexpr int SMALL_NP = 1024;  // bins of up to SMALL_NP particles are split by warp 0 alone

// level 1 of the two-level arg-max: group g covers bins 32 g .. 32 g + 31; key = bit pattern of the weight + 1 for a refinable bin
// (weights are >= 0, so the patterns order like integers), 0 otherwise; gid = the first bin of the group with the largest key
__device__ __forceinline__ void merge_group_update(int g, int Nbins, const double* b_w, const int32_t* b_np, const int32_t* b_depth, int max_depth,
                                                   unsigned long long* s_gkey, int* s_gid) {
    const int lane = threadIdx.x & 31;
    const int b = 32 * g + lane;
    unsigned long long key = 0ull;
    if (b < Nbins && b_np[b] > 2 && b_depth[b] < max_depth) key = (unsigned long long)__double_as_longlong(b_w[b]) + 1ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool win = key != 0ull && hi == mhi && lo == mlo;
    const unsigned bmin = __reduce_min_sync(0xffffffffu, win ? (unsigned)b : 0xffffffffu);
    if (lane == 0) {
        s_gkey[g] = ((unsigned long long)mhi << 32) | mlo;
        s_gid[g] = bmin == 0xffffffffu ? -1 : (int)bmin;
    }
}

__global__ void __launch_bounds__(MT) k_merge(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char mg_dyn[];  // group maxima: (Bmax + 31) / 32 keys + ids
    unsigned long long* s_gkey = (unsigned long long*)mg_dyn;
    int* s_gid = (int*)(s_gkey + ((a.Bmax + 31) >> 5) + 1);
    __shared__ int32_t s_p[SMALL_NP], s_q[SMALL_NP];
    __shared__ double s_pw[SMALL_NP], s_qw[SMALL_NP];  // weights of the slice, gathered together with the velocities
    __shared__ uint8_t s_oct[SMALL_NP];
    __shared__ int s_mode, s_prevN;
    __shared__ double sh[MT / 32];
    __shared__ double sh_w[MT / 32][8];
    __shared__ int sh_c[MT / 32][8];
    __shared__ int s_Nbins, s_total_post, s_refine, s_stop;
    __shared__ int s_cnt[8], s_base[8], s_run[8];
    __shared__ double s_wsum[8];
    __shared__ int s_scan[MT];
    __shared__ int s_carry;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const int B = a.Bmax;
    int32_t* b_np = a.b_np + (int64_t)blockIdx.x * B;
    int32_t* b_depth = a.b_depth + (int64_t)blockIdx.x * B;
    int32_t* b_start = a.b_start + (int64_t)blockIdx.x * B;
    int32_t* b_end = a.b_end + (int64_t)blockIdx.x * B;
    int32_t* b_out = a.b_out + (int64_t)blockIdx.x * (B + 1);
    double* b_w = a.b_w + (int64_t)blockIdx.x * B;
    double* b_vmin = a.b_vmin + (int64_t)blockIdx.x * 3 * B;
    double* b_vmax = a.b_vmax + (int64_t)blockIdx.x * 3 * B;
    double* outbuf = a.outbuf + (int64_t)blockIdx.x * 2 * B * 7;
    const double* __restrict__ PW = a.pv.a[F_W];

    // the CTA looks at cpb (<= blockDim.x) cells at a time (one coalesced read of their sizes) and merges, one after the other, the
    // ones that need it and are too large for k_merge_warp
    __shared__ int s_list[MT], s_nlist;
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
        const int64_t r = rbase + s_list[li];  // the order of the list does not matter: cells are independent
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int N = (int)q.n_local;
        int32_t* idx = a.idx + a.slice[r];
        int32_t* tmp = a.tmp + a.slice[r];
        uint8_t* oct = a.oct + a.slice[r];
        __syncthreads();
        // ---- init_octree! (:947-984)
        for (int j = tid; j < N; j += nt) idx[j] = (int32_t)mpos(q, j);
        __syncthreads();
        {
            double mn[3], mx[3];
            if (a.oc.init_bin_bounds == 3) {
                for (int d = 0; d < 3; d++) { mn[d] = -c_light; mx[d] = c_light; }
            } else {
                slice_bounds(a, idx, 0, N - 1, sh, mn, mx);
                if (a.oc.init_bin_bounds == 2)
                    for (int d = 0; d < 3; d++) { const double m = fmax(fabs(mn[d]), fabs(mx[d])); mn[d] = -m; mx[d] = m; }
            }
            if (tid == 0) {
                b_np[0] = N; b_w[0] = 1e50; b_depth[0] = 0; b_start[0] = 0; b_end[0] = N - 1;
                for (int d = 0; d < 3; d++) { b_vmin[d] = mn[d]; b_vmax[d] = mx[d]; }
                s_Nbins = 1;
                s_total_post = N >= 2 ? 2 : N;
                s_stop = 0;
            }
        }
        __syncthreads();
        // ---- compute_octree! (:998-1039).  The greedy refinement is sequential; most of its ~target/7 splits act on bins of a few
        //      dozen particles, so warp 0 runs the loop alone without block barriers -- two-level arg-max (group maxima in shared
        //      memory, hardware redux), octant classification and the order-exact partition staged in shared memory -- and calls in
        //      the whole CTA only for bins of more than SMALL_NP particles (the first few levels of the tree).
        if (wid == 0) merge_group_update(0, 1, b_w, b_np, b_depth, a.oc.max_depth, s_gkey, s_gid);
        __syncthreads();
        while (true) {
            if (wid == 0) {
                int Nbins = s_Nbins, total_post = s_total_post, mode = 0;
                while (!s_stop) {
                    // first bin with the strictly largest weight among the refinable ones
                    int bin;
                    {
                        const int ng = (Nbins + 31) >> 5;
                        unsigned long long key = 0ull;
                        int g_best = 0x7fffffff;
                        for (int g = lane; g < ng; g += 32) {
                            const unsigned long long k = s_gkey[g];
                            if (k > key) { key = k; g_best = g; }
                        }
                        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                        const bool win = key != 0ull && hi == mhi && lo == mlo;
                        const unsigned g_min = __reduce_min_sync(0xffffffffu, win ? (unsigned)g_best : 0xffffffffu);
                        bin = g_min == 0xffffffffu ? -1 : s_gid[g_min];
                    }
                    if (bin < 0 || total_post + 14 > a.target) break;
                    const int bs = b_start[bin], be = b_end[bin], depth = b_depth[bin];
                    const int n = be - bs + 1;
                    if (n > SMALL_NP) { if (lane == 0) s_refine = bin; mode = 1; break; }
                    // ---- split_bin! (:503-633) by one warp
                    double pmn[3], pmx[3];
                    if (a.oc.bin_bounds_compute == 2) {
                        for (int d = 0; d < 3; d++) { pmn[d] = 9299792458.0; pmx[d] = -9299792458.0; }
                        for (int j = lane; j < n; j += 32) {
                            const int64_t pj = idx[bs + j];
#pragma unroll
                            for (int d = 0; d < 3; d++) { const double v = a.pv.a[F_VX + d][pj]; pmn[d] = fmin(pmn[d], v); pmx[d] = fmax(pmx[d], v); }
                        }
#pragma unroll
                        for (int d = 0; d < 3; d++)
                            for (int o = 16; o > 0; o >>= 1) {
                                pmn[d] = fmin(pmn[d], __shfl_xor_sync(0xffffffffu, pmn[d], o));
                                pmx[d] = fmax(pmx[d], __shfl_xor_sync(0xffffffffu, pmx[d], o));
                            }
                    } else {
                        for (int d = 0; d < 3; d++) { pmn[d] = b_vmin[3 * bin + d]; pmx[d] = b_vmax[3 * bin + d]; }
                    }
                    double mid[3];
                    if (a.oc.split == 1) {
                        for (int d = 0; d < 3; d++) mid[d] = 0.5 * (pmn[d] + pmx[d]);
                    } else {
                        double sw = 0, sv[3] = {0, 0, 0};
                        for (int j = lane; j < n; j += 32) {
                            const int64_t pj = idx[bs + j];
                            const double w = PW[pj];
                            sw += w;
                            for (int d = 0; d < 3; d++) sv[d] += w * a.pv.a[F_VX + d][pj];
                        }
                        for (int o = 16; o > 0; o >>= 1) {
                            sw += __shfl_xor_sync(0xffffffffu, sw, o);
                            for (int d = 0; d < 3; d++) sv[d] += __shfl_xor_sync(0xffffffffu, sv[d], o);
                        }
                        for (int d = 0; d < 3; d++) mid[d] = sv[d] / sw;
                    }
                    int cnt[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) cnt[k] = 0;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int j = c0 + lane;
                        int o = 8;
                        if (j < n) {
                            const int32_t pj = idx[bs + j];
                            s_p[j] = pj;
                            const double pvx = a.pv.a[F_VX][pj], pvy = a.pv.a[F_VY][pj], pvz = a.pv.a[F_VZ][pj];
                            s_pw[j] = PW[pj];  // same round trip as the velocities: the child weights below need no second gather
                            o = (pvx > mid[0] ? 1 : 0) + (pvy > mid[1] ? 2 : 0) + (pvz > mid[2] ? 4 : 0);
                            s_oct[j] = (uint8_t)o;
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++) cnt[k] += __popc(__ballot_sync(0xffffffffu, o == k));
                    }
                    int basek[8], run = 0, n_ne = 0, tp = total_post - 2, my_id = -1, my_cnt = 0, my_base = 0, w_id = -1, w_cnt = 0, w_base = 0;
#pragma unroll
                    for (int o = 0; o < 8; o++) {
                        basek[o] = run;
                        const int c = cnt[o];
                        if (c > 0) {
                            const int id = n_ne == 0 ? bin : Nbins + n_ne - 1;
                            n_ne += 1;
                            tp += c >= 2 ? 2 : c;
                            if (lane == o) { my_id = id; my_cnt = c; my_base = run; }
                            if ((lane >> 2) == o) { w_id = id; w_cnt = c; w_base = run; }
                        }
                        run += c;
                    }
                    __syncwarp();  // every lane has read the parent's descriptors and bounds
                    if (my_id >= 0) {
                        b_start[my_id] = bs + my_base;
                        b_end[my_id] = bs + my_base + my_cnt - 1;
                        b_np[my_id] = my_cnt;
                        b_depth[my_id] = depth + 1;
                        for (int d = 0; d < 3; d++) {
                            if (a.oc.bin_bounds_compute == 1) {
                                const bool upper = (lane >> d) & 1;
                                b_vmin[3 * my_id + d] = upper ? mid[d] : pmn[d];
                                b_vmax[3 * my_id + d] = upper ? pmx[d] : mid[d];
                            } else {
                                b_vmin[3 * my_id + d] = pmn[d];
                                b_vmax[3 * my_id + d] = pmx[d];
                            }
                        }
                    }
                    __syncwarp();
                    int runk[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) runk[k] = 0;
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int j = c0 + lane;
                        const int o = j < n ? (int)s_oct[j] : 8;
                        int dest = -1;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const unsigned bal = __ballot_sync(0xffffffffu, o == k);
                            if (o == k) dest = basek[k] + (cnt[k] - 1 - (runk[k] + __popc(bal & lt)));
                            runk[k] += __popc(bal);
                        }
                        if (dest >= 0) { s_q[dest] = s_p[j]; s_qw[dest] = s_pw[j]; }
                    }
                    __syncwarp();
                    for (int j = lane; j < n; j += 32) idx[bs + j] = s_q[j];
                    {
                        double w = 0.0;
                        if (w_id >= 0) {
                            const int chunk = (w_cnt + 3) >> 2, part = lane & 3;
                            const int hi_j = w_base + w_cnt - 1 - part * chunk;
                            int lo_j = hi_j - chunk + 1;
                            if (lo_j < w_base) lo_j = w_base;
                            for (int j = hi_j; j >= lo_j; j--) w += s_qw[j];
                        }
                        w += __shfl_xor_sync(0xffffffffu, w, 1);
                        w += __shfl_xor_sync(0xffffffffu, w, 2);
                        if (w_id >= 0 && (lane & 3) == 0) b_w[w_id] = w;
                    }
                    const int Nb0 = Nbins;
                    Nbins += n_ne - 1;
                    total_post = tp;
                    __syncwarp();
                    {   // the (at most three) groups the split touched: loads of all of them in flight together, then the reductions
                        const int g0 = bin >> 5, g1 = Nb0 >> 5, g2 = (Nbins - 1) >> 5;
                        const int gs[3] = {g0, g1, g2};
                        unsigned long long key[3];
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            const int b = 32 * gs[t] + lane;
                            key[t] = 0ull;
                            if (b < Nbins && b_np[b] > 2 && b_depth[b] < a.oc.max_depth) key[t] = (unsigned long long)__double_as_longlong(b_w[b]) + 1ull;
                        }
#pragma unroll
                        for (int t = 0; t < 3; t++) {
                            if ((t == 1 && g1 == g0) || (t == 2 && (g2 == g0 || g2 == g1))) continue;  // warp-uniform
                            const unsigned hi = (unsigned)(key[t] >> 32), lo = (unsigned)key[t];
                            const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                            const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                            const bool win = key[t] != 0ull && hi == mhi && lo == mlo;
                            const unsigned bmin = __reduce_min_sync(0xffffffffu, win ? (unsigned)(32 * gs[t] + lane) : 0xffffffffu);
                            if (lane == 0) {
                                s_gkey[gs[t]] = ((unsigned long long)mhi << 32) | mlo;
                                s_gid[gs[t]] = bmin == 0xffffffffu ? -1 : (int)bmin;
                            }
                        }
                    }
                    __syncwarp();
                    if (Nbins + 7 > a.oc.max_Nbins) break;
                }
                if (lane == 0) { s_mode = mode; s_Nbins = Nbins; s_total_post = total_post; s_prevN = Nbins; }
            }
            __syncthreads();
            if (s_mode == 0) break;
            // ---- split_bin! (:503-633) by the whole CTA (bins of more than SMALL_NP particles)
            const int bin = s_refine;
            const int bs = b_start[bin], be = b_end[bin];
            const int depth = b_depth[bin];
            double pmn[3], pmx[3];
            if (a.oc.bin_bounds_compute == 2) {
                slice_bounds(a, idx, bs, be, sh, pmn, pmx);
            } else {
                for (int d = 0; d < 3; d++) { pmn[d] = b_vmin[3 * bin + d]; pmx[d] = b_vmax[3 * bin + d]; }
            }
            double mid[3];
            if (a.oc.split == 1) {
                for (int d = 0; d < 3; d++) mid[d] = 0.5 * (pmn[d] + pmx[d]);
            } else {  // OctreeBinMeanSplit: weighted mean of the bin (the evident intent of compute_v_mean! :431-439)
                double sw = 0, sv[3] = {0, 0, 0};
                for (int j = bs + tid; j <= be; j += nt) {
                    const int64_t p = idx[j];
                    const double w = PW[p];
                    sw += w;
                    for (int d = 0; d < 3; d++) sv[d] += w * a.pv.a[F_VX + d][p];
                }
                sw = block_sum(sw, sh);
                for (int d = 0; d < 3; d++) mid[d] = block_sum(sv[d], sh) / sw;
            }
            // octants, per-octant counts and weights
            int cnt[8];
            double ws[8];
#pragma unroll
            for (int o = 0; o < 8; o++) { cnt[o] = 0; ws[o] = 0.0; }
            for (int j = bs + tid; j <= be; j += nt) {
                const int64_t p = idx[j];
                const int o = (a.pv.a[F_VX][p] > mid[0] ? 1 : 0) + (a.pv.a[F_VY][p] > mid[1] ? 2 : 0) + (a.pv.a[F_VZ][p] > mid[2] ? 4 : 0);
                oct[j] = (uint8_t)o;
                const double w = PW[p];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (o == k) { cnt[k] += 1; ws[k] += w; }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                for (int o = 16; o > 0; o >>= 1) {
                    cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
                    ws[k] += __shfl_xor_sync(0xffffffffu, ws[k], o);
                }
                if (lane == 0) { sh_c[wid][k] = cnt[k]; sh_w[wid][k] = ws[k]; }
            }
            __syncthreads();
            if (tid < 8) {
                int c = 0;
                double w = 0;
                for (int i = 0; i < nw; i++) { c += sh_c[i][tid]; w += sh_w[i][tid]; }
                s_cnt[tid] = c;
                s_wsum[tid] = w;
            }
            __syncthreads();
            if (tid == 0) {
                // children: the first non-empty octant reuses the parent id, the others get Nbins+1.. in octant order (:487-489,:563-572)
                int run = 0, n_ne = 0, nb = s_Nbins, tp = s_total_post - 2;
                for (int o = 0; o < 8; o++) {
                    s_base[o] = run;
                    s_run[o] = 0;
                    const int c = s_cnt[o];
                    if (c > 0) {
                        const int id = n_ne == 0 ? bin : nb + n_ne - 1;
                        n_ne += 1;
                        b_start[id] = bs + run;
                        b_end[id] = bs + run + c - 1;
                        b_np[id] = c;
                        b_w[id] = s_wsum[o];
                        b_depth[id] = depth + 1;
                        if (a.oc.bin_bounds_compute == 1) {  // bin_bounds_inherit! (:341-367)
                            for (int d = 0; d < 3; d++) {
                                const bool upper = (o >> d) & 1;
                                b_vmin[3 * id + d] = upper ? mid[d] : pmn[d];
                                b_vmax[3 * id + d] = upper ? pmx[d] : mid[d];
                            }
                        } else {
                            for (int d = 0; d < 3; d++) { b_vmin[3 * id + d] = pmn[d]; b_vmax[3 * id + d] = pmx[d]; }
                        }
                        tp += c >= 2 ? 2 : c;
                    }
                    run += c;
                }
                s_Nbins = nb + n_ne - 1;
                s_total_post = tp;
                if (s_Nbins + 7 > a.oc.max_Nbins) s_stop = 1;
            }
            __syncthreads();
            // stable partition of idx[bs..be] by octant
            const int n = be - bs + 1;
            for (int c0 = 0; c0 < n; c0 += nt) {
                const int j = c0 + tid;
                const bool valid = j < n;
                int o = -1, rank = 0;
                const unsigned act = __ballot_sync(0xffffffffu, valid);
                if (tid < 8 * nw) sh_c[tid >> 3][tid & 7] = 0;
                __syncthreads();
                if (valid) {
                    o = oct[bs + j];
                    const unsigned peers = __match_any_sync(act, o);
                    rank = __popc(peers & lt);
                    if (rank == 0) sh_c[wid][o] = __popc(peers);
                }
                __syncthreads();
                if (valid) {
                    // the reference walks the slice forwards while filling each octant's range from its END
                    // (merging_octree_N2.jl:618-623), i.e. the order inside an octant is reversed at every split
                    int before = s_run[o] + rank;
                    for (int i = 0; i < wid; i++) before += sh_c[i][o];
                    tmp[bs + s_base[o] + (s_cnt[o] - 1 - before)] = idx[bs + j];
                }
                __syncthreads();
                if (tid < 8) {
                    int t = 0;
                    for (int i = 0; i < nw; i++) t += sh_c[i][tid];
                    s_run[tid] += t;
                }
                __syncthreads();
            }
            for (int j = tid; j < n; j += nt) idx[bs + j] = tmp[bs + j];
            __syncthreads();
            if (wid == 0) {  // refresh the group maxima the split touched
                const int Nb0 = s_prevN, Nb1 = s_Nbins, bin0 = s_refine;
                merge_group_update(bin0 >> 5, Nb1, b_w, b_np, b_depth, a.oc.max_depth, s_gkey, s_gid);
                for (int g = Nb0 >> 5; g <= (Nb1 - 1) >> 5; g++)
                    if (g != (bin0 >> 5)) merge_group_update(g, Nb1, b_w, b_np, b_depth, a.oc.max_depth, s_gkey, s_gid);
            }
            __syncthreads();
        }
        __syncthreads();
        const int Nbins = s_Nbins;
        if (Nbins == 1) {  // :1028-1034
            double sw = 0;
            for (int j = tid; j < N; j += nt) sw += PW[idx[j]];
            sw = block_sum(sw, sh);
            if (tid == 0) b_w[0] = sw;
            __syncthreads();
        }
        // ---- compute_bin_props! (:646-700) + the per-bin part of compute_new_particles! (:742-784): one THREAD per bin of up to 32
        //      particles (sequential in slice order, like the reference), one warp per larger bin
        for (int b = tid; b < Nbins; b += nt) {
            int np = b_np[b];
            const double w = b_w[b];
            if (w == 0) np = 0;
            const int bs = b_start[b], be = b_end[b];
            double* o1 = outbuf + (int64_t)(2 * b) * 7;
            double* o2 = o1 + 7;
            if (np > 2 && np <= 32) {
                const double inv_w = 1.0 / w;
                double m[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
                for (int j = bs; j <= be; j++) {
                    const int64_t p = idx[j];
                    const double pw = PW[p];
#pragma unroll
                    for (int d = 0; d < 6; d++) m[d] = m[d] + pw * a.pv.a[F_VX + d][p];
                }
#pragma unroll
                for (int d = 0; d < 6; d++) m[d] *= inv_w;
                for (int j = bs; j <= be; j++) {
                    const int64_t p = idx[j];
                    const double pw = PW[p];
#pragma unroll
                    for (int d = 0; d < 6; d++) { const double dd = a.pv.a[F_VX + d][p] - m[d]; s2[d] = s2[d] + pw * dd * dd; }
                }
                uint32_t rb[4];
                philox4x32_10((uint32_t)b, (uint32_t)cell, a.timestep, (OP_MERGE & 0xFFu) | (a.substream << 8), (uint32_t)a.seed,
                              (uint32_t)(a.seed >> 32), rb);
                o1[0] = 0.5 * w; o2[0] = 0.5 * w;
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    const double sd = sqrt(s2[d] * inv_w);
                    const double sg = ((rb[0] >> d) & 1u) ? 1.0 : -1.0;
                    double x1 = m[d] + sg * sd, x2 = m[d] - sg * sd;
                    if (d == 3 && a.has_grid) {  // :878-897: clamp x1 of the np > 2 outputs into [min_x, max_x]
                        x1 = x1 < a.min_x ? a.min_x : (x1 > a.max_x ? a.max_x : x1);
                        x2 = x2 < a.min_x ? a.min_x : (x2 > a.max_x ? a.max_x : x2);
                    }
                    o1[1 + d] = x1;
                    o2[1 + d] = x2;
                }
            } else if (np >= 1 && np <= 2) {
                const int64_t p1 = idx[bs];
#pragma unroll
                for (int f = 0; f < 7; f++) o1[f] = a.pv.a[f][p1];
                if (np == 2) {
                    const int64_t p2 = idx[bs + 1];
#pragma unroll
                    for (int f = 0; f < 7; f++) o2[f] = a.pv.a[f][p2];
                }
            }
            b_out[b] = np >= 2 ? 2 : np;
        }
        for (int b0 = wid * 32; b0 < Nbins; b0 += nw * 32) {
            const int mb_ = b0 + lane;
            unsigned big = __ballot_sync(0xffffffffu, mb_ < Nbins && b_np[mb_] > 32 && b_w[mb_] != 0);
            while (big) {
                const int b = b0 + (__ffs(big) - 1);
                big &= big - 1;
                const double w = b_w[b];
                const int bs = b_start[b], be = b_end[b];
                double* o1 = outbuf + (int64_t)(2 * b) * 7;
                double* o2 = o1 + 7;
                const double inv_w = 1.0 / w;
                double m[6] = {0, 0, 0, 0, 0, 0};
                for (int j = bs + lane; j <= be; j += 32) {
                    const int64_t p = idx[j];
                    const double pw = PW[p];
#pragma unroll
                    for (int d = 0; d < 6; d++) m[d] += pw * a.pv.a[F_VX + d][p];  // vx,vy,vz,x,y,z are consecutive fields
                }
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    for (int o = 16; o > 0; o >>= 1) m[d] += __shfl_xor_sync(0xffffffffu, m[d], o);
                    m[d] *= inv_w;
                }
                double s2[6] = {0, 0, 0, 0, 0, 0};
                for (int j = bs + lane; j <= be; j += 32) {
                    const int64_t p = idx[j];
                    const double pw = PW[p];
#pragma unroll
                    for (int d = 0; d < 6; d++) { const double dd = a.pv.a[F_VX + d][p] - m[d]; s2[d] += pw * dd * dd; }
                }
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    for (int o = 16; o > 0; o >>= 1) s2[d] += __shfl_xor_sync(0xffffffffu, s2[d], o);
                    s2[d] = sqrt(s2[d] * inv_w);
                }
                if (lane == 0) {
                    // rand(rng, direction_signs, 3) twice (:749,:753): Philox block (bin_id - 1) of the cell's merge stream
                    uint32_t rb[4];
                    philox4x32_10((uint32_t)b, (uint32_t)cell, a.timestep, (OP_MERGE & 0xFFu) | (a.substream << 8), (uint32_t)a.seed,
                                  (uint32_t)(a.seed >> 32), rb);
                    o1[0] = 0.5 * w; o2[0] = 0.5 * w;
#pragma unroll
                    for (int d = 0; d < 6; d++) {
                        const double sg = ((rb[0] >> d) & 1u) ? 1.0 : -1.0;
                        double x1 = m[d] + sg * s2[d], x2 = m[d] - sg * s2[d];
                        if (d == 3 && a.has_grid) {
                            x1 = x1 < a.min_x ? a.min_x : (x1 > a.max_x ? a.max_x : x1);
                            x2 = x2 < a.min_x ? a.min_x : (x2 > a.max_x ? a.max_x : x2);
                        }
                        o1[1 + d] = x1;
                        o2[1 + d] = x2;
                    }
                }
            }
        }
        __syncthreads();
        // exclusive scan of the per-bin output counts (bins in id order)
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int c0 = 0; c0 < Nbins; c0 += nt) {
            const int b = c0 + tid;
            const int v = b < Nbins ? b_out[b] : 0;
            s_scan[tid] = v;
            __syncthreads();
            for (int o = 1; o < nt; o <<= 1) {
                const int t = tid >= o ? s_scan[tid - o] : 0;
                __syncthreads();
                s_scan[tid] += t;
                __syncthreads();
            }
            const int incl = s_scan[tid], carry = s_carry;
            if (b < Nbins) b_out[b] = carry + incl - v;
            __syncthreads();
            if (tid == nt - 1) s_carry = carry + incl;
            __syncthreads();
            // keep the count recoverable: stash it in the sign bit-free upper part is not needed -- recompute below from b_np / b_w
        }
        const int curr = s_carry;
        // ---- write the merged particles into the first `curr` logical slots of the cell (:786-799)
        for (int b = wid; b < Nbins; b += nw) {
            int np = b_np[b];
            if (b_w[b] == 0) np = 0;
            const int no = np >= 2 ? 2 : np;
            const int off = b_out[b];
            for (int k = 0; k < no; k++) {
                const int64_t p = mpos(q, off + k);
                if (lane < 7) a.pv.a[lane][p] = outbuf[(int64_t)(2 * b + k) * 7 + lane];
            }
        }
        // ---- delete_particle_end! x n_delete (:800-812): group 2 shrinks first, then group 1; deleted slots get w = 0
        const int n_del = N - curr;
        for (int j = curr + tid; j < N; j += nt) a.pv.a[F_W][mpos(q, j)] = 0.0;
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
            if (!(cell == a.n_cells_total) || n_del > q.n_group2) *a.noncontig = 1;  // :806-808
        }
        __syncthreads();
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small cells (N <= NS particles; the C4 / small-cell C2 shapes: 130-500 sampled, threshold 130-150, target 100): ONE WARP per
// merging cell, the cell staged once in shared memory (w, v, x: 56 B/particle read once from HBM), every step of the octree
// refinement done by the warp with ballots / shuffles and __syncwarp only -- no block barrier, no global index arrays.
// Summation orders follow the reference's sequential loops wherever one lane does the sum (child weights: forward slice order;
// bin means / variances of bins of <= 32 particles: slice order), so those results are bit-identical to the CPU oracle.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NS_MAX = 256;    // largest cell handled by the warp kernel (instantiated for NS = 160 and 256)
constexpr int MW_WARPS = 4;    // warps (= concurrent cells) per CTA

template <int NS>
__host__ __device__ inline size_t merge_warp_smem(int BC) {
    size_t b = (size_t)7 * NS * 8 + (size_t)BC * 8 + (size_t)NS * 2 * 2 + (size_t)BC * 2 * 5 + (size_t)NS;
    return (b + 15) / 16 * 16;
}

template <int NS>
__global__ void __launch_bounds__(32 * MW_WARPS) k_merge_warp(MergeArgs a, int BC, double* __restrict__ gbounds, int ch) {
    extern __shared__ __align__(16) unsigned char mw_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    unsigned char* base = mw_smem + (size_t)wid * merge_warp_smem<NS>(BC);
    double* P = (double*)base;                 // [7][NS]: field f of local particle j at P[f * NS + j]
    double* bw = P + 7 * NS;                   // [BC] bin weight
    int16_t* idx = (int16_t*)(bw + BC);        // [NS] local particle index per slice position
    int16_t* tmp = idx + NS;                   // [NS]
    int16_t* bnp = tmp + NS;                   // [BC]
    int16_t* bstart = bnp + BC;
    int16_t* bend = bstart + BC;
    int16_t* bdepth = bend + BC;
    int16_t* bout = bdepth + BC;
    uint8_t* oct = (uint8_t*)(bout + BC);      // [NS]
    const int64_t gw = (int64_t)blockIdx.x * MW_WARPS + wid, nwarps = (int64_t)gridDim.x * MW_WARPS;
    double* bvmin = gbounds + gw * 6 * BC;     // [3 * BC] (global, L1/L2 resident: touched once per split)
    double* bvmax = bvmin + 3 * BC;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const double* Pw = P;

    // a warp takes ch (<= 32) consecutive cells at a time (one read of their sizes) and merges the ones that need it
    for (int64_t r0 = gw * ch; r0 < nr; r0 += nwarps * ch) {
      const int64_t myr = r0 + lane;
      const int64_t my_n = (lane < ch && myr < nr) ? a.ix[a.cell_lo - 1 + myr].n_local : 0;
      unsigned todo = __ballot_sync(FULL, my_n > 0 && my_n <= NS && (a.threshold < 0 || my_n > a.threshold));
      while (todo) {
        const int64_t r = r0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int N = (int)q.n_local;
        __syncwarp();
        // ---- stage the cell; init_octree! (:947-984)
        for (int j = lane; j < N; j += 32) {
            const int64_t p = mpos(q, j);
#pragma unroll
            for (int f = 0; f < 7; f++) P[f * NS + j] = a.pv.a[f][p];
            idx[j] = (int16_t)j;
        }
        __syncwarp();
        {
            double mn[3], mx[3];
            if (a.oc.init_bin_bounds == 3) {
                for (int d = 0; d < 3; d++) { mn[d] = -c_light; mx[d] = c_light; }
            } else {
                for (int d = 0; d < 3; d++) { mn[d] = 9299792458.0; mx[d] = -9299792458.0; }
                for (int j = lane; j < N; j += 32)
#pragma unroll
                    for (int d = 0; d < 3; d++) { const double v = P[(1 + d) * NS + j]; mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v); }
#pragma unroll
                for (int d = 0; d < 3; d++)
                    for (int o = 16; o > 0; o >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(FULL, mn[d], o)); mx[d] = fmax(mx[d], __shfl_xor_sync(FULL, mx[d], o)); }
                if (a.oc.init_bin_bounds == 2)
                    for (int d = 0; d < 3; d++) { const double m = fmax(fabs(mn[d]), fabs(mx[d])); mn[d] = -m; mx[d] = m; }
            }
            if (lane == 0) {
                bnp[0] = (int16_t)N; bw[0] = 1e50; bdepth[0] = 0; bstart[0] = 0; bend[0] = (int16_t)(N - 1);
                for (int d = 0; d < 3; d++) { bvmin[d] = mn[d]; bvmax[d] = mx[d]; }
            }
        }
        int Nbins = 1, total_post = N >= 2 ? 2 : N;
        __syncwarp();
        // ---- compute_octree! (:998-1039)
        while (true) {
            double bwm = -1.0;
            int bid = -1;
            for (int b = lane; b < Nbins; b += 32) {
                const double w = bw[b];
                if (w > bwm && bnp[b] > 2 && bdepth[b] < a.oc.max_depth) { bwm = w; bid = b; }
            }
            {   // first bin with the strictly largest weight: weights are >= 0, so their bit patterns order like integers and the
                // hardware warp reductions (redux.sync) replace a 5-round shuffle tournament
                const unsigned long long key = bid >= 0 ? (unsigned long long)__double_as_longlong(bwm) + 1ull : 0ull;
                const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                const unsigned mhi = __reduce_max_sync(FULL, hi);
                const unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
                const bool win = key != 0ull && hi == mhi && lo == mlo;
                const unsigned mid_id = __reduce_min_sync(FULL, win ? (unsigned)bid : 0xffffffffu);
                bid = mid_id == 0xffffffffu ? -1 : (int)mid_id;
            }
            if (bid < 0 || total_post + 14 > a.target) break;
            // ---- split_bin! (:503-633)
            const int bin = bid;
            const int bs = bstart[bin], be = bend[bin], depth = bdepth[bin];
            double pmn[3], pmx[3];
            if (a.oc.bin_bounds_compute == 2) {  // bin_bounds_recompute! (:382-418)
                for (int d = 0; d < 3; d++) { pmn[d] = 9299792458.0; pmx[d] = -9299792458.0; }
                for (int j = bs + lane; j <= be; j += 32) {
                    const int pj = idx[j];
#pragma unroll
                    for (int d = 0; d < 3; d++) { const double v = P[(1 + d) * NS + pj]; pmn[d] = fmin(pmn[d], v); pmx[d] = fmax(pmx[d], v); }
                }
#pragma unroll
                for (int d = 0; d < 3; d++)
                    for (int o = 16; o > 0; o >>= 1) { pmn[d] = fmin(pmn[d], __shfl_xor_sync(FULL, pmn[d], o)); pmx[d] = fmax(pmx[d], __shfl_xor_sync(FULL, pmx[d], o)); }
            } else {
                for (int d = 0; d < 3; d++) { pmn[d] = bvmin[3 * bin + d]; pmx[d] = bvmax[3 * bin + d]; }
            }
            double mid[3];
            if (a.oc.split == 1) {
                for (int d = 0; d < 3; d++) mid[d] = 0.5 * (pmn[d] + pmx[d]);
            } else {  // OctreeBinMeanSplit
                double sw = 0, sv[3] = {0, 0, 0};
                for (int j = bs + lane; j <= be; j += 32) {
                    const int pj = idx[j];
                    const double w = Pw[pj];
                    sw += w;
                    for (int d = 0; d < 3; d++) sv[d] += w * P[(1 + d) * NS + pj];
                }
                for (int o = 16; o > 0; o >>= 1) {
                    sw += __shfl_xor_sync(FULL, sw, o);
                    for (int d = 0; d < 3; d++) sv[d] += __shfl_xor_sync(FULL, sv[d], o);
                }
                for (int d = 0; d < 3; d++) mid[d] = sv[d] / sw;
            }
            __syncwarp();
            // octant of every particle of the slice; per-octant counts (uniform registers)
            int cnt[8];
#pragma unroll
            for (int k = 0; k < 8; k++) cnt[k] = 0;
            for (int c0 = bs; c0 <= be; c0 += 32) {
                const int j = c0 + lane;
                int o = 8;
                if (j <= be) {
                    const int pj = idx[j];
                    o = (P[1 * NS + pj] > mid[0] ? 1 : 0) + (P[2 * NS + pj] > mid[1] ? 2 : 0) + (P[3 * NS + pj] > mid[2] ? 4 : 0);
                    oct[j] = (uint8_t)o;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) cnt[k] += __popc(__ballot_sync(FULL, o == k));
            }
            // children: the first non-empty octant reuses the parent id, the others get Nbins+1.. in octant order (:487-489,:563-572)
            int basek[8], run = 0, n_ne = 0, tp = total_post - 2, my_id = -1, my_cnt = 0, my_base = 0, w_id = -1, w_cnt = 0, w_base = 0;
#pragma unroll
            for (int o = 0; o < 8; o++) {
                basek[o] = run;
                const int c = cnt[o];
                if (c > 0) {
                    const int id = n_ne == 0 ? bin : Nbins + n_ne - 1;
                    n_ne += 1;
                    tp += c >= 2 ? 2 : c;
                    if (lane == o) { my_id = id; my_cnt = c; my_base = run; }
                    if ((lane >> 2) == o) { w_id = id; w_cnt = c; w_base = run; }  // lanes 4o .. 4o+3 sum the weights of octant o
                }
                run += c;
            }
            __syncwarp();  // every lane has read the parent's descriptors and bounds
            if (my_id >= 0) {  // lane o describes the child of octant o
                bstart[my_id] = (int16_t)(bs + my_base);
                bend[my_id] = (int16_t)(bs + my_base + my_cnt - 1);
                bnp[my_id] = (int16_t)my_cnt;
                bdepth[my_id] = (int16_t)(depth + 1);
                for (int d = 0; d < 3; d++) {
                    if (a.oc.bin_bounds_compute == 1) {  // bin_bounds_inherit! (:341-367)
                        const bool upper = (lane >> d) & 1;
                        bvmin[3 * my_id + d] = upper ? mid[d] : pmn[d];
                        bvmax[3 * my_id + d] = upper ? pmx[d] : mid[d];
                    } else {
                        bvmin[3 * my_id + d] = pmn[d];
                        bvmax[3 * my_id + d] = pmx[d];
                    }
                }
            }
            // order-exact partition: the reference walks the slice forwards while filling each octant's range from its END (:618-623)
            int runk[8];
#pragma unroll
            for (int k = 0; k < 8; k++) runk[k] = 0;
            for (int c0 = bs; c0 <= be; c0 += 32) {
                const int j = c0 + lane;
                const int o = j <= be ? (int)oct[j] : 8;
                int dest = -1;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned bal = __ballot_sync(FULL, o == k);
                    if (o == k) dest = bs + basek[k] + (cnt[k] - 1 - (runk[k] + __popc(bal & lt)));
                    runk[k] += __popc(bal);
                }
                if (dest >= 0) tmp[dest] = idx[j];
            }
            __syncwarp();
            for (int j = bs + lane; j <= be; j += 32) idx[j] = tmp[j];
            __syncwarp();
            // child weights (:109): four lanes per octant, each a quarter of the child's range walked backwards (== the reference's
            // forward order inside the quarter), combined as (q0 + q1) + (q2 + q3)
            {
                double w = 0.0;
                if (w_id >= 0) {
                    const int chunk = (w_cnt + 3) >> 2, part = lane & 3;
                    const int hi_j = bs + w_base + w_cnt - 1 - part * chunk;
                    int lo_j = hi_j - chunk + 1;
                    if (lo_j < bs + w_base) lo_j = bs + w_base;
                    for (int j = hi_j; j >= lo_j; j--) w += Pw[idx[j]];
                }
                w += __shfl_xor_sync(FULL, w, 1);
                w += __shfl_xor_sync(FULL, w, 2);
                if (w_id >= 0 && (lane & 3) == 0) bw[w_id] = w;
            }
            Nbins += n_ne - 1;
            total_post = tp;
            __syncwarp();
            if (Nbins + 7 > a.oc.max_Nbins) break;
        }
        if (Nbins == 1) {  // :1028-1034
            double sw = 0;
            if (lane == 0) {
                for (int j = 0; j < N; j++) sw += Pw[idx[j]];
                bw[0] = sw;
            }
            __syncwarp();
        }
        // ---- output slots: exclusive scan of the per-bin output counts in bin-id order
        int curr = 0;
        for (int c0 = 0; c0 < Nbins; c0 += 32) {
            const int b = c0 + lane;
            int no = 0;
            if (b < Nbins) { const int np = bw[b] == 0 ? 0 : bnp[b]; no = np >= 2 ? 2 : np; }
            int incl = no;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (b < Nbins) bout[b] = (int16_t)(curr + incl - no);
            curr += __shfl_sync(FULL, incl, 31);
        }
        __syncwarp();
        // ---- compute_bin_props! (:646-700) + compute_new_particles! (:736-813 / :830-933); outputs go straight to the cell's first
        //      `curr` logical slots (all inputs are staged in shared memory, so nothing that is still needed is overwritten)
        const uint32_t c3 = (OP_MERGE & 0xFFu) | (a.substream << 8);
        for (int c0 = 0; c0 < Nbins; c0 += 32) {
            const int b = c0 + lane;
            int np = 0, bs = 0, be = -1, off = 0;
            double w = 0;
            if (b < Nbins) { w = bw[b]; np = w == 0 ? 0 : bnp[b]; bs = bstart[b]; be = bend[b]; off = bout[b]; }
            if (np > 2 && np <= 32) {  // one lane per bin, sequential in slice order like the reference
                const double inv_w = 1.0 / w;
                double m[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
                for (int j = bs; j <= be; j++) {
                    const int pj = idx[j];
                    const double pw = Pw[pj];
#pragma unroll
                    for (int d = 0; d < 6; d++) m[d] = m[d] + pw * P[(1 + d) * NS + pj];
                }
#pragma unroll
                for (int d = 0; d < 6; d++) m[d] *= inv_w;
                for (int j = bs; j <= be; j++) {
                    const int pj = idx[j];
                    const double pw = Pw[pj];
#pragma unroll
                    for (int d = 0; d < 6; d++) { const double dd = P[(1 + d) * NS + pj] - m[d]; s2[d] = s2[d] + pw * dd * dd; }
                }
                uint32_t rb[4];
                philox4x32_10((uint32_t)b, (uint32_t)cell, a.timestep, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rb);
                const int64_t p1 = mpos(q, off), p2 = mpos(q, off + 1);
                a.pv.a[F_W][p1] = 0.5 * w;
                a.pv.a[F_W][p2] = 0.5 * w;
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    const double sd = sqrt(s2[d] * inv_w);
                    const double sg = ((rb[0] >> d) & 1u) ? 1.0 : -1.0;
                    double x1 = m[d] + sg * sd, x2 = m[d] - sg * sd;
                    if (d == 3 && a.has_grid) {
                        x1 = x1 < a.min_x ? a.min_x : (x1 > a.max_x ? a.max_x : x1);
                        x2 = x2 < a.min_x ? a.min_x : (x2 > a.max_x ? a.max_x : x2);
                    }
                    a.pv.a[1 + d][p1] = x1;
                    a.pv.a[1 + d][p2] = x2;
                }
            } else if (np >= 1 && np <= 2) {
                for (int k = 0; k < np; k++) {
                    const int pj = idx[bs + k];
                    const int64_t p = mpos(q, off + k);
#pragma unroll
                    for (int f = 0; f < 7; f++) a.pv.a[f][p] = P[f * NS + pj];
                }
            }
            // bins of more than 32 particles: the whole warp on one bin at a time
            unsigned big = __ballot_sync(FULL, np > 32);
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const int Bb = c0 + src;
                const int Bs = __shfl_sync(FULL, bs, src), Be = __shfl_sync(FULL, be, src), Boff = __shfl_sync(FULL, off, src);
                const double Bw = __shfl_sync(FULL, w, src);
                const double inv_w = 1.0 / Bw;
                double m[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
                for (int j = Bs + lane; j <= Be; j += 32) {
                    const int pj = idx[j];
                    const double pw = Pw[pj];
#pragma unroll
                    for (int d = 0; d < 6; d++) m[d] += pw * P[(1 + d) * NS + pj];
                }
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    for (int o = 16; o > 0; o >>= 1) m[d] += __shfl_xor_sync(FULL, m[d], o);
                    m[d] *= inv_w;
                }
                for (int j = Bs + lane; j <= Be; j += 32) {
                    const int pj = idx[j];
                    const double pw = Pw[pj];
#pragma unroll
                    for (int d = 0; d < 6; d++) { const double dd = P[(1 + d) * NS + pj] - m[d]; s2[d] += pw * dd * dd; }
                }
#pragma unroll
                for (int d = 0; d < 6; d++)
                    for (int o = 16; o > 0; o >>= 1) s2[d] += __shfl_xor_sync(FULL, s2[d], o);
                if (lane == 0) {
                    uint32_t rb[4];
                    philox4x32_10((uint32_t)Bb, (uint32_t)cell, a.timestep, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rb);
                    const int64_t p1 = mpos(q, Boff), p2 = mpos(q, Boff + 1);
                    a.pv.a[F_W][p1] = 0.5 * Bw;
                    a.pv.a[F_W][p2] = 0.5 * Bw;
                    for (int d = 0; d < 6; d++) {
                        const double sd = sqrt(s2[d] * inv_w);
                        const double sg = ((rb[0] >> d) & 1u) ? 1.0 : -1.0;
                        double x1 = m[d] + sg * sd, x2 = m[d] - sg * sd;
                        if (d == 3 && a.has_grid) {
                            x1 = x1 < a.min_x ? a.min_x : (x1 > a.max_x ? a.max_x : x1);
                            x2 = x2 < a.min_x ? a.min_x : (x2 > a.max_x ? a.max_x : x2);
                        }
                        a.pv.a[1 + d][p1] = x1;
                        a.pv.a[1 + d][p2] = x2;
                    }
                }
            }
        }
        // ---- delete_particle_end! x n_delete (:800-812)
        const int n_del = N - curr;
        for (int j = curr + lane; j < N; j += 32) a.pv.a[F_W][mpos(q, j)] = 0.0;
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
            if (!(cell == a.n_cells_total) || n_del > q.n_group2) *a.noncontig = 1;  // :806-808
        }
      }
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_merge_octree_N2(mb_ctx* ctx, const mb_octree_params* oc, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi,
                                  int64_t species, int64_t threshold, int64_t target_np, const mb_grid1d* grid, uint32_t timestep,
                                  uint32_t substream) {
    MB_ARG(ctx && oc && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_ARG(target_np >= 1, "target_np");
    MB_ARG(oc->max_Nbins >= 8 && oc->max_depth >= 0, "OctreeN2Merge: max_Nbins >= 8");
    if (!(oc->split == 1 || oc->split == 2)) {
        set_error("OctreeBinMedianSplit is not supported (documented 'probably not fully correct' in the reference, merging_octree_N2.jl:444)");
        return MB_ERR_UNSUPPORTED;
    }
    MB_ARG(oc->init_bin_bounds >= 1 && oc->init_bin_bounds <= 3 && (oc->bin_bounds_compute == 1 || oc->bin_bounds_compute == 2), "octree enums");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    const int64_t nc = pia->n_cells, nr = cell_hi - cell_lo + 1, cap = pv->cap;
    ProfScope ps(ctx, PROF_MERGE);
    ctx->state_gen++;
    cudaStream_t st = ctx->stream;
    MergeArgs a;
    a.pv = pv->cur;
    a.ix = pia->d_indexer + (int64_t)s * nc;
    a.n_total = pia->d_n_total + s;
    a.cell_lo = cell_lo; a.cell_hi = cell_hi; a.n_cells_total = nc;
    a.threshold = threshold; a.target = target_np;
    a.oc = *oc;
    a.has_grid = grid != nullptr;
    a.min_x = grid ? grid->min_x : 0.0;
    a.max_x = grid ? grid->max_x : 0.0;
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, species, species);
    a.flags = ctx->d_flags;
    // index slices
    a.idx = (int32_t*)ctx_scratch(ctx, 0, (size_t)cap * 4);
    a.tmp = (int32_t*)ctx_scratch(ctx, 3, (size_t)cap * 4);
    a.oct = (uint8_t*)ctx_scratch(ctx, 8, (size_t)cap);
    int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)nr * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)(nr + 1) + gs_partial_count(nr)) * 8);
    if (!a.idx || !a.tmp || !a.oct || !cnt || !p64) return MB_ERR_CUDA;
    a.slice = p64;
    k_merge_counts<<<grid_for(nr, 256), 256, 0, st>>>(a.ix, cell_lo, nr, threshold, cnt);
    MB_LAUNCH_CHECK(ctx);
    int r = device_exclusive_scan(ctx, cnt, nr, p64, p64 + (nr + 1));
    if (r) return r;
    // per-CTA bin workspace
    const int64_t bcap = (oc->max_Nbins < target_np ? oc->max_Nbins : target_np) + 8;
    a.Bmax = (int)bcap;
    const int64_t avg = (pia->n_bound[s] > 0 ? pia->n_bound[s] : cap) / nc;
    // the refinement loop is run by warp 0 of each CTA, so many small CTAs (= concurrent cells) beat few large ones
    (void)avg;
    static const int env_threads = getenv("MB_MERGE_THREADS") ? atoi(getenv("MB_MERGE_THREADS")) : 0;   // experiment knobs
    static const int env_per_sm = getenv("MB_MERGE_PER_SM") ? atoi(getenv("MB_MERGE_PER_SM")) : 0;
    const int threads = env_threads > 0 ? env_threads : 128;
    const int per_sm_m = env_per_sm > 0 ? env_per_sm : 4;
    int64_t nCTA = nr < (int64_t)N_SM * per_sm_m ? nr : (int64_t)N_SM * per_sm_m;
    const size_t per = (size_t)bcap * (5 * 4 + 8 + 24 + 24 + 2 * 7 * 8) + 64;
    char* ws = (char*)ctx_scratch(ctx, 9, per * (size_t)nCTA + 4096);
    if (!ws) return MB_ERR_CUDA;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = ws + off; off += ((bytes + 255) / 256) * 256; return p; };
    a.b_w = (double*)take((size_t)nCTA * bcap * 8);
    a.b_vmin = (double*)take((size_t)nCTA * bcap * 24);
    a.b_vmax = (double*)take((size_t)nCTA * bcap * 24);
    a.outbuf = (double*)take((size_t)nCTA * bcap * 2 * 7 * 8);
    a.b_np = (int32_t*)take((size_t)nCTA * bcap * 4);
    a.b_depth = (int32_t*)take((size_t)nCTA * bcap * 4);
    a.b_start = (int32_t*)take((size_t)nCTA * bcap * 4);
    a.b_end = (int32_t*)take((size_t)nCTA * bcap * 4);
    a.b_out = (int32_t*)take((size_t)nCTA * (bcap + 1) * 4);
    if (off > per * (size_t)nCTA + 4096) {
        ws = (char*)ctx_scratch(ctx, 9, off + 4096);
        if (!ws) return MB_ERR_CUDA;
        set_error("internal: merge workspace sizing");
        return MB_ERR_UNSUPPORTED;
    }
    a.noncontig = pia->d_holes + s;
    if (!pia->contig_pending[s]) MB_CUDA(cudaMemsetAsync(a.noncontig, 0, sizeof(int), st));
    // small cells: one warp per cell in shared memory; the CTA kernel takes the rest.  The staging area is sized for 160 particles
    // when the threshold says that cells are merged long before they reach that size (a cell is merged as soon as it exceeds the
    // threshold, so it rarely exceeds it by much), else for 256; larger cells simply fall through to the CTA kernel.
    {
        const int ns = (threshold > 0 && threshold + 24 <= 160) ? 160 : NS_MAX;
        const int BCs = (int)(bcap < ns + 8 ? bcap : ns + 8);
        const size_t smem = (ns == 160 ? merge_warp_smem<160>(BCs) : merge_warp_smem<NS_MAX>(BCs)) * MW_WARPS;
        a.small_max = smem <= 200 * 1024 ? ns : 0;
        if (a.small_max) {
            static size_t attr_smem[64][2] = {{0}};  // function attributes are per device
            size_t& as = attr_smem[ctx->device & 63][ns == 160 ? 0 : 1];
            if (smem > as) {
                if (ns == 160) MB_CUDA(cudaFuncSetAttribute(k_merge_warp<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                else MB_CUDA(cudaFuncSetAttribute(k_merge_warp<NS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                as = smem;
            }
            int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
            if (per_sm > 4) per_sm = 4;  // 128 registers x 128 threads
            if (per_sm < 1) per_sm = 1;
            // a warp takes ch cells at a time: 32 when there are enough cells to keep every warp busy, fewer for short ranges
            const int64_t max_ctas = (int64_t)N_SM * per_sm;
            int ch = 32;
            while (ch > 1 && nr < max_ctas * MW_WARPS * ch) ch >>= 1;
            int64_t nW = (nr + (int64_t)ch * MW_WARPS - 1) / ((int64_t)ch * MW_WARPS);
            if (nW > max_ctas) nW = max_ctas;
            double* gb = (double*)ctx_scratch(ctx, 7, (size_t)nW * MW_WARPS * 6 * BCs * 8 + 256);
            if (!gb) return MB_ERR_CUDA;
            if (ns == 160) k_merge_warp<160><<<(int)nW, 32 * MW_WARPS, smem, st>>>(a, BCs, gb, ch);
            else k_merge_warp<NS_MAX><<<(int)nW, 32 * MW_WARPS, smem, st>>>(a, BCs, gb, ch);
            MB_LAUNCH_CHECK(ctx);
        }
    }
    {
        const size_t gsm = ((size_t)((bcap + 31) / 32) + 2) * 12 + 16;
        if (gsm > 160 * 1024) {
            set_error("merge_octree_N2_based!: max_Nbins / target_np too large for the arg-max group table");
            return MB_ERR_UNSUPPORTED;
        }
        static size_t attr_gsm[64] = {0};  // function attributes are per device
        size_t& ag = attr_gsm[ctx->device & 63];
        if (gsm > 40 * 1024 && gsm > ag) {
            MB_CUDA(cudaFuncSetAttribute(k_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
            ag = gsm;
        }
        a.cpb = threads;  // chunks of cells per CTA round: only as large as still leaves >= 8 chunks per CTA (load balance)
        while (a.cpb > 1 && nr < 8 * nCTA * a.cpb) a.cpb >>= 1;
        k_merge<<<(int)nCTA, threads, gsm, st>>>(a);
    }
    MB_LAUNCH_CHECK(ctx);
    // conservative on the host (operators dispatch on it); mb_pia_download resolves the exact reference value (:806-808)
    if (pia->contiguous[s]) pia->contig_pending[s] = 1;
    pia->contiguous[s] = 0;
    pia->sorted_layout[s] = 0;
    pia->h_valid = false;
    return MB_OK;
}

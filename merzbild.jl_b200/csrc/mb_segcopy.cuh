// Segment-wise copies whose segment sizes span five orders of magnitude (a Couette cell's 2 split particles ... a 0-D cell's 3e4
// particles): squash_pia! payload moves, the logical -> physical map of the squash folded into the sort, packing of the
// variable-weight split windows.  Two launches:
//   k_seg_small: a warp takes up to 32 consecutive segments at a time (one coalesced read of their descriptors) and copies the ones of up to
//                SEG_BIG elements lane-strided; larger segments are appended to a queue;
//   k_seg_big:   the queued segments, whole segments per CTA when they are many, cut into parts across CTAs when they are few.
// Desc::get(seg, n, src, dst) describes segment `seg`; Act::elem(seg, src, dst) moves one element; Act::seg(seg, n, src, dst) runs
// once per non-empty segment (lane 0).
#pragma once
#include "mb_common.cuh"
#include "mb_scan.cuh"

namespace mb {

constexpr int SEG_BIG = 4096;

template <class Desc, class Act>
__global__ void __launch_bounds__(256) k_seg_small(Desc D, Act A, int64_t nseg, int32_t* __restrict__ queue, int* __restrict__ qcount, int ch) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s0 = warp0 * ch; s0 < nseg; s0 += nwarps * ch) {
        const int64_t seg = s0 + lane;
        int64_t n = 0, src = 0, dst = 0;
        if (lane < ch && seg < nseg) D.get(seg, n, src, dst);
        if (n > SEG_BIG) {
            queue[atomicAdd(qcount, 1)] = (int32_t)seg;
            n = 0;
        } else if (n > 0) {
            A.seg(seg, n, src, dst);
        }
        unsigned todo = __ballot_sync(0xffffffffu, n > 0);
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t N = __shfl_sync(0xffffffffu, n, sl), S = __shfl_sync(0xffffffffu, src, sl), T = __shfl_sync(0xffffffffu, dst, sl);
            const int64_t sg = s0 + sl;
            for (int64_t j = lane; j < N; j += 32) A.elem(sg, S + j, T + j);
        }
    }
}
template <class Desc, class Act>
__global__ void __launch_bounds__(256) k_seg_big(Desc D, Act A, const int32_t* __restrict__ queue, const int* __restrict__ qcount) {
    const int nq = *qcount;
    if (nq == 0) return;
    // work unit = (queued segment, part): with few big segments every segment is cut into gridDim.x / nq parts, with many each CTA
    // takes whole segments
    const int parts = (int)gridDim.x / nq > 1 ? (int)gridDim.x / nq : 1;
    for (int64_t u = blockIdx.x; u < (int64_t)nq * parts; u += gridDim.x) {
        const int qi = (int)(u / parts), part = (int)(u - (int64_t)qi * parts);
        const int64_t seg = queue[qi];
        int64_t n, src, dst;
        D.get(seg, n, src, dst);
        const int64_t chunk = (n + parts - 1) / parts;
        const int64_t j0 = part * chunk, j1 = j0 + chunk < n ? j0 + chunk : n;
        if (part == 0 && threadIdx.x == 0) A.seg(seg, n, src, dst);
        for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) A.elem(seg, src + j, dst + j);
    }
}

// queue storage: capacity / SEG_BIG + 2 entries and a counter, in scratch slot `slot`
static inline int seg_queue(mb_ctx* ctx, int slot, int64_t capacity, int32_t** queue, int** qcount) {
    const size_t nq = (size_t)(capacity / SEG_BIG) + 4;
    int32_t* p = (int32_t*)ctx_scratch(ctx, slot, (nq + 4) * 4);
    if (!p) return MB_ERR_CUDA;
    *qcount = (int*)p;
    *queue = p + 4;
    MB_CUDA(cudaMemsetAsync(p, 0, 4, ctx->stream));
    return MB_OK;
}
template <class Desc, class Act>
static inline int seg_copy(mb_ctx* ctx, int qslot, int64_t capacity, int64_t nseg, const Desc& D, const Act& A) {
    int32_t* queue;
    int* qcount;
    int r = seg_queue(ctx, qslot, capacity, &queue, &qcount);
    if (r) return r;
    // a warp takes ch segments at a time: 32 when there are enough segments to keep every warp busy, fewer otherwise
    int ch = 32;
    while (ch > 1 && nseg < (int64_t)N_SM * 8 * 8 * ch) ch >>= 1;
    k_seg_small<<<grid_for((nseg + ch - 1) / ch * 32, 256, 8), 256, 0, ctx->stream>>>(D, A, nseg, queue, qcount, ch);
    MB_LAUNCH_CHECK(ctx);
    k_seg_big<<<N_SM * 4, 256, 0, ctx->stream>>>(D, A, queue, qcount);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

// the segments of squash_pia!: group 1 of all cells, then group 2 of all cells (particles.jl:622-682); newlo = exclusive scan of the sizes
struct SquashDesc {
    const Indexer* ix;
    int64_t nc;
    const int64_t* newlo;
    int* flags;
    __device__ __forceinline__ void get(int64_t seg, int64_t& n, int64_t& src, int64_t& dst) const {
        const bool g2 = seg >= nc;
        const Indexer q = ix[g2 ? seg - nc : seg];
        n = g2 ? q.n_group2 : q.n_group1;
        src = (g2 ? q.start2 : q.start1) - 1;
        dst = newlo[seg];
        if (n > 0 && src < dst) {  // the reference only ever shifts left (particles.jl:641,659,672: `if offset > 0`)
            atomicOr(&flags[0], DEVERR_PRECONDITION);
            n = 0;
        }
    }
};

// The map logical (squashed) position -> physical position of a non-contiguous species: group 1 of all cells, group 2 of all cells
// (the order squash_pia! produces, particles.jl:622-682), then -- after a slab exchange on a non-contiguous layout -- the arrivals,
// which mb_exchange_slab parks at the END of the capacity (they are in no indexer yet).  The sort's general path and the exchange's
// pack read the particles through this map instead of moving the payload twice.
struct SrcMapDesc {
    const Indexer* ix;
    int64_t nc;
    const int64_t* newlo;
    const int64_t* n_arr;  // nullable
    int64_t cap;
    int* flags;
    __device__ __forceinline__ void get(int64_t seg, int64_t& n, int64_t& src, int64_t& dst) const {
        const int64_t na = n_arr ? *n_arr : 0;
        dst = newlo[seg];
        if (seg == 2 * nc) { n = na; src = cap - na; return; }
        const bool g2 = seg >= nc;
        const Indexer q = ix[g2 ? seg - nc : seg];
        n = g2 ? q.n_group2 : q.n_group1;
        src = (g2 ? q.start2 : q.start1) - 1;
        if (n > 0 && src + n > cap - na) atomicOr(&flags[0], DEVERR_CAPACITY);  // the live particles reach into the parked arrivals
    }
};
struct SrcMapAct {
    int32_t* srcmap;
    __device__ __forceinline__ void seg(int64_t, int64_t, int64_t, int64_t) const {}
    __device__ __forceinline__ void elem(int64_t, int64_t src, int64_t dst) const { srcmap[dst] = (int32_t)src; }
};
static __global__ void k_srcmap_counts(const Indexer* __restrict__ ix, int64_t nc, const int64_t* n_arr, int32_t* __restrict__ cnt) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        const Indexer q = ix[c];
        cnt[c] = (int32_t)q.n_group1;
        cnt[nc + c] = (int32_t)q.n_group2;
        if (c == 0) cnt[2 * nc] = n_arr ? (int32_t)*n_arr : 0;
    }
}
// scratch slots: 8 (map), 4 (counts), 5 (offsets), 7 (queue of the big segments)
static inline int build_src_map(mb_ctx* ctx, int64_t cap, const Indexer* ix, int64_t nc, const int64_t* d_n_arr, int32_t** src_out) {
    int32_t* src = (int32_t*)ctx_scratch(ctx, 8, (size_t)cap * 4);
    int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)(2 * nc + 1) * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)(2 * nc + 2) + gs_partial_count(2 * nc + 1)) * 8);
    if (!src || !cnt || !p64) return MB_ERR_CUDA;
    ProfScope ps(ctx, PROF_SQUASH);
    k_srcmap_counts<<<grid_for(nc, 256), 256, 0, ctx->stream>>>(ix, nc, d_n_arr, cnt);
    MB_LAUNCH_CHECK(ctx);
    int r = device_exclusive_scan(ctx, cnt, 2 * nc + 1, p64, p64 + (2 * nc + 2));
    if (r) return r;
    SrcMapDesc D{ix, nc, p64, d_n_arr, cap, ctx->d_flags};
    SrcMapAct A{src};
    r = seg_copy(ctx, 7, cap, 2 * nc + 1, D, A);
    if (r) return r;
    *src_out = src;
    return MB_OK;
}

}  // namespace mb

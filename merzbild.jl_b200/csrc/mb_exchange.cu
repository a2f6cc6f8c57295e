// Slab exchange over NCCL: the B200 replacement of the reference's shared-memory chunk exchange (src/parallel.jl:21-581:
// ChunkExchanger, exchange_particles!, push_particles!, sort_particles_after_exchange!).  Every GPU owns a contiguous slab of
// cells (mb_grid1d_slab, the ChunkSplitters rule); after convection the particles whose cell lies outside the slab are packed
// (stable, in logical order) into a left and a right send buffer, the neighbours swap counts and payloads with
// ncclSend/ncclRecv in one group over NVLink, and the arrivals are appended after n_total (left arrivals first).  The
// leavers stay where they are and are dropped by the next sort_particles! (their cell is outside the slab), which also
// places the arrivals -- they are just more keys -- so no hole filling or free-list bookkeeping is needed (the reference's
// swap-vs-push distinction, parallel.jl:373-413, exists only because its sort moves indices, not particles).
//
// NCCL is loaded lazily with dlopen (libnccl.so.2, e.g. the one PyTorch bundles): a single-GPU user never needs it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "mb_common.cuh"
#include "mb_scan.cuh"
#include "mb_segcopy.cuh"

namespace mb {

// ---- minimal NCCL binding (nccl.h, 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclFloat64 = 8 };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.h) return MB_OK;
    const char* env = getenv("MB_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL not found: dlopen(libnccl.so.2) failed (import torch first, or set MB_NCCL_LIB / LD_LIBRARY_PATH)");
        return MB_ERR_NCCL;
    }
#define MB_SYM(field, name)                                            \
    *(void**)(&g_nccl.field) = dlsym(h, name);                         \
    if (!g_nccl.field) { set_error(std::string("NCCL symbol missing: ") + name); return MB_ERR_NCCL; }
    MB_SYM(GetUniqueId, "ncclGetUniqueId");
    MB_SYM(CommInitRank, "ncclCommInitRank");
    MB_SYM(CommDestroy, "ncclCommDestroy");
    MB_SYM(Send, "ncclSend");
    MB_SYM(Recv, "ncclRecv");
    MB_SYM(AllReduce, "ncclAllReduce");
    MB_SYM(GroupStart, "ncclGroupStart");
    MB_SYM(GroupEnd, "ncclGroupEnd");
    MB_SYM(GetErrorString, "ncclGetErrorString");
#undef MB_SYM
    g_nccl.h = h;
    return MB_OK;
}
#define MB_NCCL(x)                                                                        \
    do {                                                                                  \
        int e__ = (x);                                                                    \
        if (e__ != ncclSuccess) {                                                         \
            set_error(std::string("NCCL error in " #x ": ") + g_nccl.GetErrorString(e__)); \
            return MB_ERR_NCCL;                                                           \
        }                                                                                 \
    } while (0)

// sum over the ranks of the context's communicator (SurfProps: 22 doubles; surface_props.jl:232-252 across GPUs)
int nccl_allreduce_sum_f64(mb_ctx* ctx, double* buf, size_t n) {
    if (!ctx->nccl_comm) {
        set_error("reduction across ranks: call mb_comm_init first");
        return MB_ERR_NCCL;
    }
    MB_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat64, /* ncclSum */ 0, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return MB_OK;
}

constexpr int XB = 256;      // threads per block
constexpr int XI = 8;        // particles per thread (blocked)
constexpr int XT = XB * XI;  // particles per block

// direction of a particle: 0 stays, 1 leaves to the left neighbour, 2 to the right
__device__ __forceinline__ int xch_dir(double x, double inv_dx, int64_t cell_offset, int64_t n_cells) {
    const int64_t c = (int64_t)floor(x * inv_dx) - cell_offset;
    return c < 0 ? 1 : (c >= n_cells ? 2 : 0);
}

// src (nullable): logical -> physical position of a non-contiguous species (build_src_map)
static __global__ void __launch_bounds__(XB) k_xch_count(const double* __restrict__ X, const int64_t* n_total_p, double inv_dx, int64_t cell_offset,
                                                        int64_t n_cells, int32_t* __restrict__ cntL, int32_t* __restrict__ cntR,
                                                        const int32_t* __restrict__ src) {
    __shared__ int sl[XB / 32], sr[XB / 32];
    const int64_t n = *n_total_p;
    const int64_t base = (int64_t)blockIdx.x * XT + (int64_t)threadIdx.x * XI;
    int l = 0, r = 0;
    for (int k = 0; k < XI; k++) {
        const int64_t i = base + k;
        if (i < n) {
            const int d = xch_dir(X[src ? (int64_t)src[i] : i], inv_dx, cell_offset, n_cells);
            l += d == 1;
            r += d == 2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) { l += __shfl_down_sync(0xffffffffu, l, o); r += __shfl_down_sync(0xffffffffu, r, o); }
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = l; sr[threadIdx.x >> 5] = r; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int i = 0; i < XB / 32; i++) { a += sl[i]; b += sr[i]; }
        cntL[blockIdx.x] = a;
        cntR[blockIdx.x] = b;
    }
}

// stable pack (logical order) of the leavers into particle-major send buffers (7 doubles per particle)
static __global__ void __launch_bounds__(XB) k_xch_pack(SoA pv, const int64_t* n_total_p, double inv_dx, int64_t cell_offset, int64_t n_cells,
                                                       const int64_t* __restrict__ offL, const int64_t* __restrict__ offR, double* __restrict__ sendL,
                                                       double* __restrict__ sendR, int64_t cap, int64_t nblocks, int64_t* counts, int* flags,
                                                       const int32_t* __restrict__ src) {
    __shared__ int sl[XB], sr[XB];
    const int64_t n = *n_total_p;
    const int64_t base = (int64_t)blockIdx.x * XT + (int64_t)threadIdx.x * XI;
    int dir[XI];
    int l = 0, r = 0;
#pragma unroll
    for (int k = 0; k < XI; k++) {
        const int64_t i = base + k;
        dir[k] = i < n ? xch_dir(pv.a[F_X][src ? (int64_t)src[i] : i], inv_dx, cell_offset, n_cells) : 0;
        l += dir[k] == 1;
        r += dir[k] == 2;
    }
    sl[threadIdx.x] = l;
    sr[threadIdx.x] = r;
    __syncthreads();
    for (int o = 1; o < XB; o <<= 1) {  // inclusive scans over the block
        const int a = threadIdx.x >= o ? sl[threadIdx.x - o] : 0, b = threadIdx.x >= o ? sr[threadIdx.x - o] : 0;
        __syncthreads();
        sl[threadIdx.x] += a;
        sr[threadIdx.x] += b;
        __syncthreads();
    }
    int64_t pl = offL[blockIdx.x] + sl[threadIdx.x] - l, pr = offR[blockIdx.x] + sr[threadIdx.x] - r;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t tl = offL[nblocks], tr = offR[nblocks];
        counts[0] = tl < cap ? tl : cap;  // what the send buffer holds; layout [nL, cap, nR, cap]: one 2-word message per neighbour
        counts[1] = cap;
        counts[2] = tr < cap ? tr : cap;
        counts[3] = cap;
        if (tl > cap || tr > cap) {
            atomicOr(&flags[0], DEVERR_CAPACITY);
            flags[1] = (int)(tl > tr ? tl : tr);
        }
    }
#pragma unroll
    for (int k = 0; k < XI; k++) {
        if (dir[k] == 0) continue;
        const int64_t i = src ? (int64_t)src[base + k] : base + k;
        double* dst;
        if (dir[k] == 1) { if (pl >= cap) { pl++; continue; } dst = sendL + 7 * pl; pl++; }
        else { if (pr >= cap) { pr++; continue; } dst = sendR + 7 * pr; pr++; }
#pragma unroll
        for (int f = 0; f < 7; f++) dst[f] = pv.a[f][i];
    }
}

// arrivals (left neighbour's first, then the right neighbour's) are appended after n_total
// at_end: the species is not contiguous -- the arrivals are parked at [cap - nL - nR, cap) (SrcMapDesc checks that no live particle is there)
static __global__ void __launch_bounds__(256) k_xch_unpack(SoA pv, int64_t* n_total_p, int64_t cap, const double* __restrict__ recvL, int64_t nL,
                                                          const double* __restrict__ recvR, int64_t nR, int* flags, int at_end) {
    const int64_t n0 = at_end ? cap - nL - nR : *n_total_p;
    if ((at_end ? *n_total_p : n0) + nL + nR > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(&flags[0], DEVERR_CAPACITY);
            flags[1] = (int)(*n_total_p + nL + nR);
        }
        return;
    }
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nL + nR; t += (int64_t)gridDim.x * blockDim.x) {
        const double* src = t < nL ? recvL + 7 * t : recvR + 7 * (t - nL);
#pragma unroll
        for (int f = 0; f < 7; f++) pv.a[f][n0 + t] = src[f];
    }
}
// ------------------------------------------------------------------------------------------------ edge exchange
// On a sorted layout whose particles move at most w cells per step (the sort's band assumption) the leavers to the left can only
// sit in the first w cells of the slab and the leavers to the right in the last w: two CTAs scan those few thousand particles,
// pack the leavers (stable, logical order) behind a count header, and the neighbours swap FIXED-size messages -- no host
// round trip, no scan over all particles.  A leaver from any other cell is a band overflow: the fused convect kernel (F_FAR)
// or the sort's classify pass reports it as a device error instead of losing the particle silently.
constexpr int XE_CAP = 8192;             // leavers per direction and step
constexpr int XE_MSG = 8 + 7 * XE_CAP;   // doubles per message: [0] = count (int64 bits), payload from [8] (64-byte aligned)

// Works on any layout the indexers describe (sorted, or group 2 at the tail / holes after a merge): the ranges of the w cells next
// to the face are scanned cell by cell, group 1 then group 2, so the leavers are packed in (cell, group, position) order.
static __global__ void __launch_bounds__(256) k_xch_edge_pack(SoA pv, const Indexer* __restrict__ ix, int64_t n_cells, int w, double inv_dx,
                                                            int64_t cell_offset, double* __restrict__ sendL, double* __restrict__ sendR, int hasL,
                                                            int hasR, int* flags, int64_t* __restrict__ sent2) {
    __shared__ int s_w[8];
    const int side = blockIdx.x;  // 0: left face, 1: right face
    const int want = side + 1;
    double* __restrict__ send = side == 0 ? sendL : sendR;
    const int64_t c_lo = side == 0 ? 0 : (n_cells - w > 0 ? n_cells - w : 0);
    const int64_t c_hi = side == 0 ? (w < n_cells ? w : n_cells) : n_cells;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
    for (int64_t c = c_lo; c < c_hi; c++) {
        const Indexer q = ix[c];
        for (int g = 0; g < 2; g++) {
            const int64_t lo = (g == 0 ? q.start1 : q.start2) - 1;
            const int64_t hi = lo + (g == 0 ? q.n_group1 : q.n_group2);
            for (int64_t i0 = lo; i0 < hi; i0 += 256) {  // block-uniform bounds
                const int64_t i = i0 + threadIdx.x;
                const bool is = i < hi && xch_dir(pv.a[F_X][i], inv_dx, cell_offset, n_cells) == want;
                const unsigned bal = __ballot_sync(0xffffffffu, is);
                if (lane == 0) s_w[wid] = __popc(bal);
                __syncthreads();
                int off = 0, tot = 0;
                for (int k = 0; k < 8; k++) { if (k < wid) off += s_w[k]; tot += s_w[k]; }
                const int pos = base + off + __popc(bal & lt);
                if (is && pos < XE_CAP) {
#pragma unroll
                    for (int f = 0; f < 7; f++) send[8 + 7 * (int64_t)pos + f] = pv.a[f][i];
                }
                base += tot;
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0) {
        send[0] = __longlong_as_double((long long)(base < XE_CAP ? base : XE_CAP));
        sent2[side] = base;  // the sort that drops the leavers checks that it dropped exactly the particles sent
        if (base > XE_CAP) { atomicOr(&flags[0], DEVERR_CAPACITY); flags[1] = base; }
        if (base > 0 && !(side == 0 ? hasL : hasR)) atomicOr(&flags[0], DEVERR_PRECONDITION);  // left the global domain
        if (side == 0 && flags[F_FAR] != 0) atomicOr(&flags[0], DEVERR_BAND_OVERFLOW);
    }
}

static __global__ void __launch_bounds__(256) k_xch_edge_unpack(SoA pv, const int64_t* n_total_p, int64_t cap, const double* __restrict__ recvL,
                                                              const double* __restrict__ recvR, int hasL, int hasR, int* flags, int at_end) {
    int64_t nL = hasL ? (int64_t)__double_as_longlong(recvL[0]) : 0, nR = hasR ? (int64_t)__double_as_longlong(recvR[0]) : 0;
    nL = nL < 0 ? 0 : (nL > XE_CAP ? XE_CAP : nL);
    nR = nR < 0 ? 0 : (nR > XE_CAP ? XE_CAP : nR);
    if (*n_total_p + nL + nR > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&flags[0], DEVERR_CAPACITY); flags[1] = (int)(*n_total_p + nL + nR); }
        return;
    }
    const int64_t n0 = at_end ? cap - nL - nR : *n_total_p;  // non-contiguous species: parked at the end of the capacity
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nL + nR; t += (int64_t)gridDim.x * blockDim.x) {
        const double* src = t < nL ? recvL + 8 + 7 * t : recvR + 8 + 7 * (t - nL);
#pragma unroll
        for (int f = 0; f < 7; f++) pv.a[f][n0 + t] = src[f];
    }
}
// after the unpack (every block of it read the old n_total): n_total += arrivals, and the sort learns their number
static __global__ void k_xch_edge_commit(int64_t* n_total_p, int64_t cap, const double* recvL, const double* recvR, int hasL, int hasR,
                                         int64_t* d_n_arr) {
    int64_t nL = hasL ? (int64_t)__double_as_longlong(recvL[0]) : 0, nR = hasR ? (int64_t)__double_as_longlong(recvR[0]) : 0;
    nL = nL < 0 ? 0 : (nL > XE_CAP ? XE_CAP : nL);
    nR = nR < 0 ? 0 : (nR > XE_CAP ? XE_CAP : nR);
    if (*n_total_p + nL + nR <= cap) { *n_total_p += nL + nR; *d_n_arr += nL + nR; }
}
static __global__ void k_add_i64(int64_t* p, int64_t v) { *p += v; }

static __global__ void k_xch_add_total(int64_t* n_total_p, int64_t cap, int64_t add) {
    if (*n_total_p + add <= cap) *n_total_p += add;
}


// ------------------------------------------------------------------------------------------------ host phases
// mb_exchange_slab (one rank per process, NCCL transport) and mb_exchange_chunks (all chunks in one process, copy transport) share
// these: xch_begin packs the leavers, the transport moves the staging buffers, xch_finish_* appends the arrivals.
struct XchPlan {
    bool edge, hasL, hasR, at_end;
    int left, right, s;
};
struct XchCounts {
    int64_t sL, sR, rL, rR;   // particles actually moved (clamped to the receiver's staging capacity)
    bool clamped;
};

static int xch_begin(mb_ctx* ctx, const mb_grid1d* slab, mb_pv* pv, mb_pia* pia, int64_t species, int rank, int nranks, bool want_counts,
                     XchPlan& P) {
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(pia->n_cells == slab->n_cells, "slab.n_cells != pia.n_cells");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    if (!pia->contiguous[s] && pv->n_arrivals != 0) {
        set_error("mb_exchange_slab on a non-contiguous species with arrivals pending: sort first");
        return MB_ERR_PRECONDITION;
    }
    {   // the own particles are not touched: a classification cached by the fused convect kernel stays valid
        const bool keep = ctx->cls_gen == ctx->state_gen;
        ctx->state_gen++;
        if (keep) ctx->cls_gen = ctx->state_gen;
    }
    cudaStream_t st = ctx->stream;
    // staging buffers: capacity / 8 particles per direction
    const size_t want = (size_t)pv->cap / 8 + 65536;
    if (ctx->xch_cap < want) {
        MB_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 2; i++) {
            if (ctx->xch_send[i]) cudaFree(ctx->xch_send[i]);
            if (ctx->xch_recv[i]) cudaFree(ctx->xch_recv[i]);
            MB_CUDA(cudaMalloc(&ctx->xch_send[i], want * 56));
            MB_CUDA(cudaMalloc(&ctx->xch_recv[i], want * 56));
        }
        ctx->xch_cap = want;
    }
    P.s = s;
    P.left = rank - 1;
    P.right = rank + 1;
    P.hasL = nranks > 1 && P.left >= 0;
    P.hasR = nranks > 1 && P.right < nranks;
    // edge exchange: every rank takes the same decision (it only depends on the operator sequence)
    P.edge = ctx->xch_mode == 0 && ctx->band_w > 0 && pv->n_arrivals == 0 && !want_counts;
    P.at_end = !pia->contiguous[s];
    int64_t* d_nt = pia->d_n_total + s;
    if (P.edge) {
        k_xch_edge_pack<<<2, 256, 0, st>>>(pv->cur, pia->d_indexer + (int64_t)s * pia->n_cells, pia->n_cells, ctx->band_w, slab->inv_dx,
                                         slab->cell_offset, (double*)ctx->xch_send[0], (double*)ctx->xch_send[1], P.hasL ? 1 : 0, P.hasR ? 1 : 0,
                                         ctx->d_flags, ctx->d_xch_counts + 8);
        MB_LAUNCH_CHECK(ctx);
        return MB_OK;
    }
    const int64_t nb_part = pia->n_bound[s] > 0 ? pia->n_bound[s] : pv->cap;
    const int64_t nblocks = (nb_part + XT - 1) / XT;
    int32_t* cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)(2 * nblocks) * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)2 * (nblocks + 1) + gs_partial_count(nblocks)) * 8);
    if (!cnt || !p64) return MB_ERR_CUDA;
    int64_t* offL = p64;
    int64_t* offR = p64 + (nblocks + 1);
    int64_t* partial = p64 + 2 * (nblocks + 1);
    // a non-contiguous species (after a merge) is read through the logical -> physical map instead of being squashed first
    int32_t* src = nullptr;
    if (P.at_end) {
        int rr = build_src_map(ctx, pv->cap, pia->d_indexer + (int64_t)s * pia->n_cells, pia->n_cells, nullptr, &src);
        if (rr) return rr;
        cnt = (int32_t*)ctx_scratch(ctx, 13, (size_t)(2 * nblocks) * 4);  // (slots 4 / 5 were used by the map)
        p64 = (int64_t*)ctx_scratch(ctx, 14, ((size_t)2 * (nblocks + 1) + gs_partial_count(nblocks)) * 8);
        if (!cnt || !p64) return MB_ERR_CUDA;
        offL = p64; offR = p64 + (nblocks + 1); partial = p64 + 2 * (nblocks + 1);
    }
    k_xch_count<<<(int)nblocks, XB, 0, st>>>(pv->cur.a[F_X], d_nt, slab->inv_dx, slab->cell_offset, slab->n_cells, cnt, cnt + nblocks, src);
    MB_LAUNCH_CHECK(ctx);
    int r = device_exclusive_scan(ctx, cnt, nblocks, offL, partial);
    if (r) return r;
    r = device_exclusive_scan(ctx, cnt + nblocks, nblocks, offR, partial);
    if (r) return r;
    k_xch_pack<<<(int)nblocks, XB, 0, st>>>(pv->cur, d_nt, slab->inv_dx, slab->cell_offset, slab->n_cells, offL, offR, (double*)ctx->xch_send[0],
                                            (double*)ctx->xch_send[1], (int64_t)ctx->xch_cap, nblocks, ctx->d_xch_counts, ctx->d_flags, src);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

// h: [0] nL [1] my cap [2] nR [3] my cap [4] rL [5] capL [6] rR [7] capR.  Both sides of a face compute the same number.
static XchCounts xch_negotiate(const int64_t* h, int64_t my_cap, const XchPlan& P) {
    XchCounts K;
    K.sL = P.hasL ? (h[0] < h[5] ? h[0] : h[5]) : 0;
    K.sR = P.hasR ? (h[2] < h[7] ? h[2] : h[7]) : 0;
    K.rL = P.hasL ? (h[4] < my_cap ? h[4] : my_cap) : 0;
    K.rR = P.hasR ? (h[6] < my_cap ? h[6] : my_cap) : 0;
    K.clamped = (P.hasL && (K.sL != h[0] || K.rL != h[4])) || (P.hasR && (K.sR != h[2] || K.rR != h[6]));
    return K;
}

static int xch_finish_edge(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, const XchPlan& P) {
    cudaStream_t st = ctx->stream;
    int64_t* d_nt = pia->d_n_total + P.s;
    if (P.hasL || P.hasR) {
        const double* rL = (const double*)ctx->xch_recv[0];
        const double* rR = (const double*)ctx->xch_recv[1];
        k_xch_edge_unpack<<<16, 256, 0, st>>>(pv->cur, d_nt, pv->cap, rL, rR, P.hasL ? 1 : 0, P.hasR ? 1 : 0, ctx->d_flags, P.at_end ? 1 : 0);
        MB_LAUNCH_CHECK(ctx);
        k_xch_edge_commit<<<1, 1, 0, st>>>(d_nt, pv->cap, rL, rR, P.hasL ? 1 : 0, P.hasR ? 1 : 0, pv->d_n_arr);
        MB_LAUNCH_CHECK(ctx);
        pv->n_arrivals += 2 * XE_CAP;  // upper bound; the exact number stays on the device
        pv->arrivals_at_end = P.at_end ? 1 : 0;
    }
    pv->drop_oob = 2;
    pia->h_valid = false;
    pia->n_bound[P.s] = pv->cap;
    return MB_OK;
}

static int xch_finish_full(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, const XchPlan& P, const XchCounts& K, int64_t* n_sent2, int64_t* n_recv2) {
    cudaStream_t st = ctx->stream;
    int64_t* d_nt = pia->d_n_total + P.s;
    const int64_t* h = ctx->h_xch_counts;
    if (K.rL + K.rR > 0) {
        k_xch_unpack<<<grid_for(K.rL + K.rR, 256), 256, 0, st>>>(pv->cur, d_nt, pv->cap, (const double*)ctx->xch_recv[0], K.rL,
                                                               (const double*)ctx->xch_recv[1], K.rR, ctx->d_flags, P.at_end ? 1 : 0);
        MB_LAUNCH_CHECK(ctx);
        k_xch_add_total<<<1, 1, 0, st>>>(d_nt, pv->cap, K.rL + K.rR);
        MB_LAUNCH_CHECK(ctx);
        k_add_i64<<<1, 1, 0, st>>>(pv->d_n_arr, K.rL + K.rR);
        MB_LAUNCH_CHECK(ctx);
    }
    // the layout of the own particles is untouched: the next sort drops the leavers and merges the arrivals (band path if sorted)
    if (h[0] + h[2] > 0 && pv->drop_oob == 0) pv->drop_oob = 1;
    pv->n_arrivals += K.rL + K.rR;
    if (K.rL + K.rR > 0) pv->arrivals_at_end = P.at_end ? 1 : 0;
    pia->h_valid = false;
    pia->n_bound[P.s] = pv->cap;
    if (n_sent2) { n_sent2[0] = K.sL; n_sent2[1] = K.sR; }
    if (n_recv2) { n_recv2[0] = K.rL; n_recv2[1] = K.rR; }
    // errors are reported only now, after every neighbour got the messages it was waiting for
    if ((!P.hasL && h[0] > 0) || (!P.hasR && h[2] > 0)) {
        set_error("mb_exchange_slab: particles left the global domain (convect must clamp them to [min_x, max_x])");
        return MB_ERR_PRECONDITION;
    }
    if (K.clamped) {
        set_error("mb_exchange_slab: more leavers than the staging buffers hold (capacity / 8 + 65536 particles per direction): particles were lost");
        return MB_ERR_CAPACITY;
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_comm_unique_id(void* out128) {
    MB_ARG(out128 != nullptr, "NULL");
    int r = nccl_load();
    if (r) return r;
    ncclUniqueId id;
    MB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return MB_OK;
}

int mb_comm_init(mb_ctx* ctx, const void* id128, int rank, int nranks) {
    MB_ARG(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init");
    int r = nccl_load();
    if (r) return r;
    MB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    MB_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return MB_OK;
}

int mb_exchange_set_mode(mb_ctx* ctx, int32_t mode) {
    MB_ARG(ctx && (mode == 0 || mode == 1), "exchange mode must be 0 (edge exchange when possible) or 1 (always the full exchange)");
    ctx->xch_mode = mode;
    return MB_OK;
}

int mb_exchange_slab(mb_ctx* ctx, const mb_grid1d* slab, mb_pv* pv, mb_pia* pia, int64_t species, int64_t* n_sent2, int64_t* n_recv2) {
    MB_ARG(ctx && slab && pv && pia, "NULL handle");
    if (ctx->nranks > 1 && !ctx->nccl_comm) {
        set_error("mb_exchange_slab: call mb_comm_init first");
        return MB_ERR_NCCL;
    }
    ProfScope ps(ctx, PROF_EXCHANGE);
    XchPlan P;
    int r = xch_begin(ctx, slab, pv, pia, species, ctx->rank, ctx->nranks, n_sent2 || n_recv2, P);
    if (r) return r;
    cudaStream_t st = ctx->stream;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    if (P.edge) {
        if (P.hasL || P.hasR) {
            MB_NCCL(g_nccl.GroupStart());
            if (P.hasL) {
                MB_NCCL(g_nccl.Send(ctx->xch_send[0], XE_MSG, ncclFloat64, P.left, comm, st));
                MB_NCCL(g_nccl.Recv(ctx->xch_recv[0], XE_MSG, ncclFloat64, P.left, comm, st));
            }
            if (P.hasR) {
                MB_NCCL(g_nccl.Send(ctx->xch_send[1], XE_MSG, ncclFloat64, P.right, comm, st));
                MB_NCCL(g_nccl.Recv(ctx->xch_recv[1], XE_MSG, ncclFloat64, P.right, comm, st));
            }
            MB_NCCL(g_nccl.GroupEnd());
        }
        return xch_finish_edge(ctx, pv, pia, P);
    }
    // full exchange: the neighbours swap (count, receive capacity), then the payloads.  Between the two phases NOTHING returns early:
    // a rank that bailed out here would leave its neighbours blocked in their payload receive.  Local errors are collected and
    // reported after the payload phase; counts are clamped to what the receiver can take, identically on both sides.
    int64_t* dc = ctx->d_xch_counts;  // [0] nL [1] cap [2] nR [3] cap | received: [4] rL [5] capL [6] rR [7] capR
    MB_CUDA(cudaMemsetAsync(dc + 4, 0, 4 * 8, st));
    if (P.hasL || P.hasR) {
        MB_NCCL(g_nccl.GroupStart());
        if (P.hasL) {
            MB_NCCL(g_nccl.Send(dc + 0, 2, ncclInt64, P.left, comm, st));
            MB_NCCL(g_nccl.Recv(dc + 4, 2, ncclInt64, P.left, comm, st));
        }
        if (P.hasR) {
            MB_NCCL(g_nccl.Send(dc + 2, 2, ncclInt64, P.right, comm, st));
            MB_NCCL(g_nccl.Recv(dc + 6, 2, ncclInt64, P.right, comm, st));
        }
        MB_NCCL(g_nccl.GroupEnd());
    }
    MB_CUDA(cudaMemcpyAsync(ctx->h_xch_counts, dc, 8 * 8, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));  // the payload sizes must be known on the host to post the receives (no error flags consumed here)
    XchCounts K = xch_negotiate(ctx->h_xch_counts, (int64_t)ctx->xch_cap, P);
    if (P.hasL || P.hasR) {
        MB_NCCL(g_nccl.GroupStart());
        if (P.hasL) {
            if (K.sL > 0) MB_NCCL(g_nccl.Send(ctx->xch_send[0], (size_t)K.sL * 7, ncclFloat64, P.left, comm, st));
            if (K.rL > 0) MB_NCCL(g_nccl.Recv(ctx->xch_recv[0], (size_t)K.rL * 7, ncclFloat64, P.left, comm, st));
        }
        if (P.hasR) {
            if (K.sR > 0) MB_NCCL(g_nccl.Send(ctx->xch_send[1], (size_t)K.sR * 7, ncclFloat64, P.right, comm, st));
            if (K.rR > 0) MB_NCCL(g_nccl.Recv(ctx->xch_recv[1], (size_t)K.rR * 7, ncclFloat64, P.right, comm, st));
        }
        MB_NCCL(g_nccl.GroupEnd());
    }
    return xch_finish_full(ctx, pv, pia, P, K, n_sent2, n_recv2);
}

/* exchange_particles!(exchanger, pv_chunks, pia_chunks, cell_chunks, species) parallel.jl:443-450 for chunks that live in ONE process:
 * the same pack / unpack kernels as mb_exchange_slab, the transport is a device-to-device copy between the chunks' staging buffers. */
int mb_exchange_chunks(int32_t n_chunks, mb_ctx* const* ctxs, const mb_grid1d* slabs, mb_pv* const* pvs, mb_pia* const* pias, int64_t species) {
    MB_ARG(n_chunks >= 1 && ctxs && slabs && pvs && pias, "NULL");
    for (int i = 0; i < n_chunks; i++) MB_ARG(ctxs[i] && pvs[i] && pias[i] && pvs[i]->ctx == ctxs[i], "chunk handles");
    for (int i = 0; i < n_chunks; i++)
        for (int j = 0; j < i; j++) MB_ARG(ctxs[i] != ctxs[j], "every chunk needs its own context (its staging buffers live there)");
    std::vector<XchPlan> P(n_chunks);
    int first_err = MB_OK;
    for (int i = 0; i < n_chunks; i++) {
        ProfScope ps(ctxs[i], PROF_EXCHANGE);
        int r = xch_begin(ctxs[i], slabs + i, pvs[i], pias[i], species, i, n_chunks, false, P[i]);
        if (r) return r;  // nothing has been moved yet
        MB_ARG(P[i].edge == P[0].edge, "the chunks disagree on the exchange mode (same operator sequence and mb_exchange_set_mode on every chunk)");
    }
    std::vector<XchCounts> K(n_chunks);
    if (!P[0].edge) {
        for (int i = 0; i < n_chunks; i++) {
            MB_CUDA(cudaSetDevice(ctxs[i]->device));
            MB_CUDA(cudaMemcpyAsync(ctxs[i]->h_xch_counts, ctxs[i]->d_xch_counts, 4 * 8, cudaMemcpyDeviceToHost, ctxs[i]->stream));
        }
    }
    for (int i = 0; i < n_chunks; i++) {  // every pack has finished before a neighbour reads the staging buffer
        MB_CUDA(cudaSetDevice(ctxs[i]->device));
        MB_CUDA(cudaStreamSynchronize(ctxs[i]->stream));
    }
    if (!P[0].edge) {
        for (int i = 0; i < n_chunks; i++) {  // what NCCL's count exchange delivers
            int64_t* h = ctxs[i]->h_xch_counts;
            h[4] = h[5] = h[6] = h[7] = 0;
            if (i > 0) { h[4] = ctxs[i - 1]->h_xch_counts[2]; h[5] = (int64_t)ctxs[i - 1]->xch_cap; }
            if (i + 1 < n_chunks) { h[6] = ctxs[i + 1]->h_xch_counts[0]; h[7] = (int64_t)ctxs[i + 1]->xch_cap; }
        }
        for (int i = 0; i < n_chunks; i++) K[i] = xch_negotiate(ctxs[i]->h_xch_counts, (int64_t)ctxs[i]->xch_cap, P[i]);
    }
    for (int i = 0; i < n_chunks; i++) {
        MB_CUDA(cudaSetDevice(ctxs[i]->device));
        cudaStream_t st = ctxs[i]->stream;
        if (P[0].edge) {
            if (i > 0) MB_CUDA(cudaMemcpyAsync(ctxs[i]->xch_recv[0], ctxs[i - 1]->xch_send[1], (size_t)XE_MSG * 8, cudaMemcpyDefault, st));
            if (i + 1 < n_chunks) MB_CUDA(cudaMemcpyAsync(ctxs[i]->xch_recv[1], ctxs[i + 1]->xch_send[0], (size_t)XE_MSG * 8, cudaMemcpyDefault, st));
        } else {
            if (K[i].rL > 0) MB_CUDA(cudaMemcpyAsync(ctxs[i]->xch_recv[0], ctxs[i - 1]->xch_send[1], (size_t)K[i].rL * 56, cudaMemcpyDefault, st));
            if (K[i].rR > 0) MB_CUDA(cudaMemcpyAsync(ctxs[i]->xch_recv[1], ctxs[i + 1]->xch_send[0], (size_t)K[i].rR * 56, cudaMemcpyDefault, st));
        }
        ProfScope ps(ctxs[i], PROF_EXCHANGE);
        int r = P[0].edge ? xch_finish_edge(ctxs[i], pvs[i], pias[i], P[i]) : xch_finish_full(ctxs[i], pvs[i], pias[i], P[i], K[i], nullptr, nullptr);
        if (r && !first_err) first_err = r;
    }
    for (int i = 0; i < n_chunks; i++) {  // the staging buffers may be repacked by the next call
        MB_CUDA(cudaSetDevice(ctxs[i]->device));
        MB_CUDA(cudaStreamSynchronize(ctxs[i]->stream));
    }
    return first_err;
}

}  // extern "C"

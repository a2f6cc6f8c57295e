// Generic device-wide exclusive scan (int32 counts -> int64 offsets, n + 1 outputs: out[n] = total).
// Three launches (block sums, scan of block sums, apply); deterministic; used for per-cell offsets in
// squash / NTC split windows / merge compaction.  The sort has its own fused variant in mb_sort.cu.
#pragma once
#include "mb_common.cuh"

namespace mb {

constexpr int GS_BLOCK = 256;
constexpr int GS_ITEMS = 8;
constexpr int GS_TILE = GS_BLOCK * GS_ITEMS;

static __global__ void __launch_bounds__(GS_BLOCK) k_gs_reduce(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ partial) {
    __shared__ int64_t red[GS_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * GS_TILE;
    int64_t s = 0;
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k * GS_BLOCK + threadIdx.x;
        if (i < n) s += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int i = 0; i < GS_BLOCK / 32; i++) t += red[i];
        partial[blockIdx.x] = t;
    }
}
static __global__ void __launch_bounds__(1024) k_gs_partials(int64_t* __restrict__ partial, int64_t nb) {
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
            const int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
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
static __global__ void __launch_bounds__(GS_BLOCK) k_gs_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ partial,
                                                             int64_t* __restrict__ out) {
    __shared__ int64_t wsum[GS_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * GS_TILE + (int64_t)threadIdx.x * GS_ITEMS;
    int h[GS_ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k;
        h[k] = i < n ? in[i] : 0;
        tsum += h[k];
    }
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
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k;
        if (i < n) { out[i] = run; run += h[k]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[gridDim.x];
}

// partial must hold ceil(n / GS_TILE) + 1 int64
static inline size_t gs_partial_count(int64_t n) { return (size_t)((n + GS_TILE - 1) / GS_TILE) + 2; }
static inline int device_exclusive_scan(mb_ctx* ctx, const int32_t* in, int64_t n, int64_t* out, int64_t* partial) {
    const int nb = (int)((n + GS_TILE - 1) / GS_TILE);
    k_gs_reduce<<<nb, GS_BLOCK, 0, ctx->stream>>>(in, n, partial);
    MB_LAUNCH_CHECK(ctx);
    k_gs_partials<<<1, 1024, 0, ctx->stream>>>(partial, nb);
    MB_LAUNCH_CHECK(ctx);
    k_gs_apply<<<nb, GS_BLOCK, 0, ctx->stream>>>(in, n, partial, out);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

}  // namespace mb

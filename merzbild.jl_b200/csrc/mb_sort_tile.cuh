// Pass B of the band sort as a TILE kernel (sort_particles!, grids/grid_sorting.jl:58-113; included by mb_sort.cu).
//
// The warp-per-cell scatter (k_band_scatter*) stores every mover with seven isolated 8-byte writes; that is fine while 95 % of a cell
// stays (dx = 1e-5 m), and halves the bandwidth when sigma_v dt is a few cells (the reference's published grid, BENCHMARKS.md:93-99:
// 85 % of a cell moves, spread over ~17 neighbours).  Here a CTA takes a TILE of consecutive old cells -- one contiguous slice of the
// input of at most NCAP particles -- and uses the fact that the particles a tile sends to one destination cell are CONTIGUOUS in the
// output (consecutive source cells are consecutive in the stable order): the tile's output is a handful of runs, one per destination
// cell, and runs that are neighbours in global memory (complete interior cells) form one SEGMENT.
//
//   producer warp : 1-D TMA loads (cp.async.bulk global -> shared, mbarrier complete_tx; SASS UBLKCP.S.G) of the tile's slice of dr and
//                   of the seven fields, SUB particles per stage of an S-deep ring, running ahead of the consumers across tile boundaries
//   table warp    : for the NEXT tile (double-buffered): the band-matrix rows of the tile and of the 2 w cells before it staged in
//                   shared memory, run sizes R(r) and global starts G(r), the local layout Lp(r) (warp scans; parity padded, below),
//                   the segment list, the local offset LO(c', d) of every (source cell, destination) group, the moments' shifts
//   consumer warps: local output index l = LO(c', d) + rank of every particle (registers), then per field: A[j] -> B[l] (permutation
//                   in shared memory), fence.proxy.async + barrier, write-out a warp per segment: a long segment leaves as ONE bulk
//                   store (cp.async.bulk shared -> global; SASS UBLKCP.G.S) of its 16-byte aligned middle, a short one lane by lane.
//
// Parity padding: a bulk store needs 16-byte aligned addresses on both sides, so every run starts at a local index of the same parity
// as its global start:  Lp(r) = E(r) + pad(r),  E = exclusive prefix of the run sizes,  pad(r) = 2 (parity changes so far) + b(r),
// b(r) = (G(r) - E(r)) & 1 -- monotone, and equal for runs that are neighbours in global memory.
//
// HBM traffic is what the warp-per-cell kernel moves (4 + 56 + 56 B per particle; ncu: 14.6 GB at 1.25e8 particles = 1.006 x the
// algorithmic bytes), but the loads are issued by the copy engine (no registers tied up by bytes in flight) and the stores are full
// sectors.  Measured (profiles/README.md): published grid 3.87 -> 3.08 ms (72 % of the measured HBM peak; 2.61 ms = 85 % without the
// moments); at dx = 1e-5 m the warp-per-cell kernel stays faster (2.62 against 2.77 ms), so the sort picks the tile kernel from w = 4 on.
//
// Moments (compute_props_sorted! for free, every band width): when w and a velocity component of a tile are in shared memory in OUTPUT
// order, every run is a contiguous slice.  A thread sums a strip of <= TL_PIECE consecutive elements of one run (shifted by K(c) =
// velocity of the first particle of old cell c, like the warp-per-cell kernel; no shuffles), after vz a warp per run adds the strips
// (one halving reduction of the seven sums), and the partial sums of (tile kappa, run r) go to Pp[2 w kappa + c + w] -- a collision-free
// slot, because tile kappa starts at cell ca(kappa) >= kappa.  k_tile_combine adds the (1-3) partials of a cell in tile order.
//
// Tiles: old cell c belongs to tile floor(old_start(c) / NI), NI = NCAP - (largest cell), so a tile never exceeds NCAP particles;
// chunk_first[kappa] = first cell of tile kappa.  A tile with more than TL_CMAX cells, or a cell that does not fit, is scattered
// directly (every particle stored on its own, correct for any population); the cached moments are then recomputed by
// k_tile_moments_fallback.  MB_TILE_DEBUG=1 makes CTA 0 print where its consumers and its table warp spent their cycles.
#pragma once
#include "mb_common.cuh"

namespace mb {

constexpr int TL_CMAX = 64;     // cells per tile
constexpr int TL_WMAX = 31;     // widest band
constexpr int TL_NRMAX = TL_CMAX + TL_WMAX - 1;  // runs (destination cells) per tile
constexpr int TL_LOS = 32;      // row stride of LO
constexpr int TL_PIECE = 16;    // elements of a run one thread sums (a "strip": strips never cross runs)
#ifndef MB_TL_SPLIT
#define MB_TL_SPLIT 2
#endif
constexpr int TL_SPLIT = MB_TL_SPLIT;  // threads per strip in the moments pass; measured (published / same-dx pass B): 1: 3.16 / 2.93 ms, 2: 3.13 / 2.79, 4: 3.18 / 2.92
#ifndef MB_TL_BULK_MIN
#define MB_TL_BULK_MIN 64
#endif
constexpr int TL_BULK_MIN = MB_TL_BULK_MIN;  // shortest segment that leaves as a bulk store; measured (published-grid pass B): 16: 3.08 ms, 32 / 64: 3.07, 192: 3.11, 512: 3.31
enum { F_MOM_BAD = 5 };         // ctx->d_flags slot: a tile was scattered directly, its moments are missing

// ---- PTX helpers (sm_100a): mbarrier + 1-D bulk copy
__device__ __forceinline__ uint32_t tl_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tl_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tl_smem(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tl_smem(bar)), "r"(bytes) : "memory");
}
// A wait that cannot hang the GPU: a protocol error traps (the launch fails with an error) instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = tl_smem(bar);
    for (uint32_t spin = 0;; spin++) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tl_smem(dst)), "l"(src), "r"(bytes),
                 "r"(tl_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tl_smem(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NTHREADS>
__device__ __forceinline__ void tl_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// ---- tiles: old_start = exclusive scan of the old cell sizes, the largest cell, NI, NK, chunk_first
// tp[0] = NI, tp[1] = NK (number of tiles), tp[2] = largest old cell, tp[3] = n_old (sum of the old cells)
static __global__ void __launch_bounds__(GS_BLOCK) k_tile_reduce(const int32_t* __restrict__ seg_n, int64_t n, int64_t* __restrict__ partial,
                                                                int64_t* __restrict__ tp, const int* flags) {
    if (flags[2] != 0) return;
    __shared__ int64_t red[GS_BLOCK / 32];
    __shared__ int redm[GS_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * GS_TILE;
    int64_t s = 0;
    int m = 0;
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k * GS_BLOCK + threadIdx.x;
        if (i < n) { const int v = seg_n[i]; s += v; m = v > m ? v : m; }
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); m = max(m, __shfl_down_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; redm[threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        int mm = 0;
        for (int i = 0; i < GS_BLOCK / 32; i++) { t += red[i]; mm = max(mm, redm[i]); }
        partial[blockIdx.x] = t;
        atomicMax((unsigned long long*)(tp + 2), (unsigned long long)mm);
    }
}
static __global__ void __launch_bounds__(1024) k_tile_partials(int64_t* __restrict__ partial, int64_t nb, int64_t* __restrict__ tp, int ncap,
                                                              const int* flags) {
    if (flags[2] != 0) return;
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
    if (threadIdx.x == 0) {
        partial[nb] = carry;
        const int64_t maxn = tp[2];
        int64_t NI = (int64_t)ncap - maxn;
        if (NI < ncap / 4) NI = ncap / 4;  // cells that do not fit are scattered directly
        tp[0] = NI;
        tp[1] = (carry + NI - 1) / NI;
        tp[3] = carry;
    }
}
static __global__ void __launch_bounds__(GS_BLOCK) k_tile_apply(const int32_t* __restrict__ seg_n, int64_t n, const int64_t* __restrict__ partial,
                                                               int64_t* __restrict__ old_start, int32_t* __restrict__ chunk_first,
                                                               const int64_t* __restrict__ tp, const int* flags) {
    if (flags[2] != 0) return;
    __shared__ int64_t wsum[GS_BLOCK / 32];
    const int64_t base = (int64_t)blockIdx.x * GS_TILE + (int64_t)threadIdx.x * GS_ITEMS;
    int h[GS_ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k;
        h[k] = i < n ? seg_n[i] : 0;
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
    const int64_t NI = tp[0], NK = tp[1];
    // tile of the cell before this thread's first one
    int64_t kprev = -1;
    if (base > 0 && base <= n) kprev = (run - seg_n[base - 1]) / NI;
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t i = base + k;
        if (i < n) {
            old_start[i] = run;
            const int64_t kap = run / NI;
            for (int64_t q = kprev + 1; q <= kap && q <= NK; q++) chunk_first[q] = (int32_t)i;
            kprev = kap;
            run += h[k];
            if (i == n - 1) {
                old_start[n] = run;
                for (int64_t q = kap + 1; q <= NK; q++) chunk_first[q] = (int32_t)n;
            }
        }
    }
}

// ---- shared-memory layout
template <int NCAP, int SUB, int S, bool MOM>
struct TileSmem {
    static constexpr int NB = MOM ? 3 : 2;                                    // output-order buffers (MOM: B0 keeps w until vz is through)
    static constexpr int A_BYTES = (SUB + 4) * 8;                             // one stage: SUB particles of a field slice (16-byte granular)
    static constexpr int B_ELEMS = NCAP + 2 * TL_NRMAX + 4;                   // + parity padding (at most two elements per run)
    static constexpr int B_BYTES = (B_ELEMS * 8 + 15) / 16 * 16;
    static constexpr int UMAX = NCAP / TL_PIECE + TL_NRMAX;                   // moment units (run, piece)
    static constexpr int RQ = ((TL_NRMAX + 1) * 4 + 15) / 16 * 16;            // an int32 [NRMAX + 1] array
    // one table buffer
    static constexpr int T_HDR = 0;                                           // ints: mode ncell n nseg NR ca nunits pad | int64 p0 kappa
    static constexpr int T_SEGB = 64;                                         // int32 [CMAX + 1] cell boundaries relative to p0
    static constexpr int T_LP = T_SEGB + ((TL_CMAX + 1) * 4 + 15) / 16 * 16;   // int32 [NRMAX + 1] local start of a run (parity padded)
    static constexpr int T_R = T_LP + RQ;                                     // int32 [NRMAX] run sizes
    static constexpr int T_UST = T_R + RQ;                                    // int32 [NRMAX + 1] first moment unit of a run
    static constexpr int T_SL = T_UST + RQ;                                   // int32 [NRMAX] segment: local start
    static constexpr int T_SN = T_SL + RQ;                                    // int32 [NRMAX] segment: elements
    static constexpr int T_SG = T_SN + RQ;                                    // int64 [NRMAX] segment: global start
    static constexpr int T_K = T_SG + TL_NRMAX * 8;                            // double [NRMAX][3] shifts
    static constexpr int T_LO = T_K + (MOM ? TL_NRMAX * 24 : 0);              // uint16 [CMAX][32]
    static constexpr int T_URUN = T_LO + TL_CMAX * TL_LOS * 2;                // uint8 [UMAX] run of a unit
    static constexpr int T_COF = T_URUN + (MOM ? UMAX : 0);                   // uint8 [NCAP / 32] cell of position 32 i
    static constexpr int T_BYTES = (T_COF + NCAP / 32 + 15) / 16 * 16;
    // whole CTA
    static constexpr int O_A = 0;
    static constexpr int O_B = O_A + S * A_BYTES;
    static constexpr int O_T = O_B + NB * B_BYTES;
    static constexpr int O_PP = O_T + 2 * T_BYTES;                            // double [UMAX][7]
    static constexpr int O_BAR = O_PP + (MOM ? UMAX * 56 : 0);                // full[S] empty[S] tfull[2] tempty[2]
    static constexpr int O_OFF = O_BAR + (2 * S + 4) * 8;                     // int64 [32] direct mode: group offsets
    static constexpr int O_MSM = O_OFF + 32 * 8;                              // int32 [NRMAX][W]: band-matrix rows of the next tile (table warp)
    static constexpr int BYTES = O_MSM + TL_NRMAX * TL_WMAX * 4;
};

struct TileArgs {
    SoA in, out;
    const uint32_t* dr;
    const int32_t* M;
    const int64_t* old_start;
    const int32_t* chunk_first;
    const int64_t* tp;
    const int64_t* start;   // new cell starts (scan)
    const int32_t* cntB;    // extras in front of the band groups
    int64_t n_cells;
    int W;
    double* Pp;             // nullable: partial moments
    int* flags;
    int debug;              // MB_TILE_DEBUG: CTA 0 prints where its consumer thread 0 waited (cycles)
};

// one tile scattered directly (any population): a cell at a time, the whole CTA streams it
template <int NT>
__device__ __forceinline__ void tile_direct(const TileArgs& a, int64_t ca, int64_t cb, int64_t* off_s, int tid) {
    const int W = a.W, w = W / 2;
    for (int64_t c = ca; c < cb; c++) {
        const int64_t lo = a.old_start[c];
        const int n = (int)(a.old_start[c + 1] - lo);
        if (n == 0) continue;  // CTA-uniform
        tl_consumer_sync<NT>();
        if (tid < W) {
            const int64_t cd = c + tid - w;
            int64_t off = 0;
            if (cd >= 0 && cd < a.n_cells) {
                int acc = 0;
                for (int k = 1; k < W; k++) {
                    const int64_t cs = c - k;
                    if (tid + k < W && cs >= 0) acc += a.M[cs * W + tid + k];
                }
                off = a.start[cd] + a.cntB[cd] + acc;
            }
            off_s[tid] = off;
        }
        tl_consumer_sync<NT>();
        for (int j = tid; j < n; j += NT) {
            const uint32_t v = a.dr[lo + j];
            if (v == 0xFFFFFFFFu) continue;
            const int64_t pos = off_s[v >> 24] + (int64_t)(v & 0xFFFFFFu);
#pragma unroll
            for (int f = 0; f < 7; f++) a.out.a[f][pos] = a.in.a[f][lo + j];
        }
    }
}

// NT consumer threads + one table warp + one producer warp
template <int NCAP, int NT, int SUB, int S, bool MOM, int MINB>
__global__ void __launch_bounds__(NT + 64, MINB) k_band_tile(TileArgs a) {
    if (a.flags[2] != 0) return;  // general path takes over
    using L = TileSmem<NCAP, SUB, S, MOM>;
    static_assert(SUB % NT == 0 && NCAP % SUB == 0 && SUB % 4 == 0, "tile shape");
    constexpr int PPT = NCAP / NT;  // particles per consumer thread
    constexpr int KPS = SUB / NT;   // of which per stage
    constexpr int NSUB = NCAP / SUB;
    constexpr int NCW = NT / 32;    // consumer warps
    extern __shared__ __align__(128) unsigned char tl_smem_raw[];
    unsigned char* sm = tl_smem_raw;
    uint64_t* full = (uint64_t*)(sm + L::O_BAR);
    uint64_t* empty = full + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.W, w = W / 2;
    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        for (int b = 0; b < 2; b++) { mbar_init(tfull + b, 1); mbar_init(tempty + b, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t NK = a.tp[1];
    const int64_t n_cells = a.n_cells;

    if (warp == NCW + 1) {
        // ------------------------------------------------------------------ producer: TMA loads, one field per stage
        if (lane != 0) return;
        uint32_t q = 0;  // load counter of this CTA
        for (int64_t kap = blockIdx.x; kap < NK; kap += gridDim.x) {
            const int64_t ca = a.chunk_first[kap], cb = a.chunk_first[kap + 1];
            if (cb <= ca) continue;
            const int64_t p0 = a.old_start[ca];
            const int64_t n = a.old_start[cb] - p0;
            if (n == 0 || n > NCAP || cb - ca > TL_CMAX) continue;  // empty, or scattered directly
            for (int f = -1; f < 7; f++) {
                for (int m = 0; m * SUB < n; m++, q++) {  // stages of SUB particles: many small loads in flight, released one by one
                    const int s = q % S;
                    mbar_wait(empty + s, ((q / S) & 1) ^ 1);
                    const void* src;
                    uint32_t bytes;
                    if (f < 0) {  // dr: 4-byte words
                        const int64_t hi_all = ((p0 & 3) + n + 3) & ~(int64_t)3;
                        const int64_t lo = (int64_t)m * SUB, hi = lo + SUB + 4 < hi_all ? lo + SUB + 4 : hi_all;
                        src = a.dr + (p0 & ~(int64_t)3) + lo;
                        bytes = (uint32_t)(hi - lo) * 4;
                    } else {
                        const int64_t hi_all = ((p0 & 1) + n + 1) & ~(int64_t)1;
                        const int64_t lo = (int64_t)m * SUB, hi = lo + SUB + 2 < hi_all ? lo + SUB + 2 : hi_all;
                        src = a.in.a[f] + (p0 & ~(int64_t)1) + lo;
                        bytes = (uint32_t)(hi - lo) * 8;
                    }
                    mbar_expect_tx(full + s, bytes);
                    tma_load_1d(sm + L::O_A + s * L::A_BYTES, src, bytes, full + s);
                }
            }
        }
        return;
    }

    if (warp == NCW) {
        // ------------------------------------------------------------------ table warp
        uint32_t it = 0;
        const bool tdbg = a.debug && blockIdx.x == 0 && lane == 0;
        long long tw_wait = 0, tw_load = 0, tw_rg = 0, tw_serial = 0, tw_lo = 0, tq = 0;
        for (int64_t kap = blockIdx.x; kap < NK; kap += gridDim.x) {
            // (the consumers learn everything about a tile from its table buffer, also that there is nothing to do: no global loads,
            // no load latency on their side)
            const int64_t ca = a.chunk_first[kap], cb = a.chunk_first[kap + 1];
            const int64_t p0 = cb > ca ? a.old_start[ca] : 0;
            const int64_t n = cb > ca ? a.old_start[cb] - p0 : 0;
            const int tb = it & 1;
            if (tdbg) tq = clock64();
            mbar_wait(tempty + tb, ((it >> 1) & 1) ^ 1);
            if (tdbg) { const long long t1 = clock64(); tw_wait += t1 - tq; tq = t1; }
            unsigned char* T = sm + L::O_T + tb * L::T_BYTES;
            int* hdr = (int*)(T + L::T_HDR);
            const int ncell = (int)(cb - ca);
            const bool direct = n > NCAP || ncell > TL_CMAX;
            if (n == 0) {
                if (lane == 0) hdr[0] = 2;  // empty tile
            } else if (direct) {
                if (lane == 0) { hdr[0] = 1; hdr[1] = ncell; hdr[5] = (int)ca; }
            } else {
                int* segb = (int*)(T + L::T_SEGB);
                int* Lp = (int*)(T + L::T_LP);
                int* Rs = (int*)(T + L::T_R);
                int* ust = (int*)(T + L::T_UST);
                int* seg_l = (int*)(T + L::T_SL);
                int* seg_n = (int*)(T + L::T_SN);
                int64_t* seg_g = (int64_t*)(T + L::T_SG);
                double* Ksh = (double*)(T + L::T_K);
                uint16_t* LO = (uint16_t*)(T + L::T_LO);
                uint8_t* urun = (uint8_t*)(T + L::T_URUN);
                int* Msm = (int*)(sm + L::O_MSM);
                const int NR = ncell + 2 * w;
                // everything the tables need from global memory in two round trips of independent loads: the band-matrix rows of the
                // cells ca - 2w .. cb - 1 (one contiguous block) into shared memory, and per run (destination cell) the scanned start,
                // the extras in front, and the first particle of the old cell (the moments' shift)
                const int64_t c_lo = ca - 2 * w;
                const int mtot = NR * W;
                {
                    const int tneg = c_lo < 0 ? (int)(-c_lo) * W : 0;  // rows of cells < 0 (the first tiles): zeros
                    const int32_t* __restrict__ Mg = a.M + c_lo * W;
                    for (int t0 = lane; t0 < mtot; t0 += 32 * 8) {
                        int mv[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int t = t0 + 32 * u;
                            mv[u] = (t < mtot && t >= tneg) ? Mg[t] : 0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (t0 + 32 * u < mtot) Msm[t0 + 32 * u] = mv[u];
                    }
                }
                for (int i = lane; i <= ncell; i += 32) segb[i] = (int)(a.old_start[ca + i] - p0);
                constexpr int RR = (TL_NRMAX + 31) / 32;  // rounds of 32 runs
                int64_t gst[RR], f0[RR];
                bool rv[RR], hasp[RR];
#pragma unroll
                for (int i = 0; i < RR; i++) {
                    const int r = lane + 32 * i;
                    const int64_t cd = ca - w + r;
                    rv[i] = r < NR && cd >= 0 && cd < n_cells;
                    gst[i] = 0; f0[i] = 0; hasp[i] = false;
                    if (rv[i]) {
                        gst[i] = a.start[cd] + a.cntB[cd];
                        if (MOM) { f0[i] = a.old_start[cd]; hasp[i] = a.old_start[cd + 1] > f0[i]; }
                    }
                }
                if (MOM) {
#pragma unroll
                    for (int i = 0; i < RR; i++) {
                        const int r = lane + 32 * i;
                        double k1 = 0, k2 = 0, k3 = 0;
                        if (hasp[i]) { k1 = a.in.a[1][f0[i]]; k2 = a.in.a[2][f0[i]]; k3 = a.in.a[3][f0[i]]; }
                        if (r < NR) { Ksh[3 * r] = k1; Ksh[3 * r + 1] = k2; Ksh[3 * r + 2] = k3; }
                    }
                }
                __syncwarp();
                if (tdbg) { const long long t1 = clock64(); tw_load += t1 - tq; tq = t1; }
                {   // cell of every 32nd position: the last cell that starts at or before it
                    uint8_t* cof = (uint8_t*)(T + L::T_COF);
                    for (int i = lane; i < (int)((n + 31) >> 5); i += 32) {
                        const int pos = i << 5;
                        int lo_ = 0, hi_ = ncell;
                        while (hi_ - lo_ > 1) {
                            const int mid = (lo_ + hi_) >> 1;
                            if (segb[mid] <= pos) lo_ = mid; else hi_ = mid;
                        }
                        cof[i] = (uint8_t)lo_;
                    }
                }
                // run sizes and global starts; row i of Msm = cell c_lo + i, the tile's cells are rows 2w .. 2w + ncell - 1
                int Rv[RR];
                int64_t Gv[RR];
#pragma unroll
                for (int i = 0; i < RR; i++) {
                    const int r = lane + 32 * i;
                    int R = 0;
                    int64_t G = gst[i];
                    if (r < NR && rv[i]) {
                        const int row_cd = r + w;  // row of the destination cell itself
                        const int t0 = row_cd - w > 2 * w ? row_cd - w : 2 * w, t1 = row_cd + w < 2 * w + ncell - 1 ? row_cd + w : 2 * w + ncell - 1;
                        for (int row = t0; row <= t1; row++) R += Msm[row * W + (row_cd - row + w)];
                        const int e0 = row_cd - w > 0 ? row_cd - w : 0;  // sources in earlier tiles: rows e0 .. 2w - 1
                        for (int row = e0; row < 2 * w; row++) G += Msm[row * W + (row_cd - row + w)];
                    }
                    Rv[i] = R;
                    Gv[i] = G;
                    if (r < NR) Rs[r] = R;
                }
                if (tdbg) { const long long t1 = clock64(); tw_rg += t1 - tq; tq = t1; }
                // The local layout (output order).  Runs that are neighbours in global memory stay neighbours (one SEGMENT = one bulk
                // store per field); a run starts at a local index of the same parity as its global start, so that the 16-byte aligned
                // middle of a segment is 16-byte aligned in shared memory too:  Lp(r) = E(r) + pad(r),  E = exclusive prefix of the run
                // sizes,  pad(r) = 2 (number of parity changes among the non-empty runs up to r) + b(r),  b(r) = (G(r) - E(r)) & 1.
                // pad never decreases, and neighbours in global memory have the same b, hence the same pad.  All with warp scans.
                int Lv[RR];
                {
                    const unsigned le = 0xffffffffu >> (31 - lane);  // lanes <= this one
                    int cE = 0, cK = 0, cS = 0, cU = 0, cB = -1;     // carries: prefix of R, parity changes, segments, units, last parity
                    int64_t cGend = -1;
#pragma unroll
                    for (int i = 0; i < RR; i++) {
                        if (32 * i >= NR) break;  // warp-uniform
                        const int r = lane + 32 * i;
                        const int R = Rv[i];
                        int incl = R;
                        const int U = MOM ? (R + TL_PIECE - 1) / TL_PIECE : 0;
                        int uincl = U;
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(0xffffffffu, incl, o), tu = __shfl_up_sync(0xffffffffu, uincl, o);
                            if (lane >= o) { incl += t; uincl += tu; }
                        }
                        const int E = cE + incl - R;
                        const int b = (int)((Gv[i] - E) & 1);
                        const int64_t gend = Gv[i] + R;
                        const unsigned ne = __ballot_sync(0xffffffffu, R > 0);
                        const unsigned lower = ne & (le >> 1);           // non-empty runs in lower lanes
                        const int src = lower ? 31 - __clz(lower) : 0;
                        const int64_t gend_s = __shfl_sync(0xffffffffu, gend, src);
                        const int b_s = __shfl_sync(0xffffffffu, b, src);
                        const int64_t gprev = lower ? gend_s : cGend;
                        const int bprev = lower ? b_s : cB;
                        const bool segstart = R > 0 && Gv[i] != gprev;
                        const bool change = R > 0 && bprev >= 0 && b != bprev;
                        const unsigned mch = __ballot_sync(0xffffffffu, change), mss = __ballot_sync(0xffffffffu, segstart);
                        const int kc = cK + __popc(mch & le);
                        // an empty run takes the parity of the last non-empty run before it (it only needs a valid slot)
                        const int lp = E + 2 * kc + (R > 0 ? b : (bprev >= 0 ? bprev : 0));
                        Lv[i] = lp;
                        if (r < NR) Lp[r] = lp;
                        if (segstart) {
                            const int si = cS + __popc(mss & le) - 1;
                            seg_l[si] = lp; seg_g[si] = Gv[i]; seg_n[si] = E;  // (E for now: the sizes follow from the next segment's E)
                        }
                        if (MOM && r < NR) {
                            const int u0 = cU + uincl - U;
                            ust[r] = u0;
                            for (int k = 0; k < U; k++) urun[u0 + k] = (uint8_t)r;
                        }
                        cE += __shfl_sync(0xffffffffu, incl, 31);
                        cU += __shfl_sync(0xffffffffu, uincl, 31);
                        cK += __popc(mch);
                        cS += __popc(mss);
                        if (ne) {
                            const int last = 31 - __clz(ne);
                            cGend = __shfl_sync(0xffffffffu, gend, last);
                            cB = __shfl_sync(0xffffffffu, b, last);
                        }
                    }
                    __syncwarp();
                    int e0v[RR], e1v[RR];
#pragma unroll
                    for (int i = 0; i < RR; i++) {
                        const int si = lane + 32 * i;
                        e0v[i] = si < cS ? seg_n[si] : 0;
                        e1v[i] = si + 1 < cS ? seg_n[si + 1] : cE;
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < RR; i++) {
                        const int si = lane + 32 * i;
                        if (si < cS) seg_n[si] = e1v[i] - e0v[i];
                    }
                    if (lane == 0) {
                        if (MOM) ust[NR] = cU;
                        hdr[0] = 0; hdr[1] = ncell; hdr[2] = (int)n; hdr[3] = cS; hdr[4] = NR; hdr[5] = (int)ca; hdr[6] = cU;
                        *(int64_t*)(hdr + 8) = p0;
                        *(int64_t*)(hdr + 10) = kap;
                    }
                }
                __syncwarp();
                if (tdbg) { const long long t1 = clock64(); tw_serial += t1 - tq; tq = t1; }
                // local offset of every (source cell, destination) group
#pragma unroll
                for (int i = 0; i < RR; i++) {
                    const int r = lane + 32 * i;
                    if (r < NR && rv[i]) {
                        const int row_cd = r + w;
                        const int t0 = row_cd - w > 2 * w ? row_cd - w : 2 * w, t1 = row_cd + w < 2 * w + ncell - 1 ? row_cd + w : 2 * w + ncell - 1;
                        int acc = Lv[i];
                        for (int row = t0; row <= t1; row++) {
                            LO[(row - 2 * w) * TL_LOS + (row_cd - row + w)] = (uint16_t)acc;
                            acc += Msm[row * W + (row_cd - row + w)];
                        }
                    }
                }
            }
            __syncwarp();
            if (tdbg) { const long long t1 = clock64(); tw_lo += t1 - tq; tq = t1; }
            if (lane == 0) mbar_arrive(tfull + tb);
            it++;
        }
        if (tdbg)
            printf("k_band_tile CTA 0 table warp: %u tiles; waiting for a free buffer %lld, loads %lld, cof + run sizes %lld, layout (lane 0) %lld, LO %lld cycles\n", it,
                   tw_wait, tw_load, tw_rg, tw_serial, tw_lo);
        return;
    }

    // ---------------------------------------------------------------------- consumers
    uint32_t q = 0, it = 0, bsel = 0;
    long long t_tab = 0, t_full = 0, t_bar = 0, t_all = clock64(), tt = 0;
    const bool dbg = a.debug && blockIdx.x == 0 && tid == 0;
    const double* Bbase = (const double*)(sm + L::O_B);
    double* pp = (double*)(sm + L::O_PP);
    for (int64_t kap = blockIdx.x; kap < NK; kap += gridDim.x) {
        const int tb = it & 1;
        if (dbg) tt = clock64();
        mbar_wait(tfull + tb, (it >> 1) & 1);
        if (dbg) t_tab += clock64() - tt;
        unsigned char* T = sm + L::O_T + tb * L::T_BYTES;
        const int* hdr = (const int*)(T + L::T_HDR);
        if (hdr[0] != 0) {  // empty (2) or direct (1)
            if (hdr[0] == 1) {
                tile_direct<NT>(a, hdr[5], (int64_t)hdr[5] + hdr[1], (int64_t*)(sm + L::O_OFF), tid);
                if (MOM && tid == 0) a.flags[F_MOM_BAD] = 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + tb);
            it++;
            continue;
        }
        const int n = hdr[2], nseg = hdr[3], NR = hdr[4], nunits = hdr[6];
        const int64_t p0 = *(const int64_t*)(hdr + 8);
        const int* segb = (const int*)(T + L::T_SEGB);
        const int* Lp = (const int*)(T + L::T_LP);
        const int* Rs = (const int*)(T + L::T_R);
        const int* ust = (const int*)(T + L::T_UST);
        const double* Ksh = (const double*)(T + L::T_K);
        const uint16_t* LO = (const uint16_t*)(T + L::T_LO);
        const uint8_t* urun = (const uint8_t*)(T + L::T_URUN);
        const int64_t* seg_g = (const int64_t*)(T + L::T_SG);
        const int* seg_l = (const int*)(T + L::T_SL);
        const int* seg_n = (const int*)(T + L::T_SN);
        const uint8_t* cof = (const uint8_t*)(T + L::T_COF);
        // ---- dr: local output index of every particle
        int li[PPT];
#pragma unroll
        for (int m = 0; m < NSUB; m++) {
            if (m * SUB < n) {  // CTA-uniform
                const int s = q % S;
                if (dbg) tt = clock64();
                mbar_wait(full + s, (q / S) & 1);
                if (dbg) t_full += clock64() - tt;
                const uint32_t* A32 = (const uint32_t*)(sm + L::O_A + s * L::A_BYTES) + (int)(p0 & 3) - m * SUB;
                // all loads of a step are issued before the first use (the compiler cannot reorder them: everything is shared memory)
                uint32_t v[KPS];
                int cc[KPS];
#pragma unroll
                for (int kk = 0; kk < KPS; kk++) {
                    const int j = tid + (m * KPS + kk) * NT;
                    v[kk] = A32[j];                                       // (beyond n: stale words of the stage, ignored below)
                    cc[kk] = cof[(j < n ? j : n - 1) >> 5];               // the cell of position j & ~31 ...
                }
#pragma unroll
                for (int kk = 0; kk < KPS; kk++) {
                    const int j = tid + (m * KPS + kk) * NT;
                    if (j >= n) v[kk] = 0xFFFFFFFFu;
                    else
                        while (segb[cc[kk] + 1] <= j) cc[kk]++;           // ... and a short walk to the cell of j (empty cells are skipped)
                }
                int lo16[KPS];
#pragma unroll
                for (int kk = 0; kk < KPS; kk++) lo16[kk] = v[kk] != 0xFFFFFFFFu ? (int)LO[cc[kk] * TL_LOS + (int)(v[kk] >> 24)] : 0;
#pragma unroll
                for (int kk = 0; kk < KPS; kk++) {
                    const int k = m * KPS + kk;
                    li[k] = -1;
                    if (v[kk] != 0xFFFFFFFFu) li[k] = lo16[kk] + (int)(v[kk] & 0xFFFFFFu);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
                q++;
            } else {
#pragma unroll
                for (int kk = 0; kk < KPS; kk++) li[m * KPS + kk] = -1;
            }
        }
        const int sh = (int)(p0 & 1);
#pragma unroll 1
        for (int f = 0; f < 7; f++) {
            // MOM: w stays in B0 while vx, vy, vz alternate between B1 and B2; x, y, z then rotate over all three
            // (a fourth buffer -- all of w, v resident, one moments pass instead of three -- was measured: the shared memory it takes from
            // the load stages costs more than the two passes it saves)
            int b;
            if (MOM) b = f == 0 ? 0 : (f <= 3 ? 1 + ((f - 1) & 1) : (f - 4) % 3);
            else { b = bsel & 1; bsel++; }
            double* B = (double*)(sm + L::O_B + b * L::B_BYTES);
#pragma unroll
            for (int m = 0; m < NSUB; m++) {
                if (m * SUB < n) {
                    const int s = q % S;
                    if (dbg) tt = clock64();
                    mbar_wait(full + s, (q / S) & 1);
                    if (dbg) t_full += clock64() - tt;
                    const double* A = (const double*)(sm + L::O_A + s * L::A_BYTES) + sh - m * SUB;
                    double v[KPS];
#pragma unroll
                    for (int kk = 0; kk < KPS; kk++) v[kk] = A[tid + (m * KPS + kk) * NT];
#pragma unroll
                    for (int kk = 0; kk < KPS; kk++)
                        if (li[m * KPS + kk] >= 0) B[li[m * KPS + kk]] = v[kk];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + s);
                    q++;
                }
            }
            // B is complete after the barrier; the copy engine reads it through the async proxy, and the buffer the NEXT field is
            // permuted into must have been read by its previous bulk stores (issued a field or two ago)
            fence_async_smem();
            tma_store_wait_read();  // (every thread: the issuers of the previous tile may not be issuers of this one)
            if (dbg) tt = clock64();
            tl_consumer_sync<NT>();
            if (dbg) t_bar += clock64() - tt;
            if (MOM && f >= 1 && f <= 3) {
                // partial moments of component f, a thread per strip of <= TL_PIECE consecutive elements of one run: sum w, sum w c,
                // sum w c^2 (c = v - K(run)); no shuffles, every thread busy.  The lanes walk their strips rotated against each other
                // (strips start TL_PIECE doubles apart: without the rotation all lanes of a wavefront would hit one bank).
                const double* Bw = Bbase;
                // TL_SPLIT threads per strip (a part of the strip each, combined with shuffles): more threads busy, shorter dependent chains
                constexpr int SP = TL_SPLIT, PART = TL_PIECE / SP;
                for (int uu0 = 0; uu0 < SP * nunits; uu0 += NT) {  // CTA-uniform trip count: every lane takes part in the shuffles
                    const int uu = uu0 + tid;
                    const bool valid = uu < SP * nunits;
                    const int u = valid ? uu / SP : 0, part = uu % SP;
                    const int r = urun[u];
                    const int lb = Lp[r] + (u - ust[r]) * TL_PIECE;
                    const int l0 = lb + part * PART;
                    const int l1 = valid ? min(lb + TL_PIECE, Lp[r] + Rs[r]) : 0;
                    const double K = Ksh[3 * r + (f - 1)];
                    double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                    for (int k = 0; k < PART; k++) {
                        const int l = l0 + ((k + lane / SP) & (PART - 1));
                        if (l < l1) {
                            const double pw = Bw[l], c = B[l] - K;
                            s0 += pw; s1 += pw * c; s2 += pw * (c * c);
                        }
                    }
#pragma unroll
                    for (int o = 1; o < SP; o <<= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    }
                    if (valid && part == 0) {
                        double* o_ = pp + u * 7;
                        if (f == 1) o_[0] = s0;
                        o_[f] = s1;
                        o_[3 + f] = s2;
                    }
                }
            }
            {
                // write-out, a warp per segment: a long segment leaves as ONE bulk store (its 16-byte aligned middle; at most one element
                // at either end is stored by hand), a short one lane by lane (consecutive lanes, consecutive addresses)
                double* __restrict__ of = a.out.a[f];
                for (int sgi = warp; sgi < nseg; sgi += NCW) {
                    const int64_t g = seg_g[sgi];
                    const int l = seg_l[sgi], ns = seg_n[sgi];
                    if (ns >= TL_BULK_MIN) {
                        if (lane == 0) {
                            const int head = (int)(g & 1);
                            const int nb = (ns - head) & ~1;
                            if (head) of[g] = B[l];
                            tma_store_1d(of + g + head, B + l + head, (uint32_t)nb * 8);
                            if (ns - head - nb) of[g + ns - 1] = B[l + ns - 1];
                            tma_store_commit();
                        }
                    } else {
                        for (int e = lane; e < ns; e += 32) of[g + e] = B[l + e];
                    }
                }
            }
            if (MOM && f == 3) {
                tl_consumer_sync<NT>();  // pp complete; B0 (w) may be overwritten by x afterwards
                // the strips of a run -> the run's sums, a warp per run: lanes stride over the strips, then ONE reduction of the seven values
                // together (the lanes halve the set of values they carry in each of the first three steps: 9 shuffles instead of 35)
                const int ca = hdr[5];
                for (int r = warp; r < NR; r += NCW) {
                    const int u0 = ust[r], u1 = ust[r + 1];
                    if (u1 <= u0) continue;  // warp-uniform
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = 0.0;
                    for (int u = u0 + lane; u < u1; u += 32) {
                        const double* o_ = pp + u * 7;
#pragma unroll
                        for (int i = 0; i < 7; i++) v[i] += o_[i];
                    }
                    double a4[4], a2[2];
                    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const double recv = __shfl_xor_sync(0xffffffffu, h16 ? v[i] : v[i + 4], 16);
                        a4[i] = (h16 ? v[i + 4] : v[i]) + recv;
                    }
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const double recv = __shfl_xor_sync(0xffffffffu, h8 ? a4[i] : a4[i + 2], 8);
                        a2[i] = (h8 ? a4[i + 2] : a4[i]) + recv;
                    }
                    double c1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
                    c1 += __shfl_xor_sync(0xffffffffu, c1, 2);
                    c1 += __shfl_xor_sync(0xffffffffu, c1, 1);
                    // value i = 4 b16 + 2 b8 + b4 sits in the lanes with those bits
                    double t[7];
#pragma unroll
                    for (int i = 0; i < 7; i++) t[i] = __shfl_sync(0xffffffffu, c1, ((i >> 2) & 1) * 16 + ((i >> 1) & 1) * 8 + (i & 1) * 4);
                    if (lane == 0) {
                        double* P = a.Pp + ((int64_t)2 * w * kap + ca + r) * 5;
                        P[0] = t[0]; P[1] = t[1]; P[2] = t[2]; P[3] = t[3]; P[4] = t[4] + t[5] + t[6];
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty + tb);
        it++;
    }
    tma_store_wait_read();  // the bulk stores read shared memory: it must outlive them
    if (dbg)
        printf("k_band_tile CTA 0: %u tiles, %lld cycles; waiting for tables %lld, for TMA stages %lld, at the field barrier %lld\n", it,
               clock64() - t_all, t_tab, t_full, t_bar);
}

// Cached moments of the freshly sorted cells from the tiles' partial sums (in tile order: deterministic) plus the extras in front of /
// behind the band groups, read back from the output.  Same shift K(c) as the tiles used.
static __global__ void __launch_bounds__(128) k_tile_combine(const double* __restrict__ Pp, const int32_t* __restrict__ M, const int32_t* __restrict__ cntB,
                                                            const int32_t* __restrict__ cntA, const int64_t* __restrict__ old_start,
                                                            const int64_t* __restrict__ tp, const int64_t* __restrict__ start, SoA in_, SoA out_,
                                                            int64_t n_cells, int W, double* __restrict__ pcache, const int* flags) {
    if (flags[2] != 0 || flags[F_MOM_BAD] != 0) return;
    const int w = W / 2;
    const int64_t NI = tp[0];
    const double* __restrict__ o0 = out_.a[0]; const double* __restrict__ o1 = out_.a[1]; const double* __restrict__ o2 = out_.a[2];
    const double* __restrict__ o3 = out_.a[3];
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        double K1 = 0, K2 = 0, K3 = 0;
        {
            const int64_t f0 = old_start[c];
            if (old_start[c + 1] > f0) { K1 = in_.a[1][f0]; K2 = in_.a[2][f0]; K3 = in_.a[3][f0]; }
        }
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;
        int64_t count = 0, kprev = -1;
        const int64_t s0 = c - w > 0 ? c - w : 0, s1 = c + w < n_cells - 1 ? c + w : n_cells - 1;
        for (int64_t cs = s0; cs <= s1; cs++) {
            const int m = M[cs * W + (int)(c - cs + w)];
            if (m == 0) continue;
            count += m;
            const int64_t kap = old_start[cs] / NI;
            if (kap != kprev) {
                const double* P = Pp + ((int64_t)2 * w * kap + c + w) * 5;
                an += P[0]; ax += P[1]; ay += P[2]; az += P[3]; aq += P[4];
                kprev = kap;
            }
        }
        const int nb = cntB[c], na = cntA[c];
        const int64_t pb = start[c], pa = start[c] + nb + count;
        for (int t = 0; t < nb + na; t++) {
            const int64_t p = t < nb ? pb + t : pa + (t - nb);
            const double pw = o0[p], cx = o1[p] - K1, cy = o2[p] - K2, cz = o3[p] - K3;
            an += pw; ax += pw * cx; ay += pw * cy; az += pw * cz;
            aq += pw * (cx * cx + cy * cy + cz * cz);
        }
        double* pc = pcache + 6 * c;
        pc[0] = (double)(count + nb + na);
        if (an > 0.0) {
            const double mx = ax / an, my = ay / an, mz = az / an;  // mean of (v - K)
            pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
            pc[5] = aq - an * (mx * mx + my * my + mz * mz);        // sum w |v - vbar|^2
        } else {
            pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
        }
    }
}

// A tile went the direct way (no partial sums): the cell moments from the sorted output, a warp per cell.
static __global__ void __launch_bounds__(256) k_tile_moments_fallback(SoA out_, const int64_t* __restrict__ start, int64_t n_cells,
                                                                     double* __restrict__ pcache, const int* flags) {
    if (flags[2] != 0 || flags[F_MOM_BAD] == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < n_cells; c += nwarps) {
        const int64_t lo = start[c];
        const int64_t n = start[c + 1] - lo;
        double K1 = 0, K2 = 0, K3 = 0;
        if (n > 0) { K1 = out_.a[1][lo]; K2 = out_.a[2][lo]; K3 = out_.a[3][lo]; }
        double an = 0, ax = 0, ay = 0, az = 0, aq = 0;
        for (int64_t j = lane; j < n; j += 32) {
            const double pw = out_.a[0][lo + j], cx = out_.a[1][lo + j] - K1, cy = out_.a[2][lo + j] - K2, cz = out_.a[3][lo + j] - K3;
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
            double* pc = pcache + 6 * c;
            pc[0] = (double)n;
            if (an > 0.0) {
                const double mx = ax / an, my = ay / an, mz = az / an;
                pc[1] = an; pc[2] = K1 + mx; pc[3] = K2 + my; pc[4] = K3 + mz;
                pc[5] = aq - an * (mx * mx + my * my + mz * mz);
            } else {
                pc[1] = 0; pc[2] = 0; pc[3] = 0; pc[4] = 0; pc[5] = 0;
            }
        }
    }
}

}  // namespace mb

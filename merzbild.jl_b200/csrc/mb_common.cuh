// Internal definitions shared by all kernels of libmerzbild_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/merzbild_b200.h"

#define MB_HD __host__ __device__ __forceinline__

namespace mb {

constexpr double k_B = 1.380649e-23;              // constants.jl:4
constexpr double twopi = 6.283185307179586;       // constants.jl:29 (2 * pi in fp64)
constexpr double c_light = 299792458.0;           // constants.jl:9
constexpr double EPS = 2.220446049250313e-16;     // Julia eps()
constexpr int N_SM = 148;                         // B200

// device-side error flag bits (ctx->d_flags[0])
// other slots of ctx->d_flags: [2] the sort's general path must run, [3] extras of the last band sort, [4..11] unused,
// [12..15] written by the fused convect + classify kernel (mb_sort.cu)
enum { F_OUTSIDE = 12, F_CLS_BAD = 13, F_FAR = 14, F_CLS_REDO = 15 };
enum : int { DEVERR_CAPACITY = 1, DEVERR_PRECONDITION = 2, DEVERR_BAND_OVERFLOW = 4, DEVERR_BAD_CELL = 8, DEVERR_OCTREE = 16 };

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 streams; the convention (key = seed, counter = (block, entity, timestep, op | substream << 8),
// draw d = the (d & 1)-th double of block d >> 1) is shared with the CPU oracle so that the sequential-per-entity
// kernels replay the oracle draw for draw.
// ---------------------------------------------------------------------------------------------------------------
enum : uint32_t { OP_NTC = 1, OP_CONVECT = 2, OP_MERGE = 3, OP_SWPM = 4, OP_FP = 5, OP_SAMPLE = 6, OP_USER = 7, OP_MERGE_GRID = 8 };

MB_HD void philox_round(uint32_t c[4], const uint32_t k0, const uint32_t k1) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c[1] ^ k0;
    const uint32_t n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
MB_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
MB_HD double u64_to_unit_double(uint32_t lo, uint32_t hi) {
    const uint64_t u = ((uint64_t)hi << 32) | lo;
    return (double)(u >> 11) * (1.0 / 9007199254740992.0);
}

struct PhiloxStream {
    uint32_t k0, k1, c0, c1, c2, c3;
    uint32_t b2, b3;
    int have;
    MB_HD PhiloxStream(uint64_t seed, uint32_t op, uint32_t substream, uint32_t timestep, uint32_t entity)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(0), c1(entity), c2(timestep), c3((op & 0xFFu) | (substream << 8)), b2(0), b3(0),
          have(0) {}
    MB_HD double rand() {
        if (have == 0) {
            uint32_t o[4];
            philox4x32_10(c0, c1, c2, c3, k0, k1, o);
            c0++;
            b2 = o[2]; b3 = o[3];
            have = 1;
            return u64_to_unit_double(o[0], o[1]);
        }
        have = 0;
        return u64_to_unit_double(b2, b3);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// containers
// ---------------------------------------------------------------------------------------------------------------
struct SoA {  // seven fp64 arrays: w, vx, vy, vz, x, y, z
    double* a[7];
};
enum { F_W = 0, F_VX = 1, F_VY = 2, F_VZ = 3, F_X = 4, F_Y = 5, F_Z = 6 };

struct Indexer {  // particles.jl:56-66
    int64_t n_local, start1, end1, n_group1, start2, end2, n_group2;
};

}  // namespace mb

struct mb_ctx {
    int device;
    cudaStream_t stream;
    uint64_t seed;
    int* d_flags;           // [0] error bits, [1] aux (required capacity lo), ...
    int* h_flags;           // pinned mirror
    int64_t n_launch;
    cudaEvent_t ev0, ev1;
    void* l2_scratch;
    size_t l2_scratch_bytes;
    // generic scratch arena (grown on demand, stream-ordered reuse)
    void* scratch[24];
    size_t scratch_bytes[24];
    int sort_last_path;
    int sort_last_tile;     // pass B of the last band sort: 1 = tile kernel
    int band_w;
    // moments cached by the band sort's gather pass (valid while state_gen == pc_gen)
    uint64_t state_gen, pc_gen;
    void* pc_pv;
    void* pc_pia;
    int pc_species;
    int pc_general;         // the general sort path fills the moment cache (k_gen_gather_cells)
    int pc_band;            // the band path fills it (k_band_scatter + k_band_combine; narrow bands only)
    // band classification cached by the fused convect kernel (valid while state_gen == cls_gen)
    uint64_t cls_gen;
    void* cls_pv;
    void* cls_pia;
    int cls_species, cls_w;
    double cls_inv_dx;
    int64_t cls_cell_offset, cls_cap;
    // per-section event profiling
    int prof_on;
    std::vector<cudaEvent_t>* prof_ev;   // pairs (begin, end)
    std::vector<int>* prof_sec;          // section of pair i
    size_t prof_used;                    // pairs in use
    std::vector<size_t>* prof_stack;     // open sections (they may nest)
    // NCCL (dlopen'ed lazily)
    void* nccl_comm;
    int rank, nranks;
    // exchange staging
    void* xch_send[2];
    void* xch_recv[2];
    size_t xch_cap;  // particles per direction
    int64_t* d_xch_counts;   // [0..1] send counts, [2..3] recv counts
    int64_t* h_xch_counts;   // pinned
    int xch_mode;            // 0: edge exchange when the layout allows it, 1: always the full exchange
};

struct mb_pv {
    mb_ctx* ctx;
    int64_t cap;
    mb::SoA cur, alt;   // alt allocated lazily (sort ping-pong)
    bool has_alt;
    int32_t* cell;      // 1-based cell id per logical position (pv.cell); int32 on device
    int drop_oob;       // set by the slab exchange: the next sort drops particles whose cell is outside the slab
                        // (1: every leaver was sent; 2: edge exchange -- only leavers from the w cells next to a slab face were sent)
    int64_t n_arrivals; // host upper bound on the slab-exchange arrivals appended after the sorted layout (merged by the next sort)
    int64_t* d_n_arr;   // device: exact number of those arrivals (the edge exchange never tells the host)
    int arrivals_at_end; // the species was not contiguous at the exchange: the arrivals are parked at [cap - n_arr, cap), not behind n_total
};

struct mb_pia {
    mb_ctx* ctx;
    int64_t n_cells, n_species;
    mb::Indexer* d_indexer;  // [species][cell]
    int64_t* d_n_total;      // [species]
    int64_t* h_n_total;      // pinned host mirror
    bool h_valid;
    std::vector<uint8_t> contiguous;   // host-side flag per species (changes are statically known per operator)
    std::vector<uint8_t> sorted_layout; // host-side: group1 ranges tile 1..n_total in cell order and group2 is empty everywhere
    std::vector<int64_t> n_bound;      // host upper bound on n_total (for launch sizing only)
    std::vector<uint8_t> contig_pending; // a merge ran: the exact contiguous flag is !d_holes[species] (resolved at download)
    int* d_holes;            // [n_species] device flags "a merge left holes", owned by this pia (one per species)
};

struct mb_cf {
    mb_ctx* ctx;
    int64_t n_cells;
    double* sigma_g_w_max;
    int64_t* n_coll;
    int64_t* n_coll_performed;
    int64_t* n_eq_w;
};

struct mb_props {
    mb_ctx* ctx;
    int64_t n_cells, n_species, n_moments;
    int32_t ndens_not_Np;
    double Tref;
    std::vector<int32_t> powers;
    int32_t* d_powers;
    double *lpa, *np, *n, *v, *T, *moments;
};

namespace mb {

void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void* ctx_scratch(mb_ctx* ctx, int slot, size_t bytes);  // returns nullptr on failure (error set)
int pv_ensure_alt(mb_pv* pv);

#define MB_CUDA(x)                                                        \
    do {                                                                  \
        cudaError_t e__ = (x);                                            \
        if (e__ != cudaSuccess) return mb::cuda_fail(e__, #x, __FILE__, __LINE__); \
    } while (0)
#define MB_ARG(cond, msg)                         \
    do {                                          \
        if (!(cond)) {                            \
            mb::set_error(std::string("invalid argument: ") + (msg)); \
            return MB_ERR_ARG;                    \
        }                                         \
    } while (0)
#define MB_LAUNCH_CHECK(ctx)                      \
    do {                                          \
        (ctx)->n_launch++;                        \
        MB_CUDA(cudaPeekAtLastError());           \
    } while (0)

enum { PROF_SORT_CLASSIFY = 0, PROF_SORT_SCAN = 1, PROF_SORT_SCATTER = 2, PROF_SORT_GENERAL = 3, PROF_NTC = 4, PROF_CONVECT = 5, PROF_PROPS = 6,
       PROF_MERGE = 7, PROF_FP = 8, PROF_EXCHANGE = 9, PROF_SQUASH = 10, PROF_SORT_EXTRAS = 11, PROF_NSEC = 12 };
void prof_begin(mb_ctx* ctx, int section);
void prof_end(mb_ctx* ctx);
struct ProfScope {
    mb_ctx* c;
    ProfScope(mb_ctx* ctx, int section) : c(ctx) { if (c->prof_on) prof_begin(c, section); }
    ~ProfScope() { if (c->prof_on) prof_end(c); }
};

// Stream keys.  The rank of a slab-partitioned run and the species (pair) an operator works on are part of every Philox key, so a
// caller who ports the reference's pattern -- one rng handed to ntc!(s1), ntc!(s2), ntc!(s1, s2) and to the convection of every
// species in the same step, the same seed on every rank -- gets independent streams without managing substreams by hand.
// Rank 0 / species 1 leave seed and substream unchanged.  The caller's substream has 12 bits.
inline uint64_t stream_seed(const mb_ctx* c) { return c->seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)c->rank); }
inline uint32_t stream_substream(uint32_t substream, int64_t s1, int64_t s2) {
    return (substream & 0xFFFu) | ((uint32_t)((s1 - 1) & 0x3F) << 12) | ((uint32_t)((s2 - 1) & 0x3F) << 18);
}

inline int grid_for(int64_t n, int block, int per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)N_SM * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace mb

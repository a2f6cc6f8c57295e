// fp_linear! -- linear Fokker-Planck collision operator (collisions/collision_fp.jl:24-125), compute_relaxation_time
// (:143-151), sample_normal_rands! (:164-170), scale_norm_rands! (:182-211).  Defined for sorted cells (n_group2 == 0), as
// in the reference (its group-2 branch uses an undefined variable, collision_fp.jl:57); cells with n_local < 7 are skipped.
//
// One warp per cell; the cell is streamed from HBM once (w, v: 32 B/particle) and written once (v: 24 B); the intermediate
// passes hit L1/L2.  The standard normals are counter-based: particle j of the cell uses Philox block j of the
// (OP_FP, substream, timestep, cell) stream through two fp32 Box-Muller transforms (mb_normals.h: explicitly rounded fp32
// operations, bit-identical in the CPU oracle) -- regenerated in each pass that needs them instead of being stored in the streaming
// kernel, generated once in the register kernels -- and are standardised exactly (mean 0, variance 1 over the cell) like
// scale_norm_rands!.
#include "mb_common.cuh"
#include "mb_normals.h"

namespace mb {

struct FpArgs {
    SoA pv;
    const Indexer* ix;
    int64_t cell_lo, cell_hi;
    mb_interaction it;
    double mass, dt, V;
    uint64_t seed;
    uint32_t timestep, substream;
    int* flags;
};

__device__ __forceinline__ void fp_normals(const FpArgs& a, uint32_t cell, int64_t j, double o[3]) {
    const uint32_t c3 = (OP_FP & 0xFFu) | (a.substream << 8);
    uint32_t r[4];
    philox4x32_10((uint32_t)j, cell, a.timestep, c3, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), r);
    float n0, n1, n2, n3;
    mbn_box_muller(r[0], r[1], &n0, &n1);  // fp32, explicitly rounded operations: bit-identical in the CPU oracle (mb_normals.h)
    mbn_box_muller(r[2], r[3], &n2, &n3);
    o[0] = (double)n0;
    o[1] = (double)n1;
    o[2] = (double)n2;
}
__device__ __forceinline__ double wsum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// All-reduce of 8 (4) values over the warp in ONE pass: in each of the first steps a lane hands half of the values it carries to its
// partner and keeps the other half, so the set halves while the sums grow; after the last step every value sits (complete) in a
// group of lanes and is broadcast from there.  16 (10) 64-bit shuffles instead of the 40 (20) of eight (four) separate butterflies.
__device__ __forceinline__ void wallreduce8(double v[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    double a4[4], a2[2];
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
    double c = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __shfl_sync(0xffffffffu, c, ((i >> 2) & 1) * 16 + ((i >> 1) & 1) * 8 + (i & 1) * 4);
}
__device__ __forceinline__ void wallreduce4(double v[4], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8;
    double a2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double recv = __shfl_xor_sync(0xffffffffu, h16 ? v[i] : v[i + 2], 16);
        a2[i] = (h16 ? v[i + 2] : v[i]) + recv;
    }
    double c = (h8 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h8 ? a2[0] : a2[1], 8);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __shfl_sync(0xffffffffu, c, ((i >> 1) & 1) * 16 + (i & 1) * 8);
}

static __global__ void __launch_bounds__(256) k_fp_linear(FpArgs a, int n_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    double* __restrict__ VX = a.pv.a[F_VX];
    double* __restrict__ VY = a.pv.a[F_VY];
    double* __restrict__ VZ = a.pv.a[F_VZ];
    const double* __restrict__ W = a.pv.a[F_W];
    // a warp takes 32 consecutive cells at a time: one read of their sizes, then only the cells of this kernel's size class
    for (int64_t r0 = warp0 * 32; r0 < nr; r0 += nwarps * 32) {
      const int64_t myr = r0 + lane;
      const int64_t my_n = myr < nr ? a.ix[a.cell_lo - 1 + myr].n_local : 0;
      unsigned todo = __ballot_sync(0xffffffffu, my_n >= 7 && my_n > n_lo);  // :35-37; smaller cells are handled by k_fp_linear_reg
      while (todo) {
        const int64_t r = r0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int64_t n = q.n_local, lo = q.start1 - 1;
        if (q.n_group2 != 0 || q.n_group1 != q.n_local) {  // only defined for sorted cells (the reference's group-2 branch throws, collision_fp.jl:57)
            if (lane == 0) atomicOr(&a.flags[0], DEVERR_PRECONDITION);
            continue;
        }
        // scale_norm_rands!: exact standardisation over the n draws of each component
        double m[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
        for (int64_t j = lane; j < n; j += 32) {
            double o[3];
            fp_normals(a, (uint32_t)cell, j, o);
            for (int d = 0; d < 3; d++) m[d] += o[d];
        }
        for (int d = 0; d < 3; d++) m[d] = wsum(m[d]) / (double)n;
        for (int64_t j = lane; j < n; j += 32) {
            double o[3];
            fp_normals(a, (uint32_t)cell, j, o);
            for (int d = 0; d < 3; d++) { const double x = o[d] - m[d]; s2[d] += x * x; }
        }
        for (int d = 0; d < 3; d++) s2[d] = sqrt((double)n / wsum(s2[d]));
        // weighted mean velocity and thermal energy (:60-77)
        double lw = 0, ux = 0, uy = 0, uz = 0;
        for (int64_t j = lane; j < n; j += 32) {
            const double w = W[lo + j];
            lw += w;
            ux += VX[lo + j] * w; uy += VY[lo + j] * w; uz += VZ[lo + j] * w;
        }
        lw = wsum(lw);
        ux = wsum(ux) / lw; uy = wsum(uy) / lw; uz = wsum(uz) / lw;
        double es_old = 0;
        for (int64_t j = lane; j < n; j += 32) {
            const double cx = VX[lo + j] - ux, cy = VY[lo + j] - uy, cz = VZ[lo + j] - uz;
            es_old += (cx * cx + cy * cy + cz * cz) * W[lo + j];
        }
        es_old = 0.5 * wsum(es_old) / lw;
        // compute_relaxation_time (:143-151)
        const double T = es_old * a.mass / ((3.0 / 2.0) * k_B);
        const double p = (lw / a.V) * k_B * T;
        const double mu = a.it.vhs_muref * pow(T / a.it.vhs_Tref, a.it.vhs_o);
        const double tau = 2.0 * mu / p;
        const double A = exp(-a.dt / tau);
        const double C = sqrt(((2.0 / 3.0) * es_old) * (1.0 - exp(-2.0 * a.dt / tau)));
        // v <- (v - u) A + C xi ; energy of the new thermal velocities (:95-110)
        double es_new = 0;
        for (int64_t j = lane; j < n; j += 32) {
            double o[3];
            fp_normals(a, (uint32_t)cell, j, o);
            const double vx = (VX[lo + j] - ux) * A + C * ((o[0] - m[0]) * s2[0]);
            const double vy = (VY[lo + j] - uy) * A + C * ((o[1] - m[1]) * s2[1]);
            const double vz = (VZ[lo + j] - uz) * A + C * ((o[2] - m[2]) * s2[2]);
            VX[lo + j] = vx; VY[lo + j] = vy; VZ[lo + j] = vz;
            es_new += (vx * vx + vy * vy + vz * vz) * W[lo + j];
        }
        es_new = 0.5 * wsum(es_new) / lw;
        const double alpha = sqrt(es_old / es_new);  // exact energy conservation (:112-123)
        __syncwarp();
        for (int64_t j = lane; j < n; j += 32) {
            VX[lo + j] = alpha * VX[lo + j] + ux;
            VY[lo + j] = alpha * VY[lo + j] + uy;
            VZ[lo + j] = alpha * VZ[lo + j] + uz;
        }
      }
    }
}

// Small cells (n <= 32 K): the whole cell lives in registers -- K particles per lane, the normals are generated ONCE, the cell is
// read once and written once.  The per-lane accumulation (ascending j) is the same as in k_fp_linear; the cross-lane sums are taken
// with the multi-value reductions above (a different association: the two kernels agree to round-off, not bit for bit).  Handles the cells with n_lo < n_local <= 32 K; the others are skipped.
template <int K>
static __global__ void __launch_bounds__(128, K == 4 ? 4 : 2) k_fp_linear_reg(FpArgs a, int n_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    double* __restrict__ VX = a.pv.a[F_VX];
    double* __restrict__ VY = a.pv.a[F_VY];
    double* __restrict__ VZ = a.pv.a[F_VZ];
    const double* __restrict__ W = a.pv.a[F_W];
    for (int64_t r0 = warp0 * 32; r0 < nr; r0 += nwarps * 32) {
      const int64_t myr = r0 + lane;
      const int64_t my_n = myr < nr ? a.ix[a.cell_lo - 1 + myr].n_local : 0;
      unsigned todo = __ballot_sync(0xffffffffu, my_n >= 7 && my_n > n_lo && my_n <= 32 * K);
      while (todo) {
        const int64_t r = r0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t cell = a.cell_lo + r;
        const Indexer q = a.ix[cell - 1];
        const int n = (int)q.n_local;
        const int64_t lo = q.start1 - 1;
        if (q.n_group2 != 0 || q.n_group1 != q.n_local) {  // see k_fp_linear
            if (lane == 0) atomicOr(&a.flags[0], DEVERR_PRECONDITION);
            continue;
        }
        double w[K], vx[K], vy[K], vz[K], o0[K], o1[K], o2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int j = lane + 32 * k;
            const bool valid = j < n;
            w[k] = valid ? W[lo + j] : 0.0;
            vx[k] = valid ? VX[lo + j] : 0.0;
            vy[k] = valid ? VY[lo + j] : 0.0;
            vz[k] = valid ? VZ[lo + j] : 0.0;
        }
        double m[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
        double lw = 0, ux = 0, uy = 0, uz = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int j = lane + 32 * k;
            o0[k] = o1[k] = o2[k] = 0.0;
            if (j < n) {
                double o[3];
                fp_normals(a, (uint32_t)cell, j, o);
                o0[k] = o[0]; o1[k] = o[1]; o2[k] = o[2];
                m[0] += o[0]; m[1] += o[1]; m[2] += o[2];
                lw += w[k];
                ux += vx[k] * w[k]; uy += vy[k] * w[k]; uz += vz[k] * w[k];
            }
        }
        {   // the sums of the normals and the weighted velocity sums in one reduction
            double r8[8] = {m[0], m[1], m[2], lw, ux, uy, uz, 0.0};
            wallreduce8(r8, lane);
            for (int d = 0; d < 3; d++) m[d] = r8[d] / (double)n;
            lw = r8[3];
            ux = r8[4] / lw; uy = r8[5] / lw; uz = r8[6] / lw;
        }
        double es_old = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane + 32 * k < n) {
                const double x0 = o0[k] - m[0], x1 = o1[k] - m[1], x2 = o2[k] - m[2];
                s2[0] += x0 * x0; s2[1] += x1 * x1; s2[2] += x2 * x2;
                const double cx = vx[k] - ux, cy = vy[k] - uy, cz = vz[k] - uz;
                es_old += (cx * cx + cy * cy + cz * cz) * w[k];
            }
        {   // the variances of the normals and the thermal energy in one reduction
            double r4[4] = {s2[0], s2[1], s2[2], es_old};
            wallreduce4(r4, lane);
            for (int d = 0; d < 3; d++) s2[d] = sqrt((double)n / r4[d]);
            es_old = 0.5 * r4[3] / lw;
        }
        const double T = es_old * a.mass / ((3.0 / 2.0) * k_B);
        const double p = (lw / a.V) * k_B * T;
        const double mu = a.it.vhs_muref * pow(T / a.it.vhs_Tref, a.it.vhs_o);
        const double tau = 2.0 * mu / p;
        const double A = exp(-a.dt / tau);
        const double C = sqrt(((2.0 / 3.0) * es_old) * (1.0 - exp(-2.0 * a.dt / tau)));
        double es_new = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane + 32 * k < n) {
                vx[k] = (vx[k] - ux) * A + C * ((o0[k] - m[0]) * s2[0]);
                vy[k] = (vy[k] - uy) * A + C * ((o1[k] - m[1]) * s2[1]);
                vz[k] = (vz[k] - uz) * A + C * ((o2[k] - m[2]) * s2[2]);
                es_new += (vx[k] * vx[k] + vy[k] * vy[k] + vz[k] * vz[k]) * w[k];
            }
        es_new = 0.5 * wsum(es_new) / lw;
        const double alpha = sqrt(es_old / es_new);
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int j = lane + 32 * k;
            if (j < n) {
                VX[lo + j] = alpha * vx[k] + ux;
                VY[lo + j] = alpha * vy[k] + uy;
                VZ[lo + j] = alpha * vz[k] + uz;
            }
        }
      }
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_fp_linear(mb_ctx* ctx, const mb_interaction* it, double mass, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi,
                            int64_t species, double dt, double V, uint32_t timestep, uint32_t substream) {
    MB_ARG(ctx && it && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_ARG(mass > 0 && V > 0, "mass, V");
    MB_CUDA(cudaSetDevice(ctx->device));
    FpArgs a;
    a.pv = pv->cur;
    a.ix = pia->d_indexer + (species - 1) * pia->n_cells;
    a.cell_lo = cell_lo; a.cell_hi = cell_hi;
    a.it = *it;
    a.mass = mass; a.dt = dt; a.V = V;
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, species, species);
    a.flags = ctx->d_flags;
    ProfScope ps(ctx, PROF_FP);
    ctx->state_gen++;
    // three size classes, each kernel skips the cells of the others: n <= 128 and n <= 256 in registers, larger cells streamed
    const int64_t nr = cell_hi - cell_lo + 1;
    k_fp_linear_reg<4><<<grid_for(nr * 32, 128, 12), 128, 0, ctx->stream>>>(a, 0);
    MB_LAUNCH_CHECK(ctx);
    k_fp_linear_reg<8><<<grid_for(nr * 32, 128, 8), 128, 0, ctx->stream>>>(a, 128);
    MB_LAUNCH_CHECK(ctx);
    k_fp_linear<<<grid_for(nr * 32, 256, 8), 256, 0, ctx->stream>>>(a, 256);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}

// compute_props! (properties/physical_props.jl:104-154), compute_props_with_total_moments! (:168-245),
// compute_props_sorted! (:317-454), avg_props! (:281-299), clear_props! (:256-266) as segmented moment reductions.
//
// One thread group per (cell, species): a warp for ordinary DSMC cells, a 256-thread CTA for big (0-D) cells.  The
// reference's two-pass formulation is kept (mean first, then the centred second moment) -- a one-pass
// sum(w v^2) - n vbar^2 loses ~6 digits at |vbar| = 500 m/s, sigma = 250 m/s, and the parity bar is 1e-12 relative.
// The second pass re-reads the cell from L1/L2 (a 1000-particle cell is 32 KB), so HBM traffic stays at 32 B/particle.
#include "mb_common.cuh"

namespace mb {

struct PropsArgs {
    SoA pv;
    const Indexer* ix;     // indexer row of the species
    int64_t cell_lo, cell_hi;  // 1-based inclusive
    double mass;
    double *np, *n, *v, *T, *moments;  // already offset to the species
    int sorted;            // group 1 only
    int with_moments, n_moments;
    const int32_t* powers;
    double moment_factor, moment_vref;  // physical_props.jl:176-177
    double n_scale;        // 1 or inv_V (ndens variant :425)
    const int* run_if_flag;  // nullable: run only if (*run_if_flag != 0) == run_if_nonzero (the sort's cached moments are not valid)
    int run_if_nonzero;
    double* lpa;           // nullable: phys_props.lpa[species] = length(particles[species]) (:151), written by the kernel (no host sync)
    double lpa_val;
    int64_t n_lo;          // k_props<32>: cells with more than n_lo particles (the smaller ones are taken by k_props_reg)
};

template <int G>
__device__ __forceinline__ double group_sum(double x, double* sh) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (G == 32) return x;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[wid] = x;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < G / 32; i++) t += sh[i];
    return t;
}

template <int G>
__global__ void __launch_bounds__(256) k_props(PropsArgs a) {
    __shared__ double sh[8];
    if (a.run_if_flag != nullptr && (*a.run_if_flag != 0) != (a.run_if_nonzero != 0)) return;
    const int tid = G == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int64_t grp0 = G == 32 ? ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) : blockIdx.x;
    const int64_t ngrp = G == 32 ? (((int64_t)gridDim.x * blockDim.x) >> 5) : gridDim.x;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const double* __restrict__ W = a.pv.a[F_W];
    const double* __restrict__ VX = a.pv.a[F_VX];
    const double* __restrict__ VY = a.pv.a[F_VY];
    const double* __restrict__ VZ = a.pv.a[F_VZ];
    if (a.lpa != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.lpa = a.lpa_val;
    for (int64_t r = grp0; r < nr; r += ngrp) {
        const int64_t c = a.cell_lo - 1 + r;
        const Indexer q = a.ix[c];
        const int64_t lo1 = q.start1 - 1, n1 = q.end1 >= q.start1 ? q.end1 - q.start1 + 1 : 0;
        const int64_t lo2 = q.start2 - 1, n2 = (!a.sorted && q.n_group2 > 0) ? q.end2 - q.start2 + 1 : 0;
        const int64_t nn = n1 + n2;
        if (G == 32 && nn <= a.n_lo && nn > 0) continue;  // group-uniform; k_props_reg
        double n = 0, sx = 0, sy = 0, sz = 0;
        for (int64_t j = tid; j < nn; j += G) {
            const int64_t i = j < n1 ? lo1 + j : lo2 + (j - n1);
            const double w = W[i];
            n += w;
            sx += VX[i] * w; sy += VY[i] * w; sz += VZ[i] * w;
        }
        n = group_sum<G>(n, sh);
        sx = group_sum<G>(sx, sh); sy = group_sum<G>(sy, sh); sz = group_sum<G>(sz, sh);
        double vx = 0, vy = 0, vz = 0, T = 0;
        if (n > 0.0) {
            vx = sx / n; vy = sy / n; vz = sz / n;
            double E = 0;
            for (int64_t j = tid; j < nn; j += G) {
                const int64_t i = j < n1 ? lo1 + j : lo2 + (j - n1);
                const double cx = VX[i] - vx, cy = VY[i] - vy, cz = VZ[i] - vz;
                const double c2 = cx * cx + cy * cy + cz * cz;
                if (a.with_moments) {
                    const double nv = sqrt(c2);
                    E += W[i] * nv * nv;
                } else {
                    E += W[i] * c2;
                }
            }
            E = group_sum<G>(E, sh);
            E *= 0.5 * a.mass / (n * k_B);
            T = (2.0 / 3.0) * E;
        }
        if (a.with_moments && a.n_moments <= 8) {
            // all moments in ONE more pass over the cell; even powers (the total moments M4, M6, ... of the BKW tests,
            // physical_props.jl:205-207) are products of |c|^2 instead of pow()
            double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int pwr[8];
#pragma unroll
            for (int m = 0; m < 8; m++) pwr[m] = m < a.n_moments ? a.powers[m] : 0;
            if (n > 0.0) {
                for (int64_t j = tid; j < nn; j += G) {
                    const int64_t i = j < n1 ? lo1 + j : lo2 + (j - n1);
                    const double cx = VX[i] - vx, cy = VY[i] - vy, cz = VZ[i] - vz;
                    const double c2 = cx * cx + cy * cy + cz * cz, w = W[i];
                    const double nv = sqrt(c2);
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        if (m >= a.n_moments) break;
                        const int pw = pwr[m];
                        double val;
                        if (pw >= 0 && pw <= 32 && (pw & 1) == 0) {
                            val = 1.0;
                            double b = c2;
                            for (int e = pw >> 1; e > 0; e >>= 1) { if (e & 1) val *= b; b *= b; }
                        } else {
                            val = pow(nv, (double)pw);
                        }
                        s[m] += w * val;
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < 8; m++) {
                if (m >= a.n_moments) break;
                const double sm = group_sum<G>(s[m], sh);
                if (tid == 0) {
                    const int pw = pwr[m];
                    const double scaling = a.moment_factor * pow(a.moment_vref, -(double)(3 + pw)) * tgamma((3 + pw) / 2.0);
                    a.moments[(int64_t)c * a.n_moments + m] = sm / (scaling * n);
                }
            }
        } else if (a.with_moments) {
            for (int m = 0; m < a.n_moments; m++) {
                const int pw = a.powers[m];
                double s = 0;
                if (n > 0.0) {
                    for (int64_t j = tid; j < nn; j += G) {
                        const int64_t i = j < n1 ? lo1 + j : lo2 + (j - n1);
                        const double cx = VX[i] - vx, cy = VY[i] - vy, cz = VZ[i] - vz;
                        s += W[i] * pow(sqrt(cx * cx + cy * cy + cz * cz), (double)pw);
                    }
                }
                s = group_sum<G>(s, sh);
                if (tid == 0) {
                    const double scaling = a.moment_factor * pow(a.moment_vref, -(double)(3 + pw)) * tgamma((3 + pw) / 2.0);
                    a.moments[(int64_t)c * a.n_moments + m] = s / (scaling * n);
                }
            }
        }
        if (tid == 0) {
            a.np[c] = (double)nn;
            a.n[c] = n * a.n_scale;
            a.T[c] = T;
            a.v[3 * c + 0] = vx; a.v[3 * c + 1] = vy; a.v[3 * c + 2] = vz;
        }
    }
}

// Cells of 1 .. 32 K particles (both groups together): warp per cell with the cell in registers -- read once from HBM, both passes
// of the reference's two-pass formulation from registers.  Same per-lane accumulation order and butterfly as k_props<32>, so the
// results are bit-identical.  A warp takes 32 consecutive cells at a time (one read of their sizes).
template <int K>
__global__ void __launch_bounds__(128) k_props_reg(PropsArgs a, int n_lo) {
    if (a.run_if_flag != nullptr && (*a.run_if_flag != 0) != (a.run_if_nonzero != 0)) return;
    const int lane = threadIdx.x & 31;
    const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const double* __restrict__ W = a.pv.a[F_W];
    const double* __restrict__ VX = a.pv.a[F_VX];
    const double* __restrict__ VY = a.pv.a[F_VY];
    const double* __restrict__ VZ = a.pv.a[F_VZ];
    for (int64_t r0 = gw * 32; r0 < nr; r0 += nwarps * 32) {
        const int64_t myr = r0 + lane;
        int64_t my_nn = 0;
        if (myr < nr) {
            const Indexer q = a.ix[a.cell_lo - 1 + myr];
            my_nn = (q.end1 >= q.start1 ? q.end1 - q.start1 + 1 : 0) + ((!a.sorted && q.n_group2 > 0) ? q.end2 - q.start2 + 1 : 0);
        }
        unsigned todo = __ballot_sync(0xffffffffu, my_nn > n_lo && my_nn <= 32 * K);
        while (todo) {
            const int64_t c = a.cell_lo - 1 + r0 + (__ffs(todo) - 1);
            todo &= todo - 1;
            const Indexer q = a.ix[c];
            const int64_t lo1 = q.start1 - 1, n1 = q.end1 >= q.start1 ? q.end1 - q.start1 + 1 : 0;
            const int64_t lo2 = q.start2 - 1, n2 = (!a.sorted && q.n_group2 > 0) ? q.end2 - q.start2 + 1 : 0;
            const int nn = (int)(n1 + n2);
            double w[K], vx[K], vy[K], vz[K];
            double n = 0, sx = 0, sy = 0, sz = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int j = lane + 32 * k;
                w[k] = 0.0; vx[k] = 0.0; vy[k] = 0.0; vz[k] = 0.0;
                if (j < nn) {
                    const int64_t i = j < n1 ? lo1 + j : lo2 + (j - n1);
                    w[k] = W[i]; vx[k] = VX[i]; vy[k] = VY[i]; vz[k] = VZ[i];
                    n += w[k];
                    sx += vx[k] * w[k]; sy += vy[k] * w[k]; sz += vz[k] * w[k];
                }
            }
            n = group_sum<32>(n, nullptr);
            sx = group_sum<32>(sx, nullptr); sy = group_sum<32>(sy, nullptr); sz = group_sum<32>(sz, nullptr);
            double mx = 0, my = 0, mz = 0, T = 0;
            if (n > 0.0) {
                mx = sx / n; my = sy / n; mz = sz / n;
                double E = 0;
#pragma unroll
                for (int k = 0; k < K; k++)
                    if (lane + 32 * k < nn) {
                        const double cx = vx[k] - mx, cy = vy[k] - my, cz = vz[k] - mz;
                        E += w[k] * (cx * cx + cy * cy + cz * cz);
                    }
                E = group_sum<32>(E, nullptr);
                E *= 0.5 * a.mass / (n * k_B);
                T = (2.0 / 3.0) * E;
            }
            if (lane == 0) {
                a.np[c] = (double)nn;
                a.n[c] = n * a.n_scale;
                a.T[c] = T;
                a.v[3 * c + 0] = mx; a.v[3 * c + 1] = my; a.v[3 * c + 2] = mz;
            }
        }
    }
}

// compute_props_sorted! right after a band sort: the gather pass already holds np, n, vbar and sum w |v - vbar|^2 per cell
static __global__ void k_props_cached(PropsArgs a, const double* __restrict__ pcache, const int* flags, int valid_if_general) {
    // flags != nullptr: only one of the two sort paths fills the cache; flags[2] != 0 tells that the general path ran
    if (flags != nullptr && (flags[2] != 0) != (valid_if_general != 0)) return;
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = a.cell_lo - 1 + r;
        const double* pc = pcache + 6 * c;
        const double n = pc[1];
        a.np[c] = pc[0];
        a.n[c] = n * a.n_scale;
        a.v[3 * c + 0] = pc[2]; a.v[3 * c + 1] = pc[3]; a.v[3 * c + 2] = pc[4];
        double T = 0.0;
        if (n > 0.0) {
            const double E = pc[5] * (0.5 * a.mass / (n * k_B));
            T = (2.0 / 3.0) * E;
        }
        a.T[c] = T;
    }
}

static __global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, int64_t n, double a) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i] * a;
}

static int props_launch(mb_ctx* ctx, mb_pv* const* pvs, mb_pia* pia, const double* masses, mb_props* P, int sorted, int with_moments, double n_scale,
                        int64_t cell_lo, int64_t cell_hi) {
    const int64_t nc = pia->n_cells;
    ProfScope ps(ctx, PROF_PROPS);
    for (int64_t s = 0; s < pia->n_species; s++) {
        MB_ARG(pvs[s] != nullptr, "pvs[s] == NULL");
        PropsArgs a;
        a.pv = pvs[s]->cur;
        a.ix = pia->d_indexer + s * nc;
        a.cell_lo = cell_lo; a.cell_hi = cell_hi;
        a.mass = masses[s];
        a.np = P->np + s * nc; a.n = P->n + s * nc; a.T = P->T + s * nc; a.v = P->v + 3 * s * nc;
        a.moments = P->moments ? P->moments + s * nc * P->n_moments : nullptr;
        a.sorted = sorted;
        a.with_moments = with_moments && P->n_moments > 0;
        a.n_moments = (int)P->n_moments;
        a.powers = P->d_powers;
        a.moment_factor = 4 * M_PI * std::pow(masses[s] / (twopi * k_B * P->Tref), 1.5) * 0.5;  // :176
        a.moment_vref = std::pow(masses[s] / (2 * k_B * P->Tref), 0.5);                         // :177
        a.n_scale = n_scale;
        a.run_if_flag = nullptr;
        a.run_if_nonzero = 1;
        const int64_t nr = cell_hi - cell_lo + 1;
        if (sorted && ctx->pc_gen == ctx->state_gen && ctx->pc_pv == (void*)pvs[s] && ctx->pc_pia == (void*)pia && ctx->pc_species == (int)s + 1 &&
            ctx->scratch[10] != nullptr) {
            const bool both = ctx->pc_general && ctx->pc_band;
            k_props_cached<<<grid_for(nr, 256), 256, 0, ctx->stream>>>(a, (const double*)ctx->scratch[10], both ? nullptr : ctx->d_flags,
                                                                      ctx->pc_general ? 1 : 0);
            MB_LAUNCH_CHECK(ctx);
            if (both) continue;                // both sort paths fill the cache: nothing left to compute
            a.run_if_flag = ctx->d_flags + 2;  // the regular kernel below only runs if the sort took the path that does not fill the cache
            a.run_if_nonzero = ctx->pc_band ? 1 : 0;
        }
        const int64_t avg = (pia->n_bound[s] > 0 ? pia->n_bound[s] : pvs[s]->cap) / (nc > 0 ? nc : 1);
        a.lpa = sorted ? nullptr : P->lpa + s;  // phys_props.lpa[species] = length(particles[species]) :151
        a.lpa_val = (double)pvs[s]->cap;
        a.n_lo = 0;
        if (avg > 4096) {
            k_props<256><<<(int)(nr < N_SM * 8 ? nr : N_SM * 8), 256, 0, ctx->stream>>>(a);
        } else {
            if (!a.with_moments) {  // small cells from registers; the streaming kernel takes the rest
                a.n_lo = 256;
                k_props_reg<4><<<grid_for(nr, 128, 12), 128, 0, ctx->stream>>>(a, 0);
                MB_LAUNCH_CHECK(ctx);
                k_props_reg<8><<<grid_for(nr, 128, 8), 128, 0, ctx->stream>>>(a, 128);
                MB_LAUNCH_CHECK(ctx);
            }
            k_props<32><<<grid_for(nr * 32, 256, 8), 256, 0, ctx->stream>>>(a);
        }
        MB_LAUNCH_CHECK(ctx);
    }
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_props_create(mb_ctx* ctx, int64_t n_cells, int64_t n_species, int64_t n_moments, const int32_t* powers, double Tref, int32_t ndens,
                    mb_props** out) {
    MB_ARG(ctx && out && n_cells > 0 && n_species > 0 && n_moments >= 0, "props_create");
    MB_ARG(n_moments == 0 || powers != nullptr, "moment_powers == NULL");
    MB_CUDA(cudaSetDevice(ctx->device));
    mb_props* p = new mb_props();
    p->ctx = ctx;
    p->n_cells = n_cells; p->n_species = n_species; p->n_moments = n_moments;
    p->ndens_not_Np = ndens;
    p->Tref = Tref;
    p->powers.assign(powers, powers + n_moments);
    const size_t N = (size_t)n_cells * n_species;
    const size_t total = (size_t)n_species + N * 6 + N * (size_t)n_moments;
    double* base = nullptr;
    MB_CUDA(cudaMalloc(&base, total * 8));
    MB_CUDA(cudaMemsetAsync(base, 0, total * 8, ctx->stream));
    p->lpa = base;
    p->np = base + n_species;
    p->n = p->np + N;
    p->T = p->n + N;
    p->v = p->T + N;
    p->moments = n_moments > 0 ? p->v + 3 * N : nullptr;
    p->d_powers = nullptr;
    if (n_moments > 0) {
        MB_CUDA(cudaMalloc(&p->d_powers, (size_t)n_moments * 4));
        MB_CUDA(cudaMemcpyAsync(p->d_powers, p->powers.data(), (size_t)n_moments * 4, cudaMemcpyHostToDevice, ctx->stream));
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *out = p;
    return MB_OK;
}
int mb_props_destroy(mb_props* p) {
    if (!p) return MB_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->lpa);
    if (p->d_powers) cudaFree(p->d_powers);
    delete p;
    return MB_OK;
}
int mb_props_download(mb_props* p, double* lpa, double* np, double* n, double* v, double* T, double* moments) {
    MB_ARG(p != nullptr, "NULL");
    MB_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    const size_t N = (size_t)p->n_cells * p->n_species;
    if (lpa) MB_CUDA(cudaMemcpyAsync(lpa, p->lpa, (size_t)p->n_species * 8, cudaMemcpyDeviceToHost, st));
    if (np) MB_CUDA(cudaMemcpyAsync(np, p->np, N * 8, cudaMemcpyDeviceToHost, st));
    if (n) MB_CUDA(cudaMemcpyAsync(n, p->n, N * 8, cudaMemcpyDeviceToHost, st));
    if (T) MB_CUDA(cudaMemcpyAsync(T, p->T, N * 8, cudaMemcpyDeviceToHost, st));
    if (v) MB_CUDA(cudaMemcpyAsync(v, p->v, N * 24, cudaMemcpyDeviceToHost, st));
    if (moments && p->moments) MB_CUDA(cudaMemcpyAsync(moments, p->moments, N * (size_t)p->n_moments * 8, cudaMemcpyDeviceToHost, st));
    return mb_sync(p->ctx);
}
int mb_props_clear(mb_props* p) {
    MB_ARG(p != nullptr, "NULL");
    MB_CUDA(cudaSetDevice(p->ctx->device));
    const size_t N = (size_t)p->n_cells * p->n_species;
    MB_CUDA(cudaMemsetAsync(p->lpa, 0, ((size_t)p->n_species + N * 6 + N * (size_t)p->n_moments) * 8, p->ctx->stream));
    return MB_OK;
}
int mb_props_avg(mb_props* avg, mb_props* p, int64_t n_avg) {  // physical_props.jl:281-299 (lpa, np, n, v, T; not the moments)
    MB_ARG(avg && p && n_avg > 0, "avg_props");
    if (avg->n_cells != p->n_cells || avg->n_species != p->n_species || avg->ndens_not_Np != p->ndens_not_Np) {
        set_error("avg_props!: inconsistent PhysProps (physical_props.jl:282-284)");
        return MB_ERR_ARG;
    }
    MB_CUDA(cudaSetDevice(p->ctx->device));
    const int64_t N = p->n_cells * p->n_species;
    const int64_t n = p->n_species + N * 6;  // lpa, np, n, T, v are contiguous
    k_axpy<<<grid_for(n, 256), 256, 0, p->ctx->stream>>>(avg->lpa, p->lpa, n, 1.0 / (double)n_avg);
    MB_LAUNCH_CHECK(p->ctx);
    return MB_OK;
}
int mb_compute_props(mb_ctx* ctx, mb_pv* const* pvs, mb_pia* pia, const double* masses, mb_props* props, int32_t with_moments) {
    MB_ARG(ctx && pvs && pia && masses && props, "NULL");
    MB_ARG(props->n_cells == pia->n_cells && props->n_species == pia->n_species, "props shape != pia shape");
    MB_CUDA(cudaSetDevice(ctx->device));
    return props_launch(ctx, pvs, pia, masses, props, 0, with_moments, 1.0, 1, pia->n_cells);
}
int mb_compute_props_sorted(mb_ctx* ctx, mb_pv* const* pvs, mb_pia* pia, const double* masses, mb_props* props, const mb_grid1d* grid,
                            int64_t cell_lo, int64_t cell_hi) {
    MB_ARG(ctx && pvs && pia && masses && props, "NULL");
    MB_ARG(props->n_cells == pia->n_cells && props->n_species == pia->n_species, "props shape != pia shape");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_CUDA(cudaSetDevice(ctx->device));
    const double n_scale = (props->ndens_not_Np && grid) ? 1.0 / grid->dx : 1.0;  // :425, cell volume = dx
    return props_launch(ctx, pvs, pia, masses, props, 1, 0, n_scale, cell_lo, cell_hi);
}

}  // extern "C"

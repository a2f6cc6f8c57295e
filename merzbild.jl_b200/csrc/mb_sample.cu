// Device-side initial conditions: the populations of the configs are 1e8-1e9 particles, which cannot
// sensibly be sampled on the host and copied.  Replaces
//   sample_particles_equal_weight!(rng, particles, pia, cell, species, nparticles, m, T, Fnum, xlo..zhi; distribution, v0)
//                                                                              distributions_and_sampling.jl:477-509
//   sample_particles_equal_weight!(rng, grid1duniform, particles, pia, species, species_data, ppc::Integer | ndens::Float64, T, Fnum[, cell_chunk])
//                                                                              grids/grid_uniform1D.jl:117-219
//   sample_maxwellian! :432-443, sample_bkw! :195-213, sample_on_grid! :312-346 (+ evaluate_distribution_on_grid! :253-268)
// for a range of cells in one launch.
//
// RNG: one Philox stream per cell, (seed, OP_SAMPLE, substream, timestep, entity = cell).  Within the cell's stream the draws
// are consumed in the reference's order for that cell -- [R for the fractional particle (ndens variant)], 3 per particle for
// the positions (x, y, z; particle-major), then the velocities (Maxwellian: vn, vr, theta1, theta2 per particle; BKW: 6 per
// particle for chi_5, then all polar angles, then all azimuths) -- so that the CPU oracle replays it draw for draw.  Because
// the stream is counter based every thread computes the draws of its own particle directly.
#include "mb_common.cuh"
#include "mb_scan.cuh"
#include "mb_jlexp.h"

namespace mb {

// draw d (0-based) of the stream (key, entity, timestep, opword): the (d & 1)-th double of block d >> 1
struct CellStream {
    uint32_t k0, k1, c1, c2, c3;
    __device__ __forceinline__ double draw(int64_t d) const {
        uint32_t o[4];
        philox4x32_10((uint32_t)(d >> 1), c1, c2, c3, k0, k1, o);
        return (d & 1) ? u64_to_unit_double(o[2], o[3]) : u64_to_unit_double(o[0], o[1]);
    }
    // draws d and d + 1 for even d: one block
    __device__ __forceinline__ void draw2(int64_t d_even, double& a, double& b) const {
        uint32_t o[4];
        philox4x32_10((uint32_t)(d_even >> 1), c1, c2, c3, k0, k1, o);
        a = u64_to_unit_double(o[0], o[1]);
        b = u64_to_unit_double(o[2], o[3]);
    }
};

struct SampleArgs {
    SoA p;
    int32_t* cell;
    Indexer* ix;         // row of the species
    int64_t* n_total;    // entry of the species
    int64_t cell_lo, nr; // cells cell_lo .. cell_lo + nr - 1 (1-based)
    int32_t* cnt;        // per cell: particles to sample
    int64_t* off;        // exclusive scan of cnt (nr + 1)
    int64_t cap;
    int* flags;
    uint32_t k0, k1, timestep, opword;
    int has_R, distribution, use_grid;
    int64_t ppc;         // >= 0: fixed count; < 0: ndens variant
    double ndens, Fnum, vscale, v0[3], box[6];
    double dx;           // grid variant: xlo = (cell - 1) * dx, xhi = cell * dx with the GLOBAL cell number
    int64_t cell_offset; // global 0-based index of local cell 1
    // sample_on_grid
    const double* tab_w;     // weights of the non-empty grid points, in (k, j, i) order
    const int32_t* tab_ijk;  // packed i | j << 10 | k << 20
    int64_t tab_n;
    double vmax, dv, noise;  // velocity grid extent, spacing, noise amplitude
    int64_t nv;
};

// ---- counts: fixed ppc, or floor(ndens V / Fnum) + [R < remainder] with R = draw 0 of the cell's stream (grid_uniform1D.jl:200-210)
__global__ void k_sample_counts(SampleArgs a) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < a.nr; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = a.ppc;
        if (a.ppc < 0) {
            const double n_in_cell = a.ndens * a.dx;
            const double ppc = n_in_cell / a.Fnum;
            n = (int64_t)floor(ppc);
            const double remainder = ppc - (double)n;
            const CellStream s{a.k0, a.k1, (uint32_t)(a.cell_lo + r), a.timestep, a.opword};
            if (s.draw(0) < remainder) n += 1;
        }
        a.cnt[r] = (int32_t)n;
    }
}

// ---- pia: the cell's particles are appended at n_total + 1 in cell order (distributions_and_sampling.jl:481-491)
__global__ void k_sample_indexer(SampleArgs a) {
    const int64_t nt0 = *a.n_total;
    const int64_t total = a.off[a.nr];
    if (nt0 + total > a.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(&a.flags[0], DEVERR_CAPACITY);
            const int64_t need = nt0 + total;
            a.flags[1] = need > 0x7fffffff ? 0x7fffffff : (int)need;
        }
        return;
    }
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < a.nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = a.cnt[r], start = nt0 + a.off[r] + 1;
        // the reference writes start1 = start, end1 = start - 1 + n even for n == 0 (:484-486)
        a.ix[a.cell_lo - 1 + r] = Indexer{n, start, start - 1 + n, n, 0, -1, 0};
    }
}
__global__ void k_sample_commit(SampleArgs a) {
    const int64_t nt0 = *a.n_total, total = a.off[a.nr];
    if (nt0 + total <= a.cap) *a.n_total = nt0 + total;
}

// ---- particles: one CTA per cell (grid-stride), threads stride over the cell's particles
__global__ void __launch_bounds__(256) k_sample_equal_weight(SampleArgs a) {
    const int64_t nt0 = *a.n_total;
    if (nt0 + a.off[a.nr] > a.cap) return;
    for (int64_t r = blockIdx.x; r < a.nr; r += gridDim.x) {
        const int64_t cell = a.cell_lo + r;
        const int64_t n = a.cnt[r];
        const int64_t base = nt0 + a.off[r];  // 0-based slot of the cell's first particle
        const CellStream s{a.k0, a.k1, (uint32_t)cell, a.timestep, a.opword};
        double xlo = a.box[0], xhi = a.box[1];
        if (a.use_grid) {
            const int64_t g = a.cell_offset + cell;  // global 1-based cell
            xlo = (double)(g - 1) * a.dx;            // grid_uniform1D.jl:78-79
            xhi = (double)g * a.dx;
        }
        const int64_t d0 = a.has_R ? 1 : 0;
        const int64_t dv0 = d0 + 3 * n;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const int64_t q = base + i;
            const double r0 = s.draw(d0 + 3 * i), r1 = s.draw(d0 + 3 * i + 1), r2 = s.draw(d0 + 3 * i + 2);
            a.p.a[F_W][q] = a.Fnum;
            a.p.a[F_X][q] = xlo + r0 * (xhi - xlo);
            a.p.a[F_Y][q] = a.box[2] + r1 * (a.box[3] - a.box[2]);
            a.p.a[F_Z][q] = a.box[4] + r2 * (a.box[5] - a.box[4]);
            a.cell[q] = (int32_t)cell;
            double vx, vy, vz;
            if (a.distribution == 0) {  // sample_maxwellian! :432-443
                const int64_t d = dv0 + 4 * i;
                const double u0 = s.draw(d), u1 = s.draw(d + 1), u2 = s.draw(d + 2), u3 = s.draw(d + 3);
                const double vn = sqrt(-log(u0)), vr = sqrt(-log(u1));
                const double th1 = twopi * u2, th2 = twopi * u3;
                double s2, c2;
                sincos(th2, &s2, &c2);
                vx = a.vscale * (vn * cos(th1)) + a.v0[0];
                vy = a.vscale * (vr * c2) + a.v0[1];
                vz = a.vscale * (vr * s2) + a.v0[2];
            } else {  // sample_bkw! :195-213; chi_5 = sqrt(sum of 5 squared normals), 3 Box-Muller pairs (the 6th normal is dropped)
                const int64_t d = dv0 + 6 * i;
                double sum = 0.0;
                for (int k = 0; k < 3; k++) {
                    const double u1 = fmax(1e-300, s.draw(d + 2 * k)), u2 = s.draw(d + 2 * k + 1);
                    const double r2 = -2.0 * log(u1);
                    double sn, c;
                    sincos(twopi * u2, &sn, &c);
                    sum += r2 * c * c;
                    if (k < 2) sum += r2 * sn * sn;
                }
                const double v_abs = sqrt(sum);
                const double Th = s.draw(dv0 + 6 * n + i) * 3.141592653589793;
                const double ph = s.draw(dv0 + 7 * n + i) * twopi;
                double st, ct, sp, cp;
                sincos(Th, &st, &ct);
                sincos(ph, &sp, &cp);
                vx = a.vscale * (v_abs * st * cp) + a.v0[0];
                vy = a.vscale * (v_abs * st * sp) + a.v0[1];
                vz = a.vscale * (v_abs * ct) + a.v0[2];
            }
            a.p.a[F_VX][q] = vx;
            a.p.a[F_VY][q] = vy;
            a.p.a[F_VZ][q] = vz;
        }
    }
}

// ---- sample_on_grid! :312-346: every cell of the range receives the same weighted velocity-grid sample (one particle per
// grid point inside the cut-off sphere), with its own noise / position draws: 6 draws per particle, particle-major
__global__ void __launch_bounds__(256) k_sample_on_grid(SampleArgs a) {
    const int64_t nt0 = *a.n_total;
    if (nt0 + a.off[a.nr] > a.cap) return;
    for (int64_t r = blockIdx.x; r < a.nr; r += gridDim.x) {
        const int64_t cell = a.cell_lo + r;
        const int64_t base = nt0 + a.off[r];
        const CellStream s{a.k0, a.k1, (uint32_t)cell, a.timestep, a.opword};
        for (int64_t i = threadIdx.x; i < a.tab_n; i += blockDim.x) {
            const int64_t q = base + i;
            const int32_t ijk = a.tab_ijk[i];
            double r0, r1, r2, r3, r4, r5;
            s.draw2(6 * i, r0, r1);
            s.draw2(6 * i + 2, r2, r3);
            s.draw2(6 * i + 4, r4, r5);
            // DVGrid: vx_grid = LinRange(-1, 1, nv) * vmax; LinRange element i = lerpi(i, nv - 1, -1.0, 1.0) = (1 - t) * a + t * b
            const double ti = (double)(ijk & 1023) / (double)(a.nv - 1), tj = (double)((ijk >> 10) & 1023) / (double)(a.nv - 1),
                         tk = (double)((ijk >> 20) & 1023) / (double)(a.nv - 1);
            const double gi = (1.0 - ti) * -1.0 + ti * 1.0, gj = (1.0 - tj) * -1.0 + tj * 1.0, gk = (1.0 - tk) * -1.0 + tk * 1.0;
            a.p.a[F_W][q] = a.tab_w[i];
            a.p.a[F_VX][q] = gi * a.vmax + a.noise * a.dv * (0.5 - r0) + a.v0[0];
            a.p.a[F_VY][q] = gj * a.vmax + a.noise * a.dv * (0.5 - r1) + a.v0[1];
            a.p.a[F_VZ][q] = gk * a.vmax + a.noise * a.dv * (0.5 - r2) + a.v0[2];
            a.p.a[F_X][q] = a.box[0] + r3 * (a.box[1] - a.box[0]);
            a.p.a[F_Y][q] = a.box[2] + r4 * (a.box[3] - a.box[2]);
            a.p.a[F_Z][q] = a.box[4] + r5 * (a.box[5] - a.box[4]);
            a.cell[q] = (int32_t)cell;
        }
    }
}

static int sample_common(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species, SampleArgs& a, uint32_t timestep,
                         uint32_t substream) {
    MB_ARG(ctx && pv && pia, "sample: NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "sample: species");
    MB_ARG(cell_lo >= 1 && cell_hi >= cell_lo && cell_hi <= pia->n_cells, "sample: cell range");
    MB_CUDA(cudaSetDevice(ctx->device));
    ctx->state_gen++;
    const int64_t nr = cell_hi - cell_lo + 1;
    a.p = pv->cur;
    a.cell = pv->cell;
    a.ix = pia->d_indexer + (species - 1) * pia->n_cells;
    a.n_total = pia->d_n_total + (species - 1);
    a.cell_lo = cell_lo;
    a.nr = nr;
    a.cap = pv->cap;
    a.flags = ctx->d_flags;
    a.k0 = (uint32_t)stream_seed(ctx);
    a.k1 = (uint32_t)(stream_seed(ctx) >> 32);
    a.timestep = timestep;
    a.opword = (OP_SAMPLE & 0xFFu) | (stream_substream(substream, species, species) << 8);
    a.cnt = (int32_t*)ctx_scratch(ctx, 4, (size_t)nr * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)(nr + 1) + gs_partial_count(nr)) * 8);
    if (!a.cnt || !p64) return MB_ERR_CUDA;
    a.off = p64;
    k_sample_counts<<<grid_for(nr, 256), 256, 0, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    int r = device_exclusive_scan(ctx, a.cnt, nr, a.off, p64 + nr + 1);
    if (r) return r;
    k_sample_indexer<<<grid_for(nr, 256), 256, 0, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    return MB_OK;
}
static int sample_finish(mb_ctx* ctx, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species, const SampleArgs& a, int64_t bound_add) {
    k_sample_commit<<<1, 1, 0, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    const int64_t s = species - 1;
    // sorted layout: the species was empty and every cell was filled in ascending order
    const bool was_empty = pia->h_valid && pia->h_n_total[s] == 0;
    pia->sorted_layout[s] = (was_empty && cell_lo == 1 && cell_hi == pia->n_cells) ? 1 : 0;
    pia->n_bound[s] += bound_add;
    pia->h_valid = false;
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_sample_particles_equal_weight(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                                     int64_t nparticles, double ndens, double mass, double T, double Fnum, const double* box6, int32_t distribution,
                                     const double* v0, uint32_t timestep, uint32_t substream) {
    MB_ARG(mass > 0 && T >= 0 && Fnum > 0, "sample: mass / T / Fnum");
    MB_ARG(distribution == 0 || distribution == 1, "sample: distribution must be 0 (Maxwellian) or 1 (BKW)");
    MB_ARG(nparticles >= 0 || (grid != nullptr && ndens >= 0), "sample: the number-density variant needs a grid");
    MB_ARG(grid != nullptr || box6 != nullptr, "sample: needs a grid or a box");
    SampleArgs a{};
    a.ppc = nparticles;
    a.has_R = nparticles < 0;
    a.ndens = ndens;
    a.Fnum = Fnum;
    a.distribution = distribution;
    a.vscale = sqrt(2 * k_B * T / mass) * (distribution == 1 ? sqrt(0.3) : 1.0);  // compute_thermal_velocity; BKW :197
    for (int d = 0; d < 3; d++) a.v0[d] = v0 ? v0[d] : 0.0;
    a.use_grid = grid != nullptr;
    if (grid) {
        a.dx = grid->dx;
        a.cell_offset = grid->cell_offset;
        a.box[0] = 0; a.box[1] = 0; a.box[2] = 0.0; a.box[3] = 1.0; a.box[4] = 0.0; a.box[5] = 1.0;  // grid_uniform1D.jl:150-152
    } else {
        for (int d = 0; d < 6; d++) a.box[d] = box6[d];
    }
    int r = sample_common(ctx, pv, pia, cell_lo, cell_hi, species, a, timestep, substream);
    if (r) return r;
    const int64_t nr = cell_hi - cell_lo + 1;
    const int64_t per = nparticles >= 0 ? nparticles : (int64_t)floor(ndens * grid->dx / Fnum) + 1;
    const int block = per <= 64 ? 64 : per <= 128 ? 128 : 256;
    int64_t nb = nr < (int64_t)N_SM * 32 ? nr : (int64_t)N_SM * 32;
    k_sample_equal_weight<<<(int)nb, block, 0, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    return sample_finish(ctx, pia, cell_lo, cell_hi, species, a, per * nr);
}

int mb_sample_on_grid(mb_ctx* ctx, int32_t vdf_kind, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species, int64_t nv, double mass,
                      double T, double n_total, const double* box6, double v_mult, double cutoff_mult, double noise, const double* v_offset,
                      uint32_t timestep, uint32_t substream, int64_t* n_sampled) {
    MB_ARG(ctx && pv && pia && box6, "sample_on_grid: NULL");
    MB_ARG(vdf_kind == 0 || vdf_kind == 1, "sample_on_grid: vdf must be 0 (Maxwellian) or 1 (BKW at t = 0)");
    MB_ARG(nv >= 2 && nv <= 1023, "sample_on_grid: nv");
    MB_ARG(mass > 0 && T > 0, "sample_on_grid: mass / T");
    // evaluate_distribution_on_grid! :253-268 -- a host-side table of <= nv^3 weights, summed in the reference's (k, j, i) order
    const double v_thermal = sqrt(2 * k_B * T / mass);
    const double vmax = v_thermal * v_mult, cutoff_v = v_thermal * cutoff_mult;
    std::vector<double> vg(nv);
    std::vector<double> g(nv);
    for (int64_t i = 0; i < nv; i++) { const double t = (double)i / (double)(nv - 1); g[i] = (1.0 - t) * -1.0 + t * 1.0; }  // LinRange(-1, 1, nv)
    for (int64_t i = 0; i < nv; i++) vg[i] = g[i] * vmax;
    const double dv = (g[1] - g[0]) * vmax;  // UnitDVGrid.dx * vx_max (:98, :119)
    std::vector<double> w;
    std::vector<int32_t> ijk;
    double wsum = 0.0;
    const double PI = 3.141592653589793;
    for (int64_t k = 0; k < nv; k++)
        for (int64_t j = 0; j < nv; j++)
            for (int64_t i = 0; i < nv; i++) {
                const double Csq = vg[i] * vg[i] + vg[j] * vg[j] + vg[k] * vg[k];
                if (sqrt(Csq) <= cutoff_v) {
                    double f;
                    if (vdf_kind == 0) f = pow(mass / (2.0 * PI * k_B * T), 1.5) * mbjl::exp(-mass * Csq / (2.0 * k_B * T));  // maxwellian :150-152
                    else f = (std::fma(5.0, 0.6, -3.0) + 2 * (1.0 - 0.6) * Csq * mass / (2 * k_B * 0.6 * T)) * mbjl::exp(-Csq * mass / (2 * k_B * 0.6 * T));  // bkw :168-177, xk(0) = 0.6; `5 xk - 3` is fused in the reference (@muladd): -1.1e-16, not 0
                    wsum += f;
                    if (f > 0.0) { w.push_back(f); ijk.push_back((int32_t)(i | (j << 10) | (k << 20))); }
                }
            }
    for (auto& f : w) f = f * n_total / wsum;
    // a weight can underflow to 0 only after the scaling; the reference keeps particles with vdf.w > 0 after normalisation
    {
        size_t o = 0;
        for (size_t q = 0; q < w.size(); q++)
            if (w[q] > 0.0) { w[o] = w[q]; ijk[o] = ijk[q]; o++; }
        w.resize(o);
        ijk.resize(o);
    }
    const int64_t tn = (int64_t)w.size();
    if (n_sampled) *n_sampled = tn;
    MB_ARG(tn > 0, "sample_on_grid: empty sample");
    SampleArgs a{};
    a.ppc = tn;
    a.has_R = 0;
    a.use_grid = 0;
    for (int d = 0; d < 6; d++) a.box[d] = box6[d];
    for (int d = 0; d < 3; d++) a.v0[d] = v_offset ? v_offset[d] : 0.0;
    a.nv = nv;
    a.vmax = vmax;
    a.dv = dv;
    a.noise = noise;
    a.tab_n = tn;
    MB_CUDA(cudaSetDevice(ctx->device));
    char* tab = (char*)ctx_scratch(ctx, 9, (size_t)tn * 12 + 256);
    if (!tab) return MB_ERR_CUDA;
    MB_CUDA(cudaMemcpyAsync(tab, w.data(), (size_t)tn * 8, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(tab + (size_t)tn * 8, ijk.data(), (size_t)tn * 4, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // w / ijk are host temporaries
    a.tab_w = (const double*)tab;
    a.tab_ijk = (const int32_t*)(tab + (size_t)tn * 8);
    int r = sample_common(ctx, pv, pia, cell_lo, cell_hi, species, a, timestep, substream);
    if (r) return r;
    const int64_t nr = cell_hi - cell_lo + 1;
    int64_t nb = nr < (int64_t)N_SM * 16 ? nr : (int64_t)N_SM * 16;
    k_sample_on_grid<<<(int)nb, 256, 0, ctx->stream>>>(a);
    MB_LAUNCH_CHECK(ctx);
    return sample_finish(ctx, pia, cell_lo, cell_hi, species, a, tn * nr);
}

}  // extern "C"

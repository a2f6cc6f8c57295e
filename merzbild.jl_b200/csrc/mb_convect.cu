// convect_particles! (convection/convection_1D.jl:130-157, with SurfProps :176-206), convect_particles_and_compute_cell!
// (:225-307), convect_single_particle! (:17-54, :72-112), reflect_particle_x! / diffuse_reflection_x! / specular
// (convection/boundary_conditions.jl:63-121), update_surface_incident!/reflected! (properties/surface_props.jl:77-131),
// surface_props_scale! (:144-160).
//
// Pure streaming kernel, one thread per particle: reads x1 and v1 (16 B), writes x1 (8 B) [+ cell id 4 B]; only a
// particle that reaches a wall touches v2, v3, w (diffuse reflection rewrites the whole velocity).  Wall draws come from
// the particle's own Philox stream (OP_CONVECT, substream, timestep, logical index), the same stream the CPU oracle uses.
#include "mb_common.cuh"

namespace mb {

struct ConvectArgs {
    SoA pv;
    int32_t* cell;
    const Indexer* ix;
    const int64_t* n_total;
    int64_t n_cells;
    double L, inv_dx, min_x, max_x;
    int64_t cell_offset;
    double v_sq[2];        // reflection_velocities_sq[wall, species] = 2 k_B T_wall / m (boundary_conditions.jl:45-50)
    double wall_v[2][3];
    double acc[2];
    double dt;
    uint64_t seed;
    uint32_t timestep, substream;
    int compute_cell;
    double* surf;          // nullable: 2 x 11 accumulators (unscaled)
};

__device__ __forceinline__ void surf_update(double* s, int wall, double w, double vx, double vy, double vz, bool incident) {
    // surface_props.jl:77-98 (incident, sign +) and :111-131 (reflected, sign -); normals (1,0,0) left, (-1,0,0) right
    double* q = s + 11 * wall;
    const double nx = wall == 0 ? 1.0 : -1.0;
    const double px = w * vx, py = w * vy, pz = w * vz;
    const double pdn = px * nx;
    const double sg = incident ? 1.0 : -1.0;
    if (incident) { atomicAdd(q + 0, 1.0); atomicAdd(q + 1, w); }
    else atomicAdd(q + 2, -w);
    atomicAdd(q + 3, sg * px); atomicAdd(q + 4, sg * py); atomicAdd(q + 5, sg * pz);
    atomicAdd(q + 6, -sg * pdn);
    atomicAdd(q + 7, sg * (px - pdn * nx)); atomicAdd(q + 8, sg * py); atomicAdd(q + 9, sg * pz);
    atomicAdd(q + 10, sg * 0.5 * (px * vx + py * vy + pz * vz));
}

__device__ __forceinline__ void convect_one(const ConvectArgs& a, int64_t i) {
    double* __restrict__ X = a.pv.a[F_X];
    double* __restrict__ VX = a.pv.a[F_VX];
    double vx = VX[i];
    double x_old = X[i];
    double t_rest = a.dt;
    double x_new = fma(vx, a.dt, x_old);  // @muladd x[1] + v[1] * dt
    if (x_new >= a.L || x_new <= 0.0) {
        PhiloxStream rng(a.seed, OP_CONVECT, a.substream, a.timestep, (uint32_t)i);
        double vy = a.pv.a[F_VY][i], vz = a.pv.a[F_VZ][i];
        const double w = a.pv.a[F_W][i];
        while (x_new >= a.L || x_new <= 0.0) {
            int wall;
            double nsign;
            if (x_new >= a.L) { t_rest -= fabs((a.L - x_old) / vx); wall = 1; nsign = -1.0; x_old = a.L; }
            else { t_rest -= fabs(x_old / vx); wall = 0; nsign = 1.0; x_old = 0.0; }
            if (a.surf) surf_update(a.surf, wall, w, vx, vy, vz, true);
            // reflect_particle_x! boundary_conditions.jl:108-121
            const double acc = a.acc[wall];
            bool diffuse = acc == 1.0;
            if (acc != 0.0 && acc != 1.0) diffuse = rng.rand() < acc;
            if (diffuse) {  // diffuse_reflection_x! :79-93
                double R = fmax(1e-50, rng.rand());
                const double vn = nsign * sqrt(-a.v_sq[wall] * log(R));
                R = fmax(1e-50, rng.rand());
                const double vt = sqrt(-a.v_sq[wall] * log(R));
                R = twopi * rng.rand();
                double sn, cs;
                sincos(R, &sn, &cs);
                vx = vn + a.wall_v[wall][0];
                vy = sn * vt + a.wall_v[wall][1];
                vz = cs * vt + a.wall_v[wall][2];
            } else {
                vx = -vx;  // specular_reflection_x! :63-65
            }
            if (a.surf) surf_update(a.surf, wall, w, vx, vy, vz, false);
            x_new = fma(vx, t_rest, x_old);
        }
        VX[i] = vx;
        a.pv.a[F_VY][i] = vy;
        a.pv.a[F_VZ][i] = vz;
    }
    if (x_new < a.min_x) x_new = a.min_x;
    else if (x_new > a.max_x) x_new = a.max_x;
    X[i] = x_new;
    if (a.compute_cell) a.cell[i] = (int32_t)((int64_t)floor(x_new * a.inv_dx) - a.cell_offset) + 1;
}

static __global__ void __launch_bounds__(256) k_convect_contiguous(ConvectArgs a) {
    const int64_t n = *a.n_total;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) convect_one(a, i);
}
// species not contiguous: only the particles the pia points to move (convection_1D.jl:143-155)
static __global__ void __launch_bounds__(256) k_convect_ranges(ConvectArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < a.n_cells; c += nwarps) {
        const Indexer q = a.ix[c];
        for (int64_t j = q.start1 - 1 + lane; j < q.end1; j += 32) convect_one(a, j);
        if (q.n_group2 > 0)
            for (int64_t j = q.start2 - 1 + lane; j < q.end2; j += 32) convect_one(a, j);
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_convect_particles(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species,
                                    double mass, double* surf22, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream) {
    MB_ARG(ctx && grid && walls && pv && pia, "NULL handle");
    MB_ARG(species >= 1 && species <= pia->n_species, "species out of range");
    MB_ARG(mass > 0, "mass");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int s = (int)species - 1;
    ConvectArgs a;
    a.pv = pv->cur;
    a.cell = pv->cell;
    a.ix = pia->d_indexer + (int64_t)s * pia->n_cells;
    a.n_total = pia->d_n_total + s;
    a.n_cells = pia->n_cells;
    a.L = grid->L; a.inv_dx = grid->inv_dx; a.min_x = grid->min_x; a.max_x = grid->max_x;
    a.cell_offset = grid->cell_offset;
    for (int wl = 0; wl < 2; wl++) {
        a.v_sq[wl] = 2 * k_B * walls->T[wl] / mass;
        for (int d = 0; d < 3; d++) a.wall_v[wl][d] = walls->v[wl][d];
        a.acc[wl] = walls->accommodation[wl];
    }
    a.dt = dt;
    a.seed = ctx->seed; a.timestep = timestep; a.substream = substream;
    a.compute_cell = compute_cell;
    a.surf = nullptr;
    ctx->state_gen++;
    if (surf22) {
        a.surf = (double*)ctx_scratch(ctx, 6, 22 * 8);
        if (!a.surf) return MB_ERR_CUDA;
        MB_CUDA(cudaMemsetAsync(a.surf, 0, 22 * 8, ctx->stream));  // clear_props!(surf_props) convection_1D.jl:179
    }
    {
    ProfScope ps(ctx, PROF_CONVECT);
    if (pia->contiguous[s]) {
        const int64_t nb = pia->n_bound[s] > 0 ? pia->n_bound[s] : pv->cap;
        k_convect_contiguous<<<grid_for(nb, 256, 16), 256, 0, ctx->stream>>>(a);
    } else {
        k_convect_ranges<<<grid_for(pia->n_cells * 32, 256, 8), 256, 0, ctx->stream>>>(a);
    }
    MB_LAUNCH_CHECK(ctx);
    }
    if (surf22) {
        MB_CUDA(cudaMemcpyAsync(surf22, a.surf, 22 * 8, cudaMemcpyDeviceToHost, ctx->stream));
        int r = mb_sync(ctx);
        if (r) return r;
        const double factor = mass / dt;  // surface_props_scale! :144-160 (areas = 1); np is not scaled
        for (int e = 0; e < 2; e++)
            for (int k = 1; k < 11; k++) surf22[11 * e + k] *= factor;
    }
    return MB_OK;
}

// Shared by mb_convect.cu (thread-per-particle kernels) and mb_sort.cu (the fused convect + band-classify kernel):
// ConvectArgs and convect_one, the restatement of convect_single_particle! (convection/convection_1D.jl:17-54, :72-112).
#pragma once
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

// A particle that reaches a wall: repeated reflection until the remaining time is used up (convection_1D.jl:25-47).
// Kept out of line so that the streaming loops stay tight; stores the new velocity, returns the unclamped x1.
static __device__ __noinline__ double convect_wall(const ConvectArgs& a, int64_t i, double x_old, double vx, double x_new) {
    double t_rest = a.dt;
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
    a.pv.a[F_VX][i] = vx;
    a.pv.a[F_VY][i] = vy;
    a.pv.a[F_VZ][i] = vz;
    return x_new;
}

// convect_single_particle! (convection_1D.jl:17-54) for logical position i; returns the new x1 (already stored)
__device__ __forceinline__ double convect_one(const ConvectArgs& a, int64_t i) {
    double* __restrict__ X = a.pv.a[F_X];
    const double vx = a.pv.a[F_VX][i];
    const double x_old = X[i];
    double x_new = fma(vx, a.dt, x_old);  // @muladd x[1] + v[1] * dt
    if (x_new >= a.L || x_new <= 0.0) x_new = convect_wall(a, i, x_old, vx, x_new);
    if (x_new < a.min_x) x_new = a.min_x;
    else if (x_new > a.max_x) x_new = a.max_x;
    X[i] = x_new;
    if (a.compute_cell) a.cell[i] = (int32_t)((int64_t)floor(x_new * a.inv_dx) - a.cell_offset) + 1;
    return x_new;
}

// fused convection + band classification (mb_sort.cu); returns MB_OK and sets *done = true if it ran
int convect_band_launch(mb_ctx* ctx, const ConvectArgs& a, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species, bool* done);

}  // namespace mb

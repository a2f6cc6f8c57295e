// convect_particles! (convection/convection_1D.jl:130-157, with SurfProps :176-206), convect_particles_and_compute_cell!
// (:225-307), convect_single_particle! (:17-54, :72-112), reflect_particle_x! / diffuse_reflection_x! / specular
// (convection/boundary_conditions.jl:63-121), update_surface_incident!/reflected! (properties/surface_props.jl:77-131),
// surface_props_scale! (:144-160).
//
// Pure streaming kernel, one thread per particle: reads x1 and v1 (16 B), writes x1 (8 B) [+ cell id 4 B]; only a
// particle that reaches a wall touches v2, v3, w (diffuse reflection rewrites the whole velocity).  Wall draws come from
// the particle's own Philox stream (OP_CONVECT, substream, timestep, logical index), the same stream the CPU oracle uses.
#include "mb_convect.cuh"

namespace mb {

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

struct mb_surf;
namespace mb {
double* surf_device_ptr(mb_surf* s);            // mb_surf.cu
int surf_scale(mb_ctx* ctx, mb_surf* s, double factor);
}

static int convect_impl(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species, double mass,
                        double* surf22, mb_surf* surf, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream) {
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
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, species, species);
    a.compute_cell = compute_cell;
    a.surf = nullptr;
    ctx->state_gen++;
    if (surf) {  // device-resident SurfProps: cleared, accumulated and scaled on the stream, nothing comes back to the host
        a.surf = surf_device_ptr(surf);
        MB_CUDA(cudaMemsetAsync(a.surf, 0, 22 * 8, ctx->stream));  // clear_props!(surf_props) convection_1D.jl:179
    } else if (surf22) {
        a.surf = (double*)ctx_scratch(ctx, 6, 22 * 8);
        if (!a.surf) return MB_ERR_CUDA;
        MB_CUDA(cudaMemsetAsync(a.surf, 0, 22 * 8, ctx->stream));  // clear_props!(surf_props) convection_1D.jl:179
    }
    {
    ProfScope ps(ctx, PROF_CONVECT);
    bool fused = false;
    if (pia->contiguous[s]) {  // sorted layout: convect + band classification in one pass (mb_sort.cu)
        int r = convect_band_launch(ctx, a, grid, pv, pia, species, &fused);
        if (r) return r;
    }
    if (fused) {
    } else if (pia->contiguous[s]) {
        const int64_t nb = pia->n_bound[s] > 0 ? pia->n_bound[s] : pv->cap;
        k_convect_contiguous<<<grid_for(nb, 256, 16), 256, 0, ctx->stream>>>(a);
    } else {
        k_convect_ranges<<<grid_for(pia->n_cells * 32, 256, 8), 256, 0, ctx->stream>>>(a);
    }
    if (!fused) MB_LAUNCH_CHECK(ctx);
    }
    if (surf) return surf_scale(ctx, surf, mass / dt);  // surface_props_scale! :144-160 (areas = 1)
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

extern "C" int mb_convect_particles(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species,
                                    double mass, double* surf22, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream) {
    return convect_impl(ctx, grid, walls, pv, pia, species, mass, surf22, nullptr, dt, compute_cell, timestep, substream);
}
extern "C" int mb_convect_particles_surf(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species,
                                         double mass, mb_surf* surf, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream) {
    MB_ARG(surf != nullptr, "surf == NULL");
    return convect_impl(ctx, grid, walls, pv, pia, species, mass, nullptr, surf, dt, compute_cell, timestep, substream);
}

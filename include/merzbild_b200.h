/* merzbild_b200.h -- C ABI of libmerzbild_b200.so: the B200 (sm_100a) implementation of Merzbild.jl's
 * per-timestep DSMC particle pipeline.  This is the drop-in boundary: the Julia shim
 * (merzbild.jl_b200/julia/MerzbildB200.jl) binds these symbols with `ccall`, the Python test harness
 * (merzbild.jl_b200/merzbild_b200/) binds the very same symbols with `ctypes`.  Plain pointers, sizes and
 * scalars only -- no torch / CUDA types in any signature.
 *
 * Every entry point names the reference interface it replaces (file:line under /root/reference/src).
 *
 * Conventions (identical to the reference so a download *is* the Julia struct content):
 *   - cells, species and logical particle positions are 1-based; ranges are inclusive; an empty range is (0,-1)
 *     (particles.jl:84).
 *   - a ParticleIndexer is 7 x int64: n_local, start1, end1, n_group1, start2, end2, n_group2 (particles.jl:56-66);
 *     the device pia stores them as int64[n_species][n_cells][7] (== Julia's column-major indexer[cell, species]).
 *   - the device ParticleVector is SoA fp64 (w, vx, vy, vz, x, y, z); the reference's `index` indirection
 *     (particles.jl:225) is the identity on the device because the sort physically reorders the particles, so
 *     pv[i] (logical position i) is element i-1 of every array and restore_particle_ordering! is a no-op.
 *   - every per-cell operator of the reference (ntc!, swpm!, fp_linear!, merge_octree_N2_based!) takes an
 *     inclusive cell range [cell_lo, cell_hi]: lo == hi is the reference's per-cell call, the full range is the
 *     production path (one launch for all cells).
 *   - the reference's `rng` argument becomes (ctx seed, timestep, substream): Philox4x32-10 streams keyed per
 *     (operator, substream, timestep, entity) where entity is the cell (collisions, merging) or the logical
 *     particle index (wall reflections).  See DESIGN.md "RNG convention".
 *   - all calls are asynchronous on the context's stream; mb_sync and every download synchronise.
 *   - return value: 0 = MB_OK, otherwise an mb_status; mb_last_error_string() describes the last failure.
 *     There is NO CPU fallback: with no CUDA device mb_ctx_create returns MB_ERR_NO_DEVICE.
 */
#ifndef MERZBILD_B200_H
#define MERZBILD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MB_OK = 0,
    MB_ERR_NO_DEVICE = 1,   /* no CUDA device / driver */
    MB_ERR_CUDA = 2,        /* a CUDA runtime call failed */
    MB_ERR_ARG = 3,         /* invalid argument */
    MB_ERR_CAPACITY = 4,    /* particle arrays too small (the reference would resize! by DELTA_PARTICLES) */
    MB_ERR_PRECONDITION = 5,/* operator precondition violated (e.g. VW ntc on a cell whose group2 is not at the tail) */
    MB_ERR_NCCL = 6,        /* NCCL missing or a NCCL call failed */
    MB_ERR_UNSUPPORTED = 7
} mb_status;

typedef struct mb_ctx mb_ctx;       /* one per GPU: device, stream, Philox seed, scratch, (optional) NCCL communicator */
typedef struct mb_pv mb_pv;         /* ParticleVector (particles.jl:194-212), device SoA */
typedef struct mb_pia mb_pia;       /* ParticleIndexerArray (particles.jl:104-141) */
typedef struct mb_cf mb_cf;         /* CollisionFactors per cell for one species pair (collision_ntc.jl:18-25, :46-155) */
typedef struct mb_props mb_props;   /* PhysProps (physical_props.jl:24-37) */
typedef struct mb_surf mb_surf;     /* SurfProps (surface_props.jl:22-50) of the two walls of a 1-D grid, one species */

/* Grid1DUniform (grids/grid_uniform1D.jl:49-86); fill with mb_grid1d_init */
typedef struct {
    double L;             /* global domain length: walls at x = 0 and x = L */
    int64_t n_cells;      /* cells owned by this context (== global count on one GPU) */
    double dx, inv_dx, min_x, max_x;
    int64_t cell_offset;  /* global 0-based index of the first owned cell (0 on one GPU); slab partition, see mb_grid1d_slab */
} mb_grid1d;

/* MaxwellWalls1D (convection/boundary_conditions.jl:29-53): wall 0 = left (x=0), wall 1 = right (x=L) */
typedef struct {
    double T[2];
    double v[2][3];
    double accommodation[2];
} mb_walls1d;

/* Interaction (collisions/collision_utils.jl:73-82) */
typedef struct {
    double m_r, mu1, mu2, vhs_d, vhs_o, vhs_Tref, vhs_muref, vhs_factor;
} mb_interaction;

/* OctreeN2Merge parameters (merging/merging_octree_N2.jl:131-179, ctor :235-246) */
typedef struct {
    int32_t split;               /* 1 OctreeBinMidSplit, 2 OctreeBinMeanSplit (:12) */
    int32_t init_bin_bounds;     /* 1 MinMaxVel, 2 MinMaxVelSym, 3 C (:24) */
    int32_t bin_bounds_compute;  /* 1 inherit, 2 recompute (:34) */
    int32_t max_depth;
    int64_t max_Nbins;
} mb_octree_params;

const char* mb_last_error_string(void);
int mb_version(void);

/* ---- context ---- */
int mb_ctx_create(int device, uint64_t seed, mb_ctx** out);
int mb_ctx_destroy(mb_ctx* ctx);
int mb_sync(mb_ctx* ctx);                      /* stream sync + device-side error flags -> status */
void* mb_ctx_stream(mb_ctx* ctx);              /* the cudaStream_t every kernel of this context is launched on */
int mb_ctx_set_seed(mb_ctx* ctx, uint64_t seed);
int64_t mb_ctx_kernel_launches(mb_ctx* ctx);   /* number of kernels this context has launched so far */
/* CUDA-event timing on the context's stream (bench.py uses these; torch.cuda.Event would see only torch's stream) */
int mb_timer_start(mb_ctx* ctx);
int mb_timer_stop(mb_ctx* ctx, double* elapsed_ms);   /* synchronises */
int mb_flush_l2(mb_ctx* ctx);                  /* writes a 256 MiB scratch buffer (> 126 MB L2) */
/* per-kernel CUDA-event profiling (off by default).  Sections: 0 sort.classify, 1 sort.scan, 2 sort.scatter, 3 sort.general,
 * 4 ntc, 5 convect, 6 props, 7 merge, 8 fp, 9 exchange, 10 squash, 11 sort.extras.  mb_prof_read synchronises, returns the summed device
 * time and the number of timed launches of a section since the last read, and resets it. */
int mb_prof_enable(mb_ctx* ctx, int32_t on);
int mb_prof_read(mb_ctx* ctx, int32_t section, double* total_ms, int64_t* launches);

/* ---- Grid1DUniform(L, nx; wall_offset=1e-12) grid_uniform1D.jl:72-86 ---- */
int mb_grid1d_init(double L, int64_t nx, double wall_offset, mb_grid1d* out);
/* contiguous balanced slab of cells for `rank` of `nranks` (same rule as ChunkSplitters.chunks(1:nx; n), used by
 * simulations/1D/couette_multithreaded.jl:30-31): the first nx mod nranks slabs are one cell longer */
int mb_grid1d_slab(const mb_grid1d* global, int rank, int nranks, mb_grid1d* out);

/* ---- ParticleVector: ParticleVector(np) particles.jl:210-212, length :255, resize! :269-298, getindex/setindex! :225,:241 ---- */
int mb_pv_create(mb_ctx* ctx, int64_t np, mb_pv** out);
int mb_pv_destroy(mb_pv* pv);
int64_t mb_pv_length(mb_pv* pv);
int mb_pv_resize(mb_pv* pv, int64_t np);       /* keeps contents; the device never grows implicitly (MB_ERR_CAPACITY instead) */
/* rows = n x 7 doubles (w, vx, vy, vz, x, y, z) for logical positions lo .. lo+n-1 (1-based): pv[i] = Particle(...) */
int mb_pv_upload_rows(mb_pv* pv, int64_t lo, int64_t n, const double* rows);
int mb_pv_download_rows(mb_pv* pv, int64_t lo, int64_t n, double* rows);
/* SoA variant: 7 host arrays of n doubles each; may be pinned memory (then the copies are truly asynchronous) */
int mb_pv_upload_soa(mb_pv* pv, int64_t lo, int64_t n, const double* w, const double* vx, const double* vy, const double* vz,
                     const double* x, const double* y, const double* z);
int mb_pv_download_soa(mb_pv* pv, int64_t lo, int64_t n, double* w, double* vx, double* vy, double* vz, double* x, double* y, double* z);
/* pv.cell (particles.jl:197): int64 host view of the per-position cell ids written by convect_particles_and_compute_cell! */
int mb_pv_upload_cell(mb_pv* pv, int64_t lo, int64_t n, const int64_t* cell);
int mb_pv_download_cell(mb_pv* pv, int64_t lo, int64_t n, int64_t* cell);
/* raw device pointers of the current SoA buffers (7 pointers; valid until the next sort/resize) -- for zero-copy interop */
int mb_pv_device_ptrs(mb_pv* pv, void** out7);

/* ---- ParticleIndexerArray(n_cells, n_species) particles.jl:131-141 ---- */
int mb_pia_create(mb_ctx* ctx, int64_t n_cells, int64_t n_species, mb_pia** out);
int mb_pia_destroy(mb_pia* pia);
/* indexer: int64[n_species][n_cells][7]; n_total: int64[n_species]; contiguous: uint8[n_species]; any pointer may be NULL */
int mb_pia_upload(mb_pia* pia, const int64_t* indexer, const int64_t* n_total, const uint8_t* contiguous);
int mb_pia_download(mb_pia* pia, int64_t* indexer, int64_t* n_total, uint8_t* contiguous);
int64_t mb_pia_n_total(mb_pia* pia, int64_t species);   /* synchronises if the host mirror is stale */
/* debugging validators: check_pia_is_correct particles.jl:863-907, check_unique_index :942-988 (index is the identity here) */
int mb_check_pia(mb_pia* pia, int64_t species, int32_t* ok, int64_t* where);

/* ---- sort_particles!(gridsort, grid, pv, pia, species) grid_sorting.jl:58-113 (grid != NULL)
 *      sort_particles!(gridsort, pv, pia, species)       grid_sorting.jl:128-182 (grid == NULL: pv.cell known)
 *      squashes first if the species is not contiguous (:69-71).  The GridSortInPlace scratch lives in the context. ---- */
int mb_sort_particles(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t species);
/* which algorithm the last mb_sort_particles used: 1 = band (nearly-sorted fast path, hybrid with extras), 2 = general,
 * 3 = segments (grid == NULL and nobody changed cell: every cell's group 1 and group 2 are concatenated, no ranking) */
int mb_sort_last_path(mb_ctx* ctx);
/* band half-width w of the fast path: 0 disables it (general path only), else 1, 2, 4, 8 or 15; default 2.  Choose w of the order of
 * 3 sigma_v dt / dx: particles that move further than w cells in one step ("extras") are still sorted correctly by the band path
 * (hybrid: they are ranked separately), only more slowly, so w need not bound the displacement. */
int mb_sort_set_band_halfwidth(mb_ctx* ctx, int32_t w);
/* number of extras (band outliers + slab-exchange arrivals) the last band-path sort placed; -1 if the general path ran.  Synchronises. */
int64_t mb_sort_last_extras(mb_ctx* ctx);
/* pass B of the last band-path sort: 0 = a warp per old cell (k_band_scatter), 1 = a CTA per tile of cells with TMA bulk copies
 * (k_band_tile; chosen for bands of w >= 4 and cells of 24 .. 2048 particles).  Same result, bit for bit. */
int mb_sort_last_pass_b(mb_ctx* ctx);

/* ---- squash_pia!(pv, pia, species) particles.jl:622-682; restore_particle_ordering! :1086-1137 (no-op on device) ---- */
int mb_squash_pia(mb_ctx* ctx, mb_pv* pv, mb_pia* pia, int64_t species);
int mb_restore_particle_ordering(mb_ctx* ctx, mb_pv* pv);

/* ---- collisions ---- */
/* Interaction entry of load_interaction_data (collision_utils.jl:159-201) incl. compute_vhs_factor (:98-101) */
int mb_make_interaction(double m_i, double m_k, double vhs_d, double vhs_o, double vhs_Tref, mb_interaction* out);
/* estimate_sigma_g_w_max (collision_utils.jl:418-423) */
double mb_estimate_sigma_g_w_max(const mb_interaction* it, double m1, double m2, double T1, double T2, double Fnum, double mult_factor);
/* create_collision_factors_array for one species pair (collision_ntc.jl:46-155): per-cell sigma_g_w_max + counters */
int mb_cf_create(mb_ctx* ctx, int64_t n_cells, double sigma_g_w_max, mb_cf** out);
int mb_cf_destroy(mb_cf* cf);
int mb_cf_fill(mb_cf* cf, double sigma_g_w_max);
int mb_cf_upload(mb_cf* cf, const double* sigma_g_w_max);
/* any pointer may be NULL; arrays of n_cells */
int mb_cf_download(mb_cf* cf, double* sigma_g_w_max, int64_t* n_coll, int64_t* n_coll_performed, int64_t* n_eq_w_coll_performed);

/* ntc!(rng, cf, cd, interaction, pv, pia, cell, species, dt, V; dw_tol) collision_ntc.jl:338-380 and
 * ntc_equal_weight! :479-521 (equal_weight != 0), for cells [cell_lo, cell_hi].
 * V > 0: the cell volume (0-D usage); V <= 0 and grid != NULL: V = grid cell volume (dx), as couette drivers pass grid.cells[cell].V */
int mb_ntc(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
           double dt, double V, double dw_tol, int32_t equal_weight, uint32_t timestep, uint32_t substream);
/* two-species ntc! :412-453 / ntc_equal_weight! :554-595 */
int mb_ntc2(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv1, mb_pv* pv2, mb_pia* pia, int64_t cell_lo, int64_t cell_hi,
            int64_t s1, int64_t s2, double dt, double V, double dw_tol, int32_t equal_weight, uint32_t timestep, uint32_t substream);
/* swpm!(rng, cf_swpm, cd, interaction, pv, pia, cell, species, G, dt, V) collision_swpm.jl:201-287; cf.sigma_g_w_max holds sigma_g_max */
int mb_swpm(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
            double G, double dt, double V, uint32_t timestep, uint32_t substream);
/* fp_linear!(rng, cd_fp, interaction, species_data, pv, pia, cell, species, dt, V) collision_fp.jl:24-125 */
int mb_fp_linear(mb_ctx* ctx, const mb_interaction* it, double mass, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi,
                 int64_t species, double dt, double V, uint32_t timestep, uint32_t substream);

/* ---- convect_particles!(rng, grid, boundaries, pv, pia, species, species_data, [surf_props,] dt) convection_1D.jl:130-157,:176-206
 *      convect_particles_and_compute_cell! :225-307 (compute_cell != 0).
 *      surf22 (nullable, HOST pointer, 22 doubles): per wall (np, flux_incident, flux_reflected, force[3], normal_pressure,
 *      shear_pressure[3], kinetic_energy_flux) already scaled as surface_props_scale! does (surface_props.jl:144-160);
 *      passing it synchronises. ---- */
int mb_convect_particles(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species, double mass,
                         double* surf22, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream);

/* ---- SurfProps on the device (properties/surface_props.jl): 2 walls x 11 doubles = np, flux_incident, flux_reflected, force[3],
 *      normal_pressure, shear_pressure[3], kinetic_energy_flux; wall 0 = left.  Everything but the download is stream-ordered.
 *      mb_convect_particles_surf = convect_particles!(rng, grid, boundaries, pv, pia, species, species_data, surf_props, dt)
 *      convection_1D.jl:176-206: clears surf (:179), accumulates incident / reflected contributions in the convection kernel
 *      (surface_props.jl:77-131) and scales them (surface_props_scale! :144-160) without synchronising the stream. ---- */
int mb_surf_create(mb_ctx* ctx, mb_surf** out);
int mb_surf_destroy(mb_surf* s);
int mb_surf_clear(mb_surf* s);                                           /* clear_props!(surf_props) :173-181 */
int mb_surf_upload(mb_surf* s, const double* in22);
int mb_surf_download(mb_surf* s, double* out22);                         /* synchronises */
int mb_surf_avg(mb_surf* avg, mb_surf* cur, int64_t n_avg_timesteps);    /* avg_props!(surf_props_avg, surf_props, n) :202-222 */
/* reduce_surf_props!(target, chunks) :232-252: target = sum over the n_chunks SurfProps of this process (list order); with
 * across_ranks != 0 and a communicator of more than one rank (mb_comm_init) additionally summed over all ranks (ncclAllReduce). */
int mb_surf_reduce(mb_surf* target, mb_surf* const* chunks, int32_t n_chunks, int32_t across_ranks);
int mb_convect_particles_surf(mb_ctx* ctx, const mb_grid1d* grid, const mb_walls1d* walls, mb_pv* pv, mb_pia* pia, int64_t species, double mass,
                              mb_surf* surf, double dt, int32_t compute_cell, uint32_t timestep, uint32_t substream);

/* ---- PhysProps(n_cells, n_species, moment_powers; Tref) physical_props.jl:24-37,:55-71 ---- */
int mb_props_create(mb_ctx* ctx, int64_t n_cells, int64_t n_species, int64_t n_moments, const int32_t* moment_powers, double Tref,
                    int32_t ndens_not_Np, mb_props** out);
int mb_props_destroy(mb_props* p);
/* arrays: lpa[n_species]; np, n, T [n_species][n_cells]; v [n_species][n_cells][3]; moments [n_species][n_cells][n_moments]; NULL skips */
int mb_props_download(mb_props* p, double* lpa, double* np, double* n, double* v, double* T, double* moments);
int mb_props_clear(mb_props* p);                                                /* clear_props! :256-266 */
int mb_props_avg(mb_props* avg, mb_props* p, int64_t n_avg_timesteps);         /* avg_props! :281-299 */
/* compute_props!(particles, pia, species_data, phys_props) physical_props.jl:104-154 (with_moments == 0),
 * compute_props_with_total_moments! :168-245 (with_moments != 0); pvs = n_species handles */
int mb_compute_props(mb_ctx* ctx, mb_pv* const* pvs, mb_pia* pia, const double* masses, mb_props* props, int32_t with_moments);
/* compute_props_sorted!(particles, pia, species_data, phys_props[, grid][, cell_chunk]) :317-454 (group1 only) */
int mb_compute_props_sorted(mb_ctx* ctx, mb_pv* const* pvs, mb_pia* pia, const double* masses, mb_props* props, const mb_grid1d* grid,
                            int64_t cell_lo, int64_t cell_hi);

/* ---- merge_octree_N2_based!(rng, octree, pv, pia, cell, species, target_np[, grid]) merging_octree_N2.jl:1060-1094
 *      for every cell of [cell_lo, cell_hi] with n_local > threshold (threshold < 0: every cell), as the drivers do
 *      (simulations/1D/couette_varweight_octree.jl:93-98).  Leaves the species non-contiguous; call mb_squash_pia or sort. ---- */
int mb_merge_octree_N2(mb_ctx* ctx, const mb_octree_params* oc, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                       int64_t threshold, int64_t target_np, const mb_grid1d* grid, uint32_t timestep, uint32_t substream);

/* GridN2Merge(Nx, Ny, Nz, extent_multiplier) merging/merging_grid.jl:72-116 */
typedef struct {
    int32_t Nx, Ny, Nz;
    double extent_multiplier[3];
} mb_gridmerge_params;
/* ---- merge_grid_based!(rng, merging_grid, particles, pia, cell, species, species_data, phys_props | vx_extent, vy_extent, vz_extent
 *      [, grid::Grid1DUniform]) merging_grid.jl:597-703 for every cell of [cell_lo, cell_hi] with n_local > threshold (< 0: all).
 *      Exactly one of `props` (the velocity grid follows T and v of the cell as last computed by compute_props*, :190-197) and
 *      `extents6` (HOST: vx_lo, vx_hi, vy_lo, vy_hi, vz_lo, vz_hi, :210-221) is non-NULL.  grid != NULL: the 1-D variant, x of every
 *      np >= 2 output clamped into [min_x, max_x] (:528-549).  A velocity exactly on an upper grid bound (the reference would index
 *      past its grid) is reported as MB_ERR_PRECONDITION.  Leaves the species non-contiguous like the octree merge. ---- */
int mb_merge_grid_based(mb_ctx* ctx, const mb_gridmerge_params* mg, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                        double mass, mb_props* props, const double* extents6, int64_t threshold, const mb_grid1d* grid, uint32_t timestep,
                        uint32_t substream);

/* ---- initial conditions on the device: 1e8-1e9 particles are not sampled on the host and copied ----
 * sample_particles_equal_weight!(rng, particles, pia, cell, species, nparticles, m, T, Fnum, xlo, xhi, ylo, yhi, zlo, zhi;
 *     distribution, vx0, vy0, vz0)                                        distributions_and_sampling.jl:477-509   (grid == NULL, box6 given)
 * sample_particles_equal_weight!(rng, grid1duniform, particles, pia, species, species_data, ppc::Integer, T, Fnum[, cell_chunk])
 *                                                                          grids/grid_uniform1D.jl:117-152         (grid given, nparticles >= 0)
 * sample_particles_equal_weight!(rng, grid1duniform, ..., ndens::Float64, T, Fnum[, cell_chunk])   :154-219       (grid given, nparticles < 0)
 * for every cell of [cell_lo, cell_hi] in ascending order, appended at n_total + 1 exactly as the reference's loop does.
 * distribution: 0 Maxwellian (sample_maxwellian! :432-443), 1 BKW at t = 0 (sample_bkw! :195-213).  v0: 3 doubles or NULL.
 * With a slab grid the x bounds use the GLOBAL cell number (cell_offset + cell).  One Philox stream per cell (entity = cell). */
int mb_sample_particles_equal_weight(mb_ctx* ctx, const mb_grid1d* grid, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species,
                                     int64_t nparticles, double ndens, double mass, double T, double Fnum, const double* box6, int32_t distribution,
                                     const double* v0, uint32_t timestep, uint32_t substream);
/* sample_on_grid!(rng, vdf_func, particles, nv, m, T, n_total, xlo, xhi, ylo, yhi, zlo, zhi; v_mult, cutoff_mult, noise, v_offset)
 * distributions_and_sampling.jl:312-346 with evaluate_distribution_on_grid! :253-268 (vdf_kind 0: maxwellian :150-152, 1: bkw at
 * scaled_time 0 :168-177).  Every cell of [cell_lo, cell_hi] receives the weighted velocity-grid sample (an ensemble of 0-D cells),
 * appended at n_total + 1, and its indexer is set as ParticleIndexerArray(n_sampled) does (particles.jl:151).
 * n_sampled (nullable, host): particles per cell (the reference's return value). */
int mb_sample_on_grid(mb_ctx* ctx, int32_t vdf_kind, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species, int64_t nv, double mass,
                      double T, double n_total, const double* box6, double v_mult, double cutoff_mult, double noise, const double* v_offset,
                      uint32_t timestep, uint32_t substream, int64_t* n_sampled);

/* ---- slab exchange (replaces ChunkExchanger / exchange_particles! / sort_particles_after_exchange!, parallel.jl:21-581) ----
 * nccl_unique_id: 128 bytes from mb_comm_unique_id on rank 0, distributed by the host (torch.distributed / MPI / files). */
int mb_comm_unique_id(void* out128);
int mb_comm_init(mb_ctx* ctx, const void* nccl_unique_id128, int rank, int nranks);
/* Particles of `pv` whose x lies outside this rank's slab [x_lo, x_hi) are sent to the left / right neighbour and removed
 * locally; arrivals are appended after n_total.  Call between convect and sort (the sort then places arrivals: they are
 * just more keys).  x coordinates stay global; the slab grid (mb_grid1d_slab) carries the cell offset.
 * n_sent2/n_recv2 (nullable, host, 2 x int64: left, right) report the counts and synchronise. */
/* mode 0 (default): when the species is in the sorted layout (it was sorted and only moved by convection since) the exchange
 * only looks at the w cells next to each slab face (w = the sort's band half-width) and swaps fixed-size messages of up to 8192
 * particles per direction without any host synchronisation; a leaver from any other cell, or more than 8192 per direction, is
 * reported as an error by the next synchronising call.  Otherwise, and always with mode 1 or when the counts are requested,
 * every particle is examined and the message sizes are negotiated through the host.  All ranks must use the same mode. */
int mb_exchange_set_mode(mb_ctx* ctx, int32_t mode);
int mb_exchange_slab(mb_ctx* ctx, const mb_grid1d* slab, mb_pv* pv, mb_pia* pia, int64_t species, int64_t* n_sent2, int64_t* n_recv2);
/* exchange_particles!(exchanger, pv_chunks, pia_chunks, cell_chunks, species) parallel.jl:443-450 with the chunks of ONE process (the
 * reference's own use: logical chunks, test/test_couette_varweight_octree_chunking.jl:87-138): chunk i owns slab i of n_chunks
 * (slabs[i] = mb_grid1d_slab(global, i, n_chunks)) in its own context (same or different devices).  Same pack / unpack kernels and
 * the same two modes as mb_exchange_slab; the transport between neighbouring chunks is a device-to-device copy instead of
 * ncclSend/ncclRecv, so no communicator is needed.  Synchronises the chunks' streams. */
int mb_exchange_chunks(int32_t n_chunks, mb_ctx* const* ctxs, const mb_grid1d* slabs, mb_pv* const* pvs, mb_pia* const* pias, int64_t species);

#ifdef __cplusplus
}
#endif
#endif

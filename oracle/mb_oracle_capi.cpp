// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle for ctypes (oracle/oracle.py).
// See mb_oracle.hpp for the rules on who may use this.
#include <cstring>

#include "mb_oracle.hpp"
#include "mb_oracle_octree.hpp"
#include "mb_oracle_gridmerge.hpp"
#include "mb_oracle_parallel.hpp"

using namespace mbo;

extern "C" {

struct mbo_rng_spec {
    int32_t kind;        // 0: sequential xoshiro256++ handle in `seq`; 1: Philox streams keyed per entity
    void* seq;           // Xoshiro256pp*
    uint64_t seed;       // Philox key
    uint32_t timestep;   // Philox counter word 2
    uint32_t substream;  // Philox counter word 3, bits 8..31
};

// The device folds the species (pair) of an operator into the caller's 12-bit substream (merzbild.jl_b200/csrc/mb_common.cuh,
// stream_substream): species 1 leaves it unchanged.  Restated here so that multi-species runs replay the device's streams.
static inline uint32_t fold_substream(uint32_t substream, int64_t s1, int64_t s2) {
    return (substream & 0xFFFu) | ((uint32_t)((s1 - 1) & 0x3F) << 12) | ((uint32_t)((s2 - 1) & 0x3F) << 18);
}
// ---- philox (KAT hook) ----
void mbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox4x32::block(ctr, key, out); }
void mbo_philox_stream_doubles(uint64_t seed, uint32_t op, uint32_t substream, uint32_t timestep, uint32_t entity, int64_t n, double* out) {
    PhiloxStream s(seed, op, substream, timestep, entity);
    for (int64_t i = 0; i < n; i++) out[i] = s.rand();
}
void* mbo_rng_create(uint64_t seed) { return new Xoshiro256pp(seed); }
void* mbo_rng_create_stable(uint64_t seed) { return new Xoshiro256pp(Xoshiro256pp::stable_rng(seed)); }  // StableRNGs.jl StableRNG(seed)
void mbo_rng_free(void* r) { delete (Xoshiro256pp*)r; }
double mbo_rng_rand(void* r) { return ((Xoshiro256pp*)r)->rand(); }

// ---- ParticleVector ----
void* mbo_pv_create(int64_t np) { return new ParticleVector(np); }
void mbo_pv_free(void* p) { delete (ParticleVector*)p; }
int64_t mbo_pv_length(void* p) { return ((ParticleVector*)p)->length(); }
void mbo_pv_resize(void* p, int64_t n) { ((ParticleVector*)p)->resize(n); }
double* mbo_pv_particles(void* p) { return (double*)((ParticleVector*)p)->particles.data(); }
int64_t* mbo_pv_index(void* p) { return ((ParticleVector*)p)->index.data(); }
int64_t* mbo_pv_cell(void* p) { return ((ParticleVector*)p)->cell.data(); }
int64_t* mbo_pv_buffer(void* p) { return ((ParticleVector*)p)->buffer.data(); }
int64_t* mbo_pv_nbuffer(void* p) { return &((ParticleVector*)p)->nbuffer; }
void mbo_pv_add_particle(void* p, int64_t position, double w, const double* v, const double* x) { add_particle(*(ParticleVector*)p, position, w, v, x); }
void mbo_pv_update_buffer_new_particle(void* p, int64_t position) { update_particle_buffer_new_particle(*(ParticleVector*)p, position); }
// logical view: out[7*(i-lo)] = pv[i] for i in lo..hi (1-based inclusive)
void mbo_pv_get_logical(void* p, int64_t lo, int64_t hi, double* out) {
    ParticleVector& pv = *(ParticleVector*)p;
    for (int64_t i = lo; i <= hi; i++) std::memcpy(out + 7 * (i - lo), &pv[i], sizeof(Particle));
}
void mbo_pv_set_logical(void* p, int64_t lo, int64_t hi, const double* in) {
    ParticleVector& pv = *(ParticleVector*)p;
    for (int64_t i = lo; i <= hi; i++) std::memcpy(&pv[i], in + 7 * (i - lo), sizeof(Particle));
}

// ---- ParticleIndexerArray ----
void* mbo_pia_create(int64_t nc, int64_t ns) { return new ParticleIndexerArray(nc, ns); }
void mbo_pia_free(void* p) { delete (ParticleIndexerArray*)p; }
int64_t* mbo_pia_indexer(void* p) { return (int64_t*)((ParticleIndexerArray*)p)->indexer.data(); }
int64_t* mbo_pia_n_total(void* p) { return ((ParticleIndexerArray*)p)->n_total.data(); }
uint8_t* mbo_pia_contiguous(void* p) { return ((ParticleIndexerArray*)p)->contiguous.data(); }
int64_t mbo_map_cont_index(void* p, int64_t cell, int64_t species, int64_t i) { return map_cont_index(((ParticleIndexerArray*)p)->at(cell, species), i); }
void mbo_update_particle_indexer_new_lower_count(void* p, int64_t cell, int64_t species, int64_t n) { update_particle_indexer_new_lower_count(*(ParticleIndexerArray*)p, cell, species, n); }
void mbo_update_particle_indexer_new_particle(void* p, int64_t cell, int64_t species) { update_particle_indexer_new_particle(*(ParticleIndexerArray*)p, cell, species); }
void mbo_update_buffer_index_new_particle(void* pv, void* pia, int64_t cell, int64_t species) { update_buffer_index_new_particle(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species); }
void mbo_delete_particle(void* pv, void* pia, int64_t cell, int64_t species, int64_t i) { delete_particle(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species, i); }
void mbo_delete_particle_end(void* pv, void* pia, int64_t cell, int64_t species) { delete_particle_end(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species); }
void mbo_delete_particle_end_group1(void* pv, void* pia, int64_t cell, int64_t species) { delete_particle_end_group1(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species); }
void mbo_delete_particle_end_group2(void* pv, void* pia, int64_t cell, int64_t species) { delete_particle_end_group2(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species); }
void mbo_squash_pia(void* pv, void* pia, int64_t species) { squash_pia(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, species); }
void mbo_restore_particle_ordering(void* pv) {
    std::vector<int64_t> inv;
    restore_particle_ordering(*(ParticleVector*)pv, inv);
}
int mbo_check_pia_is_correct(void* pia, int64_t species, int64_t* where) { return check_pia_is_correct(*(ParticleIndexerArray*)pia, species, where); }
int mbo_check_unique_index(void* pv, void* pia, int64_t species, int64_t* code) { return check_unique_index(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, species, code); }

// ---- grid + sort ----
void mbo_grid_params(double L, int64_t nx, double wall_offset, double* out5) {
    Grid1DUniform g(L, nx, wall_offset);
    out5[0] = g.dx; out5[1] = g.inv_dx; out5[2] = g.min_x; out5[3] = g.max_x; out5[4] = g.L;
}
void mbo_sort_particles_grid(double L, int64_t nx, void* pv, void* pia, int64_t species) {
    Grid1DUniform g(L, nx);
    ParticleIndexerArray& P = *(ParticleIndexerArray*)pia;
    GridSortInPlace gs(nx, P.n_total[species - 1]);
    sort_particles(gs, g, *(ParticleVector*)pv, P, species);
}
void mbo_sort_particles_cells(void* pv, void* pia, int64_t species) {
    ParticleIndexerArray& P = *(ParticleIndexerArray*)pia;
    GridSortInPlace gs(P.n_cells, P.n_total[species - 1]);
    sort_particles(gs, *(ParticleVector*)pv, P, species);
}

// ---- interactions / collision factors ----
void mbo_make_interaction(double m_i, double m_k, double d, double o, double Tref, double* out8) {
    Interaction it = make_interaction(m_i, m_k, d, o, Tref);
    std::memcpy(out8, &it, sizeof(it));
}
double mbo_sigma_vhs(const double* it8, double g) { return sigma_vhs(*(const Interaction*)it8, g); }
double mbo_estimate_sigma_g_w_max(const double* it8, double m1, double m2, double T1, double T2, double Fnum, double mult) {
    return estimate_sigma_g_w_max(*(const Interaction*)it8, Species{m1, 0}, Species{m2, 0}, T1, T2, Fnum, mult);
}
void mbo_compute_com_g(const double* it8, const double* p1_7, const double* p2_7, double* vcom3, double* g) {
    CollisionData cd;
    compute_com(cd, *(const Interaction*)it8, *(const Particle*)p1_7, *(const Particle*)p2_7);
    compute_g(cd, *(const Particle*)p1_7, *(const Particle*)p2_7);
    for (int d = 0; d < 3; d++) vcom3[d] = cd.v_com[d];
    *g = cd.g;
}
// collide_2particles_vhs! (collision_ntc.jl:223-270) / collide_2particles_vhs_equal_weight! (:294-309) on particles i, k of one
// species in `cell`, as test/test_collision_vhs.jl calls them (compute_g! first); cf4 = {sigma_g_w_max, n_coll_performed,
// n_eq_w_coll_performed, unused}
void mbo_collide_2particles_vhs(const mbo_rng_spec* rs, const double* it8, void* pv_, void* pia_, int64_t i, int64_t k, int64_t cell, int64_t species,
                                double dw_tol, int equal_weight, double* cf4) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    const Interaction& it = *(const Interaction*)it8;
    CollisionData cd;
    CollisionFactors cf;
    cf.sigma_g_w_max = cf4[0];
    cf.n_coll_performed = (int64_t)cf4[1];
    cf.n_eq_w_coll_performed = (int64_t)cf4[2];
    compute_g(cd, pv[i], pv[k]);
    auto run = [&](auto& rng) {
        if (equal_weight) collide_2particles_vhs_equal_weight(rng, cd, cf, it, pv[i], pv[k]);
        else collide_2particles_vhs(rng, cd, cf, it, i, k, pv, pv, pia, cell, species, species, dw_tol);
    };
    if (rs->kind == 0) run(*(Xoshiro256pp*)rs->seq);
    else { PhiloxStream s(rs->seed, OP_USER, rs->substream, rs->timestep, 0); run(s); }
    cf4[0] = cf.sigma_g_w_max; cf4[1] = (double)cf.n_coll_performed; cf4[2] = (double)cf.n_eq_w_coll_performed;
}
int mbo_compute_octant(const double* v3, const double* mid3) { return compute_octant(v3, mid3); }
void mbo_octree_vel_middle(void* o, double* out3) { for (int d = 0; d < 3; d++) out3[d] = ((OctreeN2Merge*)o)->vel_middle[d]; }
void mbo_octree_compute_v_mean(void* o, int64_t bs, int64_t be, void* pv) { compute_v_mean(*(OctreeN2Merge*)o, bs, be, *(ParticleVector*)pv); }
void mbo_octree_bounds_recompute(void* o, int64_t bin_id, int64_t bs, int64_t be, void* pv) {
    bin_bounds_recompute(*(OctreeN2Merge*)o, bin_id, bs, be, *(ParticleVector*)pv);
}
void mbo_scatter_vhs(const mbo_rng_spec* rs, const double* it8, double* p1_7, double* p2_7) {
    CollisionData cd;
    Particle& a = *(Particle*)p1_7;
    Particle& b = *(Particle*)p2_7;
    compute_com(cd, *(const Interaction*)it8, a, b);
    compute_g(cd, a, b);
    if (rs->kind == 0) scatter_vhs(*(Xoshiro256pp*)rs->seq, cd, *(const Interaction*)it8, a, b);
    else { PhiloxStream s(rs->seed, OP_USER, rs->substream, rs->timestep, 0); scatter_vhs(s, cd, *(const Interaction*)it8, a, b); }
}

// cf arrays: per cell 6 entries laid out as CollisionFactors {n1,n2,sigma_g_w_max(double),n_coll,n_coll_performed,n_eq_w}
// passed as separate arrays (length n_cells, indexed cell-1): sgwm (double), counters int64[4*n_cells] = n1,n2.. no: explicit:
void mbo_ntc(const mbo_rng_spec* rs, double* sgwm, int64_t* n_coll, int64_t* n_perf, int64_t* n_eqw, const double* it8, void* pv_, void* pia_,
             int64_t cell_lo, int64_t cell_hi, int64_t species, double dt, double V, double dw_tol, int equal_weight) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    const Interaction& it = *(const Interaction*)it8;
    CollisionData cd;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        CollisionFactors cf;
        cf.sigma_g_w_max = sgwm[cell - 1];
        if (rs->kind == 0) ntc(*(Xoshiro256pp*)rs->seq, cf, cd, it, pv, pia, cell, species, dt, V, dw_tol, equal_weight != 0);
        else { PhiloxStream s(rs->seed, OP_NTC, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell); ntc(s, cf, cd, it, pv, pia, cell, species, dt, V, dw_tol, equal_weight != 0); }
        sgwm[cell - 1] = cf.sigma_g_w_max;
        if (n_coll) n_coll[cell - 1] = cf.n_coll;
        if (n_perf) n_perf[cell - 1] = cf.n_coll_performed;
        if (n_eqw) n_eqw[cell - 1] = cf.n_eq_w_coll_performed;
    }
}
void mbo_ntc2(const mbo_rng_spec* rs, double* sgwm, int64_t* n_coll, int64_t* n_perf, int64_t* n_eqw, const double* it8, void* pv1_, void* pv2_,
              void* pia_, int64_t cell_lo, int64_t cell_hi, int64_t s1, int64_t s2, double dt, double V, double dw_tol, int equal_weight) {
    ParticleVector& p1 = *(ParticleVector*)pv1_;
    ParticleVector& p2 = *(ParticleVector*)pv2_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    const Interaction& it = *(const Interaction*)it8;
    CollisionData cd;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        CollisionFactors cf;
        cf.sigma_g_w_max = sgwm[cell - 1];
        if (rs->kind == 0) ntc2(*(Xoshiro256pp*)rs->seq, cf, cd, it, p1, p2, pia, cell, s1, s2, dt, V, dw_tol, equal_weight != 0);
        else { PhiloxStream s(rs->seed, OP_NTC, fold_substream(rs->substream, s1, s2), rs->timestep, (uint32_t)cell); ntc2(s, cf, cd, it, p1, p2, pia, cell, s1, s2, dt, V, dw_tol, equal_weight != 0); }
        sgwm[cell - 1] = cf.sigma_g_w_max;
        if (n_coll) n_coll[cell - 1] = cf.n_coll;
        if (n_perf) n_perf[cell - 1] = cf.n_coll_performed;
        if (n_eqw) n_eqw[cell - 1] = cf.n_eq_w_coll_performed;
    }
}
void mbo_swpm(const mbo_rng_spec* rs, double* sgm, int64_t* n_coll, int64_t* n_perf, const double* it8, void* pv_, void* pia_, int64_t cell_lo,
              int64_t cell_hi, int64_t species, double G, double dt, double V) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    const Interaction& it = *(const Interaction*)it8;
    CollisionData cd;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        CollisionFactorsSWPM cf;
        cf.sigma_g_max = sgm[cell - 1];
        if (rs->kind == 0) swpm(*(Xoshiro256pp*)rs->seq, cf, cd, it, pv, pia, cell, species, G, dt, V);
        else { PhiloxStream s(rs->seed, OP_SWPM, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell); swpm(s, cf, cd, it, pv, pia, cell, species, G, dt, V); }
        sgm[cell - 1] = cf.sigma_g_max;
        if (n_coll) n_coll[cell - 1] = cf.n_coll;
        if (n_perf) n_perf[cell - 1] = cf.n_coll_performed;
    }
}
void mbo_scale_norm_rands(double* xr, double* yr, double* zr, int64_t n) {
    std::vector<double> a(xr, xr + n), b(yr, yr + n), c(zr, zr + n);
    scale_norm_rands(a, b, c, n);
    std::memcpy(xr, a.data(), n * 8); std::memcpy(yr, b.data(), n * 8); std::memcpy(zr, c.data(), n * 8);
}
void mbo_fp_linear(const mbo_rng_spec* rs, const double* it8, double mass, void* pv_, void* pia_, int64_t cell_lo, int64_t cell_hi,
                   int64_t species, double dt, double V) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    const Interaction& it = *(const Interaction*)it8;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        if (rs->kind == 0) {
            Xoshiro256pp& r = *(Xoshiro256pp*)rs->seq;
            auto src = [&](int64_t, double o[3]) {  // three sequential normals per particle like randn(rng) x3 (collision_fp.jl:164-170)
                if (r.engine == 1) { o[0] = r.randn(); o[1] = r.randn(); o[2] = r.randn(); return; }  // StableRNG: Julia's ziggurat
                for (int d = 0; d < 3; d += 2) {
                    const double u1 = std::max(1e-300, r.rand()), u2 = r.rand();
                    const double rr = std::sqrt(-2.0 * std::log(u1));
                    o[d] = rr * std::cos(twopi * u2);
                    if (d == 0) o[1] = rr * std::sin(twopi * u2);
                }
            };
            fp_linear(src, it, mass, pv, pia, cell, species, dt, V);
        } else {
            PhiloxStream base(rs->seed, OP_FP, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell);
            auto src = [&](int64_t j, double o[3]) { fp_normals_philox(base, j, o); };
            fp_linear(src, it, mass, pv, pia, cell, species, dt, V);
        }
    }
}

// ---- props ----
// out arrays sized by caller: np,n,T [n_cells*n_species]; v [3*n_cells*n_species]; moments [n_mom*n_cells*n_species]
void mbo_compute_props(void** pvs, void* pia_, const double* masses, int64_t n_moments, const int32_t* powers, double Tref, int with_moments,
                       double* lpa, double* np, double* n, double* v, double* T, double* moments) {
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    std::vector<ParticleVector*> P(pia.n_species);
    std::vector<Species> sd(pia.n_species);
    for (int64_t s = 0; s < pia.n_species; s++) { P[s] = (ParticleVector*)pvs[s]; sd[s] = Species{masses[s], 0}; }
    PhysProps pp(pia.n_cells, pia.n_species, std::vector<int>(powers, powers + n_moments), false, Tref);
    compute_props(P, pia, sd, pp, with_moments != 0);
    const int64_t N = pia.n_cells * pia.n_species;
    std::memcpy(lpa, pp.lpa.data(), pia.n_species * 8);
    std::memcpy(np, pp.np.data(), N * 8); std::memcpy(n, pp.n.data(), N * 8); std::memcpy(T, pp.T.data(), N * 8);
    std::memcpy(v, pp.v.data(), 3 * N * 8);
    if (with_moments && n_moments > 0) std::memcpy(moments, pp.moments.data(), n_moments * N * 8);
}
void mbo_compute_props_sorted(void** pvs, void* pia_, const double* masses, int64_t cell_lo, int64_t cell_hi, int ndens, double L, int64_t nx,
                              double* np, double* n, double* v, double* T) {
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    std::vector<ParticleVector*> P(pia.n_species);
    std::vector<Species> sd(pia.n_species);
    for (int64_t s = 0; s < pia.n_species; s++) { P[s] = (ParticleVector*)pvs[s]; sd[s] = Species{masses[s], 0}; }
    PhysProps pp(pia.n_cells, pia.n_species, {}, ndens != 0);
    const int64_t N = pia.n_cells * pia.n_species;
    std::memcpy(pp.np.data(), np, N * 8); std::memcpy(pp.n.data(), n, N * 8); std::memcpy(pp.T.data(), T, N * 8);
    std::memcpy(pp.v.data(), v, 3 * N * 8);
    Grid1DUniform g(ndens ? L : 1.0, ndens ? nx : 1);
    compute_props_sorted(P, pia, sd, pp, cell_lo, cell_hi, ndens ? &g : nullptr);
    std::memcpy(np, pp.np.data(), N * 8); std::memcpy(n, pp.n.data(), N * 8); std::memcpy(T, pp.T.data(), N * 8);
    std::memcpy(v, pp.v.data(), 3 * N * 8);
}
// avg_props!(phys_props_avg, phys_props, n_avg_timesteps) physical_props.jl:281-299 on flat arrays (N = n_cells * n_species entries of
// np, n, T; 3 N of v; n_species of lpa), in place on the avg arrays
void mbo_avg_props(int64_t n_cells, int64_t n_species, double* a_lpa, double* a_np, double* a_n, double* a_v, double* a_T, const double* lpa,
                   const double* np, const double* n, const double* v, const double* T, double n_avg_timesteps) {
    PhysProps avg(n_cells, n_species, {}, false), pp(n_cells, n_species, {}, false);
    const int64_t N = n_cells * n_species;
    std::memcpy(avg.lpa.data(), a_lpa, n_species * 8); std::memcpy(pp.lpa.data(), lpa, n_species * 8);
    std::memcpy(avg.np.data(), a_np, N * 8); std::memcpy(avg.n.data(), a_n, N * 8); std::memcpy(avg.T.data(), a_T, N * 8);
    std::memcpy(avg.v.data(), a_v, 3 * N * 8);
    std::memcpy(pp.np.data(), np, N * 8); std::memcpy(pp.n.data(), n, N * 8); std::memcpy(pp.T.data(), T, N * 8);
    std::memcpy(pp.v.data(), v, 3 * N * 8);
    avg_props(avg, pp, n_avg_timesteps);
    std::memcpy(a_lpa, avg.lpa.data(), n_species * 8);
    std::memcpy(a_np, avg.np.data(), N * 8); std::memcpy(a_n, avg.n.data(), N * 8); std::memcpy(a_T, avg.T.data(), N * 8);
    std::memcpy(a_v, avg.v.data(), 3 * N * 8);
}
double mbo_compute_mixed_moment(void* pv, void* pia, int64_t cell, int64_t species, const int32_t* powers, double sum_scaler, double res_scaler) {
    const int p[3] = {powers[0], powers[1], powers[2]};
    return compute_mixed_moment(*(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species, p, sum_scaler, res_scaler);
}

// ---- convection ----
// walls6 = {T_l, T_r, vy_l, vy_r, acc_l, acc_r}; surf22 (nullable) = per wall e (0,1): np, flux_inc, flux_refl, force[3], normal_p, shear[3], ke  (11 each)
void mbo_convect_particles(const mbo_rng_spec* rs, double L, int64_t nx, const double* walls6, void* pv_, void* pia_, int64_t species,
                           const double* masses, int64_t n_species, double* surf22, double dt, int compute_cell) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    Grid1DUniform g(L, nx);
    std::vector<Species> sd(n_species);
    for (int64_t s = 0; s < n_species; s++) sd[s] = Species{masses[s], 0};
    MaxwellWalls1D b(sd, walls6[0], walls6[1], walls6[2], walls6[3], walls6[4], walls6[5]);
    SurfProps sp(n_species);
    SurfProps* spp = surf22 ? &sp : nullptr;
    if (rs->kind == 0) {
        Xoshiro256pp& r = *(Xoshiro256pp*)rs->seq;
        convect_particles([&](int64_t) -> Xoshiro256pp& { return r; }, g, b, pv, pia, species, sd[species - 1].mass, spp, dt, compute_cell != 0);
    } else {
        PhiloxStream s;
        convect_particles([&](int64_t i) -> PhiloxStream& { s.reset(rs->seed, OP_CONVECT, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)(i - 1)); return s; }, g, b,
                          pv, pia, species, sd[species - 1].mass, spp, dt, compute_cell != 0);
    }
    if (surf22) {
        for (int e = 0; e < 2; e++) {
            const int64_t o = e + 2 * (species - 1);
            double* q = surf22 + 11 * e;
            q[0] = sp.np[o]; q[1] = sp.flux_incident[o]; q[2] = sp.flux_reflected[o];
            for (int d = 0; d < 3; d++) { q[3 + d] = sp.force[d + 3 * o]; q[7 + d] = sp.shear_pressure[d + 3 * o]; }
            q[6] = sp.normal_pressure[o]; q[10] = sp.kinetic_energy_flux[o];
        }
    }
}

// ---- sampling ----
void mbo_sample_equal_weight_grid(void* rng, double L, int64_t nx, void* pv, void* pia, int64_t species, double mass, double ndens, double T, double Fnum,
                                  int64_t cell_lo, int64_t cell_hi) {
    Grid1DUniform g(L, nx);
    sample_particles_equal_weight_grid(*(Xoshiro256pp*)rng, g, *(ParticleVector*)pv, *(ParticleIndexerArray*)pia, species, mass, ndens, T, Fnum, cell_lo, cell_hi);
}
void mbo_sample_equal_weight_cell(void* rng, void* pv, void* pia, int64_t cell, int64_t species, int64_t nparticles, double m, double T, double Fnum,
                                  const double* box6, int distribution, const double* v0) {
    sample_particles_equal_weight(*(Xoshiro256pp*)rng, *(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species, nparticles, m, T, Fnum, box6[0],
                                  box6[1], box6[2], box6[3], box6[4], box6[5], distribution, v0);
}
int64_t mbo_sample_on_grid(void* rng, int vdf_kind, void* pv, int64_t nv, double m, double T, double n_total, const double* box6, double v_mult,
                           double cutoff_mult, double noise, const double* v_offset) {
    return sample_on_grid(*(Xoshiro256pp*)rng, vdf_kind, *(ParticleVector*)pv, nv, m, T, n_total, box6[0], box6[1], box6[2], box6[3], box6[4], box6[5],
                          v_mult, cutoff_mult, noise, v_offset);
}

// per-cell variants: with Philox one stream per cell (entity = cell), the convention of merzbild.jl_b200/csrc/mb_sample.cu;
// grid2 = (L, nx) or NULL (then box6); nparticles < 0: the number-density variant of grid_uniform1D.jl:198-219
void mbo_sample_equal_weight_cells(const mbo_rng_spec* rs, const double* grid2, void* pv_, void* pia_, int64_t cell_lo, int64_t cell_hi, int64_t species,
                                   int64_t nparticles, double ndens, double m, double T, double Fnum, const double* box6, int distribution,
                                   const double* v0) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        auto one = [&](auto& rng) {
            double b[6] = {0, 0, 0.0, 1.0, 0.0, 1.0};
            if (grid2) {
                Grid1DUniform g(grid2[0], (int64_t)grid2[1]);
                b[0] = g.cell_xlo(cell); b[1] = g.cell_xhi(cell);
                if (nparticles < 0) { sample_particles_equal_weight_grid(rng, g, pv, pia, species, m, ndens, T, Fnum, cell, cell); return; }
            } else {
                for (int d = 0; d < 6; d++) b[d] = box6[d];
            }
            sample_particles_equal_weight(rng, pv, pia, cell, species, nparticles, m, T, Fnum, b[0], b[1], b[2], b[3], b[4], b[5], distribution, v0, true);
        };
        if (rs->kind == 0) one(*(Xoshiro256pp*)rs->seq);
        else { PhiloxStream s(rs->seed, OP_SAMPLE, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell); one(s); }
    }
}
// ensemble of 0-D cells, each the sample_on_grid! population appended at n_total + 1 with the cell's indexer set as
// ParticleIndexerArray(n_sampled) does (particles.jl:151)
int64_t mbo_sample_on_grid_cells(const mbo_rng_spec* rs, int vdf_kind, void* pv_, void* pia_, int64_t cell_lo, int64_t cell_hi, int64_t species, int64_t nv,
                                 double m, double T, double n_total, const double* box6, double v_mult, double cutoff_mult, double noise,
                                 const double* v_offset) {
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    int64_t n = 0;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        const int64_t off = pia.n_total[species - 1];
        auto one = [&](auto& rng) {
            n = sample_on_grid(rng, vdf_kind, pv, nv, m, T, n_total, box6[0], box6[1], box6[2], box6[3], box6[4], box6[5], v_mult, cutoff_mult, noise,
                               v_offset, off);
        };
        if (rs->kind == 0) one(*(Xoshiro256pp*)rs->seq);
        else { PhiloxStream s(rs->seed, OP_SAMPLE, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell); one(s); }
        ParticleIndexer& ix = pia.at(cell, species);
        ix.n_local = n; ix.start1 = off + 1; ix.end1 = off + n; ix.n_group1 = n; ix.start2 = 0; ix.end2 = -1; ix.n_group2 = 0;
        pia.n_total[species - 1] += n;
        for (int64_t i = 1; i <= n; i++) pv.cell[off + i - 1] = cell;
    }
    return n;
}

// ---- surface properties (KAT hook, test/test_surface_props_1D_uniform.jl) ----
// ops: sequence of n_ops (kind, element, row) with kind 0 = update_surface_incident!, 1 = update_surface_reflected! applied to the
// particle rows7[row]; then optionally surface_props_scale! with (mass, dt, inv_areas2).  out22 as mbo_convect_particles returns it.
void mbo_surface_props_kat(const double* rows7, const int64_t* ops3, int64_t n_ops, int do_scale, double mass, double dt, const double* inv_areas2,
                           double* out22) {
    SurfProps sp(1);
    for (int64_t o = 0; o < n_ops; o++) {
        const Particle& p = *(const Particle*)(rows7 + 7 * ops3[3 * o + 2]);
        update_surface(p, 1, sp, ops3[3 * o + 1], ops3[3 * o] == 0);
    }
    if (do_scale) {
        sp.inv_areas[0] = inv_areas2[0]; sp.inv_areas[1] = inv_areas2[1];
        surface_props_scale(1, sp, mass, dt);
    }
    for (int e = 0; e < 2; e++) {
        double* q = out22 + 11 * e;
        q[0] = sp.np[e]; q[1] = sp.flux_incident[e]; q[2] = sp.flux_reflected[e];
        for (int d = 0; d < 3; d++) { q[3 + d] = sp.force[d + 3 * e]; q[7 + d] = sp.shear_pressure[d + 3 * e]; }
        q[6] = sp.normal_pressure[e]; q[10] = sp.kinetic_energy_flux[e];
    }
}

// ---- velocity-grid merging ----
void* mbo_gridmerge_create(int64_t nx, int64_t ny, int64_t nz, const double* mult3) { return new GridN2Merge(nx, ny, nz, mult3); }
void mbo_gridmerge_free(void* g) { delete (GridN2Merge*)g; }
int64_t mbo_gridmerge_index(void* g, const double* ext6, const double* v) {
    GridN2Merge& mg = *(GridN2Merge*)g;
    compute_velocity_extent_given(mg, ext6);
    return compute_grid_index(mg, v);
}
// cell i (1-based): out = {np, w, v_mean[3], v_std_sq[3], x_mean[3], x_std_sq[3]} (14 doubles)
void mbo_gridmerge_cell(void* g, int64_t i, double* out) {
    const GridCell& c = ((GridN2Merge*)g)->cells[i - 1];
    out[0] = (double)c.np; out[1] = c.w;
    for (int d = 0; d < 3; d++) { out[2 + d] = c.v_mean[d]; out[5 + d] = c.v_std_sq[d]; out[8 + d] = c.x_mean[d]; out[11 + d] = c.x_std_sq[d]; }
}
// merge_grid_based! for the cells of [cell_lo, cell_hi] with n_local > threshold (threshold < 0: all).  Tv: per cell (T, vx, vy, vz) of the
// species as PhysProps holds them (4 doubles per cell of the range), or NULL with ext6 = explicit extents.  Signs: Philox block (grid index - 1)
// of the (OP_MERGE_GRID, timestep, cell) stream, or sequential draws.  Returns the number of cells whose particles fell off the grid index range.
int64_t mbo_merge_grid_based(const mbo_rng_spec* rs, void* g, void* pv_, void* pia_, int64_t cell_lo, int64_t cell_hi, int64_t species, double mass,
                             int64_t threshold, const double* Tv, const double* ext6, double L, int64_t nx) {
    GridN2Merge& mg = *(GridN2Merge*)g;
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    Grid1DUniform grid(L > 0 ? L : 1.0, nx > 0 ? nx : 1);
    const Grid1DUniform* gp = nx > 0 ? &grid : nullptr;
    int64_t bad = 0;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        if (!(threshold < 0 || pia.at(cell, species).n_local > threshold)) continue;
        if (pia.at(cell, species).n_local <= 0) continue;
        const double* tv = Tv ? Tv + 4 * (cell - cell_lo) : nullptr;
        const double zero[3] = {0, 0, 0};
        bool ok;
        if (rs->kind == 0) {
            SeqSigns<Xoshiro256pp> s{*(Xoshiro256pp*)rs->seq};
            ok = merge_grid_based(s, mg, pv, pia, cell, species, mass, tv ? tv[0] : 0.0, tv ? tv + 1 : zero, ext6, gp);
        } else {
            PhiloxSigns s{PhiloxStream(rs->seed, OP_MERGE_GRID, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell)};
            ok = merge_grid_based(s, mg, pv, pia, cell, species, mass, tv ? tv[0] : 0.0, tv ? tv + 1 : zero, ext6, gp);
        }
        if (!ok) bad++;
    }
    return bad;
}

// ---- octree ----
void* mbo_octree_create(int split, int init_bounds, int bounds_compute, int64_t max_Nbins, int64_t max_depth) {
    return new OctreeN2Merge((OctreeBinSplit)split, (OctreeInitBin)init_bounds, (OctreeBinBounds)bounds_compute, max_Nbins, max_depth);
}
void mbo_octree_free(void* o) { delete (OctreeN2Merge*)o; }
int64_t mbo_octree_nbins(void* o) { return ((OctreeN2Merge*)o)->Nbins; }
int64_t mbo_octree_n_particles(void* o) { return ((OctreeN2Merge*)o)->n_particles; }
int64_t mbo_octree_total_post_merge_np(void* o) { return ((OctreeN2Merge*)o)->total_post_merge_np; }
// bin i (1-based): out = {np, w, vmin[3], vmax[3], depth, can_be_refined, bin_start, bin_end} (12 doubles)
void mbo_octree_bin(void* o_, int64_t i, double* out) {
    OctreeN2Merge& o = *(OctreeN2Merge*)o_;
    const OctreeCell& b = o.bins[i - 1];
    out[0] = (double)b.np; out[1] = b.w;
    for (int d = 0; d < 3; d++) { out[2 + d] = b.v_min[d]; out[5 + d] = b.v_max[d]; }
    out[8] = (double)b.depth; out[9] = b.can_be_refined ? 1.0 : 0.0; out[10] = (double)o.bin_start[i - 1]; out[11] = (double)o.bin_end[i - 1];
}
// full bin i: out = {v_mean[3], v_std_sq[3], x_mean[3], x_std_sq[3]}
void mbo_octree_full_bin(void* o_, int64_t i, double* out) {
    const OctreeFullCell& f = ((OctreeN2Merge*)o_)->full_bins[i - 1];
    for (int d = 0; d < 3; d++) { out[d] = f.v_mean[d]; out[3 + d] = f.v_std_sq[d]; out[6 + d] = f.x_mean[d]; out[9 + d] = f.x_std_sq[d]; }
}
int64_t* mbo_octree_particle_indexes_sorted(void* o) { return ((OctreeN2Merge*)o)->particle_indexes_sorted.data(); }
void mbo_octree_init(void* o, void* pv, void* pia, int64_t cell, int64_t species) {
    OctreeN2Merge& oc = *(OctreeN2Merge*)o;
    oc.Nbins = 0;
    resize_octree_buffers(oc, ((ParticleIndexerArray*)pia)->at(cell, species).n_local);
    init_octree(oc, *(ParticleVector*)pv, *(ParticleIndexerArray*)pia, cell, species);
}
void mbo_octree_split_bin(void* o, int64_t bin_id, void* pv) { split_bin(*(OctreeN2Merge*)o, bin_id, *(ParticleVector*)pv); }
void mbo_octree_compute_bin_props(void* o, int64_t bin_id, void* pv) { compute_bin_props(*(OctreeN2Merge*)o, bin_id, *(ParticleVector*)pv); }
void mbo_octree_compute(void* o, void* pv, int64_t target_np) { compute_octree(*(OctreeN2Merge*)o, *(ParticleVector*)pv, target_np); }
// merge cells [cell_lo, cell_hi] whose n_local > threshold (threshold < 0: merge unconditionally); grid: L <= 0 -> 0-D variant
void mbo_merge_octree_N2(const mbo_rng_spec* rs, void* o, void* pv_, void* pia_, int64_t cell_lo, int64_t cell_hi, int64_t species, int64_t threshold,
                         int64_t target_np, double L, int64_t nx, int squash_after_each) {
    OctreeN2Merge& oc = *(OctreeN2Merge*)o;
    ParticleVector& pv = *(ParticleVector*)pv_;
    ParticleIndexerArray& pia = *(ParticleIndexerArray*)pia_;
    Grid1DUniform g(L > 0 ? L : 1.0, L > 0 ? nx : 1);
    const Grid1DUniform* gp = L > 0 ? &g : nullptr;
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        if (threshold >= 0 && !(pia.at(cell, species).n_local > threshold)) continue;
        if (rs->kind == 0) { SeqSigns<Xoshiro256pp> s{*(Xoshiro256pp*)rs->seq}; merge_octree_N2_based(s, oc, pv, pia, cell, species, target_np, gp); }
        else { PhiloxSigns s{PhiloxStream(rs->seed, OP_MERGE, fold_substream(rs->substream, species, species), rs->timestep, (uint32_t)cell)}; merge_octree_N2_based(s, oc, pv, pia, cell, species, target_np, gp); }
        if (squash_after_each) squash_pia(pv, pia, species);
    }
}

// ---- parallel.jl: chunk exchange ----
void* mbo_exchanger_create(int64_t n_chunks, int64_t n_cells) { return new ChunkExchanger(n_chunks, n_cells); }
void mbo_exchanger_free(void* e) { delete (ChunkExchanger*)e; }
int64_t* mbo_exchanger_indexer(void* e) { return (int64_t*)((ChunkExchanger*)e)->indexer.data(); }
void mbo_exchanger_reset(void* e, int64_t chunk_id) { reset_exchanger(*(ChunkExchanger*)e, chunk_id); }
// chunks: cell ranges [chunk_lo[c], chunk_hi[c]] (1-based inclusive), pvs/pias arrays of handles (one species -> one pv per chunk)
void mbo_exchange_particles(void* e, void** pvs, void** pias, const int64_t* chunk_lo, const int64_t* chunk_hi, int64_t n_chunks, int64_t species,
                            int64_t i, int64_t j) {
    std::vector<ParticleVector*> P(n_chunks);
    std::vector<ParticleIndexerArray*> A(n_chunks);
    std::vector<CellChunk> C(n_chunks);
    for (int64_t c = 0; c < n_chunks; c++) { P[c] = (ParticleVector*)pvs[c]; A[c] = (ParticleIndexerArray*)pias[c]; C[c] = CellChunk{chunk_lo[c], chunk_hi[c]}; }
    if (i > 0) exchange_particles(*(ChunkExchanger*)e, P, A, C, species, i, j);
    else exchange_particles_all(*(ChunkExchanger*)e, P, A, C, species);
}
void mbo_sort_particles_after_exchange(void* e, void* pv, void* pia, int64_t cell_lo, int64_t cell_hi, int64_t species) {
    ParticleIndexerArray& P = *(ParticleIndexerArray*)pia;
    GridSortInPlace gs(P.n_cells, P.n_total[species - 1]);
    sort_particles_after_exchange(*(ChunkExchanger*)e, gs, *(ParticleVector*)pv, P, CellChunk{cell_lo, cell_hi}, species);
}
// returns number of rounds; pairs written as (i,j) 1-based into out[2*k], round id into round_of[k]
int64_t mbo_generate_1_factorization(int64_t n_chunks, int64_t* out_pairs, int64_t* round_of) {
    auto lol = generate_1_factorization(n_chunks);
    int64_t k = 0;
    for (size_t r = 0; r < lol.size(); r++)
        for (auto& pr : lol[r]) { out_pairs[2 * k] = pr.first; out_pairs[2 * k + 1] = pr.second; round_of[k] = (int64_t)r; k++; }
    return (int64_t)lol.size();
}

}  // extern "C"

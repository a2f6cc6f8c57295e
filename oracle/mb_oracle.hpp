// TEST INFRASTRUCTURE ONLY -- CPU oracle: a restatement in C++17 of the reference's per-timestep DSMC
// particle pipeline (merzbild/Merzbild.jl v0.7.10, 100 % Julia, not runnable in this image: no Julia).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
// import, link or execute anything under oracle/.  The product path (libmerzbild_b200.so) never does.
//
// Parity pin: (1) the deterministic functions are pinned by the known-answer vectors the reference's own tests hold
// (tests/test_oracle_kat_*.py restate test/test_grid_sorting.jl, test_convection_1D.jl, test_computes.jl, test_octree_*.jl,
// test_collision_utils.jl, test_collision_fp.jl, test_pia_contiguous.jl, test_particle_index_sorting.jl, test_indexing*.jl ...);
// (2) the stochastic pipeline is pinned BIT-LEVEL by the reference's own seeded golden runs (test/data/*.nc, extracted into
// tests/golden/reference_histories.json): with the StableRNGs.jl generator, Julia's randn and exp restated (philox.hpp, mb_jlexp.h)
// the oracle reproduces every golden run of this path -- 0-D two-species and BKW histories (equal weight, octree / grid merging,
// SWPM), 1-D Couette (NTC, SWPM, Fokker-Planck, octree merging, surface properties, index re-sorting) -- to round-off at every
// recorded step (tests/test_oracle_reference_bitlevel.py); (3) generator-independent pins on top (tests/test_oracle_reference_runs.py,
// tests/test_oracle_stat.py: golden histories as draws of the oracle ensemble, BKW analytic moments test/test_bkw.jl:25-29, T_eq
// test/test_2species.jl:25, SPARTA Couette profile test/data/external/).
//
// Every function cites the reference file:line (relative to /root/reference/src) it follows.
// Indices stored in the containers are 1-based and inclusive exactly as in the reference, so a dump of
// `index`, `buffer`, or a ParticleIndexer is directly comparable to the Julia structs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <type_traits>
#include <vector>

#include "philox.hpp"
#include "../merzbild.jl_b200/csrc/mb_jlexp.h"    // exp() as Julia computes it: grid-sampled weights identical to the reference's, shared with the device library
#include "../merzbild.jl_b200/csrc/mb_normals.h"  // fp32 Box-Muller shared with the CUDA kernel (bit-identical normals)

namespace mbo {

constexpr double k_B = 1.380649e-23;         // constants.jl:4
constexpr double twopi = 2.0 * M_PI;         // constants.jl:29
constexpr double c_light = 299792458.0;      // constants.jl:9
constexpr int64_t DELTA_PARTICLES = 256;     // Merzbild.jl:24
constexpr double EPS = std::numeric_limits<double>::epsilon();  // Julia eps()

// ---------------------------------------------------------------------------------------------
// particles.jl
// ---------------------------------------------------------------------------------------------
struct Particle {  // particles.jl:14-18
    double w;
    double v[3];
    double x[3];
};

struct ParticleIndexer {  // particles.jl:56-66; empty = (0,0,-1,0,0,-1,0) particles.jl:84
    int64_t n_local = 0, start1 = 0, end1 = -1, n_group1 = 0, start2 = 0, end2 = -1, n_group2 = 0;
};

struct ParticleIndexerArray {  // particles.jl:104-141
    int64_t n_cells, n_species;
    std::vector<ParticleIndexer> indexer;  // [cell + n_cells * species], 0-based storage
    std::vector<int64_t> n_total;
    std::vector<uint8_t> contiguous;
    ParticleIndexerArray(int64_t nc, int64_t ns)
        : n_cells(nc), n_species(ns), indexer(nc * ns), n_total(ns, 0), contiguous(ns, 1) {}
    ParticleIndexer& at(int64_t cell, int64_t species) { return indexer[(cell - 1) + n_cells * (species - 1)]; }
    const ParticleIndexer& at(int64_t cell, int64_t species) const {
        return indexer[(cell - 1) + n_cells * (species - 1)];
    }
};

struct ParticleVector {  // particles.jl:194-212
    std::vector<Particle> particles;
    std::vector<int64_t> index, cell, buffer;
    int64_t nbuffer;
    explicit ParticleVector(int64_t np) : particles(np), index(np), cell(np, 0), buffer(np), nbuffer(np) {
        for (int64_t i = 0; i < np; i++) {
            particles[i] = Particle{0.0, {0, 0, 0}, {0, 0, 0}};
            index[i] = i + 1;
            buffer[i] = np - i;  // np:-1:1
        }
    }
    int64_t length() const { return (int64_t)particles.size(); }
    Particle& operator[](int64_t i) { return particles[index[i - 1] - 1]; }  // particles.jl:225
    const Particle& operator[](int64_t i) const { return particles[index[i - 1] - 1]; }

    void resize(int64_t n) {  // particles.jl:269-298
        const int64_t old_len = length();
        particles.resize(n, Particle{0.0, {0, 0, 0}, {0, 0, 0}});
        index.resize(n);
        cell.resize(n, 0);
        std::vector<int64_t> oldbuf(buffer.begin(), buffer.begin() + old_len);
        buffer.resize(n);
        const int64_t n_diff = n - old_len;
        for (int64_t i = old_len; i < n; i++) index[i] = i + 1;
        for (int64_t i = 0; i < old_len; i++) buffer[n_diff + i] = oldbuf[i];
        for (int64_t i = 1; i <= n_diff; i++) buffer[i - 1] = n + 1 - i;
        nbuffer += n_diff;
    }
};

// particles.jl:311-315
inline void update_particle_buffer_new_particle(ParticleVector& pv, int64_t position) {
    pv.index[position - 1] = pv.buffer[pv.nbuffer - 1];
    pv.nbuffer -= 1;
}
// particles.jl:364-366
inline int64_t map_cont_index(const ParticleIndexer& pi, int64_t i) {
    return i < pi.n_group1 ? i + pi.start1 : (i - pi.n_group1) + pi.start2;
}
// particles.jl:395-413
inline void update_particle_indexer_new_lower_count(ParticleIndexerArray& pia, int64_t cell, int64_t species,
                                                    int64_t new_lower_count) {
    ParticleIndexer& ix = pia.at(cell, species);
    int64_t diff = ix.n_local - new_lower_count;
    ix.n_local = new_lower_count;
    pia.n_total[species - 1] -= diff;
    if (new_lower_count > ix.n_group1) {
        ix.end2 -= diff;
        ix.n_group2 -= diff;
    } else {
        diff -= ix.n_group2;
        ix.start2 = 0; ix.end2 = -1; ix.n_group2 = 0;
        ix.end1 -= diff;
        ix.n_group1 -= diff;
    }
}
// particles.jl:426-433
inline void update_particle_indexer_new_particle(ParticleIndexerArray& pia, int64_t cell, int64_t species) {
    ParticleIndexer& ix = pia.at(cell, species);
    pia.n_total[species - 1] += 1;
    ix.n_local += 1;
    ix.n_group2 += 1;
    ix.start2 = ix.start2 > 0 ? ix.start2 : pia.n_total[species - 1];
    ix.end2 = pia.n_total[species - 1];
}
// particles.jl:718-721
inline void update_buffer_index_new_particle(ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell,
                                             int64_t species) {
    update_particle_indexer_new_particle(pia, cell, species);
    update_particle_buffer_new_particle(pv, pia.n_total[species - 1]);
}
// particles.jl:739-742
inline void add_particle(ParticleVector& pv, int64_t position, double w, const double v[3], const double x[3]) {
    update_particle_buffer_new_particle(pv, position);
    pv[position] = Particle{w, {v[0], v[1], v[2]}, {x[0], x[1], x[2]}};
}
// particles.jl:501-525
inline void delete_particle_end_group1(ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species) {
    ParticleIndexer& ix = pia.at(cell, species);
    if (ix.n_group1 == 0) return;
    const int64_t index_of_deleted = ix.end1;
    pv[index_of_deleted].w = 0.0;
    ix.n_local -= 1; ix.end1 -= 1; ix.n_group1 -= 1;
    pia.n_total[species - 1] -= 1;
    if (ix.end1 < ix.start1) { ix.start1 = 0; ix.end1 = -1; }
    pv.nbuffer += 1;
    pv.buffer[pv.nbuffer - 1] = pv.index[index_of_deleted - 1];
}
// particles.jl:542-566
inline void delete_particle_end_group2(ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species) {
    ParticleIndexer& ix = pia.at(cell, species);
    if (ix.n_group2 == 0) return;
    const int64_t index_of_deleted = ix.end2;
    pv[index_of_deleted].w = 0.0;
    ix.n_local -= 1; ix.end2 -= 1; ix.n_group2 -= 1;
    pia.n_total[species - 1] -= 1;
    if (ix.end2 < ix.start2) { ix.start2 = 0; ix.end2 = -1; }
    pv.nbuffer += 1;
    pv.buffer[pv.nbuffer - 1] = pv.index[index_of_deleted - 1];
}
// particles.jl:478-484
inline void delete_particle_end(ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species) {
    if (pia.at(cell, species).n_group2 > 0) delete_particle_end_group2(pv, pia, cell, species);
    else delete_particle_end_group1(pv, pia, cell, species);
}
// particles.jl:448-461
inline void delete_particle(ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species, int64_t i) {
    ParticleIndexer& ix = pia.at(cell, species);
    if (ix.n_group2 > 0 && i >= ix.start2 && i <= ix.end2) {
        std::swap(pv.index[ix.end2 - 1], pv.index[i - 1]);
        delete_particle_end_group2(pv, pia, cell, species);
    } else {
        std::swap(pv.index[ix.end1 - 1], pv.index[i - 1]);
        delete_particle_end_group1(pv, pia, cell, species);
    }
}

// particles.jl:622-682
inline void squash_pia(ParticleVector& pv, ParticleIndexerArray& pia, int64_t species) {
    if (pia.contiguous[species - 1]) return;
    const int64_t n_cells = pia.n_cells;
    auto shift = [&](int64_t s, int64_t e, int64_t offset) {
        for (int64_t j = s; j <= e; j++) {
            pv.index[j - 1] = pv.index[j + offset - 1];
            pv.cell[j - 1] = pv.cell[j + offset - 1];
        }
    };
    if (n_cells == 1) {
        ParticleIndexer& ix = pia.at(1, species);
        if (ix.n_group2 > 0) {
            const int64_t e1 = ix.end1 > 0 ? ix.end1 : 0;
            const int64_t offset = ix.start2 - (e1 + 1);
            if (offset > 0) {
                ix.start2 -= offset; ix.end2 -= offset;
                shift(ix.start2, ix.end2, offset);
            }
        }
    } else {
        int64_t last_end = pia.at(1, species).end1 > 0 ? pia.at(1, species).end1 : 0;
        for (int64_t i = 1; i <= n_cells - 1; i++) {
            ParticleIndexer& nx = pia.at(i + 1, species);
            const int64_t offset = nx.start1 - (last_end + 1);
            if (offset > 0) {
                nx.start1 -= offset; nx.end1 -= offset;
                shift(nx.start1, nx.end1, offset);
            }
            last_end = nx.end1 > 0 ? nx.end1 : last_end;
        }
        for (int64_t i = 1; i <= n_cells; i++) {
            ParticleIndexer& ix = pia.at(i, species);
            if (ix.n_group2 > 0) {
                const int64_t offset = ix.start2 - (last_end + 1);
                if (offset > 0) {
                    ix.start2 -= offset; ix.end2 -= offset;
                    shift(ix.start2, ix.end2, offset);
                }
                last_end = ix.end2 > 0 ? ix.end2 : last_end;
            }
        }
    }
    pia.contiguous[species - 1] = 1;
}

// particles.jl:1039-1052
inline void swap_particles_true_index(ParticleVector& a, ParticleVector& b, int64_t i, int64_t j) {
    std::swap(a.particles[i - 1], b.particles[j - 1]);
}
// particles.jl:1065-1071
inline void swap_particles(ParticleVector& a, ParticleVector& b, int64_t i, int64_t j) {
    swap_particles_true_index(a, b, a.index[i - 1], b.index[j - 1]);
}
// particles.jl:1086-1137
inline void restore_particle_ordering(ParticleVector& pv, std::vector<int64_t>& inv_map) {
    const int64_t n_total = (int64_t)pv.index.size();
    const int64_t n_used = n_total - pv.nbuffer;
    if ((int64_t)inv_map.size() < n_total) inv_map.resize(n_total);
    std::fill(inv_map.begin(), inv_map.end(), 0);
    for (int64_t i = 1; i <= n_used; i++) inv_map[pv.index[i - 1] - 1] = i;
    for (int64_t i = 1; i <= n_used; i++) {
        const int64_t cur = pv.index[i - 1];
        if (cur != i) {
            const int64_t displaced = inv_map[i - 1];
            swap_particles_true_index(pv, pv, i, cur);
            pv.index[i - 1] = i;
            inv_map[i - 1] = i;
            if (displaced > 0) {
                pv.index[displaced - 1] = cur;
                inv_map[cur - 1] = displaced;
            } else {
                inv_map[cur - 1] = 0;
            }
        }
    }
    for (int64_t i = n_used + 1; i <= n_total; i++) pv.index[i - 1] = i;
    for (int64_t i = 1; i <= pv.nbuffer; i++) pv.buffer[i - 1] = n_total - i + 1;
}

// particles.jl:863-907 ; returns 1 if ok, else 0 and *where = offending cell (0: n_total mismatch)
inline int check_pia_is_correct(const ParticleIndexerArray& pia, int64_t species, int64_t* where) {
    int64_t n_tot = 0;
    for (int64_t i = 1; i <= pia.n_cells; i++) {
        const ParticleIndexer& ix = pia.at(i, species);
        if (ix.n_local != ix.n_group1 + ix.n_group2) { *where = i; return 0; }
        n_tot += ix.n_local;
        if (ix.n_group1 > 0) { if (ix.n_group1 != ix.end1 - ix.start1 + 1) { *where = i; return 0; } }
        else if (ix.start1 != 0 || ix.end1 != -1) { *where = i; return 0; }
        if (ix.n_group2 > 0) { if (ix.n_group2 != ix.end2 - ix.start2 + 1) { *where = i; return 0; } }
        else if (ix.start2 != 0 || ix.end2 != -1) { *where = i; return 0; }
    }
    *where = 0;
    return n_tot == pia.n_total[species - 1] ? 1 : 0;
}
// particles.jl:942-988 ; returns 1 if ok, else 0 and *code as in the reference
inline int check_unique_index(const ParticleVector& pv, const ParticleIndexerArray& pia, int64_t species, int64_t* code) {
    std::vector<int64_t> n_counts(pv.index.size(), 0);
    int64_t retcode = 1;
    for (int64_t cell = 1; cell <= pia.n_cells; cell++) {
        const ParticleIndexer& ix = pia.at(cell, species);
        for (int64_t i = ix.start1; i <= ix.end1; i++) n_counts[pv.index[i - 1] - 1] += 1;
        for (int64_t i = ix.start2; i <= ix.end2; i++) n_counts[pv.index[i - 1] - 1] += 1;
    }
    int64_t n_index = 0;
    for (auto c : n_counts) n_index = std::max(n_index, c);
    for (int64_t b = 0; b < pv.nbuffer; b++)
        if (n_counts[pv.buffer[b] - 1] > 0) { retcode = -1; break; }
    if (n_index > 1) { *code = n_index * retcode; return 0; }
    if (retcode == -1) { *code = -1; return 0; }
    *code = 0;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// grids/grid_uniform1D.jl
// ---------------------------------------------------------------------------------------------
struct Grid1DUniform {  // grid_uniform1D.jl:49-86
    double L;
    int64_t n_cells;
    double dx, inv_dx, min_x, max_x;
    Grid1DUniform(double L_, int64_t nx, double wall_offset = 1e-12)
        : L(L_), n_cells(nx), dx(L_ / nx), inv_dx(1.0 / (L_ / nx)), min_x((L_ / nx) * wall_offset),
          max_x(L_ - (L_ / nx) * wall_offset) {}
    double cell_V(int64_t) const { return dx; }          // grid_uniform1D.jl:80
    double cell_xlo(int64_t cell) const { return (cell - 1) * dx; }  // :78
    double cell_xhi(int64_t cell) const { return cell * dx; }        // :79
};
inline int64_t get_cell(const Grid1DUniform& g, double x1) {  // grid_uniform1D.jl:97-99
    return (int64_t)std::floor(x1 * g.inv_dx) + 1;
}

// ---------------------------------------------------------------------------------------------
// grids/grid_sorting.jl
// ---------------------------------------------------------------------------------------------
struct GridSortInPlace {  // grid_sorting.jl:10-13
    std::vector<int64_t> cell_counts, sorted_indices;
    GridSortInPlace(int64_t n_cells, int64_t n_particles) : cell_counts(n_cells + 1, 0), sorted_indices(n_particles, 0) {}
};

// shared tail of both sort_particles! methods, grid_sorting.jl:76-112 == :145-181
inline void sort_finish(GridSortInPlace& gs, int64_t n_cells, ParticleVector& pv, ParticleIndexerArray& pia,
                        int64_t species, int64_t n_tot) {
    for (int64_t cell = 1; cell <= n_cells; cell++) {
        gs.cell_counts[cell] = gs.cell_counts[cell] + gs.cell_counts[cell - 1];
        const int64_t cell_start = gs.cell_counts[cell - 1] + 1;
        const int64_t cell_np = gs.cell_counts[cell] - gs.cell_counts[cell - 1];
        const int64_t cell_end = gs.cell_counts[cell];
        ParticleIndexer& ix = pia.at(cell, species);
        ix.start2 = 0; ix.end2 = -1; ix.n_group2 = 0;
        if (cell_np > 0) { ix.start1 = cell_start; ix.end1 = cell_end; }
        else { ix.start1 = 0; ix.end1 = -1; }
        ix.n_group1 = cell_np;
        ix.n_local = cell_np;
    }
    for (int64_t i = n_tot; i >= 1; i--) {
        const int64_t curr_cell = pv.cell[i - 1];
        gs.sorted_indices[gs.cell_counts[curr_cell] - 1] = pv.index[i - 1];
        gs.cell_counts[curr_cell] -= 1;
    }
    for (int64_t i = 1; i <= n_tot; i++) pv.index[i - 1] = gs.sorted_indices[i - 1];
    pia.contiguous[species - 1] = 1;
}
// grid_sorting.jl:58-113 (cells unknown)
inline void sort_particles(GridSortInPlace& gs, const Grid1DUniform& grid, ParticleVector& pv,
                           ParticleIndexerArray& pia, int64_t species) {
    const int64_t n_tot = pia.n_total[species - 1];
    if (n_tot > (int64_t)gs.sorted_indices.size()) gs.sorted_indices.resize(n_tot + DELTA_PARTICLES);
    std::fill(gs.cell_counts.begin(), gs.cell_counts.end(), 0);
    if (!pia.contiguous[species - 1]) squash_pia(pv, pia, species);
    for (int64_t i = 1; i <= n_tot; i++) {
        const int64_t newcell = get_cell(grid, pv[i].x[0]);
        pv.cell[i - 1] = newcell;
        gs.cell_counts[newcell] += 1;
    }
    sort_finish(gs, grid.n_cells, pv, pia, species, n_tot);
}
// grid_sorting.jl:128-182 (cells known)
inline void sort_particles(GridSortInPlace& gs, ParticleVector& pv, ParticleIndexerArray& pia, int64_t species) {
    const int64_t n_cells = pia.n_cells;
    const int64_t n_tot = pia.n_total[species - 1];
    if (n_tot > (int64_t)gs.sorted_indices.size()) gs.sorted_indices.resize(n_tot + DELTA_PARTICLES);
    std::fill(gs.cell_counts.begin(), gs.cell_counts.end(), 0);
    if (!pia.contiguous[species - 1]) squash_pia(pv, pia, species);
    for (int64_t i = 1; i <= n_tot; i++) gs.cell_counts[pv.cell[i - 1]] += 1;
    sort_finish(gs, n_cells, pv, pia, species, n_tot);
}

// ---------------------------------------------------------------------------------------------
// collisions: collision_utils.jl, collision_cross_sections.jl, collision_scattering.jl, collision_ntc.jl
// ---------------------------------------------------------------------------------------------
struct Species { double mass, charge; };  // particles.jl:31-36 (name omitted)

struct Interaction {  // collision_utils.jl:73-82
    double m_r, mu1, mu2, vhs_d, vhs_o, vhs_Tref, vhs_muref, vhs_factor;
};
inline double compute_vhs_factor(double Tref, double d, double o, double m_r) {  // collision_utils.jl:98-101
    return M_PI * d * d * std::pow(2 * k_B * Tref / m_r, o - 0.5) / std::tgamma(2.5 - o);
}
inline double compute_mu_ref(double m_r, double o, double Tref, double d) {  // collision_utils.jl:249-253 (load_interaction_data helper)
    const double numerator = 30.0 * std::sqrt(m_r * k_B * Tref);
    const double denumerator = 4.0 * std::sqrt(M_PI) * (5.0 - 2.0 * o) * (7.0 - 2.0 * o) * d * d;
    return numerator / denumerator;
}
// one entry [i,k] of load_interaction_data, collision_utils.jl:159-201 (note compute_mu_ref gets 0.5*(m_i+m_k))
inline Interaction make_interaction(double m_i, double m_k, double d, double o, double Tref) {
    Interaction it;
    it.m_r = m_i * m_k / (m_i + m_k);
    it.mu1 = m_i / (m_i + m_k);
    it.mu2 = m_k / (m_i + m_k);
    it.vhs_d = d; it.vhs_o = o; it.vhs_Tref = Tref;
    it.vhs_muref = compute_mu_ref(0.5 * (m_i + m_k), o, Tref, d);
    it.vhs_factor = compute_vhs_factor(Tref, d, o, it.m_r);
    return it;
}
inline double sigma_vhs(const Interaction& it, double g) {  // collision_cross_sections.jl:179-181
    return it.vhs_factor * std::pow(g, 1.0 - 2 * it.vhs_o);
}
// collision_utils.jl:418-423
inline double estimate_sigma_g_w_max(const Interaction& it, const Species& s1, const Species& s2, double T1, double T2,
                                     double Fnum, double mult_factor = 1.0) {
    const double g1 = std::sqrt(2 * T1 * k_B / s1.mass);
    const double g2 = std::sqrt(2 * T2 * k_B / s2.mass);
    const double g = 0.5 * (g1 + g2);
    return mult_factor * sigma_vhs(it, g) * g * Fnum;
}

struct CollisionData {  // collision_utils.jl:22-31 (fields used on the hot path)
    double v_com[3] = {0, 0, 0};
    double g = 0.0;
    double g_vec[3] = {0, 0, 0};
    double g_vec_new[3] = {0, 0, 0};
};
inline void compute_com(CollisionData& cd, const Interaction& it, const Particle& p1, const Particle& p2) {  // :374-376
    for (int d = 0; d < 3; d++) cd.v_com[d] = it.mu1 * p1.v[d] + it.mu2 * p2.v[d];
}
inline void compute_g(CollisionData& cd, const Particle& p1, const Particle& p2) {  // :388-391
    for (int d = 0; d < 3; d++) cd.g_vec[d] = p1.v[d] - p2.v[d];
    cd.g = std::sqrt(cd.g_vec[0] * cd.g_vec[0] + cd.g_vec[1] * cd.g_vec[1] + cd.g_vec[2] * cd.g_vec[2]);
}
template <class R>
inline void scatter_vhs(R& rng, CollisionData& cd, const Interaction& it, Particle& p1, Particle& p2) {  // collision_scattering.jl:17-29
    const double phi = twopi * rng.rand();
    const double cphi = std::cos(phi), sphi = std::sin(phi);
    const double ctheta = 2.0 * rng.rand() - 1.0;
    const double stheta = std::sqrt(1.0 - ctheta * ctheta);
    cd.g_vec_new[0] = cd.g * (stheta * cphi);
    cd.g_vec_new[1] = cd.g * (stheta * sphi);
    cd.g_vec_new[2] = cd.g * ctheta;
    for (int d = 0; d < 3; d++) {
        p1.v[d] = cd.v_com[d] + it.mu2 * cd.g_vec_new[d];
        p2.v[d] = cd.v_com[d] - it.mu1 * cd.g_vec_new[d];
    }
}

struct CollisionFactors {  // collision_ntc.jl:18-25
    int64_t n1 = 0, n2 = 0;
    double sigma_g_w_max = 0.0;
    int64_t n_coll = 0, n_coll_performed = 0, n_eq_w_coll_performed = 0;
};

// collision_ntc.jl:223-270 (variable weight; splits the heavier particle)
template <class R>
inline void collide_2particles_vhs(R& rng, CollisionData& cd, CollisionFactors& cf, const Interaction& it, int64_t i,
                                   int64_t k, ParticleVector& p1, ParticleVector& p2, ParticleIndexerArray& pia,
                                   int64_t cell, int64_t s1, int64_t s2, double dw_tol) {
    const double sigma = sigma_vhs(it, cd.g);
    const double sgw = sigma * cd.g * std::max(p1[i].w, p2[k].w);
    cf.sigma_g_w_max = std::max(sgw, cf.sigma_g_w_max);
    if (rng.rand() < sgw / cf.sigma_g_w_max) {
        cf.n_coll_performed += 1;
        compute_com(cd, it, p1[i], p2[k]);
        if (std::fabs(p1[i].w - p2[k].w) < dw_tol) {
            cf.n_eq_w_coll_performed += 1;
        } else if (p1[i].w > p2[k].w) {
            if (p1.length() <= pia.n_total[s1 - 1]) p1.resize(p1.length() + DELTA_PARTICLES);
            update_buffer_index_new_particle(p1, pia, cell, s1);
            const double dw = p1[i].w - p2[k].w;
            p1[i].w = p2[k].w;
            Particle& nw = p1[pia.n_total[s1 - 1]];
            nw.w = dw;
            for (int d = 0; d < 3; d++) { nw.v[d] = p1[i].v[d]; nw.x[d] = p1[i].x[d]; }
        } else {
            if (p2.length() <= pia.n_total[s2 - 1]) p2.resize(p2.length() + DELTA_PARTICLES);
            update_buffer_index_new_particle(p2, pia, cell, s2);
            const double dw = p2[k].w - p1[i].w;
            p2[k].w = p1[i].w;
            Particle& nw = p2[pia.n_total[s2 - 1]];
            nw.w = dw;
            for (int d = 0; d < 3; d++) { nw.v[d] = p2[k].v[d]; nw.x[d] = p2[k].x[d]; }
        }
        scatter_vhs(rng, cd, it, p1[i], p2[k]);
    }
}
// collision_ntc.jl:294-309
template <class R>
inline void collide_2particles_vhs_equal_weight(R& rng, CollisionData& cd, CollisionFactors& cf, const Interaction& it,
                                                Particle& pa_i, Particle& pa_k) {
    const double sigma = sigma_vhs(it, cd.g);
    const double sgw = sigma * cd.g * std::max(pa_i.w, pa_k.w);
    cf.sigma_g_w_max = std::max(sgw, cf.sigma_g_w_max);
    if (rng.rand() < sgw / cf.sigma_g_w_max) {
        cf.n_coll_performed += 1;
        cf.n_eq_w_coll_performed += 1;
        compute_com(cd, it, pa_i, pa_k);
        scatter_vhs(rng, cd, it, pa_i, pa_k);
    }
}
// collision_ntc.jl:338-380 (ntc!, one species) and :479-521 (ntc_equal_weight!, one species)
template <class R>
inline void ntc(R& rng, CollisionFactors& cf, CollisionData& cd, const Interaction& it, ParticleVector& pv,
                ParticleIndexerArray& pia, int64_t cell, int64_t species, double dt, double V, double dw_tol,
                bool equal_weight) {
    ParticleIndexer& ix = pia.at(cell, species);
    cf.n1 = ix.n_local;
    cf.n2 = ix.n_local;
    const double n_coll_float =
        0.5 * dt * (double)ix.n_local * (double)(ix.n_local - 1) * cf.sigma_g_w_max / V + rng.rand();  // :173-176
    const int64_t n_coll_int = (int64_t)std::floor(n_coll_float);
    cf.n_coll = n_coll_int;
    cf.n_coll_performed = 0;
    cf.n_eq_w_coll_performed = 0;
    for (int64_t c = 0; c < n_coll_int; c++) {
        int64_t i = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        int64_t k = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        while (i == k) k = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        i = map_cont_index(ix, i);
        k = map_cont_index(ix, k);
        compute_g(cd, pv[i], pv[k]);
        if (cd.g > EPS) {
            if (equal_weight) collide_2particles_vhs_equal_weight(rng, cd, cf, it, pv[i], pv[k]);
            else collide_2particles_vhs(rng, cd, cf, it, i, k, pv, pv, pia, cell, species, species, dw_tol);
        }
    }
}
// collision_ntc.jl:412-453 (ntc!, two species) and :554-595 (ntc_equal_weight!, two species)
template <class R>
inline void ntc2(R& rng, CollisionFactors& cf, CollisionData& cd, const Interaction& it, ParticleVector& p1,
                 ParticleVector& p2, ParticleIndexerArray& pia, int64_t cell, int64_t s1, int64_t s2, double dt, double V,
                 double dw_tol, bool equal_weight) {
    ParticleIndexer& ix1 = pia.at(cell, s1);
    ParticleIndexer& ix2 = pia.at(cell, s2);
    cf.n1 = ix1.n_local;
    cf.n2 = ix2.n_local;
    const double n_coll_float = dt * (double)ix1.n_local * (double)ix2.n_local * cf.sigma_g_w_max / V + rng.rand();  // :195-197
    const int64_t n_coll_int = (int64_t)std::floor(n_coll_float);
    cf.n_coll = n_coll_int;
    cf.n_coll_performed = 0;
    cf.n_eq_w_coll_performed = 0;
    for (int64_t c = 0; c < n_coll_int; c++) {
        int64_t i = (int64_t)std::floor(rng.rand() * (double)ix1.n_local);
        int64_t k = (int64_t)std::floor(rng.rand() * (double)ix2.n_local);
        i = map_cont_index(ix1, i);
        k = map_cont_index(ix2, k);
        compute_g(cd, p1[i], p2[k]);
        if (cd.g > EPS) {
            if (equal_weight) collide_2particles_vhs_equal_weight(rng, cd, cf, it, p1[i], p2[k]);
            else collide_2particles_vhs(rng, cd, cf, it, i, k, p1, p2, pia, cell, s1, s2, dw_tol);
        }
    }
}

// collision_swpm.jl:17-23
struct CollisionFactorsSWPM {
    int64_t n1 = 0, n2 = 0;
    double sigma_g_max = 0.0;
    int64_t n_coll = 0, n_coll_performed = 0;
};
// collision_swpm.jl:201-287
template <class R>
inline void swpm(R& rng, CollisionFactorsSWPM& cf, CollisionData& cd, const Interaction& it, ParticleVector& pv,
                 ParticleIndexerArray& pia, int64_t cell, int64_t species, double G, double dt, double V) {
    ParticleIndexer& ix = pia.at(cell, species);
    double w_max = 0.0;
    for (int64_t i = ix.start1; i <= ix.end1; i++) w_max = std::max(w_max, pv[i].w);
    if (ix.n_group2 > 0)
        for (int64_t i = ix.start2; i <= ix.end2; i++) w_max = std::max(w_max, pv[i].w);
    const double wtf = 1.0 / (1.0 + G);
    const double inv_w_max = 1.0 / w_max;
    cf.n1 = ix.n_local;
    cf.n2 = ix.n_local;
    const double n_coll_float = 0.5 * dt * (double)ix.n_local * (double)(ix.n_local - 1) * cf.sigma_g_max * w_max * (G + 1) / V +
                                rng.rand();  // :170-173
    const int64_t n_coll_int = (int64_t)std::floor(n_coll_float);
    cf.n_coll = n_coll_int;
    cf.n_coll_performed = 0;
    for (int64_t c = 0; c < n_coll_int; c++) {
        int64_t i = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        int64_t k = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        while (i == k) k = (int64_t)std::floor(rng.rand() * (double)ix.n_local);
        i = map_cont_index(ix, i);
        k = map_cont_index(ix, k);
        compute_g(cd, pv[i], pv[k]);
        if (cd.g > EPS) {
            const double sigma = sigma_vhs(it, cd.g);
            const double sg = sigma * cd.g;
            cf.sigma_g_max = std::max(sg, cf.sigma_g_max);
            if (rng.rand() < sg * std::max(pv[i].w, pv[k].w) * inv_w_max / cf.sigma_g_max) {
                cf.n_coll_performed += 1;
                compute_com(cd, it, pv[i], pv[k]);
                if (pv.length() <= pia.n_total[species - 1] + 1) pv.resize(pv.length() + DELTA_PARTICLES);  // room for 2
                const double dw = std::min(pv[i].w, pv[k].w) * wtf;
                pv[i].w -= dw;
                pv[k].w -= dw;
                update_buffer_index_new_particle(pv, pia, cell, species);
                {
                    Particle& a = pv[pia.n_total[species - 1]];
                    a.w = dw;
                    for (int d = 0; d < 3; d++) { a.v[d] = pv[i].v[d]; a.x[d] = pv[i].x[d]; }
                }
                update_buffer_index_new_particle(pv, pia, cell, species);
                {
                    Particle& b = pv[pia.n_total[species - 1]];
                    b.w = dw;
                    for (int d = 0; d < 3; d++) { b.v[d] = pv[k].v[d]; b.x[d] = pv[k].x[d]; }
                }
                scatter_vhs(rng, cd, it, pv[pia.n_total[species - 1] - 1], pv[pia.n_total[species - 1]]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// collisions/collision_fp.jl -- linear Fokker-Planck.  Only the sorted-cell path (n_group2 == 0) is
// defined in the reference (collision_fp.jl:57 uses an undefined variable on the group2 branch).
// The standard normals come from `normals(j, out3)`: the reference calls randn(rng) three times per
// particle (:164-170); the stream here is Box-Muller on Philox (see fp_normals_philox).
// ---------------------------------------------------------------------------------------------
// collision_fp.jl:182-211
inline void scale_norm_rands(std::vector<double>& xr, std::vector<double>& yr, std::vector<double>& zr, int64_t n) {
    double mean[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; i++) { mean[0] += xr[i]; mean[1] += yr[i]; mean[2] += zr[i]; }
    for (int d = 0; d < 3; d++) mean[d] /= (double)n;
    for (int64_t i = 0; i < n; i++) {
        xr[i] -= mean[0]; yr[i] -= mean[1]; zr[i] -= mean[2];
        sd[0] += xr[i] * xr[i]; sd[1] += yr[i] * yr[i]; sd[2] += zr[i] * zr[i];
    }
    for (int d = 0; d < 3; d++) sd[d] = std::sqrt((double)n / sd[d]);
    for (int64_t i = 0; i < n; i++) { xr[i] *= sd[0]; yr[i] *= sd[1]; zr[i] *= sd[2]; }
}
// collision_fp.jl:143-151
inline double compute_relaxation_time(const Interaction& it, double mass, double V, double es_old, double local_w) {
    const double T = es_old * mass / ((3.0 / 2.0) * k_B);
    const double nrho = local_w / V;
    const double p = nrho * k_B * T;
    const double mu = it.vhs_muref * std::pow(T / it.vhs_Tref, it.vhs_o);
    return 2.0 * mu / p;
}
// Normals for local particle j of a cell stream: Philox block j of `base`, two fp32 Box-Muller transforms with explicitly rounded
// operations (the header is shared with the CUDA kernel so that both sides produce the same bits; see mb_normals.h).
inline void fp_normals_philox(const PhiloxStream& base, int64_t j, double out[3]) {
    uint32_t c[4] = {(uint32_t)j, base.ctr[1], base.ctr[2], base.ctr[3]};
    uint32_t o[4];
    Philox4x32::block(c, base.key, o);
    float n0, n1, n2, n3;
    mbn_box_muller(o[0], o[1], &n0, &n1);
    mbn_box_muller(o[2], o[3], &n2, &n3);
    out[0] = (double)n0;
    out[1] = (double)n1;
    out[2] = (double)n2;
}
// collision_fp.jl:24-125; NormalSrc: void operator()(int64_t j, double out[3])
template <class NormalSrc>
inline void fp_linear(NormalSrc&& normals, const Interaction& it, double mass, ParticleVector& pv,
                      ParticleIndexerArray& pia, int64_t cell, int64_t species, double dt, double V) {
    const ParticleIndexer& ix = pia.at(cell, species);
    const int64_t n_local = ix.n_local, n_begin = ix.start1, n_end = ix.end1;
    if (n_local < 7) return;
    std::vector<double> xr(n_local), yr(n_local), zr(n_local);
    for (int64_t j = 0; j < n_local; j++) {
        double o[3];
        normals(j, o);
        xr[j] = o[0]; yr[j] = o[1]; zr[j] = o[2];
    }
    scale_norm_rands(xr, yr, zr, n_local);
    double local_w = 0.0, es_old = 0.0, es_new = 0.0;
    const double Krot = 0.0, Kvib = 0.0;
    double vel_ave[3] = {0, 0, 0};
    for (int64_t p = n_begin; p <= n_end; p++) {
        for (int d = 0; d < 3; d++) vel_ave[d] += pv[p].v[d] * pv[p].w;
        local_w += pv[p].w;
    }
    for (int d = 0; d < 3; d++) vel_ave[d] /= local_w;
    for (int64_t p = n_begin; p <= n_end; p++) {
        Particle& q = pv[p];
        for (int d = 0; d < 3; d++) q.v[d] -= vel_ave[d];
        es_old += (q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2]) * q.w;
    }
    es_old = 0.5 * es_old / local_w;
    const double tau = compute_relaxation_time(it, mass, V, es_old, local_w);
    const double A = std::exp(-dt / tau);
    const double C = std::sqrt(((2.0 / 3.0) * es_old + (Krot + Kvib) * tau) * (1.0 - std::exp(-2.0 * dt / tau)));
    for (int64_t p = n_begin; p <= n_end; p++) {
        const int64_t i = p - n_begin;
        Particle& q = pv[p];
        q.v[0] = q.v[0] * A + C * xr[i];
        q.v[1] = q.v[1] * A + C * yr[i];
        q.v[2] = q.v[2] * A + C * zr[i];
        es_new += (q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2]) * q.w;
    }
    es_new = 0.5 * es_new / local_w;
    const double alpha = std::sqrt(es_old / es_new);
    for (int64_t p = n_begin; p <= n_end; p++) {
        Particle& q = pv[p];
        for (int d = 0; d < 3; d++) q.v[d] = alpha * q.v[d] + vel_ave[d];
    }
}

// ---------------------------------------------------------------------------------------------
// properties/physical_props.jl
// ---------------------------------------------------------------------------------------------
struct PhysProps {  // physical_props.jl:24-37
    bool ndens_not_Np;
    int64_t n_cells, n_species, n_moments;
    std::vector<double> lpa, np, n, v, T;  // np,n,T: [cell + n_cells*species]; v: [comp + 3*(cell + n_cells*species)]
    std::vector<int> moment_powers;
    std::vector<double> moments;  // [m + n_moments*(cell + n_cells*species)]
    double Tref;
    PhysProps(int64_t nc, int64_t ns, std::vector<int> powers = {}, bool ndens = false, double Tref_ = 300.0)
        : ndens_not_Np(ndens), n_cells(nc), n_species(ns), n_moments((int64_t)powers.size()), lpa(ns, 0.0),
          np(nc * ns, 0.0), n(nc * ns, 0.0), v(3 * nc * ns, 0.0), T(nc * ns, 0.0), moment_powers(powers),
          moments(powers.size() * nc * ns, 0.0), Tref(Tref_) {}
};
// physical_props.jl:104-154 (moments == false) and :168-245 (moments == true)
inline void compute_props(std::vector<ParticleVector*>& particles, const ParticleIndexerArray& pia,
                          const std::vector<Species>& sd, PhysProps& pp, bool with_moments) {
    with_moments = with_moments && pp.n_moments > 0;
    for (int64_t s = 1; s <= pp.n_species; s++) {
        ParticleVector& pv = *particles[s - 1];
        double moment_factor = 0, moment_vref = 0;
        if (with_moments) {
            moment_factor = 4 * M_PI * std::pow(sd[s - 1].mass / (twopi * k_B * pp.Tref), 1.5) * 0.5;
            moment_vref = std::pow(sd[s - 1].mass / (2 * k_B * pp.Tref), 0.5);
        }
        for (int64_t cell = 1; cell <= pp.n_cells; cell++) {
            const ParticleIndexer& ix = pia.at(cell, s);
            const int64_t o = (cell - 1) + pp.n_cells * (s - 1);
            double np = 0, n = 0.0, E = 0.0, T = 0.0, v[3] = {0, 0, 0};
            if (with_moments) for (int64_t m = 0; m < pp.n_moments; m++) pp.moments[m + pp.n_moments * o] = 0.0;
            auto pass1 = [&](int64_t s0, int64_t e0) {
                for (int64_t i = s0; i <= e0; i++) {
                    const Particle& p = pv[i];
                    n += p.w;
                    for (int d = 0; d < 3; d++) v[d] = v[d] + p.v[d] * p.w;
                    np += 1;
                }
            };
            pass1(ix.start1, ix.end1);
            if (ix.n_group2 > 0) pass1(ix.start2, ix.end2);
            if (n > 0.0) {
                for (int d = 0; d < 3; d++) v[d] /= n;
                auto pass2 = [&](int64_t s0, int64_t e0) {
                    for (int64_t i = s0; i <= e0; i++) {
                        const Particle& p = pv[i];
                        const double c2 = (p.v[0] - v[0]) * (p.v[0] - v[0]) + (p.v[1] - v[1]) * (p.v[1] - v[1]) +
                                          (p.v[2] - v[2]) * (p.v[2] - v[2]);
                        if (with_moments) {
                            const double normv = std::sqrt(c2);
                            E = E + p.w * normv * normv;
                            for (int64_t m = 0; m < pp.n_moments; m++)
                                pp.moments[m + pp.n_moments * o] += p.w * std::pow(normv, pp.moment_powers[m]);
                        } else {
                            E = E + p.w * c2;
                        }
                    }
                };
                pass2(ix.start1, ix.end1);
                if (ix.n_group2 > 0) pass2(ix.start2, ix.end2);
                E *= 0.5 * sd[s - 1].mass / (n * k_B);
                T = (2.0 / 3.0) * E;
            }
            if (with_moments) {
                for (int64_t m = 0; m < pp.n_moments; m++) {
                    const int pw = pp.moment_powers[m];
                    const double scaling = moment_factor * std::pow(moment_vref, -(3 + pw)) * std::tgamma((3 + pw) / 2.0);
                    pp.moments[m + pp.n_moments * o] /= (scaling * n);
                }
            }
            pp.lpa[s - 1] = (double)pv.length();
            pp.np[o] = np; pp.n[o] = n; pp.T[o] = T;
            for (int d = 0; d < 3; d++) pp.v[d + 3 * o] = v[d];
        }
    }
}
// physical_props.jl:317-352 (Np) and :393-432 (ndens variant, inv_V = 1/dx) ; cells [cell_lo, cell_hi]
inline void compute_props_sorted(std::vector<ParticleVector*>& particles, const ParticleIndexerArray& pia,
                                 const std::vector<Species>& sd, PhysProps& pp, int64_t cell_lo, int64_t cell_hi,
                                 const Grid1DUniform* grid) {
    const bool ndens = pp.ndens_not_Np && grid != nullptr;
    for (int64_t s = 1; s <= pp.n_species; s++) {
        ParticleVector& pv = *particles[s - 1];
        for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
            const ParticleIndexer& ix = pia.at(cell, s);
            const int64_t o = (cell - 1) + pp.n_cells * (s - 1);
            double n = 0.0, E = 0.0, T = 0.0, v[3] = {0, 0, 0};
            const int64_t s1 = ix.start1, e1 = ix.end1;
            for (int64_t i = s1; i <= e1; i++) {
                const Particle& p = pv[i];
                n += p.w;
                for (int d = 0; d < 3; d++) v[d] = v[d] + p.v[d] * p.w;
            }
            const double np = e1 >= s1 ? (double)(e1 - s1) + 1.0 : 0.0;
            if (n > 0.0) {
                for (int d = 0; d < 3; d++) v[d] /= n;
                for (int64_t i = s1; i <= e1; i++) {
                    const Particle& p = pv[i];
                    E = E + p.w * ((p.v[0] - v[0]) * (p.v[0] - v[0]) + (p.v[1] - v[1]) * (p.v[1] - v[1]) +
                                   (p.v[2] - v[2]) * (p.v[2] - v[2]));
                }
                E *= 0.5 * sd[s - 1].mass / (n * k_B);
                T = (2.0 / 3.0) * E;
            }
            pp.np[o] = np;
            pp.n[o] = ndens ? n * (1.0 / grid->cell_V(cell)) : n;
            pp.T[o] = T;
            for (int d = 0; d < 3; d++) pp.v[d + 3 * o] = v[d];
        }
    }
}
// physical_props.jl:281-299
inline void avg_props(PhysProps& avg, const PhysProps& pp, double n_avg_timesteps) {
    const double inv = 1.0 / n_avg_timesteps;
    for (size_t i = 0; i < pp.lpa.size(); i++) avg.lpa[i] += pp.lpa[i] * inv;
    for (size_t i = 0; i < pp.np.size(); i++) {
        avg.np[i] += pp.np[i] * inv;
        avg.n[i] += pp.n[i] * inv;
        avg.T[i] += pp.T[i] * inv;
    }
    for (size_t i = 0; i < pp.v.size(); i++) avg.v[i] += pp.v[i] * inv;
}
// physical_props.jl:473-501
inline double compute_mixed_moment(ParticleVector& pv, const ParticleIndexerArray& pia, int64_t cell, int64_t species,
                                   const int powers[3], double sum_scaler, double res_scaler) {
    double result = 0.0;
    const ParticleIndexer& ix = pia.at(cell, species);
    auto acc = [&](int64_t s0, int64_t e0) {
        for (int64_t i = s0; i <= e0; i++) {
            const Particle& p = pv[i];
            result += p.w * sum_scaler * std::pow(p.v[0], powers[0]) * std::pow(p.v[1], powers[1]) * std::pow(p.v[2], powers[2]);
        }
    };
    acc(ix.start1, ix.end1);
    if (ix.n_group2 > 0) acc(ix.start2, ix.end2);
    return result * res_scaler;
}

// ---------------------------------------------------------------------------------------------
// properties/surface_props.jl (1-D: 2 elements, normals (1,0,0) and (-1,0,0), areas 1)
// ---------------------------------------------------------------------------------------------
struct SurfProps {  // surface_props.jl:22-50,63-64 ; layout [element + 2*species] / [comp + 3*(element + 2*species)]
    int64_t n_elements = 2, n_species;
    double normals[2][3] = {{1, 0, 0}, {-1, 0, 0}};
    double inv_areas[2] = {1.0, 1.0};
    std::vector<double> np, flux_incident, flux_reflected, force, normal_pressure, shear_pressure, kinetic_energy_flux;
    explicit SurfProps(int64_t ns)
        : n_species(ns), np(2 * ns, 0.0), flux_incident(2 * ns, 0.0), flux_reflected(2 * ns, 0.0), force(6 * ns, 0.0),
          normal_pressure(2 * ns, 0.0), shear_pressure(6 * ns, 0.0), kinetic_energy_flux(2 * ns, 0.0) {}
    void clear() {  // surface_props.jl:171-178
        for (auto* a : {&np, &flux_incident, &flux_reflected, &force, &normal_pressure, &shear_pressure, &kinetic_energy_flux})
            std::fill(a->begin(), a->end(), 0.0);
    }
};
// surface_props.jl:77-98 (sign=+1) and :111-131 (sign=-1); element is 1-based
inline void update_surface(const Particle& p, int64_t species, SurfProps& sp, int64_t element, bool incident) {
    const double px = p.w * p.v[0], py = p.w * p.v[1], pz = p.w * p.v[2];
    const double* nrm = sp.normals[element - 1];
    const double p_dot_n = px * nrm[0] + py * nrm[1] + pz * nrm[2];
    const int64_t o = (element - 1) + 2 * (species - 1);
    const double sgn = incident ? 1.0 : -1.0;
    if (incident) { sp.np[o] += 1; sp.flux_incident[o] += p.w; }
    else { sp.flux_reflected[o] -= p.w; }
    sp.force[0 + 3 * o] += sgn * px;
    sp.force[1 + 3 * o] += sgn * py;
    sp.force[2 + 3 * o] += sgn * pz;
    sp.normal_pressure[o] -= sgn * p_dot_n;
    sp.shear_pressure[0 + 3 * o] += sgn * (px - p_dot_n * nrm[0]);
    sp.shear_pressure[1 + 3 * o] += sgn * (py - p_dot_n * nrm[1]);
    sp.shear_pressure[2 + 3 * o] += sgn * (pz - p_dot_n * nrm[2]);
    sp.kinetic_energy_flux[o] += sgn * 0.5 * (px * p.v[0] + py * p.v[1] + pz * p.v[2]);
}
// surface_props.jl:144-160
inline void surface_props_scale(int64_t species, SurfProps& sp, double mass, double dt) {
    const double factor_base = mass / dt;
    for (int64_t e = 0; e < 2; e++) {
        const double factor = factor_base * sp.inv_areas[e];
        const int64_t o = e + 2 * (species - 1);
        sp.flux_incident[o] *= factor;
        sp.flux_reflected[o] *= factor;
        for (int d = 0; d < 3; d++) { sp.force[d + 3 * o] *= factor; sp.shear_pressure[d + 3 * o] *= factor; }
        sp.normal_pressure[o] *= factor;
        sp.kinetic_energy_flux[o] *= factor;
    }
}

// ---------------------------------------------------------------------------------------------
// convection/boundary_conditions.jl, convection/convection_1D.jl
// ---------------------------------------------------------------------------------------------
struct MaxwellWalls1D {  // boundary_conditions.jl:14-53
    double T[2], v[2][3], accommodation[2];
    std::vector<double> reflection_velocities_sq;  // [wall + 2*species]
    MaxwellWalls1D(const std::vector<Species>& sd, double T_l, double T_r, double vy_l, double vy_r, double a_l, double a_r) {
        T[0] = T_l; T[1] = T_r;
        v[0][0] = 0; v[0][1] = vy_l; v[0][2] = 0;
        v[1][0] = 0; v[1][1] = vy_r; v[1][2] = 0;
        accommodation[0] = a_l; accommodation[1] = a_r;
        reflection_velocities_sq.resize(2 * sd.size());
        for (size_t s = 0; s < sd.size(); s++) {
            reflection_velocities_sq[0 + 2 * s] = 2 * k_B * T_l / sd[s].mass;
            reflection_velocities_sq[1 + 2 * s] = 2 * k_B * T_r / sd[s].mass;
        }
    }
};
// boundary_conditions.jl:79-93
template <class R>
inline void diffuse_reflection_x(R& rng, Particle& p, double v_sq, double wall_normal_sign, const double wall_v[3]) {
    double Rr = std::max(1e-50, rng.rand());
    const double v_normal = wall_normal_sign * std::sqrt(-v_sq * std::log(Rr));
    Rr = std::max(1e-50, rng.rand());
    const double v_tang = std::sqrt(-v_sq * std::log(Rr));
    Rr = twopi * rng.rand();
    const double v_tang1 = std::sin(Rr) * v_tang;
    const double v_tang2 = std::cos(Rr) * v_tang;
    p.v[0] = v_normal + wall_v[0];
    p.v[1] = v_tang1 + wall_v[1];
    p.v[2] = v_tang2 + wall_v[2];
}
// boundary_conditions.jl:108-121
template <class R>
inline void reflect_particle_x(R& rng, Particle& p, double v_sq, double wall_normal_sign, const double wall_v[3], double acc) {
    if (acc == 0.0) { p.v[0] = -p.v[0]; }  // specular_reflection_x! :63-65
    else if (acc == 1.0) { diffuse_reflection_x(rng, p, v_sq, wall_normal_sign, wall_v); }
    else {
        const double Rr = rng.rand();
        if (Rr < acc) diffuse_reflection_x(rng, p, v_sq, wall_normal_sign, wall_v);
        else p.v[0] = -p.v[0];
    }
}
// convection_1D.jl:17-54 and :72-112 (with SurfProps when surf != nullptr)
template <class R>
inline void convect_single_particle(R& rng, const Grid1DUniform& grid, const MaxwellWalls1D& b, Particle& p, int64_t species,
                                    SurfProps* surf, double dt) {
    double t_rest = dt;
    double x_old = p.x[0];
    double x_new = std::fma(p.v[0], dt, p.x[0]);  // @muladd
    while (x_new >= grid.L || x_new <= 0.0) {
        int bc_id;
        double wall_normal;
        if (x_new >= grid.L) { t_rest -= std::fabs((grid.L - x_old) / p.v[0]); bc_id = 2; wall_normal = -1.0; x_old = grid.L; }
        else { t_rest -= std::fabs(x_old / p.v[0]); bc_id = 1; wall_normal = 1.0; x_old = 0.0; }
        if (surf) update_surface(p, species, *surf, bc_id, true);
        reflect_particle_x(rng, p, b.reflection_velocities_sq[(bc_id - 1) + 2 * (species - 1)], wall_normal, b.v[bc_id - 1],
                           b.accommodation[bc_id - 1]);
        if (surf) update_surface(p, species, *surf, bc_id, false);
        x_new = std::fma(p.v[0], t_rest, x_old);
    }
    if (x_new < grid.min_x) x_new = grid.min_x;
    else if (x_new > grid.max_x) x_new = grid.max_x;
    p.x[0] = x_new;
}
// convection_1D.jl:130-157, :176-206, :225-255, :274-307.  RngFor: R& operator()(int64_t logical_index_1based)
template <class RngFor>
inline void convect_particles(RngFor&& rng_for, const Grid1DUniform& grid, const MaxwellWalls1D& b, ParticleVector& pv,
                              const ParticleIndexerArray& pia, int64_t species, double mass, SurfProps* surf, double dt,
                              bool compute_cell) {
    if (surf) surf->clear();
    auto one = [&](int64_t i) {
        convect_single_particle(rng_for(i), grid, b, pv[i], species, surf, dt);
        if (compute_cell) pv.cell[i - 1] = get_cell(grid, pv[i].x[0]);
    };
    if (pia.contiguous[species - 1]) {
        const int64_t n_tot = pia.n_total[species - 1];
        for (int64_t i = 1; i <= n_tot; i++) one(i);
    } else {
        for (int64_t cell = 1; cell <= grid.n_cells; cell++) {
            const ParticleIndexer& ix = pia.at(cell, species);
            for (int64_t i = ix.start1; i <= ix.end1; i++) one(i);
            if (ix.n_group2 > 0)
                for (int64_t i = ix.start2; i <= ix.end2; i++) one(i);
        }
    }
    if (surf) surface_props_scale(species, *surf, mass, dt);
}

// ---------------------------------------------------------------------------------------------
// distributions_and_sampling.jl (initial conditions, run once before the time loop)
// ---------------------------------------------------------------------------------------------
template <class R>
inline void sample_maxwellian(R& rng, ParticleVector& pv, int64_t nparticles, int64_t offset, double m, double T,
                              const double v0[3]) {  // :432-443
    const double vscale = std::sqrt(2 * k_B * T / m);
    for (int64_t i = 1; i <= nparticles; i++) {
        const double vn = std::sqrt(-std::log(rng.rand()));
        const double vr = std::sqrt(-std::log(rng.rand()));
        const double theta1 = twopi * rng.rand();
        const double theta2 = twopi * rng.rand();
        Particle& p = pv[i + offset];
        p.v[0] = vscale * (vn * std::cos(theta1)) + v0[0];
        p.v[1] = vscale * (vr * std::cos(theta2)) + v0[1];
        p.v[2] = vscale * (vr * std::sin(theta2)) + v0[2];
    }
}
// BKW(t=0) speed magnitude: the reference draws |v| from Distributions.Chi(5) (:199-200); chi_5 = sqrt(sum of 5 N(0,1)^2)
template <class R>
inline void sample_bkw(R& rng, ParticleVector& pv, int64_t nparticles, int64_t offset, double m, double T, const double v0[3]) {  // :195-213
    const double vscale = std::sqrt(2 * k_B * T / m) * std::sqrt(0.3);
    std::vector<double> v_abs(nparticles), Th(nparticles), ph(nparticles);
    bool julia_chi = false;
    if constexpr (std::is_same_v<R, Xoshiro256pp>) julia_chi = rng.engine == 1;
    if (julia_chi) {
        // StableRNG replay: rand(rng, Chi(5), n) of Distributions.jl = sqrt of Gamma(5/2, 2) drawn with the Marsaglia-Tsang sampler
        // (samplers/gamma.jl GammaMTSampler: d = shape - 1/3, c = 1 / (3 sqrt(d)), squeeze constant 0.0331, randn + rand per trial)
        if constexpr (std::is_same_v<R, Xoshiro256pp>) {
            const double d = 2.5 - 1.0 / 3.0, c = 1.0 / (3.0 * std::sqrt(d)), kappa = d * 2.0, r = 331.0 / 10000.0;
            for (int64_t i = 0; i < nparticles; i++) {
                while (true) {
                    double x = rng.randn();
                    double cbrt_v = 1.0 + c * x;
                    while (cbrt_v <= 0.0) { x = rng.randn(); cbrt_v = 1.0 + c * x; }
                    const double v = cbrt_v * cbrt_v * cbrt_v;
                    const double u = rng.rand();
                    const double xsq = x * x;
                    if (u < 1.0 - r * (xsq * xsq) || std::log(u) < xsq / 2.0 + d * (1.0 - v + std::log(v))) { v_abs[i] = std::sqrt(v * kappa); break; }
                }
            }
        }
    }
    for (int64_t i = 0; i < nparticles && !julia_chi; i++) {
        double s = 0.0;
        for (int q = 0; q < 3; q++) {  // 3 Box-Muller pairs -> 6 normals, use 5
            const double u1 = std::max(1e-300, rng.rand()), u2 = rng.rand();
            const double r2 = -2.0 * std::log(u1);
            const double c = std::cos(twopi * u2), sn = std::sin(twopi * u2);
            s += r2 * c * c;
            if (q < 2) s += r2 * sn * sn;
        }
        v_abs[i] = std::sqrt(s);
    }
    for (int64_t i = 0; i < nparticles; i++) Th[i] = rng.rand() * M_PI;
    for (int64_t i = 0; i < nparticles; i++) ph[i] = rng.rand() * twopi;
    for (int64_t i = 1; i <= nparticles; i++) {
        const double st = std::sin(Th[i - 1]);
        Particle& p = pv[i + offset];
        p.v[0] = vscale * (v_abs[i - 1] * st * std::cos(ph[i - 1])) + v0[0];
        p.v[1] = vscale * (v_abs[i - 1] * st * std::sin(ph[i - 1])) + v0[1];
        p.v[2] = vscale * (v_abs[i - 1] * std::cos(Th[i - 1])) + v0[2];
    }
}
// :477-509 ; distribution: 0 Maxwellian, 1 BKW
template <class R>
inline void sample_particles_equal_weight(R& rng, ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species,
                                          int64_t nparticles, double m, double T, double Fnum, double xlo, double xhi,
                                          double ylo, double yhi, double zlo, double zhi, int distribution,
                                          const double v0[3], bool bkw_use_offset = false) {
    const int64_t start = pia.n_total[species - 1] + 1;
    ParticleIndexer& ix = pia.at(cell, species);
    ix.n_local = nparticles;
    pia.n_total[species - 1] += nparticles;
    ix.start1 = start; ix.end1 = start - 1 + nparticles; ix.n_group1 = nparticles;
    ix.start2 = 0; ix.end2 = -1; ix.n_group2 = 0;
    const int64_t offset = start - 1;
    const double zero[3] = {0, 0, 0};
    for (int64_t i = 1; i <= nparticles; i++) {
        double x[3];
        x[0] = xlo + rng.rand() * (xhi - xlo);
        x[1] = ylo + rng.rand() * (yhi - ylo);
        x[2] = zlo + rng.rand() * (zhi - zlo);
        add_particle(pv, i + offset, Fnum, zero, x);
        pv.cell[i + offset - 1] = cell;
    }
    if (distribution == 0) sample_maxwellian(rng, pv, nparticles, offset, m, T, v0);
    // the reference passes no offset for BKW (:507), which is only right for the first cell sampled; the device follows the
    // evident intent (bkw_use_offset) when a range of cells is sampled
    else sample_bkw(rng, pv, nparticles, bkw_use_offset ? offset : 0, m, T, v0);
}
// grid_uniform1D.jl:198-219 (ndens version) over cells [cell_lo, cell_hi]
template <class R>
inline void sample_particles_equal_weight_grid(R& rng, const Grid1DUniform& grid, ParticleVector& pv, ParticleIndexerArray& pia,
                                               int64_t species, double mass, double ndens, double T, double Fnum,
                                               int64_t cell_lo, int64_t cell_hi) {
    const double zero[3] = {0, 0, 0};
    for (int64_t cell = cell_lo; cell <= cell_hi; cell++) {
        const double n_in_cell = ndens * grid.cell_V(cell);
        const double ppc = n_in_cell / Fnum;
        int64_t ppc_int = (int64_t)std::floor(ppc);
        const double remainder = ppc - (double)ppc_int;
        if (rng.rand() < remainder) ppc_int += 1;
        sample_particles_equal_weight(rng, pv, pia, cell, species, ppc_int, mass, T, Fnum, grid.cell_xlo(cell),
                                      grid.cell_xhi(cell), 0.0, 1.0, 0.0, 1.0, 0, zero);
    }
}
inline double bkw_vdf(double vx, double vy, double vz, double m, double T, double scaled_time) {  // :168-177
    const double xk = 1.0 - 0.4 * std::exp(-scaled_time / 6.0);
    const double Csq = vx * vx + vy * vy + vz * vz;
    // `5 * xk - 3` is a fused multiply-add in the reference (@muladd): with xk = 0.6 it gives -1.1e-16 instead of 0, which moves the last
    // bit of many weights -- and the octree merge's choice between mirror-image bins of this symmetric lattice hangs on those bits
    // exp as Julia evaluates it (mb_jlexp.h): its last bit differs from glibc's for ~1 % of the arguments, same consequence
    return (std::fma(5.0, xk, -3.0) + 2 * (1.0 - xk) * Csq * m / (2 * k_B * xk * T)) * mbjl::exp(-Csq * m / (2 * k_B * xk * T));
}
inline double maxwellian_vdf(double vx, double vy, double vz, double m, double T) {  // :150-152
    return std::pow(m / (2.0 * M_PI * k_B * T), 1.5) * mbjl::exp(-m * (vx * vx + vy * vy + vz * vz) / (2.0 * k_B * T));
}
// sample_on_grid! :312-346 with evaluate_distribution_on_grid! :253-268 ; vdf: 0 Maxwellian, 1 BKW(t=0)
template <class R>
inline int64_t sample_on_grid(R& rng, int vdf_kind, ParticleVector& pv, int64_t nv, double m, double T, double n_total,
                              double xlo, double xhi, double ylo, double yhi, double zlo, double zhi, double v_mult,
                              double cutoff_mult, double noise, const double v_offset[3], int64_t offset = 0) {
    const double v_thermal = std::sqrt(2 * k_B * T / m);
    std::vector<double> g(nv);
    // LinRange(-1.0, 1.0, nv)[i + 1] = lerpi(i, nv - 1, -1.0, 1.0) = (1 - t) * a + t * b with t = i / (nv - 1) (Julia Base range.jl)
    for (int64_t i = 0; i < nv; i++) { const double t = (double)i / (double)(nv - 1); g[i] = (1.0 - t) * -1.0 + t * 1.0; }
    const double dxu = g[1] - g[0];
    const double vmax = v_thermal * v_mult;
    std::vector<double> vg(nv);
    for (int64_t i = 0; i < nv; i++) vg[i] = g[i] * vmax;
    const double dv = dxu * vmax;
    const double cutoff_v = v_thermal * cutoff_mult;
    std::vector<double> w(nv * nv * nv, 0.0);
    double wsum = 0.0;
    for (int64_t k = 0; k < nv; k++)
        for (int64_t j = 0; j < nv; j++)
            for (int64_t i = 0; i < nv; i++)
                if (std::sqrt(vg[i] * vg[i] + vg[j] * vg[j] + vg[k] * vg[k]) <= cutoff_v) {
                    const double f = vdf_kind == 0 ? maxwellian_vdf(vg[i], vg[j], vg[k], m, T) : bkw_vdf(vg[i], vg[j], vg[k], m, T, 0.0);
                    w[i + nv * (j + nv * k)] = f;
                    wsum += f;
                }
    for (auto& f : w) f = f * n_total / wsum;
    int64_t pid = 0;
    for (int64_t k = 0; k < nv; k++)
        for (int64_t j = 0; j < nv; j++)
            for (int64_t i = 0; i < nv; i++) {
                const double f = w[i + nv * (j + nv * k)];
                if (f > 0.0) {
                    pid += 1;
                    double v[3], x[3];
                    v[0] = vg[i] + noise * dv * (0.5 - rng.rand()) + v_offset[0];
                    v[1] = vg[j] + noise * dv * (0.5 - rng.rand()) + v_offset[1];
                    v[2] = vg[k] + noise * dv * (0.5 - rng.rand()) + v_offset[2];
                    x[0] = xlo + rng.rand() * (xhi - xlo);
                    x[1] = ylo + rng.rand() * (yhi - ylo);
                    x[2] = zlo + rng.rand() * (zhi - zlo);
                    add_particle(pv, pid + offset, f, v, x);
                }
            }
    return pid;
}

}  // namespace mbo

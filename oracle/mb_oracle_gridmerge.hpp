// TEST INFRASTRUCTURE ONLY -- CPU oracle, velocity-grid N:2 merging (see mb_oracle.hpp header for the rules).
// Restates /root/reference/src/merging/merging_grid.jl; line numbers cite that file.
#pragma once
#include "mb_oracle.hpp"

namespace mbo {

struct GridCell {  // :26-41
    int64_t np = 0;
    double w = 0.0;
    double v_mean[3] = {0, 0, 0}, v_std_sq[3] = {0, 0, 0}, x_mean[3] = {0, 0, 0}, x_std_sq[3] = {0, 0, 0};
    int64_t particle_index1 = 0, particle_index2 = 0;
    double w1 = 0, w2 = 0, v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0}, x1[3] = {0, 0, 0}, x2[3] = {0, 0, 0};
};

struct GridN2Merge {  // :72-116
    int64_t Nx, Ny, Nz, NyNz, Ntotal;
    double extent_multiplier[3];
    double extent_v_lower[3] = {0, 0, 0}, extent_v_upper[3] = {0, 0, 0}, extent_v_mid[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dv_inv[3] = {0, 0, 0};
    std::vector<GridCell> cells;
    GridN2Merge(int64_t nx, int64_t ny, int64_t nz, const double mult[3])
        : Nx(nx), Ny(ny), Nz(nz), NyNz(ny * nz), Ntotal(nx * ny * nz + 8), cells(nx * ny * nz + 8) {
        for (int d = 0; d < 3; d++) extent_multiplier[d] = mult[d];
    }
};

// :190-197 (extent from the cell's temperature and mean velocity as stored in PhysProps)
inline void compute_velocity_extent_props(GridN2Merge& mg, double T, const double v[3], double mass) {
    const double N[3] = {(double)mg.Nx, (double)mg.Ny, (double)mg.Nz};
    for (int d = 0; d < 3; d++) {
        const double dv = mg.extent_multiplier[d] * std::sqrt(2 * T * k_B / mass);
        mg.extent_v_lower[d] = v[d] - dv;
        mg.extent_v_upper[d] = v[d] + dv;
        mg.extent_v_mid[d] = v[d];
        mg.dv[d] = 2 * dv / N[d];
        mg.dv_inv[d] = 1.0 / mg.dv[d];
    }
}
// :210-221 (explicit extents)
inline void compute_velocity_extent_given(GridN2Merge& mg, const double ext6[6]) {
    const double N[3] = {(double)mg.Nx, (double)mg.Ny, (double)mg.Nz};
    for (int d = 0; d < 3; d++) {
        mg.extent_v_lower[d] = ext6[2 * d];
        mg.extent_v_upper[d] = ext6[2 * d + 1];
        mg.extent_v_mid[d] = 0.5 * (ext6[2 * d] + ext6[2 * d + 1]);
        mg.dv[d] = (ext6[2 * d + 1] - ext6[2 * d]) / N[d];
        mg.dv_inv[d] = 1.0 / mg.dv[d];
    }
}
// :235-264 (1-based index; the last 8 cells are the octants outside the grid)
inline int64_t compute_grid_index(const GridN2Merge& mg, const double v[3]) {
    bool outside = false;
    if (v[0] < mg.extent_v_lower[0] || v[0] > mg.extent_v_upper[0]) outside = true;
    else if (v[1] < mg.extent_v_lower[1] || v[1] > mg.extent_v_upper[1]) outside = true;
    else if (v[2] < mg.extent_v_lower[2] || v[2] > mg.extent_v_upper[2]) outside = true;
    if (!outside) {
        const double ix = (v[0] - mg.extent_v_lower[0]) * mg.dv_inv[0];
        const double iy = (v[1] - mg.extent_v_lower[1]) * mg.dv_inv[1];
        const double iz = (v[2] - mg.extent_v_lower[2]) * mg.dv_inv[2];
        return (int64_t)std::floor(ix) * mg.NyNz + (int64_t)std::floor(iy) * mg.Nz + (int64_t)std::floor(iz) + 1;
    }
    int64_t index = mg.Ntotal - 7;
    if (v[0] > mg.extent_v_mid[0]) index += 1;
    if (v[1] > mg.extent_v_mid[1]) index += 2;
    if (v[2] > mg.extent_v_mid[2]) index += 4;
    return index;
}
// :278-290
inline void clear_merging_grid(GridN2Merge& mg) {
    for (auto& c : mg.cells) {
        c.w = 0.0; c.np = 0; c.particle_index1 = 0; c.particle_index2 = 0;
        for (int d = 0; d < 3; d++) { c.v_mean[d] = 0; c.v_std_sq[d] = 0; c.x_mean[d] = 0; c.x_std_sq[d] = 0; }
    }
}
// :304-375 ; returns false if a particle fell outside 1..Ntotal (v exactly on the upper bound: the reference would index out of range)
inline bool compute_grid(GridN2Merge& mg, ParticleVector& pv, const ParticleIndexerArray& pia, int64_t cell, int64_t species) {
    clear_merging_grid(mg);
    const ParticleIndexer& ix = pia.at(cell, species);
    bool ok = true;
    auto pass1 = [&](int64_t i) {
        const Particle& p = pv[i];
        const int64_t index = compute_grid_index(mg, p.v);
        if (index < 1 || index > mg.Ntotal) { ok = false; return; }
        GridCell& c = mg.cells[index - 1];
        c.np += 1;
        c.w += p.w;
        for (int d = 0; d < 3; d++) { c.v_mean[d] = c.v_mean[d] + p.v[d] * p.w; c.x_mean[d] = c.x_mean[d] + p.x[d] * p.w; }
        if (c.np == 1) c.particle_index1 = i;
        else if (c.np == 2) c.particle_index2 = i;
    };
    for (int64_t i = ix.start1; i <= ix.end1; i++) pass1(i);
    if (ix.start2 > 0)
        for (int64_t i = ix.start2; i <= ix.end2; i++) pass1(i);
    if (!ok) return false;
    for (auto& c : mg.cells) {
        if (c.w > 0.0) {
            for (int d = 0; d < 3; d++) { c.v_mean[d] = c.v_mean[d] / c.w; c.x_mean[d] = c.x_mean[d] / c.w; }
        } else {
            c.np = 0;
        }
    }
    auto pass2 = [&](int64_t i) {
        const Particle& p = pv[i];
        GridCell& c = mg.cells[compute_grid_index(mg, p.v) - 1];
        for (int d = 0; d < 3; d++) {
            const double dvv = p.v[d] - c.v_mean[d], dxx = p.x[d] - c.x_mean[d];
            c.v_std_sq[d] = c.v_std_sq[d] + (dvv * dvv) * p.w;
            c.x_std_sq[d] = c.x_std_sq[d] + (dxx * dxx) * p.w;
        }
    };
    for (int64_t i = ix.start1; i <= ix.end1; i++) pass2(i);
    if (ix.start2 > 0)
        for (int64_t i = ix.start2; i <= ix.end2; i++) pass2(i);
    for (auto& c : mg.cells)
        if (c.w > 0.0)
            for (int d = 0; d < 3; d++) { c.v_std_sq[d] = c.v_std_sq[d] / c.w; c.x_std_sq[d] = c.x_std_sq[d] / c.w; }
    return true;
}
// :393-468 (grid == nullptr) and :484-575 (1-D: x1 of every np >= 2 output clamped into [min_x, max_x]).
// SignSrc: void operator()(int64_t grid_index_1based, double sv[3], double sx[3])
template <class SignSrc>
inline void compute_new_particles_grid(SignSrc&& signs, GridN2Merge& mg, ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species,
                                       const Grid1DUniform* grid) {
    for (int64_t index = 1; index <= mg.Ntotal; index++) {
        GridCell& c = mg.cells[index - 1];
        if (c.np > 2) {
            c.w1 = 0.5 * c.w;
            c.w2 = c.w1;
            for (int d = 0; d < 3; d++) { c.v_std_sq[d] = std::sqrt(c.v_std_sq[d]); c.x_std_sq[d] = std::sqrt(c.x_std_sq[d]); }
            double sv[3], sx[3];
            signs(index, sv, sx);
            for (int d = 0; d < 3; d++) {
                c.v1[d] = c.v_mean[d] + sv[d] * c.v_std_sq[d];
                c.v2[d] = c.v_mean[d] - sv[d] * c.v_std_sq[d];
                c.x1[d] = c.x_mean[d] + sx[d] * c.x_std_sq[d];
                c.x2[d] = c.x_mean[d] - sx[d] * c.x_std_sq[d];
            }
        } else if (c.np == 2) {
            const Particle& a = pv[c.particle_index1];
            c.w1 = a.w;
            for (int d = 0; d < 3; d++) { c.v1[d] = a.v[d]; c.x1[d] = a.x[d]; }
            const Particle& b = pv[c.particle_index2];
            c.w2 = b.w;
            for (int d = 0; d < 3; d++) { c.v2[d] = b.v[d]; c.x2[d] = b.x[d]; }
        } else if (c.np == 1) {
            const Particle& a = pv[c.particle_index1];
            c.w1 = a.w;
            for (int d = 0; d < 3; d++) { c.v1[d] = a.v[d]; c.x1[d] = a.x[d]; }
        }
    }
    const ParticleIndexer ixc = pia.at(cell, species);
    int64_t curr = 0;
    auto put = [&](double w, const double v[3], const double x[3], bool clampx) {
        const int64_t i = map_cont_index(ixc, curr);
        curr += 1;
        Particle& p = pv[i];
        p.w = w;
        for (int d = 0; d < 3; d++) { p.v[d] = v[d]; p.x[d] = x[d]; }
        if (clampx && grid) {
            if (x[0] < grid->min_x) p.x[0] = grid->min_x;
            else if (x[0] > grid->max_x) p.x[0] = grid->max_x;
        }
    };
    for (int64_t index = 1; index <= mg.Ntotal; index++) {
        const GridCell& c = mg.cells[index - 1];
        if (c.np >= 2) { put(c.w1, c.v1, c.x1, true); put(c.w2, c.v2, c.x2, true); }  // the 1-D variant clamps both np > 2 and np == 2 outputs (:528-549)
        else if (c.np == 1) put(c.w1, c.v1, c.x1, false);
    }
    const int64_t n_delete = ixc.n_local - curr;
    if (!(cell == pia.n_cells) || (n_delete > ixc.n_group2)) pia.contiguous[species - 1] = 0;
    for (int64_t q = 0; q < n_delete; q++) delete_particle_end(pv, pia, cell, species);
}

// :597-703: ext6 == nullptr -> extents from (T, v) of the cell; else the explicit extents
template <class SignSrc>
inline bool merge_grid_based(SignSrc&& signs, GridN2Merge& mg, ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell, int64_t species, double mass,
                             double T, const double v[3], const double* ext6, const Grid1DUniform* grid) {
    if (ext6) compute_velocity_extent_given(mg, ext6);
    else compute_velocity_extent_props(mg, T, v, mass);
    if (!compute_grid(mg, pv, pia, cell, species)) return false;
    compute_new_particles_grid(signs, mg, pv, pia, cell, species, grid);
    return true;
}

}  // namespace mbo

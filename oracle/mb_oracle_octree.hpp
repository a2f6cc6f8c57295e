// TEST INFRASTRUCTURE ONLY -- CPU oracle, octree N:2 merging (see mb_oracle.hpp header for the rules).
// Restates /root/reference/src/merging/merging_octree_N2.jl; line numbers cite that file.
#pragma once
#include "mb_oracle.hpp"

namespace mbo {

enum OctreeBinSplit { OctreeBinMidSplit = 1, OctreeBinMeanSplit = 2, OctreeBinMedianSplit = 3 };       // :12
enum OctreeInitBin { OctreeInitBinMinMaxVel = 1, OctreeInitBinMinMaxVelSym = 2, OctreeInitBinC = 3 };  // :24
enum OctreeBinBounds { OctreeBinBoundsInherit = 1, OctreeBinBoundsRecompute = 2 };                     // :34

struct OctreeCell {  // :49-59
    int64_t np = 0;
    double w = 0.0;
    double v_min[3] = {0, 0, 0}, v_max[3] = {0, 0, 0};
    int64_t depth = 0;
    bool can_be_refined = true;
};
struct OctreeFullCell {  // :82-96
    double v_mean[3] = {0, 0, 0}, v_std_sq[3] = {0, 0, 0}, x_mean[3] = {0, 0, 0}, x_std_sq[3] = {0, 0, 0};
    int64_t particle_index1 = 0, particle_index2 = 0;
    double w1 = 0, w2 = 0, v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0}, x1[3] = {0, 0, 0}, x2[3] = {0, 0, 0};
};

struct OctreeN2Merge {  // :131-179, ctor :235-246
    int64_t max_Nbins, Nbins = 0;
    std::vector<OctreeCell> bins;
    std::vector<OctreeFullCell> full_bins;
    int64_t n_particles = 0;
    std::vector<int64_t> bin_start, bin_end, particle_indexes_sorted, particle_octants, particles_sort_output;
    int64_t particle_in_bin_counter[8], nonempty_counter[8], nonempty_bins[8];
    double ndens_counter[8];
    OctreeBinBounds bin_bounds_compute;
    OctreeBinSplit split;
    double vel_middle[3] = {0, 0, 0}, v_min_parent[3] = {0, 0, 0}, v_max_parent[3] = {0, 0, 0}, direction_vec[3] = {0, 0, 0};
    OctreeInitBin init_bin_bounds;
    int64_t max_depth, total_post_merge_np = 0;
    OctreeN2Merge(OctreeBinSplit split_, OctreeInitBin init = OctreeInitBinMinMaxVel,
                  OctreeBinBounds bounds = OctreeBinBoundsInherit, int64_t max_Nbins_ = 4096, int64_t max_depth_ = 10)
        : max_Nbins(max_Nbins_), bins(max_Nbins_), full_bins(max_Nbins_), bin_start(max_Nbins_, 0), bin_end(max_Nbins_, 0),
          particle_indexes_sorted(8192, 0), particle_octants(8192, 0), particles_sort_output(8192, 0),
          bin_bounds_compute(bounds), split(split_), init_bin_bounds(init), max_depth(max_depth_) {}
};

inline void resize_octree_buffers(OctreeN2Merge& oc, int64_t n) {  // :270-280
    if ((int64_t)oc.particle_indexes_sorted.size() < n) oc.particle_indexes_sorted.resize(n + DELTA_PARTICLES);
    if ((int64_t)oc.particle_octants.size() < n) oc.particle_octants.resize(n + DELTA_PARTICLES);
    if ((int64_t)oc.particles_sort_output.size() < n) oc.particles_sort_output.resize(n + DELTA_PARTICLES);
}
inline int compute_octant(const double v[3], const double mid[3]) {  // :304-316 (strict >)
    int oct = 1;
    if (v[0] > mid[0]) oct += 1;
    if (v[1] > mid[1]) oct += 2;
    if (v[2] > mid[2]) oct += 4;
    return oct;
}
inline void bin_bounds_inherit(OctreeN2Merge& oc, int64_t bin_id, const double lo[3], const double hi[3], const double mid[3],
                               int octant) {  // :341-367
    OctreeCell& b = oc.bins[bin_id - 1];
    const int o = octant - 1;
    for (int d = 0; d < 3; d++) {
        const bool upper = (o >> d) & 1;
        b.v_min[d] = upper ? mid[d] : lo[d];
        b.v_max[d] = upper ? hi[d] : mid[d];
    }
}
inline void bin_bounds_recompute(OctreeN2Merge& oc, int64_t bin_id, int64_t bs, int64_t be, ParticleVector& pv) {  // :382-418
    double mn[3] = {9299792458.0, 9299792458.0, 9299792458.0};
    double mx[3] = {-9299792458.0, -9299792458.0, -9299792458.0};
    for (int64_t i = bs; i <= be; i++) {
        const Particle& p = pv[oc.particle_indexes_sorted[i - 1]];
        for (int d = 0; d < 3; d++) {
            if (p.v[d] < mn[d]) mn[d] = p.v[d];
            if (p.v[d] > mx[d]) mx[d] = p.v[d];
        }
        for (int d = 0; d < 3; d++) { oc.bins[bin_id - 1].v_min[d] = mn[d]; oc.bins[bin_id - 1].v_max[d] = mx[d]; }
    }
}
inline void compute_v_mean(OctreeN2Merge& oc, int64_t bs, int64_t be, ParticleVector& pv) {  // :431-439 (vel_middle NOT zeroed first)
    double n_tot = 0.0;
    for (int64_t i = bs; i <= be; i++) {
        const Particle& p = pv[oc.particle_indexes_sorted[i - 1]];
        n_tot += p.w;
        for (int d = 0; d < 3; d++) oc.vel_middle[d] = oc.vel_middle[d] + p.w * p.v[d];
    }
    for (int d = 0; d < 3; d++) oc.vel_middle[d] = oc.vel_middle[d] / n_tot;
}
inline int64_t get_new_bin_id(int64_t i, int64_t bin_id, int64_t Nbins) { return i == 1 ? bin_id : Nbins + i - 1; }  // :487-489
inline int64_t get_bin_post_merge_np(const OctreeN2Merge& oc, int64_t bin_id) {  // :715-720
    return oc.bins[bin_id - 1].np >= 2 ? 2 : oc.bins[bin_id - 1].np;
}

inline void split_bin(OctreeN2Merge& oc, int64_t bin_id, ParticleVector& pv) {  // :503-633
    int64_t n_nonempty = 0;
    for (int q = 0; q < 8; q++) { oc.particle_in_bin_counter[q] = 0; oc.ndens_counter[q] = 0.0; oc.nonempty_bins[q] = 0; oc.nonempty_counter[q] = 0; }
    const int64_t current_depth = oc.bins[bin_id - 1].depth;
    const int64_t bs = oc.bin_start[bin_id - 1], be = oc.bin_end[bin_id - 1];
    if (oc.bin_bounds_compute == OctreeBinBoundsRecompute) bin_bounds_recompute(oc, bin_id, bs, be, pv);
    if (oc.split == OctreeBinMidSplit) {
        for (int d = 0; d < 3; d++) oc.vel_middle[d] = 0.5 * (oc.bins[bin_id - 1].v_min[d] + oc.bins[bin_id - 1].v_max[d]);
    } else if (oc.split == OctreeBinMeanSplit) {
        compute_v_mean(oc, bs, be, pv);
    }  // OctreeBinMedianSplit (:452-465, documented "probably not fully correct") is out of scope
    for (int64_t i = bs; i <= be; i++) {
        const Particle& p = pv[oc.particle_indexes_sorted[i - 1]];
        const int oct = compute_octant(p.v, oc.vel_middle);
        oc.particle_in_bin_counter[oct - 1] += 1;
        oc.particle_octants[i - bs] = oct;
        oc.ndens_counter[oct - 1] += p.w;
    }
    int64_t n_eb = 0;
    if (oc.particle_in_bin_counter[0] > 0) {
        n_nonempty += 1; n_eb += 1;
        oc.nonempty_counter[n_eb - 1] = oc.particle_in_bin_counter[0];
        oc.nonempty_bins[n_eb - 1] = 1;
    }
    for (int i = 2; i <= 8; i++) {
        if (oc.particle_in_bin_counter[i - 1] > 0) {
            n_nonempty += 1; n_eb += 1;
            oc.nonempty_counter[n_eb - 1] = oc.particle_in_bin_counter[i - 1];
            oc.nonempty_bins[n_eb - 1] = i;
        }
        oc.particle_in_bin_counter[i - 1] += oc.particle_in_bin_counter[i - 2];
    }
    oc.bin_end[bin_id - 1] = oc.bin_start[bin_id - 1] + oc.nonempty_counter[0] - 1;
    for (int64_t i = 2; i <= n_nonempty; i++) {
        const int64_t bi = get_new_bin_id(i, bin_id, oc.Nbins), bim1 = get_new_bin_id(i - 1, bin_id, oc.Nbins);
        oc.bin_start[bi - 1] = oc.bin_end[bim1 - 1] + 1;
        oc.bin_end[bi - 1] = oc.bin_start[bi - 1] + oc.nonempty_counter[i - 1] - 1;
    }
    oc.total_post_merge_np -= 2;
    if (oc.bin_bounds_compute == OctreeBinBoundsInherit) {
        for (int d = 0; d < 3; d++) { oc.v_min_parent[d] = oc.bins[bin_id - 1].v_min[d]; oc.v_max_parent[d] = oc.bins[bin_id - 1].v_max[d]; }
    }
    for (int64_t i = 1; i <= n_nonempty; i++) {
        const int64_t bi = get_new_bin_id(i, bin_id, oc.Nbins);
        if (oc.bin_bounds_compute == OctreeBinBoundsInherit)
            bin_bounds_inherit(oc, bi, oc.v_min_parent, oc.v_max_parent, oc.vel_middle, (int)oc.nonempty_bins[i - 1]);
        OctreeCell& b = oc.bins[bi - 1];
        b.np = oc.nonempty_counter[i - 1];
        b.w = oc.ndens_counter[oc.nonempty_bins[i - 1] - 1];
        b.depth = current_depth + 1;
        oc.total_post_merge_np += get_bin_post_merge_np(oc, bi);
        b.can_be_refined = (b.np > 2) && (b.depth < oc.max_depth);
    }
    for (int64_t i = bs; i <= be; i++) {
        const int64_t pin = oc.particle_indexes_sorted[i - 1];
        const int64_t j = oc.particle_octants[i - bs];
        oc.particles_sort_output[oc.particle_in_bin_counter[j - 1] - 1] = pin;
        oc.particle_in_bin_counter[j - 1] -= 1;
    }
    for (int64_t i = bs; i <= be; i++) oc.particle_indexes_sorted[i - 1] = oc.particles_sort_output[i - bs];
    oc.Nbins += n_nonempty - 1;
}

inline void compute_bin_props(OctreeN2Merge& oc, int64_t bin_id, ParticleVector& pv) {  // :646-700
    const int64_t bs = oc.bin_start[bin_id - 1], be = oc.bin_end[bin_id - 1];
    OctreeCell& b = oc.bins[bin_id - 1];
    OctreeFullCell& f = oc.full_bins[bin_id - 1];
    if (b.w == 0) { b.np = 0; return; }
    if (b.np == 1) f.particle_index1 = oc.particle_indexes_sorted[bs - 1];
    else if (b.np >= 2) { f.particle_index1 = oc.particle_indexes_sorted[bs - 1]; f.particle_index2 = oc.particle_indexes_sorted[bs]; }
    if (b.np <= 2) return;
    double vm[3] = {0, 0, 0}, vs[3] = {0, 0, 0}, xm[3] = {0, 0, 0}, xs[3] = {0, 0, 0};
    const double inv_w = 1.0 / b.w;
    for (int64_t i = bs; i <= be; i++) {
        const Particle& p = pv[oc.particle_indexes_sorted[i - 1]];
        for (int d = 0; d < 3; d++) { vm[d] = vm[d] + p.w * p.v[d]; xm[d] = xm[d] + p.w * p.x[d]; }
    }
    for (int d = 0; d < 3; d++) { vm[d] *= inv_w; xm[d] *= inv_w; }
    for (int64_t i = bs; i <= be; i++) {
        const Particle& p = pv[oc.particle_indexes_sorted[i - 1]];
        for (int d = 0; d < 3; d++) {
            vs[d] = vs[d] + p.w * (p.v[d] - vm[d]) * (p.v[d] - vm[d]);
            xs[d] = xs[d] + p.w * (p.x[d] - xm[d]) * (p.x[d] - xm[d]);
        }
    }
    for (int d = 0; d < 3; d++) { f.v_mean[d] = vm[d]; f.v_std_sq[d] = vs[d] * inv_w; f.x_mean[d] = xm[d]; f.x_std_sq[d] = xs[d] * inv_w; }
}

inline void init_octree(OctreeN2Merge& oc, ParticleVector& pv, const ParticleIndexerArray& pia, int64_t cell, int64_t species) {  // :947-984
    const ParticleIndexer& ix = pia.at(cell, species);
    oc.Nbins = 1;
    for (int64_t q = 0; q < ix.n_group1; q++) oc.particle_indexes_sorted[q] = ix.start1 + q;
    if (ix.n_group2 > 0)
        for (int64_t q = 0; q < ix.n_group2; q++) oc.particle_indexes_sorted[ix.n_group1 + q] = ix.start2 + q;
    oc.n_particles = ix.n_local;
    OctreeCell& b = oc.bins[0];
    b.depth = 0;
    oc.bin_start[0] = 1;
    oc.bin_end[0] = oc.n_particles;
    b.np = oc.n_particles;
    b.w = 1e50;
    oc.total_post_merge_np = get_bin_post_merge_np(oc, 1);
    b.can_be_refined = (b.np > 2) && (oc.max_depth > 0);
    if (oc.init_bin_bounds == OctreeInitBinC) {
        for (int d = 0; d < 3; d++) { b.v_min[d] = -c_light; b.v_max[d] = c_light; }
    } else {
        bin_bounds_recompute(oc, 1, 1, oc.n_particles, pv);
        if (oc.init_bin_bounds == OctreeInitBinMinMaxVelSym) {
            for (int d = 0; d < 3; d++) {
                const double m = std::max(std::fabs(b.v_min[d]), std::fabs(b.v_max[d]));
                b.v_min[d] = -m; b.v_max[d] = m;
            }
        }
    }
}

inline void compute_octree(OctreeN2Merge& oc, ParticleVector& pv, int64_t target_np) {  // :998-1039
    while (true) {
        int64_t refine_id = -1;
        double max_w = -1.0;
        for (int64_t bin_id = 1; bin_id <= oc.Nbins; bin_id++) {
            if (oc.bins[bin_id - 1].w > max_w && oc.bins[bin_id - 1].can_be_refined) { max_w = oc.bins[bin_id - 1].w; refine_id = bin_id; }
        }
        if (refine_id == -1) break;
        else if (oc.total_post_merge_np + 14 > target_np) break;
        else split_bin(oc, refine_id, pv);
        if (oc.Nbins + 7 > oc.max_Nbins) break;
    }
    if (oc.Nbins == 1) {
        oc.bins[0].w = 0.0;
        for (int64_t ii = 1; ii <= oc.n_particles; ii++) oc.bins[0].w += pv[oc.particle_indexes_sorted[ii - 1]].w;
    }
    for (int64_t bin_id = 1; bin_id <= oc.Nbins; bin_id++) compute_bin_props(oc, bin_id, pv);
}

// :736-813 (grid == nullptr) and :830-933 (1-D grid: clamp x1 of np>2 outputs into [min_x, max_x]).
// Signs: `rand(rng, direction_signs, 3)` twice per np>2 bin (:749,:753). SignSrc: void operator()(int64_t bin_id_1based, double sv[3], double sx[3])
template <class SignSrc>
inline void compute_new_particles(SignSrc&& signs, OctreeN2Merge& oc, ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell,
                                  int64_t species, const Grid1DUniform* grid) {
    for (int64_t bin_id = 1; bin_id <= oc.Nbins; bin_id++) {
        const int64_t loc_np = oc.bins[bin_id - 1].np;
        OctreeFullCell& f = oc.full_bins[bin_id - 1];
        if (loc_np > 2) {
            f.w1 = 0.5 * oc.bins[bin_id - 1].w;
            f.w2 = f.w1;
            for (int d = 0; d < 3; d++) { f.v_std_sq[d] = std::sqrt(f.v_std_sq[d]); f.x_std_sq[d] = std::sqrt(f.x_std_sq[d]); }
            double sv[3], sx[3];
            signs(bin_id, sv, sx);
            for (int d = 0; d < 3; d++) {
                f.v1[d] = f.v_mean[d] + sv[d] * f.v_std_sq[d];
                f.v2[d] = f.v_mean[d] - sv[d] * f.v_std_sq[d];
                f.x1[d] = f.x_mean[d] + sx[d] * f.x_std_sq[d];
                f.x2[d] = f.x_mean[d] - sx[d] * f.x_std_sq[d];
            }
        } else if (loc_np == 2) {
            const Particle& a = pv[f.particle_index1];
            f.w1 = a.w;
            for (int d = 0; d < 3; d++) { f.v1[d] = a.v[d]; f.x1[d] = a.x[d]; }
            const Particle& b = pv[f.particle_index2];
            f.w2 = b.w;
            for (int d = 0; d < 3; d++) { f.v2[d] = b.v[d]; f.x2[d] = b.x[d]; }
        } else if (loc_np == 1) {
            const Particle& a = pv[f.particle_index1];
            f.w1 = a.w;
            for (int d = 0; d < 3; d++) { f.v1[d] = a.v[d]; f.x1[d] = a.x[d]; }
        }
    }
    const ParticleIndexer& ixc = pia.at(cell, species);
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
    for (int64_t bin_id = 1; bin_id <= oc.Nbins; bin_id++) {
        const int64_t loc_np = oc.bins[bin_id - 1].np;
        const OctreeFullCell& f = oc.full_bins[bin_id - 1];
        if (loc_np >= 2) { put(f.w1, f.v1, f.x1, loc_np > 2); put(f.w2, f.v2, f.x2, loc_np > 2); }
        else if (loc_np == 1) { put(f.w1, f.v1, f.x1, false); }
    }
    const int64_t old_count = pia.at(cell, species).n_local;
    const int64_t n_delete = old_count - curr;
    if (!(cell == pia.n_cells) || (n_delete > pia.at(cell, species).n_group2)) pia.contiguous[species - 1] = 0;
    for (int64_t q = 0; q < n_delete; q++) delete_particle_end(pv, pia, cell, species);
}

// :1060-1066 and :1088-1094
template <class SignSrc>
inline void merge_octree_N2_based(SignSrc&& signs, OctreeN2Merge& oc, ParticleVector& pv, ParticleIndexerArray& pia, int64_t cell,
                                  int64_t species, int64_t target_np, const Grid1DUniform* grid) {
    oc.Nbins = 0;  // clear_octree! :256-258
    resize_octree_buffers(oc, pia.at(cell, species).n_local);
    init_octree(oc, pv, pia, cell, species);
    compute_octree(oc, pv, target_np);
    compute_new_particles(signs, oc, pv, pia, cell, species, grid);
}

// Sign source of the GPU convention: Philox block (bin_id-1) of the (OP_MERGE, timestep, cell) stream; bit d of word 0 set -> +1
// for v_d, bit 3+d -> x_d.  (The reference draws rand(rng, [-1.0, 1.0], 3) from its sequential rng.)
struct PhiloxSigns {
    PhiloxStream base;
    void operator()(int64_t bin_id, double sv[3], double sx[3]) const {
        uint32_t c[4] = {(uint32_t)(bin_id - 1), base.ctr[1], base.ctr[2], base.ctr[3]};
        uint32_t o[4];
        Philox4x32::block(c, base.key, o);
        for (int d = 0; d < 3; d++) {
            sv[d] = ((o[0] >> d) & 1u) ? 1.0 : -1.0;
            sx[d] = ((o[0] >> (3 + d)) & 1u) ? 1.0 : -1.0;
        }
    }
};
template <class R>
struct SeqSigns {  // sequential draws like the reference: 3 for v then 3 for x
    R& rng;
    void operator()(int64_t, double sv[3], double sx[3]) {
        for (int d = 0; d < 3; d++) sv[d] = rng.sign();
        for (int d = 0; d < 3; d++) sx[d] = rng.sign();
    }
};

}  // namespace mbo

// NTC / VHS collisions: ntc! (collision_ntc.jl:338-380 one species, :412-453 two species), ntc_equal_weight!
// (:479-521, :554-595), collide_2particles_vhs! (:223-270, variable weight: the heavier particle is split),
// collide_2particles_vhs_equal_weight! (:294-309), compute_n_coll_* (:173-197), compute_g!/compute_com!
// (collision_utils.jl:374-391), sigma_vhs (collision_cross_sections.jl:179-181), scatter_vhs! (collision_scattering.jl:17-29).
//
// Parallelisation: NTC is sequential within a cell (sigma_g_w_max is raised inside the candidate loop and a particle can
// be picked twice), and at DSMC conditions only a few % of a cell's particles are touched per step, so the kernel runs
// ONE THREAD PER CELL: each thread replays the reference's candidate loop for its cell with its own Philox stream
// (OP_NTC, substream, timestep, cell) -- draw for draw the same sequence the CPU oracle consumes -- and gathers only the
// picked particles.  All cells of the range run concurrently.
//
// Variable-weight splits: the reference appends each split particle at logical position n_total + 1 (particles.jl:426-433);
// processing cells sequentially this packs the new group-2 ranges at the tail in cell order.  On the device every cell gets
// a private window at the tail, sized by its candidate count (an upper bound on its splits, known before the loop from
// the first draw): n_coll pre-pass -> exclusive scan -> collide -> pack the windows (scan of actual split counts).  The
// state after the call is the reference's: same logical positions, same pia.
#include "mb_common.cuh"
#include "mb_scan.cuh"
#include "mb_append.cuh"

namespace mb {

struct NtcArgs {
    SoA p1, p2;            // p2 == p1 for one species
    Indexer* ix1;          // indexer rows of the two species
    Indexer* ix2;
    int64_t* n_total1;     // device n_total of the species
    int64_t* n_total2;
    int64_t cap1, cap2;
    double* sgwm;
    int64_t *n_coll, *n_perf, *n_eqw;
    mb_interaction it;
    int64_t cell_lo, cell_hi;   // 1-based inclusive
    double dt, V, dw_tol;
    uint64_t seed;
    uint32_t timestep, substream;
    int equal_weight;
    int32_t* ncoll32;      // VW: candidate count per cell of the range (pre-pass)
    int64_t* win;          // VW: exclusive scan of ncoll32 (n_range + 1)
    int32_t* nsplit1;      // VW: number of new particles per cell (species 1 / 2)
    int32_t* nsplit2;
    int* flags;
    int single_cell_tail;  // VW: one cell whose existing group 2 ends at n_total (0-D usage)
    int64_t warp_min;      // one species: cells with n_local >= warp_min are collided by k_ntc_warp (0: never)
    int64_t n_cells_all;   // cells of the pia (for the mean population)
    int thin;              // one species: with a mean population >= NTC_THIN_PPC only the first N_SM CTAs work (see ntc_impl)
};

__device__ __forceinline__ int64_t ncoll_of(double dt, double V, double sgwm, int64_t n1, int64_t n2, bool two, double R) {
    // compute_n_coll_single_species (:173-176) / compute_n_coll_two_species (:195-197)
    const double f = two ? dt * (double)n1 * (double)n2 * sgwm / V + R : 0.5 * dt * (double)n1 * (double)(n1 - 1) * sgwm / V + R;
    return (int64_t)floor(f);
}

constexpr int64_t NTC_THIN_PPC = 200;

template <bool TWO>
__global__ void __launch_bounds__(128) k_ntc_prepass(NtcArgs a) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = a.cell_lo + r;
        const int64_t n1 = a.ix1[cell - 1].n_local, n2 = TWO ? a.ix2[cell - 1].n_local : n1;
        PhiloxStream rng(a.seed, OP_NTC, a.substream, a.timestep, (uint32_t)cell);
        int64_t nc = ncoll_of(a.dt, a.V, a.sgwm[cell - 1], n1, n2, TWO, rng.rand());
        if (nc < 0) nc = 0;
        if (nc > 0x7fffffff) nc = 0x7fffffff;
        a.ncoll32[r] = (int32_t)nc;
    }
}

// The windows are sized by the candidate counts -- an upper bound on the splits, far above the accepted ones when the weights are
// widely spread (0-D variable-weight cases test thousands of candidates per cell and accept a few per cent).  If their sum does not
// fit the free capacity they are shrunk proportionally instead of failing the call: what the call really needs is room for the
// splits that happen, and a cell whose window overflows raises MB_ERR_CAPACITY on its own (append_split).
template <bool TWO>
static __global__ void k_ntc_fit_windows(NtcArgs a) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const int64_t T = a.win[nr];
    int64_t B = a.cap1 - *a.n_total1;
    if (TWO) { const int64_t B2 = a.cap2 - *a.n_total2; B = B2 < B ? B2 : B; }
    if (B < 0) B = 0;
    if (T <= B) return;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += (int64_t)gridDim.x * blockDim.x)
        a.ncoll32[r] = (int32_t)((double)a.ncoll32[r] * ((double)B / (double)T));  // floor: the sum stays within B
}

// MINB: resident CTAs per SM the register allocation must allow.  Measured on the Couette bench (1.25e8 particles): 4 CTAs/SM
// (118 registers) 1.04 ms, 6 CTAs (80 registers) 1.16 ms, 8 CTAs (64 registers, all cells resident in one wave) 1.21 ms -- the
// kernel is bound by the DRAM random-access rate (one 64 B atom per picked field), not by latency, so more resident cells only
// add spills.  MB_NTC_MINB selects the variant for experiments.
template <bool TWO, int MINB>
__global__ void __launch_bounds__(128, MINB) k_ntc(NtcArgs a) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const bool vw = !a.equal_weight;
    int64_t nt1 = 0, nt2 = 0;
    if (vw) {
        nt1 = *a.n_total1;
        nt2 = TWO ? *a.n_total2 : nt1;
        const int64_t wtot = a.win[nr];
        if (nt1 + wtot > a.cap1 || (TWO && nt2 + wtot > a.cap2)) {  // uniform: nobody collides, the host reports MB_ERR_CAPACITY
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                atomicOr(&a.flags[0], DEVERR_CAPACITY);
                const int64_t need = (nt1 > nt2 ? nt1 : nt2) + wtot;
                a.flags[1] = need > 0x7fffffff ? 0x7fffffff : (int)need;
            }
            return;
        }
    }
    const mb_interaction it = a.it;
    const double pw = 1.0 - 2 * it.vhs_o;
    // resident threads per SM (see ntc_impl): large cells -> one CTA per SM, decided here from the exact particle count
    int64_t nblocks = gridDim.x;
    if (!TWO && a.thin && nblocks > N_SM && *a.n_total1 / a.n_cells_all >= NTC_THIN_PPC) nblocks = N_SM;
    if ((int64_t)blockIdx.x >= nblocks) return;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nr; r += nblocks * blockDim.x) {
        const int64_t cell = a.cell_lo + r;
        Indexer q1 = a.ix1[cell - 1];
        if (!TWO && a.warp_min > 0 && q1.n_local >= a.warp_min) continue;  // large cells: k_ntc_warp
        Indexer q2 = TWO ? a.ix2[cell - 1] : q1;
        int64_t win1 = 0, win2 = 0;
        const int64_t g2_before1 = q1.n_group2, g2_before2 = q2.n_group2;
        int64_t wend1 = 0, wend2 = 0;
        if (vw) {
            win1 = nt1 + a.win[r];
            win2 = nt2 + a.win[r];
            wend1 = nt1 + a.win[r + 1];
            wend2 = nt2 + a.win[r + 1];
            // precondition of the reference (new particles go to n_total + 1): an existing group 2 must end at n_total
            bool ok1 = q1.n_group2 == 0 || (a.single_cell_tail && q1.end2 == nt1);
            bool ok2 = !TWO || q2.n_group2 == 0 || (a.single_cell_tail && q2.end2 == nt2);
            if (!(ok1 && ok2)) { atomicOr(&a.flags[0], DEVERR_PRECONDITION); a.nsplit1[r] = 0; if (TWO) a.nsplit2[r] = 0; continue; }
        }
        PhiloxStream rng(a.seed, OP_NTC, a.substream, a.timestep, (uint32_t)cell);
        double sgwm = a.sgwm[cell - 1];
        const int64_t n_coll = ncoll_of(a.dt, a.V, sgwm, q1.n_local, q2.n_local, TWO, rng.rand());
        int64_t n_perf = 0, n_eqw = 0;
        for (int64_t c = 0; c < n_coll; c++) {
            int64_t i = (int64_t)floor(rng.rand() * (double)q1.n_local);
            int64_t k = (int64_t)floor(rng.rand() * (double)(TWO ? q2.n_local : q1.n_local));
            if (!TWO)
                while (i == k) k = (int64_t)floor(rng.rand() * (double)q1.n_local);
            PRef pi, pk;
            load_p(a.p1, map_cont(q1, i), pi);
            load_p(TWO ? a.p2 : a.p1, map_cont(TWO ? q2 : q1, k), pk);
            const double gx = pi.vx - pk.vx, gy = pi.vy - pk.vy, gz = pi.vz - pk.vz;  // compute_g! collision_utils.jl:388-391
            const double g = sqrt(gx * gx + gy * gy + gz * gz);
            if (!(g > EPS)) continue;
            const double sigma = it.vhs_factor * pow(g, pw);  // sigma_vhs
            const double sgw = sigma * g * fmax(pi.w, pk.w);
            sgwm = fmax(sgw, sgwm);
            if (rng.rand() < sgw / sgwm) {
                n_perf += 1;
                const double cx = it.mu1 * pi.vx + it.mu2 * pk.vx, cy = it.mu1 * pi.vy + it.mu2 * pk.vy,
                             cz = it.mu1 * pi.vz + it.mu2 * pk.vz;  // compute_com!
                if (!vw) {
                    n_eqw += 1;
                } else if (fabs(pi.w - pk.w) < a.dw_tol) {
                    n_eqw += 1;
                } else if (pi.w > pk.w) {
                    if (append_split(a.p1, q1, win1, pi.pos, pi.w - pk.w, pi.vx, pi.vy, pi.vz, wend1, a.flags)) a.p1.a[F_W][pi.pos] = pk.w;
                    if (!TWO) q2 = q1;
                } else {
                    if (TWO) {
                        if (append_split(a.p2, q2, win2, pk.pos, pk.w - pi.w, pk.vx, pk.vy, pk.vz, wend2, a.flags)) a.p2.a[F_W][pk.pos] = pi.w;
                    } else {
                        if (append_split(a.p1, q1, win1, pk.pos, pk.w - pi.w, pk.vx, pk.vy, pk.vz, wend1, a.flags)) a.p1.a[F_W][pk.pos] = pi.w;
                        q2 = q1;
                    }
                }
                // scatter_vhs! collision_scattering.jl:17-29
                const double phi = twopi * rng.rand();
                double sphi, cphi;
                sincos(phi, &sphi, &cphi);
                const double ctheta = 2.0 * rng.rand() - 1.0;
                const double stheta = sqrt(1.0 - ctheta * ctheta);
                const double nx = g * (stheta * cphi), ny = g * (stheta * sphi), nz = g * ctheta;
                a.p1.a[F_VX][pi.pos] = cx + it.mu2 * nx;
                a.p1.a[F_VY][pi.pos] = cy + it.mu2 * ny;
                a.p1.a[F_VZ][pi.pos] = cz + it.mu2 * nz;
                const SoA& sk = TWO ? a.p2 : a.p1;
                sk.a[F_VX][pk.pos] = cx - it.mu1 * nx;
                sk.a[F_VY][pk.pos] = cy - it.mu1 * ny;
                sk.a[F_VZ][pk.pos] = cz - it.mu1 * nz;
            }
        }
        a.sgwm[cell - 1] = sgwm;
        a.n_coll[cell - 1] = n_coll;
        a.n_perf[cell - 1] = n_perf;
        a.n_eqw[cell - 1] = n_eqw;
        if (vw) {
            a.ix1[cell - 1] = q1;
            a.nsplit1[r] = (int32_t)(q1.n_group2 - g2_before1);
            if (TWO) { a.ix2[cell - 1] = q2; a.nsplit2[r] = (int32_t)(q2.n_group2 - g2_before2); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Large cells (0-D ensembles: 1e3 - 1e5 particles per cell, thousands of candidates per step, few accepted when the weights span
// orders of magnitude): ONE WARP per cell, still the reference's exact sequential candidate loop.  The 32 lanes evaluate the next
// 32 candidates speculatively -- lane L assumes that the candidates before it in the batch were plain rejections (3 draws each,
// nothing modified), so its draws are draws d + 3 L .. d + 3 L + 2 of the cell's counter-based stream.  A candidate is an EVENT if
// it is anything else: i == k (a retry draw), g <= eps, sigma g w above the running maximum, or accepted.  The lanes before the
// first event were indeed plain rejections; the event lane then runs its candidate for real (retry draws, maximum update, split,
// scattering) from its stream position, and the next batch starts after it.  Results are identical, draw for draw, to the one-thread
// loop (and to the CPU oracle); the gathers of a batch are in flight together instead of one pair at a time.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double stream_draw(uint32_t k0, uint32_t k1, uint32_t c1, uint32_t c2, uint32_t c3, int64_t d) {
    uint32_t o[4];
    philox4x32_10((uint32_t)(d >> 1), c1, c2, c3, k0, k1, o);
    return (d & 1) ? u64_to_unit_double(o[2], o[3]) : u64_to_unit_double(o[0], o[1]);
}

#ifndef MB_NTCW_MINB
#define MB_NTCW_MINB 4  // resident CTAs per SM k_ntc_warp is compiled for (register budget: 128; measured on C2: 12.0 ms, 13.8 ms at 160 registers, 17 ms at 80)
#endif
__global__ void __launch_bounds__(128, MB_NTCW_MINB) k_ntc_warp(NtcArgs a, int ch) {
    const int64_t nr = a.cell_hi - a.cell_lo + 1;
    const bool vw = !a.equal_weight;
    const unsigned FULL = 0xffffffffu;
    int64_t nt1 = 0;
    if (vw) {
        nt1 = *a.n_total1;
        if (nt1 + a.win[nr] > a.cap1) return;  // k_ntc reports MB_ERR_CAPACITY
    }
    const mb_interaction it = a.it;
    const double pw = 1.0 - 2 * it.vhs_o;
    const int lane = threadIdx.x & 31;
    const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), c3 = (OP_NTC & 0xFFu) | (a.substream << 8);
    // a warp takes ch (<= 32) consecutive cells at a time: one read of their sizes, then the large ones
    for (int64_t r0 = gw * ch; r0 < nr; r0 += nwarps * ch) {
      const int64_t myr = r0 + lane;
      const int64_t my_n = (lane < ch && myr < nr) ? a.ix1[a.cell_lo - 1 + myr].n_local : 0;
      unsigned todo = __ballot_sync(FULL, my_n >= a.warp_min);
      while (todo) {
        const int64_t r = r0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t cell = a.cell_lo + r;
        Indexer q = a.ix1[cell - 1];
        const int64_t g2_before = q.n_group2;
        int64_t win1 = 0, wend1 = 0;
        if (vw) {
            win1 = nt1 + a.win[r];
            wend1 = nt1 + a.win[r + 1];
            if (!(q.n_group2 == 0 || (a.single_cell_tail && q.end2 == nt1))) {  // same precondition as k_ntc
                if (lane == 0) { atomicOr(&a.flags[0], DEVERR_PRECONDITION); a.nsplit1[r] = 0; }
                continue;
            }
        }
        double sgwm = a.sgwm[cell - 1];
        const int64_t n_coll = ncoll_of(a.dt, a.V, sgwm, q.n_local, q.n_local, false, stream_draw(k0, k1, (uint32_t)cell, a.timestep, c3, 0));
        int64_t d = 1, c = 0, n_perf = 0, n_eqw = 0;
        while (c < n_coll) {
            const bool active = c + lane < n_coll;
            bool ev = false;
            if (active) {
                const int64_t p0 = d + 3 * lane;
                const double Nd = (double)q.n_local;
                const int64_t i = (int64_t)floor(stream_draw(k0, k1, (uint32_t)cell, a.timestep, c3, p0) * Nd);
                const int64_t k = (int64_t)floor(stream_draw(k0, k1, (uint32_t)cell, a.timestep, c3, p0 + 1) * Nd);
                if (i == k) {
                    ev = true;
                } else {
                    PRef pi, pk;
                    load_p(a.p1, map_cont(q, i), pi);
                    load_p(a.p1, map_cont(q, k), pk);
                    const double gx = pi.vx - pk.vx, gy = pi.vy - pk.vy, gz = pi.vz - pk.vz;
                    const double g = sqrt(gx * gx + gy * gy + gz * gz);
                    if (!(g > EPS)) {
                        ev = true;
                    } else {
                        const double sigma = it.vhs_factor * pow(g, pw);
                        const double sgw = sigma * g * fmax(pi.w, pk.w);
                        ev = sgw > sgwm || stream_draw(k0, k1, (uint32_t)cell, a.timestep, c3, p0 + 2) < sgw / sgwm;
                    }
                }
            }
            const unsigned m = __ballot_sync(FULL, ev);
            if (m == 0) {
                const int64_t nact = n_coll - c < 32 ? n_coll - c : 32;
                c += nact;
                d += 3 * nact;
                continue;
            }
            const int al = __ffs(m) - 1;
            if (lane == al) {  // the reference's candidate, for real, from stream position d + 3 al
                const int64_t p0 = d + 3 * al;
                PhiloxStream rng(a.seed, OP_NTC, a.substream, a.timestep, (uint32_t)cell);
                rng.c0 = (uint32_t)(p0 >> 1);
                if (p0 & 1) (void)rng.rand();
                int64_t i = (int64_t)floor(rng.rand() * (double)q.n_local);
                int64_t k = (int64_t)floor(rng.rand() * (double)q.n_local);
                while (i == k) k = (int64_t)floor(rng.rand() * (double)q.n_local);
                PRef pi, pk;
                load_p(a.p1, map_cont(q, i), pi);
                load_p(a.p1, map_cont(q, k), pk);
                const double gx = pi.vx - pk.vx, gy = pi.vy - pk.vy, gz = pi.vz - pk.vz;
                const double g = sqrt(gx * gx + gy * gy + gz * gz);
                if (g > EPS) {
                    const double sigma = it.vhs_factor * pow(g, pw);
                    const double sgw = sigma * g * fmax(pi.w, pk.w);
                    sgwm = fmax(sgw, sgwm);
                    if (rng.rand() < sgw / sgwm) {
                        n_perf += 1;
                        const double cx = it.mu1 * pi.vx + it.mu2 * pk.vx, cy = it.mu1 * pi.vy + it.mu2 * pk.vy,
                                     cz = it.mu1 * pi.vz + it.mu2 * pk.vz;
                        if (!vw) {
                            n_eqw += 1;
                        } else if (fabs(pi.w - pk.w) < a.dw_tol) {
                            n_eqw += 1;
                        } else if (pi.w > pk.w) {
                            if (append_split(a.p1, q, win1, pi.pos, pi.w - pk.w, pi.vx, pi.vy, pi.vz, wend1, a.flags)) a.p1.a[F_W][pi.pos] = pk.w;
                        } else {
                            if (append_split(a.p1, q, win1, pk.pos, pk.w - pi.w, pk.vx, pk.vy, pk.vz, wend1, a.flags)) a.p1.a[F_W][pk.pos] = pi.w;
                        }
                        const double phi = twopi * rng.rand();
                        double sphi, cphi;
                        sincos(phi, &sphi, &cphi);
                        const double ctheta = 2.0 * rng.rand() - 1.0;
                        const double stheta = sqrt(1.0 - ctheta * ctheta);
                        const double nx = g * (stheta * cphi), ny = g * (stheta * sphi), nz = g * ctheta;
                        a.p1.a[F_VX][pi.pos] = cx + it.mu2 * nx;
                        a.p1.a[F_VY][pi.pos] = cy + it.mu2 * ny;
                        a.p1.a[F_VZ][pi.pos] = cz + it.mu2 * nz;
                        a.p1.a[F_VX][pk.pos] = cx - it.mu1 * nx;
                        a.p1.a[F_VY][pk.pos] = cy - it.mu1 * ny;
                        a.p1.a[F_VZ][pk.pos] = cz - it.mu1 * nz;
                    }
                }
                d = 2 * (int64_t)rng.c0 - rng.have;  // draws consumed so far
            }
            // everybody continues from the event lane's state
            d = __shfl_sync(FULL, d, al);
            sgwm = __shfl_sync(FULL, sgwm, al);
            n_perf = __shfl_sync(FULL, n_perf, al);
            n_eqw = __shfl_sync(FULL, n_eqw, al);
            q.n_local = __shfl_sync(FULL, q.n_local, al);
            q.n_group2 = __shfl_sync(FULL, q.n_group2, al);
            q.start2 = __shfl_sync(FULL, q.start2, al);
            q.end2 = __shfl_sync(FULL, q.end2, al);
            c += al + 1;
            __syncwarp();  // the event lane's stores are ordered before the next batch's gathers
        }
        if (lane == 0) {
            a.sgwm[cell - 1] = sgwm;
            a.n_coll[cell - 1] = n_coll;
            a.n_perf[cell - 1] = n_perf;
            a.n_eqw[cell - 1] = n_eqw;
            if (vw) {
                a.ix1[cell - 1] = q;
                a.nsplit1[r] = (int32_t)(q.n_group2 - g2_before);
            }
        }
      }
    }
}

static int ntc_impl(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv1, mb_pv* pv2, mb_pia* pia, int64_t cell_lo, int64_t cell_hi,
                    int64_t s1, int64_t s2, double dt, double V, double dw_tol, int equal_weight, uint32_t timestep, uint32_t substream,
                    bool two) {
    MB_ARG(ctx && cf && it && pv1 && pv2 && pia, "NULL handle");
    MB_ARG(s1 >= 1 && s1 <= pia->n_species && s2 >= 1 && s2 <= pia->n_species, "species out of range");
    MB_ARG(cell_lo >= 1 && cell_hi <= pia->n_cells && cell_lo <= cell_hi, "cell range");
    MB_ARG(cf->n_cells == pia->n_cells, "cf.n_cells != pia.n_cells");
    MB_ARG(V > 0.0, "V must be > 0");
    MB_ARG(!two || s1 != s2, "two-species ntc needs two different species");
    MB_CUDA(cudaSetDevice(ctx->device));
    const int64_t nc = pia->n_cells, nr = cell_hi - cell_lo + 1;
    NtcArgs a;
    a.p1 = pv1->cur; a.p2 = pv2->cur;
    a.ix1 = pia->d_indexer + (s1 - 1) * nc;
    a.ix2 = pia->d_indexer + (s2 - 1) * nc;
    a.n_total1 = pia->d_n_total + (s1 - 1);
    a.n_total2 = pia->d_n_total + (s2 - 1);
    a.cap1 = pv1->cap; a.cap2 = pv2->cap;
    a.sgwm = cf->sigma_g_w_max;
    a.n_coll = cf->n_coll; a.n_perf = cf->n_coll_performed; a.n_eqw = cf->n_eq_w;
    a.it = *it;
    a.cell_lo = cell_lo; a.cell_hi = cell_hi;
    a.dt = dt; a.V = V; a.dw_tol = dw_tol;
    a.seed = stream_seed(ctx); a.timestep = timestep; a.substream = stream_substream(substream, s1, s2);
    a.equal_weight = equal_weight;
    a.flags = ctx->d_flags;
    a.single_cell_tail = (nr == 1);
    a.ncoll32 = nullptr; a.win = nullptr; a.nsplit1 = nullptr; a.nsplit2 = nullptr;
    cudaStream_t st = ctx->stream;
    static const int env_grid = getenv("MB_NTC_GRID_PER_SM") ? atoi(getenv("MB_NTC_GRID_PER_SM")) : 0;  // experiment knob: CTAs per SM in the grid
    static const int env_block = getenv("MB_NTC_BLOCK") ? atoi(getenv("MB_NTC_BLOCK")) : 0;  // experiment knob: threads per CTA of k_ntc
    const int nblk = env_block > 0 ? env_block : 128;
    // Resident threads per SM.  The kernel is bound by the DRAM random-access rate, and past the point where the memory system is saturated
    // more concurrent gathers only thrash it: measured on the Couette step at 1.25e8 particles (ppc = 1000 / 250), 128 threads per SM
    // take 0.90 / 0.81 ms, 256: 0.97 / -, 512 (what the registers allow): 1.05 / 0.85, 64: 1.20 / 1.29.  Small cells (C4: ~110 particles,
    // where a thread spends more of its time on the per-cell bookkeeping than on gathers) want all the threads they can get:
    // 4.06 ms at 512 per SM, 4.20 / 4.30 / 6.08 at 384 / 256 / 128.  The kernel decides from the exact mean population (device n_total).
    const int g = grid_for(nr, nblk, env_grid > 0 ? env_grid : 16);
    a.n_cells_all = nc;
    a.thin = env_grid > 0 ? 0 : 1;
    ProfScope ps(ctx, PROF_NTC);
    ctx->state_gen++;
    a.warp_min = two ? 0 : 2048;
    // k_ntc_warp: a warp takes ch cells at a time -- 32 when the range is long enough to keep every warp busy, fewer otherwise
    int ch = 32;
    while (ch > 1 && nr < (int64_t)N_SM * 8 * 4 * ch) ch >>= 1;
    static const int env_wgrid = getenv("MB_NTCW_PER_SM") ? atoi(getenv("MB_NTCW_PER_SM")) : 0;  // experiment knob: CTAs per SM of k_ntc_warp
    const int gwarp = grid_for((nr + ch - 1) / ch * 32, 128, env_wgrid > 0 ? env_wgrid : 8);
    static int minb = -1;
    if (minb < 0) {
        const char* e = getenv("MB_NTC_MINB");
        minb = e ? atoi(e) : 4;
    }
    auto launch_one = [&]() {
        if (minb >= 8) k_ntc<false, 8><<<g, nblk, 0, st>>>(a);
        else if (minb >= 6) k_ntc<false, 6><<<g, nblk, 0, st>>>(a);
        else k_ntc<false, 4><<<g, nblk, 0, st>>>(a);
    };
    if (equal_weight) {
        if (two) k_ntc<true, 4><<<g, 128, 0, st>>>(a);
        else launch_one();
        MB_LAUNCH_CHECK(ctx);
        if (!two) {
            k_ntc_warp<<<gwarp, 128, 0, st>>>(a, ch);
            MB_LAUNCH_CHECK(ctx);
        }
        return MB_OK;
    }
    // variable weight
    int r = pv_ensure_alt(pv1);
    if (r) return r;
    if (two) { r = pv_ensure_alt(pv2); if (r) return r; }
    int32_t* p32 = (int32_t*)ctx_scratch(ctx, 4, (size_t)(3 * nr) * 4);
    int64_t* p64 = (int64_t*)ctx_scratch(ctx, 5, ((size_t)3 * (nr + 1) + gs_partial_count(nr)) * 8);
    if (!p32 || !p64) return MB_ERR_CUDA;
    a.ncoll32 = p32; a.nsplit1 = p32 + nr; a.nsplit2 = p32 + 2 * nr;
    a.win = p64;
    int64_t* packed1 = p64 + (nr + 1);
    int64_t* packed2 = p64 + 2 * (nr + 1);
    int64_t* partial = p64 + 3 * (nr + 1);
    if (two) k_ntc_prepass<true><<<g, 128, 0, st>>>(a);
    else k_ntc_prepass<false><<<g, 128, 0, st>>>(a);
    MB_LAUNCH_CHECK(ctx);
    r = device_exclusive_scan(ctx, a.ncoll32, nr, a.win, partial);
    if (r) return r;
    if (two) k_ntc_fit_windows<true><<<g, 128, 0, st>>>(a);
    else k_ntc_fit_windows<false><<<g, 128, 0, st>>>(a);
    MB_LAUNCH_CHECK(ctx);
    r = device_exclusive_scan(ctx, a.ncoll32, nr, a.win, partial);  // unchanged unless the windows were shrunk
    if (r) return r;
    MB_CUDA(cudaMemsetAsync(a.nsplit1, 0, (size_t)(2 * nr) * 4, st));
    if (two) k_ntc<true, 4><<<g, 128, 0, st>>>(a);
    else launch_one();
    MB_LAUNCH_CHECK(ctx);
    if (!two) {
        k_ntc_warp<<<gwarp, 128, 0, st>>>(a, ch);
        MB_LAUNCH_CHECK(ctx);
    }
    for (int sp = 0; sp < (two ? 2 : 1); sp++) {
        mb_pv* pv = sp == 0 ? pv1 : pv2;
        r = pack_windows(ctx, pv, sp == 0 ? a.ix1 : a.ix2, cell_lo, nr, a.win, sp == 0 ? a.nsplit1 : a.nsplit2, sp == 0 ? packed1 : packed2, partial,
                         sp == 0 ? a.n_total1 : a.n_total2);
        if (r) return r;
        const int64_t sidx = (sp == 0 ? s1 : s2) - 1;
        pia->sorted_layout[sidx] = 0;
        pia->n_bound[sidx] = pv->cap;
    }
    pia->h_valid = false;
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_ntc(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t species, double dt,
           double V, double dw_tol, int32_t equal_weight, uint32_t timestep, uint32_t substream) {
    return ntc_impl(ctx, cf, it, pv, pv, pia, cell_lo, cell_hi, species, species, dt, V, dw_tol, equal_weight, timestep, substream, false);
}
int mb_ntc2(mb_ctx* ctx, mb_cf* cf, const mb_interaction* it, mb_pv* pv1, mb_pv* pv2, mb_pia* pia, int64_t cell_lo, int64_t cell_hi, int64_t s1,
            int64_t s2, double dt, double V, double dw_tol, int32_t equal_weight, uint32_t timestep, uint32_t substream) {
    return ntc_impl(ctx, cf, it, pv1, pv2, pia, cell_lo, cell_hi, s1, s2, dt, V, dw_tol, equal_weight, timestep, substream, true);
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY -- CPU baseline driver built on the oracle (see mb_oracle.hpp header for the rules).
// Restates the multithreaded Couette time loop of /root/reference/simulations/1D/couette_multithreaded.jl:97-173
// (per-chunk ParticleVector / ParticleIndexerArray / rng / GridSortInPlace, ntc -> convect -> reset -> sort per chunk,
// pairwise exchange over a 1-factorisation, sort_particles_after_exchange + compute_props_sorted per chunk) with OpenMP
// threads in place of Julia threads.  It is "a C++ restatement of the reference's multithreaded path", NOT Julia.
//
// usage: couette_cpu nx ppc n_steps n_warmup n_threads [seed=1234] [L=nx*1e-5] [equal_weight=1]
// prints one JSON line: particle-timesteps/s over the timed steps plus the per-section times named as the reference names them.
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "mb_oracle.hpp"
#include "mb_oracle_parallel.hpp"

using namespace mbo;
using clk = std::chrono::steady_clock;

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s nx ppc n_steps n_warmup n_threads [seed] [L] [equal_weight]\n", argv[0]); return 2; }
    const int64_t nx = std::atoll(argv[1]), ppc = std::atoll(argv[2]);
    const int n_steps = std::atoi(argv[3]), n_warm = std::atoi(argv[4]);
    const int n_threads = std::atoi(argv[5]);
    const uint64_t seed = argc > 6 ? std::strtoull(argv[6], nullptr, 10) : 1234;
    const double L = argc > 7 ? std::atof(argv[7]) : (double)nx * 1e-5;
    const bool equal_weight = argc > 8 ? std::atoi(argv[8]) != 0 : true;
    const double T_wall = 300.0, v_wall = 500.0, ndens = 5e22, dt = 2.59e-9;
    const double margin = 1.5;  // preallocation_margin_multiplier, couette_multithreaded.jl:187
    omp_set_num_threads(n_threads);
    const int n_chunks = n_threads;

    std::vector<Species> sd = {Species{66.3e-27, 0.0}};
    const Interaction it = make_interaction(sd[0].mass, sd[0].mass, 4.11e-10, 0.81, 273.0);  // data/vhs.toml "Ar,Ar"
    Grid1DUniform grid(L, nx);
    MaxwellWalls1D walls(sd, T_wall, T_wall, -v_wall, v_wall, 1.0, 1.0);

    std::vector<CellChunk> chunks(n_chunks);
    {
        int64_t base = nx / n_chunks, rem = nx % n_chunks, lo = 1;
        for (int c = 0; c < n_chunks; c++) { int64_t ln = base + (c < rem ? 1 : 0); chunks[c] = CellChunk{lo, lo + ln - 1}; lo += ln; }
    }
    std::vector<std::unique_ptr<ParticleVector>> pv(n_chunks);
    std::vector<std::unique_ptr<ParticleIndexerArray>> pia(n_chunks);
    std::vector<std::unique_ptr<GridSortInPlace>> gs(n_chunks);
    std::vector<Xoshiro256pp> rng;
    std::vector<std::vector<CollisionFactors>> cf(n_chunks);
    for (int c = 0; c < n_chunks; c++) {
        const int64_t np = (int64_t)std::floor((double)(ppc * (chunks[c].hi - chunks[c].lo + 1)) * margin);
        pv[c] = std::make_unique<ParticleVector>(np);
        pia[c] = std::make_unique<ParticleIndexerArray>(nx, 1);
        gs[c] = std::make_unique<GridSortInPlace>(nx, np);
        rng.emplace_back(seed + c);
    }
    ChunkExchanger ex(n_chunks, nx);
    const double Fnum = grid.cell_V(1) * ndens / (double)ppc;
#pragma omp parallel for schedule(static, 1)
    for (int c = 0; c < n_chunks; c++)
        sample_particles_equal_weight_grid(rng[c], grid, *pv[c], *pia[c], 1, sd[0].mass, ndens, T_wall, Fnum, chunks[c].lo, chunks[c].hi);
    const double sgwm0 = estimate_sigma_g_w_max(it, sd[0], sd[0], T_wall, T_wall, Fnum);
    for (int c = 0; c < n_chunks; c++) { cf[c].resize(nx); for (auto& f : cf[c]) f.sigma_g_w_max = sgwm0; }
    PhysProps props(nx, 1);
    auto fact = generate_1_factorization(n_chunks);
    std::vector<ParticleVector*> pvp(n_chunks);
    std::vector<ParticleIndexerArray*> piap(n_chunks);
    for (int c = 0; c < n_chunks; c++) { pvp[c] = pv[c].get(); piap[c] = pia[c].get(); }

    double t_ccs = 0, t_ex = 0, t_rp = 0;
    int64_t particle_steps = 0;
    auto T0 = clk::now();
    for (int t = 1; t <= n_warm + n_steps; t++) {
        if (t == n_warm + 1) { t_ccs = t_ex = t_rp = 0; particle_steps = 0; T0 = clk::now(); }
        auto a = clk::now();
#pragma omp parallel for schedule(static, 1)
        for (int c = 0; c < n_chunks; c++) {
            CollisionData cd;
            for (int64_t cell = chunks[c].lo; cell <= chunks[c].hi; cell++)
                ntc(rng[c], cf[c][cell - 1], cd, it, *pv[c], *pia[c], cell, 1, dt, grid.cell_V(cell), 1e-16, equal_weight);
            convect_particles([&](int64_t) -> Xoshiro256pp& { return rng[c]; }, grid, walls, *pv[c], *pia[c], 1, sd[0].mass, nullptr, dt, false);
            reset_exchanger(ex, c + 1);
            sort_particles(*gs[c], grid, *pv[c], *pia[c], 1);
        }
        auto b = clk::now();
        for (auto& round : fact) {
#pragma omp parallel for schedule(static, 1)
            for (size_t q = 0; q < round.size(); q++) exchange_particles(ex, pvp, piap, chunks, 1, round[q].first, round[q].second);
        }
        auto c2 = clk::now();
#pragma omp parallel for schedule(static, 1)
        for (int c = 0; c < n_chunks; c++) {
            sort_particles_after_exchange(ex, *gs[c], *pv[c], *pia[c], chunks[c], 1);
            std::vector<ParticleVector*> one = {pv[c].get()};
            compute_props_sorted(one, *pia[c], sd, props, chunks[c].lo, chunks[c].hi, nullptr);
        }
        auto d = clk::now();
        t_ccs += std::chrono::duration<double>(b - a).count();
        t_ex += std::chrono::duration<double>(c2 - b).count();
        t_rp += std::chrono::duration<double>(d - c2).count();
        for (int c = 0; c < n_chunks; c++) particle_steps += pia[c]->n_total[0];
    }
    const double total = std::chrono::duration<double>(clk::now() - T0).count();
    double Tsum = 0;
    for (int64_t c = 0; c < nx; c++) Tsum += props.T[c];
    std::printf("{\"particle_steps_per_s\": %.6e, \"particle_steps\": %lld, \"seconds\": %.6f, \"threads\": %d, \"nx\": %lld, \"ppc\": %lld, "
                "\"steps\": %d, \"collide_convect_sort_s\": %.6f, \"exchange_s\": %.6f, \"resort_props_s\": %.6f, \"mean_T\": %.4f}\n",
                (double)particle_steps / total, (long long)particle_steps, total, n_threads, (long long)nx, (long long)ppc, n_steps, t_ccs, t_ex, t_rp,
                Tsum / (double)nx);
    return 0;
}

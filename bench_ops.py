#!/usr/bin/env python
"""bench_ops.py -- per-operator device timings of the hot path at BASELINE.json's single-GPU shapes (C2 / C4 / C5), next to the
algorithmic bytes of SURVEY.md 8(d).  Not the driver's bench (that is bench.py, config C3); prints one JSON line per operator.

    python bench_ops.py [--particles 1e8] [--reps 5]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "merzbild.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np

import merzbild_b200 as mb

AR, K_B, DX, NDENS, DT = 66.3e-27, 1.380649e-23, 1e-5, 5e22, 2.59e-9



def population(n_cells, ppc, seed, vw=False):
    rng = np.random.default_rng(seed)
    n = n_cells * ppc
    sig = math.sqrt(K_B * 300.0 / AR)
    a = [np.empty(n) for _ in range(7)]
    a[0][:] = DX * NDENS / ppc
    if vw:
        a[0] *= rng.uniform(0.5, 1.5, n)
    for f in (1, 2, 3):
        a[f][:] = rng.standard_normal(n) * sig
    a[4][:] = (np.repeat(np.arange(n_cells, dtype=np.float64), ppc) + rng.uniform(0.001, 0.999, n)) * DX
    a[5][:] = 0.5
    a[6][:] = 0.5
    ix = np.zeros((1, n_cells, 7), dtype=np.int64)
    c = np.arange(n_cells, dtype=np.int64)
    ix[0, :, 0] = ppc
    ix[0, :, 1] = c * ppc + 1
    ix[0, :, 2] = (c + 1) * ppc
    ix[0, :, 3] = ppc
    ix[0, :, 5] = -1
    return a, ix, n


def timed(ctx, fn, reps, setup=None):
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        ctx.sync()
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop())
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=float, default=1e8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="run one section: c5 (fp_linear!, props), ops (merge / squash / sort / ntc of the 150-particle cells), c4 (variable-weight Couette loop), c2 (BKW 0-D ensemble)")
    args = ap.parse_args()
    peak = 6533.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    ctx = mb.Context(0, 1234)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)

    def one(fn):
        ctx.sync()
        ctx.timer_start()
        fn()
        return ctx.timer_stop()

    def report(name, cfg, n, ms, bytes_per_particle, note=""):
        if bytes_per_particle is None:  # operators that touch a data-dependent fraction of the particles: no streaming roofline
            print(json.dumps({"op": name, "config": cfg, "particles": n, "ms": ms, "particles_per_s": n / (ms * 1e-3), "note": note}), flush=True)
            return
        gbs = bytes_per_particle * n / (ms * 1e-3) / 1e9
        print(json.dumps({"op": name, "config": cfg, "particles": n, "ms": ms, "particles_per_s": n / (ms * 1e-3),
                          "algorithmic_bytes_per_particle": bytes_per_particle, "achieved_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak, "note": note}),
              flush=True)

    if args.only in ("", "c5"):
        # ---- C5: fp_linear!, 1e6 cells x 100
        ppc = 100
        nc = int(args.particles // ppc)
        a, ix, n = population(nc, ppc, 1)
        pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(nc, 1, ctx)
        pv.upload_soa(1, n, a)
        pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
        step = [0]

        def fp():
            step[0] += 1
            mb.fp_linear(mb.PhiloxRng(step[0]), None, it, AR, pv, pia, (1, nc), 1, DT, DX)

        fp()
        best, med = timed(ctx, fp, args.reps)
        report("fp_linear", "C5: %d cells x %d" % (nc, ppc), n, med, 56, "read w,v 32 B + write v 24 B; 3 normals per particle regenerated from Philox counters")

        # ---- props stand-alone on the same population (two-pass, 32 B/particle)
        pp = mb.PhysProps(nc, 1, ctx=ctx)
        best, med = timed(ctx, lambda: mb.compute_props_sorted([pv], pia, [AR], pp), args.reps)
        report("compute_props_sorted (uncached)", "C5 population", n, med, 32, "two-pass; the second pass re-reads the cell from L1/L2")
        best, med = timed(ctx, lambda: mb.compute_props([pv], pia, [AR], pp), args.reps)
        report("compute_props", "C5 population", n, med, 32, "both pia groups")
        pv.close()
        pia.close()
        del a

    if args.only in ("", "ops"):
        # ---- C2 / C4: variable-weight ntc! + octree merge (150 -> 100) + squash, cells of 150
        ppc = 150
        nc = int(args.particles * 0.6 // ppc)
        a, ix, n = population(nc, ppc, 2, vw=True)
        cap = int(n * 1.3)
        pv, pia = mb.ParticleVector(cap, ctx), mb.ParticleIndexerArray(nc, 1, ctx)
        # wall_offset 1e-6: with L = nc dx ~ 1 m the default offset dx * 1e-12 is below ulp(L), so a merged particle clamped to max_x would sit exactly on L
        grid = mb.Grid1DUniform(nc * DX, nc, wall_offset=1e-6)
        cf = mb.CollisionFactors(nc, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, DX * NDENS / ppc * 1.5), ctx)
        oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)

        pp2 = mb.PhysProps(nc, 1, ctx=ctx)

        def reset():
            pv.upload_soa(1, n, a)
            pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))

        step = [0]

        def merge():
            mb.merge_octree_N2_based(mb.PhiloxRng(1), oc, pv, pia, (1, nc), 1, 100, grid, threshold=130)

        def vw_ntc():
            step[0] += 1
            mb.ntc(mb.PhiloxRng(step[0]), cf, None, it, pv, pia, (1, nc), 1, DT * 4, DX)

        def one(fn):
            ctx.sync()
            ctx.timer_start()
            fn()
            return ctx.timer_stop()

        # one untimed cycle first: lazy allocations (sort ping-pong buffer, scratch arena) must not be inside a timed region
        res = {}
        for rep in range(max(args.reps, 2)):
            reset()
            t_merge = one(merge)
            t_squash = one(lambda: mb.squash_pia(pv, pia, 1))
            n1 = int(pia.n_total[0])
            t_sort = one(lambda: mb.sort_particles(None, grid, pv, pia, 1))
            t_ntc = one(vw_ntc)
            n2 = int(pia.n_total[0])
            t_props = one(lambda: mb.compute_props([pv], pia, [AR], pp2))
            if rep > 0:
                for k, v in (("merge", t_merge), ("squash", t_squash), ("sort", t_sort), ("ntc", t_ntc), ("props", t_props)):
                    res.setdefault(k, []).append(v)
        # velocity-grid merging of the same cells (merge_grid_based!, 3 x 3 x 3 + 8 velocity cells, extents from the cell's PhysProps)
        mgp = mb.GridN2Merge(3, 3, 3, 3.0)
        tg = []
        for rep in range(max(args.reps, 2)):
            reset()
            mb.compute_props([pv], pia, [AR], pp2)
            tg.append(one(lambda: mb.merge_grid_based(mb.PhiloxRng(1), mgp, pv, pia, (1, nc), 1, AR, pp2, grid=grid, threshold=130)))
        ng = int(pia.n_total[0])
        report("merge_grid_based (3x3x3 + 8 velocity cells)", "C4: %d cells x %d -> %d particles" % (nc, ppc, ng), n, sorted(tg[1:])[len(tg[1:]) // 2],
               56 * (n + ng) / n, "CTA per cell, one thread per velocity cell")
        med = {k: sorted(v)[len(v) // 2] for k, v in res.items()}
        report("merge_octree_N2_based (150 -> 100)", "C4: %d cells x %d" % (nc, ppc), n, med["merge"], 56 * (150 + 100) / 150.0,
               "56 (N + N_target) / N bytes per particle of a merged cell")
        report("squash_pia", "after the merge: %d particles" % n1, n1, med["squash"], 112, "payload moves (index indirection is the identity on the device)")
        report("sort_particles (general path)", "after squash", n1, med["sort"], 128, "first sort after a merge: general path")
        report("ntc! variable weight (splits)", "C4 population after merge, dt x 4", n1, med["ntc"], None, "only the picked pairs are gathered: ~%d new particles" % (n2 - n1))
        report("compute_props (both groups)", "C4 population after ntc", n2, med["props"], 32, "group 2 at the tail")
        pv.close()
        pia.close()
        del a

    if args.only in ("", "c4"):
        # ---- C4: 1-D Couette, variable weight, octree merging (couette_varweight_octree.jl:30-135; 500 sampled per cell, merged to
        #      100 at t = 0, threshold 130): per step ntc! -> merge_octree_N2_based! where n_local > 130 -> convect_particles! ->
        #      sort_particles! (squashes first) -> compute_props_sorted!.  Sampled on the device.
        nx = max(int(args.particles // 100), 64)
        ppc_s, thr, tgt = 500, 130, 100
        grid4 = mb.Grid1DUniform(nx * DX, nx, wall_offset=1e-6)
        Fnum = DX * NDENS / ppc_s
        pv, pia = mb.ParticleVector(int(nx * ppc_s * 1.01) + 1024, ctx), mb.ParticleIndexerArray(nx, 1, ctx)
        oc4 = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
        # untimed pass first: the scratch arena (index slices, bin workspace) is allocated by the first merge of this size
        mb.sample_particles_equal_weight(mb.PhiloxRng(0), grid4, pv, pia, 1, AR, float(NDENS), 300.0, Fnum)
        mb.merge_octree_N2_based(mb.PhiloxRng(0), oc4, pv, pia, (1, nx), 1, tgt, grid4, threshold=thr)
        ctx.sync()
        pia.upload(np.tile(np.array([0, 0, -1, 0, 0, -1, 0], dtype=np.int64), (1, nx, 1)), np.array([0]), np.array([1], dtype=np.uint8))
        t_s = one(lambda: mb.sample_particles_equal_weight(mb.PhiloxRng(0), grid4, pv, pia, 1, AR, float(NDENS), 300.0, Fnum))
        n0 = int(pia.n_total[0])
        report("sample_particles_equal_weight! (grid, number density)", "C4: %d cells x ~%d" % (nx, ppc_s), n0, t_s, 60, "write 56 B + cell id per particle")
        t_m = one(lambda: mb.merge_octree_N2_based(mb.PhiloxRng(0), oc4, pv, pia, (1, nx), 1, tgt, grid4, threshold=thr))
        report("merge_octree_N2_based (t = 0: 500 -> 100)", "C4: %d cells" % nx, n0, t_m, 56 * 600 / 500.0, "CTA kernel (cells above 256 particles)")
        mb.squash_pia(pv, pia, 1)
        walls4 = mb.MaxwellWalls1D(300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
        cf4 = mb.CollisionFactors(nx, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, DX * NDENS / tgt), ctx)
        pp4 = mb.PhysProps(nx, 1, ctx=ctx)
        acc, nsteps, n_hist = {}, 40, []
        for t in range(1, nsteps + 1):
            r = mb.PhiloxRng(t)
            tt = {"ntc": one(lambda: mb.ntc(r, cf4, None, it, pv, pia, (1, nx), 1, DT, DX)),
                  "merge": one(lambda: mb.merge_octree_N2_based(r, oc4, pv, pia, (1, nx), 1, tgt, grid4, threshold=thr)),
                  "convect": one(lambda: mb.convect_particles(r, grid4, walls4, pv, pia, 1, AR, DT)),
                  "sort": one(lambda: mb.sort_particles(None, grid4, pv, pia, 1)),
                  "props": one(lambda: mb.compute_props_sorted([pv], pia, [AR], pp4))}
            if t > nsteps // 2:
                n_hist.append(int(pia.n_total[0]))
                for k, v in tt.items():
                    acc.setdefault(k, []).append(v)
        n_mean = sum(n_hist) / len(n_hist)
        tot = 0.0
        # ntc! gathers only the picked pairs and the merge only touches cells above the threshold: no per-particle byte count for them;
        # the sort is the general path with the squash folded in (key 4 + map 4 + perm 4 + record 56 read + 56 written)
        # props: the general path's gather by cell has cached the moments, so compute_props_sorted! moves no particle data
        for k, bpp in (("ntc", None), ("merge", None), ("convect", 44), ("sort", 124), ("props", None)):
            v = acc[k]
            tot += sum(v) / len(v)
            report("C4 step: " + k, "%d cells, ~%.3g particles, mean of steps %d-%d" % (nx, n_mean, nsteps // 2 + 1, nsteps), int(n_mean), sum(v) / len(v), bpp,
                   "max %.2f ms" % max(v))
        d = pp4.download()
        print(json.dumps({"op": "C4 step total", "ms": tot, "particle_steps_per_s": n_mean / (tot * 1e-3), "mean_T_K": float(d["T"].mean()),
                          "mean_np_per_cell": float(d["np"].mean())}), flush=True)
        pv.close()
        pia.close()
    if args.only in ("c4", "c5", "ops"):
        ctx.close()
        return

    # ---- C2: 0-D BKW variable-weight relaxation with octree N:2 merging (bkw_varweight_octree.jl / test_bkw_varweight_octree.jl:43-47):
    #      an ensemble of independent cells, each the nv = 40 grid sample (~33.5k particles) merged to 8000 at t = 0, then per step
    #      ntc! -> merge if n_local > 10000 -> squash_pia! / re-sort by cell id -> compute_props_with_total_moments!.  Sampled on the device.
    itm = mb.make_interaction(AR, AR, 4.11e-10, 1.0, 273.0)  # data/pseudo_maxwell.toml
    T0, n_dens, nv = 273.0, 1e23, 40
    probe_pv, probe_pia = mb.ParticleVector(nv ** 3, ctx), mb.ParticleIndexerArray(1, 1, ctx)
    n_s = mb.sample_on_grid(mb.PhiloxRng(0), "bkw", probe_pv, probe_pia, 1, 1, nv, AR, T0, n_dens)
    probe_pv.close()
    probe_pia.close()
    ncell = max(int(args.particles // n_s), 1)
    n = ncell * n_s
    pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(ncell, 1, ctx)
    oc2 = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
    # untimed pass first (scratch arena), then the population is sampled again
    mb.sample_on_grid(mb.PhiloxRng(0), "bkw", pv, pia, (1, ncell), 1, nv, AR, T0, n_dens)
    mb.merge_octree_N2_based(mb.PhiloxRng(0), oc2, pv, pia, (1, ncell), 1, 8000, threshold=10000)
    ctx.sync()
    pia.upload(np.tile(np.array([0, 0, -1, 0, 0, -1, 0], dtype=np.int64), (1, ncell, 1)), np.array([0]), np.array([1], dtype=np.uint8))
    t_sample = one(lambda: mb.sample_on_grid(mb.PhiloxRng(0), "bkw", pv, pia, (1, ncell), 1, nv, AR, T0, n_dens))
    report("sample_on_grid! (BKW, nv = 40)", "C2: %d cells x %d" % (ncell, n_s), n, t_sample, 60, "write 56 B + cell id 4 B per particle (includes the host weight table)")
    ppm = mb.PhysProps(ncell, 1, (4, 6, 8, 10), Tref=T0, ctx=ctx)
    t_m0 = one(lambda: mb.merge_octree_N2_based(mb.PhiloxRng(0), oc2, pv, pia, (1, ncell), 1, 8000, threshold=10000))
    report("merge_octree_N2_based (%d -> 8000)" % n_s, "C2 initial merge", n, t_m0, 56 * (n_s + 8000) / n_s, "CTA per cell")
    mb.sort_particles(None, pv, pia, 1)
    n1 = int(pia.n_total[0])
    Fnum = n_dens / 8000.0
    cf2 = mb.CollisionFactors(ncell, mb.estimate_sigma_g_w_max(itm, AR, AR, T0, T0, Fnum), ctx)
    sigma_ref = math.pi * 4.11e-10 ** 2
    tref = 1.0 / (n_dens * sigma_ref) / math.sqrt(2 * K_B * T0 / AR)
    dt2 = 0.025 * tref
    acc = {}
    n_steps = 12
    for t in range(1, n_steps + 1):
        tt = {"ntc": one(lambda: mb.ntc(mb.PhiloxRng(t), cf2, None, itm, pv, pia, (1, ncell), 1, dt2, 1.0)),
              "merge": one(lambda: mb.merge_octree_N2_based(mb.PhiloxRng(t), oc2, pv, pia, (1, ncell), 1, 8000, threshold=10000)),
              # an ensemble of 0-D cells shares one ParticleVector: the split particles appended at the tail are folded back into
              # their cells by the cell-id sort (which squashes first when a merge left holes)
              "sort": one(lambda: mb.sort_particles(None, pv, pia, 1)),
              "props": one(lambda: mb.compute_props_with_total_moments([pv], pia, [AR], ppm))}
        if t > 2:
            for k, v in tt.items():
                acc.setdefault(k, []).append(v)
    n2 = int(pia.n_total[0])
    d = ppm.download()
    tot = sum(sum(v) / len(v) for v in acc.values())
    for k, bpp in (("ntc", None), ("merge", None), ("sort", 124), ("props", 32 * 6)):
        v = acc[k]
        report("C2 step: " + k, "%d cells, %d..%d particles, mean of steps 3-%d (max %.2f ms)" % (ncell, n1, n2, n_steps, max(v)), n2, sum(v) / len(v), bpp,
               "T = %.2f K (T0 %.0f), M4 = %.4f" % (d["T"].mean(), T0, d["moments"][0, :, 0].mean()))
    print(json.dumps({"op": "C2 step total", "ms": tot, "particle_steps_per_s": n2 / (tot * 1e-3)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

"""Worker of tests/test_gpu_multirank.py: launched with torch.distributed.run, one rank per GPU.
Checks the NCCL slab exchange (mb_exchange_slab) + sort against a single-domain CPU-oracle run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "merzbild.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import merzbild_b200 as mb

AR = 66.3e-27


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mb.Context(local, 1234)
    uid = [mb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    mb.comm_init(ctx, uid[0], rank, world)

    nx, ppc, steps = 40 + 3, 150, 80  # 43 cells: uneven slabs
    L = nx * 1e-5
    n = nx * ppc
    rng = np.random.default_rng(99)
    rows = np.zeros((n, 7))
    rows[:, 0] = 1.0 + np.arange(n)  # unique weights identify the particles
    rows[:, 1:4] = rng.normal(0, 300.0, (n, 3))
    rows[:, 4] = rng.uniform(0, L, n)
    rows[:, 5:7] = rng.uniform(0, 1, (n, 2))
    G = mb.Grid1DUniform(L, nx)
    slab = G.slab(rank, world)
    gcell = np.floor(rows[:, 4] * G.inv_dx).astype(np.int64)
    mine = rows[(gcell >= slab.cell_offset) & (gcell < slab.cell_offset + slab.n_cells)]
    cap = 3 * n
    pv = mb.ParticleVector(cap, ctx)
    pia = mb.ParticleIndexerArray(slab.n_cells, 1, ctx)
    pv.set_logical(1, mine)
    ix = np.zeros((1, slab.n_cells, 7), dtype=np.int64)
    ix[0, :, 2] = -1
    ix[0, :, 5] = -1
    if len(mine):
        ix[0, 0] = (len(mine), 1, len(mine), len(mine), 0, -1, 0)
    pia.upload(ix, np.array([len(mine)]), np.array([1], dtype=np.uint8))
    mb.sort_particles(None, slab, pv, pia, 1)
    walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)  # specular: exact arithmetic, no draws
    dt = 0.3e-5 / 300.0  # ~0.3 cells per step at sigma_v: the sort's band (w = 2) holds, arrivals are merged in the band path
    sent_total = 0
    paths = []
    edge = len(sys.argv) > 1 and sys.argv[1] == "edge"  # edge exchange: no host round trip, counts stay on the device
    for t in range(1, steps + 1):
        mb.convect_particles(mb.PhiloxRng(t), slab, walls, pv, pia, 1, AR, dt)
        if edge:
            mb.exchange_slab(ctx, slab, pv, pia, 1)
            sent_total += 2
        else:
            s, r = mb.exchange_slab(ctx, slab, pv, pia, 1, counts=True)
            sent_total += int(s.sum())
        mb.sort_particles(None, slab, pv, pia, 1)
        paths.append(ctx.sort_last_path)
        ok, where = pia.check(1)
        assert ok, (rank, t, where)
    assert paths.count(1) >= steps - 2, paths  # band path with drops + arrivals
    nt = int(pia.n_total[0])
    loc = pv.logical(1, nt)
    lc = np.floor(loc[:, 4] * G.inv_dx).astype(np.int64) - slab.cell_offset
    assert lc.min() >= 0 and lc.max() < slab.n_cells, "a particle outside the slab survived the sort"
    assert np.all(np.diff(lc) >= 0), "not sorted by cell"
    counts = np.bincount(lc, minlength=slab.n_cells)
    np.testing.assert_array_equal(pia.indexer[0, :, 0], counts)
    gathered = [None] * world
    dist.all_gather_object(gathered, (loc, sent_total))
    if rank == 0:
        from oracle import oracle

        allrows = np.concatenate([g[0] for g in gathered])
        assert allrows.shape[0] == n, "particles lost or duplicated"
        assert sum(g[1] for g in gathered) > 100, "the test did not exchange anything"
        opv, opia = oracle.OPV(n), oracle.OPIA(nx, 1)
        opv.particles[:n] = rows
        opv.nbuffer = 0
        opia.indexer[0, 0] = (n, 1, n, n, 0, -1, 0)
        opia.n_total[0] = n
        oracle.sort_particles(opv, opia, 1, grid=(L, nx))
        for t in range(1, steps + 1):
            oracle.convect_particles(oracle.Rng.philox(1234, t), (L, nx), (300.0, 300.0, 0, 0, 0, 0), opv, opia, 1, [AR], dt)
            oracle.sort_particles(opv, opia, 1, grid=(L, nx))
        ref = opv.logical(1, n)
        a = allrows[np.argsort(allrows[:, 0])]
        b = ref[np.argsort(ref[:, 0])]
        np.testing.assert_array_equal(a, b)  # specular walls: bit-exact trajectories, nobody lost
        # per-cell populations equal the single-domain run
        np.testing.assert_array_equal(np.concatenate([np.bincount(np.floor(g[0][:, 4] * G.inv_dx).astype(np.int64), minlength=nx) for g in gathered]).reshape(world, nx).sum(0),
                                      opia.indexer[0, :, 0])
        print("MULTIRANK_OK exchange+sort bit-exact vs single-domain oracle; exchanged", sum(g[1] for g in gathered), "particles", flush=True)
    dist.barrier()

    # phase 2: the full Couette step (collide -> convect with diffuse walls -> exchange -> sort -> props) conserves the global population
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    cf = mb.CollisionFactors(slab.n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 1e12), ctx)
    walls2 = mb.MaxwellWalls1D(300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
    pp = mb.PhysProps(slab.n_cells, 1, ctx=ctx)
    for t in range(1, 21):
        mb.ntc_equal_weight(mb.PhiloxRng(100 + t, rank), cf, None, it, pv, pia, (1, slab.n_cells), 1, dt, slab.dx)
        mb.convect_particles(mb.PhiloxRng(100 + t, rank), slab, walls2, pv, pia, 1, AR, dt)
        mb.exchange_slab(ctx, slab, pv, pia, 1)
        mb.sort_particles(None, slab, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], pp)
    tot = torch.tensor([float(pia.n_total[0]), float(pp.download()["np"].sum())], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    assert int(tot[0].item()) == n and int(tot[1].item()) == n, tot
    if rank == 0:
        print("MULTIRANK_OK couette step conserves the global population over", world, "ranks", flush=True)
    ctx.sync()

    # phase 3: the variable-weight Couette loop over slabs (C4: couette_multithreaded_varweight_octree.jl): ntc! with splits ->
    # merge_octree_N2_based! above the threshold -> squash_pia! -> convect (specular walls) -> exchange -> sort.  Splits, merges,
    # specular walls and the exchange all conserve weight and kinetic energy, so the GLOBAL sums must stay put
    # (test_couette_varweight_octree_chunking.jl:137 asks 4 eps per step of the density; energy to 1e-11 here).
    nx3, ppc3 = 64, 160
    G3 = mb.Grid1DUniform(nx3 * 1e-5, nx3)
    slab3 = G3.slab(rank, world)
    pv3 = mb.ParticleVector(6 * slab3.n_cells * ppc3, ctx)
    pia3 = mb.ParticleIndexerArray(slab3.n_cells, 1, ctx)
    Fnum3 = 1e-5 * 5e22 / ppc3
    ctx.set_seed(4321 + rank)
    mb.sample_particles_equal_weight(mb.PhiloxRng(0), slab3, pv3, pia3, 1, AR, ppc3, 300.0, Fnum3)
    oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
    mb.merge_octree_N2_based(mb.PhiloxRng(0), oc, pv3, pia3, (1, slab3.n_cells), 1, 100, slab3, threshold=130)
    mb.squash_pia(pv3, pia3, 1)
    cf3 = mb.CollisionFactors(slab3.n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 1e-5 * 5e22 / 100), ctx)
    walls3 = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)

    def totals():
        nt = int(pia3.n_total[0])
        a = pv3.logical(1, nt)
        t = torch.tensor([a[:, 0].sum(), (a[:, 0] * (a[:, 1:4] ** 2).sum(1)).sum(), float(nt)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return t.cpu().numpy()

    t0 = totals()
    merged_cells, sent3 = 0, 0
    dt3 = 2.59e-9 * 8
    for t in range(1, 31):
        r = mb.PhiloxRng(t, rank)
        mb.ntc(r, cf3, None, it, pv3, pia3, (1, slab3.n_cells), 1, dt3, slab3.dx)
        merged_cells += int((pia3.indexer[0, :, 0] > 130).sum())
        mb.merge_octree_N2_based(r, oc, pv3, pia3, (1, slab3.n_cells), 1, 100, slab3, threshold=130)
        if t % 2 == 0:  # the exchange takes the non-contiguous layout a merge leaves as well as the squashed one
            mb.squash_pia(pv3, pia3, 1)
        mb.convect_particles(r, slab3, walls3, pv3, pia3, 1, AR, dt3)
        sc, rc = mb.exchange_slab(ctx, slab3, pv3, pia3, 1, counts=True)
        sent3 += int(sc.sum())
        mb.sort_particles(None, slab3, pv3, pia3, 1)
        ok, where = pia3.check(1)
        assert ok, (rank, t, where)
        if t % 10 == 0:
            t1 = totals()
            assert abs(t1[0] - t0[0]) <= 4e-16 * t * t0[0] * 8, (t, t1[0], t0[0])
            assert abs(t1[1] - t0[1]) <= 1e-11 * t0[1], (t, t1[1], t0[1])
    nt = int(pia3.n_total[0])
    a = pv3.logical(1, nt)
    lc = np.floor(a[:, 4] * G3.inv_dx).astype(np.int64) - slab3.cell_offset
    assert lc.min() >= 0 and lc.max() < slab3.n_cells and np.all(np.diff(lc) >= 0)
    stats = torch.tensor([float(merged_cells), float(sent3)], device="cuda", dtype=torch.float64)
    dist.all_reduce(stats)
    assert stats[0].item() > 20 and stats[1].item() > 100, stats
    if rank == 0:
        print("MULTIRANK_OK variable-weight loop (ntc splits, octree merge, squash, exchange, sort) conserves global weight and energy;",
              int(stats[0].item()), "cell merges,", int(stats[1].item()), "particles exchanged", flush=True)
    ctx.sync()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

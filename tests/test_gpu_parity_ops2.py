"""GPU parity tests, part 2: octree N:2 merging, squash_pia!, SWPM, linear Fokker-Planck -- CUDA path (through the C ABI)
against the CPU oracle.  Merging: identical bin structure (same surviving particle counts per cell, same pia), merged
particles to 1e-12, mass / momentum / energy conserved to 1e-12 relative."""
import numpy as np
import pytest

from parity_util import AR, assert_rows_close, assert_same_pia, maxwellian_rows, mirror_to_device, oracle_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(mb):
    c = mb.Context(0, 1234)
    yield c
    c.close()


def _moments(rows):
    w = rows[:, 0]
    return np.array([w.sum(), *(w[:, None] * rows[:, 1:4]).sum(0), (w[:, None] * rows[:, 1:4] ** 2).sum(0).sum()])


def _octant_particles():
    """test/test_octree_merging.jl:3-40: 24 particles, 3 per velocity octant, weight = octant id."""
    rows = []
    signs = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]
    for i, s in enumerate(signs, start=1):
        for k, dv in enumerate((-1.0, 0.0, 1.0)):
            v = np.array(s, dtype=float) * (9.0 - i) + dv * 0.25 * np.array([1.0, -1.0, 0.5])
            rows.append([float(i), *v, 0.1 * i + 0.01 * k, 0.5, 0.25])
    return np.array(rows)


def test_octree_merge_reference_kat(mb, oracle, ctx):
    """24 particles in 8 octants -> merge to 16 (8 bins x 2) and then to 2; n, v, T conserved (test_octree_merging.jl:70-163)."""
    rows = _octant_particles()
    for target, expect in ((16, 16), (2, 2)):
        opv, opia = oracle_state(oracle, rows, 1)
        pv, pia = mirror_to_device(mb, ctx, opv, opia)
        oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C, oracle.BOUNDS_INHERIT, 4096, 10)
        oracle.merge_octree_N2(oracle.Rng.philox(1234, 7, 0), oc, opv, opia, 1, 1, 1, target)
        moc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinC)
        mb.merge_octree_N2_based(mb.PhiloxRng(7, 0), moc, pv, pia, 1, 1, target)
        assert_same_pia(opia, pia)
        nt = int(opia.n_total[0])
        assert nt == expect
        a, b = pv.logical(1, 24), opv.logical(1, 24)
        assert_rows_close(a[:nt], b[:nt], 1e-13, "merged particles")
        assert np.all(a[nt:, 0] == 0.0)  # deleted particles carry zero weight (particles.jl:506,547)
        m0, m1 = _moments(rows), _moments(a[:nt])
        np.testing.assert_allclose(m1, m0, rtol=1e-14, atol=1e-12)


@pytest.mark.parametrize("init,bounds,split", [(1, 1, 1), (2, 1, 1), (1, 2, 1), (3, 1, 1), (1, 1, 2)])
def test_octree_merge_single_cell_parity(mb, oracle, ctx, init, bounds, split):
    """0-D shape of test_bkw_varweight_octree.jl (scaled down): 6000 variable-weight particles -> 800."""
    rng = np.random.default_rng(17)
    n, target = 6000, 800
    rows = maxwellian_rows(rng, n, 1.0, vw=True, w=1e15)
    rows[:, 1:4] *= rng.uniform(0.5, 1.5, (n, 1))
    opv, opia = oracle_state(oracle, rows, 1)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    oc = oracle.Octree(split, init, bounds, 6000, 10)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 3, 1), oc, opv, opia, 1, 1, 1, target)
    moc = mb.OctreeN2Merge(split, init, bounds, max_Nbins=6000)
    mb.merge_octree_N2_based(mb.PhiloxRng(3, 1), moc, pv, pia, 1, 1, target)
    nt_dev = int(pia.n_total[0])
    a = pv.logical(1, n)
    m0, m1 = _moments(rows), _moments(a[:nt_dev])
    np.testing.assert_allclose(m1, m0, rtol=1e-12)  # mass, momentum, energy
    if init != 3:  # with +-c root bounds 10 levels of halving never resolve thermal velocities: 8 bins, as in the reference
        assert target - 14 <= nt_dev <= target
    assert pia.check(1) == (True, 0)
    if split == 1:  # mid split: the octree is identical, merged particles agree to rounding
        assert_same_pia(opia, pia)
        b = opv.logical(1, n)
        assert_rows_close(a[:nt_dev], b[:nt_dev], 1e-12, "merged particles")
    # squash is a no-op for a single cell whose deletions came from the tail
    mb.squash_pia(pv, pia, 1)
    oracle.squash_pia(opv, opia, 1)
    if split == 1:
        assert_same_pia(opia, pia)


def test_octree_merge_1d_cells_and_squash(mb, oracle, ctx):
    """Couette variable-weight shape (couette_varweight_octree.jl:86-135): ntc! -> merge cells over the threshold (1-D variant
    with the x clamp) -> squash_pia!, then sort; state identical to the oracle's per-cell loop."""
    n_cells, ppc = 30, 400
    rng = np.random.default_rng(23)
    L = n_cells * 1e-5
    n = n_cells * ppc
    Fnum = 1e-5 * 5e22 / ppc
    rows = maxwellian_rows(rng, n, L, w=Fnum, vw=True)
    rows[:, 4] = rng.beta(0.7, 0.7, n) * L  # uneven cells: some over the threshold, some under; crowd the walls
    cap = 3 * n
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=cap)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia, capacity=cap)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 2 * Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    g = mb.Grid1DUniform(L, n_cells)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    moc = mb.OctreeN2Merge(max_Nbins=6000)
    threshold, target = 390, 250
    dt, V = 2.59e-9 * 6, L / n_cells
    for t in range(1, 4):
        before = _moments(opv.logical(1, int(opia.n_total[0])))
        mb.ntc(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, dt, V)
        mb.merge_octree_N2_based(mb.PhiloxRng(t), moc, pv, pia, (1, n_cells), 1, target, grid=g, threshold=threshold)
        oracle.ntc(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, dt, V)
        oracle.merge_octree_N2(oracle.Rng.philox(1234, t), oc, opv, opia, 1, n_cells, 1, target, threshold=threshold, grid=(L, n_cells))
        assert_same_pia(opia, pia)  # incl. contiguous == 0
        assert pia.check(1) == (True, 0)
        mb.squash_pia(pv, pia, 1)
        oracle.squash_pia(opv, opia, 1)
        assert_same_pia(opia, pia)
        nt = int(opia.n_total[0])
        a, b = pv.logical(1, nt), opv.logical(1, nt)
        assert_rows_close(a, b, 1e-11, "after merge + squash")
        np.testing.assert_allclose(_moments(a), before, rtol=1e-12)
        assert a[:, 4].min() >= g.min_x and a[:, 4].max() <= g.max_x
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert_same_pia(opia, pia)
        assert_rows_close(pv.logical(1, nt), opv.logical(1, nt), 1e-11, "after sort")
    assert int(opia.indexer[0, :, 0].max()) <= threshold + 60


def test_sort_squashes_noncontiguous(mb, oracle, ctx):
    """sort_particles! calls squash_pia! first when the species is not contiguous (grid_sorting.jl:69-71)."""
    n_cells, ppc = 8, 100
    rng = np.random.default_rng(29)
    L = n_cells * 1e-5
    rows = maxwellian_rows(rng, n_cells * ppc, L, vw=True)
    opv, opia = oracle_state(oracle, rows, n_cells)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 1), oc, opv, opia, 1, n_cells, 1, 40, threshold=-1, grid=(L, n_cells))
    mb.merge_octree_N2_based(mb.PhiloxRng(1), mb.OctreeN2Merge(max_Nbins=6000), pv, pia, (1, n_cells), 1, 40, grid=mb.Grid1DUniform(L, n_cells))
    assert_same_pia(opia, pia)
    mb.sort_particles(None, mb.Grid1DUniform(L, n_cells), pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    assert_same_pia(opia, pia)
    nt = int(opia.n_total[0])
    assert_rows_close(pv.logical(1, nt), opv.logical(1, nt), 1e-12, "sort after merge")


# ------------------------------------------------------------------------------------------------------------ SWPM
def test_swpm_parity(mb, oracle, ctx):
    """swpm! (collision_swpm.jl:201-287): two children per accepted pair, parents lose dw; vs the oracle, 3 steps, multi-cell."""
    n_cells, ppc = 10, 300
    rng = np.random.default_rng(31)
    L = n_cells * 1e-5
    n = n_cells * ppc
    Fnum = 1e-5 * 5e22 / ppc
    rows = maxwellian_rows(rng, n, L, w=Fnum, vw=True)
    cap = 6 * n
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=cap)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia, capacity=cap)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 1.0)  # sigma_g_max (no weight) for SWPM
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    g = mb.Grid1DUniform(L, n_cells)
    G, dt, V = 0.5, 2.59e-9 * 4, L / n_cells
    m0 = _moments(rows)
    for t in range(1, 4):
        mb.swpm(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, G, dt, V)
        oracle.swpm(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, G, dt, V)
        d = cf.download()
        np.testing.assert_array_equal(d["n_coll"], ocf.n_coll)
        np.testing.assert_array_equal(d["n_coll_performed"], ocf.n_coll_performed)
        assert_same_pia(opia, pia)
        nt = int(opia.n_total[0])
        a, b = pv.logical(1, nt), opv.logical(1, nt)
        assert_rows_close(a, b, 1e-12, "swpm")
        np.testing.assert_allclose(_moments(a), m0, rtol=1e-12)
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert_same_pia(opia, pia)
    assert nt > n


# ------------------------------------------------------------------------------------------------------------ FP
def test_fp_linear_parity(mb, oracle, ctx):
    """fp_linear! (collision_fp.jl:24-125) over cells of 100-200 particles (test_collision_fp.jl / test_1D_couette_fp.jl shapes):
    same Philox normals as the oracle; per-cell momentum and energy conserved exactly; cells with n < 7 untouched."""
    n_cells = 64
    rng = np.random.default_rng(37)
    L = n_cells * 1e-5
    counts = rng.integers(100, 200, n_cells)
    counts[5] = 6
    counts[9] = 0
    n = int(counts.sum())
    Fnum = 2.5e15
    rows = maxwellian_rows(rng, n, L, w=Fnum, vw=False)  # equal weights: sum(xi) = 0 makes the momentum exactly conserved
    cells = np.repeat(np.arange(n_cells), counts)
    rows[:, 4] = (cells + rng.uniform(0.01, 0.99, n)) * 1e-5
    rows[:, 1] *= 1.8  # anisotropic: T_x > T_y, relaxes towards isotropy
    rows[:, 2] += 300.0
    opv, opia = oracle_state(oracle, rows, n_cells)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    dt, V = 2.59e-9 * 50, 1e-5
    start = opv.logical(1, n)
    for t in range(1, 4):
        mb.fp_linear(mb.PhiloxRng(t, 4), None, it, AR, pv, pia, (1, n_cells), 1, dt, V)
        oracle.fp_linear(oracle.Rng.philox(1234, t, 4), oit, AR, opv, opia, 1, n_cells, 1, dt, V)
    a, b = pv.logical(1, n), opv.logical(1, n)
    assert_rows_close(a, b, 1e-11, "fp_linear")
    np.testing.assert_array_equal(a[:, [0, 4, 5, 6]], start[:, [0, 4, 5, 6]])
    off = np.concatenate(([0], np.cumsum(counts)))
    Tx0 = Tx1 = 0.0
    for c in range(n_cells):
        s, e = off[c], off[c + 1]
        if counts[c] < 7:
            np.testing.assert_array_equal(a[s:e], start[s:e])
            continue
        np.testing.assert_allclose(_moments(a[s:e]), _moments(start[s:e]), rtol=1e-12)
        Tx0 += np.var(start[s:e, 1])
        Tx1 += np.var(a[s:e, 1])
    assert Tx1 < Tx0  # the hot component cools


def test_octree_merging_buffer_sorting_reference_kat(mb, oracle, ctx):
    """test_octree_merging_buffer_sorting.jl:43-200 through the C ABI: a cell whose particles sit in two index groups around another
    cell's particles; post-merge counts (target 6 -> 2 particles, target 16 -> 12), the pia with its hole after the merge
    (contiguous == false), and after sort_particles! (squashes first): contiguous again, cell 2 directly behind cell 1."""
    from test_oracle_kat_octree_vhs import _buffer_sorting_state

    for n_gr1, target, expect_after, expect_c2 in ((50, 6, (2, 1, 2, 2), (10, 51, 60, 10)), (5, 16, (12, 1, 5, 5, 16, 22, 7), (10, 6, 15, 10))):
        rows, opv, opia = _buffer_sorting_state(oracle, n_gr1)
        pv, pia = mirror_to_device(mb, ctx, opv, opia)
        pp = mb.PhysProps(2, 1, ctx=ctx)
        mb.compute_props([pv], pia, [AR], pp)
        d = pp.download()
        assert d["np"][0].tolist() == [90.0, 10.0] and d["n"][0].tolist() == [90.0, 10000.0]
        oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX)
        oracle.merge_octree_N2(oracle.Rng.philox(1234, 1), oc, opv, opia, 1, 1, 1, target)
        moc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel)
        mb.merge_octree_N2_based(mb.PhiloxRng(1), moc, pv, pia, 1, 1, target)
        ix, nt, ct = pia.download()
        assert tuple(ix[0, 0][:len(expect_after)]) == expect_after and tuple(ix[0, 1][:4]) == expect_c2 and ct[0] == 0
        assert_same_pia(opia, pia)
        mb.compute_props([pv], pia, [AR], pp)
        d = pp.download()
        assert d["np"][0].tolist() == [float(expect_after[0]), 10.0] and abs(d["n"][0, 0] - 90.0) < 1e-12 and d["n"][0, 1] == 10000.0
        g = mb.Grid1DUniform(8.0, 2)
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(8.0, 2))
        ix, nt, ct = pia.download()
        k = expect_after[0]
        assert tuple(ix[0, 0]) == (k, 1, k, k, 0, -1, 0) and tuple(ix[0, 1]) == (10, k + 1, k + 10, 10, 0, -1, 0) and ct[0] == 1 and nt[0] == k + 10
        assert_rows_close(pv.logical(1, k + 10), opv.logical(1, k + 10), 1e-13, "after merge + sort")


# ------------------------------------------------------------------------------------------------------------ SurfProps / averaging
def test_surf_props_device_handle_avg_and_reduce(mb, oracle, ctx):
    """SurfProps kept on the device (surface_props.jl:22-252): convect_particles!(..., surf_props, dt) clears, accumulates and scales it
    without a host round trip; avg_props!(surf_avg, surf, n) and reduce_surf_props!(target, chunks) run on the device too.  Against the
    oracle's convect_particles! with SurfProps on the same Philox streams, step by step, and against the host-pointer variant."""
    rng = np.random.default_rng(33)
    n, nx, L, dt = 40000, 40, 4e-4, 2.59e-7
    rows = maxwellian_rows(rng, n, L, w=1e10, vw=True)
    opv, opia = oracle_state(oracle, rows, nx)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pv2, pia2 = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, nx)
    walls, owalls = mb.MaxwellWalls1D(300.0, 450.0, -500.0, 500.0, 0.7, 1.0), (300.0, 450.0, -500.0, 500.0, 0.7, 1.0)
    surf, avg = mb.SurfProps(ctx), mb.SurfProps(ctx)
    n_avg = 3
    avg_ref = np.zeros((2, 11))
    for t in range(1, n_avg + 1):
        l0 = ctx.kernel_launches
        out = mb.convect_particles(mb.PhiloxRng(t, 1), g, walls, pv, pia, 1, AR, dt, surf_props=surf)
        assert out is surf and ctx.kernel_launches - l0 >= 2  # convection + scaling kernels, nothing downloaded
        mb.avg_surf_props(avg, surf, n_avg)
        s_host = mb.convect_particles(mb.PhiloxRng(t, 1), g, walls, pv2, pia2, 1, AR, dt, surf_props=True)
        so = oracle.convect_particles(oracle.Rng.philox(1234, t, 1), (L, nx), owalls, opv, opia, 1, [AR], dt, surf=True)
        s = surf.download()
        np.testing.assert_array_equal(s[:, 0], so[:, 0])  # np: exact counts
        np.testing.assert_allclose(s, so, rtol=1e-10, atol=1e-10 * np.abs(so).max())
        np.testing.assert_allclose(s, s_host, rtol=1e-12, atol=1e-12 * np.abs(so).max())  # same kernels, atomics in another order
        avg_ref = avg_ref + so * (1.0 / n_avg)  # avg_props! surface_props.jl:202-222
    np.testing.assert_allclose(avg.download(), avg_ref, rtol=1e-10, atol=1e-10 * np.abs(avg_ref).max())
    assert avg_ref[0, 0] > 100 and avg_ref[1, 0] > 100
    # reduce_surf_props! (surface_props.jl:232-252): the target is cleared, then the chunks are added in list order
    a, b, target = mb.SurfProps(ctx), mb.SurfProps(ctx), mb.SurfProps(ctx)
    ra, rb = rng.normal(size=(2, 11)), rng.normal(size=(2, 11))
    a.upload(ra)
    b.upload(rb)
    target.upload(np.full((2, 11), 7.0))
    mb.reduce_surf_props(target, [a, b, surf])
    np.testing.assert_array_equal(target.download(), (ra + rb) + surf.download())
    surf.clear()
    assert not surf.download().any()
    for o in (surf, avg, a, b, target, pv, pia, pv2, pia2):
        o.close()


def test_avg_props_parity(mb, oracle, ctx):
    """avg_props!(phys_props_avg, phys_props, n_avg_timesteps) (physical_props.jl:281-299) on the device against the oracle's, over
    several steps of changing cell properties: bit-identical accumulation (one multiply and one add per entry)."""
    rng = np.random.default_rng(34)
    n, n_cells, L = 30000, 25, 2.5
    rows = maxwellian_rows(rng, n, L, vw=True)
    opv, opia = oracle_state(oracle, rows, n_cells)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pp, pavg = mb.PhysProps(n_cells, 1, ctx=ctx), mb.PhysProps(n_cells, 1, ctx=ctx)
    oavg = oracle.Props(n_cells, 1)
    n_avg = 5
    for t in range(n_avg):
        cur = opv.logical(1, n)
        cur[:, 1:4] *= 1.0 + 0.05 * t  # the properties change from step to step
        opv.set_logical(1, cur)
        pv.set_logical(1, cur)
        mb.compute_props([pv], pia, [AR], pp)
        mb.avg_props(pavg, pp, n_avg)
        d = pp.download()
        cur_props = oracle.Props(n_cells, 1)
        cur_props.lpa[:] = d["lpa"]
        cur_props.np[:], cur_props.n[:], cur_props.v[:], cur_props.T[:] = d["np"], d["n"], d["v"], d["T"]
        oracle.avg_props(oavg, cur_props, n_avg)
    a = pavg.download()
    for k, ref in (("np", oavg.np), ("n", oavg.n), ("v", oavg.v), ("T", oavg.T), ("lpa", oavg.lpa)):
        np.testing.assert_array_equal(a[k], ref)
    o = oracle.compute_props([opv], opia, [AR])
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)
    pv.close()
    pia.close()


def test_fp_linear_rejects_cells_with_a_second_group(mb, oracle, ctx):
    """fp_linear! is only defined for sorted cells (the reference's group-2 branch throws, collision_fp.jl:57): a cell whose particles
    sit in two index groups is skipped and reported as MB_ERR_PRECONDITION instead of being relaxed with its neighbours' particles."""
    rng = np.random.default_rng(35)
    n = 600
    rows = maxwellian_rows(rng, n, 3.0)
    pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(3, 1, ctx)
    pv.set_logical(1, rows)
    ix = np.zeros((1, 3, 7), dtype=np.int64)
    ix[0, 0] = (150, 1, 100, 100, 551, 600, 50)   # group 2 at the tail (as variable-weight ntc! leaves it)
    ix[0, 1] = (200, 101, 300, 200, 0, -1, 0)
    ix[0, 2] = (250, 301, 550, 250, 0, -1, 0)
    pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    mb.fp_linear(mb.PhiloxRng(1), None, it, AR, pv, pia, (1, 3), 1, 2.59e-9 * 50, 1e-5)
    with pytest.raises(mb.MerzbildError) as e:
        ctx.sync()
    assert e.value.status == mb.MB_ERR_PRECONDITION
    after = pv.logical(1, n)
    np.testing.assert_array_equal(after[:100], rows[:100])      # the offending cell was left alone, both groups
    np.testing.assert_array_equal(after[550:], rows[550:])
    assert np.any(after[100:550, 1:4] != rows[100:550, 1:4])    # the sorted cells were relaxed
    pv.close()
    pia.close()

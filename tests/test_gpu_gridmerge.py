"""GPU parity: merge_grid_based! (mb_merge_grid_based) against the CPU oracle: pia bit-exact, merged particles bit-exact in weight and
to 1e-13 elsewhere (one thread per velocity cell walks its particles in the reference's order, so the sums are the oracle's), mass /
momentum / energy conserved to 1e-13 relative; the reference's own KATs (empty octants, 1-D x clamp) through the device path."""
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
    return np.array([w.sum(), *(w[:, None] * rows[:, 1:4]).sum(0), (w[:, None] * rows[:, 1:4] ** 2).sum()])


def _props_Tv(oracle, opv, opia):
    p = oracle.compute_props([opv], opia, [AR])
    return np.column_stack([p.T[0], p.v[0]])


@pytest.mark.parametrize("N,nb,mult", [(5000, (2, 2, 2), 3.5), (20000, (6, 5, 4), 2.0), (300, (1, 1, 1), 10.0), (40000, (16, 16, 16), 3.5)])
def test_grid_merge_single_cell_parity(mb, oracle, ctx, N, nb, mult):
    """0-D usage (bkw_varweight_grid.jl:78-91, test_merging_grid_merging.jl:60-95): extents from PhysProps of the cell."""
    rng = np.random.default_rng(N + nb[0])
    rows = maxwellian_rows(rng, N, 1.0, vw=True, w=1e12)
    rows[:, 1:4] += np.array([2000.0, 500.0, -400.0])
    opv, opia = oracle_state(oracle, rows, 1)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pp = mb.PhysProps(1, 1, ctx=ctx)
    mb.compute_props([pv], pia, [AR], pp)
    d = pp.download()
    Tv = np.column_stack([d["T"][0], d["v"][0]])  # the device props feed both sides: the grid extents are then bit-identical
    omg, mg = oracle.GridMerge(*nb, mult), mb.GridN2Merge(*nb, mult)
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 5, 1), omg, opv, opia, 1, 1, 1, AR, T_v=Tv) == 0
    mb.merge_grid_based(mb.PhiloxRng(5, 1), mg, pv, pia, 1, 1, AR, pp)
    assert_same_pia(opia, pia)
    nt = int(opia.n_total[0])
    assert nt < N and nt <= 2 * (nb[0] * nb[1] * nb[2] + 8)
    a, b = pv.logical(1, N), opv.logical(1, N)
    np.testing.assert_array_equal(a[:nt, 0], b[:nt, 0])
    assert_rows_close(a[:nt], b[:nt], 1e-13, "grid-merged particles")
    assert np.all(a[nt:, 0] == 0.0)
    m0, m1 = _moments(rows), _moments(a[:nt])
    np.testing.assert_allclose(m1, m0, rtol=1e-13)


def test_grid_merge_cell_range_parity(mb, oracle, ctx):
    """1-D usage: a range of cells, only those above the threshold are merged; explicit extents and the PhysProps variant; group 2
    present (particles appended at the tail by a variable-weight ntc!); x clamped into the domain."""
    rng = np.random.default_rng(9)
    n_cells, L = 40, 40e-5
    sizes = rng.integers(20, 400, n_cells)
    n = int(sizes.sum())
    rows = maxwellian_rows(rng, n, L, vw=True, w=1e15)
    rows[:, 4] = (np.repeat(np.arange(n_cells), sizes) + rng.uniform(0.0, 1.0, n)) * 1e-5
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=4 * n)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia, capacity=4 * n)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 2e15)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    mb.ntc(mb.PhiloxRng(1), cf, None, it, pv, pia, (1, n_cells), 1, 2.59e-9 * 4, 1e-5)      # splits -> group 2 at the tail
    oracle.ntc(oracle.Rng.philox(1234, 1), ocf, oit, opv, opia, 1, n_cells, 1, 2.59e-9 * 4, 1e-5)
    assert_same_pia(opia, pia)
    assert opia.indexer[0, :, 6].sum() > 50
    g = mb.Grid1DUniform(L, n_cells)
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    mb.compute_props([pv], pia, [AR], pp)
    d = pp.download()
    Tv = np.column_stack([d["T"][0], d["v"][0]])
    before = pv.logical(1, int(opia.n_total[0]))
    omg, mg = oracle.GridMerge(3, 3, 3, 3.0), mb.GridN2Merge(3, 3, 3, 3.0)
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 2), omg, opv, opia, 1, n_cells, 1, AR, T_v=Tv, threshold=150, grid=(L, n_cells)) == 0
    mb.merge_grid_based(mb.PhiloxRng(2), mg, pv, pia, (1, n_cells), 1, AR, pp, grid=g, threshold=150)
    assert_same_pia(opia, pia)
    merged = [c for c in range(n_cells) if sizes[c] > 150]
    assert len(merged) > 5 and all(opia.indexer[0, c, 0] <= 2 * 35 for c in merged)
    cap = len(pv)
    a, b = pv.logical(1, cap), opv.logical(1, cap)
    live = np.zeros(cap, dtype=bool)
    for c in range(n_cells):
        q = opia.indexer[0, c]
        if q[3] > 0:
            live[q[1] - 1:q[2]] = True
        if q[6] > 0:
            live[q[4] - 1:q[5]] = True
    np.testing.assert_array_equal(a[live, 0], b[live, 0])
    assert_rows_close(a[live], b[live], 1e-13, "grid merge over a cell range")
    assert a[live, 4].min() >= g.min_x and a[live, 4].max() <= g.max_x
    np.testing.assert_allclose(_moments(a[live]), _moments(before), rtol=1e-13)
    # explicit extents on what is left, then squash + sort like the drivers do
    ext = ((-900.0, 900.0), (-800.0, 850.0), (-1000.0, 700.0))
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 3), omg, opv, opia, 1, n_cells, 1, AR, extents=ext, threshold=60, grid=(L, n_cells)) == 0
    mb.merge_grid_based(mb.PhiloxRng(3), mg, pv, pia, (1, n_cells), 1, AR, ext[0], ext[1], ext[2], grid=g, threshold=60)
    assert_same_pia(opia, pia)
    mb.sort_particles(None, g, pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    assert_same_pia(opia, pia)
    nt = int(opia.n_total[0])
    assert_rows_close(pv.logical(1, nt), opv.logical(1, nt), 1e-13, "after squash + sort")
    assert pia.check(1) == (True, 0)


def test_grid_merge_reference_kats_on_device(mb, oracle, ctx):
    """test_merging_grid_merging.jl:97-137 (an octant of zero weight is dropped) and test_merging_grid_merging_1D.jl (x clamp)."""
    from test_oracle_kat_gridmerge import _octant_particles

    rows = _octant_particles([1.0] * 7 + [0.0])
    opv, opia = oracle_state(oracle, rows, 1)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pp = mb.PhysProps(1, 1, ctx=ctx)
    mb.compute_props([pv], pia, [AR], pp)
    d0 = pp.download()
    mb.merge_grid_based(mb.PhiloxRng(1), mb.GridN2Merge(2, 2, 2, 0.5), pv, pia, 1, 1, AR, pp)
    ix, nt, ct = pia.download()
    n1 = int(nt[0])
    assert n1 < 24 and tuple(ix[0, 0]) == (n1, 1, n1, n1, 0, -1, 0)
    a = pv.logical(1, n1)
    assert np.all(np.isfinite(a)) and np.all(a[:, 0] > 0)
    mb.compute_props([pv], pia, [AR], pp)
    d1 = pp.download()
    assert d1["np"][0, 0] == n1 and abs(d1["n"][0, 0] - d0["n"][0, 0]) <= np.finfo(float).eps * d0["n"][0, 0]
    assert abs(d1["T"][0, 0] - d0["T"][0, 0]) < 1e-12 * d0["T"][0, 0] and np.all(np.abs(d1["v"][0, 0] - d0["v"][0, 0]) < 1e-13)
    # 1-D: particles placed so that mean(x) +- std(x) leaves [0, 1]
    r2 = []
    for i, x in zip(range(1, 5), (0.05, 0.05, 0.05, 0.45)):
        r2.append([2.0, 0.5 - i ** 2, -3.0 + i, 4.0 + 0.3 * i, x, 0.0, 0.0])
    for i, x in zip(range(5, 9), (0.55, 0.95, 0.85, 0.99)):
        r2.append([3.0, 0.5 + i ** 2, -3.0 + 2 * i, 4.0 - i, x, 0.0, 0.0])
    g = mb.Grid1DUniform(1.0, 2)
    for t in range(1, 12):
        pv2, pia2 = mb.ParticleVector(8, ctx), mb.ParticleIndexerArray(2, 1, ctx)
        pv2.set_logical(1, np.array(r2))
        ixh = np.zeros((1, 2, 7), dtype=np.int64)
        ixh[0, 0] = (4, 1, 4, 4, 0, -1, 0)
        ixh[0, 1] = (4, 5, 8, 4, 0, -1, 0)
        pia2.upload(ixh, np.array([8]), np.array([1], dtype=np.uint8))
        pp2 = mb.PhysProps(2, 1, ctx=ctx)
        mb.compute_props([pv2], pia2, [AR], pp2)
        mb.merge_grid_based(mb.PhiloxRng(t), mb.GridN2Merge(1, 1, 1, 5.5), pv2, pia2, (1, 2), 1, AR, pp2, grid=g)
        ix2, nt2, ct2 = pia2.download()
        assert tuple(ix2[0, 0][:4]) == (2, 1, 2, 2) and tuple(ix2[0, 1][:4]) == (2, 5, 6, 2) and nt2[0] == 4 and ct2[0] == 0
        x = np.concatenate([pv2.logical(1, 2)[:, 4], pv2.logical(5, 6)[:, 4]])
        assert x.min() >= g.min_x and x.max() <= g.max_x


def test_grid_merge_argument_errors(mb, ctx):
    pv, pia = mb.ParticleVector(16, ctx), mb.ParticleIndexerArray(1, 1, ctx)
    with pytest.raises(mb.MerzbildError):
        mb.merge_grid_based(mb.PhiloxRng(1), mb.GridN2Merge(32, 32, 32, 3.5), pv, pia, 1, 1, AR, (-1.0, 1.0), (-1.0, 1.0), (-1.0, 1.0))


def test_grid_merging_buffer_sorting_reference_kat(mb, oracle, ctx):
    """test_merging_grid_buffer_sorting.jl:44-204 through the C ABI: the reference's post-merge counts and indexers (2 particles with one
    velocity cell at 10 thermal speeds; 10 with the inner cell + outer octants at 1 thermal speed; the second case keeps every
    deletion inside group 2), and the sort that closes the hole between the cells."""
    from test_oracle_kat_octree_vhs import _buffer_sorting_state

    for n_gr1, mult, expect, expect_c2 in ((50, 10.0, (2, 1, 2, 2), (10, 51, 60, 10)), (5, 1.0, (10, 1, 5, 5, 16, 20, 5), (10, 6, 15, 10))):
        rows, opv, opia = _buffer_sorting_state(oracle, n_gr1)
        pv, pia = mirror_to_device(mb, ctx, opv, opia)
        pp = mb.PhysProps(2, 1, ctx=ctx)
        mb.compute_props([pv], pia, [AR], pp)
        mb.merge_grid_based(mb.PhiloxRng(1), mb.GridN2Merge(1, 1, 1, mult), pv, pia, 1, 1, AR, pp)
        ix, nt, ct = pia.download()
        assert tuple(ix[0, 0][:len(expect)]) == expect and tuple(ix[0, 1][:4]) == expect_c2 and ct[0] == 0 and nt[0] == expect[0] + 10
        mb.compute_props([pv], pia, [AR], pp)
        d = pp.download()
        assert d["np"][0].tolist() == [float(expect[0]), 10.0] and abs(d["n"][0, 0] - 90.0) < 1e-12 and d["n"][0, 1] == 10000.0
        mb.sort_particles(None, mb.Grid1DUniform(8.0, 2), pv, pia, 1)
        ix, nt, ct = pia.download()
        k = expect[0]
        assert tuple(ix[0, 0]) == (k, 1, k, k, 0, -1, 0) and tuple(ix[0, 1]) == (10, k + 1, k + 10, 10, 0, -1, 0) and ct[0] == 1

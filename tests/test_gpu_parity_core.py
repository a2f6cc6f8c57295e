"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bit-exact for sort / indexing; 1e-12 relative for floating-point moments; the sequential-per-
cell stochastic kernels replay the oracle's Philox streams draw for draw, so counters must match exactly and velocities
to 1e-12 relative (libm pow/sincos/log differ from glibc in the last bits)."""
import numpy as np
import pytest

from parity_util import AR, HE, assert_rows_close, assert_same_pia, maxwellian_rows, mirror_to_device, oracle_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(mb):
    c = mb.Context(0, 1234)
    yield c
    c.close()


# ------------------------------------------------------------------------------------------------------------ sort
def test_sort_reference_kat(mb, oracle, ctx):
    """test/test_grid_sorting.jl:26-125: 8 particles in reversed order, 2 then 4 cells, uneven 3/1/0/4 split."""
    rows = np.zeros((8, 7))
    rows[:, 0] = 2.5e9
    rows[:, 4] = [0.99 * 8.0 * (9.0 - i) / 8 for i in range(1, 9)]
    opv, opia = oracle_state(oracle, rows, 2)
    opia.indexer[0, 0] = (4, 1, 4, 4, 0, -1, 0)
    opia.indexer[0, 1] = (4, 5, 8, 4, 0, -1, 0)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(8.0, 2)
    mb.sort_particles(None, g, pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(8.0, 2))
    assert_same_pia(opia, pia)
    np.testing.assert_array_equal(pv.logical(1, 8), opv.logical(1, 8))
    np.testing.assert_array_equal(pv.logical(1, 8)[:, 4], rows[[4, 5, 6, 7, 0, 1, 2, 3], 4])  # index [5,6,7,8,1,2,3,4]
    # 4 cells, everything indexed from cell 1
    opv, opia = oracle_state(oracle, rows, 4)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g4 = mb.Grid1DUniform(8.0, 4)
    mb.sort_particles(None, g4, pv, pia, 1)
    np.testing.assert_array_equal(pv.logical(1, 8)[:, 4], rows[[6, 7, 4, 5, 2, 3, 0, 1], 4])  # index [7,8,5,6,3,4,1,2]
    ix = pia.indexer
    for cell in range(1, 5):
        assert ix[0, cell - 1].tolist() == [2, 1 + 2 * (cell - 1), 2 + 2 * (cell - 1), 2, 0, -1, 0]
    # uneven 3/1/0/4 incl. an empty cell (0,-1)
    rows2 = rows.copy()
    rows2[0:4, 4] = 6.75
    rows2[4, 4] = 2.5
    rows2[5:8, 4] = 0.5
    pv.set_logical(1, rows2)
    mb.sort_particles(None, g4, pv, pia, 1)
    counts, starts, ends = [3, 1, 0, 4], [1, 4, 0, 5], [3, 4, -1, 8]
    ix = pia.indexer
    for c in range(4):
        assert ix[0, c].tolist() == [counts[c], starts[c], ends[c], counts[c], 0, -1, 0]
    np.testing.assert_array_equal(pv.logical(1, 8)[:, 4], rows2[[5, 6, 7, 4, 0, 1, 2, 3], 4])  # index [6,7,8,5,1,2,3,4]
    # cells-known variant (grid_sorting.jl:128)
    pv.set_logical(1, rows2)
    pv.set_cell(1, [4, 4, 4, 4, 2, 1, 1, 1])
    mb.sort_particles(None, pv, pia, 1)
    ix = pia.indexer
    for c in range(4):
        assert ix[0, c].tolist() == [counts[c], starts[c], ends[c], counts[c], 0, -1, 0]
    np.testing.assert_array_equal(pv.logical(1, 8)[:, 4], rows2[[5, 6, 7, 4, 0, 1, 2, 3], 4])


@pytest.mark.parametrize("n,n_cells", [(0, 5), (1, 1), (33, 7), (5000, 1), (20000, 37), (100000, 1024), (30000, 3)])
def test_sort_general_random(mb, oracle, ctx, n, n_cells):
    """arbitrary (unsorted) input -> general path; bit-exact logical order and pia against the oracle."""
    rng = np.random.default_rng(n + n_cells)
    L = 2.0
    rows = maxwellian_rows(rng, n, L, vw=True)
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=max(n, 1))
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    mb.sort_particles(None, mb.Grid1DUniform(L, n_cells), pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    assert ctx.sort_last_path == 2
    assert_same_pia(opia, pia)
    np.testing.assert_array_equal(pv.logical(1, max(n, 1)), opv.logical(1, max(n, 1)))
    assert pia.check(1) == (True, 0)
    if n > 0:
        np.testing.assert_array_equal(pv.cell(1, n), opv.cell[:n])  # pv.cell is written by the grid variant and not permuted


@pytest.mark.parametrize("w", [1, 2, 4, 8, 15])
def test_sort_band_path(mb, oracle, ctx, w):
    """the per-timestep case: sorted layout + small displacements -> band path; same bits as the oracle and as the general path."""
    rng = np.random.default_rng(7 + w)
    n, n_cells, L = 60000, 300, 3.0
    dx = L / n_cells
    rows = maxwellian_rows(rng, n, L, vw=True)
    opv, opia = oracle_state(oracle, rows, n_cells)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, n_cells)
    ctx.set_band_halfwidth(w)
    try:
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        for step in range(3):
            cur = opv.logical(1, n)
            # displacement up to +-0.95 * w cells, clipped at the walls
            c0 = np.floor(cur[:, 4] / dx)
            move = rng.uniform(-0.95 * w * dx, 0.95 * w * dx, n)
            if step == 1:  # cells 3..7 keep still except cell 5, which empties into cell 6 (an empty cell for w <= 2)
                move[(c0 >= 3) & (c0 <= 7)] = 0.0
            cur[:, 4] = np.clip(cur[:, 4] + move, 1e-9, L - 1e-9)
            if step == 1:
                m5 = c0 == 5
                cur[m5, 4] = 6 * dx + (cur[m5, 4] - 5 * dx) * 0.5
            opv.set_logical(1, cur)
            pv.set_logical(1, cur)
            mb.sort_particles(None, g, pv, pia, 1)
            oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
            assert ctx.sort_last_path == 1, "band path expected"
            assert_same_pia(opia, pia)
            np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))
            assert ctx.sort_last_extras == 0
        # particles that leave the band are ranked separately (hybrid): still the band path, same result
        cur = opv.logical(1, n)
        c_old = np.floor(cur[:, 4] * (1.0 / dx))
        cur[::997, 4] = rng.uniform(0, L, len(cur[::997]))
        n_out = int(np.sum(np.abs(np.floor(cur[:, 4] * (1.0 / dx)) - c_old) > w))
        opv.set_logical(1, cur)
        pv.set_logical(1, cur)
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert ctx.sort_last_path == 1
        assert ctx.sort_last_extras == n_out > 0
        assert_same_pia(opia, pia)
        np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))
    finally:
        ctx.set_band_halfwidth(2)


@pytest.mark.parametrize("w", [1, 2, 15])
@pytest.mark.parametrize("frac", [0.001, 0.01, 0.1, 0.6])
def test_sort_band_with_outliers(mb, oracle, ctx, w, frac):
    """Hybrid band sort (grid_sorting.jl:58-113 is the contract: stable counting sort): a fraction of the particles jumps anywhere in the
    domain, far outside the band.  They are ranked among themselves by original position in front of / behind the band groups of their
    destination cell; logical order and pia stay bit-identical to the oracle's, the band path is kept (no fall back), and for narrow
    bands the cached cell moments (movers, extras included) still serve compute_props_sorted!."""
    rng = np.random.default_rng(int(1000 * frac) + w)
    n, n_cells, L = 70000, 350, 3.5
    dx = L / n_cells
    inv_dx = 1.0 / dx  # get_cell multiplies by inv_dx (grid_uniform1D.jl:97-99)
    rows = maxwellian_rows(rng, n, L, vw=True)
    rows[:, 1:4] += np.array([500.0, -200.0, 100.0])
    opv, opia = oracle_state(oracle, rows, n_cells)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, n_cells)
    ctx.set_band_halfwidth(w)
    try:
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        pp = mb.PhysProps(n_cells, 1, ctx=ctx)
        for step in range(3):
            cur = opv.logical(1, n)
            c_old = np.floor(cur[:, 4] * inv_dx)
            cur[:, 4] = np.clip(cur[:, 4] + rng.uniform(-0.9 * w * dx, 0.9 * w * dx, n), 1e-9, L - 1e-9)
            jump = rng.uniform(0, 1, n) < frac
            if step == 2:  # a crowd lands in two cells: long extras regions in front of and behind the band
                far = np.flatnonzero(jump)
                cur[far[: len(far) // 2], 4] = rng.uniform(10 * dx, 11 * dx, len(far) // 2)
                cur[far[len(far) // 2:], 4] = rng.uniform((n_cells - 20) * dx, (n_cells - 19) * dx, len(far) - len(far) // 2)
            else:
                cur[jump, 4] = rng.uniform(1e-9, L - 1e-9, int(jump.sum()))
            n_out = int(np.sum(np.abs(np.floor(cur[:, 4] * inv_dx) - c_old) > w))
            opv.set_logical(1, cur)
            pv.set_logical(1, cur)
            mb.sort_particles(None, g, pv, pia, 1)
            oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
            crowd = step == 2 and n_out // 2 > 1024  # more extras in one cell than the ranking handles: general path, same result
            assert ctx.sort_last_path == (2 if crowd else 1)
            if not crowd:
                assert ctx.sort_last_extras == n_out
            assert_same_pia(opia, pia)
            np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))
            assert pia.check(1) == (True, 0)
            mb.compute_props_sorted([pv], pia, [AR], pp)
            d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
            np.testing.assert_array_equal(d["np"], o.np)
            np.testing.assert_allclose(d["n"], o.n, rtol=1e-13)
            np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-12 * 500)
            np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)
    finally:
        ctx.set_band_halfwidth(2)


def test_sort_large_cell_segments(mb, oracle, ctx):
    """cells larger than the shared-memory segment sort (8192) use the global-memory network."""
    rng = np.random.default_rng(5)
    n, n_cells, L = 50000, 2, 1.0
    rows = maxwellian_rows(rng, n, L)
    rows[: n // 2, 4] = rng.uniform(0.5, 1.0, n // 2)  # first half of the input goes to cell 2 -> heavy reordering
    rows[n // 2:, 4] = rng.uniform(0.0, 1.0, n - n // 2)
    opv, opia = oracle_state(oracle, rows, n_cells)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    mb.sort_particles(None, mb.Grid1DUniform(L, n_cells), pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    assert_same_pia(opia, pia)
    np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))


# ------------------------------------------------------------------------------------------------------------ props
@pytest.mark.parametrize("band", [2, 0])
@pytest.mark.parametrize("n,n_cells", [(20000, 37), (3000, 64), (50000, 5), (40000, 700)])
def test_props_after_general_sort_use_the_cached_moments(mb, oracle, ctx, n, n_cells, band):
    """compute_props_sorted! right after a general-path sort of small cells reads the moments the gather-by-cell pass cached (no particle
    traffic); they must match the oracle's two-pass values like the band path's do (1e-12 on T and v, 1e-13 on n).  band = 0 (the caller
    switched the band path off: a fully scattered sort) takes the gather through 64-byte records."""
    ctx.set_band_halfwidth(band)
    try:
        _props_after_general_sort(mb, oracle, ctx, n, n_cells)
    finally:
        ctx.set_band_halfwidth(2)


def _props_after_general_sort(mb, oracle, ctx, n, n_cells):
    rng = np.random.default_rng(n + n_cells)
    L = 2.0
    rows = maxwellian_rows(rng, n, L, vw=True)
    rows[:, 1:4] += np.array([400.0, -300.0, 150.0])  # a drift: the shifted one-pass sums must not lose digits
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=n)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, n_cells)
    mb.sort_particles(None, g, pv, pia, 1)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    assert ctx.sort_last_path == 2
    assert_same_pia(opia, pia)
    np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    l0 = ctx.kernel_launches
    mb.compute_props_sorted([pv], pia, [AR], pp)
    if n // n_cells <= 2048:
        assert ctx.kernel_launches - l0 == 1  # the cached kernel only
    d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
    np.testing.assert_array_equal(d["np"], o.np)
    np.testing.assert_allclose(d["n"], o.n, rtol=1e-13)
    np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-12 * 500)
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)
    # touching the particles in between voids the cache: the regular kernels run and give the same numbers
    pv.set_logical(1, pv.logical(1, 1))
    l0 = ctx.kernel_launches
    mb.compute_props_sorted([pv], pia, [AR], pp)
    if n // n_cells <= 2048:
        assert ctx.kernel_launches - l0 > 1
    d2 = pp.download()
    np.testing.assert_allclose(d2["T"], o.T, rtol=1e-12)
    np.testing.assert_allclose(d2["v"], o.v, rtol=1e-12, atol=1e-12 * 500)


def test_compute_props_sorted_chunks_reference_kat(mb, ctx):
    """test/test_chunking.jl:57-105 through the C ABI: a cell chunk only touches its own cells (the others keep their values), the
    grid variant of a PhysProps(...; ndens_not_Np=true) divides by the cell volume."""
    rows = np.array([[i, 0.0, 0.0, 0.0, x, 0.0, 1.0] for i, x in zip((1.0, 2.0, 3.0, 4.0), (3.0, 3.0, 3.0, 7.0))])
    pv, pia = mb.ParticleVector(4, ctx), mb.ParticleIndexerArray(5, 1, ctx)
    pv.set_logical(1, rows)
    ix = np.tile(np.array([0, 0, -1, 0, 0, -1, 0], dtype=np.int64), (1, 5, 1))
    ix[0, 1] = (3, 1, 3, 3, 0, -1, 0)
    ix[0, 3] = (1, 4, 4, 1, 0, -1, 0)
    pia.upload(ix, np.array([4]), np.array([1], dtype=np.uint8))
    pp = mb.PhysProps(5, 1, ctx=ctx)
    mb.compute_props_sorted([pv], pia, [AR], pp, cell_chunk=(1, 1))
    assert pp.download()["np"][0].tolist() == [0.0] * 5
    mb.compute_props_sorted([pv], pia, [AR], pp, cell_chunk=(1, 3))
    d = pp.download()
    assert d["np"][0].tolist() == [0.0, 3.0, 0.0, 0.0, 0.0] and d["n"][0, 1] == 6.0
    mb.compute_props_sorted([pv], pia, [AR], pp, cell_chunk=(3, 4))
    d = pp.download()
    assert d["np"][0].tolist() == [0.0, 3.0, 0.0, 1.0, 0.0] and d["n"][0, 3] == 4.0 and d["n"][0, 1] == 6.0  # cell 2 is not reset
    g = mb.Grid1DUniform(10.0, 5)
    pn = mb.PhysProps(5, 1, ndens_not_Np=True, ctx=ctx)
    mb.compute_props_sorted([pv], pia, [AR], pn, g, cell_chunk=(1, 3))
    d = pn.download()
    assert d["np"][0].tolist() == [0.0, 3.0, 0.0, 0.0, 0.0] and d["n"][0, 1] == 3.0
    mb.compute_props_sorted([pv], pia, [AR], pn, g, cell_chunk=(4, 4))
    d = pn.download()
    assert d["np"][0].tolist() == [0.0, 3.0, 0.0, 1.0, 0.0] and d["n"][0, 1] == 3.0 and d["n"][0, 3] == 2.0


def test_compute_props_reference_kat(mb, oracle, ctx):
    """test/test_computes.jl:6-67: 2000 identical particles -> n, v exact, T ~ 0."""
    n = 2000
    rows = np.tile(np.array([2.0, 1.0, -2.0, 3.0, 0.5, 0.5, 0.5]), (n, 1))
    opv, opia = oracle_state(oracle, rows, 1)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pp = mb.PhysProps(1, 1, ctx=ctx)
    mb.compute_props([pv], pia, [AR], pp)
    d = pp.download()
    assert d["np"][0, 0] == n and d["lpa"][0] == n
    assert abs(d["n"][0, 0] - 2.0 * n) < 1e-9
    assert np.max(np.abs(d["v"][0, 0] - [1.0, -2.0, 3.0])) < 4e-15
    assert abs(d["T"][0, 0]) < 1e-10


@pytest.mark.parametrize("n,n_cells", [(40000, 64), (60000, 2)])
def test_compute_props_parity(mb, oracle, ctx, n, n_cells):
    """compute_props!, compute_props_sorted! (Np and ndens variants) and total moments vs the oracle, 1e-12 relative."""
    rng = np.random.default_rng(11)
    L = 1.0
    rows = maxwellian_rows(rng, n, L, vw=True, w=1e10)
    rows[:, 2] += 500.0  # |vbar| = 500 m/s: a one-pass variance would lose ~6 digits here
    opv, opia = oracle_state(oracle, rows, n_cells)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    powers = [4, 6, 8]
    pp = mb.PhysProps(n_cells, 1, powers, Tref=300.0, ctx=ctx)
    mb.compute_props_with_total_moments([pv], pia, [AR], pp)
    d = pp.download()
    o = oracle.compute_props([opv], opia, [AR], powers, 300.0, True)
    np.testing.assert_array_equal(d["np"], o.np)
    for k in ("n", "T", "moments"):
        np.testing.assert_allclose(d[k], getattr(o, k), rtol=1e-12, atol=0)
    np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-12 * 500)
    g = mb.Grid1DUniform(L, n_cells)
    for ndens in (False, True):
        pps = mb.PhysProps(n_cells, 1, ndens_not_Np=ndens, ctx=ctx)
        mb.compute_props_sorted([pv], pia, [AR], pps, grid=g if ndens else None)
        ds = pps.download()
        os_ = oracle.compute_props_sorted([opv], opia, [AR], grid=(L, n_cells) if ndens else None)
        np.testing.assert_array_equal(ds["np"], os_.np)
        np.testing.assert_allclose(ds["n"], os_.n, rtol=1e-12)
        np.testing.assert_allclose(ds["T"], os_.T, rtol=1e-12)
        np.testing.assert_allclose(ds["v"], os_.v, rtol=1e-12, atol=1e-12 * 500)


def test_compute_props_two_groups(mb, oracle, ctx):
    """compute_props! walks group 1 and group 2 (physical_props.jl:118-146); the sorted variant ignores group 2."""
    rng = np.random.default_rng(3)
    rows = maxwellian_rows(rng, 30, 1.0, vw=True)
    opv, opia = oracle_state(oracle, rows, 2)
    opia.indexer[0, 0] = (12, 1, 8, 8, 21, 24, 4)
    opia.indexer[0, 1] = (18, 9, 20, 12, 25, 30, 6)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    pp = mb.PhysProps(2, 1, ctx=ctx)
    mb.compute_props([pv], pia, [AR], pp)
    d, o = pp.download(), oracle.compute_props([opv], opia, [AR])
    np.testing.assert_array_equal(d["np"], o.np)
    np.testing.assert_allclose(d["n"], o.n, rtol=1e-13)
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)
    np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-10)
    mb.compute_props_sorted([pv], pia, [AR], pp)
    d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
    np.testing.assert_array_equal(d["np"], o.np)
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)


# ------------------------------------------------------------------------------------------------------------ convect
def test_convection_specular_kat(mb, oracle, ctx):
    """test/test_convection_1D.jl:1-75: 4 particles, specular walls, L = 50, dt = 2 -> x = 20.5, 29.0, 23.0, 3.55; cells 8,42,47,59."""
    rows = np.array([
        [1.0, -1.25, -1.5, 4.0, 23.0, -8.0, 7.5],
        [2.0, 11.0, -3.0, 1.0, 49.0, 6.0, -3.0],
        [3.0, -20.0, 0.0, 2.0, 17.0, 1.0, 3.0],
        [4.0, -49.0, -20.0, 13.0, 1.55, -1.0, 9.0],
    ])
    for compute_cell in (False, True):
        opv, opia = oracle_state(oracle, rows, 100)
        pv, pia = mirror_to_device(mb, ctx, opv, opia)
        g = mb.Grid1DUniform(50.0, 100)
        walls = mb.MaxwellWalls1D(1.0, 1.0, 0.0, 0.0, 0.0, 0.0)
        mb.convect_particles(mb.PhiloxRng(1), g, walls, pv, pia, 1, AR, 2.0, compute_cell=compute_cell)
        if compute_cell:
            assert pv.cell(1, 4).tolist() == [42, 59, 47, 8]
            mb.sort_particles(None, pv, pia, 1)
        else:
            mb.sort_particles(None, g, pv, pia, 1)
        lg = pv.logical(1, 4)
        assert np.max(np.abs(lg[0, 4:7] - [3.55, -1.0, 9.0])) < 3.65e-15 and lg[0, :4].tolist() == [4.0, -49.0, -20.0, 13.0]
        assert np.max(np.abs(lg[1, 4:7] - [20.5, -8.0, 7.5])) < 5e-16 and lg[1, :4].tolist() == [1.0, -1.25, -1.5, 4.0]
        assert np.max(np.abs(lg[2, 4:7] - [23.0, 1.0, 3.0])) < 5e-16 and lg[2, :4].tolist() == [3.0, 20.0, 0.0, 2.0]
        assert np.max(np.abs(lg[3, 4:7] - [29.0, 6.0, -3.0])) < 5e-16 and lg[3, :4].tolist() == [2.0, -11.0, -3.0, 1.0]


@pytest.mark.parametrize("acc", [(1.0, 1.0), (0.3, 0.8), (0.0, 1.0)])
def test_convection_diffuse_parity(mb, oracle, ctx, acc):
    """Maxwell walls with the shared per-particle Philox streams: positions/velocities and SurfProps vs the oracle."""
    rng = np.random.default_rng(21)
    n, nx, L, dt = 50000, 50, 5e-4, 2.59e-7  # large dt: ~half of the particles reach a wall, some reflect twice
    rows = maxwellian_rows(rng, n, L, w=1e10, vw=True)
    opv, opia = oracle_state(oracle, rows, nx)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, nx)
    walls = mb.MaxwellWalls1D(300.0, 450.0, -500.0, 500.0, acc[0], acc[1])
    ctx.set_seed(77)
    s = mb.convect_particles(mb.PhiloxRng(timestep=5, substream=2), g, walls, pv, pia, 1, AR, dt, surf_props=True, compute_cell=True)
    so = oracle.convect_particles(oracle.Rng.philox(77, 5, 2), (L, nx), (300.0, 450.0, -500.0, 500.0, acc[0], acc[1]), opv, opia, 1, [AR], dt, surf=True,
                                  compute_cell=True)
    a, b = pv.logical(1, n), opv.logical(1, n)
    assert np.count_nonzero(a[:, 1] != rows[:, 1]) > n // 25  # plenty of wall hits
    assert_rows_close(a, b, 1e-12, "convect")
    np.testing.assert_array_equal(pv.cell(1, n), opv.cell[:n])
    assert s[0, 0] == so[0, 0] and s[1, 0] == so[1, 0]
    np.testing.assert_allclose(s, so, rtol=1e-10, atol=1e-10 * np.abs(so).max())
    ctx.set_seed(1234)


def test_convection_noncontiguous(mb, oracle, ctx):
    """test/test_convection_1D.jl:236-290: only particles the pia points to are moved."""
    rows = np.zeros((30, 7))
    for i in range(1, 31):
        heavy = 11 <= i <= 25
        rows[i - 1] = [10000.0 if heavy else 1.0, -1000.0 if heavy else 1.0, 0, 0, 0.75, 0, 0]
    opv, opia = oracle_state(oracle, rows, 100)
    opia.n_total[0] = 15
    opia.indexer[0, 0] = (0, 0, -1, 0, 0, -1, 0)
    opia.indexer[0, 1] = (15, 1, 10, 10, 26, 30, 5)
    opia.contiguous[0] = 0
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(50.0, 100)
    mb.convect_particles(mb.PhiloxRng(0), g, mb.MaxwellWalls1D(1.0, 1.0, 0, 0, 0, 0), pv, pia, 1, AR, 1.0)
    oracle.convect_particles(oracle.Rng.philox(1234, 0, 0), (50.0, 100), (1.0, 1.0, 0, 0, 0, 0), opv, opia, 1, [AR], 1.0)
    np.testing.assert_array_equal(pv.logical(1, 30), opv.logical(1, 30))
    lg = pv.logical(1, 30)
    assert np.all(lg[10:25, 4] == 0.75) and np.all(lg[:10, 4] == 1.75) and np.all(lg[25:, 4] == 1.75)


# ------------------------------------------------------------------------------------------------------------ NTC
def _couette_like(oracle, mb, ctx, n_cells, ppc, seed, vw=False, capacity_mult=1.0):
    rng = np.random.default_rng(seed)
    L = n_cells * 1e-5
    n = n_cells * ppc
    Fnum = 1e-5 * 5e22 / ppc
    rows = maxwellian_rows(rng, n, L, w=Fnum, vw=vw)
    cap = int(n * capacity_mult)
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=cap)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia, capacity=cap)
    return L, n, Fnum, opv, opia, pv, pia


def test_ntc_equal_weight_parity(mb, oracle, ctx):
    """ntc_equal_weight! over all cells, 5 steps: candidate / collision counters identical, sigma_g_w_max and velocities 1e-12."""
    n_cells, ppc, dt = 40, 400, 2.59e-9 * 40  # enlarged dt: ~20-40 candidates per cell and step
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 101)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    V = L / n_cells
    total = 0
    for t in range(1, 6):
        mb.ntc_equal_weight(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, dt, V)
        oracle.ntc(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, dt, V, equal_weight=True)
        d = cf.download()
        np.testing.assert_array_equal(d["n_coll"], ocf.n_coll)
        np.testing.assert_array_equal(d["n_coll_performed"], ocf.n_coll_performed)
        np.testing.assert_array_equal(d["n_eq_w_coll_performed"], ocf.n_eq_w_coll_performed)
        np.testing.assert_allclose(d["sigma_g_w_max"], ocf.sigma_g_w_max, rtol=1e-13)
        total += int(ocf.n_coll_performed.sum())
    assert total > 1000
    a, b = pv.logical(1, n), opv.logical(1, n)
    assert_rows_close(a, b, 1e-12, "ntc equal weight")
    np.testing.assert_array_equal(a[:, [0, 4, 5, 6]], b[:, [0, 4, 5, 6]])
    # momentum and energy of the whole population conserved
    for d in range(1, 4):
        assert abs(a[:, d].sum() - rows_sum(opv, n, d)) <= 1e-9 * np.abs(a[:, d]).sum()


def rows_sum(opv, n, d):
    return opv.logical(1, n)[:, d].sum()


def test_ntc_per_cell_call_equals_range_call(mb, oracle, ctx):
    """the reference's per-cell call (cell_lo == cell_hi) gives the same state as one launch over the range."""
    n_cells, ppc, dt = 12, 200, 2.59e-9 * 60
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 55)
    pv2, pia2 = mirror_to_device(mb, ctx, opv, opia)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum)
    cf, cf2 = mb.CollisionFactors(n_cells, s0, ctx), mb.CollisionFactors(n_cells, s0, ctx)
    mb.ntc_equal_weight(mb.PhiloxRng(3), cf, None, it, pv, pia, (1, n_cells), 1, dt, L / n_cells)
    for cell in range(1, n_cells + 1):
        mb.ntc_equal_weight(mb.PhiloxRng(3), cf2, None, it, pv2, pia2, cell, 1, dt, L / n_cells)
    np.testing.assert_array_equal(pv.logical(1, n), pv2.logical(1, n))
    assert cf.download()["n_coll_performed"].tolist() == cf2.download()["n_coll_performed"].tolist()


def test_ntc_variable_weight_parity(mb, oracle, ctx):
    """ntc! with splitting (collision_ntc.jl:223-270): new particles, group-2 ranges, n_total and weights vs the oracle."""
    n_cells, ppc, dt = 24, 300, 2.59e-9 * 8  # ~100 candidates per cell and step (the split windows are sized by the candidates)
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 202, vw=True, capacity_mult=4.0)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 2 * Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    V = L / n_cells
    g = mb.Grid1DUniform(L, n_cells)
    for t in range(1, 4):
        mb.ntc(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, dt, V)
        oracle.ntc(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, dt, V)
        d = cf.download()
        np.testing.assert_array_equal(d["n_coll"], ocf.n_coll)
        np.testing.assert_array_equal(d["n_coll_performed"], ocf.n_coll_performed)
        np.testing.assert_array_equal(d["n_eq_w_coll_performed"], ocf.n_eq_w_coll_performed)
        assert_same_pia(opia, pia)
        nt = int(opia.n_total[0])
        assert nt > n
        a, b = pv.logical(1, nt), opv.logical(1, nt)
        assert_rows_close(a, b, 1e-12, "ntc vw")
        np.testing.assert_array_equal(a[:, 0], b[:, 0])  # weights are exact
        assert pia.check(1) == (True, 0)
        # the sort merges group 2 back (and must take the general path: the layout is no longer sorted)
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert_same_pia(opia, pia)
        np.testing.assert_array_equal(pv.logical(1, nt), a[np.lexsort((np.arange(nt), np.floor(a[:, 4] * g.inv_dx)))])
        n = nt
    # mass is conserved by splitting
    assert abs(pv.logical(1, n)[:, 0].sum() - opv.logical(1, n)[:, 0].sum()) == 0.0


@pytest.mark.parametrize("vw", [True, False])
@pytest.mark.parametrize("sizes", [(5000,), (3000, 150, 2600, 40, 2048)])
def test_ntc_large_cells_parity(mb, oracle, ctx, vw, sizes):
    """Cells of >= 2048 particles are collided by the speculative warp kernel (one warp per cell, 32 candidates evaluated at a time,
    the first non-trivial one executed for real): it must replay the reference's sequential loop draw for draw -- candidate and
    collision counters, sigma_g_w_max, split particles, pia identical to the oracle -- also next to small cells in the same range."""
    rng = np.random.default_rng(4242 + len(sizes))
    n_cells, n = len(sizes), sum(sizes)
    L = n_cells * 1e-5
    Fnum = 1e-5 * 5e22 / 300
    rows = maxwellian_rows(rng, n, L, w=Fnum, vw=False)
    if vw:
        rows[:, 0] *= 10.0 ** rng.uniform(-2.0, 1.0, n)  # weights over three decades: low acceptance, long rejection runs
    rows[:, 4] = (np.repeat(np.arange(n_cells), sizes) + rng.uniform(0.01, 0.99, n)) * 1e-5
    cap = 12 * n  # the split windows are sized by the candidate counts (thousands per large cell)
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=cap)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mirror_to_device(mb, ctx, opv, opia, capacity=cap)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    g = mb.Grid1DUniform(L, n_cells)
    dt, V = 2.59e-9 * 0.3, L / n_cells
    total = 0
    for t in range(1, 4):
        mb.ntc(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, dt, V, equal_weight=not vw)
        oracle.ntc(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, dt, V, equal_weight=not vw)
        d = cf.download()
        np.testing.assert_array_equal(d["n_coll"], ocf.n_coll)
        np.testing.assert_array_equal(d["n_coll_performed"], ocf.n_coll_performed)
        np.testing.assert_array_equal(d["n_eq_w_coll_performed"], ocf.n_eq_w_coll_performed)
        np.testing.assert_allclose(d["sigma_g_w_max"], ocf.sigma_g_w_max, rtol=1e-13)
        assert_same_pia(opia, pia)
        nt = int(opia.n_total[0])
        a, b = pv.logical(1, nt), opv.logical(1, nt)
        assert_rows_close(a, b, 1e-12, "ntc large cells")
        np.testing.assert_array_equal(a[:, 0], b[:, 0])
        total += int(ocf.n_coll.sum())
        # fold the split particles back into their cells (general sort path) so that the next step may split again
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert_same_pia(opia, pia)
    assert total > 200 * len([x for x in sizes if x >= 2048])  # the large cells really ran long candidate loops
    if vw:
        assert int(opia.n_total[0]) > n


def test_ntc_capacity_error(mb, oracle, ctx):
    """the reference would resize! (collision_ntc.jl:241-243); the device never grows implicitly: a split that finds no room raises
    MB_ERR_CAPACITY (the step's particle state is then invalid: restore it, resize, repeat)."""
    n_cells, ppc, dt = 4, 300, 2.59e-9 * 60
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 203, vw=True, capacity_mult=1.0)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    cf = mb.CollisionFactors(n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 2 * Fnum), ctx)
    before = pv.logical(1, n)
    ix0, nt0, ct0 = pia.download()
    mb.ntc(mb.PhiloxRng(1), cf, None, it, pv, pia, (1, n_cells), 1, dt, L / n_cells)
    with pytest.raises(mb.CapacityError):
        ctx.sync()
    pv.resize(int(4 * n))
    pv.set_logical(1, before)
    pia.upload(ix0, nt0, ct0)
    cf.fill(mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 2 * Fnum))
    mb.ntc(mb.PhiloxRng(1), cf, None, it, pv, pia, (1, n_cells), 1, dt, L / n_cells)
    ctx.sync()
    assert int(pia.n_total[0]) > n


def test_ntc_capacity_is_the_splits_not_the_candidates(mb, oracle, ctx):
    """Variable-weight ntc! needs room for the splits that HAPPEN.  With widely spread weights sigma_g_w_max follows the heaviest
    particle, a cell tests far more candidate pairs than it holds particles and accepts a few per cent of them: the per-cell windows
    (sized by the candidates) are shrunk to the free capacity, and the call succeeds and matches the oracle -- pia and particles --
    although n_total + candidates exceeds the capacity several times."""
    n_cells, ppc, dt = 6, 400, 2.59e-9 * 40
    L = n_cells * 1e-5
    rng = np.random.default_rng(77)
    n = n_cells * ppc
    Fnum = 1e-5 * 5e22 / ppc
    rows = maxwellian_rows(rng, n, L, w=Fnum)
    rows[:, 0] = Fnum * 10.0 ** rng.uniform(-3.0, 0.0, n)  # three decades of weights
    rows[:, 4] = (np.repeat(np.arange(n_cells), ppc) + rng.uniform(0.01, 0.99, n)) * 1e-5
    cap = int(1.6 * n)
    opv, opia = oracle_state(oracle, rows, n_cells, capacity=4 * n)
    oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
    pv, pia = mb.ParticleVector(cap, ctx), mb.ParticleIndexerArray(n_cells, 1, ctx)
    pv.set_logical(1, opv.logical(1, n))
    pia.upload(opia.indexer.copy(), opia.n_total.copy(), opia.contiguous.copy())
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    # sigma_g_w_max as a rare heavy, fast pair leaves it: 20 x the estimate -> 20 x the candidates, 1 / 20 of the acceptance probability
    s0 = 20.0 * mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    mb.ntc(mb.PhiloxRng(1), cf, None, it, pv, pia, (1, n_cells), 1, dt, L / n_cells)
    ctx.sync()
    oracle.ntc(oracle.Rng.philox(1234, 1), ocf, oit, opv, opia, 1, n_cells, 1, dt, L / n_cells)
    ncoll = cf.download()["n_coll"]
    nt = int(pia.n_total[0])
    assert n + int(ncoll.sum()) > 2 * cap, (n, int(ncoll.sum()), cap)  # the old requirement was far out of reach ...
    assert n < nt <= cap                                                   # ... the splits fit
    assert_same_pia(opia, pia)
    assert_rows_close(pv.logical(1, nt), opv.logical(1, nt), 1e-12, "vw ntc with shrunk windows")
    pv.close()
    pia.close()


def test_ntc_two_species_parity(mb, oracle, ctx):
    """README 2-species case shape (400 Ar @3000 K + 4000 He @360 K, Fnum 5e12, dt 2.5e-3, V = 1): the three ntc! calls of a
    step in the reference's order (Ar-Ar, He-Ar via the two-ParticleVector method, He-He), equal weight, 20 steps."""
    rng = np.random.default_rng(1)
    nAr, nHe, Fnum, dt, V = 400, 4000, 5e12, 2.5e-3, 1.0
    rAr = maxwellian_rows(rng, nAr, 1.0, T=3000.0, m=AR, w=Fnum)
    rHe = maxwellian_rows(rng, nHe, 1.0, T=360.0, m=HE, w=Fnum)
    opvA, opvH = oracle.OPV(nAr), oracle.OPV(nHe)
    opvA.fill_identity(rAr)
    opvH.particles[:nHe] = rHe
    opvH.index[:nHe] = np.arange(1, nHe + 1)
    opvH.nbuffer = 0
    opia = oracle.OPIA(1, 2)
    opia.indexer[0, 0] = (nAr, 1, nAr, nAr, 0, -1, 0)
    opia.indexer[1, 0] = (nHe, 1, nHe, nHe, 0, -1, 0)
    opia.n_total[:] = (nAr, nHe)
    pvA, pvH = mb.ParticleVector(nAr, ctx), mb.ParticleVector(nHe, ctx)
    pvA.set_logical(1, rAr)
    pvH.set_logical(1, rHe)
    pia = mb.ParticleIndexerArray(1, 2, ctx)
    pia.upload(opia.indexer.copy(), opia.n_total.copy(), opia.contiguous.copy())
    names = [("Ar", "Ar"), ("He", "Ar"), ("He", "He")]
    m = {"Ar": AR, "He": HE}
    T0 = {"Ar": 3000.0, "He": 360.0}
    its, oits, cfs, ocfs = [], [], [], []
    for a, b in names:
        d, o, Tref = oracle.VHS.get((a, b)) or oracle.VHS[(b, a)]
        its.append(mb.make_interaction(m[a], m[b], d, o, Tref))
        oits.append(oracle.make_interaction(m[a], m[b], d, o, Tref))
        s0 = mb.estimate_sigma_g_w_max(its[-1], m[a], m[b], T0[a], T0[b], Fnum)
        cfs.append(mb.CollisionFactors(1, s0, ctx))
        ocfs.append(oracle.CF(1, s0))
    for t in range(1, 21):
        mb.ntc(mb.PhiloxRng(t, 0), cfs[0], None, its[0], pvA, pia, 1, 1, dt, V, equal_weight=True)
        mb.ntc2(mb.PhiloxRng(t, 1), cfs[1], None, its[1], pvH, pvA, pia, 1, 2, 1, dt, V, equal_weight=True)
        mb.ntc(mb.PhiloxRng(t, 2), cfs[2], None, its[2], pvH, pia, 1, 2, dt, V, equal_weight=True)
        oracle.ntc(oracle.Rng.philox(1234, t, 0), ocfs[0], oits[0], opvA, opia, 1, 1, 1, dt, V, equal_weight=True)
        oracle.ntc2(oracle.Rng.philox(1234, t, 1), ocfs[1], oits[1], opvH, opvA, opia, 1, 1, 2, 1, dt, V, equal_weight=True)
        oracle.ntc(oracle.Rng.philox(1234, t, 2), ocfs[2], oits[2], opvH, opia, 1, 1, 2, dt, V, equal_weight=True)
        for cf, ocf in zip(cfs, ocfs):
            d = cf.download()
            assert d["n_coll"][0] == ocf.n_coll[0] and d["n_coll_performed"][0] == ocf.n_coll_performed[0]
    assert sum(int(o.n_coll_performed[0]) for o in ocfs) > 0
    assert_rows_close(pvA.logical(1, nAr), opvA.logical(1, nAr), 1e-11, "Ar")
    assert_rows_close(pvH.logical(1, nHe), opvH.logical(1, nHe), 1e-11, "He")
    pp = mb.PhysProps(1, 2, ctx=ctx)
    mb.compute_props([pvA, pvH], pia, [AR, HE], pp)
    d, o = pp.download(), oracle.compute_props([opvA, opvH], opia, [AR, HE])
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-11)


def test_ntc_three_species_all_pairs_parity(mb, oracle, ctx):
    """More than two species: the reference's drivers loop ntc! over every species pair of a cell
    (collision_factors[s1, s2, cell], collision_ntc.jl:46-155); here 3 species in 6 cells, variable weight, the six pair calls of a
    step -- (1,1) (2,2) (3,3) one-ParticleVector, (2,1) (3,1) (3,2) two-ParticleVector -- over 4 steps with a sort per species in
    between, device vs oracle draw for draw."""
    rng = np.random.default_rng(33)
    n_cells, L = 6, 6e-5
    g, og = mb.Grid1DUniform(L, n_cells), (L, n_cells)
    masses = [AR, HE, 2.0 * HE]
    temps = [900.0, 360.0, 500.0]
    counts = [300 * n_cells, 800 * n_cells, 500 * n_cells]
    Fnum = 3e14
    opvs, pvs = [], []
    opia = oracle.OPIA(n_cells, 3)
    for s, (m, T, n) in enumerate(zip(masses, temps, counts)):
        rows = maxwellian_rows(rng, n, L, T=T, m=m, w=Fnum, vw=True)
        opv = oracle.OPV(10 * n)  # the split windows are sized by the candidate counts
        if n <= 2000:
            opv.fill_identity(rows)
        else:
            opv.particles[:n] = rows
            opv.index[:n] = np.arange(1, n + 1)
            opv.nbuffer = len(opv) - n
        opia.indexer[s, 0] = (n, 1, n, n, 0, -1, 0)
        opia.n_total[s] = n
        opvs.append(opv)
    for s in range(3):
        oracle.sort_particles(opvs[s], opia, s + 1, grid=og)
    pia = mb.ParticleIndexerArray(n_cells, 3, ctx)
    for s in range(3):
        pv = mb.ParticleVector(len(opvs[s]), ctx)
        pv.set_logical(1, opvs[s].logical(1, counts[s]))
        pvs.append(pv)
    pia.upload(opia.indexer.copy(), opia.n_total.copy(), opia.contiguous.copy())
    pairs = [(1, 1), (2, 2), (3, 3), (2, 1), (3, 1), (3, 2)]
    its, oits, cfs, ocfs = {}, {}, {}, {}
    for a, b in pairs:
        d, o, Tref = 3.5e-10 + 0.2e-10 * (a + b), 0.75 + 0.02 * a, 273.0
        its[a, b] = mb.make_interaction(masses[a - 1], masses[b - 1], d, o, Tref)
        oits[a, b] = oracle.make_interaction(masses[a - 1], masses[b - 1], d, o, Tref)
        s0 = mb.estimate_sigma_g_w_max(its[a, b], masses[a - 1], masses[b - 1], temps[a - 1], temps[b - 1], 2 * Fnum)
        cfs[a, b], ocfs[a, b] = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    dt, V = 2.59e-9, L / n_cells
    for t in range(1, 5):
        for k, (a, b) in enumerate(pairs):
            if a == b:
                mb.ntc(mb.PhiloxRng(t, k), cfs[a, b], None, its[a, b], pvs[a - 1], pia, (1, n_cells), a, dt, V)
                oracle.ntc(oracle.Rng.philox(1234, t, k), ocfs[a, b], oits[a, b], opvs[a - 1], opia, 1, n_cells, a, dt, V)
            else:
                mb.ntc2(mb.PhiloxRng(t, k), cfs[a, b], None, its[a, b], pvs[a - 1], pvs[b - 1], pia, (1, n_cells), a, b, dt, V)
                oracle.ntc2(oracle.Rng.philox(1234, t, k), ocfs[a, b], oits[a, b], opvs[a - 1], opvs[b - 1], opia, 1, n_cells, a, b, dt, V)
            d = cfs[a, b].download()
            np.testing.assert_array_equal(d["n_coll"], ocfs[a, b].n_coll)
            np.testing.assert_array_equal(d["n_coll_performed"], ocfs[a, b].n_coll_performed)
            # the split particles of this pair call go to group 2: fold them back before the next call of the same species
            for sp in {a, b}:
                mb.sort_particles(None, g, pvs[sp - 1], pia, sp)
                oracle.sort_particles(opvs[sp - 1], opia, sp, grid=og)
        assert_same_pia(opia, pia)
    assert sum(int(o.n_coll_performed.sum()) for o in ocfs.values()) > 100
    for s in range(3):
        nt = int(opia.n_total[s])
        assert nt > counts[s]  # splits happened
        a_, b_ = pvs[s].logical(1, nt), opvs[s].logical(1, nt)
        np.testing.assert_array_equal(a_[:, 0], b_[:, 0])
        assert_rows_close(a_, b_, 1e-11, "species %d" % (s + 1))
    pp = mb.PhysProps(n_cells, 3, ctx=ctx)
    mb.compute_props(pvs, pia, masses, pp)
    d, o = pp.download(), oracle.compute_props(opvs, opia, masses)
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-10)
    np.testing.assert_array_equal(d["np"], o.np)


# ------------------------------------------------------------------------------------------------------------ time loop
def test_couette_loop_parity(mb, oracle, ctx):
    """simulations/1D/couette_benchmarking.jl:58-85 order (collide all cells -> convect -> sort -> props), 25 steps on a small
    Couette case, device vs oracle with shared Philox streams: same cell populations, same pia, profiles to 1e-10."""
    n_cells, ppc = 50, 200
    dt = 2.59e-9 * 4  # sigma_v * dt = 0.26 cells: the band (w = 2) holds
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 404)
    it, oit = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0), oracle.interaction("Ar", "Ar")
    s0 = mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum)
    cf, ocf = mb.CollisionFactors(n_cells, s0, ctx), oracle.CF(n_cells, s0)
    g = mb.Grid1DUniform(L, n_cells)
    walls = mb.MaxwellWalls1D(300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
    owalls = (300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    V = L / n_cells
    paths = []
    for t in range(1, 26):
        mb.ntc_equal_weight(mb.PhiloxRng(t), cf, None, it, pv, pia, (1, n_cells), 1, dt, V)
        mb.convect_particles(mb.PhiloxRng(t), g, walls, pv, pia, 1, AR, dt)
        mb.sort_particles(None, g, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], pp)
        oracle.ntc(oracle.Rng.philox(1234, t), ocf, oit, opv, opia, 1, n_cells, 1, dt, V, equal_weight=True)
        oracle.convect_particles(oracle.Rng.philox(1234, t), (L, n_cells), owalls, opv, opia, 1, [AR], dt)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        paths.append(ctx.sort_last_path)
        assert_same_pia(opia, pia)
    assert paths.count(1) >= 20, paths  # the steady-state steps take the band path
    assert_rows_close(pv.logical(1, n), opv.logical(1, n), 1e-10, "couette loop")
    d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
    np.testing.assert_array_equal(d["np"], o.np)
    np.testing.assert_allclose(d["T"], o.T, rtol=1e-10)  # 25 steps of libm-level differences in the particles themselves
    np.testing.assert_allclose(d["v"], o.v, rtol=1e-9, atol=1e-9 * 500)
    # the moments the bench reads (cached by the band sort) against the two-pass oracle ON THE DEVICE'S OWN PARTICLES: 1e-12
    opv2, opia2 = oracle_state(oracle, pv.logical(1, n), n_cells)
    opia2.indexer[:] = pia.indexer
    o2 = oracle.compute_props_sorted([opv2], opia2, [AR])
    np.testing.assert_allclose(d["n"], o2.n, rtol=1e-13)
    np.testing.assert_allclose(d["T"], o2.T, rtol=1e-12)
    np.testing.assert_allclose(d["v"], o2.v, rtol=1e-12, atol=1e-12 * 500)


@pytest.mark.parametrize("w", [1, 2])
def test_cached_moments_meet_1e12(mb, oracle, ctx, w):
    """compute_props_sorted! (physical_props.jl:317-454) from the moments the band sort accumulates on the fly -- the path bench.py
    times -- at ppc = 1000 with a mean velocity of 500 m/s on top of a 250 m/s thermal spread: n to 1e-13, v and T to 1e-12 of the
    two-pass oracle, step after step (stayers summed in the scatter pass, movers and extras added by the combine pass)."""
    rng = np.random.default_rng(40 + w)
    n_cells, ppc, L = 64, 1000, 64e-5
    n = n_cells * ppc
    dx = L / n_cells
    rows = maxwellian_rows(rng, n, L, vw=True)
    rows[:, 1:4] += np.array([500.0, -500.0, 500.0])
    opv, opia = oracle_state(oracle, rows, n_cells)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, n_cells)
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    ctx.set_band_halfwidth(w)
    try:
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        for step in range(4):
            cur = opv.logical(1, n)
            cur[:, 4] = np.clip(cur[:, 4] + rng.normal(0, 0.3 * w * dx, n).clip(-0.95 * w * dx, 0.95 * w * dx), 1e-12, L - 1e-12)
            if step == 3:
                cur[::501, 4] = rng.uniform(1e-12, L - 1e-12, len(cur[::501]))  # a few extras
            opv.set_logical(1, cur)
            pv.set_logical(1, cur)
            mb.sort_particles(None, g, pv, pia, 1)
            oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
            assert ctx.sort_last_path == 1
            l0 = ctx.kernel_launches
            mb.compute_props_sorted([pv], pia, [AR], pp)
            assert ctx.kernel_launches - l0 == 1  # the cached kernel only
            d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
            np.testing.assert_array_equal(d["np"], o.np)
            np.testing.assert_allclose(d["n"], o.n, rtol=1e-13)
            np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-12 * 500)
            np.testing.assert_allclose(d["T"], o.T, rtol=1e-12)
    finally:
        ctx.set_band_halfwidth(2)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n_cells,ppc,w", [(64, 1000, 2), (300, 250, 15), (500, 100, 4), (40, 3000, 1), (900, 9, 2), (7, 5000, 8), (257, 333, 1)])
def test_sort_tile_pass_b(mb, oracle, ctx, cfg, n_cells, ppc, w):
    """Pass B of the band sort as the tile kernel (TMA slices of consecutive old cells, permutation in shared memory, coalesced runs;
    mb_sort_tile.cuh) in every compiled shape: the bit-exact stable counting sort of grid_sorting.jl:58-113 and the cached moments at
    1e-12 for every band width -- tiles of a few big cells, of dozens of small ones (more than a tile's table holds: scattered
    directly), cells that do not fit a tile at all, a few outliers (hybrid extras) and empty cells."""
    import os
    rng = np.random.default_rng(7000 + 13 * cfg + n_cells)
    L = n_cells * 1e-5
    n = n_cells * ppc
    dx = L / n_cells
    rows = maxwellian_rows(rng, n, L, vw=True)
    rows[:, 1:4] += np.array([500.0, -500.0, 500.0])
    if n_cells > 20:
        lo, hi = 5 * dx, 8 * dx   # three empty cells
        inside = (rows[:, 4] >= lo) & (rows[:, 4] < hi)
        rows[inside, 4] = rng.uniform(10 * dx, 12 * dx, int(inside.sum()))
    opv, opia = oracle_state(oracle, rows, n_cells)
    pv, pia = mirror_to_device(mb, ctx, opv, opia)
    g = mb.Grid1DUniform(L, n_cells)
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    ctx.set_band_halfwidth(w)
    old = {k: os.environ.get(k) for k in ("MB_SORT_TILE", "MB_TILE_CFG")}
    os.environ["MB_SORT_TILE"] = "2"
    os.environ["MB_TILE_CFG"] = str(cfg)
    try:
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        for step in range(4):
            cur = opv.logical(1, n)
            sig = (0.35 if step % 2 else 0.1) * w * dx
            cur[:, 4] = np.clip(cur[:, 4] + rng.normal(0, sig, n).clip(-0.95 * w * dx, 0.95 * w * dx), 1e-12, L - 1e-12)
            if step == 2:
                cur[::701, 4] = rng.uniform(1e-12, L - 1e-12, len(cur[::701]))  # a few extras
            opv.set_logical(1, cur)
            pv.set_logical(1, cur)
            mb.sort_particles(None, g, pv, pia, 1)
            oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
            assert ctx.sort_last_path == 1
            assert_same_pia(opia, pia)
            np.testing.assert_array_equal(pv.logical(1, n), opv.logical(1, n))
            assert pia.check(1) == (True, 0)
            l0 = ctx.kernel_launches
            mb.compute_props_sorted([pv], pia, [AR], pp)
            if ppc <= 2048:  # (above that the general path does not fill the cache, and its stand-by kernels are launched as stubs)
                assert ctx.kernel_launches - l0 == 1  # the cached kernel only
            d, o = pp.download(), oracle.compute_props_sorted([opv], opia, [AR])
            np.testing.assert_array_equal(d["np"], o.np)
            np.testing.assert_allclose(d["n"], o.n, rtol=1e-13)
            np.testing.assert_allclose(d["v"], o.v, rtol=1e-12, atol=1e-12 * 500)
            np.testing.assert_allclose(d["T"], o.T, rtol=1e-12, atol=1e-10)  # (atol: a cell of one particle has T = 0 up to round-off)
    finally:
        ctx.set_band_halfwidth(2)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# --------------------------------------------------------------------------------------- fused convect + band classification
@pytest.mark.parametrize("n_cells,ppc,w,dt_mult,acc", [(40, 300, 2, 4, 1.0), (6, 3000, 1, 2, 0.5), (64, 50, 4, 12, 1.0), (3, 700, 8, 30, 0.0),
                                                        (200, 7, 2, 4, 1.0), (64, 50, 1, 24, 1.0), (300, 40, 15, 100, 0.7)])
def test_convect_then_sort_uses_cached_classification(mb, oracle, ctx, n_cells, ppc, w, dt_mult, acc):
    """convect_particles! on a sorted layout classifies while it moves the particles and the following sort_particles! starts at the
    scan: the result must be the bit-exact stable counting sort of the oracle (grid_sorting.jl:58-113), including cells bigger
    than the shared-memory staging (3000 ppc), tiny cells, wide bands and walls of every accommodation."""
    ctx.set_band_halfwidth(w)
    try:
        dt = 2.59e-9 * dt_mult
        L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 900 + n_cells)
        g = mb.Grid1DUniform(L, n_cells)
        walls, owalls = mb.MaxwellWalls1D(300.0, 350.0, -500.0, 500.0, acc, acc), (300.0, 350.0, -500.0, 500.0, acc, acc)
        mb.sort_particles(None, g, pv, pia, 1)  # establishes the sorted layout on the device (general path)
        paths = []
        for t in range(1, 7):
            l0 = ctx.kernel_launches
            mb.convect_particles(mb.PhiloxRng(t), g, walls, pv, pia, 1, AR, dt)
            mb.sort_particles(None, g, pv, pia, 1)
            launches = ctx.kernel_launches - l0
            oracle.convect_particles(oracle.Rng.philox(1234, t), (L, n_cells), owalls, opv, opia, 1, [AR], dt)
            oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
            paths.append(ctx.sort_last_path)
            assert_same_pia(opia, pia)
            assert_rows_close(pv.logical(1, n), opv.logical(1, n), 1e-12, f"fused convect+sort step {t}")
        assert paths.count(1) == len(paths), paths  # outliers (dt_mult 24 with w = 1: every fifth particle) no longer leave the band path
        # clear + convect_band | flag, classify stub, 3 scan, scatter, 3 extras, combine (narrow bands), 8 general-path stubs (7 for small cells);
        # with the tile pass B (cells of 24 .. 2048 particles): 3 tile-setup kernels + the tile kernel, combine + its fallback stub
        # (cells of <= 2048 particles on average: the general path's gather orders the indices of a cell itself, one stub less)
        assert launches == (24 if (24 <= ppc <= 2048 and w >= 4) else (20 if w <= 2 else 19)) - (1 if ppc <= 2048 else 0), launches
        if (w, dt_mult) == (1, 24):
            assert ctx.sort_last_extras > n // 20
    finally:
        ctx.set_band_halfwidth(2)


def test_cached_classification_is_invalidated(mb, oracle, ctx):
    """Anything that touches the particles between convect and sort (here: an upload that moves particles) voids the cache."""
    n_cells, ppc, dt = 30, 200, 2.59e-9 * 4
    L, n, Fnum, opv, opia, pv, pia = _couette_like(oracle, mb, ctx, n_cells, ppc, 77)
    g = mb.Grid1DUniform(L, n_cells)
    walls, owalls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 1.0, 1.0), (300.0, 300.0, 0.0, 0.0, 1.0, 1.0)
    mb.sort_particles(None, g, pv, pia, 1)
    for t in (1, 2, 3):
        mb.convect_particles(mb.PhiloxRng(t), g, walls, pv, pia, 1, AR, dt)
        oracle.convect_particles(oracle.Rng.philox(1234, t), (L, n_cells), owalls, opv, opia, 1, [AR], dt)
        # teleport a few particles (same edit on both sides): the cached classification is now wrong and must not be used
        rows = pv.logical(1, n)
        for i in (5, 1234, n - 3):
            rows[i - 1, 4] = (L - rows[i - 1, 4]) * 0.999 + 1e-9
        pv.set_logical(1, rows)
        opv.set_logical(1, rows)
        mb.sort_particles(None, g, pv, pia, 1)
        oracle.sort_particles(opv, opia, 1, grid=(L, n_cells))
        assert_same_pia(opia, pia)
        assert_rows_close(pv.logical(1, n), opv.logical(1, n), 1e-12, "invalidate")

"""CPU suite: the oracle's initial-condition samplers with the per-cell Philox convention of the device
(oracle/mb_oracle_capi.cpp: mbo_sample_equal_weight_cells / mbo_sample_on_grid_cells), pinned by the reference's own
distribution-level test test/test_sampling.jl:17-63 and by the structural facts of sample_on_grid!
(distributions_and_sampling.jl:312-346: one particle per grid point inside the cut-off sphere, weights sum to n_total)."""
import numpy as np
import pytest

AR = 66.3e-27
K_B = 1.380649e-23


@pytest.mark.parametrize("v0,T0", [((0.0, 0.0, 0.0), 273.0), ((20.0, -10.0, 30.0), 1000.0), ((3000.0, 2000.0, -1000.0), 500.0)])
def test_sample_equal_weight_reference_pins(oracle, v0, T0):
    """test/test_sampling.jl:17-63 with the Philox stream of cell 1."""
    n, n_dens = 20000, 1e20
    Fnum = n_dens / n
    pv, pia = oracle.OPV(n), oracle.OPIA(1, 1)
    oracle.sample_equal_weight_cells(oracle.Rng.philox(1234, 0), pv, pia, 1, 1, 1, n, AR, T0, Fnum, box=(0.0, 0.5, 0.0, 1.0, 0.0, 2.0), v0=v0)
    p = oracle.compute_props([pv], pia, [AR], (4, 6, 8), T0, with_moments=True)
    assert abs((p.n[0, 0] - n_dens) / n_dens) < 4 * np.finfo(float).eps  # Fnum * 20000 summed in fp64
    assert np.all(np.abs(p.v[0, 0] - np.array(v0)) < 10.0)
    assert abs((p.T[0, 0] - T0) / T0) < 1e-2
    assert abs(p.moments[0, 0, 0] - 1.0) < 0.05 and abs(p.moments[0, 0, 1] - 1.0) < 0.05 and abs(p.moments[0, 0, 2] - 1.0) < 0.12
    assert tuple(pia.indexer[0, 0]) == (n, 1, n, n, 0, -1, 0)
    rows = pv.logical(1, n)
    assert rows[:, 4].min() >= 0.0 and rows[:, 4].max() <= 0.5 and rows[:, 5].max() <= 1.0 and rows[:, 6].max() <= 2.0
    assert np.all(pv.cell[:n] == 1)


def test_sample_grid_variants_indexing(oracle):
    """grid_uniform1D.jl:117-219: cells filled in ascending order, appended at n_total + 1; the number-density variant draws
    floor(ndens V / Fnum) or one more."""
    nx, L, ppc = 7, 7e-5, 13
    pv, pia = oracle.OPV(nx * 40), oracle.OPIA(nx, 1)
    oracle.sample_equal_weight_cells(oracle.Rng.philox(7, 0), pv, pia, 1, nx, 1, ppc, AR, 300.0, 1e10, grid=(L, nx))
    for c in range(nx):
        assert tuple(pia.indexer[0, c]) == (ppc, c * ppc + 1, (c + 1) * ppc, ppc, 0, -1, 0)
    rows = pv.logical(1, nx * ppc)
    cells = np.floor(rows[:, 4] * (nx / L)).astype(int)
    assert np.array_equal(cells, np.repeat(np.arange(nx), ppc))
    # number-density variant: expected 20.4 particles per cell
    pv, pia = oracle.OPV(nx * 40), oracle.OPIA(nx, 1)
    Fnum = 1e10
    ndens = 20.4 * Fnum / (L / nx)
    oracle.sample_equal_weight_cells(oracle.Rng.philox(7, 0), pv, pia, 1, nx, 1, -1, AR, 300.0, Fnum, grid=(L, nx), ndens=ndens)
    counts = pia.indexer[0, :, 0]
    assert set(counts.tolist()) <= {20, 21} and pia.n_total[0] == counts.sum()


def test_sample_on_grid_structure(oracle):
    """sample_on_grid! with the BKW vdf at t = 0 (bkw_varweight_octree.jl:62-66 at nv = 12): weights sum to n_total, velocities on
    the LinRange grid, all inside the cut-off sphere; every cell of the ensemble gets the same weights / velocities (noise = 0)."""
    nv, T0, n_dens = 12, 273.0, 1e23
    pv, pia = oracle.OPV(2 * nv ** 3), oracle.OPIA(2, 1)
    n = oracle.sample_on_grid_cells(oracle.Rng.philox(3, 0), "bkw", pv, pia, 1, 2, 1, nv, AR, T0, n_dens)
    assert 0 < n < nv ** 3 and pia.n_total[0] == 2 * n
    assert tuple(pia.indexer[0, 0]) == (n, 1, n, n, 0, -1, 0) and tuple(pia.indexer[0, 1]) == (n, n + 1, 2 * n, n, 0, -1, 0)
    a, b = pv.logical(1, n), pv.logical(n + 1, 2 * n)
    assert np.array_equal(a[:, :4], b[:, :4]) and not np.array_equal(a[:, 4:], b[:, 4:])
    assert abs(a[:, 0].sum() / n_dens - 1.0) < 1e-13 and np.all(a[:, 0] > 0)
    vth = np.sqrt(2 * K_B * T0 / AR)
    t = np.arange(nv) / (nv - 1)
    g = ((1.0 - t) * -1.0 + t * 1.0) * (3.5 * vth)  # Julia's LinRange lerp
    assert np.all(np.isin(a[:, 1], g)) and np.all(np.sqrt((a[:, 1:4] ** 2).sum(1)) <= 3.5 * vth)
    # BKW(t=0) on a grid reproduces the temperature to the quadrature error of the grid
    p = oracle.compute_props([pv], pia, [AR], (4,), T0, with_moments=True)
    assert abs(p.T[0, 0] / T0 - 1.0) < 0.05


@pytest.mark.parametrize("v0,T0", [((0.0, 0.0, 0.0), 273.0), ((20.0, -10.0, 30.0), 1000.0), ((3000.0, 2000.0, -1000.0), 500.0)])
def test_maxwellian_on_grid_reference_pins(oracle, v0, T0):
    """test/test_grid_sampling.jl:19-60: a Maxwellian evaluated on a 20^3 velocity grid of half-width 3.5 v_th (cut-off 8 v_th: the whole
    cube), no velocity noise, streaming velocity v0: n to 1e-13, v to 0.5 m/s, T to 1 %, total moments M4 / M6 / M8 to 1e-4 / 2e-4 / 5e-4
    (quadrature accuracy -- nothing random enters the velocities), positions inside the box."""
    nv, n_dens = 20, 1e20
    pv, pia = oracle.OPV(nv ** 3), oracle.OPIA(1, 1)
    n = int(oracle.sample_on_grid(oracle.Rng.stable(1234), "maxwellian", pv, nv, AR, T0, n_dens, box=(0.0, 0.5, 0.0, 1.0, 0.0, 2.0), v_mult=3.5,
                                  cutoff_mult=8.0, noise=0.0, v_offset=v0))
    assert n == nv ** 3
    pia.set_single_cell(1, 1, n)
    p = oracle.compute_props([pv], pia, [AR], (4, 6, 8), T0, with_moments=True)
    assert abs(p.n[0, 0] / n_dens - 1.0) < 1e-13
    assert np.all(np.abs(p.v[0, 0] - np.array(v0)) < 0.5)
    assert abs(p.T[0, 0] / T0 - 1.0) < 1e-2
    assert abs(p.moments[0, 0, 0] - 1.0) < 1e-4 and abs(p.moments[0, 0, 1] - 1.0) < 2e-4 and abs(p.moments[0, 0, 2] - 1.0) < 5e-4
    rows = pv.logical(1, n)
    assert rows[:, 4].min() >= 0.0 and rows[:, 4].max() <= 0.5 and rows[:, 5].min() >= 0.0 and rows[:, 5].max() <= 1.0
    assert rows[:, 6].min() >= 0.0 and rows[:, 6].max() <= 2.0


def test_grid_1d_uniform_sampling_and_computes(oracle):
    """test/test_grid_1D_uniform.jl:8-130: Grid1DUniform(4.0, 8) -> dx = 0.5; 1000 particles per cell sampled cell after cell (indexer
    ranges 1 + 1000 (i - 1) .. 1000 i, group 2 empty, cell ids written); compute_props! gives n = ppc Fnum exactly, compute_props_sorted!
    with the grid (ndens_not_Np) divides by the cell volume and otherwise agrees to round-off; get_cell of 0.001 / 0.4 / 0.501 / 3.999 is
    1 / 1 / 2 / 8; the number-density variant fills 0.5 * ndens / Fnum = 500 particles per cell."""
    L, nx, ppc, T = 4.0, 8, 1000, 500.0
    g = oracle.grid_params(L, nx)
    assert g["dx"] == 0.5 and g["inv_dx"] == 2.0 and g["L"] == 4.0 and g["n_cells"] == 8
    n_per_cell = 1e10
    Fnum = n_per_cell / ppc
    pv, pia = oracle.OPV(ppc * nx), oracle.OPIA(nx, 1)
    rng = oracle.Rng.stable(1234)
    oracle.sample_equal_weight_cells(rng, pv, pia, 1, nx, 1, ppc, AR, T, Fnum, grid=(L, nx))
    assert pia.n_total[0] == ppc * nx
    for i in range(nx):
        assert tuple(pia.indexer[0, i]) == (ppc, 1 + ppc * i, ppc * (i + 1), ppc, 0, -1, 0)
        assert pv.cell[i * ppc] == i + 1
    p = oracle.compute_props([pv], pia, [AR], Tref=1.0)
    assert np.all(np.abs(p.n[0] - n_per_cell) < 2 * np.finfo(float).eps * n_per_cell) and np.all(p.np[0] == ppc) and p.np.sum() == ppc * nx
    # the reference's bounds (7.5 %, 24 m/s) are 2.3 sigma for its Xoshiro(1234) draw; 4 sigma of the mean of 1000 particles here
    assert np.all(np.abs(p.T[0] - T) / T < 0.105) and np.all(np.abs(p.v[0]) < 41.0)
    q = oracle.compute_props_sorted([pv], pia, [AR], grid=(L, nx))
    assert np.all(np.abs(q.n[0] - n_per_cell / 0.5) <= 2 * np.finfo(float).eps * n_per_cell / 0.5) and np.all(q.np[0] == ppc)
    assert np.all(np.abs(q.T[0] - p.T[0]) / p.T[0] < 2.5 * np.finfo(float).eps) and np.all(np.abs(q.v[0] - p.v[0]) < 1e-13)
    # get_cell (grid_uniform1D.jl:97-99) through the sort
    pv2, pia2 = oracle.OPV(4), oracle.OPIA(nx, 1)
    for i, x in enumerate((0.001, 0.4, 0.501, 3.999)):
        pv2.add_particle(i + 1, 1.0, [0, 0, 0], [x, 0, 0])
    pia2.indexer[0, 0] = (4, 1, 4, 4, 0, -1, 0)
    pia2.n_total[0] = 4
    oracle.sort_particles(pv2, pia2, 1, grid=(L, nx))
    assert list(pia2.indexer[0, :, 0]) == [2, 1, 0, 0, 0, 0, 0, 1]
    # number-density variant: cell volume 0.5, ndens / Fnum = 1000 -> 500 per cell
    pv3, pia3 = oracle.OPV(ppc * nx), oracle.OPIA(nx, 1)
    oracle.sample_equal_weight_cells(rng, pv3, pia3, 1, nx, 1, -1, AR, T, 1e20, grid=(L, nx), ndens=1e23)
    p3 = oracle.compute_props([pv3], pia3, [AR], Tref=1.0)
    assert p3.np.sum() == 0.5 * round(1e23 / 1e20) * nx and np.all(p3.np[0] == 500)
    assert np.all(np.abs(p3.T[0] - T) / T < 0.15) and np.all(np.abs(p3.v[0]) < 58.0)  # 500 per cell: 4 sigma (reference: 10 %, 36 m/s)

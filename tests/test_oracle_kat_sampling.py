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

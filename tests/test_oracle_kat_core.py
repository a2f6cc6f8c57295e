"""Pins the CPU oracle against the known-answer vectors the reference's own tests hold (no RNG dependence).

Each test names the reference test it restates (paths relative to /root/reference/test).
"""
import numpy as np
import pytest

EPS = np.finfo(float).eps


def test_philox_random123_kat(oracle):
    # Random123 v1.14 kat_vectors: philox4x32 10 rounds
    L = oracle.lib()
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in cases:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        L.mbo_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(v) for v in o) == exp


def test_philox_stream_uniform(oracle):
    out = np.empty(100000)
    oracle.lib().mbo_philox_stream_doubles(1234, 1, 0, 7, 42, len(out), out.ctypes.data)
    assert out.min() >= 0.0 and out.max() < 1.0
    assert abs(out.mean() - 0.5) < 5e-3 and abs(out.var() - 1 / 12) < 2e-3


def _reversed_8(oracle):
    pv = oracle.OPV(8)
    pia = oracle.OPIA(2, 1)
    rows = np.zeros((8, 7))
    rows[:, 0] = 2.5e9
    rows[:, 4] = [0.99 * 8.0 * (9.0 - i) / 8 for i in range(1, 9)]
    pv.fill_identity(rows)
    pia.indexer[0, 0] = (4, 1, 4, 4, 0, -1, 0)
    pia.indexer[0, 1] = (4, 5, 8, 4, 0, -1, 0)
    pia.n_total[0] = 8
    return pv, pia


def test_grid_sorting(oracle):
    """test_grid_sorting.jl:26-125,145-235"""
    pv, pia = _reversed_8(oracle)
    oracle.sort_particles(pv, pia, 1, grid=(8.0, 2))
    assert pia.indexer[0, 0].tolist() == [4, 1, 4, 4, 0, -1, 0]
    assert pia.indexer[0, 1].tolist() == [4, 5, 8, 4, 0, -1, 0]
    assert pv.index.tolist() == [5, 6, 7, 8, 1, 2, 3, 4]
    lg = pv.logical(1, 8)
    for i in range(1, 5):
        assert lg[i - 1, 4] == 0.99 * 8.0 * (9.0 - i - 4) / 8
    for i in range(5, 9):
        assert lg[i - 1, 4] == 0.99 * 8.0 * (9.0 - i + 4) / 8
    # finer grid: 4 cells, everything indexed from cell 1 before the sort
    pia4 = oracle.OPIA(4, 1)
    pia4.set_single_cell(1, 1, 8)
    oracle.sort_particles(pv, pia4, 1, grid=(8.0, 4))
    assert pv.index.tolist() == [7, 8, 5, 6, 3, 4, 1, 2]
    for cell in range(1, 5):
        assert pia4.indexer[0, cell - 1].tolist() == [2, 1 + 2 * (cell - 1), 2 + 2 * (cell - 1), 2, 0, -1, 0]
    # uneven split 3/1/0/4
    pv.index[:] = np.arange(1, 9)
    lg = pv.logical(1, 8)
    lg[0:4, 4] = 6.75
    lg[4, 4] = 2.5
    lg[5:8, 4] = 0.5
    pv.set_logical(1, lg)
    oracle.sort_particles(pv, pia4, 1, grid=(8.0, 4))
    assert pv.index.tolist() == [6, 7, 8, 5, 1, 2, 3, 4]
    counts, starts, ends = [3, 1, 0, 4], [1, 4, 0, 5], [3, 4, -1, 8]
    for c in range(4):
        assert pia4.indexer[0, c].tolist() == [counts[c], starts[c], ends[c], counts[c], 0, -1, 0]
    props = oracle.compute_props([pv], pia4, [oracle.MASS["Ar"]])
    assert np.all(np.abs(props.n[0] - 2.5e9 * np.array(counts)) < 2 * EPS)
    assert props.np[0].tolist() == counts
    # cells-known variant
    pv.index[:] = np.arange(1, 9)
    pv.cell[:] = [4, 4, 4, 4, 2, 1, 1, 1]
    oracle.sort_particles(pv, pia4, 1, grid=None)
    assert pv.index.tolist() == [6, 7, 8, 5, 1, 2, 3, 4]
    for c in range(4):
        assert pia4.indexer[0, c].tolist() == [counts[c], starts[c], ends[c], counts[c], 0, -1, 0]


def _four_particles(oracle):
    pv = oracle.OPV(4)
    pia = oracle.OPIA(100, 1)
    pia.set_single_cell(1, 1, 4)
    rows = np.array([
        [1.0, -1.25, -1.5, 4.0, 23.0, -8.0, 7.5],
        [2.0, 11.0, -3.0, 1.0, 49.0, 6.0, -3.0],
        [3.0, -20.0, 0.0, 2.0, 17.0, 1.0, 3.0],
        [4.0, -49.0, -20.0, 13.0, 1.55, -1.0, 9.0],
    ])
    pv.set_logical(1, rows)
    return pv, pia


@pytest.mark.parametrize("compute_cell", [False, True])
def test_convection_specular(oracle, compute_cell):
    """test_convection_1D.jl:1-75 and :330-380"""
    pv, pia = _four_particles(oracle)
    rng = oracle.Rng.seq(1234)
    grid = (50.0, 100)
    walls = (1.0, 1.0, 0.0, 0.0, 0.0, 0.0)
    oracle.convect_particles(rng, grid, walls, pv, pia, 1, [oracle.MASS["Ar"]], 2.0, compute_cell=compute_cell)
    if compute_cell:
        assert pv.cell[:4].tolist() == [42, 59, 47, 8]
        oracle.sort_particles(pv, pia, 1, grid=None)
    else:
        oracle.sort_particles(pv, pia, 1, grid=grid)
    assert pv.index.tolist() == [4, 1, 3, 2]
    lg = pv.logical(1, 4)
    assert np.max(np.abs(lg[0, 4:7] - [3.55, -1.0, 9.0])) < 3.65e-15
    assert lg[0, 1:4].tolist() == [-49.0, -20.0, 13.0] and lg[0, 0] == 4.0
    assert np.max(np.abs(lg[1, 4:7] - [20.5, -8.0, 7.5])) < 2 * EPS
    assert lg[1, 1:4].tolist() == [-1.25, -1.5, 4.0] and lg[1, 0] == 1.0
    assert np.max(np.abs(lg[2, 4:7] - [23.0, 1.0, 3.0])) < 2 * EPS
    assert lg[2, 1:4].tolist() == [20.0, 0.0, 2.0] and lg[2, 0] == 3.0
    assert np.max(np.abs(lg[3, 4:7] - [29.0, 6.0, -3.0])) < 2 * EPS
    assert lg[3, 1:4].tolist() == [-11.0, -3.0, 1.0] and lg[3, 0] == 2.0
    props = oracle.compute_props([pv], pia, [oracle.MASS["Ar"]])
    expect = np.zeros(100)
    expect[[7, 41, 46, 58]] = [4.0, 1.0, 3.0, 2.0]
    assert np.all(np.abs(props.n[0] - expect) < EPS)


def test_convection_diffuse_walls(oracle):
    """test_convection_1D.jl:76-200: half-Maxwellian reflection statistics, accommodation 0.2"""
    n = 10000
    grid = (50.0, 100)
    mass = [oracle.MASS["Ar"]]
    rng = oracle.Rng.seq(1234)
    for philox in (False, True):
        r = oracle.Rng.philox(99, 3) if philox else rng
        pv = oracle.OPV(n)
        pia = oracle.OPIA(100, 1)
        pia.set_single_cell(1, 1, n)
        rows = np.tile(np.array([1e10 / n, -1000.0, 0, 0, 0.999e-4, 0, 0]), (n, 1))
        pv.set_logical(1, rows)
        walls = (2000.0, 500.0, 1100.0, -820.0, 1.0, 1.0)
        oracle.convect_particles(r, grid, walls, pv, pia, 1, mass, 1e-7)
        oracle.sort_particles(pv, pia, 1, grid=grid)
        lg = pv.logical(1, n)
        assert np.all(lg[:, 1] >= 0)
        props = oracle.compute_props([pv], pia, mass)
        assert abs(props.n[0, 0] - 1e10) < 1e-5
        assert abs((props.v[0, 0, 1] - 1100.0) / 1100.0) < 2.25e-2
        assert abs(props.v[0, 0, 2]) < 30.0
        # T of the reflected half-Maxwellian flux: <vx> = sqrt(pi)/2 * sqrt(2kT/m)
        c = np.sqrt(2 * oracle.K_B * 2000.0 / mass[0])
        assert abs(lg[:, 1].mean() / (np.sqrt(np.pi) / 2 * c) - 1) < 0.03
        # right wall
        pv.index[:] = np.arange(1, n + 1)
        rows[:, 1] = 1000.0
        rows[:, 4] = 50.0 - 0.999e-4
        pv.set_logical(1, rows)
        pia = oracle.OPIA(100, 1)
        pia.set_single_cell(1, 1, n)
        oracle.convect_particles(r, grid, walls, pv, pia, 1, mass, 1e-7)
        oracle.sort_particles(pv, pia, 1, grid=grid)
        lg = pv.logical(1, n)
        assert np.all(lg[:, 1] <= 0)
        props = oracle.compute_props([pv], pia, mass)
        assert abs(props.n[0, 99] - 1e10) < 1e-5
        assert abs((props.v[0, 99, 1] + 820.0) / 820.0) < 2.25e-2
        # accommodation 0.2 on a cold wall: ~80 % specular
        pv.index[:] = np.arange(1, n + 1)
        rows[:, 1] = -1000.0
        rows[:, 4] = 0.999e-4
        pv.set_logical(1, rows)
        pia = oracle.OPIA(100, 1)
        pia.set_single_cell(1, 1, n)
        oracle.convect_particles(r, grid, (10.0, 10.0, 1100.0, -820.0, 0.2, 1.0), pv, pia, 1, mass, 1e-7)
        lg = pv.logical(1, n)
        assert np.all(lg[:, 1] >= 0)
        nspec = int(np.sum(np.abs(lg[:, 1] - 1000.0) < 2 * EPS))
        assert abs(nspec - 8000) < 150


def test_convection_noncontiguous(oracle):
    """test_convection_1D.jl:236-290: only particles the pia points to are moved; sort squashes first"""
    grid = (50.0, 100)
    for compute_cell in (False, True):
        pv = oracle.OPV(30)
        pia = oracle.OPIA(100, 1)
        for i in range(1, 31):
            heavy = 11 <= i <= 25
            pv.add_particle(i, 10000.0 if heavy else 1.0, [-1000.0 if heavy else 1.0, 0, 0], [0.75, 0, 0])
        pia.n_total[0] = 15
        pia.indexer[0, 1] = (15, 1, 10, 10, 26, 30, 5)
        pia.contiguous[0] = 0
        props = oracle.compute_props([pv], pia, [oracle.MASS["Ar"]])
        assert abs(props.n[0, 1] - 15.0) < EPS and abs(props.v[0, 1, 0] - 1.0) < EPS
        oracle.convect_particles(oracle.Rng.seq(1), grid, (1.0, 1.0, 0, 0, 0, 0), pv, pia, 1, [oracle.MASS["Ar"]], 1.0, compute_cell=compute_cell)
        oracle.sort_particles(pv, pia, 1, grid=None if compute_cell else grid)
        props = oracle.compute_props([pv], pia, [oracle.MASS["Ar"]])
        assert abs(props.n[0, 3] - 15.0) < EPS
        assert np.all(np.abs(np.delete(props.n[0], 3)) < EPS)
        assert abs(props.v[0, 3, 0] - 1.0) < EPS
        assert pia.contiguous[0] == 1 and pia.check()[0]


def test_phys_props_compute(oracle):
    """test_computes.jl:6-67"""
    n = 2000
    pv = oracle.OPV(n)
    pia = oracle.OPIA(1, 1)
    pia.set_single_cell(1, 1, n)
    Fnum = 1e25 / n
    rows = np.tile(np.array([Fnum, -1.0, 2.0, -4.0, 10.0, 20.0, 30.0]), (n, 1))
    pv.set_logical(1, rows)
    m = [oracle.MASS["Ar"]]
    p = oracle.compute_props([pv], pia, m)
    assert abs(p.np[0, 0] - n) / n <= EPS
    assert abs(p.n[0, 0] - 1e25) / 1e25 < 1e-8
    assert np.all(np.abs(p.v[0, 0] - [-1.0, 2.0, -4.0]) < 1e-8)
    assert abs(p.T[0, 0]) < EPS
    ps = oracle.compute_props_sorted([pv], pia, m)
    assert abs(ps.np[0, 0] - p.np[0, 0]) < EPS and abs(ps.n[0, 0] - p.n[0, 0]) < EPS
    assert np.all(np.abs(ps.v[0, 0] - p.v[0, 0]) < 1e-8) and abs(ps.T[0, 0]) < EPS
    n0 = p.n[0, 0]
    assert abs(oracle.compute_mixed_moment(pv, pia, 1, 1, [0, 0, 0]) - n0) < EPS * n0
    assert abs(oracle.compute_mixed_moment(pv, pia, 1, 1, [1, 1, 0]) / n0 - (-2)) < 1e-11
    assert abs(oracle.compute_mixed_moment(pv, pia, 1, 1, [1, 0, 1]) / n0 - 4) < 1e-11
    assert abs(oracle.compute_mixed_moment(pv, pia, 1, 1, [0, 1, 1]) / n0 - (-8)) < 1e-11
    assert abs(oracle.compute_mixed_moment(pv, pia, 1, 1, [1, 3, 2], sum_scaler=1.0 / n0) - (-128)) < 1e-11


def test_collision_utils(oracle):
    """test_collision_utils.jl:13-40,62-69"""
    it = oracle.interaction("Ar", "Ar")
    assert abs(it[1] - 0.5) < EPS and abs(it[2] - 0.5) < EPS
    p1 = np.array([1e10, 2.0, 1.0, 0.0, 0, 0, 0])
    p2 = np.array([1e10, 0.0, -1.0, -1.0, 0, 0, 0])
    vcom = np.zeros(3)
    g = np.zeros(1)
    L = oracle.lib()
    L.mbo_compute_com_g(it.ctypes.data, p1.ctypes.data, p2.ctypes.data, vcom.ctypes.data, g.ctypes.data)
    assert np.all(np.abs(vcom - [1.0, 0.0, -0.5]) < EPS) and abs(g[0] - 3.0) < EPS
    rng = oracle.Rng.seq(1234)
    for _ in range(50):  # the reference scatters once (tolerance 2.1 eps for its draw); here 50 independent draws, <= 3 ulp(3.0)
        p1[1:4] = [2.0, 1.0, 0.0]
        p2[1:4] = [0.0, -1.0, -1.0]
        L.mbo_scatter_vhs(rng.ref, it.ctypes.data, p1.ctypes.data, p2.ctypes.data)
        L.mbo_compute_com_g(it.ctypes.data, p1.ctypes.data, p2.ctypes.data, vcom.ctypes.data, g.ctypes.data)
        assert np.all(np.abs(vcom - [1.0, 0.0, -0.5]) < 2 * EPS) and abs(g[0] - 3.0) < 6.1 * EPS
    it12 = oracle.interaction("Ar", "He")
    it21 = oracle.interaction("He", "Ar")
    assert abs(it12[1] - it21[2]) < EPS and abs(it12[2] - it21[1]) < EPS and abs(it12[1] + it12[2] - 1.0) < EPS
    # vhs factor: pi d^2 (2 k T_ref / m_r)^(w - 1/2) / Gamma(5/2 - w)
    from math import gamma, pi
    d, o, Tref = oracle.VHS[("Ar", "Ar")]
    m_r = oracle.MASS["Ar"] / 2
    assert abs(it[7] / (pi * d * d * (2 * oracle.K_B * Tref / m_r) ** (o - 0.5) / gamma(2.5 - o)) - 1) < 1e-14
    assert abs(oracle.sigma_vhs(it, 300.0) / (it[7] * 300.0 ** (1 - 2 * o)) - 1) < 1e-14


def test_fp_scale_norm_rands(oracle):
    """test_collision_fp.jl:29-48: exact standardisation of the sampled normals"""
    x, y, z = oracle.scale_norm_rands([1.0, -3.0, 1.0, 2.0], [4.0, 0.5, -2.0, 9.0], [0.0, 1.0, 2.0, 3.5])
    for a in (x, y, z):
        assert abs(a.mean()) < 1e-15
        assert abs(np.sqrt(np.mean(a * a)) - 1.0) < 1e-15

"""CPU suite: the oracle's velocity-grid merging (oracle/mb_oracle_gridmerge.hpp) pinned by the reference's known-answer tests
test/test_merging_grid_indexing.jl:10-62 (grid indices), test/test_merging_grid_merging.jl:47-137 (conservation, empty octants) and
test/test_merging_grid_merging_1D.jl (x clamping of the 1-D variant)."""
import numpy as np

AR = 66.3e-27
K_B = 1.380649e-23


def test_grid_index_reference_kat(oracle):
    """test_merging_grid_indexing.jl:10-62: Nx, Ny, Nz = 4, 3, 2 on [-1, 1]^3; interior cells counted z fastest, outside octants last."""
    mg = oracle.GridMerge(4, 3, 2, 1.0)
    assert mg.Ntotal == 32
    ext = ((-1.0, 1.0),) * 3
    dv = (2.0 / 4, 2.0 / 3, 2.0 / 2)
    counter = 0
    for i in range(4):
        for j in range(3):
            for k in range(2):
                counter += 1
                assert mg.index(ext, (-1 + 1e-5 + i * dv[0], -1 + 1e-5 + j * dv[1], -1 + 1e-5 + k * dv[2])) == counter
    for n, v in enumerate([(-2.0, -0.5, -0.5), (2.0, -0.5, -0.5), (-2.0, 0.5, -0.5), (2.0, 1.5, -0.5), (-2.0, -0.5, 3.5), (2.0, -0.5, 10.5),
                           (-1.1, 0.5, 0.7), (30.0, 0.7, 0.9)], start=1):
        assert mg.index(ext, v) == 32 + n - 8


def _props(oracle, pv, pia):
    p = oracle.compute_props([pv], pia, [AR])
    return p.n[0].copy(), p.v[0].copy(), p.T[0].copy(), p.np[0].copy()


def test_grid_merge_conservation_reference_kat(oracle):
    """test_merging_grid_merging.jl:60-95: 5000 equal-weight particles with a drift, 2x2x2 grid at 3.5 thermal speeds: density exact,
    T and v conserved to the reference's tolerances."""
    n = 5000
    pv, pia = oracle.OPV(n), oracle.OPIA(1, 1)
    oracle.sample_equal_weight_cells(oracle.Rng.philox(1234, 0), pv, pia, 1, 1, 1, n, AR, 300.0, 1e8, box=(0, 1, 0, 1, 0, 1), v0=(2000.0, 500.0, -400.0))
    n0, v0, T0, _ = _props(oracle, pv, pia)
    mg = oracle.GridMerge(2, 2, 2, 3.5)
    bad = oracle.merge_grid_based(oracle.Rng.philox(1234, 1), mg, pv, pia, 1, 1, 1, AR, T_v=[[T0[0], *v0[0]]])
    assert bad == 0
    assert pia.n_total[0] < n and pia.n_total[0] == pia.indexer[0, 0, 0] and pia.n_total[0] <= 2 * 16
    n1, v1, T1, np1 = _props(oracle, pv, pia)
    assert np1[0] == pia.n_total[0]
    assert abs(n1[0] - n0[0]) <= 4 * np.finfo(float).eps * n0[0]
    assert abs(T1[0] - T0[0]) < 3e-12 and np.all(np.abs(v1[0] - v0[0]) < 6.5e-12)


def _octant_particles(weights):
    """create_24_3particles_in_octant (test_merging_grid_merging.jl:4-44)"""
    rows = []
    for octant in range(1, 9):
        for dv in (-0.5, 0.5, 0.0):
            val = 9.0 - octant + dv
            vz = val if octant >= 5 else -val
            vx = -val if octant % 2 == 1 else val
            vy = val if octant in (3, 4, 7, 8) else -val
            rows.append([octant * weights[octant - 1], vx, vy, vz, 1.0, -10.0, 3.0])
    return np.array(rows)


def test_grid_merge_empty_octants_reference_kat(oracle):
    """test_merging_grid_merging.jl:97-137: 24 particles, 3 per octant, one octant with zero weight; 2x2x2 grid with extent 0.5 thermal
    speeds so that everything lands in the outer octants; the zero-weight octant is dropped, n / v / T conserved to 1e-14."""
    from parity_util import oracle_state

    rows = _octant_particles([1.0] * 7 + [0.0])
    pv, pia = oracle_state(oracle, rows, 1)
    n0, v0, T0, _ = _props(oracle, pv, pia)
    mg = oracle.GridMerge(2, 2, 2, 0.5)
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 1), mg, pv, pia, 1, 1, 1, AR, T_v=[[T0[0], *v0[0]]]) == 0
    nt = int(pia.n_total[0])
    assert nt < 24 and tuple(pia.indexer[0, 0]) == (nt, 1, nt, nt, 0, -1, 0)  # e1 == n_total, group 2 empty (:113-126)
    n1, v1, T1, np1 = _props(oracle, pv, pia)
    assert np1[0] == nt and abs(n1[0] - n0[0]) <= np.finfo(float).eps * n0[0]
    assert abs(T1[0] - T0[0]) < 1e-14 * max(T0[0], 1) * 50 and np.all(np.abs(v1[0] - v0[0]) < 1e-14)
    a = pv.logical(1, nt)
    assert np.all(np.isfinite(a)) and np.all(a[:, 0] > 0)


def test_grid_merge_1d_clamps_x(oracle):
    """test_merging_grid_merging_1D.jl:3-133: 4 particles per cell in 2 cells of [0, 1] placed so that mean(x) +- std(x) leaves the
    domain; without a grid the merged particles are outside, the 1-D variant clamps them to [min_x, max_x]."""
    rows = []
    for i, x in zip(range(1, 5), (0.05, 0.05, 0.05, 0.45)):
        rows.append([2.0, 0.5 - i ** 2, -3.0 + i, 4.0 + 0.3 * i, x, 0.0, 0.0])
    for i, x in zip(range(5, 9), (0.55, 0.95, 0.85, 0.99)):
        rows.append([3.0, 0.5 + i ** 2, -3.0 + 2 * i, 4.0 - i, x, 0.0, 0.0])
    rows = np.array(rows)

    def state():
        pv, pia = oracle.OPV(8), oracle.OPIA(2, 1)
        pv.fill_identity(rows)
        pia.indexer[0, 0] = (4, 1, 4, 4, 0, -1, 0)
        pia.indexer[0, 1] = (4, 5, 8, 4, 0, -1, 0)
        pia.n_total[0] = 8
        return pv, pia

    outside = {}
    for with_grid in (False, True):
        pv, pia = state()
        p = oracle.compute_props([pv], pia, [AR])
        Tv = np.column_stack([p.T[0], p.v[0]])
        mg = oracle.GridMerge(1, 1, 1, 5.5)
        # the sign draws decide on which side each merged particle lands: try streams until one pushes a particle out of the domain
        for t in range(1, 40):
            pv, pia = state()
            oracle.merge_grid_based(oracle.Rng.philox(1234, t), mg, pv, pia, 1, 2, 1, AR, T_v=Tv, grid=(1.0, 2) if with_grid else None)
            x = np.concatenate([pv.logical(int(pia.indexer[0, c, 1]), int(pia.indexer[0, c, 2]))[:, 4] for c in range(2)])
            outside[(with_grid, t)] = bool(np.any(x < 0) or np.any(x > 1))
        assert tuple(pia.indexer[0, 0][:4]) == (2, 1, 2, 2) and tuple(pia.indexer[0, 1][:4]) == (2, 5, 6, 2) and pia.contiguous[0] == 0
    assert any(v for (g, t), v in outside.items() if not g)      # unclamped: some draw leaves the domain
    assert not any(v for (g, t), v in outside.items() if g)      # clamped: never


def test_grid_merging_buffer_sorting_reference_kat(oracle):
    """test_merging_grid_buffer_sorting.jl:44-204: one velocity cell at 10 thermal speeds merges 90 particles of two index groups to 2;
    at 1 thermal speed (inner cell + outer octants) to 10; freed slots, pia with a hole, and the sort that closes it."""
    from test_oracle_kat_octree_vhs import _buffer_sorting_state

    m = AR
    rows, pv, pia = _buffer_sorting_state(oracle, 50)
    p = oracle.compute_props([pv], pia, [m])
    Tv = np.column_stack([p.T[0], p.v[0]])
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 1), oracle.GridMerge(1, 1, 1, 10.0), pv, pia, 1, 1, 1, m, T_v=Tv[:1]) == 0
    p = oracle.compute_props([pv], pia, [m])
    assert pia.contiguous[0] == 0 and p.np[0].tolist() == [2.0, 10.0] and abs(p.n[0, 0] - 90.0) < 1e-12 and pv.nbuffer == 88
    assert pv.buffer[:40].tolist() == [100 - i for i in range(40)] and pv.buffer[40:88].tolist() == [50 - i for i in range(48)]
    assert tuple(pia.indexer[0, 0][:4]) == (2, 1, 2, 2) and pia.indexer[0, 0][6] <= 0 and tuple(pia.indexer[0, 1][:4]) == (10, 51, 60, 10)
    oracle.sort_particles(pv, pia, 1, grid=(8.0, 2))
    assert pia.contiguous[0] == 1
    assert tuple(pia.indexer[0, 0]) == (2, 1, 2, 2, 0, -1, 0) and tuple(pia.indexer[0, 1]) == (10, 3, 12, 10, 0, -1, 0)
    rows, pv, pia = _buffer_sorting_state(oracle, 5)
    p = oracle.compute_props([pv], pia, [m])
    Tv = np.column_stack([p.T[0], p.v[0]])
    assert oracle.merge_grid_based(oracle.Rng.philox(1234, 2), oracle.GridMerge(1, 1, 1, 1.0), pv, pia, 1, 1, 1, m, T_v=Tv[:1]) == 0
    p = oracle.compute_props([pv], pia, [m])
    assert pia.contiguous[0] == 0 and p.np[0].tolist() == [10.0, 10.0] and abs(p.n[0, 0] - 90.0) < 1e-12 and pv.nbuffer == 80
    assert pv.buffer[:80].tolist() == [100 - i for i in range(80)]
    assert tuple(pia.indexer[0, 0]) == (10, 1, 5, 5, 16, 20, 5) and tuple(pia.indexer[0, 1][:4]) == (10, 6, 15, 10)

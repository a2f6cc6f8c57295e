"""CPU suite: more of the reference's known-answer tests restated on the oracle --
test/test_octree_sorting.jl (octant numbering, the order-exact 8-way split with its end-filled ranges, two index groups, an empty
octant, nested second split), test/test_octree_bounds_and_splitting.jl (initial bin bounds of the three OctreeInitBin modes, mean
split, bounds recompute) and test/test_collision_vhs.jl / test_collision_vhs_equal_weight.jl (one VHS collision: energy conservation,
the split particle carries the weight difference and the PRE-collision velocity and position)."""
import numpy as np

SIGNS = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]
C_LIGHT = 299_792_458.0


def _in_octant(octant, v_val, w=1.0, v0=(0.0, 0.0, 0.0)):
    return [w, *(np.array(SIGNS[octant - 1], dtype=float) * v_val + np.array(v0)), 0.0, 0.0, 0.0]


def _state(oracle, rows, indexer=None):
    rows = np.array(rows, dtype=float)
    n = rows.shape[0]
    pv, pia = oracle.OPV(n), oracle.OPIA(1, 1)
    for i, r in enumerate(rows):
        pv.add_particle(i + 1, r[0], r[1:4], r[4:7])
    if indexer is None:
        pia.set_single_cell(1, 1, n)
    else:
        pia.indexer[0, 0] = indexer
        pia.n_total[0] = n
    return pv, pia


def _nine(extra_octant, reverse):
    rows = []
    for octant in (range(8, 0, -1) if reverse else range(1, 9)):
        rows.append(_in_octant(octant, 1.0, w=octant))
        if octant == extra_octant:
            rows.append(_in_octant(octant, 2.0, w=octant))
    return rows


def _fifteen_nested(w_inner=0.5):
    """create_15particles_nested (test_octree_sorting.jl:26-57): 8 particles inside octant 3 around (-2, 2, -2), one in every other octant"""
    mia = [12, 15, 2, 14, 13, 7, 8, 10, 5, 6, 1, 9, 4, 3, 11]
    vp = [None] * 15
    i = 0
    for vx in (-1.0, -3.0):
        for vy in (1.0, 3.0):
            for vz in (-3.0, -1.0):
                vp[mia[i] - 1] = [w_inner, vx, vy, vz, 0.0, 0.0, 0.0]
                i += 1
    for octant in range(8, 0, -1):
        if octant != 3:
            vp[mia[i] - 1] = _in_octant(octant, 1.0)
            i += 1
    return vp


def test_compute_octant(oracle):
    """test_octree_sorting.jl:114-140: octant = 1 + (vx > mx) + 2 (vy > my) + 4 (vz > mz)"""
    mid = (-1.0, -1.0, -1.0)
    for o, s in enumerate(SIGNS, start=1):
        assert oracle.compute_octant(2.0 * np.array(s, dtype=float), mid) == o


def test_split_bin_order_reference_kat(oracle):
    """test_octree_sorting.jl:142-205"""
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C)
    pv, pia = _state(oracle, _nine(3, True))  # octants 8, 7, 6, 5, 4, 3, 3, 2, 1
    oc.init(pv, pia, 1, 1)
    assert sorted(oc.particle_indexes_sorted(9)) == list(range(1, 10))
    oc.split_bin(1, pv)
    assert oc.Nbins == 8 and oc.particle_indexes_sorted(9).tolist() == [9, 8, 7, 6, 5, 4, 3, 2, 1]
    # particles 1..9 in octants 4, 5, 6, 7, 8, 3, 3, 2, 1
    rows = _nine(3, False)[4:9] + _nine(3, True)[5:9]
    for indexer in (None, (9, 1, 3, 3, 4, 9, 6)):  # one index group, then two groups (1..3, 4..9)
        pv, pia = _state(oracle, rows, indexer)
        oc.init(pv, pia, 1, 1)
        oc.split_bin(1, pv)
        assert oc.Nbins == 8 and oc.particle_indexes_sorted(9).tolist() == [9, 8, 7, 6, 1, 2, 3, 4, 5]
        assert [oc.bin(i)["start"] for i in range(1, 9)] == [1, 2, 3, 5, 6, 7, 8, 9]
        assert [oc.bin(i)["end"] for i in range(1, 9)] == [1, 2, 4, 5, 6, 7, 8, 9]
    assert [oc.bin(i)["np"] for i in range(1, 9)] == [1, 1, 2, 1, 1, 1, 1, 1]
    assert [oc.bin(i)["w"] for i in range(1, 9)] == [1, 2, 6, 4, 5, 6, 7, 8]
    # one empty octant (5): 7 particles in octants 8, 7, 6, 4, 3, 2, 1
    rows7 = [_in_octant(o, 1.0, w=o) for o in range(8, 0, -1) if o != 5]
    pv, pia = _state(oracle, rows7)
    oc.init(pv, pia, 1, 1)
    oc.split_bin(1, pv)
    assert oc.Nbins == 7 and oc.particle_indexes_sorted(7).tolist() == [7, 6, 5, 4, 3, 2, 1]
    assert [oc.bin(i)["start"] for i in range(1, 8)] == [1, 2, 3, 4, 5, 6, 7] == [oc.bin(i)["end"] for i in range(1, 8)]
    assert [oc.bin(i)["w"] for i in range(1, 8)] == [1, 2, 3, 4, 6, 7, 8]


def test_nested_split_reference_kat(oracle):
    """test_octree_sorting.jl:207-260: 15 particles, 8 of them in octant 3; the first split leaves bin 3 = positions 3..10, the
    second split of bin 3 gives 8 more bins of one particle each"""
    pv, pia = _state(oracle, _fifteen_nested())
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX_SYM)
    oc.init(pv, pia, 1, 1)
    b = oc.bin(1)
    assert (b["start"], b["end"]) == (1, 15)
    np.testing.assert_allclose(b["v_max"], [3.0, 3.0, 3.0], atol=1e-12)
    np.testing.assert_allclose(b["v_min"], [-3.0, -3.0, -3.0], atol=1e-12)
    oc.split_bin(1, pv)
    assert oc.Nbins == 8
    s = oc.particle_indexes_sorted(15)
    assert s[:2].tolist() == [11, 3] and s[10:].tolist() == [4, 9, 1, 6, 5]
    assert (oc.bin(3)["start"], oc.bin(3)["end"]) == (3, 10)
    assert sorted(s[2:10].tolist()) == sorted([12, 15, 2, 14, 13, 7, 8, 10])
    oc.split_bin(3, pv)
    assert oc.Nbins == 15
    assert all(oc.bin(i)["np"] == 1 for i in range(1, 16))
    assert sorted(oc.particle_indexes_sorted(15).tolist()) == list(range(1, 16))
    assert oc.bin(3)["depth"] == 2 and oc.bin(9)["depth"] == 2 and oc.bin(1)["depth"] == 1


def test_init_bounds_and_mean_split_reference_kat(oracle):
    """test_octree_bounds_and_splitting.jl:76-135"""
    rows8 = [_in_octant(o, 1.0, v0=(1.0, 1.0, 1.0)) for o in range(1, 9)]
    pv, pia = _state(oracle, rows8)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C)
    oc.init(pv, pia, 1, 1)
    np.testing.assert_allclose(oc.bin(1)["v_min"], [-C_LIGHT] * 3, atol=1e-12)
    np.testing.assert_allclose(oc.bin(1)["v_max"], [C_LIGHT] * 3, atol=1e-12)
    oc.split_bin(1, pv)
    assert np.max(np.abs(oc.vel_middle)) < 1e-12
    rows15 = _fifteen_nested(w_inner=1.0)
    pv, pia = _state(oracle, rows15)
    oc2 = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX_SYM)
    oc2.init(pv, pia, 1, 1)
    assert np.max(np.abs(oc2.bin(1)["v_min"] + oc2.bin(1)["v_max"])) < 1e-12
    np.testing.assert_allclose(oc2.bin(1)["v_max"], [3.0, 3.0, 3.0], atol=1e-12)
    oc3 = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX)
    oc3.init(pv, pia, 1, 1)
    np.testing.assert_allclose(oc3.bin(1)["v_min"], [-3.0, -1.0, -3.0], atol=1e-11)
    np.testing.assert_allclose(oc3.bin(1)["v_max"], [1.0, 3.0, 1.0], atol=1e-11)
    oc4 = oracle.Octree(oracle.MEAN_SPLIT, oracle.INIT_MINMAX)
    oc4.init(pv, pia, 1, 1)
    oc4.compute_v_mean(1, 15, pv)
    np.testing.assert_allclose(oc4.vel_middle, np.array(rows15)[:, 1:4].mean(0), atol=1e-11)
    # a far outlier moves exactly the bounds it exceeds (bin_bounds_recompute!)
    prev = oc4.bin(1)
    far = [1.0, 120_000.0, -440_000.0, 920_000.0, 0.0, 0.0, 0.0]
    pv.set_logical(15, np.array([far]))
    oc4.bin_bounds_recompute(1, 1, 15, pv)
    b = oc4.bin(1)
    assert b["v_min"][0] == prev["v_min"][0] and b["v_max"][0] == far[1]
    assert b["v_min"][1] == far[2] and b["v_max"][1] == prev["v_max"][1]
    assert b["v_min"][2] == prev["v_min"][2] and b["v_max"][2] == far[3]


def _two(oracle, w1, w2, cap=3):
    pv, pia = oracle.OPV(cap), oracle.OPIA(1, 1)
    pv.add_particle(1, w1, (1.0, 0.0, 0.0), (0.0, 0.0, 0.0))
    pv.add_particle(2, w2, (-1.0, 0.0, 0.0), (0.5, 0.25, 0.125))
    pia.set_single_cell(1, 1, 2)
    return pv, pia


def test_collide_2particles_vhs_reference_kat(oracle):
    """test_collision_vhs.jl: sigma_g_w_max reset to 0 makes the collision certain (R < 1)."""
    m = oracle.MASS["Ar"]
    it = oracle.interaction("Ar", "Ar")
    for t in range(1, 6):
        rng = oracle.Rng.philox(1234, t)
        # equal weights: no split, energy conserved
        pv, pia = _two(oracle, 1.0, 1.0)
        sg, n_perf, n_eq = oracle.collide_2particles_vhs(rng, it, pv, pia, 1, 2)
        a = pv.logical(1, 2)
        assert n_perf == 1 and n_eq == 1 and pia.n_total[0] == 2 and sg > 0
        assert abs(0.5 * m * (a[:, 1:4] ** 2).sum() - 0.5 * m * 2.0) < 1e-14 * m
        np.testing.assert_allclose(a[:, 1:4].sum(0), [0.0, 0.0, 0.0], atol=1e-15)  # momentum
        # w1 > w2: particle 1 is split; the new particle keeps the pre-collision velocity and position
        pv, pia = _two(oracle, 2.0, 1.0)
        sg, n_perf, n_eq = oracle.collide_2particles_vhs(rng, it, pv, pia, 1, 2)
        a = pv.logical(1, 3)
        assert n_perf == 1 and n_eq == 0 and pia.n_total[0] == 3
        assert a[2, 0] == 1.0 and a[2, 1:4].tolist() == [1.0, 0.0, 0.0] and a[2, 4:7].tolist() == [0.0, 0.0, 0.0] and a[0, 0] == 1.0
        assert tuple(pia.indexer[0, 0]) == (3, 1, 2, 2, 3, 3, 1)
        ke0 = 0.5 * m * (2.0 * 1.0 + 1.0 * 1.0)
        assert abs(0.5 * m * (a[:, 0] * (a[:, 1:4] ** 2).sum(1)).sum() - ke0) < 1e-14 * m
        # w1 < w2: particle 2 is split
        pv, pia = _two(oracle, 1.0, 3.5)
        sg, n_perf, n_eq = oracle.collide_2particles_vhs(rng, it, pv, pia, 1, 2)
        a = pv.logical(1, 3)
        assert pia.n_total[0] == 3 and a[2, 0] == 2.5 and a[2, 1:4].tolist() == [-1.0, 0.0, 0.0] and a[2, 4:7].tolist() == [0.5, 0.25, 0.125]
        assert a[1, 0] == 1.0
        assert abs(0.5 * m * (a[:, 0] * (a[:, 1:4] ** 2).sum(1)).sum() - 0.5 * m * 4.5) < 1e-14 * m
        # the equal-weight routine never splits, whatever the weights (test_collision_vhs_equal_weight.jl)
        pv, pia = _two(oracle, 2.0, 1.0)
        sg, n_perf, n_eq = oracle.collide_2particles_vhs(rng, it, pv, pia, 1, 2, equal_weight=True)
        assert n_perf == 1 and n_eq == 1 and pia.n_total[0] == 2 and pv.logical(1, 2)[:, 0].tolist() == [2.0, 1.0]
    # a sigma_g_w_max far above sigma g w makes the collision (almost) impossible: nothing changes
    pv, pia = _two(oracle, 1.0, 1.0)
    sg, n_perf, n_eq = oracle.collide_2particles_vhs(oracle.Rng.philox(1234, 1), it, pv, pia, 1, 2, sigma_g_w_max=1e30)
    assert n_perf == 0 and sg == 1e30 and pv.logical(1, 2)[:, 1].tolist() == [1.0, -1.0]


def test_surface_props_reference_kat(oracle):
    """test/test_surface_props_1D_uniform.jl:33-182: one particle hits each wall (incident, then reflected with a new velocity / weight);
    np, fluxes, force, normal / shear pressure and kinetic energy flux, then the m / (dt A) scaling."""
    eps2 = 2 * np.finfo(float).eps
    left_in, right_in = [3.5, -9.0, 9.0, 0.0, 0.25, 0, 0], [10.5, 10.0, 0.0, 3.0, 3.9, 0, 0]
    left_out, right_out = [3.5, 9.0, 9.0, 0.0, 0.25, 0, 0], [10.0, -1.0, -2.0, -4.0, 3.9, 0, 0]
    rows = [left_in, right_in, left_out, right_out]
    s = oracle.surface_props_kat(rows, [(0, 1, 0), (0, 2, 1)])
    assert s[:, 0].tolist() == [1.0, 1.0] and s[:, 1].tolist() == [3.5, 10.5] and s[:, 2].tolist() == [0.0, 0.0]
    np.testing.assert_allclose(s[0, 3:6], [3.5 * -9, 3.5 * 9, 0.0], atol=eps2 * 40)
    np.testing.assert_allclose(s[1, 3:6], [10.5 * 10, 0.0, 10.5 * 3], atol=eps2 * 110)
    assert abs(s[0, 6] - 9 * 3.5) < eps2 * 40 and abs(s[1, 6] - 10 * 10.5) < eps2 * 110
    np.testing.assert_allclose(s[0, 7:10], [0.0, 9 * 3.5, 0.0], atol=eps2 * 40)
    np.testing.assert_allclose(s[1, 7:10], [0.0, 0.0, 3 * 10.5], atol=eps2 * 40)
    assert abs(s[0, 10] - 0.5 * 3.5 * 162) < eps2 * 300 and abs(s[1, 10] - 0.5 * 10.5 * 109) < eps2 * 600
    ops = [(0, 1, 0), (0, 2, 1), (1, 1, 2), (1, 2, 3)]
    s = oracle.surface_props_kat(rows, ops)
    assert s[:, 0].tolist() == [1.0, 1.0] and s[:, 1].tolist() == [3.5, 10.5] and s[:, 2].tolist() == [-3.5, -10.0]
    np.testing.assert_allclose(s[0, 3:6], [3.5 * -9 - 3.5 * 9, 0.0, 0.0], atol=eps2 * 70)
    np.testing.assert_allclose(s[1, 3:6], [10.5 * 10 + 10.0, 20.0, 10.5 * 3 + 40.0], atol=eps2 * 120)
    assert abs(s[0, 6] - (9 * 3.5 + 9 * 3.5)) < eps2 * 70 and abs(s[1, 6] - (10 * 10.5 + 1 * 10.0)) < eps2 * 120
    np.testing.assert_allclose(s[0, 7:10], [0.0, 0.0, 0.0], atol=eps2 * 40)
    np.testing.assert_allclose(s[1, 7:10], [0.0, 2 * 10.0, 3 * 10.5 + 4 * 10.0], atol=eps2 * 80)
    ke_r = 0.5 * 10.5 * 109 - 0.5 * 10.0 * 21
    assert abs(s[0, 10]) < eps2 * 300 and abs(s[1, 10] - ke_r) < eps2 * 600
    m, dt, ia = oracle.MASS["Ar"], 1e-20, (0.5, 0.25)
    sc = oracle.surface_props_kat(rows, ops, scale=(m, dt, ia))
    f = m * np.array(ia) / dt
    np.testing.assert_allclose(sc[:, 1], f * [3.5, 10.5], rtol=4e-16)
    np.testing.assert_allclose(sc[:, 2], -f * [3.5, 10.0], rtol=4e-16)
    np.testing.assert_allclose(sc[:, 6], f * [63.0, 115.0], rtol=4e-16)
    np.testing.assert_allclose(sc[1, 7:10], f[1] * np.array([0.0, 20.0, 71.5]), rtol=4e-16)
    assert abs(sc[1, 10] - f[1] * ke_r) <= 4e-16 * abs(f[1] * ke_r) and sc[:, 0].tolist() == [1.0, 1.0]


def _chunk_state(oracle):
    """test/test_chunking.jl:6-55: 4 particles of weights 1..4, three in cell 2 and one in cell 4 of a 5-cell grid of length 10"""
    rows = np.array([[i, 0.0, 0.0, 0.0, x, 0.0, 1.0] for i, x in zip((1.0, 2.0, 3.0, 4.0), (3.0, 3.0, 3.0, 7.0))])
    pv, pia = oracle.OPV(4), oracle.OPIA(5, 1)
    pv.fill_identity(rows)
    pia.indexer[0, 1] = (3, 1, 3, 3, 0, -1, 0)
    pia.indexer[0, 3] = (1, 4, 4, 1, 0, -1, 0)
    pia.n_total[0] = 4
    return rows, pv, pia


def test_compute_props_sorted_chunks_reference_kat(oracle):
    """test/test_chunking.jl:57-105: compute_props_sorted! on a cell chunk fills the chunk's cells and leaves the others as they were
    (not reset); the grid variant divides by the cell volume (2.0)."""
    m = oracle.MASS["Ar"]
    rows, pv, pia = _chunk_state(oracle)
    out = oracle.Props(5, 1)
    oracle.compute_props_sorted([pv], pia, [m], 1, 1, out=out)
    assert out.np[0].tolist() == [0.0] * 5
    oracle.compute_props_sorted([pv], pia, [m], 1, 3, out=out)
    assert out.np[0].tolist() == [0.0, 3.0, 0.0, 0.0, 0.0] and out.n[0, 1] == 6.0
    oracle.compute_props_sorted([pv], pia, [m], 3, 4, out=out)
    assert out.np[0].tolist() == [0.0, 3.0, 0.0, 1.0, 0.0] and out.n[0, 3] == 4.0 and out.n[0, 1] == 6.0  # cell 2 is not reset
    out = oracle.Props(5, 1)
    oracle.compute_props_sorted([pv], pia, [m], 1, 3, grid=(10.0, 5), out=out)
    assert out.np[0].tolist() == [0.0, 3.0, 0.0, 0.0, 0.0] and out.n[0, 1] == 3.0
    oracle.compute_props_sorted([pv], pia, [m], 4, 4, grid=(10.0, 5), out=out)
    assert out.np[0].tolist() == [0.0, 3.0, 0.0, 1.0, 0.0] and out.n[0, 1] == 3.0 and out.n[0, 3] == 2.0


def _buffer_sorting_state(oracle, n_gr1):
    """create_particles_and_pia (test_octree_merging_buffer_sorting.jl:4-41): 90 particles of cell 1 in two index groups
    [1, n_gr1] and [n_gr1 + 11, 100], 10 heavy particles of cell 2 in between"""
    rows = []
    for i in range(1, 101):
        heavy = n_gr1 + 1 <= i <= n_gr1 + 10
        rows.append([1000.0 if heavy else 1.0, i - 1.0, i ** 2 + 3.0, np.sqrt(i + 5.0) ** (2 * (i % 2) - 1.0), 5.0 if heavy else 1.0, 0.0, 0.0])
    rows = np.array(rows)
    pv, pia = oracle.OPV(100), oracle.OPIA(2, 1)
    pv.fill_identity(rows)
    pia.indexer[0, 0] = (90, 1, n_gr1, n_gr1, n_gr1 + 11, 100, 90 - n_gr1)
    pia.indexer[0, 1] = (10, n_gr1 + 1, n_gr1 + 10, 10, 0, -1, 0)
    pia.n_total[0] = 100
    return rows, pv, pia


def test_octree_merging_buffer_sorting_reference_kat(oracle):
    """test_octree_merging_buffer_sorting.jl:43-200: post-merge counts (target 6 -> 2 particles, target 16 -> 12), the LIFO buffer of
    freed slots, pia after the merge (hole between the cells, contiguous == false) and after the sort (contiguous again)."""
    m = oracle.MASS["Ar"]
    rows, pv, pia = _buffer_sorting_state(oracle, 50)
    p = oracle.compute_props([pv], pia, [m])
    assert p.np[0].tolist() == [90.0, 10.0] and p.n[0].tolist() == [90.0, 10000.0] and pv.nbuffer == 0 and pia.contiguous[0] == 1
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 1), oc, pv, pia, 1, 1, 1, 6)
    p = oracle.compute_props([pv], pia, [m])
    assert pia.contiguous[0] == 0 and p.np[0].tolist() == [2.0, 10.0]
    assert abs(p.n[0, 0] - 90.0) < 1e-12 and p.n[0, 1] == 10000.0 and pv.nbuffer == 88
    assert pv.buffer[:40].tolist() == [100 - i for i in range(40)]      # group 2 (61..100) freed from the end first
    assert pv.buffer[40:88].tolist() == [50 - i for i in range(48)]     # then group 1 from 50 down to 3
    assert tuple(pia.indexer[0, 0][:4]) == (2, 1, 2, 2) and pia.indexer[0, 0][6] <= 0 and pia.indexer[0, 0][4] <= 0
    assert tuple(pia.indexer[0, 1][:4]) == (10, 51, 60, 10)
    oracle.sort_particles(pv, pia, 1, grid=(8.0, 2))
    p = oracle.compute_props([pv], pia, [m])
    assert pia.contiguous[0] == 1 and p.np[0].tolist() == [2.0, 10.0] and abs(p.n[0, 0] - 90.0) < 1e-12
    assert tuple(pia.indexer[0, 0]) == (2, 1, 2, 2, 0, -1, 0) and tuple(pia.indexer[0, 1]) == (10, 3, 12, 10, 0, -1, 0)
    # 5 particles in group 1, 85 in group 2: target 16 gives 12 post-merge particles, all deletions fit into group 2
    rows, pv, pia = _buffer_sorting_state(oracle, 5)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 2), oc, pv, pia, 1, 1, 1, 16)
    p = oracle.compute_props([pv], pia, [m])
    assert pia.contiguous[0] == 0 and p.np[0].tolist() == [12.0, 10.0] and abs(p.n[0, 0] - 90.0) < 1e-12 and pv.nbuffer == 78
    assert pv.buffer[:78].tolist() == [100 - i for i in range(78)]
    assert tuple(pia.indexer[0, 0]) == (12, 1, 5, 5, 16, 22, 7) and tuple(pia.indexer[0, 1][:4]) == (10, 6, 15, 10)


def test_swpm_collision_factor_estimate_is_the_ntc_estimate_at_unit_weight(oracle):
    """test/test_collision_utils_swpm.jl:53-85: create_collision_factors_swpm_array(pia, interactions, species, T; mult_factor) fills
    sigma_g_max with estimate_sigma_g_w_max at Fnum = 1 -- for one temperature and for a temperature per species (Ar 300 K, He 600 K),
    all four species pairs; the estimate is linear in Fnum and in mult_factor."""
    mA, mH = oracle.MASS["Ar"], oracle.MASS["He"]
    its = {("Ar", "Ar"): (oracle.interaction("Ar", "Ar"), mA, mA), ("He", "He"): (oracle.interaction("He", "He"), mH, mH),
           ("Ar", "He"): (oracle.make_interaction(mA, mH, *oracle.VHS[("Ar", "He")]), mA, mH),
           ("He", "Ar"): (oracle.make_interaction(mH, mA, *oracle.VHS[("Ar", "He")]), mH, mA)}
    for Ts in ({"Ar": 300.0, "He": 300.0}, {"Ar": 300.0, "He": 600.0}):
        for (a, b), (it, m1, m2) in its.items():
            swpm = oracle.estimate_sigma_g_w_max(it, m1, m2, Ts[a], Ts[b], 1.0, 2.0)
            ntc = oracle.estimate_sigma_g_w_max(it, m1, m2, Ts[a], Ts[b], 1.0, 1.0)
            assert swpm > 0 and abs(swpm - 2.0 * ntc) <= 2 * np.finfo(float).eps * swpm
            assert abs(oracle.estimate_sigma_g_w_max(it, m1, m2, Ts[a], Ts[b], 5e12, 2.0) / (5e12 * swpm) - 1.0) < 4e-16
    # symmetric in the pair
    a = oracle.estimate_sigma_g_w_max(its[("Ar", "He")][0], mA, mH, 300.0, 600.0, 1.0)
    b = oracle.estimate_sigma_g_w_max(its[("He", "Ar")][0], mH, mA, 600.0, 300.0, 1.0)
    assert abs(a / b - 1.0) < 1e-15

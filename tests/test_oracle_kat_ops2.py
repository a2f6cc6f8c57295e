"""Known-answer tests that pin the CPU oracle (part 2): the deterministic vectors of the reference's own tests for indexing,
squash_pia!, restore_particle_ordering!, octree N:2 merging, the 1-factorisation and the chunk exchange -- restated from
test/test_indexing.jl, test_pia_contiguous.jl, test_particle_index_sorting.jl, test_octree_merging.jl,
test_octree_merging_1D.jl, test_1_factorization.jl, test_particle_exchange.jl (file:line in each test).  No GPU."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rows(n, fn):
    return np.array([fn(i) for i in range(1, n + 1)], dtype=np.float64)


# ------------------------------------------------------------------------------------------------ golden data files
def test_oracle_constants_match_the_reference_data_files(oracle):
    """data/particles.toml, data/vhs.toml, data/pseudo_maxwell.toml as captured in tests/golden/reference_vectors.json."""
    g = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))
    assert oracle.MASS["Ar"] == g["particles_toml"]["Ar"]["mass"] and oracle.MASS["He"] == g["particles_toml"]["He"]["mass"]
    for table, key in ((oracle.VHS, "vhs_toml"), (oracle.PSEUDO_MAXWELL, "pseudo_maxwell_toml")):
        for (a, b), (d, o, Tref) in table.items():
            e = g[key].get(f"{a},{b}") or g[key][f"{b},{a}"]
            assert (d, o, Tref) == (e["vhs_d"], e["vhs_o"], e["vhs_Tref"]), (key, a, b)


# ------------------------------------------------------------------------------------------------ indexing
def test_indexing_new_particles_and_map_cont_index(oracle):
    """test/test_indexing.jl:1-95"""
    L = oracle.lib()
    pia = oracle.OPIA(1, 1)
    pia.set_single_cell(1, 1, 20)  # ParticleIndexerArray(20)
    assert tuple(pia.indexer[0, 0]) == (20, 1, 20, 20, 0, -1, 0) and pia.n_total[0] == 20
    L.mbo_update_particle_indexer_new_particle(pia.h, 1, 1)
    L.mbo_update_particle_indexer_new_particle(pia.h, 1, 1)
    assert tuple(pia.indexer[0, 0]) == (22, 1, 20, 20, 21, 22, 2) and pia.n_total[0] == 22

    pia.set_single_cell(1, 1, 10)
    pia.n_total[0] += 5
    pia.indexer[0, 0] = (15, 1, 10, 10, 31, 35, 5)
    for i0 in (1, 4, 10):
        assert L.mbo_map_cont_index(pia.h, 1, 1, i0 - 1) == 1 + i0 - 1
    for i0 in (11, 12, 15):
        assert L.mbo_map_cont_index(pia.h, 1, 1, i0 - 1) == 31 + i0 - 10 - 1
    pia.n_total[0] = 10
    pia.indexer[0, 0] = (10, 24, 28, 5, 31, 35, 5)
    for i0 in (1, 4, 5):
        assert L.mbo_map_cont_index(pia.h, 1, 1, i0 - 1) == 24 + i0 - 1
    for i0 in (6, 8, 10):
        assert L.mbo_map_cont_index(pia.h, 1, 1, i0 - 1) == 31 + i0 - 5 - 1
    L.mbo_update_particle_indexer_new_lower_count(pia.h, 1, 1, 8)
    assert tuple(pia.indexer[0, 0]) == (8, 24, 28, 5, 31, 33, 3) and pia.n_total[0] == 8
    L.mbo_update_particle_indexer_new_lower_count(pia.h, 1, 1, 5)
    assert tuple(pia.indexer[0, 0]) == (5, 24, 28, 5, 0, -1, 0) and pia.n_total[0] == 5
    pia.n_total[0] = 10
    pia.indexer[0, 0] = (10, 24, 28, 5, 31, 35, 5)
    L.mbo_update_particle_indexer_new_lower_count(pia.h, 1, 1, 2)
    assert tuple(pia.indexer[0, 0]) == (2, 24, 25, 2, 0, -1, 0) and pia.n_total[0] == 2


# ------------------------------------------------------------------------------------------------ squash_pia!
def test_squash_pia_one_cell(oracle):
    """test/test_pia_contiguous.jl:9-60 (one group) and :63-110 (two groups): delete at the end, squash, density 55 -> 45."""
    L = oracle.lib()
    m = oracle.MASS["Ar"]
    for two_groups in (False, True):
        pv, pia = oracle.OPV(10), oracle.OPIA(1, 1)
        pv.particles[:10] = _rows(10, lambda i: [i, 0, 0, 0, 10.0, 0.0, 1.0])
        pv.nbuffer = 0
        pia.indexer[0, 0] = (10, 1, 5, 5, 6, 10, 5) if two_groups else (10, 1, 10, 10, -1, -1, 0)
        pia.n_total[0] = 10
        p = oracle.compute_props([pv], pia, [m])
        assert p.np[0, 0] == 10 and abs(p.n[0, 0] - 55.0) / 55.0 < 1e-15
        pv.cell[:10] = np.arange(11, 21)
        L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)
        pia.contiguous[0] = 0
        assert pv.nbuffer == 1
        freed = int(pv.buffer[0])
        oracle.squash_pia(pv, pia, 1)
        assert pia.contiguous[0] == 1
        if two_groups:
            assert tuple(pia.indexer[0, 0][[1, 2, 3, 4, 5, 6]]) == (1, 4, 4, 5, 9, 5)
            assert list(pv.cell[:9]) == [11, 12, 13, 14, 16, 17, 18, 19, 20]
            expect_n = 55.0 - 5.0
        else:
            assert pia.indexer[0, 0][3] == 9 and pia.indexer[0, 0][2] == 9 and pia.indexer[0, 0][6] == 0
            assert list(pv.cell[:9]) == [11, 12, 13, 14, 15, 16, 17, 18, 19]
            expect_n = 45.0
        p = oracle.compute_props([pv], pia, [m])
        assert p.np[0, 0] == 9 and abs(p.n[0, 0] - expect_n) / expect_n < 1e-15
        assert freed not in list(pv.index[:9])


# ------------------------------------------------------------------------------------------------ restore_particle_ordering!
@pytest.mark.parametrize("index,buffer,n_used", [
    ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1], 10),
    ([5, 1, 7, 3, 2, 4, 10, 8, 6, 9], [7, 3, 2, 4, 10, 8, 6, 9, 5, 1], 2),
    ([10, 9, 5, 6, 1, 8, 3, 7, 2, 4], [4, 7, 2, 10, 9, 5, 6, 1, 8, 3], 7),
])
def test_restore_particle_ordering(oracle, index, buffer, n_used):
    """test/test_particle_index_sorting.jl:9-100: payloads are permuted so that index == 1:n, the buffer becomes descending."""
    pv = oracle.OPV(10)
    assert list(pv.index) == list(range(1, 11)) and list(pv.buffer) == list(range(10, 0, -1)) and pv.nbuffer == 10
    pv.index[:] = index
    rows = np.zeros((10, 7))
    rows[:n_used] = _rows(n_used, lambda i: [i, i, 2 * i, 3 * i, -i, -4 * i, -5 * i])
    pv.set_logical(1, rows)
    pv.buffer[:] = buffer
    nb = 10 if n_used == 10 else 10 - n_used
    pv.nbuffer = nb
    oracle.restore_particle_ordering(pv)
    assert list(pv.index) == list(range(1, 11))
    assert list(pv.buffer[:nb]) == list(range(10, 10 - nb, -1))
    assert pv.nbuffer == nb
    np.testing.assert_array_equal(pv.logical(1, n_used), rows[:n_used])
    np.testing.assert_array_equal(pv.particles[:n_used], rows[:n_used])  # physically in order now


# ------------------------------------------------------------------------------------------------ octree N:2 merging
SIGNS = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]


def _octant_24(weights=(1.0,) * 8):
    """create_24_3particles_in_octant (test/test_octree_merging.jl:26-44)"""
    rows = []
    for o in range(1, 9):
        for dv in (-0.5, 0.5, 0.0):
            v = np.array(SIGNS[o - 1], dtype=float) * (9.0 - o + dv)
            rows.append([o * weights[o - 1], *v, 1.0, -10.0, 3.0])
    return np.array(rows)


def _state(oracle, rows):
    n = rows.shape[0]
    pv, pia = oracle.OPV(n), oracle.OPIA(1, 1)
    for i, r in enumerate(rows):
        pv.add_particle(i + 1, r[0], r[1:4], r[4:7])
    pia.set_single_cell(1, 1, n)
    return pv, pia


def test_octree_24_particles_bins_and_merge(oracle):
    """test/test_octree_merging.jl:66-163: split at v0 = 0 gives 8 bins of 3 particles with w = 3 i, means (9 - i) * sign, x means
    (1, -10, 3); merging to 16 keeps 8 bins of depth 1, merging to 2 gives one bin, two particles of weight total / 2; n, v, T
    conserved to 1e-14."""
    m = oracle.MASS["Ar"]
    rows = _octant_24()
    pv, pia = _state(oracle, rows)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C)
    oc.init(pv, pia, 1, 1)
    oc.split_bin(1, pv)
    assert oc.Nbins == 8
    for i in range(1, 9):
        b = oc.bin(i)
        assert b["np"] == 3 and b["w"] == 3 * i
        oc.compute_bin_props(i, pv)
        f = oc.full_bin(i)
        np.testing.assert_allclose(f["v_mean"], (9.0 - i) * np.array(SIGNS[i - 1]), atol=1e-14)
        np.testing.assert_allclose(f["x_mean"], [1.0, -10.0, 3.0], atol=1e-14)
    total_w = sum(3 * i for i in range(1, 9))
    p0 = oracle.compute_props([pv], pia, [m], Tref=1.0)
    assert p0.n[0, 0] == total_w

    oc2 = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C)
    rng = oracle.Rng.seq(1234)
    oracle.merge_octree_N2(rng, oc2, pv, pia, 1, 1, 1, 16)
    assert oc2.Nbins == 8 and pia.n_total[0] == 16 == pia.indexer[0, 0, 0]
    assert all(oc2.bin(i)["depth"] == 1 for i in range(1, 9))
    p = oracle.compute_props([pv], pia, [m], Tref=1.0)
    assert p.np[0, 0] == 16 and abs(p.n[0, 0] - p0.n[0, 0]) < np.finfo(float).eps
    assert abs(p.T[0, 0] - p0.T[0, 0]) < 1e-14
    np.testing.assert_allclose(p.v[0, 0], p0.v[0, 0], atol=1e-14)

    oracle.merge_octree_N2(rng, oc2, pv, pia, 1, 1, 1, 2)
    assert oc2.Nbins == 1 and pia.n_total[0] == 2 and oc2.bin(1)["depth"] == 0
    p = oracle.compute_props([pv], pia, [m], Tref=1.0)
    two = pv.logical(1, 2)
    assert two[0, 0] == 0.5 * total_w == two[1, 0]
    assert abs(p.n[0, 0] - p0.n[0, 0]) < np.finfo(float).eps and abs(p.T[0, 0] - p0.T[0, 0]) < 1e-14
    np.testing.assert_allclose(p.v[0, 0], p0.v[0, 0], atol=1e-14)


def test_octree_merge_1d_clamps_positions(oracle):
    """test/test_octree_merging_1D.jl: with a grid the merged particles' x1 is clamped to [min_x, max_x] and the pia of the merged
    cell shrinks while the species becomes non-contiguous."""
    rows = _octant_24()
    rows[:, 4] = np.linspace(1e-9, 4.9e-4, 24)  # spread over the domain so mean +- sigma would leave it
    pv, pia = oracle.OPV(48), oracle.OPIA(2, 1)
    for i, r in enumerate(np.concatenate([rows, rows])):
        pv.add_particle(i + 1, r[0], r[1:4], r[4:7])
    pia.indexer[0, 0] = (24, 1, 24, 24, 0, -1, 0)
    pia.indexer[0, 1] = (24, 25, 48, 24, 0, -1, 0)
    pia.n_total[0] = 48
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_C)
    Lg, nx = 5e-4, 2
    oracle.merge_octree_N2(oracle.Rng.seq(3), oc, pv, pia, 1, 1, 1, 2, grid=(Lg, nx))
    assert tuple(pia.indexer[0, 0][:4]) == (2, 1, 2, 2) and tuple(pia.indexer[0, 1][:4]) == (24, 25, 48, 24)
    assert pia.contiguous[0] == 0 and pia.n_total[0] == 26
    g = oracle.grid_params(Lg, nx)
    x = pv.logical(1, 2)[:, 4]
    assert np.all(x >= g["min_x"]) and np.all(x <= g["max_x"])
    oracle.squash_pia(pv, pia, 1)
    assert tuple(pia.indexer[0, 1][:4]) == (24, 3, 26, 24) and pia.contiguous[0] == 1


# ------------------------------------------------------------------------------------------------ parallel.jl
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 5, 20, 31])
def test_1_factorization(oracle, n):
    """test/test_1_factorization.jl: every pair exactly once, nobody twice in a round, the first rounds are full."""
    fac = oracle.generate_1_factorization(n)
    if n == 2:
        assert len(fac) == 1 and len(fac[0]) == 1 and set(fac[0][0]) == {1, 2}
        return
    if n in (4, 8, 16, 32):  # the reference asserts the optimal round count only for these (test_1_factorization.jl:9-13)
        assert len(fac) == n - 1
    assert len(fac[0]) == n // 2 and len(fac[1]) == n // 2
    seen = set()
    for rnd in fac:
        used = [x for tp in rnd for x in tp]
        assert len(used) == len(set(used))
        for a, b in rnd:
            assert a != b
            seen.add(frozenset((a, b)))
    assert len(seen) == n * (n - 1) // 2


def test_chunk_exchange_moves_particles_to_their_owner(oracle):
    """test/test_particle_exchange.jl + test_particle_resort_after_exchange.jl in miniature: two chunks of a 4-cell grid; after
    convection-like displacement some particles sit in the other chunk's cells; exchange_particles! + sort_particles_after_exchange!
    leave every chunk with exactly the particles of its own cells, sorted by cell, nobody lost."""
    Lg, nx = 4.0, 4
    chunks = oracle.chunks(nx, 2)
    assert chunks == [(1, 2), (3, 4)]
    rng = np.random.default_rng(5)
    pvs, pias = [], []
    allrows = []
    for ci, (lo, hi) in enumerate(chunks):
        n = 30
        rows = np.zeros((n, 7))
        rows[:, 0] = 100 * (ci + 1) + np.arange(n)  # unique ids
        rows[:, 4] = rng.uniform(lo - 1, hi, n)
        rows[::5, 4] = rng.uniform(0, Lg, len(rows[::5]))  # some wander into the other chunk
        pv, pia = oracle.OPV(4 * n), oracle.OPIA(nx, 1)
        for i, r in enumerate(rows):
            pv.add_particle(i + 1, r[0], r[1:4], r[4:7])
        pia.indexer[0, lo - 1] = (n, 1, n, n, 0, -1, 0)
        pia.n_total[0] = n
        oracle.sort_particles(pv, pia, 1, grid=(Lg, nx))
        pvs.append(pv)
        pias.append(pia)
        allrows.append(rows)
    ex = oracle.Exchanger(chunks, nx)
    ex.reset(1)
    ex.reset(2)
    ex.exchange(pvs, pias, 1)
    got = []
    for ci, (lo, hi) in enumerate(chunks):
        ex.sort_after_exchange(pvs[ci], pias[ci], ci + 1, 1)
        nt = int(pias[ci].n_total[0])
        rows = pvs[ci].logical(1, nt)
        cells = np.floor(rows[:, 4] / (Lg / nx)).astype(int) + 1
        assert np.all((cells >= lo) & (cells <= hi)), (ci, cells)
        assert np.all(np.diff(cells) >= 0)
        ok, where = pias[ci].check(1)
        assert ok, where
        got.append(rows)
    ids = np.sort(np.concatenate(got)[:, 0])
    np.testing.assert_array_equal(ids, np.sort(np.concatenate(allrows)[:, 0]))


def test_chunked_varweight_couette_loses_no_particles(oracle):
    """test/test_couette_varweight_octree_chunking.jl:5-139 restated: 50 cells in 4 chunks, each with its own particle vector, indexer,
    octree and StableRNG(1234 + chunk); 400 particles per cell merged to 150 at t = 0; 50 steps of ntc! + merge (180 -> 150, squash after
    every merged cell) + convect + sort per chunk, then exchange_particles! and sort_particles_after_exchange!.  At every step the
    indexers are consistent, no index is used twice, and the total number density is conserved to 4 eps (:137)."""
    m, it = oracle.MASS["Ar"], oracle.interaction("Ar", "Ar")
    T_wall, v_wall, L, ndens, nx, ppc, dt = 300.0, 500.0, 5e-4, 5e22, 50, 400, 2.59e-9
    n_chunks, thr, tgt = 4, 180, 150
    V = L / nx
    Fnum = V * ndens / ppc
    grid, walls = (L, nx), (T_wall, T_wall, -v_wall, v_wall, 1.0, 1.0)
    chunks = oracle.chunks(nx, n_chunks)
    assert chunks == [(1, 13), (14, 26), (27, 38), (39, 50)]  # ChunkSplitters: the first nx mod n chunks are one longer
    rngs = [oracle.Rng.stable(1234 + i) for i in range(n_chunks)]
    pvs = [oracle.OPV(int(ppc * (hi - lo + 1) * 1.5)) for lo, hi in chunks]
    pias = [oracle.OPIA(nx, 1) for _ in chunks]
    ocs = [oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10) for _ in chunks]
    cfs = [oracle.CF(nx, oracle.estimate_sigma_g_w_max(it, m, m, T_wall, T_wall, Fnum)) for _ in chunks]
    ex = oracle.Exchanger(chunks, nx)
    for c, (lo, hi) in enumerate(chunks):
        oracle.sample_equal_weight_grid(rngs[c], grid, pvs[c], pias[c], 1, m, ndens, T_wall, Fnum, lo, hi)
        oracle.merge_octree_N2(rngs[c], ocs[c], pvs[c], pias[c], lo, hi, 1, tgt, grid=grid)
        oracle.squash_pia(pvs[c], pias[c], 1)
        assert np.all(pias[c].indexer[0, lo - 1:hi, 0] <= tgt) and pias[c].n_total[0] == pias[c].indexer[0, :, 0].sum()
    props = oracle.Props(nx, 1)
    moved = 0
    for t in range(50):
        for c, (lo, hi) in enumerate(chunks):
            assert pias[c].check(1) == (True, 0)
            assert oracle.check_unique_index(pvs[c], pias[c], 1) == (True, 0)
        for c, (lo, hi) in enumerate(chunks):
            for cell in range(lo, hi + 1):
                oracle.ntc(rngs[c], cfs[c], it, pvs[c], pias[c], cell, cell, 1, dt, V)
                if pias[c].indexer[0, cell - 1, 0] > thr:
                    oracle.merge_octree_N2(rngs[c], ocs[c], pvs[c], pias[c], cell, cell, 1, tgt, grid=grid)
                    oracle.squash_pia(pvs[c], pias[c], 1)
            oracle.convect_particles(rngs[c], grid, walls, pvs[c], pias[c], 1, [m], dt)
            ex.reset(c + 1)
            oracle.sort_particles(pvs[c], pias[c], 1, grid=grid)
            own = pias[c].indexer[0, lo - 1:hi, 0].sum()
            moved += int(pias[c].n_total[0] - own)  # particles now sitting in cells of other chunks
        ex.exchange(pvs, pias, 1)
        for c, (lo, hi) in enumerate(chunks):
            ex.sort_after_exchange(pvs[c], pias[c], c + 1, 1)
            assert pias[c].n_total[0] == pias[c].indexer[0, lo - 1:hi, 0].sum()  # only own cells are populated afterwards
            oracle.compute_props_sorted([pvs[c]], pias[c], [m], lo, hi, out=props)
        assert abs(props.n.sum() - ndens * L) / (ndens * L) < 4 * np.finfo(float).eps, t
    assert moved > 50  # the exchange did carry particles across chunk borders


def test_particle_vector_index_indirection_and_buffer(oracle):
    """test/test_indexing_particlevector.jl:8-83: pv[i] is particles[index[i]]; a fresh vector has the identity index, zeroed particles
    and cell ids, and a descending buffer of free slots; resize!(+3) grows every array and pushes the new slots on top of the buffer
    (buffer == [13, 12, ..., 1]); add_particle! pops one slot per particle."""
    pv = oracle.OPV(10)
    assert list(pv.index) == list(range(1, 11)) and len(pv) == 10 and np.all(pv.cell == 0)
    rows = np.zeros((10, 7))
    rows[:, 0] = np.arange(1, 11)
    rows[:, 4] = 10.0 - np.arange(1, 11)
    rows[:, 6] = 1.0
    pv.particles[:] = rows
    got = pv.logical(1, 10)
    assert np.array_equal(got[:, 0], np.arange(1, 11)) and np.array_equal(got[:, 4], 10.0 - np.arange(1, 11))
    pv.index[:] = np.arange(10, 0, -1)
    got = pv.logical(1, 10)
    assert np.array_equal(got[:, 0], 11.0 - np.arange(1, 11)) and np.array_equal(got[:, 4], np.arange(1, 11) - 1.0)
    assert np.array_equal(pv.particles[:, 0], np.arange(1, 11))  # the storage itself did not move
    old_nbuffer = pv.nbuffer
    pv.resize(13)
    assert len(pv) == 13 and len(pv.index) == 13 and len(pv.cell) == 13 and len(pv.buffer) == 13 and pv.particles.shape[0] == 13
    assert pv.nbuffer == old_nbuffer + 3
    assert list(pv.buffer) == list(range(13, 0, -1))

    pv = oracle.OPV(3)
    assert np.all(pv.logical(1, 3) == 0.0) and pv.nbuffer == 3
    pv.add_particle(1, 20.0, [2.0, 2.0, 3.0], [11.0, 12.0, 14.0])
    got = pv.logical(1, 3)
    assert list(got[0]) == [20.0, 2.0, 2.0, 3.0, 11.0, 12.0, 14.0] and np.all(got[1:] == 0.0)
    pv.add_particle(2, 40.0, [2.0, 2.0, 3.0], [11.0, 12.0, 14.0])
    got = pv.logical(1, 3)
    assert got[0, 0] == 20.0 and got[1, 0] == 40.0 and np.all(got[2] == 0.0) and pv.nbuffer == 1


def test_particle_buffer_lifo_reference_kat(oracle):
    """test/test_indexing_particlebuffer.jl:15-185: the free-slot buffer after sampling, resize!, and deletions from the end of either
    group -- exact buffer contents and lengths, number density after each deletion, the emptied indexer (0, -1)."""
    L = oracle.lib()
    m = oracle.MASS["Ar"]
    pv, pia = oracle.OPV(10), oracle.OPIA(1, 1)
    assert list(pv.buffer) == list(range(10, 0, -1)) and pv.nbuffer == 10
    rng = oracle.Rng.stable(1234)
    oracle.sample_equal_weight_cell(rng, pv, pia, 1, 1, 6, m, 237.0, 1e10)
    assert np.all(pv.logical(1, 6)[:, 0] == 1e10)
    assert pv.nbuffer == 4 and list(pv.buffer) == list(range(10, 0, -1))
    pv.resize(14)
    assert list(pv.index) == list(range(1, 15)) and pv.nbuffer == 8
    assert list(pv.buffer) == list(range(14, 0, -1))  # new slots go on top: older slots are used up first
    oracle.sample_equal_weight_cell(rng, pv, pia, 1, 1, 3, 1.0, 237.0, 1e6)
    lenbuf = pv.nbuffer
    assert lenbuf == 5
    rows = pv.logical(1, 9)
    assert np.all(rows[:6, 0] == 1e10) and np.all(rows[6:, 0] == 1e6)
    pia.indexer[0, 0] = (9, 1, 6, 6, 7, 9, 3)
    pia.n_total[0] = 9

    def dens():
        p = oracle.compute_props([pv], pia, [m], Tref=1.0)
        return p.np[0, 0], p.n[0, 0]

    assert dens()[0] == 9.0 and abs(dens()[1] / (6e10 + 3e6) - 1) < 1e-15
    L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)
    assert tuple(pia.indexer[0, 0][1:3]) == (1, 5) and pv.nbuffer == lenbuf + 1
    assert list(pv.buffer) == [14, 13, 12, 11, 10, 6, 8, 7, 6, 5, 4, 3, 2, 1]
    assert list(pv.index) == list(range(1, 15))
    assert dens()[0] == 8.0 and abs(dens()[1] / (5e10 + 3e6) - 1) < 1e-15
    L.mbo_delete_particle_end_group2(pv.h, pia.h, 1, 1)
    assert pv.nbuffer == lenbuf + 2 and list(pv.buffer) == [14, 13, 12, 11, 10, 6, 9, 7, 6, 5, 4, 3, 2, 1]
    assert dens()[0] == 7.0 and abs(dens()[1] / (5e10 + 2e6) - 1) < 1e-15
    L.mbo_delete_particle_end_group2(pv.h, pia.h, 1, 1)
    L.mbo_delete_particle_end_group2(pv.h, pia.h, 1, 1)
    assert list(pv.buffer) == [14, 13, 12, 11, 10, 6, 9, 8, 7, 5, 4, 3, 2, 1] and pv.nbuffer == lenbuf + 4
    assert dens() == (5.0, 5e10) and pia.indexer[0, 0][4] == 0 and pia.indexer[0, 0][6] == 0
    L.mbo_delete_particle_end_group2(pv.h, pia.h, 1, 1)  # nothing left in group 2: no change
    assert list(pv.buffer) == [14, 13, 12, 11, 10, 6, 9, 8, 7, 5, 4, 3, 2, 1] and pv.nbuffer == lenbuf + 4
    for _ in range(4):
        L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)
    assert list(pv.buffer) == [14, 13, 12, 11, 10, 6, 9, 8, 7, 5, 4, 3, 2, 1] and pv.nbuffer == lenbuf + 8
    assert dens() == (1.0, 1e10) and tuple(pia.indexer[0, 0]) == (1, 1, 1, 1, 0, -1, 0)
    L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)
    assert pv.nbuffer == 14 and dens() == (0.0, 0.0) and tuple(pia.indexer[0, 0]) == (0, 0, -1, 0, 0, -1, 0)
    L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)  # empty cell: no change
    assert pv.nbuffer == 14 and tuple(pia.indexer[0, 0]) == (0, 0, -1, 0, 0, -1, 0)
    assert list(pv.buffer) == [14, 13, 12, 11, 10, 6, 9, 8, 7, 5, 4, 3, 2, 1]


def test_particle_deletion_with_unsorted_index_reference_kat(oracle):
    """test/test_indexing_particlebuffer.jl:190-310: delete_particle_end_group1! / _group2! / delete_particle_end! / delete_particle!
    on a scrambled index array -- the exact index and buffer arrays after every call (delete_particle! swaps the victim with the last
    particle of its group first), and the density bookkeeping (weights = storage index)."""
    L = oracle.lib()
    m = oracle.MASS["Ar"]
    pv, pia = oracle.OPV(14), oracle.OPIA(1, 1)
    pia.indexer[0, 0] = (7, 1, 3, 3, 7, 10, 4)
    pia.n_total[0] = 7
    index0 = [10, 7, 9, 12, 1, 5, 11, 2, 13, 3, 4, 6, 14, 8]
    pv.index[:] = index0
    pv.buffer[:] = [12, 1, 5, 4, 6, 14, 8, 10, 7, 9, 11, 2, 13, 3]
    pv.nbuffer = 7
    for ind in index0:
        pv.particles[ind - 1] = [float(ind), 0, 0, 0, 0, 0, 0]

    def state():
        p = oracle.compute_props([pv], pia, [m], Tref=1.0)
        return list(pv.index), list(pv.buffer), pv.nbuffer, p.np[0, 0], p.n[0, 0]

    assert state()[3:] == (7.0, 55.0)
    L.mbo_delete_particle_end_group1(pv.h, pia.h, 1, 1)  # storage index 9 leaves
    assert state() == (index0, [12, 1, 5, 4, 6, 14, 8, 9, 7, 9, 11, 2, 13, 3], 8, 6.0, 46.0)
    L.mbo_delete_particle_end_group2(pv.h, pia.h, 1, 1)  # 3 leaves
    assert state() == (index0, [12, 1, 5, 4, 6, 14, 8, 9, 3, 9, 11, 2, 13, 3], 9, 5.0, 43.0)
    L.mbo_delete_particle_end(pv.h, pia.h, 1, 1)  # end of group 2: 13 leaves
    assert state() == (index0, [12, 1, 5, 4, 6, 14, 8, 9, 3, 13, 11, 2, 13, 3], 10, 4.0, 30.0)
    L.mbo_delete_particle(pv.h, pia.h, 1, 1, 1)  # logical 1 (storage 10): swapped with the last of group 1, then dropped
    index1 = [7, 10, 9, 12, 1, 5, 11, 2, 13, 3, 4, 6, 14, 8]
    assert state() == (index1, [12, 1, 5, 4, 6, 14, 8, 9, 3, 13, 10, 2, 13, 3], 11, 3.0, 20.0)
    L.mbo_delete_particle(pv.h, pia.h, 1, 1, 7)  # logical 7 (storage 11)
    index2 = [7, 10, 9, 12, 1, 5, 2, 11, 13, 3, 4, 6, 14, 8]
    assert state() == (index2, [12, 1, 5, 4, 6, 14, 8, 9, 3, 13, 10, 11, 13, 3], 12, 2.0, 9.0)
    L.mbo_delete_particle(pv.h, pia.h, 1, 1, 7)  # logical 7 again (storage 2)
    assert state() == (index2, [12, 1, 5, 4, 6, 14, 8, 9, 3, 13, 10, 11, 2, 3], 13, 1.0, 7.0)
    L.mbo_delete_particle_end(pv.h, pia.h, 1, 1)  # group 2 is empty: end of group 1 (storage 7)
    assert state() == (index2, [12, 1, 5, 4, 6, 14, 8, 9, 3, 13, 10, 11, 2, 7], 14, 0.0, 0.0)


def _two_chunk_setup(oracle, positions, caps, ranges):
    """Two chunks of a 2-cell grid (chunk i owns cell i); particle np of chunk c has w = c, v = (c^3 np, 1 - np, np), x = (pos, 0.5, 1 + c);
    ranges[c] = [(n, start, end) for cell 1, cell 2]."""
    pvs, pias = [], []
    for c in (1, 2):
        pv, pia = oracle.OPV(caps[c - 1]), oracle.OPIA(2, 1)
        for k, x in enumerate(positions[c - 1], start=1):
            pv.add_particle(k, float(c), [c ** 3 * k, -k + 1.0, k], [x, 0.5, 1.0 + c])
        pia.n_total[0] = len(positions[c - 1])
        for cell, (n, s, e) in enumerate(ranges[c - 1]):
            pia.indexer[0, cell] = (n, s, e, n, 0, -1, 0) if n else (0, 0, -1, 0, 0, -1, 0)
        pvs.append(pv)
        pias.append(pia)
    return pvs, pias


def test_particle_exchange_reference_kat(oracle):
    """test/test_particle_exchange.jl:46-330: reset!, and exchange_particles! between two chunks -- nothing to move; equal numbers of
    strangers (pure swap: the exchanger's indexer points at the swapped-in slots); unequal numbers (swap, then push with a resize by
    the shortfall + DELTA_PARTICLES = 256; the sender's freed slots go to its buffer)."""
    ex = oracle.Exchanger([(1, 2), (3, 4)], 4)
    assert ex.indexer.shape == (4, 2, 7)
    ex.indexer[:] = 7
    ex.reset(1)
    ex.reset(2)
    assert np.all(ex.indexer[:, :, 1:] == np.array([0, -1, 0, 0, -1, 0]))  # n_local is not used by the exchanger (:57-70)

    ex = oracle.Exchanger([(1, 1), (2, 2)], 2)
    # scenario 1: every particle already sits in its owner's cell
    pvs, pias = _two_chunk_setup(oracle, [[3.0] * 4, [6.0] * 4], [4, 4], [[(4, 1, 4), (0, 0, -1)], [(0, 0, -1), (4, 1, 4)]])
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    for c in (1, 2):
        rows = pvs[c - 1].logical(1, 4)
        assert len(pvs[c - 1]) == 4 and np.all(rows[:, 0] == c) and np.array_equal(rows[:, 1], c ** 3 * np.arange(1, 5)) and np.all(rows[:, 6] == 1.0 + c)
    # scenario 2: 1,1,2,2 in both chunks -> pure swap
    pvs, pias = _two_chunk_setup(oracle, [[0.0, 1.0, 3.0, 4.0], [-2.0, -1.0, 6.0, 5.0]], [4, 4], [[(2, 1, 2), (2, 3, 4)], [(2, 1, 2), (2, 3, 4)]])
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    for c, (pos, w) in enumerate((([0.0, 1.0, -2.0, -1.0], [1, 1, 2, 2]), ([3.0, 4.0, 6.0, 5.0], [1, 1, 2, 2]))):
        rows = pvs[c].logical(1, 4)
        assert len(pvs[c]) == 4 and list(rows[:, 4]) == pos and list(rows[:, 0]) == w
    I = ex.indexer  # [cell, chunk]: (n_local, start1, end1, n_group1, start2, end2, n_group2)
    assert tuple(I[0, 1, 1:4]) == (3, 4, 2) and tuple(I[0, 0, 1:3]) == (0, -1)  # cell 1 received two particles from chunk 2, in slots 3..4
    assert tuple(I[1, 0, 1:4]) == (1, 2, 2) and tuple(I[1, 1, 1:3]) == (0, -1)
    assert np.all(I[:, :, 4:] == np.array([0, -1, 0]))
    assert tuple(pias[0].indexer[0, 0]) == (2, 1, 2, 2, 0, -1, 0) and tuple(pias[0].indexer[0, 1][:4]) == (0, 0, -1, 0)
    assert tuple(pias[1].indexer[0, 0][:4]) == (0, 0, -1, 0) and tuple(pias[1].indexer[0, 1]) == (2, 3, 4, 2, 0, -1, 0)
    # scenario 3: 1,1,2 and 1,1,1,2 -> one swap, two pushes into chunk 1 (resized), chunk 2 keeps two freed slots
    pvs, pias = _two_chunk_setup(oracle, [[0.0, 1.0, 4.0], [-2.0, -1.0, 1.5, 6.0]], [3, 4], [[(2, 1, 2), (1, 3, 3)], [(3, 1, 3), (1, 4, 4)]])
    assert pvs[0].nbuffer == 0 and pvs[1].nbuffer == 0
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    assert len(pvs[0]) == 5 + 256 and len(pvs[1]) == 4
    r0, r1 = pvs[0].logical(1, 5), pvs[1].logical(1, 4)
    assert list(r0[:, 4]) == [0.0, 1.0, -2.0, -1.0, 1.5] and list(r0[:, 0]) == [1, 1, 2, 2, 2]
    assert list(r1[:, 4]) == [4.0, -1.0, 1.5, 6.0] and list(r1[:, 0]) == [1, 2, 2, 2]  # the sent particles are not erased, only freed
    assert pvs[0].nbuffer == 256 and np.all(pvs[0].buffer[:256] > 5)
    assert pvs[1].nbuffer == 2 and list(pvs[1].buffer[:2]) == [2, 3]
    assert pias[0].n_total[0] == 5 and pias[1].n_total[0] == 4
    assert tuple(pias[0].indexer[0, 0][:4]) == (2, 1, 2, 2) and tuple(pias[0].indexer[0, 1][:4]) == (0, 0, -1, 0)
    assert tuple(pias[1].indexer[0, 0][:4]) == (0, 0, -1, 0) and tuple(pias[1].indexer[0, 1][:4]) == (1, 4, 4, 1)
    I = ex.indexer
    assert tuple(I[0, 1, 1:]) == (3, 3, 1, 4, 5, 2)  # from chunk 2 into cell 1: one swapped (slot 3), two pushed (slots 4..5)
    assert tuple(I[1, 0, 1:]) == (1, 1, 1, 0, -1, 0)  # from chunk 1 into cell 2: one swapped, none pushed
    for i in (0, 1):
        assert tuple(I[i, i, 1:]) == (0, -1, 0, 0, -1, 0)


def test_particle_exchange_three_chunks_reference_kat(oracle):
    """test/test_particle_exchange.jl:540-735: 4 cells in 3 chunks ([1], [2, 3], [4]); chunk contents by cell [1,1,2,3,4,4], [1,3,4],
    [1,2,3,3].  exchange_particles! visits the pairs in 1-factorisation order (1-2, 1-3, 2-3), swapping strangers pairwise and pushing
    the rest: exact particle order, lengths, freed-slot buffers, the senders' indexers and the exchanger's (chunk, cell) ranges."""
    chunks, n_cells = [(1, 1), (2, 3), (4, 4)], 4
    positions = [[0.0, 0.5, 2.0, 3.0, 4.5, 5.0], [1.5, 3.5, 5.5], [-1.0, 2.5, 4.0, 4.5]]
    cells = [[1, 1, 2, 3, 4, 4], [1, 3, 4], [1, 2, 3, 3]]
    np_in_cells = [[2, 1, 1, 2], [1, 0, 1, 1], [1, 1, 2, 0]]
    offsets = [[1, 3, 4, 5], [1, 1, 2, 3], [1, 2, 3, 3]]
    pvs, pias = [oracle.OPV(8), oracle.OPV(4), oracle.OPV(5)], [oracle.OPIA(n_cells, 1) for _ in range(3)]
    for c in range(3):
        for k, (x, cell) in enumerate(zip(positions[c], cells[c]), start=1):
            pvs[c].add_particle(k, float(cell), [c + 1.0, -(c + 1.0), c + 1.0], [x, 0.5, 0.0])
        pias[c].n_total[0] = len(positions[c])
        for cell in range(n_cells):
            n = np_in_cells[c][cell]
            pias[c].indexer[0, cell] = (n, offsets[c][cell], offsets[c][cell] + n - 1, n, 0, -1, 0) if n else (0, 0, -1, 0, 0, -1, 0)
    ex = oracle.Exchanger(chunks, n_cells)
    for c in (1, 2, 3):
        ex.reset(c)
    ex.exchange(pvs, pias, 1)
    assert [int(p.n_total[0]) for p in pias] == [6, 6, 5]  # includes particles that were pushed away
    assert [len(p) for p in pvs] == [8, 6 + 256, 5]
    assert pvs[0].nbuffer == 4 and list(pvs[0].buffer[:4]) == [8, 7, 4, 6]
    assert pvs[1].nbuffer == 256
    assert pvs[2].nbuffer == 2 and list(pvs[2].buffer[:2]) == [3, 4]
    new_positions = [[0.0, 0.5, 1.5, 3.0, -1.0, 5.0], [2.0, 3.5, 2.5, 3.0, 4.0, 4.5], [4.5, 5.5, 4.0, 4.5, 5.0]]
    new_weights = [[1.0, 1.0, 1.0, 3.0, 1.0, 4.0], [2.0, 3.0, 2.0, 3.0, 3.0, 3.0], [4.0, 4.0, 3.0, 3.0, 4.0]]
    new_vx = [[1.0, 1.0, 2.0, 1.0, 3.0, 1.0], [1.0, 2.0, 3.0, 1.0, 3.0, 3.0], [1.0, 2.0, 3.0, 3.0, 1.0]]
    for c in range(3):
        rows = pvs[c].logical(1, len(new_positions[c]))
        assert list(rows[:, 4]) == new_positions[c] and list(rows[:, 0]) == new_weights[c] and list(rows[:, 1]) == new_vx[c]
    kept = [[2, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]]
    for c in range(3):
        for cell in range(n_cells):
            ix = tuple(pias[c].indexer[0, cell])
            n = kept[c][cell]
            assert ix[0] == n and ix[3] == n and ix[4:] == (0, -1, 0)
            if n:
                assert ix[1:3] == (offsets[c][cell], offsets[c][cell] + n - 1)
    I = ex.indexer  # [cell, chunk, (n_local, start1, end1, n_group1, start2, end2, n_group2)]
    E = (0, -1, 0)
    want = {  # (cell, from chunk): (start1, end1, n_group1, start2, end2, n_group2)
        (1, 1): E + E, (1, 2): (3, 3, 1) + E, (1, 3): (5, 5, 1) + E,
        (2, 1): (1, 1, 1) + E, (2, 2): E + E, (2, 3): (3, 3, 1) + E,
        (3, 1): E + (4, 4, 1), (3, 2): E + E, (3, 3): E + (5, 6, 2),
        (4, 1): (1, 1, 1) + (5, 5, 1), (4, 2): (2, 2, 1) + E, (4, 3): E + E,
    }
    for (cell, chunk), w in want.items():
        assert tuple(I[cell - 1, chunk - 1, 1:]) == w, (cell, chunk, tuple(I[cell - 1, chunk - 1, 1:]))
    # every cell's particles (own range + what the exchanger points at) are all there: counts 4, 2, 4, 3 with weight = cell id
    owner = [1, 2, 2, 3]
    for cell in range(1, 5):
        c = owner[cell - 1] - 1
        ix = pias[c].indexer[0, cell - 1]
        got = [pvs[c].logical(i, i)[0, 0] for i in range(ix[1], ix[2] + 1)]
        for ch in range(3):
            s1, e1, _, s2, e2, _ = I[cell - 1, ch, 1:]
            got += [pvs[c].logical(i, i)[0, 0] for i in range(s1, e1 + 1)] + [pvs[c].logical(i, i)[0, 0] for i in range(s2, e2 + 1)]
        assert got == [float(cell)] * [4, 2, 4, 3][cell - 1], (cell, got)


def test_sort_particles_after_exchange_reference_kat(oracle):
    """test/test_particle_resort_after_exchange.jl:3-135 (case 1): the three-chunk exchange above followed by
    sort_particles_after_exchange!: every chunk ends with exactly its own cells populated (n_total 4 / 6 / 3), group 2 empty, cell ranges
    (1..4), (1..2, 3..6), (1..3), weights equal to the cell id, no index used twice."""
    chunks, n_cells = [(1, 1), (2, 3), (4, 4)], 4
    positions = [[0.0, 0.5, 2.0, 3.0, 4.5, 5.0], [1.5, 3.5, 5.5], [-1.0, 2.5, 4.0, 4.5]]
    cells = [[1, 1, 2, 3, 4, 4], [1, 3, 4], [1, 2, 3, 3]]
    np_in_cells = [[2, 1, 1, 2], [1, 0, 1, 1], [1, 1, 2, 0]]
    offsets = [[1, 3, 4, 5], [1, 1, 2, 3], [1, 2, 3, 3]]
    pvs, pias = [oracle.OPV(8), oracle.OPV(4), oracle.OPV(5)], [oracle.OPIA(n_cells, 1) for _ in range(3)]
    for c in range(3):
        for k, (x, cell) in enumerate(zip(positions[c], cells[c]), start=1):
            pvs[c].add_particle(k, float(cell), [c + 1.0, -(c + 1.0), c + 1.0], [x, 0.5, 0.0])
        pias[c].n_total[0] = len(positions[c])
        for cell in range(n_cells):
            n = np_in_cells[c][cell]
            pias[c].indexer[0, cell] = (n, offsets[c][cell], offsets[c][cell] + n - 1, n, 0, -1, 0) if n else (0, 0, -1, 0, 0, -1, 0)
    ex = oracle.Exchanger(chunks, n_cells)
    for c in (1, 2, 3):
        ex.reset(c)
    ex.exchange(pvs, pias, 1)
    for c in (1, 2, 3):
        ex.sort_after_exchange(pvs[c - 1], pias[c - 1], c, 1)
    assert [int(p.n_total[0]) for p in pias] == [4, 6, 3]
    want = [{1: (4, 1, 4)}, {2: (2, 1, 2), 3: (4, 3, 6)}, {4: (3, 1, 3)}]
    for c in range(3):
        assert oracle.check_unique_index(pvs[c], pias[c], 1) == (True, 0)
        for cell in range(1, n_cells + 1):
            ix = tuple(pias[c].indexer[0, cell - 1])
            assert ix[4:] == (0, -1, 0)
            if cell in want[c]:
                n, s, e = want[c][cell]
                assert ix[:4] == (n, s, e, n)
                assert np.all(pvs[c].logical(s, e)[:, 0] == float(cell))
            else:
                assert ix[:4] == (0, 0, -1, 0)


def test_particle_exchange_push_only_reference_kats(oracle):
    """test/test_particle_exchange.jl:332-538: (a) every particle of chunk 2 belongs to chunk 1's cell -- four pushes into a resized chunk 1,
    chunk 2 keeps all four slots in its buffer in order [1, 2, 3, 4]; (b) 3 cells in 2 chunks, chunk 2 empty but pre-allocated: the two
    cell-3 particles of chunk 1 are pushed into slots 1..2 of chunk 2 without a resize, chunk 1 frees slots [4, 5]."""
    E = (0, -1, 0)
    # (a)
    ex = oracle.Exchanger([(1, 1), (2, 2)], 2)
    pvs, pias = _two_chunk_setup(oracle, [[0.0, 1.0], [-2.0, -1.0, 1.5, 0.5]], [2, 4], [[(2, 1, 2), (0, 0, -1)], [(4, 1, 4), (0, 0, -1)]])
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    assert [int(p.n_total[0]) for p in pias] == [6, 4] and [len(p) for p in pvs] == [6 + 256, 4]
    assert pvs[0].nbuffer == 256 and pvs[1].nbuffer == 4 and list(pvs[1].buffer[:4]) == [1, 2, 3, 4]
    r0 = pvs[0].logical(1, 6)
    assert list(r0[:, 4]) == [0.0, 1.0, -2.0, -1.0, 1.5, 0.5] and list(r0[:, 0]) == [1, 1, 2, 2, 2, 2]
    assert tuple(pias[0].indexer[0, 0]) == (2, 1, 2, 2, 0, -1, 0) and tuple(pias[1].indexer[0, 0]) == (0, 0, -1, 0, 0, -1, 0)
    for c in (0, 1):
        assert tuple(pias[c].indexer[0, 1]) == (0, 0, -1, 0, 0, -1, 0)
    I = ex.indexer
    assert tuple(I[0, 1, 1:]) == E + (3, 6, 4)  # cell 1 from chunk 2: nothing swapped, four pushed into slots 3..6
    assert tuple(I[1, 0, 1:]) == E + E and tuple(I[0, 0, 1:]) == E + E and tuple(I[1, 1, 1:]) == E + E
    # (b)
    chunks, n_cells = [(1, 2), (3, 3)], 3
    pvs, pias = [oracle.OPV(5), oracle.OPV(4)], [oracle.OPIA(n_cells, 1), oracle.OPIA(n_cells, 1)]
    for k, (x, cell) in enumerate(zip([0.0, 1.0, 2.0, 6.0, 6.5], [1, 1, 2, 3, 3]), start=1):
        pvs[0].add_particle(k, float(cell), [1.0, -1.0, 1.0], [x, 0.5, 0.0])
    pias[0].n_total[0] = 5
    for cell, (n, off) in enumerate(zip([2, 1, 2], [1, 3, 4])):
        pias[0].indexer[0, cell] = (n, off, off + n - 1, n, 0, -1, 0)
    ex = oracle.Exchanger(chunks, n_cells)
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    assert [int(p.n_total[0]) for p in pias] == [5, 2] and [len(p) for p in pvs] == [5, 4]
    assert pvs[0].nbuffer == 2 and list(pvs[0].buffer[:2]) == [4, 5] and pvs[1].nbuffer == 2
    r0, r1 = pvs[0].logical(1, 5), pvs[1].logical(1, 2)
    assert list(r0[:, 4]) == [0.0, 1.0, 2.0, 6.0, 6.5] and list(r0[:, 0]) == [1, 1, 2, 3, 3]
    assert list(r1[:, 4]) == [6.0, 6.5] and list(r1[:, 0]) == [3, 3] and list(r1[:, 1]) == [1.0, 1.0]
    assert [tuple(pias[0].indexer[0, c][:4]) for c in range(3)] == [(2, 1, 2, 2), (1, 3, 3, 1), (0, 0, -1, 0)]
    assert all(tuple(pias[1].indexer[0, c]) == (0, 0, -1, 0, 0, -1, 0) for c in range(3))
    I = ex.indexer
    for cell in (0, 1):
        for ch in (0, 1):
            assert tuple(I[cell, ch, 1:]) == E + E
    assert tuple(I[2, 1, 1:]) == E + E and tuple(I[2, 0, 1:]) == E + (1, 2, 2)


def test_sort_particles_after_exchange_full_swap_reference_kat(oracle):
    """test/test_particle_resort_after_exchange.jl:159-296 (case 2): chunks [1] and [2, 3, 4] hold exactly each other's particles
    ([2, 3, 4] and [1, 1, 1, 1]); after exchange + re-sort chunk 1 has four particles in cell 1 (resized by the push), chunk 2 one per
    cell in slots 1, 2, 3; writing through the logical view afterwards touches every particle exactly once."""
    chunks, n_cells = [(1, 1), (2, 4)], 4
    positions, cells = [[2.0, 3.0, 4.0], [-1.0, -0.5, 0.5, 1.0]], [[2, 3, 4], [1, 1, 1, 1]]
    np_in_cells, offsets = [[0, 1, 1, 1], [4, 0, 0, 0]], [[0, 1, 2, 3], [1, 0, 0, 0]]
    pvs, pias = [oracle.OPV(3), oracle.OPV(4)], [oracle.OPIA(n_cells, 1), oracle.OPIA(n_cells, 1)]
    for c in range(2):
        for k, (x, cell) in enumerate(zip(positions[c], cells[c]), start=1):
            pvs[c].add_particle(k, float(cell), [c + 1.0, -(c + 1.0), c + 1.0], [x, 0.5, 0.0])
        pias[c].n_total[0] = len(positions[c])
        for cell in range(n_cells):
            n = np_in_cells[c][cell]
            pias[c].indexer[0, cell] = (n, offsets[c][cell], offsets[c][cell] + n - 1, n, 0, -1, 0) if n else (0, 0, -1, 0, 0, -1, 0)
    ex = oracle.Exchanger(chunks, n_cells)
    ex.reset(1); ex.reset(2)
    ex.exchange(pvs, pias, 1)
    for c in (1, 2):
        ex.sort_after_exchange(pvs[c - 1], pias[c - 1], c, 1)
    assert [len(p) for p in pvs] == [4 + 256, 4] and [int(p.n_total[0]) for p in pias] == [4, 3]
    want = [{1: (4, 1, 4)}, {2: (1, 1, 1), 3: (1, 2, 2), 4: (1, 3, 3)}]
    for c in range(2):
        assert oracle.check_unique_index(pvs[c], pias[c], 1) == (True, 0)
        for cell in range(1, n_cells + 1):
            ix = tuple(pias[c].indexer[0, cell - 1])
            assert ix[4:] == (0, -1, 0)
            if cell in want[c]:
                n, s, e = want[c][cell]
                assert ix[:4] == (n, s, e, n) and np.all(pvs[c].logical(s, e)[:, 0] == float(cell))
            else:
                assert ix[:4] == (0, 0, -1, 0)
    # no storage slot is shared between logical positions: rewrite the weights through the logical view and read them back
    for c in range(2):
        nt = int(pias[c].n_total[0])
        rows = pvs[c].logical(1, nt).copy()
        rows[:, 0] = 100.0 * (c + 1) + np.arange(nt)
        pvs[c].set_logical(1, rows)
        assert np.array_equal(pvs[c].logical(1, nt)[:, 0], 100.0 * (c + 1) + np.arange(nt))

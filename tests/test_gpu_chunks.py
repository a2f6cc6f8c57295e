"""Slab exchange on ONE GPU with logical chunks (the reference tests its exchange the same way: chunks looped over serially in one
process, test/test_particle_exchange.jl:82-140, test/test_couette_varweight_octree_chunking.jl:87-138).  Every chunk owns a slab of
the grid in its own context; mb_exchange_chunks runs the very pack / unpack kernels of the NCCL exchange (mb_exchange_slab) with a
device-to-device copy as transport, so pack, drop and the arrival merge of the sort are covered by `pytest -m gpu` on a one-GPU box.
tests/test_gpu_multirank.py runs the same assertions over NCCL when more GPUs are present."""
import numpy as np
import pytest

from parity_util import AR

pytestmark = pytest.mark.gpu

XE_CAP = 8192  # mb_exchange.cu: leavers per direction of the edge exchange


class Chunks:
    def __init__(self, mb, n_chunks, G, rows, cap, seed=1234, band=2, mode=0):
        self.mb, self.n, self.G = mb, n_chunks, G
        self.ctx = [mb.Context(0, seed + (i if seed != 1234 else 0)) for i in range(n_chunks)]
        self.slab = [G.slab(i, n_chunks) for i in range(n_chunks)]
        self.pv, self.pia = [], []
        gcell = np.floor(rows[:, 4] * G.inv_dx).astype(np.int64)
        for i in range(n_chunks):
            c, s = self.ctx[i], self.slab[i]
            c.set_band_halfwidth(band)
            mb.exchange_set_mode(c, mode)
            mine = rows[(gcell >= s.cell_offset) & (gcell < s.cell_offset + s.n_cells)]
            pv, pia = mb.ParticleVector(cap, c), mb.ParticleIndexerArray(s.n_cells, 1, c)
            ix = np.zeros((1, s.n_cells, 7), dtype=np.int64)
            ix[0, :, 2] = -1
            ix[0, :, 5] = -1
            if len(mine):
                pv.set_logical(1, mine)
                ix[0, 0] = (len(mine), 1, len(mine), len(mine), 0, -1, 0)
            pia.upload(ix, np.array([len(mine)]), np.array([1], dtype=np.uint8))
            mb.sort_particles(None, s, pv, pia, 1)
            self.pv.append(pv)
            self.pia.append(pia)

    def exchange(self):
        self.mb.exchange_particles(self.ctx, self.slab, self.pv, self.pia, 1)

    def rows(self):
        return [self.pv[i].logical(1, int(self.pia[i].n_total[0])) for i in range(self.n)]

    def check_sorted(self):
        for i in range(self.n):
            ok, where = self.pia[i].check(1)
            assert ok, (i, where)
            loc = self.pv[i].logical(1, int(self.pia[i].n_total[0]))
            lc = np.floor(loc[:, 4] * self.G.inv_dx).astype(np.int64) - self.slab[i].cell_offset
            assert len(lc) == 0 or (lc.min() >= 0 and lc.max() < self.slab[i].n_cells), "a particle outside the slab survived the sort"
            assert np.all(np.diff(lc) >= 0), "not sorted by cell"
            np.testing.assert_array_equal(self.pia[i].indexer[0, :, 0], np.bincount(lc, minlength=self.slab[i].n_cells))

    def close(self):
        for c in self.ctx:
            c.close()


def _unique_rows(rng, n, L, sigma=300.0):
    rows = np.zeros((n, 7))
    rows[:, 0] = 1.0 + np.arange(n)  # unique weights identify the particles
    rows[:, 1:4] = rng.normal(0, sigma, (n, 3))
    rows[:, 4] = rng.uniform(0, L, n)
    rows[:, 5:7] = rng.uniform(0, 1, (n, 2))
    return rows


def _emulate_chunk_step(oracle, G, slabs, chunk_rows, t, dt):
    """What one step (convect with specular walls -> slab exchange -> sort) must leave in every chunk, order included: the oracle moves
    the particles (convection_1D.jl:130-157), the leavers go to the neighbour, arrivals are appended behind the chunk's own particles
    (the left neighbour's first, each in the sender's logical order) and the stable counting sort (grid_sorting.jl:58-113) is numpy's
    stable argsort.  The reference's chunk exchange orders a cell the same way: own particles, then swapped-in, then pushed-in
    (parallel.jl:467-532)."""
    moved = []
    for rows in chunk_rows:
        n = len(rows)
        opv, opia = oracle.OPV(max(n, 1)), oracle.OPIA(G.n_cells, 1)
        opv.particles[:n] = rows
        opv.nbuffer = 0
        if n:
            opia.indexer[0, 0] = (n, 1, n, n, 0, -1, 0)
        opia.n_total[0] = n
        oracle.convect_particles(oracle.Rng.philox(1234, t), (G.L, G.n_cells), (300.0, 300.0, 0, 0, 0, 0), opv, opia, 1, [AR], dt)
        moved.append(opv.logical(1, n) if n else rows)
    out = []
    for i, s in enumerate(slabs):
        cell = lambda r: np.floor(r[:, 4] * G.inv_dx).astype(np.int64) - s.cell_offset
        parts = [moved[i]]
        if i > 0:
            lr = moved[i - 1]
            parts.append(lr[np.floor(lr[:, 4] * G.inv_dx).astype(np.int64) >= s.cell_offset])
        if i + 1 < len(slabs):
            rr = moved[i + 1]
            parts.append(rr[np.floor(rr[:, 4] * G.inv_dx).astype(np.int64) < s.cell_offset + s.n_cells])
        allp = np.concatenate(parts)
        c = cell(allp)
        keep = (c >= 0) & (c < s.n_cells)
        allp, c = allp[keep], c[keep]
        out.append(allp[np.argsort(c, kind="stable")])
    return out


def _oracle_single_domain(oracle, rows, L, nx, steps, dt):
    n = len(rows)
    opv, opia = oracle.OPV(n), oracle.OPIA(nx, 1)
    opv.particles[:n] = rows
    opv.nbuffer = 0
    opia.indexer[0, 0] = (n, 1, n, n, 0, -1, 0)
    opia.n_total[0] = n
    oracle.sort_particles(opv, opia, 1, grid=(L, nx))
    for t in range(1, steps + 1):
        oracle.convect_particles(oracle.Rng.philox(1234, t), (L, nx), (300.0, 300.0, 0, 0, 0, 0), opv, opia, 1, [AR], dt)
        oracle.sort_particles(opv, opia, 1, grid=(L, nx))
    return opv, opia


@pytest.mark.parametrize("mode", ["full", "edge"])
@pytest.mark.parametrize("n_chunks", [2, 4])
def test_chunk_exchange_matches_single_domain(mb, oracle, n_chunks, mode):
    """convect (specular walls: exact arithmetic) -> exchange -> sort over 2 / 4 chunks == the single-domain oracle run, bit for bit:
    nobody lost, nobody duplicated, every chunk sorted with a correct pia (parallel.jl:281-532)."""
    nx, ppc, steps = 43, 150, 80  # 43 cells: uneven slabs
    L, n = nx * 1e-5, nx * ppc
    rows = _unique_rows(np.random.default_rng(99), n, L)
    G = mb.Grid1DUniform(L, nx)
    ch = Chunks(mb, n_chunks, G, rows, 3 * n, mode=0 if mode == "edge" else 1)
    try:
        walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)
        dt = 0.3e-5 / 300.0  # ~0.3 cells per step at sigma_v
        paths, n_before = [], [int(p.n_total[0]) for p in ch.pia]
        moved = 0
        expect = ch.rows()
        for t in range(1, steps + 1):
            expect = _emulate_chunk_step(oracle, G, ch.slab, expect, t, dt)
            for i in range(n_chunks):
                mb.convect_particles(mb.PhiloxRng(t), ch.slab[i], walls, ch.pv[i], ch.pia[i], 1, AR, dt)
            ch.exchange()
            for i in range(n_chunks):
                mb.sort_particles(None, ch.slab[i], ch.pv[i], ch.pia[i], 1)
                paths.append(ch.ctx[i].sort_last_path)
            n_now = [int(p.n_total[0]) for p in ch.pia]
            moved += sum(abs(a - b) for a, b in zip(n_now, n_before))
            n_before = n_now
            assert sum(n_now) == n
            if t % 20 == 0 or t < 4:  # logical order of every chunk, bit for bit
                for got, want in zip(ch.rows(), expect):
                    np.testing.assert_array_equal(got, want)
        assert paths.count(1) >= len(paths) - 2 * n_chunks, paths  # band path with drops + arrivals
        ch.check_sorted()
        allrows = np.concatenate(ch.rows())
        assert allrows.shape[0] == n, "particles lost or duplicated"
        assert moved > 50, "the test did not exchange anything"
        opv, opia = _oracle_single_domain(oracle, rows, L, nx, steps, dt)
        ref = opv.logical(1, n)
        np.testing.assert_array_equal(allrows[np.argsort(allrows[:, 0])], ref[np.argsort(ref[:, 0])])
        # the chunks' cells side by side have the single-domain populations (the order inside a cell differs by construction: arrivals
        # queue behind the cell's own particles, as in the reference's chunked runs)
        np.testing.assert_array_equal(np.concatenate([p.indexer[0, :, 0] for p in ch.pia]), opia.indexer[0, :, 0])
    finally:
        ch.close()


@pytest.mark.parametrize("n_chunks", [2, 4])
def test_chunk_couette_step_conserves_population(mb, n_chunks):
    """the full Couette step (collide -> convect with diffuse walls -> exchange -> sort -> props) over chunks"""
    nx, ppc = 48, 200
    L, n = nx * 1e-5, nx * ppc
    rows = _unique_rows(np.random.default_rng(5), n, L, sigma=250.0)
    rows[:, 0] = 1e12
    G = mb.Grid1DUniform(L, nx)
    ch = Chunks(mb, n_chunks, G, rows, 3 * n, seed=777)
    try:
        it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
        walls = mb.MaxwellWalls1D(300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
        cf = [mb.CollisionFactors(s.n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 1e12), c) for s, c in zip(ch.slab, ch.ctx)]
        pp = [mb.PhysProps(s.n_cells, 1, ctx=c) for s, c in zip(ch.slab, ch.ctx)]
        dt = 2.59e-9 * 4
        for t in range(1, 31):
            for i in range(n_chunks):
                r = mb.PhiloxRng(t, i)
                mb.ntc_equal_weight(r, cf[i], None, it, ch.pv[i], ch.pia[i], (1, ch.slab[i].n_cells), 1, dt, ch.slab[i].dx)
                mb.convect_particles(r, ch.slab[i], walls, ch.pv[i], ch.pia[i], 1, AR, dt)
            ch.exchange()
            for i in range(n_chunks):
                mb.sort_particles(None, ch.slab[i], ch.pv[i], ch.pia[i], 1)
                mb.compute_props_sorted([ch.pv[i]], ch.pia[i], [AR], pp[i])
        assert sum(int(p.n_total[0]) for p in ch.pia) == n
        assert sum(int(p.download()["np"].sum()) for p in pp) == n
        ch.check_sorted()
    finally:
        ch.close()


def test_chunk_variable_weight_loop_conserves_weight_and_energy(mb):
    """C4 over chunks (couette_multithreaded_varweight_octree.jl; test_couette_varweight_octree_chunking.jl:87-138): ntc! with splits ->
    merge_octree_N2_based! above the threshold -> squash_pia! -> convect (specular) -> exchange -> sort.  Every operator conserves weight
    and kinetic energy, so the global sums stay put: 4 eps per step on the weight (the reference's bar, :137), 1e-11 on the energy."""
    n_chunks, nx, ppc = 3, 60, 160
    G = mb.Grid1DUniform(nx * 1e-5, nx)
    ctx = [mb.Context(0, 4321 + i) for i in range(n_chunks)]
    for c in ctx:
        c.set_band_halfwidth(4)  # the edge exchange looks at the w cells next to a face: sigma_v dt = 0.5 cells here, 4 cells = 8 sigma
    try:
        slab = [G.slab(i, n_chunks) for i in range(n_chunks)]
        pv = [mb.ParticleVector(6 * s.n_cells * ppc, c) for s, c in zip(slab, ctx)]
        pia = [mb.ParticleIndexerArray(s.n_cells, 1, c) for s, c in zip(slab, ctx)]
        it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
        oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
        Fnum = 1e-5 * 5e22 / ppc
        for i in range(n_chunks):
            mb.sample_particles_equal_weight(mb.PhiloxRng(0), slab[i], pv[i], pia[i], 1, AR, ppc, 300.0, Fnum)
            mb.merge_octree_N2_based(mb.PhiloxRng(0), oc, pv[i], pia[i], (1, slab[i].n_cells), 1, 100, slab[i], threshold=130)
            mb.squash_pia(pv[i], pia[i], 1)
        cf = [mb.CollisionFactors(s.n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, 1e-5 * 5e22 / 100), c) for s, c in zip(slab, ctx)]
        walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)

        def totals():
            w = e = 0.0
            cnt = 0
            for i in range(n_chunks):
                a = pv[i].logical(1, int(pia[i].n_total[0]))
                w += a[:, 0].sum()
                e += (a[:, 0] * (a[:, 1:4] ** 2).sum(1)).sum()
                cnt += len(a)
            return w, e, cnt

        w0, e0, _ = totals()
        merged = 0
        dt = 2.59e-9 * 8
        for t in range(1, 31):
            for i in range(n_chunks):
                r = mb.PhiloxRng(t, i)
                mb.ntc(r, cf[i], None, it, pv[i], pia[i], (1, slab[i].n_cells), 1, dt, slab[i].dx)
                merged += int((pia[i].indexer[0, :, 0] > 130).sum())
                mb.merge_octree_N2_based(r, oc, pv[i], pia[i], (1, slab[i].n_cells), 1, 100, slab[i], threshold=130)
                if t % 2 == 0:  # the exchange takes the non-contiguous layout a merge leaves as well as the squashed one
                    mb.squash_pia(pv[i], pia[i], 1)
                mb.convect_particles(r, slab[i], walls, pv[i], pia[i], 1, AR, dt)
            mb.exchange_particles(ctx, slab, pv, pia, 1)
            for i in range(n_chunks):
                mb.sort_particles(None, slab[i], pv[i], pia[i], 1)
                ok, where = pia[i].check(1)
                assert ok, (i, t, where)
            if t % 10 == 0:
                w1, e1, _ = totals()
                assert abs(w1 - w0) <= 4e-16 * t * w0 * 8, (t, w1, w0)
                assert abs(e1 - e0) <= 1e-11 * e0, (t, e1, e0)
        assert merged > 20
    finally:
        for c in ctx:
            c.close()


@pytest.mark.parametrize("leavers", [XE_CAP, XE_CAP + 1])
def test_edge_exchange_at_its_capacity(mb, leavers):
    """The edge exchange moves at most XE_CAP particles per direction and step in fixed-size messages.  Exactly XE_CAP leavers through
    one face arrive complete and in order; one more is reported as MB_ERR_CAPACITY by the next synchronising call instead of being
    lost silently."""
    nx, L = 8, 8.0
    G = mb.Grid1DUniform(L, nx)
    n = 4 * leavers + 1000
    rng = np.random.default_rng(leavers)
    rows = _unique_rows(rng, n, L, sigma=0.0)
    rows[:, 4] = rng.uniform(0.05, 3.5, n)               # everything in chunk 0 (cells 0..3) ...
    rows[:leavers, 4] = rng.uniform(3.6, 3.95, leavers)  # ... the leavers in its last cell, about to cross the face at x = 4
    rows[:leavers, 1] = 1.0
    ch = Chunks(mb, 2, G, rows, 2 * n, mode=0)
    try:
        walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)
        for i in range(2):
            mb.convect_particles(mb.PhiloxRng(1), ch.slab[i], walls, ch.pv[i], ch.pia[i], 1, AR, 0.5)
        ch.exchange()
        if leavers > XE_CAP:
            with pytest.raises(mb.MerzbildError) as e:
                ch.ctx[0].sync()
            assert e.value.status == mb.MB_ERR_CAPACITY
            return
        for i in range(2):
            mb.sort_particles(None, ch.slab[i], ch.pv[i], ch.pia[i], 1)
        assert ch.ctx[0].sort_last_path == 1  # (the receiver gets 8192 arrivals in ONE cell: more than the extras ranking takes, general path)
        assert int(ch.pia[0].n_total[0]) == n - leavers and int(ch.pia[1].n_total[0]) == leavers
        ch.check_sorted()
        got = ch.rows()[1]
        np.testing.assert_array_equal(np.sort(got[:, 0]), 1.0 + np.arange(leavers))
        # stable: the arrivals keep the order they had in the sender's layout (ascending original position inside the cell)
        sent = rows[:leavers]
        order = np.argsort(np.floor(sent[:, 4] * G.inv_dx), kind="stable")  # all from one cell: the sender's sort kept the upload order
        np.testing.assert_array_equal(got[:, 0], sent[order][:, 0])
    finally:
        ch.close()


def test_chunk_exchange_many_leavers_per_face(mb, oracle):
    """The reference's own way of scaling (same L, finer cells, BENCHMARKS.md:93-99) makes a particle cross hundreds of cells per step and
    > 1e5 particles cross every slab face: full exchange, and the arrivals (far too many and too spread for the band) go through
    the sort's general path.  Two steps against the chunk emulation (order included) and the single-domain oracle, bit for bit."""
    nx, n = 4000, 1_200_000
    L = 5e-4
    G = mb.Grid1DUniform(L, nx)  # dx = 1.25e-7; sigma_v dt = 250 * 2.59e-9 * 200 = 1.3e-4 = 1000 cells
    rows = _unique_rows(np.random.default_rng(3), n, L, sigma=250.0)
    ch = Chunks(mb, 2, G, rows, 3 * n, band=0, mode=1)
    try:
        walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)
        dt = 2.59e-9 * 200
        crossed = 0
        expect = ch.rows()
        for t in (1, 2):
            expect = _emulate_chunk_step(oracle, G, ch.slab, expect, t, dt)
            for i in range(2):
                mb.convect_particles(mb.PhiloxRng(t), ch.slab[i], walls, ch.pv[i], ch.pia[i], 1, AR, dt)
            before = [int(p.n_total[0]) for p in ch.pia]
            ch.exchange()
            crossed = max(crossed, int(ch.pia[0].n_total[0]) - before[0], int(ch.pia[1].n_total[0]) - before[1])
            for i in range(2):
                mb.sort_particles(None, ch.slab[i], ch.pv[i], ch.pia[i], 1)
        assert crossed >= 100_000, crossed
        ch.check_sorted()
        for got, want in zip(ch.rows(), expect):
            np.testing.assert_array_equal(got, want)
        opv, opia = _oracle_single_domain(oracle, rows, L, nx, 2, dt)
        a, b = np.concatenate(ch.rows()), opv.logical(1, n)
        np.testing.assert_array_equal(a[np.argsort(a[:, 0])], b[np.argsort(b[:, 0])])
    finally:
        ch.close()


def test_chunk_exchange_arrivals_as_band_extras(mb, oracle):
    """Thousands of arrivals per step spread over many cells next to the face are merged by the band path as extras (hybrid sort), in the
    stable order of the single-domain run."""
    nx, n = 600, 300_000
    L = nx * 1e-5
    G = mb.Grid1DUniform(L, nx)
    rows = _unique_rows(np.random.default_rng(8), n, L, sigma=300.0)
    ch = Chunks(mb, 3, G, rows, 2 * n, band=15, mode=1)
    try:
        walls = mb.MaxwellWalls1D(300.0, 300.0, 0.0, 0.0, 0.0, 0.0)
        dt = 4e-5 / 300.0  # 4 cells per step at sigma_v: band w = 15 holds (3.75 sigma), a few outliers per step
        extras = 0
        expect = ch.rows()
        for t in range(1, 6):
            expect = _emulate_chunk_step(oracle, G, ch.slab, expect, t, dt)
            for i in range(3):
                mb.convect_particles(mb.PhiloxRng(t), ch.slab[i], walls, ch.pv[i], ch.pia[i], 1, AR, dt)
            ch.exchange()
            for i in range(3):
                mb.sort_particles(None, ch.slab[i], ch.pv[i], ch.pia[i], 1)
                assert ch.ctx[i].sort_last_path == 1
                extras += ch.ctx[i].sort_last_extras
        assert extras > 5000, extras
        ch.check_sorted()
        for got, want in zip(ch.rows(), expect):
            np.testing.assert_array_equal(got, want)
        opv, opia = _oracle_single_domain(oracle, rows, L, nx, 5, dt)
        a, b = np.concatenate(ch.rows()), opv.logical(1, n)
        np.testing.assert_array_equal(a[np.argsort(a[:, 0])], b[np.argsort(b[:, 0])])
    finally:
        ch.close()

"""Full-size runs of BASELINE.json's single-GPU shapes, checked through size-independent properties (the oracle would take hours
at these sizes): sortedness and idempotence of the sort, conservation of particle count / mass / momentum / energy, per-cell
invariants of fp_linear!, N:2 merge conservation.  Each test generates its population on the host (numpy) and runs every
timestep on the device through the C ABI."""
import math

import numpy as np
import pytest

from parity_util import AR

pytestmark = pytest.mark.gpu
K_B = 1.380649e-23
DX, NDENS, DT = 1e-5, 5e22, 2.59e-9


@pytest.fixture(scope="module")
def ctx(mb):
    c = mb.Context(0, 777)
    yield c
    c.close()


def _population(n_cells, ppc, seed, vw=False, aniso=1.0):
    """`ppc` particles per cell, Maxwellian 300 K, x uniform inside the cell; SoA host arrays (w, vx, vy, vz, x, y, z)."""
    rng = np.random.default_rng(seed)
    n = n_cells * ppc
    sig = math.sqrt(K_B * 300.0 / AR)
    Fnum = DX * NDENS / ppc
    a = [np.empty(n) for _ in range(7)]
    a[0][:] = Fnum
    if vw:
        a[0] *= rng.uniform(0.5, 1.5, n)
    for f in (1, 2, 3):
        a[f][:] = rng.standard_normal(n) * sig
    a[1] *= aniso
    a[4][:] = (np.repeat(np.arange(n_cells, dtype=np.float64), ppc) + rng.uniform(0.001, 0.999, n)) * DX
    a[5][:] = 0.5
    a[6][:] = 0.5
    ix = np.zeros((1, n_cells, 7), dtype=np.int64)
    c = np.arange(n_cells, dtype=np.int64)
    ix[0, :, 0] = ppc
    ix[0, :, 1] = c * ppc + 1
    ix[0, :, 2] = (c + 1) * ppc
    ix[0, :, 3] = ppc
    ix[0, :, 5] = -1
    return a, ix, n, Fnum


def _totals(a):
    w = a[0]
    return np.array([w.sum(), (w * a[1]).sum(), (w * a[2]).sum(), (w * a[3]).sum(), (w * (a[1] ** 2 + a[2] ** 2 + a[3] ** 2)).sum()])


def test_couette_step_full_size(mb, ctx):
    """C3 per-GPU shape: 1.0e8 particles, 1e5 cells x 1000 ppc, the full step (ntc_equal_weight -> convect -> sort -> props) x 4.
    Properties: pia tiles 1..n exactly (device validator), every particle sits in the cell the pia says, the particle count and
    the total mass are conserved exactly, the props' per-cell counts add up to n, a second sort changes nothing (idempotence),
    and the energy changes only through the (few) wall hits."""
    n_cells, ppc = 100_000, 1000
    a, ix, n, Fnum = _population(n_cells, ppc, 11)
    L = n_cells * DX
    pv, pia = mb.ParticleVector(int(1.02 * n), ctx), mb.ParticleIndexerArray(n_cells, 1, ctx)
    pv.upload_soa(1, n, a)
    pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
    t0 = _totals(a)
    del a
    grid = mb.Grid1DUniform(L, n_cells)
    walls = mb.MaxwellWalls1D(300.0, 300.0, -500.0, 500.0, 1.0, 1.0)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    cf = mb.CollisionFactors(n_cells, mb.estimate_sigma_g_w_max(it, AR, AR, 300.0, 300.0, Fnum), ctx)
    pp = mb.PhysProps(n_cells, 1, ctx=ctx)
    for t in range(1, 5):
        r = mb.PhiloxRng(t)
        mb.ntc_equal_weight(r, cf, None, it, pv, pia, (1, n_cells), 1, DT, DX)
        mb.convect_particles(r, grid, walls, pv, pia, 1, AR, DT)
        mb.sort_particles(None, grid, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], pp)
    assert ctx.sort_last_path == 1
    ok, where = pia.check(1)
    assert ok, where
    ixd, nt, ct = pia.download()
    assert nt[0] == n and ct[0] == 1
    d = pp.download()
    assert d["np"].sum() == n
    np.testing.assert_array_equal(d["np"][0], ixd[0, :, 0])
    b = [np.empty(n) for _ in range(7)]
    pv.download_soa(1, n, b)
    cell = np.floor(b[4] * grid.inv_dx).astype(np.int64)
    np.testing.assert_array_equal(cell, np.repeat(np.arange(n_cells), ixd[0, :, 0]))  # sorted, and where the pia says
    t1 = _totals(b)
    assert t1[0] == t0[0]                                   # equal weights: mass bit-exact
    assert abs(t1[4] - t0[4]) / t0[4] < 1e-4                # elastic collisions; only ~50 wall hits per step change the energy
    assert abs(d["T"].mean() - 300.0) < 1.0
    cs = d["n"][0].sum()
    assert abs(cs - t0[0]) / t0[0] < 1e-12                  # checksum of the per-cell densities
    mb.sort_particles(None, grid, pv, pia, 1)               # idempotence
    c2 = [np.empty(n) for _ in range(7)]
    pv.download_soa(1, n, c2)
    for f in range(7):
        np.testing.assert_array_equal(b[f], c2[f])
    np.testing.assert_array_equal(pia.download()[0], ixd)


def test_fp_linear_full_size(mb, ctx):
    """C5: 1e8 particles in 1e6 independent cells of 100 (test_collision_fp.jl / test_1D_couette_fp.jl shape).  Two fp_linear! steps:
    every cell's momentum and energy are conserved (checked on all 1e6 cells, 1e-11 relative), positions and weights untouched,
    the anisotropy T_x / T_y decreases."""
    n_cells, ppc = 1_000_000, 100
    a, ix, n, Fnum = _population(n_cells, ppc, 12, aniso=1.8)
    pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(n_cells, 1, ctx)
    pv.upload_soa(1, n, a)
    pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)

    def per_cell(arr):
        v = [arr[f].reshape(n_cells, ppc) for f in (1, 2, 3)]
        return np.stack([x.sum(1) for x in v], 1), sum((x ** 2).sum(1) for x in v), v[0].var(1).mean() / v[1].var(1).mean()

    p0, e0, r0 = per_cell(a)
    for t in (1, 2):
        mb.fp_linear(mb.PhiloxRng(t), None, it, AR, pv, pia, (1, n_cells), 1, DT * 200, DX)
    b = [np.empty(n) for _ in range(7)]
    pv.download_soa(1, n, b)
    p1, e1, r1 = per_cell(b)
    sig = math.sqrt(K_B * 300.0 / AR)
    assert np.max(np.abs(p1 - p0)) < 1e-9 * sig * ppc
    assert np.max(np.abs(e1 - e0) / e0) < 1e-11
    for f in (0, 4, 5, 6):
        np.testing.assert_array_equal(a[f], b[f])
    assert r1 < r0 and r0 > 3.0


def test_varweight_merge_full_size(mb, ctx):
    """C2 / C4 shape: 6e7 variable-weight particles in 4e5 cells of 150; merge_octree_N2_based! (threshold 130, target 100, as
    couette_multithreaded_varweight_octree.jl:205-206) on all cells in one launch, then squash_pia!.  Every cell ends with
    <= 100 particles, and every cell's mass, momentum and energy are conserved to 1e-12 relative (merging_octree_N2.jl:736-933)."""
    n_cells, ppc, target = 400_000, 150, 100
    a, ix, n, Fnum = _population(n_cells, ppc, 13, vw=True)
    L = n_cells * DX
    pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(n_cells, 1, ctx)
    pv.upload_soa(1, n, a)
    pia.upload(ix, np.array([n]), np.array([1], dtype=np.uint8))
    grid = mb.Grid1DUniform(L, n_cells)

    def per_cell(arr, counts):
        off = np.concatenate(([0], np.cumsum(counts)))[:-1]
        w = arr[0]
        out = [np.add.reduceat(w, off)]
        for f in (1, 2, 3):
            out.append(np.add.reduceat(w * arr[f], off))
        out.append(np.add.reduceat(w * (arr[1] ** 2 + arr[2] ** 2 + arr[3] ** 2), off))
        return np.stack(out, 1)

    m0 = per_cell(a, np.full(n_cells, ppc))
    oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
    mb.merge_octree_N2_based(mb.PhiloxRng(1), oc, pv, pia, (1, n_cells), 1, target, grid, threshold=130)
    mb.squash_pia(pv, pia, 1)  # closes the holes; the cells keep their (merged) particles, in cell order
    ixd, nt, ct = pia.download()
    counts = ixd[0, :, 0]
    assert counts.max() <= target and counts.min() > 0 and nt[0] == counts.sum() and ct[0] == 1
    np.testing.assert_array_equal(ixd[0, :, 1], np.concatenate(([0], np.cumsum(counts)))[:-1] + 1)
    n1 = int(nt[0])
    b = [np.empty(n1) for _ in range(7)]
    pv.download_soa(1, n1, b)
    assert b[4].min() >= grid.min_x and b[4].max() <= grid.max_x  # the 1-D variant clamps x1 to the domain (:878-897)
    m1 = per_cell(b, counts)
    np.testing.assert_allclose(m1[:, 0], m0[:, 0], rtol=1e-12)
    sig = math.sqrt(K_B * 300.0 / AR)
    assert np.max(np.abs(m1[:, 1:4] - m0[:, 1:4]) / (m0[:, :1] * sig)) < 1e-11
    np.testing.assert_allclose(m1[:, 4], m0[:, 4], rtol=1e-12)
    # the merged particles sit at mean +- sigma_x and may cross into a neighbour cell: the sort re-bins them, nothing is lost
    mb.sort_particles(None, grid, pv, pia, 1)
    ixs, nts, _ = pia.download()
    assert nts[0] == n1
    pv.download_soa(1, n1, b)
    np.testing.assert_array_equal(np.floor(b[4] * grid.inv_dx).astype(np.int64), np.repeat(np.arange(n_cells), ixs[0, :, 0]))
    np.testing.assert_allclose(_totals(b), m0.sum(0), rtol=1e-11, atol=1e-9 * m0[:, 0].sum() * sig)

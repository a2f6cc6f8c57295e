"""CPU suite: the oracle against runs OF THE REFERENCE ITSELF -- the golden netCDF files its regression tests compare to at 1e-13
(test/data/*.nc; extracted by tests/golden/make_golden.py into tests/golden/reference_histories.json).  tests/test_oracle_reference_bitlevel.py
reproduces them to round-off on the reference's own random stream; this file is the generator-independent complement.  The comparison
is at distribution level: deterministic quantities (grid-sampled
initial moments, counts, densities) must match to round-off; stochastic histories must be one plausible draw of the oracle's own
ensemble (z-scores against the ensemble mean / spread at every recorded step); Couette cell profiles must agree within the
per-cell sampling noise (chi-square over the 50 cells)."""
import json
import math
import os

import numpy as np
import pytest

from test_oracle_stat import _bkw_setup

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(GOLDEN, "reference_histories.json")))


def _z(ref_hist, ens):
    """z-score of the reference's single run against an ensemble of K oracle runs (axis 0), per record."""
    K = ens.shape[0]
    mean, sd = ens.mean(0), ens.std(0, ddof=1)
    return (ref_hist - mean) / (sd * math.sqrt(1.0 + 1.0 / K) + 1e-300)


# ---------------------------------------------------------------------------------------------------------------------------------
# 0-D, Ar + He relaxation
# ---------------------------------------------------------------------------------------------------------------------------------
def _two_species_run(oracle, seed, nA, nH, FA, FH, merge, every=25, n_t=800):
    mA, mH = oracle.MASS["Ar"], oracle.MASS["He"]
    TA, TH, dt, V = 3000.0, 360.0, 2.5e-3, 1.0
    pvA, pvH, pia = oracle.OPV(3 * nA), oracle.OPV(3 * nH), oracle.OPIA(1, 2)
    srng = oracle.Rng.seq(1000 + seed)
    oracle.sample_equal_weight_cell(srng, pvA, pia, 1, 1, nA, mA, TA, FA)
    oracle.sample_equal_weight_cell(srng, pvH, pia, 1, 2, nH, mH, TH, FH)
    itAA, itHH = oracle.interaction("Ar", "Ar"), oracle.interaction("He", "He")
    d, o, Tr = oracle.VHS[("Ar", "He")]
    itHA = oracle.make_interaction(mH, mA, d, o, Tr)
    cfAA = oracle.CF(1, oracle.estimate_sigma_g_w_max(itAA, mA, mA, TA, TA, FA))
    cfHH = oracle.CF(1, oracle.estimate_sigma_g_w_max(itHH, mH, mH, TH, TH, FH))
    cfHA = oracle.CF(1, oracle.estimate_sigma_g_w_max(itHA, mH, mA, TH, TA, max(FA, FH)))
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    rng = oracle.Rng.seq(seed)
    T = [oracle.compute_props([pvA, pvH], pia, [mA, mH]).T[:, 0].copy()]
    for ts in range(1, n_t + 1):
        oracle.ntc(rng, cfAA, itAA, pvA, pia, 1, 1, 1, dt, V)
        oracle.ntc2(rng, cfHA, itHA, pvH, pvA, pia, 1, 1, 2, 1, dt, V)
        oracle.ntc(rng, cfHH, itHH, pvH, pia, 1, 1, 2, dt, V)
        if merge:
            for s, (pv, n0) in enumerate(((pvA, nA), (pvH, nH))):
                if pia.indexer[s, 0, 0] > round(1.2 * n0):
                    oracle.merge_octree_N2(rng, oc, pv, pia, 1, 1, s + 1, n0)
                    oracle.squash_pia(pv, pia, s + 1)
        if ts % every == 0:
            T.append(oracle.compute_props([pvA, pvH], pia, [mA, mH]).T[:, 0].copy())
    return np.array(T)


@pytest.mark.parametrize("key,nA,nH,FA,FH,merge", [("two_species", 400, 4000, 5e12, 5e12, False),
                                                   ("two_species_varweight_octree", 4000, 4000, 5e11, 5e12, True)])
def test_two_species_history_is_a_draw_of_the_oracle_ensemble(oracle, ref, key, nA, nH, FA, FH, merge):
    """test/test_2species.jl / test_2species_varweight_octree.jl: T_Ar(t), T_He(t) of the reference's golden run, every 25 steps,
    against 8 oracle runs (sampling and collision seeds varied): |z| < 4.5 at every record, rms z < 2 per species."""
    r = ref[key]
    T_ref = np.array(r["T"])
    assert T_ref.shape == (33, 2) and r["timestep"][1] == 25.0
    nd = np.array(r["ndens"])
    assert np.all(np.abs(nd[:, 0] / 2e15 - 1) < 2e-15) and np.all(np.abs(nd[:, 1] / 2e16 - 1) < 6e-15)  # both runs conserve n
    ens = np.array([_two_species_run(oracle, s, nA, nH, FA, FH, merge) for s in range(8)])
    z = _z(T_ref, ens)
    assert np.max(np.abs(z)) < 4.5, (np.max(np.abs(z)), np.unravel_index(np.argmax(np.abs(z)), z.shape))
    assert np.all(np.sqrt((z ** 2).mean(0)) < 2.0), np.sqrt((z ** 2).mean(0))
    # and the relaxation itself: the ensemble-mean Ar excess temperature follows the reference's to a few per cent of its start value
    Tm = (2e15 * T_ref[:, 0] + 2e16 * T_ref[:, 1]) / 2.2e16
    Em = (2e15 * ens[:, :, 0] + 2e16 * ens[:, :, 1]).mean(0) / 2.2e16
    ex_ref, ex_or = (T_ref[:, 0] - Tm) / (T_ref[0, 0] - Tm[0]), (ens[:, :, 0].mean(0) - Em) / (ens[:, 0, 0].mean() - Em[0])
    assert np.max(np.abs(ex_ref - ex_or)) < (0.05 if nA == 400 else 0.02), np.max(np.abs(ex_ref - ex_or))


# ---------------------------------------------------------------------------------------------------------------------------------
# 0-D BKW relaxation: equal weights, variable weights with octree / velocity-grid merging, SWPM
# ---------------------------------------------------------------------------------------------------------------------------------
def test_bkw_grid_sampled_initial_state_matches_the_reference_run(oracle, ref):
    """No random numbers involved: sample_on_grid! of the BKW distribution on the 40^3 velocity grid (the initial state of
    test_bkw_varweight_octree.jl / _grid.jl / _octree_swpm.jl) gives the reference's particle count exactly and its T, n and total
    moments M4..M10 (compute_props_with_total_moments!) to round-off."""
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)
    pv, pia = oracle.OPV(40 ** 3), oracle.OPIA(1, 1)
    n_s = int(oracle.sample_on_grid(oracle.Rng.seq(1), "bkw", pv, 40, m, T0, n_dens))
    pia.set_single_cell(1, 1, n_s)
    p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
    for key in ("bkw_vw_octree", "bkw_vw_grid", "bkw_vw_octree_swpm"):
        r = ref[key]
        assert r["moment_powers"] == [4, 6, 8, 10]
        assert n_s == int(r["np"][0]) == 30976
        assert abs(p.T[0, 0] - r["T"][0]) < 1e-12
        assert abs(p.n[0, 0] / r["ndens"][0] - 1.0) < 1e-15
        np.testing.assert_allclose(p.moments[0, 0], r["moments"][0], rtol=1e-15)  # with the fused `5 xk - 3` of the reference's bkw()


def _bkw_octree_history(oracle, seed, n_t=500):
    """the loop of test/test_bkw_varweight_octree.jl:84-96"""
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)
    pv, pia = oracle.OPV(40 ** 3), oracle.OPIA(1, 1)
    n_s = int(oracle.sample_on_grid(oracle.Rng.seq(1), "bkw", pv, 40, m, T0, n_dens))
    pia.set_single_cell(1, 1, n_s)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    p = oracle.compute_props([pv], pia, [m], [4, 6, 8], Tref=T0, with_moments=True)
    cf = oracle.CF(1, oracle.estimate_sigma_g_w_max(it, m, m, T0, T0, n_dens / n_s))
    rng = oracle.Rng.seq(seed)
    hist = np.zeros((n_t + 1, 3))
    hist[0] = p.moments[0, 0]
    for ts in range(1, n_t + 1):
        oracle.ntc(rng, cf, it, pv, pia, 1, 1, 1, 0.025 * tref, 1.0)
        if p.np[0, 0] > 10000:
            oracle.merge_octree_N2(rng, oc, pv, pia, 1, 1, 1, 8000)
            oracle.squash_pia(pv, pia, 1)
        p = oracle.compute_props([pv], pia, [m], [4, 6, 8], Tref=T0, with_moments=True)
        hist[ts] = p.moments[0, 0]
    assert abs(p.T[0, 0] - 272.99978354) < 5e-4 and abs(p.n[0, 0] / n_dens - 1.0) < 1e-11
    return hist


@pytest.mark.parametrize("key,tol", [("bkw_20k", (0.05, 0.055, 0.15)), ("bkw_vw_octree", (0.025, 0.078, 0.21)),
                                     ("bkw_vw_grid", (0.025, 0.06, 0.13)), ("bkw_vw_octree_swpm", (0.02, 0.07, 0.15))])
def test_bkw_moment_history_is_a_draw_of_the_oracle_ensemble(oracle, ref, key, tol):
    """M4, M6, M8 of the reference's golden BKW runs (every 10th of 500 steps) against 8 oracle runs of the same loop.  The moment
    histories are strongly correlated in time, so the run is compared as a whole: (1) it stays within 1.25x the reference test's own
    tolerance (the one it uses against the analytic solution; the margin covers the merging bias of the ensemble mean, ~2 % in M8) of
    the oracle's ensemble mean; (2) its fluctuation level (rms relative
    deviation from the ensemble mean) and (3) its time-averaged offset lie in the range the oracle's own runs span (with a margin of
    that range's width: the golden octree run is a noisy one, at the edge of 24 oracle runs)."""
    import test_oracle_stat as st

    run = {"bkw_20k": lambda s: st._bkw_history(oracle, 500 + s, 40 + s)[:, :3],
           "bkw_vw_octree": lambda s: _bkw_octree_history(oracle, 60 + s),
           "bkw_vw_grid": lambda s: st._bkw_grid_merging_history(oracle, 80 + s),
           "bkw_vw_octree_swpm": lambda s: st._bkw_swpm_history(oracle, 90 + s)}[key]
    r = ref[key]
    assert r["moment_powers"][:3] == [4, 6, 8] and len(r["timestep"]) == 51
    M_ref = np.array(r["moments"])[:, :3]
    ens = np.array([run(s)[::10] for s in range(8)])
    mean = ens.mean(0)
    dev = np.max(np.abs(M_ref - mean) / mean, axis=0)
    assert np.all(dev < 1.25 * np.array(tol)), dev
    rel = lambda h: (h[10:] - mean[10:]) / mean[10:]  # records past the initial transient (t > 2.5 t_ref)
    rms = lambda h: np.sqrt((rel(h) ** 2).mean(0))
    rms_k, off_k = np.array([rms(h) for h in ens]), np.array([rel(h).mean(0) for h in ens])
    assert np.all(rms(M_ref) < 2.0 * rms_k.max(0)) and np.all(rms(M_ref) > 0.5 * rms_k.min(0)), (rms(M_ref), rms_k.min(0), rms_k.max(0))
    span = off_k.max(0) - off_k.min(0)
    off = rel(M_ref).mean(0)
    assert np.all(off < off_k.max(0) + span) and np.all(off > off_k.min(0) - span), (off, off_k.min(0), off_k.max(0))


# ---------------------------------------------------------------------------------------------------------------------------------
# 1-D Couette flow, 50 cells: snapshots every 1000 steps
# ---------------------------------------------------------------------------------------------------------------------------------
def _couette_run(oracle, seed, variant, n_steps, ppc, every=1000):
    """The time loops of test/test_1D_couette.jl:64-95 ("ntc"), test_1D_couette_varweight.jl:62-105 ("vw": octree merge 200 -> 150 per
    cell), test_1D_couette_varweight_swpm.jl:66-100 ("swpm": G = 1.5, same merge), test_1D_couette_varweight_index_resort.jl:62-114
    ("resort": merge 150 -> 100, sigma_g_w_max estimated at the merged weight, restore_particle_ordering! every 500 steps) and
    test_1D_couette_fp.jl:56-76 ("fp").  Returns snapshots [record][quantity][cell] of (np, n, T, vy)."""
    m = oracle.MASS["Ar"]
    it = oracle.interaction("Ar", "Ar")
    T_wall, v_wall, L, ndens, nx, dt = 300.0, 500.0, 5e-4, 5e22, 50, 2.59e-9
    V = L / nx
    Fnum = V * ndens / ppc
    thr, tgt = {"vw": (200, 150), "swpm": (200, 150), "resort": (150, 100)}.get(variant, (0, 0))
    grid, walls = (L, nx), (T_wall, T_wall, -v_wall, v_wall, 1.0, 1.0)
    pv, pia = oracle.OPV(ppc * nx), oracle.OPIA(nx, 1)
    oracle.sample_equal_weight_grid(oracle.Rng.seq(3000 + seed), grid, pv, pia, 1, m, ndens, T_wall, Fnum)
    F_cf = {"swpm": 1.0, "resort": Fnum * ppc / 100}.get(variant, Fnum)
    cf = oracle.CF(nx, oracle.estimate_sigma_g_w_max(it, m, m, T_wall, T_wall, F_cf))
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    rng = oracle.Rng.seq(seed)
    if thr:
        oracle.merge_octree_N2(rng, oc, pv, pia, 1, nx, 1, tgt, threshold=thr, grid=grid)
        oracle.squash_pia(pv, pia, 1)

    def snap():
        p = oracle.compute_props([pv], pia, [m])
        return np.stack([p.np[0], p.n[0], p.T[0], p.v[0, :, 1]])

    out = [snap()]
    for t in range(1, n_steps + 1):
        if variant == "fp":
            oracle.fp_linear(rng, it, m, pv, pia, 1, nx, 1, dt, V)
        elif variant == "swpm":
            oracle.swpm(rng, cf, it, pv, pia, 1, nx, 1, 1.5, dt, V)
        else:
            oracle.ntc(rng, cf, it, pv, pia, 1, nx, 1, dt, V)
        if thr:
            oracle.merge_octree_N2(rng, oc, pv, pia, 1, nx, 1, tgt, threshold=thr, grid=grid, squash_after_each=True)
        oracle.convect_particles(rng, grid, walls, pv, pia, 1, [m], dt)
        oracle.sort_particles(pv, pia, 1, grid=grid)
        if variant == "resort" and t % 500 == 0:
            oracle.restore_particle_ordering(pv)
            assert oracle.check_unique_index(pv, pia, 1) == (True, 0)
        if t % every == 0:
            out.append(snap())
    if thr:
        assert pia.check(1) == (True, 0) and oracle.check_unique_index(pv, pia, 1) == (True, 0)
    return np.array(out)


@pytest.mark.parametrize("key,variant,ppc,n_steps,K", [("couette", "ntc", 1000, 3000, 4), ("couette_vw200to150", "vw", 1000, 6000, 6),
                                                       ("couette_vw200to150_swpm", "swpm", 1000, 4000, 6),
                                                       ("couette_vw150to100_resort", "resort", 500, 3000, 6),
                                                       ("couette_fp_linear", "fp", 200, 3000, 4)])
def test_couette_snapshots_agree_within_the_cell_noise(oracle, ref, key, variant, ppc, n_steps, K):
    """The reference's golden Couette runs (NTC; NTC + octree merging; SWPM + merging; merging + index re-sorting; Fokker-Planck)
    record np, n, T, v per cell every 1000 steps.  K oracle runs of the same loop give the mean profile and the per-cell spread of a
    single run (pooled over cells: relative for n and T, absolute for v_y); the reference snapshot must differ from the mean profile
    like one more run does: chi^2 / 50 cells < 2.2 per record and quantity (expected 1 +- 0.2), < 1.4 averaged over the records."""
    r = ref[key]
    n_rec = n_steps // 1000 + 1
    assert r["timestep"][:n_rec] == [1000.0 * i for i in range(n_rec)]
    R = np.stack([np.array(r["np"])[:n_rec], np.array(r["ndens"])[:n_rec], np.array(r["T"])[:n_rec], np.array(r["v"])[:n_rec, :, 1]], 1)
    ens = np.array([_couette_run(oracle, 10 + s, variant, n_steps, ppc) for s in range(K)])  # [run][record][quantity][cell]
    merging = variant in ("vw", "swpm", "resort")
    # exact facts first: particle count at t = 0, total number density at every record
    if not merging:
        assert np.all(R[:, 0].sum(1) == 50 * ppc) and np.all(ens[:, :, 0].sum(2) == 50 * ppc)
        assert np.all(R[0, 0] == ppc) and np.all(ens[:, 0, 0] == ppc)
    n_tot = 50 * 5e22 * 1e-5
    assert np.all(np.abs(R[:, 1].sum(1) / n_tot - 1) < 1e-12) and np.all(np.abs(ens[:, :, 1].sum(2) / n_tot - 1) < 1e-12)
    mean = ens.mean(0)
    chi = np.zeros((n_rec - 1, 3))
    for rec in range(1, n_rec):
        for q, relative in ((1, True), (2, True), (3, False)):
            scale = mean[rec, q] if relative else 1.0
            resid = (ens[:, rec, q] - mean[rec, q]) / scale
            var = (resid ** 2).sum() / ((K - 1) * 50)  # pooled single-run variance
            d = (R[rec, q] - mean[rec, q]) / scale
            chi[rec - 1, q - 1] = (d ** 2).mean() / (var * (1.0 + 1.0 / K))
    assert np.all(chi < 2.2), chi
    assert np.all(chi.mean(0) < 1.4), chi.mean(0)
    if merging:  # the per-cell counts stay between the target and the threshold + one step of splits, with the reference's mean
        assert np.all(R[1:, 0] <= 1.3 * mean[1:, 0].max()) and abs(R[1:, 0].mean() / ens[:, 1:, 0].mean() - 1) < 0.03, (R[1:, 0].mean(), ens[:, 1:, 0].mean())
    if merging:  # the merged initial state: counts per cell are deterministic up to the sampled velocities -- compare their means
        assert abs(R[0, 0].mean() / ens[:, 0, 0].mean() - 1) < 0.02, (R[0, 0].mean(), ens[:, 0, 0].mean())

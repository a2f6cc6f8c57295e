"""CPU suite: the oracle REPRODUCES THE REFERENCE'S OWN RUNS.  The reference's regression tests run seeded simulations with
`StableRNG(1234)` and compare against golden netCDF files (test/data/*.nc) at ~1e-13.  With the StableRNGs.jl generator restated in
oracle/philox.hpp (LehmerRNG: 128-bit multiplicative congruential state, high 64 bits per draw, Julia's 52-bit mantissa conversion
for rand(Float64), low bit for rand(rng, [-1.0, 1.0])) the oracle, driven through the same call sequence as the reference's test
scripts, lands on the golden files' values to round-off -- every recorded step, every cell:

  0-D  test_2species.jl, test_2species_equal_weight.jl   sampling + ntc! / ntc_equal_weight! (1 and 2 species), 800 steps
  0-D  test_2species_varweight_octree.jl          variable-weight ntc! (splits) + merge_octree_N2_based!, 800 steps, ~130 merges
  0-D  test_bkw.jl                                sample_bkw! (Chi(5) of Distributions.jl: Marsaglia-Tsang gamma sampler on Julia's
                                                  ziggurat randn) + ntc!, 20 000 particles, total moments M4..M10, 500 steps
  0-D  test_bkw_varweight_octree.jl, test_bkw_varweight_octree_swpm.jl   sample_on_grid!(bkw) + ntc! / swpm! + octree merge
                                                  10 000 -> 8 000 of a cell that starts on a mirror-symmetric velocity lattice, 500 steps
  0-D  test_bkw_varweight_grid.jl                 sample_on_grid!(bkw) + ntc! + merge_grid_based!, total moments M4..M10, 500 steps
  1-D  test_1D_couette.jl                         sample on grid + ntc! + convect_particles! (diffuse walls) + sort_particles!
  1-D  test_1D_couette_varweight.jl               + per-cell octree merging with position clamping + squash_pia! + SurfProps
  1-D  test_1D_couette_varweight_swpm.jl          swpm! instead of ntc!
  1-D  test_1D_couette_varweight_index_resort.jl  + restore_particle_ordering! every 500 steps
  1-D  test_1D_couette_fp.jl                      fp_linear! with Julia's randn (256-strip ziggurat, tables regenerated in philox.hpp)

(golden values: tests/golden/reference_histories.json, extracted from the .nc files by tests/golden/make_golden.py).  Tolerances are
those of the reference's own comparisons, or a few ulp of the quantity: the Julia build fuses multiply-adds (@muladd) and uses its own
libm, the oracle is compiled with -ffp-contract=off against glibc, so the last bits differ while every random decision is the same.

The octree BKW runs are the delicate ones: the first merge acts on the mirror-symmetric lattice of sample_on_grid!, mirror-image
octree bins carry weights that are equal up to the last bits, and the strict `w > max_w` choice of the bin to refine is decided by
those bits.  They only come out right because the oracle evaluates the distribution exactly like the reference does: the `5 xk - 3`
of bkw() as a fused multiply-add (@muladd) and exp() with Julia's own algorithm (merzbild.jl_b200/csrc/mb_jlexp.h; glibc's exp differs
from it in the last bit for ~1 % of the arguments) -- see the two tests at the end of this file.  tests/test_oracle_reference_runs.py
adds the generator-independent view (golden histories as draws of the oracle ensemble)."""
import json
import os

import numpy as np
import pytest

from test_oracle_stat import _bkw_setup

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(GOLDEN, "reference_histories.json")))


def test_stable_rng_stream(oracle):
    """StableRNGs.jl LehmerRNG: state = (seed << 1) | 1, state *= 0x45a3...8add mod 2^128, output = high 64 bits;
    rand(Float64) = bitcast(0x3ff0000000000000 | (u & (2^52 - 1))) - 1 -- against a big-integer restatement in Python."""
    mult, mask = 0x45A31EFC5A35D971261FD0407A968ADD, (1 << 128) - 1
    state = (1234 << 1) | 1
    rng = oracle.Rng.stable(1234)
    for _ in range(1000):
        state = (state * mult) & mask
        u = state >> 64
        expect = np.array([0x3FF0000000000000 | (u & ((1 << 52) - 1))], dtype=np.uint64).view(np.float64)[0] - 1.0
        assert rng.rand() == expect


def _two_species(oracle, variable_weight, equal_weight_api=False):
    """test/test_2species.jl:27-71, test/test_2species_varweight_octree.jl:14-83; equal_weight_api: ntc_equal_weight! instead of ntc!
    (test/test_2species_equal_weight.jl:52-66, same golden file)"""
    mA, mH = oracle.MASS["Ar"], oracle.MASS["He"]
    TA, TH, dt, V = 3000.0, 360.0, 2.5e-3, 1.0
    nA, nH, FA, FH = (4000, 4000, 5e11, 5e12) if variable_weight else (400, 4000, 5e12, 5e12)
    pvA, pvH, pia = oracle.OPV(nA), oracle.OPV(nH), oracle.OPIA(1, 2)
    rng = oracle.Rng.stable(1234)
    oracle.sample_equal_weight_cell(rng, pvA, pia, 1, 1, nA, mA, TA, FA)
    oracle.sample_equal_weight_cell(rng, pvH, pia, 1, 2, nH, mH, TH, FH)
    itAA, itHH = oracle.interaction("Ar", "Ar"), oracle.interaction("He", "He")
    d, o, Tr = oracle.VHS[("Ar", "He")]
    itHA = oracle.make_interaction(mH, mA, d, o, Tr)
    cfAA = oracle.CF(1, oracle.estimate_sigma_g_w_max(itAA, mA, mA, TA, TA, FA))
    cfHH = oracle.CF(1, oracle.estimate_sigma_g_w_max(itHH, mH, mH, TH, TH, FH))
    cfHA = oracle.CF(1, oracle.estimate_sigma_g_w_max(itHA, mH, mA, TH, TA, max(FA, FH)))
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    out = [oracle.compute_props([pvA, pvH], pia, [mA, mH])]
    n_merges = 0
    for ts in range(1, 801):
        oracle.ntc(rng, cfAA, itAA, pvA, pia, 1, 1, 1, dt, V, equal_weight=equal_weight_api)
        oracle.ntc2(rng, cfHA, itHA, pvH, pvA, pia, 1, 1, 2, 1, dt, V, equal_weight=equal_weight_api)
        oracle.ntc(rng, cfHH, itHH, pvH, pia, 1, 1, 2, dt, V, equal_weight=equal_weight_api)
        if variable_weight:
            for s, (pv, n0) in enumerate(((pvA, nA), (pvH, nH))):
                if pia.indexer[s, 0, 0] > round(1.2 * n0):
                    oracle.merge_octree_N2(rng, oc, pv, pia, 1, 1, s + 1, n0)
                    n_merges += 1
        if ts % 25 == 0:
            out.append(oracle.compute_props([pvA, pvH], pia, [mA, mH]))
    return out, n_merges


@pytest.mark.parametrize("key,variable_weight,eq_api", [("two_species", False, False), ("two_species", False, True),
                                                        ("two_species_varweight_octree", True, False)])
def test_two_species_runs_reproduce_the_golden_files(oracle, ref, key, variable_weight, eq_api):
    """T to 9.3e-13 K (the reference's own tolerance, test_2species_varweight_octree.jl:93-95), v to 1e-12 m/s, n to 6e-15 relative,
    particle counts exactly -- at every 25th of the 800 steps, through ~130 octree merges in the variable-weight run."""
    r = ref[key]
    out, n_merges = _two_species(oracle, variable_weight, eq_api)
    assert len(out) == 33 and (n_merges > 100) == variable_weight
    for rec, p in enumerate(out):
        assert np.array_equal(p.np[:, 0], r["np"][rec])
        assert np.max(np.abs(p.T[:, 0] - r["T"][rec])) < 9.3e-13
        assert np.max(np.abs(p.v[:, 0] - np.array(r["v"][rec]))) < 1e-12
        assert np.max(np.abs(p.n[:, 0] / np.array(r["ndens"][rec]) - 1.0)) < 6e-15


def test_bkw_equal_weight_run_reproduces_the_golden_file(oracle, ref):
    """test/test_bkw.jl:56-95: 20 000 equal-weight particles sampled from BKW(t = 0) -- |v| = sqrt(0.3) v_th chi_5 with chi_5 from
    Distributions.jl's Chi(5) (sqrt of a Marsaglia-Tsang Gamma(5/2, 2) variate: randn + rand per trial), angles from two uniform arrays --
    then 500 steps of ntc!: T, v and M4..M10 of the golden file at every 10th step to round-off."""
    r = ref["bkw_20k"]
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)
    n_p = 20000
    Fnum = n_dens / n_p
    pv, pia = oracle.OPV(n_p), oracle.OPIA(1, 1)
    rng = oracle.Rng.stable(1234)
    oracle.sample_equal_weight_cell(rng, pv, pia, 1, 1, n_p, m, T0, Fnum, distribution="BKW")
    cf = oracle.CF(1, oracle.estimate_sigma_g_w_max(it, m, m, T0, T0, Fnum))
    for ts in range(0, 501):
        if ts:
            oracle.ntc(rng, cf, it, pv, pia, 1, 1, 1, 0.025 * tref, 1.0)
        if ts % 10 == 0:
            rec = ts // 10
            p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
            assert p.np[0, 0] == n_p == r["np"][rec]
            np.testing.assert_allclose(p.moments[0, 0], r["moments"][rec], rtol=1e-13)
            assert abs(p.T[0, 0] - r["T"][rec]) < 1e-12 and np.max(np.abs(p.v[0, 0] - np.array(r["v"][rec]))) < 1e-12


def test_bkw_grid_merging_run_reproduces_the_golden_file(oracle, ref):
    """test/test_bkw_varweight_grid.jl:62-100: counts exactly (4929 after the first merge ...), M4..M10 to 1e-13 relative (the reference
    compares at 1e-15 absolute on its own platform, :115-118), T to 5e-12 K."""
    r = ref["bkw_vw_grid"]
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)
    pv, pia = oracle.OPV(40 ** 3), oracle.OPIA(1, 1)
    rng = oracle.Rng.stable(1234)
    n_s = int(oracle.sample_on_grid(rng, "bkw", pv, 40, m, T0, n_dens))
    pia.set_single_cell(1, 1, n_s)
    mg = oracle.GridMerge(16, 16, 16, 3.5)
    p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
    cf = oracle.CF(1, oracle.estimate_sigma_g_w_max(it, m, m, T0, T0, n_dens / n_s))
    n_merges = 0
    for ts in range(1, 501):
        oracle.ntc(rng, cf, it, pv, pia, 1, 1, 1, 0.025 * tref, 1.0)
        if p.np[0, 0] > 10000:
            oracle.merge_grid_based(rng, mg, pv, pia, 1, 1, 1, m, T_v=[[p.T[0, 0], *p.v[0, 0]]])
            n_merges += 1
        p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
        if ts % 10 == 0:
            rec = ts // 10
            assert int(p.np[0, 0]) == int(r["np"][rec]), (ts, p.np[0, 0], r["np"][rec])
            np.testing.assert_allclose(p.moments[0, 0], r["moments"][rec], rtol=1e-13)
            assert abs(p.T[0, 0] - r["T"][rec]) < 5e-12 and abs(p.n[0, 0] / r["ndens"][rec] - 1.0) < 1e-13
            assert np.max(np.abs(p.v[0, 0] - np.array(r["v"][rec]))) < 1e-12
    assert n_merges >= 8


@pytest.mark.parametrize("key,swpm", [("bkw_vw_octree", False), ("bkw_vw_octree_swpm", True)])
def test_bkw_octree_runs_reproduce_the_golden_files(oracle, ref, key, swpm):
    """test/test_bkw_varweight_octree.jl:62-100, test_bkw_varweight_octree_swpm.jl:66-100: counts exactly (7995 after the first merge,
    a merge every ~9 steps), M4..M10 to 1e-13 relative, T to 5e-12 K, at every 10th of the 500 steps."""
    r = ref[key]
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)
    pv, pia = oracle.OPV(40 ** 3), oracle.OPIA(1, 1)
    rng = oracle.Rng.stable(1234)
    n_s = int(oracle.sample_on_grid(rng, "bkw", pv, 40, m, T0, n_dens))
    pia.set_single_cell(1, 1, n_s)
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    cf = oracle.CF(1, oracle.estimate_sigma_g_w_max(it, m, m, T0, T0, 1.0 if swpm else n_dens / n_s))
    p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
    n_merges = 0
    for ts in range(1, 501):
        if swpm:
            oracle.swpm(rng, cf, it, pv, pia, 1, 1, 1, 1.0, 0.025 * tref, 1.0)
        else:
            oracle.ntc(rng, cf, it, pv, pia, 1, 1, 1, 0.025 * tref, 1.0)
        if p.np[0, 0] > 10000:
            oracle.merge_octree_N2(rng, oc, pv, pia, 1, 1, 1, 8000)
            n_merges += 1
        p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
        if ts == 1:
            f = r["first_step"]
            assert int(p.np[0, 0]) == f["np"]
            np.testing.assert_allclose(p.moments[0, 0], f["moments"], rtol=1e-13)
        if ts % 10 == 0:
            rec = ts // 10
            assert int(p.np[0, 0]) == int(r["np"][rec]), (ts, p.np[0, 0], r["np"][rec])
            np.testing.assert_allclose(p.moments[0, 0], r["moments"][rec], rtol=1e-13)
            assert abs(p.T[0, 0] - r["T"][rec]) < 5e-12 and abs(p.n[0, 0] / r["ndens"][rec] - 1.0) < 1e-13
    assert n_merges >= 20


# ---------------------------------------------------------------------------------------------------------------------------------
# 1-D Couette
# ---------------------------------------------------------------------------------------------------------------------------------
def _surf_columns(s, rec):
    c = lambda k: np.array(s[k][rec])
    return np.concatenate([c("np")[:, None], c("flux_incident")[:, None], c("flux_reflected")[:, None], c("force"), c("normal_pressure")[:, None],
                           c("shear_pressure"), c("kinetic_energy_flux")[:, None]], 1)


@pytest.mark.parametrize("key,variant,ppc,n_steps,thr,tgt", [("couette", "ntc", 1000, 6000, 0, 0),
                                                             ("couette_vw200to150", "vw", 1000, 6000, 200, 150),
                                                             ("couette_vw200to150_swpm", "swpm", 1000, 4000, 200, 150),
                                                             ("couette_vw150to100_resort", "resort", 500, 3000, 150, 100),
                                                             ("couette_fp_linear", "fp", 200, 3000, 0, 0)])
def test_couette_runs_reproduce_the_golden_files(oracle, ref, key, variant, ppc, n_steps, thr, tgt):
    """Cell profiles every 1000 steps: T to 2.4e-13 x 2 K (the reference compares its first five cells at 2.4e-13, test_1D_couette.jl:111;
    all 50 cells here), v to 5e-13 m/s, n exactly (sums of identical weights / merged halves), counts exactly; wall properties
    (hits, fluxes, force, pressures, kinetic-energy flux of the step that was recorded) to 1e-13 relative.  Fokker-Planck run: same cells
    exactly (every decision of the ziggurat agrees), T to 1e-11 K and v to 5e-12 m/s (the reference's own bar is 2.75e-12 K,
    test_1D_couette_fp.jl:90; the regenerated ziggurat tables and libm's pow / exp differ from Julia's in the last bits)."""
    r = ref[key]
    m, it = oracle.MASS["Ar"], oracle.interaction("Ar", "Ar")
    T_wall, v_wall, L, ndens, nx, dt = 300.0, 500.0, 5e-4, 5e22, 50, 2.59e-9
    V = L / nx
    Fnum = V * ndens / ppc
    grid, walls = (L, nx), (T_wall, T_wall, -v_wall, v_wall, 1.0, 1.0)
    pv, pia = oracle.OPV(ppc * nx), oracle.OPIA(nx, 1)
    rng = oracle.Rng.stable(1234)
    oracle.sample_equal_weight_grid(rng, grid, pv, pia, 1, m, ndens, T_wall, Fnum)
    F_cf = {"swpm": 1.0, "resort": Fnum * ppc / 100}.get(variant, Fnum)
    cf = oracle.CF(nx, oracle.estimate_sigma_g_w_max(it, m, m, T_wall, T_wall, F_cf))
    oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
    if thr:
        oracle.merge_octree_N2(rng, oc, pv, pia, 1, nx, 1, tgt, threshold=thr, grid=grid)
        oracle.squash_pia(pv, pia, 1)

    def check(rec, surf_rows):
        p = oracle.compute_props([pv], pia, [m])
        if key != "couette":  # the plain-Couette golden file predates np being filled by compute_props_sorted! (holds 1000 everywhere)
            assert np.array_equal(p.np[0], r["np"][rec]), rec
        assert np.array_equal(p.n[0], np.array(r["ndens"][rec])) or np.max(np.abs(p.n[0] / np.array(r["ndens"][rec]) - 1)) < 4e-16, rec
        assert np.max(np.abs(p.T[0] - np.array(r["T"][rec]))) < (1e-11 if variant == "fp" else 4.8e-13), rec
        assert np.max(np.abs(p.v[0] - np.array(r["v"][rec]))) < (5e-12 if variant == "fp" else 5e-13), rec
        if surf_rows is not None and "surf" in r:
            s = r["surf"]
            want = _surf_columns(s, s["timestep"].index(1000.0 * rec))
            assert np.array_equal(surf_rows[:, 0], want[:, 0])
            assert np.max(np.abs(surf_rows - want)) <= 1e-13 * np.max(np.abs(want)), rec

    check(0, None)
    for t in range(1, n_steps + 1):
        if variant == "fp":
            oracle.fp_linear(rng, it, m, pv, pia, 1, nx, 1, dt, V)
        elif not thr:
            oracle.ntc(rng, cf, it, pv, pia, 1, nx, 1, dt, V)
        else:  # the reference merges a cell right after colliding it, so the stream interleaves cell by cell (:80-88)
            for cell in range(1, nx + 1):
                if variant == "swpm":
                    oracle.swpm(rng, cf, it, pv, pia, cell, cell, 1, 1.5, dt, V)
                else:
                    oracle.ntc(rng, cf, it, pv, pia, cell, cell, 1, dt, V)
                if pia.indexer[0, cell - 1, 0] > thr:
                    oracle.merge_octree_N2(rng, oc, pv, pia, cell, cell, 1, tgt, grid=grid)
                    oracle.squash_pia(pv, pia, 1)
        s = oracle.convect_particles(rng, grid, walls, pv, pia, 1, [m], dt, surf=bool(thr))
        oracle.sort_particles(pv, pia, 1, grid=grid)
        if variant == "resort" and t % 500 == 0:
            oracle.restore_particle_ordering(pv)
            assert oracle.check_unique_index(pv, pia, 1) == (True, 0)
        if t % 1000 == 0:
            check(t // 1000, s)
    assert pia.check(1) == (True, 0) and oracle.check_unique_index(pv, pia, 1) == (True, 0)


def test_bkw_octree_first_merge_depends_on_the_last_bits_of_the_weights(oracle, ref):
    """Why the octree BKW replay needs the reference's exact arithmetic for the sampled weights: moving every weight by +-1 ulp leaves the
    post-merge count (7995) and the number of bins unchanged but shifts the merged moments by 1e-5 ... 1e-3 (tie-breaking between
    mirror-image bins) -- whereas the unperturbed oracle sits on the golden file to round-off."""
    m, it, T0, n_dens, tref, magic = _bkw_setup(oracle)

    def first_step(perturb_seed):
        pv, pia = oracle.OPV(40 ** 3), oracle.OPIA(1, 1)
        rng = oracle.Rng.stable(1234)
        n_s = int(oracle.sample_on_grid(rng, "bkw", pv, 40, m, T0, n_dens))
        pia.set_single_cell(1, 1, n_s)
        if perturb_seed:
            rows = pv.logical(1, n_s).copy()
            g = np.random.default_rng(perturb_seed)
            rows[:, 0] = np.nextafter(rows[:, 0], np.where(g.random(n_s) < 0.5, 0.0, np.inf))
            pv.set_logical(1, rows)
        oc = oracle.Octree(oracle.MID_SPLIT, oracle.INIT_MINMAX, oracle.BOUNDS_INHERIT, 6000, 10)
        cf = oracle.CF(1, oracle.estimate_sigma_g_w_max(it, m, m, T0, T0, n_dens / n_s))
        oracle.ntc(rng, cf, it, pv, pia, 1, 1, 1, 0.025 * tref, 1.0)
        oracle.merge_octree_N2(rng, oc, pv, pia, 1, 1, 1, 8000)
        p = oracle.compute_props([pv], pia, [m], [4, 6, 8, 10], Tref=T0, with_moments=True)
        return int(p.np[0, 0]), oc.Nbins, p.moments[0, 0].copy()

    base = first_step(0)
    golden = np.array(ref["bkw_vw_octree"]["first_step"]["moments"])
    assert np.max(np.abs(base[2] - golden)) < 1e-14
    for s in (1, 2, 3):
        n, nb, mom = first_step(s)
        assert n == base[0] == 7995 and nb == base[1]
        assert 1e-7 < np.max(np.abs(mom - golden)) < 2e-3


def test_julia_exp_table_and_accuracy(oracle):
    """mb_jlexp.h: the packed table equals its definition (tests/golden/make_jlexp_table.py, needs mpmath); the first entries are the
    literals of Julia's base/special/exp.jl (0xaac00b1afa5abcbe, 0x9b60163da9fb3335, ...)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "merzbild.jl_b200", "csrc", "mb_jlexp.h")).read()
    body = src[src.index("J_TABLE[256] = {"):src.index("};")]
    got = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ull", body)]
    assert len(got) == 256 and got[:5] == [0x0, 0xAAC00B1AFA5ABCBE, 0x9B60163DA9FB3335, 0xAB502168143B0280, 0xADC02C9A3E778060]
    pytest.importorskip("mpmath")
    import sys

    sys.path.insert(0, GOLDEN)
    from make_jlexp_table import table

    assert got == table()

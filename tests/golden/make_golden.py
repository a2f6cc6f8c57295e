#!/usr/bin/env python
"""Writes tests/golden/*.json from the reference checkout (/root/reference, only present in the build container).

The fixtures are DATA the reference's own tests hold for the hot path (no reference source code is copied):
  * sparta_couette.json    -- test/data/external/{avg,boundary}_seed_12345.grid.50000: the SPARTA time-averaged Couette profile
                              (50 cells, steps 14 001-50 000 of in.Couette) the reference validates its 1-D Couette runs against.
  * reference_vectors.json -- known-answer vectors quoted from the reference's test files (file:line given per entry) and
                              physical constants from data/*.toml, used to pin the CPU oracle.
  * reference_histories.json -- time histories OUTPUT BY THE REFERENCE ITSELF: its committed golden runs test/data/*.nc (netCDF-4 files
                              its regression tests compare against to 1e-13), read with the small HDF5 reader hdf5_min.py: 0-D temperature /
                              moment histories (2 species, BKW with every collision + merging variant) and the 1-D Couette snapshots
                              (cell profiles and wall properties every 1000 steps).  The oracle is held to them at distribution level
                              (different generator, same physics): tests/test_oracle_reference_runs.py.
Run:  python tests/golden/make_golden.py   (idempotent; commit the JSON it writes)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def sparta():
    ext = os.path.join(REF, "test", "data", "external")
    rows = []
    with open(os.path.join(ext, "avg_seed_12345.grid.50000")) as f:
        lines = f.read().splitlines()
    k = next(i for i, l in enumerate(lines) if l.startswith("ITEM: CELLS"))
    cols = lines[k].split()[2:]
    for l in lines[k + 1:]:
        if l.strip():
            rows.append([float(x) for x in l.split()])
    with open(os.path.join(ext, "boundary_seed_12345.grid.50000")) as f:
        b = [l for l in f.read().splitlines() if l and not l.startswith("#")]
    brow = [[float(x) for x in l.split()] for l in b[1:3]]
    inp = open(os.path.join(ext, "in.Couette")).read()
    return {
        "source": "test/data/external/avg_seed_12345.grid.50000, boundary_seed_12345.grid.50000, in.Couette",
        "columns": cols, "note": "f_1 = thermal/grid temp, press; grid n, nrho, u, v (in.Couette: compute 1, compute 2, fix 1 ave/grid 1 36000 50000)",
        "cells": rows,
        "boundary_columns": ["row", "nflux", "mflux", "press", "shx", "shy", "shz", "ke"],
        "boundary": brow,
        "setup": {"L": 5e-4, "nx": 50, "fnum": float(re.search(r"fnum\s+(\S+)", inp).group(1)), "nrho": 5e22, "T_init": 273.0, "T_wall": 300.0,
                  "v_wall": 500.0, "dt": float(re.search(r"timestep\s+(\S+)", inp).group(1)), "n_steps": 50000, "avg_steps": 36000},
    }


def toml_table(path):
    out, cur = {}, None
    for l in open(path).read().splitlines():
        l = l.split("#")[0].strip()
        if not l:
            continue
        m = re.match(r"\[\"?([^\]\"]+)\"?\]", l)
        if m:
            cur = out.setdefault(m.group(1), {})
            continue
        k, v = [s.strip() for s in l.split("=", 1)]
        try:
            cur[k] = float(v)
        except ValueError:
            cur[k] = v.strip('"')
    return out


def vectors():
    return {
        "particles_toml": {k: v for k, v in toml_table(os.path.join(REF, "data", "particles.toml")).items() if k in ("Ar", "He")},
        "vhs_toml": toml_table(os.path.join(REF, "data", "vhs.toml")),
        "pseudo_maxwell_toml": toml_table(os.path.join(REF, "data", "pseudo_maxwell.toml")),
        "grid_sorting": {
            "source": "test/test_grid_sorting.jl:26-54",
            "index_after_sort": [[5, 6, 7, 8, 1, 2, 3, 4], [7, 8, 5, 6, 3, 4, 1, 2], [6, 7, 8, 5, 1, 2, 3, 4]],
        },
        "convection_specular": {
            "source": "test/test_convection_1D.jl:1-75", "L": 50.0, "dt": 2.0, "x_after": [20.5, 29.0, 23.0, 3.55], "index_after_sort": [4, 1, 3, 2],
        },
        "computes": {"source": "test/test_computes.jl:6-67", "mixed_moments": [-2.0, 4.0, -8.0, -128.0]},
        "collision_utils": {"source": "test/test_collision_utils.jl:13-40", "v_com": [1.0, 0.0, -0.5], "g": 3.0},
        "octree_24": {"source": "test/test_octree_merging.jl:3-163", "bins": 8, "np_per_bin": 3, "post_merge_np": [16, 2]},
        "bkw": {"source": "test/test_bkw.jl:4-29,108-118", "magic_factor_Ar": 1.59577, "tolerances": {"4": 0.05, "6": 0.055, "8": 0.15}},
        "two_species": {"source": "test/test_2species.jl:15-25,92-94", "T_eq": 600.0, "tolerance": 0.12},
        "bkw_vw_octree": {"source": "test/test_bkw_varweight_octree.jl:104-106", "T_abs_tol_K": 5e-4, "ndens_rel_tol": 1e-11},
    }


def histories():
    import sys

    sys.path.insert(0, HERE)
    from hdf5_min import H5File

    data = os.path.join(REF, "test", "data")

    def rd(name):
        return H5File(os.path.join(data, name + ".nc"))

    def lst(a):
        return [float(x) for x in a.ravel()] if a.ndim == 1 else [lst(r) for r in a]

    out = {"note": "arrays are in netCDF (C) order of the reference files: [timestep][species][cell](component)"}
    # 0-D, two species (test/test_2species.jl, test/test_2species_varweight_octree.jl): every 25th of 801 records
    for key, name in (("two_species", "2species_seed1234"), ("two_species_varweight_octree", "2species_varweight_octree_seed1234")):
        f = rd(name)
        sl = slice(0, 801, 25)
        out[key] = {"source": f"test/data/{name}.nc", "timestep": lst(f.read("timestep")[sl]), "T": lst(f.read("T")[sl, :, 0]),
                    "np": lst(f.read("np")[sl, :, 0]), "ndens": lst(f.read("ndens")[sl, :, 0]), "v": lst(f.read("v")[sl, :, 0, :])}
    # 0-D BKW (test/test_bkw.jl, test_bkw_varweight_octree.jl, test_bkw_varweight_grid.jl, test_bkw_varweight_octree_swpm.jl): every 10th of 501
    for key, name in (("bkw_20k", "bkw_20k_seed1234"), ("bkw_vw_octree", "bkw_vw_octree_seed1234"), ("bkw_vw_grid", "bkw_vw_grid_seed1234"),
                      ("bkw_vw_octree_swpm", "bkw_vw_octree_swpm_seed1234")):
        f = rd(name)
        sl = slice(0, 501, 10)
        out[key] = {"source": f"test/data/{name}.nc", "timestep": lst(f.read("timestep")[sl]), "moment_powers": [int(x) for x in f.read("moment_powers")],
                    "moments": lst(f.read("moments")[sl, 0, 0, :]), "np": lst(f.read("np")[sl, 0, 0]), "T": lst(f.read("T")[sl, 0, 0]),
                    "ndens": lst(f.read("ndens")[sl, 0, 0]), "v": lst(f.read("v")[sl, 0, 0, :]),
                    "first_step": {"np": int(f.read("np")[1, 0, 0]), "T": float(f.read("T")[1, 0, 0]), "ndens": float(f.read("ndens")[1, 0, 0]),
                                   "moments": lst(f.read("moments")[1, 0, 0, :])}}
    # 1-D Couette, 50 cells (test/test_1D_couette*.jl): snapshots every 1000 steps, cell profiles and wall properties
    pre = "couette_0.0005_50_500.0_300.0_"
    for key, name in (("couette", "1000"), ("couette_vw200to150", "1000_vw200to150"), ("couette_vw200to150_swpm", "1000_vw200to150_swpm"),
                      ("couette_fp_linear", "100_fp_linear"), ("couette_vw150to100_resort", "500_vw150to100_resort")):
        f = rd(pre + name)
        out[key] = {"source": f"test/data/{pre}{name}.nc", "timestep": lst(f.read("timestep")), "np": lst(f.read("np")[:, 0, :]),
                    "ndens": lst(f.read("ndens")[:, 0, :]), "T": lst(f.read("T")[:, 0, :]), "v": lst(f.read("v")[:, 0, :, :])}
        sname = pre + name + "_surf"
        if os.path.exists(os.path.join(data, sname + ".nc")):
            g = rd(sname)
            out[key]["surf"] = {"source": f"test/data/{sname}.nc", "timestep": lst(g.read("timestep")),
                                **{k: lst(g.read(k)[:, 0]) for k in ("np", "flux_incident", "flux_reflected", "force", "normal_pressure",
                                                                     "shear_pressure", "kinetic_energy_flux")}}
    return out


if __name__ == "__main__":
    for name, obj in (("sparta_couette.json", sparta()), ("reference_vectors.json", vectors()), ("reference_histories.json", histories())):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, indent=1)
        print("wrote", name)

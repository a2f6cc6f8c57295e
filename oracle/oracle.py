"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/mb_oracle*.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module.  The product package (merzbild.jl_b200/merzbild_b200) never does.

All indices are 1-based and inclusive exactly as in the reference (Merzbild.jl); ``OPIA.indexer[s, c]`` is the
7-tuple (n_local, start1, end1, n_group1, start2, end2, n_group2) of ``pia.indexer[c+1, s+1]``
(/root/reference/src/particles.jl:56-66).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmb_oracle.so")

# field order of ParticleIndexer (particles.jl:56-66)
N_LOCAL, START1, END1, N_GROUP1, START2, END2, N_GROUP2 = range(7)


def build(force=False):
    """Compile the oracle with g++ (Makefile in this directory)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) or f == "Makefile"]
    if not force and os.path.exists(_SO) and os.path.exists(os.path.join(_HERE, "_build", "couette_cpu")):
        newest = max(os.path.getmtime(s) for s in srcs)
        if os.path.getmtime(_SO) >= newest:
            return _SO
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _SO


_lib = None


class RngSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("seq", C.c_void_p), ("seed", C.c_uint64), ("timestep", C.c_uint32), ("substream", C.c_uint32)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    i64, f64, vp, i32 = C.c_int64, C.c_double, C.c_void_p, C.c_int32
    pd = C.POINTER(C.c_double)
    pi = C.POINTER(C.c_int64)
    prs = C.POINTER(RngSpec)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("mbo_philox4x32_10", None, vp, vp, vp)
    sig("mbo_philox_stream_doubles", None, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, i64, vp)
    sig("mbo_rng_create", vp, C.c_uint64)
    sig("mbo_rng_create_stable", vp, C.c_uint64)
    sig("mbo_rng_free", None, vp)
    sig("mbo_rng_rand", f64, vp)
    sig("mbo_pv_create", vp, i64)
    sig("mbo_pv_free", None, vp)
    sig("mbo_pv_length", i64, vp)
    sig("mbo_pv_resize", None, vp, i64)
    sig("mbo_pv_particles", pd, vp)
    for n in ("index", "cell", "buffer", "nbuffer"):
        sig("mbo_pv_" + n, pi, vp)
    sig("mbo_pv_add_particle", None, vp, i64, f64, vp, vp)
    sig("mbo_pv_update_buffer_new_particle", None, vp, i64)
    sig("mbo_pv_get_logical", None, vp, i64, i64, vp)
    sig("mbo_pv_set_logical", None, vp, i64, i64, vp)
    sig("mbo_pia_create", vp, i64, i64)
    sig("mbo_pia_free", None, vp)
    sig("mbo_pia_indexer", pi, vp)
    sig("mbo_pia_n_total", pi, vp)
    sig("mbo_pia_contiguous", C.POINTER(C.c_uint8), vp)
    sig("mbo_map_cont_index", i64, vp, i64, i64, i64)
    sig("mbo_update_particle_indexer_new_lower_count", None, vp, i64, i64, i64)
    sig("mbo_update_particle_indexer_new_particle", None, vp, i64, i64)
    sig("mbo_update_buffer_index_new_particle", None, vp, vp, i64, i64)
    sig("mbo_delete_particle", None, vp, vp, i64, i64, i64)
    for n in ("mbo_delete_particle_end", "mbo_delete_particle_end_group1", "mbo_delete_particle_end_group2"):
        sig(n, None, vp, vp, i64, i64)
    sig("mbo_squash_pia", None, vp, vp, i64)
    sig("mbo_restore_particle_ordering", None, vp)
    sig("mbo_check_pia_is_correct", C.c_int, vp, i64, pi)
    sig("mbo_check_unique_index", C.c_int, vp, vp, i64, pi)
    sig("mbo_grid_params", None, f64, i64, f64, vp)
    sig("mbo_sort_particles_grid", None, f64, i64, vp, vp, i64)
    sig("mbo_sort_particles_cells", None, vp, vp, i64)
    sig("mbo_make_interaction", None, f64, f64, f64, f64, f64, vp)
    sig("mbo_sigma_vhs", f64, vp, f64)
    sig("mbo_estimate_sigma_g_w_max", f64, vp, f64, f64, f64, f64, f64, f64)
    sig("mbo_compute_com_g", None, vp, vp, vp, vp, vp)
    sig("mbo_scatter_vhs", None, prs, vp, vp, vp)
    sig("mbo_collide_2particles_vhs", None, prs, vp, vp, vp, i64, i64, i64, i64, f64, C.c_int, vp)
    sig("mbo_compute_octant", C.c_int, vp, vp)
    sig("mbo_octree_vel_middle", None, vp, vp)
    sig("mbo_octree_compute_v_mean", None, vp, i64, i64, vp)
    sig("mbo_octree_bounds_recompute", None, vp, i64, i64, i64, vp)
    sig("mbo_ntc", None, prs, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, f64, f64, f64, C.c_int)
    sig("mbo_ntc2", None, prs, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, f64, f64, f64, C.c_int)
    sig("mbo_swpm", None, prs, vp, vp, vp, vp, vp, vp, i64, i64, i64, f64, f64, f64)
    sig("mbo_scale_norm_rands", None, vp, vp, vp, i64)
    sig("mbo_fp_linear", None, prs, vp, f64, vp, vp, i64, i64, i64, f64, f64)
    sig("mbo_compute_props", None, vp, vp, vp, i64, vp, f64, C.c_int, vp, vp, vp, vp, vp, vp)
    sig("mbo_compute_props_sorted", None, vp, vp, vp, i64, i64, C.c_int, f64, i64, vp, vp, vp, vp)
    sig("mbo_avg_props", None, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f64)
    sig("mbo_compute_mixed_moment", f64, vp, vp, i64, i64, vp, f64, f64)
    sig("mbo_convect_particles", None, prs, f64, i64, vp, vp, vp, i64, vp, i64, vp, f64, C.c_int)
    sig("mbo_sample_equal_weight_grid", None, vp, f64, i64, vp, vp, i64, f64, f64, f64, f64, i64, i64)
    sig("mbo_sample_equal_weight_cell", None, vp, vp, vp, i64, i64, i64, f64, f64, f64, vp, C.c_int, vp)
    sig("mbo_sample_on_grid", i64, vp, C.c_int, vp, i64, f64, f64, f64, vp, f64, f64, f64, vp)
    sig("mbo_sample_equal_weight_cells", None, vp, vp, vp, vp, i64, i64, i64, i64, f64, f64, f64, f64, vp, C.c_int, vp)
    sig("mbo_surface_props_kat", None, vp, vp, i64, C.c_int, f64, f64, vp, vp)
    sig("mbo_gridmerge_create", vp, i64, i64, i64, vp)
    sig("mbo_gridmerge_free", None, vp)
    sig("mbo_gridmerge_index", i64, vp, vp, vp)
    sig("mbo_gridmerge_cell", None, vp, i64, vp)
    sig("mbo_merge_grid_based", i64, vp, vp, vp, vp, i64, i64, i64, f64, i64, vp, vp, f64, i64)
    sig("mbo_sample_on_grid_cells", i64, vp, C.c_int, vp, vp, i64, i64, i64, i64, f64, f64, f64, vp, f64, f64, f64, vp)
    sig("mbo_octree_create", vp, C.c_int, C.c_int, C.c_int, i64, i64)
    sig("mbo_octree_free", None, vp)
    sig("mbo_octree_nbins", i64, vp)
    sig("mbo_octree_n_particles", i64, vp)
    sig("mbo_octree_total_post_merge_np", i64, vp)
    sig("mbo_octree_bin", None, vp, i64, vp)
    sig("mbo_octree_full_bin", None, vp, i64, vp)
    sig("mbo_octree_particle_indexes_sorted", pi, vp)
    sig("mbo_octree_init", None, vp, vp, vp, i64, i64)
    sig("mbo_octree_split_bin", None, vp, i64, vp)
    sig("mbo_octree_compute_bin_props", None, vp, i64, vp)
    sig("mbo_octree_compute", None, vp, vp, i64)
    sig("mbo_merge_octree_N2", None, prs, vp, vp, vp, i64, i64, i64, i64, i64, f64, i64, C.c_int)
    sig("mbo_exchanger_create", vp, i64, i64)
    sig("mbo_exchanger_free", None, vp)
    sig("mbo_exchanger_indexer", pi, vp)
    sig("mbo_exchanger_reset", None, vp, i64)
    sig("mbo_exchange_particles", None, vp, vp, vp, vp, vp, i64, i64, i64, i64)
    sig("mbo_sort_particles_after_exchange", None, vp, vp, vp, i64, i64, i64)
    sig("mbo_generate_1_factorization", i64, i64, vp, vp)
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# physical constants / data (data/particles.toml, data/vhs.toml, data/pseudo_maxwell.toml of the reference)
K_B = 1.380649e-23
MASS = {"Ar": 66.3e-27, "He": 6.65e-27}
VHS = {("Ar", "Ar"): (4.11e-10, 0.81, 273.0), ("Ar", "He"): (3.25e-10, 0.735, 273.0), ("He", "He"): (2.33e-10, 0.66, 273.0)}
PSEUDO_MAXWELL = {("Ar", "Ar"): (4.11e-10, 1.0, 273.0), ("Ar", "He"): (3.25e-10, 1.0, 273.0), ("He", "He"): (2.33e-10, 1.0, 273.0)}


class Rng:
    """Either one sequential generator (``Rng.seq(seed)``, like the reference's single rng object) or per-entity
    Philox streams (``Rng.philox(seed, timestep, substream)``, the convention shared with the CUDA path)."""

    def __init__(self, spec, handle=None):
        self.spec = spec
        self._h = handle

    @staticmethod
    def seq(seed=1234):
        h = lib().mbo_rng_create(seed)
        return Rng(RngSpec(0, h, 0, 0, 0), h)

    @staticmethod
    def stable(seed=1234):
        """StableRNGs.jl ``StableRNG(seed)``: the generator of the reference's test suite (oracle/philox.hpp)."""
        h = lib().mbo_rng_create_stable(seed)
        return Rng(RngSpec(0, h, 0, 0, 0), h)

    @staticmethod
    def philox(seed, timestep=0, substream=0):
        return Rng(RngSpec(1, None, seed, timestep, substream))

    def at(self, timestep, substream=None):
        if self.spec.kind == 0:
            return self
        return Rng(RngSpec(1, None, self.spec.seed, timestep, self.spec.substream if substream is None else substream))

    def rand(self):
        assert self.spec.kind == 0
        return lib().mbo_rng_rand(self._h)

    @property
    def ref(self):
        return C.byref(self.spec)

    def __del__(self):
        if self._h is not None and _lib is not None:
            _lib.mbo_rng_free(self._h)
            self._h = None


class OPV:
    """ParticleVector of the reference (particles.jl:194-212): AoS particles + index/cell/buffer."""

    def __init__(self, n):
        self.h = lib().mbo_pv_create(int(n))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mbo_pv_free(self.h)
            self.h = None

    def __len__(self):
        return lib().mbo_pv_length(self.h)

    def resize(self, n):
        lib().mbo_pv_resize(self.h, int(n))

    def _view(self, fn, shape, dtype=np.int64):
        ptr = fn(self.h)
        return np.ctypeslib.as_array(ptr, shape=shape)

    @property
    def particles(self):  # physical storage, rows (w, vx, vy, vz, x, y, z)
        return self._view(lib().mbo_pv_particles, (len(self), 7))

    @property
    def index(self):
        return self._view(lib().mbo_pv_index, (len(self),))

    @property
    def cell(self):
        return self._view(lib().mbo_pv_cell, (len(self),))

    @property
    def buffer(self):
        return self._view(lib().mbo_pv_buffer, (len(self),))

    @property
    def nbuffer(self):
        return int(lib().mbo_pv_nbuffer(self.h)[0])

    @nbuffer.setter
    def nbuffer(self, v):
        lib().mbo_pv_nbuffer(self.h)[0] = int(v)

    def logical(self, lo=1, hi=None):
        """rows pv[lo..hi] (1-based inclusive) through the index indirection (particles.jl:225)."""
        hi = len(self) if hi is None else hi
        out = np.empty((max(hi - lo + 1, 0), 7))
        if hi >= lo:
            lib().mbo_pv_get_logical(self.h, lo, hi, _p(out))
        return out

    def set_logical(self, lo, rows):
        rows = _f64(rows).reshape(-1, 7)
        lib().mbo_pv_set_logical(self.h, lo, lo + rows.shape[0] - 1, _p(rows))

    def add_particle(self, position, w, v, x):
        v = _f64(v)
        x = _f64(x)
        lib().mbo_pv_add_particle(self.h, position, w, _p(v), _p(x))

    def fill_identity(self, rows):
        """Place rows at logical positions 1..n the way repeated add_particle!(pv, i, ...) does."""
        rows = _f64(rows).reshape(-1, 7)
        for i, r in enumerate(rows):
            self.add_particle(i + 1, r[0], r[1:4], r[4:7])


class OPIA:
    """ParticleIndexerArray (particles.jl:104-174)."""

    def __init__(self, n_cells, n_species=1):
        self.n_cells, self.n_species = int(n_cells), int(n_species)
        self.h = lib().mbo_pia_create(self.n_cells, self.n_species)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mbo_pia_free(self.h)
            self.h = None

    @property
    def indexer(self):  # [species, cell, 7]
        return np.ctypeslib.as_array(lib().mbo_pia_indexer(self.h), shape=(self.n_species, self.n_cells, 7))

    @property
    def n_total(self):
        return np.ctypeslib.as_array(lib().mbo_pia_n_total(self.h), shape=(self.n_species,))

    @property
    def contiguous(self):
        return np.ctypeslib.as_array(lib().mbo_pia_contiguous(self.h), shape=(self.n_species,))

    def set_single_cell(self, species, cell, n):
        """ParticleIndexer(n) into [cell, species] (particles.jl:77)."""
        self.indexer[species - 1, cell - 1] = (n, 1, n, n, 0, -1, 0) if n > 0 else (0, 0, -1, 0, 0, -1, 0)
        self.n_total[species - 1] = n

    def check(self, species=1):
        where = C.c_int64(0)
        ok = lib().mbo_check_pia_is_correct(self.h, species, C.byref(where))
        return bool(ok), int(where.value)


def check_unique_index(pv, pia, species=1):
    code = C.c_int64(0)
    ok = lib().mbo_check_unique_index(pv.h, pia.h, species, C.byref(code))
    return bool(ok), int(code.value)


def grid_params(L, nx, wall_offset=1e-12):
    out = np.empty(5)
    lib().mbo_grid_params(L, nx, wall_offset, _p(out))
    return dict(dx=out[0], inv_dx=out[1], min_x=out[2], max_x=out[3], L=out[4], n_cells=nx)


def make_interaction(m_i, m_k, d, o, Tref):
    """One entry of load_interaction_data (collision_utils.jl:159-201) -> 8 doubles (m_r, mu1, mu2, d, o, Tref, muref, factor)."""
    out = np.empty(8)
    lib().mbo_make_interaction(m_i, m_k, d, o, Tref, _p(out))
    return out


def interaction(sp1, sp2, table=None):
    table = VHS if table is None else table
    d, o, Tref = table.get((sp1, sp2)) or table[(sp2, sp1)]
    return make_interaction(MASS[sp1], MASS[sp2], d, o, Tref)


def sigma_vhs(it, g):
    return lib().mbo_sigma_vhs(_p(it), g)


def estimate_sigma_g_w_max(it, m1, m2, T1, T2, Fnum, mult=1.0):
    return lib().mbo_estimate_sigma_g_w_max(_p(it), m1, m2, T1, T2, Fnum, mult)


def sort_particles(pv, pia, species=1, grid=None):
    """sort_particles! (grid_sorting.jl:58 with grid=(L, nx); :128 with grid=None, cells known)."""
    if grid is not None:
        lib().mbo_sort_particles_grid(grid[0], grid[1], pv.h, pia.h, species)
    else:
        lib().mbo_sort_particles_cells(pv.h, pia.h, species)


def squash_pia(pv, pia, species=1):
    lib().mbo_squash_pia(pv.h, pia.h, species)


def restore_particle_ordering(pv):
    lib().mbo_restore_particle_ordering(pv.h)


class CF:
    """Per-cell CollisionFactors for one species pair (collision_ntc.jl:18-25) as arrays over cells."""

    def __init__(self, n_cells, sigma_g_w_max=0.0):
        self.sigma_g_w_max = np.full(n_cells, sigma_g_w_max, dtype=np.float64)
        self.n_coll = np.zeros(n_cells, dtype=np.int64)
        self.n_coll_performed = np.zeros(n_cells, dtype=np.int64)
        self.n_eq_w_coll_performed = np.zeros(n_cells, dtype=np.int64)


def ntc(rng, cf, it, pv, pia, cell_lo, cell_hi, species, dt, V, dw_tol=1e-16, equal_weight=False):
    lib().mbo_ntc(rng.ref, _p(cf.sigma_g_w_max), _p(cf.n_coll), _p(cf.n_coll_performed), _p(cf.n_eq_w_coll_performed), _p(it), pv.h, pia.h,
                  cell_lo, cell_hi, species, dt, V, dw_tol, int(equal_weight))


def ntc2(rng, cf, it, pv1, pv2, pia, cell_lo, cell_hi, s1, s2, dt, V, dw_tol=1e-16, equal_weight=False):
    lib().mbo_ntc2(rng.ref, _p(cf.sigma_g_w_max), _p(cf.n_coll), _p(cf.n_coll_performed), _p(cf.n_eq_w_coll_performed), _p(it), pv1.h, pv2.h,
                   pia.h, cell_lo, cell_hi, s1, s2, dt, V, dw_tol, int(equal_weight))


def swpm(rng, cf, it, pv, pia, cell_lo, cell_hi, species, G, dt, V):
    lib().mbo_swpm(rng.ref, _p(cf.sigma_g_w_max), _p(cf.n_coll), _p(cf.n_coll_performed), _p(it), pv.h, pia.h, cell_lo, cell_hi, species, G, dt, V)


def scale_norm_rands(x, y, z):
    x, y, z = _f64(x).copy(), _f64(y).copy(), _f64(z).copy()
    lib().mbo_scale_norm_rands(_p(x), _p(y), _p(z), len(x))
    return x, y, z


def fp_linear(rng, it, mass, pv, pia, cell_lo, cell_hi, species, dt, V):
    lib().mbo_fp_linear(rng.ref, _p(it), mass, pv.h, pia.h, cell_lo, cell_hi, species, dt, V)


def _handles(pvs):
    arr = (C.c_void_p * len(pvs))(*[p.h for p in pvs])
    return arr


class Props:
    def __init__(self, n_cells, n_species, n_moments=0):
        self.lpa = np.zeros(n_species)
        self.np = np.zeros((n_species, n_cells))
        self.n = np.zeros((n_species, n_cells))
        self.v = np.zeros((n_species, n_cells, 3))
        self.T = np.zeros((n_species, n_cells))
        self.moments = np.zeros((n_species, n_cells, max(n_moments, 1)))


def compute_props(pvs, pia, masses, moment_powers=(), Tref=300.0, with_moments=False):
    """compute_props! (physical_props.jl:104) / compute_props_with_total_moments! (:168)."""
    pw = np.asarray(moment_powers, dtype=np.int32)
    out = Props(pia.n_cells, pia.n_species, len(pw))
    masses = _f64(masses)
    lib().mbo_compute_props(_handles(pvs), pia.h, _p(masses), len(pw), _p(pw), Tref, int(with_moments), _p(out.lpa), _p(out.np), _p(out.n),
                            _p(out.v), _p(out.T), _p(out.moments))
    return out


def compute_props_sorted(pvs, pia, masses, cell_lo=1, cell_hi=None, grid=None, out=None):
    """compute_props_sorted! (physical_props.jl:317, :393 with grid=(L, nx) -> number density)."""
    out = Props(pia.n_cells, pia.n_species) if out is None else out
    cell_hi = pia.n_cells if cell_hi is None else cell_hi
    masses = _f64(masses)
    L, nx = grid if grid is not None else (0.0, 0)
    lib().mbo_compute_props_sorted(_handles(pvs), pia.h, _p(masses), cell_lo, cell_hi, int(grid is not None), L, nx, _p(out.np), _p(out.n), _p(out.v),
                                   _p(out.T))
    return out


def avg_props(avg, props, n_avg_timesteps):
    """avg_props!(phys_props_avg, phys_props, n_avg_timesteps) (physical_props.jl:281-299), in place on ``avg`` (a Props)."""
    ns, nc = props.np.shape
    lib().mbo_avg_props(nc, ns, _p(avg.lpa), _p(avg.np), _p(avg.n), _p(avg.v), _p(avg.T), _p(props.lpa), _p(props.np), _p(props.n), _p(props.v),
                        _p(props.T), float(n_avg_timesteps))
    return avg


def compute_mixed_moment(pv, pia, cell, species, powers, sum_scaler=1.0, res_scaler=1.0):
    pw = np.asarray(powers, dtype=np.int32)
    return lib().mbo_compute_mixed_moment(pv.h, pia.h, cell, species, _p(pw), sum_scaler, res_scaler)


def convect_particles(rng, grid, walls, pv, pia, species, masses, dt, surf=False, compute_cell=False):
    """convect_particles! / convect_particles_and_compute_cell! (convection_1D.jl:130,176,225,274).
    grid=(L, nx); walls=(T_l, T_r, vy_l, vy_r, acc_l, acc_r). Returns the 2x11 SurfProps rows if surf."""
    masses = _f64(masses)
    walls = _f64(walls)
    s = np.zeros((2, 11)) if surf else None
    lib().mbo_convect_particles(rng.ref, grid[0], grid[1], _p(walls), pv.h, pia.h, species, _p(masses), len(masses), _p(s), dt, int(compute_cell))
    return s


def sample_equal_weight_grid(rng, grid, pv, pia, species, mass, ndens, T, Fnum, cell_lo=1, cell_hi=None):
    cell_hi = grid[1] if cell_hi is None else cell_hi
    lib().mbo_sample_equal_weight_grid(rng._h, grid[0], grid[1], pv.h, pia.h, species, mass, ndens, T, Fnum, cell_lo, cell_hi)


def sample_equal_weight_cell(rng, pv, pia, cell, species, n, m, T, Fnum, box=(0, 1, 0, 1, 0, 1), distribution="Maxwellian", v0=(0, 0, 0)):
    box = _f64(box)
    v0 = _f64(v0)
    lib().mbo_sample_equal_weight_cell(rng._h, pv.h, pia.h, cell, species, n, m, T, Fnum, _p(box), 0 if distribution == "Maxwellian" else 1, _p(v0))


def sample_on_grid(rng, vdf, pv, nv, m, T, n_total, box=(0, 1, 0, 1, 0, 1), v_mult=3.5, cutoff_mult=3.5, noise=0.0, v_offset=(0, 0, 0)):
    box = _f64(box)
    vo = _f64(v_offset)
    return lib().mbo_sample_on_grid(rng._h, 0 if vdf == "maxwellian" else 1, pv.h, nv, m, T, n_total, _p(box), v_mult, cutoff_mult, noise, _p(vo))


def sample_equal_weight_cells(rng, pv, pia, cell_lo, cell_hi, species, nparticles, m, T, Fnum, grid=None, ndens=0.0, box=(0, 1, 0, 1, 0, 1),
                              distribution="Maxwellian", v0=(0, 0, 0)):
    """sample_particles_equal_weight! for cells cell_lo..cell_hi; with ``Rng.philox`` one stream per cell (the device convention).
    grid = (L, nx): the 1-D grid variants (nparticles < 0: number-density variant)."""
    box, v0 = _f64(box), _f64(v0)
    g = None if grid is None else _f64([grid[0], grid[1]])
    lib().mbo_sample_equal_weight_cells(rng.ref, _p(g), pv.h, pia.h, cell_lo, cell_hi, species, nparticles, ndens, m, T, Fnum, _p(box),
                                        0 if distribution == "Maxwellian" else 1, _p(v0))


def sample_on_grid_cells(rng, vdf, pv, pia, cell_lo, cell_hi, species, nv, m, T, n_total, box=(0, 1, 0, 1, 0, 1), v_mult=3.5, cutoff_mult=3.5, noise=0.0,
                         v_offset=(0, 0, 0)):
    box, vo = _f64(box), _f64(v_offset)
    return lib().mbo_sample_on_grid_cells(rng.ref, 0 if vdf == "maxwellian" else 1, pv.h, pia.h, cell_lo, cell_hi, species, nv, m, T, n_total, _p(box),
                                          v_mult, cutoff_mult, noise, _p(vo))


MID_SPLIT, MEAN_SPLIT = 1, 2
INIT_MINMAX, INIT_MINMAX_SYM, INIT_C = 1, 2, 3
BOUNDS_INHERIT, BOUNDS_RECOMPUTE = 1, 2


class Octree:
    """OctreeN2Merge (merging_octree_N2.jl:131-246)."""

    def __init__(self, split=MID_SPLIT, init_bin_bounds=INIT_MINMAX, bin_bounds_compute=BOUNDS_INHERIT, max_Nbins=4096, max_depth=10):
        self.h = lib().mbo_octree_create(split, init_bin_bounds, bin_bounds_compute, max_Nbins, max_depth)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mbo_octree_free(self.h)
            self.h = None

    @property
    def Nbins(self):
        return lib().mbo_octree_nbins(self.h)

    @property
    def n_particles(self):
        return lib().mbo_octree_n_particles(self.h)

    @property
    def total_post_merge_np(self):
        return lib().mbo_octree_total_post_merge_np(self.h)

    def bin(self, i):
        o = np.empty(12)
        lib().mbo_octree_bin(self.h, i, _p(o))
        return dict(np=int(o[0]), w=o[1], v_min=o[2:5].copy(), v_max=o[5:8].copy(), depth=int(o[8]), can_be_refined=bool(o[9]), start=int(o[10]),
                    end=int(o[11]))

    def full_bin(self, i):
        o = np.empty(12)
        lib().mbo_octree_full_bin(self.h, i, _p(o))
        return dict(v_mean=o[0:3].copy(), v_std_sq=o[3:6].copy(), x_mean=o[6:9].copy(), x_std_sq=o[9:12].copy())

    def particle_indexes_sorted(self, n):
        return np.ctypeslib.as_array(lib().mbo_octree_particle_indexes_sorted(self.h), shape=(n,)).copy()

    def init(self, pv, pia, cell=1, species=1):
        lib().mbo_octree_init(self.h, pv.h, pia.h, cell, species)

    def split_bin(self, bin_id, pv):
        lib().mbo_octree_split_bin(self.h, bin_id, pv.h)

    def compute_bin_props(self, bin_id, pv):
        lib().mbo_octree_compute_bin_props(self.h, bin_id, pv.h)

    def compute(self, pv, target_np):
        lib().mbo_octree_compute(self.h, pv.h, target_np)

    @property
    def vel_middle(self):
        o = np.empty(3)
        lib().mbo_octree_vel_middle(self.h, _p(o))
        return o

    def compute_v_mean(self, bs, be, pv):
        lib().mbo_octree_compute_v_mean(self.h, bs, be, pv.h)

    def bin_bounds_recompute(self, bin_id, bs, be, pv):
        lib().mbo_octree_bounds_recompute(self.h, bin_id, bs, be, pv.h)


def surface_props_kat(rows, ops, scale=None):
    """update_surface_incident! / update_surface_reflected! (surface_props.jl:77-131) applied in the order of ``ops`` =
    [(kind, element, row)], kind 0 incident / 1 reflected; scale = (mass, dt, inv_areas) applies surface_props_scale! (:144-160).
    Returns the 2 x 11 rows (np, flux_incident, flux_reflected, force[3], normal_pressure, shear_pressure[3], kinetic_energy_flux)."""
    rows = _f64(np.asarray(rows).reshape(-1, 7))
    o = np.ascontiguousarray(np.asarray(ops, dtype=np.int64).reshape(-1, 3))
    out = np.zeros((2, 11))
    m, dt, ia = scale if scale is not None else (0.0, 1.0, (1.0, 1.0))
    ia = _f64(ia)
    lib().mbo_surface_props_kat(_p(rows), _p(o), o.shape[0], int(scale is not None), m, dt, _p(ia), _p(out))
    return out


def compute_octant(v, mid):
    """compute_octant (merging_octree_N2.jl:304-316)"""
    v, mid = _f64(v), _f64(mid)
    return lib().mbo_compute_octant(_p(v), _p(mid))


def collide_2particles_vhs(rng, it, pv, pia, i, k, cell=1, species=1, dw_tol=1e-16, equal_weight=False, sigma_g_w_max=0.0):
    """collide_2particles_vhs! (collision_ntc.jl:223-270) / ..._equal_weight! (:294-309) after compute_g!; returns
    (sigma_g_w_max, n_coll_performed, n_eq_w_coll_performed)."""
    cf = _f64([sigma_g_w_max, 0.0, 0.0, 0.0])
    lib().mbo_collide_2particles_vhs(rng.ref, _p(it), pv.h, pia.h, i, k, cell, species, dw_tol, int(equal_weight), _p(cf))
    return cf[0], int(cf[1]), int(cf[2])


def merge_octree_N2(rng, oc, pv, pia, cell_lo, cell_hi, species, target_np, threshold=-1, grid=None, squash_after_each=False):
    """merge_octree_N2_based! (merging_octree_N2.jl:1060,1088) over the cells with n_local > threshold."""
    L, nx = grid if grid is not None else (0.0, 0)
    lib().mbo_merge_octree_N2(rng.ref, oc.h, pv.h, pia.h, cell_lo, cell_hi, species, threshold, target_np, L, nx, int(squash_after_each))


class GridMerge:
    """GridN2Merge(Nx, Ny, Nz, extent_multiplier) (merging_grid.jl:72-116)."""

    def __init__(self, Nx, Ny=None, Nz=None, extent_multiplier=3.5):
        Ny = Nx if Ny is None else Ny
        Nz = Nx if Nz is None else Nz
        m = _f64(np.broadcast_to(np.asarray(extent_multiplier, dtype=np.float64), (3,)))
        self.Nx, self.Ny, self.Nz, self.Ntotal = Nx, Ny, Nz, Nx * Ny * Nz + 8
        self.h = lib().mbo_gridmerge_create(Nx, Ny, Nz, _p(m))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mbo_gridmerge_free(self.h)
            self.h = None

    def index(self, extents, v):
        """compute_grid_index (merging_grid.jl:235) for explicit extents ((vx_lo, vx_hi), (vy_lo, vy_hi), (vz_lo, vz_hi))."""
        e, v = _f64(np.asarray(extents).reshape(6)), _f64(v)
        return lib().mbo_gridmerge_index(self.h, _p(e), _p(v))

    def cell(self, i):
        out = np.zeros(14)
        lib().mbo_gridmerge_cell(self.h, i, _p(out))
        return dict(np=int(out[0]), w=out[1], v_mean=out[2:5], v_std_sq=out[5:8], x_mean=out[8:11], x_std_sq=out[11:14])


def merge_grid_based(rng, mg, pv, pia, cell_lo, cell_hi, species, mass, T_v=None, extents=None, threshold=-1, grid=None):
    """merge_grid_based! (merging_grid.jl:597-703) over cells cell_lo..cell_hi: T_v = array [n_range, 4] of (T, vx, vy, vz) per cell (the
    PhysProps variant) or extents = ((vx_lo, vx_hi), (vy_lo, vy_hi), (vz_lo, vz_hi)); grid = (L, nx) for the 1-D variant."""
    tv = None if T_v is None else _f64(np.asarray(T_v).reshape(-1, 4))
    e = None if extents is None else _f64(np.asarray(extents).reshape(6))
    L, nx = grid if grid is not None else (0.0, 0)
    return lib().mbo_merge_grid_based(rng.ref, mg.h, pv.h, pia.h, cell_lo, cell_hi, species, mass, threshold, _p(tv), _p(e), L, nx)


class Exchanger:
    """ChunkExchanger (parallel.jl:21-46)."""

    def __init__(self, chunks, n_cells):
        self.chunks = [(int(a), int(b)) for a, b in chunks]
        self.n_cells = n_cells
        self.h = lib().mbo_exchanger_create(len(self.chunks), n_cells)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mbo_exchanger_free(self.h)
            self.h = None

    @property
    def indexer(self):  # [cell, chunk, 7]
        return np.ctypeslib.as_array(lib().mbo_exchanger_indexer(self.h), shape=(self.n_cells, len(self.chunks), 7))

    def reset(self, chunk_id):
        lib().mbo_exchanger_reset(self.h, chunk_id)

    def exchange(self, pvs, pias, species=1, i=0, j=0):
        lo = np.array([c[0] for c in self.chunks], dtype=np.int64)
        hi = np.array([c[1] for c in self.chunks], dtype=np.int64)
        ph = (C.c_void_p * len(pias))(*[p.h for p in pias])
        lib().mbo_exchange_particles(self.h, _handles(pvs), ph, _p(lo), _p(hi), len(self.chunks), species, i, j)

    def sort_after_exchange(self, pv, pia, chunk_id, species=1):
        lo, hi = self.chunks[chunk_id - 1]
        lib().mbo_sort_particles_after_exchange(self.h, pv.h, pia.h, lo, hi, species)


def generate_1_factorization(n_chunks):
    npairs = n_chunks * (n_chunks - 1) // 2
    pairs = np.zeros((max(npairs, 1), 2), dtype=np.int64)
    rnd = np.zeros(max(npairs, 1), dtype=np.int64)
    nr = lib().mbo_generate_1_factorization(n_chunks, _p(pairs), _p(rnd))
    return [[tuple(int(v) for v in pairs[k]) for k in range(npairs) if rnd[k] == r] for r in range(nr)]


def chunks(n, n_chunks):
    """ChunkSplitters.chunks(1:n; n=n_chunks): contiguous balanced ranges, the first n % n_chunks one longer."""
    base, rem = divmod(n, n_chunks)
    out, lo = [], 1
    for c in range(n_chunks):
        ln = base + (1 if c < rem else 0)
        out.append((lo, lo + ln - 1))
        lo += ln
    return out

"""merzbild_b200 -- host-side mirror of Merzbild.jl's per-timestep DSMC API over libmerzbild_b200.so (CUDA, sm_100a).

The reference is Julia; there is no Julia toolchain in the build image, so this ctypes module is the executable twin of the
Julia shim ``merzbild.jl_b200/julia/MerzbildB200.jl``: both bind exactly the symbols of ``include/merzbild_b200.h`` and keep
the reference's names, argument order and 1-based conventions (``/root/reference/src/Merzbild.jl:26-71`` export list):

    ParticleVector, ParticleIndexerArray, Grid1DUniform, GridSortInPlace, MaxwellWalls1D, PhysProps,
    sort_particles, ntc, ntc_equal_weight, swpm, fp_linear, compute_props, compute_props_sorted,
    compute_props_with_total_moments, convect_particles, convect_particles_and_compute_cell,
    merge_octree_N2_based, squash_pia, restore_particle_ordering, ...

Differences forced by the device (documented in DESIGN.md): the ``rng`` argument is a :class:`PhiloxRng`
(seed, timestep, substream) instead of a sequential generator, and every per-cell operator accepts either a cell or an
inclusive ``(cell_lo, cell_hi)`` range so that a ``for cell in 1:n_cells`` loop becomes one launch.

There is NO CPU fallback: if the shared library is missing, or no CUDA device is present, construction of a
:class:`Context` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MERZBILD_B200_LIB") or os.path.join(_HERE, "libmerzbild_b200.so")  # the override serves kernel-variant experiments

MB_OK, MB_ERR_NO_DEVICE, MB_ERR_CUDA, MB_ERR_ARG, MB_ERR_CAPACITY, MB_ERR_PRECONDITION, MB_ERR_NCCL, MB_ERR_UNSUPPORTED = range(8)
K_B = 1.380649e-23

# field order of ParticleIndexer (particles.jl:56-66)
N_LOCAL, START1, END1, N_GROUP1, START2, END2, N_GROUP2 = range(7)


class MerzbildError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libmerzbild_b200 status {status}: {msg}")
        self.status = status


class CapacityError(MerzbildError):
    pass


class Grid1D(C.Structure):
    _fields_ = [("L", C.c_double), ("n_cells", C.c_int64), ("dx", C.c_double), ("inv_dx", C.c_double), ("min_x", C.c_double),
                ("max_x", C.c_double), ("cell_offset", C.c_int64)]


class Walls1D(C.Structure):
    _fields_ = [("T", C.c_double * 2), ("v", (C.c_double * 3) * 2), ("accommodation", C.c_double * 2)]


class Interaction(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("m_r", "mu1", "mu2", "vhs_d", "vhs_o", "vhs_Tref", "vhs_muref", "vhs_factor")]


class GridMergeParams(C.Structure):
    _fields_ = [("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32), ("extent_multiplier", C.c_double * 3)]


class OctreeParams(C.Structure):
    _fields_ = [("split", C.c_int32), ("init_bin_bounds", C.c_int32), ("bin_bounds_compute", C.c_int32), ("max_depth", C.c_int32),
                ("max_Nbins", C.c_int64)]


_lib = None

# name -> (restype, argtypes): the complete export list of include/merzbild_b200.h
_vp, _i64, _i32, _u32, _u64, _f64, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_double, C.c_int
SIGNATURES = {
    "mb_last_error_string": (C.c_char_p, []),
    "mb_version": (_int, []),
    "mb_ctx_create": (_int, [_int, _u64, C.POINTER(_vp)]),
    "mb_ctx_destroy": (_int, [_vp]),
    "mb_sync": (_int, [_vp]),
    "mb_ctx_stream": (_vp, [_vp]),
    "mb_ctx_set_seed": (_int, [_vp, _u64]),
    "mb_ctx_kernel_launches": (_i64, [_vp]),
    "mb_timer_start": (_int, [_vp]),
    "mb_timer_stop": (_int, [_vp, C.POINTER(_f64)]),
    "mb_flush_l2": (_int, [_vp]),
    "mb_prof_enable": (_int, [_vp, _i32]),
    "mb_prof_read": (_int, [_vp, _i32, C.POINTER(_f64), C.POINTER(_i64)]),
    "mb_grid1d_init": (_int, [_f64, _i64, _f64, C.POINTER(Grid1D)]),
    "mb_grid1d_slab": (_int, [C.POINTER(Grid1D), _int, _int, C.POINTER(Grid1D)]),
    "mb_pv_create": (_int, [_vp, _i64, C.POINTER(_vp)]),
    "mb_pv_destroy": (_int, [_vp]),
    "mb_pv_length": (_i64, [_vp]),
    "mb_pv_resize": (_int, [_vp, _i64]),
    "mb_pv_upload_rows": (_int, [_vp, _i64, _i64, _vp]),
    "mb_pv_download_rows": (_int, [_vp, _i64, _i64, _vp]),
    "mb_pv_upload_soa": (_int, [_vp, _i64, _i64] + [_vp] * 7),
    "mb_pv_download_soa": (_int, [_vp, _i64, _i64] + [_vp] * 7),
    "mb_pv_upload_cell": (_int, [_vp, _i64, _i64, _vp]),
    "mb_pv_download_cell": (_int, [_vp, _i64, _i64, _vp]),
    "mb_pv_device_ptrs": (_int, [_vp, C.POINTER(_vp)]),
    "mb_pia_create": (_int, [_vp, _i64, _i64, C.POINTER(_vp)]),
    "mb_pia_destroy": (_int, [_vp]),
    "mb_pia_upload": (_int, [_vp, _vp, _vp, _vp]),
    "mb_pia_download": (_int, [_vp, _vp, _vp, _vp]),
    "mb_pia_n_total": (_i64, [_vp, _i64]),
    "mb_check_pia": (_int, [_vp, _i64, C.POINTER(_i32), C.POINTER(_i64)]),
    "mb_sort_particles": (_int, [_vp, C.POINTER(Grid1D), _vp, _vp, _i64]),
    "mb_sort_last_path": (_int, [_vp]),
    "mb_sort_set_band_halfwidth": (_int, [_vp, _i32]),
    "mb_sort_last_extras": (_i64, [_vp]),
    "mb_sort_last_pass_b": (_int, [_vp]),
    "mb_squash_pia": (_int, [_vp, _vp, _vp, _i64]),
    "mb_restore_particle_ordering": (_int, [_vp, _vp]),
    "mb_make_interaction": (_int, [_f64, _f64, _f64, _f64, _f64, C.POINTER(Interaction)]),
    "mb_estimate_sigma_g_w_max": (_f64, [C.POINTER(Interaction), _f64, _f64, _f64, _f64, _f64, _f64]),
    "mb_cf_create": (_int, [_vp, _i64, _f64, C.POINTER(_vp)]),
    "mb_cf_destroy": (_int, [_vp]),
    "mb_cf_fill": (_int, [_vp, _f64]),
    "mb_cf_upload": (_int, [_vp, _vp]),
    "mb_cf_download": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "mb_ntc": (_int, [_vp, _vp, C.POINTER(Interaction), _vp, _vp, _i64, _i64, _i64, _f64, _f64, _f64, _i32, _u32, _u32]),
    "mb_ntc2": (_int, [_vp, _vp, C.POINTER(Interaction), _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _i32, _u32, _u32]),
    "mb_swpm": (_int, [_vp, _vp, C.POINTER(Interaction), _vp, _vp, _i64, _i64, _i64, _f64, _f64, _f64, _u32, _u32]),
    "mb_fp_linear": (_int, [_vp, C.POINTER(Interaction), _f64, _vp, _vp, _i64, _i64, _i64, _f64, _f64, _u32, _u32]),
    "mb_convect_particles": (_int, [_vp, C.POINTER(Grid1D), C.POINTER(Walls1D), _vp, _vp, _i64, _f64, _vp, _f64, _i32, _u32, _u32]),
    "mb_props_create": (_int, [_vp, _i64, _i64, _i64, _vp, _f64, _i32, C.POINTER(_vp)]),
    "mb_props_destroy": (_int, [_vp]),
    "mb_props_download": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mb_props_clear": (_int, [_vp]),
    "mb_props_avg": (_int, [_vp, _vp, _i64]),
    "mb_compute_props": (_int, [_vp, C.POINTER(_vp), _vp, _vp, _vp, _i32]),
    "mb_compute_props_sorted": (_int, [_vp, C.POINTER(_vp), _vp, _vp, _vp, C.POINTER(Grid1D), _i64, _i64]),
    "mb_merge_octree_N2": (_int, [_vp, C.POINTER(OctreeParams), _vp, _vp, _i64, _i64, _i64, _i64, _i64, C.POINTER(Grid1D), _u32, _u32]),
    "mb_merge_grid_based": (_int, [_vp, C.POINTER(GridMergeParams), _vp, _vp, _i64, _i64, _i64, _f64, _vp, _vp, _i64, C.POINTER(Grid1D), _u32, _u32]),
    "mb_sample_particles_equal_weight": (_int, [_vp, C.POINTER(Grid1D), _vp, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _f64, _vp, _i32, _vp, _u32, _u32]),
    "mb_sample_on_grid": (_int, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _vp, _f64, _f64, _f64, _vp, _u32, _u32, C.POINTER(_i64)]),
    "mb_comm_unique_id": (_int, [_vp]),
    "mb_comm_init": (_int, [_vp, _vp, _int, _int]),
    "mb_exchange_set_mode": (_int, [_vp, _i32]),
    "mb_surf_create": (_int, [_vp, C.POINTER(_vp)]),
    "mb_surf_destroy": (_int, [_vp]),
    "mb_surf_clear": (_int, [_vp]),
    "mb_surf_upload": (_int, [_vp, _vp]),
    "mb_surf_download": (_int, [_vp, _vp]),
    "mb_surf_avg": (_int, [_vp, _vp, _i64]),
    "mb_surf_reduce": (_int, [_vp, _vp, _i32, _i32]),
    "mb_convect_particles_surf": (_int, [_vp, C.POINTER(Grid1D), C.POINTER(Walls1D), _vp, _vp, _i64, _f64, _vp, _f64, _i32, _u32, _u32]),
    "mb_exchange_slab": (_int, [_vp, C.POINTER(Grid1D), _vp, _vp, _i64, _vp, _vp]),
    "mb_exchange_chunks": (_int, [_i32, _vp, _vp, _vp, _vp, _i64]),
}


def lib():
    """Load libmerzbild_b200.so (built by __graft_entry__.build() / csrc/Makefile).  No fallback if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MerzbildError(-1, f"{LIB_PATH} not found -- build it with `make -C merzbild.jl_b200/csrc` "
                                    "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _ck(status):
    if status != MB_OK:
        msg = lib().mb_last_error_string().decode()
        raise (CapacityError if status == MB_ERR_CAPACITY else MerzbildError)(status, msg)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PhiloxRng:
    """Replaces the reference's ``rng::AbstractRNG`` first argument: Philox4x32-10 streams keyed per
    (operator, substream, timestep, entity); the seed lives in the :class:`Context`."""

    def __init__(self, timestep=0, substream=0):
        self.timestep, self.substream = int(timestep), int(substream)

    def at(self, timestep, substream=None):
        return PhiloxRng(timestep, self.substream if substream is None else substream)


class Context:
    """One per GPU: device, stream, Philox seed, scratch."""

    def __init__(self, device=0, seed=1234):
        h = C.c_void_p()
        _ck(lib().mb_ctx_create(int(device), int(seed), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().mb_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        _ck(lib().mb_sync(self.h))

    def set_seed(self, seed):
        _ck(lib().mb_ctx_set_seed(self.h, int(seed)))

    @property
    def kernel_launches(self):
        return int(lib().mb_ctx_kernel_launches(self.h))

    @property
    def stream(self):
        return lib().mb_ctx_stream(self.h)

    def timer_start(self):
        _ck(lib().mb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _ck(lib().mb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _ck(lib().mb_flush_l2(self.h))

    PROF_SECTIONS = ("sort.classify", "sort.scan", "sort.scatter", "sort.general", "ntc", "convect", "props", "merge", "fp", "exchange",
                     "squash", "sort.extras")

    def prof_enable(self, on=True):
        _ck(lib().mb_prof_enable(self.h, int(on)))

    def prof_read(self):
        """{section: (total_ms, launches)} since the last read."""
        out = {}
        for i, name in enumerate(self.PROF_SECTIONS):
            ms, n = C.c_double(), C.c_int64()
            _ck(lib().mb_prof_read(self.h, i, C.byref(ms), C.byref(n)))
            if n.value:
                out[name] = (ms.value, int(n.value))
        return out

    def set_band_halfwidth(self, w):
        _ck(lib().mb_sort_set_band_halfwidth(self.h, int(w)))

    @property
    def sort_last_path(self):
        return lib().mb_sort_last_path(self.h)

    @property
    def sort_last_extras(self):
        """band outliers + slab-exchange arrivals placed by the last band-path sort (-1: the general path ran)"""
        return int(lib().mb_sort_last_extras(self.h))

    @property
    def sort_last_pass_b(self):
        """pass B of the last band-path sort: 0 = warp per old cell, 1 = tile kernel (TMA bulk copies)"""
        return int(lib().mb_sort_last_pass_b(self.h))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0, 1234)
    return _default_ctx


class Grid1DUniform:
    """Grid1DUniform(L, nx; wall_offset=1e-12) (grids/grid_uniform1D.jl:49-86)."""

    def __init__(self, L, nx, wall_offset=1e-12):
        self.c = Grid1D()
        _ck(lib().mb_grid1d_init(float(L), int(nx), float(wall_offset), C.byref(self.c)))

    def slab(self, rank, nranks):
        g = Grid1DUniform.__new__(Grid1DUniform)
        g.c = Grid1D()
        _ck(lib().mb_grid1d_slab(C.byref(self.c), int(rank), int(nranks), C.byref(g.c)))
        return g

    L = property(lambda s: s.c.L)
    n_cells = property(lambda s: s.c.n_cells)
    dx = property(lambda s: s.c.dx)
    inv_dx = property(lambda s: s.c.inv_dx)
    min_x = property(lambda s: s.c.min_x)
    max_x = property(lambda s: s.c.max_x)
    cell_offset = property(lambda s: s.c.cell_offset)

    def cell_V(self, cell=1):
        return self.c.dx

    @property
    def ref(self):
        return C.byref(self.c)


class MaxwellWalls1D:
    """MaxwellWalls1D(species_data, T_l, T_r, vy_l, vy_r, accomodation_l, accomodation_r) (boundary_conditions.jl:29-53)."""

    def __init__(self, T_l, T_r, vy_l, vy_r, accommodation_l, accommodation_r):
        w = Walls1D()
        w.T[0], w.T[1] = T_l, T_r
        w.v[0][1], w.v[1][1] = vy_l, vy_r
        w.accommodation[0], w.accommodation[1] = accommodation_l, accommodation_r
        self.c = w

    @property
    def ref(self):
        return C.byref(self.c)


class ParticleVector:
    """ParticleVector(np) (particles.jl:194-212) as device-resident fp64 SoA."""

    def __init__(self, np_, ctx=None):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        _ck(lib().mb_pv_create(self.ctx.h, int(np_), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mb_pv_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().mb_pv_length(self.h))

    def resize(self, n):
        _ck(lib().mb_pv_resize(self.h, int(n)))

    def set_logical(self, lo, rows):
        """pv[lo + i] = Particle(rows[i]) with rows (w, vx, vy, vz, x, y, z)."""
        rows = _f64arr(rows).reshape(-1, 7)
        _ck(lib().mb_pv_upload_rows(self.h, int(lo), rows.shape[0], _p(rows)))

    def logical(self, lo=1, hi=None):
        hi = len(self) if hi is None else hi
        out = np.empty((max(hi - lo + 1, 0), 7))
        if hi >= lo:
            _ck(lib().mb_pv_download_rows(self.h, int(lo), out.shape[0], _p(out)))
        return out

    def upload_soa(self, lo, n, arrays):
        ptrs = [_p(a) if a is not None else None for a in arrays]
        _ck(lib().mb_pv_upload_soa(self.h, int(lo), int(n), *ptrs))

    def download_soa(self, lo, n, arrays):
        ptrs = [_p(a) if a is not None else None for a in arrays]
        _ck(lib().mb_pv_download_soa(self.h, int(lo), int(n), *ptrs))

    def __getitem__(self, i):
        return self.logical(i, i)[0]

    def __setitem__(self, i, row):
        self.set_logical(i, np.asarray(row, dtype=np.float64).reshape(1, 7))

    def set_cell(self, lo, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        _ck(lib().mb_pv_upload_cell(self.h, int(lo), len(cells), _p(cells)))

    def cell(self, lo=1, hi=None):
        hi = len(self) if hi is None else hi
        out = np.empty(max(hi - lo + 1, 0), dtype=np.int64)
        if hi >= lo:
            _ck(lib().mb_pv_download_cell(self.h, int(lo), len(out), _p(out)))
        return out


class ParticleIndexerArray:
    """ParticleIndexerArray(n_cells, n_species) (particles.jl:104-141); host views are downloaded on demand."""

    def __init__(self, n_cells, n_species=1, ctx=None):
        self.ctx = ctx or default_context()
        self.n_cells, self.n_species = int(n_cells), int(n_species)
        h = C.c_void_p()
        _ck(lib().mb_pia_create(self.ctx.h, self.n_cells, self.n_species, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mb_pia_destroy(self.h)
            self.h = None

    def upload(self, indexer=None, n_total=None, contiguous=None):
        ix = None if indexer is None else np.ascontiguousarray(indexer, dtype=np.int64).reshape(self.n_species, self.n_cells, 7)
        nt = None if n_total is None else np.ascontiguousarray(n_total, dtype=np.int64).reshape(self.n_species)
        ct = None if contiguous is None else np.ascontiguousarray(contiguous, dtype=np.uint8).reshape(self.n_species)
        _ck(lib().mb_pia_upload(self.h, _p(ix), _p(nt), _p(ct)))

    def download(self):
        ix = np.empty((self.n_species, self.n_cells, 7), dtype=np.int64)
        nt = np.empty(self.n_species, dtype=np.int64)
        ct = np.empty(self.n_species, dtype=np.uint8)
        _ck(lib().mb_pia_download(self.h, _p(ix), _p(nt), _p(ct)))
        return ix, nt, ct

    @property
    def indexer(self):
        return self.download()[0]

    @property
    def n_total(self):
        return self.download()[1]

    @property
    def contiguous(self):
        return self.download()[2]

    def check(self, species=1):
        ok, where = C.c_int32(), C.c_int64()
        _ck(lib().mb_check_pia(self.h, species, C.byref(ok), C.byref(where)))
        return bool(ok.value), int(where.value)


class GridSortInPlace:
    """GridSortInPlace(grid | n_cells, n_particles) (grid_sorting.jl:10-41): the scratch lives in the Context."""

    def __init__(self, grid_or_n_cells=None, n_particles=None):
        pass


def make_interaction(m_i, m_k, d, o, Tref):
    it = Interaction()
    _ck(lib().mb_make_interaction(m_i, m_k, d, o, Tref, C.byref(it)))
    return it


def estimate_sigma_g_w_max(it, m1, m2, T1, T2, Fnum, mult_factor=1.0):
    return lib().mb_estimate_sigma_g_w_max(C.byref(it), m1, m2, T1, T2, Fnum, mult_factor)


class CollisionFactors:
    """create_collision_factors_array for one species pair (collision_ntc.jl:46-155): per-cell sigma_g_w_max + counters."""

    def __init__(self, n_cells, sigma_g_w_max=0.0, ctx=None):
        self.ctx = ctx or default_context()
        self.n_cells = int(n_cells)
        h = C.c_void_p()
        _ck(lib().mb_cf_create(self.ctx.h, self.n_cells, float(sigma_g_w_max), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mb_cf_destroy(self.h)
            self.h = None

    def fill(self, v):
        _ck(lib().mb_cf_fill(self.h, float(v)))

    def upload(self, sgwm):
        a = _f64arr(sgwm)
        _ck(lib().mb_cf_upload(self.h, _p(a)))

    def download(self):
        s = np.empty(self.n_cells)
        a, b, c = (np.empty(self.n_cells, dtype=np.int64) for _ in range(3))
        _ck(lib().mb_cf_download(self.h, _p(s), _p(a), _p(b), _p(c)))
        return dict(sigma_g_w_max=s, n_coll=a, n_coll_performed=b, n_eq_w_coll_performed=c)


def _range(cell):
    if isinstance(cell, (tuple, list)):
        return int(cell[0]), int(cell[1])
    return int(cell), int(cell)


def sort_particles(gridsort, *args):
    """sort_particles!(gridsort, grid, pv, pia, species) / sort_particles!(gridsort, pv, pia, species) (grid_sorting.jl:58,128)."""
    if isinstance(args[0], Grid1DUniform):
        grid, pv, pia, species = args
        _ck(lib().mb_sort_particles(pv.ctx.h, grid.ref, pv.h, pia.h, int(species)))
    else:
        pv, pia, species = args
        _ck(lib().mb_sort_particles(pv.ctx.h, None, pv.h, pia.h, int(species)))


def squash_pia(pv, pia, species=1):
    _ck(lib().mb_squash_pia(pv.ctx.h, pv.h, pia.h, int(species)))


def restore_particle_ordering(pv, inv_map=None):
    _ck(lib().mb_restore_particle_ordering(pv.ctx.h, pv.h))


def ntc(rng, cf, cd, interaction, pv, pia, cell, species, dt, V, dw_tol=1e-16, equal_weight=False):
    """ntc!(rng, collision_factors, collision_data, interaction, particles, pia, cell, species, Δt, V) (collision_ntc.jl:338)."""
    lo, hi = _range(cell)
    _ck(lib().mb_ntc(pv.ctx.h, cf.h, C.byref(interaction), pv.h, pia.h, lo, hi, int(species), dt, V, dw_tol, int(equal_weight), rng.timestep,
                     rng.substream))


def ntc_equal_weight(rng, cf, cd, interaction, pv, pia, cell, species, dt, V):
    """ntc_equal_weight! (collision_ntc.jl:479)."""
    ntc(rng, cf, cd, interaction, pv, pia, cell, species, dt, V, equal_weight=True)


def ntc2(rng, cf, cd, interaction, pv1, pv2, pia, cell, s1, s2, dt, V, dw_tol=1e-16, equal_weight=False):
    """two-species ntc! (collision_ntc.jl:412) / ntc_equal_weight! (:554)."""
    lo, hi = _range(cell)
    _ck(lib().mb_ntc2(pv1.ctx.h, cf.h, C.byref(interaction), pv1.h, pv2.h, pia.h, lo, hi, int(s1), int(s2), dt, V, dw_tol, int(equal_weight),
                      rng.timestep, rng.substream))


def swpm(rng, cf, cd, interaction, pv, pia, cell, species, G, dt, V):
    """swpm! (collision_swpm.jl:201)."""
    lo, hi = _range(cell)
    _ck(lib().mb_swpm(pv.ctx.h, cf.h, C.byref(interaction), pv.h, pia.h, lo, hi, int(species), G, dt, V, rng.timestep, rng.substream))


def fp_linear(rng, cd_fp, interaction, mass, pv, pia, cell, species, dt, V):
    """fp_linear! (collision_fp.jl:24)."""
    lo, hi = _range(cell)
    _ck(lib().mb_fp_linear(pv.ctx.h, C.byref(interaction), mass, pv.h, pia.h, lo, hi, int(species), dt, V, rng.timestep, rng.substream))


class SurfProps:
    """SurfProps(pia, grid) (surface_props.jl:22-50) for the two walls of a 1-D grid and one species, device resident: rows = walls,
    columns = np, flux_incident, flux_reflected, force[3], normal_pressure, shear_pressure[3], kinetic_energy_flux."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        _ck(lib().mb_surf_create(self.ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mb_surf_destroy(self.h)
            self.h = None

    def clear(self):
        _ck(lib().mb_surf_clear(self.h))

    def upload(self, rows):
        _ck(lib().mb_surf_upload(self.h, _p(_f64arr(rows).reshape(2, 11))))

    def download(self):
        out = np.empty((2, 11))
        _ck(lib().mb_surf_download(self.h, _p(out)))
        return out


def avg_surf_props(avg, surf, n_avg_timesteps):
    """avg_props!(surf_props_avg, surf_props, n_avg_timesteps) (surface_props.jl:202-222)"""
    _ck(lib().mb_surf_avg(avg.h, surf.h, int(n_avg_timesteps)))


def reduce_surf_props(target, chunks, across_ranks=False):
    """reduce_surf_props!(surf_props_target, surf_props_chunks) (surface_props.jl:232-252); across_ranks: also summed over the ranks of
    the target context's communicator"""
    hs = (C.c_void_p * max(len(chunks), 1))(*[c.h for c in chunks])
    _ck(lib().mb_surf_reduce(target.h, hs, len(chunks), int(across_ranks)))


def convect_particles(rng, grid, boundaries, pv, pia, species, mass, dt, surf_props=False, compute_cell=False):
    """convect_particles!(rng, grid, boundaries, particles, pia, species, species_data, [surf_props,] Δt) (convection_1D.jl:130,176);
    returns the 2x11 SurfProps rows (np, flux_incident, flux_reflected, force[3], normal_pressure, shear_pressure[3],
    kinetic_energy_flux) if surf_props."""
    if isinstance(surf_props, SurfProps):  # device-resident SurfProps: no host synchronisation
        _ck(lib().mb_convect_particles_surf(pv.ctx.h, grid.ref, boundaries.ref, pv.h, pia.h, int(species), float(mass), surf_props.h, dt,
                                            int(compute_cell), rng.timestep, rng.substream))
        return surf_props
    s = np.zeros((2, 11)) if surf_props else None
    _ck(lib().mb_convect_particles(pv.ctx.h, grid.ref, boundaries.ref, pv.h, pia.h, int(species), float(mass), _p(s), dt, int(compute_cell),
                                   rng.timestep, rng.substream))
    return s


def convect_particles_and_compute_cell(rng, grid, boundaries, pv, pia, species, mass, dt, surf_props=False):
    """convect_particles_and_compute_cell! (convection_1D.jl:225,274)."""
    return convect_particles(rng, grid, boundaries, pv, pia, species, mass, dt, surf_props, True)


class PhysProps:
    """PhysProps(n_cells, n_species, moment_powers; Tref=300) (physical_props.jl:24-71), device resident."""

    def __init__(self, n_cells, n_species=1, moment_powers=(), Tref=300.0, ndens_not_Np=False, ctx=None):
        self.ctx = ctx or default_context()
        self.n_cells, self.n_species = int(n_cells), int(n_species)
        self.moment_powers = np.asarray(moment_powers, dtype=np.int32)
        h = C.c_void_p()
        _ck(lib().mb_props_create(self.ctx.h, self.n_cells, self.n_species, len(self.moment_powers), _p(self.moment_powers), float(Tref),
                                  int(ndens_not_Np), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mb_props_destroy(self.h)
            self.h = None

    def download(self):
        ns, nc, nm = self.n_species, self.n_cells, len(self.moment_powers)
        out = dict(lpa=np.empty(ns), np=np.empty((ns, nc)), n=np.empty((ns, nc)), v=np.empty((ns, nc, 3)), T=np.empty((ns, nc)),
                   moments=np.empty((ns, nc, nm)) if nm else None)
        _ck(lib().mb_props_download(self.h, _p(out["lpa"]), _p(out["np"]), _p(out["n"]), _p(out["v"]), _p(out["T"]), _p(out["moments"])))
        return out

    def clear(self):
        _ck(lib().mb_props_clear(self.h))


def _handles(pvs):
    return (C.c_void_p * len(pvs))(*[p.h for p in pvs])


def compute_props(particles, pia, masses, phys_props, with_moments=False):
    """compute_props!(particles, pia, species_data, phys_props) (physical_props.jl:104)."""
    m = _f64arr(masses)
    _ck(lib().mb_compute_props(particles[0].ctx.h, _handles(particles), pia.h, _p(m), phys_props.h, int(with_moments)))


def compute_props_with_total_moments(particles, pia, masses, phys_props):
    """compute_props_with_total_moments! (physical_props.jl:168)."""
    compute_props(particles, pia, masses, phys_props, True)


def compute_props_sorted(particles, pia, masses, phys_props, grid=None, cell_chunk=None):
    """compute_props_sorted!(particles, pia, species_data, phys_props[, grid][, cell_chunk]) (physical_props.jl:317-454)."""
    lo, hi = (1, pia.n_cells) if cell_chunk is None else _range(cell_chunk)
    m = _f64arr(masses)
    _ck(lib().mb_compute_props_sorted(particles[0].ctx.h, _handles(particles), pia.h, _p(m), phys_props.h, grid.ref if grid is not None else None,
                                      lo, hi))


def avg_props(avg, props, n_avg_timesteps):
    """avg_props! (physical_props.jl:281)."""
    _ck(lib().mb_props_avg(avg.h, props.h, int(n_avg_timesteps)))


class OctreeN2Merge:
    """OctreeN2Merge(split; init_bin_bounds, bin_bounds_compute, max_Nbins=4096, max_depth=10) (merging_octree_N2.jl:235-246)."""
    OctreeBinMidSplit, OctreeBinMeanSplit = 1, 2
    OctreeInitBinMinMaxVel, OctreeInitBinMinMaxVelSym, OctreeInitBinC = 1, 2, 3
    OctreeBinBoundsInherit, OctreeBinBoundsRecompute = 1, 2

    def __init__(self, split=1, init_bin_bounds=1, bin_bounds_compute=1, max_Nbins=4096, max_depth=10):
        self.c = OctreeParams(int(split), int(init_bin_bounds), int(bin_bounds_compute), int(max_depth), int(max_Nbins))


def merge_octree_N2_based(rng, octree, pv, pia, cell, species, target_np, grid=None, threshold=-1):
    """merge_octree_N2_based!(rng, octree, particles, pia, cell, species, target_np[, grid]) (merging_octree_N2.jl:1060,1088);
    with a cell range, only cells with n_local > threshold are merged (threshold < 0: all)."""
    lo, hi = _range(cell)
    _ck(lib().mb_merge_octree_N2(pv.ctx.h, C.byref(octree.c), pv.h, pia.h, lo, hi, int(species), int(threshold), int(target_np),
                                 grid.ref if grid is not None else None, rng.timestep, rng.substream))


class GridN2Merge:
    """GridN2Merge(Nx, Ny, Nz, extent_multiplier) (merging_grid.jl:72-116); extent_multiplier a scalar or 3 values."""

    def __init__(self, Nx, Ny=None, Nz=None, extent_multiplier=3.5):
        Ny = Nx if Ny is None else Ny
        Nz = Nx if Nz is None else Nz
        m = np.broadcast_to(np.asarray(extent_multiplier, dtype=np.float64), (3,))
        self.c = GridMergeParams(int(Nx), int(Ny), int(Nz), (C.c_double * 3)(*m))


def merge_grid_based(rng, merging_grid, pv, pia, cell, species, mass, phys_props_or_vx_extent, vy_extent=None, vz_extent=None, grid=None, threshold=-1):
    """merge_grid_based!(rng, merging_grid, particles, pia, cell, species, species_data, phys_props[, grid]) or
    merge_grid_based!(..., species_data, vx_extent, vy_extent, vz_extent[, grid]) (merging_grid.jl:597-703); ``cell`` may be a range."""
    lo, hi = _range(cell)
    if isinstance(phys_props_or_vx_extent, PhysProps):
        props, ext = phys_props_or_vx_extent.h, None
    else:
        props, ext = None, _f64arr([*phys_props_or_vx_extent, *vy_extent, *vz_extent])
    _ck(lib().mb_merge_grid_based(pv.ctx.h, C.byref(merging_grid.c), pv.h, pia.h, lo, hi, int(species), float(mass), props, _p(ext), int(threshold),
                                  grid.ref if grid is not None else None, rng.timestep, rng.substream))


def sample_particles_equal_weight(rng, *args, distribution="Maxwellian", vx0=0.0, vy0=0.0, vz0=0.0, cell_chunk=None):
    """sample_particles_equal_weight!(rng, particles, pia, cell, species, nparticles, m, T, Fnum, xlo, xhi, ylo, yhi, zlo, zhi; ...)
    (distributions_and_sampling.jl:477) -- ``cell`` may be a range -- or
    sample_particles_equal_weight!(rng, grid1duniform, particles, pia, species, mass, ppc::Integer | ndens::Float64, T, Fnum[, cell_chunk])
    (grids/grid_uniform1D.jl:117-219)."""
    v0 = _f64arr([vx0, vy0, vz0])
    dist = {"Maxwellian": 0, "BKW": 1}[distribution]
    if isinstance(args[0], Grid1DUniform):
        grid, pv, pia, species, mass, ppc_or_ndens, T, Fnum = args[:8]
        if len(args) > 8:
            cell_chunk = args[8]
        lo, hi = (1, pia.n_cells) if cell_chunk is None else _range(cell_chunk)
        if isinstance(ppc_or_ndens, (int, np.integer)):
            npart, ndens = int(ppc_or_ndens), 0.0
        else:
            npart, ndens = -1, float(ppc_or_ndens)
        _ck(lib().mb_sample_particles_equal_weight(pv.ctx.h, grid.ref, pv.h, pia.h, lo, hi, int(species), npart, ndens, float(mass), float(T),
                                                   float(Fnum), None, dist, _p(v0), rng.timestep, rng.substream))
    else:
        pv, pia, cell, species, nparticles, m, T, Fnum = args[:8]
        box = _f64arr(args[8:14])
        lo, hi = _range(cell)
        _ck(lib().mb_sample_particles_equal_weight(pv.ctx.h, None, pv.h, pia.h, lo, hi, int(species), int(nparticles), 0.0, float(m), float(T),
                                                   float(Fnum), _p(box), dist, _p(v0), rng.timestep, rng.substream))


def sample_on_grid(rng, vdf, pv, pia, cell, species, nv, m, T, n_total, xlo=0.0, xhi=1.0, ylo=0.0, yhi=1.0, zlo=0.0, zhi=1.0, v_mult=3.5,
                   cutoff_mult=3.5, noise=0.0, v_offset=(0.0, 0.0, 0.0)):
    """sample_on_grid!(rng, vdf_func, particles, nv, m, T, n_total, xlo, ..., zhi; v_mult, cutoff_mult, noise, v_offset)
    (distributions_and_sampling.jl:312) for every cell of ``cell`` (an ensemble of 0-D cells); vdf is "maxwellian" or "bkw".
    Returns n_sampled (particles per cell) and sets the cells' indexers."""
    box = _f64arr([xlo, xhi, ylo, yhi, zlo, zhi])
    vo = _f64arr(v_offset)
    lo, hi = _range(cell)
    n = C.c_int64()
    _ck(lib().mb_sample_on_grid(pv.ctx.h, {"maxwellian": 0, "bkw": 1}[vdf], pv.h, pia.h, lo, hi, int(species), int(nv), float(m), float(T),
                                float(n_total), _p(box), v_mult, cutoff_mult, noise, _p(vo), rng.timestep, rng.substream, C.byref(n)))
    return int(n.value)


def comm_unique_id():
    buf = (C.c_char * 128)()
    _ck(lib().mb_comm_unique_id(buf))
    return bytes(buf)


def comm_init(ctx, unique_id, rank, nranks):
    buf = C.create_string_buffer(unique_id, 128)
    _ck(lib().mb_comm_init(ctx.h, buf, int(rank), int(nranks)))


def exchange_set_mode(ctx, mode):
    """0: edge exchange when the layout allows it (no host synchronisation), 1: always the full exchange."""
    _ck(lib().mb_exchange_set_mode(ctx.h, int(mode)))


def exchange_particles(ctxs, slabs, pv_chunks, pia_chunks, species=1):
    """exchange_particles!(exchanger, pv_chunks, pia_chunks, cell_chunks, species) (parallel.jl:443-450) for logical chunks that live in
    this process, one context per chunk (slabs[i] = grid.slab(i, n_chunks)); sort_particles! of every chunk then plays the role of
    sort_particles_after_exchange! (parallel.jl:467-532)."""
    n = len(ctxs)
    assert len(slabs) == len(pv_chunks) == len(pia_chunks) == n
    hc = (C.c_void_p * n)(*[c.h for c in ctxs])
    hp = (C.c_void_p * n)(*[p.h for p in pv_chunks])
    hi = (C.c_void_p * n)(*[p.h for p in pia_chunks])
    gs = (Grid1D * n)()
    for i, g in enumerate(slabs):
        C.memmove(C.byref(gs[i]), C.byref(g.c), C.sizeof(Grid1D))
    _ck(lib().mb_exchange_chunks(n, hc, gs, hp, hi, int(species)))


def exchange_slab(ctx, slab, pv, pia, species=1, counts=False):
    """Slab replacement of exchange_particles! + sort_particles_after_exchange! (parallel.jl:281-532)."""
    if counts:
        s, r = np.zeros(2, dtype=np.int64), np.zeros(2, dtype=np.int64)
        _ck(lib().mb_exchange_slab(ctx.h, slab.ref, pv.h, pia.h, int(species), _p(s), _p(r)))
        return s, r
    _ck(lib().mb_exchange_slab(ctx.h, slab.ref, pv.h, pia.h, int(species), None, None))

# MerzbildB200.jl -- Julia host shim over libmerzbild_b200.so (hand-written sm_100a CUDA kernels behind a C ABI).
#
# The host stays Julia: this module re-defines the per-timestep operators of Merzbild.jl on device-backed containers and
# forwards each of them with one `ccall` to the entry point declared in include/merzbild_b200.h.  Names, argument order,
# 1-based cells/species and the in-place `!` convention are the reference's (src/Merzbild.jl:26-71 export list); the
# reference structs that are plain data (Species, Interaction, Grid1DUniform, MaxwellWalls1D, OctreeN2Merge, PhysProps)
# are accepted as they are and converted to the C structs below.
#
# Differences forced by the device (see DESIGN.md "Boundary"):
#   * `ParticleVector` / `ParticleIndexerArray` become `DeviceParticleVector` / `DeviceParticleIndexerArray` (opaque handles);
#     `pv[i]`, `pv[i] = p`, `length`, `resize!`, `pia.indexer[c, s]`-style reads go through explicit H2D/D2H accessors.
#   * the `rng` argument is a `PhiloxRng(timestep, substream)`; the seed lives in the `Context`.  An `AbstractRNG` is also
#     accepted (and ignored) where the reference takes one, together with the keyword `timestep`.
#   * every per-cell operator accepts either `cell::Integer` (the reference's call) or `cells::UnitRange` (one launch for all
#     cells -- the production path: `for cell in 1:n_cells; ntc!(...) end` collapses to `ntc!(..., 1:n_cells, ...)`).
#   * there is NO CPU fallback: if the library or a CUDA device is missing, `Context()` throws.
#
# This file cannot be executed in the build image (no Julia toolchain there).  It is kept in lock-step with the Python
# ctypes mirror merzbild.jl_b200/merzbild_b200/__init__.py, which binds the same symbols with the same argument lists and
# is what the test-suite drives; tests/test_abi.py::test_julia_shim_binds_the_header checks that every symbol of the header
# is bound here with the right arity.
module MerzbildB200

export Context, PhiloxRng, DeviceParticleVector, DeviceParticleIndexerArray, DeviceCollisionFactors, DevicePhysProps, DeviceSurfProps,
       reduce_surf_props!,
       DeviceGrid1D, slab, sort_particles!, squash_pia!, restore_particle_ordering!, ntc!, ntc_equal_weight!, swpm!, fp_linear!,
       convect_particles!, convect_particles_and_compute_cell!, compute_props!, compute_props_sorted!,
       compute_props_with_total_moments!, avg_props!, clear_props!, merge_octree_N2_based!, exchange_particles!,
       sample_particles_equal_weight!, sample_on_grid!, merge_grid_based!,
       comm_unique_id, comm_init!, upload!, download, download_indexer, n_total, synchronize, kernel_launches

const libmb = get(ENV, "MERZBILD_B200_LIB", joinpath(@__DIR__, "..", "merzbild_b200", "libmerzbild_b200.so"))

# ---------------------------------------------------------------------------------------------------------------- errors
const MB_OK = Cint(0)
const MB_ERR_CAPACITY = Cint(4)

struct MerzbildB200Error <: Exception
    status::Cint
    msg::String
end
Base.showerror(io::IO, e::MerzbildB200Error) = print(io, "libmerzbild_b200 status ", e.status, ": ", e.msg)

last_error() = unsafe_string(ccall((:mb_last_error_string, libmb), Cstring, ()))
@inline function check(status::Cint)
    status == MB_OK || throw(MerzbildB200Error(status, last_error()))
    nothing
end
version() = ccall((:mb_version, libmb), Cint, ())

# ------------------------------------------------------------------------------------------------- C structs (isbits)
# mb_grid1d  <->  Grid1DUniform (grids/grid_uniform1D.jl:49-86)
struct CGrid1D
    L::Cdouble
    n_cells::Int64
    dx::Cdouble
    inv_dx::Cdouble
    min_x::Cdouble
    max_x::Cdouble
    cell_offset::Int64
end
# mb_walls1d  <->  MaxwellWalls1D (convection/boundary_conditions.jl:29-53)
struct CWalls1D
    T::NTuple{2,Cdouble}
    v::NTuple{6,Cdouble}            # v[wall][component], row-major
    accommodation::NTuple{2,Cdouble}
end
# mb_interaction  <->  Interaction (collisions/collision_utils.jl:73-82)
struct CInteraction
    m_r::Cdouble
    mu1::Cdouble
    mu2::Cdouble
    vhs_d::Cdouble
    vhs_o::Cdouble
    vhs_Tref::Cdouble
    vhs_muref::Cdouble
    vhs_factor::Cdouble
end
# mb_octree_params  <->  OctreeN2Merge (merging/merging_octree_N2.jl:131-179)
struct COctreeParams
    split::Int32
    init_bin_bounds::Int32
    bin_bounds_compute::Int32
    max_depth::Int32
    max_Nbins::Int64
end

# conversions from the reference's structs (duck-typed on field names so that this file loads without Merzbild.jl too)
CInteraction(it) = CInteraction(it.m_r, it.μ1, it.μ2, it.vhs_d, it.vhs_o, it.vhs_Tref, it.vhs_muref, it.vhs_factor)
function CWalls1D(w)
    b = w.boundaries
    CWalls1D((b[1].T, b[2].T), (b[1].v[1], b[1].v[2], b[1].v[3], b[2].v[1], b[2].v[2], b[2].v[3]),
             (b[1].accommodation, b[2].accommodation))
end
# enum order of the reference: OctreeBinMidSplit=1, OctreeBinMeanSplit=2 (:12); OctreeInitBinMinMaxVel=1,
# OctreeInitBinMinMaxVelSym=2, OctreeInitBinC=3 (:24); OctreeBinBoundsInherit=1, OctreeBinBoundsRecompute=2 (:34)
COctreeParams(o) = COctreeParams(Int32(Integer(o.split)), Int32(Integer(o.init_bin_bounds)), Int32(Integer(o.bin_bounds_compute)),
                                 Int32(o.max_depth), Int64(o.max_Nbins))

# ---------------------------------------------------------------------------------------------------------- context
"""
    Context(device=0; seed=1234)

One per GPU: device, stream, Philox seed, scratch arena and (optionally) the NCCL communicator.
Single caller thread per context, exactly like the reference's per-chunk ownership rule (docs/src/multithreaded.md:10-17).
"""
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0; seed::Integer=1234)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb_ctx_create, libmb), Cint, (Cint, UInt64, Ptr{Ptr{Cvoid}}), device, seed, out))
        ctx = new(out[])
        finalizer(c -> (c.h != C_NULL && ccall((:mb_ctx_destroy, libmb), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        ctx
    end
end
synchronize(ctx::Context) = check(ccall((:mb_sync, libmb), Cint, (Ptr{Cvoid},), ctx.h))
stream(ctx::Context) = ccall((:mb_ctx_stream, libmb), Ptr{Cvoid}, (Ptr{Cvoid},), ctx.h)
set_seed!(ctx::Context, seed::Integer) = check(ccall((:mb_ctx_set_seed, libmb), Cint, (Ptr{Cvoid}, UInt64), ctx.h, seed))
kernel_launches(ctx::Context) = ccall((:mb_ctx_kernel_launches, libmb), Int64, (Ptr{Cvoid},), ctx.h)
timer_start!(ctx::Context) = check(ccall((:mb_timer_start, libmb), Cint, (Ptr{Cvoid},), ctx.h))
function timer_stop!(ctx::Context)
    ms = Ref{Cdouble}(0.0)
    check(ccall((:mb_timer_stop, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.h, ms))
    ms[]
end
flush_l2!(ctx::Context) = check(ccall((:mb_flush_l2, libmb), Cint, (Ptr{Cvoid},), ctx.h))
prof_enable!(ctx::Context, on::Bool=true) = check(ccall((:mb_prof_enable, libmb), Cint, (Ptr{Cvoid}, Int32), ctx.h, on))
function prof_read(ctx::Context, section::Integer)
    ms = Ref{Cdouble}(0.0); n = Ref{Int64}(0)
    check(ccall((:mb_prof_read, libmb), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Int64}), ctx.h, section, ms, n))
    (ms[], n[])
end
sort_last_path(ctx::Context) = ccall((:mb_sort_last_path, libmb), Cint, (Ptr{Cvoid},), ctx.h)
sort_last_extras(ctx::Context) = ccall((:mb_sort_last_extras, libmb), Int64, (Ptr{Cvoid},), ctx.h)
sort_last_pass_b(ctx::Context) = ccall((:mb_sort_last_pass_b, libmb), Cint, (Ptr{Cvoid},), ctx.h)
set_band_halfwidth!(ctx::Context, w::Integer) = check(ccall((:mb_sort_set_band_halfwidth, libmb), Cint, (Ptr{Cvoid}, Int32), ctx.h, w))

"""
    PhiloxRng(timestep, substream=0)

Replaces the reference's `rng::AbstractRNG`: Philox4x32-10 streams keyed per (operator, substream, timestep, entity).
"""
struct PhiloxRng
    timestep::UInt32
    substream::UInt32
end
PhiloxRng(t::Integer) = PhiloxRng(UInt32(t), UInt32(0))

# ------------------------------------------------------------------------------------------------------------- grid
"""
    DeviceGrid1D(L, nx; wall_offset=1e-12)  |  DeviceGrid1D(grid::Grid1DUniform)

Grid1DUniform (grids/grid_uniform1D.jl:72-86) plus the slab offset used by the multi-GPU partition.
"""
struct DeviceGrid1D
    c::CGrid1D
end
function DeviceGrid1D(L::Real, nx::Integer; wall_offset::Real=1e-12)
    out = Ref{CGrid1D}()
    check(ccall((:mb_grid1d_init, libmb), Cint, (Cdouble, Int64, Cdouble, Ptr{CGrid1D}), L, nx, wall_offset, out))
    DeviceGrid1D(out[])
end
DeviceGrid1D(g) = DeviceGrid1D(CGrid1D(g.L, g.n_cells, g.Δx, g.inv_Δx, g.min_x, g.max_x, 0))
"""
    slab(grid, rank, nranks)

Contiguous balanced slab of cells for 0-based `rank` (the ChunkSplitters.chunks rule of couette_multithreaded.jl:30-31).
"""
function slab(g::DeviceGrid1D, rank::Integer, nranks::Integer)
    out = Ref{CGrid1D}()
    check(ccall((:mb_grid1d_slab, libmb), Cint, (Ptr{CGrid1D}, Cint, Cint, Ptr{CGrid1D}), Ref(g.c), rank, nranks, out))
    DeviceGrid1D(out[])
end
gridref(g::DeviceGrid1D) = Ref(g.c)
gridref(g) = Ref(DeviceGrid1D(g).c)

# ------------------------------------------------------------------------------------------------- ParticleVector
"""
    DeviceParticleVector(ctx, np)

ParticleVector(np) (particles.jl:210-212) as device-resident fp64 SoA (w, vx, vy, vz, x, y, z).  The reference's `index`
indirection is the identity on the device (the sort physically reorders), so `pv[i]` is element `i` of every array.
"""
mutable struct DeviceParticleVector
    ctx::Context
    h::Ptr{Cvoid}
    function DeviceParticleVector(ctx::Context, np::Integer)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb_pv_create, libmb), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), ctx.h, np, out))
        pv = new(ctx, out[])
        finalizer(p -> (p.h != C_NULL && ccall((:mb_pv_destroy, libmb), Cint, (Ptr{Cvoid},), p.h); p.h = C_NULL), pv)
        pv
    end
end
Base.length(pv::DeviceParticleVector) = Int(ccall((:mb_pv_length, libmb), Int64, (Ptr{Cvoid},), pv.h))
Base.resize!(pv::DeviceParticleVector, n::Integer) = (check(ccall((:mb_pv_resize, libmb), Cint, (Ptr{Cvoid}, Int64), pv.h, n)); pv)

"""
    upload!(pv, lo, rows::Matrix{Float64})     # rows is 7 x n: (w, vx, vy, vz, x, y, z) per column == pv[lo + j - 1]
"""
function upload!(pv::DeviceParticleVector, lo::Integer, rows::Matrix{Float64})
    size(rows, 1) == 7 || throw(ArgumentError("rows must be 7 x n"))
    check(ccall((:mb_pv_upload_rows, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), pv.h, lo, size(rows, 2), rows))
end
function download(pv::DeviceParticleVector, lo::Integer, n::Integer)
    rows = Matrix{Float64}(undef, 7, n)
    check(ccall((:mb_pv_download_rows, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), pv.h, lo, n, rows))
    rows
end
"SoA upload of 7 host vectors (may be pinned); `nothing` skips a field."
function upload_soa!(pv::DeviceParticleVector, lo::Integer, n::Integer, w, vx, vy, vz, x, y, z)
    p(a) = a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a)
    GC.@preserve w vx vy vz x y z check(ccall((:mb_pv_upload_soa, libmb), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        pv.h, lo, n, p(w), p(vx), p(vy), p(vz), p(x), p(y), p(z)))
end
function download_soa!(pv::DeviceParticleVector, lo::Integer, n::Integer, w, vx, vy, vz, x, y, z)
    p(a) = a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a)
    GC.@preserve w vx vy vz x y z check(ccall((:mb_pv_download_soa, libmb), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        pv.h, lo, n, p(w), p(vx), p(vy), p(vz), p(x), p(y), p(z)))
end
# pv[i] -> (w, v, x) like the reference's Particle (particles.jl:14-18, getindex :225)
function Base.getindex(pv::DeviceParticleVector, i::Integer)
    r = download(pv, i, 1)
    (w = r[1, 1], v = (r[2, 1], r[3, 1], r[4, 1]), x = (r[5, 1], r[6, 1], r[7, 1]))
end
function Base.setindex!(pv::DeviceParticleVector, p, i::Integer)   # p has fields w, v, x (a Merzbild.Particle works)
    upload!(pv, i, reshape(Float64[p.w, p.v[1], p.v[2], p.v[3], p.x[1], p.x[2], p.x[3]], 7, 1))
end
"pv.cell (particles.jl:197) for logical positions lo .. lo+n-1"
function upload_cell!(pv::DeviceParticleVector, lo::Integer, cell::Vector{Int64})
    check(ccall((:mb_pv_upload_cell, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), pv.h, lo, length(cell), cell))
end
function download_cell(pv::DeviceParticleVector, lo::Integer, n::Integer)
    cell = Vector{Int64}(undef, n)
    check(ccall((:mb_pv_download_cell, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), pv.h, lo, n, cell))
    cell
end
"raw device pointers of the 7 SoA arrays (valid until the next sort / resize) for zero-copy interop with CUDA.jl"
function device_ptrs(pv::DeviceParticleVector)
    out = Vector{Ptr{Cvoid}}(undef, 7)
    check(ccall((:mb_pv_device_ptrs, libmb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), pv.h, out))
    out
end

# ------------------------------------------------------------------------------------------ ParticleIndexerArray
"""
    DeviceParticleIndexerArray(ctx, n_cells, n_species)

ParticleIndexerArray(n_cells, n_species) (particles.jl:131-141).  `download_indexer(pia)` returns an `Array{Int64,3}` of size
(7, n_cells, n_species) whose slice `[:, c, s]` is (n_local, start1, end1, n_group1, start2, end2, n_group2) of
`pia.indexer[c, s]` (particles.jl:56-66).
"""
mutable struct DeviceParticleIndexerArray
    ctx::Context
    h::Ptr{Cvoid}
    n_cells::Int64
    n_species::Int64
    function DeviceParticleIndexerArray(ctx::Context, n_cells::Integer, n_species::Integer)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb_pia_create, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Ptr{Cvoid}}), ctx.h, n_cells, n_species, out))
        pia = new(ctx, out[], n_cells, n_species)
        finalizer(p -> (p.h != C_NULL && ccall((:mb_pia_destroy, libmb), Cint, (Ptr{Cvoid},), p.h); p.h = C_NULL), pia)
        pia
    end
end
function upload!(pia::DeviceParticleIndexerArray, indexer::Union{Nothing,Array{Int64,3}}, n_total::Union{Nothing,Vector{Int64}},
                 contiguous::Union{Nothing,Vector{UInt8}})
    p(a, T) = a === nothing ? Ptr{T}(C_NULL) : pointer(a)
    GC.@preserve indexer n_total contiguous check(ccall((:mb_pia_upload, libmb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}),
        pia.h, p(indexer, Int64), p(n_total, Int64), p(contiguous, UInt8)))
end
"upload a reference `ParticleIndexerArray` (pia.indexer[c, s] structs -> the 7-int rows)"
function upload!(pia::DeviceParticleIndexerArray, ref)
    ix = Array{Int64,3}(undef, 7, pia.n_cells, pia.n_species)
    for s in 1:pia.n_species, c in 1:pia.n_cells
        q = ref.indexer[c, s]
        ix[:, c, s] .= (q.n_local, q.start1, q.end1, q.n_group1, q.start2, q.end2, q.n_group2)
    end
    upload!(pia, ix, Vector{Int64}(ref.n_total), UInt8.(ref.contiguous))
end
function download_indexer(pia::DeviceParticleIndexerArray)
    ix = Array{Int64,3}(undef, 7, pia.n_cells, pia.n_species)
    nt = Vector{Int64}(undef, pia.n_species)
    ct = Vector{UInt8}(undef, pia.n_species)
    check(ccall((:mb_pia_download, libmb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}), pia.h, ix, nt, ct))
    (indexer = ix, n_total = nt, contiguous = ct .!= 0)
end
n_total(pia::DeviceParticleIndexerArray, species::Integer) = Int(ccall((:mb_pia_n_total, libmb), Int64, (Ptr{Cvoid}, Int64), pia.h, species))
"check_pia_is_correct (particles.jl:863-907)"
function check_pia(pia::DeviceParticleIndexerArray, species::Integer)
    ok = Ref{Int32}(0); wh = Ref{Int64}(0)
    check(ccall((:mb_check_pia, libmb), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int64}), pia.h, species, ok, wh))
    (ok[] != 0, wh[])
end

# ------------------------------------------------------------------------------------------------------------ sort
"""
    sort_particles!(gridsort, grid, pv, pia, species)      grid_sorting.jl:58-113
    sort_particles!(gridsort, pv, pia, species)            grid_sorting.jl:128-182 (pv.cell known)

Stable counting sort by cell; `gridsort` is accepted for signature compatibility (the scratch lives in the Context).
"""
function sort_particles!(gridsort, grid, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, species::Integer)
    check(ccall((:mb_sort_particles, libmb), Cint, (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                pv.ctx.h, gridref(grid), pv.h, pia.h, species))
end
function sort_particles!(gridsort, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, species::Integer)
    check(ccall((:mb_sort_particles, libmb), Cint, (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                pv.ctx.h, Ptr{CGrid1D}(C_NULL), pv.h, pia.h, species))
end
"squash_pia!(pv, pia, species) particles.jl:622-682"
squash_pia!(pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, species::Integer=1) =
    check(ccall((:mb_squash_pia, libmb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), pv.ctx.h, pv.h, pia.h, species))
"restore_particle_ordering!(pv, inv_map) particles.jl:1086-1137 -- a no-op on the device (index is the identity)"
restore_particle_ordering!(pv::DeviceParticleVector, inv_map=nothing) =
    check(ccall((:mb_restore_particle_ordering, libmb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), pv.ctx.h, pv.h))

# ------------------------------------------------------------------------------------------------------ collisions
"load_interaction_data entry incl. compute_vhs_factor (collision_utils.jl:98-101,159-201)"
function make_interaction(m_i::Real, m_k::Real, vhs_d::Real, vhs_o::Real, vhs_Tref::Real)
    out = Ref{CInteraction}()
    check(ccall((:mb_make_interaction, libmb), Cint, (Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{CInteraction}),
                m_i, m_k, vhs_d, vhs_o, vhs_Tref, out))
    out[]
end
"estimate_sigma_g_w_max (collision_utils.jl:418-423)"
estimate_sigma_g_w_max(it, m1::Real, m2::Real, T1::Real, T2::Real, Fnum::Real; mult_factor::Real=1.0) =
    ccall((:mb_estimate_sigma_g_w_max, libmb), Cdouble, (Ptr{CInteraction}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble),
          Ref(CInteraction(it)), m1, m2, T1, T2, Fnum, mult_factor)
CInteraction(it::CInteraction) = it

"""
    DeviceCollisionFactors(ctx, n_cells, sigma_g_w_max)

create_collision_factors_array for one species pair (collision_ntc.jl:46-155): per-cell sigma_g_w_max, n_coll,
n_coll_performed, n_eq_w_coll_performed.
"""
mutable struct DeviceCollisionFactors
    ctx::Context
    h::Ptr{Cvoid}
    n_cells::Int64
    function DeviceCollisionFactors(ctx::Context, n_cells::Integer, sigma_g_w_max::Real=0.0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb_cf_create, libmb), Cint, (Ptr{Cvoid}, Int64, Cdouble, Ptr{Ptr{Cvoid}}), ctx.h, n_cells, sigma_g_w_max, out))
        cf = new(ctx, out[], n_cells)
        finalizer(c -> (c.h != C_NULL && ccall((:mb_cf_destroy, libmb), Cint, (Ptr{Cvoid},), c.h); c.h = C_NULL), cf)
        cf
    end
end
Base.fill!(cf::DeviceCollisionFactors, v::Real) = (check(ccall((:mb_cf_fill, libmb), Cint, (Ptr{Cvoid}, Cdouble), cf.h, v)); cf)
upload!(cf::DeviceCollisionFactors, sgwm::Vector{Float64}) = check(ccall((:mb_cf_upload, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), cf.h, sgwm))
function download(cf::DeviceCollisionFactors)
    s = Vector{Float64}(undef, cf.n_cells)
    a = Vector{Int64}(undef, cf.n_cells); b = similar(a); c = similar(a)
    check(ccall((:mb_cf_download, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), cf.h, s, a, b, c))
    (sigma_g_w_max = s, n_coll = a, n_coll_performed = b, n_eq_w_coll_performed = c)
end

cellrange(c::Integer) = (Int64(c), Int64(c))
cellrange(r::AbstractUnitRange) = (Int64(first(r)), Int64(last(r)))

"""
    ntc!(rng, cf, cd, interaction, pv, pia, cell, species, Δt, V; dw_tol=1e-16)                 collision_ntc.jl:338-380
    ntc!(rng, cf, cd, interaction, pv1, pv2, pia, cell, s1, s2, Δt, V; dw_tol=1e-16)            collision_ntc.jl:412-453

`interaction` is the reference's `Matrix{Interaction}` (indexed `[species, species]`), a single `Interaction` or a
`CInteraction`; `cd` (CollisionData scratch) is accepted and ignored; `cell` may be a range.
"""
function ntc!(rng::PhiloxRng, cf::DeviceCollisionFactors, cd, interaction, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray,
              cell, species::Integer, Δt::Real, V::Real; dw_tol::Real=1e-16, equal_weight::Bool=false)
    lo, hi = cellrange(cell)
    it = interaction isa AbstractMatrix ? interaction[species, species] : interaction
    check(ccall((:mb_ntc, libmb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CInteraction}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Cdouble, Cdouble, Int32, UInt32, UInt32),
                pv.ctx.h, cf.h, Ref(CInteraction(it)), pv.h, pia.h, lo, hi, species, Δt, V, dw_tol, equal_weight, rng.timestep, rng.substream))
end
function ntc!(rng::PhiloxRng, cf::DeviceCollisionFactors, cd, interaction, pv1::DeviceParticleVector, pv2::DeviceParticleVector,
              pia::DeviceParticleIndexerArray, cell, s1::Integer, s2::Integer, Δt::Real, V::Real; dw_tol::Real=1e-16, equal_weight::Bool=false)
    lo, hi = cellrange(cell)
    it = interaction isa AbstractMatrix ? interaction[s1, s2] : interaction
    check(ccall((:mb_ntc2, libmb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CInteraction}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cdouble, Cdouble, Cdouble,
                 Int32, UInt32, UInt32),
                pv1.ctx.h, cf.h, Ref(CInteraction(it)), pv1.h, pv2.h, pia.h, lo, hi, s1, s2, Δt, V, dw_tol, equal_weight, rng.timestep, rng.substream))
end
"ntc_equal_weight! (collision_ntc.jl:479-521, :554-595)"
ntc_equal_weight!(rng::PhiloxRng, cf, cd, interaction, pv::DeviceParticleVector, pia, cell, species::Integer, Δt, V) =
    ntc!(rng, cf, cd, interaction, pv, pia, cell, species, Δt, V; equal_weight=true)
ntc_equal_weight!(rng::PhiloxRng, cf, cd, interaction, pv1::DeviceParticleVector, pv2::DeviceParticleVector, pia, cell, s1::Integer,
                  s2::Integer, Δt, V) = ntc!(rng, cf, cd, interaction, pv1, pv2, pia, cell, s1, s2, Δt, V; equal_weight=true)

"swpm!(rng, cf_swpm, cd, interaction, pv, pia, cell, species, G, Δt, V) collision_swpm.jl:201-287 (cf holds sigma_g_max)"
function swpm!(rng::PhiloxRng, cf::DeviceCollisionFactors, cd, interaction, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray,
               cell, species::Integer, G::Real, Δt::Real, V::Real)
    lo, hi = cellrange(cell)
    it = interaction isa AbstractMatrix ? interaction[species, species] : interaction
    check(ccall((:mb_swpm, libmb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CInteraction}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Cdouble, Cdouble, UInt32, UInt32),
                pv.ctx.h, cf.h, Ref(CInteraction(it)), pv.h, pia.h, lo, hi, species, G, Δt, V, rng.timestep, rng.substream))
end

"fp_linear!(rng, cd_fp, interaction, species_data, pv, pia, cell, species, Δt, V) collision_fp.jl:24-125"
function fp_linear!(rng::PhiloxRng, cd_fp, interaction, species_data, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray,
                    cell, species::Integer, Δt::Real, V::Real)
    lo, hi = cellrange(cell)
    it = interaction isa AbstractMatrix ? interaction[species, species] : interaction
    check(ccall((:mb_fp_linear, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CInteraction}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Cdouble, UInt32, UInt32),
                pv.ctx.h, Ref(CInteraction(it)), species_data[species].mass, pv.h, pia.h, lo, hi, species, Δt, V, rng.timestep, rng.substream))
end

# ------------------------------------------------------------------------------------------------------ convection
"""
    convect_particles!(rng, grid, boundaries, pv, pia, species, species_data, Δt)                    convection_1D.jl:130-157
    convect_particles!(rng, grid, boundaries, pv, pia, species, species_data, surf_props, Δt)        convection_1D.jl:176-206
    convect_particles_and_compute_cell!(...same two arities...)                                      convection_1D.jl:225-307

With `surf_props` the call returns a 11 x 2 matrix (per wall: np, flux_incident, flux_reflected, force[1:3], normal_pressure,
shear_pressure[1:3], kinetic_energy_flux), already scaled like surface_props_scale! (surface_props.jl:144-160), and also
stores it into `surf_props` if that is a reference `SurfProps`.
"""
function convect_impl(rng::PhiloxRng, grid, boundaries, pv, pia, species, species_data, Δt, surf::Bool, compute_cell::Bool)
    s22 = surf ? zeros(Float64, 11, 2) : nothing
    check(ccall((:mb_convect_particles, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{CWalls1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cdouble}, Cdouble, Int32, UInt32, UInt32),
                pv.ctx.h, gridref(grid), Ref(boundaries isa CWalls1D ? boundaries : CWalls1D(boundaries)), pv.h, pia.h, species,
                species_data[species].mass, s22 === nothing ? Ptr{Cdouble}(C_NULL) : pointer(s22), Δt, compute_cell, rng.timestep, rng.substream))
    s22
end
convect_particles!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data, Δt::Real) =
    (convect_impl(rng, grid, boundaries, pv, pia, species, species_data, Δt, false, false); nothing)
convect_particles!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data, surf_props, Δt::Real) =
    convect_impl(rng, grid, boundaries, pv, pia, species, species_data, Δt, true, false)
convect_particles_and_compute_cell!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data, Δt::Real) =
    (convect_impl(rng, grid, boundaries, pv, pia, species, species_data, Δt, false, true); nothing)
convect_particles_and_compute_cell!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data, surf_props,
                                    Δt::Real) = convect_impl(rng, grid, boundaries, pv, pia, species, species_data, Δt, true, true)

# ------------------------------------------------------------------------------------------------------ surface properties
"""
    DeviceSurfProps(ctx)                                                                            surface_props.jl:22-50

SurfProps of the two walls of a 1-D grid (one species) on the device.  Passing it as `surf_props` to `convect_particles!` keeps the
whole step asynchronous: cleared, accumulated and scaled on the stream (convection_1D.jl:176-206, surface_props.jl:77-160).
`download(s)` returns the 11 x 2 matrix (per wall: np, flux_incident, flux_reflected, force[1:3], normal_pressure,
shear_pressure[1:3], kinetic_energy_flux).
"""
mutable struct DeviceSurfProps
    ctx::Context
    h::Ptr{Cvoid}
    function DeviceSurfProps(ctx::Context)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mb_surf_create, libmb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), ctx.h, out))
        s = new(ctx, out[])
        finalizer(q -> (q.h != C_NULL && ccall((:mb_surf_destroy, libmb), Cint, (Ptr{Cvoid},), q.h); q.h = C_NULL), s)
        s
    end
end
clear_props!(s::DeviceSurfProps) = check(ccall((:mb_surf_clear, libmb), Cint, (Ptr{Cvoid},), s.h))
upload!(s::DeviceSurfProps, m::Matrix{Float64}) = check(ccall((:mb_surf_upload, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, m))
function download(s::DeviceSurfProps)
    m = zeros(Float64, 11, 2)
    check(ccall((:mb_surf_download, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, m))
    m
end
"avg_props!(surf_props_avg, surf_props, n_avg_timesteps)   surface_props.jl:202-222"
avg_props!(avg::DeviceSurfProps, s::DeviceSurfProps, n_avg_timesteps::Integer) =
    check(ccall((:mb_surf_avg, libmb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), avg.h, s.h, n_avg_timesteps))
"reduce_surf_props!(surf_props_target, surf_props_chunks)   surface_props.jl:232-252; `across_ranks`: also over the NCCL communicator"
function reduce_surf_props!(target::DeviceSurfProps, chunks::Vector{DeviceSurfProps}; across_ranks::Bool=false)
    hs = Ptr{Cvoid}[c.h for c in chunks]
    check(ccall((:mb_surf_reduce, libmb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Int32), target.h, hs, length(hs), across_ranks))
end
function convect_surf_impl(rng::PhiloxRng, grid, boundaries, pv, pia, species, species_data, surf::DeviceSurfProps, Δt, compute_cell::Bool)
    check(ccall((:mb_convect_particles_surf, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{CWalls1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Int32, UInt32, UInt32),
                pv.ctx.h, gridref(grid), Ref(boundaries isa CWalls1D ? boundaries : CWalls1D(boundaries)), pv.h, pia.h, species,
                species_data[species].mass, surf.h, Δt, compute_cell, rng.timestep, rng.substream))
    nothing
end
convect_particles!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data, surf_props::DeviceSurfProps,
                   Δt::Real) = convect_surf_impl(rng, grid, boundaries, pv, pia, species, species_data, surf_props, Δt, false)
convect_particles_and_compute_cell!(rng::PhiloxRng, grid, boundaries, pv::DeviceParticleVector, pia, species::Integer, species_data,
                                    surf_props::DeviceSurfProps, Δt::Real) =
    convect_surf_impl(rng, grid, boundaries, pv, pia, species, species_data, surf_props, Δt, true)

# ------------------------------------------------------------------------------------------------------ properties
"""
    DevicePhysProps(ctx, n_cells, n_species, moment_powers; Tref=300.0, ndens_not_Np=false)       physical_props.jl:24-71
"""
mutable struct DevicePhysProps
    ctx::Context
    h::Ptr{Cvoid}
    n_cells::Int64
    n_species::Int64
    n_moments::Int64
    function DevicePhysProps(ctx::Context, n_cells::Integer, n_species::Integer, moment_powers::Vector{<:Integer}=Int[]; Tref::Real=300.0,
                             ndens_not_Np::Bool=false)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        mp = Int32.(moment_powers)
        check(ccall((:mb_props_create, libmb), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int32}, Cdouble, Int32, Ptr{Ptr{Cvoid}}),
                    ctx.h, n_cells, n_species, length(mp), mp, Tref, ndens_not_Np, out))
        p = new(ctx, out[], n_cells, n_species, length(mp))
        finalizer(q -> (q.h != C_NULL && ccall((:mb_props_destroy, libmb), Cint, (Ptr{Cvoid},), q.h); q.h = C_NULL), p)
        p
    end
end
"""
    download(props) -> (lpa, np, n, v, T, moments) with the reference's shapes: np/n/T [cell, species], v [3, cell, species],
    moments [moment, cell, species] (physical_props.jl:24-37; Julia column-major == the library's C order reversed)
"""
function download(p::DevicePhysProps)
    lpa = Vector{Float64}(undef, p.n_species)
    np = Matrix{Float64}(undef, p.n_cells, p.n_species); n = similar(np); T = similar(np)
    v = Array{Float64,3}(undef, 3, p.n_cells, p.n_species)
    mom = Array{Float64,3}(undef, p.n_moments, p.n_cells, p.n_species)
    check(ccall((:mb_props_download, libmb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                p.h, lpa, np, n, v, T, p.n_moments > 0 ? pointer(mom) : Ptr{Cdouble}(C_NULL)))
    (lpa = lpa, np = np, n = n, v = v, T = T, moments = mom)
end
"clear_props! (physical_props.jl:256-266)"
clear_props!(p::DevicePhysProps) = check(ccall((:mb_props_clear, libmb), Cint, (Ptr{Cvoid},), p.h))
"avg_props!(avg, props, n_avg_timesteps) (physical_props.jl:281-299)"
avg_props!(avg::DevicePhysProps, p::DevicePhysProps, n_avg_timesteps::Integer) =
    check(ccall((:mb_props_avg, libmb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), avg.h, p.h, n_avg_timesteps))

handles(pvs::Vector{DeviceParticleVector}) = Ptr{Cvoid}[pv.h for pv in pvs]
masses(species_data) = Float64[sd.mass for sd in species_data]

"compute_props!(particles, pia, species_data, phys_props) physical_props.jl:104-154"
function compute_props!(particles::Vector{DeviceParticleVector}, pia::DeviceParticleIndexerArray, species_data, props::DevicePhysProps;
                        with_moments::Bool=false)
    check(ccall((:mb_compute_props, libmb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}, Int32),
                pia.ctx.h, handles(particles), pia.h, masses(species_data), props.h, with_moments))
end
"compute_props_with_total_moments! physical_props.jl:168-245"
compute_props_with_total_moments!(particles::Vector{DeviceParticleVector}, pia, species_data, props::DevicePhysProps) =
    compute_props!(particles, pia, species_data, props; with_moments=true)
"compute_props_sorted!(particles, pia, species_data, phys_props[, grid][, cell_chunk]) physical_props.jl:317-454"
function compute_props_sorted!(particles::Vector{DeviceParticleVector}, pia::DeviceParticleIndexerArray, species_data, props::DevicePhysProps,
                               grid=nothing, cell_chunk::AbstractUnitRange=1:pia.n_cells)
    g = grid === nothing ? Ptr{CGrid1D}(C_NULL) : gridref(grid)
    check(ccall((:mb_compute_props_sorted, libmb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{CGrid1D}, Int64, Int64),
                pia.ctx.h, handles(particles), pia.h, masses(species_data), props.h, g, first(cell_chunk), last(cell_chunk)))
end
compute_props_sorted!(particles::Vector{DeviceParticleVector}, pia::DeviceParticleIndexerArray, species_data, props::DevicePhysProps,
                      cell_chunk::AbstractUnitRange) = compute_props_sorted!(particles, pia, species_data, props, nothing, cell_chunk)

# --------------------------------------------------------------------------------------------------------- merging
"""
    merge_octree_N2_based!(rng, octree, pv, pia, cell, species, target_np[, grid]; threshold=-1)   merging_octree_N2.jl:1060-1094

With a cell range only the cells with `n_local > threshold` are merged (threshold < 0: all), which is the
`if pia.indexer[cell, species].n_local > threshold` test of the drivers (couette_varweight_octree.jl:93-98) moved onto the device.
"""
function merge_octree_N2_based!(rng::PhiloxRng, octree, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, cell, species::Integer,
                                target_np::Integer, grid=nothing; threshold::Integer=-1)
    lo, hi = cellrange(cell)
    g = grid === nothing ? Ptr{CGrid1D}(C_NULL) : gridref(grid)
    oc = octree isa COctreeParams ? octree : COctreeParams(octree)
    check(ccall((:mb_merge_octree_N2, libmb), Cint,
                (Ptr{Cvoid}, Ptr{COctreeParams}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Ptr{CGrid1D}, UInt32, UInt32),
                pv.ctx.h, Ref(oc), pv.h, pia.h, lo, hi, species, threshold, target_np, g, rng.timestep, rng.substream))
end

"GridN2Merge parameters (merging_grid.jl:72-116) as the C ABI takes them"
struct CGridMergeParams
    Nx::Int32
    Ny::Int32
    Nz::Int32
    extent_multiplier::NTuple{3,Float64}
end
CGridMergeParams(mg) = CGridMergeParams(mg.Nx, mg.Ny, mg.Nz, (mg.extent_multiplier[1], mg.extent_multiplier[2], mg.extent_multiplier[3]))
"""
    merge_grid_based!(rng, merging_grid, pv, pia, cell, species, species_data, phys_props::DevicePhysProps[, grid]; threshold=-1)
    merge_grid_based!(rng, merging_grid, pv, pia, cell, species, species_data, vx_extent, vy_extent, vz_extent[, grid]; threshold=-1)
                                                                                                     merging_grid.jl:597-703
`cell` may be a range; only cells with `n_local > threshold` are merged (threshold < 0: all).
"""
function merge_grid_based!(rng::PhiloxRng, merging_grid, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, cell, species::Integer,
                           species_data, phys_props::DevicePhysProps, grid=nothing; threshold::Integer=-1)
    lo, hi = cellrange(cell)
    g = grid === nothing ? Ptr{CGrid1D}(C_NULL) : gridref(grid)
    mg = merging_grid isa CGridMergeParams ? merging_grid : CGridMergeParams(merging_grid)
    check(ccall((:mb_merge_grid_based, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGridMergeParams}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{CGrid1D},
                 UInt32, UInt32),
                pv.ctx.h, Ref(mg), pv.h, pia.h, lo, hi, species, species_data[species].mass, phys_props.h, Ptr{Float64}(C_NULL), threshold, g,
                rng.timestep, rng.substream))
end
function merge_grid_based!(rng::PhiloxRng, merging_grid, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, cell, species::Integer,
                           species_data, vx_extent, vy_extent, vz_extent, grid=nothing; threshold::Integer=-1)
    lo, hi = cellrange(cell)
    g = grid === nothing ? Ptr{CGrid1D}(C_NULL) : gridref(grid)
    mg = merging_grid isa CGridMergeParams ? merging_grid : CGridMergeParams(merging_grid)
    ext = Float64[vx_extent[1], vx_extent[2], vy_extent[1], vy_extent[2], vz_extent[1], vz_extent[2]]
    check(ccall((:mb_merge_grid_based, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGridMergeParams}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{CGrid1D},
                 UInt32, UInt32),
                pv.ctx.h, Ref(mg), pv.h, pia.h, lo, hi, species, species_data[species].mass, Ptr{Cvoid}(C_NULL), ext, threshold, g,
                rng.timestep, rng.substream))
end

# ------------------------------------------------------------------------------------------ initial conditions
"""
    sample_particles_equal_weight!(rng, pv, pia, cell, species, nparticles, m, T, Fnum, xlo, xhi, ylo, yhi, zlo, zhi;
                                   distribution=:Maxwellian, vx0=0.0, vy0=0.0, vz0=0.0)         distributions_and_sampling.jl:477-509
    sample_particles_equal_weight!(rng, grid, pv, pia, species, species_data, ppc::Integer, T, Fnum[, cell_chunk])   grid_uniform1D.jl:117-152
    sample_particles_equal_weight!(rng, grid, pv, pia, species, species_data, ndens::Float64, T, Fnum[, cell_chunk]) grid_uniform1D.jl:154-219

Sampled on the device (one Philox stream per cell); `cell` may be a range.
"""
function sample_particles_equal_weight!(rng::PhiloxRng, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, cell, species::Integer,
                                        nparticles::Integer, m::Real, T::Real, Fnum::Real, xlo, xhi, ylo, yhi, zlo, zhi;
                                        distribution=:Maxwellian, vx0=0.0, vy0=0.0, vz0=0.0)
    lo, hi = cellrange(cell)
    box = Float64[xlo, xhi, ylo, yhi, zlo, zhi]
    v0 = Float64[vx0, vy0, vz0]
    check(ccall((:mb_sample_particles_equal_weight, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Float64, Float64, Float64, Float64, Ptr{Float64},
                 Int32, Ptr{Float64}, UInt32, UInt32),
                pv.ctx.h, Ptr{CGrid1D}(C_NULL), pv.h, pia.h, lo, hi, species, nparticles, 0.0, m, T, Fnum, box,
                distribution == :BKW ? 1 : 0, v0, rng.timestep, rng.substream))
end
function sample_particles_equal_weight!(rng::PhiloxRng, grid, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, species::Integer,
                                        species_data, ppc_or_ndens::Real, T::Real, Fnum::Real, cell_chunk=1:pia.n_cells)
    lo, hi = cellrange(cell_chunk)
    np_ = ppc_or_ndens isa Integer ? Int64(ppc_or_ndens) : Int64(-1)
    nd = ppc_or_ndens isa Integer ? 0.0 : Float64(ppc_or_ndens)
    check(ccall((:mb_sample_particles_equal_weight, libmb), Cint,
                (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Float64, Float64, Float64, Float64, Ptr{Float64},
                 Int32, Ptr{Float64}, UInt32, UInt32),
                pv.ctx.h, gridref(grid), pv.h, pia.h, lo, hi, species, np_, nd, species_data[species].mass, T, Fnum, Ptr{Float64}(C_NULL),
                0, Ptr{Float64}(C_NULL), rng.timestep, rng.substream))
end
"""
    sample_on_grid!(rng, vdf::Symbol, pv, pia, cell, species, nv, m, T, n_total, xlo, xhi, ylo, yhi, zlo, zhi;
                    v_mult=3.5, cutoff_mult=3.5, noise=0.0, v_offset=[0.0, 0.0, 0.0])              distributions_and_sampling.jl:312-346

`vdf` is `:maxwellian` or `:bkw` (the BKW distribution at scaled time 0); every cell of `cell` receives the weighted grid sample
and its indexer is set as `ParticleIndexerArray(n_sampled)` does.  Returns `n_sampled` per cell.
"""
function sample_on_grid!(rng::PhiloxRng, vdf::Symbol, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, cell, species::Integer,
                         nv::Integer, m::Real, T::Real, n_total::Real, xlo, xhi, ylo, yhi, zlo, zhi;
                         v_mult=3.5, cutoff_mult=3.5, noise=0.0, v_offset=[0.0, 0.0, 0.0])
    lo, hi = cellrange(cell)
    box = Float64[xlo, xhi, ylo, yhi, zlo, zhi]
    vo = Float64.(v_offset)
    n = Ref{Int64}(0)
    check(ccall((:mb_sample_on_grid, libmb), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Float64, Float64, Float64, Ptr{Float64}, Float64, Float64,
                 Float64, Ptr{Float64}, UInt32, UInt32, Ptr{Int64}),
                pv.ctx.h, vdf == :bkw ? 1 : 0, pv.h, pia.h, lo, hi, species, nv, m, T, n_total, box, v_mult, cutoff_mult, noise, vo,
                rng.timestep, rng.substream, n))
    n[]
end

# ------------------------------------------------------------------------------------------------- slab exchange
"128-byte NCCL unique id (rank 0); distribute it with MPI.jl / Distributed / a file"
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:mb_comm_unique_id, libmb), Cint, (Ptr{UInt8},), id))
    id
end
comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ccall((:mb_comm_init, libmb), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), ctx.h, id, rank, nranks))
"""
    exchange_particles!(ctxs, slabs, pv_chunks, pia_chunks, species)

`exchange_particles!(exchanger, pv_chunks, pia_chunks, cell_chunks, species)` (parallel.jl:443-450) for logical chunks of one process:
chunk `i` owns `slabs[i] = slab(grid, i - 1, n_chunks)` in its own `Context`; `sort_particles!` of every chunk afterwards plays the
role of `sort_particles_after_exchange!` (parallel.jl:467-532).
"""
function exchange_particles!(ctxs::Vector{Context}, slabs::Vector{DeviceGrid1D}, pv_chunks::Vector{DeviceParticleVector},
                             pia_chunks::Vector{DeviceParticleIndexerArray}, species::Integer=1)
    n = length(ctxs)
    hc = Ptr{Cvoid}[c.h for c in ctxs]; hp = Ptr{Cvoid}[p.h for p in pv_chunks]; hi = Ptr{Cvoid}[p.h for p in pia_chunks]
    gs = CGrid1D[s.c for s in slabs]
    check(ccall((:mb_exchange_chunks, libmb), Cint, (Int32, Ptr{Ptr{Cvoid}}, Ptr{CGrid1D}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Int64),
                n, hc, gs, hp, hi, species))
    nothing
end
"0: edge exchange when the layout allows it (no host synchronisation), 1: always the full exchange; same on all ranks"
exchange_set_mode!(ctx::Context, mode::Integer) = check(ccall((:mb_exchange_set_mode, libmb), Cint, (Ptr{Cvoid}, Int32), ctx.h, mode))
"""
    exchange_particles!(ctx, slab, pv, pia, species; counts=false)

Slab replacement of `exchange_particles!` + `sort_particles_after_exchange!` (parallel.jl:281-532): call between
`convect_particles!` and `sort_particles!`; the sort drops the leavers and places the arrivals.
"""
function exchange_particles!(ctx::Context, slab::DeviceGrid1D, pv::DeviceParticleVector, pia::DeviceParticleIndexerArray, species::Integer=1;
                             counts::Bool=false)
    if counts
        s = zeros(Int64, 2); r = zeros(Int64, 2)
        check(ccall((:mb_exchange_slab, libmb), Cint, (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}),
                    ctx.h, Ref(slab.c), pv.h, pia.h, species, s, r))
        return (sent = s, received = r)
    end
    check(ccall((:mb_exchange_slab, libmb), Cint, (Ptr{Cvoid}, Ptr{CGrid1D}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}),
                ctx.h, Ref(slab.c), pv.h, pia.h, species, Ptr{Int64}(C_NULL), Ptr{Int64}(C_NULL)))
    nothing
end

end # module

# GPIFdtdB200.jl -- the reference-side binding of libgpifdtd.so (include/gpifdtd.h).
#
# What a GeoPhyInv maintainer adds to `src/fdtd/` to swap the ParallelStencil kernels for the B200 engine.
# Every function below REPLACES THE BODY of the reference function named in its comment with `ccall`s;
# construction (`SeisForwExpt`), `Medium`, `AGeom`, `Srcs/Recs`, `update!` call order and the records layout
# are untouched.  Julia is not installed in the build image, so this file is exercised only through its
# line-by-line Python twin (geophyinv.jl_b200/engine.py + host/fdtd.py), which the test-suite runs.
#
# Conventions: Julia arrays are column-major [z,(y),x] Float32 -- exactly what the ABI takes; sparse matrices
# are `SparseMatrixCSC{Float32,Int64}` whose `colptr/rowval/nzval` are passed as they are (1-based).
#
# Loading: `include("GPIFdtdB200.jl")` from inside `module GeoPhyInv` (after src/fdtd/fdtd.jl, so that the names below exist).  A
# submodule does not see its parent's bindings, so everything it uses from GeoPhyInv is imported explicitly:
#   FdtdElastic (src/physics_types.jl:17-49), _fd_order / _fd_npml / _fd_nbound / _fd_npextend (src/GeoPhyInv.jl:85-92).
module GPIFdtdB200

using SparseArrays
using ..GeoPhyInv: FdtdElastic, _fd_order, _fd_npml, _fd_nbound, _fd_npextend

const LIB = get(ENV, "GPI_LIB", joinpath(@__DIR__, "..", "geophyinv.jl_b200", "libgpifdtd.so"))

const ABI_VERSION = Int32(2)
const FIELDS = [:p, :vx, :vy, :vz, :tauxx, :tauyy, :tauzz, :tauxy, :tauxz, :tauyz,
    :dpdx, :dpdy, :dpdz, :dvxdx, :dvydy, :dvzdz, :dvxdy, :dvxdz, :dvydx, :dvydz, :dvzdx, :dvzdy,
    :dtauxxdx, :dtauyydy, :dtauzzdz, :dtauxydx, :dtauxydy, :dtauxzdx, :dtauxzdz, :dtauyzdy, :dtauyzdz]
field_id(f::Symbol) = Int32(findfirst(==(f), FIELDS) - 1)
const PARAMS = Dict(:invK => 0, :rho => 1, :invlambda => 2, :invmu => 3)
const FACES = Dict(:zmin => 1, :zmax => 2, :ymin => 4, :ymax => 8, :xmin => 16, :xmax => 32)
face_mask(faces) = Int32(mapreduce(f -> get(FACES, f, 0), |, faces; init = 0))
const MODES = Dict(:forward => 0, :forward_save => 1, :adjoint => 2)
# OR-ed into the mode of gpi_run (include/gpifdtd.h): FD-Born scattering sources pw 1 -> pw 2 (FdtdAcoustic{Born}, born.jl:1-12);
# exact-transpose rho imaging for LinearMap's adjoint (combine_gmodrho! without the one-cell shift, gradient.jl:53-56)
const GPI_RUN_BORN = Int32(0x100)
const GPI_RUN_UNSHIFTED_RHO = Int32(0x200)

# mirrors `gpi_config` (include/gpifdtd.h); isbits, passed by reference
struct GpiConfig
    abi_version::Int32; ndims::Int32; physics::Int32; order::Int32
    n::NTuple{3,Int32}; nt::Int32; npml::Int32; nbound::Int32
    pml_faces::Int32; rigid_faces::Int32; stressfree_faces::Int32
    npw::Int32; nshots::Int32; store_boundary::Int32
    nsnaps::Int32; snaps_field::Int32; device::Int32; shot_batch::Int32
    slab_rank::Int32; slab_nranks::Int32
    dt::Float64; dtI::Float64; d::NTuple{3,Float64}; dI::NTuple{3,Float64}
end

mutable struct Engine
    h::Ptr{Cvoid}
    nt::Int
end

lasterror(h) = unsafe_string(ccall((:gpi_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(e::Engine, rc) = rc == 0 || error("gpifdtd: ", lasterror(e.h))

# ---- P_x_worker_x_pw(...) / P_x_worker_x_pw_x_ss(...)  (src/fdtd/fdtd.jl:340-528) --------------------
# called once per worker from the `ddata` init closure (fdtd.jl:261-266) instead of allocating Data.Arrays
function Engine(pac, sschunk; device = -1, shot_batch = 0, slab = (0, 1))
    N = ndims(pac.medium)
    n = length.(pac.exmedium.grid)
    cfg = GpiConfig(ABI_VERSION, N, pac.attrib_mod isa FdtdElastic ? 1 : 0, _fd_order,
        (n[1], N == 3 ? n[2] : 1, n[end]), pac.ic[:nt], _fd_npml, _fd_nbound,
        face_mask(pac.pml_faces), face_mask(pac.rigid_faces), face_mask(pac.stressfree_faces),
        pac.ic[:npw], length(sschunk), pac.attrib_mod.mode == :forward_save ? 1 : 0,   # fdtd.jl:445-455
        length(pac.itsnaps), field_id(pac.snaps_field), device, shot_batch, slab[1], slab[2],
        pac.fc[:dt], pac.fc[:dtI],
        (pac.fc[:dz], N == 3 ? pac.fc[:dy] : 1.0, pac.fc[:dx]), (pac.fc[:dzI], N == 3 ? pac.fc[:dyI] : 1.0, pac.fc[:dxI]))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:gpi_create, LIB), Cint, (Ref{GpiConfig}, Ref{Ptr{Cvoid}}), cfg, out)
    rc == 0 || error("gpifdtd: ", lasterror(C_NULL))
    e = Engine(out[], pac.ic[:nt])
    finalizer(x -> ccall((:gpi_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), e)
    return e
end

# ---- update!(pac::P_common, medium)  (src/fdtd/medium.jl:131-140): copyto!(mod[name], ...) + update_dmod! ------
# `padarray!(pac.exmedium, ...)` (medium.jl:134, media.jl:260-275) is not needed for the upload: the un-extended
# array goes to the device and the engine replicates it into the PML there (gpi_set_medium_interior).
function update_medium!(e::Engine, pac)
    N = ndims(pac.medium)
    dims = N == 3 ? (:z, :y, :x) : (:z, :x)
    lo3 = Int32[0, 0, 0]; n3 = Int32[1, 1, 1]
    for (q, d) in enumerate(dims)
        slot = N == 3 ? q : (q == 1 ? 1 : 3)
        lo3[slot] = Symbol(d, :min) in pac.pml_faces ? _fd_npextend : 0
        n3[slot] = length(pac.medium.grid[q])
    end
    if Set(names(pac.mod)[1]) in (Set([:invK, :rho]), Set([:invlambda, :invmu, :rho]))
        # the stock parameterisations: vp, (vs,) rho as they are; the getters of media.jl:103-130 are broadcast on the device
        m = pac.medium
        vs = hasproperty(m, :vs) ? Array{Float32}(m.vs.m) : Ptr{Float32}(C_NULL)
        check(e, ccall((:gpi_set_medium_fields, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}),
            e.h, Array{Float32}(m.vp.m), vs, Array{Float32}(m.rho.m), n3, lo3))
    else
        for name in names(pac.mod)[1]
            a = Array{Float32}(pac.medium[name])                         # [mz,(my),mx], no padding
            check(e, ccall((:gpi_set_medium_interior, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}),
                e.h, PARAMS[name], a, n3, lo3))
        end
    end
    check(e, ccall((:gpi_update_dmod, LIB), Cint, (Ptr{Cvoid},), e.h))   # update_dmod! + store_invav*! (medium.jl:143-221)
end

# model-vector path (medium.jl:31-52): only the interior view of `mod` changes and the padding keeps its old
# values, so the extended array is uploaded as it is
function update_mod!(e::Engine, pac)
    for name in names(pac.mod)[1]
        a = Array{Float32}(pac.mod[name])                            # [nz,(ny),nx] on the extended grid
        check(e, ccall((:gpi_set_medium, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), e.h, PARAMS[name], a))
    end
    check(e, ccall((:gpi_update_dmod, LIB), Cint, (Ptr{Cvoid},), e.h))
end

# ---- update_pml!(pac)  (src/fdtd/cpml.jl:144-155): the host loop stays, the three copyto! become one call -------
function set_pml!(e::Engine, dfield::Symbol, a::Vector{Float32}, b::Vector{Float32}, kI::Vector{Float32})
    check(e, ccall((:gpi_set_pml, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
        e.h, field_id(dfield), a, b, kI))
end

# ---- update!(pass, ipw, iss, ageomss, pac, ::Srcs / ::Recs)  (src/fdtd/ageom.jl:33-58) ---------------------------
# `S` is the SparseMatrixCSC the reference builds with get_proj_matrix; kind 0 = spray, 1 = interpolation
function set_sparse!(e::Engine, kind, ipw, issp, field::Symbol, S::SparseMatrixCSC{Float32,Int64})
    check(e, ccall((:gpi_set_sparse, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}),
        e.h, kind, ipw - 1, issp - 1, field_id(field), size(S, 2), S.colptr, S.rowval, S.nzval))
end

# ---- fill_wavelets!(ipw, iss, wavelets, srcwav, src_types)  (src/fdtd/source.jl:24-58) ---------------------------
# w[nt, ns] already transformed by get_source (source.jl:3-19)
function set_wavelets!(e::Engine, ipw, issp, field::Symbol, w::Matrix{Float32})
    check(e, ccall((:gpi_set_wavelets, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Float32}),
        e.h, ipw - 1, issp - 1, field_id(field), size(w, 2), w))
end

# ---- initialize!(pap), initialize_boundary!, reset_w2!  (src/fdtd/types.jl:41-113) -------------------------------
reset!(e::Engine, what) = check(e, ccall((:gpi_reset, LIB), Cint, (Ptr{Cvoid}, Cint), e.h, what))

# ---- mod_x_proc!(pac, pap, activepw, src_flags)  (src/fdtd/propagate.jl:138-261) ---------------------------------
# the whole shot loop x time loop of this worker; blocking like the reference's remotecall_wait
# `mode_flags`: GPI_RUN_BORN for `FdtdAcoustic{Born}` forward runs (propagate.jl:53-60 selects activepw = [1, 2]),
# GPI_RUN_UNSHIFTED_RHO for the adjoint of `LinearMap(pa)` (func_grad.jl:106-120); by default Born runs are recognised from the
# attribute's type parameter the way the reference dispatches add_born_sources_*! (born.jl:1-12).
function mod_x_proc!(e::Engine, pac, activepw, src_flags; mode_flags::Integer = default_mode_flags(pac))
    am = mapreduce(p -> 1 << (p - 1), |, activepw; init = 0)
    sm = mapreduce(i -> src_flags[i] ? 1 << (i - 1) : 0, |, eachindex(src_flags); init = 0)
    mode = Int32(MODES[pac.attrib_mod.mode]) | Int32(mode_flags)
    check(e, ccall((:gpi_run, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint), e.h, mode, am, sm))
end
default_mode_flags(pac) =
    (occursin("Born", string(typeof(pac.attrib_mod))) && pac.attrib_mod.mode != :adjoint) ? GPI_RUN_BORN : Int32(0)

# ---- update_datamat!(rfield, ipw, pac, pap)  (src/fdtd/receiver.jl:17-34): ONE copy per (shot, field) -------------
function update_datamat!(e::Engine, datamat, rfield::Symbol, ipw, issp, iss, nr)
    buf = Matrix{Float32}(undef, e.nt, nr)
    check(e, ccall((:gpi_get_records, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}), e.h, ipw - 1, issp - 1, field_id(rfield), buf))
    datamat[:, 1:nr, iss] .= buf
end

# ---- sum_grads!(pac, pap)  (src/fdtd/gradient.jl:2-11): device stack + one NCCL all-reduce, then one copy ----------
function sum_grads!(e::Engine, pac; allreduce = false)
    allreduce && check(e, ccall((:gpi_allreduce_gradients, LIB), Cint, (Ptr{Cvoid},), e.h))
    for name in names(pac.gradients)[1]
        g = Array{Float32}(undef, size(pac.gradients[name]))
        check(e, ccall((:gpi_get_gradient, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), e.h, PARAMS[name], g))
        pac.gradients[name] .= g
    end
end

# ---- pa[:snaps, i]  (src/fdtd/getprop.jl:10-25) -------------------------------------------------------------------
function get_snap(e::Engine, shape, ipw, issp, isnap)
    out = Array{Float32}(undef, shape)
    check(e, ccall((:gpi_get_snap, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}), e.h, ipw - 1, issp - 1, isnap - 1, out))
    return out
end

# ---- illum_flag (src/fdtd/fdtd.jl:59,73): compute_illum! / stack_illums! (fdtd.jl:556-581; their calls are commented out upstream,
#      propagate.jl:114,236).  set_illum! once after the constructor when pac.illum_flag; stack_illums! after mod_x_proc! -------------
set_illum!(e::Engine, on::Bool = true) = check(e, ccall((:gpi_set_illum, LIB), Cint, (Ptr{Cvoid}, Cint), e.h, on ? 1 : 0))
function stack_illums!(e::Engine, pac)
    ex = Array{Float64}(undef, field_shape(pac, :p))
    check(e, ccall((:gpi_get_illum, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), e.h, ex))
    inner = ntuple(d -> (_fd_npml+1):(size(ex, d)-_fd_npml), ndims(ex))          # the view stack_illums! takes (fdtd.jl:562)
    pac.illum_stack .+= view(ex, inner...)
end

# ---- one process per GPU: the ncclUniqueId travels over Julia Distributed (remotecall_fetch) -----------------------
function nccl_unique_id()
    id = zeros(UInt8, 128)
    ccall((:gpi_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error("ncclGetUniqueId failed")
    return id
end
nccl_init!(e::Engine, id::Vector{UInt8}, rank, nranks) =
    check(e, ccall((:gpi_nccl_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), e.h, id, rank, nranks))

# ---- staggered array shapes for the compiled-in order (src/fields.jl:92-671; `zeros(::Field, attrib_mod, n...)`) ---------
function field_shape(pac, field::Symbol)
    N = ndims(pac.medium)
    n = length.(pac.exmedium.grid)
    out = zeros(Int32, 3)
    rc = ccall((:gpi_field_shape_order, LIB), Cint, (Cint, Cint, Cint, Cint, Ptr{Int32}, Ptr{Int32}),
        N, pac.attrib_mod isa FdtdElastic ? 1 : 0, _fd_order, field_id(field), Int32[n[1], N == 3 ? n[2] : 1, n[end]], out)
    rc == 0 || error("gpifdtd: field ", field, " does not exist for this physics / dimensionality")
    return N == 3 ? Tuple(out) : (out[1], out[3])
end

# ---- FD-Born: update!(pac, medium, medium_pert)  (src/fdtd/medium.jl:103-127) ---------------------------------------
function update_medium_pert!(e::Engine, pac)
    for name in names(pac.δmod)[1]
        check(e, ccall((:gpi_set_medium_pert, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), e.h, PARAMS[name], Array{Float32}(pac.δmod[name])))
    end
    check(e, ccall((:gpi_update_born, LIB), Cint, (Ptr{Cvoid},), e.h))
end

# which stencil kernels the engine launches for this experiment: 0 scalar, 1 float4, 2 TMA-pipelined tiles, 4 order 4
kernel_family(e::Engine) = ccall((:gpi_kernel_family, LIB), Cint, (Ptr{Cvoid},), e.h)

# ---- the remaining exports of include/gpifdtd.h, one thin wrapper each -------------------------------------------------
abi_version() = ccall((:gpi_abi_version, LIB), Cint, ())
function field_shape_order2(ndims, elastic::Bool, field::Symbol, n::NTuple{3,Int})      # the order-2 entry point kept for old callers
    out = zeros(Int32, 3)
    ccall((:gpi_field_shape, LIB), Cint, (Cint, Cint, Cint, Ptr{Int32}, Ptr{Int32}), ndims, elastic ? 1 : 0, field_id(field), Int32[n...], out) == 0 ||
        error("gpifdtd: field ", field, " does not exist for this physics / dimensionality")
    return Tuple(out)
end

# pa[:snaps, i] needs the snapshot steps the constructor derived from `tsnaps` (src/fdtd/fdtd.jl:181-185)
set_snap_steps!(e::Engine, itsnaps::Vector{<:Integer}) =
    check(e, ccall((:gpi_set_snap_steps, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), e.h, length(itsnaps), Int32.(itsnaps)))

# wavefields and medium arrays in the reference's own shapes (debugging, `pap[ipw].w1[:t][field]`, `pac.mod[name]`)
function get_field(e::Engine, pac, ipw, field::Symbol; batch = 1)
    out = Array{Float32}(undef, field_shape(pac, field))
    check(e, ccall((:gpi_get_field, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}), e.h, ipw - 1, batch - 1, field_id(field), out))
    return out
end
set_field!(e::Engine, ipw, field::Symbol, a::Array{Float32}; batch = 1) =
    check(e, ccall((:gpi_set_field, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float32}), e.h, ipw - 1, batch - 1, field_id(field), a))
function get_medium(e::Engine, pac, name::Symbol)
    out = Array{Float32}(undef, size(pac.mod[name]))
    check(e, ccall((:gpi_get_medium, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), e.h, PARAMS[name], out))
    return out
end

# z-slab handles: the global z range this engine owns and a window of rows of a medium array (the host never needs the whole
# 594 x 1106 x 1106 array on every worker)
function slab_range(e::Engine)
    ka = Ref{Int32}(0); kb = Ref{Int32}(0)
    check(e, ccall((:gpi_slab_range, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}), e.h, ka, kb))
    return Int(ka[]), Int(kb[])
end
set_medium_rows!(e::Engine, name::Symbol, rows::Array{Float32}, k_first::Integer) =
    check(e, ccall((:gpi_set_medium_rows, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Cint, Cint), e.h, PARAMS[name], rows, k_first, size(rows, 1)))

# zero-copy hand-off to CUDA.jl: wrap the engine's device buffers with `unsafe_wrap(CuArray, CuPtr{Float32}(ptr), dims)`
function records_device_ptr(e::Engine, ipw, issp, rfield::Symbol)
    p = Ref{Ptr{Cvoid}}(C_NULL); nb = Ref{Int64}(0)
    check(e, ccall((:gpi_records_device_ptr, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}, Ref{Int64}), e.h, ipw - 1, issp - 1, field_id(rfield), p, nb))
    return p[], Int(nb[])
end
function gradient_device_ptr(e::Engine, name::Symbol)
    p = Ref{Ptr{Cvoid}}(C_NULL); nf = Ref{Int64}(0)
    check(e, ccall((:gpi_gradient_device_ptr, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}, Ref{Int64}), e.h, PARAMS[name], p, nf))
    return p[], Int(nf[])
end
set_stream!(e::Engine, stream::Ptr{Cvoid}) = check(e, ccall((:gpi_set_stream, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.h, stream))
synchronize(e::Engine) = check(e, ccall((:gpi_synchronize, LIB), Cint, (Ptr{Cvoid},), e.h))

# the engine-side analogue of the TimerOutputs sections of mod_x_proc! (src/fdtd/propagate.jl:176-233)
struct GpiTimers
    run_ms::Float64; steps::Float64; cell_updates::Float64; stencil_ms::Float64; launches::Float64
    vel_ms::Float64; vel_n::Float64; stress_ms::Float64; stress_n::Float64
    exch_ms::Float64; exch_n::Float64; allreduce_ms::Float64
end
function timers(e::Engine)
    t = Ref(GpiTimers(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0))
    check(e, ccall((:gpi_get_timers, LIB), Cint, (Ptr{Cvoid}, Ref{GpiTimers}), e.h, t))
    return t[]
end

end # module

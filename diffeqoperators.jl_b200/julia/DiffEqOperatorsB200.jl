# DiffEqOperatorsB200.jl -- Julia side of the drop-in boundary (include/deo_b200.h).
#
# The reference has no FFI: its hot path sits behind multiple dispatch (LinearAlgebra.mul! / Base.:* on
# DerivativeOperator, GhostDerivativeOperator, DiffEqOperatorCombination).  This file adds ONE array type
# (`DeviceArray`, an opaque library-owned device buffer) and the `mul!` / `*` methods for it; constructors,
# operator objects, Array/sparse/BandedMatrix concretization stay the reference's own.  Every method reads
# the *fields* of the reference's operator objects and hands them to libdeo_b200.so through `ccall`.
#
# NOTE: Julia is not installed in the build container nor on the GPU box, so this file has never been
# executed.  It is a 1:1 transcription of diffeqoperators.jl_b200/{_lib,apply,device}.py, which bind the
# same entry points through ctypes and ARE exercised by tests/ (see INTEGRATION.md).
module DiffEqOperatorsB200

using LinearAlgebra
using DiffEqOperators
using DiffEqOperators: DerivativeOperator, GhostDerivativeOperator, DiffEqOperatorCombination,
                       AffineBC, PeriodicBC, AtomicBC, MultiDimDirectionalBC, ComposedMultiDimBC

export DeviceArray, deo_init, deo_sync

const libdeo = get(ENV, "DEO_LIB_PATH", joinpath(@__DIR__, "..", "libdeo_b200.so"))

# ---- status / errors (deo_b200.h: every entry point returns int32, 0 == DEO_OK) ----------------------
struct DeoError <: Exception
    code::Int32
    msg::String
end
function last_error()
    buf = Vector{UInt8}(undef, 1024)
    ccall((:deo_last_error, libdeo), Int32, (Ptr{UInt8}, Csize_t), buf, length(buf))
    unsafe_string(pointer(buf))
end
check(rc::Int32) = rc == 0 ? nothing : throw(DeoError(rc, last_error()))

deo_init(device::Integer = 0) = check(ccall((:deo_init, libdeo), Int32, (Int32,), device))
deo_sync() = check(ccall((:deo_sync, libdeo), Int32, ()))

# ---- DeviceArray: the handle type (library-owned device buffer + dims, column-major like Array) -------
mutable struct DeviceArray{T <: Union{Float32, Float64}, N} <: AbstractArray{T, N}
    handle::Ptr{Cvoid}          # deo_buffer*
    dims::NTuple{N, Int}
    function DeviceArray{T, N}(::UndefInitializer, dims::NTuple{N, Int}) where {T, N}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:deo_buffer_create, libdeo), Int32, (Csize_t, Ptr{Ptr{Cvoid}}), prod(dims) * sizeof(T), h))
        a = new{T, N}(h[], dims)
        finalizer(x -> (ccall((:deo_buffer_free, libdeo), Int32, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), a)
        a
    end
end
DeviceArray{T}(::UndefInitializer, dims::Vararg{Int, N}) where {T, N} = DeviceArray{T, N}(undef, dims)
Base.size(a::DeviceArray) = a.dims
Base.similar(a::DeviceArray{T}, dims::Dims = size(a)) where {T} = DeviceArray{T, length(dims)}(undef, dims)
# scalar indexing would be one PCIe transfer per element: forbidden, like CuArray's allowscalar(false)
Base.getindex(::DeviceArray, ::Int...) = error("scalar indexing of a DeviceArray is not supported; use Array(a)")

function Base.copyto!(dst::DeviceArray{T}, src::Array{T}) where {T}          # deo_buffer_upload
    @assert length(dst) == length(src)
    check(ccall((:deo_buffer_upload, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), dst.handle, src, sizeof(src)))
    dst
end
function Base.copyto!(dst::Array{T}, src::DeviceArray{T}) where {T}          # deo_buffer_download (synchronises)
    @assert length(dst) == length(src)
    check(ccall((:deo_buffer_download, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), dst, src.handle, sizeof(dst)))
    dst
end
DeviceArray(a::Array{T, N}) where {T, N} = copyto!(DeviceArray{T, N}(undef, size(a)), a)
Base.Array(a::DeviceArray{T, N}) where {T, N} = copyto!(Array{T, N}(undef, size(a)), a)

# ---- descriptors: field-for-field images of deo_op_desc / deo_bc_desc / deo_plan_desc ----------------------
struct OpDesc
    axis::Int32; kind::Int32; nonuniform::Int32; derivative_order::Int32; len::Int32
    stencil_length::Int32; boundary_stencil_length::Int32; boundary_point_count::Int32
    offside::Int32; reserved::Int32
    stencil_coefs::Ptr{Cvoid}; low_boundary_coefs::Ptr{Cvoid}; high_boundary_coefs::Ptr{Cvoid}; coefficients::Ptr{Cvoid}
end
struct BcDesc
    kind::Int32; per_face::Int32; K_l::Int32; K_r::Int32
    a_l::Ptr{Cvoid}; b_l::Ptr{Cvoid}; a_r::Ptr{Cvoid}; b_r::Ptr{Cvoid}
end
const BC_NONE = BcDesc(0, 0, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL)
struct PlanDesc
    dtype::Int32; ndims::Int32
    dims::NTuple{3, Int64}
    padded::NTuple{3, Int32}
    nops::Int32; accumulate::Int32
    ops::Ptr{OpDesc}
    bc::NTuple{3, BcDesc}
    flags::Int32; reserved::Int32
end

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)

# Flatten the SVector-of-SVector storage of the reference into the dense row-major tables the ABI takes
# (derivative_operator.jl:14-29; shapes in deo_b200.h).
flat(::Type{T}, x::Number) where {T} = T[x]
flat(::Type{T}, x::AbstractArray{<:Number}) where {T} = collect(T, vec(x))
flat(::Type{T}, x) where {T} = isempty(x) ? T[] : reduce(vcat, [flat(T, xi) for xi in x])
# non-uniform upwind: stencil_coefs is a 2 x (len-2bpc) SMatrix of SVectors; the ABI wants set-major order
flat_sets(::Type{T}, m::AbstractMatrix) where {T} = vcat((flat(T, m[s, :]) for s in 1:size(m, 1))...)
flat_sets(::Type{T}, m) where {T} = flat(T, m)

function op_desc(A::DerivativeOperator{T, N, Wind}, keep::Vector{Any}) where {T, N, Wind}
    nonuniform = !(A.dx isa Number)
    st = (Wind && nonuniform) ? flat_sets(T, A.stencil_coefs) : flat(T, A.stencil_coefs)
    lo = (Wind && nonuniform) ? flat_sets(T, A.low_boundary_coefs) : flat(T, A.low_boundary_coefs)
    hi = (Wind && nonuniform) ? flat_sets(T, A.high_boundary_coefs) : flat(T, A.high_boundary_coefs)
    co = collect(T, A.coefficients)
    push!(keep, st, lo, hi, co)
    OpDesc(N - 1, Wind ? 1 : 0, nonuniform ? 1 : 0, A.derivative_order, A.len, A.stencil_length,
           A.boundary_stencil_length, A.boundary_point_count, A.offside, 0,
           pointer(st), isempty(lo) ? C_NULL : pointer(lo), isempty(hi) ? C_NULL : pointer(hi), pointer(co))
end

# One atomic BC for every boundary pencil of an axis, or an (N-1)-D array of them (MultiDimDirectionalBC.BCs).
function bc_desc(::Type{T}, bc::AffineBC, keep) where {T}
    al, ar, bl, br = collect(T, bc.a_l), collect(T, bc.a_r), T[bc.b_l], T[bc.b_r]
    push!(keep, al, ar, bl, br)
    BcDesc(1, 0, length(al), length(ar), pointer(al), pointer(bl), pointer(ar), pointer(br))
end
# vectors: l = u[end], r = u[1] (bc_operators.jl:192) -> DEO_BC_PERIODIC, the wrap-around read
bc_desc(::Type{T}, ::PeriodicBC, keep) where {T} = BcDesc(2, 0, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL)
# arrays: the reference's periodic ghosts are lower = u[1, ...], upper = u[end, ...] of the same pencil
# (multi_dim_bc_operators.jl:221-228), i.e. literally the affine BC a = [1], b = 0 -- which every tiled kernel takes
function periodic_nd_desc(::Type{T}, keep) where {T}
    one_, zero_ = T[1], T[0]
    push!(keep, one_, zero_)
    BcDesc(1, 0, 1, 1, pointer(one_), pointer(zero_), pointer(one_), pointer(zero_))
end
function bc_desc(::Type{T}, bcs::AbstractArray{<:AtomicBC}, keep) where {T}
    if all(b -> b === first(bcs), bcs)                                         # fill(BC, perpsize(...)) : :97-100
        return first(bcs) isa PeriodicBC ? periodic_nd_desc(T, keep) : bc_desc(T, first(bcs), keep)
    end
    all(b -> b isa AffineBC, bcs) || error("per-pencil BC arrays must hold affine BCs")
    Kl = maximum(b -> length(b.a_l), bcs); Kr = maximum(b -> length(b.a_r), bcs)
    nf = length(bcs)
    al = zeros(T, Kl, nf); ar = zeros(T, Kr, nf)          # column f = face pencil f (row-major [face][K] for C)
    for (f, b) in enumerate(vec(bcs))
        al[1:length(b.a_l), f] .= b.a_l
        ar[(Kr - length(b.a_r) + 1):Kr, f] .= b.a_r
    end
    bl = T[b.b_l for b in vec(bcs)]; br = T[b.b_r for b in vec(bcs)]
    push!(keep, al, ar, bl, br)
    BcDesc(1, 1, Kl, Kr, pointer(al), pointer(bl), pointer(ar), pointer(br))
end

# Which boundary operator extends dimension `ax` (multi_dim_bc_operators.jl:54-61, :178-192, :210)
bc_for_axis(Q::AtomicBC, ax, nd) = ax == 1 ? Q : nothing
bc_for_axis(Q::MultiDimDirectionalBC{T, B, D}, ax, nd) where {T, B, D} = ax == D ? Q.BCs : nothing
bc_for_axis(Q::ComposedMultiDimBC, ax, nd) = Q.BCs[ax]
bc_for_axis(::Nothing, ax, nd) = nothing

# (L, Q) terms in application order: ghost_derivative_operator.jl:7-13, composite_operators.jl:64-65
terms(A::DerivativeOperator) = Any[(A, nothing)]
terms(A::GhostDerivativeOperator) = A.L isa DiffEqOperatorCombination ?
    Any[(t[1], A.Q) for op in A.L.ops for t in terms(op)] : Any[(A.L, A.Q)]
terms(A::DiffEqOperatorCombination) = reduce(vcat, [terms(op) for op in A.ops])
axis_of(::DerivativeOperator{T, N}) where {T, N} = N

# ---- plans: built lazily from the operator object and cached per (operator, sizes) --------------------------
mutable struct Plan
    handle::Ptr{Cvoid}
    coeffs::Vector{Any}          # copies of the coefficient vectors the device tables were built from
end
const PLAN_CACHE = IdDict{Any, Dict{Any, Plan}}()

function build_plan(A, ::Type{T}, out_dims::Dims{N}, in_dims::Dims{N}, accumulate::Bool) where {T, N}
    N <= 3 || error("collapse arrays with more than 3 dimensions to (pre, n, post) first (see apply.py:_collapse)")
    keep = Any[]
    ts = terms(A)
    padded = ntuple(a -> a <= N ? Int32(in_dims[a] - out_dims[a] == 2) : Int32(0), 3)
    ops = OpDesc[op_desc(L, keep) for (L, _) in ts]
    bcs = BcDesc[BC_NONE, BC_NONE, BC_NONE]
    for (L, Q) in ts
        ax = axis_of(L)
        padded[ax] == 1 && continue
        b = bc_for_axis(Q, ax, N)
        b === nothing && throw(AssertionError("the differentiated dimension must be padded or carry a boundary condition"))
        bcs[ax] = bc_desc(T, b, keep)
    end
    desc = Ref(PlanDesc(dtype_code(T), N, ntuple(a -> a <= N ? Int64(out_dims[a]) : Int64(1), 3), padded,
                        length(ops), accumulate ? 1 : 0, pointer(ops), (bcs[1], bcs[2], bcs[3]), 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep ops bcs check(ccall((:deo_plan_create, libdeo), Int32, (Ptr{PlanDesc}, Ptr{Ptr{Cvoid}}), desc, h))
    p = Plan(h[], Any[copy(L.coefficients) for (L, _) in ts])
    finalizer(x -> ccall((:deo_plan_destroy, libdeo), Int32, (Ptr{Cvoid},), x.handle), p)
    p
end

# One plan per (operator, sizes).  update_coefficients! mutates A.coefficients in place on the host: the cached plan is
# refreshed with deo_plan_update_coefficients for the operators whose vector changed -- never rebuilt, so a time-stepping
# loop with a time-dependent coeff_func neither leaks device tables nor pays a plan build per step.
function plan_for(A, ::Type{T}, out_dims, in_dims, accumulate) where {T}
    key = (T, out_dims, in_dims, accumulate)
    p = get!(() -> build_plan(A, T, out_dims, in_dims, accumulate), get!(() -> Dict{Any, Plan}(), PLAN_CACHE, A), key)
    for (k, (L, _)) in enumerate(terms(A))
        if L.coefficients != p.coeffs[k]
            c = collect(T, L.coefficients)
            GC.@preserve c check(ccall((:deo_plan_update_coefficients, libdeo), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), p.handle, k - 1, c))
            p.coeffs[k] = copy(L.coefficients)
        end
    end
    p
end

# ---- the methods the path sits behind -------------------------------------------------------------------------
const FusedOperator = Union{DerivativeOperator, GhostDerivativeOperator, DiffEqOperatorCombination}

# mul!(du, A, u): convolutions.jl:17-22, derivative_operator_functions.jl:18-69, ghost_derivative_operator.jl:15-24,
# composite_operators.jl:76-83.  Asynchronous on the library stream; Array(du) / deo_sync() synchronise.
function LinearAlgebra.mul!(du::DeviceArray{T, N}, A::FusedOperator, u::DeviceArray{T, N}; overwrite = true) where {T, N}
    p = plan_for(A, T, size(du), size(u), !overwrite)
    check(ccall((:deo_plan_apply, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.handle, du.handle, u.handle))
    du
end

# A * u: ghost_derivative_operator.jl:26-37 (output = unpadded_size(u)), derivative_operator_functions.jl:150-163
function Base.:*(A::GhostDerivativeOperator, u::DeviceArray{T, N}) where {T, N}
    mul!(DeviceArray{T, N}(undef, size(u)), A, u)
end
function Base.:*(A::DerivativeOperator{T, D}, u::DeviceArray{T, N}) where {T, D, N}
    mul!(DeviceArray{T, N}(undef, ntuple(a -> a == D ? size(u, a) - 2 : size(u, a), N)), A, u)
end
function Base.:*(A::DiffEqOperatorCombination, u::DeviceArray{T, N}) where {T, N}
    all(t -> t[2] !== nothing, terms(A)) || error("`*` of a sum of bare operators along different axes needs an explicit du (pre-padded input)")
    mul!(DeviceArray{T, N}(undef, size(u)), A, u)
end

# Host-buffer form of mul!: one H2D copy of u, the fused kernel, one D2H copy of du (deo_plan_apply_host).
function mul_host!(du::Array{T, N}, A::FusedOperator, u::Array{T, N}) where {T <: Union{Float32, Float64}, N}
    p = plan_for(A, T, size(du), size(u), false)
    check(ccall((:deo_plan_apply_host, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.handle, du, u))
    du
end

# out = u + dt * (A*u): the update an explicit stepper performs right after mul! (test/DerivativeOperators/
# 3D_laplacian.jl:20-24), fused into the store of the tiled kernels (deo_plan_apply_axpy).
function step!(out::DeviceArray{T, N}, A::FusedOperator, u::DeviceArray{T, N}, dt::Real) where {T, N}
    p = plan_for(A, T, size(out), size(u), false)
    check(ccall((:deo_plan_apply_axpy, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), p.handle, out.handle, u.handle, Float64(dt)))
    out
end

# Slab-decomposed plans (one process per GPU): host-buffer mul! on this rank's planes (deo_dist_plan_apply_host).
function mul_host_slab!(du::Array{T, 3}, plan::Plan, u_own::Array{T, 3}) where {T <: Union{Float32, Float64}}
    check(ccall((:deo_dist_plan_apply_host, libdeo), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), plan.handle, du, u_own))
    du
end

# The (du,u,p,t) functor of src/DiffEqOperators.jl:66-75 works unchanged: update_coefficients! mutates
# A.coefficients on the host, plan_for() notices and refreshes the cached plan in place.

end # module

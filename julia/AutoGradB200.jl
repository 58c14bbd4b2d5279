# AutoGradB200.jl -- the Julia side of the drop-in: a device array type whose methods `ccall`
# libagb200.so (include/agb200.h).  Load AFTER `using AutoGrad`; then `Param(DArray(x))` records and
# differentiates through the unmodified AutoGrad.jl tape (src/core.jl) with every array operation on B200.
#
# NOT EXECUTED IN THIS REPOSITORY: no Julia toolchain exists in the build image (DESIGN.md section 1), so this
# file has been written against the reference's sources and the C header, and checked mechanically only as far as
# tests/test_abi_exports.py goes (every ccall below is parsed and its symbol, arity and C types are compared with the
# prototype in include/agb200.h).  The Python package autograd.jl_b200/
# issues the same C calls and is what the tests and the bench drive.
#
# Map (reference file:line -> section of this file)
#   array type, storage, views ............ src/core.jl:66 (the unboxed array the primitives are called on)        §1
#   dotted calls f.(x), f.(a, b) .......... src/core.jl:64-70, src/math.jl:9-74, src/base.jl:20-49,73-74          §2
#   device `back` methods (fused rules) ... src/macros.jl:88-134 (what @primitive emits), src/math.jl, src/base.jl  §3
#   sum / maximum / minimum / prod ........ src/base.jl:202-206,221,245-247; unbroadcast src/unbroadcast.jl:9-26   §4
#   *, adjoint, products with epilogue .... src/base.jl:26, src/linearalgebra.jl:16                                 §5
#   getindex / view / cat / permutedims ... src/getindex.jl:32-37, src/cat.jl:27-31,122-123, src/base.jl:217-218    §6
#   Sparse on a device container .......... src/sparse.jl:50-59, src/addto.jl:41-52,91-98                           §7
#   axpy!, lmul!, multi-tensor update ..... README.md:70-77, examples/charlm.jl:116-118,148-152                     §8
#   deferred elementwise chains ........... src/core.jl:68-70 (one launch per chain; mirrors lazy.py)              §9
#   CUDA graphs, events, @gs, gcnode ...... src/AutoGrad.jl:11, src/core.jl:168,235-237                            §10
#   data parallel ......................... BASELINE.json north_star (no reference counterpart)                    §11
#   input pipeline ........................ examples/mnist.jl:77-96                                                §12
module AutoGradB200

using AutoGrad
import AutoGrad: back, Arg, Value, Param, Sparse, addto!, full, zeroslike, recording, value, unbroadcast
import Base: size, length, eltype, similar, copy, sum, maximum, minimum, prod, *, +, -, adjoint, transpose, reshape,
             vec, zero, fill!, getindex, view, cat, vcat, hcat, permutedims, show, Array, convert
import Base.Broadcast: broadcasted
import LinearAlgebra: axpy!, lmul!, norm, Adjoint, Transpose

export DArray, IArray, init, sync, gemm_bias_act, allreduce_update!, clipped_update!, multi_axpy!, capture, launch,
       destroy, @graph, ubyte_batch!, LazyDArray, lazy, force!, Event, record!, elapsed_ms

const LIB = get(ENV, "AGB200_LIB", joinpath(@__DIR__, "..", "autograd.jl_b200", "libagb200.so"))
const F32, F64 = Int32(0), Int32(1)
dtcode(::Type{Float32}) = F32
dtcode(::Type{Float64}) = F64
const DT = Union{Float32,Float64}

# status codes of include/agb200.h -> the exception the reference would raise
function check(status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:ag_last_error, LIB), Int32, (Ptr{UInt8}, Int64), buf, 1024)
    msg = unsafe_string(pointer(buf))
    status == 1 && (occursin("BoundsError", msg) ? throw(BoundsError()) : throw(DimensionMismatch(msg)))
    status == 2 && throw(OutOfMemoryError())
    status == 5 && throw(ArgumentError("no device method (and no CPU fallback): " * msg))   # MethodError-class
    error("libagb200 status $status: $msg")
end

init(dev::Integer=0) = check(ccall((:ag_init, LIB), Int32, (Int32,), dev))
shutdown() = check(ccall((:ag_shutdown, LIB), Int32, ()))
sync() = check(ccall((:ag_sync, LIB), Int32, ()))          # the `@gs` hook, src/AutoGrad.jl:11
function launch_count()
    r = Ref{Int64}(0)
    check(ccall((:ag_launch_count, LIB), Int32, (Ref{Int64},), r))
    r[]
end

# ---- §1 storage ------------------------------------------------------------------------------------------
mutable struct Buffer
    ptr::Ptr{Cvoid}
    function Buffer(nbytes::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ag_alloc, LIB), Int32, (Int64, Ref{Ptr{Cvoid}}), nbytes, r))
        b = new(r[])
        finalizer(b -> ccall((:ag_free, LIB), Int32, (Ptr{Cvoid},), b.ptr), b)   # stream-ordered, never syncs
        b
    end
end

# Dense column-major device array.  Views (reshape, vec) share the Buffer; `x'` is Julia's own lazy Adjoint wrapper.
struct DArray{T,N} <: AbstractArray{T,N}
    buf::Buffer
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
end
DArray{T}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N} =
    (b = Buffer(max(prod(dims), 1) * sizeof(T)); DArray{T,N}(b, b.ptr, dims))
DArray{T}(::UndefInitializer, dims::Int...) where T = DArray{T}(undef, dims)
function DArray(x::Array{T,N}) where {T<:DT,N}
    a = DArray{T}(undef, size(x))
    isempty(x) || check(ccall((:ag_copy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), a.ptr, x, sizeof(x)))
    sync()                                   # x is pageable: the copy must finish before x may be collected
    a
end
DArray(x::AbstractArray{T}) where {T<:DT} = DArray(Array(x))
function Array(a::DArray{T,N}) where {T,N}
    x = Array{T,N}(undef, a.dims)
    isempty(x) || check(ccall((:ag_copy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), x, a.ptr, sizeof(x)))
    x
end
convert(::Type{Array}, a::DArray) = Array(a)
size(a::DArray) = a.dims
length(a::DArray) = prod(a.dims)
eltype(::DArray{T}) where T = T
similar(a::DArray{T}, dims::Dims=size(a)) where T = DArray{T}(undef, dims)
similar(a::DArray, ::Type{T}, dims::Dims) where {T<:DT} = DArray{T}(undef, dims)
reshape(a::DArray{T}, dims::Dims) where T =
    (prod(dims) == length(a) || throw(DimensionMismatch("reshape $(size(a)) -> $dims")); DArray{T,length(dims)}(a.buf, a.ptr, dims))
reshape(a::DArray, dims::Int...) = reshape(a, dims)
vec(a::DArray) = reshape(a, (length(a),))
function copy(a::DArray{T}) where T                       # src/base.jl:338
    b = similar(a)
    isempty(a) || check(ccall((:ag_copy_d2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), b.ptr, a.ptr, length(a) * sizeof(T)))
    b
end
fill!(a::DArray{T}, v) where T =
    (check(ccall((:ag_fill, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Float64), a.ptr, length(a), dtcode(T), Float64(v))); a)
zero(a::DArray) = fill!(similar(a), 0)
show(io::IO, a::DArray{T}) where T = print(io, "DArray{$T}", size(a))
show(io::IO, ::MIME"text/plain", a::DArray) = show(io, a)
# scalar indexing is a host round trip: allowed for 0-dimensional results only (sum(x) is a device scalar)
getindex(a::DArray{T,0}) where T = Array(a)[]

# Device-resident Int64 index vector, 0-based (the indices of a minibatch stay on the device so that `W[:, ids]` and
# its Sparse gradient need no host synchronisation and a whole recurrent step replays as one CUDA graph).
struct IArray
    buf::Buffer
    n::Int
    lo::Int
    hi::Int
end
function IArray(ids::AbstractVector{<:Integer})            # ids are Julia (1-based) indices
    z = Vector{Int64}(ids .- 1)
    b = Buffer(max(8 * length(z), 8))
    isempty(z) || check(ccall((:ag_copy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), b.ptr, z, 8 * length(z)))
    sync()
    IArray(b, length(z), isempty(z) ? 0 : minimum(z), isempty(z) ? -1 : maximum(z))
end
length(i::IArray) = i.n

# ---- §2 elementwise forward (op ids: include/agb200.h) ------------------------------------------------------
sigm(x::Number) = 1 / (1 + exp(-x))                        # the user primitive of examples/charlm.jl:46-47
relu(x::Number) = max(zero(x), x)
const UNARY = Dict{Symbol,Int32}(
    :- => 1, :abs => 2, :abs2 => 3, :exp => 4, :log => 5, :tanh => 6, :sigm => 7, :sqrt => 8, :sin => 9, :cos => 10,
    :tan => 11, :sinh => 12, :cosh => 13, :asin => 14, :acos => 15, :atan => 16, :asinh => 17, :acosh => 18,
    :atanh => 19, :exp2 => 20, :exp10 => 21, :expm1 => 22, :log2 => 23, :log10 => 24, :log1p => 25, :cbrt => 26,
    :sign => 27, :relu => 28, :one => 29)
# (x read by the rule?, y read by the rule?) per unary op: the table of ewise_ops.cuh / rules.py
const UNARY_NEEDS = Dict{Symbol,Tuple{Bool,Bool}}(
    :- => (false, false), :abs => (true, false), :abs2 => (true, false), :exp => (false, true), :log => (true, false),
    :tanh => (false, true), :sigm => (false, true), :sqrt => (false, true), :sin => (true, false), :cos => (true, false),
    :tan => (false, true), :sinh => (true, false), :cosh => (true, false), :asin => (true, false), :acos => (true, false),
    :atan => (true, false), :asinh => (true, false), :acosh => (true, false), :atanh => (true, false),
    :exp2 => (false, true), :exp10 => (false, true), :expm1 => (false, true), :log2 => (true, false),
    :log10 => (true, false), :log1p => (true, false), :cbrt => (false, true), :relu => (true, false))
const BINARY = Dict{Symbol,Int32}(:+ => 0, :- => 1, :* => 2, :/ => 3, :max => 4, :min => 5, :^ => 6, :(==) => 7,
                                  :< => 8, :<= => 9, :> => 10, :>= => 11, :atan => 12, :hypot => 13)

function unary_fwd(op::Integer, x::DArray{T}) where T
    y = similar(x)
    isempty(x) || check(ccall((:ag_unary_fwd, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                              op, dtcode(T), x.ptr, y.ptr, length(x)))
    y
end
# dx = dy .* g(x,y): ONE fused kernel per rule of src/math.jl (only outside of recording, SURVEY.md 3.4).
# dy: array of x's shape, 0-dimensional device scalar, or a host Number.
function unary_bwd(op::Integer, dy, x::Union{DArray{T},Nothing}, y::Union{DArray{T},Nothing}, ref::DArray{T}) where T
    dx = similar(ref)
    mode, dyp, dys = dy isa Number ? (Int32(2), C_NULL, Float64(dy)) :
                     (length(dy) == 1 && length(ref) != 1) ? (Int32(1), dy.ptr, 0.0) : (Int32(0), dy.ptr, 0.0)
    check(ccall((:ag_unary_bwd, LIB), Int32,
                (Int32, Int32, Ptr{Cvoid}, Int32, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                op, dtcode(T), dyp, mode, dys, x === nothing ? C_NULL : x.ptr, y === nothing ? C_NULL : y.ptr,
                dx.ptr, length(ref)))
    dx
end

# strides of an operand inside the broadcast shape (0 = broadcast), Julia leading-dimension alignment
function bstrides(shape::Dims, out::Dims)
    st, acc = Int64[], 1
    for d in 1:length(out)
        e = d <= length(shape) ? shape[d] : 1
        push!(st, (e == 1 && out[d] != 1) ? 0 : acc); acc *= e
    end
    st
end
# drop extent-1 dims and merge dims every operand traverses contiguously (the C side takes <= 4 collapsed dims)
function collapse(out::Dims, strides::Vector{Vector{Int64}})
    keep = [d for d in 1:length(out) if out[d] != 1]
    isempty(keep) && return Int64[1], [Int64[0] for _ in strides]
    shp = Int64[out[keep[1]]]; sts = [Int64[s[keep[1]]] for s in strides]
    for d in keep[2:end]
        if all(s[d] == st[end] * shp[end] for (s, st) in zip(strides, sts))
            shp[end] *= out[d]
        else
            push!(shp, out[d]); foreach((s, st) -> push!(st, s[d]), strides, sts)
        end
    end
    shp, sts
end
bshape(a) = a isa DArray ? size(a) : ()
function binary_fwd(op::Integer, a, b)
    T = eltype(a isa DArray ? a : b)
    out = Base.Broadcast.broadcast_shape(bshape(a), bshape(b))
    y = DArray{T}(undef, out)
    isempty(y) && return y
    shp, sts = collapse(out, [a isa DArray ? bstrides(size(a), out) : zeros(Int64, length(out)),
                              b isa DArray ? bstrides(size(b), out) : zeros(Int64, length(out)),
                              bstrides(out, out)])
    length(shp) <= 4 || throw(ArgumentError("broadcast needs $(length(shp)) collapsed dims (max 4)"))
    check(ccall((:ag_binary_fwd, LIB), Int32,
                (Int32, Int32, Ptr{Cvoid}, Float64, Ptr{Int64}, Ptr{Cvoid}, Float64, Ptr{Int64}, Ptr{Cvoid}, Int32, Ptr{Int64}),
                op, dtcode(T), a isa DArray ? a.ptr : C_NULL, a isa Number ? Float64(a) : 0.0, sts[1],
                b isa DArray ? b.ptr : C_NULL, b isa Number ? Float64(b) : 0.0, sts[2], y.ptr, length(shp), shp))
    y
end
# dy .* d(op)/d(x_argno) before unbroadcast (ag_binary_bwd); x1 / x2 may be Numbers, y may be nothing
function binary_bwd(op::Integer, argno::Integer, dy::DArray{T}, x1, x2, y) where T
    out = Base.Broadcast.broadcast_shape(size(dy), bshape(x1), bshape(x2), y === nothing ? () : size(y))
    dx = DArray{T}(undef, out)
    isempty(dx) && return dx
    st(v) = v isa DArray ? bstrides(size(v), out) : zeros(Int64, length(out))
    shp, sts = collapse(out, [st(dy), st(x1), st(x2), st(y), bstrides(out, out)])
    length(shp) <= 4 || throw(ArgumentError("broadcast needs $(length(shp)) collapsed dims (max 4)"))
    pv(v) = v isa DArray ? v.ptr : C_NULL
    sv(v) = v isa Number ? Float64(v) : 0.0
    check(ccall((:ag_binary_bwd, LIB), Int32,
                (Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Float64, Ptr{Int64}, Ptr{Cvoid}, Float64, Ptr{Int64},
                 Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Int32, Ptr{Int64}),
                op, argno, dtcode(T), dy.ptr, sts[1], pv(x1), sv(x1), sts[2], pv(x2), sv(x2), sts[3], pv(y), sts[4],
                dx.ptr, length(shp), shp))
    dx
end

# forward methods for the whole tables: what `forw` calls on the unboxed arrays (src/core.jl:66-70)
for (f, id) in UNARY
    f in (:sigm, :relu, :one, :sign) && continue
    @eval broadcasted(::typeof($f), x::DArray) = unary_fwd($id, x)
end
broadcasted(::typeof(sigm), x::DArray) = unary_fwd(UNARY[:sigm], x)
broadcasted(::typeof(relu), x::DArray) = unary_fwd(UNARY[:relu], x)
broadcasted(::typeof(sign), x::DArray) = unary_fwd(UNARY[:sign], x)
broadcasted(::typeof(one), x::DArray) = unary_fwd(UNARY[:one], x)
for (f, id) in BINARY
    @eval broadcasted(::typeof($f), a::DArray, b::Union{DArray,Number}) = binary_fwd($id, a, b)
    @eval broadcasted(::typeof($f), a::Number, b::DArray) = binary_fwd($id, a, b)
end
broadcasted(::typeof(Base.literal_pow), ::typeof(^), a::DArray, ::Val{p}) where p = binary_fwd(BINARY[:^], a, p)
+(a::DArray{T,N}, b::DArray{T,N}) where {T,N} = add(a, b)   # addto!(a,b) = a+b, src/addto.jl:23
-(a::DArray{T,N}, b::DArray{T,N}) where {T,N} = binary_fwd(BINARY[:-], a, b)
-(a::DArray) = unary_fwd(UNARY[:-], a)
*(a::Number, b::DArray) = binary_fwd(BINARY[:*], a, b)
*(a::DArray, b::Number) = binary_fwd(BINARY[:*], a, b)
function add(a::DArray{T}, b::DArray{T}) where T
    size(a) == size(b) || throw(DimensionMismatch("addto!: $(size(a)) vs $(size(b))"))
    out = similar(a)
    isempty(a) || check(ccall((:ag_add, LIB), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                              dtcode(T), a.ptr, b.ptr, out.ptr, length(a)))
    out
end

# user-defined primitives: `@primitive f(x),dy,y (expr)` (src/macros.jl:88-105) lowered to the postfix bytecode of
# ag_ewise_vm.  `code` one Int32 per instruction (op | arg << 8), see include/agb200.h.
function ewise_vm(code::Vector{Int32}, consts::Vector{Float64}, ins::Vector, out::DArray{T}) where T
    ptrs = Ptr{Cvoid}[v isa DArray ? v.ptr : C_NULL for v in ins]
    modes = Int32[v isa Number ? 2 : (length(v) == 1 && length(out) != 1 ? 1 : 0) for v in ins]
    imms = Float64[v isa Number ? v : 0.0 for v in ins]
    check(ccall((:ag_ewise_vm, LIB), Int32,
                (Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Float64}, Ptr{Cvoid}, Int64),
                dtcode(T), code, length(code), consts, length(consts), length(ins), ptrs, modes, imms, out.ptr, length(out)))
    out
end

# ---- §3 device-specialised gradient methods ------------------------------------------------------------------
# What `@primitive f(x),dy,y (dy .* g)` emits for Value arguments (src/macros.jl:100-103) is a `back` method whose
# body is the broadcast expression; for a Value holding a DArray the same signature gets a body that is ONE fused
# kernel.  Under an outer tape (recording() true, higher-order gradients) the reference's own expression must run so
# that it records itself: the methods fall through to it with `invoke`.
const VD = Value{<:DArray}
for (f, id) in UNARY
    haskey(UNARY_NEEDS, f) || continue
    nx, ny = UNARY_NEEDS[f]
    fn = f in (:sigm, :relu) ? f : f
    @eval function back(::typeof(broadcasted), ::Type{Arg{2}}, dy, y, ::typeof($fn), x::VD)
        recording() && return invoke(back, Tuple{typeof(broadcasted),Type{Arg{2}},Any,Any,typeof($fn),Value}, broadcasted, Arg{2}, dy, y, $fn, x)
        xv, yv, dyv = value(x), value(y), value(dy)
        unary_bwd($id, dyv, $nx ? xv : nothing, $ny ? yv : nothing, xv)
    end
end
# the user primitives of the examples, defined here so that they have device rules out of the box
AutoGrad.@primitive sigm(x),dy,y (dy .* y .* (1 .- y))        # examples/charlm.jl:46-47
AutoGrad.@primitive relu(x),dy,y (dy .* (y .== x))            # max.(0,x): src/math.jl:54 with x1 = 0
# special functions (src/specialfunctions.jl): forward methods only -- the reference's own `@primitive` rules
# (:5-8,20-21,25,29-30,33-44,59,63-67) are broadcast expressions over these same functions, so they run on the device
# through the methods below.  Ids 30-42 are in the unary table (CUDA's math library has them); the rest run as their own
# pass (ag_special_param: integer-order Bessel functions, polygamma(m, x), invdigamma, dawson, erfi, Airy functions).
import SpecialFunctions
const SPECIAL_UNARY = Dict{Symbol,Int32}(
    :erf => 30, :erfc => 31, :erfcx => 32, :erfinv => 33, :erfcinv => 34, :gamma => 35, :loggamma => 36, :digamma => 37,
    :trigamma => 38, :besselj0 => 39, :besselj1 => 40, :bessely0 => 41, :bessely1 => 42)
const SPECIAL_PARAM = Dict{Symbol,Int32}(
    :besselj => 0, :bessely => 1, :polygamma => 2, :invdigamma => 3, :dawson => 4, :erfi => 5, :airyai => 6,
    :airyaiprime => 7, :airybi => 8, :airybiprime => 9, :besseli => 10, :besselk => 11)
function special_param(op::Integer, m::Integer, x::DArray{T}) where T
    y = similar(x)
    isempty(x) || check(ccall((:ag_special_param, LIB), Int32, (Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                              op, dtcode(T), m, x.ptr, y.ptr, length(x)))
    y
end
for (f, id) in SPECIAL_UNARY
    isdefined(SpecialFunctions, f) || continue
    @eval broadcasted(::typeof(SpecialFunctions.$f), x::DArray) = unary_fwd($id, x)
end
for f in (:invdigamma, :dawson, :erfi, :airyai, :airyaiprime, :airybi, :airybiprime)
    @eval broadcasted(::typeof(SpecialFunctions.$f), x::DArray) = special_param($(SPECIAL_PARAM[f]), 0, x)
end
for f in (:besselj, :bessely, :polygamma, :besseli, :besselk)      # integer parameter only: other orders have no device function
    @eval broadcasted(::typeof(SpecialFunctions.$f), m::Integer, x::DArray) = special_param($(SPECIAL_PARAM[f]), m, x)
end
# the rest of src/math.jl: forward methods composed from the table's ops (the reference's rules for them, src/math.jl:10-73,
# are broadcast expressions over these same functions and run on the device through the methods here); the @zerograd
# rounding functions of src/base.jl:79,100,145,151 are ag_special_param ops 12-15
const D180, R180 = 180 / pi, pi / 180
rcp(x::DArray) = binary_fwd(BINARY[:/], 1.0, x)
scal(x::DArray, c) = binary_fwd(BINARY[:*], x, Float64(c))
for (f, inner) in ((:acot, :atan), (:acoth, :atanh), (:acsc, :asin), (:acsch, :asinh), (:asec, :acos), (:asech, :acosh))
    @eval broadcasted(::typeof($f), x::DArray) = unary_fwd($(UNARY[inner]), rcp(x))
end
for (f, inner) in ((:acotd, :atan), (:acscd, :asin), (:asecd, :acos))
    @eval broadcasted(::typeof($f), x::DArray) = scal(unary_fwd($(UNARY[inner]), rcp(x)), D180)
end
for (f, inner) in ((:acosd, :acos), (:asind, :asin), (:atand, :atan))
    @eval broadcasted(::typeof($f), x::DArray) = scal(unary_fwd($(UNARY[inner]), x), D180)
end
for (f, inner, c) in ((:sind, :sin, R180), (:cosd, :cos, R180), (:tand, :tan, R180), (:sinpi, :sin, pi), (:cospi, :cos, pi))
    @eval broadcasted(::typeof($f), x::DArray) = unary_fwd($(UNARY[inner]), scal(x, $c))
end
for (f, inner) in ((:sec, :cos), (:csc, :sin), (:cot, :tan), (:sech, :cosh), (:csch, :sinh), (:coth, :tanh))
    @eval broadcasted(::typeof($f), x::DArray) = rcp(unary_fwd($(UNARY[inner]), x))
end
for (f, inner) in ((:secd, :cos), (:cscd, :sin), (:cotd, :tan))
    @eval broadcasted(::typeof($f), x::DArray) = rcp(unary_fwd($(UNARY[inner]), scal(x, R180)))
end
broadcasted(::typeof(deg2rad), x::DArray) = scal(x, R180)
broadcasted(::typeof(rad2deg), x::DArray) = scal(x, D180)
function broadcasted(::typeof(sinc), x::DArray)            # sinc(0) = 1 without a branch
    e = binary_fwd(BINARY[:(==)], x, 0.0)
    px = scal(x, pi)
    add(binary_fwd(BINARY[:/], unary_fwd(UNARY[:sin], px), add(px, e)), e)
end
function broadcasted(::typeof(cosc), x::DArray)            # cosc(0) = 0
    e = binary_fwd(BINARY[:(==)], x, 0.0)
    binary_fwd(BINARY[:/], binary_fwd(BINARY[:-], unary_fwd(UNARY[:cos], scal(x, pi)), broadcasted(sinc, x)), add(x, e))
end
broadcasted(::typeof(clamp), x::DArray, lo::Number, hi::Number) = binary_fwd(BINARY[:min], binary_fwd(BINARY[:max], x, lo), hi)
broadcasted(::typeof(ldexp), x::DArray, n::Integer) = scal(x, 2.0^n)
broadcasted(::typeof(log), b::Number, x::DArray) = scal(unary_fwd(UNARY[:log], x), 1 / log(b))
for (f, id) in ((:floor, 12), (:ceil, 13), (:round, 14), (:trunc, 15))
    @eval broadcasted(::typeof($f), x::DArray) = special_param($id, 0, x)
end
broadcasted(::typeof(exponent), x::DArray) = special_param(16, 0, x)
broadcasted(::typeof(significand), x::DArray) = special_param(17, 0, x)
broadcasted(::typeof(mod2pi), x::DArray) = special_param(18, 2, x)
rmode(::RoundingMode{:Nearest}) = 0; rmode(::RoundingMode{:ToZero}) = 1; rmode(::RoundingMode{:Down}) = 2; rmode(::RoundingMode{:Up}) = 3
broadcasted(::typeof(rem2pi), x::DArray, r::RoundingMode) = special_param(18, rmode(r), x)
# det / inv / logdet / logabsdet (src/linearalgebra.jl:18,28,57-58): forward methods through ag_inv_det (one launch);
# the reference's rules (dy*y*inv(x)', -y'*dy*y', dy*inv(x)') are matrix products over these
import LinearAlgebra
function inv_det(x::DArray{T,2}, want_inv::Bool) where T
    n = size(x, 1)
    n == size(x, 2) || throw(DimensionMismatch("matrix is not square: dimensions are $(size(x))"))
    iv = want_inv ? similar(x) : nothing
    info = DArray{T}(undef, (3,))
    check(ccall((:ag_inv_det, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                dtcode(T), x.ptr, n, want_inv ? iv.ptr : C_NULL, info.ptr))
    iv, Array(info)                                   # (inv or nothing, [det, log|det|, sign])
end
LinearAlgebra.inv(x::DArray{T,2}) where T = inv_det(x, true)[1]
LinearAlgebra.det(x::DArray{T,2}) where T = inv_det(x, false)[2][1]
LinearAlgebra.logabsdet(x::DArray{T,2}) where T = (d = inv_det(x, false)[2]; (d[2], d[3]))
function LinearAlgebra.logdet(x::DArray{T,2}) where T
    d = inv_det(x, false)[2]
    d[3] < 0 && throw(DomainError(d[1], "determinant is negative"))
    d[2]
end
# argmax / argmin (@zerograd, src/base.jl:60-61): first position of the extremum, 1-based linear index on this side
function argext(x::DArray{T}, want_max::Bool) where T
    isempty(x) && throw(ArgumentError("collection must be non-empty"))
    out = Buffer(8)
    check(ccall((:ag_argext, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}), dtcode(T), x.ptr, length(x), want_max ? 1 : 0, out.ptr))
    host = Ref{Int64}(0)
    check(ccall((:ag_copy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), host, out.ptr, 8))
    sync()
    host[] + 1
end
Base.argmax(x::DArray) = argext(x, true)
Base.argmin(x::DArray) = argext(x, false)
# eltype conversion on the device (oftype / big / Float64.(x), src/base.jl:76,365)
function cast(::Type{D}, x::DArray{S}) where {D<:DT,S}
    D === S && return x
    y = DArray{D}(undef, size(x))
    isempty(x) || check(ccall((:ag_cast, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64), dtcode(S), dtcode(D), x.ptr, y.ptr, length(x)))
    y
end
Base.oftype(t::DArray{D}, x::DArray) where D = cast(D, x)
Base.big(x::DArray) = cast(Float64, x)
broadcasted(::Type{D}, x::DArray) where {D<:DT} = cast(D, x)
# binary broadcasting rules: fused kernel + unbroadcast (src/base.jl:21-49, src/math.jl:24,48,54-55)
for (f, id) in BINARY
    f in (:(==), :<, :<=, :>, :>=) && continue                # @zerograd (src/base.jl:52-60)
    for argno in (1, 2)
        A = argno == 1 ? :(Arg{2}) : :(Arg{3})
        @eval function back(::typeof(broadcasted), ::Type{$A}, dy, y, ::typeof($f), x1::Union{VD,DArray,Number}, x2::Union{VD,DArray,Number})
            (x1 isa VD || x2 isa VD) || throw(MethodError(back, (broadcasted, $A, dy, y, $f, x1, x2)))
            xi = $argno == 1 ? x1 : x2
            if recording() || !(value(dy) isa DArray)
                return invoke(back, Tuple{typeof(broadcasted),Type{$A},Any,Any,typeof($f),Any,Any}, broadcasted, $A, dy, y, $f, x1, x2)
            end
            unbroadcast(xi, binary_bwd($id, $argno, value(dy), value(x1), value(x2), value(y)))
        end
    end
end

# ---- §4 reductions ---------------------------------------------------------------------------------------------
const RMAP = Dict{Any,Int32}(identity => 0, abs2 => 1, abs => 2)
function sum_all(map::Integer, x::DArray{T}) where T
    out = DArray{T}(undef, ())
    isempty(x) ? fill!(out, 0) :
        check(ccall((:ag_sum_all, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}), map, dtcode(T), x.ptr, length(x), out.ptr))
    out
end
# x viewed as (inner, red, outer): one call per maximal run of consecutive reduced dims
function sum_dims(map::Integer, x::DArray{T}, dims) where T
    ds = sort(unique(Int[d for d in (dims isa Integer ? (dims,) : dims) if d <= ndims(x)]))
    cur, cs, m = x, collect(size(x)), map
    i = 1
    while i <= length(ds)
        j = i
        while j < length(ds) && ds[j+1] == ds[j] + 1; j += 1; end
        lo, hi = ds[i], ds[j]
        inner, red, outer = prod(cs[1:lo-1]), prod(cs[lo:hi]), prod(cs[hi+1:end])
        ns = copy(cs); ns[lo:hi] .= 1
        out = DArray{T}(undef, Tuple(ns))
        isempty(out) || check(ccall((:ag_sum_dim, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Cvoid}),
                                    m, dtcode(T), cur.ptr, inner, red, outer, out.ptr))
        cur, cs, m = out, ns, Int32(0)
        i = j + 1
    end
    cur
end
# the reference returns a host Number from sum(x) (SURVEY.md 7.3 item 6); `device_scalar = true` keeps it on the
# device (0-dimensional DArray) so that a whole step can be enqueued / captured without a host synchronisation
const DEVICE_SCALARS = Ref(false)
scalar(out::DArray{T,0}) where T = DEVICE_SCALARS[] ? out : Array(out)[]
sum(x::DArray; dims=:) = dims === Colon() ? scalar(sum_all(0, x)) : sum_dims(0, x, dims)
sum(f::Union{typeof(abs2),typeof(abs)}, x::DArray; dims=:) =
    dims === Colon() ? scalar(sum_all(RMAP[f], x)) : sum_dims(RMAP[f], x, dims)
function reduce_dims(op::Integer, x::DArray{T}, dims) where T
    ds = dims === Colon() ? collect(1:ndims(x)) : sort(unique(Int[d for d in (dims isa Integer ? (dims,) : dims) if d <= ndims(x)]))
    cur, cs = x, collect(size(x))
    i = 1
    while i <= length(ds)
        j = i
        while j < length(ds) && ds[j+1] == ds[j] + 1; j += 1; end
        lo, hi = ds[i], ds[j]
        ns = copy(cs); ns[lo:hi] .= 1
        out = DArray{T}(undef, Tuple(ns))
        isempty(out) || check(ccall((:ag_reduce_dim, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Cvoid}),
                                    op, dtcode(T), cur.ptr, prod(cs[1:lo-1]), prod(cs[lo:hi]), prod(cs[hi+1:end]), out.ptr))
        cur, cs = out, ns
        i = j + 1
    end
    dims === Colon() ? scalar(reshape(cur, ())) : cur
end
maximum(x::DArray; dims=:) = reduce_dims(0, x, dims)       # src/base.jl:202; rule dy.*(y.==x) via binary_bwd
minimum(x::DArray; dims=:) = reduce_dims(1, x, dims)       # src/base.jl:205
prod(x::DArray; dims=:) = reduce_dims(2, x, dims)          # src/base.jl:221

# ---- §5 matmul: forward `*` and both gradient products of src/base.jl:26 -------------------------------------------
const Mat{T} = Union{DArray{T,2},Adjoint{T,DArray{T,2}},Transpose{T,DArray{T,2}}}
storage(a::DArray) = (a, Int32(0))
storage(a::Union{Adjoint,Transpose}) = (parent(a), Int32(1))
function *(a::Mat{T}, b::Mat{T}) where T<:DT
    m, k = size(a); k2, n = size(b)
    k == k2 || throw(DimensionMismatch("A has dimensions ($m,$k) but B has dimensions ($k2,$n)"))
    (sa, ta), (sb, tb) = storage(a), storage(b)
    c = DArray{T}(undef, (m, n))
    (m == 0 || n == 0) && return c
    k == 0 && return fill!(c, 0)
    check(ccall((:ag_gemm, LIB), Int32,
                (Int32, Int32, Int32, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}, Int64),
                dtcode(T), ta, tb, m, n, k, 1.0, sa.ptr, size(sa, 1), sb.ptr, size(sb, 1), 0.0, c.ptr, m))
    c
end
*(a::Mat{T}, b::DArray{T,1}) where T<:DT = vec(a * reshape(b, (length(b), 1)))
# `act.(w*x .+ b)` as ONE launch: bias and activation in the product's epilogue (ag_gemm_epilogue).  act in
# (identity, relu, tanh, sigm); keep_z = true also returns the pre-activation sum (relu's rule reads it).
const ACTS = Dict{Any,Int32}(identity => 0, relu => 28, tanh => 6, sigm => 7)
function gemm_bias_act(w::Mat{T}, x::Mat{T}, b::Union{DArray{T},Nothing}, act=identity; keep_z::Bool=false) where T<:DT
    m, k = size(w); k2, n = size(x)
    k == k2 || throw(DimensionMismatch("A has dimensions ($m,$k) but B has dimensions ($k2,$n)"))
    (b === nothing || length(b) == m) || throw(DimensionMismatch("bias of length $(length(b)) for $m rows"))
    (sa, ta), (sb, tb) = storage(w), storage(x)
    c = DArray{T}(undef, (m, n))
    c2 = (keep_z && act !== identity) ? DArray{T}(undef, (m, n)) : nothing
    check(ccall((:ag_gemm_epilogue, LIB), Int32,
                (Int32, Int32, Int32, Int64, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32,
                 Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                dtcode(T), ta, tb, m, n, k, sa.ptr, size(sa, 1), sb.ptr, size(sb, 1), b === nothing ? C_NULL : b.ptr,
                ACTS[act], c.ptr, c2 === nothing ? C_NULL : c2.ptr, m))
    c2 === nothing ? c : (c, c2)
end
function gemm_runs_fused(w::Mat{T}, x::Mat{T}) where T<:DT
    m, k = size(w); n = size(x, 2)
    (sa, ta), (sb, tb) = storage(w), storage(x)
    r = Ref{Int32}(0)
    check(ccall((:ag_gemm_epilogue_query, LIB), Int32,
                (Int32, Int32, Int32, Int64, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ref{Int32}),
                dtcode(T), ta, tb, m, n, k, sa.ptr, size(sa, 1), sb.ptr, size(sb, 1), r))
    r[] != 0
end
# Several independent products handed over together (the four gate products of an LSTM step, examples/charlm.jl:69-73
# "we can probably combine these four operations into one", their dx products, the dW products of several steps):
# the library runs those of one class as ONE launch.  GemmDesc mirrors `ag_gemm_desc` of include/agb200.h field by field.
struct GemmDesc
    transA::Int32
    transB::Int32
    M::Int64
    N::Int64
    K::Int64
    A::Ptr{Cvoid}
    lda::Int64
    B::Ptr{Cvoid}
    ldb::Int64
    bias::Ptr{Cvoid}
    act::Int32
    C::Ptr{Cvoid}
    C2::Ptr{Cvoid}
    ldc::Int64
end
# jobs: (w, x, b or nothing, act) tuples; returns the outputs in order
function gemm_group(jobs::Vector{<:Tuple}, ::Type{T}=Float32) where T<:DT
    descs, outs = GemmDesc[], DArray{T}[]
    for (w, x, b, act) in jobs
        m, k = size(w); k2, n = size(x)
        k == k2 || throw(DimensionMismatch("A has dimensions ($m,$k) but B has dimensions ($k2,$n)"))
        (sa, ta), (sb, tb) = storage(w), storage(x)
        c = DArray{T}(undef, (m, n))
        push!(outs, c)
        push!(descs, GemmDesc(ta, tb, m, n, k, sa.ptr, size(sa, 1), sb.ptr, size(sb, 1),
                              b === nothing ? C_NULL : b.ptr, ACTS[act], c.ptr, C_NULL, m))
    end
    check(ccall((:ag_gemm_group, LIB), Int32, (Int32, Int32, Ptr{GemmDesc}), dtcode(T), length(descs), descs))
    outs
end
function gemm_engine()
    buf = Vector{UInt8}(undef, 32)
    check(ccall((:ag_gemm_last_engine, LIB), Int32, (Ptr{UInt8}, Int64), buf, 32))
    unsafe_string(pointer(buf))
end
set_gemm_engine(e::Integer) = check(ccall((:ag_gemm_set_engine, LIB), Int32, (Int32,), e))

# ---- §6 getindex / view / cat / permutedims ------------------------------------------------------------------------
# W[:, ids] (the embedding lookup of examples/charlm.jl:99): blocked gather of whole columns
function getindex(x::DArray{T,2}, ::Colon, ids::IArray) where T
    (ids.n == 0 || (ids.lo >= 0 && ids.hi < size(x, 2))) || throw(BoundsError(x, (:, ids)))
    out = DArray{T}(undef, (size(x, 1), ids.n))
    isempty(out) || check(ccall((:ag_gather_cols, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Int64),
                                dtcode(T), x.ptr, size(x, 1), ids.buf.ptr, out.ptr, ids.n))
    out
end
getindex(x::DArray{T,2}, c::Colon, ids::AbstractVector{<:Integer}) where T = getindex(x, c, IArray(ids))
# general indexing: the host normalises the index tuple to flat column-major destinations (bit-exact bookkeeping,
# src/addto.jl:95 `to_indices`), the device gathers
function flat_dest(shape::Dims, I...)
    idx = to_indices(zeros(Bool, ntuple(_ -> 0, length(shape))...), axes_of(shape), I)
    lin = LinearIndices(shape)
    vec(Int64[lin[i...] - 1 for i in Iterators.product(idx...)]), map(length, filter(i -> !(i isa Integer), collect(idx)))
end
axes_of(shape::Dims) = map(Base.OneTo, shape)
function getindex(x::DArray{T}, I...) where T
    dest, oshape = flat_dest(size(x), I...)
    out = DArray{T}(undef, Tuple(oshape))
    ib = IArray(dest .+ 1)
    isempty(out) || check(ccall((:ag_gather, LIB), Int32, (Int32, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Int64),
                                dtcode(T), x.ptr, ib.buf.ptr, out.ptr, length(out)))
    out
end
# a contiguous range of columns is a view that shares the Buffer (src/getindex.jl:35-37 `view`)
function view(x::DArray{T,2}, ::Colon, r::UnitRange{Int}) where T
    (first(r) >= 1 && last(r) <= size(x, 2)) || throw(BoundsError(x, (:, r)))
    DArray{T,2}(x.buf, x.ptr + (first(r) - 1) * size(x, 1) * sizeof(T), (size(x, 1), length(r)))
end
# slab copy: dst[i, doff + r, o] = src[i, soff + r, o] (cat forward / uncat gradient slices, src/cat.jl:27-67)
function slab_copy!(dst::DArray{T}, dred, doff, src::DArray{T}, sred, soff, inner, len, outer) where T
    check(ccall((:ag_slab_copy, LIB), Int32,
                (Int32, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64),
                dtcode(T), src.ptr, sred, soff, dst.ptr, dred, doff, inner, len, outer))
    dst
end
function cat(xs::DArray{T}...; dims) where T
    d = dims isa Val ? typeof(dims).parameters[1] : dims
    nd = max(maximum(ndims, xs), d)
    shp(x) = ntuple(i -> size(x, i), nd)
    for x in xs, i in 1:nd
        (i == d || size(x, i) == size(xs[1], i)) || throw(DimensionMismatch("cat: mismatch in dimension $i"))
    end
    tot = sum(size(x, d) for x in xs)
    oshape = ntuple(i -> i == d ? tot : size(xs[1], i), nd)
    out = DArray{T}(undef, oshape)
    inner, outer, off = prod(oshape[1:d-1]), prod(oshape[d+1:end]), 0
    if length(xs) <= 8 && !isempty(out)                     # one launch for all sources (ag_cat)
        srcs = Ptr{Cvoid}[size(x, d) > 0 ? x.ptr : C_NULL for x in xs]
        lens = Int64[size(x, d) for x in xs]
        check(ccall((:ag_cat, LIB), Int32, (Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cvoid}, Int64, Int64),
                    dtcode(T), length(xs), srcs, lens, out.ptr, inner, outer))
        return out
    end
    for x in xs
        size(x, d) > 0 && slab_copy!(out, tot, off, x, size(x, d), 0, inner, size(x, d), outer)
        off += size(x, d)
    end
    out
end
vcat(xs::DArray{T}...) where T = cat(xs...; dims=1)       # src/cat.jl:122
hcat(xs::DArray{T}...) where T = cat(xs...; dims=2)       # src/cat.jl:123
# the slice of dy that belongs to argument argn of a cat (what `uncat` computes, src/cat.jl:46-67)
function uncat_slice(dy::DArray{T}, lo::Int, len::Int, d::Int, target::Dims) where T
    shape = ntuple(i -> size(dy, i), max(ndims(dy), d))
    out = DArray{T}(undef, target)
    isempty(out) || slab_copy!(out, len, 0, dy, shape[d], lo, prod(shape[1:d-1]), len, prod(shape[d+1:end]))
    out
end
function permutedims(x::DArray{T,N}, perm) where {T,N}      # src/base.jl:217-218
    p = collect(Int32, perm) .- Int32(1)
    sort(p) == collect(Int32(0):Int32(N - 1)) || throw(ArgumentError("$perm is not a permutation of $N dims"))
    out = DArray{T}(undef, ntuple(i -> size(x, perm[i]), N))
    isempty(out) || check(ccall((:ag_permute, LIB), Int32, (Int32, Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}),
                                dtcode(T), x.ptr, N, collect(Int64, size(x)), p, out.ptr))
    out
end
function materialize(a::Union{Adjoint{T,DArray{T,2}},Transpose{T,DArray{T,2}}}) where T   # permutedims(x, (2,1))
    x = parent(a)
    out = DArray{T}(undef, (size(x, 2), size(x, 1)))
    isempty(out) || check(ccall((:ag_transpose, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                                dtcode(T), x.ptr, size(x, 1), size(x, 2), out.ptr))
    out
end

# ---- §7 Sparse on a device container ----------------------------------------------------------------------------------
# All (index, value) blocks of a Sparse are flattened into ONE destination vector (first index fastest, blocks in
# append order: the order of the reference's sequential loop, src/addto.jl:91-98,107-129) and applied by one
# deterministic sort + segmented reduce.  The embedding case (Colon(), ids) is block-granular: one destination per
# column, `rows` values each.
zeroslike(a::DArray) = zero(a)
function concat_flat(vals::Vector{<:DArray{T}}) where T
    length(vals) == 1 && return vec(vals[1])
    tot = sum(length, vals)
    out = DArray{T}(undef, (tot,))
    off = 0
    for v in vals
        isempty(v) || slab_copy!(out, tot, off, vec(v), length(v), 0, 1, length(v), 1)
        off += length(v)
    end
    out
end
function concat_index(blocks::Vector{IArray})
    length(blocks) == 1 && return blocks[1]
    n = sum(b -> b.n, blocks)
    buf = Buffer(max(8n, 8))
    off = 0
    for b in blocks                                       # raw 8-byte copies through the Float64 slab kernel
        b.n > 0 && check(ccall((:ag_slab_copy, LIB), Int32,
                               (Int32, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64),
                               F64, b.buf.ptr, b.n, 0, buf.ptr, n, off, 1, b.n, 1))
        off += b.n
    end
    IArray(buf, n, minimum(b -> b.lo, blocks), maximum(b -> b.hi, blocks))
end
function scatter_add!(dest::DArray{T}, idx::IArray, vals::DArray{T}, block::Integer) where T
    idx.n == 0 && return dest
    length(vals) == idx.n * block || throw(DimensionMismatch("scatter-add: $(length(vals)) values for $(idx.n) destinations"))
    (idx.lo >= 0 && (idx.hi + 1) * block <= length(dest)) || throw(BoundsError())
    check(ccall((:ag_scatter_add, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Int64, Int64),
                dtcode(T), dest.ptr, length(dest), idx.buf.ptr, vals.ptr, idx.n, block))
    dest
end
iscolblock(a::DArray, idx) = ndims(a) == 2 && length(idx) == 2 && idx[1] isa Colon && (idx[2] isa IArray || idx[2] isa AbstractVector{<:Integer})
function addto!(a::DArray{T}, b::Sparse) where T
    size(a) == size(b.container) || throw(DimensionMismatch("addto!: accumulator does not match the Sparse container"))
    recording() && (a = copy(a))                          # do not overwrite in a higher-order context (src/addto.jl:93)
    isempty(b.values) && return a
    vals = DArray{T}[value(v) for v in b.values]
    if all(iscolblock(a, i) for i in b.indices)
        idx = concat_index(IArray[i[2] isa IArray ? i[2] : IArray(i[2]) for i in b.indices])
        scatter_add!(a, idx, concat_flat(vals), size(a, 1))
    else
        flat = IArray[IArray(flat_dest(size(a), i...)[1] .+ 1) for i in b.indices]
        scatter_add!(a, concat_index(flat), concat_flat(vals), 1)
    end
    a
end
addto!(a::Sparse, b::DArray) = addto!(b, a)
full(b::Sparse{T,N}) where {T,N} = b.container isa DArray ? addto!(zeroslike(b.container), b) : invoke(full, Tuple{Any}, b)
axpy!(a::Number, x::Sparse, y::DArray) = addto!(y, lmul!(a, x))                              # src/sparse.jl:50
axpy!(a::Number, x::Sparse, y::Param) = (axpy!(a, x, value(y)); y)
# lmul!(a, ::Sparse) of the reference builds `a*v` per block (src/sparse.jl:51): `*(::Number, ::DArray)` above
norm(x::DArray) = sqrt(Float64(Array(sum_all(1, x))[]))
norm(x::Sparse) = sqrt(sum(abs2, norm(v) for v in x.values))                                   # src/sparse.jl:53-54

# ---- §8 update --------------------------------------------------------------------------------------------------------
axpy!(alpha::Number, x::DArray{T}, y::DArray{T}) where T =
    (check(ccall((:ag_axpy, LIB), Int32, (Int32, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Int64), dtcode(T), Float64(alpha), x.ptr, y.ptr, length(y))); y)
axpy!(alpha::Number, x::DArray, y::Param) = (axpy!(alpha, x, value(y)); y)
lmul!(alpha::Number, x::DArray{T}) where T =
    (check(ccall((:ag_scale, LIB), Int32, (Int32, Float64, Ptr{Cvoid}, Int64), dtcode(T), Float64(alpha), x.ptr, length(x))); x)
ptrs(xs) = Ptr{Cvoid}[x.ptr for x in xs]
# `for w in params(f); axpy!(-lr, grad(J,w), w); end` (README.md:73-76) as ONE launch (<= 16 tensors)
function multi_axpy!(alpha::Number, gs::Vector{<:DArray{T}}, ws::Vector{<:DArray{T}}) where T
    check(ccall((:ag_multi_axpy, LIB), Int32, (Int32, Int32, Float64, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}),
                dtcode(T), length(gs), Float64(alpha), ptrs(gs), ptrs(ws), Int64[length(g) for g in gs]))
    ws
end
# gradient clipping on the device (examples/charlm.jl:116-118,148-152): gnorm = sqrt(sum_k sum(abs2, g_k)); the
# step is scaled by min(1, gclip/gnorm) read from device memory -- no host round trip, replays as a graph
function clipped_update!(alpha::Number, gs::Vector{<:DArray{T}}, ws::Vector{<:DArray{T}}, gclip::Number) where T
    sizes = Int64[length(g) for g in gs]
    nrm2 = DArray{T}(undef, ())
    check(ccall((:ag_multi_sumabs2, LIB), Int32, (Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cvoid}),
                dtcode(T), length(gs), ptrs(gs), sizes, nrm2.ptr))
    scale = binary_fwd(BINARY[:min], binary_fwd(BINARY[:/], Float64(gclip), unary_fwd(UNARY[:sqrt], reshape(nrm2, (1,)))), 1.0)
    check(ccall((:ag_multi_axpy_scaled, LIB), Int32,
                (Int32, Int32, Float64, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}),
                dtcode(T), length(gs), Float64(alpha), scale.ptr, ptrs(gs), ptrs(ws), sizes))
    ws
end

# ---- §9 deferred elementwise chains (horizontal fusion; executable logic: autograd.jl_b200/lazy.py) ----------------------
# A LazyDArray is an expression node over DArrays of one shape; `force!` compiles the DAG to the bytecode of
# ag_ewise_fused (include/agb200.h) and runs it as ONE launch: each leaf read once, each requested value written once.
# This reduced form covers full-shape operands, column / row vectors and immediates (modes 0, 2, 3, 4).
mutable struct LazyDArray{T}
    shape::Dims
    kind::Symbol                  # :leaf, :unary, :binary
    op::Int32
    args::Vector{Any}
    store::Union{DArray{T},Nothing}
end
lazy(x::DArray{T}) where T = LazyDArray{T}(size(x), :leaf, Int32(0), Any[x], x)
lazy(f::Symbol, a::LazyDArray{T}) where T = LazyDArray{T}(a.shape, :unary, UNARY[f], Any[a], nothing)
lazy(f::Symbol, a::LazyDArray{T}, b::Union{LazyDArray{T},DArray{T},Number}) where T =
    LazyDArray{T}(a.shape, :binary, BINARY[f], Any[a, b], nothing)
const LD_IN, LD_CONST, OP_UNARY, OP_BIN, OP_BIN_IN, OP_BIN_CONST, OP_STORE = Int32(0), Int32(1), Int32(4), Int32(5), Int32(6), Int32(7), Int32(10)
function force!(root::LazyDArray{T}) where T
    root.store !== nothing && return root.store
    code, consts, ins, modes = Int32[], Float64[], Ptr{Cvoid}[], Int32[]
    rows, cols = length(root.shape) >= 1 ? root.shape[1] : 1, prod(root.shape[2:end])
    function leaf!(x::DArray)
        i = findfirst(==(x.ptr), ins)
        i === nothing || return Int32(i - 1)
        push!(ins, x.ptr)
        push!(modes, size(x) == root.shape ? 0 : (length(x) == rows && size(x, 1) == rows) ? 3 : length(x) == cols ? 4 :
                     throw(ArgumentError("operand layout $(size(x)) in a chain of shape $(root.shape)")))
        Int32(length(ins) - 1)
    end
    function const!(v)
        push!(consts, Float64(v)); Int32(length(consts) - 1)
    end
    function gen!(n::LazyDArray, push::Bool)
        if n.store !== nothing
            Base.push!(code, LD_IN | (leaf!(n.store) << 8) | (Int32(push) << 16))
        elseif n.kind == :unary
            gen!(n.args[1], push); Base.push!(code, OP_UNARY | (n.op << 8))
        else
            a, b = n.args
            gen!(a, push)
            if b isa Number
                Base.push!(code, OP_BIN_CONST | (n.op << 8) | (const!(b) << 16))
            elseif b isa DArray || b.store !== nothing
                Base.push!(code, OP_BIN_IN | (n.op << 8) | (leaf!(b isa DArray ? b : b.store) << 16))
            else
                gen!(b, true); Base.push!(code, OP_BIN | (n.op << 8))
            end
        end
    end
    gen!(root, false)
    push!(code, OP_STORE)
    out = DArray{T}(undef, root.shape)
    ewise_fused!([out], code, consts, ins, modes, zeros(Float64, length(ins)), rows, cols)
    root.store = out
end
function ewise_fused!(outs::Vector{<:DArray{T}}, code::Vector{Int32}, consts::Vector{Float64},
                      ins::Vector{Ptr{Cvoid}}, modes::Vector{Int32}, imms::Vector{Float64}, rows::Integer, cols::Integer) where T
    op = Ptr{Cvoid}[o.ptr for o in outs]
    check(ccall((:ag_ewise_fused, LIB), Int32,
                (Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Float64}, Int32,
                 Ptr{Ptr{Cvoid}}, Int64, Int64),
                dtcode(T), code, length(code), consts, length(consts), length(ins), ins, modes, imms, length(outs), op,
                rows, cols))
    outs
end
# ... and the variant that also sums the chain (sum(x; dims) / sum(abs2, x) of a pending x, src/unbroadcast.jl:17-24):
# same arguments plus the (inner, red, outer) view of the operand and the result array; status 5 means "not fusable
# here" and the caller forces x and calls ag_sum_dim.
function ewise_fused_reduce!(r::DArray{T}, outs::Vector{<:DArray{T}}, code::Vector{Int32}, consts::Vector{Float64},
                             ins::Vector{Ptr{Cvoid}}, modes::Vector{Int32}, imms::Vector{Float64}, rows::Integer,
                             cols::Integer, inner::Integer, red::Integer, outer::Integer) where T
    op = Ptr{Cvoid}[o.ptr for o in outs]
    ccall((:ag_ewise_fused_reduce, LIB), Int32,
          (Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Float64}, Int32,
           Ptr{Ptr{Cvoid}}, Int64, Int64, Int64, Int64, Int64, Ptr{Cvoid}),
          dtcode(T), code, length(code), consts, length(consts), length(ins), ins, modes, imms, length(outs), op,
          rows, cols, inner, red, outer, r.ptr)
end

# Column programs: a chain that sums over the rows of an m x n matrix and uses the sums again (the softmax normaliser
# of examples/mnist.jl:40 and the unbroadcast of its gradient) as one launch; see include/agb200.h for the segment
# instructions.  r === nothing: no whole-space sum.
function ewise_colprog!(r::Union{DArray{T},Nothing}, outs::Vector{<:DArray{T}}, code::Vector{Int32}, consts::Vector{Float64},
                        ins::Vector{Ptr{Cvoid}}, modes::Vector{Int32}, m::Integer, n::Integer) where T
    op = Ptr{Cvoid}[o.ptr for o in outs]
    imms = zeros(Float64, max(length(ins), 1))
    check(ccall((:ag_ewise_colprog, LIB), Int32,
                (Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Float64}, Int32,
                 Ptr{Ptr{Cvoid}}, Int64, Int64, Ptr{Cvoid}),
                dtcode(T), code, length(code), consts, length(consts), length(ins), ins, modes, imms, length(outs), op,
                m, n, r === nothing ? C_NULL : r.ptr))
    outs
end

# ---- §10 graphs, events, hooks ------------------------------------------------------------------------------------------
struct Graph
    handle::UInt64
end
# capture(f): everything f enqueues (one recorded tape's forward + backward + update) becomes a CUDA graph; the tape
# logic runs at capture time only, `launch` replays the device work (src/core.jl:64-79,156-169 on replay).
function capture(f)
    check(ccall((:ag_graph_begin, LIB), Int32, ()))
    h = Ref{UInt64}(0)
    try
        f()
    catch
        ccall((:ag_graph_end, LIB), Int32, (Ref{UInt64},), h)
        rethrow()
    end
    check(ccall((:ag_graph_end, LIB), Int32, (Ref{UInt64},), h))
    Graph(h[])
end
launch(g::Graph) = check(ccall((:ag_graph_launch, LIB), Int32, (UInt64,), g.handle))
destroy(g::Graph) = check(ccall((:ag_graph_destroy, LIB), Int32, (UInt64,), g.handle))
macro graph(body)
    :(capture(() -> $(esc(body))))
end
struct Event
    h::UInt64
end
Event() = (r = Ref{UInt64}(0); check(ccall((:ag_event_create, LIB), Int32, (Ref{UInt64},), r)); Event(r[]))
record!(e::Event) = (check(ccall((:ag_event_record, LIB), Int32, (UInt64,), e.h)); e)
function elapsed_ms(a::Event, b::Event)
    ms = Ref{Float32}(0)
    check(ccall((:ag_event_elapsed_ms, LIB), Int32, (UInt64, UInt64, Ref{Float32}), a.h, b.h, ms))
    ms[]
end
# gcnode hook (src/core.jl:168,235-237): release a consumed node's buffers to the pool early
release_node(n, tape) = (n.outgrad = nothing; n.Value isa AutoGrad.Result && (n.Value.value = nothing); nothing)
use_release_node!() = AutoGrad.set_gc_function(release_node)

# C-side tape (csrc/tape.cu): one ccall per recorded operation, ONE for the whole backward sweep (src/core.jl:153-170)
struct CTape
    h::UInt64
end
CTape() = (r = Ref{UInt64}(0); check(ccall((:ag_tape_new, LIB), Int32, (Ref{UInt64},), r)); CTape(r[]))
free!(t::CTape) = check(ccall((:ag_tape_free, LIB), Int32, (UInt64,), t.h))
function leaf!(t::CTape, x::DArray{T,N}, tracked::Bool) where {T,N}
    id = Ref{Int32}(0)
    check(ccall((:ag_tape_leaf, LIB), Int32, (UInt64, Int32, Ptr{Cvoid}, Int32, Ptr{Int64}, Int32, Ref{Int32}),
                t.h, dtcode(T), x.ptr, N, collect(Int64, size(x)), tracked, id))
    id[]
end
function record!(t::CTape, op::Integer, arg::Integer, ins::Vector{Int32}, scalar::Real=0.0)
    id = Ref{Int32}(0)
    check(ccall((:ag_tape_record, LIB), Int32, (UInt64, Int32, Int32, Int32, Ptr{Int32}, Float64, Ref{Int32}),
                t.h, op, arg, length(ins), ins, Float64(scalar), id))
    id[]
end
backward!(t::CTape, result::Integer) = check(ccall((:ag_tape_backward, LIB), Int32, (UInt64, Int32), t.h, result))
function tape_grad(t::CTape, id::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ag_tape_grad, LIB), Int32, (UInt64, Int32, Ref{Ptr{Cvoid}}), t.h, id, p))
    p[]                                   # C_NULL = nothing (src/core.jl:215-216)
end
function tape_value(t::CTape, id::Integer)
    p, nd, shp = Ref{Ptr{Cvoid}}(C_NULL), Ref{Int32}(0), zeros(Int64, 4)
    check(ccall((:ag_tape_value, LIB), Int32, (UInt64, Int32, Ref{Ptr{Cvoid}}, Ref{Int32}, Ptr{Int64}), t.h, id, p, nd, shp))
    p[], Tuple(shp[1:nd[]])
end

# ---- §11 data parallel: one process per GPU, gradients summed over the ranks, then the same axpy! everywhere -------------
# The launcher (MPI.jl, Distributed, torchrun-style env) moves two small byte strings between the ranks: rank 0's
# 128-byte NCCL id and every rank's 64-byte CUDA-IPC handle.  `exchange_all(bytes)::Vector{Vector{UInt8}}` is that
# launcher-provided allgather.
mutable struct DataParallel
    rank::Int
    world::Int
    p2p::Bool
end
function DataParallel(rank::Integer, world::Integer, exchange_all; p2p_bytes::Integer=8 << 20)
    dp = DataParallel(rank, world, false)
    world == 1 && return dp
    id = zeros(UInt8, 128)
    rank == 0 && check(ccall((:ag_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    id = exchange_all(id)[1]
    check(ccall((:ag_comm_init, LIB), Int32, (Int32, Int32, Ptr{UInt8}), world, rank, id))
    h = zeros(UInt8, 64)
    ok = ccall((:ag_p2p_export, LIB), Int32, (Int64, Int32, Ptr{UInt8}), p2p_bytes, world, h) == 0
    got = exchange_all(vcat(UInt8(ok), h))
    if all(g -> g[1] == 1, got)
        hs = reduce(vcat, [g[2:end] for g in got])
        ok = ccall((:ag_p2p_connect, LIB), Int32, (Int32, Int32, Ptr{UInt8}), world, rank, hs) == 0
        dp.p2p = all(g -> g[1] == 1, exchange_all(UInt8[ok]))
    end
    dp.p2p || ccall((:ag_p2p_destroy, LIB), Int32, ())
    dp
end
# params[k] += alpha * (sum over ranks of grads[k]): ONE launch when the peer-memory kernel applies (Float32, <= 1 MB:
# allreduce and axpy! fused), otherwise pack -> NCCL allreduce -> multi-tensor axpy!
function allreduce_update!(dp::DataParallel, alpha::Number, gs::Vector{<:DArray{T}}, ws::Vector{<:DArray{T}}) where T
    dp.world == 1 && return multi_axpy!(alpha, gs, ws)
    sizes = Int64[length(g) for g in gs]
    if dp.p2p && T == Float32 && 4 * sum(cld.(sizes, 4) .* 4) <= (1 << 20) && length(gs) <= 16
        check(ccall((:ag_p2p_allreduce_multi, LIB), Int32, (Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Float64),
                    dtcode(T), length(gs), ptrs(gs), ptrs(ws), sizes, Float64(alpha)))
        return ws
    end
    bucket = DArray{T}(undef, (sum(sizes),))
    views, off = DArray{T,1}[], 0
    for n in sizes
        push!(views, DArray{T,1}(bucket.buf, bucket.ptr + off * sizeof(T), (n,))); off += n
    end
    check(ccall((:ag_multi_copy, LIB), Int32, (Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}),
                dtcode(T), length(gs), ptrs(gs), ptrs(views), sizes))
    check(ccall((:ag_allreduce_sum, LIB), Int32, (Int32, Ptr{Cvoid}, Int64), dtcode(T), bucket.ptr, length(bucket)))
    multi_axpy!(alpha, views, ws)
end
close(dp::DataParallel) = check(ccall((:ag_comm_destroy, LIB), Int32, ()))

# ---- §12 minibatches in the dataset's byte format (examples/mnist.jl:80-81 rebuilt on the device) ------------------------------
function ubyte_batch!(x::DArray{T,2}, y::DArray{T,2}, px_dev::Ptr{Cvoid}, lb_dev::Ptr{Cvoid},
                      pixels::Array{UInt8}, labels::Vector{UInt8}) where T
    check(ccall((:ag_copy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), px_dev, pixels, length(pixels)))
    check(ccall((:ag_copy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), lb_dev, labels, length(labels)))
    check(ccall((:ag_ubyte_to_float, LIB), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float64),
                dtcode(T), px_dev, x.ptr, length(x), 255.0))                       # xshape: a ./ 255f0
    check(ccall((:ag_onehot_labels, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Int32, Int32, Ptr{Cvoid}),
                dtcode(T), lb_dev, size(y, 2), size(y, 1), 10, y.ptr))             # yshape: a[a.==0] = 10; one-hot
    x, y
end

end # module

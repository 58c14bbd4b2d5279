# AutoGradB200.jl -- the Julia side of the drop-in: a device array type whose methods `ccall`
# libagb200.so (include/agb200.h).  Load AFTER `using AutoGrad`; then `Param(DArray(x))` records and
# differentiates through the unmodified AutoGrad.jl tape (src/core.jl) with every array operation on B200.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: no Julia toolchain exists in the build image (see DESIGN.md §1).
# The Python package autograd.jl_b200/ issues exactly the same C calls in the same order and is what the
# tests drive; this file is the binding a maintainer would add on the reference side.
module AutoGradB200

using AutoGrad
import AutoGrad: back, Arg, unbroadcast, addto!, Sparse, full, zeroslike, recording, value
import Base: size, length, eltype, similar, copy, sum, *, adjoint, reshape, vec, zero, fill!, getindex, show
import Base.Broadcast: broadcasted, materialize
import LinearAlgebra: axpy!, lmul!

const LIB = get(ENV, "AGB200_LIB", joinpath(@__DIR__, "..", "autograd.jl_b200", "libagb200.so"))
const F32, F64 = Int32(0), Int32(1)
dtcode(::Type{Float32}) = F32
dtcode(::Type{Float64}) = F64

function check(status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:ag_last_error, LIB), Int32, (Ptr{UInt8}, Int64), buf, 1024)
    msg = unsafe_string(pointer(buf))
    status == 1 && throw(DimensionMismatch(msg))
    status == 2 && throw(OutOfMemoryError())
    status == 5 && throw(MethodError(ccall, (msg,)))       # no device method, no CPU fallback
    error("libagb200 status $status: $msg")
end

init(dev::Integer=0) = check(ccall((:ag_init, LIB), Int32, (Int32,), dev))
sync() = check(ccall((:ag_sync, LIB), Int32, ()))          # the `@gs` hook, src/AutoGrad.jl:11

# ---- storage --------------------------------------------------------------------------------------
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

struct DArray{T,N} <: AbstractArray{T,N}
    buf::Buffer
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    tr::Bool                   # lazy adjoint of a matrix (src/linearalgebra.jl:16)
end
DArray{T}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N} =
    (b = Buffer(max(prod(dims), 1) * sizeof(T)); DArray{T,N}(b, b.ptr, dims, false))
function DArray(x::Array{T,N}) where {T<:Union{Float32,Float64},N}
    a = DArray{T}(undef, size(x))
    check(ccall((:ag_copy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), a.ptr, x, sizeof(x))); sync()
    a
end
function Base.Array(a::DArray{T,N}) where {T,N}
    x = Array{T,N}(undef, a.dims)
    check(ccall((:ag_copy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), x, a.ptr, sizeof(x)))
    x
end
size(a::DArray) = a.dims
eltype(::DArray{T}) where T = T
similar(a::DArray{T}, dims::Dims=size(a)) where T = DArray{T}(undef, dims)
reshape(a::DArray{T}, dims::Dims) where T = DArray{T,length(dims)}(a.buf, a.ptr, dims, false)
vec(a::DArray) = reshape(a, (length(a),))
adjoint(a::DArray{T,2}) where T = DArray{T,2}(a.buf, a.ptr, reverse(a.dims), !a.tr)
fill!(a::DArray{T}, v) where T =
    (check(ccall((:ag_fill, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Float64), a.ptr, length(a), dtcode(T), v)); a)
zero(a::DArray) = fill!(similar(a), 0)
show(io::IO, a::DArray) = print(io, "DArray", size(a))

# ---- elementwise (ids: include/agb200.h) -------------------------------------------------------------
const UNARY = Dict(:- => 1, :abs => 2, :abs2 => 3, :exp => 4, :log => 5, :tanh => 6, :sqrt => 8, :sin => 9, :cos => 10)
const BINARY = Dict(:+ => 0, :- => 1, :* => 2, :/ => 3, :max => 4, :min => 5, :^ => 6)

function unary_fwd(op::Integer, x::DArray{T}) where T
    y = similar(x)
    check(ccall((:ag_unary_fwd, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64), op, dtcode(T), x.ptr, y.ptr, length(x)))
    y
end
# dx = dy .* g(x,y): ONE fused kernel per rule of src/math.jl (only outside of recording, SURVEY.md 3.4)
function unary_bwd(op::Integer, dy, x::DArray{T}, y) where T
    dx = similar(x)
    mode, dyp, dys = dy isa Number ? (Int32(2), C_NULL, Float64(dy)) : (Int32(0), dy.ptr, 0.0)
    check(ccall((:ag_unary_bwd, LIB), Int32,
                (Int32, Int32, Ptr{Cvoid}, Int32, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64),
                op, dtcode(T), dyp, mode, dys, x.ptr, y === nothing ? C_NULL : y.ptr, dx.ptr, length(x)))
    dx
end
for (f, id) in UNARY
    @eval broadcasted(::typeof($f), x::DArray) = unary_fwd($id, x)
end
# device-specialised gradient methods, e.g. tanh (src/math.jl:74) and exp (src/math.jl:42):
back(::typeof(broadcasted), ::Type{Arg{2}}, dy, y, ::typeof(tanh), x::AutoGrad.Value{<:DArray}) =
    recording() ? dy .* (1 .- abs2.(y)) : unary_bwd(6, dy, value(x), value(y))
back(::typeof(broadcasted), ::Type{Arg{2}}, dy, y, ::typeof(exp), x::AutoGrad.Value{<:DArray}) =
    recording() ? dy .* y : unary_bwd(4, dy, value(x), value(y))

# strides of an operand inside the broadcast shape (0 = broadcast), Julia leading-dimension alignment
function bstrides(shape::Dims, out::Dims)
    st, acc = Int64[], 1
    for d in 1:length(out)
        e = d <= length(shape) ? shape[d] : 1
        push!(st, (e == 1 && out[d] != 1) ? 0 : acc); acc *= e
    end
    st
end
function binary_fwd(op::Integer, a, b)
    T = eltype(a isa DArray ? a : b)
    out = Base.Broadcast.broadcast_shape(a isa DArray ? size(a) : (), b isa DArray ? size(b) : ())
    y = DArray{T}(undef, out)
    sa = a isa DArray ? bstrides(size(a), out) : zeros(Int64, length(out))
    sb = b isa DArray ? bstrides(size(b), out) : zeros(Int64, length(out))
    shp = collect(Int64, out)      # (the Python mirror additionally collapses dims; <= 4 are accepted)
    check(ccall((:ag_binary_fwd, LIB), Int32,
                (Int32, Int32, Ptr{Cvoid}, Float64, Ptr{Int64}, Ptr{Cvoid}, Float64, Ptr{Int64}, Ptr{Cvoid}, Int32, Ptr{Int64}),
                op, dtcode(T), a isa DArray ? a.ptr : C_NULL, a isa Number ? a : 0.0, sa,
                b isa DArray ? b.ptr : C_NULL, b isa Number ? b : 0.0, sb, y.ptr, length(shp), shp))
    y
end
for (f, id) in BINARY
    @eval broadcasted(::typeof($f), a::DArray, b::Union{DArray,Number}) = binary_fwd($id, a, b)
    @eval broadcasted(::typeof($f), a::Number, b::DArray) = binary_fwd($id, a, b)
end
Base.:+(a::DArray, b::DArray) = binary_fwd(0, a, b)       # addto!(a,b) = a+b, src/addto.jl:23
Base.:-(a::DArray) = unary_fwd(1, a)

# ---- reductions: sum(x), sum(abs2,x), sum(x;dims) and unbroadcast (src/unbroadcast.jl) -----------------
function sum(x::DArray{T}; dims=:) where T
    if dims === Colon()
        out = DArray{T}(undef, ())
        check(ccall((:ag_sum_all, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}), 0, dtcode(T), x.ptr, length(x), out.ptr))
        return Array(out)[]      # the reference returns a host Number
    end
    d = dims isa Integer ? dims : only(dims)
    inner, red, outer = prod(size(x)[1:d-1]), size(x, d), prod(size(x)[d+1:end])
    out = DArray{T}(undef, ntuple(i -> i == d ? 1 : size(x, i), ndims(x)))
    check(ccall((:ag_sum_dim, LIB), Int32, (Int32, Int32, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Cvoid}),
                0, dtcode(T), x.ptr, inner, red, outer, out.ptr))
    out
end

# ---- matmul: forward `*` and both gradient products of src/base.jl:26 ---------------------------------
function *(a::DArray{T,2}, b::DArray{T,2}) where T
    m, k = size(a); k2, n = size(b)
    k == k2 || throw(DimensionMismatch("A has dimensions ($m,$k) but B has dimensions ($k2,$n)"))
    c = DArray{T}(undef, (m, n))
    lda = a.tr ? size(a, 2) : size(a, 1); ldb = b.tr ? size(b, 2) : size(b, 1)
    check(ccall((:ag_gemm, LIB), Int32,
                (Int32, Int32, Int32, Int64, Int64, Int64, Float64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Float64, Ptr{Cvoid}, Int64),
                dtcode(T), a.tr, b.tr, m, n, k, 1.0, a.ptr, lda, b.ptr, ldb, 0.0, c.ptr, m))
    c
end

# ---- update ------------------------------------------------------------------------------------------
axpy!(alpha::Number, x::DArray{T}, y::DArray{T}) where T =
    (check(ccall((:ag_axpy, LIB), Int32, (Int32, Float64, Ptr{Cvoid}, Ptr{Cvoid}, Int64), dtcode(T), alpha, x.ptr, y.ptr, length(y))); y)
axpy!(alpha::Number, x::DArray, y::AutoGrad.Param) = (axpy!(alpha, x, value(y)); y)

# ---- Sparse on a device container: one batched deterministic scatter-add (src/addto.jl:91-98) ---------
function addto!(a::DArray{T,2}, b::Sparse) where T
    recording() && (a = copy(a))
    cols = Int64[]; vals = DArray[]
    for (idx, val) in zip(b.indices, b.values)         # (Colon(), ids) blocks of the embedding lookup
        append!(cols, idx[2] .- 1); push!(vals, val)
    end
    v = length(vals) == 1 ? vals[1] : cat1d_device(vals)
    d_idx = DArray(reinterpret(Float64, cols))          # raw int64 bits on device
    check(ccall((:ag_scatter_add, LIB), Int32, (Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64),
                dtcode(T), a.ptr, length(a), d_idx.ptr, v.ptr, length(cols), size(a, 1)))
    a
end

# ---- horizontal fusion (SURVEY.md 8f-1): deferred elementwise chains -> ONE launch ------------------------
# The executable host logic lives in autograd.jl_b200/lazy.py; in Julia the same thing is a `LazyDArray <: DArray`
# whose `getproperty(:ptr)` forces the pending expression DAG.  The ccall it ends in:
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
# same arguments plus the (inner, red, outer) view of the operand and the result array; a MethodError-class status
# (5) means "not fusable here" and the caller falls back to forcing x and calling ag_sum_dim.
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

# ---- minibatches in the dataset's byte format (examples/mnist.jl:80-81 rebuilt on the device) -------------
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

# gcnode hook (src/core.jl:168,235-237): release a consumed node's buffers to the pool early
AutoGrad.set_gc_function((n, tape) -> (n.outgrad = nothing; n.Value isa AutoGrad.Result && (n.Value.value = nothing)))

end # module

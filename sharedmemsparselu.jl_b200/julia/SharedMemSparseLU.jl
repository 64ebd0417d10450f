# SharedMemSparseLU.jl -- drop-in shim: the reference package's Julia API on top of libsmslu.so.
#
# Same module name, same exports and the same call signatures as the reference
# (/root/reference/src/SharedMemSparseLU.jl, cited as src:N):
#     ParallelSparseLU(A[, chunk_size])   src:64     -> smslu_create + smslu_analyze + smslu_refactor
#     lu!(F, A)                           src:245    -> smslu_refactor
#     ldiv!(x, F, b)                      src:286    -> smslu_solve
#     lsolve!(F, x), rsolve!(F, x)        src:349,374-> smslu_lsolve / smslu_rsolve
#     F.m F.n F.L F.U F.p F.q F.Rs        src:45-51  -> smslu_get_factors (lazy, cached per factorization)
#     cleanup_ParallelSparseLU!(F)        src:31 (exported, undefined there) -> smslu_destroy
#     allocate_shared                     src:31 (exported, undefined there) -> documented no-op
# Every numeric operation happens in hand-written sm_100a CUDA kernels behind the C ABI
# (include/smslu.h); this file only marshals arguments with `ccall`.  No CUDA.jl, no CPU fallback.
#
# NOT EXECUTED in the build environment (no Julia toolchain there); the same C entry points are
# exercised by the Python mirror (sharedmemsparselu.jl_b200/__init__.py) in tests/.
#
# Where the pivot order comes from:
#   pivots = :umfpack (default)  `lu(A)` from SparseArrays (UMFPACK, what the reference calls at
#                                src:74) is run on the host for its (p, q, Rs); the GPU then
#                                factors with exactly those pivots, so F.p, F.q, F.Rs are the
#                                reference's and L, U agree to rounding.  `lu!(F, A)` keeps them while the
#                                new values pass the threshold test on the GPU (every |l_ij| <= 1/pivot_tol);
#                                when the test fails (SMSLU_E_REPIVOT) or a pivot is zero, UMFPACK is asked
#                                again and the object is re-analysed -- the reference's re-pivot /
#                                re-chunk branch (src:252-273).
#   pivots = :native             the library's own fill-reducing ordering (nested dissection on
#                                A+A'), diagonal pivots, UMFPACK-style row scaling Rs = 1/sum|a_ij|:
#                                for matrices known to need no pivoting (M-matrices, diagonally dominant
#                                blocks); this is also the ordering the multi-GPU partition needs.  A failed
#                                threshold test falls back to :umfpack pivots unless `strict=true`.
module SharedMemSparseLU

export ParallelSparseLU, cleanup_ParallelSparseLU!, allocate_shared, comm_unique_id

using LinearAlgebra
using SparseArrays

import LinearAlgebra: ldiv!, lu!

const libsmslu = get(ENV, "SMSLU_LIB", joinpath(@__DIR__, "..", "libsmslu.so"))

# ---- error codes (include/smslu.h)
const SMSLU_E_DIM = -1
const SMSLU_E_PIVOT = -2
const SMSLU_E_PATTERN = -3
const SMSLU_E_REPIVOT = -9

"Thrown by `lu!` (strict mode only) when the static pivot order fails the threshold test for the new values."
struct PivotThresholdError <: Exception
    msg::String
end

# mirror of smslu_options_t (12 Int32, one Float64, 4 Int32 = 72 bytes)
struct SmsluOptions
    ordering::Int32
    grid::NTuple{3,Int32}
    nd_leaf::Int32
    relax::Int32
    max_width::Int32
    scaling::Int32
    device::Int32
    reserved0::Int32
    nranks::Int32
    rank::Int32
    pivot_tol::Float64
    reserved::NTuple{4,Int32}
end

const ORDERINGS = Dict(:auto => 0, :natural => 1, :given => 2, :nd_graph => 3, :nd_grid => 4)

function default_options()
    o = Ref{SmsluOptions}()
    ccall((:smslu_options_default, libsmslu), Cint, (Ptr{SmsluOptions},), o)
    return o[]
end

function last_error(h::Ptr{Cvoid})
    s = ccall((:smslu_last_error, libsmslu), Cstring, (Ptr{Cvoid},), h)
    return s == C_NULL ? "" : unsafe_string(s)
end

function check(h::Ptr{Cvoid}, rc::Integer)
    rc == 0 && return nothing
    msg = last_error(h)
    rc == SMSLU_E_DIM && throw(DimensionMismatch(msg))
    rc == SMSLU_E_PIVOT && throw(SingularException(0))
    rc == SMSLU_E_PATTERN && throw(ArgumentError(msg))
    rc == SMSLU_E_REPIVOT && throw(PivotThresholdError(msg))
    error("smslu error $rc: $msg")
end

"""
    ParallelSparseLU(A::SparseMatrixCSC{Float64,Int64}, chunk_size=nothing; pivots=:umfpack,
                     ordering=:auto, grid=nothing, device=-1, strict=false)

Factorize `A` on the GPU.  `chunk_size` is accepted for compatibility with the reference
(src:64-72) and ignored: the dense column-chunk layout it sized (src:101-178) does not exist here.
"""
mutable struct ParallelSparseLU{Tf,Ti}
    m::Ti
    n::Ti
    handle::Ptr{Cvoid}
    colptr::Vector{Int64}
    rowval::Vector{Int64}
    chunk_size::Ti
    Rs_given::Union{Vector{Float64},Nothing}
    cache::Dict{Symbol,Any}
    pivots::Symbol
    strict::Bool
    opts::SmsluOptions
    comm_id::Union{Vector{UInt8},Nothing}

    function ParallelSparseLU(A::SparseMatrixCSC{Tf,Ti}, chunk_size=nothing;
                              pivots::Symbol=:umfpack, ordering::Symbol=:auto,
                              grid=nothing, device::Integer=-1, strict::Bool=false,
                              nranks::Integer=1, rank::Integer=0, pivot_tol::Real=0.0,
                              comm_id::Union{Vector{UInt8},Nothing}=nothing) where {Tf<:Float64,Ti<:Int64}
        size(A, 1) == size(A, 2) || throw(DimensionMismatch("matrix is not square: $(size(A))"))
        chunk_size === nothing && (chunk_size = 8)                       # src:67-70
        chunk_size = min(chunk_size, A.n)                                # src:72
        nranks > 1 && (pivots = :native)                                 # the partition is the top of OUR nested dissection
        o = default_options()
        g = grid === nothing ? (Int32(0), Int32(0), Int32(0)) :
            (Int32(grid[1]), Int32(length(grid) > 1 ? grid[2] : 1), Int32(length(grid) > 2 ? grid[3] : 1))
        o = SmsluOptions(Int32(ORDERINGS[ordering]), g, o.nd_leaf, o.relax, o.max_width, o.scaling,
                         Int32(device), Int32(0), Int32(nranks), Int32(rank), Float64(pivot_tol), o.reserved)
        F = new{Tf,Ti}(A.m, A.n, C_NULL, copy(A.colptr), copy(A.rowval), chunk_size, nothing, Dict{Symbol,Any}(),
                       pivots, strict, o, comm_id)
        finalizer(cleanup_ParallelSparseLU!, F)
        if pivots === :umfpack
            build!(F, A, true)
        else
            try
                build!(F, A, false)
            catch e
                (e isa PivotThresholdError || e isa SingularException) && !strict || rethrow()
                build!(F, A, true)                                       # the diagonal pivots fail for this matrix
            end
        end
        return F
    end
end

# (Re)create the handle: host pivot search with UMFPACK when `umfpack`, analysis, first numeric factorization.
function build!(F::ParallelSparseLU, A::SparseMatrixCSC, umfpack::Bool)
    cleanup_ParallelSparseLU!(F)
    o = getfield(F, :opts)
    p = q = nothing
    setfield!(F, :Rs_given, nothing)
    if umfpack
        F0 = lu(A)                                                       # host UMFPACK (src:74): ordering + threshold pivoting
        p, q = Vector{Int64}(F0.p), Vector{Int64}(F0.q)
        setfield!(F, :Rs_given, Vector{Float64}(F0.Rs))
        o = SmsluOptions(Int32(ORDERINGS[:given]), o.grid, o.nd_leaf, o.relax, o.max_width, o.scaling,
                         o.device, o.reserved0, o.nranks, o.rank, o.pivot_tol, o.reserved)
        setfield!(F, :pivots, :umfpack)
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:smslu_create, libsmslu), Cint,
               (Ptr{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{SmsluOptions}),
               h, A.n, A.colptr, A.rowval, 1, Ref(o))
    rc == 0 || error("smslu_create failed with code $rc")
    setfield!(F, :handle, h[])
    check(h[], ccall((:smslu_analyze, libsmslu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}),
                     h[], p === nothing ? C_NULL : p, q === nothing ? C_NULL : q))
    if o.nranks > 1
        # one process (MPI rank) per GPU: every rank analyses the same pattern; the 128-byte id comes from
        # `comm_unique_id()` on rank 0, e.g.  id = MPI.bcast(rank == 0 ? comm_unique_id() : nothing, 0, comm)
        id = getfield(F, :comm_id)
        id === nothing && throw(ArgumentError("nranks > 1 needs comm_id (see comm_unique_id)"))
        check(h[], ccall((:smslu_comm_init, libsmslu), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), h[], id, length(id)))
    end
    refactor!(F, A)
    # later lu!(F, A) calls recompute the row scaling from the new values on the GPU (same formula as UMFPACK's
    # default, Rs[i] = 1 / sum_j |a_ij|); the first factorization used UMFPACK's own vector
    setfield!(F, :Rs_given, nothing)
    return nothing
end

function refactor!(F::ParallelSparseLU, A::SparseMatrixCSC)
    empty!(getfield(F, :cache))
    Rs = getfield(F, :Rs_given)
    h = getfield(F, :handle)
    check(h, ccall((:smslu_refactor, libsmslu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                   h, A.nzval, Rs === nothing ? C_NULL : Rs))
    return nothing
end

"""
    comm_unique_id() -> Vector{UInt8}

128-byte NCCL id for a multi-GPU `ParallelSparseLU`; call on rank 0 and broadcast to the other ranks.
"""
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:smslu_comm_unique_id, libsmslu), Cint, (Ptr{UInt8}, Int64), id, 128)
    rc == 0 || error("smslu_comm_unique_id failed with code $rc")
    return id
end

"""
    lu!(F::ParallelSparseLU, A::SparseMatrixCSC)

Numeric refactorization with the sparsity pattern `F` was analysed for (reference src:245-279).  The pivot order of
the previous factorization is kept as long as the new values pass the threshold test on the GPU; when they do not
(or a pivot is zero) UMFPACK is asked for fresh pivots on the host and `F` is re-analysed, which is what the
reference's re-pivot / re-chunk branch (src:252-273) amounts to.  With `strict=true` (or on a multi-GPU object) a
`PivotThresholdError` / `SingularException` is thrown instead.  A pattern change raises `ArgumentError`.  Returns `nothing`.
"""
function lu!(F::ParallelSparseLU{Tf,Ti}, A::Union{SparseMatrixCSC{Tf,Ti},Nothing}) where {Tf,Ti}
    A === nothing && throw(ArgumentError("lu!(F, nothing) is not supported (nor does the reference's Nothing arm work, src:246-247)"))
    (A.m == F.m && A.n == F.n) || throw(DimensionMismatch("matrix size differs from the factor object"))
    (A.colptr == F.colptr && A.rowval == F.rowval) ||
        throw(ArgumentError("sparsity pattern differs from the analysed one"))
    try
        refactor!(F, A)
    catch e
        (e isa PivotThresholdError || e isa SingularException) || rethrow()
        (getfield(F, :strict) || getfield(F, :opts).nranks > 1) && rethrow()
        build!(F, A, true)
    end
    return nothing
end

function fetch_factors!(F::ParallelSparseLU)
    c = getfield(F, :cache)
    haskey(c, :L) && return c
    h = getfield(F, :handle)
    n = getfield(F, :n)
    nl = Ref{Int64}(0); nu = Ref{Int64}(0)
    check(h, ccall((:smslu_get_nnz, libsmslu), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h, nl, nu))
    lp = Vector{Int64}(undef, n + 1); li = Vector{Int64}(undef, nl[]); lx = Vector{Float64}(undef, nl[])
    up = Vector{Int64}(undef, n + 1); ui = Vector{Int64}(undef, nu[]); ux = Vector{Float64}(undef, nu[])
    p = Vector{Int64}(undef, n); q = Vector{Int64}(undef, n); Rs = Vector{Float64}(undef, n)
    check(h, ccall((:smslu_get_factors, libsmslu), Cint,
                   (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64},
                    Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32),
                   h, lp, li, lx, up, ui, ux, p, q, Rs, 1))
    c[:L] = SparseMatrixCSC(n, n, lp, li, lx)      # unit diagonal stored explicitly, rows sorted
    c[:U] = SparseMatrixCSC(n, n, up, ui, ux)
    c[:p] = p; c[:q] = q; c[:Rs] = Rs
    return c
end

function Base.getproperty(F::ParallelSparseLU, s::Symbol)
    if s === :L || s === :U || s === :p || s === :q || s === :Rs
        return fetch_factors!(F)[s]          # F.L*F.U == (F.Rs .* A)[F.p, F.q]   (src:307)
    end
    return getfield(F, s)
end

"""
    ldiv!(x::AbstractVector, F::ParallelSparseLU, b::AbstractVector)

Solves `A*x=b` where `F` is the LU factorisation of `A`, overwriting `x` (reference src:281-342).
`b` is not modified.  Matrices `X`, `B` (column-major, one right-hand side per column) are accepted
as an extension.
"""
function ldiv!(x::StridedVecOrMat{Float64}, F::ParallelSparseLU{Tf,Ti}, b::StridedVecOrMat{Float64}) where {Tf,Ti}
    F.m == F.n || throw(DimensionMismatch("`F` is not square: F.m=$(F.m), F.n=$(F.n)"))                         # src:288
    size(x, 1) == F.n || throw(DimensionMismatch("`x` does not have same size as F: length(x)=$(size(x,1)), F.n=$(F.n)"))  # src:289
    size(b, 1) == F.n || throw(DimensionMismatch("`b` does not have same size as F: length(b)=$(size(b,1)), F.n=$(F.n)"))  # src:290
    size(x, 2) == size(b, 2) || throw(DimensionMismatch("x and b have different numbers of columns"))
    (stride(x, 1) == 1 && stride(b, 1) == 1) || throw(ArgumentError("x and b must have unit stride"))
    check(F.handle, ccall((:smslu_solve, libsmslu), Cint,
                          (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64),
                          F.handle, x, size(x, 1), b, size(b, 1), size(x, 2),
                          ndims(x) == 1 ? F.n : stride(x, 2), ndims(b) == 1 ? F.n : stride(b, 2)))
    return x
end

"""
    lsolve!(F, x)

Solve `L*x = b` in place, where `F.L` is the lower triangular factor (reference src:344-367).
"""
function lsolve!(F::ParallelSparseLU, x::StridedVecOrMat{Float64})
    check(F.handle, ccall((:smslu_lsolve, libsmslu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64),
                          F.handle, x, size(x, 1), size(x, 2), ndims(x) == 1 ? F.n : stride(x, 2)))
    return nothing
end

"""
    rsolve!(F, x)

Solve `U*x = b` in place, where `F.U` is the upper triangular factor (reference src:369-392).
"""
function rsolve!(F::ParallelSparseLU, x::StridedVecOrMat{Float64})
    check(F.handle, ccall((:smslu_rsolve, libsmslu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64),
                          F.handle, x, size(x, 1), size(x, 2), ndims(x) == 1 ? F.n : stride(x, 2)))
    return nothing
end

"""
    cleanup_ParallelSparseLU!(F)

Release the GPU memory, streams and events owned by `F` (the reference exports this name without
defining it, src:31).  Safe to call more than once; also installed as the finalizer.
"""
function cleanup_ParallelSparseLU!(F::ParallelSparseLU)
    h = getfield(F, :handle)
    if h != C_NULL
        ccall((:smslu_destroy, libsmslu), Cint, (Ptr{Cvoid},), h)
        setfield!(F, :handle, C_NULL)
    end
    return nothing
end

"""
    allocate_shared(args...)

Exported by the reference without a definition or a call site (src:31); kept as a no-op so that
`using SharedMemSparseLU` exposes the same names.
"""
allocate_shared(args...) = (ccall((:smslu_allocate_shared, libsmslu), Cint, ()); nothing)

# Diagnostics: nnz, supernodes, levels, per-phase device times of the last calls.
function stats(F::ParallelSparseLU)
    buf = zeros(UInt8, 8 * (14 + 9 + 5 + 16 + 16 + 8))     # sizeof(smslu_stats_t)
    check(F.handle, ccall((:smslu_get_stats, libsmslu), Cint, (Ptr{Cvoid}, Ptr{UInt8}), F.handle, buf))
    i64 = reinterpret(Int64, buf); f64 = reinterpret(Float64, buf)
    return (n=i64[1], nnz_a=i64[2], nnz_l=i64[3], nnz_u=i64[4], supernodes=i64[7], levels=i64[8],
            flops=f64[15], ms_analyze=f64[17], ms_refactor=f64[19], ms_solve=f64[20])
end

end # module SharedMemSparseLU

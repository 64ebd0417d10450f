# Parity of the shim with the REAL reference (Julia + UMFPACK), for whoever has both a Julia toolchain and a B200
# (SURVEY.md section 8c item 5).  NOT EXECUTED in the build environment.
#
#   julia sharedmemsparselu.jl_b200/julia/test/parity_umfpack.jl /path/to/SharedMemSparseLU.jl-checkout
#
# Checks, on the reference's own fixture families (test:12-21 FE-like matrices, dense rand(n,n)) and on the BASELINE
# Laplacians: with pivots = :umfpack the shim's F.p, F.q, F.Rs are the reference's bit for bit; the patterns of F.L and
# F.U are identical; entries agree to 1e-12 relative (column-scale floor for cancelled entries); ldiv!, lsolve!, rsolve!
# agree at the reference's tolerances (1e-12; 1e-10 for dense / rsolve, test:25-26); lu! with new values follows.
using LinearAlgebra, SparseArrays, Random, Test

length(ARGS) >= 1 || error("usage: julia parity_umfpack.jl /path/to/reference/checkout")
const RefMod = Module(:RefHost)
Base.include(RefMod, joinpath(ARGS[1], "src", "SharedMemSparseLU.jl"))      # the reference, under RefHost.SharedMemSparseLU
const Ref = RefMod.SharedMemSparseLU
pushfirst!(LOAD_PATH, normpath(joinpath(@__DIR__, "..")))
import SharedMemSparseLU as Shim

function fe_matrix(rng, nel, ngr=5)                                         # test:12-21
    n = nel * (ngr - 1) + 1
    mat = zeros(n, n)
    for iel in 1:nel
        imin = (iel - 1) * (ngr - 1) + 1; imax = iel * (ngr - 1) + 1
        mat[imin:imax, imin:imax] .= rand(rng, ngr, ngr)
    end
    return sparse(mat)
end
lap1d(n) = spdiagm(-1 => -ones(n - 1), 1 => -ones(n - 1))
lap2d(n) = kron(sparse(I, n, n), lap1d(n)) + kron(lap1d(n), sparse(I, n, n)) + 4.0 * sparse(I, n * n, n * n)

function relerr_cols(X::SparseMatrixCSC, Y::SparseMatrixCSC)
    worst = 0.0
    for j in 1:size(X, 2)
        r = nzrange(Y, j); isempty(r) && continue
        cmax = maximum(abs, view(Y.nzval, r))
        for t in r
            worst = max(worst, abs(X.nzval[t] - Y.nzval[t]) / max(abs(Y.nzval[t]), 1e-2 * cmax, 1e-300))
        end
    end
    return worst
end

function compare(A; tol=1e-12, soltol=1e-12)
    Fr = Ref.ParallelSparseLU(A)
    Fs = Shim.ParallelSparseLU(A; pivots=:umfpack)
    @test Fs.p == Fr.p && Fs.q == Fr.q && Fs.Rs == Fr.Rs                   # bit-exact permutations, pivots and scaling
    @test Fs.L.colptr == Fr.L.colptr && Fs.L.rowval == Fr.L.rowval          # bit-exact symbolic structure
    @test Fs.U.colptr == Fr.U.colptr && Fs.U.rowval == Fr.U.rowval
    @test relerr_cols(Fs.L, Fr.L) < tol && relerr_cols(Fs.U, Fr.U) < tol
    n = size(A, 1)
    b = rand(MersenneTwister(n), n)
    xr = similar(b); xs = similar(b)
    ldiv!(xr, Fr, b); ldiv!(xs, Fs, b)
    @test isapprox(xs, xr, rtol=soltol, atol=soltol)
    @test norm(A * xs - b) <= 4 * norm(A * xr - b) + 1e-15 * norm(b)        # residual no worse than the reference's
    yr = copy(b); ys = copy(b)
    Ref.lsolve!(Fr, yr); Shim.lsolve!(Fs, ys)
    @test isapprox(ys, yr, rtol=1e-12, atol=1e-12)
    yr = copy(b); ys = copy(b)
    Ref.rsolve!(Fr, yr); Shim.rsolve!(Fs, ys)
    @test isapprox(ys, yr, rtol=1e-10, atol=1e-10)
    return Fr, Fs
end

@testset "parity with the reference (UMFPACK)" begin
    rng = MersenneTwister(47)
    @testset "FE-like fixture nel=$nel" for nel in (1, 2, 6, 20, 64, 200)
        A = fe_matrix(rng, nel)
        Fr, Fs = compare(A)
        A2 = fe_matrix(rng, nel)                                            # lu! with new values, same pattern (test:171-186)
        lu!(Fr, A2); lu!(Fs, A2)
        b = rand(rng, size(A, 1)); xr = similar(b); xs = similar(b)
        ldiv!(xr, Fr, b); ldiv!(xs, Fs, b)
        @test isapprox(xs, xr, rtol=1e-12, atol=1e-12)
    end
    @testset "dense n=$n" for n in (1, 2, 8, 33, 100, 200)
        A = sparse(rand(rng, n, n))
        Fr, Fs = compare(A; tol=1e-10, soltol=1e-10)
        A2 = sparse(rand(rng, n, n))                                        # test:129-131
        lu!(Fr, A2); lu!(Fs, A2)
        b = rand(rng, n); xr = similar(b); xs = similar(b)
        ldiv!(xr, Fr, b); ldiv!(xs, Fs, b)
        @test isapprox(xs, xr, rtol=1e-10, atol=1e-10)
    end
    @testset "2D Laplacian $n x $n (BASELINE configs[0] at n = 100)" for n in (10, 37, 100)
        compare(lap2d(n))
    end
end

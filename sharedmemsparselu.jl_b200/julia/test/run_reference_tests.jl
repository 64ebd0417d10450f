# Run the reference's OWN test file, unchanged, against the drop-in shim.
#
#   julia sharedmemsparselu.jl_b200/julia/test/run_reference_tests.jl /path/to/SharedMemSparseLU.jl-checkout
#
# The reference's test/runtests.jl does `using SharedMemSparseLU` and `using SharedMemSparseLU: lsolve!, rsolve!`
# (test:8-9); putting this directory's parent first on LOAD_PATH makes that name resolve to the shim
# (sharedmemsparselu.jl_b200/julia/SharedMemSparseLU.jl) instead of the reference package.  Needs libsmslu.so
# (python sharedmemsparselu.jl_b200/build.py; or ENV["SMSLU_LIB"]) and a B200.
#
# NOT EXECUTED in the build environment: there is no Julia toolchain in that image (SURVEY.md F5).  The six testsets
# (2 400 assertions, test:38-188) are mirrored in Python by tests/test_gpu_parity.py on top of the same C entry points.
pushfirst!(LOAD_PATH, normpath(joinpath(@__DIR__, "..")))
length(ARGS) >= 1 || error("usage: julia run_reference_tests.jl /path/to/reference/checkout")
reftests = joinpath(ARGS[1], "test", "runtests.jl")
isfile(reftests) || error("not found: $reftests")
using SharedMemSparseLU
@assert pathof(SharedMemSparseLU) == normpath(joinpath(@__DIR__, "..", "SharedMemSparseLU.jl")) "the reference package shadows the shim"
include(reftests)

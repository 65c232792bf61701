# CloudAtlasB200.jl -- drop-in Julia binding of libcloudatlas_b200.so for CloudAtlas.jl's hot path.
#
#     using CloudAtlas
#     import CloudAtlasB200: ODEModel, hookstepsolve      # instead of CloudAtlas's own two names
#
# keeps every script and notebook of the package working unchanged: `ODEModel(α, γ, J, K, L, H; normalize)` returns an
# object with the reference struct's fields (src/ODEModels.jl:6-20)
#     α, γ, H, ijkl, Ψ, B, A1, A2, N, Bfact, f, Df
# where the Galerkin assembly, `N(x)`, `derivative(N, x)`, `f(x, R)`, `Df(x, R)` and `hookstepsolve(model, ...)` run
# on the B200 through the C ABI of include/cloudatlas_b200.h.  Additive entry points: ensembles (`model.f(X, R)` with
# one state per row of `X`, `model.N(X)`, `Df(X, R)`), `ngpus = n` (sharded assembly with the library's own NCCL
# all-gather, ensembles split by state), `cnab2`, `observables`.
# Only Float64 goes to the library; Rational / Dual element types stay on CloudAtlas.jl's generic methods.
#
# NOTE: Julia is not installed in the build image, so this file is not executed there.  It is held to the C ABI by
#   * tests/test_julia_binding_cpu.py: every ccall below is parsed and checked against the header (symbol exported,
#     argument count, Julia type vs C type) and the struct fields against src/ODEModels.jl:6-20;
#   * tests/c_harness/find_eqb.c: the same call sequence as `ODEModel(...)`, `model.f`, `hookstepsolve(model, ...)`
#     below, compiled with gcc and run against the library with host buffers (tests/test_c_harness_gpu.py).
#
# Acceptance script -- test/runtests.jl:40-107 of the reference, unchanged but for the import line:
#
#     using CloudAtlas, LinearAlgebra, Test
#     import CloudAtlasB200: ODEModel, hookstepsolve
#     sx, sy, sz, tx, tz = halfbox_symmetries()
#     α, γ = 1.0, 2.0; H = [sx*sy*sz, sz*tx*tz]; J, K, L = 1, 1, 3; R = 200.0
#     xeqb = [0.16764501289529937, -0.018552429231415948, 0.48640800040772025, ...]      # runtests.jl:49-53
#     model = ODEModel(α, γ, J, K, L, H)
#     @test norm(model.f(xeqb, R))/norm(xeqb) < 1e-10                                     # runtests.jl:57
#     xguess = [0.105, -0.0539, 0.388, ...]                                               # runtests.jl:70-71
#     xsoln, success = hookstepsolve(x -> model.f(x,R), x -> model.Df(x,R), xguess)       # runtests.jl:78-80 (closures)
#     @test norm(xsoln - xeqb)/norm(xeqb) < 1e-08
#     xsoln, success = hookstepsolve(model, R, xguess)                                    # same search, device resident
#     @test norm(xsoln - xeqb)/norm(xeqb) < 1e-08
#     model = ODEModel(α, γ, J, K, L, H, normalize=true)                                  # runtests.jl:90
#     xsoln, success = hookstepsolve(model, R, xguess_normalized, SearchParams(; δ=0.01)) # runtests.jl:91-105
#     @test norm(xsoln - xeqb_normalized)/norm(xeqb_normalized) < 1e-08
module CloudAtlasB200

using CloudAtlas
using LinearAlgebra
import CloudAtlas: derivative, SearchParams, Symmetry, SparseBilinear, BasisFunction, basisSet

const lib = get(ENV, "CLOUDATLAS_B200_LIB", joinpath(@__DIR__, "..", "lib", "libcloudatlas_b200.so"))

check(status::Int32) = status == 0 || error(unsafe_string(ccall((:ca_last_error, lib), Cstring, ())))

# ---- Symmetry / basisIndices (src/Symmetries.jl:3-14, src/BasisFunctions.jl:566-574) -------------------------
struct CSymmetry
    sx::Int32; sy::Int32; sz::Int32; ax2::Int32; az2::Int32; s::Int32
end
CSymmetry(σ::Symmetry) = CSymmetry(σ.sx, σ.sy, σ.sz, Int32(2σ.ax), Int32(2σ.az), σ.s)

function b200_basisIndices(J::Int, K::Int, L::Int, H::Vector{Symmetry}=Symmetry[])
    cH = CSymmetry.(H)
    m = Ref{Int64}(0)
    check(ccall((:ca_basis_indices, lib), Int32, (Int32, Int32, Int32, Ptr{CSymmetry}, Int32, Ptr{Int64}, Ref{Int64}),
                J, K, L, cH, length(cH), C_NULL, m))
    ijkl = Matrix{Int64}(undef, m[], 4)
    check(ccall((:ca_basis_indices, lib), Int32, (Int32, Int32, Int32, Ptr{CSymmetry}, Int32, Ptr{Int64}, Ref{Int64}),
                J, K, L, cH, length(cH), ijkl, m))
    ijkl
end

# page-lock a Julia array for the duration of a call: the host-pointer entry points overlap copies with kernels only
# from page-locked memory (bench.py's e2e figure is measured that way)
function with_pinned(fn, arrays::Array{Float64}...)
    for a in arrays
        check(ccall((:ca_pin, lib), Int32, (Ptr{Cvoid}, Int64), a, sizeof(a)))
    end
    try
        GC.@preserve arrays fn()
    finally
        for a in arrays
            ccall((:ca_unpin, lib), Int32, (Ptr{Cvoid},), a)
        end
    end
end

# ---- SparseBilinear (src/SparseBilinear.jl) ------------------------------------------------------------------
"SparseBilinear{Float64} evaluated on the GPU: same fields (ijk, val, m), same call syntax"
mutable struct B200Bilinear <: Function
    ijk::Matrix{Int64}
    val::Vector{Float64}
    m::Int
    handle::Ptr{Cvoid}
    function B200Bilinear(ijk::Matrix{Int64}, val::Vector{Float64}, m::Int)
        size(ijk, 2) == 3 || error("ijk must have three columns")                         # SparseBilinear.jl:8
        size(ijk, 1) == length(val) || error("ijk and val must have same number of rows") # SparseBilinear.jl:9
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ca_bilinear_create, lib), Int32, (Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    ijk, val, length(val), m, h))
        obj = new(ijk, val, m, h[])
        finalizer(o -> ccall((:ca_bilinear_destroy, lib), Int32, (Ptr{Cvoid},), o.handle), obj)
    end
end
B200Bilinear(N::SparseBilinear{Float64}) = B200Bilinear(Matrix{Int64}(N.ijk), N.val, N.m)

function (N::B200Bilinear)(x::Vector{Float64}, y::Vector{Float64})    # SparseBilinear.jl:56-65
    out = Vector{Float64}(undef, N.m)
    check(ccall((:ca_bilinear_eval, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), N.handle, x, y, out))
    out
end
(N::B200Bilinear)(x::Vector{Float64}) = N(x, x)                        # SparseBilinear.jl:70

function (N::B200Bilinear)(X::Matrix{Float64})                         # ensemble: nstates × m, one state per row
    size(X, 2) == N.m || error("ensemble has $(size(X,2)) columns, expected $(N.m)")
    OUT = similar(X)
    with_pinned(X, OUT) do
        check(ccall((:ca_bilinear_eval_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                    N.handle, X, size(X, 1), OUT))
    end
    OUT
end

function derivative(N::B200Bilinear, x::Vector{Float64})               # SparseBilinear.jl:75-91
    DN = Matrix{Float64}(undef, N.m, N.m)
    check(ccall((:ca_bilinear_jacobian, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), N.handle, x, DN))
    DN
end

function derivative(N::B200Bilinear, X::Matrix{Float64})               # ensemble: DN[s,:,:] = derivative(N, X[s,:])
    size(X, 2) == N.m || error("ensemble has $(size(X,2)) columns, expected $(N.m)")
    DN = Array{Float64,3}(undef, size(X, 1), N.m, N.m)
    check(ccall((:ca_bilinear_jacobian_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                N.handle, X, size(X, 1), DN))
    DN
end

# ---- ODEModel (src/ODEModels.jl:6-63) ------------------------------------------------------------------------
"the reference's ODEModel struct, field for field (src/ODEModels.jl:6-20), backed by device handles"
mutable struct B200Model{TB, TF, TDF}
    α::Float64                           # streamwise wavenumber
    γ::Float64                           # spanwise wavenumber
    H::Vector{Symmetry}                  # symmetry subgroup
    ijkl::Matrix{Int}                    # allowed i,j,k,l indices
    Ψ::Vector{BasisFunction{Float64}}    # basis set (CloudAtlas.basisSet: symbolic, host)
    B::Matrix{Float64}                   # inner product matrix
    A1::Matrix{Float64}                  # linear term from the base flow
    A2::Matrix{Float64}                  # linear term from the Laplacian
    N::B200Bilinear                      # nonlinear term, evaluated on the GPU
    Bfact::TB                            # lu(B)
    f::TF                                # f(x,R), f(X::Matrix,R)
    Df::TDF                              # Df(x,R), Df(X::Matrix,R)
    handles::Vector{Ptr{Cvoid}}          # one ca_model per GPU (replicas)
    comm::Ptr{Cvoid}                     # ca_comm of a multi-GPU model, C_NULL otherwise
end
Base.length(model::B200Model) = length(model.Ψ)                         # ODEModels.jl:63

function _closures(handles::Vector{Ptr{Cvoid}}, m::Int)
    handle = handles[1]
    function f(x::Vector{Float64}, R)                                  # ODEModels.jl:57
        out = Vector{Float64}(undef, m)
        check(ccall((:ca_model_rhs, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Float64}), handle, x, R, out))
        out
    end
    function f(X::Matrix{Float64}, R)                                  # ensemble, nstates × m; states split over the GPUs
        size(X, 2) == m || error("ensemble has $(size(X,2)) columns, expected $m")
        OUT = similar(X)
        with_pinned(X, OUT) do
            if length(handles) == 1
                check(ccall((:ca_model_rhs_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}),
                            handle, X, size(X, 1), R, OUT))
            else
                check(ccall((:ca_models_rhs_batched, lib), Int32,
                            (Ptr{Ptr{Cvoid}}, Int32, Ptr{Float64}, Int64, Float64, Ptr{Float64}),
                            handles, length(handles), X, size(X, 1), R, OUT))
            end
        end
        OUT
    end
    function Df(x::Vector{Float64}, R)                                 # ODEModels.jl:58
        D = Matrix{Float64}(undef, m, m)
        check(ccall((:ca_model_jac, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Float64}), handle, x, R, D))
        D
    end
    function Df(X::Matrix{Float64}, R)                                 # ensemble: D[s,:,:] = Df(X[s,:], R)
        D = Array{Float64,3}(undef, size(X, 1), m, m)
        check(ccall((:ca_model_jac_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}),
                    handle, X, size(X, 1), R, D))
        D
    end
    f, Df
end

function _finish(α, γ, H, ijkl, Ψ, B, A1, A2, N::B200Bilinear, handles, comm)
    f, Df = _closures(handles, size(B, 1))
    obj = B200Model(α, γ, H, Matrix{Int}(ijkl), Ψ, B, A1, A2, N, lu(B), f, Df, handles, comm)
    finalizer(obj) do o
        for h in o.handles
            ccall((:ca_model_destroy, lib), Int32, (Ptr{Cvoid},), h)
        end
        o.comm == C_NULL || ccall((:ca_comm_destroy, lib), Int32, (Ptr{Cvoid},), o.comm)
    end
end

"move an already-built CloudAtlas.ODEModel to the GPU (its operators are uploaded and folded there)"
function B200Model(model::CloudAtlas.ODEModel{Float64})
    B, A1, A2 = Matrix(model.B), Matrix(model.A1), Matrix(model.A2)
    m = size(B, 1)
    N = B200Bilinear(model.N)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ca_model_create, lib), Int32,
                (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                B, A1, A2, m, N.ijk, N.val, length(N.val), h))
    shear = CloudAtlas.shear_function(model.Ψ)                           # ODEModels.jl:65-78: x -> 1 + dot(x, Ψshear)
    Ψshear = [shear(Float64.(1:m .== n)) - 1 for n in 1:m]
    D = dissipation_matrix(model)
    check(ccall((:ca_model_set_observables, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], Ψshear, D))
    _finish(model.α, model.γ, model.H, model.ijkl, model.Ψ, B, A1, A2, N, [h[]], C_NULL)
end

"D_ij of dissipation(x) = 1 + x'Dx (ODEModels.jl:87-137) from the reference's own function, by polarisation"
function dissipation_matrix(model)
    d = CloudAtlas.dissipation_function(model.Ψ)
    m = length(model.Ψ)
    e(n) = Float64.(1:m .== n)
    diag = [d(e(n)) - 1 for n in 1:m]
    [i == j ? diag[i] : (d(e(i) + e(j)) - 1 - diag[i] - diag[j]) / 2 for i in 1:m, j in 1:m]
end

"""
    ODEModel(α, γ, J, K, L, H; normalize=false, ngpus=1, host_coo=true)

The reference constructor (src/ODEModels.jl:22-61) with the Galerkin assembly, the B⁻¹ folding and the packing done on
the GPU(s).  `ngpus > 1`: rows of the operators are sharded by output mode over the GPUs of this box and exchanged with
one NCCL all-gather inside the library (no NCCL.jl needed); every GPU then holds the full model and ensembles are split
by state.  `host_coo=false` skips the download of `N.ijk` / `N.val` (9 GB at m ≈ 3000).
"""
function ODEModel(α::Float64, γ::Float64, J::Int, K::Int, L::Int, H::Vector{Symmetry}; normalize=false, ngpus::Int=1,
                  host_coo::Bool=true)
    ijkl = b200_basisIndices(J, K, L, H)
    m = size(ijkl, 1)
    println("J,K,L,m == $J,$K,$L,$m")                                    # as the reference (ODEModels.jl:25-26)
    println("(2J+1)(2K+1)(2L+1) + 1 == $((2J+1)*(2K+1)*(2L+1) + 1)")
    Ψ = basisSet(α, γ, Matrix{Int}(ijkl), normalize=normalize)           # host, symbolic: kept for plotting / printing
    comm = Ref{Ptr{Cvoid}}(C_NULL)
    assemblies = Vector{Ptr{Cvoid}}(undef, ngpus)
    if ngpus == 1
        a = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ca_assemble, lib), Int32, (Float64, Float64, Ptr{Int64}, Int64, Int32, Int64, Int64, Ref{Ptr{Cvoid}}),
                    α, γ, ijkl, m, normalize, 0, m, a))
        assemblies[1] = a[]
    else
        check(ccall((:ca_comm_init_all, lib), Int32, (Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}), ngpus, C_NULL, comm))
        check(ccall((:ca_assemble_sharded, lib), Int32,
                    (Ptr{Cvoid}, Float64, Float64, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Int32, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
                    comm[], α, γ, ijkl, m, normalize, C_NULL, 0, 0, 0, 0, assemblies))
    end
    handles = Vector{Ptr{Cvoid}}(undef, ngpus)
    for d in 1:ngpus                                                     # the model is built on the assembly's device
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ca_model_create_from_assembly, lib), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), assemblies[d], h))
        handles[d] = h[]
    end
    nnz = Ref{Int64}(0)
    check(ccall((:ca_assembly_info, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ptr{Float64}),
                assemblies[1], C_NULL, C_NULL, C_NULL, nnz, C_NULL))
    B, A1, A2 = zeros(m, m), zeros(m, m), zeros(m, m)
    check(ccall((:ca_assembly_get_linear, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), assemblies[1], B, A1, A2))
    ijk, val = Matrix{Int64}(undef, host_coo ? nnz[] : 0, 3), Vector{Float64}(undef, host_coo ? nnz[] : 0)
    host_coo && check(ccall((:ca_assembly_get_coo, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}), assemblies[1], ijk, val))
    for a in assemblies
        ccall((:ca_assembly_destroy, lib), Int32, (Ptr{Cvoid},), a)
    end
    ccall((:ca_set_device, lib), Int32, (Int32,), 0)
    _finish(α, γ, H, ijkl, Ψ, B, A1, A2, B200Bilinear(ijk, val, m), handles, comm[])
end

"cnab2(x₀, model, R, Δt, Nsteps, nsave) for one trajectory (Vector) or an ensemble (Matrix, one state per row) -- ODEModels.jl:147"
function CloudAtlas.cnab2(x0::VecOrMat{Float64}, model::B200Model, R, Δt, Nsteps, nsave)
    X0 = x0 isa Vector ? reshape(x0, 1, :) : x0
    ns, Nsave = size(X0, 1), Nsteps ÷ nsave
    Xsave = Array{Float64}(undef, ns, length(model), Nsave)
    check(ccall((:ca_model_cnab2_batched, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                model.handles[1], X0, ns, R, Δt, Nsteps, nsave, Xsave, C_NULL))
    t = (0:Nsave-1) * Δt                                   # as the reference (ODEModels.jl:162)
    x0 isa Vector ? (t, permutedims(Xsave[1, :, :])) : (t, Xsave)
end

"(shear, dissipation) per state -- shear_function / dissipation_function, ODEModels.jl:65-137 (installed by the constructors)"
function observables(model::B200Model, X::Matrix{Float64})
    OBS = Matrix{Float64}(undef, size(X, 1), 2)
    check(ccall((:ca_model_observables_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                model.handles[1], X, size(X, 1), OBS))
    OBS
end

# ---- Newton-hookstep (src/Hookstep.jl) -----------------------------------------------------------------------
struct CSearchParams
    ftol::Float64; xtol::Float64; δ::Float64
    Nnewton::Int32; Nhook::Int32; Nmusearch::Int32; verbosity::Int32
end
CSearchParams(p::SearchParams) = CSearchParams(p.ftol, p.xtol, p.δ, p.Nnewton, p.Nhook, p.Nmusearch, p.verbosity)

"device-resident hookstepsolve for a model (any size) at Reynolds number R; returns (x, success) like Hookstep.jl:109"
function hookstepsolve(model::B200Model, R::Real, xguess::Vector{Float64}, params=SearchParams())
    x = similar(xguess)
    success = Ref{Int32}(0)
    check(ccall((:ca_hookstep_solve_model, lib), Int32,
                (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}, Int64, Ref{CSearchParams}, Ptr{Float64}, Ref{Int32}, Ptr{Cvoid}),
                model.handles[1], Float64(R), xguess, 1, CSearchParams(params), x, success, C_NULL))
    x, success[] == 1
end
# the closure forms keep the reference's meaning (Hookstep.jl:95,109): with x -> model.f(x,R), x -> model.Df(x,R)
# every evaluation runs on the GPU and the dense solves in Julia
hookstepsolve(f, Df, xguess, params=SearchParams()) = CloudAtlas.hookstepsolve(f, Df, xguess, params)
hookstepsolve(f, xguess, params::SearchParams=SearchParams()) = CloudAtlas.hookstepsolve(f, xguess, params)

# ---- Newton-Krylov (additive; BASELINE configs[4]) -----------------------------------------------------------
struct CKrylovParams
    ftol::Float64; gmres_rtol::Float64
    max_newton::Int32; restart::Int32; max_restarts::Int32; verbosity::Int32; precondition::Int32
end
struct CKrylovInfo
    converged::Int32; newton_steps::Int32; jvps::Int32; f_evals::Int32; world::Int32
    final_residual::Float64; device_seconds::Float64
    residual::NTuple{64,Float64}
end

"""
    newton_krylov(model, R, xguess; ftol=1e-10, gmres_rtol=1e-8, max_newton=30, restart=60, max_restarts=5)

Matrix-free Newton-GMRES on the device (preconditioned with L1 + L2/R); with a multi-GPU model the Jacobian action is
sharded by row slab over the GPUs.  Returns (x, converged, jacobian_actions).
"""
function newton_krylov(model::B200Model, R::Real, xguess::Vector{Float64}; ftol=1e-10, gmres_rtol=1e-8, max_newton=30,
                       restart=60, max_restarts=5)
    x = similar(xguess)
    success = Ref{Int32}(0)
    info = Ref{CKrylovInfo}()
    p = CKrylovParams(ftol, gmres_rtol, max_newton, restart, max_restarts, 0, 1)
    check(ccall((:ca_newton_krylov, lib), Int32,
                (Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{CKrylovParams}, Ptr{Float64}, Ref{Int32}, Ref{CKrylovInfo}),
                model.handles, length(model.handles), model.comm, Float64(R), xguess, p, x, success, info))
    x, success[] == 1, Int(info[].jvps)
end

export B200Bilinear, B200Model, b200_basisIndices, observables, newton_krylov

end # module

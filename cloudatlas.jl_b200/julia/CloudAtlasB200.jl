# CloudAtlasB200.jl -- drop-in Julia binding of libcloudatlas_b200.so for CloudAtlas.jl's hot path.
#
# `using CloudAtlas, CloudAtlasB200` keeps the package's API (src/CloudAtlas.jl:16-32) and adds GPU methods:
#   B200Bilinear(N::SparseBilinear{Float64})            N(x), N(x,y), derivative(N,x), N(X::Matrix)
#   B200Model(model::ODEModel{Float64}) / B200Model(α,γ,J,K,L,H; normalize)      f, Df, ensemble f
#   hookstepsolve(m::B200Model, R, xguess, params)      device-resident search, returns (x, success)
# Only Float64 goes to the library; Rational / Dual element types stay on CloudAtlas.jl's generic methods.
# NOTE: Julia is not installed in the build image; this file is exercised through the identical C ABI by the
# Python mirror (tests/), not executed there.  Every ccall below matches include/cloudatlas_b200.h.
module CloudAtlasB200

using CloudAtlas
import CloudAtlas: derivative, hookstepsolve, SearchParams, Symmetry, ODEModel, SparseBilinear

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

# ---- SparseBilinear (src/SparseBilinear.jl) ------------------------------------------------------------------
mutable struct B200Bilinear
    handle::Ptr{Cvoid}
    m::Int
    function B200Bilinear(N::SparseBilinear{Float64})
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ca_bilinear_create, lib), Int32, (Ptr{Int64}, Ptr{Float64}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    N.ijk, N.val, length(N.val), N.m, h))
        obj = new(h[], N.m)
        finalizer(o -> ccall((:ca_bilinear_destroy, lib), Int32, (Ptr{Cvoid},), o.handle), obj)
    end
end

function (N::B200Bilinear)(x::Vector{Float64}, y::Vector{Float64})    # SparseBilinear.jl:56-65
    out = Vector{Float64}(undef, N.m)
    check(ccall((:ca_bilinear_eval, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), N.handle, x, y, out))
    out
end
(N::B200Bilinear)(x::Vector{Float64}) = N(x, x)                        # SparseBilinear.jl:70

function (N::B200Bilinear)(X::Matrix{Float64})                         # ensemble: nstates × m, one state per row
    size(X, 2) == N.m || error("ensemble has $(size(X,2)) columns, expected $(N.m)")
    OUT = similar(X)
    check(ccall((:ca_bilinear_eval_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                N.handle, X, size(X, 1), OUT))
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
mutable struct B200Model
    handle::Ptr{Cvoid}
    m::Int
    ijkl::Matrix{Int64}
    B::Matrix{Float64}; A1::Matrix{Float64}; A2::Matrix{Float64}
    N::SparseBilinear{Float64}
    f::Function          # f(x,R) and f(X::Matrix,R)
    Df::Function
end

function _model_from_operators(ijkl, B, A1, A2, N::SparseBilinear{Float64})
    m = size(B, 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ca_model_create, lib), Int32,
                (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                B, A1, A2, m, N.ijk, N.val, length(N.val), h))
    handle = h[]
    function f(x::Vector{Float64}, R)                                  # ODEModels.jl:57
        out = Vector{Float64}(undef, m)
        check(ccall((:ca_model_rhs, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Float64}), handle, x, R, out))
        out
    end
    function f(X::Matrix{Float64}, R)                                  # ensemble, nstates × m
        OUT = similar(X)
        check(ccall((:ca_model_rhs_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Ptr{Float64}),
                    handle, X, size(X, 1), R, OUT))
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
    obj = B200Model(handle, m, ijkl, B, A1, A2, N, f, Df)
    finalizer(o -> ccall((:ca_model_destroy, lib), Int32, (Ptr{Cvoid},), o.handle), obj)
end

"move an already-built CloudAtlas.ODEModel to the GPU"
B200Model(model::ODEModel{Float64}) =
    _model_from_operators(model.ijkl, Matrix(model.B), Matrix(model.A1), Matrix(model.A2), model.N)

"ODEModel(α,γ,J,K,L,H; normalize) with the Galerkin assembly done on the GPU (ODEModels.jl:22-61)"
function B200Model(α::Float64, γ::Float64, J::Int, K::Int, L::Int, H::Vector{Symmetry}; normalize=false)
    ijkl = b200_basisIndices(J, K, L, H)
    m = size(ijkl, 1)
    a = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ca_assemble, lib), Int32, (Float64, Float64, Ptr{Int64}, Int64, Int32, Int64, Int64, Ref{Ptr{Cvoid}}),
                α, γ, ijkl, m, normalize, 0, m, a))
    nnz = Ref{Int64}(0)
    check(ccall((:ca_assembly_info, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ptr{Float64}),
                a[], C_NULL, C_NULL, C_NULL, nnz, C_NULL))
    B, A1, A2 = zeros(m, m), zeros(m, m), zeros(m, m)
    ijk, val = Matrix{Int64}(undef, nnz[], 3), Vector{Float64}(undef, nnz[])
    check(ccall((:ca_assembly_get_linear, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), a[], B, A1, A2))
    check(ccall((:ca_assembly_get_coo, lib), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}), a[], ijk, val))
    ccall((:ca_assembly_destroy, lib), Int32, (Ptr{Cvoid},), a[])
    _model_from_operators(ijkl, B, A1, A2, SparseBilinear(ijk, val, m))
end

"cnab2(x₀, model, R, Δt, Nsteps, nsave) for one trajectory (Vector) or an ensemble (Matrix, one state per row) -- ODEModels.jl:147"
function CloudAtlas.cnab2(x0::VecOrMat{Float64}, model::B200Model, R, Δt, Nsteps, nsave)
    X0 = x0 isa Vector ? reshape(x0, 1, :) : x0
    ns, Nsave = size(X0, 1), Nsteps ÷ nsave
    Xsave = Array{Float64}(undef, ns, model.m, Nsave)
    check(ccall((:ca_model_cnab2_batched, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Float64, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                model.handle, X0, ns, R, Δt, Nsteps, nsave, Xsave, C_NULL))
    t = (0:Nsave-1) * Δt                                   # as the reference (ODEModels.jl:162)
    x0 isa Vector ? (t, permutedims(Xsave[1, :, :])) : (t, Xsave)
end

"(shear, dissipation) per state -- shear_function / dissipation_function, ODEModels.jl:65-137; needs ca_model_set_observables"
function observables(model::B200Model, X::Matrix{Float64})
    OBS = Matrix{Float64}(undef, size(X, 1), 2)
    check(ccall((:ca_model_observables_batched, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
                model.handle, X, size(X, 1), OBS))
    OBS
end

Base.length(model::B200Model) = model.m                                 # ODEModels.jl:63

# ---- Newton-hookstep (src/Hookstep.jl) -----------------------------------------------------------------------
struct CSearchParams
    ftol::Float64; xtol::Float64; δ::Float64
    Nnewton::Int32; Nhook::Int32; Nmusearch::Int32; verbosity::Int32
end
CSearchParams(p::SearchParams) = CSearchParams(p.ftol, p.xtol, p.δ, p.Nnewton, p.Nhook, p.Nmusearch, p.verbosity)

"device-resident hookstepsolve for a model at Reynolds number R; returns (x, success) like Hookstep.jl:109"
function hookstepsolve(model::B200Model, R::Real, xguess::Vector{Float64}, params=SearchParams())
    x = similar(xguess)
    success = Ref{Int32}(0)
    check(ccall((:ca_hookstep_solve_model, lib), Int32,
                (Ptr{Cvoid}, Ref{Float64}, Ptr{Float64}, Int64, Ref{CSearchParams}, Ptr{Float64}, Ref{Int32}, Ptr{Cvoid}),
                model.handle, Float64(R), xguess, 1, CSearchParams(params), x, success, C_NULL))
    x, success[] == 1
end
# the closure form hookstepsolve(f, Df, xguess, params) is CloudAtlas.jl's own (Hookstep.jl:109); pass
# x -> model.f(x,R), x -> model.Df(x,R) to run its evaluations on the GPU.

export B200Bilinear, B200Model, b200_basisIndices

end # module

"""The Julia binding (cloudatlas.jl_b200/julia/CloudAtlasB200.jl) cannot run here (no Julia in the image), so its
ccalls are checked statically against the C header: every symbol exists in the library, and the argument count and
the Julia argument types agree with the C prototype."""
import os
import re

import __graft_entry__ as g

ca = g.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "cloudatlas.jl_b200", "julia", "CloudAtlasB200.jl")
HDR = os.path.join(ROOT, "include", "cloudatlas_b200.h")


def c_prototypes():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|const char\*|void)\s+(ca_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        protos[name] = (ret, params)
    return protos


def julia_type_of(cparam):
    t = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", cparam).strip()      # drop the parameter name
    t = t.replace("const ", "").replace(" ", "")
    table = {"double*": {"Ptr{Float64}", "Ref{Float64}"}, "double": {"Float64"}, "int64_t": {"Int64"},
             "int32_t": {"Int32"}, "int64_t*": {"Ptr{Int64}", "Ref{Int64}"}, "int32_t*": {"Ptr{Int32}", "Ref{Int32}"},
             "void*": {"Ptr{Cvoid}"}, "char*": {"Cstring", "Ptr{UInt8}"}, "char**": {"Ptr{Cstring}", "Ref{Cstring}", "Ptr{Ptr{UInt8}}"}}
    if t in table:
        return table[t]
    if re.fullmatch(r"ca_[a-z_]+\*\*", t):
        return {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"}
    if re.fullmatch(r"ca_[a-z_]+\*", t):                               # opaque handles and plain structs by pointer
        return {"Ptr{Cvoid}", "Ref{SearchParams}", "Ref{CSearchParams}", "Ptr{CSearchParams}", "Ref{CSymmetry}",
                "Ptr{CSymmetry}", "Ptr{Nothing}", "Ref{CKrylovParams}", "Ref{CKrylovInfo}"}
    return None


def julia_ccalls():
    src = open(JL).read()
    calls = []
    for name, ret, args in re.findall(r"ccall\(\(:(ca_[a-z0-9_]+),\s*lib\),\s*([A-Za-z0-9{}]+),\s*\(([^)]*)\)", src):
        types = [a.strip() for a in args.split(",") if a.strip()]
        calls.append((name, ret, types))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 20
    problems = []
    for name, ret, types in calls:
        if not hasattr(ca.lib, name):
            problems.append(f"{name}: not exported by the library")
            continue
        if name not in protos:
            problems.append(f"{name}: not declared in the header")
            continue
        cret, cparams = protos[name]
        want_ret = {"int32_t": "Int32", "const char*": "Cstring", "void": "Cvoid"}[cret]
        if ret != want_ret:
            problems.append(f"{name}: returns {ret}, header says {cret}")
        if len(types) != len(cparams):
            problems.append(f"{name}: {len(types)} arguments in the ccall, {len(cparams)} in the header")
            continue
        for pos, (jt, cp) in enumerate(zip(types, cparams)):
            ok = julia_type_of(cp)
            if ok is None:
                problems.append(f"{name}: argument {pos + 1} '{cp}' has no Julia mapping in this test")
            elif jt not in ok:
                problems.append(f"{name}: argument {pos + 1} is {jt}, header has '{cp}'")
    assert not problems, "\n".join(problems)


def test_model_struct_has_the_reference_fields_in_order():
    """src/ODEModels.jl:6-20: struct ODEModel has the fields α, γ, H, ijkl, Ψ, B, A1, A2, N, Bfact, f, Df.  The binding's
    B200Model must start with exactly these, in this order (positional construction and field access in the notebooks)."""
    src = open(JL).read()
    body = re.search(r"mutable struct B200Model\{[^}]*\}\n(.*?)\nend", src, flags=re.S).group(1)
    fields = [re.match(r"\s*([^\s:#]+)::", line).group(1) for line in body.split("\n") if re.match(r"\s*[^\s:#]+::", line)]
    assert fields[:12] == ["α", "γ", "H", "ijkl", "Ψ", "B", "A1", "A2", "N", "Bfact", "f", "Df"]
    # N is callable and carries SparseBilinear's fields (src/SparseBilinear.jl:3-6)
    nb = re.search(r"mutable struct B200Bilinear <: Function\n(.*?)\n    function", src, flags=re.S).group(1)
    nfields = [re.match(r"\s*([a-z]+)::", line).group(1) for line in nb.split("\n") if re.match(r"\s*[a-z]+::", line)]
    assert nfields[:3] == ["ijk", "val", "m"]
    # the reference's entry points the notebooks call are all defined for the binding's types
    for needle in ("function ODEModel(α::Float64, γ::Float64, J::Int, K::Int, L::Int, H::Vector{Symmetry}; normalize=false",
                   "function hookstepsolve(model::B200Model, R::Real, xguess::Vector{Float64}, params=SearchParams())",
                   "Base.length(model::B200Model)", "function derivative(N::B200Bilinear, x::Vector{Float64})",
                   "function CloudAtlas.cnab2(", "ca_pin", "ca_assemble_sharded", "ca_comm_init_all"):
        assert needle in src, needle

# Julia shim: keeps TensorNetworkAD.jl's API (src/TensorNetworkAD.jl:6-10) and Zygote rules, forwards the
# bodies of the hot path to libtnad_b200.so through `ccall`.  NOT EXECUTED in the build image (no Julia);
# the identical C ABI is exercised by the ctypes layer in tensornetworkad.jl_b200/.
#
# Usage inside the reference package: `include("TensorNetworkAD_b200.jl")` after the existing includes; the
# methods below replace `trg`, `ctmrg`, `energy` and add Zygote adjoints for them (the rules in
# src/autodiff.jl stay as they are and keep serving user code that composes the pieces by hand).
module TNADB200

using Zygote
const libtnad = get(ENV, "TNAD_B200_LIB", "libtnad_b200.so")

mutable struct Ctx
    h::Ptr{Cvoid}
end
const _ctx = Ref{Union{Nothing,Ctx}}(nothing)

function ctx()
    if _ctx[] === nothing
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:tnad_create, libtnad), Cint, (Cint, Ref{Ptr{Cvoid}}), parse(Int, get(ENV, "TNAD_DEVICE", "0")), h)
        rc == 0 || error(unsafe_string(ccall((:tnad_last_error, libtnad), Cstring, (Ptr{Cvoid},), C_NULL)))
        c = Ctx(h[])
        finalizer(c -> ccall((:tnad_destroy, libtnad), Cint, (Ptr{Cvoid},), c.h), c)
        _ctx[] = c
    end
    _ctx[].h
end

function check(rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:tnad_last_error, libtnad), Cstring, (Ptr{Cvoid},), ctx()))
    rc == 1 ? throw(DimensionMismatch(msg)) : error("tnad error $rc: $msg")
end

# ---- trg(a, χ, niter; tol)  (src/trg.jl:13-30) ---------------------------------------------------------
function trg_forward(a::Array{Float64,4}, χ::Integer, niter::Integer, tol::Float64, want_tape::Bool)
    lnZ = Ref{Cdouble}(0.0)
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve a check(ccall((:tnad_trg_forward, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Cint, Cdouble, Ref{Cdouble}, Ptr{Ptr{Cvoid}}),
        ctx(), a, size(a, 1), size(a, 2), χ, niter, tol, lnZ, want_tape ? tape : C_NULL))
    lnZ[], tape[]
end

trg(a::Array{Float64,4}, χ, niter; tol::Float64 = 1e-16) = trg_forward(a, χ, niter, tol, false)[1]

Zygote.@adjoint function trg(a::Array{Float64,4}, χ, niter; tol::Float64 = 1e-16)
    lnZ, tape = trg_forward(a, χ, niter, tol, true)
    function back(Δ)
        ā = similar(a)
        GC.@preserve ā check(ccall((:tnad_trg_backward, libtnad), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}),
                                   ctx(), tape, Float64(Δ), ā))
        ccall((:tnad_tape_free, libtnad), Cint, (Ptr{Cvoid},), tape)
        (ā, nothing, nothing)               # Zygote then chains through model_tensor (README.md:67-70)
    end
    lnZ, back
end

# ---- ctmrg(rt; tol, maxit)  (src/ctmrg.jl:110-117) ------------------------------------------------------
function ctmrg_arrays(bulk::Array{Float64,4}, corner::Matrix{Float64}, edge::Array{Float64,3}, tol, maxit, want_tape)
    D, χ = size(bulk, 1), size(corner, 1)
    c, e = copy(corner), copy(edge)
    steps = Ref{Cint}(0); vals = Vector{Float64}(undef, χ * D); tape = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bulk c e vals check(ccall((:tnad_ctmrg, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ref{Cint}, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}),
        ctx(), bulk, D, χ, c, e, Float64(tol), maxit, steps, vals, want_tape ? tape : C_NULL))
    c, e, tape[]
end

# ctmrg(rt::CTMRGRuntime; tol, maxit) = SquareCTMRGRuntime(rt.bulk, ctmrg_arrays(rt.bulk, rt.corner, rt.edge, tol, maxit, false)[1:2]...)
#
# Zygote.@adjoint ctmrg(rt; tol, maxit): forward with want_tape = true; pullback
#   Δ -> begin
#       b̄ = similar(rt.bulk); c̄0 = similar(rt.corner); ē0 = similar(rt.edge)
#       ccall((:tnad_ctmrg_backward, libtnad), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
#             ctx(), tape, Δ.corner, Δ.edge, b̄, c̄0, ē0)
#       ((bulk = b̄, corner = c̄0, edge = ē0),)          # NamedTuple cotangent of the struct (autodiff.jl:11-14)
#   end

# ---- energy(h, ipeps; χ, tol, maxit)  (src/variationalipeps.jl:28-40) -------------------------------------
function energy_call(h::Array{Float64,4}, A::Array{Float64,5}, χ, tol, maxit, want_grad::Bool)
    size(A, 1) == size(A, 2) == size(A, 3) == size(A, 4) ||
        throw(DimensionMismatch("size of tensor error, should be `(d, d, d, d, s)`, got $(size(A))."))
    e = Ref{Cdouble}(0.0); g = want_grad ? similar(A) : A
    GC.@preserve h A g check(ccall((:tnad_energy, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint, Cdouble, Cint, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
        ctx(), h, A, size(A, 1), size(A, 5), χ, Float64(tol), maxit, e, want_grad ? pointer(g) : C_NULL, C_NULL))
    e[], (want_grad ? g : nothing)
end

# energy(h, ipeps::IPEPS; χ, tol, maxit) = energy_call(h, ipeps.bulk, χ, tol, maxit, false)[1]
#
# Zygote.@adjoint energy(h, ipeps; χ, tol, maxit):
#   e, Ā = energy_call(h, ipeps.bulk, χ, tol, maxit, true)
#   e, Δ -> (nothing, (bulk = Δ .* Ā,))      # test/variationalipeps.jl:126-128 reads `.bulk`
#
# optimiseipeps (src/variationalipeps.jl:67-75) is unchanged: Optim.jl calls `energy` and `Zygote.gradient(energy, x)`.

end # module

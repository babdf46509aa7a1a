# Julia shim: keeps TensorNetworkAD.jl's API (src/TensorNetworkAD.jl:6-10) and Zygote rules, forwards the
# bodies of the hot path to libtnad_b200.so through `ccall`.  NOT EXECUTED in the build image (no Julia);
# the identical C ABI is exercised by the ctypes layer in tensornetworkad.jl_b200/ (every call below has a
# line-by-line twin in tensornetworkad.jl_b200/_lib.py, which the GPU tests run).
#
# Usage inside the reference package, after its own includes (src/TensorNetworkAD.jl:12-18):
#
#     include("TensorNetworkAD_b200.jl")      # adds the methods below to TensorNetworkAD
#
# It defines, on the reference's own types,
#     trg(a, χ, niter; tol)                      (src/trg.jl:13-30)             + Zygote.@adjoint
#     ctmrg(rt::CTMRGRuntime; tol, maxit)        (src/ctmrg.jl:110-117)         + Zygote.@adjoint
#     energy(h, ipeps::IPEPS; χ, tol, maxit)     (src/variationalipeps.jl:28-40)+ Zygote.@adjoint
#     magnetisation(model, β, χ)                 (src/exampletensors.jl:57-69)  (differentiable through ctmrg's adjoint)
# for Float64 arrays (the complex methods of the reference stay as they are).  `optimiseipeps`
# (src/variationalipeps.jl:67-75), the struct adjoints and the `norm` rule of src/autodiff.jl are untouched: Optim.jl
# keeps calling `energy` and `Zygote.gradient`.

using Zygote
using LinearAlgebra

const libtnad = get(ENV, "TNAD_B200_LIB", "libtnad_b200.so")

# ---- context: one device + its streams, created on first use ------------------------------------------------
mutable struct TnadCtx
    h::Ptr{Cvoid}
end
const _tnad_ctx = Ref{Union{Nothing,TnadCtx}}(nothing)

function tnad_ctx()
    if _tnad_ctx[] === nothing
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:tnad_create, libtnad), Cint, (Cint, Ref{Ptr{Cvoid}}), parse(Int, get(ENV, "TNAD_DEVICE", "0")), h)
        rc == 0 || error(unsafe_string(ccall((:tnad_last_error, libtnad), Cstring, (Ptr{Cvoid},), C_NULL)))
        _tnad_ctx[] = TnadCtx(h[])      # lives as long as the session: tapes must not outlive their context
    end
    _tnad_ctx[].h
end

function tnad_check(rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:tnad_last_error, libtnad), Cstring, (Ptr{Cvoid},), tnad_ctx()))
    rc == 1 ? throw(DimensionMismatch(msg)) : error("tnad error $rc: $msg")
end

# ---- tapes: device-resident record of a forward pass.  Owned by a mutable struct whose finalizer frees it, so a
#      pullback may be called any number of times (Zygote.jacobian, repeated back(Δ)) and an unused pullback leaks nothing.
mutable struct TnadTape
    h::Ptr{Cvoid}
    function TnadTape(h::Ptr{Cvoid})
        t = new(h)
        finalizer(t) do x
            x.h == C_NULL || ccall((:tnad_tape_free, libtnad), Cint, (Ptr{Cvoid},), x.h)
            x.h = C_NULL
        end
        t
    end
end

# ---- trg(a, χ, niter; tol)  (src/trg.jl:13-30) ---------------------------------------------------------------
function tnad_trg_forward(a::Array{Float64,4}, χ::Integer, niter::Integer, tol::Float64, want_tape::Bool)
    lnZ = Ref{Cdouble}(0.0)
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve a tnad_check(ccall((:tnad_trg_forward, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Cint, Cint, Cdouble, Ref{Cdouble}, Ptr{Ptr{Cvoid}}),
        tnad_ctx(), a, size(a, 1), size(a, 2), χ, niter, tol, lnZ, want_tape ? tape : C_NULL))
    lnZ[], (want_tape ? TnadTape(tape[]) : nothing)
end

trg(a::Array{Float64,4}, χ, niter; tol::Float64 = 1e-16) = tnad_trg_forward(a, χ, niter, tol, false)[1]

Zygote.@adjoint function trg(a::Array{Float64,4}, χ, niter; tol::Float64 = 1e-16)
    lnZ, tape = tnad_trg_forward(a, χ, niter, tol, true)
    function back(Δ)
        ā = similar(a)
        GC.@preserve ā tape tnad_check(ccall((:tnad_trg_backward, libtnad), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), tnad_ctx(), tape.h, Float64(Δ), ā))
        (ā, nothing, nothing)               # Zygote then chains through model_tensor (README.md:67-70)
    end
    lnZ, back
end

# ---- ctmrg(rt; tol, maxit)  (src/ctmrg.jl:110-117, fixedpoint.jl:11-41 inside the library) ---------------------
function tnad_ctmrg_arrays(bulk::Array{Float64,4}, corner::Matrix{Float64}, edge::Array{Float64,3}, tol, maxit, want_tape::Bool)
    D, χ = size(bulk, 1), size(corner, 1)
    c, e = copy(corner), copy(edge)                       # in/out arguments of the C call
    steps = Ref{Cint}(0)
    vals = Vector{Float64}(undef, χ * D)
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bulk c e vals tnad_check(ccall((:tnad_ctmrg, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ref{Cint}, Ptr{Cdouble}, Ptr{Ptr{Cvoid}}),
        tnad_ctx(), bulk, D, χ, c, e, Float64(tol), maxit, steps, vals, want_tape ? tape : C_NULL))
    c, e, (want_tape ? TnadTape(tape[]) : nothing)
end

function ctmrg(rt::CTMRGRuntime{LT,Float64}; tol, maxit) where LT
    c, e, _ = tnad_ctmrg_arrays(rt.bulk, rt.corner, rt.edge, tol, maxit, false)
    CTMRGRuntime{LT}(rt.bulk, c, e)
end

# The cotangent of a struct is a NamedTuple of its fields (src/autodiff.jl:11-14).  The library returns the cotangents
# of the INITIAL corner / edge as well: with the `@nograd` initialisation of the reference (autodiff.jl:5) Zygote drops
# them, with a user-supplied differentiable environment they flow on.
Zygote.@adjoint function ctmrg(rt::CTMRGRuntime{LT,Float64}; tol, maxit) where LT
    c, e, tape = tnad_ctmrg_arrays(rt.bulk, rt.corner, rt.edge, tol, maxit, true)
    function back(Δ)
        c̄ = (Δ === nothing || Δ.corner === nothing) ? zeros(size(c)) : Array{Float64}(Δ.corner)
        ē = (Δ === nothing || Δ.edge === nothing) ? zeros(size(e)) : Array{Float64}(Δ.edge)
        b̄ = similar(rt.bulk); c̄0 = similar(rt.corner); ē0 = similar(rt.edge)
        GC.@preserve c̄ ē b̄ c̄0 ē0 tape tnad_check(ccall((:tnad_ctmrg_backward, libtnad), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            tnad_ctx(), tape.h, c̄, ē, b̄, c̄0, ē0))
        if Δ !== nothing && Δ.bulk !== nothing
            b̄ .+= Δ.bulk                                  # the returned runtime carries the same bulk
        end
        ((bulk = b̄, corner = c̄0, edge = ē0),)
    end
    CTMRGRuntime{LT}(rt.bulk, c, e), back
end

# ---- energy(h, ipeps; χ, tol, maxit)  (src/variationalipeps.jl:28-40): value and gradient in one call ----------
function tnad_energy_call(h::Array{Float64,4}, A::Array{Float64,5}, χ, tol, maxit, want_grad::Bool)
    size(A, 1) == size(A, 2) == size(A, 3) == size(A, 4) ||
        throw(DimensionMismatch("size of tensor error, should be `(d, d, d, d, s)`, got $(size(A))."))   # src/ipeps.jl:19-20
    e = Ref{Cdouble}(0.0)
    g = want_grad ? similar(A) : A
    GC.@preserve h A g tnad_check(ccall((:tnad_energy, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint, Cdouble, Cint, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
        tnad_ctx(), h, A, size(A, 1), size(A, 5), χ, Float64(tol), maxit, e, want_grad ? pointer(g) : C_NULL, C_NULL))
    e[], (want_grad ? g : nothing)
end

energy(h::Array{Float64,4}, ipeps::IPEPS{LT,Float64}; χ::Int, tol::Real, maxit::Int) where LT =
    tnad_energy_call(h, ipeps.bulk, χ, tol, maxit, false)[1]

# `Zygote.gradient(x -> energy(h, x; ...), ipeps)[1]` is a NamedTuple with the field `bulk`
# (test/variationalipeps.jl:126-128); `optimiseipeps` reaches this rule through the IPEPS constructor adjoint
# (src/autodiff.jl:7-9), which unwraps `.bulk` again (src/variationalipeps.jl:71-73).
Zygote.@adjoint function energy(h::Array{Float64,4}, ipeps::IPEPS{LT,Float64}; χ::Int, tol::Real, maxit::Int) where LT
    e, Ā = tnad_energy_call(h, ipeps.bulk, χ, tol, maxit, true)
    e, Δ -> (nothing, (bulk = Δ .* Ā,))
end

# ---- magnetisation(model, β, χ)  (src/exampletensors.jl:57-69) --------------------------------------------------
# The read-out is five small einsums on χ x D x χ tensors; it stays the reference's own OMEinsum code and becomes
# differentiable through the `ctmrg` adjoint above (test/ctmrg.jl:44-46).  For callers that want it in one C call:
function tnad_magnetisation_readout(a::Array{Float64,4}, m::Array{Float64,4}, corner::Matrix{Float64}, edge::Array{Float64,3})
    mag = Ref{Cdouble}(0.0)
    GC.@preserve a m corner edge tnad_check(ccall((:tnad_magnetisation_readout, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ref{Cdouble}),
        tnad_ctx(), a, m, size(a, 1), corner, edge, size(corner, 1), mag))
    mag[]
end

Zygote.@adjoint function tnad_magnetisation_readout(a::Array{Float64,4}, m::Array{Float64,4}, corner::Matrix{Float64}, edge::Array{Float64,3})
    y = tnad_magnetisation_readout(a, m, corner, edge)
    function back(Δ)
        ā = similar(a); m̄ = similar(m); c̄ = similar(corner); ē = similar(edge)
        GC.@preserve a m corner edge ā m̄ c̄ ē tnad_check(ccall((:tnad_magnetisation_backward, libtnad), Cint,
            (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble,
             Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            tnad_ctx(), a, m, size(a, 1), corner, edge, size(corner, 1), Float64(Δ), ā, m̄, c̄, ē))
        (ā, m̄, c̄, ē)
    end
    y, back
end

# ---- independent instances over several GPUs (β sweeps, parameter scans): one C call ---------------------------
function trg_sweep(tensors::Vector{Array{Float64,4}}, χ::Integer, niter::Integer; tol::Float64 = 1e-16, ngpu::Integer = 1,
                   want_grad::Bool = false)
    ninst = length(tensors)
    d0, d1 = size(tensors[1], 1), size(tensors[1], 2)
    flat = reduce(vcat, vec.(tensors))
    lnZ = Vector{Float64}(undef, ninst)
    grads = want_grad ? similar(flat) : flat
    err = Vector{UInt8}(undef, 512)
    rc = GC.@preserve flat lnZ grads err ccall((:tnad_trg_sweep, libtnad), Cint,
        (Ptr{Cdouble}, Cint, Cint, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Cint),
        flat, ninst, d0, d1, χ, niter, tol, ngpu, C_NULL, lnZ, want_grad ? pointer(grads) : C_NULL, err, 512)
    rc == 0 || error("tnad_trg_sweep: " * unsafe_string(pointer(err)))
    want_grad ? (lnZ, [reshape(grads[(i - 1) * length(tensors[1]) + 1:i * length(tensors[1])], size(tensors[1])) for i in 1:ninst]) : lnZ
end

# ---- χ-sharded ctmrgstep over the ranks of an MPI job (one process per GPU) --------------------------------------
# rank 0 asks NCCL (dlopen'ed by the library) for a unique id, MPI carries the 128 bytes, every rank joins:
#     id = rank == 0 ? nccl_unique_id() : zeros(UInt8, 128);  MPI.Bcast!(id, 0, comm);  comm_init(id, rank, nranks)
function nccl_unique_id(nccl_path::Union{Nothing,String} = nothing)
    id = zeros(UInt8, 128)
    rc = ccall((:tnad_nccl_unique_id, libtnad), Cint, (Cstring, Ptr{UInt8}), something(nccl_path, C_NULL), id)
    rc == 0 || error("tnad_nccl_unique_id failed ($rc)")
    id
end
comm_init(id::Vector{UInt8}, rank::Integer, nranks::Integer; nccl_path::Union{Nothing,String} = nothing) =
    tnad_check(ccall((:tnad_comm_init, libtnad), Cint, (Ptr{Cvoid}, Cstring, Ptr{UInt8}, Cint, Cint),
                tnad_ctx(), something(nccl_path, C_NULL), id, rank, nranks))

# ctmrgstep((c, t, vals), (a, χ, D)) of src/ctmrg.jl:126-153 with the contractions sliced along χ over the ranks,
# ncclAllGather around svd(cpmat + cpmat') and a shared back-transformation; every rank passes and receives the full
# environment (a few MB)
function ctmrgstep_sharded(a::Array{Float64,4}, c::Matrix{Float64}, t::Array{Float64,3})
    D, χ = size(a, 1), size(c, 1)
    c′, t′, vals = similar(c), similar(t), Vector{Float64}(undef, χ * D)
    GC.@preserve a c t c′ t′ vals tnad_check(ccall((:tnad_ctmrgstep_sharded, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        tnad_ctx(), a, D, c, t, χ, c′, t′, vals, C_NULL))
    c′, t′, vals
end

# ctmrg(rt; tol, maxit) (src/ctmrg.jl:110-117) over the communicator: the fixed-point loop with the reference's stop rule
function ctmrg_sharded(a::Array{Float64,4}, c::Matrix{Float64}, t::Array{Float64,3}; tol::Real, maxit::Integer)
    D, χ = size(a, 1), size(c, 1)
    c′, t′, vals, steps = copy(c), copy(t), Vector{Float64}(undef, χ * D), Ref{Cint}(0)
    GC.@preserve a c′ t′ vals tnad_check(ccall((:tnad_ctmrg_sharded, libtnad), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ref{Cint}, Ptr{Cdouble}),
        tnad_ctx(), a, D, χ, c′, t′, Float64(tol), maxit, steps, vals))
    c′, t′, vals, Int(steps[])
end

# _initializect_square(bulk, Val(:random), χ) (src/ctmrg.jl:66-72) generated on the device from a seed
function initializect_random_device(D::Integer, χ::Integer, seed::Integer)
    c, t = Matrix{Float64}(undef, χ, χ), Array{Float64,3}(undef, χ, D, χ)
    GC.@preserve c t tnad_check(ccall((:tnad_ctmrg_init_random, libtnad), Cint,
        (Ptr{Cvoid}, Cint, Cint, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}), tnad_ctx(), D, χ, UInt64(seed), c, t))
    c, t
end

# A/B switches of the context (DESIGN.md 7a), e.g. tnad_set_option("TNAD_SHARDED_LOOP", "1") after comm_init
tnad_set_option(name::String, value::Union{Nothing,String}) =
    ccall((:tnad_set_option, libtnad), Cint, (Ptr{Cvoid}, Cstring, Cstring), tnad_ctx(), name, something(value, C_NULL))

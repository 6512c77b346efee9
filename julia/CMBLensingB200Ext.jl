# CMBLensingB200Ext.jl — the package extension a CMBLensing.jl maintainer would add next to ext/CMBLensingCUDAExt.jl
# (Project.toml: [extensions] CMBLensingB200Ext = "CUDA"; loaded after CMBLensingCUDAExt).
#
# It keeps CuArray storage — every non-hot method keeps working through CUDA.jl — and re-routes ONLY the hot path to
# libcmbl_b200.so (C ABI: include/cmbl_b200.h) at the seams the reference itself dispatches on:
#
#   seam in the reference                                          what this file adds
#   -------------------------------------------------------------  ------------------------------------------------------------
#   m_plan_rfft(::Type{A}, dims, sz...)   src/util_fft.jl:32-35     a method for A<:CuArray, dims == (1,2): returns a B200RFFTPlan;
#     (every Fourier(f′,f)/Map(f′,f)/QUFourier/… of                  mul!/ldiv!/*/\ on it call cmbl_rfft2 / cmbl_irfft2, so ALL the
#      src/proj_lambert.jl:245-300 goes through it)                  reference's basis conversions use our FFT with no other change
#   precompute!!(::LenseFlow, f)          src/lenseflow.jl:80-115   a method for CuArray fields: builds the CachedLenseFlow WITHOUT the
#                                                                    15×6-map Julia-side cache; the p / M⁻¹ cache lives in the library
#   precompute!(::CachedLenseFlow)        src/lenseflow.jl:131-142  one-argument method, as in the reference: refills the device cache
#   *, \ on CachedLenseFlow and Adjoint   src/flowops.jl:11-14      cmbl_lenseflow_apply ops 0..3
#   Zygote @adjoint of L*f, L\f           src/flowops.jl:40-68      cmbl_lenseflow_grad (negδvelocityᴴ) or L'Δ when :ϕ ∈ AD_constants
#   dot(::LambertField, ::LambertField)   src/proj_lambert.jl:318   cmbl_dot
#   argmaxf_logpdf(ds::BaseDataSet, Ω, d) src/maximization.jl:17-42 cmbl_cg_create + cmbl_wiener_cg when ds is "load_sim-shaped"
#                                                                    (diagonal Cf, Cn, B, M = Mfourier*Mpix); anything else falls through
#   get_max_lensing_step(ϕ, η)            src/lenseflow.jl:242-256  cmbl_max_lensing_step
#
# STATUS: Julia is not installed in the build image, so this file has never been executed; it is written against the
# reference's sources at 8e75a7c and the same ABI is exercised call for call from Python (cmblensing.jl_b200/__init__.py) by
# tests/ and bench.py.  julia/make_fixtures.jl writes reference outputs that tests/test_reference_fixtures.py consumes, which is
# the first thing to run wherever a Julia toolchain exists.
module CMBLensingB200Ext

using CMBLensing, CUDA, LinearAlgebra, AbstractFFTs, StaticArrays
using Zygote: @adjoint
using CMBLensing: BaseField, BaseFourier, LambertField, LambertMap, LambertFourier, ProjLambert, LenseFlow, CachedLenseFlow,
    RK4Solver, DiagOp, BlockDiagIEB, BaseDataSet, FieldTuple, Map, Fourier, IEBFourier, Ł, Ð, batch, unbatch,
    SpatialBasis, Hessian_logpdf_preconditioner, spin_adjoint
import CMBLensing: m_plan_rfft, precompute!, precompute!!, argmaxf_logpdf, get_max_lensing_step
import LinearAlgebra: dot, mul!, ldiv!
import Base: *, \

const lib = get(ENV, "CMBL_B200_LIB", "libcmbl_b200.so")
const CuLambertField{B,T} = BaseField{B,<:ProjLambert,T,<:CuArray}
const FT = Union{Float32,Float64}

check(rc) = rc == 0 || error(unsafe_string(ccall((:cmbl_last_error, lib), Cstring, ())))   # same exception class as the reference's error()
dtype(::Type{Float32}) = Cint(0); dtype(::Type{Float64}) = Cint(1)
stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))
ispow2ge4(n) = n >= 4 && ispow2(n)

# ---- plans --------------------------------------------------------------------------------------------------------------------
# One library plan per (Ny, Nx, θpix, T, device).  The FFT tables depend on (Ny, Nx, T) only; θpix enters the ℓ grids (∇, QU↔EB),
# so the FFT-only seam below uses θpix = 1 and the field-level entry points use the field's own ProjLambert.
const plans = Dict{Any,Ptr{Cvoid}}()
function plan(Ny::Int, Nx::Int, θpix::Real, ::Type{T}) where {T<:FT}
    get!(plans, (Ny, Nx, Float64(θpix), T, CUDA.deviceid())) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:cmbl_plan_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cdouble, Cint), h, CUDA.deviceid(), Ny, Nx, θpix, dtype(T)))
        h[]
    end
end
plan(p::ProjLambert{T}) where {T} = plan(p.Ny, p.Nx, p.θpix, T)
supported(p::ProjLambert) = ispow2ge4(p.Ny) && ispow2ge4(p.Nx) && p.Ny <= 8192 && p.Nx <= 8192 && p.T <: FT

# ---- FFT seam: the memoised plan the reference asks for (src/util_fft.jl:32-35) --------------------------------------------------
struct B200RFFTPlan{T,N} <: AbstractFFTs.Plan{T}
    sz :: NTuple{N,Int}              # size of the REAL array (Ny, Nx, ...)
    h  :: Ptr{Cvoid}
end
Base.size(p::B200RFFTPlan) = p.sz
planes(sz) = prod(sz[3:end]; init=1)
function mul!(dst::CuArray{Complex{T},N}, p::B200RFFTPlan{T,N}, src::CuArray{T,N}) where {T,N}        # m_rfft!: unnormalised R2C
    check(ccall((:cmbl_rfft2, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}), p.h, src, dst, planes(p.sz), stream())); dst
end
function ldiv!(dst::CuArray{T,N}, p::B200RFFTPlan{T,N}, src::CuArray{Complex{T},N}) where {T,N}       # m_irfft!: 1/(Ny·Nx), input intact
    check(ccall((:cmbl_irfft2, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}), p.h, src, dst, planes(p.sz), stream())); dst
end
*(p::B200RFFTPlan{T,N}, src::CuArray{T,N}) where {T,N} = mul!(similar(src, Complex{T}, (p.sz[1]÷2+1, p.sz[2:end]...)), p, src)
\(p::B200RFFTPlan{T,N}, src::CuArray{Complex{T},N}) where {T,N} = ldiv!(similar(src, T, p.sz), p, src)
# more specific than the reference's generic @memoize'd method, so it is what m_rfft / m_irfft / m_rfft! / m_irfft! pick up for
# CuArrays; anything the library does not cover (other dims, odd sizes, Dual numbers) goes to CUFFT exactly as before
function m_plan_rfft(::Type{A}, dims::Tuple{Int,Int}, sz::Int...) where {T<:FT, N, A<:CuArray{T,N}}
    if dims == (1,2) && ispow2ge4(sz[1]) && ispow2ge4(sz[2])
        B200RFFTPlan{T,N}(sz, plan(sz[1], sz[2], 1.0, T))
    else
        invoke(m_plan_rfft, Tuple{Type{<:AbstractArray{T,N}}, Any, Vararg{Any}}, A, dims, sz...)
    end
end

# ---- dot (src/proj_lambert.jl:318-328): per-batch values, no intermediate arrays ------------------------------------------------
function b200_dot(a::CuLambertField{B,T}, b::CuLambertField{B,T}, basis::Cint) where {B,T}
    nb = max(a.Nbatch, b.Nbatch); out = Vector{Float64}(undef, nb)
    (size(a.arr) == size(b.arr)) || return invoke(dot, Tuple{LambertField{B},LambertField{B}}, a, b)     # broadcasting batches: reference path
    check(ccall((:cmbl_dot, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cvoid}),
                plan(a.metadata), basis, a.arr, b.arr, size(a.arr, 3), nb, out, stream()))
    nb == 1 ? T(out[1]) : batch(T.(out))
end
dot(a::CuLambertField{B,T}, b::CuLambertField{B,T}) where {B<:SpatialBasis{Map},T<:FT}     = supported(a.metadata) ? b200_dot(a, b, Cint(0)) : invoke(dot, Tuple{LambertField{B},LambertField{B}}, a, b)
dot(a::CuLambertField{B,T}, b::CuLambertField{B,T}) where {B<:SpatialBasis{Fourier},T<:FT} = supported(a.metadata) ? b200_dot(a, b, Cint(1)) : invoke(dot, Tuple{LambertField{B},LambertField{B}}, a, b)

# ---- LenseFlow -----------------------------------------------------------------------------------------------------------------
# Device-side state of a CachedLenseFlow: one library handle per (Npol, Nbatch of f) the operator is applied to, refilled in place
# when ϕ changes (precompute!!), destroyed by a finalizer when the Julia object dies (the p-cache is GBs).
mutable struct DeviceFlow
    handles :: Dict{Tuple{Int,Int},Ptr{Cvoid}}     # (Npol, Nb_f) => cmbl_flow*
    fresh   :: Dict{Tuple{Int,Int},Bool}           # p-cache of that handle matches L.ϕ[]
end
const devflows = WeakKeyDict{Any,DeviceFlow}()
function devflow(L::CachedLenseFlow)
    get!(devflows, L) do
        d = DeviceFlow(Dict(), Dict())
        finalizer(L) do _
            for h in values(d.handles); ccall((:cmbl_lenseflow_destroy, lib), Cint, (Ptr{Cvoid},), h); end
        end
        d
    end
end
function handle(L::CachedLenseFlow, f::CuLambertField{B,T}) where {B,T}
    d = devflow(L); key = (size(f.arr, 3), f.Nbatch); ϕ = L.ϕ[]
    h = get!(d.handles, key) do
        r = Ref{Ptr{Cvoid}}()
        check(ccall((:cmbl_lenseflow_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Cint, Cint, Cint, Cint),
                    r, plan(f.metadata), L.odesolve.nsteps, key[1], key[2], ϕ.Nbatch))
        d.fresh[key] = false
        r[]
    end
    if !get(d.fresh, key, false)                     # precompute! (src/lenseflow.jl:131-142) on the device; with M⁻¹ for the pullbacks
        ϕm = Map(ϕ)
        check(ccall((:cmbl_lenseflow_precompute, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Cvoid}), h, ϕm.arr, 0, 1, stream()))
        d.fresh[key] = true
    end
    h
end
usable(L::CachedLenseFlow, f) = f isa CuLambertField && supported(f.metadata) && L.odesolve isa RK4Solver && L.t₀ == 0 && L.t₁ == 1

# precompute!!(::LenseFlow, f): the reference allocates 15·(2+4) ϕ-sized maps here; on this path they live in the library, so the
# Julia-side dictionaries stay empty and only the (small) "wide" scratch fields that size(L) and adapt() look at are created
function precompute!!(Lϕ::LenseFlow{S,T}, f::CuLambertField) where {S<:RK4Solver,T}
    supported(f.metadata) || return invoke(precompute!!, Tuple{LenseFlow{S,T},Any}, Lϕ, f)
    ϕ = Lϕ.ϕ
    Łϕ = Ł(ϕ); D = typeof(Diagonal(Łϕ))
    f′ = Ł(ϕ) .* Ł(f); ϕ′ = spin_adjoint(f′) * f′
    Łϕ′, Ðϕ′, Łf′, Ðf′ = Ł(ϕ′), Ð(ϕ′), Ł(f′), Ð(f′)
    CachedLenseFlow(Ref{Any}(ϕ), Ref(false), Lϕ.odesolve, Lϕ.t₀, Lϕ.t₁,
        Dict{Float16,SVector{2,D}}(), Dict{Float16,SMatrix{2,2,D,4}}(),
        Łf′, Ðf′, @SVector[Łf′, Łf′], @SVector[Ðf′, Ðf′], Łϕ′, Ðϕ′, @SVector[Łϕ′, Łϕ′], @SVector[Ðϕ′, Ðϕ′])
end
# precompute!(Lϕ) — ONE argument, as in the reference; called by precompute!!(::CachedLenseFlow, f) when (Lϕ)(ϕ′) flagged a new ϕ.
# Here it only invalidates the device caches; they are refilled (in place, no reallocation) by the next apply.
function precompute!(Lϕ::CachedLenseFlow{<:Any,<:Any,<:Any,<:CuLambertField})
    haskey(devflows, Lϕ) && (d = devflows[Lϕ]; for k in keys(d.fresh); d.fresh[k] = false; end)
    isempty(Lϕ.p) || invoke(precompute!, Tuple{CachedLenseFlow}, Lϕ)      # a cache built by the reference path keeps working
    Lϕ
end

function apply(L::CachedLenseFlow, op::Integer, g::CuLambertField)
    out = similar(g)
    check(ccall((:cmbl_lenseflow_apply, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}), handle(L, g), op, g.arr, out.arr, stream()))
    out
end
const CuCLF = CachedLenseFlow{<:Any,<:Any,<:Any,<:CuLambertField}
# src/flowops.jl:11-14 (results in the basis the reference's ODE state lives in: Ł for L and L\, Ð for L' and L'\)
*(L::CuCLF, f::CuLambertField)                  = usable(L, f) ? apply(precompute!!(L, f), 0, Ł(f)) : invoke(*, Tuple{CMBLensing.FlowOp,CMBLensing.Field}, L, f)
\(L::CuCLF, f::CuLambertField)                  = usable(L, f) ? apply(precompute!!(L, f), 2, Ł(f)) : invoke(\, Tuple{CMBLensing.FlowOp,CMBLensing.Field}, L, f)
*(L::Adjoint{<:Any,<:CuCLF}, f::CuLambertField) = usable(parent(L), f) ? apply(precompute!!(parent(L), f), 1, Ð(f)) : invoke(*, Tuple{Adjoint{<:Any,<:CMBLensing.FlowOp},CMBLensing.Field}, L, f)
\(L::Adjoint{<:Any,<:CuCLF}, f::CuLambertField) = usable(parent(L), f) ? apply(precompute!!(parent(L), f), 3, Ð(f)) : invoke(\, Tuple{Adjoint{<:Any,<:CMBLensing.FlowOp},CMBLensing.Field}, L, f)

# pullbacks (src/flowops.jl:40-68).  bug_compat = true reproduces the reference's aliased 2×2 product in the δϕ flow
# (src/lenseflow.jl:198-200 with src/field_vectors.jl:48-49); set CMBL_B200_EXACT_GRAD=1 for the exact product.
const bug_compat = get(ENV, "CMBL_B200_EXACT_GRAD", "0") == "0"
function pullback_flow(L::CachedLenseFlow, op::Integer, f_out::CuLambertField, Δ)
    g, δ = Ł(f_out), Ð(Δ)
    δf = similar(δ)
    δϕ = similar(Ð(L.ϕ[]), Complex{real(eltype(g))}, size(δ.arr, 1), size(δ.arr, 2), 1, g.Nbatch)      # one ϕ-gradient per batch item of f
    check(ccall((:cmbl_lenseflow_grad, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                handle(L, g), op, g.arr, δ.arr, δf.arr, δϕ.arr, bug_compat, stream()))
    L.ϕ[].Nbatch == 1 && g.Nbatch > 1 && (δϕ = sum(unbatch(δϕ)))                                        # shared ϕ: sum over the batch
    δf, δϕ
end
@adjoint function *(Lϕ::CuCLF, f::CuLambertField{B}) where {B}
    usable(Lϕ, f) || return Zygote.pullback((L, f) -> invoke(*, Tuple{CMBLensing.FlowOp,CMBLensing.Field}, L, f), Lϕ, f)
    L = precompute!!(Lϕ, f); f̃ = L * f
    back(Δ) = :ϕ in get(task_local_storage(), :AD_constants, ()) ? (nothing, B(L' * Δ)) : ((δf, δϕ) = pullback_flow(L, 0, f̃, Δ); (δϕ, B(δf)))
    f̃, back
end
@adjoint function \(Lϕ::CuCLF, f̃::CuLambertField{B}) where {B}
    usable(Lϕ, f̃) || return Zygote.pullback((L, f) -> invoke(\, Tuple{CMBLensing.FlowOp,CMBLensing.Field}, L, f), Lϕ, f̃)
    L = precompute!!(Lϕ, f̃); f = L \ f̃
    back(Δ) = :ϕ in get(task_local_storage(), :AD_constants, ()) ? (nothing, B(L' \ Δ)) : ((δf, δϕ) = pullback_flow(L, 2, f, Δ); (δϕ, B(δf)))
    f, back
end

function get_max_lensing_step(ϕ::CuLambertField{B,T}, η::CuLambertField) where {B,T<:FT}
    (supported(ϕ.metadata) && size(ϕ.arr) == size(η.arr)) || return invoke(get_max_lensing_step, Tuple{Any,Any}, ϕ, η)
    a, b = Fourier(ϕ), Fourier(η); out = Vector{Float64}(undef, a.Nbatch)
    check(ccall((:cmbl_max_lensing_step, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Cint, CuPtr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cvoid}),
                plan(a.metadata), a.arr, 1, b.arr, 1, a.Nbatch, out, stream()))
    T(minimum(out))                                   # the reference returns one number: the minimum over pixels and batch
end

# ---- argmaxf_logpdf (src/maximization.jl:17-42) for a load_sim-shaped BaseDataSet -------------------------------------------------
struct DatasetDesc                                   # mirrors cmbl_dataset_desc (include/cmbl_b200.h)
    Npol::Cint; Nb::Cint
    Cf::CuPtr{Cvoid}; Cn::CuPtr{Cvoid}; Cnhat::CuPtr{Cvoid}; B::CuPtr{Cvoid}; Bhat::CuPtr{Cvoid}; Mf::CuPtr{Cvoid}
    mask_pix::CuPtr{Cvoid}; d::CuPtr{Cvoid}
end
# real (Ny÷2+1, Nx, Npol) planes of a harmonic-basis diagonal; BlockDiagIEB -> its four planes [ΣTE[1,1], ΣTE[2,1], ΣTE[2,2], ΣB]
planes_of(D::DiagOp{<:BaseFourier}) = real.(diag(D).arr)
planes_of(L::BlockDiagIEB) = cat(real.(diag(L.ΣTE[1,1]).arr), real.(diag(L.ΣTE[2,1]).arr), real.(diag(L.ΣTE[2,2]).arr), real.(diag(L.ΣB).arr); dims=3)
harmonic_op(x) = x isa DiagOp{<:BaseFourier} || x isa BlockDiagIEB
# M = Mfourier * Mpix (src/dataset.jl:277-286) is a LazyBinaryOp{*}; a bare harmonic diagonal means "no pixel mask"
split_mask(M) = harmonic_op(M) ? (M, nothing) :
    (M isa CMBLensing.LazyBinaryOp{*} && harmonic_op(M.X) && M.Y isa DiagOp{<:CuLambertField{<:SpatialBasis{Map}}}) ? (M.X, M.Y) : nothing
function b200_dataset(ds::BaseDataSet, θ, d)
    ops = (ds.Cf(θ), ds.Cn(θ), ds.Cn̂(θ), ds.B(θ), ds.B̂(θ)); mm = split_mask(ds.M(θ))
    (all(harmonic_op, ops) && mm !== nothing && d isa CuLambertField && supported(d.metadata)) || return nothing
    (ops..., mm...)
end
function argmaxf_logpdf(ds::BaseDataSet, Ω::NamedTuple, d = ds.d; fstart = nothing, preconditioner = :diag,
                        conjgrad_kwargs = (tol=1e-1, nsteps=500), offset = false)
    θ = get(Ω, :θ, (;)); parts = haskey(Ω, :ϕ) ? b200_dataset(ds, θ, d) : nothing
    fallback() = invoke(argmaxf_logpdf, Tuple{CMBLensing.DataSet,NamedTuple,Any}, ds, Ω, d; fstart, preconditioner, conjgrad_kwargs, offset)
    (parts === nothing || preconditioner != :diag || !isempty(setdiff(keys(conjgrad_kwargs), (:tol, :nsteps, :history_keys, :progress)))) && return fallback()
    Cf, Cn, Cn̂, B, B̂, Mf, Mpix = parts
    d isa BaseFourier || return fallback()            # load_sim stores d in the harmonic basis (Fourier / EBFourier / IEBFourier)
    dh = d
    Npol, Nb = size(dh.arr, 3), dh.Nbatch
    Lϕ = precompute!!(ds.L(Ω.ϕ), Ł(dh)); usable(Lϕ, Ł(dh)) || return fallback()
    keep = (planes_of(Cf), planes_of(Cn), planes_of(Cn̂), planes_of(B), planes_of(B̂), planes_of(Mf), Mpix === nothing ? nothing : diag(Mpix).arr)
    desc = Ref(DatasetDesc(Npol, Nb, pointer.(keep[1:6])..., Mpix === nothing ? CU_NULL : pointer(keep[7]), pointer(dh.arr)))
    cg = Ref{Ptr{Cvoid}}()
    nsteps = get(conjgrad_kwargs, :nsteps, length(dh)); tol = Float64(get(conjgrad_kwargs, :tol, sqrt(eps())))
    f = similar(dh); hist = Matrix{Float64}(undef, Nb, nsteps); iters = Ref{Cint}(0)
    GC.@preserve keep dh begin
        check(ccall((:cmbl_cg_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{DatasetDesc}, Ptr{Cvoid}), cg, handle(Lϕ, Ł(dh)), desc, stream()))
        try
            check(ccall((:cmbl_wiener_cg, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cdouble, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cvoid}),
                        cg[], fstart === nothing ? CU_NULL : pointer(typeof(dh)(fstart).arr), f.arr, nsteps, tol, offset, iters, hist, stream()))
        finally
            ccall((:cmbl_cg_destroy, lib), Cint, (Ptr{Cvoid},), cg[])
        end
    end
    T = real(eltype(dh)); hk = get(conjgrad_kwargs, :history_keys, nothing)
    res(i) = Nb == 1 ? T(hist[1, i]) : batch(T.(hist[:, i]))
    history = hk === nothing ? fill(nothing, iters[]) : [CMBLensing.select((; i, res = res(i)), hk) for i in 1:iters[]]      # (:i, :res) — what MAP_joint asks for
    (f, history)
end

end

# CMBLensingB200Ext.jl — the package extension a CMBLensing.jl maintainer would add next to ext/CMBLensingCUDAExt.jl.
# It keeps CuArray storage (so every non-hot method keeps working through CUDA.jl) and overrides only the hot-path methods
# with ccalls into libcmbl_b200.so (C ABI: include/cmbl_b200.h).  UNTESTED HERE: Julia is not installed in the build image;
# the same ABI is exercised from Python (cmblensing.jl_b200/__init__.py) by tests/ and bench.py.
module CMBLensingB200Ext

using CMBLensing, CUDA, LinearAlgebra
using CMBLensing: BaseField, LambertField, FlatField, ProjLambert, CachedLenseFlow, FlowOp, DiagOp, BatchedReal,
    Map, Fourier, QUMap, QUFourier, EBFourier, Ł, Ð, batch
import CMBLensing: m_rfft!, m_irfft!, precompute!, argmaxf_logpdf
import LinearAlgebra: dot
import Base: *, \

const lib = get(ENV, "CMBL_B200_LIB", "libcmbl_b200.so")
const CuLambertField{B,T} = LambertField{B,<:Any,T,<:CuArray}

check(rc) = rc == 0 || error(unsafe_string(ccall((:cmbl_last_error, lib), Cstring, ())))
dtype(::Type{Float32}) = Cint(0); dtype(::Type{Float64}) = Cint(1)
stream() = Ptr{Cvoid}(UInt(CUDA.stream().handle))

# ---- plan: one per ProjLambert (memoised like m_plan_rfft, src/util_fft.jl:32-39) --------------------------------------
const plans = Dict{Any,Ptr{Cvoid}}()
function plan(p::ProjLambert{T}) where {T}
    get!(plans, (p.Ny, p.Nx, p.θpix, T, CUDA.device())) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:cmbl_plan_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cdouble, Cint),
                    h, CUDA.deviceid(), p.Ny, p.Nx, p.θpix, dtype(T)))
        h[]
    end
end

# ---- FFT: m_rfft! / m_irfft! on CuArrays (src/util_fft.jl:26-27) -------------------------------------------------------
planes(a) = prod(size(a)[3:end])
function m_rfft!(dst::CuArray{Complex{T}}, src::CuArray{T}, dims; proj) where {T}
    check(ccall((:cmbl_rfft2, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                plan(proj), src, dst, planes(src), stream())); dst
end
function m_irfft!(dst::CuArray{T}, src::CuArray{Complex{T}}, dims; proj) where {T}
    check(ccall((:cmbl_irfft2, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                plan(proj), src, dst, planes(dst), stream())); dst
end

# ---- dot (src/proj_lambert.jl:318-328): per-batch values without a device->host collect of the field --------------------
function dot(a::CuLambertField{B,T}, b::CuLambertField{B,T}) where {B,T}
    nb = max(a.Nbatch, b.Nbatch); out = Vector{Float64}(undef, nb)
    check(ccall((:cmbl_dot, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cvoid}),
                plan(a.metadata), B <: CMBLensing.Basislike{Fourier} ? 1 : 0, a.arr, b.arr, size(a.arr, 3), nb, out, stream()))
    nb == 1 ? T(out[1]) : batch(T.(out))
end

# ---- LenseFlow: one handle per CachedLenseFlow; precompute! and the four flow ops (src/flowops.jl:11-14) -----------------
const flows = WeakKeyDict{Any,Ptr{Cvoid}}()
function handle(L::CachedLenseFlow, f)
    get!(flows, L) do
        h = Ref{Ptr{Cvoid}}(); Npol = size(f.arr, 3)
        check(ccall((:cmbl_lenseflow_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Cint, Cint, Cint, Cint),
                    h, plan(f.metadata), L.ODESolver.nsteps, Npol, f.Nbatch, L.ϕ[].Nbatch))
        ϕ = Map(L.ϕ[])
        check(ccall((:cmbl_lenseflow_precompute, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Cvoid}), h[], ϕ.arr, 0, 1, stream()))
        h[]
    end
end
function apply(L::CachedLenseFlow, op, f, out)
    check(ccall((:cmbl_lenseflow_apply, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                handle(L, f), op, f.arr, out.arr, stream())); out
end
*(L::CachedLenseFlow, f::CuLambertField)                    = (g = Ł(f); apply(L, 0, g, similar(g)))       # Map  -> Map
*(L::Adjoint{<:Any,<:CachedLenseFlow}, f::CuLambertField)   = (g = Ð(f); apply(parent(L), 1, g, similar(g)))  # Fourier -> Fourier
\(L::CachedLenseFlow, f::CuLambertField)                    = (g = Ł(f); apply(L, 2, g, similar(g)))
\(L::Adjoint{<:Any,<:CachedLenseFlow}, f::CuLambertField)   = (g = Ð(f); apply(parent(L), 3, g, similar(g)))

# ---- pullbacks of L*f and L\f (src/flowops.jl:40-68): the transpose flow negδvelocityᴴ on the device --------------------------
# (the handle must have been precomputed with with_minv = 1, as `handle` above does)
function lenseflow_pullback(L::CachedLenseFlow, op, f_out, Δ; bug_compat=true)
    g, δ = Ł(f_out), Ð(Δ)
    δf = similar(δ); δϕ = similar(Ð(L.ϕ[]), f_out.Nbatch)
    check(ccall((:cmbl_lenseflow_grad, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                handle(L, g), op, g.arr, δ.arr, δf.arr, δϕ.arr, bug_compat, stream()))
    δf, δϕ
end
@adjoint *(L::CachedLenseFlow, f::CuLambertField) = (Lf = L * f; (Lf, Δ -> reverse(lenseflow_pullback(L, 0, Lf, Δ))))
@adjoint \(L::CachedLenseFlow, f::CuLambertField) = (Lf = L \ f; (Lf, Δ -> reverse(lenseflow_pullback(L, 2, Lf, Δ))))

# ---- precompute!! (src/lenseflow.jl:80-129): refill the SAME device cache when ϕ changes (no reallocation in a line search) ----
function precompute!(L::CachedLenseFlow{<:Any,<:Any,<:Any,<:CuLambertField}, f)
    ϕ = Map(L.ϕ[])
    check(ccall((:cmbl_lenseflow_precompute, lib), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Cint, Cint, Ptr{Cvoid}), handle(L, f), ϕ.arr, 0, 1, stream())); L
end

# ---- BlockDiagIEB (src/specialops.jl:77-82,87-88): [ΣTT ΣTE; ΣTE ΣEE] ⊕ ΣBB as four real half-planes --------------------------
blockplanes(L::BlockDiagIEB) = cat(real.(L.ΣTE[1,1].diag.arr), real.(L.ΣTE[2,1].diag.arr), real.(L.ΣTE[2,2].diag.arr), real.(L.ΣB.diag.arr); dims=3)
function blockdiag_apply(L::BlockDiagIEB, f::CuLambertField, mode)      # mode 0: L*f, 1: L\f = pinv(L)*f, 2: sqrt(L)*f
    g = IEBFourier(f); out = similar(g)
    check(ccall((:cmbl_blockdiag_ieb, lib), Cint, (Ptr{Cvoid}, Cint, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Cint, Ptr{Cvoid}),
                plan(g.metadata), mode, blockplanes(L), g.arr, out.arr, g.Nbatch, stream())); out
end
*(L::BlockDiagIEB, f::CuLambertField)  = blockdiag_apply(L, f, 0)
\(L::BlockDiagIEB, f::CuLambertField) = blockdiag_apply(L, f, 1)

# ---- argmaxf_logpdf for a BaseDataSet on the GPU (src/maximization.jl:17-42) ------------------------------------------
struct DatasetDesc
    Npol::Cint; Nb::Cint
    Cf::CuPtr{Cvoid}; Cn::CuPtr{Cvoid}; Cnhat::CuPtr{Cvoid}; B::CuPtr{Cvoid}; Bhat::CuPtr{Cvoid}; Mf::CuPtr{Cvoid}
    mask_pix::CuPtr{Cvoid}; d::CuPtr{Cvoid}
end
# (for pol = :IP every operator pointer is `blockplanes(op)`, Npol = 3)
# (construction of the descriptor from ds.Cf, ds.Cn, ds.Cn̂, ds.B, ds.B̂, ds.M (= Mfourier * Mpix) and ds.d, then
#  cmbl_cg_create + cmbl_wiener_cg; returns (f, history) with history[i] = (i=i, res=batch(res_hist[:,i])).)

end

# make_fixtures.jl — writes OUTPUTS OF THE REFERENCE ITSELF (CMBLensing.jl on the CPU, Float64, FFTW) for the hot path, together with
# the exact inputs, as .npy files under tests/golden/ref/.  tests/test_reference_fixtures.py consumes them when present: the oracle
# (-m "not gpu") and the CUDA library (-m gpu) are then compared with the reference's own numbers — that is the pin that turns
# "parity unpinned" (oracle/cmbl_oracle.py header) into "pinned".  Julia is not available in the build image, so this script could
# not be run there; run it once wherever CMBLensing.jl @ 8e75a7c is installed:
#
#     julia --project=/path/to/CMBLensing.jl julia/make_fixtures.jl tests/golden/ref
#
# Setup mirrors test/runtests.jl:533-581 ("Lensing") and :587-616 ("Posterior"): ProjLambert, Cℓ = camb() (default parameters: read
# from dat/default_camb_Cls.jld2, no Python needed), fields from simulate(rng, C), LenseFlow with 7 RK4 steps.
using CMBLensing, Random, LinearAlgebra, Zygote
using CMBLensing: precompute!!, QUFourier, QUMap, EBFourier, Map, Fourier

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "ref")
mkpath(outdir)

# minimal .npy (v1.0) writer: Julia arrays are column-major, so fortran_order = True and the bytes go out as they are
npy_descr(::Type{Float64}) = "<f8"; npy_descr(::Type{Float32}) = "<f4"; npy_descr(::Type{ComplexF64}) = "<c16"; npy_descr(::Type{ComplexF32}) = "<c8"
function write_npy(name, a::AbstractArray{T}) where {T}
    a = Array(a)
    shape = join(size(a), ", ") * (ndims(a) == 1 ? "," : "")
    hdr = "{'descr': '$(npy_descr(T))', 'fortran_order': True, 'shape': ($shape), }"
    pad = 64 - mod(10 + length(hdr) + 1, 64); hdr *= " "^pad * "\n"
    open(joinpath(outdir, name * ".npy"), "w") do io
        write(io, UInt8[0x93], "NUMPY", UInt8[1, 0], UInt16(length(hdr)), hdr, a)
    end
end
write_txt(name, x) = open(io -> print(io, x), joinpath(outdir, name * ".txt"), "w")

T = Float64
Cℓ = camb()
for (Ny, Nx) in [(128, 128), (64, 32)]
    tag = "$(Ny)x$(Nx)"
    rng = MersenneTwister(1000 + Ny + Nx)
    proj = ProjLambert(; Ny, Nx, θpix = 2, T)
    Cϕ = Cℓ_to_Cov(:I, proj, Cℓ.unlensed_total.ϕϕ)
    Cf = Cℓ_to_Cov(:P, proj, Cℓ.unlensed_total.EE, Cℓ.unlensed_total.BB)
    ϕ = simulate(rng, Cϕ); f = simulate(rng, Cf)
    Lϕ = precompute!!(LenseFlow(ϕ, 7), f)
    write_npy("lf_$(tag)_phi_fourier", Fourier(ϕ).arr)
    write_npy("lf_$(tag)_f_qumap", QUMap(f).arr)
    write_npy("lf_$(tag)_L_f_qumap", QUMap(Lϕ * f).arr)                       # Lϕ * f      (src/flowops.jl:11)
    write_npy("lf_$(tag)_LH_f_qufourier", QUFourier(Lϕ' * f).arr)              # Lϕ' * f     (:12)
    write_npy("lf_$(tag)_Linv_f_qumap", QUMap(Lϕ \ f).arr)                     # Lϕ \ f      (:13)
    write_npy("lf_$(tag)_LHinv_f_qufourier", QUFourier(Lϕ' \ f).arr)           # Lϕ' \ f     (:14)
    # pullback of (ϕ, f) -> Lϕ*f with the cotangent Δ = f (the transpose-δ flow, src/flowops.jl:40-55), as computed by the reference
    _, back = Zygote.pullback((ϕ, f) -> LenseFlow(ϕ, 7) * f, ϕ, f)
    δϕ, δf = back(QUMap(f))
    write_npy("lf_$(tag)_grad_phi_fourier", Fourier(δϕ).arr)
    write_npy("lf_$(tag)_grad_f_qufourier", QUFourier(δf).arr)
    # p cache at t = 0, 1/2, 1 (src/lenseflow.jl:131-142)
    for t in (0.0, 0.5, 1.0)
        p = Lϕ.p[Float16(t)]
        write_npy("lf_$(tag)_p_t$(t)", cat(Map(diag(p[1])).arr, Map(diag(p[2])).arr; dims = 3))
    end
end

# CG Wiener filter on a load_sim dataset (src/dataset.jl:186-340, src/maximization.jl:17-42, src/numerical_algorithms.jl:73-134)
let Nside = 128
    (; ds, f, ϕ) = load_sim(; θpix = 2, Nside, T, pol = :P, rng = MersenneTwister(7), μKarcminT = 3, ℓknee = 100, αknee = 3, beamFWHM = 0,
                              bandpass_mask = LowPass(3000), pixel_mask_kwargs = (edge_padding_deg = 0.5, apodization_deg = 0.5, num_ptsrcs = 0))
    harm(D) = real.(diag(D).arr)
    write_npy("cg_d_ebfourier", ds.d.arr); write_npy("cg_phi_fourier", Fourier(ϕ).arr); write_npy("cg_f_ebfourier", EBFourier(f).arr)
    write_npy("cg_Cf", harm(ds.Cf())); write_npy("cg_Cn", harm(ds.Cn())); write_npy("cg_Cnhat", harm(ds.Cn̂())); write_npy("cg_B", harm(ds.B())); write_npy("cg_Bhat", harm(ds.B̂()))
    M = ds.M()
    write_npy("cg_Mf", harm(M.X)); write_npy("cg_Mpix", diag(M.Y).arr)        # M = Mfourier * Mpix (src/dataset.jl:277-286)
    write_npy("cg_gradientf", EBFourier(gradientf_logpdf(ds; f, ϕ)).arr)      # src/dataset.jl:76-80
    fwf, hist = argmaxf_logpdf(ds, (; ϕ); conjgrad_kwargs = (tol = 0, nsteps = 8, history_keys = (:i, :res)))
    write_npy("cg_fwf_ebfourier", EBFourier(fwf).arr)
    write_npy("cg_res_history", Float64[h.res for h in hist])
    write_txt("cg_setup", "load_sim θpix=2 Nside=$Nside pol=:P T=Float64 MersenneTwister(7); LenseFlow n=7; conjugate_gradient tol=0 nsteps=8")
end
println("reference fixtures written to ", outdir)

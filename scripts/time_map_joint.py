"""Development aid: wall time of MAP_joint steps at Nside=1024 QU batch 8 (BASELINE config 3 shape) or another shape.
usage: time_map_joint.py [f64|f32] [N] [pol P|IP] [NB] [steps]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g
import cmbl_oracle as O                      # only Cℓ tables / mask profile for the synthetic inputs
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pol = sys.argv[3] if len(sys.argv) > 3 else "P"
NB = int(sys.argv[4]) if len(sys.argv) > 4 else 8
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
tT = torch.float64 if dtype == "f64" else torch.float32
dev = "cuda:0"
proj = pkg.ProjLambert(N, N, 2.0, tT, dev)
cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
npol = {"P": 2, "IP": 3}[pol]; lense = ("Map", "QUMap", "IQUMap")[npol - 1]
gen = torch.Generator(device=dev).manual_seed(5)
w = lambda p: pkg.Field(("Map", "QUMap", "IQUMap")[p - 1], torch.randn((NB, p, N, N), dtype=tT, device=dev, generator=gen), proj)
nT = O.noise_cls(ell); zero = np.zeros_like(nT); one = np.ones_like(nT); lb, wl = O.lowpass_wl(3000)
if pol == "IP":
    Cf = pkg.Cℓ_to_Cov("IP", proj, ell, cls["ut_TT"], cls["ut_EE"], cls["ut_BB"], cls["ut_TE"])
    Cn = pkg.Cℓ_to_Cov("IP", proj, ell, nT, 2 * nT, 2 * nT, zero)
    Mf = pkg.Cℓ_to_Cov("IP", proj, lb, wl, wl, wl, np.zeros_like(wl), units=1)
    B = pkg.Cℓ_to_Cov("IP", proj, ell, one, one, one, zero, units=1)
else:
    Cf = pkg.Cℓ_to_Cov("P", proj, ell, cls["ut_EE"], cls["ut_BB"])
    Cn = pkg.Cℓ_to_Cov("P", proj, ell, 2 * nT, 2 * nT)
    Mf = pkg.Cℓ_to_Cov("P", proj, lb, wl, wl, units=1)
    B = pkg.Cℓ_to_Cov("P", proj, ell, one, one, units=1)
Cϕ = pkg.Cℓ_to_Cov("I", proj, ell, cls["pp"])
Cft = (pkg.Cℓ_to_Cov("P", proj, ell, cls["tot_EE"], cls["tot_BB"]) if pol == "P" else None)          # Cf̃ for the quadratic estimate
L_ = np.arange(2, 16000, dtype=float)      # stand-in for quadratic_estimate(ds).Nϕ: a flat [L(L+1)]²N_L/2π = 1e-8 plateau rising beyond L ~ 1500
Nϕ = pkg.Cℓ_to_Cov("I", proj, L_, 2 * np.pi * 1e-8 / (L_ * (L_ + 1)) ** 2 * (1 + (L_ / 1500) ** 4))
mask = torch.from_numpy(O.cosine_border_mask(O.ProjLambert(N, N, 2.0, np.float32 if dtype == "f32" else np.float64), 1.0))
Mpix = pkg.DiagOp(pkg.Field(lense, mask[None, None].expand(1, npol, N, N).contiguous(), proj))
ϕ_true = pkg.DiagOp(pkg.Field("Fourier", torch.sqrt(Cϕ._real), proj)) * w(1)
ds0 = pkg.BaseDataSet(pkg.HarmonicBasis(w(npol)), Cf, Cn, B, Mf, Mpix, nsteps=7, Cϕ=Cϕ, Nϕ=Nϕ)
sim = pkg.simulate(ds0, ϕ_true, generator=gen)
ds = pkg.BaseDataSet(sim["d"], Cf, Cn, B, Mf, Mpix, nsteps=7, Cϕ=Cϕ, Nϕ=Nϕ, Cf̃=Cft)
if pol == "P":                                 # load_sim: Nϕ = quadratic_estimate(ds).Nϕ / Nϕ_fac, Nϕ_fac = 2 (src/dataset.jl:222,316)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    qe = pkg.quadratic_estimate(ds)
    torch.cuda.synchronize(); tq = time.perf_counter() - t0
    ds.Nϕ = pkg.DiagOp(pkg.Field("Fourier", (qe["Nϕ"]._real / 2).to(proj.cT), proj))
    ds.D = pkg.mixing_D(ds)                      # load_sim's mixing matrix (src/dataset.jl:325-332)
    a, b = pkg.Map(qe["ϕqe"]).arr, pkg.Map(ϕ_true).arr
    q = slice(N // 4, 3 * N // 4)
    cc = [float(torch.corrcoef(torch.stack([a[i, 0, q, q].flatten(), b[i, 0, q, q].flatten()]))[0, 1]) for i in range(NB)]
    print(f"quadratic_estimate (EB, Wiener-filtered): {tq*1e3:.0f} ms; corr(ϕqe, ϕ_true) = {np.round(cc, 3)}")
lib = pkg.load()
torch.cuda.synchronize(); n0 = lib.launch_count(); t0 = time.perf_counter()
f, ϕ, hist = pkg.MAP_joint(ds, nsteps=steps, conjgrad_kwargs=dict(tol=1e-1, nsteps=500))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"MAP_joint {dtype} N={N} pol={pol} NB={NB}: {steps} steps in {dt:.2f} s ({dt/steps:.2f} s/step), {lib.launch_count()-n0} kernel launches")
for h in hist:
    print(f"   step {h['step']}: CG iterations {h['cg_iters']}, line-search evaluations {h['linesearch_evals']}, α={h['α']:.4f}, Σ logpdf={float(np.sum(h['logpdf'])):.6e}")
a, b = pkg.Map(ϕ).arr, pkg.Map(ϕ_true).arr
q = slice(N // 4, 3 * N // 4)
cc = [float(torch.corrcoef(torch.stack([a[i, 0, q, q].flatten(), b[i, 0, q, q].flatten()]))[0, 1]) for i in range(NB)]
print("   corr(ϕ_MAP, ϕ_true) inside the mask per batch item:", np.round(cc, 3))
# time of the pieces of one ϕ step
f_m, ϕ_m = pkg.mix(ds, f, ϕ)
torch.cuda.synchronize(); t0 = time.perf_counter(); gf, gp = pkg.gradient_logpdf_mixed(ds, f_m, ϕ_m); torch.cuda.synchronize(); tg = time.perf_counter() - t0
t0 = time.perf_counter(); lp = pkg.logpdf(pkg.Mixed(ds), f_m, ϕ_m); torch.cuda.synchronize(); tl = time.perf_counter() - t0
print(f"   gradient of logpdf(Mixed) (2 flows + 2 δ-flows): {tg*1e3:.1f} ms;  one logpdf(Mixed) evaluation (precompute + 2 flows): {tl*1e3:.1f} ms")
# one HMC update of ϕ° (src/sampling.jl:397-417): N leap-frog steps, each one gradient of logpdf(Mixed(ds))
NL = 5
torch.cuda.synchronize(); t0 = time.perf_counter()
x, ΔH, acc = pkg.gibbs_sample_ϕ(ds, f_m, ϕ_m, symp_kwargs=(dict(N=NL, ϵ=0.01),), always_accept=False)
torch.cuda.synchronize(); th = time.perf_counter() - t0
if pol == "IP": ΔH = np.full_like(ΔH, np.nan)      # timing only: with the stand-in Nϕ of the IP branch (no quadratic estimate, no mixing matrix) ΔH is not meaningful
print(f"   HMC ϕ° update, {NL} leap-frog steps: {th*1e3:.0f} ms = {NL/th:.1f} leap-frog steps/s for the batch of {NB} ({NB*NL/th:.1f} chain-steps/s); ΔH = {np.round(ΔH, 3)}, accept = {acc}")
# per-kernel device time of one gradient (CUDA events around every launch)
lib.cdll.cmbl_profile_begin()
pkg.gradient_logpdf_mixed(ds, f_m, ϕ_m)
rows = []
for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
    nm, cnt, t = line.split(); rows.append((float(t), nm, int(cnt)))
tot = sum(r[0] for r in rows)
print("   gradient kernel profile (ms, launches): " + ", ".join(f"{nm} {t:.1f} ({cnt})" for t, nm, cnt in sorted(rows, reverse=True)[:12]) + f"; sum {tot:.1f} ms")

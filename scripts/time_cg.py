"""Development aid: per-kernel device time of one CG-Wiener iteration at the bench workload (Nside=1024 QU batch 8)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g
import cmbl_oracle as O
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
tT = torch.float64 if dtype == "f64" else torch.float32
npT = np.float64 if dtype == "f64" else np.float32
N, NB, dev = 1024, 8, "cuda:0"
proj = pkg.ProjLambert(N, N, 2.0, tT, dev)
op = O.ProjLambert(N, N, 2.0, npT)
cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
dg = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
Cphi = dg(O.cl_to_cov(op, ell, cls["pp"]))
Cf_np = np.stack([O.cl_to_cov(op, ell, cls[k]) for k in ("ut_EE", "ut_BB")])[None]
Cn_np = np.stack([O.cl_to_cov(op, ell, O.noise_cls(ell, pol=True)) for _ in range(2)])[None]
lb, wl = O.lowpass_wl(3000)
Mf_np = np.stack([O.cl_to_cov(op, lb, wl, units=1) for _ in range(2)])[None]
B_np = np.ones_like(Mf_np)
mask_np = np.broadcast_to(O.cosine_border_mask(op, 1.0), (1, 2, N, N)).copy()
gen = torch.Generator(device=dev).manual_seed(1)
white = lambda n, p: torch.randn((n, p, N, N), dtype=tT, device=dev, generator=gen)
phi = pkg.Fourier(pkg.Field("Map", white(NB, 1), proj)); phi = phi._like(phi.arr * torch.sqrt(Cphi))
f = pkg.Fourier(pkg.Field("QUMap", white(NB, 2), proj)); f = pkg.Field("EBFourier", f.arr * torch.sqrt(dg(Cf_np)), proj)
L = pkg.LenseFlow(phi, 7)
D = lambda a, basis="EBFourier": pkg.DiagOp(pkg.Field(basis, dg(a), proj))
ds = pkg.BaseDataSet(f, D(Cf_np), D(Cn_np), D(B_np), D(Mf_np), D(mask_np, "QUMap"), L=L, nsteps=7)
lib = pkg.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
h, *_ = ds._solver(phi)
res = (ctypes.c_double * NB)()
lib.call("cmbl_cg_begin", h, ctypes.c_void_p(0), 0, res, st)
for _ in range(2): lib.call("cmbl_cg_step", h, res, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): lib.call("cmbl_cg_step", h, res, st)
e1.record(); torch.cuda.synchronize()
print(f"{dtype}: {e0.elapsed_time(e1)/4:.3f} ms per CG iteration, res {res[0]:.4g}")
lib.cdll.cmbl_profile_begin()
lib.call("cmbl_cg_step", h, res, st)
tot = 0
for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
    nm, cnt, t = line.split(); tot += float(t)
    print(f"   {nm}: {int(cnt)} launches, total {float(t):.3f} ms, avg {float(t)/int(cnt)*1e3:.1f} us")
print("   sum", round(tot, 3), "ms")

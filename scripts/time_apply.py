"""Development aid: device time of one batched LenseFlow apply at the bench workload (Nside=1024 QU batch 8).
usage: python scripts/time_apply.py [f64|f32] [op] ; env CMBL_FLOW_CHUNK / CMBL_TILE_KB are read by the library."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
op = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tT = torch.float64 if dtype == "f64" else torch.float32
N, NB, NPOL = int(os.environ.get("N", "1024")), int(os.environ.get("NB", "8")), int(os.environ.get("NPOL", "2"))
LENSE, DERIV = ("Map", "QUMap", "IQUMap")[NPOL - 1], ("Fourier", "QUFourier", "IQUFourier")[NPOL - 1]
proj = pkg.ProjLambert(N, N, 2.0, tT, "cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
k = torch.fft.fftfreq(N, device="cuda:0")
kk = torch.sqrt(k[:, None] ** 2 + k[None, :] ** 2) + 1e-3
smooth = lambda a: torch.fft.ifft2(torch.fft.fft2(a.double()) / kk ** 2).real
phi = smooth(torch.randn((NB, 1, N, N), generator=gen, device="cuda:0")); phi = (phi / phi.std() * 2e-5 * 1024 / N).to(tT)          # ~arcminute deflections
f = pkg.Field(LENSE, torch.randn((NB, NPOL, N, N), dtype=tT, device="cuda:0", generator=gen), proj)
L = pkg.LenseFlow(pkg.Field("Map", phi, proj), 7)
cache = L.cache(f)
lib = pkg.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
if op in (0, 2):
    x = f.arr; out = torch.empty_like(x)
else:
    x = pkg.convert(f, DERIV).arr; out = torch.empty_like(x)
run = lambda: lib.call("cmbl_lenseflow_apply", cache.handle, op, P(x), P(out), st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
s = 8 if dtype == "f64" else 4
AL = 28 * (7 * NPOL * NB + 2 * NB) * N * N * s
print(f"N={N} NPOL={NPOL} NB={NB} path={lib.cdll.cmbl_lenseflow_kernel_path(cache.handle)} ", end="")
print(f"{dtype} op{op} chunk={os.environ.get('CMBL_FLOW_CHUNK','all')} tile={os.environ.get('CMBL_TILE_KB','70')}KB: {ms:.3f} ms/apply  alg {AL/ms/1e6:.0f} GB/s  frac {AL/ms/1e6/6552.6:.3f}  finite={bool(torch.isfinite(out).all())}")
lib.cdll.cmbl_profile_begin.restype = ctypes.c_int
lib.cdll.cmbl_profile_begin()
run(); run()
for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
    nm, cnt, tot = line.split()
    print(f"   {nm}: {int(cnt)} launches, avg {float(tot)/int(cnt)*1e3:.1f} us")

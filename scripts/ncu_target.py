"""Small target for ncu: one LenseFlow apply (and optionally one adjoint apply / CG step) at the bench workload.
usage: python scripts/ncu_target.py [f64|f32] [fwd|adj|cg]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
what = sys.argv[2] if len(sys.argv) > 2 else "fwd"
tT = torch.float64 if dtype == "f64" else torch.float32
N, NB, NPOL = int(os.environ.get("N", "1024")), int(os.environ.get("NB", "8")), int(os.environ.get("NPOL", "2"))
proj = pkg.ProjLambert(N, N, 2.0, tT, "cuda:0")
phi = pkg.Field("Map", torch.randn((NB, 1, N, N), dtype=tT, device="cuda:0") * 1e-6, proj)
f = pkg.Field(("Map", "QUMap", "IQUMap")[NPOL - 1], torch.randn((NB, NPOL, N, N), dtype=tT, device="cuda:0"), proj)
L = pkg.LenseFlow(phi, 7)
if what == "fwd":
    o = L * f
else:
    o = L.H * pkg.convert(f, ("Fourier", "QUFourier", "IQUFourier")[NPOL - 1])
torch.cuda.synchronize()
print("done", float(o.arr.abs().mean()))

"""Development aid: device time of the batched 2-D transforms (cmbl_rfft2 / cmbl_irfft2) with the per-kernel split.
usage: python scripts/time_fft.py [f64|f32] ; env N (1024), C (16 planes)"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
tT, cT, s = (torch.float64, torch.complex128, 8) if dtype == "f64" else (torch.float32, torch.complex64, 4)
N, C = int(os.environ.get("N", "1024")), int(os.environ.get("C", "16"))
proj = pkg.ProjLambert(N, N, 2.0, tT, "cuda:0")
lib = pkg.load()
m = torch.randn((C, N, N), dtype=tT, device="cuda:0"); F = torch.empty((C, N, N // 2 + 1), dtype=cT, device="cuda:0"); back = torch.empty_like(m)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
fwd = lambda: lib.call("cmbl_rfft2", proj.handle, P(m), P(F), C, st)
inv = lambda: lib.call("cmbl_irfft2", proj.handle, P(F), P(back), C, st)
for _ in range(3): fwd(); inv()
torch.cuda.synchronize()
err = float((back - m).abs().max())
ref = torch.fft.rfft2(m[:2])                                            # arrays are [c][x][y]: y is the fast (half-spectrum) axis
rel = float((F[:2] - ref).abs().max() / ref.abs().max())
for name, fn in (("rfft2", fwd), ("irfft2", inv)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    alg = C * (N * N * s + N * (N // 2 + 1) * 2 * s)                    # one real plane in, one half-spectrum out (or the reverse)
    print(f"N={N} C={C} {dtype} {name}: {ms*1e3:.1f} us  algorithmic {alg/ms/1e6:.0f} GB/s = {alg/ms/1e6/6552.6:.3f} of the measured roofline", end="")
    print(f"   (round trip max err {err:.2e}, vs cuFFT rel {rel:.2e})" if name == "rfft2" else "")
lib.cdll.cmbl_profile_begin.restype = ctypes.c_int
lib.cdll.cmbl_profile_begin()
fwd(); inv()
for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
    nm, cnt, tot = line.split()
    print(f"   {nm}: {int(cnt)} launches, avg {float(tot)/int(cnt)*1e3:.1f} us")

"""Development aid: end-to-end time of cmbl_lenseflow_apply_host (pinned host buffers) at the bench workload for the
current CMBL_HOST_CHUNKS; checks the result against the device-resident apply bit for bit.  usage: time_e2e.py [f64|f32]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
tT = torch.float64 if dtype == "f64" else torch.float32
N, NB = 1024, 8
proj = pkg.ProjLambert(N, N, 2.0, tT, "cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(1)
phi = pkg.Field("Map", torch.randn((NB, 1, N, N), dtype=tT, device="cuda:0", generator=gen) * 1e-6, proj)
f = pkg.Field("QUMap", torch.randn((NB, 2, N, N), dtype=tT, device="cuda:0", generator=gen), proj)
cache = pkg.LenseFlow(phi, 7).cache(f)
lib = pkg.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
out = torch.empty_like(f.arr)
lib.call("cmbl_lenseflow_apply", cache.handle, 0, P(f.arr), P(out), st)
hin = torch.empty(f.arr.shape, dtype=tT).pin_memory(); hin.copy_(f.arr)
hout = torch.empty(f.arr.shape, dtype=tT).pin_memory()
for op in (0, 2):
    run = lambda: lib.call("cmbl_lenseflow_apply_host", cache.handle, op, P(hin), P(hout), st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ok = ""
    if op == 0:
        ok = f" max|host-dev|={float((hout.to('cuda:0') - out).abs().max()):.1e}"
    print(f"{dtype} op{op} CMBL_HOST_CHUNKS={os.environ.get('CMBL_HOST_CHUNKS', 'default(4)')}: {ms:.3f} ms/apply e2e = {1e3/ms:.1f} applies/s{ok}")

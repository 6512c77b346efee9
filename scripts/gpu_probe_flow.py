"""GPU probe (development aid): parity of rfft2/irfft2/LenseFlow vs the oracle at small sizes, then device timing of the
stage kernels at Nside=1024.  Run on the GPU box: python scripts/gpu_probe_flow.py"""
import ctypes, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cmbl_oracle as O
lib = ctypes.CDLL(os.path.join(ROOT, "cmblensing.jl_b200", "libcmbl_b200.so"))
lib.cmbl_last_error.restype = ctypes.c_char_p
lib.cmbl_launch_count.restype = ctypes.c_longlong
def chk(r):
    if r != 0: raise RuntimeError(lib.cmbl_last_error().decode())
P = lambda t: ctypes.c_void_p(t.data_ptr())
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {}
rng = np.random.default_rng(0)
# ---- parity ------------------------------------------------------------------------------------
for dt, T, tT, tol in [(1, np.float64, torch.float64, 1e-11), (0, np.float32, torch.float32, 1e-4)]:
    for Ny, Nx, Npol, Nb, Nbphi in [(32, 16, 1, 1, 1), (64, 128, 2, 3, 3), (256, 256, 2, 2, 1), (1024, 512, 1, 1, 1)]:
        nst = 4
        d = O.make_dataset(Ny, Nx, 2.0, pol='I' if Npol == 1 else 'P', T=T, nb=Nb, seed=3, nsteps=nst, mask=False)
        proj = d['proj']; cT = proj.cT
        phi = d['phi'][:Nbphi]
        L = O.precompute(proj, phi, nst, phi_is_fourier=True)
        f = O.to_lense_basis('I' if Npol == 1 else 'P', proj, d['f']).astype(T)
        plan = ctypes.c_void_p(); flow = ctypes.c_void_p()
        chk(lib.cmbl_plan_create(ctypes.byref(plan), 0, Ny, Nx, ctypes.c_double(2.0), dt))
        chk(lib.cmbl_lenseflow_create(ctypes.byref(flow), plan, nst, Npol, Nb, Nbphi))
        phid = torch.from_numpy(np.ascontiguousarray(phi.astype(cT))).to(dev)
        chk(lib.cmbl_lenseflow_precompute(flow, P(phid), 1, 0, st()))
        fd = torch.from_numpy(np.ascontiguousarray(f)).to(dev)
        Fd = torch.zeros((Nb, Npol, Nx, Ny // 2 + 1), dtype=torch.complex128 if dt else torch.complex64, device=dev)
        chk(lib.cmbl_rfft2(plan, P(fd), P(Fd), Nb * Npol, st()))
        ref = O.rfft2(f.astype(np.float64))
        e_r = float(np.abs(Fd.cpu().numpy() - ref).max() / np.abs(ref).max())
        bd = torch.zeros_like(fd)
        chk(lib.cmbl_irfft2(plan, P(Fd), P(bd), Nb * Npol, st()))
        e_i = float(np.abs(bd.cpu().numpy() - f).max() / np.abs(f).max())
        errs = {}
        for op in range(4):
            if op in (0, 2): x = np.ascontiguousarray(f)
            else:
                F0 = O.rfft2(f)
                x = np.ascontiguousarray((F0 + 0.1 * np.abs(F0).mean() * (rng.standard_normal(F0.shape) + 1j * rng.standard_normal(F0.shape))).astype(cT))
            refo = O.lenseflow_apply(L, op, x)
            xd = torch.from_numpy(x).to(dev); od = torch.zeros_like(xd)
            chk(lib.cmbl_lenseflow_apply(flow, op, P(xd), P(od), st()))
            torch.cuda.synchronize()
            o = od.cpu().numpy()
            errs[op] = float(np.sqrt((np.abs(o - refo) ** 2).sum() / (np.abs(refo) ** 2).sum()))
        ok = max(e_r, e_i, *errs.values()) < tol
        print("parity", dt, Ny, Nx, Npol, Nb, Nbphi, "rfft %.1e irfft %.1e" % (e_r, e_i), errs, "OK" if ok else "FAIL", flush=True)
        out["parity_%d_%d_%d_%d_%d" % (dt, Ny, Nx, Npol, Nb)] = dict(rfft=e_r, irfft=e_i, ops=errs, ok=ok)
        chk(lib.cmbl_lenseflow_destroy(flow)); chk(lib.cmbl_plan_destroy(plan))
# ---- timing ------------------------------------------------------------------------------------
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dt, tT, cT, s in [(0, torch.float32, torch.complex64, 4), (1, torch.float64, torch.complex128, 8)]:
    for Npol, Nb, Nbphi in [(1, 1, 1), (2, 8, 8)]:
        N = 1024; C = Npol * Nb
        plan = ctypes.c_void_p(); flow = ctypes.c_void_p()
        chk(lib.cmbl_plan_create(ctypes.byref(plan), 0, N, N, ctypes.c_double(2.0), dt))
        chk(lib.cmbl_lenseflow_create(ctypes.byref(flow), plan, 7, Npol, Nb, Nbphi))
        phid = (torch.randn((Nbphi, 1, N, N), dtype=tT, device=dev) * 1e-6)
        chk(lib.cmbl_lenseflow_precompute(flow, P(phid), 0, 0, st()))
        fd = torch.randn((Nb, Npol, N, N), dtype=tT, device=dev); od = torch.zeros_like(fd)
        Fd = torch.zeros((Nb, Npol, N, N // 2 + 1), dtype=cT, device=dev); Od = torch.zeros_like(Fd)
        t_r = timeit(lambda: chk(lib.cmbl_rfft2(plan, P(fd), P(Fd), C, st())))
        t_i = timeit(lambda: chk(lib.cmbl_irfft2(plan, P(Fd), P(od), C, st())))
        t_L = timeit(lambda: chk(lib.cmbl_lenseflow_apply(flow, 0, P(fd), P(od), st())), n=3, warm=1)
        t_LH = timeit(lambda: chk(lib.cmbl_lenseflow_apply(flow, 1, P(Fd), P(Od), st())), n=3, warm=1)
        t_cu = timeit(lambda: torch.fft.rfft2(fd))
        passb = N * N * s
        A_L = 28 * (7 * C + 2 * Nbphi) * passb
        r = dict(rfft2_ms=t_r, irfft2_ms=t_i, L_ms=t_L, LH_ms=t_LH, cufft_rfft2_ms=t_cu,
                 rfft2_GBs=C * passb * 2 / t_r / 1e6, L_alg_GBs=A_L / t_L / 1e6, LH_alg_GBs=A_L / t_LH / 1e6)
        print("time", dt, Npol, Nb, json.dumps(r), flush=True)
        out["time_%d_%d_%d" % (dt, Npol, Nb)] = r
        chk(lib.cmbl_lenseflow_destroy(flow)); chk(lib.cmbl_plan_destroy(plan))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_flow.json"), "w"), indent=1)

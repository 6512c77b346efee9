"""MAP_joint with the batch sharded over the GPUs of one box (BASELINE config 4 shape: IQU fields do not have a quadratic-estimate
Nϕ here, so the polarisation-only dataset is used): every rank owns NB_LOCAL batch items, the CG keeps the reference's lock-step
stopping rule through one flag all-reduce per iteration and the line search sums logpdf over ranks (one scalar all-reduce per
evaluation) — NCCL carries 8 bytes at a time, no field ever leaves its GPU.
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/map_joint_multi.py [f32|f64] [Nside] [NB_LOCAL] [steps]"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g
import cmbl_oracle as O                      # only Cℓ tables / mask profile for the synthetic inputs
pkg = g.load_package()
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 1
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
tT = torch.float64 if dtype == "f64" else torch.float32
proj = pkg.ProjLambert(N, N, 2.0, tT, dev)
cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
w = lambda p: pkg.Field(("Map", "QUMap")[p - 1], torch.randn((NB, p, N, N), dtype=tT, device=dev, generator=gen), proj)
nT = O.noise_cls(ell); one = np.ones_like(nT); lb, wl = O.lowpass_wl(3000)
Cf = pkg.Cℓ_to_Cov("P", proj, ell, cls["ut_EE"], cls["ut_BB"]); Cft = pkg.Cℓ_to_Cov("P", proj, ell, cls["tot_EE"], cls["tot_BB"])
Cn = pkg.Cℓ_to_Cov("P", proj, ell, 2 * nT, 2 * nT); Mf = pkg.Cℓ_to_Cov("P", proj, lb, wl, wl, units=1); B = pkg.Cℓ_to_Cov("P", proj, ell, one, one, units=1)
Cϕ = pkg.Cℓ_to_Cov("I", proj, ell, cls["pp"])
mask = torch.from_numpy(O.cosine_border_mask(O.ProjLambert(N, N, 2.0, np.float32 if dtype == "f32" else np.float64), 1.0))
Mpix = pkg.DiagOp(pkg.Field("QUMap", mask[None, None].expand(1, 2, N, N).contiguous(), proj))
ϕ_true = pkg.DiagOp(pkg.Field("Fourier", torch.sqrt(Cϕ._real), proj)) * w(1)
ds0 = pkg.BaseDataSet(pkg.HarmonicBasis(w(2)), Cf, Cn, B, Mf, Mpix, nsteps=7, Cϕ=Cϕ, Cf̃=Cft)
sim = pkg.simulate(ds0, ϕ_true, generator=gen)
ds = pkg.BaseDataSet(sim["d"], Cf, Cn, B, Mf, Mpix, nsteps=7, Cϕ=Cϕ, Cf̃=Cft)
qe = pkg.quadratic_estimate(ds)
ds.Nϕ = pkg.DiagOp(pkg.Field("Fourier", (qe["Nϕ"]._real / 2).to(proj.cT), proj)); ds.D = pkg.mixing_D(ds)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
f, ϕ, hist = pkg.MAP_joint(ds, nsteps=steps, conjgrad_kwargs=dict(tol=1e-1, nsteps=500), group=dist.group.WORLD)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
a, b = pkg.Map(ϕ).arr, pkg.Map(ϕ_true).arr
q = slice(N // 4, 3 * N // 4)
cc = [float(torch.corrcoef(torch.stack([a[i, 0, q, q].flatten(), b[i, 0, q, q].flatten()]))[0, 1]) for i in range(NB)]
out = [None] * world
dist.all_gather_object(out, dict(rank=rank, alphas=[round(h["α"], 6) for h in hist], cg=[h["cg_iters"] for h in hist], corr=np.round(cc, 3).tolist()))
if rank == 0:
    print(f"MAP_joint sharded over {world} GPU(s) (NCCL): {dtype} Nside={N} QU, {NB} item(s) per GPU, {steps} steps in {dt:.2f} s = {dt/steps:.2f} s/step")
    for o in out:
        print("   ", o)
    assert all(o["alphas"] == out[0]["alphas"] and o["cg"] == out[0]["cg"] for o in out), "ranks must take the same α and CG iteration counts"
    print("    every rank took the same step lengths and CG iteration counts")
dist.destroy_process_group()

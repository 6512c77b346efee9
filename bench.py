#!/usr/bin/env python
"""bench.py — headline benchmark of the flat-sky hot path (BASELINE.json: "LenseFlow applies/sec and CG-Wiener iters/sec at
Nside=1024 batch=8; HBM GB/s vs roofline").

Workload (config.workload): Nside=1024, QU polarisation, batch=8 with 8 distinct ϕ (BASELINE configs[2]); RK4 with 7 steps.
A "step" is ONE batched LenseFlow apply  Lϕ*f  over the whole batch-8 QU field (16 planes, 28 RK stages).  The same line
also carries the CG-Wiener iteration rate (`cg`), measured in the same run on the same workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

N>1: launched by torchrun, one rank per GPU; every rank applies its own independent batch-8 field (weak scaling — the path
shards over independent batch items with no data-path collective); time = max over ranks.
`--impl reference`: the reference algorithm on the host cores (the NumPy oracle port — Julia is not installed; FFT
provider = the faster of scipy.fft/pocketfft and torch.fft/MKL, timed on the spot), same metric and workload string; each step
is the WHOLE batch-8 apply (unscaled) whenever K+W such steps fit in ~4 minutes, otherwise the stated subset of items.  Under
torchrun only rank 0 runs it (one CPU job on the box, NOT multiplied by N); the other ranks exit 0.

Extra keys of the b200 line (all measured in the same run, at the N ranks of the launch):
  cg        CG-Wiener iterations/s on the headline workload (>= 20 iterations timed) + its own cpu_baseline
  map_joint BASELINE configs[3]: Nside=2048 IQU, one batch item per GPU, MAP_joint steps with the scalar all-reduce of the line
            search (and the CG stop flag) inside the timed region
  hmc       BASELINE configs[4]: Nside=512, 8 chains per GPU, leap-frog steps/s of the HMC update of ϕ°
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDE, NPOL, NB, NSTEPS_RK, THETA = 1024, 2, 8, 7, 2.0
WORKLOAD = f"LenseFlow apply Lphi*f, Nside={NSIDE} QU batch={NB} (Cphi={NB}) per GPU, RK4 n={NSTEPS_RK}, theta_pix={THETA}'"


def algorithmic_bytes(s):
    """pass = s·Ny·Nx bytes; LenseFlow stage = 7C + 2Cϕ passes (row kernel 2C, column kernel 5C + 2Cϕ);
    apply = 4·n stages; CG iteration = 2 applies + (22C + 8) passes."""
    C, Cphi = NPOL * NB, NB
    p = s * NSIDE * NSIDE
    stage_rows, stage_cols = 2 * C * p, (5 * C + 2 * Cphi) * p
    apply_b = 4 * NSTEPS_RK * (stage_rows + stage_cols)
    return dict(pass_bytes=p, rows=stage_rows, cols=stage_cols, apply=apply_b, cg_iter=2 * apply_b + (22 * C + 8) * p)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML poll every ~5 ms between begin()
    and end() (a 10-step timed region lasts < 100 ms — too short for `nvidia-smi -lms`, whose first sample arrives after
    ~1 s); falls back to one `nvidia-smi` query issued while the GPU is kept busy when NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.mx, self.power = index, [], 0, None, []
        self._run, self._thr, self.nv, self.h = False, None, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve by PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if nv.nvmlDeviceGetPciInfo(hi).bus == bus:
                        self.h = hi
                        break
            self.mx = int(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while self._run:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def begin(self):
        if self.nv is not None:
            sys.setswitchinterval(0.0005)             # let the poll thread in between the launch calls of the timed loop
            self._run = True
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def end(self):
        if self.nv is not None:
            self._run = False
            self._thr.join()
            sys.setswitchinterval(0.005)
            try:
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            except Exception:
                pass

    def smi_fallback(self):
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20)
            v = [x.strip() for x in r.stdout.strip().splitlines()[0].split(",")]
            reasons = [n for n, a in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), v[2:6]) if a == "Active"]
            return {"sm_mhz": int(v[0]), "sm_max_mhz": int(v[1]), "reasons": reasons, "samples": 1, "source": "nvidia-smi (one query under load)"}
        except Exception as e:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock query unavailable: {e}"], "samples": 0}

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b),
                "samples": len(sm), "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(self.power) if self.power else None,
                "source": "NVML poll (5 ms) inside the timed region"}


def csrc_sha16():
    """sha256[:16] over the kernel sources the library is built from (sorted csrc/*.cu, *.cuh and the Makefile)."""
    import glob, hashlib
    d = os.path.join(ROOT, "cmblensing.jl_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + [os.path.join(d, "Makefile")]):
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def numa_interleave():
    """Spread this process's future page allocations (the pinned host buffers of the e2e leg) over all NUMA nodes: at 8 ranks the host side
    of the e2e path moves ~300 GB/s, more than one socket's DRAM delivers.  set_mempolicy(MPOL_INTERLEAVE); returns the node count or None."""
    try:
        nodes = [int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if len(nodes) < 2:
            return len(nodes)
        mask = ctypes.c_ulong(sum(1 << n for n in nodes))
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.syscall(238, 3, ctypes.byref(mask), ctypes.c_ulong(max(nodes) + 2))      # SYS_set_mempolicy (x86-64), MPOL_INTERLEAVE
        return len(nodes) if rc == 0 else None
    except Exception:
        return None


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's CPU algorithm (oracle port) on all host threads of the box, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import cmbl_oracle as O
    cores = os.cpu_count() or 1
    npT = np.float64 if args.dtype == "f64" else np.float32
    proj = O.ProjLambert(NSIDE, NSIDE, THETA, npT)
    rng = np.random.default_rng(0)
    cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
    # same synthetic inputs as the b200 arm: ϕ and (E,B) drawn from the fiducial spectra, 8 distinct ϕ
    phi = O.simulate_diag(proj, O.cl_to_cov(proj, ell, cls["pp"])[None, None], rng, nb=NB)
    Cf = np.stack([O.cl_to_cov(proj, ell, cls[k]) for k in ("ut_EE", "ut_BB")])[None]
    f = O.to_lense_basis("P", proj, O.simulate_diag(proj, Cf, rng, nb=NB)).astype(npT)
    # FFT provider: pocketfft vs MKL (the reference's recommended provider, README.md:56) on one item, keep the faster
    t_be = {}
    for be in ("pocketfft", "mkl"):
        try:
            O.set_fft_backend(be); O.set_workers(cores)
            L1 = O.precompute(proj, phi[:1], 1, phi_is_fourier=True)
            O.lenseflow_apply(L1, O.OP_L, f[:1])
            t0 = time.perf_counter(); O.lenseflow_apply(L1, O.OP_L, f[:1]); t_be[be] = time.perf_counter() - t0
        except Exception as e:                                     # torch missing on the box: pocketfft only
            t_be[be] = float("inf")
    best = min(t_be, key=t_be.get)
    O.set_fft_backend(best); O.set_workers(cores)
    # the warm-up step is timed: if K steps of the whole batch would not end within ~4 minutes, halve the number of items and say so
    W = 1
    nit = NB
    while True:
        L = O.precompute(proj, phi[:nit], NSTEPS_RK, phi_is_fourier=True)
        t0 = time.perf_counter()
        O.lenseflow_apply(L, O.OP_L, f[:nit])
        t_w = time.perf_counter() - t0
        if nit == 1 or t_w * args.steps <= 240.0:
            break
        nit //= 2
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.lenseflow_apply(L, O.OP_L, f[:nit])
    dt = (time.perf_counter() - t0) / args.steps
    ms = dt * (NB / nit) * 1e3
    val = 1e3 / ms
    sample = (f"the whole workload per step (all {NB} batch items, unscaled)" if nit == NB else
              f"{nit} of {NB} batch items per step, time scaled x{NB // nit} (the whole batch would not fit {args.steps}+{W} steps in 4 minutes)")
    sample += f"; NumPy + {best} ({'scipy.fft' if best == 'pocketfft' else 'torch.fft/MKL'}) on {cores} threads; one-item n=1 probe: " + \
              ", ".join(f"{k} {v*1e3:.0f} ms" for k, v in t_be.items())
    print_line({
        "impl": "reference", "metric": "lenseflow_batched_applies_per_sec", "value": val, "unit": "applies/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": "applies/s", "cores": cores, "kind": "port", "sample": sample, "fft": best},
        "e2e": {"value": val, "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ranks": {"world": world, "ran_on": "rank 0 only",
                  "note": "ONE CPU job on the box's host cores whatever N is: the value is per batch-8 apply and is not multiplied by the number of GPUs"},
    })


def build_sim_dataset(pkg, O, N, pol, nb, tT, dev, gen, theta=THETA):
    """load_sim-like synthetic dataset on the device (src/dataset.jl:186-340): fiducial spectra, 3 µK-arcmin noise with a knee,
    LowPass(3000) Fourier mask, cosine-apodised 1° pixel mask, d = M B L(ϕ) f + n; Nϕ = quadratic_estimate(ds).Nϕ / 2 and the
    mixing matrix D of load_sim (:316-332).  Returns (ds, ϕ_true)."""
    import numpy as np
    import torch
    proj = pkg.ProjLambert(N, N, theta, tT, dev)
    cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
    npol = {"P": 2, "IP": 3}[pol]; lense = ("Map", "QUMap", "IQUMap")[npol - 1]
    w = lambda p: pkg.Field(("Map", "QUMap", "IQUMap")[p - 1], torch.randn((nb, p, N, N), dtype=tT, device=dev, generator=gen), proj)
    nT = O.noise_cls(ell); zero = np.zeros_like(nT); one = np.ones_like(nT); lb, wl = O.lowpass_wl(3000)
    if pol == "IP":
        Cf = pkg.Cℓ_to_Cov("IP", proj, ell, cls["ut_TT"], cls["ut_EE"], cls["ut_BB"], cls["ut_TE"])
        Cft = pkg.Cℓ_to_Cov("IP", proj, ell, cls["tot_TT"], cls["tot_EE"], cls["tot_BB"], cls["tot_TE"])
        Cn = pkg.Cℓ_to_Cov("IP", proj, ell, nT, 2 * nT, 2 * nT, zero)
        Mf = pkg.Cℓ_to_Cov("IP", proj, lb, wl, wl, wl, np.zeros_like(wl), units=1)
        B = pkg.Cℓ_to_Cov("IP", proj, ell, one, one, one, zero, units=1)
    else:
        Cf = pkg.Cℓ_to_Cov("P", proj, ell, cls["ut_EE"], cls["ut_BB"]); Cft = pkg.Cℓ_to_Cov("P", proj, ell, cls["tot_EE"], cls["tot_BB"])
        Cn = pkg.Cℓ_to_Cov("P", proj, ell, 2 * nT, 2 * nT); Mf = pkg.Cℓ_to_Cov("P", proj, lb, wl, wl, units=1)
        B = pkg.Cℓ_to_Cov("P", proj, ell, one, one, units=1)
    Cϕ = pkg.Cℓ_to_Cov("I", proj, ell, cls["pp"])
    mask = torch.from_numpy(O.cosine_border_mask(O.ProjLambert(N, N, theta, np.float32 if tT == torch.float32 else np.float64), 1.0))
    Mpix = pkg.DiagOp(pkg.Field(lense, mask[None, None].expand(1, npol, N, N).contiguous(), proj))
    ϕ_true = pkg.DiagOp(pkg.Field("Fourier", torch.sqrt(Cϕ._real), proj)) * w(1)
    ds0 = pkg.BaseDataSet(pkg.HarmonicBasis(w(npol)), Cf, Cn, B, Mf, Mpix, nsteps=NSTEPS_RK, Cϕ=Cϕ, Cf̃=Cft)
    sim = pkg.simulate(ds0, ϕ_true, generator=gen)
    ds = pkg.BaseDataSet(sim["d"], Cf, Cn, B, Mf, Mpix, nsteps=NSTEPS_RK, Cϕ=Cϕ, Cf̃=Cft)
    qe = pkg.quadratic_estimate(ds)
    ds.Nϕ = pkg.DiagOp(pkg.Field("Fourier", (qe["Nϕ"]._real / 2).to(proj.cT), proj))
    ds.D = pkg.mixing_D(ds)
    return ds, ϕ_true


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cg-iters", type=int, default=20)
    ap.add_argument("--skip", default="", help="comma list of optional sections to skip: cg,map_joint,hmc,other,cpu")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    skip = set(x for x in args.skip.split(",") if x)
    if args.no_cpu_baseline:
        skip.add("cpu")
    # stdout carries exactly ONE JSON line: library chatter (e.g. "NCCL version ..." printed at communicator creation) is sent
    # to stderr by pointing fd 1 at fd 2 for the duration of the run; the line itself is written to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_line
    def print_line(obj):
        _real_stdout.write(json.dumps(obj) + "\n"); _real_stdout.flush()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = pkg.load()
    tT = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if args.dtype == "f64" else 4
    AB = algorithmic_bytes(s)
    peak, peak_src = measured_peak()
    peak_gbs = peak
    proj = pkg.ProjLambert(NSIDE, NSIDE, THETA, tT, dev)

    # ---- synthetic inputs (seeded per rank), SURVEY §8(d): ϕ, f drawn from the fiducial spectra ---------------------
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cmbl_oracle as O                                   # only for Cℓ→2-D setup tables and the cpu_baseline leg
    npT = np.float64 if args.dtype == "f64" else np.float32
    op = O.ProjLambert(NSIDE, NSIDE, THETA, npT)
    cls = O.load_fiducial_cls(); ell = cls["ell"].astype(float)
    dg = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Cphi = dg(O.cl_to_cov(op, ell, cls["pp"]))
    Cf_np = np.stack([O.cl_to_cov(op, ell, cls[k]) for k in ("ut_EE", "ut_BB")])[None]
    Cn_np = np.stack([O.cl_to_cov(op, ell, O.noise_cls(ell, pol=True)) for _ in range(2)])[None]
    lb, wl = O.lowpass_wl(3000)
    Mf_np = np.stack([O.cl_to_cov(op, lb, wl, units=1) for _ in range(2)])[None]
    B_np = np.ones_like(Mf_np)
    mask_np = np.broadcast_to(O.cosine_border_mask(op, 1.0), (1, 2, NSIDE, NSIDE)).copy()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    white = lambda n, p: torch.randn((n, p, NSIDE, NSIDE), dtype=tT, device=dev, generator=gen)
    ϕ = pkg.Fourier(pkg.Field("Map", white(NB, 1), proj)); ϕ = ϕ._like(ϕ.arr * torch.sqrt(Cphi))
    f = pkg.Fourier(pkg.Field("QUMap", white(NB, 2), proj)); f = pkg.Field("EBFourier", f.arr * torch.sqrt(dg(Cf_np)), proj)
    L = pkg.LenseFlow(ϕ, NSTEPS_RK)
    fmap = pkg.LenseBasis(f)
    cache = L.cache(fmap)
    D = lambda a, basis="EBFourier": pkg.DiagOp(pkg.Field(basis, dg(a), proj))
    noise = pkg.Fourier(pkg.Field("QUMap", white(NB, 2), proj)); noise = pkg.Field("EBFourier", noise.arr * torch.sqrt(dg(Cn_np)), proj)
    ds = pkg.BaseDataSet(f, D(Cf_np), D(Cn_np), D(B_np), D(Mf_np), D(mask_np, "QUMap"), L=L, nsteps=NSTEPS_RK)   # d replaced below
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.empty_like(fmap.arr)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def step_device():
        lib.call("cmbl_lenseflow_apply", cache.handle, 0, P(fmap.arr), P(out), st)

    # ---- device-resident timing ------------------------------------------------------------------------------------
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.launch_count()
        if sampler:
            sampler.begin()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if sampler:
            sampler.end()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, (lib.launch_count() - n0)

    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, launches = timed(step_device, args.steps, max(args.warmup, 3), sampler if (sampler and sampler.nv) else None)
    clocks = None
    if rank == 0:
        if sampler.nv is not None and sampler.sm:
            clocks = sampler.summary()
        else:                                            # keep the GPU busy with the same step while nvidia-smi answers
            import concurrent.futures as cf
            with cf.ThreadPoolExecutor(1) as ex:
                fut = ex.submit(sampler.smi_fallback)
                while not fut.done():
                    step_device(); torch.cuda.synchronize()
                clocks = fut.result()

    # ---- end-to-end: HOST buffers through the C ABI, copies inside the timed region ---------------------------------
    numa = numa_interleave() if (world >= 4 and os.environ.get("CMBL_BENCH_NUMA", "interleave") == "interleave") else None
    hin = torch.empty(fmap.arr.shape, dtype=tT).pin_memory(); hin.copy_(fmap.arr)
    hout = torch.empty(fmap.arr.shape, dtype=tT).pin_memory()

    def step_host():
        lib.call("cmbl_lenseflow_apply_host", cache.handle, 0, P(hin), P(hout), st)
    ms_e2e, _ = timed(step_host, max(3, args.steps // 2), 3)
    assert float((hout.to(dev) - out).abs().max()) == 0.0, "host path and device path disagree"
    # streaming variant of the same API: successive applies overlap their transfers (cmbl_lenseflow_apply_host_async, two staging
    # slots); every step still copies ITS input from pinned host memory and ITS result back — the copies are inside the timed region,
    # they just run during the neighbouring steps' integrations.  Results are checked after the final sync.
    hins = [hin, torch.empty_like(hin).pin_memory()]; hins[1].copy_(hin)
    houts = [hout, torch.empty_like(hout).pin_memory()]
    kstep = [0]

    def step_host_async():
        i = kstep[0] & 1; kstep[0] += 1
        lib.call("cmbl_lenseflow_apply_host_async", cache.handle, 0, P(hins[i]), P(houts[i]), st)

    def timed_stream(steps, warmup):
        for _ in range(warmup):
            step_host_async()
        lib.call("cmbl_lenseflow_host_sync"); barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host_async()
        lib.call("cmbl_lenseflow_host_sync")                # every out_host of the timed steps is valid here
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt / steps
    ms_e2e_stream = timed_stream(max(6, args.steps), 3)
    for ho in houts:
        assert float((ho.to(dev) - out).abs().max()) == 0.0, "streaming host path and device path disagree"
    # the same bytes with no compute in between (H2D and D2H back to back on one stream): what the host link alone costs
    stage = torch.empty_like(fmap.arr)

    def step_copy():
        stage.copy_(hin, non_blocking=True); hout.copy_(stage, non_blocking=True); torch.cuda.current_stream().synchronize()
    ms_copy, _ = timed(step_copy, max(3, args.steps // 2), 2)
    del stage

    # ---- CG-Wiener iterations on the same workload ------------------------------------------------------------------
    cg_line = None
    if "cg" not in skip:
        ft = L * fmap
        data = pkg.HarmonicBasis(ft)                                           # d = M B L f + n  (mask applied inside M)
        data = pkg.Field("EBFourier", data.arr + noise.arr, proj)
        ds = pkg.BaseDataSet(data, D(Cf_np), D(Cn_np), D(B_np), D(Mf_np), D(mask_np, "QUMap"), L=L, nsteps=NSTEPS_RK)
        h, *_ = ds._solver(ϕ)
        res = (ctypes.c_double * NB)()
        lib.call("cmbl_cg_begin", h, ctypes.c_void_p(0), 0, res, st)
        res0 = list(res)
        nul = ctypes.POINTER(ctypes.c_double)()                                # NULL: residuals stay on the device between polls
        def cg_iter():
            lib.call("cmbl_cg_step", h, nul, st)
        ms_cg, launches_cg = timed(cg_iter, max(args.cg_iters, 20), 3)
        ms_cg_poll, _ = timed(lambda: lib.call("cmbl_cg_step", h, res, st), 8, 1)      # host reads res after every iteration
        res1 = list(res)
        assert all(r == r and r > 0 for r in res1), "CG residual went non-finite"
        cg_line = {"metric": "cg_wiener_iters_per_sec", "value": world * 1e3 / ms_cg, "unit": "iters/s", "ms_per_iter": ms_cg,
                   "iters_timed": max(args.cg_iters, 20), "gpu_launches": launches_cg, "algorithmic_GBs": AB["cg_iter"] / ms_cg / 1e6,
                   "frac": AB["cg_iter"] / ms_cg / 1e6 / peak_gbs, "frac_of_8TBs_nominal": AB["cg_iter"] / ms_cg / 1e6 / 8000.0,
                   "res_first": res0[0], "res_last": res1[0], "cpu_baseline": None, "ms_per_iter_host_poll_every_iter": ms_cg_poll,
                   "residual_polling": "cmbl_cg_step(res_host=NULL): α, β, res stay on the device; the host reads them only when it asks"}
        if rank == 0 and world == 1 and "cpu" not in skip:
            # the oracle's CG operator on ONE batch item, one iteration's worth of work (A·p = gradientf_logpdf(p, d=0): Lϕ, the
            # masking chain, Lϕ'), scaled to the batch of 8
            cores = os.cpu_count() or 1
            O.set_workers(cores)
            dso = O.DataSet(proj=op, pol="P", d=data.arr[:1].cpu().numpy(), Cf=Cf_np, Cn=Cn_np, Cnhat=Cn_np, B=B_np, Bhat=B_np, Mf=Mf_np, Mpix=mask_np,
                            L=O.precompute(op, ϕ.arr[:1].cpu().numpy(), NSTEPS_RK, phi_is_fourier=True))
            pvec = f.arr[:1].cpu().numpy()
            t0 = time.perf_counter()
            O.gradientf_logpdf(dso, pvec, np.zeros_like(pvec))
            dtc = time.perf_counter() - t0
            cg_line["cpu_baseline"] = {"value": 1.0 / (dtc * NB), "unit": "iters/s", "cores": cores, "kind": "port", "seconds_sample": dtc,
                                       "sample": f"1 of {NB} batch items, one operator application A*p (the >97 % of a CG iteration, numerical_algorithms.jl:99), time scaled x{NB}; NumPy + scipy.fft workers={cores}",
                                       "gpu_over_cpu": (1e3 / ms_cg) * dtc * NB}

    # ---- per-kernel durations (CUDA events around each launch, on the launching stream) for the roofline ------------
    barrier()
    lib.cdll.cmbl_profile_begin()
    step_device(); step_device()
    prof = {}
    for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
        nm, cnt, tot = line.split()
        prof[nm] = (int(cnt), float(tot))
    tot_prof = sum(v[1] for v in prof.values())
    # algorithmic bytes per launch: row kernel 2C passes, column kernel 5C + 2Cϕ passes; the two layout
    # conversions of an apply (2C passes each) are overhead outside the model
    kbytes = {"flow_rows": AB["rows"], "flow_cols": AB["cols"], "layout_to_rg": 2 * NPOL * NB * AB["pass_bytes"], "layout_from_rg": 2 * NPOL * NB * AB["pass_bytes"]}
    dom = max(("flow_cols", "flow_rows"), key=lambda k: prof[k][1])
    dom_ms, dom_bytes = prof[dom][1] / prof[dom][0], kbytes[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / dom_ms / 1e6, "peak": peak, "unit": "GB/s",
                "frac": dom_bytes / dom_ms / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                "share_of_step": prof[dom][1] / tot_prof,
                "algorithmic_bytes_per_launch": dom_bytes,
                "kernels": {k: {"launches_per_step": v[0] // 2, "avg_ms": v[1] / v[0], "share": v[1] / tot_prof,
                                "algorithmic_GBs": (kbytes[k] / (v[1] / v[0]) / 1e6 if k in kbytes else None)} for k, v in prof.items()}}
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture — reported only when that capture was taken from the
    # code that is loaded now: the same library file (sha256 of libcmbl_b200.so) or, because nvcc builds are not bit-reproducible, a library
    # built from the same kernel sources (sha256 over csrc/*.cu, *.cuh, Makefile recorded next to the numbers); else null
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            import hashlib
            tj = json.load(open(traffic_file))
            sha = hashlib.sha256(open(lib.path, "rb").read()).hexdigest()[:16]
            src = csrc_sha16()
            roofline["traffic_capture"] = {"binary_sha16": tj.get("binary_sha16"), "loaded_sha16": sha, "source_sha16": tj.get("source_sha16"),
                                           "tree_source_sha16": src, "source": tj.get("source")}
            if tj.get("binary_sha16") == sha or (tj.get("source_sha16") and tj.get("source_sha16") == src):
                roofline["traffic"] = tj.get(args.dtype, {}).get(dom)
        except Exception:
            pass
    apply_gbs = AB["apply"] / ms_step / 1e6

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample --------------------------
    cpu = None
    if rank == 0 and world == 1 and "cpu" not in skip:
        cores = os.cpu_count() or 1
        O.set_workers(cores)
        nit = 2
        phi_np = ϕ.arr[:nit].cpu().numpy()
        Lo = O.precompute(op, phi_np, NSTEPS_RK, phi_is_fourier=True)
        f_np = fmap.arr[:nit].cpu().numpy()
        t0 = time.perf_counter()
        ref = O.lenseflow_apply(Lo, O.OP_L, f_np)
        dt = time.perf_counter() - t0
        err = float(np.linalg.norm(out[:nit].cpu().numpy() - ref) / np.linalg.norm(ref))
        cpu = {"value": 1.0 / (dt * NB / nit), "unit": "applies/s", "cores": cores, "kind": "port",
               "sample": f"{nit} of {NB} batch items of the same workload, one apply, time scaled x{NB // nit}; NumPy + scipy.fft(pocketfft) workers={cores}",
               "seconds_sample": dt, "gpu_vs_oracle_rel_l2": err}

    # ---- the same apply in the other precision (context for the headline; Float32 is what the reference runs on GPUs) ----
    other = None
    if rank == 0 and world == 1 and "other" not in skip:
        oT, odt, osz = (torch.float32, "f32", 4) if args.dtype == "f64" else (torch.float64, "f64", 8)
        proj2 = pkg.ProjLambert(NSIDE, NSIDE, THETA, oT, dev)
        ϕ2 = pkg.Field("Fourier", ϕ.arr.to(torch.complex64 if oT == torch.float32 else torch.complex128), proj2)
        f2 = pkg.Field("QUMap", fmap.arr.to(oT), proj2)
        c2 = pkg.LenseFlow(ϕ2, NSTEPS_RK).cache(f2)
        out2 = torch.empty_like(f2.arr)
        ms2, _ = timed(lambda: lib.call("cmbl_lenseflow_apply", c2.handle, 0, P(f2.arr), P(out2), st), args.steps, 3)
        AB2 = algorithmic_bytes(osz)
        other = {"dtype": odt, "value": 1e3 / ms2, "unit": "applies/s", "ms_per_step": ms2, "apply_algorithmic_GBs": AB2["apply"] / ms2 / 1e6,
                 "apply_frac_of_measured_peak": AB2["apply"] / ms2 / 1e6 / peak}
        del c2, out2, f2, ϕ2

    # ---- the batched 2-D transforms on their own (m_rfft! / m_irfft!, src/util_fft.jl:26-27): 16 planes of the headline shape ------------
    fft_line = None
    if rank == 0 and world == 1 and "other" not in skip:
        C16 = NPOL * NB
        fm16 = fmap.arr.reshape(C16, NSIDE, NSIDE)
        F16 = torch.empty((C16, NSIDE, NSIDE // 2 + 1), dtype=proj.cT, device=dev); back16 = torch.empty_like(fm16)
        ms_r, _ = timed(lambda: lib.call("cmbl_rfft2", proj.handle, P(fm16), P(F16), C16, st), 20, 3)
        ms_i, _ = timed(lambda: lib.call("cmbl_irfft2", proj.handle, P(F16), P(back16), C16, st), 20, 3)
        plane, half = NSIDE * NSIDE * s, NSIDE * (NSIDE // 2 + 1) * 2 * s
        per = C16 * (plane + 3 * half)                                  # column pass: plane <-> half-spectrum; row pass: half-spectrum in and out
        fft_line = {"workload": f"cmbl_rfft2 / cmbl_irfft2, {C16} planes of {NSIDE}x{NSIDE}, {args.dtype}", "rfft2_us": ms_r * 1e3, "irfft2_us": ms_i * 1e3,
                    "algorithmic_bytes_two_passes": per, "rfft2_GBs": per / ms_r / 1e6, "irfft2_GBs": per / ms_i / 1e6,
                    "rfft2_frac_of_measured_peak": per / ms_r / 1e6 / peak, "irfft2_frac_of_measured_peak": per / ms_i / 1e6 / peak,
                    "round_trip_max_abs_err": float((back16 - fm16).abs().max())}
        del F16, back16

    # free the headline workload before the other configs
    del cache, L, ds, noise, hin, hout
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3]: Nside=2048 IQU MAP_joint, one batch item per GPU, scalar all-reduces inside the timed region -------
    mj_line = None
    if "map_joint" not in skip:
        gen2 = torch.Generator(device=dev).manual_seed(4321 + rank)
        ds4, ϕ4 = build_sim_dataset(pkg, O, 2048, "IP", 1, tT, dev, gen2)
        grp = dist.group.WORLD if world > 1 else None
        kw = dict(tol=0.0, nsteps=20)                                          # fixed CG length: identical work on every rank and run
        f0, p0, _ = pkg.MAP_joint(ds4, nsteps=1, conjgrad_kwargs=kw, group=grp)            # warm-up step (allocations, caches)
        box = {}
        def mj_steps():
            box["r"] = pkg.MAP_joint(ds4, ϕstart=p0, fstart=f0, nsteps=2, conjgrad_kwargs=kw, group=grp)
        ms_mj, launches_mj = timed(mj_steps, 1, 0)
        hist = box["r"][2]
        a_, b_ = pkg.Map(box["r"][1]).arr, pkg.Map(ϕ4).arr
        q = slice(512, 1536)
        cc = float(torch.corrcoef(torch.stack([a_[0, 0, q, q].flatten(), b_[0, 0, q, q].flatten()]))[0, 1])
        mj_line = {"metric": "map_joint_seconds_per_step", "value": ms_mj / 2e3, "unit": "s/step", "higher_is_better": False,
                   "workload": f"MAP_joint, Nside=2048 IQU (BlockDiagIEB operators), 1 batch item per GPU x {world} GPUs, CG fixed at 20 iterations per step, Brent line search on the batch-summed logpdf",
                   "steps_timed": 2, "gpu_launches_per_step": launches_mj // 2, "dtype": args.dtype,
                   "collectives": ("none (1 rank)" if world == 1 else f"NCCL all-reduce of 1 scalar per line-search evaluation + 1 flag per CG iteration over {world} ranks, inside the timed region"),
                   "linesearch_evals": [h["linesearch_evals"] for h in hist], "alpha": [round(float(h["α"]), 5) for h in hist],
                   "corr_phi_map_vs_truth_rank0": cc, "items_per_sec_aggregate": world / (ms_mj / 2e3)}
        del ds4, ϕ4, f0, p0, box
        torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: Nside=512, 8 chains per GPU, HMC leap-frog steps ---------------------------------------------------
    hmc_line = None
    if "hmc" not in skip:
        gen3 = torch.Generator(device=dev).manual_seed(8765 + rank)
        NCH, NL = 8, 5
        ds5, ϕ5 = build_sim_dataset(pkg, O, 512, "P", NCH, tT, dev, gen3)
        f5, _ = pkg.argmaxf_logpdf(ds5, ϕ5, conjgrad_kwargs=dict(tol=0.0, nsteps=10))
        fm5, pm5 = pkg.mix(ds5, f5, ϕ5)
        box = {}
        def hmc_steps():
            box["r"] = pkg.gibbs_sample_ϕ(ds5, fm5, pm5, symp_kwargs=(dict(N=NL, ϵ=0.01),), always_accept=False)
        ms_h, launches_h = timed(hmc_steps, 2, 1)
        dH = box["r"][1]
        hmc_line = {"metric": "hmc_leapfrog_chain_steps_per_sec", "value": world * NCH * NL * 1e3 / ms_h, "unit": "chain-steps/s", "higher_is_better": True,
                    "workload": f"HMC update of phi° (sample_joint's gibbs_sample_phi), Nside=512 QU, {NCH} chains per GPU x {world} GPUs = {NCH * world} chains, {NL} leap-frog steps per update (each: gradient of logpdf(Mixed) = 2 flows + 2 delta-flows)",
                    "ms_per_leapfrog_step": ms_h / NL, "gpu_launches_per_leapfrog_step": launches_h // (2 * NL), "dtype": args.dtype,
                    "collectives": "none: chains are independent (pmap over chains, src/sampling.jl:292-307)", "abs_dH_max_rank0": float(np.max(np.abs(dH)))}
        del ds5, ϕ5, f5, fm5, pm5, box
        torch.cuda.empty_cache()

    if rank == 0:
        nbytes = fmap.arr.numel() * fmap.arr.element_size()
        line = {
            "metric": "lenseflow_batched_applies_per_sec", "value": world * 1e3 / ms_step, "unit": "applies/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "l2": "working set (4 state buffers 4x%.0f MB + p-cache %.1f GB) exceeds the 126 MB L2" % (nbytes / 1e6, 15 * NB * 2 * AB["pass_bytes"] / 1e9),
                       "map_applies_per_sec": world * NB * 1e3 / ms_step},
            "clocks": clocks,
            "e2e": {"value": world * 1e3 / ms_e2e_stream, "unit": "applies/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e_stream,
                    "api": "cmbl_lenseflow_apply_host_async + cmbl_lenseflow_host_sync (pinned host buffers; every step copies its own input H2D and its own result D2H; "
                           "the copies of step i overlap the integrations of steps i-1 / i+1 through two staging slots and two copy streams)",
                    "timer": "host wall clock around the K queued calls and the final cmbl_lenseflow_host_sync, bracketed by device synchronisation; max over ranks "
                             "(the D2H tail runs on the library's copy stream, which CUDA events on the caller's stream do not see)",
                    "synchronous": {"value": world * 1e3 / ms_e2e, "unit": "applies/s", "ms_per_step": ms_e2e,
                                    "api": "cmbl_lenseflow_apply_host (returns when out_host is valid: H2D + apply + D2H of ONE call, item groups pipelined inside the call)"},
                    "copies_alone_ms_per_step": ms_copy, "device_apply_ms_per_step": ms_step,
                    "host_memory_policy": (f"MPOL_INTERLEAVE over {numa} NUMA nodes (ranks >= 4)" if numa else "default (first touch)")},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_apply": {"bound": "hbm", "achieved": apply_gbs, "peak": peak, "unit": "GB/s", "frac": apply_gbs / peak, "frac_of_8TBs_nominal": apply_gbs / 8000.0,
                               "algorithmic_bytes_per_apply": AB["apply"]},
            "cpu_baseline": cpu,
            "other_precision": other,
            "fft": fft_line,
            "cg": cg_line,
            "map_joint": mj_line,
            "hmc": hmc_line,
        }
        print_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

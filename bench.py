#!/usr/bin/env python
"""bench.py — headline benchmark of the flat-sky hot path (BASELINE.json: "LenseFlow applies/sec and CG-Wiener iters/sec at
Nside=1024 batch=8; HBM GB/s vs roofline").

Workload (config.workload): Nside=1024, QU polarisation, batch=8 with 8 distinct ϕ (BASELINE configs[2]); RK4 with 7 steps.
A "step" is ONE batched LenseFlow apply  Lϕ*f  over the whole batch-8 QU field (16 planes, 28 RK stages).  The same line
also carries the CG-Wiener iteration rate (`cg`), measured in the same run on the same workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

N>1: launched by torchrun, one rank per GPU; every rank applies its own independent batch-8 field (weak scaling — the path
shards over independent batch items with no data-path collective); time = max over ranks.
`--impl reference`: the reference algorithm on the host cores (the NumPy/pocketfft oracle port — Julia is not installed, see
DESIGN.md), same metric and config, each step a bounded sample (1 of the 8 batch items) scaled to the batch.
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


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank):
    """The reference's CPU algorithm (oracle port) on all host threads; bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import cmbl_oracle as O
    cores = os.cpu_count() or 1
    O.set_workers(cores)
    npT = np.float64 if args.dtype == "f64" else np.float32
    proj = O.ProjLambert(NSIDE, NSIDE, THETA, npT)
    rng = np.random.default_rng(0)
    phi = (rng.standard_normal((1, 1, NSIDE, NSIDE)) * 1e-6).astype(npT)
    L = O.precompute(proj, phi, NSTEPS_RK)
    f = rng.standard_normal((1, NPOL, NSIDE, NSIDE)).astype(npT)
    for _ in range(min(args.warmup, 1)):
        O.lenseflow_apply(L, O.OP_L, f)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.lenseflow_apply(L, O.OP_L, f)
    dt = (time.perf_counter() - t0) / args.steps          # one batch item
    ms = dt * NB * 1e3
    val = 1e3 / ms
    sample = f"1 of {NB} batch items per step (QU pair, Nside={NSIDE}, n={NSTEPS_RK}), time scaled x{NB}; scipy.fft/pocketfft workers={cores}"
    print_line({
        "impl": "reference", "metric": "lenseflow_batched_applies_per_sec", "value": val, "unit": "applies/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"LenseFlow apply Lphi*f, Nside={NSIDE} QU batch={NB} (Cphi={NB}), RK4 n={NSTEPS_RK}, theta_pix={THETA}'"},
        "cpu_baseline": {"value": val, "unit": "applies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cg-iters", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
        return reference_arm(args, rank)

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
    hin = torch.empty(fmap.arr.shape, dtype=tT).pin_memory(); hin.copy_(fmap.arr)
    hout = torch.empty(fmap.arr.shape, dtype=tT).pin_memory()

    def step_host():
        lib.call("cmbl_lenseflow_apply_host", cache.handle, 0, P(hin), P(hout), st)
    ms_e2e, _ = timed(step_host, max(3, args.steps // 2), 2)
    assert float((hout.to(dev) - out).abs().max()) == 0.0, "host path and device path disagree"

    # ---- CG-Wiener iterations on the same workload ------------------------------------------------------------------
    ft = L * fmap
    d = pkg.gradientf_logpdf  # noqa (keep name visible)
    data = pkg.HarmonicBasis(ft)                                           # d = M B L f + n  (mask applied inside M)
    data = pkg.Field("EBFourier", data.arr + noise.arr, proj)
    ds = pkg.BaseDataSet(data, D(Cf_np), D(Cn_np), D(B_np), D(Mf_np), D(mask_np, "QUMap"), L=L, nsteps=NSTEPS_RK)
    h, *_ = ds._solver(ϕ)
    res = (ctypes.c_double * NB)()
    lib.call("cmbl_cg_begin", h, ctypes.c_void_p(0), 0, res, st)
    res0 = list(res)
    ms_cg, launches_cg = timed(lambda: lib.call("cmbl_cg_step", h, res, st), args.cg_iters, 2)
    res1 = list(res)

    # ---- per-kernel durations (CUDA events around each launch, on the launching stream) for the roofline ------------
    barrier()
    lib.cdll.cmbl_profile_begin()
    step_device(); step_device()
    prof = {}
    for line in lib.cdll.cmbl_profile_end().decode().strip().splitlines():
        nm, cnt, tot = line.split()
        prof[nm] = (int(cnt), float(tot))
    peak, peak_src = measured_peak()
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
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(args.dtype, {}).get(dom)
        except Exception:
            pass
    apply_gbs = AB["apply"] / ms_step / 1e6

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample --------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
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
    if rank == 0 and world == 1:
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

    if rank == 0:
        nbytes = fmap.arr.numel() * fmap.arr.element_size()
        line = {
            "metric": "lenseflow_batched_applies_per_sec", "value": world * 1e3 / ms_step, "unit": "applies/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"LenseFlow apply Lphi*f, Nside={NSIDE} QU batch={NB} (Cphi={NB}) per GPU, RK4 n={NSTEPS_RK}, theta_pix={THETA}'",
                       "l2": "working set (4 state buffers 4x%.0f MB + p-cache %.1f GB) exceeds the 126 MB L2" % (nbytes / 1e6, 15 * NB * 2 * AB["pass_bytes"] / 1e9),
                       "map_applies_per_sec": world * NB * 1e3 / ms_step},
            "clocks": clocks,
            "e2e": {"value": world * 1e3 / ms_e2e, "unit": "applies/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e,
                    "api": "cmbl_lenseflow_apply_host (pinned host buffers, H2D + apply + D2H)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_apply": {"bound": "hbm", "achieved": apply_gbs, "peak": peak, "unit": "GB/s", "frac": apply_gbs / peak, "frac_of_8TBs_nominal": apply_gbs / 8000.0,
                               "algorithmic_bytes_per_apply": AB["apply"]},
            "cpu_baseline": cpu,
            "other_precision": other,
            "cg": {"metric": "cg_wiener_iters_per_sec", "value": world * 1e3 / ms_cg, "unit": "iters/s", "ms_per_iter": ms_cg, "iters_timed": args.cg_iters,
                   "gpu_launches": launches_cg, "algorithmic_GBs": AB["cg_iter"] / ms_cg / 1e6, "frac": AB["cg_iter"] / ms_cg / 1e6 / peak,
                   "res_first": res0[0], "res_last": res1[0]},
        }
        print_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""World-size-2 gloo test of the N>1 path: the batch is sharded over ranks (no data-path collective), CG keeps the
reference's lock-step stopping / bestx rules through one flag all-reduce per iteration.  Runs on CPU through the host
emulator build; the same host logic drives NCCL ranks on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CMBL_EMU_THREADS="2")
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as g
    from common import make_problem
    pkg = g.load_package()
    emu = pkg._lib.Library(emu_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the global problem has batch 2·world; rank r owns items [2r, 2r+2)
        pr = make_problem(pkg, 32, 32, "P", "f64", nb=2 * world, nsteps=3, mask=True, seed=21, theta=3.0, lib=emu)
        sl = slice(2 * rank, 2 * rank + 2)
        F = pr["F"]
        d_loc = F(pr["sim"]["d"][sl], "EBFourier"); phi_loc = F(pr["sim"]["phi"][sl], "Fourier")
        ds = pr["ds"]
        ds_loc = pkg.BaseDataSet(d_loc, ds.Cf, ds.Cn, ds.B, ds.Mf, ds.Mpix, nsteps=3)
        if q.get("map_joint"):
            ds_loc.Cϕ, ds_loc.Nϕ = ds.Cϕ, ds.Nϕ
            f, ϕ, hist = pkg.MAP_joint(ds_loc, nsteps=2, conjgrad_kwargs=dict(tol=1e-1, nsteps=100), group=dist.group.WORLD)
            q["out"].put((rank, ϕ.cpu_numpy(), [(h["α"], h["cg_iters"]) for h in hist]))
            return
        x, hist = pkg.argmaxf_logpdf(ds_loc, phi_loc, conjgrad_kwargs=dict(tol=q["tol"], nsteps=40), group=dist.group.WORLD)
        q["out"].put((rank, x.cpu_numpy(), [(i, r.copy()) for i, r in hist]))
    finally:
        dist.destroy_process_group()


def test_cg_sharded_over_two_ranks_matches_single_batch(pkg, emu):
    import cmbl_oracle as O
    from common import make_problem, relerr
    world = 2
    pr = make_problem(pkg, 32, 32, "P", "f64", nb=2 * world, nsteps=3, mask=True, seed=21, theta=3.0, lib=emu)
    _, h0 = O.argmaxf_logpdf(pr["dso"], nsteps=40, tol=0.0)
    tol = float(np.max(h0[14][1])) * 1.0001
    xo, histo = O.argmaxf_logpdf(pr["dso"], nsteps=40, tol=tol)         # reference semantics on the whole batch
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, emu.path, {"tol": tol, "out": out})) for r in range(world)]
    [p.start() for p in procs]
    res = dict()
    for _ in range(world):
        r, x, hist = out.get(timeout=300)
        res[r] = (x, hist)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # every rank stops at the SAME iteration as the unsharded lock-step solve, and returns its slice of the same bestx
    for r in range(world):
        x, hist = res[r]
        assert len(hist) == len(histo)
        assert relerr(x, xo[2 * r: 2 * r + 2]) < 1e-9
        for (i, rr), (io, ro) in zip(hist, histo):
            assert i == io and np.allclose(rr, ro[2 * r: 2 * r + 2], rtol=1e-9)


def test_map_joint_sharded_over_two_ranks_matches_single_batch(pkg, emu):
    """MAP_joint with the batch sharded over 2 ranks: the line search sums logpdf over the batch (src/maximization.jl:173), so the
    one all-reduced scalar per Brent evaluation makes both ranks take the α of the unsharded run; CG stays in lock step."""
    import cmbl_oracle as O
    from common import make_problem, relerr
    world = 2
    pr = make_problem(pkg, 32, 32, "P", "f64", nb=2 * world, nsteps=3, mask=True, seed=21, theta=3.0, lib=emu)
    _, ϕo, histo = O.MAP_joint(pr["dso"], nsteps=2, conjgrad_kwargs=dict(tol=1e-1, nsteps=100))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, emu.path, {"map_joint": True, "out": out})) for r in range(world)]
    [p.start() for p in procs]
    res = dict()
    for _ in range(world):
        r, ϕ, hist = out.get(timeout=600)
        res[r] = (ϕ, hist)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for r in range(world):
        ϕ, hist = res[r]
        for (α, it), ho in zip(hist, histo):
            assert it == ho["cg_iters"] and abs(α - ho["alpha"]) < 1e-6
        assert relerr(ϕ, ϕo[2 * r: 2 * r + 2]) < 1e-6

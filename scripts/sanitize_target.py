"""Small workload for compute-sanitizer: every kernel family once at the smallest fast-path size (256², QU, 2 items), a generic-path size, and
transform length 2048 as column and as row length (64 KB tiles, 256 threads, half-bundle radix-16 sweep) in both precisions.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_target.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from common import make_problem
for (Ny, Nx, pol, dt, nb) in ((256, 256, "P", "f64", 2), (64, 64, "IP", "f64", 2), (2048, 256, "I", "f64", 1), (256, 2048, "P", "f32", 1)):
    N = Ny
    pr = make_problem(pkg, Ny, Nx, pol, dt, nb=nb, nsteps=2, mask=True, seed=3, theta=2.0, device="cuda:0")
    L = pkg.LenseFlow(pr["phi"], 2)
    fm = pkg.LenseBasis(pr["f"])
    a = L * fm; b = L.ldiv(a); c = L.H * pkg.DerivBasis(fm); d = L.H.ldiv(c)
    cache = L.cache(fm, with_minv=True)
    out = cache.apply(pkg.OP_L, fm)
    gf, gp = cache.pullback(pkg.OP_L, out, pkg.DerivBasis(fm))
    x, hist = pkg.argmaxf_logpdf(pr["ds"], pr["phi"], conjgrad_kwargs=dict(tol=0.0, nsteps=3))
    am = pkg.get_max_lensing_step(pr["phi"], pr["phi"] * 2.0)
    torch.cuda.synchronize()
    print(f"{Ny}x{Nx}", pol, dt, "ok", float(a.arr.abs().mean()), float(gp.arr.abs().mean()), hist[-1][1], am)

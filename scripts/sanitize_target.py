"""Small workload for compute-sanitizer: every kernel family once at the smallest fast-path size (256², QU, 2 items) plus a generic-path size.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_target.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from common import make_problem
for (N, pol) in ((256, "P"), (64, "IP")):
    pr = make_problem(pkg, N, N, pol, "f64", nb=2, nsteps=2, mask=True, seed=3, theta=2.0, device="cuda:0")
    L = pkg.LenseFlow(pr["phi"], 2)
    fm = pkg.LenseBasis(pr["f"])
    a = L * fm; b = L.ldiv(a); c = L.H * pkg.DerivBasis(fm); d = L.H.ldiv(c)
    cache = L.cache(fm, with_minv=True)
    out = cache.apply(pkg.OP_L, fm)
    gf, gp = cache.pullback(pkg.OP_L, out, pkg.DerivBasis(fm))
    x, hist = pkg.argmaxf_logpdf(pr["ds"], pr["phi"], conjgrad_kwargs=dict(tol=0.0, nsteps=3))
    am = pkg.get_max_lensing_step(pr["phi"], pr["phi"] * 2.0)
    torch.cuda.synchronize()
    print(N, pol, "ok", float(a.arr.abs().mean()), float(gp.arr.abs().mean()), hist[-1][1], am)

import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from common import make_problem
import cmbl_oracle as O
for dtype in ("f64", "f32"):
    for (Ny, Nx, pol, nb) in ((256, 256, "I", 1), (256, 256, "P", 1), (512, 256, "I", 1)):
        pr = make_problem(pkg, Ny, Nx, pol, dtype, nb=nb, nbphi=nb, nsteps=7, mask=False, seed=3, device="cuda:0")
        L = pkg.LenseFlow(pr["phi"], 7)
        fm = pkg.LenseBasis(pr["f"])
        for rep in range(3):
            a = L * fm; b = L.ldiv(a); c = L.H * pkg.DerivBasis(fm)
            torch.cuda.synchronize()
        print(dtype, Ny, Nx, pol, "ok", float(a.arr.abs().mean()))

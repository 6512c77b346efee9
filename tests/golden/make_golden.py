"""Generates tests/golden/lenseflow_golden.npz with the CPU oracle (run in the build container:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors for this path and Julia is not
installed here, so these freeze the oracle's outputs — a regression anchor for the CUDA path, not an independent pin."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import cmbl_oracle as O

Ny, Nx, theta, n = 32, 16, 3.0, 7
sim = O.make_dataset(Ny, Nx, theta, pol="P", T=np.float64, nb=2, seed=11, nsteps=n, mask=False)
proj, L = sim["proj"], sim["ds"].L
f = O.to_lense_basis("P", proj, sim["f"])
np.savez_compressed(os.path.join(HERE, "lenseflow_golden.npz"), Ny=Ny, Nx=Nx, theta=theta, nsteps=n, phi=sim["phi"], f_qumap=f,
                    L_f=O.lenseflow_apply(L, O.OP_L, f), LH_f=O.lenseflow_apply(L, O.OP_LH, O.rfft2(f)),
                    Linv_f=O.lenseflow_apply(L, O.OP_LINV, f))
print("wrote lenseflow_golden.npz")

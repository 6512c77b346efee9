"""Shared helpers: build the same seeded problem for the oracle and for the package (emulator on CPU / CUDA on GPU)."""
import numpy as np
import torch

import cmbl_oracle as O


def T_of(dtype):
    return (np.float64, torch.float64) if dtype == "f64" else (np.float32, torch.float32)


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.sqrt((np.abs(a - b) ** 2).sum() / max((np.abs(b) ** 2).sum(), 1e-300)))


def make_problem(pkg, Ny, Nx, pol, dtype, nb, nbphi=None, nsteps=7, mask=True, seed=0, theta=2.0, device="cpu", lib=None):
    npT, tT = T_of(dtype)
    sim = O.make_dataset(Ny, Nx, theta, pol=pol, T=npT, nb=nb, seed=seed, nsteps=nsteps, mask=mask)
    nbphi = nb if nbphi is None else nbphi
    proj = pkg.ProjLambert(Ny, Nx, theta, tT, device, lib)
    phi_np = sim["phi"][:nbphi]
    Lo = O.precompute(sim["proj"], phi_np, nsteps, phi_is_fourier=True)
    sim["ds"].L = Lo
    if nbphi != nb:       # data must be consistent with the ϕ actually used
        ft = O.lenseflow_apply(Lo, O.OP_L, O.to_lense_basis(pol, sim["proj"], sim["f"]))
        sim["d"] = sim["ds"].d = (O.apply_M(sim["ds"], O.op_mul(pol, sim["ds"].B, O.to_harmonic_basis(pol, sim["proj"], ft)))).astype(sim["proj"].cT)
    harm = {"I": "Fourier", "P": "EBFourier", "IP": "IEBFourier"}[pol]
    lense = {"I": "Map", "P": "QUMap", "IP": "IQUMap"}[pol]
    F = lambda a, basis: pkg.Field(basis, torch.from_numpy(np.ascontiguousarray(a)), proj)
    def D(a, basis=harm):
        if basis == "IEBFourier":            # BlockDiagIEB from its four half-planes [TT, TE, EE, BB]
            return pkg.BlockDiagIEB(*(a[0, i] for i in range(4)), proj=proj)
        return pkg.DiagOp(F(a, basis))
    dso = sim["ds"]
    ds = pkg.BaseDataSet(F(sim["d"], harm), D(dso.Cf), D(dso.Cn), D(dso.B), D(dso.Mf),
                         D(dso.Mpix, lense) if dso.Mpix is not None else None, nsteps=nsteps,
                         Cϕ=pkg.DiagOp(F(dso.Cphi, "Fourier")), Nϕ=pkg.DiagOp(F(dso.Nphi, "Fourier")),
                         Cf̃=D(dso.Cftilde) if dso.Cftilde is not None else None)
    return dict(sim=sim, proj=proj, oproj=sim["proj"], phi=F(phi_np, "Fourier"), f=F(sim["f"], harm), ds=ds, dso=dso, Lo=Lo,
                harm=harm, lense=lense, F=F)

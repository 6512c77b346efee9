"""Consumes OUTPUTS OF THE REFERENCE ITSELF (julia/make_fixtures.jl → tests/golden/ref/*.npy) when they are present: the oracle
(-m "not gpu") and the sm_100a library (-m gpu) against CMBLensing.jl's own numbers on identical inputs.  The build image has no Julia, so
the directory is empty there and these tests skip with the command that creates the files; wherever the fixtures exist this is the pin of
the oracle (oracle/cmbl_oracle.py header: "parity unpinned")."""
import glob
import os

import numpy as np
import pytest
import torch

import cmbl_oracle as O
from common import relerr

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")
HAVE = bool(glob.glob(os.path.join(REF, "lf_*_L_f_qumap.npy")))
need = pytest.mark.skipif(not HAVE, reason="no reference fixtures: run `julia --project=<CMBLensing.jl> julia/make_fixtures.jl tests/golden/ref`")
TOL = 1e-10          # fp64 FFT/ODE output, relative L2 (1e-11 against the same FFT; FFTW vs pocketfft/ours differ in rounding)


def ld(name):
    """Julia (Ny, Nx, Npol[, Nb]) column-major → (Nb, Npol, Nx, Ny) C-order."""
    a = np.load(os.path.join(REF, name + ".npy"))
    a = a.reshape(a.shape + (1,) * (4 - a.ndim)) if a.ndim < 4 else a
    return np.ascontiguousarray(a.transpose(3, 2, 1, 0))


def cases():
    return sorted(os.path.basename(p)[3:].split("_")[0] for p in glob.glob(os.path.join(REF, "lf_*_L_f_qumap.npy")))


@need
@pytest.mark.parametrize("tag", cases() or ["none"])
def test_oracle_lenseflow_vs_reference(tag):
    Ny, Nx = map(int, tag.split("x"))
    proj = O.ProjLambert(Ny, Nx, 2.0, np.float64)
    L = O.precompute(proj, ld(f"lf_{tag}_phi_fourier"), 7, phi_is_fourier=True)
    f = ld(f"lf_{tag}_f_qumap"); F = O.rfft2(f)
    assert relerr(O.lenseflow_apply(L, O.OP_L, f), ld(f"lf_{tag}_L_f_qumap")) < TOL
    assert relerr(O.lenseflow_apply(L, O.OP_LINV, f), ld(f"lf_{tag}_Linv_f_qumap")) < TOL
    assert relerr(O.lenseflow_apply(L, O.OP_LH, F), ld(f"lf_{tag}_LH_f_qufourier")) < TOL
    assert relerr(O.lenseflow_apply(L, O.OP_LHINV, F), ld(f"lf_{tag}_LHinv_f_qufourier")) < TOL
    for t, k in ((0.0, 0), (0.5, 7), (1.0, 14)):
        p = np.concatenate([L.p[k][0], L.p[k][1]], axis=1)
        assert relerr(p, ld(f"lf_{tag}_p_t{t}")) < TOL
    out = O.lenseflow_apply(L, O.OP_L, f)
    gf, gphi = O.lenseflow_grad(L, O.OP_L, out, F, bug_compat=True)
    assert relerr(gf, ld(f"lf_{tag}_grad_f_qufourier")) < 1e-9 and relerr(gphi, ld(f"lf_{tag}_grad_phi_fourier")) < 1e-9


def _cg_dataset():
    proj = O.ProjLambert(128, 128, 2.0, np.float64)
    g = lambda n: ld("cg_" + n)[:1].real.astype(np.float64)
    ds = O.DataSet(proj=proj, pol="P", Cf=g("Cf"), Cn=g("Cn"), Cnhat=g("Cnhat"), B=g("B"), Bhat=g("Bhat"), Mf=g("Mf"), Mpix=ld("cg_Mpix"),
                   d=ld("cg_d_ebfourier"), L=O.precompute(proj, ld("cg_phi_fourier"), 7, phi_is_fourier=True))
    return proj, ds


@need
def test_oracle_cg_vs_reference():
    proj, ds = _cg_dataset()
    assert relerr(O.gradientf_logpdf(ds, ld("cg_f_ebfourier"), ds.d), ld("cg_gradientf")) < 1e-9
    x, hist = O.argmaxf_logpdf(ds, nsteps=8, tol=0.0)
    ref = np.load(os.path.join(REF, "cg_res_history.npy")).ravel()
    assert len(hist) == len(ref) and np.allclose([h[1][0] for h in hist], ref, rtol=1e-8)
    assert relerr(x, ld("cg_fwf_ebfourier")) < 1e-8


@need
@pytest.mark.gpu
@pytest.mark.parametrize("tag", cases() or ["none"])
def test_cuda_lenseflow_vs_reference(cuda_pkg, tag):
    pkg = cuda_pkg
    Ny, Nx = map(int, tag.split("x"))
    proj = pkg.ProjLambert(Ny, Nx, 2.0, torch.float64, "cuda:0")
    F = lambda a, b: pkg.Field(b, torch.from_numpy(a), proj)
    L = pkg.LenseFlow(F(ld(f"lf_{tag}_phi_fourier"), "Fourier"), 7)
    f = F(ld(f"lf_{tag}_f_qumap"), "QUMap")
    assert relerr((L * f).cpu_numpy(), ld(f"lf_{tag}_L_f_qumap")) < TOL
    assert relerr(L.ldiv(f).cpu_numpy(), ld(f"lf_{tag}_Linv_f_qumap")) < TOL
    assert relerr((L.H * pkg.QUFourier(f)).cpu_numpy(), ld(f"lf_{tag}_LH_f_qufourier")) < TOL
    assert relerr(L.H.ldiv(pkg.QUFourier(f)).cpu_numpy(), ld(f"lf_{tag}_LHinv_f_qufourier")) < TOL
    cache = L.cache(f, with_minv=True)
    out = cache.apply(pkg.OP_L, f)
    gf, gphi = cache.pullback(pkg.OP_L, out, pkg.QUFourier(f), bug_compat=True)
    assert relerr(gf.cpu_numpy(), ld(f"lf_{tag}_grad_f_qufourier")) < 1e-9 and relerr(gphi.cpu_numpy(), ld(f"lf_{tag}_grad_phi_fourier")) < 1e-9


@need
@pytest.mark.gpu
def test_cuda_cg_vs_reference(cuda_pkg):
    pkg = cuda_pkg
    _, dso = _cg_dataset()
    proj = pkg.ProjLambert(128, 128, 2.0, torch.float64, "cuda:0")
    F = lambda a, b: pkg.Field(b, torch.from_numpy(np.ascontiguousarray(a)), proj)
    D = lambda a, b="EBFourier": pkg.DiagOp(F(a, b))
    ds = pkg.BaseDataSet(F(dso.d, "EBFourier"), D(dso.Cf), D(dso.Cn), D(dso.B), D(dso.Mf), D(dso.Mpix, "QUMap"), D(dso.Cnhat), D(dso.Bhat), nsteps=7)
    ϕ = F(ld("cg_phi_fourier"), "Fourier")
    assert relerr(pkg.gradientf_logpdf(ds, F(ld("cg_f_ebfourier"), "EBFourier"), ϕ).cpu_numpy(), ld("cg_gradientf")) < 1e-9
    x, hist = pkg.argmaxf_logpdf(ds, ϕ, conjgrad_kwargs=dict(tol=0.0, nsteps=8))
    ref = np.load(os.path.join(REF, "cg_res_history.npy")).ravel()
    assert np.allclose([h[1][0] for h in hist], ref, rtol=1e-8) and relerr(x.cpu_numpy(), ld("cg_fwf_ebfourier")) < 1e-8

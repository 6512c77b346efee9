"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package
(`cmblensing.jl_b200/`); only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` leg may import it, and there only as the checker /
the CPU baseline.

A NumPy/pocketfft restatement of the flat-sky hot path of marius311/CMBLensing.jl
@ 8e75a7c (v0.10.1).  Every function cites the reference file:line it follows
(paths relative to /root/reference).  The reference is pure Julia and Julia is not
installed in this image, so the reference itself cannot be executed here.

PARITY STATUS: **parity unpinned** beyond (a) the known-answer cases the reference's own
tests hold for this path (test/runtests.jl:252-256,269-273 logdet/tr; 116-131 basis
round trips; 259-283 logdet/tr vs dense fft) and (b) its property tests (adjoint identity
:556,:570; finite-difference gradient :559,:573) — the reference ships no golden output
vectors for LenseFlow / CG, and its FFT lives in un-vendored FFTW.jl 1.7.1 /
FFTW_jll 3.3.10 / MKL_jll 2023.2 (docs/Manifest.toml).  The DFT definition, normalisation and
half-plane layout are pinned by those tests; bit patterns are not.

Array layout: Julia `arr[iy, ix, ipol, ibatch]` column-major == NumPy C-order
`a[ibatch, ipol, ix, iy]` (iy fastest).  Map arrays: (Nb, Npol, Nx, Ny); Fourier arrays:
(Nb, Npol, Nx, Ny//2+1) complex (the halved dimension is y, `plan_rfft(A,(1,2))`,
src/util_fft.jl:32-35).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field as _dc_field

import numpy as np
import scipy.fft as sfft

_WORKERS = int(os.environ.get("CMBL_ORACLE_WORKERS", "1"))


_FFT_BACKEND = "pocketfft"


def set_workers(n: int):
    global _WORKERS
    _WORKERS = max(1, int(n))
    if _FFT_BACKEND == "mkl":
        import torch
        torch.set_num_threads(_WORKERS)


def set_fft_backend(name: str):
    """FFT provider of the restatement: "pocketfft" (scipy.fft, the default and the one every parity test uses) or "mkl"
    (torch.fft on the CPU = Intel MKL, the provider the reference recommends, README.md:56).  Same transform definition either way;
    bench.py's reference arm times both and runs the faster one."""
    global _FFT_BACKEND
    if name not in ("pocketfft", "mkl"):
        raise ValueError("fft backend must be 'pocketfft' or 'mkl'")
    _FFT_BACKEND = name
    if name == "mkl":
        import torch
        torch.set_num_threads(_WORKERS)


# ----------------------------------------------------------------------------------------------
# ProjLambert  (src/proj_lambert.jl:24-75)
# ----------------------------------------------------------------------------------------------
class ProjLambert:
    """ℓ-grids and metadata of a flat-sky Lambert projection, computed in type T exactly in the
    order of src/proj_lambert.jl:58-71 (Nyquist frequencies carry NEGATIVE ℓ, :63-64)."""

    def __init__(self, Ny: int, Nx: int, theta_pix: float = 1.0, T=np.float64):
        T = np.dtype(T).type
        self.Ny, self.Nx, self.theta_pix, self.T = int(Ny), int(Nx), float(theta_pix), T
        self.cT = np.complex64 if T is np.float32 else np.complex128
        dx = T(np.deg2rad(self.theta_pix / 60.0))                    # :58
        self.dx = dx
        self.dlx = T(2 * np.pi / np.float64(T(Nx) * dx))             # :59  (2π is Float64 in Julia)
        self.dly = T(2 * np.pi / np.float64(T(Ny) * dx))             # :60
        self.nyquist = T(2 * np.pi / np.float64(T(2) * dx))          # :61
        self.omega_pix = T(dx * dx)                                  # :62
        ky = np.fft.ifftshift(np.arange(-(Ny // 2), (Ny - 1) // 2 + 1))
        kx = np.fft.ifftshift(np.arange(-(Nx // 2), (Nx - 1) // 2 + 1))
        self.ly = (ky.astype(T) * self.dly)[: Ny // 2 + 1].astype(T)  # :63  last entry = -(Ny/2)Δℓy
        self.lx = (kx.astype(T) * self.dlx).astype(T)                 # :64
        LX, LY = self.lx[:, None], self.ly[None, :]                   # [ix, iy]
        self.lmag = np.sqrt(LX * LX + LY * LY).astype(T)              # :65
        phi = np.arctan2(LY + 0 * LX, LX + 0 * LY).astype(T)          # :66 angle(ℓx' + im ℓy)
        self.sin2phi = np.sin(T(2) * phi).astype(T)                   # :67
        self.cos2phi = np.cos(T(2) * phi).astype(T)
        self.lam_rfft = rfft_degeneracy_fac(Ny).astype(T)             # :68
        if Ny % 2 == 0:                                               # :69-71 (1-based end:-1:Nx÷2+2 ← 2:Nx÷2)
            self.sin2phi[Nx - 1: Nx // 2: -1, -1] = self.sin2phi[1: Nx // 2, -1]

    @property
    def map_shape(self):
        return (self.Nx, self.Ny)

    @property
    def fourier_shape(self):
        return (self.Nx, self.Ny // 2 + 1)


def rfft_degeneracy_fac(n: int) -> np.ndarray:
    """src/util_fft.jl:137-143"""
    if n % 2 == 0:
        return np.array([1.0] + [2.0] * (n // 2 - 1) + [1.0])
    return np.array([1.0] + [2.0] * (n // 2))


# ----------------------------------------------------------------------------------------------
# FFT conventions (src/util_fft.jl:26-27,44 ; call sites src/proj_lambert.jl:245-300)
# ----------------------------------------------------------------------------------------------
def rfft2(a: np.ndarray) -> np.ndarray:
    """m_rfft!(dst, arr, (1,2)): unnormalised batched 2-D R2C, halved dim = y (last NumPy axis)."""
    if _FFT_BACKEND == "mkl":
        import torch
        return torch.fft.rfftn(torch.from_numpy(np.ascontiguousarray(a)), dim=(-2, -1)).numpy()
    return sfft.rfftn(a, axes=(-2, -1), workers=_WORKERS)


def irfft2(F: np.ndarray, Ny: int) -> np.ndarray:
    """m_irfft!(dst, arr, (1,2)) = ldiv!(dst, plan, arr): normalised 1/(Ny·Nx); complex IFFT along x
    then c2r along y, which ignores Im of the ky=0 and ky=Ny/2 rows (FFTW/MKL/cuFFT/pocketfft)."""
    Nx = F.shape[-2]
    if _FFT_BACKEND == "mkl":
        import torch
        return torch.fft.irfftn(torch.from_numpy(np.ascontiguousarray(F)), s=(Nx, Ny), dim=(-2, -1)).numpy()
    return sfft.irfftn(F, s=(Nx, Ny), axes=(-2, -1), workers=_WORKERS)


# ----------------------------------------------------------------------------------------------
# Derivative diagonals (src/specialops.jl:144-177 ; src/proj_lambert.jl:146-159)
# ----------------------------------------------------------------------------------------------
def grad_diag(proj: ProjLambert, coord: int, prefactor: int = 1) -> np.ndarray:
    """∇diag(coord, ·, prefactor) materialised on the half-plane: (prefactor·im)·ℓx' or ·ℓy."""
    if coord == 1:
        return (prefactor * 1j * proj.lx[:, None]).astype(proj.cT)
    return (prefactor * 1j * proj.ly[None, :]).astype(proj.cT)


def gradhess(proj: ProjLambert, phi_four: np.ndarray):
    """src/specialops.jl:184-188: g = ∇ⁱ f ; H = [∇₁g₁ ∇₂g₁; ∇₁g₂ ∇₂g₂], all in Fourier basis."""
    d1, d2 = grad_diag(proj, 1), grad_diag(proj, 2)
    g = (d1 * phi_four, d2 * phi_four)
    H = ((d1 * g[0], d2 * g[0]), (d1 * g[1], d2 * g[1]))
    return g, H


# ----------------------------------------------------------------------------------------------
# Basis rotations (src/proj_lambert.jl:253-271)
# ----------------------------------------------------------------------------------------------
def eb_to_qu(proj: ProjLambert, F: np.ndarray, pol0: int = 0) -> np.ndarray:
    """QUFourier(f::LambertEBFourier) :253-258 on components (pol0, pol0+1)."""
    out = F.copy()
    E, B = F[:, pol0], F[:, pol0 + 1]
    c, s = proj.cos2phi, proj.sin2phi
    out[:, pol0] = -E * c + B * s
    out[:, pol0 + 1] = -E * s - B * c
    return out


def qu_to_eb(proj: ProjLambert, F: np.ndarray, pol0: int = 0) -> np.ndarray:
    """EBFourier(f::LambertQUFourier) :265-271."""
    out = F.copy()
    Q, U = F[:, pol0], F[:, pol0 + 1]
    c, s = proj.cos2phi, proj.sin2phi
    out[:, pol0] = -Q * c - U * s
    out[:, pol0 + 1] = Q * s - U * c
    return out


# ----------------------------------------------------------------------------------------------
# DiagOp (src/specialops.jl:9-10,18) and nan2zero (src/util.jl:32)
# ----------------------------------------------------------------------------------------------
def nan2zero(x: np.ndarray) -> np.ndarray:
    return np.where(np.isfinite(x), x, np.zeros((), dtype=x.dtype))


def pinv_diag(d: np.ndarray) -> np.ndarray:
    """pinv(D::DiagOp) = Diagonal(pinv.(diag(D))): scalar pinv, 0 → 0."""
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(d == 0, np.zeros((), dtype=d.dtype), 1 / d)
    return r.astype(d.dtype)


def diag_mul(d: np.ndarray, f: np.ndarray) -> np.ndarray:
    return d * f


def diag_ldiv(d: np.ndarray, f: np.ndarray) -> np.ndarray:
    with np.errstate(divide="ignore", invalid="ignore"):
        return nan2zero(f / d).astype(f.dtype)


# ----------------------------------------------------------------------------------------------
# BlockDiagIEB (src/specialops.jl:61-118) with the 2×2 helpers of src/field_vectors.jl:64-84.
# Stored as 4 real half-planes [ΣTE[1,1], ΣTE[2,1], ΣTE[2,2], ΣB] on axis 1: Cℓ_to_Cov(:IP) builds the
# 2×2 block symmetric ([ΣTT ΣTE; ΣTE ΣEE], src/proj_lambert.jl:368-371) and sqrt/pinv/det read the
# off-diagonal from [2,1] twice (`a, c, b, d = A[1,1], A[2,1], A[2,1], A[2,2]`).
# ----------------------------------------------------------------------------------------------
def block_from_diag(d3: np.ndarray) -> np.ndarray:
    """DiagOp{IEBFourier} (3 planes) as a block with zero off-diagonal."""
    return np.stack([d3[:, 0], np.zeros_like(d3[:, 0]), d3[:, 1], d3[:, 2]], axis=1)


def block_mul(A: np.ndarray, F: np.ndarray) -> np.ndarray:
    """L * f::IEBFourier (:79-82): (i,e) = ΣTE·(I,E), b = ΣB·B."""
    out = np.empty(np.broadcast_shapes(A[:, :3].shape, F.shape), dtype=F.dtype)
    out[:, 0] = A[:, 0] * F[:, 0] + A[:, 1] * F[:, 1]
    out[:, 1] = A[:, 1] * F[:, 0] + A[:, 2] * F[:, 1]
    out[:, 2] = A[:, 3] * F[:, 2]
    return out


def block_pinv(A: np.ndarray) -> np.ndarray:
    """pinv(L) (:88) with pinv(::2×2) of src/field_vectors.jl:74-78: idet = pinv(ad − bc); [d·idet −b·idet; −c·idet a·idet]."""
    a, c, d = A[:, 0], A[:, 1], A[:, 2]
    idet = pinv_diag(a * d - c * c)
    return np.stack([d * idet, -(c * idet), a * idet, pinv_diag(A[:, 3])], axis=1).astype(A.dtype)


def block_sqrt(A: np.ndarray) -> np.ndarray:
    """sqrt(L) (:87) with sqrt(::2×2) of src/field_vectors.jl:62-67: s = √(ad−bc), t = pinv(√(a+(d+2s)))."""
    a, c, d = A[:, 0], A[:, 1], A[:, 2]
    s = np.sqrt(a * d - c * c)
    t = pinv_diag(np.sqrt(a + (d + 2 * s)))
    return np.stack([t * (a + s), t * c, t * (d + s), np.sqrt(A[:, 3])], axis=1).astype(A.dtype)


def block_hess_sandwich(X: np.ndarray, M: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """X'·M'·Y·M·X for BlockDiagIEBs (`*(La, Lb)` = 2×2 matrix product ⊕ product of the B parts, src/specialops.jl:101),
    evaluated left to right with full (non-symmetric) 2×2 intermediates; the result is symmetric and is returned as the
    4 planes [1,1], [2,1], [2,2], B."""
    full = lambda A: (A[:, 0], A[:, 1], A[:, 1], A[:, 2], A[:, 3])                       # [a b; c d] ⊕ e
    mm = lambda x, y: (x[0] * y[0] + x[1] * y[2], x[0] * y[1] + x[1] * y[3], x[2] * y[0] + x[3] * y[2], x[2] * y[1] + x[3] * y[3], x[4] * y[4])
    x, m, y = full(X), full(M), full(Y)
    h = mm(mm(mm(mm(x, m), y), m), x)
    return np.stack([h[0], h[2], h[3], h[4]], axis=1).astype(X.dtype)


def op_mul(pol: str, A: np.ndarray, F: np.ndarray) -> np.ndarray:
    """Harmonic-basis operator times field: DiagOp (Npol planes), or BlockDiagIEB (4 planes) for IP."""
    return block_mul(A, F).astype(F.dtype) if (pol == "IP" and A.shape[1] == 4) else A * F


def op_pinv(pol: str, A: np.ndarray) -> np.ndarray:
    return block_pinv(A) if pol == "IP" else pinv_diag(A)


def op_ldiv(pol: str, A: np.ndarray, F: np.ndarray) -> np.ndarray:
    """A \\ f: nan2zero(f ./ diag) for a DiagOp, pinv(L) * f for a BlockDiagIEB (src/specialops.jl:10,78)."""
    if pol == "IP" and A.shape[1] == 4:
        return block_mul(block_pinv(A), F).astype(F.dtype)
    return diag_ldiv(A, F)


# ----------------------------------------------------------------------------------------------
# dot / logdet / tr (src/proj_lambert.jl:318-353)
# ----------------------------------------------------------------------------------------------
def dot_map(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """:318-321  per-batch Σ a·b (BatchedReal)."""
    z = (a * b).astype(np.float64)
    return z.reshape(z.shape[0], -1).sum(axis=1)


def dot_fourier(proj: ProjLambert, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """:322-325  per-batch Σ Re(conj(a)·b)·λ_rfft / (Ny·Nx)."""
    z = np.real(np.conj(a) * b).astype(np.float64) * proj.lam_rfft.astype(np.float64)
    return z.reshape(z.shape[0], -1).sum(axis=1) / (proj.Ny * proj.Nx)


def logdet_fourier(proj: ProjLambert, d: np.ndarray) -> np.ndarray:
    """:331-336"""
    with np.errstate(divide="ignore", invalid="ignore"):
        z = nan2zero(np.log(np.abs(d)) * proj.lam_rfft)
    return np.real(z.reshape(z.shape[0], -1).sum(axis=1))


def logdet_map(d: np.ndarray) -> np.ndarray:
    """:337-342"""
    n = d.shape[0]
    flat = d.reshape(n, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.abs(flat)).sum(axis=1) + np.log(np.prod(np.sign(flat), axis=1))


def tr_fourier(proj: ProjLambert, d: np.ndarray) -> np.ndarray:
    """:346-350"""
    z = d * proj.lam_rfft
    return np.real(z.reshape(z.shape[0], -1).sum(axis=1))


def tr_map(d: np.ndarray) -> np.ndarray:
    return d.reshape(d.shape[0], -1).sum(axis=1)


# ----------------------------------------------------------------------------------------------
# Cℓ → 2-D (src/numerical_algorithms.jl:148-177, src/cls.jl:9-30, src/proj_lambert.jl:173-175,361-371)
# ----------------------------------------------------------------------------------------------
def linear_interp_nan(xdat: np.ndarray, ydat: np.ndarray, x: np.ndarray) -> np.ndarray:
    """LinearInterpolation(xdat, ydat; extrapolation_bc=NaN)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.interp(x, xdat, ydat)
    out = np.where((x < xdat[0]) | (x > xdat[-1]), np.nan, out)
    return out


def cl_to_2d(proj: ProjLambert, ell: np.ndarray, cl: np.ndarray) -> np.ndarray:
    """Cℓ_to_2D: T.(nan2zero.(Cℓ(ℓmag)))."""
    v = linear_interp_nan(np.asarray(ell, dtype=np.float64), np.asarray(cl, dtype=np.float64), proj.lmag)
    return nan2zero(v).astype(proj.T)


def cl_to_cov(proj: ProjLambert, ell, cl, units=None) -> np.ndarray:
    """Cℓ_to_Cov(:I, ...) : Cℓ_to_2D / units, default units = Ωpix (:361-364)."""
    units = proj.omega_pix if units is None else proj.T(units)
    return (cl_to_2d(proj, ell, cl) / units).astype(proj.T)


def noise_cls(ell: np.ndarray, muK_arcmin_T: float = 3.0, lknee: float = 100.0, aknee: float = 3.0, pol: bool = False):
    """noiseCℓs (src/cls.jl:288-299) with beamFWHM=0 (Bℓ≡1): white + 1/f; ×2 for EE/BB."""
    ell = np.asarray(ell, dtype=np.float64)
    n1f = 1.0 + (lknee / ell) ** aknee
    return (2.0 if pol else 1.0) * np.deg2rad(muK_arcmin_T / 60.0) ** 2 * n1f


def beam_cls(ell: np.ndarray, fwhm_arcmin: float):
    """beamCℓs (src/cls.jl:307-309)."""
    ell = np.asarray(ell, dtype=np.float64)
    return np.exp(-ell ** 2 * np.deg2rad(fwhm_arcmin / 60.0) ** 2 / (8 * np.log(2)))


def lowpass_wl(lmax: int, dl: int = 50):
    """LowPass(ℓ; Δℓ=50) = BandPassOp(0:ℓ, [ones(ℓ-Δℓ+1); cos_ramp_down(Δℓ)]) (src/specialops.jl:236-241)."""
    ramp_up = (np.cos(np.linspace(np.pi, 0, dl)) + 1) / 2
    return np.arange(0, lmax + 1), np.concatenate([np.ones(lmax - dl + 1), 1 - ramp_up])


# ----------------------------------------------------------------------------------------------
# LenseFlow (src/lenseflow.jl, src/flowops.jl, src/field_vectors.jl, src/numerical_algorithms.jl)
# ----------------------------------------------------------------------------------------------
@dataclass
class CachedLenseFlow:
    """CachedLenseFlow (src/lenseflow.jl:33-60): p[τ], M⁻¹[τ] at the 2n+1 times k/(2n)."""
    proj: ProjLambert
    nsteps: int
    p: list = _dc_field(default_factory=list)       # p[k] = (p1, p2) maps, shape (Nbϕ,1,Nx,Ny)
    Minv: list = _dc_field(default_factory=list)    # Minv[k] = ((m11,m12),(m21,m22))
    ts: list = _dc_field(default_factory=list)


def precompute(proj: ProjLambert, phi, nsteps: int = 7, phi_is_fourier: bool = False) -> CachedLenseFlow:
    """precompute! (src/lenseflow.jl:131-142); pinv! (src/field_vectors.jl:86-94, note a,c,b,d =
    H11,H21,H21,H22: the off-diagonal is read twice from [2,1]); p = M⁻¹' ∇ϕ (:46-47)."""
    T = proj.T
    phi = np.asarray(phi)
    if phi.ndim == 2:
        phi = phi[None, None]
    F = phi.astype(proj.cT) if phi_is_fourier else rfft2(phi.astype(T))
    g, H = gradhess(proj, F)
    g = [irfft2(x, proj.Ny).astype(T) for x in g]
    H11, H21, H22 = (irfft2(H[0][0], proj.Ny).astype(T), irfft2(H[1][0], proj.Ny).astype(T),
                     irfft2(H[1][1], proj.Ny).astype(T))
    L = CachedLenseFlow(proj, nsteps)
    for k in range(2 * nsteps + 1):
        t = T(k / (2 * nsteps))
        a = T(1) + t * H11
        d = T(1) + t * H22
        b = c = t * H21
        det = a * d - b * c
        idet = pinv_diag(det)
        m11, m12, m21, m22 = idet * d, -idet * b, -idet * c, idet * a
        p1 = m11 * g[0] + m21 * g[1]
        p2 = m12 * g[0] + m22 * g[1]
        L.ts.append(t)
        L.p.append((p1.astype(T), p2.astype(T)))
        L.Minv.append(((m11, m12), (m21, m22)))
    return L


def velocity(L: CachedLenseFlow, k: int, f: np.ndarray) -> np.ndarray:
    """velocity → v! (src/lenseflow.jl:150-161): state in Map basis."""
    proj = L.proj
    p1, p2 = L.p[k]
    F = rfft2(f)
    dx = irfft2(grad_diag(proj, 1) * F, proj.Ny)
    dy = irfft2(grad_diag(proj, 2) * F, proj.Ny)
    return (p1 * dx + p2 * dy).astype(f.dtype)


def velocity_H(L: CachedLenseFlow, k: int, F: np.ndarray) -> np.ndarray:
    """velocityᴴ → v! (src/lenseflow.jl:163-174): state in Fourier basis."""
    proj = L.proj
    p1, p2 = L.p[k]
    f = irfft2(F, proj.Ny)
    return (grad_diag(proj, 1) * rfft2(p1 * f) + grad_diag(proj, 2) * rfft2(p2 * f)).astype(F.dtype)


def neg_delta_velocity_H(L: CachedLenseFlow, k: int, state, bug_compat: bool = True):
    """negδvelocityᴴ → v! (src/lenseflow.jl:176-214) on (f Map, δf Fourier, δϕ Fourier).
    bug_compat=True reproduces the aliased in-place 2×2 product (:198-200 with
    src/field_vectors.jl:48-49, both operands bound to L.memŁvϕ): m₂ uses the *new* m₁."""
    proj = L.proj
    f, df, dphi = state
    t = L.ts[k]
    p1, p2 = L.p[k]
    (m11, m12), (m21, m22) = L.Minv[k]
    d1, d2 = grad_diag(proj, 1), grad_diag(proj, 2)
    Ldf = irfft2(df, proj.Ny)
    ddf_dt = d1 * rfft2(p1 * Ldf) + d2 * rfft2(p2 * Ldf)
    Ff = rfft2(f)
    gx, gy = irfft2(d1 * Ff, proj.Ny), irfft2(d2 * Ff, proj.Ny)
    df_dt = p1 * gx + p2 * gy
    w1 = (Ldf * gx).sum(axis=1, keepdims=True)          # spin_adjoint(Łδf) * Ł∇f  (src/proj_lambert.jl:423-430)
    w2 = (Ldf * gy).sum(axis=1, keepdims=True)
    m1 = m11 * w1 + m12 * w2
    m2 = (m21 * m1 + m22 * w2) if bug_compat else (m21 * w1 + m22 * w2)
    ddphi_dt = d1 * rfft2(m1) + d2 * rfft2(m2)
    nd = (-d1, -d2)                                     # ∇'  = conj → −iℓ
    for i, mi in enumerate((m1, m2)):
        for j, pj in enumerate((p1, p2)):
            ddphi_dt = ddphi_dt + nd[i] * (nd[j] * rfft2(t * pj * mi))
    return (df_dt.astype(f.dtype), ddf_dt.astype(df.dtype), ddphi_dt.astype(dphi.dtype))


def get_max_lensing_step(proj: ProjLambert, phi_four: np.ndarray, eta_four: np.ndarray) -> np.ndarray:
    """get_max_lensing_step(ϕ, η) (src/lenseflow.jl:242-256): smallest positive root α of det(𝕀 + ∇∇(ϕ + α η)) = 0 over the pixels, per batch
    item (the reference takes the minimum over the batch as well).  Both off-diagonal Hessian entries are ϕ₁₂ = H[2,1] as in the reference."""
    T = proj.T
    _, Hp = gradhess(proj, phi_four); _, He = gradhess(proj, eta_four)
    m = lambda h: irfft2(h, proj.Ny).astype(T)
    # `ϕ₁₁, ϕ₁₂, ϕ₂₁, ϕ₂₂ = Map.(H)` walks the SMatrix column-major: the variable called ϕ₁₂ is H[2,1] = ∇₁(∇₂ϕ)
    p11, p12, p22 = m(Hp[0][0]), m(Hp[1][0]), m(Hp[1][1])
    e11, e12, e22 = m(He[0][0]), m(He[1][0]), m(He[1][1])
    one = T(1)
    a = e11 * e22 - e12 * e12
    b = e11 * (one + p22) + e22 * (one + p11) - T(2) * e12 * p12
    c = (one + p11) * (one + p22) - p12 * p12
    with np.errstate(invalid="ignore", divide="ignore"):
        sq = np.sqrt(b * b - T(4) * a * c)
        a1 = (-b + sq) / (T(2) * a); a2 = (-b - sq) / (T(2) * a)
    out = []
    for i in range(a1.shape[0]):
        cand = np.concatenate([a1[i][a1[i] > 0].ravel(), a2[i][a2[i] > 0].ravel()])
        out.append(float(cand.min()) if cand.size else np.inf)
    return np.array(out)


def _axpy(y, a, k):
    if isinstance(y, tuple):
        return tuple(_axpy(yi, a, ki) for yi, ki in zip(y, k))
    return (y + y.real.dtype.type(a) * k).astype(y.dtype)


def rk4(F, y0, k0: int, k1: int, nsteps: int):
    """RK4Solver (src/numerical_algorithms.jl:11-24) on the stage grid k=0..2n (t=k/2n):
    h=(t₁−t₀)/n; stages at k, k±1, k±1, k±2."""
    sgn = 1 if k1 > k0 else -1
    h = sgn / nsteps
    y = y0
    k = k0
    for _ in range(nsteps):
        a = F(k, y)
        b = F(k + sgn, _axpy(y, h / 2, a))
        c = F(k + sgn, _axpy(y, h / 2, b))
        d = F(k + 2 * sgn, _axpy(y, h, c))
        if isinstance(y, tuple):
            y = tuple((yi + yi.real.dtype.type(h) * (ai + 2 * (bi + ci) + di) / 6).astype(yi.dtype)
                      for yi, ai, bi, ci, di in zip(y, a, b, c, d))
        else:
            y = (y + y.real.dtype.type(h) * (a + 2 * (b + c) + d) / 6).astype(y.dtype)
        k += 2 * sgn
    return y


OP_L, OP_LH, OP_LINV, OP_LHINV = 0, 1, 2, 3


def lenseflow_apply(L: CachedLenseFlow, op: int, x: np.ndarray) -> np.ndarray:
    """*, \\ on FlowOp / Adjoint (src/flowops.jl:11-14).  op 0: L*f (Map→Map, t 0→1);
    1: L'*f (Fourier→Fourier, t 1→0); 2: L\\f (Map, t 1→0); 3: L'\\f (Fourier, t 0→1)."""
    n = L.nsteps
    if op == OP_L:
        return rk4(lambda k, y: velocity(L, k, y), x, 0, 2 * n, n)
    if op == OP_LH:
        return rk4(lambda k, y: velocity_H(L, k, y), x, 2 * n, 0, n)
    if op == OP_LINV:
        return rk4(lambda k, y: velocity(L, k, y), x, 2 * n, 0, n)
    if op == OP_LHINV:
        return rk4(lambda k, y: velocity_H(L, k, y), x, 0, 2 * n, n)
    raise ValueError(op)


def lenseflow_grad(L: CachedLenseFlow, op: int, f_out: np.ndarray, delta_four: np.ndarray, bug_compat=True):
    """Pullback of L*f (op 0: δ-flow t 1→0 from (L f, Δ, 0)) or L\\f (op 2: t 0→1)
    (src/flowops.jl:40-68).  Returns (δf Fourier, δϕ Fourier)."""
    n = L.nsteps
    nbphi = L.p[0][0].shape[0]
    dphi0 = np.zeros((nbphi, 1) + L.proj.fourier_shape, dtype=L.proj.cT)
    k0, k1 = (2 * n, 0) if op == OP_L else (0, 2 * n)
    _, df, dphi = rk4(lambda k, y: neg_delta_velocity_H(L, k, y, bug_compat), (f_out, delta_four, dphi0), k0, k1, n)
    return df, dphi


# ----------------------------------------------------------------------------------------------
# DataSet / Wiener filter (src/dataset.jl:37-137, src/maximization.jl:17-42, src/numerical_algorithms.jl:73-134)
# ----------------------------------------------------------------------------------------------
@dataclass
class DataSet:
    """BaseDataSet restricted to what load_sim builds for pol ∈ {I, P} (src/dataset.jl:186-338):
    Cf, Cn, Cn̂, B, B̂ and the Fourier part of M are diagonals in the harmonic basis of the field
    (Fourier for I, EBFourier for P; arrays (1|Nb, Npol, Nx, Ny/2+1) real); the pixel part of M is
    a Map-basis diagonal (QUMap for P).  M = Mfourier * Mpix ; M̂ = Mfourier (:283-296)."""
    proj: ProjLambert
    pol: str                      # "I", "P" or "IP" (IP: every harmonic operator is a BlockDiagIEB of 4 planes)
    Cf: np.ndarray
    Cn: np.ndarray
    Cnhat: np.ndarray
    B: np.ndarray
    Bhat: np.ndarray
    Mf: np.ndarray                # Fourier part of M
    Mpix: np.ndarray | None       # pixel mask (Map basis), or None for I
    d: np.ndarray | None = None   # data, harmonic basis
    L: CachedLenseFlow | None = None
    Cphi: np.ndarray | None = None    # ϕ prior covariance (Fourier diagonal, (1,1,Nx,Ny/2+1))
    Nphi: np.ndarray | None = None    # ϕ noise estimate used by the ϕ° Hessian preconditioner (src/dataset.jl:134-137)
    D: np.ndarray | None = None       # mixing matrices of the Mixed parametrisation (None = identity)
    G: np.ndarray | None = None
    Cftilde: np.ndarray | None = None # lensed ("total") field covariance Cf̃ used by the quadratic estimate (src/dataset.jl:270)

    @property
    def npol(self):
        return {"I": 1, "P": 2, "IP": 3}[self.pol]


def to_lense_basis(ds_or_pol, proj, F):
    """Ł(f) for a harmonic-basis field: (EB→QU) then irfft2 (src/generic.jl:88-93)."""
    pol = ds_or_pol if isinstance(ds_or_pol, str) else ds_or_pol.pol
    if pol != "I":
        F = eb_to_qu(proj, F, 1 if pol == "IP" else 0)        # IQU: components 2:3 (src/proj_lambert.jl:284)
    return irfft2(F, proj.Ny)


def to_harmonic_basis(ds_or_pol, proj, f):
    """EBFourier(f::QUMap) / Fourier(f::Map)."""
    pol = ds_or_pol if isinstance(ds_or_pol, str) else ds_or_pol.pol
    F = rfft2(f)
    if pol != "I":
        F = qu_to_eb(proj, F, 1 if pol == "IP" else 0)        # :292
    return F


def apply_M(ds: DataSet, F):
    """M*f = Mfourier * (Mpix * f)."""
    if ds.Mpix is not None:
        F = to_harmonic_basis(ds, ds.proj, ds.Mpix * to_lense_basis(ds, ds.proj, F))
    return op_mul(ds.pol, ds.Mf, F)


def apply_MH(ds: DataSet, F):
    """M'*f = Mpix' * (Mfourier' * f); result left in the Map (QUMap) basis when Mpix is present."""
    F = op_mul(ds.pol, ds.Mf, F)
    if ds.Mpix is not None:
        return ds.Mpix * to_lense_basis(ds, ds.proj, F), True
    return F, False


def gradientf_logpdf(ds: DataSet, f_harm: np.ndarray, d: np.ndarray) -> np.ndarray:
    """gradientf_logpdf(::BaseDataSet) (src/dataset.jl:76-80):
       Lϕ'*(B'*(M'*(pinv(Cn)*(d − M*(B*(Lϕ*f)))))) − pinv(Cf)*f.
    Returned in the DerivBasis Fourier representation the reference ends up in
    (QUFourier for P by basis promotion, src/generic.jl:185-200); we convert back to the harmonic
    (EB) basis so the caller sees one basis throughout — a unitary change, not a numerical one."""
    proj = ds.proj
    ft = lenseflow_apply(ds.L, OP_L, to_lense_basis(ds, proj, f_harm))             # Lϕ*f   (Map)
    pol0 = 1 if ds.pol == "IP" else 0
    Bf = op_mul(ds.pol, ds.B, to_harmonic_basis(ds, proj, ft))                    # B*f̃
    res = d - apply_M(ds, Bf)
    x = op_mul(ds.pol, op_pinv(ds.pol, ds.Cn), res)
    x, is_map = apply_MH(ds, x)
    if is_map:
        x = to_harmonic_basis(ds, proj, x)
    x = op_mul(ds.pol, ds.B, x)                                                   # B' (real, symmetric)
    xq = eb_to_qu(proj, x, pol0) if ds.pol != "I" else x                          # Ð(·) → (I)QUFourier
    y = lenseflow_apply(ds.L, OP_LH, xq.astype(proj.cT))                          # Lϕ'*  (QU/Fourier)
    if ds.pol != "I":
        y = qu_to_eb(proj, y, pol0)
    return (y - op_mul(ds.pol, op_pinv(ds.pol, ds.Cf), f_harm)).astype(proj.cT)


def mix(ds: DataSet, proj: ProjLambert, pol: str, f_harm: np.ndarray, phi_four: np.ndarray, D=None, G=None, nsteps: int = 7):
    """mix(ds; f, ϕ) (src/dataset.jl:96-101): f° = L(ϕ)·D·f (returned in the Map/QUMap basis), ϕ° = G·ϕ."""
    L = precompute(proj, phi_four, nsteps, phi_is_fourier=True)
    Df = f_harm if D is None else op_mul(pol, D, f_harm)
    return lenseflow_apply(L, OP_L, to_lense_basis(pol, proj, Df)), (phi_four if G is None else diag_mul(G, phi_four))


def unmix(ds: DataSet, proj: ProjLambert, pol: str, f_mixed_map: np.ndarray, phi_mixed: np.ndarray, D=None, G=None, nsteps: int = 7):
    """unmix(ds; f°, ϕ°) (src/dataset.jl:111-116): ϕ = G \\ ϕ°, f = D \\ (L(ϕ) \\ f°) (harmonic basis)."""
    phi = phi_mixed if G is None else diag_ldiv(G, phi_mixed)
    L = precompute(proj, phi, nsteps, phi_is_fourier=True)
    f = to_harmonic_basis(pol, proj, lenseflow_apply(L, OP_LINV, f_mixed_map))
    return (f if D is None else op_ldiv(pol, D, f)), phi


def hess_preconditioner(ds: DataSet) -> np.ndarray:
    """Hessian_logpdf_preconditioner(:f) (src/dataset.jl:129-132): pinv(Cf) + B̂'M̂'pinv(Cn̂)M̂B̂."""
    if ds.pol == "IP":       # BlockDiagIEB sums / products (src/specialops.jl:101-102)
        return (block_pinv(ds.Cf) + block_hess_sandwich(ds.Bhat, ds.Mf, block_pinv(ds.Cnhat))).astype(ds.proj.T)
    return (pinv_diag(ds.Cf) + ds.Bhat * ds.Mf * pinv_diag(ds.Cnhat) * ds.Mf * ds.Bhat).astype(ds.proj.T)


def conjugate_gradient(Mdiag, A, dot, b, x0, nsteps: int, tol: float, pol: str = "I"):
    """conjugate_gradient (src/numerical_algorithms.jl:73-134), exact update order; `Mdiag` is the
    diagonal preconditioner (M \\ r = nan2zero(r ./ diag)); per-batch scalars broadcast over axis 0
    (BatchedReal, src/batching.jl:9-45).  Returns (bestx, history[(i, res)])."""
    def bc(s, like):
        return np.asarray(s, dtype=like.real.dtype).reshape((-1,) + (1,) * (like.ndim - 1))
    if pol == "IP":          # M \\ r = pinv(M) * IEBFourier(r) (src/specialops.jl:78)
        Minv = block_pinv(Mdiag)
        ldiv = lambda M_, r_: block_mul(Minv, r_).astype(r_.dtype)
    else:
        ldiv = diag_ldiv
    x = x0
    r = b - A(x)
    z = ldiv(Mdiag, r)
    p = z
    res = dot(r, z)
    assert not np.any(np.isnan(res))
    bestres, bestx = res.copy(), x
    hist = [(1, res.copy())]
    for i in range(2, nsteps + 1):
        Ap = A(p)
        alpha = res / dot(p, Ap)
        x = (x + bc(alpha, x) * p).astype(x.dtype)
        r = (r - bc(alpha, r) * Ap).astype(r.dtype)
        z = ldiv(Mdiag, r)
        res2 = dot(r, z)
        p = (z + bc(res2 / res, p) * p).astype(p.dtype)
        res = res2
        if np.all(res < bestres):
            bestres, bestx = res.copy(), x
        hist.append((i, res.copy()))
        if np.all(res < tol):
            break
    return bestx, hist


def argmaxf_logpdf(ds: DataSet, d=None, fstart=None, nsteps: int = 500, tol: float = 1e-1, offset: bool = False):
    """argmaxf_logpdf (src/maximization.jl:17-42)."""
    proj = ds.proj
    d = ds.d if d is None else d
    zero_f = np.zeros(d.shape, dtype=proj.cT)
    b = -gradientf_logpdf(ds, zero_f, d)
    a0 = gradientf_logpdf(ds, zero_f, np.zeros_like(d))
    if offset:
        b = b + a0
    A = lambda f: gradientf_logpdf(ds, f, np.zeros_like(d)) - a0
    dot = lambda u, v: dot_fourier(proj, u, v)
    return conjugate_gradient(hess_preconditioner(ds), A, dot, b, zero_f if fstart is None else fstart, nsteps, tol, pol=ds.pol)


def op_sqrt_mul(pol: str, C: np.ndarray, F: np.ndarray) -> np.ndarray:
    """simulate(rng, L) = sqrt(L) * randn (src/specialops.jl:6,94) applied to the transform F of a unit white map."""
    return (block_mul(block_sqrt(C), F) if pol == "IP" else np.sqrt(C) * F).astype(F.dtype)


def simulate_ds(ds: DataSet, white_f: np.ndarray, white_n: np.ndarray):
    """simulate(rng, ds; ϕ) (src/dataset.jl:60-67) with the unit white-noise maps given explicitly."""
    proj = ds.proj
    f = op_sqrt_mul(ds.pol, ds.Cf, to_harmonic_basis(ds, proj, white_f))
    n = op_sqrt_mul(ds.pol, ds.Cn, to_harmonic_basis(ds, proj, white_n))
    ft = lenseflow_apply(ds.L, OP_L, to_lense_basis(ds, proj, f))
    d = (apply_M(ds, op_mul(ds.pol, ds.B, to_harmonic_basis(ds, proj, ft))) + n).astype(proj.cT)
    return dict(f=f, ft=ft, d=d)


def sample_f(ds: DataSet, white_f: np.ndarray, white_n: np.ndarray, nsteps: int = 500, tol: float = 1e-1):
    """sample_f (src/maximization.jl:56-62): sim.f + argmaxf_logpdf(ds, Ω, d − sim.d; offset=true)."""
    sim = simulate_ds(ds, white_f, white_n)
    df, hist = argmaxf_logpdf(ds, d=(ds.d - sim["d"]).astype(ds.proj.cT), nsteps=nsteps, tol=tol, offset=True)
    return (sim["f"] + df).astype(ds.proj.cT), hist


# ----------------------------------------------------------------------------------------------
# Joint posterior: logpdf, its gradient in the mixed parametrisation, MAP_joint
# (src/dataset.jl:60-67,84-117,134-137, src/distributions.jl:11-15, src/maximization.jl:115-222)
# ----------------------------------------------------------------------------------------------
def op_logdet(pol: str, proj: ProjLambert, A: np.ndarray) -> np.ndarray:
    """logdet(Diagonal) (src/proj_lambert.jl:331-336); logdet(BlockDiagIEB) = logdet(det ΣTE) + logdet ΣB (src/specialops.jl:94)."""
    if pol == "IP":
        det = A[:, 0] * A[:, 2] - A[:, 1] * A[:, 1]
        return logdet_fourier(proj, det[:, None]) + logdet_fourier(proj, A[:, 3:4])
    return logdet_fourier(proj, A)


def logpdf(ds: DataSet, f_harm: np.ndarray, phi_four: np.ndarray, d=None) -> np.ndarray:
    """logpdf(ds; f, ϕ) of the BaseDataSet forward model (src/dataset.jl:60-67) with the MvNormal terms of
    src/distributions.jl:11-15: −(z'pinv(Σ)z + logdet Σ)/2 for f ~ N(0,Cf), ϕ ~ N(0,Cϕ), d ~ N(M B L(ϕ) f, Cn).  Per batch item."""
    proj, pol = ds.proj, ds.pol
    d = ds.d if d is None else d
    L = precompute(proj, phi_four, ds.L.nsteps, phi_is_fourier=True)
    ft = lenseflow_apply(L, OP_L, to_lense_basis(ds, proj, f_harm))
    z = apply_M(ds, op_mul(pol, ds.B, to_harmonic_basis(ds, proj, ft))) - d
    quad = lambda C, v, pl: dot_fourier(proj, v, op_mul(pl, op_pinv(pl, C), v)) + op_logdet(pl, proj, C)
    return -(quad(ds.Cn, z, pol) + quad(ds.Cf, f_harm, pol) + quad(ds.Cphi, phi_four, "I")) / 2


def mixing_D(ds: DataSet, sigma_len_arcmin: float = 5.0) -> np.ndarray:
    """load_sim's D = sqrt((Cf + (σ²len + 2Cn̂)) pinv(Cf)), σ²len = deg2rad(5/60)² (src/dataset.jl:325-332).  For pol = IP the
    BlockDiagIEB algebra of src/specialops.jl:99-105 (UniformScaling adds to the diagonal entries and to B; `*` is the 2×2 matrix
    product) followed by the 2×2 sqrt of src/field_vectors.jl:62-67, which reads the off-diagonal of the product from [2,1]."""
    s2 = ds.proj.T(np.deg2rad(sigma_len_arcmin / 60.0) ** 2)
    if ds.pol == "IP":
        cf, cn, pi = ds.Cf, ds.Cnhat, block_pinv(ds.Cf)
        a, c, d, e = cf[:, 0] + (s2 + 2 * cn[:, 0]), cf[:, 1] + 2 * cn[:, 1], cf[:, 2] + (s2 + 2 * cn[:, 2]), cf[:, 3] + (s2 + 2 * cn[:, 3])
        P = np.stack([a * pi[:, 0] + c * pi[:, 1], c * pi[:, 0] + d * pi[:, 1], c * pi[:, 1] + d * pi[:, 2], e * pi[:, 3]], axis=1)   # [1,1], [2,1], [2,2], B
        return block_sqrt(P).astype(ds.proj.T)
    return np.sqrt((ds.Cf + (s2 + 2 * ds.Cnhat)) * pinv_diag(ds.Cf)).astype(ds.proj.T)


def logpdf_mixed(ds: DataSet, f_mixed_map: np.ndarray, phi_mixed: np.ndarray) -> np.ndarray:
    """logpdf(Mixed(ds); f°, ϕ°) (src/dataset.jl:84-87); logdet(D,θ) = logdet(G,θ) = 0 without θ dependence (src/generic.jl:269)."""
    f, phi = unmix(ds, ds.proj, ds.pol, f_mixed_map, phi_mixed, D=ds.D, G=ds.G, nsteps=ds.L.nsteps)
    return logpdf(ds, f, phi)


def gradient_logpdf_mixed(ds: DataSet, f_mixed_map: np.ndarray, phi_mixed: np.ndarray, bug_compat: bool = True):
    """gradient(Ω° -> logpdf(Mixed(ds); f°, Ω°...)) (src/maximization.jl:151) by the reference's pullbacks: the chain
    f° → f₁ = L(ϕ)\f° → f = D\f₁ → f̃ = L(ϕ)f → r = d − M B f̃ is differentiated backwards with the δ-flows of `L*f` and
    `L\f` (src/flowops.jl:40-68).  Returns (∇f° in the Ð basis, ∇ϕ° Fourier); d lnP = ⟨∇f°, δf°⟩ + ⟨∇ϕ°, δϕ°⟩."""
    proj, pol = ds.proj, ds.pol
    pol0 = 1 if pol == "IP" else 0
    phi = phi_mixed if ds.G is None else diag_ldiv(ds.G, phi_mixed)
    L = precompute(proj, phi, ds.L.nsteps, phi_is_fourier=True)
    f1 = lenseflow_apply(L, OP_LINV, f_mixed_map)                                   # Map
    f1h = to_harmonic_basis(ds, proj, f1)
    f = f1h if ds.D is None else op_ldiv(pol, ds.D, f1h)
    ft = lenseflow_apply(L, OP_L, to_lense_basis(ds, proj, f))
    r = ds.d - apply_M(ds, op_mul(pol, ds.B, to_harmonic_basis(ds, proj, ft)))
    x, is_map = apply_MH(ds, op_mul(pol, op_pinv(pol, ds.Cn), r))
    if is_map:
        x = to_harmonic_basis(ds, proj, x)
    g_ft = op_mul(pol, ds.B, x)                                                    # ∂lnP/∂f̃ = B'M'pinv(Cn) r
    deriv = lambda h: (eb_to_qu(proj, h, pol0) if pol != "I" else h).astype(proj.cT)
    df_a, dphi_a = lenseflow_grad(L, OP_L, ft, deriv(g_ft), bug_compat)
    df_a = qu_to_eb(proj, df_a, pol0) if pol != "I" else df_a
    g_f = df_a - op_mul(pol, op_pinv(pol, ds.Cf), f)
    g_f1 = g_f if ds.D is None else op_ldiv(pol, ds.D, g_f)                         # D real and symmetric: D⁻ᵀ = D⁻¹
    df0, dphi_b = lenseflow_grad(L, OP_LINV, f1, deriv(g_f1), bug_compat)
    g_phi = dphi_a + dphi_b - pinv_diag(ds.Cphi) * phi
    if ds.G is not None:
        g_phi = diag_ldiv(ds.G, g_phi)
    return df0.astype(proj.cT), g_phi.astype(proj.cT)


def MAP_joint(ds: DataSet, nsteps: int = 5, conjgrad_kwargs=dict(tol=1e-1, nsteps=500), alpha_tol: float = 1e-4, bug_compat: bool = True):
    """MAP_joint (src/maximization.jl:115-222) for (f, ϕ): coordinate descent alternating the CG Wiener filter at fixed ϕ with
    one preconditioned gradient step in ϕ° whose length is found by Brent's bounded line search on [0, 2α] (the reference
    calls Optim.Brent; here scipy's bounded Brent).  G = 1 as in the reference (:137).  Returns (f, ϕ, history)."""
    from scipy.optimize import minimize_scalar
    proj, pol = ds.proj, ds.pol
    G_save, ds.G = ds.G, None
    phi = np.zeros((ds.d.shape[0], 1) + proj.fourier_shape, dtype=proj.cT)
    f, alpha, hist = None, 1.0, []
    H = pinv_diag(ds.Cphi) + pinv_diag(ds.Nphi)                                     # Hessian_logpdf_preconditioner((:ϕ°,)) :134-137
    try:
        for step in range(nsteps):
            ds.L = precompute(proj, phi, ds.L.nsteps, phi_is_fourier=True)
            f, cg_hist = argmaxf_logpdf(ds, fstart=f, **conjgrad_kwargs)
            f_mixed, phi_mixed = mix(ds, proj, pol, f, phi, D=ds.D, G=None, nsteps=ds.L.nsteps)
            _, g = gradient_logpdf_mixed(ds, f_mixed, phi_mixed, bug_compat)
            step_dir = diag_ldiv(H, g)
            amax = 2 * alpha
            obj = lambda a: -float(logpdf_mixed(ds, f_mixed, (phi_mixed + proj.T(a) * step_dir).astype(proj.cT)).sum())
            sol = minimize_scalar(obj, bounds=(0.0, amax), method="bounded", options=dict(xatol=alpha_tol))
            alpha = float(sol.x)
            phi_mixed = (phi_mixed + proj.T(alpha) * step_dir).astype(proj.cT)
            lp = logpdf_mixed(ds, f_mixed, phi_mixed)
            # Ω = delete(unmix(...), (:f, :θ)) (src/maximization.jl:206): only ϕ is taken from unmix; f stays the CG solution, which is
            # both the next step's fstart (prevf, :230) and the value returned (:224)
            _, phi = unmix(ds, proj, pol, f_mixed, phi_mixed, D=ds.D, G=None, nsteps=ds.L.nsteps)
            hist.append(dict(step=step + 1, logpdf=lp, alpha=alpha, cg_iters=len(cg_hist), linesearch_evals=int(sol.nfev)))
    finally:
        ds.G = G_save
    return f, phi, hist


# ----------------------------------------------------------------------------------------------
# Quadratic estimate of ϕ and its analytic N⁰ (src/quadratic_estimate.jl:30-199)
# ----------------------------------------------------------------------------------------------
class _QELegs:
    """QE_leg (src/quadratic_estimate.jl:84-93): Map(nan2zero(C · ∇[1].diag^p₁ · ∇[2].diag^p₂ / sqrt(∇².diag)^n)) for an index list
    in which a bracketed index [i] contributes the wave-vector factor iℓ_i and a plain index j the factor iℓ_j/|ℓ|; memoised on
    (C, n, p₁, p₂) like the reference (all terms are symmetric in their indices)."""
    def __init__(self, proj):
        self.proj, self.memo = proj, {}
        self.d1, self.d2 = grad_diag(proj, 1), grad_diag(proj, 2)
        self.lmag = proj.lmag.astype(proj.T)

    def __call__(self, C, *inds):
        n = sum(1 for x in inds if isinstance(x, int))
        first = [x if isinstance(x, int) else x[0] for x in inds]
        p1, p2 = first.count(1), first.count(2)
        key = (id(C), n, p1, p2)
        if key not in self.memo:
            with np.errstate(divide="ignore", invalid="ignore"):
                F = nan2zero(C * self.d1 ** p1 * self.d2 ** p2 / self.lmag ** n).astype(self.proj.cT)
            self.memo[key] = (irfft2(F, self.proj.Ny), C)          # keep C alive so that id(C) stays unique
        return self.memo[key][0]


def _eps(a, b):          # levicivita([a, b, 3])
    return 0 if a == b else (1 if (a, b) == (1, 2) else -1)


def quadratic_estimate(ds: DataSet, which: str | None = None, wiener_filtered: bool = True, weights: str = "unlensed", AL=None, d2=None,
                       abs_each_term: bool = True):
    """quadratic_estimate(ds, which; wiener_filtered, weights, AL) (src/quadratic_estimate.jl:30-199) for which ∈ {TT, EE, EB}:
    returns dict(phi_qe, AL, Nphi) with Nϕ = AL.  Only the Fourier-diagonal B̂, M̂, Cn̂ enter (TF = M̂·B̂), as in the reference.
    `abs_each_term=True` reproduces the reference's normalisation `pinv(Σ_ij abs.(∇ᵢ∇ⱼ·Fourier(A(i,j))))` (:117,151,190), which takes
    the absolute value of every (i,j) term before summing; the cross terms are not sign-definite, so for EB this under-normalises
    the estimate by ≈40 % (measured response 0.53–0.62; tests/test_oracle.py).  `False` uses |Σ_ij …|, whose response is 1."""
    from itertools import product
    proj, pol = ds.proj, ds.pol
    assert weights in ("lensed", "unlensed")
    which = which or ("TT" if pol == "I" else "EB")
    assert which in ("TT", "EE", "EB")
    if pol == "IP":          # ds.d[pol], Cf[pol], ... (:44): the I or the P part of an IQU dataset; BlockDiagIEB[:P] = Diagonal(E, B), [:I] = ΣTT
        part = (lambda A: A[:, 0:1]) if which == "TT" else (lambda A: A[:, 2:4])
        dpart = (lambda F: F[:, 0:1]) if which == "TT" else (lambda F: F[:, 1:3])
        sub = DataSet(proj, "I" if which == "TT" else "P", part(ds.Cf), part(ds.Cn), part(ds.Cnhat), part(ds.B), part(ds.Bhat), part(ds.Mf), None,
                      dpart(ds.d), ds.L, Cphi=ds.Cphi, Cftilde=part(ds.Cftilde))
        return quadratic_estimate(sub, which, wiener_filtered, weights, AL, None if d2 is None else dpart(d2), abs_each_term)
    assert pol in (("I",) if which == "TT" else ("P",))
    leg = _QELegs(proj)
    grad = {1: leg.d1, 2: leg.d2}
    TF = ds.Mf * ds.Bhat
    d1 = ds.d
    d2 = d1 if d2 is None else d2
    inds = lambda D: list(product((1, 2), repeat=D))
    div = lambda a, b: (np.divide(a, b, out=np.full(np.broadcast_shapes(a.shape, b.shape), np.nan, dtype=np.result_type(a, b)), where=(b != 0)))
    sl = lambda A, c: A[:, c:c + 1]
    fou = lambda m: rfft2(m.astype(proj.T))
    def norm(A):
        terms = [grad[i] * grad[j] * fou(A(i, j)) for i, j in inds(2)]
        tot = sum(np.abs(t) for t in terms) if abs_each_term else np.abs(sum(terms))
        return pinv_diag(tot.astype(proj.T))
    if which == "TT":
        S = TF ** 2 * ds.Cftilde + ds.Cnhat
        CT = ds.Cf if weights == "unlensed" else ds.Cftilde
        a, b = diag_ldiv(S, TF * d1), CT * diag_ldiv(S, TF * d2)
        unnorm = -sum(grad[i] * fou(leg(a) * leg(b, [i])) for i in (1, 2))
        if AL is None:
            X2, X1, X0 = div(TF ** 2 * CT ** 2, S), div(TF ** 2 * CT, S), div(TF ** 2, S)
            A = lambda i, j: leg(X2, [i], [j]) * leg(X0) + leg(X1, [i]) * leg(X1, [j])
            AL = norm(A)
    else:
        E, B = 0, 1
        TF2E, TF2B = sl(TF, E) ** 2, sl(TF, B) ** 2
        SE, SB = TF2E * sl(ds.Cftilde, E) + sl(ds.Cnhat, E), TF2B * sl(ds.Cftilde, B) + sl(ds.Cnhat, B)
        Cw = ds.Cf if weights == "unlensed" else ds.Cftilde
        CE, CB = sl(Cw, E), sl(Cw, B)
        tE1, tE2, tB2 = sl(TF * d1, E), sl(TF * d2, E), sl(TF * d2, B)
        if which == "EE":
            a1, a2, b2 = CE * diag_ldiv(SE, tE1), diag_ldiv(SE, tE1), diag_ldiv(SE, tE2)
            a1 = CE * diag_ldiv(SE, tE1)
            I = lambda i: -(2 * sum(leg(a1, [i], j, k) * leg(b2, j, k) for j, k in inds(2)) - leg(a1, [i]) * leg(b2))
            unnorm = sum(grad[i] * fou(I(i)) for i in (1, 2))
            if AL is None:
                X2, X1, X0 = div(TF2E * CE ** 2, SE), div(TF2E * CE, SE), div(TF2E, SE)
                A1 = lambda i, j: -4 * sum(_eps(m, p) * _eps(n, q) * (leg(X2, [i], [j], k, l, m, n) * leg(X0, k, l, p, q)
                                                                        + leg(X1, [i], k, l, m, n) * leg(X1, [j], k, l, p, q))
                                           for k, l, m, n, p, q in inds(6) if _eps(m, p) and _eps(n, q))
                A2 = lambda i, j: leg(X2, [i], [j]) * leg(X0) + leg(X1, [i]) * leg(X1, [j])
                AL = norm(lambda i, j: A1(i, j) + A2(i, j))
        else:
            aE, aE0 = CE * diag_ldiv(SE, tE1), diag_ldiv(SE, tE1)
            bB, bB0 = CB * diag_ldiv(SB, tB2), diag_ldiv(SB, tB2)
            I = lambda i: 2 * sum(_eps(k, l) * (leg(aE, [i], j, k) * leg(bB0, j, l) - leg(aE0, j, k) * leg(bB, [i], j, l))
                                  for j, k, l in inds(3) if _eps(k, l))
            unnorm = sum(grad[i] * fou(I(i)) for i in (1, 2))
            if AL is None:
                XE2, XE1, XE0 = div(TF2E * CE ** 2, SE), div(TF2E * CE, SE), div(TF2E, SE)
                XB2, XB1, XB0 = div(TF2B * CB ** 2, SB), div(TF2B * CB, SB), div(TF2B, SB)
                A = lambda i, j: 4 * sum(_eps(m, p) * _eps(n, q) * (leg(XE2, [i], [j], k, l, m, n) * leg(XB0, k, l, p, q)
                                                                    - 2 * leg(XE1, [i], k, l, m, n) * leg(XB1, [j], k, l, p, q)
                                                                    + leg(XE0, k, l, m, n) * leg(XB2, [i], [j], k, l, p, q))
                                         for k, l, m, n, p, q in inds(6) if _eps(m, p) and _eps(n, q))
                AL = norm(A)
    Nphi = AL
    phi = (AL * unnorm).astype(proj.cT)
    if wiener_filtered:
        phi = (ds.Cphi * pinv_diag(ds.Cphi + Nphi) * phi).astype(proj.cT)
    return dict(phi_qe=phi, AL=AL, Nphi=Nphi)


# ----------------------------------------------------------------------------------------------
# HMC step in ϕ° of the Gibbs sampler (src/sampling.jl:14-55 symplectic_integrate, :397-425 gibbs_sample_ϕ!, hmc_step, mass_matrix_ϕ)
# ----------------------------------------------------------------------------------------------
def symplectic_integrate(x0, p0, Lam, U, dUdx, dot, N: int = 50, eps: float = 0.1):
    """src/sampling.jl:14-55: leap-frog with mass matrix Λ on the *log-density* U (H = U − p·Λ⁻¹p/2).  Returns (ΔH, x, p)."""
    bc = lambda v: np.asarray(v).reshape((-1,) + (1,) * (x0.ndim - 1))
    T = x0.real.dtype.type
    H = lambda x, p: U(x) - dot(p, diag_ldiv(Lam, p)) / 2
    x, p = x0, p0
    g = dUdx(x)
    for _ in range(N):
        x1 = (x - T(eps) * diag_ldiv(Lam, p - T(eps) / 2 * g)).astype(x0.dtype)
        g1 = dUdx(x1)
        p = (p - T(eps) / 2 * (g1 + g)).astype(x0.dtype)
        x, g = x1, g1
    return H(x, p) - H(x0, p0), x, p


def mass_matrix_phi(ds: DataSet) -> np.ndarray:
    """mass_matrix_ϕ (src/sampling.jl:422-425): pinv(G)² (pinv(Cϕ) + pinv(Nϕ))."""
    m = pinv_diag(ds.Cphi) + pinv_diag(ds.Nphi)
    return m if ds.G is None else (pinv_diag(ds.G) ** 2 * m).astype(m.dtype)


def hmc_step_phi(ds: DataSet, f_mixed_map, phi_mixed, white_map, uniforms, N: int = 25, eps: float = 0.01, always_accept: bool = False,
                 bug_compat: bool = True):
    """gibbs_sample_ϕ! / hmc_step (src/sampling.jl:397-417) with the momentum draw p = √Λ·rfft(white) and the accept draws given
    explicitly.  Returns (ϕ°, ΔH per batch item, accept)."""
    proj = ds.proj
    Lam = mass_matrix_phi(ds)
    dot = lambda a, b: dot_fourier(proj, a, b)
    U = lambda x: logpdf_mixed(ds, f_mixed_map, x)
    dU = lambda x: gradient_logpdf_mixed(ds, f_mixed_map, x, bug_compat)[1]
    p0 = (np.sqrt(Lam) * rfft2(white_map)).astype(proj.cT)
    dH, xt, _ = symplectic_integrate(phi_mixed, p0, Lam, U, dU, dot, N, eps)
    accept = np.logical_or(always_accept, np.log(uniforms) < dH)
    a = accept.astype(proj.T).reshape(-1, 1, 1, 1)
    return (a * xt + (1 - a) * phi_mixed).astype(proj.cT), dH, accept


def sample_joint(ds: DataSet, phi_start, draws, symp_N: int = 25, symp_eps: float = 0.01, nburnin_always_accept: int = 10,
                 conjgrad_kwargs=dict(tol=1e-1, nsteps=500), bug_compat: bool = True):
    """The (f, ϕ) Gibbs sampler of sample_joint (src/sampling.jl:180-336) without θ: per step gibbs_sample_f! (sample_f), gibbs_mix!,
    gibbs_sample_ϕ! (one HMC update of ϕ°), gibbs_unmix! (:388-451); steps are numbered from 2 like the reference and proposals
    are always accepted while step < nburnin_always_accept.  `draws[i]` = dict(wf, wn, wp, u): the white maps of sample_f, the
    momentum white map and the accept uniforms of step i.  Returns the chain as a list of dicts."""
    proj, pol = ds.proj, ds.pol
    phi, chain = phi_start, []
    for i, dr in enumerate(draws):
        step = i + 2
        ds.L = precompute(proj, phi, ds.L.nsteps, phi_is_fourier=True)
        f, _ = sample_f(ds, dr["wf"], dr["wn"], **conjgrad_kwargs)
        fm, pm = mix(ds, proj, pol, f, phi, D=ds.D, G=ds.G, nsteps=ds.L.nsteps)
        pm, dH, acc = hmc_step_phi(ds, fm, pm, dr["wp"], dr["u"], N=symp_N, eps=symp_eps, always_accept=(step < nburnin_always_accept), bug_compat=bug_compat)
        f, phi = unmix(ds, proj, pol, fm, pm, D=ds.D, G=ds.G, nsteps=ds.L.nsteps)
        chain.append(dict(step=step, f=f, phi=phi, dH=dH, accept=acc, logpdf=logpdf(ds, f, phi)))
    return chain


# ----------------------------------------------------------------------------------------------
# Synthetic flat-sky inputs (harness; mirrors load_sim defaults, src/dataset.jl:186-338)
# ----------------------------------------------------------------------------------------------
def load_fiducial_cls(path=None):
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fiducial_cls.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def simulate_diag(proj: ProjLambert, cov_half: np.ndarray, rng, nb=1) -> np.ndarray:
    """simulate(rng, D::DiagOp) = sqrt(D) * randn (src/specialops.jl:6): white Map noise → Fourier → ×√C."""
    npol = cov_half.shape[1]
    w = rng.standard_normal((nb, npol) + proj.map_shape).astype(proj.T)
    return (np.sqrt(cov_half) * rfft2(w)).astype(proj.cT)


def cosine_border_mask(proj: ProjLambert, border_deg: float = 1.0) -> np.ndarray:
    """Harness-made pixel mask: 1 inside, cosine-apodised to 0 over `border_deg` at each edge."""
    def prof(n):
        w = max(2, int(round(border_deg * 60.0 / proj.theta_pix)))
        w = min(w, n // 4)
        x = np.ones(n)
        ramp = 0.5 * (1 - np.cos(np.pi * (np.arange(w) + 0.5) / w))
        x[:w] = ramp
        x[-w:] = ramp[::-1]
        return x
    return np.outer(prof(proj.Nx), prof(proj.Ny)).astype(proj.T)


def make_dataset(Ny, Nx, theta_pix, pol="I", T=np.float64, nb=1, seed=0, nsteps=7, mask=True,
                 muK_arcmin_T=3.0, lknee=100.0, aknee=3.0, beam_fwhm=0.0, lowpass=3000, cls=None,
                 mask_border_deg=1.0):
    """load_sim (src/dataset.jl:186-338) for pol ∈ {I, P, IP}; returns dict(f, phi, d, ds, proj) with harmonic-basis
    fields.  For IP every harmonic operator is a BlockDiagIEB (4 planes [TT, TE, EE, BB]); the TE entry of the noise, mask
    and beam blocks is zero (:279,:299, src/cls.jl:298)."""
    cls = cls or load_fiducial_cls()
    proj = ProjLambert(Ny, Nx, theta_pix, T)
    ell = cls["ell"].astype(np.float64)
    rng = np.random.default_rng(seed)
    keys = {"I": ("TT",), "P": ("EE", "BB"), "IP": ("TT", "TE", "EE", "BB")}[pol]
    npol = {"I": 1, "P": 2, "IP": 3}[pol]
    lb, wl = lowpass_wl(lowpass)
    zero = lambda: np.zeros(proj.fourier_shape, dtype=proj.T)
    Cf = np.stack([cl_to_cov(proj, ell, cls["ut_" + k]) for k in keys])[None]
    Cphi = cl_to_cov(proj, ell, cls["pp"])[None, None]
    Cn = np.stack([zero() if k == "TE" else cl_to_cov(proj, ell, noise_cls(ell, muK_arcmin_T, lknee, aknee, pol=(k != "TT"))) for k in keys])[None]
    Mf = np.stack([zero() if k == "TE" else cl_to_cov(proj, lb, wl, units=1) for k in keys])[None]
    B = np.stack([zero() if k == "TE" else cl_to_cov(proj, ell, np.sqrt(beam_cls(ell, beam_fwhm)), units=1) for k in keys])[None]
    Mpix = None
    if mask:
        m = cosine_border_mask(proj, mask_border_deg)
        Mpix = np.broadcast_to(m, (1, npol) + m.shape).copy()
    if pol == "IP":          # simulate(rng, L::BlockDiagIEB) = sqrt(L) * randn (src/specialops.jl:94)
        sim = lambda C: block_mul(block_sqrt(C), rfft2(rng.standard_normal((nb, 3) + proj.map_shape).astype(proj.T))).astype(proj.cT)
    else:
        sim = lambda C: simulate_diag(proj, C, rng, nb)
    f = sim(Cf)
    phi = simulate_diag(proj, Cphi, rng, nb)
    n = sim(Cn)
    L = precompute(proj, phi, nsteps, phi_is_fourier=True)
    ds = DataSet(proj, pol, Cf, Cn, Cn.copy(), B, B.copy(), Mf, Mpix, None, L)
    ds.Cphi = Cphi
    ds.Nphi = (Cphi * 0 + np.median(Cphi[Cphi > 0]) if np.any(Cphi > 0) else Cphi + 1).astype(proj.T)   # stand-in; load_sim's value is quadratic_estimate(ds).Nϕ / 2 (:316), see `nphi_from_qe`
    if "tot_TT" in cls:
        ds.Cftilde = np.stack([cl_to_cov(proj, ell, cls["tot_" + k]) for k in keys])[None]
    ft = lenseflow_apply(L, OP_L, to_lense_basis(pol, proj, f))
    d = (apply_M(ds, op_mul(pol, B, to_harmonic_basis(pol, proj, ft))) + n).astype(proj.cT)
    ds.d = d
    return dict(f=f, phi=phi, d=d, ds=ds, proj=proj, Cphi=Cphi)

"""cmblensing.jl_b200 — host-side mirror of CMBLensing.jl's flat-sky operator API over libcmbl_b200.so.

The reference's host language is Julia, which is not available in this image, so this thin Python layer stands where
`CMBLensingB200Ext.jl` would: same names, argument meaning and error behaviour as the reference for the
hot path — `ProjLambert`, `FlatMap/FlatFourier/FlatQUMap/...`, basis conversions `Map(f)`, `Fourier(f)`, `QUFourier(f)`,
`EBFourier(f)`, `LenseFlow(ϕ) * f`, `L.H * f`, `L.ldiv(f)` (Julia `L \\ f`), `DiagOp`, `dot`, `BaseDataSet`,
`gradientf_logpdf`, `argmaxf_logpdf`, `conjugate_gradient`.  All arithmetic runs in the CUDA library (sm_100a); torch
is used only to own device memory and streams.  There is no CPU fallback.

Array layout: Julia's column-major (Ny, Nx, Npol, Nb) is the C-order tensor [Nb, Npol, Nx, Ny] (Ny fastest); Fourier
arrays are [Nb, Npol, Nx, Ny//2+1] complex (src/proj_cartesian.jl:51-56).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import CmblError, DatasetDesc, FOURIER, MAP, OP_L, OP_LH, OP_LHINV, OP_LINV, load

__all__ = [
    "ProjLambert", "Field", "FlatMap", "FlatFourier", "FlatQUMap", "FlatQUFourier", "FlatEBMap", "FlatEBFourier",
    "FlatIQUMap", "FlatIQUFourier", "FlatIEBMap", "FlatIEBFourier",
    "Map", "Fourier", "QUMap", "QUFourier", "EBMap", "EBFourier", "IQUMap", "IQUFourier", "IEBMap", "IEBFourier",
    "LenseBasis", "DerivBasis", "HarmonicBasis",
    "LenseFlow", "CachedLenseFlow", "DiagOp", "Diagonal", "BlockDiagIEB", "dot", "BaseDataSet", "gradientf_logpdf",
    "Hessian_logpdf_preconditioner", "mix", "unmix", "argmaxf_logpdf", "argmaxf_lnP", "conjugate_gradient_wiener", "batch", "unbatch",
    "Cℓ_to_2D", "Cℓ_to_Cov", "Cl_to_Cov", "simulate", "sample_f", "convert",
    "quadratic_estimate", "mixing_D", "logdet", "logpdf", "Mixed", "gradient_logpdf_mixed", "MAP_joint", "symplectic_integrate", "hmc_step", "mass_matrix_ϕ", "gibbs_sample_ϕ", "sample_joint",
    "CmblError", "load",
]

_TORCH_REAL = {0: torch.float32, 1: torch.float64}
_TORCH_CPLX = {0: torch.complex64, 1: torch.complex128}
_NP_REAL = {0: np.float32, 1: np.float64}


def _ptr(t: torch.Tensor) -> c_void_p:
    return c_void_p(t.data_ptr())


def _stream(t: torch.Tensor) -> c_void_p:
    if t.is_cuda:
        return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return c_void_p(0)


# ------------------------------------------------------------------------------------------------------------------
# ProjLambert (src/proj_lambert.jl:24-75) + FFT plan (src/util_fft.jl:32-39)
# ------------------------------------------------------------------------------------------------------------------
class ProjLambert:
    _memo: dict = {}

    def __new__(cls, Ny, Nx=None, θpix=1.0, T=torch.float32, device="cuda:0", lib=None):
        Nx = Ny if Nx is None else Nx
        lib = lib or load()
        key = (int(Ny), int(Nx), float(θpix), T, str(device), lib.path)
        if key in cls._memo:                                      # @memoize, src/proj_lambert.jl:47
            return cls._memo[key]
        self = super().__new__(cls)
        self._init(int(Ny), int(Nx), float(θpix), T, torch.device(device), lib)
        cls._memo[key] = self
        return self

    def _init(self, Ny, Nx, θpix, T, device, lib):
        if T not in (torch.float32, torch.float64):
            raise CmblError("T must be torch.float32 or torch.float64")
        if device.type == "cuda" and lib.is_emulator:
            raise CmblError("the host emulator build cannot drive CUDA tensors")
        if device.type != "cuda" and not lib.is_emulator:
            raise CmblError("cmblensing.jl_b200 needs a CUDA device (no CPU fallback)")
        self.lib, self.Ny, self.Nx, self.θpix, self.T, self.device = lib, Ny, Nx, θpix, T, device
        self.dtype_code = 0 if T == torch.float32 else 1
        self.cT = _TORCH_CPLX[self.dtype_code]
        self.handle = c_void_p()
        dev_index = device.index or 0
        lib.call("cmbl_plan_create", byref(self.handle), dev_index, Ny, Nx, c_double(θpix), self.dtype_code)
        npT = _NP_REAL[self.dtype_code]
        Nyh = Ny // 2 + 1
        self.ℓx, self.ℓy, self.λ_rfft = np.zeros(Nx, npT), np.zeros(Nyh, npT), np.zeros(Nyh, npT)
        self.sin2ϕ, self.cos2ϕ = np.zeros((Nx, Nyh), npT), np.zeros((Nx, Nyh), npT)
        sc = (c_double * 5)()
        P = lambda a: a.ctypes.data_as(c_void_p)
        lib.call("cmbl_plan_grids", self.handle, P(self.ℓx), P(self.ℓy), P(self.λ_rfft), P(self.sin2ϕ), P(self.cos2ϕ), sc)
        self.Δx, self.Δℓx, self.Δℓy, self.Ωpix, self.nyquist = (float(v) for v in sc)
        self.ℓmag = np.sqrt(self.ℓx[:, None] ** 2 + self.ℓy[None, :] ** 2).astype(npT)     # :65

    @property
    def Nyh(self):
        return self.Ny // 2 + 1

    def map_shape(self, Npol, Nb):
        return (Nb, Npol, self.Nx, self.Ny)

    def fourier_shape(self, Npol, Nb):
        return (Nb, Npol, self.Nx, self.Nyh)


# ------------------------------------------------------------------------------------------------------------------
# Fields (src/base_fields.jl:14-21; bases src/generic.jl:9-16)
# ------------------------------------------------------------------------------------------------------------------
_BASES = {"Map": (1, False, None), "Fourier": (1, True, None), "QUMap": (2, False, "QU"), "QUFourier": (2, True, "QU"),
          "EBMap": (2, False, "EB"), "EBFourier": (2, True, "EB"),
          "IQUMap": (3, False, "IQU"), "IQUFourier": (3, True, "IQU"), "IEBMap": (3, False, "IEB"), "IEBFourier": (3, True, "IEB")}


class Field:
    """BaseField{B,M,T,A}: an array plus ProjLambert metadata, tagged with its basis."""

    def __init__(self, basis: str, arr: torch.Tensor, proj: ProjLambert):
        if basis not in _BASES:
            raise CmblError(f"unknown basis {basis}")
        npol, four, _ = _BASES[basis]
        if arr.dim() == 2:
            arr = arr[None, None]
        elif arr.dim() == 3:
            arr = arr[None]
        want = proj.fourier_shape(npol, arr.shape[0]) if four else proj.map_shape(npol, arr.shape[0])
        if tuple(arr.shape) != tuple(want):
            raise CmblError(f"Size-mismatched array {tuple(arr.shape)} and metadata {want} in Field constructor.")   # base_fields.jl:19
        want_dt = proj.cT if four else proj.T
        self.basis, self.proj = basis, proj
        self.arr = arr.to(device=proj.device, dtype=want_dt).contiguous()

    # -- metadata -------------------------------------------------------------------------------------------
    @property
    def Npol(self): return _BASES[self.basis][0]
    @property
    def is_fourier(self): return _BASES[self.basis][1]
    @property
    def Nbatch(self): return self.arr.shape[0]
    @property
    def C(self): return self.arr.shape[0] * self.arr.shape[1]

    def _like(self, arr, basis=None):
        return Field(basis or self.basis, arr, self.proj)

    def _check(self, o):
        if o.proj is not self.proj:
            raise CmblError("Can't broadcast fields with mismatched metadata (src/proj_lambert.jl:111-114)")

    # -- broadcast-style arithmetic (generic, not the hot path) ------------------------------------------------
    def _bin(self, o, fn):
        if isinstance(o, Field):
            self._check(o)
            o = convert(o, self.basis).arr
        elif isinstance(o, (list, tuple, np.ndarray, torch.Tensor)):                 # BatchedReal (src/batching.jl:9-45)
            o = torch.as_tensor(o, device=self.arr.device, dtype=self.proj.T).reshape(-1, 1, 1, 1)
        return self._like(fn(self.arr, o))

    def _axpby(self, a, y=None, b=1.0):
        """out = a .* self + b .* y with per-batch scalars (cmbl_field_axpby): the linear combinations of the callers (leap-frog, RK and
        line-search updates, `x + α*Δ`) run on the library's kernel, not on eager tensor ops.  None if the case is not covered."""
        if self.Nbatch > 64:
            return None
        av = np.atleast_1d(np.asarray(a, dtype=np.float64)).ravel(); bv = np.atleast_1d(np.asarray(b, dtype=np.float64)).ravel()
        if y is not None:
            self._check(y)
            if y.basis != self.basis:
                y = convert(y, self.basis)
            if y.arr.shape != self.arr.shape:
                return None                                         # broadcasting batches: generic path
        if av.size not in (1, self.Nbatch) or bv.size not in (1, self.Nbatch):
            return None
        out = torch.empty_like(self.arr)
        A = (c_double * av.size)(*av); B = (c_double * bv.size)(*bv)
        self.proj.lib.call("cmbl_field_axpby", self.proj.handle, FOURIER if self.is_fourier else MAP, A, int(av.size), _ptr(self.arr), B, int(bv.size),
                           _ptr(y.arr) if y is not None else c_void_p(0), _ptr(out), self.Npol, self.Nbatch, _stream(out))
        return self._like(out)

    @staticmethod
    def _is_scalars(o):
        return isinstance(o, (int, float, np.floating, np.integer)) or (isinstance(o, (list, tuple, np.ndarray)) and np.ndim(o) <= 1)

    def __add__(self, o):
        r = self._axpby(1.0, o, 1.0) if isinstance(o, Field) else None
        return r if r is not None else self._bin(o, torch.add)

    def __sub__(self, o):
        r = self._axpby(1.0, o, -1.0) if isinstance(o, Field) else None
        return r if r is not None else self._bin(o, torch.sub)

    def __mul__(self, o):
        r = self._axpby(o) if self._is_scalars(o) else None
        return r if r is not None else self._bin(o, torch.mul)

    def __truediv__(self, o): return self._bin(o, torch.div)
    __radd__ = __add__
    __rmul__ = __mul__
    def __neg__(self):
        r = self._axpby(-1.0)
        return r if r is not None else self._like(-self.arr)
    def __rsub__(self, o): return (-self) + o
    def zero(self): return self._like(torch.zeros_like(self.arr))
    def copy(self): return self._like(self.arr.clone())
    def cpu_numpy(self): return self.arr.detach().cpu().numpy()
    def batch_index(self, i): return self._like(self.arr[i:i + 1].clone())                  # src/proj_lambert.jl:456-459
    def __repr__(self): return f"Flat{self.basis}(Ny={self.proj.Ny}, Nx={self.proj.Nx}, Nbatch={self.Nbatch}, T={self.proj.T})"


def _mk(basis):
    def ctor(arr, proj=None, θpix=1.0, device="cuda:0", lib=None):
        if isinstance(arr, np.ndarray):
            arr = torch.from_numpy(np.ascontiguousarray(arr))
        if proj is None:
            Nx, Ny = arr.shape[-2], arr.shape[-1]
            if _BASES[basis][1]:
                raise CmblError("pass proj= when constructing a Fourier-basis field (Ny is ambiguous)")
            T = torch.float64 if arr.dtype in (torch.float64, torch.complex128) else torch.float32
            proj = ProjLambert(Ny, Nx, θpix, T, device, lib)
        return Field(basis, arr, proj)
    ctor.__name__ = "Flat" + basis
    return ctor


FlatMap, FlatFourier, FlatQUMap, FlatQUFourier, FlatEBMap, FlatEBFourier = (
    _mk(b) for b in ("Map", "Fourier", "QUMap", "QUFourier", "EBMap", "EBFourier"))
FlatIQUMap, FlatIQUFourier, FlatIEBMap, FlatIEBFourier = (_mk(b) for b in ("IQUMap", "IQUFourier", "IEBMap", "IEBFourier"))


def batch(fields):
    """batch(fs...) — concatenate along the batch dimension (src/proj_lambert.jl:446-449)."""
    f0 = fields[0]
    return f0._like(torch.cat([f.arr for f in fields], dim=0))


def unbatch(f):
    return [f.batch_index(i) for i in range(f.Nbatch)]


# -- basis conversion: the FFT call sites (src/proj_lambert.jl:245-300) -------------------------------------------------
def _fft(f: Field, inverse: bool) -> torch.Tensor:
    p = f.proj
    if inverse:
        out = torch.empty(p.map_shape(f.Npol, f.Nbatch), dtype=p.T, device=p.device)
        p.lib.call("cmbl_irfft2", p.handle, _ptr(f.arr), _ptr(out), f.C, _stream(out))
    else:
        out = torch.empty(p.fourier_shape(f.Npol, f.Nbatch), dtype=p.cT, device=p.device)
        p.lib.call("cmbl_rfft2", p.handle, _ptr(f.arr), _ptr(out), f.C, _stream(out))
    return out


def _rot(f: Field, to_eb: bool) -> torch.Tensor:
    p = f.proj
    out = torch.empty_like(f.arr)
    if f.Npol == 3:                                   # IQU <-> IEB: I is copied, the rotation acts on planes 2:3 (src/proj_lambert.jl:284,292)
        out[:, 0] = f.arr[:, 0]
    p.lib.call("cmbl_qu_eb", p.handle, 1 if to_eb else 0, _ptr(f.arr), _ptr(out), f.Nbatch, f.Npol, f.Npol - 2, _stream(out))
    return out


def convert(f: Field, basis: str) -> Field:
    if f.basis == basis:
        return f
    npol, four, pb = _BASES[basis]
    if npol != f.Npol:
        raise CmblError(f"cannot convert {f.basis} to {basis}")
    cur = f
    cur_pb = _BASES[cur.basis][2]
    if pb != cur_pb:                                  # the rotation lives in Fourier space
        if not cur.is_fourier:
            cur = cur._like(_fft(cur, False), cur_pb + "Fourier")
        cur = cur._like(_rot(cur, to_eb=pb.endswith("EB")), pb + "Fourier")
    if cur.is_fourier != four:
        name = (pb or "") + ("Fourier" if four else "Map")
        cur = cur._like(_fft(cur, inverse=not four), name)
    return cur


def Map(f): return convert(f, "Map")
def Fourier(f): return convert(f, "Fourier" if f.Npol == 1 else f.basis.replace("Map", "Fourier"))
def QUMap(f): return convert(f, "QUMap")
def QUFourier(f): return convert(f, "QUFourier")
def EBMap(f): return convert(f, "EBMap")
def EBFourier(f): return convert(f, "EBFourier")
def IQUMap(f): return convert(f, "IQUMap")
def IQUFourier(f): return convert(f, "IQUFourier")
def IEBMap(f): return convert(f, "IEBMap")
def IEBFourier(f): return convert(f, "IEBFourier")
def LenseBasis(f): return convert(f, ("Map", "QUMap", "IQUMap")[f.Npol - 1])                  # Ł, src/generic.jl:88-93
def DerivBasis(f): return convert(f, ("Fourier", "QUFourier", "IQUFourier")[f.Npol - 1])      # Ð
def HarmonicBasis(f): return convert(f, ("Fourier", "EBFourier", "IEBFourier")[f.Npol - 1])


def dot(a: Field, b: Field) -> np.ndarray:
    """dot(a,b) (src/proj_lambert.jl:318-328): per-batch values (BatchedReal) as a float64 array."""
    a._check(b)
    if a.basis != b.basis:                                          # mixed bases → both to the Ð basis (:328)
        a, b = DerivBasis(a), DerivBasis(b)
    p = a.proj
    nb = max(a.Nbatch, b.Nbatch)
    if a.Nbatch != b.Nbatch:
        a = a._like(a.arr.expand(nb, -1, -1, -1).contiguous()); b = b._like(b.arr.expand(nb, -1, -1, -1).contiguous())
    out = (c_double * nb)()
    p.lib.call("cmbl_dot", p.handle, FOURIER if a.is_fourier else MAP, _ptr(a.arr), _ptr(b.arr), a.Npol, nb, out, _stream(a.arr))
    return np.array(out[:], dtype=np.float64)


def get_max_lensing_step(ϕ: Field, η: Field, per_batch: bool = False):
    """get_max_lensing_step(ϕ, η) (src/lenseflow.jl:242-256): αmax such that 𝕀 + ∇∇(ϕ + α η) has non-zero determinant in every pixel for
    all α in [0, αmax) — the largest step along η that keeps ϕ + α η in the weak-lensing regime LenseFlow can handle.  The reference
    returns ONE number (the minimum over pixels and batch); `per_batch=True` returns the per-item values the device computes."""
    ϕ._check(η)
    if ϕ.Npol != 1 or η.Npol != 1:
        raise CmblError("get_max_lensing_step expects spin-0 fields (ϕ, η)")
    p = ϕ.proj
    nb = max(ϕ.Nbatch, η.Nbatch)
    if ϕ.Nbatch != η.Nbatch:
        ϕ = ϕ._like(ϕ.arr.expand(nb, -1, -1, -1).contiguous()); η = η._like(η.arr.expand(nb, -1, -1, -1).contiguous())
    out = (c_double * nb)()
    p.lib.call("cmbl_max_lensing_step", p.handle, _ptr(ϕ.arr), FOURIER if ϕ.is_fourier else MAP, _ptr(η.arr), FOURIER if η.is_fourier else MAP,
               nb, out, _stream(ϕ.arr))
    v = np.array(out[:], dtype=np.float64)
    return v if per_batch else float(v.min())


# ------------------------------------------------------------------------------------------------------------------
# DiagOp (src/specialops.jl:9-22)
# ------------------------------------------------------------------------------------------------------------------
class DiagOp:
    """Diagonal(f): `D * f` converts f to D's basis and multiplies; `D.ldiv(f)` is Julia's `D \\ f` with nan2zero."""

    def __init__(self, diag: Field):
        d = diag
        if d.is_fourier and d.arr.is_complex():
            if float(d.arr.imag.abs().max()) != 0.0:
                raise CmblError("DiagOp on this path needs a real diagonal (Cf, Cn, B, masks)")
            self._real = d.arr.real.contiguous()
        else:
            self._real = d.arr.contiguous()
        self.diag = d

    def _apply(self, f: Field, ldiv: bool) -> Field:
        f = convert(f, self.diag.basis)
        p = f.proj
        out = torch.empty_like(f.arr)
        Cd = self._real.shape[0] * self._real.shape[1]
        if self._real.shape[0] not in (1, f.Nbatch):
            raise CmblError("batch sizes must be equal or 1 (src/batching.jl)")
        p.lib.call("cmbl_diag_mul", p.handle, FOURIER if f.is_fourier else MAP, _ptr(self._real), Cd, _ptr(f.arr), _ptr(out),
                   f.C, 1 if ldiv else 0, _stream(out))
        return f._like(out)

    def __mul__(self, f): return self._apply(f, False)
    def ldiv(self, f): return self._apply(f, True)

    def pinv(self):
        r = torch.where(self._real == 0, torch.zeros_like(self._real), 1 / self._real)
        return DiagOp(Field(self.diag.basis, r.to(self.diag.arr.dtype), self.diag.proj))


Diagonal = DiagOp


class BlockDiagIEB:
    """BlockDiagIEB(ΣTE, ΣB) (src/specialops.jl:61-118): [ΣTT ΣTE; ΣTE ΣEE] ⊕ ΣBB acting on IEBFourier fields.  Built from
    four real Fourier-basis half-planes (what Cℓ_to_Cov(:IP) produces, src/proj_lambert.jl:368-371); `L * f`, `L.ldiv(f)`
    (= pinv(L) * f, :78) and `L.sqrt_mul(f)` (simulate, :94) run on the device."""

    def __init__(self, ΣTT, ΣTE, ΣEE, ΣBB, proj: ProjLambert | None = None):
        planes = []
        for a in (ΣTT, ΣTE, ΣEE, ΣBB):
            if isinstance(a, DiagOp):
                proj = proj or a.diag.proj
                a = a._real
            if isinstance(a, np.ndarray):
                a = torch.from_numpy(np.ascontiguousarray(a))
            planes.append(a.reshape(a.shape[-2], a.shape[-1]))
        if proj is None:
            raise CmblError("BlockDiagIEB from raw arrays needs proj=")
        self.proj = proj
        self._real = torch.stack(planes)[None].to(device=proj.device, dtype=proj.T).contiguous()      # [1, 4, Nx, Nyh]
        if tuple(self._real.shape) != (1, 4, proj.Nx, proj.Nyh):
            raise CmblError("BlockDiagIEB planes must be (Ny÷2+1, Nx) half-planes")

    def _apply(self, f: Field, mode: int) -> Field:
        f = convert(f, "IEBFourier")                                          # L * IEBFourier(f), :77
        p = f.proj
        out = torch.empty_like(f.arr)
        p.lib.call("cmbl_blockdiag_ieb", p.handle, mode, _ptr(self._real), _ptr(f.arr), _ptr(out), f.Nbatch, _stream(out))
        return f._like(out)

    def __mul__(self, f): return self._apply(f, 0)
    def ldiv(self, f): return self._apply(f, 1)
    def sqrt_mul(self, f): return self._apply(f, 2)


# -- covariance operators from power spectra (setup; src/proj_lambert.jl:173-175,361-371, src/numerical_algorithms.jl:148-177) ---
def Cℓ_to_2D(proj: ProjLambert, ℓ, Cℓ) -> np.ndarray:
    """nan2zero.(Cℓ.(ℓmag)) with the reference's LinearInterpolation (NaN outside the table, so ℓ = 0 and ℓ > ℓmax map to 0)."""
    ℓ, Cℓ = np.asarray(ℓ, dtype=np.float64), np.asarray(Cℓ, dtype=np.float64)
    x = proj.ℓmag.astype(np.float64)
    i = np.clip(np.searchsorted(ℓ, x, side="left") - 1, 0, len(ℓ) - 2)
    m = np.diff(Cℓ) / np.diff(ℓ)
    y = Cℓ[i] + m[i] * (x - ℓ[i])
    y = np.where((x < ℓ[0]) | (x > ℓ[-1]) | ~np.isfinite(y), 0.0, y)
    return y.astype(_NP_REAL[proj.dtype_code])


def Cℓ_to_Cov(pol: str, proj: ProjLambert, ℓ, *Cℓs, units=None):
    """Cℓ_to_Cov(:I | :P | :IP, proj, Cℓ...; units=Ωpix): Diagonal(Fourier), Diagonal(EBFourier) from (EE, BB), or
    BlockDiagIEB from (TT, EE, BB, TE) — argument order as the reference (src/proj_lambert.jl:361-371)."""
    want = {"I": 1, "P": 2, "IP": 4}.get(pol)
    if want is None:
        raise CmblError("`pol` should be one of I, P, or IP")                     # src/dataset.jl:263
    if len(Cℓs) != want:
        raise CmblError(f"Cℓ_to_Cov({pol}) takes {want} spectra")
    # the interpolation onto the ℓmag grid runs on the device (cmbl_cl_to_cov): one real half-plane per spectrum
    ell = np.ascontiguousarray(ℓ, dtype=np.float64)
    out = torch.empty((want,) + proj.fourier_shape(1, 1)[2:], dtype=proj.T, device=proj.device)
    for i, c in enumerate(Cℓs):
        cl = np.ascontiguousarray(c, dtype=np.float64)
        if cl.shape != ell.shape:
            raise CmblError("Cℓ_to_Cov: ℓ and Cℓ must have the same length")
        proj.lib.call("cmbl_cl_to_cov", proj.handle, ell.ctypes.data_as(c_void_p), cl.ctypes.data_as(c_void_p), int(ell.size),
                      c_double(0.0 if units is None else float(units)), c_void_p(out[i].data_ptr()), _stream(out))
    if pol == "IP":
        TT, EE, BB, TE = (out[i] for i in range(4))
        return BlockDiagIEB(TT, TE, EE, BB, proj=proj)
    return DiagOp(Field("Fourier" if pol == "I" else "EBFourier", out[None], proj))


Cl_to_Cov = Cℓ_to_Cov


# ------------------------------------------------------------------------------------------------------------------
# LenseFlow (src/lenseflow.jl:19-60, src/flowops.jl:11-14)
# ------------------------------------------------------------------------------------------------------------------
class CachedLenseFlow:
    """CachedLenseFlow (src/lenseflow.jl:33-60): the library handle that owns the p (and M⁻¹) cache and the RK scratch.
    `precompute(ϕ)` refills the SAME device memory for a new ϕ — the reference's `precompute!!`, which re-caches only when
    `ϕ !== old` (src/lenseflow.jl:80-129) — so a line search or a sampler that moves ϕ never reallocates the multi-GB cache."""

    def __init__(self, ϕ: Field, n: int, Npol: int, Nb_f: int, with_minv=False):
        if ϕ.Npol != 1:
            raise CmblError("ϕ must be a spin-0 field")
        p = ϕ.proj
        self.proj, self.n, self.Npol, self.Nb_f, self.Nb_ϕ, self.ϕ = p, n, Npol, Nb_f, ϕ.Nbatch, None
        self.handle = c_void_p()
        self.with_minv = with_minv
        p.lib.call("cmbl_lenseflow_create", byref(self.handle), p.handle, n, Npol, Nb_f, ϕ.Nbatch)
        self.precompute(ϕ)

    def precompute(self, ϕ: Field):
        if ϕ is self.ϕ:
            return self
        if ϕ.Nbatch != self.Nb_ϕ or ϕ.proj is not self.proj:
            raise CmblError("precompute!!: ϕ does not match the cached LenseFlow (batch size / metadata)")
        self.proj.lib.call("cmbl_lenseflow_precompute", self.handle, _ptr(ϕ.arr), FOURIER if ϕ.is_fourier else MAP,
                           1 if self.with_minv else 0, _stream(ϕ.arr))
        self.ϕ = ϕ
        return self

    def get_p(self, k: int) -> np.ndarray:
        """p[τ_k] = M⁻¹ᵀ∇ϕ at stage k of the 2n+1 cached times (src/lenseflow.jl:131-142), as a host array [Nb_ϕ, 2, Nx, Ny] in the
        reference's layout (the library may hold it row-grouped internally)."""
        out = np.empty((self.Nb_ϕ, 2, self.proj.Nx, self.proj.Ny), dtype=_NP_REAL[self.proj.dtype_code])
        self.proj.lib.call("cmbl_lenseflow_get_p", self.handle, int(k), out.ctypes.data_as(c_void_p))
        return out

    def __del__(self):
        try:
            self.proj.lib.call("cmbl_lenseflow_destroy", self.handle)
        except Exception:
            pass

    def apply(self, op: int, f: Field) -> Field:
        p = self.proj
        if op in (OP_L, OP_LINV):
            f = LenseBasis(f)
        else:
            f = DerivBasis(f)
        out = torch.empty_like(f.arr)
        p.lib.call("cmbl_lenseflow_apply", self.handle, op, _ptr(f.arr), _ptr(out), _stream(out))
        return f._like(out)

    def pullback(self, op: int, f_out: Field, Δ: Field, bug_compat: bool = True):
        """Pullback of `Lϕ*f` (op = OP_L) or `Lϕ\\f` (op = OP_LINV) — the Zygote rule of src/flowops.jl:40-68 integrating
        negδvelocityᴴ (src/lenseflow.jl:176-214).  `f_out` is the forward result, `Δ` the cotangent; returns (δf, δϕ) in the
        Fourier basis (δϕ has one plane per batch item).  `bug_compat=True` reproduces the reference's aliased 2×2 product."""
        if not self.with_minv:
            raise CmblError("pullback needs LenseFlow(...).cache(f, with_minv=True)")
        p = self.proj
        f_out, Δ = LenseBasis(f_out), DerivBasis(Δ)
        δf = torch.empty_like(Δ.arr)
        δϕ = torch.empty((self.Nb_f, 1) + tuple(Δ.arr.shape[2:]), dtype=Δ.arr.dtype, device=Δ.arr.device)
        p.lib.call("cmbl_lenseflow_grad", self.handle, op, _ptr(f_out.arr), _ptr(Δ.arr), _ptr(δf), _ptr(δϕ), 1 if bug_compat else 0, _stream(δf))
        return Δ._like(δf), Field("Fourier", δϕ, p)


_FLOW_POOL_MAX = 6
_FLOW_POOL: dict = {}        # (proj, n, Npol, Nb_f, Nb_ϕ, with_minv) -> CachedLenseFlow: one device cache per shape, refilled per ϕ


class LenseFlow:
    """LenseFlow(ϕ, n=7): lazy wrapper.  Applying it fetches the cached handle for the field's shape and (re)fills it for this
    ϕ when it currently holds another one (precompute!!, src/lenseflow.jl:80-129): handles are pooled per shape, so distinct
    LenseFlow(ϕ) objects of the same shape share one p-cache allocation."""

    def __init__(self, ϕ: Field, n: int = 7):
        self.ϕ, self.n = ϕ, n

    def cache(self, f: Field, with_minv=False) -> CachedLenseFlow:
        if self.ϕ.Nbatch not in (1, f.Nbatch):
            raise CmblError("batch sizes must be equal or 1")
        key = (id(self.ϕ.proj), self.n, f.Npol, f.Nbatch, self.ϕ.Nbatch, bool(with_minv))
        c = _FLOW_POOL.pop(key, None)
        if c is None:
            c = CachedLenseFlow(self.ϕ, self.n, f.Npol, f.Nbatch, with_minv)
        _FLOW_POOL[key] = c                                                   # most recently used last
        while len(_FLOW_POOL) > _FLOW_POOL_MAX:
            _FLOW_POOL.pop(next(iter(_FLOW_POOL)))                            # handles still referenced elsewhere stay alive
        return c.precompute(self.ϕ)

    def __mul__(self, f): return self.cache(f).apply(OP_L, f)                    # Lϕ * f
    def ldiv(self, f): return self.cache(f).apply(OP_LINV, f)                    # Lϕ \ f
    @property
    def H(self): return _AdjointFlow(self)                                       # Lϕ'
    adjoint = H


class _AdjointFlow:
    def __init__(self, L): self.L = L
    def __mul__(self, f): return self.L.cache(f).apply(OP_LH, f)                 # Lϕ' * f
    def ldiv(self, f): return self.L.cache(f).apply(OP_LHINV, f)                 # Lϕ' \ f


# ------------------------------------------------------------------------------------------------------------------
# DataSet + CG Wiener filter (src/dataset.jl:37-137, src/maximization.jl:17-42, src/numerical_algorithms.jl:73-134)
# ------------------------------------------------------------------------------------------------------------------
class BaseDataSet:
    """The fields of BaseDataSet used by argmaxf_logpdf: d, Cf, Cn, Cn̂, B, B̂, M = Mf∘Mpix, M̂ = Mf, L (src/dataset.jl:37-57).
    For pol = I / P the operators are DiagOps over real harmonic-basis fields with batch 1; for pol = IP (IEBFourier data) they
    are BlockDiagIEBs, as load_sim builds them (src/dataset.jl:262-301).  `Mpix` is a Map-basis DiagOp or None."""

    def __init__(self, d: Field, Cf: DiagOp, Cn: DiagOp, B: DiagOp, Mf: DiagOp, Mpix: DiagOp | None = None,
                 Cnhat: DiagOp | None = None, Bhat: DiagOp | None = None, L=LenseFlow, nsteps: int = 7,
                 D: DiagOp | None = None, G: DiagOp | None = None, Cϕ: DiagOp | None = None, Nϕ: DiagOp | None = None,
                 Cf̃: DiagOp | None = None):
        self.d, self.Cf, self.Cn, self.B, self.Mf, self.Mpix = HarmonicBasis(d), Cf, Cn, B, Mf, Mpix
        self.Cnhat, self.Bhat, self.L, self.nsteps = Cnhat or Cn, Bhat or B, L, nsteps
        self.D, self.G = D, G                      # mixing matrices of the Mixed parametrisation (src/dataset.jl:96-117); None = identity
        self.Cf̃ = Cf̃                              # lensed ("total") field covariance, used by quadratic_estimate (src/dataset.jl:270)
        self.Cϕ, self.Nϕ = Cϕ, Nϕ                  # ϕ prior and ϕ-noise estimate (logpdf, ϕ° Hessian preconditioner, src/dataset.jl:45-57,134-137)
        self._cg = {}

    def _solver(self, ϕ: Field):
        """(CG handle, CachedLenseFlow, LenseFlow, ϕ).  The CG handle (diagonals, CG vectors) is built once per dataset and
        reused for every ϕ: it references the pooled CachedLenseFlow, which is refilled in place for the ϕ at hand."""
        d = self.d
        p = d.proj
        L = self.L(ϕ, self.nsteps) if isinstance(self.L, type) else self.L
        cache = L.cache(d)                                                    # pooled per shape; precompute!! for L.ϕ
        # the handle holds raw device pointers to d and to the operators' planes: reuse it only while every one of them is the same
        # tensor (load_sim-style `ds.d = d` or a replaced operator builds a new handle); the key keeps the tensors alive
        ops = (self.Cf, self.Cn, self.Cnhat, self.B, self.Bhat, self.Mf, self.Mpix)
        key = (cache, d.arr) + tuple(None if D is None else D._real for D in ops)
        cur = self._cg.get("solver")
        old = self._cg.get("key")
        if cur is not None and old is not None and len(old) == len(key) and all(a is b for a, b in zip(old, key)):
            self._cg["solver"] = (cur[0], cache, L, ϕ)
            return self._cg["solver"]
        want = BlockDiagIEB if d.Npol == 3 else DiagOp
        for nm in ("Cf", "Cn", "Cnhat", "B", "Bhat", "Mf"):
            if not isinstance(getattr(self, nm), want):
                raise CmblError(f"BaseDataSet.{nm} must be a {want.__name__} for {d.basis} data")
        desc = DatasetDesc(d.Npol, d.Nbatch, *(c_void_p(D._real.data_ptr()) for D in (self.Cf, self.Cn, self.Cnhat, self.B, self.Bhat, self.Mf)),
                           c_void_p(self.Mpix._real.data_ptr()) if self.Mpix is not None else c_void_p(0), _ptr(d.arr))
        h = _CgHandle(p.lib)
        p.lib.call("cmbl_cg_create", byref(h.h), cache.handle, byref(desc), _stream(d.arr))
        self._cg["owner"] = h                                                  # previous owner (if any) is released here
        self._cg["key"] = key
        self._cg["solver"] = (h.h, cache, L, ϕ)
        return self._cg["solver"]


class _CgHandle:
    """Owns a cmbl_cg* (≈10 field-sized device buffers): destroyed with the dataset that created it."""
    def __init__(self, lib): self.lib, self.h = lib, c_void_p()
    def __del__(self):
        try:
            if self.h: self.lib.call("cmbl_cg_destroy", self.h)
        except Exception:
            pass


def mix(ds: BaseDataSet, f: Field, ϕ: Field):
    """mix(ds; f, ϕ) (src/dataset.jl:96-101): f° = L(ϕ)·D·f (Map basis, like the reference's L*f), ϕ° = G·ϕ."""
    L = ds.L(ϕ, ds.nsteps) if isinstance(ds.L, type) else ds.L
    Df = ds.D * f if ds.D is not None else f
    return L * Df, (ds.G * ϕ if ds.G is not None else ϕ)


def unmix(ds: BaseDataSet, f_mixed: Field, ϕ_mixed: Field):
    """unmix(ds; f°, ϕ°) (src/dataset.jl:111-116): ϕ = G \\ ϕ°, f = D \\ (L(ϕ) \\ f°)."""
    ϕ = ds.G.ldiv(ϕ_mixed) if ds.G is not None else ϕ_mixed
    L = ds.L(ϕ, ds.nsteps) if isinstance(ds.L, type) else LenseFlow(ϕ, ds.nsteps)
    f = L.ldiv(f_mixed)
    return (ds.D.ldiv(f) if ds.D is not None else f), ϕ


def gradientf_logpdf(ds: BaseDataSet, f: Field, ϕ: Field, d: Field | None = None, d_zero=False) -> Field:
    """gradientf_logpdf(ds; f, ϕ, d) (src/dataset.jl:76-80), returned in the harmonic basis."""
    h, *_ = ds._solver(ϕ)
    f = HarmonicBasis(f)
    out = torch.empty_like(f.arr)
    p = f.proj
    dh = HarmonicBasis(d) if d is not None else None          # keep the converted field alive across the call
    dptr = _ptr(dh.arr) if dh is not None else c_void_p(0)
    p.lib.call("cmbl_gradientf_logpdf", h, _ptr(f.arr), dptr, 1 if d_zero else 0, _ptr(out), _stream(out))
    return f._like(out)


def Hessian_logpdf_preconditioner(ds: BaseDataSet):
    """pinv(Cf) + B̂'M̂'pinv(Cn̂)M̂B̂ (src/dataset.jl:129-132); a DiagOp, or a BlockDiagIEB for pol = IP (operator algebra of
    src/specialops.jl:88,99-102 — setup-time host mirror; the solver forms the same block on the device in cg_setup)."""
    pinv = lambda t: torch.where(t == 0, torch.zeros_like(t), 1 / t)
    if isinstance(ds.Cf, BlockDiagIEB):
        full = lambda L: (L._real[0, 0], L._real[0, 1], L._real[0, 1], L._real[0, 2], L._real[0, 3])          # [a b; c d] ⊕ e
        def inv(m):
            a, _, c, d, e = m
            idet = pinv(a * d - c * c)
            return (d * idet, -(c * idet), -(c * idet), a * idet, pinv(e))
        mm = lambda x, y: (x[0] * y[0] + x[1] * y[2], x[0] * y[1] + x[1] * y[3], x[2] * y[0] + x[3] * y[2], x[2] * y[1] + x[3] * y[3], x[4] * y[4])
        bh, mf = full(ds.Bhat), full(ds.Mf)
        h, icf = mm(mm(mm(mm(bh, mf), inv(full(ds.Cnhat))), mf), bh), inv(full(ds.Cf))
        s = tuple(u + v for u, v in zip(icf, h))
        return BlockDiagIEB(s[0], s[2], s[3], s[4], proj=ds.Cf.proj)
    r = pinv(ds.Cf._real) + ds.Bhat._real * ds.Mf._real * pinv(ds.Cnhat._real) * ds.Mf._real * ds.Bhat._real
    return DiagOp(Field(ds.Cf.diag.basis, r.to(ds.Cf.diag.arr.dtype), ds.Cf.diag.proj))


class Comm:
    """cmbl_comm (include/cmbl_b200.h): the NCCL communicator of the C ABI for a batch sharded over the GPUs of a box — what a host
    program without torch.distributed (the Julia shim) uses.  Rank 0 creates the 128-byte id with `Comm.unique_id(lib)` and hands
    it to the other ranks by its own means; every rank then constructs `Comm(lib, nranks, rank, id)` on its device."""

    def __init__(self, lib, nranks: int, rank: int, uid: bytes):
        import ctypes
        self.lib, self.nranks, self.rank = lib, nranks, rank
        self.handle = c_void_p()
        self._id = ctypes.create_string_buffer(bytes(uid), 128)
        lib.call("cmbl_comm_init", byref(self.handle), nranks, rank, self._id)

    @staticmethod
    def unique_id(lib) -> bytes:
        import ctypes
        buf = ctypes.create_string_buffer(128)
        lib.call("cmbl_comm_unique_id", buf)
        return buf.raw

    def allreduce(self, values, op: str = "sum", stream=None) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        arr = (c_double * len(v))(*v)
        self.lib.call("cmbl_comm_allreduce", self.handle, arr, len(v), {"sum": 0, "min": 1, "max": 2}[op], stream or c_void_p(0))
        return np.array(arr[:])

    def close(self):
        if self.handle:
            self.lib.call("cmbl_comm_destroy", self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def conjugate_gradient_wiener(ds: BaseDataSet, ϕ: Field, fstart: Field | None = None, nsteps=500, tol=1e-1, offset=False,
                              group=None, comm: Comm | None = None):
    """conjugate_gradient (src/numerical_algorithms.jl:73-134) specialised to the Wiener-filter Hessian.  Returns
    (bestx, history) with history = [(i, res[Nb]), ...].  With a torch.distributed `group` the batch is sharded over ranks and
    the lock-step rules `all(res<bestres)` / `all(res<tol)` are taken across ranks (one tiny all-reduce per iteration)."""
    h, cache, L, _ = ds._solver(ϕ)
    p = ds.d.proj
    nb = ds.d.Nbatch
    st = _stream(ds.d.arr)
    res = (c_double * nb)()
    fsf = HarmonicBasis(fstart) if fstart is not None else None   # keep the converted field alive while the solver reads it
    fs = _ptr(fsf.arr) if fsf is not None else c_void_p(0)
    out = torch.empty_like(ds.d.arr)
    if group is None:
        hist = (c_double * (nsteps * nb))()
        iters = c_int(0)
        if comm is not None:                      # batch sharded over the ranks of a C-ABI communicator: the whole loop stays in the library
            p.lib.call("cmbl_wiener_cg_sharded", h, comm.handle, fs, _ptr(out), nsteps, c_double(tol), 1 if offset else 0, byref(iters), hist, st)
        else:
            p.lib.call("cmbl_wiener_cg", h, fs, _ptr(out), nsteps, c_double(tol), 1 if offset else 0, byref(iters), hist, st)
        H = np.array(hist[: iters.value * nb]).reshape(iters.value, nb)
        return ds.d._like(out), [(i + 1, H[i]) for i in range(iters.value)]
    import torch.distributed as dist

    def all_true(flag: bool) -> bool:
        t = torch.tensor([0 if flag else 1], dtype=torch.int32, device=ds.d.arr.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return int(t.item()) == 0

    p.lib.call("cmbl_cg_begin", h, fs, 1 if offset else 0, res, st)
    r = np.array(res[:]); best = r.copy(); hist = [(1, r.copy())]
    for i in range(2, nsteps + 1):
        p.lib.call("cmbl_cg_step", h, res, st)
        r = np.array(res[:])
        if all_true(bool(np.all(r < best))):
            best = r.copy()
            p.lib.call("cmbl_cg_mark_best", h, st)
        hist.append((i, r.copy()))
        if all_true(bool(np.all(r < tol))):
            break
    p.lib.call("cmbl_cg_result", h, 0, _ptr(out), st)
    return ds.d._like(out), hist


def argmaxf_logpdf(ds: BaseDataSet, ϕ: Field, fstart: Field | None = None, offset=False,
                   conjgrad_kwargs=dict(tol=1e-1, nsteps=500), group=None, comm=None):
    """argmaxf_logpdf(ds, (;ϕ)) (src/maximization.jl:17-42): the Wiener filter of ds.d at fixed ϕ; returns (f, history).  `group` (a
    torch.distributed group) or `comm` (a C-ABI `Comm`) shard the batch over ranks with the reference's lock-step stopping rule."""
    return conjugate_gradient_wiener(ds, ϕ, fstart=fstart, offset=offset, group=group, comm=comm, **conjgrad_kwargs)


argmaxf_lnP = argmaxf_logpdf          # the name BASELINE.json uses (pre-0.10 spelling)


def _sqrt_mul(op, w: Field) -> Field:
    """sqrt(C) * w for a DiagOp or BlockDiagIEB covariance (simulate, src/specialops.jl:6,94)."""
    if isinstance(op, BlockDiagIEB):
        return op.sqrt_mul(w)
    return DiagOp(Field(op.diag.basis, torch.sqrt(op._real).to(op.diag.arr.dtype), op.diag.proj)) * w


def _apply_MB(ds: BaseDataSet, ft: Field) -> Field:
    """M * (B * f̃) with M = Mfourier * Mpix (src/dataset.jl:60-67,283-290)."""
    x = ds.B * HarmonicBasis(ft)
    if ds.Mpix is not None:
        x = ds.Mpix * LenseBasis(x)
    return ds.Mf * HarmonicBasis(x)


def simulate(ds: BaseDataSet, ϕ: Field, white_f: Field | None = None, white_n: Field | None = None, generator=None):
    """simulate(rng, ds; ϕ) of the BaseDataSet forward model (src/dataset.jl:60-67): f ~ N(0, Cf), d ~ N(M B L(ϕ) f, Cn).
    The unit white-noise maps may be passed in (LenseBasis fields) so that a CPU restatement can be fed the same draws."""
    d0 = ds.d
    p = d0.proj
    lense = ("Map", "QUMap", "IQUMap")[d0.Npol - 1]
    draw = lambda: Field(lense, torch.randn(p.map_shape(d0.Npol, d0.Nbatch), dtype=p.T, device=p.device, generator=generator), p)
    wf = white_f if white_f is not None else draw()
    wn = white_n if white_n is not None else draw()
    f = _sqrt_mul(ds.Cf, HarmonicBasis(wf))
    n = _sqrt_mul(ds.Cn, HarmonicBasis(wn))
    L = ds.L(ϕ, ds.nsteps) if isinstance(ds.L, type) else ds.L
    f̃ = L * f
    d = _apply_MB(ds, f̃) + n
    return {"f": f, "f̃": f̃, "ϕ": ϕ, "d": d}


def sample_f(ds: BaseDataSet, ϕ: Field, white_f: Field | None = None, white_n: Field | None = None, generator=None,
             conjgrad_kwargs=dict(tol=1e-1, nsteps=500), fstart: Field | None = None):
    """sample_f(rng, ds, (;ϕ)) (src/maximization.jl:56-62): a posterior sample of f at fixed ϕ,
    sim.f + argmaxf_logpdf(ds, Ω, d − sim.d; offset=true).  Returns (f, history)."""
    sim = simulate(ds, ϕ, white_f, white_n, generator)
    ds2 = BaseDataSet(ds.d - sim["d"], ds.Cf, ds.Cn, ds.B, ds.Mf, ds.Mpix, ds.Cnhat, ds.Bhat, L=ds.L, nsteps=ds.nsteps, D=ds.D, G=ds.G)
    Δf, hist = argmaxf_logpdf(ds2, ϕ, fstart=fstart, offset=True, conjgrad_kwargs=conjgrad_kwargs)
    return sim["f"] + Δf, hist


# ------------------------------------------------------------------------------------------------------------------
# Joint posterior: logpdf, Mixed(ds), its gradient, MAP_joint (src/dataset.jl:60-67,84-117, src/distributions.jl:11-15,
# src/maximization.jl:115-222).  Control flow lives here on the host, as it does in the reference; every field operation
# is a call into the library (flows, δ-flows, transforms, diagonal products, dots).
# ------------------------------------------------------------------------------------------------------------------
def logdet(op) -> float:
    """logdet(Diagonal(::Fourier-basis field)) = Σ nan2zero(log|diag|)·λ_rfft (src/proj_lambert.jl:331-336);
    logdet(BlockDiagIEB) = logdet(det ΣTE) + logdet ΣB (src/specialops.jl:94).  A setup-time constant of the operator."""
    p = op.proj if isinstance(op, BlockDiagIEB) else op.diag.proj
    λ = torch.from_numpy(p.λ_rfft).to(op._real.device, torch.float64)
    def ld(t):
        v = torch.log(t.abs().to(torch.float64)) * λ
        return float(torch.where(torch.isfinite(v), v, torch.zeros_like(v)).sum())
    r = op._real
    if isinstance(op, BlockDiagIEB):
        return ld(r[0, 0] * r[0, 2] - r[0, 1] * r[0, 1]) + ld(r[0, 3])
    if not op.diag.is_fourier:
        raise CmblError("logdet on this path is implemented for Fourier-basis diagonals and BlockDiagIEB")
    return ld(r)


def _quad(C, v: Field) -> np.ndarray:
    """v' pinv(Σ) v + logdet Σ per batch item (src/distributions.jl:11-15)."""
    return dot(v, C.ldiv(v)) + logdet(C)


def logpdf(ds, f: Field | None = None, ϕ: Field | None = None, **kw) -> np.ndarray:
    """logpdf(ds; f, ϕ) of the BaseDataSet forward model, or logpdf(Mixed(ds); f°=…, ϕ°=…) (src/dataset.jl:60-67,84-87).
    Returns one value per batch item."""
    if isinstance(ds, Mixed):
        f, ϕ = unmix(ds.ds, kw.get("f°", f), kw.get("ϕ°", ϕ))
        return logpdf(ds.ds, f, ϕ)                       # − logdet(D,θ) − logdet(G,θ), both 0 without θ dependence (src/generic.jl:269)
    if ds.Cϕ is None:
        raise CmblError("logpdf needs BaseDataSet(..., Cϕ=...)")
    L = ds.L(ϕ, ds.nsteps) if isinstance(ds.L, type) else LenseFlow(ϕ, ds.nsteps)
    f = HarmonicBasis(f)
    z = _apply_MB(ds, L * f) - ds.d
    return -(_quad(ds.Cn, z) + _quad(ds.Cf, f) + _quad(ds.Cϕ, Fourier(ϕ))) / 2


def mixing_D(ds: BaseDataSet, σ_len_arcmin: float = 5.0) -> DiagOp:
    """The mixing matrix load_sim attaches to a dataset: D = sqrt((Cf + (σ²len + 2Cn̂)) · pinv(Cf)), σ²len = deg2rad(5/60)²
    (src/dataset.jl:325-332).  It decorrelates f° from ϕ°, which is what makes pinv(Cϕ)+pinv(Nϕ) a usable ϕ° Hessian in MAP_joint
    (with D = 1 the line search collapses to α ~ 1e-4)."""
    σ2 = float(np.deg2rad(σ_len_arcmin / 60.0) ** 2)
    if isinstance(ds.Cf, BlockDiagIEB):
        # BlockDiagIEB algebra (src/specialops.jl:99-105: UniformScaling adds to the diagonal entries and to B, `*` is the 2×2 matrix
        # product) and the 2×2 sqrt of src/field_vectors.jl:62-67, which reads the off-diagonal of the product from [2,1]
        pinv = lambda t: torch.where(t == 0, torch.zeros_like(t), 1 / t)
        cf, cn = ds.Cf._real[0], ds.Cnhat._real[0]
        idet = pinv(cf[0] * cf[2] - cf[1] * cf[1])
        pa, pc, pd, pe = cf[2] * idet, -(cf[1] * idet), cf[0] * idet, pinv(cf[3])
        a, c, d, e = cf[0] + (σ2 + 2 * cn[0]), cf[1] + 2 * cn[1], cf[2] + (σ2 + 2 * cn[2]), cf[3] + (σ2 + 2 * cn[3])
        A, C, D_, E = a * pa + c * pc, c * pa + d * pc, c * pc + d * pd, e * pe                   # [1,1], [2,1], [2,2], B of the product
        s_ = torch.sqrt(A * D_ - C * C)
        t_ = pinv(torch.sqrt(A + (D_ + 2 * s_)))
        return BlockDiagIEB(t_ * (A + s_), t_ * C, t_ * (D_ + s_), torch.sqrt(E), proj=ds.Cf.proj)
    cf, cn = ds.Cf._real, ds.Cnhat._real
    r = torch.sqrt((cf + (σ2 + 2 * cn)) * torch.where(cf == 0, torch.zeros_like(cf), 1 / cf))
    return DiagOp(Field(ds.Cf.diag.basis, r.to(ds.Cf.diag.arr.dtype), ds.Cf.diag.proj))


class Mixed:
    """Mixed(ds): the same posterior in the mixed variables f° = L(ϕ) D f, ϕ° = G ϕ (src/dataset.jl:28-30,84-117)."""
    def __init__(self, ds: BaseDataSet): self.ds = ds


def gradient_logpdf_mixed(ds: BaseDataSet, f_mixed: Field, ϕ_mixed: Field, bug_compat: bool = True):
    """gradient((f°, ϕ°) -> logpdf(Mixed(ds); f°, ϕ°)) (src/maximization.jl:151) through the reference's pullbacks: the chain
    f° → f₁ = L(ϕ)\\f° → f = D\\f₁ → f̃ = L(ϕ) f → r = d − M B f̃ is walked backwards with the transpose δ-flows of `L*f` and `L\\f`
    (negδvelocityᴴ, src/lenseflow.jl:176-214; rules src/flowops.jl:40-68).  Two flows and two δ-flows on the device.
    Returns (∇f° in the Ð basis, ∇ϕ° Fourier):  d lnP = ⟨∇f°, δf°⟩ + ⟨∇ϕ°, δϕ°⟩."""
    if ds.Cϕ is None:
        raise CmblError("gradient_logpdf_mixed needs BaseDataSet(..., Cϕ=...)")
    ϕ = ds.G.ldiv(ϕ_mixed) if ds.G is not None else Fourier(ϕ_mixed)
    f_mixed = LenseBasis(f_mixed)
    Lc = LenseFlow(ϕ, ds.nsteps).cache(f_mixed, with_minv=True)
    f1 = Lc.apply(OP_LINV, f_mixed)
    f = ds.D.ldiv(f1) if ds.D is not None else HarmonicBasis(f1)
    f̃ = Lc.apply(OP_L, f)
    r = ds.d - _apply_MB(ds, f̃)
    x = ds.Mf * ds.Cn.ldiv(r)                                                   # M' = Mpix'·Mf' after pinv(Cn)
    if ds.Mpix is not None:
        x = HarmonicBasis(ds.Mpix * LenseBasis(x))
    g_f̃ = ds.B * x                                                              # ∂lnP/∂f̃ = B'M'pinv(Cn) r
    δf_a, δϕ_a = Lc.pullback(OP_L, f̃, g_f̃, bug_compat)
    g_f = HarmonicBasis(δf_a) - ds.Cf.ldiv(f)
    g_f1 = ds.D.ldiv(g_f) if ds.D is not None else g_f                           # D is real and diagonal: D⁻ᵀ = D⁻¹
    δf0, δϕ_b = Lc.pullback(OP_LINV, f1, g_f1, bug_compat)
    g_ϕ = δϕ_a + δϕ_b - ds.Cϕ.ldiv(ϕ)
    if ds.G is not None:
        g_ϕ = ds.G.ldiv(g_ϕ)
    return δf0, g_ϕ


def _brent_bounded(fun, a: float, b: float, xatol: float, maxiter: int = 500):
    """Brent's derivative-free minimiser on [a, b] (golden section + successive parabolic interpolation) — the algorithm
    behind Optim.Brent() that MAP_joint's line search calls (src/maximization.jl:171-176).  Returns (x, f(x), evaluations)."""
    golden = 0.5 * (3.0 - 5.0 ** 0.5)
    sqrt_eps = float(np.sqrt(2.2e-16))
    x = w = v = a + golden * (b - a)
    fx = fw = fv = fun(x)
    d = e = 0.0
    n = 1
    while n < maxiter:
        m = 0.5 * (a + b)
        tol1 = sqrt_eps * abs(x) + xatol / 3.0
        tol2 = 2.0 * tol1
        if abs(x - m) <= tol2 - 0.5 * (b - a):
            break
        use_golden = True
        if abs(e) > tol1:                                   # parabolic fit through (v, w, x)
            r_ = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            p_ = (x - v) * q - (x - w) * r_
            q = 2.0 * (q - r_)
            if q > 0:
                p_ = -p_
            q = abs(q)
            r_, e = e, d
            if abs(p_) < abs(0.5 * q * r_) and p_ > q * (a - x) and p_ < q * (b - x):
                d = p_ / q
                u = x + d
                if (u - a) < tol2 or (b - u) < tol2:
                    d = tol1 if m >= x else -tol1
                use_golden = False
        if use_golden:
            e = (b - x) if x < m else (a - x)
            d = golden * e
        u = x + (d if abs(d) >= tol1 else (tol1 if d > 0 else -tol1))
        fu = fun(u); n += 1
        if fu <= fx:
            if u >= x: a = x
            else: b = x
            v, fv, w, fw, x, fx = w, fw, x, fx, u, fu
        else:
            if u < x: a = u
            else: b = u
            if fu <= fw or w == x:
                v, fv, w, fw = w, fw, u, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
    return x, fx, n


def MAP_joint(ds: BaseDataSet, ϕstart: Field | None = None, nsteps: int = 20, fstart: Field | None = None, αtol: float = 1e-4,
              αmax: float | None = None, conjgrad_kwargs=dict(tol=1e-1, nsteps=500), bug_compat: bool = True, group=None):
    """MAP_joint(ds; nsteps, αtol, conjgrad_kwargs) (src/maximization.jl:115-222) for Ω = (ϕ,): coordinate descent that alternates
    the CG Wiener filter at fixed ϕ (`argmaxf_logpdf`) with one step along pinv(H)·∇ϕ° logpdf(Mixed(ds)), H = pinv(Cϕ) + pinv(Nϕ)
    (src/dataset.jl:134-137), whose length comes from a Brent line search on [0, 2α] of the batch-summed −logpdf(Mixed(ds)).
    G = 1 during the maximisation, as in the reference (:137).  With a torch.distributed `group` the batch is sharded over ranks:
    the line-search objective is all-reduced (one scalar per evaluation), so every rank takes the same α.
    Returns (f, ϕ, history)."""
    if ds.Cϕ is None or ds.Nϕ is None:
        raise CmblError("MAP_joint needs BaseDataSet(..., Cϕ=..., Nϕ=...)")
    d = ds.d
    p = d.proj
    G_save, ds.G = ds.G, None
    ϕ = Fourier(ϕstart) if ϕstart is not None else Field("Fourier", torch.zeros(p.fourier_shape(1, d.Nbatch), dtype=p.cT, device=p.device), p)
    H = DiagOp(Field("Fourier", (ds.Cϕ.pinv()._real + ds.Nϕ.pinv()._real).to(p.cT), p))
    f, α, history = fstart, 1.0, []

    def total(v: np.ndarray) -> float:
        s = float(np.sum(v))
        if group is not None:
            import torch.distributed as dist
            t = torch.tensor([s], dtype=torch.float64, device=d.arr.device if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(t, group=group)
            s = float(t.item())
        return s

    try:
        for step in range(1, nsteps + 1):
            f, cg_hist = argmaxf_logpdf(ds, ϕ, fstart=f, conjgrad_kwargs=conjgrad_kwargs, group=group)      # f step
            f_m, ϕ_m = mix(ds, f, ϕ)                                                                          # ϕ step
            _, g = gradient_logpdf_mixed(ds, f_m, ϕ_m, bug_compat)
            Δ = H.ldiv(g)
            amax = αmax if αmax is not None else 2 * α
            mds = Mixed(ds)

            def obj(a):
                v = -total(logpdf(mds, f_m, ϕ_m + Δ * float(a)))
                return v if np.isfinite(v) else (a / amax) * np.finfo(np.float64).max                           # :174
            α, _, nev = _brent_bounded(obj, 0.0, float(amax), αtol)
            ϕ_m = ϕ_m + Δ * float(α)
            lp = logpdf(mds, f_m, ϕ_m)
            # :206 deletes :f from unmix(...): f stays the CG solution (the next step's fstart, :230, and the value returned, :224),
            # so only the ϕ half of unmix is evaluated (G = 1 here, :137)
            ϕ = ds.G.ldiv(ϕ_m) if ds.G is not None else ϕ_m
            history.append(dict(step=step, logpdf=lp, α=α, cg_iters=len(cg_hist), linesearch_evals=nev))
    finally:
        ds.G = G_save
    return f, ϕ, history


# ------------------------------------------------------------------------------------------------------------------
# HMC step in ϕ° of the Gibbs sampler `sample_joint` (src/sampling.jl:14-55,397-425).  Host control flow; every leap-frog step
# costs one gradient of logpdf(Mixed(ds)) = two flows + two δ-flows on the device.
# ------------------------------------------------------------------------------------------------------------------
def symplectic_integrate(x0: Field, p0: Field, Λ: DiagOp, U, δUδx, N: int = 50, ϵ: float = 0.1):
    """symplectic_integrate(x₀, p₀, Λ, U, δUδx; N, ϵ) (src/sampling.jl:14-55): leap-frog on the log-density U with mass matrix Λ,
    H = U − p·(Λ\\p)/2.  Returns (ΔH per batch item, x, p)."""
    H = lambda x, p: U(x) - dot(p, Λ.ldiv(p)) / 2
    x, p = x0, p0
    g = δUδx(x)
    for _ in range(N):
        x1 = x - Λ.ldiv(p - g * (ϵ / 2)) * ϵ
        g1 = δUδx(x1)
        p = p - (g1 + g) * (ϵ / 2)
        x, g = x1, g1
    return H(x, p) - H(x0, p0), x, p


def mass_matrix_ϕ(ds: BaseDataSet) -> DiagOp:
    """mass_matrix_ϕ(θ, ds) = pinv(G)² (pinv(Cϕ) + pinv(Nϕ)) (src/sampling.jl:422-425)."""
    if ds.Cϕ is None or ds.Nϕ is None:
        raise CmblError("mass_matrix_ϕ needs BaseDataSet(..., Cϕ=..., Nϕ=...)")
    m = ds.Cϕ.pinv()._real + ds.Nϕ.pinv()._real
    if ds.G is not None:
        m = ds.G.pinv()._real ** 2 * m
    p = ds.d.proj
    return DiagOp(Field("Fourier", m.to(p.cT), p))


def hmc_step(U, x: Field, Λ: DiagOp, δUδx, symp_kwargs=(dict(N=25, ϵ=0.01),), always_accept=False, white: Field | None = None,
             uniforms=None, generator=None):
    """hmc_step (src/sampling.jl:405-418): draw p ~ N(0, Λ), integrate, accept per batch item with probability min(1, e^ΔH).
    `white` (unit white Map field) and `uniforms` (one U(0,1) per batch item) may be passed in for reproducibility."""
    pr = x.proj
    ΔH = accept = None
    whites = list(white) if isinstance(white, (list, tuple)) else [white] * len(symp_kwargs)       # one momentum draw per entry (src/sampling.jl:408)
    unis = list(uniforms) if (isinstance(uniforms, (list, tuple)) and len(uniforms) and np.ndim(uniforms[0]) > 0) else [uniforms] * len(symp_kwargs)
    if len(whites) != len(symp_kwargs) or len(unis) != len(symp_kwargs):
        raise CmblError("hmc_step: pass one `white` / `uniforms` per symp_kwargs entry (or a single one / None)")
    for kw, white, uniforms in zip(symp_kwargs, whites, unis):
        w = white if white is not None else Field("Map", torch.randn(pr.map_shape(1, x.Nbatch), dtype=pr.T, device=pr.device, generator=generator), pr)
        p0 = DiagOp(Field("Fourier", torch.sqrt(Λ._real).to(pr.cT), pr)) * Fourier(w)          # simulate(rng, Λ)
        ΔH, xtest, _ = symplectic_integrate(x, p0, Λ, U, δUδx, **kw)
        u = np.asarray(uniforms, dtype=np.float64) if uniforms is not None else \
            torch.rand(x.Nbatch, dtype=torch.float64, device=pr.device if generator is not None and generator.device.type != "cpu" else "cpu", generator=generator).cpu().numpy()
        accept = np.logical_or(always_accept, np.log(u) < ΔH)
        a = accept.astype(np.float64)
        x = xtest * a + x * (1 - a)
    return x, ΔH, accept


def gibbs_sample_ϕ(ds: BaseDataSet, f_mixed: Field, ϕ_mixed: Field, symp_kwargs=(dict(N=25, ϵ=0.01),), always_accept=False,
                   white: Field | None = None, uniforms=None, bug_compat: bool = True, generator=None):
    """gibbs_sample_ϕ! (src/sampling.jl:397-403): one HMC update of ϕ° at fixed f° under logpdf(Mixed(ds)).  `generator` seeds both the
    momentum draw and the accept/reject uniforms when `white` / `uniforms` are not supplied."""
    mds = Mixed(ds)
    U = lambda x: logpdf(mds, f_mixed, x)
    δU = lambda x: gradient_logpdf_mixed(ds, f_mixed, x, bug_compat)[1]
    return hmc_step(U, Fourier(ϕ_mixed), mass_matrix_ϕ(ds), δU, symp_kwargs, always_accept, white, uniforms, generator)


# ------------------------------------------------------------------------------------------------------------------
# Quadratic estimate of ϕ and its analytic N⁰ (src/quadratic_estimate.jl:30-199): products of inverse-variance-filtered "legs"
# in map space; every leg is one cmbl_irfft2, every product goes back through cmbl_rfft2.  A setup-time computation
# (load_sim uses it once for Nϕ, src/dataset.jl:316), written with the host mirror's broadcast arithmetic like the reference's.
# ------------------------------------------------------------------------------------------------------------------
def quadratic_estimate(ds: BaseDataSet, which: str | None = None, wiener_filtered: bool = True, weights: str = "unlensed", AL: DiagOp | None = None,
                       abs_each_term: bool = True):
    """quadratic_estimate(ds, which; wiener_filtered, weights, AL) for which ∈ {TT, EE, EB}: returns dict(ϕqe, AL, Nϕ), Nϕ = AL.
    `abs_each_term=True` is the reference's normalisation pinv(Σ_ij abs.(∇ᵢ∇ⱼ·Fourier(A(i,j)))) (:117,151,190), which under-normalises
    EB by ≈40 % because the cross terms are not sign-definite; `False` takes |Σ_ij …| (unit response)."""
    from itertools import product
    if weights not in ("lensed", "unlensed"):
        raise CmblError("weights should be lensed or unlensed")
    d = ds.d
    p = d.proj
    which = which or ("TT" if d.Npol == 1 else "EB")
    if which not in ("TT", "EE", "EB"):
        raise CmblError(f"which='{which}' not implemented")
    if ds.Cf̃ is None or ds.Cϕ is None:
        raise CmblError("quadratic_estimate needs BaseDataSet(..., Cf̃=..., Cϕ=...)")
    if d.Npol == 3:
        # ds.d[pol], Cf[pol], ... (src/quadratic_estimate.jl:44): the I or the P part of an IQU dataset;
        # BlockDiagIEB[:P] = Diagonal(EBFourier(E, B)), [:I] = ΣTT (src/specialops.jl:107-113)
        sl_op = (lambda L: L._real[:, 0:1]) if which == "TT" else (lambda L: L._real[:, 2:4])
        basis = "Fourier" if which == "TT" else "EBFourier"
        D = lambda L: DiagOp(Field(basis, sl_op(L).contiguous().to(p.cT), p))
        dsub = Field(basis, (d.arr[:, 0:1] if which == "TT" else d.arr[:, 1:3]).contiguous(), p)
        sub = BaseDataSet(dsub, D(ds.Cf), D(ds.Cn), D(ds.B), D(ds.Mf), None, D(ds.Cnhat), D(ds.Bhat), L=ds.L, nsteps=ds.nsteps, Cϕ=ds.Cϕ, Cf̃=D(ds.Cf̃))
        return quadratic_estimate(sub, which, wiener_filtered, weights, AL, abs_each_term)
    if d.Npol != (1 if which == "TT" else 2):
        raise CmblError(f"which='{which}' not implemented for {d.basis} data")
    dev, cT = p.device, p.cT
    lx = torch.from_numpy(p.ℓx).to(dev)[:, None]; ly = torch.from_numpy(p.ℓy).to(dev)[None, :]
    grad = {1: (1j * lx).to(cT).expand(p.Nx, p.Nyh), 2: (1j * ly).to(cT).expand(p.Nx, p.Nyh)}
    lmag = torch.from_numpy(p.ℓmag).to(dev)
    nz = lambda t: torch.where(torch.isfinite(t.real) & (torch.isfinite(t.imag) if t.is_complex() else True), t, torch.zeros_like(t))
    irf = lambda F: Map(Field("Fourier", F.to(cT), p)).arr                    # cmbl_irfft2
    fou = lambda m: Fourier(Field("Map", m, p)).arr                           # cmbl_rfft2
    memo = {}

    def leg(C, *inds):                                                        # QE_leg (:84-93), memoised on (C, n, p₁, p₂)
        n = sum(1 for x in inds if isinstance(x, int))
        first = [x if isinstance(x, int) else x[0] for x in inds]
        key = (id(C), n, first.count(1), first.count(2))
        if key not in memo:
            memo[key] = (irf(nz(C * grad[1] ** key[2] * grad[2] ** key[3] / lmag ** n)), C)
        return memo[key][0]

    eps = lambda a, b: 0 if a == b else (1 if (a, b) == (1, 2) else -1)      # levicivita([a, b, 3])
    inds = lambda D: list(product((1, 2), repeat=D))
    pinv = lambda t: torch.where(t == 0, torch.zeros_like(t), 1 / t)
    ldiv = lambda S, x: nz(x / S)
    sl = lambda A, c: A[:, c:c + 1]
    TF = ds.Mf._real * ds.Bhat._real
    Cf, Cft, Cn = ds.Cf._real, ds.Cf̃._real, ds.Cnhat._real
    Cw = Cf if weights == "unlensed" else Cft

    def norm(A):
        terms = [grad[i] * grad[j] * fou(A(i, j)) for i, j in inds(2)]
        tot = sum(t.abs() for t in terms) if abs_each_term else sum(terms).abs()
        return pinv(tot.to(p.T))

    if which == "TT":
        S = TF ** 2 * Cft + Cn
        a, b = ldiv(S, TF * d.arr), Cw * ldiv(S, TF * d.arr)
        unnorm = -sum(grad[i] * fou(leg(a) * leg(b, [i])) for i in (1, 2))
        if AL is None:
            X2, X1, X0 = TF ** 2 * Cw ** 2 / S, TF ** 2 * Cw / S, TF ** 2 / S
            ALr = norm(lambda i, j: leg(X2, [i], [j]) * leg(X0) + leg(X1, [i]) * leg(X1, [j]))
    else:
        TF2E, TF2B = sl(TF, 0) ** 2, sl(TF, 1) ** 2
        SE, SB = TF2E * sl(Cft, 0) + sl(Cn, 0), TF2B * sl(Cft, 1) + sl(Cn, 1)
        CE, CB = sl(Cw, 0), sl(Cw, 1)
        tE, tB = sl(TF * d.arr, 0), sl(TF * d.arr, 1)
        if which == "EE":
            a1, b2 = CE * ldiv(SE, tE), ldiv(SE, tE)
            I = lambda i: -(2 * sum(leg(a1, [i], j, k) * leg(b2, j, k) for j, k in inds(2)) - leg(a1, [i]) * leg(b2))
            unnorm = sum(grad[i] * fou(I(i)) for i in (1, 2))
            if AL is None:
                X2, X1, X0 = TF2E * CE ** 2 / SE, TF2E * CE / SE, TF2E / SE
                A1 = lambda i, j: -4 * sum(eps(m, q_) * eps(n, r_) * (leg(X2, [i], [j], k, l, m, n) * leg(X0, k, l, q_, r_)
                                                                        + leg(X1, [i], k, l, m, n) * leg(X1, [j], k, l, q_, r_))
                                           for k, l, m, n, q_, r_ in inds(6) if eps(m, q_) and eps(n, r_))
                A2 = lambda i, j: leg(X2, [i], [j]) * leg(X0) + leg(X1, [i]) * leg(X1, [j])
                ALr = norm(lambda i, j: A1(i, j) + A2(i, j))
        else:
            aE, aE0 = CE * ldiv(SE, tE), ldiv(SE, tE)
            bB, bB0 = CB * ldiv(SB, tB), ldiv(SB, tB)
            I = lambda i: 2 * sum(eps(k, l) * (leg(aE, [i], j, k) * leg(bB0, j, l) - leg(aE0, j, k) * leg(bB, [i], j, l))
                                  for j, k, l in inds(3) if eps(k, l))
            unnorm = sum(grad[i] * fou(I(i)) for i in (1, 2))
            if AL is None:
                XE2, XE1, XE0 = TF2E * CE ** 2 / SE, TF2E * CE / SE, TF2E / SE
                XB2, XB1, XB0 = TF2B * CB ** 2 / SB, TF2B * CB / SB, TF2B / SB
                A = lambda i, j: 4 * sum(eps(m, q_) * eps(n, r_) * (leg(XE2, [i], [j], k, l, m, n) * leg(XB0, k, l, q_, r_)
                                                                    - 2 * leg(XE1, [i], k, l, m, n) * leg(XB1, [j], k, l, q_, r_)
                                                                    + leg(XE0, k, l, m, n) * leg(XB2, [i], [j], k, l, q_, r_))
                                         for k, l, m, n, q_, r_ in inds(6) if eps(m, q_) and eps(n, r_))
                ALr = norm(A)
    if AL is not None:
        ALr = AL._real
    ϕ = ALr * unnorm
    if wiener_filtered:
        Cp = ds.Cϕ._real
        ϕ = Cp * pinv(Cp + ALr) * ϕ
    ALop = DiagOp(Field("Fourier", ALr.to(cT), p))
    return {"ϕqe": Field("Fourier", ϕ.to(cT), p), "AL": ALop, "Nϕ": ALop}      # string keys: identifiers would be NFKC-normalised (ϕ → φ)


def sample_joint(ds: BaseDataSet, ϕstart: Field, nsamps_per_chain: int | None = None, symp_kwargs=(dict(N=25, ϵ=0.01),), nburnin_always_accept: int = 10,
                 conjgrad_kwargs=dict(tol=1e-1, nsteps=500), draws=None, generator=None, bug_compat: bool = True):
    """sample_joint (src/sampling.jl:180-336) for (f, ϕ) at fixed θ: the Gibbs passes gibbs_sample_f! (sample_f), gibbs_mix!,
    gibbs_sample_ϕ! (one HMC update of ϕ° per entry of symp_kwargs), gibbs_unmix! (:388-451).  The batch dimension carries independent
    chains (the reference's `Nbatch` chains per worker); steps are numbered from 2 and proposals are always accepted while
    step < nburnin_always_accept, like the reference.  `draws` (optional, one dict(wf, wn, wp, u) per step) fixes the random numbers.
    Returns the chain as a list of dicts (f, ϕ, ΔH, accept, logpdf)."""
    if draws is None and nsamps_per_chain is None:
        raise CmblError("sample_joint: pass nsamps_per_chain (or explicit `draws`)")
    nsteps = len(draws) if draws is not None else (nsamps_per_chain - 1)
    ϕ, chain = Fourier(ϕstart), []
    for i in range(nsteps):
        step = i + 2
        dr = draws[i] if draws is not None else {}
        f, _ = sample_f(ds, ϕ, dr.get("wf"), dr.get("wn"), generator, conjgrad_kwargs)
        f_m, ϕ_m = mix(ds, f, ϕ)
        ϕ_m, ΔH, acc = gibbs_sample_ϕ(ds, f_m, ϕ_m, symp_kwargs, step < nburnin_always_accept, dr.get("wp"), dr.get("u"), bug_compat, generator)
        f, ϕ = unmix(ds, f_m, ϕ_m)
        chain.append({"step": step, "f": f, "ϕ": ϕ, "ΔH": ΔH, "accept": acc, "logpdf": logpdf(ds, f, ϕ)})      # string keys (no NFKC folding)
    return chain

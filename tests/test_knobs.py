"""The run-time knobs that switch between kernel families give the same answers: each variant runs in its own process
(the library reads its environment once) on the emulator (CPU suite) or on cuda:0 (-m gpu), and the digests are compared."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import hashlib, json, os, sys
import numpy as np, torch
ROOT = sys.argv[1]; backend = sys.argv[2]
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
lib = pkg._lib.Library(os.path.join(ROOT, "tests", "_emu", "libcmbl_emu.so")) if backend == "emu" else None
dev = "cpu" if backend == "emu" else "cuda:0"
out = {}
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_problem                                             # the seeded ΛCDM problem of the kernel tests (oracle = input generator only)
for dt in ("f64", "f32"):
    pr = make_problem(pkg, 256, 512, "P", dt, nb=2, nsteps=3, mask=False, seed=11, lib=lib, device=dev)   # Ny = 256, Nx = 512: fast stage kernels, fast transform columns
    f = pkg.LenseBasis(pr["f"])
    L = pkg.LenseFlow(pr["phi"], 3)
    F = pkg.QUFourier(f)
    res = {"rfft2": F.arr, "irfft2": pkg.QUMap(F).arr, "L": (L * f).arr, "LH": (L.H * F).arr, "LHinv": L.H.ldiv(F).arr}
    cache = L.cache(f, with_minv=True)
    δf, δϕ = cache.pullback(pkg.OP_L, cache.apply(pkg.OP_L, f), F)
    res["δf"], res["δϕ"] = δf.arr, δϕ.arr
    for kname, v in res.items():
        a = v.detach().cpu().numpy()
        out[dt + ":" + kname] = [hashlib.sha1(a.tobytes()).hexdigest(), float(np.abs(a).sum())]
print(json.dumps(out))
'''


def _run(backend, env):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CODE, ROOT, backend], capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("backend", ["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_kernel_family_knobs_agree(backend, request):
    if backend == "emu":
        request.getfixturevalue("emu")
    else:
        request.getfixturevalue("cuda_pkg")
    base = _run(backend, {})
    # same arithmetic, different data movement: bit-identical
    for env in ({"CMBL_RG_DIRECT": "0"}, {"CMBL_FFT_FAST": "0"}, {"CMBL_PDL": "0"}, {"CMBL_PDL": "1"}, {"CMBL_COL_JN_RED": "0"}):
        other = _run(backend, env)
        diff = [k for k in base if base[k][0] != other[k][0]]
        assert not diff, (env, diff)
    # different kernels (generic stage kernels on the reference layout; general transpose-δ flow): same numbers within rounding
    for env, keys in (({"CMBL_FLOW_FAST": "0"}, None), ({"CMBL_GRAD_FUSED": "0"}, ("δf", "δϕ"))):
        other = _run(backend, env)
        for k in base:
            tol = 1e-11 if k.startswith("f64") else 2e-4
            assert abs(base[k][1] - other[k][1]) <= tol * abs(base[k][1]), (env, k, base[k][1], other[k][1])
            if keys is not None and k.split(":")[1] not in keys:
                assert base[k][0] == other[k][0], (env, k)


CODE_2048 = r'''
import json, os, sys
import numpy as np, torch
ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
dev, N = "cuda:0", 2048
out = {}
for dt, tT in (("f64", torch.float64), ("f32", torch.float32)):
    proj = pkg.ProjLambert(N, N, 2.0, tT, dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    k = torch.fft.fftfreq(N, device=dev, dtype=torch.float64)
    kk = torch.sqrt(k[:, None] ** 2 + k[None, :] ** 2) + 1e-3
    phi = torch.fft.ifft2(torch.fft.fft2(torch.randn((1, 1, N, N), generator=gen, device=dev, dtype=torch.float64)) / kk ** 3).real
    phi = (phi / phi.std() * 1e-5).to(tT)                                  # arcminute-scale deflections, red spectrum (curvature from the large scales: weak lensing)
    f = pkg.Field("IQUMap", torch.randn((1, 3, N, N), generator=gen, device=dev, dtype=tT), proj)
    L = pkg.LenseFlow(pkg.Field("Map", phi, proj), 3)
    F = pkg.convert(f, "IQUFourier")
    res = {"L": (L * f).arr, "Linv": L.ldiv(f).arr, "LH": (L.H * F).arr}
    probe = torch.Generator(device=dev).manual_seed(11)
    for kname, v in res.items():
        a = torch.view_as_real(v) if v.is_complex() else v
        w = torch.randn(a.shape, generator=probe, device=dev, dtype=torch.float64)
        out[dt + ":" + kname] = [float((a.double() * w).sum()), float(a.double().norm())]      # a random projection and the norm
    out[dt + ":path"] = pkg.load().cdll.cmbl_lenseflow_kernel_path(L.cache(f).handle)
print(json.dumps(out))
'''


@pytest.mark.gpu
def test_fast_stage_kernels_at_2048_match_generic_full_size(request):
    """Nside=2048 IQU (BASELINE config 4's shape): the fast stage kernels (64 KB tiles, 256 threads, [8,16,16]) against the generic kernels on the
    reference layout, whole maps, through a random projection and the norm of every result."""
    request.getfixturevalue("cuda_pkg")
    def run(env):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", CODE_2048, ROOT], capture_output=True, text=True, timeout=900, env=e)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    fast, gen = run({}), run({"CMBL_FLOW_FAST": "0"})
    assert fast["f64:path"] == 3 and gen["f64:path"] == 0
    for k in fast:
        if k.endswith(":path"):
            continue
        tol = 1e-11 if k.startswith("f64") else 5e-5
        (pf, nf), (pg, ng) = fast[k], gen[k]
        assert abs(nf - ng) <= tol * ng and abs(pf - pg) <= tol * ng * 10, (k, fast[k], gen[k])     # |projection| ~ norm: compare on that scale

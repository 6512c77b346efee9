"""Extract the fiducial CAMB spectra shipped with the reference into a small fixture.

Run HERE (container with /root/reference mounted); the GPU box never reads /root/reference.
Source: /root/reference/dat/default_camb_Cls.jld2 (written by dat/compute_default_camb_Cls.jl,
`camb(ℓmax=16000)`, src/cls.jl:181-196).  The JLD2/HDF5 container stores 21 zlib-deflated
Float64[15998] chunks (ℓ = 2..15999); we scan for zlib headers and inflate.
Output: tests/golden/fiducial_cls.npz  (ℓ, unlensed_total TT/EE/BB/TE = unlensed_scalar + tensor(r=0.2), ϕϕ, and the lensed
`total` TT/EE/BB/TE that load_sim uses for Cf̃, src/dataset.jl:270)
stored as float64 for ℓ = 2..LMAX_KEEP.
"""
import zlib, sys, os
import numpy as np

SRC = "/root/reference/dat/default_camb_Cls.jld2"
LMAX_KEEP = 8200   # covers ℓmax=ceil(√2·nyquist)+1 for θpix ≥ 2′ (7638)
ORDER = ["us_TT", "us_EE", "us_BB", "us_TE", "pp",
         "ls_TT", "ls_EE", "ls_BB", "ls_TE",
         "t_TT", "t_EE", "t_BB", "t_TE",
         "ut_TT", "ut_EE", "ut_BB", "ut_TE",
         "tot_TT", "tot_EE", "tot_BB", "tot_TE"]

def main():
    b = open(SRC, "rb").read()
    chunks, i = [], 0
    while i < len(b) - 2:
        if b[i] == 0x78 and b[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                d = zlib.decompressobj()
                raw = d.decompress(b[i:])
                if len(raw) == 15998 * 8:
                    chunks.append(np.frombuffer(raw, dtype="<f8").copy())
                    i += len(b) - i - len(d.unused_data)
                    continue
            except zlib.error:
                pass
        i += 1
    assert len(chunks) == 21, len(chunks)
    C = dict(zip(ORDER, chunks))
    ell = np.arange(2, 16000)
    assert abs(C["ut_TT"][0] - 1071.5229487157596) < 1e-9      # SURVEY App. B anchor
    keep = ell <= LMAX_KEEP
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fiducial_cls.npz")
    np.savez_compressed(out, ell=ell[keep].astype(np.int32),
                        **{k: C[k][keep] for k in ("ut_TT", "ut_EE", "ut_BB", "ut_TE", "pp", "tot_TT", "tot_EE", "tot_BB", "tot_TE")})
    print("wrote", out, os.path.getsize(out), "bytes")

if __name__ == "__main__":
    main()

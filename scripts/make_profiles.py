"""Turns the ncu captures of a GPU pass (gpurun_out/r02_ncu_*.ncu-rep + gpurun_out/binary_sha16.txt) into the tracked summaries under
profiles/: one text summary per capture (key metrics per launch + stall reasons + hottest SASS locations) and profiles/traffic.json
(DRAM bytes per launch of the two stage kernels, keyed by dtype, with the sha256 of the libcmbl_b200.so the capture was taken from — bench.py
reports `roofline.traffic` only when that hash matches the library it has loaded).
usage: python scripts/make_profiles.py [tag] [outdir]   (run on the GPU box right after the captures — the .ncu-rep files are too big to travel)"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = os.path.join(ROOT, "gpurun_out"); P = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")     # (on the GPU box: write next to the captures)
os.makedirs(P, exist_ok=True)
CMD = {"flow_f64": "ncu --set full --clock-control none --import-source on -s 38 -c 2  python scripts/ncu_target.py f64 fwd   (Nside=1024 QU batch 8; one flow_rows + one flow_cols launch of the first RK step)",
       "flow_f32": "ncu --set full --clock-control none --import-source on -s 38 -c 2  python scripts/ncu_target.py f32 fwd",
       "adj_f64": "ncu --set full --clock-control none --import-source on -s 12 -c 2  python scripts/ncu_target.py f64 adj   (adjoint stage kernels)",
       "fft_f64": "ncu --set full --clock-control none --import-source on -s 2 -c 3  python scripts/ncu_target.py f64 adj   (general 2-D transform kernels of the rfft2 inside precompute: persistent column kernel + row pass)"}
sha = open(os.path.join(G, "binary_sha16.txt")).read().strip() if os.path.exists(os.path.join(G, "binary_sha16.txt")) else None
def csrc_sha16():          # same definition as bench.py
    import glob, hashlib
    d = os.path.join(ROOT, "cmblensing.jl_b200", "csrc"); h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + [os.path.join(d, "Makefile")]):
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    return h.hexdigest()[:16]
traffic = {"binary_sha16": sha, "source_sha16": csrc_sha16(), "source": f"profiles/{tag}_ncu_flow_{{f64,f32}}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
for key, cmd in CMD.items():
    rep = os.path.join(G, f"{tag}_ncu_{key}.ncu-rep")
    if not os.path.exists(rep):
        continue
    s1 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    s2 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), rep, "12"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_{key}.txt"), "w").write(cmd + f"\nlibcmbl_b200.so sha256[:16] = {sha}\n" + s1 + s2)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr = rows[0]; kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    for r in rows[2:]:
        nm = "flow_rows" if "RowBody" in r[kn] and "C2C" not in r[kn] else ("flow_cols" if "ColBody" in r[kn] and "R2C" not in r[kn] and "C2R" not in r[kn] else None)
        if nm and key.startswith("flow_"):
            b = float(r[rd]) * unit[rows[1][rd]] + float(r[wr]) * unit[rows[1][wr]]
            traffic.setdefault(key[5:], {})[nm] = b
    print("wrote", f"profiles/{tag}_ncu_{key}.txt")
# The --set full captures above hold ONE launch of each stage kernel (a middle RK4 stage: 6C + 2Cϕ plane passes for the column kernel, against
# 4C + 2Cϕ for the first and the last stage).  bench.py's `roofline.achieved` is an average over all launches, so the traffic it is compared
# with is the mean over the four stages of one RK4 step, from the metrics-only pass gpurun_out/<tag>_stage_traffic_by_kind_<dtype>.csv.
unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
for dt in ("f64", "f32"):
    f = os.path.join(G, f"{tag}_stage_traffic_by_kind_{dt}.csv")
    if not os.path.exists(f):
        continue
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki, idc, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        per.setdefault((r[idc], r[ki]), {})[r[mn]] = float(r[mv].replace(",", "")) * unit.get(r[mu], 1.0)
    out = {"flow_rows": [], "flow_cols": []}; lines = []
    for (i, k), m in sorted(per.items(), key=lambda kv: int(kv[0][0])):
        nm = "flow_rows" if "RowBody" in k else ("flow_cols" if "FastColBody" in k else None)
        if nm:
            out[nm].append(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"])
            lines.append(f"launch {i} {nm}: read {m['dram__bytes_read.sum']/1e6:.1f} MB, written {m['dram__bytes_write.sum']/1e6:.1f} MB, {m['gpu__time_duration.sum']:.0f} ns, L2 read hit {m['lts__t_sector_op_read_hit_rate.pct']:.1f} %")
    if out["flow_cols"]:
        traffic[dt + "_full_capture_launch"] = traffic.get(dt)
        traffic[dt] = {k: sum(v) / len(v) for k, v in out.items() if v}
        traffic[dt + "_by_stage"] = out
        open(os.path.join(P, f"{tag}_stage_traffic_by_kind_{dt}.txt"), "w").write(
            "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct --clock-control none -s 20 -c 8  python scripts/ncu_target.py "
            + dt + " fwd   (one RK4 step: 4 row + 4 column launches)\nlibcmbl_b200.so sha256[:16] = " + str(sha) + "\n" + "\n".join(lines) + "\n")
traffic["source"] = f"mean over the four stages of one RK4 step (profiles/{tag}_stage_traffic_by_kind_*.txt, ncu metrics pass); *_full_capture_launch: the single launch of the --set full capture (profiles/{tag}_ncu_flow_*.txt)"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic))

"""Stall samples of a kernel split at its BAR.SYNC instructions (phases of a block-synchronous kernel).
usage: ncu_phases.py rep kernel-substring"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out)); want = sys.argv[2]
kern, hdr, body, done = None, None, [], False
def flush():
    global done
    if not body or want not in kern or done: return
    done = True
    si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
    tot = sum(int(r[si]) for r in body)
    print(kern, 'samples', tot)
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    ridx = [hdr.index(h) for h in reasons]
    acc, n, start, iex = 0, 0, 0, 0
    racc = [0] * len(reasons)
    for i, r in enumerate(body):
        acc += int(r[si]); n += 1; iex += int(r[ex])
        for k, j in enumerate(ridx): racc[k] += int(r[j] or 0)
        if 'BAR.SYNC' in r[src] or i == len(body) - 1:
            top = sorted(zip(racc, reasons), reverse=True)[:4]
            print(f"  sass {start:5d}-{i:5d}: {acc:6d} samples {100*acc/tot:5.1f}%  instr {iex:9d}  " + ' '.join(f"{nm[6:]}={v}" for v, nm in top))
            acc, n, start, iex = 0, 0, i + 1, 0
            racc = [0] * len(reasons)
for r in rows:
    if r and r[0] == "Kernel Name": flush(); kern = r[1]; body = []; hdr = None
    elif r and r[0] == "Address": hdr = r
    elif r and hdr: body.append(r)
flush()

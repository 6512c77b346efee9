"""Top stall locations (SASS) of each kernel in an .ncu-rep, from the source page.  usage: ncu_hot.py rep [topN]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(out))
kern, hdr, body = None, None, []
def flush():
    if not body: return
    si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source")
    tot = sum(int(r[si]) for r in body)
    print(f"== {kern}  total samples {tot}, {len(body)} SASS instructions")
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:top]
    for i in sorted(idx):
        prev = body[i - 1][src].strip() if i else ''
        print(f"  {i:5d} {int(body[i][si]):6d} {100*int(body[i][si])/max(tot,1):5.1f}%  {body[i][src].strip()[:90]}")
for r in rows:
    if r and r[0] == "Kernel Name": flush(); kern = r[1]; body = []; hdr = None
    elif r and r[0] == "Address": hdr = r
    elif r and hdr: body.append(r)
flush()

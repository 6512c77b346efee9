import csv, subprocess, sys
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','source','--csv'],capture_output=True,text=True).stdout.splitlines()
rows=list(csv.reader(out)); want=sys.argv[2]
kern=None; hdr=None; body=[]; done=False
def flush():
    global done
    if not body or want not in kern or done: return
    done=True
    ex=hdr.index("L1 Wavefronts Shared Excessive"); wf=hdr.index("L1 Wavefronts Shared"); src=hdr.index("Source")
    tot=sum(int(r[wf] or 0) for r in body); totex=sum(int(r[ex] or 0) for r in body)
    print(kern[:70],'wavefronts',tot,'excessive',totex)
    # split by BAR.SYNC phases
    acc=0;accx=0;start=0
    for i,r in enumerate(body):
        acc+=int(r[wf] or 0); accx+=int(r[ex] or 0)
        if 'BAR.SYNC' in r[src] or i==len(body)-1:
            if acc: print(f"  sass {start:5d}-{i:5d}: wavefronts {acc:9d} excessive {accx:9d} ({100*accx/max(acc,1):.0f}%)")
            acc=0;accx=0;start=i+1
    idx=sorted(range(len(body)),key=lambda i:-int(body[i][ex] or 0))[:12]
    for i in sorted(idx): print(f"   {i:5d} ex {int(body[i][ex] or 0):8d} wf {int(body[i][wf] or 0):8d}  {body[i][src].strip()[:80]}")
for r in rows:
    if r and r[0]=="Kernel Name": flush(); kern=r[1]; body=[]; hdr=None
    elif r and r[0]=="Address": hdr=r
    elif r and hdr: body.append(r)
flush()

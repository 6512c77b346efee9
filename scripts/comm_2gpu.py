"""The C-ABI communicator on real ranks (one process per GPU): rank 0 creates the NCCL id with cmbl_comm_unique_id, the launcher's own
channel (here torch.distributed/gloo, only for those 128 bytes) hands it out, every rank calls cmbl_comm_init and then runs
cmbl_wiener_cg_sharded on ITS shard of one batch.  Checked: every rank stops at the iteration the unsharded solve of the whole batch stops at,
and returns the bits of its items.
usage: torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/comm_2gpu.py"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from common import make_problem
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("gloo")
lib = pkg.load()
uid = [pkg.Comm.unique_id(lib) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = pkg.Comm(lib, world, rank, uid[0])
s = comm.allreduce([rank + 1.0, 1.0], "sum"); m = comm.allreduce([float(rank)], "max")
assert s[0] == world * (world + 1) / 2 and s[1] == world and m[0] == world - 1
nb = 2 * world
pr = make_problem(pkg, 256, 256, "P", "f64", nb=nb, nsteps=7, mask=True, seed=5, theta=2.0, device=dev)
kw = dict(tol=3e1, nsteps=200)
x_all, h_all = pkg.argmaxf_logpdf(pr["ds"], pr["phi"], conjgrad_kwargs=kw)                      # the whole batch on every rank: the reference answer
sl = slice(2 * rank, 2 * rank + 2)
ds = pr["ds"]
F = lambda f: pkg.Field(f.basis, f.arr[sl].contiguous(), f.proj)
dsr = pkg.BaseDataSet(F(ds.d), ds.Cf, ds.Cn, ds.B, ds.Mf, ds.Mpix, nsteps=7)
x, h = pkg.argmaxf_logpdf(dsr, F(pr["phi"]), conjgrad_kwargs=kw, comm=comm)
ok = len(h) == len(h_all) and torch.equal(x.arr, x_all.arr[sl]) and all(np.array_equal(a[1], b[1][sl]) for a, b in zip(h, h_all))
out = [None] * world
dist.all_gather_object(out, dict(rank=rank, iters=len(h), iters_unsharded=len(h_all), bits_equal=bool(ok)))
if rank == 0:
    print(f"cmbl_comm over {world} ranks (NCCL via dlopen): all-reduce ok; sharded CG-Wiener, 2 batch items per rank, Nside=256 QU fp64")
    for o in out:
        print("   ", o)
    assert all(o["bits_equal"] for o in out)
    print("    every rank stopped at the unsharded iteration and returned the unsharded bits of its items")
comm.close()
dist.destroy_process_group()

#!/usr/bin/env python
"""Multi-GPU checks on real devices (run under torchrun, one rank per GPU, NCCL):

  1. data-subsampling estimator: X row-sharded over the ranks, one all-reduce of 1 + d (+ d*d)
     doubles; the result must equal the single-GPU evaluation of the same (beta, u) to 1e-12
     (the sum order differs), and is timed at n = 11 M x 28, m = 550 000;
  2. chain sharding: B independent SV problems split into contiguous blocks over the ranks, outputs
     all-gathered; must equal the single-GPU batch bit for bit.

usage: torchrun --nproc-per-node W tools/check_multi_gpu.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_inputs as gi  # noqa: E402
from pmmh_qn_b200 import kernels as K, sharding  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


# ---- 1. subsampling estimator, row-sharded + one all-reduce
n, d, m = 11_000_000, 28, 550_000
g = torch.Generator(device=dev)
g.manual_seed(0)                      # every rank generates the same data, keeps its rows
per = (n + world - 1) // world
b, e = min(n, rank * per), min(n, (rank + 1) * per)
beta = 0.1 * torch.randn((d,), dtype=torch.float64, device=dev, generator=g)
u = torch.randn((m,), dtype=torch.float64, device=dev, generator=g)
rows = []
ys = []
chunk = 1_000_000
full_ref = None
if rank == 0:
    full_x = torch.empty((n, d), dtype=torch.float64, device=dev)
    full_y = torch.empty((n,), dtype=torch.float64, device=dev)
for c0 in range(0, n, chunk):         # same stream of random numbers on every rank
    c1 = min(n, c0 + chunk)
    xc = torch.randn((c1 - c0, d), dtype=torch.float64, device=dev, generator=g)
    yc = (torch.rand((c1 - c0,), dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(xc @ beta)).to(torch.float64)
    lo, hi = max(c0, b), min(c1, e)
    if hi > lo:
        rows.append(xc[lo - c0:hi - c0].clone())
        ys.append(yc[lo - c0:hi - c0].clone())
    if rank == 0:
        full_x[c0:c1] = xc
        full_y[c0:c1] = yc
x_sh = torch.cat(rows) if rows else torch.empty((0, d), dtype=torch.float64, device=dev)
y_sh = torch.cat(ys) if ys else torch.empty((0,), dtype=torch.float64, device=dev)
del rows, ys
idx = K.subsample_indices(u, n)       # redundantly on every rank (deterministic)
for hess in (False, True):
    ws = K.Workspace()

    def step():
        out = K.logistic_loglike(x_sh, y_sh, idx, beta, compute_hessian=hess, row_begin=b, row_end=e, workspace=ws)
        sharding.allreduce_sum_(out)
        return out

    out = step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = K.logistic_loglike(full_x, full_y, idx, beta, compute_hessian=hess)
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        assert err <= 1e-12, err
        say(check="subsampling estimator sharded over %d GPUs == single GPU" % world, hessian=int(hess),
            max_rel_err=err, ms_per_evaluation=float(tt.item()), rows=m)
if rank == 0:
    del full_x, full_y
del x_sh, y_sh
torch.cuda.empty_cache()

# ---- 2. chain sharding
B, nn, nobs, lag = 12, 3000, 120, 10
rs = np.random.RandomState(5)
obs = torch.from_numpy(gi.sv_obs(nobs)).to(dev)
params = np.array(gi.SV_PARAM_SETS[0]) + 0.02 * rs.normal(size=(B, 4))
uu = rs.normal(size=(B, nobs, nn))
rvr = rs.uniform(size=(B, nobs))
(bb, ee) = sharding.block_range(B, rank, world)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
mine = K.flps_sv_corr(obs, t(params[bb:ee]), t(rvr[bb:ee]), t(uu[bb:ee]), lag=lag) if ee > bb else None
keys = ("log_like", "gradient", "filt")
gathered = {}
for k in keys:
    shape = {"log_like": (0,), "gradient": (0, 4, nobs), "filt": (0, nobs)}[k]
    loc = mine[k] if mine is not None else torch.empty(shape, dtype=torch.float64, device=dev)
    gathered[k] = sharding.allgather_blocks(loc, B)
if rank == 0:
    whole = K.flps_sv_corr(obs, t(params), t(rvr), t(uu), lag=lag)
    torch.cuda.synchronize()
    for k in keys:
        assert torch.equal(gathered[k], whole[k]), k
    say(check="%d SV chains sharded over %d GPUs == single-GPU batch (bit for bit)" % (B, world), ok=True)
dist.barrier()
dist.destroy_process_group()

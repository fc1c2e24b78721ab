"""Split particle filter over real ranks (torchrun, NCCL): parity against the in-process run and the
oracle at a size the oracle finishes quickly, then timings with a Philox stream.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
      --master-port 29533 tools/run_split_dist.py [logN ...] [--T 100] [--lag 10] [--out file.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import golden_inputs as gi
from pmmh_qn_b200 import kernels as K
from pmmh_qn_b200.state.particle_methods import split as SP

ap = argparse.ArgumentParser()
ap.add_argument("logn", nargs="*", type=int, default=[22])
ap.add_argument("--T", type=int, default=100)
ap.add_argument("--lag", type=int, default=10)
ap.add_argument("--out", default="")
ap.add_argument("--no-parity", action="store_true")
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = SP.DistComm()
lines = []


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)
        lines.append(d)


# ---- parity: NCCL ranks vs the oracle (and vs in-process ranks on rank 0's GPU)
if not args.no_parity:
    import oracle
    from helpers import relerr, to_time_major
    n, nobs, lag = 1 << 18, 41, 10
    obs, params, rvr, rvp = gi.sv_inputs(n, nobs, 11)
    u = torch.from_numpy(to_time_major(rvp, n, nobs)).to(dev)
    rvr_d = torch.from_numpy(rvr[:nobs].copy()).to(dev)
    out = SP.run_split_smoother(comm, obs, params, n, lag, rvr_d, u_d=u, device=dev, keep_history=True)
    torch.cuda.synchronize()
    # gather the last generation for a value check
    xl = out["x_hist"][-1][0]
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([xl.numel()], dtype=torch.int64, device=dev))
    pad = torch.zeros(n, dtype=torch.float64, device=dev)
    pad[:xl.numel()] = xl
    allx = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allx, pad)
    if rank == 0:
        ref = oracle.flps_sv_corr(obs, params, rvr, rvp, n, lag, 0, dumps=True)
        xg = np.concatenate([allx[r][:int(sizes[r].item())].cpu().numpy() for r in range(world)])
        g, gr = out["gradient"].cpu().numpy(), ref["gradient"]
        emit({"check": "nccl_vs_oracle", "world": world, "N": n, "T": nobs - 1, "lag": lag,
              "log_like": float(out["log_like"].item()), "oracle_log_like": float(ref["log_like"]),
              "rel_log_like": abs(float(out["log_like"].item()) - ref["log_like"]) / abs(ref["log_like"]),
              "grad_err_over_max": float(np.max(np.abs(g - gr)) / np.max(np.abs(gr))),
              "smo_rel": relerr(out["smo"].cpu().numpy(), ref["smo"]),
              "final_generation_rel": relerr(xg, ref["X"][-1]),
              "near_ties_rank0": out["diag"][0], "status": out["diag"][2],
              "ok": bool(abs(float(out["log_like"].item()) - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
                         and np.max(np.abs(g - gr)) <= 1e-9 * np.max(np.abs(gr))
                         and relerr(xg, ref["X"][-1]) <= 1e-12)})
    del u, out
    torch.cuda.empty_cache()

# ---- timings with a Philox stream (u is never stored)
for logn in args.logn:
    n, nobs = 1 << logn, args.T + 1
    obs = gi.sv_obs(nobs)
    params = np.array(gi.SV_PARAM_SETS[0], dtype=np.float64)
    ph = SP.PhiloxRVS(seed=5, offset=0)
    rvr = K.norm_cdf(ph.resampling_normals(nobs, n, dev))
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        out = SP.run_split_smoother(comm, obs, params, n, args.lag, rvr, philox=(5, 0), device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        mx = torch.tensor([float(out["counts"].max())], dtype=torch.float64, device=dev)
    emit({"bench": "split_pf", "world": world, "N": n, "T": args.T, "lag": args.lag, "seconds": float(dt.item()),
          "particle_steps_per_s": n * args.T / float(dt.item()), "log_like": float(out["log_like"].item()),
          "max_arrivals_per_rank": int(mx.item()), "near_ties_rank0": out["diag"][0], "status": out["diag"][2]})
    del out
    torch.cuda.empty_cache()
if rank == 0 and args.out:
    with open(args.out, "a") as fh:
        for d in lines:
            fh.write(json.dumps(d) + "\n")
dist.destroy_process_group()

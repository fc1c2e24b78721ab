"""CPU, world_size = 2, gloo: the two multi-GPU decompositions of the path.

(1) subsampling estimator: row-sharded partial sums + ONE all-reduce == unsharded oracle;
(2) chain-batched sharding: block shards of a batch of problems, results all-gathered."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_inputs as gi
    import oracle
    from pmmh_qn_b200 import sharding as S

    # (1) sharded logistic sum with a single all-reduce
    n_data, d, m = 20000, 28, 1500
    x, y, beta = gi.logit_data(n_data, d, seed=4)
    idx = oracle.subsample_indices(gi.logit_u(m, 1), n_data)
    b, e = S.block_range(n_data, rank, world)
    part = torch.from_numpy(S.subsample_partial_sums(x[b:e], y[b:e], idx, beta, b, compute_hessian=True))
    S.allreduce_sum_(part)
    ref = oracle.logistic_loglike_gradient(beta, x, y, idx, True, True)
    got = part.numpy()
    ok1 = (abs(got[0] - ref["log_like"]) <= 1e-11 * abs(ref["log_like"])
           and np.max(np.abs(got[1:1 + d] - ref["gradient"])) <= 1e-10 * np.max(np.abs(ref["gradient"]))
           and np.max(np.abs(got[1 + d:].reshape(d, d) - ref["hessian"])) <= 1e-10 * np.max(np.abs(ref["hessian"])))

    # (2) chain-batched sharding: B problems, each rank evaluates its block, all-gather
    B, n, nobs = 5, 64, 40
    obs = gi.sv_obs(nobs)
    lls = []
    arrays = {"seed": np.arange(B)}
    mine, (pb, pe) = S.shard_batch(arrays, rank, world)
    for s in mine["seed"]:
        rvr, rvp = gi.split_particle(gi.sv_rvs(n, nobs, int(s)), nobs)
        lls.append(oracle.flps_sv_corr(obs, np.array(gi.SV_PARAM_SETS[0]), rvr, rvp, n, 4, 0)["log_like"])
    local = torch.tensor(lls, dtype=torch.float64).reshape(-1, 1)
    allv = S.allgather_blocks(local, B).numpy().reshape(-1)
    want = []
    for s in range(B):
        rvr, rvp = gi.split_particle(gi.sv_rvs(n, nobs, s), nobs)
        want.append(oracle.flps_sv_corr(obs, np.array(gi.SV_PARAM_SETS[0]), rvr, rvp, n, 4, 0)["log_like"])
    ok2 = np.array_equal(allv, np.array(want))
    ret[rank] = bool(ok1 and ok2)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

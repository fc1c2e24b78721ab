"""CPU, world_size = 2, gloo: the two multi-GPU decompositions of the path.

(1) subsampling estimator: row-sharded partial sums + ONE all-reduce == unsharded oracle;
(2) chain-batched sharding: block shards of a batch of problems, results all-gathered;
(3) split particle filter: the interval plan of the tail exchange and the communicator wrapper."""
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

    # (3) split particle filter, host side of the tail exchange: weights of generation i live in
    # blocks of n_src[r] sorted positions, the final generation wants blocks of n_dst[r]; one
    # all_to_all_single with the interval plan must deliver exactly the wanted slice, in order
    from pmmh_qn_b200.state.particle_methods.split import DistComm
    comm = DistComm()
    n_src, n_dst = [700, 324], [301, 723]
    full = np.arange(1024, dtype=np.float64) * 0.5
    s0, d0 = sum(n_src[:rank]), sum(n_dst[:rank])
    mine = torch.from_numpy(full[s0:s0 + n_src[rank]].copy())
    got = torch.zeros(800, dtype=torch.float64)
    sc, rc = S.interval_exchange_counts(n_src, n_dst, rank)
    comm.all_to_all([mine], [sc], [got], [rc])
    ok3 = np.array_equal(got[:n_dst[rank]].numpy(), full[d0:d0 + n_dst[rank]])
    # the per-step exchange after pmmh_svsplit_pack_direct: the self part is already in place and must stay
    # untouched, the other parts land at the unchanged offsets (rows of LR doubles)
    LR = 3
    sc2, rc2 = ([2, 3], [2, 4]) if rank == 0 else ([4, 1], [3, 1])
    send2 = torch.arange(5 * LR, dtype=torch.float64).reshape(5, LR) + 100.0 * (rank + 1)
    recv2 = torch.full((8, LR), -1.0, dtype=torch.float64)
    comm.all_to_all([send2], [sc2], [recv2], [rc2], skip_self=True)
    if rank == 0:    # rows 0..1 = self (untouched), rows 2..5 = rank 1's rows 0..3
        want2 = np.vstack([np.full((2, LR), -1.0), np.arange(4 * LR).reshape(4, LR) + 200.0, np.full((2, LR), -1.0)])
    else:            # rows 0..2 = rank 0's rows 2..4, row 3 = self (untouched)
        want2 = np.vstack([np.arange(2 * LR, 5 * LR).reshape(3, LR) + 100.0, np.full((5, LR), -1.0)])
    ok3 = ok3 and np.array_equal(recv2.numpy(), want2)
    # and the small all-gather / all-reduce the time loop uses
    g_out = torch.zeros((world, 4), dtype=torch.float64)
    comm.all_gather([torch.full((4,), float(rank + 1), dtype=torch.float64)], [g_out])
    ok3 = ok3 and g_out[:, 0].tolist() == [1.0, 2.0]
    red = torch.full((3, 8), float(rank + 1), dtype=torch.float64)
    comm.all_reduce_sum([red])
    ok3 = ok3 and bool((red == 3.0).all())
    # a rank-local abandon (capacity overflow on one rank only) is agreed on before anybody acts on it:
    # every rank sees the largest status word, so all of them leave the time loop in the same step
    ok3 = ok3 and comm.agree_status(8 if rank == 1 else 0, torch.device("cpu")) == 8
    ok3 = ok3 and comm.agree_status(0, torch.device("cpu")) == 0
    ret[rank] = bool(ok1 and ok2 and ok3)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

"""Host-side sharding helpers for the 8 GPUs of one B200 box (one process per GPU,
``torch.distributed``).  Only two things on this path shard:

* independent chains / parameter proposals: a batch of B (theta, u) problems is split into
  contiguous blocks, one block per rank, with NO collective on the data path; the O(T) outputs
  are all-gathered at the end if every rank needs them;
* the data-subsampling estimator: the regressor matrix is row-sharded, every rank reduces the
  sampled rows it owns and ONE all-reduce(sum) of 1 + d (+ d*d) doubles combines them
  (state/direct/cuda.py).

Everything here works on CPU tensors with the gloo backend as well (that is how it is tested
without GPUs); the device kernels never appear in this file.
"""
import numpy as np
import torch


def world_info(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def block_range(n, rank, world):
    """Contiguous block [begin, end) of n items owned by `rank` (ceil-sized blocks; trailing
    ranks may own nothing)."""
    per = (int(n) + world - 1) // world
    begin = min(int(n), rank * per)
    return begin, min(int(n), begin + per)


def shard_batch(arrays, rank, world):
    """Slice the leading (problem) axis of every array in `arrays` for this rank."""
    n = len(next(iter(arrays.values())))
    b, e = block_range(n, rank, world)
    return {k: v[b:e] for k, v in arrays.items()}, (b, e)


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks (the single exchange step of the subsampling estimator)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allgather_blocks(local, n_total, group=None):
    """All-gather per-problem results that were computed on block_range shards.
    `local` is a tensor [n_local, ...]; returns [n_total, ...] on every rank."""
    import torch.distributed as dist
    rank, world = world_info(group)
    if world == 1:
        return local
    per = (int(n_total) + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out, dim=0)[:n_total]


def subsample_partial_sums(x_shard, y_shard, idx, beta, row_begin, compute_hessian=False):
    """NumPy restatement of what one rank contributes to the sharded logistic sum (used by the
    gloo test and as documentation of the decomposition; the device path is
    kernels.logistic_loglike).  Returns [1 + d + d*d]."""
    d = x_shard.shape[1]
    idx = np.asarray(idx, dtype=np.int64)
    mine = idx[(idx >= row_begin) & (idx < row_begin + x_shard.shape[0])] - row_begin
    out = np.zeros(1 + d + d * d)
    if mine.size == 0:
        return out
    x = x_shard[mine]
    y = y_shard[mine]
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        xb = x.dot(beta)
        en, ep = np.exp(-xb), np.exp(xb)
        eta = 1.0 / (1.0 + en)
        e1, e0 = np.log(eta), np.log(1.0 - eta)
        e1[np.isinf(e1)] = 0.0
        e0[np.isinf(e0)] = 0.0
        out[0] = np.sum(y * e1 + (1.0 - y) * e0)
        out[1:1 + d] = (x.T * (y / (1.0 + ep) - (1.0 - y) / (1.0 + en))).sum(axis=1)
        if compute_hessian:
            s = y * (-ep / (1.0 + ep) ** 2) + (1.0 - y) * (-en / (1.0 + en) ** 2)
            out[1 + d:] = (-(x.T * s).dot(x)).reshape(-1)
    return out


def interval_exchange_counts(n_src, n_dst, rank):
    """A vector of length sum(n_src) is stored as consecutive blocks, block r (n_src[r] entries) on
    rank r; it is wanted as consecutive blocks of n_dst[r] entries.  Returns (send_counts,
    recv_counts) of `rank` for one all_to_all_single: what it sends to every destination is a
    run of its block in destination order, what it receives arrives in source order -- so both
    buffers are contiguous and no index list travels.  Used by the split particle filter to put
    the weights of generation i where the tail terms of the final generation need them
    (state/particle_methods/split.py)."""
    n_src = np.asarray(n_src, dtype=np.int64)
    n_dst = np.asarray(n_dst, dtype=np.int64)
    assert n_src.sum() == n_dst.sum()
    s0 = np.concatenate([[0], np.cumsum(n_src)])
    d0 = np.concatenate([[0], np.cumsum(n_dst)])
    world = len(n_src)

    def overlap(a0, a1, b0, b1):
        return int(max(0, min(a1, b1) - max(a0, b0)))

    send = [overlap(s0[rank], s0[rank + 1], d0[d], d0[d + 1]) for d in range(world)]
    recv = [overlap(s0[s], s0[s + 1], d0[rank], d0[rank + 1]) for s in range(world)]
    return send, recv

"""ONE SV particle filter / fixed-lag smoother with its particles split over several GPUs
(BASELINE.json configs[4]; SURVEY.md 8e "single large PF").

The algorithm is the reference's ``flps_sv_corr``
(/root/reference/python/state/particle_methods/stochastic_volatility.pyx:205-655, called from
``ParticleMethodsCython.smoother`` at state/particle_methods/cython.py:97).  The reference has
no multi-device call, so the split is new API; the estimator class below keeps the reference's
estimator contract (``smoother(model, **kw) -> bool``, ``results`` keys, ``settings``,
``dim_rvs``, ``alg_type``).

One process per GPU (``torch.distributed``, NCCL).  Rank r owns a contiguous value range of the
sorted generation.  Per time step the device phases of ``csrc/sv_split.cu`` alternate with three
exchanges, all enqueued on the current CUDA stream:

    weights(t)  -> all_gather  4 doubles / rank   (weight total, count, min, max)
    children    -> all_gather  4096 ints / rank   (value histogram of the children)
    plan, pack  -> all_to_all  records by value range (counts read back from pinned memory)
    sort        -> weights(t+1)

Parents are never fetched remotely: the children of local parents are a contiguous range of the
N systematic resampling points, known from the all-gathered weight totals.  The fixed-lag score
terms need the values of a particle's ancestors lag-1 and lag-2 steps back; they travel with the
particle (records of ``lag`` doubles).  At the end one all_reduce combines the O(T) local sums
and ``lag - 1`` small all_to_alls put the weights of the last generations where the tail terms
(:540-562) need them.

``LocalComm`` drives several ranks inside ONE process on one device with the same code path
(device copies instead of NCCL); the GPU tests use it to check the partition logic on a single
GPU, and it is how a single GPU runs N too large for the shared-memory kernels.
"""
import ctypes

import numpy as np
import torch

from ... import _lib
from ... import kernels as K
from ...sharding import block_range, interval_exchange_counts
from ..base_state_inference import BaseStateInference

_F64 = torch.float64
HIST_BINS = 4096


class PhiloxRVS(object):
    """Auxiliary variables defined by a Philox4x32-10 stream instead of an array: element
    t * N + j of stream (seed, offset) is u[t][j] (Box-Muller, exactly the numbers
    ``pmmh_crank_nicolson`` draws), the n_obs resampling normals follow it.  At N = 2^26,
    T = 1000 the array would be 537 GB; any rank regenerates any entry instead."""
    __slots__ = ("seed", "offset")

    def __init__(self, seed, offset=0):
        self.seed = int(seed)
        self.offset = int(offset)

    def __deepcopy__(self, memo):
        return self

    def resampling_normals(self, n_obs, n_particles, device):
        off = self.offset + (n_obs * n_particles + 1) // 2
        z = torch.zeros(n_obs, dtype=_F64, device=device)
        return K.crank_nicolson(z, 1.0, seed=self.seed, philox_offset=off)

    def materialise(self, n_obs, n_particles, device):
        """u as an [n_obs, N] tensor (tests / small N only)."""
        z = torch.zeros(n_obs * n_particles, dtype=_F64, device=device)
        return K.crank_nicolson(z, 1.0, seed=self.seed, philox_offset=self.offset).reshape(n_obs, n_particles)


def default_capacities(n_total, world):
    """(cap_particles, cap_children) per rank.  Arrivals are balanced to one histogram bin
    (~0.2 % of N).  Children follow the weight mass of the local parents, which is NOT balanced:
    after an outlying observation the rank with the largest values can own most of the weight
    (seen: > N/2 on one of 8 ranks), so the children buffers are sized for all N (94 bytes per
    child at lag 10: 6.3 GB of the 180 GB at N = 2^26)."""
    per = -(-n_total // world)
    if world == 1:
        return n_total, n_total
    cap = min(n_total, int(per * 1.3) + 65536)
    return cap, n_total


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class _Rank(object):
    """Device state of one rank and the C-ABI calls on it."""

    def __init__(self, rank, world, n_total, n_obs, lag, device, cap=None, capc=None):
        self.lib = _lib.load()
        self.rank, self.world, self.N, self.n_obs, self.lag = rank, world, int(n_total), int(n_obs), int(lag)
        self.LR = max(1, self.lag)
        self.device = device
        dcap, dcapc = default_capacities(self.N, world)
        self.cap = int(cap or dcap)
        self.capc = int(capc or dcapc)
        rows = max(self.cap, self.capc) if world == 1 else self.cap
        nbytes = ctypes.c_size_t()
        _lib.check(self.lib.pmmh_svsplit_workspace_bytes(self.cap, self.capc, ctypes.byref(nbytes)),
                   "pmmh_svsplit_workspace_bytes")
        self.ws_bytes = nbytes.value
        dev = device
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.xs = torch.empty(self.cap, dtype=_F64, device=dev)
        self.perm = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.rec = torch.empty((rows, self.LR), dtype=_F64, device=dev)
        self.rec_next = torch.empty((rows, self.LR), dtype=_F64, device=dev)
        self.send = torch.empty((max(self.capc, rows if world == 1 else 0), self.LR), dtype=_F64, device=dev)
        # values alone (8 bytes per particle): exchanged first so that the sort can start while the
        # 8 * lag byte records are still crossing NVLink
        # (measured, N = 2^24, lag 10: 2 GPUs 4.8e9 -> 6.4e9 particle-steps/s; 8 GPUs 1.29e10 -> 1.17e10,
        # where a second all-to-all per step costs more than the overlap returns: used up to 4 ranks)
        split_keys = 1 < world <= 4 and self.LR > 1
        self.send_keys = torch.empty(self.capc, dtype=_F64, device=dev) if split_keys else None
        self.recv_keys = torch.empty(self.cap, dtype=_F64, device=dev) if split_keys else None
        self.sums = torch.zeros((self.n_obs, 8), dtype=_F64, device=dev)
        self.shift = torch.zeros(self.n_obs, dtype=_F64, device=dev)
        self.xmin = torch.zeros(self.n_obs, dtype=_F64, device=dev)
        self.gather_send = torch.zeros(4, dtype=_F64, device=dev)
        self.gather = torch.zeros((world, 4), dtype=_F64, device=dev)
        self.gather_hist = torch.zeros((self.n_obs, world, 4), dtype=_F64, device=dev)
        self.hist_send = torch.zeros(HIST_BINS, dtype=torch.int32, device=dev)
        self.hist = torch.zeros((world, HIST_BINS), dtype=torch.int32, device=dev)
        self.h_counts = torch.zeros(2 * world + 4, dtype=torch.int32).pin_memory()
        self.n_local = 0
        # sh of the last `lag` generations (tail terms), slot = t % lag
        self.sh_keep = torch.empty((self.LR, self.cap), dtype=_F64, device=dev) if self.lag > 0 else None
        self.n_keep = {}

    def _c(self):
        return (_p(self.ws), self.ws_bytes, self.cap, self.capc)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def init(self, params_d):
        b, e = block_range(self.N, self.rank, self.world)
        self.n_local = e - b
        if self.n_local > self.cap:
            raise _lib.PmmhError("cap_particles too small for the initial split")
        _lib.check(self.lib.pmmh_svsplit_init(_p(self.ws), self.ws_bytes, self.N, self.n_obs, self.world,
                                              self.rank, self.lag, self.cap, self.capc, self.n_local,
                                              _p(params_d), _p(self.xs),
                                              _p(self.perm), _p(self.rec), self._stream()),
                   "pmmh_svsplit_init")

    def weights(self, t, obs_d, params_d):
        keep = None
        if self.sh_keep is not None and t >= self.n_obs - self.lag:
            keep = self.sh_keep[t % self.lag]
            self.n_keep[t] = self.n_local
        _lib.check(self.lib.pmmh_svsplit_weights(*self._c(), t, self.n_local, self.lag, self.n_obs,
                                                 _p(obs_d), _p(params_d),
                                                 _p(self.xs), _p(self.perm), _p(self.rec), _p(self.sums),
                                                 _p(self.gather_send), None if keep is None else _p(keep),
                                                 self._stream()), "pmmh_svsplit_weights")

    def children(self, t, obs_d, params_d, rvr_d, u_d, seed, offset):
        _lib.check(self.lib.pmmh_svsplit_children(*self._c(), t, self.n_local, _p(obs_d), _p(params_d),
                                                  _p(rvr_d), None if u_d is None else _p(u_d), seed, offset,
                                                  _p(self.gather), _p(self.xs), _p(self.hist_send),
                                                  _p(self.shift), _p(self.xmin), self._stream()),
                   "pmmh_svsplit_children")

    def plan_and_pack(self):
        _lib.check(self.lib.pmmh_svsplit_plan(*self._c(), self.world, _p(self.hist),
                                              ctypes.c_void_p(self.h_counts.data_ptr()), self._stream()),
                   "pmmh_svsplit_plan")
        if self.world > 1:
            # children that stay on this rank go straight to their arrival slots in the receive buffers: the
            # exchange moves only what crosses ranks (the self part was a device copy of up to 80 B per particle)
            _lib.check(self.lib.pmmh_svsplit_pack_direct(*self._c(), _p(self.perm), _p(self.rec), _p(self.send),
                                                         None if self.send_keys is None else _p(self.send_keys),
                                                         _p(self.rec_next),
                                                         None if self.recv_keys is None else _p(self.recv_keys),
                                                         self._stream()), "pmmh_svsplit_pack_direct")
            return
        _lib.check(self.lib.pmmh_svsplit_pack(*self._c(), _p(self.perm), _p(self.rec), _p(self.send),
                                              None if self.send_keys is None else _p(self.send_keys),
                                              self._stream()), "pmmh_svsplit_pack")

    def counts(self):
        """(send counts, recv counts, arrivals, children, fine bins, status) -- after a stream sync."""
        c = self.h_counts.numpy()
        G = self.world
        return c[:G].tolist(), c[G:2 * G].tolist(), int(c[2 * G]), int(c[2 * G + 1]), int(c[2 * G + 2]), int(c[2 * G + 3])

    def sort(self, n_arrivals, n_fine):
        self.rec, self.rec_next = self.rec_next, self.rec
        self.n_local = n_arrivals
        _lib.check(self.lib.pmmh_svsplit_sort(*self._c(), n_arrivals, n_fine, self.lag, _p(self.rec),
                                              None if self.recv_keys is None else _p(self.recv_keys),
                                              1 if self.world == 1 else 0, _p(self.xs), _p(self.perm),
                                              self._stream()), "pmmh_svsplit_sort")

    def diag(self):
        out = (ctypes.c_longlong * _lib.DIAG_COUNT)()
        _lib.check(self.lib.pmmh_svsplit_diag(*self._c(), out), "pmmh_svsplit_diag")
        return [int(v) for v in out]


class LocalComm(object):
    """All ranks live in this process on one device: exchanges are device copies."""

    def __init__(self, world):
        self.world = int(world)
        self.local_ranks = list(range(self.world))

    def all_gather(self, sends, outs):
        for o in outs:
            for r, s in enumerate(sends):
                o[r].copy_(s)

    def all_to_all(self, sends, send_counts, recvs, recv_counts, skip_self=False):
        """skip_self: the part a rank sends to itself is already in place (pmmh_svsplit_pack_direct)."""
        G = self.world
        soff = [np.concatenate([[0], np.cumsum(c)]) for c in send_counts]
        roff = [np.concatenate([[0], np.cumsum(c)]) for c in recv_counts]
        for s in range(G):
            for d in range(G):
                n = send_counts[s][d]
                assert n == recv_counts[d][s]
                if n and not (skip_self and s == d):
                    recvs[d][roff[d][s]:roff[d][s] + n].copy_(sends[s][soff[s][d]:soff[s][d] + n])

    def all_to_all_async(self, sends, send_counts, recvs, recv_counts, skip_self=False):
        self.all_to_all(sends, send_counts, recvs, recv_counts, skip_self)
        return None

    def all_reduce_sum(self, tensors):
        tot = torch.stack(tensors).sum(dim=0) if len(tensors) > 1 else tensors[0]
        for t in tensors:
            t.copy_(tot)

    def agree_status(self, status, device):
        """The largest status word over all ranks (every rank acts on the same value)."""
        return int(status)


class DistComm(object):
    """One rank per process: torch.distributed (NCCL on GPUs; gloo works for the host logic)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.local_ranks = [dist.get_rank(group)]

    def all_gather(self, sends, outs):
        self.dist.all_gather_into_tensor(outs[0].view(-1), sends[0].view(-1), group=self.group)

    def _views(self, buf, counts, skip):
        """One view of ``buf`` per peer at the cumulative offsets of ``counts``; an empty one for ``skip``."""
        out, off = [], 0
        for r, n in enumerate(counts):
            n = int(n)
            out.append(buf[off:off] if r == skip else buf[off:off + n])
            off += n
        return out

    def _exchange_views(self, send, send_counts, recv, recv_counts, async_op=False):
        """Everything but the self part: NCCL = one grouped send / recv per peer (all_to_all over views);
        other backends (gloo, CPU tests) = point-to-point pairs."""
        me = self.local_ranks[0]
        outs, ins = self._views(recv, recv_counts, me), self._views(send, send_counts, me)
        if self.dist.get_backend(self.group) == "nccl":
            return self.dist.all_to_all(outs, ins, group=self.group, async_op=async_op)
        ops = []
        for r in range(self.world):
            if r != me and ins[r].numel():
                ops.append(self.dist.P2POp(self.dist.isend, ins[r].contiguous(), r, self.group))
            if r != me and outs[r].numel():
                ops.append(self.dist.P2POp(self.dist.irecv, outs[r], r, self.group))
        for w in (self.dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None

    def all_to_all(self, sends, send_counts, recvs, recv_counts, skip_self=False):
        if skip_self:
            # the self part is already in place (pmmh_svsplit_pack_direct): grouped send / recv per peer over views at
            # the unchanged offsets, nothing for this rank itself
            self._exchange_views(sends[0], send_counts[0], recvs[0], recv_counts[0])
            return
        nrecv, nsend = int(sum(recv_counts[0])), int(sum(send_counts[0]))
        self.dist.all_to_all_single(recvs[0][:nrecv], sends[0][:nsend], list(recv_counts[0]),
                                    list(send_counts[0]), group=self.group)

    def all_to_all_async(self, sends, send_counts, recvs, recv_counts, skip_self=False):
        """Enqueued behind the current stream's work; later kernels on the current stream overlap
        it until ``handle.wait()`` (which makes the current stream wait, not the host)."""
        if skip_self:
            return self._exchange_views(sends[0], send_counts[0], recvs[0], recv_counts[0], async_op=True)
        nrecv, nsend = int(sum(recv_counts[0])), int(sum(send_counts[0]))
        return self.dist.all_to_all_single(recvs[0][:nrecv], sends[0][:nsend], list(recv_counts[0]),
                                           list(send_counts[0]), group=self.group, async_op=True)

    def all_reduce_sum(self, tensors):
        self.dist.all_reduce(tensors[0], op=self.dist.ReduceOp.SUM, group=self.group)

    def agree_status(self, status, device):
        """The largest status word over all ranks: cap overflows (status 4 / 8) are rank-local, and a
        rank that raised alone would leave the others blocked in the next collective."""
        t = torch.tensor([int(status)], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(t.item())


def run_split_smoother(comm, obs, params, n_total, lag, rvr_d, u_d=None, philox=None, device=None,
                       cap=None, capc=None, keep_history=False):
    """Runs one evaluation.  obs [n_obs] and params [4] are host arrays; rvr_d [n_obs] device
    uniforms (Phi already applied); u_d [n_obs, N] device tensor (every rank holds all of it) or
    None with ``philox`` = (seed, offset).  Returns a dict of device tensors (identical on every
    rank) plus per-rank diagnostics; ``keep_history`` also returns the sorted local generations
    (tests)."""
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    n_obs = obs.shape[0]
    G = comm.world
    device = device or torch.device("cuda", torch.cuda.current_device())
    if lag and n_obs < 2 * lag:
        raise _lib.PmmhError("n_obs must be at least 2 * lag")
    ranks = [_Rank(r, G, n_total, n_obs, lag, device, cap, capc) for r in comm.local_ranks]
    obs_d = torch.from_numpy(obs).to(device)
    params_d = torch.from_numpy(params).to(device)
    seed, offset = (0, 0) if philox is None else (int(philox[0]), int(philox[1]))
    if u_d is None and philox is None:
        raise _lib.PmmhError("either u or a Philox stream is needed")
    hist = [] if keep_history else None
    nc_hist = []

    for rk in ranks:
        rk.init(params_d)
        rk.weights(0, obs_d, params_d)
    comm.all_gather([rk.gather_send for rk in ranks], [rk.gather for rk in ranks])
    if keep_history:
        hist.append([rk.xs[:rk.n_local].clone() for rk in ranks])
    for t in range(1, n_obs):
        for rk in ranks:
            rk.gather_hist[t - 1].copy_(rk.gather)
            rk.children(t, obs_d, params_d, rvr_d, u_d, seed, offset)
        comm.all_gather([rk.hist_send for rk in ranks], [rk.hist for rk in ranks])
        for rk in ranks:
            rk.plan_and_pack()
        torch.cuda.current_stream(device).synchronize()      # the counts are on the host now
        cnt = [rk.counts() for rk in ranks]
        nc_hist.append([c[3] for c in cnt])
        status = max(int(c[5]) for c in cnt)
        if G > 1:
            status = comm.agree_status(status, device)     # the same decision on every rank
        if status != 0:
            raise FloatingPointError("split particle filter abandoned at t=%d (largest status over the ranks %d; "
                                     "local %s: 1/2 non-finite weights, 4 children > cap_children, 8 arrivals > "
                                     "cap_particles)" % (t, status, [int(c[5]) for c in cnt]))
        if G == 1:
            rk = ranks[0]
            rk.rec_next, rk.send = rk.send, rk.rec_next       # the exchange is the identity
            pending = None
        elif ranks[0].send_keys is not None:
            # values first (8 B per particle), then the records asynchronously: the sort needs only
            # the values and runs while the records are in flight
            comm.all_to_all([rk.send_keys for rk in ranks], [c[0] for c in cnt],
                            [rk.recv_keys for rk in ranks], [c[1] for c in cnt], skip_self=True)
            pending = comm.all_to_all_async([rk.send for rk in ranks], [c[0] for c in cnt],
                                            [rk.rec_next for rk in ranks], [c[1] for c in cnt], skip_self=True)
        else:
            comm.all_to_all([rk.send for rk in ranks], [c[0] for c in cnt],
                            [rk.rec_next for rk in ranks], [c[1] for c in cnt], skip_self=True)
            pending = None
        for rk, c in zip(ranks, cnt):
            rk.sort(c[2], c[4])
        if pending is not None:
            pending.wait()
        for rk in ranks:
            rk.weights(t, obs_d, params_d)
        comm.all_gather([rk.gather_send for rk in ranks], [rk.gather for rk in ranks])
        if keep_history:
            hist.append([rk.xs[:rk.n_local].clone() for rk in ranks])
    for rk in ranks:
        rk.gather_hist[n_obs - 1].copy_(rk.gather)

    # ---- O(T) sums over ranks, tail terms, outputs ---------------------------------------
    comm.all_reduce_sum([rk.sums for rk in ranks])
    tails = [None] * len(ranks)
    if lag:
        counts_hist = ranks[0].gather_hist[:, :, 1].cpu().numpy().astype(np.int64)   # [n_obs][G]
        n_final = counts_hist[n_obs - 1]
        w_lag = [torch.zeros((lag, rk.cap), dtype=_F64, device=device) for rk in ranks]
        w_fin = []
        scratch = [torch.empty(rk.cap, dtype=_F64, device=device) for rk in ranks]
        for irel in range(lag):
            i = n_obs - lag + irel
            for k, rk in enumerate(ranks):
                _lib.check(rk.lib.pmmh_svsplit_normalise(_p(rk.sh_keep[i % lag]), rk.n_keep[i],
                                                         _p(rk.sums[i]), _p(scratch[k]), rk._stream()),
                           "pmmh_svsplit_normalise")
            if irel == lag - 1:
                w_fin = [s.clone() for s in scratch]
                break
            plans = [interval_exchange_counts(counts_hist[i], n_final, rk.rank) for rk in ranks]
            if G == 1:
                w_lag[0][irel, :n_final[0]].copy_(scratch[0][:n_final[0]])
            else:
                comm.all_to_all([s for s in scratch], [p[0] for p in plans],
                                [w[irel] for w in w_lag], [p[1] for p in plans])
        tails = [torch.zeros((lag, 8), dtype=_F64, device=device) for _ in ranks]
        for k, rk in enumerate(ranks):
            _lib.check(rk.lib.pmmh_svsplit_tail(*rk._c(), rk.n_local, lag, n_obs, _p(obs_d), _p(params_d),
                                                _p(rk.perm), _p(rk.rec), _p(w_fin[k]), _p(w_lag[k]),
                                                rk.cap, _p(tails[k]), rk._stream()), "pmmh_svsplit_tail")
        comm.all_reduce_sum(tails)
    outs = []
    for k, rk in enumerate(ranks):
        o = {"log_like": torch.zeros(1, dtype=_F64, device=device),
             "filt": torch.zeros(n_obs, dtype=_F64, device=device),
             "smo": torch.zeros(n_obs, dtype=_F64, device=device),
             "gradient": torch.zeros((4, n_obs), dtype=_F64, device=device),
             "traj": torch.zeros(n_obs, dtype=_F64, device=device)}
        _lib.check(rk.lib.pmmh_svsplit_finish(_p(rk.sums), _p(rk.shift), _p(rk.xmin),
                                              None if tails[k] is None else _p(tails[k]), _p(rk.gather), G,
                                              _p(params_d), n_obs, lag, int(n_total), _p(o["log_like"]),
                                              _p(o["filt"]), _p(o["smo"]), _p(o["gradient"]), _p(o["traj"]),
                                              rk._stream()), "pmmh_svsplit_finish")
        o["diag"] = rk.diag()
        o["n_local"] = rk.n_local
        outs.append(o)
    res = dict(outs[0])
    res["per_rank"] = outs
    res["counts"] = ranks[0].gather_hist[:, :, 1].cpu().numpy().astype(np.int64)
    res["children"] = np.array(nc_hist, dtype=np.int64)     # [T][local ranks] children generated per step
    if keep_history:
        res["x_hist"] = hist
    return res


class SplitParticleMethodsCUDA(BaseStateInference):
    """Estimator with the reference's contract (state/particle_methods/cython.py:30-168) whose
    particle system is split over the ranks of ``group`` (or over ``local_world`` in-process
    ranks on one GPU).  ``rvs={'rvs': x}``: a ``DeviceRVS`` handle (u resident on every rank) or
    a ``PhiloxRVS``.  Log-likelihood, filter / smoother means and the gradient; the Hessian
    branch is not split (use ParticleMethodsCUDA)."""

    def __init__(self, model, no_particles, fixed_lag=10, group=None, local_world=None, device=None,
                 cap_particles=None, cap_children=None):
        if model.short_name != 'sv':
            raise NameError("CUDA implementation for model missing.")
        if not torch.cuda.is_available():
            raise RuntimeError("SplitParticleMethodsCUDA needs a CUDA device; there is no CPU fallback.")
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.comm = LocalComm(local_world) if local_world else DistComm(group)
        self.alg_type = 'particle'
        self.name = "Split particle method (CUDA) for " + model.short_name + " model"
        self.no_obs = model.no_obs + 1
        self.no_particles = int(no_particles)
        self.dim_rvs = (self.no_obs, self.no_particles + 1)
        self.settings = {'no_particles': self.no_particles, 'no_obs': self.no_obs,
                         'resampling_method': 'systematic', 'fixed_lag': int(fixed_lag),
                         'initial_state': 0.0, 'generate_initial_state': True,
                         'estimate_gradient': True, 'estimate_hessian': False}
        self.cap = (cap_particles, cap_children)
        self.results = {}
        self.diagnostics = {}

    def _run(self, model, lag, kwargs):
        from ...parameter.rvs import DeviceRVS
        obs = np.array(model.obs.flatten()).astype(float)
        params = np.asarray(model.get_all_params(), dtype=np.float64)
        rvs = kwargs['rvs']['rvs']
        if isinstance(rvs, PhiloxRVS):
            rvr = K.norm_cdf(rvs.resampling_normals(self.no_obs, self.no_particles, self.device))
            u, philox = None, (rvs.seed, rvs.offset)
        elif isinstance(rvs, DeviceRVS):
            rvr, u, philox = K.norm_cdf(rvs.tensors['r_raw']), rvs.tensors['u'], None
        else:
            raise TypeError("rvs must be a DeviceRVS or PhiloxRVS handle")
        out = run_split_smoother(self.comm, obs, params, self.no_particles, lag, rvr, u, philox,
                                 self.device, self.cap[0], self.cap[1])
        self.diagnostics = {'near_ties': sum(o['diag'][_lib.DIAG_NEAR_TIES] for o in out['per_rank']),
                            'status': max(o['diag'][_lib.DIAG_STATUS] for o in out['per_rank'])}
        return out

    def filter(self, model, **kwargs):
        try:
            out = self._run(model, 0, kwargs)
            self.results.update({'filt_state_est': out['filt'].cpu().numpy(),
                                 'state_trajectory': out['traj'].cpu().numpy(),
                                 'log_like': float(out['log_like'].item())})
            return bool(np.isfinite(self.results['log_like']))
        except Exception as e:
            print("Error in CUDA code for split particle filter.")
            print(e)
            return False

    def smoother(self, model, **kwargs):
        try:
            out = self._run(model, self.settings['fixed_lag'], kwargs)
            grad = out['gradient'].cpu().numpy()
            grad[~np.isfinite(grad)] = 0.0            # cython.py:104-106
            self.results.update({'filt_state_est': out['filt'].cpu().numpy(),
                                 'state_trajectory': out['traj'].cpu().numpy(),
                                 'smo_state_est': out['smo'].cpu().numpy(),
                                 'log_like': float(out['log_like'].item())})
            if model.using_gradients:
                self.results.update({'log_joint_gradient_estimate': np.nansum(grad, axis=1)})
            if not np.isfinite(self.results['log_like']):
                return False
            return bool(self._estimate_gradient_and_hessian(model))
        except Exception as e:
            print("Error in CUDA code for split particle smoother.")
            print(e)
            return False

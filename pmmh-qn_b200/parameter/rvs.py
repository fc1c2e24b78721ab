"""Device-resident auxiliary random variables u and their Crank-Nicolson proposal.

Replaces ``MarkovChainMonteCarlo._propose_rvs`` (/root/reference/python/parameter/mcmc/
base_class.py:221-241) for large problems: at T=1000, N=2^20 one u array is 8.4 GB, so it must
live in HBM, be updated there and never be deep-copied by the sampler's state history
(mh_quasi_newton.py:232 calls ``copy.deepcopy`` on every state).

``DeviceRVS`` is what the CUDA estimators accept in ``rvs={'rvs': ...}`` next to plain NumPy
arrays.  For the particle estimators it holds u already split the way the kernels read it:
``r_raw`` [n_obs] (the first n_obs flat entries of the reference's (n_obs, N+1) array) and
``u`` [n_obs, N] time-major.  ``copy.deepcopy`` of a handle is a reference copy: handles are
treated as immutable, every proposal allocates a new one.
"""
import numpy as np
import torch

from .. import kernels as K


class DeviceRVS(object):
    __slots__ = ("_tensors", "shape", "kind", "_owner", "_generation")

    def __init__(self, tensors, shape, kind, owner=None, generation=0):
        self._tensors = tensors     # dict name -> CUDA tensor
        self.shape = tuple(shape) if not np.isscalar(shape) else (int(shape),)
        self.kind = kind            # 'particle' | 'importance' | 'direct'
        self._owner = owner         # CorrelatedRVSState whose slot this handle views (None: owns its tensors)
        self._generation = generation

    @property
    def stale(self):
        """True for a handle into a CorrelatedRVSState slot that has been written again since."""
        return self._owner is not None and self._owner._slot_generation(self) != self._generation

    @property
    def tensors(self):
        if self.stale:
            raise RuntimeError("stale DeviceRVS handle: its slot of the CorrelatedRVSState has been reused "
                               "(only the current and the proposed state are kept on the device)")
        return self._tensors

    # handles are immutable: history copies share the device buffers
    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors.values())

    @classmethod
    def from_numpy_particle(cls, rvs, device, non_blocking=True):
        """(n_obs, N+1) host array in the reference layout -> handle (H2D + layout kernel)."""
        rvs = np.ascontiguousarray(rvs, dtype=np.float64)
        n_obs, n1 = rvs.shape
        d = torch.from_numpy(rvs).to(device, non_blocking=non_blocking)
        r_raw, u = K.split_rvs(d, n_obs, n1 - 1)
        return cls({"r_raw": r_raw[0], "u": u[0]}, rvs.shape, "particle")

    @classmethod
    def randn_particle(cls, n_obs, n_particles, device, seed=0):
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        r_raw = torch.randn((n_obs,), dtype=torch.float64, device=device, generator=g)
        u = torch.randn((n_obs, n_particles), dtype=torch.float64, device=device, generator=g)
        return cls({"r_raw": r_raw, "u": u}, (n_obs, n_particles + 1), "particle")

    @classmethod
    def from_numpy_flat(cls, rvs, device, kind):
        rvs = np.ascontiguousarray(rvs, dtype=np.float64)
        d = torch.from_numpy(rvs.reshape(-1)).to(device, non_blocking=True)
        return cls({"u": d}, rvs.shape, kind)

    def to_numpy_particle(self):
        """Back to the reference's (n_obs, N+1) host layout (tests / debugging)."""
        r = self.tensors["r_raw"].cpu().numpy()
        u = self.tensors["u"].cpu().numpy()
        n_obs, n = u.shape
        flat = np.concatenate([r, np.ascontiguousarray(u.T).reshape(-1)])
        return flat.reshape(n_obs, n + 1)


def propose_rvs(current, sigma_u, xi=None, seed=0, philox_offset=0):
    """Crank-Nicolson proposal u' = sqrt(1 - sigma_u^2) u + sigma_u xi on the device
    (base_class.py:231-233).  ``xi`` may be a DeviceRVS with the same structure (parity tests)
    or None: standard normals are then drawn on the device (Philox4x32-10, counter based).
    Returns a new DeviceRVS."""
    out = {}
    offset = int(philox_offset)
    for name in sorted(current.tensors):
        t = current.tensors[name]
        x = None if xi is None else xi.tensors[name]
        out[name] = K.crank_nicolson(t, sigma_u, xi=x, seed=seed, philox_offset=offset)
        offset += (t.numel() + 1) // 2
    return DeviceRVS(out, current.shape, current.kind)


class CorrelatedRVSState(object):
    """Device-resident state of the correlated pseudo-marginal chain for ONE estimator: the auxiliary
    variables u of the current state and of the proposal in two preallocated slots (8.4 GB each at
    T = 1000, N = 2^20), the Crank-Nicolson proposal drawn on the device from a counter-based Philox
    stream, accept / reject as a swap of the two slot indices.

    Mirrors what the reference's sampler does with ``state['rvs']``
    (/root/reference/python/parameter/mcmc/base_class.py:221-241 propose, :269-300 accept / reject with
    ``copy.deepcopy``, :290-300 ``_delete_corr_rvs_history``; mh_quasi_newton.py:226-233 keeps
    ``memory_length`` states, whose rvs it never reads again) without ever moving u: handles given to
    the state history are views of a slot (deep copies are reference copies) and turn *stale* when the
    slot is written again; reading a stale handle raises instead of returning another state's numbers.

        st = CorrelatedRVSState.randn_particle(n_obs, N, device, sigma_u=0.05, seed=1)
        cur = st.current                      # DeviceRVS for estimator.smoother(model, rvs={'rvs': cur})
        prop = st.propose()                   # CN into the spare slot (one pass over u, no allocation)
        ... estimator.smoother(model, rvs={'rvs': prop}) ...
        st.accept()  or  st.reject()
    """

    def __init__(self, tensors, shape, kind, sigma_u, seed=0):
        self.shape, self.kind = shape, kind
        self.sigma_u = float(sigma_u)
        self.seed = int(seed)
        self._slots = [tensors, {k: torch.empty_like(v) for k, v in tensors.items()}]
        self._gen = [1, 0]
        self._cur = 0
        self._have_proposal = False
        self._offset = 0          # Philox counters consumed so far (two normals per counter)
        self.proposals = 0
        self.accepted = 0

    @classmethod
    def randn_particle(cls, n_obs, n_particles, device, sigma_u, seed=0):
        h = DeviceRVS.randn_particle(n_obs, n_particles, device, seed)
        return cls(h._tensors, h.shape, "particle", sigma_u, seed)

    @classmethod
    def from_numpy_particle(cls, rvs, device, sigma_u, seed=0):
        h = DeviceRVS.from_numpy_particle(rvs, device)
        return cls(h._tensors, h.shape, "particle", sigma_u, seed)

    def _slot_generation(self, handle):
        for k in (0, 1):
            if self._slots[k] is handle._tensors:
                return self._gen[k]
        return -1

    def _handle(self, k):
        return DeviceRVS(self._slots[k], self.shape, self.kind, owner=self, generation=self._gen[k])

    @property
    def current(self):
        return self._handle(self._cur)

    @property
    def nbytes(self):
        return 2 * sum(t.numel() * t.element_size() for t in self._slots[0].values())

    def propose(self, xi=None):
        """u' = sqrt(1 - sigma_u^2) u + sigma_u xi into the spare slot (base_class.py:231-233).  xi: a
        DeviceRVS of the same structure (parity tests) or None = Philox normals drawn on the device."""
        spare = 1 - self._cur
        self._gen[spare] = max(self._gen) + 1
        for name in sorted(self._slots[self._cur]):
            t = self._slots[self._cur][name]
            x = None if xi is None else xi.tensors[name]
            K.crank_nicolson(t, self.sigma_u, xi=x, seed=self.seed, philox_offset=self._offset,
                             out=self._slots[spare][name])
            self._offset += (t.numel() + 1) // 2
        self._have_proposal = True
        self.proposals += 1
        return self._handle(spare)

    def accept(self):
        if not self._have_proposal:
            raise RuntimeError("accept() without a proposal")
        self._cur = 1 - self._cur
        self._have_proposal = False
        self.accepted += 1

    def reject(self):
        self._have_proposal = False

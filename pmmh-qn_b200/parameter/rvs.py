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
    __slots__ = ("tensors", "shape", "kind")

    def __init__(self, tensors, shape, kind):
        self.tensors = tensors      # dict name -> CUDA tensor
        self.shape = tuple(shape) if not np.isscalar(shape) else (int(shape),)
        self.kind = kind            # 'particle' | 'importance' | 'direct'

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

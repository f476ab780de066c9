"""Host logic around the ensemble: synthetic initial conditions, trajectory sharding, spectrum assembly.

* sharding: trajectory itrj -> rank itrj mod G; replaces `pqcxms` farming one process per TMP.<n> directory
  (reference bin/pqcxms:88-98).  No inter-GPU traffic during MD.
* the only collective: one all-reduce(sum) of the fragment-mass histogram, in place of concatenating the
  per-directory qcxms.res files (reference bin/pqcxms:101-103).
* synthetic initial conditions follow SURVEY.md 8(d) (no Fortran set-up run is available offline).
"""
import numpy as np

from .api import AMUTOAU, AUTOEV, FSTOAU, KB

# NIST masses (amu) of the elements of the benchmark molecules, reference src/atomic_masses.f90:25-31
ATOMIC_MASS_AMU = {1: 1.00794075, 2: 4.00260193, 6: 12.01073590, 7: 14.00670321, 8: 15.99940492, 9: 18.99840316,
                   16: 32.06478741, 17: 35.45293758, 18: 39.94779856}


def masses_au(num):
    """setmass(): atomic masses in electron masses (reference src/mass.f90:14-19)."""
    return np.array([ATOMIC_MASS_AMU[int(z)] * AMUTOAU for z in num])


def shard_indices(ntraj, world_size, rank):
    """Static partition of trajectory ids over ranks (itrj mod G)."""
    return np.arange(rank, ntraj, world_size)


def synthetic_initial_conditions(num, xyz_eq, ntraj, first_id=0, temperature=500.0, sigma=0.05, ieeatm=0.6, tadd_fs=400.0, ids=None):
    """Per-trajectory start geometry/velocities/IEE: counter-based RNG seeded by the GLOBAL trajectory id, so a
    trajectory gets the same initial conditions whichever rank runs it.  ids: explicit global ids (a rank's shard,
    see shard_indices); default first_id .. first_id + ntraj - 1."""
    nat = len(num)
    mass = masses_au(num)
    ids = np.arange(first_id, first_id + ntraj) if ids is None else np.asarray(ids)
    ntraj = len(ids)
    xyz = np.empty((ntraj, nat, 3)); velo = np.empty((ntraj, nat, 3)); eimp = np.empty(ntraj)
    for k in range(ntraj):
        rng = np.random.Generator(np.random.Philox(key=0x5EED0000 + int(ids[k])))
        xyz[k] = xyz_eq + sigma * rng.standard_normal((nat, 3))
        sign = np.where(rng.random((nat, 3)) < 0.5, -1.0, 1.0)
        velo[k] = sign * np.sqrt(KB * temperature / mass)[:, None]       # mdinitu rule, reference src/mdinit.f90:10-52
        ev = ieeatm * nat * np.exp(0.3 * rng.standard_normal())
        eimp[k] = min(max(ev, 1.0), 60.0) / AUTOEV
    velof = np.ones((ntraj, nat))
    tadd = np.full(ntraj, tadd_fs * FSTOAU)
    return dict(xyz=xyz, velo=velo, velof=velof, eimp=eimp, tadd=tadd, mass=mass)


def allreduce_histogram(bins):
    """Sum the per-rank fragment histograms (torch tensor, any device) across the process group in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bins, op=dist.ReduceOp.SUM)
    return bins


def make_comm(device=0):
    """The C-ABI communicator (qcxms_b200_comm_*, NCCL) for this process: bootstrapped through torch.distributed when a process
    group exists (rank 0's unique id is broadcast), a one-rank communicator otherwise."""
    from .api import Comm
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [Comm.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return Comm(box[0], dist.get_world_size(), dist.get_rank(), device)
    return Comm(Comm.unique_id(), 1, 0, device)


def spectrum_from_histogram(bins):
    """Normalised stick spectrum (base peak = 100) from the summed fragment-mass histogram."""
    bins = np.asarray(bins, dtype=np.float64)
    top = bins.max()
    return 100.0 * bins / top if top > 0 else bins


def cosine_similarity(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(a) * np.linalg.norm(b)
    return float(a @ b / den) if den > 0 else 0.0

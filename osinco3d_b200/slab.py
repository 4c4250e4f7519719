"""Host-side z-slab decomposition helpers (the reference has no parallelism,
README.md:61).  Mirrors what libo3d_b200 does internally (o3d_slab_partition, comm.cu) so that
drivers, bench.py and the tests can scatter / gather global fields and reason about halos.
Pure NumPy: usable with any launcher (torch.distributed over NCCL or gloo, MPI)."""
import numpy as np

R = 3   # ghost planes per side = widest stencil radius (6th-order first derivative)


def slab_range(nz, rank, nranks):
    """(z0, nz_local): contiguous slabs, remainder planes to the low ranks"""
    q, r = divmod(nz, nranks)
    return rank * q + min(rank, r), q + (1 if rank < r else 0)


def neighbours(rank, nranks, periodic_z):
    """(down, up) ranks of a slab (-1: a physical wall closes that side)"""
    up = rank + 1 if rank + 1 < nranks else (0 if periodic_z else -1)
    dn = rank - 1 if rank > 0 else (nranks - 1 if periodic_z else -1)
    if nranks == 1:
        return -1, -1
    return dn, up


def take_slab(a, rank, nranks):
    """this rank's planes of a global (nx, ny, nz) array, Fortran-ordered"""
    z0, nzl = slab_range(a.shape[2], rank, nranks)
    return np.asfortranarray(a[:, :, z0:z0 + nzl])


def halo_plan(rank, nranks, nz_local, width, periodic_z):
    """The exchange comm.cu performs before a z-stencil: list of
    (peer, send_planes, recv_ghost) with plane ranges in LOCAL indices; ghost planes are
    addressed as negative indices (below) or >= nz_local (above)."""
    dn, up = neighbours(rank, nranks, periodic_z)
    plan = []
    if up >= 0:
        plan.append((up, (nz_local - width, nz_local), (nz_local, nz_local + width)))
    if dn >= 0:
        plan.append((dn, (0, width), (-width, 0)))
    return plan


def wall_ghosts(slab, side, width, odd):
    """free-slip closure of src/derivation.f90 (*p_11 even / *i_11 odd) as ghost planes of a slab
    that owns a wall: f(-g) = +-f(g)"""
    if side == "lo":
        g = slab[:, :, 1:width + 1][:, :, ::-1]
    else:
        n = slab.shape[2]
        g = slab[:, :, n - 1 - width:n - 1][:, :, ::-1]
    return -g if odd else g.copy()

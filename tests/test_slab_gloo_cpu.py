"""N > 1 host logic on CPU: world_size-2 `gloo` processes exercise the z-slab decomposition the
multi-GPU path uses (osinco3d_b200/slab.py == o3d_slab_partition / comm.cu's exchange pattern):
partition, halo exchange with periodic wrap links and free-slip wall closures, the all-reduced
SOR residual, and bench.py's max-over-ranks timing reduction.  No CUDA is touched."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_python_partition_mirrors_the_library(built_lib):
    from osinco3d_b200 import slab
    lib = built_lib.lib()
    for nz in (7, 81, 129, 256, 1024):
        for nr in (1, 2, 3, 4, 8):
            tot = 0
            for r in range(nr):
                z0, nzl = C.c_int(), C.c_int()
                assert lib.o3d_slab_partition(nz, nr, r, C.byref(z0), C.byref(nzl)) == 0
                assert (z0.value, nzl.value) == slab.slab_range(nz, r, nr)
                assert z0.value == tot
                tot += nzl.value
            assert tot == nz
    assert lib.o3d_slab_partition(10, 2, 2, None, None) != 0


def _worker(rank, world, port, periodic, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from osinco3d_b200 import slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz, W = 9, 8, 23, slab.R
        rng = np.random.default_rng(7)
        glob = np.asfortranarray(rng.standard_normal((nx, ny, nz)))
        z0, nzl = slab.slab_range(nz, rank, world)
        mine = slab.take_slab(glob, rank, world)
        # padded slab: W ghost planes on each side, like the device layout
        pad = np.zeros((nx, ny, nzl + 2 * W))
        pad[:, :, W:W + nzl] = mine
        plan = slab.halo_plan(rank, world, nzl, W, periodic)
        reqs, bufs = [], []
        for peer, (s0, s1), (g0, g1) in plan:
            send = torch.from_numpy(np.ascontiguousarray(pad[:, :, W + s0:W + s1]))
            recv = torch.empty_like(send)
            tag_s = 0 if s0 == 0 else 1          # 0: my bottom planes, 1: my top planes
            tag_r = 1 if g0 < 0 else 0           # ghost below receives the peer's top planes
            reqs.append(dist.isend(send, peer, tag=tag_s))
            reqs.append(dist.irecv(recv, peer, tag=tag_r))
            bufs.append((recv, g0, g1))
        for r in reqs:
            r.wait()
        for recv, g0, g1 in bufs:
            pad[:, :, W + g0:W + g1] = recv.numpy()
        dn, up = slab.neighbours(rank, world, periodic)
        odd = True
        if dn < 0:
            pad[:, :, 0:W] = slab.wall_ghosts(mine, "lo", W, odd)
        if up < 0:
            pad[:, :, W + nzl:] = slab.wall_ghosts(mine, "hi", W, odd)
        # expected ghosts straight from the global field and the closure rule (o3d_common.cuh)
        for g in range(1, W + 1):
            for side, q_glob in (("lo", z0 - g), ("hi", z0 + nzl - 1 + g)):
                if 0 <= q_glob < nz:
                    exp = glob[:, :, q_glob]
                elif periodic:
                    exp = glob[:, :, q_glob % nz]
                else:
                    src = -q_glob if q_glob < 0 else 2 * (nz - 1) - q_glob
                    exp = -glob[:, :, src]
                got = pad[:, :, W - g] if side == "lo" else pad[:, :, W + nzl - 1 + g]
                assert np.array_equal(got, exp), (rank, side, g)
        # SOR residual: every rank must take the same exit decision (max all-reduce on the bits)
        local = float(np.max(np.abs(mine)))
        t = torch.tensor([np.float64(local).view(np.int64)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gmax = np.int64(t.item()).view(np.float64)
        assert gmax == np.max(np.abs(glob))
        # bench.py timing: max over ranks
        ms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert ms.item() == 10.0 + world - 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_halo_exchange_world_size_2_gloo(periodic):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if periodic else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, periodic, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res

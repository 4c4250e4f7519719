#!/usr/bin/env python
"""Multi-GPU parity check of the z-slab decomposition, run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank holds a z-slab; halo planes travel with ncclSend/ncclRecv and the SOR residual with
ncclAllReduce(max).  Rank 0 additionally runs the SAME problem in a single-GPU session on its
own device.  Per point the arithmetic is identical and the halo planes carry the neighbour's
values verbatim, so the P-GPU fields must equal the 1-GPU fields BIT FOR BIT -- stencil kernels
and red-black SOR iterates alike (identical dmax -> identical exit decisions on every rank).
Prints one line per case and exits non-zero on any mismatch.  tests/test_gpu_multi.py wraps it.
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tgv_like(n, nz, d, bc):
    """TGV plus a non-symmetric perturbation and uz != 0, so that every halo plane and parity
    matters (a pure TGV has uz = 0 and many symmetric planes)"""
    x = (d * np.arange(n))[:, None, None]
    y = (d * np.arange(n))[None, :, None]
    z = (d * np.arange(nz))[None, None, :]
    if bc[2] == 1:
        Lz = d * (nz - 1)
        kz = np.pi / Lz
        ux = np.sin(x) * np.cos(y) * np.cos(kz * z) + 0.1 * np.sin(2 * x) * np.cos(3 * y) * np.cos(2 * kz * z)
        uy = -np.cos(x) * np.sin(y) * np.cos(kz * z) + 0.05 * np.cos(x) * np.sin(2 * y) * np.cos(3 * kz * z)
        uz = 0.2 * np.cos(x) * np.cos(y) * np.sin(kz * z)
        pp = 0.0625 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * kz * z) + 2.0)
        phi = 0.5 + 0.4 * np.cos(x) * np.cos(y) * np.cos(kz * z)
    else:
        Lz = d * nz
        kz = 2 * np.pi / Lz
        ux = np.sin(x) * np.cos(y) * np.cos(kz * z) + 0.1 * np.sin(2 * x + 0.3) * np.cos(kz * z + 0.2)
        uy = -np.cos(x) * np.sin(y) * np.cos(kz * z) + 0.05 * np.sin(y + 2 * kz * z)
        uz = 0.2 * np.cos(x + 0.1) * np.cos(y) * np.sin(kz * z + 0.4)
        pp = 0.0625 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * kz * z) + 2.0)
        phi = 0.5 + 0.4 * np.cos(x) * np.cos(y) * np.cos(kz * z + 0.3)
    f = np.asfortranarray
    return f(ux + 0 * z), f(uy + 0 * z), f(uz + 0 * z), f(pp + 0 * z), f(phi + 0 * z)


def main():
    import torch
    import torch.distributed as dist
    import osinco3d_b200 as o3d
    from osinco3d_b200 import slab
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    o3d._lib.check(o3d.lib().o3d_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = 0

    def make_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(o3d.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    cases = [
        # name, n, nz, bc, iles, nscr, eps, idyn, steps [, multigrid]
        ("freeslip_dns_fusedSOR", 48, 24 * world, (1, 1, 1), 0, 0, 1e-7, 0, 4),
        ("freeslip_les_scalar_dynomega", 40, 16 * world + 1, (1, 1, 1), 1, 1, 1e-6, 1, 4),
        ("periodic_even_fusedSOR", 32, 16 * world, (0, 0, 0), 0, 1, 1e-7, 0, 3),
        ("periodic_odd_seamSOR", 33, 16 * world + 1, (0, 0, 0), 1, 0, 1e-6, 0, 3),
        ("mixed_0011_ragged", 33, 16 * world + 3, (0, 1, 0), 0, 1, 1e-6, 1, 3),
        # slabs of >= 32 planes: kernels are split into interior + boundary launches and the halo
        # exchange runs on the communication stream underneath the interior launch
        ("freeslip_overlap_les_scalar", 40, 40 * world + 1, (1, 1, 1), 1, 1, 1e-6, 1, 4),
        ("periodic_overlap_fusedSOR", 32, 36 * world, (0, 0, 0), 0, 0, 1e-7, 0, 3),
        ("mixed_0011_overlap_seamSOR", 33, 34 * world + 1, (0, 1, 0), 0, 0, 1e-6, 0, 3),
        # multigrid V-cycles (src/integration.f90:244, multigrid = 1): level 0 distributed in z,
        # coarser levels replicated; nested (even periodic / odd mirrored) and non-nested extents
        ("freeslip_multigrid", 33, 16 * world + 1, (1, 1, 1), 0, 0, 1e-8, 0, 3, 1),
        ("periodic_even_multigrid", 32, 16 * world, (0, 0, 0), 0, 1, 1e-8, 0, 3, 1),
        ("periodic_odd_multigrid_ragged", 33, 16 * world + 3, (0, 0, 0), 1, 0, 1e-8, 0, 3, 1),
        ("mixed_0011_multigrid", 40, 20 * world + 1, (0, 1, 0), 0, 0, 1e-8, 0, 3, 1),
        # nbcz differs from the Poisson variant's z rule (src/initialization.f90:283-301 picks the
        # variant from the x / y flags only): _0000 / _0011 wrap z across the ranks although the
        # stencils mirror it, _111111 mirrors z on the end ranks although the stencils wrap it
        ("zrule_0000_freeslipz_SOR", 32, 16 * world, (0, 0, 1), 0, 0, 1e-7, 0, 3),
        ("zrule_0011_freeslipz_SOR", 33, 16 * world + 1, (0, 1, 1), 0, 1, 1e-6, 1, 3),
        ("zrule_111111_periodicz_SOR", 33, 16 * world + 2, (1, 1, 0), 1, 0, 1e-6, 0, 3),
        ("zrule_0000_freeslipz_multigrid", 32, 16 * world, (0, 0, 1), 0, 0, 1e-8, 0, 3, 1),
        ("zrule_0011_freeslipz_multigrid", 33, 16 * world + 1, (0, 1, 1), 0, 0, 1e-8, 0, 3, 1),
        ("zrule_111111_periodicz_multigrid", 33, 16 * world + 2, (1, 1, 0), 0, 0, 1e-8, 0, 3, 1),
    ]
    only = os.environ.get("O3D_MGPU_ONLY")
    for case in cases:
        name, n, nz, bc, iles, nscr, eps, idyn, steps = case[:9]
        mg = case[9] if len(case) > 9 else 0
        if only and only not in name:
            continue
        L = np.pi if bc[0] == 1 else 2 * np.pi
        d = L / (n - 1)
        fields = dict(zip(("ux", "uy", "uz", "pp", "phi"), tgv_like(n, nz, d, bc)))
        kw = dict(bc=bc, re=800.0, cs=0.17, dt=0.02 * d, itscheme=3, iles=iles, nscr=nscr,
                  omega=1.6, eps=eps, kmax=3000, idyn=idyn, multigrid=mg)
        cfg = o3d.make_config(n, n, nz, d, d, d, rank=rank, nranks=world, nccl_id=make_id(), **kw)
        ses = o3d.Session(cfg)
        assert (ses.z0, ses.nz_local) == slab.slab_range(nz, rank, world)
        ses.set(**{k: slab.take_slab(v, rank, world) for k, v in fields.items()})
        iters = [ses.step() for _ in range(steps - 1)]
        ses.old_values()
        iters.append(ses.step())
        resid = ses.calculate_residuals(0.02 * d, 1.0, 1.0)   # src/utils.f90:93-160
        diag = ses.step_diagnostics()
        stats = ses.statistics()
        red = ses.reduce("ux", o3d.RED_ABSMAX)
        out = {k: ses.download(k) for k in ("ux", "uy", "uz", "pp", "phi", "ux_pred", "rhs")}
        # shared-file output: every rank writes its own byte range (src/IOfunctions.f90:360-402,
        # src/visualization.f90:243-276); compared below with the files of the 1-GPU session
        io_dir = None
        if name in ("freeslip_les_scalar_dynomega", "mixed_0011_ragged"):
            io_dir = "/tmp/o3d_mgpu_io_%s_x%d" % (name, world)
            if rank == 0:
                shutil.rmtree(io_dir, ignore_errors=True)
                os.makedirs(io_dir)
            dist.barrier()
            xyz = [d * np.arange(m) for m in (n, n, nz)]
            ses.save_fields(io_dir + "/fields.bin", 0.5, *xyz)
            ses.write_all_data(io_dir + "/outputs", 3)
            ses.io_wait()
            dist.barrier()
        ses.close()
        ok = True
        msg = ""
        if rank == 0:
            one = o3d.Session(o3d.make_config(n, n, nz, d, d, d, **kw))
            one.set(**fields)
            it1 = [one.step() for _ in range(steps - 1)]
            one.old_values()
            it1.append(one.step())
            resid1 = one.calculate_residuals(0.02 * d, 1.0, 1.0)
            # Linf values and their indices are exact, the sums differ in association only
            if not (np.array_equal(resid[3:], resid1[3:]) and
                    np.allclose(resid[:3], resid1[:3], rtol=1e-12, atol=0)):
                ok, msg = False, "residuals %s vs single-GPU %s" % (resid, resid1)
            diag1 = one.step_diagnostics()
            for key in diag:
                a, b = np.array(diag[key]), np.array(diag1[key])
                if key.startswith("divu"):   # the mean (index 2) is a differently associated sum
                    same = np.array_equal(np.delete(a, 2), np.delete(b, 2)) and \
                        abs(a[2] - b[2]) <= 1e-12 * max(abs(b[0]), abs(b[1]))
                else:
                    same = np.array_equal(a, b)
                if not same:
                    ok, msg = False, msg + " diagnostics[%s] %s vs %s;" % (key, a, b)
            st1 = one.statistics()
            red1 = one.reduce("ux", o3d.RED_ABSMAX)
            ref = {k: one.download(k) for k in out}
            if io_dir:
                one.save_fields(io_dir + "/fields_1gpu.bin", 0.5, *xyz)
                one.write_all_data(io_dir + "/outputs_1gpu", 3)
                one.io_wait()
                pairs = [("fields.bin", "fields_1gpu.bin")] + [
                    ("outputs/" + f, "outputs_1gpu/" + f)
                    for f in sorted(os.listdir(io_dir + "/outputs_1gpu"))]
                for a, b in pairs:
                    if open(io_dir + "/" + a, "rb").read() != open(io_dir + "/" + b, "rb").read():
                        ok, msg = False, msg + " file %s differs from the 1-GPU file;" % a
                msg += " [%d output files byte-identical]" % len(pairs) if ok else ""
                shutil.rmtree(io_dir, ignore_errors=True)
            one.close()
            if it1 != iters:
                ok, msg = False, msg + " iterations %s vs single-GPU %s" % (iters, it1)
            if red1 != red:
                ok, msg = False, msg + " absmax differs"
            # sums are reduced in a different association across ranks: round-off only
            if not np.allclose(stats, st1, rtol=1e-12, atol=1e-300):
                ok, msg = False, msg + " statistics differ"
        # gather slabs on rank 0 and compare bitwise
        for k in sorted(out):
            mine = torch.from_numpy(np.ascontiguousarray(out[k].transpose(2, 1, 0))).cuda()
            if rank == 0:
                parts = [mine.cpu().numpy()]
                for r in range(1, world):
                    z0, nzl = slab.slab_range(nz, r, world)
                    buf = torch.empty((nzl, n, n), dtype=torch.float64, device="cuda")
                    dist.recv(buf, r)
                    parts.append(buf.cpu().numpy())
                full = np.concatenate(parts, axis=0).transpose(2, 1, 0)
                if k == "phi":
                    # the conservative clipping divides by three GLOBAL sums whose association
                    # differs across ranks (per-rank partial sums + ncclAllReduce): round-off only
                    if not np.allclose(full, ref[k], rtol=0, atol=1e-13):
                        ok = False
                        msg += " phi: max |d| %.3e;" % np.max(np.abs(full - ref[k]))
                elif not np.array_equal(full, ref[k]):
                    bad = np.argwhere(full != ref[k])
                    ok = False
                    msg += " %s: %d points differ, max |d| %.3e, first at %s;" % (
                        k, len(bad), np.max(np.abs(full - ref[k])), bad[0])
            else:
                dist.send(mine, 0)
        if rank == 0:
            print("[mgpu x%d] %-32s %s  SOR iters/step %s %s" % (
                world, name, "BITWISE-EQUAL to 1 GPU" if ok else "MISMATCH", iters, msg), flush=True)
            failures += 0 if ok else 1
    flag = torch.tensor([failures], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        # a rank that fails must not leave its peers blocked inside NCCL: die at once so that
        # torchrun tears the whole group down
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)

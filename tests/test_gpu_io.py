"""Field output in the reference's binary formats from device-resident state:
save_fields / read_fields (src/IOfunctions.f90:360-470) and write_binary / write_all_data
(src/visualization.f90:224-276).  The expected bytes are assembled with numpy from the oracle
(stream-access unformatted Fortran = raw little-endian values, i fastest)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PI = 3.141592653589793


def session(gpu, O, n=(24, 20, 18), bc=(1, 1, 1), nscr=1, iles=1):
    d = tuple((PI if b else 2 * PI) / (m - 1) for b, m in zip(bc, n))
    g = O.grid(*n, *d, bc)
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=1)
    rng = np.random.default_rng(5)
    uz = np.asfortranarray(uz + 0.05 * np.cos(g_x(n, d, 0)) * np.sin(2 * g_x(n, d, 2)))
    cfg = gpu.make_config(*n, *d, bc=bc, re=400.0, dt=0.01 * d[0], itscheme=3, iles=iles, cs=0.17,
                          nscr=nscr, omega=1.7, eps=1e-6, kmax=500)
    ses = gpu.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp, phi=phi)
    del rng
    return ses, g, d


def g_x(n, d, axis):
    shape = [1, 1, 1]
    shape[axis] = n[axis]
    return (d[axis] * np.arange(n[axis])).reshape(shape)


def raw(a):
    return np.asfortranarray(a).tobytes(order="F")


def test_save_fields_bytes_and_restart_round_trip(gpu, O, tmp_path):
    ses, g, d = session(gpu, O)
    n = (ses.cfg.nx, ses.cfg.ny, ses.cfg.nz)
    for _ in range(2):
        ses.step()
    x, y, z = (d[a] * np.arange(n[a]) for a in range(3))
    f = {k: ses.download(k) for k in ("ux", "uy", "uz", "pp", "phi")}
    path = str(tmp_path / "fields_000002.bin")
    ses.save_fields(path, 0.125, x, y, z)
    # the write is asynchronous: the time loop goes on and overwrites every field meanwhile
    for _ in range(2):
        ses.step()
    ses.io_wait()
    expect = (np.float64(0.125).tobytes() + np.array(n, dtype=np.int32).tobytes() + x.tobytes() +
              y.tobytes() + z.tobytes() + b"".join(raw(f[k]) for k in ("ux", "uy", "uz", "pp", "phi")))
    got = open(path, "rb").read()
    assert len(got) == len(expect) == 8 + 12 + 8 * sum(n) + 5 * 8 * n[0] * n[1] * n[2]
    assert got == expect
    after = {k: ses.download(k) for k in ("ux", "pp")}
    assert not np.array_equal(after["ux"], f["ux"])
    # restart: a fresh session that reads the file holds exactly the saved state
    cfg2 = gpu.make_config(*n, *d, bc=(1, 1, 1), re=400.0, dt=0.01 * d[0], nscr=1)
    s2 = gpu.Session(cfg2)
    t, x2, y2, z2 = s2.read_fields(path)
    assert t == 0.125 and np.array_equal(x2, x) and np.array_equal(y2, y) and np.array_equal(z2, z)
    for k in f:
        assert np.array_equal(s2.download(k), f[k]), k
    # overwriting an existing, longer file leaves exactly the new content
    with open(path, "ab") as fh:
        fh.write(b"x" * 1000)
    ses.save_fields(path, 0.25, x, y, z)
    ses.io_wait()
    assert os.path.getsize(path) == len(expect)
    s2.close()
    ses.close()


def test_read_fields_rejects_other_grid_and_missing_file(gpu, O, tmp_path):
    ses, g, d = session(gpu, O)
    n = (ses.cfg.nx, ses.cfg.ny, ses.cfg.nz)
    x, y, z = (d[a] * np.arange(n[a]) for a in range(3))
    path = str(tmp_path / "f.bin")
    ses.save_fields(path, 1.0, x, y, z)
    ses.io_wait()
    other = gpu.Session(gpu.make_config(n[0] + 1, n[1], n[2], *d))
    before = other.download("ux")
    with pytest.raises(gpu.O3DError) as e:      # src/IOfunctions.f90:452-459
        other.read_fields(path)
    assert e.value.code == gpu._lib.ERR_INVALID and "number of cells" in str(e.value)
    assert np.array_equal(other.download("ux"), before)
    with pytest.raises(gpu.O3DError) as e:      # :389-392 "Error opening file"
        other.read_fields(str(tmp_path / "nope.bin"))
    assert e.value.code == gpu._lib.ERR_IO
    with pytest.raises(gpu.O3DError) as e:
        ses.write_binary(str(tmp_path / "no_such_dir" / "a.bin"), "ux")
    assert e.value.code == gpu._lib.ERR_IO
    other.close()
    ses.close()


@pytest.mark.parametrize("bc,nscr,iles", [((1, 1, 1), 1, 1), ((0, 0, 0), 0, 0), ((0, 1, 0), 1, 0)])
def test_write_all_data_matches_oracle_operators(gpu, O, tmp_path, bc, nscr, iles):
    n = (23, 20, 19) if bc != (1, 1, 1) else (24, 20, 18)
    ses, g, d = session(gpu, O, n=n, bc=bc, nscr=nscr, iles=iles)
    ses.step()
    f = {k: ses.download(k) for k in ("ux", "uy", "uz", "pp", "phi", "nu_t")}
    out = str(tmp_path / "outputs")
    ses.write_all_data(out, 7)
    ses.step()                   # overlaps the drain
    ses.io_wait()
    rx, ry, rz = O.rotational(g, f["ux"], f["uy"], f["uz"])
    expect = {"ux": f["ux"], "uy": f["uy"], "uz": f["uz"], "pp": f["pp"],
              "vort": np.sqrt(rx ** 2 + ry ** 2 + rz ** 2),      # src/visualization.f90:259
              "qcrit": O.q_criterion(g, f["ux"], f["uy"], f["uz"])}
    if nscr:
        expect["phi"] = f["phi"]
    if iles:
        expect["nu_t"] = f["nu_t"]
    names = sorted(os.listdir(out))
    assert names == sorted("%s_7.bin" % k for k in expect)
    for k, a in expect.items():
        got = np.fromfile(os.path.join(out, "%s_7.bin" % k)).reshape(n, order="F")
        assert np.array_equal(got, a), (k, np.max(np.abs(got - a)))
    ses.close()


def test_many_queued_writes_back_pressure(gpu, O, tmp_path):
    """more fields queued than staging buffers: the producer blocks, nothing is lost or reordered"""
    ses, g, d = session(gpu, O)
    snaps = []
    for q in range(9):
        ses.step()
        snaps.append(ses.download("ux"))
        ses.write_binary(str(tmp_path / ("ux_%d.bin" % q)), "ux")
    ses.io_wait()
    for q, a in enumerate(snaps):
        got = np.fromfile(str(tmp_path / ("ux_%d.bin" % q))).reshape(a.shape, order="F")
        assert np.array_equal(got, a), q
    ses.close()

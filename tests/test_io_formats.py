"""CPU tests of the reference's binary formats: big-endian direct-access grid
records and the sequential `iced` restart layout, round-tripped through synthetic files; the
shipped gx3 grid is checked against the committed fixture when the reference tree is present."""
import os

import numpy as np
import pytest

from cice4_b200 import io as IO
from conftest import GX3_FIXTURE


def test_rda8_ida4_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    nx, ny = 13, 7
    recs = [rng.standard_normal((nx, ny)) for _ in range(7)]
    p = tmp_path / "grid"
    IO.write_rda8(str(p), recs)
    got = IO.read_rda8(str(p), nx, ny, 7)
    for k in range(7):
        np.testing.assert_array_equal(got[k], recs[k])
    # Fortran order on disk: i fastest, big-endian
    raw = np.fromfile(str(p), dtype=">f8")
    assert raw[1] == recs[0][1, 0] and raw[nx] == recs[0][0, 1]
    kmt = rng.integers(0, 30, size=(nx, ny)).astype(np.int32)
    pk = tmp_path / "kmt"
    kmt.T.astype(">i4").tofile(str(pk))
    np.testing.assert_array_equal(IO.read_ida4(str(pk), nx, ny), kmt)


def test_restart_dynamics_roundtrip(tmp_path):
    rng = np.random.default_rng(2)
    nx, ny, ncat = 11, 9, 5
    f = {k: np.asfortranarray(rng.standard_normal((nx, ny))) for k in
         ["uvel", "vvel", "strocnxT", "strocnyT"] + IO.STRESS_FILE_ORDER}
    for k in ("aicen", "vicen", "vsnon"):
        f[k] = np.asfortranarray(rng.random((nx, ny, ncat)))
    f["iceumask"] = np.asfortranarray((rng.random((nx, ny)) > 0.5).astype(np.int32))
    p = tmp_path / "iced"
    IO.write_restart_dynamics(str(p), dict(istep1=744, time=2678400.0, time_forc=0.0), f)
    h, g = IO.read_restart_dynamics(str(p), nx, ny)
    assert h == dict(istep1=744, time=2678400.0, time_forc=0.0)
    for k, v in f.items():
        np.testing.assert_array_equal(g[k], v), k
    # record framing and the 1,3,2,4 stress order (source/ice_restart.F90:219-235)
    recs = list(IO._records(open(str(p), "rb")))
    assert len(recs) == 1 + 4 * ncat + 25 + 2 + 5 + 2 + 12 + 1
    first_stress = 1 + 4 * ncat + 25 + 2 + 5 + 2
    np.testing.assert_array_equal(IO._field(recs[first_stress + 1], nx, ny), f["stressp_3"])
    np.testing.assert_array_equal(IO._field(recs[first_stress + 2], nx, ny), f["stressp_2"])


def test_gx3_reader_matches_fixture():
    ref = "/root/reference/input_templates/gx3"
    if not os.path.exists(os.path.join(ref, "global_gx3.grid")):
        pytest.skip("reference tree not present (GPU box)")
    g = IO.read_pop_grid(os.path.join(ref, "global_gx3.grid"), os.path.join(ref, "global_gx3.kmt"), 100, 116)
    z = np.load(GX3_FIXTURE)
    for k in ("ULAT", "ULON", "HTN", "HTE", "KMT"):
        np.testing.assert_array_equal(g[k], z[k])


def test_netcdf_grid_reader(tmp_path):
    """popgrid_nc layout (ice_grid.F90:617-839): NetCDF-3 files with (ny, nx) variables."""
    from scipy.io import netcdf_file
    rng = np.random.default_rng(4)
    nx, ny = 12, 9
    fields = {n: rng.standard_normal((nx, ny)) for n in
              ["ulat", "ulon", "htn", "hte", "angle", "tlat", "tlon", "angleT", "tarea", "uarea"]}
    fields["angle"][0, 0] = 4.0                      # clipped to pi like the reference does
    gp, kp = str(tmp_path / "grid.nc"), str(tmp_path / "kmt.nc")
    with netcdf_file(gp, "w") as f:
        f.createDimension("ny", ny)
        f.createDimension("nx", nx)
        for n, a in fields.items():
            v = f.createVariable(n, "d", ("ny", "nx"))
            v[:] = a.T
    kmt = rng.integers(0, 2, size=(nx, ny))
    with netcdf_file(kp, "w") as f:
        f.createDimension("time", 1)
        f.createDimension("ny", ny)
        f.createDimension("nx", nx)
        v = f.createVariable("kmt", "d", ("time", "ny", "nx"))
        v[0] = kmt.T
    g = IO.read_pop_grid_nc(gp, kp)
    assert set(g) == {"ULAT", "ULON", "HTN", "HTE", "ANGLE", "KMT"}
    np.testing.assert_array_equal(g["HTN"], fields["htn"])
    np.testing.assert_array_equal(g["KMT"], kmt)
    assert g["ANGLE"][0, 0] == np.pi
    ga = IO.read_pop_grid_nc(gp, kp, auscom=True)
    np.testing.assert_array_equal(ga["TAREA"], fields["tarea"])
    assert "ANGLET" in ga and "KMU" not in ga
    with pytest.raises(ValueError):
        IO.read_pop_grid_nc(kp, kp)


def test_block_layout_land_block_helpers():
    """BlockLayout.land_blocks / .without (land-block elimination of the caller's layout)"""
    from cice4_b200 import evp as E
    nx, ny = 20, 12
    lay = E.BlockLayout.cartesian(nx, ny, 8, 5)          # 3 x 3 blocks, padded east / north edge
    assert lay.nblocks == 9 and lay.shape == (10, 7, 9)
    tmask = np.ones((nx + 2, ny + 2), dtype=np.int32)
    tmask[1:9, 1:6] = 0                                  # block 0 entirely land
    tmask[17:21, 11:13] = 0                              # the small north-east corner block entirely land
    tmask[9:17, 6:11] = 0
    tmask[12, 8] = 1                                     # block 4: one ocean cell left
    land = lay.land_blocks(tmask)
    assert land == [0, 8]
    sub = lay.without(land)
    assert sub.nblocks == 7 and sub.nx_block == lay.nx_block
    assert list(sub.iglob_lo) == [int(lay.iglob_lo[b]) for b in range(9) if b not in land]
    # split / merge on the reduced layout: covered cells round-trip, uncovered cells stay zero
    a = np.asfortranarray(np.arange((nx + 2) * (ny + 2), dtype=np.float64).reshape(nx + 2, ny + 2))
    back = E.merge_blocks(E.split_blocks(a, sub, "cyclic", "open"), sub)
    assert np.array_equal(back[9:17, 1:6], a[9:17, 1:6])
    assert not back[1:9, 1:6].any()

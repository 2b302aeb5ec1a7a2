"""Readers/writers for the reference's binary grid and restart formats (host harness, numpy).

SURVEY.md 8(f) row 3: lets real grids and `iced_*` restarts drive the EVP parity tests when the
blobs are available (they are absent from this repository's copy of the reference, see
/root/reference/.MISSING_LARGE_BLOBS; only the gx3 grid + kmt and the gx1 kmt ship).

Formats (/root/reference/source/ice_read_write.F90:52-95,97-243):
  'rda8' / 'ida4'  direct access, big-endian, one record = nx_global*ny_global real*8 / integer*4,
                   i fastest (Fortran order) -- POP grid file (ULAT, ULON, HTN, HTE, HUS, HUW, ANGLE;
                   ice_grid.F90:497-607) and the kmt file;
  'ruf8'           sequential unformatted, big-endian: every record is framed by two 4-byte lengths.
Restart record order: /root/reference/source/ice_restart.F90:160-246 (header `istep1, time,
time_forc`; per category aicen, vicen, vsnon, Tsfc; eicen (ntilyr), esnon (ntslyr); uvel, vvel;
scale_factor, swvdr, swvdf, swidr, swidf; strocnxT, strocnyT; the 12 stresses in the order
1,3,2,4 per kind; iceumask as real*8).
"""
from __future__ import annotations

import struct
from typing import BinaryIO, Dict, Iterator, List, Optional, Tuple

import numpy as np

GRID_RECORDS = ["ULAT", "ULON", "HTN", "HTE", "HUS", "HUW", "ANGLE"]
# file order of the stress records, source/ice_restart.F90:219-235
STRESS_FILE_ORDER = ["stressp_1", "stressp_3", "stressp_2", "stressp_4",
                     "stressm_1", "stressm_3", "stressm_2", "stressm_4",
                     "stress12_1", "stress12_3", "stress12_2", "stress12_4"]


def read_rda8(path: str, nx: int, ny: int, nrec: Optional[int] = None) -> np.ndarray:
    """Direct-access real*8 records -> array (nrec, nx, ny), Fortran (i,j) indexing."""
    raw = np.fromfile(path, dtype=">f8")
    n = raw.size // (nx * ny)
    if raw.size != n * nx * ny or (nrec is not None and n < nrec):
        raise ValueError(f"{path}: size does not match {nx}x{ny} real*8 records")
    return np.ascontiguousarray(raw.reshape(n, ny, nx).transpose(0, 2, 1)).astype(np.float64)


def read_ida4(path: str, nx: int, ny: int, rec: int = 1) -> np.ndarray:
    raw = np.fromfile(path, dtype=">i4")
    if raw.size < rec * nx * ny:
        raise ValueError(f"{path}: size does not match {nx}x{ny} integer*4 records")
    return np.ascontiguousarray(raw[(rec - 1) * nx * ny: rec * nx * ny].reshape(ny, nx).T).astype(np.int32)


def read_pop_grid(grid_path: str, kmt_path: str, nx: int, ny: int) -> Dict[str, np.ndarray]:
    """popgrid (ice_grid.F90:497-607): the 7 grid records (radians / cm) and KMT."""
    rec = read_rda8(grid_path, nx, ny, 7)
    out = {n: rec[k] for k, n in enumerate(GRID_RECORDS)}
    out["KMT"] = read_ida4(kmt_path, nx, ny)
    return out


def read_pop_grid_nc(grid_path: str, kmt_path: str, auscom: bool = False) -> Dict[str, np.ndarray]:
    """popgrid_nc (ice_grid.F90:617-839): NetCDF-3 grid file with the variables `ulat, ulon, htn, hte,
    angle` (radians / cm, dimensions (ny, nx)) and a kmt file with `kmt`; the AusCOM build also reads
    `tlat, tlon, angleT, tarea, uarea` from the grid file (:790-818) and, if present, `kmu` from the kmt
    file (:713-731).  Returned arrays use Fortran (i, j) indexing like read_pop_grid."""
    from scipy.io import netcdf_file

    def var(nc, name):
        if name not in nc.variables:
            raise ValueError(f"variable {name!r} missing")
        a = np.array(nc.variables[name][:], dtype=np.float64)
        a = a.reshape(a.shape[-2:])          # a leading time / record dimension of length 1 is dropped
        return np.ascontiguousarray(a.T)

    out: Dict[str, np.ndarray] = {}
    with netcdf_file(grid_path, "r", mmap=False) as g:
        for name in ["ulat", "ulon", "htn", "hte", "angle"] + (["tlat", "tlon", "angleT", "tarea", "uarea"] if auscom else []):
            out[name.upper() if name != "angleT" else "ANGLET"] = var(g, name)
    with netcdf_file(kmt_path, "r", mmap=False) as k:
        out["KMT"] = var(k, "kmt").astype(np.int32)
        if auscom and "kmu" in k.variables:
            out["KMU"] = var(k, "kmu").astype(np.int32)
    for a in ("ANGLE", "ANGLET"):                # :766-767, :806-807
        if a in out:
            out[a] = np.clip(out[a], -np.pi, np.pi)
    return out


def write_rda8(path: str, records: List[np.ndarray]) -> None:
    with open(path, "wb") as f:
        for a in records:
            f.write(np.asarray(a, dtype=np.float64).T.astype(">f8").tobytes())


# ---------------------------------------------------------------------------------------------
# sequential unformatted ('ruf8')
# ---------------------------------------------------------------------------------------------
def _records(f: BinaryIO) -> Iterator[bytes]:
    while True:
        head = f.read(4)
        if not head:
            return
        (n,) = struct.unpack(">i", head)
        payload = f.read(n)
        (m,) = struct.unpack(">i", f.read(4))
        if m != n or len(payload) != n:
            raise ValueError("corrupt Fortran sequential record")
        yield payload


def _write_record(f: BinaryIO, payload: bytes) -> None:
    f.write(struct.pack(">i", len(payload)))
    f.write(payload)
    f.write(struct.pack(">i", len(payload)))


def _field(payload: bytes, nx: int, ny: int) -> np.ndarray:
    a = np.frombuffer(payload, dtype=">f8")
    if a.size != nx * ny:
        raise ValueError("record is not an nx_global*ny_global real*8 field")
    return np.asfortranarray(a.reshape(ny, nx).T.astype(np.float64))


def read_restart_dynamics(path: str, nx: int, ny: int, ncat: int = 5, ntilyr: int = 20,
                          ntslyr: int = 5) -> Tuple[Dict[str, float], Dict[str, np.ndarray]]:
    """The dynamics part of an `iced` restart: uvel, vvel, strocnxT/yT, the 12 stresses (mapped back
    from the 1,3,2,4 file order) and iceumask (int32), as global (nx, ny) arrays; plus the category
    state (aicen, vicen, vsnon) that `ice_strength` needs.  Ghost cells are the caller's business
    (scatter_global / scatter_global_stress, serial/ice_gather_scatter.F90:341-585,1140-1228)."""
    with open(path, "rb") as f:
        it = _records(f)
        h = next(it)
        if len(h) == 20:      # integer*4, real*8, real*8
            istep1, time, time_forc = struct.unpack(">idd", h)
        elif len(h) == 24:    # integer*8 build
            istep1, time, time_forc = struct.unpack(">qdd", h)
        else:
            raise ValueError("unexpected restart header")
        header = {"istep1": istep1, "time": time, "time_forc": time_forc}
        out: Dict[str, np.ndarray] = {}
        aicen = np.zeros((nx, ny, ncat), order="F")
        vicen = np.zeros((nx, ny, ncat), order="F")
        vsnon = np.zeros((nx, ny, ncat), order="F")
        for n in range(ncat):
            aicen[:, :, n] = _field(next(it), nx, ny)
            vicen[:, :, n] = _field(next(it), nx, ny)
            vsnon[:, :, n] = _field(next(it), nx, ny)
            next(it)  # Tsfc
        out.update(aicen=aicen, vicen=vicen, vsnon=vsnon)
        for _ in range(ntilyr + ntslyr):
            next(it)
        out["uvel"] = _field(next(it), nx, ny)
        out["vvel"] = _field(next(it), nx, ny)
        for _ in range(5):   # scale_factor, swvdr, swvdf, swidr, swidf
            next(it)
        out["strocnxT"] = _field(next(it), nx, ny)
        out["strocnyT"] = _field(next(it), nx, ny)
        for name in STRESS_FILE_ORDER:
            out[name] = _field(next(it), nx, ny)
        out["iceumask"] = np.asfortranarray((_field(next(it), nx, ny) > 0.5).astype(np.int32))
    return header, out


def write_restart_dynamics(path: str, header: Dict[str, float], fields: Dict[str, np.ndarray],
                           ncat: int = 5, ntilyr: int = 20, ntslyr: int = 5) -> None:
    """Inverse of read_restart_dynamics (records this package does not model are written as zeros);
    used to round-trip the format in the tests."""
    nx, ny = fields["uvel"].shape
    zero = np.zeros((nx, ny))
    enc = lambda a: np.asarray(a, dtype=np.float64).T.astype(">f8").tobytes()
    with open(path, "wb") as f:
        _write_record(f, struct.pack(">idd", int(header["istep1"]), float(header["time"]), float(header["time_forc"])))
        for n in range(ncat):
            for key in ("aicen", "vicen", "vsnon"):
                _write_record(f, enc(fields[key][:, :, n]))
            _write_record(f, enc(zero))
        for _ in range(ntilyr + ntslyr):
            _write_record(f, enc(zero))
        _write_record(f, enc(fields["uvel"]))
        _write_record(f, enc(fields["vvel"]))
        for _ in range(5):
            _write_record(f, enc(zero))
        _write_record(f, enc(fields["strocnxT"]))
        _write_record(f, enc(fields["strocnyT"]))
        for name in STRESS_FILE_ORDER:
            _write_record(f, enc(fields[name]))
        _write_record(f, enc(fields["iceumask"].astype(np.float64)))

"""Host-side mirror of `module ice_dyn_evp` over the C ABI of libevp_b200.so.

The reference is Fortran and no Fortran compiler exists in this image, so the host side above
the C ABI is written here in Python with the reference's names and call order
(/root/reference/source/ice_dyn_evp.F90): module variables `kdyn, ndte, evp_damping,
yield_curve, dragio, cosw, sinw` (:64-97), `init_evp(dt)` (:441-526), `evp(dt)` (:119-432),
`principal_stress` (:1558-1609) and `set_evp_parameters` (:535-577).  The module arrays of
ice_state / ice_flux that `evp` reads and writes are numpy arrays in Fortran order
`(nx_block, ny_block, max_blocks)` held in `self.state`, `self.flux`.

The Fortran shim a CICE maintainer would compile instead is cice4_b200/fortran/ice_dyn_evp_b200.F90
(see INTEGRATION.md); both bind exactly the entry points declared in include/evp_b200.h.

There is NO CPU fallback: importing works without a GPU (so that symbols can be checked), but
every compute call fails with EvpB200Error when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libevp_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

BND = {"open": 0, "closed": 1, "cyclic": 2, "tripole": 3, "tripoleT": 4}

EXPORTS = [
    "evp_b200_abi_version", "evp_b200_last_error", "evp_b200_default_params", "evp_b200_init",
    "evp_b200_prep", "evp_b200_run", "evp_b200_step", "evp_b200_subcycle_resident",
    "evp_b200_principal_stress", "evp_b200_get_timings", "evp_b200_diagnostics", "evp_b200_download_state",
    "evp_b200_invalidate_device_state", "evp_b200_comm_unique_id", "evp_b200_comm_init", "evp_b200_finalize",
    "evp_b200_unpin", "evp_b200_selftest_ieee", "evp_b200_get_info", "evp_b200_download_velocity",
    "evp_b200_device_velocity", "evp_b200_step_device", "evp_b200_principal_stress_n", "evp_b200_diagnostics_energy",
]


class EvpB200Error(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("nx_block", "ny_block", "max_blocks", "nblocks", "nx_global",
                                          "ny_global", "ew_boundary", "ns_boundary")] +
                [(n, c_ip) for n in ("ilo", "ihi", "jlo", "jhi", "iglob_lo", "jglob_lo")] +
                [(n, C.c_int32) for n in ("slab_jlo", "slab_jhi", "rank", "nranks", "device")])


class Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("ndte", C.c_int32), ("evp_damping", C.c_int32),
                ("dragio", C.c_double), ("cosw", C.c_double), ("sinw", C.c_double),
                ("rhoi", C.c_double), ("rhos", C.c_double), ("rhow", C.c_double),
                ("gravit", C.c_double), ("puny", C.c_double),
                ("coupled_tilt", C.c_int32), ("use_ocnslope", C.c_int32),
                ("hemisphere_turning", C.c_int32), ("wind_from_strax", C.c_int32),
                ("kstrength", C.c_int32), ("krdg_partic", C.c_int32), ("krdg_redist", C.c_int32),
                ("ncat", C.c_int32), ("mu_rdg", C.c_double),
                ("math_mode", C.c_int32), ("pin_host", C.c_int32), ("use_graph", C.c_int32),
                ("tile_threads", C.c_int32), ("tile_rows", C.c_int32), ("kernel_variant", C.c_int32),
                ("state_residency", C.c_int32), ("exchange_mode", C.c_int32)]


STATIC_D = ["dxt", "dyt", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym",
            "tarea", "tarear", "tinyarea", "uarea", "uarear", "fcor"]
STATIC_I = ["tmask", "umask"]
STATIC_OPT = ["HTE", "HTN"]   # optional: enable the 2-plane metric path of the subcycle kernel
INPUT_D = ["aice", "vice", "vsno", "strairxT", "strairyT", "uocn", "vocn", "ss_tltx", "ss_tlty",
           "aice0", "aicen", "vicen"]
STATE_D = ["uvel", "vvel",
           "stressp_1", "stressp_2", "stressp_3", "stressp_4",
           "stressm_1", "stressm_2", "stressm_3", "stressm_4",
           "stress12_1", "stress12_2", "stress12_3", "stress12_4"]
OUTPUT_D = ["strairx", "strairy", "strtltx", "strtlty", "strintx", "strinty",
            "strocnx", "strocny", "strocnxT", "strocnyT", "fm", "prs_sig",
            "divu", "shear", "rdg_conv", "rdg_shear", "strength", "sicemass", "sig1", "sig2"]


class StaticFields(C.Structure):
    _fields_ = [(n, c_dp) for n in STATIC_D] + [(n, c_ip) for n in STATIC_I] + [(n, c_dp) for n in STATIC_OPT]


class Inputs(C.Structure):
    _fields_ = [(n, c_dp) for n in INPUT_D]


class State(C.Structure):
    _fields_ = [(n, c_dp) for n in STATE_D] + [("iceumask", c_ip)]


class Outputs(C.Structure):
    _fields_ = [(n, c_dp) for n in OUTPUT_D]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("upload_ms", "prep_ms", "subcycle_ms", "finish_ms",
                                         "download_ms", "total_ms")] + \
               [("kernel_launches", C.c_int32), ("subcycle_launches", C.c_int32),
                ("exchange_mode_used", C.c_int32), ("reserved", C.c_int32)]


_lib: Optional[C.CDLL] = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libevp_b200.so and declare the prototypes of include/evp_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise EvpB200Error(f"{path} is missing: build it with `python -m cice4_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(path)
    H = C.c_void_p
    L.evp_b200_abi_version.restype = C.c_int
    L.evp_b200_last_error.restype = C.c_char_p
    L.evp_b200_default_params.argtypes = [C.POINTER(Params)]
    L.evp_b200_default_params.restype = None
    L.evp_b200_init.argtypes = [C.POINTER(Dims), C.POINTER(Params), C.POINTER(StaticFields), C.POINTER(H)]
    L.evp_b200_prep.argtypes = [H, C.POINTER(Inputs), C.POINTER(State), c_ip]
    L.evp_b200_run.argtypes = [H, c_dp, C.POINTER(State), C.POINTER(Outputs)]
    L.evp_b200_step.argtypes = [H, C.POINTER(Inputs), c_dp, C.POINTER(State), C.POINTER(Outputs)]
    L.evp_b200_subcycle_resident.argtypes = [H, C.c_int32, C.POINTER(C.c_float)]
    L.evp_b200_principal_stress.argtypes = [H, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.evp_b200_get_timings.argtypes = [H, C.POINTER(Timings)]
    L.evp_b200_diagnostics.argtypes = [H, c_dp]
    L.evp_b200_download_state.argtypes = [H, C.POINTER(State)]
    L.evp_b200_invalidate_device_state.argtypes = [H]
    L.evp_b200_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.evp_b200_comm_init.argtypes = [H, C.POINTER(C.c_uint8)]
    L.evp_b200_finalize.argtypes = [H]
    L.evp_b200_unpin.argtypes = [H, C.c_void_p]
    L.evp_b200_get_info.argtypes = [H, c_ip]
    L.evp_b200_download_velocity.argtypes = [H, c_dp, c_dp]
    L.evp_b200_device_velocity.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ip, c_ip]
    L.evp_b200_step_device.argtypes = [H, C.POINTER(Inputs), C.c_void_p, C.POINTER(State), C.POINTER(Outputs)]
    L.evp_b200_principal_stress_n.argtypes = [H, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.evp_b200_diagnostics_energy.argtypes = [H, c_dp]
    L.evp_b200_selftest_ieee.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_uint64)]
    for n in EXPORTS:
        if n not in ("evp_b200_last_error", "evp_b200_default_params"):
            getattr(L, n).restype = C.c_int
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        msg = load_library().evp_b200_last_error().decode(errors="replace")
        raise EvpB200Error(f"libevp_b200 error {rc}: {msg}")


def _dptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    if a.dtype != np.float64 or not a.flags.f_contiguous:
        raise EvpB200Error("arrays must be float64 in Fortran order (nx_block, ny_block, max_blocks)")
    return a.ctypes.data_as(c_dp)


def _iptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    if a.dtype != np.int32 or not a.flags.f_contiguous:
        raise EvpB200Error("masks must be int32 in Fortran order")
    return a.ctypes.data_as(c_ip)


def default_params(**over) -> Params:
    p = Params()
    load_library().evp_b200_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise EvpB200Error(f"unknown parameter {k}")
        setattr(p, k, v)
    return p


class BlockLayout:
    """`ice_blocks` for the caller's arrays: nx_block, ny_block, max_blocks and, per local block,
    ilo/ihi/jlo/jhi and the global index of the first physical cell
    (/root/reference/source/ice_blocks.F90:133-350)."""

    def __init__(self, nx_global: int, ny_global: int, nx_block: int, ny_block: int,
                 ilo: Sequence[int], ihi: Sequence[int], jlo: Sequence[int], jhi: Sequence[int],
                 iglob_lo: Sequence[int], jglob_lo: Sequence[int], max_blocks: Optional[int] = None):
        self.nx_global, self.ny_global = nx_global, ny_global
        self.nx_block, self.ny_block = nx_block, ny_block
        self.nblocks = len(ilo)
        self.max_blocks = max_blocks or self.nblocks
        mk = lambda v: np.ascontiguousarray(np.asarray(v, dtype=np.int32))
        self.ilo, self.ihi, self.jlo, self.jhi = mk(ilo), mk(ihi), mk(jlo), mk(jhi)
        self.iglob_lo, self.jglob_lo = mk(iglob_lo), mk(jglob_lo)

    @classmethod
    def single_block(cls, nx: int, ny: int) -> "BlockLayout":
        """One block spanning the domain (BLCKX=NXGLOB, BLCKY=NYGLOB, as bld/config.ubuntu)."""
        return cls(nx, ny, nx + 2, ny + 2, [2], [nx + 1], [2], [ny + 1], [1], [1])

    @classmethod
    def cartesian(cls, nx: int, ny: int, bx: int, by: int) -> "BlockLayout":
        """create_blocks (/root/reference/source/ice_blocks.F90:196-222): blocks of bx x by,
        j-outer / i-inner numbering, padded at the east/north edge when sizes do not divide."""
        nbx = (nx - 1) // bx + 1
        nby = (ny - 1) // by + 1
        ilo, ihi, jlo, jhi, ig, jg = [], [], [], [], [], []
        for jb in range(nby):
            js = jb * by + 1
            je = min(js + by - 1, ny)
            for ib in range(nbx):
                is_ = ib * bx + 1
                ie = min(is_ + bx - 1, nx)
                ilo.append(2)
                ihi.append(2 + (ie - is_))
                jlo.append(2)
                jhi.append(2 + (je - js))
                ig.append(is_)
                jg.append(js)
        return cls(nx, ny, bx + 2, by + 2, ilo, ihi, jlo, jhi, ig, jg)

    def without(self, blocks: Sequence[int]) -> "BlockLayout":
        """The layout with the listed blocks removed: land-block elimination of the reference's
        distribution (/root/reference/source/ice_distribution.F90: blocks without ocean points are
        assigned to no task).  The library treats cells that no block covers as land."""
        keep = [b for b in range(self.nblocks) if b not in set(blocks)]
        pick = lambda v: [int(v[b]) for b in keep]
        return BlockLayout(self.nx_global, self.ny_global, self.nx_block, self.ny_block, pick(self.ilo),
                           pick(self.ihi), pick(self.jlo), pick(self.jhi), pick(self.iglob_lo), pick(self.jglob_lo))

    def land_blocks(self, tmask_padded: np.ndarray) -> list:
        """Indices of the blocks without a single ocean T cell (tmask: padded single-block array)."""
        out = []
        for b in range(self.nblocks):
            i0, j0 = self.iglob_lo[b], self.jglob_lo[b]
            ni, nj = self.ihi[b] - self.ilo[b] + 1, self.jhi[b] - self.jlo[b] + 1
            if not tmask_padded[i0:i0 + ni, j0:j0 + nj].any():
                out.append(b)
        return out

    @property
    def shape(self):
        return (self.nx_block, self.ny_block, self.max_blocks)


class IceDynEvp:
    """`module ice_dyn_evp` on one B200 (one y-slab)."""

    def __init__(self, layout: BlockLayout, ew_boundary: str = "cyclic", ns_boundary: str = "open",
                 device: int = -1, rank: int = 0, nranks: int = 1, slab: Optional[Sequence[int]] = None,
                 **params):
        # namelist / module variables, source/ice_dyn_evp.F90:64-97
        self.kdyn = 1
        self.yield_curve = "ellipse"
        self.layout = layout
        self.ew_boundary, self.ns_boundary = ew_boundary, ns_boundary
        self.device = device
        # y-slab of the global domain owned by this rank (1-based inclusive rows), SURVEY 8(e)
        self.rank, self.nranks = rank, nranks
        self.slab = tuple(slab) if slab is not None else (1, layout.ny_global)
        self._param_over = dict(params)
        self.params: Optional[Params] = None
        self._h = C.c_void_p(None)
        self._keep = []
        # module state of ice_state / ice_flux that evp owns across steps
        sh = layout.shape
        self.state: Dict[str, np.ndarray] = {n: np.zeros(sh, order="F") for n in STATE_D}
        self.state["iceumask"] = np.zeros(sh, dtype=np.int32, order="F")
        self.flux: Dict[str, np.ndarray] = {}

    # --- module variables -----------------------------------------------------------------
    @property
    def ndte(self) -> int:
        return self.params.ndte if self.params is not None else self._param_over.get("ndte", 120)

    @property
    def evp_damping(self) -> bool:
        return bool(self.params.evp_damping) if self.params is not None else bool(self._param_over.get("evp_damping", 0))

    def set_evp_parameters(self, dt: float) -> Dict[str, float]:
        """source/ice_dyn_evp.F90:535-577 (host copy; the library computes the same scalars)."""
        ndte = self.ndte
        eyc = 0.36
        dte = dt / float(ndte)
        dtei = 1.0 / dte
        ecc = 4.0
        tdamp2 = 2.0 * eyc * dt
        dte2T = dte / tdamp2
        return dict(dte=dte, dtei=dtei, ecci=0.25, tdamp=eyc * dt, dte2T=dte2T,
                    denom1=1.0 / (1.0 + dte2T), denom2=1.0 / (1.0 + dte2T * ecc),
                    rcon=1230.0 * eyc * dt * dtei ** 2)

    # --- init_evp -------------------------------------------------------------------------
    def init_evp(self, dt: float, grid_fields: Dict[str, np.ndarray]) -> None:
        """source/ice_dyn_evp.F90:441-526.  `grid_fields` are the module ice_grid arrays
        (block layout) plus fcor; velocities, stresses and iceumask are zeroed as :487-524."""
        L = load_library()
        self.finalize()
        lay = self.layout
        p = default_params(**self._param_over)
        p.dt = dt
        self.params = p
        d = Dims()
        d.nx_block, d.ny_block, d.max_blocks, d.nblocks = lay.nx_block, lay.ny_block, lay.max_blocks, lay.nblocks
        d.nx_global, d.ny_global = lay.nx_global, lay.ny_global
        d.ew_boundary, d.ns_boundary = BND[self.ew_boundary], BND[self.ns_boundary]
        d.ilo, d.ihi, d.jlo, d.jhi = _iptr(lay.ilo), _iptr(lay.ihi), _iptr(lay.jlo), _iptr(lay.jhi)
        d.iglob_lo, d.jglob_lo = _iptr(lay.iglob_lo), _iptr(lay.jglob_lo)
        d.slab_jlo, d.slab_jhi = self.slab
        d.rank, d.nranks, d.device = self.rank, self.nranks, self.device
        sf = StaticFields()
        keep = []
        for n in STATIC_D:
            a = self._as_block(grid_fields[n], np.float64)
            keep.append(a)
            setattr(sf, n, _dptr(a))
        for n in STATIC_I:
            a = self._as_block(grid_fields[n], np.int32)
            keep.append(a)
            setattr(sf, n, _iptr(a))
        for n in STATIC_OPT:
            if grid_fields.get(n) is not None:
                a = self._as_block(grid_fields[n], np.float64)
                keep.append(a)
                setattr(sf, n, _dptr(a))
        h = C.c_void_p(None)
        _check(L.evp_b200_init(C.byref(d), C.byref(p), C.byref(sf), C.byref(h)))
        self._h = h
        for n in STATE_D:
            self.state[n][...] = 0.0
        self.state["iceumask"][...] = 0

    # --- multi-GPU: NCCL communicator over the chain of y-slabs ---------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclUniqueId (128 bytes) made on rank 0; broadcast it with the host's own transport
        (MPI_Bcast in the Fortran world, torch.distributed in the Python harness)."""
        buf = (C.c_uint8 * 128)()
        _check(load_library().evp_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid: bytes) -> None:
        if not self._h:
            raise EvpB200Error("init_evp has not been called")
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        _check(load_library().evp_b200_comm_init(self._h, buf))

    def _as_block(self, a: np.ndarray, dtype, temps: Optional[list] = None) -> np.ndarray:
        """`a` in the block layout; a copy is made only when `a` does not conform, and such per-call
        temporaries are collected in `temps` so that the library forgets them before Python frees them."""
        sh = self.layout.shape
        orig = a
        if a.ndim == 2:
            a = a.reshape(a.shape[0], a.shape[1], 1, order="F")
        if a.shape != sh:
            raise EvpB200Error(f"array shape {a.shape} does not match the block layout {sh}")
        a = np.asfortranarray(a, dtype=dtype)
        if temps is not None and not np.shares_memory(a, orig):
            temps.append(a)
        return a

    def _unpin(self, temps) -> None:
        if self.params is not None and self.params.pin_host and self._h:
            L = load_library()
            for a in temps:
                L.evp_b200_unpin(self._h, C.c_void_p(a.ctypes.data))

    # --- evp ------------------------------------------------------------------------------
    def evp(self, dt: float, inputs: Dict[str, np.ndarray], strength: Optional[np.ndarray] = None,
            want: Optional[Sequence[str]] = None, two_phase: bool = False) -> Dict[str, np.ndarray]:
        """One dynamics step, source/ice_dyn_evp.F90:119-432.  `dt` is unused at run time exactly as
        in the reference (the parameters are frozen by init_evp, :195-197).  Updates `self.state` in
        place and returns the output fields named in `want` (default: all of include/evp_b200.h's
        evp_b200_outputs except sig1/sig2).  `strength=None` runs ice_strength on the device.
        `two_phase` uses evp_b200_prep + evp_b200_run (what the Fortran shim does when ice_strength
        stays on the host) and returns the halo-updated icetmask as outputs['icetmask']."""
        if not self._h:
            raise EvpB200Error("init_evp has not been called")
        L = load_library()
        sh = self.layout.shape
        inp = Inputs()
        keep = []
        temps = []   # arrays made for this call only (non-conforming inputs): unpinned before they are freed
        for n in INPUT_D:
            a = inputs.get(n)
            if a is None or (strength is not None and n in ("aice0", "aicen", "vicen")):
                continue   # the category arrays feed only the device ice_strength
            if n in ("aicen", "vicen"):
                orig = a
                if a.ndim == 3:
                    a = a.reshape(a.shape[0], a.shape[1], a.shape[2], 1, order="F")
                a = np.asfortranarray(a, dtype=np.float64)
                if not np.shares_memory(a, orig):
                    temps.append(a)
            else:
                a = self._as_block(a, np.float64, temps)
            keep.append(a)
            setattr(inp, n, _dptr(a))
        st = State()
        for n in STATE_D:
            setattr(st, n, _dptr(self.state[n]))
        st.iceumask = _iptr(self.state["iceumask"])
        names = list(want) if want is not None else [n for n in OUTPUT_D if n not in ("sig1", "sig2")]
        out = Outputs()
        res: Dict[str, np.ndarray] = {}
        for n in names:
            # module arrays of ice_flux: allocated once and reused, like the Fortran module variables
            # (their addresses stay stable, so the library pins them once)
            a = self.flux.get(n)
            if a is None or a.shape != sh:
                a = np.zeros(sh, order="F")
                self.flux[n] = a
            res[n] = a
            setattr(out, n, _dptr(a))
        sp = None
        if strength is not None:
            sarr = self._as_block(strength, np.float64, temps)
            keep.append(sarr)
            sp = _dptr(sarr)
        try:
            if two_phase:
                if sp is None:
                    raise EvpB200Error("two_phase needs a host strength array")
                icet = self.flux.get("icetmask")
                if icet is None or icet.shape != sh:
                    icet = np.zeros(sh, dtype=np.int32, order="F")
                _check(L.evp_b200_prep(self._h, C.byref(inp), C.byref(st), _iptr(icet)))
                _check(L.evp_b200_run(self._h, sp, C.byref(st), C.byref(out)))
                res["icetmask"] = icet
            else:
                _check(L.evp_b200_step(self._h, C.byref(inp), sp, C.byref(st), C.byref(out)))
        finally:
            self._unpin(temps)
        self.flux.update(res)
        return res

    def _state_struct(self) -> State:
        st = State()
        for n in STATE_D:
            setattr(st, n, _dptr(self.state[n]))
        st.iceumask = _iptr(self.state["iceumask"])
        return st

    def download_state(self) -> None:
        """state_residency = 1: bring the device-resident stresses back into `self.state` (what the
        shim calls before dumpfile / ice_write_hist)."""
        st = self._state_struct()
        _check(load_library().evp_b200_download_state(self._h, C.byref(st)))

    def download_velocity(self) -> None:
        """state_residency = 2: bring uvel / vvel (with their ghost ring) into `self.state` -- what the
        transport scheme reads right after evp (/root/reference/source/ice_step_mod.F90:575-585)."""
        _check(load_library().evp_b200_download_velocity(self._h, _dptr(self.state["uvel"]), _dptr(self.state["vvel"])))

    def device_velocity(self):
        """(uvel_ptr, vvel_ptr, pitch, nrows): the device planes of this slab (transport hand-off on the GPU)."""
        pu, pv = C.c_void_p(None), C.c_void_p(None)
        pitch, nrows = C.c_int32(0), C.c_int32(0)
        _check(load_library().evp_b200_device_velocity(self._h, C.byref(pu), C.byref(pv), C.byref(pitch), C.byref(nrows)))
        return pu.value, pv.value, pitch.value, nrows.value

    def evp_device(self, inputs: Dict[str, int], state: Dict[str, int], outputs: Dict[str, int],
                   strength: Optional[int] = None) -> None:
        """evp_b200_step_device: every argument is a DEVICE address (int) of an array in the block layout
        (nx_block, ny_block[, ncat], max_blocks): inputs by INPUT_D name, state by STATE_D name + 'iceumask'
        (int32), outputs by OUTPUT_D name.  `strength=None` runs ice_strength on the device."""
        inp, st, out = Inputs(), State(), Outputs()
        for n, a in inputs.items():
            setattr(inp, n, C.cast(C.c_void_p(a), c_dp))
        for n in STATE_D:
            setattr(st, n, C.cast(C.c_void_p(state[n]), c_dp))
        st.iceumask = C.cast(C.c_void_p(state["iceumask"]), c_ip)
        for n, a in outputs.items():
            setattr(out, n, C.cast(C.c_void_p(a), c_dp))
        _check(load_library().evp_b200_step_device(self._h, C.byref(inp), C.c_void_p(strength), C.byref(st), C.byref(out)))

    def invalidate_device_state(self) -> None:
        """The host arrays of `self.state` were changed by the caller (restartfile): upload them again."""
        _check(load_library().evp_b200_invalidate_device_state(self._h))

    def principal_stress_block(self, stressp_1, stressm_1, stress12_1, prs_sig):
        """principal_stress on one (nx_block, ny_block) slice, as ice_history calls it block by block
        (/root/reference/source/ice_history.F90:1939-1945) -> (sig1, sig2)."""
        a = [np.asfortranarray(x, dtype=np.float64) for x in (stressp_1, stressm_1, stress12_1, prs_sig)]
        sig1 = np.zeros(a[0].shape, order="F")
        sig2 = np.zeros(a[0].shape, order="F")
        _check(load_library().evp_b200_principal_stress_n(self._h, a[0].size, *[_dptr(x) for x in a], _dptr(sig1), _dptr(sig2)))
        return sig1, sig2

    def principal_stress(self, stressp_1, stressm_1, stress12_1, prs_sig):
        """source/ice_dyn_evp.F90:1558-1609 -> (sig1, sig2)."""
        if not self._h:
            raise EvpB200Error("init_evp has not been called")
        a = [self._as_block(x, np.float64) for x in (stressp_1, stressm_1, stress12_1, prs_sig)]
        sig1 = np.zeros(self.layout.shape, order="F")
        sig2 = np.zeros(self.layout.shape, order="F")
        _check(load_library().evp_b200_principal_stress(self._h, *[_dptr(x) for x in a], _dptr(sig1), _dptr(sig2)))
        return sig1, sig2

    # --- measurement hooks (the reference's timer_dynamics, mpi/ice_timers.F90) -------------
    def subcycle_resident(self, repeats: int = 1) -> float:
        """Mean device milliseconds of one ndte subcycle loop on the resident state."""
        ms = C.c_float(0.0)
        _check(load_library().evp_b200_subcycle_resident(self._h, repeats, C.byref(ms)))
        return ms.value

    def timings(self) -> Dict[str, float]:
        t = Timings()
        _check(load_library().evp_b200_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    def info(self) -> Dict[str, int]:
        """How the ndte loop of this handle runs (evp_b200_get_info)."""
        out = (C.c_int32 * 8)()
        _check(load_library().evp_b200_get_info(self._h, out))
        return dict(tiled=int(out[0] == 1), fused=int(out[0] == 2), grid_x=out[1], grid_y=out[2], threads=out[3],
                    strip_w=out[4], stages=out[5], p2p=out[6], persistent=out[7] & 1, warp_strips=(out[7] >> 1) & 1,
                    finish_fused=(out[7] >> 2) & 1)

    def diagnostics_energy(self) -> Dict[str, float]:
        """total ice-snow kinetic energy, ice / snow volume and rms ice speed per hemisphere of this slab
        (/root/reference/source/ice_diagnostics.F90:199-234), deterministic fixed-order sums."""
        out = (C.c_double * 8)()
        _check(load_library().evp_b200_diagnostics_energy(self._h, out))
        return dict(ketotn=out[0], ketots=out[1], shmaxn=out[2], shmaxs=out[3], snwmxn=out[4], snwmxs=out[5],
                    urmsn=out[6], urmss=out[7])

    def diagnostics(self) -> Dict[str, float]:
        """max ice speed / max strength per hemisphere of this slab, as runtime_diags prints them
        (/root/reference/source/ice_diagnostics.F90:294-346)."""
        out = (C.c_double * 4)()
        _check(load_library().evp_b200_diagnostics(self._h, out))
        return dict(umaxn=out[0], umaxs=out[1], pmaxn=out[2], pmaxs=out[3])

    @staticmethod
    def selftest_ieee(n: int = 1 << 27, seed: int = 1) -> Dict[str, int]:
        """evp_b200_selftest_ieee: the straight-line IEEE sqrt / division of the subcycle kernel against
        sqrt() and operator/ on `n` generated operands (on the current CUDA device)."""
        out = (C.c_uint64 * 6)()
        _check(load_library().evp_b200_selftest_ieee(n, seed, out))
        return dict(sqrt_mismatch=out[0], sqrt_fast=out[1], div_mismatch=out[2], div_fast=out[3],
                    div_shared_rcp_mismatch=out[4], n=out[5])

    def finalize(self) -> None:
        if self._h:
            load_library().evp_b200_finalize(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass


def grid_fields_in_blocks(grid, layout: "BlockLayout", ew: str, ns: str, with_ht: bool = True) -> Dict[str, np.ndarray]:
    """Module ice_grid arrays of a cice4_b200.grid.Grid in the caller's block layout (what
    init_evp takes); HTE/HTN are included when present so that the 2-plane metric path can be used."""
    names = STATIC_D + STATIC_I + ([n for n in STATIC_OPT if n in grid.f] if with_ht else [])
    return {n: split_blocks(grid.f[n], layout, ew, ns) for n in names}


def split_blocks(a: np.ndarray, layout: BlockLayout, ew: str, ns: str) -> np.ndarray:
    """Padded single-block array (nx+2, ny+2[, k]) -> block layout, ghost cells taken from the
    neighbouring cells of the padded array (what scatter_global + ice_HaloUpdate leave in a
    multi-block run).  Host helper for tests and the harness."""
    extra = a.shape[2:]
    out = np.zeros((layout.nx_block, layout.ny_block) + extra + (layout.max_blocks,), dtype=a.dtype, order="F")
    for b in range(layout.nblocks):
        ni = layout.ihi[b] - layout.ilo[b] + 1
        nj = layout.jhi[b] - layout.jlo[b] + 1
        i0 = layout.iglob_lo[b] - 1   # padded index of the west ghost column of this block
        j0 = layout.jglob_lo[b] - 1
        out[0:ni + 2, 0:nj + 2, ..., b] = a[i0:i0 + ni + 2, j0:j0 + nj + 2, ...]
    return out


def merge_blocks(blk: np.ndarray, layout: BlockLayout, base: Optional[np.ndarray] = None) -> np.ndarray:
    """Block layout -> padded single-block array: physical cells from every block, the outer
    ghost ring from the edge blocks.  Inverse of split_blocks for consistent arrays."""
    nx, ny = layout.nx_global, layout.ny_global
    out = np.zeros((nx + 2, ny + 2), dtype=blk.dtype, order="F") if base is None else base.copy(order="F")
    for b in range(layout.nblocks):
        ni = layout.ihi[b] - layout.ilo[b] + 1
        nj = layout.jhi[b] - layout.jlo[b] + 1
        ig, jg = layout.iglob_lo[b], layout.jglob_lo[b]
        w = 0 if ig == 1 else 1
        e = ni + 2 if ig + ni - 1 == nx else ni + 1
        s = 0 if jg == 1 else 1
        n = nj + 2 if jg + nj - 1 == ny else nj + 1
        out[ig - 1 + w:ig - 1 + e, jg - 1 + s:jg - 1 + n] = blk[w:e, s:n, b]
    return out

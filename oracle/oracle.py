"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/evp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The oracle is pinned bit for bit against the machine-translated
reference (oracle/_ref, loaded by ref_lib() below; see evp_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OrcGrid(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("nx_block", "ny_block", "ilo", "ihi", "jlo", "jhi", "ew_boundary", "ns_boundary")]


class OrcParams(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ("dtei", "ecci", "dte2T", "denom1", "denom2", "rcon")] +
                [("ndte", C.c_int32), ("evp_damping", C.c_int32)] +
                [(n, C.c_double) for n in ("rhoi", "rhos", "rhow", "dragio", "gravit", "puny", "cosw", "sinw")] +
                [(n, C.c_int32) for n in ("auscom", "coupled", "use_ocnslope", "access_wind",
                                          "kstrength", "krdg_partic", "krdg_redist")] +
                [("mu_rdg", C.c_double), ("ncat", C.c_int32)])


_D_STATIC = ["dxt", "dyt", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym",
             "tarea", "tarear", "tinyarea", "uarea", "uarear", "fcor"]
_I_STATIC = ["tmask", "umask"]
_D_INPUT = ["aice", "vice", "vsno", "strairxT", "strairyT", "strax", "stray",
            "uocn", "vocn", "ss_tltx", "ss_tlty", "aice0", "aicen", "vicen", "strength_in"]
STATE_D = ["uvel", "vvel",
           "stressp_1", "stressp_2", "stressp_3", "stressp_4",
           "stressm_1", "stressm_2", "stressm_3", "stressm_4",
           "stress12_1", "stress12_2", "stress12_3", "stress12_4"]
_I_STATE = ["iceumask"]
OUT_D = ["strength", "strairx", "strairy", "strtltx", "strtlty", "strintx", "strinty",
         "strocnx", "strocny", "strocnxT", "strocnyT", "fm", "prs_sig",
         "divu", "shear", "rdg_conv", "rdg_shear", "sicemass"]
_I_OUT = ["icetmask"]
SCRATCH_D = ["tmass", "umass", "aiu", "umassdtei", "waterx", "watery", "forcex", "forcey"]


class OrcFields(C.Structure):
    _fields_ = ([(n, c_dp) for n in _D_STATIC] + [(n, c_ip) for n in _I_STATIC] +
                [(n, c_dp) for n in _D_INPUT] +
                [(n, c_dp) for n in STATE_D] + [(n, c_ip) for n in _I_STATE] +
                [(n, c_dp) for n in OUT_D] + [(n, c_ip) for n in _I_OUT] +
                [(n, c_dp) for n in SCRATCH_D])


def build(force: bool = False) -> None:
    """Compile the oracle libraries with oracle/Makefile (gcc only)."""
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)


_libs: Dict[str, C.CDLL] = {}


def lib(kind: str = "strict") -> C.CDLL:
    if kind not in _libs:
        path = os.path.join(HERE, f"liboracle_{kind}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_evp.restype = C.c_int
        L.orc_evp.argtypes = [C.POINTER(OrcGrid), C.POINTER(OrcParams), C.POINTER(OrcFields), c_dp]
        L.orc_subcycle_only.restype = C.c_int
        L.orc_subcycle_only.argtypes = [C.POINTER(OrcGrid), C.POINTER(OrcParams), C.POINTER(OrcFields),
                                        C.c_int, c_dp]
        L.orc_set_evp_parameters.argtypes = [C.POINTER(OrcParams), C.c_double, C.c_int]
        L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
        L.orc_halo_r8.argtypes = [c_dp, C.POINTER(OrcGrid), C.c_int, C.c_int, C.c_double]
        L.orc_halo_i4.argtypes = [c_ip, C.POINTER(OrcGrid), C.c_int, C.c_int, C.c_int32]
        L.orc_principal_stress.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_double, c_dp, c_dp]
        # primitives, for the slab (multi-rank) restatement in tests/test_slab_gloo.py
        G_, P_, F_ = C.POINTER(OrcGrid), C.POINTER(OrcParams), C.POINTER(OrcFields)
        L.orc_evp_prep1.argtypes = [G_, P_, F_]
        L.orc_evp_prep2.argtypes = [G_, P_, F_, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip]
        L.orc_ice_strength.argtypes = [G_, P_, F_, C.c_int32, c_ip, c_ip]
        L.orc_stress.argtypes = [G_, P_, F_, C.c_int, C.c_int32, c_ip, c_ip, c_dp]
        L.orc_stepu.argtypes = [G_, P_, F_, C.c_int32, c_ip, c_ip, c_dp]
        L.orc_evp_finish.argtypes = [G_, P_, F_, C.c_int32, c_ip, c_ip]
        L.orc_to_ugrid.argtypes = [G_, c_dp, c_dp, c_dp, c_dp]
        L.orc_to_tgrid.argtypes = [G_, c_dp, c_dp, c_dp, c_dp]
        for fn in ("orc_evp_prep1", "orc_evp_prep2", "orc_ice_strength", "orc_stress", "orc_stepu",
                   "orc_evp_finish", "orc_to_ugrid", "orc_to_tgrid", "orc_halo_r8", "orc_halo_i4"):
            getattr(L, fn).restype = None
        _libs[kind] = L
    return _libs[kind]


def make_params(dt: float = 3600.0, ndte: int = 120, kind: str = "strict", **over) -> OrcParams:
    p = OrcParams()
    L = lib(kind)
    L.orc_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    L.orc_set_evp_parameters(C.byref(p), dt, ndte)
    return p


def make_grid(nx_block: int, ny_block: int, ew: int, ns: int) -> OrcGrid:
    return OrcGrid(nx_block, ny_block, 2, nx_block - 1, 2, ny_block - 1, ew, ns)


def _ptr(a: Optional[np.ndarray], ip: bool = False):
    if a is None:
        return None
    assert a.flags.f_contiguous, "oracle arrays must be Fortran-ordered"
    assert a.dtype == (np.int32 if ip else np.float64)
    return a.ctypes.data_as(c_ip if ip else c_dp)


def halo_r8(a: np.ndarray, ew: int, ns: int, loc: int, kind: int, lib_kind: str = "strict") -> None:
    g = make_grid(a.shape[0], a.shape[1], ew, ns)
    lib(lib_kind).orc_halo_r8(_ptr(a), C.byref(g), loc, kind, 0.0)


def principal_stress(sp1, sm1, s12, prs, puny=1e-11, lib_kind="strict"):
    sig1 = np.zeros_like(sp1, order="F")
    sig2 = np.zeros_like(sp1, order="F")
    lib(lib_kind).orc_principal_stress(sp1.shape[0], sp1.shape[1], _ptr(sp1), _ptr(sm1), _ptr(s12),
                                       _ptr(prs), puny, _ptr(sig1), _ptr(sig2))
    return sig1, sig2


class Fields:
    """Owns every array of one oracle run and the ctypes view onto them."""

    def __init__(self, grid_fields: Dict[str, np.ndarray], inputs: Dict[str, np.ndarray],
                 state: Dict[str, np.ndarray], strength_in: Optional[np.ndarray] = None):
        shape = grid_fields["dxt"].shape      # (nx_block, ny_block) or (nx_block, ny_block, nblocks)
        nxb, nyb = shape[:2]
        self.shape = shape
        self.arr: Dict[str, Optional[np.ndarray]] = {}
        for n in _D_STATIC + _I_STATIC:
            self.arr[n] = grid_fields[n]
        for n in _D_INPUT:
            self.arr[n] = inputs.get(n)
        self.arr["strength_in"] = strength_in
        for n in STATE_D + _I_STATE:
            self.arr[n] = state[n]
        for n in OUT_D + SCRATCH_D:
            self.arr[n] = np.zeros(shape, order="F")
        self.arr["icetmask"] = np.zeros(shape, dtype=np.int32, order="F")
        self.c = OrcFields()
        for n, _t in OrcFields._fields_:
            a = self.arr[n]
            setattr(self.c, n, _ptr(a, ip=(a is not None and a.dtype == np.int32)))

    def __getitem__(self, k):
        return self.arr[k]


def run_evp(grid, inputs, state, params: Optional[OrcParams] = None, strength_in=None,
            lib_kind: str = "strict"):
    """One `evp(dt)` call.  `grid` is cice4_b200.grid.Grid; `state` arrays are updated
    in place.  Returns (Fields, subcycle_seconds)."""
    p = params if params is not None else make_params(kind=lib_kind)
    f = Fields(grid.f, inputs, state, strength_in)
    g = make_grid(grid.nx_block, grid.ny_block, grid.ew, grid.ns)
    sec = C.c_double(0.0)
    L = lib(lib_kind)
    rc = L.orc_evp(C.byref(g), C.byref(p), C.byref(f.c), C.byref(sec))
    if rc != 0:
        raise RuntimeError("orc_evp failed")
    f.strength_pre = _strength_pre(L.orc_last_strength_prehalo, f["strength"])
    return f, sec.value


def _strength_pre(fn, like: np.ndarray) -> np.ndarray:
    """ice_strength's result of the last call BEFORE evp's halo update of it: what a host that runs ice_strength
    itself hands to the two-phase entry (on the T-fold the halo update is not idempotent, so the post-halo array
    would not do)"""
    out = np.zeros(like.shape, order="F")
    fn.restype = C.c_size_t
    fn.argtypes = [c_dp]
    n = fn(_ptr(out))
    assert n == out.size, (n, out.size)
    return out


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own Fortran, machine-translated and compiled (oracle/build_ref.py)
# ---------------------------------------------------------------------------------------------
REF_DIR = os.path.join(HERE, "_ref")
_ref_libs: Dict[str, C.CDLL] = {}


def ref_variant(p: OrcParams) -> str:
    """CPP variant of the reference build that the run-time flags of `p` stand for."""
    if p.auscom:
        if not p.coupled:
            raise ValueError("the reference builds AusCOM only together with coupled")
        return "access" if p.access_wind else "auscom"
    if p.access_wind:
        raise ValueError("the reference builds ACCESS only together with AusCOM")
    return "coupled" if p.coupled else "cice4"


def ref_available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f"libevp_ref_{v}.so"))
               for v in ("cice4", "coupled", "auscom", "access"))


def ref_lib(variant: str) -> C.CDLL:
    if variant not in _ref_libs:
        L = C.CDLL(os.path.join(REF_DIR, f"libevp_ref_{variant}.so"))
        L.ref_evp.restype = C.c_int
        L.ref_evp.argtypes = [C.POINTER(OrcGrid), C.POINTER(OrcParams), C.POINTER(OrcFields), C.c_double]
        L.ref_set_evp_parameters.restype = None
        L.ref_set_evp_parameters.argtypes = [C.POINTER(OrcParams), C.c_double, c_dp]
        L.ref_subcycle_only.restype = C.c_int
        L.ref_subcycle_only.argtypes = [C.POINTER(OrcGrid), C.POINTER(OrcParams), C.POINTER(OrcFields),
                                        C.c_double, C.c_int, c_dp]
        L.ref_principal_stress.restype = None
        L.ref_principal_stress.argtypes = [C.POINTER(OrcParams), C.c_int32, C.c_int32] + [c_dp] * 6
        _ref_libs[variant] = L
    return _ref_libs[variant]


def ref_abort_message(params: OrcParams) -> str:
    """what the translated reference passed to abort_ice, or why the host refused the domain"""
    L = ref_lib(ref_variant(params))
    L.ref_abort_message.restype = C.c_char_p
    return L.ref_abort_message().decode()


def run_evp_ref(grid, inputs, state, params: OrcParams, dt: float):
    """One `evp(dt)` call of the translated reference (same contract as run_evp; the locals of the
    reference's `evp` -- icetmask, tmass, umass, aiu, ... -- are not exported)."""
    f = Fields(grid.f, inputs, state, None)
    g = make_grid(grid.nx_block, grid.ny_block, grid.ew, grid.ns)
    L = ref_lib(ref_variant(params))
    rc = L.ref_evp(C.byref(g), C.byref(params), C.byref(f.c), dt)
    if rc != 0:
        raise RuntimeError("ref_evp failed: " + ref_abort_message(params))
    f.strength_pre = _strength_pre(L.ref_last_strength_prehalo, f["strength"])
    return f


class RefLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nblocks", "nx_global", "ny_global")] + \
               [(n, c_ip) for n in ("ilo", "ihi", "jlo", "jhi", "iglob_lo", "jglob_lo")]


def run_evp_ref_blocks(layout, ew: int, ns: int, grid_fields_blk, inputs_blk, state_blk, params: OrcParams,
                       dt: float):
    """One `evp(dt)` call of the translated reference on a create_blocks decomposition.  `layout` is a
    cice4_b200.evp.BlockLayout; every array is in block layout (nx_block, ny_block[, ncat], nblocks) with
    the ghost cells a multi-block run would hold (cice4_b200.evp.split_blocks); state arrays are updated
    in place."""
    L = ref_lib(ref_variant(params))
    L.ref_evp_blocks.restype = C.c_int
    L.ref_evp_blocks.argtypes = [C.POINTER(OrcGrid), C.POINTER(RefLayout), C.POINTER(OrcParams),
                                 C.POINTER(OrcFields), C.c_double]
    f = Fields(grid_fields_blk, inputs_blk, state_blk, None)
    g = OrcGrid(layout.nx_block, layout.ny_block, 2, layout.nx_block - 1, 2, layout.ny_block - 1, ew, ns)
    lay = RefLayout(layout.nblocks, layout.nx_global, layout.ny_global,
                    *[a.ctypes.data_as(c_ip) for a in (layout.ilo, layout.ihi, layout.jlo, layout.jhi,
                                                      layout.iglob_lo, layout.jglob_lo)])
    rc = L.ref_evp_blocks(C.byref(g), C.byref(lay), C.byref(params), C.byref(f.c), dt)
    if rc != 0:
        raise RuntimeError("ref_evp_blocks failed: " + ref_abort_message(params))
    f.strength_pre = _strength_pre(L.ref_last_strength_prehalo, f["strength"])
    return f


def ref_halo(a: np.ndarray, ew: int, ns: int, loc: int, kind: int, layout=None, variant: str = "cice4") -> None:
    """The REFERENCE's own halo update (translated ice_HaloCreate message loop / ice_HaloMsgCreate /
    ice_HaloUpdate2DR8|2DI4, serial/ice_boundary.F90) applied in place to `a`: one block (nx_block, ny_block) or, with
    `layout` (a cice4_b200.evp.BlockLayout), a block array (nx_block, ny_block, nblocks)."""
    L = ref_lib(variant)
    is_int = a.dtype == np.int32
    fn = L.ref_halo_update_i4 if is_int else L.ref_halo_update_r8
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(OrcGrid), C.POINTER(RefLayout), c_ip if is_int else c_dp, C.c_int32, C.c_int32]
    if layout is None:
        g = make_grid(a.shape[0], a.shape[1], ew, ns)
        rc = fn(C.byref(g), None, _ptr(a, is_int), loc, kind)
    else:
        g = OrcGrid(layout.nx_block, layout.ny_block, 2, layout.nx_block - 1, 2, layout.ny_block - 1, ew, ns)
        lay = RefLayout(layout.nblocks, layout.nx_global, layout.ny_global,
                        *[x.ctypes.data_as(c_ip) for x in (layout.ilo, layout.ihi, layout.jlo, layout.jhi,
                                                          layout.iglob_lo, layout.jglob_lo)])
        rc = fn(C.byref(g), C.byref(lay), _ptr(a, is_int), loc, kind)
    if rc != 0:
        L.ref_abort_message.restype = C.c_char_p
        raise RuntimeError("reference halo update failed: " + L.ref_abort_message().decode())


def omp_set_num_threads(n: int) -> None:
    """Thread count of the OpenMP regions of the libraries loaded in this process (libgomp)."""
    C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))


def time_subcycles_ref(grid, fields: Fields, params: OrcParams, dt: float, nsub: int,
                       threads: Optional[int] = None) -> float:
    """Seconds for `nsub` subcycles of the translated reference's own stress + stepu (+ 2 halo updates)
    on fields prepared by run_evp: the -O3 build whose cell loops run on `threads` OpenMP threads
    (None = the current setting; 1 = the reference's serial build)."""
    g = make_grid(grid.nx_block, grid.ny_block, grid.ew, grid.ns)
    sec = C.c_double(0.0)
    L = ref_lib("cice4_fast")
    if threads is not None:
        omp_set_num_threads(threads)
    rc = L.ref_subcycle_only(C.byref(g), C.byref(params), C.byref(fields.c), dt, nsub, C.byref(sec))
    if rc != 0:
        raise RuntimeError("ref_subcycle_only failed")
    return sec.value


def time_subcycles(grid, fields: Fields, params: OrcParams, nsub: int, lib_kind: str = "fast") -> float:
    """Seconds for `nsub` subcycles (stress+stepu+2 halos) on already-prepared fields."""
    g = make_grid(grid.nx_block, grid.ny_block, grid.ew, grid.ns)
    sec = C.c_double(0.0)
    rc = lib(lib_kind).orc_subcycle_only(C.byref(g), C.byref(params), C.byref(fields.c), nsub, C.byref(sec))
    if rc != 0:
        raise RuntimeError("orc_subcycle_only failed")
    return sec.value

"""TEST INFRASTRUCTURE: `evp` for one y-slab of a multi-rank run, built from the oracle's
primitives, with the row exchanges of cice4_b200.slab.EXCHANGES done through a caller-supplied
`exchange(arrays)` function (torch.distributed/gloo in tests/test_slab_gloo.py).  It follows
source/ice_dyn_evp.F90:119-432 exactly like oracle/evp_oracle.c:orc_evp, but every
ice_HaloUpdate is "local halo (east-west wrap, tripole fold on the top slab) + row exchange"."""
import ctypes as C

import numpy as np

from cice4_b200 import grid as G


def slab_fields(O, case, rows, state):
    """Cut rows [jlo-1 .. jhi+1] (ghost rows included) out of the padded global arrays."""
    jlo, jhi = rows
    cut = lambda a: np.asfortranarray(a[:, jlo - 1:jhi + 2, ...])
    gf = {k: cut(v) for k, v in case.grid.f.items() if k in O._D_STATIC + O._I_STATIC}
    inp = {k: cut(v) for k, v in case.inputs.items()}
    st = {k: cut(v) for k, v in state.items()}
    return O.Fields(gf, inp, st), st


def slab_evp(O, f, nx, rows, ny, ew, ns, rank, world, p, exchange):
    """One evp call on a slab.  f: oracle Fields of the slab; exchange(list of 2-D arrays) swaps
    row nyl -> north ghost row 0 and row 1 -> south ghost row nyl+1 with the neighbours."""
    L = O.lib("strict")
    nyl = rows[1] - rows[0] + 1
    ns_local = ns if rank == world - 1 else G.BND_OPEN      # the fold lives on the last slab
    if ns_local == G.BND_CYCLIC:
        raise ValueError("north-south cyclic is single-slab only")
    g = O.make_grid(nx + 2, nyl + 2, ew, ns_local)
    gp, pp, fp = C.byref(g), C.byref(p), C.byref(f.c)
    n = (nx + 2) * (nyl + 2)
    idx = [np.zeros(n, dtype=np.int32) for _ in range(4)]
    ip = [a.ctypes.data_as(O.c_ip) for a in idx]
    icellt, icellu = C.c_int32(0), C.c_int32(0)
    dptr = lambda a: a.ctypes.data_as(O.c_dp)

    def halo(a, loc, kind):
        if a.dtype == np.int32:
            L.orc_halo_i4(a.ctypes.data_as(O.c_ip), gp, loc, kind, 0)
        else:
            L.orc_halo_r8(dptr(a), gp, loc, kind, 0.0)
        exchange([a])

    for k in ("rdg_conv", "rdg_shear", "divu", "shear", "prs_sig"):
        f[k][...] = 0.0
    L.orc_evp_prep1(gp, pp, fp)
    halo(f["icetmask"], G.LOC_CENTER, G.TYPE_SCALAR)
    L.orc_to_ugrid(gp, dptr(f["tarea"]), dptr(f["uarea"]), dptr(f["tmass"]), dptr(f["umass"]))
    L.orc_to_ugrid(gp, dptr(f["tarea"]), dptr(f["uarea"]), dptr(f["aice"]), dptr(f["aiu"]))
    for k in ("strairx", "strairy"):
        w = f[k].copy(order="F")
        halo(w, G.LOC_CENTER, G.TYPE_VECTOR)
        L.orc_to_ugrid(gp, dptr(f["tarea"]), dptr(f["uarea"]), dptr(w), dptr(f[k]))
    L.orc_evp_prep2(gp, pp, fp, C.byref(icellt), C.byref(icellu), *ip)
    L.orc_ice_strength(gp, pp, fp, icellt, ip[0], ip[1])
    halo(f["strength"], G.LOC_CENTER, G.TYPE_SCALAR)
    halo(f["uvel"], G.LOC_NECORNER, G.TYPE_VECTOR)
    halo(f["vvel"], G.LOC_NECORNER, G.TYPE_VECTOR)
    str_ = np.zeros(n * 8)
    for ksub in range(1, p.ndte + 1):
        L.orc_stress(gp, pp, fp, ksub, icellt, ip[0], ip[1], dptr(str_))
        L.orc_stepu(gp, pp, fp, icellu, ip[2], ip[3], dptr(str_))
        L.orc_halo_r8(dptr(f["uvel"]), gp, G.LOC_NECORNER, G.TYPE_VECTOR, 0.0)
        L.orc_halo_r8(dptr(f["vvel"]), gp, G.LOC_NECORNER, G.TYPE_VECTOR, 0.0)
        exchange([f["uvel"], f["vvel"]])
    L.orc_evp_finish(gp, pp, fp, icellu, ip[2], ip[3])
    for k in ("strocnxT", "strocnyT"):
        w = f[k].copy(order="F")
        halo(w, G.LOC_NECORNER, G.TYPE_VECTOR)
        L.orc_to_tgrid(gp, dptr(f["tarea"]), dptr(f["uarea"]), dptr(w), dptr(f[k]))

"""Host-side grid construction for the EVP path (numpy, no GPU, no oracle).

Builds the padded single-block arrays `evp` reads -- the same arrays module
ice_grid owns in the reference -- from global HTN/HTE/ULAT/land-mask arrays,
following `init_grid2` (/root/reference/source/ice_grid.F90:263-487),
`primary_grid_lengths_HTN/HTE` (:1139-1289), `makemask` (:1298-1399) and the
ghost-cell fill of `scatter_global` (/root/reference/serial/ice_gather_scatter.F90
:341-585) for one block that spans the whole domain (nghost = 1,
/root/reference/source/ice_blocks.F90:56-61).

Arrays are Fortran-ordered `(nx_block, ny_block)` so that their memory layout is
the reference's `(nx_block, ny_block, 1)`; index `[i-1, j-1]` is Fortran `(i,j)`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict

import numpy as np

# boundary types (same numeric values as include/evp_b200.h)
BND_OPEN, BND_CLOSED, BND_CYCLIC, BND_TRIPOLE, BND_TRIPOLET = 0, 1, 2, 3, 4
BND_NAMES = {"open": BND_OPEN, "closed": BND_CLOSED, "cyclic": BND_CYCLIC, "tripole": BND_TRIPOLE,
             "tripoleT": BND_TRIPOLET}
LOC_CENTER, LOC_NECORNER, LOC_NFACE, LOC_EFACE = 1, 2, 3, 4
TYPE_SCALAR, TYPE_VECTOR, TYPE_ANGLE = 1, 2, 3

# drivers/cice4/ice_constants.F90:65-67,166-176
OMEGA = 7.292e-5
RADIUS = 6.37e6
PUNY = 1.0e-11
CM_TO_M = 0.01


def fzeros(shape, dtype=np.float64):
    return np.zeros(shape, dtype=dtype, order="F")


def _ghost_index(n: int, bnd: int, north: bool):
    """Global index map for local 1..n+2 (ice_blocks.F90:237-343), 1-based; 0 = none,
    negative = tripole fold row."""
    idx = np.arange(0, n + 2)  # global index of local cell: 0..n+1
    out = idx.copy()
    # low side
    if bnd == BND_CYCLIC:
        out[0] = n
    elif bnd in (BND_OPEN, BND_TRIPOLE, BND_TRIPOLET):
        out[0] = 1  # nghost - j + 1
    else:
        out[0] = 0
    # high side
    if bnd == BND_CYCLIC:
        out[n + 1] = 1
    elif bnd == BND_OPEN:
        out[n + 1] = n  # 2*n - (n+1) + 1
    elif bnd in (BND_TRIPOLE, BND_TRIPOLET) and north:
        out[n + 1] = -(n + 1)
    else:
        out[n + 1] = 0
    return out


def scatter_global(a_g: np.ndarray, ew: int, ns: int, loc: int = LOC_CENTER,
                   kind: int = TYPE_SCALAR) -> np.ndarray:
    """Global (nx,ny) -> padded single block with ghost fill
    (serial/ice_gather_scatter.F90:404-534; u-fold offsets :426-443, T-fold offsets :407-425 with the
    second source row `yoffset2` for U rows on the T-fold, :520-534)."""
    nx, ny = a_g.shape
    out = fzeros((nx + 2, ny + 2), a_g.dtype)
    ig = _ghost_index(nx, ew, north=False)
    jg = _ghost_index(ny, ns, north=True)
    if ns == BND_TRIPOLET:
        xoff, yoff = {LOC_CENTER: (2, 0), LOC_NECORNER: (1, -1), LOC_EFACE: (1, 0), LOC_NFACE: (2, -1)}[loc]
    else:
        xoff, yoff = {LOC_CENTER: (1, 1), LOC_NECORNER: (0, 0), LOC_EFACE: (0, 1), LOC_NFACE: (1, 0)}[loc]
    isign = 1 if kind == TYPE_SCALAR else -1
    for j in range(ny + 2):
        if jg[j] > 0:
            cols = ig != 0
            out[cols, j] = a_g[ig[cols] - 1, jg[j] - 1]
    for j in range(ny + 2):      # fold rows last: on the T-fold they also overwrite the row below (:531)
        if jg[j] < 0:
            for yoff2 in range(0, max(yoff, 0) - yoff + 1):
                jsrc = ny + yoff + yoff2 + (jg[j] + ny)
                for i in range(nx + 2):
                    if ig[i] != 0:
                        isrc = nx + xoff - ig[i]
                        if isrc < 1:
                            isrc += nx
                        if isrc > nx:
                            isrc -= nx
                        out[i, j - yoff2] = isign * a_g[isrc - 1, jsrc - 1]
    return out


def halo_update(a: np.ndarray, ew: int, ns: int, loc: int = LOC_CENTER,
                kind: int = TYPE_SCALAR) -> None:
    """In-place ghost update of one padded block holding the whole domain
    (serial/ice_boundary.F90:591-873 with the address lists of :3494-4202): independent numpy restatement,
    used to build the grids and to cross-check oracle/evp_oracle.c's halo update."""
    nxb, nyb = a.shape
    nx, ny = nxb - 2, nyb - 2
    tfold = ns == BND_TRIPOLET
    trip = ns == BND_TRIPOLE or tfold
    trows = 3 if tfold else 2
    if trip:
        buf = a[1:nx + 1, ny - trows + 1:ny + 1].copy()  # rows jhi-trows+1 .. jhi ('north' message, :3699-3733)
        # the 'northeast' / 'northwest' messages of a tripole block (:3833-3848) come last in the list and copy
        # rows jhi-1, jhi into buffer rows 1, 2 whatever tripoleRows is: a repeat on the u-fold, an overwrite of
        # rows 1 and 2 on the T-fold -- the reference's actual behaviour (its own translated halo shows it)
        buf[:, 0:2] = a[1:nx + 1, ny - 1:ny + 1]
    if ew == BND_CYCLIC:
        a[0, 1:ny + 1] = a[nx, 1:ny + 1]
        a[nx + 1, 1:ny + 1] = a[1, 1:ny + 1]
    if ns == BND_CYCLIC:
        a[1:nx + 1, 0] = a[1:nx + 1, ny]
        a[1:nx + 1, ny + 1] = a[1:nx + 1, 1]
    if ew == BND_CYCLIC and ns == BND_CYCLIC:
        a[0, 0] = a[nx, ny]
        a[nx + 1, 0] = a[1, ny]
        a[0, ny + 1] = a[nx, 1]
        a[nx + 1, ny + 1] = a[1, 1]
    if trip:
        isign = 1 if kind == TYPE_SCALAR else -1
        integer = np.issubdtype(a.dtype, np.integer)

        def avg(x1, x2):
            v = 0.5 * (float(x1) + isign * float(x2))
            if integer:   # nint: half away from zero
                return int(np.floor(v + 0.5)) if v >= 0 else -int(np.floor(-v + 0.5))
            return v

        def symmetrise(lo, hi, mirror):
            for i in range(lo, hi + 1):
                idst = mirror - i
                xavg = avg(buf[i - 1, trows - 1], buf[idst - 1, trows - 1])
                buf[i - 1, trows - 1] = xavg
                buf[idst - 1, trows - 1] = isign * xavg

        if tfold:   # :725-773
            ioff, joff = {LOC_CENTER: (-1, 0), LOC_NECORNER: (0, 1), LOC_EFACE: (0, 0), LOC_NFACE: (-1, 1)}[loc]
            if loc == LOC_CENTER:
                symmetrise(2, nx // 2, nx + 2)
            elif loc == LOC_EFACE:
                symmetrise(1, nx // 2, nx + 1)
        else:       # :777-827
            ioff, joff = {LOC_CENTER: (0, 0), LOC_NECORNER: (1, 1), LOC_EFACE: (1, 0), LOC_NFACE: (0, 1)}[loc]
            if loc == LOC_NECORNER:
                symmetrise(1, nx // 2 - 1, nx)
            elif loc == LOC_NFACE:
                symmetrise(1, nx // 2, nx + 1)
        ig = _ghost_index(nx, ew, north=False)
        for j in (1, 2):
            jsrc = 4 - j - joff
            jdst = ny + j - 1  # local 1-based row jhi + j - 1 -> 0-based index
            if not (0 < jsrc <= trows):
                continue
            for i in range(nx + 2):
                isrc = nx - ig[i] + 1 - ioff
                if isrc == 0:
                    isrc = nx
                if isrc > nx:
                    isrc -= nx
                a[i, jdst] = isign * buf[isrc - 1, jsrc - 1]


@dataclass
class Grid:
    """Padded single-block grid fields consumed by `evp` (module ice_grid state)."""
    nx: int
    ny: int
    ew: int
    ns: int
    f: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def nx_block(self):
        return self.nx + 2

    @property
    def ny_block(self):
        return self.ny + 2

    def __getattr__(self, name):
        f = object.__getattribute__(self, "f")
        if name in f:
            return f[name]
        raise AttributeError(name)


def build_grid(htn_m: np.ndarray, hte_m: np.ndarray, ulat: np.ndarray, hm: np.ndarray,
               ew: int, ns: int) -> Grid:
    """init_grid2 for one whole-domain block.  htn_m/hte_m in metres, ulat in radians,
    hm = 1 ocean / 0 land; all (nx, ny) global arrays."""
    nx, ny = htn_m.shape
    g = Grid(nx, ny, ew, ns)
    ip1 = np.r_[1:nx, 0]
    im1 = np.r_[nx - 1, 0:nx - 1]
    # primary_grid_lengths_HTN (ice_grid.F90:1173-1205)
    dxu_g = 0.5 * (htn_m + htn_m[ip1, :])
    dxt_g = np.empty_like(htn_m)
    dxt_g[:, 1:] = 0.5 * (htn_m[:, 1:] + htn_m[:, :-1])
    dxt_g[:, 0] = 2.0 * htn_m[:, 1] - htn_m[:, 2]
    # primary_grid_lengths_HTE (:1252-1285)
    dyu_g = np.empty_like(hte_m)
    dyu_g[:, :-1] = 0.5 * (hte_m[:, :-1] + hte_m[:, 1:])
    dyu_g[:, -1] = 2.0 * hte_m[:, -2] - hte_m[:, -3]
    dyt_g = 0.5 * (hte_m + hte_m[im1, :])

    HTN = scatter_global(htn_m, ew, ns, LOC_NFACE)
    HTE = scatter_global(hte_m, ew, ns, LOC_EFACE)
    dxu = scatter_global(dxu_g, ew, ns, LOC_NECORNER)
    dyu = scatter_global(dyu_g, ew, ns, LOC_NECORNER)
    dxt = scatter_global(dxt_g, ew, ns, LOC_CENTER)
    dyt = scatter_global(dyt_g, ew, ns, LOC_CENTER)
    ULAT = scatter_global(ulat, ew, ns, LOC_NECORNER)
    hmp = scatter_global(hm.astype(np.float64), ew, ns, LOC_CENTER)

    sh = (nx + 2, ny + 2)
    tarea, uarea, tarear, uarear, tinyarea = (fzeros(sh) for _ in range(5))
    dxhy, dyhx, cyp, cxp, cym, cxm = (fzeros(sh) for _ in range(6))
    I = slice(1, nx + 1)
    J = slice(1, ny + 1)
    # :332-353
    tarea[I, J] = dxt[I, J] * dyt[I, J]
    uarea[I, J] = dxu[I, J] * dyu[I, J]
    with np.errstate(divide="ignore"):
        tarear[I, J] = np.where(tarea[I, J] > 0, 1.0 / tarea[I, J], 0.0)
        uarear[I, J] = np.where(uarea[I, J] > 0, 1.0 / uarea[I, J], 0.0)
    tinyarea[I, J] = PUNY * tarea[I, J]
    dxhy[I, J] = 0.5 * (HTE[I, J] - HTE[0:nx, J])
    dyhx[I, J] = 0.5 * (HTN[I, J] - HTN[I, 0:ny])
    # :355-363 (N and E ghost cells included)
    I1 = slice(1, nx + 2)
    J1 = slice(1, ny + 2)
    cyp[I1, J1] = 1.5 * HTE[I1, J1] - 0.5 * HTE[0:nx + 1, J1]
    cxp[I1, J1] = 1.5 * HTN[I1, J1] - 0.5 * HTN[I1, 0:ny + 1]
    cym[I1, J1] = -(1.5 * HTE[0:nx + 1, J1] - 0.5 * HTE[I1, J1])
    cxm[I1, J1] = -(1.5 * HTN[I1, 0:ny + 1] - 0.5 * HTN[I1, J1])
    # :377-399
    for a, loc, kind in ((tarea, LOC_CENTER, TYPE_SCALAR), (uarea, LOC_NECORNER, TYPE_SCALAR),
                         (tarear, LOC_CENTER, TYPE_SCALAR), (uarear, LOC_NECORNER, TYPE_SCALAR),
                         (tinyarea, LOC_CENTER, TYPE_SCALAR), (dxhy, LOC_CENTER, TYPE_VECTOR),
                         (dyhx, LOC_CENTER, TYPE_VECTOR)):
        halo_update(a, ew, ns, loc, kind)
    # makemask :1323-1367
    halo_update(hmp, ew, ns, LOC_CENTER, TYPE_SCALAR)
    uvm = fzeros(sh)
    uvm[I, J] = np.minimum(np.minimum(hmp[I, J], hmp[2:nx + 2, J]),
                           np.minimum(hmp[I, 2:ny + 2], hmp[2:nx + 2, 2:ny + 2]))
    halo_update(uvm, ew, ns, LOC_NECORNER, TYPE_SCALAR)
    tmask = np.asfortranarray((hmp > 0.5).astype(np.int32))
    umask = np.asfortranarray((uvm > 0.5).astype(np.int32))
    # init_evp, ice_dyn_evp.F90:503
    fcor = np.asfortranarray(2.0 * OMEGA * np.sin(ULAT))

    g.f.update(dxt=dxt, dyt=dyt, dxhy=dxhy, dyhx=dyhx, cxp=cxp, cyp=cyp, cxm=cxm, cym=cym,
               tarea=tarea, tarear=tarear, tinyarea=tinyarea, uarea=uarea, uarear=uarear,
               fcor=fcor, tmask=tmask, umask=umask, ULAT=ULAT, HTN=HTN, HTE=HTE, hm=hmp)
    return g


# ---------------------------------------------------------------------------
# Analytic grids: the real gx1/access-om grid blobs are absent.
# ---------------------------------------------------------------------------

def analytic_global(nx: int, ny: int, lat_s: float = -78.0, lat_n: float = 89.5,
                    lat_cap: float = 65.0, tfold: bool = False):
    """Lat-lon-like metrics, periodic in x, symmetric under the tripole fold
    (u-fold: i -> nx+1-i for T columns, i -> nx-i for U columns; T-fold (`tfold`): i -> nx+2-i for T
    columns, i -> nx+1-i for U columns).  Cell widths follow
    cos(lat) up to |lat| = lat_cap and stay constant poleward of it (a real
    tripole grid keeps cells finite by putting its poles on land)."""
    dlam = 2.0 * np.pi / nx
    dphi = np.deg2rad(lat_n - lat_s) / ny
    j = np.arange(1, ny + 1)
    ulat_1d = np.deg2rad(lat_s) + dphi * j  # latitude of the N edge / U points of row j
    lat_eff = np.clip(ulat_1d, -np.deg2rad(lat_cap), np.deg2rad(lat_cap))
    shift = 0.5 if tfold else 0.0      # the fold axis passes through T points instead of U points
    xt = (np.arange(1, nx + 1) - 0.5 - shift) / nx
    xu = (np.arange(1, nx + 1) - shift) / nx
    htn = RADIUS * np.cos(lat_eff)[None, :] * dlam * (1.0 + 0.05 * np.cos(4.0 * np.pi * xt))[:, None]
    tlat_1d = ulat_1d - 0.5 * dphi
    hte = RADIUS * dphi * (1.0 + 0.03 * np.sin(2.0 * tlat_1d))[None, :] * \
        (1.0 + 0.05 * np.cos(4.0 * np.pi * xu))[:, None]
    ulat = np.broadcast_to(ulat_1d[None, :], (nx, ny)).copy()
    ulon = np.broadcast_to((2.0 * np.pi * xu)[:, None], (nx, ny)).copy()
    return np.asfortranarray(htn), np.asfortranarray(hte), np.asfortranarray(ulat), np.asfortranarray(ulon)


def synthetic_land(nx: int, ny: int, ulat: np.ndarray, ulon: np.ndarray, ns: int,
                   realistic: bool) -> np.ndarray:
    """hm (1 ocean / 0 land).  `realistic`: two continents + Antarctica + polar caps;
    otherwise all ocean.  On a tripole the two pole points of the fold row
    (i = nx/2 and nx) sit on land patches in both variants, as on real grids."""
    hm = np.ones((nx, ny), dtype=np.float64, order="F")
    lat = np.rad2deg(ulat)
    lon = np.rad2deg(ulon)
    if realistic:
        hm[(lon > 20) & (lon < 60) & (lat > -35) & (lat < 72)] = 0
        hm[(lon > 190) & (lon < 250) & (lat > -55) & (lat < 68)] = 0
        hm[lat < -74] = 0
    if ns in (BND_TRIPOLE, BND_TRIPOLET):
        w = max(2, nx // 60)
        h = max(2, ny // 60)
        # the two pole points of the fold row: U columns nx/2 and nx (u-fold), T columns nx/2+1 and 1 (T-fold)
        for ic in ((nx // 2 + 1, 1) if ns == BND_TRIPOLET else (nx // 2, nx)):
            cols = np.arange(ic - w, ic + w + 1) % nx
            hm[cols, ny - h:] = 0
    else:
        hm[:, 0] = 0
        hm[:, -1] = 0
    return hm

"""Seeded synthetic ice / wind / ocean state for the EVP path (numpy only).

The reference's restart and forcing blobs are absent
(/root/reference/.MISSING_LARGE_BLOBS), so inputs follow SURVEY.md 8(d): the
reference's own default initial ice state (`init_state`/`set_state_var`,
/root/reference/source/ice_init.F90:1041-1113) modulated by a smooth
concentration field, an analytic two-cyclone wind stress and an ocean gyre.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import grid as G

NCAT = 5
CONFIGS = {
    # name: (nx, ny, ew, ns)   -- BASELINE.json configs
    "gx3": (100, 116, "cyclic", "open"),
    "gx1": (320, 384, "cyclic", "open"),
    "om1deg": (360, 300, "cyclic", "tripole"),
    "om025": (1440, 1080, "cyclic", "tripole"),
    "p01": (3600, 2700, "cyclic", "tripole"),
}


# dynamics time step per config (s): the elastic subcycle dte = dt/ndte must shrink with the
# grid spacing or the EVP iteration amplifies rounding noise (DESIGN.md "FMA contraction");
# 1 deg and coarser use the reference's dt = 3600 s (input_templates/gx3/ice_in), 0.25 deg
# dt = 1800 s and 0.1 deg dt = 600 s as the ACCESS-OM2 configurations of those grids do.
CONFIG_DT = {"gx3": 3600.0, "gx1": 3600.0, "om1deg": 3600.0, "om025": 1800.0, "p01": 600.0}


def hin_max(ncat: int = NCAT) -> np.ndarray:
    """Category bounds, kcatbound=0, kitd=1 (/root/reference/source/ice_itd.F90:162-186).
    Known answers: ice.log.Linux.LANL.coyote:185-190."""
    cc1 = 3.0 / ncat
    cc2 = 15.0 * cc1
    cc3 = 3.0
    h = np.zeros(ncat + 1)
    for n in range(1, ncat + 1):
        x1 = float(n - 1) / ncat
        h[n] = h[n - 1] + cc1 + cc2 * (1.0 + np.tanh(cc3 * (x1 - 1.0)))
    return h


def default_itd(ncat: int = NCAT):
    """ainit/hinit of the `ice_ic='default'` recipe (ice_init.F90:1041-1058)."""
    hm = hin_max(ncat)
    hbar = 3.0
    hinit = np.zeros(ncat)
    ainit = np.zeros(ncat)
    for n in range(1, ncat + 1):
        hinit[n - 1] = 0.5 * (hm[n - 1] + hm[n]) if n < ncat else hm[n - 1] + 1.0
        ainit[n - 1] = max(0.0, 2.0 * hbar * hinit[n - 1] - hinit[n - 1] ** 2)
    ainit = ainit / (ainit.sum() + G.PUNY / ncat)
    return ainit, hinit


@dataclass
class Case:
    name: str
    grid: G.Grid
    inputs: Dict[str, np.ndarray]     # aice vice vsno aice0 aicen vicen strairxT strairyT uocn vocn ss_tltx ss_tlty
    active_fraction: float


def _smooth(nx, ny, rng, kx=3, ky=2):
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    ph = rng.uniform(0, 2 * np.pi, size=4)
    s = (np.sin(2 * np.pi * kx * x + ph[0])[:, None] * np.sin(np.pi * ky * y + ph[1])[None, :] +
         0.5 * np.cos(2 * np.pi * (kx + 1) * x + ph[2])[:, None] * np.cos(np.pi * (ky + 1) * y + ph[3])[None, :])
    s = (s - s.min()) / (s.max() - s.min())
    return s


def make_case(name: str = "gx3", nx: Optional[int] = None, ny: Optional[int] = None,
              ew: Optional[str] = None, ns: Optional[str] = None, realistic: bool = False,
              seed: int = 20260101, gx3_fixture: Optional[str] = None) -> Case:
    """Build grid + inputs for a named BASELINE config or an explicit nx x ny."""
    if name in CONFIGS:
        cnx, cny, cew, cns = CONFIGS[name]
        nx = nx or cnx
        ny = ny or cny
        ew = ew or cew
        ns = ns or cns
    assert nx and ny and ew and ns
    ewb, nsb = G.BND_NAMES[ew], G.BND_NAMES[ns]
    rng = np.random.default_rng(seed)

    if name == "gx3" and gx3_fixture is not None:
        z = np.load(gx3_fixture)
        htn = np.asfortranarray(z["HTN"] * G.CM_TO_M)
        hte = np.asfortranarray(z["HTE"] * G.CM_TO_M)
        ulat = np.asfortranarray(z["ULAT"])
        ulon = np.asfortranarray(z["ULON"])
        hm = np.asfortranarray((z["KMT"] >= 1).astype(np.float64))
        if not realistic:
            hm[:, 1:-1] = 1.0
    else:
        htn, hte, ulat, ulon = G.analytic_global(nx, ny, tfold=(nsb == G.BND_TRIPOLET))
        hm = G.synthetic_land(nx, ny, ulat, ulon, nsb, realistic)
    grid = G.build_grid(htn, hte, ulat, hm, ewb, nsb)

    # --- ice state on the global grid -------------------------------------------------
    ainit, hinit = default_itd(NCAT)
    conc = 0.15 + 0.84 * _smooth(nx, ny, rng) + 0.01 * rng.random((nx, ny))
    conc = np.clip(conc, 0.0, 0.999)
    if realistic:
        lat = np.rad2deg(ulat)
        conc = np.where((lat > 70.0) | (lat < -60.0), conc, 0.0)
    conc = conc * (hm > 0.5)
    aicen_g = conc[:, :, None] * ainit[None, None, :]
    thick = 1.0 + 0.1 * (rng.random((nx, ny)) - 0.5)
    vicen_g = aicen_g * hinit[None, None, :] * thick[:, :, None]
    vsnon_g = np.minimum(aicen_g * 0.2, 0.2 * vicen_g)
    aice_g = aicen_g.sum(axis=2)
    vice_g = vicen_g.sum(axis=2)
    vsno_g = vsnon_g.sum(axis=2)
    aice0_g = 1.0 - aice_g

    sc = lambda a: G.scatter_global(np.asfortranarray(a), ewb, nsb, G.LOC_CENTER, G.TYPE_SCALAR)
    inputs: Dict[str, np.ndarray] = {}
    inputs["aice"] = sc(aice_g)
    inputs["vice"] = sc(vice_g)
    inputs["vsno"] = sc(vsno_g)
    inputs["aice0"] = sc(aice0_g)
    aicen = np.zeros((nx + 2, ny + 2, NCAT), order="F")
    vicen = np.zeros((nx + 2, ny + 2, NCAT), order="F")
    for n in range(NCAT):
        aicen[:, :, n] = sc(aicen_g[:, :, n])
        vicen[:, :, n] = sc(vicen_g[:, :, n])
    inputs["aicen"] = aicen
    inputs["vicen"] = vicen

    # --- forcing ------------------------------------------------------------------------
    x = (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    X, Y = np.meshgrid(x, y, indexing="ij")

    def cyclone(x0, y0, sgn):
        dx = (X - x0 + 0.5) % 1.0 - 0.5
        dy = Y - y0
        r2 = dx * dx + dy * dy
        amp = 10.0 * np.exp(-r2 / (2 * 0.12 ** 2)) * np.sqrt(r2) / 0.12 * 1.6487
        th = np.arctan2(dy, dx)
        return -sgn * amp * np.sin(th), sgn * amp * np.cos(th)

    ua1, va1 = cyclone(0.3, 0.8, 1.0)
    ua2, va2 = cyclone(0.7, 0.2, -1.0)
    ua = ua1 + ua2 + 2.0
    va = va1 + va2 - 1.0
    wsp = np.sqrt(ua * ua + va * va)
    strx = aice_g * 1.3 * 0.0015 * wsp * ua
    stry = aice_g * 1.3 * 0.0015 * wsp * va
    inputs["strairxT"] = sc(strx)
    inputs["strairyT"] = sc(stry)
    uo = 0.1 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + 0.02 * (rng.random((nx, ny)) - 0.5)
    vo = -0.1 * np.cos(2 * np.pi * X) * np.sin(np.pi * Y) + 0.02 * (rng.random((nx, ny)) - 0.5)
    scu = lambda a: G.scatter_global(np.asfortranarray(a), ewb, nsb, G.LOC_NECORNER, G.TYPE_VECTOR)
    inputs["uocn"] = scu(uo)
    inputs["vocn"] = scu(vo)
    inputs["ss_tltx"] = G.fzeros((nx + 2, ny + 2))
    inputs["ss_tlty"] = G.fzeros((nx + 2, ny + 2))

    active = float(((aice_g > 0.001) & (hm > 0.5)).sum()) / float(nx * ny)
    return Case(name=name, grid=grid, inputs=inputs, active_fraction=active)


def zero_state(nx_block: int, ny_block: int) -> Dict[str, np.ndarray]:
    """State after `init_evp` (/root/reference/source/ice_dyn_evp.F90:487-524)."""
    st = {k: G.fzeros((nx_block, ny_block)) for k in
          ("uvel", "vvel",
           "stressp_1", "stressp_2", "stressp_3", "stressp_4",
           "stressm_1", "stressm_2", "stressm_3", "stressm_4",
           "stress12_1", "stress12_2", "stress12_3", "stress12_4")}
    st["iceumask"] = G.fzeros((nx_block, ny_block), np.int32)
    return st

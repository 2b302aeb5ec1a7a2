"""Shared helpers of the parity tests: run the oracle and the CUDA path on the same case."""
import copy

import numpy as np

from cice4_b200 import evp as E
from cice4_b200 import synth

STATE = E.STATE_D + ["iceumask"]
OUT_CMP = ["strairx", "strairy", "strtltx", "strtlty", "strintx", "strinty", "strocnx", "strocny",
           "strocnxT", "strocnyT", "fm", "prs_sig", "divu", "shear", "rdg_conv", "rdg_shear"]


def oracle_steps(O, case, nsteps=1, strength_from_oracle=True, **pover):
    """nsteps consecutive evp calls with the oracle; returns (state, fields of last call, strengths)."""
    g = case.grid
    st = synth.zero_state(g.nx_block, g.ny_block)
    dt = pover.pop("dt", 3600.0)
    ndte = pover.pop("ndte", 120)
    p = O.make_params(dt=dt, ndte=ndte, **pover)
    f = None
    strengths = []
    for _ in range(nsteps):
        f, _sec = O.run_evp(g, case.inputs, st, p)
        strengths.append(f.strength_pre)    # what ice_strength returned, BEFORE evp's halo update of it
    return st, f, strengths, p


def cuda_steps(case, nsteps=1, strengths=None, layout=None, two_phase=False, want=None, **params):
    """Same through the C ABI.  strengths: list of host strength arrays as ice_strength returns them (evp
    halo-updates strength itself; on the T-fold that update is not idempotent, so pre-halo values are required) or
    None for the device ice_strength."""
    g = case.grid
    lay = layout or E.BlockLayout.single_block(g.nx, g.ny)
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    dt = params.pop("dt", 3600.0)
    dyn = E.IceDynEvp(lay, ew, ns, **params)
    gf = E.grid_fields_in_blocks(g, lay, ew, ns)
    dyn.init_evp(dt, gf)
    inputs = {k: E.split_blocks(v, lay, ew, ns) for k, v in case.inputs.items()}
    out = None
    for k in range(nsteps):
        s = None
        if strengths is not None:
            s = E.split_blocks(strengths[k], lay, ew, ns)
        out = dyn.evp(dt, inputs, strength=s, two_phase=two_phase, want=want)
    return dyn, out


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) if a.size else 0.0


def relerr(a, b):
    scale = float(np.max(np.abs(b)))
    return maxabs(a, b) / scale if scale > 0 else maxabs(a, b)


# ------------------------------------------------------------------------------------------------
# committed outputs of the reference itself (tests/golden/make_ref_golden.py)
# ------------------------------------------------------------------------------------------------
import glob
import os
from types import SimpleNamespace

from cice4_b200 import grid as G

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_GOLDEN = sorted(glob.glob(os.path.join(GOLDEN_DIR, "ref_evp_*.npz")))


def load_ref_golden(path):
    """-> namespace(grid, inputs, over, dt, ndte, nsteps, ref_state, ref_out)"""
    z = np.load(path)
    grid = G.Grid(int(z["meta_nx"]), int(z["meta_ny"]), int(z["meta_ew"]), int(z["meta_ns"]))
    inputs, over, ref_state, ref_out = {}, {}, {}, {}
    for k in z.files:
        a = z[k]
        if k.startswith("grid_"):
            grid.f[k[5:]] = np.asfortranarray(a)
        elif k.startswith("in_"):
            inputs[k[3:]] = np.asfortranarray(a)
        elif k.startswith("param_"):
            over[k[6:]] = a.item()
        elif k.startswith("ref_strength_"):
            pass
        elif k.startswith("ref_state_"):
            ref_state[k[10:]] = a
        elif k.startswith("ref_out_"):
            ref_out[k[8:]] = a
    return SimpleNamespace(grid=grid, inputs=inputs, over=over, dt=float(z["meta_dt"]), ndte=int(z["meta_ndte"]),
                           nsteps=int(z["meta_nsteps"]), ref_state=ref_state, ref_out=ref_out,
                           ref_strengths=[np.asfortranarray(z["ref_strength_%d" % k])
                                          for k in range(int(z["meta_nsteps"]))],
                           label=os.path.basename(path)[8:-4])


# which cells of a block `evp` defines, per field (the CUDA marshalling policies PACK_FULL / PACK_TNE_* /
# PACK_INT_* of csrc/evp_abi.cu): lo / hi = how far the region extends below ilo / beyond ihi
BLOCK_REGION = {"uvel": (1, 1), "vvel": (1, 1), "iceumask": (0, 0)}
BLOCK_REGION.update({n: (0, 1) for n in E.STATE_D[2:]})                                  # stresses: + N/E ghosts
BLOCK_REGION.update({n: (0, 1) for n in ("prs_sig", "divu", "shear", "rdg_conv", "rdg_shear")})
BLOCK_REGION.update({n: (0, 0) for n in ("strairx", "strairy", "strtltx", "strtlty", "strintx", "strinty",
                                         "strocnx", "strocny", "strocnxT", "strocnyT", "fm")})
BLOCK_REGION.update({"strength": (1, 1), "sicemass": (1, 1)})


def block_region_mismatches(name, got_blk, want_blk, layout):
    """blocks in which `got_blk` differs from `want_blk` on the cells evp defines for field `name`"""
    lo, hi = BLOCK_REGION[name]
    bad = []
    for b in range(layout.nblocks):
        i0, i1 = layout.ilo[b] - 1 - lo, layout.ihi[b] + hi
        j0, j1 = layout.jlo[b] - 1 - lo, layout.jhi[b] + hi
        if not np.array_equal(got_blk[i0:i1, j0:j1, b], want_blk[i0:i1, j0:j1, b]):
            bad.append(b)
    return bad


def energy_sums_fixed_order(grid, inputs, state, rhoi=917.0, rhos=330.0, puny=1e-11):
    """runtime_diags' kinetic energy / volume sums and rms ice speed (/root/reference/source/
    ice_diagnostics.F90:199-234) with the summation order of the device reduction (csrc/evp_aux.cu
    k_energy_rows / k_energy_total): per row, thread t of 256 adds columns t+1, t+257, ... in order, a binary
    tree (stride 128 .. 1) combines the 256 partial sums; the rows are combined the same way."""
    nx, ny = grid.nx, grid.ny
    I = (slice(1, nx + 1), slice(1, ny + 1))
    u, v = state["uvel"][I], state["vvel"][I]
    vice, vsno = inputs["vice"][I], inputs["vsno"][I]
    area = np.where(grid.f["tmask"][I] != 0, grid.f["tarea"][I], 0.0)
    south = grid.f["ULAT"][I] < -puny
    ke = 0.5 * (rhos * vsno + rhoi * vice) * (u * u + v * v)
    terms = [ke * area, vice * area, vsno * area]

    def tree(part):            # part: (256, n)
        s = 128
        part = part.copy()
        while s > 0:
            part[:s] = part[:s] + part[s:2 * s]
            s >>= 1
        return part[0]

    def strided(a):            # a: (n_items, n_lines): sum over items in the device order, per line
        n = a.shape[0]
        part = np.zeros((256, a.shape[1]))
        for k in range(0, n, 256):
            chunk = a[k:k + 256]
            part[:chunk.shape[0]] = part[:chunk.shape[0]] + chunk
        return tree(part)

    out = {}
    for name, t in zip(("ketot", "shmax", "snwmx"), terms):
        for hemi, m in (("n", ~south), ("s", south)):
            rows = strided(np.where(m, t, 0.0))          # one value per row j
            out[name + hemi] = float(strided(rows[:, None])[0])
    for hemi in ("n", "s"):
        ur = 2.0 * out["ketot" + hemi] / (rhoi * out["shmax" + hemi] + rhos * out["snwmx" + hemi] + puny)
        out["urms" + hemi] = float(np.sqrt(ur)) if ur > puny else 0.0
    return out

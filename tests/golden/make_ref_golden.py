"""Generate tests/golden/ref_evp_*.npz: inputs and OUTPUTS OF THE REFERENCE ITSELF for `evp(dt)`.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ref_golden.py

The reference is Fortran and the image has no Fortran compiler, so "the reference itself" is
oracle/_ref/libevp_ref_<variant>.so: the reference's own source text of `evp` and everything it calls
(source/ice_dyn_evp.F90, source/ice_grid.F90, source/ice_mechred.F90, drivers/*/ice_constants.F90),
translated statement by statement into C by oracle/f90_to_c.py at build time and compiled with gcc
-O2 -ffp-contract=off (oracle/build_ref.py).  Since round 2 the halo updates are the reference's own too: the
block loop of create_blocks, ice_blocksGetNbrID, the message loop of ice_HaloCreate, ice_HaloMsgCreate and
ice_HaloUpdate2DR8 / 2DI4 (serial/ice_boundary.F90) are translated as well; hand-written in oracle/ref_glue.c are only
the allocations, get_block / get_block_parameter and abort_ice.

Each fixture holds the complete problem (grid fields, inputs, initial state = init_evp zeros, the
run-time options) and the reference's state and outputs after `nsteps` consecutive calls, so the
tests need neither /root/reference nor the synthetic-input generator to reproduce them bit for bit:
  * tests/test_oracle_golden.py::test_oracle_matches_reference_golden   (oracle/evp_oracle.c, CPU)
  * tests/test_parity_gpu.py::test_cuda_matches_reference_golden        (CUDA path through the C ABI)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cice4_b200 import synth  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import oracle as O  # noqa: E402

# label -> (make_case kwargs, oracle-parameter overrides, dt, ndte, nsteps)
CASES = {
    "cice4_tripole_28x22": (dict(name="om1deg", nx=28, ny=22, realistic=True), dict(), 3600.0, 120, 2),
    "auscom_cyclic_open_26x20": (dict(name="x", nx=26, ny=20, ew="cyclic", ns="open"),
                                 dict(auscom=1, coupled=1, use_ocnslope=0, cosw=0.9063077870366499,
                                      sinw=0.42261826174069944), 3600.0, 120, 2),
    # T-fold: the fold as the reference's own halo executes it (the corner messages overwrite two rows of the
    # three-row tripole buffer, serial/ice_boundary.F90:3833-3848)
    "cice4_tripoleT_26x20": (dict(name="x", nx=26, ny=20, ew="cyclic", ns="tripoleT", realistic=True), dict(),
                             3600.0, 120, 2),
    "access_tripole_damping_24x20": (dict(name="om1deg", nx=24, ny=20),
                                     dict(auscom=1, coupled=1, use_ocnslope=1, access_wind=1, evp_damping=1),
                                     1800.0, 120, 2),
    "coupled_open_closed_hibler_22x18": (dict(name="x", nx=22, ny=18, ew="open", ns="closed"),
                                         dict(coupled=1, kstrength=0), 3600.0, 61, 1),
}


def add_variant_inputs(case, over):
    """inputs the synthetic recipe leaves at zero / absent but the AusCOM / ACCESS builds read"""
    inp = dict(case.inputs)
    g = case.grid
    rng = np.random.default_rng(7)
    if over.get("access_wind"):
        inp["strax"] = np.asfortranarray(0.9 * inp["strairxT"] + 0.001)
        inp["stray"] = np.asfortranarray(1.1 * inp["strairyT"] - 0.002)
    if over.get("coupled"):
        inp["ss_tltx"] = np.asfortranarray(1e-6 * (rng.random((g.nx_block, g.ny_block)) - 0.5))
        inp["ss_tlty"] = np.asfortranarray(1e-6 * (rng.random((g.nx_block, g.ny_block)) - 0.5))
    return inp


# block decompositions of one of the problems above: (label of the problem, bx, by)
BLOCK_CASES = [("cice4_tripole_28x22", 10, 8), ("cice4_tripoleT_26x20", 10, 8)]


def blocks_golden():
    """The reference run on a create_blocks decomposition: its state and outputs in BLOCK layout (the
    problem itself is the fixture of the same label)."""
    from cice4_b200 import evp as E
    for label, bx, by in BLOCK_CASES:
        kw, over, dt, ndte, nsteps = CASES[label]
        case = synth.make_case(**kw)
        g = case.grid
        inp = add_variant_inputs(case, over)
        p = O.make_params(dt=dt, ndte=ndte, **over)
        ew = {v: k for k, v in E.BND.items()}[g.ew]
        ns = {v: k for k, v in E.BND.items()}[g.ns]
        lay = E.BlockLayout.cartesian(g.nx, g.ny, bx, by)
        gfb = {k: np.asfortranarray(v) for k, v in E.grid_fields_in_blocks(g, lay, ew, ns).items() if v is not None}
        inb = {k: np.asfortranarray(E.split_blocks(v, lay, ew, ns)) for k, v in inp.items()}
        stb = {k: np.zeros(lay.shape, dtype=v.dtype, order="F") for k, v in synth.zero_state(3, 3).items()}
        strengths = []
        for _ in range(nsteps):
            fb = O.run_evp_ref_blocks(lay, g.ew, g.ns, gfb, inb, stb, p, dt)
            strengths.append(fb.strength_pre)
        out = {"meta_bx": bx, "meta_by": by}
        for k, a in enumerate(strengths):
            out["ref_strength_%d" % k] = a
        for k in O.STATE_D + ["iceumask"]:
            out["ref_state_" + k] = stb[k]
        for k in O.OUT_D:
            if k == "sicemass" and not over.get("auscom"):
                continue
            out["ref_out_" + k] = fb[k]
        path = os.path.join(HERE, "refblocks_%s_b%dx%d.npz" % (label, bx, by))
        np.savez_compressed(path, **out)
        print("%-36s %d blocks of %dx%d  %6.1f KB" % (label, lay.nblocks, bx, by, os.path.getsize(path) / 1024))


def main():
    build_ref.build()
    blocks_golden()
    for label, (kw, over, dt, ndte, nsteps) in CASES.items():
        case = synth.make_case(**kw)
        g = case.grid
        inp = add_variant_inputs(case, over)
        p = O.make_params(dt=dt, ndte=ndte, **over)
        st = synth.zero_state(g.nx_block, g.ny_block)
        f = None
        strengths = []
        for _ in range(nsteps):
            f = O.run_evp_ref(g, inp, st, p, dt)
            strengths.append(f.strength_pre)   # ice_strength's result before evp's halo update (input of the two-phase ABI)
        out = {"meta_nx": g.nx, "meta_ny": g.ny, "meta_ew": g.ew, "meta_ns": g.ns, "meta_dt": dt,
               "meta_ndte": ndte, "meta_nsteps": nsteps}
        for k, v in over.items():
            out["param_" + k] = v
        for k, v in g.f.items():
            out["grid_" + k] = v
        for k, v in inp.items():
            out["in_" + k] = v
        for k, a in enumerate(strengths):  # ice_strength of every call (input of the two-phase ABI)
            out["ref_strength_%d" % k] = a
        for k in O.STATE_D + ["iceumask"]:
            out["ref_state_" + k] = st[k]
        for k in O.OUT_D:
            if k == "sicemass" and not over.get("auscom"):
                continue
            out["ref_out_" + k] = f[k]
        path = os.path.join(HERE, "ref_evp_%s.npz" % label)
        np.savez_compressed(path, **out)
        print("%-36s |u|max %.4f  %6.1f KB" % (label, np.abs(st["uvel"]).max(), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

"""GPU parity tests: the CUDA path, called through the C ABI (ctypes over libevp_b200.so),
against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): after a full ndte = 120 timestep
    max |du|, |dv| <= 1e-10 m/s,   relative stress error <= 1e-10.
math_mode = 0 (unfused) is additionally required to be BIT-EXACT against the unfused oracle for
every field; math_mode = 1 (FMA-contracted) must stay within the tolerance.
"""
import numpy as np
import pytest

from cice4_b200 import evp as E
from cice4_b200 import synth
from conftest import GX3_FIXTURE
from helpers import OUT_CMP, REF_GOLDEN, STATE, cuda_steps, load_ref_golden, maxabs, oracle_steps, relerr

pytestmark = pytest.mark.gpu

TOL_U = 1e-10      # m/s
TOL_S = 1e-10      # relative


def _merge(dyn_or_arr, lay):
    return E.merge_blocks(dyn_or_arr, lay)


def _compare_exact(dyn, out, st, f, lay, names_out=OUT_CMP):
    bad = []
    for n in STATE:
        a = _merge(dyn.state[n], lay)
        if not np.array_equal(a, st[n]):
            bad.append((n, maxabs(a, st[n])))
    for n in names_out:
        a = _merge(out[n], lay)
        if not np.array_equal(a, f[n]):
            bad.append((n, maxabs(a, f[n])))
    assert not bad, f"not bit-exact: {bad}"


def _compare_tol(dyn, out, st, f, lay):
    for n in ("uvel", "vvel"):
        assert maxabs(_merge(dyn.state[n], lay), st[n]) <= TOL_U, n
    for n in STATE[2:14]:
        assert relerr(_merge(dyn.state[n], lay), st[n]) <= TOL_S, n
    assert np.array_equal(_merge(dyn.state["iceumask"], lay), st["iceumask"])
    for n in OUT_CMP:
        assert relerr(_merge(out[n], lay), f[n]) <= 1e-9, n


CASES = [
    ("gx3-real-grid", dict(name="gx3", realistic=True, gx3_fixture=GX3_FIXTURE)),
    ("gx3-dense", dict(name="gx3", realistic=False, gx3_fixture=GX3_FIXTURE)),
    ("tripole-64x48", dict(name="om1deg", nx=64, ny=48)),
    ("tripole-130x70-realistic", dict(name="om1deg", nx=130, ny=70, realistic=True)),
    ("cyclic-cyclic-40x33", dict(name="x", nx=40, ny=33, ew="cyclic", ns="cyclic")),
    ("open-open-37x29", dict(name="x", nx=37, ny=29, ew="open", ns="open")),
    ("tripoleT-64x48", dict(name="x", nx=64, ny=48, ew="cyclic", ns="tripoleT")),
    ("tripoleT-130x70-realistic", dict(name="x", nx=130, ny=70, ew="cyclic", ns="tripoleT", realistic=True)),
]


@pytest.mark.parametrize("path", REF_GOLDEN, ids=[p.split("ref_evp_")[-1][:-4] for p in REF_GOLDEN])
def test_cuda_matches_reference_golden(evp_lib, path):
    """The CUDA path against the committed OUTPUTS OF THE REFERENCE ITSELF
    (tests/golden/make_ref_golden.py): unfused math is bit-exact for the state and every output,
    FMA mode within the north_star tolerance.  No oracle involved."""
    from types import SimpleNamespace
    c = load_ref_golden(path)
    inputs = dict(c.inputs)
    cpar = {}
    for k, v in c.over.items():
        cpar[{"auscom": "hemisphere_turning", "coupled": "coupled_tilt", "access_wind": "wind_from_strax"}.get(k, k)] = v
    if c.over.get("access_wind"):  # the ABI takes strax/stray through the strairxT/yT pointers
        inputs["strairxT"], inputs["strairyT"] = inputs.pop("strax"), inputs.pop("stray")
    inputs = {k: v for k, v in inputs.items() if k in E.INPUT_D}
    case = SimpleNamespace(grid=c.grid, inputs=inputs)
    lay = E.BlockLayout.single_block(c.grid.nx, c.grid.ny)
    want = [n for n in c.ref_out if n != "strength"]
    # (1) the reference's ice_strength of every call supplied by the host, as the Fortran shim does in
    #     its two-phase mode: unfused math must reproduce the reference bit for bit
    dyn, out = cuda_steps(case, nsteps=c.nsteps, strengths=c.ref_strengths, want=want, dt=c.dt, ndte=c.ndte,
                          math_mode=0, **cpar)
    bad = [n for n in STATE if not np.array_equal(_merge(dyn.state[n], lay), c.ref_state[n])]
    bad += [n for n in want if not np.array_equal(_merge(out[n], lay), c.ref_out[n])]
    assert not bad, f"CUDA path differs from the reference in {bad}"
    dyn.finalize()
    # (2) FMA-contracted math and (3) ice_strength on the device (device exp): within the tolerance
    for par in (dict(strengths=c.ref_strengths, math_mode=1), dict(strengths=None, math_mode=0)):
        dyn, out = cuda_steps(case, nsteps=c.nsteps, want=want + ["strength"], dt=c.dt, ndte=c.ndte, **par, **cpar)
        for n in ("uvel", "vvel"):
            assert maxabs(_merge(dyn.state[n], lay), c.ref_state[n]) <= TOL_U, n
        for n in STATE[2:14]:
            assert relerr(_merge(dyn.state[n], lay), c.ref_state[n]) <= TOL_S, n
        assert np.array_equal(_merge(dyn.state["iceumask"], lay), c.ref_state["iceumask"])
        assert relerr(_merge(out["strength"], lay), c.ref_out["strength"]) <= 1e-12
        dyn.finalize()


@pytest.mark.parametrize("label,nb", [("cice4_tripole_28x22", 9), ("cice4_tripoleT_26x20", 9)])
def test_cuda_block_layout_matches_reference_blocks(evp_lib, label, nb):
    """The caller's block layout against THE REFERENCE RUN ON THE SAME BLOCKS (committed outputs of the
    translated reference on a 10 x 8 create_blocks decomposition of the 28 x 22 tripole problem, padded
    edge blocks included): on the cells evp defines in every block -- velocities with their ghost ring,
    stresses and T-point diagnostics with their north/east ghost cells, U-point outputs on the physical
    cells -- the CUDA path returns the reference's block arrays bit for bit."""
    import os
    from types import SimpleNamespace
    from helpers import BLOCK_REGION, GOLDEN_DIR, block_region_mismatches
    base = [p for p in REF_GOLDEN if p.endswith("ref_evp_%s.npz" % label)][0]
    c = load_ref_golden(base)
    z = np.load(os.path.join(GOLDEN_DIR, "refblocks_%s_b10x8.npz" % label))
    lay = E.BlockLayout.cartesian(c.grid.nx, c.grid.ny, int(z["meta_bx"]), int(z["meta_by"]))
    assert lay.nblocks == nb
    case = SimpleNamespace(grid=c.grid, inputs={k: v for k, v in c.inputs.items() if k in E.INPUT_D})
    want = [n for n in BLOCK_REGION if "ref_out_" + n in z.files]
    dyn, out = cuda_steps(case, nsteps=c.nsteps, strengths=c.ref_strengths, layout=lay, want=want, dt=c.dt,
                          ndte=c.ndte, math_mode=0, **c.over)
    bad = {}
    for n in BLOCK_REGION:
        if "ref_state_" + n in z.files:
            b = block_region_mismatches(n, dyn.state[n], z["ref_state_" + n], lay)
        elif "ref_out_" + n in z.files:
            b = block_region_mismatches(n, out[n], z["ref_out_" + n], lay)
        else:
            continue
        if b:
            bad[n] = b
    assert not bad, f"blocks that differ from the reference's: {bad}"
    dyn.finalize()


@pytest.mark.parametrize("kernel", [0, 32768], ids=["default-kernel", "plane-kernel"])
@pytest.mark.parametrize("label,kw", CASES, ids=[c[0] for c in CASES])
def test_bit_exact_vs_oracle_two_steps(oracle, evp_lib, label, kw, kernel):
    """Cold start + warm second call (the timed configuration), unfused math: bit-exact.  Once with the
    kernel the library picks itself (the strip-tiled TMA-fed kernel on slabs this short) and once with the
    plane kernel forced (kernel_variant bit 15), which is what the tall slabs run."""
    case = synth.make_case(**kw)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, kernel_variant=kernel)
    # north-south cyclic domains and the T-fold run the plane kernels (the in-kernel fold is the u-fold)
    assert dyn.info()["tiled"] == (0 if kernel or label.startswith("cyclic-cyclic") or "tripoleT" in label else 1)
    _compare_exact(dyn, out, st, f, lay)
    assert np.abs(st["uvel"]).max() > 1e-3   # the case is not trivially zero


@pytest.mark.parametrize("label,kw", CASES[:4], ids=[c[0] for c in CASES[:4]])
def test_fma_mode_within_tolerance(oracle, evp_lib, label, kw):
    case = synth.make_case(**kw)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=1)
    _compare_tol(dyn, out, st, f, lay)


WARP_STRIPS, CTA_STRIPS, FINISH_KERNEL = 1048576, 2097152, 4194304   # kernel_variant bits 20, 21, 22


@pytest.mark.parametrize("variant", [WARP_STRIPS, CTA_STRIPS, WARP_STRIPS + FINISH_KERNEL, CTA_STRIPS + FINISH_KERNEL,
                                     WARP_STRIPS + 16, WARP_STRIPS + 4],
                         ids=["warp-strips", "cta-strips", "warp-strips-finish-kernel", "cta-strips-finish-kernel",
                              "warp-strips-2plane", "warp-strips-fold-kernel"])
@pytest.mark.parametrize("label,kw", CASES, ids=[c[0] for c in CASES])
def test_plane_kernel_strip_and_finish_variants(oracle, evp_lib, label, kw, variant):
    """The plane kernel of 128-thread CTAs with one strip per warp (bit 20: shuffles, no row barrier) or one strip
    per CTA (bit 21: exchange line + barrier), evp_finish as an epilogue of the last subcycle kernel (default) or
    as its own kernel (bit 22): every combination is bit-exact, cold + warm call, on every domain type."""
    case = synth.make_case(**kw)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, tile_threads=128,
                          kernel_variant=32768 + variant)
    info = dyn.info()
    assert info["tiled"] == 0 and info["threads"] == 128
    assert info["warp_strips"] == (1 if variant & WARP_STRIPS else 0), info
    if variant & WARP_STRIPS:
        assert info["strip_w"] % 4 == 0 and info["strip_w"] <= 124, info
    # the fused finish needs the final velocities inside the kernel: not with a fold outside it (T-fold, bit 2)
    fold_outside = "tripoleT" in label or (label.startswith("tripole-") and (variant & 4))
    assert info["finish_fused"] == (0 if (variant & FINISH_KERNEL) or fold_outside else 1), info
    _compare_exact(dyn, out, st, f, lay)


@pytest.mark.parametrize("variant", [0, FINISH_KERNEL], ids=["finish-fused", "finish-kernel"])
def test_finish_epilogue_tiled_and_variants(oracle, evp_lib, variant):
    """evp_finish inside the last subcycle kernel on the strip-tiled layout (fold row completed by the fold), with the
    AusCOM hemisphere turning (sign flip where fm < 0), in the FMA build, behind a block layout, with the TMA-staged and
    the persistent kernel -- against the same call with the separate k_finish launch and against the oracle."""
    case = synth.make_case("om1deg", nx=130, ny=70, realistic=True)
    lay1 = E.BlockLayout.single_block(130, 70)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, kernel_variant=2048 + variant)
    assert dyn.info()["tiled"] == 1 and dyn.info()["finish_fused"] == (0 if variant else 1)
    _compare_exact(dyn, out, st, f, lay1)
    lay = E.BlockLayout.cartesian(130, 70, 33, 18)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay, kernel_variant=variant)
    _compare_exact(dyn, out, st, f, lay)
    for kv, par in ((256, {}), (128, dict(tile_threads=64)), (1024, dict(tile_threads=128))):
        dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, kernel_variant=32768 + kv + variant, **par)
        _compare_exact(dyn, out, st, f, lay1)
    pover = dict(auscom=1, coupled=1, use_ocnslope=0, cosw=0.9063077870366499, sinw=0.42261826174069944)
    cpar = {{"auscom": "hemisphere_turning", "coupled": "coupled_tilt"}.get(k, k): v for k, v in pover.items()}
    st, f, strengths, _ = oracle_steps(oracle, case, **dict(pover))
    for kv in (0, 32768 + WARP_STRIPS):
        dyn, out = cuda_steps(case, strengths=strengths, kernel_variant=kv + variant,
                              **(dict(tile_threads=128) if kv else {}), **cpar)
        _compare_exact(dyn, out, st, f, lay1)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=1, kernel_variant=variant)
    _compare_tol(dyn, out, st, f, lay1)


@pytest.mark.parametrize("threads,rows,variant", [(64, 5, 0), (128, 7, 4), (256, 0, 0), (128, 1000, 0),
                                                  (128, 0, 16), (64, 9, 16 + 4), (128, 0, 64), (256, 11, 64), (128, 0, 1024),
                                                  (128, 3, WARP_STRIPS), (128, 0, WARP_STRIPS + 64), (128, 1, CTA_STRIPS),
                                                  (128, 0, WARP_STRIPS + 1024), (128, 5, WARP_STRIPS + 1024 + 16)])
def test_tiling_invariance(oracle, evp_lib, threads, rows, variant):
    """Strip width, rows per CTA and the prefetch variant must not change a single bit."""
    case = synth.make_case("om1deg", nx=300, ny=90)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, ndte=40)
    lay = E.BlockLayout.single_block(300, 90)
    dyn, out = cuda_steps(case, strengths=strengths, ndte=40, tile_threads=threads, tile_rows=rows,
                          kernel_variant=variant)
    _compare_exact(dyn, out, st, f, lay)


PERSIST_CASES = [
    ("gx3-real-grid", dict(name="gx3", realistic=True, gx3_fixture=GX3_FIXTURE), dict()),
    ("tripole-130x70-realistic", dict(name="om1deg", nx=130, ny=70, realistic=True), dict()),
    ("tripole-300x90-t64", dict(name="om1deg", nx=300, ny=90), dict(tile_threads=64)),
    ("tripole-300x200-t128-2plane", dict(name="om1deg", nx=300, ny=200), dict(tile_threads=128, kernel_variant=128 + 16)),
    ("open-open-37x29", dict(name="x", nx=37, ny=29, ew="open", ns="open"), dict()),
    ("cyclic-open-50x40-odd-ndte", dict(name="gx3", nx=50, ny=40, ew="cyclic", ns="open"), dict(ndte=7)),
    ("closed-closed-31x30", dict(name="x", nx=31, ny=30, ew="closed", ns="closed"), dict(ndte=2)),
]


@pytest.mark.parametrize("label,kw,par", PERSIST_CASES, ids=[c[0] for c in PERSIST_CASES])
def test_persistent_kernel_bit_exact(oracle, evp_lib, label, kw, par):
    """kernel_variant bit 7: the whole ndte loop as ONE cooperative launch whose CTAs synchronise with
    their neighbours only (per-CTA epochs, no grid barrier).  Two consecutive calls, bit-exact."""
    case = synth.make_case(**kw)
    par = dict(par)
    ndte = par.pop("ndte", 120)
    variant = par.pop("kernel_variant", 128)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, ndte=ndte)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, ndte=ndte, kernel_variant=variant, **par)
    assert dyn.timings()["subcycle_launches"] == 1, "the persistent kernel was not used"
    _compare_exact(dyn, out, st, f, lay)
    dyn.finalize()


@pytest.mark.parametrize("variant", [256, 512], ids=["3-rows-deep-2-ctas", "2-rows-deep-3-ctas"])
@pytest.mark.parametrize("label,kw", [CASES[0], CASES[3], CASES[5], ("tripole-300x200", dict(name="om1deg", nx=300, ny=200))],
                         ids=["gx3-real-grid", "tripole-130x70-realistic", "open-open-37x29", "tripole-300x200"])
def test_tma_staged_kernel_bit_exact(oracle, evp_lib, label, kw, variant):
    """kernel_variant bits 8 / 9: the T-row planes reach the SM through TMA bulk copies into shared
    memory (mbarrier pipeline) instead of per-thread register prefetch.  Same arithmetic: bit-exact."""
    case = synth.make_case(**kw)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, kernel_variant=variant)
    _compare_exact(dyn, out, st, f, lay)
    dyn.finalize()


TILED_CASES = CASES + [
    ("tripole-300x200", dict(name="om1deg", nx=300, ny=200)),
    ("cyclic-open-50x40-odd-ndte", dict(name="gx3", nx=50, ny=40, ew="cyclic", ns="open")),
    ("closed-closed-31x30", dict(name="x", nx=31, ny=30, ew="closed", ns="closed")),
    ("tripole-62x40", dict(name="om1deg", nx=62, ny=40)),       # nx a multiple of the strip width
    ("tripole-63x21", dict(name="om1deg", nx=63, ny=21)),       # a last strip of one U column
    ("open-cyclic-32x9", dict(name="x", nx=32, ny=9, ew="cyclic", ns="open")),
]


@pytest.mark.parametrize("variant", [2048 + 4096, 2048], ids=["2-stages-3-ctas", "3-stages-2-ctas"])
@pytest.mark.parametrize("label,kw", TILED_CASES, ids=[c[0] for c in TILED_CASES])
def test_tiled_kernel_bit_exact(oracle, evp_lib, label, kw, variant):
    """kernel_variant bit 11: the strip-tiled layout with the warp-autonomous TMA-fed kernel (one bulk copy
    per tile row, warp shuffles, no CTA barrier).  Same arithmetic as the plane kernels: cold + warm call
    bit-exact against the strict oracle, every state and output field."""
    case = synth.make_case(**kw)
    ndte = 7 if "odd-ndte" in label else 120
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, ndte=ndte)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, ndte=ndte, kernel_variant=variant)
    info = dyn.info()
    if label.startswith("cyclic-cyclic") or "tripoleT" in label:
        assert info["tiled"] == 0      # north-south cyclic / T-fold: the plane kernels run
    else:
        assert info["tiled"] == 1 and info["strip_w"] == 31 and info["stages"] == (2 if variant & 4096 else 3), info
    _compare_exact(dyn, out, st, f, lay)
    # the resident loop continues from the tiles; the state comes back through download_state
    if info["tiled"]:
        dyn.subcycle_resident(1)
        d = dyn.diagnostics()
        assert np.isfinite(d["umaxn"])
    dyn.finalize()


FUSED_CASES = [
    ("gx3-real-grid", dict(name="gx3", realistic=True, gx3_fixture=GX3_FIXTURE), dict()),
    ("gx3-dense", dict(name="gx3", realistic=False, gx3_fixture=GX3_FIXTURE), dict()),
    ("tripole-64x48", dict(name="om1deg", nx=64, ny=48), dict()),
    ("tripole-130x70-realistic", dict(name="om1deg", nx=130, ny=70, realistic=True), dict()),
    ("tripole-300x200", dict(name="om1deg", nx=300, ny=200), dict()),
    ("tripole-112x40-one-cta", dict(name="om1deg", nx=112, ny=40), dict()),
    ("tripole-113x33", dict(name="om1deg", nx=113, ny=33), dict()),
    ("tripole-225x30-rows3", dict(name="om1deg", nx=225, ny=30), dict(tile_rows=3)),
    ("tripole-170x64-rows2", dict(name="om1deg", nx=170, ny=64), dict(tile_rows=2)),
    ("open-open-70x40", dict(name="x", nx=70, ny=40, ew="open", ns="open"), dict()),
    ("closed-closed-90x35", dict(name="x", nx=90, ny=35, ew="closed", ns="closed"), dict(tile_rows=4)),
    ("cyclic-open-150x40-odd-ndte", dict(name="gx3", nx=150, ny=40, ew="cyclic", ns="open"), dict(ndte=7)),
    ("cyclic-closed-84x50-ndte2", dict(name="x", nx=84, ny=50, ew="cyclic", ns="closed"), dict(ndte=2)),
    ("open-tripole-100x36", dict(name="x", nx=100, ny=36, ew="open", ns="tripole"), dict(ndte=12)),
]


@pytest.mark.parametrize("label,kw,par", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_two_subcycle_kernel_bit_exact(oracle, evp_lib, label, kw, par):
    """kernel_variant bit 17: TWO subcycles per launch (csrc/evp_fused.cuh) -- the second subcycle runs two rows
    behind the first one on values kept in shared memory; east-west seam, closed / open boundaries and the tripole
    fold of the intermediate velocities included.  Same arithmetic as the one-subcycle kernels: cold + warm call
    bit-exact against the strict oracle, every state and output field."""
    par = dict(par)
    case = synth.make_case(**kw)
    ndte = par.pop("ndte", 120)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, ndte=ndte)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, ndte=ndte, kernel_variant=131072, **par)
    info = dyn.info()
    assert info["fused"] == 1 and info["threads"] == 128 and info["strip_w"] <= 112, info
    _compare_exact(dyn, out, st, f, lay)
    dyn.subcycle_resident(1)
    assert np.isfinite(dyn.diagnostics()["umaxn"])
    dyn.finalize()


def test_fused_kernel_fma_mode_and_blocks(oracle, evp_lib):
    """The two-subcycle kernel behind a block layout, with evp_damping, and its FMA build within the tolerance."""
    case = synth.make_case("om1deg", nx=130, ny=70, realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.cartesian(130, 70, 23, 19)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay, kernel_variant=131072)
    assert dyn.info()["fused"] == 1
    _compare_exact(dyn, out, st, f, lay)
    dyn.finalize()
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=1, kernel_variant=131072)
    _compare_tol(dyn, out, st, f, E.BlockLayout.single_block(130, 70))
    dyn.finalize()
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, evp_damping=1)
    dyn, out = cuda_steps(case, strengths=strengths, kernel_variant=131072, evp_damping=1)
    _compare_exact(dyn, out, st, f, E.BlockLayout.single_block(130, 70))
    dyn.finalize()


@pytest.mark.parametrize("rows", [1, 3, 1000])
def test_tiled_kernel_tiling_invariance(oracle, evp_lib, rows):
    case = synth.make_case("om1deg", nx=300, ny=90)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, ndte=40)
    lay = E.BlockLayout.single_block(300, 90)
    dyn, out = cuda_steps(case, strengths=strengths, ndte=40, tile_rows=rows, kernel_variant=2048)
    if rows >= 2:
        assert dyn.info()["tiled"] == 1
    _compare_exact(dyn, out, st, f, lay)


def test_tiled_kernel_variants_and_blocks(oracle, evp_lib):
    """The tiled kernel behind the reference's block layouts, the AusCOM variant, evp_damping and the FMA build."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.cartesian(64, 48, 23, 19)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay, kernel_variant=2048)
    assert dyn.info()["tiled"] == 1
    _compare_exact(dyn, out, st, f, lay)
    lay1 = E.BlockLayout.single_block(64, 48)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=1, kernel_variant=2048)
    _compare_tol(dyn, out, st, f, lay1)
    for pover in (dict(evp_damping=1), dict(auscom=1, coupled=1, use_ocnslope=0, cosw=0.9063077870366499, sinw=0.42261826174069944)):
        st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, **dict(pover))
        cpar = {{"auscom": "hemisphere_turning", "coupled": "coupled_tilt"}.get(k, k): v for k, v in pover.items()}
        dyn, out = cuda_steps(case, strengths=strengths, kernel_variant=2048, **cpar)
        _compare_exact(dyn, out, st, f, lay1)


def test_graph_and_stream_launch_agree(oracle, evp_lib):
    case = synth.make_case("om1deg", nx=64, ny=48)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, ndte=30)
    lay = E.BlockLayout.single_block(64, 48)
    for use_graph in (0, 1):
        dyn, out = cuda_steps(case, strengths=strengths, ndte=30, use_graph=use_graph)
        _compare_exact(dyn, out, st, f, lay)


def test_odd_ndte(oracle, evp_lib):
    case = synth.make_case("gx3", nx=50, ny=40, ew="cyclic", ns="open")
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, ndte=7)
    lay = E.BlockLayout.single_block(50, 40)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, ndte=7)
    _compare_exact(dyn, out, st, f, lay)


@pytest.mark.parametrize("bx,by", [(16, 12), (64, 7), (10, 48), (23, 19)])
def test_block_decomposition_invariance(oracle, evp_lib, bx, by):
    """Reference block layouts (incl. padded edge blocks) give the single-block answer
    (doc/cicedoc.pdf 4.6: bit-for-bit regardless of decomposition)."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.cartesian(64, 48, bx, by)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay)
    _compare_exact(dyn, out, st, f, lay)
    # ghost cells of every block hold the neighbour's values after the final halo update
    ublk = dyn.state["uvel"]
    want = E.split_blocks(st["uvel"], lay, "cyclic", "tripole")
    for b in range(lay.nblocks):
        ni, nj = lay.ihi[b] - lay.ilo[b] + 3, lay.jhi[b] - lay.jlo[b] + 3
        np.testing.assert_array_equal(ublk[:ni, :nj, b], want[:ni, :nj, b])


def test_tripoleT_block_layout(oracle, evp_lib):
    """T-fold behind a reference block decomposition (padded edge blocks), two calls: bit-exact."""
    case = synth.make_case("x", nx=64, ny=48, ew="cyclic", ns="tripoleT", realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    lay = E.BlockLayout.cartesian(64, 48, 23, 19)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay)
    _compare_exact(dyn, out, st, f, lay)
    u = _merge(dyn.state["uvel"], lay)
    for i in range(1, 65):      # the fold as the reference executes it: the ghost row mirrors the row below the top row
        assert u[i, 49] == -u[65 - i, 47]


def test_land_block_elimination(oracle, evp_lib):
    """Blocks without ocean cells left out of the caller's layout (the reference's distribution assigns
    them to no task): the cells no block covers are land to the library, and every remaining block gets
    the single-block result bit for bit (tests/test_oracle_vs_ref.py shows the reference does too)."""
    from helpers import BLOCK_REGION, block_region_mismatches
    case = synth.make_case("om1deg", nx=96, ny=64, realistic=True)
    g = case.grid
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    full = E.BlockLayout.cartesian(g.nx, g.ny, 8, 8)
    land = full.land_blocks(g.f["tmask"])
    assert len(land) >= 5
    lay = full.without(land)
    names = [n for n in BLOCK_REGION if n != "sicemass"]
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, layout=lay, want=[n for n in names if n not in STATE])
    bad = {}
    for n in names:
        got = dyn.state[n] if n in dyn.state else out[n]
        b = block_region_mismatches(n, got, E.split_blocks(st[n] if n in st else f[n], lay, "cyclic", "tripole"), lay)
        if b:
            bad[n] = b
    assert not bad, bad
    dyn.finalize()


def test_two_phase_prep_run(oracle, evp_lib):
    """evp_b200_prep + evp_b200_run (host ice_strength in between, as the Fortran shim does)."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    lay = E.BlockLayout.single_block(64, 48)
    dyn, out = cuda_steps(case, strengths=strengths, two_phase=True)
    _compare_exact(dyn, out, st, f, lay)
    np.testing.assert_array_equal(_merge(out["icetmask"], lay), f["icetmask"])


def test_device_ice_strength(oracle, evp_lib):
    """strength == NULL: ice_strength on the device.  exp() differs from glibc in the last
    bits, so the comparison is by tolerance (relative 1e-12 on strength itself)."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    lay = E.BlockLayout.single_block(64, 48)
    dyn, out = cuda_steps(case, strengths=None)
    assert relerr(_merge(out["strength"], lay), f["strength"]) <= 1e-12
    _compare_tol(dyn, out, st, f, lay)


@pytest.mark.parametrize("pover", [
    dict(evp_damping=1),
    dict(auscom=1, coupled=1, use_ocnslope=0, cosw=0.9063077870366499, sinw=0.42261826174069944),
    dict(auscom=1, coupled=1, use_ocnslope=1),
    dict(coupled=1),
], ids=["evp_damping", "auscom-turning", "auscom-ocnslope", "coupled"])
def test_namelist_and_cpp_variants(oracle, evp_lib, pover):
    case = synth.make_case("om1deg", nx=48, ny=40)
    rng = np.random.default_rng(5)
    case.inputs["ss_tltx"][...] = 1e-6 * rng.standard_normal(case.inputs["ss_tltx"].shape)
    case.inputs["ss_tlty"][...] = 1e-6 * rng.standard_normal(case.inputs["ss_tlty"].shape)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1, **dict(pover))
    lay = E.BlockLayout.single_block(48, 40)
    cpar = {}
    for k, v in pover.items():
        cpar[{"auscom": "hemisphere_turning", "coupled": "coupled_tilt"}.get(k, k)] = v
    want = OUT_CMP + (["sicemass"] if pover.get("auscom") else [])
    dyn, out = cuda_steps(case, strengths=strengths, want=want + ["strength"], **cpar)
    _compare_exact(dyn, out, st, f, lay, names_out=want)


def test_principal_stress(oracle, evp_lib):
    case = synth.make_case("om1deg", nx=48, ny=40)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    lay = E.BlockLayout.single_block(48, 40)
    dyn, out = cuda_steps(case, strengths=strengths, want=["prs_sig", "sig1", "sig2"])
    s1, s2 = oracle.principal_stress(st["stressp_1"], st["stressm_1"], st["stress12_1"], f["prs_sig"])
    np.testing.assert_array_equal(_merge(out["sig1"], lay), s1)
    np.testing.assert_array_equal(_merge(out["sig2"], lay), s2)
    g1, g2 = dyn.principal_stress(dyn.state["stressp_1"], dyn.state["stressm_1"], dyn.state["stress12_1"],
                                  out["prs_sig"])
    np.testing.assert_array_equal(_merge(g1, lay), s1)
    np.testing.assert_array_equal(_merge(g2, lay), s2)


FULL_SIZE = [
    # BASELINE.json configs at their full size, with the dt the bench uses (synth.CONFIG_DT)
    ("gx1-320x384", dict(name="gx1"), 3600.0),
    ("om1deg-360x300-tripole", dict(name="om1deg"), 3600.0),
    ("om025-1440x1080-tripole", dict(name="om025"), 1800.0),
]


@pytest.mark.parametrize("label,kw,dt", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_vs_oracle(oracle, evp_lib, label, kw, dt):
    """The benchmarked configurations themselves (gx1, access-om 1 deg, access-om 0.25 deg at full size,
    dense mask as in bench.py): cold start + warm second call through the C ABI against the strict
    serial oracle (source/ice_dyn_evp.F90:347-404 for the loop).  math_mode 0: every state and output
    field BIT-EXACT; math_mode 1 (FMA-contracted): max |du|,|dv| <= 1e-10 m/s and relative stress error
    <= 1e-10 (BASELINE north_star tolerance) -- except where contraction alone moves the CPU code by more
    than that: the same C code built with and without contraction (gcc -ffp-contract=fast vs off) differs
    by 1.6e-10 in the stresses of the warm gx1 call (rigid cells with Delta near tinyarea amplify
    rounding), so where that happens the bound is twice the measured freedom of the reference arithmetic
    itself on the same case (computed here, printed with -s)."""
    case = synth.make_case(**kw)
    assert synth.CONFIG_DT[kw["name"]] == dt
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, dt=dt)
    lay = E.BlockLayout.single_block(case.grid.nx, case.grid.ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=0, dt=dt)
    _compare_exact(dyn, out, st, f, lay)
    assert np.abs(st["uvel"]).max() > 1e-3
    dyn.finalize()
    # FMA freedom of the CPU arithmetic itself on this case: contracted build of the same oracle source
    g = case.grid
    st_c = synth.zero_state(g.nx_block, g.ny_block)
    pc = oracle.make_params(dt=dt, ndte=120, kind="fast")
    for _ in range(2):
        oracle.run_evp(g, case.inputs, st_c, pc, lib_kind="fast")
    freedom = max(relerr(st_c[n], st[n]) for n in STATE[2:14])
    freedom_u = max(maxabs(st_c[n], st[n]) for n in ("uvel", "vvel"))
    tol_s = max(TOL_S, 2.0 * freedom)
    tol_u = max(TOL_U, 2.0 * freedom_u)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, math_mode=1, dt=dt)
    worst_u = max(maxabs(_merge(dyn.state[n], lay), st[n]) for n in ("uvel", "vvel"))
    assert worst_u <= tol_u, f"velocity error {worst_u:.3e} m/s > {tol_u:.3e} (CPU contraction freedom {freedom_u:.3e})"
    worst = max(relerr(_merge(dyn.state[n], lay), st[n]) for n in STATE[2:14])
    assert worst <= tol_s, f"relative stress error {worst:.3e} > {tol_s:.3e} (CPU contraction freedom {freedom:.3e})"
    print(f"\n[{label}] math_mode=1 vs strict oracle after cold + warm call: max|du| {worst_u:.3e} m/s (CPU contraction "
          f"freedom {freedom_u:.3e}), relative stress error {worst:.3e} (CPU contraction freedom {freedom:.3e})")
    assert np.array_equal(_merge(dyn.state["iceumask"], lay), st["iceumask"])
    dyn.finalize()


def test_ieee_sequences_match_sqrt_and_division(evp_lib):
    """csrc/evp_ieee.cuh (sqrt_fast / rcp_refined / div_fast, the interleaved straight-line expansions the
    subcycle kernel uses for :1095-1098, :1131-1134, :1426-1427) against sqrt() and operator/ on 2^27
    (1.3e8) generated operands: random bit patterns, EVP magnitudes, operands around each range check
    (high word 0x03500000 / 0x7ff00000 for sqrt, |n| ~ 2^-969 and tiny quotients for the division),
    subnormal, huge, zero, exact and one-ulp-around-powers-of-two cases.  Whenever the fast path accepts
    an operand its result must equal the IEEE result in every bit."""
    r = E.IceDynEvp.selftest_ieee(1 << 27, seed=20260101)
    assert r["n"] == 1 << 27
    assert r["sqrt_mismatch"] == 0 and r["div_mismatch"] == 0 and r["div_shared_rcp_mismatch"] == 0, r
    # the fast path must actually be exercised (and rejected for the special operands)
    assert 0.3 * r["n"] < r["sqrt_fast"] < r["n"] and 0.25 * r["n"] < r["div_fast"] < r["n"], r
    r2 = E.IceDynEvp.selftest_ieee(1 << 22, seed=7)
    assert r2["sqrt_mismatch"] == 0 and r2["div_mismatch"] == 0 and r2["div_shared_rcp_mismatch"] == 0, r2


def test_full_size_properties_om025(evp_lib):
    """BASELINE metric size (1440 x 1080, tripole): size-independent properties instead of the
    (slow) oracle: tripole symmetry, masked cells stay zero, unfused vs FMA within tolerance,
    tiling invariance bit-exact, finite and plausible magnitudes."""
    case = synth.make_case("om025", realistic=True)
    nx, ny = 1440, 1080
    lay = E.BlockLayout.single_block(nx, ny)
    res = {}
    for tag, par in (("a", dict(math_mode=0)), ("b", dict(math_mode=0, tile_threads=256, tile_rows=19)),
                     ("c", dict(math_mode=1))):
        dyn, out = cuda_steps(case, strengths=None, want=["strength", "divu", "prs_sig"],
                              dt=synth.CONFIG_DT["om025"], **par)
        res[tag] = ({k: v[:, :, 0].copy() for k, v in dyn.state.items()}, out)
        dyn.finalize()
    sa, sb, sc = res["a"][0], res["b"][0], res["c"][0]
    u = sa["uvel"]
    assert np.isfinite(u).all() and 0.01 < np.abs(u).max() < 3.0
    for i in range(1, nx // 2):
        assert u[i, ny] == -u[nx - i, ny]
    I = (slice(1, nx + 1), slice(1, ny + 1))
    assert np.all(u[I][sa["iceumask"][I] == 0] == 0.0)
    for n in STATE:
        assert np.array_equal(sa[n], sb[n]), f"tiling changed {n}"
    assert maxabs(sa["uvel"], sc["uvel"]) <= TOL_U and maxabs(sa["vvel"], sc["vvel"]) <= TOL_U
    for n in STATE[2:14]:
        assert relerr(sc[n], sa[n]) <= TOL_S, n


def test_two_gpus_bit_exact_vs_oracle(evp_lib):
    """y-slabs on 2 GPUs with the NCCL row exchange (skipped on a single-GPU box; run by hand with
    `gpurun --gpus 2`, see tests/multigpu_parity.py)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multigpu_parity.py"),
           "--realistic"]
    # graph of launches; persistent; tiled; plane kernel with warp strips / CTA strips
    for extra in ([], ["--kernel-variant", "128"], ["--kernel-variant", "2048"],
                  ["--kernel-variant", "32768", "--tile-threads", "128"],
                  ["--kernel-variant", str(32768 + 2097152), "--tile-threads", "128"]):
        r = subprocess.run(cmd + extra, capture_output=True, text=True, timeout=240)
        assert r.returncode == 0 and "BIT-EXACT" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_plane_metric_path_and_fallback(oracle, evp_lib):
    """HTE/HTN given: rows whose metric planes equal the init_grid2 formulas bit for bit are
    re-derived in the kernel (2 planes streamed instead of 8); rows that do not (here: a perturbed
    dxt value, and row 1 whose dxt is extrapolated, ice_grid.F90:1199-1202) fall back to the 8
    planes.  Either way the result is bit-exact."""
    case = synth.make_case("om1deg", nx=64, ny=48)
    lay = E.BlockLayout.single_block(64, 48)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    dyn, out = cuda_steps(case, strengths=strengths, kernel_variant=16)
    rows_on = dyn.timings()["reserved"]
    assert rows_on >= 45, rows_on
    _compare_exact(dyn, out, st, f, lay)
    dyn0, out0 = cuda_steps(case, strengths=strengths, kernel_variant=0)   # default: 8 planes
    assert dyn0.timings()["reserved"] == rows_on   # still verified, just not used
    _compare_exact(dyn0, out0, st, f, lay)
    # perturb one metric value: that row must drop out of the fast path, results follow the arrays
    case.grid.f["dxt"][20, 30] *= 1.0 + 1e-12
    st2, f2, strengths2, _ = oracle_steps(oracle, case, nsteps=1)
    dyn2, out2 = cuda_steps(case, strengths=strengths2, kernel_variant=16)
    assert dyn2.timings()["reserved"] == rows_on - 1
    _compare_exact(dyn2, out2, st2, f2, lay)
    assert not np.array_equal(st2["uvel"], st["uvel"])


def test_stress_residency(oracle, evp_lib):
    """state_residency = 1: stresses stay on the device between calls, the host
    arrays are refreshed only by download_state; invalidate_device_state re-uploads them."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    lay = E.BlockLayout.cartesian(64, 48, 32, 24)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=3)
    dyn, out = cuda_steps(case, nsteps=3, strengths=strengths, layout=lay, state_residency=1)
    assert np.all(dyn.state["stressp_1"] == 0.0)          # never downloaded so far
    for n in ("uvel", "vvel", "iceumask"):                 # these still round-trip
        assert np.array_equal(_merge(dyn.state[n], lay), st[n]), n
    dyn.download_state()
    _compare_exact(dyn, out, st, f, lay)
    # the caller overwrites its arrays (restartfile): the next call must use them
    for n in STATE[2:14]:
        dyn.state[n][...] *= 0.5
    st_h = {n: _merge(dyn.state[n], lay).copy(order="F") for n in STATE}
    dyn.invalidate_device_state()
    p = oracle.make_params()
    f2, _ = oracle.run_evp(case.grid, case.inputs, st_h, p)
    inputs = {k: E.split_blocks(v, lay, "cyclic", "tripole") for k, v in case.inputs.items()}
    out2 = dyn.evp(3600.0, inputs, strength=E.split_blocks(f2["strength"], lay, "cyclic", "tripole"))
    dyn.download_state()
    _compare_exact(dyn, out2, st_h, f2, lay)


def test_full_state_residency_and_velocity_handoff(oracle, evp_lib):
    """state_residency = 2: nothing of the state travels; the transport hand-off gets
    uvel / vvel through download_velocity or as device planes; download_state brings everything back."""
    import torch
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    lay = E.BlockLayout.cartesian(64, 48, 32, 24)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=3)
    dyn, out = cuda_steps(case, nsteps=3, strengths=strengths, layout=lay, state_residency=2)
    assert np.all(dyn.state["stressp_1"] == 0.0) and np.all(dyn.state["uvel"] == 0.0)   # never downloaded so far
    for n in OUT_CMP:                                          # outputs do travel
        assert np.array_equal(_merge(out[n], lay), f[n]), n
    dyn.download_velocity()
    for n in ("uvel", "vvel"):
        assert np.array_equal(_merge(dyn.state[n], lay), st[n]), n
    pu, pv, pitch, nrows = dyn.device_velocity()
    assert pitch >= 66 and nrows == 50
    # the device planes themselves: plane (i, j) is element (i, j) of the padded single-block array
    buf = torch.empty(pitch * nrows, dtype=torch.float64, device="cuda")
    from cuda import cudart
    err, = cudart.cudaMemcpy(buf.data_ptr(), pu, 8 * pitch * nrows, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    torch.cuda.synchronize()
    plane = buf.cpu().numpy().reshape(nrows, pitch)[:, :66].T
    assert np.array_equal(plane, st["uvel"])
    dyn.download_state()
    _compare_exact(dyn, out, st, f, lay)
    dyn.finalize()


def test_step_device_pointers(oracle, evp_lib):
    """evp_b200_step_device: inputs, state and outputs as DEVICE arrays in the block layout (a caller whose
    other components already live on the GPU); bit-exact against the host-pointer path's oracle result."""
    import torch
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    lay = E.BlockLayout.cartesian(64, 48, 23, 19)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2)
    dyn = E.IceDynEvp(lay, "cyclic", "tripole")
    dyn.init_evp(3600.0, E.grid_fields_in_blocks(case.grid, lay, "cyclic", "tripole"))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).cuda()
    inputs = {k: dev(E.split_blocks(v, lay, "cyclic", "tripole")) for k, v in case.inputs.items()
              if k in E.INPUT_D and k not in ("aice0", "aicen", "vicen")}
    n = lay.nx_block * lay.ny_block * lay.max_blocks
    state = {k: torch.zeros(n, dtype=torch.float64, device="cuda") for k in E.STATE_D}
    state["iceumask"] = torch.zeros(n, dtype=torch.int32, device="cuda")
    outs = {k: torch.zeros(n, dtype=torch.float64, device="cuda") for k in OUT_CMP}
    for k in range(2):
        sdev = dev(E.split_blocks(strengths[k], lay, "cyclic", "tripole"))
        dyn.evp_device({k2: v.data_ptr() for k2, v in inputs.items()}, {k2: v.data_ptr() for k2, v in state.items()},
                       {k2: v.data_ptr() for k2, v in outs.items()}, strength=sdev.data_ptr())
    back = lambda t: t.cpu().numpy().reshape(lay.shape, order="F")
    for k in STATE:
        assert np.array_equal(_merge(back(state[k]), lay), st[k]), k
    for k in OUT_CMP:
        assert np.array_equal(_merge(back(outs[k]), lay), f[k]), k
    dyn.finalize()


def test_energy_diagnostics_fixed_order(oracle, evp_lib):
    """Kinetic energy, ice / snow volume and rms ice speed per hemisphere (runtime_diags,
    ice_diagnostics.F90:199-234) reduced on the device in a fixed order: bit-identical to the same order in
    numpy, and equal to the reference's sequential global_sum order within rounding."""
    from helpers import energy_sums_fixed_order
    for kw in (dict(name="om1deg", nx=300, ny=90), dict(name="gx3", realistic=True, gx3_fixture=GX3_FIXTURE)):
        case = synth.make_case(**kw)
        st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
        dyn, out = cuda_steps(case, strengths=strengths)
        d = dyn.diagnostics_energy()
        ref = energy_sums_fixed_order(case.grid, case.inputs, st)
        for k in ref:
            assert d[k] == ref[k], (k, d[k], ref[k])
        # the reference's own (sequential, j outer / i inner) order agrees to rounding
        g = case.grid
        I = (slice(1, g.nx + 1), slice(1, g.ny + 1))
        ke = 0.5 * (330.0 * case.inputs["vsno"][I] + 917.0 * case.inputs["vice"][I]) * (st["uvel"][I] ** 2 + st["vvel"][I] ** 2)
        area_n = np.where((g.f["tmask"][I] != 0) & (g.f["ULAT"][I] >= -1e-11), g.f["tarea"][I], 0.0)
        seq = float(np.sum((ke * area_n).T.ravel()))
        assert abs(d["ketotn"] - seq) <= 1e-12 * abs(seq)
        assert d["urmsn"] > 0.0
        dyn.finalize()


def test_principal_stress_per_block(oracle, evp_lib):
    case = synth.make_case("om1deg", nx=48, ny=40)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    dyn, out = cuda_steps(case, strengths=strengths, want=["prs_sig"])
    s1, s2 = oracle.principal_stress(st["stressp_1"], st["stressm_1"], st["stress12_1"], f["prs_sig"])
    g1, g2 = dyn.principal_stress_block(st["stressp_1"], st["stressm_1"], st["stress12_1"], f["prs_sig"])
    np.testing.assert_array_equal(g1, s1)
    np.testing.assert_array_equal(g2, s2)


def test_edge_no_ice_and_all_land(oracle, evp_lib):
    """Empty inputs: an ice-free ocean and an all-land domain give all-zero dynamics, and the state
    left by a previous call is cleared exactly as evp_prep2 does (:827-840, :886-894)."""
    case = synth.make_case("om1deg", nx=48, ny=36)
    lay = E.BlockLayout.single_block(48, 36)
    # previous step with ice, then the ice disappears
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    for k in ("aice", "vice", "vsno", "aicen", "vicen", "strairxT", "strairyT"):
        case.inputs[k][...] = 0.0
    case.inputs["aice0"][...] = 1.0
    p = oracle.make_params()
    f2, _ = oracle.run_evp(case.grid, case.inputs, st, p)
    assert np.all(st["uvel"] == 0.0) and np.all(st["stressp_1"] == 0.0) and np.all(st["iceumask"] == 0)
    case_ice = synth.make_case("om1deg", nx=48, ny=36)
    dyn, _ = cuda_steps(case_ice, nsteps=1, strengths=strengths)
    inputs = {k: E.split_blocks(v, lay, "cyclic", "tripole") for k, v in case.inputs.items()}
    out = dyn.evp(3600.0, inputs, strength=E.split_blocks(f2["strength"], lay, "cyclic", "tripole"))
    _compare_exact(dyn, out, st, f2, lay)
    # all land
    case2 = synth.make_case("gx3", nx=20, ny=16, ew="cyclic", ns="open")
    case2.grid.f["tmask"][...] = 0
    case2.grid.f["umask"][...] = 0
    st3, f3, s3, _ = oracle_steps(oracle, case2, nsteps=1)
    dyn3, out3 = cuda_steps(case2, nsteps=1, strengths=s3)
    _compare_exact(dyn3, out3, st3, f3, E.BlockLayout.single_block(20, 16))
    assert np.all(dyn3.state["uvel"] == 0.0)


@pytest.mark.parametrize("nx,ny,ew,ns", [(8, 6, "cyclic", "tripole"), (5, 4, "open", "open"), (6, 3, "cyclic", "open"),
                                         (257, 3, "cyclic", "tripole"), (129, 47, "cyclic", "cyclic")])
def test_edge_small_and_ragged_sizes(oracle, evp_lib, nx, ny, ew, ns):
    """Minimum and ragged sizes: fewer columns than one strip, strips that do not divide nx, slabs of
    2-3 rows, a fold row shorter than a warp."""
    case = synth.make_case("x", nx=nx, ny=ny, ew=ew, ns=ns)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=2, ndte=24)
    lay = E.BlockLayout.single_block(nx, ny)
    dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, ndte=24)
    _compare_exact(dyn, out, st, f, lay)


def test_restart_like_state(oracle, evp_lib):
    """Non-zero incoming state incl. ghost cells (what restartfile leaves: scattered velocities,
    N/E ghost stresses filled, W/S ghost stresses zeroed, source/ice_restart.F90:427-539)."""
    case = synth.make_case("om1deg", nx=64, ny=48, realistic=True)
    lay = E.BlockLayout.cartesian(64, 48, 16, 24)
    rng = np.random.default_rng(11)
    st0 = synth.zero_state(66, 50)
    for n in STATE[:2]:
        st0[n][...] = 0.05 * rng.standard_normal(st0[n].shape)
    for n in STATE[2:14]:
        st0[n][...] = 1.0e3 * rng.standard_normal(st0[n].shape)
        st0[n][0, :] = 0.0
        st0[n][:, 0] = 0.0
    st0["iceumask"][1:-1, 1:-1] = (rng.random((64, 48)) > 0.5).astype(np.int32)
    st_o = {k: v.copy(order="F") for k, v in st0.items()}
    p = oracle.make_params()
    f, _ = oracle.run_evp(case.grid, case.inputs, st_o, p)
    dyn = E.IceDynEvp(lay, "cyclic", "tripole")
    dyn.init_evp(3600.0, E.grid_fields_in_blocks(case.grid, lay, "cyclic", "tripole"))
    for k in STATE:
        dyn.state[k][...] = E.split_blocks(st0[k], lay, "cyclic", "tripole")
    inputs = {k: E.split_blocks(v, lay, "cyclic", "tripole") for k, v in case.inputs.items()}
    out = dyn.evp(3600.0, inputs, strength=E.split_blocks(f["strength"], lay, "cyclic", "tripole"))
    _compare_exact(dyn, out, st_o, f, lay)


def test_device_diagnostics(oracle, evp_lib):
    """max ice speed / max strength per hemisphere (runtime_diags, ice_diagnostics.F90:294-346)
    reduced on the device; maxima are order-independent, so the comparison is exact."""
    case = synth.make_case("om1deg", nx=64, ny=48)
    st, f, strengths, _ = oracle_steps(oracle, case, nsteps=1)
    lay = E.BlockLayout.single_block(64, 48)
    dyn, out = cuda_steps(case, strengths=strengths)
    d = dyn.diagnostics()
    I = (slice(1, 65), slice(1, 49))
    speed = np.sqrt(st["uvel"][I] ** 2 + st["vvel"][I] ** 2)
    north = case.grid.f["ULAT"][I] >= -1e-11
    assert d["umaxn"] == speed[north].max() and d["umaxs"] == speed[~north].max()
    assert d["pmaxn"] == (f["strength"][I][north] / 1000.0).max()
    assert d["pmaxs"] == (f["strength"][I][~north] / 1000.0).max()
    assert 0.01 < d["umaxn"] < 2.0

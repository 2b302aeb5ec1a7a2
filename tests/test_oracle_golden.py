"""CPU tests: pin the oracle (and the host-side builders) against every known answer the
reference offers for this path, and check the invariants the reference guarantees.

Known answers: /root/reference/ice.log.Linux.LANL.coyote
  :101-119 grid record min/max, :181-183 dte / tdamp, :185-190 hin_max.
The reference ships no golden vectors for the EVP outputs themselves; those are pinned by the
committed OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_evp_*.npz, made by
tests/golden/make_ref_golden.py from the reference's own Fortran text, machine-translated to C at
build time because the image has no Fortran compiler) and, in the build container, by direct
comparison with that translated reference (tests/test_oracle_vs_ref.py).
"""
import numpy as np
import pytest

from cice4_b200 import grid as G
from cice4_b200 import synth
from conftest import GX3_FIXTURE
from helpers import REF_GOLDEN, load_ref_golden


@pytest.mark.parametrize("path", REF_GOLDEN, ids=[p.split("ref_evp_")[-1][:-4] for p in REF_GOLDEN])
def test_oracle_matches_reference_golden(oracle, path):
    """oracle/evp_oracle.c reproduces the reference's own evp() outputs BIT FOR BIT (state after
    `nsteps` calls + every output field), all CPP variants and boundary types of the fixtures."""
    c = load_ref_golden(path)
    p = oracle.make_params(dt=c.dt, ndte=c.ndte, **c.over)
    st = synth.zero_state(c.grid.nx_block, c.grid.ny_block)
    f = None
    for _ in range(c.nsteps):
        f, _sec = oracle.run_evp(c.grid, c.inputs, st, p)
    bad = [n for n, a in c.ref_state.items() if not np.array_equal(st[n], a)]
    bad += [n for n, a in c.ref_out.items() if not np.array_equal(f[n], a)]
    assert not bad, f"oracle differs from the reference in {bad}"
    assert len(c.ref_state) == 15 and len(c.ref_out) >= 17
    assert np.abs(c.ref_state["uvel"]).max() > 1e-2


def test_reference_golden_fixtures_present():
    assert len(REF_GOLDEN) >= 4


def test_set_evp_parameters_known_answers(oracle):
    # ice.log.Linux.LANL.coyote:181-183: dt = 3600, dte = 30, tdamp = 1296
    p = oracle.make_params(dt=3600.0, ndte=120)
    assert 1.0 / p.dtei == pytest.approx(30.0, rel=1e-15)
    assert p.ecci == 0.25
    assert 0.36 * 3600.0 == pytest.approx(1296.0, rel=1e-15)
    dte2T = 30.0 / (2.0 * 0.36 * 3600.0)
    assert p.dte2T == dte2T
    assert p.denom1 == 1.0 / (1.0 + dte2T)
    assert p.denom2 == 1.0 / (1.0 + dte2T * 4.0)
    assert p.rcon == 1230.0 * 0.36 * 3600.0 * (p.dtei * p.dtei)


def test_host_set_evp_parameters_matches_oracle(oracle):
    from cice4_b200.evp import BlockLayout, IceDynEvp
    dyn = IceDynEvp(BlockLayout.single_block(8, 8), ndte=120)
    h = dyn.set_evp_parameters(3600.0)
    p = oracle.make_params(dt=3600.0, ndte=120)
    for k in ("dtei", "ecci", "dte2T", "denom1", "denom2", "rcon"):
        assert h[k] == getattr(p, k), k
    assert h["dte"] == 30.0 and h["tdamp"] == pytest.approx(1296.0, rel=1e-15)


def test_hin_max_known_answers():
    # ice.log.Linux.LANL.coyote:185-190
    want = [0.644507216819426, 1.39143349757630, 2.47017938195989, 4.56728791885049, 9.33384181586817]
    got = synth.hin_max(5)[1:]
    np.testing.assert_allclose(got, want, rtol=2e-14)


def test_gx3_grid_fixture_known_answers():
    # ice.log.Linux.LANL.coyote:101-119 (values printed with 15 significant digits)
    z = np.load(GX3_FIXTURE)
    assert z["ULAT"].shape == (100, 116)
    assert z["ULAT"].min() == pytest.approx(-1.36148077740934, rel=1e-14)
    assert z["ULAT"].max() == pytest.approx(1.56905100613449, rel=1e-14)
    assert z["HTN"].min() == pytest.approx(372424.009403068, rel=1e-14)
    assert z["HTN"].max() == pytest.approx(40023891.5205005, rel=1e-14)
    assert z["HTE"].min() == pytest.approx(9205227.47129144, rel=1e-14)
    assert z["HTE"].max() == pytest.approx(25237286.5505761, rel=1e-14)
    assert z["KMT"].max() == 25 and z["KMT"].min() == 0
    assert int((z["KMT"] >= 1).sum()) == 8006


@pytest.mark.parametrize("ew,ns", [("cyclic", "open"), ("cyclic", "tripole"), ("cyclic", "cyclic"),
                                   ("open", "open"), ("cyclic", "tripoleT"), ("open", "tripoleT")])
@pytest.mark.parametrize("loc,kind", [(G.LOC_CENTER, G.TYPE_SCALAR), (G.LOC_CENTER, G.TYPE_VECTOR),
                                      (G.LOC_NECORNER, G.TYPE_VECTOR), (G.LOC_NECORNER, G.TYPE_SCALAR),
                                      (G.LOC_NFACE, G.TYPE_SCALAR), (G.LOC_EFACE, G.TYPE_VECTOR)])
def test_numpy_halo_equals_oracle_halo(oracle, ew, ns, loc, kind):
    """The host-side numpy halo (grid construction) and the oracle's C halo restate the same
    reference routine independently; they must agree bit for bit."""
    rng = np.random.default_rng(7)
    a = np.asfortranarray(rng.standard_normal((18, 13)))
    b = a.copy(order="F")
    G.halo_update(a, G.BND_NAMES[ew], G.BND_NAMES[ns], loc, kind)
    oracle.halo_r8(b, G.BND_NAMES[ew], G.BND_NAMES[ns], loc, kind)
    np.testing.assert_array_equal(a, b)


def test_tripole_fold_semantics(oracle):
    """NE-corner vector update symmetrises and overwrites the top physical row."""
    nx, ny = 16, 6
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.standard_normal((nx + 2, ny + 2)))
    a0 = a.copy()
    oracle.halo_r8(a, G.BND_CYCLIC, G.BND_TRIPOLE, G.LOC_NECORNER, G.TYPE_VECTOR)
    top = a[1:nx + 1, ny]       # physical top row, global i = 1..nx
    # u(i) = -u(nx - i) for i = 1..nx/2-1 after the update
    for i in range(1, nx // 2):
        assert top[i - 1] == -top[nx - i - 1]
    # the self-mapped points flip sign every update
    assert top[nx // 2 - 1] == -a0[nx // 2, ny]
    assert top[nx - 1] == -a0[nx, ny]
    # ghost row = -(row ny-1) mirrored with the U-point offset
    for i in range(1, nx):
        assert a[i, ny + 1] == -a0[nx - i, ny - 1]
    # rows below the top two are untouched
    np.testing.assert_array_equal(a[1:nx + 1, 1:ny - 1], a0[1:nx + 1, 1:ny - 1])
    # a second update leaves the symmetrised pairs unchanged up to the double sign flip
    b = a.copy(order="F")
    oracle.halo_r8(b, G.BND_CYCLIC, G.BND_TRIPOLE, G.LOC_NECORNER, G.TYPE_VECTOR)
    np.testing.assert_array_equal(b[1:nx // 2, ny], a[1:nx // 2, ny])


def test_tripoleT_fold_semantics(oracle):
    """T-fold AS THE REFERENCE EXECUTES IT (serial/ice_boundary.F90:725-773 with the address lists of
    ice_HaloCreate / ice_HaloMsgCreate, established by running the reference's own translated halo,
    tests/test_oracle_vs_ref.py::test_oracle_halo_equals_reference_halo): the 'north' message fills the three-row
    buffer with the rows ny-2 .. ny, then the 'northeast' / 'northwest' messages of the tripole blocks (:3833-3848,
    written for the two-row u-fold buffer and last in the list) overwrite buffer rows 1 and 2 with the rows ny-1, ny.
    So for a NE-corner vector the top physical row becomes the sign-flipped mirror image of ITSELF and the ghost
    row that of the row below it (not of rows ny-1 / ny-2, which a geometric T-fold would give); for a centre
    field the degenerate top row is symmetrised about the two pole T points (columns 1 and nx/2+1 keep their
    values) and the ghost row mirrors the RAW top row."""
    nx, ny = 16, 7
    rng = np.random.default_rng(4)
    a = np.asfortranarray(rng.standard_normal((nx + 2, ny + 2)))
    a0 = a.copy()
    oracle.halo_r8(a, G.BND_CYCLIC, G.BND_TRIPOLET, G.LOC_NECORNER, G.TYPE_VECTOR)
    for i in range(1, nx + 1):                       # U column i <-> nx + 1 - i
        assert a[i, ny] == -a0[nx + 1 - i, ny]
        assert a[i, ny + 1] == -a0[nx + 1 - i, ny - 1]
    np.testing.assert_array_equal(a[1:nx + 1, 1:ny], a0[1:nx + 1, 1:ny])
    c = np.asfortranarray(rng.standard_normal((nx + 2, ny + 2)))
    c0 = c.copy()
    oracle.halo_r8(c, G.BND_CYCLIC, G.BND_TRIPOLET, G.LOC_CENTER, G.TYPE_SCALAR)
    top = c[1:nx + 1, ny]
    for i in range(2, nx // 2 + 1):                  # T column i <-> nx + 2 - i, averaged
        assert top[i - 1] == top[nx + 2 - i - 1] == 0.5 * (c0[i, ny] + c0[nx + 2 - i, ny])
    assert top[0] == c0[1, ny] and top[nx // 2] == c0[nx // 2 + 1, ny]     # the pole points map onto themselves
    for i in range(2, nx + 1):
        assert c[i, ny + 1] == c0[nx + 2 - i, ny]                          # the raw top row (buffer row 2)
    # integer fields average with nint (:1318): a 0/1 mask becomes the OR of the pair
    import ctypes as C
    m = np.asfortranarray(rng.integers(0, 2, (nx + 2, ny + 2)).astype(np.int32))
    m0 = m.copy()
    g = oracle.make_grid(nx + 2, ny + 2, G.BND_CYCLIC, G.BND_TRIPOLET)
    oracle.lib().orc_halo_i4(m.ctypes.data_as(oracle.c_ip), C.byref(g), G.LOC_CENTER, G.TYPE_SCALAR, 0)
    for i in range(2, nx // 2 + 1):
        assert m[i, ny] == m[nx + 2 - i, ny] == (m0[i, ny] | m0[nx + 2 - i, ny])


def test_tripoleT_evp_invariants(oracle):
    """One evp step on a T-fold grid, with the fold as the reference executes it (see
    test_tripoleT_fold_semantics): after the final halo update the ghost row is the sign-flipped mirror image of
    the row below the top physical row (which no halo update touches)."""
    case = synth.make_case("x", nx=48, ny=30, ew="cyclic", ns="tripoleT")
    st, f = _run(oracle, case)
    u, v = st["uvel"], st["vvel"]
    nx, ny = 48, 30
    assert np.abs(u).max() > 1e-3
    for i in range(1, nx + 1):
        assert u[i, ny + 1] == -u[nx + 1 - i, ny - 1] and v[i, ny + 1] == -v[nx + 1 - i, ny - 1]


def _run(oracle, case, **kw):
    st = synth.zero_state(case.grid.nx_block, case.grid.ny_block)
    p = oracle.make_params(**kw)
    f, _ = oracle.run_evp(case.grid, case.inputs, st, p)
    return st, f


def test_zero_forcing_stays_zero(oracle):
    case = synth.make_case("om1deg", nx=24, ny=20)
    for k in ("strairxT", "strairyT", "uocn", "vocn"):
        case.inputs[k][...] = 0.0
    st, f = _run(oracle, case, ndte=20)
    assert np.all(st["uvel"] == 0.0) and np.all(st["vvel"] == 0.0)
    for n in range(1, 5):
        assert np.all(st[f"stress12_{n}"] == 0.0)


def test_uniform_grid_uniform_flow_has_zero_strain(oracle):
    """Uniform rectangular grid, uniform velocity field and matching ocean current, no wind:
    strain rates and Delta vanish, stresses decay by denom1/denom2 only (SURVEY 8c iii)."""
    nx, ny = 20, 16
    htn = np.full((nx, ny), 3.0e4, order="F")
    hte = np.full((nx, ny), 3.0e4, order="F")
    ulat = np.zeros((nx, ny), order="F")     # no Coriolis: fm = 0, geostrophic tilt = 0
    hm = np.ones((nx, ny), order="F")
    g = G.build_grid(htn, hte, ulat, hm, G.BND_CYCLIC, G.BND_CYCLIC)
    case = synth.make_case("x", nx=nx, ny=ny, ew="cyclic", ns="cyclic")
    case.grid = g
    for k in ("strairxT", "strairyT"):
        case.inputs[k][...] = 0.0
    case.inputs["uocn"][...] = 0.05
    case.inputs["vocn"][...] = -0.02
    ainit, hinit = synth.default_itd()
    for n in range(synth.NCAT):
        case.inputs["aicen"][:, :, n] = 0.9 * ainit[n]
        case.inputs["vicen"][:, :, n] = 0.9 * ainit[n] * hinit[n]
    case.inputs["aice"][...] = case.inputs["aicen"].sum(axis=2)
    case.inputs["vice"][...] = case.inputs["vicen"].sum(axis=2)
    case.inputs["vsno"][...] = 0.1
    case.inputs["aice0"][...] = 1.0 - case.inputs["aice"]
    st = synth.zero_state(nx + 2, ny + 2)
    st["stressp_1"][...] = -100.0
    st["stressm_2"][...] = 10.0
    p = oracle.make_params(ndte=10)
    f, _ = oracle.run_evp(g, case.inputs, st, p)
    I = (slice(1, nx + 1), slice(1, ny + 1))
    assert np.all(f["iceumask"][I] == 1)
    # (m*(m*u))/(m*m) rounds, identically in every cell: the field stays exactly uniform
    np.testing.assert_allclose(st["uvel"][I], 0.05, rtol=1e-14)
    np.testing.assert_allclose(st["vvel"][I], -0.02, rtol=1e-14)
    assert np.ptp(st["uvel"][I]) == 0.0 and np.ptp(st["vvel"][I]) == 0.0
    assert np.all(f["divu"] == 0.0) and np.all(f["shear"] == 0.0)
    want_p = -100.0
    want_m = 10.0
    for _ in range(10):
        want_p = (want_p + 0.0) * p.denom1
        want_m = (want_m + 0.0) * p.denom2
    np.testing.assert_array_equal(st["stressp_1"][I], want_p)
    np.testing.assert_array_equal(st["stressm_2"][I], want_m)


def test_tripole_symmetry_after_evp(oracle):
    case = synth.make_case("om1deg", nx=32, ny=24)
    st, f = _run(oracle, case, ndte=30)
    nx, ny = 32, 24
    u = st["uvel"]
    for i in range(1, nx // 2):
        assert u[i, ny] == -u[nx - i, ny]
    assert np.isfinite(u).all()
    # plausible magnitudes (ice.log.Linux.LANL.coyote:296-300,385-386: |u| <~ 0.3 m/s)
    assert 0.01 < np.abs(u).max() < 2.0


def test_gx3_real_grid_runs_and_is_plausible(oracle):
    case = synth.make_case("gx3", realistic=True, gx3_fixture=GX3_FIXTURE)
    st, f = _run(oracle, case, ndte=120)
    assert np.isfinite(st["uvel"]).all() and np.isfinite(st["stressp_1"]).all()
    assert 0.01 < np.abs(st["uvel"]).max() < 2.0
    assert 1e3 < f["strength"].max() < 3e5   # log: max strength ~113 kN/m
    # land and ice-free cells stay masked out
    I = (slice(1, 101), slice(1, 117))   # interior: ghost cells of uvel hold the cyclic wrap
    assert np.all(st["uvel"][I][f["iceumask"][I] == 0] == 0.0)
    assert np.all(st["stressp_1"][f["icetmask"] == 0] == 0.0)


def test_oracle_fast_build_agrees_with_strict(oracle):
    """The -O3/-march/OpenMP timing build may contract FMAs; it must stay within the parity
    tolerance of the strict build (it is only ever used as the timed CPU baseline)."""
    case = synth.make_case("gx3", nx=40, ny=36, ew="cyclic", ns="open")
    st1 = synth.zero_state(42, 38)
    st2 = synth.zero_state(42, 38)
    p1 = oracle.make_params(ndte=120, kind="strict")
    p2 = oracle.make_params(ndte=120, kind="fast")
    oracle.run_evp(case.grid, case.inputs, st1, p1, lib_kind="strict")
    oracle.run_evp(case.grid, case.inputs, st2, p2, lib_kind="fast")
    assert np.max(np.abs(st1["uvel"] - st2["uvel"])) <= 1e-10
    assert np.max(np.abs(st1["vvel"] - st2["vvel"])) <= 1e-10

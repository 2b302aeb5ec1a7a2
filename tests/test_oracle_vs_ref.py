"""CPU tests (build container): the hand-written oracle against THE REFERENCE ITSELF.

oracle/_ref/libevp_ref_<variant>.so is the reference's own Fortran source of `evp` and everything it
calls, translated statement by statement to C by oracle/f90_to_c.py at build time (the image has no
Fortran compiler) and compiled with gcc -O2 -ffp-contract=off -- see oracle/build_ref.py and
oracle/ref_glue.c.  It exists only where /root/reference does (this container, built by
__graft_entry__.build()); elsewhere these tests skip and the committed outputs of the same library
(tests/golden/ref_evp_*.npz, checked by tests/test_oracle_golden.py) stand in.

Every comparison is bit-exact: state after consecutive calls and every output field.
"""
import numpy as np
import pytest

from cice4_b200 import synth
from conftest import GX3_FIXTURE
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")

DOMAINS = [
    ("gx3-real-grid", dict(name="gx3", realistic=True, gx3_fixture=GX3_FIXTURE)),
    ("gx3-dense", dict(name="gx3", realistic=False, gx3_fixture=GX3_FIXTURE)),
    ("tripole-64x48-realistic", dict(name="om1deg", nx=64, ny=48, realistic=True)),
    ("tripole-91x37", dict(name="om1deg", nx=91, ny=37)),
    ("cyclic-cyclic-40x33", dict(name="x", nx=40, ny=33, ew="cyclic", ns="cyclic")),
    ("open-open-37x29", dict(name="x", nx=37, ny=29, ew="open", ns="open")),
    ("closed-closed-30x31", dict(name="x", nx=30, ny=31, ew="closed", ns="closed")),
    ("tripoleT-64x48-realistic", dict(name="x", nx=64, ny=48, ew="cyclic", ns="tripoleT", realistic=True)),
    ("tripoleT-45x31", dict(name="x", nx=45, ny=31, ew="cyclic", ns="tripoleT")),
]
TURN = dict(cosw=0.9063077870366499, sinw=0.42261826174069944)
VARIANTS = [
    ("cice4", dict()),
    ("coupled", dict(coupled=1)),
    ("auscom-geostrophic-tilt", dict(auscom=1, coupled=1, use_ocnslope=0, **TURN)),
    ("auscom-ocnslope", dict(auscom=1, coupled=1, use_ocnslope=1, **TURN)),
    ("access", dict(auscom=1, coupled=1, use_ocnslope=1, access_wind=1)),
]


def _inputs(case, over, seed=11):
    inp = dict(case.inputs)
    g = case.grid
    rng = np.random.default_rng(seed)
    inp["strax"] = np.asfortranarray(0.9 * inp["strairxT"] + 0.001)
    inp["stray"] = np.asfortranarray(1.1 * inp["strairyT"] - 0.002)
    inp["ss_tltx"] = np.asfortranarray(1e-6 * rng.standard_normal((g.nx_block, g.ny_block)))
    inp["ss_tlty"] = np.asfortranarray(1e-6 * rng.standard_normal((g.nx_block, g.ny_block)))
    return inp


def _both(case, over, dt=3600.0, ndte=120, nsteps=2):
    g = case.grid
    inp = _inputs(case, over)
    p = O.make_params(dt=dt, ndte=ndte, **over)
    st_o = synth.zero_state(g.nx_block, g.ny_block)
    st_r = synth.zero_state(g.nx_block, g.ny_block)
    for _ in range(nsteps):
        f_o, _sec = O.run_evp(g, inp, st_o, p)
        f_r = O.run_evp_ref(g, inp, st_r, p, dt)
    bad = [n for n in O.STATE_D + ["iceumask"] if not np.array_equal(st_o[n], st_r[n])]
    outs = [n for n in O.OUT_D if n != "sicemass" or over.get("auscom")]
    bad += [n for n in outs if not np.array_equal(f_o[n], f_r[n])]
    return bad, st_r


@pytest.mark.parametrize("dom", DOMAINS, ids=[d[0] for d in DOMAINS])
@pytest.mark.parametrize("var", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_oracle_bit_exact_vs_reference(dom, var):
    """cold start + warm second call, ndte = 120, every domain type x every CPP variant"""
    case = synth.make_case(**dom[1])
    bad, st = _both(case, var[1])
    assert not bad, f"oracle differs from the reference in {bad}"
    assert np.abs(st["uvel"]).max() > 1e-2


@pytest.mark.parametrize("over", [
    dict(evp_damping=1), dict(kstrength=0), dict(krdg_partic=0), dict(krdg_redist=0),
    dict(krdg_partic=0, krdg_redist=0), dict(mu_rdg=3.0), dict(evp_damping=1, auscom=1, coupled=1, **TURN),
], ids=["evp_damping", "hibler79", "partic0", "redist0", "partic0-redist0", "mu_rdg3", "damping-auscom"])
def test_namelist_options_bit_exact_vs_reference(over):
    case = synth.make_case("om1deg", nx=48, ny=40, realistic=True)
    bad, _ = _both(case, over)
    assert not bad, f"oracle differs from the reference in {bad}"


@pytest.mark.parametrize("dt,ndte", [(3600.0, 120), (1800.0, 120), (600.0, 240), (3600.0, 7), (900.0, 1)])
def test_time_step_and_ndte_bit_exact_vs_reference(dt, ndte):
    case = synth.make_case("x", nx=33, ny=26, ew="cyclic", ns="open")
    bad, _ = _both(case, dict(), dt=dt, ndte=ndte, nsteps=3)
    assert not bad, f"oracle differs from the reference in {bad}"


@pytest.mark.parametrize("bx,by", [(10, 8), (7, 11), (14, 5), (28, 22), (13, 22)])
@pytest.mark.parametrize("var", [VARIANTS[0], VARIANTS[4]], ids=["cice4", "access"])
def test_reference_on_block_decompositions(bx, by, var):
    """The translated reference itself run on create_blocks decompositions (padded edge blocks included;
    several blocks exercise the block loops of `evp`, the per-block index lists and get_block): on the
    cells evp defines in each block it must equal the single-block oracle -- the reference's own
    decomposition invariance (doc/cicedoc.pdf 4.6), and the cell sets the CUDA marshalling reproduces."""
    from cice4_b200 import evp as E
    from helpers import BLOCK_REGION, block_region_mismatches
    case = synth.make_case("om1deg", nx=28, ny=22, realistic=True)
    g = case.grid
    over = var[1]
    inp = _inputs(case, over)
    p = O.make_params(dt=3600.0, ndte=120, **over)
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    st = synth.zero_state(g.nx_block, g.ny_block)
    for _ in range(2):
        f, _sec = O.run_evp(g, inp, st, p)
    lay = E.BlockLayout.cartesian(g.nx, g.ny, bx, by)
    gfb = {k: np.asfortranarray(v) for k, v in E.grid_fields_in_blocks(g, lay, ew, ns).items() if v is not None}
    inb = {k: np.asfortranarray(E.split_blocks(v, lay, ew, ns)) for k, v in inp.items()}
    stb = {k: np.zeros(lay.shape, dtype=v.dtype, order="F") for k, v in st.items()}
    for _ in range(2):
        fb = O.run_evp_ref_blocks(lay, g.ew, g.ns, gfb, inb, stb, p, 3600.0)
    bad = {}
    for n in BLOCK_REGION:
        if n == "sicemass" and not over.get("auscom"):
            continue
        got = stb[n] if n in stb else fb[n]
        want = E.split_blocks(st[n] if n in st else f[n], lay, ew, ns)
        b = block_region_mismatches(n, got, want, lay)
        if b:
            bad[n] = b
    assert not bad, f"blocks that differ from the single-block result: {bad}"


def test_reference_with_land_blocks_eliminated():
    """Land-block elimination (source/ice_distribution.F90): the translated reference run WITHOUT the
    blocks that hold no ocean cell (their neighbours' ghost cells are zero-filled by the halo update)
    gives the single-block result on every remaining block -- the property the CUDA path relies on when
    it treats cells that no block covers as land."""
    from cice4_b200 import evp as E
    from helpers import BLOCK_REGION, block_region_mismatches
    case = synth.make_case("om1deg", nx=96, ny=64, realistic=True)
    g = case.grid
    p = O.make_params(dt=3600.0, ndte=120)
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    st = synth.zero_state(g.nx_block, g.ny_block)
    for _ in range(2):
        f, _sec = O.run_evp(g, case.inputs, st, p)
    full = E.BlockLayout.cartesian(g.nx, g.ny, 8, 8)
    land = full.land_blocks(g.f["tmask"])
    assert len(land) >= 5
    lay = full.without(land)
    gfb = {k: np.asfortranarray(v) for k, v in E.grid_fields_in_blocks(g, lay, ew, ns).items() if v is not None}
    inb = {k: np.asfortranarray(E.split_blocks(v, lay, ew, ns)) for k, v in case.inputs.items()}
    stb = {k: np.zeros(lay.shape, dtype=v.dtype, order="F") for k, v in st.items()}
    for _ in range(2):
        fb = O.run_evp_ref_blocks(lay, g.ew, g.ns, gfb, inb, stb, p, 3600.0)
    bad = {}
    for n in BLOCK_REGION:
        if n == "sicemass":
            continue
        got = stb[n] if n in stb else fb[n]
        b = block_region_mismatches(n, got, E.split_blocks(st[n] if n in st else f[n], lay, ew, ns), lay)
        if b:
            bad[n] = b
    assert not bad, bad


@pytest.mark.parametrize("dt,ndte", [(3600.0, 120), (1800.0, 120), (600.0, 240), (7200.0, 77)])
def test_set_evp_parameters_vs_reference(dt, ndte):
    """source/ice_dyn_evp.F90:535-577 as the reference computes it"""
    import ctypes as C
    p = O.make_params(dt=dt, ndte=ndte)
    out = (C.c_double * 6)()
    O.ref_lib("cice4").ref_set_evp_parameters(C.byref(p), dt, out)
    assert list(out) == [p.dtei, p.ecci, p.dte2T, p.denom1, p.denom2, p.rcon]


def test_principal_stress_vs_reference():
    """source/ice_dyn_evp.F90:1558-1609"""
    import ctypes as C
    case = synth.make_case("om1deg", nx=48, ny=40, realistic=True)
    g = case.grid
    p = O.make_params()
    st = synth.zero_state(g.nx_block, g.ny_block)
    f, _ = O.run_evp(g, case.inputs, st, p)
    s1, s2 = O.principal_stress(st["stressp_1"], st["stressm_1"], st["stress12_1"], f["prs_sig"])
    r1, r2 = np.zeros_like(s1, order="F"), np.zeros_like(s1, order="F")
    O.ref_lib("cice4").ref_principal_stress(C.byref(p), g.nx_block, g.ny_block, O._ptr(st["stressp_1"]),
                                            O._ptr(st["stressm_1"]), O._ptr(st["stress12_1"]),
                                            O._ptr(f["prs_sig"]), O._ptr(r1), O._ptr(r2))
    assert np.array_equal(s1, r1) and np.array_equal(s2, r2)
    assert (r1 < 1e29).any() and (r1 == 1e30).any()


def test_reference_fma_build_within_tolerance():
    """The reference's own code under a contracting compiler (gcc -O3 -march=x86-64-v3, the stand-in for
    the production `ifort -O3 -xHost`, bld/Macros.nci:26) against its strict build: the difference is the
    freedom a Fortran compiler has, and it stays inside the north_star tolerance (1e-10 m/s, 1e-10
    relative stress) that the FMA mode of the CUDA path is held to."""
    import ctypes as C
    import os
    if not os.path.exists(os.path.join(O.REF_DIR, "libevp_ref_cice4_fast.so")):
        pytest.skip("fast build of the translated reference not present")
    case = synth.make_case("om1deg", nx=120, ny=100)
    g = case.grid
    p = O.make_params(dt=3600.0, ndte=120)
    res = {}
    for variant in ("cice4", "cice4_fast"):
        st = synth.zero_state(g.nx_block, g.ny_block)
        gg = O.make_grid(g.nx_block, g.ny_block, g.ew, g.ns)
        for _ in range(2):
            f = O.Fields(g.f, case.inputs, st, None)
            assert O.ref_lib(variant).ref_evp(C.byref(gg), C.byref(p), C.byref(f.c), 3600.0) == 0
        res[variant] = st
    a, b = res["cice4"], res["cice4_fast"]
    assert max(np.abs(a[n] - b[n]).max() for n in ("uvel", "vvel")) <= 1e-10
    for n in O.STATE_D[2:]:
        assert np.abs(a[n] - b[n]).max() <= 1e-10 * np.abs(a[n]).max(), n
    assert any(not np.array_equal(a[n], b[n]) for n in O.STATE_D)   # the builds do differ


def test_reference_timing_build_is_thread_invariant():
    """The timing build of the translated reference runs the cell loops of stress / stepu on OpenMP threads
    (independent iterations over the index lists): one thread and many give the same bits."""
    import ctypes as C
    import os
    if not os.path.exists(os.path.join(O.REF_DIR, "libevp_ref_cice4_fast.so")):
        pytest.skip("fast build of the translated reference not present")
    case = synth.make_case("x", nx=70, ny=55, ew="cyclic", ns="open")
    g = case.grid
    p = O.make_params(dt=3600.0, ndte=40)
    res = []
    for threads in (1, 4):
        O.omp_set_num_threads(threads)
        st = synth.zero_state(g.nx_block, g.ny_block)
        gg = O.make_grid(g.nx_block, g.ny_block, g.ew, g.ns)
        f = O.Fields(g.f, case.inputs, st, None)
        assert O.ref_lib("cice4_fast").ref_evp(C.byref(gg), C.byref(p), C.byref(f.c), 3600.0) == 0
        res.append((st, f))
    for n in O.STATE_D:
        assert np.array_equal(res[0][0][n], res[1][0][n]), n
    for n in ("divu", "strintx", "strocnxT", "prs_sig"):
        assert np.array_equal(res[0][1][n], res[1][1][n]), n


# ------------------------------------------------------------------------------------------------
# the reference's OWN halo machinery (translated create_blocks loop, ice_blocksGetNbrID, message loop of
# ice_HaloCreate, ice_HaloMsgCreate, ice_HaloUpdate2DR8 / 2DI4): what pins oracle/evp_oracle.c's halo
# ------------------------------------------------------------------------------------------------
BND_ALL = {"open": 0, "closed": 1, "cyclic": 2, "tripole": 3, "tripoleT": 4}


@pytest.mark.parametrize("ns", ["open", "closed", "cyclic", "tripole", "tripoleT"])
@pytest.mark.parametrize("ew", ["open", "closed", "cyclic"])
def test_oracle_halo_equals_reference_halo(ew, ns):
    """Every boundary combination x field location x field kind x three sizes (even, odd, tiny), real and integer
    fields: the oracle's halo update, its independent numpy restatement and the reference's own translated halo
    update give the same array bit for bit -- ghost ring, fill cells and, on the tripole, the top physical row."""
    from cice4_b200 import grid as G
    rng = np.random.default_rng(7)
    for nx, ny in ((12, 9), (13, 7), (8, 6)):
        for loc in (1, 2, 3, 4):
            for kind in (1, 2, 3):
                a = np.asfortranarray(rng.standard_normal((nx + 2, ny + 2)))
                b, c, d = a.copy(order="F"), a.copy(order="F"), a.copy(order="F")
                O.halo_r8(b, BND_ALL[ew], BND_ALL[ns], loc, kind)
                O.ref_halo(c, BND_ALL[ew], BND_ALL[ns], loc, kind)
                G.halo_update(d, BND_ALL[ew], BND_ALL[ns], loc, kind)
                assert np.array_equal(b, c), (nx, ny, loc, kind, np.argwhere(b != c)[:5].tolist())
                assert np.array_equal(d, c), (nx, ny, loc, kind, np.argwhere(d != c)[:5].tolist())
        m = np.asfortranarray(rng.integers(0, 2, (nx + 2, ny + 2)).astype(np.int32))   # icetmask: centre scalar
        mb, mc = m.copy(order="F"), m.copy(order="F")
        import ctypes as C
        g = O.make_grid(nx + 2, ny + 2, BND_ALL[ew], BND_ALL[ns])
        O.lib().orc_halo_i4(mb.ctypes.data_as(O.c_ip), C.byref(g), 1, 1, 0)
        O.ref_halo(mc, BND_ALL[ew], BND_ALL[ns], 1, 1)
        assert np.array_equal(mb, mc)


@pytest.mark.parametrize("ew,ns", [("cyclic", "tripole"), ("cyclic", "tripoleT"), ("cyclic", "open"), ("open", "closed"),
                                   ("cyclic", "cyclic")])
@pytest.mark.parametrize("bx,by", [(10, 8), (7, 11), (14, 5), (13, 22)])
def test_reference_multiblock_halo_equals_single_block(ew, ns, bx, by):
    """The reference's halo update on create_blocks decompositions (its own address lists, several blocks, padded
    edge blocks, corner neighbours, tripole buffer): every block ends up with the ring -- and on the tripole the
    top physical row -- that the one-block update gives the same field."""
    from cice4_b200 import evp as E
    nx, ny = 28, 22
    rng = np.random.default_rng(3)
    lay = E.BlockLayout.cartesian(nx, ny, bx, by)
    if ns == "tripoleT" and by == 5:      # top blocks of 2 rows < tripoleRows = 3: the reference itself stops
        with pytest.raises(RuntimeError, match="not enough points in block for tripole"):
            O.ref_halo(np.zeros(lay.shape, order="F"), BND_ALL[ew], BND_ALL[ns], 1, 1, layout=lay)
        return
    for loc, kind in ((1, 1), (2, 2), (1, 2)):
        glob = np.zeros((nx + 2, ny + 2), order="F")     # ghost cells no update defines (open / closed) stay 0 in both
        glob[1:nx + 1, 1:ny + 1] = rng.standard_normal((nx, ny))
        one = glob.copy(order="F")
        O.ref_halo(one, BND_ALL[ew], BND_ALL[ns], loc, kind)
        blk = np.zeros(lay.shape, order="F")
        for b in range(lay.nblocks):   # physical cells only: the update must produce every ghost cell itself
            i0, j0 = lay.iglob_lo[b], lay.jglob_lo[b]
            for j in range(lay.jlo[b], lay.jhi[b] + 1):
                for i in range(lay.ilo[b], lay.ihi[b] + 1):
                    blk[i - 1, j - 1, b] = glob[i0 + i - lay.ilo[b], j0 + j - lay.jlo[b]]
        O.ref_halo(blk, BND_ALL[ew], BND_ALL[ns], loc, kind, layout=lay)
        for b in range(lay.nblocks):
            i0, j0 = lay.iglob_lo[b], lay.jglob_lo[b]
            for j in range(lay.jlo[b] - 1, lay.jhi[b] + 2):
                for i in range(lay.ilo[b] - 1, lay.ihi[b] + 2):
                    want = one[i0 + i - lay.ilo[b], j0 + j - lay.jlo[b]]
                    assert blk[i - 1, j - 1, b] == want, (loc, kind, b, i, j)


def test_reference_create_blocks_one_column_edge_block():
    """A quirk of the reference that the translated create_blocks exposes: a padded edge block of exactly ONE physical
    column keeps its full width, because ihi is only shrunk for i > ilo (source/ice_blocks.F90:332-335).  The host
    refuses a caller layout that differs from what create_blocks makes instead of silently using either."""
    from cice4_b200 import evp as E
    lay = E.BlockLayout.cartesian(28, 22, 9, 22)      # 28 = 3 x 9 + 1
    assert lay.ihi[-1] == lay.ilo[-1]
    a = np.zeros(lay.shape, order="F")
    with pytest.raises(RuntimeError, match="not what create_blocks makes"):
        O.ref_halo(a, 2, 3, 2, 2, layout=lay)


@pytest.mark.parametrize("name", ["ice_HaloUpdate2DR8", "ice_HaloUpdate2DI4"])
def test_mpi_halo_update_is_the_serial_one_plus_message_plumbing(name):
    """oracle/_ref executes the SERIAL build's ice_HaloUpdate (serial/ice_boundary.F90:591-873, 1169-1451); the MPI
    build's (mpi/ice_boundary.F90:1028-1417, ...) is what a multi-task run of the reference executes.  Text-level pin:
    every statement of the serial routine appears, in order, in the MPI routine, and what the MPI routine adds is only
    message plumbing -- request / status arrays, posting the receives, packing + sending, waiting, and unpacking the
    received values into `array` or into the same tripole buffer the local copies fill.  So the local copies, the
    tripole buffer arithmetic and the copy-out that the translated reference runs are the MPI build's too."""
    import difflib
    import os
    import re
    ref = "/root/reference"

    def routine(path):
        src = open(os.path.join(ref, path)).read().splitlines()
        a = [i for i, l in enumerate(src) if re.match(r"\s*subroutine\s+" + name + r"\b", l)][0]
        b = [i for i, l in enumerate(src) if re.match(r"\s*end subroutine\s+" + name + r"\b", l)][0]
        out = []
        for l in src[a:b + 1]:
            l = re.sub(r"\s+", " ", l.split("!")[0].strip().lower())
            if l:
                out.append(l)
        return out

    s, m = routine("serial/ice_boundary.F90"), routine("mpi/ice_boundary.F90")
    ops = difflib.SequenceMatcher(None, s, m, autojunk=False).get_opcodes()
    assert {t for t, *_ in ops} == {"equal", "insert"}, [o for o in ops if o[0] not in ("equal", "insert")]
    kind = "r8" if name.endswith("R8") else "i4"
    plumbing = [r"^integer \(int_kind\)(, dimension\(:(,:)?\), allocatable)? ::", r"^(snd|rcv)(request|status)(, &)?$",
                r"^(allocate|deallocate)\(", r"^(snd|rcv)(request|status)\(", r"^if \(ierr > 0\) then$", r"^call abort_ice\(",
                r"^'ice_haloupdate2d" + kind + r": error (de)?allocating req,status arrays'\)$", r"^return$", r"^end ?if$",
                r"^do nmsg=1,halo%nummsg(send|recv)$", r"^do n=", r"^end do$", r"^len = halo%size(send|recv)\(nmsg\)$",
                r"^call mpi_(irecv|isend|waitall)\(", r"^halo%(recv|send)task\(nmsg\), &$", r"^mpitaghalo \+ ",
                r"^halo%communicator, (snd|rcv)request\(nmsg\), ierr\)$",
                r"^(isrc|jsrc|srcblock) = halo%sendaddr\(\d,n,nmsg\)$", r"^(idst|jdst|dstblock) = halo%recvaddr\(\d,n,nmsg\)$",
                r"^bufsend" + kind + r"\(n,nmsg\) = (array\(isrc,jsrc,srcblock\)|fill)$",
                r"^if \(dstblock > 0\) then$", r"^else if \(dstblock < 0\) then$",
                r"^array\(idst,jdst,dstblock\) = bufrecv" + kind + r"\(n,nmsg\)$",
                r"^buftripole" + kind + r"\(idst,jdst\) = bufrecv" + kind + r"\(n,nmsg\)$"]
    added = [l for t, i1, i2, j1, j2 in ops if t == "insert" for l in m[j1:j2]]
    assert len(added) > 40
    for l in added:
        assert any(re.match(p, l) for p in plumbing), f"MPI-only statement that is not message plumbing: {l}"

"""CPU tests of the drop-in boundary: libevp_b200.so builds (nvcc cross-compiles), loads and
exports every symbol include/evp_b200.h declares; without a GPU every compute entry fails
loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "evp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(evp_b200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(evp_lib):
    from cice4_b200 import evp as E
    names = _declared()
    assert len(names) >= 12
    assert sorted(E.EXPORTS) == names, "cice4_b200.evp.EXPORTS must list exactly the header's entry points"
    for n in names:
        assert hasattr(evp_lib, n), f"{n} declared in include/evp_b200.h but not exported"


def test_abi_version_and_defaults(evp_lib):
    from cice4_b200 import evp as E
    assert evp_lib.evp_b200_abi_version() == 1
    p = E.default_params()
    # reference defaults: source/ice_init.F90:216-222, drivers/cice4/ice_constants.F90:50-60
    assert p.ndte == 120 and p.evp_damping == 0 and p.kstrength == 1
    assert p.krdg_partic == 1 and p.krdg_redist == 1 and p.mu_rdg == 3.0 and p.ncat == 5
    assert (p.rhoi, p.rhos, p.rhow, p.dragio) == (917.0, 330.0, 1026.0, 0.00536)
    assert (p.cosw, p.sinw, p.puny) == (1.0, 0.0, 1.0e-11)


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (guards against drift between header and mirror)."""
    from cice4_b200 import evp as E
    assert C.sizeof(E.Dims) == 8 * 4 + 6 * 8 + 5 * 4 + 4   # trailing pad to 8
    assert C.sizeof(E.StaticFields) == 18 * 8
    assert C.sizeof(E.Inputs) == 12 * 8
    assert C.sizeof(E.State) == 15 * 8
    assert C.sizeof(E.Outputs) == 20 * 8
    assert C.sizeof(E.Timings) == 10 * 4


def test_bad_arguments_are_rejected(evp_lib):
    from cice4_b200 import evp as E
    h = C.c_void_p(None)
    rc = evp_lib.evp_b200_init(None, None, None, C.byref(h))
    assert rc == 1 and b"NULL" in evp_lib.evp_b200_last_error()
    lay = E.BlockLayout.single_block(8, 8)
    dyn = E.IceDynEvp(lay, "cyclic", "open")
    with pytest.raises(E.EvpB200Error):
        dyn.evp(3600.0, {})          # init_evp not called


def test_no_cpu_fallback_without_gpu(evp_lib):
    """On a box without a CUDA device init must fail with EVP_B200_ERR_CUDA, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    from cice4_b200 import evp as E, synth
    case = synth.make_case("gx3", nx=12, ny=10, ew="cyclic", ns="open")
    lay = E.BlockLayout.single_block(12, 10)
    dyn = E.IceDynEvp(lay, "cyclic", "open")
    gf = E.grid_fields_in_blocks(case.grid, lay, "cyclic", "open")
    with pytest.raises(E.EvpB200Error, match="error 2"):
        dyn.init_evp(3600.0, gf)


def test_block_layout_split_merge_roundtrip():
    from cice4_b200 import evp as E
    rng = np.random.default_rng(0)
    nx, ny = 23, 17
    a = np.asfortranarray(rng.standard_normal((nx + 2, ny + 2)))
    for bx, by in ((23, 17), (8, 6), (5, 17), (23, 4)):
        lay = E.BlockLayout.cartesian(nx, ny, bx, by)
        blk = E.split_blocks(a, lay, "cyclic", "open")
        assert blk.shape == (bx + 2, by + 2, lay.nblocks)
        back = E.merge_blocks(blk, lay)
        np.testing.assert_array_equal(back, a)

#!/usr/bin/env python3
"""build_ref.py -- TEST INFRASTRUCTURE ONLY.

Builds `oracle/_ref/libevp_ref_<variant>.so`: the reference's own EVP path, machine-translated from
the Fortran under /root/reference (oracle/f90_to_c.py) and hosted by oracle/ref_glue.c.  One library
per CPP variant the reference is built with:

    cice4   no defines                     (drivers/cice4)
    auscom  -DAusCOM -Dcoupled             (drivers/access-om, bld/Macros.nci:56-57)
    access  -DAusCOM -Dcoupled -DACCESS    (drivers/access-cm)
    coupled -Dcoupled                      (slope tilt without the AusCOM changes)

Needs /root/reference (this container only).  The generated C lives in a temporary directory for the
duration of the build (EVP_REF_KEEP_C=1 keeps it for inspection); only the libraries are written to
oracle/_ref/ (git-ignored): no reference source, translated or not, enters the repository tree.
Run:  python oracle/build_ref.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
if __package__:
    from . import f90_to_c as T
else:  # run as a script: import the sibling module without putting oracle/ itself on sys.path
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import f90_to_c as T  # noqa: E402

REF = os.environ.get("CICE4_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# variant -> (cpp defines, driver directory whose ice_constants.F90 the build uses)
VARIANTS = {
    "cice4": ((), "cice4"),
    "coupled": (("coupled",), "cice4"),
    "auscom": (("AusCOM", "coupled"), "access-om"),
    "access": (("AusCOM", "coupled", "ACCESS"), "access-cm"),
}

BLOCK3 = ["nx_block", "ny_block", "max_blocks"]


def available():
    return os.path.isfile(os.path.join(REF, "source", "ice_dyn_evp.F90"))


def translate(defines, driver, omp=False):
    tr = T.Translator(defines)
    src = os.path.join(REF, "source")
    # modules whose variables the path uses -- the few that are not plain declarations are listed
    # by hand: ice_blocks / ice_domain_size make these compile-time parameters of the executable
    for n in ("nx_block", "ny_block", "max_blocks", "ncat", "nblocks", "halo_info", "timer_dynamics",
              "timer_bound"):
        tr.add_global(T.Decl(n, "integer"))
    tr.add_global(T.Decl("blocks_ice", "integer", ["max_blocks"]))                 # source/ice_domain.F90
    tr.add_global(T.Decl("work1", "real", BLOCK3))                                 # source/ice_work.F90
    if "AusCOM" in defines:
        tr.add_global(T.Decl("use_ocnslope", "logical"))                           # drivers/access-om/cpl_parameters.F90:40
        tr.add_global(T.Decl("sicemass", "real", BLOCK3))                          # drivers/access-om/cpl_interface.F90:418
    tr.module_decls(os.path.join(REF, "drivers", driver, "ice_constants.F90"))
    tr.module_decls(os.path.join(src, "ice_dyn_evp.F90"), dims_override={"fcor_blk": BLOCK3})
    tr.module_decls(os.path.join(src, "ice_mechred.F90"))
    tr.module_decls(os.path.join(src, "ice_state.F90"),
                    only={"aice", "vice", "vsno", "aice0", "aicen", "vicen", "uvel", "vvel", "strength", "divu",
                          "shear"})
    tr.module_decls(os.path.join(src, "ice_flux.F90"),
                    only={"strairxt", "strairyt", "strax", "stray", "uocn", "vocn", "ss_tltx", "ss_tlty",
                          "stressp_1", "stressp_2", "stressp_3", "stressp_4", "stressm_1", "stressm_2",
                          "stressm_3", "stressm_4", "stress12_1", "stress12_2", "stress12_3", "stress12_4",
                          "iceumask", "strairx", "strairy", "strtltx", "strtlty", "strintx", "strinty",
                          "strocnx", "strocny", "strocnxt", "strocnyt", "fm", "prs_sig", "rdg_conv", "rdg_shear"})
    tr.module_decls(os.path.join(src, "ice_grid.F90"),
                    only={"dxt", "dyt", "dxhy", "dyhx", "cxp", "cyp", "cxm", "cym", "tarea", "tarear", "tinyarea",
                          "uarea", "uarear", "tmask", "umask"})
    missing = set()
    for path, subs in (
        (os.path.join(src, "ice_mechred.F90"), ["asum_ridging", "ridge_itd", "ice_strength"]),
        (os.path.join(src, "ice_grid.F90"), ["to_ugrid", "to_tgrid", "t2ugrid_vector", "u2tgrid_vector"]),
        (os.path.join(src, "ice_dyn_evp.F90"), ["set_evp_parameters", "evp_prep1", "evp_prep2", "stress", "stepu",
                                                "evp_finish", "principal_stress", "evp"]),
    ):
        for s in subs:
            # timing build: the cell loops of stress / stepu (independent iterations over the index lists,
            # as the OpenMP directives of later CICE versions assume) run on all host threads
            missing.update(tr.subroutine(path, s, omp={"ij"} if omp and s in ("stress", "stepu") else None))
    missing.update(translate_halo(tr))
    missing -= set(tr.subs) | {"get_block", "ice_haloupdate", "ice_timer_start", "ice_timer_stop",
                               "get_block_parameter", "abort_ice"}
    if missing:
        raise T.TranslateError("unresolved names: " + ", ".join(sorted(missing)))
    return tr.emit_file()


def translate_halo(tr):
    """The reference's own block decomposition and halo machinery (serial build): the block loop of `create_blocks`,
    `ice_blocksGetNbrID` (source/ice_blocks.F90), `ice_distributionGetBlockLoc` (source/ice_distribution.F90), the
    message-configuration loop of `ice_HaloCreate`, `ice_HaloMsgCreate`, `ice_HaloUpdate2DR8` and `ice_HaloUpdate2DI4`
    (serial/ice_boundary.F90), with the derived types `block`, `distrb` and `ice_halo`.  Allocation (the head of
    create_blocks / ice_HaloCreate, whose sizes come from a counting pass over the same loop) and
    `get_block_parameter` (optional pointer results) are hosted by oracle/ref_glue.c."""
    src = os.path.join(REF, "source")
    blk = os.path.join(src, "ice_blocks.F90")
    dst = os.path.join(src, "ice_distribution.F90")
    bnd = os.path.join(REF, "serial", "ice_boundary.F90")
    for n in ("my_task", "nblocks_tot", "nblocks_x", "nblocks_y", "block_size_x", "block_size_y", "buf_nx", "buf_rows"):
        tr.add_global(T.Decl(n, "integer"))
    tr.module_decls(blk, only={"nghost", "ice_blocksnorth", "ice_blockssouth", "ice_blockseast", "ice_blockswest",
                               "ice_blocksnortheast", "ice_blocksnorthwest", "ice_blockssoutheast",
                               "ice_blockssouthwest", "ice_blockseast2", "ice_blockswest2",
                               "ice_blockseastnortheast", "ice_blockswestnorthwest"})
    tr.type_def(blk, "block", {"i_glob": [None], "j_glob": [None]})
    tr.type_def(dst, "distrb", {"blocklocation": [None], "blocklocalid": [None], "blockglobalid": [None]})
    tr.type_def(bnd, "ice_halo", {"srclocaladdr": ["3", None], "dstlocaladdr": ["3", None]})
    tr.add_global(T.Decl("all_blocks", "type", ["nblocks_tot"], tname="block", pointer=True))
    tr.add_global(T.Decl("all_blocks_ij", "integer", ["nblocks_x", "nblocks_y"], pointer=True))
    tr.add_global(T.Decl("i_global", "integer", ["nx_block", "nblocks_tot"], pointer=True))
    tr.add_global(T.Decl("j_global", "integer", ["ny_block", "nblocks_tot"], pointer=True))
    tr.add_global(T.Decl("buftripoler8", "real", ["buf_nx", "buf_rows"], pointer=True))
    tr.add_global(T.Decl("buftripolei4", "integer", ["buf_nx", "buf_rows"], pointer=True))
    missing = set()
    missing.update(tr.translate_range(blk, "create_blocks", "subroutine", r"^do\s+jblock\s*=", None,
                                      "create_blocks_loop",
                                      ["nx_global", "ny_global", "ew_boundary_type", "ns_boundary_type"]))
    missing.update(tr.function(blk, "ice_blocksgetnbrid"))
    missing.update(tr.subroutine(dst, "ice_distributiongetblockloc"))
    missing.update(tr.subroutine(bnd, "ice_halomsgcreate"))
    missing.update(tr.translate_range(bnd, "ice_halocreate", "function", r"^msgconfigloop\s*:", r"^end\s*do\s+msgconfigloop$",
                                      "ice_halocreate_msgconfig", ["halo", "dist", "nsboundarytype", "ewboundarytype"]))
    missing.update(tr.subroutine(bnd, "ice_haloupdate2dr8", dims_override={"array": BLOCK3}))
    missing.update(tr.subroutine(bnd, "ice_haloupdate2di4", dims_override={"array": BLOCK3}))
    return missing


def build(verbose=False):
    if not available():
        raise RuntimeError("reference sources not found under " + REF)
    os.makedirs(OUT, exist_ok=True)
    for stale in os.listdir(OUT):      # generated C of earlier builds: only libraries stay in oracle/_ref
        if stale.endswith(".c"):
            os.remove(os.path.join(OUT, stale))
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    tmp = tempfile.mkdtemp(prefix="evp_ref_")   # the translated text never enters the repository tree
    for name, (defines, driver) in VARIANTS.items():
        gen = os.path.join(tmp, "evp_ref_%s.c" % name)
        with open(gen, "w") as fh:
            fh.write(translate(defines, driver))
        # strict: the parity authority (no FMA contraction).  fast (cice4 only): the reference's
        # production optimisation level (bld/Macros.nci:26 is -O3 -xHost) with the cell loops of stress
        # and stepu on all host threads, timed by bench.py
        builds = [("", ["-O2", "-ffp-contract=off", "-fno-fast-math"], gen)]
        if name == "cice4":
            gen_omp = os.path.join(tmp, "evp_ref_%s_omp.c" % name)
            with open(gen_omp, "w") as fh:
                fh.write(translate(defines, driver, omp=True))
            builds.append(("_fast", ["-O3", "-march=x86-64-v3", "-fopenmp"], gen_omp))
        for suffix, opt, src in builds:
            cmd = [cc, "-std=gnu11"] + opt + ["-fPIC", "-shared", "-Wall", "-Wno-unused", "-Wno-parentheses",
                                              "-Wno-maybe-uninitialized", '-DREF_GEN="%s"' % src, "-I", HERE]
            if "AusCOM" in defines:
                cmd.append("-DREF_AUSCOM")
            cmd += ["-o", os.path.join(OUT, "libevp_ref_%s%s.so" % (name, suffix)),
                    os.path.join(HERE, "ref_glue.c"), os.path.join(HERE, "evp_oracle.c"), "-lm"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if os.environ.get("EVP_REF_KEEP_C"):
        print("generated C kept in", tmp)
    else:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    build(verbose=True)

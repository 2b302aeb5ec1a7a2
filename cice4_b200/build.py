"""In-tree build of libevp_b200.so (nvcc, sm_100a only).

    python -m cice4_b200.build [--force]

The subcycle kernel is compiled twice from one body: evp_subcycle_strict.cu with -fmad=false
(unfused IEEE arithmetic, bit-identical to the unfused CPU oracle) and evp_subcycle_fast.cu with
-fmad=true.  Everything else is compiled with -fmad=false.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libevp_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math"]
UNITS = [
    ("evp_subcycle_strict.cu", ["-fmad=false"]),
    ("evp_subcycle_fast.cu", ["-fmad=true"]),
    ("evp_fused_strict.cu", ["-fmad=false"]),
    ("evp_fused_fast.cu", ["-fmad=true"]),
    ("evp_aux.cu", ["-fmad=false"]),
    ("evp_abi.cu", ["-fmad=false"]),
]
DEPS = ["evp_common.cuh", "evp_aux.cuh", "evp_subcycle_body.cuh", "evp_ieee.cuh", "evp_tiled.cuh", "evp_fused.cuh", os.path.join("..", "..", "include", "evp_b200.h")]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libevp_b200 cannot be built (there is no CPU fallback)")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    nvcc = _nvcc()
    deps = [os.path.join(CSRC, d) for d in DEPS]
    objs, cmds = [], []
    for src, flags in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + deps):
            cmds.append([nvcc] + ARCH + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    if cmds:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(subprocess.check_call, cmds))
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

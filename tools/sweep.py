#!/usr/bin/env python
"""Kernel-configuration sweep on one GPU: device-resident subcycle loop time for a list of
(math_mode, tile_threads, tile_rows, kernel_variant) on one workload.  Development tool."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cice4_b200 import build as B, evp as E, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="om025")
    ap.add_argument("--realistic", action="store_true")
    ap.add_argument("--ndte", type=int, default=120)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--configs", default="1:128:0:0,1:128:0:1,1:128:0:2,1:128:0:3,0:128:0:0")
    args = ap.parse_args()
    B.build()
    fixture = os.path.join(ROOT, "tests", "golden", "gx3_grid.npz") if args.workload == "gx3" else None
    case = synth.make_case(args.workload, nx=args.nx or None, ny=args.ny or None, realistic=args.realistic,
                           gx3_fixture=fixture)
    g = case.grid
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    lay = E.BlockLayout.single_block(g.nx, g.ny)
    gf = E.grid_fields_in_blocks(g, lay, ew, ns)
    inputs = {k: E.split_blocks(v, lay, ew, ns) for k, v in case.inputs.items()}
    ref = None
    print(f"# {args.workload} {g.nx}x{g.ny} ndte={args.ndte} active={case.active_fraction:.3f}")
    print("# math threads rows variant | ms/loop  us/subcycle  Gcell-sub/s  GB/s(384B*cells)  max|du| vs first")
    for cfg in args.configs.split(","):
        mm, nt, rows, var = (int(x) for x in cfg.split(":"))
        dyn = E.IceDynEvp(lay, ew, ns, ndte=args.ndte, math_mode=mm, tile_threads=nt, tile_rows=rows,
                          kernel_variant=var)
        dyn.init_evp(3600.0, gf)
        out = dyn.evp(3600.0, inputs, strength=None, want=["strength"])
        out = dyn.evp(3600.0, inputs, strength=out["strength"], want=["strength"])
        u = dyn.state["uvel"].copy()
        if ref is None:
            ref = u
        dyn.subcycle_resident(2)
        ms = min(dyn.subcycle_resident(args.reps) for _ in range(3))
        cells = g.nx * g.ny
        print(f"{mm:5d} {nt:7d} {rows:4d} {var:7d} | {ms:8.3f} {1e3 * ms / args.ndte:10.2f} "
              f"{cells * args.ndte / ms / 1e6:10.3f} {384.0 * cells * case.active_fraction * args.ndte / ms / 1e6:12.1f} "
              f"{np.abs(u - ref).max():.3e}", flush=True)
        dyn.finalize()


if __name__ == "__main__":
    main()

"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_parity.py [--case om1deg --nx 96 --ny 64 --steps 2]

Every rank owns one y-slab (cice4_b200.slab), runs `evp` through the C ABI with the per-subcycle
row exchange over NCCL, and rank 0 compares the gathered result bit for bit with the single-domain
CPU oracle (doc/cicedoc.pdf 4.6: results do not depend on the decomposition).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    from cice4_b200 import build as B, evp as E, slab, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="om1deg")
    ap.add_argument("--nx", type=int, default=96)
    ap.add_argument("--ny", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--ndte", type=int, default=120)
    ap.add_argument("--realistic", action="store_true")
    ap.add_argument("--math-mode", type=int, default=0)
    ap.add_argument("--use-graph", type=int, default=1)
    ap.add_argument("--exchange-mode", type=int, default=0)
    ap.add_argument("--kernel-variant", type=int, default=0, help="evp_b200_params.kernel_variant (128 = persistent kernel)")
    ap.add_argument("--tile-threads", type=int, default=0, help="evp_b200_params.tile_threads (128 + variant 32768: warp-strip plane kernel)")
    ap.add_argument("--blocks", default="", help="bx,by: reference block layout regrouped into slabs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        B.build()
    dist.barrier()

    case = synth.make_case(args.case, nx=args.nx, ny=args.ny, realistic=args.realistic)
    g = case.grid
    nx, ny = g.nx, g.ny
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    if args.blocks:
        bx, by = (int(x) for x in args.blocks.split(","))
        lay = slab.slab_blocks_from_reference(nx, ny, bx, by, world, rank)
    else:
        lay = slab.slab_layout(nx, ny, world, rank)
    rows = slab.layout_rows(lay)

    dyn = E.IceDynEvp(lay, ew, ns, device=local, rank=rank, nranks=world, slab=rows, ndte=args.ndte,
                      math_mode=args.math_mode, use_graph=args.use_graph, exchange_mode=args.exchange_mode,
                      kernel_variant=args.kernel_variant, tile_threads=args.tile_threads)
    gf = E.grid_fields_in_blocks(g, lay, ew, ns)
    dyn.init_evp(3600.0, gf)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.tensor(list(E.IceDynEvp.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    dyn.comm_init(bytes(uid.cpu().tolist()))

    # oracle on rank 0 (also provides the strength fields so that the comparison is bit-exact)
    strengths = [None] * args.steps
    if rank == 0:
        from helpers import oracle_steps
        from oracle import oracle as O
        st, f, strengths, _ = oracle_steps(O, case, nsteps=args.steps, ndte=args.ndte)
    obj = [strengths]
    dist.broadcast_object_list(obj, 0)
    strengths = obj[0]

    inputs = {k: E.split_blocks(v, lay, ew, ns) for k, v in case.inputs.items()}
    out = None
    for k in range(args.steps):
        out = dyn.evp(3600.0, inputs, strength=E.split_blocks(strengths[k], lay, ew, ns))
    ms = dyn.subcycle_resident(2)

    # gather slabs on rank 0
    names = E.STATE_D + ["iceumask"]
    onames = ["strintx", "strocnxT", "strocnyT", "divu", "prs_sig", "strairx", "fm"]
    mine = {}
    for n in names:
        mine[n] = _local_padded(dyn.state[n], lay, rows, nx)
    for n in onames:
        mine["o_" + n] = _local_padded(out[n], lay, rows, nx)
    gathered = [None] * world
    dist.gather_object((rows, mine), gathered if rank == 0 else None, 0)
    ok = True
    if rank == 0:
        bounds = [gp[0] for gp in gathered]
        for n in names + ["o_" + x for x in onames]:
            full = slab.gather_slabs([gp[1][n] for gp in gathered], bounds, nx, ny)
            ref = st[n] if n in st else f[n[2:]]
            I = (slice(0, nx + 2), slice(0, ny + 2)) if n in ("uvel", "vvel") else (slice(1, nx + 1), slice(1, ny + 1))
            if not np.array_equal(full[I], ref[I]):
                ok = False
                print(f"MISMATCH {n}: max abs diff {np.abs(full[I] - ref[I]).max():.3e}")
        print(f"multigpu_parity: world={world} {args.case} {nx}x{ny} steps={args.steps} blocks='{args.blocks}' "
              f"graph={args.use_graph} variant={args.kernel_variant} info={dyn.info()} launches/loop={dyn.timings()['subcycle_launches']} exchange_mode_used={dyn.timings()['exchange_mode_used']}: "
              f"{'BIT-EXACT vs oracle' if ok else 'FAILED'}; "
              f"resident loop {ms:.3f} ms on rank 0", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dyn.finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def _local_padded(blk, lay, rows, nx):
    """block layout of one slab -> padded (nx+2, rows+2) array (physical cells + outer ring)."""
    from cice4_b200 import evp as E
    jlo, jhi = rows
    n = jhi - jlo + 1
    out = np.zeros((nx + 2, n + 2), dtype=blk.dtype, order="F")
    for b in range(lay.nblocks):
        ni = lay.ihi[b] - lay.ilo[b] + 1
        nj = lay.jhi[b] - lay.jlo[b] + 1
        ig, jg = lay.iglob_lo[b], lay.jglob_lo[b] - jlo + 1
        w = 0 if ig == 1 else 1
        e = ni + 2 if ig + ni - 1 == nx else ni + 1
        s = 0 if jg == 1 else 1
        t = nj + 2 if jg + nj - 1 == n else nj + 1
        out[ig - 1 + w:ig - 1 + e, jg - 1 + s:jg - 1 + t] = blk[w:e, s:t, b]
    return out


if __name__ == "__main__":
    main()

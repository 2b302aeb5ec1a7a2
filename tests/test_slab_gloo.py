"""CPU test of the N > 1 path (world_size 2 and 3, torch.distributed `gloo`): the y-slab partition
and row-exchange schedule of cice4_b200.slab, executed with the oracle's primitives on every rank,
must reproduce the single-domain oracle bit for bit (doc/cicedoc.pdf 4.6)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kw, ndte, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from cice4_b200 import slab, synth
    from oracle import oracle as O
    from slab_oracle import slab_evp, slab_fields

    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.make_case(**kw)
    g = case.grid
    nx, ny = g.nx, g.ny
    rows = slab.slab_bounds(ny, world, rank)
    nyl = rows[1] - rows[0] + 1
    state = synth.zero_state(nx + 2, ny + 2)
    f, st = slab_fields(O, case, rows, state)
    p = O.make_params(ndte=ndte)
    north = rank + 1 if rank < world - 1 else None
    south = rank - 1 if rank > 0 else None

    def exchange(arrs):
        # whole padded rows; sends first (gloo isend), then receives
        reqs, recv = [], []
        for a in arrs:
            if north is not None:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[:, nyl])), north))
                t = torch.empty(nx + 2, dtype=torch.from_numpy(a[:1, 0]).dtype)
                reqs.append(dist.irecv(t, north))
                recv.append((a, nyl + 1, t))
            if south is not None:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[:, 1])), south))
                t = torch.empty(nx + 2, dtype=torch.from_numpy(a[:1, 0]).dtype)
                reqs.append(dist.irecv(t, south))
                recv.append((a, 0, t))
        for r in reqs:
            r.wait()
        for a, row, t in recv:
            a[:, row] = t.numpy()

    for _ in range(2):   # cold + warm call
        slab_evp(O, f, nx, rows, ny, g.ew, g.ns, rank, world, p, exchange)
    names = ["uvel", "vvel", "stressp_1", "stressm_3", "stress12_4", "strocnxT", "strintx", "divu", "prs_sig", "strength"]
    mine = {n: f[n].copy() for n in names}
    gathered = [None] * world
    dist.gather_object((rows, mine), gathered if rank == 0 else None, 0)
    if rank == 0:
        bounds = [x[0] for x in gathered]
        full = {n: slab.gather_slabs([x[1][n] for x in gathered], bounds, nx, ny) for n in names}
        st1 = synth.zero_state(nx + 2, ny + 2)
        f1 = None
        for _ in range(2):
            f1, _s = O.run_evp(g, case.inputs, st1, p)
        bad = []
        for n in names:
            ref = st1[n] if n in st1 else f1[n]
            I = (slice(None), slice(None)) if n in ("uvel", "vvel") else (slice(1, nx + 1), slice(1, ny + 1))
            if not np.array_equal(full[n][I], ref[I]):
                bad.append((n, float(np.abs(full[n][I] - ref[I]).max())))
        q.put(bad)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kw", [
    (2, dict(name="om1deg", nx=40, ny=30, realistic=True)),
    (3, dict(name="om1deg", nx=36, ny=31)),
    (2, dict(name="gx3", nx=30, ny=24, ew="cyclic", ns="open")),
], ids=["tripole-2", "tripole-3-uneven", "open-2"])
def test_slab_partition_reproduces_single_domain(oracle, world, kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, kw, 30, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    bad = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert bad == [], f"slab run differs from the single-domain oracle: {bad}"


def test_slab_bounds_cover_domain():
    from cice4_b200 import slab
    for ny in (16, 31, 1080, 2700):
        for n in (1, 2, 3, 4, 8):
            b = [slab.slab_bounds(ny, n, r) for r in range(n)]
            assert b[0][0] == 1 and b[-1][1] == ny
            for a, c in zip(b[:-1], b[1:]):
                assert c[0] == a[1] + 1
            sizes = [j1 - j0 + 1 for j0, j1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_reference_blocks_regroup_into_slabs():
    from cice4_b200 import slab
    # bld/config.nci.access-om.1440x1080: 96 x 2 blocks of 15 x 540 -> 2 slabs of 540 rows
    for r in range(2):
        lay = slab.slab_blocks_from_reference(1440, 1080, 15, 540, 2, r)
        assert lay.nblocks == 96 and slab.layout_rows(lay) == (1 + 540 * r, 540 * (r + 1))
    with pytest.raises(ValueError):
        slab.slab_blocks_from_reference(1440, 1080, 15, 540, 4, 0)

#!/usr/bin/env python
"""bench.py -- EVP grid-cell-subcycles/s (fp64) on B200, BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload om025]
                    [--configs auto|none]

A "step" is one pass of the hot path over one batch of synthetic input: one full ndte = 120
subcycle loop (stress + stepu + velocity halo update per subcycle) of one `evp(dt)` call on the
named grid.  `value` = nx*ny*ndte*K / t with the fields already resident in HBM (CUDA events on
the library's stream, max over ranks); `e2e` is the same metric through the public call
(`IceDynEvp.evp` -> evp_b200_step) with pinned HOST buffers: upload, prep, device ice_strength,
subcycles, finish and download inside the timed region.  One JSON line on stdout (rank 0).

The line's headline workload is access-om 0.25 deg (1440 x 1080, BASELINE.json configs[3], the grid the
metric is quoted on).  `configs` (default `--configs auto`) appends a table with the other BASELINE
configurations that make sense at this GPU count -- gx3, gx1, access-om 1 deg, the 0.1-degree class grid
at ndte = 120 and 240 and a weak-scaling slab of it -- each with its device-resident value, the
effective roofline fraction, an L2-residency flag and, on several GPUs, a bit-for-bit comparison of
the N-GPU result with the same problem run on one GPU (`parity_vs_1gpu`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EVP grid-cell-subcycles/sec (fp64)"
UNIT = "grid-cell-subcycles/s"
BYTES_T = 34 * 8   # per active T cell and subcycle: 12+12 stresses, strength, 9 metrics
BYTES_U = 14 * 8   # per active U cell and subcycle: u,v read+write, 10 U fields
L2_BYTES = 126e6


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return int(d["dram_bytes_read_per_launch"]) + int(d["dram_bytes_write_per_launch"]), d["source"]
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in o.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()),
                "samples": len(self.samples), "reasons": reasons}


# workloads beyond cice4_b200.synth.CONFIGS: name -> (base config, nx, ny)
def build_case(workload: str, realistic: bool, world: int = 1):
    from cice4_b200 import synth
    if workload == "p01w":   # weak scaling of the 0.1-degree class grid: 3600 x 338 rows per GPU
        c = synth.make_case("p01", nx=3600, ny=338 * world, realistic=realistic)
        c.name = "p01w"
        return c
    if "@" in workload:      # development: NAME@NXxNY = the named configuration on another grid size
        name, dims = workload.split("@")
        nx, ny = (int(v) for v in dims.split("x"))
        c = synth.make_case(name, nx=nx, ny=ny, realistic=realistic)
        c.name = name
        return c
    fixture = os.path.join(ROOT, "tests", "golden", "gx3_grid.npz") if workload == "gx3" else None
    return synth.make_case(workload, realistic=realistic, gx3_fixture=fixture)


def case_dt(case) -> float:
    from cice4_b200 import synth
    return synth.CONFIG_DT.get("p01" if case.name == "p01w" else case.name, 3600.0)


def workload_string(case, ndte: int, realistic: bool) -> str:
    """The one description of a workload both arms print (config.workload): grid, ndte, dt, mask, boundaries."""
    g = case.grid
    names = {0: "open", 1: "closed", 2: "cyclic", 3: "tripole", 4: "tripoleT"}
    return (f"{case.name} {g.nx}x{g.ny} ndte={ndte} dt={case_dt(case):g}s {'realistic' if realistic else 'dense'} mask, "
            f"ew={names.get(g.ew, g.ew)} ns={names.get(g.ns, g.ns)}, warm second call")


DT_NOTE = ("dt per grid as the ACCESS-OM2 configurations of these grids use it (0.25 deg: 1800 s, 0.1 deg: 600 s; "
           "1 deg and coarser: the reference's 3600 s): a builder choice where SURVEY 8(d) says 3600 s -- with dte = 30 s "
           "on the 0.25-degree grid the EVP iteration amplifies rounding noise to 1.9e-7 m/s between two CPU builds of "
           "the same code; the cost per subcycle does not depend on dt")


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own code on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_prepare(case, ndte: int):
    """Fields prepared by one (untimed) oracle call so that the timed legs run stress/stepu only."""
    from cice4_b200 import synth
    from oracle import oracle as O
    O.build()
    g = case.grid
    dt = case_dt(case)
    st = synth.zero_state(g.nx_block, g.ny_block)
    p = O.make_params(dt=dt, ndte=2, kind="fast")
    f, _ = O.run_evp(g, case.inputs, st, p, lib_kind="fast")
    p = O.make_params(dt=dt, ndte=ndte, kind="fast")
    have_ref = O.ref_available() and os.path.exists(os.path.join(O.REF_DIR, "libevp_ref_cice4_fast.so"))
    return O, g, f, p, dt, have_ref


def cpu_loop_seconds(O, g, f, p, dt, nsub, have_ref, threads):
    """Seconds for `nsub` subcycles of the reference's own stress/stepu (+ halo updates) on `threads`
    host threads (the translated Fortran, oracle/_ref) or, when that is absent, of the oracle port."""
    if have_ref:
        return O.time_subcycles_ref(g, f, p, dt, nsub, threads=threads), "reference"
    return O.time_subcycles(g, f, p, nsub, lib_kind="fast"), "port"


def cpu_baseline(case, ndte: int, budget_s: float = 12.0, threads: int = 0):
    """Bounded CPU sample inside the b200 arm (rank 0, N = 1): full ndte loops of the reference's own
    stress / stepu -- its Fortran machine-translated to C at build time (oracle/_ref), gcc -O3, the cell
    loops on all host threads (kind "reference") -- repeated for about `budget_s` seconds; next to it one
    loop of the oracle port (kind "port") and a few subcycles of the reference on ONE thread (its serial
    build).  The value is the reference's when it is available."""
    cores = threads or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    O, g, f, p, dt, have_ref = cpu_prepare(case, ndte)
    what = f"{workload_string(case, ndte, False).split(',')[0]} (stress+stepu+2 halo updates per subcycle)"
    O.omp_set_num_threads(cores)
    t1, kind = cpu_loop_seconds(O, g, f, p, dt, 2, have_ref, cores)      # warm-up / probe
    per_loop = t1 / 2.0 * ndte
    loops = int(max(1, min(10, budget_s / max(per_loop, 1e-3))))
    nsub = ndte if per_loop <= budget_s else int(max(2, ndte * budget_s / per_loop))
    secs = []
    for _ in range(loops):
        s, kind = cpu_loop_seconds(O, g, f, p, dt, nsub, have_ref, cores)
        secs.append(s)
    tot = sum(secs)
    base = {"value": g.nx * g.ny * nsub * loops / tot, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{loops} x {nsub} subcycles of {what}; "
                      + ("the reference's own stress/stepu (Fortran machine-translated to C, oracle/_ref), gcc -O3 "
                         "-fopenmp over the cell lists" if kind == "reference" else "oracle C port gcc -O3 -fopenmp")
                      + f"; {tot:.2f} s, per-loop min/max {min(secs):.2f}/{max(secs):.2f} s"}
    if have_ref:
        psub = int(max(2, min(ndte, 3.0 / max(per_loop / ndte, 1e-6))))
        psec = O.time_subcycles(g, f, p, psub, lib_kind="fast")
        base["port"] = {"value": g.nx * g.ny * psub / psec, "cores": cores, "kind": "port",
                        "sample": f"{psub} subcycles, oracle C port gcc -O3 -fopenmp, {psec:.2f} s"}
        ssub = int(max(1, min(ndte, 2.0 / max(per_loop / ndte * cores * 0.7, 1e-6))))
        ssec = O.time_subcycles_ref(g, f, p, dt, ssub, threads=1)
        O.omp_set_num_threads(cores)
        base["reference_serial"] = {"value": g.nx * g.ny * ssub / ssec, "cores": 1, "kind": "reference",
                                    "sample": f"{ssub} subcycles, same code on one thread (the reference's serial build), {ssec:.2f} s"}
    return base


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (its
    Fortran stress / stepu machine-translated to C, oracle/_ref; the oracle port only where that is
    absent).  Same workload, dt, metric and unit as the b200 arm; a step is one full ndte subcycle loop
    (bounded to a share of it only when one loop would take more than ~20 s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = build_case(args.workload, args.realistic, 1)
    g = case.grid
    ndte = args.ndte
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    O, g, f, p, dt, have_ref = cpu_prepare(case, ndte)
    O.omp_set_num_threads(cores)
    t1, kind = cpu_loop_seconds(O, g, f, p, dt, 2, have_ref, cores)
    per_loop = t1 / 2.0 * ndte
    budget = 240.0 / max(1, args.steps + args.warmup)
    nsub = ndte if per_loop <= max(20.0, budget) and per_loop * (args.steps + args.warmup) <= 300.0 else \
        int(max(2, ndte * min(budget, 20.0) / per_loop))
    secs = []
    for k in range(args.warmup + args.steps):
        s, kind = cpu_loop_seconds(O, g, f, p, dt, nsub, have_ref, cores)
        if k >= args.warmup:
            secs.append(s)
    tot = sum(secs)
    v = g.nx * g.ny * nsub * len(secs) / tot
    base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(secs)} x {nsub} subcycles; "
                      + ("the reference's own stress/stepu (Fortran machine-translated to C, oracle/_ref), gcc -O3 -fopenmp "
                         "over the cell lists" if kind == "reference" else "oracle C port gcc -O3 -fopenmp")
                      + f"; per-step min/max {min(secs):.3f}/{max(secs):.3f} s"}
    if have_ref:   # for context: the oracle port on the same threads and the reference on one thread (serial build)
        per_sub = tot / (nsub * len(secs))
        psub = int(max(2, min(ndte, 3.0 / max(per_sub, 1e-6))))
        psec = O.time_subcycles(g, f, p, psub, lib_kind="fast")
        base["port"] = {"value": g.nx * g.ny * psub / psec, "cores": cores, "kind": "port",
                        "sample": f"{psub} subcycles, oracle C port gcc -O3 -fopenmp, {psec:.2f} s"}
        ssub = int(max(1, min(ndte, 3.0 / max(per_sub * cores * 0.7, 1e-6))))
        ssec = O.time_subcycles_ref(g, f, p, dt, ssub, threads=1)
        O.omp_set_num_threads(cores)
        base["reference_serial"] = {"value": g.nx * g.ny * ssub / ssec, "cores": 1, "kind": "reference",
                                    "sample": f"{ssub} subcycles, same code on one thread (the reference's serial build), {ssec:.2f} s"}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, len(secs)) * (ndte / nsub),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(case, ndte, args.realistic),
                       "step": "one ndte subcycle loop (stress+stepu+halo)"
                               + ("" if nsub == ndte else f"; bounded sample: {nsub} of {ndte} subcycles per step, scaled"),
                       "dt_note": DT_NOTE},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class Ctx:
    """torch / torch.distributed plumbing of one bench process (one rank per GPU)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def allmax(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def uid(self, E):
        torch = self.torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            uid.copy_(torch.tensor(list(E.IceDynEvp.comm_unique_id()), dtype=torch.uint8))
        self.dist.broadcast(uid, 0)
        return bytes(uid.cpu().tolist())


def make_dyn(ctx, case, args, ndte, world=None, rank=None, **over):
    """One handle for this rank's y-slab of `case` (world = 1: the whole domain on this rank's GPU)."""
    from cice4_b200 import evp as E
    from cice4_b200 import slab
    world = ctx.world if world is None else world
    rank = ctx.rank if rank is None else rank
    g = case.grid
    ew = {v: k for k, v in E.BND.items()}[g.ew]
    ns = {v: k for k, v in E.BND.items()}[g.ns]
    lay = slab.slab_layout(g.nx, g.ny, world, rank)
    rows = slab.layout_rows(lay)
    par = dict(ndte=ndte, math_mode=args.math_mode, pin_host=1, tile_threads=args.tile_threads,
               tile_rows=args.tile_rows, kernel_variant=args.variant)
    par.update(over)
    dyn = E.IceDynEvp(lay, ew, ns, device=ctx.local_rank, rank=rank, nranks=world, slab=rows, **par)
    gf = E.grid_fields_in_blocks(g, lay, ew, ns)
    dyn.init_evp(case_dt(case), gf)
    if world > 1:
        dyn.comm_init(ctx.uid(E))
    inputs = {k: E.split_blocks(v, lay, ew, ns) for k, v in case.inputs.items()}
    return dyn, lay, rows, inputs


def local_padded(blk, lay, rows, nx):
    """block layout of one slab (one block per slab here) -> padded (nx+2, rows+2) array"""
    return np.ascontiguousarray(blk[:, :, 0])


def parity_vs_one_gpu(ctx, case, args, ndte, dyn, rows, want):
    """N > 1: rank 0 runs the same two calls (cold start, warm call) for the WHOLE domain on its own GPU
    and compares uvel, vvel and stressp_1 gathered from the N slabs bit for bit with it."""
    from cice4_b200 import slab
    g = case.grid
    names = ("uvel", "vvel", "stressp_1")
    mine = {n: local_padded(dyn.state[n], None, rows, g.nx) for n in names}
    gathered = [None] * ctx.world
    ctx.dist.gather_object((rows, mine), gathered if ctx.rank == 0 else None, 0)
    verdict = None
    if ctx.rank == 0:
        try:
            one, lay1, rows1, inputs1 = make_dyn(ctx, case, args, ndte, world=1, rank=0, pin_host=0)
            dt = case_dt(case)
            o = one.evp(dt, inputs1, strength=None, want=want)
            s1 = o["strength"].copy(order="F")
            one.evp(dt, inputs1, strength=s1, want=want)
            bad = []
            bounds = [gp[0] for gp in gathered]
            for n in names:
                full = slab.gather_slabs([gp[1][n] for gp in gathered], bounds, g.nx, g.ny)
                ref = one.state[n][:, :, 0]
                I = (slice(0, g.nx + 2), slice(0, g.ny + 2)) if n in ("uvel", "vvel") else (slice(1, g.nx + 1), slice(1, g.ny + 1))
                if not np.array_equal(full[I], ref[I]):
                    bad.append(f"{n}: max|diff| {np.abs(full[I] - ref[I]).max():.3e}")
            one.finalize()
            verdict = "bit-exact (uvel, vvel, stressp_1 after cold + warm call)" if not bad else "MISMATCH " + "; ".join(bad)
        except Exception as e:   # the other ranks wait in the broadcast below: never leave them there
            verdict = f"NOT CHECKED: {type(e).__name__}: {e}"
    obj = [verdict]
    ctx.dist.broadcast_object_list(obj, 0)
    return obj[0]


def measure(ctx, case, args, ndte, steps, warmup, full):
    """Cold + warm evp call, parity against one GPU (N > 1), the device-resident loop (value) and the
    end-to-end call.  full = the headline workload: clocks, resident-state e2e and per-phase times too."""
    g = case.grid
    nx, ny = g.nx, g.ny
    dt = case_dt(case)
    world, rank = ctx.world, ctx.rank
    dyn, lay, rows, inputs = make_dyn(ctx, case, args, ndte)
    from cice4_b200 import evp as E
    want = [n for n in E.OUTPUT_D if n not in ("sig1", "sig2", "sicemass")]
    # cold start (iceumask = .false. => u = uocn), then warm calls are what is timed
    out = dyn.evp(dt, inputs, strength=None, want=want)
    strength = out["strength"].copy(order="F")
    out = dyn.evp(dt, inputs, strength=strength, want=want, two_phase=True)
    res = {}
    if world > 1:
        res["parity_vs_1gpu"] = parity_vs_one_gpu(ctx, case, args, ndte, dyn, rows, want)
    nyl = rows[1] - rows[0] + 1
    top = nyl + 2 if rank == world - 1 else nyl + 1        # the domain's north ghost row belongs to the last slab
    icellt = int(ctx.allsum(float(out["icetmask"][1:, 1:top, 0].sum())))
    icellu = int(ctx.allsum(float(dyn.state["iceumask"].sum())))
    bytes_per_sub = BYTES_T * icellt + BYTES_U * icellu

    # ---- device-resident subcycle loop: the headline value ---------------------------------------
    for _ in range(warmup):
        dyn.subcycle_resident(1)
    ctx.barrier()
    clocks = None
    if full:
        with ClockSampler(ctx.local_rank) as cs:
            t0 = time.perf_counter()
            ms_loop = ctx.allmax(dyn.subcycle_resident(steps))
            ctx.barrier()
            wall = time.perf_counter() - t0
            if wall < 1.5:   # give nvidia-smi a few samples under the same load (not part of the number)
                dyn.subcycle_resident(max(1, int(1.5 / max(ms_loop * 1e-3, 1e-4))))
        clocks = cs.summary()
    else:
        ms_loop = ctx.allmax(dyn.subcycle_resident(steps))
        ctx.barrier()
    tm = dyn.timings()
    try:
        res["kernel_info"] = {k: int(v) for k, v in dyn.info().items()}   # which subcycle kernel ran (evp_b200_get_info)
    except Exception as e:   # informative only
        res["kernel_info"] = {"error": str(e)}
    kernel_s = ms_loop * 1e-3 / ndte
    achieved = bytes_per_sub / kernel_s / 1e9        # whole job, all GPUs
    peak, peak_src = measured_peaks()
    ws = bytes_per_sub / world
    res.update(value=nx * ny * ndte / (ms_loop * 1e-3), ms_loop=ms_loop, kernel_us=kernel_s * 1e6,
               achieved=achieved / world, frac=achieved / world / peak, peak=peak, peak_src=peak_src,
               bytes_per_sub=bytes_per_sub, icellt=icellt, icellu=icellu, tm=tm, clocks=clocks,
               l2=(f"working set {ws / 1e6:.0f} MB per subcycle per GPU vs 126 MB L2: "
                   + ("inputs larger than L2" if ws > 1.6 * L2_BYTES else "fits L2 -> effective bandwidth, latency-bound")),
               l2_resident=bool(ws <= 1.6 * L2_BYTES))

    # ---- end to end through the public call, host buffers, device ice_strength inside --------------
    e2e_steps = steps if full else max(2, min(steps, 5))
    for _ in range(2):
        dyn.evp(dt, inputs, strength=None, want=want)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dyn.evp(dt, inputs, strength=None, want=want)
    ctx.barrier()
    e2e_s = ctx.allmax((time.perf_counter() - t0) / e2e_steps)
    tme = dyn.timings()
    plane = lay.nx_block * lay.ny_block * lay.max_blocks
    ncat = case.inputs["aicen"].shape[2]
    res.update(e2e_s=e2e_s, tme=tme,
               h2d=int(ctx.allsum((7 + 14 + 1 + 2 * ncat) * plane * 8 + plane * 4)),   # inputs, state, aice0 + categories, iceumask
               d2h=int(ctx.allsum((14 + len(want)) * plane * 8 + plane * 4)))
    dyn.finalize()

    if full:
        # the same call with the state resident on the device: state_residency = 1 keeps the 12
        # stress arrays there, 2 also uvel / vvel / iceumask; with 2 the timed region includes the download of
        # uvel and vvel that the transport scheme needs right after evp (evp_b200_download_velocity)
        for mode, key in ((1, "e2e_res1_s"), (2, "e2e_res2_s")):
            dyn, lay, rows, inputs = make_dyn(ctx, case, args, ndte, state_residency=mode)
            for _ in range(3):
                dyn.evp(dt, inputs, strength=None, want=want)
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                dyn.evp(dt, inputs, strength=None, want=want)
                if mode == 2:
                    dyn.download_velocity()
            ctx.barrier()
            res[key] = ctx.allmax((time.perf_counter() - t0) / e2e_steps)
            dyn.finalize()
        if not args.math_mode:
            # the same device-resident loop with the FMA-contracted build (math_mode = 1: inside the 1e-10 tolerance
            # of the north star, tests/test_parity_gpu.py::test_fma_mode_within_tolerance, but not bit-exact)
            dyn, lay, rows, inputs = make_dyn(ctx, case, args, ndte, math_mode=1)
            dyn.evp(dt, inputs, strength=None, want=want)
            dyn.evp(dt, inputs, strength=None, want=want)
            for _ in range(warmup):
                dyn.subcycle_resident(1)
            ctx.barrier()
            res["ms_loop_fma"] = ctx.allmax(dyn.subcycle_resident(steps))
            ctx.barrier()
            dyn.finalize()
    return res


def extra_configs(world: int):
    """(workload, ndte) of the `configs` table at this GPU count (BASELINE.json configs 0-2 and 4)."""
    if world == 1:
        return [("gx3", 120), ("gx1", 120), ("om1deg", 120), ("p01", 120), ("p01", 240), ("p01w", 120),
                ("om025:realistic", 120)]   # the headline grid with a realistic ice mask (20 % of the cells active)
    if world == 2:
        return [("om1deg", 120), ("p01", 120), ("p01w", 120)]
    if world == 4:
        return [("p01", 120), ("p01w", 120)]
    return [("p01", 120), ("p01", 240), ("p01w", 120)]


def run_b200(args):
    from cice4_b200 import build as B
    ctx = Ctx(args)
    if ctx.rank == 0:
        B.build()
    if ctx.dist:
        ctx.dist.barrier()
    world, rank = ctx.world, ctx.rank
    warmup = max(3, args.warmup)
    case = build_case(args.workload, args.realistic, world)
    ndte = args.ndte
    r = measure(ctx, case, args, ndte, args.steps, warmup, full=True)
    g = case.grid
    nx, ny = g.nx, g.ny

    base = {"value": None}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(case, ndte)

    table = []
    if args.configs == "auto" and args.workload == "om025" and not args.realistic:
        for wl, nd in extra_configs(world):
            realistic = wl.endswith(":realistic")
            wl = wl.split(":")[0]
            c = build_case(wl, realistic, world)
            x = measure(ctx, c, args, nd, max(2, min(args.steps, 5)), warmup, full=False)
            table.append({"workload": workload_string(c, nd, realistic), "n_gpus": world, "value": x["value"], "unit": UNIT,
                          "us_per_subcycle": x["kernel_us"], "roofline_frac_per_gpu": x["frac"],
                          "l2_resident": x["l2_resident"], "active_T_cells": x["icellt"], "active_U_cells": x["icellu"],
                          "e2e_value": c.grid.nx * c.grid.ny * nd / x["e2e_s"],
                          "e2e_ms_per_call": x["e2e_s"] * 1e3,
                          "scaling": "weak (3600 x 338 rows per GPU)" if wl == "p01w" else "strong",
                          "exchange_mode_used": int(x["tm"]["exchange_mode_used"]),
                          **({"parity_vs_1gpu": x["parity_vs_1gpu"]} if "parity_vs_1gpu" in x else {})})
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if rank != 0:
        return

    traffic, traffic_src = measured_traffic()
    if world != 1 or args.workload != "om025" or args.realistic:
        traffic, traffic_src = None, None            # the capture is of the 1-GPU om025 dense launch
    tm, tme = r["tm"], r["tme"]
    xmode = int(tm["exchange_mode_used"])
    xname = {-1: "single rank: no exchange", 0: "peer-to-peer stores from the subcycle kernel into the neighbour's "
             "ghost rows (CUDA IPC over NVLink, per-strip epoch flags)", 1: "NCCL send/recv of the boundary rows after every subcycle"}[xmode]
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": r["ms_loop"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(case, ndte, args.realistic),
                   "step": "one ndte subcycle loop (stress+stepu+halo) on device-resident fields",
                   "dt_note": DT_NOTE,
                   "active_T_cells": r["icellt"], "active_U_cells": r["icellu"],
                   "l2": r["l2"],
                   "math_mode": "fma-contracted (<=1e-10 of the unfused oracle)" if args.math_mode else "unfused (bit-exact vs oracle)",
                   "tile": {"threads": args.tile_threads, "rows": args.tile_rows, "variant": args.variant},
                   "kernel": r.get("kernel_info"),
                   "parallelism": f"{world} y-slab(s), one process per GPU; velocity halo inside the loop: {xname}",
                   "exchange_mode_used": xmode},
        "roofline": {"bound": "hbm", "achieved": r["achieved"], "peak": r["peak"], "unit": "GB/s",
                     "frac": r["frac"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": r["peak_src"],
                     "per": "GPU", "kernel": "fused stress+stepu subcycle kernel", "kernel_us": r["kernel_us"],
                     "algorithmic_bytes_per_launch": r["bytes_per_sub"] / world,
                     "frac_of_nominal_8TBs": r["achieved"] / 8000.0},
        "cpu_baseline": base,
        "e2e": {"value": nx * ny * ndte / r["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "ms_per_call": r["e2e_s"] * 1e3,
                "device_breakdown_ms_rank0": {k: round(v, 3) for k, v in tme.items() if k.endswith("_ms")},
                "call": "IceDynEvp.evp(dt, inputs, strength=None): upload, prep, device ice_strength, ndte loop, finish, download",
                "state": "full round trip of uvel, vvel, 12 stresses, iceumask every call (restart-exact drop-in)",
                "resident_stresses": {"value": nx * ny * ndte / r["e2e_res1_s"], "ms_per_call": r["e2e_res1_s"] * 1e3,
                                      "note": "state_residency=1: the 12 stress arrays stay on the device"},
                "resident_state": {"value": nx * ny * ndte / r["e2e_res2_s"], "ms_per_call": r["e2e_res2_s"] * 1e3,
                                   "note": "state_residency=2: stresses, uvel, vvel, iceumask stay on the device; the call "
                                           "is followed by evp_b200_download_velocity (transport hand-off) inside the timed region"}},
        "gpu_launches": int(tm["subcycle_launches"]) * args.steps * world,
        "clocks": r["clocks"],
    }
    if "ms_loop_fma" in r:
        t = r["ms_loop_fma"] * 1e-3
        line["fma_mode"] = {"value": nx * ny * ndte / t, "unit": UNIT, "ms_per_step": r["ms_loop_fma"],
                            "kernel_us": t / ndte * 1e6,
                            "roofline_frac": r["bytes_per_sub"] / (t / ndte) / 1e9 / world / r["peak"],
                            "note": "math_mode=1 (nvcc contracts a*b+c into DFMA, as the reference's -O3 -xHost build "
                                    "may): within the 1e-10 tolerance of the unfused oracle, not bit-exact; the headline "
                                    "value above is the bit-exact build"}
    if "parity_vs_1gpu" in r:
        line["parity_vs_1gpu"] = r["parity_vs_1gpu"]
    if table:
        line["configs"] = table
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="om025")
    ap.add_argument("--configs", default="auto", choices=["auto", "none"],
                    help="auto: append the table of the other BASELINE configurations (om025 dense runs only)")
    ap.add_argument("--realistic", action="store_true")
    ap.add_argument("--ndte", type=int, default=120)
    ap.add_argument("--math-mode", type=int, default=0)
    ap.add_argument("--tile-threads", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    except BaseException:
        # a rank that fails must not linger in destructors (NCCL teardown, stream syncs on peer flags) while
        # the other ranks wait for it in a collective: report and leave at once so that the launcher ends the job
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        if isinstance(sys.exc_info()[1], SystemExit) and sys.exc_info()[1].code in (0, None):
            return
        os._exit(1)


if __name__ == "__main__":
    main()

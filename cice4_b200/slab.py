"""y-slab partition of the global domain across GPUs (host logic, numpy only).

Re-maps the reference's block decomposition (/root/reference/source/ice_blocks.F90:133-350,
ice_distribution.F90 `create_distrb_cart`) so that each GPU owns a contiguous slab of rows
[jlo, jhi] with full x extent: the east-west wrap stays inside a GPU, the tripole fold lives on the
last rank, and every `ice_HaloUpdate` that `evp` performs becomes one exchange of whole padded
rows with the south and north neighbour.  The exchange points of one `evp` call,
in order, are listed in EXCHANGES; the CUDA library and the CPU (gloo) test harness follow it.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .evp import BlockLayout

# (field, location, kind, when) -- every halo update of source/ice_dyn_evp.F90:119-432
EXCHANGES = [
    ("icetmask", "center", "scalar", "after evp_prep1 (:250-253)"),
    ("strairx", "center", "vector", "t2ugrid_vector (:276)"),
    ("strairy", "center", "vector", "t2ugrid_vector (:277)"),
    ("strength", "center", "scalar", "after ice_strength (:337-338)"),
    ("uvel", "NEcorner", "vector", "before the subcycle loop (:340-341)"),
    ("vvel", "NEcorner", "vector", "before the subcycle loop (:342-343)"),
    ("uvel+vvel", "NEcorner", "vector", "after stepu, every subcycle (:397-402)"),
    ("strocnxT", "NEcorner", "vector", "u2tgrid_vector (:427)"),
    ("strocnyT", "NEcorner", "vector", "u2tgrid_vector (:428)"),
]


def slab_bounds(ny_global: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Rows [jlo, jhi] (1-based, inclusive) of `rank`; remainders go to the southern slabs."""
    base, rem = divmod(ny_global, nranks)
    jlo = rank * base + min(rank, rem) + 1
    n = base + (1 if rank < rem else 0)
    if n < 2:
        raise ValueError("each slab needs at least 2 rows")
    return jlo, jlo + n - 1


def slab_layout(nx_global: int, ny_global: int, nranks: int, rank: int) -> BlockLayout:
    """One block per rank covering its slab: nx_block = nx+2, ny_block = rows+2."""
    jlo, jhi = slab_bounds(ny_global, nranks, rank)
    n = jhi - jlo + 1
    return BlockLayout(nx_global, ny_global, nx_global + 2, n + 2, [2], [nx_global + 1], [2], [n + 1], [1], [jlo])


def slab_blocks_from_reference(nx_global: int, ny_global: int, bx: int, by: int, nranks: int, rank: int) -> BlockLayout:
    """Reference block layout (bx x by blocks) regrouped into slabs: rank r gets the block rows
    whose j range falls into its share; requires nblocks_y to be a multiple of nranks."""
    full = BlockLayout.cartesian(nx_global, ny_global, bx, by)
    nby = (ny_global - 1) // by + 1
    if nby % nranks:
        raise ValueError("nblocks_y must be a multiple of the number of GPUs")
    per = nby // nranks
    jlo = rank * per * by + 1
    jhi = min((rank + 1) * per * by, ny_global)
    keep = [b for b in range(full.nblocks) if jlo <= full.jglob_lo[b] <= jhi]
    pick = lambda v: [int(v[b]) for b in keep]
    return BlockLayout(nx_global, ny_global, full.nx_block, full.ny_block, pick(full.ilo), pick(full.ihi),
                       pick(full.jlo), pick(full.jhi), pick(full.iglob_lo), pick(full.jglob_lo))


def layout_rows(layout: BlockLayout) -> Tuple[int, int]:
    jlo = int(min(layout.jglob_lo))
    jhi = int(max(layout.jglob_lo[b] + (layout.jhi[b] - layout.jlo[b]) for b in range(layout.nblocks)))
    return jlo, jhi


def gather_slabs(parts: List[np.ndarray], bounds: List[Tuple[int, int]], nx: int, ny: int) -> np.ndarray:
    """Padded per-slab arrays (nx+2, rows+2) -> padded global array; ghost rows of the domain come
    from the first / last slab."""
    out = np.zeros((nx + 2, ny + 2), dtype=parts[0].dtype, order="F")
    for a, (jlo, jhi) in zip(parts, bounds):
        out[:, jlo:jhi + 1] = a[:, 1:jhi - jlo + 2]
    out[:, 0] = parts[0][:, 0]
    out[:, ny + 1] = parts[-1][:, -1]
    return out

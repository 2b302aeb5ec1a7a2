// evp_tiled.cuh -- the strip-tiled (AoSoA) layout of everything the ndte subcycle loop touches.
//
// The once-per-call kernels work on planes (evp_common.cuh).  The subcycle loop -- 120 launches per call
// that each stream the whole working set once -- reads and writes a second layout made for it:
//
//   * the domain is cut into STRIPS of 31 U columns; strip w holds U columns 31w+1 .. 31w+31 and the T
//     columns 31w+1 .. 31w+32 (32 "slots": a T cell needs the U point west of it and the U point of its
//     own column, a U point needs the T cell east of it -- slot 31 duplicates the first column of strip w+1
//     and is recomputed redundantly from the same inputs, exactly like the reference's redundant N/E ghost
//     cells, source/ice_dyn_evp.F90:846-859);
//   * one warp owns one strip and marches north; everything the warp needs for T row j and U row j-1 is
//     ONE contiguous "tile row" in memory, so a single bulk copy (cp.async.bulk, TMA) brings it into
//     shared memory, and consecutive rows of a strip are consecutive in memory:
//
//       tile row (strip w, row j), 1552 doubles = 12416 bytes:
//         [ state copy 0 | loop-invariant block | state copy 1 ]
//       state copy (452 doubles): u[32] v[32] stress[12][32] halo{u_west, v_west, -, -}
//           u/v slot l = U column 31w+1+l (slot 31 = first U column of strip w+1, kept equal by the writer),
//           halo = U column 31w (last column of strip w-1, kept equal by the writer)
//       loop-invariant block (648 doubles): strength, dxt, dyt, dxhy, dyhx, cxp, cyp, cxm, cym, tinyarea of
//           T row j (10 x 32); aiu, uocn, vocn, waterx, watery, forcex, forcey, umassdtei, fm, uarear of U
//           row j-1 (10 x 32: the U row that is completed together with T row j); 32 icetmask bytes of T row
//           j and 32 iceumask bytes of U row j-1
//     The invariant block sits BETWEEN the two ping-pong copies so that "old copy + invariants" is one
//     contiguous 8800-byte window whichever copy is the old one.
//
// Traffic per strip row: 8800 B read + 3616 B written for 31 x 1 cells = 400.5 B per cell (the 384 B of the
// plane layout x 32/31 for the duplicated slot, + masks and halo words).
#pragma once

#include <cstddef>
#include <cstdint>

#define EVT_SLOTS 32                       // slots per segment = T columns of a strip
#define EVT_UW 31                          // U columns owned by a strip
#define EVT_STATE_D (14 * 32 + 4)          // doubles per state copy: u, v, 12 stresses, halo
#define EVT_INV_D (20 * 32 + 8)            // doubles of the loop-invariant block (masks: 64 bytes)
#define EVT_ROW_D (2 * EVT_STATE_D + EVT_INV_D)   // doubles per tile row (1552)
#define EVT_WIN_D (EVT_STATE_D + EVT_INV_D)       // doubles of the read window (1100 = 8800 bytes)
// offsets inside a state copy
#define EVT_U 0
#define EVT_V 32
#define EVT_S(k) (64 + 32 * (k))
#define EVT_HALO 448                       // u_west at +0, v_west at +1
// offsets inside the loop-invariant block
#define EVT_T(q) (32 * (q))                // q: 0 strength 1 dxt 2 dyt 3 dxhy 4 dyhx 5 cxp 6 cyp 7 cxm 8 cym 9 tinyarea
#define EVT_UF(q) (320 + 32 * (q))         // q: 0 aiu 1 uocn 2 vocn 3 waterx 4 watery 5 forcex 6 forcey 7 umassdtei 8 fm 9 uarear
#define EVT_MASK 640                       // byte l: icetmask of T row j; byte 32 + l: iceumask of U row j-1

// where the pieces of a tile row start (doubles from the start of the tile row)
__host__ __device__ inline int evt_state_off(int copy) { return copy ? EVT_STATE_D + EVT_INV_D : 0; }
__host__ __device__ inline int evt_inv_off() { return EVT_STATE_D; }
__host__ __device__ inline int evt_window_off(int old_copy) { return old_copy ? EVT_STATE_D : 0; }
__host__ __device__ inline int evt_nstrips(int nx) { return (nx + EVT_UW - 1) / EVT_UW; }

// Tile row (strip w, row j) starts at  w * sw + j * sj  doubles.  Two orders: strip-major (sw = nr * EVT_ROW_D,
// sj = EVT_ROW_D: the rows a warp marches over are consecutive in memory) and row-major (sw = EVT_ROW_D,
// sj = ns * EVT_ROW_D: the tile rows that the warps of one row chunk read at about the same time are
// consecutive, i.e. every row front of the sweep is one long contiguous run for the DRAM pages).
struct TileGeom {
    double *tiles;
    int ns, nr;        // strips, rows per strip (nyl + 2)
    long long sw, sj;  // strides (doubles) between strips / between rows
};
__host__ __device__ inline void evt_set_order(TileGeom &g, int row_major) {
    g.sw = row_major ? (long long)EVT_ROW_D : (long long)g.nr * EVT_ROW_D;
    g.sj = row_major ? (long long)g.ns * EVT_ROW_D : (long long)EVT_ROW_D;
}

// U column i (0 .. nx+1) of row j: where its PRIMARY copy lives.  Returns the offset (doubles) of u from
// the start of the tile pool; v sits dv doubles behind it.
__host__ __device__ inline size_t evt_u_primary(int i, int j, int nx, int ns, long long sw, long long sj, int copy, int &dv) {
    if (i == 0) { // west ghost column: the halo word of strip 0
        dv = 1;
        return (size_t)(j * sj) + evt_state_off(copy) + EVT_HALO;
    }
    int w = (i - 1) / EVT_UW, l = (i - 1) - w * EVT_UW;
    if (w >= ns) { // column nx+1 when nx is a multiple of 31: slot 31 of the last strip
        w = ns - 1;
        l = i - 1 - w * EVT_UW;
    }
    dv = 32;
    return (size_t)(w * sw + j * sj) + evt_state_off(copy) + EVT_U + l;
}

// evp_aux.cuh -- launchers of the once-per-call kernels (marshalling, prep, halo, finish).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "evp_common.cuh"
#include "evp_tiled.cuh"

// geometry of one handle's slab planes
struct PlaneGeom {
    int nx, nyl, pitch;
    int ew_cyclic, ns_cyclic;
    int tripole;   // 1 when this slab holds the tripole fold (top rank, ns = tripole or tripoleT)
    int tfold;     // 1: the fold is a T-fold (ns = tripoleT), else a u-fold
    size_t cells;  // pitch * (nyl + 2)
};

// block layout of the caller's host arrays; device copy of per-block ints:
// tab[b*6 + {0..5}] = ilo, ihi, jlo, jhi (1-based), ishift, jshift with
// plane_i = local_i(1-based) + ishift, plane_j = local_j(1-based) + jshift
struct BlockGeom {
    int nx_block, ny_block, nblocks;
    const int *tab;
};

// download policies (which cells of a block the reference's evp defines)
enum {
    PACK_FULL = 0,        // every cell of the block (uvel, vvel, icetmask, tmass, strength)
    PACK_TNE_KEEP = 1,    // [ilo..ihi+1]x[jlo..jhi+1]; elsewhere 0 where icetmask==0, else untouched (stresses)
    PACK_TNE_ZERO = 2,    // [ilo..ihi+1]x[jlo..jhi+1]; elsewhere 0 (prs_sig, divu, shear, rdg_*)
    PACK_INT_ZERO = 3,    // interior; elsewhere 0 (U-point outputs, strocnxT)
    PACK_INT_KEEP = 4     // interior; elsewhere untouched (iceumask)
};

void aux_unblock_r8(const BlockGeom &bg, const PlaneGeom &pg, const double *blocked, double *plane, cudaStream_t s);
void aux_unblock_mask(const BlockGeom &bg, const PlaneGeom &pg, const int32_t *blocked, uint8_t *plane, cudaStream_t s);
void aux_block_r8(const BlockGeom &bg, const PlaneGeom &pg, const double *plane, const uint8_t *icetmask,
                  double *blocked, int policy, cudaStream_t s);
void aux_block_mask(const BlockGeom &bg, const PlaneGeom &pg, const uint8_t *plane, int32_t *blocked,
                    int policy, cudaStream_t s);

// ice_HaloUpdate for a slab plane: east-west wrap, north-south cyclic, tripole u-fold.
// loc: 1 centre, 2 NE corner; isign: +1 scalar, -1 vector.  (slab-to-slab rows are exchanged by the caller)
void aux_halo_r8(const PlaneGeom &pg, double *plane, int loc, int isign, cudaStream_t s);
void aux_halo_u8(const PlaneGeom &pg, uint8_t *plane, cudaStream_t s);
// uvel+vvel together (NE corner, vector): north-south part only (east-west is done by the subcycle kernel)
void aux_halo_uv_ns(const PlaneGeom &pg, double *u, double *v, cudaStream_t s);

struct PrepArgs {
    // static
    const double *tarea, *uarea, *fcor;
    const uint8_t *tmask, *umask;
    // inputs
    const double *aice, *vice, *vsno, *uocn, *vocn, *ss_tltx, *ss_tlty;
    // scratch / outputs
    double *tmass, *umass, *aiu, *umassdtei, *waterx, *watery, *forcex, *forcey;
    double *strairx, *strairy, *strtltx, *strtlty, *strintx, *strinty, *strocnx, *strocny, *fm;
    double *uvel, *vvel;
    double *stress[EVP_NSTRESS];
    uint8_t *tmphm, *icetmask, *iceumask;
    double rhoi, rhos, dtei, cosw, sinw, gravit;
    int hemisphere_turning, coupled_tilt, use_ocnslope;
};

void aux_prep1(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s);      // tmass, tmphm   (:643-677)
void aux_icetmask(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s);   // icetmask       (:679-691)
void aux_to_ugrid(const PlaneGeom &pg, const double *w1, const double *tarea, const double *uarea,
                  double *w2, cudaStream_t s);                               // ice_grid.F90:1612-1631
void aux_to_tgrid(const PlaneGeom &pg, const double *w1, const double *tarea, const double *uarea,
                  double *w2, cudaStream_t s);                               // ice_grid.F90:1720-1730
void aux_prep2(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s);      // :819-936

struct FinishArgs {
    const double *uvel, *vvel, *uocn, *vocn, *aiu, *fm;
    const uint8_t *iceumask;
    double *strocnx, *strocny, *strocnxT, *strocnyT;
    double dragw, cosw, sinw;
    int hemisphere_turning;
};
void aux_finish(const PlaneGeom &pg, const FinishArgs &a, cudaStream_t s);   // :1510-1547

void aux_principal_stress(size_t n, const double *sp1, const double *sm1, const double *s12,
                          const double *prs, double puny, double *sig1, double *sig2, cudaStream_t s);

struct StrengthArgs {
    const double *aice, *vice, *aice0, *aicen, *vicen; // aicen/vicen: ncat planes, stride `cells`
    const uint8_t *icetmask;
    double *strength;
    int ncat, kstrength, krdg_partic, krdg_redist;
    double mu_rdg, puny, gravit, rhow, rhoi;
};
void aux_ice_strength(const PlaneGeom &pg, const StrengthArgs &a, cudaStream_t s); // ice_mechred.F90:1869-2036

// spin until both neighbours have published at least this rank's epoch (see SubArgs::sync)
void aux_wait_peers(int *sync, int has_north, int has_south, int ncx, cudaStream_t s);

// row_ht[j] = 1 iff on every ocean T cell (tmask) of row j, columns 1..nx+1, the eight metric planes
// equal the init_grid2 formulas applied to HTE/HTN bit for bit (rows 1..nyl+1; others 0)
struct MetricCheckArgs {
    const double *hte, *htn, *dxt, *dyt, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym;
    const uint8_t *tmask;
    uint8_t *row_ht;
};
void aux_check_metrics(const PlaneGeom &pg, const MetricCheckArgs &a, cudaStream_t s);

// Load balance of the subcycle kernel's row chunks by ACTIVE work (device-side replacement of the
// reference's compressed index lists icellt/indxti, source/ice_dyn_evp.F90:850-859): rebuilds the
// chunk table (j0, n) in launch order [south, north, interior south->north] so that every chunk
// carries about the same cost, cost(row) = row_overhead + active T cells of the row; the boundary
// chunks get w_bot / w_top of an interior chunk's cost.  ncy chunks, each >= 1 row (north >= min_top).
// Rows without any active T or U cell are trimmed from the ends of the interior chunks (they belong to
// no chunk); an all-inactive chunk becomes empty.
void aux_balance_chunks(const PlaneGeom &pg, const uint8_t *icetmask, const uint8_t *iceumask, int *rowcnt,
                        int *chunks, int ncy, float w_bot, float w_top, int min_top, float row_overhead,
                        int keep_bot, int keep_top, cudaStream_t s);

// max ice speed / max strength per hemisphere (source/ice_diagnostics.F90:294-346); out4 must be zeroed
void aux_diagnostics(const PlaneGeom &pg, const double *u, const double *v, const double *strength,
                     const double *fcor, double fcor_south, double *out4, cudaStream_t s);

// evp_ieee.cuh (the straight-line IEEE sqrt / division of the subcycle kernel) against sqrt() and operator/
// on n generated operands; see k_selftest_ieee for out[0..5].  Returns a cudaError_t value.
int aux_selftest_ieee(long long n, unsigned long long seed, unsigned long long out[6]);

// ---- strip-tiled layout of the subcycle loop (evp_tiled.cuh) -----------------------------------------
// loop-invariant block, static part (once at init): the 9 T metrics of row j and uarear of U row j-1
struct TileStaticArgs {
    const double *dxt, *dyt, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym, *tinyarea, *uarear;
};
void aux_tile_pack_static(const PlaneGeom &pg, const TileGeom &tg, const TileStaticArgs &a, cudaStream_t s);
// per call, before the ndte loop: strength, the 9 per-call U fields, the masks and the state.  State copy 0
// receives u, v and the stresses of the planes (with the halo words and duplicated slots), copy 1 the same
// velocities and zero stresses (what the plane path does with its second copy, evp_abi.cu).
struct TileCallArgs {
    const double *strength, *aiu, *uocn, *vocn, *waterx, *watery, *forcex, *forcey, *umassdtei, *fm;
    const uint8_t *icetmask, *iceumask;
    const double *u, *v;
    const double *s[EVP_NSTRESS];
};
void aux_tile_pack_call(const PlaneGeom &pg, const TileGeom &tg, const TileCallArgs &a, cudaStream_t s);
// after the loop: state copy `copy` of the tiles -> planes (u, v with their ghost ring; stresses)
struct TileStateArgs {
    double *u, *v;
    double *s[EVP_NSTRESS];
};
void aux_tile_unpack_state(const PlaneGeom &pg, const TileGeom &tg, int copy, const TileStateArgs &a, cudaStream_t s);

// kinetic-energy / volume sums of runtime_diags (source/ice_diagnostics.F90:199-234), fixed summation order
struct EnergyArgs {
    const double *u, *v, *vice, *vsno, *tarea, *fcor;
    const uint8_t *tmask;
    double rhoi, rhos, fcor_south;
    double *rowsum; // 6 * (nyl + 2) doubles of scratch
    double *out6;   // ke north, ke south, ice volume north, south, snow volume north, south
};
void aux_energy_sums(const PlaneGeom &pg, const EnergyArgs &a, cudaStream_t s);

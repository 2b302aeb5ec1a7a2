/*
 * ref_glue.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Host for the machine-translated reference (oracle/f90_to_c.py, oracle/build_ref.py): the
 * reference's own `evp`, `set_evp_parameters`, `evp_prep1/2`, `stress`, `stepu`, `evp_finish`,
 * `principal_stress` (source/ice_dyn_evp.F90), `to_ugrid`, `to_tgrid`, `t2ugrid_vector`,
 * `u2tgrid_vector` (source/ice_grid.F90) and `ice_strength`, `asum_ridging`, `ridge_itd`
 * (source/ice_mechred.F90) are translated statement by statement into REF_GEN (a generated file in a
 * temporary build directory, never committed) and #included below.  What the translated code calls but the
 * translator does not cover is supplied here by hand:
 *
 *   get_block        (source/ice_blocks.F90:349-378)  -> the single block of the test domain
 *   ice_HaloUpdate   (serial/ice_boundary.F90:591-873) -> orc_halo_r8 / orc_halo_i4 of evp_oracle.c
 *                    (index copying only; cross-checked against an independent numpy restatement
 *                    in tests/test_oracle_golden.py)
 *   ice_timer_start/stop -> no-ops
 *
 * ref_evp() / ref_evp_blocks() bind the module variables of ice_state / ice_flux / ice_grid /
 * ice_dyn_evp to the caller's arrays (one block, or a create_blocks decomposition) and call the
 * translated `evp(dt)`.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "evp_oracle.h"

struct f_block {
    int32_t v_ilo, v_ihi, v_jlo, v_jhi;
};
static struct f_block v_get_block(int32_t block_id, int32_t local_id);
static void shim_halo_r8(double *a, int32_t *halo, int32_t *loc, int32_t *kind);
static void shim_halo_i4(int32_t *a, int32_t *halo, int32_t *loc, int32_t *kind);
#define v_ice_haloupdate(f, h, l, t) _Generic((f), double *: shim_halo_r8, int32_t *: shim_halo_i4)(f, h, l, t)
static void v_ice_timer_start(int32_t *t) { (void)t; }
static void v_ice_timer_stop(int32_t *t) { (void)t; }

#include REF_GEN

/* Domain of the current call.  One block: g_grid describes the block itself.  Several blocks
 * (ref_evp_blocks): g_grid carries nx_block/ny_block of a block and the boundary types, g_lay the
 * decomposition (source/ice_blocks.F90:196-222: per block the physical index range inside the block and
 * the global index of its first physical cell), g_glob the whole domain as one padded block. */
typedef struct {
    int32_t nblocks, nx_global, ny_global;
    const int32_t *ilo, *ihi, *jlo, *jhi, *iglob_lo, *jglob_lo;
} ref_layout;

static orc_grid g_grid, g_glob;
static ref_layout g_lay;
static int32_t *g_block_ids;

static struct f_block v_get_block(int32_t block_id, int32_t local_id) {
    struct f_block b = {g_grid.ilo, g_grid.ihi, g_grid.jlo, g_grid.jhi};
    (void)block_id;
    if (g_lay.nblocks > 0) {
        const int k = local_id - 1;
        b.v_ilo = g_lay.ilo[k]; b.v_ihi = g_lay.ihi[k]; b.v_jlo = g_lay.jlo[k]; b.v_jhi = g_lay.jhi[k];
    }
    return b;
}

/* ice_HaloUpdate over several blocks: what the reference's message/copy lists achieve is that every
 * ghost cell holds the value of the neighbouring block's physical cell, or the boundary condition of
 * the domain.  Done here by assembling the domain as one padded block, applying the one-block halo
 * update (orc_halo_*) and handing every block its ring [ilo-1, ihi+1] x [jlo-1, jhi+1] back (the
 * physical cells too: the tripole fold also symmetrises the top physical row). */
#define HALO_BLOCKS(T, HALO, FILL)                                                                          \
    static void halo_blocks_##T(T *a, int loc, int kind) {                                                    \
        const int nxb = g_grid.nx_block, nyb = g_grid.ny_block, nxg = g_glob.nx_block, nyg = g_glob.ny_block; \
        T *glob = (T *)calloc((size_t)nxg * nyg, sizeof(T));                                                  \
        if (!glob) abort();                                                                                   \
        for (int b = 0; b < g_lay.nblocks; ++b)                                                               \
            for (int j = g_lay.jlo[b]; j <= g_lay.jhi[b]; ++j)                                                \
                for (int i = g_lay.ilo[b]; i <= g_lay.ihi[b]; ++i)                                            \
                    glob[(size_t)(g_lay.jglob_lo[b] + j - g_lay.jlo[b]) * nxg + (g_lay.iglob_lo[b] + i - g_lay.ilo[b])] = \
                        a[((size_t)b * nyb + (j - 1)) * nxb + (i - 1)];                                       \
        HALO(glob, &g_glob, loc, kind, FILL);                                                                 \
        for (int b = 0; b < g_lay.nblocks; ++b)                                                               \
            for (int j = g_lay.jlo[b] - 1; j <= g_lay.jhi[b] + 1; ++j)                                        \
                for (int i = g_lay.ilo[b] - 1; i <= g_lay.ihi[b] + 1; ++i)                                    \
                    a[((size_t)b * nyb + (j - 1)) * nxb + (i - 1)] =                                          \
                        glob[(size_t)(g_lay.jglob_lo[b] + j - g_lay.jlo[b]) * nxg + (g_lay.iglob_lo[b] + i - g_lay.ilo[b])]; \
        free(glob);                                                                                           \
    }
HALO_BLOCKS(double, orc_halo_r8, 0.0)
HALO_BLOCKS(int32_t, orc_halo_i4, 0)

/* field_loc_* / field_type_* of drivers/cice4/ice_constants.F90 arrive as the translated parameter
 * values; map them by value onto the oracle's enums */
static int map_loc(int32_t loc) {
    if (loc == v_field_loc_center) return ORC_LOC_CENTER;
    if (loc == v_field_loc_necorner) return ORC_LOC_NECORNER;
    if (loc == v_field_loc_nface) return ORC_LOC_NFACE;
    if (loc == v_field_loc_eface) return ORC_LOC_EFACE;
    abort();
}
static int map_kind(int32_t kind) {
    if (kind == v_field_type_scalar) return ORC_TYPE_SCALAR;
    if (kind == v_field_type_vector) return ORC_TYPE_VECTOR;
    if (kind == v_field_type_angle) return ORC_TYPE_ANGLE;
    abort();
}
static void shim_halo_r8(double *a, int32_t *halo, int32_t *loc, int32_t *kind) {
    (void)halo;
    if (g_lay.nblocks > 0) halo_blocks_double(a, map_loc(*loc), map_kind(*kind));
    else orc_halo_r8(a, &g_grid, map_loc(*loc), map_kind(*kind), 0.0);
}
static void shim_halo_i4(int32_t *a, int32_t *halo, int32_t *loc, int32_t *kind) {
    (void)halo;
    if (g_lay.nblocks > 0) halo_blocks_int32_t(a, map_loc(*loc), map_kind(*kind));
    else orc_halo_i4(a, &g_grid, map_loc(*loc), map_kind(*kind), 0);
}

static int32_t one_block[1] = {1};

/* lay == NULL: one block spanning the domain */
static void bind_domain(const orc_grid *g, const orc_params *p, const ref_layout *lay) {
    g_grid = *g;
    memset(&g_lay, 0, sizeof(g_lay));
    v_nx_block = g->nx_block;
    v_ny_block = g->ny_block;
    v_max_blocks = 1;
    v_nblocks = 1;
    v_blocks_ice_ = one_block;
    if (lay) {
        g_lay = *lay;
        g_glob = *g;
        g_glob.nx_block = lay->nx_global + 2; g_glob.ny_block = lay->ny_global + 2;
        g_glob.ilo = 2; g_glob.ihi = lay->nx_global + 1; g_glob.jlo = 2; g_glob.jhi = lay->ny_global + 1;
        v_max_blocks = v_nblocks = lay->nblocks;
        free(g_block_ids);
        g_block_ids = (int32_t *)malloc(sizeof(int32_t) * lay->nblocks);
        if (!g_block_ids) abort();
        for (int b = 0; b < lay->nblocks; ++b) g_block_ids[b] = b + 1;
        v_blocks_ice_ = g_block_ids;
    }
    v_ncat = p->ncat;
    ref_init_parameters();
}

/* namelist / run-time module variables (source/ice_init.F90:219-222,258-264) */
static void bind_scalars(const orc_params *p) {
    v_ndte = p->ndte;
    v_evp_damping = p->evp_damping;
    v_kstrength = p->kstrength;
    v_krdg_partic = p->krdg_partic;
    v_krdg_redist = p->krdg_redist;
    v_mu_rdg = p->mu_rdg;
#ifdef REF_AUSCOM
    v_dragio = p->dragio;
    v_cosw = p->cosw;
    v_sinw = p->sinw;
    v_use_ocnslope = p->use_ocnslope;
#endif
}

/* set_evp_parameters as the reference computes it: returns the six derived module scalars */
void ref_set_evp_parameters(const orc_params *p, double dt, double *out6) {
    orc_grid g = {3, 3, 2, 2, 2, 2, 0, 0};
    bind_domain(&g, p, NULL);
    bind_scalars(p);
    v_set_evp_parameters(&dt);
    out6[0] = v_dtei; out6[1] = v_ecci; out6[2] = v_dte2t;
    out6[3] = v_denom1; out6[4] = v_denom2; out6[5] = v_rcon;
}

/* the reference's evp(dt); same argument structs as orc_evp, every array (nx_block, ny_block, nblocks)
 * (aicen / vicen: (nx_block, ny_block, ncat, nblocks)).  f->strength_in is ignored (the reference
 * always calls ice_strength). */
static int ref_evp_impl(const orc_grid *g, const ref_layout *lay, const orc_params *p, const orc_fields *f, double dt) {
    const size_t plane = (size_t)g->nx_block * g->ny_block * (lay ? lay->nblocks : 1);
    bind_domain(g, p, lay);
    bind_scalars(p);
    v_set_evp_parameters(&dt); /* init_evp, source/ice_dyn_evp.F90:476 */

    v_work1_ = (double *)calloc(plane, sizeof(double));
    if (!v_work1_) return -1;
    /* ice_grid */
    v_dxt_ = (double *)f->dxt; v_dyt_ = (double *)f->dyt; v_dxhy_ = (double *)f->dxhy; v_dyhx_ = (double *)f->dyhx;
    v_cxp_ = (double *)f->cxp; v_cyp_ = (double *)f->cyp; v_cxm_ = (double *)f->cxm; v_cym_ = (double *)f->cym;
    v_tarea_ = (double *)f->tarea; v_tarear_ = (double *)f->tarear; v_tinyarea_ = (double *)f->tinyarea;
    v_uarea_ = (double *)f->uarea; v_uarear_ = (double *)f->uarear;
    v_tmask_ = (int32_t *)f->tmask; v_umask_ = (int32_t *)f->umask;
    v_fcor_blk_ = (double *)f->fcor;
    /* ice_state */
    v_aice_ = (double *)f->aice; v_vice_ = (double *)f->vice; v_vsno_ = (double *)f->vsno;
    v_aice0_ = (double *)f->aice0; v_aicen_ = (double *)f->aicen; v_vicen_ = (double *)f->vicen;
    v_uvel_ = f->uvel; v_vvel_ = f->vvel; v_strength_ = f->strength;
    v_divu_ = f->divu; v_shear_ = f->shear;
    /* ice_flux */
    v_strairxt_ = (double *)f->strairxT; v_strairyt_ = (double *)f->strairyT;
    v_strax_ = (double *)f->strax; v_stray_ = (double *)f->stray;
    v_uocn_ = (double *)f->uocn; v_vocn_ = (double *)f->vocn;
    v_ss_tltx_ = (double *)f->ss_tltx; v_ss_tlty_ = (double *)f->ss_tlty;
    v_stressp_1_ = f->stressp_1; v_stressp_2_ = f->stressp_2; v_stressp_3_ = f->stressp_3; v_stressp_4_ = f->stressp_4;
    v_stressm_1_ = f->stressm_1; v_stressm_2_ = f->stressm_2; v_stressm_3_ = f->stressm_3; v_stressm_4_ = f->stressm_4;
    v_stress12_1_ = f->stress12_1; v_stress12_2_ = f->stress12_2; v_stress12_3_ = f->stress12_3; v_stress12_4_ = f->stress12_4;
    v_iceumask_ = f->iceumask;
    v_strairx_ = f->strairx; v_strairy_ = f->strairy; v_strtltx_ = f->strtltx; v_strtlty_ = f->strtlty;
    v_strintx_ = f->strintx; v_strinty_ = f->strinty; v_strocnx_ = f->strocnx; v_strocny_ = f->strocny;
    v_strocnxt_ = f->strocnxT; v_strocnyt_ = f->strocnyT; v_fm_ = f->fm; v_prs_sig_ = f->prs_sig;
    v_rdg_conv_ = f->rdg_conv; v_rdg_shear_ = f->rdg_shear;
#ifdef REF_AUSCOM
    v_sicemass_ = f->sicemass;
#endif
    v_evp(&dt);
    free(v_work1_);
    v_work1_ = NULL;
    return 0;
}

/* one block spanning the domain */
int ref_evp(const orc_grid *g, const orc_params *p, const orc_fields *f, double dt) {
    return ref_evp_impl(g, NULL, p, f, dt);
}

/* the domain decomposed into blocks (create_blocks, source/ice_blocks.F90:196-222): g holds nx_block,
 * ny_block of a block and the domain's boundary types */
int ref_evp_blocks(const orc_grid *g, const ref_layout *lay, const orc_params *p, const orc_fields *f, double dt) {
    return ref_evp_impl(g, lay, p, f, dt);
}

/* Timing helper (bench.py cpu_baseline kind "reference"): nsub subcycles of the reference's own
 * stress + stepu + the two velocity halo updates (source/ice_dyn_evp.F90:347-404) on fields that a
 * previous orc_evp / ref_evp call has prepared (icetmask, iceumask, aiu, waterx, ... valid in f).
 * Serial, like the reference's serial build. */
#include <time.h>
int ref_subcycle_only(const orc_grid *g, const orc_params *p, const orc_fields *f, double dt, int nsub,
                      double *seconds) {
    const size_t plane = (size_t)g->nx_block * g->ny_block;
    int32_t nxb = g->nx_block, nyb = g->ny_block, icellt = 0, icellu = 0, ksub;
    int32_t *ix = (int32_t *)malloc(4 * plane * sizeof(int32_t));
    double *str = (double *)malloc(8 * plane * sizeof(double));
    struct timespec t0, t1;
    if (!ix || !str) return -1;
    int32_t *indxti = ix, *indxtj = ix + plane, *indxui = ix + 2 * plane, *indxuj = ix + 3 * plane;
    bind_domain(g, p, NULL);
    bind_scalars(p);
    v_set_evp_parameters(&dt);
    for (int j = g->jlo; j <= g->jhi + 1; ++j) /* the lists evp_prep2 builds, :850-859, :867-880 */
        for (int i = g->ilo; i <= g->ihi + 1; ++i)
            if (f->icetmask[(size_t)(j - 1) * nxb + (i - 1)] == 1) { indxti[icellt] = i; indxtj[icellt] = j; ++icellt; }
    for (int j = g->jlo; j <= g->jhi; ++j)
        for (int i = g->ilo; i <= g->ihi; ++i)
            if (f->iceumask[(size_t)(j - 1) * nxb + (i - 1)]) { indxui[icellu] = i; indxuj[icellu] = j; ++icellu; }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 1; k <= nsub; ++k) {
        ksub = (k == nsub) ? p->ndte : 0;
        v_stress(&nxb, &nyb, &ksub, &icellt, indxti, indxtj, f->uvel, f->vvel, (double *)f->dxt, (double *)f->dyt,
                 (double *)f->dxhy, (double *)f->dyhx, (double *)f->cxp, (double *)f->cyp, (double *)f->cxm,
                 (double *)f->cym, (double *)f->tarear, (double *)f->tinyarea, f->strength, f->stressp_1,
                 f->stressp_2, f->stressp_3, f->stressp_4, f->stressm_1, f->stressm_2, f->stressm_3, f->stressm_4,
                 f->stress12_1, f->stress12_2, f->stress12_3, f->stress12_4, f->shear, f->divu, f->prs_sig,
                 f->rdg_conv, f->rdg_shear, str);
        v_stepu(&nxb, &nyb, &icellu, indxui, indxuj, f->aiu, str, (double *)f->uocn, (double *)f->vocn, f->waterx,
                f->watery, f->forcex, f->forcey, f->umassdtei, f->fm, (double *)f->uarear, f->strocnx, f->strocny,
                f->strintx, f->strinty, f->uvel, f->vvel);
        orc_halo_r8(f->uvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
        orc_halo_r8(f->vvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(ix);
    free(str);
    return 0;
}

/* principal_stress (source/ice_dyn_evp.F90:1558-1609) on caller arrays */
void ref_principal_stress(const orc_params *p, int32_t nx_block, int32_t ny_block, double *stressp_1,
                          double *stressm_1, double *stress12_1, double *prs_sig, double *sig1, double *sig2) {
    orc_grid g = {nx_block, ny_block, 2, nx_block - 1, 2, ny_block - 1, 0, 0};
    bind_domain(&g, p, NULL);
    v_principal_stress(&nx_block, &ny_block, stressp_1, stressm_1, stress12_1, prs_sig, sig1, sig2);
}

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
 *   allocation        the heads of create_blocks (source/ice_blocks.F90:163-195) and ice_HaloCreate
 *                     (serial/ice_boundary.F90:160-215,433-460): array sizes; the capacity of the local-copy
 *                     lists is a bound per block instead of the reference's counting pass
 *   get_block, get_block_parameter (source/ice_blocks.F90:349-378,880-938): reads of the translated all_blocks
 *   abort_ice         records the message; the host returns an error
 *   ice_timer_start/stop -> no-ops
 *
 * Translated as well, since round 2 (oracle/build_ref.py translate_halo): the block loop of create_blocks,
 * ice_blocksGetNbrID, ice_distributionGetBlockLoc, the message-configuration loop of ice_HaloCreate,
 * ice_HaloMsgCreate, ice_HaloUpdate2DR8 and ice_HaloUpdate2DI4 with the derived types block, distrb and ice_halo --
 * every halo update of the translated evp runs the reference's OWN address lists and copy / tripole-fold code, on one
 * block and on create_blocks decompositions (eliminated land blocks = blocks without a task) alike.
 *
 * ref_evp() / ref_evp_blocks() bind the module variables of ice_state / ice_flux / ice_grid /
 * ice_dyn_evp to the caller's arrays (one block, or a create_blocks decomposition) and call the
 * translated `evp(dt)`.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "evp_oracle.h"

/* hosts of what the translated code calls but the translator does not cover (see the header comment) */
struct f_block;
struct f_distrb;
struct f_ice_halo;
struct f_kw_get_block_parameter { /* keyword arguments of get_block_parameter (source/ice_blocks.F90:880-938) */
    int32_t *v_local_id, *v_ilo, *v_ihi, *v_jlo, *v_jhi, *v_iblock, *v_jblock, *v_tripole;
    int32_t **v_i_glob, **v_j_glob;
};
static struct f_block v_get_block(int32_t block_id, int32_t local_id);
static void v_get_block_parameter(int32_t *block_id, struct f_kw_get_block_parameter kw);
static void v_abort_ice(const char *msg);
static void shim_halo_r8(double *a, int32_t *halo, int32_t *loc, int32_t *kind);
static void shim_halo_i4(int32_t *a, int32_t *halo, int32_t *loc, int32_t *kind);
#define v_ice_haloupdate(f, h, l, t) _Generic((f), double *: shim_halo_r8, int32_t *: shim_halo_i4)(f, h, l, t)
static void v_ice_timer_start(int32_t *t) { (void)t; }
static void v_ice_timer_stop(int32_t *t) { (void)t; }

#include REF_GEN

#include <stdio.h>

/* Domain of the current call.  g_grid: nx_block / ny_block of a block and the boundary types of the domain.
 * g_lay (several blocks, ref_evp_blocks): the caller's decomposition -- per LOCAL block the physical index range
 * inside the block and the global index of its first physical cell; blocks of the full create_blocks grid that the
 * caller does not hold are eliminated land blocks (no task: source/ice_distribution.F90). */
typedef struct {
    int32_t nblocks, nx_global, ny_global;
    const int32_t *ilo, *ihi, *jlo, *jhi, *iglob_lo, *jglob_lo;
} ref_layout;

static orc_grid g_grid;
static ref_layout g_lay;
static int32_t *g_block_ids;          /* blocks_ice: local block -> global block id */
static struct f_distrb g_dist;        /* the (serial) distribution: task 1 or 0 (eliminated) per global block */
static struct f_ice_halo g_halo;      /* halo_info of source/ice_domain.F90, built per domain */
static char g_abort_msg[256];
static int g_aborted;

static void v_abort_ice(const char *msg) { /* the reference stops the model; the host reports the message */
    if (!g_aborted) snprintf(g_abort_msg, sizeof g_abort_msg, "%s", msg);
    g_aborted = 1;
}
const char *ref_abort_message(void) { return g_aborted ? g_abort_msg : ""; }

/* source/ice_blocks.F90:349-378 */
static struct f_block v_get_block(int32_t block_id, int32_t local_id) {
    (void)local_id;
    return v_all_blocks(block_id);
}

/* source/ice_blocks.F90:880-938: every result is optional; i_glob / j_glob are pointer results */
static void v_get_block_parameter(int32_t *block_id, struct f_kw_get_block_parameter kw) {
    if (*block_id < 1 || *block_id > v_nblocks_tot) {
        v_abort_ice("ice: get_block_parameter: invalid block_id");
        return;
    }
    const struct f_block *b = &v_all_blocks(*block_id);
    if (kw.v_local_id) *kw.v_local_id = b->v_local_id;
    if (kw.v_ilo) *kw.v_ilo = b->v_ilo;
    if (kw.v_ihi) *kw.v_ihi = b->v_ihi;
    if (kw.v_jlo) *kw.v_jlo = b->v_jlo;
    if (kw.v_jhi) *kw.v_jhi = b->v_jhi;
    if (kw.v_iblock) *kw.v_iblock = b->v_iblock;
    if (kw.v_jblock) *kw.v_jblock = b->v_jblock;
    if (kw.v_i_glob) *kw.v_i_glob = b->v_i_glob_;
    if (kw.v_j_glob) *kw.v_j_glob = b->v_j_glob_;
    if (kw.v_tripole) *kw.v_tripole = b->v_tripole;
}

static const char *bnd_name(int b) { /* the translator lower-cases the reference's text, string literals included */
    switch (b) {
    case ORC_BND_OPEN: return "open";
    case ORC_BND_CLOSED: return "closed";
    case ORC_BND_CYCLIC: return "cyclic";
    case ORC_BND_TRIPOLE: return "tripole";
    case ORC_BND_TRIPOLET: return "tripolet";
    }
    abort();
}

static void free_decomposition(void) {
    free(v_all_blocks_); free(v_all_blocks_ij_); free(v_i_global_); free(v_j_global_);
    free(g_dist.v_blocklocation_); free(g_dist.v_blocklocalid_); free(g_dist.v_blockglobalid_);
    free(g_halo.v_srclocaladdr_); free(g_halo.v_dstlocaladdr_);
    free(v_buftripoler8_); free(v_buftripolei4_);
    free(g_block_ids);
    v_all_blocks_ = NULL; v_all_blocks_ij_ = NULL; v_i_global_ = NULL; v_j_global_ = NULL;
    memset(&g_dist, 0, sizeof g_dist);
    memset(&g_halo, 0, sizeof g_halo);
    v_buftripoler8_ = NULL; v_buftripolei4_ = NULL;
    g_block_ids = NULL;
}

/* The reference's own decomposition and halo structure for the domain of this call: allocation as in the heads of
 * create_blocks (source/ice_blocks.F90:163-195) and ice_HaloCreate (serial/ice_boundary.F90:160-215, 433-460), then
 * the TRANSLATED block loop of create_blocks, the distribution (one task; blocks the caller does not hold have no
 * task), and the TRANSLATED message-configuration loop of ice_HaloCreate with ice_HaloMsgCreate.  The number of
 * local copies is bounded per block by its ring + the tripole rows instead of the reference's counting pass.
 * Returns 0, or -1 when the caller's layout is not what create_blocks makes of this domain. */
static int build_decomposition(const orc_grid *g, const ref_layout *lay) {
    int32_t nxg = lay ? lay->nx_global : g->nx_block - 2, nyg = lay ? lay->ny_global : g->ny_block - 2;
    const char *ew = bnd_name(g->ew_boundary), *ns = bnd_name(g->ns_boundary);
    free_decomposition();
    g_aborted = 0;
    v_my_task = 0;
    v_block_size_x = g->nx_block - 2 * v_nghost;
    v_block_size_y = g->ny_block - 2 * v_nghost;
    v_nblocks_x = (nxg - 1) / v_block_size_x + 1;
    v_nblocks_y = (nyg - 1) / v_block_size_y + 1;
    v_nblocks_tot = v_nblocks_x * v_nblocks_y;
    const size_t nt = (size_t)v_nblocks_tot;
    v_all_blocks_ = (struct f_block *)calloc(nt, sizeof(struct f_block));
    v_all_blocks_ij_ = (int32_t *)calloc(nt, sizeof(int32_t));
    v_i_global_ = (int32_t *)calloc(nt * g->nx_block, sizeof(int32_t));
    v_j_global_ = (int32_t *)calloc(nt * g->ny_block, sizeof(int32_t));
    g_dist.v_blocklocation_ = (int32_t *)calloc(nt, sizeof(int32_t));
    g_dist.v_blocklocalid_ = (int32_t *)calloc(nt, sizeof(int32_t));
    g_dist.v_blockglobalid_ = (int32_t *)calloc(nt, sizeof(int32_t));
    const int nlocal = lay ? lay->nblocks : 1;
    g_block_ids = (int32_t *)calloc((size_t)nlocal, sizeof(int32_t));
    if (!v_all_blocks_ || !v_all_blocks_ij_ || !v_i_global_ || !v_j_global_ || !g_dist.v_blocklocation_ ||
        !g_dist.v_blocklocalid_ || !g_dist.v_blockglobalid_ || !g_block_ids)
        abort();
    v_create_blocks_loop(&nxg, &nyg, ew, ns);
    if (g_aborted) return -1;
    /* distribution: the caller's blocks, in the caller's order, on task 1 */
    g_dist.v_nprocs = 1;
    g_dist.v_numlocalblocks = nlocal;
    for (int k = 0; k < nlocal; ++k) {
        const int ig = lay ? lay->iglob_lo[k] : 1, jg = lay ? lay->jglob_lo[k] : 1;
        if ((ig - 1) % v_block_size_x || (jg - 1) % v_block_size_y) return -1;
        const int ib = (ig - 1) / v_block_size_x + 1, jb = (jg - 1) / v_block_size_y + 1;
        if (ib < 1 || ib > v_nblocks_x || jb < 1 || jb > v_nblocks_y) return -1;
        const int n = v_all_blocks_ij(ib, jb);
        const struct f_block *b = &v_all_blocks(n);
        const int ilo = lay ? lay->ilo[k] : g->ilo, ihi = lay ? lay->ihi[k] : g->ihi;
        const int jlo = lay ? lay->jlo[k] : g->jlo, jhi = lay ? lay->jhi[k] : g->jhi;
        if (b->v_ilo != ilo || b->v_ihi != ihi || b->v_jlo != jlo || b->v_jhi != jhi) {
            /* e.g. a padded edge block of exactly ONE physical column / row: create_blocks shrinks ihi only for
             * i > ilo (source/ice_blocks.F90:332-335), so the reference leaves such a block at full width */
            snprintf(g_abort_msg, sizeof g_abort_msg,
                     "block %d: the caller's layout (ilo..ihi %d..%d, jlo..jhi %d..%d) is not what create_blocks makes "
                     "(%d..%d, %d..%d)", n, ilo, ihi, jlo, jhi, b->v_ilo, b->v_ihi, b->v_jlo, b->v_jhi);
            g_aborted = 1;
            return -1;
        }
        if (g_dist.v_blocklocation(n) != 0) return -1; /* the same block twice */
        g_dist.v_blocklocation(n) = 1;
        g_dist.v_blocklocalid(n) = k + 1;
        g_dist.v_blockglobalid(k + 1) = n;
        v_all_blocks(n).v_local_id = k + 1;
        g_block_ids[k] = n;
    }
    /* halo structure */
    const int tripole = g->ns_boundary == ORC_BND_TRIPOLE || g->ns_boundary == ORC_BND_TRIPOLET;
    g_halo.v_tripoletflag = g->ns_boundary == ORC_BND_TRIPOLET;
    g_halo.v_tripolerows = v_nghost + 1 + (g->ns_boundary == ORC_BND_TRIPOLET ? 1 : 0);
    if (tripole) {
        v_buf_nx = nxg;
        v_buf_rows = g_halo.v_tripolerows;
        v_buftripoler8_ = (double *)calloc((size_t)nxg * v_buf_rows, sizeof(double));
        v_buftripolei4_ = (int32_t *)calloc((size_t)nxg * v_buf_rows, sizeof(int32_t));
        if (!v_buftripoler8_ || !v_buftripolei4_) abort();
    }
    const size_t cap = nt * (16 * (size_t)(g->nx_block + g->ny_block) + 16);
    g_halo.v_srclocaladdr_ = (int32_t *)calloc(3 * cap, sizeof(int32_t));
    g_halo.v_dstlocaladdr_ = (int32_t *)calloc(3 * cap, sizeof(int32_t));
    if (!g_halo.v_srclocaladdr_ || !g_halo.v_dstlocaladdr_) abort();
    g_halo.v_numlocalcopies = 0;
    v_ice_halocreate_msgconfig(&g_halo, &g_dist, ns, ew);
    if (g_aborted || (size_t)g_halo.v_numlocalcopies > cap) return -1;
    return 0;
}

/* ice_HaloUpdate (generic interface, serial/ice_boundary.F90:71-81) -> the translated specific routines on the
 * halo structure built above; fillValue is absent in every call of the path */
static double *g_strength_pre;
static size_t g_strength_pre_n;
static void shim_halo_r8(double *a, int32_t *halo, int32_t *loc, int32_t *kind) {
    (void)halo;
    if (a == v_strength_) { /* tests: ice_strength's result before evp halo-updates it (:337-343) */
        const size_t n = (size_t)v_nx_block * v_ny_block * v_max_blocks;
        free(g_strength_pre);
        g_strength_pre = (double *)malloc(sizeof(double) * n);
        g_strength_pre_n = g_strength_pre ? n : 0;
        if (g_strength_pre) memcpy(g_strength_pre, a, sizeof(double) * n);
    }
    v_ice_haloupdate2dr8(a, &g_halo, loc, kind, NULL);
}
size_t ref_last_strength_prehalo(double *out) {
    if (out && g_strength_pre) memcpy(out, g_strength_pre, sizeof(double) * g_strength_pre_n);
    return g_strength_pre_n;
}
static void shim_halo_i4(int32_t *a, int32_t *halo, int32_t *loc, int32_t *kind) {
    (void)halo;
    v_ice_haloupdate2di4(a, &g_halo, loc, kind, NULL);
}

/* for tests: the reference's halo update on a caller array (nx_block, ny_block, nblocks) of the bound domain */
int ref_halo_update_r8(const orc_grid *g, const ref_layout *lay, double *a, int32_t loc, int32_t kind);
int ref_halo_update_i4(const orc_grid *g, const ref_layout *lay, int32_t *a, int32_t loc, int32_t kind);

/* lay == NULL: one block spanning the domain.  Returns 0 or -1 (layout / decomposition refused). */
static int bind_domain(const orc_grid *g, const orc_params *p, const ref_layout *lay) {
    g_grid = *g;
    memset(&g_lay, 0, sizeof(g_lay));
    if (lay) g_lay = *lay;
    v_nx_block = g->nx_block;
    v_ny_block = g->ny_block;
    v_max_blocks = v_nblocks = lay ? lay->nblocks : 1;
    if (p) v_ncat = p->ncat;
    ref_init_parameters();
    if (build_decomposition(g, lay)) return -1;
    v_blocks_ice_ = g_block_ids;
    return 0;
}

int ref_halo_update_r8(const orc_grid *g, const ref_layout *lay, double *a, int32_t loc, int32_t kind) {
    if (bind_domain(g, NULL, lay)) return -1;
    int32_t l = loc == ORC_LOC_CENTER ? v_field_loc_center : loc == ORC_LOC_NECORNER ? v_field_loc_necorner
              : loc == ORC_LOC_NFACE ? v_field_loc_nface : v_field_loc_eface;
    int32_t k = kind == ORC_TYPE_SCALAR ? v_field_type_scalar : kind == ORC_TYPE_VECTOR ? v_field_type_vector
              : v_field_type_angle;
    v_ice_haloupdate2dr8(a, &g_halo, &l, &k, NULL);
    return g_aborted ? -1 : 0;
}
int ref_halo_update_i4(const orc_grid *g, const ref_layout *lay, int32_t *a, int32_t loc, int32_t kind) {
    if (bind_domain(g, NULL, lay)) return -1;
    int32_t l = loc == ORC_LOC_CENTER ? v_field_loc_center : loc == ORC_LOC_NECORNER ? v_field_loc_necorner
              : loc == ORC_LOC_NFACE ? v_field_loc_nface : v_field_loc_eface;
    int32_t k = kind == ORC_TYPE_SCALAR ? v_field_type_scalar : kind == ORC_TYPE_VECTOR ? v_field_type_vector
              : v_field_type_angle;
    v_ice_haloupdate2di4(a, &g_halo, &l, &k, NULL);
    return g_aborted ? -1 : 0;
}
/* the halo structure of the bound domain (tests): number of local copies and the address lists */
int32_t ref_halo_num_copies(void) { return g_halo.v_numlocalcopies; }
const int32_t *ref_halo_src(void) { return g_halo.v_srclocaladdr_; }
const int32_t *ref_halo_dst(void) { return g_halo.v_dstlocaladdr_; }

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
    if (bind_domain(&g, p, NULL)) abort();
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
    if (bind_domain(g, p, lay)) return -2;
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
    return g_aborted ? -3 : 0;
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
    if (bind_domain(g, p, NULL)) return -2;
    bind_scalars(p);
    v_set_evp_parameters(&dt);
    int32_t loc = v_field_loc_necorner, kind = v_field_type_vector;
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
        v_ice_haloupdate2dr8(f->uvel, &g_halo, &loc, &kind, NULL); /* the reference's own halo update */
        v_ice_haloupdate2dr8(f->vvel, &g_halo, &loc, &kind, NULL);
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
    if (bind_domain(&g, p, NULL)) abort();
    v_principal_stress(&nx_block, &ny_block, stressp_1, stressm_1, stress12_1, prs_sig, sig1, sig2);
}

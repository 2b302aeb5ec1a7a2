/*
 * evp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64) of the EVP sea-ice dynamics path of
 * COSIMA/cice4, used as the parity checker for the CUDA implementation and as
 * the "port" CPU baseline in bench.py.  Nothing in the product
 * (cice4_b200/, include/) may call, link or import anything in oracle/.
 *
 * PARITY PINNED against the reference itself.  The reference is Fortran and this image has no
 * Fortran compiler, so the reference's own source text of `evp` and everything it calls is
 * machine-translated to C at build time (oracle/f90_to_c.py, oracle/build_ref.py ->
 * oracle/_ref/libevp_ref_<variant>.so, git-ignored) and this restatement is required to agree with it
 * BIT FOR BIT: directly where /root/reference exists (tests/test_oracle_vs_ref.py: 7 domain types x
 * 5 CPP variants, namelist options, dt/ndte sweeps) and everywhere through the committed outputs of
 * that library (tests/golden/ref_evp_*.npz, tests/test_oracle_golden.py).  Hand-written in the
 * translated reference are only get_block and the index copying of ice_HaloUpdate (oracle/ref_glue.c,
 * which reuses orc_halo_* below; cross-checked against an independent numpy restatement).  The
 * reference's known answers are pinned too: grid decoding and set_evp_parameters
 * (ice.log.Linux.LANL.coyote:101-119,181-183) and the ITD bounds (:185-190).
 *
 * Build: -O2 -ffp-contract=off (strict: no FMA contraction) for parity;
 *        -O3 -march=native -fopenmp for the timed CPU baseline.
 *
 * Layout: one block with a 1-cell ghost ring, Fortran order (i fastest):
 * a[(j-1)*nx_block + (i-1)] is Fortran a(i,j).  A global grid nx x ny held in
 * one block has nx_block = nx+2, ilo=2, ihi=nx+1 (source/ice_blocks.F90:56-61,
 * :219-222).
 */
#ifndef EVP_ORACLE_H
#define EVP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boundary types, source/ice_blocks.F90:237-343 */
enum { ORC_BND_OPEN = 0, ORC_BND_CLOSED = 1, ORC_BND_CYCLIC = 2, ORC_BND_TRIPOLE = 3, ORC_BND_TRIPOLET = 4 };
/* field locations / kinds, drivers/cice4/ice_constants.F90 (field_loc_*, field_type_*) */
enum { ORC_LOC_CENTER = 1, ORC_LOC_NECORNER = 2, ORC_LOC_NFACE = 3, ORC_LOC_EFACE = 4 };
enum { ORC_TYPE_SCALAR = 1, ORC_TYPE_VECTOR = 2, ORC_TYPE_ANGLE = 3 };

typedef struct {
    int32_t nx_block, ny_block;      /* padded block size */
    int32_t ilo, ihi, jlo, jhi;      /* physical range, 1-based */
    int32_t ew_boundary, ns_boundary;
} orc_grid;

/* scalars of module ice_dyn_evp + constants it uses */
typedef struct {
    /* set by orc_set_evp_parameters (source/ice_dyn_evp.F90:535-577) */
    double dtei, ecci, dte2T, denom1, denom2, rcon;
    int32_t ndte;
    int32_t evp_damping;
    /* constants / namelist (drivers/cice4/ice_constants.F90:50-60,65-66) */
    double rhoi, rhos, rhow, dragio, gravit, puny;
    double cosw, sinw;
    /* CPP variants as runtime flags */
    int32_t auscom;        /* #ifdef AusCOM: hemisphere turning, sicemass */
    int32_t coupled;       /* #ifdef coupled: tilt from ss_tlt */
    int32_t use_ocnslope;  /* AusCOM namelist; .false. reverts to geostrophic tilt */
    int32_t access_wind;   /* #ifdef ACCESS: strairx = strax */
    /* ice_strength options (source/ice_init.F90:219-222) */
    int32_t kstrength, krdg_partic, krdg_redist;
    double mu_rdg;
    int32_t ncat;
} orc_params;

/* all fields are (nx_block, ny_block) fp64 unless noted */
typedef struct {
    /* static grid (source/ice_grid.F90:58-123) */
    const double *dxt, *dyt, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym;
    const double *tarea, *tarear, *tinyarea, *uarea, *uarear, *fcor;
    const int32_t *tmask, *umask;
    /* inputs */
    const double *aice, *vice, *vsno, *strairxT, *strairyT, *strax, *stray;
    const double *uocn, *vocn, *ss_tltx, *ss_tlty;
    const double *aice0, *aicen, *vicen;   /* aicen/vicen: (nx_block,ny_block,ncat) */
    const double *strength_in;             /* if non-NULL, used instead of ice_strength */
    /* state, in/out */
    double *uvel, *vvel;
    double *stressp_1, *stressp_2, *stressp_3, *stressp_4;
    double *stressm_1, *stressm_2, *stressm_3, *stressm_4;
    double *stress12_1, *stress12_2, *stress12_3, *stress12_4;
    int32_t *iceumask;
    /* outputs */
    double *strength;
    double *strairx, *strairy, *strtltx, *strtlty, *strintx, *strinty;
    double *strocnx, *strocny, *strocnxT, *strocnyT, *fm, *prs_sig;
    double *divu, *shear, *rdg_conv, *rdg_shear;
    double *sicemass;                      /* AusCOM only, may be NULL */
    int32_t *icetmask;
    /* scratch (locals of evp, source/ice_dyn_evp.F90:160-184) exposed for tests */
    double *tmass, *umass, *aiu, *umassdtei, *waterx, *watery, *forcex, *forcey;
} orc_fields;

void orc_default_params(orc_params *p);
void orc_set_evp_parameters(orc_params *p, double dt, int ndte);

void orc_halo_r8(double *a, const orc_grid *g, int loc, int kind, double fill);
void orc_halo_i4(int32_t *a, const orc_grid *g, int loc, int kind, int32_t fill);

void orc_to_ugrid(const orc_grid *g, const double *tarea, const double *uarea,
                  const double *work1, double *work2);
void orc_to_tgrid(const orc_grid *g, const double *tarea, const double *uarea,
                  const double *work1, double *work2);
void orc_t2ugrid_vector(const orc_grid *g, const double *tarea, const double *uarea,
                        double *work);
void orc_u2tgrid_vector(const orc_grid *g, const double *tarea, const double *uarea,
                        double *work);

void orc_evp_prep1(const orc_grid *g, const orc_params *p, const orc_fields *f);
void orc_evp_prep2(const orc_grid *g, const orc_params *p, const orc_fields *f,
                   int32_t *icellt, int32_t *icellu, int32_t *indxti, int32_t *indxtj,
                   int32_t *indxui, int32_t *indxuj);
void orc_ice_strength(const orc_grid *g, const orc_params *p, const orc_fields *f,
                      int32_t icellt, const int32_t *indxti, const int32_t *indxtj);
void orc_stress(const orc_grid *g, const orc_params *p, const orc_fields *f, int ksub,
                int32_t icellt, const int32_t *indxti, const int32_t *indxtj, double *str);
void orc_stepu(const orc_grid *g, const orc_params *p, const orc_fields *f,
               int32_t icellu, const int32_t *indxui, const int32_t *indxuj, const double *str);
void orc_evp_finish(const orc_grid *g, const orc_params *p, const orc_fields *f,
                    int32_t icellu, const int32_t *indxui, const int32_t *indxuj);
void orc_principal_stress(int nx_block, int ny_block, const double *stressp_1,
                          const double *stressm_1, const double *stress12_1,
                          const double *prs_sig, double puny, double *sig1, double *sig2);

/* full driver, source/ice_dyn_evp.F90:119-432.  Returns 0, or -1 on allocation failure.
 * subcycle_seconds (may be NULL) receives the wall time of the ndte loop only. */
int orc_evp(const orc_grid *g, const orc_params *p, const orc_fields *f, double *subcycle_seconds);
/* tests: the strength array of the last orc_evp call before evp's halo update of it; returns its size */
size_t orc_last_strength_prehalo(double *out);

/* ndte subcycles of stress+stepu+halo only, on prepared fields; for CPU timing.
 * nthreads>1 splits the T/U lists into j-bands (OpenMP) with a barrier per phase. */
int orc_subcycle_only(const orc_grid *g, const orc_params *p, const orc_fields *f, int nsub,
                      double *seconds);

#ifdef __cplusplus
}
#endif
#endif

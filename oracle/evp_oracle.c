/*
 * evp_oracle.c -- TEST INFRASTRUCTURE ONLY (see evp_oracle.h).
 *
 * Statement-by-statement CPU restatement of the EVP dynamics path of
 * COSIMA/cice4.  Every function cites the reference lines it follows; the
 * order of floating-point operations is the Fortran's (left-to-right
 * evaluation, same parenthesisation).  Build with -ffp-contract=off so the
 * compiler does not fuse multiply-adds.
 *
 * Pinned bit for bit against the machine-translated reference (oracle/_ref)
 * -- see the header.
 */
#include "evp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* Fortran (i,j), 1-based, column-major */
#define IX(i, j) ((size_t)((j) - 1) * (size_t)nxb + (size_t)((i) - 1))

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ */
/* constants: drivers/cice4/ice_constants.F90:50-60,65-66,166-176      */
/* ------------------------------------------------------------------ */
static const double c0 = 0.0, c1 = 1.0, c2 = 2.0, c4 = 4.0;
static const double p5 = 0.5, p25 = 0.25;
#define P166 (1.0 / 6.0)
#define P333 (1.0 / 3.0)
#define P111 (1.0 / 9.0)
#define P055 (P111 * 0.5)
#define P027 (P055 * 0.5)
#define P222 (2.0 / 9.0)

void orc_default_params(orc_params *p) {
    memset(p, 0, sizeof(*p));
    p->rhoi = 917.0;
    p->rhos = 330.0;
    p->rhow = 1026.0;
    p->dragio = 0.00536;
    p->gravit = 9.80616;
    p->puny = 1.0e-11;
    p->cosw = 1.0; /* source/ice_dyn_evp.F90:84-85 */
    p->sinw = 0.0;
    p->ndte = 120; /* source/ice_init.F90:216-222 */
    p->evp_damping = 0;
    p->kstrength = 1;
    p->krdg_partic = 1;
    p->krdg_redist = 1;
    p->mu_rdg = 3.0;
    p->ncat = 5;
    p->use_ocnslope = 0;
}

/* source/ice_dyn_evp.F90:535-577 */
void orc_set_evp_parameters(orc_params *p, double dt, int ndte) {
    const double eyc = 0.36; /* :81 */
    double dte, ecc, tdamp2;
    p->ndte = ndte;
    dte = dt / (double)ndte;      /* :563 */
    p->dtei = c1 / dte;           /* :564 */
    ecc = c4;                     /* :567 */
    p->ecci = p25;                /* :568 */
    tdamp2 = c2 * eyc * dt;       /* :571 */
    p->dte2T = dte / tdamp2;      /* :572 */
    p->denom1 = c1 / (c1 + p->dte2T);       /* :573 */
    p->denom2 = c1 / (c1 + p->dte2T * ecc); /* :574 */
    p->rcon = 1230.0 * eyc * dt * (p->dtei * p->dtei); /* :575, dtei**2 */
}

/* ------------------------------------------------------------------ */
/* Halo update for one block that holds the whole global domain.        */
/* Follows serial/ice_boundary.F90:591-873 (ice_HaloUpdate2DR8) with the */
/* address lists of ice_HaloMsgCreate (:3494-4202) and the ghost index   */
/* maps of source/ice_blocks.F90:237-343 collapsed for nblocks = 1:      */
/*   - east/west neighbour exists only for 'cyclic' (the block itself);  */
/*     copies cover physical rows jlo..jhi only (:3629-3668);            */
/*   - north/south neighbour exists only for 'cyclic'; copies cover      */
/*     physical columns only (:3681-3695,:3774-3788);                    */
/*   - corner neighbours exist only when both directions are cyclic      */
/*     (1x1 copies, :3797-3817 ...), or via the tripole buffer;          */
/*   - open/closed: ice_blocksGetNbrID returns 0, ice_HaloMsgCreate      */
/*     returns at :3580, ghost cells are left untouched;                 */
/*   - tripole (u-fold): :3699-3733 copy-in of the top nghost+1 physical */
/*     rows, :777-827 symmetrisation, :3735-3763 + :837-866 copy-out     */
/*     over local i = 1..ihi+nghost, rows jhi and jhi+1;                 */
/*   - tripoleT (T-fold): tripoleRows = nghost+2 (:199-205), so the top  */
/*     THREE physical rows go into the buffer, and the corner messages   */
/*     then overwrite buffer rows 1-2 with rows jhi-1, jhi (see below);  */
/*     symmetrisation and offsets of :725-773 (integer fields: nint of   */
/*     the average, :1303-1321); the copy-out message of buffer row 3    */
/*     has jDst = -1 and is skipped (:3753-3757).                        */
/* All regular copies read physical cells and write ghost cells, so      */
/* their order is immaterial; the tripole copy-out runs last (:837).     */
/* ------------------------------------------------------------------ */
#define HALO_BODY(T)                                                                        \
    const int nxb = g->nx_block;                                                            \
    const int ilo = g->ilo, ihi = g->ihi, jlo = g->jlo, jhi = g->jhi;                       \
    const int nxg = ihi - ilo + 1;                                                          \
    const int ewc = (g->ew_boundary == ORC_BND_CYCLIC);                                     \
    const int nsc = (g->ns_boundary == ORC_BND_CYCLIC);                                     \
    const int tfold = (g->ns_boundary == ORC_BND_TRIPOLET);                                 \
    const int trip = (g->ns_boundary == ORC_BND_TRIPOLE) || tfold;                          \
    const int trows = tfold ? 3 : 2; /* tripoleRows, :199-205 */                            \
    T *buf = NULL;                                                                          \
    int i, j;                                                                               \
    if (trip) {                                                                             \
        /* copy-in: buf(i_glob, 1:trows) = array(:, jhi-trows+1:jhi), :3717-3731 */         \
        buf = (T *)malloc(sizeof(T) * (size_t)nxg * trows);                                 \
        for (j = 1; j <= trows; ++j)                                                        \
            for (i = 1; i <= nxg; ++i)                                                      \
                buf[(size_t)(j - 1) * nxg + (i - 1)] = a[IX(ilo + i - 1, jhi - trows + j)]; \
        /* The 'northeast' / 'northwest' messages of a tripole block (:3833-3848, :3868-3883) copy the top \
         * nghost+1 rows into buffer rows 1 .. nghost+1 -- whatever tripoleRows is -- and they are the last \
         * copies into the buffer in ice_HaloCreate's list order (:519-580).  On the u-fold that repeats the \
         * 'north' message; on the T-fold (3 buffer rows) it OVERWRITES rows 1 and 2, so the buffer the update \
         * works on holds the physical rows jhi-1, jhi, jhi (found by running the reference's own translated \
         * halo, tests/test_oracle_vs_ref.py). */                                             \
        for (j = 1; j <= 2; ++j)                                                            \
            for (i = 1; i <= nxg; ++i)                                                      \
                buf[(size_t)(j - 1) * nxg + (i - 1)] = a[IX(ilo + i - 1, jhi - 2 + j)];     \
    }                                                                                       \
    if (ewc) {                                                                              \
        for (j = jlo; j <= jhi; ++j) {                                                      \
            a[IX(ilo - 1, j)] = a[IX(ihi, j)]; /* 'east' msg: src east edge -> dst west halo */ \
            a[IX(ihi + 1, j)] = a[IX(ilo, j)]; /* 'west' msg */                             \
        }                                                                                   \
    }                                                                                       \
    if (nsc) {                                                                              \
        for (i = ilo; i <= ihi; ++i) {                                                      \
            a[IX(i, jlo - 1)] = a[IX(i, jhi)]; /* 'north' msg: src top -> dst south halo */ \
            a[IX(i, jhi + 1)] = a[IX(i, jlo)]; /* 'south' msg */                            \
        }                                                                                   \
    }                                                                                       \
    if (ewc && nsc) {                                                                       \
        a[IX(ilo - 1, jlo - 1)] = a[IX(ihi, jhi)]; /* northeast */                          \
        a[IX(ihi + 1, jlo - 1)] = a[IX(ilo, jhi)]; /* northwest */                          \
        a[IX(ilo - 1, jhi + 1)] = a[IX(ihi, jlo)]; /* southeast */                          \
        a[IX(ihi + 1, jhi + 1)] = a[IX(ilo, jlo)]; /* southwest */                          \
    }                                                                                       \
    if (trip) {                                                                             \
        int isign = (kind == ORC_TYPE_SCALAR) ? 1 : -1; /* :713-723 */                      \
        int ioffset = 0, joffset = 0;                                                       \
        T *row2 = buf + (size_t)(trows - 1) * nxg; /* bufTripole(:, tripoleRows) */         \
        if (tfold) switch (loc) { /* T-fold branch, :725-773 */                             \
        case ORC_LOC_CENTER:                                                                \
            ioffset = -1; joffset = 0;                                                      \
            for (i = 2; i <= nxg / 2; ++i) {                                                \
                int iDst = nxg - i + 2;                                                     \
                T x1 = row2[i - 1], x2 = row2[iDst - 1];                                    \
                T xavg = (T)AVG(x1, isign * x2);                                            \
                row2[i - 1] = xavg;                                                         \
                row2[iDst - 1] = isign * xavg;                                              \
            }                                                                               \
            break;                                                                          \
        case ORC_LOC_NECORNER: ioffset = 0; joffset = 1; break;                             \
        case ORC_LOC_EFACE:                                                                 \
            ioffset = 0; joffset = 0;                                                       \
            for (i = 1; i <= nxg / 2; ++i) {                                                \
                int iDst = nxg + 1 - i;                                                     \
                T x1 = row2[i - 1], x2 = row2[iDst - 1];                                    \
                T xavg = (T)AVG(x1, isign * x2);                                            \
                row2[i - 1] = xavg;                                                         \
                row2[iDst - 1] = isign * xavg;                                              \
            }                                                                               \
            break;                                                                          \
        case ORC_LOC_NFACE: ioffset = -1; joffset = 1; break;                               \
        default: break;                                                                     \
        }                                                                                   \
        else switch (loc) { /* u-fold branch, :777-827 */                                   \
        case ORC_LOC_CENTER: ioffset = 0; joffset = 0; break;                               \
        case ORC_LOC_NECORNER:                                                              \
            ioffset = 1; joffset = 1;                                                       \
            for (i = 1; i <= nxg / 2 - 1; ++i) {                                            \
                int iDst = nxg - i;                                                         \
                T x1 = row2[i - 1], x2 = row2[iDst - 1];                                    \
                T xavg = (T)AVG(x1, isign * x2);                                            \
                row2[i - 1] = xavg;                                                         \
                row2[iDst - 1] = isign * xavg;                                              \
            }                                                                               \
            break;                                                                          \
        case ORC_LOC_EFACE: ioffset = 1; joffset = 0; break;                                \
        case ORC_LOC_NFACE:                                                                 \
            ioffset = 0; joffset = 1;                                                       \
            for (i = 1; i <= nxg / 2; ++i) {                                                \
                int iDst = nxg + 1 - i;                                                     \
                T x1 = row2[i - 1], x2 = row2[iDst - 1];                                    \
                T xavg = (T)AVG(x1, isign * x2);                                            \
                row2[i - 1] = xavg;                                                         \
                row2[iDst - 1] = isign * xavg;                                              \
            }                                                                               \
            break;                                                                          \
        default: break;                                                                     \
        }                                                                                   \
        /* copy-out, :3743-3761 and :837-866 (the message of buffer row 3 has jDst = -1) */ \
        for (j = 1; j <= 2; ++j) {                                                          \
            for (i = 1; i <= ihi + 1; ++i) {                                                \
                int ig = i - ilo + 1; /* i_glob(i), source/ice_blocks.F90:291-330 */        \
                int iSrc, jSrc, jDst;                                                       \
                if (ig < 1) ig = ewc ? ig + nxg : (g->ew_boundary == ORC_BND_OPEN ? 1 - ig : 0); \
                else if (ig > nxg) ig = ewc ? ig - nxg : (g->ew_boundary == ORC_BND_OPEN ? 2 * nxg - ig + 1 : 0); \
                iSrc = nxg - ig + 1;                                                        \
                jSrc = 1 + 3 - j; /* nghost + 3 - j */                                      \
                jDst = jhi + j - 1;                                                         \
                iSrc -= ioffset;                                                            \
                jSrc -= joffset;                                                            \
                if (iSrc == 0) iSrc = nxg;                                                  \
                if (iSrc > nxg) iSrc -= nxg;                                                \
                if (jSrc <= trows && jSrc > 0 && jDst > 0)                                  \
                    a[IX(i, jDst)] = isign * buf[(size_t)(jSrc - 1) * nxg + (iSrc - 1)];    \
            }                                                                               \
        }                                                                                   \
        free(buf);                                                                          \
    }                                                                                       \
    (void)fill;

#define AVG(x1, x2) (0.5 * ((x1) + (x2)))
void orc_halo_r8(double *a, const orc_grid *g, int loc, int kind, double fill) {
    HALO_BODY(double)
}
#undef AVG
/* integer variant (serial/ice_boundary.F90:1169-1451): xavg = nint(0.5*(x1+isign*x2)) -- Fortran nint rounds
 * half away from zero.  Reached by icetmask (a centre scalar) on the T-fold only. */
static int32_t orc_nint(double x) { return (int32_t)(x >= 0.0 ? floor(x + 0.5) : -floor(-x + 0.5)); }
#define AVG(x1, x2) orc_nint(0.5 * ((double)(x1) + (double)(x2)))
void orc_halo_i4(int32_t *a, const orc_grid *g, int loc, int kind, int32_t fill) {
    HALO_BODY(int32_t)
}
#undef AVG

/* ------------------------------------------------------------------ */
/* source/ice_grid.F90:1580-1633                                        */
/* ------------------------------------------------------------------ */
void orc_to_ugrid(const orc_grid *g, const double *tarea, const double *uarea,
                  const double *work1, double *work2) {
    const int nxb = g->nx_block;
    int i, j;
    memset(work2, 0, sizeof(double) * (size_t)g->nx_block * g->ny_block); /* :1612 */
    for (j = g->jlo; j <= g->jhi; ++j)
        for (i = g->ilo; i <= g->ihi; ++i)
            work2[IX(i, j)] = p25 *
                              (work1[IX(i, j)] * tarea[IX(i, j)] +
                               work1[IX(i + 1, j)] * tarea[IX(i + 1, j)] +
                               work1[IX(i, j + 1)] * tarea[IX(i, j + 1)] +
                               work1[IX(i + 1, j + 1)] * tarea[IX(i + 1, j + 1)]) /
                              uarea[IX(i, j)]; /* :1623-1628 */
}

/* source/ice_grid.F90:1684-1732 (ghosts of work2 are NOT touched) */
void orc_to_tgrid(const orc_grid *g, const double *tarea, const double *uarea,
                  const double *work1, double *work2) {
    const int nxb = g->nx_block;
    int i, j;
    for (j = g->jlo; j <= g->jhi; ++j)
        for (i = g->ilo; i <= g->ihi; ++i)
            work2[IX(i, j)] = p25 *
                              (work1[IX(i, j)] * uarea[IX(i, j)] +
                               work1[IX(i - 1, j)] * uarea[IX(i - 1, j)] +
                               work1[IX(i, j - 1)] * uarea[IX(i, j - 1)] +
                               work1[IX(i - 1, j - 1)] * uarea[IX(i - 1, j - 1)]) /
                              tarea[IX(i, j)]; /* :1722-1727 */
}

/* source/ice_grid.F90:1540-1571 */
void orc_t2ugrid_vector(const orc_grid *g, const double *tarea, const double *uarea,
                        double *work) {
    size_t n = (size_t)g->nx_block * g->ny_block;
    double *work1 = (double *)malloc(sizeof(double) * n);
    memcpy(work1, work, sizeof(double) * n);                         /* :1562 */
    orc_halo_r8(work1, g, ORC_LOC_CENTER, ORC_TYPE_VECTOR, 0.0);     /* :1565 */
    orc_to_ugrid(g, tarea, uarea, work1, work);                      /* :1569 */
    free(work1);
}

/* source/ice_grid.F90:1642-1675 */
void orc_u2tgrid_vector(const orc_grid *g, const double *tarea, const double *uarea,
                        double *work) {
    size_t n = (size_t)g->nx_block * g->ny_block;
    double *work1 = (double *)malloc(sizeof(double) * n);
    memcpy(work1, work, sizeof(double) * n);                         /* :1666 */
    orc_halo_r8(work1, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);   /* :1669 */
    orc_to_tgrid(g, tarea, uarea, work1, work);                      /* :1673 */
    free(work1);
}

/* ------------------------------------------------------------------ */
/* source/ice_dyn_evp.F90:586-694                                       */
/* ------------------------------------------------------------------ */
void orc_evp_prep1(const orc_grid *g, const orc_params *p, const orc_fields *f) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const double a_min = 0.001, m_min = 0.01; /* :87-88 */
    int i, j;
    unsigned char *tmphm = (unsigned char *)malloc((size_t)nxb * nyb);
    for (j = 1; j <= nyb; ++j)
        for (i = 1; i <= nxb; ++i) {
            size_t k = IX(i, j);
            if (f->tmask[k])
                f->tmass[k] = (p->rhoi * f->vice[k] + p->rhos * f->vsno[k]); /* :652 */
            else
                f->tmass[k] = c0;
            tmphm[k] = f->tmask[k] && (f->aice[k] > a_min) && (f->tmass[k] > m_min); /* :660 */
            f->strairx[k] = f->strairxT[k]; /* :668 */
            f->strairy[k] = f->strairyT[k];
            f->icetmask[k] = 0; /* :674 */
        }
    for (j = g->jlo; j <= g->jhi; ++j)
        for (i = g->ilo; i <= g->ihi; ++i) {
            if (tmphm[IX(i - 1, j + 1)] || tmphm[IX(i, j + 1)] || tmphm[IX(i + 1, j + 1)] ||
                tmphm[IX(i - 1, j)] || tmphm[IX(i, j)] || tmphm[IX(i + 1, j)] ||
                tmphm[IX(i - 1, j - 1)] || tmphm[IX(i, j - 1)] || tmphm[IX(i + 1, j - 1)])
                f->icetmask[IX(i, j)] = 1; /* :683-687 */
            if (!f->tmask[IX(i, j)]) f->icetmask[IX(i, j)] = 0; /* :689 */
        }
    free(tmphm);
}

/* single-precision sign(1., real(fm)) of source/ice_dyn_evp.F90:912-913 */
static inline double sign1_real(double fm) {
    float r = (float)fm;
    return signbit(r) ? -1.0 : 1.0;
}

/* ------------------------------------------------------------------ */
/* source/ice_dyn_evp.F90:703-938                                       */
/* ------------------------------------------------------------------ */
void orc_evp_prep2(const orc_grid *g, const orc_params *p, const orc_fields *f,
                   int32_t *icellt, int32_t *icellu, int32_t *indxti, int32_t *indxtj,
                   int32_t *indxui, int32_t *indxuj) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const double a_min = 0.001, m_min = 0.01;
    int i, j, ij;
    for (j = 1; j <= nyb; ++j)
        for (i = 1; i <= nxb; ++i) {
            size_t k = IX(i, j);
            f->waterx[k] = c0; /* :821-825 */
            f->watery[k] = c0;
            f->forcex[k] = c0;
            f->forcey[k] = c0;
            f->umassdtei[k] = c0;
            if (f->icetmask[k] == 0) { /* :827-840 */
                f->stressp_1[k] = c0; f->stressp_2[k] = c0; f->stressp_3[k] = c0; f->stressp_4[k] = c0;
                f->stressm_1[k] = c0; f->stressm_2[k] = c0; f->stressm_3[k] = c0; f->stressm_4[k] = c0;
                f->stress12_1[k] = c0; f->stress12_2[k] = c0; f->stress12_3[k] = c0; f->stress12_4[k] = c0;
            }
        }
    *icellt = 0; /* :850-859 */
    for (j = g->jlo; j <= g->jhi + 1; ++j)
        for (i = g->ilo; i <= g->ihi + 1; ++i)
            if (f->icetmask[IX(i, j)] == 1) {
                indxti[*icellt] = i;
                indxtj[*icellt] = j;
                *icellt += 1;
            }
    *icellu = 0; /* :867-896 */
    for (j = g->jlo; j <= g->jhi; ++j)
        for (i = g->ilo; i <= g->ihi; ++i) {
            size_t k = IX(i, j);
            int old = f->iceumask[k];
            f->iceumask[k] = f->umask[k] && (f->aiu[k] > a_min) && (f->umass[k] > m_min);
            if (f->iceumask[k]) {
                indxui[*icellu] = i;
                indxuj[*icellu] = j;
                *icellu += 1;
                if (!old) {
                    f->uvel[k] = f->uocn[k];
                    f->vvel[k] = f->vocn[k];
                }
            } else {
                f->uvel[k] = c0; f->vvel[k] = c0;
                f->strintx[k] = c0; f->strinty[k] = c0;
                f->strocnx[k] = c0; f->strocny[k] = c0;
            }
        }
    for (ij = 0; ij < *icellu; ++ij) { /* :902-936 */
        size_t k;
        i = indxui[ij];
        j = indxuj[ij];
        k = IX(i, j);
        f->umassdtei[k] = f->umass[k] * p->dtei;
        f->fm[k] = f->fcor[k] * f->umass[k];
        if (p->auscom) { /* :910-913 */
            double s = sign1_real(f->fm[k]);
            f->waterx[k] = f->uocn[k] * p->cosw - f->vocn[k] * p->sinw * s;
            f->watery[k] = f->vocn[k] * p->cosw + f->uocn[k] * p->sinw * s;
        } else { /* :915-916 */
            f->waterx[k] = f->uocn[k] * p->cosw - f->vocn[k] * p->sinw;
            f->watery[k] = f->vocn[k] * p->cosw + f->uocn[k] * p->sinw;
        }
        if (!p->coupled) { /* :919-922 */
            f->strtltx[k] = -f->fm[k] * f->vocn[k];
            f->strtlty[k] = f->fm[k] * f->uocn[k];
        } else { /* :924-925 */
            f->strtltx[k] = -p->gravit * f->umass[k] * f->ss_tltx[k];
            f->strtlty[k] = -p->gravit * f->umass[k] * f->ss_tlty[k];
        }
        if (p->auscom && !p->use_ocnslope) { /* :928-933 */
            f->strtltx[k] = -f->fm[k] * f->vocn[k];
            f->strtlty[k] = f->fm[k] * f->uocn[k];
        }
        f->forcex[k] = f->strairx[k] + f->strtltx[k]; /* :934-935 */
        f->forcey[k] = f->strairy[k] + f->strtlty[k];
    }
}

/* ------------------------------------------------------------------ */
/* source/ice_mechred.F90:1869-2036 with asum_ridging :573-631 and      */
/* ridge_itd :773-1081 (per-cell restatement of the compressed loops)    */
/* ------------------------------------------------------------------ */
#define ORC_MAXCAT 16
void orc_ice_strength(const orc_grid *g, const orc_params *p, const orc_fields *f,
                      int32_t icellt, const int32_t *indxti, const int32_t *indxtj) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const size_t plane = (size_t)nxb * nyb;
    const int ncat = p->ncat;
    const double puny = p->puny;
    const double Cf = 17.0;                                             /* :83 */
    const double Cp = p5 * p->gravit * (p->rhow - p->rhoi) * p->rhoi / p->rhow; /* :86 */
    const double Gstar = 0.15, astar = 0.05, maxraft = 1.0, Hstar = 25.0; /* :89-95 */
    const double Pstar = 2.75e4, Cstar = 20.0;                          /* :98-100 */
    const double Gstari = c1 / Gstar, astari = c1 / astar;              /* :825-826 */
    int i, j, n, ij;
    memset(f->strength, 0, sizeof(double) * plane); /* :1942 */
    if (p->kstrength == 1) {
        for (ij = 0; ij < icellt; ++ij) {
            double Gsum_[ORC_MAXCAT + 2], *Gsum = Gsum_ + 1; /* Gsum(-1:ncat) */
            double apartic[ORC_MAXCAT + 1], hrmin[ORC_MAXCAT + 1], hrmax[ORC_MAXCAT + 1];
            double hrexp[ORC_MAXCAT + 1], krdg[ORC_MAXCAT + 1];
            double work, aksum, hi, hrmean, xtmp, h2rdg, dh2rdg, s;
            size_t k;
            i = indxti[ij];
            j = indxtj[ij];
            k = IX(i, j);
            /* asum_ridging (:610-629) computes asum, which ice_strength never uses */
            Gsum[-1] = c0; /* :845-858 */
            Gsum[0] = c1;
            apartic[0] = c0;
            for (n = 1; n <= ncat; ++n) {
                Gsum[n] = c1; apartic[n] = c0; hrmin[n] = c0; hrmax[n] = c0; hrexp[n] = c0; krdg[n] = c1;
            }
            if (f->aice0[k] > puny) Gsum[0] = f->aice0[k]; else Gsum[0] = Gsum[-1]; /* :877-881 */
            for (n = 1; n <= ncat; ++n) { /* :884-897 */
                double an = f->aicen[(size_t)(n - 1) * plane + k];
                if (an > puny) Gsum[n] = Gsum[n - 1] + an; else Gsum[n] = Gsum[n - 1];
            }
            work = c1 / Gsum[ncat]; /* :902 */
            for (n = 0; n <= ncat; ++n) Gsum[n] = Gsum[n] * work; /* :904-911 */
            if (p->krdg_partic == 0) { /* :931-941 */
                for (n = 0; n <= ncat; ++n) {
                    if (Gsum[n] < Gstar)
                        apartic[n] = Gstari * (Gsum[n] - Gsum[n - 1]) * (c2 - (Gsum[n - 1] + Gsum[n]) * Gstari);
                    else if (Gsum[n - 1] < Gstar)
                        apartic[n] = Gstari * (Gstar - Gsum[n - 1]) * (c2 - (Gsum[n - 1] + Gstar) * Gstari);
                }
            } else { /* :954-969 */
                xtmp = c1 / (c1 - exp(-astari));
                for (n = -1; n <= ncat; ++n) Gsum[n] = exp(-Gsum[n] * astari) * xtmp;
                for (n = 0; n <= ncat; ++n) apartic[n] = Gsum[n - 1] - Gsum[n];
            }
            for (n = 1; n <= ncat; ++n) {
                double an = f->aicen[(size_t)(n - 1) * plane + k];
                double vn = f->vicen[(size_t)(n - 1) * plane + k];
                if (an > puny) {
                    if (p->krdg_redist == 0) { /* :1001-1008 */
                        hi = vn / an;
                        hrmin[n] = dmin(c2 * hi, hi + maxraft);
                        hrmax[n] = c2 * sqrt(Hstar * hi);
                        hrmax[n] = dmax(hrmax[n], hrmin[n] + puny);
                        hrmean = p5 * (hrmin[n] + hrmax[n]);
                        krdg[n] = hrmean / hi;
                    } else { /* :1047-1053 */
                        hi = vn / an;
                        hi = dmax(hi, puny);
                        hrmin[n] = dmin(c2 * hi, hi + maxraft);
                        hrexp[n] = p->mu_rdg * sqrt(hi);
                        krdg[n] = (hrmin[n] + hrexp[n]) / hi;
                    }
                }
            }
            aksum = apartic[0]; /* :1066-1079 */
            for (n = 1; n <= ncat; ++n) aksum = aksum + apartic[n] * (c1 - c1 / krdg[n]);
            s = c0;
            for (n = 1; n <= ncat; ++n) { /* :1970-2007 */
                double an = f->aicen[(size_t)(n - 1) * plane + k];
                double vn = f->vicen[(size_t)(n - 1) * plane + k];
                if (an > puny && apartic[n] > c0) {
                    hi = vn / an;
                    if (p->krdg_redist == 0)
                        h2rdg = P333 * (hrmax[n] * hrmax[n] * hrmax[n] - hrmin[n] * hrmin[n] * hrmin[n]) /
                                (hrmax[n] - hrmin[n]);
                    else
                        h2rdg = hrmin[n] * hrmin[n] + c2 * hrmin[n] * hrexp[n] + c2 * hrexp[n] * hrexp[n];
                    dh2rdg = -hi * hi + h2rdg / krdg[n];
                    s = s + apartic[n] * dh2rdg;
                }
            }
            f->strength[k] = Cf * Cp * s / aksum; /* :2017 */
        }
    } else { /* :2028-2032 */
        for (j = g->jlo; j <= g->jhi; ++j)
            for (i = g->ilo; i <= g->ihi; ++i)
                f->strength[IX(i, j)] = Pstar * f->vice[IX(i, j)] * exp(-Cstar * (c1 - f->aice[IX(i, j)]));
    }
}

/* ------------------------------------------------------------------ */
/* source/ice_dyn_evp.F90:947-1293                                      */
/* str is (nx_block, ny_block, 8)                                        */
/* ------------------------------------------------------------------ */
void orc_stress(const orc_grid *g, const orc_params *p, const orc_fields *f, int ksub,
                int32_t icellt, const int32_t *indxti, const int32_t *indxtj, double *str) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const size_t plane = (size_t)nxb * nyb;
    const double ecci = p->ecci, dte2T = p->dte2T, denom1 = p->denom1, denom2 = p->denom2;
    const double rcon = p->rcon;
    const double p055 = P055, p027 = P027, p111 = P111, p222 = P222, p333 = P333, p166 = P166;
    const double *uvel = f->uvel, *vvel = f->vvel;
    const double *cyp = f->cyp, *cxp = f->cxp, *cym = f->cym, *cxm = f->cxm;
    const double *dxt = f->dxt, *dyt = f->dyt, *dxhy = f->dxhy, *dyhx = f->dyhx;
    int ij;

    memset(str, 0, sizeof(double) * plane * 8); /* :1051 */

#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (ij = 0; ij < icellt; ++ij) {
        const int i = indxti[ij], j = indxtj[ij];
        const size_t k = IX(i, j);
        const size_t kw = IX(i - 1, j), ks = IX(i, j - 1), ksw = IX(i - 1, j - 1);
        double divune, divunw, divuse, divusw, tensionne, tensionnw, tensionse, tensionsw;
        double shearne, shearnw, shearse, shearsw, Deltane, Deltanw, Deltase, Deltasw;
        double c0ne, c0nw, c0se, c0sw, c1ne, c1nw, c1se, c1sw;
        double ssigpn, ssigps, ssigpe, ssigpw, ssigmn, ssigms, ssigme, ssigmw;
        double ssig12n, ssig12s, ssig12e, ssig12w, ssigp1, ssigp2, ssigm1, ssigm2, ssig121, ssig122;
        double csigpne, csigpnw, csigpse, csigpsw, csigmne, csigmnw, csigmse, csigmsw;
        double csig12ne, csig12nw, csig12se, csig12sw, str12ew, str12we, str12ns, str12sn;
        double strp_tmp, strm_tmp, tmp;

        /* :1065-1072 */
        divune = cyp[k] * uvel[k] - dyt[k] * uvel[kw] + cxp[k] * vvel[k] - dxt[k] * vvel[ks];
        divunw = cym[k] * uvel[kw] + dyt[k] * uvel[k] + cxp[k] * vvel[kw] - dxt[k] * vvel[ksw];
        divusw = cym[k] * uvel[ksw] + dyt[k] * uvel[ks] + cxm[k] * vvel[ksw] + dxt[k] * vvel[kw];
        divuse = cyp[k] * uvel[ks] - dyt[k] * uvel[ksw] + cxm[k] * vvel[ks] + dxt[k] * vvel[k];
        /* :1075-1082 */
        tensionne = -cym[k] * uvel[k] - dyt[k] * uvel[kw] + cxm[k] * vvel[k] + dxt[k] * vvel[ks];
        tensionnw = -cyp[k] * uvel[kw] + dyt[k] * uvel[k] + cxm[k] * vvel[kw] + dxt[k] * vvel[ksw];
        tensionsw = -cyp[k] * uvel[ksw] + dyt[k] * uvel[ks] + cxp[k] * vvel[ksw] - dxt[k] * vvel[kw];
        tensionse = -cym[k] * uvel[ks] - dyt[k] * uvel[ksw] + cxp[k] * vvel[ks] - dxt[k] * vvel[k];
        /* :1085-1092 */
        shearne = -cym[k] * vvel[k] - dyt[k] * vvel[kw] - cxm[k] * uvel[k] - dxt[k] * uvel[ks];
        shearnw = -cyp[k] * vvel[kw] + dyt[k] * vvel[k] - cxm[k] * uvel[kw] - dxt[k] * uvel[ksw];
        shearsw = -cyp[k] * vvel[ksw] + dyt[k] * vvel[ks] - cxp[k] * uvel[ksw] + dxt[k] * uvel[kw];
        shearse = -cym[k] * vvel[ks] - dyt[k] * vvel[ksw] - cxp[k] * uvel[ks] + dxt[k] * uvel[k];
        /* :1095-1098 */
        Deltane = sqrt(divune * divune + ecci * (tensionne * tensionne + shearne * shearne));
        Deltanw = sqrt(divunw * divunw + ecci * (tensionnw * tensionnw + shearnw * shearnw));
        Deltase = sqrt(divuse * divuse + ecci * (tensionse * tensionse + shearse * shearse));
        Deltasw = sqrt(divusw * divusw + ecci * (tensionsw * tensionsw + shearsw * shearsw));

        if (ksub == p->ndte) { /* :1103-1115 */
            f->divu[k] = p25 * (divune + divunw + divuse + divusw) * f->tarear[k];
            tmp = p25 * (Deltane + Deltanw + Deltase + Deltasw) * f->tarear[k];
            f->rdg_conv[k] = -dmin(f->divu[k], c0);
            f->rdg_shear[k] = p5 * (tmp - fabs(f->divu[k]));
            f->shear[k] = p25 * f->tarear[k] *
                          sqrt((tensionne + tensionnw + tensionse + tensionsw) *
                                   (tensionne + tensionnw + tensionse + tensionsw) +
                               (shearne + shearnw + shearse + shearsw) *
                                   (shearne + shearnw + shearse + shearsw));
        }

        if (p->evp_damping) { /* :1121-1128 */
            c0ne = dmin(f->strength[k] / dmax(Deltane, c4 * f->tinyarea[k]), rcon);
            c0nw = dmin(f->strength[k] / dmax(Deltanw, c4 * f->tinyarea[k]), rcon);
            c0sw = dmin(f->strength[k] / dmax(Deltasw, c4 * f->tinyarea[k]), rcon);
            c0se = dmin(f->strength[k] / dmax(Deltase, c4 * f->tinyarea[k]), rcon);
            f->prs_sig[k] = f->strength[k] * Deltane / dmax(Deltane, c4 * f->tinyarea[k]);
        } else { /* :1131-1135 */
            c0ne = f->strength[k] / dmax(Deltane, f->tinyarea[k]);
            c0nw = f->strength[k] / dmax(Deltanw, f->tinyarea[k]);
            c0sw = f->strength[k] / dmax(Deltasw, f->tinyarea[k]);
            c0se = f->strength[k] / dmax(Deltase, f->tinyarea[k]);
            f->prs_sig[k] = c0ne * Deltane;
        }
        c1ne = c0ne * dte2T; /* :1138-1141 */
        c1nw = c0nw * dte2T;
        c1sw = c0sw * dte2T;
        c1se = c0se * dte2T;

        /* :1148-1165 */
        f->stressp_1[k] = (f->stressp_1[k] + c1ne * (divune - Deltane)) * denom1;
        f->stressp_2[k] = (f->stressp_2[k] + c1nw * (divunw - Deltanw)) * denom1;
        f->stressp_3[k] = (f->stressp_3[k] + c1sw * (divusw - Deltasw)) * denom1;
        f->stressp_4[k] = (f->stressp_4[k] + c1se * (divuse - Deltase)) * denom1;

        f->stressm_1[k] = (f->stressm_1[k] + c1ne * tensionne) * denom2;
        f->stressm_2[k] = (f->stressm_2[k] + c1nw * tensionnw) * denom2;
        f->stressm_3[k] = (f->stressm_3[k] + c1sw * tensionsw) * denom2;
        f->stressm_4[k] = (f->stressm_4[k] + c1se * tensionse) * denom2;

        f->stress12_1[k] = (f->stress12_1[k] + c1ne * shearne * p5) * denom2;
        f->stress12_2[k] = (f->stress12_2[k] + c1nw * shearnw * p5) * denom2;
        f->stress12_3[k] = (f->stress12_3[k] + c1sw * shearsw * p5) * denom2;
        f->stress12_4[k] = (f->stress12_4[k] + c1se * shearse * p5) * denom2;

        /* :1196-1215 */
        ssigpn = f->stressp_1[k] + f->stressp_2[k];
        ssigps = f->stressp_3[k] + f->stressp_4[k];
        ssigpe = f->stressp_1[k] + f->stressp_4[k];
        ssigpw = f->stressp_2[k] + f->stressp_3[k];
        ssigp1 = (f->stressp_1[k] + f->stressp_3[k]) * p055;
        ssigp2 = (f->stressp_2[k] + f->stressp_4[k]) * p055;

        ssigmn = f->stressm_1[k] + f->stressm_2[k];
        ssigms = f->stressm_3[k] + f->stressm_4[k];
        ssigme = f->stressm_1[k] + f->stressm_4[k];
        ssigmw = f->stressm_2[k] + f->stressm_3[k];
        ssigm1 = (f->stressm_1[k] + f->stressm_3[k]) * p055;
        ssigm2 = (f->stressm_2[k] + f->stressm_4[k]) * p055;

        ssig12n = f->stress12_1[k] + f->stress12_2[k];
        ssig12s = f->stress12_3[k] + f->stress12_4[k];
        ssig12e = f->stress12_1[k] + f->stress12_4[k];
        ssig12w = f->stress12_2[k] + f->stress12_3[k];
        ssig121 = (f->stress12_1[k] + f->stress12_3[k]) * p111;
        ssig122 = (f->stress12_2[k] + f->stress12_4[k]) * p111;

        /* :1217-1234 */
        csigpne = p111 * f->stressp_1[k] + ssigp2 + p027 * f->stressp_3[k];
        csigpnw = p111 * f->stressp_2[k] + ssigp1 + p027 * f->stressp_4[k];
        csigpsw = p111 * f->stressp_3[k] + ssigp2 + p027 * f->stressp_1[k];
        csigpse = p111 * f->stressp_4[k] + ssigp1 + p027 * f->stressp_2[k];

        csigmne = p111 * f->stressm_1[k] + ssigm2 + p027 * f->stressm_3[k];
        csigmnw = p111 * f->stressm_2[k] + ssigm1 + p027 * f->stressm_4[k];
        csigmsw = p111 * f->stressm_3[k] + ssigm2 + p027 * f->stressm_1[k];
        csigmse = p111 * f->stressm_4[k] + ssigm1 + p027 * f->stressm_2[k];

        csig12ne = p222 * f->stress12_1[k] + ssig122 + p055 * f->stress12_3[k];
        csig12nw = p222 * f->stress12_2[k] + ssig121 + p055 * f->stress12_4[k];
        csig12sw = p222 * f->stress12_3[k] + ssig122 + p055 * f->stress12_1[k];
        csig12se = p222 * f->stress12_4[k] + ssig121 + p055 * f->stress12_2[k];

        /* :1236-1239 */
        str12ew = p5 * dxt[k] * (p333 * ssig12e + p166 * ssig12w);
        str12we = p5 * dxt[k] * (p333 * ssig12w + p166 * ssig12e);
        str12ns = p5 * dyt[k] * (p333 * ssig12n + p166 * ssig12s);
        str12sn = p5 * dyt[k] * (p333 * ssig12s + p166 * ssig12n);

        /* :1244-1264 */
        strp_tmp = p25 * dyt[k] * (p333 * ssigpn + p166 * ssigps);
        strm_tmp = p25 * dyt[k] * (p333 * ssigmn + p166 * ssigms);
        str[0 * plane + k] = -strp_tmp - strm_tmp - str12ew + dxhy[k] * (-csigpne + csigmne) + dyhx[k] * csig12ne;
        str[1 * plane + k] = strp_tmp + strm_tmp - str12we + dxhy[k] * (-csigpnw + csigmnw) + dyhx[k] * csig12nw;
        strp_tmp = p25 * dyt[k] * (p333 * ssigps + p166 * ssigpn);
        strm_tmp = p25 * dyt[k] * (p333 * ssigms + p166 * ssigmn);
        str[2 * plane + k] = -strp_tmp - strm_tmp + str12ew + dxhy[k] * (-csigpse + csigmse) + dyhx[k] * csig12se;
        str[3 * plane + k] = strp_tmp + strm_tmp + str12we + dxhy[k] * (-csigpsw + csigmsw) + dyhx[k] * csig12sw;

        /* :1269-1289 */
        strp_tmp = p25 * dxt[k] * (p333 * ssigpe + p166 * ssigpw);
        strm_tmp = p25 * dxt[k] * (p333 * ssigme + p166 * ssigmw);
        str[4 * plane + k] = -strp_tmp + strm_tmp - str12ns - dyhx[k] * (csigpne + csigmne) + dxhy[k] * csig12ne;
        str[5 * plane + k] = strp_tmp - strm_tmp - str12sn - dyhx[k] * (csigpse + csigmse) + dxhy[k] * csig12se;
        strp_tmp = p25 * dxt[k] * (p333 * ssigpw + p166 * ssigpe);
        strm_tmp = p25 * dxt[k] * (p333 * ssigmw + p166 * ssigme);
        str[6 * plane + k] = -strp_tmp + strm_tmp + str12ns - dyhx[k] * (csigpnw + csigmnw) + dxhy[k] * csig12nw;
        str[7 * plane + k] = strp_tmp - strm_tmp + str12sn - dyhx[k] * (csigpsw + csigmsw) + dxhy[k] * csig12sw;
    }
}

/* ------------------------------------------------------------------ */
/* source/ice_dyn_evp.F90:1302-1443                                     */
/* ------------------------------------------------------------------ */
void orc_stepu(const orc_grid *g, const orc_params *p, const orc_fields *f,
               int32_t icellu, const int32_t *indxui, const int32_t *indxuj, const double *str) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const size_t plane = (size_t)nxb * nyb;
    const double dragw = p->dragio * p->rhow; /* :78 / :1382 */
    const double cosw = p->cosw, sinw = p->sinw;
    int ij;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (ij = 0; ij < icellu; ++ij) {
        const int i = indxui[ij], j = indxuj[ij];
        const size_t k = IX(i, j), ke = IX(i + 1, j), kn = IX(i, j + 1), kne = IX(i + 1, j + 1);
        double uold, vold, vrel, cca, ccb, ab2, cc1, cc2, taux, tauy;
        uold = f->uvel[k];
        vold = f->vvel[k];
        /* :1394-1395 */
        vrel = f->aiu[k] * dragw *
               sqrt((f->uocn[k] - uold) * (f->uocn[k] - uold) + (f->vocn[k] - vold) * (f->vocn[k] - vold));
        taux = vrel * f->waterx[k]; /* :1397-1398 */
        tauy = vrel * f->watery[k];
        cca = f->umassdtei[k] + vrel * cosw; /* :1401 */
        if (p->auscom && f->fm[k] < 0.0) /* :1403-1408 */
            ccb = f->fm[k] - vrel * sinw;
        else
            ccb = f->fm[k] + vrel * sinw; /* :1410 */
        ab2 = cca * cca + ccb * ccb;      /* :1412 */
        /* :1415-1418 */
        f->strintx[k] = f->uarear[k] * (str[0 * plane + k] + str[1 * plane + ke] + str[2 * plane + kn] + str[3 * plane + kne]);
        f->strinty[k] = f->uarear[k] * (str[4 * plane + k] + str[5 * plane + kn] + str[6 * plane + ke] + str[7 * plane + kne]);
        /* :1421-1427 */
        cc1 = f->strintx[k] + f->forcex[k] + taux + f->umassdtei[k] * uold;
        cc2 = f->strinty[k] + f->forcey[k] + tauy + f->umassdtei[k] * vold;
        f->uvel[k] = (cca * cc1 + ccb * cc2) / ab2;
        f->vvel[k] = (cca * cc2 - ccb * cc1) / ab2;
        f->strocnx[k] = taux; /* :1434-1435 */
        f->strocny[k] = tauy;
    }
}

/* ------------------------------------------------------------------ */
/* source/ice_dyn_evp.F90:1452-1549                                     */
/* ------------------------------------------------------------------ */
void orc_evp_finish(const orc_grid *g, const orc_params *p, const orc_fields *f,
                    int32_t icellu, const int32_t *indxui, const int32_t *indxuj) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const double dragw = p->dragio * p->rhow;
    const double cosw = p->cosw, sinw = p->sinw;
    int ij;
    memset(f->strocnxT, 0, sizeof(double) * (size_t)nxb * nyb); /* :1510-1515 */
    memset(f->strocnyT, 0, sizeof(double) * (size_t)nxb * nyb);
    for (ij = 0; ij < icellu; ++ij) {
        const int i = indxui[ij], j = indxuj[ij];
        const size_t k = IX(i, j);
        double vrel = dragw * sqrt((f->uocn[k] - f->uvel[k]) * (f->uocn[k] - f->uvel[k]) +
                                   (f->vocn[k] - f->vvel[k]) * (f->vocn[k] - f->vvel[k])); /* :1522 */
        if (p->auscom && f->fm[k] < 0.0) { /* :1525-1530 */
            f->strocnx[k] = f->strocnx[k] - vrel * (f->uvel[k] * cosw + f->vvel[k] * sinw) * f->aiu[k];
            f->strocny[k] = f->strocny[k] - vrel * (f->vvel[k] * cosw - f->uvel[k] * sinw) * f->aiu[k];
        } else { /* :1532-1541 */
            f->strocnx[k] = f->strocnx[k] - vrel * (f->uvel[k] * cosw - f->vvel[k] * sinw) * f->aiu[k];
            f->strocny[k] = f->strocny[k] - vrel * (f->vvel[k] * cosw + f->uvel[k] * sinw) * f->aiu[k];
        }
        f->strocnxT[k] = f->strocnx[k] / f->aiu[k]; /* :1545-1546 */
        f->strocnyT[k] = f->strocny[k] / f->aiu[k];
    }
}

/* source/ice_dyn_evp.F90:1558-1609 */
void orc_principal_stress(int nx_block, int ny_block, const double *stressp_1,
                          const double *stressm_1, const double *stress12_1,
                          const double *prs_sig, double puny, double *sig1, double *sig2) {
    const double spval_dbl = 1.0e30;
    size_t k, n = (size_t)nx_block * ny_block;
    for (k = 0; k < n; ++k) {
        if (prs_sig[k] > puny) {
            sig1[k] = (p5 * (stressp_1[k] + sqrt(stressm_1[k] * stressm_1[k] + c4 * (stress12_1[k] * stress12_1[k])))) / prs_sig[k];
            sig2[k] = (p5 * (stressp_1[k] - sqrt(stressm_1[k] * stressm_1[k] + c4 * (stress12_1[k] * stress12_1[k])))) / prs_sig[k];
        } else {
            sig1[k] = spval_dbl;
            sig2[k] = spval_dbl;
        }
    }
}

/* ------------------------------------------------------------------ */
/* driver: source/ice_dyn_evp.F90:119-432                               */
/* ------------------------------------------------------------------ */
static double *g_strength_pre;
static size_t g_strength_pre_n;
/* copies the strength array of the last orc_evp call as it was BEFORE the halo update; returns its size */
size_t orc_last_strength_prehalo(double *out) {
    if (out && g_strength_pre) memcpy(out, g_strength_pre, sizeof(double) * g_strength_pre_n);
    return g_strength_pre_n;
}

int orc_evp(const orc_grid *g, const orc_params *p, const orc_fields *f, double *subcycle_seconds) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const size_t plane = (size_t)nxb * nyb;
    int32_t icellt = 0, icellu = 0;
    int32_t *indxti = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxtj = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxui = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxuj = (int32_t *)malloc(sizeof(int32_t) * plane);
    double *str = (double *)malloc(sizeof(double) * plane * 8);
    int ksub;
    double t0;
    if (!indxti || !indxtj || !indxui || !indxuj || !str) return -1;

    /* :214-224 */
    memset(f->rdg_conv, 0, sizeof(double) * plane);
    memset(f->rdg_shear, 0, sizeof(double) * plane);
    memset(f->divu, 0, sizeof(double) * plane);
    memset(f->shear, 0, sizeof(double) * plane);
    memset(f->prs_sig, 0, sizeof(double) * plane);

    orc_evp_prep1(g, p, f); /* :236-242 */
    if (p->auscom && f->sicemass) memcpy(f->sicemass, f->tmass, sizeof(double) * plane); /* :246-248 */

    orc_halo_i4(f->icetmask, g, ORC_LOC_CENTER, ORC_TYPE_SCALAR, 0); /* :250-253 */

    orc_to_ugrid(g, f->tarea, f->uarea, f->tmass, f->umass); /* :259-260 */
    orc_to_ugrid(g, f->tarea, f->uarea, f->aice, f->aiu);

    if (p->access_wind) { /* :271-275 */
        memcpy(f->strairx, f->strax, sizeof(double) * plane);
        memcpy(f->strairy, f->stray, sizeof(double) * plane);
    }
    orc_t2ugrid_vector(g, f->tarea, f->uarea, f->strairx); /* :276-277 */
    orc_t2ugrid_vector(g, f->tarea, f->uarea, f->strairy);

    orc_evp_prep2(g, p, f, &icellt, &icellu, indxti, indxtj, indxui, indxuj); /* :292-316 */

    if (f->strength_in)
        memcpy(f->strength, f->strength_in, sizeof(double) * plane);
    else
        orc_ice_strength(g, p, f, icellt, indxti, indxtj); /* :322-332 */

    /* tests: what ice_strength returned, before evp's halo update of it (the array a host that runs
     * ice_strength itself passes to the two-phase entry; on the T-fold the update is not idempotent) */
    free(g_strength_pre);
    g_strength_pre = (double *)malloc(sizeof(double) * plane);
    g_strength_pre_n = g_strength_pre ? plane : 0;
    if (g_strength_pre) memcpy(g_strength_pre, f->strength, sizeof(double) * plane);
    orc_halo_r8(f->strength, g, ORC_LOC_CENTER, ORC_TYPE_SCALAR, 0.0); /* :337-343 */
    orc_halo_r8(f->uvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
    orc_halo_r8(f->vvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);

    t0 = now_s();
    for (ksub = 1; ksub <= p->ndte; ++ksub) { /* :347-404 */
        orc_stress(g, p, f, ksub, icellt, indxti, indxtj, str);
        orc_stepu(g, p, f, icellu, indxui, indxuj, str);
        orc_halo_r8(f->uvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
        orc_halo_r8(f->vvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
    }
    if (subcycle_seconds) *subcycle_seconds = now_s() - t0;

    orc_evp_finish(g, p, f, icellu, indxui, indxuj); /* :410-425 */
    orc_u2tgrid_vector(g, f->tarea, f->uarea, f->strocnxT); /* :427-428 */
    orc_u2tgrid_vector(g, f->tarea, f->uarea, f->strocnyT);

    free(indxti); free(indxtj); free(indxui); free(indxuj); free(str);
    return 0;
}

/* Timing helper: the ndte loop only (stress + stepu + 2 halos), on fields that a
 * previous orc_evp call has prepared (icetmask, iceumask, aiu, umassdtei, ... valid). */
int orc_subcycle_only(const orc_grid *g, const orc_params *p, const orc_fields *f, int nsub,
                      double *seconds) {
    const int nxb = g->nx_block, nyb = g->ny_block;
    const size_t plane = (size_t)nxb * nyb;
    int32_t icellt = 0, icellu = 0;
    int32_t *indxti = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxtj = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxui = (int32_t *)malloc(sizeof(int32_t) * plane);
    int32_t *indxuj = (int32_t *)malloc(sizeof(int32_t) * plane);
    double *str = (double *)malloc(sizeof(double) * plane * 8);
    int i, j, ksub;
    double t0;
    if (!indxti || !indxtj || !indxui || !indxuj || !str) return -1;
    for (j = g->jlo; j <= g->jhi + 1; ++j) /* source/ice_dyn_evp.F90:850-859 */
        for (i = g->ilo; i <= g->ihi + 1; ++i)
            if (f->icetmask[IX(i, j)] == 1) { indxti[icellt] = i; indxtj[icellt] = j; ++icellt; }
    for (j = g->jlo; j <= g->jhi; ++j) /* :867-880 */
        for (i = g->ilo; i <= g->ihi; ++i)
            if (f->iceumask[IX(i, j)]) { indxui[icellu] = i; indxuj[icellu] = j; ++icellu; }
    t0 = now_s();
    for (ksub = 1; ksub <= nsub; ++ksub) {
        orc_stress(g, p, f, ksub == nsub ? p->ndte : 0, icellt, indxti, indxtj, str);
        orc_stepu(g, p, f, icellu, indxui, indxuj, str);
        orc_halo_r8(f->uvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
        orc_halo_r8(f->vvel, g, ORC_LOC_NECORNER, ORC_TYPE_VECTOR, 0.0);
    }
    if (seconds) *seconds = now_s() - t0;
    free(indxti); free(indxtj); free(indxui); free(indxuj); free(str);
    return 0;
}

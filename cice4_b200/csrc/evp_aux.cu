// evp_aux.cu -- once-per-call kernels of libevp_b200: host-layout marshalling, evp_prep1/2,
// T<->U grid operators, halo updates (east-west wrap, north-south cyclic, tripole u-fold),
// evp_finish, principal_stress and the device ice_strength.
// Compiled with -fmad=false: these kernels always use unfused IEEE arithmetic in the
// reference's operation order, so their results are bit-identical to the unfused CPU oracle
// (ice_strength excepted: it calls exp()).
#include "evp_aux.cuh"

#include <cmath>

#include "evp_ieee.cuh"

namespace {

constexpr int TPB = 256;
inline unsigned nblk(size_t n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------------
// host block layout <-> slab plane
// ---------------------------------------------------------------------------------------------
struct Loc {
    int i, j, pi, pj;
    bool pad, phys, tne, ighost, jghost;
};

__device__ __forceinline__ Loc locate(const BlockGeom &bg, int b, int cell) {
    Loc L;
    const int *t = bg.tab + b * 6;
    const int ilo = t[0], ihi = t[1], jlo = t[2], jhi = t[3];
    L.i = cell % bg.nx_block + 1;
    L.j = cell / bg.nx_block + 1;
    L.pi = L.i + t[4];
    L.pj = L.j + t[5];
    L.pad = (L.i > ihi + 1) || (L.j > jhi + 1) || (L.i < ilo - 1) || (L.j < jlo - 1);
    L.ighost = (L.i < ilo) || (L.i > ihi);
    L.jghost = (L.j < jlo) || (L.j > jhi);
    L.phys = !L.ighost && !L.jghost;
    L.tne = (L.i >= ilo) && (L.i <= ihi + 1) && (L.j >= jlo) && (L.j <= jhi + 1);
    return L;
}

// a block cell is a source for the plane if it is physical, or a ghost cell that lands on the
// plane's own ghost ring in every direction in which it is a ghost (domain boundary or the row
// owned by the neighbouring slab); interior ghost cells duplicate another block's physical cell.
template <typename TS, typename TD>
__global__ void k_unblock(BlockGeom bg, PlaneGeom pg, const TS *__restrict__ blocked, TD *__restrict__ plane) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (cell >= bg.nx_block * bg.ny_block) return;
    const Loc L = locate(bg, b, cell);
    if (L.pad) return;
    const bool ring_i = (L.pi == 0) || (L.pi == pg.nx + 1);
    const bool ring_j = (L.pj == 0) || (L.pj == pg.nyl + 1);
    if ((L.ighost && !ring_i) || (L.jghost && !ring_j)) return;
    if (L.pi < 0 || L.pi > pg.nx + 1 || L.pj < 0 || L.pj > pg.nyl + 1) return;
    const TS v = blocked[(size_t)b * bg.nx_block * bg.ny_block + cell];
    plane[(size_t)L.pj * pg.pitch + L.pi] = (TD)v;
}

template <typename TS, typename TD>
__global__ void k_block(BlockGeom bg, PlaneGeom pg, const TS *__restrict__ plane,
                        const uint8_t *__restrict__ icetmask, TD *__restrict__ blocked, int policy) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (cell >= bg.nx_block * bg.ny_block) return;
    const Loc L = locate(bg, b, cell);
    if (L.pad) return;
    if (L.pi < 0 || L.pi > pg.nx + 1 || L.pj < 0 || L.pj > pg.nyl + 1) return;
    const size_t pidx = (size_t)L.pj * pg.pitch + L.pi;
    TD *dst = blocked + (size_t)b * bg.nx_block * bg.ny_block + cell;
    const TD v = (TD)plane[pidx];
    switch (policy) {
    case PACK_FULL: *dst = v; break;
    case PACK_TNE_KEEP:
        if (L.tne) *dst = v;
        else if (icetmask[pidx] == 0) *dst = (TD)0;
        break;
    case PACK_TNE_ZERO: *dst = L.tne ? v : (TD)0; break;
    case PACK_INT_ZERO: *dst = L.phys ? v : (TD)0; break;
    case PACK_INT_KEEP:
        if (L.phys) *dst = v;
        break;
    default: break;
    }
}

// ---------------------------------------------------------------------------------------------
// halo updates on a slab plane (serial/ice_boundary.F90:591-873; address lists :3494-4202)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_halo_ew(PlaneGeom pg, T *__restrict__ a) {
    const int j = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j > pg.nyl) return;
    const size_t r = (size_t)j * pg.pitch;
    a[r] = a[r + pg.nx];         // 'east' message: east edge -> west ghost column (:3629-3643)
    a[r + pg.nx + 1] = a[r + 1]; // 'west' message (:3654-3668)
}

template <typename T>
__global__ void k_halo_ns_cyclic(PlaneGeom pg, T *__restrict__ a, int with_corners) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x; // 0..nx+1
    if (i > pg.nx + 1) return;
    if (!with_corners && (i == 0 || i == pg.nx + 1)) return;
    const size_t top = (size_t)pg.nyl * pg.pitch, gn = (size_t)(pg.nyl + 1) * pg.pitch;
    // columns 0 and nx+1 hold the east-west wrap already, so copying them fills the corner cells
    a[i] = a[top + i];                // 'north' message -> south ghost row (:3681-3695)
    a[gn + i] = a[(size_t)pg.pitch + i]; // 'south' message (:3774-3788)
}

// Tripole u-fold for up to 2 planes (blockIdx.x selects).  One CTA per plane: all reads of the
// top physical rows happen before the barrier, all writes after it, so the in-place update of
// row nyl is safe.  loc 1 = centre (ghost row only), 2 = NE corner (symmetrise + overwrite the
// top physical row, :777-800, :837-866).
constexpr int TRIP_MAXC = 8;
// average of two degenerate points: floating point 0.5*(x1 + isign*x2); integer fields nint(...) (serial/
// ice_boundary.F90:1318) -- for the 0/1 masks that travel here nint(0.5*(x1 + x2)) = x1 | x2
template <typename T> __device__ __forceinline__ T trip_avg(T x1, T x2, int isign) { return (T)(0.5 * (x1 + isign * x2)); }
template <> __device__ __forceinline__ uint8_t trip_avg<uint8_t>(uint8_t x1, uint8_t x2, int) { return (uint8_t)(x1 | x2); }

template <typename T>
__global__ void __launch_bounds__(1024) k_halo_tripole(PlaneGeom pg, T *a0, T *a1, int loc, int isign) {
    T *a = blockIdx.x == 0 ? a0 : a1;
    const int nx = pg.nx, nyl = pg.nyl;
    const size_t rtop = (size_t)nyl * pg.pitch, rbelow = (size_t)(nyl - 1) * pg.pitch,
                 rghost = (size_t)(nyl + 1) * pg.pitch;
    T vtop[TRIP_MAXC], vghost[TRIP_MAXC];
    bool wtop = false;
#pragma unroll
    for (int c = 0; c < TRIP_MAXC; ++c) {
        const int i = threadIdx.x + c * 1024; // plane column 0..nx+1
        vtop[c] = (T)0;
        vghost[c] = (T)0;
        if (i > nx + 1) continue;
        int ig = i; // i_glob of the column (source/ice_blocks.F90:291-330)
        if (i == 0) ig = pg.ew_cyclic ? nx : 1;
        if (i == nx + 1) ig = pg.ew_cyclic ? 1 : nx;
        if (pg.tfold) {
            // T-fold (:725-773), iSrc = nx - i_glob + 1 - ioffset.  The three-row buffer the reference works on holds
            // the physical rows nyl-1, nyl, nyl: the 'north' message fills it with rows nyl-2 .. nyl, then the
            // 'northeast' / 'northwest' messages of the tripole blocks -- last in ice_HaloCreate's list, written
            // for the two-row u-fold buffer -- overwrite buffer rows 1 and 2 with rows nyl-1, nyl
            // (serial/ice_boundary.F90:3833-3848; seen by running the reference's own translated halo).
            if (loc == 2) { // NE corner: ioffset 0, joffset 1 -- row nyl <- buffer row 2 (raw row nyl), row nyl+1 <- buffer row 1
                int k = nx - ig + 1;
                if (k == 0) k = nx;
                if (k > nx) k -= nx;
                vtop[c] = (T)(isign * a[rtop + k]);
                vghost[c] = (T)(isign * a[rbelow + k]);
                wtop = true;
            } else { // centre: ioffset -1, joffset 0 -- the top row is degenerate: symmetrise pairs (k, nx-k+2), k = 2..nx/2
                int k = nx - ig + 2;
                if (k == 0) k = nx;
                if (k > nx) k -= nx;
                T x = a[rtop + k];
                if (k >= 2 && k <= nx / 2) {
                    x = trip_avg<T>(a[rtop + k], a[rtop + (nx - k + 2)], isign);
                } else if (k >= nx - nx / 2 + 2 && k <= nx) { // the partner iDst = nx - i + 2 of the loop index i
                    const int kk = nx - k + 2;
                    x = (T)(isign * trip_avg<T>(a[rtop + kk], a[rtop + k], isign));
                }
                vtop[c] = (T)(isign * x);                  // row nyl   <- buffer row 3 (symmetrised row nyl)
                vghost[c] = (T)(isign * a[rtop + k]);      // row nyl+1 <- buffer row 2 (RAW row nyl)
                wtop = true;
            }
        } else if (loc == 2) {
            int k = nx - ig; // iSrc = nxGlobal - i_glob + 1 - ioffset, ioffset = 1
            if (k == 0) k = nx;
            // bufTripole(k, 2) after symmetrisation of the top physical row
            T x;
            if (k >= 1 && k <= nx / 2 - 1) {
                const T x1 = a[rtop + k], x2 = a[rtop + (nx - k)];
                x = (T)(0.5 * (x1 + isign * x2));
            } else if (k >= nx - (nx / 2 - 1) && k <= nx - 1) {
                const int kk = nx - k; // the partner index i of the loop at :793-800
                const T x1 = a[rtop + kk], x2 = a[rtop + k];
                x = (T)(isign * (T)(0.5 * (x1 + isign * x2)));
            } else {
                x = a[rtop + k];
            }
            vtop[c] = (T)(isign * x);                  // j=1: row jhi   <- isign*buf(iSrc, 2)
            vghost[c] = (T)(isign * a[rbelow + k]);    // j=2: row jhi+1 <- isign*buf(iSrc, 1)
            wtop = true;
        } else {
            int k = nx - ig + 1; // ioffset = 0
            if (k > nx) k -= nx;
            vghost[c] = (T)(isign * a[rtop + k]);      // j=2: jSrc = 2; j=1 skipped (jSrc = 3)
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < TRIP_MAXC; ++c) {
        const int i = threadIdx.x + c * 1024;
        if (i > nx + 1) continue;
        if (wtop) a[rtop + i] = vtop[c];
        a[rghost + i] = vghost[c];
    }
}

// ---------------------------------------------------------------------------------------------
// evp_prep1 (source/ice_dyn_evp.F90:643-691)
// ---------------------------------------------------------------------------------------------
__global__ void k_prep1(PlaneGeom pg, PrepArgs a) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pg.cells) return;
    const double a_min = 0.001, m_min = 0.01;
    double tmass = 0.0;
    const bool tm = a.tmask[idx] != 0;
    if (tm) tmass = (a.rhoi * a.vice[idx] + a.rhos * a.vsno[idx]); // :652
    a.tmass[idx] = tmass;
    a.tmphm[idx] = (tm && (a.aice[idx] > a_min) && (tmass > m_min)) ? 1 : 0; // :660
    a.icetmask[idx] = 0; // :674
}

__global__ void k_icetmask(PlaneGeom pg, PrepArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y;
    if (i > pg.nx || j > pg.nyl) return;
    const size_t idx = (size_t)j * pg.pitch + i;
    const uint8_t *m = a.tmphm;
    const size_t p = pg.pitch;
    uint8_t any = m[idx - 1 + p] | m[idx + p] | m[idx + 1 + p] | m[idx - 1] | m[idx] | m[idx + 1] |
                  m[idx - 1 - p] | m[idx - p] | m[idx + 1 - p]; // :683-687
    if (!a.tmask[idx]) any = 0;                                  // :689
    a.icetmask[idx] = any ? 1 : 0;
}

// to_ugrid (source/ice_grid.F90:1612-1631): zero everywhere, 4-point area-weighted mean inside
__global__ void k_to_ugrid(PlaneGeom pg, const double *__restrict__ w1, const double *__restrict__ tarea,
                           const double *__restrict__ uarea, double *__restrict__ w2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= pg.pitch) return;
    const size_t idx = (size_t)j * pg.pitch + i, p = pg.pitch;
    double r = 0.0;
    if (i >= 1 && i <= pg.nx && j >= 1 && j <= pg.nyl)
        r = 0.25 * (w1[idx] * tarea[idx] + w1[idx + 1] * tarea[idx + 1] + w1[idx + p] * tarea[idx + p] +
                    w1[idx + p + 1] * tarea[idx + p + 1]) / uarea[idx];
    w2[idx] = r;
}

// to_tgrid (source/ice_grid.F90:1720-1730): interior only, ghosts of w2 untouched
__global__ void k_to_tgrid(PlaneGeom pg, const double *__restrict__ w1, const double *__restrict__ tarea,
                           const double *__restrict__ uarea, double *__restrict__ w2) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y;
    if (i > pg.nx || j > pg.nyl) return;
    const size_t idx = (size_t)j * pg.pitch + i, p = pg.pitch;
    w2[idx] = 0.25 * (w1[idx] * uarea[idx] + w1[idx - 1] * uarea[idx - 1] + w1[idx - p] * uarea[idx - p] +
                      w1[idx - p - 1] * uarea[idx - p - 1]) / tarea[idx];
}

// evp_prep2 (source/ice_dyn_evp.F90:819-936), dense: the index lists become the masks themselves
__global__ void k_prep2(PlaneGeom pg, PrepArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > pg.nx + 1) return;
    const size_t idx = (size_t)j * pg.pitch + i;
    const double a_min = 0.001, m_min = 0.01;
    double waterx = 0.0, watery = 0.0, forcex = 0.0, forcey = 0.0, umassdtei = 0.0; // :821-825
    if (a.icetmask[idx] == 0) { // :827-840
#pragma unroll
        for (int k = 0; k < EVP_NSTRESS; ++k) a.stress[k][idx] = 0.0;
    }
    if (i >= 1 && i <= pg.nx && j >= 1 && j <= pg.nyl) {
        const bool old = a.iceumask[idx] != 0; // :872-874
        const double aiu = a.aiu[idx], umass = a.umass[idx];
        const bool now = (a.umask[idx] != 0) && (aiu > a_min) && (umass > m_min);
        a.iceumask[idx] = now ? 1 : 0;
        if (now) {
            const double uocn = a.uocn[idx], vocn = a.vocn[idx];
            if (!old) { // :882-885
                a.uvel[idx] = uocn;
                a.vvel[idx] = vocn;
            }
            umassdtei = umass * a.dtei;           // :906
            const double fm = a.fcor[idx] * umass; // :907
            a.fm[idx] = fm;
            if (a.hemisphere_turning) { // :912-913, sign(1., real(fm))
                const double sg = signbit(__double2float_rn(fm)) ? -1.0 : 1.0;
                waterx = uocn * a.cosw - vocn * a.sinw * sg;
                watery = vocn * a.cosw + uocn * a.sinw * sg;
            } else { // :915-916
                waterx = uocn * a.cosw - vocn * a.sinw;
                watery = vocn * a.cosw + uocn * a.sinw;
            }
            double tx, ty;
            if (!a.coupled_tilt) { // :921-922
                tx = -fm * vocn;
                ty = fm * uocn;
            } else { // :924-925
                tx = -a.gravit * umass * a.ss_tltx[idx];
                ty = -a.gravit * umass * a.ss_tlty[idx];
            }
            if (a.hemisphere_turning && !a.use_ocnslope) { // :929-932 (AusCOM)
                tx = -fm * vocn;
                ty = fm * uocn;
            }
            a.strtltx[idx] = tx;
            a.strtlty[idx] = ty;
            forcex = a.strairx[idx] + tx; // :934-935
            forcey = a.strairy[idx] + ty;
        } else { // :888-893
            a.uvel[idx] = 0.0;
            a.vvel[idx] = 0.0;
            a.strintx[idx] = 0.0;
            a.strinty[idx] = 0.0;
            a.strocnx[idx] = 0.0;
            a.strocny[idx] = 0.0;
        }
    }
    a.waterx[idx] = waterx;
    a.watery[idx] = watery;
    a.forcex[idx] = forcex;
    a.forcey[idx] = forcey;
    a.umassdtei[idx] = umassdtei;
}

// evp_finish (source/ice_dyn_evp.F90:1510-1547)
__global__ void k_finish(PlaneGeom pg, FinishArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > pg.nx + 1) return;
    const size_t idx = (size_t)j * pg.pitch + i;
    double xT = 0.0, yT = 0.0; // :1512-1513
    if (i >= 1 && i <= pg.nx && j >= 1 && j <= pg.nyl && a.iceumask[idx]) {
        const double u = a.uvel[idx], v = a.vvel[idx], aiu = a.aiu[idx];
        const double du = a.uocn[idx] - u, dv = a.vocn[idx] - v;
        const double vrel = a.dragw * sqrt(du * du + dv * dv); // :1522
        double sx = a.strocnx[idx], sy = a.strocny[idx];
        if (a.hemisphere_turning && a.fm[idx] < 0.0) { // :1525-1530
            sx = sx - vrel * (u * a.cosw + v * a.sinw) * aiu;
            sy = sy - vrel * (v * a.cosw - u * a.sinw) * aiu;
        } else { // :1532-1541
            sx = sx - vrel * (u * a.cosw - v * a.sinw) * aiu;
            sy = sy - vrel * (v * a.cosw + u * a.sinw) * aiu;
        }
        a.strocnx[idx] = sx;
        a.strocny[idx] = sy;
        xT = sx / aiu; // :1545-1546
        yT = sy / aiu;
    }
    a.strocnxT[idx] = xT;
    a.strocnyT[idx] = yT;
}

// principal_stress (source/ice_dyn_evp.F90:1593-1607)
__global__ void k_principal_stress(size_t n, const double *__restrict__ sp1, const double *__restrict__ sm1,
                                   const double *__restrict__ s12, const double *__restrict__ prs,
                                   double puny, double *__restrict__ sig1, double *__restrict__ sig2) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double r1 = 1.0e30, r2 = 1.0e30; // spval_dbl
    const double p = prs[k];
    if (p > puny) {
        const double rt = sqrt(sm1[k] * sm1[k] + 4.0 * (s12[k] * s12[k]));
        r1 = (0.5 * (sp1[k] + rt)) / p;
        r2 = (0.5 * (sp1[k] - rt)) / p;
    }
    sig1[k] = r1;
    sig2[k] = r2;
}

// ice_strength (source/ice_mechred.F90:1869-2036 with ridge_itd :773-1081) for one T cell of the
// T list (icetmask == 1 on [1..nx+1] x [1..nyl+1]); 0 elsewhere (:1942).
constexpr int MAXCAT = 16;
__global__ void k_ice_strength(PlaneGeom pg, StrengthArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > pg.nx + 1) return;
    const size_t idx = (size_t)j * pg.pitch + i;
    double result = 0.0;
    const int ncat = a.ncat;
    if (a.kstrength == 1) {
        if (i >= 1 && j >= 1 && a.icetmask[idx]) {
            const double puny = a.puny;
            const double Cf = 17.0;
            const double Cp = 0.5 * a.gravit * (a.rhow - a.rhoi) * a.rhoi / a.rhow;
            const double Gstar = 0.15, astar = 0.05, maxraft = 1.0, Hstar = 25.0;
            const double Gstari = 1.0 / Gstar, astari = 1.0 / astar;
            double Gs[MAXCAT + 2], apartic[MAXCAT + 1], hrmin[MAXCAT + 1], hrmax[MAXCAT + 1],
                hrexp[MAXCAT + 1], krdg[MAXCAT + 1];
            double *Gsum = Gs + 1;
            Gsum[-1] = 0.0;
            apartic[0] = 0.0;
            for (int n = 1; n <= ncat; ++n) {
                apartic[n] = 0.0; hrmin[n] = 0.0; hrmax[n] = 0.0; hrexp[n] = 0.0; krdg[n] = 1.0;
            }
            const double a0 = a.aice0[idx];
            Gsum[0] = (a0 > puny) ? a0 : Gsum[-1];
            for (int n = 1; n <= ncat; ++n) {
                const double an = a.aicen[(size_t)(n - 1) * pg.cells + idx];
                Gsum[n] = (an > puny) ? Gsum[n - 1] + an : Gsum[n - 1];
            }
            const double work = 1.0 / Gsum[ncat];
            for (int n = 0; n <= ncat; ++n) Gsum[n] = Gsum[n] * work;
            if (a.krdg_partic == 0) {
                for (int n = 0; n <= ncat; ++n) {
                    if (Gsum[n] < Gstar)
                        apartic[n] = Gstari * (Gsum[n] - Gsum[n - 1]) * (2.0 - (Gsum[n - 1] + Gsum[n]) * Gstari);
                    else if (Gsum[n - 1] < Gstar)
                        apartic[n] = Gstari * (Gstar - Gsum[n - 1]) * (2.0 - (Gsum[n - 1] + Gstar) * Gstari);
                }
            } else {
                const double xtmp = 1.0 / (1.0 - exp(-astari));
                for (int n = -1; n <= ncat; ++n) Gsum[n] = exp(-Gsum[n] * astari) * xtmp;
                for (int n = 0; n <= ncat; ++n) apartic[n] = Gsum[n - 1] - Gsum[n];
            }
            for (int n = 1; n <= ncat; ++n) {
                const double an = a.aicen[(size_t)(n - 1) * pg.cells + idx];
                const double vn = a.vicen[(size_t)(n - 1) * pg.cells + idx];
                if (an > puny) {
                    double hi = vn / an;
                    if (a.krdg_redist == 0) {
                        hrmin[n] = fmin(2.0 * hi, hi + maxraft);
                        hrmax[n] = 2.0 * sqrt(Hstar * hi);
                        hrmax[n] = fmax(hrmax[n], hrmin[n] + puny);
                        const double hrmean = 0.5 * (hrmin[n] + hrmax[n]);
                        krdg[n] = hrmean / hi;
                    } else {
                        hi = fmax(hi, puny);
                        hrmin[n] = fmin(2.0 * hi, hi + maxraft);
                        hrexp[n] = a.mu_rdg * sqrt(hi);
                        krdg[n] = (hrmin[n] + hrexp[n]) / hi;
                    }
                }
            }
            double aksum = apartic[0];
            for (int n = 1; n <= ncat; ++n) aksum = aksum + apartic[n] * (1.0 - 1.0 / krdg[n]);
            double s = 0.0;
            for (int n = 1; n <= ncat; ++n) {
                const double an = a.aicen[(size_t)(n - 1) * pg.cells + idx];
                const double vn = a.vicen[(size_t)(n - 1) * pg.cells + idx];
                if (an > puny && apartic[n] > 0.0) {
                    const double hi = vn / an;
                    double h2rdg;
                    if (a.krdg_redist == 0)
                        h2rdg = (1.0 / 3.0) * (hrmax[n] * hrmax[n] * hrmax[n] - hrmin[n] * hrmin[n] * hrmin[n]) /
                                (hrmax[n] - hrmin[n]);
                    else
                        h2rdg = hrmin[n] * hrmin[n] + 2.0 * hrmin[n] * hrexp[n] + 2.0 * hrexp[n] * hrexp[n];
                    const double dh2rdg = -hi * hi + h2rdg / krdg[n];
                    s = s + apartic[n] * dh2rdg;
                }
            }
            result = Cf * Cp * s / aksum;
        }
    } else {
        if (i >= 1 && i <= pg.nx && j >= 1 && j <= pg.nyl)
            result = 2.75e4 * a.vice[idx] * exp(-20.0 * (1.0 - a.aice[idx]));
    }
    a.strength[idx] = result;
}

// one CTA per row: any mismatching ocean cell clears the row's flag
__global__ void k_check_metrics(PlaneGeom pg, MetricCheckArgs a) {
    const int j = blockIdx.x; // 0..nyl+1
    __shared__ int bad;
    if (threadIdx.x == 0) bad = (j < 1) ? 1 : 0;
    __syncthreads();
    if (j >= 1) {
        for (int i = 1 + threadIdx.x; i <= pg.nx + 1; i += blockDim.x) {
            const size_t idx = (size_t)j * pg.pitch + i;
            if (!a.tmask[idx]) continue;
            const double hte = a.hte[idx], htew = a.hte[idx - 1], htn = a.htn[idx], htns = a.htn[idx - pg.pitch];
            const bool ok = (a.cyp[idx] == 1.5 * hte - 0.5 * htew) && (a.cxp[idx] == 1.5 * htn - 0.5 * htns) &&
                            (a.cym[idx] == -(1.5 * htew - 0.5 * hte)) && (a.cxm[idx] == -(1.5 * htns - 0.5 * htn)) &&
                            (a.dxhy[idx] == 0.5 * (hte - htew)) && (a.dyhx[idx] == 0.5 * (htn - htns)) &&
                            (a.dxt[idx] == 0.5 * (htn + htns)) && (a.dyt[idx] == 0.5 * (hte + htew));
            if (!ok) bad = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.row_ht[j] = bad ? 0 : 1;
}

__global__ void k_row_active(PlaneGeom pg, const uint8_t *__restrict__ icetmask,
                             const uint8_t *__restrict__ iceumask, int *__restrict__ rowcnt) {
    const int j = blockIdx.x; // 0..nyl+1
    int c = 0;
    for (int i = 1 + threadIdx.x; i <= pg.nx + 1; i += blockDim.x) {
        const size_t idx = (size_t)j * pg.pitch + i;
        c += (icetmask[idx] ? 1 : 0) + ((iceumask[idx] && !icetmask[idx]) ? 1 : 0); // rows with T or U work
    }
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) rowcnt[j] = tot;
}

// one CTA: the row costs are staged in shared memory by all threads, then thread 0 walks them (a few
// thousand rows at most; the serial pass over global memory took 110 us at 1080 rows)
__global__ void k_balance_chunks(PlaneGeom pg, const int *__restrict__ rowcnt_g, int *__restrict__ chunks, int ncy,
                                 float w_bot, float w_top, int min_top, float row_overhead, int keep_bot, int keep_top) {
    extern __shared__ int rowcnt[]; // nyl + 2 entries
    for (int j = threadIdx.x; j <= pg.nyl + 1; j += blockDim.x) rowcnt[j] = rowcnt_g[j];
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int nyl = pg.nyl;
    // A row without any active cell costs ~30 % of an active one when a CTA has to march through it (a boundary
    // chunk that keeps its rows, or a gap between two ice bands inside a chunk), and nothing where it is trimmed
    // from the end of a chunk: long ice-free stretches must not be given chunks of their own (they would all be
    // trimmed to nothing while the few chunks left over share the ice).
    auto cost_b = [&](int j) { return row_overhead + (float)rowcnt[j]; };
    auto cost = [&](int j) { return rowcnt[j] > 0 ? row_overhead + (float)rowcnt[j] : 0.02f * row_overhead; };
    if (ncy == 1) {
        chunks[0] = 1; chunks[1] = nyl;
        return;
    }
    float total = 0.f;
    for (int j = 1; j <= nyl; ++j) total += cost(j);
    const int n_int = ncy - 2;
    const float unit = total / ((float)n_int + w_bot + w_top);
    // north chunk: from the top down
    int n_top = 0;
    float acc = 0.f;
    while (n_top < nyl - (ncy - 1) && (n_top < min_top || acc < w_top * unit)) {
        acc += keep_top ? cost_b(nyl - n_top) : cost(nyl - n_top);
        ++n_top;
    }
    if (n_top < 1) n_top = 1;
    // south chunk: from the bottom up, leaving one row for each interior chunk
    int n_bot = 0;
    acc = 0.f;
    while (n_bot < nyl - n_top - n_int && (n_bot < 1 || acc < w_bot * unit)) {
        acc += keep_bot ? cost_b(1 + n_bot) : cost(1 + n_bot);
        ++n_bot;
    }
    chunks[0] = 1; chunks[1] = n_bot;
    chunks[2] = nyl - n_top + 1; chunks[3] = n_top;
    // interior: greedy with a running target so that rounding does not pile up on the last chunk
    int j = 1 + n_bot;
    const int jend = nyl - n_top; // last interior row
    float rest = 0.f;
    for (int r = j; r <= jend; ++r) rest += cost(r);
    for (int k = 0; k < n_int; ++k) {
        const int chunks_left = n_int - k;
        const float target = rest / (float)chunks_left;
        int n = 0;
        acc = 0.f;
        // every remaining chunk must keep at least one row
        while (j + n <= jend - (chunks_left - 1) && (n < 1 || acc + 0.5f * cost(j + n) < target)) {
            acc += cost(j + n);
            ++n;
        }
        if (k == n_int - 1) { // last interior chunk takes what is left
            while (j + n <= jend) { acc += cost(j + n); ++n; }
        }
        chunks[2 * (2 + k)] = j;
        chunks[2 * (2 + k) + 1] = n;
        j += n;
        rest -= acc;
    }
    // Trim rows without any active T or U cell from both ends of the chunks: such rows belong to no chunk
    // (nothing is computed there; velocities and stresses are 0 in both copies), and an all-inactive chunk
    // becomes empty (its CTAs exit at once).  A boundary chunk that carries the tripole fold or the
    // peer-to-peer halo keeps its rows; the northernmost chunk also owns the ghost T row nyl+1.
    for (int k = 0; k < ncy; ++k) {
        if ((k == 0 && keep_bot) || (k == 1 && keep_top)) continue;
        int j0 = chunks[2 * k], n = chunks[2 * k + 1];
        while (n > 0 && rowcnt[j0] == 0) { ++j0; --n; }
        while (n > 0 && rowcnt[j0 + n - 1] == 0 && !(j0 + n - 1 == nyl && rowcnt[nyl + 1] != 0)) --n;
        chunks[2 * k] = j0;
        chunks[2 * k + 1] = n;
    }
}

// non-negative doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    atomicMax((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

__global__ void k_diagnostics(PlaneGeom pg, const double *__restrict__ u, const double *__restrict__ v,
                              const double *__restrict__ strength, const double *__restrict__ fcor,
                              double fcor_south, double *out4) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y;
    double sp[2] = {0.0, 0.0}, pm[2] = {0.0, 0.0};
    if (i <= pg.nx && j <= pg.nyl) {
        const size_t idx = (size_t)j * pg.pitch + i;
        const int south = fcor[idx] < fcor_south ? 1 : 0; // lmask_s: ULAT < -puny
        sp[south] = sqrt(u[idx] * u[idx] + v[idx] * v[idx]);
        pm[south] = fmax(strength[idx], 0.0) / 1000.0;
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 2; ++k) {
            sp[k] = fmax(sp[k], __shfl_down_sync(0xffffffffu, sp[k], o));
            pm[k] = fmax(pm[k], __shfl_down_sync(0xffffffffu, pm[k], o));
        }
    if ((threadIdx.x & 31) == 0) {
        if (sp[0] > 0.0) atomic_max_nonneg(out4 + 0, sp[0]);
        if (sp[1] > 0.0) atomic_max_nonneg(out4 + 1, sp[1]);
        if (pm[0] > 0.0) atomic_max_nonneg(out4 + 2, pm[0]);
        if (pm[1] > 0.0) atomic_max_nonneg(out4 + 3, pm[1]);
    }
}

// ---- kinetic energy / ice and snow volume sums of runtime_diags (source/ice_diagnostics.F90:199-234) ----
// Deterministic, fixed-order sums: row j is summed by one CTA of 256 threads -- thread t adds columns
// t+1, t+257, ... in that order, then a binary tree (stride 128, 64, .. 1) combines the 256 partial sums --
// and one CTA sums the rows the same way.  tests/helpers.py restates exactly this order in numpy.
// q: 0/1 kinetic energy north/south, 2/3 ice volume, 4/5 snow volume (all times the hemisphere's T-cell area).
__device__ __forceinline__ double tree256(double v, double *sh) {
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if (t < s2) sh[t] = sh[t] + sh[t + s2];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) k_energy_rows(PlaneGeom pg, EnergyArgs a) {
    __shared__ double sh[256];
    const int j = 1 + blockIdx.x; // physical rows 1..nyl
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = 1 + threadIdx.x; i <= pg.nx; i += 256) {
        const size_t idx = (size_t)j * pg.pitch + i;
        const double area = a.tmask[idx] ? a.tarea[idx] : 0.0; // tarea * hm (source/ice_grid.F90:1386-1391)
        const int south = a.fcor[idx] < a.fcor_south ? 1 : 0;  // lmask_s: ULAT < -puny
        const double vsno = a.vsno[idx], vice = a.vice[idx], u = a.u[idx], v = a.v[idx];
        const double ke = 0.5 * (a.rhos * vsno + a.rhoi * vice) * (u * u + v * v); // :210-212
        acc[0 + south] = acc[0 + south] + ke * area;
        acc[2 + south] = acc[2 + south] + vice * area;
        acc[4 + south] = acc[4 + south] + vsno * area;
        // the other hemisphere's mask is zero there: adding array * 0 leaves its sum unchanged
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double r = tree256(acc[q], sh);
        if (threadIdx.x == 0) a.rowsum[(size_t)q * (pg.nyl + 2) + j] = r;
    }
}

__global__ void __launch_bounds__(256) k_energy_total(PlaneGeom pg, EnergyArgs a) {
    __shared__ double sh[256];
    for (int q = 0; q < 6; ++q) {
        double acc = 0.0;
        for (int j = 1 + threadIdx.x; j <= pg.nyl; j += 256) acc = acc + a.rowsum[(size_t)q * (pg.nyl + 2) + j];
        const double r = tree256(acc, sh);
        if (threadIdx.x == 0) a.out6[q] = r;
    }
}

// Bounded like the in-kernel waits (wait_flag_ge in evp_subcycle_body.cuh): a neighbour that returned
// early on an error, crashed or timed out must not hang this GPU.  After ~2^24 polls (seconds) the error
// flag sync[6] is raised and the host reports EVP_B200_ERR_STATE.
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_peer_flag(const int *flag, int want, int *sync) {
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(64);
        if (++spins > (1u << 24)) {
            *(volatile int *)(sync + 6) = 1;
            break;
        }
    }
}
__global__ void k_wait_peers(int *sync, int has_north, int has_south, int ncx) {
    const int e = *(volatile int *)(sync + 1);
    for (int x = threadIdx.x; x < ncx; x += blockDim.x) {
        if (has_north) wait_peer_flag(sync + EVP_SYNC_FN + x, e, sync);
        if (has_south) wait_peer_flag(sync + EVP_SYNC_FS + x, e, sync);
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// strip-tiled layout (evp_tiled.cuh): one warp per (strip, tile row); lane = slot
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tile_coords(const PlaneGeom &pg, const TileGeom &tg, int &w, int &j, int &lane, int &i) {
    w = blockIdx.y;
    j = blockIdx.x * blockDim.y + threadIdx.y;
    lane = threadIdx.x;
    i = EVT_UW * w + 1 + lane; // plane column of this slot
    return j < tg.nr;
}

__global__ void k_tile_pack_static(PlaneGeom pg, TileGeom tg, TileStaticArgs a) {
    int w, j, lane, i;
    if (!tile_coords(pg, tg, w, j, lane, i)) return;
    double *inv = tg.tiles + (w * tg.sw + j * tg.sj) + evt_inv_off();
    const bool colT = i <= pg.nx + 1;
    const size_t idx = (size_t)j * pg.pitch + i;
    const double *tp[9] = {a.dxt, a.dyt, a.dxhy, a.dyhx, a.cxp, a.cyp, a.cxm, a.cym, a.tinyarea};
#pragma unroll
    for (int q = 0; q < 9; ++q) inv[EVT_T(1 + q) + lane] = colT ? tp[q][idx] : 0.0;
    const bool colU = lane < EVT_UW && i <= pg.nx && j >= 1;
    inv[EVT_UF(9) + lane] = colU ? a.uarear[idx - pg.pitch] : 0.0;
}

__global__ void k_tile_pack_call(PlaneGeom pg, TileGeom tg, TileCallArgs a) {
    int w, j, lane, i;
    if (!tile_coords(pg, tg, w, j, lane, i)) return;
    double *row = tg.tiles + (w * tg.sw + j * tg.sj);
    double *inv = row + evt_inv_off();
    const bool colT = i <= pg.nx + 1;
    const bool colU = lane < EVT_UW && i <= pg.nx && j >= 1;
    const size_t idx = (size_t)j * pg.pitch + i;
    inv[EVT_T(0) + lane] = colT ? a.strength[idx] : 0.0;
    const double *up[9] = {a.aiu, a.uocn, a.vocn, a.waterx, a.watery, a.forcex, a.forcey, a.umassdtei, a.fm};
#pragma unroll
    for (int q = 0; q < 9; ++q) inv[EVT_UF(q) + lane] = colU ? up[q][idx - pg.pitch] : 0.0;
    unsigned char *mk = (unsigned char *)(inv + EVT_MASK);
    // T cells are computed on [1..nx+1] x [1..nyl+1] (source/ice_dyn_evp.F90:850-859), U cells on the interior
    mk[lane] = (colT && j >= 1 && a.icetmask[idx]) ? 1 : 0;
    mk[32 + lane] = (colU && a.iceumask[idx - pg.pitch]) ? 1 : 0;
    double *c0 = row + evt_state_off(0), *c1 = row + evt_state_off(1);
    const double u = colT ? a.u[idx] : 0.0, v = colT ? a.v[idx] : 0.0;
    c0[EVT_U + lane] = u; c0[EVT_V + lane] = v;
    c1[EVT_U + lane] = u; c1[EVT_V + lane] = v;
#pragma unroll
    for (int k = 0; k < EVP_NSTRESS; ++k) {
        c0[EVT_S(k) + lane] = colT ? a.s[k][idx] : 0.0;
        c1[EVT_S(k) + lane] = 0.0;
    }
    if (lane < 4) { // halo words: U column 31 w (the last column of strip w-1, or the west ghost column)
        const size_t hidx = (size_t)j * pg.pitch + EVT_UW * w;
        const double h = lane == 0 ? a.u[hidx] : (lane == 1 ? a.v[hidx] : 0.0);
        c0[EVT_HALO + lane] = h;
        c1[EVT_HALO + lane] = h;
    }
}

__global__ void k_tile_unpack_state(PlaneGeom pg, TileGeom tg, int copy, TileStateArgs a) {
    int w, j, lane, i;
    if (!tile_coords(pg, tg, w, j, lane, i)) return;
    const double *c = tg.tiles + (w * tg.sw + j * tg.sj) + evt_state_off(copy);
    // every plane column 1 .. nx+1 once: the 31 own slots of a strip, and slot 31 only where it is column nx+1
    const bool own = (lane < EVT_UW && i <= pg.nx + 1) || (i == pg.nx + 1);
    const size_t idx = (size_t)j * pg.pitch + i;
    if (own) {
        a.u[idx] = c[EVT_U + lane];
        a.v[idx] = c[EVT_V + lane];
        if (j >= 1) { // T row 0 is never computed: the planes keep their south ghost stresses
#pragma unroll
            for (int k = 0; k < EVP_NSTRESS; ++k) a.s[k][idx] = c[EVT_S(k) + lane];
        }
    }
    if (w == 0 && lane < 2) { // west ghost column of u, v
        (lane == 0 ? a.u : a.v)[(size_t)j * pg.pitch] = c[EVT_HALO + lane];
    }
}

// ---------------------------------------------------------------------------------------------
// self-test of evp_ieee.cuh against the compiler's own IEEE sqrt() and operator/
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &x) {
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// a double with the given 11-bit exponent field and random sign / mantissa
__device__ __forceinline__ double with_exponent(unsigned long long r, int e) {
    e = e < 0 ? 0 : (e > 2047 ? 2047 : e);
    return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | ((unsigned long long)e << 52)));
}
__device__ __forceinline__ bool same_bits(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}
// out[0] sqrt mismatches (ok but different bits), out[1] sqrt fast-path count, out[2] div mismatches,
// out[3] div fast-path count, out[4] shared-reciprocal (two quotients, one rcp_refined) mismatches,
// out[5] operands tested per function
__global__ void k_selftest_ieee(long long n, unsigned long long seed, unsigned long long *out) {
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        unsigned long long st = seed + 0x632be59bd9b4e019ull * (unsigned long long)(k + 1);
        const unsigned long long r1 = splitmix64(st), r2 = splitmix64(st), r3 = splitmix64(st), r4 = splitmix64(st);
        double a, num, den;
        switch ((int)(k & 7)) {
        case 0: // any bit pattern: all exponents, subnormals, infinities, NaNs, both signs
            a = __longlong_as_double((long long)r1); num = __longlong_as_double((long long)r2);
            den = __longlong_as_double((long long)r3);
            break;
        case 1: // the magnitudes EVP works with (1e-30 .. 1e+30)
        case 2:
            a = with_exponent(r1, 1023 - 100 + (int)(r4 % 200)) ; a = fabs(a);
            num = with_exponent(r2, 1023 - 100 + (int)((r4 >> 8) % 200));
            den = with_exponent(r3, 1023 - 100 + (int)((r4 >> 16) % 200));
            break;
        case 3: // sqrt: around the lower range check (high word 0x03500000) and the upper one (0x7ff00000)
            a = fabs(with_exponent(r1, ((r4 & 1) ? 0x035 : 0x7ff) - 2 + (int)((r4 >> 1) % 5)));
            // division: numerator around the |n| range check (exponent field 0x036)
            num = with_exponent(r2, 0x036 - 2 + (int)((r4 >> 8) % 5));
            den = with_exponent(r3, 1023 - 8 + (int)((r4 >> 16) % 16));
            break;
        case 4: // division: quotient around the tiny-result check (exponent field of q near 1)
            a = fabs(with_exponent(r1, (int)(r4 % 4)));                     // subnormal / smallest normals
            num = with_exponent(r2, 1023 - 400 + (int)((r4 >> 8) % 16));
            den = with_exponent(r3, 1023 + 620 - 8 + (int)((r4 >> 16) % 16));  // q ~ 2^-1020
            break;
        case 5: // huge quotients / overflow, zero and signed-zero operands
            a = (r4 & 3) == 0 ? 0.0 : fabs(with_exponent(r1, 2046 - (int)(r4 % 3)));
            num = (r4 & 12) == 0 ? 0.0 : with_exponent(r2, 2046 - (int)((r4 >> 8) % 40));
            den = (r4 & 48) == 0 ? -0.0 : with_exponent(r3, 1 + (int)((r4 >> 16) % 40));
            break;
        case 6: // perfect squares and exactly representable quotients (ties / exact cases)
            { const double m = (double)(r1 & 0x3ffffff); a = m * m; den = (double)((r3 & 0xfffff) + 1); num = den * (double)(r2 & 0xfffff); }
            break;
        default: // operands one ulp around powers of two
            a = fabs(with_exponent(((r4 & 1) ? 0ull : 0x000fffffffffffffull) ^ (r1 & 3), 1023 - 60 + (int)(r4 % 120)));
            num = with_exponent(((r4 & 2) ? 0ull : 0x000fffffffffffffull) ^ (r2 & 3), 1023 - 60 + (int)((r4 >> 8) % 120));
            den = with_exponent(((r4 & 4) ? 0ull : 0x000fffffffffffffull) ^ (r3 & 3), 1023 - 60 + (int)((r4 >> 16) % 120));
            break;
        }
        bool ok;
        const double sf = evp_ieee::sqrt_fast(a, ok);
        if (ok) { ++c[1]; if (!same_bits(sf, sqrt(a))) ++c[0]; }
        const double r = evp_ieee::rcp_refined(den);
        const double qf = evp_ieee::div_fast(num, den, r, ok);
        if (ok) { ++c[3]; if (!same_bits(qf, num / den)) ++c[2]; }
        // two numerators over one denominator share the refined reciprocal (stepu, :1426-1427)
        bool ok2;
        const double q2 = evp_ieee::div_fast(a, den, r, ok2);
        if (ok2 && !same_bits(q2, a / den)) ++c[4];
    }
    for (int q = 0; q < 5; ++q)
        if (c[q]) atomicAdd(out + q, c[q]);
}

} // namespace

int aux_selftest_ieee(long long n, unsigned long long seed, unsigned long long out[6]) {
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 6 * sizeof(unsigned long long));
    if (e != cudaSuccess) return (int)e;
    cudaMemset(d, 0, 6 * sizeof(unsigned long long));
    k_selftest_ieee<<<148 * 8, 256>>>(n, seed, d);
    e = cudaMemcpy(out, d, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    out[5] = (unsigned long long)n;
    return (int)e;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
void aux_unblock_r8(const BlockGeom &bg, const PlaneGeom &pg, const double *blocked, double *plane, cudaStream_t s) {
    dim3 grid(nblk((size_t)bg.nx_block * bg.ny_block), bg.nblocks);
    k_unblock<double, double><<<grid, TPB, 0, s>>>(bg, pg, blocked, plane);
}
void aux_unblock_mask(const BlockGeom &bg, const PlaneGeom &pg, const int32_t *blocked, uint8_t *plane, cudaStream_t s) {
    dim3 grid(nblk((size_t)bg.nx_block * bg.ny_block), bg.nblocks);
    k_unblock<int32_t, uint8_t><<<grid, TPB, 0, s>>>(bg, pg, blocked, plane);
}
void aux_block_r8(const BlockGeom &bg, const PlaneGeom &pg, const double *plane, const uint8_t *icetmask,
                  double *blocked, int policy, cudaStream_t s) {
    dim3 grid(nblk((size_t)bg.nx_block * bg.ny_block), bg.nblocks);
    k_block<double, double><<<grid, TPB, 0, s>>>(bg, pg, plane, icetmask, blocked, policy);
}
void aux_block_mask(const BlockGeom &bg, const PlaneGeom &pg, const uint8_t *plane, int32_t *blocked,
                    int policy, cudaStream_t s) {
    dim3 grid(nblk((size_t)bg.nx_block * bg.ny_block), bg.nblocks);
    k_block<uint8_t, int32_t><<<grid, TPB, 0, s>>>(bg, pg, plane, nullptr, blocked, policy);
}

void aux_halo_r8(const PlaneGeom &pg, double *plane, int loc, int isign, cudaStream_t s) {
    if (pg.ew_cyclic) k_halo_ew<double><<<nblk(pg.nyl), TPB, 0, s>>>(pg, plane);
    if (pg.ns_cyclic) k_halo_ns_cyclic<double><<<nblk(pg.nx + 2), TPB, 0, s>>>(pg, plane, pg.ew_cyclic);
    if (pg.tripole) k_halo_tripole<double><<<1, 1024, 0, s>>>(pg, plane, plane, loc, isign);
}
void aux_halo_u8(const PlaneGeom &pg, uint8_t *plane, cudaStream_t s) {
    if (pg.ew_cyclic) k_halo_ew<uint8_t><<<nblk(pg.nyl), TPB, 0, s>>>(pg, plane);
    if (pg.ns_cyclic) k_halo_ns_cyclic<uint8_t><<<nblk(pg.nx + 2), TPB, 0, s>>>(pg, plane, pg.ew_cyclic);
    if (pg.tripole) k_halo_tripole<uint8_t><<<1, 1024, 0, s>>>(pg, plane, plane, 1, 1);
}
void aux_halo_uv_ns(const PlaneGeom &pg, double *u, double *v, cudaStream_t s) {
    if (pg.ns_cyclic) {
        k_halo_ns_cyclic<double><<<nblk(pg.nx + 2), TPB, 0, s>>>(pg, u, pg.ew_cyclic);
        k_halo_ns_cyclic<double><<<nblk(pg.nx + 2), TPB, 0, s>>>(pg, v, pg.ew_cyclic);
    }
    if (pg.tripole) k_halo_tripole<double><<<2, 1024, 0, s>>>(pg, u, v, 2, -1);
}

void aux_prep1(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s) {
    k_prep1<<<nblk(pg.cells), TPB, 0, s>>>(pg, a);
}
void aux_icetmask(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s) {
    dim3 grid(nblk(pg.nx), pg.nyl);
    k_icetmask<<<grid, TPB, 0, s>>>(pg, a);
}
void aux_to_ugrid(const PlaneGeom &pg, const double *w1, const double *tarea, const double *uarea,
                  double *w2, cudaStream_t s) {
    dim3 grid(nblk(pg.pitch), pg.nyl + 2);
    k_to_ugrid<<<grid, TPB, 0, s>>>(pg, w1, tarea, uarea, w2);
}
void aux_to_tgrid(const PlaneGeom &pg, const double *w1, const double *tarea, const double *uarea,
                  double *w2, cudaStream_t s) {
    dim3 grid(nblk(pg.nx), pg.nyl);
    k_to_tgrid<<<grid, TPB, 0, s>>>(pg, w1, tarea, uarea, w2);
}
void aux_prep2(const PlaneGeom &pg, const PrepArgs &a, cudaStream_t s) {
    dim3 grid(nblk(pg.nx + 2), pg.nyl + 2);
    k_prep2<<<grid, TPB, 0, s>>>(pg, a);
}
void aux_finish(const PlaneGeom &pg, const FinishArgs &a, cudaStream_t s) {
    dim3 grid(nblk(pg.nx + 2), pg.nyl + 2);
    k_finish<<<grid, TPB, 0, s>>>(pg, a);
}
void aux_principal_stress(size_t n, const double *sp1, const double *sm1, const double *s12,
                          const double *prs, double puny, double *sig1, double *sig2, cudaStream_t s) {
    k_principal_stress<<<nblk(n), TPB, 0, s>>>(n, sp1, sm1, s12, prs, puny, sig1, sig2);
}
void aux_ice_strength(const PlaneGeom &pg, const StrengthArgs &a, cudaStream_t s) {
    dim3 grid(nblk(pg.nx + 2), pg.nyl + 2);
    k_ice_strength<<<grid, TPB, 0, s>>>(pg, a);
}
void aux_tile_pack_static(const PlaneGeom &pg, const TileGeom &tg, const TileStaticArgs &a, cudaStream_t s) {
    dim3 grid((tg.nr + 3) / 4, tg.ns), block(32, 4);
    k_tile_pack_static<<<grid, block, 0, s>>>(pg, tg, a);
}
void aux_tile_pack_call(const PlaneGeom &pg, const TileGeom &tg, const TileCallArgs &a, cudaStream_t s) {
    dim3 grid((tg.nr + 3) / 4, tg.ns), block(32, 4);
    k_tile_pack_call<<<grid, block, 0, s>>>(pg, tg, a);
}
void aux_tile_unpack_state(const PlaneGeom &pg, const TileGeom &tg, int copy, const TileStateArgs &a, cudaStream_t s) {
    dim3 grid((tg.nr + 3) / 4, tg.ns), block(32, 4);
    k_tile_unpack_state<<<grid, block, 0, s>>>(pg, tg, copy, a);
}
void aux_wait_peers(int *sync, int has_north, int has_south, int ncx, cudaStream_t s) {
    k_wait_peers<<<1, 128, 0, s>>>(sync, has_north, has_south, ncx);
}
void aux_check_metrics(const PlaneGeom &pg, const MetricCheckArgs &a, cudaStream_t s) {
    k_check_metrics<<<pg.nyl + 2, 256, 0, s>>>(pg, a);
}
void aux_balance_chunks(const PlaneGeom &pg, const uint8_t *icetmask, const uint8_t *iceumask, int *rowcnt,
                        int *chunks, int ncy, float w_bot, float w_top, int min_top, float row_overhead,
                        int keep_bot, int keep_top, cudaStream_t s) {
    k_row_active<<<pg.nyl + 2, 128, 0, s>>>(pg, icetmask, iceumask, rowcnt);
    k_balance_chunks<<<1, 256, sizeof(int) * (pg.nyl + 2), s>>>(pg, rowcnt, chunks, ncy, w_bot, w_top, min_top,
                                                                 row_overhead, keep_bot, keep_top);
}
void aux_energy_sums(const PlaneGeom &pg, const EnergyArgs &a, cudaStream_t s) {
    k_energy_rows<<<pg.nyl, 256, 0, s>>>(pg, a);
    k_energy_total<<<1, 256, 0, s>>>(pg, a);
}
void aux_diagnostics(const PlaneGeom &pg, const double *u, const double *v, const double *strength,
                     const double *fcor, double fcor_south, double *out4, cudaStream_t s) {
    dim3 grid(nblk(pg.nx), pg.nyl);
    k_diagnostics<<<grid, TPB, 0, s>>>(pg, u, v, strength, fcor, fcor_south, out4);
}

// evp_subcycle_body.cuh -- the fused stress + stepu subcycle kernel (one EVP subcycle per launch).
//
// Included by evp_subcycle_strict.cu (nvcc -fmad=false: unfused IEEE arithmetic, bit-identical to
// the unfused CPU oracle) and evp_subcycle_fast.cu (-fmad=true: nvcc contracts a*b+c into DFMA).
// EVP_SUB_LAUNCH / EVP_PERSIST_LAUNCH name the exported launchers: k_subcycle runs one subcycle per
// launch, k_persist (end of this file) all ndte subcycles in one cooperative launch.
//
// What it replaces: `stress` (source/ice_dyn_evp.F90:947-1293) and `stepu` (:1302-1443) for one
// ksub, plus the east-west part of the two ice_HaloUpdate calls (:397-402).  The reference writes
// str(nx_block,ny_block,8) to memory in `stress` and reads it back in `stepu`; here the eight
// stress combinations never leave the SM:
//
//   * a CTA owns a strip of `strip_w` U columns and marches north over `rows` U rows;
//     thread t holds T column i0+t and U column i0+t (threads 0..strip_w are active, the last
//     one only supplies the T column east of the strip).  DEFAULT since round 2 (WARPX, 128-thread
//     CTAs): the CTA's strip is shared out among its warps, each warp owns strip_w / 4 <= 31 columns
//     of its own (+ the redundant T column east of them) and is autonomous;
//   * for T row j a thread computes the four-corner strain rates / Delta / stress update of its
//     T cell from u,v at (i-1..i, j-1..j) -- west values by an offset load that hits L1, south
//     values carried in registers from the previous row -- and forms str(1:8);
//   * str(2,4,7,8) of the east neighbour arrive by warp shuffle (WARPX: no barrier in the row loop;
//     measured 107.4 vs 113.95 us per launch at 1440 x 1080) or, with one strip per CTA, through a
//     double-buffered shared-memory line (one __syncthreads per row); str(1,5) (+ east 2,7) of the
//     row below are carried in
//     registers, so U(i, j-1) is finished in the same iteration with the reference's summation
//     order  ((s1 + s2) + s3) + s4  and  ((s5 + s6) + s7) + s8  (:1415-1418);
//   * u,v and the 12 stresses are ping-ponged (read `old`, write `new`): the first T row of the
//     CTA above and the T column east of the strip are recomputed redundantly from `old` values,
//     exactly like the reference's redundant N/E ghost-cell stresses (:846-859), so results do
//     not depend on the tiling;
//   * loads for T row j+1 are issued before the arithmetic of row j (software prefetch in
//     registers), the U-row loads before the stress arithmetic whose result they are combined with;
//     mask bytes run one more row ahead.  The load sequences are branch-free on purpose: a run-time
//     branch between them cost 30 % (measured), and staging through shared memory with cp.async
//     (12 instead of 8 warps per SM) was 50 % slower.
//
// Algorithmic traffic per active cell and subcycle: 12+12 stresses, 2+2 velocities, strength +
// 9 T metrics, 10 U fields = 48 fp64 words = 384 B (+2 mask bytes).
#include <cuda_runtime.h>

#include "evp_common.cuh"
#include "evp_ieee.cuh"
#include "evp_tiled.cuh"

namespace EVP_SUB_NS {

// element offsets inside the plane pool fit 32 bits (checked at init): one IMAD.WIDE per address
typedef int idx_t;

// State that other CTAs (persistent kernel) or the neighbour GPUs (peer-to-peer halo) rewrite while
// the kernel runs must not be served by the non-coherent path: COH selects ld.global.cg; the
// one-subcycle kernel on a single rank keeps the read-only path (ld.global.nc).
template <bool COH>
__device__ __forceinline__ double ld_state(const double *p) {
    return COH ? __ldcg(p) : __ldg(p);
}

struct TRow {
    double s[EVP_NSTRESS];
    double strength, dxt, dyt, dxhy, dyhx, cxp, cyp, cxm, cym, tiny, tarear;
    double u, v, uw, vw;
    bool act;
    bool ht; // cyp/cym/cxp/cxm hold raw HTE(i,j), HTE(i-1,j), HTN(i,j), HTN(i,j-1): derive before use
};

// init_grid2 (source/ice_grid.F90:350-361) and primary_grid_lengths (:1196,:1280) for one T cell,
// with explicitly unfused arithmetic so that the values equal the host-computed module arrays.
template <bool HT>
__device__ __forceinline__ void derive_metrics(TRow &t) {
    if (!HT || !t.ht) return;
    const double hte = t.cyp, htew = t.cym, htn = t.cxp, htns = t.cxm;
    t.cyp = __dsub_rn(__dmul_rn(1.5, hte), __dmul_rn(0.5, htew));
    t.cxp = __dsub_rn(__dmul_rn(1.5, htn), __dmul_rn(0.5, htns));
    t.cym = -__dsub_rn(__dmul_rn(1.5, htew), __dmul_rn(0.5, hte));
    t.cxm = -__dsub_rn(__dmul_rn(1.5, htns), __dmul_rn(0.5, htn));
    t.dxhy = __dmul_rn(0.5, __dsub_rn(hte, htew));
    t.dyhx = __dmul_rn(0.5, __dsub_rn(htn, htns));
    t.dxt = __dmul_rn(0.5, __dadd_rn(htn, htns));
    t.dyt = __dmul_rn(0.5, __dadd_rn(hte, htew));
    t.ht = false;
}

struct URow {
    double aiu, uocn, vocn, waterx, watery, forcex, forcey, umassdtei, fm, uarear;
    bool act;
};

// `act` comes from a mask byte that was loaded one iteration earlier, so the data loads below
// are issued without waiting on a dependent mask load.
// `so` / `sn`: offset (in doubles) of the state copy that is read / written; copy 1 of every state plane
// lies a.copy_stride behind copy 0, so one register selects the copy for all 14 planes and the plane
// pointers themselves stay constant-bank operands.
template <bool LAST, bool HT, bool COH>
__device__ __forceinline__ void load_T(const SubArgs &a, idx_t so, TRow &t, int i, int j, bool colT, bool act) {
    idx_t idx = j * a.pitch + i + so;
    t.act = act;
    if (colT) {
        t.u = ld_state<COH>(a.u + idx);
        t.v = ld_state<COH>(a.v + idx);
        t.uw = ld_state<COH>(a.u + idx - 1);
        t.vw = ld_state<COH>(a.v + idx - 1);
    } else {
        t.u = t.v = t.uw = t.vw = 0.0;
    }
    t.ht = false;
    if (t.act) {
#pragma unroll
        for (int k = 0; k < EVP_NSTRESS; ++k) t.s[k] = ld_state<COH>(a.s[k] + idx);
        idx -= so;
        t.strength = __ldg(a.strength + idx);
        if (HT && __ldg(a.row_ht + j)) { // uniform per row
            t.ht = true;
            t.cyp = __ldg(a.hte + idx);
            t.cym = __ldg(a.hte + idx - 1);
            t.cxp = __ldg(a.htn + idx);
            t.cxm = __ldg(a.htn + idx - a.pitch);
        } else {
            t.dxt = __ldg(a.dxt + idx);
            t.dyt = __ldg(a.dyt + idx);
            t.dxhy = __ldg(a.dxhy + idx);
            t.dyhx = __ldg(a.dyhx + idx);
            t.cxp = __ldg(a.cxp + idx);
            t.cyp = __ldg(a.cyp + idx);
            t.cxm = __ldg(a.cxm + idx);
            t.cym = __ldg(a.cym + idx);
        }
        t.tiny = __ldg(a.tinyarea + idx);
        if (LAST) t.tarear = __ldg(a.tarear + idx);
    }
}

__device__ __forceinline__ void load_U(const SubArgs &a, URow &u, int i, int j, bool act) {
    const idx_t idx = j * a.pitch + i;
    u.act = act;
    if (u.act) {
        u.aiu = __ldg(a.aiu + idx);
        u.uocn = __ldg(a.uocn + idx);
        u.vocn = __ldg(a.vocn + idx);
        u.waterx = __ldg(a.waterx + idx);
        u.watery = __ldg(a.watery + idx);
        u.forcex = __ldg(a.forcex + idx);
        u.forcey = __ldg(a.forcey + idx);
        u.umassdtei = __ldg(a.umassdtei + idx);
        u.fm = __ldg(a.fm + idx);
        u.uarear = __ldg(a.uarear + idx);
    }
}

// Where the results of stress_cell / stepu_cell go.  PlaneStore: the plane layout (one pointer per field,
// 32-bit element offsets); TileStore (evp_tiled.cuh): the strip-tiled layout of the TMA-fed kernel.  The
// arithmetic is shared; the policies only differ in addressing.
struct PlaneStoreT {
    const SubArgs &a;
    idx_t sn, idx; // offset of the state copy that is written; plane index of the T cell
    bool on;       // this thread owns the cell (stores its stresses / diagnostics)
    __device__ __forceinline__ void diag(double divu, double rdg_conv, double rdg_shear, double shear) const {
        a.divu[idx] = divu; a.rdg_conv[idx] = rdg_conv; a.rdg_shear[idx] = rdg_shear; a.shear[idx] = shear;
    }
    __device__ __forceinline__ void prs(double v) const { a.prs_sig[idx] = v; }
    __device__ __forceinline__ void stress(int k, double v) const { a.s[k][idx + sn] = v; }
};

struct PlaneStoreU {
    const SubArgs &a;
    idx_t sn, idx;
    int i, j;
    __device__ __forceinline__ void uv(double unew, double vnew) const {
        double *const u_new = a.u + sn, *const v_new = a.v + sn;
        u_new[idx] = unew;
        v_new[idx] = vnew;
        if (a.ew_cyclic) { // east-west part of ice_HaloUpdate(uvel/vvel), serial/ice_boundary.F90:3629-3668
            if (i == a.nx) {
                u_new[idx - a.nx] = unew;
                v_new[idx - a.nx] = vnew;
            }
            if (i == 1) {
                u_new[idx + a.nx] = unew;
                v_new[idx + a.nx] = vnew;
            }
        }
        if (a.p2p) {
            // slab-to-slab part of the halo update: the top / bottom physical row goes straight into the
            // neighbour GPU's ghost row over NVLink (whole padded row: the wrap columns travel with it)
            double *pu = nullptr, *pv = nullptr;
            const bool second = sn != 0; // the neighbours' copies are laid out like ours
            if (j == a.nyl && a.peer_n_flag) {
                pu = a.peer_n_u + (second ? a.peer_n_stride : 0);
                pv = a.peer_n_v + (second ? a.peer_n_stride : 0);
            }
            if (j == 1 && a.peer_s_flag) {
                if (pu) { // one-row slab: both neighbours
                    pu[i] = unew; pv[i] = vnew;
                    if (a.ew_cyclic && i == a.nx) { pu[0] = unew; pv[0] = vnew; }
                    if (a.ew_cyclic && i == 1) { pu[a.nx + 1] = unew; pv[a.nx + 1] = vnew; }
                }
                pu = a.peer_s_u + (second ? a.peer_s_stride : 0);
                pv = a.peer_s_v + (second ? a.peer_s_stride : 0);
            }
            if (pu) {
                pu[i] = unew; pv[i] = vnew;
                if (a.ew_cyclic && i == a.nx) { pu[0] = unew; pv[0] = vnew; }
                if (a.ew_cyclic && i == 1) { pu[a.nx + 1] = unew; pv[a.nx + 1] = vnew; }
            }
        }
    }
    __device__ __forceinline__ void last(double strintx, double strinty, double taux, double tauy) const {
        a.strintx[idx] = strintx; a.strinty[idx] = strinty; a.strocnx[idx] = taux; a.strocny[idx] = tauy;
    }
    // fused evp_finish: not on the row whose velocities the in-kernel tripole fold is still going to change
    __device__ __forceinline__ bool finish_here() const { return a.fuse_finish && !(a.fold && j == a.nyl); }
    __device__ __forceinline__ void finish(double xT, double yT) const { a.finx[idx] = xT; a.finy[idx] = yT; }
};

// n/d1 .. n/d4: four independent IEEE divisions, interleaved (evp_ieee.cuh); bit-identical to operator/
__device__ __forceinline__ void div4(double n, double d1, double d2, double d3, double d4, double &q1,
                                     double &q2, double &q3, double &q4) {
    bool ok1, ok2, ok3, ok4;
    const double r1 = evp_ieee::rcp_refined(d1), r2 = evp_ieee::rcp_refined(d2);
    const double r3 = evp_ieee::rcp_refined(d3), r4 = evp_ieee::rcp_refined(d4);
    q1 = evp_ieee::div_fast(n, d1, r1, ok1);
    q2 = evp_ieee::div_fast(n, d2, r2, ok2);
    q3 = evp_ieee::div_fast(n, d3, r3, ok3);
    q4 = evp_ieee::div_fast(n, d4, r4, ok4);
    if (!(ok1 && ok2 && ok3 && ok4)) {
        q1 = n / d1;
        q2 = n / d2;
        q3 = n / d3;
        q4 = n / d4;
    }
}

// source/ice_dyn_evp.F90:1056-1291 for one T cell.  (un,vn)=(i,j) (uw,vw)=(i-1,j) (us,vs)=(i,j-1)
// (usw,vsw)=(i-1,j-1).  Operation order is the Fortran's.
template <bool LAST, class ST>
__device__ __forceinline__ void stress_cell(const SubArgs &a, const ST &out, const TRow &t, double us, double vs,
                                            double usw, double vsw, double (&str)[8]) {
    const bool store = out.on;
    const double p5 = 0.5, p25 = 0.25, c4 = 4.0;
    const double p166 = 1.0 / 6.0, p333 = 1.0 / 3.0, p111 = 1.0 / 9.0;
    const double p055 = p111 * 0.5, p027 = p055 * 0.5, p222 = 2.0 / 9.0;
    const double un = t.u, vn = t.v, uw = t.uw, vw = t.vw;
    const double cyp = t.cyp, cxp = t.cxp, cym = t.cym, cxm = t.cxm, dxt = t.dxt, dyt = t.dyt;

    // :1065-1072
    const double divune = cyp * un - dyt * uw + cxp * vn - dxt * vs;
    const double divunw = cym * uw + dyt * un + cxp * vw - dxt * vsw;
    const double divusw = cym * usw + dyt * us + cxm * vsw + dxt * vw;
    const double divuse = cyp * us - dyt * usw + cxm * vs + dxt * vn;
    // :1075-1082
    const double tensionne = -cym * un - dyt * uw + cxm * vn + dxt * vs;
    const double tensionnw = -cyp * uw + dyt * un + cxm * vw + dxt * vsw;
    const double tensionsw = -cyp * usw + dyt * us + cxp * vsw - dxt * vw;
    const double tensionse = -cym * us - dyt * usw + cxp * vs - dxt * vn;
    // :1085-1092
    const double shearne = -cym * vn - dyt * vw - cxm * un - dxt * us;
    const double shearnw = -cyp * vw + dyt * vn - cxm * uw - dxt * usw;
    const double shearsw = -cyp * vsw + dyt * vs - cxp * usw + dxt * uw;
    const double shearse = -cym * vs - dyt * vsw - cxp * us + dxt * un;
    // :1095-1098
    // four independent IEEE square roots, interleaved (evp_ieee.cuh); bit-identical to sqrt()
    const double r2ne = divune * divune + a.ecci * (tensionne * tensionne + shearne * shearne);
    const double r2nw = divunw * divunw + a.ecci * (tensionnw * tensionnw + shearnw * shearnw);
    const double r2se = divuse * divuse + a.ecci * (tensionse * tensionse + shearse * shearse);
    const double r2sw = divusw * divusw + a.ecci * (tensionsw * tensionsw + shearsw * shearsw);
    bool ok1, ok2, ok3, ok4;
    double Deltane = evp_ieee::sqrt_fast(r2ne, ok1);
    double Deltanw = evp_ieee::sqrt_fast(r2nw, ok2);
    double Deltase = evp_ieee::sqrt_fast(r2se, ok3);
    double Deltasw = evp_ieee::sqrt_fast(r2sw, ok4);
    if (!(ok1 && ok2 && ok3 && ok4)) {
        Deltane = sqrt(r2ne);
        Deltanw = sqrt(r2nw);
        Deltase = sqrt(r2se);
        Deltasw = sqrt(r2sw);
    }

    if (LAST && store) { // :1103-1115
        const double divu = p25 * (divune + divunw + divuse + divusw) * t.tarear;
        const double tmp = p25 * (Deltane + Deltanw + Deltase + Deltasw) * t.tarear;
        const double tsum = tensionne + tensionnw + tensionse + tensionsw;
        const double ssum = shearne + shearnw + shearse + shearsw;
        out.diag(divu, -fmin(divu, 0.0), p5 * (tmp - fabs(divu)), p25 * t.tarear * sqrt(tsum * tsum + ssum * ssum));
    }

    double c0ne, c0nw, c0sw, c0se;
    if (a.evp_damping) { // :1121-1128
        const double t4 = c4 * t.tiny;
        div4(t.strength, fmax(Deltane, t4), fmax(Deltanw, t4), fmax(Deltasw, t4), fmax(Deltase, t4), c0ne, c0nw,
             c0sw, c0se);
        c0ne = fmin(c0ne, a.rcon);
        c0nw = fmin(c0nw, a.rcon);
        c0sw = fmin(c0sw, a.rcon);
        c0se = fmin(c0se, a.rcon);
        if (LAST && store) out.prs(t.strength * Deltane / fmax(Deltane, t4));
    } else { // :1131-1135
        div4(t.strength, fmax(Deltane, t.tiny), fmax(Deltanw, t.tiny), fmax(Deltasw, t.tiny), fmax(Deltase, t.tiny),
             c0ne, c0nw, c0sw, c0se);
        if (LAST && store) out.prs(c0ne * Deltane);
    }
    const double c1ne = c0ne * a.dte2T; // :1138-1141
    const double c1nw = c0nw * a.dte2T;
    const double c1sw = c0sw * a.dte2T;
    const double c1se = c0se * a.dte2T;

    // :1148-1165
    const double sp1 = (t.s[0] + c1ne * (divune - Deltane)) * a.denom1;
    const double sp2 = (t.s[1] + c1nw * (divunw - Deltanw)) * a.denom1;
    const double sp3 = (t.s[2] + c1sw * (divusw - Deltasw)) * a.denom1;
    const double sp4 = (t.s[3] + c1se * (divuse - Deltase)) * a.denom1;
    const double sm1 = (t.s[4] + c1ne * tensionne) * a.denom2;
    const double sm2 = (t.s[5] + c1nw * tensionnw) * a.denom2;
    const double sm3 = (t.s[6] + c1sw * tensionsw) * a.denom2;
    const double sm4 = (t.s[7] + c1se * tensionse) * a.denom2;
    const double s121 = (t.s[8] + c1ne * shearne * p5) * a.denom2;
    const double s122 = (t.s[9] + c1nw * shearnw * p5) * a.denom2;
    const double s123 = (t.s[10] + c1sw * shearsw * p5) * a.denom2;
    const double s124 = (t.s[11] + c1se * shearse * p5) * a.denom2;

    if (store) {
        out.stress(0, sp1); out.stress(1, sp2); out.stress(2, sp3); out.stress(3, sp4);
        out.stress(4, sm1); out.stress(5, sm2); out.stress(6, sm3); out.stress(7, sm4);
        out.stress(8, s121); out.stress(9, s122); out.stress(10, s123); out.stress(11, s124);
    }

    // :1196-1215
    const double ssigpn = sp1 + sp2, ssigps = sp3 + sp4, ssigpe = sp1 + sp4, ssigpw = sp2 + sp3;
    const double ssigp1 = (sp1 + sp3) * p055, ssigp2 = (sp2 + sp4) * p055;
    const double ssigmn = sm1 + sm2, ssigms = sm3 + sm4, ssigme = sm1 + sm4, ssigmw = sm2 + sm3;
    const double ssigm1 = (sm1 + sm3) * p055, ssigm2 = (sm2 + sm4) * p055;
    const double ssig12n = s121 + s122, ssig12s = s123 + s124, ssig12e = s121 + s124, ssig12w = s122 + s123;
    const double ssig121 = (s121 + s123) * p111, ssig122 = (s122 + s124) * p111;
    // :1217-1234
    const double csigpne = p111 * sp1 + ssigp2 + p027 * sp3;
    const double csigpnw = p111 * sp2 + ssigp1 + p027 * sp4;
    const double csigpsw = p111 * sp3 + ssigp2 + p027 * sp1;
    const double csigpse = p111 * sp4 + ssigp1 + p027 * sp2;
    const double csigmne = p111 * sm1 + ssigm2 + p027 * sm3;
    const double csigmnw = p111 * sm2 + ssigm1 + p027 * sm4;
    const double csigmsw = p111 * sm3 + ssigm2 + p027 * sm1;
    const double csigmse = p111 * sm4 + ssigm1 + p027 * sm2;
    const double csig12ne = p222 * s121 + ssig122 + p055 * s123;
    const double csig12nw = p222 * s122 + ssig121 + p055 * s124;
    const double csig12sw = p222 * s123 + ssig122 + p055 * s121;
    const double csig12se = p222 * s124 + ssig121 + p055 * s122;
    // :1236-1239
    const double str12ew = p5 * dxt * (p333 * ssig12e + p166 * ssig12w);
    const double str12we = p5 * dxt * (p333 * ssig12w + p166 * ssig12e);
    const double str12ns = p5 * dyt * (p333 * ssig12n + p166 * ssig12s);
    const double str12sn = p5 * dyt * (p333 * ssig12s + p166 * ssig12n);
    // :1244-1264
    double strp_tmp = p25 * dyt * (p333 * ssigpn + p166 * ssigps);
    double strm_tmp = p25 * dyt * (p333 * ssigmn + p166 * ssigms);
    str[0] = -strp_tmp - strm_tmp - str12ew + t.dxhy * (-csigpne + csigmne) + t.dyhx * csig12ne;
    str[1] = strp_tmp + strm_tmp - str12we + t.dxhy * (-csigpnw + csigmnw) + t.dyhx * csig12nw;
    strp_tmp = p25 * dyt * (p333 * ssigps + p166 * ssigpn);
    strm_tmp = p25 * dyt * (p333 * ssigms + p166 * ssigmn);
    str[2] = -strp_tmp - strm_tmp + str12ew + t.dxhy * (-csigpse + csigmse) + t.dyhx * csig12se;
    str[3] = strp_tmp + strm_tmp + str12we + t.dxhy * (-csigpsw + csigmsw) + t.dyhx * csig12sw;
    // :1269-1289
    strp_tmp = p25 * dxt * (p333 * ssigpe + p166 * ssigpw);
    strm_tmp = p25 * dxt * (p333 * ssigme + p166 * ssigmw);
    str[4] = -strp_tmp + strm_tmp - str12ns - t.dyhx * (csigpne + csigmne) + t.dxhy * csig12ne;
    str[5] = strp_tmp - strm_tmp - str12sn - t.dyhx * (csigpse + csigmse) + t.dxhy * csig12se;
    strp_tmp = p25 * dxt * (p333 * ssigpw + p166 * ssigpe);
    strm_tmp = p25 * dxt * (p333 * ssigmw + p166 * ssigme);
    str[6] = -strp_tmp + strm_tmp + str12ns - t.dyhx * (csigpnw + csigmnw) + t.dxhy * csig12nw;
    str[7] = strp_tmp - strm_tmp + str12sn - t.dyhx * (csigpsw + csigmsw) + t.dxhy * csig12sw;
}

// evp_finish (source/ice_dyn_evp.F90:1520-1546) for one U cell of the U list: (u, v) are the final velocities (after
// the halo update of the last subcycle), sx / sy come in as taux / tauy of the last stepu (:1434-1435) and leave as
// strocnx / strocny; xT / yT = strocnx / aiu, strocny / aiu (the input of u2tgrid_vector).  Same expressions as k_finish.
__device__ __forceinline__ void finish_cell(const SubArgs &a, double u, double v, double uocn, double vocn, double aiu,
                                            double fm, double &sx, double &sy, double &xT, double &yT) {
    const double du = uocn - u, dv = vocn - v;
    const double vrel = a.dragw * sqrt(du * du + dv * dv); // :1522
    if (a.hemisphere_turning && fm < 0.0) {                 // :1525-1530
        sx = sx - vrel * (u * a.cosw + v * a.sinw) * aiu;
        sy = sy - vrel * (v * a.cosw - u * a.sinw) * aiu;
    } else {                                                // :1532-1541
        sx = sx - vrel * (u * a.cosw - v * a.sinw) * aiu;
        sy = sy - vrel * (v * a.cosw + u * a.sinw) * aiu;
    }
    xT = sx / aiu; // :1545-1546
    yT = sy / aiu;
}

// evp_finish for U column c of the top physical row, by the thread(s) that applied the in-kernel tripole fold to it
// (the fold changes u, v of that row after stepu): strocnx / strocny still hold the raw taux / tauy of the last stepu.
__device__ __forceinline__ void finish_fold_column(const SubArgs &a, int c, double u, double v) {
    const size_t idx = (size_t)a.nyl * a.pitch + c;
    if (__ldg(a.iceumask + idx) == 0) return;
    double sx = __ldcg(a.strocnx + idx), sy = __ldcg(a.strocny + idx), xT, yT;
    finish_cell(a, u, v, __ldg(a.uocn + idx), __ldg(a.vocn + idx), __ldg(a.aiu + idx), __ldg(a.fm + idx), sx, sy, xT, yT);
    a.strocnx[idx] = sx;
    a.strocny[idx] = sy;
    a.finx[idx] = xT;
    a.finy[idx] = yT;
}

// source/ice_dyn_evp.F90:1386-1441 for one U cell; sx, sy are the two str sums of :1415-1418
template <bool LAST, class SU>
__device__ __forceinline__ void stepu_cell(const SubArgs &a, const SU &out, const URow &u, double uold, double vold,
                                           double sx, double sy) {
    const double du = u.uocn - uold, dv = u.vocn - vold;
    const double vrel = u.aiu * a.dragw * sqrt(du * du + dv * dv); // :1394
    const double taux = vrel * u.waterx;                            // :1397-1398
    const double tauy = vrel * u.watery;
    const double cca = u.umassdtei + vrel * a.cosw;                 // :1401
    double ccb;
    if (a.hemisphere_turning && u.fm < 0.0)                         // :1403-1410
        ccb = u.fm - vrel * a.sinw;
    else
        ccb = u.fm + vrel * a.sinw;
    const double ab2 = cca * cca + ccb * ccb;                       // :1412
    const double strintx = u.uarear * sx;                           // :1415-1418
    const double strinty = u.uarear * sy;
    const double cc1 = strintx + u.forcex + taux + u.umassdtei * uold; // :1421-1424
    const double cc2 = strinty + u.forcey + tauy + u.umassdtei * vold;
    // :1426-1427: two IEEE divisions by the same denominator share the refined reciprocal
    const double nu = cca * cc1 + ccb * cc2, nv = cca * cc2 - ccb * cc1;
    bool oku, okv;
    const double rab = evp_ieee::rcp_refined(ab2);
    double unew = evp_ieee::div_fast(nu, ab2, rab, oku);
    double vnew = evp_ieee::div_fast(nv, ab2, rab, okv);
    if (!(oku && okv)) {
        unew = nu / ab2;
        vnew = nv / ab2;
    }
    out.uv(unew, vnew); // + the east-west / slab-to-slab part of ice_HaloUpdate(uvel/vvel) (:397-402)
    if constexpr (LAST) { // only the last subcycle's values are observable (:1415-1418,:1434-1435)
        double ox = taux, oy = tauy;
        if (out.finish_here()) { // evp_finish as an epilogue of the thread that holds the final u, v
            double xT, yT;
            finish_cell(a, unew, vnew, u.uocn, u.vocn, u.aiu, u.fm, ox, oy, xT, yT);
            out.finish(xT, yT);
        }
        out.last(strintx, strinty, ox, oy);
    }
}

// Bounded spin on a flag that another CTA (or another GPU) advances: a protocol bug or a dead peer must
// not hang the device.  After ~2^26 polls the error flag sync[6] is raised and the wait gives up; the
// host reports EVP_B200_ERR_STATE after the loop.
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag_ge(const int *flag, int want, int *sync) {
    unsigned spins = 0;
    while (ld_acquire(flag) < want) {
        if (++spins > (1u << 26)) {
            *(volatile int *)(sync + 6) = 1;
            break;
        }
    }
}

// Start of a subcycle on a boundary CTA of a multi-rank run (peer-to-peer halo): wait until strips
// x-1, x, x+1 of the neighbour's adjacent chunk have published `e` finished subcycles.
// A boundary CTA reads ghost-row columns that those strips stored during that rank's previous
// subcycle, and stores into ghost-row columns those strips read during it.  Both are safe once they
// have published at least as many finished subcycles as this rank has completed (per-strip epochs:
// the neighbour's boundary CTAs run first and are short, so normally nothing waits).
__device__ __forceinline__ void p2p_wait(const SubArgs &a, int tid, bool top, bool bot, int e) {
    if (tid < 3) {
        const int ncx = (int)gridDim.x;
        const int x = ((int)blockIdx.x + tid - 1 + ncx) % ncx;
        if (top && a.peer_n_flag) wait_flag_ge(a.sync + EVP_SYNC_FN + x, e, a.sync);
        if (bot && a.peer_s_flag) wait_flag_ge(a.sync + EVP_SYNC_FS + x, e, a.sync);
        __threadfence_system();
    }
    __syncthreads();
}

// Tripole u-fold of the rows nyl (symmetrised in place) and nyl+1 (ghost row) of the planes u_new, v_new by ONE CTA,
// for the whole width (north-south part of the halo update on the top slab, serial/ice_boundary.F90:777-866).
// The raw top row goes through a scratch copy because the update is in place.  All other CTAs' stores to these
// rows must be complete and visible (the caller orders that).
template <int NT>
__device__ __forceinline__ void fold_top_rows(const SubArgs &a, double *u_new, double *v_new, int tid, bool finish = false) {
    const size_t rtop = (size_t)a.nyl * a.pitch;
    const int ncol = a.nx + 2;
    // raw top row of u_new and v_new -> scratch (independent loads, issued in batches)
    for (int c0 = 0; c0 < 2 * ncol; c0 += 4 * NT) {
        double val[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * NT + tid;
            if (c < 2 * ncol) {
                const int f = c >= ncol, cc = f ? c - ncol : c;
                val[q] = __ldcg((f ? v_new : u_new) + rtop + cc);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * NT + tid;
            if (c < 2 * ncol) {
                const int f = c >= ncol, cc = f ? c - ncol : c;
                a.fold_scratch[(size_t)f * a.pitch + cc] = val[q];
            }
        }
    }
    __syncthreads();
    for (int c0 = 0; c0 < 2 * ncol; c0 += 4 * NT) {
        double vt[4], vg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * NT + tid;
            if (c < 2 * ncol) {
                const int f = c >= ncol, cc = f ? c - ncol : c;
                const double *fld = f ? v_new : u_new;
                int ig = cc;
                if (cc == 0) ig = a.ew_cyclic ? a.nx : 1;
                if (cc == a.nx + 1) ig = a.ew_cyclic ? 1 : a.nx;
                int k = a.nx - ig;
                if (k == 0) k = a.nx;
                const double *scr = a.fold_scratch + (size_t)f * a.pitch;
                double dummy;
                evp_fold_necorner(scr, scr, cc, a.nx, a.ew_cyclic, -1.0, vt[q], dummy);
                vg[q] = -__ldcg(fld + rtop - a.pitch + k);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * NT + tid;
            if (c < 2 * ncol) {
                const int f = c >= ncol, cc = f ? c - ncol : c;
                double *fld = f ? v_new : u_new;
                fld[rtop + cc] = vt[q];
                fld[rtop + a.pitch + cc] = vg[q];
            }
        }
    }
    if (finish) { // last subcycle: evp_finish of the row whose velocities have just been folded
        __syncthreads();
        for (int c = 1 + tid; c <= a.nx; c += NT) finish_fold_column(a, c, u_new[rtop + c], v_new[rtop + c]);
    }
}

// End of a subcycle: publication of this rank's epoch to the neighbours (peer-to-peer halo), then the
// tripole fold by the last CTA of the northernmost chunk (top slab).  `sn` = offset of the copy just written.
// PERSIST: the fold's completion is published in sync[5] (the top chunk waits for it before the next
// subcycle) and the rank-level counter is left to the end of the persistent kernel.
template <int NT, bool PERSIST>
__device__ __forceinline__ void subcycle_epilogue(const SubArgs &a, idx_t sn, int tid, bool top, bool bot, int epoch,
                                                  bool last = false) {
    double *const u_new = a.u + sn, *const v_new = a.v + sn;
    if (a.p2p && ((top && a.peer_n_flag) || (bot && a.peer_s_flag))) {
        // boundary CTA done: its stores into the neighbour's ghost row are made visible system-wide,
        // then its per-strip epoch is published in the neighbour's sync block
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            if (top && a.peer_n_flag) *(volatile int *)(a.peer_n_flag + blockIdx.x) = epoch + 1;
            if (bot && a.peer_s_flag) *(volatile int *)(a.peer_s_flag + blockIdx.x) = epoch + 1;
        }
    }
    if (a.fold && top) {
        // Tripole u-fold (north-south part of the halo update on the top slab): the CTAs of the
        // northernmost chunk hold rows nyl-1 and nyl; the last of them to finish symmetrises the
        // top row and fills the ghost row for the whole width.  The raw top row goes through a
        // scratch copy because the update is in place.
        __shared__ int is_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            is_last = (atomicAdd((unsigned *)a.sync + 4, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            fold_top_rows<NT>(a, u_new, v_new, tid, last && a.fuse_finish);
            if (PERSIST) {
                __syncthreads();
                if (tid == 0) {
                    a.sync[4] = 0;
                    __threadfence();
                    *(volatile int *)(a.sync + 5) = epoch + 1;
                }
            } else if (tid == 0) {
                a.sync[4] = 0;
            }
        }
    }
    if (!PERSIST && a.p2p) {
        // the last CTA of the grid to finish advances this rank's count of completed kernels
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned total = gridDim.x * gridDim.y;
            const unsigned prev = atomicAdd((unsigned *)a.sync, 1u);
            if (prev == total - 1) {
                a.sync[0] = 0;
                a.sync[1] = a.sync[1] + 1;
            }
        }
    }
}

// One subcycle of one CTA: march north over the U rows j0 .. j0+nrows-1 of strip blockIdx.x, reading
// the state copy at offset `so` and writing the one at offset `sn`.  COH: state loads bypass the
// non-coherent L1 (persistent kernel).
// LATE: the loads of T row j+1 are issued after the arithmetic of row j instead of before it: nothing is
// prefetched across the stress computation, which frees ~50 registers (three warps per scheduler fit)
// at the price of an exposed load latency per row that the third warp has to hide.
// WARPX: every WARP owns its own strip of strip_w / 4 <= 31 U columns (lane strip_w / 4 supplies the T column east
// of it, recomputed redundantly like the column east of a CTA strip) and takes str(2,4,7,8) of the east neighbour by
// warp shuffle: no exchange line, no CTA barrier in the row loop, the warps of a CTA drift apart freely.
template <int NT, bool LAST, bool HT, bool COH, bool LATE = false, bool WARPX = false>
__device__ __forceinline__ void march(const SubArgs &a, idx_t so, idx_t sn, int tid, int i, int j0, int nrows) {
    extern __shared__ double evp_xch[]; // [2][4][NT]: str(2,4,7,8) handed to the west neighbour thread (not WARPX)
    const int jlast = min(j0 + nrows, a.nyl + 1); // last T row of this CTA
    const int lt = WARPX ? (tid & 31) : tid;              // position inside the strip of this thread's owner
    const int lw = WARPX ? a.strip_w / (NT / 32) : a.strip_w; // U columns of that strip
    const bool colT = (lt <= lw) && (i <= a.nx + 1);
    const bool colU = (lt < lw) && (i <= a.nx);
    const bool ownT = colT && (lt < lw || i == a.nx + 1);

    double us = 0.0, vs = 0.0, usw = 0.0, vsw = 0.0;
    if (colT) {
        const idx_t idx = (j0 - 1) * a.pitch + i + so;
        us = ld_state<COH>(a.u + idx);
        vs = ld_state<COH>(a.v + idx);
        usw = ld_state<COH>(a.u + idx - 1);
        vsw = ld_state<COH>(a.v + idx - 1);
    }
    double px = 0.0, s5c = 0.0, s7c = 0.0;
    // Mask bytes run two rows (T) / one row (U) ahead of the data they gate and are kept RAW in a
    // register: the compare that consumes a byte sits one iteration after its load, so no data load
    // ever waits on a mask load.
    const uint8_t *tmk = a.icetmask + (size_t)j0 * a.pitch + i;
    const uint8_t *umk = a.iceumask + (size_t)j0 * a.pitch + i;
    unsigned tm_raw = (colT && j0 + 1 <= jlast) ? __ldg(tmk + a.pitch) : 0u; // T row j+1
    unsigned um_raw = 0u;                                                     // U row j-1
    TRow t;
    load_T<LAST, HT, COH>(a, so, t, i, j0, colT, colT && (__ldg(tmk) != 0));
    int par = 0;

    for (int j = j0; j <= jlast; ++j) {
        TRow tn;
        URow uc;
        const bool tm_next = tm_raw != 0u;
        const bool um_cur = um_raw != 0u;
        const unsigned tm_raw2 = (colT && j + 2 <= jlast) ? __ldg(tmk + 2 * (size_t)a.pitch) : 0u;
        const unsigned um_raw1 = (colU && j + 1 <= jlast) ? __ldg(umk) : 0u; // U row j
        tmk += a.pitch;
        umk += a.pitch;
        if (!LATE && j < jlast) load_T<LAST, HT, COH>(a, so, tn, i, j + 1, colT, tm_next); // software prefetch of the next row
        uc.act = false;
        if (j > j0) load_U(a, uc, i, j - 1, um_cur);

        const idx_t idx = j * a.pitch + i;
        double str[8];
        if (t.act) {
            derive_metrics<HT>(t);
            const bool store = ownT && (j < j0 + nrows || j == a.nyl + 1);
            stress_cell<LAST>(a, PlaneStoreT{a, sn, idx, store}, t, us, vs, usw, vsw, str);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) str[k] = 0.0; // str(:,:,:) = c0, :1051
        }
        double s2r = 0.0, s4r = 0.0, s7r = 0.0, s8r = 0.0;
        if (WARPX) { // str(2,4,7,8) of the T cell east of this lane's U point (lane 31 never holds a U column)
            s2r = __shfl_down_sync(0xffffffffu, str[1], 1);
            s4r = __shfl_down_sync(0xffffffffu, str[3], 1);
            s7r = __shfl_down_sync(0xffffffffu, str[6], 1);
            s8r = __shfl_down_sync(0xffffffffu, str[7], 1);
        } else {
            double *const xl = evp_xch + par * 4 * NT + tid;
            xl[0 * NT] = str[1];
            xl[1 * NT] = str[3];
            xl[2 * NT] = str[6];
            xl[3 * NT] = str[7];
            __syncthreads();
            if (tid < NT - 1) {
                s2r = xl[0 * NT + 1];
                s4r = xl[1 * NT + 1];
                s7r = xl[2 * NT + 1];
                s8r = xl[3 * NT + 1];
            }
            par ^= 1;
        }

        if (uc.act) {
            const double sx = px + str[2] + s4r;          // ((s1 + s2) + s3) + s4
            const double sy = s5c + str[5] + s7c + s8r;    // ((s5 + s6) + s7) + s8
            stepu_cell<LAST>(a, PlaneStoreU{a, sn, idx - a.pitch, i, j - 1}, uc, us, vs, sx, sy);
        }
        px = str[0] + s2r;
        s5c = str[4];
        s7c = s7r;
        us = t.u;
        vs = t.v;
        usw = t.uw;
        vsw = t.vw;
        if (LATE && j < jlast) load_T<LAST, HT, COH>(a, so, tn, i, j + 1, colT, tm_next);
        t = tn;
        tm_raw = tm_raw2;
        um_raw = um_raw1;
    }
}

// At ~210 registers per thread every scheduler (16384 registers) holds two warps: 8 warps per SM
// whatever the CTA shape (96 x 3 or 160 x 2 would need <= 168 registers and spill).
// COH: u, v and the stresses are read with ld.global.cg instead of the read-only (non-coherent) path.
// Needed with the peer-to-peer halo: the neighbour GPU stores into this slab's ghost rows while this
// kernel may already be running (the boundary CTAs only wait on the flag before READING those rows), and
// ld.global.nc requires the data to be read-only for the whole kernel.
template <int NT, bool LAST, bool HT, bool LATE = false, int MINB = 1, bool COH = false, bool WARPX = false>
__global__ void __launch_bounds__(NT, MINB) k_subcycle(const __grid_constant__ SubArgs a) {
    // programmatic dependent launch: let the next subcycle kernel be scheduled as SMs drain, and
    // wait here until the previous grid has completed and flushed (no-ops without the attribute)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    // WARPX: the CTA's strip_w columns are shared out among its warps, strip_w / (NT / 32) <= 31 to each
    const int i = WARPX ? 1 + blockIdx.x * a.strip_w + (tid >> 5) * (a.strip_w / (NT / 32)) + (tid & 31)
                        : 1 + blockIdx.x * a.strip_w + tid;
    // Row chunks come from a table in launch order: the southernmost and northernmost chunk first
    // (with the peer-to-peer halo the rows the neighbours wait for are produced first), and they are
    // shorter than the interior ones, so that the wait for the neighbours / the tripole fold at their
    // end overlaps with the interior CTAs instead of extending the kernel.
    const int j0 = __ldg(a.chunks + 2 * blockIdx.y);
    const int nrows = __ldg(a.chunks + 2 * blockIdx.y + 1);
    const bool top = (j0 + nrows - 1 == a.nyl), bot = (j0 == 1);
    int epoch = 0;
    if (a.p2p && ((top && a.peer_n_flag) || (bot && a.peer_s_flag))) {
        epoch = *(volatile int *)(a.sync + 1); // subcycle kernels completed on this rank
        p2p_wait(a, tid, top, bot, epoch);
    }
    // an empty chunk (all rows inactive, trimmed by the load balancer) has nothing to do
    const idx_t so = a.flip ? (idx_t)a.copy_stride : 0, sn = a.flip ? 0 : (idx_t)a.copy_stride;
    if (nrows > 0) march<NT, LAST, HT, COH, LATE, WARPX>(a, so, sn, tid, i, j0, nrows);
    subcycle_epilogue<NT, false>(a, sn, tid, top, bot, epoch, LAST);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant of the one-subcycle kernel (kernel_variant bits 8/9).  The 24 (LAST: 25) planes a
// T row needs -- u, v, 12 stresses, strength, 9 metrics (+ tarear) -- are brought into shared memory
// by one thread per CTA with 1-D bulk copies (cp.async.bulk, one 16-byte-aligned row segment of
// strip_w + 2 columns per plane) that complete on an mbarrier, S rows deep, instead of being
// prefetched one row ahead in registers by every thread.  That frees ~54 registers per thread (three
// warps per scheduler fit) and lengthens the prefetch distance from one row to S - 1.  Whole row
// segments are copied whether or not their cells are active; the U-point fields, the masks and all
// stores stay as in march().  The arithmetic is the same code (stress_cell / stepu_cell): results are
// bit-identical.
// ------------------------------------------------------------------------------------------------
template <bool LAST> struct TmaPlanes { static constexpr int N = LAST ? 25 : 24; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Warp 0 issues one row: lane 0 arms the stage's mbarrier with the byte count, then lanes 0 .. NP-1 issue
// one bulk copy each (plane `lane` of SubArgs::tplane; the first 14 are state planes and take the copy
// offset).  Issued by a single thread the copies cost ~250 serial instructions per row and made warp 0
// the straggler of every row barrier.
template <bool LAST, int W>
__device__ __forceinline__ void tma_issue_row(const SubArgs &a, int lane, idx_t so, idx_t goff, unsigned bytes,
                                              double *stage, unsigned bar) {
    constexpr int NP = TmaPlanes<LAST>::N;
    if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (unsigned)NP)
                     : "memory");
    __syncwarp();
    if (lane < NP) {
        const double *src = a.tplane[lane] + (lane < 2 + EVP_NSTRESS ? so : 0) + goff;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(stage + lane * W)),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
    }
}

template <int NT, bool LAST, int S>
__device__ __forceinline__ void march_tma(const SubArgs &a, idx_t so, idx_t sn, int tid, int i, int j0, int nrows) {
    constexpr int W = NT + 2;                 // doubles per staged plane row (16-byte multiple)
    constexpr int NP = TmaPlanes<LAST>::N;
    extern __shared__ double evp_xch[];       // [2][4][NT] exchange line, then S stages of NP x W, then S mbarriers
    double *const stages = evp_xch + 2 * 4 * NT;
    unsigned long long *const bars = (unsigned long long *)(stages + S * NP * W);
    const int jlast = min(j0 + nrows, a.nyl + 1); // last T row of this CTA
    const int nT = jlast - j0 + 1;
    const bool colT = (tid <= a.strip_w) && (i <= a.nx + 1);
    const bool colU = (tid < a.strip_w) && (i <= a.nx);
    const bool ownT = colT && (tid < a.strip_w || i == a.nx + 1);
    // staged segment: plane columns c0 .. c0 + strip_w + 1, c0 = i0 - 1 (even: strip_w is even)
    const int c0 = blockIdx.x * a.strip_w;
    const unsigned bytes = (unsigned)(a.strip_w + 2) * 8u;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < S; ++q)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + q)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
#pragma unroll
        for (int q = 0; q < S; ++q)
            if (q < nT)
                tma_issue_row<LAST, W>(a, tid, so, (idx_t)(j0 + q) * a.pitch + c0, bytes, stages + q * NP * W,
                                       smem_u32(bars + q));
    }

    double us = 0.0, vs = 0.0, usw = 0.0, vsw = 0.0;
    if (colT) {
        const idx_t idx = (j0 - 1) * a.pitch + i + so;
        us = __ldcg(a.u + idx);
        vs = __ldcg(a.v + idx);
        usw = __ldcg(a.u + idx - 1);
        vsw = __ldcg(a.v + idx - 1);
    }
    double px = 0.0, s5c = 0.0, s7c = 0.0;
    const uint8_t *tmk = a.icetmask + (size_t)j0 * a.pitch + i;
    const uint8_t *umk = a.iceumask + (size_t)j0 * a.pitch + i;
    unsigned tm_raw = colT ? __ldg(tmk) : 0u; // T row j (one row ahead of its use is enough here)
    unsigned um_raw = 0u;                     // U row j-1
    int par = 0, st = 0;
    unsigned phase = 0;

    for (int r = 0; r < nT; ++r) {
        const int j = j0 + r;
        URow uc;
        const bool tact = tm_raw != 0u;
        const bool um_cur = um_raw != 0u;
        const unsigned tm_raw1 = (colT && j + 1 <= jlast) ? __ldg(tmk + a.pitch) : 0u; // T row j+1
        const unsigned um_raw1 = (colU && j + 1 <= jlast) ? __ldg(umk) : 0u;           // U row j
        tmk += a.pitch;
        umk += a.pitch;
        uc.act = false;
        if (j > j0) load_U(a, uc, i, j - 1, um_cur);

        // wait for the row's bulk copies (bounded: a protocol bug must not hang the device)
        {
            const unsigned bar = smem_u32(bars + st);
            unsigned done = 0, spins = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done)
                             : "r"(bar), "r"(phase)
                             : "memory");
                if (!done && ++spins > (1u << 22)) {
                    *(volatile int *)(a.sync + 6) = 1;
                    break;
                }
            }
        }
        const double *sg = stages + st * NP * W + tid;
        TRow t;
        t.act = tact;
        t.ht = false;
        if (colT) {
            t.uw = sg[0 * W];
            t.u = sg[0 * W + 1];
            t.vw = sg[1 * W];
            t.v = sg[1 * W + 1];
        } else {
            t.u = t.v = t.uw = t.vw = 0.0;
        }
        const idx_t idx = j * a.pitch + i;
        double str[8];
        if (t.act) {
#pragma unroll
            for (int k = 0; k < EVP_NSTRESS; ++k) t.s[k] = sg[(2 + k) * W + 1];
            t.strength = sg[14 * W + 1];
            t.dxt = sg[15 * W + 1];
            t.dyt = sg[16 * W + 1];
            t.dxhy = sg[17 * W + 1];
            t.dyhx = sg[18 * W + 1];
            t.cxp = sg[19 * W + 1];
            t.cyp = sg[20 * W + 1];
            t.cxm = sg[21 * W + 1];
            t.cym = sg[22 * W + 1];
            t.tiny = sg[23 * W + 1];
            if (LAST) t.tarear = sg[(NP - 1) * W + 1];
            const bool store = ownT && (j < j0 + nrows || j == a.nyl + 1);
            stress_cell<LAST>(a, PlaneStoreT{a, sn, idx, store}, t, us, vs, usw, vsw, str);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) str[k] = 0.0; // str(:,:,:) = c0, :1051
        }
        double *const xl = evp_xch + par * 4 * NT + tid;
        xl[0 * NT] = str[1];
        xl[1 * NT] = str[3];
        xl[2 * NT] = str[6];
        xl[3 * NT] = str[7];
        __syncthreads();
        // every thread has read stage `st` (its values went into str before the barrier): refill it
        if (tid < 32 && r + S < nT) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_issue_row<LAST, W>(a, tid, so, (idx_t)(j + S) * a.pitch + c0, bytes, stages + st * NP * W,
                                   smem_u32(bars + st));
        }
        double s2r = 0.0, s4r = 0.0, s7r = 0.0, s8r = 0.0;
        if (tid < NT - 1) {
            s2r = xl[0 * NT + 1];
            s4r = xl[1 * NT + 1];
            s7r = xl[2 * NT + 1];
            s8r = xl[3 * NT + 1];
        }
        par ^= 1;

        if (uc.act) {
            const double sx = px + str[2] + s4r;          // ((s1 + s2) + s3) + s4
            const double sy = s5c + str[5] + s7c + s8r;    // ((s5 + s6) + s7) + s8
            stepu_cell<LAST>(a, PlaneStoreU{a, sn, idx - a.pitch, i, j - 1}, uc, us, vs, sx, sy);
        }
        px = str[0] + s2r;
        s5c = str[4];
        s7c = s7r;
        us = t.u;
        vs = t.v;
        usw = t.uw;
        vsw = t.vw;
        tm_raw = tm_raw1;
        um_raw = um_raw1;
        if (++st == S) {
            st = 0;
            phase ^= 1u;
        }
    }
}

template <int NT, bool LAST, int S, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_subcycle_tma(const __grid_constant__ SubArgs a) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    const int i = 1 + blockIdx.x * a.strip_w + tid;
    const int j0 = __ldg(a.chunks + 2 * blockIdx.y);
    const int nrows = __ldg(a.chunks + 2 * blockIdx.y + 1);
    const bool top = (j0 + nrows - 1 == a.nyl), bot = (j0 == 1);
    int epoch = 0;
    if (a.p2p && ((top && a.peer_n_flag) || (bot && a.peer_s_flag))) {
        epoch = *(volatile int *)(a.sync + 1);
        p2p_wait(a, tid, top, bot, epoch);
    }
    const idx_t so = a.flip ? (idx_t)a.copy_stride : 0, sn = a.flip ? 0 : (idx_t)a.copy_stride;
    if (nrows > 0) march_tma<NT, LAST, S>(a, so, sn, tid, i, j0, nrows);
    subcycle_epilogue<NT, false>(a, sn, tid, top, bot, epoch, LAST);
}

template <int NT, int S>
static constexpr size_t tma_smem_bytes() {
    return (size_t)(2 * 4 * NT + S * 25 * (NT + 2)) * sizeof(double) + S * sizeof(unsigned long long);
}

// ------------------------------------------------------------------------------------------------
// Strip-tiled, warp-autonomous, TMA-fed subcycle kernel (layout: evp_tiled.cuh).
//
// One WARP owns a strip of 31 U columns (32 T columns) and marches north over its row chunk on its own:
//   * every tile row (the old state copy + all loop-invariant fields of T row j and U row j-1, 8800
//     contiguous bytes) arrives in shared memory by ONE bulk copy (cp.async.bulk, completing on an
//     mbarrier), S rows deep per warp -- no per-thread global address arithmetic, no register prefetch,
//     no register rotation;
//   * str(2,4,7,8) of the east neighbour cell come by warp shuffle, the west velocities too; there is
//     no shared-memory exchange line and no CTA barrier in the loop, so the warps of an SM drift apart
//     and overlap each other's copy and arithmetic phases;
//   * the new stresses and velocities go straight from registers to the new state copy of the tile rows
//     (one base pointer, immediate offsets; 256-byte coalesced segments).  The writer of a U column that a
//     neighbouring strip also holds (its east slot 31 / its west halo word, the east-west wrap columns and,
//     on several GPUs, the neighbour slab's ghost row) stores it there as well -- this is the per-subcycle
//     ice_HaloUpdate(uvel, vvel) (source/ice_dyn_evp.F90:397-402).
// The arithmetic is stress_cell / stepu_cell, shared with the plane kernels: results are bit-identical.
// ------------------------------------------------------------------------------------------------
#define EVT_STAGE_D 1104 // doubles per pipeline stage (EVT_WIN_D rounded up to a multiple of 128 bytes)

struct TileStoreT {
    const SubArgs &a;
    double *row_new; // new state copy of tile row (strip, j)
    idx_t pidx;      // plane index of the T cell (last subcycle's diagnostics stay in planes)
    int lane;
    bool on;
    __device__ __forceinline__ void diag(double divu, double rdg_conv, double rdg_shear, double shear) const {
        a.divu[pidx] = divu; a.rdg_conv[pidx] = rdg_conv; a.rdg_shear[pidx] = rdg_shear; a.shear[pidx] = shear;
    }
    __device__ __forceinline__ void prs(double v) const { a.prs_sig[pidx] = v; }
    __device__ __forceinline__ void stress(int k, double v) const { row_new[EVT_S(k) + lane] = v; }
};

// which other places hold a copy of this lane's U column (constant over the rows of a strip)
struct TileDup {
    int flags;      // bit 0: an east-type slot (u, v 32 apart), bit 1: a west halo word (u, v adjacent)
    int wdE, slotE; // strip and slot of the east-type duplicate
    int wdW;        // strip whose west halo duplicates this column
};

__device__ __forceinline__ TileDup tile_dups(const SubArgs &a, int w, int lane, int i, bool colU) {
    TileDup d = {0, 0, 0, 0};
    if (!colU) return d;
    const int ns = a.t_ns;
    if (lane == 0) {
        if (w > 0) { d.flags |= 1; d.wdE = w - 1; d.slotE = EVT_UW; }                                  // east slot of strip w-1
        else if (a.ew_cyclic) { d.flags |= 1; d.wdE = ns - 1; d.slotE = a.nx - EVT_UW * (ns - 1); } // ghost column nx+1 <- U(1)
    }
    if (lane == EVT_UW - 1 && w < ns - 1) { d.flags |= 2; d.wdW = w + 1; }                             // west halo of strip w+1
    if (i == a.nx && a.ew_cyclic) { d.flags |= 2; d.wdW = 0; }                                         // ghost column 0 <- U(nx)
    return d;
}

// u, v of U column (strip w, lane) of tile row `row` of a tile pool, with its duplicates
__device__ __forceinline__ void tile_store_uv(double *pool, long long sw, long long sj, int row, int copy, int w, int lane,
                                              const TileDup &d, double u, double v) {
    const int so = evt_state_off(copy);
    double *p = pool + (w * sw + row * sj) + so;
    p[EVT_U + lane] = u;
    p[EVT_V + lane] = v;
    if (d.flags & 1) {
        double *q = pool + (d.wdE * sw + row * sj) + so + EVT_U + d.slotE;
        q[0] = u; q[32] = v;
    }
    if (d.flags & 2) {
        double *q = pool + (d.wdW * sw + row * sj) + so + EVT_HALO;
        q[0] = u; q[1] = v;
    }
}

struct TileStoreU {
    const SubArgs &a;
    double *u_new;        // this lane's u slot in the new copy of tile row (strip, j); v is 32 doubles behind
    long long dupE, dupW; // offsets (doubles) from u_new to the duplicates
    const TileDup &d;
    idx_t pidx;
    int w, lane, j, newc;
    __device__ __forceinline__ void uv(double unew, double vnew) const {
        u_new[0] = unew;
        u_new[32] = vnew;
        if (d.flags & 1) { u_new[dupE] = unew; u_new[dupE + 32] = vnew; }
        if (d.flags & 2) { u_new[dupW] = unew; u_new[dupW + 1] = vnew; }
        if (a.p2p) { // slab-to-slab part of the halo update: straight into the neighbour GPU's ghost tile rows
            if (j == a.nyl && a.peer_n_flag)
                tile_store_uv(a.peer_n_tiles, a.peer_n_sw, a.peer_n_sj, 0, newc, w, lane, d, unew, vnew);
            if (j == 1 && a.peer_s_flag)
                tile_store_uv(a.peer_s_tiles, a.peer_s_sw, a.peer_s_sj, a.peer_s_nr - 1, newc, w, lane, d, unew, vnew);
        }
    }
    __device__ __forceinline__ void last(double strintx, double strinty, double taux, double tauy) const {
        a.strintx[pidx] = strintx; a.strinty[pidx] = strinty; a.strocnx[pidx] = taux; a.strocny[pidx] = tauy;
    }
    __device__ __forceinline__ bool finish_here() const { return a.fuse_finish && !(a.fold && j == a.nyl); }
    __device__ __forceinline__ void finish(double xT, double yT) const { a.finx[pidx] = xT; a.finy[pidx] = yT; }
};

// lane 0: arm the stage's mbarrier with the byte count and issue the bulk copy of one read window
__device__ __forceinline__ void tile_issue(const double *src, double *stage, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(EVT_WIN_D * 8))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(stage)),
                 "l"(src), "r"((unsigned)(EVT_WIN_D * 8)), "r"(bar)
                 : "memory");
}

// MEMONLY (measurement only, kernel_variant bit 14, results invalid): the same loads, copies and stores without
// the stress / stepu arithmetic -- the streaming ceiling of this access pattern.
template <bool LAST, int S, bool MEMONLY = false>
__device__ __forceinline__ void march_tiled(const SubArgs &a, double *stages, unsigned long long *bars, int w, int lane,
                                            int j0, int nrows, bool peer_top) {
    const int nx = a.nx;
    const long long sw = a.t_sw, sj = a.t_sj;
    const int i = EVT_UW * w + 1 + lane;
    const bool colT = i <= nx + 1;
    const bool colU = lane < EVT_UW && i <= nx;
    const int jlast = min(j0 + nrows, a.nyl + 1); // last T row of this warp
    const int nT = jlast - j0 + 1;
    const int oldc = a.flip ? 1 : 0, newc = oldc ^ 1;
    // where the old state copy and the loop-invariant block sit inside a stage (evp_tiled.cuh)
    const int st_s = oldc ? EVT_INV_D : 0, inv_s = oldc ? 0 : EVT_STATE_D;
    double *const strip = a.tiles + w * sw;
    const double *const win = strip + j0 * sj + evt_window_off(oldc);

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < S; ++q)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + q)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int q = 0; q < S; ++q)
            if (q < nT) tile_issue(win + q * sj, stages + q * EVT_STAGE_D, smem_u32(bars + q));
    }
    __syncwarp();

    // velocities of the row south of the chunk
    double us = 0.0, vs = 0.0, usw, vsw;
    {
        const double *sr = strip + (j0 - 1) * sj + evt_state_off(oldc);
        if (colT) {
            us = __ldcg(sr + EVT_U + lane);
            vs = __ldcg(sr + EVT_V + lane);
        }
        usw = __shfl_up_sync(0xffffffffu, us, 1);
        vsw = __shfl_up_sync(0xffffffffu, vs, 1);
        if (lane == 0) {
            usw = __ldcg(sr + EVT_HALO);
            vsw = __ldcg(sr + EVT_HALO + 1);
        }
    }
    const TileDup dup = tile_dups(a, w, lane, i, colU);
    const long long dupE = (dup.wdE - w) * sw + (dup.slotE - lane);
    const long long dupW = (dup.wdW - w) * sw + (EVT_HALO - (EVT_U + lane));
    double *rown = strip + j0 * sj + evt_state_off(newc); // new state copy of tile row j
    double px = 0.0, s5c = 0.0, s7c = 0.0;
    int st = 0;
    unsigned phase = 0;

    for (int r = 0; r < nT; ++r) {
        const int j = j0 + r;
        { // wait for the row's bulk copy (bounded: a protocol bug must not hang the device)
            const unsigned bar = smem_u32(bars + st);
            unsigned done = 0, spins = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done)
                             : "r"(bar), "r"(phase)
                             : "memory");
                if (!done && ++spins > (1u << 22)) {
                    *(volatile int *)(a.sync + 6) = 1;
                    break;
                }
            }
        }
        double *const stage = stages + st * EVT_STAGE_D;
        const double *sg = stage + st_s, *si = stage + inv_s;
        const unsigned char *mk = (const unsigned char *)(si + EVT_MASK);
        TRow t;
        URow uc;
        t.ht = false;
        t.act = colT && mk[lane] != 0;
        uc.act = colU && r > 0 && mk[32 + lane] != 0;
        double hu, hv;
        if (peer_top && j == a.nyl + 1) {
            // ghost row written by the north neighbour GPU while this kernel runs: read it with coherent
            // generic loads (after the flag wait) instead of trusting the bulk copy's view of it
            const double *gr = strip + j * sj + evt_state_off(oldc);
            t.u = colT ? __ldcg(gr + EVT_U + lane) : 0.0;
            t.v = colT ? __ldcg(gr + EVT_V + lane) : 0.0;
            hu = __ldcg(gr + EVT_HALO);
            hv = __ldcg(gr + EVT_HALO + 1);
        } else {
            t.u = colT ? sg[EVT_U + lane] : 0.0;
            t.v = colT ? sg[EVT_V + lane] : 0.0;
            hu = sg[EVT_HALO];
            hv = sg[EVT_HALO + 1];
        }
        t.uw = __shfl_up_sync(0xffffffffu, t.u, 1);
        t.vw = __shfl_up_sync(0xffffffffu, t.v, 1);
        if (lane == 0) {
            t.uw = hu;
            t.vw = hv;
        }
        const idx_t pidx = j * a.pitch + i;
        if (t.act) {
#pragma unroll
            for (int k = 0; k < EVP_NSTRESS; ++k) t.s[k] = sg[EVT_S(k) + lane];
            t.strength = si[EVT_T(0) + lane];
            t.dxt = si[EVT_T(1) + lane];
            t.dyt = si[EVT_T(2) + lane];
            t.dxhy = si[EVT_T(3) + lane];
            t.dyhx = si[EVT_T(4) + lane];
            t.cxp = si[EVT_T(5) + lane];
            t.cyp = si[EVT_T(6) + lane];
            t.cxm = si[EVT_T(7) + lane];
            t.cym = si[EVT_T(8) + lane];
            t.tiny = si[EVT_T(9) + lane];
            if (LAST) t.tarear = __ldg(a.tarear + pidx);
        }
        if (uc.act) {
            uc.aiu = si[EVT_UF(0) + lane];
            uc.uocn = si[EVT_UF(1) + lane];
            uc.vocn = si[EVT_UF(2) + lane];
            uc.waterx = si[EVT_UF(3) + lane];
            uc.watery = si[EVT_UF(4) + lane];
            uc.forcex = si[EVT_UF(5) + lane];
            uc.forcey = si[EVT_UF(6) + lane];
            uc.umassdtei = si[EVT_UF(7) + lane];
            uc.fm = si[EVT_UF(8) + lane];
            uc.uarear = si[EVT_UF(9) + lane];
        }
        // Every lane holds its values in registers: the stage can be refilled with row j + S right away.  The
        // warp barrier orders the lanes' shared-memory reads before the issue of the copy (write-after-read
        // across the proxies needs no proxy fence: the copy is issued after the barrier, like a TMA producer
        // that acquired an "empty" mbarrier).
        __syncwarp();
        if (lane == 0 && r + S < nT) tile_issue(win + (r + S) * sj, stage, smem_u32(bars + st));

        double str[8];
        if (MEMONLY) {
#pragma unroll
            for (int k = 0; k < 8; ++k) str[k] = 0.0;
            if (t.act) {
                const double extra = t.strength + t.dxt + t.dyt + t.dxhy + t.dyhx + t.cxp + t.cyp + t.cxm + t.cym + t.tiny;
                if (j < j0 + nrows || j == a.nyl + 1) {
#pragma unroll
                    for (int k = 0; k < EVP_NSTRESS; ++k) rown[EVT_S(k) + lane] = t.s[k] + (k == 0 ? extra * 1e-300 : 0.0);
                }
                str[0] = t.uw + t.vw;
            }
            if (uc.act) {
                double *un = rown - sj + EVT_U + lane;
                const double e2 = uc.aiu + uc.uocn + uc.vocn + uc.waterx + uc.watery + uc.forcex + uc.forcey + uc.umassdtei + uc.fm + uc.uarear;
                un[0] = us + (e2 + usw) * 1e-300;
                un[32] = vs + vsw * 1e-300;
            }
        } else if (t.act) {
            const bool store = j < j0 + nrows || j == a.nyl + 1;
            stress_cell<LAST>(a, TileStoreT{a, rown, pidx, lane, store}, t, us, vs, usw, vsw, str);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) str[k] = 0.0; // str(:,:,:) = c0, :1051
        }
        // str(2,4,7,8) of the T cell east of this lane's U point
        const double s2r = __shfl_down_sync(0xffffffffu, str[1], 1);
        const double s4r = __shfl_down_sync(0xffffffffu, str[3], 1);
        const double s7r = __shfl_down_sync(0xffffffffu, str[6], 1);
        const double s8r = __shfl_down_sync(0xffffffffu, str[7], 1);
        if (!MEMONLY && uc.act) {
            const double sx = px + str[2] + s4r;       // ((s1 + s2) + s3) + s4
            const double sy = s5c + str[5] + s7c + s8r; // ((s5 + s6) + s7) + s8
            stepu_cell<LAST>(a, TileStoreU{a, rown - sj + EVT_U + lane, dupE, dupW, dup, pidx - a.pitch, w, lane, j - 1, newc},
                             uc, us, vs, sx, sy);
        }
        px = str[0] + s2r;
        s5c = str[4];
        s7c = s7r;
        us = t.u;
        vs = t.v;
        usw = t.uw;
        vsw = t.vw;
        rown += sj;
        if (++st == S) {
            st = 0;
            phase ^= 1u;
        }
    }
}

// Tripole u-fold on the tiled layout (north-south part of the halo update on the top slab): the last CTA of
// the northernmost chunk to finish symmetrises the top physical row and fills the ghost row, through a scratch
// copy of the raw top row (same arithmetic as k_halo_tripole / evp_fold_necorner).
__device__ __forceinline__ void tile_write_col(const SubArgs &a, int copy, int cc, int row, double u, double v) {
    int dv;
    double *pool = a.tiles;
    double *p = pool + evt_u_primary(cc, row, a.nx, a.t_ns, a.t_sw, a.t_sj, copy, dv);
    p[0] = u;
    p[dv] = v;
    if (cc >= 1 && cc <= a.nx) {
        const int w = (cc - 1) / EVT_UW, l = (cc - 1) - w * EVT_UW;
        const int so = evt_state_off(copy);
        if (l == 0 && w > 0) {
            double *q = pool + ((w - 1) * a.t_sw + row * a.t_sj) + so + EVT_U + EVT_UW;
            q[0] = u; q[32] = v;
        }
        if (l == EVT_UW - 1 && w < a.t_ns - 1) {
            double *q = pool + ((w + 1) * a.t_sw + row * a.t_sj) + so + EVT_HALO;
            q[0] = u; q[1] = v;
        }
    }
}

// Barrier among the CTAs of the northernmost chunk (they are all resident: the grid is one wave).  The counter
// only grows; the target is the next multiple of the number of participants, so no reset is needed between
// kernels.  Bounded like every other wait.
__device__ __forceinline__ void top_chunk_barrier(int *counter, int tid, int *sync) {
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned n = gridDim.x;
        const unsigned prev = atomicAdd((unsigned *)counter, 1u);
        const unsigned target = (prev / n + 1u) * n;
        unsigned spins = 0;
        while ((int)((unsigned)ld_acquire(counter) - target) < 0) {
            if (++spins > (1u << 24)) {
                *(volatile int *)(sync + 6) = 1;
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// End of a subcycle on the tiled layout.  Tripole u-fold (north-south part of the halo update on the top slab,
// serial/ice_boundary.F90:777-866) by ALL CTAs of the northernmost chunk together: once every one of them has
// finished its rows, each thread takes one column of its own strip, loads the raw top-row values of that
// column's fold partner (and the row below it), and -- after a second barrier, because the update is in place --
// writes the symmetrised top row and the ghost row.  Same arithmetic as evp_fold_necorner / k_halo_tripole.
template <int NT>
__device__ __forceinline__ void tiled_epilogue(const SubArgs &a, int newc, int tid, bool top, bool last) {
    if (a.fold && top) {
        top_chunk_barrier(a.sync + 4, tid, a.sync);
        const int nx = a.nx, nyl = a.nyl, warp = tid >> 5, lane = tid & 31;
        const int w = blockIdx.x * 4 + warp;
        // lanes 0..30: the U columns of strip w; lane 31 of the first / last strip: the ghost columns 0 / nx+1
        int cc = EVT_UW * w + 1 + lane;
        bool mine = w < a.t_ns && lane < EVT_UW && cc <= nx;
        if (lane == EVT_UW && w == 0) { cc = 0; mine = true; }
        if (lane == EVT_UW && w == a.t_ns - 1 && a.t_ns > 1) { cc = nx + 1; mine = true; }
        double ut = 0.0, vt = 0.0, ug = 0.0, vg = 0.0, ut2 = 0.0, vt2 = 0.0, ug2 = 0.0, vg2 = 0.0;
        auto fold_col = [&](int c, double &o_ut, double &o_vt, double &o_ug, double &o_vg) {
            int ig = c; // i_glob of the column (source/ice_blocks.F90:291-330)
            if (c == 0) ig = a.ew_cyclic ? nx : 1;
            if (c == nx + 1) ig = a.ew_cyclic ? 1 : nx;
            int k = nx - ig; // iSrc = nxGlobal - i_glob + 1 - ioffset, ioffset = 1
            if (k == 0) k = nx;
            int dv, dv2 = 32, dvb;
            const double *pk = a.tiles + evt_u_primary(k, nyl, nx, a.t_ns, a.t_sw, a.t_sj, newc, dv);
            const bool pair = k >= 1 && k <= nx - 1;
            const double *pn = pair ? a.tiles + evt_u_primary(nx - k, nyl, nx, a.t_ns, a.t_sw, a.t_sj, newc, dv2) : pk;
            const double *pb = a.tiles + evt_u_primary(k, nyl - 1, nx, a.t_ns, a.t_sw, a.t_sj, newc, dvb);
            const double uk = __ldcg(pk), vk = __ldcg(pk + dv), un = __ldcg(pn), vn = __ldcg(pn + (pair ? dv2 : dv));
            const double ub = __ldcg(pb), vb = __ldcg(pb + dvb);
            const double isign = -1.0;
            double xu, xv;
            if (k >= 1 && k <= nx / 2 - 1) {
                xu = 0.5 * (uk + isign * un);
                xv = 0.5 * (vk + isign * vn);
            } else if (k >= nx - (nx / 2 - 1) && k <= nx - 1) {
                xu = isign * (0.5 * (un + isign * uk)); // partner of the loop index i = nx - k
                xv = isign * (0.5 * (vn + isign * vk));
            } else {
                xu = uk;
                xv = vk;
            }
            o_ut = isign * xu; o_vt = isign * xv; // row nyl   <- isign * buf(iSrc, 2)
            o_ug = isign * ub; o_vg = isign * vb; // row nyl+1 <- isign * buf(iSrc, 1)
        };
        // a single strip holds both ghost columns on lane 31: the second one goes through the *2 set
        const bool both = a.t_ns == 1 && lane == EVT_UW && w == 0;
        if (mine) fold_col(cc, ut, vt, ug, vg);
        if (both) fold_col(nx + 1, ut2, vt2, ug2, vg2);
        top_chunk_barrier(a.sync + 5, tid, a.sync);
        if (mine) {
            tile_write_col(a, newc, cc, nyl, ut, vt);
            tile_write_col(a, newc, cc, nyl + 1, ug, vg);
            // last subcycle: evp_finish of the row whose velocities have just been folded
            if (last && a.fuse_finish && cc >= 1 && cc <= nx) finish_fold_column(a, cc, ut, vt);
        }
        if (both) {
            tile_write_col(a, newc, nx + 1, nyl, ut2, vt2);
            tile_write_col(a, newc, nx + 1, nyl + 1, ug2, vg2);
        }
    }
    if (a.p2p) { // the last CTA of the grid to finish advances this rank's count of completed kernels
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned total = gridDim.x * gridDim.y;
            const unsigned prev = atomicAdd((unsigned *)a.sync, 1u);
            if (prev == total - 1) {
                a.sync[0] = 0;
                a.sync[1] = a.sync[1] + 1;
            }
        }
    }
}

template <bool LAST, int S, int MINB, bool MEMONLY = false>
__global__ void __launch_bounds__(128, MINB) k_subcycle_tiled(const __grid_constant__ SubArgs a) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    extern __shared__ __align__(128) double evt_smem[]; // [4 warps][S stages][EVT_STAGE_D], then 4 * S mbarriers
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *const stages = evt_smem + (size_t)warp * S * EVT_STAGE_D;
    unsigned long long *const bars = (unsigned long long *)(evt_smem + 4 * S * EVT_STAGE_D) + warp * S;
    const int w = blockIdx.x * 4 + warp; // strip of this warp
    const int j0 = __ldg(a.chunks + 2 * blockIdx.y);
    const int nrows = __ldg(a.chunks + 2 * blockIdx.y + 1);
    const bool top = (j0 + nrows - 1 == a.nyl), bot = (j0 == 1);
    const bool live = w < a.t_ns && nrows > 0;
    const bool peer_cta = a.p2p && ((top && a.peer_n_flag) || (bot && a.peer_s_flag));
    int epoch = 0;
    if (peer_cta && live) {
        // wait until strips w-1, w, w+1 of the neighbour's adjacent chunk have published as many finished
        // subcycles as this rank has completed (see p2p_wait)
        epoch = *(volatile int *)(a.sync + 1);
        if (lane < 3) {
            const int x = (w + lane - 1 + a.t_ns) % a.t_ns;
            if (top && a.peer_n_flag) wait_flag_ge(a.sync + EVP_SYNC_FN + x, epoch, a.sync);
            if (bot && a.peer_s_flag) wait_flag_ge(a.sync + EVP_SYNC_FS + x, epoch, a.sync);
            __threadfence_system();
        }
        __syncwarp();
    }
    if (live) march_tiled<LAST, S, MEMONLY>(a, stages, bars, w, lane, j0, nrows, top && a.p2p && a.peer_n_flag != nullptr);
    if (peer_cta && live) {
        // this strip's boundary rows are in the neighbour's ghost tile rows: make them visible system-wide,
        // then publish the strip's epoch in the neighbour's sync block
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            if (top && a.peer_n_flag) *(volatile int *)(a.peer_n_flag + w) = epoch + 1;
            if (bot && a.peer_s_flag) *(volatile int *)(a.peer_s_flag + w) = epoch + 1;
        }
    }
    tiled_epilogue<128>(a, a.flip ? 0 : 1, tid, top, LAST);
}

template <int S>
static constexpr size_t tiled_smem_bytes() {
    return (size_t)4 * S * EVT_STAGE_D * sizeof(double) + 4 * S * sizeof(unsigned long long);
}

template <int S, int MINB>
static int tiled_launch(const SubArgs &a, bool last, bool pdl, bool memonly, unsigned gx, unsigned gy, cudaStream_t s,
                        int *ctas_per_sm) {
    auto k0 = memonly ? k_subcycle_tiled<false, S, MINB, true> : k_subcycle_tiled<false, S, MINB>;
    auto k1 = k_subcycle_tiled<true, S, MINB>;
    const size_t smem = tiled_smem_bytes<S>();
    static bool configured = false; // once per process and instantiation (outside any stream capture)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_subcycle_tiled<false, S, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_subcycle_tiled<false, S, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (ctas_per_sm) {
        int n0 = 0, n1 = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, k0, 128, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k1, 128, smem);
        *ctas_per_sm = n0 < n1 ? n0 : n1;
        return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)(last ? cudaLaunchKernelEx(&cfg, k1, a) : cudaLaunchKernelEx(&cfg, k0, a));
}

// ------------------------------------------------------------------------------------------------
// Persistent kernel: all a.nsub subcycles of one evp call in ONE cooperative launch (every CTA is
// resident; the grid is the same one-wave grid as k_subcycle's and each CTA keeps its strip and row
// chunk for the whole loop).  There is no grid-wide barrier: because u, v and the stresses are
// ping-ponged, a CTA may start subcycle k as soon as its (up to eight) neighbouring CTAs have
// finished subcycle k-1 -- that single condition covers the read-after-write of the neighbours' new
// values and the write-after-read of the copy they were still reading.  Each CTA publishes the number
// of subcycles it has finished in cta_epoch[] (zeroed by the host before the launch) after a
// __threadfence; state loads use ld.global.cg, because the L1 is not coherent with the other SMs'
// stores.  What this removes per subcycle: the kernel launch gap, the ramp-up and tail of a grid, and
// the wait of every CTA for the slowest one.
// ------------------------------------------------------------------------------------------------
// one subcycle of the persistent loop on one CTA
template <int NT, bool LAST, bool HT>
__device__ __forceinline__ void persist_step(const SubArgs &a, int k, const int *s_nb, int *my_epoch, int tid, int i,
                                             int j0, int nrows, bool top, bool bot, bool p2p_cta) {
    const bool odd = ((a.flip + k) & 1) != 0;
    const idx_t so = odd ? (idx_t)a.copy_stride : 0, sn = odd ? 0 : (idx_t)a.copy_stride;
    if (k > 0) {
        if (tid < 8 && s_nb[tid] >= 0) wait_flag_ge(a.cta_epoch + s_nb[tid], k, a.sync);
        if (tid == 8 && a.fold && top) wait_flag_ge(a.sync + 5, a.epoch0 + k, a.sync);
        __syncthreads();
    }
    if (p2p_cta) p2p_wait(a, tid, top, bot, a.epoch0 + k);
    march<NT, LAST, HT, true>(a, so, sn, tid, i, j0, nrows);
    subcycle_epilogue<NT, true>(a, sn, tid, top, bot, a.epoch0 + k, LAST);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        *(volatile int *)my_epoch = k + 1;
    }
}

template <int NT, bool HT>
__global__ void __launch_bounds__(NT) k_persist(const __grid_constant__ SubArgs a) {
    __shared__ int s_nb[8];
    const int tid = threadIdx.x;
    const int i = 1 + blockIdx.x * a.strip_w + tid;
    const int ncx = (int)gridDim.x, ncy = (int)gridDim.y;
    const int j0 = __ldg(a.chunks + 2 * blockIdx.y);
    const int nrows = __ldg(a.chunks + 2 * blockIdx.y + 1);
    const bool top = (j0 + nrows - 1 == a.nyl), bot = (j0 == 1);
    int *const my_epoch = a.cta_epoch + blockIdx.y * ncx + blockIdx.x;
    if (nrows <= 0) { // empty chunk: nobody depends on it
        if (tid == 0) *(volatile int *)my_epoch = 0x7fffffff;
        return;
    }
    // the CTAs whose cells this one reads / whose reads it overwrites: strips x-1, x, x+1 of this chunk
    // and of the chunks that own the rows just south and just north of it (rows without active cells
    // belong to no chunk: nothing is read or written there)
    if (tid < 8) {
        int ys = -1, yn = -1;
        for (int c = 0; c < ncy; ++c) {
            const int cj = __ldg(a.chunks + 2 * c), cn = __ldg(a.chunks + 2 * c + 1);
            if (cn <= 0) continue;
            if (cj + cn == j0) ys = c;
            if (cj == j0 + nrows) yn = c;
        }
        const int bx = (int)blockIdx.x;
        int xw = bx - 1, xe = bx + 1;
        if (xw < 0) xw = a.ew_cyclic ? ncx - 1 : -1;
        if (xe >= ncx) xe = a.ew_cyclic ? 0 : -1;
        // tid: 0 W, 1 E, 2 S, 3 SW, 4 SE, 5 N, 6 NW, 7 NE
        const int y = tid < 2 ? (int)blockIdx.y : (tid < 5 ? ys : yn);
        const int x = (tid == 2 || tid == 5) ? bx : ((tid == 0 || tid == 3 || tid == 6) ? xw : xe);
        int nb = (x >= 0 && y >= 0) ? y * ncx + x : -1;
        if (nb == (int)(blockIdx.y * ncx + blockIdx.x)) nb = -1; // a single strip wraps onto itself
        s_nb[tid] = nb;
    }
    __syncthreads();
    const bool p2p_cta = a.p2p && ((top && a.peer_n_flag) || (bot && a.peer_s_flag));
    // only the last subcycle's diagnostics are observable: it runs the LAST instance of the march
    for (int k = 0; k < a.nsub - 1; ++k)
        persist_step<NT, false, HT>(a, k, s_nb, my_epoch, tid, i, j0, nrows, top, bot, p2p_cta);
    persist_step<NT, true, HT>(a, a.nsub - 1, s_nb, my_epoch, tid, i, j0, nrows, top, bot, p2p_cta);
    if (a.p2p && tid == 0) { // rank-level count of completed subcycles (read by k_wait_peers and k_subcycle)
        int live = 0; // empty chunks returned at once and are not counted
        for (int c = 0; c < ncy; ++c) live += __ldg(a.chunks + 2 * c + 1) > 0 ? 1 : 0;
        const unsigned prev = atomicAdd((unsigned *)a.sync, 1u);
        if (prev == (unsigned)(live * ncx) - 1u) {
            a.sync[0] = 0;
            a.sync[1] = a.epoch0 + a.nsub;
        }
    }
}

// Launch with programmatic dependent launch (PDL): the next subcycle kernel may be scheduled while
// this one drains; its CTAs block in griddepcontrol.wait at their first instruction until this grid
// has completed and its stores are visible, so only launch latency and ramp-up overlap.
template <typename K>
static int launch_k(K kernel, const SubArgs &a, dim3 grid, dim3 block, bool pdl, cudaStream_t s, int xch = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = xch ? 2 * 4 * block.x * sizeof(double) : 0; // evp_xch
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kernel, a);
}

template <typename K>
static int launch_tma(K kernel, size_t smem, const SubArgs &a, dim3 grid, dim3 block, cudaStream_t s) {
    kernel<<<grid, block, smem, s>>>(a);
    return (int)cudaPeekAtLastError();
}

// run-time selection of the LAST / HT (2-plane metric path, a.row_ht) / COH (peer-to-peer halo) instances
template <int NT, bool LATE, int MINB, bool WARPX>
static int launch_sel(const SubArgs &a, bool last, dim3 grid, dim3 block, bool pdl, cudaStream_t s) {
    const int xch = WARPX ? 0 : 1;
    if (a.p2p) {
        if (a.row_ht) {
            if (last) return launch_k(k_subcycle<NT, true, true, LATE, MINB, true, WARPX>, a, grid, block, pdl, s, xch);
            return launch_k(k_subcycle<NT, false, true, LATE, MINB, true, WARPX>, a, grid, block, pdl, s, xch);
        }
        if (last) return launch_k(k_subcycle<NT, true, false, LATE, MINB, true, WARPX>, a, grid, block, pdl, s, xch);
        return launch_k(k_subcycle<NT, false, false, LATE, MINB, true, WARPX>, a, grid, block, pdl, s, xch);
    }
    if (a.row_ht) {
        if (last) return launch_k(k_subcycle<NT, true, true, LATE, MINB, false, WARPX>, a, grid, block, pdl, s, xch);
        return launch_k(k_subcycle<NT, false, true, LATE, MINB, false, WARPX>, a, grid, block, pdl, s, xch);
    }
    if (last) return launch_k(k_subcycle<NT, true, false, LATE, MINB, false, WARPX>, a, grid, block, pdl, s, xch);
    return launch_k(k_subcycle<NT, false, false, LATE, MINB, false, WARPX>, a, grid, block, pdl, s, xch);
}

// HT (2-plane metric path) is chosen by a.row_ht; variant bit 8 (256): TMA staging, 3 rows deep, 2 CTAs
// per SM; bit 9 (512): TMA staging, 2 rows deep, 3 CTAs per SM (<= 168 registers) -- 128 threads only
template <int NT>
static int launch_nt(const SubArgs &a, bool last, bool pdl, int variant, unsigned gx, unsigned gy, cudaStream_t s) {
    dim3 grid(gx, gy), block(NT);
    if constexpr (NT == 128) {
        if ((variant & 1024) && (variant & 1048576)) // the same with one strip per warp (2-plane path allowed)
            return launch_sel<NT, true, 3, true>(a, last, grid, block, pdl, s);
        if (variant & 1024) { // no prefetch across the arithmetic, 3 CTAs per SM (<= 168 registers)
            if (a.p2p) {
                if (last) return launch_k(k_subcycle<NT, true, false, true, 3, true>, a, grid, block, pdl, s);
                return launch_k(k_subcycle<NT, false, false, true, 3, true>, a, grid, block, pdl, s);
            }
            if (last) return launch_k(k_subcycle<NT, true, false, true, 3>, a, grid, block, pdl, s);
            return launch_k(k_subcycle<NT, false, false, true, 3>, a, grid, block, pdl, s);
        }
        if (variant & 256) {
            if (last) return launch_tma(k_subcycle_tma<NT, true, 3, 2>, tma_smem_bytes<NT, 3>(), a, grid, block, s);
            return launch_tma(k_subcycle_tma<NT, false, 3, 2>, tma_smem_bytes<NT, 3>(), a, grid, block, s);
        }
        if (variant & 512) {
            if (last) return launch_tma(k_subcycle_tma<NT, true, 2, 3>, tma_smem_bytes<NT, 2>(), a, grid, block, s);
            return launch_tma(k_subcycle_tma<NT, false, 2, 3>, tma_smem_bytes<NT, 2>(), a, grid, block, s);
        }
    }
    if constexpr (NT == 128) {
        // warp-autonomous strips: shuffles instead of the exchange line, no row barrier
        if (variant & 1048576) return launch_sel<NT, false, 1, true>(a, last, grid, block, pdl, s);
    }
    if (a.p2p) { // the neighbours write this slab's ghost rows during the kernel: coherent state loads
        if (a.row_ht) {
            if (last) return launch_k(k_subcycle<NT, true, true, false, 1, true>, a, grid, block, pdl, s);
            return launch_k(k_subcycle<NT, false, true, false, 1, true>, a, grid, block, pdl, s);
        }
        if (last) return launch_k(k_subcycle<NT, true, false, false, 1, true>, a, grid, block, pdl, s);
        return launch_k(k_subcycle<NT, false, false, false, 1, true>, a, grid, block, pdl, s);
    }
    if (a.row_ht) {
        if (last) return launch_k(k_subcycle<NT, true, true>, a, grid, block, pdl, s);
        return launch_k(k_subcycle<NT, false, true>, a, grid, block, pdl, s);
    }
    if (last) return launch_k(k_subcycle<NT, true, false>, a, grid, block, pdl, s);
    return launch_k(k_subcycle<NT, false, false>, a, grid, block, pdl, s);
}

template <int NT>
static int persist_nt(const SubArgs &a, unsigned gx, unsigned gy, cudaStream_t s, int *ctas_per_sm) {
    auto kern = a.row_ht ? k_persist<NT, true> : k_persist<NT, false>;
    const size_t smem = 2 * 4 * NT * sizeof(double); // evp_xch
    if (ctas_per_sm) {
        int n = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, NT, smem);
        *ctas_per_sm = n;
        return (int)e;
    }
    SubArgs arg = a;
    void *params[1] = {(void *)&arg};
    return (int)cudaLaunchCooperativeKernel((const void *)kern, dim3(gx, gy), dim3(NT), params, smem, s);
}

} // namespace EVP_SUB_NS

#ifndef EVP_BODY_NO_LAUNCHERS
// returns a cudaError_t value (the launch status)
int EVP_SUB_LAUNCH(const SubArgs &a, bool last, int variant, int threads, unsigned grid_x,
                   unsigned grid_y, void *stream) {
    const bool pdl = (variant & 64) != 0;
    cudaStream_t s = (cudaStream_t)stream;
    switch (threads) {
    case 64: return EVP_SUB_NS::launch_nt<64>(a, last, pdl, variant, grid_x, grid_y, s);
    case 256: return EVP_SUB_NS::launch_nt<256>(a, last, pdl, variant, grid_x, grid_y, s);
    default: return EVP_SUB_NS::launch_nt<128>(a, last, pdl, variant, grid_x, grid_y, s);
    }
}

// opt in to more than 48 KB of dynamic shared memory for the TMA-staged kernels (once per device, outside
// any stream capture); returns a cudaError_t value
int EVP_SUB_CONFIGURE(void) {
    using namespace EVP_SUB_NS;
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kernel, size_t smem) {
        const cudaError_t r = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = r;
    };
    set(k_subcycle_tma<128, false, 3, 2>, tma_smem_bytes<128, 3>());
    set(k_subcycle_tma<128, true, 3, 2>, tma_smem_bytes<128, 3>());
    set(k_subcycle_tma<128, false, 2, 3>, tma_smem_bytes<128, 2>());
    set(k_subcycle_tma<128, true, 2, 3>, tma_smem_bytes<128, 2>());
    return (int)e;
}

// strip-tiled TMA-fed kernel: launch (ctas_per_sm == nullptr) or configure + occupancy query
int EVP_TILED_LAUNCH(const SubArgs &a, bool last, int stages, int flags, unsigned grid_x, unsigned grid_y, void *stream,
                     int *ctas_per_sm) {
    cudaStream_t s = (cudaStream_t)stream;
    const bool pdl = (flags & 1) != 0, memonly = (flags & 2) != 0;
    if (stages == 3) return EVP_SUB_NS::tiled_launch<3, 2>(a, last, pdl, memonly, grid_x, grid_y, s, ctas_per_sm);
    return EVP_SUB_NS::tiled_launch<2, 3>(a, last, pdl, memonly, grid_x, grid_y, s, ctas_per_sm);
}

// persistent kernel: launch (ctas_per_sm == nullptr) or occupancy query; returns a cudaError_t value
int EVP_PERSIST_LAUNCH(const SubArgs &a, int threads, unsigned grid_x, unsigned grid_y, void *stream,
                       int *ctas_per_sm) {
    cudaStream_t s = (cudaStream_t)stream;
    switch (threads) {
    case 64: return EVP_SUB_NS::persist_nt<64>(a, grid_x, grid_y, s, ctas_per_sm);
    case 256: return EVP_SUB_NS::persist_nt<256>(a, grid_x, grid_y, s, ctas_per_sm);
    default: return EVP_SUB_NS::persist_nt<128>(a, grid_x, grid_y, s, ctas_per_sm);
    }
}
#endif // EVP_BODY_NO_LAUNCHERS

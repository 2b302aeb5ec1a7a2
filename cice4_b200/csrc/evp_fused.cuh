// evp_fused.cuh -- TWO EVP subcycles per launch (temporal blocking of the ndte loop) on the plane layout.
//
// Included after evp_subcycle_body.cuh (with EVP_BODY_NO_LAUNCHERS) by evp_fused_strict.cu / evp_fused_fast.cu; the
// arithmetic is stress_cell / stepu_cell of that file (source/ice_dyn_evp.F90:947-1443), so the results are
// bit-identical to two launches of k_subcycle.
//
// Why: one subcycle streams 384 B per active cell (12+12 stresses, 2+2 velocities, 20 loop-invariant fields) and
// the one-subcycle kernels sit on the DRAM ceiling of that access pattern.  The dependency radius of a subcycle
// is one cell (stress of T(i,j) needs u at (i-1..i, j-1..j); stepu of U(i,j) needs str of T(i..i+1, j..j+1)), so
// a second subcycle can run two rows behind the first one on data that never left the SM: the state is read once
// and written once per TWO subcycles and the loop-invariant fields are read once instead of twice (~200 B per
// cell and subcycle instead of 384).
//
// How: one WARP owns a strip of up to 29 U columns and marches north over its row chunk on its own (no CTA
// barrier; cross-lane values by warp shuffle).  In iteration jA it
//   (1) issues the global loads of stage A's T row jA and U row jA-1 (old state copy + loop-invariant fields),
//   (2) runs stage B -- the SECOND subcycle -- for T row jA-2 and U row jA-3 from operands that stage A left in
//       shared memory two iterations earlier (stresses and velocities after the first subcycle, loop-invariant
//       fields), storing the final stresses / velocities into the new state copy; this hides the latency of (1),
//   (3) runs stage A -- the FIRST subcycle -- for T row jA and U row jA-1 and leaves its results in a 3-row ring
//       in shared memory (26 KB per warp).
// Stage B of lane l needs the first-subcycle velocities of lanes l-1 and l, stage A's stepu the stresses of lane
// l+1: of the 32 T columns a warp holds, lanes own_lo..own_hi (up to 29) produce final values, the rest is
// redundant work (like the reference's redundant N/E ghost-cell stresses, :846-859), and a chunk runs stage A
// on three more rows than it owns.  Everything is computed from the old copy only, so the result does not
// depend on the tiling.
//
// Halo updates (ice_HaloUpdate(uvel/vvel) after EVERY subcycle, :397-402):
//   * east-west wrap: the first / last warp strip of a cyclic domain holds the columns around the seam in the
//     order  ..., nx, nx+1 (ghost T column), 1, 2, ...  (`G` = lane of the ghost column): the intermediate
//     velocity of ghost column nx+1 is taken from the lane of column 1 and the west neighbour of column 1 is the
//     lane of column nx; the final velocities are stored with their wrap duplicates as in k_subcycle;
//   * closed / open boundaries: ghost velocities keep their plane value in both subcycles;
//   * tripole u-fold (non-local): the CTAs of the northernmost chunk do not fuse.  They run the one-subcycle
//     march twice on their (short) chunk through a third state copy, with the fold of the intermediate and of the
//     final velocities in between (last CTA to arrive folds, the others wait on a flag) -- the chunk below them
//     needs nothing of that: its stage A recomputes what it needs from the old copy.
// Not handled here (the caller falls back to one subcycle per launch): several ranks (slab-to-slab rows),
// north-south cyclic domains, the T-fold, the 2-plane metric path.
#pragma once

namespace EVP_SUB_NS {

// shared-memory ring of one warp (doubles): 3 rows each of the first-subcycle stresses [12][32], the loop-invariant
// T fields [10][32], the loop-invariant U fields [10][32] and the first-subcycle velocities [2][32]
#define EVF_SIG 0
#define EVF_TIN (EVF_SIG + 3 * 12 * 32)
#define EVF_UIN (EVF_TIN + 3 * 10 * 32)
#define EVF_U1 (EVF_UIN + 3 * 10 * 32)
#define EVF_WARP_D (EVF_U1 + 3 * 2 * 32)

// stage A keeps its results on the SM
struct StashStoreT {
    double *sig; // this lane's column of the ring row: stress k at sig[k * 32]
    bool on;
    __device__ __forceinline__ void diag(double, double, double, double) const {}
    __device__ __forceinline__ void prs(double) const {}
    __device__ __forceinline__ void stress(int k, double v) const { sig[k * 32] = v; }
};
struct RegStoreU {
    double &u1, &v1;
    __device__ __forceinline__ void uv(double unew, double vnew) const { u1 = unew; v1 = vnew; }
    __device__ __forceinline__ void last(double, double, double, double) const {}
};

template <bool LASTB>
__device__ __forceinline__ void wmarch2(const SubArgs &a, idx_t so, idx_t sn, double *wsm, int lane, int j0, int nrows,
                                        int wi) {
    const int4 d = __ldg((const int4 *)a.wstrips + wi);
    const int vcol0 = d.x, own_lo = d.y & 255, own_hi = d.y >> 8, G = d.z, lane_lo = d.w & 255, lane_hi = d.w >> 8;
    if (own_hi < own_lo) return; // a warp without columns (last CTA of the row)
    const int nx = a.nx, nyl = a.nyl, pitch = a.pitch;
    // plane column of this lane (0 / nx+1: ghost columns)
    const int pc = (G < 0 || lane <= G) ? vcol0 + lane : lane - G;
    const bool inl = lane >= lane_lo && lane <= lane_hi;
    const bool colV = inl && pc >= 0 && pc <= nx + 1;                          // holds a velocity column
    const bool colT = inl && pc >= 1 && pc <= nx + 1;                          // holds a T column
    const bool colUA = colT && pc <= nx && lane != G && lane + 1 <= lane_hi;   // stage A can finish its U point
    const bool needTB = colT && lane >= own_lo && lane <= own_hi + 1;          // stage B computes its T cell
    const bool ownU = lane >= own_lo && lane <= own_hi;
    const bool ownT = ownU || (lane == own_hi + 1 && pc == nx + 1);
    // where stage B finds the first-subcycle velocities of its own and of its west column
    const int wsrc = (G >= 0 && lane == G + 1) ? G - 1 : (lane > 0 ? lane - 1 : 0);
    const int osrc = (G >= 0 && lane == G) ? G + 1 : lane;
    const int j1 = j0 + nrows - 1;              // last U / T row this chunk owns
    const int jBlast = min(j1 + 1, nyl + 1);    // last T row of stage B
    const int a0 = max(j0 - 1, 1);              // first / last T row of stage A
    const int a1 = min(jBlast + 1, nyl + 1);
    double *const SIG = wsm + EVF_SIG, *const TIN = wsm + EVF_TIN, *const UIN = wsm + EVF_UIN, *const U1 = wsm + EVF_U1;

    // velocities of the row south of stage A's first row; they are also the first-subcycle values of that row
    // where nothing is computed (ghost row 0)
    double us = 0.0, vs = 0.0, usw = 0.0, vsw = 0.0;
    if (colV) {
        const idx_t idx = (a0 - 1) * pitch + pc + so;
        us = __ldg(a.u + idx);
        vs = __ldg(a.v + idx);
        if (pc >= 1) {
            usw = __ldg(a.u + idx - 1);
            vsw = __ldg(a.v + idx - 1);
        }
    }
    {
        double *r = U1 + ((a0 - 1) % 3) * 64;
        r[lane] = us;
        r[32 + lane] = vs;
    }
    // mask bytes run one row ahead of the data they gate (kept raw: no data load waits on a mask load)
    const uint8_t *tmk = a.icetmask + (size_t)a0 * pitch + max(pc, 0);
    const uint8_t *umk = a.iceumask + (size_t)a0 * pitch + max(pc, 0);
    unsigned tm_raw = colT ? __ldg(tmk) : 0u; // T row jA
    unsigned um_raw = 0u;                     // U row jA-1 (row a0-1 is never computed)
    double pxA = 0.0, s5A = 0.0, s7A = 0.0, pxB = 0.0, s5B = 0.0, s7B = 0.0;
    bool t1 = false, t2 = false, u1a = false, u2a = false; // stage A's T rows jA-1, jA-2 / U rows jA-2, jA-3 were active
    __syncwarp();

    for (int jA = a0; jA <= jBlast + 2; ++jA) {
        const bool tA = tm_raw != 0u;
        const bool uA = um_raw != 0u;
        const unsigned tm_next = (colT && jA + 1 <= a1) ? __ldg(tmk + pitch) : 0u;            // T row jA+1
        const unsigned um_next = (colUA && jA + 1 <= a1 && jA <= nyl) ? __ldg(umk) : 0u;      // U row jA
        tmk += pitch;
        umk += pitch;

        // ---- (1) global loads of stage A: T row jA (old state copy), U row jA-1 ---------------------------------
        TRow t;
        URow uc;
        t.act = tA;
        t.ht = false;
        t.u = t.v = t.uw = t.vw = 0.0;
        if (jA <= a1) {
            idx_t idx = jA * pitch + pc + so;
            if (colV) {
                t.u = __ldg(a.u + idx);
                t.v = __ldg(a.v + idx);
                if (pc >= 1) {
                    t.uw = __ldg(a.u + idx - 1);
                    t.vw = __ldg(a.v + idx - 1);
                }
            }
            if (tA) {
#pragma unroll
                for (int k = 0; k < EVP_NSTRESS; ++k) t.s[k] = __ldg(a.s[k] + idx);
                idx -= so;
                t.strength = __ldg(a.strength + idx);
                t.dxt = __ldg(a.dxt + idx);
                t.dyt = __ldg(a.dyt + idx);
                t.dxhy = __ldg(a.dxhy + idx);
                t.dyhx = __ldg(a.dyhx + idx);
                t.cxp = __ldg(a.cxp + idx);
                t.cyp = __ldg(a.cyp + idx);
                t.cxm = __ldg(a.cxm + idx);
                t.cym = __ldg(a.cym + idx);
                t.tiny = __ldg(a.tinyarea + idx);
            }
        }
        load_U(a, uc, pc, jA - 1, uA);

        // ---- (2) stage B: SECOND subcycle of T row jB = jA-2 and U row jB-1, operands from the ring ---------------
        const int jB = jA - 2;
        if (jB >= j0 && jB <= jBlast) {
            const double *r1 = U1 + (jB % 3) * 64, *r0 = U1 + ((jB + 2) % 3) * 64; // velocities of rows jB, jB-1
            TRow tb;
            tb.ht = false;
            tb.u = r1[osrc];
            tb.v = r1[32 + osrc];
            tb.uw = r1[wsrc];
            tb.vw = r1[32 + wsrc];
            const double usb = r0[osrc], vsb = r0[32 + osrc], uswb = r0[wsrc], vswb = r0[32 + wsrc];
            tb.act = needTB && t2;
            const idx_t pidx = jB * pitch + pc;
            double str[8];
            if (tb.act) {
                const double *sg = SIG + (jB % 3) * 384 + lane, *ti = TIN + (jB % 3) * 320 + lane;
#pragma unroll
                for (int k = 0; k < EVP_NSTRESS; ++k) tb.s[k] = sg[k * 32];
                tb.strength = ti[0 * 32];
                tb.dxt = ti[1 * 32];
                tb.dyt = ti[2 * 32];
                tb.dxhy = ti[3 * 32];
                tb.dyhx = ti[4 * 32];
                tb.cxp = ti[5 * 32];
                tb.cyp = ti[6 * 32];
                tb.cxm = ti[7 * 32];
                tb.cym = ti[8 * 32];
                tb.tiny = ti[9 * 32];
                if (LASTB) tb.tarear = __ldg(a.tarear + pidx);
                const bool store = ownT && (jB <= j1 || jB == nyl + 1);
                stress_cell<LASTB>(a, PlaneStoreT{a, sn, pidx, store}, tb, usb, vsb, uswb, vswb, str);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) str[k] = 0.0; // str(:,:,:) = c0, :1051
            }
            const double s2r = __shfl_down_sync(0xffffffffu, str[1], 1);
            const double s4r = __shfl_down_sync(0xffffffffu, str[3], 1);
            const double s7r = __shfl_down_sync(0xffffffffu, str[6], 1);
            const double s8r = __shfl_down_sync(0xffffffffu, str[7], 1);
            if (ownU && u2a && jB - 1 >= j0) {
                URow ub;
                ub.act = true;
                const double *ui = UIN + ((jB + 2) % 3) * 320 + lane; // U row jB-1
                ub.aiu = ui[0 * 32];
                ub.uocn = ui[1 * 32];
                ub.vocn = ui[2 * 32];
                ub.waterx = ui[3 * 32];
                ub.watery = ui[4 * 32];
                ub.forcex = ui[5 * 32];
                ub.forcey = ui[6 * 32];
                ub.umassdtei = ui[7 * 32];
                ub.fm = ui[8 * 32];
                ub.uarear = ui[9 * 32];
                const double sx = pxB + str[2] + s4r;         // ((s1 + s2) + s3) + s4
                const double sy = s5B + str[5] + s7B + s8r;    // ((s5 + s6) + s7) + s8
                stepu_cell<LASTB>(a, PlaneStoreU{a, sn, pidx - pitch, pc, jB - 1}, ub, usb, vsb, sx, sy);
            }
            pxB = str[0] + s2r;
            s5B = str[4];
            s7B = s7r;
        }

        // ---- (3) stage A: FIRST subcycle of T row jA and U row jA-1, results into the ring -----------------------
        if (jA <= a1 + 1) {
            double str[8];
            if (t.act) {
                double *ti = TIN + (jA % 3) * 320 + lane;
                ti[0 * 32] = t.strength;
                ti[1 * 32] = t.dxt;
                ti[2 * 32] = t.dyt;
                ti[3 * 32] = t.dxhy;
                ti[4 * 32] = t.dyhx;
                ti[5 * 32] = t.cxp;
                ti[6 * 32] = t.cyp;
                ti[7 * 32] = t.cxm;
                ti[8 * 32] = t.cym;
                ti[9 * 32] = t.tiny;
                stress_cell<false>(a, StashStoreT{SIG + (jA % 3) * 384 + lane, true}, t, us, vs, usw, vsw, str);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) str[k] = 0.0;
            }
            const double s2r = __shfl_down_sync(0xffffffffu, str[1], 1);
            const double s4r = __shfl_down_sync(0xffffffffu, str[3], 1);
            const double s7r = __shfl_down_sync(0xffffffffu, str[6], 1);
            const double s8r = __shfl_down_sync(0xffffffffu, str[7], 1);
            double u1 = us, v1 = vs; // U row jA-1: points that are not computed keep their value
            if (uc.act) {
                double *ui = UIN + ((jA + 2) % 3) * 320 + lane; // U row jA-1
                ui[0 * 32] = uc.aiu;
                ui[1 * 32] = uc.uocn;
                ui[2 * 32] = uc.vocn;
                ui[3 * 32] = uc.waterx;
                ui[4 * 32] = uc.watery;
                ui[5 * 32] = uc.forcex;
                ui[6 * 32] = uc.forcey;
                ui[7 * 32] = uc.umassdtei;
                ui[8 * 32] = uc.fm;
                ui[9 * 32] = uc.uarear;
                const double sx = pxA + str[2] + s4r;
                const double sy = s5A + str[5] + s7A + s8r;
                stepu_cell<false>(a, RegStoreU{u1, v1}, uc, us, vs, sx, sy);
            }
            double *r = U1 + ((jA + 2) % 3) * 64; // row jA-1
            r[lane] = u1;
            r[32 + lane] = v1;
            pxA = str[0] + s2r;
            s5A = str[4];
            s7A = s7r;
            us = t.u;
            vs = t.v;
            usw = t.uw;
            vsw = t.vw;
        }
        t2 = t1;
        t1 = tA;
        u2a = u1a;
        u1a = uA;
        tm_raw = tm_next;
        um_raw = um_next;
        __syncwarp(); // this iteration's ring writes are visible to the lanes that read them in the next two
    }
}

template <bool LASTB>
__global__ void __launch_bounds__(128, 2) k_subcycle2(const __grid_constant__ SubArgs a) {
    extern __shared__ double evp_xch[]; // 4 warp rings (wmarch2) or, in the tripole top chunk, march()'s exchange line
    __shared__ int is_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j0 = __ldg(a.chunks + 2 * blockIdx.y);
    const int nrows = __ldg(a.chunks + 2 * blockIdx.y + 1);
    if (nrows <= 0) return; // an empty chunk (all rows inactive, trimmed by the load balancer)
    const bool top = (j0 + nrows - 1 == a.nyl);
    const idx_t so = a.flip ? (idx_t)a.copy_stride : 0, sn = a.flip ? 0 : (idx_t)a.copy_stride;
    if (a.fold && top) {
        // Tripole top chunk: the u-fold between the two subcycles is not local, so these CTAs run the one-subcycle
        // march twice through the third state copy.  The first pass starts one row lower: it leaves the
        // first-subcycle velocities of row j0-1 and the stresses of T rows j0 .. nyl+1 for the second pass.
        const idx_t sc = 2 * (idx_t)a.copy_stride;
        const int i = 1 + blockIdx.x * a.strip_w + tid;
        march<128, false, false, true>(a, so, sc, tid, i, j0 - 1, nrows + 1);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            is_last = (atomicAdd((unsigned *)a.sync + 4, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (is_last) { // every CTA of the chunk has stored its rows: fold the intermediate velocities
            __threadfence();
            fold_top_rows<128>(a, a.u + sc, a.v + sc, tid);
            __syncthreads();
            if (tid == 0) {
                a.sync[4] = 0;
                __threadfence();
                *(volatile int *)(a.sync + 5) = 1;
            }
        }
        if (tid == 0) {
            wait_flag_ge(a.sync + 5, 1, a.sync);
            __threadfence();
        }
        __syncthreads();
        march<128, LASTB, false, true>(a, sc, sn, tid, i, j0, nrows);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            is_last = (atomicAdd((unsigned *)a.sync + 4, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (is_last) { // fold of the final velocities; every CTA of the chunk is past its wait on sync[5]
            __threadfence();
            fold_top_rows<128>(a, a.u + sn, a.v + sn, tid);
            if (tid == 0) {
                a.sync[4] = 0;
                a.sync[5] = 0;
            }
        }
    } else {
        wmarch2<LASTB>(a, so, sn, evp_xch + warp * EVF_WARP_D, lane, j0, nrows, 4 * (int)blockIdx.x + warp);
    }
}

static constexpr size_t fused_smem_bytes() { return (size_t)4 * EVF_WARP_D * sizeof(double); }

} // namespace EVP_SUB_NS

// two subcycles per launch: launch (ctas_per_sm == nullptr) or configure + occupancy query
int EVP_FUSED_LAUNCH(const SubArgs &a, bool last, int flags, unsigned grid_x, unsigned grid_y, void *stream,
                     int *ctas_per_sm) {
    using namespace EVP_SUB_NS;
    (void)flags;
    const size_t smem = fused_smem_bytes();
    static bool configured = false; // once per process (outside any stream capture)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_subcycle2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_subcycle2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (ctas_per_sm) {
        int n0 = 0, n1 = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, k_subcycle2<false>, 128, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k_subcycle2<true>, 128, smem);
        *ctas_per_sm = n0 < n1 ? n0 : n1;
        return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_x, grid_y);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    return (int)(last ? cudaLaunchKernelEx(&cfg, k_subcycle2<true>, a) : cudaLaunchKernelEx(&cfg, k_subcycle2<false>, a));
}

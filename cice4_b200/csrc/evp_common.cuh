// evp_common.cuh -- shared declarations of libevp_b200 (device layout + kernel argument blocks).
//
// Device layout ("plane"): one padded slab per field, i fastest, index
//   idx(i, j) = j * pitch + i,   i in [0, nx+1],  j in [0, nyl+1]
// where nx = nx_global, nyl = rows of this handle's y-slab, (0 / nx+1 / 0 / nyl+1) is the
// 1-cell ghost ring (domain boundary or neighbouring slab) and pitch is nx+2 rounded up to 16
// doubles.  Plane (i, j) is Fortran (i+1, j+1) of a single whole-slab block, so indices map
// 1:1 to source/ice_dyn_evp.F90.
#pragma once

#include <cstddef>
#include <cstdint>

#define EVP_NSTRESS 12
#define EVP_SYNC_MAXCX 512
#define EVP_SYNC_FN 32
#define EVP_SYNC_FS (EVP_SYNC_FN + EVP_SYNC_MAXCX)
#define EVP_SYNC_INTS (EVP_SYNC_FS + EVP_SYNC_MAXCX)

// argument block of the fused stress+stepu subcycle kernel
struct SubArgs {
    // T-cell statics (source/ice_dyn_evp.F90:992-1005)
    const double *dxt, *dyt, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym, *tinyarea, *tarear, *strength;
    // U-cell statics (:1339-1349)
    const double *aiu, *uocn, *vocn, *waterx, *watery, *forcex, *forcey, *umassdtei, *fm, *uarear;
    const uint8_t *icetmask, *iceumask;
    // the planes a T row needs, in the order the TMA-staged kernel lays them out in shared memory:
    // u, v, 12 stresses (copy 0), strength, dxt, dyt, dxhy, dyhx, cxp, cyp, cxm, cym, tinyarea, tarear
    const double *tplane[25];
    // ping-pong state: two copies of u, v and of the 12 stresses (stressp_1..4, stressm_1..4,
    // stress12_1..4).  The pointers are copy 0; copy 1 of every plane lies copy_stride doubles behind
    // it.  A one-subcycle launch reads copy `flip` and writes copy `flip ^ 1`; the persistent kernel
    // alternates by itself starting from `flip`.
    double *u, *v;
    double *s[EVP_NSTRESS];
    long long copy_stride;
    int flip;
    // written on the last subcycle only (ksub == ndte, :1103-1115; :1415-1418,:1434-1435)
    double *divu, *shear, *rdg_conv, *rdg_shear, *prs_sig, *strintx, *strinty, *strocnx, *strocny;
    // evp_finish (:1510-1547) as an epilogue of the last subcycle: 1 = the thread that produces the final u, v of a U
    // cell also completes strocnx/y and writes strocnxT/yT = strocnx/y / aiu into finx / finy (zero-filled by the
    // host, like :1512-1513); on the row that the in-kernel tripole fold rewrites, the fold does it
    int fuse_finish;
    double *finx, *finy;
    int nx, nyl, pitch;
    int ew_cyclic;
    int strip_w;   // U columns produced per CTA (threads 0..strip_w hold T columns)
    int rows;      // U rows marched per interior CTA (informative; the kernel uses `chunks`)
    // row chunks in launch order (blockIdx.y): chunks[2k] = first U row, chunks[2k+1] = number of U
    // rows.  Boundary chunks come first and are shorter (see choose_tiling in evp_abi.cu).
    const int *chunks;
    // persistent kernel (all ndte subcycles in one cooperative launch): subcycles to run; one epoch
    // per CTA (cta_epoch[blockIdx.y * gridDim.x + blockIdx.x] = subcycles of THIS launch the CTA has
    // finished; zeroed by the host before the launch); subcycles this rank had completed before the
    // launch (== sync[1] at launch; the peer-to-peer flags and the fold epoch count from there)
    int nsub;
    int *cta_epoch;
    int epoch0;
    int evp_damping, hemisphere_turning;
    // 2-plane metric path: on rows where row_ht[j] != 0 the eight metrics are re-derived from the
    // primary cell widths with the (unfused) init_grid2 formulas instead of being loaded
    const double *hte, *htn;
    const uint8_t *row_ht;
    double ecci, dte2T, denom1, denom2, rcon, dragw, cosw, sinw;
    // ---- multi-rank peer-to-peer velocity halo (exchange_mode 0); all null/0 on one rank -------
    // ghost rows of the neighbours' u/v planes (copy 0; copy 1 lies peer_*_stride doubles behind),
    // mapped through CUDA IPC: peer_n_* = row 0 of the north neighbour, peer_s_* = row nyl+1 of the
    // south neighbour
    double *peer_n_u, *peer_n_v, *peer_s_u, *peer_s_v;
    long long peer_n_stride, peer_s_stride;
    // sync block in local memory (EVP_SYNC_INTS ints): [0] CTAs finished (counter), [1] subcycles
    // completed on this rank (epoch), [4] fold counter, [5] subcycles whose tripole fold is complete
    // (persistent kernel), [6] error flag (a bounded wait gave up), then two arrays of per-strip epochs
    // written by the neighbours' boundary CTAs through their mapping of this block:
    //   [EVP_SYNC_FN + x] = epochs finished by strip x of the NORTH neighbour's southernmost chunk,
    //   [EVP_SYNC_FS + x] = epochs finished by strip x of the SOUTH neighbour's northernmost chunk.
    int *sync;
    // where this rank's boundary CTAs publish: the north neighbour's FS array, the south neighbour's FN array
    int *peer_n_flag, *peer_s_flag;
    int p2p;                         // 1 = the above are in use
    // ---- strip-tiled layout of the TMA-fed kernel (evp_tiled.cuh); null when the plane kernels run ----
    double *tiles;                   // tile row (strip w, row j) at tiles + w * t_sw + j * t_sj
    int t_ns, t_nr;                  // strips; rows per strip (nyl + 2)
    long long t_sw, t_sj;
    double *peer_n_tiles, *peer_s_tiles; // the neighbours' tile pools (same strips; their own row counts / strides)
    int peer_n_nr, peer_s_nr;
    long long peer_n_sw, peer_n_sj, peer_s_sw, peer_s_sj;
    // ---- tripole u-fold of u_new/v_new inside the kernel (top slab only) ------------------------
    int fold;              // 1 = the last CTA of the northernmost chunk to finish applies the fold
    double *fold_scratch;  // 2 * pitch doubles: copy of the raw top physical row of u_new, v_new
    // ---- two subcycles per launch (evp_fused.cuh) ------------------------------------------------
    // per warp strip (4 per CTA, index 4 * blockIdx.x + warp) four ints: virtual column of lane 0; first | last << 8
    // owned lane; lane that holds the ghost T column nx+1 next to the east-west wrap (-1: none); first | last << 8
    // lane in use.  A third state copy (2 * copy_stride behind copy 0) holds the intermediate rows of the tripole
    // top chunk.
    const int *wstrips;
};

// Tripole u-fold of a NE-corner vector field (serial/ice_boundary.F90:777-800 symmetrisation,
// :837-866 copy-out) for plane column i in [0, nx+1]: `top` is the raw top physical row (row nyl),
// `below` row nyl-1.  Returns the new values of rows nyl (vtop) and nyl+1 (vghost).
__host__ __device__ inline void evp_fold_necorner(const double *top, const double *below, int i, int nx,
                                                  int ew_cyclic, double isign, double &vtop, double &vghost) {
    int ig = i; // i_glob of the column (source/ice_blocks.F90:291-330)
    if (i == 0) ig = ew_cyclic ? nx : 1;
    if (i == nx + 1) ig = ew_cyclic ? 1 : nx;
    int k = nx - ig; // iSrc = nxGlobal - i_glob + 1 - ioffset, ioffset = 1
    if (k == 0) k = nx;
    double x;
    if (k >= 1 && k <= nx / 2 - 1) {
        x = 0.5 * (top[k] + isign * top[nx - k]);
    } else if (k >= nx - (nx / 2 - 1) && k <= nx - 1) {
        x = isign * (0.5 * (top[nx - k] + isign * top[k])); // partner of the loop index i = nx - k
    } else {
        x = top[k];
    }
    vtop = isign * x;          // j=1: row jhi   <- isign*buf(iSrc, 2)
    vghost = isign * below[k]; // j=2: row jhi+1 <- isign*buf(iSrc, 1)
}

// returns a cudaError_t value (the launch status)
typedef int (*subcycle_launch_fn)(const SubArgs &a, bool last, int variant, int threads,
                                  unsigned grid_x, unsigned grid_y, void *stream);

// defined in evp_subcycle_strict.cu (-fmad=false) and evp_subcycle_fast.cu (-fmad=true)
int evp_subcycle_launch_strict(const SubArgs &a, bool last, int variant, int threads,
                               unsigned grid_x, unsigned grid_y, void *stream);
int evp_subcycle_launch_fast(const SubArgs &a, bool last, int variant, int threads,
                             unsigned grid_x, unsigned grid_y, void *stream);
// persistent kernel: a.nsub subcycles in one cooperative launch (ctas_per_sm == nullptr), or, with
// ctas_per_sm != nullptr, only the occupancy query (resident CTAs per SM; the grid must fit in one
// wave).  Returns a cudaError_t value.
typedef int (*persist_launch_fn)(const SubArgs &a, int threads, unsigned grid_x, unsigned grid_y, void *stream,
                                 int *ctas_per_sm);
int evp_persist_launch_strict(const SubArgs &a, int threads, unsigned grid_x, unsigned grid_y, void *stream,
                              int *ctas_per_sm);
int evp_persist_launch_fast(const SubArgs &a, int threads, unsigned grid_x, unsigned grid_y, void *stream,
                            int *ctas_per_sm);
int evp_subcycle_configure_strict(void);
int evp_subcycle_configure_fast(void);
// strip-tiled TMA-fed kernel (k_subcycle_tiled): stages = 2 (3 CTAs of 4 warps per SM) or 3 (2 CTAs per SM);
// ctas_per_sm != nullptr: only the occupancy query.  Returns a cudaError_t value.
// flags: bit 0 programmatic dependent launch, bit 1 memory-only measurement variant (results invalid)
typedef int (*tiled_launch_fn)(const SubArgs &a, bool last, int stages, int flags, unsigned grid_x, unsigned grid_y,
                               void *stream, int *ctas_per_sm);
int evp_tiled_launch_strict(const SubArgs &a, bool last, int stages, int flags, unsigned grid_x, unsigned grid_y,
                            void *stream, int *ctas_per_sm);
int evp_tiled_launch_fast(const SubArgs &a, bool last, int stages, int flags, unsigned grid_x, unsigned grid_y,
                          void *stream, int *ctas_per_sm);
// two subcycles per launch (k_subcycle2): reads state copy a.flip, writes copy a.flip ^ 1; `last`: the second of
// the two is subcycle ndte.  ctas_per_sm != nullptr: only configure + occupancy query.  Returns a cudaError_t value.
typedef int (*fused_launch_fn)(const SubArgs &a, bool last, int flags, unsigned grid_x, unsigned grid_y, void *stream,
                               int *ctas_per_sm);
int evp_fused_launch_strict(const SubArgs &a, bool last, int flags, unsigned grid_x, unsigned grid_y, void *stream,
                            int *ctas_per_sm);
int evp_fused_launch_fast(const SubArgs &a, bool last, int flags, unsigned grid_x, unsigned grid_y, void *stream,
                          int *ctas_per_sm);
int evp_subcycle_max_threads(void);

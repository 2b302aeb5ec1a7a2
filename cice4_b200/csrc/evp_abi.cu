// evp_abi.cu -- the C ABI of libevp_b200.so (include/evp_b200.h): handle, device memory,
// marshalling and the evp() sequence of source/ice_dyn_evp.F90:119-432 on one B200.
// There is no CPU fallback anywhere in this file: every compute entry needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types only: libnccl.so.2 is dlopen()ed by evp_b200_comm_init, never linked

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/evp_b200.h"
#include "evp_aux.cuh"
#include "evp_common.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// default of the plane kernels with 128-thread CTAs: one strip per warp (1) or one strip per CTA (0)
#ifndef EVP_WARPX_DEFAULT
#define EVP_WARPX_DEFAULT 1
#endif

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(EVP_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// names of the double planes held on the device
enum PlaneId {
    // static grid
    P_DXT, P_DYT, P_DXHY, P_DYHX, P_CXP, P_CYP, P_CXM, P_CYM, P_TAREA, P_TAREAR, P_TINYAREA,
    P_UAREA, P_UAREAR, P_FCOR, P_HTE, P_HTN,
    // inputs
    P_AICE, P_VICE, P_VSNO, P_UOCN, P_VOCN, P_SSTLTX, P_SSTLTY, P_AICE0,
    // scratch (locals of evp, :170-178) and work1 of ice_work
    P_TMASS, P_UMASS, P_AIU, P_UMASSDTEI, P_WATERX, P_WATERY, P_FORCEX, P_FORCEY, P_WRKX, P_WRKY,
    // outputs
    P_STRAIRX, P_STRAIRY, P_STRTLTX, P_STRTLTY, P_STRINTX, P_STRINTY, P_STROCNX, P_STROCNY,
    P_STROCNXT, P_STROCNYT, P_FM, P_PRS_SIG, P_DIVU, P_SHEAR, P_RDG_CONV, P_RDG_SHEAR, P_STRENGTH,
    P_SIG1, P_SIG2,
    // ping-pong state: u, v, 12 stresses, two copies each
    // state copy 0 (u, v, 12 stresses) and, EVP_STATE_PLANES planes behind each, copy 1 (SubArgs::copy_stride)
    // copy 2 (2 * copy_stride behind copy 0): intermediate rows of the tripole top chunk of the two-subcycle kernel
    P_U0, P_V0, P_S0, P_U1 = P_S0 + EVP_NSTRESS, P_V1, P_S1, P_U2 = P_S1 + EVP_NSTRESS, P_V2, P_S2,
    P_COUNT = P_S2 + EVP_NSTRESS
};
enum MaskId { M_TMASK, M_UMASK, M_TMPHM, M_ICETMASK, M_ICEUMASK, M_COUNT };

struct Timer {
    cudaEvent_t ev[8];
};

} // namespace

struct evp_b200_handle {
    evp_b200_dims dims;
    evp_b200_params par;
    std::vector<int32_t> blk_tab; // host copy of the per-block table
    int *d_blk_tab = nullptr;
    BlockGeom bg;
    PlaneGeom pg;
    size_t blocked_elems = 0; // nx_block*ny_block*max_blocks
    // derived scalars (set_evp_parameters, source/ice_dyn_evp.F90:563-575)
    double dtei, ecci, dte2T, denom1, denom2, rcon, dragw;
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;       // device-to-host copies (and the early outputs' pack kernels)
    cudaStream_t st_up = nullptr;     // host-to-device copies: the copy engine runs ahead of the unpack kernels on st
    cudaEvent_t ring[128] = {};       // events that order a copy on st_up / st2 against its kernel on st
    int ring_pos = 0;
    bool dn_pending = false;          // copies on st2 that st has not been ordered after yet
    cudaEvent_t ev_early = nullptr, ev_early_done = nullptr;
    double *pool = nullptr;
    double *pl[P_COUNT];
    double *cat = nullptr; // aicen, vicen planes (2*ncat), allocated on first device ice_strength
    uint8_t *mpool = nullptr;
    uint8_t *mk[M_COUNT];
    // staging in the caller's block layout
    double *stage = nullptr; // n_stage slots of blocked_elems doubles
    int n_stage = 0;
    int32_t *stage_i = nullptr; // 2 slots of blocked_elems int32
    double *stage_cat = nullptr;
    cudaEvent_t ev[8];
    cudaGraph_t graph[2] = {nullptr, nullptr};            // ndte loop starting from state copy 0 / 1
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    // strip-tiled layout of the subcycle loop (evp_tiled.cuh) and the TMA-fed kernel that runs on it
    bool tiled = false;               // the ndte loop runs k_subcycle_tiled on `tiles`
    int tiled_stages = 2;             // pipeline depth per warp: 2 (3 CTAs per SM) or 3 (2 CTAs per SM)
    TileGeom tg = {nullptr, 0, 0, 0, 0};
    bool tiles_static = false;        // static part of the loop-invariant block packed
    bool planes_stale = false;        // the current state lives in the tiles only (after evp_b200_subcycle_resident)
    double *peer_tiles[2] = {nullptr, nullptr};
    int cur = 0; // which ping-pong copy holds the current state
    bool prepared = false, resident = false;
    bool stress_on_device = false;    // state_residency = 1: copy 0 of the stress planes is current
    evp_b200_timings tm;
    std::unordered_map<const void *, size_t> pinned;
    bool pin_enabled = false;         // false during evp_b200_init (static fields are not pinned)
    bool io_device = false;           // evp_b200_step_device: the caller's arrays are DEVICE memory (block layout)
    bool vel_on_device = false;       // state_residency = 2: uvel, vvel, iceumask of the planes are current
    int grid_x = 0, grid_y = 0, threads = 128, strip_w = 0, rows = 0;
    bool warpx = false; // plane kernels with one strip per warp (shuffles, no row barrier)
    int *d_chunks = nullptr;          // row-chunk table of the subcycle kernel (2 ints per chunk)
    int *d_cta_epoch = nullptr;       // persistent kernel: subcycles finished per CTA (grid_x * grid_y ints)
    bool persistent = false;          // run the ndte loop as one cooperative launch (k_persist)
    bool fused = false;               // two subcycles per launch (k_subcycle2, evp_fused.cuh)
    int *d_wstrips = nullptr;         // its per-warp strip table (4 ints per warp strip)
    int epoch_count = 0;              // subcycles completed on this rank since init (== sync[1] with p2p)
    long eliminated_cells = 0;        // cells of the slab without a block (eliminated land blocks)
    int *d_rowcnt = nullptr;          // active T cells per row (load balance of the chunks)
    bool balance = false;             // rebuild the chunk table from icetmask every call
    float w_bot = 1.f, w_top = 1.f;   // relative cost targets of the boundary chunks
    int sub_launches_per_loop = 0;
    // y-slab chain: neighbour ranks (-1 = none) and the NCCL communicator (dlopen()ed entry points)
    int north = -1, south = -1;
    void *nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*pGroupStart)() = nullptr;
    ncclResult_t (*pGroupEnd)() = nullptr;
    ncclResult_t (*pSend)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pRecv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
    const char *(*pGetErrorString)(ncclResult_t) = nullptr;
    // peer-to-peer halo (exchange_mode 0): neighbours' plane pools and sync blocks mapped via CUDA IPC
    int *sync = nullptr;              // local sync block (64 ints): see SubArgs::sync
    uint8_t *row_ht = nullptr;        // per-row flag of the 2-plane metric path (nullptr = off)
    int rows_ht = 0;                  // rows on which it is active
    double *d_energy = nullptr;       // scratch of evp_b200_diagnostics_energy
    double *fold_scratch = nullptr;   // 2 * pitch doubles
    bool fold_in_kernel = false;      // tripole fold done by the subcycle kernel (else k_halo_tripole)
    bool p2p = false;
    double *peer_pool[2] = {nullptr, nullptr}; // [0] north, [1] south
    int *peer_sync[2] = {nullptr, nullptr};
    int peer_nyl[2] = {0, 0};
    size_t peer_cells[2] = {0, 0};
};

namespace {

// Page-lock a caller array on first use (pin_host = 1).  Only the per-call arrays are pinned -- the state, input
// and output arrays of evp(), which the caller keeps for the whole run (Fortran module arrays; evp_b200_unpin
// otherwise): the static grid fields of evp_b200_init are uploaded once from pageable memory, because their
// arrays need not outlive the call and a registration left on freed memory poisons whatever the allocator
// places there next (a partly registered range makes cudaMemcpyAsync fail with "invalid argument").
void pin(evp_b200_handle *h, const void *p, size_t bytes) {
    if (!h->par.pin_host || !h->pin_enabled || !p) return;
    auto it = h->pinned.find(p);
    if (it != h->pinned.end() && it->second >= bytes) return;
    if (it != h->pinned.end()) cudaHostUnregister(const_cast<void *>(p));
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess)
        h->pinned[p] = bytes;
    else
        cudaGetLastError(); // pageable copy still works
}

cudaEvent_t next_event(evp_b200_handle *h) {
    cudaEvent_t &e = h->ring[h->ring_pos];
    h->ring_pos = (h->ring_pos + 1) % 128;
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}

// Host arrays reach the planes in two steps on the handle's stream: the copy into the staging slot and the unpack
// kernel.  (Running the copies on a second stream ahead of the kernels gained 0.3 ms of a 28 ms call and was
// taken out again.)
int upload_r8(evp_b200_handle *h, const double *host, int slot, double *plane) {
    if (!host) return fail(EVP_B200_ERR_ARG, "required host array is NULL (stage slot %d)", slot);
    if (h->io_device) { // device-pointer hand-off: the block array is read where it lies
        aux_unblock_r8(h->bg, h->pg, host, plane, h->st);
        return 0;
    }
    double *stg = h->stage + (size_t)slot * h->blocked_elems;
    const size_t bytes = h->blocked_elems * sizeof(double);
    pin(h, host, bytes);
    CU(cudaMemcpyAsync(stg, host, bytes, cudaMemcpyHostToDevice, h->st));
    aux_unblock_r8(h->bg, h->pg, stg, plane, h->st);
    return 0;
}

int upload_mask(evp_b200_handle *h, const int32_t *host, int slot, uint8_t *plane) {
    if (!host) return fail(EVP_B200_ERR_ARG, "required host mask is NULL");
    if (h->io_device) {
        aux_unblock_mask(h->bg, h->pg, host, plane, h->st);
        return 0;
    }
    int32_t *stg = h->stage_i + (size_t)slot * h->blocked_elems;
    const size_t bytes = h->blocked_elems * sizeof(int32_t);
    pin(h, host, bytes);
    CU(cudaMemcpyAsync(stg, host, bytes, cudaMemcpyHostToDevice, h->st));
    aux_unblock_mask(h->bg, h->pg, stg, plane, h->st);
    return 0;
}

// fresh = the staging slot does not hold the caller's current values: fill it from the host
// first for KEEP policies (not needed for state fields, whose slot was uploaded this call)
int download_r8(evp_b200_handle *h, double *host, int slot, const double *plane, int policy,
                cudaStream_t s = nullptr) {
    if (!host) return 0;
    if (!s) s = h->st;
    if (h->io_device) { // the KEEP policies leave the caller's other cells as they are
        aux_block_r8(h->bg, h->pg, plane, h->mk[M_ICETMASK], host, policy, s);
        if (s != h->st) h->dn_pending = true;
        return 0;
    }
    double *stg = h->stage + (size_t)slot * h->blocked_elems;
    const size_t bytes = h->blocked_elems * sizeof(double);
    pin(h, host, bytes);
    aux_block_r8(h->bg, h->pg, plane, h->mk[M_ICETMASK], stg, policy, s);
    CU(cudaMemcpyAsync(host, stg, bytes, cudaMemcpyDeviceToHost, s));
    if (s != h->st) h->dn_pending = true;
    return 0;
}

// the early outputs travel on st2 while the ndte loop runs: order them into st (no-op when there were none)
int join_downloads(evp_b200_handle *h) {
    if (!h->dn_pending) return 0;
    cudaEvent_t e = next_event(h);
    CU(cudaEventRecord(e, h->st2));
    CU(cudaStreamWaitEvent(h->st, e, 0));
    h->dn_pending = false;
    return 0;
}

int download_mask(evp_b200_handle *h, int32_t *host, int slot, const uint8_t *plane, int policy) {
    if (!host) return 0;
    if (h->io_device) {
        aux_block_mask(h->bg, h->pg, plane, host, policy, h->st);
        return 0;
    }
    int32_t *stg = h->stage_i + (size_t)slot * h->blocked_elems;
    const size_t bytes = h->blocked_elems * sizeof(int32_t);
    pin(h, host, bytes);
    aux_block_mask(h->bg, h->pg, plane, stg, policy, h->st);
    CU(cudaMemcpyAsync(host, stg, bytes, cudaMemcpyDeviceToHost, h->st));
    return 0;
}

// staging slots
enum {
    SL_AICE, SL_VICE, SL_VSNO, SL_STRAIRX, SL_STRAIRY, SL_UOCN, SL_VOCN, SL_SSTLTX, SL_SSTLTY, SL_AICE0,
    SL_STRENGTH, SL_U, SL_V, SL_S0, SL_OUT0 = SL_S0 + EVP_NSTRESS, SL_COUNT = SL_OUT0 + 20
};

// Slab-to-slab part of ice_HaloUpdate (replaces the MPI_ISEND/IRECV of mpi/ice_boundary.F90:1145-1215
// between tasks): whole padded rows -- columns 0 and nx+1 already hold the east-west wrap, so the
// corner cells travel with the row -- straight out of / into the planes, one grouped NCCL call for
// all planes.  Row nyl -> north neighbour's row 0, row 1 -> south neighbour's row nyl+1.
int exchange_rows(evp_b200_handle *h, void *const *planes, int n, size_t elem) {
    if (h->dims.nranks == 1) return 0;
    if (!h->comm) return fail(EVP_B200_ERR_STATE, "nranks > 1 but evp_b200_comm_init has not been called");
    const PlaneGeom &pg = h->pg;
    const size_t rowb = (size_t)pg.pitch * elem, cnt = (size_t)(pg.nx + 2) * elem;
    ncclResult_t r = h->pGroupStart();
    for (int k = 0; k < n && r == ncclSuccess; ++k) {
        char *p = (char *)planes[k];
        if (h->north >= 0) {
            r = h->pSend(p + rowb * pg.nyl, cnt, ncclChar, h->north, h->comm, h->st);
            if (r == ncclSuccess) r = h->pRecv(p + rowb * (pg.nyl + 1), cnt, ncclChar, h->north, h->comm, h->st);
        }
        if (h->south >= 0 && r == ncclSuccess) {
            r = h->pSend(p + rowb, cnt, ncclChar, h->south, h->comm, h->st);
            if (r == ncclSuccess) r = h->pRecv(p, cnt, ncclChar, h->south, h->comm, h->st);
        }
    }
    ncclResult_t r2 = h->pGroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return fail(EVP_B200_ERR_COMM, "NCCL row exchange failed: %s", h->pGetErrorString(r));
    return 0;
}

// sync[6]: a bounded flag wait inside the subcycle kernels (or k_wait_peers) gave up.  Reported once and
// cleared, so that one time-out does not poison every later call; the epochs of a multi-rank chain are
// undefined after it, so the caller has to re-initialise all ranks.  Synchronises the stream.
int check_wait_flag(evp_b200_handle *h) {
    int gave_up = 0;
    CU(cudaMemcpyAsync(&gave_up, h->sync + 6, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (gave_up) {
        CU(cudaMemsetAsync(h->sync + 6, 0, sizeof(int), h->st));
        CU(cudaStreamSynchronize(h->st));
        return fail(EVP_B200_ERR_STATE, "subcycle kernel: a wait for a neighbouring CTA / GPU timed out (results invalid%s)",
                    h->dims.nranks > 1 ? "; re-initialise every rank of the chain" : "");
    }
    return 0;
}

int halo_r8(evp_b200_handle *h, double *plane, int loc, int isign) {
    aux_halo_r8(h->pg, plane, loc, isign, h->st);
    void *pp[1] = {plane};
    return exchange_rows(h, pp, 1, sizeof(double));
}

// evp_finish (:1510-1547) runs as an epilogue of the last subcycle kernel (stepu's thread completes strocnx/y and
// writes strocnxT/yT's input; the in-kernel tripole fold does it for the row it rewrites) unless the velocities of
// physical cells still change after that kernel -- a tripole fold outside the kernel (T-fold, kernel_variant bit 2,
// tiny slabs) -- or the two-subcycle kernel runs; kernel_variant bit 22 (4194304) keeps the separate k_finish launch.
static bool fuse_finish(const evp_b200_handle *h) {
    if (h->par.kernel_variant & 4194304) return false;
    if (h->fused) return false;
    return !h->pg.tripole || h->fold_in_kernel;
}

void fill_subargs(evp_b200_handle *h, SubArgs &a, int cur) {
    double **p = h->pl;
    a.dxt = p[P_DXT]; a.dyt = p[P_DYT]; a.dxhy = p[P_DXHY]; a.dyhx = p[P_DYHX];
    a.cxp = p[P_CXP]; a.cyp = p[P_CYP]; a.cxm = p[P_CXM]; a.cym = p[P_CYM];
    a.tinyarea = p[P_TINYAREA]; a.tarear = p[P_TAREAR]; a.strength = p[P_STRENGTH];
    a.aiu = p[P_AIU]; a.uocn = p[P_UOCN]; a.vocn = p[P_VOCN]; a.waterx = p[P_WATERX]; a.watery = p[P_WATERY];
    a.forcex = p[P_FORCEX]; a.forcey = p[P_FORCEY]; a.umassdtei = p[P_UMASSDTEI]; a.fm = p[P_FM];
    a.uarear = p[P_UAREAR];
    a.icetmask = h->mk[M_ICETMASK]; a.iceumask = h->mk[M_ICEUMASK];
    static_assert(P_U1 - P_U0 == P_V1 - P_V0 && P_U1 - P_U0 == P_S1 - P_S0, "state copies must be equidistant");
    static_assert(P_U2 - P_U1 == P_U1 - P_U0 && P_V2 - P_V1 == P_U1 - P_U0 && P_S2 - P_S1 == P_U1 - P_U0, "state copies must be equidistant");
    a.u = p[P_U0];
    a.v = p[P_V0];
    for (int k = 0; k < EVP_NSTRESS; ++k) a.s[k] = p[P_S0 + k];
    a.copy_stride = (long long)(P_U1 - P_U0) * (long long)h->pg.cells;
    {
        const double *tp[25] = {a.u, a.v, a.s[0], a.s[1], a.s[2], a.s[3], a.s[4], a.s[5], a.s[6], a.s[7], a.s[8],
                                a.s[9], a.s[10], a.s[11], a.strength, a.dxt, a.dyt, a.dxhy, a.dyhx, a.cxp, a.cyp,
                                a.cxm, a.cym, a.tinyarea, a.tarear};
        for (int k = 0; k < 25; ++k) a.tplane[k] = tp[k];
    }
    a.flip = cur ? 1 : 0; // copy that is read; copy flip ^ 1 is written
    a.nsub = 1;
    a.cta_epoch = h->d_cta_epoch;
    a.epoch0 = h->epoch_count;
    a.divu = p[P_DIVU]; a.shear = p[P_SHEAR]; a.rdg_conv = p[P_RDG_CONV]; a.rdg_shear = p[P_RDG_SHEAR];
    a.prs_sig = p[P_PRS_SIG]; a.strintx = p[P_STRINTX]; a.strinty = p[P_STRINTY];
    a.strocnx = p[P_STROCNX]; a.strocny = p[P_STROCNY];
    a.fuse_finish = fuse_finish(h) ? 1 : 0; a.finx = p[P_WRKX]; a.finy = p[P_WRKY];
    a.nx = h->pg.nx; a.nyl = h->pg.nyl; a.pitch = h->pg.pitch;
    a.ew_cyclic = h->pg.ew_cyclic;
    a.strip_w = h->strip_w; a.rows = h->rows; a.chunks = h->d_chunks;
    a.evp_damping = h->par.evp_damping; a.hemisphere_turning = h->par.hemisphere_turning;
    a.hte = p[P_HTE]; a.htn = p[P_HTN];
    // The 2-plane metric path removes 6 of 48 words of DRAM traffic.  In a burst the kernel is not purely
    // bandwidth-bound and does not run faster for it, but sustained at the box's power cap the saved DRAM power goes
    // to the SM clock: 113.4 instead of 116.0 us per subcycle at 1440 x 1080 inside bench.py's timed region
    // (profiles/r02_experiments.txt 7).  Default on the plane kernels of tall slabs (>= 450 rows; short slabs are
    // issue-bound, where the ~14 extra fp64 operations per cell cost more than the bytes save) when HTE / HTN were
    // given and verified; kernel_variant bit 4 forces it on, bit 19 off.
    const bool ht_auto = !h->tiled && !h->fused && h->dims.ny_global / h->dims.nranks >= 450 &&
                         (h->par.kernel_variant & (256 | 512 | 524288)) == 0 &&
                         ((h->par.kernel_variant & 1024) == 0 || h->warpx); // late loads: only the warp-strip instances have it
    a.row_ht = ((h->par.kernel_variant & 16) || ht_auto) ? h->row_ht : nullptr;
    a.ecci = h->ecci; a.dte2T = h->dte2T; a.denom1 = h->denom1; a.denom2 = h->denom2; a.rcon = h->rcon;
    a.dragw = h->dragw; a.cosw = h->par.cosw; a.sinw = h->par.sinw;
    a.fold = h->fold_in_kernel ? 1 : 0;
    a.fold_scratch = h->fold_scratch;
    a.wstrips = h->d_wstrips;
    a.peer_n_u = a.peer_n_v = a.peer_s_u = a.peer_s_v = nullptr;
    a.peer_n_stride = a.peer_s_stride = 0;
    a.peer_n_flag = a.peer_s_flag = nullptr;
    a.sync = h->sync;
    a.p2p = h->p2p ? 1 : 0;
    a.tiles = h->tiled ? h->tg.tiles : nullptr;
    a.t_ns = h->tg.ns;
    a.t_nr = h->tg.nr;
    a.t_sw = h->tg.sw;
    a.t_sj = h->tg.sj;
    a.peer_n_tiles = a.peer_s_tiles = nullptr;
    a.peer_n_nr = a.peer_s_nr = 0;
    a.peer_n_sw = a.peer_n_sj = a.peer_s_sw = a.peer_s_sj = 0;
    if (h->p2p && h->tiled) {
        const int row_major = h->tg.sw == EVT_ROW_D; // every rank of a chain uses the same order
        TileGeom pgm = h->tg;
        if (h->north >= 0) {
            pgm.nr = h->peer_nyl[0] + 2;
            evt_set_order(pgm, row_major);
            a.peer_n_tiles = h->peer_tiles[0]; a.peer_n_nr = pgm.nr; a.peer_n_sw = pgm.sw; a.peer_n_sj = pgm.sj;
        }
        if (h->south >= 0) {
            pgm.nr = h->peer_nyl[1] + 2;
            evt_set_order(pgm, row_major);
            a.peer_s_tiles = h->peer_tiles[1]; a.peer_s_nr = pgm.nr; a.peer_s_sw = pgm.sw; a.peer_s_sj = pgm.sj;
        }
    }
    if (h->p2p) {
        // the neighbours' pools are laid out like ours (same plane ids, their own plane size)
        if (h->north >= 0) { // north neighbour's south ghost row = its row 0
            a.peer_n_u = h->peer_pool[0] + (size_t)P_U0 * h->peer_cells[0];
            a.peer_n_v = h->peer_pool[0] + (size_t)P_V0 * h->peer_cells[0];
            a.peer_n_stride = (long long)(P_U1 - P_U0) * (long long)h->peer_cells[0];
            a.peer_n_flag = h->peer_sync[0] + EVP_SYNC_FS;
        }
        if (h->south >= 0) { // south neighbour's north ghost row = its row nyl+1
            const size_t off = (size_t)(h->peer_nyl[1] + 1) * h->pg.pitch;
            a.peer_s_u = h->peer_pool[1] + (size_t)P_U0 * h->peer_cells[1] + off;
            a.peer_s_v = h->peer_pool[1] + (size_t)P_V0 * h->peer_cells[1] + off;
            a.peer_s_stride = (long long)(P_U1 - P_U0) * (long long)h->peer_cells[1];
            a.peer_s_flag = h->peer_sync[1] + EVP_SYNC_FN;
        }
    }
}

// one subcycle: fused stress+stepu (+ east-west halo), then the north-south part of
// ice_HaloUpdate(uvel), ice_HaloUpdate(vvel) (:397-402).  Returns kernels launched.
int launch_subcycle(evp_b200_handle *h, int cur, bool last) {
    SubArgs a;
    fill_subargs(h, a, cur);
    int e;
    if (h->tiled) {
        tiled_launch_fn fn = h->par.math_mode == 1 ? evp_tiled_launch_fast : evp_tiled_launch_strict;
        const int flags = ((h->par.kernel_variant & 64) ? 1 : 0) | ((h->par.kernel_variant & 16384) ? 2 : 0);
        e = fn(a, last, h->tiled_stages, flags, (unsigned)h->grid_x, (unsigned)h->grid_y, (void *)h->st, nullptr);
    } else {
        subcycle_launch_fn fn = h->par.math_mode == 1 ? evp_subcycle_launch_fast : evp_subcycle_launch_strict;
        const int kv = (h->par.kernel_variant & ~1048576) | (h->warpx ? 1048576 : 0);
        e = fn(a, last, kv, h->threads, (unsigned)h->grid_x, (unsigned)h->grid_y, (void *)h->st);
    }
    if (e != 0) {
        fail(EVP_B200_ERR_CUDA, "subcycle kernel launch: %s", cudaGetErrorString((cudaError_t)e));
        return -1;
    }
    int n = 1;
    if (h->tiled) return n; // east-west wrap, tripole fold and slab-to-slab rows are all inside the kernel
    if (h->pg.ns_cyclic) n += 2;
    if (h->pg.tripole && !h->fold_in_kernel) n += 1;
    double *u_new = a.flip ? a.u : a.u + a.copy_stride, *v_new = a.flip ? a.v : a.v + a.copy_stride;
    if (!h->fold_in_kernel) aux_halo_uv_ns(h->pg, u_new, v_new, h->st);
    if (h->dims.nranks > 1 && !h->p2p) {
        void *pp[2] = {u_new, v_new};
        if (exchange_rows(h, pp, 2, sizeof(double))) return -1;
        n += 1;
    }
    return n;
}

// two subcycles in one launch (k_subcycle2): east-west wrap and tripole fold of both are inside the kernel
int launch_fused(evp_b200_handle *h, int cur, bool last) {
    SubArgs a;
    fill_subargs(h, a, cur);
    fused_launch_fn fn = h->par.math_mode == 1 ? evp_fused_launch_fast : evp_fused_launch_strict;
    const int e = fn(a, last, 0, (unsigned)h->grid_x, (unsigned)h->grid_y, (void *)h->st, nullptr);
    if (e != 0) {
        fail(EVP_B200_ERR_CUDA, "two-subcycle kernel launch: %s", cudaGetErrorString((cudaError_t)e));
        return -1;
    }
    return 1;
}

// the launches of one ndte loop starting from state copy `cur`; returns the number of kernels (< 0: error) and
// leaves the copy that holds the result in `cur`
int launch_loop(evp_b200_handle *h, int &cur) {
    const int ndte = h->par.ndte;
    int n = 0, k = 1;
    if (h->fused) {
        if (ndte & 1) { // an odd ndte starts with one single subcycle
            const int m = launch_subcycle(h, cur, ndte == 1);
            if (m < 0) return -1;
            n += m;
            cur ^= 1;
            k = 2;
        }
        for (; k + 1 <= ndte; k += 2) {
            const int m = launch_fused(h, cur, k + 1 == ndte);
            if (m < 0) return -1;
            n += m;
            cur ^= 1;
        }
        return n;
    }
    for (; k <= ndte; ++k) {
        const int m = launch_subcycle(h, cur, k == ndte);
        if (m < 0) return -1;
        n += m;
        cur ^= 1;
    }
    return n;
}

int run_subcycle_loop(evp_b200_handle *h) {
    const int ndte = h->par.ndte;
    // the plane kernels always start from state copy 0 (an odd ndte is copied back below); the tiled
    // kernel starts from whichever copy is current
    if (h->cur != 0 && !h->tiled) return fail(EVP_B200_ERR_STATE, "subcycle loop must start from state copy 0");
    const int c0 = h->cur;
    // NCCL send/recv inside a captured graph dead-locked on 2 x B200 (NCCL 2.28.9): with the NCCL
    // exchange the loop is launched on the stream; the peer-to-peer exchange has no host calls
    const bool graph_ok = h->par.use_graph && (h->dims.nranks == 1 || h->p2p);
    if (h->persistent) {
        // all ndte subcycles in one cooperative launch (k_persist): CTAs synchronise with their
        // neighbours only, through per-CTA epochs that count from zero in every launch
        SubArgs a;
        fill_subargs(h, a, 0);
        a.nsub = ndte;
        CU(cudaMemsetAsync(h->d_cta_epoch, 0, sizeof(int) * (size_t)h->grid_x * h->grid_y, h->st));
        persist_launch_fn fn = h->par.math_mode == 1 ? evp_persist_launch_fast : evp_persist_launch_strict;
        const int e = fn(a, h->threads, (unsigned)h->grid_x, (unsigned)h->grid_y, (void *)h->st, nullptr);
        if (e != 0) return fail(EVP_B200_ERR_CUDA, "persistent subcycle kernel: %s", cudaGetErrorString((cudaError_t)e));
        h->sub_launches_per_loop = 1;
        h->cur = ndte & 1;
    } else if (graph_ok) {
        if (!h->graph_exec[c0]) {
            CU(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            int cur = c0;
            const int n = launch_loop(h, cur);
            const bool bad = n < 0;
            h->sub_launches_per_loop = n;
            const std::string why = g_err;
            CU(cudaStreamEndCapture(h->st, &h->graph[c0]));
            if (bad) {
                cudaGraphDestroy(h->graph[c0]);
                h->graph[c0] = nullptr;
                cudaGetLastError();
                g_err = why;
                return EVP_B200_ERR_CUDA;
            }
            CU(cudaGraphInstantiate(&h->graph_exec[c0], h->graph[c0], 0));
        }
        CU(cudaGraphLaunch(h->graph_exec[c0], h->st));
        h->cur = c0 ^ ((h->fused ? (ndte + 1) / 2 : ndte) & 1);
    } else {
        const int n = launch_loop(h, h->cur);
        if (n < 0) return EVP_B200_ERR_CUDA;
        h->sub_launches_per_loop = n;
    }
    CU(cudaGetLastError());
    h->epoch_count += ndte;
    // peer-to-peer halo: the ghost rows of the final copy are complete once both neighbours have
    // published the epoch of their last subcycle kernel
    if (h->p2p) aux_wait_peers(h->sync, h->north >= 0, h->south >= 0, h->tiled ? h->tg.ns : h->grid_x, h->st);
    if (h->tiled) {
        h->planes_stale = true; // the result is in tile copy h->cur
        return 0;
    }
    if (h->cur != 0) { // odd ndte: bring the result back to copy 0 so the next loop starts there
        const size_t bytes = h->pg.cells * sizeof(double);
        if (h->p2p) {
            // The neighbours have seen this rank's last epoch and may already run their next loop, whose
            // first kernel stores into THIS rank's copy-1 ghost rows: copy the rows this rank owns only,
            // and fetch the ghost rows of copy 0 from the neighbours' copy 0 (two-sided, so it is also
            // ordered after their copy-back).
            const size_t off = (size_t)h->pg.pitch, rows = (size_t)h->pg.pitch * h->pg.nyl * sizeof(double);
            CU(cudaMemcpyAsync(h->pl[P_U0] + off, h->pl[P_U1] + off, rows, cudaMemcpyDeviceToDevice, h->st));
            CU(cudaMemcpyAsync(h->pl[P_V0] + off, h->pl[P_V1] + off, rows, cudaMemcpyDeviceToDevice, h->st));
            if (h->pg.tripole) { // the fold's ghost row belongs to this rank
                const size_t g = (size_t)h->pg.pitch * (h->pg.nyl + 1), gb = (size_t)h->pg.pitch * sizeof(double);
                CU(cudaMemcpyAsync(h->pl[P_U0] + g, h->pl[P_U1] + g, gb, cudaMemcpyDeviceToDevice, h->st));
                CU(cudaMemcpyAsync(h->pl[P_V0] + g, h->pl[P_V1] + g, gb, cudaMemcpyDeviceToDevice, h->st));
            }
            void *pp[2] = {h->pl[P_U0], h->pl[P_V0]};
            if (int rc = exchange_rows(h, pp, 2, sizeof(double))) return rc;
        } else {
            CU(cudaMemcpyAsync(h->pl[P_U0], h->pl[P_U1], bytes, cudaMemcpyDeviceToDevice, h->st));
            CU(cudaMemcpyAsync(h->pl[P_V0], h->pl[P_V1], bytes, cudaMemcpyDeviceToDevice, h->st));
        }
        for (int k = 0; k < EVP_NSTRESS; ++k)
            CU(cudaMemcpyAsync(h->pl[P_S0 + k], h->pl[P_S1 + k], bytes, cudaMemcpyDeviceToDevice, h->st));
        h->cur = 0;
    }
    return 0;
}

// tiled layout: bring the current state (tile copy h->cur) back into the planes U0, V0, S0
int sync_planes(evp_b200_handle *h) {
    if (!h->tiled || !h->planes_stale) return 0;
    TileStateArgs ta;
    ta.u = h->pl[P_U0];
    ta.v = h->pl[P_V0];
    for (int k = 0; k < EVP_NSTRESS; ++k) ta.s[k] = h->pl[P_S0 + k];
    aux_tile_unpack_state(h->pg, h->tg, h->cur, ta, h->st);
    CU(cudaGetLastError());
    h->planes_stale = false;
    return 0;
}

// tiled layout: planes U0, V0, S0 and the loop-invariant fields of this call -> tiles (state copy 0 current)
int pack_tiles(evp_b200_handle *h) {
    double **p = h->pl;
    if (!h->tiles_static) {
        TileStaticArgs sa = {p[P_DXT], p[P_DYT], p[P_DXHY], p[P_DYHX], p[P_CXP], p[P_CYP], p[P_CXM], p[P_CYM],
                             p[P_TINYAREA], p[P_UAREAR]};
        aux_tile_pack_static(h->pg, h->tg, sa, h->st);
        h->tiles_static = true;
    }
    TileCallArgs ca;
    ca.strength = p[P_STRENGTH]; ca.aiu = p[P_AIU]; ca.uocn = p[P_UOCN]; ca.vocn = p[P_VOCN];
    ca.waterx = p[P_WATERX]; ca.watery = p[P_WATERY]; ca.forcex = p[P_FORCEX]; ca.forcey = p[P_FORCEY];
    ca.umassdtei = p[P_UMASSDTEI]; ca.fm = p[P_FM];
    ca.icetmask = h->mk[M_ICETMASK]; ca.iceumask = h->mk[M_ICEUMASK];
    ca.u = p[P_U0]; ca.v = p[P_V0];
    for (int k = 0; k < EVP_NSTRESS; ++k) ca.s[k] = p[P_S0 + k];
    aux_tile_pack_call(h->pg, h->tg, ca, h->st);
    CU(cudaGetLastError());
    h->cur = 0;
    h->planes_stale = false;
    return 0;
}

// Whether the ndte loop runs as ONE cooperative launch (k_persist) instead of one launch per subcycle:
// opt-in through kernel_variant bit 7 (128), where it is possible.  Measured on B200 it is bit-identical
// but not faster: the neighbour waits cost ~3 us per subcycle against a ~2 us launch gap
// inside a CUDA graph, so the graph of k_subcycle launches stays the default.
void decide_persistent(evp_b200_handle *h) {
    h->persistent = false;
    if ((h->par.kernel_variant & 128) == 0) return;
    if (h->pg.ns_cyclic) return;                          // the north-south wrap runs between the kernels
    if (h->pg.tripole && !h->fold_in_kernel) return;      // so does the separate tripole fold
    if (h->dims.nranks > 1 && !h->p2p) return;            // and the NCCL row exchange
    if (h->par.ndte < 2) return;
    int sms = 148, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    SubArgs a;
    fill_subargs(h, a, 0);
    persist_launch_fn fn = h->par.math_mode == 1 ? evp_persist_launch_fast : evp_persist_launch_strict;
    if (fn(a, h->threads, 0, 0, nullptr, &per_sm) != 0) { cudaGetLastError(); return; }
    if ((long)h->grid_x * h->grid_y > (long)per_sm * sms) return; // every CTA must be resident
    h->persistent = true;
}

// Tiling of the subcycle kernel: balanced strips in x, one wave of CTAs in total, row chunks in
// launch order with short boundary chunks.  Returns 0 or a CUDA error code via fail().
int choose_tiling(evp_b200_handle *h) {
    h->warpx = false;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int nx = h->pg.nx, nyl = h->pg.nyl;
    int nt, ncx, strip_w, per_sm;
    bool tma = false;
    if (h->fused) {
        // two-subcycle kernel: one warp per strip of up to 29 U columns, 4 warps per CTA (at most 112 columns: every
        // warp can own at least 28, see the strip table below), two CTAs per SM (255 registers, 102 KB of rings)
        nt = 128;
        ncx = (nx + 111) / 112;
        strip_w = (nx + ncx - 1) / ncx;
        SubArgs a;
        fill_subargs(h, a, 0);
        fused_launch_fn fn = h->par.math_mode == 1 ? evp_fused_launch_fast : evp_fused_launch_strict;
        per_sm = 0;
        const int e = fn(a, false, 0, 0, 0, nullptr, &per_sm);
        if (e != 0 || per_sm < 1) {
            cudaGetLastError();
            return fail(EVP_B200_ERR_CUDA, "two-subcycle kernel cannot be configured: %s", cudaGetErrorString((cudaError_t)e));
        }
    } else if (h->tiled) {
        // strip-tiled TMA-fed kernel: one warp per strip of 31 U columns, 4 warps per CTA; resident CTAs per
        // SM from the occupancy of the kernel itself (shared memory: 2 or 3 pipeline stages per warp)
        // 3 stages / 2 CTAs per SM measured equal or better than 2 stages / 3 CTAs on every short slab but one
        // (360 x 300: 13.7 vs 13.3 us); kernel_variant bit 12 selects the latter
        h->tiled_stages = (h->par.kernel_variant & 4096) ? 2 : 3;
        nt = 128;
        strip_w = EVT_UW;
        ncx = (evt_nstrips(nx) + 3) / 4;
        SubArgs a;
        fill_subargs(h, a, 0);
        tiled_launch_fn fn = h->par.math_mode == 1 ? evp_tiled_launch_fast : evp_tiled_launch_strict;
        per_sm = 0;
        const int e = fn(a, false, h->tiled_stages, 0, 0, 0, nullptr, &per_sm);
        if (e != 0 || per_sm < 1) {
            cudaGetLastError();
            return fail(EVP_B200_ERR_CUDA, "tiled subcycle kernel cannot be configured: %s", cudaGetErrorString((cudaError_t)e));
        }
    } else {
        // default: 128 threads per CTA; short slabs (multi-GPU: 19.5 vs 23.0 us at 135 rows) run better with one
        // 256-thread CTA per SM, 300 rows and more with two 128-thread CTAs (1 degree: 12.7 vs 13.6 us)
        // (decided from the mean slab height so that all ranks of a chain use the same strips)
        const int nyl_mean = h->dims.ny_global / h->dims.nranks;
        nt = h->par.tile_threads > 0 ? h->par.tile_threads : (nyl_mean < 300 ? 256 : 128);
        if (nt != 64 && nt != 128 && nt != 256) nt = 128;
        // TMA-staged kernel (kernel_variant bits 8 / 9): 128 threads, even strips of at most nt - 2 columns
        // (16-byte aligned row segments), 2 or 3 CTAs per SM
        tma = (h->par.kernel_variant & (256 | 512)) != 0;
        if (tma) {
            nt = 128;
            const int e = h->par.math_mode == 1 ? evp_subcycle_configure_fast() : evp_subcycle_configure_strict();
            if (e != 0) return fail(EVP_B200_ERR_CUDA, "TMA-staged subcycle kernel: %s", cudaGetErrorString((cudaError_t)e));
        }
        // Warp-autonomous strips (k_subcycle<.., WARPX>): each of the 4 warps of a 128-thread CTA owns strip_w / 4 <= 31
        // U columns of its own, shuffles replace the exchange line and the per-row CTA barrier.  kernel_variant
        // bit 20 (1048576) selects it, bit 21 (2097152) forbids it; see h->warpx below for the default.
        const int kvv = h->par.kernel_variant;
        h->warpx = nt == 128 && !tma && (kvv & (128 | 2097152)) == 0 &&
                   ((kvv & 1048576) != 0 || EVP_WARPX_DEFAULT);
        // balanced strips: ncx strips of strip_w U columns, strip_w + 1 <= nt threads hold T columns
        const int wmax = h->warpx ? 124 : (tma ? nt - 2 : nt - 1);
        ncx = (nx + wmax - 1) / wmax;
        strip_w = (nx + ncx - 1) / ncx;
        if (tma && (strip_w & 1)) ++strip_w;
        if (h->warpx) strip_w = (strip_w + 3) & ~3;
        // resident CTAs per SM at ~210 registers per thread: 256 threads
        per_sm = (nt == 256) ? 1 : (nt == 128 ? 2 : 4);
        if (tma && (h->par.kernel_variant & 256) == 0) per_sm = 3;
        if (!tma && nt == 128 && (h->par.kernel_variant & 1024)) per_sm = 3; // late-load kernel: 3 CTAs per SM
    }
    int ncy = (per_sm * sms) / ncx; // one wave
    if (ncy < 1) ncy = 1;
    if (ncy > nyl) ncy = nyl;
    // the in-kernel fold is the u-fold; the T-fold (rows nyl, nyl+1 mirror rows nyl-1, nyl-2) runs as k_halo_tripole
    const bool fold_wanted = h->pg.tripole && !h->pg.tfold && (h->par.kernel_variant & 4) == 0;
    // relative length of the boundary chunks: the northernmost chunk of the top slab also folds, a
    // chunk next to another slab waits for that slab's flag at its start
    const bool multi = h->dims.nranks > 1 && h->par.exchange_mode == 0;
    double w_top = 1.0, w_bot = 1.0;
    if (ncy >= 3 && h->par.tile_rows <= 0) {
        // (two-subcycle kernel: the fold chunk runs two one-subcycle passes over its rows + 1 and waits for its
        // neighbours twice, against one fused pass over rows + 3 at ~1.7 x the cost per row)
        // A chunk next to another slab pays a fixed cost per kernel (flag wait, coherent ghost-row loads, remote
        // stores, system fence before its flag): ~1.5 rows' worth in the plane kernels, ~4.5 rows' worth in the tiled
        // kernel, whose rows are twice as fast (measured on 4 x B200, 135 rows per slab: 22.4 / 19.8 / 17.4 us per
        // subcycle with boundary chunks of 1.0 / 0.6 / 0.3 interior lengths)
        double w_peer = 0.6;
        if (h->tiled) {
            const double r0 = (double)nyl / ncy;
            w_peer = std::min(1.0, std::max(0.2, 1.0 - 4.5 / r0));
        }
        if (fold_wanted) w_top = h->fused ? 0.6 : 0.5;
        else if (multi && h->north >= 0) w_top = w_peer;
        if (multi && h->south >= 0) w_bot = w_peer;
    }
    if (const char *e = getenv("EVP_B200_WBND")) { // development: relative length of the chunks next to another slab
        const double w = atof(e);
        if (w > 0.05 && w <= 2.0 && multi) {
            if (h->north >= 0 && !fold_wanted) w_top = w;
            if (h->south >= 0) w_bot = w;
        }
    }
    std::vector<int> tab; // (j0, n) in launch order: bottom, top, then interior south to north
    int rows = h->par.tile_rows;
    if (rows > 0) { // explicit uniform chunks (tests, sweeps)
        if (rows > nyl) rows = nyl;
        ncy = (nyl + rows - 1) / rows;
        std::vector<int> j0s;
        for (int k = 0; k < ncy; ++k) j0s.push_back(1 + k * rows);
        auto len = [&](int k) { return std::min(rows, nyl - (j0s[k] - 1)); };
        tab.push_back(j0s[0]); tab.push_back(len(0));
        if (ncy > 1) { tab.push_back(j0s[ncy - 1]); tab.push_back(len(ncy - 1)); }
        for (int k = 1; k < ncy - 1; ++k) { tab.push_back(j0s[k]); tab.push_back(len(k)); }
    } else {
        if (ncy < 3) { w_top = w_bot = 1.0; }
        const double units = (ncy >= 2 ? (ncy - 2) + w_top + w_bot : 1.0);
        rows = (int)(nyl / units + 0.999);
        if (rows < 2) rows = 2;
        int n_bot = ncy >= 2 ? std::max(2, (int)(rows * w_bot + 0.5)) : nyl;
        int n_top = ncy >= 2 ? std::max(2, (int)(rows * w_top + 0.5)) : 0;
        if (n_bot + n_top > nyl) { n_bot = nyl; n_top = 0; }
        const int mid = nyl - n_bot - n_top;
        int n_mid = mid > 0 ? (mid + rows - 1) / rows : 0;
        tab.push_back(1); tab.push_back(n_bot);
        if (n_top > 0) { tab.push_back(nyl - n_top + 1); tab.push_back(n_top); }
        int j = 1 + n_bot;
        for (int k = 0; k < n_mid; ++k) { // spread the interior rows evenly
            const int n = mid / n_mid + (k < mid % n_mid ? 1 : 0);
            tab.push_back(j); tab.push_back(n);
            j += n;
        }
        ncy = (int)tab.size() / 2;
    }
    if (h->par.kernel_variant & 65536) {
        // MEASUREMENT ONLY (results invalid): every chunk marches the same rows in the middle of the slab, so the
        // whole grid works on an L2-resident set -- the kernel's time without DRAM
        const int n = std::min(rows, std::max(2, nyl - 4));
        for (int k = 0; k < ncy; ++k) { tab[2 * k] = std::max(2, nyl / 2 - n / 2); tab[2 * k + 1] = n; }
    }
    h->threads = nt;
    h->strip_w = strip_w;
    h->rows = rows;
    h->grid_x = ncx;
    h->grid_y = ncy;
    // in-kernel tripole fold needs rows nyl-1 and nyl in the northernmost chunk
    int top_rows = 0;
    for (int k = 0; k < ncy; ++k)
        if (tab[2 * k] + tab[2 * k + 1] - 1 == nyl) top_rows = tab[2 * k + 1];
    h->fold_in_kernel = fold_wanted && (top_rows >= 2 || (h->par.kernel_variant & 65536));
    if (h->fused) {
        // the fold chunk reads two rows below itself; every chunk of the table must be one wave (the fold chunk's
        // CTAs wait for each other)
        int top_j0 = 0;
        for (int k = 0; k < ncy; ++k)
            if (tab[2 * k] + tab[2 * k + 1] - 1 == nyl) top_j0 = tab[2 * k];
        const bool ok = (!h->pg.tripole || (h->fold_in_kernel && (top_j0 >= 3 || (h->par.kernel_variant & 65536)))) &&
                        (long)ncx * ncy <= (long)per_sm * sms;
        if (!ok) {
            h->fused = false;
            return choose_tiling(h);
        }
        // Strip table.  Per warp: virtual column of lane 0; owned lanes; lane G of the ghost T column nx+1 where the
        // warp holds the seam of a cyclic domain (lanes .., nx, nx+1, 1, 2, ..); lanes in use.  A warp owns lanes
        // 1..29 (a first cyclic warp 2..29 behind its two seam lanes), a warp that owns column nx of a cyclic domain
        // at most up to lane 28 (three seam lanes follow), so four warps always cover a CTA's <= 112 columns.
        std::vector<int> ws((size_t)ncx * 16, 0);
        const bool cyc = h->pg.ew_cyclic != 0;
        for (int x = 0; x < ncx; ++x) {
            const int ca = 1 + strip_w * x, cb = std::min(strip_w * (x + 1), nx);
            int c = ca;
            for (int w = 0; w < 4; ++w) {
                int *e = &ws[((size_t)x * 4 + w) * 4];
                if (c > cb) { e[0] = 0; e[1] = 1 | (0 << 8); e[2] = -1; e[3] = 1 | (0 << 8); continue; }
                const bool first_cyc = cyc && c == 1;
                const int l0 = first_cyc ? 2 : 1;
                int n = std::min(cb - c + 1, 29 - l0 + 1);
                if (cyc && c + n - 1 == nx && l0 + n - 1 > 28) --n; // column nx moves to the next warp
                const int lo = l0, hi = l0 + n - 1;
                const bool owns_nx = c + n - 1 == nx;
                int G = -1, v0 = c - l0, llo = lo - 1, lhi = hi + 2;
                if (first_cyc) { G = 1; v0 = nx; llo = 0; }
                else if (cyc && owns_nx) { G = hi + 1; lhi = hi + 3; }
                e[0] = v0; e[1] = lo | (hi << 8); e[2] = G; e[3] = llo | (lhi << 8);
                c += n;
            }
            if (c <= cb) { // cannot happen for strips of <= 112 columns
                h->fused = false;
                return choose_tiling(h);
            }
        }
        if (h->d_wstrips) cudaFree(h->d_wstrips);
        CU(cudaMalloc(&h->d_wstrips, sizeof(int) * ws.size()));
        CU(cudaMemcpy(h->d_wstrips, ws.data(), sizeof(int) * ws.size(), cudaMemcpyHostToDevice));
    }
    if (h->tiled && h->pg.tripole && !h->fold_in_kernel) {
        // the separate fold kernel works on planes: tiny slabs / variant bit 2 fall back to the plane kernels
        h->tiled = false;
        return choose_tiling(h);
    }
    h->w_bot = (float)w_bot;
    h->w_top = (float)w_top;
    // with the default tiling the chunk table is re-balanced by active cells on the device each call
    h->balance = h->par.tile_rows <= 0 && ncy >= 3 && (h->par.kernel_variant & (32 | 65536)) == 0;
    if (h->d_chunks) cudaFree(h->d_chunks);
    if (h->d_rowcnt) cudaFree(h->d_rowcnt);
    if (h->d_cta_epoch) cudaFree(h->d_cta_epoch);
    CU(cudaMalloc(&h->d_rowcnt, sizeof(int) * (nyl + 2)));
    CU(cudaMalloc(&h->d_chunks, sizeof(int) * tab.size()));
    CU(cudaMalloc(&h->d_cta_epoch, sizeof(int) * (size_t)ncx * ncy));
    CU(cudaMemcpy(h->d_chunks, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice));
    return 0;
}

} // namespace

extern "C" {

int evp_b200_abi_version(void) { return EVP_B200_ABI_VERSION; }

const char *evp_b200_last_error(void) { return g_err.c_str(); }

void evp_b200_default_params(evp_b200_params *p) {
    memset(p, 0, sizeof(*p));
    p->dt = 3600.0;
    p->ndte = 120;          // source/ice_init.F90:216
    p->evp_damping = 0;     // :217
    p->dragio = 0.00536;    // drivers/cice4/ice_constants.F90:59
    p->cosw = 1.0;          // source/ice_dyn_evp.F90:84-85
    p->sinw = 0.0;
    p->rhoi = 917.0;
    p->rhos = 330.0;
    p->rhow = 1026.0;
    p->gravit = 9.80616;
    p->puny = 1.0e-11;
    p->kstrength = 1;       // source/ice_init.F90:219-222
    p->krdg_partic = 1;
    p->krdg_redist = 1;
    p->mu_rdg = 3.0;
    p->ncat = 5;
    p->math_mode = 0;
    p->pin_host = 0;
    p->use_graph = 1;
}

// everything of evp_b200_init that can fail after the handle exists; the caller finalizes on error
static int init_handle(evp_b200_handle *h, const evp_b200_dims *d, const evp_b200_params *p,
                       const evp_b200_static_fields *g) {
    h->dims = *d;
    h->par = *p;
    if (d->device >= 0) {
        CU(cudaSetDevice(d->device));
        h->device = d->device;
    } else {
        CU(cudaGetDevice(&h->device));
    }
    // block table + coverage check
    const int nyl = d->slab_jhi - d->slab_jlo + 1;
    h->blk_tab.resize((size_t)d->nblocks * 6);
    long covered = 0;
    // cells of the slab not covered by any block are the land blocks the reference's distribution has
    // eliminated (source/ice_distribution.F90: blocks without ocean points get no task): they stay land
    // with zero fields in the planes, which is what the halo update's zero fill gives their neighbours'
    // ghost cells in the reference (mpi|serial/ice_boundary.F90, "fill out halo region")
    std::vector<uint8_t> cover((size_t)d->nx_global * nyl, 0);
    for (int b = 0; b < d->nblocks; ++b) {
        if (d->ilo[b] < 2 || d->jlo[b] < 2 || d->ihi[b] > d->nx_block - 1 || d->jhi[b] > d->ny_block - 1 ||
            d->ihi[b] < d->ilo[b] || d->jhi[b] < d->jlo[b]) {
            return fail(EVP_B200_ERR_ARG, "block %d: bad ilo/ihi/jlo/jhi (nghost must be 1)", b);
        }
        h->blk_tab[b * 6 + 0] = d->ilo[b];
        h->blk_tab[b * 6 + 1] = d->ihi[b];
        h->blk_tab[b * 6 + 2] = d->jlo[b];
        h->blk_tab[b * 6 + 3] = d->jhi[b];
        h->blk_tab[b * 6 + 4] = d->iglob_lo[b] - d->ilo[b];
        h->blk_tab[b * 6 + 5] = d->jglob_lo[b] - d->jlo[b] - (d->slab_jlo - 1);
        const int ig1 = d->iglob_lo[b] + (d->ihi[b] - d->ilo[b]);
        const int jg1 = d->jglob_lo[b] + (d->jhi[b] - d->jlo[b]);
        if (d->iglob_lo[b] < 1 || ig1 > d->nx_global || d->jglob_lo[b] < d->slab_jlo || jg1 > d->slab_jhi) {
            return fail(EVP_B200_ERR_ARG, "block %d lies outside the slab", b);
        }
        for (int jg = d->jglob_lo[b]; jg <= jg1; ++jg)
            for (int ig = d->iglob_lo[b]; ig <= ig1; ++ig) {
                uint8_t &c = cover[(size_t)(jg - d->slab_jlo) * d->nx_global + (ig - 1)];
                if (c) return fail(EVP_B200_ERR_ARG, "block %d overlaps another block at global cell (%d, %d)", b, ig, jg);
                c = 1;
            }
        covered += (long)(d->ihi[b] - d->ilo[b] + 1) * (d->jhi[b] - d->jlo[b] + 1);
    }
    h->eliminated_cells = (long)d->nx_global * nyl - covered;

    // set_evp_parameters, source/ice_dyn_evp.F90:563-575
    {
        const double eyc = 0.36;
        const double dte = p->dt / (double)p->ndte;
        h->dtei = 1.0 / dte;
        const double ecc = 4.0;
        h->ecci = 0.25;
        const double tdamp2 = 2.0 * eyc * p->dt;
        h->dte2T = dte / tdamp2;
        h->denom1 = 1.0 / (1.0 + h->dte2T);
        h->denom2 = 1.0 / (1.0 + h->dte2T * ecc);
        h->rcon = 1230.0 * eyc * p->dt * (h->dtei * h->dtei);
        h->dragw = p->dragio * p->rhow; // :78 / :1382
    }

    PlaneGeom &pg = h->pg;
    pg.nx = d->nx_global;
    pg.nyl = nyl;
    pg.pitch = ((pg.nx + 2 + 15) / 16) * 16;
    pg.ew_cyclic = d->ew_boundary == EVP_B200_BND_CYCLIC;
    pg.ns_cyclic = d->ns_boundary == EVP_B200_BND_CYCLIC;
    pg.tfold = d->ns_boundary == EVP_B200_BND_TRIPOLET;
    pg.tripole = (d->ns_boundary == EVP_B200_BND_TRIPOLE || pg.tfold) && d->rank == d->nranks - 1;
    h->north = (d->rank < d->nranks - 1) ? d->rank + 1 : -1;
    h->south = (d->rank > 0) ? d->rank - 1 : -1;
    pg.cells = (size_t)pg.pitch * (pg.nyl + 2);
    // Plane spacing: consecutive planes are streamed concurrently by the subcycle kernel (~40 of
    // them); a spacing that is a multiple of 4 KiB puts all streams on the same L2 slice / HBM
    // channel phase.  Skew the spacing by an odd number of 256-byte segments.
    {
        long skew = 1 * 32 + 0; // doubles
        if (const char *e = getenv("EVP_B200_PLANE_SKEW")) skew = atol(e);
        if (skew < 0) skew = 0;
        if (skew > 4096) skew = 4096;
        pg.cells += (size_t)skew;
    }
    // the subcycle kernel addresses a plane and its second copy with 32-bit element offsets: the largest
    // one is copy_stride (P_U1 - P_U0 planes) plus an index inside the last plane
    if ((unsigned long long)(P_U2 - P_U0 + 1) * pg.cells + (unsigned long long)pg.pitch >= (1ull << 31))
        return fail(EVP_B200_ERR_ARG, "slab too large for 32-bit plane offsets (%zu cells per plane): use more ranks",
                    pg.cells);
    h->blocked_elems = (size_t)d->nx_block * d->ny_block * d->max_blocks;

    CU(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->st_up, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_early_done, cudaEventDisableTiming));
    for (auto &e : h->ev) CU(cudaEventCreate(&e));
    // + one row: the TMA-staged kernel copies whole 16-byte aligned row segments and may read past the
    // last column of a row (into the next row; past the pool only for the last row of the last plane)
    CU(cudaMalloc(&h->pool, sizeof(double) * (pg.cells * P_COUNT + pg.pitch)));
    CU(cudaMemsetAsync(h->pool, 0, sizeof(double) * pg.cells * P_COUNT, h->st));
    for (int k = 0; k < P_COUNT; ++k) h->pl[k] = h->pool + (size_t)k * pg.cells;
    CU(cudaMalloc(&h->mpool, pg.cells * M_COUNT));
    CU(cudaMemsetAsync(h->mpool, 0, pg.cells * M_COUNT, h->st));
    for (int k = 0; k < M_COUNT; ++k) h->mk[k] = h->mpool + (size_t)k * pg.cells;
    h->n_stage = SL_COUNT;
    CU(cudaMalloc(&h->stage, sizeof(double) * h->blocked_elems * h->n_stage));
    CU(cudaMalloc(&h->stage_i, sizeof(int32_t) * h->blocked_elems * 2));
    // padding cells of padded blocks / unused blocks are never written by the pack kernels: keep them 0
    CU(cudaMemsetAsync(h->stage, 0, sizeof(double) * h->blocked_elems * h->n_stage, h->st));
    CU(cudaMemsetAsync(h->stage_i, 0, sizeof(int32_t) * h->blocked_elems * 2, h->st));
    CU(cudaMalloc(&h->fold_scratch, sizeof(double) * 2 * pg.pitch));
    CU(cudaMalloc(&h->sync, sizeof(int) * EVP_SYNC_INTS + 4 * sizeof(double))); // + diagnostics scratch
    CU(cudaMemsetAsync(h->sync, 0, sizeof(int) * EVP_SYNC_INTS, h->st));
    CU(cudaMalloc(&h->d_blk_tab, sizeof(int) * h->blk_tab.size()));
    CU(cudaMemcpyAsync(h->d_blk_tab, h->blk_tab.data(), sizeof(int) * h->blk_tab.size(), cudaMemcpyHostToDevice, h->st));
    h->bg.nx_block = d->nx_block;
    h->bg.ny_block = d->ny_block;
    h->bg.nblocks = d->nblocks;
    h->bg.tab = h->d_blk_tab;

    // static fields
    struct { const double *src; int id; } statics[] = {
        {g->dxt, P_DXT}, {g->dyt, P_DYT}, {g->dxhy, P_DXHY}, {g->dyhx, P_DYHX}, {g->cxp, P_CXP}, {g->cyp, P_CYP},
        {g->cxm, P_CXM}, {g->cym, P_CYM}, {g->tarea, P_TAREA}, {g->tarear, P_TAREAR}, {g->tinyarea, P_TINYAREA},
        {g->uarea, P_UAREA}, {g->uarear, P_UAREAR}, {g->fcor, P_FCOR}};
    int slot = 0;
    for (auto &s : statics) {
        int rc = upload_r8(h, s.src, slot++, h->pl[s.id]);
        if (rc) return rc;
    }
    int rc = upload_mask(h, g->tmask, 0, h->mk[M_TMASK]);
    if (!rc) rc = upload_mask(h, g->umask, 1, h->mk[M_UMASK]);
    if (!rc && g->HTE && g->HTN) {
        rc = upload_r8(h, g->HTE, slot++, h->pl[P_HTE]);
        if (!rc) rc = upload_r8(h, g->HTN, slot++, h->pl[P_HTN]);
        if (!rc) {
            CU(cudaMalloc(&h->row_ht, pg.nyl + 2));
            MetricCheckArgs mc;
            mc.hte = h->pl[P_HTE]; mc.htn = h->pl[P_HTN];
            mc.dxt = h->pl[P_DXT]; mc.dyt = h->pl[P_DYT]; mc.dxhy = h->pl[P_DXHY]; mc.dyhx = h->pl[P_DYHX];
            mc.cxp = h->pl[P_CXP]; mc.cyp = h->pl[P_CYP]; mc.cxm = h->pl[P_CXM]; mc.cym = h->pl[P_CYM];
            mc.tmask = h->mk[M_TMASK]; mc.row_ht = h->row_ht;
            aux_check_metrics(pg, mc, h->st);
            std::vector<uint8_t> flags(pg.nyl + 2);
            CU(cudaMemcpyAsync(flags.data(), h->row_ht, pg.nyl + 2, cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            for (uint8_t f : flags) h->rows_ht += f ? 1 : 0;
        }
    }
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->st));
    // Which kernel runs the ndte loop.  The strip-tiled TMA-fed kernel needs everything of the per-subcycle halo
    // update inside the kernel: no north-south cyclic wrap, no NCCL exchange in the loop, per-strip flags for all
    // strips.  Measured on B200 (profiles/r02_*): it wins on short slabs -- 1440 x 135: 15.0 vs 19.2 us per
    // subcycle, 1440 x 270: 28.9 vs 32.4, 3600 x 338: 95.7 vs 103.8, 100 x 116: 6.3 vs 6.8 -- where the plane
    // kernel's few rows per CTA cost it redundant T rows and exposed row latency, and loses on tall ones (1440 x
    // 540: 61.6 vs 57.0, 1440 x 1080: 123.6 vs 112.6), where the duplicated slot of every strip (+3 %), the
    // shorter row chunks and the halo words make it stream 9 % more bytes.  Default: tiled below 450 rows per
    // slab (decided from the mean slab height so that all ranks of a chain agree); kernel_variant bit 11 forces
    // it, bit 15 forces the plane kernels.
    const int kv = p->kernel_variant;
    const bool tiled_ok = (kv & (16 | 128 | 256 | 512 | 1024 | 32768)) == 0 && !pg.ns_cyclic &&
                          !(d->nranks > 1 && p->exchange_mode != 0) && evt_nstrips(pg.nx) <= EVP_SYNC_MAXCX;
    const bool tiled_auto = p->tile_threads == 0 && d->ny_global / d->nranks < 450;
    h->tiled = tiled_ok && ((kv & 2048) != 0 || tiled_auto);
    // Two subcycles per launch (evp_fused.cuh): single rank, no north-south wrap, u-fold or no fold, plane layout.
    // kernel_variant bit 17 (131072) selects it.
    const bool fused_ok = (kv & (4 | 16 | 128 | 256 | 512 | 1024 | 2048)) == 0 && d->nranks == 1 && !pg.ns_cyclic &&
                          !pg.tfold && p->ndte >= 2 && pg.nx >= 64 && pg.nyl >= 16 && p->tile_threads == 0;
    h->fused = fused_ok && (kv & 131072) != 0;
    if (h->fused) h->tiled = false;
    if (h->tiled) {
        h->tg.ns = evt_nstrips(pg.nx);
        h->tg.nr = pg.nyl + 2;
        evt_set_order(h->tg, (p->kernel_variant & 8192) ? 0 : 1); // bit 13: strip-major instead of row-major
        const size_t tb = sizeof(double) * (size_t)h->tg.ns * h->tg.nr * EVT_ROW_D;
        CU(cudaMalloc(&h->tg.tiles, tb));
        CU(cudaMemsetAsync(h->tg.tiles, 0, tb, h->st));
    }
    if (int trc = choose_tiling(h)) return trc;
    if (!h->tiled && h->tg.tiles) {
        cudaFree(h->tg.tiles);
        h->tg.tiles = nullptr;
    }
    decide_persistent(h); // multi-rank: decided again once the peer-to-peer halo is set up (comm_init)
    memset(&h->tm, 0, sizeof(h->tm));
    h->pin_enabled = true;
    return EVP_B200_OK;
}

int evp_b200_init(const evp_b200_dims *d, const evp_b200_params *p, const evp_b200_static_fields *g,
                  evp_b200_handle **out) {
    if (!d || !p || !g || !out) return fail(EVP_B200_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (d->nblocks < 1 || d->nblocks > d->max_blocks) return fail(EVP_B200_ERR_ARG, "bad nblocks");
    if (d->nx_block < 3 || d->ny_block < 3) return fail(EVP_B200_ERR_ARG, "bad block size");
    if (!d->ilo || !d->ihi || !d->jlo || !d->jhi || !d->iglob_lo || !d->jglob_lo)
        return fail(EVP_B200_ERR_ARG, "block index arrays are NULL");
    if (d->ns_boundary > EVP_B200_BND_TRIPOLET || d->ew_boundary > EVP_B200_BND_CYCLIC || d->ew_boundary < 0 ||
        d->ns_boundary < 0)
        return fail(EVP_B200_ERR_UNSUPPORTED, "boundary type not supported");
    if (d->ns_boundary == EVP_B200_BND_TRIPOLET && d->rank == d->nranks - 1 && d->slab_jhi - d->slab_jlo + 1 < 3)
        return fail(EVP_B200_ERR_ARG, "the T-fold needs at least 3 rows in the top slab");
    if (d->nranks < 1 || d->rank < 0 || d->rank >= d->nranks) return fail(EVP_B200_ERR_ARG, "bad rank/nranks");
    if (d->slab_jlo < 1 || d->slab_jhi > d->ny_global || d->slab_jhi - d->slab_jlo + 1 < 2)
        return fail(EVP_B200_ERR_ARG, "bad slab rows (each slab needs at least 2 rows)");
    if (d->nranks == 1 && (d->slab_jlo != 1 || d->slab_jhi != d->ny_global))
        return fail(EVP_B200_ERR_ARG, "slab must span the domain when nranks == 1");
    if (d->nranks > 1 && d->ns_boundary == EVP_B200_BND_CYCLIC)
        return fail(EVP_B200_ERR_UNSUPPORTED, "north-south cyclic domains are single-slab only");
    if (d->nranks > 1 && ((d->rank == 0) != (d->slab_jlo == 1) || (d->rank == d->nranks - 1) != (d->slab_jhi == d->ny_global)))
        return fail(EVP_B200_ERR_ARG, "slabs must be ordered south to north by rank");
    if (p->ndte < 1 || !(p->dt > 0.0)) return fail(EVP_B200_ERR_ARG, "bad dt/ndte");
    if (p->ncat < 1 || p->ncat > 16) return fail(EVP_B200_ERR_ARG, "ncat out of range");
    if ((d->ns_boundary == EVP_B200_BND_TRIPOLE || d->ns_boundary == EVP_B200_BND_TRIPOLET) && d->nx_global + 2 > 8 * 1024)
        return fail(EVP_B200_ERR_UNSUPPORTED, "tripole fold kernel supports nx_global <= 8190");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(EVP_B200_ERR_CUDA, "no CUDA device: libevp_b200 has no CPU fallback");
    }
    evp_b200_handle *h = new evp_b200_handle();
    const int rc = init_handle(h, d, p, g);
    if (rc) {
        const std::string msg = g_err; // finalize must not lose the reason
        evp_b200_finalize(h);
        g_err = msg;
        return rc;
    }
    *out = h;
    return EVP_B200_OK;
}

// On every error return the streams are drained before control goes back to the caller: asynchronous
// downloads into the caller's arrays must not still be in flight when the caller frees or reuses them.
static int drain_on_error(evp_b200_handle *h, int rc) {
    if (rc && h) {
        const std::string why = g_err;
        if (h->st) cudaStreamSynchronize(h->st);
        if (h->st2) cudaStreamSynchronize(h->st2);
        if (h->st_up) cudaStreamSynchronize(h->st_up);
        h->dn_pending = false;
        cudaGetLastError();
        g_err = why;
    }
    return rc;
}

static int do_prep_impl(evp_b200_handle *h, const evp_b200_inputs *in, evp_b200_state *st, int32_t *icetmask_out) {
    if (!h || !in || !st) return fail(EVP_B200_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(h->device));
    const PlaneGeom &pg = h->pg;
    double **p = h->pl;
    const size_t pbytes = pg.cells * sizeof(double);
    if (int src = sync_planes(h)) return src; // resident stresses must be current in the planes
    CU(cudaEventRecord(h->ev[0], h->st));
    // ---- upload inputs and state --------------------------------------------------------------
    int rc = 0;
    if ((rc = upload_r8(h, in->aice, SL_AICE, p[P_AICE]))) return rc;
    if ((rc = upload_r8(h, in->vice, SL_VICE, p[P_VICE]))) return rc;
    if ((rc = upload_r8(h, in->vsno, SL_VSNO, p[P_VSNO]))) return rc;
    // strairx = strairxT (:668-669) or strax (:273-274): uploaded straight into work1 of t2ugrid_vector
    if ((rc = upload_r8(h, in->strairxT, SL_STRAIRX, p[P_WRKX]))) return rc;
    if ((rc = upload_r8(h, in->strairyT, SL_STRAIRY, p[P_WRKY]))) return rc;
    if ((rc = upload_r8(h, in->uocn, SL_UOCN, p[P_UOCN]))) return rc;
    if ((rc = upload_r8(h, in->vocn, SL_VOCN, p[P_VOCN]))) return rc;
    const bool need_slope = h->par.coupled_tilt && !(h->par.hemisphere_turning && !h->par.use_ocnslope);
    if (need_slope) {
        if ((rc = upload_r8(h, in->ss_tltx, SL_SSTLTX, p[P_SSTLTX]))) return rc;
        if ((rc = upload_r8(h, in->ss_tlty, SL_SSTLTY, p[P_SSTLTY]))) return rc;
    }
    // state_residency: 0 = the whole state travels both ways every call; 1 = the stresses stay on the device;
    // 2 = uvel, vvel and iceumask too (the planes hold the result of the previous call)
    const bool keep_vel = h->par.state_residency == 2 && h->vel_on_device;
    const bool keep_stress = h->par.state_residency >= 1 && h->stress_on_device;
    if (!keep_vel) {
        if ((rc = upload_r8(h, st->uvel, SL_U, p[P_U0]))) return rc;
        if ((rc = upload_r8(h, st->vvel, SL_V, p[P_V0]))) return rc;
    }
    double *sh[EVP_NSTRESS] = {st->stressp_1, st->stressp_2, st->stressp_3, st->stressp_4,
                               st->stressm_1, st->stressm_2, st->stressm_3, st->stressm_4,
                               st->stress12_1, st->stress12_2, st->stress12_3, st->stress12_4};
    if (!keep_stress)
        for (int k = 0; k < EVP_NSTRESS; ++k)
            if ((rc = upload_r8(h, sh[k], SL_S0 + k, p[P_S0 + k]))) return rc;
    if (!keep_vel && (rc = upload_mask(h, st->iceumask, 0, h->mk[M_ICEUMASK]))) return rc;
    CU(cudaEventRecord(h->ev[1], h->st));

    // ---- :214-224 and init_history_dyn (source/ice_flux.F90:585-602) ---------------------------
    const int zero_ids[] = {P_RDG_CONV, P_RDG_SHEAR, P_DIVU, P_SHEAR, P_PRS_SIG, P_STRTLTX, P_STRTLTY,
                            P_STRINTX, P_STRINTY, P_STROCNX, P_STROCNY, P_FM};
    for (int id : zero_ids) CU(cudaMemsetAsync(p[id], 0, pbytes, h->st));

    PrepArgs a;
    memset(&a, 0, sizeof(a));
    a.tarea = p[P_TAREA]; a.uarea = p[P_UAREA]; a.fcor = p[P_FCOR];
    a.tmask = h->mk[M_TMASK]; a.umask = h->mk[M_UMASK];
    a.aice = p[P_AICE]; a.vice = p[P_VICE]; a.vsno = p[P_VSNO]; a.uocn = p[P_UOCN]; a.vocn = p[P_VOCN];
    a.ss_tltx = p[P_SSTLTX]; a.ss_tlty = p[P_SSTLTY];
    a.tmass = p[P_TMASS]; a.umass = p[P_UMASS]; a.aiu = p[P_AIU]; a.umassdtei = p[P_UMASSDTEI];
    a.waterx = p[P_WATERX]; a.watery = p[P_WATERY]; a.forcex = p[P_FORCEX]; a.forcey = p[P_FORCEY];
    a.strairx = p[P_STRAIRX]; a.strairy = p[P_STRAIRY]; a.strtltx = p[P_STRTLTX]; a.strtlty = p[P_STRTLTY];
    a.strintx = p[P_STRINTX]; a.strinty = p[P_STRINTY]; a.strocnx = p[P_STROCNX]; a.strocny = p[P_STROCNY];
    a.fm = p[P_FM]; a.uvel = p[P_U0]; a.vvel = p[P_V0];
    for (int k = 0; k < EVP_NSTRESS; ++k) a.stress[k] = p[P_S0 + k];
    a.tmphm = h->mk[M_TMPHM]; a.icetmask = h->mk[M_ICETMASK]; a.iceumask = h->mk[M_ICEUMASK];
    a.rhoi = h->par.rhoi; a.rhos = h->par.rhos; a.dtei = h->dtei; a.cosw = h->par.cosw; a.sinw = h->par.sinw;
    a.gravit = h->par.gravit;
    a.hemisphere_turning = h->par.hemisphere_turning; a.coupled_tilt = h->par.coupled_tilt;
    a.use_ocnslope = h->par.use_ocnslope;

    aux_prep1(pg, a, h->st);                                             // :236-242
    aux_icetmask(pg, a, h->st);
    aux_halo_u8(pg, h->mk[M_ICETMASK], h->st);                           // :250-253
    {
        void *pp[1] = {h->mk[M_ICETMASK]};
        if ((rc = exchange_rows(h, pp, 1, 1))) return rc;
    }
    aux_to_ugrid(pg, p[P_TMASS], p[P_TAREA], p[P_UAREA], p[P_UMASS], h->st); // :259-260
    aux_to_ugrid(pg, p[P_AICE], p[P_TAREA], p[P_UAREA], p[P_AIU], h->st);
    if ((rc = halo_r8(h, p[P_WRKX], 1, -1))) return rc;                  // t2ugrid_vector, :276-277
    if ((rc = halo_r8(h, p[P_WRKY], 1, -1))) return rc;
    aux_to_ugrid(pg, p[P_WRKX], p[P_TAREA], p[P_UAREA], p[P_STRAIRX], h->st);
    aux_to_ugrid(pg, p[P_WRKY], p[P_TAREA], p[P_UAREA], p[P_STRAIRY], h->st);
    aux_prep2(pg, a, h->st);                                             // :292-316
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev[2], h->st));
    h->tm.kernel_launches = 0;
    if (icetmask_out) {
        if ((rc = download_mask(h, icetmask_out, 1, h->mk[M_ICETMASK], PACK_FULL))) return rc;
        if ((rc = join_downloads(h))) return rc;
        CU(cudaStreamSynchronize(h->st));
    }
    h->prepared = true;
    h->resident = false;
    h->cur = 0;
    return 0;
}

static int do_prep(evp_b200_handle *h, const evp_b200_inputs *in, evp_b200_state *st, int32_t *icetmask_out) {
    return drain_on_error(h, do_prep_impl(h, in, st, icetmask_out));
}

static int do_run_impl(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength, evp_b200_state *st,
                       evp_b200_outputs *out) {
    if (!h || !st) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->prepared) return fail(EVP_B200_ERR_STATE, "evp_b200_run called before evp_b200_prep");
    CU(cudaSetDevice(h->device));
    const PlaneGeom &pg = h->pg;
    double **p = h->pl;
    const size_t pbytes = pg.cells * sizeof(double);
    int rc = 0;
    // ---- ice strength (:322-332) ----------------------------------------------------------------
    if (strength) {
        if ((rc = upload_r8(h, strength, SL_STRENGTH, p[P_STRENGTH]))) return rc;
    } else {
        if (!in || !in->aice0 || !in->aicen || !in->vicen)
            return fail(EVP_B200_ERR_ARG, "strength == NULL needs aice0/aicen/vicen for the device ice_strength");
        const int ncat = h->par.ncat;
        if (!h->cat) {
            CU(cudaMalloc(&h->cat, pbytes * 2 * ncat));
            CU(cudaMemsetAsync(h->cat, 0, pbytes * 2 * ncat, h->st));
            CU(cudaMalloc(&h->stage_cat, sizeof(double) * h->blocked_elems));
        }
        if ((rc = upload_r8(h, in->aice0, SL_AICE0, p[P_AICE0]))) return rc;
        // (nx_block, ny_block, ncat, max_blocks): category n of block b is a strided set of planes;
        // with max_blocks == 1 it is contiguous.  General case: copy per (block, category).
        const size_t be = (size_t)h->dims.nx_block * h->dims.ny_block;
        for (int which = 0; which < 2; ++which) {
            const double *src = which == 0 ? in->aicen : in->vicen;
            if (!h->io_device) pin(h, src, be * ncat * h->dims.max_blocks * sizeof(double));
            for (int n = 0; n < ncat; ++n) {
                for (int b = 0; b < h->dims.nblocks; ++b)
                    CU(cudaMemcpyAsync(h->stage_cat + (size_t)b * be, src + ((size_t)b * ncat + n) * be,
                                       be * sizeof(double), cudaMemcpyDefault, h->st));
                aux_unblock_r8(h->bg, pg, h->stage_cat, h->cat + (size_t)(which * ncat + n) * pg.cells, h->st);
            }
        }
        StrengthArgs sa;
        sa.aice = p[P_AICE]; sa.vice = p[P_VICE]; sa.aice0 = p[P_AICE0];
        sa.aicen = h->cat; sa.vicen = h->cat + (size_t)ncat * pg.cells;
        sa.icetmask = h->mk[M_ICETMASK]; sa.strength = p[P_STRENGTH];
        sa.ncat = ncat; sa.kstrength = h->par.kstrength; sa.krdg_partic = h->par.krdg_partic;
        sa.krdg_redist = h->par.krdg_redist; sa.mu_rdg = h->par.mu_rdg; sa.puny = h->par.puny;
        sa.gravit = h->par.gravit; sa.rhow = h->par.rhow; sa.rhoi = h->par.rhoi;
        aux_ice_strength(pg, sa, h->st);
    }
    // ---- :336-344 ------------------------------------------------------------------------------
    if ((rc = halo_r8(h, p[P_STRENGTH], 1, 1))) return rc;
    if ((rc = halo_r8(h, p[P_U0], 2, -1))) return rc;
    if ((rc = halo_r8(h, p[P_V0], 2, -1))) return rc;
    if (h->tiled) {
        // the loop runs on the strip-tiled layout: state and loop-invariant fields -> tiles (both copies)
        if ((rc = pack_tiles(h))) return rc;
    } else {
        // second ping-pong copy: same ghost / masked-out values as copy 0; stresses outside the T list are 0
        CU(cudaMemcpyAsync(p[P_U1], p[P_U0], pbytes, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(p[P_V1], p[P_V0], pbytes, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemsetAsync(p[P_S1], 0, pbytes * EVP_NSTRESS, h->st));
        if (h->fused && h->pg.tripole) { // third copy: the rows of the fold chunk (ghost / masked-out values)
            CU(cudaMemcpyAsync(p[P_U2], p[P_U0], pbytes, cudaMemcpyDeviceToDevice, h->st));
            CU(cudaMemcpyAsync(p[P_V2], p[P_V0], pbytes, cudaMemcpyDeviceToDevice, h->st));
            CU(cudaMemsetAsync(p[P_S2], 0, pbytes * EVP_NSTRESS, h->st));
        }
    }
    if (h->p2p) {
        // the neighbours store into the ghost rows of copy 1 from their first subcycle kernel on:
        // exchanging those rows once more (same values; with the tiled layout only as a two-sided ordering
        // point) orders their stores after the copy / the packing above
        void *pp[2] = {p[P_U1], p[P_V1]};
        if ((rc = exchange_rows(h, pp, 2, sizeof(double)))) return rc;
    }
    h->cur = 0;
    if (h->balance) // chunks of equal ACTIVE work; an inactive row inside a chunk still costs ~30 % of an active one
        aux_balance_chunks(pg, h->mk[M_ICETMASK], h->mk[M_ICEUMASK], h->d_rowcnt, h->d_chunks, h->grid_y,
                           h->w_bot, h->w_top, h->fold_in_kernel ? 2 : 1, 0.3f * (float)(pg.nx + 1),
                           h->south >= 0 ? 1 : 0, (h->fold_in_kernel || h->north >= 0) ? 1 : 0, h->st);
    CU(cudaEventRecord(h->ev[3], h->st));
    // Outputs that are final before the subcycle loop travel to the host on a second stream while
    // the loop runs (copy engine and SMs overlap): strairx/y, strtltx/y, fm, strength, sicemass.
    struct EarlyOut { double *dst; int id; int policy; int slot; };
    const EarlyOut early[] = {
        {out ? out->strairx : nullptr, P_STRAIRX, PACK_INT_ZERO, SL_OUT0 + 0},
        {out ? out->strairy : nullptr, P_STRAIRY, PACK_INT_ZERO, SL_OUT0 + 1},
        {out ? out->strtltx : nullptr, P_STRTLTX, PACK_INT_ZERO, SL_OUT0 + 2},
        {out ? out->strtlty : nullptr, P_STRTLTY, PACK_INT_ZERO, SL_OUT0 + 3},
        {out ? out->fm : nullptr, P_FM, PACK_INT_ZERO, SL_OUT0 + 10},
        {out ? out->strength : nullptr, P_STRENGTH, PACK_FULL, SL_OUT0 + 16},
        {out ? out->sicemass : nullptr, P_TMASS, PACK_FULL, SL_OUT0 + 17}};
    CU(cudaEventRecord(h->ev_early, h->st));
    CU(cudaStreamWaitEvent(h->st2, h->ev_early, 0));
    for (const EarlyOut &e : early)
        if ((rc = download_r8(h, e.dst, e.slot, p[e.id], e.policy, h->st2))) return rc;
    CU(cudaEventRecord(h->ev_early_done, h->st2));
    // ---- :347-404 ------------------------------------------------------------------------------
    const bool fin_fused = fuse_finish(h);
    if (fin_fused) { // evp_finish inside the last subcycle kernel: strocnxT / strocnyT = 0 everywhere first (:1512-1513)
        CU(cudaMemsetAsync(p[P_WRKX], 0, pbytes, h->st));
        CU(cudaMemsetAsync(p[P_WRKY], 0, pbytes, h->st));
    }
    if ((rc = run_subcycle_loop(h))) return rc;
    if ((rc = sync_planes(h))) return rc; // tiled layout: the result goes back into the planes
    CU(cudaEventRecord(h->ev[4], h->st));
    h->resident = true;
    // ---- evp_finish + u2tgrid_vector (:410-428) -------------------------------------------------
    FinishArgs fa;
    fa.uvel = p[P_U0]; fa.vvel = p[P_V0]; fa.uocn = p[P_UOCN]; fa.vocn = p[P_VOCN]; fa.aiu = p[P_AIU];
    fa.fm = p[P_FM]; fa.iceumask = h->mk[M_ICEUMASK];
    fa.strocnx = p[P_STROCNX]; fa.strocny = p[P_STROCNY]; fa.strocnxT = p[P_WRKX]; fa.strocnyT = p[P_WRKY];
    fa.dragw = h->dragw; fa.cosw = h->par.cosw; fa.sinw = h->par.sinw;
    fa.hemisphere_turning = h->par.hemisphere_turning;
    if (!fin_fused) aux_finish(pg, fa, h->st);
    // work1 = strocnxT; HALO(work1); to_tgrid(work1, strocnxT): the ghosts of strocnxT keep the 0 of :1512
    CU(cudaMemsetAsync(p[P_STROCNXT], 0, pbytes, h->st));
    CU(cudaMemsetAsync(p[P_STROCNYT], 0, pbytes, h->st));
    if ((rc = halo_r8(h, p[P_WRKX], 2, -1))) return rc;
    if ((rc = halo_r8(h, p[P_WRKY], 2, -1))) return rc;
    aux_to_tgrid(pg, p[P_WRKX], p[P_TAREA], p[P_UAREA], p[P_STROCNXT], h->st);
    aux_to_tgrid(pg, p[P_WRKY], p[P_TAREA], p[P_UAREA], p[P_STROCNYT], h->st);
    if (out && (out->sig1 || out->sig2))
        aux_principal_stress(pg.cells, p[P_S0], p[P_S0 + 4], p[P_S0 + 8], p[P_PRS_SIG], h->par.puny,
                             p[P_SIG1], p[P_SIG2], h->st);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev[5], h->st));
    // ---- download ------------------------------------------------------------------------------
    if (h->par.state_residency != 2) {
        if ((rc = download_r8(h, st->uvel, SL_U, p[P_U0], PACK_FULL))) return rc;
        if ((rc = download_r8(h, st->vvel, SL_V, p[P_V0], PACK_FULL))) return rc;
    }
    double *sh[EVP_NSTRESS] = {st->stressp_1, st->stressp_2, st->stressp_3, st->stressp_4,
                               st->stressm_1, st->stressm_2, st->stressm_3, st->stressm_4,
                               st->stress12_1, st->stress12_2, st->stress12_3, st->stress12_4};
    if (h->par.state_residency == 0)
        for (int k = 0; k < EVP_NSTRESS; ++k)
            if ((rc = download_r8(h, sh[k], SL_S0 + k, p[P_S0 + k], PACK_TNE_KEEP))) return rc;
    h->stress_on_device = true;
    h->vel_on_device = true;
    if (h->par.state_residency != 2 && (rc = download_mask(h, st->iceumask, 0, h->mk[M_ICEUMASK], PACK_INT_KEEP))) return rc;
    if (out) {
        struct { double *dst; int id; int policy; } outs[] = {
            {out->strairx, P_STRAIRX, PACK_INT_ZERO}, {out->strairy, P_STRAIRY, PACK_INT_ZERO},
            {out->strtltx, P_STRTLTX, PACK_INT_ZERO}, {out->strtlty, P_STRTLTY, PACK_INT_ZERO},
            {out->strintx, P_STRINTX, PACK_INT_ZERO}, {out->strinty, P_STRINTY, PACK_INT_ZERO},
            {out->strocnx, P_STROCNX, PACK_INT_ZERO}, {out->strocny, P_STROCNY, PACK_INT_ZERO},
            {out->strocnxT, P_STROCNXT, PACK_INT_ZERO}, {out->strocnyT, P_STROCNYT, PACK_INT_ZERO},
            {out->fm, P_FM, PACK_INT_ZERO}, {out->prs_sig, P_PRS_SIG, PACK_TNE_ZERO},
            {out->divu, P_DIVU, PACK_TNE_ZERO}, {out->shear, P_SHEAR, PACK_TNE_ZERO},
            {out->rdg_conv, P_RDG_CONV, PACK_TNE_ZERO}, {out->rdg_shear, P_RDG_SHEAR, PACK_TNE_ZERO},
            {out->strength, P_STRENGTH, PACK_FULL}, {out->sicemass, P_TMASS, PACK_FULL},
            {out->sig1, P_SIG1, PACK_FULL}, {out->sig2, P_SIG2, PACK_FULL}};
        int slot = SL_OUT0;
        for (auto &o : outs) {
            const bool was_early = o.id == P_STRAIRX || o.id == P_STRAIRY || o.id == P_STRTLTX || o.id == P_STRTLTY ||
                                   o.id == P_FM || o.id == P_STRENGTH || o.id == P_TMASS;
            if (!was_early && (rc = download_r8(h, o.dst, slot, p[o.id], o.policy))) return rc;
            ++slot;
        }
    }
    if ((rc = join_downloads(h))) return rc; // early and late outputs, state: all copies on st2
    CU(cudaEventRecord(h->ev[6], h->st));
    if ((rc = check_wait_flag(h))) return rc;
    CU(cudaEventElapsedTime(&h->tm.upload_ms, h->ev[0], h->ev[1]));
    CU(cudaEventElapsedTime(&h->tm.prep_ms, h->ev[1], h->ev[3]));
    CU(cudaEventElapsedTime(&h->tm.subcycle_ms, h->ev[3], h->ev[4]));
    CU(cudaEventElapsedTime(&h->tm.finish_ms, h->ev[4], h->ev[5]));
    CU(cudaEventElapsedTime(&h->tm.download_ms, h->ev[5], h->ev[6]));
    CU(cudaEventElapsedTime(&h->tm.total_ms, h->ev[0], h->ev[6]));
    h->tm.subcycle_launches = h->sub_launches_per_loop;
    h->tm.exchange_mode_used = h->dims.nranks == 1 ? -1 : (h->p2p ? 0 : 1);
    h->tm.reserved = h->rows_ht; // rows on the 2-plane metric path
    h->prepared = false;
    return 0;
}

static int do_run(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength, evp_b200_state *st,
                  evp_b200_outputs *out) {
    return drain_on_error(h, do_run_impl(h, in, strength, st, out));
}

int evp_b200_prep(evp_b200_handle *h, const evp_b200_inputs *in, evp_b200_state *st, int32_t *icetmask_out) {
    return do_prep(h, in, st, icetmask_out);
}

int evp_b200_run(evp_b200_handle *h, const double *strength, evp_b200_state *st, evp_b200_outputs *out) {
    if (!strength) return fail(EVP_B200_ERR_ARG, "evp_b200_run needs the strength array (use evp_b200_step for the device ice_strength)");
    return do_run(h, nullptr, strength, st, out);
}

int evp_b200_step(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength, evp_b200_state *st,
                  evp_b200_outputs *out) {
    int rc = do_prep(h, in, st, nullptr);
    if (rc) return rc;
    return do_run(h, in, strength, st, out);
}

int evp_b200_subcycle_resident(evp_b200_handle *h, int32_t repeats, float *ms_per_loop) {
    if (!h || repeats < 1) return fail(EVP_B200_ERR_ARG, "bad argument");
    if (!h->resident) return fail(EVP_B200_ERR_STATE, "no device-resident state: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[3], h->st));
    for (int r = 0; r < repeats; ++r) {
        int rc = run_subcycle_loop(h);
        if (rc) return rc;
    }
    CU(cudaEventRecord(h->ev[4], h->st));
    if (int rc = check_wait_flag(h)) return rc; // a timing of an invalid loop is not reported
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]));
    if (ms_per_loop) *ms_per_loop = ms / (float)repeats;
    return 0;
}

int evp_b200_principal_stress(evp_b200_handle *h, const double *sp1, const double *sm1, const double *s12,
                              const double *prs, double *sig1, double *sig2) {
    if (!h || !sp1 || !sm1 || !s12 || !prs || !sig1 || !sig2) return fail(EVP_B200_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(h->device));
    // pointwise over the whole block array (source/ice_dyn_evp.F90:1593-1607): no layout change needed
    const size_t n = h->blocked_elems, bytes = n * sizeof(double);
    double *s = h->stage;
    const double *src[4] = {sp1, sm1, s12, prs};
    for (int k = 0; k < 4; ++k) CU(cudaMemcpyAsync(s + k * n, src[k], bytes, cudaMemcpyHostToDevice, h->st));
    aux_principal_stress(n, s, s + n, s + 2 * n, s + 3 * n, h->par.puny, s + 4 * n, s + 5 * n, h->st);
    CU(cudaMemcpyAsync(sig1, s + 4 * n, bytes, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(sig2, s + 5 * n, bytes, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int evp_b200_principal_stress_n(evp_b200_handle *h, int64_t n, const double *sp1, const double *sm1, const double *s12,
                                const double *prs, double *sig1, double *sig2) {
    if (!h || !sp1 || !sm1 || !s12 || !prs || !sig1 || !sig2) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (n < 1 || (size_t)n > h->blocked_elems) return fail(EVP_B200_ERR_ARG, "n must be 1 .. nx_block*ny_block*max_blocks");
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)n * sizeof(double), m = h->blocked_elems;
    double *s = h->stage;
    const double *src[4] = {sp1, sm1, s12, prs};
    for (int k = 0; k < 4; ++k) CU(cudaMemcpyAsync(s + k * m, src[k], bytes, cudaMemcpyHostToDevice, h->st));
    aux_principal_stress((size_t)n, s, s + m, s + 2 * m, s + 3 * m, h->par.puny, s + 4 * m, s + 5 * m, h->st);
    CU(cudaMemcpyAsync(sig1, s + 4 * m, bytes, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(sig2, s + 5 * m, bytes, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int evp_b200_diagnostics(evp_b200_handle *h, double out[4]) {
    if (!h || !out) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->resident) return fail(EVP_B200_ERR_STATE, "no device-resident result: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    if (int src = sync_planes(h)) return src;
    double *d = (double *)(h->sync + EVP_SYNC_INTS); // 4 doubles behind the sync block
    CU(cudaMemsetAsync(d, 0, 4 * sizeof(double), h->st));
    // lmask_s: ULAT < -puny  <=>  fcor = 2*omega*sin(ULAT) < 2*omega*sin(-puny)
    const double fcor_south = 2.0 * 7.292e-5 * sin(-h->par.puny);
    aux_diagnostics(h->pg, h->pl[P_U0], h->pl[P_V0], h->pl[P_STRENGTH], h->pl[P_FCOR], fcor_south, d, h->st);
    CU(cudaMemcpyAsync(out, d, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int evp_b200_diagnostics_energy(evp_b200_handle *h, double out[8]) {
    if (!h || !out) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->resident) return fail(EVP_B200_ERR_STATE, "no device-resident result: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    if (int src = sync_planes(h)) return src;
    if (!h->d_energy) CU(cudaMalloc(&h->d_energy, sizeof(double) * (6 * (size_t)(h->pg.nyl + 2) + 8)));
    EnergyArgs ea;
    ea.u = h->pl[P_U0]; ea.v = h->pl[P_V0]; ea.vice = h->pl[P_VICE]; ea.vsno = h->pl[P_VSNO];
    ea.tarea = h->pl[P_TAREA]; ea.fcor = h->pl[P_FCOR]; ea.tmask = h->mk[M_TMASK];
    ea.rhoi = h->par.rhoi; ea.rhos = h->par.rhos;
    ea.fcor_south = 2.0 * 7.292e-5 * sin(-h->par.puny);
    ea.rowsum = h->d_energy;
    ea.out6 = h->d_energy + 6 * (size_t)(h->pg.nyl + 2);
    aux_energy_sums(h->pg, ea, h->st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ea.out6, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    // rms ice speed, source/ice_diagnostics.F90:221-234 (for one slab; several slabs: sum out[0..5], then this)
    for (int k = 0; k < 2; ++k) {
        double urms = 2.0 * out[k] / (h->par.rhoi * out[2 + k] + h->par.rhos * out[4 + k] + h->par.puny);
        out[6 + k] = urms > h->par.puny ? sqrt(urms) : 0.0;
    }
    return 0;
}

int evp_b200_download_state(evp_b200_handle *h, evp_b200_state *st) {
    if (!h || !st) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->stress_on_device) return fail(EVP_B200_ERR_STATE, "no device state yet: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    double **p = h->pl;
    int rc = 0;
    if ((rc = sync_planes(h))) return rc;
    if ((rc = download_r8(h, st->uvel, SL_U, p[P_U0], PACK_FULL))) return rc;
    if ((rc = download_r8(h, st->vvel, SL_V, p[P_V0], PACK_FULL))) return rc;
    double *sh[EVP_NSTRESS] = {st->stressp_1, st->stressp_2, st->stressp_3, st->stressp_4,
                               st->stressm_1, st->stressm_2, st->stressm_3, st->stressm_4,
                               st->stress12_1, st->stress12_2, st->stress12_3, st->stress12_4};
    for (int k = 0; k < EVP_NSTRESS; ++k)
        if ((rc = download_r8(h, sh[k], SL_S0 + k, p[P_S0 + k], PACK_TNE_KEEP))) return rc;
    if ((rc = download_mask(h, st->iceumask, 0, h->mk[M_ICEUMASK], PACK_INT_KEEP))) return rc;
    if ((rc = join_downloads(h))) return rc;
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int evp_b200_invalidate_device_state(evp_b200_handle *h) {
    if (!h) return fail(EVP_B200_ERR_ARG, "NULL argument");
    h->stress_on_device = false;
    h->vel_on_device = false;
    return 0;
}

int evp_b200_download_velocity(evp_b200_handle *h, double *uvel, double *vvel) {
    if (!h || !uvel || !vvel) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->vel_on_device) return fail(EVP_B200_ERR_STATE, "no device state yet: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    int rc = 0;
    if ((rc = sync_planes(h))) return rc;
    if ((rc = download_r8(h, uvel, SL_U, h->pl[P_U0], PACK_FULL))) return drain_on_error(h, rc);
    if ((rc = download_r8(h, vvel, SL_V, h->pl[P_V0], PACK_FULL))) return drain_on_error(h, rc);
    if ((rc = join_downloads(h))) return rc;
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int evp_b200_device_velocity(evp_b200_handle *h, const double **uvel, const double **vvel, int32_t *pitch,
                             int32_t *nrows) {
    if (!h || !uvel || !vvel || !pitch || !nrows) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (!h->vel_on_device) return fail(EVP_B200_ERR_STATE, "no device state yet: call evp_b200_step/run first");
    CU(cudaSetDevice(h->device));
    if (int rc = sync_planes(h)) return rc;
    CU(cudaStreamSynchronize(h->st));
    *uvel = h->pl[P_U0];
    *vvel = h->pl[P_V0];
    *pitch = h->pg.pitch;
    *nrows = h->pg.nyl + 2;
    return 0;
}

int evp_b200_step_device(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength, evp_b200_state *st,
                         evp_b200_outputs *out) {
    if (!h) return fail(EVP_B200_ERR_ARG, "NULL argument");
    h->io_device = true;
    int rc = do_prep(h, in, st, nullptr);
    if (!rc) rc = do_run(h, in, strength, st, out);
    h->io_device = false;
    return rc;
}

int evp_b200_get_info(const evp_b200_handle *h, int32_t out[8]) {
    if (!h || !out) return fail(EVP_B200_ERR_ARG, "NULL argument");
    out[0] = h->tiled ? 1 : (h->fused ? 2 : 0); // kernel of the ndte loop: 0 plane, 1 strip-tiled, 2 two subcycles per launch
    out[1] = h->grid_x;
    out[2] = h->grid_y;
    out[3] = h->threads;
    out[4] = h->strip_w;
    out[5] = h->tiled ? h->tiled_stages : 0;
    out[6] = h->p2p ? 1 : 0;
    out[7] = (h->persistent ? 1 : 0) | (h->warpx ? 2 : 0) | (fuse_finish(h) ? 4 : 0);
    return 0;
}

int evp_b200_selftest_ieee(int64_t n, uint64_t seed, uint64_t out[6]) {
    if (n < 1 || !out) return fail(EVP_B200_ERR_ARG, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(EVP_B200_ERR_CUDA, "no CUDA device: libevp_b200 has no CPU fallback");
    }
    unsigned long long o[6] = {0, 0, 0, 0, 0, 0};
    const int e = aux_selftest_ieee((long long)n, (unsigned long long)seed, o);
    if (e != 0) return fail(EVP_B200_ERR_CUDA, "selftest kernel: %s", cudaGetErrorString((cudaError_t)e));
    for (int k = 0; k < 6; ++k) out[k] = o[k];
    return 0;
}

int evp_b200_unpin(evp_b200_handle *h, const void *host_ptr) {
    if (!h) return fail(EVP_B200_ERR_ARG, "NULL argument");
    auto it = h->pinned.find(host_ptr);
    if (it == h->pinned.end()) return 0;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->st));
    CU(cudaStreamSynchronize(h->st2));
    CU(cudaStreamSynchronize(h->st_up));
    cudaHostUnregister(const_cast<void *>(host_ptr));
    cudaGetLastError();
    h->pinned.erase(it);
    return 0;
}

int evp_b200_get_timings(const evp_b200_handle *h, evp_b200_timings *t) {
    if (!h || !t) return fail(EVP_B200_ERR_ARG, "NULL argument");
    *t = h->tm;
    return 0;
}

static void *open_nccl() {
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return lib;
}

int evp_b200_comm_unique_id(uint8_t id[128]) {
    if (!id) return fail(EVP_B200_ERR_ARG, "NULL argument");
    void *lib = open_nccl();
    if (!lib) return fail(EVP_B200_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    auto fn = (ncclResult_t(*)(ncclUniqueId *))dlsym(lib, "ncclGetUniqueId");
    if (!fn) return fail(EVP_B200_ERR_COMM, "ncclGetUniqueId not found");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    ncclResult_t r = fn(&uid);
    if (r != ncclSuccess) return fail(EVP_B200_ERR_COMM, "ncclGetUniqueId failed (%d)", (int)r);
    memcpy(id, &uid, 128);
    return 0;
}

int evp_b200_comm_init(evp_b200_handle *h, const uint8_t id[128]) {
    if (!h || !id) return fail(EVP_B200_ERR_ARG, "NULL argument");
    if (h->dims.nranks == 1) return 0;
    if (h->comm) return 0;
    CU(cudaSetDevice(h->device));
    h->nccl_lib = open_nccl();
    if (!h->nccl_lib) return fail(EVP_B200_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    auto initRank = (ncclResult_t(*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h->nccl_lib, "ncclCommInitRank");
    h->pGroupStart = (ncclResult_t(*)())dlsym(h->nccl_lib, "ncclGroupStart");
    h->pGroupEnd = (ncclResult_t(*)())dlsym(h->nccl_lib, "ncclGroupEnd");
    h->pSend = (decltype(h->pSend))dlsym(h->nccl_lib, "ncclSend");
    h->pRecv = (decltype(h->pRecv))dlsym(h->nccl_lib, "ncclRecv");
    h->pCommDestroy = (decltype(h->pCommDestroy))dlsym(h->nccl_lib, "ncclCommDestroy");
    h->pGetErrorString = (decltype(h->pGetErrorString))dlsym(h->nccl_lib, "ncclGetErrorString");
    if (!initRank || !h->pGroupStart || !h->pGroupEnd || !h->pSend || !h->pRecv || !h->pCommDestroy || !h->pGetErrorString)
        return fail(EVP_B200_ERR_COMM, "libnccl lacks a required entry point");
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclResult_t r = initRank(&h->comm, h->dims.nranks, uid, h->dims.rank);
    if (r != ncclSuccess) {
        h->comm = nullptr;
        return fail(EVP_B200_ERR_COMM, "ncclCommInitRank failed: %s", h->pGetErrorString(r));
    }
    // the epochs of the peer-to-peer halo count from here on every rank (calls made before the
    // communicator existed did not advance sync[1]); the neighbours cannot write into this block before
    // they have received its handle below, which is ordered after this memset on the stream
    h->epoch_count = 0;
    CU(cudaMemsetAsync(h->sync, 0, sizeof(int) * EVP_SYNC_INTS, h->st));
    if (h->par.exchange_mode == 0) {
        // Peer-to-peer halo: swap CUDA IPC handles of the plane pool and the sync block with both
        // neighbours (through the communicator just made), map them, and let the subcycle kernel
        // store its boundary rows straight into the neighbours' ghost rows.
        struct PeerInfo {
            cudaIpcMemHandle_t pool, sync, tiles;
            int nyl, pitch, tiled;
            unsigned long long cells;
        } mine, theirs[2];
        memset(&mine, 0, sizeof(mine));
        memset(theirs, 0, sizeof(theirs));
        CU(cudaIpcGetMemHandle(&mine.pool, h->pool));
        CU(cudaIpcGetMemHandle(&mine.sync, h->sync));
        if (h->tiled) CU(cudaIpcGetMemHandle(&mine.tiles, h->tg.tiles));
        mine.tiled = h->tiled ? 1 : 0;
        mine.nyl = h->pg.nyl;
        mine.pitch = h->pg.pitch;
        mine.cells = h->pg.cells;
        auto allReduce = (ncclResult_t(*)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                                          cudaStream_t))dlsym(h->nccl_lib, "ncclAllReduce");
        if (!allReduce) return fail(EVP_B200_ERR_COMM, "libnccl lacks ncclAllReduce");
        char *d = nullptr;
        CU(cudaMalloc(&d, 3 * sizeof(PeerInfo) + sizeof(int)));
        struct Free { char *p; ~Free() { cudaFree(p); } } free_d{d}; // also on the error returns below
        CU(cudaMemcpyAsync(d, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, h->st));
        ncclResult_t q = h->pGroupStart();
        const int nb[2] = {h->north, h->south};
        for (int k = 0; k < 2 && q == ncclSuccess; ++k)
            if (nb[k] >= 0) {
                q = h->pSend(d, sizeof(PeerInfo), ncclChar, nb[k], h->comm, h->st);
                if (q == ncclSuccess) q = h->pRecv(d + (k + 1) * sizeof(PeerInfo), sizeof(PeerInfo), ncclChar, nb[k], h->comm, h->st);
            }
        ncclResult_t q2 = h->pGroupEnd();
        if (q == ncclSuccess) q = q2;
        if (q != ncclSuccess) return fail(EVP_B200_ERR_COMM, "IPC handle exchange failed: %s", h->pGetErrorString(q));
        CU(cudaMemcpyAsync(theirs, d + sizeof(PeerInfo), 2 * sizeof(PeerInfo), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        bool ok = true;
        for (int k = 0; k < 2 && ok; ++k)
            if (nb[k] >= 0) {
                if (theirs[k].pitch != h->pg.pitch) ok = false;
                void *pp = nullptr, *ps = nullptr;
                if (ok && cudaIpcOpenMemHandle(&pp, theirs[k].pool, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
                if (ok && cudaIpcOpenMemHandle(&ps, theirs[k].sync, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
                if (ok && theirs[k].tiled != (h->tiled ? 1 : 0)) ok = false;
                if (ok && h->tiled) {
                    void *pt = nullptr;
                    if (cudaIpcOpenMemHandle(&pt, theirs[k].tiles, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
                    h->peer_tiles[k] = (double *)pt;
                }
                h->peer_pool[k] = (double *)pp;
                h->peer_sync[k] = (int *)ps;
                h->peer_nyl[k] = theirs[k].nyl;
                h->peer_cells[k] = (size_t)theirs[k].cells;
            }
        cudaGetLastError();
        if ((h->tiled ? h->tg.ns : h->grid_x) > EVP_SYNC_MAXCX) ok = false;
        // every rank must use the same exchange inside the loop: peer-to-peer only if ALL ranks can
        int mine_ok = ok ? 1 : 0, all_ok = 0;
        int *d_ok = (int *)(d + 3 * sizeof(PeerInfo));
        CU(cudaMemcpyAsync(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, h->st));
        q = allReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->st);
        if (q != ncclSuccess) return fail(EVP_B200_ERR_COMM, "ncclAllReduce failed: %s", h->pGetErrorString(q));
        CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        h->p2p = all_ok != 0; // otherwise the NCCL exchange stays in use (evp_b200_get_timings reports the mode)
        if (!h->p2p && h->tiled) { // the tiled kernel has no NCCL exchange: back to the plane kernels
            h->tiled = false;
            if (int trc = choose_tiling(h)) return trc;
        }
        decide_persistent(h);
    }
    return 0;
}

int evp_b200_finalize(evp_b200_handle *h) {
    if (!h) return EVP_B200_OK;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (auto &kv : h->pinned) cudaHostUnregister(const_cast<void *>(kv.first));
    for (int k = 0; k < 2; ++k) {
        if (h->peer_pool[k]) cudaIpcCloseMemHandle(h->peer_pool[k]);
        if (h->peer_sync[k]) cudaIpcCloseMemHandle(h->peer_sync[k]);
    }
    cudaFree(h->sync);
    cudaFree(h->d_energy);
    cudaFree(h->fold_scratch);
    cudaFree(h->d_chunks);
    cudaFree(h->d_wstrips);
    cudaFree(h->d_rowcnt);
    cudaFree(h->d_cta_epoch);
    cudaFree(h->row_ht);
    if (h->comm && h->pCommDestroy) h->pCommDestroy(h->comm);
    for (int k = 0; k < 2; ++k) {
        if (h->graph_exec[k]) cudaGraphExecDestroy(h->graph_exec[k]);
        if (h->graph[k]) cudaGraphDestroy(h->graph[k]);
        if (h->peer_tiles[k]) cudaIpcCloseMemHandle(h->peer_tiles[k]);
    }
    cudaFree(h->tg.tiles);
    cudaFree(h->pool);
    cudaFree(h->cat);
    cudaFree(h->mpool);
    cudaFree(h->stage);
    cudaFree(h->stage_i);
    cudaFree(h->stage_cat);
    cudaFree(h->d_blk_tab);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->ev_early) cudaEventDestroy(h->ev_early);
    if (h->ev_early_done) cudaEventDestroy(h->ev_early_done);
    if (h->st2) cudaStreamDestroy(h->st2);
    if (h->st_up) cudaStreamDestroy(h->st_up);
    for (auto &e : h->ring)
        if (e) cudaEventDestroy(e);
    if (h->st) cudaStreamDestroy(h->st);
    cudaGetLastError();
    delete h;
    return EVP_B200_OK;
}

} // extern "C"

/*
 * evp_b200.h -- C ABI of libevp_b200.so: the B200 (sm_100a) implementation of
 * CICE4's elastic-viscous-plastic dynamics step.
 *
 * This is the drop-in boundary for ONE path of COSIMA/cice4: `subroutine evp(dt)`
 * (source/ice_dyn_evp.F90:119-432) with `init_evp` (:441-526) and
 * `principal_stress` (:1558-1609).  The reference has no FFI of its own; a
 * replacement ice_dyn_evp.F90 (cice4_b200/fortran/ice_dyn_evp_b200.F90, see
 * INTEGRATION.md) keeps the module's public names and marshals `c_loc` pointers
 * of the module arrays of ice_state / ice_flux / ice_grid into these calls.
 *
 * Conventions
 *  - every array pointer is CALLER-OWNED HOST memory, fp64 unless stated,
 *    Fortran order (nx_block, ny_block, max_blocks) contiguous, valid for the
 *    duration of the call only; the library owns all device memory;
 *  - Fortran `logical` arrays (tmask, umask, iceumask) are passed as int32 0/1;
 *  - ghost cells of inputs are inputs (evp_prep1 reads aice/vice/vsno on the whole
 *    block, source/ice_dyn_evp.F90:643-677); the library performs only the halo
 *    updates evp itself performs (:250-253,:276-277,:336-344,:397-402,:427-428);
 *  - every function returns 0 on success, an EVP_B200_ERR_* code otherwise;
 *    evp_b200_last_error() returns a thread-local message.  No CPU fallback
 *    exists: without a CUDA device every compute entry fails with
 *    EVP_B200_ERR_CUDA.
 *  - one handle is driven by one host thread (the reference has no threads).
 */
#ifndef EVP_B200_H
#define EVP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVP_B200_ABI_VERSION 1

enum {
    EVP_B200_OK = 0,
    EVP_B200_ERR_ARG = 1,         /* bad argument / inconsistent dims */
    EVP_B200_ERR_CUDA = 2,        /* CUDA runtime error or no device */
    EVP_B200_ERR_UNSUPPORTED = 3, /* valid in the reference, not implemented here */
    EVP_B200_ERR_STATE = 4,       /* call order violated (run before prep, ...) */
    EVP_B200_ERR_COMM = 5         /* inter-GPU exchange failed */
};

/* ew_boundary_type / ns_boundary_type of domain_nml (source/ice_domain.F90:126-131,
 * index maps source/ice_blocks.F90:237-343).  'tripole' is the u-fold, 'tripoleT' the T-fold
 * (serial/ice_boundary.F90:725-773: three buffer rows, U rows jhi and jhi+1 mirror rows jhi-1 and
 * jhi-2, the degenerate top T row is symmetrised); the T-fold needs at least 3 rows in the top slab and
 * is folded by a separate kernel after every subcycle (the in-kernel fold and the tiled kernel are
 * u-fold only). */
enum { EVP_B200_BND_OPEN = 0, EVP_B200_BND_CLOSED = 1, EVP_B200_BND_CYCLIC = 2, EVP_B200_BND_TRIPOLE = 3,
       EVP_B200_BND_TRIPOLET = 4 };

/* Replaces: nx_block/ny_block/max_blocks (source/ice_blocks.F90:56-61,
 * source/ice_domain_size.F90:34-64), nblocks + blocks_ice (source/ice_domain.F90),
 * type block (source/ice_blocks.F90:32-45) and the y-slab re-mapping of
 * ice_distribution for GPUs. */
typedef struct {
    int32_t nx_block, ny_block; /* block size incl. the 1-cell ghost ring */
    int32_t max_blocks;         /* 3rd extent of every host array */
    int32_t nblocks;            /* local blocks actually used (<= max_blocks) */
    int32_t nx_global, ny_global;
    int32_t ew_boundary, ns_boundary;
    /* per local block, length nblocks, Fortran 1-based: this_block%ilo..jhi and the
     * global index of the first physical cell, this_block%i_glob(ilo), %j_glob(jlo) */
    const int32_t *ilo, *ihi, *jlo, *jhi;
    const int32_t *iglob_lo, *jglob_lo;
    /* rows of the global domain owned by this handle (1-based, inclusive); the local blocks
     * lie inside [1..nx_global] x [slab_jlo..slab_jhi] and do not overlap; cells that no
     * block covers are land (the reference's land-block elimination, ice_distribution.F90) */
    int32_t slab_jlo, slab_jhi;
    int32_t rank, nranks;       /* position in the south->north chain of y-slabs */
    int32_t device;             /* CUDA device ordinal, -1 = current */
} evp_b200_dims;

/* Replaces the scalars of module ice_dyn_evp (source/ice_dyn_evp.F90:64-103), the
 * constants it uses (drivers/cice4/ice_constants.F90:50-60,65-66) and the CPP
 * variants AusCOM / coupled / ACCESS as run-time flags (bld/Macros.nci:56-82). */
typedef struct {
    double dt;              /* dynamics time step passed to init_evp (s) */
    int32_t ndte;           /* subcycles, namelist ice_nml */
    int32_t evp_damping;    /* namelist ice_nml */
    double dragio;          /* ice-ocean drag coefficient */
    double cosw, sinw;      /* ocean turning angle */
    double rhoi, rhos, rhow, gravit, puny;
    int32_t coupled_tilt;       /* #ifdef coupled: tilt = -gravit*umass*ss_tlt (:924-925) */
    int32_t use_ocnslope;       /* AusCOM namelist; 0 reverts to geostrophic tilt (:928-933) */
    int32_t hemisphere_turning; /* #ifdef AusCOM/ACCICE: sign(1.,real(fm)) turning (:912-913,:1403-1408,:1525-1536) */
    int32_t wind_from_strax;    /* #ifdef ACCESS: strairx = strax (:271-275) */
    /* ice_strength options (source/ice_init.F90:219-222); only used when the library
     * computes strength itself (strength == NULL in evp_b200_step) */
    int32_t kstrength, krdg_partic, krdg_redist, ncat;
    double mu_rdg;
    /* library knobs */
    int32_t math_mode;      /* 0 = unfused IEEE (bit-exact vs the unfused CPU oracle); 1 = FMA-contracted */
    int32_t pin_host;       /* 1 = cudaHostRegister caller arrays on first use (keyed by address).  The caller
                               promises that a pinned array stays allocated until evp_b200_finalize, or tells the
                               library with evp_b200_unpin before freeing it (the Fortran module arrays live for
                               the whole run; the Python mirror unpins its per-call temporaries) */
    int32_t use_graph;      /* 1 = replay the ndte loop as one CUDA graph */
    int32_t tile_threads;   /* 0 = default; threads per CTA of the subcycle kernel */
    int32_t tile_rows;      /* 0 = default; U rows marched per CTA */
    int32_t kernel_variant; /* 0 = default; bit 2 (4): tripole fold as a separate kernel; bit 4 (16): force the 2-plane
                               metric path (needs HTE/HTN; the default on plane kernels of slabs >= 450 rows; bit 19
                               (524288) switches it off); bit 5 (32): uniform row chunks instead of the
                               per-call active-work balance; bit 6 (64): programmatic dependent launch;
                               bit 7 (128): run the ndte loop as one persistent cooperative launch whose CTAs
                               synchronise with their neighbours only (bit-identical; measured slower than
                               the default graph of per-subcycle launches); bit 8 (256) / bit 9
                               (512): T-row planes staged through shared memory by TMA bulk copies, 3 rows
                               deep with 2 CTAs per SM / 2 rows deep with 3 CTAs per SM; bit 10 (1024): no
                               register prefetch across the arithmetic (<= 168 registers, 3 CTAs per SM);
                               bit 11 (2048): force the strip-tiled layout + warp-autonomous TMA-fed kernel
                               (csrc/evp_tiled.cuh), which is the default below 450 rows per slab; it falls
                               back to the plane kernels where it does not apply (north-south cyclic domains,
                               exchange_mode 1, slabs too small for the in-kernel fold, tile_threads set);
                               bit 12 (4096): tiled kernel with 2 pipeline stages per warp and 3 CTAs per SM
                               instead of 3 stages and 2 CTAs; bit 13 (8192): strip-major instead of row-major
                               tile order; bit 14 (16384): MEASUREMENT ONLY, results invalid: tiled kernel
                               without the arithmetic (streaming ceiling of the access pattern); bit 15
                               (32768): force the plane kernels; bit 16 (65536): MEASUREMENT ONLY, results
                               invalid: every row chunk marches the same L2-resident rows (the kernel's time
                               without DRAM traffic); bit 17 (131072): TWO subcycles per launch (temporal
                               blocking, csrc/evp_fused.cuh; bit-identical; single rank, no north-south wrap,
                               no T-fold; halves the DRAM traffic but is issue-bound and measured slower,
                               DESIGN.md 4); bit 20 (1048576): plane kernel of 128-thread CTAs with one strip of
                               <= 31 U columns per WARP (east-neighbour str terms by warp shuffle, no exchange
                               line and no CTA barrier in the row loop), bit 21 (2097152): one strip per CTA
                               with the shared-memory exchange line and a barrier per row instead (DESIGN.md 4
                               says which of the two is the default); bit 22 (4194304): evp_finish as a separate
                               kernel after the loop instead of an epilogue of the last subcycle kernel */
    int32_t state_residency; /* 0 = the whole state is uploaded and downloaded by every call (host arrays always
                               current: restart-exact drop-in); 1 = the 12 stress arrays stay on the device
                               between calls: uploaded by the first call after init or after
                               evp_b200_invalidate_device_state, brought back only by evp_b200_download_state
                               (before dumpfile / ice_write_hist); uvel, vvel, iceumask still round-trip;
                               2 = uvel, vvel and iceumask stay on the device as well: nothing of the state
                               travels in either direction; the transport scheme gets the velocities from
                               evp_b200_download_velocity or, on the device, evp_b200_device_velocity */
    int32_t exchange_mode;  /* multi-rank velocity halo inside the ndte loop: 0 = peer-to-peer stores from the
                               subcycle kernel into the neighbour's ghost rows (CUDA IPC over NVLink, flags for
                               ordering), 1 = NCCL send/recv after every subcycle */
} evp_b200_params;

/* module ice_grid arrays read by evp (source/ice_grid.F90:58-85,114-123) + fcor_blk
 * (source/ice_dyn_evp.F90:105-106,503).  Uploaded once by evp_b200_init. */
typedef struct {
    const double *dxt, *dyt, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym;
    const double *tarea, *tarear, *tinyarea, *uarea, *uarear, *fcor;
    const int32_t *tmask, *umask;
    /* optional (may be NULL): the primary cell widths HTE, HTN of module ice_grid
     * (source/ice_grid.F90:75-76).  When given, the library checks at init, row by row and bit for
     * bit, that dxt,dyt,dxhy,dyhx,cxp,cyp,cxm,cym equal the init_grid2 formulas (:350-361,:1196,:1280)
     * applied to them; on rows where they do, the subcycle kernel streams 2 planes instead of 8. */
    const double *HTE, *HTN;
} evp_b200_static_fields;

/* per-call inputs (source/ice_state.F90:66-89, source/ice_flux.F90:45-56) */
typedef struct {
    const double *aice, *vice, *vsno;
    const double *strairxT, *strairyT; /* or strax/stray when wind_from_strax */
    const double *uocn, *vocn;
    const double *ss_tltx, *ss_tlty;   /* may be NULL unless coupled_tilt && use_ocnslope */
    /* only for the device ice_strength (strength == NULL): (nx_block,ny_block,ncat,max_blocks) */
    const double *aice0, *aicen, *vicen;
} evp_b200_inputs;

/* persistent module state, in/out (source/ice_state.F90:132-134, source/ice_flux.F90:85-93) */
typedef struct {
    double *uvel, *vvel;
    double *stressp_1, *stressp_2, *stressp_3, *stressp_4;
    double *stressm_1, *stressm_2, *stressm_3, *stressm_4;
    double *stress12_1, *stress12_2, *stress12_3, *stress12_4;
    int32_t *iceumask;
} evp_b200_state;

/* outputs; any pointer may be NULL = not downloaded.  Cells evp never writes are 0,
 * as after init_history_dyn (source/ice_flux.F90:585-602) in step_dynamics. */
typedef struct {
    double *strairx, *strairy, *strtltx, *strtlty, *strintx, *strinty;
    double *strocnx, *strocny, *strocnxT, *strocnyT, *fm, *prs_sig;
    double *divu, *shear, *rdg_conv, *rdg_shear;
    double *strength;  /* halo-updated strength as left by evp (:337-338) */
    double *sicemass;  /* AusCOM: = tmass (:246-248) */
    double *sig1, *sig2; /* principal_stress epilogue (source/ice_history.F90:1939-1945) */
} evp_b200_outputs;

/* device timings of the last call, milliseconds, CUDA events on the library stream */
typedef struct {
    float upload_ms, prep_ms, subcycle_ms, finish_ms, download_ms, total_ms;
    int32_t kernel_launches;   /* kernels launched (graph nodes count) in the last call */
    int32_t subcycle_launches; /* of which inside the ndte loop */
    int32_t exchange_mode_used; /* multi-rank: 0 = peer-to-peer stores, 1 = NCCL per subcycle; -1 = single rank */
    int32_t reserved;           /* rows of this slab on the 2-plane metric path (HTE/HTN given and verified) */
} evp_b200_timings;

typedef struct evp_b200_handle evp_b200_handle;

int evp_b200_abi_version(void);
const char *evp_b200_last_error(void);
void evp_b200_default_params(evp_b200_params *p);

/* init_evp (source/ice_dyn_evp.F90:441-526): set_evp_parameters, device allocation,
 * static-field upload.  Velocity/stress zeroing stays with the caller's module arrays. */
int evp_b200_init(const evp_b200_dims *dims, const evp_b200_params *params,
                  const evp_b200_static_fields *grid, evp_b200_handle **out);

/* source/ice_dyn_evp.F90:214-316: zero diagnostics, evp_prep1, HALO icetmask, to_ugrid x2,
 * t2ugrid_vector x2, evp_prep2.  icetmask_out (int32, block layout, halo-updated) lets the
 * Fortran shim rebuild icellt/indxti/indxtj (:850-859) for its own ice_strength call. */
int evp_b200_prep(evp_b200_handle *h, const evp_b200_inputs *in, evp_b200_state *st,
                  int32_t *icetmask_out);

/* source/ice_dyn_evp.F90:336-428: HALO strength,u,v; ndte x (stress, stepu, HALO u,v);
 * evp_finish; u2tgrid_vector x2; download of state and outputs. */
int evp_b200_run(evp_b200_handle *h, const double *strength, evp_b200_state *st,
                 evp_b200_outputs *out);

/* prep + run in one call; strength == NULL computes ice_strength on the device
 * (source/ice_mechred.F90:1869-2036; needs aice0/aicen/vicen). */
int evp_b200_step(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength,
                  evp_b200_state *st, evp_b200_outputs *out);

/* Re-run the ndte subcycle loop `repeats` times on the device-resident fields left by the
 * last prep/run (no host traffic); ms_per_loop = mean device time of one ndte loop.  This is
 * the timed region of the headline metric (grid-cell-subcycles/s). */
int evp_b200_subcycle_resident(evp_b200_handle *h, int32_t repeats, float *ms_per_loop);

/* principal_stress (source/ice_dyn_evp.F90:1558-1609) on caller arrays */
int evp_b200_principal_stress(evp_b200_handle *h, const double *stressp_1, const double *stressm_1,
                              const double *stress12_1, const double *prs_sig,
                              double *sig1, double *sig2);

/* principal_stress on the first n elements of the arrays (the reference calls it block by block with
 * (nx_block, ny_block) slices, source/ice_history.F90:1939-1945); n <= nx_block*ny_block*max_blocks */
int evp_b200_principal_stress_n(evp_b200_handle *h, int64_t n, const double *stressp_1, const double *stressm_1,
                                const double *stress12_1, const double *prs_sig, double *sig1, double *sig2);

/* Kinetic energy and rms ice speed of runtime_diags (source/ice_diagnostics.F90:199-234) from the device-
 * resident result and the vice / vsno of the last call, this slab only: out[0] / out[1] = total ice-snow
 * kinetic energy north / south (sum of 0.5*(rhos*vsno + rhoi*vice)*(uvel**2 + vvel**2) * tarean|tareas),
 * out[2] / out[3] = ice volume, out[4] / out[5] = snow volume, out[6] / out[7] = rms ice speed computed from
 * them (several slabs: add out[0..5] over the slabs -- the reference's global_sum -- and apply :221-234).
 * The sums are deterministic with a fixed order (csrc/evp_aux.cu k_energy_rows): thread-strided partial sums
 * of a row combined by a binary tree, rows combined the same way; a value depends on the summation order
 * exactly as the reference's does on its block decomposition (serial/ice_global_reductions.F90:222-246). */
int evp_b200_diagnostics_energy(evp_b200_handle *h, double out[8]);

int evp_b200_get_timings(const evp_b200_handle *h, evp_b200_timings *t);

/* How the ndte loop of this handle runs: out[0] = 1 when the strip-tiled TMA-fed kernel is in use (0: the
 * plane kernels), out[1..2] = grid of the subcycle kernel, out[3] = threads per CTA, out[4] = U columns per
 * strip, out[5] = pipeline stages per warp (tiled kernel), out[6] = 1 with the peer-to-peer halo, out[7] = bit 0
 * the persistent cooperative kernel, bit 1 the plane kernel with one strip per warp (shuffles, no row barrier),
 * bit 2 evp_finish runs as an epilogue of the last subcycle kernel. */
int evp_b200_get_info(const evp_b200_handle *h, int32_t out[8]);

/* pin_host = 1: forget (cudaHostUnregister) a caller array before the caller frees it; unknown pointers
 * are ignored.  Waits for the handle's streams first. */
int evp_b200_unpin(evp_b200_handle *h, const void *host_ptr);

/* Velocity / strength diagnostics of runtime_diags (source/ice_diagnostics.F90:294-346) from the
 * device-resident result of the last call, this slab only (the caller combines slabs with
 * global_maxval): out[0] = max ice speed (m/s), northern hemisphere (lmask_n: ULAT >= -puny),
 * out[1] = southern; out[2] = max strength (kN/m) north, out[3] = south.  Interior cells only. */
int evp_b200_diagnostics(evp_b200_handle *h, double out[4]);

/* state_residency = 1: copy the device-resident stresses (and uvel, vvel, iceumask) into the caller's
 * arrays -- what dumpfile (source/ice_restart.F90:197-246) and ice_write_hist need; and tell the library
 * that the caller changed the host arrays (restartfile, :427-487) so the next call uploads them again. */
int evp_b200_download_state(evp_b200_handle *h, evp_b200_state *st);
int evp_b200_invalidate_device_state(evp_b200_handle *h);

/* state_residency = 2, transport hand-off (source/ice_step_mod.F90:575-585, source/ice_transport_driver.F90:
 * 498-505 read uvel / vvel with their ghost ring right after evp): only the two velocity arrays into the
 * caller's host arrays (block layout) ... */
int evp_b200_download_velocity(evp_b200_handle *h, double *uvel, double *vvel);
/* ... or where they lie on the device: the library's planes of this slab, element (i, j) of plane row j at
 * ptr[j * pitch + i], i in 0 .. nx_global+1, j in 0 .. nrows-1 (ghost ring included; plane (i, j) is Fortran
 * (i+1, j+1) of one block spanning the slab).  Valid until the next call on this handle. */
int evp_b200_device_velocity(evp_b200_handle *h, const double **uvel, const double **vvel, int32_t *pitch,
                             int32_t *nrows);

/* evp_b200_step with every array pointer of in / strength / st / out being DEVICE memory of this handle's
 * device (same block layout (nx_block, ny_block, max_blocks) as the host arrays): for a caller whose
 * thermodynamics / transport already live on the GPU.  No host traffic; the call still synchronises. */
int evp_b200_step_device(evp_b200_handle *h, const evp_b200_inputs *in, const double *strength,
                         evp_b200_state *st, evp_b200_outputs *out);

/* multi-GPU: one handle per rank (y-slab).  id is the 128-byte ncclUniqueId made by rank 0
 * (evp_b200_comm_unique_id) and broadcast by the host (MPI_Bcast in the Fortran world,
 * torch.distributed in the Python harness).  Replaces the per-subcycle ice_HaloUpdate of
 * mpi/ice_boundary.F90:1028-1417 between tasks. */
int evp_b200_comm_unique_id(uint8_t id[128]);
int evp_b200_comm_init(evp_b200_handle *h, const uint8_t id[128]);

/* Self-test of the straight-line IEEE sqrt / division sequences the subcycle kernel uses
 * (csrc/evp_ieee.cuh) against the compiler's sqrt() and operator/ on n generated operands (random bit
 * patterns, EVP magnitudes, operands around every range check, subnormal / huge / zero / exact cases),
 * on the current device.  out[0] = sqrt results that took the fast path and differ in any bit, out[1] =
 * sqrt operands on the fast path, out[2] / out[3] the same for n/d, out[4] = mismatches of a second
 * quotient sharing the refined reciprocal, out[5] = n.  Bit-exactness of the kernel against
 * source/ice_dyn_evp.F90:1095-1098,1131-1134,1426-1427 rests on out[0] = out[2] = out[4] = 0. */
int evp_b200_selftest_ieee(int64_t n, uint64_t seed, uint64_t out[6]);

int evp_b200_finalize(evp_b200_handle *h);

#ifdef __cplusplus
}
#endif
#endif

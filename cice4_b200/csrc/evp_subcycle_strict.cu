// Unfused build of the subcycle kernel: compiled with -fmad=false so every multiply and add
// rounds separately, bit-identical to the -ffp-contract=off CPU oracle.
#define EVP_SUB_NS evp_sub_strict
#define EVP_SUB_LAUNCH evp_subcycle_launch_strict
#define EVP_PERSIST_LAUNCH evp_persist_launch_strict
#define EVP_SUB_CONFIGURE evp_subcycle_configure_strict
#define EVP_TILED_LAUNCH evp_tiled_launch_strict
#include "evp_subcycle_body.cuh"

int evp_subcycle_max_threads(void) { return 256; }

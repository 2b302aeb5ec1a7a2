// FMA-contracted build of the subcycle kernel: compiled with -fmad=true.  Differences from the
// unfused build are rounding-level only (DESIGN.md states the measured bound).
#define EVP_SUB_NS evp_sub_fast
#define EVP_SUB_LAUNCH evp_subcycle_launch_fast
#define EVP_PERSIST_LAUNCH evp_persist_launch_fast
#define EVP_SUB_CONFIGURE evp_subcycle_configure_fast
#define EVP_TILED_LAUNCH evp_tiled_launch_fast
#include "evp_subcycle_body.cuh"

// Unfused build of the two-subcycle kernel (evp_fused.cuh): compiled with -fmad=false, bit-identical to the
// unfused CPU oracle and to two launches of the one-subcycle kernel.
#define EVP_SUB_NS evp_fused_strict
#define EVP_BODY_NO_LAUNCHERS
#define EVP_FUSED_LAUNCH evp_fused_launch_strict
#include "evp_subcycle_body.cuh"
#include "evp_fused.cuh"

// FMA-contracted build of the two-subcycle kernel (evp_fused.cuh): compiled with -fmad=true.
#define EVP_SUB_NS evp_fused_fast
#define EVP_BODY_NO_LAUNCHERS
#define EVP_FUSED_LAUNCH evp_fused_launch_fast
#include "evp_subcycle_body.cuh"
#include "evp_fused.cuh"

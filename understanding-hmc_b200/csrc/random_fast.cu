// placeholder until the FFMA2 kernel lands
#include "hmc_common.cuh"
bool hmc_random_fast_supported(const hmc_random_args& a, const char** why) { *why = "fast kernel not built yet"; return false; }
int hmc_random_run_fast(const hmc_random_args& a, cudaStream_t stream) { hmc_set_error("fast kernel not built yet"); return HMC_E_UNSUPPORTED; }

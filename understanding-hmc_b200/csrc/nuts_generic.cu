// placeholder until the NUTS kernel lands
#include "hmc_common.cuh"
int hmc_nuts_run_generic(const hmc_nuts_args& a, cudaStream_t stream) { hmc_set_error("NUTS kernel not built yet"); return HMC_E_UNSUPPORTED; }

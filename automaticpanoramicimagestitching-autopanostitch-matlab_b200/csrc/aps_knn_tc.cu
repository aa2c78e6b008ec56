// placeholder replaced by the tcgen05 kernel
#include "aps_common.cuh"
int aps_k_knn_tc_supported(int Dp) { (void)Dp; return 0; }
int aps_k_knn_tc(cudaStream_t s, int sm_count, const aps_tc_problem& p) {
  (void)s; (void)sm_count; (void)p;
  aps_set_error(APS_ERR_ARGS, "", "tcgen05 path not built");
  return APS_ERR_ARGS;
}

/* placeholder: filled in by the GICP milestone */
#include "b2icp_oracle.h"
int b2o_align_gicp_impl(const b2icp_params*, const float*, size_t, const float*, size_t, const float*,
                        b2icp_result*, float*, int, int32_t*, float*, b2o_stage_ms*, int) {
  return B2ICP_ERR_INVALID_ARG;
}
extern "C" int b2o_covariances(const float*, size_t, int, double, double*) { return B2ICP_ERR_INVALID_ARG; }

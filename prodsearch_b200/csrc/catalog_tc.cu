// G5 (tensor-core mode): tcgen05 TF32 shortlist + exact rescoring.  Placeholder until the
// kernel lands; the exact mode in catalog_topk.cu is complete.
#include "catalog_common.cuh"

namespace psb {

int64_t tc_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }

int catalog_topk_tc(const float*, int64_t, const float*, int64_t, int64_t, const float*, int64_t, int64_t, int64_t,
                    void*, int64_t, int64_t*, float*, cudaStream_t) {
  return PSB_E_UNSUPPORTED;
}

}  // namespace psb

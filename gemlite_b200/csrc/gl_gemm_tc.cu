// tcgen05 tensor-core GEMM (placeholder until the kernel lands; keeps the link closed).
#include "gl_common.cuh"
namespace glb {
bool gemm_tc_supported(const glb_linear&, int64_t) { return false; }
size_t gemm_tc_workspace(const glb_linear&, int64_t) { return 0; }
int launch_gemm_tc(const glb_linear&, const void*, const float*, void*, int64_t, void*, size_t, const void*, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return GLB_ERR_UNSUPPORTED;
}
}  // namespace glb

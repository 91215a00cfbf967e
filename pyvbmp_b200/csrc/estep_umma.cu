// K2 (tcgen05 variant) — placeholder until the UMMA kernel lands: reports "not supported" so
// vbmp_estep takes the CUDA-core kernel.
#include "common.cuh"
namespace vbmp {
bool estep_umma_supported(long long, int, int, int, int, int, int) { return false; }
size_t estep_umma_workspace_bytes(long long, int, int, int, int) { return 0; }
int launch_estep_umma(const EstepArgs&, int, void*, size_t, float*, float*, cudaStream_t) {
  set_error("estep_umma: not built"); return VBMP_ERR_UNSUPPORTED;
}
}  // namespace vbmp

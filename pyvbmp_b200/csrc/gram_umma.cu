// K3 (tcgen05 variant) — placeholder until the UMMA kernel lands.
#include "common.cuh"
namespace vbmp {
bool gram_umma_supported(long long, int, int, int, int, int, int, int, bool) { return false; }
size_t gram_umma_workspace_bytes(long long, int, int, int, int, int) { return 0; }
int launch_gram_umma(const GramArgs&, float*, void*, size_t, cudaStream_t) {
  set_error("gram_umma: not built"); return VBMP_ERR_UNSUPPORTED;
}
}  // namespace vbmp

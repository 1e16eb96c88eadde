// tcgen05 / TMEM scan -- placeholder until the kernel lands (next commit).
#include "scan_tcgen05.cuh"

namespace rl {
bool tcgen05_supported(const rl_scan_params*) { return false; }
size_t tcgen05_qimg_bytes(int, int) { return 0; }
int tcgen05_prepare_queries(const rl_scan_params*, const float*, float*, void*, cudaStream_t) {
  set_error("tcgen05 scan not built");
  return RL_EUNSUPPORTED;
}
int launch_scan_tcgen05(const ScanArgs&, const rl_scan_params*, const float*, const void*, int, cudaStream_t) {
  set_error("tcgen05 scan not built");
  return RL_EUNSUPPORTED;
}
}  // namespace rl

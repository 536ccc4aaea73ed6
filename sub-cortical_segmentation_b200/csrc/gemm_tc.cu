// tcgen05 / TMEM / TMA back-end (placeholder until the kernel lands)
#include "common.cuh"
namespace sc {
int tc_init(sc_ctx*) { return SC_ERR_UNSUPPORTED; }
void tc_destroy(sc_ctx*) {}
int launch_gemm_tc(sc_ctx*, const GemmProblem&, const GemmW&, cudaStream_t) {
  set_error("tcgen05 back-end not built");
  return SC_ERR_UNSUPPORTED;
}
}  // namespace sc

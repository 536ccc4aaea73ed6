// training step (placeholder until the kernels land)
#include "common.cuh"
namespace sc {
int train_forward_backward(sc_ctx*, const float*, const float*, const float*, const float*, const uint8_t*, int64_t,
                           int64_t, uint64_t, const uint8_t*, float*, cudaStream_t) {
  set_error("training kernels not built");
  return SC_ERR_UNSUPPORTED;
}
int adam_step(sc_ctx*, float, float, float, float, float, cudaStream_t) {
  set_error("training kernels not built");
  return SC_ERR_UNSUPPORTED;
}
int eval_batch(sc_ctx*, const float*, const float*, const float*, const float*, const uint8_t*, int64_t, float*,
               cudaStream_t) {
  set_error("training kernels not built");
  return SC_ERR_UNSUPPORTED;
}
}  // namespace sc

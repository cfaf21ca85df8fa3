// Rollout kernel instantiations, noise mode: philox_f64.
#include "discrete_launch.h"

namespace mdpp {
int launch_rollout_philox_f64(mdpp_ctx* ctx, RolloutParams& p, cudaStream_t stream) {
  return launch_rollout<MDPP_NOISE_PHILOX, 0>(ctx, p, stream);
}
}  // namespace mdpp

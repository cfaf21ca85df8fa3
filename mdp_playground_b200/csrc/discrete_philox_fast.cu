// Rollout kernel instantiations, noise mode: philox_fast.
#include "discrete_launch.h"

namespace mdpp {
int launch_rollout_philox_fast(mdpp_ctx* ctx, RolloutParams& p, cudaStream_t stream) {
  return launch_rollout<MDPP_NOISE_PHILOX, 1>(ctx, p, stream);
}
}  // namespace mdpp

// Rollout kernel instantiations, noise mode: off.
#include "discrete_launch.h"

namespace mdpp {
int launch_rollout_off(mdpp_ctx* ctx, RolloutParams& p, cudaStream_t stream) {
  return launch_rollout<MDPP_NOISE_OFF, 0>(ctx, p, stream);
}
}  // namespace mdpp

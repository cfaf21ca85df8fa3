// Rollout kernel instantiations, noise mode: replay.
#include "discrete_launch.h"

namespace mdpp {
int launch_rollout_replay(mdpp_ctx* ctx, RolloutParams& p, cudaStream_t stream) {
  return launch_rollout<MDPP_NOISE_REPLAY, 0>(ctx, p, stream);
}
}  // namespace mdpp

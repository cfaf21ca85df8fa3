// Rollout kernel instantiations, noise mode: philox, ziggurat fp64 normals.
#include "discrete_launch.h"

namespace mdpp {
int launch_rollout_philox_zig(mdpp_ctx* ctx, RolloutParams& p, cudaStream_t stream) {
  return launch_rollout<MDPP_NOISE_PHILOX, MDPP_NORMAL_ZIGGURAT>(ctx, p, stream);
}
}  // namespace mdpp

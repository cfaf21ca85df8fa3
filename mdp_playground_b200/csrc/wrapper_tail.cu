// GymEnvWrapper's post-processing tail for external vector environments
// (envs/gym_env_wrapper.py:350-439, :523-618): action
// substitution noise, observation noise, reward delay FIFO with the flush at
// episode end, reward noise / scale / shift, padded image shift.  Elementwise
// over environments, HBM-bound streams; oracle/wrapper_tail.py restates it and
// is pinned to the reference wrapper's golden vectors.
#include <cstring>

#include "internal.h"
#include "philox.cuh"

namespace mdpp {
namespace {

// Noise normals: Box-Muller on Philox words, in fp64 (MDPP_NORMAL_F64, the
// default) or on the SFU in fp32 (MDPP_NORMAL_FAST).  (numpy's ziggurat was
// tried here too: with one or two draws per thread, half of all warps run its
// slow path for a single lane -- 346 instead of 230 instructions per pair.)
enum : uint32_t {
  STREAM_TAIL_ACTION = 48,   // counter = (env, step): w0 decides / picks
  STREAM_TAIL_NORMAL = 49,   // (w0, w1) -> reward normal
  STREAM_TAIL_SHIFT = 50,    // w0, w1 -> the two shift integers
  STREAM_TAIL_OBS = 0x10000, // + pair index: (w0, w1) -> the normals of
                             // dimensions 2 pair, 2 pair + 1
};
constexpr int kTBlock = 128;

struct TailParams {
  mdpp_tail_config cfg;
  mdpp_tail_state st;
  int64_t n;
  int32_t noise_mode, normal_mode;
  uint32_t k0, k1;
  uint32_t rk[20];     // Philox round keys
  uint64_t step_index;
  int64_t env_id_offset;
};

__global__ void __launch_bounds__(kTBlock)
tail_actions_kernel(const __grid_constant__ TailParams p, const int32_t* actions,
                    int32_t* applied, const double* replay_u) {
  const int64_t i = (int64_t)blockIdx.x * kTBlock + threadIdx.x;
  if (i >= p.n) return;
  const int n = p.cfg.n_actions;
  int32_t a = actions[i];
  if (p.cfg.has_transition_noise && n > 1) {
    const double pn = p.cfg.transition_noise;
    if (p.noise_mode == MDPP_NOISE_REPLAY) {
      // probs = ones(n) * p / (n - 1); probs[a] = 1 - p; cdf = cumsum(probs);
      // cdf /= cdf[-1]; searchsorted(cdf, u, side="right")
      const double q = __ddiv_rn(__dmul_rn(1.0, pn), (double)(n - 1));
      const double keep = __dadd_rn(1.0, -pn);
      double total = 0.0;
      for (int k = 0; k < n; ++k) total = __dadd_rn(total, k == a ? keep : q);
      const double u = replay_u[i];
      double s = 0.0;
      int idx = 0;
      for (int k = 0; k < n; ++k) {
        s = __dadd_rn(s, k == a ? keep : q);
        idx += __ddiv_rn(s, total) <= u;
      }
      a = min(idx, n - 1);
    } else {
      const uint32_t gid = (uint32_t)(p.env_id_offset + i);
      const U4 w = philox4x32_10(gid, (uint32_t)p.step_index,
                                 (uint32_t)(p.step_index >> 32), STREAM_TAIL_ACTION,
                                 p.k0, p.k1);
      // noisy iff w0 < round(p 2^32); then uniform over the n - 1 others
      const double t = floor(pn * 4294967296.0 + 0.5);
      const uint64_t T = t <= 0.0 ? 0ull : t >= 4294967296.0 ? (1ull << 32) : (uint64_t)t;
      if ((uint64_t)w.x < T) a = (a + 1 + (int32_t)__umulhi(w.y, (uint32_t)(n - 1))) % n;
    }
  }
  applied[i] = a;
}

// Observation noise (:367-376, :405-406): next_obs += noise, a float64 draw
// added into the observation's dtype.  One thread per (env, pair of
// dimensions): consecutive threads touch consecutive elements.
template <typename R>
__global__ void __launch_bounds__(kTBlock)
tail_obs_kernel(const __grid_constant__ TailParams p, const R* obs, R* out_obs,
                const double* replay_on) {
  const mdpp_tail_config& c = p.cfg;
  const int D = c.obs_dim, P = (D + 1) / 2;
  const int64_t idx = (int64_t)blockIdx.x * kTBlock + threadIdx.x;
  if (idx >= p.n * P) return;
  const int64_t i = idx / P;
  const int pr = (int)(idx - i * P), d0 = 2 * pr;
  const bool two = d0 + 1 < D;
  const int64_t at = i * D + d0;
  double z0 = 0.0, z1 = 0.0;
  if (c.has_transition_noise) {
    if (p.noise_mode == MDPP_NOISE_REPLAY) {
      z0 = replay_on[at];
      if (two) z1 = replay_on[at + 1];
    } else {
      const uint32_t gid = (uint32_t)(p.env_id_offset + i);
      const uint32_t s0 = (uint32_t)p.step_index, s1 = (uint32_t)(p.step_index >> 32);
      const U4 w = philox4x32_10_rk(gid, s0, s1, STREAM_TAIL_OBS + (uint32_t)pr, p.rk);
      if (p.normal_mode == MDPP_NORMAL_FAST) normal_pair_fast(w.x, w.y, &z0, &z1);
      else normal_pair_f64(w.x, w.y, &z0, &z1);
      z0 = __dmul_rn(c.transition_noise, z0);
      z1 = __dmul_rn(c.transition_noise, z1);
    }
  }
  out_obs[at] = (R)__dadd_rn((double)obs[at], z0);
  if (two) out_obs[at + 1] = (R)__dadd_rn((double)obs[at + 1], z1);
}

__global__ void __launch_bounds__(kTBlock)
tail_post_kernel(const __grid_constant__ TailParams p, const double* reward,
                 const uint8_t* done, double* out_reward, const double* replay_rn) {
  const int64_t i = (int64_t)blockIdx.x * kTBlock + threadIdx.x;
  if (i >= p.n) return;
  const mdpp_tail_config& c = p.cfg;
  const uint32_t gid = (uint32_t)(p.env_id_offset + i);
  const uint32_t s0 = (uint32_t)p.step_index, s1 = (uint32_t)(p.step_index >> 32);
  // ---- reward tail (:411-436) ---------------------------------------------
  double r = reward[i];
  const int delay = c.delay;
  int32_t tl = p.st.t_episode[i];
  const int64_t N = p.st.n_envs;
  if (done[i]) {
    // reward += np.sum(reward_buffer * reward_scale + reward_shift): all `delay`
    // entries, oldest first (entries older than the last clear are the 0.0s the
    // reference's reset() puts there); numpy sums < 8 elements sequentially and
    // up to 128 with 8 accumulators
    auto term = [&](int k) {  // k-th oldest entry
      const int age = delay - k;  // written `age` steps ago
      const double b = age <= tl ? p.st.ring[((p.step_index + (uint64_t)delay - (uint64_t)age)
                                              % (uint64_t)delay) * N + i] : 0.0;
      return __dadd_rn(__dmul_rn(b, c.reward_scale), c.reward_shift);
    };
    double sum = 0.0;
    if (delay < 8) {
      for (int k = 0; k < delay; ++k) sum = __dadd_rn(sum, term(k));
    } else {
      double acc[8];
      for (int j = 0; j < 8; ++j) acc[j] = term(j);
      int k = 8;
      for (; k < delay - (delay % 8); k += 8)
        for (int j = 0; j < 8; ++j) acc[j] = __dadd_rn(acc[j], term(k + j));
      sum = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                      __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
      for (; k < delay; ++k) sum = __dadd_rn(sum, term(k));
    }
    if (delay > 0) r = __dadd_rn(r, sum);
    else r = __dadd_rn(r, 0.0);  // np.sum([]) = 0.0
    r = __dadd_rn(r, __dmul_rn(c.term_state_reward, c.reward_scale));
    tl = 0;  // the reset() that follows re-creates the buffer (:466)
  } else {
    if (delay > 0) {
      double* slot = p.st.ring + (p.step_index % (uint64_t)delay) * N + i;
      const double delayed = tl >= delay ? *slot : 0.0;
      *slot = r;
      r = delayed;
    }
    tl += 1;
  }
  p.st.t_episode[i] = tl;
  if (c.has_reward_noise) {
    double nz;
    if (p.noise_mode == MDPP_NOISE_REPLAY) {
      nz = replay_rn[i];
    } else {
      const U4 w = philox4x32_10_rk(gid, s0, s1, STREAM_TAIL_NORMAL, p.rk);
      double z0, z1;
      if (p.normal_mode == MDPP_NORMAL_FAST) normal_pair_fast(w.x, w.y, &z0, &z1);
      else normal_pair_f64(w.x, w.y, &z0, &z1);
      nz = __dmul_rn(c.reward_noise_std, z0);
    }
    r = __dadd_rn(r, nz);
  }
  r = __dmul_rn(r, c.reward_scale);
  r = __dadd_rn(r, c.reward_shift);
  out_reward[i] = r;
}

// One CTA per image: out[x][y][c] = canvas[y][x][c], the env image pasted at
// rows [sh - side/2, sh + side/2), columns [sw - side/2, sw + side/2).
__global__ void __launch_bounds__(256)
tail_image_shift_kernel(const __grid_constant__ TailParams p, const uint8_t* img,
                        uint8_t* out, const int32_t* replay_shift, int32_t* shift_out) {
  const int64_t i = blockIdx.x;
  const mdpp_tail_config& c = p.cfg;
  const int side = c.image_side, pad = c.image_padding, tot = side + 2 * pad;
  int raw_w = 0, raw_h = 0;
  if (c.has_shift) {
    if (p.noise_mode == MDPP_NOISE_REPLAY) {
      raw_w = replay_shift[2 * i];
      raw_h = replay_shift[2 * i + 1];
    } else {
      // integers(-m + 1, m), m = (tot - side) // 2 = padding: 2 m - 1 values
      const U4 w = philox4x32_10((uint32_t)(p.env_id_offset + i), (uint32_t)p.step_index,
                                 (uint32_t)(p.step_index >> 32), STREAM_TAIL_SHIFT,
                                 p.k0, p.k1);
      const uint32_t span = (uint32_t)max(2 * pad - 1, 1);
      raw_w = -pad + 1 + (int)__umulhi(w.x, span);
      raw_h = -pad + 1 + (int)__umulhi(w.y, span);
    }
  }
  if (shift_out && threadIdx.x == 0) {
    shift_out[2 * i] = raw_w;
    shift_out[2 * i + 1] = raw_h;
  }
  const int q = max(c.sh_quant, 1);
  const int sw = tot / 2 + (c.has_shift ? (raw_w / q) * q : 0);  // int(x / q) * q: truncation
  const int sh = tot / 2 + (c.has_shift ? (raw_h / q) * q : 0);
  const int top = sh - side / 2, left = sw - side / 2;
  const int half2 = 2 * (side / 2);  // rows / columns actually pasted
  const uint8_t* src = img + i * (int64_t)side * side * 3;
  uint8_t* dst = out + i * (int64_t)tot * tot * 3;
  for (int e = threadIdx.x; e < tot * tot; e += blockDim.x) {
    const int x = e / tot, y = e - x * tot;  // out[x][y] = canvas[y][x]
    const int r = y - top, col = x - left;
    uint8_t v0 = 0, v1 = 0, v2 = 0;
    if (r >= 0 && r < half2 && col >= 0 && col < half2) {
      const uint8_t* px = src + ((int64_t)r * side + col) * 3;
      v0 = px[0]; v1 = px[1]; v2 = px[2];
    }
    uint8_t* o = dst + (int64_t)e * 3;
    o[0] = v0; o[1] = v1; o[2] = v2;
  }
}

// The same paste for even canvas sides, HBM-friendly: the env image goes
// TRANSPOSED into shared memory (coalesced global loads), and every thread
// emits 4 consecutive pixels of the transposed canvas = 12 contiguous output
// bytes, read from the tile as four aligned words and funnel-shifted, stored
// as three 32-bit words.
__device__ __forceinline__ void tail_shift_params(const TailParams& p, int64_t i,
                                                  const int32_t* replay_shift,
                                                  int32_t* shift_out, int* top, int* left) {
  const mdpp_tail_config& c = p.cfg;
  const int side = c.image_side, pad = c.image_padding, tot = side + 2 * pad;
  int raw_w = 0, raw_h = 0;
  if (c.has_shift) {
    if (p.noise_mode == MDPP_NOISE_REPLAY) {
      raw_w = replay_shift[2 * i];
      raw_h = replay_shift[2 * i + 1];
    } else {
      // integers(-m + 1, m), m = (tot - side) // 2 = padding: 2 m - 1 values
      const U4 w = philox4x32_10((uint32_t)(p.env_id_offset + i), (uint32_t)p.step_index,
                                 (uint32_t)(p.step_index >> 32), STREAM_TAIL_SHIFT,
                                 p.k0, p.k1);
      const uint32_t span = (uint32_t)max(2 * pad - 1, 1);
      raw_w = -pad + 1 + (int)__umulhi(w.x, span);
      raw_h = -pad + 1 + (int)__umulhi(w.y, span);
    }
  }
  if (shift_out && threadIdx.x == 0) {
    shift_out[2 * i] = raw_w;
    shift_out[2 * i + 1] = raw_h;
  }
  const int q = max(c.sh_quant, 1);
  const int sw = tot / 2 + (c.has_shift ? (raw_w / q) * q : 0);  // int(x / q) * q: truncation
  const int sh = tot / 2 + (c.has_shift ? (raw_h / q) * q : 0);
  *top = sh - side / 2;
  *left = sw - side / 2;
}

// Bytes of the shared-memory tile: the env image TRANSPOSED, tile[col][4 + row][3]
// with 4 zero rows before and after each column (an item that the image's first
// or last row cuts through reads its zeros from there), an odd number of 32-bit
// words per column (conflict-free in both phases) and 16 bytes of slack for the
// 4-word reads of the last item.
constexpr int kTailRowPad = 4;
__host__ __device__ inline int tail_tile_col_bytes(int side) {
  const int words = ((side + 2 * kTailRowPad) * 3 + 3) / 4;
  return 4 * (words | 1);
}
__host__ __device__ inline size_t tail_tile_bytes(int side) {
  return (size_t)side * tail_tile_col_bytes(side) + 16;
}

__global__ void __launch_bounds__(256)
tail_image_shift_smem_kernel(const __grid_constant__ TailParams p, const uint8_t* img,
                             uint8_t* out, const int32_t* replay_shift,
                             int32_t* shift_out) {
  extern __shared__ __align__(16) uint8_t tile[];  // [col][4 + row][3], see above
  const int64_t i = blockIdx.x;
  const mdpp_tail_config& c = p.cfg;
  const int side = c.image_side, tot = side + 2 * c.image_padding;  // (side is even)
  const int cstride = tail_tile_col_bytes(side);
  const uint8_t* src = img + i * (int64_t)side * side * 3;
  // the zero rows of every column
  for (int k = threadIdx.x; k < side * 2 * kTailRowPad * 3; k += blockDim.x) {
    const int col = k / (2 * kTailRowPad * 3), b = k - col * (2 * kTailRowPad * 3);
    tile[col * cstride + (b < kTailRowPad * 3 ? b : b + side * 3)] = 0;
  }
  // phase 1: consecutive lanes read consecutive pixels and write them a column
  // apart (an odd number of words: 32 different banks).  Sides that are a
  // multiple of 4: four pixels = three aligned words per thread, two such
  // groups in flight.
  const float inv_side = 1.0f / (float)side;
  if ((side & 3) == 0) {
    const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src);
    const int groups = side * side / 4;
    auto scatter = [&](int g, uint32_t a0, uint32_t a1, uint32_t a2) {
      const int px = 4 * g;
      const int r = (int)(((float)px + 0.5f) * inv_side);  // exact: see render.cu
      const int col = px - r * side;                       // col .. col + 3: one row
      uint8_t* d = tile + col * cstride + 3 * (r + kTailRowPad);
      d[0] = (uint8_t)a0; d[1] = (uint8_t)(a0 >> 8); d[2] = (uint8_t)(a0 >> 16);
      d += cstride;
      d[0] = (uint8_t)(a0 >> 24); d[1] = (uint8_t)a1; d[2] = (uint8_t)(a1 >> 8);
      d += cstride;
      d[0] = (uint8_t)(a1 >> 16); d[1] = (uint8_t)(a1 >> 24); d[2] = (uint8_t)a2;
      d += cstride;
      d[0] = (uint8_t)(a2 >> 8); d[1] = (uint8_t)(a2 >> 16); d[2] = (uint8_t)(a2 >> 24);
    };
    int g = threadIdx.x;
    for (; g + (int)blockDim.x < groups; g += 2 * blockDim.x) {
      const int h = g + blockDim.x;
      const uint32_t a0 = __ldcs(src4 + 3 * g), a1 = __ldcs(src4 + 3 * g + 1),
                     a2 = __ldcs(src4 + 3 * g + 2);
      const uint32_t b0 = __ldcs(src4 + 3 * h), b1 = __ldcs(src4 + 3 * h + 1),
                     b2 = __ldcs(src4 + 3 * h + 2);
      scatter(g, a0, a1, a2);
      scatter(h, b0, b1, b2);
    }
    if (g < groups)
      scatter(g, __ldcs(src4 + 3 * g), __ldcs(src4 + 3 * g + 1), __ldcs(src4 + 3 * g + 2));
  } else {
    for (int px = threadIdx.x; px < side * side; px += blockDim.x) {
      const int r = (int)(((float)px + 0.5f) * inv_side);
      const int col = px - r * side;
      const uint8_t* s3 = src + 3 * px;
      const uint8_t v0 = __ldcs(s3), v1 = __ldcs(s3 + 1), v2 = __ldcs(s3 + 2);
      uint8_t* d = tile + col * cstride + 3 * (r + kTailRowPad);
      d[0] = v0; d[1] = v1; d[2] = v2;
    }
  }
  int top, left;
  tail_shift_params(p, i, replay_shift, shift_out, &top, &left);
  __syncthreads();
  // phase 2: 4 consecutive pixels of the transposed canvas per thread = 12
  // contiguous output bytes = three 32-bit stores.  out[x][y] = canvas[y][x],
  // flat pixel index e = x tot + y.
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + i * (int64_t)tot * tot * 3);
  const float inv_tot = 1.0f / (float)tot;
  for (int e4 = threadIdx.x; e4 < tot * tot / 4; e4 += blockDim.x) {
    const int e = 4 * e4;
    const int x = (int)(((float)e + 0.5f) * inv_tot);
    const int y = e - x * tot;
    const int r = y - top, col = x - left;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (y + 3 < tot) {  // the four pixels share column x of the output
      if ((unsigned)col < (unsigned)side && r + 3 >= 0 && r < side) {
        // 12 contiguous tile bytes (zero rows included) at any alignment: four
        // aligned words, funnel-shifted
        const uint32_t a = (uint32_t)(col * cstride + 3 * (r + kTailRowPad));
        const uint32_t* q = reinterpret_cast<const uint32_t*>(tile + (a & ~3u));
        const uint32_t sh = (a & 3u) * 8u;
        const uint32_t t0 = q[0], t1 = q[1], t2 = q[2], t3 = q[3];
        w0 = __funnelshift_r(t0, t1, sh);
        w1 = __funnelshift_r(t1, t2, sh);
        w2 = __funnelshift_r(t2, t3, sh);
      }
    } else {  // (canvas sides that are not a multiple of 4: the item wraps)
      uint32_t b[12];
      int xx = x, yy = y;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = yy - top, cc = xx - left;
        const bool in = (unsigned)rr < (unsigned)side && (unsigned)cc < (unsigned)side;
        const uint8_t* px = tile + (in ? cc * cstride + 3 * (rr + kTailRowPad) : 0);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) b[3 * k + ch] = in ? px[ch] : 0u;
        if (++yy == tot) { yy = 0; ++xx; }
      }
      w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
    uint32_t* o = dst + 3 * e4;
    __stcs(o, w0); __stcs(o + 1, w1); __stcs(o + 2, w2);
  }
}

int fill(mdpp_ctx* ctx, const mdpp_tail_config* cfg, const mdpp_step_opts* opts,
         int64_t n, TailParams* p) {
  if (!ctx) return MDPP_EINVAL;
  if (!cfg || !opts || n < 1) return fail(ctx, MDPP_EINVAL, "tail: bad arguments");
  if (opts->noise_mode != MDPP_NOISE_REPLAY && opts->noise_mode != MDPP_NOISE_PHILOX &&
      opts->noise_mode != MDPP_NOISE_OFF)
    return fail(ctx, MDPP_EINVAL, "tail: unknown noise_mode");
  std::memset(p, 0, sizeof *p);
  p->cfg = *cfg;
  p->n = n;
  p->noise_mode = opts->noise_mode;
  p->k0 = (uint32_t)opts->seed;
  p->k1 = (uint32_t)(opts->seed >> 32);
  philox_round_keys(p->k0, p->k1, p->rk);
  p->normal_mode = opts->normal_mode == MDPP_NORMAL_FAST ? MDPP_NORMAL_FAST : MDPP_NORMAL_F64;
  p->step_index = opts->step_index;
  p->env_id_offset = opts->env_id_offset;
  return MDPP_OK;
}

}  // namespace
}  // namespace mdpp

using namespace mdpp;

extern "C" int mdpp_tail_actions(mdpp_ctx* ctx, const mdpp_tail_config* cfg,
                                 const int32_t* actions, int32_t* applied,
                                 const double* replay_u, int64_t n_envs,
                                 const mdpp_step_opts* opts, void* cuda_stream) {
  TailParams p;
  int rc = fill(ctx, cfg, opts, n_envs, &p);
  if (rc) return rc;
  if (!actions || !applied || cfg->n_actions < 1)
    return fail(ctx, MDPP_EINVAL, "tail_actions: bad arguments");
  if (cfg->has_transition_noise && opts->noise_mode == MDPP_NOISE_REPLAY && !replay_u)
    return fail(ctx, MDPP_EINVAL, "tail_actions: replay mode needs replay_u");
  MDPP_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)((n_envs + kTBlock - 1) / kTBlock);
  tail_actions_kernel<<<grid, kTBlock, 0, (cudaStream_t)cuda_stream>>>(
      p, actions, applied, replay_u);
  MDPP_CUDA(ctx, cudaGetLastError());
  return MDPP_OK;
}

extern "C" int mdpp_tail_post(mdpp_ctx* ctx, const mdpp_tail_config* cfg,
                              const mdpp_tail_state* st, const void* obs,
                              void* out_obs, const double* reward,
                              const uint8_t* done, double* out_reward,
                              const double* replay_reward_noise,
                              const double* replay_obs_noise,
                              const mdpp_step_opts* opts, void* cuda_stream) {
  TailParams p;
  int rc = fill(ctx, cfg, opts, st ? st->n_envs : 0, &p);
  if (rc) return rc;
  if (!st->t_episode || (cfg->delay > 0 && !st->ring) || cfg->delay < 0)
    return fail(ctx, MDPP_EINVAL, "tail_post: state arrays missing");
  if (!reward || !done || !out_reward)
    return fail(ctx, MDPP_EINVAL, "tail_post: reward / done / out_reward are required");
  if (opts->noise_mode == MDPP_NOISE_REPLAY &&
      ((cfg->has_reward_noise && !replay_reward_noise) ||
       (!cfg->discrete && cfg->has_transition_noise && obs && !replay_obs_noise)))
    return fail(ctx, MDPP_EINVAL, "tail_post: replay mode needs the replay arrays");
  p.st = *st;
  MDPP_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)((st->n_envs + kTBlock - 1) / kTBlock);
  cudaStream_t s = (cudaStream_t)cuda_stream;
  if (!cfg->discrete && obs && out_obs) {
    if (cfg->obs_dim < 1) return fail(ctx, MDPP_EINVAL, "tail_post: obs_dim < 1");
    const int64_t pairs = st->n_envs * ((cfg->obs_dim + 1) / 2);
    if (pairs > 0x7fffffffLL * kTBlock)
      return fail(ctx, MDPP_EINVAL, "tail_post: too many observation elements per call");
    const unsigned og = (unsigned)((pairs + kTBlock - 1) / kTBlock);
    if (cfg->obs_is_f64)
      tail_obs_kernel<double><<<og, kTBlock, 0, s>>>(p, (const double*)obs,
                                                     (double*)out_obs, replay_obs_noise);
    else
      tail_obs_kernel<float><<<og, kTBlock, 0, s>>>(p, (const float*)obs, (float*)out_obs,
                                                    replay_obs_noise);
  }
  tail_post_kernel<<<grid, kTBlock, 0, s>>>(p, reward, done, out_reward,
                                            replay_reward_noise);
  MDPP_CUDA(ctx, cudaGetLastError());
  return MDPP_OK;
}

extern "C" int mdpp_tail_image_shift(mdpp_ctx* ctx, const mdpp_tail_config* cfg,
                                     const uint8_t* img, uint8_t* out,
                                     const int32_t* replay_shift, int32_t* shift_out,
                                     int64_t n_envs, const mdpp_step_opts* opts,
                                     void* cuda_stream) {
  TailParams p;
  int rc = fill(ctx, cfg, opts, n_envs, &p);
  if (rc) return rc;
  if (!img || !out || cfg->image_side < 1 || cfg->image_padding < 0 ||
      cfg->image_channels != 3)
    return fail(ctx, MDPP_EINVAL, "tail_image_shift: square RGB images only");
  if (cfg->has_shift && opts->noise_mode == MDPP_NOISE_REPLAY && !replay_shift)
    return fail(ctx, MDPP_EINVAL, "tail_image_shift: replay mode needs replay_shift");
  MDPP_CUDA(ctx, cudaSetDevice(ctx->device));
  const int tot = cfg->image_side + 2 * cfg->image_padding;
  const size_t tile_bytes = tail_tile_bytes(cfg->image_side);
  if (cfg->image_side % 2 == 0 && tile_bytes <= (size_t)ctx->max_smem_optin - 1024 &&
      ((uintptr_t)img & 3) == 0 && ((uintptr_t)out & 3) == 0) {
    if (tile_bytes > 48 * 1024 - 512)
      MDPP_CUDA(ctx, cudaFuncSetAttribute(tail_image_shift_smem_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)tile_bytes));
    tail_image_shift_smem_kernel<<<(unsigned)n_envs, 256, tile_bytes,
                                   (cudaStream_t)cuda_stream>>>(
        p, img, out, replay_shift, shift_out);
  } else {
    tail_image_shift_kernel<<<(unsigned)n_envs, 256, 0, (cudaStream_t)cuda_stream>>>(
        p, img, out, replay_shift, shift_out);
  }
  MDPP_CUDA(ctx, cudaGetLastError());
  return MDPP_OK;
}

// 256-layer ziggurat for N(0,1) in fp64: numpy's `random_standard_normal`
// (what `Generator.normal` runs, i.e. the reference's reward noise at
// rl_toy_env.py:1982) fed with Philox words instead of PCG64's.
// oracle/ziggurat.py restates it on the CPU and is pinned to numpy bit for bit.
//
//   r (64 bits): idx = r & 0xff, sign = (r >> 8) & 1, rabs = (r >> 9) & (2^52-1)
//   x = rabs * wi[idx], negated if sign; accepted when rabs < ki[idx] (98.5 %)
//   else: idx == 0 -> tail beyond R (Marsaglia), idx > 0 -> wedge test with a
//   fresh double; on rejection start over with a fresh word.
//
// Word supply (counter based, so a draw depends only on seed, env, step):
//   first attempt of step t : Philox(env, t >> 1, STREAM_ZIG), words (0,1) for
//                             even t, (2,3) for odd t, r = hi:lo
//   everything after it     : the sequence q_0, q_1, ... with
//                             (q_2c, q_2c+1) = Philox(env, t, STREAM_ZIG_RETRY + c)
//                             consumed in numpy's order
// Tables: the fast path reads {wi, (double)ki} pairs (16 B per layer, 4 KB,
// staged in shared memory by the rollout kernel; ki < 2^52 is exact as a
// double, so `rabs < ki` is one fp64 compare -- the fp64 pipe idles on this
// path while the integer ALU is the busiest); the slow path reads the pairs
// and fi from global memory (context.cu uploads them once per context).
#pragma once
#include "philox.cuh"

namespace mdpp {

constexpr int kZigLayers = 256;
constexpr int kZigFastBytes = kZigLayers * 16;
// byte offsets inside the context's ziggurat buffer
constexpr int kZigOffFast = 0;                       // {wi, (double)ki}[256]
constexpr int kZigOffFi = kZigFastBytes;             // fi[256]
constexpr int kZigBytes = kZigFastBytes + kZigLayers * 8;
// Staging of the rollout kernels (discrete_kernels.cuh, zig_fill): normals of
// kZigWindow consecutive steps per thread in shared memory, and per warp a
// queue of the draws whose first attempt was rejected.
#ifdef MDPP_ZIG_WINDOW  // (16 in specialisations whose tables leave no room for 32)
constexpr int kZigWindow = MDPP_ZIG_WINDOW;
#else
constexpr int kZigWindow = 32;
#endif
constexpr int kZigQueueCap = 64;
constexpr int kZigQueueBytes = 16 + 2 * kZigQueueCap;  // count, then u16 items
__host__ __device__ constexpr int zig_stage_bytes(int block, int window = kZigWindow) {
  return window * block * 8 + (block / 32) * kZigQueueBytes;
}

// First attempt.  `kw` -> {wi, (double)ki}[256]; returns x, sets *accepted.
__device__ __forceinline__ double zig_first(uint32_t lo, uint32_t hi, const uint4* kw,
                                            bool* accepted) {
  const uint4 e = kw[lo & 0xffu];  // {wi, (double)ki} as 4 words
  const uint32_t rl = __funnelshift_r(lo, hi, 9);
  const uint32_t rh = (hi >> 9) & 0xFFFFFu;
  // (double)rabs without an integer conversion: 2^52 + rabs has the bit
  // pattern (0x433 << 20 | rh) : rl, and the subtraction is exact
  const double d = __dadd_rn(__hiloint2double((int)(0x43300000u | rh), (int)rl),
                             -4503599627370496.0);
  double x = __dmul_rn(d, __hiloint2double((int)e.y, (int)e.x));
  x = __hiloint2double((int)((uint32_t)__double2hiint(x) ^ ((lo << 23) & 0x80000000u)),
                       __double2loint(x));
  *accepted = d < __hiloint2double((int)e.w, (int)e.z);
  return x;
}

__device__ __forceinline__ double zig_u53(uint64_t w) {  // numpy next_double
  return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

// Everything after a rejected first attempt hi:lo, numpy's loop verbatim.
// Fresh words q_0, q_1, ...: (q_2c, q_2c+1) = Philox(gid, s0, s1, retry_stream + c).
__device__ __forceinline__ double zig_resolve_body(uint32_t lo, uint32_t hi, uint32_t gid,
                                                   uint32_t s0, uint32_t s1,
                                                   uint32_t retry_stream,
                                                   const uint32_t* rk, const uint8_t* zig) {
  const uint4* kw = reinterpret_cast<const uint4*>(zig + kZigOffFast);
  const double* fi = reinterpret_cast<const double*>(zig + kZigOffFi);
  uint32_t call = 0;
  bool have = false;
  U4 q = {0, 0, 0, 0};
  auto next_word = [&]() -> uint64_t {
    if (!have) {
      q = philox4x32_10_rk(gid, s0, s1, retry_stream + call, rk);
      ++call;
      have = true;
      return ((uint64_t)q.y << 32) | q.x;
    }
    have = false;
    return ((uint64_t)q.w << 32) | q.z;
  };
  constexpr double kR = 3.6541528853610087963519472518;
  constexpr double kInvR = 0.27366123732975827203338247596;
  bool first = true;
  for (;;) {
    bool ok;
    const double x = zig_first(lo, hi, kw, &ok);
    if (ok && !first) return x;  // (the first attempt is known to be rejected)
    first = false;
    const uint32_t idx = lo & 0xffu;
    if (idx == 0) {
      for (;;) {
        const double xx = __dmul_rn(-kInvR, log1p(-zig_u53(next_word())));
        const double yy = -log1p(-zig_u53(next_word()));
        if (__dadd_rn(yy, yy) > __dmul_rn(xx, xx)) {
          const uint32_t rl = __funnelshift_r(lo, hi, 9);  // low word of rabs
          return ((rl >> 8) & 1u) ? -__dadd_rn(kR, xx) : __dadd_rn(kR, xx);
        }
      }
    } else {
      const double f1 = fi[idx], f0 = fi[idx - 1];
      const double u = zig_u53(next_word());
      const double lhs = __dadd_rn(__dmul_rn(__dadd_rn(f0, -f1), u), f1);
      if (lhs < exp(__dmul_rn(__dmul_rn(-0.5, x), x))) return x;
    }
    const uint64_t r = next_word();
    lo = (uint32_t)r;
    hi = (uint32_t)(r >> 32);
  }
}

// Everything after a rejected first attempt of (env gid, step) of the discrete
// rollout's reward normal.  Out of line: 1.5 % of the draws get here.  `zig`
// -> a copy of the context's ziggurat buffer (kZigBytes: the rollout kernels
// stage all of it in shared memory).
static __device__ __noinline__ double zig_slow(uint32_t gid, uint64_t step,
                                        const uint32_t* rk, const uint8_t* zig) {
  const uint32_t s0 = (uint32_t)step, s1 = (uint32_t)(step >> 32);
  const uint64_t pair = step >> 1;
  const U4 w = philox4x32_10_rk(gid, (uint32_t)pair, (uint32_t)(pair >> 32),
                                STREAM_ZIG, rk);
  return zig_resolve_body((step & 1) ? w.z : w.x, (step & 1) ? w.w : w.y, gid, s0, s1,
                          STREAM_ZIG_RETRY, rk, zig);
}

// Kernels with several normals per (env, step) -- the continuous transition
// noise, the continuous and grid reward noise.  Draw j of a step: j < 16 =
// state-noise dimension j, first word = (w0,w1) / (w2,w3) of
// Philox(env, step, STREAM_STATE_NOISE + j / 2) for even / odd j; 16 = the
// continuous reward normal, (w0,w1) of Philox(env, step, STREAM_NORMAL); 17 =
// the grid reward normal, the word pair of the step's parity of
// Philox(env, step >> 1, STREAM_GRID_ZIG).  After a rejected first attempt the
// draw continues on the words of STREAM_ZIG_DRAW_RETRY + 64 j + c.
constexpr uint32_t kZigDrawReward = 16, kZigDrawGridReward = 17;

// The rejected draw `draw` of (env gid, step), from scratch (the first word is
// recomputed: the call site passes three scalars and keeps nothing else alive
// for it).  Out of line: 1.5 % of the draws.  `zig` -> the context's buffer.
static __device__ __noinline__ double zig_resolve_draw(uint32_t gid, uint64_t step,
                                                       uint32_t draw, const uint32_t* rk,
                                                       const uint8_t* zig) {
  const uint32_t s0 = (uint32_t)step, s1 = (uint32_t)(step >> 32);
  U4 w;
  bool second;
  if (draw == kZigDrawGridReward) {
    const uint64_t pair = step >> 1;
    w = philox4x32_10_rk(gid, (uint32_t)pair, (uint32_t)(pair >> 32), STREAM_GRID_ZIG, rk);
    second = step & 1;
  } else if (draw == kZigDrawReward) {
    w = philox4x32_10_rk(gid, s0, s1, STREAM_NORMAL, rk);
    second = false;
  } else {
    w = philox4x32_10_rk(gid, s0, s1, STREAM_STATE_NOISE + (draw >> 1), rk);
    second = draw & 1;
  }
  return zig_resolve_body(second ? w.z : w.x, second ? w.w : w.y, gid, s0, s1,
                          STREAM_ZIG_DRAW_RETRY + 64u * draw, rk, zig);
}

}  // namespace mdpp

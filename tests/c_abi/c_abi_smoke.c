/* Pure-C caller of libmdpp_b200.so: no Python, no PyTorch.  Proves that the
 * drop-in boundary is the C ABI of include/mdpp_b200.h.  A hand-written
 * 4-state / 2-action MDP (noise off) is stepped for T steps on N envs with a
 * fixed action pattern and checked against a scalar loop in this file.
 *   gcc c_abi_smoke.c -I../../include -I/usr/local/cuda/include \
 *       -L<dir of .so> -lmdpp_b200 -L/usr/local/cuda/lib64 -lcudart -o smoke */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mdpp_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define MK(x) do { int rc_ = (x); if (rc_ != 0) { \
  printf("mdpp error %d: %s\n", rc_, mdpp_last_error(ctx)); return 3; } } while (0)

enum { S = 4, A = 2, L = 2, D = 1, N = 1000, T = 37 };

int main(void) {
  /* P[s][a]; state 3 is terminal (self loop); sequence (0,1) pays 1.0, (1,2) pays 0.5 */
  int32_t P[S * A] = {1, 2, 2, 0, 3, 1, 3, 3};
  uint8_t term[S] = {0, 0, 0, 1};
  double init_cdf[S] = {1.0 / 3, 2.0 / 3, 1.0, 1.0};
  int32_t seqs[2 * L] = {0, 1, 1, 2};
  double seq_rewards[2] = {1.0, 0.5};
  mdpp_ctx* ctx = NULL;
  if (mdpp_abi_version() != MDPP_ABI_VERSION) { printf("ABI mismatch\n"); return 1; }
  MK(mdpp_create(0, &ctx));
  mdpp_discrete_group g;
  memset(&g, 0, sizeof g);
  g.n_states = S; g.n_actions = A; g.sequence_length = L; g.delay = D;
  g.reward_every_n_steps = 1; g.n_sequences = 2;
  g.reward_scale = 2.0; g.reward_shift = -0.25; g.term_state_reward = 10.0;
  g.transition = P; g.terminal = term; g.init_cdf = init_cdf;
  g.sequences = seqs; g.sequence_rewards = seq_rewards;
  g.env_begin = 0; g.env_count = N;
  MK(mdpp_set_discrete_groups(ctx, &g, 1));

  mdpp_discrete_state st;
  memset(&st, 0, sizeof st);
  st.n_envs = N; st.ring_depth = D;
  CK(cudaMalloc((void**)&st.cur_state, N * 4)); CK(cudaMalloc((void**)&st.seq_key, N * 8));
  CK(cudaMalloc((void**)&st.t_episode, N * 4)); CK(cudaMalloc((void**)&st.episode, N * 4));
  CK(cudaMalloc((void**)&st.ring, D * N * 8)); CK(cudaMalloc((void**)&st.stats, MDPP_N_STATS * 8));
  CK(cudaMemset(st.t_episode, 0, N * 4)); CK(cudaMemset(st.episode, 0, N * 4));
  CK(cudaMemset(st.ring, 0, D * N * 8)); CK(cudaMemset(st.stats, 0, MDPP_N_STATS * 8));

  /* reset every env to a chosen initial state (i % 3) */
  int32_t* h_init = (int32_t*)malloc(N * 4);
  for (int i = 0; i < N; ++i) h_init[i] = i % 3;
  int32_t* d_init; CK(cudaMalloc((void**)&d_init, N * 4));
  CK(cudaMemcpy(d_init, h_init, N * 4, cudaMemcpyHostToDevice));
  mdpp_step_opts opts;
  memset(&opts, 0, sizeof opts);
  opts.n_steps = 1; opts.noise_mode = MDPP_NOISE_OFF;
  MK(mdpp_discrete_reset(ctx, &st, NULL, d_init, NULL, NULL, &opts, NULL));

  int32_t* h_act = (int32_t*)malloc((size_t)T * N * 4);
  for (int t = 0; t < T; ++t) for (int i = 0; i < N; ++i) h_act[t * N + i] = (t * 7 + i * 3) % 5 < 2;
  mdpp_discrete_io io;
  memset(&io, 0, sizeof io);
  int32_t* d_act; CK(cudaMalloc((void**)&d_act, (size_t)T * N * 4));
  CK(cudaMemcpy(d_act, h_act, (size_t)T * N * 4, cudaMemcpyHostToDevice));
  io.actions = d_act;
  CK(cudaMalloc((void**)&io.obs, (size_t)T * N * 8)); CK(cudaMalloc((void**)&io.reward, (size_t)T * N * 8));
  CK(cudaMalloc((void**)&io.terminated, (size_t)T * N)); CK(cudaMalloc((void**)&io.truncated, (size_t)T * N));
  opts.n_steps = T;
  MK(mdpp_discrete_rollout(ctx, &st, &io, &opts, NULL));
  CK(cudaDeviceSynchronize());

  int64_t* obs = (int64_t*)malloc((size_t)T * N * 8);
  double* rew = (double*)malloc((size_t)T * N * 8);
  uint8_t* ter = (uint8_t*)malloc((size_t)T * N);
  CK(cudaMemcpy(obs, io.obs, (size_t)T * N * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rew, io.reward, (size_t)T * N * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ter, io.terminated, (size_t)T * N, cudaMemcpyDeviceToHost));

  /* scalar restatement (rl_toy_env.py:1602-1603, :1837-1841, :1970-1990, :2102-2109) */
  long bad = 0;
  for (int i = 0; i < N; ++i) {
    int s = h_init[i], prev = -1, tl = 0;
    double fifo = 0.0;
    for (int t = 0; t < T; ++t) {
      int nxt = P[s * A + h_act[t * N + i]];
      ++tl;
      double r = 0.0;
      if (tl >= L) { if (s == 0 && nxt == 1) r = 1.0; else if (s == 1 && nxt == 2) r = 0.5; }
      (void)prev;
      double delayed = fifo; fifo = r; r = delayed;  /* delay 1, zero-initialised */
      r = r * 2.0 + -0.25;
      int done = term[nxt];
      if (done) r += 10.0 * 2.0;
      if (obs[t * N + i] != nxt || rew[t * N + i] != r || ter[t * N + i] != done) ++bad;
      prev = s; s = nxt;
    }
  }
  double stats[MDPP_N_STATS];
  CK(cudaMemcpy(stats, st.stats, sizeof stats, cudaMemcpyDeviceToHost));
  printf("c_abi_smoke: %d envs x %d steps, mismatches=%ld, transitions=%.0f\n", N, T, bad,
         stats[MDPP_STAT_TRANSITIONS]);
  /* error convention: NULL io must be refused with a message, not crash */
  int rc = mdpp_discrete_rollout(ctx, &st, NULL, &opts, NULL);
  printf("expected failure rc=%d msg=%s\n", rc, mdpp_last_error(ctx));
  mdpp_destroy(ctx);
  return (bad == 0 && stats[MDPP_STAT_TRANSITIONS] == (double)N * T && rc == MDPP_EINVAL) ? 0 : 1;
}

// Internal declarations shared by the translation units of libmdpp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mdpp_b200.h"

namespace mdpp {

// Device-side descriptor of one discrete configuration group.  The tables of
// the group live in one contiguous, 16-byte aligned blob so that a CTA can
// stage them into shared memory with 128-bit copies.
struct DiscreteGroupDev {
  int32_t S, A, L, delay, every_n, custom_reward, n_seq, key_bits;
  int32_t has_pnoise, has_rnoise, lookup_kind, hash_shift;  // kind: 0 LUT 1 hash 2 R
  uint32_t hash_mask;
  int32_t cdf_log2;    // cdf rows hold 2^cdf_log2 entries (sentinel padded)
  int32_t cdf_stride;  // = 2^cdf_log2
  int32_t has_guide;
  uint64_t key_mask;
  double p_noise, r_std, scale, shift, term_reward_scaled;
  // byte offsets inside the blob
  int32_t off_P, off_term, off_init_cdf, off_noise_cdf;
  int32_t off_lut, off_hash_keys, off_hash_vals, off_values, off_R, off_guide;
  int32_t pad1;
  int32_t blob_bytes;
  int64_t blob_offset;  // of this group's blob inside the context blob buffer
  int64_t env_begin, env_count;
};

struct CtaMapEntry {
  int32_t group;
  int32_t chunk;  // chunk index inside the group (units of block size)
};

enum { LOOKUP_LUT = 0, LOOKUP_HASH = 1, LOOKUP_MATRIX = 2 };
constexpr int kLutMaxBits = 12;
constexpr uint64_t kHashEmpty = ~0ull;
constexpr int kGuideBits = 12;
constexpr int kGuideEntries = 1 << kGuideBits;
constexpr uint8_t kGuideMiss = 0xFF;

}  // namespace mdpp

struct mdpp_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  std::string error;
  // discrete tables
  std::vector<mdpp::DiscreteGroupDev> d_groups_host;
  mdpp::DiscreteGroupDev* d_groups = nullptr;
  uint8_t* d_blob = nullptr;
  size_t d_blob_bytes = 0;
  int max_group_blob = 0;
  int max_delay = 0;
  int64_t d_total_envs = 0;
  // CTA -> (group, chunk) map, one per supported block size
  mdpp::CtaMapEntry* d_cta_map = nullptr;
  int cta_map_block = 0;
  int64_t n_ctas = 0;
};

namespace mdpp {
int fail(mdpp_ctx* ctx, int code, const std::string& msg);
int cuda_fail(mdpp_ctx* ctx, cudaError_t e, const char* what);
}  // namespace mdpp

#define MDPP_CUDA(ctx, call)                                      \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return mdpp::cuda_fail(ctx, e_, #call); \
  } while (0)

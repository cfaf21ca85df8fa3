// Image-observation renderers (K4 ImageMultiDiscrete, K5 ImageContinuous).
//
// K4 restates spaces/image_multi_discrete.py:129-270: polygon -> rotate ->
// flip -> transpose, as ONE gather per output byte: the final pixel is mapped
// back through the flip and Pillow's 16.16 fixed-point rotation to the polygon
// image, where a bit of the (state, R, vertex-variant) mask decides 0 / 255.
// K5 restates spaces/image_continuous.py:116-246: background 208, black
// terminal rectangles, green target disc, blue agent disc (painter's order),
// relevant and irrelevant sub-images stacked along x.
//
// One CTA renders one image; a thread produces 16 consecutive output bytes and
// writes them with one 128-bit streaming store.  Chunks outside the shape's
// bounding box are emitted without per-pixel work (most of the image is
// background), so the kernel stays store-bandwidth bound.
#include "internal.h"
#include "philox.cuh"

namespace mdpp {

constexpr int kRBlock = 128;
constexpr int kMaskRows = 64;
constexpr int kMaskCentre = 31;
constexpr int kRotCols = 80;  // bounding box of the rotated polygon: <= 2*30+7+2

struct RenderDParams {
  mdpp_image_discrete_tables tb;
  const int64_t* states;
  const int32_t* params_in;
  int32_t* params_out;
  uint8_t* out;
  int64_t n_envs;
  uint32_t k0, k1, stream;
  uint64_t step_index;
  const uint64_t* step_index_dev;
  int64_t env_id_offset;
};

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// 12 CTAs per SM (40 registers): the kernel is short, occupancy hides the
// per-image prologue latency
__global__ void __launch_bounds__(kRBlock, 12)
render_discrete_kernel(const __grid_constant__ RenderDParams p) {
  __shared__ uint64_t mask[kMaskRows];   // column bitmaps of the polygon
  __shared__ uint64_t rmask[kRotCols][2]; // rotated + flipped polygon, box-local
  __shared__ int prm[8];
  const int64_t m = blockIdx.x;
  const mdpp_image_discrete_tables& tb = p.tb;
  const int W = tb.width, H = tb.height;
  if (threadIdx.x < 32) {  // warp 0 decides the transform parameters
    const int lane = threadIdx.x;
    int state = (int)p.states[m];
    state = min(max(state, 0), tb.n_states - 1);
    int R, sw, sh, rot, flip;
    if (p.params_in) {
      const int32_t* q = p.params_in + m * 5;
      R = q[0]; sw = q[1]; sh = q[2]; rot = q[3]; flip = q[4];
    } else {
      // draw order of the reference (:149-181, :251, :258-259); one Philox
      // call per image: w0 scale, w1 / w2 shifts, w3 rotation + flip bits
      const uint32_t gid = (uint32_t)(p.env_id_offset + m % p.n_envs);
      const uint64_t step = p.step_index + (uint64_t)(m / p.n_envs) +
                            (p.step_index_dev ? *p.step_index_dev : 0ull);
      U4 w = philox4x32_10(gid, (uint32_t)step, (uint32_t)(step >> 32), p.stream,
                           p.k0, p.k1);
      R = tb.r_min;
      if (tb.has_scale) {  // R = r_min + #{thresholds <= u}, one lane each
        const double u = uniform32(w.x);
        for (int base = 0; base < tb.n_radii - 1; base += 32) {
          const int k = base + lane;
          const bool le = k < tb.n_radii - 1 && tb.r_thresholds[k] <= u;
          R += __popc(__ballot_sync(0xffffffffu, le));
        }
      }
      sw = W / 2; sh = H / 2;
      if (tb.has_shift) {  // integers(-m + 1, m), m = W/2 - R, then quantise
        const int mw = W / 2 - R, mh = H / 2 - R;
        const int aw = -mw + 1 + (int)__umulhi(w.y, (uint32_t)max(2 * mw - 1, 1));
        const int ah = -mh + 1 + (int)__umulhi(w.z, (uint32_t)max(2 * mh - 1, 1));
        sw += floor_div(aw, tb.sh_quant) * tb.sh_quant;
        sh += floor_div(ah, tb.sh_quant) * tb.sh_quant;
      }
      rot = -1;
      if (tb.has_rotate)
        rot = ((int)__umulhi(w.w, 360u) / tb.ro_quant) * tb.ro_quant;
      flip = 0;
      if (tb.has_flip && (w.w & 1u) == 0) flip = (w.w & 2u) == 0 ? 1 : 2;
    }
    if (lane == 0) {
      prm[0] = state; prm[1] = R; prm[2] = sw; prm[3] = sh; prm[4] = rot;
      prm[5] = flip;
      if (p.params_out) {
        int32_t* q = p.params_out + m * 5;
        q[0] = R; q[1] = sw; q[2] = sh; q[3] = rot; q[4] = flip;
      }
    }
  }
  __syncthreads();
  const int state = prm[0], R = prm[1], sw = prm[2], sh = prm[3], rot = prm[4],
            flip = prm[5];
  if (threadIdx.x < kMaskRows) {
    const int ri = min(max(R - tb.r_min, 0), tb.n_radii - 1);
    const int cell = state * tb.n_radii + ri;
    const int xv = tb.xvar[(int64_t)cell * W + min(max(sw, 0), W - 1)];
    const int yv = tb.yvar[(int64_t)cell * H + min(max(sh, 0), H - 1)];
    const int id = tb.mask_index[((int64_t)cell * tb.n_xvar + xv) * tb.n_yvar + yv];
    uint64_t col = tb.mask_bits[(int64_t)id * kMaskRows + threadIdx.x];
    // A quantised shift can push the polygon up to q-1 pixels over the image
    // edge (Pillow clips it there): clear the mask bits whose pixel lies
    // outside the image, so "inside the mask" implies "inside the image".
    const int rx = threadIdx.x + sw - kMaskCentre;
    const int lo = max(0, kMaskCentre - sh), hi = min(63, H - 1 + kMaskCentre - sh);
    uint64_t rows = 0;
    if (hi >= lo) rows = (hi - lo == 63 ? ~0ull : ((1ull << (hi - lo + 1)) - 1ull)) << lo;
    if ((unsigned)rx >= (unsigned)W) rows = 0;
    mask[threadIdx.x] = col & rows;
  }
  int a0 = 65536, a1 = 0, a2 = 0, a3 = 0, a4 = 65536, a5 = 0;
  if (rot >= 0) {
    const int32_t* c = tb.rot_coeff + (rot % 360) * 6;
    a0 = c[0]; a1 = c[1]; a2 = c[2]; a3 = c[3]; a4 = c[4]; a5 = c[5];
  }
  // bounding box of the polygon (disc of radius R around the centre) in the
  // FINAL image: forward-map the centre through the rotation and the flip
  float cx = (float)sw, cy = (float)sh;
  if (rot >= 0) {
    const float X = (float)sw * 65536.f - (float)a2, Y = (float)sh * 65536.f - (float)a5;
    const float det = (float)a0 * (float)a4 - (float)a1 * (float)a3;
    cx = ((float)a4 * X - (float)a1 * Y) / det;
    cy = ((float)a0 * Y - (float)a3 * X) / det;
  }
  if (flip == 1) cx = (float)(W - 1) - cx;
  if (flip == 2) cy = (float)(H - 1) - cy;
  const int bx0 = (int)floorf(cx) - R - 3, bx1 = (int)ceilf(cx) + R + 3;
  const int by0 = (int)floorf(cy) - R - 3, by1 = (int)ceilf(cy) + R + 3;
  __syncthreads();
  // 4 mask bits -> 4 bytes of 0 / 255: bit k of b lands on bit 8k of
  // b * (1 + 2^7 + 2^14 + 2^21); the isolated 0/1 bytes times 255 fill up
  auto expand4 = [](uint32_t b) -> uint32_t {
    return (((b & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
  };

  // Unrotated images: 16 consecutive output bytes = 16 consecutive y at one x
  // = 16 consecutive bits of one column bitmap (reversed under a top-bottom
  // flip).  n pixels of column x starting at y, as a bit string (bit j = y+j).
  auto column_bits = [&](int x, int y, int n) -> uint32_t {
    const int fx = flip == 1 ? W - 1 - x : x;
    const int mx = fx - sw + kMaskCentre;
    if ((unsigned)mx >= (unsigned)kMaskRows) return 0u;
    uint64_t col = mask[mx];
    int start;  // mask bit of pixel y, ascending with y after the optional flip
    if (flip == 2) {
      col = __brevll(col);                       // bit k <- bit 63-k
      start = 63 - ((H - 1 - y) - sh + kMaskCentre);
    } else {
      start = y - sh + kMaskCentre;
    }
    uint64_t v;
    if (start >= 64 || start <= -64) v = 0;
    else v = start >= 0 ? col >> start : col << (-start);
    return (uint32_t)v & ((1u << n) - 1u);
  };
  auto pixel = [&](int x, int y) -> uint32_t {  // rotated images: one gather
    int fx = x, fy = y;
    if (flip == 1) fx = W - 1 - x;
    if (flip == 2) fy = H - 1 - y;
    const int rx = (a2 + a1 * fy + a0 * fx) >> 16;
    const int ry = (a5 + a4 * fy + a3 * fx) >> 16;
    if ((unsigned)rx >= (unsigned)W || (unsigned)ry >= (unsigned)H) return 0u;
    const int mx = rx - sw + kMaskCentre, my = ry - sh + kMaskCentre;
    if ((unsigned)mx >= (unsigned)kMaskRows || (unsigned)my >= 64u) return 0u;
    return (uint32_t)((mask[mx] >> my) & 1ull);
  };

  // Rotated images: gather the polygon ONCE into box-local column bitmaps of
  // the final image (two threads per column, the inverse affine stepped
  // incrementally along y); the output pass then reads bit strings exactly
  // like the unrotated case.  Bit k of column c is final pixel
  // (rcx0 + c, rcy0 + k).
  const int rcx0 = max(bx0, 0), rcx1 = min(min(bx1, W - 1), rcx0 + kRotCols - 1);
  const int rcy0 = max(by0, 0), rcy1 = min(min(by1, H - 1), rcy0 + 127);
  if (rot >= 0) {
    const int ncols = rcx1 - rcx0 + 1, nrows = rcy1 - rcy0 + 1;
    for (int w = threadIdx.x; w < 2 * kRotCols; w += kRBlock) (&rmask[0][0])[w] = 0ull;
    __syncthreads();
    // work item = 8 consecutive rows of one column (keeps all lanes busy);
    // items of a column OR their bits into the column's two words
    const int segs = (nrows + 7) >> 3;
    for (int w = threadIdx.x; w < ncols * segs; w += kRBlock) {
      const int c = w / segs, sgm = w - c * segs;
      const int x = rcx0 + c;
      const int fx = flip == 1 ? W - 1 - x : x;
      const int k0 = sgm * 8;                               // box-local rows
      const int sy = flip == 2 ? -1 : 1;                    // d(fy) / d(y)
      const int fy = flip == 2 ? H - 1 - (rcy0 + k0) : rcy0 + k0;
      // 16.16 source coordinates, pre-shifted into mask space (mx = X >> 16,
      // my = Y >> 16).  Mask bits outside the image were cleared at load, so
      // the image-bounds test of the rotation is implied by the mask-bounds
      // test -- one compare on (mx | my).
      int X = a2 + a1 * fy + a0 * fx + (kMaskCentre - sw) * 65536;
      int Y = a5 + a4 * fy + a3 * fx + (kMaskCentre - sh) * 65536;
      const int dX = a1 * sy, dY = a4 * sy;
      uint32_t bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // bit k enters at the top and slides down
        const int mx = X >> 16, my = Y >> 16;
        uint32_t v = 0;
        if ((unsigned)(mx | my) < 64u) v = (uint32_t)(mask[mx] >> my);
        bits = __funnelshift_r(bits, v, 1);
        X += dX; Y += dY;
      }
      bits >>= 24;
      if (k0 + 8 > nrows) bits &= (1u << (nrows - k0)) - 1u;
      if (bits)
        atomicOr(reinterpret_cast<unsigned long long*>(&rmask[c][k0 >> 6]),
                 (unsigned long long)bits << (k0 & 63));
    }
    __syncthreads();
  }
  // n (<= 16) pixels of final column x starting at y, from the rotated bitmaps
  auto rot_bits = [&](int x, int y, int n) -> uint32_t {
    const int c = x - rcx0;
    if ((unsigned)c > (unsigned)(rcx1 - rcx0)) return 0u;
    const uint64_t lo = rmask[c][0], hi = rmask[c][1];
    const int k = y - rcy0;  // bit of pixel y
    uint64_t v;
    if (k <= -64 || k >= 128) v = 0;
    else if (k < 0) v = lo << (-k);
    else if (k == 0) v = lo;
    else if (k < 64) v = (lo >> k) | (hi << (64 - k));
    else v = hi >> (k - 64);
    return (uint32_t)v & ((1u << n) - 1u);
  };

  const int total = W * H;
  uint8_t* out = p.out + m * (int64_t)total;
  if (total % 16 == 0) {  // every image starts 16-byte aligned
    uint4* out4 = reinterpret_cast<uint4*>(out);
    // phase 1: the image is mostly background -- stream zeros everywhere
    for (int i = threadIdx.x; i < total / 16; i += kRBlock)
      __stcs(out4 + i, make_uint4(0u, 0u, 0u, 0u));
    // phase 2: recompute only the 16-byte chunks that meet the shape's
    // bounding box (after the barrier, so these stores land last)
    const int cx0 = max(bx0, 0), cx1 = min(bx1, W - 1);
    const int cy0 = max(by0, 0), cy1 = min(by1, H - 1);
    __syncthreads();
    if (cx0 > cx1 || cy0 > cy1) return;
    // chunks a column span can touch, rounded up to a power of two so the
    // (column, k) decomposition of a work item is a shift and a mask
    const int per_col = ((cy1 - cy0) >> 4) + 2;
    const int lg = per_col <= 4 ? 2 : per_col <= 8 ? 3 : per_col <= 16 ? 4 : 5;
    const int n_work = (cx1 - cx0 + 1) << lg;
    for (int wi = threadIdx.x; wi < n_work; wi += kRBlock) {
      const int col = cx0 + (wi >> lg), k = wi & ((1 << lg) - 1);
      const int chunk = ((col * H + cy0) >> 4) + k;
      if (chunk > ((col * H + cy1) >> 4)) continue;
      const int idx0 = chunk * 16;
      const int x0 = idx0 >= col * H ? col : col - 1;
      const int y0 = idx0 - x0 * H;
      const int n0 = min(16, H - y0);  // pixels of the chunk in column x0
      uint32_t bits = 0;
      if (rot < 0) {
        bits = column_bits(x0, y0, n0);
        if (n0 < 16 && x0 + 1 < W) bits |= column_bits(x0 + 1, 0, 16 - n0) << n0;
      } else {
        bits = rot_bits(x0, y0, n0);
        if (n0 < 16 && x0 + 1 < W) bits |= rot_bits(x0 + 1, 0, 16 - n0) << n0;
      }
      __stcs(out4 + chunk, make_uint4(expand4(bits), expand4(bits >> 4),
                                      expand4(bits >> 8), expand4(bits >> 12)));
    }
  } else {
    for (int idx = threadIdx.x; idx < total; idx += kRBlock) {
      const int x = idx / H, y = idx % H;
      out[idx] = (rot < 0 ? column_bits(x, y, 1) : pixel(x, y)) ? 255 : 0;
    }
  }
}

struct RenderCParams {
  mdpp_image_continuous_config cfg;
  const void* states;
  uint8_t* out;
};

template <typename R>
__global__ void __launch_bounds__(kRBlock)
render_continuous_kernel(const __grid_constant__ RenderCParams p) {
  const mdpp_image_continuous_config& c = p.cfg;
  const int W = c.width, H = c.height;
  const int64_t m = blockIdx.x / c.n_sub_images;
  const int sub = blockIdx.x % c.n_sub_images;  // 0 relevant, 1 irrelevant
  const R* st = reinterpret_cast<const R*>(p.states) + m * c.dim;
  // convert_to_pixel (:248-277): ((v - lo) / (hi - lo)) in dtype_s, promoted
  // to float64 by the integer image shape, truncated
  int px[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const R v = st[sub == 0 ? c.rel_index[k] : c.irr_index[k]];
    const R lo = (R)c.feat_low[k], hi = (R)c.feat_high[k];
    R frac;
    if (sizeof(R) == 4) frac = (R)__fdiv_rn(__fsub_rn((float)v, (float)lo),
                                           __fsub_rn((float)hi, (float)lo));
    else frac = (R)__ddiv_rn(__dsub_rn((double)v, (double)lo),
                             __dsub_rn((double)hi, (double)lo));
    px[k] = (int)__dmul_rn((double)frac, (double)(k == 0 ? W : H));
  }
  const int rad = c.stamp_radius;
  const bool rel = sub == 0;
  auto in_stamp = [&](int x, int y, int cx, int cy) -> bool {
    const int r = y - cy + rad;
    if ((unsigned)r >= (unsigned)c.stamp_rows) return false;
    const int dx = x - cx - c.stamp[r][0];
    return (unsigned)dx < (unsigned)c.stamp[r][1];
  };
  auto colour = [&](int x, int y, int ch) -> uint32_t {
    if (in_stamp(x, y, px[0], px[1])) return ch == 2 ? 255u : 0u;  // agent, blue
    if (rel) {
      if (c.has_target && in_stamp(x, y, c.target_pixel[0], c.target_pixel[1]))
        return ch == 1 ? 255u : 0u;                                 // target, green
      for (int b = 0; b < c.n_rects; ++b)
        if (x >= c.rect[b][0] && x <= c.rect[b][2] && y >= c.rect[b][1] &&
            y <= c.rect[b][3])
          return 0u;                                                // terminal, black
    }
    return 208u;
  };
  const int total = W * H * 3;
  uint8_t* out = p.out + ((int64_t)m * c.n_sub_images + sub) * (int64_t)total;
  if (total % 16 == 0) {
    uint4* out4 = reinterpret_cast<uint4*>(out);
    for (int i = threadIdx.x; i < total / 16; i += kRBlock) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      int pix = (i * 16) / 3, ch = (i * 16) % 3;
      int x = pix / H, y = pix - x * H;
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        w[b >> 2] |= colour(x, y, ch) << ((b & 3) * 8);
        if (++ch == 3) { ch = 0; if (++y == H) { y = 0; ++x; } }
      }
      __stcs(out4 + i, make_uint4(w[0], w[1], w[2], w[3]));
    }
  } else {
    for (int idx = threadIdx.x; idx < total; idx += kRBlock) {
      const int pix = idx / 3;
      out[idx] = (uint8_t)colour(pix / H, pix % H, idx % 3);
    }
  }
}

}  // namespace mdpp

using namespace mdpp;

extern "C" int mdpp_render_discrete(mdpp_ctx* ctx,
                                    const mdpp_image_discrete_tables* tb,
                                    const int64_t* states,
                                    const int32_t* params_in,
                                    int32_t* params_out, uint8_t* out,
                                    int64_t n_images, int64_t n_envs,
                                    int32_t image_stream,
                                    const mdpp_step_opts* opts,
                                    void* cuda_stream) {
  if (!ctx) return MDPP_EINVAL;
  if (!tb || !states || !out || !opts || n_images < 1 || n_envs < 1)
    return fail(ctx, MDPP_EINVAL, "render_discrete: bad arguments");
  if (tb->width < 1 || tb->height < 1 || !tb->mask_bits || !tb->mask_index ||
      !tb->xvar || !tb->yvar || !tb->rot_coeff)
    return fail(ctx, MDPP_EINVAL, "render_discrete: incomplete tables");
  if (tb->has_scale && tb->n_radii > 1 && !tb->r_thresholds)
    return fail(ctx, MDPP_EINVAL, "render_discrete: missing r_thresholds");
  if (n_images > 0x7fffffffLL)
    return fail(ctx, MDPP_EINVAL, "render_discrete: too many images per call");
  MDPP_CUDA(ctx, cudaSetDevice(ctx->device));
  RenderDParams p;
  p.tb = *tb;
  p.states = states;
  p.params_in = params_in;
  p.params_out = params_out;
  p.out = out;
  p.n_envs = n_envs;
  p.k0 = (uint32_t)opts->seed;
  p.k1 = (uint32_t)(opts->seed >> 32);
  p.stream = (uint32_t)image_stream;
  p.step_index = opts->step_index;
  p.step_index_dev = opts->step_index_dev;
  p.env_id_offset = opts->env_id_offset;
  render_discrete_kernel<<<(unsigned)n_images, kRBlock, 0,
                           (cudaStream_t)cuda_stream>>>(p);
  MDPP_CUDA(ctx, cudaGetLastError());
  return MDPP_OK;
}

extern "C" int mdpp_render_continuous(mdpp_ctx* ctx,
                                      const mdpp_image_continuous_config* cfg,
                                      const void* states, uint8_t* out,
                                      int64_t n_images, void* cuda_stream) {
  if (!ctx) return MDPP_EINVAL;
  if (!cfg || !states || !out || n_images < 1)
    return fail(ctx, MDPP_EINVAL, "render_continuous: bad arguments");
  if (cfg->n_sub_images < 1 || cfg->n_sub_images > 2 ||
      cfg->stamp_rows > MDPP_MAX_STAMP_ROWS || cfg->n_rects > MDPP_MAX_TERM_BOXES)
    return fail(ctx, MDPP_EINVAL, "render_continuous: bad configuration");
  if (n_images * cfg->n_sub_images > 0x7fffffffLL)
    return fail(ctx, MDPP_EINVAL, "render_continuous: too many images per call");
  MDPP_CUDA(ctx, cudaSetDevice(ctx->device));
  RenderCParams p;
  p.cfg = *cfg;
  p.states = states;
  p.out = out;
  const unsigned grid = (unsigned)(n_images * cfg->n_sub_images);
  if (cfg->is_f64)
    render_continuous_kernel<double><<<grid, kRBlock, 0, (cudaStream_t)cuda_stream>>>(p);
  else
    render_continuous_kernel<float><<<grid, kRBlock, 0, (cudaStream_t)cuda_stream>>>(p);
  MDPP_CUDA(ctx, cudaGetLastError());
  return MDPP_OK;
}

// HBM-bound helper kernels of the Wan VAE decoder on channels-last activations [pixels, C] bf16:
//   b200_rmsnorm_silu_cl   WanRMS_norm over channels (+ SiLU)        vae/wan/model.py:216-222, :404-405
//   b200_upsample2x_cl     nearest-exact 2x spatial upsample          vae/wan/model.py:226-237 (WanUpsample)
//   b200_softmax_rows      row softmax of fp32 scores -> bf16 probs   mid-block attention, vae/wan/model.py:478
//   b200_blend_tile        tile blend + crop + clamp                  vae/wan/model.py:1404-1422, 1600-1619
#include "host_util.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace vae {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// One pixel per SEG-lane segment of a warp (SEG = 16 or 32); lane i of the segment owns 16-byte chunks
// i, i + SEG, ...  y = x / max(||x||_2, 1e-12) * sqrt(C) * gamma ; optional SiLU.  fp32 math, one rounding.
// SiLU is evaluated as 0.5 v (1 + tanh(v / 2)) with tanh.approx: the exp + divide form costs two MUFU ops per element and
// made these kernels SFU-bound at 38 % of the HBM bandwidth (profiles/r01_row_kernels_bw.json).
template <int SEG, int ITERS>
__global__ void __launch_bounds__(256)
rmsnorm_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                    const __nv_bfloat16* __restrict__ gamma, int64_t pixels, int C, int silu) {
  const int lane = threadIdx.x & 31;
  const int seg_lane = lane % SEG;
  const int segs_per_warp = 32 / SEG;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t pix = warp_global * segs_per_warp + lane / SEG;
  const int nchunks = C >> 3;
  const bool active = pix < pixels;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (active ? pix : 0) * C);
  uint4 raw[ITERS];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = seg_lane + i * SEG;
    if (active && c < nchunks) {
      raw[i] = xr[c];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
#pragma unroll
  for (int o = SEG / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = sqrtf(static_cast<float>(C)) / fmaxf(sqrtf(ss), 1e-12f);
  uint4* yr = reinterpret_cast<uint4*>(y + (active ? pix : 0) * C);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = seg_lane + i * SEG;
    if (active && c < nchunks) {
      float f[8], g[8];
      unpack8(raw[i], f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + c), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = f[j] * inv * g[j];
        if (silu) v = 0.5f * v * (1.0f + fast_tanh(0.5f * v));  // silu = v * sigmoid(v), ONE MUFU op (tanh.approx) per element
        f[j] = v;
      }
      yr[c] = pack8(f);
    }
  }
}

// out[t, y, x, :] = in[t, y / 2, x / 2, :]   (nearest-exact with scale 2)
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int H, int W, int chunks) {
  const int64_t total = static_cast<int64_t>(T) * (2 * H) * (2 * W) * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % chunks);
    int64_t r = i / chunks;
    const int xo = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int yo = static_cast<int>(r % (2 * H));
    const int t = static_cast<int>(r / (2 * H));
    out[i] = in[((static_cast<int64_t>(t) * H + (yo >> 1)) * W + (xo >> 1)) * chunks + c];
  }
}

// One warp per row: P = softmax(S * scale) over `cols` fp32 scores, bf16 out (row stride ldp).
// `block` > 0: block-causal mask -- row r sees the columns [0, (r / block + 1) * block) (frame-causal attention of the
// HunyuanVideo-1.5 VAE mid block, vae/hunyuanvideo15/model.py:143-165, block = H*W); masked probabilities are written as 0.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int rows, int cols_all, int64_t lds,
                    int64_t ldp, float scale, int block) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int cols = block > 0 ? min(cols_all, (static_cast<int>(row) / block + 1) * block) : cols_all;
  const float* sr = s + row * lds;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += __expf((sr[c] - mx) * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + row * ldp;
  for (int c = lane; c < cols; c += 32) pr[c] = __float2bfloat16(__expf((sr[c] - mx) * scale) * inv);
  for (int c = cols + lane; c < cols_all; c += 32) pr[c] = __float2bfloat16(0.f);
}

// Blend one decoded tile [C, T, th, tw] (planar bf16) with its upper / left neighbours and write the cropped,
// clamped result into the frame buffer [C, T, OH, OW] at (y0, x0).  The reference blends IN PLACE in row-major
// tile order (vae/wan/model.py:1600-1614): the upper neighbour `up` has already been blended with ITS upper and
// left neighbours, the left neighbour likewise -- so tiles are processed in that order and `tile` is updated in
// place before being consumed by later tiles.
__global__ void __launch_bounds__(256)
blend_tile_kernel(__nv_bfloat16* __restrict__ tile, const __nv_bfloat16* __restrict__ up,
                  const __nv_bfloat16* __restrict__ left, __nv_bfloat16* __restrict__ frame, int planes, int th, int tw,
                  int up_h, int up_w, int left_h, int left_w, int blend, int crop_h, int crop_w, int y0, int x0, int OH,
                  int OW, int clamp) {
  const int64_t total = static_cast<int64_t>(planes) * th * tw;
  const int bv = up ? min(min(up_h, th), blend) : 0;
  const int bh = left ? min(min(left_w, tw), blend) : 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % tw);
    const int y = static_cast<int>((i / tw) % th);
    const int64_t pl = i / (static_cast<int64_t>(tw) * th);
    float v = __bfloat162float(tile[i]);
    if (y < bv && x < up_w) {
      const float a = __bfloat162float(up[(pl * up_h + (up_h - bv + y)) * up_w + x]);
      // a * (1 - y / e) + b * (y / e): python floats multiply a bf16 tensor -> each product / the sum rounds to bf16
      const float w1 = static_cast<float>(1.0 - static_cast<double>(y) / bv), w2 = static_cast<float>(static_cast<double>(y) / bv);
      const float t1 = __bfloat162float(__float2bfloat16(a * w1));
      const float t2 = __bfloat162float(__float2bfloat16(v * w2));
      v = __bfloat162float(__float2bfloat16(t1 + t2));
    }
    if (x < bh && y < left_h) {
      const float a = __bfloat162float(left[(pl * left_h + y) * left_w + (left_w - bh + x)]);
      const float w1 = static_cast<float>(1.0 - static_cast<double>(x) / bh), w2 = static_cast<float>(static_cast<double>(x) / bh);
      const float t1 = __bfloat162float(__float2bfloat16(a * w1));
      const float t2 = __bfloat162float(__float2bfloat16(v * w2));
      v = __bfloat162float(__float2bfloat16(t1 + t2));
    }
    tile[i] = __float2bfloat16(v);
    if (y < crop_h && x < crop_w && y0 + y < OH && x0 + x < OW) {
      const float c = clamp ? fminf(fmaxf(v, -1.0f), 1.0f) : v;
      frame[(pl * OH + (y0 + y)) * OW + (x0 + x)] = __float2bfloat16(c);
    }
  }
}

// Frame hand-off: planar bf16 video [3, T, H, W] in [-1, 1] -> uint8 [T, H, W, 3], with the arithmetic of diffusers'
// VideoProcessor.postprocess_video: (x * 0.5 + 0.5) in bf16, clamp(0, 1), float * 255, round-half-even, uint8.
__global__ void __launch_bounds__(256)
frames_to_uint8_kernel(const __nv_bfloat16* __restrict__ in, uint8_t* __restrict__ out, int T, int64_t hw) {
  const int64_t total = static_cast<int64_t>(T) * hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint8_t px[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __bfloat162float(in[c * total + i]);
      float y = __bfloat162float(__float2bfloat16(__bfloat162float(__float2bfloat16(x * 0.5f)) + 0.5f));
      y = fminf(fmaxf(y, 0.0f), 1.0f);
      px[c] = static_cast<uint8_t>(rintf(y * 255.0f));
    }
    out[i * 3 + 0] = px[0];
    out[i * 3 + 1] = px[1];
    out[i * 3 + 2] = px[2];
  }
}

// Replicate-pad gather fused with the channel RMS-norm (+ SiLU) that precedes every causal conv of the HunyuanVideo-1.5 VAE
// (HunyuanVideo15CausalConv3d pads with mode="replicate": kt-1 frames in FRONT of time, k/2 on each side of H and W,
// vae/hunyuanvideo15/model.py:72-90; the producer is norm -> SiLU, :366-376, or nothing for conv_in / the upsample conv).
// One OUTPUT (padded) pixel per SEG-lane segment: source pixel = clamp of the padded coordinate; the norm of the few edge
// pixels is recomputed instead of re-read.  y = x / max(||x||, 1e-12) * sqrt(C) * gamma ; fp32 math, one rounding.
template <int SEG, int ITERS>
__global__ void __launch_bounds__(256)
pad_norm_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                     const __nv_bfloat16* __restrict__ gamma, int T, int H, int W, int C, int pt, int ph, int pw, int silu) {
  const int lane = threadIdx.x & 31;
  const int seg_lane = lane % SEG;
  const int segs_per_warp = 32 / SEG;
  const int To = T + pt, Ho = H + 2 * ph, Wo = W + 2 * pw;
  const int64_t pixels = static_cast<int64_t>(To) * Ho * Wo;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t pix = warp_global * segs_per_warp + lane / SEG;
  const bool active = pix < pixels;
  int64_t src = 0;
  if (active) {
    const int xo = static_cast<int>(pix % Wo);
    const int yo = static_cast<int>((pix / Wo) % Ho);
    const int to = static_cast<int>(pix / (static_cast<int64_t>(Wo) * Ho));
    const int xs = min(max(xo - pw, 0), W - 1), ys = min(max(yo - ph, 0), H - 1), ts = max(to - pt, 0);
    src = (static_cast<int64_t>(ts) * H + ys) * W + xs;
  }
  const int nchunks = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + src * C);
  uint4 raw[ITERS];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = seg_lane + i * SEG;
    if (active && c < nchunks) {
      raw[i] = xr[c];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
#pragma unroll
  for (int o = SEG / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = sqrtf(static_cast<float>(C)) / fmaxf(sqrtf(ss), 1e-12f);
  uint4* yr = reinterpret_cast<uint4*>(y + (active ? pix : 0) * C);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const int c = seg_lane + i * SEG;
    if (active && c < nchunks) {
      if (gamma != nullptr) {
        float f[8], g[8];
        unpack8(raw[i], f);
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + c), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = f[j] * inv * g[j];
          if (silu) v = 0.5f * v * (1.0f + fast_tanh(0.5f * v));  // silu = v * sigmoid(v), ONE MUFU op (tanh.approx) per element
          f[j] = v;
        }
        yr[c] = pack8(f);
      } else {
        yr[c] = raw[i];
      }
    }
  }
}

// DCAE "channel to space" upsample + shortcut of HunyuanVideo15Upsample.forward (vae/hunyuanvideo15/model.py:231-274), on
// channels-last tensors, one 16-byte chunk per thread:
//   h [T, H, W, F*Co] (conv output, F = 8 with temporal upsampling else 4), x [T, H, W, Ci] (conv input)
//   out [T', 2H, 2W, Co],  T' = 2T - 1 with temporal upsampling (the first frame is not doubled) else T
//   out(t', y, x, c) = h(f, y/2, x/2)[(a*4 + (y%2)*2 + (x%2)) * Co' + c]  +  x(f, y/2, x/2)[(a*4 + (y%2)*2 + (x%2)) * Cs + c / rep]
// with (f, a) = (0, 0) for t' = 0 (where Co' = 2*Co: `h_first[:, :C/2]` keeps the first half of the 2*Co channels and
// Cs = Ci/4, rep = repeats/2) and (f, a) = (1 + (t'-1)/2, (t'-1)%2), Co' = Co, Cs = Ci/8, rep = repeats otherwise.
__global__ void __launch_bounds__(256)
dcae_upsample_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                     int T, int H, int W, int Co, int Ci, int temporal) {
  const int chunks = Co >> 3;
  const int To = temporal ? 2 * T - 1 : T;
  const int F = temporal ? 8 : 4;
  const int repeats = F * Co / Ci;
  const int64_t total = static_cast<int64_t>(To) * 2 * H * 2 * W * chunks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % chunks);
    int64_t r = i / chunks;
    const int xo = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int yo = static_cast<int>(r % (2 * H));
    const int to = static_cast<int>(r / (2 * H));
    int f, a, co_eff, cs, rep;
    if (temporal) {
      if (to == 0) { f = 0; a = 0; co_eff = 2 * Co; cs = Ci / 4; rep = repeats / 2; }
      else { f = 1 + (to - 1) / 2; a = (to - 1) & 1; co_eff = Co; cs = Ci / 8; rep = repeats; }
    } else { f = to; a = 0; co_eff = Co; cs = Ci / 4; rep = repeats; }
    const int sub = a * 4 + (yo & 1) * 2 + (xo & 1);
    const int64_t pix = (static_cast<int64_t>(f) * H + (yo >> 1)) * W + (xo >> 1);
    const uint4 hv = *(reinterpret_cast<const uint4*>(h + pix * (static_cast<int64_t>(F) * Co) + static_cast<int64_t>(sub) * co_eff) + cc);
    float hf[8];
    unpack8(hv, hf);
    const __nv_bfloat16* xs = x + pix * Ci + static_cast<int64_t>(sub) * cs;
    const int c0 = cc << 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) hf[j] += __bfloat162float(xs[(c0 + j) / rep]);
    *(reinterpret_cast<uint4*>(out) + i) = pack8(hf);
  }
}

inline int grid_for(int64_t total, int block) {
  int64_t b = (total + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  return static_cast<int>(b < cap ? b : cap);
}

}  // namespace vae
}  // namespace b200

using namespace b200;
using namespace b200::vae;

extern "C" int b200_rmsnorm_silu_cl(const void* x, void* y, const void* gamma, int64_t pixels, int C, int silu,
                                    void* stream) {
  if (!x || !y || !gamma) return B200_ERR_ARG;
  if (pixels <= 0 || C <= 0) return B200_ERR_SHAPE;
  if (C % 8) return B200_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma)) & 15)
    return B200_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nchunks = C / 8;
  const int block = 256;
  auto blocks_for = [&](int seg) {
    const int64_t warps = (pixels + (32 / seg) - 1) / (32 / seg);
    return static_cast<unsigned>((warps * 32 + block - 1) / block);
  };
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x;
  __nv_bfloat16* yp = (__nv_bfloat16*)y;
  const __nv_bfloat16* gp = (const __nv_bfloat16*)gamma;
  if (nchunks <= 16) rmsnorm_silu_kernel<16, 1><<<blocks_for(16), block, 0, st>>>(xp, yp, gp, pixels, C, silu);
  else if (nchunks <= 32) rmsnorm_silu_kernel<32, 1><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, pixels, C, silu);
  else if (nchunks <= 64) rmsnorm_silu_kernel<32, 2><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, pixels, C, silu);
  else if (nchunks <= 128) rmsnorm_silu_kernel<32, 4><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, pixels, C, silu);
  else return B200_ERR_SHAPE;
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_upsample2x_cl(const void* in, void* out, int T, int H, int W, int C, void* stream) {
  if (!in || !out) return B200_ERR_ARG;
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0) return B200_ERR_SHAPE;
  if (C % 8) return B200_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return B200_ERR_ALIGN;
  const int64_t total = static_cast<int64_t>(T) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint4*)in, (uint4*)out, T, H, W, C / 8);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_softmax_rows(const float* s, void* p, int rows, int cols, int64_t lds, int64_t ldp, float scale,
                                 void* stream) {
  if (!s || !p) return B200_ERR_ARG;
  if (rows <= 0 || cols <= 0) return B200_ERR_SHAPE;
  const int64_t threads = static_cast<int64_t>(rows) * 32;
  softmax_rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s, (__nv_bfloat16*)p, rows, cols, lds, ldp, scale, 0);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_softmax_rows_block_causal(const float* s, void* p, int rows, int cols, int64_t lds, int64_t ldp,
                                              float scale, int block, void* stream) {
  if (!s || !p) return B200_ERR_ARG;
  if (rows <= 0 || cols <= 0 || block <= 0) return B200_ERR_SHAPE;
  const int64_t threads = static_cast<int64_t>(rows) * 32;
  softmax_rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s, (__nv_bfloat16*)p, rows, cols, lds, ldp, scale, block);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_frames_to_uint8(const void* video, void* out, int T, int H, int W, void* stream) {
  if (!video || !out) return B200_ERR_ARG;
  if (T <= 0 || H <= 0 || W <= 0) return B200_ERR_SHAPE;
  const int64_t hw = static_cast<int64_t>(H) * W;
  frames_to_uint8_kernel<<<grid_for(static_cast<int64_t>(T) * hw, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)video, (uint8_t*)out, T, hw);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

static int blend_tile_impl(void* tile, const void* up, const void* left, void* frame, int planes, int th, int tw,
                           int up_h, int up_w, int left_h, int left_w, int blend, int crop_h, int crop_w, int y0,
                           int x0, int OH, int OW, int clamp, void* stream) {
  if (!tile || !frame) return B200_ERR_ARG;
  if (planes <= 0 || th <= 0 || tw <= 0 || blend < 0) return B200_ERR_SHAPE;
  const int64_t total = static_cast<int64_t>(planes) * th * tw;
  blend_tile_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (__nv_bfloat16*)tile, (const __nv_bfloat16*)up, (const __nv_bfloat16*)left, (__nv_bfloat16*)frame, planes, th, tw,
      up_h, up_w, left_h, left_w, blend, crop_h, crop_w, y0, x0, OH, OW, clamp);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_blend_tile(void* tile, const void* up, const void* left, void* frame, int planes, int th, int tw,
                               int up_h, int up_w, int left_h, int left_w, int blend, int crop_h, int crop_w, int y0,
                               int x0, int OH, int OW, void* stream) {
  return blend_tile_impl(tile, up, left, frame, planes, th, tw, up_h, up_w, left_h, left_w, blend, crop_h, crop_w, y0, x0, OH,
                         OW, 1, stream);
}

extern "C" int b200_blend_tile_noclamp(void* tile, const void* up, const void* left, void* frame, int planes, int th,
                                       int tw, int up_h, int up_w, int left_h, int left_w, int blend, int crop_h,
                                       int crop_w, int y0, int x0, int OH, int OW, void* stream) {
  return blend_tile_impl(tile, up, left, frame, planes, th, tw, up_h, up_w, left_h, left_w, blend, crop_h, crop_w, y0, x0, OH,
                         OW, 0, stream);
}

extern "C" int b200_pad_norm_silu_cl(const void* x, void* y, const void* gamma, int T, int H, int W, int C, int pad_t,
                                     int pad_h, int pad_w, int silu, void* stream) {
  if (!x || !y) return B200_ERR_ARG;
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || pad_t < 0 || pad_h < 0 || pad_w < 0) return B200_ERR_SHAPE;
  if (C % 8) return B200_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma)) & 15)
    return B200_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t pixels = static_cast<int64_t>(T + pad_t) * (H + 2 * pad_h) * (W + 2 * pad_w);
  const int nchunks = C / 8;
  const int block = 256;
  auto blocks_for = [&](int seg) {
    const int64_t warps = (pixels + (32 / seg) - 1) / (32 / seg);
    return static_cast<unsigned>((warps * 32 + block - 1) / block);
  };
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x;
  __nv_bfloat16* yp = (__nv_bfloat16*)y;
  const __nv_bfloat16* gp = (const __nv_bfloat16*)gamma;
  if (nchunks <= 4) pad_norm_silu_kernel<4, 1><<<blocks_for(4), block, 0, st>>>(xp, yp, gp, T, H, W, C, pad_t, pad_h, pad_w, silu);
  else if (nchunks <= 16) pad_norm_silu_kernel<16, 1><<<blocks_for(16), block, 0, st>>>(xp, yp, gp, T, H, W, C, pad_t, pad_h, pad_w, silu);
  else if (nchunks <= 32) pad_norm_silu_kernel<32, 1><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, T, H, W, C, pad_t, pad_h, pad_w, silu);
  else if (nchunks <= 64) pad_norm_silu_kernel<32, 2><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, T, H, W, C, pad_t, pad_h, pad_w, silu);
  else if (nchunks <= 128) pad_norm_silu_kernel<32, 4><<<blocks_for(32), block, 0, st>>>(xp, yp, gp, T, H, W, C, pad_t, pad_h, pad_w, silu);
  else return B200_ERR_SHAPE;
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_dcae_upsample_cl(const void* h, const void* x, void* out, int T, int H, int W, int Cout, int Cin,
                                     int temporal, void* stream) {
  if (!h || !x || !out) return B200_ERR_ARG;
  if (T <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cin <= 0) return B200_ERR_SHAPE;
  const int F = temporal ? 8 : 4;
  if ((Cout % 8) || (Cin % F) || ((F * Cout) % Cin) || (temporal && (((F * Cout) / Cin) % 2))) return B200_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out)) & 15) return B200_ERR_ALIGN;
  const int To = temporal ? 2 * T - 1 : T;
  const int64_t total = static_cast<int64_t>(To) * 4 * H * W * (Cout / 8);
  dcae_upsample_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)h, (const __nv_bfloat16*)x, (__nv_bfloat16*)out, T, H, W, Cout, Cin, temporal);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// Row kernels of the dual-stream (MMDiT) block families -- Flux, HunyuanVideo-1.5, QwenImage:
// per-HEAD q/k RMS-norm + rotary embedding on interleaved (even, odd) channel pairs, in place on the q and k column
// blocks of the fused QKV buffer, both tensors in ONE launch.  HBM-bound: every element is read once and written once
// with 16-byte accesses; half a warp owns one (token, head) so the 128-channel statistic is four shuffles.
//
// The three families round differently between the norm and the rotation; `norm_mode` selects the reference's
// rounding points so the result is bit-comparable with the reference's bf16 arithmetic:
//   1  torch.nn.RMSNorm                       bf16(x * rsqrt(ms + eps) * w), one rounding
//        (flux/base/model.py:102-107, called flux/base/attention.py:70-71)
//   2  InplaceRMSNorm (efficiency/mod.py:24-35) r = bf16(rsqrt(ms + eps)); x = bf16(x * r); x = bf16(x * w)
//        (hunyuanvideo15/base/model.py:603-618, called :115-116)
//   3  diffusers RMSNorm                      x = bf16(x * rsqrt(ms + eps)); x = bf16(x * w)
//        (qwenimage/base/attention.py:115-122)
// Rotation (all three): with the pair (re, im) and the table entry (c, s) in fp32,
//   re' = re * c - im * s ;  im' = im * c + re * s ; one rounding to bf16
//   (diffusers apply_rotary_emb as called flux/base/attention.py:87-88; apply_cos_sin_rope_inplace efficiency/ops.py:163-235;
//    apply_rotary_emb_qwen qwenimage/base/attention.py:10-60).  The host bakes each family's table rounding
//   (fp64 -> fp32, or fp64 -> bf16 -> fp32 for HunyuanVideo-1.5) into the fp32 (cos, sin) table.
#include "host_util.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace mmdit {

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct HeadNormArgs {
  __nv_bfloat16* x[2];        // q and k column blocks (x[1] may be null)
  const __nv_bfloat16* w[2];  // [128] norm gains (may be null = no affine)
  const float2* rope;         // [rows, 64] (cos, sin) or null
  int64_t ldx;                // row stride of x in elements
  int rows, heads, ntens;
  int norm_mode;              // 0 = no norm
  float eps;
};

// 256 threads = 16 half-warps = 16 (tensor, token, head) items per CTA; grid-stride over items.
__global__ void __launch_bounds__(256) headnorm_rope_kernel(const HeadNormArgs a) {
  const int lane16 = threadIdx.x & 15;
  const int64_t items = static_cast<int64_t>(a.rows) * a.heads * a.ntens;
  const int64_t per_tensor = static_cast<int64_t>(a.rows) * a.heads;
  // a half-warp leaves the loop as a unit; the shuffles below name only its own 16 lanes
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  for (int64_t item = blockIdx.x * 16 + (threadIdx.x >> 4); item < items;
       item += static_cast<int64_t>(gridDim.x) * 16) {
    const int t = static_cast<int>(item / per_tensor);
    const int64_t rem = item - t * per_tensor;
    const int64_t row = rem / a.heads;
    const int head = static_cast<int>(rem - row * a.heads);
    __nv_bfloat16* xt = t ? a.x[1] : a.x[0];
    const __nv_bfloat16* wt = t ? a.w[1] : a.w[0];
    uint4* p = reinterpret_cast<uint4*>(xt + row * a.ldx + head * 128) + lane16;
    const uint4 raw = *p;
    float f[8];
    f[0] = bf16_lo(raw.x); f[1] = bf16_hi(raw.x); f[2] = bf16_lo(raw.y); f[3] = bf16_hi(raw.y);
    f[4] = bf16_lo(raw.z); f[5] = bf16_hi(raw.z); f[6] = bf16_lo(raw.w); f[7] = bf16_hi(raw.w);
    if (a.norm_mode != 0) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(mask, ss, o);
      float r = rsqrtf(ss * (1.0f / 128.0f) + a.eps);
      float w[8];
      if (wt != nullptr) {
        const uint4 wr = __ldg(reinterpret_cast<const uint4*>(wt) + lane16);
        w[0] = bf16_lo(wr.x); w[1] = bf16_hi(wr.x); w[2] = bf16_lo(wr.y); w[3] = bf16_hi(wr.y);
        w[4] = bf16_lo(wr.z); w[5] = bf16_hi(wr.z); w[6] = bf16_lo(wr.w); w[7] = bf16_hi(wr.w);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 1.0f;
      }
      if (a.norm_mode == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = round_bf16(__fmul_rn(__fmul_rn(f[j], r), w[j]));
      } else {
        if (a.norm_mode == 2) r = round_bf16(r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = round_bf16(round_bf16(f[j] * r) * w[j]);
      }
    }
    if (a.rope != nullptr) {
      const float4* rp = reinterpret_cast<const float4*>(a.rope + row * 64 + lane16 * 4);
      const float4 cs0 = __ldg(rp), cs1 = __ldg(rp + 1);
      const float c[4] = {cs0.x, cs0.z, cs1.x, cs1.z};
      const float s[4] = {cs0.y, cs0.w, cs1.y, cs1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float re = f[2 * j], im = f[2 * j + 1];
        // separate fp32 roundings of the two products, like the reference's tensor-level mul / add
        f[2 * j] = __fsub_rn(__fmul_rn(re, c[j]), __fmul_rn(im, s[j]));
        f[2 * j + 1] = __fadd_rn(__fmul_rn(im, c[j]), __fmul_rn(re, s[j]));
      }
    }
    uint4 out;
    out.x = pack_bf16x2(f[0], f[1]); out.y = pack_bf16x2(f[2], f[3]);
    out.z = pack_bf16x2(f[4], f[5]); out.w = pack_bf16x2(f[6], f[7]);
    *p = out;
  }
}

// SwiGLU / GEGLU-style gating of a [rows, 2*inner] projection: y[:, j] = act(a[:, j]) * a[:, inner + j]
// (Flux2FeedForward: flux2/base/model.py:91-130 -- silu(x1) * x2, rounded once like the reference's bf16 tensor ops:
//  bf16(silu) then bf16(mul)).
__global__ void __launch_bounds__(256)
swiglu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int rows, int inner, int64_t ldx,
              int64_t ldy) {
  const int nchunks = inner >> 3;
  const int64_t total = static_cast<int64_t>(rows) * nchunks;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / nchunks;
    const int c = static_cast<int>(idx - row * nchunks);
    const uint4 a = *(reinterpret_cast<const uint4*>(x + row * ldx) + c);
    const uint4 b = *(reinterpret_cast<const uint4*>(x + row * ldx + inner) + c);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = bf16_lo(av[j]), a1 = bf16_hi(av[j]);
      const float s0 = round_bf16(a0 / (1.0f + expf(-a0))), s1 = round_bf16(a1 / (1.0f + expf(-a1)));
      o[j] = pack_bf16x2(s0 * bf16_lo(bv[j]), s1 * bf16_hi(bv[j]));
    }
    *(reinterpret_cast<uint4*>(y + row * ldy) + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}


// Row RMS-norm with gain over the whole row (text-stream input norm of QwenImage: diffusers RMSNorm(joint_dim),
// qwenimage/base/model.py:826,920), out of place; `mode` as in headnorm_rope.  One CTA per row; the row is read twice
// (the second read is an L1/L2 hit) -- the text stream has a few hundred rows.
__global__ void __launch_bounds__(256)
rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ w,
                    int dim, int64_t ldx, int64_t ldy, float eps, int mode) {
  __shared__ float red[8];
  const int64_t row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
  const int nchunks = dim >> 3;
  float ss = 0.f;
  for (int c = threadIdx.x; c < nchunks; c += 256) {
    const uint4 u = xr[c];
    const uint32_t v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(v[j]), b = bf16_hi(v[j]);
      ss += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  float r = rsqrtf(tot / static_cast<float>(dim) + eps);
  if (mode == 2) r = round_bf16(r);
  for (int c = threadIdx.x; c < nchunks; c += 256) {
    const uint4 u = xr[c];
    const uint4 g = w ? __ldg(reinterpret_cast<const uint4*>(w) + c) : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    const uint32_t v[4] = {u.x, u.y, u.z, u.w}, gv[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = bf16_lo(v[j]), b = bf16_hi(v[j]);
      if (mode == 1) {
        a = __fmul_rn(__fmul_rn(a, r), bf16_lo(gv[j]));
        b = __fmul_rn(__fmul_rn(b, r), bf16_hi(gv[j]));
      } else {
        a = round_bf16(a * r) * bf16_lo(gv[j]);
        b = round_bf16(b * r) * bf16_hi(gv[j]);
      }
      o[j] = pack_bf16x2(a, b);
    }
    yr[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace mmdit
}  // namespace b200

using namespace b200;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int b200_headnorm_rope(void* q, void* k, const void* wq, const void* wk, const float* rope, int rows,
                                  int heads, int head_dim, int64_t ldx, float eps, int norm_mode, void* stream) {
  if (!q) return B200_ERR_ARG;
  if (norm_mode < 0 || norm_mode > 3) return B200_ERR_ARG;
  if (rows <= 0 || heads <= 0) return B200_ERR_SHAPE;
  if (head_dim != 128) return B200_ERR_SHAPE;
  if (ldx % 8) return B200_ERR_ALIGN;
  if (!aligned16(q) || !aligned16(k) || !aligned16(wq) || !aligned16(wk) || !aligned16(rope)) return B200_ERR_ALIGN;
  mmdit::HeadNormArgs a;
  a.x[0] = static_cast<__nv_bfloat16*>(q);
  a.x[1] = static_cast<__nv_bfloat16*>(k);
  a.w[0] = static_cast<const __nv_bfloat16*>(wq);
  a.w[1] = static_cast<const __nv_bfloat16*>(wk);
  a.rope = reinterpret_cast<const float2*>(rope);
  a.ldx = ldx;
  a.rows = rows;
  a.heads = heads;
  a.ntens = k ? 2 : 1;
  a.norm_mode = norm_mode;
  a.eps = eps;
  const int64_t items = static_cast<int64_t>(rows) * heads * a.ntens;
  int64_t blocks = (items + 15) / 16;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  mmdit::headnorm_rope_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_swiglu(const void* x, void* y, int rows, int inner, int64_t ldx, int64_t ldy, void* stream) {
  if (!x || !y) return B200_ERR_ARG;
  if (rows <= 0 || inner <= 0) return B200_ERR_SHAPE;
  if ((inner % 8) || (ldx % 8) || (ldy % 8) || !aligned16(x) || !aligned16(y)) return B200_ERR_ALIGN;
  const int64_t total = static_cast<int64_t>(rows) * (inner / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  mmdit::swiglu_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), rows, inner, ldx, ldy);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_rmsnorm_rows(const void* x, void* y, const void* w, int rows, int dim, int64_t ldx, int64_t ldy,
                                 float eps, int norm_mode, void* stream) {
  if (!x || !y) return B200_ERR_ARG;
  if (norm_mode < 1 || norm_mode > 3) return B200_ERR_ARG;
  if (rows <= 0 || dim <= 0) return B200_ERR_SHAPE;
  if ((dim % 8) || (ldx % 8) || (ldy % 8) || !aligned16(x) || !aligned16(y) || !aligned16(w)) return B200_ERR_ALIGN;
  mmdit::rmsnorm_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(w), dim, ldx, ldy,
      eps, norm_mode);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// HBM-bound row kernels of the Wan block: adaLN modulated LayerNorm, q/k RMS-norm (+ 3-axis RoPE),
// gate * y + residual, CFG combine.  One CTA per row, 16-byte vector loads, the row lives in registers
// between the statistics pass and the write, so every tensor is read once and written once.
//
// These kernels reproduce the reference's bf16 rounding points exactly ("Rounding
// points"), so against the reference's own bf16 arithmetic they differ only through the fp32 reduction
// order of the row statistics.
#include "host_util.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace ew {

constexpr int MAX_CHUNKS = 8;  // uint4 (8 x bf16) chunks per thread held in registers

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16(x)); }

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

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics, two-pass) + optional affine + optional (1+scale), shift modulation.
// reference: FP32LayerNorm -> .to(bf16); x.addcmul_(x, scale); x.add_(shift)   (model.py:56-116, ops.py:37-56)
// ------------------------------------------------------------------------------------------------
// MODE 0: Wan (above).  MODE 1: diffusers AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous and the
// explicit `norm2(x) * (1 + scale) + shift` of the dual-stream blocks (flux/base/model.py:266,297-300,
// hunyuanvideo15/base/model.py:617-694): nn.LayerNorm in bf16 (fp32 inside, one rounding), then three bf16 tensor ops
// s1 = bf16(1 + scale); y = bf16(y * s1); y = bf16(y + shift).
template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS)
layernorm_modulate_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                          const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift,
                          const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b, int dim,
                          int64_t ldx, int64_t ldy, int64_t mod_stride, float eps) {
  __shared__ float red[THREADS / 32];
  const int64_t row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
  const int nchunks = dim >> 3;

  uint4 raw[MAX_CHUNKS];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * THREADS;
    if (c < nchunks) {
      raw[i] = xr[c];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  }
  const float mean = block_sum<THREADS>(s, red) / static_cast<float>(dim);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * THREADS;
    if (c < nchunks) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[j] - mean;
        ss += d * d;
      }
    }
  }
  const float var = block_sum<THREADS>(ss, red) / static_cast<float>(dim);
  const float rstd = rsqrtf(var + eps);

  const uint4* sc = scale ? reinterpret_cast<const uint4*>(scale + row * mod_stride) : nullptr;
  const uint4* sh = shift ? reinterpret_cast<const uint4*>(shift + row * mod_stride) : nullptr;
  const uint4* lw = ln_w ? reinterpret_cast<const uint4*>(ln_w) : nullptr;
  const uint4* lb = ln_b ? reinterpret_cast<const uint4*>(ln_b) : nullptr;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * THREADS;
    if (c < nchunks) {
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd;
      if (lw) {
        float w[8], b[8];
        unpack8(__ldg(lw + c), w);
        if (lb) unpack8(__ldg(lb + c), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = lb ? f[j] * w[j] + b[j] : f[j] * w[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = round_bf16(f[j]);  // FP32LayerNorm(...).to(bf16)
      if (sc) {
        float a[8], b[8];
        unpack8(__ldg(sc + c), a);
        unpack8(__ldg(sh + c), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = MODE == 0 ? round_bf16(f[j] + f[j] * a[j])              // addcmul_(x, scale)
                                    : round_bf16(f[j] * round_bf16(1.0f + a[j])); // norm(x) * (1 + scale)
          f[j] = t + b[j];                                 // add_(shift), rounded by pack8
        }
      }
      yr[c] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// In-place RMS-norm over the whole row (all heads) + RoPE on (even, odd) pairs.
// reference: InplaceRMSNorm (mod.py:24-35), apply_wan_rope_inplace (ops.py:101-160)
// rope table: bf16 [rows, head_dim] = (cos0, sin0, cos1, sin1, ...) already cast to bf16 as the reference does.
// ------------------------------------------------------------------------------------------------
// Scatter mode (sequence-parallel "tokens -> heads" exchange fused into this kernel): instead of updating x in
// place, channel chunk c of token `row` is stored straight into the peer GPU that owns its head group, over
// NVLink peer memory: dst = peer[ch / width] + elem_off + (row0 + row) * width + ch % width.
struct PeerScatter {
  void* peer[8];
  int n_peers;        // 0 = in place
  int width;          // channels per peer = (heads / P) * head_dim
  int row0;           // first global token of this rank's shard
  int64_t elem_off;   // element offset of the q / k / v plane inside each peer's receive buffer
  int64_t batch_elem_off;  // added per batch index (blockIdx.y): the next plane
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                    const __nv_bfloat16* __restrict__ rope, int dim, int head_dim, int64_t ldx, float eps,
                    const PeerScatter sc, int64_t x_batch_stride, int64_t w_batch_stride) {
  __shared__ float red[THREADS / 32];
  const int64_t row = blockIdx.x;
  // batch index (grid.y): another column block of the same rows with its own weight vector -- q and k of the fused q|k|v
  // buffer in ONE launch, or the K blocks of all layers' cross-attention projections
  x += static_cast<int64_t>(blockIdx.y) * x_batch_stride;
  if (w != nullptr) w += static_cast<int64_t>(blockIdx.y) * w_batch_stride;
  uint4* xr = reinterpret_cast<uint4*>(x + row * ldx);
  const int nchunks = dim >> 3;
  const int chunks_per_head = head_dim >> 3;

  uint4 raw[MAX_CHUNKS];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * THREADS;
    if (c < nchunks) {
      raw[i] = xr[c];
      float f[8];
      unpack8(raw[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  float r = 1.0f;
  if (w != nullptr || eps >= 0.f) {
    const float ms = block_sum<THREADS>(ss, red) / static_cast<float>(dim);
    r = round_bf16(rsqrtf(ms + eps));  // y.to(dtype=x.dtype)
  }
  const uint4* wr = w ? reinterpret_cast<const uint4*>(w) : nullptr;
  const uint4* rr = rope ? reinterpret_cast<const uint4*>(rope + row * head_dim) : nullptr;
#pragma unroll
  for (int i = 0; i < MAX_CHUNKS; ++i) {
    const int c = threadIdx.x + i * THREADS;
    if (c < nchunks) {
      float f[8];
      unpack8(raw[i], f);
      if (eps >= 0.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = round_bf16(f[j] * r);  // x.mul_(rsqrt)
        if (wr) {
          float ww[8];
          unpack8(__ldg(wr + c), ww);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = round_bf16(f[j] * ww[j]);  // x.mul_(weight)
        }
      }
      if (rr) {
        float cs[8];
        unpack8(__ldg(rr + (c % chunks_per_head)), cs);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float re = f[2 * j], im = f[2 * j + 1];
          const float co = cs[2 * j], si = cs[2 * j + 1];
          // x_real.mul_(c).addcmul_(x_imag, s, value=-1); x_imag.mul_(c).addcmul_(x_real_orig, s, value=1)
          f[2 * j] = round_bf16(re * co) - im * si;
          f[2 * j + 1] = round_bf16(im * co) + re * si;
        }
      }
      if (sc.n_peers == 0) {
        xr[c] = pack8(f);
      } else {
        const int ch = c << 3;
        const int d = ch / sc.width;
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sc.peer[d]) + sc.elem_off + blockIdx.y * sc.batch_elem_off +
                             (static_cast<int64_t>(sc.row0) + row) * sc.width + (ch - d * sc.width);
        *reinterpret_cast<uint4*>(dst) = pack8(f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// h += y * gate    (y.mul_(gate) rounds to bf16, then h.add_(y))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_residual_kernel(__nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ y,
                     const __nv_bfloat16* __restrict__ gate, int rows, int dim, int64_t ldh, int64_t ldy) {
  const int nchunks = dim >> 3;
  const int64_t total = static_cast<int64_t>(rows) * nchunks;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / nchunks;
    const int c = static_cast<int>(idx - row * nchunks);
    uint4* hp = reinterpret_cast<uint4*>(h + row * ldh) + c;
    const uint4 hv = *hp;
    const uint4 yv = *(reinterpret_cast<const uint4*>(y + row * ldy) + c);
    float hf[8], yf[8];
    unpack8(hv, hf);
    unpack8(yv, yf);
    if (gate) {
      float g[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gate) + c), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) yf[j] = round_bf16(yf[j] * g[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) hf[j] += yf[j];
    *hp = pack8(hf);
  }
}

// ------------------------------------------------------------------------------------------------
// noise = u + g * (c - u) in bf16 tensor arithmetic (three roundings), bf16 out like the reference.
// reference: engine/wan/shared/__init__.py:565
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cfg_combine_kernel(const __nv_bfloat16* __restrict__ c, const __nv_bfloat16* __restrict__ u,
                   __nv_bfloat16* __restrict__ out, float g, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float cf = __bfloat162float(c[i]);
    const float uf = __bfloat162float(u[i]);
    const float d = round_bf16(cf - uf);
    const float gd = round_bf16(g * d);
    out[i] = __float2bfloat16(uf + gd);
  }
}

template <typename F>
int dispatch_threads(int nchunks, F&& f) {
  if (nchunks <= 64 * MAX_CHUNKS) return f(std::integral_constant<int, 64>());
  if (nchunks <= 128 * MAX_CHUNKS) return f(std::integral_constant<int, 128>());
  if (nchunks <= 256 * MAX_CHUNKS) return f(std::integral_constant<int, 256>());
  if (nchunks <= 512 * MAX_CHUNKS) return f(std::integral_constant<int, 512>());
  if (nchunks <= 1024 * MAX_CHUNKS) return f(std::integral_constant<int, 1024>());
  return B200_ERR_SHAPE;
}

}  // namespace ew
}  // namespace b200

using namespace b200;
using namespace b200::ew;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int layernorm_modulate_launch(const void* x, void* y, const void* scale, const void* shift, const void* ln_w,
                                     const void* ln_b, int rows, int dim, int64_t ldx, int64_t ldy,
                                     int64_t mod_stride, float eps, int mode, void* stream) {
  if (!x || !y) return B200_ERR_ARG;
  if ((scale == nullptr) != (shift == nullptr)) return B200_ERR_ARG;
  if (rows <= 0 || dim <= 0) return B200_ERR_SHAPE;
  if ((dim % 8) || (ldx % 8) || (ldy % 8) || (mod_stride % 8)) return B200_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(y) || !aligned16(scale) || !aligned16(shift) || !aligned16(ln_w) || !aligned16(ln_b))
    return B200_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = dispatch_threads(dim / 8, [&](auto T) {
    constexpr int THREADS = decltype(T)::value;
    if (mode == 0)
      layernorm_modulate_kernel<THREADS, 0><<<rows, THREADS, 0, st>>>(
          (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)scale, (const __nv_bfloat16*)shift,
          (const __nv_bfloat16*)ln_w, (const __nv_bfloat16*)ln_b, dim, ldx, ldy, mod_stride, eps);
    else
      layernorm_modulate_kernel<THREADS, 1><<<rows, THREADS, 0, st>>>(
          (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)scale, (const __nv_bfloat16*)shift,
          (const __nv_bfloat16*)ln_w, (const __nv_bfloat16*)ln_b, dim, ldx, ldy, mod_stride, eps);
    return B200_OK;
  });
  if (rc) return rc;
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_layernorm_modulate(const void* x, void* y, const void* scale, const void* shift, const void* ln_w,
                                       const void* ln_b, int rows, int dim, int64_t ldx, int64_t ldy,
                                       int64_t mod_stride, float eps, void* stream) {
  return layernorm_modulate_launch(x, y, scale, shift, ln_w, ln_b, rows, dim, ldx, ldy, mod_stride, eps, 0, stream);
}

extern "C" int b200_adaln_zero_modulate(const void* x, void* y, const void* scale, const void* shift, int rows, int dim,
                                        int64_t ldx, int64_t ldy, float eps, void* stream) {
  return layernorm_modulate_launch(x, y, scale, shift, nullptr, nullptr, rows, dim, ldx, ldy, 0, eps, 1, stream);
}

static int rmsnorm_rope_launch(void* x, const void* w, const void* rope, int rows, int heads, int head_dim,
                               int64_t ldx, float eps, const PeerScatter& sc, void* stream, int n_batch = 1,
                               int64_t x_batch_stride = 0, int64_t w_batch_stride = 0) {
  if (!x) return B200_ERR_ARG;
  if (rows <= 0 || heads <= 0 || head_dim <= 0 || n_batch <= 0 || n_batch > 65535) return B200_ERR_SHAPE;
  if ((head_dim % 8) || (ldx % 8) || (x_batch_stride % 8) || (w_batch_stride % 8)) return B200_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(w) || !aligned16(rope)) return B200_ERR_ALIGN;
  const int dim = heads * head_dim;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = dispatch_threads(dim / 8, [&](auto T) {
    constexpr int THREADS = decltype(T)::value;
    rmsnorm_rope_kernel<THREADS><<<dim3(rows, n_batch), THREADS, 0, st>>>((__nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                                          (const __nv_bfloat16*)rope, dim, head_dim, ldx, eps, sc,
                                                                          x_batch_stride, w_batch_stride);
    return B200_OK;
  });
  if (rc) return rc;
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_rmsnorm_rope(void* x, const void* w, const void* rope, int rows, int heads, int head_dim,
                                 int64_t ldx, float eps, void* stream) {
  PeerScatter sc;
  sc.n_peers = 0;
  sc.batch_elem_off = 0;
  return rmsnorm_rope_launch(x, w, rope, rows, heads, head_dim, ldx, eps, sc, stream);
}

extern "C" int b200_rmsnorm_rope_batched(void* x, const void* w, const void* rope, int rows, int heads, int head_dim,
                                         int64_t ldx, float eps, int n_batch, int64_t x_batch_stride, int64_t w_batch_stride,
                                         void* stream) {
  PeerScatter sc;
  sc.n_peers = 0;
  sc.batch_elem_off = 0;
  return rmsnorm_rope_launch(x, w, rope, rows, heads, head_dim, ldx, eps, sc, stream, n_batch, x_batch_stride, w_batch_stride);
}

extern "C" int b200_rmsnorm_rope_scatter(const void* x, const void* w, const void* rope, int rows, int heads,
                                         int head_dim, int64_t ldx, float eps, void* const* peers, int n_peers,
                                         int64_t dst_elem_offset, int row0, void* stream) {
  if (!peers || n_peers < 1 || n_peers > 8 || (heads % n_peers)) return B200_ERR_ARG;
  PeerScatter sc;
  sc.n_peers = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    if (!peers[i] || !aligned16(peers[i])) return B200_ERR_ALIGN;
    sc.peer[i] = peers[i];
  }
  sc.width = (heads / n_peers) * head_dim;
  sc.row0 = row0;
  sc.elem_off = dst_elem_offset;
  sc.batch_elem_off = 0;
  if (dst_elem_offset % 8) return B200_ERR_ALIGN;
  return rmsnorm_rope_launch(const_cast<void*>(x), w, rope, rows, heads, head_dim, ldx, eps, sc, stream);
}

extern "C" int b200_gate_residual(void* h, const void* y, const void* gate, int rows, int dim, int64_t ldh, int64_t ldy,
                                  void* stream) {
  if (!h || !y) return B200_ERR_ARG;
  if (rows <= 0 || dim <= 0) return B200_ERR_SHAPE;
  if ((dim % 8) || (ldh % 8) || (ldy % 8)) return B200_ERR_ALIGN;
  if (!aligned16(h) || !aligned16(y) || !aligned16(gate)) return B200_ERR_ALIGN;
  const int64_t total = static_cast<int64_t>(rows) * (dim / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  gate_residual_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (__nv_bfloat16*)h, (const __nv_bfloat16*)y, (const __nv_bfloat16*)gate, rows, dim, ldh, ldy);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_cfg_combine(const void* cond, const void* uncond, void* out, float guidance, int64_t n,
                                void* stream) {
  if (!cond || !uncond || !out) return B200_ERR_ARG;
  if (n <= 0) return B200_ERR_SHAPE;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cfg_combine_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)cond, (const __nv_bfloat16*)uncond, (__nv_bfloat16*)out, guidance, n);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

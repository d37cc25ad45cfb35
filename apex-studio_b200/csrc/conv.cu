// b200_conv3d_cl: causal 3-D convolution on channels-last activations as an implicit GEMM on tcgen05.
//
//   y[t, h, w, :] = bias + sum_{kt,kh,kw} W[kt,kh,kw] . x[t + kt - (KT-1), h + kh - KH/2, w + kw - KW/2, :]
//
// (zero outside the volume: KT-1 zero frames on the LEFT of time only -- WanCausalConv3d, vae/wan/model.py:136-185 --
// and symmetric zero padding in H/W.)  One kernel covers the 3x3x3 convs of the residual blocks, conv_in/conv_out, the
// (3,1,1) time_conv of upsample3d and the per-frame 3x3 Conv2d of WanResample.
//
// GEMM view: M = output pixels (tiles of 128 = BH rows x BW columns of ONE frame), N = C_out, K = taps x C_in.
// The A operand of tap (kt,kh,kw) is a shifted [BH x BW x BK] window of x, fetched by ONE 4-D TMA box whose
// out-of-bounds elements (halo, causal left pad) are zero-filled by the TMA unit -- no im2col buffer, no padded copy.
// The B operand is the [BN x BK] slice of the tap's weight matrix (host layout [tap][C_out][C_in]).
// Forms (round 2): KC channel chunks per pipeline stage (KC = 3 puts all 96 channels of a tap behind ONE barrier round trip);
// 32-byte epilogue loads / stores; a fused consumer epilogue (RMS-norm + SiLU of the conv output: b200_conv3d_cl_norm_silu);
// an opt-in temporally blocked kernel (conv3d_tb_kernel, B200_CONV_TB) kept with its measurements.
// Warp roles / pipeline are those of linear.cu: warp 0 TMA, warp 1 MMA (fp32 accumulators in TMEM, two buffers),
// warps 2-5 epilogue (bias, optional residual add, bf16 channels-last store, or planar store for conv_out, or the
// 2x temporal interleave of upsample3d, vae/wan/model.py:332-334).
#include "host_util.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace conv {

constexpr int BM = 128;
constexpr int BN_MAX = 256;
constexpr int NUM_THREADS = 192;

// KC = channel chunks of BK per pipeline stage.  With BK = 32 and N = 96 (the VAE's full-resolution stage) a stage of ONE chunk
// is two MMAs of 48 tensor cycles each behind a barrier wait and a commit: the single issuing thread, not the tensor pipe,
// bounds the kernel (630 TFLOP/s; halving the L2 traffic did not help: profiles/r02_conv_tb_ab.log).  KC = 3 puts all 96
// channels of a tap in one stage: one wait + one commit per six MMAs (N <= 128 so that four 48 KB stages fit).
template <int BK, int KC = 1>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;                                  // one chunk of the activation window
  static constexpr int B_BYTES = (KC == 1 ? BN_MAX : 128) * BK * 2;            // one chunk of the weight slice
  static constexpr int BN_LIMIT = KC == 1 ? BN_MAX : 128;
  static constexpr int STAGE_BYTES = KC * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (KC > 1) ? 4 : ((BK == 64) ? 4 : 8);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr uint32_t SBO = 8 * BK * 2;          // bytes between 8-row groups (dense rows of BK bf16)
  static constexpr uint64_t LAYOUT = (BK == 64) ? 2 : 4;  // SWIZZLE_128B : SWIZZLE_64B
};

struct Params {
  int T, H, W, Cin, Cout;     // output volume == input volume (stride 1)
  int KT, KH, KW;
  int BW, BH, BN;             // pixel tile = BH x BW (BH*BW == 128), N tile
  int tiles_w, tiles_h, tiles_n;
  const __nv_bfloat16* bias;      // [Cout] or null
  const __nv_bfloat16* residual;  // channels-last [T,H,W,Cout] or null
  void* out;
  int out_mode;               // 0: channels-last bf16 [T',H,W,Csplit]; 1: planar bf16 [Cvalid,T,H,W]
  int out_t_mul, out_t_off;   // destination frame = t * out_t_mul + out_t_off + (n / Csplit)
  int Csplit;                 // channels per destination frame (== Cout unless temporal interleave)
  int Cvalid;                 // number of real output channels (planar mode; <= Cout)
  // Pre-padded input (replicate padding materialised by the producer kernel, HunyuanVideo-1.5 VAE): the input tensor is
  // [T + pad_t, H + 2*pad_h, W + 2*pad_w, Cin] and tap coordinates are shifted by the pads; 0 = zero fill by TMA.
  int pad_t, pad_h, pad_w;
  int TB, tblocks;            // conv3d_tb_kernel: output frames per CTA block, ceil(T / TB)
  int vec32;                  // channels-last rows of out / residual are 32-byte aligned: 32-byte loads and stores
  // Fused consumer (WanResidualBlock: conv1 -> norm2 -> SiLU, vae/wan/model.py:404-413): non-null = the epilogue applies
  // WanRMS_norm (x / max(||x||_2, 1e-12) * sqrt(C) * gamma over the channels of a pixel) + SiLU to the bf16-rounded conv
  // output and stores only that.  Needs the whole channel vector in one N tile (tiles_n == 1).
  const __nv_bfloat16* norm_gamma;
};

template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((Cfg<BK>::SBO >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= Cfg<BK>::LAYOUT << 61;
  return d;
}

B200_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Epilogue of one [128 pixels x BN channels] accumulator (TMEM address t_addr = lane quadrant + first column) of output frame t:
// bias, optional residual add, bf16 channels-last / planar / temporally interleaved store.  One pixel row per thread.
__device__ __forceinline__ void store_rows(const Params& p, int t, int h0, int w0, int tn, uint32_t t_addr, int quad, int lane) {
  const int64_t frame_px = static_cast<int64_t>(p.H) * p.W;
  const int r = quad * 32 + lane;
  const int h = h0 + r / p.BW;
  const int w = w0 + r % p.BW;
  const bool ok = (h < p.H) && (w < p.W);
  const int64_t pix = static_cast<int64_t>(h) * p.W + w;
  const int n_base = tn * p.BN;
  for (int c0 = 0; c0 < p.BN; c0 += 16) {
    uint32_t rr[16];
    tmem_ld_x16(t_addr + c0, rr);
    tmem_ld_wait();
    const int n0 = n_base + c0;  // first output channel of this 16-wide chunk
    if (n0 >= p.Cout) break;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
    if (p.bias != nullptr) {
      const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 b = __ldg(bp + q);
        v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
        v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
        v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
        v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
      }
    }
    if (!ok) continue;
    if (p.residual != nullptr) {
      const __nv_bfloat16* rsrc = p.residual + (static_cast<int64_t>(t) * frame_px + pix) * p.Cout + n0;
      uint4 rv[2];
      if (p.vec32) {
        ld_global_v8(rsrc, rv);   // one full 32-byte sector per lane
      } else {
        rv[0] = reinterpret_cast<const uint4*>(rsrc)[0];
        rv[1] = reinterpret_cast<const uint4*>(rsrc)[1];
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 b = rv[q];
        v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
        v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
        v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
        v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
      }
    }
    if (p.out_mode == 0) {
      const int g = n0 / p.Csplit;
      const int cdst = n0 - g * p.Csplit;
      const int64_t tf = static_cast<int64_t>(t) * p.out_t_mul + p.out_t_off + g;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (tf * frame_px + pix) * p.Csplit + cdst;
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]);  o0.y = pack_bf16x2(v[2], v[3]);
      o0.z = pack_bf16x2(v[4], v[5]);  o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]);  o1.y = pack_bf16x2(v[10], v[11]);
      o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
      if (p.vec32) {
        // 32 bytes per lane: a full sector (two 16-byte stores per lane wrote every sector in two partial pieces and doubled
        // the L1 -> L2 write transactions: ncu 63.7 M sector writes for 31.9 M sectors of output)
        const uint32_t ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        st_global_v8(op, ov);
      } else {
        reinterpret_cast<uint4*>(op)[0] = o0;
        reinterpret_cast<uint4*>(op)[1] = o1;
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < p.Cvalid)
          op[(static_cast<int64_t>(n0 + j) * p.T + t) * frame_px + pix] = __float2bfloat16(v[j]);
    }
  }
}

// Epilogue with the fused RMS-norm + SiLU (Params::norm_gamma): two passes over the accumulator row -- sum of squares of the
// bf16-rounded conv output (what the unfused path stores and the norm kernel reads back), then normalise, SiLU, store.  Saves
// the norm kernel's launch and its read + write pass over the activation (HBM-bound: 2 GB per launch at the 96-channel stage).
__device__ __forceinline__ void store_rows_norm(const Params& p, int t, int h0, int w0, uint32_t t_addr, int quad, int lane) {
  const int64_t frame_px = static_cast<int64_t>(p.H) * p.W;
  const int r = quad * 32 + lane;
  const int h = h0 + r / p.BW;
  const int w = w0 + r % p.BW;
  const bool ok = (h < p.H) && (w < p.W);
  const int64_t pix = static_cast<int64_t>(h) * p.W + w;
  auto load_chunk = [&](int c0, float (&v)[16]) {
    uint32_t rr[16];
    tmem_ld_x16(t_addr + c0, rr);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
    if (p.bias != nullptr) {
      const uint4* bp = reinterpret_cast<const uint4*>(p.bias + c0);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 b = __ldg(bp + q);
        v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
        v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
        v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
        v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __bfloat162float(__float2bfloat16(v[j]));   // the conv output as the reference stores it
  };
  float ss = 0.f;
  for (int c0 = 0; c0 < p.BN; c0 += 16) {
    float v[16];
    load_chunk(c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) ss += v[j] * v[j];
  }
  const float inv = sqrtf(static_cast<float>(p.Cout)) / fmaxf(sqrtf(ss), 1e-12f);
  for (int c0 = 0; c0 < p.BN; c0 += 16) {
    float v[16];
    load_chunk(c0, v);
    if (!ok) continue;
    const uint4* gp = reinterpret_cast<const uint4*>(p.norm_gamma + c0);
    uint32_t ov[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint4 g = __ldg(gp + q);
      const float gg[8] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y), bf16_hi(g.y), bf16_lo(g.z), bf16_hi(g.z), bf16_lo(g.w), bf16_hi(g.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = v[q * 8 + j] * inv * gg[j];
        v[q * 8 + j] = 0.5f * y * (1.0f + fast_tanh(0.5f * y));   // SiLU as in vae_ops.cu (one MUFU op per element)
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ov[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (static_cast<int64_t>(t) * frame_px + pix) * p.Cout + c0;
    if (p.vec32) {
      st_global_v8(op, ov);
    } else {
      reinterpret_cast<uint4*>(op)[0] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      reinterpret_cast<uint4*>(op)[1] = make_uint4(ov[4], ov[5], ov[6], ov[7]);
    }
  }
}

template <int BK, int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, Params p) {
  using C = Cfg<BK, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int taps = p.KT * p.KH * p.KW;
  const int kchunks = (p.Cin + BK * KC - 1) / (BK * KC);   // pipeline stages per tap (a partial last chunk is zero-filled by TMA)
  const int tiles_per_frame = p.tiles_h * p.tiles_w;
  const int total_tiles = p.T * tiles_per_frame * p.tiles_n;
  const uint32_t stage_tx = static_cast<uint32_t>(KC * (BM * BK + p.BN * BK) * 2);

  // tile -> (n tile, frame, pixel-tile origin); n fastest so that neighbouring CTAs share the activation window
  auto decode = [&](int tile, int& tn, int& t, int& h0, int& w0) {
    tn = tile % p.tiles_n;
    int m = tile / p.tiles_n;
    t = m / tiles_per_frame;
    int r = m - t * tiles_per_frame;
    h0 = (r / p.tiles_w) * p.BH;
    w0 = (r % p.tiles_w) * p.BW;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int tn, t, h0, w0;
        decode(tile, tn, t, h0, w0);
        for (int tap = 0; tap < taps; ++tap) {
          const int kt = tap / (p.KH * p.KW);
          const int kh = (tap / p.KW) % p.KH;
          const int kw = tap % p.KW;
          const int ti = t + kt - (p.KT - 1) + p.pad_t;
          const int hi = h0 + kh - p.KH / 2 + p.pad_h;
          const int wi = w0 + kw - p.KW / 2 + p.pad_w;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            uint8_t* sb = sa + KC * C::A_BYTES;
            mbar_arrive_expect_tx(&full[stage], stage_tx);
#pragma unroll
            for (int c = 0; c < KC; ++c) {
              tma_load_4d(sa + c * C::A_BYTES, &tmX, &full[stage], (kc * KC + c) * BK, wi, hi, ti);
              tma_load_2d(sb + c * C::B_BYTES, &tmW, &full[stage], (kc * KC + c) * BK, tap * p.Cout + tn * p.BN);
            }
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(BM, p.BN, 0);
      const int k_iters = taps * kchunks;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          // descriptors of the stage's first chunk; chunk c / 16-channel slice k only move the 16-byte-unit start address
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_desc = make_desc<BK>(a_addr);
          const uint64_t b_desc = make_desc<BK>(a_addr + KC * C::A_BYTES);
#pragma unroll
          for (int c = 0; c < KC; ++c)
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(d_tmem, a_desc + ((c * C::A_BYTES + k * 32) >> 4), b_desc + ((c * C::B_BYTES + k * 32) >> 4), idesc,
                      (ki | c | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int tn, t, h0, w0;
      decode(tile, tn, t, h0, w0);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      if (p.norm_gamma != nullptr) store_rows_norm(p, t, h0, w0, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN_MAX, quad, lane);
      else store_rows(p, t, h0, w0, tn, tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN_MAX, quad, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------
// conv3d_tb_kernel: the same implicit GEMM with TEMPORAL BLOCKING for kernels with KT > 1.
//
// conv3d_kernel above fetches, per output frame and tap, one activation window AND one weight slice from L2.  At N = 96 (the
// VAE's full-resolution stage) that is 43 KB per 2.36 MFLOP = 55 FLOP per L2 byte, and the chip's L2 delivers ~6300 B/clk
// (42.6 B/clk/SM): 192 FLOP/B would be needed to keep the tensor pipe busy -- the kernel runs at 29 % of it, L2-throughput
// bound (ncu: 11.6 TB/s of lts traffic, profiles/r02_conv_tb_ab.log).  Here ONE CTA owns TB consecutive output frames of a
// pixel tile (TB accumulators of BN columns in TMEM).  The window of input frame f at spatial shift (kh, kw) is the A operand
// of tap kt for output frame f + (KT-1) - kt: it is loaded ONCE and multiplied by the KT weight slices W[kt, kh, kw], which
// stay resident for the whole frame block (loop order: (kh, kw) -> channel group -> input frame -> kt).  Loads per block:
// KH*KW*(TB+KT-1) windows + KH*KW*KT weight slices instead of TB*KT*KH*KW of each: 139 FLOP/B at TB = 4, N = 96.  Input
// frames left of the causal boundary are skipped (the plain kernel multiplies TMA zero fill).
// Latency: a stage holds KC channel chunks (KC = 3: all 96 channels of a window, 24 KB, ~1000 tensor cycles), 4 stages; the
// weight buffer holds the KT x KC slices of one (kh, kw) (54 KB, ~6000 tensor cycles), double buffered -- with one chunk per
// buffer the refill (one iteration = 0.3 us ahead) was exposed on every iteration: 68 us per block instead of 16.
// Accumulators are released one by one (acc_empty[o]): the next block's MMAs into accumulator o only wait for the epilogue
// of o, so the epilogue (o = 0, 1, ...) runs under the next block's first windows (which feed o = 0, then 0-1, ...).
// ---------------------------------------------------------------------------------------------------------
constexpr int TB_KT_MAX = 3;
constexpr int TB_MAX = 4;
template <int BK, int KC>
struct CfgTB {
  static constexpr int A_BYTES = BM * BK * 2;                     // one channel chunk of a window
  static constexpr int A_STAGE_BYTES = KC * A_BYTES;
  static constexpr int A_STAGES = (KC > 1) ? 4 : ((BK == 64) ? 4 : 8);
  static constexpr int BN_LIMIT = (KC > 1) ? 96 : ((BK == 64) ? 192 : 256);   // weight slice rows that fit the buffers below
  static constexpr int B_TILE_MAX = BN_LIMIT * BK * 2;            // one [BN x BK] weight slice
  static constexpr int B_BUF_BYTES = TB_KT_MAX * KC * B_TILE_MAX; // the KT x KC slices of one (kh, kw, channel group)
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + 2 * B_BUF_BYTES + 1024 + 256;
};

// store_rows with the loads hoisted: all tcgen05.ld of the accumulator row and the residual loads are issued before the
// first use (store_rows waits for each 16-column piece and its residual in turn: ~1000 cycles of exposed latency per piece).
template <int NCH>   // BN / 16
__device__ __forceinline__ void store_rows_ilp(const Params& p, int t, int h0, int w0, int tn, uint32_t t_addr, int quad, int lane) {
  const int64_t frame_px = static_cast<int64_t>(p.H) * p.W;
  const int r = quad * 32 + lane;
  const int h = h0 + r / p.BW;
  const int w = w0 + r % p.BW;
  const bool ok = (h < p.H) && (w < p.W);
  const int64_t pix = static_cast<int64_t>(h) * p.W + w;
  const int n_base = tn * p.BN;
  uint32_t acc[NCH][16];
  uint4 res[NCH][2];
#pragma unroll
  for (int c = 0; c < NCH; ++c) tmem_ld_x16(t_addr + c * 16, acc[c]);
  const bool has_res = p.residual != nullptr && ok;
  if (has_res) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (static_cast<int64_t>(t) * frame_px + pix) * p.Cout + n_base);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (p.vec32) {
        ld_global_v8(rp + 2 * c, res[c]);
      } else {
        res[c][0] = rp[2 * c];
        res[c][1] = rp[2 * c + 1];
      }
    }
  }
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int n0 = n_base + c * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[c][j]);
    if (p.bias != nullptr) {
      const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 b = __ldg(bp + q);
        v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
        v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
        v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
        v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
      }
    }
    if (has_res) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 b = res[c][q];
        v[q * 8 + 0] += bf16_lo(b.x); v[q * 8 + 1] += bf16_hi(b.x);
        v[q * 8 + 2] += bf16_lo(b.y); v[q * 8 + 3] += bf16_hi(b.y);
        v[q * 8 + 4] += bf16_lo(b.z); v[q * 8 + 5] += bf16_hi(b.z);
        v[q * 8 + 6] += bf16_lo(b.w); v[q * 8 + 7] += bf16_hi(b.w);
      }
    }
    if (!ok) continue;
    if (p.out_mode == 0) {
      const int g = n0 / p.Csplit;
      const int cdst = n0 - g * p.Csplit;
      const int64_t tf = static_cast<int64_t>(t) * p.out_t_mul + p.out_t_off + g;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (tf * frame_px + pix) * p.Csplit + cdst;
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]);  o0.y = pack_bf16x2(v[2], v[3]);
      o0.z = pack_bf16x2(v[4], v[5]);  o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]);  o1.y = pack_bf16x2(v[10], v[11]);
      o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
      if (p.vec32) {
        // 32 bytes per lane: a full sector (two 16-byte stores per lane wrote every sector in two partial pieces and doubled
        // the L1 -> L2 write transactions: ncu 63.7 M sector writes for 31.9 M sectors of output)
        const uint32_t ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        st_global_v8(op, ov);
      } else {
        reinterpret_cast<uint4*>(op)[0] = o0;
        reinterpret_cast<uint4*>(op)[1] = o1;
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < p.Cvalid)
          op[(static_cast<int64_t>(n0 + j) * p.T + t) * frame_px + pix] = __float2bfloat16(v[j]);
    }
  }
}

template <int BK, int KC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3d_tb_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, Params p) {
  using C = CfgTB<BK, KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + C::A_STAGES * C::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + 2 * C::B_BUF_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::A_STAGES;
  uint64_t* b_full = a_empty + C::A_STAGES;   // [2]
  uint64_t* b_empty = b_full + 2;             // [2]
  uint64_t* acc_full = b_empty + 2;           // [1]
  uint64_t* acc_empty = acc_full + 1;         // [TB_MAX]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + TB_MAX);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < C::A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&b_full[b], 1);
      mbar_init(&b_empty[b], 1);
    }
    mbar_init(acc_full, 1);
    for (int o = 0; o < TB_MAX; ++o) mbar_init(&acc_empty[o], 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int khw_n = p.KH * p.KW;
  const int kgroups = (p.Cin + BK * KC - 1) / (BK * KC);   // a partial last chunk is zero-filled by TMA
  const int tiles_per_frame = p.tiles_h * p.tiles_w;
  const int total_tiles = p.tblocks * tiles_per_frame * p.tiles_n;
  const int b_tile_bytes = p.BN * BK * 2;   // a multiple of 1024
  const int f_n = p.TB + p.KT - 1;          // input frames touched by one frame block

  // tile -> (n tile, frame block, pixel-tile origin); n fastest so that neighbouring CTAs share the activation windows
  auto decode = [&](int tile, int& tn, int& tb0, int& h0, int& w0) {
    tn = tile % p.tiles_n;
    int m = tile / p.tiles_n;
    const int tb = m / tiles_per_frame;
    int r = m - tb * tiles_per_frame;
    tb0 = tb * p.TB;
    h0 = (r / p.tiles_w) * p.BH;
    w0 = (r % p.tiles_w) * p.BW;
  };
  // input frame tin exists in the tensor the TMA map describes
  auto frame_present = [&](int tin) { return tin >= -p.pad_t && tin < p.T; };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, bbuf = 0;
      uint32_t phase = 0, bphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int tn, tb0, h0, w0;
        decode(tile, tn, tb0, h0, w0);
        for (int khw = 0; khw < khw_n; ++khw) {
          const int kh = khw / p.KW, kw = khw - kh * p.KW;
          const int hi = h0 + kh - p.KH / 2 + p.pad_h;
          const int wi = w0 + kw - p.KW / 2 + p.pad_w;
          for (int kg = 0; kg < kgroups; ++kg) {
            // the KT x KC weight slices of (kh, kw, channel group)
            mbar_wait(&b_empty[bbuf], bphase ^ 1);
            mbar_arrive_expect_tx(&b_full[bbuf], static_cast<uint32_t>(p.KT * KC * b_tile_bytes));
            for (int kt = 0; kt < p.KT; ++kt)
#pragma unroll
              for (int c = 0; c < KC; ++c)
                tma_load_2d(b_smem + bbuf * C::B_BUF_BYTES + (kt * KC + c) * b_tile_bytes, &tmW, &b_full[bbuf],
                            (kg * KC + c) * BK, (kt * khw_n + khw) * p.Cout + tn * p.BN);
            if (++bbuf == 2) { bbuf = 0; bphase ^= 1; }
            for (int f = 0; f < f_n; ++f) {
              const int tin = tb0 + f - (p.KT - 1);
              if (!frame_present(tin)) continue;
              mbar_wait(&a_empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&a_full[stage], C::A_STAGE_BYTES);
#pragma unroll
              for (int c = 0; c < KC; ++c)
                tma_load_4d(a_smem + stage * C::A_STAGE_BYTES + c * C::A_BYTES, &tmX, &a_full[stage], (kg * KC + c) * BK, wi, hi,
                            tin + p.pad_t);
              if (++stage == C::A_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_f32(BM, p.BN, 0);
      int stage = 0, bbuf = 0;
      uint32_t phase = 0, bphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int tn, tb0, h0, w0;
        decode(tile, tn, tb0, h0, w0);
        uint32_t started = 0;   // bit o: accumulator o has received its first MMA of this block (and its release was awaited)
        for (int khw = 0; khw < khw_n; ++khw) {
          for (int kg = 0; kg < kgroups; ++kg) {
            mbar_wait(&b_full[bbuf], bphase);
            tc_fence_after();
            const uint64_t b_desc0 = make_desc<BK>(smem_u32(b_smem + bbuf * C::B_BUF_BYTES));
            for (int f = 0; f < f_n; ++f) {
              const int tin = tb0 + f - (p.KT - 1);
              if (!frame_present(tin)) continue;
              mbar_wait(&a_full[stage], phase);
              tc_fence_after();
              const uint64_t a_desc0 = make_desc<BK>(smem_u32(a_smem + stage * C::A_STAGE_BYTES));
              for (int kt = 0; kt < p.KT; ++kt) {
                const int o = f - kt;   // output frame tb0 + o takes input frame tin through tap kt
                if (o < 0 || o >= p.TB || tb0 + o >= p.T) continue;
                const bool first = ((started >> o) & 1u) == 0;
                if (first) {
                  // the epilogue of the previous block has read accumulator o
                  mbar_wait(&acc_empty[o], (it & 1) ^ 1);
                  tc_fence_after();
                  started |= 1u << o;
                }
                const uint32_t d_tmem = tmem_base + o * p.BN;
                const uint64_t b_desc = b_desc0 + ((kt * KC * b_tile_bytes) >> 4);
#pragma unroll
                for (int c = 0; c < KC; ++c)
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_ss(d_tmem, a_desc0 + ((c * C::A_BYTES + k * 32) >> 4), b_desc + ((c * b_tile_bytes + k * 32) >> 4), idesc,
                            (first && c == 0 && k == 0) ? 0u : 1u);
              }
              umma_commit(&a_empty[stage]);
              if (++stage == C::A_STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&b_empty[bbuf]);
            if (++bbuf == 2) { bbuf = 0; bphase ^= 1; }
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int tn, tb0, h0, w0;
      decode(tile, tn, tb0, h0, w0);
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      for (int o = 0; o < p.TB; ++o) {
        if (tb0 + o < p.T) {
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + o * p.BN;
          if (p.BN == 96) store_rows_ilp<6>(p, tb0 + o, h0, w0, tn, t_addr, quad, lane);
          else if (p.BN == 16) store_rows_ilp<1>(p, tb0 + o, h0, w0, tn, t_addr, quad, lane);
          else store_rows(p, tb0 + o, h0, w0, tn, t_addr, quad, lane);
        }
        // every accumulator is released in every block (unused ones too): the barrier phases stay in step with `it`
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[o]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_elems, const uint32_t* box, bool sw64) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return B200_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return B200_ERR_ALIGN;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_elems[i] * 2;
      if (gstr[i - 1] % 16 != 0) return B200_ERR_ALIGN;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[apex_b200] cuTensorMapEncodeTiled (conv) failed: %d\n", (int)r);
    return B200_ERR_TMAP;
  }
  return B200_OK;
}

template <int BK, int KC = 1>
int launch(const void* x, const void* wt, Params& p, cudaStream_t st) {
  using C = Cfg<BK, KC>;
  CUtensorMap tmX, tmW;
  {
    const uint64_t wi = p.W + 2 * p.pad_w, hi = p.H + 2 * p.pad_h, ti = p.T + p.pad_t;
    uint64_t dims[4] = {(uint64_t)p.Cin, wi, hi, ti};
    uint64_t str[4] = {1, (uint64_t)p.Cin, wi * p.Cin, hi * wi * p.Cin};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    int rc = make_tmap_bf16_sw(&tmX, x, 4, dims, str, box, BK == 32);
    if (rc) return rc;
  }
  {
    const int taps = p.KT * p.KH * p.KW;
    uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)taps * p.Cout};
    uint64_t str[2] = {1, (uint64_t)p.Cin};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)p.BN};
    int rc = make_tmap_bf16_sw(&tmW, wt, 2, dims, str, box, BK == 32);
    if (rc) return rc;
  }
  if (p.TB > 1) {
    using CT = CfgTB<BK, KC>;
    static std::atomic<bool> attr_done_tb[kMaxDevices];   // one array per (BK, KC) instantiation
    if (!once_per_device(attr_done_tb, [] {
          return cudaFuncSetAttribute(conv3d_tb_kernel<BK, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT::SMEM_BYTES) == cudaSuccess;
        }))
      return B200_ERR_LAUNCH;
    const int total = p.tblocks * p.tiles_h * p.tiles_w * p.tiles_n;
    const int grid = total < num_sms() ? total : num_sms();
    conv3d_tb_kernel<BK, KC><<<grid, NUM_THREADS, CT::SMEM_BYTES, st>>>(tmX, tmW, p);
    B200_CHECK_LAUNCH();
    return B200_OK;
  }
  static std::atomic<bool> attr_done[kMaxDevices];   // one array per (BK, KC) instantiation
  if (!once_per_device(attr_done, [] {
        return cudaFuncSetAttribute(conv3d_kernel<BK, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess;
      }))
    return B200_ERR_LAUNCH;
  const int total = p.T * p.tiles_h * p.tiles_w * p.tiles_n;
  const int grid = total < num_sms() ? total : num_sms();
  conv3d_kernel<BK, KC><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tmX, tmW, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace conv
}  // namespace b200

static int conv3d_cl_impl(const void* x, const void* w, const void* bias, const void* residual, void* out, int T,
                          int H, int W, int Cin, int Cout, int KT, int KH, int KW, int out_mode, int out_t_mul,
                          int out_t_off, int c_split, int c_valid, int prepadded, void* stream, const void* norm_gamma = nullptr) {
  using namespace b200;
  using namespace b200::conv;
  if (!x || !w || !out) return B200_ERR_ARG;
  if (T <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return B200_ERR_SHAPE;
  if (KT < 1 || KH < 1 || KW < 1 || !(KH & 1) || !(KW & 1)) return B200_ERR_SHAPE;
  if ((Cin % 32) || (Cout % 16)) return B200_ERR_SHAPE;
  if (out_mode != 0 && out_mode != 1) return B200_ERR_ARG;
  if (c_split <= 0 || (Cout % c_split) || (c_split % 16)) return B200_ERR_SHAPE;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return B200_ERR_ALIGN;
  if (residual && ((reinterpret_cast<uintptr_t>(residual) & 15) || c_split != Cout)) return B200_ERR_ALIGN;
  if (reinterpret_cast<uintptr_t>(out) & 15) return B200_ERR_ALIGN;
  Params p;
  p.T = T; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.KT = KT; p.KH = KH; p.KW = KW;
  int bw = 1;
  while (bw * 2 <= W && bw * 2 <= BM) bw *= 2;
  p.BW = bw;
  p.BH = BM / bw;
  // N tile: the largest divisor of Cout that is a multiple of 16, at most 256 and inside one interleave group
  int bn = 0;
  for (int cand = 256; cand >= 16; cand -= 16)
    if (Cout % cand == 0 && c_split % cand == 0) { bn = cand; break; }
  if (bn == 0) return B200_ERR_SHAPE;
  p.BN = bn;
  p.tiles_w = (W + p.BW - 1) / p.BW;
  p.tiles_h = (H + p.BH - 1) / p.BH;
  p.tiles_n = Cout / bn;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.norm_gamma = reinterpret_cast<const __nv_bfloat16*>(norm_gamma);
  if (norm_gamma != nullptr) {
    // the fused norm needs a pixel's whole channel vector in one accumulator row and the plain channels-last store
    if (p.tiles_n != 1 || residual != nullptr || out_mode != 0 || c_split != Cout || out_t_mul != 1 || out_t_off != 0) return B200_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(norm_gamma) & 15) return B200_ERR_ALIGN;
  }
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = out;
  p.out_mode = out_mode;
  p.out_t_mul = out_t_mul; p.out_t_off = out_t_off;
  p.Csplit = c_split; p.Cvalid = c_valid;
  p.vec32 = ((reinterpret_cast<uintptr_t>(out) & 31) == 0) && (c_split % 16 == 0) &&
            (!residual || (reinterpret_cast<uintptr_t>(residual) & 31) == 0);
  p.pad_t = prepadded ? KT - 1 : 0;
  p.pad_h = prepadded ? KH / 2 : 0;
  p.pad_w = prepadded ? KW / 2 : 0;
  // Kernel form.  BK = 64 for Cin % 64 == 0, else BK = 32 (Cin = 96: KC = 3 chunks per stage).
  // B200_CONV_TB = n >= 2: temporal blocking (conv3d_tb_kernel) for every eligible shape with TB <= n.  OFF by default: it
  // halves the L2 traffic of the N = 96 stage (ncu: 49 -> 20 GB per launch) but does not beat the one-frame kernel
  // (3.80 vs 3.75 ms at 81 x 256 x 256 x 96 -> 96; 3.03 vs 1.91 ms at N = 192) -- with one accumulator set per block the
  // epilogue of four frames and the short first/last windows of every (kh, kw) expose latency the one-frame kernel hides
  // (profiles/r02_conv_*_ab.log).  B200_CONV_KC = 1: one chunk per stage.  B200_CONV_PAD64 = 1: Cin = 96 as two 64-channel
  // chunks whose upper half is TMA zero fill (128-byte rows, 33 % more MMA work; measured slower than KC = 3).
  static int tb_mode = -2, kc_mode = -1, pad64 = -1;
  if (tb_mode == -2) {
    const char* ev = getenv("B200_CONV_TB");
    tb_mode = ev ? atoi(ev) : -1;
    ev = getenv("B200_CONV_KC");
    kc_mode = ev ? atoi(ev) : 3;
    ev = getenv("B200_CONV_PAD64");
    pad64 = ev ? atoi(ev) : 0;
  }
  const int bk = (Cin % 64 == 0 || (pad64 && Cin > 64)) ? 64 : 32;
  const int kc = (bk == 32 && kc_mode == 3 && (Cin / 32) % 3 == 0) ? 3 : 1;
  const int bn_limit_tb = bk == 64 ? CfgTB<64, 1>::BN_LIMIT : (kc == 3 ? CfgTB<32, 3>::BN_LIMIT : CfgTB<32, 1>::BN_LIMIT);
  int tb = 512 / bn;
  const int tb_cap = tb_mode < 0 ? 0 : (tb_mode > TB_MAX ? TB_MAX : tb_mode);
  if (tb > tb_cap) tb = tb_cap;
  if (tb > T) tb = T;
  p.TB = (KT >= 2 && KT <= TB_KT_MAX && bn <= bn_limit_tb && tb >= 2 && norm_gamma == nullptr) ? tb : 1;
  p.tblocks = (T + p.TB - 1) / p.TB;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bk == 64) return launch<64, 1>(x, w, p, st);
  if (kc == 3 && (p.TB > 1 || bn <= Cfg<32, 3>::BN_LIMIT)) return launch<32, 3>(x, w, p, st);
  return launch<32, 1>(x, w, p, st);
}

extern "C" int b200_conv3d_cl(const void* x, const void* w, const void* bias, const void* residual, void* out, int T,
                              int H, int W, int Cin, int Cout, int KT, int KH, int KW, int out_mode, int out_t_mul,
                              int out_t_off, int c_split, int c_valid, void* stream) {
  return conv3d_cl_impl(x, w, bias, residual, out, T, H, W, Cin, Cout, KT, KH, KW, out_mode, out_t_mul, out_t_off, c_split,
                        c_valid, 0, stream);
}

extern "C" int b200_conv3d_cl_norm_silu(const void* x, const void* w, const void* bias, const void* gamma, void* out, int T,
                                        int H, int W, int Cin, int Cout, int KT, int KH, int KW, void* stream) {
  if (!gamma) return B200_ERR_ARG;
  return conv3d_cl_impl(x, w, bias, nullptr, out, T, H, W, Cin, Cout, KT, KH, KW, 0, 1, 0, Cout, Cout, 0, stream, gamma);
}

extern "C" int b200_conv3d_cl_padded(const void* x_padded, const void* w, const void* bias, const void* residual, void* out,
                                     int T, int H, int W, int Cin, int Cout, int KT, int KH, int KW, int out_mode,
                                     int c_valid, void* stream) {
  return conv3d_cl_impl(x_padded, w, bias, residual, out, T, H, W, Cin, Cout, KT, KH, KW, out_mode, 1, 0, Cout, c_valid, 1,
                        stream);
}

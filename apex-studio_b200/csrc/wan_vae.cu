// b200_wan_vae_decode: the whole Wan 3D-VAE decoder (post_quant_conv + WanDecoder3d, non-residual, all frames of one
// latent tile) issued from ONE C call -- the launch sequence of AutoencoderKLWan._decode / WanDecoder3d.forward
// (vae/wan/model.py:1333-1375, 972-1021; WanResidualBlock :389-441, WanAttentionBlock :461-490, WanResample :291-353)
// on the kernels of this library (conv.cu, vae_ops.cu, linear.cu).  ~130 launches per 32 x 32 x 21 tile (197 before conv1 + norm2 became one kernel and the mid-attention projections ran once over all frames), no allocation:
// every intermediate lives in a caller-supplied workspace, handed out by a two-region ping-pong arena (a block reads its
// input from one region and builds its output + temporaries in the other), sized by b200_wan_vae_decode_workspace.
//
// The arithmetic, kernel by kernel and in the same order, is that of apex-studio_b200/vae/wan.py::decode_tile (the Python
// host path that issues the same entry points one by one): results are bit-identical (tests/test_gpu_vae.py).
#include "host_util.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

struct Arena {
  char* base;
  int64_t size, top;
  int64_t peak;
  void* take(int64_t bytes) {
    const int64_t a = (top + 255) & ~int64_t(255);
    if (base != nullptr && a + bytes > size) return nullptr;
    top = a + bytes;
    if (top > peak) peak = top;
    return base ? base + a : reinterpret_cast<void*>(1);   // dry run (sizing): any non-null value
  }
};

struct Ctx {
  Arena reg[2];
  int cur;          // region that holds the CURRENT activation
  bool dry;         // sizing pass: no launches
  void* stream;
  int rc;
  Arena& in() { return reg[cur]; }
  Arena& out() { return reg[cur ^ 1]; }
  // start building the next activation in the other region
  bool noreuse;     // debugging aid (B200_VAE_NOREUSE=1): never hand a byte out twice
  void begin_block() { if (!noreuse) out().top = 0; }
  void end_block() { cur ^= 1; }
};

#define VAE_TRY(expr)                       \
  do {                                      \
    if (!c.dry && c.rc == B200_OK) {        \
      const int _rc = (expr);               \
      if (_rc != B200_OK) c.rc = _rc;       \
    }                                       \
  } while (0)

inline int64_t bf16_bytes(int64_t n) { return n * 2; }

// conv1 -> norm2 -> SiLU of a residual block as ONE kernel (b200_conv3d_cl_norm_silu) when the conv keeps a pixel's channel
// vector in one N tile: Cout <= 256 (96, 192; the 384-channel stage runs two N tiles of 192).  B200_VAE_FUSE_NORM=0: never
// (the A/B partner; vae/wan.py applies the same rule so that both host paths stay bit-identical).
inline bool conv_norm_fusable(int cout) {
  static int mode = -1;
  if (mode < 0) {
    const char* ev = getenv("B200_VAE_FUSE_NORM");
    mode = (ev && ev[0] == '0') ? 0 : 1;
  }
  return mode == 1 && cout <= 256 && (cout % 16) == 0;
}

// y = conv2(silu(norm2(conv1(silu(norm1(x)))))) + shortcut(x)      (WanResidualBlock)
void* res_block(Ctx& c, const void* x, const B200WanResBlock& w, int T, int H, int W) {
  const int64_t px = static_cast<int64_t>(T) * H * W;
  c.begin_block();
  Arena& a = c.out();
  void* out = a.take(bf16_bytes(px * w.cout));
  const void* h = x;
  if (w.shortcut_w != nullptr) {
    void* hs = a.take(bf16_bytes(px * w.cout));
    if (!hs) { c.rc = B200_ERR_ARG; return nullptr; }
    VAE_TRY(b200_linear(x, w.shortcut_w, w.shortcut_b, hs, nullptr, static_cast<int>(px), w.cout, w.cin, w.cin, w.cin, w.cout,
                        B200_EPI_BIAS, c.stream));
    h = hs;
  }
  void* n = a.take(bf16_bytes(px * w.cin));
  void* y = a.take(bf16_bytes(px * w.cout));
  if (!out || !n || !y) { c.rc = B200_ERR_ARG; return nullptr; }
  VAE_TRY(b200_rmsnorm_silu_cl(x, n, w.norm1_gamma, px, w.cin, 1, c.stream));
  if (conv_norm_fusable(w.cout)) {   // conv1 -> norm2 -> SiLU in one kernel (the conv output has no other consumer)
    VAE_TRY(b200_conv3d_cl_norm_silu(n, w.conv1_w, w.conv1_b, w.norm2_gamma, y, T, H, W, w.cin, w.cout, 3, 3, 3, c.stream));
  } else {
    VAE_TRY(b200_conv3d_cl(n, w.conv1_w, w.conv1_b, nullptr, y, T, H, W, w.cin, w.cout, 3, 3, 3, 0, 1, 0, w.cout, w.cout, c.stream));
    VAE_TRY(b200_rmsnorm_silu_cl(y, y, w.norm2_gamma, px, w.cout, 1, c.stream));
  }
  VAE_TRY(b200_conv3d_cl(y, w.conv2_w, w.conv2_b, h, out, T, H, W, w.cout, w.cout, 3, 3, 3, 0, 1, 0, w.cout, w.cout, c.stream));
  c.end_block();
  return out;
}

// single-head attention over the H*W positions of each frame, x += proj(attn)   (WanAttentionBlock), in place on x
void attn_block(Ctx& c, void* x, const B200WanAttn& w, int T, int H, int W) {
  const int C = w.channels;
  const int N = H * W;
  const int64_t px = static_cast<int64_t>(T) * N;
  c.begin_block();
  Arena& a = c.out();
  // The three weight GEMMs do not mix frames: q | k, V^T and the output projection run ONCE over all T frames (M = T * N rows)
  // instead of once per frame; only the scores and P V are per frame.  66 launches instead of 126, and the projections fill the
  // GPU (per frame they were 16-CTA grids).  Element-wise identical to the per-frame form (same K loop per output element).
  void* xn = a.take(bf16_bytes(px * C));
  void* qk = a.take(bf16_bytes(px * 2 * C));                       // [T*N, 2C]
  void* vT = a.take(bf16_bytes(static_cast<int64_t>(C) * px));     // [C, T*N]  (V^T of every frame side by side)
  void* s = a.take(static_cast<int64_t>(N) * N * 4);
  void* pm = a.take(bf16_bytes(static_cast<int64_t>(N) * N));
  void* o = a.take(bf16_bytes(px * C));                            // [T*N, C]
  if (!xn || !qk || !vT || !s || !pm || !o) { c.rc = B200_ERR_ARG; return; }
  VAE_TRY(b200_rmsnorm_silu_cl(x, xn, w.norm_gamma, px, C, 0, c.stream));
  const char* wq = static_cast<const char*>(w.to_qkv_w);
  const char* bq = static_cast<const char*>(w.to_qkv_b);
  const float scale = static_cast<float>(pow(static_cast<double>(C), -0.5));   // C ** -0.5 in double, rounded once (as the host path does)
  const int M = static_cast<int>(px);
  // q | k = xn Wqk^T + b  [T*N, 2C];  V^T = Wv xn^T + b (per-row bias)  [C, T*N]
  VAE_TRY(b200_linear(xn, wq, bq, qk, nullptr, M, 2 * C, C, C, C, 2 * C, B200_EPI_BIAS, c.stream));
  VAE_TRY(b200_linear(wq + bf16_bytes(static_cast<int64_t>(2) * C * C), xn, bq + bf16_bytes(2 * C), vT, nullptr, C, M, C, C, C, M,
                      B200_EPI_BIAS | B200_EPI_ROW_BIAS, c.stream));
  for (int t = 0; t < T; ++t) {
    const char* qt = static_cast<const char*>(qk) + bf16_bytes(static_cast<int64_t>(t) * N * 2 * C);
    const char* vt = static_cast<const char*>(vT) + bf16_bytes(static_cast<int64_t>(t) * N);
    char* ot = static_cast<char*>(o) + bf16_bytes(static_cast<int64_t>(t) * N * C);
    // scores = q k^T (fp32), P = softmax(scores / sqrt(C)), o = P V
    VAE_TRY(b200_linear(qt, qt + bf16_bytes(C), nullptr, s, nullptr, N, N, C, 2 * C, 2 * C, N, B200_EPI_BIAS_F32, c.stream));
    VAE_TRY(b200_softmax_rows(static_cast<const float*>(s), pm, N, N, N, N, scale, c.stream));
    VAE_TRY(b200_linear(pm, vt, nullptr, ot, nullptr, N, C, N, N, M, C, B200_EPI_BIAS, c.stream));
  }
  // x += proj(o) for all frames
  VAE_TRY(b200_linear(o, w.proj_w, w.proj_b, x, nullptr, M, C, C, C, C, C, B200_EPI_GATE_RES, c.stream));
  // x stays where it is: the block was in place, the other region only held temporaries
}

// WanResample upsample2d / upsample3d: (time_conv + 2x temporal interleave, first frame bypasses) -> nearest 2x -> 3x3 conv C -> C/2
void* upsample(Ctx& c, const void* x, const B200WanUpsample& w, int& T, int& H, int& W) {
  const int C = w.channels;
  c.begin_block();
  Arena& a = c.out();
  const int64_t frame = static_cast<int64_t>(H) * W * C;
  int T2 = T;
  if (w.temporal && T > 1) T2 = 1 + 2 * (T - 1);
  void* out = a.take(bf16_bytes(static_cast<int64_t>(T2) * (2 * H) * (2 * W) * (C / 2)));
  const void* xt = x;
  if (T2 != T) {
    void* y = a.take(bf16_bytes(static_cast<int64_t>(T2) * frame));
    if (!y) { c.rc = B200_ERR_ARG; return nullptr; }
    if (!c.dry && c.rc == B200_OK) {
      if (cudaMemcpyAsync(y, x, bf16_bytes(frame), cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(c.stream)) != cudaSuccess)
        c.rc = B200_ERR_LAUNCH;
    }
    VAE_TRY(b200_conv3d_cl(static_cast<const char*>(x) + bf16_bytes(frame), w.time_conv_w, w.time_conv_b, nullptr, y, T - 1, H, W, C,
                           2 * C, 3, 1, 1, 0, 2, 1, C, 2 * C, c.stream));
    xt = y;
  }
  void* u = a.take(bf16_bytes(static_cast<int64_t>(T2) * (2 * H) * (2 * W) * C));
  if (!out || !u) { c.rc = B200_ERR_ARG; return nullptr; }
  VAE_TRY(b200_upsample2x_cl(xt, u, T2, H, W, C, c.stream));
  VAE_TRY(b200_conv3d_cl(u, w.resample_w, w.resample_b, nullptr, out, T2, 2 * H, 2 * W, C, C / 2, 1, 3, 3, 0, 1, 0, C / 2, C / 2,
                         c.stream));
  T = T2;
  H *= 2;
  W *= 2;
  c.end_block();
  return out;
}

int run(Ctx& c, const void* z_cl, const B200WanVaeWeights& w, void* out, int T, int h, int wd) {
  int H = h, W = wd;
  const int64_t px0 = static_cast<int64_t>(T) * H * W;
  // post_quant_conv (1x1x1) on the 64-channel padded latent, then conv_in
  c.begin_block();
  void* x = c.out().take(bf16_bytes(px0 * w.z_pad));
  if (!x) return B200_ERR_ARG;
  VAE_TRY(b200_linear(z_cl, w.post_quant_w, w.post_quant_b, x, nullptr, static_cast<int>(px0), w.z_pad, 64, 64, 64, w.z_pad, B200_EPI_BIAS,
                      c.stream));
  c.end_block();
  c.begin_block();
  void* y = c.out().take(bf16_bytes(px0 * w.dims[0]));
  if (!y) return B200_ERR_ARG;
  VAE_TRY(b200_conv3d_cl(x, w.conv_in_w, w.conv_in_b, nullptr, y, T, H, W, w.z_pad, w.dims[0], 3, 3, 3, 0, 1, 0, w.dims[0], w.dims[0],
                         c.stream));
  c.end_block();
  x = y;
  x = res_block(c, x, w.mid_res[0], T, H, W);
  if (!x) return c.rc ? c.rc : B200_ERR_ARG;
  attn_block(c, x, w.mid_attn, T, H, W);
  x = res_block(c, x, w.mid_res[1], T, H, W);
  for (int i = 0; i < 4 && x; ++i) {
    for (int j = 0; j < 3 && x; ++j) x = res_block(c, x, w.up_res[i][j], T, H, W);
    if (i != 3 && x) x = upsample(c, x, w.up_samp[i], T, H, W);
  }
  if (!x) return c.rc ? c.rc : B200_ERR_ARG;
  const int64_t px = static_cast<int64_t>(T) * H * W;
  VAE_TRY(b200_rmsnorm_silu_cl(x, x, w.norm_out_gamma, px, w.dims[4], 1, c.stream));
  VAE_TRY(b200_conv3d_cl(x, w.conv_out_w, w.conv_out_b, nullptr, out, T, H, W, w.dims[4], 16, 3, 3, 3, 1, 1, 0, 16, 3, c.stream));
  return c.rc;
}

}  // namespace

extern "C" int64_t b200_wan_vae_decode_workspace(const B200WanVaeWeights* w, int T, int h, int wd) {
  if (!w || T <= 0 || h <= 0 || wd <= 0) return B200_ERR_ARG;
  Ctx c = {};
  c.dry = true;
  c.noreuse = getenv("B200_VAE_NOREUSE") != nullptr;
  const int rc = run(c, reinterpret_cast<const void*>(1), *w, reinterpret_cast<void*>(1), T, h, wd);
  if (rc != B200_OK) return rc;
  const int64_t region = ((c.reg[0].peak > c.reg[1].peak ? c.reg[0].peak : c.reg[1].peak) + 255) & ~int64_t(255);
  return 2 * region;
}

extern "C" int b200_wan_vae_decode(const void* z_cl, const B200WanVaeWeights* w, void* out, void* workspace, int64_t workspace_bytes,
                                   int T, int h, int wd, void* stream) {
  if (!z_cl || !w || !out || !workspace) return B200_ERR_ARG;
  if (T <= 0 || h <= 0 || wd <= 0) return B200_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return B200_ERR_ALIGN;
  const int64_t need = b200_wan_vae_decode_workspace(w, T, h, wd);
  if (need < 0) return static_cast<int>(need);
  if (workspace_bytes < need) return B200_ERR_ARG;
  Ctx c = {};
  c.stream = stream;
  c.noreuse = getenv("B200_VAE_NOREUSE") != nullptr;
  const int64_t region = need / 2;
  c.reg[0].base = static_cast<char*>(workspace);
  c.reg[1].base = static_cast<char*>(workspace) + region;
  c.reg[0].size = c.reg[1].size = region;
  return run(c, z_cl, *w, out, T, h, wd);
}

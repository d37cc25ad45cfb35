// Library identity + error strings of the C ABI (include/apex_b200.h).
#include "../../include/apex_b200.h"

extern "C" int b200_version(void) { return 100; }

extern "C" const char* b200_strerror(int code) {
  switch (code) {
    case B200_OK: return "ok";
    case B200_ERR_SHAPE: return "unsupported shape";
    case B200_ERR_ALIGN: return "pointer or stride is not 16-byte aligned";
    case B200_ERR_DRIVER: return "CUDA driver entry point unavailable (no GPU / driver?)";
    case B200_ERR_TMAP: return "cuTensorMapEncodeTiled rejected the tensor layout";
    case B200_ERR_LAUNCH: return "kernel launch failed";
    case B200_ERR_ARG: return "null pointer or invalid enum argument";
    default: return "unknown error";
  }
}

// ---------------------------------------------------------------------------------------------------------
// Composite entry points: the granularity SURVEY.md section 8(b) lists for the reference's call sites.  Each one
// enqueues the kernels of one call site on `stream`, back to back; the caller supplies outputs and workspace.
// ---------------------------------------------------------------------------------------------------------
extern "C" int b200_ln_modulate(const void* x, const void* scale, const void* shift, void* y, int rows, int dim,
                                float eps, void* stream) {
  if (!scale || !shift) return B200_ERR_ARG;
  return b200_layernorm_modulate(x, y, scale, shift, nullptr, nullptr, rows, dim, dim, dim, 0, eps, stream);
}

extern "C" int b200_qkv_rmsnorm_rope(const void* x, const void* w_qkv, const void* b_qkv, const void* wq_norm,
                                     const void* wk_norm, const void* rope, void* qkv, int rows, int dim, int heads,
                                     int64_t ldx, float eps, void* stream) {
  if (!x || !w_qkv || !qkv) return B200_ERR_ARG;
  if (heads <= 0 || dim % heads) return B200_ERR_SHAPE;
  int rc = b200_linear(x, w_qkv, b_qkv, qkv, nullptr, rows, 3 * dim, dim, ldx, dim, 3 * (int64_t)dim, B200_EPI_BIAS,
                       stream);
  if (rc) return rc;
  char* base = static_cast<char*>(qkv);
  // q and k in ONE launch when their norm weights are stacked ([2, dim] contiguous, as the model stores them)
  if (wq_norm && wk_norm && static_cast<const char*>(wk_norm) == static_cast<const char*>(wq_norm) + 2 * (int64_t)dim)
    return b200_rmsnorm_rope_batched(base, wq_norm, rope, rows, heads, dim / heads, 3 * (int64_t)dim, eps, 2, dim, dim, stream);
  rc = b200_rmsnorm_rope(base, wq_norm, rope, rows, heads, dim / heads, 3 * (int64_t)dim, eps, stream);
  if (rc) return rc;
  return b200_rmsnorm_rope(base + 2 * (int64_t)dim, wk_norm, rope, rows, heads, dim / heads, 3 * (int64_t)dim, eps,
                           stream);
}

extern "C" int b200_mlp_gelu(const void* x, const void* w1, const void* b1, const void* w2, const void* b2,
                             const void* gate, void* h, void* workspace, int rows, int dim, int ffn_dim,
                             void* stream) {
  if (!x || !w1 || !w2 || !h || !workspace) return B200_ERR_ARG;
  int rc = b200_linear(x, w1, b1, workspace, nullptr, rows, ffn_dim, dim, dim, dim, ffn_dim, B200_EPI_GELU_TANH, stream);
  if (rc) return rc;
  return b200_linear(workspace, w2, b2, h, gate, rows, dim, ffn_dim, ffn_dim, ffn_dim, dim, B200_EPI_GATE_RES, stream);
}

// b200_attn_fwd: o = softmax(q k^T * scale) v  (non-causal, no mask), head_dim 128, bf16 in/out.
//
// One CTA owns TWO 128-row query tiles of one (batch, head) and streams all key/value tiles past them:
//   warps 0-3   softmax for query tile 0 (one query row per thread, row == TMEM lane)
//   warps 4-7   softmax for query tile 1
//   warp  8     TMA producer: Q once, then K_j, V_j through a ring of shared-memory slots (128B-swizzled tiles)
//   warp  9     MMA issuer:   S_t = Q_t K_j^T  (tcgen05.mma, A and B from shared memory, fp32 S in TMEM)
//                             O_t += P_t V_j   (A = bf16 P read straight from TMEM, B = V MN-major in smem)
//
// PIPE = 1 (default, round 2) -- the DECOUPLED pipeline.  TMEM (512 columns): S | P0 | P1 | O0 | O1 = 128 + 64 + 64 +
//   128 + 128.  ONE S buffer is shared by both query tiles: a softmax warp pulls its whole 128-column S row into
//   registers in ~50 cycles and hands the buffer back ("S free"), so the tensor pipe computes S_t(j+1) WHILE softmax_t(j)
//   is still working in registers, and P_t has its own columns.  Issue order  S1(j+1) PV0(j) S0(j+2) PV1(j).  The round-1
//   layout (S0 | S1 | O0 | O1, P_t aliasing S_t) forced S_t(j+1) behind PV_t(j), which put the whole chain
//   PV_t(j) -> S_t(j+1) -> softmax_t(j+1) -> "P ready" in series: 512 + 512 + ~1700 + ~290 = 3036 cycles per key tile against
//   2048 cycles of tensor work (measured with b200_attn_fwd_prof: period 3084, profiles/r02_attn_timeline_pipe0.log).
//   Decoupled, the period is 2425 cycles, bound by the hand-over of the single S buffer between the tiles (2 x [S MMAs
//   512 + "S ready" -> pull ~250 + "S free" -> issue ~300 + queueing behind a PV]).  Hazards: (1) P_t(j+1) may only be stored
//   after PV_t(j) has read P_t(j) -> the softmax warps wait for "O done"(j) right before their P stores (it has long
//   completed in steady state); (2) the lazy rescale of O_t (only when the running max grows by more than 2^8) waits for
//   the same barrier.  Measured (40 heads x 75,600^2, sustained, same box): 1335 TFLOP/s vs 1237 (PIPE 0) vs cuDNN SDPA 1372;
//   the SM clock settles at 1.30 GHz instead of 1.54 GHz -- the 1 kW power cap converts most of the recovered cycles into a
//   lower clock (profiles/r02_attn_pipe1_sharedS_ab.log, r02_attn_roles_ab.log).
//   Tried and rejected, with logs under profiles/: (a) S as two N = 64 MMA groups (64-key double buffering): operand fetch
//   starves the pipe, r02_attn_exp_n64.log; (b) S two key tiles ahead with P aliasing S and the softmax pulling S(j+1)
//   before storing P(j) (no hand-back barrier): the extra pull/store chain lengthens every softmax step, 1248-1278 TFLOP/s,
//   r02_attn_designC_*.log; (c) one issuing thread per MMA stream: the pipe interleaves them and delays S, 1195,
//   r02_attn_dual_issuer.log; (d) more / fewer polynomial exp2 (12.5 % .. 37.5 %): within 1 %, r02_attn_designC_poly_sweep.log.
// PIPE = 0 -- the round-1 pipeline (1-CTA only), kept as the A/B partner (B200_ATTN_PIPE=0).
//
// NCTA = 2 (opt-in, B200_ATTN_2CTA=1): the CTAs work in PAIRS (cluster of 2 on one TPC, tcgen05 cta_group::2).  A pair owns
// 512 query rows; every MMA is M = 256 (128 rows in each CTA's TMEM).  The B operands are SPLIT across the pair: each CTA
// stages only 64 of the 128 keys of K_j (S = Q K^T: N = 128 keys, N/2 per CTA) and only 64 of the 128 channels of V_j (O += P V:
// N = 128 channels, N/2 per CTA), so every K/V byte is fetched from L2 and written to shared memory once per 512 query rows
// instead of once per 256, and the operand reads of the tensor pipe drop from 192 KB to 128 KB per key tile per SM.  Both
// CTAs' TMA bytes complete on the LEADER's mbarrier; the leader's single MMA thread issues for both;
// tcgen05.commit.multicast publishes "S ready", "O done" and "slot free" in both CTAs; the peer's softmax warps signal
// "S free" / "P ready" on the leader's barriers through remote (shared::cluster) arrives.  Correct on every test shape, but
// SLOWER (1029-1080 TFLOP/s): the remote arrives and multicast commits add ~400 cycles to hand-overs that sit on the
// critical path of this kernel (unlike the GEMM, where the pair gained 10 %), so single CTAs stay the default.
//
// LAUNCH FORMS of the decoupled 1-CTA pipeline (host side of attn_fwd_impl; work item = (256 query rows, head, batch)):
//   long key sequences  (> MULTI_MAX_KV_TILES key tiles; Wan self-attention: 591): one work item per CTA, 5-slot K/V ring,
//                        epilogue = 32-byte direct stores.  The per-CTA prologue is amortised; the item loops fold away (MULTI = false).
//   mid                 (<= 128 key tiles; FLUX / QwenImage joint attention: 36 / 68): PERSISTENT grid, one CTA per SM walks the
//                        items: barriers, TMEM and descriptors are set up once; the next item's Q / K_0 / K_1 loads and its S_0(0)
//                        MMAs run under the current epilogue ("Q empty" after the item's last S MMA, "O free" before its first PV).
//   short               (<= STAGE_MAX_KV_TILES; Wan cross-attention: 4): persistent + STAGED epilogue -- O goes through shared
//                        memory (two 128B-swizzled 4 KB boxes per softmax warp) and out by TMA bulk stores; the 64 KB staging area
//                        replaces two ring slots.  One thread owns one output row, so direct stores touch 32 lines per instruction:
//                        5,000-7,000 cycles per item, a third of a 4-key-tile item (profiles/r02_attn_cross_timeline.log).
//   Sequence-parallel scatter epilogues (b200_attn_fwd_scatter, b200_attn_fwd_scatter_joint) use the direct-store form.
//   Measured (same box): Wan cross-attention 650 -> 1080 TFLOP/s (cuDNN 1100-1200), self-attention 1337 -> 1365-1377 (cuDNN 1366-1388).
//
// Replaces attention_register.call(q, k, v) -- attention/functions.py:84 (`sdpa` :338-377 is the gold
// backend) as called by transformer/wan/base/attention.py:397.
#include "host_util.cuh"
#include <stdlib.h>
#include <type_traits>

#include "sm100_ptx.cuh"

namespace b200 {
namespace attn {

constexpr int D = 128;
constexpr int BQ = 128;   // rows per query tile (per CTA)
constexpr int BKV = 128;  // keys per tile
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two [128 x 64] swizzled half tiles
constexpr int HALF_BYTES = TILE_BYTES / 2;
// warpgroups: softmax0 | softmax1 | {TMA, MMA issuer, 2 idle warps}.  The single-thread roles park their mbarrier waits in
// hardware (mbar_wait_parked) instead of re-polling.  (Roles in warps 0-3 vs 8-11, parked vs polling waits: all four combinations
// measure the same, 1332-1336 TFLOP/s, profiles/r02_attn_roles_ab.log.)
constexpr int NUM_THREADS = 384;
constexpr int ROLE_WG = 2;          // warpgroup of the single-thread roles
constexpr int TMA_WARP = 4 * ROLE_WG, MMA_WARP = 4 * ROLE_WG + 1;
B200_DEVICE void role_wait(uint64_t* bar, uint32_t parity) { mbar_wait_parked(bar, parity); }
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units
// setmaxnreg budget.  The CTA owns 168 regs x 384 threads = 504 per (softmax0, softmax1, other) warp triple; the
// increase BLOCKS until the pool has enough registers, so 2 * REGS_SOFTMAX + REGS_OTHER must not exceed 504
// (216/80 = 512 deadlocked every CTA in bring-up).
constexpr int REGS_SOFTMAX = 208;
constexpr int REGS_OTHER = 88;
constexpr int STAGE_MAX_KV_TILES = 8;     // staged epilogue + 3-slot ring up to this many key tiles (Wan cross-attention: 4)
constexpr int MULTI_MAX_KV_TILES = 128;    // persistent launch up to this many key tiles per work item
constexpr bool PAIR_BY_DEFAULT = false;   // CTA pairs: set once measured faster than single CTAs on the B200
static_assert(2 * REGS_SOFTMAX + REGS_OTHER <= 504, "setmaxnreg.inc would wait forever");

// STAGE (1-CTA, decoupled pipeline): the epilogue stages O in shared memory and writes it with TMA bulk stores instead of
// 16-byte stores per lane.  One thread owns one output ROW (= TMEM lane), so a warp-wide 16-byte store touches 32 different
// 128-byte lines: 512 line accesses per warp and tile, ~6,000 cycles of LSU time per work item (measured with the in-kernel
// timeline: profiles/r02_attn_cross_timeline.log) -- irrelevant next to 591 key tiles, but 1/3 of a 4-key-tile
// cross-attention item.  The staging area (8 warps x two [32 rows x 64 channels] boxes = 64 KB) takes the place of TWO K/V ring
// slots (3 remain: K runs one tile ahead instead of two), so this form is used for short key sequences only -- their K/V
// come from L2.
template <int NCTA, bool STAGE = false>
struct Cfg {
  static constexpr int KV_BYTES = TILE_BYTES / NCTA;   // bytes of one K (or V) tile staged by ONE CTA
  // ring depth: the decoupled pipeline issues S_t(j+2) while V_j is still in use, so K_{j+2} must land about one key tile
  // ahead -- 5 slots of 32 KB (4 stalled the MMA issuer ~800 cycles per key tile on the K wait: profiles/r02_attn_timeline_*.log)
  static constexpr int KV_SLOTS = NCTA == 2 ? 8 : (STAGE ? 3 : 5);
  static constexpr int STAGE_BYTES = STAGE ? 8 * 8192 : 0;
  static constexpr int SMEM_BYTES = 2 * TILE_BYTES + KV_SLOTS * KV_BYTES + STAGE_BYTES + 1024 + 256;
};

struct Params {
  int B, H, Sq, Sk;
  int n_kv;           // key tiles: ceil(Sk / 128)
  int n_qt;           // query tile groups per (batch, head): ceil(Sq / (2 * BQ * NCTA))
  int n_hb;           // H * B
  int n_items;        // n_qt * H * B work items; a CTA (pair) walks items blockIdx.x / NCTA + i * gridDim.x / NCTA
  __nv_bfloat16* o;
  int64_t o_sb, o_sh, o_ss;
  float scale_log2;
  // Sequence-parallel "heads -> tokens" exchange fused into the epilogue: query row r belongs to the rank that owns
  // token r, so its output row is stored straight into that peer's [S/P, H_total*128] buffer over NVLink.
  void* o_peer[8];
  int o_vec32;        // every output row is 32-byte aligned: the direct-store epilogue may use 32-byte stores
  int n_peers;        // 0 = plain store into `o`
  int rows_per_rank;  // S / P
  int head_off;       // first global head computed by this rank
  // Joint (dual-stream) sequences: `rep_rows` query rows -- the replicated text stream -- are stored to EVERY peer; the other
  // rows are token-sharded as above.  rep_first: the replicated rows come first (Flux / QwenImage), else last (HunyuanVideo-1.5).
  // Every peer's buffer is laid out like its local joint sequence: [rep_rows + rows_per_rank, H_total * 128] in the same order.
  int rep_rows, rep_first;
  // b200_attn_fwd_qnorm: the RMS-norm of q over all H * 128 channels of a token, folded into the logits: row r of every head is
  // scaled by rstd[r] = bf16(rsqrt(sum_i q_rowsumsq[(batch * Sq + r) * q_parts + i] * q_inv_dim + q_eps)) (the per-channel norm
  // weight was applied by the projection's epilogue, b200_linear_normw).  NULL = plain attention.
  const float* q_rowsumsq;
  int q_parts;
  float q_inv_dim, q_eps;
  long long* prof;    // PROF kernels only: [steps][16] SM-clock timestamps of CTA (0,0,0) (b200_attn_fwd_prof)
  int prof_steps;
};

// ---------------------------------------------------------------------------------------------------------
// Softmax of one 128-key tile for one query row (thread == row == TMEM lane): one pass -- the whole S row is loaded
// into registers with four back-to-back tcgen05.ld, packed fma/add (.f32x2) for the scale-and-subtract and the row
// sum, and POLY_PAIRS of every 16 column pairs take a Cody-Waite + cubic-polynomial exp2 on the FMA pipe instead of
// MUFU.EX2 (the SFU is the co-bottleneck: 128x128 exps at 16/clk/SM take as long as the two 128x128x128 MMAs of the tile).
// ---------------------------------------------------------------------------------------------------------
B200_DEVICE float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  float2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
B200_DEVICE float2 fadd2(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  float2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// 2^x for x <= ~9 on the FMA/ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a minimax cubic (max rel err
// 7.6e-5, far below the bf16 rounding of P), exponent added with an integer shift-add.
B200_DEVICE float2 exp2_poly2(float2 x) {
  const float magic = 12582912.0f;  // 1.5 * 2^23: adding it rounds x to an integer in the low mantissa bits
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = fadd2(x, make_float2(magic, magic));
  const float2 n = fadd2(t, make_float2(-magic, -magic));
  const float2 f = fadd2(x, make_float2(-n.x, -n.y));
  // minimax cubic for 2^f on [-0.5, 0.5]: max relative error 7.6e-5 (bf16 rounding of P is 2e-3)
  float2 p = ffma2(f, make_float2(0.055205505f, 0.055205505f), make_float2(0.24261397f, 0.24261397f));
  p = ffma2(p, f, make_float2(0.69325477f, 0.69325477f));
  p = ffma2(p, f, make_float2(0.9999277f, 0.9999277f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

B200_DEVICE float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

#ifndef ATTN_POLY_PAIRS
#define ATTN_POLY_PAIRS 4
#endif
constexpr int POLY_PAIRS = ATTN_POLY_PAIRS;  // of every 16 column pairs (32 columns) -> 25 % of the exps leave the SFU

// row = key-tile step counted over all work items of CTA 0 (n_steps is the role's running step count)
#define ATTN_STAMP(step, k)                                                                                       \
  do {                                                                                                           \
    if (PROF && prof_cta && (n_steps + (step)) < p.prof_steps) p.prof[(n_steps + (step)) * 32 + (k)] = clock64(); \
  } while (0)

// MULTI: the CTA walks several work items (persistent launch).  Without it the item loops run exactly once and the
// cross-item bookkeeping folds away at compile time: the long self-attention launches (591 key tiles per item, prologue
// amortised) keep the leaner single-item issue loop.
template <int NCTA, int PIPE, bool PROF = false, bool STAGE = false, bool MULTI = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, Params p) {
  static_assert(!STAGE || (NCTA == 1 && PIPE == 1), "staged epilogue: 1-CTA decoupled pipeline only");
  static_assert(!MULTI || (NCTA == 1 && PIPE == 1), "persistent launch: 1-CTA decoupled pipeline only");
  constexpr int KV_SLOTS = Cfg<NCTA, STAGE>::KV_SLOTS;
  constexpr int KV_BYTES = Cfg<NCTA, STAGE>::KV_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                      // 2 tiles
  uint8_t* kv_smem = smem + 2 * TILE_BYTES;    // KV_SLOTS slots
  uint8_t* o_stage = kv_smem + KV_SLOTS * KV_BYTES;   // STAGE: two 4 KB boxes per softmax warp (1024-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_stage + Cfg<NCTA, STAGE>::STAGE_BYTES);
  uint64_t* q_full = bars;                 // [1]          (pair: the leader's collects both CTAs' bytes)
  uint64_t* kv_full = bars + 1;            // [KV_SLOTS]   (pair: leader's)
  uint64_t* kv_empty = kv_full + KV_SLOTS; // [KV_SLOTS]   (pair: multicast commit -> both CTAs)
  uint64_t* s_full = kv_empty + KV_SLOTS;  // [2]          (multicast)
  uint64_t* p_full = s_full + 2;           // [2]          (pair: leader's, 4 local + 4 remote warps arrive)
  uint64_t* o_done = p_full + 2;           // [2]          (multicast)
  uint64_t* s_free = o_done + 2;           // [1]  PIPE 1: the shared S buffer has been pulled into registers (pair: leader's)
  uint64_t* q_empty = s_free + 1;          // [1]  persistent: the last S MMA of a work item has read the Q tiles
  uint64_t* o_free = q_empty + 1;          // [2]  persistent: the epilogue of tile t has read O_t out of TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);
  // TMEM columns
  constexpr uint32_t S_COL0 = 0, S_COL1 = PIPE ? 0 : 128;                  // PIPE 1: one S buffer for both tiles
  constexpr uint32_t P_COL0 = PIPE ? 128 : 0, P_COL1 = PIPE ? 192 : 128;   // PIPE 0: P_t aliases S_t
  constexpr uint32_t O_COL0 = 256, O_COL1 = 384;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // (n_kv = p.n_kv, w_step and p.n_items are read from the constant bank where used: as locals computed up here they
  // were spilled across the role split and re-loaded from local memory inside the MMA issuer's loop)
#define n_kv (p.n_kv)
#define w_step (static_cast<int>(gridDim.x / NCTA))
  // Work items (query tile group, head, batch), query tiles fastest so that the CTAs resident at any moment share one
  // head's K/V in L2.  A non-persistent launch has one item per CTA (pair); a persistent launch (1-CTA, PIPE 1) has one
  // CTA per SM walking items w0, w0 + w_step, ...: barriers, TMEM and descriptors are set up once, and the Q / K loads
  // and the first S MMAs of the next item run under the epilogue of the current one.
  const int w0 = static_cast<int>(blockIdx.x / NCTA);
  constexpr int TILE_ROW_STEP = BQ * NCTA;
  struct Item { int head, batch, row_base; };
  // Walking the items without a division per item (decoding w = qt + n_qt * (head + H * batch) cost ~500 cycles at the top
  // of every epilogue): the iterator keeps (qt, hb = head + H * batch) and advances by w_step.
  struct ItemIter { int qt, hb; };
  auto iter_first = [&]() {
    ItemIter it;
    it.qt = w0 % p.n_qt;
    it.hb = w0 / p.n_qt;
    return it;
  };
  auto iter_next = [&](ItemIter& it) {
    it.qt += w_step;
    if (it.qt >= p.n_qt) {
      it.qt -= p.n_qt;
      ++it.hb;
      if (it.qt >= p.n_qt) {   // more than one (batch, head) per stride: few query tiles
        it.hb += it.qt / p.n_qt;
        it.qt %= p.n_qt;
      }
    }
  };
  // row_base: first query row of tile 0 of this CTA (a pair owns 2 * 256 rows, MMA tile t = rows [t * 256, +256) of them)
  auto decode_item = [&](const ItemIter& it) {
    Item wi;
    if (it.hb < p.H) {
      wi.head = it.hb;
      wi.batch = 0;
    } else {
      wi.batch = it.hb / p.H;
      wi.head = it.hb - wi.batch * p.H;
    }
    wi.row_base = it.qt * (2 * BQ * NCTA) + static_cast<int>(cta_rank) * BQ;
    return wi;
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if constexpr (STAGE) tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_SLOTS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4 * NCTA);
      mbar_init(&o_done[t], 1);
    }
    mbar_init(s_free, 4 * NCTA);
    mbar_init(q_empty, 1);
    mbar_init(&o_free[0], 4 * NCTA);
    mbar_init(&o_free[1], 4 * NCTA);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    if constexpr (NCTA == 2) {
      tmem_alloc_2sm(tmem_ptr, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  // Every role reads the TMEM base address from shared memory INSIDE its branch (read_tmem_base): read up here it was
  // spilled across the role split and re-loaded from local memory inside the MMA issuer's loop; as a compile-time constant
  // (the CTA allocates all 512 columns, the base can only be 0) ptxas materialised one uniform register per MMA.
  auto read_tmem_base = [&]() {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(tmem_ptr)) : "memory");
    return v;
  };

  // Register budget per SMSP slot is 512 / 3 warps: give the two softmax warpgroups 208 registers each (a whole
  // 128-column S row lives in registers) and shrink the TMA/MMA warpgroup to 88.
  // (one setmaxnreg per warpgroup, executed by all four of its warps at the same instruction, at the top of the
  // warpgroup's branch so that ptxas budgets the branch accordingly.)
  if ((warp >> 2) == ROLE_WG) {
   setmaxnreg_dec<REGS_OTHER>();
   if (warp == TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer (every CTA stages its own part)
    if (elect_one()) {
      if constexpr (NCTA == 2) {
        // Both CTAs' bytes complete on the LEADER's barriers; only the leader arrives (expect_tx of both CTAs' bytes).
        // The peer may only refill a slot after the leader's MMAs of the previous round were committed (its own
        // kv_empty, multicast), i.e. after the leader's barrier finished that round, so its bytes can land before the
        // leader's expect_tx of the same round (the tx-count goes negative transiently, which is legal).
        // (pairs are launched non-persistently: one work item)
        const Item wi = decode_item(iter_first());
        const int head = wi.head, batch = wi.batch;
        const uint32_t qbar = mapa_u32(q_full, 0);
        if (leader) mbar_arrive_expect_tx(q_full, 2 * 2 * TILE_BYTES);
        for (int t = 0; t < 2; ++t) {
          const int row0 = wi.row_base + t * TILE_ROW_STEP;
          tma_load_4d_2sm(q_smem + t * TILE_BYTES, &tmQ, qbar, 0, row0, head, batch);
          tma_load_4d_2sm(q_smem + t * TILE_BYTES + HALF_BYTES, &tmQ, qbar, 64, row0, head, batch);
        }
        int slot = 0;
        uint32_t phase = 0;
        const int r64 = static_cast<int>(cta_rank) * 64;
        auto load_k = [&](int j) {   // K_j: this CTA's 64 keys x 128 channels = two [64 x 64] swizzled sub-tiles
          role_wait(&kv_empty[slot], phase ^ 1);
          uint8_t* dst = kv_smem + slot * KV_BYTES;
          const uint32_t fbar = mapa_u32(&kv_full[slot], 0);
          if (leader) mbar_arrive_expect_tx(&kv_full[slot], 2 * KV_BYTES);
          tma_load_4d_2sm(dst, &tmK, fbar, 0, j * BKV + r64, head, batch);
          tma_load_4d_2sm(dst + KV_BYTES / 2, &tmK, fbar, 64, j * BKV + r64, head, batch);
          if (++slot == KV_SLOTS) { slot = 0; phase ^= 1; }
        };
        auto load_v = [&](int j) {   // V_j: all 128 keys x this CTA's 64 channels = one [128 x 64] swizzled half tile
          role_wait(&kv_empty[slot], phase ^ 1);
          uint8_t* dst = kv_smem + slot * KV_BYTES;
          const uint32_t fbar = mapa_u32(&kv_full[slot], 0);
          if (leader) mbar_arrive_expect_tx(&kv_full[slot], 2 * KV_BYTES);
          tma_load_4d_2sm(dst, &tmV, fbar, r64, j * BKV, head, batch);
          if (++slot == KV_SLOTS) { slot = 0; phase ^= 1; }
        };
        load_k(0);
        if (n_kv > 1) load_k(1);
        for (int j = 0; j < n_kv; ++j) {
          load_v(j);
          if (j + 2 < n_kv) load_k(j + 2);
        }
      } else {
        // The K/V ring runs on across work items: the next item's Q (once the last S MMA of this item has read the Q
        // tiles: "Q empty") and its first K/V tiles are in flight while this item's last key tiles and epilogue run.
        int slot = 0;
        uint32_t phase = 0;
        int n_it = 0;
        for (ItemIter it = iter_first(); it.hb < p.n_hb; iter_next(it), ++n_it) {
          if (!MULTI && n_it > 0) break;
          const Item wi = decode_item(it);
          const int head = wi.head, batch = wi.batch;
          if (n_it > 0) role_wait(q_empty, (n_it - 1) & 1);
          mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
          for (int t = 0; t < 2; ++t) {
            const int row0 = wi.row_base + t * TILE_ROW_STEP;
            tma_load_4d(q_smem + t * TILE_BYTES, &tmQ, q_full, 0, row0, head, batch);
            tma_load_4d(q_smem + t * TILE_BYTES + HALF_BYTES, &tmQ, q_full, 64, row0, head, batch);
          }
          auto load_tile = [&](const CUtensorMap* tm, int j) {
            role_wait(&kv_empty[slot], phase ^ 1);
            uint8_t* dst = kv_smem + slot * KV_BYTES;
            mbar_arrive_expect_tx(&kv_full[slot], TILE_BYTES);
            tma_load_4d(dst, tm, &kv_full[slot], 0, j * BKV, head, batch);
            tma_load_4d(dst + HALF_BYTES, tm, &kv_full[slot], 64, j * BKV, head, batch);
            if (++slot == KV_SLOTS) { slot = 0; phase ^= 1; }
          };
          if constexpr (PIPE == 1) {
            load_tile(&tmK, 0);
            if (n_kv > 1) load_tile(&tmK, 1);
            for (int j = 0; j < n_kv; ++j) {
              load_tile(&tmV, j);
              if (j + 2 < n_kv) load_tile(&tmK, j + 2);
            }
          } else {
            for (int j = 0; j < n_kv; ++j) {
              load_tile(&tmK, j);
              load_tile(&tmV, j);
            }
          }
        }
      }
    }
   } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    // ATTN_UNIFORM_ISSUE=1 (experiment, off): ALL 32 lanes of the MMA warp run the issue loop (waits, ring cursor, descriptor
    // arithmetic) and one elected lane issues the tcgen05 instructions.  Warp-uniform control flow lets ptxas keep more of the
    // descriptor arithmetic on the uniform datapath (R2UR 93 -> 51 per key tile) but the loop grows (329 -> 357 instructions)
    // and 32 lanes poll every barrier: measured SLOWER, 1303 vs 1375 TFLOP/s on 40 x 75,600^2 and 1010 vs 1080 on the Wan
    // cross-attention (profiles/r02_attn_uniform_issue_ab.log).  Default: only the elected lane runs the loop.
#ifndef ATTN_UNIFORM_ISSUE
#define ATTN_UNIFORM_ISSUE 0
#endif
    const bool issuer = elect_one();
    if (leader && (ATTN_UNIFORM_ISSUE || issuer)) {
      const uint32_t tmem_base = read_tmem_base();
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(BQ * NCTA, BKV, 0);  // B = K tile, K-major
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(BQ * NCTA, D, 1);    // B = V tile, MN-major
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t kv_addr = smem_u32(kv_smem);
      auto commit = [&](uint64_t* bar) {
        if (!issuer) return;
        if constexpr (NCTA == 2) umma_commit_2sm(bar); else umma_commit(bar);
      };
      auto issue_s = [&](int t, int slot) {
        const uint32_t a0 = q_addr + t * TILE_BYTES;
        const uint32_t b0 = kv_addr + slot * KV_BYTES;
        const uint32_t d_tmem = tmem_base + (t ? S_COL1 : S_COL0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          // channel slice kk: sub-tile kk / 4 (64 channels each), 32 bytes per 16 channels inside the 128-byte row
          const uint64_t da = make_smem_desc_sw128(a0 + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b0 + (kk >> 2) * (KV_BYTES / 2) + (kk & 3) * 32, 16, 1024);
          if (issuer) {
            if constexpr (NCTA == 2) umma_ss_2sm(d_tmem, da, db, idesc_s, kk != 0 ? 1u : 0u);
            else umma_ss(d_tmem, da, db, idesc_s, kk != 0 ? 1u : 0u);
          }
        }
      };
      auto issue_pv = [&](int t, int slot, bool first, uint32_t parity) {
        role_wait(&p_full[t], parity);
        tc_fence_after();
        const uint32_t b0 = kv_addr + slot * KV_BYTES;
        const uint32_t d_tmem = tmem_base + (t ? O_COL1 : O_COL0);
        const uint32_t a_tmem = tmem_base + (t ? P_COL1 : P_COL0);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // 16 keys = 16 rows of 128 bytes; pair: this CTA's 64 channels are ONE swizzle atom (no leading-dim stride)
          const uint64_t db = make_smem_desc_sw128(b0 + kk * 2048, HALF_BYTES, 1024);
          const uint32_t acc = (first && kk == 0) ? 0u : 1u;
          if (issuer) {
            if constexpr (NCTA == 2) umma_ts_2sm(d_tmem, a_tmem + kk * 8, db, idesc_o, acc);
            else umma_ts(d_tmem, a_tmem + kk * 8, db, idesc_o, acc);
          }
        }
      };
      // Ring items in load order -> item i lives in slot i % KV_SLOTS, phase (i / KV_SLOTS) & 1.
      //   PIPE 0:  K_0 V_0 K_1 V_1 ...
      //   PIPE 1:  K_0 K_1 V_0 K_2 V_1 K_3 ... V_{n-3} K_{n-1} V_{n-2} V_{n-1}: K runs TWO tiles ahead of V because S_t(j+2) is
      //            issued while V_j is in use and a 32 KB TMA load takes ~1000 cycles to land; item i + 5 reuses the slot of
      //            item i, so K_{j+2} waits for V_{j-2}'s release and V_j for K_{j-1}'s: >= 1.5 key tiles of lead each.
      uint32_t n_steps = 0;   // key-tile steps of the work items already done by this CTA: parity base of p_full / o_done
      // PIPE 0 (one work item): ring item i lives in slot i % KV_SLOTS, phase (i / KV_SLOTS) & 1
      auto wait_item = [&](int item) {
        role_wait(&kv_full[item % KV_SLOTS], (item / KV_SLOTS) & 1);
        tc_fence_after();
      };
      auto release_item = [&](int item) { commit(&kv_empty[item % KV_SLOTS]); };
      // PIPE 1: the MMA thread first touches the ring items in their LOAD order (K_0 K_1 V_0 K_2 V_1 K_3 ...), so one running
      // (slot, phase) cursor replaces the item arithmetic (the % and / by 5 on a running item index cost ~80 instructions per
      // key tile on the issuing thread -- the thread every hand-over waits for).
      int ring_slot = 0;
      uint32_t ring_phase = 0;
      auto take = [&]() {   // wait for the next ring item to land; returns its slot
        role_wait(&kv_full[ring_slot], ring_phase);
        tc_fence_after();
        const int sl = ring_slot;
        if (++ring_slot == KV_SLOTS) { ring_slot = 0; ring_phase ^= 1; }
        return sl;
      };
      auto release = [&](int sl) { commit(&kv_empty[sl]); };
      if constexpr (PIPE == 1) {
        // Issue order  S1(j+1) PV0(j) S0(j+2) PV1(j): an S only needs the shared S buffer back (the other tile's softmax has
        // pulled its row into registers), a PV only needs its P.  S_t(j+1) is therefore in TMEM long before softmax_t(j) ends.
        // The c-th hand-back of the S buffer (S0(0), S1(0), S0(1), S1(1), ...) completes phase c of s_free; every S issue
        // except the CTA's very first waits for the previous hand-back, also across work items.
        // (Two issuing threads -- one per stream -- were tried: the pipe interleaves their MMAs, an S group then takes ~1000
        // instead of ~600 cycles from issue to "S ready", and S is the stream on the critical path: 1195 vs 1330 TFLOP/s,
        // profiles/r02_attn_dual_issuer.log.)
        // Persistent launch: work item i + 1 starts with S0(0) as soon as its Q and K_0 have landed -- under the epilogue of
        // item i; only its first PV_t waits for that epilogue ("O free": the softmax warps have read O_t out of TMEM).
        uint32_t n_free = 0;
        auto wait_s_free = [&]() {
          role_wait(s_free, n_free & 1);
          ++n_free;
          tc_fence_after();
        };
        const int n_my = !MULTI ? 1 : (w0 < p.n_items ? (p.n_items - w0 + w_step - 1) / w_step : 0);   // work items of this CTA
#pragma unroll 1
        for (int n_it = 0; n_it < n_my; ++n_it) {
          const bool prof_cta = PROF && blockIdx.x == 0 && issuer;
          (void)prof_cta;
          role_wait(q_full, n_it & 1);
          ATTN_STAMP(0, 20);   // Q landed
          const int k0 = take();                             // K_0
          ATTN_STAMP(0, 21);   // K_0 landed
          if (n_it > 0) wait_s_free();                       // softmax 1 pulled the previous item's last S1
          issue_s(0, k0);
          commit(&s_full[0]);
          ATTN_STAMP(0, 22);   // S0(0) issued
          wait_s_free();
          issue_s(1, k0);
          commit(&s_full[1]);
          if (n_kv == 1) commit(q_empty);
          release(k0);
          int k_pend = 0;                                    // slot of K_{j+1}: S0(j+1) issued, S1(j+1) still to come
          if (n_kv > 1) {
            k_pend = take();                                 // K_1
            wait_s_free();
            issue_s(0, k_pend);
            commit(&s_full[0]);
          }
          for (int j = 0; j < n_kv; ++j) {
            const uint32_t parity = ((MULTI ? n_steps : 0u) + j) & 1;
            if (j + 1 < n_kv) {
              wait_s_free();                                   // softmax 0 pulled S0(j+1)
              issue_s(1, k_pend);                              // K_{j+1} landed before S0(j+1) was issued
              commit(&s_full[1]);
              if (j + 2 == n_kv) commit(q_empty);              // that was the item's last S MMA: the Q tiles may be refilled
              release(k_pend);
            }
            ATTN_STAMP(j, 10);
            const int v_slot = take();                         // V_j
            if (j == 0 && n_it > 0) {
              role_wait(&o_free[0], (n_it - 1) & 1);
              tc_fence_after();
            }
            issue_pv(0, v_slot, j == 0, parity);
            ATTN_STAMP(j, 11);   // P0 seen ready, PV0 issued
            commit(&o_done[0]);
            if (j + 2 < n_kv) {
              k_pend = take();                                 // K_{j+2}
              ATTN_STAMP(j, 14);
              wait_s_free();                                   // softmax 1 pulled S1(j+1)
              ATTN_STAMP(j, 15);
              issue_s(0, k_pend);
              commit(&s_full[0]);
            }
            ATTN_STAMP(j, 12);
            if (j == 0 && n_it > 0) {
              role_wait(&o_free[1], (n_it - 1) & 1);
              tc_fence_after();
            }
            issue_pv(1, v_slot, j == 0, parity);
            ATTN_STAMP(j, 13);   // P1 seen ready, PV1 issued
            commit(&o_done[1]);
            release(v_slot);
          }
          if constexpr (MULTI) n_steps += n_kv;
        }
      } else {
        // round-1 order  PV0(j) S0(j+1) PV1(j) S1(j+1)  (P_t aliases S_t); launched non-persistently (one work item)
        const bool prof_cta = PROF && blockIdx.x == 0 && issuer;
        (void)prof_cta;
        role_wait(q_full, 0);
        wait_item(0);
        issue_s(0, 0);
        commit(&s_full[0]);
        issue_s(1, 0);
        commit(&s_full[1]);
        release_item(0);
        for (int j = 0; j < n_kv; ++j) {
          const int v_item = 2 * j + 1, k_item = 2 * j + 2;
          const bool has_next = (j + 1 < n_kv);
          wait_item(v_item);
          ATTN_STAMP(j, 10);
          issue_pv(0, v_item % KV_SLOTS, j == 0, j & 1);
          ATTN_STAMP(j, 11);   // P0 seen ready, PV0 issued
          commit(&o_done[0]);
          if (has_next) {
            wait_item(k_item);
            issue_s(0, k_item % KV_SLOTS);
            commit(&s_full[0]);
          }
          ATTN_STAMP(j, 12);
          issue_pv(1, v_item % KV_SLOTS, j == 0, j & 1);
          ATTN_STAMP(j, 13);   // P1 seen ready, PV1 issued
          commit(&o_done[1]);
          release_item(v_item);
          if (has_next) {
            issue_s(1, k_item % KV_SLOTS);
            commit(&s_full[1]);
            release_item(k_item);
          }
        }
      }
    }
   }  // the remaining warps of this warpgroup idle until the final barrier
  } else {
    // ------------------------------------------------------------------ softmax warps
    setmaxnreg_inc<REGS_SOFTMAX>();
    const uint32_t tmem_base = read_tmem_base();
    const int t = warp >> 2;     // query tile
    const int quad = warp & 3;   // TMEM lane quadrant
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + (t ? S_COL1 : S_COL0);
    const uint32_t p_addr = tmem_base + lane_base + (t ? P_COL1 : P_COL0);
    const uint32_t o_addr = tmem_base + lane_base + (t ? O_COL1 : O_COL0);
    const uint32_t p_full_remote = NCTA == 2 ? mapa_u32(&p_full[t], 0) : 0u;
    const uint32_t s_free_remote = NCTA == 2 ? mapa_u32(s_free, 0) : 0u;
    const uint32_t o_free_remote = NCTA == 2 ? mapa_u32(&o_free[t], 0) : 0u;
    uint32_t n_steps = 0;  // key-tile steps of the work items already done by this CTA: parity base of s_full / o_done
    bool first_item = true;
    for (ItemIter it = iter_first(); it.hb < p.n_hb; iter_next(it)) {
    if (!MULTI && !first_item) break;   // one work item per CTA
    first_item = false;
    const uint32_t step_base = MULTI ? n_steps : 0u;
    const bool prof_cta = PROF && blockIdx.x == 0;
    (void)prof_cta;
    float sl2 = p.scale_log2;
    if (p.q_rowsumsq != nullptr) {
      // logits of this thread's query row carry the row's RMS-norm factor (every head of a token shares it)
      const Item wq = decode_item(it);
      const int qrow = wq.row_base + t * TILE_ROW_STEP + quad * 32 + lane;
      float ss = 0.f;
      if (qrow < p.Sq) {
        const float* part = p.q_rowsumsq + (static_cast<int64_t>(wq.batch) * p.Sq + qrow) * p.q_parts;
        for (int i = 0; i < p.q_parts; ++i) ss += __ldg(part + i);
      }
      sl2 *= __bfloat162float(__float2bfloat16(rsqrtf(ss * p.q_inv_dim + p.q_eps)));   // y.to(dtype=x.dtype), efficiency/mod.py:24-35
    }
    float m = -INFINITY;  // running max of s * scale_log2 actually used for P
    float l = 0.f;        // running sum of P
    // One key tile.  MASKED is a compile-time flag: only the LAST tile can be partial (Sk % 128 != 0); written as a run-time
    // `if` the compiler turned the masking into 128 ISETP + 128 SEL executed for EVERY tile -- 30 % of the loop's instructions.
    auto step = [&](const int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const uint32_t parity = (step_base + j) & 1;
      mbar_wait(&s_full[t], parity);
      tc_fence_after();
      if (quad == 0 && lane == 0) ATTN_STAMP(j, t * 5 + 0);
      const int valid = p.Sk - j * BKV;  // >= 128 except possibly for the last tile
      uint32_t s[128];
      tmem_ld_x32(s_addr + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      tmem_ld_x32(s_addr + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
      tmem_ld_wait();
      if constexpr (PIPE == 1) {
        // the whole row is in registers: hand the shared S buffer back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2 && !leader) mbar_arrive_release_cluster(s_free_remote);
          else mbar_arrive(s_free);
        }
      }
      if (quad == 0 && lane == 0) ATTN_STAMP(j, t * 5 + 1);
      if constexpr (MASKED) {
#pragma unroll
        for (int k = 0; k < 128; ++k)
          if (k >= valid) s[k] = 0xff800000u;  // -inf
      }
      // FMNMX3: one instruction folds two new values into an accumulator -> 62 instead of 124 max ops per row
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int k = 4; k < 124; k += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[k + 4]), __uint_as_float(s[k + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[k + 6]), __uint_as_float(s[k + 7]));
      }
      mx0 = fmax3(mx0, __uint_as_float(s[124]), __uint_as_float(s[125]));
      mx1 = fmax3(mx1, __uint_as_float(s[126]), __uint_as_float(s[127]));
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m, mx * sl2);
      bool o_waited = false;   // PIPE 1: has this warp already seen PV_t(j-1) complete?
      if (j == 0) {
        m = m_new;
      } else if (__any_sync(0xffffffffu, (m_new - m) > RESCALE_THRESHOLD)) {
        if constexpr (PIPE == 1) {
          // O_t is still being accumulated by PV_t(j-1) unless its "O done" phase has completed
          mbar_wait(&o_done[t], parity ^ 1);
          tc_fence_after();
          o_waited = true;
        }
        const float alpha = fast_exp2(m - m_new);
        l *= alpha;
        m = m_new;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_x32(o_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
          tmem_st_x32(o_addr + c * 32, r);
        }
        tmem_st_wait();
      }
      if (quad == 0 && lane == 0) ATTN_STAMP(j, t * 5 + 2);
      const float2 sl2_2 = make_float2(sl2, sl2);
      const float2 negm_2 = make_float2(-m, -m);
      float2 sum2 = make_float2(0.f, 0.f);
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int col = c * 32 + 2 * k;
          const float2 x = ffma2(make_float2(__uint_as_float(s[col]), __uint_as_float(s[col + 1])), sl2_2, negm_2);
          float2 e;
          if (k < POLY_PAIRS) {
            e = exp2_poly2(x);
          } else {
            e.x = fast_exp2(x.x);
            e.y = fast_exp2(x.y);
          }
          sum2 = fadd2(sum2, e);
          pk[c * 16 + k] = pack_bf16x2(e.x, e.y);
        }
        if constexpr (PIPE == 0) {
          // chunk c of S (fp32 columns [32c, 32c+32)) becomes P columns [16c, 16c+16): always inside the part of S this
          // thread has already consumed (the whole row is in registers).
          tmem_st_x16(p_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[c * 16]));
        }
      }
      l += sum2.x + sum2.y;
      if constexpr (PIPE == 1) {
        // P_t(j) goes where PV_t(j-1) reads P_t(j-1): that MMA must have completed (it long has, in steady state)
        if (j > 0 && !o_waited) {
          mbar_wait(&o_done[t], parity ^ 1);
          tc_fence_after();
        }
        tmem_st_x32(p_addr, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tmem_st_x32(p_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      }
      if (quad == 0 && lane == 0) ATTN_STAMP(j, t * 5 + 3);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2 && !leader) mbar_arrive_release_cluster(p_full_remote);
        else mbar_arrive(&p_full[t]);
      }
      if (quad == 0 && lane == 0) ATTN_STAMP(j, t * 5 + 4);
    };
    const int n_full = p.Sk / BKV;   // key tiles without a tail
    for (int j = 0; j < n_full; ++j) step(j, std::false_type{});
    if (n_full < n_kv) step(n_kv - 1, std::true_type{});
    // epilogue: O / l -> bf16 -> global
    mbar_wait(&o_done[t], (step_base + n_kv - 1) & 1);
    tc_fence_after();
    if (quad == 0 && lane == 0) ATTN_STAMP(n_kv - 1, 16 + 2 * t);   // last PV_t seen complete: epilogue begins
    const float inv_l = 1.0f / l;
    const Item wi = decode_item(it);
    const int row = wi.row_base + t * TILE_ROW_STEP + quad * 32 + lane;
    __nv_bfloat16* orow = p.o + wi.batch * p.o_sb + wi.head * p.o_sh + static_cast<int64_t>(row) * p.o_ss;
    int n_dst = 1;               // peers this row is stored to (replicated text rows of a joint sequence: all of them)
    int64_t peer_row_off = 0;    // element offset of the row inside a peer's buffer
    if (p.n_peers > 0 && row < p.Sq) {
      const int n_shard = p.Sq - p.rep_rows;
      const int rs = p.rep_first ? row - p.rep_rows : row;   // index among the token-sharded rows (if it is one)
      if (rs >= 0 && rs < n_shard) {
        const int d = rs / p.rows_per_rank;
        const int local = rs - d * p.rows_per_rank + (p.rep_first ? p.rep_rows : 0);
        orow = reinterpret_cast<__nv_bfloat16*>(p.o_peer[d]) + static_cast<int64_t>(local) * p.o_ss +
               static_cast<int64_t>(wi.head + p.head_off) * p.o_sh;
      } else {
        const int local = p.rep_first ? row : p.rows_per_rank + (row - n_shard);
        peer_row_off = static_cast<int64_t>(local) * p.o_ss + static_cast<int64_t>(wi.head + p.head_off) * p.o_sh;
        orow = reinterpret_cast<__nv_bfloat16*>(p.o_peer[0]) + peer_row_off;
        n_dst = p.n_peers;
      }
    }
    if constexpr (STAGE) {
      // O_t rows [32 * quad, +32) of this warp: registers -> this warp's two 4 KB boxes (channels 0-63 and 64-127; 128-byte
      // rows, 16-byte chunks XOR-swizzled by row & 7 = the tensor map's SWIZZLE_128B, conflict-free) -> ONE proxy fence and
      // two bulk tensor stores of [32 rows x 64 channels]; rows >= Sq are clipped by the tensor map.  The whole row block has
      // its own staging bytes, so nothing inside the epilogue waits for the TMA engine: the stores drain under the next
      // work item, and the wait below (for the PREVIOUS item's stores) has long been satisfied.  (Smaller boxes that had to
      // be re-used inside one epilogue cost a fence + store issue + drain wait per re-use: 2100-2700 cycles per item,
      // profiles/r02_attn_cross_timeline.log.)
      uint8_t* boxes = o_stage + (t * 4 + quad) * 8192;
      const int row0 = wi.row_base + t * TILE_ROW_STEP + quad * 32;
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(o_addr + c * 32, r);
        tmem_ld_wait();
        if (c == 3) {
          // O_t is in registers: the next work item's first PV_t may overwrite the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[t]);
        }
        uint8_t* box = boxes + (c >> 1) * 4096 + lane * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[q4 * 8 + 0]) * inv_l, __uint_as_float(r[q4 * 8 + 1]) * inv_l);
          o.y = pack_bf16x2(__uint_as_float(r[q4 * 8 + 2]) * inv_l, __uint_as_float(r[q4 * 8 + 3]) * inv_l);
          o.z = pack_bf16x2(__uint_as_float(r[q4 * 8 + 4]) * inv_l, __uint_as_float(r[q4 * 8 + 5]) * inv_l);
          o.w = pack_bf16x2(__uint_as_float(r[q4 * 8 + 6]) * inv_l, __uint_as_float(r[q4 * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(box + (((q4 + 4 * (c & 1)) ^ (lane & 7)) << 4)) = o;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && row0 < p.Sq) {
        tma_store_4d(&tmO, boxes, 0, row0, wi.head, wi.batch);
        tma_store_4d(&tmO, boxes + 4096, 64, row0, wi.head, wi.batch);
        tma_store_commit();
      }
    } else {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_x32(o_addr + c * 32, r);
      tmem_ld_wait();
      if (c == 3) {
        // O_t is in registers: the next work item's first PV_t may overwrite the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2 && !leader) mbar_arrive_release_cluster(o_free_remote);
          else mbar_arrive(&o_free[t]);
        }
      }
      if (row < p.Sq) {
        uint32_t pk8[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          pk8[e] = pack_bf16x2(__uint_as_float(r[2 * e]) * inv_l, __uint_as_float(r[2 * e + 1]) * inv_l);
        for (int dst = 0; dst < n_dst; ++dst) {
          __nv_bfloat16* od = dst == 0 ? orow : reinterpret_cast<__nv_bfloat16*>(p.o_peer[dst]) + peer_row_off;
          if (p.o_vec32) {
            // 32 bytes (a whole sector) per lane and instruction: STG.256 halves the line accesses of the 16-byte form
            st_global_v8(od + c * 32, *reinterpret_cast<uint32_t(*)[8]>(&pk8[0]));
            st_global_v8(od + c * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&pk8[8]));
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              reinterpret_cast<uint4*>(od + c * 32)[q4] = make_uint4(pk8[q4 * 4], pk8[q4 * 4 + 1], pk8[q4 * 4 + 2], pk8[q4 * 4 + 3]);
          }
        }
      }
    }
    }
    if (quad == 0 && lane == 0) ATTN_STAMP(n_kv - 1, 17 + 2 * t);   // epilogue stores issued
    if constexpr (MULTI) n_steps += n_kv;
    }  // work items
    if constexpr (STAGE) {
      if (lane == 0) tma_store_wait<0>();   // bulk stores read shared memory: drain before the CTA exits
    }
  }

  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    const uint32_t tmem_base = read_tmem_base();
    if constexpr (NCTA == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

#undef n_kv
#undef w_step
}  // namespace attn
}  // namespace b200

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                         int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                         int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                         float scale, void* const* o_peers, int n_peers, int rows_per_rank, int head_off, void* stream,
                         long long* prof = nullptr, int prof_steps = 0, int rep_rows = 0, int rep_first = 0,
                         const float* q_rowsumsq = nullptr, int q_parts = 0, float q_inv_dim = 0.f, float q_eps = 0.f);

// Diagnostics: b200_attn_fwd with SM-clock timestamps of the softmax / MMA hand-offs of CTA (0,0,0) written to
// prof[prof_steps][32] (device memory): columns 0-4 softmax tile 0 (S seen ready, S in registers, row max done, P stores
// issued, "P ready" arrived), 5-9 the same for tile 1, 10-13 the MMA thread (V tile landed, PV0 issued, S0(j+1) issued, PV1 issued).
extern "C" int b200_attn_fwd_prof(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                                  int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                                  int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                                  float scale, long long* prof, int prof_steps, void* stream) {
  if (!prof || prof_steps <= 0) return B200_ERR_ARG;
  return attn_fwd_impl(q, k, v, o, B, H, Sq, Sk, D, q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss, o_sb, o_sh, o_ss,
                       scale, nullptr, 0, 0, 0, stream, prof, prof_steps);
}

extern "C" int b200_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                             int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                             int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                             float scale, void* stream) {
  return attn_fwd_impl(q, k, v, o, B, H, Sq, Sk, D, q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss, o_sb, o_sh, o_ss,
                       scale, nullptr, 0, 0, 0, stream);
}

extern "C" int b200_attn_fwd_qnorm(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                                   int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                                   int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                                   float scale, const float* q_rowsumsq, int q_parts, int q_norm_dim, float q_eps, void* stream) {
  if (!q_rowsumsq || q_parts <= 0 || q_norm_dim <= 0) return B200_ERR_ARG;
  return attn_fwd_impl(q, k, v, o, B, H, Sq, Sk, D, q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss, o_sb, o_sh, o_ss,
                       scale, nullptr, 0, 0, 0, stream, nullptr, 0, 0, 0, q_rowsumsq, q_parts, 1.0f / static_cast<float>(q_norm_dim),
                       q_eps);
}

extern "C" int b200_attn_fwd_scatter_joint(const void* q, const void* k, const void* v, int H, int Sq, int Sk, int D,
                                           int64_t q_sh, int64_t q_ss, int64_t k_sh, int64_t k_ss, int64_t v_sh, int64_t v_ss,
                                           void* const* o_peers, int n_peers, int rows_per_rank, int head_off, int64_t o_sh,
                                           int64_t o_ss, int rep_rows, int rep_first, float scale, void* stream) {
  if (!o_peers || n_peers < 1 || n_peers > 8 || rows_per_rank <= 0 || rep_rows < 0 || rep_rows > Sq) return B200_ERR_ARG;
  if (static_cast<int64_t>(rows_per_rank) * n_peers < Sq - rep_rows) return B200_ERR_SHAPE;
  for (int i = 0; i < n_peers; ++i)
    if (!o_peers[i] || (reinterpret_cast<uintptr_t>(o_peers[i]) & 15)) return B200_ERR_ALIGN;
  return attn_fwd_impl(q, k, v, o_peers[0], 1, H, Sq, Sk, D, 0, q_sh, q_ss, 0, k_sh, k_ss, 0, v_sh, v_ss, 0, o_sh, o_ss,
                       scale, o_peers, n_peers, rows_per_rank, head_off, stream, nullptr, 0, rep_rows, rep_first ? 1 : 0);
}

extern "C" int b200_attn_fwd_scatter(const void* q, const void* k, const void* v, int H, int Sq, int Sk, int D,
                                     int64_t q_sh, int64_t q_ss, int64_t k_sh, int64_t k_ss, int64_t v_sh, int64_t v_ss,
                                     void* const* o_peers, int n_peers, int rows_per_rank, int head_off, int64_t o_sh,
                                     int64_t o_ss, float scale, void* stream) {
  return b200_attn_fwd_scatter_joint(q, k, v, H, Sq, Sk, D, q_sh, q_ss, k_sh, k_ss, v_sh, v_ss, o_peers, n_peers, rows_per_rank,
                                     head_off, o_sh, o_ss, 0, 0, scale, stream);
}

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                         int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                         int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                         float scale, void* const* o_peers, int n_peers, int rows_per_rank, int head_off, void* stream,
                         long long* prof, int prof_steps, int rep_rows, int rep_first, const float* q_rowsumsq, int q_parts,
                         float q_inv_dim, float q_eps) {
  using namespace b200;
  using namespace b200::attn;
  if (!q || !k || !v || !o) return B200_ERR_ARG;
  if (D != 128) return B200_ERR_SHAPE;
  if (B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) return B200_ERR_SHAPE;
  if ((o_sb % 8) || (o_sh % 8) || (o_ss % 8) || (reinterpret_cast<uintptr_t>(o) & 15)) return B200_ERR_ALIGN;

  // B200_ATTN_2CTA=1 selects CTA pairs (0 / default: single CTAs); B200_ATTN_PIPE=0 selects the round-1 pipeline (1-CTA
  // only; the A/B partner), default the decoupled pipeline.  The opt-in to > 48 KB of dynamic shared memory is per device.
  // B200_ATTN_PERSIST=0 launches one CTA per work item (the round-1 grid); default: one CTA per SM walks the items.
  // B200_ATTN_STAGE_KV=n: key sequences of up to n tiles use the staged (TMA-store) epilogue with a 3-slot K/V ring (0 = never).
  // B200_ATTN_MULTI_KV=n: key sequences of up to n tiles are launched persistently.
  static int pair_mode = -2, pipe_mode = -2, persist_mode = -2, stage_kv = STAGE_MAX_KV_TILES, multi_kv = MULTI_MAX_KV_TILES;
  if (pair_mode == -2) {
    const char* ev = getenv("B200_ATTN_STAGE_KV");
    if (ev) stage_kv = atoi(ev);
    ev = getenv("B200_ATTN_MULTI_KV");
    if (ev) multi_kv = atoi(ev);
    ev = getenv("B200_ATTN_PERSIST");
    persist_mode = ev ? (ev[0] == '0' ? 0 : 1) : 1;
    ev = getenv("B200_ATTN_PIPE");
    pipe_mode = ev ? (ev[0] == '0' ? 0 : 1) : 1;
    ev = getenv("B200_ATTN_2CTA");
    pair_mode = ev ? (ev[0] == '1' ? 1 : 0) : -1;
  }
  const bool use_pair = pipe_mode == 1 && (pair_mode == 1 || (pair_mode == -1 && PAIR_BY_DEFAULT && Sq > 2 * BQ));
  static std::atomic<bool> attr_done[kMaxDevices];
  if (!once_per_device(attr_done, [] {
        bool ok = true;
#define B200_ATTN_OPT_IN(NCTA, PIPE, STAGE, MULTI)                                                                        \
  ok &= cudaFuncSetAttribute(attn_fwd_kernel<NCTA, PIPE, false, STAGE, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             Cfg<NCTA, STAGE>::SMEM_BYTES) == cudaSuccess;                                                \
  ok &= cudaFuncSetAttribute(attn_fwd_kernel<NCTA, PIPE, true, STAGE, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             Cfg<NCTA, STAGE>::SMEM_BYTES) == cudaSuccess;
        B200_ATTN_OPT_IN(1, 0, false, false)
        B200_ATTN_OPT_IN(2, 1, false, false)
        B200_ATTN_OPT_IN(1, 1, false, false)
        B200_ATTN_OPT_IN(1, 1, false, true)
        B200_ATTN_OPT_IN(1, 1, true, false)
        B200_ATTN_OPT_IN(1, 1, true, true)
#undef B200_ATTN_OPT_IN
        return ok;
      }))
    return B200_ERR_LAUNCH;

  const bool use_stage = pipe_mode == 1 && !use_pair && n_peers == 0 && (Sk + BKV - 1) / BKV <= stage_kv;
  CUtensorMap tmQ, tmK, tmV, tmO;
  const uint32_t box_o[4] = {64, 32, 1, 1};        // staged epilogue: one warp's 32 rows x 64 channels
  const uint32_t box[4] = {64, 128, 1, 1};
  const uint32_t box_k_pair[4] = {64, 64, 1, 1};   // pair: each CTA stages 64 of the 128 keys of a K tile
  {
    uint64_t dims[4] = {128, (uint64_t)Sq, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {1, (uint64_t)q_ss, (uint64_t)q_sh, (uint64_t)q_sb};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {128, (uint64_t)Sk, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {1, (uint64_t)k_ss, (uint64_t)k_sh, (uint64_t)k_sb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, str, use_pair ? box_k_pair : box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {128, (uint64_t)Sk, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {1, (uint64_t)v_ss, (uint64_t)v_sh, (uint64_t)v_sb};
    int rc = make_tmap_bf16(&tmV, v, 4, dims, str, box);
    if (rc) return rc;
  }
  if (use_stage) {
    uint64_t dims[4] = {128, (uint64_t)Sq, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {1, (uint64_t)o_ss, (uint64_t)o_sh, (uint64_t)o_sb};
    int rc = make_tmap_bf16(&tmO, o, 4, dims, str, box_o);
    if (rc) return rc;
  } else {
    tmO = tmQ;   // unused
  }
  Params p;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.o_sb = o_sb; p.o_sh = o_sh; p.o_ss = o_ss;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.n_peers = n_peers;
  p.rows_per_rank = rows_per_rank;
  p.head_off = head_off;
  p.rep_rows = rep_rows;
  p.rep_first = rep_first;
  p.q_rowsumsq = q_rowsumsq;
  p.q_parts = q_parts;
  p.q_inv_dim = q_inv_dim;
  p.q_eps = q_eps;
  for (int i = 0; i < 8; ++i) p.o_peer[i] = (o_peers && i < n_peers) ? o_peers[i] : nullptr;
  p.prof = prof;
  p.prof_steps = prof_steps;
  p.o_vec32 = !((o_sb % 16) || (o_sh % 16) || (o_ss % 16) || (reinterpret_cast<uintptr_t>(o) & 31));
  for (int i = 0; i < n_peers; ++i)
    if (reinterpret_cast<uintptr_t>(o_peers[i]) & 31) p.o_vec32 = 0;
  const int rows_per_item = 2 * BQ * (use_pair ? 2 : 1);
  const int64_t n_qt = (Sq + rows_per_item - 1) / rows_per_item;
  const int64_t n_items = n_qt * H * B;
  if (n_items > (1 << 30)) return B200_ERR_SHAPE;
  p.n_kv = (Sk + BKV - 1) / BKV;
  p.n_qt = static_cast<int>(n_qt);
  p.n_hb = H * B;
  p.n_items = static_cast<int>(n_items);

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (use_pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * p.n_items, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<2>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = prof ? cudaLaunchKernelEx(&cfg, attn_fwd_kernel<2, 1, true>, tmQ, tmK, tmV, tmO, p)
                                : cudaLaunchKernelEx(&cfg, attn_fwd_kernel<2, 1>, tmQ, tmK, tmV, tmO, p);
    if (le != cudaSuccess) {
      cudaGetLastError();
      return B200_ERR_LAUNCH;
    }
  } else {
    // persistent (decoupled pipeline only): one CTA per SM walks the work items -- for key sequences up to multi_kv tiles;
    // beyond that the per-CTA prologue is amortised and the leaner single-item kernel is used.
    const bool persist = pipe_mode == 1 && persist_mode == 1 && p.n_items > num_sms() && p.n_kv <= multi_kv;
    dim3 grid(persist ? num_sms() : p.n_items, 1, 1);
#define B200_ATTN_LAUNCH(PIPE, STAGE, MULTI)                                                                                        \
  do {                                                                                                                             \
    if (prof) attn_fwd_kernel<1, PIPE, true, STAGE, MULTI><<<grid, NUM_THREADS, Cfg<1, STAGE>::SMEM_BYTES, st>>>(tmQ, tmK, tmV, tmO, p); \
    else attn_fwd_kernel<1, PIPE, false, STAGE, MULTI><<<grid, NUM_THREADS, Cfg<1, STAGE>::SMEM_BYTES, st>>>(tmQ, tmK, tmV, tmO, p);     \
  } while (0)
    if (pipe_mode != 1) B200_ATTN_LAUNCH(0, false, false);
    else if (use_stage && persist) B200_ATTN_LAUNCH(1, true, true);
    else if (use_stage) B200_ATTN_LAUNCH(1, true, false);
    else if (persist) B200_ATTN_LAUNCH(1, false, true);
    else B200_ATTN_LAUNCH(1, false, false);
#undef B200_ATTN_LAUNCH
  }
  B200_CHECK_LAUNCH();
  return B200_OK;
}

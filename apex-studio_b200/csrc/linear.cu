// b200_linear: C = epilogue(A @ W^T + bias) on the 5th-gen tensor cores.
//
// Persistent, warp-specialised kernel, one CTA per SM; by default the CTAs work in PAIRS (cluster of 2, cta_group::2):
//   warp 0      TMA producer   (A tile 128x64 + this CTA's 128 of the 256 W rows, 128B swizzle, 6-stage mbarrier ring;
//                               both CTAs' loads complete on the LEADER's barrier)
//   warp 1      MMA issuer     (leader CTA only: tcgen05.mma cta_group::2, M=256 N=256 K=16 -- 128 rows per CTA, the W
//                               halves of both CTAs -- fp32 accumulators in the TMEM of both CTAs, two accumulator
//                               buffers = all 512 TMEM columns, so the epilogue of tile i overlaps the main loop of tile
//                               i+1; tcgen05.commit.multicast releases the smem slots / publishes the accumulator in both)
//   (M <= 128, or B200_LINEAR_2CTA=0: the 1-CTA form -- 128x256 tile, W tile 256x64, 4 stages, cta_group::1)
//   warps 2-5   epilogue       (tcgen05.ld 32x32b, bias / gelu-tanh / gate*acc+residual, 16-byte stores)
// Tiles are walked in groups of GROUP_M row-tiles so that the W panel of a group stays in L2.
//
// Reference call sites this replaces: transformer/wan/base/attention.py:345-347,407 (to_q/to_k/to_v/to_out),
// diffusers FeedForward built at transformer/wan/base/model.py:1062 and called :1270, and the gate/residual
// updates at model.py:1212-1213,1245-1251,1278-1279 (fused into the epilogue of the preceding projection).
#include "host_util.cuh"
#include <stdlib.h>
#include "sm100_ptx.cuh"

namespace b200 {
namespace linear {

#ifndef LINEAR_PARKED
#define LINEAR_PARKED 1
#endif
// every wait of this kernel is long (a k-block for the producer / issuer, a whole tile for the epilogue warps): park in
// hardware instead of re-polling (see mbar_wait_parked)
B200_DEVICE void lin_wait(uint64_t* bar, uint32_t parity) {
  if (LINEAR_PARKED) mbar_wait_parked(bar, parity); else mbar_wait(bar, parity);
}

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int GROUP_M = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 192;
// epilogue staging for the TMA-store path: per epilogue warp two [32 rows x 64 columns] bf16 boxes (128-byte rows, 128B swizzle)
constexpr int EPI_BOX_BYTES = 32 * 128;
constexpr int EPI_STAGING_BYTES = 4 * 2 * EPI_BOX_BYTES;
constexpr bool TALL_BY_DEFAULT = false;   // set once measured on the B200 (profiles/r02_gemm_tall_ab.log)
constexpr bool PAIR_BY_DEFAULT = true;    // validated on the B200: +10 % over the 1-CTA kernel (profiles/r01_gpu_session26_linear_pair.log)

// NCTA = 1: one CTA per 128 x 256 output tile.  NCTA = 2 (cta_group::2): a CTA PAIR (cluster of 2 on one TPC) owns a
// 256 x 256 tile -- each CTA stages its own 128 rows of A and only HALF of the W tile (128 of its 256 rows); the leader's
// tcgen05.mma.cta_group::2 (M = 256) reads both halves, so every W byte is fetched from L2 and written to shared memory
// once per 256 output rows instead of once per 128: 64 instead of 96 bytes per clock per SM, and 6 pipeline stages fit.
// BN_T: N extent of the output tile.  256 everywhere except the SMALL-M form <1, 64>: the 512-row text streams of the
// dual-stream image models give only 24-96 tiles of 256 columns on 148 SMs; 64-column tiles quadruple the tile count
// (profiles/r01_gemm_shapes_vs_cublas_pair.txt: 236-755 TFLOP/s on those shapes with 256-column tiles).
// TALL (NCTA = 2 only): every CTA owns 256 rows (two M = 128 halves, i.e. two M = 256 pair-MMAs per k-step into the two
// halves of TMEM), so a CTA pair owns a 512 x 256 tile and each W byte is fetched once per 512 output rows: 96 KB from L2
// per 512 x 256 x 64 block instead of 128 KB -- the traffic of the QUAD form, but on all 148 SMs.  Both 256-column
// accumulators belong to ONE tile, so the epilogue (1.6k cycles) is no longer hidden behind the next tile's main loop
// (82k cycles at K = 5120): a 2 % cost.  This is the tile shape cuBLAS picks for these GEMMs (nvjet 256x256_64x4 2cta,
// profiles/r02_cublas_qkv_ncu.txt).
template <int NCTA, int BN_T = BN, bool TALL = false>
struct Cfg {
  static constexpr int A_TILE_BYTES = A_BYTES * (TALL ? 2 : 1);
  static constexpr int B_ROWS = BN_T / NCTA;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_BYTES;
  static constexpr int STAGES = TALL ? 4 : (NCTA == 2 ? 6 : (BN_T == 64 ? 8 : 4));   // 64-column tiles: 128 MMA cycles per k-block, deeper ring
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + EPI_STAGING_BYTES;
};

struct Params {
  int M, N, K;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* gate;
  void* C;
  int64_t ldc;
  int epi;
  int bias_row;
  int tiles_m, tiles_n;
  int tma_store; // 1: bf16 output written through shared memory + TMA (128-byte lines) instead of 16-byte stores per lane
  int c_vec32;   // rows of C are 32-byte aligned (bf16 outputs): the direct epilogues use 32-byte loads / stores
  // B200_EPI_NORMW (the RMS-norm of the cross-attention query folded into its projection, attention.py:345-370): the epilogue
  // stores bf16(q * w[n]) with q = bf16(acc + bias) and writes sum_n q^2 of every row and column tile to row_sumsq[row, part];
  // the attention kernel multiplies its logits by rsqrt(mean + eps) of the row (b200_attn_fwd_qnorm).
  float* row_sumsq;
  int n_parts;
  int group_m;   // row-tiles per rasterisation group (the A rows of a group stay in L2 while its n-tiles are swept)
  int panel_n;   // column-tiles per panel: the W panel (panel_n x BN x K) stays in L2 while ALL row groups sweep it
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  -- torch F.gelu(approximate="tanh")
  const float k0 = 0.7978845608028654f;
  const float k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + fast_tanh(inner));
}

// Two-level rasterisation for L2 residency (the output is walked panel by panel, inside a panel group by group, inside a group
// column by column, rows fastest):  W panel (panel_n x BN x K bytes x 2) + A group (group_m x tile rows x K x 2) are sized by the
// host to fit the 126 MB L2 together, so W is read from HBM once per panel... once in total, and A once per panel.  With one
// 16-row-tile group over the whole N (round 1) the 75600 x 15360 x 5120 GEMM read 7.8 GB from HBM for 0.93 GB of operands and
// the N = 5120 shapes ran 25 % faster with group_m = 4 (profiles/r02_gemm_groupm_sweep.log).
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int panel_n, int& tm, int& tn) {
  const int per_panel = tiles_m * panel_n;
  const int panel = tile / per_panel;
  const int first_n = panel * panel_n;
  const int pw = min(tiles_n - first_n, panel_n);          // width of this panel (the last one may be narrower)
  const int rp = tile - panel * per_panel;                 // all full panels precede a narrower last one
  const int per_group = group_m * pw;
  const int group = rp / per_group;
  const int first_m = group * group_m;
  const int gsz = min(tiles_m - first_m, group_m);
  const int r = rp - group * per_group;
  tm = first_m + r % gsz;
  tn = first_n + r / gsz;
}

// QUAD (NCTA = 2 only): clusters of FOUR CTAs = two pairs that own horizontally adjacent 256 x 256 tiles (same rows, columns tn
// and tn + 1).  Both pairs need the same A rows, so every A box is fetched from L2 ONCE and TMA-multicast into both pairs'
// shared memory (each CTA issues one 64-row box for itself and its counterpart in the other pair): L2 -> SM traffic per tile and
// k-block drops from 64 KB to 48 KB.  A slot may only be refilled when BOTH pairs have consumed it: their commits are multicast
// to all four CTAs and the `empty` barriers count two arrivals.
template <int NCTA, int BN_T, bool QUAD = false, bool TALL = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmC, Params p) {
  static_assert(!QUAD || NCTA == 2, "QUAD is a form of the CTA-pair kernel");
  static_assert(!TALL || (NCTA == 2 && !QUAD && BN_T == BN), "TALL is a form of the plain CTA-pair kernel");
  constexpr int STAGES = Cfg<NCTA, BN_T, TALL>::STAGES;
  constexpr int STAGE_BYTES = Cfg<NCTA, BN_T, TALL>::STAGE_BYTES;
  constexpr int A_TILE_BYTES = Cfg<NCTA, BN_T, TALL>::A_TILE_BYTES;
  constexpr int B_ROWS = Cfg<NCTA, BN_T, TALL>::B_ROWS;
  constexpr int ROWS_PER_CTA = TALL ? 2 * BM : BM;
  constexpr int CLUSTER = QUAD ? 4 : NCTA;
  const uint32_t cl_rank = NCTA == 2 ? cluster_ctarank() : 0u;   // rank inside the cluster
  const uint32_t cta_rank = cl_rank & 1u;                         // rank inside the CTA pair
  const uint32_t pair_id = cl_rank >> 1;                          // QUAD: which of the two pairs (column tile offset)
  const bool leader = cta_rank == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES + 1024;   // 1024-aligned: a swizzle atom is 8 rows x 128 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);         // pair: the LEADER's barrier collects the bytes of both CTAs (only the leader arrives)
      mbar_init(&empty[s], QUAD ? 2 : 1);   // QUAD: both pairs must have consumed the slot
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4 * NCTA);   // pair: the epilogue warps of BOTH CTAs release the leader's MMA issuer
    }
    fence_mbar_init();
  }
  if (warp == 1) {
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
  const uint32_t tmem_base = *tmem_ptr;

  const int num_kb = (p.K + BK - 1) / BK;
  // QUAD: p.tiles_n counts column tile PAIRS; this CTA pair owns column tile 2 * tn + pair_id of each
  const int total_tiles = p.tiles_m * p.tiles_n;      // tiles of (BM * NCTA) x BN (QUAD: x 2 BN)
  const int first_tile = blockIdx.x / CLUSTER;          // one tile stream per CTA pair (QUAD: per quad)
  const int tile_step = gridDim.x / CLUSTER;
  const uint16_t commit_mask = QUAD ? static_cast<uint16_t>(3u << (2 * pair_id)) : static_cast<uint16_t>(3);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        int tm, tn;
        tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, p.panel_n, tm, tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          lin_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          if constexpr (NCTA == 2) {
            // both CTAs fill their own shared memory; the bytes of both complete on the LEADER's barrier.  The peer never
            // arrives: it may only refill a slot after the leader's MMAs of the previous round were committed (its own
            // `empty` barrier, multicast), i.e. after the leader's barrier finished that round -- so its bytes can land
            // before the leader's expect_tx of the same round (the tx-count goes negative transiently, which is legal).
            const uint32_t lbar = leader_bar(&full[stage]);
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
            if constexpr (QUAD) {
              // rows [64 * pair_id, +64) of this CTA's 128 A rows, multicast to this CTA and its counterpart in the other pair
              const uint16_t mc = static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2)));
              tma_load_2d_2sm_mc(sa + pair_id * (A_BYTES / 2), &tmA, lbar, kb * BK,
                                 (tm * 2 + static_cast<int>(cta_rank)) * BM + static_cast<int>(pair_id) * (BM / 2), mc);
              tma_load_2d_2sm(sb, &tmB, lbar, kb * BK, (tn * 2 + static_cast<int>(pair_id)) * BN_T + static_cast<int>(cta_rank) * B_ROWS);
            } else {
              // TALL: one box of 256 rows (the tensor map's box is ROWS_PER_CTA tall)
              tma_load_2d_2sm(sa, &tmA, lbar, kb * BK, (tm * 2 + static_cast<int>(cta_rank)) * ROWS_PER_CTA);
              tma_load_2d_2sm(sb, &tmB, lbar, kb * BK, tn * BN_T + static_cast<int>(cta_rank) * B_ROWS);
            }
          } else {
            mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, tm * BM);
            tma_load_2d(sb, &tmB, &full[stage], kb * BK, tn * BN_T);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (pair: the leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM * NCTA, BN_T, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
        // TALL: ONE accumulator set (both 256-column halves belong to this tile), phase = tile parity
        const int acc = TALL ? 0 : (it & 1);
        const uint32_t acc_phase = TALL ? (it & 1) : ((it >> 1) & 1);
        lin_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          lin_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            if constexpr (NCTA == 2) umma_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (TALL) {   // rows 128..255 of both CTAs -> TMEM columns 256..511
              const uint64_t da2 = make_smem_desc_sw128(a_addr + A_BYTES + k * 32, 16, 1024);
              umma_ss_2sm(d_tmem + BN, da2, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (QUAD) umma_commit_2sm_mask(&empty[stage], 0xF);       // all four CTAs: a slot is shared by both pairs
          else if constexpr (NCTA == 2) umma_commit_2sm(&empty[stage]);
          else umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (NCTA == 2) umma_commit_2sm_mask(&acc_full[acc], commit_mask); else umma_commit(&acc_full[acc]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps (2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++it) {
      int tm, tn;
      tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, p.panel_n, tm, tn);
      const int acc = TALL ? 0 : (it & 1);
      const uint32_t acc_phase = TALL ? (it & 1) : ((it >> 1) & 1);
      lin_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < (TALL ? 2 : 1); ++half) {
      const int row = (tm * NCTA + static_cast<int>(cta_rank)) * ROWS_PER_CTA + half * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (acc + half) * BN;
      if (p.tma_store) {
        // Output through shared memory + TMA: every lane owns ONE row of the accumulator, so direct stores are 32 scattered
        // 16-byte pieces per instruction (262,144 L2 write transactions per 256 x 256 tile -- more than the whole operand
        // stream); here a warp stages [32 rows x 64 columns] (128-byte rows, 128B swizzle = conflict-free 16-byte chunks)
        // and one elected lane issues a bulk tensor store of full 128-byte lines; the M / N tails are clipped by the TMA unit.
        uint8_t* my_stage = epi_stage + (warp - 2) * 2 * EPI_BOX_BYTES;
        const int row_base = row - lane;
        float ss = 0.f;   // B200_EPI_NORMW: sum of squares of this row over the tile's columns
#pragma unroll 1
        for (int c = 0; c < BN_T / 64; ++c) {
          const int col0 = (QUAD ? tn * 2 + static_cast<int>(pair_id) : tn) * BN_T + c * 64;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r0[32], r1[32];
          tmem_ld_x32(t_addr + c * 64, r0);
          tmem_ld_x32(t_addr + c * 64 + 32, r1);
          tmem_ld_wait();
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[32 + j] = __uint_as_float(r1[j]);
          }
          if (p.bias != nullptr && p.bias_row) {
            const float br = row_ok ? __bfloat162float(p.bias[row]) : 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] += br;
          } else if (p.bias != nullptr) {
            if (col0 + 64 <= p.N) {
              const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const uint4 b = __ldg(bp + q4);
                v[q4 * 8 + 0] += bf16_lo(b.x); v[q4 * 8 + 1] += bf16_hi(b.x);
                v[q4 * 8 + 2] += bf16_lo(b.y); v[q4 * 8 + 3] += bf16_hi(b.y);
                v[q4 * 8 + 4] += bf16_lo(b.z); v[q4 * 8 + 5] += bf16_hi(b.z);
                v[q4 * 8 + 6] += bf16_lo(b.w); v[q4 * 8 + 7] += bf16_hi(b.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j)
                if (col0 + j < p.N) v[j] += __bfloat162float(p.bias[col0 + j]);
            }
          }
          if (p.epi == B200_EPI_NORMW) {
            const bool full64 = (col0 + 64 <= p.N);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              float wv[8];
              if (full64) {
                const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.gate + col0) + q4);
                wv[0] = bf16_lo(g.x); wv[1] = bf16_hi(g.x); wv[2] = bf16_lo(g.y); wv[3] = bf16_hi(g.y);
                wv[4] = bf16_lo(g.z); wv[5] = bf16_hi(g.z); wv[6] = bf16_lo(g.w); wv[7] = bf16_hi(g.w);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[j] = (col0 + q4 * 8 + j < p.N) ? __bfloat162float(p.gate[col0 + q4 * 8 + j]) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float q = __bfloat162float(__float2bfloat16(v[q4 * 8 + j]));   // the projection output as the reference stores it
                if (full64 || col0 + q4 * 8 + j < p.N) ss += q * q;
                v[q4 * 8 + j] = q * wv[j];
              }
            }
          } else if (p.epi == B200_EPI_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = gelu_tanh(v[j]);
          } else if (p.epi == B200_EPI_SILU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
          } else if (p.epi == B200_EPI_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
          }
          uint8_t* box = my_stage + (c & 1) * EPI_BOX_BYTES;
          if (lane == 0) tma_store_wait_read<1>();     // the store that last read this box (two chunks ago) has drained it
          __syncwarp();
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            uint4 o;
            o.x = pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]);
            o.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
            o.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]);
            o.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
            *reinterpret_cast<uint4*>(box + lane * 128 + ((q4 ^ (lane & 7)) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, box, col0, row_base);
            tma_store_commit();
          }
        }
        if (p.epi == B200_EPI_NORMW && row_ok) {
          const int part = QUAD ? tn * 2 + static_cast<int>(pair_id) : tn;
          if (part < p.n_parts) p.row_sumsq[static_cast<int64_t>(row) * p.n_parts + part] = ss;   // (QUAD: an odd tile count leaves a phantom tile)
        }
        continue;   // next half / tile
      }
#pragma unroll 1
      for (int c = 0; c < BN_T / 32; ++c) {
        const int col0 = (QUAD ? tn * 2 + static_cast<int>(pair_id) : tn) * BN_T + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_x32(t_addr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool full_chunk = (col0 + 32 <= p.N);
        if (p.bias != nullptr && p.bias_row) {
          const float br = row_ok ? __bfloat162float(p.bias[row]) : 0.f;  // per-ROW bias (transposed projections)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += br;
        } else if (p.bias != nullptr) {
          if (full_chunk) {
            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint4 b = __ldg(bp + q4);
              v[q4 * 8 + 0] += bf16_lo(b.x); v[q4 * 8 + 1] += bf16_hi(b.x);
              v[q4 * 8 + 2] += bf16_lo(b.y); v[q4 * 8 + 3] += bf16_hi(b.y);
              v[q4 * 8 + 4] += bf16_lo(b.z); v[q4 * 8 + 5] += bf16_hi(b.z);
              v[q4 * 8 + 6] += bf16_lo(b.w); v[q4 * 8 + 7] += bf16_hi(b.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __bfloat162float(p.bias[col0 + j]);
          }
        }
        if (p.epi == B200_EPI_GELU_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
        } else if (p.epi == B200_EPI_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
        } else if (p.epi == B200_EPI_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
        }
        if (p.epi == B200_EPI_BIAS_F32) {
          if (row_ok) {
            float* cp = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
            if (full_chunk) {
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4)
                reinterpret_cast<float4*>(cp)[q4] = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) cp[j] = v[j];
            }
          }
        } else {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
          if (p.epi == B200_EPI_GATE_RES) {
            // residual stream update: h = h + gate * (acc + bias), fp32, one rounding.
            if (row_ok) {
              if (full_chunk) {
                // The 64 bytes of h this thread updates are loaded up front -- as two 32-byte (full-sector) loads when the row
                // is 32-byte aligned -- and written back the same way.  (One thread owns one output row: every warp-wide
                // access touches 32 different lines, and interleaved 16-byte load / store pairs on the same pointer kept four
                // dependent L2 round trips per chunk in flight one at a time.)
                uint4 hv[4];
                if (p.c_vec32) {
                  ld_global_v8(cp, *reinterpret_cast<uint4(*)[2]>(&hv[0]));
                  ld_global_v8(cp + 16, *reinterpret_cast<uint4(*)[2]>(&hv[2]));
                } else {
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) hv[q4] = reinterpret_cast<const uint4*>(cp)[q4];
                }
                uint4 ov[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const uint4 h = hv[q4];
                  float g[8];
                  if (p.gate != nullptr) {
                    uint4 gg = __ldg(reinterpret_cast<const uint4*>(p.gate + col0) + q4);
                    g[0] = bf16_lo(gg.x); g[1] = bf16_hi(gg.x); g[2] = bf16_lo(gg.y); g[3] = bf16_hi(gg.y);
                    g[4] = bf16_lo(gg.z); g[5] = bf16_hi(gg.z); g[6] = bf16_lo(gg.w); g[7] = bf16_hi(gg.w);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = 1.0f;
                  }
                  float* vv = v + q4 * 8;
                  uint4 o;
                  o.x = pack_bf16x2(bf16_lo(h.x) + g[0] * vv[0], bf16_hi(h.x) + g[1] * vv[1]);
                  o.y = pack_bf16x2(bf16_lo(h.y) + g[2] * vv[2], bf16_hi(h.y) + g[3] * vv[3]);
                  o.z = pack_bf16x2(bf16_lo(h.z) + g[4] * vv[4], bf16_hi(h.z) + g[5] * vv[5]);
                  o.w = pack_bf16x2(bf16_lo(h.w) + g[6] * vv[6], bf16_hi(h.w) + g[7] * vv[7]);
                  ov[q4] = o;
                }
                if (p.c_vec32) {
                  st_global_v8(cp, *reinterpret_cast<uint32_t(*)[8]>(&ov[0]));
                  st_global_v8(cp + 16, *reinterpret_cast<uint32_t(*)[8]>(&ov[2]));
                } else {
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) reinterpret_cast<uint4*>(cp)[q4] = ov[q4];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) {
                    float g = p.gate ? __bfloat162float(p.gate[col0 + j]) : 1.0f;
                    cp[j] = __float2bfloat16(__bfloat162float(cp[j]) + g * v[j]);
                  }
              }
            }
          } else if (row_ok) {
            if (full_chunk) {
              uint4 ov[4];
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float* vv = v + q4 * 8;
                ov[q4].x = pack_bf16x2(vv[0], vv[1]);
                ov[q4].y = pack_bf16x2(vv[2], vv[3]);
                ov[q4].z = pack_bf16x2(vv[4], vv[5]);
                ov[q4].w = pack_bf16x2(vv[6], vv[7]);
              }
              if (p.c_vec32) {
                st_global_v8(cp, *reinterpret_cast<uint32_t(*)[8]>(&ov[0]));
                st_global_v8(cp + 16, *reinterpret_cast<uint32_t(*)[8]>(&ov[2]));
              } else {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) reinterpret_cast<uint4*>(cp)[q4] = ov[q4];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) cp[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      }   // half
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (NCTA == 2) mbar_arrive_cluster(leader_bar(&acc_empty[acc]));
        else mbar_arrive(&acc_empty[acc]);
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait<0>();   // bulk stores read shared memory: drain before the CTA exits
  }

  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (NCTA == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace linear
}  // namespace b200

static int linear_impl(const void* A, const void* W, const void* bias, void* C, const void* gate, int M, int N, int K,
                       int64_t lda, int64_t ldw, int64_t ldc, int epilogue, void* stream, float* row_sumsq, int* n_parts_out);

extern "C" int b200_linear(const void* A, const void* W, const void* bias, void* C, const void* gate, int M, int N,
                           int K, int64_t lda, int64_t ldw, int64_t ldc, int epilogue, void* stream) {
  if ((epilogue & ~B200_EPI_ROW_BIAS) == B200_EPI_NORMW) return B200_ERR_ARG;   // needs b200_linear_normw (row_sumsq output)
  return linear_impl(A, W, bias, C, gate, M, N, K, lda, ldw, ldc, epilogue, stream, nullptr, nullptr);
}

extern "C" int b200_linear_normw(const void* A, const void* W, const void* bias, const void* norm_w, void* C, float* row_sumsq,
                                 int row_sumsq_capacity, int* n_parts, int M, int N, int K, int64_t lda, int64_t ldw,
                                 int64_t ldc, void* stream) {
  if (!norm_w || !row_sumsq || !n_parts) return B200_ERR_ARG;
  if (row_sumsq_capacity < (N + 63) / 64) return B200_ERR_SHAPE;   // parts per row for the narrowest column tile (64)
  return linear_impl(A, W, bias, C, norm_w, M, N, K, lda, ldw, ldc, B200_EPI_NORMW, stream, row_sumsq, n_parts);
}

static int linear_impl(const void* A, const void* W, const void* bias, void* C, const void* gate, int M, int N, int K,
                       int64_t lda, int64_t ldw, int64_t ldc, int epilogue, void* stream, float* row_sumsq, int* n_parts_out) {
  using namespace b200;
  using namespace b200::linear;
  if (!A || !W || !C) return B200_ERR_ARG;
  const int bias_row = (epilogue & B200_EPI_ROW_BIAS) ? 1 : 0;
  epilogue &= ~B200_EPI_ROW_BIAS;
  if (epilogue < 0 || epilogue > 6) return B200_ERR_ARG;
  if (epilogue == B200_EPI_NORMW && (!gate || !row_sumsq || bias_row || N < 64)) return B200_ERR_ARG;
  if (M <= 0 || N <= 0 || K <= 0) return B200_ERR_SHAPE;
  if ((K % 8) || (lda % 8) || (ldw % 8)) return B200_ERR_ALIGN;
  if (epilogue == B200_EPI_BIAS_F32 ? (ldc % 4) : (ldc % 8)) return B200_ERR_ALIGN;
  if (reinterpret_cast<uintptr_t>(C) & 15) return B200_ERR_ALIGN;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return B200_ERR_ALIGN;
  if (gate && (reinterpret_cast<uintptr_t>(gate) & 15)) return B200_ERR_ALIGN;

  // CTA-pair (cta_group::2) kernel: B200_LINEAR_2CTA=1 forces it, =0 disables it; default: see use_pair below.
  static int pair_mode = -2;
  if (pair_mode == -2) {
    const char* ev = getenv("B200_LINEAR_2CTA");
    pair_mode = ev ? (ev[0] == '1' ? 1 : 0) : -1;
  }
  bool use_pair = pair_mode == 1 || (pair_mode == -1 && PAIR_BY_DEFAULT && M > BM);
  // small-M form: fewer than one wave of 128 x 256 tiles -> 128 x 64 tiles on single CTAs.  OFF by default (B200_LINEAR_SMALLM=1
  // enables): measured on the FLUX text-stream shapes it wins only 3-5 us where there are <= 48 tiles (512x3072x3072: 25.6 -> 20.6 us)
  // and loses where there are 144 (512x9216x3072: 27.2 -> 35.9 us) -- N = 64 MMAs and 4x the A re-reads eat the extra parallelism
  // (profiles/r01_gpu_session29_smallm.log).  Split-K is the better tool for these shapes.
  static int smallm_mode = -2;
  if (smallm_mode == -2) {
    const char* ev = getenv("B200_LINEAR_SMALLM");
    smallm_mode = ev ? (ev[0] == '1' ? 1 : 0) : -1;
  }
  const int tiles256 = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const bool small_m = smallm_mode == 1 && pair_mode != 1 && M <= 512 && M > 16 && tiles256 < num_sms() && N >= 256;
  if (small_m) use_pair = false;
  const int ncta = use_pair ? 2 : 1;
  const int bn = small_m ? 64 : BN;
  // QUAD (clusters of 4 = two pairs sharing A through TMA multicast).  A cluster of 4 can only use 132 of the 148 SMs (33 quads
  // fit: GPC sizes 16 / 18 / 20) yet measures >= the pair form on every big shape under the power cap (Wan qkv +2.5 %, ffn.2
  // +13 %, profiles/r02_gemm_quad_ab.log): default for problems of >= 1000 pair tiles, where 66 vs 74 tile slots per wave does
  // not matter.  B200_LINEAR_QUAD=1 / 0 forces / disables it.
  static int quad_mode = -2;
  if (quad_mode == -2) {
    const char* ev = getenv("B200_LINEAR_QUAD");
    quad_mode = ev ? (ev[0] == '1' ? 1 : 0) : -1;
  }
  const int64_t pair_tiles = static_cast<int64_t>((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  // TALL (512 x 256 tile per CTA pair, see Cfg): B200_LINEAR_TALL=1 / 0 forces / disables it
  static int tall_mode = -2;
  if (tall_mode == -2) {
    const char* ev = getenv("B200_LINEAR_TALL");
    tall_mode = ev ? (ev[0] == '1' ? 1 : 0) : -1;
  }
  const bool use_tall = use_pair && (tall_mode == 1 || (tall_mode == -1 && TALL_BY_DEFAULT && pair_tiles >= 1000));
  const bool use_quad = !use_tall && use_pair && ((N + BN - 1) / BN) >= 2 &&
                        (quad_mode == 1 || (quad_mode == -1 && pair_tiles >= 1000));
  const int rows_per_cta = use_tall ? 2 * BM : BM;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[2] = {1, (uint64_t)lda};
    uint32_t box[2] = {BK, static_cast<uint32_t>(use_quad ? BM / 2 : rows_per_cta)};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[2] = {1, (uint64_t)ldw};
    uint32_t box[2] = {BK, (uint32_t)(bn / ncta)};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box);
    if (rc) return rc;
  }
  // TMA-store epilogue for the bf16, non-residual epilogues (B200_LINEAR_TMA_STORE=0 disables)
  static int tma_store_mode = -2;
  if (tma_store_mode == -2) {
    const char* ev = getenv("B200_LINEAR_TMA_STORE");
    tma_store_mode = ev ? (ev[0] == '1' ? 1 : 0) : 1;
  }
  const bool use_tma_store = (tma_store_mode == 1 || epilogue == B200_EPI_NORMW) && epilogue != B200_EPI_GATE_RES &&
                             epilogue != B200_EPI_BIAS_F32 && N >= 64;   // NORMW lives in the staged epilogue only
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  if (use_tma_store) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str[2] = {1, (uint64_t)ldc};
    uint32_t box[2] = {64, 32};
    int rc = make_tmap_bf16(&tmC, C, 2, dims, str, box);
    if (rc) return rc;
  }
  Params p;
  p.tma_store = use_tma_store ? 1 : 0;
  p.c_vec32 = (epilogue != B200_EPI_BIAS_F32 && (ldc % 16) == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0) ? 1 : 0;
  p.M = M; p.N = N; p.K = K;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.C = C;
  p.ldc = ldc;
  p.epi = epilogue;
  p.bias_row = bias_row;
  p.row_sumsq = row_sumsq;
  p.n_parts = (N + bn - 1) / bn;
  if (n_parts_out) *n_parts_out = p.n_parts;
  p.tiles_m = (M + rows_per_cta * ncta - 1) / (rows_per_cta * ncta);
  p.tiles_n = (N + bn - 1) / bn;
  if (use_quad) p.tiles_n = (p.tiles_n + 1) / 2;      // column tile PAIRS
  {
    // rasterisation sized for L2: ~24 MB of A rows per group (4..16 row tiles), ~64 MB of W per panel (>= 8 column tiles), panels
    // of equal width (see tile_coords).
    // Measured effect under the power cap: within +-3 % of the round-1 order on every Wan shape (profiles/r02_gemm_raster_sweep.log)
    // -- HBM traffic drops 3x but the kernel is bound by energy per FLOP elsewhere.
    const double row_tile_bytes = 2.0 * rows_per_cta * ncta * K, col_tile_bytes = 2.0 * bn * (use_quad ? 2 : 1) * K;
    int gm = static_cast<int>(24.0e6 / row_tile_bytes);
    int pn = static_cast<int>(64.0e6 / col_tile_bytes);
    p.group_m = gm < 4 ? 4 : (gm > GROUP_M ? GROUP_M : gm);
    pn = pn < 8 ? 8 : pn;
    const int n_panels = (p.tiles_n + pn - 1) / pn;
    p.panel_n = (p.tiles_n + n_panels - 1) / n_panels;
    if (const char* ev = getenv("B200_LINEAR_GROUP_M")) p.group_m = atoi(ev) > 0 ? atoi(ev) : p.group_m;   // experiments
    if (const char* ev = getenv("B200_LINEAR_PANEL_N")) p.panel_n = atoi(ev) > 0 ? atoi(ev) : p.panel_n;
    if (p.panel_n > p.tiles_n) p.panel_n = p.tiles_n;
    if (p.group_m > p.tiles_m) p.group_m = p.tiles_m;
  }
  const int total = p.tiles_m * p.tiles_n;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  static std::atomic<bool> attr_done[kMaxDevices];
  if (!once_per_device(attr_done, [] {
        return cudaFuncSetAttribute(linear_kernel<1, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(linear_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1, 64>::SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(linear_kernel<2, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(linear_kernel<2, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(linear_kernel<2, BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2, BN, true>::SMEM_BYTES) == cudaSuccess;
      }))
    return B200_ERR_LAUNCH;
  if (use_quad) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg<2>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static std::atomic<int> quads_fit[kMaxDevices];
    const int dev = current_device();
    int fit = dev >= 0 ? quads_fit[dev].load() : 0;
    if (fit == 0) {
      cfg.gridDim = dim3(4 * (num_sms() / 4), 1, 1);
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, linear_kernel<2, BN, true>, &cfg) != cudaSuccess || ncl <= 0) ncl = num_sms() / 4 - 4;
      fit = ncl;
      if (dev >= 0) quads_fit[dev].store(fit);
      if (getenv("B200_LINEAR_DEBUG")) fprintf(stderr, "[apex_b200] linear quad kernel: %d clusters of 4 fit\n", fit);
    }
    const int quads = total < fit ? total : fit;
    cfg.gridDim = dim3(4 * quads, 1, 1);
    if (cudaLaunchKernelEx(&cfg, linear_kernel<2, BN, true>, tmA, tmB, tmC, p) != cudaSuccess) {
      cudaGetLastError();
      return B200_ERR_LAUNCH;
    }
  } else if (use_pair) {
    int pairs_avail = num_sms() / 2;
    if (const char* ev = getenv("B200_LINEAR_PAIRS")) pairs_avail = atoi(ev) > 0 ? atoi(ev) : pairs_avail;   // experiments
    const int pairs = total < pairs_avail ? total : pairs_avail;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = use_tall ? Cfg<2, BN, true>::SMEM_BYTES : Cfg<2>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static bool dbg_done = false;
    if (!dbg_done && getenv("B200_LINEAR_DEBUG")) {
      int ncl = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, linear_kernel<2, BN>, &cfg);
      fprintf(stderr, "[apex_b200] linear pair kernel: max active clusters %d (err %d), launching %d pairs\n", ncl, (int)e, pairs);
      dbg_done = true;
    }
    const cudaError_t le = use_tall ? cudaLaunchKernelEx(&cfg, linear_kernel<2, BN, false, true>, tmA, tmB, tmC, p)
                                    : cudaLaunchKernelEx(&cfg, linear_kernel<2, BN>, tmA, tmB, tmC, p);
    if (le != cudaSuccess) {
      cudaGetLastError();
      return B200_ERR_LAUNCH;
    }
  } else {
    const int grid = total < num_sms() ? total : num_sms();
    if (small_m) linear_kernel<1, 64><<<grid, NUM_THREADS, Cfg<1, 64>::SMEM_BYTES, st>>>(tmA, tmB, tmC, p);
    else linear_kernel<1, BN><<<grid, NUM_THREADS, Cfg<1>::SMEM_BYTES, st>>>(tmA, tmB, tmC, p);
  }
  B200_CHECK_LAUNCH();
  return B200_OK;
}

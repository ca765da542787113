// One CrossAttentionBLK (modules/attention.py:436-452) minus its causal self-attention core as ONE persistent
// tcgen05 kernel per 128-row tile of one utterance ("row kernel").  Everything in the block that is local to a
// query row -- given the projected text memory K/V of the utterance -- runs back to back on one SM with the
// activations resident in shared memory (fp16 UMMA operand panels) and the fp32 residual stream resident in TMEM:
//
//   s   = LN1( [x ; a1] Wp1 + b + x )                 attention.py:440-443   (a1 = causal self-attention context)
//   q   = s Wcq ; a2 = softmax(mask(q K^T / 8)) V     attention.py:217-246   (cross attention, 4 heads, S kept in TMEM)
//   c   = LN2( [s ; a2] Wp2 + b + s )                 attention.py:447-450
//   x'  = LN3( relu(c W1 + b1) W2 + b2 + c )          modules/utils.py:48-53 (hidden never leaves the SM)
//   q|k|v of the NEXT block's self-attention = x' Wqkv                       (attention.py:217-226 of block i+1)
//
// Data flow inside the CTA
//   * TMEM (512 columns): R = [0,256) fp32 residual / GEMM accumulator of the three LayerNorm inputs (x is stored
//     into R at kernel start, every Dense ACCUMULATES onto it, each LayerNorm writes its fp32 result back into R --
//     the residual never touches memory inside a block); scratch = [256,512): cq result, attention scores S (<=192
//     columns) + O (64), double-buffered 128-column FFN-hidden / next-QKV chunks.
//   * shared memory: ACT0 / ACT1 = 2 x [128 x 256] fp16 as four SW128 K-major panels each (A operands written by the
//     epilogue warps, or by TMA at kernel start); a 5-slot x 16 KB ring streams every B operand (weights [128 x 64],
//     memory-K [128 keys x 64], memory-V^T panels [64 x 64 keys]) with TMA.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..17 = epilogue / softmax (four column
//     groups per TMEM lane quadrant, thread == row).
// Training forward (tape): with the tp_* pointers of XRowParams set, the epilogue warps additionally store -- straight from the
// registers that already hold them, 64-256 contiguous bytes per thread, no change to the barrier protocol -- every tensor the
// backward pass of the block reads (LN1 / LN2 outputs fp32 + fp16 and their reciprocal standard deviations, the
// cross-attention queries, context and log2-sum-exp, the FFN hidden, LN3's reciprocal standard deviation); the block output
// goes to its own buffers (x_f_out, tensor map tmXo) so that the input survives, and the next block's q | k | v are written
// row-major with pitch 768 next to V^T.  The flow tail is inference-only.
// Masking semantics of the reference are kept: fully masked query rows attend uniformly over all T_text keys (their
// context is the column mean of V, alignments 1/T_text), masked keys of live rows get exactly 0.
#pragma once
#include "ptx.cuh"

namespace vb {

constexpr int XR_THREADS = 576;
constexpr int XR_EPI_THREADS = 512;
constexpr int XR_D = 256;        // block width
constexpr int XR_F = 1024;       // FFN hidden
constexpr int XR_H = 4;          // heads (x 64)
constexpr int XR_PANEL = 128 * 64 * 2;        // [128 rows x 64 fp16], 128-byte swizzle
constexpr int XR_ACT = 4 * XR_PANEL;          // 64 KB
constexpr int XR_SLOT = 16384;
#ifndef VB_XR_NSLOT
#define VB_XR_NSLOT 5
#endif
constexpr int XR_NSLOT = VB_XR_NSLOT;
constexpr int XR_NGRP = 8;                     // tile-group "full" barriers (power of two > groups in flight)
constexpr int XR_NGRP_LOG2 = 3;
constexpr int XR_EPI_WARPS = 16;
constexpr int XR_TK_MAX = 192;                // S (<= 192 columns) + O (64) share the 256 scratch columns
constexpr int XR_OFF_RING = 2 * XR_ACT;
constexpr int XR_OFF_BARS = XR_OFF_RING + XR_NSLOT * XR_SLOT;
constexpr int XR_OFF_RED = XR_OFF_BARS + 512;
constexpr int XR_NVEC = 9;                     // bias / gamma / beta vectors of the three LayerNorm stages, staged in smem
constexpr int XR_RED_FLOATS = 2 * 4 * 128 + XR_H * 64 + XR_NVEC * XR_D + XR_TK_MAX;
constexpr int XR_SMEM = XR_OFF_RED + XR_RED_FLOATS * 4 + 1024;

struct XRowParams {
  int T;                 // rows (frames) per utterance
  int Tt;                // memory (text) length
  int TKP;               // Tt rounded up to a multiple of 16 (<= XR_TK_MAX)
  int kv_col0;           // column of head 0 of this block's K inside the memory-K tensor
  int vt_row0;           // first V^T row of (batch 0, head 0) of this block
  const __half* vt;      // memory V^T base / row pitch (column mean for fully masked rows)
  int vt_ld;
  const int* q_len;      // [B]
  const int* k_len;      // [B]
  float scale;           // 1 / sqrt(64)
  float ln_eps;
  // [256] each: att_proj1 bias, layer_norm1 gamma, beta; att_proj2 bias, layer_norm2 gamma, beta; ffn dense2 bias,
  // ffn layer_norm gamma, beta
  const float* vec[XR_NVEC];
  const float* bf1;      // [1024] ffn dense1 bias
  float* x_f;            // [B*T, 256] fp32 residual stream (in / out)
  int has_next;          // tail bit 0: also project q|k|v of the next block (of this module, or of the next flow step's net)
  __half* qk_next;       // [B*T, 512] Q | K row-major
  __half* vt_next;       // [B*H*64, vt_next_ld] V transposed
  int vt_next_ld;
  // ---- flow tail of the LAST block of a coupling net (modules/flow.py:223-257, modules/prior.py:135-168): the block
  // output never leaves the SM; the coupling update of z, the next ActNorm (+) InvertibleLinear map (split-fp16 on the tensor
  // cores) and the next step's pre-projection + positional encoding run in the same kernel.
  int tail_coupling;     // z[:, zp_off : zp_off + 64] <- affine coupling with (log_scale, shift) = x' W_out + b
  int tail_flow;         // z <- z M + c  (M = folded 128 x 128 map of the NEXT flow operation)
  int tail_pre;          // x <- z[:, cond_off_next : +64] W_pre + b + pos_weight * PE[t]  (input of the next coupling net)
  float* z;              // [B*T, 128] fp32 flow state (in / out)
  int zp_off;            // transformed half of this step
  int cp_backward;       // 0: zp * scale + shift ; 1: (zp - shift) / (scale + 1e-12)
  const float* cp_bias;  // [128] log_scale_proj bias | shift_proj bias
  float* row_acc;        // [B*T] per-row log-det accumulator (+=)
  const float* fl_c;     // [128] offset of the folded flow map
  int cond_off_next;     // conditioning half of the next step
  const float* pre_bias; // [256]
  const float* pre_pw;   // [1] pos_weight of the next net
  const float* pe;       // [T, 256] positional-encoding table
  float* ali;            // optional [B, H, T, Tt] fp32 cross-attention alignments
  int mma_n256;          // 256-wide outputs: one N = 256 MMA per k-step when the two weight tiles sit in adjacent ring slots
  // ---- training forward: the block output goes to its own buffers and the tensors the backward pass needs are written on
  // the way (all optional; inference: x_f_out == x_f, the output tensor map == the input one, everything else null)
  float* x_f_out;        // [B*T, 256] fp32 block output (o_f of the tape)
  float* tp_s_f;         // LN1 output fp32 / fp16, reciprocal standard deviation
  __half* tp_s_h;
  float* tp_rstd1;
  __half* tp_q2;         // cross-attention queries [B*T, 256]
  float* tp_lse2;        // [B, H, T] log2 of the cross-attention softmax denominator incl. the maximum
  __half* tp_ctx2;       // cross-attention context [B*T, 256]
  float* tp_c_f;         // LN2 output
  __half* tp_c_h;
  float* tp_rstd2;
  __half* tp_hid;        // [B*T, 1024] FFN hidden (post-ReLU)
  float* tp_rstd3;
  int qk_next_ld;        // row pitch of qk_next: 512 (Q | K), or 768 when V row-major follows in the same rows (training)
  __half* v_rm_next;     // training: V of the next block row-major (pitch qk_next_ld), written next to V^T
  unsigned long long* dbg;   // optional per-CTA phase timestamps (tuning aid), 128 x u64 per CTA (globaltimer ns)
};

__global__ void __launch_bounds__(XR_THREADS, 1)   // 18 warps (allocated as 20): at most 96 registers per thread
xblk_row_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVt,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmWq,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmF1,
                const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmWn,
                const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmFl,
                const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmXo,
                const __grid_constant__ XRowParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* act0 = smem;
  uint8_t* act1 = smem + XR_ACT;
  uint8_t* ring = smem + XR_OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XR_OFF_BARS);
  uint64_t* empty = bars + 48;           // [5] per ring slot: free again (MMA -> TMA); every completion is waited exactly once
  uint64_t* a_full = bars + 10;          // x (fp16) panels landed in ACT0
  uint64_t* a1_full = bars + 11;         // self-attention context panels landed in ACT1
  uint64_t* r_init = bars + 12;          // residual stored into R (epilogue -> MMA)
  uint64_t* stage_free = bars + 13;      // ACT1 staging no longer used by the residual load (epilogue -> TMA)
  uint64_t* r_full = bars + 14;          // R holds a complete LayerNorm input (MMA -> epilogue), 3 phases
  uint64_t* act_ready = bars + 15;       // LayerNorm output in ACT0 (+R) (epilogue -> MMA), 3 phases
  uint64_t* sc_full = bars + 16;         // scratch holds cq / S_h (MMA -> epilogue), 5 phases
  uint64_t* p_ready = bars + 17;         // q panels / P_h written (epilogue -> MMA), 5 phases
  uint64_t* o_full = bars + 18;          // O_h complete (MMA -> epilogue), 4 phases
  uint64_t* o_done = bars + 19;          // O_h drained, a2 panel h written (epilogue -> MMA), 4 phases
  uint64_t* h_full = bars + 20;          // [2] scratch chunk buffer complete (MMA -> epilogue)
  uint64_t* hid_ready = bars + 22;       // [2] chunk drained (+ hidden panel written) (epilogue -> MMA)
  uint64_t* s_read = bars + 24;          // S_h copied into registers (epilogue -> MMA), 4 phases
  uint64_t* gfull = bars + 32;           // [8] per tile group, TMA -> MMA
  uint64_t* t_full = bars + 40;          // [3] flow tail: coupling / flow / pre accumulators complete (MMA -> epilogue)
  uint64_t* t_ready = bars + 43;         // [2] flow tail: z (hi | lo) panels / conditioning panel written (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  float* lred = reinterpret_cast<float*>(smem + XR_OFF_RED);   // [2][4][128] LayerNorm row statistics
  float* sred = lred;                                          // [2][4][128] softmax row statistics (never live together)
  float* vmean = lred + 2 * 4 * 128;                           // [4][64] column mean of V (fully masked rows)
  float* pvec = vmean + XR_H * 64;                             // [9][256] bias / gamma / beta vectors
  float* kbias = pvec + XR_NVEC * XR_D;                        // [192] additive key mask of this utterance (0 / -inf)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * 128;
  const int b = blockIdx.y;
  unsigned long long* dbg = p.dbg ? p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 128 : nullptr;
  auto stamp = [&](int i) {
    if (dbg) {
      unsigned long long tt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
      dbg[i] = tt;
    }
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < XR_NGRP; ++i) mbar_init(&gfull[i], 1);
    for (int i = 0; i < XR_NSLOT; ++i) mbar_init(&empty[i], 1);
    mbar_init(a_full, 1);
    mbar_init(a1_full, 1);
    mbar_init(r_init, XR_EPI_WARPS);
    mbar_init(stage_free, XR_EPI_WARPS);
    mbar_init(r_full, 1);
    mbar_init(act_ready, XR_EPI_WARPS);
    mbar_init(sc_full, 1);
    mbar_init(p_ready, XR_EPI_WARPS);
    mbar_init(o_full, 1);
    mbar_init(o_done, XR_EPI_WARPS);
    mbar_init(&h_full[0], 1);
    mbar_init(&h_full[1], 1);
    mbar_init(&hid_ready[0], XR_EPI_WARPS);
    mbar_init(&hid_ready[1], XR_EPI_WARPS);
    mbar_init(s_read, XR_EPI_WARPS);
    for (int i = 0; i < 3; ++i) mbar_init(&t_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&t_ready[i], XR_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmVt);
    tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmWq); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmF1);
    tma_prefetch_desc(&tmF2); tma_prefetch_desc(&tmWn);
    if (p.tail_coupling) { tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmFl); tma_prefetch_desc(&tmWp); }
    tma_prefetch_desc(&tmXo);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: only the threads that touch memory written by the preceding
  // kernels wait for them (TMA producer before the activation loads, epilogue warps before the residual load); the
  // weight stream -- packed long before, with fully serialising launches in between -- starts right away.

  const int TKP = p.TKP;
  const int nK = TKP > 128 ? 2 : 1;          // 128-key chunks of the memory K tile
  const int npan = (TKP + 63) >> 6;          // 64-key panels of P / V^T

  if (warp == 0) {
    // =========================================================== TMA producer
    // The ring streams 16 KB tiles; consecutive tiles form GROUPS (two weight tiles, or the K / V^T tiles of one head)
    // that share one "full" barrier, so that the MMA issuer pays one barrier wait (~180 cycles) per 32 KB instead of
    // per tile.  Slots are released per tile (one commit each); the producer waits for every release exactly once.
    if (elect_one()) {
      int f = 0;           // tile index (slot = f % XR_NSLOT)
      int g = 0;           // group index (full barrier = g % XR_NGRP)
      uint64_t* gbar = nullptr;
      auto begin_group = [&](uint32_t bytes) {
        gbar = &gfull[g & (XR_NGRP - 1)];
        mbar_arrive_expect_tx(gbar, bytes);
        ++g;
      };
      auto acquire = [&]() -> uint8_t* {   // slots are released per tile; the previous use of this slot is use (f / NSLOT) - 1
        if (f >= XR_NSLOT) mbar_wait(&empty[f % XR_NSLOT], ((f / XR_NSLOT) - 1) & 1);
        return ring + (f % XR_NSLOT) * XR_SLOT;
      };
      // two weight tiles [128 out-rows x 64 k] at (k0, n0) and (k1, n1)
      auto group_w = [&](const CUtensorMap* m, int k0, int n0, int k1, int n1) {
        begin_group(2 * XR_SLOT);
        uint8_t* dst = acquire();
        tma_load_2d(dst, m, gbar, k0, n0);
        ++f;
        dst = acquire();
        tma_load_2d(dst, m, gbar, k1, n1);
        ++f;
      };
      auto group_k = [&](int h) {
        begin_group(nK * XR_SLOT);
        for (int c = 0; c < nK; ++c) {
          uint8_t* dst = acquire();
          tma_load_3d(dst, &tmK, gbar, p.kv_col0 + h * 64, c * 128, b);
          ++f;
        }
      };
      auto group_v = [&](int h) {
        const int vrow = p.vt_row0 + (b * XR_H + h) * 64;
        begin_group(npan * 8192);
        for (int q0 = 0; q0 < npan; q0 += 2) {
          uint8_t* dst = acquire();
          for (int q = q0; q < min(npan, q0 + 2); ++q) tma_load_2d(dst + (q - q0) * 8192, &tmVt, gbar, q * 64, vrow);
          ++f;
        }
      };
      // att_proj1, x half of the concat (k-panels 0..3); the first two groups never block (ring empty)
      for (int kp = 0; kp < 2; ++kp) group_w(&tmW1, kp * 64, 0, kp * 64, 128);
      pdl_wait();
      mbar_arrive_expect_tx(a_full, XR_ACT);
      for (int pn = 0; pn < 4; ++pn) tma_load_3d(act0 + pn * XR_PANEL, &tmX, a_full, pn * 64, t0, b);
      stamp(96);
      mbar_wait(stage_free, 0);
      stamp(97);
      mbar_arrive_expect_tx(a1_full, XR_ACT);
      for (int pn = 0; pn < 4; ++pn) tma_load_3d(act1 + pn * XR_PANEL, &tmA1, a1_full, pn * 64, t0, b);
      for (int kp = 2; kp < 8; ++kp) group_w(&tmW1, kp * 64, 0, kp * 64, 128);
      stamp(98);
      // the cross-attention query projection, then att_proj2's s half (its MMAs run while the epilogue converts q)
      for (int kp = 0; kp < 4; ++kp) group_w(&tmWq, kp * 64, 0, kp * 64, 128);
      for (int kp = 0; kp < 4; ++kp) group_w(&tmW2, kp * 64, 0, kp * 64, 128);
      stamp(99);
      // cross attention: K_0, then per head K_{h+1}, V_h (S_{h+1} is issued before P_h V_h)
      group_k(0);
      for (int h = 0; h < XR_H; ++h) {
        if (h + 1 < XR_H) group_k(h + 1);
        group_v(h);
      }
      stamp(100);
      // att_proj2, context half
      for (int kp = 4; kp < 8; ++kp) group_w(&tmW2, kp * 64, 0, kp * 64, 128);
      stamp(101);
      // FFN: dense1 chunk j = hidden columns [128 j, 128 j + 128); dense2 consumes it as a K-slice
      auto fill_f1 = [&](int j) {
        group_w(&tmF1, 0, j * 128, 64, j * 128);
        group_w(&tmF1, 128, j * 128, 192, j * 128);
      };
      auto fill_f2 = [&](int j) {
        for (int kp = 0; kp < 2; ++kp) group_w(&tmF2, j * 128 + kp * 64, 0, j * 128 + kp * 64, 128);
      };
      fill_f1(0);
      fill_f1(1);
      for (int j = 0; j < 8; ++j) {
        fill_f2(j);
        if (j + 2 < 8) fill_f1(j + 2);
      }
      stamp(102);
      if (p.tail_coupling) {   // W_out [128 x 256]: two groups of two k-panels
        group_w(&tmWo, 0, 0, 64, 0);
        group_w(&tmWo, 128, 0, 192, 0);
      }
      if (p.tail_flow) {       // folded map, rows [0,128) = hi, [128,256) = lo: (hi x z_hi), (hi x z_lo), (lo x z_hi)
        group_w(&tmFl, 0, 0, 64, 0);
        group_w(&tmFl, 0, 0, 64, 0);
        group_w(&tmFl, 0, 128, 64, 128);
      }
      if (p.tail_pre) group_w(&tmWp, 0, 0, 0, 128);   // W_pre [256 x 64]: the two column halves
      if (p.has_next)
        for (int j = 0; j < 6; ++j) {
          group_w(&tmWn, 0, j * 128, 64, j * 128);
          group_w(&tmWn, 128, j * 128, 192, j * 128);
        }
      stamp(103);
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer
    if (elect_one()) {
      int f = 0, g = 0;
      const uint32_t R = tmem_base, SCR = tmem_base + 256;
      constexpr uint32_t idesc128 = umma_idesc_f16(128, 128);
      constexpr uint32_t idesc256 = umma_idesc_f16(128, 256);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64);
      const uint32_t a0 = smem_u32(act0), a1 = smem_u32(act1);
      const uint32_t ring_u32 = smem_u32(ring);
      auto wait_group = [&]() {
        mbar_wait(&gfull[g & (XR_NGRP - 1)], (g >> XR_NGRP_LOG2) & 1);
        tc_fence_after();
        ++g;
      };
      auto tile_addr = [&]() -> uint32_t { return ring_u32 + (f % XR_NSLOT) * XR_SLOT; };
      auto release_tile = [&]() {
        umma_commit(&empty[f % XR_NSLOT]);
        ++f;
      };
      // one tile = a [128 x 64] B operand: D[128 x 128] (+)= A_panel[128 x 64] * B^T, four K=16 steps
      auto mma_tile = [&](uint32_t a_panel, uint32_t d_tmem, bool acc) {
        const uint64_t ad = umma_desc_sw128(a_panel), bd = umma_desc_sw128(tile_addr());
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc128, (acc || k > 0) ? 1u : 0u);
        release_tile();
      };
      // group of the two column halves of a 256-wide output for one k-panel
      // When the group's two ring slots are adjacent in shared memory they form ONE [256 x 64] K-major operand (same
      // 8-row atom stride): one N = 256 instruction per k-step reads A once for both halves -- 12 KB of shared-memory
      // operand traffic per 128 x 256 x 16 instead of 16 KB (the SS-operand MMA rate is bound by the shared-memory port).
      auto mma_pair_n = [&](uint32_t a_panel, uint32_t d_tmem, bool acc) {
        wait_group();
        if (p.mma_n256 && (f % XR_NSLOT) + 1 < XR_NSLOT) {
          const uint64_t ad = umma_desc_sw128(a_panel), bd = umma_desc_sw128(tile_addr());
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc256, (acc || k > 0) ? 1u : 0u);
          release_tile();
          release_tile();
          return;
        }
        mma_tile(a_panel, d_tmem, acc);
        mma_tile(a_panel, d_tmem + 128, acc);
      };
      // group of two consecutive k-panels of a 128-wide output
      auto mma_pair_k = [&](uint32_t a_panel, uint32_t d_tmem, bool acc) {
        wait_group();
        mma_tile(a_panel, d_tmem, acc);
        mma_tile(a_panel + XR_PANEL, d_tmem, true);
      };
      // ---- 1. R (= x) += [x ; a1] Wp1
      mbar_wait(a_full, 0);
      mbar_wait(r_init, 0);
      tc_fence_after();
      stamp(64);
      for (int kp = 0; kp < 4; ++kp) mma_pair_n(a0 + kp * XR_PANEL, R, true);
      stamp(65);
      mbar_wait(a1_full, 0);
      tc_fence_after();
      stamp(66);
      for (int kp = 0; kp < 4; ++kp) mma_pair_n(a1 + kp * XR_PANEL, R, true);
      umma_commit(r_full);
      stamp(67);
      // ---- 2. scratch = s Wcq ; R (= s) += s Wp2[:256]
      mbar_wait(act_ready, 0);
      tc_fence_after();
      stamp(68);
      for (int kp = 0; kp < 4; ++kp) mma_pair_n(a0 + kp * XR_PANEL, SCR, kp > 0);
      umma_commit(sc_full);
      for (int kp = 0; kp < 4; ++kp) mma_pair_n(a0 + kp * XR_PANEL, R, true);
      stamp(69);
      // ---- 3. cross attention, S_h = Q_h K_h^T into scratch [0, TKP), O_h = P_h V_h into scratch [192, 256)
      auto mma_s = [&](int h) {
        wait_group();
        for (int c = 0; c < nK; ++c) {
          const uint32_t n = c == 0 ? static_cast<uint32_t>(min(128, TKP)) : static_cast<uint32_t>(TKP - 128);
          const uint32_t idesc = umma_idesc_f16(128, n);
          const uint64_t ad = umma_desc_sw128(a1 + h * XR_PANEL), bd = umma_desc_sw128(tile_addr());
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(SCR + c * 128, ad + 2 * k, bd + 2 * k, idesc, k > 0 ? 1u : 0u);
          release_tile();
        }
        umma_commit(sc_full);
      };
      mbar_wait(p_ready, 0);   // q panels
      tc_fence_after();
      stamp(70);
      mma_s(0);
      for (int h = 0; h < XR_H; ++h) {
        if (h + 1 < XR_H) {
          mbar_wait(s_read, h & 1);   // S_h copied into registers: the score columns are free
          tc_fence_after();
          mma_s(h + 1);
        }
        mbar_wait(p_ready, (h + 1) & 1);   // P_h in ACT0
        if (h > 0) mbar_wait(o_done, (h - 1) & 1);   // O_{h-1} drained
        tc_fence_after();
        stamp(71 + 2 * h);
        wait_group();   // V^T_h
        const int nsteps = TKP >> 4;
        for (int kk = 0; kk < nsteps; ++kk) {
          const int panel = kk >> 2;
          if (kk > 0 && (kk & 7) == 0) release_tile();
          const uint64_t ad = umma_desc_sw128(a0 + panel * XR_PANEL) + 2 * (kk & 3);
          const uint64_t bd = umma_desc_sw128(tile_addr() + (panel & 1) * 8192) + 2 * (kk & 3);
          umma_f16(SCR + 192, ad, bd, idesc_o, kk > 0 ? 1u : 0u);
        }
        release_tile();
        umma_commit(o_full);
        stamp(72 + 2 * h);
      }
      mbar_wait(o_done, (XR_H - 1) & 1);   // a2 complete in ACT1
      tc_fence_after();
      stamp(79);
      // ---- 4. R += a2 Wp2[256:]
      for (int kp = 0; kp < 4; ++kp) mma_pair_n(a1 + kp * XR_PANEL, R, true);
      umma_commit(r_full);
      stamp(80);
      // ---- 5. FFN: hidden chunk j -> scratch buffer (j & 1); R (= c) += hidden_j W2[128 j : 128 j + 128]
      mbar_wait(act_ready, 1);
      tc_fence_after();
      stamp(81);
      auto f1 = [&](int j) {
        mma_pair_k(a0, SCR + (j & 1) * 128, false);
        mma_pair_k(a0 + 2 * XR_PANEL, SCR + (j & 1) * 128, true);
        umma_commit(&h_full[j & 1]);
      };
      auto f2 = [&](int j) {
        mbar_wait(&hid_ready[j & 1], (j >> 1) & 1);
        tc_fence_after();
        for (int kp = 0; kp < 2; ++kp) mma_pair_n(a1 + ((j & 1) * 2 + kp) * XR_PANEL, R, true);
      };
      f1(0);
      f1(1);
      for (int j = 0; j < 8; ++j) {
        f2(j);
        if (j + 2 < 8) f1(j + 2);
      }
      umma_commit(r_full);
      stamp(82);
      // ---- 6a. flow tail: coupling projection, folded flow map (split-fp16), next pre-projection
      uint32_t x_par = 0;   // parity of the act_ready phase that announces the A operand of the next stage
      if (p.tail_coupling) {
        mbar_wait(act_ready, x_par);
        x_par ^= 1;
        tc_fence_after();
        mma_pair_k(a0, SCR, false);
        mma_pair_k(a0 + 2 * XR_PANEL, SCR, true);
        umma_commit(&t_full[0]);
      }
      if (p.tail_flow) {
        mbar_wait(&t_ready[0], 0);   // z as fp16 hi (ACT1 panels 0, 1) and lo (panels 2, 3)
        tc_fence_after();
        mma_pair_k(a1, SCR + 128, false);                  // z_hi M_hi
        mma_pair_k(a1 + 2 * XR_PANEL, SCR + 128, true);    // z_lo M_hi
        mma_pair_k(a1, SCR + 128, true);                   // z_hi M_lo
        umma_commit(&t_full[1]);
      }
      if (p.tail_pre) {
        mbar_wait(&t_ready[1], 0);   // conditioning half of the new z in ACT1 panel 0
        tc_fence_after();
        mma_pair_n(a1, R, false);
        umma_commit(&t_full[2]);
      }
      // ---- 6. q|k|v of the next block: six 128-column chunks through the scratch buffers
      if (p.has_next) {
        mbar_wait(act_ready, x_par);
        tc_fence_after();
        stamp(83);
        for (int j = 0; j < 6; ++j) {
          if (j >= 2) {   // chunk j-2 drained (hid_ready use index 4 + ((j-2) >> 1))
            mbar_wait(&hid_ready[j & 1], ((j - 2) >> 1) & 1);
            tc_fence_after();
          }
          mma_pair_k(a0, SCR + (j & 1) * 128, false);
          mma_pair_k(a0 + 2 * XR_PANEL, SCR + (j & 1) * 128, true);
          umma_commit(&h_full[j & 1]);
        }
        stamp(84);
      }
    }
  } else {
    // =========================================================== epilogue / softmax warps
    const int quad = warp & 3;                  // TMEM lane quadrant (hardware rule: warp id % 4)
    const int grp = (warp - 2) >> 2;            // column group 0..3
    const int ew = warp - 2;                    // 0..15
    const int r = quad * 32 + lane;             // row of the tile == TMEM lane
    const int t = t0 + r;
    const bool row_ok = t < p.T;
    const uint32_t R = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t SCR = R + 256;
    const long grow0 = static_cast<long>(b) * p.T + t0 + quad * 32;   // first global row of this warp
    const int rows_here = min(32, p.T - (t0 + quad * 32));            // valid rows of this warp (may be <= 0)
    uint8_t* slab = act1 + ew * 4096;           // warp-private staging: 32 rows x 128 B, chunk ^= row & 7
    auto sw = [&](int row, int chunk) -> uint4* {
      return reinterpret_cast<uint4*>(slab + row * 128 + ((chunk ^ (row & 7)) << 4));
    };
    auto bar_all = []() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    // one arrival per warp: every lane has issued its fences, __syncwarp orders them before lane 0's release-arrive
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    uint32_t v[32];
    const bool st = (ew == 0 && lane == 0);
    auto estamp = [&](int i) { if (st) stamp(i); };

    // bias / gamma / beta vectors -> shared memory (weights: no dependency on the preceding kernels)
    for (int i = ew * 32 + lane; i < XR_NVEC * XR_D; i += XR_EPI_THREADS) pvec[i] = __ldg(p.vec[i >> 8] + (i & 255));
    pdl_wait();
    estamp(0);

    // ---- 0. residual stream x (fp32) -> R, coalesced through the staging slab (all loads in flight first)
    {
      uint4 pre[2][8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), chunk = lane & 7;
          pre[hh][it] = make_uint4(0u, 0u, 0u, 0u);
          if (row < rows_here)
            pre[hh][it] = *reinterpret_cast<const uint4*>(p.x_f + (grow0 + row) * XR_D + grp * 64 + hh * 32 + chunk * 4);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) *sw(it * 4 + (lane >> 3), lane & 7) = pre[hh][it];
        if (hh == 0) estamp(118);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 q = *sw(lane, g);
          v[g * 4 + 0] = q.x; v[g * 4 + 1] = q.y; v[g * 4 + 2] = q.z; v[g * 4 + 3] = q.w;
        }
        tmem_st32(R + grp * 64 + hh * 32, v);
      }
    }
    estamp(119);
    tmem_wait_st();
    fence_proxy_async_smem();   // generic accesses of the slab are ordered before the TMA write of a1 into ACT1
    tc_fence_before();
    warp_arrive(r_init);
    warp_arrive(stage_free);
    estamp(1);

    const int qlen = __ldg(p.q_len + b);
    const int klen = __ldg(p.k_len + b);
    const bool row_dead = (t >= qlen) || (klen <= 0);   // fully masked query row -> uniform attention
    if (max(t0, qlen) < min(t0 + 128, p.T) || klen <= 0) {
      // column mean of V over ALL Tt keys for every head (attention.py:240-242), two threads per channel
      const int idx = ew * 32 + lane;
      const int hd = idx >> 1, part = idx & 1;
      const __half* vr = p.vt + (static_cast<long>(p.vt_row0) + static_cast<long>(b) * XR_H * 64 + hd) * p.vt_ld;
      float acc = 0.f;
      for (int tt = part; tt < p.Tt; tt += 2) acc += __half2float(vr[tt]);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (part == 0) vmean[hd] = acc / static_cast<float>(p.Tt);
    }
    for (int i = ew * 32 + lane; i < XR_TK_MAX; i += XR_EPI_THREADS) kbias[i] = i < klen ? 0.f : -INFINITY;   // key mask, additive
    bar_all();   // pvec / vmean / kbias visible to every epilogue warp

    // LayerNorm of R (+ bias) over the 256 columns; fp16 result -> ACT0 panel `grp`; fp32 result back into R, or
    // (final) to global memory.  vi = index of the bias vector in pvec (gamma, beta follow).
    // x' (fp16 panels in ACT0 already written + fenced by every thread) -> global by TMA store; x' (fp32, in xu) through the slab
    auto store_x_global = [&](const uint32_t* xu) {
      bar_all();   // all four panels written and fenced
      if (st) {
        for (int pn = 0; pn < 4; ++pn) tma_store_3d(&tmXo, act0 + pn * XR_PANEL, pn * 64, t0, b);
        tma_store_commit();
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *sw(lane, g) = make_uint4(xu[hh * 32 + g * 4 + 0], xu[hh * 32 + g * 4 + 1], xu[hh * 32 + g * 4 + 2],
                                    xu[hh * 32 + g * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), chunk = lane & 7;
          if (row < rows_here)
            *reinterpret_cast<uint4*>(p.x_f_out + (grow0 + row) * XR_D + grp * 64 + hh * 32 + chunk * 4) = *sw(row, chunk);
        }
      }
    };
    // out_mode 0: fp32 result back into R (residual of the next stage); 1: block output -> global; 2: block output stays on chip
    const long grow = static_cast<long>(b) * p.T + t;   // this thread's global row (valid when row_ok)
    auto ln_epi = [&](int vi, int out_mode, float* tape_f, __half* tape_h, float* tape_rstd) {
      const bool final_out = out_mode != 0;
      const float* bias = pvec + vi * XR_D + grp * 64;
      const float* gamma = bias + XR_D;
      const float* beta = gamma + XR_D;
      float xs[64];
      uint32_t* xu = reinterpret_cast<uint32_t*>(xs);
      tmem_ld32(R + grp * 64, xu);
      tmem_ld32(R + grp * 64 + 32, xu + 32);
      tmem_wait_ld();
      if (vi == 3) estamp(110);
      float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const float4 bq = *reinterpret_cast<const float4*>(bias + g * 4);
        const float x0 = xs[g * 4 + 0] + bq.x, x1 = xs[g * 4 + 1] + bq.y, x2 = xs[g * 4 + 2] + bq.z, x3 = xs[g * 4 + 3] + bq.w;
        s4[0] += x0; s4[1] += x1; s4[2] += x2; s4[3] += x3;
        q4[0] = fmaf(x0, x0, q4[0]); q4[1] = fmaf(x1, x1, q4[1]); q4[2] = fmaf(x2, x2, q4[2]); q4[3] = fmaf(x3, x3, q4[3]);
        xs[g * 4 + 0] = x0; xs[g * 4 + 1] = x1; xs[g * 4 + 2] = x2; xs[g * 4 + 3] = x3;
      }
      bar_all();   // the statistics scratch is shared with the softmax (ordering would otherwise only follow from the mbarrier chain)
      lred[grp * 128 + r] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      lred[512 + grp * 128 + r] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      if (vi == 3) estamp(111);
      bar_all();
      if (vi == 3) estamp(112);
      const float tot = (lred[r] + lred[128 + r]) + (lred[256 + r] + lred[384 + r]);
      const float tsq = (lred[512 + r] + lred[640 + r]) + (lred[768 + r] + lred[896 + r]);
      const float mean = tot * (1.f / XR_D);
      const float rstd = rsqrtf(fmaxf(tsq * (1.f / XR_D) - mean * mean, 0.f) + p.ln_eps);
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const float4 gq = *reinterpret_cast<const float4*>(gamma + g * 4);
        const float4 bq = *reinterpret_cast<const float4*>(beta + g * 4);
        xs[g * 4 + 0] = (xs[g * 4 + 0] - mean) * rstd * gq.x + bq.x;
        xs[g * 4 + 1] = (xs[g * 4 + 1] - mean) * rstd * gq.y + bq.y;
        xs[g * 4 + 2] = (xs[g * 4 + 2] - mean) * rstd * gq.z + bq.z;
        xs[g * 4 + 3] = (xs[g * 4 + 3] - mean) * rstd * gq.w + bq.w;
      }
      if (vi == 3) estamp(113);
      if (tape_rstd && grp == 0 && row_ok) tape_rstd[grow] = rstd;
      if (tape_f && row_ok) {   // 256 contiguous bytes per thread
        float4* dst = reinterpret_cast<float4*>(tape_f + grow * XR_D + grp * 64);
#pragma unroll
        for (int g = 0; g < 16; ++g) dst[g] = make_float4(xs[g * 4 + 0], xs[g * 4 + 1], xs[g * 4 + 2], xs[g * 4 + 3]);
      }
      if (!final_out) {
        tmem_st32(R + grp * 64, xu);
        tmem_st32(R + grp * 64 + 32, xu + 32);
      }
      if (vi == 3) estamp(114);
      uint8_t* prow = act0 + grp * XR_PANEL + r * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_half2(xs[g * 8 + 0], xs[g * 8 + 1]);
        u.y = pack_half2(xs[g * 8 + 2], xs[g * 8 + 3]);
        u.z = pack_half2(xs[g * 8 + 4], xs[g * 8 + 5]);
        u.w = pack_half2(xs[g * 8 + 6], xs[g * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = u;
        if (tape_h && row_ok) *reinterpret_cast<uint4*>(tape_h + grow * XR_D + grp * 64 + g * 8) = u;
      }
      if (vi == 3) estamp(115);
      if (!final_out) tmem_wait_st();
      if (vi == 3) estamp(116);
      fence_proxy_async_smem();
      if (vi == 3) estamp(117);
      tc_fence_before();
      warp_arrive(act_ready);
      if (out_mode == 1) store_x_global(xu);
    };

    // ---- 1. s = LN1(R)
    mbar_wait(r_full, 0);
    tc_fence_after();
    estamp(2);
    ln_epi(0, 0, p.tp_s_f, p.tp_s_h, p.tp_rstd1);
    estamp(3);

    // ---- 2. cross-attention queries: scratch -> fp16 -> ACT1 panel (== head) `grp`
    {
      uint32_t w[32];
      mbar_wait(sc_full, 0);
      tc_fence_after();
      estamp(4);
      tmem_ld32(SCR + grp * 64, v);
      tmem_ld32(SCR + grp * 64 + 32, w);
      tmem_wait_ld();
      uint8_t* prow = act1 + grp * XR_PANEL + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u, u2;
        u.x = pack_half2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
        u.y = pack_half2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
        u.z = pack_half2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
        u.w = pack_half2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
        u2.x = pack_half2(__uint_as_float(w[g * 8 + 0]), __uint_as_float(w[g * 8 + 1]));
        u2.y = pack_half2(__uint_as_float(w[g * 8 + 2]), __uint_as_float(w[g * 8 + 3]));
        u2.z = pack_half2(__uint_as_float(w[g * 8 + 4]), __uint_as_float(w[g * 8 + 5]));
        u2.w = pack_half2(__uint_as_float(w[g * 8 + 6]), __uint_as_float(w[g * 8 + 7]));
        *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = u;
        *reinterpret_cast<uint4*>(prow + (((4 + g) ^ (r & 7)) << 4)) = u2;
        if (p.tp_q2 && row_ok) {
          *reinterpret_cast<uint4*>(p.tp_q2 + grow * XR_D + grp * 64 + g * 8) = u;
          *reinterpret_cast<uint4*>(p.tp_q2 + grow * XR_D + grp * 64 + 32 + g * 8) = u2;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      warp_arrive(p_ready);
      estamp(5);
    }

    // ---- 3. cross attention.  This thread owns the 8-key units [u0, u1) of row r.  The softmax of head h+1 runs while
    //         the tensor core computes P_h V_h: S_{h+1} is issued as soon as S_h sits in registers (s_read).
    {
      const int U8 = TKP >> 3;
      const int per = (U8 + 3) >> 2;             // <= 6
      const int u0 = grp * per;
      const int u1 = min(U8, u0 + per);
      const float sl2 = p.scale * 1.4426950408889634f;   // scale * log2(e)
      const float inv_tk = 1.0f / static_cast<float>(p.Tt);
      const bool ali_vec = (p.Tt & 3) == 0;
      float sv[6][8];
      // scores of head h -> unnormalised probabilities in sv; returns 1 / denominator (0 for a fully masked row)
      auto sm_compute = [&](int h) -> float {
        mbar_wait(sc_full, (h + 1) & 1);
        tc_fence_after();
        estamp(6 + 4 * h);
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (u0 + j < u1) tmem_ld8(SCR + (u0 + j) * 8, reinterpret_cast<uint32_t*>(sv[j]));
        tmem_wait_ld();
        if (h == 1) estamp(105);
        tc_fence_before();
        warp_arrive(s_read);
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (u0 + j < u1) {
            const float4 ka = *reinterpret_cast<const float4*>(kbias + (u0 + j) * 8);
            const float4 kb = *reinterpret_cast<const float4*>(kbias + (u0 + j) * 8 + 4);
            sv[j][0] += ka.x; sv[j][1] += ka.y; sv[j][2] += ka.z; sv[j][3] += ka.w;
            sv[j][4] += kb.x; sv[j][5] += kb.y; sv[j][6] += kb.z; sv[j][7] += kb.w;
#pragma unroll
            for (int e = 0; e < 8; ++e) m = fmaxf(m, sv[j][e]);
          }
        }
        sred[grp * 128 + r] = m;
        if (h == 1) estamp(106);
        bar_all();
        if (h == 1) estamp(107);
        m = fmaxf(fmaxf(sred[r], sred[128 + r]), fmaxf(sred[256 + r], sred[384 + r]));
        const float msl2 = row_dead ? INFINITY : m * sl2;   // a live row has key 0 unmasked: m is finite; dead row: p = 0
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (u0 + j < u1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pe = ex2_approx(fmaf(sv[j][e], sl2, -msl2));   // ex2(-inf) = 0: masked keys, dead rows
              l += pe;
              sv[j][e] = pe;
            }
          }
        }
        sred[512 + grp * 128 + r] = l;
        if (h == 1) estamp(108);
        bar_all();
        if (h == 1) estamp(109);
        l = (sred[512 + r] + sred[640 + r]) + (sred[768 + r] + sred[896 + r]);
        if (p.tp_lse2 && grp == 0 && row_ok)   // same statistic as attention_tc_kernel writes for the backward pass
          p.tp_lse2[(static_cast<long>(b) * XR_H + h) * p.T + t] = row_dead ? 0.f : m * sl2 + log2f(l);
        return row_dead ? 0.f : 1.0f / l;
      };
      // P_h (fp16, unnormalised) -> ACT0 panels; alignments of head h (normalised, fp32) -> global
      auto write_p = [&](int h, float inv_l) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (u0 + j < u1) {
            const int k0 = (u0 + j) * 8;
            uint4 u;
            u.x = pack_half2(sv[j][0], sv[j][1]); u.y = pack_half2(sv[j][2], sv[j][3]);
            u.z = pack_half2(sv[j][4], sv[j][5]); u.w = pack_half2(sv[j][6], sv[j][7]);
            *reinterpret_cast<uint4*>(act0 + (k0 >> 6) * XR_PANEL + r * 128 + ((((k0 & 63) >> 3) ^ (r & 7)) << 4)) = u;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        warp_arrive(p_ready);
        estamp(7 + 4 * h);
        if (p.ali && row_ok) {
          float* arow = p.ali + ((static_cast<long>(b) * XR_H + h) * p.T + t) * p.Tt;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (u0 + j < u1) {
#pragma unroll
              for (int e4 = 0; e4 < 2; ++e4) {
                const int kk = (u0 + j) * 8 + e4 * 4;
                float4 o;
                o.x = row_dead ? inv_tk : sv[j][e4 * 4 + 0] * inv_l;
                o.y = row_dead ? inv_tk : sv[j][e4 * 4 + 1] * inv_l;
                o.z = row_dead ? inv_tk : sv[j][e4 * 4 + 2] * inv_l;
                o.w = row_dead ? inv_tk : sv[j][e4 * 4 + 3] * inv_l;
                if (ali_vec && kk + 4 <= p.Tt) {
                  *reinterpret_cast<float4*>(arow + kk) = o;
                } else {
                  if (kk + 0 < p.Tt) arow[kk + 0] = o.x;
                  if (kk + 1 < p.Tt) arow[kk + 1] = o.y;
                  if (kk + 2 < p.Tt) arow[kk + 2] = o.z;
                  if (kk + 3 < p.Tt) arow[kk + 3] = o.w;
                }
              }
            }
          }
        }
      };
      // context of head h: 16 of the 64 channels per thread -> ACT1 panel h (the Q_h panel is dead by now)
      auto o_epi = [&](int h, float inv_l) {
        mbar_wait(o_full, h & 1);
        tc_fence_after();
        estamp(8 + 4 * h);
        uint32_t vv[16];
        tmem_ld16(SCR + 192 + grp * 16, vv);
        tmem_wait_ld();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = row_dead ? vmean[h * 64 + grp * 16 + e] : __uint_as_float(vv[e]) * inv_l;
        uint8_t* prow = act1 + h * XR_PANEL + r * 128;
        uint4 ua, ub;
        ua.x = pack_half2(f[0], f[1]); ua.y = pack_half2(f[2], f[3]); ua.z = pack_half2(f[4], f[5]); ua.w = pack_half2(f[6], f[7]);
        ub.x = pack_half2(f[8], f[9]); ub.y = pack_half2(f[10], f[11]); ub.z = pack_half2(f[12], f[13]); ub.w = pack_half2(f[14], f[15]);
        *reinterpret_cast<uint4*>(prow + (((grp * 2) ^ (r & 7)) << 4)) = ua;
        *reinterpret_cast<uint4*>(prow + (((grp * 2 + 1) ^ (r & 7)) << 4)) = ub;
        if (p.tp_ctx2 && row_ok) {
          *reinterpret_cast<uint4*>(p.tp_ctx2 + grow * XR_D + h * 64 + grp * 16) = ua;
          *reinterpret_cast<uint4*>(p.tp_ctx2 + grow * XR_D + h * 64 + grp * 16 + 8) = ub;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        warp_arrive(o_done);
        estamp(9 + 4 * h);
      };
      bar_all();   // LayerNorm statistics scratch -> softmax statistics scratch
      float il = sm_compute(0);
      write_p(0, il);
      for (int h = 0; h < XR_H; ++h) {
        float il_next = 0.f;
        if (h + 1 < XR_H) il_next = sm_compute(h + 1);   // overlaps P_h V_h on the tensor core
        o_epi(h, il);                                    // o_full(h): P_h has been consumed, the P buffer is free
        if (h + 1 < XR_H) write_p(h + 1, il_next);
        il = il_next;
      }
    }

    // ---- 4. c = LN2(R)
    mbar_wait(r_full, 1);
    tc_fence_after();
    estamp(22);
    ln_epi(3, 0, p.tp_c_f, p.tp_c_h, p.tp_rstd2);
    estamp(23);

    // ---- 5. FFN hidden chunks: relu(acc + b1) -> fp16 -> ACT1 buffer (j & 1), two panels of 64 hidden columns
    for (int j = 0; j < 8; ++j) {
      const int bi = j & 1;
      float4 bq[8];
      const float* bj = p.bf1 + j * 128 + grp * 32;
#pragma unroll
      for (int g = 0; g < 8; ++g) bq[g] = __ldg(reinterpret_cast<const float4*>(bj + g * 4));
      mbar_wait(&h_full[bi], (j >> 1) & 1);
      tc_fence_after();
      estamp(24 + 2 * j);
      tmem_ld32(SCR + bi * 128 + grp * 32, v);
      tmem_wait_ld();
      uint8_t* prow = act1 + (bi * 2 + (grp >> 1)) * XR_PANEL + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 ba = bq[2 * g], bb = bq[2 * g + 1];
        uint4 u;
        u.x = pack_half2(fmaxf(__uint_as_float(v[g * 8 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[g * 8 + 1]) + ba.y, 0.f));
        u.y = pack_half2(fmaxf(__uint_as_float(v[g * 8 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[g * 8 + 3]) + ba.w, 0.f));
        u.z = pack_half2(fmaxf(__uint_as_float(v[g * 8 + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(v[g * 8 + 5]) + bb.y, 0.f));
        u.w = pack_half2(fmaxf(__uint_as_float(v[g * 8 + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(v[g * 8 + 7]) + bb.w, 0.f));
        *reinterpret_cast<uint4*>(prow + ((((grp & 1) * 4 + g) ^ (r & 7)) << 4)) = u;
        if (p.tp_hid && row_ok) *reinterpret_cast<uint4*>(p.tp_hid + grow * XR_F + j * 128 + grp * 32 + g * 8) = u;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      warp_arrive(&hid_ready[bi]);
      estamp(25 + 2 * j);
    }

    // ---- 6. x' = LN3(R): fp16 -> ACT0 (A operand of the next QKV) and, by TMA store, -> global; fp32 -> global
    mbar_wait(r_full, 0);
    tc_fence_after();
    estamp(40);
    ln_epi(6, p.tail_coupling ? 2 : 1, nullptr, nullptr, p.tp_rstd3);
    estamp(41);
    pdl_launch_dependents();   // late trigger: a parked dependent grid would only block SMs the other launch chain needs

    // ---- 6a. flow tail (last block of a coupling net)
    if (p.tail_coupling) {
      // affine coupling (modules/flow.py:223-257): thread (r, grp) owns 16 of the 64 transformed channels and, for the
      // flow map that follows, the matching 16 channels of the conditioning half
      const int cond_off = 64 - p.zp_off;
      float zp[16], zc[16];
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c4 = a;
        if (row_ok) {
          a = *reinterpret_cast<const float4*>(p.z + grow * 128 + p.zp_off + grp * 16 + e4 * 4);
          if (p.tail_flow) c4 = *reinterpret_cast<const float4*>(p.z + grow * 128 + cond_off + grp * 16 + e4 * 4);
        }
        zp[e4 * 4 + 0] = a.x; zp[e4 * 4 + 1] = a.y; zp[e4 * 4 + 2] = a.z; zp[e4 * 4 + 3] = a.w;
        zc[e4 * 4 + 0] = c4.x; zc[e4 * 4 + 1] = c4.y; zc[e4 * 4 + 2] = c4.z; zc[e4 * 4 + 3] = c4.w;
      }
      mbar_wait(&t_full[0], 0);
      tc_fence_after();
      uint32_t ls[16], sh[16];
      tmem_ld16(SCR + grp * 16, ls);
      tmem_ld16(SCR + 64 + grp * 16, sh);
      tmem_wait_ld();
      float logdet = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float l_ = __uint_as_float(ls[e]) + __ldg(p.cp_bias + grp * 16 + e);
        const float s_ = __uint_as_float(sh[e]) + __ldg(p.cp_bias + 64 + grp * 16 + e);
        const float scale = __fdividef(1.f, 1.f + __expf(-(l_ + 2.0f)));   // sigmoid(log_scale + 2), flow.py:231
        zp[e] = p.cp_backward ? __fdividef(zp[e] - s_, scale + 1e-12f) : scale * zp[e] + s_;
        logdet += __logf(scale);
      }
      bar_all();
      sred[grp * 128 + r] = logdet;
      bar_all();
      if (grp == 0 && row_ok) {
        const float tot = (sred[r] + sred[128 + r]) + (sred[256 + r] + sred[384 + r]);
        if (t < qlen) p.row_acc[grow] += p.cp_backward ? -tot : tot;
      }
      if (!p.tail_flow) {
        if (row_ok) {
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4)
            *reinterpret_cast<float4*>(p.z + grow * 128 + p.zp_off + grp * 16 + e4 * 4) =
                make_float4(zp[e4 * 4 + 0], zp[e4 * 4 + 1], zp[e4 * 4 + 2], zp[e4 * 4 + 3]);
        }
      } else {
        // the updated z row as split fp16 (hi | lo) A operands: z column c lives in panel (c >> 6) (hi) / 2 + (c >> 6) (lo)
        auto put = [&](const float* vals, int col0) {   // 16 consecutive z columns starting at col0 (multiple of 16)
          float lo[16];
          uint4 h0, h1, l0, l1;
#pragma unroll
          for (int e = 0; e < 16; ++e) lo[e] = vals[e] - __half2float(__float2half_rn(vals[e]));
          h0.x = pack_half2(vals[0], vals[1]); h0.y = pack_half2(vals[2], vals[3]); h0.z = pack_half2(vals[4], vals[5]); h0.w = pack_half2(vals[6], vals[7]);
          h1.x = pack_half2(vals[8], vals[9]); h1.y = pack_half2(vals[10], vals[11]); h1.z = pack_half2(vals[12], vals[13]); h1.w = pack_half2(vals[14], vals[15]);
          l0.x = pack_half2(lo[0], lo[1]); l0.y = pack_half2(lo[2], lo[3]); l0.z = pack_half2(lo[4], lo[5]); l0.w = pack_half2(lo[6], lo[7]);
          l1.x = pack_half2(lo[8], lo[9]); l1.y = pack_half2(lo[10], lo[11]); l1.z = pack_half2(lo[12], lo[13]); l1.w = pack_half2(lo[14], lo[15]);
          const int pn = col0 >> 6, ch = (col0 & 63) >> 3;
          uint8_t* hrow = act1 + pn * XR_PANEL + r * 128;
          uint8_t* lrow = act1 + (2 + pn) * XR_PANEL + r * 128;
          *reinterpret_cast<uint4*>(hrow + ((ch ^ (r & 7)) << 4)) = h0;
          *reinterpret_cast<uint4*>(hrow + (((ch + 1) ^ (r & 7)) << 4)) = h1;
          *reinterpret_cast<uint4*>(lrow + ((ch ^ (r & 7)) << 4)) = l0;
          *reinterpret_cast<uint4*>(lrow + (((ch + 1) ^ (r & 7)) << 4)) = l1;
        };
        put(zp, p.zp_off + grp * 16);
        put(zc, cond_off + grp * 16);
        fence_proxy_async_smem();
        tc_fence_before();
        warp_arrive(&t_ready[0]);
        // z <- z M + c: 32 of the 128 columns per thread
        mbar_wait(&t_full[1], 0);
        tc_fence_after();
        bar_all();   // explicit edge: every thread's (hi | lo) panel writes precede the re-use of ACT1 panel 0 below
        tmem_ld32(SCR + 128 + grp * 32, v);
        tmem_wait_ld();
        float z2[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) z2[e] = __uint_as_float(v[e]) + __ldg(p.fl_c + grp * 32 + e);
        if (row_ok) {
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4)
            *reinterpret_cast<float4*>(p.z + grow * 128 + grp * 32 + e4 * 4) =
                make_float4(z2[e4 * 4 + 0], z2[e4 * 4 + 1], z2[e4 * 4 + 2], z2[e4 * 4 + 3]);
        }
        if (p.tail_pre) {
          // conditioning half of the next step -> fp16 A operand (ACT1 panel 0; the flow MMAs that read ACT1 are complete)
          const int c0 = grp * 32 - p.cond_off_next;
          if (c0 >= 0 && c0 < 64) {
            uint8_t* crow = act1 + r * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_half2(z2[g * 8 + 0], z2[g * 8 + 1]); u.y = pack_half2(z2[g * 8 + 2], z2[g * 8 + 3]);
              u.z = pack_half2(z2[g * 8 + 4], z2[g * 8 + 5]); u.w = pack_half2(z2[g * 8 + 6], z2[g * 8 + 7]);
              *reinterpret_cast<uint4*>(crow + ((((c0 >> 3) + g) ^ (r & 7)) << 4)) = u;
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          warp_arrive(&t_ready[1]);
          // x <- cond W_pre + b + pos_weight * PE[t]  (modules/transform.py:47-52): the input of the next coupling net
          mbar_wait(&t_full[2], 0);
          tc_fence_after();
          float xs[64];
          uint32_t* xu = reinterpret_cast<uint32_t*>(xs);
          tmem_ld32(R + grp * 64, xu);
          tmem_ld32(R + grp * 64 + 32, xu + 32);
          tmem_wait_ld();
          const float pw = __ldg(p.pre_pw);
          const float* perow = p.pe + static_cast<long>(min(t, p.T - 1)) * XR_D + grp * 64;
#pragma unroll
          for (int g = 0; g < 16; ++g) {
            const float4 bq = __ldg(reinterpret_cast<const float4*>(p.pre_bias + grp * 64 + g * 4));
            const float4 pq = __ldg(reinterpret_cast<const float4*>(perow + g * 4));
            xs[g * 4 + 0] += bq.x + pw * pq.x; xs[g * 4 + 1] += bq.y + pw * pq.y;
            xs[g * 4 + 2] += bq.z + pw * pq.z; xs[g * 4 + 3] += bq.w + pw * pq.w;
          }
          uint8_t* prow = act0 + grp * XR_PANEL + r * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint4 u;
            u.x = pack_half2(xs[g * 8 + 0], xs[g * 8 + 1]); u.y = pack_half2(xs[g * 8 + 2], xs[g * 8 + 3]);
            u.z = pack_half2(xs[g * 8 + 4], xs[g * 8 + 5]); u.w = pack_half2(xs[g * 8 + 6], xs[g * 8 + 7]);
            *reinterpret_cast<uint4*>(prow + ((g ^ (r & 7)) << 4)) = u;
          }
          fence_proxy_async_smem();
          tc_fence_before();
          warp_arrive(act_ready);
          store_x_global(xu);
        }
      }
    }

    // ---- 7. next block's q | k (row-major) and v (transposed)
    if (p.has_next) {
      for (int j = 0; j < 6; ++j) {
        const int bi = j & 1;
        mbar_wait(&h_full[bi], (j >> 1) & 1);
        tc_fence_after();
        tmem_ld32(SCR + bi * 128 + grp * 32, v);
        tmem_wait_ld();
        tc_fence_before();
        warp_arrive(&hid_ready[bi]);   // the chunk now lives in registers
        if (j < 4) {
          // 32 rows x 64 B slab, 16-byte chunk ^= (row >> 1) & 3, then 8 rows x 64 B per store instruction
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_half2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            u.y = pack_half2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            u.z = pack_half2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            u.w = pack_half2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            *reinterpret_cast<uint4*>(slab + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = u;
          }
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + (lane >> 2), chunk = lane & 3;
            if (row < rows_here)
              *reinterpret_cast<uint4*>(p.qk_next + (grow0 + row) * p.qk_next_ld + j * 128 + grp * 32 + chunk * 8) =
                  *reinterpret_cast<const uint4*>(slab + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
          }
        } else if (row_ok) {
          // V^T [b, head, dim, t]: thread == row == consecutive t -> coalesced along t
          const int n0 = (j - 4) * 128 + grp * 32;
          __half* dst = p.vt_next + (static_cast<long>(b * XR_H + (n0 >> 6)) * 64 + (n0 & 63)) * p.vt_next_ld + t;
#pragma unroll
          for (int e = 0; e < 32; ++e) dst[static_cast<long>(e) * p.vt_next_ld] = __float2half_rn(__uint_as_float(v[e]));
          if (p.v_rm_next) {   // training: the backward pass reads V row-major
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_half2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
              u.y = pack_half2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
              u.z = pack_half2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
              u.w = pack_half2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
              *reinterpret_cast<uint4*>(p.v_rm_next + grow * p.qk_next_ld + n0 + g * 8) = u;
            }
          }
        }
        estamp(42 + j);
      }
    }
    if (st) tma_store_wait<0>();
    estamp(48);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

// Backward of MultiHeadScaledProductAttention (modules/attention.py:217-246) for sm_100a, head_dim 64, on
// tcgen05 tensor cores.  Given dO (gradient of the merged context), the saved Q, K, V, the per-row softmax
// statistic  lse2[q] = log2 sum_k 2^(s_qk * scale * log2 e)  written by the forward kernel and
// delta[q] = sum_d dO[q,d] O[q,d], the probabilities are recomputed tile by tile (never stored in HBM):
//
//   P = 2^(s * scale*log2e - lse2)        (masked keys: exactly 0;  fully masked query rows: uniform 1/Tk)
//   dV = P^T dO          dP = dO V^T          dS = P o (dP - delta)      (fully masked rows: dS = 0, the
//   dQ = scale * dS K    dK = scale * dS^T Q                              constant fill has no gradient)
//
// Two kernels, both deterministic (no atomics):
//   attn_bwd_dkdv_kernel : one CTA per (128-key block, head, batch), loops over query blocks.  Works in the
//       TRANSPOSED orientation S^T = K Q^T, dP^T = V dO^T (TMEM lane == key row) so that P^T / dS^T land in shared
//       memory directly as the K-major A operands of dV += P^T dO and dK += dS^T Q; the B operands of those two
//       products are the SAME Q / dO tiles, re-described as MN-major (no transposed copies).
//   attn_bwd_dq_kernel   : one CTA per (128-query block, head, batch), loops over key blocks: S = Q K^T,
//       dP = dO V^T, dS -> shared memory, dQ += dS K with K re-described as MN-major.
// Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 softmax/epilogue (four per TMEM
// lane quadrant, each owning 32 of the 128 columns of a tile).
#pragma once
#include "ptx.cuh"
#include "wgrad_tc.cuh"   // umma_desc_mn_sw128 / umma_idesc_f16_major

namespace vb {

constexpr int ATB_THREADS = 64 + 512;
constexpr int ATB_TILE = 128 * 64 * 2;                 // one 128-row x 64-col fp16 tile: 16 KB
constexpr int ATB_PTILE = 128 * 128 * 2;               // P / dS tile: two 64-column panels of 16 KB
constexpr int ATB_DKDV_SMEM = 6 * ATB_TILE + 2 * ATB_PTILE + 512 + 1024;   // K, V, Q[2], dO[2], P^T, dS^T
constexpr int ATB_DQ_SMEM = 6 * ATB_TILE + ATB_PTILE + 512 + 1024;         // Q, dO, K[2], V[2], dS

struct AttnBwdParams {
  int B, H, Tq, Tk;
  int q_col0, k_col0, v_col0, do_col0;   // column of head 0 inside the Q / K / V / dO tensor maps
  const int* q_len;                      // [B]
  const int* k_len;                      // [B]
  int causal;
  float scale;                           // 1 / sqrt(head_dim)
  const float* lse2;                     // [B, H, Tq]
  const float* delta;                    // [B, H, Tq]
  __half* dq; int dq_ld, dq_col0;        // [B*Tq, dq_ld], head h at columns dq_col0 + h*64
  __half* dk; int dk_ld, dk_col0;        // [B*Tk, dk_ld]
  __half* dv; int dv_ld, dv_col0;        // [B*Tk, dv_ld]
};

// delta[b,h,q] = sum_d dO[b,q,h*64+d] * O[b,q,h*64+d]; one warp per token row, 8 lanes per head.
__global__ void attn_delta_kernel(const __half* __restrict__ dO, int do_ld, int do_col0, const __half* __restrict__ O,
                                  int o_ld, float* __restrict__ delta, int B, int T, int H) {
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long>(B) * T) return;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  for (int c0 = 0; c0 < H * 64; c0 += 256) {
    const int c = c0 + lane * 8;
    float acc = 0.f;
    if (c < H * 64) {
      const uint4 a = *reinterpret_cast<const uint4*>(dO + row * do_ld + do_col0 + c);
      const uint4 o = *reinterpret_cast<const uint4*>(O + row * o_ld + c);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* oh = reinterpret_cast<const __half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __half22float2(ah[e]), y = __half22float2(oh[e]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((lane & 7) == 0 && c < H * 64) delta[(static_cast<long>(b) * H + (c >> 6)) * T + t] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = sK + ATB_TILE;
  uint8_t* sQ = sV + ATB_TILE;            // [2]
  uint8_t* sdO = sQ + 2 * ATB_TILE;       // [2]
  uint8_t* sP = sdO + 2 * ATB_TILE;       // P^T  [128 keys][128 queries] (two 64-query panels)
  uint8_t* sdS = sP + ATB_PTILE;          // dS^T
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ATB_PTILE);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* pds_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  const int n_qblk = (p.Tq + 127) / 128;
  // query blocks that can contribute to this key block (identical in every warp role):
  //   live rows (q < qlen) see the block when it holds a valid key and, if causal, some q >= k0;
  //   fully masked rows (qlen <= q < Tq) attend uniformly to ALL Tk keys -> always contribute to dV.
  auto needed = [&](int i) -> bool {
    const int q_lo = i * 128, q_hi = min(q_lo + 128, p.Tq);
    const bool live = (q_lo < qlen) && (k0 < klen) && (!p.causal || q_hi - 1 >= k0) && (min(q_hi, qlen) - 1 >= (p.causal ? k0 : 0));
    const bool dead = max(q_lo, qlen) < q_hi;
    return live || dead;
  };

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 512);
    mbar_init(pds_full, 512);
    mbar_init(pds_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;           // S^T   [128 keys x 128 queries]
  const uint32_t tmem_dP = tmem_base + 128;    // dP^T
  const uint32_t tmem_dV = tmem_base + 256;    // [128 keys x 64]
  const uint32_t tmem_dK = tmem_base + 320;

  int n_it = 0;
  for (int i = 0; i < n_qblk; ++i) n_it += needed(i) ? 1 : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (n_it > 0 && elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * ATB_TILE);
      tma_load_3d(sK, &tmK, kv_full, p.k_col0 + h * 64, k0, b);
      tma_load_3d(sV, &tmV, kv_full, p.v_col0 + h * 64, k0, b);
      int it = 0;
      for (int i = 0; i < n_qblk; ++i) {
        if (!needed(i)) continue;
        const int st = it & 1;
        mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * ATB_TILE);
        tma_load_3d(sQ + st * ATB_TILE, &tmQ, &qdo_full[st], p.q_col0 + h * 64, i * 128, b);
        tma_load_3d(sdO + st * ATB_TILE, &tmdO, &qdo_full[st], p.do_col0 + h * 64, i * 128, b);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (n_it > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);                         // K-major x K-major, N = 128
      constexpr uint32_t idesc_g = umma_idesc_f16_major(128, 64, false, true);       // A K-major, B MN-major, N = 64
      mbar_wait(kv_full, 0);
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sV));
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait(&qdo_full[st], (it >> 1) & 1);
        mbar_wait(s_empty, (it & 1) ^ 1);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + st * ATB_TILE));
        const uint64_t odesc = umma_desc_sw128(smem_u32(sdO + st * ATB_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S, kdesc + 2 * k, qdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_dP, vdesc + 2 * k, odesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // contraction over the 128 queries of the block, 16 per instruction
          const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + (k >> 2) * (ATB_PTILE / 2))) + 2 * (k & 3);
          const uint64_t bdo = umma_desc_mn_sw128(smem_u32(sdO + st * ATB_TILE) + k * 2048, ATB_TILE, 1024);
          umma_f16(tmem_dV, pdesc, bdo, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t sdesc = umma_desc_sw128(smem_u32(sdS + (k >> 2) * (ATB_PTILE / 2))) + 2 * (k & 3);
          const uint64_t bq = umma_desc_mn_sw128(smem_u32(sQ + st * ATB_TILE) + k * 2048, ATB_TILE, 1024);
          umma_f16(tmem_dK, sdesc, bq, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(pds_empty);
        umma_commit(&qdo_empty[st]);
      }
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax / epilogue warps: thread == key row =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;             // 32-query column group of every 128-query block
    const int r = quad * 32 + lane;
    const int kk = k0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float inv_tk = 1.0f / static_cast<float>(p.Tk);
    const bool key_ok = kk < klen;
    const float* lse_bh = p.lse2 + (static_cast<long>(b) * p.H + h) * p.Tq;
    const float* del_bh = p.delta + (static_cast<long>(b) * p.H + h) * p.Tq;
    uint32_t vs[32], vp[32];
    int it = 0;
    for (int i = 0; i < n_qblk; ++i) {
      if (!needed(i)) continue;
      const int qb = i * 128 + grp * 32;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      __syncwarp();
      tmem_ld32(tmem_S + lane_off + grp * 32, vs);
      tmem_ld32(tmem_dP + lane_off + grp * 32, vp);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(s_empty);
      float pr[32], ds[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int q = qb + e;
        float pe = 0.f, de = 0.f;
        if (q < p.Tq) {
          if (q >= qlen) {
            pe = inv_tk;                                    // fully masked query row: uniform, no logit gradient
          } else if (key_ok && (!p.causal || kk <= q)) {
            pe = ex2_approx(__uint_as_float(vs[e]) * sl2 - __ldg(lse_bh + q));
            de = pe * (__uint_as_float(vp[e]) - __ldg(del_bh + q));
          }
        }
        pr[e] = pe;
        ds[e] = de;
      }
      mbar_wait(pds_empty, (it & 1) ^ 1);
      uint8_t* prow = sP + (grp >> 1) * (ATB_PTILE / 2) + r * 128;
      uint8_t* drow = sdS + (grp >> 1) * (ATB_PTILE / 2) + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (grp & 1) * 4 + g;
        uint4 u, w;
        u.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]); u.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
        u.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]); u.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
        w.x = pack_half2(ds[g * 8 + 0], ds[g * 8 + 1]); w.y = pack_half2(ds[g * 8 + 2], ds[g * 8 + 3]);
        w.z = pack_half2(ds[g * 8 + 4], ds[g * 8 + 5]); w.w = pack_half2(ds[g * 8 + 6], ds[g * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = u;
        *reinterpret_cast<uint4*>(drow + ((chunk ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      ++it;
    }
    // ---- epilogue: groups 0,1 store dV, groups 2,3 store dK * scale (32 head channels each)
    const bool is_dk = grp >= 2;
    const int ch0 = (grp & 1) * 32;
    if (n_it > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      __syncwarp();
      tmem_ld32((is_dk ? tmem_dK : tmem_dV) + lane_off + ch0, vs);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) vs[e] = 0u;
    }
    if (kk < p.Tk) {
      const float mul = is_dk ? p.scale : 1.0f;
      __half* dst = is_dk ? p.dk + (static_cast<long>(b) * p.Tk + kk) * p.dk_ld + p.dk_col0 + h * 64 + ch0
                          : p.dv + (static_cast<long>(b) * p.Tk + kk) * p.dv_ld + p.dv_col0 + h * 64 + ch0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(vs[j + 0]) * mul, __uint_as_float(vs[j + 1]) * mul);
        u.y = pack_half2(__uint_as_float(vs[j + 2]) * mul, __uint_as_float(vs[j + 3]) * mul);
        u.z = pack_half2(__uint_as_float(vs[j + 4]) * mul, __uint_as_float(vs[j + 5]) * mul);
        u.w = pack_half2(__uint_as_float(vs[j + 6]) * mul, __uint_as_float(vs[j + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + ATB_TILE;
  uint8_t* sK = sdO + ATB_TILE;           // [2]
  uint8_t* sV = sK + 2 * ATB_TILE;        // [2]
  uint8_t* sdS = sV + 2 * ATB_TILE;       // dS [128 queries][128 keys] (two 64-key panels)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ATB_PTILE);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* ds_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  const int q_hi = min(q0 + 128, p.Tq);
  int nblk = 0;   // key blocks that hold an unmasked key for some live row of this tile
  if (q0 < qlen && klen > 0) {
    nblk = min((p.Tk + 127) / 128, (klen + 127) / 128);
    if (p.causal) nblk = min(nblk, (min(q_hi, qlen) - 1) / 128 + 1);
  }

  if (threadIdx.x == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 512);
    mbar_init(ds_full, 512);
    mbar_init(ds_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_dP = tmem_base + 128;
  const uint32_t tmem_dQ = tmem_base + 256;

  if (warp == 0) {
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(qdo_full, 2 * ATB_TILE);
      tma_load_3d(sQ, &tmQ, qdo_full, p.q_col0 + h * 64, q0, b);
      tma_load_3d(sdO, &tmdO, qdo_full, p.do_col0 + h * 64, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * ATB_TILE);
        tma_load_3d(sK + st * ATB_TILE, &tmK, &kv_full[st], p.k_col0 + h * 64, j * 128, b);
        tma_load_3d(sV + st * ATB_TILE, &tmV, &kv_full[st], p.v_col0 + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (nblk > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128);
      constexpr uint32_t idesc_g = umma_idesc_f16_major(128, 64, false, true);
      mbar_wait(qdo_full, 0);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t odesc = umma_desc_sw128(smem_u32(sdO));
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(s_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * ATB_TILE));
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + st * ATB_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_dP, odesc + 2 * k, vdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // contraction over the 128 keys of the block
          const uint64_t sdesc = umma_desc_sw128(smem_u32(sdS + (k >> 2) * (ATB_PTILE / 2))) + 2 * (k & 3);
          const uint64_t bk = umma_desc_mn_sw128(smem_u32(sK + st * ATB_TILE) + k * 2048, ATB_TILE, 1024);
          umma_f16(tmem_dQ, sdesc, bk, idesc_g, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(ds_empty);
        umma_commit(&kv_empty[st]);
      }
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax / epilogue warps: thread == query row =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool row_live = q < qlen && q < p.Tq;
    const long sidx = (static_cast<long>(b) * p.H + h) * p.Tq + min(q, p.Tq - 1);
    const float lse = row_live ? __ldg(p.lse2 + sidx) : 0.f;
    const float del = row_live ? __ldg(p.delta + sidx) : 0.f;
    uint32_t vs[32], vp[32];
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      __syncwarp();
      tmem_ld32(tmem_S + lane_off + grp * 32, vs);
      tmem_ld32(tmem_dP + lane_off + grp * 32, vp);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(s_empty);
      const int kb = j * 128 + grp * 32;
      float ds[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = kb + e;
        float de = 0.f;
        if (row_live && kk < klen && (!p.causal || kk <= q)) {
          const float pe = ex2_approx(__uint_as_float(vs[e]) * sl2 - lse);
          de = pe * (__uint_as_float(vp[e]) - del);
        }
        ds[e] = de;
      }
      mbar_wait(ds_empty, (j & 1) ^ 1);
      uint8_t* drow = sdS + (grp >> 1) * (ATB_PTILE / 2) + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (grp & 1) * 4 + g;
        uint4 w;
        w.x = pack_half2(ds[g * 8 + 0], ds[g * 8 + 1]); w.y = pack_half2(ds[g * 8 + 2], ds[g * 8 + 3]);
        w.z = pack_half2(ds[g * 8 + 4], ds[g * 8 + 5]); w.w = pack_half2(ds[g * 8 + 6], ds[g * 8 + 7]);
        *reinterpret_cast<uint4*>(drow + ((chunk ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    if (grp < 2) {
      if (nblk > 0) {
        mbar_wait(o_full, 0);
        tc_fence_after();
        __syncwarp();
        tmem_ld32(tmem_dQ + lane_off + grp * 32, vs);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) vs[e] = 0u;
      }
      if (q < p.Tq) {
        __half* dst = p.dq + (static_cast<long>(b) * p.Tq + q) * p.dq_ld + p.dq_col0 + h * 64 + grp * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(vs[j + 0]) * p.scale, __uint_as_float(vs[j + 1]) * p.scale);
          u.y = pack_half2(__uint_as_float(vs[j + 2]) * p.scale, __uint_as_float(vs[j + 3]) * p.scale);
          u.z = pack_half2(__uint_as_float(vs[j + 4]) * p.scale, __uint_as_float(vs[j + 5]) * p.scale);
          u.w = pack_half2(__uint_as_float(vs[j + 6]) * p.scale, __uint_as_float(vs[j + 7]) * p.scale);
          *reinterpret_cast<uint4*>(dst + j) = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ================================================================================================ v2: 64-wide sub-tiles
// Same mathematics and operand tricks as above, but the loop dimension is tiled by 64 (queries for dK/dV, keys for dQ):
// 96 / 80 KB of shared memory and 256 TMEM columns per CTA, so TWO CTAs are resident per SM and the softmax phase of
// one overlaps the tensor-core / TMA phases of the other.  320 threads: warp 0 TMA, warp 1 MMA, warps 2..9 softmax
// (two per TMEM lane quadrant, 32 columns each, processed as 2 x 16 to stay within 100 registers).
constexpr int ATB2_THREADS = 64 + 256;
constexpr int ATB2_T128 = 128 * 64 * 2;   // 128-row x 64-col fp16 tile: 16 KB
constexpr int ATB2_T64 = 64 * 64 * 2;     // 64-row tile: 8 KB
constexpr int ATB2_DKDV_SMEM = 2 * ATB2_T128 + 4 * ATB2_T64 + 2 * ATB2_T128 + 512 + 1024;   // K, V, Q[2], dO[2], P^T, dS^T
constexpr int ATB2_DQ_SMEM = 2 * ATB2_T128 + 4 * ATB2_T64 + ATB2_T128 + 512 + 1024;         // Q, dO, K[2], V[2], dS

__global__ void __launch_bounds__(ATB2_THREADS, 2)
attn_bwd_dkdv2_kernel(const __grid_constant__ CUtensorMap tmQ /*box 64 rows*/, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO /*box 64 rows*/,
                      const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = sK + ATB2_T128;
  uint8_t* sQ = sV + ATB2_T128;           // [2]
  uint8_t* sdO = sQ + 2 * ATB2_T64;       // [2]
  uint8_t* sP = sdO + 2 * ATB2_T64;       // P^T  [128 keys][64 queries]
  uint8_t* sdS = sP + ATB2_T128;          // dS^T
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ATB2_T128);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* pds_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  const int n_qblk = (p.Tq + 63) / 64;
  auto needed = [&](int i) -> bool {
    const int q_lo = i * 64, q_hi = min(q_lo + 64, p.Tq);
    const bool live = (q_lo < qlen) && (k0 < klen) && (!p.causal || min(q_hi, qlen) - 1 >= k0);
    const bool dead = max(q_lo, qlen) < q_hi;
    return live || dead;
  };

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 256);
    mbar_init(pds_full, 256);
    mbar_init(pds_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;           // S^T   [128 keys x 64 queries]
  const uint32_t tmem_dP = tmem_base + 64;     // dP^T
  const uint32_t tmem_dV = tmem_base + 128;    // [128 keys x 64]
  const uint32_t tmem_dK = tmem_base + 192;

  int n_it = 0;
  for (int i = 0; i < n_qblk; ++i) n_it += needed(i) ? 1 : 0;

  if (warp == 0) {
    if (n_it > 0 && elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * ATB2_T128);
      tma_load_3d(sK, &tmK, kv_full, p.k_col0 + h * 64, k0, b);
      tma_load_3d(sV, &tmV, kv_full, p.v_col0 + h * 64, k0, b);
      int it = 0;
      for (int i = 0; i < n_qblk; ++i) {
        if (!needed(i)) continue;
        const int st = it & 1;
        mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * ATB2_T64);
        tma_load_3d(sQ + st * ATB2_T64, &tmQ, &qdo_full[st], p.q_col0 + h * 64, i * 64, b);
        tma_load_3d(sdO + st * ATB2_T64, &tmdO, &qdo_full[st], p.do_col0 + h * 64, i * 64, b);
        ++it;
      }
    }
  } else if (warp == 1) {
    if (n_it > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 64);
      constexpr uint32_t idesc_g = umma_idesc_f16_major(128, 64, false, true);
      mbar_wait(kv_full, 0);
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK));
      const uint64_t vdesc = umma_desc_sw128(smem_u32(sV));
      const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
      const uint64_t sdesc = umma_desc_sw128(smem_u32(sdS));
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        mbar_wait(&qdo_full[st], (it >> 1) & 1);
        mbar_wait(s_empty, (it & 1) ^ 1);
        tc_fence_after();
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + st * ATB2_T64));
        const uint64_t odesc = umma_desc_sw128(smem_u32(sdO + st * ATB2_T64));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S, kdesc + 2 * k, qdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_dP, vdesc + 2 * k, odesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // contraction over the 64 queries of the sub-tile, 16 per instruction
          const uint64_t bdo = umma_desc_mn_sw128(smem_u32(sdO + st * ATB2_T64) + k * 2048, ATB2_T64, 1024);
          umma_f16(tmem_dV, pdesc + 2 * k, bdo, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bq = umma_desc_mn_sw128(smem_u32(sQ + st * ATB2_T64) + k * 2048, ATB2_T64, 1024);
          umma_f16(tmem_dK, sdesc + 2 * k, bq, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(pds_empty);
        umma_commit(&qdo_empty[st]);
      }
      umma_commit(o_full);
    }
  } else {
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;             // 0/1: which 32 queries of the 64-query sub-tile
    const int r = quad * 32 + lane;
    const int kk = k0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float inv_tk = 1.0f / static_cast<float>(p.Tk);
    const bool key_ok = kk < klen;
    const float* lse_bh = p.lse2 + (static_cast<long>(b) * p.H + h) * p.Tq;
    const float* del_bh = p.delta + (static_cast<long>(b) * p.H + h) * p.Tq;
    uint8_t* prow = sP + r * 128;
    uint8_t* drow = sdS + r * 128;
    uint32_t vs[16], vp[16];
    int it = 0;
    for (int i = 0; i < n_qblk; ++i) {
      if (!needed(i)) continue;
      mbar_wait(s_full, it & 1);
      mbar_wait(pds_empty, (it & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int qb = i * 64 + grp * 32 + hh * 16;
        __syncwarp();
        tmem_ld16(tmem_S + lane_off + grp * 32 + hh * 16, vs);
        tmem_ld16(tmem_dP + lane_off + grp * 32 + hh * 16, vp);
        tmem_wait_ld();
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float pe[2], de[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int q = qb + e + u;
            pe[u] = 0.f; de[u] = 0.f;
            if (q < p.Tq) {
              if (q >= qlen) {
                pe[u] = inv_tk;
              } else if (key_ok && (!p.causal || kk <= q)) {
                pe[u] = ex2_approx(__uint_as_float(vs[e + u]) * sl2 - __ldg(lse_bh + q));
                de[u] = pe[u] * (__uint_as_float(vp[e + u]) - __ldg(del_bh + q));
              }
            }
          }
          pk[e >> 1] = pack_half2(pe[0], pe[1]);
          dk[e >> 1] = pack_half2(de[0], de[1]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int chunk = grp * 4 + hh * 2 + g;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
          *reinterpret_cast<uint4*>(drow + ((chunk ^ (r & 7)) << 4)) = make_uint4(dk[g * 4], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(pds_full);
      ++it;
    }
    // ---- epilogue: group 0 stores dV, group 1 stores dK * scale (64 head channels, two 32-column reads)
    const bool is_dk = grp == 1;
    uint32_t v32[32];
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      if (n_it > 0) {
        if (c2 == 0) {
          mbar_wait(o_full, 0);
          tc_fence_after();
        }
        __syncwarp();
        tmem_ld32((is_dk ? tmem_dK : tmem_dV) + lane_off + c2 * 32, v32);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v32[e] = 0u;
      }
      if (kk < p.Tk) {
        const float mul = is_dk ? p.scale : 1.0f;
        __half* dst = is_dk ? p.dk + (static_cast<long>(b) * p.Tk + kk) * p.dk_ld + p.dk_col0 + h * 64 + c2 * 32
                            : p.dv + (static_cast<long>(b) * p.Tk + kk) * p.dv_ld + p.dv_col0 + h * 64 + c2 * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(v32[j + 0]) * mul, __uint_as_float(v32[j + 1]) * mul);
          u.y = pack_half2(__uint_as_float(v32[j + 2]) * mul, __uint_as_float(v32[j + 3]) * mul);
          u.z = pack_half2(__uint_as_float(v32[j + 4]) * mul, __uint_as_float(v32[j + 5]) * mul);
          u.w = pack_half2(__uint_as_float(v32[j + 6]) * mul, __uint_as_float(v32[j + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + j) = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

__global__ void __launch_bounds__(ATB2_THREADS, 2)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK /*box 64 rows*/,
                    const __grid_constant__ CUtensorMap tmV /*box 64 rows*/, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + ATB2_T128;
  uint8_t* sK = sdO + ATB2_T128;          // [2] 64-key tiles
  uint8_t* sV = sK + 2 * ATB2_T64;        // [2]
  uint8_t* sdS = sV + 2 * ATB2_T64;       // dS [128 queries][64 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + ATB2_T128);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* ds_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  const int q_hi = min(q0 + 128, p.Tq);
  int nblk = 0;   // 64-key blocks that hold an unmasked key for some live row of this tile
  if (q0 < qlen && klen > 0) {
    nblk = min((p.Tk + 63) / 64, (klen + 63) / 64);
    if (p.causal) nblk = min(nblk, (min(q_hi, qlen) - 1) / 64 + 1);
  }

  if (threadIdx.x == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 256);
    mbar_init(ds_full, 256);
    mbar_init(ds_empty, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_dP = tmem_base + 64;
  const uint32_t tmem_dQ = tmem_base + 128;

  if (warp == 0) {
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(qdo_full, 2 * ATB2_T128);
      tma_load_3d(sQ, &tmQ, qdo_full, p.q_col0 + h * 64, q0, b);
      tma_load_3d(sdO, &tmdO, qdo_full, p.do_col0 + h * 64, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * ATB2_T64);
        tma_load_3d(sK + st * ATB2_T64, &tmK, &kv_full[st], p.k_col0 + h * 64, j * 64, b);
        tma_load_3d(sV + st * ATB2_T64, &tmV, &kv_full[st], p.v_col0 + h * 64, j * 64, b);
      }
    }
  } else if (warp == 1) {
    if (nblk > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 64);
      constexpr uint32_t idesc_g = umma_idesc_f16_major(128, 64, false, true);
      mbar_wait(qdo_full, 0);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t odesc = umma_desc_sw128(smem_u32(sdO));
      const uint64_t sdesc = umma_desc_sw128(smem_u32(sdS));
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(s_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * ATB2_T64));
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + st * ATB2_T64));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_dP, odesc + 2 * k, vdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // contraction over the 64 keys of the block
          const uint64_t bk = umma_desc_mn_sw128(smem_u32(sK + st * ATB2_T64) + k * 2048, ATB2_T64, 1024);
          umma_f16(tmem_dQ, sdesc + 2 * k, bk, idesc_g, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(ds_empty);
        umma_commit(&kv_empty[st]);
      }
      umma_commit(o_full);
    }
  } else {
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool row_live = q < qlen && q < p.Tq;
    const long sidx = (static_cast<long>(b) * p.H + h) * p.Tq + min(q, p.Tq - 1);
    const float lse = row_live ? __ldg(p.lse2 + sidx) : 0.f;
    const float del = row_live ? __ldg(p.delta + sidx) : 0.f;
    uint8_t* drow = sdS + r * 128;
    uint32_t vs[16], vp[16];
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      mbar_wait(ds_empty, (j & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int kb = j * 64 + grp * 32 + hh * 16;
        __syncwarp();
        tmem_ld16(tmem_S + lane_off + grp * 32 + hh * 16, vs);
        tmem_ld16(tmem_dP + lane_off + grp * 32 + hh * 16, vp);
        tmem_wait_ld();
        uint32_t dk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float de[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int kk = kb + e + u;
            de[u] = 0.f;
            if (row_live && kk < klen && (!p.causal || kk <= q)) {
              const float pe = ex2_approx(__uint_as_float(vs[e + u]) * sl2 - lse);
              de[u] = pe * (__uint_as_float(vp[e + u]) - del);
            }
          }
          dk[e >> 1] = pack_half2(de[0], de[1]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int chunk = grp * 4 + hh * 2 + g;
          *reinterpret_cast<uint4*>(drow + ((chunk ^ (r & 7)) << 4)) = make_uint4(dk[g * 4], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();
      mbar_arrive(ds_full);
    }
    uint32_t v32[32];
    if (nblk > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      __syncwarp();
      tmem_ld32(tmem_dQ + lane_off + grp * 32, v32);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) v32[e] = 0u;
    }
    if (q < p.Tq) {
      __half* dst = p.dq + (static_cast<long>(b) * p.Tq + q) * p.dq_ld + p.dq_col0 + h * 64 + grp * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(v32[j + 0]) * p.scale, __uint_as_float(v32[j + 1]) * p.scale);
        u.y = pack_half2(__uint_as_float(v32[j + 2]) * p.scale, __uint_as_float(v32[j + 3]) * p.scale);
        u.z = pack_half2(__uint_as_float(v32[j + 4]) * p.scale, __uint_as_float(v32[j + 5]) * p.scale);
        u.w = pack_half2(__uint_as_float(v32[j + 6]) * p.scale, __uint_as_float(v32[j + 7]) * p.scale);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace vb

// Multi-head scaled-dot-product attention for sm_100a (tcgen05 + TMEM + TMA), one CTA per
// (query tile of 128, head, batch).  Restates MultiHeadScaledProductAttention.call
// (modules/attention.py:217-246) with head_dim = 64:
//
//   logits = Q K^T / sqrt(64) ; mask = key_len AND query_len (AND lower-triangular if causal)
//   where(mask, logits, -2^32+1) ; softmax over keys ; ctx = P V ; alignments = P (optional output)
//
// A fully masked query row (q >= query_len) yields the UNIFORM distribution 1/Tk over all Tk padded
// keys, exactly as the reference's constant fill does (attention.py:240-242); masked keys of a live row
// get exactly 0.  The mask is an in-register predicate from the length arrays -- no mask tensor.
//
// Two passes over the key blocks (128 keys each), both on tensor cores:
//   pass 1: S = Q K^T -> TMEM, softmax warps reduce the row maximum (and the denominator when the
//           alignments are requested);
//   pass 2: S again, p = exp(s - max) written as fp16 into 128B-swizzled shared memory (the UMMA
//           A-operand layout), O += P V accumulated in TMEM; ctx = O / l.
// No online rescaling of O is ever needed, the whole row never has to be resident, and Tk is unbounded.
//
// Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 softmax: four warps per TMEM lane
// quadrant, each owning 32 keys of every 128-key block (row statistics combined through shared memory).
// Key blocks that cannot contribute to a LIVE row are skipped (above the causal diagonal, beyond key_len).  Fully
// masked rows attend uniformly to all T_k keys, i.e. their context is the column mean of V: tiles that contain such
// rows compute that mean once from V^T (fp32) and write it directly, so dead rows / dead tiles cost no MMA work.
#pragma once
#include "ptx.cuh"

namespace vb {

constexpr int ATT_GROUPS = 4;                       // column groups (warps per TMEM lane quadrant)
constexpr int ATT_SOFTMAX_THREADS = 128 * ATT_GROUPS;
constexpr int ATT_THREADS = 64 + ATT_SOFTMAX_THREADS;
constexpr int ATT_BQ = 128;     // queries per CTA
constexpr int ATT_BK = 128;     // keys per block
constexpr int ATT_D = 64;       // head dim
constexpr int ATT_QBYTES = ATT_BQ * ATT_D * 2;       // 16 KB
constexpr int ATT_KBYTES = ATT_BK * ATT_D * 2;       // 16 KB
constexpr int ATT_VBYTES = ATT_D * ATT_BK * 2;       // 16 KB (two 64-key panels of 8 KB)
constexpr int ATT_PBYTES = ATT_BQ * ATT_BK * 2;      // 32 KB (two 64-key panels of 16 KB)
constexpr int ATT_SMEM = ATT_QBYTES + 2 * ATT_KBYTES + 2 * ATT_VBYTES + 2 * ATT_PBYTES + 256 + 3 * ATT_GROUPS * ATT_BQ * 4 + 256 + 1024;

struct AttnParams {
  int B, H, Tq, Tk;
  int q_col0;            // column of head 0 inside the Q tensor map
  int k_col0;            // column of head 0 inside the K tensor map
  long vt_row0;          // first V^T row of (batch 0, head 0) for this block
  const __half* vt;      // V^T base pointer and row pitch (for the column mean used by fully masked rows)
  int vt_ld;
  const int* q_len;      // [B]
  const int* k_len;      // [B]
  int causal;
  float scale;           // 1 / sqrt(head_dim) / temperature
  __half* ctx;           // [B*Tq, ctx_ld] fp16, head h at columns h*64
  int ctx_ld;
  float* ali;            // optional [B, H, Tq, Tk] fp32
  float* lse2;           // optional [B, H, Tq] fp32: log2 of the softmax denominator incl. the max (saved for the backward pass)
  unsigned long long* dbg;   // optional per-CTA phase timestamps (tuning aid), 8 x u64 per CTA
};

template <bool kWriteAli>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // stays in the shared address space
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_QBYTES;
  uint8_t* sV = sK + 2 * ATT_KBYTES;
  uint8_t* sP = sV + 2 * ATT_VBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_PBYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* red = reinterpret_cast<float*>(bars + 32);   // [3][2][128] row-statistic exchange between the two column halves

  unsigned long long* dbg = p.dbg ? p.dbg + ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[0] = t; }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  // key blocks that can contribute to the LIVE rows of this query tile (identical in every warp role)
  const int q_hi = min(q0 + ATT_BQ, p.Tq);
  const bool has_dead_rows = max(q0, qlen) < q_hi || klen <= 0;   // some stored row is fully masked
  const bool all_dead = q0 >= qlen || klen <= 0;                   // every stored row is fully masked
  int nblk = 0;
  if (!all_dead) {
    nblk = min((p.Tk + ATT_BK - 1) / ATT_BK, (klen + ATT_BK - 1) / ATT_BK);
    if (p.causal) nblk = min(nblk, (min(q_hi, qlen) - 1) / ATT_BK + 1);
  }
  float* vmean = red + 3 * ATT_GROUPS * ATT_BQ;   // [64] column mean of V for fully masked rows
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], ATT_SOFTMAX_THREADS);
      mbar_init(&p_full[i], ATT_SOFTMAX_THREADS);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // PDL: prologue overlapped the previous kernel's tail
  pdl_wait();
  if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[1] = t; dbg[5] = static_cast<unsigned long long>(nblk); }
  const uint32_t tmem_S = tmem_base;          // two buffers: columns [0,128) and [128,256)
  const uint32_t tmem_O = tmem_base + 256;    // columns [256, 320)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_QBYTES);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + h * ATT_D, q0, b);
      const long vrow = p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D;
      for (int i = 0; i < 2 * nblk; ++i) {
        const int j = i % nblk;
        const int st = i & 1;
        mbar_wait(&k_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], ATT_KBYTES);
        tma_load_3d(sK + st * ATT_KBYTES, &tmK, &k_full[st], p.k_col0 + h * ATT_D, j * ATT_BK, b);
        if (i >= nblk) {
          const int iv = i - nblk;
          const int sv = iv & 1;
          mbar_wait(&v_empty[sv], ((iv >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[sv], ATT_VBYTES);
          tma_load_2d(sV + sv * ATT_VBYTES, &tmVt, &v_full[sv], j * ATT_BK, static_cast<int>(vrow));
          tma_load_2d(sV + sv * ATT_VBYTES + ATT_VBYTES / 2, &tmVt, &v_full[sv], j * ATT_BK + 64,
                      static_cast<int>(vrow));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (nblk > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(ATT_BQ, ATT_BK);   // S: 128 x 128, K = 64
      constexpr uint32_t idesc_o = umma_idesc_f16(ATT_BQ, ATT_D);    // O: 128 x 64,  K = 128
      auto issue_pv = [&](int iv) {
        const int sv = iv & 1;
        const uint32_t par = (iv >> 1) & 1;
        mbar_wait(&p_full[sv], par);
        mbar_wait(&v_full[sv], par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT_BK / 16; ++k) {
          const uint64_t adesc =
              umma_desc_sw128(smem_u32(sP + sv * ATT_PBYTES + (k >> 2) * (ATT_PBYTES / 2))) + 2 * (k & 3);
          const uint64_t bdesc =
              umma_desc_sw128(smem_u32(sV + sv * ATT_VBYTES + (k >> 2) * (ATT_VBYTES / 2))) + 2 * (k & 3);
          umma_f16(tmem_O, adesc, bdesc, idesc_o, (iv > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[sv]);
        umma_commit(&v_empty[sv]);
      };
      mbar_wait(q_full, 0);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      for (int i = 0; i < 2 * nblk; ++i) {
        const int st = i & 1;
        const uint32_t par = (i >> 1) & 1;
        mbar_wait(&k_full[st], par);
        mbar_wait(&s_empty[st], par ^ 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * ATT_KBYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16(tmem_S + st * ATT_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
        if (i > nblk) issue_pv(i - nblk - 1);
      }
      issue_pv(nblk - 1);
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quad = warp & 3;                   // TMEM lane quadrant (hardware rule: warp_id % 4)
    const int grp = (warp - 2) >> 2;             // which 32 keys of each 128-key block this warp owns
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool row_dead = (q >= qlen) || (klen <= 0);      // fully masked row -> uniform over Tk
    const bool row_store = q < p.Tq;
    const float sl2 = p.scale * 1.4426950408889634f;       // scale * log2(e)
    const float inv_tk = 1.0f / static_cast<float>(p.Tk);
    uint32_t v[32];
    auto softmax_bar = []() { asm volatile("bar.sync 1, %0;" ::"n"(ATT_SOFTMAX_THREADS) : "memory"); };
    float* red_m = red;                              // [G][128]
    float* red_l = red + ATT_GROUPS * ATT_BQ;        // [G][128]
    float* red_s = red + 2 * ATT_GROUPS * ATT_BQ;    // [G][128]

    if (has_dead_rows) {
      // column mean of V over ALL Tk padded keys (the uniform distribution of attention.py:240-242), fp32
      const int sidx = (warp - 2) * 32 + lane;      // 0..511: 8 threads per head channel
      const int d = sidx >> 3, part = sidx & 7;
      const __half* vrow = p.vt + (p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D + d) * p.vt_ld;
      float acc = 0.f;
      for (int t0 = part * 8; t0 < p.Tk; t0 += 64) {
        if (t0 + 8 <= p.Tk) {
          const uint4 u = *reinterpret_cast<const uint4*>(vrow + t0);
          const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc += f.x + f.y;
          }
        } else {
          for (int t = t0; t < p.Tk; ++t) acc += __half2float(vrow[t]);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (part == 0) vmean[d] = acc * inv_tk;
    }

    // ---- pass 1: row maximum (and denominator if the alignments are written)
    float m = -INFINITY;
    float l = 0.f;
    for (int i = 0; i < nblk; ++i) {
      const int st = i & 1;
      mbar_wait(&s_full[st], (i >> 1) & 1);
      tc_fence_after();
      __syncwarp();
      tmem_ld32(tmem_S + st * ATT_BK + lane_off + grp * 32, v);
      tmem_wait_ld();
      const int kk0 = i * ATT_BK + grp * 32;
      float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = kk0 + e;
        const bool ok = (kk < klen) && (!p.causal || kk <= q);
        const float sv = ok ? __uint_as_float(v[e]) : -INFINITY;
        bm[e & 3] = fmaxf(bm[e & 3], sv);
        if (kWriteAli) v[e] = __float_as_uint(sv);
      }
      const float cm = fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3]));
      if (kWriteAli) {
        const float mn = fmaxf(m, cm);
        if (mn > -INFINITY) {
          float add[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 32; ++e) add[e & 3] += ex2_approx((__uint_as_float(v[e]) - mn) * sl2);
          l = l * ex2_approx((m - mn) * sl2) + (add[0] + add[1]) + (add[2] + add[3]);
          m = mn;
        }
      } else {
        m = fmaxf(m, cm);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[st]);
    }
    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[2] = t; }
    // combine the column groups of every row
    red_m[grp * ATT_BQ + r] = m;
    if (kWriteAli) red_l[grp * ATT_BQ + r] = l;
    softmax_bar();
    {
      float mn = m;
#pragma unroll
      for (int g = 0; g < ATT_GROUPS; ++g) mn = fmaxf(mn, red_m[g * ATT_BQ + r]);
      if (kWriteAli) {
        float lt = 0.f;
        if (mn > -INFINITY) {
#pragma unroll
          for (int g = 0; g < ATT_GROUPS; ++g) {
            const float mg = red_m[g * ATT_BQ + r];
            if (mg > -INFINITY) lt += red_l[g * ATT_BQ + r] * ex2_approx((mg - mn) * sl2);
          }
        }
        l = lt;
      }
      m = mn;
    }
    if (row_dead) { m = 0.f; l = 1.f; }

    // ---- pass 2: probabilities -> shared memory (A operand of P V), denominators, alignments
    float ls[4] = {0.f, 0.f, 0.f, 0.f};
    const float inv_l = kWriteAli ? 1.0f / l : 1.0f;
    const float msl2 = m * sl2;
    for (int iv = 0; iv < nblk; ++iv) {
      const int i = nblk + iv;
      const int st = i & 1;
      const int sp = iv & 1;
      mbar_wait(&s_full[st], (i >> 1) & 1);
      mbar_wait(&p_empty[sp], ((iv >> 1) & 1) ^ 1);
      tc_fence_after();
      // this warp's 32 keys inside the 64-key panel (grp >> 1): 128 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)
      uint8_t* prow = sP + sp * ATT_PBYTES + (grp >> 1) * (ATT_PBYTES / 2) + r * 128;
      __syncwarp();
      tmem_ld32(tmem_S + st * ATT_BK + lane_off + grp * 32, v);
      tmem_wait_ld();
      const int kk0 = iv * ATT_BK + grp * 32;
      float pr[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = kk0 + e;
        float pe;
        if (row_dead) {
          pe = 0.f;                 // dead rows take the V column mean directly (epilogue)
        } else {
          const bool ok = (kk < klen) && (!p.causal || kk <= q);
          pe = ok ? ex2_approx(__uint_as_float(v[e]) * sl2 - msl2) : 0.f;
        }
        ls[e & 3] += pe;
        pr[e] = pe * inv_l;           // normalised already when kWriteAli (inv_l == 1 otherwise)
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (grp & 1) * 4 + g;
        uint4 u;
        u.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]);
        u.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
        u.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]);
        u.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = u;
      }
      if (kWriteAli && row_store) {
        float* arow = p.ali + ((static_cast<long>(b) * p.H + h) * p.Tq + q) * p.Tk + kk0;
        for (int e = 0; e < 32 && kk0 + e < p.Tk; ++e) arow[e] = row_dead ? inv_tk : pr[e];
      }
      fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&p_full[sp]);
      mbar_arrive(&s_empty[st]);
    }
    // alignments of skipped key blocks: exact zeros for live rows, uniform 1/Tk for fully masked rows
    if (kWriteAli && row_store) {
      const int k_done = nblk * ATT_BK;
      float* arow = p.ali + ((static_cast<long>(b) * p.H + h) * p.Tq + q) * p.Tk;
      const float fill = row_dead ? inv_tk : 0.f;
      for (int kk = k_done + grp; kk < p.Tk; kk += ATT_GROUPS) arow[kk] = fill;
    }

    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[3] = t; }
    // ---- epilogue: ctx = O / l ; column groups 0 and 1 each write 32 of the 64 head channels
    float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
    red_s[grp * ATT_BQ + r] = lsum;
    softmax_bar();
    if (grp < 2) {
      lsum = 0.f;
#pragma unroll
      for (int g = 0; g < ATT_GROUPS; ++g) lsum += red_s[g * ATT_BQ + r];
      if (nblk > 0) {
        mbar_wait(o_full, 0);
        tc_fence_after();
        __syncwarp();
        tmem_ld32(tmem_O + lane_off + grp * 32, v);
        tmem_wait_ld();
      }
      const float on = row_dead ? 0.f : (kWriteAli ? 1.0f : 1.0f / lsum);
      if (p.lse2 && grp == 0 && row_store)
        p.lse2[(static_cast<long>(b) * p.H + h) * p.Tq + q] = row_dead ? 0.f : msl2 + log2f(lsum);
      if (row_store) {
        __half* dst = p.ctx + (static_cast<long>(b) * p.Tq + q) * p.ctx_ld + h * ATT_D + grp * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = row_dead ? vmean[grp * 32 + j + e] : __uint_as_float(v[j + e]) * on;
          uint4 u;
          u.x = pack_half2(f[0], f[1]);
          u.y = pack_half2(f[2], f[3]);
          u.z = pack_half2(f[4], f[5]);
          u.w = pack_half2(f[6], f[7]);
          *reinterpret_cast<uint4*>(dst + j) = u;
        }
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[4] = t; }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ================================================================================================ v2: 64-key blocks
// Same algorithm with 64-key blocks, single-panel P / V^T tiles and 8 softmax warps: 84 KB of shared memory and 256
// TMEM columns per CTA, i.e. TWO CTAs per SM -- the softmax of one overlaps the tensor-core / TMA phases of the other,
// and the two launch chains of the batch-split inference fit on the chip in one wave.
constexpr int ATT2_GROUPS = 2;
constexpr int ATT2_SOFTMAX_THREADS = 128 * ATT2_GROUPS;
constexpr int ATT2_THREADS = 64 + ATT2_SOFTMAX_THREADS;
constexpr int ATT2_BK = 64;
constexpr int ATT2_KBYTES = ATT2_BK * ATT_D * 2;       // 8 KB
constexpr int ATT2_VBYTES = ATT_D * ATT2_BK * 2;       // 8 KB
constexpr int ATT2_PBYTES = ATT_BQ * ATT2_BK * 2;      // 16 KB
constexpr int ATT2_SMEM = ATT_QBYTES + 2 * ATT2_KBYTES + 2 * ATT2_VBYTES + 2 * ATT2_PBYTES + 256 + 3 * ATT2_GROUPS * ATT_BQ * 4 + 256 + 1024;

template <bool kWriteAli>
__global__ void __launch_bounds__(ATT2_THREADS, 2)
attention2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // stays in the shared address space
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_QBYTES;
  uint8_t* sV = sK + 2 * ATT2_KBYTES;
  uint8_t* sP = sV + 2 * ATT2_VBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT2_PBYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* red = reinterpret_cast<float*>(bars + 32);   // [3][2][128] row-statistic exchange between the two column halves

  unsigned long long* dbg = p.dbg ? p.dbg + ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[0] = t; }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int qlen = __ldg(p.q_len + b);
  const int klen = __ldg(p.k_len + b);
  // key blocks that can contribute to the LIVE rows of this query tile (identical in every warp role)
  const int q_hi = min(q0 + ATT_BQ, p.Tq);
  const bool has_dead_rows = max(q0, qlen) < q_hi || klen <= 0;   // some stored row is fully masked
  const bool all_dead = q0 >= qlen || klen <= 0;                   // every stored row is fully masked
  int nblk = 0;
  if (!all_dead) {
    nblk = min((p.Tk + ATT2_BK - 1) / ATT2_BK, (klen + ATT2_BK - 1) / ATT2_BK);
    if (p.causal) nblk = min(nblk, (min(q_hi, qlen) - 1) / ATT2_BK + 1);
  }
  float* vmean = red + 3 * ATT2_GROUPS * ATT_BQ;   // [64] column mean of V for fully masked rows
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], ATT2_SOFTMAX_THREADS);
      mbar_init(&p_full[i], ATT2_SOFTMAX_THREADS);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // PDL: prologue overlapped the previous kernel's tail
  pdl_wait();
  if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[1] = t; dbg[5] = static_cast<unsigned long long>(nblk); }
  const uint32_t tmem_S = tmem_base;          // two buffers: columns [0,64) and [64,128)
  const uint32_t tmem_O = tmem_base + 128;    // columns [128, 192)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_QBYTES);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + h * ATT_D, q0, b);
      const long vrow = p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D;
      for (int i = 0; i < 2 * nblk; ++i) {
        const int j = i % nblk;
        const int st = i & 1;
        mbar_wait(&k_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], ATT2_KBYTES);
        tma_load_3d(sK + st * ATT2_KBYTES, &tmK, &k_full[st], p.k_col0 + h * ATT_D, j * ATT2_BK, b);
        if (i >= nblk) {
          const int iv = i - nblk;
          const int sv = iv & 1;
          mbar_wait(&v_empty[sv], ((iv >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[sv], ATT2_VBYTES);
          tma_load_2d(sV + sv * ATT2_VBYTES, &tmVt, &v_full[sv], j * ATT2_BK, static_cast<int>(vrow));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (nblk > 0 && elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(ATT_BQ, ATT2_BK);   // S: 128 x 128, K = 64
      constexpr uint32_t idesc_o = umma_idesc_f16(ATT_BQ, ATT_D);    // O: 128 x 64,  K = 128
      auto issue_pv = [&](int iv) {
        const int sv = iv & 1;
        const uint32_t par = (iv >> 1) & 1;
        mbar_wait(&p_full[sv], par);
        mbar_wait(&v_full[sv], par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < ATT2_BK / 16; ++k) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sP + sv * ATT2_PBYTES)) + 2 * k;
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sV + sv * ATT2_VBYTES)) + 2 * k;
          umma_f16(tmem_O, adesc, bdesc, idesc_o, (iv > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[sv]);
        umma_commit(&v_empty[sv]);
      };
      mbar_wait(q_full, 0);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      for (int i = 0; i < 2 * nblk; ++i) {
        const int st = i & 1;
        const uint32_t par = (i >> 1) & 1;
        mbar_wait(&k_full[st], par);
        mbar_wait(&s_empty[st], par ^ 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + st * ATT2_KBYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16(tmem_S + st * ATT2_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
        if (i > nblk) issue_pv(i - nblk - 1);
      }
      issue_pv(nblk - 1);
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quad = warp & 3;                   // TMEM lane quadrant (hardware rule: warp_id % 4)
    const int grp = (warp - 2) >> 2;             // which 32 keys of each 128-key block this warp owns
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool row_dead = (q >= qlen) || (klen <= 0);      // fully masked row -> uniform over Tk
    const bool row_store = q < p.Tq;
    const float sl2 = p.scale * 1.4426950408889634f;       // scale * log2(e)
    const float inv_tk = 1.0f / static_cast<float>(p.Tk);
    uint32_t v[32];
    auto softmax_bar = []() { asm volatile("bar.sync 1, %0;" ::"n"(ATT2_SOFTMAX_THREADS) : "memory"); };
    float* red_m = red;                              // [G][128]
    float* red_l = red + ATT2_GROUPS * ATT_BQ;        // [G][128]
    float* red_s = red + 2 * ATT2_GROUPS * ATT_BQ;    // [G][128]

    if (has_dead_rows) {
      // column mean of V over ALL Tk padded keys (the uniform distribution of attention.py:240-242), fp32
      const int sidx = (warp - 2) * 32 + lane;      // 0..255: 4 threads per head channel
      const int d = sidx >> 2, part = sidx & 3;
      const __half* vrow = p.vt + (p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D + d) * p.vt_ld;
      float acc = 0.f;
      for (int t0 = part * 8; t0 < p.Tk; t0 += 32) {
        if (t0 + 8 <= p.Tk) {
          const uint4 u = *reinterpret_cast<const uint4*>(vrow + t0);
          const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc += f.x + f.y;
          }
        } else {
          for (int t = t0; t < p.Tk; ++t) acc += __half2float(vrow[t]);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) vmean[d] = acc * inv_tk;
    }

    // ---- pass 1: row maximum (and denominator if the alignments are written)
    float m = -INFINITY;
    float l = 0.f;
    for (int i = 0; i < nblk; ++i) {
      const int st = i & 1;
      mbar_wait(&s_full[st], (i >> 1) & 1);
      tc_fence_after();
      __syncwarp();
      tmem_ld32(tmem_S + st * ATT2_BK + lane_off + grp * 32, v);
      tmem_wait_ld();
      const int kk0 = i * ATT2_BK + grp * 32;
      float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = kk0 + e;
        const bool ok = (kk < klen) && (!p.causal || kk <= q);
        const float sv = ok ? __uint_as_float(v[e]) : -INFINITY;
        bm[e & 3] = fmaxf(bm[e & 3], sv);
        if (kWriteAli) v[e] = __float_as_uint(sv);
      }
      const float cm = fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3]));
      if (kWriteAli) {
        const float mn = fmaxf(m, cm);
        if (mn > -INFINITY) {
          float add[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 32; ++e) add[e & 3] += ex2_approx((__uint_as_float(v[e]) - mn) * sl2);
          l = l * ex2_approx((m - mn) * sl2) + (add[0] + add[1]) + (add[2] + add[3]);
          m = mn;
        }
      } else {
        m = fmaxf(m, cm);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[st]);
    }
    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[2] = t; }
    // combine the column groups of every row
    red_m[grp * ATT_BQ + r] = m;
    if (kWriteAli) red_l[grp * ATT_BQ + r] = l;
    softmax_bar();
    {
      float mn = m;
#pragma unroll
      for (int g = 0; g < ATT2_GROUPS; ++g) mn = fmaxf(mn, red_m[g * ATT_BQ + r]);
      if (kWriteAli) {
        float lt = 0.f;
        if (mn > -INFINITY) {
#pragma unroll
          for (int g = 0; g < ATT2_GROUPS; ++g) {
            const float mg = red_m[g * ATT_BQ + r];
            if (mg > -INFINITY) lt += red_l[g * ATT_BQ + r] * ex2_approx((mg - mn) * sl2);
          }
        }
        l = lt;
      }
      m = mn;
    }
    if (row_dead) { m = 0.f; l = 1.f; }

    // ---- pass 2: probabilities -> shared memory (A operand of P V), denominators, alignments
    float ls[4] = {0.f, 0.f, 0.f, 0.f};
    const float inv_l = kWriteAli ? 1.0f / l : 1.0f;
    const float msl2 = m * sl2;
    for (int iv = 0; iv < nblk; ++iv) {
      const int i = nblk + iv;
      const int st = i & 1;
      const int sp = iv & 1;
      mbar_wait(&s_full[st], (i >> 1) & 1);
      mbar_wait(&p_empty[sp], ((iv >> 1) & 1) ^ 1);
      tc_fence_after();
      // this warp's 32 keys of the 64-key block: 128 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)
      uint8_t* prow = sP + sp * ATT2_PBYTES + r * 128;
      __syncwarp();
      tmem_ld32(tmem_S + st * ATT2_BK + lane_off + grp * 32, v);
      tmem_wait_ld();
      const int kk0 = iv * ATT2_BK + grp * 32;
      float pr[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = kk0 + e;
        float pe;
        if (row_dead) {
          pe = 0.f;                 // dead rows take the V column mean directly (epilogue)
        } else {
          const bool ok = (kk < klen) && (!p.causal || kk <= q);
          pe = ok ? ex2_approx(__uint_as_float(v[e]) * sl2 - msl2) : 0.f;
        }
        ls[e & 3] += pe;
        pr[e] = pe * inv_l;           // normalised already when kWriteAli (inv_l == 1 otherwise)
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = grp * 4 + g;
        uint4 u;
        u.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]);
        u.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
        u.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]);
        u.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = u;
      }
      if (kWriteAli && row_store) {
        float* arow = p.ali + ((static_cast<long>(b) * p.H + h) * p.Tq + q) * p.Tk + kk0;
        for (int e = 0; e < 32 && kk0 + e < p.Tk; ++e) arow[e] = row_dead ? inv_tk : pr[e];
      }
      fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&p_full[sp]);
      mbar_arrive(&s_empty[st]);
    }
    // alignments of skipped key blocks: exact zeros for live rows, uniform 1/Tk for fully masked rows
    if (kWriteAli && row_store) {
      const int k_done = nblk * ATT2_BK;
      float* arow = p.ali + ((static_cast<long>(b) * p.H + h) * p.Tq + q) * p.Tk;
      const float fill = row_dead ? inv_tk : 0.f;
      for (int kk = k_done + grp; kk < p.Tk; kk += ATT2_GROUPS) arow[kk] = fill;
    }

    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[3] = t; }
    // ---- epilogue: ctx = O / l ; column groups 0 and 1 each write 32 of the 64 head channels
    float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
    red_s[grp * ATT_BQ + r] = lsum;
    softmax_bar();
    if (grp < 2) {
      lsum = 0.f;
#pragma unroll
      for (int g = 0; g < ATT2_GROUPS; ++g) lsum += red_s[g * ATT_BQ + r];
      if (nblk > 0) {
        mbar_wait(o_full, 0);
        tc_fence_after();
        __syncwarp();
        tmem_ld32(tmem_O + lane_off + grp * 32, v);
        tmem_wait_ld();
      }
      const float on = row_dead ? 0.f : (kWriteAli ? 1.0f : 1.0f / lsum);
      if (p.lse2 && grp == 0 && row_store)
        p.lse2[(static_cast<long>(b) * p.H + h) * p.Tq + q] = row_dead ? 0.f : msl2 + log2f(lsum);
      if (row_store) {
        __half* dst = p.ctx + (static_cast<long>(b) * p.Tq + q) * p.ctx_ld + h * ATT_D + grp * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = row_dead ? vmean[grp * 32 + j + e] : __uint_as_float(v[j + e]) * on;
          uint4 u;
          u.x = pack_half2(f[0], f[1]);
          u.y = pack_half2(f[2], f[3]);
          u.z = pack_half2(f[4], f[5]);
          u.w = pack_half2(f[6], f[7]);
          *reinterpret_cast<uint4*>(dst + j) = u;
        }
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[4] = t; }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}


// ================================================================================================ v3: scores resident in TMEM
// Causal self-attention (query_lengths == key_lengths, the CrossAttentionBLK self-attention of attention.py:437-439) with
// T <= 448: ONE tensor-core pass.  All key blocks of the tile are loaded once, every S_j = Q K_j^T is issued up front
// into its own TMEM columns (4 x 128 score columns + 64 for O = 512) and stays there: the softmax warps sweep the
// resident scores twice (row maximum, then p = exp2(...) -> fp16 P_j in shared memory -> O += P_j V_j) -- no second
// Q K^T, no second sweep over K.  Only the diagonal block needs a mask (key <= query); fully masked rows (query >= length)
// attend uniformly over all T keys as in the reference (attention.py:240-242) and take the column mean of V.
// Tiles are scheduled heaviest first (blockIdx.y = 0 is the LAST query tile, which sees the most key blocks).
constexpr int AT3_THREADS = 576;
constexpr int AT3_MAXBLK = 4;
constexpr int AT3_TMAX = 448;
constexpr int AT3_OFF_K = ATT_QBYTES;
constexpr int AT3_OFF_V = AT3_OFF_K + AT3_MAXBLK * ATT_KBYTES;
constexpr int AT3_OFF_P = AT3_OFF_V + AT3_MAXBLK * ATT_VBYTES;
constexpr int AT3_OFF_BARS = AT3_OFF_P + 2 * ATT_PBYTES;
constexpr int AT3_SMEM = AT3_OFF_BARS + 256 + (2 * 4 * 128 + 64) * 4 + 1024;

__global__ void __launch_bounds__(AT3_THREADS, 1)
attention3_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT3_OFF_K;
  uint8_t* sV = smem + AT3_OFF_V;
  uint8_t* sP = smem + AT3_OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT3_OFF_BARS);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [4]
  uint64_t* v_full = bars + 5;    // [4]
  uint64_t* s_full = bars + 9;    // [4]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* red = reinterpret_cast<float*>(bars + 32);   // [2][4][128]
  float* vmean = red + 2 * 4 * 128;                   // [64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = gridDim.y;
  const int qt = nq - 1 - static_cast<int>(blockIdx.y);   // heaviest tiles first
  const int q0 = qt * ATT_BQ;
  const int h = blockIdx.x % p.H;
  const int b = blockIdx.x / p.H;
  const int len = __ldg(p.q_len + b);
  const int q_hi = min(q0 + ATT_BQ, p.Tq);
  const bool has_dead_rows = max(q0, len) < q_hi || len <= 0;
  const bool all_dead = q0 >= len || len <= 0;
  // key blocks that can contribute to the live rows: causal (keys <= last live query of the tile)
  const int nblk = all_dead ? 0 : (min(q_hi, len) - 1) / ATT_BK + 1;
  auto blk_n = [&](int j) -> int {   // score columns of block j (multiple of 16): zero-filled keys beyond T are masked causally
    return min(ATT_BK, ((p.Tk - j * ATT_BK + 15) >> 4) << 4);
  };
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < AT3_MAXBLK; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&s_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 16);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_O = tmem_base + AT3_TMAX;

  if (warp == 0) {
    // ===================== TMA producer: everything is resident, no ring =====================
    if (nblk > 0 && elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_QBYTES);
      tma_load_3d(sQ, &tmQ, q_full, p.q_col0 + h * ATT_D, q0, b);
      for (int j = 0; j < nblk; ++j) {
        mbar_arrive_expect_tx(&k_full[j], ATT_KBYTES);
        tma_load_3d(sK + j * ATT_KBYTES, &tmK, &k_full[j], p.k_col0 + h * ATT_D, j * ATT_BK, b);
      }
      const long vrow = p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D;
      for (int j = 0; j < nblk; ++j) {
        const int npanel = blk_n(j) > 64 ? 2 : 1;
        mbar_arrive_expect_tx(&v_full[j], npanel * (ATT_VBYTES / 2));
        for (int q = 0; q < npanel; ++q)
          tma_load_2d(sV + j * ATT_VBYTES + q * (ATT_VBYTES / 2), &tmVt, &v_full[j], j * ATT_BK + q * 64, static_cast<int>(vrow));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (nblk > 0 && elect_one()) {
      constexpr uint32_t idesc_o = umma_idesc_f16(ATT_BQ, ATT_D);
      mbar_wait(q_full, 0);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_full[j], 0);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_f16(ATT_BQ, static_cast<uint32_t>(blk_n(j)));
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + j * ATT_KBYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16(tmem_base + j * ATT_BK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[j]);
      }
      for (int j = 0; j < nblk; ++j) {
        const int sp = j & 1;
        mbar_wait(&p_full[sp], (j >> 1) & 1);
        mbar_wait(&v_full[j], 0);
        tc_fence_after();
        const int nsteps = blk_n(j) >> 4;
        for (int kk = 0; kk < nsteps; ++kk) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sP + sp * ATT_PBYTES + (kk >> 2) * (ATT_PBYTES / 2))) + 2 * (kk & 3);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sV + j * ATT_VBYTES + (kk >> 2) * (ATT_VBYTES / 2))) + 2 * (kk & 3);
          umma_f16(tmem_O, adesc, bdesc, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[sp]);
      }
      umma_commit(o_full);
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;             // 32 score columns of every 128-key block
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const bool row_dead = (q >= len) || (len <= 0);
    const bool row_store = q < p.Tq;
    const float sl2 = p.scale * 1.4426950408889634f;
    uint32_t v[32];
    auto softmax_bar = []() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    if (has_dead_rows) {
      // column mean of V over ALL Tk keys (uniform attention of a fully masked row), 8 threads per head channel
      const int sidx = (warp - 2) * 32 + lane;
      const int d = sidx >> 3, part = sidx & 7;
      const __half* vrow = p.vt + (p.vt_row0 + (static_cast<long>(b) * p.H + h) * ATT_D + d) * p.vt_ld;
      float acc = 0.f;
      for (int t = part; t < p.Tk; t += 8) acc += __half2float(vrow[t]);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (part == 0) vmean[d] = acc / static_cast<float>(p.Tk);
    }
    // ---- sweep 1: row maximum over the resident scores
    float m = -INFINITY;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j], 0);
      tc_fence_after();
      if (grp * 32 < blk_n(j)) {
        tmem_ld32(tmem_base + lane_off + j * ATT_BK + grp * 32, v);
        tmem_wait_ld();
        const int kk0 = j * ATT_BK + grp * 32;
        if (kk0 + 31 <= q) {                       // whole chunk below the diagonal (warp-divergent only in the diagonal block)
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, __uint_as_float(v[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, (kk0 + e <= q) ? __uint_as_float(v[e]) : -INFINITY);
        }
      }
    }
    red[grp * 128 + r] = m;
    softmax_bar();
    m = fmaxf(fmaxf(red[r], red[128 + r]), fmaxf(red[256 + r], red[384 + r]));
    const float msl2 = row_dead ? INFINITY : m * sl2;   // live row: key 0 is always visible, m is finite
    // ---- sweep 2: probabilities -> shared memory (A operand of P V)
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int sp = j & 1;
      if (j >= 2) mbar_wait(&p_empty[sp], ((j >> 1) - 1) & 1);
      if (grp * 32 < blk_n(j)) {
        tc_fence_after();
        tmem_ld32(tmem_base + lane_off + j * ATT_BK + grp * 32, v);
        tmem_wait_ld();
        const int kk0 = j * ATT_BK + grp * 32;
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float pe = ex2_approx(fmaf(__uint_as_float(v[e]), sl2, -msl2));
          if (kk0 + e > q) pe = 0.f;
          l += pe;
          pr[e] = pe;
        }
        uint8_t* prow = sP + sp * ATT_PBYTES + (grp >> 1) * (ATT_PBYTES / 2) + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]);
          u.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
          u.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]);
          u.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + ((((grp & 1) * 4 + g) ^ (r & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      warp_arrive(&p_full[sp]);
    }
    red[512 + grp * 128 + r] = l;
    softmax_bar();
    l = (red[512 + r] + red[640 + r]) + (red[768 + r] + red[896 + r]);
    // ---- epilogue: ctx = O / l, 16 head channels per thread
    if (nblk > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
      tmem_ld16(tmem_O + lane_off + grp * 16, v);
      tmem_wait_ld();
    }
    if (p.lse2 && grp == 0 && row_store)   // training: log2 of the softmax denominator incl. the maximum (backward pass)
      p.lse2[(static_cast<long>(b) * p.H + h) * p.Tq + q] = row_dead ? 0.f : msl2 + log2f(l);
    if (row_store) {
      const float on = row_dead ? 0.f : 1.0f / l;
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] = row_dead ? vmean[grp * 16 + e] : __uint_as_float(v[e]) * on;
      __half* dst = p.ctx + (static_cast<long>(b) * p.Tq + q) * p.ctx_ld + h * ATT_D + grp * 16;
      uint4 ua, ub;
      ua.x = pack_half2(f[0], f[1]); ua.y = pack_half2(f[2], f[3]); ua.z = pack_half2(f[4], f[5]); ua.w = pack_half2(f[6], f[7]);
      ub.x = pack_half2(f[8], f[9]); ub.y = pack_half2(f[10], f[11]); ub.z = pack_half2(f[12], f[13]); ub.w = pack_half2(f[14], f[15]);
      *reinterpret_cast<uint4*>(dst) = ua;
      *reinterpret_cast<uint4*>(dst + 8) = ub;
    }
    tc_fence_before();
  }
  pdl_launch_dependents();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vb

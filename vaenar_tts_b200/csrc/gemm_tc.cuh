// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues.
//
//   C[M, N] = epilogue( sum_seg A_seg[M (+row shift), K_seg] * W[N, K_total]^T )
//
// * A operands are fp16, row-major [batch, rows, K] behind 3-D TMA tensor maps (128-byte swizzle).
//   Up to kMaxSegs K-segments, each choosing one of two tensor maps and a row shift:
//     - plain Dense:            1 segment
//     - Dense over a concat:    2 segments, 2 maps  (modules/attention.py:410,440,447  [x ; ctx] W)
//     - Conv1D k=5 'same':      5 segments, row shift -2..2; TMA zero-fills rows outside [0, T)
//                               (modules/utils.py:56-85) -> implicit GEMM, no im2col buffer
//     - split-fp16 ("3x"):      segments (hi, lo, hi) against packed weights [Whi | Whi | Wlo]
// * W is packed fp16 [N, K_total] (K-major), accumulators are fp32 in TMEM.
// * One CTA = one 128 x BLOCK_N output tile.  Warp 0: TMA producer, warp 1: MMA issuer (one elected
//   thread), warps 2..5: epilogue (thread == output row, TMEM lane == row).
#pragma once
#include "ptx.cuh"

namespace vb {

constexpr int kMaxSegs = 16;
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;

enum EpiMode : int {
  EPI_PLAIN = 0,     // act(acc + bias) [*scale + shift] [+ a * table[t]] [+ residual] -> f32 / f16 / f16-lo
  EPI_LN = 1,        // LayerNorm(acc + bias + residual) -> f32 + f16        (BLOCK_N == N)
  EPI_QKV = 2,       // columns < n_rowmajor -> f16 row-major ; the rest -> V^T [blk, b, h, 64, Tpad]
  EPI_COUPLING = 3,  // affine coupling update of z in place + per-row log-det     (BLOCK_N == N == 128)
  EPI_POSTERIOR = 4, // [logvar_named | mu_named] -> z = eps*exp(.5*lv)+mu, per-row log q  (N == 256)
};

struct GemmParams {
  // ---- tiling
  int batches;           // A-map batch extent (1 for flat row tiling)
  int rows;              // rows per batch
  int tiles_per_batch;   // ceil(rows / 128)
  int N;                 // valid output columns
  int nseg;
  int seg_map[kMaxSegs];
  int seg_shift[kMaxSegs];
  int seg_kblocks[kMaxSegs];
  int alg_k;             // algorithmic K (host-side accounting only)
  // ---- sequence geometry of the flattened rows (row = b * seq_T + t)
  int seq_T;
  int seq_B;
  // ---- epilogue
  int mode;
  int act;                       // 0 none, 1 relu, 2 tanh
  const float* bias;             // [N] or null
  const float* ch_scale;         // per-channel affine after the activation (inference BatchNorm), or null
  const float* ch_shift;
  const float* add_table;        // [seq_T, add_ld] table added as (*add_scale) * table[t, n]  (positional enc.)
  const float* add_scale;
  int add_ld;
  const float* residual;         // [M, res_ld] fp32 or null
  int res_ld;
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  float* out_f32;
  int ld_f32;
  __half* out_h;
  __half* out_lo;                // optional fp16 residual part (x - fp16(x)) for split-fp16 consumers
  int ld_h;
  // EPI_QKV
  int n_rowmajor;
  __half* vt;
  int vt_ld;
  int heads;
  // EPI_COUPLING / EPI_POSTERIOR
  float* z;                      // [M, z_ld] fp32 latent (updated in place)
  __half* z_h;                   // fp16 copy of z
  int z_ld;
  int zp_off;                    // column offset of the transformed half
  int backward;                  // 0: zp*scale+shift ; 1: (zp-shift)/(scale+1e-12)
  float* row_acc;                // [M] per-row accumulator (log-det / log q), accumulated (+=)
  const int* lengths;            // [seq_B]
  const float* eps_in;           // EPI_POSTERIOR noise [M, z_ld]
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N <= 128) ? 6 : (BLOCK_N <= 256 ? 4 : 2);
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kUmmaN = BLOCK_N > 256 ? 256 : BLOCK_N;
  static constexpr int kNumUmmaN = BLOCK_N / kUmmaN;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return tanhf(v);
  return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int tile_b = m_tile / p.tiles_per_batch;
  const int tile_t0 = (m_tile % p.tiles_per_batch) * GEMM_BLOCK_M;

  int total_kblocks = 0;
  for (int s = 0; s < p.nseg; ++s) total_kblocks += p.seg_kblocks[s];

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int kglobal = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const CUtensorMap* tmA = p.seg_map[s] ? &tmA1 : &tmA0;
        const int row0 = tile_t0 + p.seg_shift[s];
        for (int kb = 0; kb < p.seg_kblocks[s]; ++kb, ++kglobal) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_3d(smem_a + stage * Cfg::kABytes, tmA, &full_bar[stage], kb * GEMM_BLOCK_K, row0, tile_b);
#pragma unroll
          for (int h = 0; h < Cfg::kNumUmmaN; ++h)
            tma_load_2d(smem_b + stage * Cfg::kBBytes + h * Cfg::kUmmaN * GEMM_BLOCK_K * 2, &tmB, &full_bar[stage],
                        kglobal * GEMM_BLOCK_K, n_tile * BLOCK_N + h * Cfg::kUmmaN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(GEMM_BLOCK_M, Cfg::kUmmaN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < total_kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
#pragma unroll
        for (int h = 0; h < Cfg::kNumUmmaN; ++h) {
          const uint64_t bdesc =
              umma_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes + h * Cfg::kUmmaN * GEMM_BLOCK_K * 2));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (>>4) address field
            umma_f16(tmem_base + h * Cfg::kUmmaN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                  // row inside the tile
    const int t_in_batch = tile_t0 + r;
    const bool row_ok = t_in_batch < p.rows;
    const long grow = static_cast<long>(tile_b) * p.rows + t_in_batch;   // flattened row
    const int sb = static_cast<int>(grow / p.seq_T);
    const int st = static_cast<int>(grow % p.seq_T);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint32_t v[32];

    if (p.mode == EPI_PLAIN || p.mode == EPI_QKV) {
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int n0 = n_tile * BLOCK_N + c * 32;
        if (n0 >= p.N) break;
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        if (!row_ok) continue;   // predicated stores only; the warp re-converges at __syncwarp()
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + j;
          float x = __uint_as_float(v[j]);
          if (n < p.N) {
            if (p.bias) x += __ldg(p.bias + n);
            x = apply_act(x, p.act);
            if (p.ch_scale) x = x * __ldg(p.ch_scale + n) + __ldg(p.ch_shift + n);
            if (p.add_table) x += __ldg(p.add_scale) * __ldg(p.add_table + static_cast<long>(st) * p.add_ld + n);
            if (p.residual) x += __ldg(p.residual + grow * p.res_ld + n);
          }
          f[j] = x;
        }
        if (p.mode == EPI_QKV && n0 >= p.n_rowmajor) {
          // V^T store: [blk][b][h][64][vt_ld]; consecutive lanes = consecutive t -> coalesced
          const int nv = n0 - p.n_rowmajor;
          const int hd = p.heads * 64;
          const int blk = nv / hd;
          const int h = (nv % hd) / 64;
          const int d0 = nv % 64;
          __half* dst = p.vt + ((static_cast<long>(blk) * p.seq_B + sb) * p.heads + h) * 64 * p.vt_ld + st;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[static_cast<long>(d0 + j) * p.vt_ld] = __float2half_rn(f[j]);
          continue;
        }
        const bool full = (n0 + 32 <= p.N);
        if (p.out_f32) {
          float* dst = p.out_f32 + grow * p.ld_f32 + n0;
          if (full && (p.ld_f32 & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32 && n0 + j < p.N; ++j) dst[j] = f[j];
          }
        }
        if (p.out_h) {
          __half* dst = p.out_h + grow * p.ld_h + n0;
          __half* dlo = p.out_lo ? p.out_lo + grow * p.ld_h + n0 : nullptr;
          if (full && (p.ld_h & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_half2(f[j], f[j + 1]);
              u.y = pack_half2(f[j + 2], f[j + 3]);
              u.z = pack_half2(f[j + 4], f[j + 5]);
              u.w = pack_half2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = u;
            }
            if (dlo) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float g[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) g[e] = f[j + e] - __half2float(__float2half_rn(f[j + e]));
                uint4 u;
                u.x = pack_half2(g[0], g[1]);
                u.y = pack_half2(g[2], g[3]);
                u.z = pack_half2(g[4], g[5]);
                u.w = pack_half2(g[6], g[7]);
                *reinterpret_cast<uint4*>(dlo + j) = u;
              }
            }
          } else {
            for (int j = 0; j < 32 && n0 + j < p.N; ++j) {
              const __half hi = __float2half_rn(f[j]);
              dst[j] = hi;
              if (dlo) dlo[j] = __float2half_rn(f[j] - __half2float(hi));
            }
          }
        }
      }
    } else if (p.mode == EPI_LN) {
      // pass 1: x = acc + bias + residual ; row sum ; x written back to TMEM
      float sum = 0.f;
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        const int n0 = c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && p.residual) rs = *reinterpret_cast<const float4*>(p.residual + grow * p.res_ld + n0 + j);
          const float rr[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = __uint_as_float(v[j + e]) + rr[e];
            if (p.bias) x += __ldg(p.bias + n0 + j + e);
            sum += x;
            v[j + e] = __float_as_uint(x);
          }
        }
        __syncwarp();
        tmem_st32(taddr + c * 32, v);
      }
      tmem_wait_st();
      const float mean = sum * (1.f / BLOCK_N);
      // pass 2: centred second moment
      float sq = 0.f;
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(v[j]) - mean;
          sq += d * d;
        }
      }
      const float rstd = 1.f / sqrtf(sq * (1.f / BLOCK_N) + p.ln_eps);
      // pass 3: normalise + store
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_wait_ld();
        if (!row_ok) continue;
        const int n0 = c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          f[j] = (__uint_as_float(v[j]) - mean) * rstd * __ldg(p.ln_gamma + n0 + j) + __ldg(p.ln_beta + n0 + j);
        if (p.out_f32) {
          float* dst = p.out_f32 + grow * p.ld_f32 + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        }
        if (p.out_h) {
          __half* dst = p.out_h + grow * p.ld_h + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_half2(f[j], f[j + 1]);
            u.y = pack_half2(f[j + 2], f[j + 3]);
            u.z = pack_half2(f[j + 4], f[j + 5]);
            u.w = pack_half2(f[j + 6], f[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = u;
          }
        }
      }
    } else if (p.mode == EPI_COUPLING) {
      // columns [0, half) = log_scale, [half, 2*half) = shift, half = N / 2 (modules/flow.py:223-257)
      const int half = p.N >> 1;
      const bool in_len = row_ok && (st < __ldg(p.lengths + sb));
      float logdet = 0.f;
      uint32_t w[32];
      for (int c = 0; c < half / 32; ++c) {
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_ld32(taddr + half + c * 32, w);
        tmem_wait_ld();
        if (!row_ok) continue;
        float* zrow = p.z + grow * p.z_ld + p.zp_off + c * 32;
        __half* zh = p.z_h + grow * p.z_ld + p.zp_off + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float ls = __uint_as_float(v[j]) + __ldg(p.bias + c * 32 + j);
          const float sh = __uint_as_float(w[j]) + __ldg(p.bias + half + c * 32 + j);
          const float scale = 1.f / (1.f + expf(-(ls + 2.0f)));
          const float zp = zrow[j];
          const float o = p.backward ? (zp - sh) / (scale + 1e-12f) : scale * zp + sh;
          zrow[j] = o;
          zh[j] = __float2half_rn(o);
          logdet += logf(scale);
        }
      }
      if (row_ok) p.row_acc[grow] += in_len ? (p.backward ? -logdet : logdet) : 0.f;
    } else if (p.mode == EPI_POSTERIOR) {
      // modules/posterior.py:20-72 with the models.py:136 name swap already applied by the packing order:
      // columns [0, L) = log-variance (mu_projection), [L, 2L) = mean (logvar_projection), L = N / 2.
      const int L = p.N >> 1;
      const bool in_len = row_ok && (st < __ldg(p.lengths + sb));
      float acc = 0.f;
      uint32_t w[32];
      for (int c = 0; c < L / 32; ++c) {
        __syncwarp();
        tmem_ld32(taddr + c * 32, v);
        tmem_ld32(taddr + L + c * 32, w);
        tmem_wait_ld();
        if (!row_ok) continue;
        const float* er = p.eps_in + grow * p.z_ld + c * 32;
        float* zrow = p.z + grow * p.z_ld + c * 32;
        __half* zh = p.z_h + grow * p.z_ld + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float lv = __uint_as_float(v[j]) + __ldg(p.bias + c * 32 + j);
          const float mu = __uint_as_float(w[j]) + __ldg(p.bias + L + c * 32 + j);
          const float e = er[j];
          const float o = e * expf(0.5f * lv) + mu;
          zrow[j] = o;
          zh[j] = __float2half_rn(o);
          acc += lv + e * e;
        }
      }
      if (row_ok) p.row_acc[grow] += in_len ? -0.5f * (static_cast<float>(L) * 1.8378770664093453f + acc) : 0.f;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace vb

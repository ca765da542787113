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
//   thread), warps 2..9: epilogue (thread == output row, TMEM lane == row; two warps per TMEM lane
//   quadrant, each taking alternate 64-column chunks; LayerNorm row statistics are combined through shared memory).
#pragma once
#include "ptx.cuh"

namespace vb {

constexpr int kMaxSegs = 16;
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int GEMM_THREADS_LN = 576;       // LayerNorm epilogue: warps 2..17 (four per TMEM lane quadrant)
__host__ __device__ constexpr int gemm_threads(int mode) { return mode == 1 /*EPI_LN*/ ? GEMM_THREADS_LN : GEMM_THREADS; }

enum EpiMode : int {
  EPI_PLAIN = 0,     // act(acc + bias) [*scale + shift] [+ a * table[t]] [+ residual] -> f32 / f16 / f16-lo
  EPI_LN = 1,        // LayerNorm(acc + bias + residual) -> f32 + f16        (BLOCK_N == N)
  EPI_QKV = 2,       // columns < n_rowmajor -> f16 row-major ; the rest -> V^T [blk, b, h, 64, Tpad]
  EPI_COUPLING = 3,  // affine coupling update of z in place + per-row log-det     (BLOCK_N == N == 128)
  EPI_POSTERIOR = 4, // [logvar_named | mu_named] -> z = eps*exp(.5*lv)+mu, per-row log q  (N == 256)
};

// Compile-time epilogue feature mask of EPI_PLAIN (F_RUNTIME = decide from the GemmParams pointers at run time)
enum : uint32_t {
  F_BIAS = 1, F_RELU = 2, F_TANH = 4, F_BN = 8, F_TABLE = 16, F_RES = 32, F_OUT_F32 = 64, F_OUT_H = 128, F_OUT_LO = 256,
  F_MASK = 512,   // out = acc where mask_h != 0, else 0 (ReLU backward inside the dgrad GEMM); excludes F_RES / F_OUT_F32
  F_RUNTIME = 0x80000000u
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
  unsigned long long* dbg;   // optional per-CTA phase timestamps (tuning aid), 8 x u64 per CTA
  int ln_cluster;        // EPI_LN: N is split over a 2-CTA cluster (blockIdx.y = rank), stats exchanged via DSMEM
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
  const __half* mask_h;          // [M, mask_ld] fp16 or null: outputs are zeroed where the mask element is 0 (post-ReLU activation)
  int mask_ld;
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
  // ---- training: tensors saved for the backward pass (all optional)
  float* ln_rstd;                // EPI_LN: [M] reciprocal standard deviation of every row
  int v_rowmajor;                // EPI_QKV: the V columns are ALSO written row-major into out_h (next to V^T)
  float* save_scale;             // EPI_COUPLING: [M, N/2] sigmoid(log_scale + 2)
  float* save_lv;                // EPI_POSTERIOR: [M, N/2] log-variance
};

// OCC = resident CTAs per SM the instance is built for.  OCC 2 (BLOCK_N 128, EPI_PLAIN without the fp16-lo output): 3
// pipeline stages (96 KB, which the 8 x 12 KB epilogue slabs fill exactly) and <= 102 registers, so that the epilogue of one
// CTA (2-5 us of TMEM reads, conversions and stores) runs under the TMA / MMA phase of its neighbour: at d_model 256 the
// mainloop of a tile is 4-16 k-blocks and the epilogue dominates a one-CTA-per-SM schedule.
template <int BLOCK_N, int OCC = 1>
struct GemmCfg {
  static_assert(OCC == 1 || BLOCK_N == 128, "two CTAs per SM: BLOCK_N 128 only");
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = OCC == 2 ? 3 : ((BLOCK_N <= 128) ? 6 : (BLOCK_N <= 256 ? 4 : 2));
  static constexpr int kSlabBytes = OCC == 2 ? 3 * 4096 : 4 * 4096;   // per epilogue warp: [res0 | f32 0][res1 | f32 1][f16]([f16-lo])
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kSmemBytes =
      (kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*cluster LN exchange*/ + 1023) / 1024 * 1024;
  static constexpr int kUmmaN = BLOCK_N > 256 ? 256 : BLOCK_N;
  static constexpr int kNumUmmaN = BLOCK_N / kUmmaN;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return tanhf(v);
  return v;
}

// optional in-kernel phase timestamps (debug / tuning only): 8 x u64 per CTA, globaltimer ns
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int BLOCK_N, int MODE, uint32_t FEAT, int OCC = 1>
__global__ void __launch_bounds__(gemm_threads(MODE), OCC)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, OCC>;
  static_assert(OCC == 1 || (MODE == EPI_PLAIN && !(FEAT & (F_OUT_LO | F_RUNTIME))),
                "two CTAs per SM: plain epilogue with a compile-time feature mask and no fp16-lo output (12 KB slabs)");
  extern __shared__ __align__(1024) uint8_t smem[];   // stays in the shared address space (LDS/STS, not generic)
  if ((smem_u32(smem) & 1023u) != 0) __trap();        // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);
  uint64_t* xbar = bars + 2 * Cfg::kStages + 2;             // [2] cluster LayerNorm exchange
  float* xred = reinterpret_cast<float*>(bars + 2 * Cfg::kStages + 4);   // [2][128] written by the peer CTA

  unsigned long long* dbg = p.dbg ? p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 64) dbg[0] = gtime();
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
    mbar_init(&xbar[0], 128);
    mbar_init(&xbar[1], 128);
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
  if (p.ln_cluster) cluster_sync_all();   // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on we touch memory produced by it.
  pdl_launch_dependents();
  pdl_wait();
  if (dbg && threadIdx.x == 64) dbg[1] = gtime();

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
    // TMEM gives thread == row.  To keep every global access coalesced the accumulator chunk is
    // transposed through shared memory (the pipeline stages are idle once tmem_full has fired):
    // T[row][col] with row stride 65 floats, then lane == column pair and the warp walks the 32 rows.
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                  // row inside the tile (thread == row view)
    const int t_in_batch = tile_t0 + r;
    const bool row_ok = t_in_batch < p.rows;
    const long grow = static_cast<long>(tile_b) * p.rows + t_in_batch;   // flattened row
    const int sb = static_cast<int>(grow / p.seq_T);
    const int st = static_cast<int>(grow % p.seq_T);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const long grow0 = static_cast<long>(tile_b) * p.rows + tile_t0 + quad * 32;   // first row of this warp
    const int rows_here = min(32, p.rows - (tile_t0 + quad * 32));                  // valid rows of this warp (may be <= 0)

    uint32_t v[32], w[32];

    // Warp-private staging slabs (the pipeline stages are idle now): 32 rows x 128 B, 16-byte chunks
    // XOR-swizzled by (row & 7) so that both the thread==row writes and the cooperative copies are
    // bank-conflict free.  [res0 | f32 0][res1 | f32 1][f16][f16-lo], 4 KB each: the fp32 output staging
    // aliases the residual slabs (a thread reads its own residual chunk before overwriting it).
    const int half = (warp - 2) >> 2;                // which alternate 64-column chunks this warp handles
    uint8_t* slab = smem_a + (warp - 2) * Cfg::kSlabBytes;
    uint8_t* s_res = slab;
    uint8_t* s_f32 = slab;
    uint8_t* s_h = slab + 2 * 4096;
    uint8_t* s_lo = slab + 3 * 4096;
    float* red = reinterpret_cast<float*>(smem_a + 8 * 4 * 4096);   // [4][128] row-statistic exchange between column halves
    auto epi_bar = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    auto sw = [&](uint8_t* base, int row, int chunk) -> uint4* {
      return reinterpret_cast<uint4*>(base + row * 128 + ((chunk ^ (row & 7)) << 4));
    };
    // cooperative, coalesced copy of a [rows_here x 32] fp32 global tile <-> slab (4 rows x 128 B per instruction)
    auto load_f32_slab = [&](uint8_t* base, const float* g, int ld, int col0, int ncols_valid) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), chunk = lane & 7;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows_here && chunk * 4 < ncols_valid)
          u = *reinterpret_cast<const uint4*>(g + (grow0 + row) * ld + col0 + chunk * 4);   // plain load: z is updated in place
        *sw(base, row, chunk) = u;
      }
    };
    auto store_f32_slab = [&](uint8_t* base, float* g, int ld, int col0, int ncols_valid) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), chunk = lane & 7;
        if (row < rows_here && chunk * 4 < ncols_valid)
          *reinterpret_cast<uint4*>(g + (grow0 + row) * ld + col0 + chunk * 4) = *sw(base, row, chunk);
      }
    };
    auto load_f16_slab = [&](uint8_t* base, const __half* g, int ld, int col0, int ncols_valid) {   // [rows_here x 64] fp16
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), chunk = lane & 7;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows_here && chunk * 8 < ncols_valid) u = __ldg(reinterpret_cast<const uint4*>(g + (grow0 + row) * ld + col0 + chunk * 8));
        *sw(base, row, chunk) = u;
      }
    };
    auto store_f16_slab = [&](uint8_t* base, __half* g, int ld, int col0, int ncols_valid) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), chunk = lane & 7;
        if (row < rows_here && chunk * 8 < ncols_valid)
          *reinterpret_cast<uint4*>(g + (grow0 + row) * ld + col0 + chunk * 8) = *sw(base, row, chunk);
      }
    };

    // LayerNorm (16 epilogue warps): column group of this warp and its first residual chunk, fetched into registers
    // while the mainloop is still running
    static_assert(MODE != EPI_LN || BLOCK_N == 128 || BLOCK_N == 256, "LayerNorm epilogue: BLOCK_N 128 or 256");
    constexpr int NCH = (MODE == EPI_LN) ? BLOCK_N / 128 : 1;    // 32-column chunks per LayerNorm warp
    const int grp = (warp - 2) >> 2;                             // 0..3 (LayerNorm) / 0..1 (other modes, == half)
    uint4 pre[8];
    if constexpr (MODE == EPI_LN) {
      const int n0p = n_tile * BLOCK_N + grp * 32;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), chunk = lane & 7;
        pre[it] = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows_here) pre[it] = __ldg(reinterpret_cast<const uint4*>(p.residual + (grow0 + row) * p.res_ld + n0p + chunk * 4));
      }
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (dbg && threadIdx.x == 64) dbg[2] = gtime();

    if constexpr (MODE == EPI_PLAIN || MODE == EPI_QKV) {
      constexpr bool RT = (FEAT & F_RUNTIME) != 0;
      const bool f_bias = RT ? (p.bias != nullptr) : ((FEAT & F_BIAS) != 0);
      const bool f_bn = RT ? (p.ch_scale != nullptr) : ((FEAT & F_BN) != 0);
      const bool f_table = RT ? (p.add_table != nullptr) : ((FEAT & F_TABLE) != 0);
      const bool f_res = RT ? (p.residual != nullptr) : ((FEAT & F_RES) != 0);
      const bool f_f32 = RT ? (p.out_f32 != nullptr) : ((FEAT & F_OUT_F32) != 0);
      const bool f_h = RT ? (p.out_h != nullptr) : ((FEAT & F_OUT_H) != 0);
      const bool f_lo = RT ? (p.out_lo != nullptr) : ((FEAT & F_OUT_LO) != 0);
      const bool f_mask = RT ? (p.mask_h != nullptr) : ((FEAT & F_MASK) != 0);
      static_assert(RT || !(FEAT & F_MASK) || !(FEAT & (F_RES | F_OUT_F32)), "the mask tile is staged in the residual / fp32 slab");
      const int act = RT ? p.act : ((FEAT & F_RELU) ? 1 : ((FEAT & F_TANH) ? 2 : 0));
      for (int dc = half; dc < BLOCK_N / 64; dc += 2) {
        const int n0 = n_tile * BLOCK_N + dc * 64;
        if (n0 >= p.N) break;
        const int nvalid = min(64, p.N - n0);
        __syncwarp();
        if (f_res) {
          load_f32_slab(s_res, p.residual, p.res_ld, n0, nvalid);
          load_f32_slab(s_res + 4096, p.residual, p.res_ld, n0 + 32, nvalid - 32);
        }
        if (f_mask) load_f16_slab(s_res, p.mask_h, p.mask_ld, n0, nvalid);
        tmem_ld32(taddr + dc * 64, v);
        tmem_ld32(taddr + dc * 64 + 32, w);
        tmem_wait_ld();
        if (MODE == EPI_QKV && n0 >= p.n_rowmajor) {
          // V^T store: [blk][b][h][64][vt_ld]; thread == row == consecutive t -> already coalesced
          if (row_ok) {
            const int nv = n0 - p.n_rowmajor;
            const int hd = p.heads * 64;
            const int blk = nv / hd;
            const int h = (nv % hd) / 64;
            __half* dst = p.vt + ((static_cast<long>(blk) * p.seq_B + sb) * p.heads + h) * 64 * p.vt_ld + st;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              dst[static_cast<long>(j) * p.vt_ld] = __float2half_rn(__uint_as_float(v[j]));
              dst[static_cast<long>(32 + j) * p.vt_ld] = __float2half_rn(__uint_as_float(w[j]));
            }
          }
          if (!p.v_rowmajor) continue;
        }
        __syncwarp();   // residual slab visible
        const float tscale = f_table ? __ldg(p.add_scale) : 0.f;
        const float* trow = f_table ? p.add_table + static_cast<long>(st) * p.add_ld + n0 : nullptr;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {          // the two 32-column halves of this 64-column chunk
          uint32_t* acc = hh ? w : v;
          const int nb = n0 + hh * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g) {           // 4 columns per step
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __uint_as_float(acc[g * 4 + e]);
            if (nb + g * 4 < p.N) {
              if (f_bias) {
                const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + nb + g * 4));
                x[0] += bq.x; x[1] += bq.y; x[2] += bq.z; x[3] += bq.w;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = apply_act(x[e], act);
              if (f_bn) {
                const float4 sq = __ldg(reinterpret_cast<const float4*>(p.ch_scale + nb + g * 4));
                const float4 hq = __ldg(reinterpret_cast<const float4*>(p.ch_shift + nb + g * 4));
                x[0] = x[0] * sq.x + hq.x; x[1] = x[1] * sq.y + hq.y; x[2] = x[2] * sq.z + hq.z; x[3] = x[3] * sq.w + hq.w;
              }
              if (f_table && row_ok) {
                const float4 tq = __ldg(reinterpret_cast<const float4*>(trow + hh * 32 + g * 4));
                x[0] += tscale * tq.x; x[1] += tscale * tq.y; x[2] += tscale * tq.z; x[3] += tscale * tq.w;
              }
              if (f_res) {
                const uint4 rq = *sw(s_res + hh * 4096, lane, g);
                x[0] += __uint_as_float(rq.x); x[1] += __uint_as_float(rq.y);
                x[2] += __uint_as_float(rq.z); x[3] += __uint_as_float(rq.w);
              }
              if (f_mask) {   // 8 mask halves per 16-byte chunk: columns hh * 32 + g * 4 .. + 3 sit in chunk hh * 4 + g / 2
                const uint4 mq = *sw(s_res, lane, hh * 4 + (g >> 1));
                const uint32_t m01 = (g & 1) ? mq.z : mq.x, m23 = (g & 1) ? mq.w : mq.y;
                x[0] = (m01 & 0x7fffu) ? x[0] : 0.f; x[1] = (m01 & 0x7fff0000u) ? x[1] : 0.f;
                x[2] = (m23 & 0x7fffu) ? x[2] : 0.f; x[3] = (m23 & 0x7fff0000u) ? x[3] : 0.f;
              }
            }
            if (f_f32)
              *sw(s_f32 + hh * 4096, lane, g) = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]),
                                                            __float_as_uint(x[2]), __float_as_uint(x[3]));
            acc[g * 4 + 0] = __float_as_uint(x[0]); acc[g * 4 + 1] = __float_as_uint(x[1]);
            acc[g * 4 + 2] = __float_as_uint(x[2]); acc[g * 4 + 3] = __float_as_uint(x[3]);
          }
          if (f_h) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {         // 8 columns -> one 16-byte fp16 chunk
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[g * 8 + e]);
              uint4 u;
              u.x = pack_half2(f[0], f[1]); u.y = pack_half2(f[2], f[3]);
              u.z = pack_half2(f[4], f[5]); u.w = pack_half2(f[6], f[7]);
              *sw(s_h, lane, hh * 4 + g) = u;
              if (f_lo) {
                float d[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) d[e] = f[e] - __half2float(__float2half_rn(f[e]));
                uint4 q;
                q.x = pack_half2(d[0], d[1]); q.y = pack_half2(d[2], d[3]);
                q.z = pack_half2(d[4], d[5]); q.w = pack_half2(d[6], d[7]);
                *sw(s_lo, lane, hh * 4 + g) = q;
              }
            }
          }
        }
        __syncwarp();
        if (f_f32) {
          store_f32_slab(s_f32, p.out_f32, p.ld_f32, n0, nvalid);
          store_f32_slab(s_f32 + 4096, p.out_f32, p.ld_f32, n0 + 32, nvalid - 32);
        }
        if (f_h) {
          store_f16_slab(s_h, p.out_h, p.ld_h, n0, nvalid);
          if (f_lo) store_f16_slab(s_lo, p.out_lo, p.ld_h, n0, nvalid);
        }
      }
    } else if constexpr (MODE == EPI_LN) {
      // 16 warps: warp (quad, grp) owns rows quad*32.. and the 32-column chunks {ci*128 + grp*32}.  The values stay in
      // registers between the statistics sweep and the normalisation (no TMEM write-back).
      uint8_t* l_f32 = smem_a + (warp - 2) * 6144;          // [32 rows x 128 B] fp32 / residual slab (swizzled)
      uint8_t* l_h = l_f32 + 4096;                          // [32 rows x 64 B] fp16 slab, chunk ^= (row >> 1) & 3
      float* lred = reinterpret_cast<float*>(smem_a + 16 * 6144);   // [2][4][128] row-statistic exchange
      auto ln_bar = []() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
      const int nbase = n_tile * BLOCK_N;                   // column offset of this CTA (cluster split of N)
      const float inv_n = 1.f / static_cast<float>(p.ln_cluster ? 2 * BLOCK_N : BLOCK_N);
      const uint32_t peer = cluster_ctarank() ^ 1u;
      float xs[NCH][32];
      float sum4[4] = {0.f, 0.f, 0.f, 0.f}, sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int col0 = ci * 128 + grp * 32;
        const int n0 = nbase + col0;
        __syncwarp();
        if (ci == 0) {
#pragma unroll
          for (int it = 0; it < 8; ++it) *sw(l_f32, it * 4 + (lane >> 3), lane & 7) = pre[it];
        } else {
          load_f32_slab(l_f32, p.residual, p.res_ld, n0, 32);
        }
        tmem_ld32(taddr + col0, v);
        tmem_wait_ld();
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + g * 4));
          const uint4 rq = *sw(l_f32, lane, g);
          const float x0 = __uint_as_float(v[g * 4 + 0]) + bq.x + __uint_as_float(rq.x);
          const float x1 = __uint_as_float(v[g * 4 + 1]) + bq.y + __uint_as_float(rq.y);
          const float x2 = __uint_as_float(v[g * 4 + 2]) + bq.z + __uint_as_float(rq.z);
          const float x3 = __uint_as_float(v[g * 4 + 3]) + bq.w + __uint_as_float(rq.w);
          sum4[0] += x0; sum4[1] += x1; sum4[2] += x2; sum4[3] += x3;
          sq4[0] = fmaf(x0, x0, sq4[0]); sq4[1] = fmaf(x1, x1, sq4[1]);
          sq4[2] = fmaf(x2, x2, sq4[2]); sq4[3] = fmaf(x3, x3, sq4[3]);
          xs[ci][g * 4 + 0] = x0; xs[ci][g * 4 + 1] = x1; xs[ci][g * 4 + 2] = x2; xs[ci][g * 4 + 3] = x3;
        }
      }
      float tot = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      float tsq = (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
      lred[grp * 128 + r] = tot;                     // combine the four column groups of every row
      lred[512 + grp * 128 + r] = tsq;
      ln_bar();
      tot = (lred[r] + lred[128 + r]) + (lred[256 + r] + lred[384 + r]);
      tsq = (lred[512 + r] + lred[640 + r]) + (lred[768 + r] + lred[896 + r]);
      if (p.ln_cluster) {                            // push this CTA's row statistics into the peer, wait for the peer's
        if (grp == 0) {
          st_cluster_f32(map_to_cta(smem_u32(&xred[r]), peer), tot);
          st_cluster_f32(map_to_cta(smem_u32(&xred[128 + r]), peer), tsq);
          mbar_arrive_cluster(map_to_cta(smem_u32(&xbar[0]), peer));
        }
        mbar_wait_cluster(&xbar[0], 0);
        tot += xred[r];
        tsq += xred[128 + r];
      }
      const float mean = tot * inv_n;
      const float rstd = rsqrtf(fmaxf(tsq * inv_n - mean * mean, 0.f) + p.ln_eps);
      if (p.ln_rstd && grp == 0 && n_tile == 0 && row_ok) p.ln_rstd[grow] = rstd;
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int n0 = nbase + ci * 128 + grp * 32;
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 gq = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + n0 + g * 4));
          const float4 bq = __ldg(reinterpret_cast<const float4*>(p.ln_beta + n0 + g * 4));
          const float y0 = (xs[ci][g * 4 + 0] - mean) * rstd * gq.x + bq.x;
          const float y1 = (xs[ci][g * 4 + 1] - mean) * rstd * gq.y + bq.y;
          const float y2 = (xs[ci][g * 4 + 2] - mean) * rstd * gq.z + bq.z;
          const float y3 = (xs[ci][g * 4 + 3] - mean) * rstd * gq.w + bq.w;
          *sw(l_f32, lane, g) = make_uint4(__float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2), __float_as_uint(y3));
          xs[ci][g * 4 + 0] = y0; xs[ci][g * 4 + 1] = y1; xs[ci][g * 4 + 2] = y2; xs[ci][g * 4 + 3] = y3;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_half2(xs[ci][g * 8 + 0], xs[ci][g * 8 + 1]);
          u.y = pack_half2(xs[ci][g * 8 + 2], xs[ci][g * 8 + 3]);
          u.z = pack_half2(xs[ci][g * 8 + 4], xs[ci][g * 8 + 5]);
          u.w = pack_half2(xs[ci][g * 8 + 6], xs[ci][g * 8 + 7]);
          *reinterpret_cast<uint4*>(l_h + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = u;
        }
        __syncwarp();
        store_f32_slab(l_f32, p.out_f32, p.ld_f32, n0, 32);
#pragma unroll
        for (int it = 0; it < 4; ++it) {               // 8 rows x 64 B per instruction
          const int row = it * 8 + (lane >> 2), chunk = lane & 3;
          if (row < rows_here)
            *reinterpret_cast<uint4*>(p.out_h + (grow0 + row) * p.ld_h + n0 + chunk * 8) =
                *reinterpret_cast<const uint4*>(l_h + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
        }
      }
    } else if constexpr (MODE == EPI_COUPLING) {
      // columns [0, hN) = log_scale, [hN, 2 hN) = shift, hN = N / 2 = 64 (modules/flow.py:223-257).
      // Each warp half owns 32 of the 64 transformed channels; z is staged through the swizzled slabs.
      const int hN = p.N >> 1;
      const int ch0 = half * 32;
      const int len_b = __ldg(p.lengths + min(sb, p.seq_B - 1));
      const bool in_len = row_ok && (st < len_b);
      __syncwarp();
      load_f32_slab(s_res, p.z, p.z_ld, p.zp_off + ch0, 32);
      tmem_ld32(taddr + ch0, v);
      tmem_ld32(taddr + hN + ch0, w);
      tmem_wait_ld();
      __syncwarp();
      float ld4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 bl = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + g * 4));
        const float4 bs = __ldg(reinterpret_cast<const float4*>(p.bias + hN + ch0 + g * 4));
        const uint4 zq = *sw(s_res, lane, g);
        const float zp[4] = {__uint_as_float(zq.x), __uint_as_float(zq.y), __uint_as_float(zq.z), __uint_as_float(zq.w)};
        const float lsb[4] = {bl.x, bl.y, bl.z, bl.w}, shb[4] = {bs.x, bs.y, bs.z, bs.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ls = __uint_as_float(v[g * 4 + e]) + lsb[e];
          const float sh = __uint_as_float(w[g * 4 + e]) + shb[e];
          const float scale = __fdividef(1.f, 1.f + __expf(-(ls + 2.0f)));   // sigmoid(ls + 2), flow.py:231
          o[e] = p.backward ? __fdividef(zp[e] - sh, scale + 1e-12f) : scale * zp[e] + sh;
          ld4[e] += __logf(scale);
          v[g * 4 + e] = __float_as_uint(o[e]);
          w[g * 4 + e] = __float_as_uint(scale);
        }
        if (p.save_scale && row_ok)
          *reinterpret_cast<uint4*>(p.save_scale + grow * hN + ch0 + g * 4) = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
        *sw(s_f32, lane, g) = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]),
                                         __float_as_uint(o[3]));
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_half2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
        u.y = pack_half2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
        u.z = pack_half2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
        u.w = pack_half2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
        *sw(s_h, lane, g) = u;
      }
      __syncwarp();
      store_f32_slab(s_f32, p.z, p.z_ld, p.zp_off + ch0, 32);
      store_f16_slab(s_h, p.z_h, p.z_ld, p.zp_off + ch0, 32);
      const float logdet = (ld4[0] + ld4[1]) + (ld4[2] + ld4[3]);
      red[half * 128 + r] = logdet;
      epi_bar();
      if (half == 0 && row_ok) {
        const float tot = logdet + red[128 + r];
        p.row_acc[grow] += in_len ? (p.backward ? -tot : tot) : 0.f;
      }
    } else if constexpr (MODE == EPI_POSTERIOR) {
      // modules/posterior.py:20-72 with the models.py:136 name swap already applied by the packing order:
      // columns [0, L) = log-variance (mu_projection), [L, 2L) = mean (logvar_projection), L = N / 2.
      const int L = p.N >> 1;
      const int len_b = __ldg(p.lengths + min(sb, p.seq_B - 1));
      const bool in_len = row_ok && (st < len_b);
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = half; c < L / 32; c += 2) {
        const int ch0 = c * 32;
        __syncwarp();
        load_f32_slab(s_res, p.eps_in, p.z_ld, ch0, 32);
        tmem_ld32(taddr + ch0, v);
        tmem_ld32(taddr + L + ch0, w);
        tmem_wait_ld();
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 bl = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + g * 4));
          const float4 bm = __ldg(reinterpret_cast<const float4*>(p.bias + L + ch0 + g * 4));
          const uint4 eq = *sw(s_res, lane, g);
          const float ee[4] = {__uint_as_float(eq.x), __uint_as_float(eq.y), __uint_as_float(eq.z), __uint_as_float(eq.w)};
          const float lvb[4] = {bl.x, bl.y, bl.z, bl.w}, mub[4] = {bm.x, bm.y, bm.z, bm.w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lv = __uint_as_float(v[g * 4 + e]) + lvb[e];
            const float mu = __uint_as_float(w[g * 4 + e]) + mub[e];
            o[e] = ee[e] * expf(0.5f * lv) + mu;
            a4[e] += lv + ee[e] * ee[e];
            v[g * 4 + e] = __float_as_uint(o[e]);
            w[g * 4 + e] = __float_as_uint(lv);
          }
          if (p.save_lv && row_ok)
            *reinterpret_cast<uint4*>(p.save_lv + grow * L + ch0 + g * 4) = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
          *sw(s_f32, lane, g) = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]),
                                           __float_as_uint(o[3]));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_half2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
          u.y = pack_half2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
          u.z = pack_half2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
          u.w = pack_half2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
          *sw(s_h, lane, g) = u;
        }
        __syncwarp();
        store_f32_slab(s_f32, p.z, p.z_ld, ch0, 32);
        store_f16_slab(s_h, p.z_h, p.z_ld, ch0, 32);
      }
      const float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
      red[half * 128 + r] = acc;
      epi_bar();
      if (half == 0 && row_ok) {
        const float tot = acc + red[128 + r];
        p.row_acc[grow] += in_len ? -0.5f * (static_cast<float>(L) * 1.8378770664093453f + tot) : 0.f;
      }
    }
    tc_fence_before();
    if (dbg && threadIdx.x == 64) dbg[3] = gtime();
  }
  __syncthreads();
  if (dbg && threadIdx.x == 64) dbg[4] = gtime();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (p.ln_cluster) cluster_sync_relaxed();   // no CTA of the pair exits while the other may still address its shared memory
}

}  // namespace vb

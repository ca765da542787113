// Weight-gradient GEMM for sm_100a (tcgen05 + TMEM + TMA):
//
//   dW[M, N] += sum_tokens  X[token (+shift), m]^T * dY[token, n]          (Keras layout [in, out], fp32)
//
// Both operands are the ROW-MAJOR fp16 activation / gradient tensors [batch, time, channel] exactly as the
// forward / backward kernels leave them in HBM: the contraction runs over tokens, so the tiles are fed to the
// tensor core as MN-major operands (channel contiguous) -- no transposed copies are ever materialised.
//   * TMA box {64 channels, 64 tokens} with the 128-byte swizzle = the canonical MN-major SW128 atom
//     ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units: 8 token rows of 128 B form one 1024-byte atom,
//     SBO = 1024 B between 8-token groups, LBO = 8 KB between 64-channel chunks (one TMA box each).
//   * Conv1D taps (modules/utils.py:56-85): the A rows are shifted by (tap - (k-1)/2); the 3-D tensor map
//     zero-fills rows outside [0, T) of every utterance, i.e. the 'same' padding.
//   * Dense over a concat [x ; ctx] (modules/attention.py:410,440,447): output rows < a_split come from map A0,
//     the rest from map A1.
//   * Split-K over token blocks across blockIdx.z; partial tiles are accumulated into the flat gradient
//     buffer with coalesced fp32 reductions (red.global.add.f32).
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quadrant, each
// taking half of the tile's columns; 16-byte vector reductions).  The epilogue was 30-50 % of a CTA's life with four warps
// and scalar reductions (ncu source view of the round-1 capture): a split-K CTA only has 12-24 token blocks of mainloop.
#pragma once
#include "ptx.cuh"

namespace vb {

constexpr int WG_BLOCK_M = 128;   // dW rows per CTA (input channels)
constexpr int WG_BLOCK_K = 64;    // tokens per pipeline stage
constexpr int WG_THREADS = 320;
constexpr int WG_TILE_LD = 36;    // floats per row of the epilogue staging tile (16-byte aligned rows, conflict-free float4 access)
constexpr int WG_CHUNK_BYTES = 64 * WG_BLOCK_K * 2;             // one TMA box: 64 channels x 64 tokens fp16 = 8 KB
// BN = dW columns per CTA (output channels): 128, or 256 for wide outputs -- a 128 x 256 tile moves 48 KB of operands per
// 128 x 256 x 64 MMA block instead of 2 x 32 KB (the mainloop is bound by the L2 -> shared-memory fill of one SM, not by
// the tensor pipe: 25 % tensor-active at 128 x 128, profiles/wgrad_r1.md)
template <int BN>
struct WgCfg {
  static constexpr int kStages = BN == 128 ? 6 : 4;
  static constexpr int kStageBytes = (2 + BN / 64) * WG_CHUNK_BYTES;   // A (2 chunks) + B (BN / 64 chunks)
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
};

struct WgradParams {
  int batches, rows;       // token geometry [batches][rows]
  int kb_per_batch;        // ceil(rows / 64)
  int total_kb;            // batches * kb_per_batch
  int kb_per_split;        // token blocks per blockIdx.z
  int M, N;                // dW extent
  int a_split;             // dW rows >= a_split read map A1 (channel m - a_split); M if unused
  int a_shift;             // token shift of the A rows (Conv1D tap)
  int a0_col0, a1_col0, b_col0;   // first channel of the operand windows inside their tensors
  float* out;              // [M, ldo] fp32, accumulated
  int ldo;
  // several parameter tensors behind one launch (Q|K|V kernels, the K / V memory projections of all blocks of a module):
  // column n belongs to outs[n / n_per_out] at local column n % n_per_out (n_per_out == 0: single output `out`)
  int n_per_out;
  float* outs[24];
};

// smem descriptor of an MN-major SW128 tile: start, LBO (64-channel chunk pitch), SBO (8-token group pitch)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// 16-byte vector reduction into global memory (sm_90+): four fp32 adds, no return value
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// kind::f16 instruction descriptor with selectable operand majors (bit 15: A MN-major, bit 16: B MN-major)
__host__ __device__ constexpr uint32_t umma_idesc_f16_major(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  constexpr int WG_STAGES = WgCfg<BN>::kStages, WG_STAGE_BYTES = WgCfg<BN>::kStageBytes, WG_BLOCK_N = BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * WG_BLOCK_M;
  const int n0 = blockIdx.y * WG_BLOCK_N;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
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
  if (warp == 1) tmem_alloc<WG_BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (nkb > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        const bool second = m0 >= p.a_split;
        const CUtensorMap* tmA = second ? &tmA1 : &tmA0;
        const int a_col = second ? p.a1_col0 + (m0 - p.a_split) : p.a0_col0 + m0;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int b = kb / p.kb_per_batch;
          const int t0 = (kb - b * p.kb_per_batch) * WG_BLOCK_K;
          uint8_t* st = smem + stage * WG_STAGE_BYTES;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
          tma_load_3d(st, tmA, &full_bar[stage], a_col, t0 + p.a_shift, b);
          tma_load_3d(st + WG_CHUNK_BYTES, tmA, &full_bar[stage], a_col + 64, t0 + p.a_shift, b);
#pragma unroll
          for (int cb = 0; cb < WG_BLOCK_N / 64; ++cb)
            tma_load_3d(st + (2 + cb) * WG_CHUNK_BYTES, &tmB, &full_bar[stage], p.b_col0 + n0 + cb * 64, t0, b);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_f16_major(WG_BLOCK_M, WG_BLOCK_N, true, true);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
          const uint32_t sb = sa + 2 * WG_CHUNK_BYTES;
#pragma unroll
          for (int k = 0; k < WG_BLOCK_K / 16; ++k) {
            // 16 tokens = two 8-token groups = 2048 bytes further into every 64-channel chunk
            const uint64_t adesc = umma_desc_mn_sw128(sa + k * 2048, WG_CHUNK_BYTES, 1024);
            const uint64_t bdesc = umma_desc_mn_sw128(sb + k * 2048, WG_CHUNK_BYTES, 1024);
            umma_f16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      // ===================== epilogue: TMEM -> smem transpose -> coalesced fp32 reductions =====================
      const int quad = warp & 3;               // TMEM lane quadrant (hardware rule: warp id % 4)
      const int half = (warp - 2) >> 2;        // column half of the tile this warp reduces
      float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * WG_TILE_LD);   // pipeline stages are idle by now
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      uint32_t v[32];
      // 16-byte reductions need 4-column groups inside one output tensor and 16-byte aligned rows
      const bool vec_ok = (p.ldo % 4 == 0) && (p.N % 4 == 0) && (p.n_per_out % 4 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.n_per_out ? p.outs[0] : p.out) & 15) == 0);
      for (int c0 = half * (WG_BLOCK_N / 2); c0 < (half + 1) * (WG_BLOCK_N / 2); c0 += 32) {
        if (n0 + c0 >= p.N) break;
        __syncwarp();
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)   // lane == row of the tile: eight 16-byte stores, rows 144 B apart
          *reinterpret_cast<uint4*>(tile + lane * WG_TILE_LD + j * 4) = make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
        __syncwarp();
        if (vec_ok) {
          // lane -> (row it * 4 + lane / 8, columns (lane % 8) * 4 .. + 3): one instruction reduces 4 rows x 128 B
          const int col = n0 + c0 + (lane & 7) * 4;
          if (col < p.N) {
            float* dst = p.n_per_out ? p.outs[col / p.n_per_out] + (col % p.n_per_out) : p.out + col;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + (lane >> 3);
              const int row = m0 + quad * 32 + rr;
              if (row < p.M) {
                const float4 q = *reinterpret_cast<const float4*>(tile + rr * WG_TILE_LD + (lane & 7) * 4);
                red_add_v4(dst + static_cast<long>(row) * p.ldo, q);
              }
            }
          }
        } else {
          const int col = n0 + c0 + lane;
          if (col < p.N) {
            float* dst = p.n_per_out ? p.outs[col / p.n_per_out] + (col % p.n_per_out) : p.out + col;
            for (int rr = 0; rr < 32; ++rr) {
              const int row = m0 + quad * 32 + rr;
              if (row < p.M) atomicAdd(dst + static_cast<long>(row) * p.ldo, tile[rr * WG_TILE_LD + lane]);
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<WG_BLOCK_N>(tmem_base);
  }
}

}  // namespace vb

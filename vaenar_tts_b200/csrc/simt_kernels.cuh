// Non-tensor-core kernels of the path: weight packing, embedding gather, positional-encoding table,
// the fp32 flow linear algebra (ActNorm (+) InvertibleLinear, 128x128 LU / inverse), reductions for
// the log-densities and losses, and a counter-based normal RNG.  All fp32 (fp64 where the reference
// uses float64, modules/flow.py:126-129,141-144).
#pragma once
#include "ptx.cuh"

namespace vb {

// ------------------------------------------------------------------ weight packing
// dst[n * ldd + k] = fp16( src[k * lds + n] )          (mode 0: hi part)
//                  = fp16( src - float(fp16(src)) )    (mode 1: lo part of the split-fp16 pair)
// dst[k * ldd + n] = fp16( src[k * lds + n] )          (mode 2: straight copy -- the backward pass (dgrad) wants the
//                                                       Keras layout itself: [in, out] = [N_gemm, K_gemm] K-major)
// Keras Dense kernels are [in, out]; the UMMA B operand wants [out, in] (K-major).
struct PackOp {
  const float* src;
  __half* dst;
  int K, N, lds, ldd, mode;
};
__global__ void pack_weights_kernel(const PackOp* __restrict__ ops) {
  const PackOp op = ops[blockIdx.z];
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  if (k0 >= op.K || n0 >= op.N) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < op.K && n < op.N) ? op.src[static_cast<long>(k) * op.lds + n] : 0.f;
  }
  __syncthreads();
  if (op.mode == 2) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int k = k0 + i, n = n0 + threadIdx.x;
      if (k < op.K && n < op.N) op.dst[static_cast<long>(k) * op.ldd + n] = __float2half_rn(tile[i][threadIdx.x]);
    }
    return;
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < op.N && k < op.K) {
      const float v = tile[threadIdx.x][i];
      const __half hi = __float2half_rn(v);
      op.dst[static_cast<long>(n) * op.ldd + k] = op.mode ? __float2half_rn(v - __half2float(hi)) : hi;
    }
  }
}

// Same, one CTA (256 threads) per 64x64 tile of the whole plan: tile_start[op] = first CTA of op (prefix sums, nops + 1
// entries).  Reads are float4 along N, writes half2 along K (mode 0/1) or along N (mode 2).
constexpr int PACK_TILE = 64;
__global__ void __launch_bounds__(256)
pack_weights_flat_kernel(const PackOp* __restrict__ ops, const int* __restrict__ tile_start, int nops) {
  int lo = 0, hi = nops - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_start[mid] <= b) lo = mid; else hi = mid - 1;
  }
  const PackOp op = ops[lo];
  const int t = b - tile_start[lo];
  const int tn = (op.N + PACK_TILE - 1) / PACK_TILE;
  const int n0 = (t % tn) * PACK_TILE, k0 = (t / tn) * PACK_TILE;
  __shared__ float tile[PACK_TILE][PACK_TILE + 1];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  for (int i = ty; i < PACK_TILE; i += 4) {
    const int k = k0 + i, n = n0 + tx;
    tile[i][tx] = (k < op.K && n < op.N) ? op.src[static_cast<long>(k) * op.lds + n] : 0.f;
  }
  __syncthreads();
  if (op.mode == 2) {
    for (int i = ty; i < PACK_TILE; i += 4) {
      const int k = k0 + i, n = n0 + tx;
      if (k < op.K && n < op.N) op.dst[static_cast<long>(k) * op.ldd + n] = __float2half_rn(tile[i][tx]);
    }
    return;
  }
  for (int i = ty; i < PACK_TILE; i += 4) {
    const int n = n0 + i, k = k0 + tx;
    if (n < op.N && k < op.K) {
      const float v = tile[tx][i];
      const __half hi16 = __float2half_rn(v);
      op.dst[static_cast<long>(n) * op.ldd + k] = op.mode ? __float2half_rn(v - __half2float(hi16)) : hi16;
    }
  }
}

// Inference-mode BatchNorm folded to a per-channel affine (modules/utils.py:72; Keras eps 1e-3):
//   y = (x - mean) * rsqrt(var + eps) * gamma + beta  =  x * scale + shift
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = gamma[c] * rsqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
  }
}

__global__ void concat2_kernel(const float* a, int na, const float* b, int nb, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) out[i] = a[i];
  else if (i < na + nb) out[i] = b[i - na];
}

// ------------------------------------------------------------------ elementwise helpers
__global__ void cast_f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long n) {
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 u;
    u.x = pack_half2(v.x, v.y);
    u.y = pack_half2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (long j = i; j < n; ++j) out[j] = __float2half_rn(in[j]);
  }
}

// Embedding lookup (modules/encoder.py:10-12,81): rows of the table as fp16 conv operands.
__global__ void embed_kernel(const int* __restrict__ ids, const float* __restrict__ table, __half* __restrict__ out,
                             int n_tokens, int dim, int vocab) {
  const int tok = blockIdx.x;
  if (tok >= n_tokens) return;
  int id = ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int d = threadIdx.x; d < dim; d += blockDim.x)
    out[static_cast<long>(tok) * dim + d] = __float2half_rn(table[static_cast<long>(id) * dim + d]);
}

// PositionalEncoding.positional_encoding (modules/utils.py:332-355):
//   even d: sin(t*step / 10000^(d/D)) ; odd d: cos(t*step / 10000^((d-1)/D))
__global__ void pe_table_kernel(float* __restrict__ out, int T, int D, float step) {
  const int t = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float pos = static_cast<float>(t) * step;
    float v;
    if ((d & 1) == 0) v = sinf(pos / powf(10000.f, static_cast<float>(d) / static_cast<float>(D)));
    else v = cosf(pos / powf(10000.f, static_cast<float>(d - 1) / static_cast<float>(D)));
    out[static_cast<long>(t) * D + d] = v;
  }
}

// mels[:, ::rf, :] (models/models.py:123) as fp16 GEMM operand [B, Tz, 80]
__global__ void reduce_mels_kernel(const float* __restrict__ mels, __half* __restrict__ out, int B, int Tm, int Tz,
                                   int rf, int C) {
  const int row = blockIdx.x;           // b * Tz + tz
  const int b = row / Tz, tz = row % Tz;
  const int tm = tz * rf;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = tm < Tm ? mels[(static_cast<long>(b) * Tm + tm) * C + c] : 0.f;
    out[static_cast<long>(row) * C + c] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------ block reduce
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  return sh[0];
}

// ------------------------------------------------------------------ length predictor
// DenseLengthPredictor.call (modules/length_predictor.py:35-42): sum_t mask * exp(x_t . w + b)
__global__ void length_predictor_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                        const float* __restrict__ bias, const int* __restrict__ lens,
                                        float* __restrict__ out, int T, int D) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int len = lens[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int t = warp; t < T && t < len; t += nw) {
    const float* row = x + (static_cast<long>(b) * T + t) * D;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot += row[d] * w[d];
    dot = warp_sum(dot);
    if (lane == 0) acc += expf(dot + bias[0]);
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) out[b] = tot;
}

// ------------------------------------------------------------------ flow linear algebra (DIM = 128)
constexpr int FLOW_DIM = 128;

// Partial-pivot search of column k over rows k..127 by warp 0 (first maximum wins, like the serial scan).
template <typename T>
__device__ __forceinline__ void pivot_search(const T* A, int ld, int k, int* piv, T* pivval) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    T bv = static_cast<T>(-1);
    int best = FLOW_DIM;
    for (int i = k + lane; i < FLOW_DIM; i += 32) {
      const T v = A[i * ld + k] < static_cast<T>(0) ? -A[i * ld + k] : A[i * ld + k];
      if (v > bv) { bv = v; best = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best, o);
      if (ov > bv || (ov == bv && oi < best)) { bv = ov; best = oi; }
    }
    if (lane == 0) {
      *piv = best;
      if (pivval) *pivval = bv;
    }
  }
}

// log|det W| in float64 with partial pivoting (modules/flow.py:126-129: slogdet(cast(W, float64))).
// One CTA of 128 threads per matrix; thread j owns column j.
__global__ void slogdet128_kernel(const float* const* __restrict__ Ws, double* __restrict__ out) {
  extern __shared__ double A[];                    // [128][129]
  __shared__ int piv;
  __shared__ double pivval;
  const float* W = Ws[blockIdx.x];
  const int j = threadIdx.x;
  constexpr int LD = FLOW_DIM + 1;
  for (int i = 0; i < FLOW_DIM; ++i) A[i * LD + j] = static_cast<double>(W[i * FLOW_DIM + j]);
  __syncthreads();
  double logdet = 0.0;
  for (int k = 0; k < FLOW_DIM; ++k) {
    pivot_search<double>(A, LD, k, &piv, &pivval);
    __syncthreads();
    const int pr = piv;
    if (pr != k) {
      const double tmp = A[k * LD + j];
      A[k * LD + j] = A[pr * LD + j];
      A[pr * LD + j] = tmp;
    }
    __syncthreads();
    logdet += log(pivval);
    const double pinv = 1.0 / A[k * LD + k];
    const double akj = A[k * LD + j];
    __syncthreads();
    if (j > k) {
      for (int i = k + 1; i < FLOW_DIM; ++i) A[i * LD + j] -= A[i * LD + k] * pinv * akj;
    }
    __syncthreads();
  }
  if (j == 0) out[blockIdx.x] = logdet;
}

// fp32 inverse by Gauss-Jordan with partial pivoting (modules/flow.py:139: tf.linalg.inv(self.weight)).
// One CTA of 256 threads per matrix; thread j owns column j of the augmented [128 x 256] system.
__global__ void inverse128_kernel(const float* const* __restrict__ Ws, float* __restrict__ out) {
  extern __shared__ float G[];                     // [128][257] + column buffer [128]
  __shared__ int piv;
  const float* W = Ws[blockIdx.x];
  float* Winv = out + static_cast<long>(blockIdx.x) * FLOW_DIM * FLOW_DIM;
  const int j = threadIdx.x;                       // 0..255
  constexpr int LD = 2 * FLOW_DIM + 1;
  float* colk = G + FLOW_DIM * LD;
  for (int i = 0; i < FLOW_DIM; ++i)
    G[i * LD + j] = j < FLOW_DIM ? W[i * FLOW_DIM + j] : ((j - FLOW_DIM) == i ? 1.f : 0.f);
  __syncthreads();
  for (int k = 0; k < FLOW_DIM; ++k) {
    pivot_search<float>(G, LD, k, &piv, static_cast<float*>(nullptr));
    __syncthreads();
    const int pr = piv;
    if (pr != k) {
      const float tmp = G[k * LD + j];
      G[k * LD + j] = G[pr * LD + j];
      G[pr * LD + j] = tmp;
    }
    __syncthreads();
    const float pinv = 1.0f / G[k * LD + k];
    if (j < FLOW_DIM) colk[j] = G[j * LD + k];
    __syncthreads();
    const float rowk = G[k * LD + j] * pinv;
    G[k * LD + j] = rowk;
    for (int i = 0; i < FLOW_DIM; ++i)
      if (i != k) G[i * LD + j] -= colk[i] * rowk;
    __syncthreads();
  }
  if (j >= FLOW_DIM)
    for (int i = 0; i < FLOW_DIM; ++i) Winv[i * FLOW_DIM + (j - FLOW_DIM)] = G[i * LD + j];
}

// Fold ActNorm into the invertible linear map, both directions (modules/flow.py:123-187):
//   forward  (sample):        y = (z * e^s + b) W            = z Mf + cf ,  Mf = diag(e^s) W , cf = b W
//   backward (log_prob):      y = (z W^-1 - b) / (e^s + 1e-8) = z Mb + cb ,  Mb = W^-1 diag(r), cb = -b r
// consts[step] = { sum(log_scale), log|det W| }.
__global__ void flow_fold_kernel(const float* const* __restrict__ Ws, const float* const* __restrict__ log_scales,
                                 const float* const* __restrict__ biases, const float* __restrict__ Winv,
                                 const double* __restrict__ logdet64, float* __restrict__ Mf, float* __restrict__ cf,
                                 float* __restrict__ Mb, float* __restrict__ cb, float* __restrict__ consts) {
  __shared__ float sh[32];
  const int s = blockIdx.x;
  const float* W = Ws[s];
  const float* ls = log_scales[s];
  const float* bi = biases[s];
  const float* Wi = Winv + static_cast<long>(s) * FLOW_DIM * FLOW_DIM;
  const int j = threadIdx.x;   // 128 threads, column j
  float c = 0.f;
  for (int i = 0; i < FLOW_DIM; ++i) {
    const float w = W[i * FLOW_DIM + j];
    Mf[(static_cast<long>(s) * FLOW_DIM + i) * FLOW_DIM + j] = expf(ls[i]) * w;
    c += bi[i] * w;
  }
  cf[s * FLOW_DIM + j] = c;
  const float r = 1.0f / (expf(ls[j]) + 1e-8f);
  for (int i = 0; i < FLOW_DIM; ++i) Mb[(static_cast<long>(s) * FLOW_DIM + i) * FLOW_DIM + j] = Wi[i * FLOW_DIM + j] * r;
  cb[s * FLOW_DIM + j] = -bi[j] * r;
  const float tot = block_sum(ls[j], sh);
  if (j == 0) {
    consts[2 * s] = tot;
    consts[2 * s + 1] = static_cast<float>(logdet64[s]);
  }
}

// Folded flow maps M[s][k][n] (fp32, z M convention) -> split-fp16 UMMA B operands out[s][0:128][k] = hi(M[k][n]) (row n),
// out[s][128:256][k] = lo(M[k][n]): hi + lo carries 22 bits, z (hi, lo) x M (hi, lo) minus the lo x lo term ~ fp32 accuracy.
__global__ void flow_pack_f16_kernel(const float* __restrict__ M, __half* __restrict__ out, int S) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(S) * FLOW_DIM * FLOW_DIM) return;
  const int s = static_cast<int>(i / (FLOW_DIM * FLOW_DIM));
  const int k = static_cast<int>((i / FLOW_DIM) % FLOW_DIM), n = static_cast<int>(i % FLOW_DIM);
  const float m = M[i];
  const __half hi = __float2half_rn(m);
  __half* o = out + static_cast<long>(s) * 2 * FLOW_DIM * FLOW_DIM;
  o[static_cast<long>(n) * FLOW_DIM + k] = hi;
  o[static_cast<long>(FLOW_DIM + n) * FLOW_DIM + k] = __float2half_rn(m - __half2float(hi));
}

// y[rows,128] = x[rows,128] M[128,128] + c  in fp32 (CUDA cores; exactness matters more than speed here:
// the flow state z carries the log-density).  In place is safe: a CTA stages its 64 rows first.
// Also emits the fp16 copy consumed by the conditioner's pre-projection GEMM.
// Thread (ty, tx) owns rows ty*8..+7 and columns tx, tx+32, tx+64, tx+96; the k loop runs in ascending order with one
// fmaf per term (the summation order every parity fixture was generated with).  transpose != 0 (backward pass:
// g_in = g_out M^T) keeps M row-major in shared memory with a row pitch of 129 floats so that the column reads are
// conflict-free.
constexpr int FLOW_ROWS = 64;
constexpr int FLOW_LIN_SMEM = (FLOW_DIM * (FLOW_DIM + 1) + FLOW_ROWS * FLOW_DIM) * 4;
__global__ void __launch_bounds__(256)
flow_linear_kernel(const float* z_in, float* z, __half* __restrict__ z_h, const float* __restrict__ M,
                   const float* __restrict__ c, long rows, int transpose) {
  extern __shared__ float smf[];
  float* Ms = smf;                               // [128][128] or [128][129]
  float* Xs = smf + FLOW_DIM * (FLOW_DIM + 1);   // [64][128]
  const long r0 = static_cast<long>(blockIdx.x) * FLOW_ROWS;
  if (transpose) {
    for (int i = threadIdx.x; i < FLOW_DIM * FLOW_DIM; i += 256) Ms[(i / FLOW_DIM) * (FLOW_DIM + 1) + (i % FLOW_DIM)] = M[i];
  } else {
    for (int i = threadIdx.x; i < FLOW_DIM * FLOW_DIM / 4; i += 256)
      reinterpret_cast<float4*>(Ms)[i] = reinterpret_cast<const float4*>(M)[i];
  }
  for (int i = threadIdx.x; i < FLOW_ROWS * FLOW_DIM / 4; i += 256) {
    const long row = r0 + (i * 4) / FLOW_DIM;
    reinterpret_cast<float4*>(Xs)[i] =
        row < rows ? reinterpret_cast<const float4*>(z_in + r0 * FLOW_DIM)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int ks = transpose ? 1 : FLOW_DIM;             // stride of k
  const int ns = transpose ? (FLOW_DIM + 1) * 32 : 32; // stride of the thread's 4 columns
  const float* mp = Ms + (transpose ? tx * (FLOW_DIM + 1) : tx);
  const float* xp = Xs + ty * 8 * FLOW_DIM;
#pragma unroll 4
  for (int k = 0; k < FLOW_DIM; ++k) {
    const float m0 = mp[k * ks], m1 = mp[k * ks + ns], m2 = mp[k * ks + 2 * ns], m3 = mp[k * ks + 3 * ns];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float x = xp[a * FLOW_DIM + k];
      acc[a][0] = fmaf(x, m0, acc[a][0]);
      acc[a][1] = fmaf(x, m1, acc[a][1]);
      acc[a][2] = fmaf(x, m2, acc[a][2]);
      acc[a][3] = fmaf(x, m3, acc[a][3]);
    }
  }
  float cc[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) cc[b] = c ? c[tx + 32 * b] : 0.f;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const long row = r0 + ty * 8 + a;
    if (row < rows) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float o = acc[a][b] + cc[b];
        z[row * FLOW_DIM + tx + 32 * b] = o;
        if (z_h) z_h[row * FLOW_DIM + tx + 32 * b] = __float2half_rn(o);
      }
    }
  }
}

// Per-batch Gaussian base log-density: sum_{t < len} sum_d -0.5 (ln 2pi + e^2)   (modules/prior.py:37-41,147-151)
__global__ void base_logprob_kernel(const float* __restrict__ e, const int* __restrict__ lens, float* __restrict__ out,
                                    int T, int D) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int len = min(lens[b], T);
  const long n = static_cast<long>(len) * D;
  const float* base = e + static_cast<long>(b) * T * D;
  float acc = 0.f;
  for (long i = static_cast<long>(blockIdx.y) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.y) * blockDim.x) {
    const float v = base[i];
    acc += -0.5f * (1.8378770664093453f + v * v);
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    if (gridDim.y == 1) out[b] = tot;
    else atomicAdd(out + b, tot);   // out zeroed by the caller
  }
}

// logp[b] = base[b] + sign * ( sum_rows row_acc  +  len_b * sum_steps (sum_s + logdetW) )
//   sample       (modules/prior.py:154-169): sign = -1, row_acc = sum of coupling forward log-dets
//   log_prob     (modules/prior.py:119-152): row_acc already holds the (negative) backward coupling log-dets,
//                                            the linear/actnorm terms enter with sign -1 as well.
__global__ void prior_logp_finalize_kernel(const float* __restrict__ base, const float* __restrict__ row_acc,
                                           const float* __restrict__ consts, int nsteps,
                                           const int* __restrict__ lens, float* __restrict__ out, int T,
                                           float row_sign) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) acc += row_acc[static_cast<long>(b) * T + t];
  const float rows = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    float lin = 0.f;
    for (int s = 0; s < nsteps; ++s) lin += consts[2 * s] + consts[2 * s + 1];
    out[b] = base[b] + row_sign * rows - static_cast<float>(lens[b]) * lin;
  }
}

// Per-batch sum of a per-row accumulator (posterior log q, modules/posterior.py:62-71)
__global__ void row_sum_kernel(const float* __restrict__ row_acc, float* __restrict__ out, int T) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) acc += row_acc[static_cast<long>(b) * T + t];
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) out[b] = tot;
}

// VAENAR._compute_l2_loss per batch element (models/models.py:67-86, n_sample = 1):
//   out[b] (+)= sum_{t < len} mean_d (rec - tgt)^2 / len
// grid (B, chunks): every CTA reduces a slice of the frames and adds its share (out must be zeroed by the caller).
__global__ void l2_loss_kernel(const float* __restrict__ rec, int rec_T, const float* __restrict__ tgt, int tgt_T,
                               const int* __restrict__ lens, float* __restrict__ out, int D) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int len = min(lens[b], tgt_T);
  const long n = static_cast<long>(len) * D;
  float acc = 0.f;
  for (long i = static_cast<long>(blockIdx.y) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.y) * blockDim.x) {
    const int t = static_cast<int>(i / D), d = static_cast<int>(i % D);
    const float df = rec[(static_cast<long>(b) * rec_T + t) * D + d] - tgt[(static_cast<long>(b) * tgt_T + t) * D + d];
    acc += df * df;
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out + b, tot / static_cast<float>(D) / static_cast<float>(lens[b]));
}

// ------------------------------------------------------------------ training-mode BatchNorm / dropout / ActNorm init
// Per-channel partial sums over a tile of rows: partial[tile][0][c] = sum x, partial[tile][1][c] = sum x^2.
// (Keras BatchNormalization in training mode normalises with the batch mean / population variance over
// (batch, time) INCLUDING padded frames, modules/utils.py:72,79-83.)
// 256 threads = row lanes x column groups of 4 (float4 loads); lanes combined in a fixed order.  C % 4 == 0, C <= 1024.
__global__ void __launch_bounds__(256)
colstats_partial_kernel(const float* __restrict__ x, long rows, int C, int rows_per_tile, float* __restrict__ partial) {
  __shared__ float red[256 * 4];
  const int groups = C / 4, lanes = 256 / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const bool active = lane < lanes;
  const long r0 = static_cast<long>(blockIdx.x) * rows_per_tile;
  const long r1 = min(rows, r0 + rows_per_tile);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (active)
    for (long r = r0 + lane; r < r1; r += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * C + g * 4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
    if (active)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[lane * C + g * 4 + e] = which ? q[e] : s[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += red[l * C + c];
      partial[(static_cast<long>(blockIdx.x) * 2 + which) * C + c] = t;
    }
  }
}
// mean / population variance (float64 accumulation over the tile partials, fixed order => deterministic),
// folded affine for the apply kernel, and the moving-average update  m <- m * momentum + batch * (1 - momentum).
__global__ void bn_train_finalize_kernel(const float* __restrict__ partial, int ntile, long rows, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ moving_mean, float* __restrict__ moving_var, float momentum,
                                         float eps, float* __restrict__ scale, float* __restrict__ shift, int update,
                                         float* __restrict__ save_mean = nullptr, float* __restrict__ save_rstd = nullptr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int t = 0; t < ntile; ++t) {
    s += partial[(static_cast<long>(t) * 2 + 0) * C + c];
    q += partial[(static_cast<long>(t) * 2 + 1) * C + c];
  }
  const double mean = s / static_cast<double>(rows);
  const double var = fmax(q / static_cast<double>(rows) - mean * mean, 0.0);
  const float sc = gamma[c] * rsqrtf(static_cast<float>(var) + eps);
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  if (save_mean) {
    save_mean[c] = static_cast<float>(mean);
    save_rstd[c] = rsqrtf(static_cast<float>(var) + eps);
  }
  if (update && isfinite(mean) && isfinite(var)) {   // an overflowed forward (fp16 operand range) must not poison the moving statistics
    moving_mean[c] = moving_mean[c] * momentum + static_cast<float>(mean) * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + static_cast<float>(var) * (1.f - momentum);
  }
}
// y * scale[c] + shift[c], times the dropout mask (values 0 or 1/(1-rate)); fp16 (+ split lo) operand for the next conv
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ mask,
                                __half* __restrict__ out_h, __half* __restrict__ out_lo, long n, int C) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  float v = y[i] * scale[c] + shift[c];
  if (mask) v *= mask[i];
  const __half h = __float2half_rn(v);
  out_h[i] = h;
  if (out_lo) out_lo[i] = __float2half_rn(v - __half2float(h));
}
// x * mask -> fp32 (optional, may alias x) and fp16 copies
__global__ void dropout_apply_kernel(const float* x, const float* __restrict__ mask, float* out_f32,
                                     __half* __restrict__ out_h, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i] * mask[i];
  if (out_f32) out_f32[i] = v;
  if (out_h) out_h[i] = __float2half_rn(v);
}
// ActNormFlow.init (modules/flow.py:189-196): data-dependent log_scale = log(1/(std+1e-8)), bias = -mean/(std+1e-8)
// from the population statistics over ALL positions, written into the parameter buffer, plus the folded forward map
// of this step (Mf = diag(e^s) W, cf = b W) so that the flow can continue immediately.  One CTA of 128 threads.
__global__ void actnorm_init_kernel(const float* __restrict__ partial, int ntile, long rows, float* __restrict__ log_scale,
                                    float* __restrict__ bias, const float* __restrict__ W, float* __restrict__ Mf,
                                    float* __restrict__ cf) {
  __shared__ float s_es[FLOW_DIM], s_b[FLOW_DIM];
  const int j = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int t = 0; t < ntile; ++t) {
    s += partial[(static_cast<long>(t) * 2 + 0) * FLOW_DIM + j];
    q += partial[(static_cast<long>(t) * 2 + 1) * FLOW_DIM + j];
  }
  const double mean = s / static_cast<double>(rows);
  const float sd = sqrtf(static_cast<float>(fmax(q / static_cast<double>(rows) - mean * mean, 0.0)));
  const float ls = logf(1.0f / (sd + 1e-8f));
  const float b = -static_cast<float>(mean) / (sd + 1e-8f);
  log_scale[j] = ls;
  bias[j] = b;
  s_es[j] = expf(ls);
  s_b[j] = b;
  __syncthreads();
  float c = 0.f;
  for (int i = 0; i < FLOW_DIM; ++i) {
    const float w = W[i * FLOW_DIM + j];
    Mf[i * FLOW_DIM + j] = s_es[i] * w;
    c += s_b[i] * w;
  }
  cf[j] = c;
}

// ------------------------------------------------------------------ counter-based N(0,1) generator
// Philox4x32-10 + Box-Muller; element i of a stream is a pure function of (seed, stream, i).
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__global__ void randn_kernel(float* __restrict__ out, long n, uint64_t seed, uint64_t stream, float stddev) {
  const long i4 = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  uint32_t c0 = static_cast<uint32_t>(i4), c1 = static_cast<uint32_t>(i4 >> 32);
  uint32_t c2 = static_cast<uint32_t>(stream), c3 = static_cast<uint32_t>(stream >> 32);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u0 = (static_cast<float>(c0) + 0.5f) * 2.3283064365386963e-10f;
  const float u1 = (static_cast<float>(c1) + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = (static_cast<float>(c2) + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = (static_cast<float>(c3) + 0.5f) * 2.3283064365386963e-10f;
  const float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float s0, co0, s1, co1;
  sincosf(6.283185307179586f * u1, &s0, &co0);
  sincosf(6.283185307179586f * u3, &s1, &co1);
  const float v[4] = {r0 * co0 * stddev, r0 * s0 * stddev, r1 * co1 * stddev, r1 * s1 * stddev};
  for (int j = 0; j < 4 && i4 * 4 + j < n; ++j) out[i4 * 4 + j] = v[j];
}

// Keras Adam over the FLAT parameter buffer (train.py:116-117, TF 2.2 ResourceApplyAdam form):
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t m / (sqrt(v) + eps)
// `trainable` is a per-element byte mask (BatchNorm moving statistics and alignment padding are skipped);
// grad_scale folds the 1/world_size of the data-parallel gradient mean.
// Number of non-finite entries of g[0, n) added to *count (fp16 gradient operands in loss-scale space can overflow:
// the optimiser kernels skip the whole update when the count is non-zero, the host then backs the loss scale off).
__global__ void nonfinite_count_kernel(const float* __restrict__ g, long n, float* __restrict__ count) {
  int bad = 0;
  const long n4 = n / 4;
  for (long j = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n4; j += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 t = *reinterpret_cast<const float4*>(g + j * 4);
    bad += !isfinite(t.x) + !isfinite(t.y) + !isfinite(t.z) + !isfinite(t.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long i = n4 * 4; i < n; ++i) bad += !isfinite(g[i]);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, static_cast<float>(bad));
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 const uint8_t* __restrict__ trainable, long n, float lr_t, float b1, float b2, float eps, float grad_scale,
                 const float* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0.f) return;   // non-finite gradients somewhere in the job: leave p, m, v untouched
  const long j = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long i = j * 4;
  if (i >= n) return;
  if (i + 3 < n) {                              // four elements per thread (the flat buffers are 16-byte aligned)
    const uchar4 tm = *reinterpret_cast<const uchar4*>(trainable + i);
    if (!(tm.x | tm.y | tm.z | tm.w)) return;
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float4 p4 = *reinterpret_cast<const float4*>(p + i), m4 = *reinterpret_cast<const float4*>(m + i),
           v4 = *reinterpret_cast<const float4*>(v + i);
    float* pe = &p4.x; float* me = &m4.x; float* ve = &v4.x; const float* ge = &g4.x;
    const unsigned char te[4] = {tm.x, tm.y, tm.z, tm.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (!te[e]) continue;
      const float gi = ge[e] * grad_scale;
      const float mi = b1 * me[e] + (1.f - b1) * gi;
      const float vi = b2 * ve[e] + (1.f - b2) * gi * gi;
      me[e] = mi;
      ve[e] = vi;
      pe[e] -= lr_t * mi / (sqrtf(vi) + eps);
    }
    *reinterpret_cast<float4*>(m + i) = m4;
    *reinterpret_cast<float4*>(v + i) = v4;
    *reinterpret_cast<float4*>(p + i) = p4;
    return;
  }
  for (long e = i; e < n; ++e) {
    if (!trainable[e]) continue;
    const float gi = g[e] * grad_scale;
    const float mi = b1 * m[e] + (1.f - b1) * gi;
    const float vi = b2 * v[e] + (1.f - b2) * gi * gi;
    m[e] = mi;
    v[e] = vi;
    p[e] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Data-parallel optimiser step fused with the gradient exchange over NVLink peer memory (one process per GPU, buffers
// shared through CUDA IPC).  Rank r owns the parameter shard [lo, hi): it READS the gradient shard of every replica
// straight from the peers' HBM (reduce-scatter), applies Keras Adam with its shard of the moments, and WRITES the
// updated parameters into every replica's parameter buffer (all-gather) -- one kernel, no staging buffers, the moments
// exist only once across the job.  Non-trainable entries (BatchNorm moving statistics: per-replica) are left alone.
// Ordering against the other ranks' kernels is the caller's job (a barrier before and after).
struct PeerPtrs {
  float* params[8];
  const float* grads[8];
};
__global__ void __launch_bounds__(256)
adam_sharded_kernel(const PeerPtrs pp, float* __restrict__ m, float* __restrict__ v, const uint8_t* __restrict__ trainable,
                    long lo, long hi, int rank, int world, float lr_t, float b1, float b2, float eps, float grad_scale,
                    const float* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag != 0.f) return;   // identical on every rank (all-reduced count): nobody writes anything
  const long n4 = (hi - lo) / 4;
  for (long j = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n4; j += static_cast<long>(gridDim.x) * blockDim.x) {
    const long i = lo + j * 4;
    const uchar4 tm = *reinterpret_cast<const uchar4*>(trainable + i);
    if (!(tm.x | tm.y | tm.z | tm.w)) continue;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 t = *reinterpret_cast<const float4*>(pp.grads[r] + i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 p = *reinterpret_cast<const float4*>(pp.params[rank] + i);
    float4 mi = *reinterpret_cast<const float4*>(m + j * 4);
    float4 vi = *reinterpret_cast<const float4*>(v + j * 4);
    float* pe = &p.x; float* me = &mi.x; float* ve = &vi.x; const float* ge = &g.x;
    const unsigned char te[4] = {tm.x, tm.y, tm.z, tm.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (!te[e]) continue;
      const float gi = ge[e] * grad_scale;
      me[e] = b1 * me[e] + (1.f - b1) * gi;
      ve[e] = b2 * ve[e] + (1.f - b2) * gi * gi;
      pe[e] -= lr_t * me[e] / (sqrtf(ve[e]) + eps);
    }
    *reinterpret_cast<float4*>(m + j * 4) = mi;
    *reinterpret_cast<float4*>(v + j * 4) = vi;
    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(pp.params[r] + i) = p;
  }
}

// Inverted-dropout keep mask (Keras Dropout): 0 with probability rate, else 1/(1-rate); same Philox stream layout.
__global__ void dropout_mask_kernel(float* __restrict__ out, long n, float rate, uint64_t seed, uint64_t stream) {
  const long i4 = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  uint32_t c0 = static_cast<uint32_t>(i4), c1 = static_cast<uint32_t>(i4 >> 32);
  uint32_t c2 = static_cast<uint32_t>(stream), c3 = static_cast<uint32_t>(stream >> 32) ^ 0x5D0Fu;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint32_t cs[4] = {c0, c1, c2, c3};
  const float keep = 1.0f / (1.0f - rate);
  for (int j = 0; j < 4 && i4 * 4 + j < n; ++j) {
    const float u = static_cast<float>(cs[j]) * 2.3283064365386963e-10f;
    out[i4 * 4 + j] = (u >= rate) ? keep : 0.f;
  }
}

}  // namespace vb

// Non-tensor-core kernels of the backward pass (train_step, train.py:120-138): LayerNorm / BatchNorm / activation /
// dropout backward with the fused bias-gradient column sums, the loss seeds, the flow (ActNorm, InvertibleLinear,
// affine coupling) and posterior backward, embedding / positional-weight / length-predictor gradients.
// All gradients live in "loss-scale space": every seed is multiplied by the loss scale S so that the fp16 operand
// copies of the activation gradients stay inside the fp16 range; Adam divides by S (grad_scale).
// Parameter gradients are ACCUMULATED (+=, atomics where several CTAs contribute) into the flat gradient buffer,
// which has the layout of the flat parameter buffer and is zeroed at the start of the step.
#pragma once
#include "simt_kernels.cuh"

namespace vb {

// ------------------------------------------------------------------ LayerNorm backward
// y = (u - mean) * rstd * gamma + beta (Keras LayerNormalization, eps 1e-3; modules/attention.py:402,428,433, utils.py:46).
// The forward saves y and rstd only; xhat = (y - beta) / gamma is reconstructed (gamma != 0; |gamma| < 1e-20 -> xhat = 0).
//   du = rstd * (gamma*dy - mean_j(gamma*dy) - xhat * mean_j(gamma*dy*xhat))
// Also: dgamma += sum_rows dy*xhat, dbeta += sum_rows dy, dbias += sum_rows du (bias of the Dense feeding the LN).
// One warp per row, 32 rows per CTA; du may alias dy.
template <int D>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* dy, const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ rstd, long rows, float* du, __half* __restrict__ du_h, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dbias) {
  constexpr int NJ = D / 128;   // float4 chunks per lane
  __shared__ float red[8][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g4[NJ][4], b4[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4);
    const float4 b = *reinterpret_cast<const float4*>(beta + (j * 32 + lane) * 4);
    g4[j][0] = g.x; g4[j][1] = g.y; g4[j][2] = g.z; g4[j][3] = g.w;
    b4[j][0] = b.x; b4[j][1] = b.y; b4[j][2] = b.z; b4[j][3] = b.w;
  }
  float ag[NJ][4], ab[NJ][4], au[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) ag[j][e] = ab[j][e] = au[j][e] = 0.f;
  const long r0 = static_cast<long>(blockIdx.x) * 32;
  // a warp owns rows warp, warp + 8, warp + 16, warp + 24 of the tile: all their loads are issued before the first store
  // (du may alias dy, which keeps the compiler from overlapping the rows itself)
  float4 dvs[4][NJ], yvs[4][NJ];
  float rss[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long row = r0 + warp + 8 * u;
    if (row < rows) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        dvs[u][j] = *reinterpret_cast<const float4*>(dy + row * D + (j * 32 + lane) * 4);
        yvs[u][j] = *reinterpret_cast<const float4*>(y + row * D + (j * 32 + lane) * 4);
      }
      rss[u] = rstd[row];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long row = r0 + warp + 8 * u;
    if (row >= rows) break;
    float d[NJ][4], xh[NJ][4];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 dv = dvs[u][j];
      const float4 yv = yvs[u][j];
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g = g4[j][e];
        const float x = fabsf(g) > 1e-20f ? (yy[e] - b4[j][e]) / g : 0.f;
        d[j][e] = dd[e];
        xh[j][e] = x;
        const float gd = g * dd[e];
        c1 += gd;
        c2 = fmaf(gd, x, c2);
        ag[j][e] = fmaf(dd[e], x, ag[j][e]);
        ab[j][e] += dd[e];
      }
    }
    c1 = warp_sum(c1) * (1.0f / D);
    c2 = warp_sum(c2) * (1.0f / D);
    const float rs = rss[u];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e] = rs * (g4[j][e] * d[j][e] - c1 - xh[j][e] * c2);
        au[j][e] += o[e];
      }
      *reinterpret_cast<float4*>(du + row * D + (j * 32 + lane) * 4) = make_float4(o[0], o[1], o[2], o[3]);
      if (du_h) {
        uint2 u2;
        u2.x = pack_half2(o[0], o[1]);
        u2.y = pack_half2(o[2], o[3]);
        *reinterpret_cast<uint2*>(du_h + row * D + (j * 32 + lane) * 4) = u2;
      }
    }
  }
  // cross-warp reduction of the three column-sum vectors, one at a time through the same shared array
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        red[warp][(j * 32 + lane) * 4 + e] = which == 0 ? ag[j][e] : (which == 1 ? ab[j][e] : au[j][e]);
    __syncthreads();
    if (dst)
      for (int c = threadIdx.x; c < D; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][c];
        atomicAdd(dst + c, s);
      }
  }
}

// ------------------------------------------------------------------ column-wise kernels: thread layout
// 256 threads over a row-major [rows, C] tile = `lanes` row lanes x `groups` column groups of V consecutive columns
// (vector loads, coalesced along the row); per-column sums are combined across the lanes through shared memory in a
// fixed order.  C must be a multiple of V and C / V <= 256.
template <int V>
struct ColLayout {
  int groups, lanes, g, lane;
  bool active;
  __device__ explicit ColLayout(int C) {
    groups = C / V;
    lanes = 256 / groups;
    g = threadIdx.x % groups;
    lane = threadIdx.x / groups;
    active = lane < lanes;
  }
};
__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __half* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h[e]);
    v[2 * e] = f.x; v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void st8h(__half* p, const float* v) {
  uint4 u;
  u.x = pack_half2(v[0], v[1]); u.y = pack_half2(v[2], v[3]); u.z = pack_half2(v[4], v[5]); u.w = pack_half2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------ activation / dropout backward + bias gradient
// out_h = g * mask * [act > 0]   (mask nullable: inverted-dropout keep mask; act: post-ReLU (and post-dropout) activation),
// dbias[c] += column sums.  RELU_ROWS rows per CTA, 8 columns per thread; C % 8 == 0, C <= 2048.  out_h may alias g
// (TG == __half), so the loads of four row steps are issued before the first store (memory-level parallelism: the
// compiler cannot hoist loads over possibly aliasing stores itself).
constexpr int RELU_ROWS = 32;
template <typename TG>
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const TG* g, const float* __restrict__ mask, const __half* __restrict__ act, long rows, int C,
                __half* out_h, float* __restrict__ dbias) {
  __shared__ float red[256 * 8];
  const ColLayout<8> L(C);
  const long r0 = static_cast<long>(blockIdx.x) * RELU_ROWS;
  const long r1 = min(rows, r0 + RELU_ROWS);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (L.active)
    for (long rb = r0 + L.lane; rb < r1; rb += 4L * L.lanes) {
      float v[4][8], a[4][8], m[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long r = rb + static_cast<long>(u) * L.lanes;
        if (r < r1) {
          const long i = r * C + L.g * 8;
          ld8(g + i, v[u]);
          if (mask) ld8(mask + i, m[u]);
          if (act) ld8(act + i, a[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long r = rb + static_cast<long>(u) * L.lanes;
        if (r < r1) {
          if (mask) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] *= m[u][e];
          }
          if (act) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] = a[u][e] > 0.f ? v[u][e] : 0.f;
          }
          st8h(out_h + r * C + L.g * 8, v[u]);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += v[u][e];
        }
      }
    }
  if (dbias) {
    __syncthreads();
    if (L.active)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[L.lane * C + L.g * 8 + e] = acc[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
      for (int l = 0; l < L.lanes; ++l) t += red[l * C + c];
      atomicAdd(dbias + c, t);
    }
  }
}

// Column sums of a row-major [rows, ld] tensor window of C columns: out0[c] += sum (c < split), out1[c - split] otherwise.
// 128 rows per CTA; V = 4 (float) / 8 (half) columns per thread.
template <typename T, int V>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ in, long rows, int C, int ld, float* __restrict__ out0, int split,
              float* __restrict__ out1) {
  __shared__ float red[256 * V];
  const ColLayout<V> L(C);
  const long r0 = static_cast<long>(blockIdx.x) * 128;
  const long r1 = min(rows, r0 + 128);
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  if (L.active)
    for (long r = r0 + L.lane; r < r1; r += L.lanes) {
      float v[8];
      if constexpr (V == 8) {
        ld8(in + r * ld + L.g * 8, v);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(in + r * ld + L.g * 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      }
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] += v[e];
    }
  __syncthreads();
  if (L.active)
#pragma unroll
    for (int e = 0; e < V; ++e) red[L.lane * C + L.g * V + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < L.lanes; ++l) t += red[l * C + c];
    if (c < split) atomicAdd(out0 + c, t);
    else atomicAdd(out1 + (c - split), t);
  }
}

// ------------------------------------------------------------------ BatchNorm (training mode) backward
// Forward (modules/utils.py:56-85): a = act(conv + bias);  xhat = (a - mean) * rstd;  out = (gamma*xhat + beta) * mask.
// partial[tile][0][c] = sum dy', partial[tile][1][c] = sum dy' * xhat with dy' = dout * mask  (fixed-order, deterministic)
// 4 columns per thread; C % 4 == 0, C <= 1024.
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ mask, const float* __restrict__ a,
                    const float* __restrict__ mean, const float* __restrict__ rstd, long rows, int C, int rows_per_tile,
                    float* __restrict__ partial) {
  __shared__ float red[256 * 4];
  const ColLayout<4> L(C);
  const long r0 = static_cast<long>(blockIdx.x) * rows_per_tile;
  const long r1 = min(rows, r0 + rows_per_tile);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (L.active) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + L.g * 4), rs = *reinterpret_cast<const float4*>(rstd + L.g * 4);
    const float mu4[4] = {mu.x, mu.y, mu.z, mu.w}, rs4[4] = {rs.x, rs.y, rs.z, rs.w};
    for (long r = r0 + L.lane; r < r1; r += L.lanes) {
      const long i = r * C + L.g * 4;
      float4 d = *reinterpret_cast<const float4*>(dout + i);
      if (mask) {
        const float4 m = *reinterpret_cast<const float4*>(mask + i);
        d.x *= m.x; d.y *= m.y; d.z *= m.z; d.w *= m.w;
      }
      const float4 av = *reinterpret_cast<const float4*>(a + i);
      const float d4[4] = {d.x, d.y, d.z, d.w}, a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[e] += d4[e];
        q[e] = fmaf(d4[e], (a4[e] - mu4[e]) * rs4[e], q[e]);
      }
    }
  }
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
    if (L.active)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[L.lane * C + L.g * 4 + e] = which ? q[e] : s[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
      for (int l = 0; l < L.lanes; ++l) t += red[l * C + c];
      partial[(static_cast<long>(blockIdx.x) * 2 + which) * C + c] = t;
    }
  }
}
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int ntile, int C, float* __restrict__ s1,
                                       float* __restrict__ s2, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int t = 0; t < ntile; ++t) {
    s += partial[(static_cast<long>(t) * 2 + 0) * C + c];
    q += partial[(static_cast<long>(t) * 2 + 1) * C + c];
  }
  s1[c] = static_cast<float>(s);
  s2[c] = static_cast<float>(q);
  dbeta[c] += static_cast<float>(s);
  dgamma[c] += static_cast<float>(q);
}
// da = gamma*rstd*(dy' - s1/N - xhat*s2/N);  dpre = da * act'(a)  (act 1: relu, 2: tanh, 0: identity) -> fp16 operand of
// the conv dgrad / wgrad, plus the conv bias gradient (column sums).  64 rows per CTA, 4 columns per thread.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ mask, const float* __restrict__ a,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ s1, const float* __restrict__ s2, long rows, int C, int act,
                    __half* __restrict__ dpre_h, float* __restrict__ dbias) {
  __shared__ float red[256 * 4];
  const ColLayout<4> L(C);
  const long r0 = static_cast<long>(blockIdx.x) * 64;
  const long r1 = min(rows, r0 + 64);
  const float inv_n = 1.0f / static_cast<float>(rows);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (L.active) {
    float mu[4], rs[4], gr[4], m1[4], m2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = L.g * 4 + e;
      mu[e] = mean[c]; rs[e] = rstd[c]; gr[e] = gamma[c] * rs[e]; m1[e] = s1[c] * inv_n; m2[e] = s2[c] * inv_n;
    }
    for (long r = r0 + L.lane; r < r1; r += L.lanes) {
      const long i = r * C + L.g * 4;
      float4 d = *reinterpret_cast<const float4*>(dout + i);
      if (mask) {
        const float4 m = *reinterpret_cast<const float4*>(mask + i);
        d.x *= m.x; d.y *= m.y; d.z *= m.z; d.w *= m.w;
      }
      const float4 av = *reinterpret_cast<const float4*>(a + i);
      const float d4[4] = {d.x, d.y, d.z, d.w}, a4[4] = {av.x, av.y, av.z, av.w};
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = gr[e] * (d4[e] - m1[e] - (a4[e] - mu[e]) * rs[e] * m2[e]);
        if (act == 1) v[e] = a4[e] > 0.f ? v[e] : 0.f;
        else if (act == 2) v[e] *= 1.f - a4[e] * a4[e];
        acc[e] += v[e];
      }
      uint2 u;
      u.x = pack_half2(v[0], v[1]);
      u.y = pack_half2(v[2], v[3]);
      *reinterpret_cast<uint2*>(dpre_h + i) = u;
    }
  }
  __syncthreads();
  if (L.active)
#pragma unroll
    for (int e = 0; e < 4; ++e) red[L.lane * C + L.g * 4 + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < L.lanes; ++l) t += red[l * C + c];
    atomicAdd(dbias + c, t);
  }
}

// ------------------------------------------------------------------ losses: scalars, seeds of the backward pass
// total = mean l2 + kl_w * max(mean kl, 0) + len_w * mean length_l2   (train.py:135)
// losses[4] = {total, l2, kl, length};  coef_q[b] = dL/dlogq[b] * S,  coef_p[b] = dL/dlogp[b] * S,
// coef_len[0] = sum_b coef_p[b] * z_len[b]  (weight of the length-proportional ActNorm / InvertibleLinear log-dets),
// dpred[b] = S * len_w * 2 (ln pred - ln len) / (pred * B).
__global__ void loss_finalize_kernel(const float* __restrict__ l2, const float* __restrict__ kl,
                                     const float* __restrict__ len_loss, const float* __restrict__ pred,
                                     const int* __restrict__ m_len, const int* __restrict__ z_len, int B, float kl_w,
                                     float len_w, float S, float* __restrict__ losses, float* __restrict__ coef_q,
                                     float* __restrict__ coef_p, float* __restrict__ coef_len, float* __restrict__ dpred) {
  __shared__ float sh[32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += l2[i]; b += kl[i]; c += len_loss[i]; }
  const float l2m = block_sum(a, sh) / B;
  const float klm = block_sum(b, sh) / B;
  const float lm = block_sum(c, sh) / B;
  const float gate = klm > 0.f ? 1.f : 0.f;
  float cl = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float cq = S * kl_w * gate / B;
    coef_q[i] = cq;
    coef_p[i] = -cq;
    cl += -cq * static_cast<float>(z_len[i]);
    dpred[i] = S * len_w * 2.f * (logf(pred[i]) - logf(static_cast<float>(m_len[i]))) / (pred[i] * B);
  }
  const float clt = block_sum(cl, sh);
  if (threadIdx.x == 0) {
    losses[0] = l2m + kl_w * fmaxf(klm, 0.f) + len_w * lm;
    losses[1] = l2m;
    losses[2] = klm;
    losses[3] = lm;
    coef_len[0] = clt;
  }
}

// Masked-MSE gradients (models/models.py:67-86,182-188): for rows t < min(len_b, Tm):
//   g_fin = coef_b (fin - mel),  g_ini = coef_b (ini - mel) + g_fin   (the final mel is postnet(ini) + ini),
//   coef_b = S * 2 / (D * len_b * B).  Rows beyond the target (cropped, models.py:182-183) get zero.
__global__ void l2_grad_kernel(const float* __restrict__ fin, const float* __restrict__ ini, int rec_T,
                               const float* __restrict__ tgt, int tgt_T, const int* __restrict__ lens, int B, int D,
                               float S, float* __restrict__ g_fin, __half* __restrict__ g_fin_h,
                               float* __restrict__ g_ini) {
  const long n = static_cast<long>(B) * rec_T * D;
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = static_cast<int>(i % D);
  const long row = i / D;
  const int t = static_cast<int>(row % rec_T), b = static_cast<int>(row / rec_T);
  const int len = lens[b];
  float gf = 0.f, gi = 0.f;
  if (t < len && t < tgt_T) {
    const float coef = S * 2.f / (static_cast<float>(D) * static_cast<float>(len) * static_cast<float>(B));
    const float m = tgt[(static_cast<long>(b) * tgt_T + t) * D + d];
    gf = coef * (fin[i] - m);
    gi = coef * (ini[i] - m) + gf;
  }
  g_fin[i] = gf;
  g_fin_h[i] = __float2half_rn(gf);
  g_ini[i] = gi;
}

// DenseLengthPredictor backward (modules/length_predictor.py:35-42; the input is stop_gradient'ed, models.py:132-135):
// pred_b = sum_{t<len} exp(x_t . w + bias)  ->  dw += dpred_b * sum_t e_t x_t,  dbias += dpred_b * sum_t e_t
__global__ void length_predictor_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                            const float* __restrict__ bias, const int* __restrict__ lens,
                                            const float* __restrict__ dpred, int T, int D, float* __restrict__ dw,
                                            float* __restrict__ dbias) {
  extern __shared__ float sm_lp[];   // [D] accumulators + [T] exp values
  float* acc = sm_lp;
  float* ev = sm_lp + D;
  const int b = blockIdx.x;
  const int len = min(lens[b], T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t < len; t += nw) {
    const float* row = x + (static_cast<long>(b) * T + t) * D;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot += row[d] * w[d];
    dot = warp_sum(dot);
    if (lane == 0) ev[t] = expf(dot + bias[0]);
  }
  __syncthreads();
  const float dp = dpred[b];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < len; ++t) s = fmaf(ev[t], x[(static_cast<long>(b) * T + t) * D + d], s);
    atomicAdd(dw + d, dp * s);
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int t = 0; t < len; ++t) s += ev[t];
    atomicAdd(dbias, dp * s);
  }
  (void)acc;
}

// ------------------------------------------------------------------ posterior / flow backward
// Posterior (modules/posterior.py:20-72 with the models.py:136 name swap): z = eps * exp(lv/2) + mu,
// log q = sum_{t<len} -1/2 (L ln 2pi + sum_d (lv + eps^2)).  Given g_z and coef_q[b] = dL/dlogq:
//   dmu = g_z ;  dlv = g_z * eps * exp(lv/2) / 2 - coef_q[b] [t < len] / 2  -> fp16 [rows, 2L] = [dlv | dmu]
__global__ void posterior_bwd_kernel(const float* __restrict__ g_z, const float* __restrict__ eps,
                                     const float* __restrict__ lv, const float* __restrict__ coef_q,
                                     const int* __restrict__ lens, int T, int L, long rows, __half* __restrict__ dout) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const long row = i / L;
  const int d = static_cast<int>(i % L);
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  const float g = g_z[i];
  float dl = 0.5f * g * eps[i] * expf(0.5f * lv[i]);
  if (t < lens[b]) dl -= 0.5f * coef_q[b];
  dout[row * 2 * L + d] = __float2half_rn(dl);
  dout[row * 2 * L + L + d] = __float2half_rn(g);
}

// Gaussian base density of the prior (modules/prior.py:147-151): g_eps = coef_p[b] * [t < len] * (-eps)
__global__ void base_logprob_bwd_kernel(const float* __restrict__ e, const float* __restrict__ coef_p,
                                        const int* __restrict__ lens, int T, int L, long rows, float* __restrict__ g) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const long row = i / L;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  g[i] = t < lens[b] ? -coef_p[b] * e[i] : 0.f;
}

// Affine coupling, training direction (modules/flow.py:240-257): zp_out = (zp - shift) / (scale + 1e-12),
// scale = sigmoid(ls + 2), log-det row term -sum_d log scale on rows t < len.  Given g (gradient w.r.t. the step OUTPUT,
// [rows, 2h], updated in place to the gradient w.r.t. the step INPUT for the transformed half; the conditioning half
// receives the conditioner's input gradient later) and coef_p[b]:
//   g_zp_in = g / (scale + 1e-12);  dshift = -g_zp_in;  dscale = -g_zp_in * zp_out - coef_p[b] [t<len] / scale;
//   dls = dscale * scale (1 - scale)        -> dout fp16 [rows, 2h] = [dls | dshift]
__global__ void coupling_bwd_kernel(float* __restrict__ g, const float* __restrict__ z_out, const float* __restrict__ scale,
                                    const float* __restrict__ coef_p, const int* __restrict__ lens, int T, int hdim,
                                    int zp_off, long rows, __half* __restrict__ dout) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * hdim) return;
  const long row = i / hdim;
  const int d = static_cast<int>(i % hdim);
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  const long zi = row * 2 * hdim + zp_off + d;
  const float sc = scale[i];
  const float gi = g[zi] / (sc + 1e-12f);
  float dsc = -gi * z_out[zi];
  if (t < lens[b]) dsc -= coef_p[b] / sc;
  g[zi] = gi;
  dout[row * 2 * hdim + d] = __float2half_rn(dsc * sc * (1.f - sc));
  dout[row * 2 * hdim + hdim + d] = __float2half_rn(-gi);
}

// Parameter gradients of one ActNorm (+) InvertibleLinear pair in the training direction (modules/flow.py:136-150,
// 176-187):  z_out = z_in Mb + cb,  Mb = V diag(D), cb = -b D,  V = W^-1,  D_j = 1 / (exp(s_j) + 1e-8);
// log-dets -len (log|det W| + sum_j s_j) enter the loss with weight coef_len = sum_b coef_p[b] len_b.
// Inputs: G = dL/dMb = z_in^T g_out [128,128], gc = dL/dcb = colsum(g_out).  One CTA (256 threads) per flow step.
//   dL/dV = G diag(D);  dL/dW = -V^T (dL/dV) V^T - coef_len V^T;
//   dD_j = sum_i V_ij G_ij - b_j gc_j;  ds_j = -dD_j exp(s_j) D_j^2 - coef_len;  db_j = -D_j gc_j.
__global__ void __launch_bounds__(256)
flow_param_grad_kernel(const float* __restrict__ G_all, const float* __restrict__ gc_all, const float* __restrict__ winv_all,
                       const float* const* __restrict__ log_scales, const float* const* __restrict__ biases,
                       const float* __restrict__ coef_len, float* const* __restrict__ dW, float* const* __restrict__ ds,
                       float* const* __restrict__ db) {
  extern __shared__ float sm_fp[];
  constexpr int N = FLOW_DIM;
  float* V = sm_fp;              // [N][N+1]
  float* A = V + N * (N + 1);    // [N][N+1]  dL/dV, later T = A V^T
  float* Dv = A + N * (N + 1);   // [N]
  const int s = blockIdx.x;
  const float* G = G_all + static_cast<long>(s) * N * N;
  const float* gc = gc_all + s * N;
  const float* Vg = winv_all + static_cast<long>(s) * N * N;
  const float cl = coef_len[0];
  if (threadIdx.x < N) Dv[threadIdx.x] = 1.f / (expf(log_scales[s][threadIdx.x]) + 1e-8f);
  __syncthreads();
  for (int i = threadIdx.x; i < N * N; i += 256) {
    const int r = i / N, c = i % N;
    V[r * (N + 1) + c] = Vg[i];
    A[r * (N + 1) + c] = G[i] * Dv[c];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int j = threadIdx.x;
    float dD = 0.f;
    for (int i = 0; i < N; ++i) dD = fmaf(V[i * (N + 1) + j], G[i * N + j], dD);
    const float bj = biases[s][j];
    dD -= bj * gc[j];
    const float es = expf(log_scales[s][j]);
    ds[s][j] += -dD * es * Dv[j] * Dv[j] - cl;
    db[s][j] += -Dv[j] * gc[j];
  }
  // T[k][j] = sum_l A[k][l] V[j][l]; each thread: row k = tid / 2, 64 columns
  {
    const int k = threadIdx.x >> 1, j0 = (threadIdx.x & 1) * 64;
    float t[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) t[j] = 0.f;
    for (int l = 0; l < N; ++l) {
      const float a = A[k * (N + 1) + l];
#pragma unroll
      for (int j = 0; j < 64; ++j) t[j] = fmaf(a, V[(j0 + j) * (N + 1) + l], t[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 64; ++j) A[k * (N + 1) + j0 + j] = t[j];
  }
  __syncthreads();
  // dW[i][j] += -sum_k V[k][i] T[k][j] - cl * V[j][i]
  {
    const int i = threadIdx.x >> 1, j0 = (threadIdx.x & 1) * 64;
    float r[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) r[j] = 0.f;
    for (int k = 0; k < N; ++k) {
      const float v = V[k * (N + 1) + i];
#pragma unroll
      for (int j = 0; j < 64; ++j) r[j] = fmaf(v, A[k * (N + 1) + j0 + j], r[j]);
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) dW[s][i * N + j0 + j] += -r[j] - cl * V[(j0 + j) * (N + 1) + i];
  }
}

// ------------------------------------------------------------------ positional weight / embedding / misc
// dpw += sum_{rows, c} g[row, c] * table[t(row), c]   (x = dense(...) + pos_weight * PE, encoder.py:84-86 etc.); C % 4 == 0
__global__ void __launch_bounds__(256)
pe_dot_kernel(const float* __restrict__ g, const float* __restrict__ mask, const float* __restrict__ table, long rows, int T,
              int C, float* __restrict__ dpw) {
  __shared__ float sh[32];
  const long n4 = rows * C / 4;
  const int c4 = C / 4;
  float acc = 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / c4;
    const int c = static_cast<int>(i - row * c4) * 4, t = static_cast<int>(row % T);
    float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    if (mask) {
      const float4 m = *reinterpret_cast<const float4*>(mask + i * 4);
      v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    }
    const float4 p = *reinterpret_cast<const float4*>(table + static_cast<long>(t) * C + c);
    acc += v.x * p.x + v.y * p.y + v.z * p.z + v.w * p.w;
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(dpw, tot);
}

// Embedding gradient (modules/encoder.py:10-12,81): dE[v, :] += sum over tokens with id v of g[token, :].
// grid (V, C / 32); 256 threads = 32 columns x 8 row lanes, lanes combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int* __restrict__ ids, const float* __restrict__ g, long rows, int C, int V, float* __restrict__ dE) {
  __shared__ float red[8][32];
  const int v = blockIdx.x;
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (long r = lane; r < rows; r += 8) {
      int id = ids[r];
      id = min(max(id, 0), V - 1);
      if (id == v) s += g[r * C + c];
    }
  red[lane][cl] = s;
  __syncthreads();
  if (lane == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += red[l][cl];
    dE[static_cast<long>(v) * C + c] += t;
  }
}

// out = a + b (fp32) ; optional fp16 copy
__global__ void add_kernel(const float* a, const float* b, float* out, __half* out_h, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] + b[i];
  out[i] = v;
  if (out_h) out_h[i] = __float2half_rn(v);
}

}  // namespace vb

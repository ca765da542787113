// libvaenar_sm100.so -- host-side engine + C ABI (include/vaenar_b200.h) of the B200-native VAENAR-TTS
// mel-synthesis path.  Owns no device memory: parameters, packed weights and workspace are caller buffers.
// Every forward below is a fixed sequence of sm_100a kernel launches on the caller's stream (CUDA-graph
// capturable: no host synchronisation, no allocation).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vaenar_b200.h"
#include "attention_tc.cuh"
#include "attention_bwd_tc.cuh"
#include "gemm_tc.cuh"
#include "simt_kernels.cuh"
#include "wgrad_tc.cuh"
#include "xblk_fused.cuh"
#include "bwd_kernels.cuh"
#include "griffin_lim.cuh"

using namespace vb;

// ============================================================================ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}
struct EngineError {
  std::string msg;
};
#define VB_THROW(...)                         \
  do {                                        \
    char _b[1024];                            \
    snprintf(_b, sizeof(_b), __VA_ARGS__);    \
    throw EngineError{_b};                    \
  } while (0)
#define VB_CUDA(x)                                                                              \
  do {                                                                                          \
    cudaError_t _e = (x);                                                                       \
    if (_e != cudaSuccess) VB_THROW("%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ============================================================================ TMA descriptor encode
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      VB_THROW("cuTensorMapEncodeTiled unavailable (no CUDA driver / device?): %s", cudaGetErrorString(e));
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}
// fp16 tensor [d2][d1][d0] with d0 contiguous; strides in ELEMENTS; box {b0, b1, 1}; 128B swizzle; OOB -> 0
static CUtensorMap make_tmap(const void* ptr, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                             uint64_t s2, uint32_t b0, uint32_t b1) {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1 * 2, s2 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (rank == 3 && (strides[1] & 15)))
    VB_THROW("tensor map alignment: ptr %p strides %llu %llu", ptr, (unsigned long long)strides[0],
             (unsigned long long)strides[1]);
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) VB_THROW("cuTensorMapEncodeTiled failed (%d) dims %llu %llu %llu", (int)r,
                                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return m;
}

// ============================================================================ model description
struct ParamInfo {
  std::string name;
  int ndim;
  int64_t dims[3];
  int64_t offset, numel;
  bool trainable;
};
struct PackedMat {
  int64_t off;   // bytes into the packed arena
  int N, K;      // [N, K] fp16, K-major
};
struct PackPlanOp {
  int param;         // source parameter index
  int64_t src_off;   // float offset inside the parameter
  int K, N, lds;
  std::string dst;
  int n_off, k_off, mode;
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct vaenar_model {
  vaenar_hparams_t hp;
  std::vector<ParamInfo> params;
  std::map<std::string, int> pidx;
  int64_t param_floats = 0;
  std::map<std::string, PackedMat> pmats;
  std::map<std::string, int64_t> pvecs;   // fp32 vectors in the packed arena (byte offsets)
  std::vector<PackPlanOp> plan;
  int64_t packed_bytes = 0;
  int64_t off_packops = 0, off_packtiles = 0, off_ptrs = 0, off_logdet64 = 0, off_winv = 0;
  std::vector<int> host_tiles;          // first CTA of every pack op (prefix sums)
  cudaEvent_t ev_flow_ready = nullptr;  // the flow constants (W^-1, log|det W|, folded maps) are produced on the side stream
  bool flow_pending = false;
  int64_t off_Mf = 0, off_cf = 0, off_Mb = 0, off_cb = 0, off_consts = 0;
  std::vector<PackOp> host_ops;
  std::vector<const float*> host_ptrs;
  std::vector<float*> host_gptrs;   // flow parameter-gradient pointers (train step)
  cudaStream_t wgrad_stream = nullptr;   // weight gradients run beside the activation-gradient chain (train step)
  cudaStream_t lane_stream = nullptr;    // train step: decoder forward + backward run beside the prior's (both hang off z only)
  cudaStream_t lane_wgrad_stream = nullptr;
  cudaStream_t main_stream = nullptr;    // train step: high-priority stream of the activation chain (forked off the caller's)
  cudaStream_t attn_stream = nullptr, lane_attn_stream = nullptr;   // dK/dV kernels of the attention backward (beside dQ)
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_cursor = 0;
  bool attrs_set = false;
  cudaStream_t side_stream = nullptr;   // second launch chain for batch-split inference (fork/join with events)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  int add_param(const std::string& n, std::initializer_list<int64_t> dims, bool trainable = true) {
    ParamInfo p;
    p.name = n;
    p.ndim = static_cast<int>(dims.size());
    p.numel = 1;
    int i = 0;
    for (auto d : dims) { p.dims[i++] = d; p.numel *= d; }
    for (; i < 3; ++i) p.dims[i] = 1;
    p.offset = param_floats;
    p.trainable = trainable;
    param_floats = align_up(param_floats + p.numel, 64);
    pidx[n] = static_cast<int>(params.size());
    params.push_back(p);
    return static_cast<int>(params.size()) - 1;
  }
  int P(const std::string& n) const {
    auto it = pidx.find(n);
    if (it == pidx.end()) VB_THROW("unknown parameter %s", n.c_str());
    return it->second;
  }
  void add_mat(const std::string& n, int N, int K) {
    packed_bytes = align_up(packed_bytes, 1024);
    pmats[n] = PackedMat{packed_bytes, N, K};
    packed_bytes += static_cast<int64_t>(N) * K * 2;
  }
  void add_vec(const std::string& n, int count) {
    packed_bytes = align_up(packed_bytes, 256);
    pvecs[n] = packed_bytes;
    packed_bytes += static_cast<int64_t>(count) * 4;
  }
  void add_op(const std::string& dst, int n_off, int k_off, const std::string& param, int64_t src_off, int K, int N,
              int lds, int mode) {
    plan.push_back(PackPlanOp{P(param), src_off, K, N, lds, dst, n_off, k_off, mode});
  }
};

// ---------------------------------------------------------------------------- manifest
static void add_dense(vaenar_model& m, const std::string& n, int din, int dout, bool bias = true) {
  m.add_param(n + ".kernel", {din, dout});
  if (bias) m.add_param(n + ".bias", {dout});
}
static void add_ln(vaenar_model& m, const std::string& n, int d) {
  m.add_param(n + ".gamma", {d});
  m.add_param(n + ".beta", {d});
}
static void add_conv_bn(vaenar_model& m, const std::string& n, int k, int cin, int cout) {
  m.add_param(n + ".conv1d.kernel", {k, cin, cout});
  m.add_param(n + ".conv1d.bias", {cout});
  m.add_param(n + ".bn.gamma", {cout});
  m.add_param(n + ".bn.beta", {cout});
  m.add_param(n + ".bn.moving_mean", {cout}, false);
  m.add_param(n + ".bn.moving_variance", {cout}, false);
}
static void add_ffn(vaenar_model& m, const std::string& n, int d, int hidden) {
  add_dense(m, n + ".dense1", d, hidden);
  add_dense(m, n + ".dense2", hidden, d);
  add_ln(m, n + ".layer_norm", d);
}
static void add_xblk(vaenar_model& m, const std::string& n, int d, int mem, int ffn) {
  for (const char* q : {"query", "key", "value"}) add_dense(m, n + ".self_attention." + q + "_layer", d, d, false);
  add_dense(m, n + ".att_proj1", 2 * d, d);
  add_ln(m, n + ".layer_norm1", d);
  add_dense(m, n + ".cross_attention.query_layer", d, d, false);
  add_dense(m, n + ".cross_attention.key_layer", mem, d, false);
  add_dense(m, n + ".cross_attention.value_layer", mem, d, false);
  add_dense(m, n + ".att_proj2", 2 * d, d);
  add_ln(m, n + ".layer_norm2", d);
  add_ffn(m, n + ".ffn", d, ffn);
}

static int kpad64(int k) { return cdiv(k, 64) * 64; }
// tuning aid: the first VAENAR_POST_PLAIN PostNet convolutions run on plain fp16 operands instead of split-fp16 (inference only)
static int post_plain() {
  static const int v = getenv("VAENAR_POST_PLAIN") ? atoi(getenv("VAENAR_POST_PLAIN")) : 0;
  return v;
}

// packed operand plan of one CrossAttentionBLK (modules/attention.py:418-452)
static void plan_xblk(vaenar_model& m, const std::string& pk, const std::string& n, int d, int ffn) {
  m.add_mat(pk + ".qkv", 3 * d, d);
  int i = 0;
  for (const char* q : {"query", "key", "value"})
    m.add_op(pk + ".qkv", (i++) * d, 0, n + ".self_attention." + q + "_layer.kernel", 0, d, d, d, 0);
  m.add_mat(pk + ".proj1", d, 2 * d);
  m.add_op(pk + ".proj1", 0, 0, n + ".att_proj1.kernel", 0, 2 * d, d, d, 0);
  m.add_mat(pk + ".cq", d, d);
  m.add_op(pk + ".cq", 0, 0, n + ".cross_attention.query_layer.kernel", 0, d, d, d, 0);
  m.add_mat(pk + ".proj2", d, 2 * d);
  m.add_op(pk + ".proj2", 0, 0, n + ".att_proj2.kernel", 0, 2 * d, d, d, 0);
  m.add_mat(pk + ".ffn1", ffn, d);
  m.add_op(pk + ".ffn1", 0, 0, n + ".ffn.dense1.kernel", 0, d, ffn, ffn, 0);
  m.add_mat(pk + ".ffn2", d, ffn);
  m.add_op(pk + ".ffn2", 0, 0, n + ".ffn.dense2.kernel", 0, ffn, d, d, 0);
}
// memory K/V projections of several blocks in ONE matrix: rows [K_0 .. K_{n-1} | V_0 .. V_{n-1}]
static void plan_memkv(vaenar_model& m, const std::string& pk, const std::vector<std::string>& blks, int mem, int d) {
  const int nb = static_cast<int>(blks.size());
  m.add_mat(pk, 2 * nb * d, mem);
  for (int i = 0; i < nb; ++i) {
    m.add_op(pk, i * d, 0, blks[i] + ".cross_attention.key_layer.kernel", 0, mem, d, d, 0);
    m.add_op(pk, nb * d + i * d, 0, blks[i] + ".cross_attention.value_layer.kernel", 0, mem, d, d, 0);
  }
}
// Conv1D kernel [k, cin, cout] -> [cout, k * parts * cpad]; split: parts (hi, hi, lo) to pair with A (hi, lo, hi)
static void plan_conv(vaenar_model& m, const std::string& pk, const std::string& n, int k, int cin, int cout,
                      bool split) {
  const int cp = kpad64(cin), parts = split ? 3 : 1;
  m.add_mat(pk, cout, k * parts * cp);
  for (int j = 0; j < k; ++j)
    for (int p = 0; p < parts; ++p)
      m.add_op(pk, 0, (j * parts + p) * cp, n + ".conv1d.kernel", static_cast<int64_t>(j) * cin * cout, cin, cout, cout,
               (split && p == 2) ? 1 : 0);
  m.add_vec(pk + ".bn_scale", cout);
  m.add_vec(pk + ".bn_shift", cout);
}

// ---- backward-pass (dgrad) operands: the Keras [in, out] layout itself as fp16 = [N_gemm = in, K_gemm = out] K-major
static void plan_bw(vaenar_model& m, const std::string& pk, const std::string& param, int in, int out) {
  m.add_mat(pk, in, out);
  m.add_op(pk, 0, 0, param, 0, in, out, out, 2);
}
static void plan_bw_xblk(vaenar_model& m, const std::string& pk, const std::string& n, int d, int ffn) {
  int i = 0;
  plan_bw(m, "bw." + pk + ".proj1", n + ".att_proj1.kernel", 2 * d, d);
  plan_bw(m, "bw." + pk + ".proj2", n + ".att_proj2.kernel", 2 * d, d);
  // K-concatenated operands: all gradient contributions to one residual-stream tensor in ONE GEMM
  //   d s = du2 Wp2[:d]^T + dq2 Wcq^T                  -> [d, d | d]
  //   d x = du1 Wp1[:d]^T + dqkv [Wq | Wk | Wv]^T      -> [d, d | 3d]
  m.add_mat("bw." + pk + ".sq", d, 2 * d);
  m.add_op("bw." + pk + ".sq", 0, 0, n + ".att_proj2.kernel", 0, d, d, d, 2);
  m.add_op("bw." + pk + ".sq", 0, d, n + ".cross_attention.query_layer.kernel", 0, d, d, d, 2);
  m.add_mat("bw." + pk + ".xq", d, 4 * d);
  m.add_op("bw." + pk + ".xq", 0, 0, n + ".att_proj1.kernel", 0, d, d, d, 2);
  i = 0;
  for (const char* q : {"query", "key", "value"})
    m.add_op("bw." + pk + ".xq", 0, d + (i++) * d, n + ".self_attention." + q + "_layer.kernel", 0, d, d, d, 2);
  plan_bw(m, "bw." + pk + ".ffn1", n + ".ffn.dense1.kernel", d, ffn);
  plan_bw(m, "bw." + pk + ".ffn2", n + ".ffn.dense2.kernel", ffn, d);
}
static void plan_bw_memkv(vaenar_model& m, const std::string& pk, const std::vector<std::string>& blks, int mem, int d) {
  const int nb = static_cast<int>(blks.size());
  m.add_mat(pk, mem, 2 * nb * d);
  for (int i = 0; i < nb; ++i) {
    m.add_op(pk, 0, i * d, blks[i] + ".cross_attention.key_layer.kernel", 0, mem, d, d, 2);
    m.add_op(pk, 0, nb * d + i * d, blks[i] + ".cross_attention.value_layer.kernel", 0, mem, d, d, 2);
  }
}
static void plan_bw_conv(vaenar_model& m, const std::string& pk, const std::string& n, int k, int cin, int cout) {
  m.add_mat(pk, cin, k * cout);
  for (int j = 0; j < k; ++j)
    m.add_op(pk, 0, j * cout, n + ".conv1d.kernel", static_cast<int64_t>(j) * cin * cout, cin, cout, cout, 2);
}

static void build_model(vaenar_model& m) {
  const vaenar_hparams_t& h = m.hp;
  // ---- supported envelope of the hand-written kernels
  if (h.enc_att_dim / h.enc_heads != 64 || h.dec_att_dim / h.dec_heads != 64 ||
      h.posterior_att_dim / h.posterior_heads != 64 || h.prior_att_dim / h.prior_heads != 64)
    VB_THROW("kernels are specialised for head_dim 64");
  if (h.latent_dim != FLOW_DIM) VB_THROW("flow kernels are specialised for latent_dim 128");
  if (h.enc_hidden != 512 || h.embd_dim != 512) VB_THROW("encoder width must be 512 (LayerNorm epilogue tile)");
  if (h.dec_att_dim != 256 || h.posterior_att_dim != 256 || h.prior_att_dim != 256 || h.enc_att_dim != 256 ||
      h.posterior_pre_hidden != 256 || h.post_filters != 256)
    VB_THROW("decoder/posterior/prior width must be 256 (LayerNorm epilogue tile)");
  if (h.enc_conv_kernel * 1 > kMaxSegs || h.post_kernel * 3 > kMaxSegs) VB_THROW("conv kernel too wide");
  const int E = h.enc_hidden, O = h.out_dim, L = h.latent_dim;

  // ---- parameters
  m.add_param("text_encoder.emb_layer.embeddings", {h.vocab_size, h.embd_dim});
  m.add_param("text_encoder.pos_weight", {1});
  {
    int cin = h.embd_dim;
    for (int i = 0; i < h.enc_n_conv; ++i) {
      add_conv_bn(m, "text_encoder.prenet.conv_stack." + std::to_string(i), h.enc_conv_kernel, cin, E);
      cin = E;
    }
  }
  add_dense(m, "text_encoder.prenet.projection", E, E);
  for (int i = 0; i < h.enc_n_blk; ++i) {
    const std::string n = "text_encoder.self_attentions." + std::to_string(i);
    for (const char* q : {"query", "key", "value"}) add_dense(m, n + ".attention." + q + "_layer", E, h.enc_att_dim, false);
    add_dense(m, n + ".att_proj", E + h.enc_att_dim, E);
    add_ln(m, n + ".layer_norm", E);
    add_ffn(m, n + ".ffn", E, h.enc_ffn);
  }
  add_dense(m, "length_predictor.projection", E, 1);
  m.add_param("posterior.pos_weight", {1});
  add_dense(m, "posterior.prenet.dense1", O, h.posterior_pre_hidden);
  add_dense(m, "posterior.prenet.dense2", h.posterior_pre_hidden, h.posterior_pre_hidden);
  for (int i = 0; i < h.posterior_nblk; ++i)
    add_xblk(m, "posterior.attentions." + std::to_string(i), h.posterior_att_dim, E, h.posterior_ffn);
  add_dense(m, "posterior.mu_projection", h.posterior_att_dim, L);
  add_dense(m, "posterior.logvar_projection", h.posterior_att_dim, L);
  for (int s = 0; s < h.prior_n_blk; ++s) {
    const std::string g = "prior.glow." + std::to_string(s);
    m.add_param(g + ".actnorm.log_scale", {L});
    m.add_param(g + ".actnorm.bias", {L});
    m.add_param(g + ".linear.weight", {L, L});
    const std::string n = g + ".affine_coupling.net";
    m.add_param(n + ".pos_weight", {1});
    add_dense(m, n + ".pre_projection", L / 2, h.prior_att_dim);
    for (int j = 0; j < h.prior_n_tblk; ++j)
      add_xblk(m, n + ".attentions." + std::to_string(j), h.prior_att_dim, E, h.prior_ffn);
    add_dense(m, n + ".log_scale_proj", h.prior_att_dim, L / 2);
    add_dense(m, n + ".shift_proj", h.prior_att_dim, L / 2);
  }
  add_dense(m, "decoder.pre_projection", L, h.dec_att_dim);
  for (int i = 0; i < h.dec_nblk; ++i) add_xblk(m, "decoder.attentions." + std::to_string(i), h.dec_att_dim, E, h.dec_ffn);
  add_dense(m, "decoder.out_projection", h.dec_att_dim, O * h.max_reduction_factor);
  {
    int cin = O;
    for (int i = 0; i < h.post_n_conv; ++i) {
      add_conv_bn(m, "decoder.postnet.conv_stack." + std::to_string(i), h.post_kernel, cin, h.post_filters);
      cin = h.post_filters;
    }
  }
  add_dense(m, "decoder.residual_projection", h.post_filters, O);

  // ---- packed operand plan
  {
    int cin = h.embd_dim;
    for (int i = 0; i < h.enc_n_conv; ++i) {
      plan_conv(m, "enc.conv" + std::to_string(i), "text_encoder.prenet.conv_stack." + std::to_string(i),
                h.enc_conv_kernel, cin, E, false);
      cin = E;
    }
  }
  m.add_mat("enc.proj", E, E);
  m.add_op("enc.proj", 0, 0, "text_encoder.prenet.projection.kernel", 0, E, E, E, 0);
  for (int i = 0; i < h.enc_n_blk; ++i) {
    const std::string n = "text_encoder.self_attentions." + std::to_string(i), pk = "enc.blk" + std::to_string(i);
    const int A = h.enc_att_dim;
    m.add_mat(pk + ".qkv", 3 * A, E);
    int c = 0;
    for (const char* q : {"query", "key", "value"})
      m.add_op(pk + ".qkv", (c++) * A, 0, n + ".attention." + q + "_layer.kernel", 0, E, A, A, 0);
    m.add_mat(pk + ".proj", E, E + A);
    m.add_op(pk + ".proj", 0, 0, n + ".att_proj.kernel", 0, E + A, E, E, 0);
    m.add_mat(pk + ".ffn1", h.enc_ffn, E);
    m.add_op(pk + ".ffn1", 0, 0, n + ".ffn.dense1.kernel", 0, E, h.enc_ffn, h.enc_ffn, 0);
    m.add_mat(pk + ".ffn2", E, h.enc_ffn);
    m.add_op(pk + ".ffn2", 0, 0, n + ".ffn.dense2.kernel", 0, h.enc_ffn, E, E, 0);
  }
  // posterior
  m.add_mat("post.pre1", h.posterior_pre_hidden, kpad64(O));
  m.add_op("post.pre1", 0, 0, "posterior.prenet.dense1.kernel", 0, O, h.posterior_pre_hidden, h.posterior_pre_hidden, 0);
  m.add_mat("post.pre2", h.posterior_pre_hidden, h.posterior_pre_hidden);
  m.add_op("post.pre2", 0, 0, "posterior.prenet.dense2.kernel", 0, h.posterior_pre_hidden, h.posterior_pre_hidden,
           h.posterior_pre_hidden, 0);
  {
    std::vector<std::string> blks;
    for (int i = 0; i < h.posterior_nblk; ++i) {
      const std::string n = "posterior.attentions." + std::to_string(i);
      plan_xblk(m, "post.blk" + std::to_string(i), n, h.posterior_att_dim, h.posterior_ffn);
      blks.push_back(n);
    }
    plan_memkv(m, "post.kv", blks, E, h.posterior_att_dim);
  }
  // models/models.py:136 swaps the names: mu_projection's output is used as the LOG-VARIANCE, logvar_projection's
  // as the MEAN.  Packed order = [log-variance | mean] as EPI_POSTERIOR expects.
  m.add_mat("post.out", 2 * L, h.posterior_att_dim);
  m.add_op("post.out", 0, 0, "posterior.mu_projection.kernel", 0, h.posterior_att_dim, L, L, 0);
  m.add_op("post.out", L, 0, "posterior.logvar_projection.kernel", 0, h.posterior_att_dim, L, L, 0);
  m.add_vec("post.out.bias", 2 * L);
  // prior
  {
    std::vector<std::string> blks;
    for (int s = 0; s < h.prior_n_blk; ++s) {
      const std::string n = "prior.glow." + std::to_string(s) + ".affine_coupling.net", pk = "prior." + std::to_string(s);
      m.add_mat(pk + ".pre", h.prior_att_dim, kpad64(L / 2));
      m.add_op(pk + ".pre", 0, 0, n + ".pre_projection.kernel", 0, L / 2, h.prior_att_dim, h.prior_att_dim, 0);
      for (int j = 0; j < h.prior_n_tblk; ++j) {
        plan_xblk(m, pk + ".blk" + std::to_string(j), n + ".attentions." + std::to_string(j), h.prior_att_dim, h.prior_ffn);
        blks.push_back(n + ".attentions." + std::to_string(j));
      }
      m.add_mat(pk + ".out", L, h.prior_att_dim);
      m.add_op(pk + ".out", 0, 0, n + ".log_scale_proj.kernel", 0, h.prior_att_dim, L / 2, L / 2, 0);
      m.add_op(pk + ".out", L / 2, 0, n + ".shift_proj.kernel", 0, h.prior_att_dim, L / 2, L / 2, 0);
      m.add_vec(pk + ".out.bias", L);
    }
    plan_memkv(m, "prior.kv", blks, E, h.prior_att_dim);
  }
  // decoder
  m.add_mat("dec.pre", h.dec_att_dim, L);
  m.add_op("dec.pre", 0, 0, "decoder.pre_projection.kernel", 0, L, h.dec_att_dim, h.dec_att_dim, 0);
  {
    std::vector<std::string> blks;
    for (int i = 0; i < h.dec_nblk; ++i) {
      const std::string n = "decoder.attentions." + std::to_string(i);
      plan_xblk(m, "dec.blk" + std::to_string(i), n, h.dec_att_dim, h.dec_ffn);
      blks.push_back(n);
    }
    plan_memkv(m, "dec.kv", blks, E, h.dec_att_dim);
  }
  m.add_mat("dec.out", O * h.max_reduction_factor, h.dec_att_dim);
  m.add_op("dec.out", 0, 0, "decoder.out_projection.kernel", 0, h.dec_att_dim, O * h.max_reduction_factor,
           O * h.max_reduction_factor, 0);
  {
    int cin = O;
    for (int i = 0; i < h.post_n_conv; ++i) {
      plan_conv(m, "dec.post" + std::to_string(i), "decoder.postnet.conv_stack." + std::to_string(i), h.post_kernel, cin,
                h.post_filters, i >= post_plain());
      cin = h.post_filters;
    }
  }
  m.add_mat("dec.res", O, 3 * h.post_filters);
  for (int p = 0; p < 3; ++p)
    m.add_op("dec.res", 0, p * h.post_filters, "decoder.residual_projection.kernel", 0, h.post_filters, O, O, p == 2 ? 1 : 0);
  // ---- backward-pass operands (training only; same arena)
  {
    int cin = h.embd_dim;
    for (int i = 0; i < h.enc_n_conv; ++i) {
      plan_bw_conv(m, "bw.enc.conv" + std::to_string(i), "text_encoder.prenet.conv_stack." + std::to_string(i),
                   h.enc_conv_kernel, cin, E);
      cin = E;
    }
    plan_bw(m, "bw.enc.proj", "text_encoder.prenet.projection.kernel", E, E);
    for (int i = 0; i < h.enc_n_blk; ++i) {
      const std::string n = "text_encoder.self_attentions." + std::to_string(i), pk = "bw.enc.blk" + std::to_string(i);
      const int A = h.enc_att_dim;
      int c = 0;
      plan_bw(m, pk + ".proj", n + ".att_proj.kernel", E + A, E);
      m.add_mat(pk + ".xq", E, E + 3 * A);
      m.add_op(pk + ".xq", 0, 0, n + ".att_proj.kernel", 0, E, E, E, 2);
      c = 0;
      for (const char* q : {"query", "key", "value"})
        m.add_op(pk + ".xq", 0, E + (c++) * A, n + ".attention." + q + "_layer.kernel", 0, E, A, A, 2);
      plan_bw(m, pk + ".ffn1", n + ".ffn.dense1.kernel", E, h.enc_ffn);
      plan_bw(m, pk + ".ffn2", n + ".ffn.dense2.kernel", h.enc_ffn, E);
    }
    plan_bw(m, "bw.post.pre2", "posterior.prenet.dense2.kernel", h.posterior_pre_hidden, h.posterior_pre_hidden);
    std::vector<std::string> blks;
    for (int i = 0; i < h.posterior_nblk; ++i) {
      const std::string n = "posterior.attentions." + std::to_string(i);
      plan_bw_xblk(m, "post.blk" + std::to_string(i), n, h.posterior_att_dim, h.posterior_ffn);
      blks.push_back(n);
    }
    plan_bw_memkv(m, "bw.post.kv", blks, E, h.posterior_att_dim);
    m.add_mat("bw.post.out", h.posterior_att_dim, 2 * L);
    m.add_op("bw.post.out", 0, 0, "posterior.mu_projection.kernel", 0, h.posterior_att_dim, L, L, 2);
    m.add_op("bw.post.out", 0, L, "posterior.logvar_projection.kernel", 0, h.posterior_att_dim, L, L, 2);
    blks.clear();
    for (int s = 0; s < h.prior_n_blk; ++s) {
      const std::string n = "prior.glow." + std::to_string(s) + ".affine_coupling.net", pk = "prior." + std::to_string(s);
      plan_bw(m, "bw." + pk + ".pre", n + ".pre_projection.kernel", L / 2, h.prior_att_dim);
      for (int j = 0; j < h.prior_n_tblk; ++j) {
        plan_bw_xblk(m, pk + ".blk" + std::to_string(j), n + ".attentions." + std::to_string(j), h.prior_att_dim, h.prior_ffn);
        blks.push_back(n + ".attentions." + std::to_string(j));
      }
      m.add_mat("bw." + pk + ".out", h.prior_att_dim, L);
      m.add_op("bw." + pk + ".out", 0, 0, n + ".log_scale_proj.kernel", 0, h.prior_att_dim, L / 2, L / 2, 2);
      m.add_op("bw." + pk + ".out", 0, L / 2, n + ".shift_proj.kernel", 0, h.prior_att_dim, L / 2, L / 2, 2);
    }
    plan_bw_memkv(m, "bw.prior.kv", blks, E, h.prior_att_dim);
    plan_bw(m, "bw.dec.pre", "decoder.pre_projection.kernel", L, h.dec_att_dim);
    blks.clear();
    for (int i = 0; i < h.dec_nblk; ++i) {
      const std::string n = "decoder.attentions." + std::to_string(i);
      plan_bw_xblk(m, "dec.blk" + std::to_string(i), n, h.dec_att_dim, h.dec_ffn);
      blks.push_back(n);
    }
    plan_bw_memkv(m, "bw.dec.kv", blks, E, h.dec_att_dim);
    plan_bw(m, "bw.dec.out", "decoder.out_projection.kernel", h.dec_att_dim, O * h.max_reduction_factor);
    cin = O;
    for (int i = 0; i < h.post_n_conv; ++i) {
      plan_bw_conv(m, "bw.dec.post" + std::to_string(i), "decoder.postnet.conv_stack." + std::to_string(i), h.post_kernel, cin,
                   h.post_filters);
      cin = h.post_filters;
    }
    plan_bw(m, "bw.dec.res", "decoder.residual_projection.kernel", h.post_filters, O);
  }
  // folded flow maps as split-fp16 tensor-core operands (fused row-kernel tail): per step [hi: 128 x 128 | lo: 128 x 128], K-major
  m.add_mat("flow.f", h.prior_n_blk * 2 * L, L);
  m.add_mat("flow.b", h.prior_n_blk * 2 * L, L);
  // flow constants
  const int S = h.prior_n_blk;
  auto region = [&](int64_t bytes) {
    m.packed_bytes = align_up(m.packed_bytes, 256);
    const int64_t o = m.packed_bytes;
    m.packed_bytes += bytes;
    return o;
  };
  m.off_Mf = region(static_cast<int64_t>(S) * L * L * 4);
  m.off_Mb = region(static_cast<int64_t>(S) * L * L * 4);
  m.off_winv = region(static_cast<int64_t>(S) * L * L * 4);
  m.off_cf = region(S * L * 4);
  m.off_cb = region(S * L * 4);
  m.off_consts = region(S * 2 * 4);
  m.off_logdet64 = region(S * 8);
  m.off_ptrs = region(3 * S * 8);
  m.off_packops = region(static_cast<int64_t>(m.plan.size()) * sizeof(PackOp));
  m.off_packtiles = region(static_cast<int64_t>(m.plan.size() + 1) * sizeof(int));
  m.packed_bytes = align_up(m.packed_bytes, 1024);
}

// ============================================================================ execution context
struct Ctx {
  vaenar_model* m;
  const float* params;
  uint8_t* packed;
  uint8_t* ws;
  int64_t ws_bytes;
  int64_t ws_off = 0, ws_peak = 0;
  bool dry = false;
  cudaStream_t stream;
  // training-mode forward (BatchNorm batch statistics, dropout)
  bool training = false;
  float* params_mut = nullptr;           // same buffer as `params`, writable (BN moving averages, ActNorm init)
  const float* const* masks = nullptr;   // injected dropout keep-masks (host array of device pointers) or null
  int n_masks = 0, mask_cursor = 0;
  uint64_t seed = 0;
  int update_bn = 1;
  cudaStream_t wstream = nullptr;        // when set, run_wgrad launches here after an event recorded on `stream`
  cudaStream_t lane = nullptr;           // train step: stream of the decoder branch (null: everything on `stream`)
  cudaStream_t lane_w = nullptr;         // weight-gradient stream of the decoder branch
  cudaStream_t astream = nullptr;        // when set, the dK/dV kernel of run_attention_bwd launches here (beside dQ)
  cudaStream_t lane_a = nullptr;         // the decoder branch's

  template <typename T>
  T* alloc(int64_t count) {
    ws_off = align_up(ws_off, 1024);
    const int64_t o = ws_off;
    ws_off += count * static_cast<int64_t>(sizeof(T));
    if (ws_off > ws_peak) ws_peak = ws_off;
    if (!dry && ws_off > ws_bytes) VB_THROW("workspace too small: need %lld bytes, have %lld", (long long)ws_off, (long long)ws_bytes);
    return dry ? nullptr : reinterpret_cast<T*>(ws + o);
  }
  const float* P(const std::string& n) const { return dry ? nullptr : params + m->params[m->P(n)].offset; }
  float* PM(const std::string& n) const {
    if (!dry && !params_mut) VB_THROW("writable parameter buffer required (%s)", n.c_str());
    return dry ? nullptr : params_mut + m->params[m->P(n)].offset;
  }
  const __half* W(const std::string& n) const {
    auto it = m->pmats.find(n);
    if (it == m->pmats.end()) VB_THROW("unknown packed matrix %s", n.c_str());
    return dry ? nullptr : reinterpret_cast<const __half*>(packed + it->second.off);
  }
  const PackedMat& WM(const std::string& n) const { return m->pmats.at(n); }
  float* V(const std::string& n) const {
    auto it = m->pvecs.find(n);
    if (it == m->pvecs.end()) VB_THROW("unknown packed vector %s", n.c_str());
    return dry ? nullptr : reinterpret_cast<float*>(packed + it->second);
  }
};

// ---- launch accounting + optional per-launch timing (bench.py roofline; never active inside graph capture)
struct KernelStat {
  double flops = 0;     // algorithmic FLOPs of the instrumented launches
  double bytes = 0;     // algorithmic HBM bytes (operands read once + outputs written once)
  long launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};
static long g_launch_count = 0;
static bool g_use_fused = true;   // fused per-CrossAttentionBLK row kernel (VAENAR_FUSED=0 selects the per-op launch chain)
static bool g_use_pdl = true;   // programmatic dependent launch between the tensor-core kernels (VAENAR_NO_PDL=1 disables)
static unsigned long long* g_dbg_cursor = nullptr;   // tuning aid: per-CTA phase timestamps of GEMM launches
static unsigned long long* g_dbg_base = nullptr;
static std::vector<std::string> g_dbg_launches;
static bool g_dbg_attn = true;
static unsigned long long* g_xrow_dbg = nullptr;   // tuning aid: per-CTA phase timestamps of the fused row kernel (128 x u64 per CTA)
static bool g_profile = false;
static std::map<std::string, KernelStat> g_stats;
static std::string g_profile_json;

static void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) VB_THROW("launch of %s failed: %s", what, cudaGetErrorString(e));
  ++g_launch_count;
}
struct ProfileScope {
  KernelStat* st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t stream;
  ProfileScope(const std::string& cls, double flops, double bytes, cudaStream_t s) : stream(s) {
    if (!g_profile) return;
    st = &g_stats[cls];
    st->flops += flops;
    st->bytes += bytes;
    st->launches += 1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
  ~ProfileScope() {
    if (!st) return;
    cudaEventRecord(e1, stream);
    st->events.emplace_back(e0, e1);
  }
};

#define VB_GEMM_INSTANCES(X)                                                                       \
  X(128, EPI_PLAIN, F_RUNTIME) X(256, EPI_PLAIN, F_RUNTIME)                                          \
  X(128, EPI_PLAIN, F_OUT_H) X(128, EPI_PLAIN, F_BIAS | F_RELU | F_OUT_H) X(128, EPI_PLAIN, F_MASK | F_OUT_H)  \
  X(128, EPI_PLAIN, F_BIAS | F_TABLE | F_OUT_F32 | F_OUT_H)                                           \
  X(128, EPI_PLAIN, F_BIAS | F_RELU | F_BN | F_OUT_H)                                                 \
  X(128, EPI_PLAIN, F_BIAS | F_TANH | F_BN | F_OUT_H | F_OUT_LO)                                      \
  X(128, EPI_PLAIN, F_BIAS | F_BN | F_OUT_H | F_OUT_LO)                                               \
  X(256, EPI_PLAIN, F_BIAS | F_TANH | F_BN | F_OUT_H | F_OUT_LO)                                      \
  X(256, EPI_PLAIN, F_BIAS | F_BN | F_OUT_H | F_OUT_LO)                                               \
  X(128, EPI_PLAIN, F_BIAS | F_OUT_F32 | F_OUT_H)                                                     \
  X(128, EPI_PLAIN, F_BIAS | F_RELU | F_TABLE | F_OUT_F32 | F_OUT_H)                                  \
  X(128, EPI_PLAIN, F_BIAS | F_OUT_F32 | F_OUT_H | F_OUT_LO)                                          \
  X(128, EPI_PLAIN, F_BIAS | F_RES | F_OUT_F32)                                                       \
  X(128, EPI_PLAIN, F_RES | F_OUT_F32) X(128, EPI_PLAIN, F_OUT_F32)                                   \
  X(128, EPI_LN, 0) X(256, EPI_LN, 0)                                                                \
  X(256, EPI_QKV, F_OUT_H) X(128, EPI_COUPLING, 0) X(256, EPI_POSTERIOR, 0)

// two-CTAs-per-SM instances (GemmCfg<128, 2>): the plain epilogues of the training step's forward / dgrad GEMMs
#define VB_GEMM_OCC2_INSTANCES(X)                                                                  \
  X(128, EPI_PLAIN, F_OUT_H) X(128, EPI_PLAIN, F_BIAS | F_RELU | F_OUT_H) X(128, EPI_PLAIN, F_MASK | F_OUT_H)  \
  X(128, EPI_PLAIN, F_RES | F_OUT_F32) X(128, EPI_PLAIN, F_OUT_F32)                                   \
  X(128, EPI_PLAIN, F_BIAS | F_RES | F_OUT_F32) X(128, EPI_PLAIN, F_BIAS | F_OUT_F32 | F_OUT_H)
static int g_train_fused = -1;   // vaenar_set_train_fused / VAENAR_TRAIN_FUSED: -1 auto, 0 per-op chain, 1 fused row kernel
static int g_gemm_occ2 = 1;   // VAENAR_GEMM_OCC2=0: one CTA per SM everywhere (A/B timing)

static void set_attrs(vaenar_model* m) {
  static bool done = false;
  if (done) return;
  if (const char* e = getenv("VAENAR_NO_PDL")) g_use_pdl = !(e[0] == '1');
  if (const char* e = getenv("VAENAR_FUSED")) g_use_fused = !(e[0] == '0');
  VB_CUDA(cudaFuncSetAttribute(xblk_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XR_SMEM));
#define VB_SET_ATTR(BN, MODE, FEAT)                                                                  \
  VB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, (FEAT)>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               GemmCfg<BN>::kSmemBytes));
  VB_GEMM_INSTANCES(VB_SET_ATTR)
#undef VB_SET_ATTR
  if (const char* e = getenv("VAENAR_GEMM_OCC2")) g_gemm_occ2 = atoi(e);
#define VB_SET_ATTR2(BN, MODE, FEAT)                                                                 \
  VB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, (FEAT), 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               GemmCfg<BN, 2>::kSmemBytes));                                         \
  VB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, (FEAT), 2>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                               cudaSharedmemCarveoutMaxShared));
  VB_GEMM_OCC2_INSTANCES(VB_SET_ATTR2)
#undef VB_SET_ATTR2
  VB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attention2_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attention2_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attention3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT3_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_DKDV_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_DQ_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB2_DKDV_SMEM));
  VB_CUDA(cudaFuncSetAttribute(attn_bwd_dq2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB2_DQ_SMEM));
  VB_CUDA(cudaFuncSetAttribute(flow_param_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (2 * FLOW_DIM * (FLOW_DIM + 1) + FLOW_DIM) * 4));
  VB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<128>::kSmem));
  VB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<256>::kSmem));
  VB_CUDA(cudaFuncSetAttribute(slogdet128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOW_DIM * (FLOW_DIM + 1) * 8));
  VB_CUDA(cudaFuncSetAttribute(inverse128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (FLOW_DIM * (2 * FLOW_DIM + 1) + FLOW_DIM) * 4));
  VB_CUDA(cudaFuncSetAttribute(flow_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOW_LIN_SMEM));
  done = true;
  (void)m;
}

// ============================================================================ launch helpers
struct AOp {
  const __half* p = nullptr;
  int ld = 0;   // row stride in elements
  int K = 0;    // valid inner extent
};

static void segs_plain(GemmParams& p, int K) {
  p.nseg = 1;
  p.seg_map[0] = 0; p.seg_shift[0] = 0; p.seg_kblocks[0] = cdiv(K, 64);
  p.alg_k = K;
}
static void segs_concat(GemmParams& p, int K0, int K1) {
  p.nseg = 2;
  p.seg_map[0] = 0; p.seg_shift[0] = 0; p.seg_kblocks[0] = cdiv(K0, 64);
  p.seg_map[1] = 1; p.seg_shift[1] = 0; p.seg_kblocks[1] = cdiv(K1, 64);
  p.alg_k = K0 + K1;
}
// taps x (hi | lo | hi) when split; row shift j - (taps-1)/2  (Keras 'same' padding, stride 1)
static void segs_conv(GemmParams& p, int taps, int cin, bool split) {
  const int parts = split ? 3 : 1;
  p.nseg = taps * parts;
  p.alg_k = taps * cin;
  for (int j = 0; j < taps; ++j)
    for (int q = 0; q < parts; ++q) {
      const int s = j * parts + q;
      p.seg_map[s] = (split && q == 1) ? 1 : 0;
      p.seg_shift[s] = j - (taps - 1) / 2;
      p.seg_kblocks[s] = cdiv(cin, 64);
    }
}

// A operands: [batches, rows, K]; W: packed [Nrows, Ktot]; output tile selection by block_n.
// W: [Nrows, Ktot] K-major with row pitch w_ld (0: Ktot); a K extent that is not a multiple of 64 is zero-filled by TMA.
static void run_gemm(Ctx& c, int block_n, AOp a0, AOp a1, int batches, int rows, const __half* W, int Ktot, int Nrows,
                     GemmParams p, int w_ld = 0) {
  if (c.dry) return;
  int tk = 0;
  for (int s = 0; s < p.nseg; ++s) tk += p.seg_kblocks[s];
  if (tk * 64 < Ktot || tk * 64 - Ktot >= 64) VB_THROW("gemm: segments cover %d of K %d", tk * 64, Ktot);
  if (p.mode == EPI_LN && !(p.bias && p.residual && p.out_f32 && p.out_h && p.ln_gamma && p.ln_beta))
    VB_THROW("gemm: the LayerNorm epilogue needs bias, residual, gamma, beta and both outputs");
  // LayerNorm over N = 2 * BLOCK_N: split the columns over a 2-CTA cluster (row statistics via DSMEM)
  p.ln_cluster = (p.mode == EPI_LN && p.N == 2 * block_n) ? 1 : 0;
  if ((p.mode == EPI_LN && !p.ln_cluster && p.N != block_n) ||
      ((p.mode == EPI_COUPLING || p.mode == EPI_POSTERIOR) && p.N != block_n))
    VB_THROW("gemm: row-wise epilogue needs BLOCK_N == N (%d vs %d)", block_n, p.N);
  p.batches = batches;
  p.rows = rows;
  p.tiles_per_batch = cdiv(rows, GEMM_BLOCK_M);
  if (p.seq_T <= 0) { p.seq_T = rows; p.seq_B = batches; }
  if (!a1.p) a1 = a0;
  const CUtensorMap tA0 = make_tmap(a0.p, 3, a0.K, rows, batches, a0.ld, static_cast<uint64_t>(rows) * a0.ld, 64, 128);
  const CUtensorMap tA1 = make_tmap(a1.p, 3, a1.K, rows, batches, a1.ld, static_cast<uint64_t>(rows) * a1.ld, 64, 128);
  const CUtensorMap tB = make_tmap(W, 2, Ktot, Nrows, 1, w_ld > 0 ? w_ld : Ktot, 0, 64, block_n > 256 ? 256 : block_n);
  dim3 grid(batches * p.tiles_per_batch, cdiv(p.N, block_n));
  const double Mrows = static_cast<double>(batches) * rows;
  const double kalg = p.alg_k > 0 ? p.alg_k : Ktot;   // algorithmic K: no padding, no split-fp16 triple
  if (g_dbg_cursor) {
    if (g_dbg_launches.empty()) g_dbg_base = g_dbg_cursor;
    p.dbg = g_dbg_cursor;
    char buf[256];
    snprintf(buf, sizeof(buf), "{\"grid\": [%u, %u], \"mode\": %d, \"N\": %d, \"K\": %d, \"bn\": %d, \"offset\": %lld}", grid.x, grid.y, p.mode, p.N, Ktot,
             block_n, static_cast<long long>(g_dbg_cursor - g_dbg_base));
    g_dbg_launches.push_back(buf);
    g_dbg_cursor += static_cast<size_t>(grid.x) * grid.y * 8;
  }
  const char* cls = p.mode == EPI_LN ? "gemm_ln" : (p.nseg >= 5 ? "gemm_conv" : (p.mode == EPI_QKV ? "gemm_qkv" : "gemm_plain"));
  ProfileScope prof(cls, 2.0 * Mrows * p.N * kalg, Mrows * kalg * 2 + static_cast<double>(p.N) * Ktot * 2 + Mrows * p.N * 6, c.stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(gemm_threads(p.mode));
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = p.ln_cluster ? 2 : 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: the kernel calls griddepcontrol.wait
  attr[1].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t le = cudaErrorInvalidValue;
  bool launched = false;
  // feature mask of this call (EPI_PLAIN); a specialised instance is used when one exists, else the run-time one
  uint32_t feat = 0;
  if (p.mode == EPI_PLAIN) {
    feat = (p.bias ? F_BIAS : 0) | (p.act == 1 ? F_RELU : 0) | (p.act == 2 ? F_TANH : 0) | (p.ch_scale ? F_BN : 0) |
           (p.add_table ? F_TABLE : 0) | (p.residual ? F_RES : 0) | (p.out_f32 ? F_OUT_F32 : 0) | (p.out_h ? F_OUT_H : 0) |
           (p.out_lo ? F_OUT_LO : 0) | (p.mask_h ? F_MASK : 0);
    if (p.mask_h && (p.residual || p.out_f32)) VB_THROW("gemm: the output mask excludes the residual and the fp32 output");
  }
#define VB_TRY(BN, MODE, FEAT)                                                                              \
  if (!launched && block_n == BN && p.mode == MODE && (MODE != EPI_PLAIN || feat == static_cast<uint32_t>(FEAT))) { \
    cfg.dynamicSmemBytes = GemmCfg<BN>::kSmemBytes;                                                         \
    le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, (FEAT)>, tA0, tA1, tB, p);                       \
    launched = true;                                                                                        \
  }
#define VB_TRY_RT(BN, MODE, FEAT)                                                                           \
  if (!launched && block_n == BN && p.mode == MODE && MODE == EPI_PLAIN && ((FEAT) & F_RUNTIME)) {          \
    cfg.dynamicSmemBytes = GemmCfg<BN>::kSmemBytes;                                                         \
    le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, (FEAT)>, tA0, tA1, tB, p);                       \
    launched = true;                                                                                        \
  }
  // multi-wave grids of the plain epilogues: two CTAs per SM (the epilogue of one under the mainloop of the other)
#define VB_TRY2(BN, MODE, FEAT)                                                                             \
  if (!launched && block_n == BN && p.mode == MODE && feat == static_cast<uint32_t>(FEAT)) {                \
    cfg.dynamicSmemBytes = GemmCfg<BN, 2>::kSmemBytes;                                                      \
    le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MODE, (FEAT), 2>, tA0, tA1, tB, p);                    \
    launched = true;                                                                                        \
  }
  if (g_gemm_occ2 && p.mode == EPI_PLAIN && !p.dbg &&
      static_cast<long>(grid.x) * grid.y > (g_gemm_occ2 > 1 ? 0 : 148)) {
    VB_GEMM_OCC2_INSTANCES(VB_TRY2)
  }
#undef VB_TRY2
  VB_GEMM_INSTANCES(VB_TRY)
  VB_GEMM_INSTANCES(VB_TRY_RT)
#undef VB_TRY
#undef VB_TRY_RT
  if (!launched) VB_THROW("no gemm_tc_kernel instance for BLOCK_N %d mode %d", block_n, p.mode);
  if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(gemm_tc_kernel<%d>) failed: %s", block_n, cudaGetErrorString(le));
  check_launch("gemm_tc_kernel");
}

static GemmParams gp() {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.ln_eps = 1e-3f;   // Keras LayerNormalization default epsilon
  return p;
}

// ---- weight gradient: dW[M, N] (+)= X^T dY over tokens (csrc/wgrad_tc.cuh).  X / dY are row-major fp16
// [batches, rows, ld] tensors; `x1` (optional) supplies the dW rows >= a_split (Dense over a concat).
struct WOp {
  const __half* p = nullptr;
  int ld = 0;     // row pitch in elements
  int C = 0;      // channel extent of the tensor map (columns >= C read as zero)
  int col0 = 0;   // first channel of the operand window
};
static cudaEvent_t next_event(vaenar_model* m) {
  if (m->ev_cursor == m->ev_pool.size()) {
    cudaEvent_t e;
    VB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    m->ev_pool.push_back(e);
  }
  return m->ev_pool[m->ev_cursor++];
}
// everything queued on the weight-gradient stream so far is ordered before what follows on the main stream
static void join_wgrad(Ctx& c) {
  if (c.dry || !c.wstream) return;
  cudaEvent_t e = next_event(c.m);
  VB_CUDA(cudaEventRecord(e, c.wstream));
  VB_CUDA(cudaStreamWaitEvent(c.stream, e, 0));
}
static void run_wgrad(Ctx& c, WOp x0, WOp x1, int a_split, WOp dy, int batches, int rows, int shift, int M, int N,
                      float* out, int ldo, const std::vector<float*>* outs = nullptr, int n_per_out = 0) {
  if (c.dry) return;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.batches = batches; p.rows = rows;
  p.kb_per_batch = cdiv(rows, WG_BLOCK_K);
  p.total_kb = batches * p.kb_per_batch;
  // 128 x 256 tiles (VAENAR_WGRAD_BN=256; less operand traffic per MMA) measured SLOWER than 128 x 128 at C3 (train step
  // 11.87 vs 11.74 ms): half as many CTAs per split-K wave outweighs the 25 % smaller fill -- kept for A/B timing only
  static const int wg_bn_env = getenv("VAENAR_WGRAD_BN") ? atoi(getenv("VAENAR_WGRAD_BN")) : 0;
  const int bn = (wg_bn_env == 256 && N % 256 == 0) ? 256 : 128;
  const int tiles = cdiv(M, WG_BLOCK_M) * cdiv(N, bn);
  // About one wave of CTAs, but at least `min_kb` token blocks per split: the step keeps several streams busy, so what counts
  // is the SM-time of a launch (CTAs x (prologue + mainloop + reduction epilogue)), not its latency -- a 256 x 256 gradient
  // cut into 37 splits of 6 token blocks spends most of its SM-time in prologues and atomics.
  static const int min_kb = getenv("VAENAR_WGRAD_MINKB") ? std::max(1, atoi(getenv("VAENAR_WGRAD_MINKB"))) : 16;
  int splits = std::max(1, std::min(std::max(1, p.total_kb / min_kb), std::max(1, 148 / tiles)));
  p.kb_per_split = cdiv(p.total_kb, splits);
  splits = cdiv(p.total_kb, p.kb_per_split);
  p.M = M; p.N = N;
  if (!x1.p) { x1 = x0; a_split = M; }
  p.a_split = a_split; p.a_shift = shift;
  p.a0_col0 = x0.col0; p.a1_col0 = x1.col0; p.b_col0 = dy.col0;
  p.out = out; p.ldo = ldo;
  if (outs) {
    if (outs->size() > 24 || n_per_out <= 0) VB_THROW("wgrad: too many output tensors");
    p.n_per_out = n_per_out;
    for (size_t i = 0; i < outs->size(); ++i) p.outs[i] = (*outs)[i];
  }
  if (a_split % WG_BLOCK_M != 0 && a_split != M) VB_THROW("wgrad: concat boundary %d must be a multiple of %d", a_split, WG_BLOCK_M);
  const CUtensorMap tA0 = make_tmap(x0.p, 3, x0.C, rows, batches, x0.ld, static_cast<uint64_t>(rows) * x0.ld, 64, 64);
  const CUtensorMap tA1 = make_tmap(x1.p, 3, x1.C, rows, batches, x1.ld, static_cast<uint64_t>(rows) * x1.ld, 64, 64);
  const CUtensorMap tB = make_tmap(dy.p, 3, dy.C, rows, batches, dy.ld, static_cast<uint64_t>(rows) * dy.ld, 64, 64);
  const double tokens = static_cast<double>(batches) * rows;
  ProfileScope prof("wgrad", 2.0 * tokens * M * N, tokens * (M + N) * 2 + static_cast<double>(M) * N * 4,
                    c.wstream ? c.wstream : c.stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cdiv(M, WG_BLOCK_M), cdiv(N, bn), splits);
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = bn == 256 ? WgCfg<256>::kSmem : WgCfg<128>::kSmem;
  cfg.stream = c.stream;
  if (c.wstream) {   // off the critical path: wait for the producer of dY, then run on the weight-gradient stream
    cudaEvent_t e = next_event(c.m);
    VB_CUDA(cudaEventRecord(e, c.stream));
    VB_CUDA(cudaStreamWaitEvent(c.wstream, e, 0));
    cfg.stream = c.wstream;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_use_pdl && !c.wstream) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t le = bn == 256 ? cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<256>, tA0, tA1, tB, p)
                                   : cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<128>, tA0, tA1, tB, p);
  if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(wgrad_tc_kernel) failed: %s", cudaGetErrorString(le));
  check_launch("wgrad_tc_kernel");
}

struct AttnCall {
  const __half* q; int q_ld; int q_col0; int Tq;
  const __half* k; int k_ld; int k_col0; int Tk;
  const __half* vt; long vt_rows; int vt_ld; long vt_row0;
  const int* q_len; const int* k_len;
  int causal;
  __half* ctx; int ctx_ld;
  float* ali;
  float* lse2 = nullptr;   // [B, H, Tq] softmax statistic saved for the backward pass (training)
};
static void run_attention(Ctx& c, int B, int H, const AttnCall& a) {
  if (c.dry) return;
  AttnParams p;
  p.B = B; p.H = H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.vt_row0 = a.vt_row0;
  p.vt = a.vt; p.vt_ld = a.vt_ld;
  p.q_len = a.q_len; p.k_len = a.k_len; p.causal = a.causal;
  p.scale = 1.0f / sqrtf(static_cast<float>(ATT_D));   // attention.py:227-229, temperature 1.0
  p.ctx = a.ctx; p.ctx_ld = a.ctx_ld; p.ali = a.ali; p.lse2 = a.lse2;
  p.dbg = nullptr;
  if (g_dbg_cursor && g_dbg_attn) {
    if (g_dbg_launches.empty()) g_dbg_base = g_dbg_cursor;
    p.dbg = g_dbg_cursor;
    const unsigned gx = cdiv(a.Tq, ATT_BQ);
    char buf[256];
    snprintf(buf, sizeof(buf), "{\"grid\": [%u, %d], \"mode\": %d, \"N\": %d, \"K\": %d, \"bn\": %d, \"offset\": %lld}", gx, H * B,
             a.causal ? 100 : (a.ali ? 102 : 101), a.Tq, a.Tk, 0, static_cast<long long>(g_dbg_cursor - g_dbg_base));
    g_dbg_launches.push_back(buf);
    g_dbg_cursor += static_cast<size_t>(gx) * H * B * 8;
  }
  const CUtensorMap tQ = make_tmap(a.q, 3, a.q_ld, a.Tq, B, a.q_ld, static_cast<uint64_t>(a.Tq) * a.q_ld, 64, 128);
  const CUtensorMap tK = make_tmap(a.k, 3, a.k_ld, a.Tk, B, a.k_ld, static_cast<uint64_t>(a.Tk) * a.k_ld, 64, 128);
  const CUtensorMap tV = make_tmap(a.vt, 2, a.Tk, a.vt_rows, 1, a.vt_ld, 0, 64, 64);
  dim3 grid(cdiv(a.Tq, ATT_BQ), H, B);
  const double work = static_cast<double>(B) * H * a.Tq * a.Tk;
  ProfileScope prof(a.causal ? "attn_self" : (a.ali ? "attn_cross_ali" : "attn_cross"), 4.0 * work * ATT_D,
                    static_cast<double>(B) * H * ATT_D * 2 * (2.0 * a.Tq + 2.0 * a.Tk) + (a.ali ? work * 4 : 0), c.stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  // Two variants: 128-key blocks / one CTA per SM (lowest latency when the grid fits the chip in one wave: the inference
  // chains), 64-key blocks / two CTAs per SM (more overlap when the grid is several waves deep: the B32 training step).
  // Measured: C2 inference 1.995 ms (v1) vs 2.022 ms (v2); C3 train step 14.30 ms (v1) vs 14.14 ms (v2).
  static const char* att_env = getenv("VAENAR_ATTN");
  // v3: causal self-attention with the scores resident in TMEM (one tensor-core pass), T <= 448; also the training forward
  if ((!att_env || att_env[0] == '3') && a.causal && a.q_len == a.k_len && a.Tq == a.Tk && a.Tk <= AT3_TMAX && !a.ali) {
    cfg.gridDim = dim3(H * B, cdiv(a.Tq, ATT_BQ));
    cfg.blockDim = dim3(AT3_THREADS);
    cfg.dynamicSmemBytes = AT3_SMEM;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr3[1];
    attr3[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr3[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr3;
    cfg.numAttrs = 1;
    const cudaError_t le3 = cudaLaunchKernelEx(&cfg, attention3_tc_kernel, tQ, tK, tV, p);
    if (le3 != cudaSuccess) VB_THROW("cudaLaunchKernelEx(attention3_tc_kernel) failed: %s", cudaGetErrorString(le3));
    check_launch("attention3_tc_kernel");
    return;
  }
  const bool att_v1 = att_env ? att_env[0] == '1' : (grid.x * grid.y * grid.z <= 2u * 148u);
  cfg.blockDim = dim3(att_v1 ? ATT_THREADS : ATT2_THREADS);
  cfg.dynamicSmemBytes = att_v1 ? ATT_SMEM : ATT2_SMEM;
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CUtensorMap tK64 = make_tmap(a.k, 3, a.k_ld, a.Tk, B, a.k_ld, static_cast<uint64_t>(a.Tk) * a.k_ld, 64, 64);
  const cudaError_t le =
      att_v1 ? (a.ali ? cudaLaunchKernelEx(&cfg, attention_tc_kernel<true>, tQ, tK, tV, p)
                      : cudaLaunchKernelEx(&cfg, attention_tc_kernel<false>, tQ, tK, tV, p))
             : (a.ali ? cudaLaunchKernelEx(&cfg, attention2_tc_kernel<true>, tQ, tK64, tV, p)
                      : cudaLaunchKernelEx(&cfg, attention2_tc_kernel<false>, tQ, tK64, tV, p));
  if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(attention_tc_kernel) failed: %s", cudaGetErrorString(le));
  check_launch("attention_tc_kernel");
}

// ---- attention backward (csrc/attention_bwd_tc.cuh).  All tensors row-major fp16 [B*T, ld]; head h at col0 + h*64.
struct AttnBwdCall {
  const __half* q; int q_ld, q_col0; int Tq;
  const __half* k; int k_ld, k_col0; int Tk;
  const __half* v; int v_ld, v_col0;
  const __half* o; int o_ld;                 // forward context [B*Tq, H*64]
  const __half* dO; int do_ld, do_col0;
  const int* q_len; const int* k_len; int causal;
  const float* lse2;
  float* delta;                              // scratch [B, H, Tq]
  __half* dq; int dq_ld, dq_col0;
  __half* dk; int dk_ld, dk_col0;
  __half* dv; int dv_ld, dv_col0;
};
// everything queued on the dK/dV stream so far is ordered before what follows on the main stream
static void join_attn(Ctx& c) {
  if (c.dry || !c.astream) return;
  cudaEvent_t e = next_event(c.m);
  VB_CUDA(cudaEventRecord(e, c.astream));
  VB_CUDA(cudaStreamWaitEvent(c.stream, e, 0));
}
// dq (main stream) and dk / dv (c.astream when set: the two kernels only share inputs) from dO.  join_dkdv: dk / dv are
// complete on the main stream when the call returns; otherwise the caller orders its consumers with join_attn().
static void run_attention_bwd(Ctx& c, int B, int H, const AttnBwdCall& a, bool join_dkdv = true) {
  if (c.dry) return;
  {
    const long rows = static_cast<long>(B) * a.Tq;
    attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, c.stream>>>(a.dO, a.do_ld, a.do_col0, a.o, a.o_ld, a.delta,
                                                                                B, a.Tq, H);
    check_launch("attn_delta");
  }
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0; p.do_col0 = a.do_col0;
  p.q_len = a.q_len; p.k_len = a.k_len; p.causal = a.causal;
  p.scale = 1.0f / sqrtf(static_cast<float>(ATT_D));
  p.lse2 = a.lse2; p.delta = a.delta;
  p.dq = a.dq; p.dq_ld = a.dq_ld; p.dq_col0 = a.dq_col0;
  p.dk = a.dk; p.dk_ld = a.dk_ld; p.dk_col0 = a.dk_col0;
  p.dv = a.dv; p.dv_ld = a.dv_ld; p.dv_col0 = a.dv_col0;
  const CUtensorMap tQ = make_tmap(a.q, 3, a.q_ld, a.Tq, B, a.q_ld, static_cast<uint64_t>(a.Tq) * a.q_ld, 64, 128);
  const CUtensorMap tK = make_tmap(a.k, 3, a.k_ld, a.Tk, B, a.k_ld, static_cast<uint64_t>(a.Tk) * a.k_ld, 64, 128);
  const CUtensorMap tV = make_tmap(a.v, 3, a.v_ld, a.Tk, B, a.v_ld, static_cast<uint64_t>(a.Tk) * a.v_ld, 64, 128);
  const CUtensorMap tO = make_tmap(a.dO, 3, a.do_ld, a.Tq, B, a.do_ld, static_cast<uint64_t>(a.Tq) * a.do_ld, 64, 128);
  const double work = static_cast<double>(B) * H * a.Tq * a.Tk;
  static const bool v1 = getenv("VAENAR_ATTN_BWD_V1") != nullptr;   // 128-wide tiles, one CTA per SM (kept for A/B timing)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(v1 ? ATB_THREADS : ATB2_THREADS);
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (c.astream) {   // inputs (incl. delta) are final at this point of the main stream
    cudaEvent_t e = next_event(c.m);
    VB_CUDA(cudaEventRecord(e, c.stream));
    VB_CUDA(cudaStreamWaitEvent(c.astream, e, 0));
    cfg.stream = c.astream;
  }
  {
    ProfileScope prof(a.causal ? "attn_bwd_dkdv_self" : "attn_bwd_dkdv_cross", 8.0 * work * ATT_D,
                      static_cast<double>(B) * H * ATT_D * 2 * (2.0 * a.Tq + 4.0 * a.Tk), cfg.stream);
    cfg.gridDim = dim3(cdiv(a.Tk, 128), H, B);
    cfg.dynamicSmemBytes = v1 ? ATB_DKDV_SMEM : ATB2_DKDV_SMEM;
    const CUtensorMap tQ64 = make_tmap(a.q, 3, a.q_ld, a.Tq, B, a.q_ld, static_cast<uint64_t>(a.Tq) * a.q_ld, 64, 64);
    const CUtensorMap tO64 = make_tmap(a.dO, 3, a.do_ld, a.Tq, B, a.do_ld, static_cast<uint64_t>(a.Tq) * a.do_ld, 64, 64);
    const cudaError_t le = v1 ? cudaLaunchKernelEx(&cfg, attn_bwd_dkdv_kernel, tQ, tK, tV, tO, p)
                              : cudaLaunchKernelEx(&cfg, attn_bwd_dkdv2_kernel, tQ64, tK, tV, tO64, p);
    if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(attn_bwd_dkdv_kernel) failed: %s", cudaGetErrorString(le));
    check_launch("attn_bwd_dkdv_kernel");
  }
  cfg.stream = c.stream;
  {
    ProfileScope prof(a.causal ? "attn_bwd_dq_self" : "attn_bwd_dq_cross", 6.0 * work * ATT_D,
                      static_cast<double>(B) * H * ATT_D * 2 * (3.0 * a.Tq + 2.0 * a.Tk), c.stream);
    cfg.gridDim = dim3(cdiv(a.Tq, 128), H, B);
    cfg.dynamicSmemBytes = v1 ? ATB_DQ_SMEM : ATB2_DQ_SMEM;
    const CUtensorMap tK64 = make_tmap(a.k, 3, a.k_ld, a.Tk, B, a.k_ld, static_cast<uint64_t>(a.Tk) * a.k_ld, 64, 64);
    const CUtensorMap tV64 = make_tmap(a.v, 3, a.v_ld, a.Tk, B, a.v_ld, static_cast<uint64_t>(a.Tk) * a.v_ld, 64, 64);
    const cudaError_t le = v1 ? cudaLaunchKernelEx(&cfg, attn_bwd_dq_kernel, tQ, tK, tV, tO, p)
                              : cudaLaunchKernelEx(&cfg, attn_bwd_dq2_kernel, tQ, tK64, tV64, tO, p);
    if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(attn_bwd_dq_kernel) failed: %s", cudaGetErrorString(le));
    check_launch("attn_bwd_dq_kernel");
  }
  if (join_dkdv) join_attn(c);
}

static void run_cast(Ctx& c, const float* in, __half* out, int64_t n) {
  if (c.dry) return;
  const int64_t thr = (n + 3) / 4;
  cast_f32_to_f16_kernel<<<static_cast<unsigned>((thr + 255) / 256), 256, 0, c.stream>>>(in, out, n);
  check_launch("cast");
}
static void run_pe(Ctx& c, float* out, int T, int D, float step) {
  if (c.dry) return;
  pe_table_kernel<<<T, 256, 0, c.stream>>>(out, T, D, step);
  check_launch("pe_table");
}
static int vt_pad(int T) { return cdiv(T, 64) * 64; }

// Next dropout keep-mask in the reference's call order; injected (parity tests) or generated from (seed, site).
static const float* next_mask(Ctx& c, int64_t n, float rate) {
  const int site = c.mask_cursor++;
  if (c.masks) {
    if (site >= c.n_masks) VB_THROW("dropout mask %d requested but only %d supplied", site, c.n_masks);
    return c.dry ? nullptr : c.masks[site];
  }
  float* mk = c.alloc<float>(n);
  if (!c.dry) {
    const int64_t thr = (n + 3) / 4;
    dropout_mask_kernel<<<static_cast<unsigned>((thr + 255) / 256), 256, 0, c.stream>>>(mk, n, rate, c.seed,
                                                                                     static_cast<uint64_t>(site));
    check_launch("dropout_mask");
  }
  return mk;
}
static void run_dropout(Ctx& c, const float* x, const float* mask, float* out_f32, __half* out_h, int64_t n) {
  if (c.dry) return;
  dropout_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(x, mask, out_f32, out_h, n);
  check_launch("dropout_apply");
}
// x = (x * mask2 + pw * table[t]) * mask3 -> fp32 + fp16  (posterior prenet tail in training mode, posterior.py:117-121)
__global__ void prenet_tail_kernel(float* x, const float* __restrict__ mask2, const float* __restrict__ table,
                                   const float* __restrict__ pw, const float* __restrict__ mask3, int seq_T, int C,
                                   __half* __restrict__ out_h, long n, __half* __restrict__ pre_h = nullptr) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long row = i / C;
  const int c = static_cast<int>(i % C), t = static_cast<int>(row % seq_T);
  if (pre_h) pre_h[i] = __float2half_rn(x[i]);   // post-ReLU, pre-dropout value: sign needed by the backward pass
  const float v = (x[i] * mask2[i] + pw[0] * table[static_cast<long>(t) * C + c]) * mask3[i];
  x[i] = v;
  out_h[i] = __float2half_rn(v);
}

// ============================================================================ blocks
struct Stream2 {   // fp32 residual stream + its fp16 operand copy, [rows, d]
  float* f;
  __half* h;
};
struct MemKV {     // projected memory of all blocks of a module
  __half* k;       // [B*Tt, nblk*d]
  __half* vt;      // [nblk*B*H*64, tpad]
  int nblk, d, tpad;
  int k_ld = 0;    // row pitch of k when it is a window of a wider tensor (training: K | V row-major), 0: nblk*d
};

// K/V projections of the text memory for all blocks of a module, one GEMM (attention.py:219-220 hoisted)
static MemKV memory_kv(Ctx& c, const std::string& wname, const __half* emb_h, int B, int Tt, int E, int nblk, int d, int H) {
  MemKV kv;
  kv.nblk = nblk; kv.d = d; kv.tpad = vt_pad(Tt);
  kv.k = c.alloc<__half>(static_cast<int64_t>(B) * Tt * nblk * d);
  kv.vt = c.alloc<__half>(static_cast<int64_t>(nblk) * B * H * 64 * kv.tpad);
  GemmParams p = gp();
  p.mode = EPI_QKV; p.N = 2 * nblk * d;
  segs_plain(p, E);
  p.seq_T = Tt; p.seq_B = B;
  p.n_rowmajor = nblk * d; p.out_h = kv.k; p.ld_h = nblk * d;
  p.vt = kv.vt; p.vt_ld = kv.tpad; p.heads = H;
  run_gemm(c, 256, AOp{emb_h, E, E}, AOp{}, 1, B * Tt, c.W(wname), E, 2 * nblk * d, p);
  return kv;
}

struct XblkBufs {
  __half* qk;    // [rows, 2d]  self Q | K
  __half* vt;    // [B*H*64, tpad]
  __half* ctx;   // [rows, d]
  __half* q;     // [rows, d]
  __half* hid;   // [rows, ffn]
  Stream2 s, cst;
  int tpad;
};
static XblkBufs xblk_bufs(Ctx& c, int B, int T, int d, int H, int ffn) {
  XblkBufs b;
  const int64_t rows = static_cast<int64_t>(B) * T;
  b.tpad = vt_pad(T);
  b.qk = c.alloc<__half>(rows * 2 * d);
  b.vt = c.alloc<__half>(static_cast<int64_t>(B) * H * 64 * b.tpad);
  b.ctx = c.alloc<__half>(rows * d);
  b.q = c.alloc<__half>(rows * d);
  b.hid = c.alloc<__half>(rows * ffn);
  b.s.f = c.alloc<float>(rows * d); b.s.h = c.alloc<__half>(rows * d);
  b.cst.f = c.alloc<float>(rows * d); b.cst.h = c.alloc<__half>(rows * d);
  return b;
}

// Fused row kernel of one CrossAttentionBLK (csrc/xblk_fused.cuh): everything after the causal self-attention core, plus
// (next_pk non-empty) the q|k|v projections of the NEXT block, in one launch.  b.ctx holds the self-attention context.
static bool xrow_supported(int d, int H, int ffn, int Tt, const MemKV& kv) {
  return g_use_fused && d == XR_D && H == XR_H && ffn == XR_F && kv.d == XR_D && Tt >= 1 && Tt <= XR_TK_MAX;
}
// What follows the last CrossAttentionBLK of a coupling net inside the same launch (csrc/xblk_fused.cuh, "flow tail")
struct XTail {
  bool coupling = false, flow = false, pre = false;
  std::string out_pk;            // packed [log_scale | shift] projection ("prior.s.out") and its bias vector
  float* z = nullptr;            // [rows, 128] flow state
  int zp_off = 0;
  bool backward = false;
  float* row_acc = nullptr;
  const __half* flow_w = nullptr;   // packed hi | lo map of the flow operation that follows
  const float* flow_c = nullptr;
  std::string pre_pk, pre_pn;    // next step's pre-projection ("prior.s.pre", "prior.glow.s.affine_coupling.net")
  int cond_off_next = 0;
  const float* pe = nullptr;
};
// Training forward: where the fused row kernel leaves the tensors the backward pass reads (XblkTape in train.inc)
struct XTapeOut {
  float* o_f = nullptr; __half* o_h = nullptr;   // block output (the input buffers stay intact)
  float* s_f = nullptr; __half* s_h = nullptr; float* rstd1 = nullptr;
  __half* q2 = nullptr; float* lse2 = nullptr; __half* ctx2 = nullptr;
  float* c_f = nullptr; __half* c_h = nullptr; float* rstd2 = nullptr;
  __half* hid = nullptr; float* rstd3 = nullptr;
};
// b.qk: Q | K of the next block, row pitch 2d -- or, with a tape, Q | K | V row-major with pitch 3d
static void run_xblk_row(Ctx& c, const std::string& pk, const std::string& pn, const std::string& next_pk, Stream2 x,
                         const XblkBufs& b, int B, int T, const int* q_len, const MemKV& kv, int kv_blk, int Tt,
                         const int* t_len, float* ali, const XTail* tail = nullptr, const XTapeOut* tape = nullptr) {
  if (c.dry) return;
  const int d = XR_D, H = XR_H;
  XRowParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.Tt = Tt; p.TKP = cdiv(Tt, 16) * 16;
  p.kv_col0 = kv_blk * kv.d;
  p.vt_row0 = kv_blk * B * H * 64;
  p.vt = kv.vt; p.vt_ld = kv.tpad;
  p.q_len = q_len; p.k_len = t_len;
  p.scale = 1.0f / sqrtf(64.f);   // attention.py:227-229, temperature 1.0
  p.ln_eps = 1e-3f;               // Keras LayerNormalization default
  p.vec[0] = c.P(pn + ".att_proj1.bias"); p.vec[1] = c.P(pn + ".layer_norm1.gamma"); p.vec[2] = c.P(pn + ".layer_norm1.beta");
  p.vec[3] = c.P(pn + ".att_proj2.bias"); p.vec[4] = c.P(pn + ".layer_norm2.gamma"); p.vec[5] = c.P(pn + ".layer_norm2.beta");
  p.vec[6] = c.P(pn + ".ffn.dense2.bias"); p.vec[7] = c.P(pn + ".ffn.layer_norm.gamma"); p.vec[8] = c.P(pn + ".ffn.layer_norm.beta");
  p.bf1 = c.P(pn + ".ffn.dense1.bias");
  p.x_f = x.f;
  p.has_next = next_pk.empty() ? 0 : 1;
  p.qk_next = b.qk; p.vt_next = b.vt; p.vt_next_ld = b.tpad;
  p.qk_next_ld = 2 * d;
  p.x_f_out = x.f;
  if (tape) {
    if (tail) VB_THROW("row kernel: the flow tail is an inference-only fusion");
    p.x_f_out = tape->o_f;
    p.tp_s_f = tape->s_f; p.tp_s_h = tape->s_h; p.tp_rstd1 = tape->rstd1;
    p.tp_q2 = tape->q2; p.tp_lse2 = tape->lse2; p.tp_ctx2 = tape->ctx2;
    p.tp_c_f = tape->c_f; p.tp_c_h = tape->c_h; p.tp_rstd2 = tape->rstd2;
    p.tp_hid = tape->hid; p.tp_rstd3 = tape->rstd3;
    if (p.has_next) { p.qk_next_ld = 3 * d; p.v_rm_next = b.qk + 2 * d; }
  }
  p.ali = ali;
  static const bool no_n256 = getenv("VAENAR_XR_N128") != nullptr;   // A/B timing: two N = 128 instructions per weight-tile pair
  p.mma_n256 = no_n256 ? 0 : 1;
  if (tail && tail->coupling) {
    if (tail->pre != !next_pk.empty()) VB_THROW("flow tail: the next q|k|v projection needs the next pre-projection and vice versa");
    p.tail_coupling = 1; p.tail_flow = tail->flow ? 1 : 0; p.tail_pre = tail->pre ? 1 : 0;
    p.z = tail->z; p.zp_off = tail->zp_off; p.cp_backward = tail->backward ? 1 : 0;
    p.cp_bias = c.V(tail->out_pk + ".bias"); p.row_acc = tail->row_acc;
    p.fl_c = tail->flow_c; p.cond_off_next = tail->cond_off_next;
    if (tail->pre) {
      p.pre_bias = c.P(tail->pre_pn + ".pre_projection.bias"); p.pre_pw = c.P(tail->pre_pn + ".pos_weight"); p.pe = tail->pe;
    }
  }
  if (g_xrow_dbg) {
    p.dbg = g_xrow_dbg;
    g_xrow_dbg += static_cast<size_t>(cdiv(T, 128)) * B * 128;
  }
  const uint64_t T64 = static_cast<uint64_t>(T);
  const CUtensorMap tX = make_tmap(x.h, 3, d, T, B, d, T64 * d, 64, 128);
  const CUtensorMap tA1 = make_tmap(b.ctx, 3, d, T, B, d, T64 * d, 64, 128);
  const int kld = kv.nblk * kv.d;
  const int kpitch = kv.k_ld > 0 ? kv.k_ld : kld;
  const CUtensorMap tK = make_tmap(kv.k, 3, kld, Tt, B, kpitch, static_cast<uint64_t>(Tt) * kpitch, 64, 128);
  const CUtensorMap tXo = tape ? make_tmap(tape->o_h, 3, d, T, B, d, T64 * d, 64, 128) : tX;
  const CUtensorMap tV = make_tmap(kv.vt, 2, Tt, static_cast<uint64_t>(kv.nblk) * B * H * 64, 1, kv.tpad, 0, 64, 64);
  const CUtensorMap tW1 = make_tmap(c.W(pk + ".proj1"), 2, 2 * d, d, 1, 2 * d, 0, 64, 128);
  const CUtensorMap tWq = make_tmap(c.W(pk + ".cq"), 2, d, d, 1, d, 0, 64, 128);
  const CUtensorMap tW2 = make_tmap(c.W(pk + ".proj2"), 2, 2 * d, d, 1, 2 * d, 0, 64, 128);
  const CUtensorMap tF1 = make_tmap(c.W(pk + ".ffn1"), 2, d, XR_F, 1, d, 0, 64, 128);
  const CUtensorMap tF2 = make_tmap(c.W(pk + ".ffn2"), 2, XR_F, d, 1, XR_F, 0, 64, 128);
  const CUtensorMap tWn = p.has_next ? make_tmap(c.W(next_pk + ".qkv"), 2, d, 3 * d, 1, d, 0, 64, 128) : tWq;
  const CUtensorMap tWo = p.tail_coupling ? make_tmap(c.W(tail->out_pk), 2, d, 128, 1, d, 0, 64, 128) : tWq;
  const CUtensorMap tFl = p.tail_flow ? make_tmap(tail->flow_w, 2, 128, 256, 1, 128, 0, 64, 128) : tWq;
  const CUtensorMap tWp = p.tail_pre ? make_tmap(c.W(tail->pre_pk), 2, 64, d, 1, 64, 0, 64, 128) : tWq;
  const double rows = static_cast<double>(B) * T;
  const double macs_row = 2.0 * d * d + d * d + 2.0 * Tt * d + 2.0 * d * d + 2.0 * d * XR_F + (p.has_next ? 3.0 * d * d : 0.0) +
                          (p.tail_coupling ? 128.0 * d : 0.0) + (p.tail_flow ? 128.0 * 128 : 0.0) + (p.tail_pre ? 64.0 * d : 0.0);
  const double wbytes = 2.0 * (4.0 * d * d + d * d + 2.0 * d * XR_F + (p.has_next ? 3.0 * d * d : 0.0));
  ProfileScope prof(tape ? "xblk_row_tape" : (ali ? "xblk_row_ali" : "xblk_row"), 2.0 * rows * macs_row,
                    wbytes + rows * d * (4 + 2 + 2 + 4 + 2) + (p.has_next ? rows * 3 * d * 2 : 0) +
                        static_cast<double>(B) * Tt * 2 * d * 2 + (ali ? rows * H * Tt * 4 : 0) +
                        (tape ? rows * (d * (4 + 2 + 2 + 2 + 4 + 2) + XR_F * 2) : 0),
                    c.stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cdiv(T, 128), B);
  cfg.blockDim = dim3(XR_THREADS);
  cfg.dynamicSmemBytes = XR_SMEM;
  cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, xblk_row_kernel, tX, tA1, tK, tV, tW1, tWq, tW2, tF1, tF2, tWn, tWo, tFl, tWp, tXo, p);
  if (le != cudaSuccess) VB_THROW("cudaLaunchKernelEx(xblk_row_kernel) failed: %s", cudaGetErrorString(le));
  check_launch("xblk_row_kernel");
}

// CrossAttentionBLK.call (modules/attention.py:436-452); x updated in place.  `qkv_ready` (in/out): the self-attention
// q|k|v of THIS block already sit in b.qk / b.vt (written by the previous block's fused row kernel); on return it tells
// the same about the block named by next_pk.
static void xblk_fwd(Ctx& c, const std::string& pk, const std::string& pn, Stream2 x, const XblkBufs& b, int B, int T,
                     int d, int H, int ffn, const int* q_len, const MemKV& kv, int kv_blk, int Tt, const int* t_len,
                     float* ali, const std::string& next_pk, bool& qkv_ready, const XTail* tail = nullptr) {
  const int rows = B * T;
  if (!qkv_ready) {  // self-attention projections: Q | K row-major, V transposed
    GemmParams p = gp();
    p.mode = EPI_QKV; p.N = 3 * d; segs_plain(p, d);
    p.seq_T = T; p.seq_B = B; p.n_rowmajor = 2 * d; p.out_h = b.qk; p.ld_h = 2 * d;
    p.vt = b.vt; p.vt_ld = b.tpad; p.heads = H;
    run_gemm(c, 256, AOp{x.h, d, d}, AOp{}, 1, rows, c.W(pk + ".qkv"), d, 3 * d, p);
  }
  qkv_ready = false;
  run_attention(c, B, H, AttnCall{b.qk, 2 * d, 0, T, b.qk, 2 * d, d, T, b.vt, static_cast<long>(B) * H * 64, b.tpad, 0,
                                  q_len, q_len, 1, b.ctx, d, nullptr});
  if (xrow_supported(d, H, ffn, Tt, kv)) {
    run_xblk_row(c, pk, pn, next_pk, x, b, B, T, q_len, kv, kv_blk, Tt, t_len, ali, tail);
    qkv_ready = !next_pk.empty();
    return;
  }
  {  // LN1(att_proj1([x ; a1]) + x)
    GemmParams p = gp();
    p.mode = EPI_LN; p.N = d; segs_concat(p, d, d);
    p.bias = c.P(pn + ".att_proj1.bias"); p.residual = x.f; p.res_ld = d;
    p.ln_gamma = c.P(pn + ".layer_norm1.gamma"); p.ln_beta = c.P(pn + ".layer_norm1.beta");
    p.out_f32 = b.s.f; p.ld_f32 = d; p.out_h = b.s.h; p.ld_h = d;
    run_gemm(c, d / 2, AOp{x.h, d, d}, AOp{b.ctx, d, d}, 1, rows, c.W(pk + ".proj1"), 2 * d, d, p);
  }
  {  // cross-attention query projection
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = d; segs_plain(p, d);
    p.out_h = b.q; p.ld_h = d;
    run_gemm(c, 128, AOp{b.s.h, d, d}, AOp{}, 1, rows, c.W(pk + ".cq"), d, d, p);
  }
  run_attention(c, B, H, AttnCall{b.q, d, 0, T, kv.k, kv.nblk * kv.d, kv_blk * kv.d, Tt, kv.vt,
                                  static_cast<long>(kv.nblk) * B * H * 64, kv.tpad, static_cast<long>(kv_blk) * B * H * 64,
                                  q_len, t_len, 0, b.ctx, d, ali});
  {  // LN2(att_proj2([s ; a2]) + s)
    GemmParams p = gp();
    p.mode = EPI_LN; p.N = d; segs_concat(p, d, d);
    p.bias = c.P(pn + ".att_proj2.bias"); p.residual = b.s.f; p.res_ld = d;
    p.ln_gamma = c.P(pn + ".layer_norm2.gamma"); p.ln_beta = c.P(pn + ".layer_norm2.beta");
    p.out_f32 = b.cst.f; p.ld_f32 = d; p.out_h = b.cst.h; p.ld_h = d;
    run_gemm(c, d / 2, AOp{b.s.h, d, d}, AOp{b.ctx, d, d}, 1, rows, c.W(pk + ".proj2"), 2 * d, d, p);
  }
  {  // FFN (modules/utils.py:48-53)
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = ffn; segs_plain(p, d); p.act = 1;
    p.bias = c.P(pn + ".ffn.dense1.bias"); p.out_h = b.hid; p.ld_h = ffn;
    run_gemm(c, 128, AOp{b.cst.h, d, d}, AOp{}, 1, rows, c.W(pk + ".ffn1"), d, ffn, p);
  }
  {
    GemmParams p = gp();
    p.mode = EPI_LN; p.N = d; segs_plain(p, ffn);
    p.bias = c.P(pn + ".ffn.dense2.bias"); p.residual = b.cst.f; p.res_ld = d;
    p.ln_gamma = c.P(pn + ".ffn.layer_norm.gamma"); p.ln_beta = c.P(pn + ".ffn.layer_norm.beta");
    p.out_f32 = x.f; p.ld_f32 = d; p.out_h = x.h; p.ld_h = d;
    run_gemm(c, d / 2, AOp{b.hid, ffn, ffn}, AOp{}, 1, rows, c.W(pk + ".ffn2"), ffn, d, p);
  }
}
// all blocks of one module in sequence (the fused row kernel of block i also projects q|k|v of block i+1)
static void xblk_stack_fwd(Ctx& c, const std::string& pk_prefix, const std::string& pn_prefix, int nblk, Stream2 x,
                           const XblkBufs& b, int B, int T, int d, int H, int ffn, const int* q_len, const MemKV& kv,
                           int kv_blk0, int Tt, const int* t_len, float* ali, int64_t ali_blk_stride,
                           bool* qkv_ready_io = nullptr, const XTail* tail = nullptr, const std::string& after_pk = std::string()) {
  // qkv_ready_io: in = q|k|v of block 0 already projected (by the previous step's flow tail); out = the same about the block
  // named by after_pk (the first block of the NEXT coupling net), which the flow tail of the last block projects
  bool qkv_ready = qkv_ready_io ? *qkv_ready_io : false;
  for (int i = 0; i < nblk; ++i) {
    const bool last = i + 1 == nblk;
    xblk_fwd(c, pk_prefix + std::to_string(i), pn_prefix + std::to_string(i), x, b, B, T, d, H, ffn, q_len, kv, kv_blk0 + i,
             Tt, t_len, ali ? ali + static_cast<int64_t>(i) * ali_blk_stride : nullptr,
             last ? after_pk : pk_prefix + std::to_string(i + 1), qkv_ready, last ? tail : nullptr);
  }
  if (qkv_ready_io) *qkv_ready_io = qkv_ready;
}

// Conv1D wrapper of the reference: conv -> activation -> BatchNorm -> dropout (modules/utils.py:56-85).
// Inference: BN folded into the GEMM epilogue.  Training: GEMM -> fp32 activations, batch statistics over all
// (batch, time) rows incl. padding (deterministic two-stage reduction, fp64 finalise), moving-average update,
// normalise + dropout + fp16 (hi/lo) operand for the next layer.
static void conv_bn(Ctx& c, const std::string& pk, const std::string& pn, AOp a0, AOp a1, int B, int T, int cin, int C,
                    int taps, int act, bool split, int block_n, float drop_rate, __half* out_h, __half* out_lo) {
  GemmParams p = gp();
  p.mode = EPI_PLAIN; p.N = C; p.act = act; segs_conv(p, taps, cin, split);
  p.bias = c.P(pn + ".conv1d.bias");
  if (!c.training) {
    p.ch_scale = c.V(pk + ".bn_scale"); p.ch_shift = c.V(pk + ".bn_shift");
    p.out_h = out_h; p.out_lo = out_lo; p.ld_h = C;
    run_gemm(c, block_n, a0, a1, B, T, c.W(pk), c.WM(pk).K, C, p);
    return;
  }
  const int64_t mark = c.ws_off;
  const int64_t rows = static_cast<int64_t>(B) * T;
  float* y = c.alloc<float>(rows * C);
  p.out_f32 = y; p.ld_f32 = C;
  run_gemm(c, block_n, a0, a1, B, T, c.W(pk), c.WM(pk).K, C, p);
  const int rpt = 64, ntile = static_cast<int>(cdiv(static_cast<int>(rows), rpt));
  float* partial = c.alloc<float>(static_cast<int64_t>(ntile) * 2 * C);
  float* scale = c.alloc<float>(C);
  float* shift = c.alloc<float>(C);
  const float* mask = drop_rate > 0.f ? next_mask(c, rows * C, drop_rate) : nullptr;
  if (!c.dry) {
    colstats_partial_kernel<<<ntile, 256, 0, c.stream>>>(y, rows, C, rpt, partial);
    bn_train_finalize_kernel<<<cdiv(C, 128), 128, 0, c.stream>>>(partial, ntile, rows, C, c.P(pn + ".bn.gamma"),
                                                                c.P(pn + ".bn.beta"), c.PM(pn + ".bn.moving_mean"),
                                                                c.PM(pn + ".bn.moving_variance"), 0.99f, 1e-3f, scale,
                                                                shift, c.update_bn);
    bn_apply_kernel<<<static_cast<unsigned>((rows * C + 255) / 256), 256, 0, c.stream>>>(y, scale, shift, mask, out_h, out_lo,
                                                                                        rows * C, C);
    check_launch("bn_train");
  }
  c.ws_off = mark;
}

// ============================================================================ modules
// TransformerEncoder.call (modules/encoder.py:79-93), inference mode. Writes text_embd fp32 [B,Tt,E].
static void encoder_fwd(Ctx& c, const int* texts, const int* t_len, int B, int Tt, float pos_step, float* text_embd) {
  const vaenar_hparams_t& h = c.m->hp;
  const int E = h.enc_hidden, A = h.enc_att_dim, H = h.enc_heads, F = h.enc_ffn;
  const int64_t rows = static_cast<int64_t>(B) * Tt;
  const int64_t mark = c.ws_off;
  __half* xa = c.alloc<__half>(rows * E);
  __half* xb = c.alloc<__half>(rows * E);
  float* hf = c.alloc<float>(rows * E);
  __half* hh = c.alloc<__half>(rows * E);
  __half* qk = c.alloc<__half>(rows * 2 * A);
  const int tpad = vt_pad(Tt);
  __half* vt = c.alloc<__half>(static_cast<int64_t>(B) * H * 64 * tpad);
  __half* ctx = c.alloc<__half>(rows * A);
  __half* hid = c.alloc<__half>(rows * F);
  float* pe = c.alloc<float>(static_cast<int64_t>(Tt) * E);
  if (!c.dry) {
    embed_kernel<<<static_cast<unsigned>(rows), 128, 0, c.stream>>>(texts, c.P("text_encoder.emb_layer.embeddings"), xa,
                                                                   static_cast<int>(rows), E, h.vocab_size);
    check_launch("embed");
  }
  run_pe(c, pe, Tt, E, pos_step);
  int cin = h.embd_dim;
  for (int i = 0; i < h.enc_n_conv; ++i) {   // ConvPreNet (modules/utils.py:21-38): conv -> relu -> BN -> dropout
    const std::string pk = "enc.conv" + std::to_string(i), pn = "text_encoder.prenet.conv_stack." + std::to_string(i);
    conv_bn(c, pk, pn, AOp{xa, cin, cin}, AOp{}, B, Tt, cin, E, h.enc_conv_kernel, 1, false, 128, h.enc_pre_drop_rate, xb,
            nullptr);
    std::swap(xa, xb);
    cin = E;
  }
  {  // prenet projection + pos_weight * PE (encoder.py:84-86)
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = E; segs_plain(p, E);
    p.bias = c.P("text_encoder.prenet.projection.bias");
    p.seq_T = Tt; p.seq_B = B;
    p.add_table = pe; p.add_ld = E; p.add_scale = c.P("text_encoder.pos_weight");
    p.out_f32 = text_embd; p.ld_f32 = E; p.out_h = xb; p.ld_h = E;
    run_gemm(c, 128, AOp{xa, E, E}, AOp{}, 1, static_cast<int>(rows), c.W("enc.proj"), E, E, p);
    std::swap(xa, xb);
    if (c.training && h.enc_pos_drop_rate > 0.f) {   // pe_dropout (encoder.py:87)
      const float* mk = next_mask(c, rows * E, h.enc_pos_drop_rate);
      run_dropout(c, text_embd, mk, text_embd, xa, rows * E);
    }
  }
  for (int i = 0; i < h.enc_n_blk; ++i) {   // SelfAttentionBLK (modules/attention.py:405-415)
    const std::string pk = "enc.blk" + std::to_string(i), pn = "text_encoder.self_attentions." + std::to_string(i);
    {
      GemmParams p = gp();
      p.mode = EPI_QKV; p.N = 3 * A; segs_plain(p, E);
      p.seq_T = Tt; p.seq_B = B; p.n_rowmajor = 2 * A; p.out_h = qk; p.ld_h = 2 * A;
      p.vt = vt; p.vt_ld = tpad; p.heads = H;
      run_gemm(c, 256, AOp{xa, E, E}, AOp{}, 1, static_cast<int>(rows), c.W(pk + ".qkv"), E, 3 * A, p);
    }
    run_attention(c, B, H, AttnCall{qk, 2 * A, 0, Tt, qk, 2 * A, A, Tt, vt, static_cast<long>(B) * H * 64, tpad, 0, t_len,
                                    t_len, 0, ctx, A, nullptr});
    {
      GemmParams p = gp();
      p.mode = EPI_LN; p.N = E; segs_concat(p, E, A);
      p.bias = c.P(pn + ".att_proj.bias"); p.residual = text_embd; p.res_ld = E;
      p.ln_gamma = c.P(pn + ".layer_norm.gamma"); p.ln_beta = c.P(pn + ".layer_norm.beta");
      p.out_f32 = hf; p.ld_f32 = E; p.out_h = hh; p.ld_h = E;
      run_gemm(c, E / 2, AOp{xa, E, E}, AOp{ctx, A, A}, 1, static_cast<int>(rows), c.W(pk + ".proj"), E + A, E, p);
    }
    {
      GemmParams p = gp();
      p.mode = EPI_PLAIN; p.N = F; p.act = 1; segs_plain(p, E);
      p.bias = c.P(pn + ".ffn.dense1.bias"); p.out_h = hid; p.ld_h = F;
      run_gemm(c, 128, AOp{hh, E, E}, AOp{}, 1, static_cast<int>(rows), c.W(pk + ".ffn1"), E, F, p);
    }
    {
      GemmParams p = gp();
      p.mode = EPI_LN; p.N = E; segs_plain(p, F);
      p.bias = c.P(pn + ".ffn.dense2.bias"); p.residual = hf; p.res_ld = E;
      p.ln_gamma = c.P(pn + ".ffn.layer_norm.gamma"); p.ln_beta = c.P(pn + ".ffn.layer_norm.beta");
      p.out_f32 = text_embd; p.ld_f32 = E; p.out_h = xa; p.ld_h = E;
      run_gemm(c, E / 2, AOp{hid, F, F}, AOp{}, 1, static_cast<int>(rows), c.W(pk + ".ffn2"), F, E, p);
    }
  }
  c.ws_off = mark;
}

static void length_predictor_fwd(Ctx& c, const float* text_embd, const int* t_len, int B, int Tt, float* out) {
  if (c.dry) return;
  length_predictor_kernel<<<B, 256, 0, c.stream>>>(text_embd, c.P("length_predictor.projection.kernel"),
                                                  c.P("length_predictor.projection.bias"), t_len, out, Tt,
                                                  c.m->hp.enc_hidden);
  check_launch("length_predictor");
}

// One flow step's conditioner + coupling (modules/flow.py:223-257, modules/transform.py:45-59).
// `fuse`: what follows the coupling inside the fused row kernel of the step's last block (csrc/xblk_fused.cuh flow tail):
//   flow_M / flow_c  the folded ActNorm (+) InvertibleLinear map applied to z right after the coupling (the NEXT flow
//                    operation of the chain: step s+1 when sampling, step s itself when evaluating log p), or null;
//   next_step        >= 0: also the pre-projection + positional encoding + first q|k|v projection of that step's net.
// `pre_done`: x / q|k|v of this step were already produced by the previous step's tail.
struct FlowFuse {
  bool enabled = false;
  const __half* flow_w = nullptr;
  const float* flow_c = nullptr;
  int next_step = -1;
};
static bool flow_tail_supported(Ctx& c, int Tt, const MemKV& kv) {
  const vaenar_hparams_t& h = c.m->hp;
  static const bool off = getenv("VAENAR_NO_FLOW_TAIL") != nullptr;
  return !off && h.latent_dim == 128 && xrow_supported(h.prior_att_dim, h.prior_heads, h.prior_ffn, Tt, kv);
}
static void coupling_step(Ctx& c, int s, bool backward, float* z, __half* zh, float* row_acc, Stream2 x, const XblkBufs& xb,
                          const float* pe, const MemKV& kv, int B, int Tz, int Tt, const int* z_len, const int* t_len,
                          bool pre_done = false, const FlowFuse& fuse = FlowFuse()) {
  const vaenar_hparams_t& h = c.m->hp;
  const int d = h.prior_att_dim, H = h.prior_heads, F = h.prior_ffn, L = h.latent_dim;
  const int rows = B * Tz;
  const bool upper = (s % 2 == 0);           // modules/prior.py:85-87
  const int cond_off = upper ? 0 : L / 2;    // conditioning half
  const int zp_off = upper ? L / 2 : 0;      // transformed half
  const std::string pk = "prior." + std::to_string(s), pn = "prior.glow." + std::to_string(s) + ".affine_coupling.net";
  if (!pre_done) {
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = d; segs_plain(p, L / 2);
    p.bias = c.P(pn + ".pre_projection.bias");
    p.seq_T = Tz; p.seq_B = B; p.add_table = pe; p.add_ld = d; p.add_scale = c.P(pn + ".pos_weight");
    p.out_f32 = x.f; p.ld_f32 = d; p.out_h = x.h; p.ld_h = d;
    run_gemm(c, 128, AOp{c.dry ? nullptr : zh + cond_off, L, L / 2}, AOp{}, 1, rows, c.W(pk + ".pre"), kpad64(L / 2), d, p);
  }
  bool qkv_ready = pre_done;
  if (fuse.enabled) {
    XTail tail;
    tail.coupling = true;
    tail.out_pk = pk + ".out";
    tail.z = z; tail.zp_off = zp_off; tail.backward = backward; tail.row_acc = row_acc;
    tail.flow = fuse.flow_w != nullptr;
    tail.flow_w = fuse.flow_w; tail.flow_c = fuse.flow_c;
    std::string after;
    if (fuse.next_step >= 0) {
      const int ns = fuse.next_step;
      tail.pre = true;
      tail.pre_pk = "prior." + std::to_string(ns) + ".pre";
      tail.pre_pn = "prior.glow." + std::to_string(ns) + ".affine_coupling.net";
      tail.cond_off_next = (ns % 2 == 0) ? 0 : L / 2;
      tail.pe = pe;
      after = "prior." + std::to_string(ns) + ".blk0";
    }
    xblk_stack_fwd(c, pk + ".blk", pn + ".attentions.", h.prior_n_tblk, x, xb, B, Tz, d, H, F, z_len, kv, s * h.prior_n_tblk, Tt,
                   t_len, nullptr, 0, &qkv_ready, &tail, after);
    return;
  }
  xblk_stack_fwd(c, pk + ".blk", pn + ".attentions.", h.prior_n_tblk, x, xb, B, Tz, d, H, F, z_len, kv, s * h.prior_n_tblk, Tt,
                 t_len, nullptr, 0, &qkv_ready);
  {
    GemmParams p = gp();
    p.mode = EPI_COUPLING; p.N = L; segs_plain(p, d);
    p.seq_T = Tz; p.seq_B = B;
    p.bias = c.V(pk + ".out.bias");
    p.z = z; p.z_h = zh; p.z_ld = L; p.zp_off = zp_off; p.backward = backward ? 1 : 0;
    p.row_acc = row_acc; p.lengths = z_len;
    run_gemm(c, 128, AOp{x.h, d, d}, AOp{}, 1, rows, c.W(pk + ".out"), d, L, p);
  }
}

static void flow_linear(Ctx& c, float* z, __half* zh, const float* M, const float* cvec, int64_t rows) {
  if (c.dry) return;
  flow_linear_kernel<<<static_cast<unsigned>((rows + FLOW_ROWS - 1) / FLOW_ROWS), 256, FLOW_LIN_SMEM, c.stream>>>(
      z, z, zh, M, cvec, rows, 0);
  check_launch("flow_linear");
}

struct PriorBufs {
  __half* emb_h;
  __half* zh;
  float* row_acc;
  float* base;
  float* pe;
  Stream2 x;
  XblkBufs xb;
  MemKV kv;
};
static PriorBufs prior_setup(Ctx& c, const float* text_embd, int B, int Tt, int Tz) {
  const vaenar_hparams_t& h = c.m->hp;
  const int d = h.prior_att_dim, E = h.enc_hidden, L = h.latent_dim;
  const int64_t rows = static_cast<int64_t>(B) * Tz;
  PriorBufs b;
  b.emb_h = c.alloc<__half>(static_cast<int64_t>(B) * Tt * E);
  run_cast(c, text_embd, b.emb_h, static_cast<int64_t>(B) * Tt * E);
  b.zh = c.alloc<__half>(rows * L);
  b.row_acc = c.alloc<float>(rows);
  b.base = c.alloc<float>(B);
  b.pe = c.alloc<float>(static_cast<int64_t>(Tz) * d);
  b.x.f = c.alloc<float>(rows * d);
  b.x.h = c.alloc<__half>(rows * d);
  b.xb = xblk_bufs(c, B, Tz, d, h.prior_heads, h.prior_ffn);
  b.kv = memory_kv(c, "prior.kv", b.emb_h, B, Tt, E, h.prior_n_blk * h.prior_n_tblk, d, h.prior_heads);
  run_pe(c, b.pe, Tz, d, 1.0f);
  if (!c.dry) VB_CUDA(cudaMemsetAsync(b.row_acc, 0, rows * sizeof(float), c.stream));
  return b;
}

// TransformerPrior.sample (modules/prior.py:154-169); z holds epsilon on entry.
static void prior_sample(Ctx& c, const float* text_embd, const int* t_len, const int* z_len, int B, int Tt, int Tz,
                         float* z, float* logp) {
  const vaenar_hparams_t& h = c.m->hp;
  const int L = h.latent_dim;
  const int64_t mark = c.ws_off;
  PriorBufs b = prior_setup(c, text_embd, B, Tt, Tz);
  if (!c.dry) {
    VB_CUDA(cudaMemsetAsync(b.base, 0, B * sizeof(float), c.stream));
    base_logprob_kernel<<<dim3(B, 16), 256, 0, c.stream>>>(z, z_len, b.base, Tz, L);
    check_launch("base_logprob");
  }
  const float* Mf = c.dry ? nullptr : reinterpret_cast<const float*>(c.packed + c.m->off_Mf);
  const float* cf = c.dry ? nullptr : reinterpret_cast<const float*>(c.packed + c.m->off_cf);
  const bool fused = flow_tail_supported(c, Tt, b.kv);
  const __half* Wf = c.dry ? nullptr : c.W("flow.f");
  for (int s = 0; s < h.prior_n_blk; ++s) {
    const bool pre_done = fused && s > 0;    // the previous step's flow tail applied this step's map and pre-projection
    if (!pre_done)
      flow_linear(c, z, b.zh, c.dry ? nullptr : Mf + static_cast<int64_t>(s) * L * L, c.dry ? nullptr : cf + s * L,
                  static_cast<int64_t>(B) * Tz);
    FlowFuse fuse;
    fuse.enabled = fused;
    if (fused && s + 1 < h.prior_n_blk) {
      fuse.flow_w = c.dry ? nullptr : Wf + static_cast<int64_t>(s + 1) * 2 * L * L;
      fuse.flow_c = c.dry ? nullptr : cf + (s + 1) * L;
      fuse.next_step = s + 1;
    }
    coupling_step(c, s, false, z, b.zh, b.row_acc, b.x, b.xb, b.pe, b.kv, B, Tz, Tt, z_len, t_len, pre_done, fuse);
  }
  if (!c.dry) {
    prior_logp_finalize_kernel<<<B, 256, 0, c.stream>>>(b.base, b.row_acc,
                                                       reinterpret_cast<const float*>(c.packed + c.m->off_consts),
                                                       h.prior_n_blk, z_len, logp, Tz, -1.0f);
    check_launch("prior_logp_finalize");
  }
  c.ws_off = mark;
}

// TransformerPrior.init (modules/prior.py:171-186): like sample, but every ActNorm is first initialised from the
// statistics of its own input over ALL B*T positions (modules/flow.py:189-196); writes the parameters.
static void prior_init(Ctx& c, const float* text_embd, const int* t_len, const int* z_len, int B, int Tt, int Tz, float* z) {
  const vaenar_hparams_t& h = c.m->hp;
  const int L = h.latent_dim;
  const int64_t rows = static_cast<int64_t>(B) * Tz;
  const int64_t mark = c.ws_off;
  PriorBufs b = prior_setup(c, text_embd, B, Tt, Tz);
  const int rpt = 64, ntile = cdiv(static_cast<int>(rows), rpt);
  float* partial = c.alloc<float>(static_cast<int64_t>(ntile) * 2 * L);
  float* Mf = c.dry ? nullptr : reinterpret_cast<float*>(c.packed + c.m->off_Mf);
  float* cf = c.dry ? nullptr : reinterpret_cast<float*>(c.packed + c.m->off_cf);
  for (int s = 0; s < h.prior_n_blk; ++s) {
    const std::string g = "prior.glow." + std::to_string(s);
    if (!c.dry) {
      colstats_partial_kernel<<<ntile, 256, 0, c.stream>>>(z, rows, L, rpt, partial);
      actnorm_init_kernel<<<1, FLOW_DIM, 0, c.stream>>>(partial, ntile, rows, c.PM(g + ".actnorm.log_scale"),
                                                       c.PM(g + ".actnorm.bias"), c.P(g + ".linear.weight"),
                                                       Mf + static_cast<int64_t>(s) * L * L, cf + s * L);
      check_launch("actnorm_init");
    }
    flow_linear(c, z, b.zh, c.dry ? nullptr : Mf + static_cast<int64_t>(s) * L * L, c.dry ? nullptr : cf + s * L, rows);
    coupling_step(c, s, false, z, b.zh, b.row_acc, b.x, b.xb, b.pe, b.kv, B, Tz, Tt, z_len, t_len);
  }
  c.ws_off = mark;
}

// TransformerPrior.log_probability (modules/prior.py:119-152)
static void prior_logprob(Ctx& c, const float* z_in, const float* text_embd, const int* t_len, const int* z_len, int B,
                          int Tt, int Tz, float* logp) {
  const vaenar_hparams_t& h = c.m->hp;
  const int L = h.latent_dim;
  const int64_t rows = static_cast<int64_t>(B) * Tz;
  const int64_t mark = c.ws_off;
  float* z = c.alloc<float>(rows * L);
  PriorBufs b = prior_setup(c, text_embd, B, Tt, Tz);
  if (!c.dry) VB_CUDA(cudaMemcpyAsync(z, z_in, rows * L * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  run_cast(c, z, b.zh, rows * L);
  const float* Mb = c.dry ? nullptr : reinterpret_cast<const float*>(c.packed + c.m->off_Mb);
  const float* cb = c.dry ? nullptr : reinterpret_cast<const float*>(c.packed + c.m->off_cb);
  const bool fused = flow_tail_supported(c, Tt, b.kv);
  const __half* Wb = c.dry ? nullptr : c.W("flow.b");
  for (int s = h.prior_n_blk - 1; s >= 0; --s) {
    FlowFuse fuse;
    fuse.enabled = fused;
    if (fused) {   // the inverse map of THIS step follows its coupling; then the pre-projection of step s - 1
      fuse.flow_w = c.dry ? nullptr : Wb + static_cast<int64_t>(s) * 2 * L * L;
      fuse.flow_c = c.dry ? nullptr : cb + s * L;
      fuse.next_step = s - 1;
    }
    coupling_step(c, s, true, z, b.zh, b.row_acc, b.x, b.xb, b.pe, b.kv, B, Tz, Tt, z_len, t_len,
                  fused && s + 1 < h.prior_n_blk, fuse);
    if (!fused) flow_linear(c, z, b.zh, c.dry ? nullptr : Mb + static_cast<int64_t>(s) * L * L, c.dry ? nullptr : cb + s * L, rows);
  }
  if (!c.dry) {
    VB_CUDA(cudaMemsetAsync(b.base, 0, B * sizeof(float), c.stream));
    base_logprob_kernel<<<dim3(B, 16), 256, 0, c.stream>>>(z, z_len, b.base, Tz, L);
    check_launch("base_logprob");
    // backward log-dets: coupling -sum log scale (already signed in row_acc), linear -len*log|det W|, actnorm -len*sum s
    prior_logp_finalize_kernel<<<B, 256, 0, c.stream>>>(b.base, b.row_acc,
                                                       reinterpret_cast<const float*>(c.packed + c.m->off_consts),
                                                       h.prior_n_blk, z_len, logp, Tz, 1.0f);
    check_launch("prior_logp_finalize");
  }
  c.ws_off = mark;
}

// TransformerPosterior.call + reparameterize + log_probability (modules/posterior.py:20-72,115-130)
static void posterior_fwd(Ctx& c, const float* mels, const float* text_embd, const int* t_len, const int* z_len,
                          const float* eps, int B, int Tt, int Tm, int Tz, int rf, float* z, float* logq,
                          float* save_lv = nullptr) {
  const vaenar_hparams_t& h = c.m->hp;
  const int d = h.posterior_att_dim, H = h.posterior_heads, F = h.posterior_ffn, E = h.enc_hidden, L = h.latent_dim,
            O = h.out_dim;
  const int64_t rows = static_cast<int64_t>(B) * Tz;
  const int64_t mark = c.ws_off;
  __half* emb_h = c.alloc<__half>(static_cast<int64_t>(B) * Tt * E);
  run_cast(c, text_embd, emb_h, static_cast<int64_t>(B) * Tt * E);
  __half* rm = c.alloc<__half>(rows * O);
  __half* a1 = c.alloc<__half>(rows * d);
  Stream2 x{c.alloc<float>(rows * d), c.alloc<__half>(rows * d)};
  __half* zh = c.alloc<__half>(rows * L);
  float* row_acc = c.alloc<float>(rows);
  float* pe = c.alloc<float>(static_cast<int64_t>(Tz) * d);
  XblkBufs xb = xblk_bufs(c, B, Tz, d, H, F);
  MemKV kv = memory_kv(c, "post.kv", emb_h, B, Tt, E, h.posterior_nblk, d, H);
  run_pe(c, pe, Tz, d, 1.0f);
  if (!c.dry) {
    reduce_mels_kernel<<<static_cast<unsigned>(rows), 96, 0, c.stream>>>(mels, rm, B, Tm, Tz, rf, O);
    check_launch("reduce_mels");
    VB_CUDA(cudaMemsetAsync(row_acc, 0, rows * sizeof(float), c.stream));
  }
  if (!c.training) {
    {  // PreNet dense1 + relu (modules/utils.py:13-18); dropout inactive (training=False)
      GemmParams p = gp();
      p.mode = EPI_PLAIN; p.N = d; p.act = 1; segs_plain(p, O);
      p.bias = c.P("posterior.prenet.dense1.bias"); p.out_h = a1; p.ld_h = d;
      run_gemm(c, 128, AOp{rm, O, O}, AOp{}, 1, static_cast<int>(rows), c.W("post.pre1"), kpad64(O), d, p);
    }
    {  // dense2 + relu, then + pos_weight * PE (posterior.py:118-121)
      GemmParams p = gp();
      p.mode = EPI_PLAIN; p.N = d; p.act = 1; segs_plain(p, d);
      p.bias = c.P("posterior.prenet.dense2.bias");
      p.seq_T = Tz; p.seq_B = B; p.add_table = pe; p.add_ld = d; p.add_scale = c.P("posterior.pos_weight");
      p.out_f32 = x.f; p.ld_f32 = d; p.out_h = x.h; p.ld_h = d;
      run_gemm(c, 128, AOp{a1, d, d}, AOp{}, 1, static_cast<int>(rows), c.W("post.pre2"), d, d, p);
    }
  } else {
    // training: relu -> dropout after each dense (the PreNet inherits the Keras call-context training flag,
    // posterior.py:117), then + pos_weight * PE, then pe_dropout (posterior.py:118-121)
    const float* mk1 = next_mask(c, rows * d, h.posterior_pre_drop_rate);
    const float* mk2 = next_mask(c, rows * d, h.posterior_pre_drop_rate);
    const float* mk3 = next_mask(c, rows * d, h.posterior_pos_drop_rate);
    {
      GemmParams p = gp();
      p.mode = EPI_PLAIN; p.N = d; p.act = 1; segs_plain(p, O);
      p.bias = c.P("posterior.prenet.dense1.bias"); p.out_f32 = x.f; p.ld_f32 = d;
      run_gemm(c, 128, AOp{rm, O, O}, AOp{}, 1, static_cast<int>(rows), c.W("post.pre1"), kpad64(O), d, p);
      run_dropout(c, x.f, mk1, nullptr, a1, rows * d);
    }
    {
      GemmParams p = gp();
      p.mode = EPI_PLAIN; p.N = d; p.act = 1; segs_plain(p, d);
      p.bias = c.P("posterior.prenet.dense2.bias"); p.out_f32 = x.f; p.ld_f32 = d;
      run_gemm(c, 128, AOp{a1, d, d}, AOp{}, 1, static_cast<int>(rows), c.W("post.pre2"), d, d, p);
      if (!c.dry) {
        prenet_tail_kernel<<<static_cast<unsigned>((rows * d + 255) / 256), 256, 0, c.stream>>>(
            x.f, mk2, pe, c.P("posterior.pos_weight"), mk3, Tz, d, x.h, rows * d);
        check_launch("prenet_tail");
      }
    }
  }
  xblk_stack_fwd(c, "post.blk", "posterior.attentions.", h.posterior_nblk, x, xb, B, Tz, d, H, F, z_len, kv, 0, Tt, t_len,
                 nullptr, 0);
  {
    GemmParams p = gp();
    p.mode = EPI_POSTERIOR; p.N = 2 * L; segs_plain(p, d);
    p.seq_T = Tz; p.seq_B = B;
    p.bias = c.V("post.out.bias"); p.eps_in = eps; p.z = z; p.z_h = zh; p.z_ld = L; p.row_acc = row_acc; p.lengths = z_len;
    p.save_lv = save_lv;
    run_gemm(c, 2 * L, AOp{x.h, d, d}, AOp{}, 1, static_cast<int>(rows), c.W("post.out"), d, 2 * L, p);
  }
  if (!c.dry) {
    row_sum_kernel<<<B, 256, 0, c.stream>>>(row_acc, logq, Tz);
    check_launch("row_sum");
  }
  c.ws_off = mark;
}

// TransformerDecoder.call (modules/decoder.py:181-199), inference mode
static void decoder_fwd(Ctx& c, const float* z, const float* text_embd, const int* z_len, const int* t_len, int B, int Tt,
                        int Tz, int rf, float* initial, float* mel, float* ali, int64_t ali_blk_stride = 0) {
  const vaenar_hparams_t& h = c.m->hp;
  const int d = h.dec_att_dim, H = h.dec_heads, F = h.dec_ffn, E = h.enc_hidden, L = h.latent_dim, O = h.out_dim,
            C = h.post_filters;
  if (rf < 1 || rf > h.max_reduction_factor) VB_THROW("reduction factor %d outside [1, %d]", rf, h.max_reduction_factor);
  const int64_t rows = static_cast<int64_t>(B) * Tz;
  const int Tm = Tz * rf;
  const int64_t mrows = static_cast<int64_t>(B) * Tm;
  const int64_t mark = c.ws_off;
  __half* emb_h = c.alloc<__half>(static_cast<int64_t>(B) * Tt * E);
  run_cast(c, text_embd, emb_h, static_cast<int64_t>(B) * Tt * E);
  __half* zh = c.alloc<__half>(rows * L);
  run_cast(c, z, zh, rows * L);
  Stream2 x{c.alloc<float>(rows * d), c.alloc<__half>(rows * d)};
  XblkBufs xb = xblk_bufs(c, B, Tz, d, H, F);
  MemKV kv = memory_kv(c, "dec.kv", emb_h, B, Tt, E, h.dec_nblk, d, H);
  __half* ini_h = c.alloc<__half>(mrows * O);
  __half* ini_l = c.alloc<__half>(mrows * O);
  __half* pa_h = c.alloc<__half>(mrows * C);
  __half* pa_l = c.alloc<__half>(mrows * C);
  __half* pb_h = c.alloc<__half>(mrows * C);
  __half* pb_l = c.alloc<__half>(mrows * C);
  {  // pre_projection (no positional encoding in the decoder, decoder.py:186-188)
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = d; segs_plain(p, L);
    p.bias = c.P("decoder.pre_projection.bias");
    p.out_f32 = x.f; p.ld_f32 = d; p.out_h = x.h; p.ld_h = d;
    run_gemm(c, 128, AOp{zh, L, L}, AOp{}, 1, static_cast<int>(rows), c.W("dec.pre"), L, d, p);
  }
  xblk_stack_fwd(c, "dec.blk", "decoder.attentions.", h.dec_nblk, x, xb, B, Tz, d, H, F, z_len, kv, 0, Tt, t_len, ali,
                 ali_blk_stride ? ali_blk_stride : static_cast<int64_t>(B) * H * Tz * Tt);
  {  // out_projection, first rf*80 columns only (decoder.py:193); [B*Tz, rf*80] == [B*Tz*rf, 80]
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = rf * O; segs_plain(p, d);
    p.bias = c.P("decoder.out_projection.bias");
    p.out_f32 = initial; p.ld_f32 = rf * O; p.out_h = ini_h; p.out_lo = ini_l; p.ld_h = rf * O;
    run_gemm(c, 128, AOp{x.h, d, d}, AOp{}, 1, static_cast<int>(rows), c.W("dec.out"), d, rf * O, p);
  }
  // PostNet (modules/utils.py:98-115): conv -> tanh (last: identity) -> BN; split-fp16 operands
  __half *in_h = ini_h, *in_l = ini_l, *out_h = pa_h, *out_l = pa_l;
  int cin = O;
  static const int post_bn_env = getenv("VAENAR_POST_BN") ? atoi(getenv("VAENAR_POST_BN")) : 0;
  const int post_bn = post_bn_env ? post_bn_env : 256;
  for (int i = 0; i < h.post_n_conv; ++i) {
    const std::string pk = "dec.post" + std::to_string(i), pn = "decoder.postnet.conv_stack." + std::to_string(i);
    const bool split = i >= post_plain();
    conv_bn(c, pk, pn, AOp{in_h, cin, cin}, split ? AOp{in_l, cin, cin} : AOp{}, B, Tm, cin, C, h.post_kernel,
            (i < h.post_n_conv - 1) ? 2 : 0, split, post_bn, h.post_drop_rate, out_h, out_l);
    in_h = out_h; in_l = out_l;
    out_h = (in_h == pa_h) ? pb_h : pa_h;
    out_l = (in_l == pa_l) ? pb_l : pa_l;
    cin = C;
  }
  {  // residual_projection + initial (decoder.py:197-198)
    GemmParams p = gp();
    p.mode = EPI_PLAIN; p.N = O;
    p.nseg = 3;
    for (int q = 0; q < 3; ++q) { p.seg_map[q] = (q == 1) ? 1 : 0; p.seg_shift[q] = 0; p.seg_kblocks[q] = cdiv(C, 64); }
    p.alg_k = C;
    p.bias = c.P("decoder.residual_projection.bias"); p.residual = initial; p.res_ld = O;
    p.out_f32 = mel; p.ld_f32 = O;
    run_gemm(c, 128, AOp{in_h, C, C}, AOp{in_l, C, C}, 1, static_cast<int>(mrows), c.W("dec.res"), 3 * C, O, p);
  }
  c.ws_off = mark;
}

// ---- small loss kernels (models/models.py:88-103)
__global__ void length_loss_kernel(const float* pred, const int* lens, float* out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const float d = logf(pred[b]) - logf(static_cast<float>(lens[b]));
    out[b] = d * d;
  }
}
__global__ void kl_kernel(const float* logq, const float* logp, float* out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = logq[b] - logp[b];
}

// VAENAR.call forward (models/models.py:105-197), training=False, n_sample = 1
static void elbo_fwd(Ctx& c, const int* texts, const float* mels, const int* m_len, const int* t_len, const int* z_len,
                     const float* eps, int B, int Tt, int Tm, int Tz, int rf, float* mel_out, float* l2, float* kl,
                     float* len_loss, float* ali) {
  const vaenar_hparams_t& h = c.m->hp;
  const int E = h.enc_hidden, L = h.latent_dim, O = h.out_dim;
  const int64_t mark = c.ws_off;
  float* emb = c.alloc<float>(static_cast<int64_t>(B) * Tt * E);
  float* z = c.alloc<float>(static_cast<int64_t>(B) * Tz * L);
  float* ini = c.alloc<float>(static_cast<int64_t>(B) * Tz * rf * O);
  float* fin = c.alloc<float>(static_cast<int64_t>(B) * Tz * rf * O);
  float* pred = c.alloc<float>(B);
  float* logq = c.alloc<float>(B);
  float* logp = c.alloc<float>(B);
  encoder_fwd(c, texts, t_len, B, Tt, h.mel_text_len_ratio / static_cast<float>(rf), emb);
  length_predictor_fwd(c, emb, t_len, B, Tt, pred);
  posterior_fwd(c, mels, emb, t_len, z_len, eps, B, Tt, Tm, Tz, rf, z, logq);
  decoder_fwd(c, z, emb, z_len, t_len, B, Tt, Tz, rf, ini, fin, ali);
  prior_logprob(c, z, emb, t_len, z_len, B, Tt, Tz, logp);
  if (!c.dry) {
    length_loss_kernel<<<cdiv(B, 128), 128, 0, c.stream>>>(pred, m_len, len_loss, B);
    VB_CUDA(cudaMemsetAsync(l2, 0, B * sizeof(float), c.stream));
    l2_loss_kernel<<<dim3(B, 16), 256, 0, c.stream>>>(fin, Tz * rf, mels, Tm, m_len, l2, O);
    l2_loss_kernel<<<dim3(B, 16), 256, 0, c.stream>>>(ini, Tz * rf, mels, Tm, m_len, l2, O);
    kl_kernel<<<cdiv(B, 128), 128, 0, c.stream>>>(logq, logp, kl, B);
    check_launch("loss kernels");
    VB_CUDA(cudaMemcpy2DAsync(mel_out, static_cast<size_t>(Tm) * O * 4, fin, static_cast<size_t>(Tz) * rf * O * 4,
                              static_cast<size_t>(Tm) * O * 4, B, cudaMemcpyDeviceToDevice, c.stream));
  }
  c.ws_off = mark;
}

static void inference_one(Ctx& c, const int* texts, const int* t_len, const int* z_len, int B, int Tt, int Tz, int rf,
                          float* z_io, float* text_embd, float* mel, float* ali, int64_t ali_blk_stride, float* logp) {
  const vaenar_hparams_t& h = c.m->hp;
  const int64_t mark = c.ws_off;
  float* ini = c.alloc<float>(static_cast<int64_t>(B) * Tz * rf * h.out_dim);
  encoder_fwd(c, texts, t_len, B, Tt, h.mel_text_len_ratio / static_cast<float>(rf), text_embd);
  prior_sample(c, text_embd, t_len, z_len, B, Tt, Tz, z_io, logp);
  decoder_fwd(c, z_io, text_embd, z_len, t_len, B, Tt, Tz, rf, ini, mel, ali, ali_blk_stride);
  c.ws_off = mark;
}

// VAENAR.inference.  One launch chain by default (96 kernels at the LJSpeech shape).  VAENAR_DUAL_CHAIN=1: utterances are
// independent in inference (BatchNorm uses moving statistics), so the batch can also run as two halves on two concurrent
// launch chains (caller stream + an internal side stream, fork/join with events; capturable into one CUDA graph) -- that
// was the faster schedule for the per-op launch chain of round 1; with the fused row kernels both schedules take the same time.
static void inference_fwd(Ctx& c, const int* texts, const int* t_len, const int* z_len, int B, int Tt, int Tz, int rf,
                          float* z_io, float* text_embd, float* mel, float* ali, float* logp) {
  const vaenar_hparams_t& h = c.m->hp;
  static const bool dual_chain = getenv("VAENAR_DUAL_CHAIN") != nullptr;
  const int B0 = (B >= 4 && dual_chain) ? B / 2 : B, B1 = B - B0;
  const int64_t ali_stride = static_cast<int64_t>(B) * h.dec_heads * Tz * Tt;
  const int64_t mark = c.ws_off;
  if (B1 == 0) {
    inference_one(c, texts, t_len, z_len, B, Tt, Tz, rf, z_io, text_embd, mel, ali, ali_stride, logp);
    return;
  }
  Ctx c1 = c;                      // second chain: its own workspace region and stream
  if (!c.dry) {
    vaenar_model* m = c.m;
    if (!m->side_stream) {
      VB_CUDA(cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking));
      VB_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
      VB_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
    c1.stream = m->side_stream;
    VB_CUDA(cudaEventRecord(m->ev_fork, c.stream));
    VB_CUDA(cudaStreamWaitEvent(c1.stream, m->ev_fork, 0));
  }
  inference_one(c, texts, t_len, z_len, B0, Tt, Tz, rf, z_io, text_embd, mel, ali, ali_stride, logp);
  // region of the second chain starts after the first chain's peak
  c1.ws_off = align_up(c.ws_peak, 1024);
  c1.ws_peak = c1.ws_off;
  const int L = h.latent_dim, E = h.enc_hidden, O = h.out_dim;
  inference_one(c1, c.dry ? nullptr : texts + static_cast<int64_t>(B0) * Tt, c.dry ? nullptr : t_len + B0,
                c.dry ? nullptr : z_len + B0, B1, Tt, Tz, rf, c.dry ? nullptr : z_io + static_cast<int64_t>(B0) * Tz * L,
                c.dry ? nullptr : text_embd + static_cast<int64_t>(B0) * Tt * E,
                c.dry ? nullptr : mel + static_cast<int64_t>(B0) * Tz * rf * O,
                (c.dry || !ali) ? nullptr : ali + static_cast<int64_t>(B0) * h.dec_heads * Tz * Tt, ali_stride,
                c.dry ? nullptr : logp + B0);
  c.ws_peak = std::max(c.ws_peak, c1.ws_peak);
  if (!c.dry) {
    VB_CUDA(cudaEventRecord(c.m->ev_join, c1.stream));
    VB_CUDA(cudaStreamWaitEvent(c.stream, c.m->ev_join, 0));
  }
  c.ws_off = mark;
}

// VAENAR.init (models/models.py:212-226): training=True, rf = max_reduction_factor
static void init_fwd(Ctx& c, const int* texts, const int* t_len, const int* z_len, int B, int Tt, int Tz, float* z_io,
                     float* mel) {
  const vaenar_hparams_t& h = c.m->hp;
  const int rf = h.max_reduction_factor;
  const int64_t mark = c.ws_off;
  float* emb = c.alloc<float>(static_cast<int64_t>(B) * Tt * h.enc_hidden);
  float* ini = c.alloc<float>(static_cast<int64_t>(B) * Tz * rf * h.out_dim);
  encoder_fwd(c, texts, t_len, B, Tt, h.mel_text_len_ratio / static_cast<float>(rf), emb);
  prior_init(c, emb, t_len, z_len, B, Tt, Tz, z_io);
  decoder_fwd(c, z_io, emb, z_len, t_len, B, Tt, Tz, rf, ini, mel, nullptr);
  c.ws_off = mark;
}

static void join_flow(Ctx& c);
#include "train.inc"

// ============================================================================ weight packing
static void pack_weights(vaenar_model* m, const float* params, uint8_t* packed, cudaStream_t stream) {
  const vaenar_hparams_t& h = m->hp;
  VB_CUDA(cudaMemsetAsync(packed, 0, m->packed_bytes, stream));
  // 1. transposed fp16 operand matrices
  m->host_ops.clear();
  int maxK = 0, maxN = 0;
  for (const PackPlanOp& o : m->plan) {
    const PackedMat& pm = m->pmats.at(o.dst);
    PackOp op;
    op.src = params + m->params[o.param].offset + o.src_off;
    op.dst = reinterpret_cast<__half*>(packed + pm.off) + static_cast<int64_t>(o.n_off) * pm.K + o.k_off;
    op.K = o.K; op.N = o.N; op.lds = o.lds; op.ldd = pm.K; op.mode = o.mode;
    if (o.mode == 2 ? (o.n_off + o.K > pm.N || o.k_off + o.N > pm.K) : (o.n_off + o.N > pm.N || o.k_off + o.K > pm.K))
      VB_THROW("pack op outside %s", o.dst.c_str());
    m->host_ops.push_back(op);
    maxK = std::max(maxK, o.K);
    maxN = std::max(maxN, o.N);
  }
  PackOp* dops = reinterpret_cast<PackOp*>(packed + m->off_packops);
  VB_CUDA(cudaMemcpyAsync(dops, m->host_ops.data(), m->host_ops.size() * sizeof(PackOp), cudaMemcpyHostToDevice, stream));
  const int nops = static_cast<int>(m->host_ops.size());
  m->host_tiles.assign(nops + 1, 0);
  for (int i = 0; i < nops; ++i)
    m->host_tiles[i + 1] = m->host_tiles[i] + cdiv(m->host_ops[i].N, PACK_TILE) * cdiv(m->host_ops[i].K, PACK_TILE);
  int* dtiles = reinterpret_cast<int*>(packed + m->off_packtiles);
  VB_CUDA(cudaMemcpyAsync(dtiles, m->host_tiles.data(), (nops + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
  pack_weights_flat_kernel<<<m->host_tiles[nops], 256, 0, stream>>>(dops, dtiles, nops);
  check_launch("pack_weights");
  (void)maxK; (void)maxN;
  auto PP = [&](const std::string& n) { return params + m->params[m->P(n)].offset; };
  auto VV = [&](const std::string& n) { return reinterpret_cast<float*>(packed + m->pvecs.at(n)); };
  // 2. inference BatchNorm -> affine
  auto fold = [&](const std::string& pk, const std::string& pn, int C) {
    bn_fold_kernel<<<cdiv(C, 128), 128, 0, stream>>>(PP(pn + ".bn.gamma"), PP(pn + ".bn.beta"), PP(pn + ".bn.moving_mean"),
                                                    PP(pn + ".bn.moving_variance"), 1e-3f, VV(pk + ".bn_scale"),
                                                    VV(pk + ".bn_shift"), C);
  };
  for (int i = 0; i < h.enc_n_conv; ++i)
    fold("enc.conv" + std::to_string(i), "text_encoder.prenet.conv_stack." + std::to_string(i), h.enc_hidden);
  for (int i = 0; i < h.post_n_conv; ++i)
    fold("dec.post" + std::to_string(i), "decoder.postnet.conv_stack." + std::to_string(i), h.post_filters);
  // 3. concatenated biases
  const int L = h.latent_dim;
  concat2_kernel<<<cdiv(2 * L, 128), 128, 0, stream>>>(PP("posterior.mu_projection.bias"), L,
                                                      PP("posterior.logvar_projection.bias"), L, VV("post.out.bias"));
  for (int s = 0; s < h.prior_n_blk; ++s) {
    const std::string n = "prior.glow." + std::to_string(s) + ".affine_coupling.net";
    concat2_kernel<<<1, 128, 0, stream>>>(PP(n + ".log_scale_proj.bias"), L / 2, PP(n + ".shift_proj.bias"), L / 2,
                                          VV("prior." + std::to_string(s) + ".out.bias"));
  }
  check_launch("bn_fold/concat");
  // 4. flow: float64 log|det W|, fp32 inverse, ActNorm folded maps
  const int S = h.prior_n_blk;
  m->host_ptrs.assign(3 * S, nullptr);
  for (int s = 0; s < S; ++s) {
    const std::string g = "prior.glow." + std::to_string(s);
    m->host_ptrs[s] = PP(g + ".linear.weight");
    m->host_ptrs[S + s] = PP(g + ".actnorm.log_scale");
    m->host_ptrs[2 * S + s] = PP(g + ".actnorm.bias");
  }
  const float** dptrs = reinterpret_cast<const float**>(packed + m->off_ptrs);
  VB_CUDA(cudaMemcpyAsync(dptrs, m->host_ptrs.data(), 3 * S * sizeof(float*), cudaMemcpyHostToDevice, stream));
  // The 128x128 LU / inverse kernels are latency-bound single-CTA work: run them on the side stream so that they
  // overlap the operand packing and whatever the caller launches next; consumers join through ev_flow_ready.
  cudaStream_t main_stream = stream;
  if (!m->wgrad_stream) VB_CUDA(cudaStreamCreateWithFlags(&m->wgrad_stream, cudaStreamNonBlocking));
  if (!m->ev_flow_ready) VB_CUDA(cudaEventCreateWithFlags(&m->ev_flow_ready, cudaEventDisableTiming));
  {
    if (!m->ev_fork) {
      VB_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
      VB_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    }
    VB_CUDA(cudaEventRecord(m->ev_fork, main_stream));
    VB_CUDA(cudaStreamWaitEvent(m->wgrad_stream, m->ev_fork, 0));
    stream = m->wgrad_stream;
  }
  double* ld64 = reinterpret_cast<double*>(packed + m->off_logdet64);
  float* winv = reinterpret_cast<float*>(packed + m->off_winv);
  slogdet128_kernel<<<S, FLOW_DIM, FLOW_DIM * (FLOW_DIM + 1) * 8, stream>>>(dptrs, ld64);
  inverse128_kernel<<<S, 2 * FLOW_DIM, (FLOW_DIM * (2 * FLOW_DIM + 1) + FLOW_DIM) * 4, stream>>>(dptrs, winv);
  flow_fold_kernel<<<S, FLOW_DIM, 0, stream>>>(dptrs, dptrs + S, dptrs + 2 * S, winv, ld64,
                                              reinterpret_cast<float*>(packed + m->off_Mf),
                                              reinterpret_cast<float*>(packed + m->off_cf),
                                              reinterpret_cast<float*>(packed + m->off_Mb),
                                              reinterpret_cast<float*>(packed + m->off_cb),
                                              reinterpret_cast<float*>(packed + m->off_consts));
  {
    const long n = static_cast<long>(S) * FLOW_DIM * FLOW_DIM;
    flow_pack_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float*>(packed + m->off_Mf), reinterpret_cast<__half*>(packed + m->pmats.at("flow.f").off), S);
    flow_pack_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float*>(packed + m->off_Mb), reinterpret_cast<__half*>(packed + m->pmats.at("flow.b").off), S);
  }
  check_launch("flow prepare");
  VB_CUDA(cudaEventRecord(m->ev_flow_ready, stream));
  m->flow_pending = true;
}
// order the flow constants produced by the last pack_weights before what follows on c.stream
static void join_flow(Ctx& c) {
  if (c.dry || !c.m || !c.m->flow_pending) return;
  VB_CUDA(cudaStreamWaitEvent(c.stream, c.m->ev_flow_ready, 0));
  c.m->flow_pending = false;
}

// ============================================================================ C ABI
#define API_BEGIN try {
#define API_END                          \
  }                                      \
  catch (const EngineError& e) {         \
    g_err = e.msg;                       \
    return -1;                           \
  }                                      \
  catch (const std::exception& e) {      \
    g_err = e.what();                    \
    return -2;                           \
  }                                      \
  return 0;

static Ctx make_ctx(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes, void* stream,
                    bool defer_flow_join = false) {
  if (!h) VB_THROW("null handle");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) VB_THROW("no CUDA device: the B200 path has no CPU fallback");
  set_attrs(h);
  Ctx c;
  c.m = h; c.params = params; c.packed = const_cast<uint8_t*>(static_cast<const uint8_t*>(packed));
  c.ws = static_cast<uint8_t*>(ws); c.ws_bytes = ws_bytes; c.stream = static_cast<cudaStream_t>(stream);
  if (!defer_flow_join) join_flow(c);
  return c;
}

extern "C" {

const char* vaenar_last_error(void) { return g_err.c_str(); }
int vaenar_abi_version(void) { return 1; }

int vaenar_create(const vaenar_hparams_t* hps, vaenar_handle_t* out) {
  API_BEGIN
  if (!hps || !out) VB_THROW("null argument");
  vaenar_model* m = new vaenar_model();
  m->hp = *hps;
  try {
    build_model(*m);
  } catch (...) {
    delete m;
    throw;
  }
  *out = m;
  API_END
}
int vaenar_destroy(vaenar_handle_t h) {
  if (h) {
    if (h->ev_flow_ready) cudaEventDestroy(h->ev_flow_ready);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->wgrad_stream) cudaStreamDestroy(h->wgrad_stream);
    if (h->lane_stream) cudaStreamDestroy(h->lane_stream);
    if (h->lane_wgrad_stream) cudaStreamDestroy(h->lane_wgrad_stream);
    if (h->main_stream) cudaStreamDestroy(h->main_stream);
    if (h->attn_stream) cudaStreamDestroy(h->attn_stream);
    if (h->lane_attn_stream) cudaStreamDestroy(h->lane_attn_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
  }
  delete h;
  return 0;
}
int vaenar_num_params(vaenar_handle_t h) { return h ? static_cast<int>(h->params.size()) : -1; }
const char* vaenar_param_name(vaenar_handle_t h, int i) { return h->params[i].name.c_str(); }
int vaenar_param_ndim(vaenar_handle_t h, int i) { return h->params[i].ndim; }
int64_t vaenar_param_dim(vaenar_handle_t h, int i, int d) { return h->params[i].dims[d]; }
int64_t vaenar_param_offset(vaenar_handle_t h, int i) { return h->params[i].offset; }
int vaenar_param_trainable(vaenar_handle_t h, int i) { return h->params[i].trainable ? 1 : 0; }
int64_t vaenar_param_floats(vaenar_handle_t h) { return h->param_floats; }
int64_t vaenar_packed_bytes(vaenar_handle_t h) { return h->packed_bytes; }

int64_t vaenar_workspace_bytes(vaenar_handle_t h, int B, int T_text, int T_z, int rf) {
  try {
    Ctx c;
    c.m = h; c.params = nullptr; c.packed = nullptr; c.ws = nullptr; c.ws_bytes = 0; c.dry = true; c.stream = nullptr;
    inference_fwd(c, nullptr, nullptr, nullptr, B, T_text, T_z, rf, nullptr, nullptr, nullptr, nullptr, nullptr);
    elbo_fwd(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, T_text, T_z * rf, T_z, rf, nullptr, nullptr, nullptr,
             nullptr, nullptr);
    c.training = true;   // training-mode forward needs fp32 conv activations, statistics and masks
    c.mask_cursor = 0;
    elbo_fwd(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, T_text, T_z * rf, T_z, rf, nullptr, nullptr, nullptr,
             nullptr, nullptr);
    return align_up(c.ws_peak, 1024) + 4096;
  } catch (const EngineError& e) {
    g_err = e.msg;
    return -1;
  }
}

int vaenar_pack_weights(vaenar_handle_t h, const float* params, void* packed, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, nullptr, 0, stream);
  pack_weights(h, params, static_cast<uint8_t*>(packed), c.stream);
  join_flow(c);   // everything, incl. the flow constants prepared on the internal stream, is ordered on `stream`
  API_END
}
int vaenar_pack_weights_async(vaenar_handle_t h, const float* params, void* packed, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, nullptr, 0, stream);
  pack_weights(h, params, static_cast<uint8_t*>(packed), c.stream);   // flow constants joined by the next library call
  API_END
}

int vaenar_text_encoder_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const int32_t* texts, const int32_t* text_lengths, int B, int T_text, float pos_step,
                            float* text_embd, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  encoder_fwd(c, texts, text_lengths, B, T_text, pos_step, text_embd);
  API_END
}

int vaenar_length_predictor_fwd(vaenar_handle_t h, const float* params, const float* text_embd,
                                const int32_t* text_lengths, int B, int T_text, float* pred_lengths, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, nullptr, nullptr, 0, stream);
  length_predictor_fwd(c, text_embd, text_lengths, B, T_text, pred_lengths);
  API_END
}

int vaenar_prior_sample(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                        const float* text_embd, const int32_t* text_lengths, const int32_t* z_lengths, int B,
                        int T_text, int T_z, float* z_io, float* logp, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  prior_sample(c, text_embd, text_lengths, z_lengths, B, T_text, T_z, z_io, logp);
  API_END
}

int vaenar_prior_log_probability(vaenar_handle_t h, const float* params, const void* packed, void* ws,
                                 int64_t ws_bytes, const float* z, const float* text_embd,
                                 const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z,
                                 float* logp, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  prior_logprob(c, z, text_embd, text_lengths, z_lengths, B, T_text, T_z, logp);
  API_END
}

int vaenar_posterior_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                         const float* mels, const float* text_embd, const int32_t* text_lengths,
                         const int32_t* z_lengths, const float* eps, int B, int T_text, int T_mel, int T_z, int rf,
                         float* z, float* logq, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  posterior_fwd(c, mels, text_embd, text_lengths, z_lengths, eps, B, T_text, T_mel, T_z, rf, z, logq);
  API_END
}

int vaenar_posterior_params(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const float* reduced_mels, const float* text_embd, const int32_t* text_lengths,
                            const int32_t* z_lengths, int B, int T_text, int T_z, float* mu_projection_out,
                            float* logvar_projection_out, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  // the fused epilogue computes z = eps * exp(lv / 2) + mean with lv = mu_projection output, mean = logvar_projection
  // output (models/models.py:136): with eps = 0, z is the logvar_projection output and the saved lv the mu_projection's
  const int64_t n = static_cast<int64_t>(B) * T_z * h->hp.latent_dim;
  float* zero = c.alloc<float>(n);
  float* logq = c.alloc<float>(B);
  VB_CUDA(cudaMemsetAsync(zero, 0, n * sizeof(float), c.stream));
  posterior_fwd(c, reduced_mels, text_embd, text_lengths, z_lengths, zero, B, T_text, T_z, T_z, 1, logvar_projection_out, logq,
                mu_projection_out);
  API_END
}

int vaenar_decoder_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                       const float* z, const float* text_embd, const int32_t* z_lengths,
                       const int32_t* text_lengths, int B, int T_text, int T_z, int rf, float* initial_mel,
                       float* mel, float* alignments, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  decoder_fwd(c, z, text_embd, z_lengths, text_lengths, B, T_text, T_z, rf, initial_mel, mel, alignments);
  API_END
}

int vaenar_inference(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                     const int32_t* texts, const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text,
                     int T_z, int rf, float* z_io, float* text_embd, float* mel, float* alignments, float* logp,
                     void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  inference_fwd(c, texts, text_lengths, z_lengths, B, T_text, T_z, rf, z_io, text_embd, mel, alignments, logp);
  API_END
}

int vaenar_elbo_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                    const int32_t* texts, const float* mels, const int32_t* mel_lengths,
                    const int32_t* text_lengths, const int32_t* z_lengths, const float* eps, int B, int T_text,
                    int T_mel, int T_z, int rf, float* mel_out, float* l2, float* kl, float* length_loss,
                    float* alignments, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  elbo_fwd(c, texts, mels, mel_lengths, text_lengths, z_lengths, eps, B, T_text, T_mel, T_z, rf, mel_out, l2, kl,
           length_loss, alignments);
  API_END
}

static void apply_train_opts(Ctx& c, float* params, const vaenar_train_opts_t* o) {
  c.training = true;
  c.params_mut = params;
  c.mask_cursor = 0;
  if (o) {
    c.masks = o->masks; c.n_masks = o->n_masks; c.seed = o->seed; c.update_bn = o->update_bn_stats;
  }
}

int vaenar_elbo_fwd_train(vaenar_handle_t h, float* params, const void* packed, void* ws, int64_t ws_bytes,
                          const int32_t* texts, const float* mels, const int32_t* mel_lengths,
                          const int32_t* text_lengths, const int32_t* z_lengths, const float* eps, int B, int T_text,
                          int T_mel, int T_z, int rf, const vaenar_train_opts_t* opts, float* mel_out, float* l2,
                          float* kl, float* length_loss, float* alignments, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  apply_train_opts(c, params, opts);
  elbo_fwd(c, texts, mels, mel_lengths, text_lengths, z_lengths, eps, B, T_text, T_mel, T_z, rf, mel_out, l2, kl,
           length_loss, alignments);
  API_END
}

int vaenar_init(vaenar_handle_t h, float* params, void* packed, void* ws, int64_t ws_bytes, const int32_t* texts,
                const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z,
                const vaenar_train_opts_t* opts, float* z_io, float* mel, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  apply_train_opts(c, params, opts);
  init_fwd(c, texts, text_lengths, z_lengths, B, T_text, T_z, z_io, mel);
  API_END
}

int vaenar_prior_init(vaenar_handle_t h, float* params, void* packed, void* ws, int64_t ws_bytes, const float* text_embd,
                      const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z, float* z_io,
                      void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  c.params_mut = params;
  prior_init(c, text_embd, text_lengths, z_lengths, B, T_text, T_z, z_io);
  API_END
}

int64_t vaenar_train_workspace_bytes(vaenar_handle_t h, int B, int T_text, int T_z, int rf) {
  try {
    Ctx c;
    c.m = h; c.params = nullptr; c.packed = nullptr; c.ws = nullptr; c.ws_bytes = 0; c.dry = true; c.stream = nullptr;
    c.training = true;
    train_grads(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, T_text, T_z * rf, T_z, rf, 0.f, 0.f, 1.f, nullptr,
                nullptr);
    return align_up(c.ws_peak, 1024) + 4096;
  } catch (const EngineError& e) {
    g_err = e.msg;
    return -1;
  }
}

int vaenar_train_step_grads(vaenar_handle_t h, float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const int32_t* texts, const float* mels, const int32_t* mel_lengths, const int32_t* text_lengths,
                            const int32_t* z_lengths, const float* eps, int B, int T_text, int T_mel, int T_z, int rf,
                            const vaenar_train_opts_t* opts, float kl_weight, float length_weight, float loss_scale, float* grads,
                            float* losses, float* mel_out, void* stream) {
  API_BEGIN
  if (!grads || !losses) VB_THROW("null gradient / loss buffer");
  if (!(loss_scale > 0.f)) VB_THROW("loss_scale must be positive");
  // The activation chain (forward, dgrad chain; the decoder branch) runs on HIGH-priority streams forked off the caller's
  // stream, the weight-gradient and dK / dV streams keep the default (lowest) priority: the block scheduler then places the
  // CTAs of the critical chain first and the one-wave weight-gradient grids fill what is left.
  h->ev_cursor = 0;
  cudaStream_t user_stream = static_cast<cudaStream_t>(stream);
  static const bool no_prio = getenv("VAENAR_NO_PRIO") != nullptr;
  const bool prio = !no_prio && !g_profile;
  int prio_hi = 0;
  if (prio) {
    int prio_lo = 0;
    VB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (!h->main_stream) VB_CUDA(cudaStreamCreateWithPriority(&h->main_stream, cudaStreamNonBlocking, prio_hi));
    cudaEvent_t e = next_event(h);
    VB_CUDA(cudaEventRecord(e, user_stream));
    VB_CUDA(cudaStreamWaitEvent(h->main_stream, e, 0));
    stream = h->main_stream;
  }
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream, /*defer_flow_join=*/true);   // joined right before the prior
  apply_train_opts(c, params, opts);
  if (!getenv("VAENAR_NO_WGRAD_STREAM")) {
    if (!h->wgrad_stream) VB_CUDA(cudaStreamCreateWithFlags(&h->wgrad_stream, cudaStreamNonBlocking));
    c.wstream = h->wgrad_stream;
  }
  // decoder forward/backward beside the prior's (not while per-class timing is on: class times must stay exclusive)
  static const bool no_lane = getenv("VAENAR_NO_DEC_LANE") != nullptr;
  if (!no_lane && !g_profile) {
    if (!h->lane_stream) VB_CUDA(cudaStreamCreateWithPriority(&h->lane_stream, cudaStreamNonBlocking, prio_hi));
    c.lane = h->lane_stream;
    if (c.wstream) {
      if (!h->lane_wgrad_stream) VB_CUDA(cudaStreamCreateWithFlags(&h->lane_wgrad_stream, cudaStreamNonBlocking));
      c.lane_w = h->lane_wgrad_stream;
    }
  }
  static const bool no_attn_stream = getenv("VAENAR_NO_ATTN_STREAM") != nullptr;
  if (!no_attn_stream && !g_profile) {
    if (!h->attn_stream) VB_CUDA(cudaStreamCreateWithFlags(&h->attn_stream, cudaStreamNonBlocking));
    c.astream = h->attn_stream;
    if (c.lane) {
      if (!h->lane_attn_stream) VB_CUDA(cudaStreamCreateWithFlags(&h->lane_attn_stream, cudaStreamNonBlocking));
      c.lane_a = h->lane_attn_stream;
    }
  }
  train_grads(c, grads, texts, mels, mel_lengths, text_lengths, z_lengths, eps, B, T_text, T_mel, T_z, rf, kl_weight, length_weight,
              loss_scale, losses, mel_out);
  if (prio) {   // every side stream has been joined into c.stream by train_grads
    cudaEvent_t e = next_event(h);
    VB_CUDA(cudaEventRecord(e, c.stream));
    VB_CUDA(cudaStreamWaitEvent(user_stream, e, 0));
  }
  API_END
}

int vaenar_trainable_mask(vaenar_handle_t h, uint8_t* host_mask) {
  API_BEGIN
  if (!h || !host_mask) VB_THROW("null argument");
  memset(host_mask, 0, static_cast<size_t>(h->param_floats));
  for (const ParamInfo& pi : h->params)
    if (pi.trainable) memset(host_mask + pi.offset, 1, static_cast<size_t>(pi.numel));
  API_END
}

int vaenar_grad_nonfinite(const float* grads, int64_t n, float* count, void* stream) {
  API_BEGIN
  if (!grads || !count) VB_THROW("null argument");
  nonfinite_count_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(grads, n, count);
  check_launch("nonfinite_count");
  API_END
}

int vaenar_adam_step(float* params, const float* grads, float* m, float* v, const uint8_t* trainable_mask, int64_t n,
                     int64_t step, float lr, float beta1, float beta2, float eps, float grad_scale, const float* skip_flag,
                     void* stream) {
  API_BEGIN
  if (step < 1) VB_THROW("Adam step counts from 1");
  for (const void* q : {static_cast<const void*>(params), static_cast<const void*>(grads), static_cast<const void*>(m),
                        static_cast<const void*>(v)})
    if (reinterpret_cast<uintptr_t>(q) & 15) VB_THROW("adam_step: buffers must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(trainable_mask) & 3) VB_THROW("adam_step: trainable mask must be 4-byte aligned");
  const double lr_t = static_cast<double>(lr) * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(step))) /
                      (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(step)));
  adam_step_kernel<<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      params, grads, m, v, trainable_mask, n, static_cast<float>(lr_t), beta1, beta2, eps, grad_scale, skip_flag);
  check_launch("adam_step");
  API_END
}

int vaenar_adam_step_sharded(float* const* peer_params, const float* const* peer_grads, float* m_shard, float* v_shard,
                             const uint8_t* trainable_mask, int64_t n, int rank, int world, int64_t step, float lr, float beta1,
                             float beta2, float eps, float grad_scale, const float* skip_flag, void* stream) {
  API_BEGIN
  if (step < 1) VB_THROW("Adam step counts from 1");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) VB_THROW("sharded Adam: rank %d / world %d (max 8)", rank, world);
  if (n % 4) VB_THROW("sharded Adam: the flat buffer length must be a multiple of 4");
  const int64_t chunk = (n / 4 + world - 1) / world * 4;
  const int64_t lo = std::min<int64_t>(n, rank * chunk), hi = std::min<int64_t>(n, lo + chunk);
  PeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int r = 0; r < world; ++r) { pp.params[r] = peer_params[r]; pp.grads[r] = peer_grads[r]; }
  const double lr_t = static_cast<double>(lr) * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(step))) /
                      (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(step)));
  if (hi > lo) {
    adam_sharded_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pp, m_shard, v_shard, trainable_mask, lo, hi, rank, world,
                                                                               static_cast<float>(lr_t), beta1, beta2, eps, grad_scale, skip_flag);
    check_launch("adam_sharded");
  }
  API_END
}
int vaenar_enable_peer_access(int peer_device) {
  API_BEGIN
  int cur = -1;
  VB_CUDA(cudaGetDevice(&cur));
  if (peer_device == cur) return 0;
  int can = 0;
  VB_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) VB_THROW("device %d cannot access device %d over NVLink / PCIe peer-to-peer", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) VB_THROW("cudaDeviceEnablePeerAccess(%d) failed: %s", peer_device, cudaGetErrorString(e));
  API_END
}
// Export the cudaMalloc allocation that contains `ptr` for other processes: its cudaIpcMemHandle_t (64 bytes) and the
// byte offset of `ptr` inside it (PyTorch's caching allocator sub-allocates from large cudaMalloc segments).
int vaenar_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out) {
  API_BEGIN
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      VB_THROW("cuMemGetAddressRange unavailable");
    fn = reinterpret_cast<PFN_range>(p);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) VB_THROW("cuMemGetAddressRange failed for %p", ptr);
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess)
    VB_THROW("cudaIpcGetMemHandle failed (%s): the buffer must come from cudaMalloc (PyTorch: do not use "
             "PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True)", cudaGetErrorString(e));
  memcpy(handle64_out, &h, sizeof(h));
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  API_END
}
// Map a peer process's allocation (cudaIpcMemHandle_t, 64 bytes) into THIS device's address space with peer access.
int vaenar_ipc_open(const void* handle64, void** out_ptr) {
  API_BEGIN
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  VB_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  API_END
}
int vaenar_ipc_close(void* ptr) {
  API_BEGIN
  VB_CUDA(cudaIpcCloseMemHandle(ptr));
  API_END
}
int64_t vaenar_adam_shard_floats(int64_t n, int world) { return (n / 4 + world - 1) / world * 4; }

// CRC32C (Castagnoli) of a host buffer -- checksums of the TF tensor-bundle / TFRecord formats (host code only).
uint32_t vaenar_crc32c(const void* data, int64_t n, uint32_t crc) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  for (int64_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

int vaenar_randn(float* out, int64_t n, uint64_t seed, uint64_t stream_id, float stddev, void* stream) {
  API_BEGIN
  const int64_t thr = (n + 3) / 4;
  randn_kernel<<<static_cast<unsigned>((thr + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed,
                                                                                                    stream_id, stddev);
  check_launch("randn");
  API_END
}

// ---------------------------------------------------------------------------- mel inversion (csrc/griffin_lim.cuh)
struct GLWorkspace {
  double2* tw;
  double* window;
  double* carry;
  double* peak;
  double* frames[2];
  int max_chunks;
};
static int64_t gl_workspace_layout(int B, int T, int win, int hop, uint8_t* base, GLWorkspace* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const int64_t L = static_cast<int64_t>(hop) * std::max(T - 1, 0);
  const int max_chunks = static_cast<int>((L + PRE_CHUNK - 1) / PRE_CHUNK) + 1;
  uint8_t* tw = take(sizeof(double2) * (GL_N / 2));
  uint8_t* window = take(sizeof(double) * GL_MAX_WIN);
  uint8_t* carry = take(sizeof(double) * static_cast<int64_t>(B) * max_chunks);
  uint8_t* peak = take(sizeof(double) * B);
  uint8_t* f0 = take(sizeof(double) * static_cast<int64_t>(B) * T * win);
  uint8_t* f1 = take(sizeof(double) * static_cast<int64_t>(B) * T * win);
  if (w) {
    w->tw = reinterpret_cast<double2*>(tw); w->window = reinterpret_cast<double*>(window);
    w->carry = reinterpret_cast<double*>(carry); w->peak = reinterpret_cast<double*>(peak);
    w->frames[0] = reinterpret_cast<double*>(f0); w->frames[1] = reinterpret_cast<double*>(f1);
    w->max_chunks = max_chunks;
  }
  return off;
}
static void gl_check_shape(int B, int T, int num_freq, int win, int hop) {
  if (num_freq != GL_BINS) VB_THROW("griffin_lim: num_freq %d unsupported (n_fft is fixed at %d, num_freq %d)", num_freq, GL_N, GL_BINS);
  if (B < 1 || T < 2) VB_THROW("griffin_lim: need B >= 1 and at least 2 frames (B %d, T %d)", B, T);
  if (win < 1 || win > GL_MAX_WIN || hop < 1 || hop > win) VB_THROW("griffin_lim: win_length %d / hop_length %d unsupported", win, hop);
}
static void gl_need_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) VB_THROW("no CUDA device: the B200 path has no CPU fallback");
}

int64_t vaenar_griffin_lim_workspace_bytes(int B, int T, int win_length, int hop_length) {
  if (B < 1 || T < 1 || win_length < 1 || hop_length < 1) return -1;
  return gl_workspace_layout(B, T, win_length, hop_length, nullptr, nullptr);
}

int vaenar_mel_to_linear(const float* mel, const int32_t* n_frames, const float* inv_basis_t, int B, int T, int n_mels,
                         int num_freq, float min_level_db, float ref_level_db, float max_abs_value, int symmetric,
                         float power, double* S, void* stream) {
  API_BEGIN
  gl_need_device();
  if (num_freq != GL_BINS) VB_THROW("mel_to_linear: num_freq %d unsupported (%d)", num_freq, GL_BINS);
  if (n_mels < 1 || n_mels > 128) VB_THROW("mel_to_linear: num_mels %d unsupported (1..128)", n_mels);
  if (B < 1 || T < 1) VB_THROW("mel_to_linear: empty batch");
  MelToLinearParams p;
  p.mel = mel; p.inv_t = inv_basis_t; p.n_frames = n_frames; p.S = S;
  p.T = T; p.n_mels = n_mels; p.s_ld = GL_BINS; p.symmetric = symmetric;
  p.min_level_db = min_level_db; p.ref_level_db = ref_level_db; p.max_abs = max_abs_value; p.power = power;
  mel_to_linear_kernel<<<dim3(T, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  check_launch("mel_to_linear_kernel");
  API_END
}

int vaenar_griffin_lim(const double* S, const int32_t* n_frames, const double* rand, uint64_t seed, int B, int T,
                       int num_freq, int win_length, int hop_length, int iters, void* ws, int64_t ws_bytes, double* wav,
                       int64_t wav_ld, void* stream) {
  API_BEGIN
  gl_need_device();
  gl_check_shape(B, T, num_freq, win_length, hop_length);
  if (iters < 0) VB_THROW("griffin_lim: iters %d", iters);
  const int64_t L = static_cast<int64_t>(hop_length) * (T - 1);
  if (wav_ld < L) VB_THROW("griffin_lim: wav row pitch %lld < %lld samples", (long long)wav_ld, (long long)L);
  GLWorkspace w;
  const int64_t need = gl_workspace_layout(B, T, win_length, hop_length, static_cast<uint8_t*>(ws), &w);
  if (!ws || ws_bytes < need) VB_THROW("griffin_lim: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  const int smem = GL_SMEM;
  if (!attr_set) {
    VB_CUDA(cudaFuncSetAttribute(gl_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  gl_tables_kernel<<<cdiv(GL_N / 2, 256), 256, 0, st>>>(w.tw, w.window, win_length);
  check_launch("gl_tables_kernel");
  GLParams p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.rand = rand; p.n_frames = n_frames; p.tw = w.tw; p.window = w.window; p.seed = seed;
  p.T = T; p.s_ld = GL_BINS; p.win = win_length; p.hop = hop_length; p.lpad = (GL_N - win_length) / 2;
  const dim3 grid(cdiv(T, 2), B);
  for (int it = 0; it <= iters; ++it) {   // pass 0: y = istft(S * random phases); passes 1..iters: audio.py:98-100
    p.first = it == 0;
    p.prev = w.frames[(it + 1) & 1];
    p.next = w.frames[it & 1];
    gl_iter_kernel<<<grid, GL_THREADS, smem, st>>>(p);
    check_launch("gl_iter_kernel");
  }
  gl_finalize_kernel<<<dim3(static_cast<unsigned>((wav_ld + 255) / 256), B), 256, 0, st>>>(
      w.frames[iters & 1], w.window, n_frames, T, win_length, hop_length, p.lpad, wav, wav_ld);
  check_launch("gl_finalize_kernel");
  API_END
}

int vaenar_inv_preemphasis(double* wav, int64_t wav_ld, const int32_t* n_frames, int B, int T, int win_length,
                           int hop_length, double k, void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  gl_need_device();
  GLWorkspace w;
  const int64_t need = gl_workspace_layout(B, T, win_length, hop_length, static_cast<uint8_t*>(ws), &w);
  if (!ws || ws_bytes < need) VB_THROW("inv_preemphasis: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)need);
  inv_preemphasis_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, wav_ld, n_frames, hop_length, k, w.carry,
                                                                            w.max_chunks);
  check_launch("inv_preemphasis_kernel");
  API_END
}

int vaenar_wav_to_int16(const double* wav, int64_t wav_ld, const int32_t* n_frames, int B, int T, int win_length,
                        int hop_length, void* ws, int64_t ws_bytes, int16_t* out, void* stream) {
  API_BEGIN
  gl_need_device();
  GLWorkspace w;
  const int64_t need = gl_workspace_layout(B, T, win_length, hop_length, static_cast<uint8_t*>(ws), &w);
  if (!ws || ws_bytes < need) VB_THROW("wav_to_int16: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wav_peak_kernel<<<B, 256, 0, st>>>(wav, wav_ld, n_frames, hop_length, w.peak);
  check_launch("wav_peak_kernel");
  wav_to_int16_kernel<<<dim3(static_cast<unsigned>((wav_ld + 255) / 256), B), 256, 0, st>>>(wav, wav_ld, n_frames, hop_length,
                                                                                             w.peak, out);
  check_launch("wav_to_int16_kernel");
  API_END
}

long vaenar_launch_count(void) { return g_launch_count; }

// Debug/tuning: device buffer (8 x u64 per CTA) receiving per-phase globaltimer stamps of every GEMM CTA; null disables.
int vaenar_debug_gemm_timestamps(void* dev_buf) {
  g_dbg_cursor = static_cast<unsigned long long*>(dev_buf);
  g_dbg_launches.clear();
  return 0;
}
// Fused row kernel: consecutive launches write 128 x u64 per CTA (grid order) into dev_buf; null disables.
int vaenar_debug_xrow_timestamps(void* dev_buf) {
  g_xrow_dbg = static_cast<unsigned long long*>(dev_buf);
  return 0;
}
// JSON list of the GEMM launches recorded since the buffer was set: [{"grid": [x, y], "cls": ..., "offset": u64 index}]
const char* vaenar_debug_gemm_launches(void) {
  static std::string out;
  out = "[";
  for (size_t i = 0; i < g_dbg_launches.size(); ++i) out += (i ? ", " : "") + g_dbg_launches[i];
  out += "]";
  return out.c_str();
}

// Enable (1) / disable (0) per-launch CUDA-event timing of the tensor-core kernels.  Not for use during
// graph capture.  vaenar_profile_report() synchronises the device and returns a JSON object
// {class: {launches, ms, flops, bytes}} for the launches seen since the last enable.
int vaenar_profile_enable(int on) {
  if (on) {
    for (auto& kv : g_stats)
      for (auto& ev : kv.second.events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    g_stats.clear();
  }
  g_profile = on != 0;
  return 0;
}
const char* vaenar_profile_report(void) {
  cudaDeviceSynchronize();
  std::string out = "{";
  bool first = true;
  for (auto& kv : g_stats) {
    double ms = 0;
    for (auto& ev : kv.second.events) {
      float t = 0;
      if (cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) ms += t;
    }
    char buf[512];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
             first ? "" : ", ", kv.first.c_str(), kv.second.launches, ms, kv.second.flops, kv.second.bytes);
    out += buf;
    first = false;
  }
  out += "}";
  g_profile_json = out;
  return g_profile_json.c_str();
}

// ---------------------------------------------------------------------------- block-level test hooks
struct TestCtx : Ctx {
  vaenar_model dummy;
};
static void test_ctx(TestCtx& c, void* ws, int64_t ws_bytes, void* stream) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) VB_THROW("no CUDA device: the B200 path has no CPU fallback");
  set_attrs(nullptr);
  c.m = &c.dummy; c.params = nullptr; c.packed = nullptr;
  c.ws = static_cast<uint8_t*>(ws); c.ws_bytes = ws_bytes; c.stream = static_cast<cudaStream_t>(stream);
}
static void launch_pack(Ctx& c, const std::vector<PackOp>& ops, PackOp* dev_ops) {
  VB_CUDA(cudaMemcpyAsync(dev_ops, ops.data(), ops.size() * sizeof(PackOp), cudaMemcpyHostToDevice, c.stream));
  int maxK = 0, maxN = 0;
  for (auto& o : ops) { maxK = std::max(maxK, o.K); maxN = std::max(maxN, o.N); }
  dim3 grid(cdiv(maxN, 32), cdiv(maxK, 32), static_cast<unsigned>(ops.size()));
  pack_weights_kernel<<<grid, dim3(32, 8), 0, c.stream>>>(dev_ops);
  check_launch("pack_weights(test)");
}
// hi/lo fp16 split of an fp32 activation matrix (test-only producer; the model's epilogues emit these directly)
__global__ void split_f16_kernel(const float* in, __half* hi, __half* lo, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const __half h = __float2half_rn(in[i]);
    hi[i] = h;
    if (lo) lo[i] = __float2half_rn(in[i] - __half2float(h));
  }
}

int vaenar_test_dense(const float* A, const float* W, const float* bias, const float* residual, const float* gamma,
                      const float* beta, int M, int K, int N, int act, int ln, int split, int block_n, float* out,
                      void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  TestCtx c;
  test_ctx(c, ws, ws_bytes, stream);
  const int Kp = kpad64(K), parts = split ? 3 : 1;
  __half* Ah = c.alloc<__half>(static_cast<int64_t>(M) * K);
  __half* Al = c.alloc<__half>(static_cast<int64_t>(M) * K);
  __half* Wp = c.alloc<__half>(static_cast<int64_t>(N) * parts * Kp);
  PackOp* dops = c.alloc<PackOp>(4);
  VB_CUDA(cudaMemsetAsync(Wp, 0, static_cast<int64_t>(N) * parts * Kp * 2, c.stream));
  std::vector<PackOp> ops;
  for (int p = 0; p < parts; ++p) ops.push_back(PackOp{W, Wp + p * Kp, K, N, N, parts * Kp, (split && p == 2) ? 1 : 0});
  launch_pack(c, ops, dops);
  const long n = static_cast<long>(M) * K;
  split_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(A, Ah, split ? Al : nullptr, n);
  check_launch("split");
  GemmParams p = gp();
  p.mode = ln ? EPI_LN : EPI_PLAIN; p.N = N; p.act = act; p.bias = bias; p.residual = residual; p.res_ld = N;
  p.ln_gamma = gamma; p.ln_beta = beta; p.out_f32 = out; p.ld_f32 = N;
  if (ln) { p.out_h = c.alloc<__half>(static_cast<int64_t>(M) * N); p.ld_h = N; }
  p.nseg = parts;
  for (int q = 0; q < parts; ++q) { p.seg_map[q] = (q == 1) ? 1 : 0; p.seg_shift[q] = 0; p.seg_kblocks[q] = Kp / 64; }
  run_gemm(c, block_n, AOp{Ah, K, K}, AOp{split ? Al : nullptr, K, K}, 1, M, Wp, parts * Kp, N, p);
  API_END
}

int vaenar_test_conv1d(const float* X, const float* W, const float* bias, int B, int T, int Cin, int Cout, int taps,
                       int act, int split, float* out, void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  TestCtx c;
  test_ctx(c, ws, ws_bytes, stream);
  const int cp = kpad64(Cin), parts = split ? 3 : 1, Ktot = taps * parts * cp;
  if (taps * parts > kMaxSegs) VB_THROW("too many segments");
  const long n = static_cast<long>(B) * T * Cin;
  __half* Xh = c.alloc<__half>(n);
  __half* Xl = c.alloc<__half>(n);
  __half* Wp = c.alloc<__half>(static_cast<int64_t>(Cout) * Ktot);
  PackOp* dops = c.alloc<PackOp>(taps * parts);
  VB_CUDA(cudaMemsetAsync(Wp, 0, static_cast<int64_t>(Cout) * Ktot * 2, c.stream));
  std::vector<PackOp> ops;
  for (int j = 0; j < taps; ++j)
    for (int p = 0; p < parts; ++p)
      ops.push_back(PackOp{W + static_cast<long>(j) * Cin * Cout, Wp + (j * parts + p) * cp, Cin, Cout, Cout, Ktot,
                           (split && p == 2) ? 1 : 0});
  launch_pack(c, ops, dops);
  split_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(X, Xh, split ? Xl : nullptr, n);
  check_launch("split");
  GemmParams p = gp();
  p.mode = EPI_PLAIN; p.N = Cout; p.act = act; p.bias = bias; p.out_f32 = out; p.ld_f32 = Cout;
  segs_conv(p, taps, Cin, split != 0);
  run_gemm(c, 128, AOp{Xh, Cin, Cin}, AOp{split ? Xl : nullptr, Cin, Cin}, B, T, Wp, Ktot, Cout, p);
  API_END
}

// [B, T, H*64] fp32 -> V^T fp16 [B*H*64, tpad]
__global__ void vt_transpose_kernel(const float* v, __half* vt, int B, int T, int H, int tpad) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long n = static_cast<long>(B) * T * H * 64;
  if (i >= n) return;
  const int c = static_cast<int>(i % (H * 64));
  const long bt = i / (H * 64);
  const int t = static_cast<int>(bt % T), b = static_cast<int>(bt / T);
  vt[(static_cast<long>(b) * H * 64 + c) * tpad + t] = __float2half_rn(v[i]);
}
__global__ void half_to_float_kernel(const __half* in, float* out, long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half2float(in[i]);
}

int vaenar_test_attention(const float* q, const float* k, const float* v, const int32_t* q_len,
                          const int32_t* k_len, int B, int H, int Tq, int Tk, int causal, float* ctx, float* ali,
                          void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  TestCtx c;
  test_ctx(c, ws, ws_bytes, stream);
  const int D = H * 64, tpad = vt_pad(Tk);
  __half* qh = c.alloc<__half>(static_cast<int64_t>(B) * Tq * D);
  __half* kh = c.alloc<__half>(static_cast<int64_t>(B) * Tk * D);
  __half* vt = c.alloc<__half>(static_cast<int64_t>(B) * D * tpad);
  __half* ch = c.alloc<__half>(static_cast<int64_t>(B) * Tq * D);
  run_cast(c, q, qh, static_cast<int64_t>(B) * Tq * D);
  run_cast(c, k, kh, static_cast<int64_t>(B) * Tk * D);
  const long nv = static_cast<long>(B) * Tk * D;
  VB_CUDA(cudaMemsetAsync(vt, 0xFF, static_cast<int64_t>(B) * D * tpad * 2, c.stream));   // NaN padding: must never leak
  vt_transpose_kernel<<<static_cast<unsigned>((nv + 255) / 256), 256, 0, c.stream>>>(v, vt, B, Tk, H, tpad);
  check_launch("vt_transpose");
  // causal self-attention masks keys with the query lengths (attention.py:437-439 passes the same tensor twice)
  run_attention(c, B, H, AttnCall{qh, D, 0, Tq, kh, D, 0, Tk, vt, static_cast<long>(B) * D, tpad, 0, q_len,
                                  (causal && Tq == Tk) ? q_len : k_len, causal, ch, D, ali});
  const long nc = static_cast<long>(B) * Tq * D;
  half_to_float_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, c.stream>>>(ch, ctx, nc);
  check_launch("half_to_float");
  API_END
}

// Selects the fused per-block row kernel (1, default) or the per-op launch chain (0) for every later forward call.
int vaenar_set_fused(int on) {
  set_attrs(nullptr);
  g_use_fused = on != 0;
  return 0;
}

// Plain-epilogue GEMMs: 1 (default) the two-CTAs-per-SM instances for grids deeper than one wave, 0 never, 2 always (parity
// tests at small shapes).
int vaenar_set_gemm_occ2(int mode) {
  set_attrs(nullptr);
  g_gemm_occ2 = mode;
  return 0;
}

// Training forward of a CrossAttentionBLK: -1 (default) fused row kernel when the row tiles fill the chip, 0 per-op chain
// always, 1 fused row kernel whenever the shapes allow it (parity tests at small shapes).
int vaenar_set_train_fused(int mode) {
  g_train_fused = mode < 0 ? -1 : (mode != 0);
  return 0;
}

/* The CrossAttentionBLK stack of one module on caller-supplied activations (block-level parity hook):
 * module 0 = decoder.attentions, 1 = posterior.attentions, 2 + s = prior.glow.s.affine_coupling.net.attentions. */
int vaenar_xblk_stack_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes, int module,
                          float* x, const float* text_embd, const int32_t* q_lengths, const int32_t* text_lengths, int B, int T,
                          int T_text, float* alignments, void* stream) {
  API_BEGIN
  Ctx c = make_ctx(h, params, packed, ws, ws_bytes, stream);
  const vaenar_hparams_t& hp = h->hp;
  const int E = hp.enc_hidden;
  std::string pk, pn, kvname;
  int nblk, d, H, F, kv_nblk, kv_blk0 = 0;
  if (module == 0) {
    pk = "dec.blk"; pn = "decoder.attentions."; kvname = "dec.kv";
    nblk = kv_nblk = hp.dec_nblk; d = hp.dec_att_dim; H = hp.dec_heads; F = hp.dec_ffn;
  } else if (module == 1) {
    pk = "post.blk"; pn = "posterior.attentions."; kvname = "post.kv";
    nblk = kv_nblk = hp.posterior_nblk; d = hp.posterior_att_dim; H = hp.posterior_heads; F = hp.posterior_ffn;
  } else if (module >= 2 && module < 2 + hp.prior_n_blk) {
    const int s = module - 2;
    pk = "prior." + std::to_string(s) + ".blk"; pn = "prior.glow." + std::to_string(s) + ".affine_coupling.net.attentions.";
    kvname = "prior.kv";
    nblk = hp.prior_n_tblk; kv_nblk = hp.prior_n_blk * hp.prior_n_tblk; kv_blk0 = s * hp.prior_n_tblk;
    d = hp.prior_att_dim; H = hp.prior_heads; F = hp.prior_ffn;
  } else {
    VB_THROW("unknown module %d", module);
  }
  const int64_t rows = static_cast<int64_t>(B) * T;
  __half* emb_h = c.alloc<__half>(static_cast<int64_t>(B) * T_text * E);
  run_cast(c, text_embd, emb_h, static_cast<int64_t>(B) * T_text * E);
  Stream2 xs{x, c.alloc<__half>(rows * d)};
  run_cast(c, x, xs.h, rows * d);
  XblkBufs xb = xblk_bufs(c, B, T, d, H, F);
  MemKV kv = memory_kv(c, kvname, emb_h, B, T_text, E, kv_nblk, d, H);
  xblk_stack_fwd(c, pk, pn, nblk, xs, xb, B, T, d, H, F, q_lengths, kv, kv_blk0, T_text, text_lengths, alignments,
                 static_cast<int64_t>(B) * H * T * T_text);
  API_END
}

/* dW[taps][Cin][Cout] = sum_{b,t} X[b, t + tap - (taps-1)/2, :]^T dY[b, t, :]  (taps = 1: Dense weight gradient).
 * X [B,T,Cin], dY [B,T,Cout] fp32 (cast to fp16 operands here); X2 (nullable) [B,T,Cin2] supplies rows Cin.. of a
 * concat-Dense gradient. */
int vaenar_test_wgrad(const float* X, const float* X2, const float* dY, int B, int T, int Cin, int Cin2, int Cout, int taps,
                      float* dW, void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  TestCtx c;
  test_ctx(c, ws, ws_bytes, stream);
  const int64_t tok = static_cast<int64_t>(B) * T;
  __half* xh = c.alloc<__half>(tok * Cin);
  __half* x2h = X2 ? c.alloc<__half>(tok * Cin2) : nullptr;
  __half* dyh = c.alloc<__half>(tok * Cout);
  run_cast(c, X, xh, tok * Cin);
  if (X2) run_cast(c, X2, x2h, tok * Cin2);
  run_cast(c, dY, dyh, tok * Cout);
  const int M = Cin + (X2 ? Cin2 : 0);
  VB_CUDA(cudaMemsetAsync(dW, 0, static_cast<int64_t>(taps) * M * Cout * 4, c.stream));
  for (int j = 0; j < taps; ++j)
    run_wgrad(c, WOp{xh, Cin, Cin, 0}, X2 ? WOp{x2h, Cin2, Cin2, 0} : WOp{}, Cin, WOp{dyh, Cout, Cout, 0}, B, T,
              j - (taps - 1) / 2, M, Cout, dW + static_cast<int64_t>(j) * M * Cout, Cout);
  API_END
}

/* Backward of the attention core: q,k,v,dctx fp32 [B,T,H*64] -> dq,dk,dv fp32 (computed on fp16 operands). */
int vaenar_test_attention_bwd(const float* q, const float* k, const float* v, const float* dctx, const int32_t* q_len,
                              const int32_t* k_len, int B, int H, int Tq, int Tk, int causal, float* dq, float* dk,
                              float* dv, void* ws, int64_t ws_bytes, void* stream) {
  API_BEGIN
  TestCtx c;
  test_ctx(c, ws, ws_bytes, stream);
  const int D = H * 64, tpad = vt_pad(Tk);
  const int64_t nq = static_cast<int64_t>(B) * Tq * D, nk = static_cast<int64_t>(B) * Tk * D;
  __half* qh = c.alloc<__half>(nq);
  __half* kh = c.alloc<__half>(nk);
  __half* vh = c.alloc<__half>(nk);
  __half* doh = c.alloc<__half>(nq);
  __half* vt = c.alloc<__half>(static_cast<int64_t>(B) * D * tpad);
  __half* ch = c.alloc<__half>(nq);
  float* lse2 = c.alloc<float>(static_cast<int64_t>(B) * H * Tq);
  float* delta = c.alloc<float>(static_cast<int64_t>(B) * H * Tq);
  __half* dqh = c.alloc<__half>(nq);
  __half* dkh = c.alloc<__half>(nk);
  __half* dvh = c.alloc<__half>(nk);
  run_cast(c, q, qh, nq);
  run_cast(c, k, kh, nk);
  run_cast(c, v, vh, nk);
  run_cast(c, dctx, doh, nq);
  VB_CUDA(cudaMemsetAsync(vt, 0, static_cast<int64_t>(B) * D * tpad * 2, c.stream));
  vt_transpose_kernel<<<static_cast<unsigned>((nk + 255) / 256), 256, 0, c.stream>>>(v, vt, B, Tk, H, tpad);
  check_launch("vt_transpose");
  AttnCall f{qh, D, 0, Tq, kh, D, 0, Tk, vt, static_cast<long>(B) * D, tpad, 0, q_len, k_len, causal, ch, D, nullptr};
  f.lse2 = lse2;
  run_attention(c, B, H, f);
  VB_CUDA(cudaMemsetAsync(dqh, 0xFF, nq * 2, c.stream));   // NaN fill: every element must be written
  VB_CUDA(cudaMemsetAsync(dkh, 0xFF, nk * 2, c.stream));
  VB_CUDA(cudaMemsetAsync(dvh, 0xFF, nk * 2, c.stream));
  run_attention_bwd(c, B, H, AttnBwdCall{qh, D, 0, Tq, kh, D, 0, Tk, vh, D, 0, ch, D, doh, D, 0, q_len, k_len, causal, lse2,
                                         delta, dqh, D, 0, dkh, D, 0, dvh, D, 0});
  half_to_float_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, c.stream>>>(dqh, dq, nq);
  half_to_float_kernel<<<static_cast<unsigned>((nk + 255) / 256), 256, 0, c.stream>>>(dkh, dk, nk);
  half_to_float_kernel<<<static_cast<unsigned>((nk + 255) / 256), 256, 0, c.stream>>>(dvh, dv, nk);
  check_launch("half_to_float");
  API_END
}

}  // extern "C"

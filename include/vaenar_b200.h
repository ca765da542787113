/* vaenar_b200.h -- C ABI of libvaenar_sm100.so: the B200-native (sm_100a) implementation of the
 * VAENAR-TTS non-autoregressive mel-synthesis hot path.
 *
 * The reference (thuhcsi/VAENAR-TTS) has no FFI layer: the operator boundary of the path is the Keras
 * object API of models/models.py:9-226 as consumed by train.py:120-179 and inference.py:55-72,125-143.
 * Each entry point below replaces one of those Python call sites (cited per function); the Python mirror
 * (vaenar_tts_b200/model.py) binds them with ctypes and re-exposes the reference's names.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on error; vaenar_last_error() gives the text;
 *   - the CALLER owns all memory: parameters (one flat fp32 buffer laid out by the manifest below), the
 *     packed fp16 weight arena, the workspace, all inputs and outputs.  Nothing is allocated or freed here
 *     and no call synchronises the host with the device;
 *   - all pointers are device pointers unless stated; tensors are row-major [batch, time, channel] fp32,
 *     token ids and lengths int32; `stream` is a cudaStream_t passed as void*;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef VAENAR_B200_H_
#define VAENAR_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vaenar_model* vaenar_handle_t;

/* Mirrors the values models/models.py:16-65 reads from configs/hparams.py:233-348 (LJHPS) / :351-474. */
typedef struct vaenar_hparams {
  int32_t vocab_size, embd_dim, enc_n_conv, enc_hidden, enc_conv_kernel, enc_n_blk, enc_att_dim, enc_heads, enc_ffn;
  int32_t dec_nblk, dec_att_dim, dec_heads, dec_ffn, post_n_conv, post_filters, post_kernel;
  int32_t posterior_pre_hidden, posterior_nblk, posterior_att_dim, posterior_heads, posterior_ffn;
  int32_t prior_n_blk, prior_n_tblk, prior_att_dim, prior_heads, prior_ffn;
  int32_t latent_dim, out_dim, max_reduction_factor, final_reduction_factor;
  float mel_text_len_ratio;
  /* dropout rates (configs/hparams.py:299-300,321,326-327): encoder prenet / positional, posterior prenet / positional, postnet */
  float enc_pre_drop_rate, enc_pos_drop_rate, posterior_pre_drop_rate, posterior_pos_drop_rate, post_drop_rate;
} vaenar_hparams_t;

/* Training-mode forward options.  Dropout keep-masks (values 0 or 1/(1-rate), fp32, shape of the tensor they multiply)
 * are consumed in the reference's call order: encoder prenet conv 0..n-1, encoder positional, [posterior prenet 1, 2,
 * posterior positional,] postnet conv 0..n-1.  masks == NULL: generated on device from (seed, site index). */
typedef struct vaenar_train_opts {
  int32_t n_masks;
  const float* const* masks; /* HOST array of device pointers, or NULL */
  uint64_t seed;
  int32_t update_bn_stats;   /* apply the BatchNorm moving-average side effect (momentum 0.99) */
} vaenar_train_opts_t;

const char* vaenar_last_error(void);
int vaenar_abi_version(void);

/* VAENAR.__init__ (models/models.py:10-65): builds the parameter manifest; allocates no device memory. */
int vaenar_create(const vaenar_hparams_t* hps, vaenar_handle_t* out);
int vaenar_destroy(vaenar_handle_t h);

/* Parameter manifest: model.trainable_variables + BN moving stats (train.py:136-137), attribute-path
 * names of SURVEY.md Appendix B, Keras layouts (Dense [in,out], Conv1D [k,in,out]).  All parameters live
 * in ONE flat fp32 buffer; offsets are in floats. */
int vaenar_num_params(vaenar_handle_t h);
const char* vaenar_param_name(vaenar_handle_t h, int i);
int vaenar_param_ndim(vaenar_handle_t h, int i);
int64_t vaenar_param_dim(vaenar_handle_t h, int i, int d);
int64_t vaenar_param_offset(vaenar_handle_t h, int i);
int vaenar_param_trainable(vaenar_handle_t h, int i);
int64_t vaenar_param_floats(vaenar_handle_t h);

int64_t vaenar_packed_bytes(vaenar_handle_t h);
int64_t vaenar_workspace_bytes(vaenar_handle_t h, int B, int T_text, int T_z, int rf);

/* Re-layout of the fp32 master weights into the kernel operand formats (fp16 [out,in] K-major matrices,
 * folded inference BatchNorm, ActNorm(+)InvertibleLinear folded 128x128 maps, float64 log|det W| and the
 * fp32 inverse of modules/flow.py:126-144).  Call after every change of the parameters. */
int vaenar_pack_weights(vaenar_handle_t h, const float* params, void* packed, void* stream);
/* Same, but the flow constants (128x128 LU / inverse, latency-bound single-CTA kernels) keep running on an internal
 * stream when the call returns; the NEXT call of this library with the same handle orders them before its own work
 * (vaenar_train_step_grads does so right before the prior, so that they overlap the encoder / posterior / decoder
 * forward).  Do not use before replaying a captured CUDA graph: the replay does not pass through the library. */
int vaenar_pack_weights_async(vaenar_handle_t h, const float* params, void* packed, void* stream);

/* TransformerEncoder.call (modules/encoder.py:79-93), training=False: text_embd [B, T_text, embd]. */
int vaenar_text_encoder_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const int32_t* texts, const int32_t* text_lengths, int B, int T_text, float pos_step,
                            float* text_embd, void* stream);

/* DenseLengthPredictor.call (modules/length_predictor.py:35-42): predicted lengths [B] (float). */
int vaenar_length_predictor_fwd(vaenar_handle_t h, const float* params, const float* text_embd,
                                const int32_t* text_lengths, int B, int T_text, float* pred_lengths, void* stream);

/* TransformerPrior.sample (modules/prior.py:154-169).  z_io holds the initial noise epsilon
 * [B, T_z, latent] on entry (already scaled by the temperature) and the sampled latents on return;
 * logp [B] receives the log-density. */
int vaenar_prior_sample(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                        const float* text_embd, const int32_t* text_lengths, const int32_t* z_lengths, int B,
                        int T_text, int T_z, float* z_io, float* logp, void* stream);

/* TransformerPrior.log_probability (modules/prior.py:119-152): logp [B] of given latents z [B,T_z,latent]
 * (z is not modified). */
int vaenar_prior_log_probability(vaenar_handle_t h, const float* params, const void* packed, void* ws,
                                 int64_t ws_bytes, const float* z, const float* text_embd,
                                 const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z,
                                 float* logp, void* stream);

/* TransformerPosterior.call + reparameterize + log_probability (modules/posterior.py:20-72,115-130) as
 * wired by VAENAR.call (models/models.py:123,136-144; the (logvar, mu) name swap of :136 is honoured):
 * mels [B,T_mel,80], eps [B,T_z,latent] -> z [B,T_z,latent], logq [B].  training=False (no dropout). */
int vaenar_posterior_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                         const float* mels, const float* text_embd, const int32_t* text_lengths,
                         const int32_t* z_lengths, const float* eps, int B, int T_text, int T_mel, int T_z, int rf,
                         float* z, float* logq, void* stream);

/* TransformerPosterior.call alone (modules/posterior.py:115-130), training=False: the outputs of mu_projection and
 * logvar_projection [B, T_z, latent] exactly as the reference returns them (`return mu, logvar, None`, :130).  Note that
 * VAENAR.call unpacks them swapped (models/models.py:136); this entry point does NOT swap.  reduced_mels [B, T_z, 80]. */
int vaenar_posterior_params(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const float* reduced_mels, const float* text_embd, const int32_t* text_lengths,
                            const int32_t* z_lengths, int B, int T_text, int T_z, float* mu_projection_out,
                            float* logvar_projection_out, void* stream);

/* TransformerDecoder.call (modules/decoder.py:181-199), training=False: initial / final mel
 * [B, T_z*rf, 80]; alignments (nullable) [dec_nblk, B, heads, T_z, T_text]. */
int vaenar_decoder_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                       const float* z, const float* text_embd, const int32_t* z_lengths,
                       const int32_t* text_lengths, int B, int T_text, int T_z, int rf, float* initial_mel,
                       float* mel, float* alignments, void* stream);

/* VAENAR.inference (models/models.py:199-210): encoder -> prior.sample -> decoder in one call (one CUDA
 * graph capturable launch sequence).  z_io: epsilon in / latents out. */
int vaenar_inference(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                     const int32_t* texts, const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text,
                     int T_z, int rf, float* z_io, float* text_embd, float* mel, float* alignments, float* logp,
                     void* stream);

/* VAENAR.call forward (models/models.py:105-197), training=False, n_sample=1, per-example losses:
 * l2 [B], kl [B], length_loss [B]; decoded mel [B, T_mel, 80] (cropped to T_mel). */
int vaenar_elbo_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes,
                    const int32_t* texts, const float* mels, const int32_t* mel_lengths,
                    const int32_t* text_lengths, const int32_t* z_lengths, const float* eps, int B, int T_text,
                    int T_mel, int T_z, int rf, float* mel_out, float* l2, float* kl, float* length_loss,
                    float* alignments, void* stream);

/* VAENAR.call forward with training=True semantics (models/models.py:105-197 as called by train.py:129-134):
 * BatchNorm uses batch statistics over (batch, time) incl. padding and updates its moving averages in `params`,
 * dropout is active (encoder, posterior prenet/positional, postnet).  Forward only (dev-style evaluation of the
 * training-mode arithmetic); the gradients come from vaenar_train_step_grads. */
int vaenar_elbo_fwd_train(vaenar_handle_t h, float* params, const void* packed, void* ws, int64_t ws_bytes,
                          const int32_t* texts, const float* mels, const int32_t* mel_lengths,
                          const int32_t* text_lengths, const int32_t* z_lengths, const float* eps, int B, int T_text,
                          int T_mel, int T_z, int rf, const vaenar_train_opts_t* opts, float* mel_out, float* l2,
                          float* kl, float* length_loss, float* alignments, void* stream);

/* VAENAR.init (models/models.py:212-226): encoder(training=True) -> prior.init (data-dependent ActNorm
 * initialisation, modules/prior.py:171-186, modules/flow.py:189-196, written into `params`) -> decoder at
 * rf = max_reduction_factor.  z_io: N(0,1) noise in, latents out; mel [B, T_z*max_rf, 80].  The packed arena must be
 * rebuilt (vaenar_pack_weights) afterwards. */
int vaenar_init(vaenar_handle_t h, float* params, void* packed, void* ws, int64_t ws_bytes, const int32_t* texts,
                const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z,
                const vaenar_train_opts_t* opts, float* z_io, float* mel, void* stream);

/* TransformerPrior.init alone (modules/prior.py:171-186): z_io holds N(0,1) noise on entry and the flow output on return;
 * every ActNorm's log_scale / bias (modules/flow.py:189-196) is written into `params`, the folded flow maps into `packed`.
 * Re-pack (vaenar_pack_weights) afterwards. */
int vaenar_prior_init(vaenar_handle_t h, float* params, void* packed, void* ws, int64_t ws_bytes, const float* text_embd,
                      const int32_t* text_lengths, const int32_t* z_lengths, int B, int T_text, int T_z, float* z_io,
                      void* stream);

/* train_step of train.py:120-138 up to the gradients: VAENAR.call(training=True) forward that records what the
 * backward pass needs, the losses (losses[4] = {total, mel_l2, kl, length_l2}, total = mel_l2 + kl_weight * max(kl, 0) +
 * length_weight * length_l2, train.py:135) and the hand-written backward pass.  `grads` (flat parameter layout, fully
 * overwritten) receives loss_scale * dTotal/dparams: all activation gradients are carried multiplied by loss_scale so
 * that their fp16 tensor-core operand copies stay in range; pass 1/loss_scale (x 1/world_size) as grad_scale to
 * vaenar_adam_step.  BatchNorm moving averages are updated in `params` (opts->update_bn_stats).  mel_out (nullable):
 * decoded mel [B, T_mel, 80].  Workspace: vaenar_train_workspace_bytes (holds every saved activation). */
int64_t vaenar_train_workspace_bytes(vaenar_handle_t h, int B, int T_text, int T_z, int rf);
int vaenar_train_step_grads(vaenar_handle_t h, float* params, const void* packed, void* ws, int64_t ws_bytes,
                            const int32_t* texts, const float* mels, const int32_t* mel_lengths, const int32_t* text_lengths,
                            const int32_t* z_lengths, const float* eps, int B, int T_text, int T_mel, int T_z, int rf,
                            const vaenar_train_opts_t* opts, float kl_weight, float length_weight, float loss_scale, float* grads,
                            float* losses, float* mel_out, void* stream);

/* Optimiser half of train_step (train.py:116-117,137): Keras Adam over the flat parameter buffer, one fused launch.
 * `trainable_mask` (device, one byte per float; host copy from vaenar_trainable_mask) skips the BatchNorm moving
 * statistics and padding; grad_scale folds the 1/world_size of the data-parallel gradient mean.  The gradients
 * come from vaenar_train_step_grads. */
int vaenar_trainable_mask(vaenar_handle_t h, uint8_t* host_mask);
/* skip_flag (device, nullable): when *skip_flag != 0 the whole update is skipped (non-finite gradients, see below). */
int vaenar_adam_step(float* params, const float* grads, float* m, float* v, const uint8_t* trainable_mask, int64_t n,
                     int64_t step, float lr, float beta1, float beta2, float eps, float grad_scale, const float* skip_flag,
                     void* stream);
/* Overflow guard of the loss-scaled fp16 backward pass (no counterpart in the fp32 reference): adds the number of
 * non-finite entries of grads[0, n) to *count (device float, zeroed by the caller).  Data-parallel runs all-reduce the
 * count so that every replica skips the same step; the host backs the loss scale off (VAENAR.train_step). */
int vaenar_grad_nonfinite(const float* grads, int64_t n, float* count, void* stream);

/* Data-parallel variant (one process per GPU, world <= 8): the gradient exchange and the optimiser in ONE kernel over
 * NVLink peer memory.  peer_params / peer_grads: HOST arrays of `world` device pointers to every replica's flat parameter
 * / gradient buffer (entry `rank` = the local ones; the others are CUDA-IPC mappings of the peers' buffers).  This rank
 * reduces the gradient shard [rank * S, (rank+1) * S), S = vaenar_adam_shard_floats(n, world), reading the peers' HBM
 * directly, applies Keras Adam with ITS shard of the moments (m_shard, v_shard: S floats each -- optimiser state is
 * sharded across the job) and writes the new parameters into every replica.  grad_scale = 1 / (loss_scale * world).
 * The caller orders the call between two cross-rank barriers on the stream (all gradients complete before; all peer
 * writes complete after).  Replaces all_reduce + vaenar_adam_step of the N > 1 train_step (train.py:136-137). */
int vaenar_enable_peer_access(int peer_device);   /* kernels of the current device may dereference peer_device's memory */
int vaenar_ipc_export(const void* ptr, void* cuda_ipc_mem_handle_64_bytes_out, int64_t* offset_out);
int vaenar_ipc_open(const void* cuda_ipc_mem_handle_64_bytes, void** out_ptr);   /* cudaIpcOpenMemHandle on the current device */
int vaenar_ipc_close(void* ptr);
int64_t vaenar_adam_shard_floats(int64_t n, int world);
int vaenar_adam_step_sharded(float* const* peer_params, const float* const* peer_grads, float* m_shard, float* v_shard,
                             const uint8_t* trainable_mask, int64_t n, int rank, int world, int64_t step, float lr, float beta1,
                             float beta2, float eps, float grad_scale, const float* skip_flag, void* stream);

/* CRC32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of the TF tensor-bundle /
 * TFRecord formats read by vaenar_tts_b200/tf_checkpoint.py (tf.train.Checkpoint files of train.py:246-248). */
uint32_t vaenar_crc32c(const void* data, int64_t n, uint32_t crc);

/* N(0, stddev) noise from the counter-based generator (replaces tf.random.normal at
 * modules/posterior.py:35 and modules/prior.py:35). */
int vaenar_randn(float* out, int64_t n, uint64_t seed, uint64_t stream_id, float stddev, void* stream);

/* Accounting for bench.py: number of kernels this library has launched (captured launches count once at
 * capture time), and per-launch CUDA-event timing of the tensor-core kernel classes. */
long vaenar_launch_count(void);
int vaenar_profile_enable(int on);
const char* vaenar_profile_report(void);
/* Tuning aid: per-CTA phase timestamps (8 x u64, globaltimer ns) of the GEMM kernel into dev_buf; NULL disables. */
int vaenar_debug_gemm_timestamps(void* dev_buf);
const char* vaenar_debug_gemm_launches(void);
/* Tuning aid: per-CTA phase timestamps (128 x u64) of consecutive fused row-kernel launches into dev_buf; NULL disables. */
int vaenar_debug_xrow_timestamps(void* dev_buf);

/* ---- block-level entry points (parity tests of single kernels against the oracle) ---- */
/* out = act(A[M,K] W[K,N] + bias) (+residual, LayerNorm if ln != 0); A, W fp32 host-layout device buffers;
 * the call converts operands to fp16 into `ws` and runs the tcgen05 GEMM.  split != 0 uses the split-fp16 path. */
int vaenar_test_dense(const float* A, const float* W, const float* bias, const float* residual, const float* gamma,
                      const float* beta, int M, int K, int N, int act, int ln, int split, int block_n, float* out,
                      void* ws, int64_t ws_bytes, void* stream);
/* Conv1D k taps 'same' over [B,T,Cin] -> [B,T,Cout] fp32 (bias + act), implicit GEMM on tcgen05. */
int vaenar_test_conv1d(const float* X, const float* W, const float* bias, int B, int T, int Cin, int Cout, int taps,
                       int act, int split, float* out, void* ws, int64_t ws_bytes, void* stream);
/* MultiHeadScaledProductAttention core (modules/attention.py:217-246) on projected q/k/v fp32
 * [B,Tq,H*64], [B,Tk,H*64]: ctx [B,Tq,H*64] fp32, ali (nullable) [B,H,Tq,Tk]. */
int vaenar_test_attention(const float* q, const float* k, const float* v, const int32_t* q_len,
                          const int32_t* k_len, int B, int H, int Tq, int Tk, int causal, float* ctx, float* ali,
                          void* ws, int64_t ws_bytes, void* stream);

/* CrossAttentionBLK stack of one module (modules/attention.py:418-452) on caller-supplied activations:
 * module 0 = decoder.attentions (modules/decoder.py:170-174), 1 = posterior.attentions (modules/posterior.py:100-106),
 * 2 + s = prior.glow[s].affine_coupling.net.attentions (modules/transform.py:37-43).  x [B,T,256] fp32 is updated in
 * place; text_embd [B,T_text,512] fp32 is the attention memory; alignments (nullable) [nblk,B,heads,T,T_text]. */
int vaenar_xblk_stack_fwd(vaenar_handle_t h, const float* params, const void* packed, void* ws, int64_t ws_bytes, int module,
                          float* x, const float* text_embd, const int32_t* q_lengths, const int32_t* text_lengths, int B, int T,
                          int T_text, float* alignments, void* stream);
/* 1 (default): one fused tcgen05 row kernel per CrossAttentionBLK (csrc/xblk_fused.cuh); 0: the per-op launch chain. */
int vaenar_set_fused(int on);
/* Training forward of a CrossAttentionBLK (the tape-recording half of train.py:129-134): -1 = automatic (the fused row
 * kernel with tape outputs when B * ceil(T / 128) >= 96 row tiles, else the per-op launch chain), 0 = per-op chain,
 * 1 = fused row kernel whenever the shapes allow it.  Both produce the same tape; the parity tests run both. */
int vaenar_set_train_fused(int mode);
/* Plain-epilogue GEMM instances built for two resident CTAs per SM (csrc/gemm_tc.cuh, GemmCfg<128, 2>): 1 = for grids deeper
 * than one wave (default), 0 = never, 2 = always. */
int vaenar_set_gemm_occ2(int mode);

/* Weight gradient of a Dense / Conv1D layer on tcgen05 with MN-major operands (csrc/wgrad_tc.cuh):
 * dW[tap][Cin (+Cin2)][Cout] = sum_{b,t} [X ; X2][b, t + tap - (taps-1)/2, :]^T dY[b, t, :]  (fp32 out, fp16 operands). */
int vaenar_test_wgrad(const float* X, const float* X2, const float* dY, int B, int T, int Cin, int Cin2, int Cout, int taps,
                      float* dW, void* ws, int64_t ws_bytes, void* stream);

/* Backward of the attention core (csrc/attention_bwd_tc.cuh; forward recomputed for the softmax statistics):
 * dq [B,Tq,H*64], dk, dv [B,Tk,H*64] from the context gradient dctx [B,Tq,H*64]. */
int vaenar_test_attention_bwd(const float* q, const float* k, const float* v, const float* dctx, const int32_t* q_len,
                              const int32_t* k_len, int B, int H, int Tq, int Tk, int causal, float* dq, float* dk,
                              float* dv, void* ws, int64_t ws_bytes, void* stream);

/* ---- mel inversion downstream of the path (csrc/griffin_lim.cuh) ----
 * Replaces audio/audio.py:81-84 (Audio.inv_mel_spectrogram) as called per utterance by
 * audio/utils.py:24-40 (TestUtils.synthesize_and_save_wavs), batched over utterances.  n_fft is fixed at 2048
 * (num_freq 1025, configs/hparams.py:268,386); n_frames[b] = mel length of utterance b (>= 2); the waveform of
 * utterance b has hop_length * (n_frames[b] - 1) samples (librosa.istft, center=True). */
int64_t vaenar_griffin_lim_workspace_bytes(int B, int T, int win_length, int hop_length);
/* audio.py:157-165,180-182,196-206: S = max(1e-10, pinv(mel_basis) @ db_to_amp(denormalize(mel) + ref_level_db)) ** power
 * in the reference's float32 arithmetic.  mel [B,T,n_mels] fp32 (normalised, as the decoder emits it); inv_basis_t
 * [n_mels,num_freq] fp32 = pinv(mel_basis) transposed; S [B,T,num_freq] fp64 (the reference's complex128 magnitudes). */
int vaenar_mel_to_linear(const float* mel, const int32_t* n_frames, const float* inv_basis_t, int B, int T, int n_mels,
                         int num_freq, float min_level_db, float ref_level_db, float max_abs_value, int symmetric,
                         float power, double* S, void* stream);
/* audio.py:93-102 (Audio._griffin_lim) with librosa 0.8.0 stft/istft semantics (audio.py:104-143): `iters` projections
 * after the random-phase start.  rand [B,T,num_freq] fp64 = the np.random.rand draw (NULL: Philox stream from `seed`).
 * wav [B, wav_ld] fp64, zero beyond each utterance's length. */
int vaenar_griffin_lim(const double* S, const int32_t* n_frames, const double* rand, uint64_t seed, int B, int T,
                       int num_freq, int win_length, int hop_length, int iters, void* ws, int64_t ws_bytes, double* wav,
                       int64_t wav_ld, void* stream);
/* audio.py:224-226 (Audio.inv_preemphasize): y[n] = x[n] + k y[n-1] in place (scipy.signal.lfilter([1], [1, -k], x)). */
int vaenar_inv_preemphasis(double* wav, int64_t wav_ld, const int32_t* n_frames, int B, int T, int win_length,
                           int hop_length, double k, void* ws, int64_t ws_bytes, void* stream);
/* audio.py:18-21 (Audio.save_wav up to the file): out = int16(wav * 32767 / max(0.01, max |wav|)) per utterance. */
int vaenar_wav_to_int16(const double* wav, int64_t wav_ld, const int32_t* n_frames, int B, int T, int win_length,
                        int hop_length, void* ws, int64_t ws_bytes, int16_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAENAR_B200_H_ */

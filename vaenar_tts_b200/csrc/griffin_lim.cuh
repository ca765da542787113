// Mel inversion downstream of the synthesis path: mel -> linear magnitudes -> Griffin-Lim
// phase reconstruction -> inverse pre-emphasis -> int16 samples.  Reference: audio/audio.py:81-102 (inv_mel_spectrogram,
// _griffin_lim), :104-151 (_stft/_istft = librosa 0.8.0 stft/istft, center=True, periodic Hann of win_length zero-padded
// to n_fft, reflect padding), :157-165 (_mel_to_linear), :224-226 (inv_preemphasize), :18-21 (save_wav scaling).
//
// One Griffin-Lim iteration  y <- istft(S * exp(i angle(stft(y))))  is ONE kernel launch: a CTA owns two consecutive
// frames, packs them as the real and imaginary part of one complex 2048-point FFT in shared memory (fp64, like the
// reference's complex128 arithmetic), separates the two spectra, re-imposes the magnitudes, transforms back and writes
// the two windowed frames.  The complex spectrogram never exists in HBM; the overlap-add and the division by the window
// sum-of-squares of the previous iterate are evaluated on the fly when the next iteration gathers its frames (ascending
// frame order: the summation order of librosa's overlap-add loop).
#pragma once
#include "ptx.cuh"
#include "simt_kernels.cuh"

namespace vb {

constexpr int GL_LOGN = 11;
constexpr int GL_N = 1 << GL_LOGN;       // n_fft = (num_freq - 1) * 2, audio.py:146
constexpr int GL_BINS = GL_N / 2 + 1;    // num_freq = 1025 (configs/hparams.py:268)
constexpr int GL_THREADS = 256;
constexpr int GL_MAX_WIN = 1024;

struct GLParams {
  const double* S;         // [B][T][s_ld]   target magnitudes (S ** power), frame-major
  const double* rand;      // [B][T][GL_BINS] uniform [0, 1) draws of the first pass (null: Philox from `seed`)
  const double* prev;      // [B][T][win]    windowed frames of the previous iterate (unused on the first pass)
  double* next;            // [B][T][win]
  const int* n_frames;     // [B]
  const double2* tw;       // [GL_N / 2]     exp(-2 pi i k / N)
  const double* window;    // [win]          periodic Hann
  unsigned long long seed;
  int T, s_ld, win, hop, lpad, first;
};

// exp(-2 pi i k / N) and scipy.signal.get_window('hann', win, fftbins=True) in fp64
__global__ void gl_tables_kernel(double2* __restrict__ tw, double* __restrict__ window, int win) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < GL_N / 2) {
    double s, c;
    sincospi(2.0 * i / GL_N, &s, &c);
    tw[i] = make_double2(c, -s);
  }
  if (i < win) window[i] = 0.5 - 0.5 * cospi(2.0 * i / win);
}

// shared-memory index of FFT element / twiddle p (16-byte double2 units; a quarter-warp of 8 lanes is conflict-free when
// its 8 addresses differ mod 8): one pad element per 8, per 64 and per 512, so that EVERY power-of-two stride -- the
// butterflies of all stages, the strided twiddle reads tw[i << s] and the bit-reversed accesses of the spectrum
// separation (stride 256) -- spreads over all banks.  (With the single pad per 8 of the first version 54 % of the kernel's
// shared-memory wavefronts were bank conflicts: ncu, profiles/griffin_lim_r2.md.)
__host__ __device__ constexpr int glp(int p) { return p + (p >> 3) + (p >> 6) + (p >> 9); }
constexpr int GL_ZSIZE = glp(GL_N - 1) + 1;         // 2337
constexpr int GL_TWSIZE = glp(GL_N / 2 - 1) + 1;    // 1167
constexpr int GL_SMEM = static_cast<int>(sizeof(double2)) * (GL_ZSIZE + GL_TWSIZE);
static_assert(2 * GL_TWSIZE >= 2 * GL_MAX_WIN, "the twiddle region also stages the win + hop signal samples of a frame pair");

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// natural order in -> bit-reversed order out (decimation in frequency), radix-4 passes + one radix-2 pass
__device__ __forceinline__ void gl_fft_dif(double2* z, const double2* tw) {
  for (int lh = GL_LOGN - 1; lh >= 1; lh -= 2) {          // h = 1024, 256, 64, 16, 4 (stage h and stage h/2 together)
    const int h = 1 << lh, q = h >> 1;
    for (int g = threadIdx.x; g < GL_N / 4; g += GL_THREADS) {
      // consecutive lanes take consecutive i (consecutive elements); in the last pass (q = 2) consecutive blocks instead
      // (elements 8 apart = 9 padded units apart), which keeps the quarter-warps conflict-free there too
      const int i = q >= 8 ? (g & (q - 1)) : (g >> (GL_LOGN - 1 - lh)), blk = q >= 8 ? (g >> (lh - 1)) : (g & ((GL_N >> (lh + 1)) - 1));
      const int p0 = (blk << (lh + 1)) + i;
      const double2 w1 = tw[glp(i << (GL_LOGN - 1 - lh))], w2 = tw[glp(i << (GL_LOGN - lh))];
      const double2 a0 = z[glp(p0)], a1 = z[glp(p0 + q)], a2 = z[glp(p0 + h)], a3 = z[glp(p0 + h + q)];
      const double2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), w1);
      const double2 b1 = cadd(a1, a3), d13 = csub(a1, a3);
      const double2 b3 = cmul(make_double2(d13.y, -d13.x), w1);   // (a1 - a3) * (-i) * w1
      z[glp(p0)] = cadd(b0, b1);
      z[glp(p0 + q)] = cmul(csub(b0, b1), w2);
      z[glp(p0 + h)] = cadd(b2, b3);
      z[glp(p0 + h + q)] = cmul(csub(b2, b3), w2);
    }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < GL_N / 2; g += GL_THREADS) {     // h = 1: twiddle 1
    const double2 a = z[glp(2 * g)], b = z[glp(2 * g + 1)];
    z[glp(2 * g)] = cadd(a, b);
    z[glp(2 * g + 1)] = csub(a, b);
  }
  __syncthreads();
}

// bit-reversed order in -> natural order out (decimation in time), conjugate twiddles, unscaled
__device__ __forceinline__ void gl_ifft_dit(double2* z, const double2* tw) {
  for (int g = threadIdx.x; g < GL_N / 2; g += GL_THREADS) {     // h = 1
    const double2 a = z[glp(2 * g)], b = z[glp(2 * g + 1)];
    z[glp(2 * g)] = cadd(a, b);
    z[glp(2 * g + 1)] = csub(a, b);
  }
  __syncthreads();
  for (int lh = 1; lh < GL_LOGN; lh += 2) {                      // h = 2, 8, 32, 128, 512 (stage h and stage 2h together)
    const int h = 1 << lh;
    for (int g = threadIdx.x; g < GL_N / 4; g += GL_THREADS) {
      const int i = h >= 8 ? (g & (h - 1)) : (g >> (GL_LOGN - 2 - lh)), blk = h >= 8 ? (g >> lh) : (g & ((GL_N >> (lh + 2)) - 1));
      const int p0 = (blk << (lh + 2)) + i;
      const double2 v1 = tw[glp(i << (GL_LOGN - 1 - lh))], v2 = tw[glp(i << (GL_LOGN - 2 - lh))];
      const double2 a0 = z[glp(p0)], a1 = z[glp(p0 + h)], a2 = z[glp(p0 + 2 * h)], a3 = z[glp(p0 + 3 * h)];
      const double2 t = cmulc(a1, v1), t2 = cmulc(a3, v1);
      const double2 b0 = cadd(a0, t), b1 = csub(a0, t), b2 = cadd(a2, t2), b3 = csub(a2, t2);
      const double2 u = cmulc(b2, v2), u3 = cmulc(b3, v2);
      const double2 u2 = make_double2(-u3.y, u3.x);               // b3 * conj(v2) * (+i)
      z[glp(p0)] = cadd(b0, u);
      z[glp(p0 + 2 * h)] = csub(b0, u);
      z[glp(p0 + h)] = cadd(b1, u2);
      z[glp(p0 + 3 * h)] = csub(b1, u2);
    }
    __syncthreads();
  }
}

// One sample of istft's untrimmed output buffer (librosa.istft + window_sumsquare): overlap-add of the windowed frames in
// ascending frame order -- the order of librosa's overlap-add loop, so the sums round identically -- divided by the
// window sum-of-squares where that exceeds `tiny`.  The sample sits `r` samples after the first windowed sample of
// frame 0; (fq, mr) = divmod(r, hop): frame fq - j holds it at window position mr + j * hop.
__device__ __forceinline__ double gl_ola(const double* __restrict__ frames, const double* w, int nf, int win, int hop,
                                         int fq, int mr) {
  int jmax = 0;
  while (mr + (jmax + 1) * hop < win) ++jmax;
  if (jmax > fq) jmax = fq;
  const int jmin = fq > nf - 1 ? fq - (nf - 1) : 0;
  double acc = 0.0, wss = 0.0;
  for (int j = jmax; j >= jmin; --j) {
    const int m = mr + j * hop;
    const double wm = w[m];
    acc += frames[static_cast<long>(fq - j) * win + m];
    wss += wm * wm;
  }
  return wss > 2.2250738585072014e-308 ? acc / wss : acc;
}

// np.pad(y, n_fft // 2, mode='reflect') coordinate -> index into y (length L >= 1); any pad width
__device__ __forceinline__ int gl_reflect(int t, int L) {
  if (L == 1) return 0;
  const int period = 2 * (L - 1);
  t %= period;
  if (t < 0) t += period;
  return t < L ? t : period - t;
}

// sample t (any integer) of the reflect-padded, trimmed istft output, for frame f / window position m with
// t = f * hop + m + lpad - n_fft / 2 and (mq, mr) = divmod(m, hop)
__device__ __forceinline__ double gl_signal(const double* __restrict__ frames, const double* w, int nf, int win, int hop,
                                            int lpad, int L, int f, int m, int mq, int mr) {
  const int t = f * hop + m + lpad - GL_N / 2;
  if (t >= 0 && t < L) return gl_ola(frames, w, nf, win, hop, f + mq, mr);
  const int r = gl_reflect(t, L) + GL_N / 2 - lpad;
  const int fq = r / hop;
  return gl_ola(frames, w, nf, win, hop, fq, r - fq * hop);
}

__device__ __forceinline__ int gl_brev(int k) { return static_cast<int>(__brev(static_cast<unsigned>(k)) >> (32 - GL_LOGN)); }

__device__ __forceinline__ double gl_uniform(unsigned long long seed, unsigned long long idx) {
  uint32_t c0 = static_cast<uint32_t>(idx), c1 = static_cast<uint32_t>(idx >> 32), c2 = 0x474c494du, c3 = 0;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ((c0 >> 5) * 67108864.0 + (c1 >> 6)) * (1.0 / 9007199254740992.0);   // 53-bit mantissa, numpy's random_sample
}

__device__ __forceinline__ double2 gl_unit_phase(double2 a) {      // exp(1j * angle(a)); angle(0) = 0
  double m = sqrt(fma(a.x, a.x, a.y * a.y));
  if (!(m > 1e-140 && m < 1e140)) m = hypot(a.x, a.y);             // squares under/overflowed (or a == 0)
  if (!(m > 0.0)) return make_double2(1.0, 0.0);
  const double inv = 1.0 / m;                                      // one division instead of two
  return make_double2(a.x * inv, a.y * inv);
}

__global__ void __launch_bounds__(GL_THREADS) gl_iter_kernel(GLParams p) {
  extern __shared__ __align__(16) unsigned char gl_smem[];
  double2* z = reinterpret_cast<double2*>(gl_smem);
  double2* tw = z + GL_ZSIZE;
  const double* __restrict__ w = p.window;                       // 8 KB, L1-resident
  const int b = blockIdx.y, fa = blockIdx.x * 2, fb = fa + 1;
  const int nf = p.n_frames[b];
  if (fa >= nf) return;
  const bool has_b = fb < nf;
  const int tid = threadIdx.x;
  const double* Sa = p.S + (static_cast<long>(b) * p.T + fa) * p.s_ld;
  const double* Sb = Sa + p.s_ld;

  if (p.first) {
    for (int i = tid; i < GL_N / 2; i += GL_THREADS) tw[glp(i)] = p.tw[i];
    // angles = exp(2j * pi * rand); y = istft(S * angles): irfft drops the imaginary part of the DC and Nyquist bins
    const long rbase = (static_cast<long>(b) * p.T + fa) * GL_BINS;
    for (int k = tid; k <= GL_N / 2; k += GL_THREADS) {
      const double ra = p.rand ? p.rand[rbase + k] : gl_uniform(p.seed, rbase + k);
      const double rb = has_b ? (p.rand ? p.rand[rbase + GL_BINS + k] : gl_uniform(p.seed, rbase + GL_BINS + k)) : 0.0;
      double sa, ca, sb, cb;
      sincospi(2.0 * ra, &sa, &ca);
      sincospi(2.0 * rb, &sb, &cb);
      const double ma = Sa[k], mb = has_b ? Sb[k] : 0.0;
      double2 A = make_double2(ma * ca, ma * sa), Bv = make_double2(mb * cb, mb * sb);
      if (k == 0 || k == GL_N / 2) {
        z[glp(gl_brev(k))] = make_double2(A.x, Bv.x);
      } else {
        z[glp(gl_brev(k))] = make_double2(A.x - Bv.y, A.y + Bv.x);
        z[glp(gl_brev(GL_N - k))] = make_double2(A.x + Bv.y, Bv.x - A.y);
      }
    }
    __syncthreads();
  } else {
    // stft frames of y = trimmed, reflect-padded overlap-add of the previous iterate
    const double* frames = p.prev + static_cast<long>(b) * p.T * p.win;
    const int L = p.hop * (nf - 1);
    // The two frames of this CTA overlap by win - hop samples (frame fb at window position m reads the signal sample frame
    // fa reads at m + hop): every distinct sample is evaluated once (win + hop instead of 2 win overlap-add gathers) and
    // staged in the twiddle region, which is filled afterwards.
    double* sig = reinterpret_cast<double*>(tw);
    const int nsig = has_b ? p.win + p.hop : p.win;
    for (int u = tid; u < nsig; u += GL_THREADS) {
      const int mq = u / p.hop, mr = u - mq * p.hop;
      sig[u] = gl_signal(frames, w, nf, p.win, p.hop, p.lpad, L, fa, u, mq, mr);
    }
    __syncthreads();
    for (int n = tid; n < GL_N; n += GL_THREADS) {
      double va = 0.0, vb_ = 0.0;
      const int m = n - p.lpad;
      if (m >= 0 && m < p.win) {
        va = w[m] * sig[m];
        if (has_b) vb_ = w[m] * sig[m + p.hop];
      }
      z[glp(n)] = make_double2(va, vb_);
    }
    __syncthreads();
    for (int i = tid; i < GL_N / 2; i += GL_THREADS) tw[glp(i)] = p.tw[i];
    __syncthreads();
    gl_fft_dif(z, tw);
    // Z = A + iB with A, B the spectra of the two real frames: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i
    for (int k = tid; k <= GL_N / 2; k += GL_THREADS) {
      const int pk = gl_brev(k);
      if (k == 0 || k == GL_N / 2) {
        const double2 Z = z[glp(pk)];
        const double a = Z.x < 0.0 ? -Sa[k] : Sa[k];
        const double bb = has_b ? (Z.y < 0.0 ? -Sb[k] : Sb[k]) : 0.0;
        z[glp(pk)] = make_double2(a, bb);
      } else {
        const int pn = gl_brev(GL_N - k);
        const double2 Zk = z[glp(pk)], Zn = z[glp(pn)];
        const double2 A = make_double2(0.5 * (Zk.x + Zn.x), 0.5 * (Zk.y - Zn.y));
        const double2 Bv = make_double2(0.5 * (Zk.y + Zn.y), 0.5 * (Zn.x - Zk.x));
        double2 ua = gl_unit_phase(A), ub = gl_unit_phase(Bv);
        const double ma = Sa[k], mb = has_b ? Sb[k] : 0.0;
        ua.x *= ma; ua.y *= ma; ub.x *= mb; ub.y *= mb;
        z[glp(pk)] = make_double2(ua.x - ub.y, ua.y + ub.x);
        z[glp(pn)] = make_double2(ua.x + ub.y, ub.x - ua.y);
      }
    }
    __syncthreads();
  }
  gl_ifft_dit(z, tw);
  double* oa = p.next + (static_cast<long>(b) * p.T + fa) * p.win;
  const double inv_n = 1.0 / GL_N;
  for (int m = tid; m < p.win; m += GL_THREADS) {
    const double2 v = z[glp(p.lpad + m)];
    oa[m] = w[m] * (v.x * inv_n);
    if (has_b) oa[p.win + m] = w[m] * (v.y * inv_n);
  }
}

// y = istft(...)[n_fft/2 : -n_fft/2] of the last iterate: wav[b][t], t < hop * (n_frames - 1); zero beyond
__global__ void gl_finalize_kernel(const double* __restrict__ frames, const double* __restrict__ window,
                                   const int* __restrict__ n_frames, int T, int win, int hop, int lpad,
                                   double* __restrict__ wav, long ld) {
  const int b = blockIdx.y;
  const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= ld) return;
  const int nf = n_frames[b];
  const long L = static_cast<long>(hop) * (nf - 1);
  double v = 0.0;
  if (t < L) {
    const int r = static_cast<int>(t) + GL_N / 2 - lpad, fq = r / hop;
    v = gl_ola(frames + static_cast<long>(b) * T * win, window, nf, win, hop, fq, r - fq * hop);
  }
  wav[b * ld + t] = v;
}

// audio.py:81-84,157-165,180-182,196-206 in the reference's float32 arithmetic (the model hands over float32 mels):
// S = max(1e-10, pinv(mel_basis) @ 10^((denormalize(mel) + ref_level_db) / 20)) ** power, stored fp64 (np.complex128)
struct MelToLinearParams {
  const float* mel;        // [B][T][n_mels]
  const float* inv_t;      // [n_mels][GL_BINS]  pinv(mel_basis) transposed
  const int* n_frames;
  double* S;               // [B][T][s_ld]
  int T, n_mels, s_ld, symmetric;
  float min_level_db, ref_level_db, max_abs, power;
};
__global__ void __launch_bounds__(256) mel_to_linear_kernel(MelToLinearParams p) {
  __shared__ float amp[128];
  const int b = blockIdx.y, f = blockIdx.x;
  if (f >= p.n_frames[b]) return;
  const float* mel = p.mel + (static_cast<long>(b) * p.T + f) * p.n_mels;
  for (int j = threadIdx.x; j < p.n_mels; j += blockDim.x) {
    const float s = mel[j];
    float db;
    if (p.symmetric) db = (fminf(fmaxf(s, -p.max_abs), p.max_abs) + p.max_abs) * (-p.min_level_db) / (2.f * p.max_abs) + p.min_level_db;
    else db = fminf(fmaxf(s, 0.f), p.max_abs) * (-p.min_level_db) / p.max_abs + p.min_level_db;
    amp[j] = powf(10.0f, (db + p.ref_level_db) * 0.05f);
  }
  __syncthreads();
  double* out = p.S + (static_cast<long>(b) * p.T + f) * p.s_ld;
  for (int k = threadIdx.x; k < GL_BINS; k += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < p.n_mels; ++j) acc = fmaf(p.inv_t[j * GL_BINS + k], amp[j], acc);
    out[k] = static_cast<double>(powf(fmaxf(1e-10f, acc), p.power));
  }
}

// scipy.signal.lfilter([1], [1, -k], x): y[n] = x[n] + k y[n-1], in place; one CTA per utterance, chunked scan
constexpr int PRE_CHUNK = 512;
__global__ void __launch_bounds__(256) inv_preemphasis_kernel(double* __restrict__ wav, long ld, const int* __restrict__ n_frames,
                                                              int hop, double k, double* __restrict__ carry_ws, int max_chunks) {
  const int b = blockIdx.x;
  const long L = static_cast<long>(hop) * (n_frames[b] - 1);
  double* y = wav + b * ld;
  double* carry = carry_ws + static_cast<long>(b) * max_chunks;
  const int nchunk = static_cast<int>((L + PRE_CHUNK - 1) / PRE_CHUNK);
  for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {       // local recurrences with zero initial state
    const long s = static_cast<long>(c) * PRE_CHUNK, e = s + PRE_CHUNK < L ? s + PRE_CHUNK : L;
    double acc = 0.0;
    for (long n = s; n < e; ++n) {
      acc = y[n] + k * acc;
      y[n] = acc;
    }
    carry[c] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                         // true value at the end of every chunk
    double kc = 1.0;
    for (int i = 0; i < PRE_CHUNK; ++i) kc *= k;
    double prev = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const long len = (static_cast<long>(c) * PRE_CHUNK + PRE_CHUNK < L ? PRE_CHUNK : L - static_cast<long>(c) * PRE_CHUNK);
      double kl = kc;
      if (len != PRE_CHUNK) {
        kl = 1.0;
        for (long i = 0; i < len; ++i) kl *= k;
      }
      const double v = carry[c] + kl * prev;
      carry[c] = prev;                                            // becomes: state entering chunk c
      prev = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
    const double y0 = carry[c];
    if (y0 == 0.0) continue;
    const long s = static_cast<long>(c) * PRE_CHUNK, e = s + PRE_CHUNK < L ? s + PRE_CHUNK : L;
    double pw = k;
    for (long n = s; n < e; ++n) {
      y[n] += pw * y0;
      pw *= k;
    }
  }
}

// save_wav (audio.py:18-21): wav *= 32767 / max(0.01, max |wav|); astype(int16) truncates toward zero
__global__ void __launch_bounds__(256) wav_peak_kernel(const double* __restrict__ wav, long ld, const int* __restrict__ n_frames,
                                                       int hop, double* __restrict__ peak) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  const long L = static_cast<long>(hop) * (n_frames[b] - 1);
  double m = 0.0;
  for (long n = threadIdx.x; n < L; n += blockDim.x) m = fmax(m, fabs(wav[b * ld + n]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) peak[b] = red[0];
}
__global__ void wav_to_int16_kernel(const double* __restrict__ wav, long ld, const int* __restrict__ n_frames, int hop,
                                    const double* __restrict__ peak, short* __restrict__ out) {
  const int b = blockIdx.y;
  const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= ld) return;
  const long L = static_cast<long>(hop) * (n_frames[b] - 1);
  const double scale = 32767.0 / fmax(0.01, peak[b]);
  out[b * ld + t] = t < L ? static_cast<short>(static_cast<int>(wav[b * ld + t] * scale)) : static_cast<short>(0);
}

}  // namespace vb

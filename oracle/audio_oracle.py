"""CPU restatement of the reference's mel-inversion path (TEST INFRASTRUCTURE -- never imported by the product path).

Follows /root/reference/audio/audio.py:81-102 (``inv_mel_spectrogram`` / ``_griffin_lim``), :104-151 (``_stft`` /
``_istft`` / ``_stft_parameters``), :157-174 (``_mel_to_linear`` / ``_build_mel_basis``), :180-182 (``_db_to_amp``),
:196-206 (``_denormalize``), :224-238 (``inv_preemphasize``), :18-21 (``save_wav``) and the caller
/root/reference/audio/utils.py:24-40 (``synthesize_and_save_wavs``).

The arithmetic of ``librosa.stft`` / ``librosa.istft`` / ``librosa.filters.mel`` lives in the third-party dependency
**librosa 0.8.0** (pinned in /root/reference/environment.yml:65; not importable here, no network).  Its published
algorithm is restated below with numpy:
  * stft: periodic Hann window of ``win_length`` zero-padded (centred) to ``n_fft``; ``center=True`` reflect-pads the
    signal by n_fft // 2; frames hop by ``hop_length``; ``rfft`` per frame; result [1 + n_fft/2, n_frames]; the complex
    dtype follows the input (float64 -> complex128);
  * istft: ``irfft`` per frame, multiply by the same padded window, overlap-add, divide by the window sum-of-squares
    where that exceeds ``tiny``; trim n_fft // 2 from both ends;
  * filters.mel: Slaney mel scale (linear below 1 kHz, log above), triangular filters, 'slaney' area normalisation,
    float32.
PARITY UNPINNED against librosa itself (no fixture of the reference exists for this path); the restatement is pinned
against two independent implementations that are designed to reproduce librosa: ``torch.stft`` / ``torch.istft`` and
``transformers.audio_utils.mel_filter_bank``, and against scipy (a dependency of the reference's audio.py and of librosa:
``get_window('hann', fftbins=True)``, ``scipy.signal.stft`` framing, ``lfilter``) -- tests/test_audio_cpu.py.
"""
import numpy as np


class LJAudio:
    """configs/hparams.py:266-282 (LJHPS.Audio)"""
    num_mels = 80
    num_freq = 1025
    min_mel_freq = 0.
    max_mel_freq = 8000.
    sample_rate = 22050
    frame_length_sample = 1024
    frame_shift_sample = 256
    preemphasize = 0.97
    min_level_db = -100.0
    ref_level_db = 20.0
    max_abs_value = 1
    symmetric_specs = False
    griffin_lim_iters = 60
    power = 1.5
    center = True


class DataBakerAudio(LJAudio):
    """configs/hparams.py:384-400 (DataBakerHPS.Audio)"""
    sample_rate = 16000
    frame_length_sample = 800
    frame_shift_sample = 200
    min_level_db = -115.


# ---------------------------------------------------------------- librosa 0.8.0 restatement
def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True)"""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pad_center(w, size):
    lpad = (size - len(w)) // 2
    return np.pad(w, (lpad, size - len(w) - lpad), mode="constant")


def stft(y, n_fft, hop_length, win_length):
    """librosa.stft(center=True, window='hann', pad_mode='reflect'): [1 + n_fft/2, n_frames]"""
    w = pad_center(hann_periodic(win_length), n_fft)
    y = np.pad(np.asarray(y), n_fft // 2, mode="reflect")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return np.fft.rfft(w[:, None] * y[idx], axis=0)


def window_sumsquare(n_frames, hop_length, win_length, n_fft):
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n)
    wsq = pad_center(hann_periodic(win_length) ** 2, n_fft)
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += wsq[:max(0, min(n_fft, n - s))]
    return x


def istft(D, hop_length, win_length):
    """librosa.istft(center=True, window='hann', length=None)"""
    n_fft = 2 * (D.shape[0] - 1)
    n_frames = D.shape[1]
    w = pad_center(hann_periodic(win_length), n_fft)
    y = np.zeros(n_fft + hop_length * (n_frames - 1))
    ytmp = w[:, None] * np.fft.irfft(D, n=n_fft, axis=0)
    for f in range(n_frames):
        y[f * hop_length:f * hop_length + n_fft] += ytmp[:, f]
    wss = window_sumsquare(n_frames, hop_length, win_length, n_fft)
    nz = wss > np.finfo(np.float64).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney', dtype=float32)"""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


# ---------------------------------------------------------------- audio/audio.py
class Audio:
    def __init__(self, hps):
        self.hps = hps

    def _stft_parameters(self):                                     # audio.py:145-151
        return (self.hps.num_freq - 1) * 2, self.hps.frame_shift_sample, self.hps.frame_length_sample

    def _build_mel_basis(self):                                     # audio.py:167-174
        n_fft = (self.hps.num_freq - 1) * 2
        return mel_filters(self.hps.sample_rate, n_fft, self.hps.num_mels, self.hps.min_mel_freq,
                           self.hps.max_mel_freq)

    def _mel_to_linear(self, mel_spectrogram):                      # audio.py:157-165
        inv = np.linalg.pinv(self._build_mel_basis())
        return np.maximum(1e-10, np.dot(inv, mel_spectrogram))

    @staticmethod
    def _db_to_amp(x):                                              # audio.py:180-182
        return np.power(10.0, x * 0.05)

    def _denormalize(self, S):                                      # audio.py:196-206
        h = self.hps
        if h.symmetric_specs:
            return ((np.clip(S, -h.max_abs_value, h.max_abs_value) + h.max_abs_value) * (-h.min_level_db)
                    / (2 * h.max_abs_value) + h.min_level_db)
        return (np.clip(S, 0, h.max_abs_value) * (-h.min_level_db) / h.max_abs_value) + h.min_level_db

    def _stft(self, y):                                             # audio.py:104-109
        n_fft, hop, win = self._stft_parameters()
        return stft(y, n_fft, hop, win)

    def _istft(self, D):                                            # audio.py:127-132
        _, hop, win = self._stft_parameters()
        return istft(D, hop, win)

    def _griffin_lim(self, S, rand=None, iters=None, trace=None):   # audio.py:93-102
        """``rand`` = the ``np.random.rand(*S.shape)`` draw of the reference, injected by the harness."""
        if rand is None:
            rand = np.random.rand(*S.shape)
        angles = np.exp(2j * np.pi * rand)
        S_complex = np.abs(S).astype(np.complex128)
        y = self._istft(S_complex * angles)
        for i in range(self.hps.griffin_lim_iters if iters is None else iters):
            if trace is not None:
                trace.append(y.copy())
            angles = np.exp(1j * np.angle(self._stft(y)))
            y = self._istft(S_complex * angles)
        return y

    def linear_magnitudes(self, mel_spectrogram):
        """The argument of ``_griffin_lim`` in ``inv_mel_spectrogram`` (audio.py:81-84): [num_freq, T]"""
        S = self._mel_to_linear(self._db_to_amp(self._denormalize(mel_spectrogram) + self.hps.ref_level_db))
        return S ** self.hps.power

    def inv_mel_spectrogram(self, mel_spectrogram, rand=None, iters=None):   # audio.py:81-84; input [num_mels, T]
        return self._griffin_lim(self.linear_magnitudes(mel_spectrogram), rand=rand, iters=iters)

    def inv_preemphasize(self, x):                                  # audio.py:224-226: lfilter([1], [1, -k], x)
        k = self.hps.preemphasize
        if k is None:
            return x
        y = np.empty_like(x)
        acc = 0.0
        for n in range(len(x)):
            acc = x[n] + k * acc
            y[n] = acc
        return y

    @staticmethod
    def to_int16(wav):                                              # audio.py:18-21 (save_wav without the file)
        wav = wav * (32767 / max(0.01, np.max(np.abs(wav))))
        return wav.astype(np.int16)


def synthesize(audio, mel, rand=None, iters=None):
    """``_synthesize`` of audio/utils.py:25-29 up to the int16 samples: mel [T, num_mels] (already cropped)"""
    wav = audio.inv_mel_spectrogram(mel.T, rand=rand, iters=iters)
    wav = audio.inv_preemphasize(wav)
    return audio.to_int16(wav), wav

"""Mel inversion on the B200: the synthesis-side half of the reference's ``audio.Audio``
(audio/audio.py:11-21, :81-102, :145-174, :196-226) and ``audio.utils.TestUtils`` (audio/utils.py:10-40) with the same
method names, batched over utterances.  All signal arithmetic -- mel -> linear magnitudes, the Griffin-Lim iterations
(fp64 2048-point FFTs, one kernel launch per iteration), inverse pre-emphasis, int16 scaling -- runs in
csrc/griffin_lim.cuh through the C ABI (include/vaenar_b200.h); there is no CPU fallback.  The host only builds the
constant mel filter bank / its pseudo-inverse once per ``Audio`` object (librosa.filters.mel + np.linalg.pinv in the
reference, audio.py:157-174) and writes files."""
import os
import wave

import numpy as np
import torch

from . import _lib
from ._lib import check, VaenarError


def _slaney_hz(mel):
    """Slaney mel scale -> Hz (librosa.mel_to_hz, htk=False): linear below 1 kHz, logarithmic above."""
    mel = np.asarray(mel, dtype=np.float64)
    lin = mel * (200.0 / 3)
    knee = 1000.0 / (200.0 / 3)
    return np.where(mel >= knee, 1000.0 * np.exp((mel - knee) * (np.log(6.4) / 27.0)), lin)


def _slaney_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    knee = 1000.0 / (200.0 / 3)
    return np.where(hz >= 1000.0, knee + np.log(np.maximum(hz, 1e-300) / 1000.0) / (np.log(6.4) / 27.0), hz / (200.0 / 3))


def mel_filter_bank(sample_rate, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (0.8.0 defaults: Slaney scale, 'slaney' norm, float32)
    as called by audio.py:167-174: [n_mels, 1 + n_fft // 2]."""
    bins = np.linspace(0.0, sample_rate / 2.0, 1 + n_fft // 2)
    edges = _slaney_hz(np.linspace(_slaney_mel(fmin), _slaney_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    dist = edges[:, None] - bins[None, :]
    fb = np.zeros((n_mels, bins.size), dtype=np.float32)
    for i in range(n_mels):
        fb[i] = np.maximum(0.0, np.minimum(-dist[i] / width[i], dist[i + 2] / width[i + 1]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


class Audio:
    """Drop-in for the methods of ``audio.Audio`` that sit downstream of the synthesis path."""

    def __init__(self, audio_hparams, device="cuda:0"):
        self.hps = audio_hparams
        self.device = torch.device(device)
        self._lib = None
        self._inv_t = None
        self._ws = None

    # ------------------------------------------------------------------ constants (host, once)
    def _stft_parameters(self):                                   # audio.py:145-151
        return (self.hps.num_freq - 1) * 2, self.hps.frame_shift_sample, self.hps.frame_length_sample

    def _build_mel_basis(self):                                   # audio.py:167-174
        n_fft = (self.hps.num_freq - 1) * 2
        return mel_filter_bank(self.hps.sample_rate, n_fft, self.hps.num_mels, self.hps.min_mel_freq,
                               self.hps.max_mel_freq)

    def _inv_basis_t(self):
        if self._inv_t is None:
            inv = np.linalg.pinv(self._build_mel_basis())        # audio.py:158, float32 like the reference
            self._inv_t = torch.from_numpy(np.ascontiguousarray(inv.T.astype(np.float32))).to(self.device)
        return self._inv_t

    # ------------------------------------------------------------------ plumbing
    def _L(self):
        if self._lib is None:
            if self.device.type != "cuda" or not torch.cuda.is_available():
                raise VaenarError("Audio needs a CUDA device: the B200 path has no CPU fallback")
            self._lib = _lib.load()
        return self._lib

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, B, T):
        _, hop, win = self._stft_parameters()
        need = int(self._L().vaenar_griffin_lim_workspace_bytes(B, T, win, hop))
        if need < 0:
            raise VaenarError(f"griffin_lim workspace: bad shape B {B} T {T}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _frames(self, n_frames, B, T):
        n = torch.as_tensor(n_frames).to(torch.int64).reshape(-1)
        if n.numel() != B:
            raise VaenarError(f"mel_lengths has {n.numel()} entries for a batch of {B}")
        if int(n.min()) < 2 or int(n.max()) > T:
            raise VaenarError(f"mel_lengths must lie in [2, {T}] (librosa's reflect padding needs a non-empty signal)")
        return n.to(device=self.device, dtype=torch.int32)

    # ------------------------------------------------------------------ batched device entry points
    def linear_magnitudes(self, mel_batch, mel_lengths):
        """``_mel_to_linear(_db_to_amp(_denormalize(mel) + ref_level_db)) ** power`` (audio.py:81-84) for a batch
        [B, T, num_mels] fp32 -> S [B, T, num_freq] fp64, frame-major."""
        h = self.hps
        self._L()
        mel = torch.as_tensor(mel_batch).to(device=self.device, dtype=torch.float32).contiguous()
        B, T, M = mel.shape
        n = self._frames(mel_lengths, B, T)
        S = torch.zeros(B, T, h.num_freq, dtype=torch.float64, device=self.device)
        check(self._L().vaenar_mel_to_linear(self._p(mel), self._p(n), self._p(self._inv_basis_t()), B, T, M, h.num_freq,
                                             float(h.min_level_db), float(h.ref_level_db), float(h.max_abs_value),
                                             int(bool(h.symmetric_specs)), float(h.power), self._p(S), self._stream()))
        return S

    def griffin_lim(self, S, n_frames, rand=None, seed=0, iters=None):
        """``_griffin_lim`` (audio.py:93-102) on S [B, T, num_freq] fp64 (frame-major).  ``rand`` [B, T, num_freq] = the
        ``np.random.rand`` draw of the first pass (default: counter-based generator seeded by ``seed``).
        Returns wav [B, hop * (T - 1)] fp64 on the device, zero beyond hop * (n_frames[b] - 1)."""
        h = self.hps
        self._L()
        _, hop, win = self._stft_parameters()
        S = torch.as_tensor(S).to(device=self.device, dtype=torch.float64).contiguous()
        B, T, F = S.shape
        n = self._frames(n_frames, B, T)
        if rand is not None:
            rand = torch.as_tensor(rand).to(device=self.device, dtype=torch.float64).contiguous()
            if tuple(rand.shape) != (B, T, F):
                raise VaenarError(f"rand shape {tuple(rand.shape)} != {(B, T, F)}")
        iters = int(h.griffin_lim_iters if iters is None else iters)
        ws = self._workspace(B, T)
        ld = hop * (T - 1)
        wav = torch.empty(B, ld, dtype=torch.float64, device=self.device)
        check(self._L().vaenar_griffin_lim(self._p(S), self._p(n), self._p(rand), int(seed) & (2 ** 64 - 1), B, T, F, win, hop,
                                           iters, self._p(ws), ws.numel(), self._p(wav), ld, self._stream()))
        return wav

    def inv_preemphasize_batch(self, wav, n_frames):
        """``inv_preemphasize`` (audio.py:224-226) in place on wav [B, ld] fp64 (device)."""
        k = self.hps.preemphasize
        if k is None:
            return wav
        self._L()
        _, hop, win = self._stft_parameters()
        B, ld = wav.shape
        T = ld // hop + 1
        n = self._frames(n_frames, B, T)
        ws = self._workspace(B, T)
        check(self._L().vaenar_inv_preemphasis(self._p(wav), ld, self._p(n), B, T, win, hop, float(k), self._p(ws),
                                               ws.numel(), self._stream()))
        return wav

    def to_int16_batch(self, wav, n_frames):
        """The scaling of ``save_wav`` (audio.py:18-21) per utterance: int16 [B, ld] on the device."""
        self._L()
        _, hop, win = self._stft_parameters()
        B, ld = wav.shape
        T = ld // hop + 1
        n = self._frames(n_frames, B, T)
        ws = self._workspace(B, T)
        out = torch.empty(B, ld, dtype=torch.int16, device=self.device)
        check(self._L().vaenar_wav_to_int16(self._p(wav), ld, self._p(n), B, T, win, hop, self._p(ws), ws.numel(),
                                            self._p(out), self._stream()))
        return out

    def inv_mel_spectrogram_batch(self, mel_batch, mel_lengths, rand=None, seed=0, iters=None):
        """``inv_mel_spectrogram`` (audio.py:81-84) of every utterance of mel_batch [B, T, num_mels]: wav [B, hop*(T-1)]
        fp64 on the device; utterance b owns the first hop * (mel_lengths[b] - 1) samples."""
        S = self.linear_magnitudes(mel_batch, mel_lengths)
        return self.griffin_lim(S, mel_lengths, rand=rand, seed=seed, iters=iters)

    # ------------------------------------------------------------------ the reference's per-utterance signatures
    def inv_mel_spectrogram(self, mel_spectrogram, rand=None, seed=0):
        """audio.py:81-84: mel_spectrogram [num_mels, T] (the reference passes ``mel.T``) -> wav [hop * (T - 1)]
        (numpy float64).  ``rand`` [num_freq, T] optionally injects the ``np.random.rand`` draw."""
        mel = torch.as_tensor(np.ascontiguousarray(np.asarray(mel_spectrogram, dtype=np.float32).T))[None]
        T = mel.shape[1]
        r = None if rand is None else np.ascontiguousarray(np.asarray(rand, dtype=np.float64).T)[None]
        return self.inv_mel_spectrogram_batch(mel, [T], rand=r, seed=seed)[0].cpu().numpy()

    def inv_preemphasize(self, x):
        """audio.py:224-238 for [time] or [1, time] arrays."""
        x = np.asarray(x, dtype=np.float64)
        if self.hps.preemphasize is None:
            return x
        self._L()
        _, hop, _ = self._stft_parameters()
        flat = x.reshape(-1)
        T = -(-flat.size // hop) + 1                       # smallest frame count whose signal covers x
        buf = torch.zeros(1, hop * (T - 1), dtype=torch.float64, device=self.device)
        buf[0, :flat.size] = torch.from_numpy(flat)
        return self.inv_preemphasize_batch(buf, [T])[0, :flat.size].cpu().numpy().reshape(x.shape)

    def save_wav(self, wav, path):
        """audio.py:18-21: scale to the int16 range by the peak (floor 0.01) and write a PCM file."""
        wav = np.asarray(wav, dtype=np.float64).reshape(-1)
        self._L()
        _, hop, _ = self._stft_parameters()
        T = -(-wav.size // hop) + 1
        buf = torch.zeros(1, hop * (T - 1), dtype=torch.float64, device=self.device)
        buf[0, :wav.size] = torch.from_numpy(wav)
        pcm = self.to_int16_batch(buf, [T])[0, :wav.size].cpu().numpy()
        write_pcm16(path, self.hps.sample_rate, pcm)


def write_pcm16(path, sample_rate, pcm):
    """scipy.io.wavfile.write(path, sr, int16 array) of audio.py:20: mono 16-bit RIFF/WAVE."""
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


class TestUtils:
    """``audio.utils.TestUtils`` (audio/utils.py:10-40): the writers ``inference.py`` / ``train.py`` call after a
    synthesis step.  ``synthesize_and_save_wavs`` inverts the whole batch in one pass on the GPU instead of one CPU
    thread per utterance."""
    __test__ = False      # not a pytest class

    def __init__(self, hps, save_dir, device="cuda:0"):
        self.prcocessor = Audio(hps.Audio, device=device)      # (sic) attribute name of audio/utils.py:12
        self.hps = hps
        self.save_dir = save_dir

    def write_mels(self, step, mel_batch, mel_lengths, ids, prefix=""):
        from .synthesis import write_mels
        return write_mels(self.save_dir, step, mel_batch, mel_lengths, ids, prefix=prefix)

    def synthesize_and_save_wavs(self, step, mel_batch, mel_lengths, ids, prefix="", rand=None, seed=0):
        """audio/utils.py:24-40: per utterance  inv_mel_spectrogram(mel[:len].T) -> inv_preemphasize -> save_wav to
        ``{prefix}-{id}-{step}.wav``.  Returns the file names."""
        os.makedirs(self.save_dir, exist_ok=True)
        a = self.prcocessor
        lens = torch.as_tensor(mel_lengths).to(torch.int64).reshape(-1).cpu()
        _, hop, _ = a._stft_parameters()
        wav = a.inv_mel_spectrogram_batch(mel_batch, lens, rand=rand, seed=seed)
        wav = a.inv_preemphasize_batch(wav, lens)
        pcm = a.to_int16_batch(wav, lens).cpu().numpy()
        names = []
        for i in range(pcm.shape[0]):
            idx = ids[i].decode("utf-8") if isinstance(ids[i], bytes) else ids[i]
            name = os.path.join(self.save_dir, "{}-{}-{}.wav".format(prefix, idx, step))
            write_pcm16(name, a.hps.sample_rate, pcm[i, :hop * (int(lens[i]) - 1)])
            names.append(name)
        return names

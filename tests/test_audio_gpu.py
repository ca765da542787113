"""Parity of the CUDA mel-inversion path (csrc/griffin_lim.cuh through the C ABI, vaenar_tts_b200.audio) against the
oracle restatement of audio/audio.py:81-102 / audio/utils.py:24-40 (oracle/audio_oracle.py).

Tolerances.  The Griffin-Lim kernels compute in fp64 like the reference's complex128 arithmetic, so with the SAME
magnitudes and random phases the waveform must agree to 1e-9 of its peak after any number of iterations (measured:
~1e-13; FFT factorisation and x/|x| vs exp(1j*angle(x)) differ in the last bits only).  The mel -> linear step follows
the reference's float32 flow (sgemm summation order is unspecified): 2e-5 relative on S, 1e-4 of the peak on the
end-to-end waveform, +-2 LSB on a few int16 samples."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import audio_oracle as A  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "audio_griffin_lim.npz")


def _audio(which):
    from vaenar_tts_b200 import LJHPS, DataBakerHPS
    from vaenar_tts_b200.audio import Audio
    return (Audio(LJHPS.Audio), A.Audio(A.LJAudio)) if which == "lj" else (Audio(DataBakerHPS.Audio), A.Audio(A.DataBakerAudio))


def _mel(rng, T):
    from golden_util import speechlike_mel
    return speechlike_mel(rng, T)


@pytest.mark.parametrize("which", ["lj", "db"])
def test_golden_vectors(which):
    g = np.load(GOLD)
    dev, _ = _audio(which)
    S, rand, iters = g[f"{which}_S"], g[f"{which}_rand"], int(g[f"{which}_iters"])
    T = S.shape[1]
    wav = dev.griffin_lim(torch.from_numpy(S.T.astype(np.float64))[None], [T], rand=rand.T[None], iters=iters)
    ref = g[f"{which}_wav"]
    assert wav.shape == (1, ref.size)
    assert np.abs(wav[0].cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
    pre = dev.inv_preemphasize_batch(wav.clone(), [T])
    assert np.abs(pre[0].cpu().numpy() - g[f"{which}_pre"]).max() <= 1e-9 * np.abs(g[f"{which}_pre"]).max()
    pcm = dev.to_int16_batch(pre, [T])[0].cpu().numpy()
    assert np.abs(pcm.astype(int) - g[f"{which}_pcm"].astype(int)).max() <= 1


@pytest.mark.parametrize("which,lens,iters", [
    ("lj", [37, 12, 2, 36], 60),        # ragged batch, odd and even frame counts, the 2-frame minimum
    ("db", [21, 20, 5], 60),            # win 800 / hop 200 (window not a power of two)
    ("lj", [9], 0),                     # random-phase start only (audio.py:94-96)
    ("lj", [8, 3], 1),
])
def test_griffin_lim_vs_oracle_same_magnitudes(which, lens, iters):
    dev, orc = _audio(which)
    rng = np.random.default_rng(sum(lens) + iters)
    B, T = len(lens), max(lens)
    hop = orc.hps.frame_shift_sample
    S = np.zeros((B, T, 1025))
    rand = rng.random((B, T, 1025))
    for b, n in enumerate(lens):
        S[b, :n] = orc.linear_magnitudes(_mel(rng, n).T).T.astype(np.float64)
    wav = dev.griffin_lim(torch.from_numpy(S), lens, rand=rand, iters=iters).cpu().numpy()
    assert wav.shape == (B, hop * (T - 1))
    for b, n in enumerate(lens):
        ref = orc._griffin_lim(S[b, :n].T, rand=rand[b, :n].T, iters=iters)
        got = wav[b, :ref.size]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print(f"{which} utt {b} frames {n} iters {iters}: rel err {err:.2e}")
        assert err <= 1e-9, (b, n, err)
        assert not wav[b, ref.size:].any()                       # zero beyond the utterance


def test_mel_to_linear_float32_flow():
    dev, orc = _audio("lj")
    rng = np.random.default_rng(11)
    lens = [30, 17]
    mel = np.zeros((2, 30, 80), np.float32)
    for b, n in enumerate(lens):
        mel[b, :n] = _mel(rng, n)
    mel[0, 3, :5] = [-0.2, 0.0, 1.0, 1.7, 0.5]                    # clipping of _denormalize (audio.py:203-206)
    S = dev.linear_magnitudes(mel, lens).cpu().numpy()
    for b, n in enumerate(lens):
        ref = orc.linear_magnitudes(mel[b, :n].T).T.astype(np.float64)
        got = S[b, :n]
        floor = ref <= 1.1e-15                                    # the 1e-10 floor of audio.py:165, ** 1.5
        assert (got[floor] <= 1.1e-15).mean() > 0.999             # entries at the floor stay at the floor
        rel = np.abs(got - ref)[~floor] / np.maximum(ref[~floor], 1e-3 * ref.max())
        print("S rel err", rel.max())
        assert rel.max() < 2e-5
        assert not S[b, n:].any()


@pytest.mark.parametrize("which", ["lj", "db"])
def test_synthesize_and_save_wavs_end_to_end(tmp_path, which):
    """audio/utils.py:24-40 on a ragged batch: files named like the reference's, samples against the oracle chain."""
    from vaenar_tts_b200 import LJHPS, DataBakerHPS
    from vaenar_tts_b200.audio import TestUtils
    hps = LJHPS if which == "lj" else DataBakerHPS
    _, orc = _audio(which)
    rng = np.random.default_rng(21)
    lens = [26, 40, 9]
    B, T = len(lens), max(lens)
    mel = np.zeros((B, T, 80), np.float32)
    for b, n in enumerate(lens):
        mel[b, :n] = _mel(rng, n)
    mel[1, 30:] = rng.random((10, 80)).astype(np.float32)         # padding garbage beyond mel_lengths must be ignored ...
    lens[1] = 30
    rand = rng.random((B, T, 1025))
    tester = TestUtils(hps, str(tmp_path))
    ids = [b"LJ001-0001", "LJ001-0002", "x"]
    names = tester.synthesize_and_save_wavs(1234, torch.from_numpy(mel).cuda(), torch.tensor(lens), ids, prefix="prior",
                                            rand=rand)
    assert [os.path.basename(n) for n in names] == ["prior-LJ001-0001-1234.wav", "prior-LJ001-0002-1234.wav",
                                                    "prior-x-1234.wav"]
    hop = orc.hps.frame_shift_sample
    for b, n in enumerate(lens):
        with wave.open(names[b], "rb") as f:
            assert (f.getnchannels(), f.getsampwidth(), f.getframerate()) == (1, 2, orc.hps.sample_rate)
            pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
        ref_pcm, ref_wav = A.synthesize(orc, mel[b, :n], rand=rand[b, :n].T)
        assert pcm.size == ref_pcm.size == hop * (n - 1)
        d = np.abs(pcm.astype(int) - ref_pcm.astype(int))
        print(f"{which} utt {b}: int16 max diff {d.max()}, differing samples {(d > 0).mean():.3%}")
        assert d.max() <= 2 and (d > 1).mean() < 1e-3


def test_reference_signatures_single_utterance(tmp_path):
    """Audio.inv_mel_spectrogram(mel.T) / inv_preemphasize(wav) / save_wav(wav, path) as audio/utils.py:25-29 calls them."""
    dev, orc = _audio("lj")
    rng = np.random.default_rng(4)
    mel = _mel(rng, 19)
    rand = rng.random((1025, 19))
    wav = dev.inv_mel_spectrogram(mel.T, rand=rand)
    ref = orc.inv_mel_spectrogram(mel.T, rand=rand)
    assert wav.shape == ref.shape and wav.dtype == np.float64
    assert np.abs(wav - ref).max() <= 1e-4 * np.abs(ref).max()
    pre = dev.inv_preemphasize(wav)
    assert np.abs(pre - orc.inv_preemphasize(wav)).max() <= 1e-10 * np.abs(pre).max()
    path = str(tmp_path / "a.wav")
    dev.save_wav(pre, path)
    with wave.open(path, "rb") as f:
        pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
    assert np.abs(pcm.astype(int) - A.Audio.to_int16(pre).astype(int)).max() <= 1
    # unseeded draws are reproducible per seed and differ between seeds
    w0, w1, w2 = (dev.inv_mel_spectrogram(mel.T, seed=s) for s in (7, 7, 8))
    assert np.array_equal(w0, w1) and not np.array_equal(w0, w2)


def test_full_size_batch_properties():
    """BASELINE config 2's mel shape (B16 x 870 frames, 60 iterations) is far beyond what the numpy oracle finishes in
    seconds, so check size-independent properties: (i) utterances are independent -- a 3-utterance slice of the batch
    gives bit-identical samples; (ii) spectral inconsistency || |stft(y)| - S || / || S || after 60 iterations is
    below the random-phase start's (Griffin-Lim's monotone-descent property; the synthetic mels are noisy, so the
    fixed point stays far from consistent); (iii) one utterance against the oracle."""
    dev, orc = _audio("lj")
    rng = np.random.default_rng(99)
    B, T = 16, 870
    lens = [T] + [int(x) for x in rng.integers(300, T, B - 1)]
    mel = np.zeros((B, T, 80), np.float32)
    for b, n in enumerate(lens):
        mel[b, :n] = _mel(rng, n)
    S = dev.linear_magnitudes(mel, lens)
    rand = torch.rand(B, T, 1025, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    w0 = dev.griffin_lim(S, lens, rand=rand, iters=0)
    w60 = dev.griffin_lim(S, lens, rand=rand, iters=60)
    sub = dev.griffin_lim(S[5:8].contiguous(), lens[5:8], rand=rand[5:8].contiguous(), iters=60)
    hop = 256
    for i, b in enumerate(range(5, 8)):
        L = hop * (lens[b] - 1)
        assert torch.equal(sub[i, :L], w60[b, :L])
    window = torch.hann_window(1024, periodic=True, dtype=torch.float64, device="cuda")

    def inconsistency(w, b):
        L = hop * (lens[b] - 1)
        D = torch.stft(w[b, :L], 2048, hop, 1024, window=window, center=True, pad_mode="reflect", return_complex=True)
        Sb = S[b, :lens[b]].T
        return float((D.abs() - Sb).norm() / Sb.norm())
    for b in (0, 3, 15):
        e0, e60 = inconsistency(w0, b), inconsistency(w60, b)
        print(f"utt {b} ({lens[b]} frames): inconsistency {e0:.3f} -> {e60:.3f}")
        assert e60 < 0.8 * e0
    b = int(np.argmin(lens))
    n = lens[b]
    ref = orc._griffin_lim(S[b, :n].T.cpu().numpy(), rand=rand[b, :n].T.cpu().numpy(), iters=60)
    got = w60[b, :ref.size].cpu().numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"utt {b} ({n} frames, 60 iterations) vs oracle: rel err {err:.2e}")
    assert err <= 1e-9

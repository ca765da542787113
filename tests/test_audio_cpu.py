"""CPU checks of the mel-inversion oracle (oracle/audio_oracle.py: restatement of audio/audio.py:81-102 with librosa
0.8.0's stft / istft / filters.mel) and of the host side of vaenar_tts_b200.audio.

librosa is not installable here, so the restatement is pinned against two independent implementations designed to
reproduce it -- torch.stft / torch.istft (center=True, reflect padding, periodic Hann zero-padded to n_fft) and
transformers.audio_utils.mel_filter_bank (Slaney scale + 'slaney' norm) -- plus the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import audio_oracle as A

GOLD = os.path.join(os.path.dirname(__file__), "golden", "audio_griffin_lim.npz")


@pytest.mark.parametrize("hps", [A.LJAudio, A.DataBakerAudio])
def test_stft_istft_match_torch(hps):
    audio = A.Audio(hps)
    n_fft, hop, win = audio._stft_parameters()
    rng = np.random.default_rng(3)
    y = rng.standard_normal(hop * 23)
    window = torch.hann_window(win, periodic=True, dtype=torch.float64)
    D = audio._stft(y)
    Dt = torch.stft(torch.from_numpy(y), n_fft, hop, win, window=window, center=True, pad_mode="reflect",
                    return_complex=True).numpy()
    assert D.shape == (1025, 24)
    assert np.abs(D - Dt).max() < 1e-11
    yi = audio._istft(D)
    yt = torch.istft(torch.from_numpy(D), n_fft, hop, win, window=window, center=True).numpy()
    assert yi.shape == yt.shape == y.shape
    assert np.abs(yi - yt).max() < 1e-12
    assert np.abs(yi - y).max() < 1e-12          # NOLA holds for Hann at 75 % overlap: perfect reconstruction


@pytest.mark.parametrize("hps", [A.LJAudio, A.DataBakerAudio])
def test_mel_basis_matches_transformers_and_product(hps):
    from vaenar_tts_b200.audio import mel_filter_bank
    basis = A.Audio(hps)._build_mel_basis()
    assert basis.shape == (80, 1025) and basis.dtype == np.float32
    mine = mel_filter_bank(hps.sample_rate, 2048, hps.num_mels, hps.min_mel_freq, hps.max_mel_freq)
    assert np.array_equal(mine, basis)
    tfb = pytest.importorskip("transformers.audio_utils").mel_filter_bank(
        1025, hps.num_mels, hps.min_mel_freq, hps.max_mel_freq, hps.sample_rate, norm="slaney", mel_scale="slaney").T
    assert np.abs(basis - tfb).max() < 1e-8
    # every filter is a non-negative triangle with 'slaney' area normalisation: sum * bin width == 1 for interior filters
    assert (basis >= 0).all()
    area = basis.sum(1) * (hps.sample_rate / 2048)
    assert np.abs(area[1:-1] - 1).max() < 0.03


def test_product_hparams_mirror_reference_audio_block():
    from vaenar_tts_b200 import LJHPS, DataBakerHPS
    for prod, orc in ((LJHPS.Audio, A.LJAudio), (DataBakerHPS.Audio, A.DataBakerAudio)):
        for k in ("num_mels", "num_freq", "min_mel_freq", "max_mel_freq", "sample_rate", "frame_length_sample",
                  "frame_shift_sample", "preemphasize", "min_level_db", "ref_level_db", "max_abs_value", "symmetric_specs",
                  "griffin_lim_iters", "power", "center"):
            assert getattr(prod, k) == getattr(orc, k), k


def test_oracle_reproduces_golden_vectors():
    g = np.load(GOLD)
    for name, hps in (("lj", A.LJAudio), ("db", A.DataBakerAudio)):
        audio = A.Audio(hps)
        mel, rand, iters = g[f"{name}_mel"], g[f"{name}_rand"], int(g[f"{name}_iters"])
        S = audio.linear_magnitudes(mel.T)
        assert S.dtype == np.float32                              # the reference's float32 flow (audio.py:157-165)
        assert np.allclose(S, g[f"{name}_S"], rtol=2e-5, atol=0)  # sgemm summation order may differ between BLAS builds
        wav = audio._griffin_lim(g[f"{name}_S"], rand=rand, iters=iters)
        assert wav.shape == ((mel.shape[0] - 1) * hps.frame_shift_sample,)
        assert np.abs(wav - g[f"{name}_wav"]).max() <= 1e-10 * np.abs(wav).max()
        pre = audio.inv_preemphasize(wav)
        assert np.abs(pre - g[f"{name}_pre"]).max() <= 1e-10 * np.abs(pre).max()
        assert np.abs(A.Audio.to_int16(pre).astype(int) - g[f"{name}_pcm"].astype(int)).max() <= 1


def test_inverse_preemphasis_is_the_lfilter_recurrence():
    signal = pytest.importorskip("scipy.signal")
    x = np.random.default_rng(0).standard_normal(5000)
    ref = signal.lfilter([1], [1, -0.97], x)
    assert np.abs(A.Audio(A.LJAudio).inv_preemphasize(x) - ref).max() < 1e-12


def test_window_and_stft_against_scipy():
    """scipy is a direct dependency of the reference's audio.py and of librosa 0.8.0: the analysis window is
    scipy.signal.get_window('hann', win, fftbins=True) (librosa.filters.get_window), and scipy.signal.stft with the same
    window / hop / reflect boundary yields librosa's frames up to its documented scaling by the window sum."""
    signal = pytest.importorskip("scipy.signal")
    for win in (1024, 800, 64):
        assert np.abs(A.hann_periodic(win) - signal.get_window("hann", win, fftbins=True)).max() < 1e-15
    audio = A.Audio(A.LJAudio)
    n_fft, hop, win = audio._stft_parameters()
    y = np.random.default_rng(9).standard_normal(hop * 17)
    D = audio._stft(y)
    w = A.pad_center(A.hann_periodic(win), n_fft)
    # scipy frames the reflect-padded signal exactly like librosa when nperseg == n_fft (the zero-padded window) and
    # padded=False; it divides by sum(window)
    _, _, Z = signal.stft(y, window=w, nperseg=n_fft, noverlap=n_fft - hop, boundary="even", padded=False, return_onesided=True)
    k = min(Z.shape[1], D.shape[1])
    assert k >= D.shape[1] - 1
    # 'even' extension == numpy 'reflect' padding
    assert np.abs(Z[:, :k] * w.sum() - D[:, :k]).max() < 1e-9 * np.abs(D).max()


def test_griffin_lim_reduces_spectral_inconsistency():
    """Each projection step cannot increase || |stft(y)| - S || (Griffin & Lim 1984): a known-answer property."""
    rng = np.random.default_rng(5)
    audio = A.Audio(A.LJAudio)
    t = np.arange(256 * 15) / 22050.0
    y = np.sin(2 * np.pi * 440 * t) * np.hanning(t.size) + 0.3 * np.sin(2 * np.pi * 1320 * t)
    S = np.abs(audio._stft(y))
    rand = rng.random(S.shape)
    errs = []
    for it in (0, 2, 8, 24):
        w = audio._griffin_lim(S, rand=rand, iters=it)
        errs.append(np.linalg.norm(np.abs(audio._stft(w)) - S) / np.linalg.norm(S))
    assert errs[0] > errs[1] > errs[2] > errs[3]
    assert errs[3] < 0.5 * errs[0]


def test_audio_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vaenar_tts_b200 import LJHPS
    from vaenar_tts_b200._lib import VaenarError
    from vaenar_tts_b200.audio import Audio
    a = Audio(LJHPS.Audio)
    with pytest.raises(VaenarError, match="no CPU fallback"):
        a.inv_mel_spectrogram(np.zeros((80, 10), np.float32))

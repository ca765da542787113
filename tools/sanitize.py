"""Tiny run of every C-ABI path (inference, call eval, call training-forward, init, two full train_steps + one with the fused training forward and the two-CTAs-per-SM GEMMs forced, mel inversion)
for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vaenar_oracle as O
from oracle.hparams import LJHPS as OH
from vaenar_tts_b200 import VAENAR, LJHPS
P = O.init_params(OH, seed=1, zero_init_std=0.02)
m = VAENAR(LJHPS, device="cuda"); m.load_state_dict(P)
texts, mels, t_len, m_len = O.synthetic_batch(OH, 4, 20, 70)
mel, ali = m.inference(texts, m_len, t_len, reduction_factor=2)
out = m(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len, reduction_factor=2, training=False, reduce_loss=True)
out2 = m(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len, reduction_factor=2, training=True, reduce_loss=True)
m.init(texts, m_len, t_len)
for _ in range(2):
    loss = m.train_step(texts, mels, t_len, m_len, 1e-5, 2)
# the variants the default only picks at C3-sized batches: training forward through the fused row kernel (tape stores),
# two-CTAs-per-SM GEMM instances (incl. the ReLU-masked dgrad)
from vaenar_tts_b200 import _lib
lib = _lib.load()
lib.vaenar_set_train_fused(1); lib.vaenar_set_gemm_occ2(2)
loss = m.train_step(texts, mels, t_len, m_len, 1e-5, 2)
lib.vaenar_set_train_fused(-1); lib.vaenar_set_gemm_occ2(1)
from vaenar_tts_b200.audio import Audio
a = Audio(LJHPS.Audio, device="cuda")
lens = [int(x) for x in m_len.clamp(min=2)]
wav = a.inv_mel_spectrogram_batch(mel.clamp(0, 1), lens, seed=3, iters=3)
pcm = a.to_int16_batch(a.inv_preemphasize_batch(wav, lens), lens)
torch.cuda.synchronize(); print("ok", float(mel.abs().mean()), float(out[2]), float(loss[0]), int(pcm.abs().max()))

"""Mel inversion on BASELINE config 2's output shape: B16 x 870 frames, 60 Griffin-Lim iterations.
Prints device time per batch (CUDA events), seconds of audio per second, and the numpy oracle on a bounded sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from vaenar_tts_b200 import LJHPS  # noqa: E402
from vaenar_tts_b200.audio import Audio  # noqa: E402
from golden_util import speechlike_mel  # noqa: E402

B, T, iters = 16, 870, 60
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
a = Audio(LJHPS.Audio)
rng = np.random.default_rng(0)
mel = torch.from_numpy(np.stack([speechlike_mel(rng, T) for _ in range(B)])).cuda()
lens = [T] * B
for _ in range(2):
    wav = a.inv_mel_spectrogram_batch(mel, lens, seed=1)
    a.to_int16_batch(a.inv_preemphasize_batch(wav, lens), lens)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    wav = a.inv_mel_spectrogram_batch(mel, lens, seed=1)
    pcm = a.to_int16_batch(a.inv_preemphasize_batch(wav, lens), lens)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms.sort()
t_gpu = ms[len(ms) // 2] * 1e-3
audio_s = B * 256 * (T - 1) / 22050.0
out = {"workload": f"mel inversion B{B} x {T} frames, {iters} Griffin-Lim iterations, fp64", "ms_per_batch": t_gpu * 1e3,
       "audio_seconds_per_second": audio_s / t_gpu, "frame_iterations_per_s": B * T * (iters + 1) / t_gpu}
if "--cpu" in sys.argv:
    from oracle import audio_oracle as A
    o = A.Audio(A.LJAudio)
    Tc = 200
    m = speechlike_mel(rng, Tc)
    S = o.linear_magnitudes(m.T)
    r = rng.random(S.shape)
    t0 = time.perf_counter()
    o._griffin_lim(S, rand=r, iters=iters)
    dt = time.perf_counter() - t0
    out["cpu_oracle"] = {"sample": f"1 utterance x {Tc} frames, {iters} iterations, numpy (threads: {os.cpu_count()})",
                         "audio_seconds_per_second": 256 * (Tc - 1) / 22050.0 / dt,
                         "frame_iterations_per_s": Tc * (iters + 1) / dt}
print(json.dumps(out))

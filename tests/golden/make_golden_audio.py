"""Golden vectors of the mel-inversion oracle (oracle/audio_oracle.py): small utterances, injected random phases.
Run from the repo root:  python tests/golden/make_golden_audio.py   (numpy only; librosa 0.8.0 is not installable here,
see the oracle's header)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import audio_oracle as A  # noqa: E402
from golden_util import speechlike_mel  # noqa: E402


def main():
    out = {}
    for name, hps, T, iters in (("lj", A.LJAudio, 14, 8), ("db", A.DataBakerAudio, 9, 5)):
        rng = np.random.default_rng(2024 + T)
        audio = A.Audio(hps)
        mel = speechlike_mel(rng, T)
        S = audio.linear_magnitudes(mel.T)                       # [num_freq, T] float32
        rand = rng.random(S.shape)
        wav = audio._griffin_lim(S, rand=rand, iters=iters)
        pcm, pre = A.synthesize(audio, mel, rand=rand, iters=iters)
        out.update({f"{name}_mel": mel, f"{name}_S": S, f"{name}_rand": rand, f"{name}_iters": np.int32(iters),
                    f"{name}_wav": wav, f"{name}_pre": pre, f"{name}_pcm": pcm})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "audio_griffin_lim.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""Helpers shared by the golden-vector tests (oracle and CUDA path)."""
import hashlib
import os

import numpy as np
import torch

from oracle import vaenar_oracle as O
from oracle.hparams import LJHPS, DataBakerHPS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "lj_b3_t21_m67_rf2.npz": LJHPS,
    "lj_b2_t9_m43_rf3.npz": LJHPS,
    "db_b2_t17_m50_rf2.npz": DataBakerHPS,
}


def weights_checksum(P):
    h = hashlib.sha256()
    for k in sorted(P):
        h.update(k.encode())
        h.update(P[k].detach().cpu().numpy().astype(np.float32).tobytes())
    return h.hexdigest()


def load_case(name):
    """Returns (hps, golden dict of torch tensors / numpy scalars, params regenerated from the seed)."""
    hps = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False))
    seed = int(g["seed"])
    P = O.init_params(hps, seed=seed, zero_init_std=0.02)
    O.randomize_bn_stats(P, seed=seed + 1)
    assert weights_checksum(P) == str(g["weights_sha256"]), "weight RNG drifted from the golden generator"
    return hps, g, P


def t(g, key):
    return torch.from_numpy(np.asarray(g[key]))


def train_masks(hps, g, prefix="train"):
    """Map the reference's dropout call order (recorded by the shim) onto the oracle's mask sites."""
    n = int(g[f"{prefix}_n_dropout"])
    ms = [t(g, f"{prefix}_dropout_{i:02d}") for i in range(n)]
    sites = [f"enc.prenet.{i}" for i in range(hps.Encoder.n_conv)] + ["enc.pos"]
    if prefix == "train":
        sites += ["post.prenet.1", "post.prenet.2", "post.pos"]
    sites += [f"dec.postnet.{i}" for i in range(hps.Decoder.post_n_conv)]
    assert len(sites) == n, (len(sites), n)
    return dict(zip(sites, ms))


def speechlike_mel(rng, T, n_mels=80):
    """Smooth normalised mel in [0, 1] with silence at both ends (float32, as the decoder emits it): the synthetic
    input of the mel-inversion tests (tests/golden/make_golden_audio.py, tests/test_audio_gpu.py)."""
    tt = np.arange(T)[:, None] / max(T - 1, 1)
    m = np.arange(n_mels)[None, :] / n_mels
    env = np.sin(np.pi * tt) ** 0.5
    mel = 0.55 * env * (1 - 0.6 * m) + 0.15 * np.sin(2 * np.pi * (3 * tt + 2 * m)) * env + 0.05 * rng.standard_normal((T, n_mels))
    return np.clip(mel, 0, 1).astype(np.float32)

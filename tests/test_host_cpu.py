"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol declared in
include/vaenar_b200.h, the parameter manifest equals the oracle's (names, shapes, 485 trainable tensors /
34 725 865 parameters), and compute refuses to run without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

from oracle import vaenar_oracle as O
from oracle.hparams import LJHPS as OLJ, DataBakerHPS as ODB

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    return g.build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "vaenar_b200.h")).read()
    declared = set(re.findall(r"\b(vaenar_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vaenar_model"}
    lib = ctypes.CDLL(built)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from vaenar_tts_b200 import EXPORTS
    assert set(EXPORTS) == declared, set(EXPORTS) ^ declared
    assert lib.vaenar_abi_version() == 1


@pytest.mark.parametrize("which", ["lj", "db"])
def test_manifest_matches_oracle(built, which):
    from vaenar_tts_b200 import VAENAR, LJHPS, DataBakerHPS
    ohps, hps = (OLJ, LJHPS) if which == "lj" else (ODB, DataBakerHPS)
    m = VAENAR(hps, device="cpu")
    P = O.init_params(ohps, seed=0)
    sd = m.state_dict()
    assert set(sd) == set(P)
    for k, v in P.items():
        assert sd[k].numel() == v.numel(), k
    tv = m.trainable_variables
    assert len(tv) == 485
    assert sum(v.numel() for v in tv) == (34725865 if which == "lj" else 34725865 - 4 * 512)
    m.load_state_dict(P)
    back = m.state_dict()
    for k, v in P.items():
        assert torch.equal(back[k].reshape(v.shape), v), k


def test_default_init_statistics(built):
    """Keras default initialisers: zero-init projections, orthogonal InvertibleLinear, unit LN/BN scales."""
    from vaenar_tts_b200 import VAENAR, LJHPS
    m = VAENAR(LJHPS, device="cpu", seed=1)
    sd = m.state_dict()
    assert float(sd["posterior.mu_projection.kernel"].abs().max()) == 0.0
    assert float(sd["prior.glow.3.affine_coupling.net.shift_proj.kernel"].abs().max()) == 0.0
    w = sd["prior.glow.0.linear.weight"]
    assert torch.allclose(w @ w.T, torch.eye(128), atol=1e-5)
    k = sd["decoder.attentions.0.ffn.dense1.kernel"]
    lim = (6.0 / (256 + 1024)) ** 0.5
    assert float(k.abs().max()) <= lim and float(k.abs().max()) > 0.9 * lim
    assert float(sd["text_encoder.pos_weight"]) == 1.0


def test_compute_fails_loudly_without_gpu(built):
    from vaenar_tts_b200 import VAENAR, LJHPS, VaenarError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = VAENAR(LJHPS, device="cpu")
    with pytest.raises(VaenarError):
        m.inference(torch.zeros(2, 4, dtype=torch.int32), [8, 8], [4, 4])
    with pytest.raises(VaenarError):
        m.text_encoder(torch.zeros(2, 4, dtype=torch.int32), [4, 4], pos_step=1.0)


def test_workspace_query_and_unsupported_hparams(built):
    from vaenar_tts_b200 import VAENAR, LJHPS, VaenarError
    m = VAENAR(LJHPS, device="cpu")
    a = m._lib.vaenar_workspace_bytes(m._h, 16, 148, 435, 2)
    b = m._lib.vaenar_workspace_bytes(m._h, 32, 148, 435, 2)
    assert 0 < a < b < (1 << 31)

    class Bad(LJHPS):
        class Common(LJHPS.Common):
            latent_dim = 64
    with pytest.raises(VaenarError):
        VAENAR(Bad, device="cpu")


def test_training_host_logic_without_gpu(built):
    """train_step must fail loudly without a GPU; the training workspace query (a dry run of the whole forward + backward
    launch sequence) and the optimiser sharding arithmetic are pure host code."""
    from vaenar_tts_b200 import VAENAR, LJHPS, VaenarError, _lib
    lib = _lib.load()
    m = VAENAR(LJHPS, device="cpu")
    small = int(lib.vaenar_train_workspace_bytes(m._h, 4, 64, 128, 2))
    big = int(lib.vaenar_train_workspace_bytes(m._h, 32, 148, 435, 2))
    assert 0 < small < big and 4e9 < big < 8e9, (small, big)          # C3: every saved activation, ~5.5 GB
    assert big > 10 * int(lib.vaenar_workspace_bytes(m._h, 32, 148, 435, 2))   # inference reuses its buffers
    n = m.flat_parameters().numel()
    for world in (1, 2, 3, 8):
        S = int(lib.vaenar_adam_shard_floats(n, world))
        assert S % 4 == 0 and S * world >= n and S * (world - 1) < n + 4 * world
    if not torch.cuda.is_available():
        with pytest.raises(VaenarError):
            m.train_step(torch.zeros(2, 4, dtype=torch.int32), torch.zeros(2, 8, 80), [4, 4], [8, 8], 1e-5, 2)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, on the host cores) prints ONE JSON line with
    the contract keys, needs no GPU, and names the same metric / workload as our arm."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mel-frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("C2:") and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-6 * d["value"]

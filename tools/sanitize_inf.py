"""compute-sanitizer target: one small VAENAR.inference + eval call (fused row kernel, flow tail, attention3)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vaenar_oracle as O
from oracle.hparams import LJHPS as OH
from vaenar_tts_b200 import VAENAR, LJHPS
P = O.init_params(OH, seed=1, zero_init_std=0.02)
m = VAENAR(LJHPS, device="cuda"); m.load_state_dict(P)
texts, mels, t_len, m_len = O.synthetic_batch(OH, 2, 20, 70)
mel, ali = m.inference(texts, m_len, t_len, reduction_factor=2)
torch.cuda.synchronize(); print("ok", float(mel.abs().mean()))

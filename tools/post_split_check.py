"""Accuracy / time of the PostNet with the first k convolutions on plain fp16 operands (VAENAR_POST_PLAIN=k)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vaenar_oracle as O
from oracle.hparams import LJHPS as OH
from vaenar_tts_b200 import VAENAR, LJHPS, InferenceSession
B, Tt, Tm, rf = 16, 148, 870, 2
P = O.init_params(OH, seed=21, zero_init_std=0.02)
O.randomize_bn_stats(P, seed=22)
texts, mels, t_len, m_len = O.synthetic_batch(OH, B, Tt, Tm, rf=rf, seed=23)
Tz = int(((m_len + rf - 1) // rf).max())
eps = torch.randn(B, Tz, 128, generator=torch.Generator().manual_seed(3))
ref_path = "/tmp/post_ref.pt"
if os.path.exists(ref_path):
    ref = torch.load(ref_path)
else:
    torch.set_num_threads(16)
    with torch.no_grad():
        ref, _, _ = O.vaenar_inference(P, OH, texts, m_len, t_len, rf, eps)
    torch.save(ref, ref_path)
model = VAENAR(LJHPS, device="cuda:0")
model.load_state_dict(P)
mel, _ = model.inference(texts, m_len, t_len, reduction_factor=rf, epsilon=eps)
torch.cuda.synchronize()
mask = O.sequence_mask(m_len, ref.shape[1], torch.float32)[:, :, None]
mae = float((((mel.cpu() - ref).abs() * mask).sum() / (mask.sum() * 80)))
sess = InferenceSession(model, B, Tt, Tz, rf=rf)
sess.set_inputs(texts, t_len, m_len); sess.run_e2e(); torch.cuda.synchronize(); sess.capture()
for _ in range(5): sess.run_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): sess.run_device()
e1.record(); torch.cuda.synchronize()
print(f"POST_PLAIN={os.environ.get('VAENAR_POST_PLAIN','0')}: mel MAE {mae:.3e}  serial step {e0.elapsed_time(e1)/20:.4f} ms")

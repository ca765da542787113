"""Profiling driver: init + N train_steps at C3 (or a given batch) -- run under ncu for the launch list."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from oracle.hparams import LJHPS as OH  # noqa: E402
from oracle.vaenar_oracle import synthetic_batch  # noqa: E402
from vaenar_tts_b200 import LJHPS, VAENAR  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
texts, mels, t_len, m_len = synthetic_batch(OH, B, 148, 870, seed=1)
m = VAENAR(LJHPS, device="cuda:0", seed=1)
d = [x.cuda() for x in (texts, mels, t_len, m_len)]
m.init(d[0], d[3], d[2])
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    out = m.train_step(d[0], d[1], d[2], d[3], 1e-5, 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print([float(x) for x in out])

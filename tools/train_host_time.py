"""Is train_step bound by the host issuing launches or by the device?  For each shape: host time until train_step
returns (all launches queued) vs device time until the stream drains.  Usage: python tools/train_host_time.py [reps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from vaenar_tts_b200 import VAENAR, LJHPS  # noqa: E402
from oracle.vaenar_oracle import synthetic_batch  # noqa: E402
from oracle.hparams import LJHPS as OLJ  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda:0"
import os
shapes = ((32, 148, 870), (12, 148, 870), (4, 148, 870), (32, 64, 256))
if os.environ.get("SHAPES") == "2":
    shapes = ((32, 148, 870), (4, 148, 870))
for (B, Tt, Tm) in shapes:
    texts, mels, t_len, m_len = synthetic_batch(OLJ, B, Tt, Tm, seed=1)
    d = [x.to(dev) for x in (texts, mels, t_len, m_len)]
    model = VAENAR(LJHPS, device=dev, seed=1)
    model.init(d[0], d[3], d[2])
    for _ in range(3):
        model.train_step(d[0], d[1], d[2], d[3], 1e-5, 2)
    torch.cuda.synchronize()
    host, total = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.train_step(d[0], d[1], d[2], d[3], 1e-5, 2)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host.append((t1 - t0) * 1e3)
        total.append((t2 - t0) * 1e3)
    host.sort(); total.sort()
    print(f"B{B} Tt{Tt} Tm{Tm}: host issue {host[len(host)//2]:.2f} ms, step {total[len(total)//2]:.2f} ms", flush=True)
    del model

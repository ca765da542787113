import sys, time, torch
sys.path.insert(0,'/root/repo')
from oracle.hparams import LJHPS as OH
from oracle.vaenar_oracle import synthetic_batch
from vaenar_tts_b200 import LJHPS, VAENAR
B=32
texts, mels, t_len, m_len = synthetic_batch(OH, B, 148, 870, seed=1)
m = VAENAR(LJHPS, device="cuda:0", seed=1)
d=[x.cuda() for x in (texts, mels, t_len, m_len)]
m.init(d[0], d[3], d[2])
for i in range(3): m.train_step(d[0], d[1], d[2], d[3], 1e-5, 2)
torch.cuda.synchronize()
for i in range(3):
    t0=time.perf_counter()
    m.train_step(d[0], d[1], d[2], d[3], 1e-5, 2)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    print(f"enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
import os
os.environ['X']='1'

"""Throughput of N independent InferenceSessions replayed concurrently on N streams (C2 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vaenar_tts_b200 import VAENAR, LJHPS, InferenceSession
from oracle.vaenar_oracle import synthetic_batch
from oracle.hparams import LJHPS as OH
B, Tt, Tm, RF = 16, 148, 870, 2
Tz = (Tm + RF - 1) // RF
model = VAENAR(LJHPS, device="cuda:0", seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nsess in (1, 2, 3, 4):
    sess, streams = [], []
    for i in range(nsess):
        texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm, seed=100 + i)
        s = InferenceSession(model, B, Tt, Tz, rf=RF, return_alignments=False, seed=i)
        s.set_inputs(texts, t_len, m_len)
        s.run_e2e(); torch.cuda.synchronize()
        s.capture()
        sess.append(s); streams.append(torch.cuda.Stream())
    steps = 40
    for rep in range(2):
        torch.cuda.synchronize()
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for k in range(steps):
            i = k % nsess
            with torch.cuda.stream(streams[i]):
                sess[i].run_device()
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{nsess} sessions in flight: {ms/steps:.4f} ms per step (amortised), {B*Tm*steps/(ms/1e3)/1e6:.2f} M frames/s")

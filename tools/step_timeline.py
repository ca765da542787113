"""Timeline of the GEMM launches inside one real VAENAR.inference step at C2 (eager launches): in-kernel span of
every GEMM (first CTA start -> last CTA end, globaltimer) and the gap to the next GEMM.  Tuning aid."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vaenar_tts_b200 import VAENAR, LJHPS, InferenceSession, _lib
from oracle.vaenar_oracle import synthetic_batch
from oracle.hparams import LJHPS as OH
lib = _lib.load()
B, Tt, Tm = 16, 148, 870
texts, mels, t_len, m_len = synthetic_batch(OH, B, Tt, Tm)
m = VAENAR(LJHPS, device="cuda")
s = InferenceSession(m, B, Tt, 435, rf=2)
s.set_inputs(texts, t_len, m_len)
for _ in range(3):
    s.run_e2e()
torch.cuda.synchronize()
buf = torch.zeros(64 << 20, dtype=torch.int64, device="cuda")
lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(buf.data_ptr()))
s._launch()
torch.cuda.synchronize()
launches = json.loads(lib.vaenar_debug_gemm_launches().decode())
lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(0))
t = buf.cpu()
rows = []
for L in launches:
    n = L["grid"][0] * L["grid"][1]
    a = t[L["offset"]: L["offset"] + n * 8].view(n, 8).double()
    rows.append(dict(L, start=a[:, 0].min().item(), end=a[:, 4].max().item(),
                     setup=(a[:, 1] - a[:, 0]).mean().item(), main=(a[:, 2] - a[:, 1]).mean().item(),
                     epi=(a[:, 3] - a[:, 2]).mean().item()))
tot_span = sum(r["end"] - r["start"] for r in rows) / 1e3
wall = (rows[-1]["end"] - rows[0]["start"]) / 1e3
print(f"{len(rows)} GEMM launches; sum of spans {tot_span:.0f} us; first->last wall {wall:.0f} us (eager, includes attention + host gaps)")
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for r in rows:
    k = (tuple(r["grid"]), r["mode"], r["N"], r["K"], r["bn"])
    a = agg[k]
    a[0] += 1; a[1] += (r["end"] - r["start"]) / 1e3; a[2] += r["setup"] / 1e3; a[3] += r["main"] / 1e3; a[4] += r["epi"] / 1e3
print("grid        mode N    K     bn  | n  span_us  setup  main  epi   (mode 100/101/102 = self/cross/cross+ali attention: setup, pass1, pass2, epilogue->[3..4])")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{str(k[0]):11s} {k[1]:4d} {k[2]:4d} {k[3]:5d} {k[4]:4d} | {a[0]:2d} {a[1]/a[0]:7.2f} {a[2]/a[0]:6.2f} {a[3]/a[0]:6.2f} {a[4]/a[0]:6.2f}")

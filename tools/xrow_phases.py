"""Phase timeline of the fused per-CrossAttentionBLK row kernel (csrc/xblk_fused.cuh): per-CTA globaltimer stamps of the
epilogue, MMA-issuer and TMA-producer roles, median over the CTAs of one launch, relative to the epilogue's first stamp."""
import ctypes
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vaenar_tts_b200 import VAENAR, LJHPS, _lib  # noqa: E402

B, T, Tt = int(os.environ.get("XB", 8)), 435, 148
lib = _lib.load()
model = VAENAR(LJHPS, device="cuda:0", seed=1)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, 256, generator=g)
mem = torch.randn(B, Tt, 512, generator=g)
ql = torch.full((B,), T, dtype=torch.int32)
tl = torch.full((B,), Tt, dtype=torch.int32)
for _ in range(3):
    model.cross_attention_blocks(("prior", 0), x, mem, ql, tl)
torch.cuda.synchronize()
tiles = (T + 127) // 128
n = tiles * B * 128
buf = torch.zeros(2 * n, dtype=torch.int64, device="cuda")
lib.vaenar_debug_xrow_timestamps(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.cross_attention_blocks(("prior", 0), x, mem, ql, tl)
e1.record()
torch.cuda.synchronize()
lib.vaenar_debug_xrow_timestamps(ctypes.c_void_p(0))
d = buf.cpu().numpy().reshape(2, tiles * B, 128).astype(np.float64)
names = {0: "epi start", 1: "R init done", 2: "LN1 in", 3: "LN1 out", 4: "cq in", 5: "q out", 22: "LN2 in", 23: "LN2 out",
         40: "LN3 in", 41: "LN3 out", 48: "epi end", 64: "mma start", 65: "proj1-x issued", 66: "a1 landed", 67: "proj1 issued",
         68: "s ready", 69: "proj2a+cq issued", 70: "q ready", 79: "a2 ready", 80: "proj2b issued", 81: "c ready",
         82: "ffn issued", 83: "x' ready", 84: "qkv issued", 96: "tma A issued", 97: "stage free", 98: "proj1 fills", 99: "w2a+cq fills",
         100: "attn fills", 101: "w2b fills", 102: "ffn fills", 103: "qkv fills"}
names.update({105: "sm1: S in regs", 106: "sm1: local max done", 107: "sm1: after bar 1", 108: "sm1: exp done", 109: "sm1: after bar 2",
              110: "LN2: R in regs", 111: "LN2: stats done", 112: "LN2: after bar", 113: "LN2: normalised", 114: "LN2: tmem st issued",
              115: "LN2: panels written", 116: "LN2: tmem st waited", 117: "LN2: proxy fence done", 118: "Rinit: loads landed", 119: "Rinit: st issued",
              })
for h in range(4):
    names[6 + 4 * h] = f"S{h} in"; names[7 + 4 * h] = f"P{h} out"; names[8 + 4 * h] = f"O{h} in"; names[9 + 4 * h] = f"O{h} out"
    names[71 + 2 * h] = f"P{h} seen (mma)"; names[72 + 2 * h] = f"PV{h}+S{h+1} issued"
for j in range(8):
    names[24 + 2 * j] = f"hid{j} in"; names[25 + 2 * j] = f"hid{j} out"
for j in range(6):
    names[42 + j] = f"qkv{j} out"
print(f"two launches (block 0 has_next=1, block 1 has_next=0), total incl. QKV GEMM + 2 self-attn: {e0.elapsed_time(e1)*1e3:.1f} us")
for li in range(2):
    dd = d[li]
    raw = buf.cpu().numpy().reshape(2, tiles * B, 128)[li]
    base = dd[:, 0:1]
    rel = (dd - base) / 1e3
    print(f"--- launch {li}: CTA spread of start {np.ptp(dd[:,0])/1e3:.1f} us, kernel span {(dd[:,48].max()-dd[:,0].min())/1e3:.1f} us")
    order = sorted(names, key=lambda k: np.median(rel[:, k]) if dd[:, k].min() > 0 else 1e18)
    for k in order:
        if dd[:, k].min() <= 0 or k >= 121:
            continue
        print(f"  {np.median(rel[:, k]):8.2f} us  (min {rel[:, k].min():7.2f} max {rel[:, k].max():7.2f})  [{k:3d}] {names[k]}")


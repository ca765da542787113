"""Per-CTA phase timing of the stand-alone attention kernel at the C2 causal self-attention shape."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_util as G
from vaenar_tts_b200 import _lib
lib = _lib.load()
g = torch.Generator().manual_seed(0)
for (B, Tq, Tk, causal) in [(16, 435, 435, True), (8, 435, 435, True), (16, 435, 148, False)]:
    H = 4
    q = torch.randn(B, Tq, 256, generator=g); k = torch.randn(B, Tk, 256, generator=g); v = torch.randn(B, Tk, 256, generator=g)
    ql = torch.full((B,), Tq, dtype=torch.int32); kl = torch.full((B,), Tk, dtype=torch.int32)
    G.attention(q, k, v, ql, kl, H, causal, want_ali=False)
    nct = ((Tq + 127) // 128) * H * B
    buf = torch.zeros(nct * 8, dtype=torch.int64, device="cuda")
    lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(buf.data_ptr()))
    G.attention(q, k, v, ql, kl, H, causal, want_ali=False)
    lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(0))
    t = buf.cpu().view(nct, 8).double()
    t0 = t[:, 0].min()
    span = (t[:, 4].max() - t0).item() / 1e3
    # grid order: x = q tile fastest
    for qt in range((Tq + 127) // 128):
        sel = t[qt::(Tq + 127) // 128]
        ph = [(sel[:, i + 1] - sel[:, i]).mean().item() / 1e3 for i in range(4)]
        print(f"B{B} Tq{Tq} Tk{Tk} causal={causal} qtile {qt}: nblk {sel[:,5].mean().item():.0f} setup {ph[0]:5.2f} pass1 {ph[1]:5.2f} pass2 {ph[2]:5.2f} epi {ph[3]:5.2f} "
              f"| start {((sel[:,0]-t0).mean()/1e3):6.2f} end {((sel[:,4]-t0).mean()/1e3):6.2f}")
    print(f"   kernel span {span:.2f} us, CTAs {nct}")

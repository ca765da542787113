"""Per-CTA phase timing of the GEMM kernel (globaltimer stamps): start -> setup done -> accumulator ready ->
epilogue done -> CTA done.  Tuning aid."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_util as G
from vaenar_tts_b200 import _lib
lib = _lib.load()
g = torch.Generator().manual_seed(0)
M = 6960
cases = {
  "cq  K256 N256 plain->h (bn128)": dict(K=256, N=256, block_n=128),
  "ffn1 K256 N1024 relu (bn128)": dict(K=256, N=1024, block_n=128, act=1),
  "ffn2 K1024 N256 LN cluster (bn128)": dict(K=1024, N=256, block_n=128, ln=True),
  "proj K512 N256 LN cluster (bn128)": dict(K=512, N=256, block_n=128, ln=True),
  "ffn2 K1024 N256 LN single (bn256)": dict(K=1024, N=256, block_n=256, ln=True),
}
for name, c in cases.items():
    K, N = c["K"], c["N"]
    A = torch.randn(M, K, generator=g); W = torch.randn(K, N, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    kw = dict(act=c.get("act", 0), block_n=c["block_n"])
    if c.get("ln"):
        kw.update(residual=torch.randn(M, N, generator=g), gamma=torch.ones(N), beta=torch.zeros(N), ln=True)
    G.dense(A, W, b, **kw)
    nct = ((M + 127) // 128) * ((N + c["block_n"] - 1) // c["block_n"])
    buf = torch.zeros(nct * 8, dtype=torch.int64, device="cuda")
    lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(buf.data_ptr()))
    G.dense(A, W, b, **kw)
    lib.vaenar_debug_gemm_timestamps(ctypes.c_void_p(0))
    t = buf.cpu().view(nct, 8).double()
    t0 = t[:, 0].min()
    ph = [(t[:, i + 1] - t[:, i]).mean().item() / 1e3 for i in range(4)]
    print(f"{name:40s} ctas={nct:4d} kernel span={(t[:,4].max()-t0).item()/1e3:7.2f}us | setup {ph[0]:6.2f} mainloop-wait {ph[1]:6.2f} "
          f"epilogue {ph[2]:6.2f} tail-sync {ph[3]:6.2f} | first-start spread {(t[:,0].max()-t0).item()/1e3:6.2f}us")

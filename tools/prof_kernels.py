"""Launch the hot kernels once each at C2 shapes through the block-level C-ABI hooks (for ncu captures)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_util as G

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B, H, Tz, Tt = 16, 4, 435, 148
g = torch.Generator().manual_seed(0)
if which in ("all", "attn"):
    q = torch.randn(B, Tz, 256, generator=g); k = torch.randn(B, Tz, 256, generator=g); v = torch.randn(B, Tz, 256, generator=g)
    ql = torch.randint(300, Tz + 1, (B,), generator=g); ql[0] = Tz
    for _ in range(2):
        G.attention(q, k, v, ql, ql, H, True, want_ali=False)                 # decoder/prior causal self-attention
    km = torch.randn(B, Tt, 256, generator=g); vm = torch.randn(B, Tt, 256, generator=g)
    kl = torch.randint(80, Tt + 1, (B,), generator=g); kl[0] = Tt
    for _ in range(2):
        G.attention(q, km, vm, ql, kl, H, False, want_ali=False)              # cross-attention over the text memory
    G.attention(q, km, vm, ql, kl, H, False, want_ali=True)                   # decoder cross-attention with alignments
if which in ("all", "gemm"):
    M = B * Tz
    A = torch.randn(M, 256, generator=g); W = torch.randn(256, 1024, generator=g) / 16; b = torch.randn(1024, generator=g)
    for _ in range(2):
        G.dense(A, W, b, act=1, block_n=128)                                  # FFN dense1
    A2 = torch.randn(M, 1024, generator=g); W2 = torch.randn(1024, 256, generator=g) / 32; b2 = torch.randn(256, generator=g)
    res = torch.randn(M, 256, generator=g); gm = torch.ones(256); bt = torch.zeros(256)
    for _ in range(2):
        G.dense(A2, W2, b2, residual=res, gamma=gm, beta=bt, ln=True, block_n=256)   # FFN dense2 + residual + LN
if which in ("all", "gemm2"):
    # the two most frequent training GEMMs at C3 (13920 rows): grids deeper than one wave -> two-CTAs-per-SM instances
    M = 32 * Tz
    A = torch.randn(M, 1024, generator=g); W = torch.randn(1024, 256, generator=g) / 32; res = torch.randn(M, 256, generator=g)
    for _ in range(2):
        G.dense(A, W, None, residual=res, block_n=128)                        # dgrad through ffn.dense1, accumulated into g (fp32)
    A2 = torch.randn(M, 256, generator=g); W2 = torch.randn(256, 1024, generator=g) / 16
    for _ in range(2):
        G.dense(A2, W2, None, block_n=128)                                    # K 256 -> N 1024, fp32 out
if which in ("all", "wgrad"):
    Bw, T = 32, 435                                                                   # C3 token count
    X = torch.randn(Bw, T, 256, generator=g); dY = torch.randn(Bw, T, 1024, generator=g)
    for _ in range(2):
        G.wgrad(X, dY)                                                                # FFN dense1 weight gradient [256, 1024]
    dY2 = torch.randn(Bw, T, 256, generator=g)
    for _ in range(2):
        G.wgrad(X, dY2, X2=X)                                                         # att_proj ([x ; ctx]) weight gradient [512, 256]
if which in ("all", "attn_bwd"):
    Bw = 32
    q = torch.randn(Bw, Tz, 256, generator=g); k = torch.randn(Bw, Tz, 256, generator=g); v = torch.randn(Bw, Tz, 256, generator=g)
    do = torch.randn(Bw, Tz, 256, generator=g)
    ql = torch.randint(300, Tz + 1, (Bw,), generator=g); ql[0] = Tz
    for _ in range(2):
        G.attention_bwd(q, k, v, do, ql, ql, H, True)                                 # causal self-attention backward
    km = torch.randn(Bw, Tt, 256, generator=g); vm = torch.randn(Bw, Tt, 256, generator=g)
    kl = torch.randint(80, Tt + 1, (Bw,), generator=g); kl[0] = Tt
    for _ in range(2):
        G.attention_bwd(q, km, vm, do, ql, kl, H, False)                              # cross-attention backward
print("done")

import sys, ctypes, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from vaenar_tts_b200 import VAENAR, LJHPS, _lib
from vaenar_tts_b200._lib import check
lib=_lib.load()
m=VAENAR(LJHPS, device="cuda:0", seed=1)
m2=VAENAR(LJHPS, device="cuda:0", seed=1)
n=m._flat.numel()
g=torch.randn(n, device="cuda")*65.0
host=torch.zeros(n,dtype=torch.uint8); check(lib.vaenar_trainable_mask(m._h, ctypes.c_void_p(host.data_ptr()))); mask=host.cuda()
S=int(lib.vaenar_adam_shard_floats(n,1)); print(n,S)
mm=torch.zeros(S,device="cuda"); vv=torch.zeros(S,device="cuda")
pp=(ctypes.c_void_p*1)(m2._flat.data_ptr()); pg=(ctypes.c_void_p*1)(g.data_ptr())
for step in (1,2):
    m.apply_gradients(g, step, grad_scale=1/65536.)
    check(lib.vaenar_adam_step_sharded(ctypes.cast(pp,ctypes.c_void_p), ctypes.cast(pg,ctypes.c_void_p), mm.data_ptr(), vv.data_ptr(), mask.data_ptr(), n, 0, 1, step, 1.25e-4, 0.9, 0.999, 1e-7, 1/65536., None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(step, float((m._flat-m2._flat).abs().max()))

"""Per-tensor comparison of the CUDA train_step gradients with autograd of the CPU oracle on a golden case
(diagnostic companion of tests/test_train_gpu.py).  usage: python tools/grad_check.py [case-index] [kl_weight]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch  # noqa: E402

from golden_util import CASES, load_case  # noqa: E402
from test_train_gpu import oracle_grads, cuda_grads  # noqa: E402
from test_model_gpu import make_model  # noqa: E402

case = list(CASES)[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
klw = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
ohps, g, P = load_case(case)
ref_losses, ref = oracle_grads(ohps, g, P, klw)
_, ref16 = oracle_grads(ohps, g, P, klw, emulate_fp16=True)
m = make_model(ohps, P)
losses, got = cuda_grads(m, g, klw)
print("case", case, "kl_weight", klw)
print("losses cuda  ", losses)
print("losses oracle", list(ref_losses))
nbad = 0
for k, r in ref.items():
    a, b = got[k].double().reshape(-1), r.double().reshape(-1)
    nb, na = float(b.norm()), float(a.norm())
    err = float((a - b).norm()) / (nb + 1e-30)
    cos = float((a @ b) / (na * nb + 1e-30))
    c = ref16[k].double().reshape(-1)
    err16 = float((a - c).norm()) / (float(c.norm()) + 1e-30)
    flag = "" if (err16 < 2e-2 and err < 1e-1 and cos > 0.995) or nb < 1e-6 else "   <<<<"
    nbad += bool(flag)
    print(f"{k:90s} |ref| {nb:10.3e} |got| {na:10.3e} err {err:9.2e} err16 {err16:9.2e} cos {cos:9.6f}{flag}")
print("bad tensors:", nbad, "of", len(ref))

"""Debug driver for the peer-memory optimizer: torchrun --nproc-per-node 2 tools/peer_check.py"""
import os, sys, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
def log(*a):
    print(f"[rank {rank}]", *a, flush=True)
try:
    from golden_util import CASES, load_case, t
    from test_model_gpu import make_model
    ohps, g, P = load_case(list(CASES)[0])
    m1, m2 = make_model(ohps, P), make_model(ohps, P)
    log("models built")
    m2.enable_peer_optimizer()
    log("peer optimizer enabled", [hex(x) for x in m2._peer_params])
    n = m1.flat_parameters().numel()
    S = 65536.0
    for step in (1, 2):
        grads = torch.randn(n, generator=torch.Generator().manual_seed(100 * step + rank)).cuda() * S * 1e-3
        ref = grads.clone(); dist.all_reduce(ref)
        m1.apply_gradients(ref, step, grad_scale=1.0 / (S * world))
        m2._grads.copy_(grads)
        m2._peer_adam(step, 1.0 / (S * world))
        torch.cuda.synchronize()
        log("step", step, "max diff", float((m1.flat_parameters() - m2.flat_parameters()).abs().max()))
    # timing of the optimiser half alone (CUDA events, 20 iterations each)
    def timed(fn, it=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / it
    gbuf = torch.randn(n, device="cuda")
    step_box = [10]
    def nccl_path():
        step_box[0] += 1
        dist.all_reduce(gbuf)
        m1.apply_gradients(gbuf, step_box[0], grad_scale=1e-9)
    def peer_path():
        step_box[0] += 1
        m2._peer_adam(step_box[0], 1e-9)
    def barrier_only():
        dist.all_reduce(m2._barrier_flag)
    log("ms: nccl all-reduce + adam", round(timed(nccl_path), 4), "| fused peer kernel + 2 barriers", round(timed(peer_path), 4),
        "| one tiny all-reduce", round(timed(barrier_only), 4))
    # real data-parallel step + replica agreement
    from vaenar_tts_b200 import parallel as PP
    full = [t(g, "texts"), t(g, "mels"), t(g, "t_len"), t(g, "m_len")]
    mine = PP.shard_batch(full, rank, world)
    m2.train_step(mine[0], mine[1], mine[2], mine[3], 1e-5, int(g["rf"]))
    torch.cuda.synchronize()
    flat = m2.flat_parameters().clone(); other = flat.clone(); dist.broadcast(other, 0)
    mask = m2._trainable_mask.bool()
    log("replica diff on trainable params", float((flat - other)[mask].abs().max()), "finite", bool(torch.isfinite(flat).all()))
except Exception:
    traceback.print_exc()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(1)
dist.barrier(); dist.destroy_process_group()

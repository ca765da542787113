"""Data-parallel train_step over NCCL: one process per GPU, ONE all-reduce of the flat gradient
buffer.  Needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:   # report instead of leaving the parent in q.get until its timeout
        import traceback
        q.put((rank, "error", traceback.format_exc()[-2000:]))
        raise


def _worker_body(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from golden_util import CASES, load_case, t
    from test_model_gpu import make_model, _masks
    from vaenar_tts_b200 import parallel as PP
    ohps, g, P = load_case(list(CASES)[0])
    m = make_model(ohps, P)
    full = [t(g, "texts"), t(g, "mels"), t(g, "t_len"), t(g, "m_len")]
    # identical data on both ranks: the all-reduced mean gradient must equal the local gradient
    losses, grads = m.train_step_grads(full[0], full[1], full[2], full[3], 1e-5, int(g["rf"]), eps=t(g, "train_eps"),
                                       dropout_masks=_masks(g, "train"), update_bn_stats=False)
    local = grads.clone()
    dist.all_reduce(grads)
    err = float((grads / world - local).norm() / local.norm())
    # sharded data (utt_ids[rank::size]): a full step runs, parameters stay identical on all ranks
    mine = PP.shard_batch(full, rank, world) if full[0].shape[0] >= world else full
    m.train_step(mine[0], mine[1], mine[2], mine[3], 1e-5, int(g["rf"]))
    flat = m.flat_parameters().clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    mask = m._trainable_mask.bool()
    same = float((flat - ref)[mask].abs().max())      # BatchNorm moving statistics are per-replica, everything else must agree
    q.put((rank, err, same, bool(torch.isfinite(flat).all())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_train_step_allreduce():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for r in res:
        assert r[1] != "error", r[2]
    res = sorted(res)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err, same, finite in res:
        assert err < 1e-3, (rank, err)        # atomics order only
        assert finite
    # trainable parameters bit-identical on both ranks after the step (same all-reduced gradients, same Adam)
    for rank, _, same, _ in res:
        assert same == 0.0, res


def _peer_worker(rank, world, port, q):
    try:
        _peer_worker_body(rank, world, port, q)
    except Exception as e:
        import traceback
        q.put((rank, "error", traceback.format_exc()[-2000:]))
        raise


def _peer_worker_body(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from golden_util import CASES, load_case, t
    from test_model_gpu import make_model
    from vaenar_tts_b200 import parallel as PP
    ohps, g, P = load_case(list(CASES)[0])
    m1, m2 = make_model(ohps, P), make_model(ohps, P)
    m2.enable_peer_optimizer()
    n = m1.flat_parameters().numel()
    S = 65536.0
    worst = 0.0
    for step in (1, 2, 3):
        grads = torch.randn(n, generator=torch.Generator().manual_seed(100 * step + rank)).cuda() * S * 1e-3
        ref = grads.clone()
        dist.all_reduce(ref)                                              # baseline: NCCL all-reduce + fused Adam
        m1.apply_gradients(ref, step, grad_scale=1.0 / (S * world))
        m2._grads.copy_(grads)                                            # fused: peer reduce-scatter -> Adam -> all-gather
        m2._ovf = torch.zeros(1, device="cuda")
        m2._peer_adam(step, 1.0 / (S * world))
        torch.cuda.synchronize()
        worst = max(worst, float((m1.flat_parameters() - m2.flat_parameters()).abs().max()))
    moved = 0.0
    # a real data-parallel step through the fused path: runs, finite, replicas agree on the trainable parameters
    full = [t(g, "texts"), t(g, "mels"), t(g, "t_len"), t(g, "m_len")]
    mine = PP.shard_batch(full, rank, world)
    m2.train_step(mine[0], mine[1], mine[2], mine[3], 1e-5, int(g["rf"]))
    torch.cuda.synchronize()
    flat = m2.flat_parameters().clone()
    other = flat.clone()
    dist.broadcast(other, 0)
    mask = m2._trainable_mask.bool()
    same = float((flat - other)[mask].abs().max())
    q.put((rank, worst, same, bool(torch.isfinite(flat).all()), moved))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_peer_memory_optimizer_matches_allreduce_adam():
    """vaenar_adam_step_sharded (gradient exchange + Adam in one kernel over NVLink peer memory, CUDA-IPC shared buffers)
    against NCCL all-reduce + vaenar_adam_step on identical gradients; then one real data-parallel train_step."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for r in res:
        assert r[1] != "error", r[2]
    res = sorted(res)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, worst, same, finite, _ in res:
        assert worst <= 1e-6, (rank, worst)       # same sums (two addends commute); Adam differs by FMA contraction (1 ulp)
        assert same == 0.0, (rank, same)          # every replica holds the same trainable parameters after the step
        assert finite


def test_sharded_adam_single_gpu_two_local_peers():
    """vaenar_adam_step_sharded with world = 2 where both "replicas" live on this GPU (plain device pointers instead of IPC
    mappings): rank 0 then rank 1 each reduce their gradient shard over both gradient buffers, apply Adam with their shard
    of the moments and write both parameter replicas.  Reference: vaenar_adam_step on the summed gradients.  Covers the
    exchange kernel on a 1-GPU box; also the overflow skip flag."""
    import ctypes
    from vaenar_tts_b200 import VAENAR, LJHPS, _lib
    from vaenar_tts_b200._lib import check
    lib = _lib.load()
    ref_m = VAENAR(LJHPS, device="cuda", seed=5)
    n = ref_m.flat_parameters().numel()
    world, S = 2, 65536.0
    p0 = ref_m.flat_parameters().clone()
    reps = [p0.clone(), p0.clone()]
    shard = int(lib.vaenar_adam_shard_floats(n, world))
    ms = [torch.zeros(shard, device="cuda") for _ in range(world)]
    vs = [torch.zeros(shard, device="cuda") for _ in range(world)]
    host = torch.zeros(n, dtype=torch.uint8)
    check(lib.vaenar_trainable_mask(ref_m._h, ctypes.c_void_p(host.data_ptr())))
    mask = host.cuda()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flag = torch.zeros(1, device="cuda")
    for step in (1, 2, 3):
        gs = [(torch.randn(n, generator=torch.Generator().manual_seed(10 * step + r)) * S * 1e-3).cuda() for r in range(world)]
        ref_m.apply_gradients(gs[0] + gs[1], step, grad_scale=1.0 / (S * world))
        pp = (ctypes.c_void_p * world)(*[t.data_ptr() for t in reps])
        pg = (ctypes.c_void_p * world)(*[t.data_ptr() for t in gs])
        for r in range(world):
            check(lib.vaenar_adam_step_sharded(ctypes.cast(pp, ctypes.c_void_p), ctypes.cast(pg, ctypes.c_void_p),
                                               ctypes.c_void_p(ms[r].data_ptr()), ctypes.c_void_p(vs[r].data_ptr()),
                                               ctypes.c_void_p(mask.data_ptr()), n, r, world, step, 1.25e-4, 0.9, 0.999, 1e-7,
                                               1.0 / (S * world), ctypes.c_void_p(flag.data_ptr()), stream))
        torch.cuda.synchronize()
        assert torch.equal(reps[0], reps[1])                                  # all-gather: both replicas identical
        tm = mask.bool()
        assert float((reps[0] - ref_m.flat_parameters())[tm].abs().max()) <= 1e-6
        assert torch.equal(reps[0][~tm], p0[~tm])                             # BatchNorm moving statistics untouched
    # overflow: a non-finite gradient entry anywhere -> the counted flag makes every rank skip the whole update
    before = reps[0].clone()
    gs[1][12345] = float("inf")
    flag.zero_()
    for r in range(world):
        check(lib.vaenar_grad_nonfinite(ctypes.c_void_p(gs[r].data_ptr()), n, ctypes.c_void_p(flag.data_ptr()), stream))
    assert float(flag) == 1.0
    pg = (ctypes.c_void_p * world)(*[t.data_ptr() for t in gs])
    for r in range(world):
        check(lib.vaenar_adam_step_sharded(ctypes.cast(pp, ctypes.c_void_p), ctypes.cast(pg, ctypes.c_void_p),
                                           ctypes.c_void_p(ms[r].data_ptr()), ctypes.c_void_p(vs[r].data_ptr()),
                                           ctypes.c_void_p(mask.data_ptr()), n, r, world, 4, 1.25e-4, 0.9, 0.999, 1e-7,
                                           1.0 / (S * world), ctypes.c_void_p(flag.data_ptr()), stream))
    torch.cuda.synchronize()
    assert torch.equal(reps[0], before) and torch.equal(reps[1], before)

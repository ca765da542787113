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
    same = float((flat - ref).abs().max())
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
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err, same, finite in res:
        assert err < 1e-3, (rank, err)        # atomics order only
        assert finite
    # trainable parameters identical on both ranks after the step (BatchNorm moving statistics are per-replica)
    assert res[1][2] < 1.0, res


def _peer_worker(rank, world, port, q):
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
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, worst, same, finite, _ in res:
        assert worst <= 1e-6, (rank, worst)       # same sums (two addends commute); Adam differs by FMA contraction (1 ulp)
        assert same == 0.0, (rank, same)          # every replica holds the same trainable parameters after the step
        assert finite

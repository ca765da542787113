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

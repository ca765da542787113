"""world_size-2 gloo tests (CPU) of the data-parallel host logic: utterance sharding (batch tensors and the sharded
TFRecord dataset: same batch count on every rank), the parameter broadcast, the all-reduced overflow count that makes every
replica skip the same train step, and the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vaenar_tts_b200 import parallel as P
    from oracle import vaenar_oracle as O
    from oracle.hparams import LJHPS
    texts, mels, t_len, m_len = O.synthetic_batch(LJHPS, 6, 12, 40)
    mine = P.shard_batch([texts, mels, t_len, m_len], rank, world)
    ids = P.shard_utterances(6, rank, world)
    g = torch.full((1000,), float(rank + 1))
    dist.all_reduce(g)                           # the single collective of the training path (sum; Adam folds 1 / world)
    g /= world
    tt = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)    # bench.py: a multi-GPU time is the max over ranks
    tmax = float(tt)
    fr = torch.tensor([int(mine[3].sum())])
    dist.all_reduce(fr)
    frames = int(fr)
    # overflow guard: only rank 1 sees a non-finite gradient; the all-reduced count makes BOTH ranks skip the step
    ovf = torch.tensor([1.0 if rank == 1 else 0.0])
    dist.all_reduce(ovf)
    assert float(ovf) == 1.0
    # replicas start from rank 0's parameters (host logic of VAENAR.broadcast_parameters, here over gloo on CPU tensors)
    import __graft_entry__ as ge
    ge.build()
    from vaenar_tts_b200 import VAENAR, LJHPS as PH
    m = VAENAR(PH, device="cpu", seed=100 + rank)
    before = m.flat_parameters().clone()
    m.broadcast_parameters(0)
    ref = m.flat_parameters().clone()
    dist.broadcast(ref, 0)
    same_as_rank0 = bool(torch.equal(m.flat_parameters(), ref))
    changed = bool(not torch.equal(before, m.flat_parameters()))
    out.put((rank, ids, [tuple(t.shape) for t in mine], float(g[0]), tmax, frames, int(m_len.sum()), same_as_rank0, changed))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_collectives():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = [r[1] for r in res]
    assert sorted(ids[0] + ids[1]) == list(range(6)) and not set(ids[0]) & set(ids[1])   # a partition
    assert ids[0] == [0, 2, 4] and ids[1] == [1, 3, 5]                                    # utt_ids[rank::size]
    for r in res:
        assert r[2][0][0] == 3 and r[2][1] == (3, 40, 80)
        assert abs(r[3] - 1.5) < 1e-6          # mean of the per-rank gradients 1 and 2
        assert r[4] == 2.0                      # max over ranks
        assert r[5] == r[6]                     # whole-job frames = sum over shards
        assert r[7]                             # every replica holds rank 0's parameters after the broadcast
    assert res[0][8] is False and res[1][8] is True   # rank 1 (different seed) was overwritten, rank 0 untouched

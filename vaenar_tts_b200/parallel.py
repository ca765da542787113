"""Data-parallel plumbing of the path (one process per GPU, torch.distributed).

The path shards by utterance: inference is pure replication with NO data-path collective;
training needs exactly one all-reduce of the flat gradient buffer.  The reference's only multi-worker vestige is
``TextMelData.get_batch(ids_file, rank, size)`` which strides utterance ids ``utt_ids[rank::size]``
(datasets/datasets.py:179-192); ``shard_utterances`` follows the same rule."""
import torch
import torch.distributed as dist


def shard_utterances(n_utterances: int, rank: int, world: int):
    """Indices of the utterances owned by ``rank`` (strided, datasets/datasets.py:191)."""
    return list(range(n_utterances))[rank::world]


def shard_batch(tensors, rank: int, world: int):
    """Slice every per-utterance tensor of a batch (leading dim = utterance) for this rank."""
    n = tensors[0].shape[0]
    idx = torch.as_tensor(shard_utterances(n, rank, world), dtype=torch.long)
    return [t.index_select(0, idx.to(t.device)) for t in tensors]


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of the flat gradient buffer (the single collective of the training path)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def max_over_ranks(value: float, device="cpu", group=None) -> float:
    """Timing reduction used by bench.py: a multi-GPU number is the MAX over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_frames(local_frames: int, device="cpu", group=None) -> int:
    """Whole-job number of mel frames processed in a step (sum over ranks)."""
    t = torch.tensor([int(local_frames)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())

"""Data-parallel plumbing of the path (one process per GPU, torch.distributed).

The path shards by utterance: inference is pure replication with NO data-path collective;
training needs exactly one all-reduce of the flat gradient buffer.  The reference's only multi-worker vestige is
``TextMelData.get_batch(ids_file, rank, size)`` which strides utterance ids ``utt_ids[rank::size]``
(datasets/datasets.py:179-192); ``shard_utterances`` follows the same rule."""
import torch


def shard_utterances(n_utterances: int, rank: int, world: int):
    """Indices of the utterances owned by ``rank`` (strided, datasets/datasets.py:191)."""
    return list(range(n_utterances))[rank::world]


def shard_batch(tensors, rank: int, world: int):
    """Slice every per-utterance tensor of a batch (leading dim = utterance) for this rank."""
    n = tensors[0].shape[0]
    idx = torch.as_tensor(shard_utterances(n, rank, world), dtype=torch.long)
    return [t.index_select(0, idx.to(t.device)) for t in tensors]

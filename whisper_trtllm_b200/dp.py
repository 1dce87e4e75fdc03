"""Data parallelism by utterance: one process per GPU, each rank holds a full weight replica and a
contiguous shard of the utterances; encode/decode need NO communication (the reference processes utterances one by
one, run.py:263-288, and oracle rows are independent).  The only collective is the final gather of the token ids
(``int32 [B/N, max_length]`` per rank, <= 458 KB in total) over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch


def world_info() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_utterances: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of rank; sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_utterances, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_shard(n_utterances: int, world_size: int) -> int:
    return -(-n_utterances // world_size)


def pad_tokens(ids: torch.Tensor, rows: int, max_length: int, pad_token_id: int) -> torch.Tensor:
    """ids [b, L<=max_length] -> int32 [rows, max_length], right/bottom padded with pad_token_id (rows finished early are
    already padded with it by the greedy loop, generation/utils.py:1506-1510)."""
    if ids.shape[1] > max_length or ids.shape[0] > rows:
        raise ValueError(f"ids {tuple(ids.shape)} do not fit the gather buffer [{rows}, {max_length}]: max_length must be the "
                         "longest sequence any rank can produce")
    out = torch.full((rows, max_length), pad_token_id, dtype=torch.int32, device=ids.device)
    out[:ids.shape[0], :ids.shape[1]] = ids.to(torch.int32)
    return out


def gather_tokens(ids: torch.Tensor, n_utterances: int, max_length: int, pad_token_id: int, group=None) -> torch.Tensor:
    """All-gather the per-rank token ids into the global ``[n_utterances, max_length]`` tensor (every rank gets it).
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return pad_tokens(ids, n_utterances, max_length, pad_token_id)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = max_shard(n_utterances, world)
    mine = pad_tokens(ids, rows, max_length, pad_token_id)
    gathered = torch.empty(world * rows, max_length, dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(n_utterances, world, r)
        parts.append(gathered[r * rows:r * rows + (e - b)])
    return torch.cat(parts, dim=0)


def transcribe_sharded(transcribe_fn: Callable[[torch.Tensor], torch.Tensor], mel_all: torch.Tensor, max_length: int,
                       pad_token_id: int, group=None, rank: Optional[int] = None, world_size: Optional[int] = None,
                       device: Optional[torch.device] = None) -> torch.Tensor:
    """Run ``transcribe_fn`` (mel shard -> ids) on this rank's shard of ``mel_all [N, 80, 3000]`` and gather.
    ``mel_all`` may live on the host: only the shard is moved by ``transcribe_fn``.  ``device`` is where the collective runs
    (the engine's / NCCL device): a rank with an EMPTY shard must enter the all-gather with a tensor on the same kind of
    device as the others.  Default: the device of the ids ``transcribe_fn`` returns, else of ``mel_all``."""
    import torch.distributed as dist
    if rank is None or world_size is None:
        if dist.is_available() and dist.is_initialized():
            rank, world_size = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world_size = 0, 1
    n = mel_all.shape[0]
    b, e = shard_range(n, world_size, rank)
    if e > b:
        ids = transcribe_fn(mel_all[b:e])
        if device is not None:
            ids = ids.to(device)
    else:
        ids = torch.empty(0, 1, dtype=torch.int32, device=device if device is not None else mel_all.device)
    return gather_tokens(ids, n, max_length, pad_token_id, group)

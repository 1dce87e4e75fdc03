"""Data-parallel sharding by utterance + the final token gather, world_size 2 over gloo on CPU.
The GPU path uses the same functions over NCCL (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_trtllm_b200 import dp


def test_shard_ranges_partition_every_count():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 31, 256, 257):
            spans = [dp.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == dp.max_shard(n, world) or n == 0
    assert dp.shard_range(256, 8, 3) == (96, 128)
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def _fake_transcribe(mel):
    """Deterministic stand-in for the per-rank engine: ids depend only on the utterance's own data (rows are
    independent), ragged lengths, pad after an 'EOS'."""
    n = mel.shape[0]
    L = 3 + int(mel[:, 0, 0].abs().sum().item()) % 4
    ids = torch.zeros(n, L, dtype=torch.int32)
    for i in range(n):
        ids[i] = (mel[i, 0, :L] * 1000).to(torch.int32)
    return ids


def _worker(rank, world, port, n, max_length, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        mel = torch.randn(n, 2, 16, generator=g)
        got = dp.transcribe_sharded(_fake_transcribe, mel, max_length, pad_token_id=50256)
        torch.save(got, os.path.join(out_dir, f"rank{rank}.pt"))
        assert dp.world_info() == (rank, world, rank)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [5, 8, 1])
def test_sharded_transcription_equals_single_process(n, tmp_path):
    world, max_length = 2, 9
    mp.spawn(_worker, args=(world, _free_port(), n, max_length, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(7)
    mel = torch.randn(n, 2, 16, generator=g)
    # reference: every utterance on its own (rows are independent), padded to max_length
    want = torch.full((n, max_length), 50256, dtype=torch.int32)
    for r in range(world):
        b, e = dp.shard_range(n, world, r)
        if e > b:
            ids = _fake_transcribe(mel[b:e])
            want[b:e, :ids.shape[1]] = ids
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert got.shape == (n, max_length) and torch.equal(got, want), f"rank {r}"


def test_gather_without_process_group_is_identity_padding():
    ids = torch.arange(6, dtype=torch.int32).reshape(2, 3)
    out = dp.gather_tokens(ids, 2, 5, 9)
    assert out.tolist() == [[0, 1, 2, 9, 9], [3, 4, 5, 9, 9]]

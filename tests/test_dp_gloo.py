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


# ---------------------------------------------------------------------------------------------------- the pipeline's sharding
def _pipeline_worker(rank, world, port, n, out_dir):
    """WhisperPipeline.transcribe_sharded under a process group, the engine replaced by a stub: each rank loads and transcribes
    only its shard (with the prefetching loader), every rank ends with the ids of ALL utterances in input order."""
    import types

    import numpy as np
    from whisper_trtllm_b200 import pipeline
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe = object.__new__(pipeline.WhisperPipeline)
        pipe.config = {"max_length": 7, "pad_token_id": 50256}
        pipe.max_batch = 2
        pipe.engine = types.SimpleNamespace(device=torch.device("cpu"))
        loaded = []

        def load(item):
            loaded.append(item)
            return np.full(3, item, dtype=np.float32)

        pipe.transcribe_waveforms = lambda waves: torch.tensor([[int(w[0]), 1, 2, 50256, 50256, 50256, 50256] for w in waves],
                                                               dtype=torch.int32).reshape(len(waves), 7)
        got = pipe.transcribe_sharded(list(range(100, 100 + n)), load=load)
        b, e = dp.shard_range(n, world, rank)
        assert sorted(loaded) == list(range(100 + b, 100 + e)), (rank, loaded)       # nothing outside the shard was read
        feats = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1).expand(n, 80, 3000)
        pipe.transcribe_features = lambda m: torch.stack([torch.full((7,), int(x[0, 0]), dtype=torch.int32) for x in m]) \
            if len(m) else torch.empty(0, 7, dtype=torch.int32)
        got_f = pipe.transcribe_sharded(feats, features=True)
        torch.save((got, got_f), os.path.join(out_dir, f"pipe_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 1])
def test_pipeline_sharding_world_2(n, tmp_path):
    world = 2
    mp.spawn(_pipeline_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got, got_f = torch.load(os.path.join(str(tmp_path), f"pipe_rank{r}.pt"))
        assert got.shape == (n, 7) and got[:, 0].tolist() == list(range(100, 100 + n)) and got[:, 3:].eq(50256).all()
        assert got_f.shape == (n, 7) and got_f[:, 0].tolist() == list(range(n))

"""Waveforms in, text (and WER) out: the main loops of the reference's two scripts, batched.

``run.py:259-290`` walks a dataset one utterance at a time — feature extractor on the CPU, encoder engine, greedy loop with a
host round trip per token, ``batch_decode`` — and ``cal_wer.py:251-287`` does the same over (mel, text) pairs and then
normalises both sides and calls ``jiwer.wer``.  ``WhisperPipeline`` is that loop over batches of utterances with every stage
on the device: ``LogMelFrontend`` -> ``WhisperEngine.generate`` (the hot path, §8a) -> ``WhisperDetokenizer``
-> ``EnglishTextNormalizer`` / ``wer`` (row 3), from a checkpoint directory read by ``checkpoint.load_hf_checkpoint`` (row 1).
Under ``torchrun`` the utterances are sharded over the ranks and the ids gathered at the end (``dp``).

Host plumbing only: no arithmetic happens here, and there is no CPU fallback (the engine and the front-end raise without a GPU).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Sequence

import torch

from . import dp
from .audio import batches, load_audio
from .checkpoint import load_hf_checkpoint
from .text import EnglishTextNormalizer, WhisperDetokenizer, wer


def first_special_token_id(checkpoint_dir: str, config: Dict) -> int:
    """Smallest special-token id.  Whisper's special tokens are one block at the top of the vocabulary: ``<|endoftext|>``
    (= eos_token_id: 50256 in the `.en` vocabularies, where it is part of ``vocab.json``; 50257 in the multilingual ones) followed
    by the tokens of ``added_tokens.json``."""
    first = int(config["eos_token_id"])
    p = os.path.join(checkpoint_dir, "added_tokens.json")
    if os.path.exists(p):
        with open(p, encoding="utf-8") as f:
            added = json.load(f)
        if added:
            first = min(first, min(int(v) for v in added.values()))
    return first


def load_text_tools(checkpoint_dir: str, config: Dict):
    """-> (detokenizer or None, normaliser).  ``vocab.json`` and ``normalizer.json`` are the tokenizer files that ship with
    every Whisper checkpoint; without ``vocab.json`` only token ids can be returned, without ``normalizer.json`` the
    normaliser runs with an empty spelling table."""
    def read_json(name):
        p = os.path.join(checkpoint_dir, name)
        if not os.path.exists(p):
            return {}
        with open(p, encoding="utf-8") as f:
            return json.load(f)
    vocab = os.path.join(checkpoint_dir, "vocab.json")
    detok = None
    if os.path.exists(vocab):
        clean_up = read_json("tokenizer_config.json").get("clean_up_tokenization_spaces", True)   # tokenization_utils_base.py:1552
        detok = WhisperDetokenizer(vocab, first_special_token_id(checkpoint_dir, config), clean_up_tokenization_spaces=clean_up,
                                   added_tokens=read_json("added_tokens.json"))
    spelling = {}
    p = os.path.join(checkpoint_dir, "normalizer.json")
    if os.path.exists(p):
        with open(p, encoding="utf-8") as f:
            spelling = json.load(f)
    return detok, EnglishTextNormalizer(spelling)


class WhisperPipeline:
    def __init__(self, checkpoint_dir: str, dtype: str = "bfloat16", max_batch: int = 64, device=None, compact_every: int = 32,
                 refill: bool = True):
        """``dtype``: "bfloat16" (tcgen05 speed path) or "float32" (token ids identical to the reference's fp32 run).
        ``compact_every``: every that many tokens the utterances that have emitted EOS leave the decode batch (0 = never).
        ``refill``: when more utterances are queued than the engine has rows, the freed rows are refilled in flight with the
        next utterances (WhisperEngine.transcribe_stream) instead of decoding batch after batch."""
        from .engine import WhisperEngine
        from .frontend import LogMelFrontend
        self.checkpoint_dir = checkpoint_dir
        self.config, state_dict = load_hf_checkpoint(checkpoint_dir)
        self.detokenizer, self.normalizer = load_text_tools(checkpoint_dir, self.config)
        self.max_batch = int(max_batch)
        self.compact_every = int(compact_every)
        self.refill = bool(refill)
        self.engine = WhisperEngine(self.config, state_dict, dtype=dtype, max_batch=self.max_batch, device=device)
        self.frontend = LogMelFrontend(self.engine.device)

    # ------------------------------------------------------------------ ids
    @torch.no_grad()
    def transcribe_features(self, input_features: torch.Tensor) -> torch.Tensor:
        """log-mel fp32 [n, 80, 3000] (host or device) -> ids int32 [n, max_length] on the host, padded with pad_token_id."""
        L, pad = self.config["max_length"], self.config["pad_token_id"]
        n = input_features.shape[0]
        if getattr(self, "refill", False) and n > self.max_batch and hasattr(self.engine, "transcribe_stream"):
            ids = self.engine.transcribe_stream(input_features.to(torch.float32), window=self.compact_every or 32)
            return dp.pad_tokens(ids, n, L, pad).cpu()
        rows = []
        for b0 in range(0, input_features.shape[0], self.max_batch):
            mel = input_features[b0:b0 + self.max_batch].to(self.engine.device, torch.float32).contiguous()
            ids = self.engine.generate(mel, compact_every=self.compact_every)
            rows.append(dp.pad_tokens(ids, ids.shape[0], L, pad).cpu())
        return torch.cat(rows, dim=0) if rows else torch.empty(0, L, dtype=torch.int32)

    @torch.no_grad()
    def transcribe_waveforms(self, waves: Sequence) -> torch.Tensor:
        """16 kHz waveforms (any lengths; padded / cut to 30 s) -> ids int32 [n, max_length] on the host."""
        L, pad = self.config["max_length"], self.config["pad_token_id"]
        waves = list(waves)
        if getattr(self, "refill", False) and len(waves) > self.max_batch and hasattr(self.engine, "transcribe_stream"):
            # log-mel of everything first (GPU front-end, batch by batch), then ONE refilled greedy loop over all utterances
            mel = torch.cat([self.frontend(chunk) for chunk in batches(waves, self.max_batch)], dim=0)
            ids = self.engine.transcribe_stream(mel, window=self.compact_every or 32)
            return dp.pad_tokens(ids, len(waves), L, pad).cpu()
        rows = []
        for chunk in batches(list(waves), self.max_batch):
            ids = self.engine.generate(self.frontend(chunk), compact_every=self.compact_every)
            rows.append(dp.pad_tokens(ids, ids.shape[0], L, pad).cpu())
        return torch.cat(rows, dim=0) if rows else torch.empty(0, L, dtype=torch.int32)

    def transcribe_sharded(self, items, load=None, features: bool = False) -> torch.Tensor:
        """ids of ALL items on every rank: each rank transcribes its contiguous shard (dp.shard_range), one final gather.
        ``load(item) -> waveform`` is applied on a thread pool (file reading and FLAC decoding release the GIL).
        ``features``: the items are log-mel windows [n, 80, 3000] (cal_wer.py's librispeech.cache), not waveforms."""
        import torch.distributed as dist
        L, pad = self.config["max_length"], self.config["pad_token_id"]
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        b, e = dp.shard_range(len(items), world, rank)
        if features:
            ids = self.transcribe_features(items[b:e])
            return ids if world == 1 else dp.gather_tokens(ids.to(self.engine.device), len(items), L, pad).cpu()
        mine = list(items[b:e])
        if load is None or not mine:
            ids = self.transcribe_waveforms(mine)
        else:
            # software pipeline: while the GPU works on batch k, the pool reads and decodes the files of batch k + 1
            from concurrent.futures import ThreadPoolExecutor
            chunks = list(batches(mine, self.max_batch))
            rows = []
            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
                pending = [pool.submit(load, it) for it in chunks[0]]
                for k in range(len(chunks)):
                    waves = [f.result() for f in pending]
                    pending = [pool.submit(load, it) for it in chunks[k + 1]] if k + 1 < len(chunks) else []
                    rows.append(self.transcribe_waveforms(waves))
            ids = torch.cat(rows, dim=0)
        if world == 1:
            return ids
        return dp.gather_tokens(ids.to(self.engine.device), len(items), L, pad).cpu()

    def transcribe_files(self, paths: Sequence[str]) -> torch.Tensor:
        """Audio files (.wav / .flac / .npy, 16 kHz) -> ids int32 [n, max_length] of all files, on every rank."""
        return self.transcribe_sharded(paths, load=load_audio)

    # ------------------------------------------------------------------ text
    def decode(self, ids) -> List[str]:
        """``hf_processor.batch_decode(predicted_ids, skip_special_tokens=True)`` (run.py:287)."""
        if self.detokenizer is None:
            raise FileNotFoundError(f"{self.checkpoint_dir} has no vocab.json: token ids cannot be turned into text")
        return self.detokenizer.batch_decode(ids, skip_special_tokens=True)

    def __call__(self, waves: Sequence) -> List[str]:
        return self.decode(self.transcribe_waveforms(waves))

    def wer(self, hypotheses: Sequence[str], references: Sequence[str]) -> float:
        """Both sides through the English normaliser, then the corpus word error rate (cal_wer.py:279-286)."""
        return wer([self.normalizer(t) for t in references], [self.normalizer(t) for t in hypotheses])

    def close(self):
        self.engine.close()


def compare_transcriptions(ours: Sequence[str], theirs: Sequence[str]) -> List[tuple]:
    """The `--compare` report of run.py:321-331: the (ours, theirs) pairs that differ."""
    if len(ours) != len(theirs):
        raise ValueError("the two runs transcribed a different number of utterances")
    return [(a, b) for a, b in zip(ours, theirs) if a != b]

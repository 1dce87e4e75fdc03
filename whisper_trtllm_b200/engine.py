"""WhisperEngine — packed weights + a session of libwhisper_b200 (the native runtime).

Takes the place of the reference's two serialized TensorRT engines plus ``config.pkl``
(examples/whisper/build_encoder.py:30-109, build_decoder.py:30-119, run.py:250-256): weights come
straight from an HF ``state_dict`` (same keys the reference's binders read) and are packed once on the
device; nothing is traced or compiled.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _abi
from ._abi import byref, c_int, c_int32, c_int64, c_size_t, c_void_p, ptr, stream_handle

_DTYPES = {"float32": _abi.F32, "fp32": _abi.F32, "bfloat16": _abi.BF16, "bf16": _abi.BF16,
           torch.float32: _abi.F32, torch.bfloat16: _abi.BF16}
_TORCH_DTYPE = {_abi.F32: torch.float32, _abi.BF16: torch.bfloat16}


def begin_index_of(config: Dict, input_ids_seq_length: int = 1) -> int:
    """begin_index of SuppressTokensAtBeginLogitsProcessor exactly as run.py:155-158 computes it."""
    b = input_ids_seq_length
    if config.get("forced_bos_token_id") is not None:
        b += 1
    return b + config["forced_decoder_ids"][-1][0]


class WhisperEngine:
    def __init__(self, config: Dict, state_dict: Dict[str, torch.Tensor], dtype="float32", max_batch: int = 1,
                 enc_chunk: Optional[int] = None, device: Optional[torch.device] = None, n_streams: int = 1):
        """``n_streams`` > 1 splits the batch into that many sub-batches, each with its own native session (KV caches,
        token loop) on its own stream; their greedy loops are interleaved (wb_decode_run_multi) so that one sub-batch's
        HBM-bound cross-attention overlaps the other's latency-bound kernels.  Rows are independent, so the ids are the same."""
        if not torch.cuda.is_available():
            raise _abi.WhisperB200Error(-101, "a CUDA device is required (no CPU fallback)")
        self.lib = _abi.load()
        self.config = dict(config)
        self.dtype_code = _DTYPES[dtype]
        self.torch_dtype = _TORCH_DTYPE[self.dtype_code]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_batch = int(max_batch)
        self.n_streams = max(1, min(int(n_streams), self.max_batch))
        self.sub_batch = -(-self.max_batch // self.n_streams)          # rows per sub-session (last one may be smaller)
        self.enc_chunk = int(min(enc_chunk or 32, self.sub_batch))
        c = config
        assert c["encoder_attention_heads"] == c["decoder_attention_heads"] and c["encoder_ffn_dim"] == c["decoder_ffn_dim"]
        self._cfg = _abi.wb_config(
            d_model=c["d_model"], n_heads=c["encoder_attention_heads"], encoder_layers=c["encoder_layers"],
            decoder_layers=c["decoder_layers"], ffn_dim=c["encoder_ffn_dim"], vocab_size=c["vocab_size"],
            num_mel_bins=c["num_mel_bins"], n_frames=2 * c["max_source_positions"],
            max_source_positions=c["max_source_positions"], max_target_positions=c["max_target_positions"],
            decoder_start_token_id=c["decoder_start_token_id"], eos_token_id=c["eos_token_id"],
            pad_token_id=c["pad_token_id"], max_length=c["max_length"])
        self._model = c_void_p()
        self._session = c_void_p()
        with torch.cuda.device(self.device):
            _abi.call("wb_model_create", byref(self._cfg), self.dtype_code, byref(self._model))
            self._load_state_dict(state_dict)
            self._set_generation()
            nbytes = c_size_t()
            _abi.call("wb_session_workspace_bytes", self._model, self.sub_batch, self.enc_chunk, byref(nbytes))
            self._subs = []   # (session handle, workspace tensor)
            for _ in range(self.n_streams):
                ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
                h = c_void_p()
                _abi.call("wb_session_create", self._model, self.sub_batch, self.enc_chunk, ptr(ws), c_size_t(nbytes.value), byref(h))
                self._subs.append((h, ws))
            self._session, self.workspace = self._subs[0]
            if self.n_streams > 1:
                # co-residency of the two streams' kernels on one SM: bulk-ring cross-attention (1 CTA/SM, bytes in flight
                # in shared memory) + the lean decode GEMM (csrc/attn_dec_bulk.cu, gemm_tc.cu kLean)
                _abi.call("wb_set_decode_attention_backend", 1)
                _abi.call("wb_set_lean_decode_gemm", 1)
        self.d_model = c["d_model"]
        self.n_ctx = c["max_source_positions"]
        self.vocab = c["vocab_size"]
        self.max_tgt = c["max_target_positions"]
        self._keepalive = []

    # ------------------------------------------------------------------ weights
    def _load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """HF state_dict -> packed device weights (mapping of build_encoder.py:71-91 / build_decoder.py:71-101,
        done natively in csrc/runtime.cu Model::load_tensor)."""
        for name, t in sd.items():
            if name == "proj_out.weight":
                emb = sd.get("model.decoder.embed_tokens.weight")
                if emb is not None and t.data_ptr() != emb.data_ptr() and not torch.equal(t, emb):
                    raise ValueError("proj_out.weight is not tied to embed_tokens (modeling_whisper.py:1335)")
                continue
            h = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            _abi.call("wb_model_load_tensor", self._model, name.encode(), c_void_p(h.data_ptr()), c_int64(h.numel()))

    def _set_generation(self):
        c = self.config
        self.set_generation(c["suppress_tokens"], c["begin_suppress_tokens"], begin_index_of(c), c["forced_decoder_ids"])

    def set_generation(self, suppress_tokens, begin_suppress_tokens, begin_index: int, forced_decoder_ids):
        """The three logits processors of run.py:150-162 as data: suppress list, begin-suppress list + begin_index,
        forced (generation index, token) pairs.  Cheap (two small H2D copies); skipped when nothing changed."""
        key = (tuple(suppress_tokens), tuple(begin_suppress_tokens), int(begin_index), tuple(tuple(p) for p in forced_decoder_ids))
        if getattr(self, "_generation_key", None) == key:
            return
        sup = (c_int32 * len(key[0]))(*key[0])
        beg = (c_int32 * len(key[1]))(*key[1])
        flat = [int(v) for pair in key[3] for v in pair]
        forced = (c_int32 * len(flat))(*flat)
        torch.cuda.current_stream().synchronize()   # the tables may still be read by queued steps
        _abi.call("wb_model_set_generation", self._model, sup, len(key[0]), beg, len(key[1]), int(begin_index), forced, len(key[3]))
        self._generation_key = key

    def set_option(self, name: str, value: int):
        """Per-session override of a process-wide A/B switch (wb_session_set_option): "small_batch_path", "decode_chain_path",
        "cuda_graphs"; 1 on, 0 off, -1 inherit."""
        for h, _ in self._subs:
            _abi.call("wb_session_set_option", h, name.encode(), int(value))

    def weight_bytes(self) -> int:
        n = c_size_t()
        _abi.call("wb_model_weight_bytes", self._model, byref(n))
        return n.value

    # ------------------------------------------------------------------ encoder
    def encode(self, mel: torch.Tensor, return_hidden: bool = True, stream=None) -> Optional[torch.Tensor]:
        """mel fp32 [B, 80, 3000] on the device -> hidden_states fp32 [B, 1500, d]; fills the cross K/V caches."""
        assert mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous(), "mel must be a contiguous fp32 CUDA tensor"
        B = mel.shape[0]
        out = torch.empty(B, self.n_ctx, self.d_model, dtype=torch.float32, device=mel.device) if return_hidden else None
        for k, (b0, b1) in enumerate(self._shards(B)):
            _abi.call("wb_encode", self._subs[k][0], ptr(mel[b0:b1]), b1 - b0, ptr(out[b0:b1]) if out is not None else None,
                      stream_handle(stream))
        return out

    def _shards(self, B: int):
        """Contiguous row ranges of the sub-sessions for a batch of B rows (empty ranges dropped)."""
        assert 0 < B <= self.max_batch, f"batch {B} exceeds max_batch {self.max_batch}"
        return [(b0, min(b0 + self.sub_batch, B)) for b0 in range(0, B, self.sub_batch)]

    def _single(self, what: str):
        if self.n_streams != 1:
            raise NotImplementedError(f"{what} is only available with n_streams == 1")

    def set_encoder_output(self, enc: torch.Tensor, stream=None):
        assert enc.is_cuda and enc.is_contiguous() and enc.dtype in (torch.float32, torch.bfloat16)
        for k, (b0, b1) in enumerate(self._shards(enc.shape[0])):
            _abi.call("wb_set_encoder_output", self._subs[k][0], ptr(enc[b0:b1]), _DTYPES[enc.dtype], b1 - b0, stream_handle(stream))

    # ------------------------------------------------------------------ greedy decode
    def decode_begin(self, batch: int, stream=None):
        if not 0 < batch <= self.max_batch:
            raise _abi.WhisperB200Error(-1, f"decode batch {batch} outside (0, {self.max_batch}]")
        self._active = self._shards(batch)
        for k, (b0, b1) in enumerate(self._active):
            _abi.call("wb_decode_begin", self._subs[k][0], b1 - b0, stream_handle(stream))

    def decode_step(self, stream=None):
        for k in range(len(getattr(self, "_active", [None]))):
            _abi.call("wb_decode_step", self._subs[k][0], stream_handle(stream))

    def decode_run(self, max_steps: int = 0, check_every: int = 32, stream=None) -> int:
        """Runs the greedy loop(s) to the end; returns the final sequence length (the longest over the sub-batches)."""
        n = len(getattr(self, "_active", [None]))
        handles = (c_void_p * n)(*[self._subs[k][0] for k in range(n)])
        lens = (c_int * n)()
        _abi.call("wb_decode_run_multi", handles, n, max_steps, check_every, lens, stream_handle(stream))
        self._final_lens = list(lens)
        return max(self._final_lens)

    def decode_compact(self, stream=None) -> int:
        """Finished-row compaction (wb_decode_compact): utterances that have emitted EOS leave the decode batch,
        the following steps run on the rows still decoding.  Returns their number (0: the loop has stopped)."""
        self._single("decode_compact()")
        n = c_int()
        _abi.call("wb_decode_compact", self._session, byref(n), stream_handle(stream))
        return n.value

    def _view(self, address: int, shape, dtype) -> torch.Tensor:
        """Zero-copy torch view of session-owned device memory (lives inside sub-session 0's workspace)."""
        return self._view_of(self.workspace, address, shape, dtype)

    def _view_of(self, ws: torch.Tensor, address: int, shape, dtype) -> torch.Tensor:
        off = address - ws.data_ptr()
        n = 1
        for s_ in shape:
            n *= s_
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        assert 0 <= off and off + nbytes <= ws.numel()
        return ws[off:off + nbytes].view(dtype).view(*shape)

    def tokens(self) -> torch.Tensor:
        """ids int32 [rows, max_target_positions]: a view into the session (n_streams == 1) or the sub-sessions' rows
        concatenated (copy)."""
        parts = []
        for k, (h, ws) in enumerate(self._subs):
            p, stride = c_void_p(), c_int()
            _abi.call("wb_decode_tokens", h, byref(p), byref(stride))
            parts.append(self._view_of(ws, p.value, (self.sub_batch, stride.value), torch.int32))
        if self.n_streams == 1:
            return parts[0]
        shards = getattr(self, "_active", self._shards(self.max_batch))
        return torch.cat([parts[k][:b1 - b0] for k, (b0, b1) in enumerate(shards)], dim=0)

    def logits(self) -> torch.Tensor:
        self._single("logits()")
        p = c_void_p()
        _abi.call("wb_decode_logits", self._session, byref(p))
        return self._view(p.value, (self.max_batch, self.vocab), torch.float32)

    def cross_kv(self, layer: int) -> torch.Tensor:
        """[2, max_batch, H, 1500, 64] view of the cross-attention cache of one decoder layer."""
        self._single("cross_kv()")
        p, stride = c_void_p(), c_int64()
        _abi.call("wb_session_cross_kv", self._session, layer, byref(p), byref(stride))
        H = self.config["decoder_attention_heads"]
        return self._view(p.value, (2, self.max_batch, H, self.n_ctx, 64), self.torch_dtype)

    def self_kv(self, layer: int, batch: int, length: int):
        """Gather the paged self-attention cache of one layer into dense [B, H, length, 64] K and V."""
        self._single("self_kv()")
        kp, vp, pt, pps, ptok = c_void_p(), c_void_p(), c_void_p(), c_int(), c_int()
        _abi.call("wb_session_self_kv", self._session, layer, byref(kp), byref(vp), byref(pt), byref(pps), byref(ptok))
        H = self.config["decoder_attention_heads"]
        num_pages = self.max_batch * pps.value
        k = self._view(kp.value, (num_pages, H, ptok.value, 64), self.torch_dtype)
        v = self._view(vp.value, (num_pages, H, ptok.value, 64), self.torch_dtype)
        table = self._view(pt.value, (self.max_batch, pps.value), torch.int32)[:batch].long()
        def dense(pages):
            g = pages[table]                                  # [B, pps, H, tok, 64]
            g = g.permute(0, 2, 1, 3, 4).reshape(batch, H, pps.value * ptok.value, 64)
            return g[:, :, :length].contiguous()
        return dense(k), dense(v)

    @torch.no_grad()
    def generate(self, mel: torch.Tensor, max_new_tokens: Optional[int] = None, forced_tokens: Optional[torch.Tensor] = None,
                 dump_logits_steps: int = 0, check_every: int = 32, stream=None, compact_every: int = 0):
        """Encoder + on-device greedy loop.  Returns ids int32 [B, L] (L as the oracle's loop would stop),
        plus per-step raw logits [steps, B, V] when ``dump_logits_steps`` > 0.  ``compact_every`` = n > 0: every n steps the
        utterances that have emitted EOS are removed from the decode batch (same ids, fewer rows per step)."""
        B = mel.shape[0]
        self.encode(mel, return_hidden=False, stream=stream)
        return self.greedy(B, max_new_tokens, forced_tokens, dump_logits_steps, check_every, stream, compact_every)

    @torch.no_grad()
    def greedy(self, B: int, max_new_tokens=None, forced_tokens=None, dump_logits_steps: int = 0, check_every: int = 32,
               stream=None, compact_every: int = 0):
        dump = None
        if forced_tokens is not None or dump_logits_steps > 0:
            self._single("teacher forcing / logits dump")
        if forced_tokens is not None:
            ft = torch.full((B, self.max_tgt), self.config["pad_token_id"], dtype=torch.int32, device=self.device)
            ft[:, :forced_tokens.shape[1]] = forced_tokens.to(device=self.device, dtype=torch.int32)
            self._keepalive = [ft]
            _abi.call("wb_decode_set_forced_tokens", self._session, ptr(ft))
        else:
            _abi.call("wb_decode_set_forced_tokens", self._session, None)
        if dump_logits_steps > 0:
            dump = torch.empty(dump_logits_steps, B, self.vocab, dtype=torch.float32, device=self.device)
        _abi.call("wb_decode_set_logits_dump", self._session, ptr(dump), dump_logits_steps)
        self.decode_begin(B, stream)
        steps = 0 if max_new_tokens is None else int(max_new_tokens)
        if compact_every > 0 and forced_tokens is None and dump_logits_steps == 0 and self.n_streams == 1:
            total = steps if steps > 0 else self.config["max_length"] - 1
            done, final_len = 0, 1
            while done < total:
                n = min(int(compact_every), total - done)
                final_len = self.decode_run(n, check_every, stream)
                done += n
                if done >= total or self.decode_compact(stream) == 0:
                    break
        else:
            final_len = self.decode_run(steps, check_every, stream)
        ids = self.tokens()[:B, :final_len].clone()
        if self.n_streams > 1:   # a sub-batch whose rows all hit EOS stopped earlier: its tail is pad (GU:1506-1510)
            for (b0, b1), n in zip(self._active, self._final_lens):
                ids[b0:b1, n:] = self.config["pad_token_id"]
        _abi.call("wb_decode_set_logits_dump", self._session, None, 0)
        _abi.call("wb_decode_set_forced_tokens", self._session, None)
        if dump is not None:
            return ids, dump[:final_len - 1]
        return ids

    @torch.no_grad()
    def transcribe_stream(self, mel_all: torch.Tensor, window: int = 32, stream=None) -> torch.Tensor:
        """Greedy transcription of ANY number of utterances with in-flight refill (wb_decode_refill): the engine
        decodes `max_batch` rows at a time; every `window` steps the utterances that have emitted EOS leave the batch and the next
        ones of `mel_all` (fp32 [N, 80, 3000], host or device) are encoded into the freed slots and start decoding while the
        others continue, so the batch stays full until the queue is empty.  Returns int32 [N, L] (L = longest result, rows
        padded with pad_token_id exactly like the reference's batched loop pads finished rows, generation/utils.py:1506-1510)."""
        import numpy as np
        self._single("transcribe_stream()")
        N = mel_all.shape[0]
        assert N > 0 and mel_all.dtype == torch.float32
        cfg = self.config
        B0 = min(N, self.max_batch)

        def dev(lo, hi):
            return mel_all[lo:hi].to(self.device, non_blocking=True).contiguous()

        _abi.call("wb_decode_set_forced_tokens", self._session, None)
        _abi.call("wb_decode_set_logits_dump", self._session, None, 0)
        self.encode(dev(0, B0), return_hidden=False, stream=stream)
        self.decode_begin(B0, stream)
        nxt = B0
        fin_utt = np.zeros(self.max_batch, dtype=np.int32)
        fin_len = np.zeros(self.max_batch, dtype=np.int32)
        fin_ids = np.zeros((self.max_batch, self.max_tgt), dtype=np.int32)
        results = [None] * N
        rows = B0
        while rows > 0:
            self.decode_run(int(window), check_every=int(window), stream=stream)
            want = min(self.max_batch, N - nxt)
            mel_new = dev(nxt, nxt + want) if want > 0 else None
            n_fin, n_adm, n_rows = c_int(), c_int(), c_int()
            _abi.call("wb_decode_refill", self._session, ptr(mel_new), want, c_void_p(fin_utt.ctypes.data), c_void_p(fin_len.ctypes.data),
                      c_void_p(fin_ids.ctypes.data), byref(n_fin), byref(n_adm), byref(n_rows), stream_handle(stream))
            for k in range(n_fin.value):
                results[int(fin_utt[k])] = fin_ids[k, :int(fin_len[k])].copy()
            nxt += n_adm.value
            rows = n_rows.value
            self._active = self._shards(max(rows, 1))
        assert all(r is not None for r in results), "an utterance never finished"
        L = max(len(r) for r in results)
        out = torch.full((N, L), cfg["pad_token_id"], dtype=torch.int32)
        for i, r in enumerate(results):
            out[i, :len(r)] = torch.from_numpy(r)
        return out

    @torch.no_grad()
    def transcribe_host(self, mel_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """End-to-end call with HOST buffers: H2D copy of the log-mel, encoder, greedy loop, D2H copy of the ids."""
        B = mel_host.shape[0]
        mel = mel_host.to(self.device, non_blocking=True)
        ids = self.generate(mel, stream=stream)
        if out_host is None:
            return ids.cpu()
        out_host[:B, :ids.shape[1]].copy_(ids, non_blocking=False)
        return out_host[:B, :ids.shape[1]]

    # ------------------------------------------------------------------ live kernel timing (bench roofline)
    PROF_CLASSES = {"cross_attn": 1, "self_attn": 2, "dec_gemm": 3, "lm_head": 4, "enc_gemm": 5, "enc_attn": 6,
                    "layernorm": 7, "greedy": 8, "stem": 9, "cross_kv": 10}

    def profile(self, kernel_class: Optional[str], decode_step: int = -1):
        """Record CUDA events around every launch of one kernel class inside the real loop (None = off).
        ``decode_step >= 0``: decode kernels are timed only in that step of each greedy loop (it runs eagerly; the other
        steps keep replaying the CUDA graph)."""
        _abi.call("wb_session_profile_at", self._session, 0 if kernel_class is None else self.PROF_CLASSES[kernel_class],
                  int(decode_step))

    def profile_read(self):
        """-> (summed device milliseconds, launches) since profile() was armed; synchronises."""
        import ctypes
        ms, n = ctypes.c_double(), ctypes.c_longlong()
        _abi.call("wb_session_profile_read", self._session, byref(ms), byref(n))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.wb_launch_count())

    def close(self):
        for h, _ in getattr(self, "_subs", []):
            if h:
                _abi.call("wb_session_destroy", h)
        self._subs = []
        self._session = c_void_p()
        if self._model:
            _abi.call("wb_model_destroy", self._model)
            self._model = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""whisper_trtllm_b200 — B200-native (sm_100a) Whisper greedy inference behind the drop-in surface of
EdVince/whisper-trtllm (WhisperEncoder / WhisperDecoder / WhisperDecoderAttention, the run.py greedy session).

Python here is host plumbing only; all arithmetic runs in hand-written CUDA kernels reached through the
C-ABI of ``libwhisper_b200.so`` (include/whisper_b200.h).  There is no CPU or library fallback.
"""
from ._abi import BF16, F32, WhisperB200Error  # noqa: F401
from .engine import WhisperEngine, begin_index_of  # noqa: F401

__version__ = "0.1.0"

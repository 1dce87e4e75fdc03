"""whisper_trtllm_b200 — B200-native (sm_100a) Whisper greedy inference behind the drop-in surface of
EdVince/whisper-trtllm (WhisperEncoder / WhisperDecoder / WhisperDecoderAttention, the run.py greedy session).

Python here is host plumbing only; all arithmetic runs in hand-written CUDA kernels reached through the
C-ABI of ``libwhisper_b200.so`` (include/whisper_b200.h).  There is no CPU or library fallback.

    whisper_trtllm_b200.models.WhisperEncoder / WhisperDecoder / WhisperDecoderAttention / WhisperDecoderLayer
                                  module-level drop-ins of tensorrt_llm.models (model.py)
    whisper_trtllm_b200.runtime.Session / TensorInfo          drop-ins of tensorrt_llm.runtime (session.py)
    whisper_trtllm_b200.run      runner classes + greedy_search / get_logits_processor / get_stopping_criteria (run.py)
    whisper_trtllm_b200.WhisperEngine   the native runtime (packed weights, paged KV, on-device greedy loop)
    whisper_trtllm_b200.dp       data-parallel sharding by utterance + the final token gather
    whisper_trtllm_b200.checkpoint / .frontend / .processor / .text / .audio / .pipeline
                                  either side of the path: HF checkpoint directory -> weights, PCM -> log-mel on the GPU,
                                  ids -> text -> WER, and the scripts' main loops (WhisperPipeline)
"""
from ._abi import BF16, F32, WhisperB200Error  # noqa: F401
from .engine import WhisperEngine, begin_index_of  # noqa: F401
from . import model as models  # noqa: F401  (tensorrt_llm.models.WhisperEncoder -> whisper_trtllm_b200.models.WhisperEncoder)
from . import session as runtime  # noqa: F401  (tensorrt_llm.runtime.Session / TensorInfo)
from . import layers, run, dp  # noqa: F401
from . import audio, checkpoint, text  # noqa: F401  (host-only modules; frontend / pipeline import on use)

__version__ = "0.2.0"

"""Synthetic configs / weights / inputs: re-exported from whisper_trtllm_b200.synthetic (pure data generation shared by the
tests, the oracle's golden scripts and bench.py; the oracle may import the package, never the other way round)."""
from whisper_trtllm_b200.synthetic import (NON_SPEECH_TOKENS, SIZES, make_config, make_mel, make_weights,  # noqa: F401
                                           weights_fingerprint, _uniform)

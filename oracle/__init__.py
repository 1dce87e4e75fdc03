"""CPU oracle for the Whisper greedy-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain fp32 PyTorch on the CPU, the algorithm of the reference's own
implementation of the path (the hand-simplified HuggingFace Whisper vendored by EdVince/whisper-trtllm
under ``transformers/src/transformers/models/whisper/modeling_whisper.py`` and
``generation/utils.py``).  Every function cites the reference file:line it follows.

Nothing under ``whisper_trtllm_b200/`` (the product) may import from here.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it,
and only as the checker / the CPU baseline, never as the thing that is shipped or measured as ours.

Parity pinning: the restatement is checked bit-for-token (and to 1e-5 on logits/encoder output)
against the *real* reference imported from ``/root/reference`` in the build container
(``oracle/hf_reference.py`` + ``oracle/make_golden.py``); the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py`` everywhere.
"""

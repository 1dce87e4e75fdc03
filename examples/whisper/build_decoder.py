#!/usr/bin/env python
"""`python build_decoder.py --whisper <checkpoint dir> --engine_precision float32 --engine_dir whisper_outputs` — the reference's
examples/whisper/build_decoder.py with the same flags: writes `WhisperDecoder.engine` into the engine directory."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--whisper", type=str, default="whisper-tiny.en")
    parser.add_argument("--engine_precision", type=str, default="float32", choices=["float32", "bfloat16"])
    parser.add_argument("--log_level", type=str, default="error")
    parser.add_argument("--engine_dir", type=str, default="whisper_outputs")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_arguments(argv)
    from whisper_trtllm_b200 import run
    from whisper_trtllm_b200.checkpoint import load_hf_checkpoint
    config, ckpt = load_hf_checkpoint(args.whisper)
    engine = run.build_decoder(config, ckpt, engine_dir=args.engine_dir, engine_precision=args.engine_precision)
    print(f"{os.path.join(args.engine_dir, run.DECODER_ENGINE)}: {len(engine)} bytes")


if __name__ == "__main__":
    main()

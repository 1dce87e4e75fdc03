#!/usr/bin/env python
"""`python build_encoder.py --whisper <checkpoint dir> --engine_precision float32 --engine_dir whisper_outputs` — the reference's
examples/whisper/build_encoder.py with the same flags: writes `WhisperEncoder.engine` and `config.pkl` into the engine directory.
The "engine" is a container of the bound weights (nothing is traced or compiled); precision is float32 or bfloat16."""
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
    config, ckpt = load_hf_checkpoint(args.whisper)          # build_encoder.py:38-45: config = hf_model.config.to_dict(), ckpt = state_dict()
    engine = run.build_encoder(config, ckpt, engine_dir=args.engine_dir, engine_precision=args.engine_precision)
    print(f"{os.path.join(args.engine_dir, run.ENCODER_ENGINE)}: {len(engine)} bytes; {os.path.join(args.engine_dir, run.CONFIG_PKL)} written")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""`python run.py --whisper <checkpoint dir> --engine_dir whisper_outputs [--compare]` — the reference's examples/whisper/run.py
main (:229-331) on whisper_trtllm_b200, same flags and the same report: engines and `config.pkl` from the engine directory
(build_encoder.py / build_decoder.py next to this file), the runner classes `WhisperEncoder(args, config)` /
`WhisperDecoder(args, config)`, `get_logits_processor`, `get_stopping_criteria`, `greedy_search`; the dataset is walked twice
(the first pass is the warm-up), one utterance at a time as the reference does (`--batch N` for more).

What differs underneath: log-mel on the GPU instead of the CPU feature extractor, `link()` packs both engines into one native
runtime so that `greedy_search` runs the whole token loop on the device, ids -> text by the repo's detokenizer.
`--dataset` is ./librispeech_asr_dummy (an HF `datasets` directory, as in the reference) or any directory / manifest that
`whisper_trtllm_b200.audio.read_manifest` understands.
"""
import argparse
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--whisper", type=str, default="whisper-tiny.en", required=True)
    parser.add_argument("--engine_precision", type=str, default="float32")
    parser.add_argument("--log_level", type=str, default="error")
    parser.add_argument("--engine_dir", type=str, default="whisper_outputs")
    parser.add_argument("--compare", action="store_true")
    parser.add_argument("--dataset", type=str, default="./librispeech_asr_dummy")
    parser.add_argument("--batch", type=int, default=1, help="utterances per greedy_search call (the reference: 1)")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_arguments(argv)
    import torch
    from whisper_trtllm_b200 import run
    from whisper_trtllm_b200.audio import batches, is_hf_dataset, load_audio, read_hf_dataset, read_manifest
    from whisper_trtllm_b200.frontend import LogMelFrontend
    from whisper_trtllm_b200.pipeline import compare_transcriptions, load_text_tools

    torch.cuda.set_device(0)
    # load dataset (run.py:241-247)
    if is_hf_dataset(args.dataset):
        print("loading dataset from disk...")
        waves, _ = read_hf_dataset(args.dataset)
    else:
        paths, _ = read_manifest(args.dataset)
        waves = [load_audio(p) for p in paths]
    # load whisper config (run.py:250-252)
    with open(os.path.join(args.engine_dir, run.CONFIG_PKL), "rb") as f:
        config = pickle.load(f)
    # init whisper (run.py:254-256)
    whisperencoder = run.WhisperEncoder(args=args, config=config)
    whisperdecoder = run.WhisperDecoder(args=args, config=config)
    run.link(whisperencoder, whisperdecoder, config, max_batch=args.batch)
    processor = LogMelFrontend()                                  # hf_processor(...).input_features, on the GPU
    detokenizer, _ = load_text_tools(args.whisper, config)        # hf_processor.batch_decode
    if detokenizer is None:
        raise SystemExit(f"{args.whisper} has no vocab.json")

    # go through the dataset twice, the first pass is the warm-up (run.py:259-291)
    for _ in range(2):
        transcriptions = []
        torch.cuda.synchronize()
        start_time = time.time()
        for chunk in batches(waves, args.batch):
            input_features = processor(chunk)
            encoder_outputs = whisperencoder(input_features)
            input_ids = torch.full((len(chunk), 1), config["decoder_start_token_id"], dtype=torch.int32, device="cuda")
            predicted_ids = run.greedy_search(
                model=whisperdecoder, encoder_outputs=encoder_outputs, input_ids=input_ids,
                logits_processor=run.get_logits_processor(config, input_ids.shape[-1]),
                stopping_criteria=run.get_stopping_criteria(config),
                pad_token_id=config["pad_token_id"], eos_token_id=config["eos_token_id"])
            transcriptions.extend(detokenizer.batch_decode(predicted_ids, skip_special_tokens=True))
        torch.cuda.synchronize()
        end_time = time.time()
    b200_time = end_time - start_time
    for t in transcriptions:
        print(t)

    if args.compare:                                              # run.py:293-331
        from transformers import WhisperForConditionalGeneration, WhisperProcessor
        hf_processor = WhisperProcessor.from_pretrained(args.whisper)
        hf_model = WhisperForConditionalGeneration.from_pretrained(args.whisper)
        for _ in range(2):
            hf_transcriptions = []
            start_time = time.time()
            for w in waves:
                feats = hf_processor(w, sampling_rate=16000, return_tensors="pt").input_features
                hf_transcriptions.extend(hf_processor.batch_decode(hf_model.generate(feats), skip_special_tokens=True))
            end_time = time.time()
        hf_time = end_time - start_time
        print("B200 time: ", b200_time)
        print("Huggingface  time: ", hf_time)
        print("Speed up: ", hf_time / b200_time)
        diff = compare_transcriptions(transcriptions, hf_transcriptions)
        print(f"Compare Result: same [{len(transcriptions) - len(diff)}], diff [{len(diff)}]")
        for a, b in diff:
            print("-------------------------")
            print(f"B200        : {a}")
            print(f"Huggingface : {b}")
    else:
        print("B200 time: ", b200_time)


if __name__ == "__main__":
    main()

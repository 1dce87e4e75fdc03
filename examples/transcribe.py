#!/usr/bin/env python
"""Transcribe audio files and (optionally) score them: the `__main__` blocks of the reference's run.py:229-331 and
cal_wer.py:227-287 on top of whisper_trtllm_b200 — batched, log-mel on the GPU, greedy loop on the device.

    python examples/transcribe.py --whisper whisper_models/whisper-medium.en --audio LibriSpeech/test-clean/1089/134686
    python examples/transcribe.py --whisper ... --audio manifest.tsv --wer                # WER against the manifest's texts
    python examples/transcribe.py --whisper ... --audio wav_dir --compare                 # vs HuggingFace on the CPU (run.py --compare)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/transcribe.py ...   # sharded by utterance

`--audio` is a directory of 16 kHz .wav / .flac / .npy files, a LibriSpeech `*.trans.txt`, a TSV `<path>\\t<text>`, a .jsonl
(`whisper_trtllm_b200.audio.read_manifest`) or an HF `datasets` directory such as the reference's ./librispeech_asr_dummy
(run.py:241-247; FLAC bytes are decoded by the repo's own decoder, libwb_audio.so), or cal_wer.py's `librispeech.cache`.  Needs a B200 (there is no CPU path) and a checkpoint directory in the HF layout
(config.json, model.safetensors, vocab.json, optionally normalizer.json).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("--whisper", required=True, help="checkpoint directory (the reference's --whisper, run.py:48)")
    p.add_argument("--audio", required=True, help="directory / *.trans.txt / .tsv / .jsonl listing the utterances")
    p.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float32"])
    p.add_argument("--batch", type=int, default=64, help="utterances per engine call (per GPU)")
    p.add_argument("--compact-every", type=int, default=32, help="drop finished utterances from the batch every N tokens (0: never)")
    p.add_argument("--wer", action="store_true", help="print the word error rate against the manifest's reference texts")
    p.add_argument("--compare", action="store_true", help="also run HuggingFace transformers on the CPU and list differences")
    p.add_argument("--out", default=None, help="write {audio, text[, reference]} JSON lines here")
    return p.parse_args(argv)


def huggingface_transcriptions(checkpoint_dir, paths):
    """The reference's comparison arm (run.py:293-313): stock transformers, fp32, CPU, one utterance at a time."""
    from transformers import WhisperForConditionalGeneration, WhisperProcessor
    from whisper_trtllm_b200.audio import load_audio
    processor = WhisperProcessor.from_pretrained(checkpoint_dir)
    model = WhisperForConditionalGeneration.from_pretrained(checkpoint_dir)
    out = []
    for p in paths:
        feats = processor(load_audio(p), sampling_rate=16000, return_tensors="pt").input_features
        out.extend(processor.batch_decode(model.generate(feats), skip_special_tokens=True))
    return out


def main(argv=None):
    args = parse_arguments(argv)
    import torch
    import torch.distributed as dist
    from whisper_trtllm_b200.audio import is_hf_dataset, read_hf_dataset, read_manifest, read_mel_cache
    from whisper_trtllm_b200.pipeline import WhisperPipeline, compare_transcriptions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")

    waves, mels = None, None
    if args.audio.endswith(".cache"):            # cal_wer.py's pickled (log-mel, text) pairs (get_LibriSpeech.py)
        mels, references = read_mel_cache(args.audio)
        paths = [f"{os.path.basename(args.audio)}[{i}]" for i in range(mels.shape[0])]
    elif is_hf_dataset(args.audio):
        waves, references = read_hf_dataset(args.audio)
        paths = [f"{os.path.basename(os.path.normpath(args.audio))}[{i}]" for i in range(len(waves))]
    else:
        paths, references = read_manifest(args.audio)
    if args.wer and references is None:
        raise SystemExit(f"--wer needs reference texts, {args.audio} has none")
    pipe = WhisperPipeline(args.whisper, dtype=args.dtype, max_batch=args.batch, compact_every=args.compact_every)

    torch.cuda.synchronize()
    t0 = time.time()
    if mels is not None:
        ids = pipe.transcribe_sharded(mels, features=True)
    else:
        ids = pipe.transcribe_files(paths) if waves is None else pipe.transcribe_sharded(waves)
    torch.cuda.synchronize()
    elapsed = time.time() - t0
    if rank == 0:
        texts = pipe.decode(ids)
        for p, t in zip(paths, texts):
            print(f"{os.path.basename(p)}\t{t}")
        print(f"{len(paths)} utterances in {elapsed:.2f} s on {world} GPU(s) (checkpoint load and first-call set-up excluded)",
              file=sys.stderr)
        if args.out:
            with open(args.out, "w", encoding="utf-8") as f:
                for i, (p, t) in enumerate(zip(paths, texts)):
                    row = {"audio": p, "text": t}
                    if references is not None:
                        row["reference"] = references[i]
                    f.write(json.dumps(row, ensure_ascii=False) + "\n")
        if args.wer:
            print(f"WER: {pipe.wer(texts, references) * 100:.2f} %")                      # cal_wer.py:287
        if args.compare:
            if waves is not None or mels is not None:
                raise SystemExit("--compare reads audio files; export the dataset's audio to a directory first")
            t0 = time.time()
            hf = huggingface_transcriptions(args.whisper, paths)
            hf_time = time.time() - t0
            diff = compare_transcriptions(texts, hf)
            print("B200 time: ", elapsed)
            print("Huggingface  time: ", hf_time)
            print("Speed up: ", hf_time / elapsed)
            print(f"Compare Result: same [{len(texts) - len(diff)}], diff [{len(diff)}]")  # run.py:322-331
            for a, b in diff:
                print("-------------------------")
                print(f"B200        : {a}")
                print(f"Huggingface : {b}")
    pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for ids -> text (test infrastructure, not shipped).

Imports the REAL reference tokenizers — /root/reference/transformers/src/transformers/models/whisper/tokenization_whisper.py
(`WhisperTokenizer`, the slow class) and tokenization_whisper_fast.py (`WhisperTokenizerFast`, the class
`WhisperProcessor.from_pretrained` hands to run.py:287 when `tokenizers` is installed) — in the build container, over a small
synthetic byte-level vocabulary (the checkpoints' vocab.json is not in the image): the 256 byte tokens, a few multi-byte
pieces, characters split across tokens, and the full set of Whisper special tokens added the way the real tokenizer files
list them.  Records `decode(ids, skip_special_tokens=True/False)` of both classes on fixed and seeded random id sequences:

    tests/golden/detokenizer.json = {"vocab", "added_tokens", "cases": [[ids, slow_skip, slow_keep, fast_skip, fast_keep], ...]}

    python oracle/make_golden_detok.py
"""
import json
import os
import random
import shutil
import sys
import tempfile
import types

REF_SRC = "/root/reference/transformers/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PIECES = [" the", " Hello", "ing", " world", " don", "'t", " ,", " .", " 's", " n't", " 're", " ?", " !", " 'm", " 've", " '", " caf", "é",
          " na", "ï", "ve", " I", " it", " we", " you", "He", " said", ":", " \"", "\"", " 5", "0", "%", " $", "20", " o", "'clock", "…", " —",
          " 日本", "語", "\n", "  ", " Mr", ".", ",", " Smith", " isn", " they", " can", " wo"]


def main():
    sys.dont_write_bytecode = True
    stub = types.ModuleType("transformers.dependency_versions_check")
    stub.dep_version_check = lambda *a, **k: None
    sys.modules["transformers.dependency_versions_check"] = stub
    sys.path.insert(0, REF_SRC)
    from transformers.models.whisper.tokenization_whisper import LANGUAGES, WhisperTokenizer, bytes_to_unicode
    from transformers.models.whisper.tokenization_whisper_fast import WhisperTokenizerFast

    b2u = bytes_to_unicode()
    enc = lambda s: "".join(b2u[b] for b in (s if isinstance(s, bytes) else s.encode("utf-8")))
    vocab = {b2u[b]: b for b in range(256)}
    for w in PIECES:
        vocab.setdefault(enc(w), len(vocab))
    smile = "😀".encode("utf-8")                       # one character split over two tokens (byte-level BPE does this)
    vocab[enc(smile[:2])] = len(vocab)
    vocab[enc(smile[2:])] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    specials = (["<|startoftranscript|>"] + [f"<|{l}|>" for l in LANGUAGES]
                + ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nocaptions|>", "<|notimestamps|>"])
    work = tempfile.mkdtemp()
    try:
        with open(os.path.join(work, "vocab.json"), "w", encoding="utf-8") as f:
            json.dump(vocab, f, ensure_ascii=False)
        with open(os.path.join(work, "merges.txt"), "w") as f:
            f.write("#version: 0.2\n")
        slow = WhisperTokenizer(os.path.join(work, "vocab.json"), os.path.join(work, "merges.txt"))
        slow.add_special_tokens({"additional_special_tokens": specials})
        slow.save_pretrained(os.path.join(work, "saved"))
        fast = WhisperTokenizerFast.from_pretrained(os.path.join(work, "saved"))
        with open(os.path.join(work, "saved", "added_tokens.json"), encoding="utf-8") as f:
            added = json.load(f)
    except Exception:
        shutil.rmtree(work, ignore_errors=True)
        raise
    eot, sot, prev, nots = vocab["<|endoftext|>"], added["<|startoftranscript|>"], added["<|startofprev|>"], added["<|notimestamps|>"]
    n_text = eot                                         # ids below <|endoftext|> are text pieces
    T = lambda w: vocab[enc(w)]
    fixed = [
        [sot, nots, T(" Hello"), T(" world"), T(" ."), eot, eot],
        [sot, nots, T(" I"), T(" don"), T(" n't"), T(" ,"), T(" it"), T(" 's"), T(" ?"), T(" we"), T(" 're"), T(" !"), T(" I"), T(" 'm"), T(" you"), T(" 've"), eot],
        [sot, nots, T(" caf"), T("é"), T(" na"), T("ï"), T("ve"), T(" 日本"), T("語"), vocab[enc(smile[:2])], vocab[enc(smile[2:])], eot],
        [prev, T(" the"), T(" world"), sot, nots, T(" Hello"), eot],                    # prompt is stripped when specials are skipped
        [prev, T(" the"), T(" world")],                                                  # a prompt without a transcript
        [sot, added["<|en|>"], added["<|transcribe|>"], nots, T("He"), T(" said"), T(":"), T(" \""), T(" Hello"), T(" '"), T(" ,"), T("\""), eot],
        [T(" Mr"), T("."), T(" Smith"), T(" isn"), T("'t"), T(" wo"), T(" n't"), T(" ."), T(" .")],
        [T(" 5"), T("0"), T("%"), T(" $"), T("20"), T(" o"), T("'clock"), T("…"), T(" —"), T("\n"), T("  "), T(" the")],
        [], [eot], [sot, nots, eot], [T(" .")], [T(" '"), T(" ")],
    ]
    rng = random.Random(20240611)
    cases = [list(map(int, c)) for c in fixed]
    printable = [i for i in range(n_text) if i >= 256 or 32 <= i < 127]   # random cases stay valid UTF-8: ASCII bytes + whole pieces
    printable = [i for i in printable if i not in (vocab[enc(smile[:2])], vocab[enc(smile[2:])], T("é"), T("ï"), T("語"), T(" 日本"), T("…"), T(" —"))]
    pairs = [[T(" caf"), T("é")], [T(" na"), T("ï"), T("ve")], [T(" 日本"), T("語")], [vocab[enc(smile[:2])], vocab[enc(smile[2:])]], [T("…")], [T(" —")]]
    for _ in range(400):
        ids = [sot, nots] if rng.random() < 0.7 else []
        for _ in range(rng.randint(0, 14)):
            ids += rng.choice(pairs) if rng.random() < 0.15 else [rng.choice(printable)]
        ids += [eot] * rng.randint(0, 3)
        cases.append(ids)
    out = []
    for ids in cases:
        out.append([ids, slow.decode(ids, skip_special_tokens=True), slow.decode(ids, skip_special_tokens=False),
                    fast.decode(ids, skip_special_tokens=True), fast.decode(ids, skip_special_tokens=False)])
    shutil.rmtree(work, ignore_errors=True)
    path = os.path.join(ROOT, "tests", "golden", "detokenizer.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"vocab": vocab, "added_tokens": added, "first_special_id": eot, "slow_clean_up": False,
                   "fast_clean_up": bool(fast.clean_up_tokenization_spaces), "cases": out}, f, ensure_ascii=False, indent=0)
    differ = sum(1 for c in out if c[1] != c[3])
    print(f"{len(out)} cases -> {path}; slow and fast outputs differ on {differ}")


if __name__ == "__main__":
    main()

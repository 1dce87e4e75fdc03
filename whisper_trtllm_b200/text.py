"""Token ids -> text and WER: the step AFTER the hot path in the reference's scripts —
``hf_processor.batch_decode(predicted_ids, skip_special_tokens=True)`` (run.py:287, cal_wer.py:275), the text
normaliser and ``jiwer.wer`` (cal_wer.py:279-287).  Pure host-side string work, no arithmetic on the path.

* ``WhisperDetokenizer``: Whisper's tokenizer is GPT-2 byte-level BPE; decoding needs only ``vocab.json``
  (token string -> id) and the byte<->unicode table (tokenization_whisper.py ``bytes_to_unicode`` / ``decode``).
  Special tokens are every id >= the first added token (``<|endoftext|>`` = 50256 for the ``.en`` vocabularies).
* ``BasicTextNormalizer``: the language-independent normaliser of english_normalizer.py (lower-case, drop bracketed
  spans, strip symbols/punctuation, collapse spaces).  The full ``EnglishTextNormalizer`` (number words, contractions,
  British->American spelling map, english_normalizer.py:510-595) depends on a 1700-entry spelling table that ships with
  the checkpoints, not with this repo — pass ``spelling_map`` to apply it when available.
* ``wer``: word error rate = (S + D + I) / N over the whole corpus, as ``jiwer.wer`` (jiwer is not installed).
"""
from __future__ import annotations

import json
import re
import unicodedata
from typing import Dict, Iterable, List, Optional, Sequence, Union


def bytes_to_unicode() -> Dict[int, str]:
    """The reversible byte -> printable unicode table of GPT-2 byte-level BPE (tokenization_whisper.py:~40-60)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


class WhisperDetokenizer:
    def __init__(self, vocab: Union[str, Dict[str, int]], first_special_id: int = 50256):
        if isinstance(vocab, str):
            with open(vocab, encoding="utf-8") as f:
                vocab = json.load(f)
        self.id_to_token = {int(i): t for t, i in vocab.items()}
        self.first_special_id = first_special_id
        self.byte_decoder = {c: b for b, c in bytes_to_unicode().items()}

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True) -> str:
        pieces: List[str] = []
        for i in ids:
            i = int(i)
            if i >= self.first_special_id:
                if not skip_special_tokens:
                    pieces.append(self.id_to_token.get(i, f"<|{i}|>"))
                continue
            tok = self.id_to_token.get(i)
            if tok is None:
                raise KeyError(f"token id {i} is not in the vocabulary")
            pieces.append(tok)
        text = "".join(pieces)
        out = bytearray()
        for ch in text:
            if ch in self.byte_decoder:
                out.append(self.byte_decoder[ch])
            else:                       # a special token kept verbatim
                out.extend(ch.encode("utf-8"))
        return out.decode("utf-8", errors="replace")

    def batch_decode(self, batch_ids, skip_special_tokens: bool = True) -> List[str]:
        rows = batch_ids.tolist() if hasattr(batch_ids, "tolist") else batch_ids
        return [self.decode(r, skip_special_tokens) for r in rows]


class BasicTextNormalizer:
    def __init__(self, spelling_map: Optional[Dict[str, str]] = None):
        self.spelling_map = spelling_map or {}

    @staticmethod
    def _strip_symbols(s: str) -> str:
        # replace markers / symbols / punctuation by a space, keep everything else (english_normalizer.py remove_symbols)
        return "".join(" " if unicodedata.category(c)[0] in "MSP" else c for c in unicodedata.normalize("NFKC", s))

    def __call__(self, s: str) -> str:
        s = s.lower()
        s = re.sub(r"[<\[][^>\]]*[>\]]", "", s)      # remove words between brackets
        s = re.sub(r"\(([^)]+?)\)", "", s)           # remove words between parentheses
        s = self._strip_symbols(s)
        words = [self.spelling_map.get(w, w) for w in s.split()]
        return " ".join(words)


def _edit_distance(ref: Sequence[str], hyp: Sequence[str]) -> int:
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def wer(references: Union[str, Sequence[str]], hypotheses: Union[str, Sequence[str]]) -> float:
    """Corpus word error rate: total word-level edit distance / total reference words."""
    if isinstance(references, str):
        references, hypotheses = [references], [hypotheses]
    if len(references) != len(hypotheses):
        raise ValueError("references and hypotheses differ in length")
    edits = words = 0
    for r, h in zip(references, hypotheses):
        rw, hw = r.split(), h.split()
        edits += _edit_distance(rw, hw)
        words += len(rw)
    if words == 0:
        raise ValueError("no reference words")
    return edits / words

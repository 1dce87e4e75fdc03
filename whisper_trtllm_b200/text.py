"""Token ids -> text and WER: the step AFTER the hot path in the reference's scripts —
``hf_processor.batch_decode(predicted_ids, skip_special_tokens=True)`` (run.py:287, cal_wer.py:275), the text
normaliser and ``jiwer.wer`` (cal_wer.py:279-287).  Pure host-side string work, no arithmetic on the path.

* ``WhisperDetokenizer``: Whisper's tokenizer is GPT-2 byte-level BPE; decoding needs only ``vocab.json``
  (token string -> id) and the byte<->unicode table (tokenization_whisper.py ``bytes_to_unicode`` / ``_decode``).
  Special tokens are every id >= the first added token (``<|endoftext|>`` = 50256 for the ``.en`` vocabularies).  Output equals
  the reference's ``WhisperTokenizer.decode`` (clean-up off) and ``WhisperTokenizerFast.decode`` (clean-up on, the default of
  ``WhisperProcessor``) — tests/golden/detokenizer.json.
* ``BasicTextNormalizer`` / ``EnglishTextNormalizer`` (+ ``EnglishNumberNormalizer``, ``EnglishSpellingNormalizer``): the
  normalisers of english_normalizer.py:75-595 — same class names, constructor arguments and outputs (pinned against the
  reference's own outputs on 2500 inputs, tests/golden/english_normalizer.json).  The British->American table is the
  checkpoint's ``normalizer.json`` (data, not in this repo): pass it as ``english_spelling_mapping`` when available.
* ``wer``: word error rate = (S + D + I) / N over the whole corpus, as ``jiwer.wer`` (jiwer is not installed).
"""
from __future__ import annotations

import json
import re
import unicodedata
from fractions import Fraction
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union


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


# `PreTrainedTokenizerBase.clean_up_tokenization` (tokenization_utils_base.py:3598-3620): English detokenisation artefacts,
# applied in this order by the FAST tokenizer's decode (tokenization_utils_fast.py:566-575) — the class
# `WhisperProcessor.from_pretrained` (run.py:239) loads when `tokenizers` is installed.  The slow WhisperTokenizer._decode
# (tokenization_whisper.py:613-645) never applies it.
_CLEAN_UP = ((" .", "."), (" ?", "?"), (" !", "!"), (" ,", ","), (" ' ", "'"), (" n't", "n't"), (" 'm", "'m"), (" 's", "'s"),
             (" 've", "'ve"), (" 're", "'re"))


def clean_up_tokenization(text: str) -> str:
    for a, b in _CLEAN_UP:
        text = text.replace(a, b)
    return text


class WhisperDetokenizer:
    def __init__(self, vocab: Union[str, Dict[str, int]], first_special_id: int = 50256, clean_up_tokenization_spaces: bool = True,
                 added_tokens: Optional[Dict[str, int]] = None):
        """``clean_up_tokenization_spaces``: the tokenizer's setting of that name (``tokenizer_config.json``; the base-class default
        is True, tokenization_utils_base.py:1552).  ``added_tokens``: the checkpoint's ``added_tokens.json`` (special-token
        strings for ``skip_special_tokens=False`` and the prompt markers)."""
        if isinstance(vocab, str):
            with open(vocab, encoding="utf-8") as f:
                vocab = json.load(f)
        self.id_to_token = {int(i): t for t, i in vocab.items()}
        for t, i in (added_tokens or {}).items():
            self.id_to_token.setdefault(int(i), t)
        self.first_special_id = first_special_id
        self.clean_up_tokenization_spaces = bool(clean_up_tokenization_spaces)
        self.byte_decoder = {c: b for b, c in bytes_to_unicode().items()}
        token_to_id = {t: i for i, t in self.id_to_token.items()}
        self.prompt_token_id = token_to_id.get("<|startofprev|>")
        self.decoder_start_token_id = token_to_id.get("<|startoftranscript|>")

    def _strip_prompt(self, ids: List[int]) -> List[int]:
        """A sequence that starts with <|startofprev|> carries a text prompt: drop it up to <|startoftranscript|>
        (tokenization_whisper.py `_strip_prompt`, applied when special tokens are skipped)."""
        if ids and self.prompt_token_id is not None and ids[0] == self.prompt_token_id:
            if self.decoder_start_token_id in ids:
                return ids[ids.index(self.decoder_start_token_id):]
            return []
        return ids

    def decode(self, ids: Iterable[int], skip_special_tokens: bool = True, clean_up_tokenization_spaces: Optional[bool] = None) -> str:
        ids = [int(i) for i in ids]
        if skip_special_tokens:
            ids = self._strip_prompt(ids)
        out = bytearray()
        for i in ids:
            if i >= self.first_special_id:
                if not skip_special_tokens:           # special tokens are kept verbatim, not byte-decoded
                    out.extend(self.id_to_token.get(i, f"<|{i}|>").encode("utf-8"))
                continue
            tok = self.id_to_token.get(i)
            if tok is None:
                raise KeyError(f"token id {i} is not in the vocabulary")
            out.extend(self.byte_decoder[ch] for ch in tok)
        text = out.decode("utf-8", errors="replace")
        if self.clean_up_tokenization_spaces if clean_up_tokenization_spaces is None else clean_up_tokenization_spaces:
            text = clean_up_tokenization(text)
        return text

    def batch_decode(self, batch_ids, skip_special_tokens: bool = True, clean_up_tokenization_spaces: Optional[bool] = None) -> List[str]:
        rows = batch_ids.tolist() if hasattr(batch_ids, "tolist") else batch_ids
        return [self.decode(r, skip_special_tokens, clean_up_tokenization_spaces) for r in rows]


# ---------------------------------------------------------------------------------------------------------------------
# Text normalisers (english_normalizer.py).  Same observable behaviour as the reference classes, quirks included
# (tests/golden/english_normalizer.json holds the reference's outputs); organised here as table-driven passes.

_LIGATURES = {"œ": "oe", "Œ": "OE", "ø": "o", "Ø": "O", "æ": "ae", "Æ": "AE", "ß": "ss", "ẞ": "SS", "đ": "d", "Đ": "D",
              "ð": "d", "Ð": "D", "þ": "th", "Þ": "th", "ł": "l", "Ł": "L"}     # letters NFKD does not decompose (:24-41)
_BRACKETED = re.compile(r"[<\[][^>\]]*[>\]]")
_PARENTHESISED = re.compile(r"\(([^)]+?)\)")
_SPACES = re.compile(r"\s+")


def strip_symbols(s: str) -> str:
    """Markers / symbols / punctuation (unicode categories M*, S*, P*) -> one space each, after NFKC
    (english_normalizer.py:68-72)."""
    return "".join(" " if unicodedata.category(c)[0] in "MSP" else c for c in unicodedata.normalize("NFKC", s))


def strip_symbols_and_diacritics(s: str, keep: str = "") -> str:
    """As `strip_symbols` but over NFKD, dropping combining marks (Mn) and spelling out the ligatures NFKD leaves alone;
    characters in `keep` pass through (english_normalizer.py:44-65)."""
    out: List[str] = []
    for c in unicodedata.normalize("NFKD", s):
        if c in keep:
            out.append(c)
        elif c in _LIGATURES:
            out.append(_LIGATURES[c])
        else:
            cat = unicodedata.category(c)
            if cat == "Mn":
                continue
            out.append(" " if cat[0] in "MSP" else c)
    return "".join(out)


class BasicTextNormalizer:
    """Language-independent normaliser (english_normalizer.py:75-93): lower-case, drop bracketed / parenthesised spans,
    symbols -> spaces, optional grapheme splitting, runs of whitespace -> one space (ends are NOT stripped)."""

    def __init__(self, remove_diacritics: bool = False, split_letters: bool = False):
        self.clean = strip_symbols_and_diacritics if remove_diacritics else strip_symbols
        self.split_letters = split_letters

    def __call__(self, s: str) -> str:
        s = _PARENTHESISED.sub("", _BRACKETED.sub("", s.lower()))
        s = self.clean(s).lower()
        if self.split_letters:
            import regex                                  # third-party `regex` for \X (extended grapheme clusters)
            s = " ".join(regex.findall(r"\X", s, regex.U))
        return _SPACES.sub(" ", s)


# word classes of the number pass
_ZERO, _UNIT, _UNIT_SFX, _TEN, _TEN_SFX, _MULT, _MULT_SFX, _SIGN, _CURRENCY, _PERCENT, _PER, _AND, _REPEAT, _POINT = range(14)
_NUMERIC = re.compile(r"^\d+(\.\d+)?$")


def _build_lexicon() -> Dict[str, tuple]:
    """word -> (class, value, suffix / symbol).  Vocabulary of english_normalizer.py:107-207."""
    lex: Dict[str, tuple] = {}
    units = ("one two three four five six seven eight nine ten eleven twelve thirteen fourteen fifteen sixteen seventeen "
             "eighteen nineteen").split()
    irregular = {1: "first", 2: "second", 3: "third", 5: "fifth", 12: "twelfth"}
    for w in ("o", "oh", "zero"):
        lex[w] = (_ZERO, 0, None)
    lex["zeroth"] = (_UNIT_SFX, 0, "th")
    for n, w in enumerate(units, 1):
        lex[w] = (_UNIT, n, None)
        lex["sixes" if w == "six" else w + "s"] = (_UNIT_SFX, n, "s")
        if n in irregular:
            lex[irregular[n]] = (_UNIT_SFX, n, "th" if n > 3 else ("st", "nd", "rd")[n - 1])
        else:
            lex[w + ("h" if w.endswith("t") else "th")] = (_UNIT_SFX, n, "th")
    for n, w in enumerate("twenty thirty forty fifty sixty seventy eighty ninety".split(), 2):
        lex[w] = (_TEN, 10 * n, None)
        lex[w.replace("y", "ies")] = (_TEN_SFX, 10 * n, "s")
        lex[w.replace("y", "ieth")] = (_TEN_SFX, 10 * n, "th")
    big = "thousand million billion trillion quadrillion quintillion sextillion septillion octillion nonillion decillion".split()
    for w, m in [("hundred", 100)] + [(w, 1000 ** k) for k, w in enumerate(big, 1)]:
        lex[w] = (_MULT, m, None)
        lex[w + "s"] = (_MULT_SFX, m, "s")
        lex[w + "th"] = (_MULT_SFX, m, "th")
    for w, sym in (("minus", "-"), ("negative", "-"), ("plus", "+"), ("positive", "+")):
        lex[w] = (_SIGN, None, sym)
    for w, sym in (("pound", "£"), ("euro", "€"), ("dollar", "$"), ("cent", "¢")):
        lex[w] = lex[w + "s"] = (_CURRENCY, None, sym)
    lex["percent"] = (_PERCENT, None, "%")
    lex["per"] = (_PER, None, None)                     # "per cent" -> %
    lex["and"] = (_AND, None, None)
    lex["double"] = (_REPEAT, 2, None)
    lex["triple"] = (_REPEAT, 3, None)
    lex["point"] = (_POINT, None, None)
    return lex


class EnglishNumberNormalizer:
    """Spelled-out numbers -> digits (english_normalizer.py:96-491): commas gone, suffixes kept (`1960s`, `274th`), currency
    words become a symbol before the number, percent a `%` after it, runs of single digits read as one nominal number
    (`one oh one` -> `101`), and a lone `1` / `1s` goes back to `one` / `ones`.

    One left-to-right pass keeps a pending number — an `int` while it is still a quantity that can be multiplied
    ("two hundred"), a `str` once digits have been concatenated ("one oh", "3.") — plus a pending sign / currency symbol.
    """

    _SIGN_OR_CURRENCY = frozenset("-+£€$¢")

    def __init__(self):
        self.lexicon = _build_lexicon()
        self.words = frozenset(self.lexicon)
        self._half_ok = frozenset(w for w, (k, _, _) in self.lexicon.items() if k in (_ZERO, _UNIT, _TEN, _MULT))

    # -- pass 1: "<number> and a half" -> "<number> point five"; split digits from letters (:433-460)
    def preprocess(self, s: str) -> str:
        pieces = re.split(r"\band\s+a\s+half\b", s)
        kept: List[str] = []
        for i, piece in enumerate(pieces):
            if not piece.strip():
                continue
            kept.append(piece)
            if i + 1 < len(pieces):
                kept.append("point five" if piece.rsplit(maxsplit=2)[-1] in self._half_ok else "and a half")
        s = " ".join(kept)
        s = re.sub(r"([a-z])([0-9])", r"\1 \2", s)
        s = re.sub(r"([0-9])([a-z])", r"\1 \2", s)
        return re.sub(r"([0-9])\s+(st|nd|rd|th|s)\b", r"\1\2", s)      # "21 st" is a suffix, re-attach

    # -- pass 2: the word scan (:209-431)
    def process_words(self, words: Sequence[str]) -> Iterator[str]:
        lex = self.lexicon
        pending: Union[None, int, str] = None      # the number being assembled
        symbol: Optional[str] = None               # sign / currency to put in front of it
        emitted: List[str] = []

        def emit(text) -> None:
            nonlocal pending, symbol
            emitted.append(str(text) if symbol is None else symbol + str(text))
            pending = symbol = None

        def flush() -> None:
            if pending is not None:
                emit(pending)

        def digits_so_far() -> str:                # the reference's `str(value or "")`: an int 0 counts as nothing
            return str(pending or "")

        def as_fraction(v):
            try:
                return Fraction(v)
            except ValueError:
                return None

        def joined(n: int, after_ten: bool) -> Union[int, str]:
            """Pending number followed by a unit word worth n (1..19)."""
            if isinstance(pending, str) or prev_is_unit:
                if after_ten and n < 10:           # "twenty" + "one" inside a digit string: overwrite the trailing 0
                    return pending[:-1] + str(n)
                return str(pending) + str(n)
            if pending % (10 if n < 10 else 100) == 0:
                return pending + n
            return str(pending) + str(n)

        def scaled(m: int):
            """Pending int times a multiplier word: only the part below 1000 is scaled."""
            return pending // 1000 * 1000 + pending % 1000 * m

        skip = False
        n_words = len(words)
        for i, word in enumerate(words):
            if skip:
                skip = False
                continue
            prev = words[i - 1] if i else None
            nxt = words[i + 1] if i + 1 < n_words else None
            nxt_numeric = nxt is not None and _NUMERIC.match(nxt) is not None
            nxt_known = nxt in lex
            prev_kind = lex[prev][0] if prev in lex else None
            prev_is_unit = prev_kind == _UNIT

            signed = word[0] in self._SIGN_OR_CURRENCY
            bare = word[1:] if signed else word
            if _NUMERIC.match(bare):               # digits, possibly "$3.50" / "-2"
                if pending is not None:
                    if isinstance(pending, str) and pending.endswith("."):
                        pending += word            # decimals / dotted quads grow as text
                        continue
                    flush()
                if signed:
                    symbol = word[0]
                frac = Fraction(bare)
                pending = frac.numerator if frac.denominator == 1 else bare
                continue

            entry = lex.get(word)
            if entry is None:                      # ordinary word
                flush()
                emit(word)
                continue
            kind, val, sfx = entry

            if kind == _ZERO:
                pending = digits_so_far() + "0"
            elif kind == _UNIT:
                pending = val if pending is None else joined(val, prev_kind == _TEN)
            elif kind == _UNIT_SFX:                # ordinal / plural closes the number
                emit(f"{val if pending is None else joined(val, prev_kind == _TEN)}{sfx}")
            elif kind == _TEN:
                if pending is None:
                    pending = val
                elif isinstance(pending, str) or pending % 100:
                    pending = str(pending) + str(val)
                else:
                    pending += val
            elif kind == _TEN_SFX:
                if pending is None:
                    emit(f"{val}{sfx}")
                elif isinstance(pending, str) or pending % 100:
                    emit(f"{pending}{val}{sfx}")
                else:
                    emit(f"{pending + val}{sfx}")
            elif kind == _MULT:
                if pending is None:
                    pending = val
                elif isinstance(pending, str) or pending == 0:
                    f = as_fraction(pending)
                    if f is not None and (f * val).denominator == 1:
                        pending = (f * val).numerator          # "2.5 million"
                    else:
                        flush()
                        pending = val
                else:
                    pending = scaled(val)
            elif kind == _MULT_SFX:
                if pending is None:
                    emit(f"{val}{sfx}")
                elif isinstance(pending, str):
                    f = as_fraction(pending)
                    if f is not None and (f * val).denominator == 1:
                        emit(f"{(f * val).numerator}{sfx}")
                    else:
                        flush()
                        emit(f"{val}{sfx}")
                else:
                    emit(f"{scaled(val)}{sfx}")
            elif kind == _SIGN:                    # only a sign when something numeric follows
                flush()
                if nxt_known or nxt_numeric:
                    symbol = sfx
                else:
                    emit(word)
            elif kind == _CURRENCY:                # only after a number; replaces a pending sign
                if pending is not None:
                    symbol = sfx
                    emit(pending)
                else:
                    emit(word)
            elif kind == _PERCENT:
                emit(word if pending is None else f"{pending}%")
            elif kind == _PER:
                if pending is None:
                    emit(word)
                elif nxt == "cent":
                    emit(f"{pending}%")
                    skip = True
                else:
                    flush()
                    emit(word)
            elif not (nxt_known or nxt_numeric):   # and / double / triple / point in front of a non-number: plain words
                flush()
                emit(word)
            elif kind == _AND:                     # "two hundred and five": dropped after a multiplier
                if prev_kind != _MULT:
                    flush()
                    emit(word)
            elif kind == _REPEAT:
                if nxt_known and lex[nxt][0] in (_UNIT, _ZERO):
                    pending = digits_so_far() + str(lex[nxt][1]) * val
                    skip = True
                else:
                    flush()
                    emit(word)
            else:                                  # _POINT: kept only if a digit word / number follows, else dropped
                if nxt_numeric or (nxt_known and lex[nxt][0] in (_ZERO, _UNIT, _TEN)):
                    pending = digits_so_far() + "."

        flush()
        yield from emitted

    # -- pass 3: "$2 and ¢7" -> "$2.07", "$0.07" -> "¢7", "1" -> "one" (:462-483)
    def postprocess(self, s: str) -> str:
        s = re.sub(r"([€£$])([0-9]+) (?:and )?¢([0-9]{1,2})\b", lambda m: f"{m.group(1)}{m.group(2)}.{int(m.group(3)):02d}", s)
        s = re.sub(r"[€£$]0.([0-9]{1,2})\b", lambda m: f"¢{int(m.group(1))}", s)
        return re.sub(r"\b1(s?)\b", r"one\1", s)

    def __call__(self, s: str) -> str:
        return self.postprocess(" ".join(self.process_words(self.preprocess(s).split())))


class EnglishSpellingNormalizer:
    """British -> American spelling, word by word (english_normalizer.py:494-505).  The table is the checkpoint's
    `normalizer.json` (≈1700 entries); it is data, not shipped with this repo."""

    def __init__(self, english_spelling_mapping: Dict[str, str]):
        self.mapping = english_spelling_mapping

    def __call__(self, s: str) -> str:
        return " ".join(self.mapping.get(w, w) for w in s.split())


_CONTRACTIONS = (                                   # whole-word rewrites, applied in this order (:511-526)
    ("won't", "will not"), ("can't", "can not"), ("let's", "let us"), ("ain't", "aint"), ("y'all", "you all"),
    ("wanna", "want to"), ("gotta", "got to"), ("gonna", "going to"), ("i'ma", "i am going to"), ("imma", "i am going to"),
    ("woulda", "would have"), ("coulda", "could have"), ("shoulda", "should have"), ("ma'am", "madam"))
_TITLES = (                                         # abbreviations; a space is appended to the expansion (:527-548)
    ("mr", "mister"), ("mrs", "missus"), ("st", "saint"), ("dr", "doctor"), ("prof", "professor"), ("capt", "captain"),
    ("gov", "governor"), ("ald", "alderman"), ("gen", "general"), ("sen", "senator"), ("rep", "representative"),
    ("pres", "president"), ("rev", "reverend"), ("hon", "honorable"), ("asst", "assistant"), ("assoc", "associate"),
    ("lt", "lieutenant"), ("col", "colonel"), ("jr", "junior"), ("sr", "senior"), ("esq", "esquire"))
_CLITICS = (                                        # suffix rewrites: perfect tenses first, then the general ones (:549-565)
    ("'d been", " had been"), ("'s been", " has been"), ("'d gone", " had gone"), ("'s gone", " has gone"),
    ("'d done", " had done"), ("'s got", " has got"),
    ("n't", " not"), ("'re", " are"), ("'s", " is"), ("'d", " would"), ("'ll", " will"), ("'t", " not"), ("'ve", " have"),
    ("'m", " am"))


class EnglishTextNormalizer:
    """The normaliser the reference's WER script applies to hypotheses and references (cal_wer.py:279-285;
    english_normalizer.py:508-595): fillers out, contractions and titles expanded, numbers to digits, spelling table,
    symbols and diacritics stripped."""

    def __init__(self, english_spelling_mapping: Optional[Dict[str, str]] = None):
        rules = [(rf"\b{re.escape(w)}\b", to) for w, to in _CONTRACTIONS]
        rules += [(rf"\b{w}\b", to + " ") for w, to in _TITLES]
        rules += [(rf"{w}\b", to) for w, to in _CLITICS]
        self.rules = [(re.compile(p), to) for p, to in rules]
        self.fillers = re.compile(r"\b(hmm|mm|mhm|mmm|uh|um)\b")
        self.standardize_numbers = EnglishNumberNormalizer()
        self.standardize_spellings = EnglishSpellingNormalizer(english_spelling_mapping or {})

    def __call__(self, s: str) -> str:
        s = _PARENTHESISED.sub("", _BRACKETED.sub("", s.lower()))
        s = self.fillers.sub("", s)
        s = re.sub(r"\s+'", "'", s)                 # "it 's" -> "it's"
        for pattern, to in self.rules:
            s = pattern.sub(to, s)
        s = re.sub(r"(\d),(\d)", r"\1\2", s)        # thousands separators
        s = re.sub(r"\.([^0-9]|$)", r" \1", s)      # full stops that are not decimal points
        s = strip_symbols_and_diacritics(s, keep=".%$¢€£")
        s = self.standardize_spellings(self.standardize_numbers(s))
        s = re.sub(r"[.$¢€£]([^0-9])", r" \1", s)   # symbols left without a number
        s = re.sub(r"([^0-9])%", r"\1 ", s)
        return _SPACES.sub(" ", s)


def _edit_distance(ref: Sequence[str], hyp: Sequence[str]) -> int:
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def wer(references: Union[str, Sequence[str]], hypotheses: Union[str, Sequence[str]]) -> float:
    """Corpus word error rate as ``jiwer.wer(reference, hypothesis)`` (cal_wer.py:286): total word-level edit distance
    (substitutions + deletions + insertions) over all sentence pairs / total reference words; words are whitespace-separated.
    Like jiwer, a reference without any word is an error (the rate of that pair would be undefined)."""
    if isinstance(references, str):
        references = [references]
    if isinstance(hypotheses, str):
        hypotheses = [hypotheses]
    if len(references) != len(hypotheses):
        raise ValueError("references and hypotheses differ in length")
    edits = words = 0
    for r, h in zip(references, hypotheses):
        rw, hw = r.split(), h.split()
        if not rw:
            raise ValueError("one or more references are empty strings")
        edits += _edit_distance(rw, hw)
        words += len(rw)
    if words == 0:
        raise ValueError("no reference words")
    return edits / words

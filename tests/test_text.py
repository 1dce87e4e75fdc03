"""Token->text and WER: known-answer tests of the byte-level BPE decode, the text normalisers
(against outputs of the reference's own english_normalizer.py, recorded by oracle/make_golden_text.py) and the corpus WER
(jiwer's documented examples)."""
import json
import os

import pytest

from whisper_trtllm_b200.text import (BasicTextNormalizer, EnglishNumberNormalizer, EnglishSpellingNormalizer,
                                      EnglishTextNormalizer, WhisperDetokenizer, bytes_to_unicode, wer)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "english_normalizer.json")


def _toy_vocab():
    b2u = bytes_to_unicode()
    enc = lambda s: "".join(b2u[b] for b in s.encode("utf-8"))
    toks = [" Hello", " world", "!", " caf", "é", " na", "ï", "ve", " 😀"[:1], "😀"]
    vocab = {enc(t): i for i, t in enumerate(toks)}
    # a multi-byte character split across two tokens (byte-level BPE does this)
    smile = "😀".encode("utf-8")
    vocab["".join(b2u[b] for b in smile[:2])] = 20
    vocab["".join(b2u[b] for b in smile[2:])] = 21
    vocab["<|endoftext|>"] = 50256
    vocab["<|startoftranscript|>"] = 50257
    vocab["<|notimestamps|>"] = 50362
    return vocab


def test_bytes_to_unicode_is_a_bijection():
    m = bytes_to_unicode()
    assert len(m) == 256 and len(set(m.values())) == 256 and m[ord("A")] == "A" and m[ord(" ")] == "Ġ"


def test_decode_skips_specials_and_joins_split_bytes():
    tok = WhisperDetokenizer(_toy_vocab())
    ids = [50257, 50362, 0, 1, 2, 3, 4, 20, 21, 50256, 50256]
    assert tok.decode(ids) == " Hello world! café😀"
    assert tok.decode(ids, skip_special_tokens=False).startswith("<|startoftranscript|><|notimestamps|> Hello")
    assert tok.batch_decode([[0, 1], [3, 4, 50256]]) == [" Hello world", " café"]
    with pytest.raises(KeyError):
        tok.decode([999])


def test_basic_normalizer_known_answers():
    n = BasicTextNormalizer()
    # ends are not stripped: one space survives on either side (english_normalizer.py:91)
    assert n("  Hello, World!  (applause) [MUSIC] It's 5 o'clock… ") == " hello world it s 5 o clock "
    assert BasicTextNormalizer(remove_diacritics=True)("Café Ærø") == "cafe aero"
    assert BasicTextNormalizer(split_letters=True)("ab c") == "a b c"


def test_english_normalizer_known_answers():
    n = EnglishTextNormalizer({"colour": "color"})
    assert n("Mr. Smith won't pay twenty dollars and seven cents!") == "mister smith will not pay $20.07"
    assert n("one oh one, the nineteen sixties, two and a half percent") == "101 the 1960s 2.5%"
    assert n("The colour (sic) um is three point one four") == "the color is 3.14"
    assert EnglishTextNormalizer()("no table given: colour") == "no table given colour"     # cal_wer.py:279 passes none
    assert EnglishNumberNormalizer()("two hundred and first of one") == "201st of one"
    assert EnglishSpellingNormalizer({"grey": "gray"})("grey  skies") == "gray skies"


def test_normalizers_match_the_reference_goldens():
    with open(GOLDEN, encoding="utf-8") as f:
        g = json.load(f)
    n = EnglishTextNormalizer(g["spelling"])
    assert len(g["cases"]) > 1000 and len(g["basic"]) > 1000
    for text, want in g["cases"]:
        assert n(text) == want, text
    basic = {}
    for text, rd, sl, want in g["basic"]:
        b = basic.setdefault((rd, sl), BasicTextNormalizer(remove_diacritics=rd, split_letters=sl))
        assert b(text) == want, (text, rd, sl)


def test_wer_of_normalised_text():
    # cal_wer.py:279-286: both sides normalised, then corpus WER
    n = EnglishTextNormalizer()
    refs = [n("Twenty one people can't come."), n("Dr. Who's gone")]
    hyps = [n("21 people can not come"), n("doctor who has left")]
    assert wer(refs, hyps) == pytest.approx(1 / 9)


def test_wer_known_answers():
    assert wer("hello world", "hello world") == 0.0
    assert wer("hello world", "hello duck") == 0.5                    # one substitution of two words
    assert wer("hello world", "hello") == 0.5                         # one deletion
    assert wer("hello", "hello big world") == 2.0                     # two insertions / one reference word
    # corpus level: edits are pooled over sentences (jiwer semantics), not averaged per sentence
    assert wer(["a b c d", "e f"], ["a x c d", "e f g"]) == pytest.approx(2 / 6)
    with pytest.raises(ValueError):
        wer(["a"], ["a", "b"])
    with pytest.raises(ValueError, match="empty"):
        wer(["a b", "  "], ["a b", "c"])                              # jiwer: "one or more references are empty strings"
    assert wer("a b", "") == 1.0                                      # an empty hypothesis is fine: every word deleted


def test_detokenizer_matches_both_reference_tokenizers():
    """oracle/make_golden_detok.py: `decode` of the reference's WhisperTokenizer (slow: no clean-up) and WhisperTokenizerFast
    (clean_up_tokenization_spaces, the class run.py:239 gets) over a synthetic byte-level vocabulary with the full set of special
    tokens; with and without skip_special_tokens; prompts (<|startofprev|> ...) stripped as the reference does."""
    with open(os.path.join(os.path.dirname(GOLDEN), "detokenizer.json"), encoding="utf-8") as f:
        g = json.load(f)
    assert len(g["cases"]) > 400 and g["fast_clean_up"] is True and g["slow_clean_up"] is False
    differ = 0
    for clean, cols in ((False, (1, 2)), (True, (3, 4))):
        d = WhisperDetokenizer(g["vocab"], g["first_special_id"], clean_up_tokenization_spaces=clean, added_tokens=g["added_tokens"])
        for case in g["cases"]:
            assert d.decode(case[0], skip_special_tokens=True) == case[cols[0]], (clean, case[0])
            assert d.decode(case[0], skip_special_tokens=False) == case[cols[1]], (clean, case[0])
            differ += case[1] != case[3]
        # the per-call argument overrides the tokenizer's setting
        assert d.decode(g["cases"][1][0], clean_up_tokenization_spaces=True) == g["cases"][1][3]
        assert d.batch_decode([g["cases"][1][0]], clean_up_tokenization_spaces=False) == [g["cases"][1][1]]
    assert differ > 100          # the clean-up matters: the two classes disagree on a third of the cases

"""Token->text and WER: known-answer tests of the byte-level BPE decode, the basic normaliser and
the corpus WER (jiwer's documented examples)."""
import pytest

from whisper_trtllm_b200.text import BasicTextNormalizer, WhisperDetokenizer, bytes_to_unicode, wer


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


def test_normalizer():
    n = BasicTextNormalizer()
    assert n("  Hello, World!  (applause) [MUSIC] It's 5 o'clock… ") == "hello world it s 5 o clock"
    assert BasicTextNormalizer({"colour": "color"})("The Colour") == "the color"


def test_wer_known_answers():
    assert wer("hello world", "hello world") == 0.0
    assert wer("hello world", "hello duck") == 0.5                    # one substitution of two words
    assert wer("hello world", "hello") == 0.5                         # one deletion
    assert wer("hello", "hello big world") == 2.0                     # two insertions / one reference word
    # corpus level: edits are pooled over sentences (jiwer semantics), not averaged per sentence
    assert wer(["a b c d", "e f"], ["a x c d", "e f g"]) == pytest.approx(2 / 6)
    with pytest.raises(ValueError):
        wer(["a"], ["a", "b"])

#!/usr/bin/env python
"""Golden vectors for the English text normaliser (test infrastructure, not shipped).

Imports the REAL reference implementation, /root/reference/transformers/src/transformers/models/whisper/english_normalizer.py
(`EnglishTextNormalizer`, used by the reference's WER script through `whisper.normalizers`, cal_wer.py:279-285), in the build
container and records its output on a fixed list of sentences plus seeded random compositions of number words, currency /
percent words, contractions and titles:  tests/golden/english_normalizer.json = [[input, output], ...].
The spelling table is the checkpoint's `normalizer.json` (absent here): the vectors use a 6-entry table.

    python oracle/make_golden_text.py
"""
import importlib.util
import json
import os
import random

REF = "/root/reference/transformers/src/transformers/models/whisper/english_normalizer.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPELLING = {"colour": "color", "favourite": "favorite", "centre": "center", "realise": "realize", "travelling": "traveling",
            "grey": "gray"}

FIXED = [
    "Hello, World!",
    "Mr. Smith won't go to St. Louis; he can't, and he ain't sorry.",
    "I'ma tell y'all: we're gonna win, I've said it, she'd gone, it's been done.",
    "um, hmm... uh, I mean (laughs) [noise] <unk> yes",
    "The colour of the centre is grey -- my favourite!",
    "twenty one", "one hundred and twenty three", "nineteen eighty four", "two thousand and twenty three",
    "one oh one", "double five three triple zero", "three point one four one five nine", "point five", "zero point zero five",
    "five and a half", "a million and a half people", "and a half", "half and a half",
    "minus five degrees", "negative twenty", "plus seven", "positive thinking", "minus the tax",
    "twenty dollars", "twenty dollars and seven cents", "five pounds", "ten euros and fifty cents", "seven cents", "dollars and cents",
    "$20 million", "$2.50", "€3 and ¢7", "£1,000,000", "ninety nine percent", "ten per cent", "per person", "percent of people",
    "the first second third fourth fifth twelfth twentieth thirtieth hundredth thousandth millionth",
    "twenty first", "one hundred and first", "ninety ninth", "the sixties", "the nineteen sixties", "twos and threes and sixes",
    "hundreds of thousands", "two hundreds", "three millions", "1960s", "274th", "32nd", "1st place", "3rd", "21 st century",
    "one", "ones", "one one", "the one and only", "no one knows", "eleven ones",
    "a 3.5 percent rise", "1,234.56", "192.168.1.1", "3.14.15", "v2.0", "covid19", "mp3 player", "3d", "b2b",
    "thirteen hundred", "twelve thousand three hundred forty five", "one million two hundred thousand", "three billion",
    "two point five million", "zero", "oh", "o", "oh no", "oh one two", "nine one one", "one two three four five",
    "fifty fifty", "twenty twenty", "nineteen ninety nine", "eighteen twelve", "ten ten", "eleven eleven", "forty two and sixty",
    "point", "the point is", "double", "double trouble", "triple", "triple seven", "and", "and one", "hundred and one",
    "thousand", "a thousand", "million dollars", "one trillion", "two quadrillion", "hundred", "hundredth",
    "It's 5 o'clock; I'll be there @ 7:30 p.m. -- don't be late!", "50% off", "100 %", "he said: \"no.\"",
    "Dr. Jekyll & Mr. Hyde", "Gen. Lee, Col. Mustard, Lt. Dan, Sen. Smith, Rep. Jones, Gov. Brown, Pres. Lincoln, Rev. King",
    "Hon. judge, Asst. Prof. Capt. Ald. Assoc. Jr. Sr. Esq.", "ma'am, let's go", "woulda coulda shoulda wanna gotta gonna imma",
    "café naïve résumé Ærø œuvre straße Łódź Þór", "it 's", "they 're here", "n't", "'s", "rock 'n' roll",
    "seven thousand eight hundred and ninety two point four five six", "minus one point five", "plus one hundred percent",
    "sixty five dollars and three cents", "one dollar", "one cent", "one pound", "one euro", "zero dollars",
    "two and a half thousand", "three and a half", "a hundred and a half", "twenty and a half percent",
    "first of all, second thoughts", "the seventh of the ninth", "twenty second of march nineteen ninety", "two thirds",
    "one third", "three quarters", "half", "", " ", "   multiple   spaces   ", "UPPER CASE TWENTY ONE", "Twenty-One", "twenty-one",
    "forty-two point-five", "1 000 000", "1st 2nd 3rd 4th", "5 s", "10 th", "3 rd", "2 nd", "7 st",
]

WORDS = ["zero", "oh", "o", "one", "two", "three", "five", "nine", "ten", "eleven", "twelve", "fifteen", "nineteen", "twenty", "thirty",
         "forty", "ninety", "hundred", "thousand", "million", "billion", "and", "point", "double", "triple", "minus", "plus", "negative",
         "dollars", "dollar", "cents", "cent", "pounds", "euros", "percent", "per", "first", "second", "third", "fifth", "ninth", "twelfth",
         "twentieth", "hundredth", "thousandth", "twos", "sixes", "tens", "twenties", "hundreds", "thousands", "millions", "a", "half", "the",
         "cat", "of", "7", "42", "3.5", "1000", "$5", "-3", "+2", "0.25", "1st", "20s", "won't", "it's", "mr", "st", "dr"]
CHARS = "abc 123 .,'$%€£¢-+()[]<>é  "
FRAGMENTS = ["one", "two", "twenty", "hundred", "point", "and", "double", "percent", "dollars", "cents", "oh", "half", "first", "s",
             "th", "st"]
BASIC_CHARS = "abc ABC 123 .,!?'\"()[]<>éßøœ\u0303\u0301…—$%€\t\n😀\u200d👩e\u0301"


def main():
    spec = importlib.util.spec_from_file_location("ref_english_normalizer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    norm = mod.EnglishTextNormalizer(SPELLING)
    rng = random.Random(20240607)
    cases = list(FIXED)
    for _ in range(900):
        n = rng.randint(1, 7)
        cases.append(" ".join(rng.choice(WORDS) for _ in range(n)))
    for _ in range(300):          # character-level compositions: punctuation / symbols glued to number words
        cases.append("".join(rng.choice(CHARS) if rng.random() < 0.6 else rng.choice(FRAGMENTS) + rng.choice([" ", ""])
                             for _ in range(rng.randint(0, 25))))
    out = []
    for c in cases:
        try:
            out.append([c, norm(c)])
        except Exception as e:   # the reference raises on a few inputs: the port must raise too
            out.append([c, {"raises": type(e).__name__}])
    # BasicTextNormalizer (english_normalizer.py:75-93), all four (remove_diacritics, split_letters) modes
    basic = []
    texts = [c for c in FIXED if c] + ["".join(rng.choice(BASIC_CHARS) for _ in range(rng.randint(0, 30))) for _ in range(150)]
    for rd in (False, True):
        for sl in (False, True):
            b = mod.BasicTextNormalizer(remove_diacritics=rd, split_letters=sl)
            basic += [[t, rd, sl, b(t)] for t in texts]
    path = os.path.join(ROOT, "tests", "golden", "english_normalizer.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"spelling": SPELLING, "cases": out, "basic": basic}, f, ensure_ascii=False, indent=0)
    print(f"{len(out)} + {len(basic)} cases -> {path}; raised on {sum(1 for _, o in out if isinstance(o, dict))}")


if __name__ == "__main__":
    main()

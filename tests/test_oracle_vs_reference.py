"""The oracle and the host ports against the REAL reference, live — build container only (auto-skipped where /root/reference is
absent, i.e. on the GPU box; there the committed goldens stand in).  Each check runs in a fresh interpreter: the vendored
`transformers` must not meet the installed one that other tests import."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import hf_reference

pytestmark = pytest.mark.skipif(not hf_reference.available(), reason="/root/reference is not mounted")


def _run(code: str, timeout=900) -> dict:
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_restatement_equals_reference_generate_on_fresh_seeds():
    """Seeds that are NOT among the committed goldens: `generate()` of the vendored model vs oracle/whisper_ref.py — token ids
    identical, encoder output and every step's logits within 1e-5."""
    out = _run("""
import json, torch
from oracle import hf_reference as HF, synth, whisper_ref as R
res = []
for size, B, wseed, mseed, max_length in (("micro", 4, 11, 99, 64), ("tiny.en", 2, 5, 42, 40)):
    cfg = synth.make_config(size, max_length=max_length)
    sd = synth.make_weights(cfg, seed=wseed)
    mel = synth.make_mel(B, seed=mseed)
    model = HF.build_reference_model(cfg, sd)
    ref_ids = HF.reference_generate(model, mel, max_length=max_length)
    ref_enc = HF.reference_encode(model, mel)
    ref_logits, _ = HF.reference_step_logits(model, ref_enc, ref_ids, ref_ids.shape[1] - 1)
    ids, enc, logits = R.greedy(mel, sd, cfg, return_logits=True)
    res.append(dict(size=size, tokens_equal=bool(torch.equal(ref_ids, ids)), length=int(ids.shape[1]),
                    enc=float((ref_enc - enc).abs().max()), logits=max(float((a - b).abs().max()) for a, b in zip(ref_logits, logits))))
print(json.dumps(res))
""")
    for case in out:
        assert case["tokens_equal"], case
        assert case["enc"] < 1e-5 and case["logits"] < 1e-5, case
    assert [c["length"] for c in out] == [64, 40]


def test_text_ports_equal_the_reference_on_fresh_inputs():
    """english_normalizer.py and the slow tokenizer's decode, live, on inputs that are not in the golden files."""
    out = _run("""
import importlib.util, json, random, sys
from whisper_trtllm_b200 import text as T
spec = importlib.util.spec_from_file_location("refn", "/root/reference/transformers/src/transformers/models/whisper/english_normalizer.py")
ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
sp = {"colour": "color", "programme": "program"}
r, m = ref.EnglishTextNormalizer(sp), T.EnglishTextNormalizer(sp)
words = sorted(ref.EnglishNumberNormalizer().words) + "the a cat programme colour 7 42 3.5 $5 -3 1st 20s won't it's mr dr half and a half o'clock 1,000 5% (x) [y] uh".split()
rng = random.Random(987654)
bad = n = 0
for _ in range(20000):
    s = " ".join(rng.choice(words) for _ in range(rng.randint(1, 8)))
    n += 1
    bad += r(s) != m(s)
for rd in (False, True):
    rb, mb = ref.BasicTextNormalizer(rd), T.BasicTextNormalizer(rd)
    for _ in range(2000):
        s = "".join(rng.choice("abc ABC 12 .,!?'()[]<>éßø…—$%") for _ in range(rng.randint(0, 30)))
        n += 1
        bad += rb(s) != mb(s)
print(json.dumps(dict(cases=n, different=bad)))
""")
    assert out["cases"] == 24000 and out["different"] == 0


def test_logmel_restatement_equals_the_reference_feature_extractor():
    out = _run("""
import json, sys, types
import numpy as np
sys.dont_write_bytecode = True
stub = types.ModuleType("transformers.dependency_versions_check"); stub.dep_version_check = lambda *a, **k: None
sys.modules["transformers.dependency_versions_check"] = stub
sys.path.insert(0, "/root/reference/transformers/src")
from transformers.models.whisper.feature_extraction_whisper import WhisperFeatureExtractor
from oracle import logmel_ref as LM
fe = WhisperFeatureExtractor()
rng = np.random.RandomState(31337)
worst = 0.0
for n in (16000, 123457, 480000, 500001):
    w = (rng.randn(n) * 0.1 * np.abs(np.sin(np.arange(n) * 3e-4))).astype(np.float32)
    ref = fe(w, sampling_rate=16000, return_tensors="np").input_features[0]
    worst = max(worst, float(np.abs(ref - LM.log_mel(w)).max()))
print(json.dumps(dict(worst=worst)))
""")
    assert out["worst"] < 1e-5

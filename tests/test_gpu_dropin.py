"""Drop-in surface parity (-m gpu): the module classes (tensorrt_llm.models.*), Session/TensorInfo and the run.py greedy
session, all executing on the CUDA kernels through the C-ABI, against the CPU oracle (oracle/whisper_ref.py, pinned to
the real reference by tests/golden) on the same seeded inputs.

fp32: tokens identical, tensors within 1e-3 relative (north_star).  bf16: stated tolerance 3e-2 relative to max."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import synth, whisper_ref as R
from oracle.make_golden import CASES
from whisper_trtllm_b200 import models, run, runtime
from whisper_trtllm_b200.model import decoder_from_config, encoder_from_config, load_decoder_from_hf, load_encoder_from_hf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL, BF16_TOL = 1e-3, 3e-2


def _rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def tiny():
    cfg = synth.make_config("tiny.en", max_length=24)
    sd = synth.make_weights(cfg, seed=0)
    mel = synth.make_mel(2, seed=1234)
    with torch.no_grad():
        enc = R.encode(mel, sd, cfg)
    return cfg, sd, mel, enc


@pytest.mark.parametrize("dtype,tol", [("float32", FP32_TOL), ("bfloat16", BF16_TOL)])
def test_whisper_encoder_module(tiny, dtype, tol):
    cfg, sd, mel, ref_enc = tiny
    enc = load_encoder_from_hf(encoder_from_config(cfg, dtype=dtype), sd)
    out = enc(mel.to(DEV))
    assert out.dtype == torch.float32 and tuple(out.shape) == (2, 1500, cfg["d_model"])
    assert _rel(out, ref_enc) < tol


@pytest.mark.parametrize("dtype,tol", [("float32", FP32_TOL), ("bfloat16", BF16_TOL)])
def test_decoder_attention_four_modes(tiny, dtype, tol):
    """self/cross x with/without cache; the cache length is carried by cache_mask.shape[0]-1, values never read
    (model.py:261-281; oracle branches modeling_whisper.py:474-503)."""
    cfg, sd, _, ref_enc = tiny
    H, d = cfg["decoder_attention_heads"], cfg["d_model"]
    dec = load_decoder_from_hf(decoder_from_config(cfg, dtype=dtype), sd)
    layer = dec.layers[1]
    p = "model.decoder.layers.1"
    g = torch.Generator().manual_seed(3)
    h1, h2, h3 = (torch.randn(2, 1, d, generator=g) for _ in range(3))
    nan_mask = lambda n: torch.full((n,), float("nan"), device=DEV)   # values must never be read

    # --- self / no cache (step 0: dummy past of length 1, mask of length 1 -> cache length 0, run.py:113-119)
    r1, (rk1, rv1) = R.decoder_attention(h1, sd, p + ".self_attn", H)
    dummy = torch.full((2, H, 1, 64), float("nan"), device=DEV)
    o1, k1, v1 = layer.self_attn(h1.to(DEV), None, dummy, dummy, nan_mask(1))
    assert tuple(k1.shape) == (2, H, 1, 64) and _rel(o1, r1) < tol and _rel(k1, rk1) < tol and _rel(v1, rv1) < tol
    # past_key None behaves the same
    o1b, k1b, _ = layer.self_attn(h1.to(DEV))
    assert torch.equal(o1b, o1) and torch.equal(k1b, k1)
    # --- self / cache: mask of length n+1 -> reuse n cached rows and append
    r2, (rk2, rv2) = R.decoder_attention(h2, sd, p + ".self_attn", H, past=(rk1, rv1))
    o2, k2, v2 = layer.self_attn(h2.to(DEV), None, k1, v1, nan_mask(2))
    assert tuple(k2.shape) == (2, H, 2, 64) and _rel(o2, r2) < tol and _rel(k2, rk2) < tol
    r3, (rk3, _) = R.decoder_attention(h3, sd, p + ".self_attn", H, past=(rk2, rv2))
    o3, k3, v3 = layer.self_attn(h3.to(DEV), None, k2, v2, nan_mask(3))
    assert tuple(k3.shape) == (2, H, 3, 64) and _rel(o3, r3) < tol and _rel(k3, rk3) < tol
    # the mask SHAPE decides: a shorter mask uses fewer cached rows (min(len(mask)-1, past.shape[2]), model.py:278)
    o2s, k2s, _ = layer.self_attn(h2.to(DEV), None, k3, v3, nan_mask(2))
    assert tuple(k2s.shape) == (2, H, 2, 64) and _rel(o2s, r2) < tol
    # --- cross / no cache: mask of length 1 -> project the encoder states
    c1, (rck, rcv) = R.decoder_attention(h1, sd, p + ".encoder_attn", H, ref_enc)
    big_dummy = torch.full((2, H, 1500, 64), float("nan"), device=DEV)
    oc1, ck, cv = layer.encoder_attn(h1.to(DEV), ref_enc.to(DEV), big_dummy, big_dummy, nan_mask(1))
    assert tuple(ck.shape) == (2, H, 1500, 64) and _rel(oc1, c1) < tol and _rel(ck, rck) < tol and _rel(cv, rcv) < tol
    # --- cross / cache: mask of length 1501 -> reuse, and the SAME storage comes back (no re-copy)
    c2, _ = R.decoder_attention(h2, sd, p + ".encoder_attn", H, ref_enc, past=(rck, rcv))
    oc2, ck2, cv2 = layer.encoder_attn(h2.to(DEV), ref_enc.to(DEV), ck, cv, nan_mask(1501))
    assert ck2.data_ptr() == ck.data_ptr() and cv2.data_ptr() == cv.data_ptr() and _rel(oc2, c2) < tol


def _run_module_steps(dec, enc_dev, ids, steps, batched):
    """Teacher-forced decode through WhisperDecoder.forward with the runner's tensor contract (run.py:105-126)."""
    L, H = dec.decoder_layers, dec.decoder_attention_heads
    B = ids.shape[0]
    lead = (L, B, H) if batched else (L, H)
    sk = torch.rand(*lead, 1, 64, device=DEV); sv = torch.rand(*lead, 1, 64, device=DEV)
    ck = torch.rand(*lead, 1500, 64, device=DEV); cv = torch.rand(*lead, 1500, 64, device=DEV)
    ms, mc = torch.rand(1, device=DEV), torch.rand(1, device=DEV)
    out = []
    for t in range(steps):
        logits, sk, sv, ck, cv = dec(ids[:, t:t + 1].to(DEV), enc_dev, sk, sv, ck, cv, ms, mc)
        out.append(logits[:, -1, :].float().cpu())
        ms, mc = torch.rand(1 + sk.shape[-2], device=DEV), torch.rand(1 + 1500, device=DEV)
    return out, (sk, sv, ck, cv)


@pytest.mark.parametrize("dtype,tol", [("float32", FP32_TOL), ("bfloat16", BF16_TOL)])
def test_whisper_decoder_module_teacher_forced(tiny, dtype, tol):
    cfg, sd, mel, ref_enc = tiny
    ref_ids, _, ref_logits = R.greedy(mel, sd, cfg, return_logits=True)
    dec = load_decoder_from_hf(decoder_from_config(cfg, dtype=dtype), sd)
    steps = 8
    logits, (sk, sv, ck, cv) = _run_module_steps(dec, ref_enc.to(DEV), ref_ids, steps, batched=True)
    for s in range(steps):
        assert _rel(logits[s], ref_logits[s]) < tol, f"step {s}"
    L, H = cfg["decoder_layers"], cfg["decoder_attention_heads"]
    assert tuple(sk.shape) == (L, 2, H, steps, 64) and tuple(ck.shape) == (L, 2, H, 1500, 64)
    # caches equal the oracle's past_key_values (modeling_whisper.py:494-495)
    past = None
    for t in range(steps):
        _, past = R.decoder_forward(ref_ids[:, t:t + 1], ref_enc, sd, cfg, past)
    assert _rel(sk[L - 1], past[-1][0]) < tol and _rel(cv[0], past[0][3]) < tol
    if dtype == "float32":    # reference layout [L, H, T, 64] for one utterance gives the same numbers
        l1, (sk1, _, _, _) = _run_module_steps(dec, ref_enc[:1].to(DEV), ref_ids[:1], 4, batched=False)
        assert tuple(sk1.shape) == (L, H, 4, 64)
        for s in range(4):
            assert _rel(l1[s], ref_logits[s][:1]) < tol


def test_session_run_with_raw_pointers(tiny):
    """Session.run(inputs, outputs, stream): caller-owned buffers passed as raw addresses, async on a raw cudaStream_t
    (session.py:148-178, run.py:35-46)."""
    cfg, sd, mel, ref_enc = tiny
    T, f32 = runtime.TensorInfo, runtime.DataType.float32
    sess = runtime.Session.from_serialized_engine(run.build_encoder(cfg, sd))
    data = mel.to(DEV)
    length = torch.ones(2, device=DEV)
    infos = sess.infer_shapes([T("data", f32, tuple(data.shape)), T("length", f32, (2,))])
    out = torch.zeros(*infos[0].shape, device=DEV)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    ok = sess.run({"data": data.data_ptr(), "length": length.data_ptr()}, {"hidden_states": out.data_ptr()}, stream.cuda_stream)
    assert ok is True
    stream.synchronize()
    assert _rel(out, ref_enc) < FP32_TOL


def test_runner_generic_and_fast_paths_match_reference_tokens(tmp_path):
    """run.py's session: engine files on disk -> runner classes -> greedy_search.  The per-step Session path (generic)
    and the on-device loop (fast) both reproduce the REAL reference's token ids (golden 'tiny')."""
    meta, g = load_golden("tiny")
    size, B, wseed, mseed, _, _ = CASES["tiny"]
    steps = 16
    cfg = synth.make_config(size, max_length=steps + 1)
    sd = synth.make_weights(cfg, seed=wseed)
    mel = synth.make_mel(B, seed=mseed).to(DEV)
    want = g["tokens"][:, :steps + 1]

    class Args:
        engine_dir = str(tmp_path)
    run.build_encoder(cfg, sd, Args.engine_dir, "float32")
    run.build_decoder(cfg, sd, Args.engine_dir, "float32")
    import pickle, os
    with open(os.path.join(Args.engine_dir, "config.pkl"), "rb") as f:
        config = pickle.load(f)
    whisperencoder = run.WhisperEncoder(args=Args, config=config)
    whisperdecoder = run.WhisperDecoder(args=Args, config=config)

    def transcribe():   # run.py:270-284
        encoder_outputs = whisperencoder(mel)
        input_ids = torch.full((B, 1), config["decoder_start_token_id"], dtype=torch.int32, device=DEV)
        return run.greedy_search(model=whisperdecoder, encoder_outputs=encoder_outputs, input_ids=input_ids,
                                 logits_processor=run.get_logits_processor(config, input_ids.shape[-1]),
                                 stopping_criteria=run.get_stopping_criteria(config),
                                 pad_token_id=config["pad_token_id"], eos_token_id=config["eos_token_id"])
    ids_generic = transcribe()
    assert np.array_equal(ids_generic.cpu().numpy(), want)
    fast = run.link(whisperencoder, whisperdecoder, config, max_batch=B)
    launches0 = fast.engine.launch_count()
    ids_fast = transcribe()
    assert fast.engine.launch_count() > launches0
    assert ids_fast.dtype == torch.int32 and np.array_equal(ids_fast.cpu().numpy(), want)
    # encoder states produced elsewhere (not by the linked encoder): the fast path re-projects the cross K/V
    enc_out = whisperencoder.session.engine(mel)
    ids2 = run.greedy_search(whisperdecoder, enc_out, torch.full((B, 1), config["decoder_start_token_id"], dtype=torch.int32, device=DEV),
                             run.get_logits_processor(config, 1), run.get_stopping_criteria(config), config["pad_token_id"],
                             config["eos_token_id"])
    assert np.array_equal(ids2.cpu().numpy(), want)
    # a non-standard processor list falls back to the generic loop and honours the callable
    class BanToken:
        def __init__(self, tok): self.tok = tok
        def __call__(self, input_ids, scores):
            scores[:, self.tok] = -float("inf")
            return scores
    banned = int(want[0, 2])
    procs = run.get_logits_processor(config, 1)
    procs.append(BanToken(banned))
    ids3 = run.greedy_search(whisperdecoder, enc_out, torch.full((B, 1), config["decoder_start_token_id"], dtype=torch.int32, device=DEV),
                             procs, run.get_stopping_criteria(config), config["pad_token_id"], config["eos_token_id"])
    assert banned not in ids3[0, 2:].tolist() and ids3.shape[1] == steps + 1
    fast.engine.close()

"""Kernel-level parity (-m gpu): every CUDA operator, called through the C-ABI, against the same op in plain
fp32 PyTorch (the arithmetic the oracle is made of).  Mirrors the reference's one-file-per-op
style (tests/functional/test_layer_norm.py, test_gelu.py, test_matmul.py ...)."""
import pytest
import torch
import torch.nn.functional as F

import gpu_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


@pytest.mark.parametrize("rows,d", [(1, 384), (7, 128), (300, 512), (1500, 768), (4097, 1024)])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_layernorm(rows, d, odt):
    g = _gen(rows + d)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(DEV)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV)
    b = (0.1 * torch.randn(d, generator=g)).to(DEV)
    ref = F.layer_norm(x, (d,), w, b, 1e-5)
    out = G.layernorm(x, w, b, odt)
    tol = 2e-6 if odt == torch.float32 else 4e-3
    assert G.rel_err(out, ref) < tol


@pytest.mark.parametrize("M,N,K", [(1, 384, 384), (16, 512, 2048), (100, 1152, 384), (130, 51864, 128), (1500, 1536, 384), (3000, 384, 256)])
@pytest.mark.parametrize("act,res", [(0, False), (1, False), (0, True)])
def test_linear_fp32_cuda_core(M, N, K, act, res):
    g = _gen(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    R = torch.randn(M, N, generator=g).to(DEV) if res else None
    ref = F.linear(A, W, bias)
    if act:
        ref = F.gelu(ref)
    if res:
        ref = ref + R
    out = G.linear(A, W, bias, R, act, torch.float32, backend=1)
    assert G.rel_err(out, ref) < 1e-5   # fp32 FFMA accumulation; differences are summation order only


@pytest.mark.parametrize("bn", [0, 32, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 1024, 1024), (1500, 1152, 384), (77, 512, 2048), (300, 51864, 384), (3001, 384, 256)])
def test_linear_bf16_tcgen05(M, N, K, bn):
    """tcgen05/TMEM/TMA GEMM vs fp32 torch on the SAME bf16-rounded operands (so only accumulation differs)."""
    g = _gen(M * 3 + N + K)
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    ref = F.linear(A.float(), W.float(), bias)
    out = G.linear(A, W, bias, None, 0, torch.float32, backend=2 if bn == 0 else 100 + bn)
    assert G.rel_err(out, ref) < 2e-5
    out16 = G.linear(A, W, bias, None, 0, torch.bfloat16, backend=2 if bn == 0 else 100 + bn)
    assert G.rel_err(out16, ref) < 6e-3


def test_linear_bf16_tcgen05_epilogues():
    g = _gen(5)
    M, N, K = 700, 1536, 384
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    R = torch.randn(M, N, generator=g).to(DEV)
    lin = F.linear(A.float(), W.float(), bias)
    assert G.rel_err(G.linear(A, W, bias, None, 1, torch.float32, backend=2), F.gelu(lin)) < 1e-5   # erf GELU
    assert G.rel_err(G.linear(A, W, bias, R, 0, torch.float32, backend=2), lin + R) < 1e-5
    # in-place residual (out aliases the residual): the encoder's x += out_proj(...)
    x = R.clone()
    from whisper_trtllm_b200 import _abi
    _abi.call("wb_linear", G.ptr(A), K, G.ptr(W), K, _abi.BF16, G.ptr(bias), G.ptr(x), N, G.ptr(x), N, _abi.F32, M, N, K, 0, 2,
              G.stream_handle())
    assert G.rel_err(x, lin + R) < 1e-5


def test_linear_bf16_cuda_core_matches_tcgen05():
    g = _gen(9)
    M, N, K = 257, 768, 768
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    a = G.linear(A, W, None, None, 0, torch.float32, backend=1)
    b = G.linear(A, W, None, None, 0, torch.float32, backend=2)
    assert G.rel_err(a, b) < 2e-5


def _attn_ref(qkv, B, S, H):
    d = H * 64
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    w = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    return (w @ v).permute(0, 2, 1, 3).reshape(B * S, d)


@pytest.mark.parametrize("B,S,H", [(1, 64, 2), (2, 200, 6), (1, 1500, 8)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_encoder_attention_cuda_core(B, S, H, dt):
    g = _gen(B + S + H)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).to(DEV)
    qkv[:, :H * 64] *= 0.125 * 3
    qkv = qkv.to(dt)
    ref = _attn_ref(qkv, B, S, H)
    out = G.encoder_attention(qkv, B, S, H, backend=1)
    assert G.rel_err(out, ref) < (2e-6 if dt == torch.float32 else 6e-3)


@pytest.mark.parametrize("B,S,H", [(1, 64, 1), (1, 128, 2), (2, 200, 6), (3, 1500, 8), (2, 1500, 16)])
def test_encoder_attention_tcgen05(B, S, H):
    """tcgen05 flash kernel (bf16 in, fp32 softmax/accumulate) vs fp32 torch on the same bf16 inputs.
    B > 1 with S % 64 != 0 exercises the tail mask (the last key tile reaches into the next utterance)."""
    g = _gen(B * 7 + S + H)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).to(DEV)
    qkv[:, :H * 64] *= 0.125 * 3
    qkv = qkv.to(torch.bfloat16)
    ref = _attn_ref(qkv, B, S, H)
    out = G.encoder_attention(qkv, B, S, H, backend=2)
    assert torch.isfinite(out.float()).all()
    assert G.rel_err(out, ref) < 8e-3          # P is rounded to bf16 before the P V product
    simt = G.encoder_attention(qkv, B, S, H, backend=1)
    assert G.rel_err(out, simt) < 8e-3


@pytest.mark.parametrize("B,H,T,n", [(1, 6, 448, 1), (3, 8, 448, 77), (2, 16, 1500, 1500), (5, 2, 1500, 1499)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_decode_attention(B, H, T, n, dt):
    g = _gen(B + H + T + n)
    q = (torch.randn(B, H * 64, generator=g) * 0.3).to(DEV).to(dt)
    k = torch.randn(B, H, T, 64, generator=g).to(DEV).to(dt)
    v = torch.randn(B, H, T, 64, generator=g).to(DEV).to(dt)
    w = torch.softmax(torch.einsum("bhd,bhtd->bht", q.float().view(B, H, 64), k.float()[:, :, :n]), dim=-1)
    ref = torch.einsum("bht,bhtd->bhd", w, v.float()[:, :, :n]).reshape(B, H * 64)
    out = G.decode_attention(q, k, v, n)
    assert G.rel_err(out, ref) < (2e-6 if dt == torch.float32 else 6e-3)


def test_masked_argmax_first_max_and_masks():
    V = 51864
    g = _gen(1)
    x = torch.randn(4, V, generator=g).to(DEV)
    x[0, 100] = 50.0
    x[0, 7] = 60.0            # masked below
    x[1, 5000] = 42.0
    x[1, 40000] = 42.0        # tie -> first index wins (torch.argmax semantics)
    x[2, V - 1] = 99.0
    mask = torch.zeros(V, dtype=torch.uint8, device=DEV)
    mask[7] = 1
    mask[V - 1] = 2
    out = G.argmax(x, mask, 1)
    xm = x.clone()
    xm[:, 7] = -float("inf")
    assert out.tolist() == torch.argmax(xm, -1).tolist()
    assert out[1].item() == 5000 and out[0].item() == 100 and out[2].item() == V - 1
    out2 = G.argmax(x, mask, 3)
    xm[:, V - 1] = -float("inf")
    assert out2.tolist() == torch.argmax(xm, -1).tolist()


@pytest.mark.parametrize("M,N,K,splits", [(256, 1024, 1024, 0), (256, 1024, 4096, 0), (256, 1024, 1024, 4), (64, 384, 1536, 2), (31, 512, 512, 8)])
def test_linear_splitk_deferred_reduction_and_layernorm_preadd(M, N, K, splits):
    """split-K tcgen05 GEMM -> raw fp32 slabs; the consumer (LayerNorm pre-add) adds bias + slabs + residual in a fixed order."""
    import ctypes
    from whisper_trtllm_b200 import _abi
    g = _gen(M + N + K + splits)
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    x = torch.randn(M, N, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(N, generator=g)).to(DEV)
    parts = torch.full((8, M, N), float("nan"), device=DEV)
    chosen = ctypes.c_int(0)
    _abi.call("wb_linear_splitk", G.ptr(A), K, G.ptr(W), K, _abi.BF16, G.ptr(parts), M * N, M, N, K, splits, 8, ctypes.byref(chosen),
              G.stream_handle())
    S = chosen.value
    assert 1 <= S <= 8 and (splits == 0 or S == splits)
    ref = F.linear(A.float(), W.float())
    assert G.rel_err(parts[:S].sum(0), ref) < 2e-5
    if S < 8:
        assert torch.isnan(parts[S:]).all()          # slabs beyond the chosen count are untouched
    # consumer: x += bias + sum(slabs); out = LN(x)
    x2 = x.clone()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    _abi.call("wb_layernorm_preadd", G.ptr(x2), G.ptr(parts), S, M * N, G.ptr(bias), G.ptr(gamma), G.ptr(beta), G.ptr(out), _abi.BF16,
              M, N, 1e-5, G.stream_handle())
    want_x = x + bias + ref
    assert G.rel_err(x2, want_x) < 2e-5
    assert G.rel_err(out, F.layer_norm(want_x, (N,), gamma, beta, 1e-5)) < 6e-3
    # determinism: same slabs, same order -> bit-identical
    x3 = x.clone()
    _abi.call("wb_layernorm_preadd", G.ptr(x3), G.ptr(parts), S, M * N, G.ptr(bias), G.ptr(gamma), G.ptr(beta), G.ptr(out), _abi.BF16,
              M, N, 1e-5, G.stream_handle())
    assert torch.equal(x2, x3)


def test_pdl_switch_gives_identical_results():
    """Programmatic dependent launch only changes WHEN a kernel's prologue runs, never its result."""
    from whisper_trtllm_b200 import _abi
    g = _gen(77)
    x = torch.randn(256, 1024, generator=g).to(DEV)
    w = (1 + 0.1 * torch.randn(1024, generator=g)).to(DEV)
    b = (0.1 * torch.randn(1024, generator=g)).to(DEV)
    W = (torch.randn(1024, 1024, generator=g) / 32).to(DEV).to(torch.bfloat16)
    outs = []
    for flag in (0, 1):
        _abi.call("wb_set_pdl", flag)
        h = x
        for _ in range(6):       # a chain of dependent kernels: LN -> GEMM(+residual) -> LN -> ...
            ln = G.layernorm(h, w, b, torch.bfloat16)
            h = G.linear(ln, W, b, h, 0, torch.float32, backend=2)
        outs.append(h.clone())
    _abi.call("wb_set_pdl", 0)
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()


@pytest.mark.parametrize("B,H,T,n", [(1, 6, 448, 1), (3, 8, 448, 77), (2, 16, 1500, 1500), (5, 2, 1500, 1499), (40, 16, 1500, 1500)])
def test_decode_attention_bulk_ring_kernel(B, H, T, n):
    """cp.async.bulk ring variant of the cross-attention kernel == the 16-byte-load kernel == fp32 torch."""
    from whisper_trtllm_b200 import _abi
    g = _gen(B + H + T + n + 1)
    q = (torch.randn(B, H * 64, generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    k = torch.randn(B, H, T, 64, generator=g).to(DEV).to(torch.bfloat16)
    v = torch.randn(B, H, T, 64, generator=g).to(DEV).to(torch.bfloat16)
    w = torch.softmax(torch.einsum("bhd,bhtd->bht", q.float().view(B, H, 64), k.float()[:, :, :n]), dim=-1)
    ref = torch.einsum("bht,bhtd->bhd", w, v.float()[:, :, :n]).reshape(B, H * 64)
    base = G.decode_attention(q, k, v, n)
    _abi.call("wb_set_decode_attention_backend", 1)
    try:
        out = G.decode_attention(q, k, v, n)
        torch.cuda.synchronize()
    finally:
        _abi.call("wb_set_decode_attention_backend", 0)
    assert G.rel_err(out, ref) < 6e-3
    assert G.rel_err(out, base) < 6e-3


def _paged_self_attention_case(B, H, n, dtype, seed, finished=()):
    """Random fused qkv rows + a paged cache holding n - 1 tokens per row behind a SHUFFLED page table; the op appends row n - 1
    in place and attends over n keys.  -> (out, fp32 torch reference, pages after, expected appended rows)."""
    from whisper_trtllm_b200 import _abi
    g = _gen(seed)
    d, pps = H * 64, 7
    qkv = (torch.randn(B, 3 * d, generator=g) * 0.5).to(DEV).to(dtype)
    num_pages = B * pps
    k_pages = torch.randn(num_pages, H, 64, 64, generator=g).to(DEV).to(dtype)
    v_pages = torch.randn(num_pages, H, 64, 64, generator=g).to(DEV).to(dtype)
    table = torch.randperm(num_pages, generator=g).to(torch.int32).view(B, pps).to(DEV)
    state = torch.zeros(8, dtype=torch.int32, device=DEV)
    state[0], state[1] = n, 1
    row_active = torch.ones(B, dtype=torch.int32, device=DEV)
    for r in finished:
        row_active[r] = 0
    out = torch.full((B, d), float("nan"), device=DEV).to(dtype)
    k_before, v_before = k_pages.clone(), v_pages.clone()
    _abi.call("wb_paged_self_attention", G.ptr(qkv), 3 * d, G.ptr(out), G.ptr(k_pages), G.ptr(v_pages), G.ptr(table), pps,
              G.DT[dtype], B, H, G.ptr(state), G.ptr(row_active), G.stream_handle())
    torch.cuda.synchronize()
    # dense reference: gather the pages, put the new row at slot n - 1
    def dense(pages):
        x = pages[table.long()]                                     # [B, pps, H, 64, 64]
        return x.permute(0, 2, 1, 3, 4).reshape(B, H, pps * 64, 64).float()
    K, V = dense(k_before), dense(v_before)
    q = qkv[:, :d].float().view(B, H, 64)
    K[:, :, n - 1] = qkv[:, d:2 * d].float().view(B, H, 64)
    V[:, :, n - 1] = qkv[:, 2 * d:].float().view(B, H, 64)
    w = torch.softmax(torch.einsum("bhd,bhtd->bht", q, K[:, :, :n]), dim=-1)
    ref = torch.einsum("bht,bhtd->bhd", w, V[:, :, :n]).reshape(B, d)
    return out, ref, (k_pages, v_pages, k_before, v_before, table, qkv)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 128, 224, 447])
@pytest.mark.parametrize("B,H", [(32, 16), (64, 16), (40, 12)])
def test_paged_self_attention_warp_per_item_bf16(B, H, n):
    """The kernel the headline configuration launches for the cached self-attention: self_attn_warp_kernel<bf16, 4> (one warp per
    (utterance, head) item; taken when B * H exceeds two items per SM) against fp32 torch on the same bf16 inputs: lengths at
    and around the page boundaries, 1 key, the maximum; the appended row lands in its page and nothing else is touched."""
    import ctypes
    from whisper_trtllm_b200 import _abi
    sms = ctypes.c_int()
    _abi.call("wb_device_info", ctypes.byref(sms), ctypes.byref(ctypes.c_int()), ctypes.byref(ctypes.c_int()))
    assert B * H > 2 * sms.value, "this case must take the warp-per-item kernel"
    out, ref, (kp, vp, kb, vb, table, qkv) = _paged_self_attention_case(B, H, n, torch.bfloat16, 1000 + n + B)
    assert torch.isfinite(out.float()).all()
    assert G.rel_err(out, ref) < 8e-3, (B, H, n)
    d = H * 64
    page = table[:, (n - 1) // 64].long()
    slot = (n - 1) % 64
    assert torch.equal(kp[page, :, slot], qkv[:, d:2 * d].view(B, H, 64))
    assert torch.equal(vp[page, :, slot], qkv[:, 2 * d:].view(B, H, 64))
    kb[page, :, slot] = kp[page, :, slot]
    vb[page, :, slot] = vp[page, :, slot]
    assert torch.equal(kp, kb) and torch.equal(vp, vb)


def test_paged_self_attention_skips_finished_rows_and_matches_the_cta_kernel():
    """Rows whose utterance has emitted EOS are skipped (output untouched, nothing appended); the warp-per-item kernel and the
    CTA-per-item kernel (small item counts) agree on the rows they both compute; fp32 instance within 1e-5."""
    out, ref, (kp, vp, kb, vb, table, qkv) = _paged_self_attention_case(48, 16, 130, torch.bfloat16, 5, finished=(0, 17, 47))
    live = [r for r in range(48) if r not in (0, 17, 47)]
    assert torch.isnan(out[[0, 17, 47]].float()).all()
    assert G.rel_err(out[live], ref[live]) < 8e-3
    for r in (0, 17, 47):
        pg = table[r].long()
        assert torch.equal(kp[pg], kb[pg]) and torch.equal(vp[pg], vb[pg])
    small, ref_s, _ = _paged_self_attention_case(4, 16, 130, torch.bfloat16, 6)        # 64 items: CTA per item
    assert G.rel_err(small, ref_s) < 8e-3
    o32, r32, _ = _paged_self_attention_case(32, 16, 200, torch.float32, 7)            # fp32 warp kernel
    assert G.rel_err(o32, r32) < 1e-5


@pytest.mark.parametrize("B,H,n", [(256, 16, 1500), (64, 16, 1500), (24, 16, 1500), (40, 12, 1500), (32, 16, 777)])
def test_cross_attention_large_batch_kernel_bf16(B, H, n):
    """decode_attn_kernel<bf16, false, 128, 8>: the cross-attention kernel of the headline configuration (persistent, 4 CTAs per
    SM, taken when B * H exceeds two items per SM) asserted DIRECTLY against fp32 torch on the same bf16 K/V."""
    g = _gen(B + H + n)
    q = (torch.randn(B, H * 64, generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    k = torch.randn(B, H, 1500, 64, generator=g).to(DEV).to(torch.bfloat16)
    v = torch.randn(B, H, 1500, 64, generator=g).to(DEV).to(torch.bfloat16)
    out = G.decode_attention(q, k, v, n)
    torch.cuda.synchronize()
    ref = torch.empty(B, H * 64, device=DEV)
    for b0 in range(0, B, 32):     # chunked: the fp32 copy of K/V of 256 utterances would not be small
        kk, vv = k[b0:b0 + 32, :, :n].float(), v[b0:b0 + 32, :, :n].float()
        w = torch.softmax(torch.einsum("bhd,bhtd->bht", q[b0:b0 + 32].float().view(-1, H, 64), kk), dim=-1)
        ref[b0:b0 + 32] = torch.einsum("bht,bhtd->bhd", w, vv).reshape(-1, H * 64)
    assert G.rel_err(out, ref) < 6e-3

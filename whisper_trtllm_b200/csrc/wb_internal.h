// Internal launcher declarations shared by the runtime (runtime.cu) and the C-ABI (abi.cu).
#pragma once
#include "wb_common.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------
// Linear layer  C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      (torch nn.Linear layout: W is [N, K])
//
//   v = acc + bias[n]                      (bias fp32, optional)
//   v = gelu_erf(v)                        (act == 1)
//   v += res[res_row, n]                   (residual fp32, optional; res_row = r_in if res_periodic else out_row)
//   store as out_dtype
//
// Row remap (lets the conv stem run as a plain GEMM):
//   g = m / m_period_in, r_in = m % m_period_in; rows with r_in >= m_valid are dropped;
//   out_row = g * m_period_out + r_in + m_out_offset.
// out_mode 1 ("head split", used by the cross-K/V projection): n -> (kv = n / d_model, h, j),
//   out[((kv * hs_batch + hs_b0 + g) * heads + h) * m_valid + r_in][j], j < 64.
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
    const void* A = nullptr; long long lda = 0;   // A[m*lda + k], dtype = in_dtype (rows may overlap: lda < K)
    const void* W = nullptr; long long ldw = 0;   // W[n*ldw + k], dtype = in_dtype
    int in_dtype = F32;
    const float* bias = nullptr;
    const float* res = nullptr; long long ldres = 0; int res_periodic = 0;
    void* out = nullptr; long long ldo = 0; int out_dtype = F32;
    int M = 0, N = 0, K = 0;
    int act = 0;
    int m_period_in = 0, m_valid = 0, m_period_out = 0, m_out_offset = 0;  // 0 period = identity
    int out_mode = 0; int hs_heads = 0; int hs_batch = 0; int hs_b0 = 0;   // head-split store
    const int* active = nullptr;  // optional device flag: kernel is a no-op when *active == 0
    // Split-K with DEFERRED reduction (skinny decode GEMMs): split s accumulates k in [s*K/k_splits, (s+1)*K/k_splits) and
    // stores its raw fp32 partial at out + s*split_stride (out_dtype must be F32, no bias/act/residual); the consumer
    // kernel (layernorm pre-add, decode attention q load) sums the slabs in a fixed order and adds the bias.
    // k_splits == 0: let the launcher choose (<= max_k_splits); the choice is written to *chosen_splits.
    int k_splits = 1; long long split_stride = 0; int max_k_splits = 1; int* chosen_splits = nullptr;
    // optional second output (same remap, row-major): fp32 copy of the result
    void* out2 = nullptr; long long ldo2 = 0;
    // Fused logits processors + argmax (LM head of the greedy loop, tcgen05 path only): instead of storing the [M, N] logits, every
    // (row, column tile, epilogue half) stores the maximum of its columns that are not suppressed and its column index:
    //   am_val / am_idx [M, am_stride] with am_stride >= ceil(N / 256) * 2.  The argmax kernel reduces them (first maximum wins).
    // am_mask: vocab bits (bit 0 always suppressed, bit 1 suppressed when the row's length == am_begin); the row's length is
    // am_row_len[row] when given, else am_state->cur_len.  `out` is not written.
    float* am_val = nullptr; int* am_idx = nullptr; int am_stride = 0;
    int* am_count = nullptr;     // out (host): entries per row actually written (depends on the tile width the launcher picked)
    const unsigned char* am_mask = nullptr; const StepState* am_state = nullptr; const int* am_row_len = nullptr; int am_begin = 0;
};
int gemm_argmax_partials(int N);   // partial entries per row the fused LM head writes for N columns

// prefer_tc: use the tcgen05 kernel when in_dtype == BF16 (falls back to SIMT when shapes do not fit)
void gemm(const GemmArgs& a, cudaStream_t stream);
void gemm_simt(const GemmArgs& a, cudaStream_t stream);
void gemm_tc(const GemmArgs& a, cudaStream_t stream);      // bf16 in, tcgen05/TMEM/TMA
bool gemm_tc_supported(const GemmArgs& a);
void set_gemm_backend(int backend);  // 0 = auto (tcgen05 for bf16), 1 = force SIMT
int get_gemm_backend();

// LayerNorm over the last dim: x fp32 [rows, d] -> out (out_dtype) [rows, d], optional fp32 copy out2
void layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_dtype, float* out2,
               int rows, int d, float eps, const int* active, cudaStream_t stream);
// Residual update fused in front of the LayerNorm (consumer side of a split-K GEMM):
//   x[r, :] += add_bias[:] + sum_{s < n_parts} parts[s * part_stride + r * d + :]   (fixed order; x updated in place)
//   out[r, :] = LayerNorm(x[r, :])
struct LnPreAdd { const float* parts = nullptr; int n_parts = 0; long long part_stride = 0; const float* bias = nullptr; };
void layernorm_preadd(float* x, const LnPreAdd& pre, const float* gamma, const float* beta, void* out, int out_dtype,
                      int rows, int d, float eps, const int* active, cudaStream_t stream);

// conv1 im2col: mel fp32 [B, n_mels, T] -> A1 [B*T, kpad] (k = tap * n_mels + c, zero padded to kpad)
void im2col_conv1(const float* mel, void* out, int out_dtype, int B, int n_mels, int T, int kpad, cudaStream_t stream);
// dtype conversion (n elements)
void cast(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream);

// Encoder self-attention over fused qkv [B*S, 3*d] (q already scaled) -> out [B*S, d]; no mask
void encoder_attention(const void* qkv, void* out, int dtype, int B, int S, int H, cudaStream_t stream);
void encoder_attention_simt(const void* qkv, void* out, int dtype, int B, int S, int H, cudaStream_t stream);
void encoder_attention_tc(const void* qkv, void* out, int B, int S, int H, cudaStream_t stream);  // bf16
void set_attn_backend(int backend);  // 0 = auto, 1 = force SIMT
int get_attn_backend();

// One-query attention for the decode step.
struct DecAttnArgs {
    int dtype = F32;
    const void* q = nullptr; long long q_stride = 0;     // q[b*q_stride + h*64 + j] (already scaled)
    // alternative q source (consumer side of a split-K q projection): q = q_bias + sum_s q_parts[s*stride + b*H*64 + ...], fp32
    const float* q_parts = nullptr; int q_n_parts = 0; long long q_part_stride = 0; const float* q_bias = nullptr;
    void* out = nullptr; long long out_stride = 0;       // out[b*out_stride + h*64 + j]
    int B = 0, H = 0;
    // --- contiguous K/V (cross attention, or explicit tensors): K[(b*kv_bstride + h*kv_hstride) + s*64 + j]
    const void* k = nullptr; const void* v = nullptr; long long kv_bstride = 0, kv_hstride = 0;
    int n_keys = 0;                // fixed key count (cross) when n_keys_dev == nullptr
    // --- paged self-attention: append k_new/v_new at slot (len-1), then attend over len keys
    const StepState* state = nullptr;   // when set: len = state->cur_len, no-op if !state->active
    const int* active = nullptr;        // optional device flag for kernels without `state` (cross attention): no-op when 0
    // optional per-utterance flags (the greedy loop's `unfinished_sequences`): rows that already emitted EOS are skipped —
    // their K/V are not read at all (their next tokens are pad regardless, generation/utils.py:1506-1510)
    const int* row_active = nullptr;
    // optional per-utterance lengths (in-flight refill: the rows of a batch are at different positions): when set, row b
    // appends at slot row_len[b] - 1 and attends over row_len[b] keys instead of state->cur_len
    const int* row_len = nullptr;
    const void* k_new = nullptr; const void* v_new = nullptr; long long new_stride = 0;  // [b*new_stride + h*64 + j]
    void* k_pages = nullptr; void* v_pages = nullptr;   // [page][H][page_tokens][64]
    const int* page_table = nullptr; int pages_per_seq = 0; int page_tokens = 64;
};
void decode_attention(const DecAttnArgs& a, cudaStream_t stream);

// Decoder embedding: x[b, :] = E[ids[b, cur_len-1], :] + P[cur_len-1, :]  (fp32 residual stream)
// row_len (optional, in-flight refill): per-row lengths instead of state->cur_len
void decoder_embed(const int* tokens, int tokens_stride, const StepState* state, const void* emb, const void* pos,
                   int dtype, float* x, int B, int d, cudaStream_t stream, const int* row_len = nullptr);

// stateless variants for the module-level drop-ins (explicit position / dense caches)
void embed_tokens(const int* ids, long long ids_stride, int B, int T, int pos0, const void* emb, const void* pos, int dtype,
                  float* x, int d, int vocab, cudaStream_t stream);
void kv_append(const void* past, long long past_bs, long long past_hs, const void* cur, long long cur_bs, void* out, int dtype,
               int B, int H, int n, cudaStream_t stream);

// Logits processors + argmax + EOS/pad bookkeeping + length advance, entirely on device.
struct GreedyArgs {
    const float* logits = nullptr; long long ld = 0;   // [B, V] fp32
    int B = 0, V = 0;
    const unsigned char* vocab_mask = nullptr;  // bit0: always suppressed, bit1: suppressed at begin_index
    int begin_index = 0;
    const int* force_map = nullptr;             // [max_length] forced token at generation index, or -1
    int pad_id = 0, eos_id = 0, max_length = 0;
    int* tokens = nullptr; int tokens_stride = 0;   // [B, max_length]
    int* unfinished = nullptr;                      // [B]
    StepState* state = nullptr;
    const int* forced_tokens = nullptr;         // teacher forcing (tests): take ids[b, n] from here instead of argmax
    // partial (maximum, index) entries written by the fused LM head epilogue (GemmArgs::am_*): reduced here instead of scanning
    // the logits (which are then not needed and `logits` may be null)
    const float* part_val = nullptr; const int* part_idx = nullptr; int n_parts = 0; int part_stride = 0;
    // in-flight refill: per-row lengths [B].  Row b writes its token at ids[b, row_len[b]], the processors see its own length,
    // a row stops growing once it has emitted EOS or reached max_length (no pad is appended: the host pads the results)
    int* row_len = nullptr;
    // optional (whole-step decoder kernel): also write the embedding of the chosen token, the input of the next step,
    // x[b, :] = E[tok, :] + P[len, :] (fp32 [B, d]; bf16 tables, d % 8 == 0)
    float* embed_x = nullptr; const void* embed_table = nullptr; const void* embed_pos = nullptr; int embed_d = 0;
};
void greedy_step(const GreedyArgs& a, cudaStream_t stream);
void greedy_init(int* tokens, int tokens_stride, int* unfinished, StepState* state, int B, int start_token,
                 int pad_id, int max_length, cudaStream_t stream);
// plain argmax with the processors applied, no bookkeeping (module-level API / tests)
void argmax_rows(const float* logits, long long ld, int B, int V, const unsigned char* vocab_mask, int mask_bits,
                 int* out, cudaStream_t stream);

}  // namespace wb

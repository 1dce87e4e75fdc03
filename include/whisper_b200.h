/* whisper_b200.h — C-ABI of libwhisper_b200.so: the B200-native (sm_100a) Whisper greedy-inference path.
 *
 * Boundary conventions (they mirror the reference's native plugin boundary):
 *   - `extern "C"`, plain pointers and sizes, no torch / C++ types        (InferPlugin.cpp:149-170 `initLibNvInferPlugins`)
 *   - every entry point returns 0 on success or a negative wb_status; the message is in wb_last_error()
 *     (plugin enqueue returns int, 0 = ok; exceptions are caught at the boundary: identityPlugin.cpp:82-108,190-203)
 *   - device pointers are NON-OWNING, work is enqueued asynchronously on the caller's cudaStream_t and the
 *     caller synchronises, exactly like Session.run(inputs, outputs, stream)      (runtime/session.py:148-178)
 *   - workspaces are provided by the caller (plugin `workspace` argument, identityPlugin.cpp:82-85)
 *   - loaded with ctypes.CDLL(..., RTLD_GLOBAL) + hand-set argtypes like tensorrt_llm/plugin/plugin.py:10-22
 *
 * dtype codes: 0 = float32, 1 = bfloat16.
 */
#ifndef WHISPER_B200_H_
#define WHISPER_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define WB_API __attribute__((visibility("default")))
#else
#define WB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb_model wb_model;     /* packed device weights of one Whisper checkpoint            */
typedef struct wb_session wb_session; /* per-batch state: encoder workspace, KV caches, token loop   */
typedef void* wb_stream;              /* cudaStream_t                                                */

enum wb_status {
    WB_OK = 0,
    WB_ERR_INVALID = -1,   /* bad argument / shape / alignment            */
    WB_ERR_CUDA = -2,      /* CUDA runtime error (message has the call)   */
    WB_ERR_DRIVER = -3,    /* driver entry point / tensor-map failure     */
    WB_ERR_STATE = -4,     /* object not ready (missing weights, ...)     */
    WB_ERR_INTERNAL = -9
};

/* Hyper-parameters: the subset of the reference's config.pkl (WhisperConfig.to_dict(), build_encoder.py:42-45)
 * that the path reads.  Mirrors the ctor kwargs of tensorrt_llm.models.WhisperEncoder / WhisperDecoder
 * (models/whisper/model.py:69-72, 372-382). */
typedef struct wb_config {
    int32_t d_model, n_heads, encoder_layers, decoder_layers, ffn_dim, vocab_size;
    int32_t num_mel_bins, n_frames, max_source_positions, max_target_positions;
    int32_t decoder_start_token_id, eos_token_id, pad_token_id, max_length;
} wb_config;

/* ---- library ---------------------------------------------------------------------------------- */
WB_API const char* wb_last_error(void);                 /* thread-local message of the last failing call   */
WB_API int wb_version(void);
WB_API int wb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* debug/bench switches: gemm 0 = tcgen05 for bf16 (default), 1 = CUDA-core; attention likewise */
WB_API int wb_set_backend(int gemm_backend, int attn_backend);
/* programmatic dependent launch between the kernels of a decode step (default 0 = plain stream order; measured slower) */
WB_API int wb_set_pdl(int enabled);
/* measurement tool: stream `bytes` of device memory once (mode 0: 16-byte L1-bypassing loads, mode 1: cp.async.bulk into a
 * shared-memory ring); time it with CUDA events to get the pure-read HBM ceiling of this GPU.  sink: 4 device bytes. */
WB_API int wb_bandwidth_probe(const void* buf, size_t bytes, int mode, int ctas_per_sm, void* sink, wb_stream stream);
/* cross-attention kernel of the decode step: 0 = 16-byte L1-bypassing loads, 4 CTAs/SM (default); 1 = cp.async.bulk into a
 * 128 KB shared-memory ring, one CTA/SM (bf16 only).  Existing CUDA graphs keep the kernel they were captured with. */
WB_API int wb_set_decode_attention_backend(int backend);
/* paged self-attention kernel: 0 = one warp per (utterance, head) item (default, measured fastest), 1..4 = one CTA per item with
 * (threads, batch depth) = (128, 4) (128, 8) (64, 8) (256, 4), 5..6 = warp-kernel tuning variants */
WB_API int wb_set_self_attention_warp_kernel(int variant);
/* decode steps of <= 16 utterances (bf16) run as ONE persistent cooperative kernel per token, phases separated by grid barriers
 * (csrc/step_mega.cu).  mode 1 (default) = on; 2 = on, with the attention phases on mma.sync blocks of 16 keys instead of
 * CUDA-core dot products (measured 4-13 % slower, kept for A/B); 0 = the multi-kernel step for every batch size (A/B, parity tests).
 * Sessions decoded concurrently by wb_decode_run_multi never use it (a cooperative grid needs every SM to itself). */
WB_API int wb_set_small_batch_path(int mode);
/* decode steps of > 16 utterances (bf16): the GEMM / LayerNorm chains between the attention kernels run as persistent
 * cooperative tcgen05 kernels with grid barriers between their phases (csrc/step_chain.cu; 4 launches per decoder layer
 * instead of 11).  1 (default) = on, 0 = one kernel per GEMM / LayerNorm (A/B, parity tests).  Same rounding points either
 * way.  Replaces nothing in the reference (its decode step is one TensorRT engine execution, run.py:93-148). */
WB_API int wb_set_decode_chain_path(int enabled);
/* measurement hook of the whole-step kernel: device buffer of 8 * (8 * decoder_layers + 2) int64; CTA 0 stores its SM clock
 * at 8 points of every phase: [0] start, [1] activations staged, [2] block barrier passed, [3] first weight tile requested,
 * [4] phase done, [5] arrived at the grid barrier + next phase prefetched, [8] = next [0] grid barrier passed
 * (tools/step_trace.py).  NULL (default) = off.  Decode-step graphs captured afterwards carry the pointer. */
WB_API int wb_set_step_trace(void* device_buffer);
/* skinny (decode-step) GEMMs use the variant sized to co-reside with the bulk-ring cross-attention CTA of a concurrent stream
 * (256 threads, <= 128 registers, <= 90 KB shared memory); set together with wb_decode_run_multi.  Default 0. */
WB_API int wb_set_lean_decode_gemm(int enabled);
/* tuning hook: force the tcgen05 GEMM's tile width (0 = automatic choice, see pick_config in csrc/gemm_tc.cu) */
WB_API int wb_set_gemm_block_n(int block_n);
/* wb_decode_run replays the decode step as a CUDA graph (default 1 = on); 0 = one launch per kernel.  Existing graphs are kept. */
WB_API int wb_set_cuda_graphs(int enabled);
/* number of kernels this library launched so far on this thread's device (bench `gpu_launches`) */
WB_API long long wb_launch_count(void);

/* ---- model: replaces build_encoder.py / build_decoder.py (HF state_dict -> engine)  ---------- */
WB_API int wb_model_create(const wb_config* cfg, int dtype, wb_model** out);
WB_API int wb_model_destroy(wb_model* m);
/* `name` is the oracle's state_dict key (binders build_encoder.py:71-91,
 * build_decoder.py:71-101); `data` is HOST fp32, contiguous, torch layout.  q/k/v are fused on the fly
 * (zero K bias), the q scale is folded in, conv weights are re-laid out for the GEMM stem. */
WB_API int wb_model_load_tensor(wb_model* m, const char* name, const float* data, int64_t numel);
/* logits processors of run.py:150-162: suppress_tokens, begin_suppress_tokens (+begin_index), forced ids */
WB_API int wb_model_set_generation(wb_model* m, const int32_t* suppress, int n_suppress, const int32_t* begin_suppress,
                            int n_begin, int begin_index, const int32_t* forced_pairs, int n_forced);
WB_API int wb_model_weight_bytes(const wb_model* m, size_t* bytes);

/* ---- session ------------------------------------------------------------------------------------ */
WB_API int wb_session_workspace_bytes(const wb_model* m, int max_batch, int enc_chunk, size_t* bytes);
WB_API int wb_session_create(wb_model* m, int max_batch, int enc_chunk, void* workspace, size_t workspace_bytes, wb_session** out);
WB_API int wb_session_destroy(wb_session* s);
/* Per-session override of a process-wide A/B switch (the wb_set_* functions below change the default for every session of the
 * process; a session option wins for that session only).  name: "small_batch_path" (wb_set_small_batch_path != 0),
 * "decode_chain_path" (wb_set_decode_chain_path), "cuda_graphs" (wb_set_cuda_graphs), "merge_attention" (no process-wide
 * switch; default off: for small batches the attention kernels run as phases of the layer's persistent launch - one launch per
 * decoder layer, measured slower than the stand-alone attention kernels); value: 1 on, 0 off, -1 inherit the process-wide
 * switch again.  Takes effect at the next decode step (a captured step graph of the other kind is rebuilt).
 * The reference has no equivalent (one TensorRT execution context per engine, runtime/session.py:53-61). */
WB_API int wb_session_set_option(wb_session* session, const char* name, int value);

/* WhisperEncoder.__call__(input) -> hidden_states (run.py:81-91): mel fp32 [B, 80, 3000] on device.
 * Also projects the cross-attention K/V of every decoder layer once (oracle branch modeling_whisper.py:484-487).
 * enc_out_f32 (optional) receives hidden_states fp32 [B, 1500, d]. */
WB_API int wb_encode(wb_session* s, const float* mel, int batch, float* enc_out_f32, wb_stream stream);
/* use encoder states produced elsewhere (decoder drop-in: `encoder_hidden_states` input, model.py:487-490) */
WB_API int wb_set_encoder_output(wb_session* s, const void* enc_states, int dtype, int batch, wb_stream stream);

/* greedy_search (run.py:171-227 == generation/utils.py:1474-1529), entirely on device.
 * wb_decode_begin resets ids to [[decoder_start_token_id]] * B; wb_decode_step enqueues ONE step
 * (WhisperDecoder.__call__ + processors + argmax + EOS bookkeeping); wb_decode_run enqueues up to max_steps
 * steps (<=0: until max_length), polls the stop flag every `check_every` steps, synchronises and returns the
 * final sequence length in *final_len. */
WB_API int wb_decode_begin(wb_session* s, int batch, wb_stream stream);
WB_API int wb_decode_step(wb_session* s, wb_stream stream);
WB_API int wb_decode_run(wb_session* s, int max_steps, int check_every, int* final_len, wb_stream stream);
/* Greedy loops of several sessions of ONE model (disjoint sub-batches of utterances) interleaved step by step, each on its
 * own stream, so that the HBM-bound cross-attention of one sub-batch overlaps the latency-bound GEMM / LayerNorm kernels of
 * the others.  wb_decode_begin must have been called on every session.  final_lens: one int per session. */
WB_API int wb_decode_run_multi(wb_session** sessions, int n_sessions, int max_steps, int check_every, int* final_lens, wb_stream stream);
/* Finished-row compaction (not in the reference, which decodes one utterance at a time): call between two
 * wb_decode_run(max_steps = n) windows.  Synchronises `stream`, reads the per-row unfinished flags
 * (generation/utils.py:1506-1520) and, if rows have emitted EOS, moves the rows still running to the front of the batch (ids,
 * page-table rows, cross K/V rows) so that the following steps run on fewer rows; the ids of the rows that left are kept in a
 * result buffer in ORIGINAL row order, which wb_decode_tokens returns from then on.  *rows_running = rows still decoding,
 * 0 when the loop has stopped.  Not available with teacher forcing / logits dumps. */
WB_API int wb_decode_compact(wb_session* s, int* rows_running, wb_stream stream);
/* In-flight refill: call between two wb_decode_run windows.  (1) Utterances that have emitted EOS or reached
 * max_length leave the batch: their ids are written to the HOST arrays finished_utt [max_batch] (utterance number: decode_begin's
 * rows are 0 .. batch-1, every admitted utterance gets the next number), finished_len [max_batch] and finished_ids
 * [max_batch, max_target_positions] (first finished_len[k] entries valid; nothing is padded after EOS).  (2) The rows still
 * decoding move to the front.  (3) Up to n_new new utterances (mel_new: DEVICE fp32 [n_new, 80, 3000]) are encoded into the
 * freed slots and start decoding at position 0 while the others continue where they are: from then on the kernels use per-row
 * lengths.  *n_admitted tells how many of mel_new were taken (the first ones), *n_rows how many rows decode now (0: nothing left,
 * the loop is over).  Synchronises `stream`.  Every utterance gets the ids it would get alone (rows are independent).
 * The reference transcribes utterances one by one (examples/whisper/run.py:263-288); TensorRT-LLM's GPT runtime has the idea
 * (docs/in_flight_batching.md), its Whisper path does not. */
WB_API int wb_decode_refill(wb_session* session, const float* mel_new, int n_new, int32_t* finished_utt, int32_t* finished_len,
                            int32_t* finished_ids, int* n_finished, int* n_admitted, int* n_rows, wb_stream stream);
/* ids int32 [B, max_target_positions] (row stride max_target_positions), original row order; device pointer owned by the session */
WB_API int wb_decode_tokens(wb_session* s, const int32_t** tokens_dev, int* row_stride);
/* raw next-token logits of the last wb_decode_step, fp32 [B, vocab] (the decoder engine's output tensor, model.py:464).
 * wb_decode_run does NOT materialise logits in the bf16 path (the LM head's epilogue applies the logits processors and reduces
 * the argmax per column tile) unless a logits dump is set (wb_decode_set_logits_dump): after it the buffer holds scratch. */
WB_API int wb_decode_logits(wb_session* s, const float** logits_dev);
/* test hooks: teacher forcing (ids taken from forced[B, max_target_positions]) and per-step logits dump */
WB_API int wb_decode_set_forced_tokens(wb_session* s, const int32_t* forced_dev);
WB_API int wb_decode_set_logits_dump(wb_session* s, float* dump_dev, int max_steps);
/* views of the caches: cross K/V of layer l: [2][max_batch][H][1500][64]; self pages [num_pages][H][64][64] */
WB_API int wb_session_cross_kv(wb_session* s, int layer, const void** kv_dev, int64_t* kv_stride_elems);
WB_API int wb_session_self_kv(wb_session* s, int layer, const void** k_pages, const void** v_pages, const int32_t** page_table,
                       int* pages_per_seq, int* page_tokens);

/* live timing for the bench roofline: CUDA events are recorded on the launching stream around every launch of
 * one kernel class inside the real loop.  classes: 1 cross-attention, 2 self-attention, 3 decode GEMMs,
 * 4 LM head, 5 encoder GEMMs, 6 encoder attention, 8 greedy/argmax, 9 conv stem, 10 cross-K/V projection; 0 = off.
 * wb_session_profile_read synchronises, returns the summed device time and the launch count, and resets. */
WB_API int wb_session_profile(wb_session* s, int kernel_class);
/* like wb_session_profile, but decode-step kernel classes are timed only in the decode step with index `decode_step` of every
 * greedy loop (that one step is launched eagerly, all others keep replaying the CUDA graph): per-launch events inside the
 * timed region of the bench without giving up graph replay.  decode_step < 0 = every step. */
WB_API int wb_session_profile_at(wb_session* s, int kernel_class, int decode_step);
WB_API int wb_session_profile_read(wb_session* s, double* total_ms, long long* launches);

/* ---- operators (module-level drop-ins and kernel tests) ---------------------------------------- */
/* LayerNorm eps: x fp32 [rows, d] -> out (out_dtype)                         (layers/normalization.py:6-30) */
WB_API int wb_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_dtype, int rows, int d,
                 float eps, wb_stream stream);
/* Linear: out = act(A W^T + bias) + residual; A [M,K] (lda), W [N,K], dtypes in_dtype, bias/residual fp32;
 * act 0 none / 1 erf-GELU; backend 0 auto / 1 CUDA-core / 2 tcgen05       (layers/linear.py:38-139) */
WB_API int wb_linear(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, const float* bias,
              const float* residual, int64_t ldres, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
              int backend, wb_stream stream);
/* Skinny (decode-step) Linear as split-K with DEFERRED reduction: split s accumulates k in [s*K/S, (s+1)*K/S) and stores its raw
 * fp32 product at parts + s*split_stride ([M, N] row-major); no bias / activation.  k_splits 0 = let the library choose
 * (<= max_k_splits; the choice is returned in *chosen_splits).  The consumer adds bias + slabs in a fixed order:
 * wb_layernorm_preadd (residual GEMMs: RowLinear `dense` / fc2, layers/linear.py:98-139) or the cross-attention q load. */
WB_API int wb_linear_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, float* parts, int64_t split_stride,
                     int M, int N, int K, int k_splits, int max_k_splits, int* chosen_splits, wb_stream stream);
/* x[r,:] += add_bias + sum_s parts[s*part_stride + r*d + :] (in place), then out[r,:] = LayerNorm(x[r,:])
 * (the residual add + pre-LN of WhisperDecoderLayer.forward, model.py:336-361, fused into one pass) */
WB_API int wb_layernorm_preadd(float* x, const float* parts, int n_parts, int64_t part_stride, const float* add_bias, const float* gamma,
                        const float* beta, void* out, int out_dtype, int rows, int d, float eps, wb_stream stream);
/* Encoder stem (conv1+GELU, conv2+GELU, +positions): mel fp32 [B,80,3000] -> x fp32 [B*1500, d]; uses the
 * session's workspace and the model's packed conv weights             (models/whisper/model.py:96-102) */
WB_API int wb_encoder_stem(wb_session* s, const float* mel, int batch, float* x_out, wb_stream stream);
/* Encoder self-attention over fused qkv [B*S, 3*H*64] (q pre-scaled) -> out [B*S, H*64]; backend as above */
WB_API int wb_encoder_attention(const void* qkv, void* out, int dtype, int batch, int seq, int heads, int backend, wb_stream stream);
/* One-query attention over explicit contiguous K/V [B, H, T, 64] (WhisperDecoderAttention cached modes,
 * model.py:240-304): q [B, H*64] pre-scaled, n_keys <= T */
WB_API int wb_decode_attention(const void* q, const void* k, const void* v, void* out, int dtype, int batch, int heads,
                        int n_keys, int64_t kv_batch_stride, int64_t kv_head_stride, wb_stream stream);
/* Cached decoder SELF-attention of one greedy step on its own (WhisperDecoderAttention, self / with-cache mode:
 * models/whisper/model.py:273-281 slice + concat, modeling_whisper.py:490-497 torch.cat): the new key / value rows are appended
 * IN PLACE at slot cur_len - 1 of the paged cache (pages of 64 tokens x all heads, [page][H][64][64]; row b owns
 * page_table[b * pages_per_seq ..]) and every (utterance, head) attends over its cur_len keys.  qkv: fused rows
 * [batch, 3 * heads * 64] (q pre-scaled | k | v), row_stride elements apart; out [batch, heads * 64].
 * state: device pointer to 8 int32 {cur_len, active, ...} (the loop state of the greedy session; active == 0 -> no-op);
 * row_active: optional [batch] int32, rows with 0 are skipped.  Same dispatch as the session's decode step (one warp per
 * item when batch * heads exceeds two items per SM, one CTA per item below). */
WB_API int wb_paged_self_attention(const void* qkv, int64_t row_stride, void* out, void* k_pages, void* v_pages,
                                   const int32_t* page_table, int pages_per_seq, int dtype, int batch, int heads,
                                   const int32_t* state, const int32_t* row_active, wb_stream stream);
/* masked argmax per row: logits fp32 [B, V]; mask (optional) uint8 [V], a token is skipped if mask & bits */
WB_API int wb_argmax(const float* logits, int64_t ld, int batch, int vocab, const uint8_t* mask, int bits, int32_t* out,
              wb_stream stream);
/* Decoder embedding with an explicit position (WhisperDecoder.forward, model.py:423-425: position row =
 * shape(past_self_cache_mask,0)-1): x[b*T+t, :] = embed_tokens[ids[b,t], :] + embed_positions[position0+t, :], fp32 out */
WB_API int wb_embed(const int32_t* ids, int64_t ids_stride, int batch, int n_tokens, int position0, const void* embed_tokens,
             const void* embed_positions, int dtype, int d_model, int vocab_size, float* x_out, wb_stream stream);
/* Dense cache growth of the engine contract (model.py:276-281 slice+concat; oracle torch.cat modeling_whisper.py:494-495):
 * out [B,H,cache_len+1,64] = concat(past[:, :, :cache_len] (strided), current [B, H*64] (row stride)) */
WB_API int wb_kv_append(const void* past, int64_t past_batch_stride, int64_t past_head_stride, int cache_len, const void* current,
                 int64_t current_batch_stride, void* out, int dtype, int batch, int heads, wb_stream stream);
/* Linear + transpose_for_scores (model.py:254-259): A [B*S, K] x W[H*64, K]^T (+bias) -> out [B, H, S, 64] */
WB_API int wb_linear_split_heads(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, const float* bias, void* out,
                          int out_dtype, int batch, int seq, int heads, int K, wb_stream stream);
/* Stateless encoder stem (WhisperEncoder.forward, model.py:94-102): conv1+GELU, conv2(stride 2)+GELU, +positions.
 * w1_packed [d, 256] with k = tap*num_mel_bins + c (zero padded), w2_packed [d, 3d] with k = tap*d + c, compute dtype;
 * biases / positions fp32; x_out fp32 [B*n_frames/2, d]; caller-provided workspace */
WB_API int wb_conv_stem_workspace_bytes(int batch, int d_model, int n_frames, int dtype, size_t* bytes);
WB_API int wb_conv_stem(const float* mel, int batch, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                 const float* positions, int dtype, int d_model, int num_mel_bins, int n_frames, void* workspace,
                 size_t workspace_bytes, float* x_out, wb_stream stream);
/* Log-mel front-end (the step BEFORE the path: WhisperFeatureExtractor, feature_extraction_whisper.py:96-109, numpy on the CPU
 * in run.py:267).  pcm fp32 [B, 480000] on the device (16 kHz mono, zero padded / trimmed to 30 s); window [400] periodic Hann;
 * dft_basis [408, 400] rows 0..200 cos(2 pi k j / 400), rows 201..401 -sin(...), rest 0; mel_filters [80, 208] (bins 201.. = 0);
 * all fp32 device constants built by the caller (whisper_trtllm_b200/frontend.py).  input_features fp32 [B, 80, 3000]. */
WB_API int wb_log_mel_workspace_bytes(int batch, size_t* bytes);
WB_API int wb_log_mel(const float* pcm, int batch, const float* window, const float* dft_basis, const float* mel_filters, void* workspace,
               size_t workspace_bytes, float* input_features, wb_stream stream);
WB_API int wb_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, wb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H_ */

// extern "C" boundary of libwhisper_b200.so (declared in include/whisper_b200.h).
// Exceptions never cross it: every entry point returns a wb_status and stores the message thread-locally,
// the same discipline as the reference's plugin boundary (identityPlugin.cpp:190-203, InferPlugin.cpp:149-170).
#include "../../include/whisper_b200.h"

#include "wb_runtime.h"

namespace {
thread_local std::string g_last_error;

template <typename F> int guarded(F&& f) {
    try {
        f();
        return WB_OK;
    } catch (const wb::Error& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WB_ERR_INTERNAL;
    } catch (...) {
        g_last_error = "unknown error";
        return WB_ERR_INTERNAL;
    }
}
inline cudaStream_t S(wb_stream s) { return reinterpret_cast<cudaStream_t>(s); }
inline wb::Model* M(wb_model* m) { return reinterpret_cast<wb::Model*>(m); }
inline const wb::Model* M(const wb_model* m) { return reinterpret_cast<const wb::Model*>(m); }
inline wb::Session* SS(wb_session* s) { return reinterpret_cast<wb::Session*>(s); }
}  // namespace

#define WB_NOT_NULL(p) WB_REQUIRE((p) != nullptr, #p " is null")

namespace wb {
void set_gemm_tc_block_n(int bn);
void set_cuda_graphs(bool on);
void set_decode_attention_backend(int b);
void set_lean_decode_gemm(bool on);
void set_small_batch_path(bool on);
void set_chain_path(bool on);
void set_mega_attention_tc(bool on);
void set_step_trace(long long* dev_ptr);
void set_self_attention_variant(int v);
size_t log_mel_workspace_bytes(int chunk);
void log_mel(const float* pcm, int B, const float* window, const float* dft_basis, const float* mel_filters, void* workspace,
             size_t workspace_bytes, float* out, cudaStream_t st);
void bandwidth_probe(const void* buf, size_t bytes, int mode, int ctas_per_sm, unsigned* sink, cudaStream_t stream);
}

extern "C" {

const char* wb_last_error(void) { return g_last_error.c_str(); }
int wb_version(void) { return 100; }

int wb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    return guarded([&] {
        int dev = 0;
        WB_CHECK_CUDA(cudaGetDevice(&dev));
        if (sm_count) WB_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
        if (cc_major) WB_CHECK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        if (cc_minor) WB_CHECK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    });
}

int wb_set_backend(int gemm_backend, int attn_backend) {
    return guarded([&] {
        WB_REQUIRE(gemm_backend >= 0 && gemm_backend <= 1 && attn_backend >= 0 && attn_backend <= 1, "backend must be 0 or 1");
        wb::set_gemm_backend(gemm_backend);
        wb::set_attn_backend(attn_backend);
    });
}

int wb_bandwidth_probe(const void* buf, size_t bytes, int mode, int ctas_per_sm, void* sink, wb_stream stream) {
    return guarded([&] { wb::bandwidth_probe(buf, bytes, mode, ctas_per_sm, (unsigned*)sink, S(stream)); });
}

int wb_set_decode_attention_backend(int backend) {
    return guarded([&] {
        WB_REQUIRE(backend >= 0 && backend <= 12, "backend must be 0 (16-byte loads), 1 (cp.async.bulk ring) or a tuning variant 2..12");
        wb::set_decode_attention_backend(backend);
    });
}

int wb_log_mel_workspace_bytes(int batch, size_t* bytes) {
    return guarded([&] {
        WB_NOT_NULL(bytes);
        WB_REQUIRE(batch > 0, "batch must be positive");
        *bytes = wb::log_mel_workspace_bytes(batch < 64 ? batch : 64);
    });
}
int wb_log_mel(const float* pcm, int batch, const float* window, const float* dft_basis, const float* mel_filters, void* workspace,
               size_t workspace_bytes, float* input_features, wb_stream stream) {
    return guarded([&] { wb::log_mel(pcm, batch, window, dft_basis, mel_filters, workspace, workspace_bytes, input_features, S(stream)); });
}

int wb_set_self_attention_warp_kernel(int variant) {
    return guarded([&] {
        WB_REQUIRE(variant >= 0 && variant <= 6, "variant must be 0..6");
        wb::set_self_attention_variant(variant);
    });
}

int wb_set_small_batch_path(int mode) {
    return guarded([&] {
        WB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (off), 1 (on) or 2 (on, mma.sync attention)");
        wb::set_small_batch_path(mode != 0);
        wb::set_mega_attention_tc(mode == 2);
    });
}

int wb_set_decode_chain_path(int enabled) {
    wb::set_chain_path(enabled != 0);
    return WB_OK;
}

int wb_set_step_trace(void* device_buffer) {
    wb::set_step_trace(reinterpret_cast<long long*>(device_buffer));
    return WB_OK;
}

int wb_set_lean_decode_gemm(int enabled) {
    wb::set_lean_decode_gemm(enabled != 0);
    return WB_OK;
}

int wb_set_gemm_block_n(int block_n) {
    return guarded([&] {
        WB_REQUIRE(block_n == 0 || block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256, "block_n must be 0 (auto), 32, 64, 128 or 256");
        wb::set_gemm_tc_block_n(block_n);
    });
}

int wb_set_cuda_graphs(int enabled) {
    wb::set_cuda_graphs(enabled != 0);
    return WB_OK;
}

int wb_set_pdl(int enabled) {
    wb::pdl_enabled() = enabled != 0;
    return WB_OK;
}

long long wb_launch_count(void) { return wb::launch_counter().load(); }

int wb_model_create(const wb_config* c, int dtype, wb_model** out) {
    return guarded([&] {
        WB_NOT_NULL(c);
        WB_NOT_NULL(out);
        wb::ModelConfig g;
        g.d_model = c->d_model; g.n_heads = c->n_heads; g.enc_layers = c->encoder_layers; g.dec_layers = c->decoder_layers;
        g.ffn = c->ffn_dim; g.vocab = c->vocab_size; g.n_mels = c->num_mel_bins; g.n_frames = c->n_frames;
        g.n_ctx = c->max_source_positions; g.max_tgt = c->max_target_positions;
        g.sot = c->decoder_start_token_id; g.eos = c->eos_token_id; g.pad = c->pad_token_id; g.max_length = c->max_length;
        *out = reinterpret_cast<wb_model*>(new wb::Model(g, dtype));
    });
}
int wb_model_destroy(wb_model* m) {
    return guarded([&] { delete M(m); });
}
int wb_model_load_tensor(wb_model* m, const char* name, const float* data, int64_t numel) {
    return guarded([&] {
        WB_NOT_NULL(m);
        WB_NOT_NULL(name);
        M(m)->load_tensor(name, data, numel);
    });
}
int wb_model_set_generation(wb_model* m, const int32_t* suppress, int n_suppress, const int32_t* begin_suppress, int n_begin,
                            int begin_index, const int32_t* forced_pairs, int n_forced) {
    return guarded([&] {
        WB_NOT_NULL(m);
        M(m)->set_generation(suppress, n_suppress, begin_suppress, n_begin, begin_index, forced_pairs, n_forced);
    });
}
int wb_model_weight_bytes(const wb_model* m, size_t* bytes) {
    return guarded([&] {
        WB_NOT_NULL(m);
        WB_NOT_NULL(bytes);
        *bytes = M(m)->weight_bytes;
    });
}

int wb_session_workspace_bytes(const wb_model* m, int max_batch, int enc_chunk, size_t* bytes) {
    return guarded([&] {
        WB_NOT_NULL(m);
        WB_NOT_NULL(bytes);
        WB_REQUIRE(max_batch > 0 && enc_chunk > 0 && enc_chunk <= max_batch, "need 0 < enc_chunk <= max_batch");
        *bytes = wb::Session::workspace_bytes(M(m), max_batch, enc_chunk);
    });
}
int wb_session_create(wb_model* m, int max_batch, int enc_chunk, void* workspace, size_t workspace_bytes, wb_session** out) {
    return guarded([&] {
        WB_NOT_NULL(out);
        *out = reinterpret_cast<wb_session*>(new wb::Session(M(m), max_batch, enc_chunk, workspace, workspace_bytes));
    });
}
int wb_session_set_option(wb_session* session, const char* name, int value) {
    return guarded([&] {
        WB_NOT_NULL(session); WB_NOT_NULL(name);
        SS(session)->set_option(name, value);
    });
}

int wb_session_destroy(wb_session* s) {
    return guarded([&] { delete SS(s); });
}

int wb_encode(wb_session* s, const float* mel, int batch, float* enc_out_f32, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->encode(mel, batch, enc_out_f32, S(stream));
    });
}
int wb_set_encoder_output(wb_session* s, const void* enc_states, int dtype, int batch, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->set_encoder_output(enc_states, dtype, batch, S(stream));
    });
}
int wb_decode_begin(wb_session* s, int batch, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->decode_begin(batch, S(stream));
    });
}
int wb_decode_step(wb_session* s, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->decode_step(S(stream));
    });
}
int wb_decode_run(wb_session* s, int max_steps, int check_every, int* final_len, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        const int n = SS(s)->decode_run(max_steps, check_every, S(stream));
        if (final_len) *final_len = n;
    });
}
int wb_decode_run_multi(wb_session** sessions, int n_sessions, int max_steps, int check_every, int* final_lens, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(sessions);
        wb::decode_run_multi(reinterpret_cast<wb::Session**>(sessions), n_sessions, max_steps, check_every, final_lens, S(stream));
    });
}
int wb_decode_compact(wb_session* s, int* rows_running, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s);
        const int n = SS(s)->decode_compact(S(stream));
        if (rows_running) *rows_running = n;
    });
}
int wb_decode_refill(wb_session* s, const float* mel_new, int n_new, int32_t* finished_utt, int32_t* finished_len, int32_t* finished_ids,
                     int* n_finished, int* n_admitted, int* n_rows, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s); WB_NOT_NULL(finished_utt); WB_NOT_NULL(finished_len); WB_NOT_NULL(finished_ids);
        WB_NOT_NULL(n_finished); WB_NOT_NULL(n_rows);
        std::vector<wb::Session::Finished> fin;
        int adm = 0;
        const int rows = SS(s)->decode_refill(mel_new, n_new, fin, &adm, S(stream));
        const int max_tgt = SS(s)->m->cfg.max_tgt;
        for (size_t k = 0; k < fin.size(); ++k) {
            finished_utt[k] = fin[k].utt;
            finished_len[k] = fin[k].len;
            std::copy(fin[k].ids.begin(), fin[k].ids.end(), finished_ids + k * (size_t)max_tgt);
        }
        *n_finished = (int)fin.size();
        if (n_admitted) *n_admitted = adm;
        *n_rows = rows;
    });
}

int wb_decode_tokens(wb_session* s, const int32_t** tokens_dev, int* row_stride) {
    return guarded([&] {
        WB_NOT_NULL(s);
        WB_NOT_NULL(tokens_dev);
        *tokens_dev = SS(s)->compacted ? SS(s)->result_tokens : SS(s)->tokens;
        if (row_stride) *row_stride = SS(s)->m->cfg.max_tgt;
    });
}
int wb_decode_logits(wb_session* s, const float** logits_dev) {
    return guarded([&] {
        WB_NOT_NULL(s);
        WB_NOT_NULL(logits_dev);
        *logits_dev = SS(s)->logits;
    });
}
int wb_decode_set_forced_tokens(wb_session* s, const int32_t* forced_dev) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->forced_tokens = forced_dev;
    });
}
int wb_decode_set_logits_dump(wb_session* s, float* dump_dev, int max_steps) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->logits_dump = dump_dev;
        SS(s)->logits_dump_steps = dump_dev ? max_steps : 0;
    });
}
int wb_session_cross_kv(wb_session* s, int layer, const void** kv_dev, int64_t* kv_stride_elems) {
    return guarded([&] {
        WB_NOT_NULL(s);
        WB_NOT_NULL(kv_dev);
        wb::Session* ss = SS(s);
        WB_REQUIRE(layer >= 0 && layer < ss->m->cfg.dec_layers, "layer out of range");
        *kv_dev = (const uint8_t*)ss->cross + (size_t)layer * ss->cross_layer_elems() * wb::dtype_size(ss->m->dtype);
        if (kv_stride_elems) *kv_stride_elems = (int64_t)(ss->cross_layer_elems() / 2);
    });
}
int wb_session_self_kv(wb_session* s, int layer, const void** k_pages, const void** v_pages, const int32_t** page_table,
                       int* pages_per_seq, int* page_tokens) {
    return guarded([&] {
        WB_NOT_NULL(s);
        wb::Session* ss = SS(s);
        WB_REQUIRE(layer >= 0 && layer < ss->m->cfg.dec_layers, "layer out of range");
        const size_t off = (size_t)layer * ss->self_layer_elems() * wb::dtype_size(ss->m->dtype);
        if (k_pages) *k_pages = (const uint8_t*)ss->self_k + off;
        if (v_pages) *v_pages = (const uint8_t*)ss->self_v + off;
        if (page_table) *page_table = ss->page_table;
        if (pages_per_seq) *pages_per_seq = ss->pages_per_seq;
        if (page_tokens) *page_tokens = wb::PAGE_TOKENS;
    });
}

int wb_session_profile(wb_session* s, int kernel_class) {
    return guarded([&] {
        WB_NOT_NULL(s);
        WB_REQUIRE(kernel_class >= 0 && kernel_class <= 10, "unknown kernel class");
        SS(s)->prof_class = kernel_class;
        SS(s)->prof_step = -1;
        SS(s)->prof_used = 0;
    });
}
int wb_session_profile_at(wb_session* s, int kernel_class, int decode_step) {
    return guarded([&] {
        WB_NOT_NULL(s);
        WB_REQUIRE(kernel_class >= 0 && kernel_class <= 10, "unknown kernel class");
        SS(s)->prof_class = kernel_class;
        SS(s)->prof_step = decode_step;
        SS(s)->prof_used = 0;
    });
}
int wb_session_profile_read(wb_session* s, double* total_ms, long long* launches) {
    return guarded([&] {
        WB_NOT_NULL(s);
        SS(s)->prof_read(total_ms, launches);
    });
}

int wb_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_dtype, int rows, int d, float eps,
                 wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(x); WB_NOT_NULL(gamma); WB_NOT_NULL(beta); WB_NOT_NULL(out);
        wb::layernorm(x, gamma, beta, out, out_dtype, nullptr, rows, d, eps, nullptr, S(stream));
    });
}

int wb_linear(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, const float* bias, const float* residual,
              int64_t ldres, void* out, int64_t ldo, int out_dtype, int Mrows, int N, int K, int act, int backend,
              wb_stream stream) {
    return guarded([&] {
        wb::GemmArgs a;
        a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.in_dtype = in_dtype; a.bias = bias; a.res = residual; a.ldres = ldres;
        a.out = out; a.ldo = ldo; a.out_dtype = out_dtype; a.M = Mrows; a.N = N; a.K = K; a.act = act;
        if (backend >= 100) {  // test hook: tcgen05 kernel with a fixed BLOCK_N (backend = 100 + BLOCK_N)
            wb::set_gemm_tc_block_n(backend - 100);
            try { wb::gemm_tc(a, S(stream)); } catch (...) { wb::set_gemm_tc_block_n(0); throw; }
            wb::set_gemm_tc_block_n(0);
        } else if (backend == 2) wb::gemm_tc(a, S(stream));
        else if (backend == 1) wb::gemm_simt(a, S(stream));
        else wb::gemm(a, S(stream));
    });
}

int wb_encoder_stem(wb_session* s, const float* mel, int batch, float* x_out, wb_stream stream);

int wb_linear_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, float* parts, int64_t split_stride,
                     int Mrows, int N, int K, int k_splits, int max_k_splits, int* chosen_splits, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(parts);
        wb::GemmArgs a;
        a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.in_dtype = in_dtype;
        a.out = parts; a.ldo = N; a.out_dtype = wb::F32; a.M = Mrows; a.N = N; a.K = K;
        a.k_splits = k_splits; a.max_k_splits = max_k_splits; a.split_stride = split_stride;
        int chosen = 1;
        a.chosen_splits = &chosen;
        WB_REQUIRE(split_stride >= (int64_t)Mrows * N || (k_splits == 1), "split_stride must cover one [M, N] slab");
        wb::gemm(a, S(stream));
        if (chosen_splits) *chosen_splits = chosen;
    });
}

int wb_layernorm_preadd(float* x, const float* parts, int n_parts, int64_t part_stride, const float* add_bias, const float* gamma,
                        const float* beta, void* out, int out_dtype, int rows, int d, float eps, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(x); WB_NOT_NULL(gamma); WB_NOT_NULL(beta); WB_NOT_NULL(out);
        wb::LnPreAdd pre;
        pre.parts = parts; pre.n_parts = n_parts; pre.part_stride = part_stride; pre.bias = add_bias;
        wb::layernorm_preadd(x, pre, gamma, beta, out, out_dtype, rows, d, eps, nullptr, S(stream));
    });
}

int wb_encoder_attention(const void* qkv, void* out, int dtype, int batch, int seq, int heads, int backend, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(qkv); WB_NOT_NULL(out);
        if (backend == 1) wb::encoder_attention_simt(qkv, out, dtype, batch, seq, heads, S(stream));
        else if (backend == 2) { WB_REQUIRE(dtype == wb::BF16, "tcgen05 attention is bf16 only"); wb::encoder_attention_tc(qkv, out, batch, seq, heads, S(stream)); }
        else wb::encoder_attention(qkv, out, dtype, batch, seq, heads, S(stream));
    });
}

int wb_decode_attention(const void* q, const void* k, const void* v, void* out, int dtype, int batch, int heads, int n_keys,
                        int64_t kv_batch_stride, int64_t kv_head_stride, wb_stream stream) {
    return guarded([&] {
        wb::DecAttnArgs a;
        a.dtype = dtype; a.q = q; a.q_stride = (long long)heads * 64; a.out = out; a.out_stride = (long long)heads * 64;
        a.B = batch; a.H = heads; a.k = k; a.v = v; a.kv_bstride = kv_batch_stride; a.kv_hstride = kv_head_stride;
        a.n_keys = n_keys;
        wb::decode_attention(a, S(stream));
    });
}

int wb_paged_self_attention(const void* qkv, int64_t row_stride, void* out, void* k_pages, void* v_pages, const int32_t* page_table,
                            int pages_per_seq, int dtype, int batch, int heads, const int32_t* state, const int32_t* row_active,
                            wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(qkv); WB_NOT_NULL(out); WB_NOT_NULL(k_pages); WB_NOT_NULL(v_pages); WB_NOT_NULL(page_table); WB_NOT_NULL(state);
        WB_REQUIRE(dtype == wb::F32 || dtype == wb::BF16, "dtype must be 0 (fp32) or 1 (bf16)");
        const size_t es = wb::dtype_size(dtype);
        const long long d = (long long)heads * 64;
        wb::DecAttnArgs a;
        a.dtype = dtype; a.q = qkv; a.q_stride = row_stride; a.out = out; a.out_stride = d; a.B = batch; a.H = heads;
        a.state = reinterpret_cast<const wb::StepState*>(state); a.row_active = row_active;
        a.k_new = (const uint8_t*)qkv + d * es; a.v_new = (const uint8_t*)qkv + 2 * d * es; a.new_stride = row_stride;
        a.k_pages = k_pages; a.v_pages = v_pages; a.page_table = page_table; a.pages_per_seq = pages_per_seq; a.page_tokens = 64;
        wb::decode_attention(a, S(stream));
    });
}

int wb_argmax(const float* logits, int64_t ld, int batch, int vocab, const uint8_t* mask, int bits, int32_t* out, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(logits); WB_NOT_NULL(out);
        wb::argmax_rows(logits, ld, batch, vocab, mask, bits, out, S(stream));
    });
}

int wb_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(in); WB_NOT_NULL(out);
        wb::cast(in, in_dtype, out, out_dtype, n, S(stream));
    });
}

int wb_embed(const int32_t* ids, int64_t ids_stride, int batch, int n_tokens, int position0, const void* embed_tokens,
             const void* embed_positions, int dtype, int d_model, int vocab_size, float* x_out, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(ids); WB_NOT_NULL(embed_tokens); WB_NOT_NULL(embed_positions); WB_NOT_NULL(x_out);
        wb::embed_tokens(ids, ids_stride, batch, n_tokens, position0, embed_tokens, embed_positions, dtype, x_out, d_model,
                         vocab_size, S(stream));
    });
}

int wb_kv_append(const void* past, int64_t past_batch_stride, int64_t past_head_stride, int cache_len, const void* current,
                 int64_t current_batch_stride, void* out, int dtype, int batch, int heads, wb_stream stream) {
    return guarded([&] {
        wb::kv_append(past, past_batch_stride, past_head_stride, current, current_batch_stride, out, dtype, batch, heads,
                      cache_len, S(stream));
    });
}

int wb_linear_split_heads(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, const float* bias, void* out,
                          int out_dtype, int batch, int seq, int heads, int K, wb_stream stream) {
    return guarded([&] {
        WB_REQUIRE(batch > 0 && seq > 0 && heads > 0, "bad split-heads shape");
        wb::GemmArgs a;
        a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.in_dtype = in_dtype; a.bias = bias;
        a.out = out; a.ldo = 0; a.out_dtype = out_dtype; a.M = batch * seq; a.N = heads * 64; a.K = K;
        a.m_period_in = seq; a.m_valid = seq; a.m_period_out = seq;
        a.out_mode = 1; a.hs_heads = heads; a.hs_batch = batch; a.hs_b0 = 0;
        wb::gemm(a, S(stream));
    });
}

int wb_conv_stem_workspace_bytes(int batch, int d_model, int n_frames, int dtype, size_t* bytes) {
    return guarded([&] {
        WB_NOT_NULL(bytes);
        WB_REQUIRE(batch > 0 && d_model > 0 && n_frames > 0, "bad stem shape");
        *bytes = wb::conv_stem_a1_bytes(batch, n_frames, dtype) + wb::conv_stem_h1p_bytes(batch, d_model, dtype) + 4096;
    });
}

int wb_conv_stem(const float* mel, int batch, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                 const float* positions, int dtype, int d_model, int num_mel_bins, int n_frames, void* workspace,
                 size_t workspace_bytes, float* x_out, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(mel); WB_NOT_NULL(w1_packed); WB_NOT_NULL(b1); WB_NOT_NULL(w2_packed); WB_NOT_NULL(b2);
        WB_NOT_NULL(positions); WB_NOT_NULL(workspace); WB_NOT_NULL(x_out);
        WB_REQUIRE(batch > 0 && d_model % 64 == 0 && 3 * num_mel_bins <= wb::CONV1_KPAD, "bad stem shape");
        WB_REQUIRE(n_frames % 2 == 0 && n_frames + 8 <= wb::H1_ROWS && n_frames / 2 <= wb::CONV2_MPERIOD, "unsupported audio window");
        const size_t a1b = (wb::conv_stem_a1_bytes(batch, n_frames, dtype) + 1023) / 1024 * 1024;
        const size_t h1b = wb::conv_stem_h1p_bytes(batch, d_model, dtype);
        uint8_t* base = (uint8_t*)workspace;
        const size_t skew = (1024 - (reinterpret_cast<uintptr_t>(base) & 1023)) & 1023;
        WB_REQUIRE(workspace_bytes >= skew + a1b + h1b, "stem workspace too small (see wb_conv_stem_workspace_bytes)");
        void* a1 = base + skew;
        void* h1p = base + skew + a1b;
        WB_CHECK_CUDA(cudaMemsetAsync(h1p, 0, h1b, S(stream)));
        wb::StemWeights w;
        w.w1 = w1_packed; w.b1 = b1; w.w2 = w2_packed; w.b2 = b2; w.pos = positions;
        w.dtype = dtype; w.d = d_model; w.n_mels = num_mel_bins; w.n_frames = n_frames;
        wb::conv_stem(w, mel, batch, a1, h1p, x_out, S(stream));
    });
}

int wb_encoder_stem(wb_session* s, const float* mel, int batch, float* x_out, wb_stream stream) {
    return guarded([&] {
        WB_NOT_NULL(s); WB_NOT_NULL(mel); WB_NOT_NULL(x_out);
        SS(s)->stem(mel, batch, x_out, S(stream));
    });
}

}  // extern "C"

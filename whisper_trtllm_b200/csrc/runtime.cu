// Native runtime: packed weights, encoder pass, cross-K/V projection, paged self-KV, on-device greedy loop.
//
// Weight-name mapping follows the reference's binders: build_encoder.py:71-91 (q/k/v fused into one `qkv`
// weight with a ZERO K bias) and build_decoder.py:71-101; oracle key names per SURVEY.md Appendix B.
// The q scale head_dim^-0.5 = 0.125 (modeling_whisper.py:472,572) is folded into the packed q weights
// and bias: multiplying by a power of two commutes with every rounding step, so results are bit-identical
// to scaling after the projection.
#include "wb_runtime.h"

#include <cstring>

namespace wb {

// ------------------------------------------------------------------------------------------------ Model
void* Model::dalloc(size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 255) / 256 * 256;
    WB_CHECK_CUDA(cudaMalloc(&p, bytes));
    WB_CHECK_CUDA(cudaMemset(p, 0, bytes));
    allocs.push_back(p);
    weight_bytes += bytes;
    return p;
}

Model::Model(const ModelConfig& c, int dt) : cfg(c), dtype(dt) {
    WB_REQUIRE(dt == F32 || dt == BF16, "dtype must be 0 (fp32) or 1 (bf16)");
    WB_REQUIRE(c.d_model > 0 && c.n_heads > 0 && c.d_model == c.n_heads * 64, "head_dim must be 64");
    WB_REQUIRE(c.d_model % 64 == 0 && c.d_model <= 1024, "d_model must be a multiple of 64, <= 1024");
    WB_REQUIRE(c.ffn % 64 == 0 && c.vocab % 8 == 0, "ffn % 64 == 0 and vocab % 8 == 0 required");
    WB_REQUIRE(c.n_frames == 2 * c.n_ctx && c.n_frames + 8 <= H1_ROWS && c.n_ctx <= CONV2_MPERIOD, "unsupported audio window");
    WB_REQUIRE(3 * c.n_mels <= CONV1_KPAD, "too many mel bins");
    WB_REQUIRE(c.max_length >= 2 && c.max_length <= c.max_tgt && c.max_tgt % PAGE_TOKENS == 0, "bad max_length / max_target_positions");
    const size_t es = dtype_size(dtype);
    const int d = c.d_model, F = c.ffn;
    auto lin = [&](Linear& l, int n, int k) {
        l.n = n; l.k = k;
        l.w = dalloc((size_t)n * k * es);
        l.b = (float*)dalloc((size_t)n * 4);
    };
    auto ln = [&](LNorm& l) {
        l.g = (float*)dalloc((size_t)d * 4);
        l.b = (float*)dalloc((size_t)d * 4);
    };
    lin(conv1, d, CONV1_KPAD);
    lin(conv2, d, 3 * d);
    enc_pos = (float*)dalloc((size_t)c.n_ctx * d * 4);
    enc.resize(c.enc_layers);
    for (auto& L : enc) { ln(L.ln1); lin(L.qkv, 3 * d, d); lin(L.out, d, d); ln(L.ln2); lin(L.fc1, F, d); lin(L.fc2, d, F); }
    ln(enc_ln);
    emb = dalloc((size_t)c.vocab * d * es);
    dec_pos = dalloc((size_t)c.max_tgt * d * es);
    dec.resize(c.dec_layers);
    for (auto& L : dec) {
        ln(L.ln1); lin(L.qkv, 3 * d, d); lin(L.out, d, d);
        ln(L.ln2); lin(L.cq, d, d); lin(L.ckv, 2 * d, d); lin(L.cout, d, d);
        ln(L.ln3); lin(L.fc1, F, d); lin(L.fc2, d, F);
    }
    ln(dec_ln);
    vocab_mask = (unsigned char*)dalloc((size_t)c.vocab);
    force_map = (int*)dalloc((size_t)(c.max_tgt + 1) * 4);
    std::vector<int> fm(c.max_tgt + 1, -1);
    WB_CHECK_CUDA(cudaMemcpy(force_map, fm.data(), fm.size() * 4, cudaMemcpyHostToDevice));
}

Model::~Model() {
    for (void* p : allocs) cudaFree(p);
}

int Model::expected_tensors() const { return 5 + cfg.enc_layers * 15 + 2 + 2 + cfg.dec_layers * 24 + 2; }

void Model::check_complete() const {
    if (loaded != expected_tensors())
        throw Error(-4, "model is missing weights: loaded " + std::to_string(loaded) + " of " + std::to_string(expected_tensors()) + " tensors");
}

namespace {
// copy `numel` host floats (optionally scaled) into dst (compute dtype) at element offset `off`
void upload(int dtype, void* dst, size_t off, const float* host, size_t numel, float scale = 1.0f) {
    std::vector<float> scaled;
    if (scale != 1.0f) {
        scaled.resize(numel);
        for (size_t i = 0; i < numel; ++i) scaled[i] = host[i] * scale;
        host = scaled.data();
    }
    if (dtype == F32) {
        WB_CHECK_CUDA(cudaMemcpy((float*)dst + off, host, numel * 4, cudaMemcpyHostToDevice));
    } else {
        float* tmp = nullptr;
        WB_CHECK_CUDA(cudaMalloc(&tmp, numel * 4));
        cudaError_t e = cudaMemcpy(tmp, host, numel * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            try {
                cast(tmp, F32, (bf16*)dst + off, BF16, (long long)numel, 0);
                e = cudaStreamSynchronize(0);
            } catch (...) { cudaFree(tmp); throw; }
        }
        cudaFree(tmp);
        WB_CHECK_CUDA(e);
    }
}
void upload_f32(float* dst, size_t off, const float* host, size_t numel, float scale = 1.0f) { upload(F32, dst, off, host, numel, scale); }

bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
}  // namespace

void Model::load_tensor(const std::string& name, const float* host, long long numel) {
    WB_REQUIRE(host != nullptr && numel > 0, "empty tensor " + name);
    const int d = cfg.d_model, F = cfg.ffn;
    const float qscale = 0.125f;
    auto expect = [&](long long n) {
        if (numel != n) throw Error(-1, "tensor " + name + ": expected " + std::to_string(n) + " elements, got " + std::to_string(numel));
    };
    auto load_ln = [&](LNorm& l, const std::string& leaf) {
        expect(d);
        upload_f32(leaf == "weight" ? l.g : l.b, 0, host, d);
    };
    auto load_lin = [&](Linear& l, const std::string& leaf, int row0, int rows, float scale) {
        if (leaf == "weight") { expect((long long)rows * l.k); upload(dtype, l.w, (size_t)row0 * l.k, host, (size_t)rows * l.k, scale); }
        else { expect(rows); upload_f32(l.b, row0, host, rows, scale); }
    };
    // attention projections of one block: fused [q*0.125 | k | v] rows (K has no bias: its slice stays zero)
    auto load_attn = [&](const std::string& proj, const std::string& leaf, Linear& qkv_or_q, Linear* kv, Linear& outp) {
        if (proj == "q_proj") load_lin(qkv_or_q, leaf, 0, d, qscale);
        else if (proj == "k_proj") { if (kv) load_lin(*kv, leaf, 0, d, 1.f); else load_lin(qkv_or_q, leaf, d, d, 1.f); }
        else if (proj == "v_proj") { if (kv) load_lin(*kv, leaf, d, d, 1.f); else load_lin(qkv_or_q, leaf, 2 * d, d, 1.f); }
        else if (proj == "out_proj") load_lin(outp, leaf, 0, d, 1.f);
        else throw Error(-1, "unknown attention tensor " + name);
    };

    if (name == "proj_out.weight") return;  // tied to embed_tokens (modeling_whisper.py:1335); not counted
    const size_t dot = name.rfind('.');
    WB_REQUIRE(dot != std::string::npos, "bad tensor name " + name);
    const std::string leaf = name.substr(dot + 1);
    const std::string stem = name.substr(0, dot);

    if (stem == "model.encoder.conv1") {
        if (leaf == "bias") { expect(d); upload_f32(conv1.b, 0, host, d); }
        else {  // [d, n_mels, 3] -> [d, KPAD] with k = tap * n_mels + c
            expect((long long)d * cfg.n_mels * 3);
            std::vector<float> w((size_t)d * CONV1_KPAD, 0.f);
            for (int n = 0; n < d; ++n)
                for (int c = 0; c < cfg.n_mels; ++c)
                    for (int t = 0; t < 3; ++t) w[(size_t)n * CONV1_KPAD + t * cfg.n_mels + c] = host[((size_t)n * cfg.n_mels + c) * 3 + t];
            upload(dtype, conv1.w, 0, w.data(), w.size());
        }
    } else if (stem == "model.encoder.conv2") {
        if (leaf == "bias") { expect(d); upload_f32(conv2.b, 0, host, d); }
        else {  // [d, d, 3] -> [d, 3d] with k = tap * d + c
            expect((long long)d * d * 3);
            std::vector<float> w((size_t)d * 3 * d);
            for (int n = 0; n < d; ++n)
                for (int c = 0; c < d; ++c)
                    for (int t = 0; t < 3; ++t) w[(size_t)n * 3 * d + (size_t)t * d + c] = host[((size_t)n * d + c) * 3 + t];
            upload(dtype, conv2.w, 0, w.data(), w.size());
        }
    } else if (stem == "model.encoder.embed_positions") {
        expect((long long)cfg.n_ctx * d);
        upload_f32(enc_pos, 0, host, (size_t)cfg.n_ctx * d);
    } else if (stem == "model.encoder.layer_norm") {
        load_ln(enc_ln, leaf);
    } else if (stem == "model.decoder.layer_norm") {
        load_ln(dec_ln, leaf);
    } else if (stem == "model.decoder.embed_tokens") {
        expect((long long)cfg.vocab * d);
        upload(dtype, emb, 0, host, (size_t)cfg.vocab * d);
    } else if (stem == "model.decoder.embed_positions") {
        expect((long long)cfg.max_tgt * d);
        upload(dtype, dec_pos, 0, host, (size_t)cfg.max_tgt * d);
    } else if (starts_with(stem, "model.encoder.layers.") || starts_with(stem, "model.decoder.layers.")) {
        const bool is_enc = stem[6] == 'e';
        const size_t p0 = std::strlen("model.encoder.layers.");
        const size_t p1 = stem.find('.', p0);
        WB_REQUIRE(p1 != std::string::npos, "bad layer tensor name " + name);
        const int li = std::stoi(stem.substr(p0, p1 - p0));
        const std::string rest = stem.substr(p1 + 1);
        WB_REQUIRE(li >= 0 && li < (is_enc ? cfg.enc_layers : cfg.dec_layers), "layer index out of range in " + name);
        if (is_enc) {
            EncLayer& L = enc[li];
            if (rest == "self_attn_layer_norm") load_ln(L.ln1, leaf);
            else if (rest == "final_layer_norm") load_ln(L.ln2, leaf);
            else if (rest == "fc1") load_lin(L.fc1, leaf, 0, F, 1.f);
            else if (rest == "fc2") load_lin(L.fc2, leaf, 0, d, 1.f);
            else if (starts_with(rest, "self_attn.")) load_attn(rest.substr(10), leaf, L.qkv, nullptr, L.out);
            else throw Error(-1, "unknown tensor " + name);
        } else {
            DecLayer& L = dec[li];
            if (rest == "self_attn_layer_norm") load_ln(L.ln1, leaf);
            else if (rest == "encoder_attn_layer_norm") load_ln(L.ln2, leaf);
            else if (rest == "final_layer_norm") load_ln(L.ln3, leaf);
            else if (rest == "fc1") load_lin(L.fc1, leaf, 0, F, 1.f);
            else if (rest == "fc2") load_lin(L.fc2, leaf, 0, d, 1.f);
            else if (starts_with(rest, "self_attn.")) load_attn(rest.substr(10), leaf, L.qkv, nullptr, L.out);
            else if (starts_with(rest, "encoder_attn.")) load_attn(rest.substr(13), leaf, L.cq, &L.ckv, L.cout);
            else throw Error(-1, "unknown tensor " + name);
        }
    } else {
        throw Error(-1, "unknown tensor " + name);
    }
    ++loaded;
}

void Model::set_generation(const int* suppress, int n_suppress, const int* begin_suppress, int n_begin, int begin_index,
                           const int* forced_pairs, int n_forced) {
    std::vector<unsigned char> mask(cfg.vocab, 0);
    for (int i = 0; i < n_suppress; ++i) {
        WB_REQUIRE(suppress[i] >= 0 && suppress[i] < cfg.vocab, "suppress token out of range");
        mask[suppress[i]] |= 1;
    }
    for (int i = 0; i < n_begin; ++i) {
        WB_REQUIRE(begin_suppress[i] >= 0 && begin_suppress[i] < cfg.vocab, "begin-suppress token out of range");
        mask[begin_suppress[i]] |= 2;
    }
    WB_CHECK_CUDA(cudaMemcpy(vocab_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
    std::vector<int> fm(cfg.max_tgt + 1, -1);
    for (int i = 0; i < n_forced; ++i) {
        const int idx = forced_pairs[2 * i], tok = forced_pairs[2 * i + 1];
        WB_REQUIRE(idx >= 0 && idx <= cfg.max_tgt && tok >= 0 && tok < cfg.vocab, "forced decoder id out of range");
        fm[idx] = tok;
    }
    WB_CHECK_CUDA(cudaMemcpy(force_map, fm.data(), fm.size() * 4, cudaMemcpyHostToDevice));
    cfg.begin_index = begin_index;
    ++generation_version;
}

// ------------------------------------------------------------------------------------------------ Session
namespace {
struct Carver {
    uint8_t* base; size_t off = 0, cap;
    Carver(void* p, size_t c) : base((uint8_t*)p), cap(c) {}
    template <typename P> void take(P*& field, size_t bytes) {
        off = (off + 1023) / 1024 * 1024;
        field = base ? reinterpret_cast<P*>(base + off) : nullptr;
        off += bytes;
        if (base && off > cap) throw Error(-1, "workspace too small: need at least " + std::to_string(off) + " bytes");
    }
};

// single source of truth for the workspace layout (base == nullptr -> sizing pass)
size_t layout(Buffers& b, const Model* m, int max_batch, int enc_chunk, void* base, size_t cap) {
    const ModelConfig& g = m->cfg;
    const size_t es = dtype_size(m->dtype), d = g.d_model, Bc = enc_chunk, B = max_batch;
    const size_t Mc = Bc * g.n_ctx, L = g.dec_layers, H = g.n_heads;
    const size_t pages = B * (g.max_tgt / PAGE_TOKENS);
    Carver c(base, cap);
    c.take(b.a1, Bc * g.n_frames * CONV1_KPAD * es);
    c.take(b.h1p, (Bc * H1_ROWS + 8) * d * es);
    c.take(b.x, Mc * d * 4);
    c.take(b.ln, Mc * d * es);
    c.take(b.qkv, Mc * 3 * d * es);
    c.take(b.att, Mc * d * es);
    c.take(b.ffn, Mc * g.ffn * es);
    c.take(b.enc, B * g.n_ctx * d * es);
    c.take(b.cross, L * 2 * B * H * g.n_ctx * 64 * es);
    c.take(b.self_k, L * pages * H * PAGE_TOKENS * 64 * es);
    c.take(b.self_v, L * pages * H * PAGE_TOKENS * 64 * es);
    c.take(b.page_table, pages * 4);
    c.take(b.dx, B * d * 4);
    c.take(b.dpart, (size_t)MAX_K_SPLITS * B * d * 4);
    c.take(b.dln, B * d * es);
    c.take(b.dqkv, B * 3 * d * es);
    c.take(b.datt, B * d * es);
    c.take(b.dq, B * d * es);
    c.take(b.dffn, B * g.ffn * es);
    c.take(b.logits, B * (size_t)g.vocab * 4);
    c.take(b.tokens, B * g.max_tgt * 4);
    c.take(b.unfinished, B * 4);
    c.take(b.state, sizeof(StepState));
    c.take(b.mega_part, mega_part_bytes(max_batch, g.n_heads));
    c.take(b.mega_sync, mega_sync_bytes(max_batch, g.n_heads));
    c.take(b.mega_table, mega_table_bytes(g.dec_layers));
    c.take(b.result_tokens, B * g.max_tgt * 4);
    c.take(b.chain_sync, chain_sync_bytes());
    c.take(b.row_len, B * 4);
    return c.off + 1024;
}
}  // namespace

size_t Session::workspace_bytes(const Model* m, int max_batch, int enc_chunk) {
    Buffers b;
    return layout(b, m, max_batch, enc_chunk, nullptr, 0);
}

Session::Session(Model* model, int mb, int ec, void* workspace, size_t workspace_bytes) : m(model), max_batch(mb), enc_chunk(ec) {
    WB_REQUIRE(model != nullptr && workspace != nullptr, "null model / workspace");
    WB_REQUIRE(mb > 0 && ec > 0 && ec <= mb, "need 0 < enc_chunk <= max_batch");
    WB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    model->check_complete();
    // align the carve base to 1024 so every buffer is TMA-friendly
    uint8_t* base = (uint8_t*)workspace;
    const size_t skew = (1024 - (reinterpret_cast<uintptr_t>(base) & 1023)) & 1023;
    WB_REQUIRE(workspace_bytes > skew, "workspace too small");
    layout(*this, model, mb, ec, base + skew, workspace_bytes - skew);
    pages_per_seq = model->cfg.max_tgt / PAGE_TOKENS;
    num_pages = mb * pages_per_seq;
    // page allocator: every row owns its pages up front (identity map); the kernels only see the table
    page_table_host.resize((size_t)num_pages);
    for (int i = 0; i < num_pages; ++i) page_table_host[i] = i;
    WB_CHECK_CUDA(cudaMemcpy(page_table, page_table_host.data(), page_table_host.size() * 4, cudaMemcpyHostToDevice));
    WB_CHECK_CUDA(cudaMemset(h1p, 0, ((size_t)ec * H1_ROWS + 8) * model->cfg.d_model * dtype_size(model->dtype)));
    WB_CHECK_CUDA(cudaMemset(state, 0, sizeof(StepState)));
    WB_CHECK_CUDA(cudaMemset(mega_sync, 0, mega_sync_bytes(mb, model->cfg.n_heads)));
    build_mega_table();
    init_chain();
    WB_CHECK_CUDA(cudaMallocHost(&host_state, sizeof(StepState)));
    std::memset(host_state, 0, sizeof(StepState));
    WB_CHECK_CUDA(cudaEventCreateWithFlags(&check_event, cudaEventDisableTiming));
    WB_CHECK_CUDA(cudaEventCreateWithFlags(&fence_event, cudaEventDisableTiming));
    WB_CHECK_CUDA(cudaStreamCreateWithFlags(&loop_stream, cudaStreamNonBlocking));
}

Session::~Session() {
    if (step_graph) cudaGraphExecDestroy(step_graph);
    if (loop_stream) cudaStreamDestroy(loop_stream);
    if (fence_event) cudaEventDestroy(fence_event);
    if (host_state) cudaFreeHost(host_state);
    if (check_event) cudaEventDestroy(check_event);
    for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
}

static inline bool prof_active(const Session* s, int cls) {
    if (cls != s->prof_class) return false;
    if (s->prof_step < 0) return true;
    return cls >= PROF_ENC_GEMM && cls != PROF_LAYERNORM && cls != PROF_GREEDY ? true : s->steps_enqueued == s->prof_step;
}

void Session::prof_begin(int cls, cudaStream_t st) {
    if (!prof_active(this, cls)) return;
    if (prof_used + 2 > prof_events.size()) {
        for (int i = 0; i < 2; ++i) {
            cudaEvent_t e;
            WB_CHECK_CUDA(cudaEventCreate(&e));
            prof_events.push_back(e);
        }
    }
    WB_CHECK_CUDA(cudaEventRecord(prof_events[prof_used], st));
}
void Session::prof_end(int cls, cudaStream_t st) {
    if (!prof_active(this, cls)) return;
    WB_CHECK_CUDA(cudaEventRecord(prof_events[prof_used + 1], st));
    prof_used += 2;
}
void Session::prof_read(double* total_ms, long long* launches) {
    double total = 0;
    for (size_t i = 0; i + 1 < prof_used; i += 2) {
        WB_CHECK_CUDA(cudaEventSynchronize(prof_events[i + 1]));
        float ms = 0;
        WB_CHECK_CUDA(cudaEventElapsedTime(&ms, prof_events[i], prof_events[i + 1]));
        total += ms;
    }
    if (total_ms) *total_ms = total;
    if (launches) *launches = (long long)(prof_used / 2);
    prof_used = 0;
}

size_t Session::cross_layer_elems() const { return (size_t)2 * max_batch * m->cfg.n_heads * m->cfg.n_ctx * 64; }
size_t Session::self_layer_elems() const { return (size_t)num_pages * m->cfg.n_heads * PAGE_TOKENS * 64; }

namespace {
inline void* eoff(void* p, size_t elems, int dtype) { return (uint8_t*)p + elems * dtype_size(dtype); }

GemmArgs linear_args(const Model* m, const void* A, long long lda, const Linear& l, void* out, long long ldo, int out_dtype, int M) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.W = l.w; g.ldw = l.k; g.in_dtype = m->dtype; g.bias = l.b;
    g.out = out; g.ldo = ldo; g.out_dtype = out_dtype; g.M = M; g.N = l.n; g.K = l.k;
    return g;
}
}  // namespace

// cross-attention K/V for utterances [b0, b0+bc): one GEMM per decoder layer, scattered head-major
static void project_cross_kv(Session* s, int b0, int bc, cudaStream_t st) {
    const Model* m = s->m;
    const ModelConfig& g = m->cfg;
    const int d = g.d_model;
    for (int l = 0; l < g.dec_layers; ++l) {
        GemmArgs a = linear_args(m, eoff(s->enc, (size_t)b0 * g.n_ctx * d, m->dtype), d, m->dec[l].ckv,
                                 eoff(s->cross, (size_t)l * s->cross_layer_elems(), m->dtype), 0, m->dtype, bc * g.n_ctx);
        a.m_period_in = g.n_ctx; a.m_valid = g.n_ctx; a.m_period_out = g.n_ctx;
        a.out_mode = 1; a.hs_heads = g.n_heads; a.hs_batch = s->max_batch; a.hs_b0 = b0;
        ProfScope ps(s, PROF_CROSS_KV, st);
        gemm(a, st);
    }
}

// conv1 (+GELU) as a GEMM over the im2col'd mel; conv2 (+GELU, +positions) as a GEMM over sliding windows of
// the zero-padded, time-major conv1 output.  Result: fp32 residual stream x [bc * n_ctx, d].
// Stateless: used by the session (packed model weights) and by wb_conv_stem (module-level WhisperEncoder).
void conv_stem(const StemWeights& w, const float* mel, int bc, void* a1, void* h1p, float* x, cudaStream_t st) {
    const int d = w.d, dt = w.dtype, n_ctx = w.n_frames / 2;
    im2col_conv1(mel, a1, dt, bc, w.n_mels, w.n_frames, CONV1_KPAD, st);
    {
        GemmArgs a;
        a.A = a1; a.lda = CONV1_KPAD; a.W = w.w1; a.ldw = CONV1_KPAD; a.in_dtype = dt; a.bias = w.b1;
        a.out = h1p; a.ldo = d; a.out_dtype = dt; a.M = bc * w.n_frames; a.N = d; a.K = CONV1_KPAD;
        a.act = 1;
        a.m_period_in = w.n_frames; a.m_valid = w.n_frames; a.m_period_out = H1_ROWS; a.m_out_offset = 1;
        gemm(a, st);
    }
    {
        GemmArgs a;
        a.A = h1p; a.lda = 2 * d; a.W = w.w2; a.ldw = 3 * d; a.in_dtype = dt; a.bias = w.b2;
        a.out = x; a.ldo = d; a.out_dtype = F32; a.M = bc * CONV2_MPERIOD; a.N = d; a.K = 3 * d;
        a.act = 1;
        a.res = w.pos; a.ldres = d; a.res_periodic = 1;
        a.m_period_in = CONV2_MPERIOD; a.m_valid = n_ctx; a.m_period_out = n_ctx;
        gemm(a, st);
    }
}

size_t conv_stem_a1_bytes(int bc, int n_frames, int dtype) { return (size_t)bc * n_frames * CONV1_KPAD * dtype_size(dtype); }
size_t conv_stem_h1p_bytes(int bc, int d, int dtype) { return ((size_t)bc * H1_ROWS + 8) * d * dtype_size(dtype); }

void Session::stem_chunk(const float* mel, int bc, cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    StemWeights w;
    w.w1 = m->conv1.w; w.b1 = m->conv1.b; w.w2 = m->conv2.w; w.b2 = m->conv2.b; w.pos = m->enc_pos;
    w.dtype = m->dtype; w.d = g.d_model; w.n_mels = g.n_mels; w.n_frames = g.n_frames;
    conv_stem(w, mel, bc, a1, h1p, x, st);
}

void Session::stem(const float* mel, int B, float* x_out, cudaStream_t st) {
    WB_REQUIRE(B > 0 && B <= enc_chunk, "stem batch must be <= enc_chunk");
    stem_chunk(mel, B, st);
    WB_CHECK_CUDA(cudaMemcpyAsync(x_out, x, (size_t)B * m->cfg.n_ctx * m->cfg.d_model * 4, cudaMemcpyDeviceToDevice, st));
}

void Session::encode(const float* mel, int B, float* enc_out_f32, cudaStream_t st) { encode_at(mel, B, 0, enc_out_f32, st); }

// encoder + cross-K/V projection of B utterances into the batch slots [slot0, slot0 + B)
void Session::encode_at(const float* mel, int B, int slot0, float* enc_out_f32, cudaStream_t st) {
    WB_REQUIRE(mel != nullptr && B > 0 && slot0 >= 0 && slot0 + B <= max_batch, "bad encode batch");
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, dt = m->dtype;
    for (int c0 = 0; c0 < B; c0 += enc_chunk) {
        const int bc = std::min(enc_chunk, B - c0);
        const int b0 = slot0 + c0;
        const int M = bc * g.n_ctx;
        { ProfScope ps(this, PROF_STEM, st); stem_chunk(mel + (size_t)c0 * g.n_mels * g.n_frames, bc, st); }
        // ---- transformer layers (pre-LN), fp32 residual stream in x
        for (int l = 0; l < g.enc_layers; ++l) {
            const EncLayer& L = m->enc[l];
            layernorm(x, L.ln1.g, L.ln1.b, ln, dt, nullptr, M, d, 1e-5f, nullptr, st);
            { ProfScope ps(this, PROF_ENC_GEMM, st); gemm(linear_args(m, ln, d, L.qkv, qkv, 3 * d, dt, M), st); }
            { ProfScope ps(this, PROF_ENC_ATTN, st); encoder_attention(qkv, att, dt, bc, g.n_ctx, g.n_heads, st); }
            {
                GemmArgs a = linear_args(m, att, d, L.out, x, d, F32, M);
                a.res = x; a.ldres = d;
                ProfScope ps(this, PROF_ENC_GEMM, st);
                gemm(a, st);
            }
            layernorm(x, L.ln2.g, L.ln2.b, ln, dt, nullptr, M, d, 1e-5f, nullptr, st);
            {
                GemmArgs a = linear_args(m, ln, d, L.fc1, ffn, g.ffn, dt, M);
                a.act = 1;
                ProfScope ps(this, PROF_ENC_GEMM, st);
                gemm(a, st);
            }
            {
                GemmArgs a = linear_args(m, ffn, g.ffn, L.fc2, x, d, F32, M);
                a.res = x; a.ldres = d;
                ProfScope ps(this, PROF_ENC_GEMM, st);
                gemm(a, st);
            }
        }
        layernorm(x, m->enc_ln.g, m->enc_ln.b, eoff(enc, (size_t)b0 * g.n_ctx * d, dt), dt,
                  enc_out_f32 ? enc_out_f32 + (size_t)c0 * g.n_ctx * d : nullptr, M, d, 1e-5f, nullptr, st);
        project_cross_kv(this, b0, bc, st);
    }
}

void Session::set_encoder_output(const void* enc_states, int dtype, int B, cudaStream_t st) {
    WB_REQUIRE(enc_states != nullptr && B > 0 && B <= max_batch, "bad encoder states");
    const ModelConfig& g = m->cfg;
    cast(enc_states, dtype, enc, m->dtype, (long long)B * g.n_ctx * g.d_model, st);
    for (int b0 = 0; b0 < B; b0 += enc_chunk) project_cross_kv(this, b0, std::min(enc_chunk, B - b0), st);
}

void Session::decode_begin(int B, cudaStream_t st) {
    WB_REQUIRE(B > 0 && B <= max_batch, "bad decode batch");
    const ModelConfig& g = m->cfg;
    batch = B;
    begin_batch = B;
    compacted = false;
    row_origin.resize(B);
    for (int i = 0; i < B; ++i) row_origin[i] = i;
    steps_enqueued = 0;
    ragged = false;
    refill_mode = false;
    next_utt = B;
    greedy_init(tokens, g.max_tgt, unfinished, state, B, g.sot, g.pad, g.max_tgt, st);
    // the whole-step kernel starts from the residual stream: embedding of the start token here, of every later token by the
    // greedy kernel of the step that chose it.  ALWAYS written (one tiny kernel): which step path runs is only known when the
    // loop starts (decode_run_multi sets `exclusive`), so the decision must not depend on the previous run's state
    decoder_embed(tokens, g.max_tgt, state, m->emb, m->dec_pos, m->dtype, dx, B, g.d_model, st);
    dx_embedded = true;
}

// small batches (B <= 16, bf16) run the whole decoder for one token in ONE persistent cooperative kernel (step_mega.cu); the
// switch exists for A/B measurements and the parity tests (0 = always the multi-kernel step)
static std::atomic<bool>& whole_step_kernel_enabled() {
    static std::atomic<bool> on{true};
    return on;
}
void set_small_batch_path(bool on) { whole_step_kernel_enabled() = on; }

// (not when several sessions decode concurrently on their own streams: a cooperative grid needs every SM to itself)
// a session option (wb_session_set_option) wins over the process-wide switch; -1 = inherit
static inline bool opt_or(int session_opt, bool global) { return session_opt < 0 ? global : session_opt != 0; }
bool Session::use_mega() const {
    return opt_or(opt_small_batch_path, whole_step_kernel_enabled()) && exclusive && !ragged && get_gemm_backend() == 0 && mega_supported();
}
bool chain_path_enabled();   // step_chain.cu
bool Session::use_chain() const {
    return opt_or(opt_decode_chain_path, chain_path_enabled()) && exclusive && get_gemm_backend() == 0 && chain_supported() && !use_mega();
}
void Session::set_option(const std::string& name, int value) {
    WB_REQUIRE(value >= -1 && value <= 1, "option values: -1 (inherit the process-wide switch), 0 (off), 1 (on)");
    if (name == "small_batch_path") opt_small_batch_path = value;
    else if (name == "decode_chain_path") opt_decode_chain_path = value;
    else if (name == "cuda_graphs") opt_cuda_graphs = value;
    else if (name == "merge_attention") { opt_merge_attention = value; chain_batch = -1; step_graph_batch = -1; }   // phase tables and step graph are rebuilt
    else throw Error(-1, "unknown session option '" + name + "' (small_batch_path, decode_chain_path, cuda_graphs, merge_attention)");
}

// ids of the current decode rows -> result buffer, original row order
void Session::publish_results(cudaStream_t st) {
    const size_t row_bytes = (size_t)m->cfg.max_tgt * 4;
    bool identity = true;
    for (int i = 0; i < batch; ++i) identity = identity && row_origin[i] == i;
    if (identity) {
        WB_CHECK_CUDA(cudaMemcpyAsync(result_tokens, tokens, row_bytes * batch, cudaMemcpyDeviceToDevice, st));
    } else {
        for (int i = 0; i < batch; ++i)
            WB_CHECK_CUDA(cudaMemcpyAsync(result_tokens + (size_t)row_origin[i] * m->cfg.max_tgt, tokens + (size_t)i * m->cfg.max_tgt,
                                          row_bytes, cudaMemcpyDeviceToDevice, st));
    }
}

// rows keep[0] < keep[1] < ... move to decode rows 0, 1, ...: ids, cross K/V rows, page-table rows (swapped: the self K/V pages
// stay where they are and no page is lost), utterance ids.  Stable: row i moves to slot j <= i, which held a row that is leaving or
// has already moved.  Does not touch `batch`, `unfinished` or the lengths.
void Session::compact_rows(const std::vector<int>& keep, cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    const size_t es = dtype_size(m->dtype);
    const size_t row_bytes = (size_t)g.max_tgt * 4;
    const size_t kv_row = (size_t)g.n_heads * g.n_ctx * 64;                   // cross K (or V) elements of one utterance and layer
    const size_t per_kv = (size_t)max_batch * kv_row;
    const int old_batch = batch;
    bool moved = false;
    for (int j = 0; j < (int)keep.size(); ++j) {
        const int i = keep[j];
        if (i == j) continue;
        moved = true;
        WB_CHECK_CUDA(cudaMemcpyAsync(tokens + (size_t)j * g.max_tgt, tokens + (size_t)i * g.max_tgt, row_bytes, cudaMemcpyDeviceToDevice, st));
        for (int l = 0; l < g.dec_layers; ++l) {
            for (int kv = 0; kv < 2; ++kv) {
                uint8_t* base_l = (uint8_t*)cross + ((size_t)l * cross_layer_elems() + (size_t)kv * per_kv) * es;
                WB_CHECK_CUDA(cudaMemcpyAsync(base_l + (size_t)j * kv_row * es, base_l + (size_t)i * kv_row * es, kv_row * es,
                                              cudaMemcpyDeviceToDevice, st));
            }
        }
        for (int q = 0; q < pages_per_seq; ++q)                               // the self K/V pages stay where they are
            std::swap(page_table_host[(size_t)j * pages_per_seq + q], page_table_host[(size_t)i * pages_per_seq + q]);
        row_origin[j] = row_origin[i];
    }
    row_origin.resize(keep.size());
    if (moved) {
        WB_CHECK_CUDA(cudaMemcpyAsync(page_table, page_table_host.data(), (size_t)old_batch * pages_per_seq * 4, cudaMemcpyHostToDevice, st));
        WB_CHECK_CUDA(cudaStreamSynchronize(st));                             // page_table_host may change again before the copy ran
    }
}

int Session::decode_compact(cudaStream_t st) {
    WB_REQUIRE(batch > 0, "decode_begin was not called");
    WB_REQUIRE(forced_tokens == nullptr && logits_dump == nullptr, "compaction is not available with teacher forcing / logits dumps");
    WB_REQUIRE(!refill_mode, "wb_decode_compact keeps results by ORIGINAL row: after wb_decode_refill use wb_decode_refill (n_new = 0) instead");
    const ModelConfig& g = m->cfg;
    std::vector<int> unf((size_t)batch);
    StepState hs;
    WB_CHECK_CUDA(cudaMemcpyAsync(unf.data(), unfinished, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(&hs, state, sizeof(StepState), cudaMemcpyDeviceToHost, st));
    WB_CHECK_CUDA(cudaStreamSynchronize(st));
    if (hs.active == 0) return 0;
    std::vector<int> keep;
    for (int i = 0; i < batch; ++i)
        if (unf[i] != 0) keep.push_back(i);
    if ((int)keep.size() == batch) return batch;
    // the ids of every current row go to the result buffer first (finished rows are final; running rows are refreshed at the end)
    publish_results(st);
    compacted = true;
    compact_rows(keep, st);
    batch = (int)keep.size();
    std::vector<int> ones((size_t)batch, 1);
    WB_CHECK_CUDA(cudaMemcpyAsync(unfinished, ones.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaStreamSynchronize(st));                                 // the host vector above goes out of scope
    // the whole-step kernel starts from the residual stream: re-embed the last token of the rows that moved
    decoder_embed(tokens, g.max_tgt, state, m->emb, m->dec_pos, m->dtype, dx, batch, g.d_model, st, ragged_len());
    dx_embedded = true;
    return batch;
}

// In-flight refill (the idea of docs/in_flight_batching.md of the reference tree, which its Whisper path does
// not have: it transcribes one utterance at a time, run.py:263-288).  Called between two windows of decode steps:
//   1. utterances that have emitted EOS (or are full) are handed to the caller with their ids and leave the batch;
//   2. the rows still decoding move to the front (compact_rows);
//   3. up to n_new NEW utterances are encoded straight into the freed slots (cross K/V projected there), start token in place;
//   4. from now on the rows are at different positions: per-row lengths (row_len) drive the embedding, the paged self-attention
//      and the logits processors / argmax (`ragged`).  When nothing was left running the batch restarts uniform.
// Rows are independent, so every utterance gets exactly the ids it would get alone (fp32: bit-identical to the reference).
int Session::decode_refill(const float* mel_new, int n_new, std::vector<Finished>& finished, int* n_admitted, cudaStream_t st) {
    WB_REQUIRE(begin_batch > 0, "decode_begin was not called");
    WB_REQUIRE(forced_tokens == nullptr && logits_dump == nullptr, "refill is not available with teacher forcing / logits dumps");
    WB_REQUIRE(n_new >= 0 && (n_new == 0 || mel_new != nullptr), "bad refill arguments");
    const ModelConfig& g = m->cfg;
    finished.clear();
    refill_mode = true;
    compacted = false;
    // ---- 1. snapshot of the loop state
    std::vector<int> unf((size_t)batch), len((size_t)batch);
    StepState hs;
    WB_CHECK_CUDA(cudaMemcpyAsync(&hs, state, sizeof(StepState), cudaMemcpyDeviceToHost, st));
    if (batch > 0) {
        WB_CHECK_CUDA(cudaMemcpyAsync(unf.data(), unfinished, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
        if (ragged) WB_CHECK_CUDA(cudaMemcpyAsync(len.data(), row_len, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    }
    WB_CHECK_CUDA(cudaStreamSynchronize(st));
    std::vector<int> keep, gone;
    for (int i = 0; i < batch; ++i) {
        if (!ragged) len[i] = hs.cur_len;
        const bool full = len[i] >= g.max_length;
        if (unf[i] != 0 && !full) keep.push_back(i);           // still decoding (the window's step budget ran out)
        else gone.push_back(i);
    }
    // ---- 2. finished utterances -> caller
    if (!gone.empty()) {
        std::vector<int> rows(gone.size() * (size_t)g.max_tgt);
        for (size_t k = 0; k < gone.size(); ++k)
            WB_CHECK_CUDA(cudaMemcpyAsync(rows.data() + k * g.max_tgt, tokens + (size_t)gone[k] * g.max_tgt, (size_t)g.max_tgt * 4,
                                          cudaMemcpyDeviceToHost, st));
        WB_CHECK_CUDA(cudaStreamSynchronize(st));
        for (size_t k = 0; k < gone.size(); ++k) {
            Finished f;
            f.utt = row_origin[gone[k]];
            int n = std::min(len[gone[k]], g.max_length);
            const int* r = rows.data() + k * g.max_tgt;
            if (!ragged)                                         // uniform batch: a row that finished early was padded after its EOS
                for (int t = 1; t < n; ++t)
                    if (r[t] == g.eos) { n = t + 1; break; }
            f.len = n;
            f.ids.assign(r, r + n);
            finished.push_back(std::move(f));
        }
    }
    // ---- 3. survivors to the front
    compact_rows(keep, st);
    std::vector<int> new_len(keep.size());
    for (size_t j = 0; j < keep.size(); ++j) new_len[j] = len[keep[j]];
    const int n_live = (int)keep.size();
    // ---- 4. new utterances into the freed slots
    const int n_adm = std::min(n_new, max_batch - n_live);
    if (n_admitted) *n_admitted = n_adm;
    if (n_adm > 0) {
        encode_at(mel_new, n_adm, n_live, nullptr, st);
        std::vector<int> first((size_t)g.max_tgt, g.pad);
        first[0] = g.sot;
        for (int k = 0; k < n_adm; ++k) {
            WB_CHECK_CUDA(cudaMemcpyAsync(tokens + (size_t)(n_live + k) * g.max_tgt, first.data(), (size_t)g.max_tgt * 4, cudaMemcpyHostToDevice, st));
            row_origin.push_back(next_utt++);
            new_len.push_back(1);
        }
        WB_CHECK_CUDA(cudaStreamSynchronize(st));               // `first` goes out of scope
    }
    batch = n_live + n_adm;
    if (batch == 0) {                                           // nothing left and nothing new: the loop is over
        hs.active = 0;
        WB_CHECK_CUDA(cudaMemcpyAsync(state, &hs, sizeof(StepState), cudaMemcpyHostToDevice, st));
        WB_CHECK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    // ---- 5. loop state of the new batch
    bool uniform = true;
    for (int i = 1; i < batch; ++i) uniform = uniform && new_len[i] == new_len[0];
    ragged = !uniform;
    hs.cur_len = uniform ? new_len[0] : std::max(hs.cur_len, 1);
    hs.active = 1; hs.final_len = 0; hs.done_counter = 0; hs.n_unfinished = batch;
    std::vector<int> ones((size_t)batch, 1);
    WB_CHECK_CUDA(cudaMemcpyAsync(state, &hs, sizeof(StepState), cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(unfinished, ones.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(row_len, new_len.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaStreamSynchronize(st));
    // the chain / whole-step paths start from the residual stream: embed the last token of every row at ITS position
    decoder_embed(tokens, g.max_tgt, state, m->emb, m->dec_pos, m->dtype, dx, batch, g.d_model, st, ragged_len());
    dx_embedded = true;
    return batch;
}

void Session::decode_step(cudaStream_t st) {
    WB_REQUIRE(batch > 0, "decode_begin was not called");
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, dt = m->dtype, B = batch;
    const int mode = step_mode() & 15;
    const bool mega = mode != 0;    // the step starts from the residual stream and its greedy kernel embeds the chosen token
    WB_REQUIRE(!ragged || forced_tokens == nullptr, "teacher forcing is not available on a refilled (ragged) batch");
    if (!in_loop_step) fuse_argmax_now = false;     // a caller's own wb_decode_step leaves the logits of the step in place (wb_decode_logits)
    prepare_step(st);
    if (mode == 1) decode_step_mega(st);
    else if (mode == 2) decode_step_chain(st);
    else decode_step_large(st);
    // logits -> processors -> argmax -> EOS / length bookkeeping, common to both paths
    if (logits_dump != nullptr && steps_enqueued < logits_dump_steps)
        WB_CHECK_CUDA(cudaMemcpyAsync(logits_dump + (size_t)steps_enqueued * B * g.vocab, logits, (size_t)B * g.vocab * 4,
                                      cudaMemcpyDeviceToDevice, st));
    {
        GreedyArgs a;
        a.logits = logits; a.ld = g.vocab; a.B = B; a.V = g.vocab;
        a.vocab_mask = m->vocab_mask; a.begin_index = g.begin_index; a.force_map = m->force_map;
        a.pad_id = g.pad; a.eos_id = g.eos; a.max_length = g.max_length;
        a.tokens = tokens; a.tokens_stride = g.max_tgt; a.unfinished = unfinished; a.state = state;
        a.forced_tokens = forced_tokens;
        a.row_len = ragged ? row_len : nullptr;
        if (fuse_argmax_now && mode != 1) {     // (the whole-step kernel computes its own LM head and writes logits)
            const int stride = gemm_argmax_partials(g.vocab);
            a.part_val = logits; a.part_idx = reinterpret_cast<const int*>(logits + (size_t)max_batch * stride);
            a.n_parts = argmax_parts; a.part_stride = stride;
        }
        // whole-step kernel: the greedy kernel also embeds the token it chose (the next step starts from the residual stream)
        if (mega) { a.embed_x = dx; a.embed_table = m->emb; a.embed_pos = m->dec_pos; a.embed_d = d; }
        ProfScope ps(this, PROF_GREEDY, st);
        greedy_step(a, st);
    }
    dx_embedded = mega;   // the greedy kernel of a whole-step launch wrote the next token's embedding
    ++steps_enqueued;
}

void Session::decode_step_large(cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    const int d = g.d_model, dt = m->dtype, B = batch;
    const int* active = &state->active;
    auto lin = [&](const void* A, long long lda, const Linear& l, void* out, long long ldo, int odt, int act, const float* res) {
        GemmArgs a = linear_args(m, A, lda, l, out, ldo, odt, B);
        a.act = act; a.res = res; a.ldres = d; a.active = active;
        ProfScope ps(this, PROF_DEC_GEMM, st);
        gemm(a, st);
    };
    // Residual GEMMs (out-proj, cross-out, fc2) and the cross-attention q projection are split-K with DEFERRED
    // reduction: they store raw fp32 partial slabs into dpart, and their consumer (the next LayerNorm, the cross-attention
    // q load) adds bias + slabs in a fixed order.  One kernel per GEMM, all SMs busy, deterministic.
    const long long part_stride = (long long)B * d;
    auto lin_split = [&](const void* A, long long lda, const Linear& l) -> LnPreAdd {
        GemmArgs a = linear_args(m, A, lda, l, dpart, d, F32, B);
        a.bias = nullptr; a.active = active;
        a.k_splits = 0; a.max_k_splits = MAX_K_SPLITS; a.split_stride = part_stride;
        int chosen = 1;
        a.chosen_splits = &chosen;
        { ProfScope ps(this, PROF_DEC_GEMM, st); gemm(a, st); }
        LnPreAdd pre;
        pre.parts = dpart; pre.n_parts = chosen; pre.part_stride = part_stride; pre.bias = l.b;
        return pre;
    };
    auto ln = [&](const LNorm& n, const LnPreAdd& pre) {
        ProfScope ps(this, PROF_LAYERNORM, st);
        if (pre.n_parts > 0) layernorm_preadd(dx, pre, n.g, n.b, dln, dt, B, d, 1e-5f, active, st);
        else layernorm(dx, n.g, n.b, dln, dt, nullptr, B, d, 1e-5f, active, st);
    };
    decoder_embed(tokens, g.max_tgt, state, m->emb, m->dec_pos, dt, dx, B, d, st, ragged_len());
    LnPreAdd pending;   // residual update still owed to dx (previous layer's fc2)
    for (int l = 0; l < g.dec_layers; ++l) {
        const DecLayer& L = m->dec[l];
        // --- self attention: fused q|k|v projection, in-place paged append, one-query attention
        ln(L.ln1, pending);
        lin(dln, d, L.qkv, dqkv, 3 * d, dt, 0, nullptr);
        {
            DecAttnArgs a;
            a.dtype = dt; a.q = dqkv; a.q_stride = 3 * d; a.out = datt; a.out_stride = d; a.B = B; a.H = g.n_heads;
            a.state = state; a.row_active = unfinished; a.row_len = ragged_len();
            a.k_new = eoff(dqkv, d, dt); a.v_new = eoff(dqkv, 2 * d, dt); a.new_stride = 3 * d;
            a.k_pages = eoff(self_k, (size_t)l * self_layer_elems(), dt);
            a.v_pages = eoff(self_v, (size_t)l * self_layer_elems(), dt);
            a.page_table = page_table; a.pages_per_seq = pages_per_seq; a.page_tokens = PAGE_TOKENS;
            ProfScope ps(this, PROF_SELF_ATTN, st);
            decode_attention(a, st);
        }
        const LnPreAdd after_self = lin_split(datt, d, L.out);
        // --- cross attention over the K/V projected once per utterance
        ln(L.ln2, after_self);
        const LnPreAdd qpre = lin_split(dln, d, L.cq);
        {
            DecAttnArgs a;
            a.dtype = dt; a.out = datt; a.out_stride = d; a.B = B; a.H = g.n_heads;
            a.q_parts = qpre.parts; a.q_n_parts = qpre.n_parts; a.q_part_stride = qpre.part_stride; a.q_bias = qpre.bias;
            a.state = nullptr; a.n_keys = g.n_ctx; a.active = active; a.row_active = unfinished;
            const size_t per_kv = (size_t)max_batch * g.n_heads * g.n_ctx * 64;
            a.k = eoff(cross, (size_t)l * cross_layer_elems(), dt);
            a.v = eoff(cross, (size_t)l * cross_layer_elems() + per_kv, dt);
            a.kv_bstride = (long long)g.n_heads * g.n_ctx * 64; a.kv_hstride = (long long)g.n_ctx * 64;
            ProfScope ps(this, PROF_CROSS_ATTN, st);
            decode_attention(a, st);
        }
        const LnPreAdd after_cross = lin_split(datt, d, L.cout);
        // --- MLP
        ln(L.ln3, after_cross);
        lin(dln, d, L.fc1, dffn, g.ffn, dt, 1, nullptr);
        pending = lin_split(dffn, g.ffn, L.fc2);
    }
    ln(m->dec_ln, pending);
    lm_head(st);
}

// the logits processors + argmax can move into the LM head's epilogue when nobody needs the logits themselves
bool Session::lm_head_fusable() const {
    static const bool off = std::getenv("WB_NO_FUSED_ARGMAX") != nullptr;   // (dev) A/B
    return !off && m->dtype == BF16 && get_gemm_backend() == 0 && logits_dump == nullptr && m->cfg.d_model % 64 == 0 &&
           (size_t)2 * gemm_argmax_partials(m->cfg.vocab) <= (size_t)m->cfg.vocab;   // the partials fit the logits buffer
}

// LM head: proj_out shares storage with embed_tokens (modeling_whisper.py:1335,1433), no bias.  Reads dln [B, d].
void Session::lm_head(cudaStream_t st) {
    const ModelConfig& g = m->cfg;
    GemmArgs a;
    a.A = dln; a.lda = g.d_model; a.W = m->emb; a.ldw = g.d_model; a.in_dtype = m->dtype;
    a.out = logits; a.ldo = g.vocab; a.out_dtype = F32; a.M = batch; a.N = g.vocab; a.K = g.d_model; a.active = &state->active;
    if (fuse_argmax_now) {
        const int stride = gemm_argmax_partials(g.vocab);
        a.am_val = logits; a.am_idx = reinterpret_cast<int*>(logits + (size_t)max_batch * stride); a.am_stride = stride;
        a.am_mask = m->vocab_mask; a.am_state = state; a.am_row_len = ragged_len(); a.am_begin = g.begin_index;
        a.am_count = &argmax_parts;
    }
    ProfScope ps(this, PROF_LM_HEAD, st);
    gemm(a, st);
}

static std::atomic<bool>& graphs_enabled() {
    static std::atomic<bool> on{true};
    return on;
}
void set_cuda_graphs(bool on) { graphs_enabled() = on; }

bool Session::graph_ok() const {
    const bool timing_this_step = prof_class != 0 && (prof_step < 0 || steps_enqueued == prof_step);
    return opt_or(opt_cuda_graphs, graphs_enabled()) && !graph_capture_failed && step_warm && forced_tokens == nullptr &&
           logits_dump == nullptr && !timing_this_step;
}

void Session::build_step_graph(cudaStream_t st) {
    if (step_graph) { cudaGraphExecDestroy(step_graph); step_graph = nullptr; }
    const long long c0 = launch_counter().load();
    const int enq0 = steps_enqueued;
    WB_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    try {
        decode_step(st);
    } catch (...) {
        cudaStreamEndCapture(st, &graph);
        if (graph) cudaGraphDestroy(graph);
        steps_enqueued = enq0;
        launch_counter().store(c0);
        throw;
    }
    WB_CHECK_CUDA(cudaStreamEndCapture(st, &graph));
    step_graph_launches = launch_counter().load() - c0;
    launch_counter().store(c0);      // nothing ran yet: replays are counted when they are launched
    steps_enqueued = enq0;
    const cudaError_t e = cudaGraphInstantiate(&step_graph, graph, 0);
    cudaGraphDestroy(graph);
    WB_CHECK_CUDA(e);
    step_graph_batch = batch;
    step_graph_generation = m->generation_version;
    step_graph_mode = step_mode();
}

// one decode step on the session's loop stream: CUDA-graph replay when possible, eager launches otherwise
// Host-side work of a step that must not end up inside a graph capture: the fused-chain phase table of the current batch,
// and the fed token's embedding in dx when a multi-kernel step (which embeds for itself and then overwrites dx with its
// residual stream) has run since it was last written there.
void Session::prepare_step(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    WB_CHECK_CUDA(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone) return;    // enqueue_step prepared the step before it began the capture
    const int mode = step_mode() & 15;
    if (mode == 2 && chain_batch != batch) {
        WB_CHECK_CUDA(cudaStreamSynchronize(st));     // queued steps may still read the table
        build_chain_table();
    }
    if (mode != 0 && !dx_embedded) {
        decoder_embed(tokens, m->cfg.max_tgt, state, m->emb, m->dec_pos, m->dtype, dx, batch, m->cfg.d_model, st, ragged_len());
        dx_embedded = true;
    }
}

void Session::enqueue_step() {
    cudaStream_t st = loop_stream;
    prepare_step(st);
    const bool fuse = lm_head_fusable() && (step_mode() & 15) != 1;
    if (fuse != fuse_argmax_now) { fuse_argmax_now = fuse; step_graph_batch = -1; }   // the captured step is of the other kind
    struct InLoop { bool& f; InLoop(bool& x) : f(x) { f = true; } ~InLoop() { f = false; } } in_loop(in_loop_step);
    if (graph_ok()) {
        if (step_graph == nullptr || step_graph_batch != batch || step_graph_generation != m->generation_version ||
            step_graph_mode != step_mode()) {
            try {
                build_step_graph(st);
            } catch (const Error& e) {
                graph_capture_failed = true;   // capture not possible for THIS session: it stays on eager launches
                cudaGetLastError();         // do not leave the (non-sticky) capture error for the next caller to trip over
                static bool warned = false;
                if (!warned) {              // a silent perf fallback is a bug report waiting to happen: say it once
                    warned = true;
                    std::fprintf(stderr, "[whisper_b200] CUDA graph capture of the decode step failed (%s): decode steps are "
                                         "launched kernel by kernel from now on (slower, same results)\n", e.what());
                }
            }
        }
    }
    if (graph_ok() && step_graph != nullptr && step_graph_batch == batch && step_graph_generation == m->generation_version &&
        step_graph_mode == step_mode()) {
        WB_CHECK_CUDA(cudaGraphLaunch(step_graph, st));
        launch_counter().fetch_add(step_graph_launches, std::memory_order_relaxed);
        dx_embedded = (step_graph_mode & 15) != 0;
        ++steps_enqueued;
    } else {
        decode_step(st);
        step_warm = true;
    }
}

// Greedy loops of n independent sessions (sub-batches of one model) interleaved step by step, each on its own stream:
// the HBM-bound cross-attention of one sub-batch overlaps the latency-bound GEMM / LayerNorm kernels of the others.
// n == 1 is the plain loop.  Loops run on the sessions' own streams, fenced against the caller's stream on both sides
// (the caller may hand us the legacy default stream, on which a CUDA graph can neither be captured nor replayed).
void decode_run_multi(Session** ss, int n, int max_steps, int check_every, int* final_lens, cudaStream_t caller) {
    WB_REQUIRE(ss != nullptr && n >= 1 && n <= 8, "1..8 sessions");
    for (int k = 0; k < n; ++k) WB_REQUIRE(ss[k] != nullptr && ss[k]->batch > 0, "decode_begin was not called on every session");
    const ModelConfig& g = ss[0]->m->cfg;
    if (max_steps <= 0 || max_steps > g.max_length - 1) max_steps = g.max_length - 1;
    if (check_every <= 0) check_every = 32;
    bool pending[8] = {}, stopped[8] = {};
    for (int k = 0; k < n; ++k) {
        ss[k]->exclusive = n == 1;
        WB_CHECK_CUDA(cudaEventRecord(ss[k]->fence_event, caller));
        WB_CHECK_CUDA(cudaStreamWaitEvent(ss[k]->loop_stream, ss[k]->fence_event, 0));
    }
    for (int i = 0; i < max_steps; ++i) {
        bool any = false;
        for (int k = 0; k < n; ++k) {
            if (stopped[k]) continue;
            any = true;
            ss[k]->enqueue_step();
        }
        if (!any) break;
        if ((i + 1) % check_every == 0 && i + 1 < max_steps) {
            // look at the PREVIOUS snapshot (long finished while a window of steps is still queued), then take a new one
            for (int k = 0; k < n; ++k) {
                if (stopped[k]) continue;
                Session* s = ss[k];
                if (pending[k]) {
                    WB_CHECK_CUDA(cudaEventSynchronize(s->check_event));
                    if (s->host_state->active == 0) { stopped[k] = true; continue; }
                }
                WB_CHECK_CUDA(cudaMemcpyAsync(s->host_state, s->state, sizeof(StepState), cudaMemcpyDeviceToHost, s->loop_stream));
                WB_CHECK_CUDA(cudaEventRecord(s->check_event, s->loop_stream));
                pending[k] = true;
            }
        }
    }
    for (int k = 0; k < n; ++k) {
        Session* s = ss[k];
        WB_CHECK_CUDA(cudaMemcpyAsync(s->host_state, s->state, sizeof(StepState), cudaMemcpyDeviceToHost, s->loop_stream));
        WB_CHECK_CUDA(cudaEventRecord(s->fence_event, s->loop_stream));
        WB_CHECK_CUDA(cudaStreamWaitEvent(caller, s->fence_event, 0));
    }
    for (int k = 0; k < n; ++k) {
        Session* s = ss[k];
        WB_CHECK_CUDA(cudaStreamSynchronize(s->loop_stream));
        if (final_lens) final_lens[k] = s->host_state->active ? s->host_state->cur_len : s->host_state->final_len;
        if (s->compacted) {       // ids of the rows still in the batch -> result buffer (original row order)
            s->publish_results(s->loop_stream);
            WB_CHECK_CUDA(cudaStreamSynchronize(s->loop_stream));
        }
    }
}

int Session::decode_run(int max_steps, int check_every, cudaStream_t caller) {
    Session* self = this;
    int final_len = 0;
    decode_run_multi(&self, 1, max_steps, check_every, &final_len, caller);
    return final_len;
}

}  // namespace wb

// Model / session objects behind the C-ABI (include/whisper_b200.h).  Native runtime: weight packing,
// encoder pass, cross-K/V projection, the on-device greedy loop.
#pragma once
#include <string>
#include <vector>

#include "wb_internal.h"

namespace wb {

struct ModelConfig {
    int d_model = 0, n_heads = 0, enc_layers = 0, dec_layers = 0, ffn = 0, vocab = 0;
    int n_mels = 80, n_frames = 3000, n_ctx = 1500, max_tgt = 448;
    int sot = 50257, eos = 50256, pad = 50256, max_length = 448, begin_index = 2;
};

struct Linear { void* w = nullptr; float* b = nullptr; int n = 0, k = 0; };
struct LNorm { float* g = nullptr; float* b = nullptr; };
struct EncLayer { LNorm ln1, ln2; Linear qkv, out, fc1, fc2; };
struct DecLayer { LNorm ln1, ln2, ln3; Linear qkv, out, cq, ckv, cout, fc1, fc2; };

constexpr int CONV1_KPAD = 256;   // 3 * 80 = 240 zero-padded to a multiple of the 64-wide K block
constexpr int H1_ROWS = 3008;     // rows per utterance of the padded conv1 output (1 zero row + 3000 + 7 zero rows)
constexpr int CONV2_MPERIOD = 1504;
constexpr int PAGE_TOKENS = 64;
constexpr int MAX_K_SPLITS = 8;   // split-K slabs of the skinny decode GEMMs (deferred reduction)
// kernel classes for wb_session_profile
enum ProfClass { PROF_NONE = 0, PROF_CROSS_ATTN = 1, PROF_SELF_ATTN = 2, PROF_DEC_GEMM = 3, PROF_LM_HEAD = 4,
                 PROF_ENC_GEMM = 5, PROF_ENC_ATTN = 6, PROF_LAYERNORM = 7, PROF_GREEDY = 8, PROF_STEM = 9, PROF_CROSS_KV = 10 };

// packed conv-stem weights: w1 [d, CONV1_KPAD] (k = tap * n_mels + c), w2 [d, 3d] (k = tap * d + c), compute dtype
struct StemWeights {
    const void* w1 = nullptr; const float* b1 = nullptr; const void* w2 = nullptr; const float* b2 = nullptr;
    const float* pos = nullptr;  // [n_frames / 2, d] fp32
    int dtype = F32, d = 0, n_mels = 80, n_frames = 3000;
};
// a1: im2col buffer, h1p: zero-initialised padded conv1 output (rows 0 and 3001.. of every utterance stay zero)
void conv_stem(const StemWeights& w, const float* mel, int bc, void* a1, void* h1p, float* x, cudaStream_t st);
size_t conv_stem_a1_bytes(int bc, int n_frames, int dtype);
size_t conv_stem_h1p_bytes(int bc, int d, int dtype);

struct Model {
    ModelConfig cfg;
    int dtype = F32;
    Linear conv1, conv2;
    float* enc_pos = nullptr;
    std::vector<EncLayer> enc;
    LNorm enc_ln;
    void* emb = nullptr;
    void* dec_pos = nullptr;
    std::vector<DecLayer> dec;
    LNorm dec_ln;
    unsigned char* vocab_mask = nullptr;
    int* force_map = nullptr;
    std::vector<void*> allocs;
    size_t weight_bytes = 0;
    long long loaded = 0;
    int generation_version = 0;   // bumped by set_generation: captured decode-step graphs hold begin_index by value

    Model(const ModelConfig& c, int dt);
    ~Model();
    void* dalloc(size_t bytes);
    // HF state_dict key -> packed device weights (fp32 host data)
    void load_tensor(const std::string& name, const float* host, long long numel);
    void set_generation(const int* suppress, int n_suppress, const int* begin_suppress, int n_begin, int begin_index,
                        const int* forced_pairs, int n_forced);
    void check_complete() const;
    int expected_tensors() const;
};

// device buffers carved out of the caller-provided workspace
struct Buffers {
    // encoder workspace (sized for enc_chunk utterances)
    void* a1 = nullptr; void* h1p = nullptr; float* x = nullptr; void* ln = nullptr; void* qkv = nullptr;
    void* att = nullptr; void* ffn = nullptr;
    void* enc = nullptr;        // [max_batch * n_ctx, d] encoder output in the compute dtype
    // caches
    void* cross = nullptr;      // [L][2][max_batch][H][n_ctx][64]
    void* self_k = nullptr;     // [L][num_pages][H][PAGE_TOKENS][64]
    void* self_v = nullptr;
    int* page_table = nullptr;  // [max_batch][pages_per_seq]
    // decode activations
    float* dx = nullptr; float* dpart = nullptr; void* dln = nullptr; void* dqkv = nullptr; void* datt = nullptr; void* dq = nullptr;
    void* dffn = nullptr; float* logits = nullptr;
    int* tokens = nullptr; int* unfinished = nullptr; StepState* state = nullptr;
    // whole-step kernel (step_mega.cu): attention partials of split items, grid-barrier / per-item arrival counters
    float* mega_part = nullptr; unsigned* mega_sync = nullptr; void* mega_table = nullptr;
    int* result_tokens = nullptr;   // [max_batch, max_tgt] ids in ORIGINAL row order once rows have been compacted away
    int* row_len = nullptr;         // [max_batch] per-row lengths of a RAGGED batch (in-flight refill); unused while every row is at cur_len
    // fused decode chains (step_chain.cu): the grid-barrier words
    unsigned* chain_sync = nullptr;
};
size_t mega_part_bytes(int max_batch, int heads);
size_t mega_sync_bytes(int max_batch, int heads);
size_t mega_table_bytes(int dec_layers);
size_t chain_table_bytes(int dec_layers);
size_t chain_sync_bytes();

struct Session : Buffers {
    Model* m;
    int max_batch, enc_chunk;
    int pages_per_seq = 0, num_pages = 0;
    const int* forced_tokens = nullptr;  // teacher forcing (tests): [B, max_length]
    float* logits_dump = nullptr; int logits_dump_steps = 0;
    int batch = 0;
    int steps_enqueued = 0;
    StepState* host_state = nullptr;  // pinned
    // the decode step captured as a CUDA graph (all kernel arguments are step-invariant: lengths / stop flag live on the
    // device), replayed by decode_run; eager launches are kept for teacher forcing, logits dumps and live profiling
    cudaStream_t loop_stream = nullptr;   // decode_run's own stream (graph capture is impossible on the legacy default stream)
    cudaEvent_t fence_event = nullptr;
    cudaGraphExec_t step_graph = nullptr;
    int step_graph_batch = 0;
    int step_graph_generation = -1;
    int step_graph_mode = 0;          // path of the captured step: 0 multi-kernel, 1 whole-step kernel (B <= 16), 2 fused chains
    int step_mode() const { return (use_mega() ? 1 : use_chain() ? 2 : 0) | (ragged ? 16 : 0); }
    bool exclusive = true;            // this session's loop is the only one running on the device (decode_run_multi, n == 1)
    long long step_graph_launches = 0;
    bool step_warm = false;           // one eager step has run (one-time kernel attribute setup done)
    // per-session overrides of the process-wide A/B switches (wb_session_set_option): -1 inherit, 0 off, 1 on
    int opt_small_batch_path = -1, opt_decode_chain_path = -1, opt_cuda_graphs = -1, opt_merge_attention = -1;
    bool graph_capture_failed = false;   // stream capture of the step failed once on this session: eager launches from then on
    void set_option(const std::string& name, int value);
    // LM head fused with the logits processors + argmax (bf16 tcgen05 path, steps of the greedy loop that need no logits):
    // the [B, V] fp32 logits are neither written nor re-read; partial (max, index) per column tile live in the logits buffer
    bool fuse_argmax_now = false;            // set by enqueue_step for the step being enqueued / captured
    bool in_loop_step = false;               // decode_step is running on behalf of the greedy loop (enqueue_step)
    int argmax_parts = 0;                    // partial entries per row the last fused LM head wrote
    bool lm_head_fusable() const;
    void lm_head(cudaStream_t s);            // LM head GEMM of the current step (fused or materialising the logits)
    bool dx_embedded = false;         // dx holds E[last token] + P[position] of every row (what the whole-step kernel starts from)
    bool graph_ok() const;
    void build_step_graph(cudaStream_t s);
    cudaEvent_t check_event = nullptr;
    bool h1p_zeroed = false;
    // live per-kernel-class timing (bench roofline): CUDA events recorded on the launching stream around every
    // launch of the selected class inside the real loop
    int prof_class = 0;
    int prof_step = -1;   // >= 0: only the decode step with this index is timed (and runs eagerly); the others replay the graph
    std::vector<cudaEvent_t> prof_events;  // pairs
    size_t prof_used = 0;
    void prof_begin(int cls, cudaStream_t s);
    void prof_end(int cls, cudaStream_t s);
    void prof_read(double* total_ms, long long* launches);  // synchronises, then resets

    Session(Model* model, int max_batch, int enc_chunk, void* workspace, size_t workspace_bytes);
    ~Session();
    static size_t workspace_bytes(const Model* m, int max_batch, int enc_chunk);

    void stem_chunk(const float* mel, int bc, cudaStream_t s);
    void stem(const float* mel, int B, float* x_out, cudaStream_t s);
    void encode(const float* mel, int B, float* enc_out_f32, cudaStream_t s);   // also projects cross K/V
    void encode_at(const float* mel, int n, int slot0, float* enc_out_f32, cudaStream_t s);   // into batch slots [slot0, slot0 + n)
    void set_encoder_output(const void* enc_states, int dtype, int B, cudaStream_t s);  // external encoder states
    void decode_begin(int B, cudaStream_t s);
    void decode_step(cudaStream_t s);
    void decode_step_large(cudaStream_t s);   // tcgen05 / CUDA-core GEMMs, split-K with deferred reduction (any batch), 268 launches
    void decode_step_mega(cudaStream_t s);    // B <= 16, bf16: ONE persistent cooperative kernel per token (step_mega.cu)
    bool mega_supported() const;
    bool use_mega() const;                    // this batch goes through the whole-step kernel
    void build_mega_table();                  // phase descriptors of the whole-step kernel (constructor)
    int mega_grid = 0;                        // CTAs of the whole-step kernel = SMs of the device the table was built for
    // B > 16, bf16: the GEMM / LayerNorm chains between the attention kernels as persistent cooperative tcgen05 kernels
    // (step_chain.cu): 4 launches per layer instead of 11
    void decode_step_chain(cudaStream_t s);
    bool chain_supported() const;
    bool use_chain() const;
    void init_chain();                        // constructor: kernel attributes, grid size
    void build_chain_table();                 // phase table for the current batch
    void launch_chain(int ph_begin, int ph_end, cudaStream_t s, bool merged = false);
    bool chain_merge_ok() const;              // this batch is small enough for the attention kernels to become phases of the layer's launch
    void prepare_step(cudaStream_t s);        // host-side work a step needs OUTSIDE a graph capture (phase table, embedding in dx)
    int chain_grid = 0, chain_batch = -1;
    std::vector<int> chain_q_parts;           // split-K slabs of each layer's cross-attention q projection
    std::vector<unsigned char> chain_merged;  // same with the attention phases in place (small batches: one launch per decoder layer); empty = not used
    std::vector<unsigned char> chain_host;    // phase descriptors of the step (tensor maps included); a launch copies its slice into its kernel parameters
    int decode_run(int max_steps, int check_every, cudaStream_t s);  // returns final length (syncs)
    // Finished-row compaction: utterances that emitted EOS leave the decode batch; the rows still running
    // move to the front (ids, page-table rows, cross K/V rows) and the following steps run on `batch` = their number.
    int decode_compact(cudaStream_t s);       // syncs; returns the rows still running (0: the loop has stopped)
    void publish_results(cudaStream_t s);     // current ids -> result_tokens[row_origin]
    // In-flight refill: between two windows of steps the utterances that have finished leave the batch, the
    // rows still decoding move to the front and NEW utterances are encoded into the freed slots; from then on the rows of the
    // batch are at different positions (`ragged`: per-row lengths in row_len, the kernels take them instead of cur_len).
    // Finished utterances are handed to the caller (ids in the order they were admitted: 0 .. for decode_begin's rows).
    struct Finished { int utt; int len; std::vector<int> ids; };
    int decode_refill(const float* mel_new, int n_new, std::vector<Finished>& finished, int* n_admitted, cudaStream_t s);
    void compact_rows(const std::vector<int>& keep, cudaStream_t s);
    bool ragged = false;
    bool refill_mode = false;                 // wb_decode_refill was used since decode_begin: results travel through it, not result_tokens
    int next_utt = 0;                         // id of the next utterance to be admitted
    const int* ragged_len() const { return ragged ? row_len : nullptr; }
    std::vector<int> row_origin;              // original row (utterance id) of decode row i
    std::vector<int> page_table_host;         // host mirror of the page table (rows are swapped, never duplicated)
    bool compacted = false;
    int begin_batch = 0;
    void enqueue_step();                                             // one step on loop_stream (graph replay or eager)
    size_t cross_layer_elems() const;
    size_t self_layer_elems() const;
};

// CUDA events around the launches of one kernel class (Session::prof_*)
struct ProfScope {
    Session* s; int cls; cudaStream_t st;
    ProfScope(Session* s_, int c, cudaStream_t t) : s(s_), cls(c), st(t) { s->prof_begin(cls, st); }
    ~ProfScope() noexcept(false) { s->prof_end(cls, st); }
};

void decode_run_multi(Session** sessions, int n, int max_steps, int check_every, int* final_lens, cudaStream_t caller);

}  // namespace wb

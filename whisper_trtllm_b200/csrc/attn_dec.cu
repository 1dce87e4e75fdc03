// One-query attention for the greedy decode step: softmax(q K^T) V per (utterance, head), head_dim 64.
//
//  * cross attention: K/V were projected ONCE per utterance into [B, H, 1500, 64] and are only read here
//    (the reference re-copies the whole 1500-frame K/V block three times per step: identity-plugin memcpy
//    models/whisper/model.py:287-288, output concat :461-462, .clone() run.py:145-146).
//  * self attention: the new key/value row is appended IN PLACE into a paged cache (page = 64 tokens x all
//    heads) and the kernel attends over cur_len keys, replacing the oracle's torch.cat growth
//    (modeling_whisper.py:494-495) and TRT-LLM's slice + concat (model.py:276-281).
//  Oracle semantics (modeling_whisper.py:468-526): no mask, q pre-scaled (folded into the q weights), fp32 softmax.
//
// HBM-bound streaming kernel, single pass (flash-decoding style online softmax):
//   - persistent grid: a few CTAs per SM walk the (utterance, head) items round-robin, so the tail of the last
//     wave costs 1/28 of the launch instead of half a wave;
//   - every key row (128 B in bf16) is read by 8 consecutive lanes with one 16-byte load each, so a warp
//     instruction covers 4 complete rows = 512 contiguous bytes; K and V rows of UNROLL iterations are all in
//     flight before the first use (16-byte L1-bypassing loads);
//   - each 8-lane group keeps its own running (max, sum, acc[8]) and the groups/warps are merged once per item.
#include <atomic>
#include <cstdlib>

#include "wb_internal.h"

namespace wb {

namespace {
constexpr int DH = 64;
// the page size is a compile-time power of two: the paged kernels were instruction-issue bound (ncu: 0.59 IPC per scheduler,
// 69 instructions per 512-byte load) with a runtime integer division and modulo in every row address
constexpr int PAGE_TOKENS_C = 64, PAGE_SHIFT = 6;
// cross attention (1500 keys per item): 256 threads; paged self attention (<= 447 keys, latency-bound per item): 128
// threads so that twice as many items are in flight per SM
constexpr int THREADS_CROSS = 256, THREADS_SELF = 128;

template <typename T> __device__ __forceinline__ float softmax_exp(float x);
template <> __device__ __forceinline__ float softmax_exp<float>(float x) { return expf(x); }      // exactness path
template <> __device__ __forceinline__ float softmax_exp<bf16>(float x) { return __expf(x); }     // speed path

template <typename T, bool kPaged, int THREADS, int UNROLL, bool kPipe = false>
__global__ void __launch_bounds__(THREADS) decode_attn_kernel(DecAttnArgs a) {
    constexpr int WARPS = THREADS / 32;
    constexpr int VEC = Vec16<T>::N;       // elements per 16-byte load: 8 (bf16) / 4 (fp32)
    constexpr int LPK = DH / VEC;          // lanes per key row: 8 / 16
    constexpr int KPW = 32 / LPK;          // key rows per warp instruction: 4 / 2
    constexpr int KPB = KPW * WARPS;       // key rows per block iteration
    __shared__ float part_m[2][WARPS], part_l[2][WARPS];
    __shared__ float part_o[2][WARPS][DH];

    pdl_wait();
    pdl_trigger();
    int n = a.n_keys;
    if (a.active != nullptr && *a.active == 0) return;
    if (a.state != nullptr) {
        if (a.state->active == 0) return;
        n = a.state->cur_len;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane % LPK, grp = lane / LPK;
    const int n_items = a.B * a.H;
    int parity = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / a.H, h = item - b * a.H;
        if (a.row_active != nullptr && a.row_active[b] == 0) continue;   // finished utterance: block-uniform skip
        if (a.row_len != nullptr) n = a.row_len[b];                       // ragged batch (in-flight refill): this row's length
        parity ^= 1;
        float qf[VEC];
        if (a.q_parts != nullptr) {   // q = bias + sum of the split-K partial slabs of the q projection (fp32, fixed order)
            const int c0 = h * DH + sub * VEC;
#pragma unroll
            for (int i = 0; i < VEC; i += 4) {
                float4 t = a.q_bias != nullptr ? *reinterpret_cast<const float4*>(a.q_bias + c0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < a.q_n_parts; ++p) {
                    const float4 u = *reinterpret_cast<const float4*>(a.q_parts + (size_t)p * a.q_part_stride + (size_t)b * a.H * DH + c0 + i);
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                qf[i] = t.x; qf[i + 1] = t.y; qf[i + 2] = t.z; qf[i + 3] = t.w;
            }
        } else {
            ld16(reinterpret_cast<const T*>(a.q) + (size_t)b * a.q_stride + h * DH + sub * VEC).unpack(qf);
        }

        const T* kbase = nullptr;
        const T* vbase = nullptr;
        const int* pt = nullptr;
        if constexpr (kPaged) {
            pt = a.page_table + (size_t)b * a.pages_per_seq;
            if (a.k_new != nullptr) {
                if (warp == 0 && grp == 0) {  // in-place append at slot n-1
                    const int s = n - 1;
                    const size_t off = (((size_t)pt[s >> PAGE_SHIFT] * a.H + h) * PAGE_TOKENS_C + (s & (PAGE_TOKENS_C - 1))) * DH + sub * VEC;
                    const size_t src = (size_t)b * a.new_stride + h * DH + sub * VEC;
                    st16(reinterpret_cast<T*>(a.k_pages) + off, ld16(reinterpret_cast<const T*>(a.k_new) + src));
                    st16(reinterpret_cast<T*>(a.v_pages) + off, ld16(reinterpret_cast<const T*>(a.v_new) + src));
                }
                // no barrier: this step reads row n-1 from k_new / v_new (below); the page copy is for the later steps
            }
        } else {
            const size_t o = (size_t)b * a.kv_bstride + (size_t)h * a.kv_hstride;
            kbase = reinterpret_cast<const T*>(a.k) + o;
            vbase = reinterpret_cast<const T*>(a.v) + o;
        }
        auto row_off = [&](int s) -> size_t {
            if constexpr (kPaged)
                return (((size_t)pt[s >> PAGE_SHIFT] * a.H + h) * PAGE_TOKENS_C + (s & (PAGE_TOKENS_C - 1))) * DH + sub * VEC;
            else
                return (size_t)s * DH + sub * VEC;
        };

        float m_run = -INFINITY, l_run = 0.f;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        // one batch = UNROLL rows per lane group: all of its 2 * UNROLL 16-byte requests are issued before the first use
        auto issue = [&](int sb, Vec16<T>* kr, Vec16<T>* vr) {
            const int s0 = sb + grp;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int s = min(s0 + u * KPB, n - 1);
                if constexpr (kPaged) {
                    const T* kp = reinterpret_cast<const T*>(a.k_pages) + row_off(s);
                    const T* vp = reinterpret_cast<const T*>(a.v_pages) + row_off(s);
                    if (s == n - 1 && a.k_new != nullptr) {      // the row appended in this very step
                        const size_t src = (size_t)b * a.new_stride + h * DH + sub * VEC;
                        kp = reinterpret_cast<const T*>(a.k_new) + src;
                        vp = reinterpret_cast<const T*>(a.v_new) + src;
                    }
                    kr[u] = ld16(kp);
                    vr[u] = ld16(vp);
                } else {
                    const size_t off = row_off(s);
                    kr[u] = ld16_stream(kbase + off);
                    vr[u] = ld16_stream(vbase + off);
                }
            }
        };
        auto consume = [&](int sb, const Vec16<T>* kr, const Vec16<T>* vr) {
            const int s0 = sb + grp;
            float sc[UNROLL];
            float mb = -INFINITY;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                float kf[VEC];
                kr[u].unpack(kf);
                float dot = 0.f;
#pragma unroll
                for (int i = 0; i < VEC; ++i) dot = fmaf(qf[i], kf[i], dot);
#pragma unroll
                for (int o = LPK / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                sc[u] = (s0 + u * KPB < n) ? dot : -INFINITY;
                mb = fmaxf(mb, sc[u]);
            }
            const float m_new = fmaxf(m_run, mb);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;   // this group has not seen a valid row yet: every p below is 0
            const float scale = softmax_exp<T>(m_run - m_use);        // exp(-inf) = 0 while l_run / acc are still 0
            l_run *= scale;
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] *= scale;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float p = softmax_exp<T>(sc[u] - m_use);        // 0 for the masked rows
                l_run += p;
                float vf[VEC];
                vr[u].unpack(vf);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
            }
            m_run = m_new;
        };
        // warp-uniform trip counts (the shuffles need all 32 lanes): groups whose rows fall past n are masked
        constexpr int STEP = KPB * UNROLL;
        if constexpr (kPipe) {
            // software pipeline: the requests of batch i+1 are in flight while batch i is reduced (two register buffers)
            Vec16<T> ka[UNROLL], va[UNROLL], kb2[UNROLL], vb2[UNROLL];
            int sb = warp * KPW;
            if (sb < n) issue(sb, ka, va);
            while (sb < n) {
                const int sb1 = sb + STEP;
                if (sb1 < n) issue(sb1, kb2, vb2);
                consume(sb, ka, va);
                if (sb1 >= n) break;
                const int sb2 = sb1 + STEP;
                if (sb2 < n) issue(sb2, ka, va);
                consume(sb1, kb2, vb2);
                sb = sb2;
            }
        } else {
            for (int sb = warp * KPW; sb < n; sb += STEP) {
                Vec16<T> kr[UNROLL], vr[UNROLL];
                issue(sb, kr, vr);
                consume(sb, kr, vr);
            }
        }

        // ---- merge the key groups of the warp (lanes with equal sub), then the warps
#pragma unroll
        for (int o = LPK; o < 32; o <<= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m_run, o);
            const float ol = __shfl_xor_sync(0xffffffffu, l_run, o);
            const float mm = fmaxf(m_run, om);
            const float s1 = (m_run == -INFINITY) ? 0.f : softmax_exp<T>(m_run - mm);
            const float s2 = (om == -INFINITY) ? 0.f : softmax_exp<T>(om - mm);
            l_run = l_run * s1 + ol * s2;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
                acc[i] = acc[i] * s1 + oa * s2;
            }
            m_run = mm;
        }
        if (grp == 0) {
            if (sub == 0) { part_m[parity][warp] = m_run; part_l[parity][warp] = l_run; }
#pragma unroll
            for (int i = 0; i < VEC; ++i) part_o[parity][warp][sub * VEC + i] = acc[i];
        }
        __syncthreads();   // partials are double buffered by item parity: one barrier per item is enough
        if (tid < DH) {
            float mm = part_m[parity][0];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) mm = fmaxf(mm, part_m[parity][w]);
            float l = 0.f, o = 0.f;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const float pm = part_m[parity][w];
                const float s = (pm == -INFINITY) ? 0.f : softmax_exp<T>(pm - mm);
                l = fmaf(part_l[parity][w], s, l);
                o = fmaf(part_o[parity][w][tid], s, o);
            }
            reinterpret_cast<T*>(a.out)[(size_t)b * a.out_stride + h * DH + tid] = from_f32<T>(o / l);
        }
    }
}

// Paged self-attention, ONE WARP PER (utterance, head) item.  The items are short (<= 447 keys = 57 KB at the mean
// length) and the CTA-per-item kernel above spends its time in per-item latency chains (q -> page table -> K/V round
// trips -> smem merge -> barrier): 55 us at length 224 where the bytes need 36 us.  Here every warp owns an item, keeps
// 2 * UNROLL 16-byte loads in flight per lane, merges its four 8-lane groups with shuffles and never touches shared
// memory or a block barrier; all 4096 items of a launch are resident at once.
template <typename T, int UNROLL, bool kForceBatch>
__global__ void __launch_bounds__(128) self_attn_warp_kernel(DecAttnArgs a) {
    constexpr int VEC = Vec16<T>::N, LPK = DH / VEC, KPW = 32 / LPK;
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= a.B * a.H) return;                       // whole warp
    const int b = item / a.H, h = item - b * a.H;
    const int sub = lane % LPK, grp = lane / LPK;
    // everything the item needs before its first K/V request is loaded in ONE round trip (independent loads, then the
    // early-outs): loop state, the row's unfinished flag, its page ids, q and the new k/v row
    const int* pt = a.page_table + (size_t)b * a.pages_per_seq;
    const int active = a.state->active;
    const int n = a.row_len != nullptr ? a.row_len[b] : a.state->cur_len;   // per row when the batch is ragged (in-flight refill)
    const int row_on = a.row_active != nullptr ? a.row_active[b] : 1;
    // the item's page ids live in the warp's registers (lane i holds page i; <= 7 pages for 448 tokens): every row address
    // costs a shuffle instead of a dependent global load in front of each batch of K/V requests
    const int my_page = lane < a.pages_per_seq ? pt[lane] : 0;
    const Vec16<T> q_raw = ld16(reinterpret_cast<const T*>(a.q) + (size_t)b * a.q_stride + h * DH + sub * VEC);
    Vec16<T> k_app, v_app;
    if (a.k_new != nullptr && grp == 0) {
        const size_t src = (size_t)b * a.new_stride + h * DH + sub * VEC;
        k_app = ld16(reinterpret_cast<const T*>(a.k_new) + src);
        v_app = ld16(reinterpret_cast<const T*>(a.v_new) + src);
    }
    if (active == 0 || row_on == 0) return;              // loop stopped / finished utterance (warp-uniform)
    auto row_off = [&](int s) -> size_t {
        const int page = __shfl_sync(0xffffffffu, my_page, s >> PAGE_SHIFT);
        return (((size_t)page * a.H + h) * PAGE_TOKENS_C + (s & (PAGE_TOKENS_C - 1))) * DH + sub * VEC;
    };
    if (a.k_new != nullptr) {
        const size_t off = row_off(n - 1);   // all lanes: row_off shuffles
        if (grp == 0) {   // in-place append at slot n-1
            st16(reinterpret_cast<T*>(a.k_pages) + off, k_app);
            st16(reinterpret_cast<T*>(a.v_pages) + off, v_app);
        }
        __syncwarp();     // the appended row is read below by the other lane groups of this warp
    }
    float qf[VEC];
    q_raw.unpack(qf);
    float m_run = -INFINITY, l_run = 0.f;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    for (int sb = 0; sb < n; sb += KPW * UNROLL) {       // warp-uniform trip count
        const int s0 = sb + grp;
        Vec16<T> kr[UNROLL], vr[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t off = row_off(min(s0 + u * KPW, n - 1));
            kr[u] = ld16_issue(reinterpret_cast<const T*>(a.k_pages) + off);
            vr[u] = ld16_issue(reinterpret_cast<const T*>(a.v_pages) + off);
        }
        if constexpr (kForceBatch) {
            // every load of the batch must have been ISSUED before the first value is consumed (ptxas otherwise sinks the
            // later loads behind the first dot products to save registers, which halves the requests in flight)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                uint32_t* kw = reinterpret_cast<uint32_t*>(&kr[u].raw);
                uint32_t* vw = reinterpret_cast<uint32_t*>(&vr[u].raw);
                asm volatile("" : "+r"(kw[0]), "+r"(kw[1]), "+r"(kw[2]), "+r"(kw[3]), "+r"(vw[0]), "+r"(vw[1]), "+r"(vw[2]), "+r"(vw[3]));
            }
        }
        float sc[UNROLL];
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float kf[VEC];
            kr[u].unpack(kf);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < VEC; ++i) dot = fmaf(qf[i], kf[i], dot);
#pragma unroll
            for (int o = LPK / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            sc[u] = (s0 + u * KPW < n) ? dot : -INFINITY;
            mb = fmaxf(mb, sc[u]);
        }
        const float m_new = fmaxf(m_run, mb);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float scale = softmax_exp<T>(m_run - m_use);
        l_run *= scale;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] *= scale;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float p = softmax_exp<T>(sc[u] - m_use);
            l_run += p;
            float vf[VEC];
            vr[u].unpack(vf);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
        }
        m_run = m_new;
    }
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {                 // merge the key groups of the warp
        const float om = __shfl_xor_sync(0xffffffffu, m_run, o);
        const float ol = __shfl_xor_sync(0xffffffffu, l_run, o);
        const float mm = fmaxf(m_run, om);
        const float s1 = (m_run == -INFINITY) ? 0.f : softmax_exp<T>(m_run - mm);
        const float s2 = (om == -INFINITY) ? 0.f : softmax_exp<T>(om - mm);
        l_run = l_run * s1 + ol * s2;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
            acc[i] = acc[i] * s1 + oa * s2;
        }
        m_run = mm;
    }
    if (grp == 0) {
        const float inv = 1.0f / l_run;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] *= inv;
        Vec16<T> o;
        o.pack(acc);
        st16(reinterpret_cast<T*>(a.out) + (size_t)b * a.out_stride + h * DH + sub * VEC, o);
    }
}

template <typename T, bool kPaged, int THREADS, int UNROLL, bool kPipe = false>
void launch(const DecAttnArgs& a, cudaStream_t stream) {
    static std::atomic<int> per_sm[WB_MAX_DEVICES];     // resident CTAs per SM of this instantiation, per device
    const int dev = current_device(), sms = device_sm_count();
    int blocks_per_sm = per_sm[dev].load(std::memory_order_relaxed);
    if (blocks_per_sm == 0) {
        WB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, decode_attn_kernel<T, kPaged, THREADS, UNROLL, kPipe>, THREADS, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        per_sm[dev].store(blocks_per_sm, std::memory_order_relaxed);
    }
    const int items = a.B * a.H;
    // persistent CTAs walk the items round-robin: size the grid so that every CTA gets the SAME number of items (2048 items on
    // 592 slots = 3.46 rounds would leave 54 % of the slots idle during the 4th; 512 CTAs x 4 items do not)
    const int slots = sms * blocks_per_sm;
    const int rounds = ceil_div(items, slots);
    const int grid = ceil_div(items, rounds);
    launch_kernel(decode_attn_kernel<T, kPaged, THREADS, UNROLL, kPipe>, dim3(grid), dim3(THREADS), 0, stream, true, a);
}
}  // namespace

static int num_sms() { return device_sm_count(); }

bool decode_attention_bulk_supported(const DecAttnArgs& a);                 // attn_dec_bulk.cu
void decode_attention_bulk(const DecAttnArgs& a, cudaStream_t stream);
// 0 = 16-byte load kernel (256 threads, 4 x 2 loads in flight per lane), 1 = cp.async.bulk ring kernel for cross attention,
// 2..9 = tuning variants of the load kernel for bf16 cross attention (threads, unroll); default (128, 8)
static std::atomic<int> g_dec_attn_backend{0};
// paged self-attention: one CTA per item (default) vs one warp per item.  Measured on B200 (B = 256, medium.en, whole
// 447-step loop): CTA-per-item 4.231 s, warp-per-item 4.263 s per 256 utterances.
// paged self-attention variants, us per launch averaged over the 447 steps (B = 256, medium.en, one B200 run):
//   0 warp per item, 4 deep 48.5 (default) | 1 CTA (128 threads, 4 deep) 55.7 | 2 CTA (128, 8) 59.3 | 3 CTA (64, 8) 53.4 |
//   4 CTA (256, 4) 70.2 | 5 warp, 8 deep, loads forced into one batch 53.1 | 6 warp, 8 deep 52.9
static std::atomic<int> g_self_attn_variant{0};
void set_self_attention_variant(int v) { g_self_attn_variant = v; }
void set_decode_attention_backend(int b) { g_dec_attn_backend = b; }

void decode_attention(const DecAttnArgs& a, cudaStream_t stream) {
    const int g_dec_attn_backend = wb::g_dec_attn_backend.load(std::memory_order_relaxed);     // one consistent read per call
    const int g_self_attn_variant = wb::g_self_attn_variant.load(std::memory_order_relaxed);
    if (g_dec_attn_backend == 1 && decode_attention_bulk_supported(a)) {
        decode_attention_bulk(a, stream);
        return;
    }
    WB_REQUIRE((a.q || a.q_parts) && a.out && a.B > 0 && a.H > 0, "bad decode attention arguments");
    const bool paged = a.k_pages != nullptr;
    WB_REQUIRE(paged || (a.k && a.v), "missing K/V");
    WB_REQUIRE(!paged || (a.page_table && a.v_pages && a.pages_per_seq > 0 && a.pages_per_seq <= 32 && a.page_tokens == PAGE_TOKENS_C),
               "bad paged cache (pages hold 64 tokens, at most 32 pages per sequence)");
    WB_REQUIRE(paged ? a.state != nullptr : a.n_keys > 0, "key count must be positive");
    // few items (small batch): one warp per item would leave most SMs idle and serialise a whole sequence behind one warp's
    // round trips -> give every item a 256-thread CTA instead
    const bool few_items = a.B * a.H <= 2 * num_sms();
    if (paged && few_items && a.dtype == BF16 && g_self_attn_variant == 0) {
        launch<bf16, true, 256, 4>(a, stream);
        return;
    }
    // medium item counts (two .. ~eight items per SM): one warp per item leaves most of every SM's load slots empty and each
    // warp walks its whole sequence alone (23 us per launch at B = 32, medium.en, where the bytes need 4 us) -> a 128-thread
    // CTA per item, the keys split over its 4 warps.  (dev) WB_SELF_CTA_ITEMS overrides the threshold.
    static const int cta_items_max = std::getenv("WB_SELF_CTA_ITEMS") ? std::atoi(std::getenv("WB_SELF_CTA_ITEMS")) : 1024;
    if (paged && a.dtype == BF16 && g_self_attn_variant == 0 && a.B * a.H <= cta_items_max) {
        static const int cta_threads = std::getenv("WB_SELF_CTA_THREADS") ? std::atoi(std::getenv("WB_SELF_CTA_THREADS")) : 128;   // (dev)
        if (cta_threads == 256) launch<bf16, true, 256, 4>(a, stream);
        else launch<bf16, true, THREADS_SELF, 4>(a, stream);
        return;
    }
    const bool warp_variant = g_self_attn_variant == 0 || g_self_attn_variant == 5 || g_self_attn_variant == 6;
    if (paged && warp_variant && a.q != nullptr) {
        const dim3 grid(ceil_div(a.B * a.H, 4)), block(128);
        if (a.dtype == F32) launch_kernel(self_attn_warp_kernel<float, 8, false>, grid, block, 0, stream, true, a);
        else if (g_self_attn_variant == 5) launch_kernel(self_attn_warp_kernel<bf16, 8, true>, grid, block, 0, stream, true, a);
        else if (g_self_attn_variant == 6) launch_kernel(self_attn_warp_kernel<bf16, 8, false>, grid, block, 0, stream, true, a);
        else launch_kernel(self_attn_warp_kernel<bf16, 4, false>, grid, block, 0, stream, true, a);
        return;
    }
    if (a.dtype == F32) {
        if (paged) launch<float, true, THREADS_SELF, 4>(a, stream); else launch<float, false, THREADS_CROSS, 4>(a, stream);
    } else if (paged) {
        switch (g_self_attn_variant) {
            case 2: launch<bf16, true, 128, 8>(a, stream); break;
            case 3: launch<bf16, true, 64, 8>(a, stream); break;
            case 4: launch<bf16, true, 256, 4>(a, stream); break;
            default: launch<bf16, true, THREADS_SELF, 4>(a, stream); break;
        }
    } else {
        // measured on B200 (B = 256, medium.en; us per launch): (256,4) 239.5 | (128,4) 247.1 | (256,8) 235.5 | (128,8) 232.1 |
        // (512,4) 247.8 | (256,2) 265.4  -> default 128 threads x 8-deep batches (16 requests in flight per lane)
        switch (g_dec_attn_backend) {
            case 2: launch<bf16, false, 128, 4>(a, stream); break;
            case 3: launch<bf16, false, 256, 8>(a, stream); break;
            case 4: launch<bf16, false, 256, 4>(a, stream); break;
            case 5: launch<bf16, false, 512, 4>(a, stream); break;
            case 6: launch<bf16, false, 256, 2>(a, stream); break;
            case 7: launch<bf16, false, 128, 12>(a, stream); break;
            case 8: launch<bf16, false, 64, 8>(a, stream); break;
            case 9: launch<bf16, false, 192, 8>(a, stream); break;
            case 10: launch<bf16, false, 128, 4, true>(a, stream); break;    // software-pipelined variants
            case 11: launch<bf16, false, 128, 8, true>(a, stream); break;
            case 12: launch<bf16, false, 256, 4, true>(a, stream); break;
            default:
                // few items (small batch): 512 threads per item so that one CTA keeps 128 KB of requests in flight
                if (a.B * a.H <= num_sms()) launch<bf16, false, 512, 8>(a, stream);
                else if (few_items) launch<bf16, false, 256, 8>(a, stream);
                else launch<bf16, false, 128, 8>(a, stream);
                break;
        }
    }
}

}  // namespace wb

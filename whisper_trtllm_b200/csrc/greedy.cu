// On-device greedy bookkeeping: decoder embedding, logits processors + argmax + EOS/pad handling + length
// advance.  With these the token loop never returns to the host (the reference syncs twice per step and
// runs three python logits processors, run.py:138-140,199-225).
//
// Oracle semantics reproduced (generation/utils.py:1485-1527, logits_process.py:1281-1328):
//   scores[:, suppress] = -inf                      every step
//   scores[:, begin_suppress] = -inf                when len(ids) == begin_index
//   forced token                                    when len(ids) in force map (all -inf, token <- 0)
//   tok = argmax (first maximal index); tok = tok*unfinished + pad*(1-unfinished); append;
//   unfinished &= tok != eos; stop when no row is unfinished or len(ids) >= max_length.
#include "wb_internal.h"

namespace wb {

namespace {

template <typename T>
__global__ void __launch_bounds__(128) decoder_embed_kernel(const int* __restrict__ tokens, int tokens_stride,
                                                            const StepState* __restrict__ state, const T* __restrict__ emb,
                                                            const T* __restrict__ pos, float* __restrict__ x, int d,
                                                            const int* __restrict__ row_len) {
    pdl_wait();
    pdl_trigger();
    if (state->active == 0) return;
    constexpr int VEC = Vec16<T>::N;
    const int b = blockIdx.x;
    // position == past length (modeling_whisper.py:307-308, model.py:424); per row when the batch is ragged (in-flight refill)
    const int p = (row_len != nullptr ? row_len[b] : state->cur_len) - 1;
    const int tok = tokens[(size_t)b * tokens_stride + p];
    const T* e = emb + (size_t)tok * d;
    const T* pp = pos + (size_t)p * d;
    for (int i = threadIdx.x * VEC; i < d; i += blockDim.x * VEC) {
        float ef[VEC], pf[VEC];
        ld16(e + i).unpack(ef);
        ld16(pp + i).unpack(pf);
#pragma unroll
        for (int j = 0; j < VEC; j += 4)
            *reinterpret_cast<float4*>(x + (size_t)b * d + i + j) =
                make_float4(ef[j] + pf[j], ef[j + 1] + pf[j + 1], ef[j + 2] + pf[j + 2], ef[j + 3] + pf[j + 3]);
    }
}

// stateless embedding (module-level drop-in): x[b*T + t, :] = E[ids[b, t], :] + P[pos0 + t, :]
template <typename T>
__global__ void __launch_bounds__(128) embed_kernel(const int* __restrict__ ids, long long ids_stride, int Tn, int pos0,
                                                    const T* __restrict__ emb, const T* __restrict__ pos, float* __restrict__ x,
                                                    int d, int vocab) {
    constexpr int VEC = Vec16<T>::N;
    const int row = blockIdx.x;
    const int b = row / Tn, t = row - b * Tn;
    int tok = ids[(size_t)b * ids_stride + t];
    tok = min(max(tok, 0), vocab - 1);
    const T* e = emb + (size_t)tok * d;
    const T* pp = pos + (size_t)(pos0 + t) * d;
    for (int i = threadIdx.x * VEC; i < d; i += blockDim.x * VEC) {
        float ef[VEC], pf[VEC];
        ld16(e + i).unpack(ef);
        ld16(pp + i).unpack(pf);
#pragma unroll
        for (int j = 0; j < VEC; j += 4)
            *reinterpret_cast<float4*>(x + (size_t)row * d + i + j) =
                make_float4(ef[j] + pf[j], ef[j + 1] + pf[j + 1], ef[j + 2] + pf[j + 2], ef[j + 3] + pf[j + 3]);
    }
}

// dense KV "concat" of the reference contract (model.py:276-281): out[b,h,0:n) = past[b,h,0:n), out[b,h,n] = new[b,h]
template <typename T>
__global__ void __launch_bounds__(256) kv_append_kernel(const T* __restrict__ past, long long past_bs, long long past_hs,
                                                        const T* __restrict__ cur, long long cur_bs, T* __restrict__ out,
                                                        int H, int n) {
    constexpr int VEC = Vec16<T>::N;
    constexpr int CPR = 64 / VEC;  // 16-byte chunks per row
    const int h = blockIdx.x, b = blockIdx.y;
    T* o = out + ((size_t)b * H + h) * (size_t)(n + 1) * 64;
    const T* p = past + (size_t)b * past_bs + (size_t)h * past_hs;
    const int total = n * CPR;
    for (int i = threadIdx.x; i < total; i += blockDim.x) st16(o + (size_t)i * VEC, ld16(p + (size_t)i * VEC));
    if (threadIdx.x < CPR)
        st16(o + (size_t)n * 64 + threadIdx.x * VEC, ld16(cur + (size_t)b * cur_bs + h * 64 + threadIdx.x * VEC));
}

__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// one block per row; returns the masked argmax in every thread of warp 0 (valid in thread 0)
__device__ int block_masked_argmax(const float* __restrict__ row, int V, const unsigned char* __restrict__ mask,
                                   int mask_bits) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -INFINITY;
    int besti = 0x7fffffff;
    const int nv = V >> 2;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const float4 f = reinterpret_cast<const float4*>(row)[i];
        float vals[4] = {f.x, f.y, f.z, f.w};
        if (mask != nullptr) {
            const uchar4 m = reinterpret_cast<const uchar4*>(mask)[i];
            if (m.x & mask_bits) vals[0] = -INFINITY;
            if (m.y & mask_bits) vals[1] = -INFINITY;
            if (m.z & mask_bits) vals[2] = -INFINITY;
            if (m.w & mask_bits) vals[3] = -INFINITY;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) argmax_combine(best, besti, vals[j], 4 * i + j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        argmax_combine(best, besti, ov, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = best; si[warp] = besti; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? sv[lane] : -INFINITY;
        besti = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            argmax_combine(best, besti, ov, oi);
        }
    }
    return besti;
}

// block-wide (value, index) reduction, first maximum wins; result valid in thread 0
__device__ int block_argmax_reduce(float best, int besti) {
    __shared__ float rv[32];
    __shared__ int ri[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        argmax_combine(best, besti, ov, oi);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { rv[warp] = best; ri[warp] = besti; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? rv[lane] : -INFINITY;
        besti = lane < nw ? ri[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            argmax_combine(best, besti, ov, oi);
        }
    }
    return besti;
}

__global__ void __launch_bounds__(1024) greedy_step_kernel(GreedyArgs a) {
    pdl_wait();
    pdl_trigger();
    StepState* st = a.state;
    if (st->active == 0) return;
    const int b = blockIdx.x;
    const bool ragged = a.row_len != nullptr;
    const int n = ragged ? a.row_len[b] : st->cur_len;  // == input_ids.shape[-1] seen by the processors (of this row)
    const bool row_on = !ragged || a.unfinished[b] != 0;    // ragged batch: a finished row is frozen (block-uniform)
    int tok = a.pad_id;
    if (row_on) {
        const int forced = a.force_map != nullptr ? a.force_map[n] : -1;
        if (forced >= 0) {
            tok = forced;
        } else if (a.part_val != nullptr) {
            // the LM head's epilogue already applied the processors and reduced every column tile (GemmArgs::am_*)
            float best = -INFINITY;
            int besti = 0x7fffffff;
            for (int i = threadIdx.x; i < a.n_parts; i += blockDim.x)
                argmax_combine(best, besti, a.part_val[(size_t)b * a.part_stride + i], a.part_idx[(size_t)b * a.part_stride + i]);
            tok = block_argmax_reduce(best, besti);
        } else {
            const int bits = 1 | (n == a.begin_index ? 2 : 0);
            tok = block_masked_argmax(a.logits + (size_t)b * a.ld, a.V, a.vocab_mask, bits);
        }
    }
    __shared__ int s_tok;
    if (threadIdx.x == 0 && row_on) {
        const int unf = a.unfinished[b];
        tok = unf ? tok : a.pad_id;
        if (a.forced_tokens != nullptr) tok = a.forced_tokens[(size_t)b * a.tokens_stride + n];
        a.tokens[(size_t)b * a.tokens_stride + n] = tok;
        s_tok = tok;
        if (tok == a.eos_id) a.unfinished[b] = 0;
        if (ragged) {
            a.row_len[b] = n + 1;
            if (n + 1 >= a.max_length) a.unfinished[b] = 0;    // this row is full (stopping_criteria.py:61-70, per row)
        }
    }
    if (a.embed_x != nullptr && n < a.tokens_stride && row_on) {
        // embedding of the token just chosen = the input of the NEXT step (position n): x[b, :] = E[tok, :] + P[n, :], so that the
        // whole-step decoder kernel (step_mega.cu) starts from the residual stream (model.py:423-425)
        __syncthreads();
        const int t = s_tok;
        const bf16* e = reinterpret_cast<const bf16*>(a.embed_table) + (size_t)t * a.embed_d;
        const bf16* pp = reinterpret_cast<const bf16*>(a.embed_pos) + (size_t)n * a.embed_d;
        for (int i = threadIdx.x * 8; i < a.embed_d; i += blockDim.x * 8) {
            float ef[8], pf[8];
            ld16(e + i).unpack(ef);
            ld16(pp + i).unpack(pf);
#pragma unroll
            for (int j = 0; j < 8; j += 4)
                *reinterpret_cast<float4*>(a.embed_x + (size_t)b * a.embed_d + i + j) =
                    make_float4(ef[j] + pf[j], ef[j + 1] + pf[j + 1], ef[j + 2] + pf[j + 2], ef[j + 3] + pf[j + 3]);
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const int prev = atomicAdd(&st->done_counter, 1);
        if (prev == a.B - 1) {  // last row of this step: advance the shared length / stop flag
            __threadfence();
            int cnt = 0;
            for (int i = 0; i < a.B; ++i) cnt += (*((volatile int*)&a.unfinished[i]) != 0);
            const int new_len = (ragged ? st->cur_len : n) + 1;   // ragged: a step counter; the rows carry their own lengths
            st->n_unfinished = cnt;
            st->done_counter = 0;
            if (cnt == 0 || (!ragged && new_len >= a.max_length)) {
                st->final_len = new_len;
                st->active = 0;
            }
            st->cur_len = new_len;
        }
    }
}

__global__ void greedy_init_kernel(int* tokens, int tokens_stride, int* unfinished, StepState* state, int B,
                                   int start_token, int pad_id, int max_length) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * max_length) {
        const int b = i / max_length, t = i - b * max_length;
        tokens[(size_t)b * tokens_stride + t] = (t == 0) ? start_token : pad_id;
    }
    if (i < B) unfinished[i] = 1;
    if (i == 0) {
        state->cur_len = 1;
        state->active = 1;
        state->final_len = 0;
        state->done_counter = 0;
        state->n_unfinished = B;
    }
}

__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, long long ld, int V,
                                                           const unsigned char* __restrict__ mask, int mask_bits,
                                                           int* __restrict__ out) {
    const int tok = block_masked_argmax(logits + (size_t)blockIdx.x * ld, V, mask, mask_bits);
    if (threadIdx.x == 0) out[blockIdx.x] = tok;
}

}  // namespace

void decoder_embed(const int* tokens, int tokens_stride, const StepState* state, const void* emb, const void* pos,
                   int dtype, float* x, int B, int d, cudaStream_t stream, const int* row_len) {
    WB_REQUIRE(d % 8 == 0, "d_model must be a multiple of 8");
    if (dtype == F32)
        launch_kernel(decoder_embed_kernel<float>, dim3(B), dim3(128), 0, stream, true, tokens, tokens_stride, state, (const float*)emb, (const float*)pos, x, d, row_len);
    else
        launch_kernel(decoder_embed_kernel<bf16>, dim3(B), dim3(128), 0, stream, true, tokens, tokens_stride, state, (const bf16*)emb, (const bf16*)pos, x, d, row_len);
}

void embed_tokens(const int* ids, long long ids_stride, int B, int T, int pos0, const void* emb, const void* pos, int dtype,
                  float* x, int d, int vocab, cudaStream_t stream) {
    WB_REQUIRE(d % 8 == 0 && B > 0 && T > 0 && pos0 >= 0, "bad embed arguments");
    if (dtype == F32)
        embed_kernel<float><<<B * T, 128, 0, stream>>>(ids, ids_stride, T, pos0, (const float*)emb, (const float*)pos, x, d, vocab);
    else
        embed_kernel<bf16><<<B * T, 128, 0, stream>>>(ids, ids_stride, T, pos0, (const bf16*)emb, (const bf16*)pos, x, d, vocab);
    WB_CHECK_LAUNCH();
}

void kv_append(const void* past, long long past_bs, long long past_hs, const void* cur, long long cur_bs, void* out, int dtype,
               int B, int H, int n, cudaStream_t stream) {
    WB_REQUIRE(cur && out && B > 0 && H > 0 && n >= 0 && (n == 0 || past), "bad kv_append arguments");
    WB_REQUIRE(past_bs % 8 == 0 && past_hs % 8 == 0 && cur_bs % 8 == 0, "kv_append strides must be multiples of 8 elements");
    dim3 grid(H, B);
    if (dtype == F32)
        kv_append_kernel<float><<<grid, 256, 0, stream>>>((const float*)past, past_bs, past_hs, (const float*)cur, cur_bs, (float*)out, H, n);
    else
        kv_append_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)past, past_bs, past_hs, (const bf16*)cur, cur_bs, (bf16*)out, H, n);
    WB_CHECK_LAUNCH();
}

void greedy_step(const GreedyArgs& a, cudaStream_t stream) {
    WB_REQUIRE(a.V % 4 == 0 && a.ld % 4 == 0, "vocab size / logits pitch must be multiples of 4");
    WB_REQUIRE(a.part_val != nullptr || a.logits != nullptr, "greedy step needs logits or the LM head's argmax partials");
    // with partials a row is a few hundred entries: one small block per row
    launch_kernel(greedy_step_kernel, dim3(a.B), dim3(a.part_val != nullptr ? 256 : 1024), 0, stream, true, a);
}

void greedy_init(int* tokens, int tokens_stride, int* unfinished, StepState* state, int B, int start_token, int pad_id,
                 int max_length, cudaStream_t stream) {
    const int n = B * max_length;
    greedy_init_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(tokens, tokens_stride, unfinished, state, B, start_token, pad_id, max_length);
    WB_CHECK_LAUNCH();
}

void argmax_rows(const float* logits, long long ld, int B, int V, const unsigned char* vocab_mask, int mask_bits, int* out,
                 cudaStream_t stream) {
    WB_REQUIRE(V % 4 == 0 && ld % 4 == 0, "vocab size / logits pitch must be multiples of 4");
    argmax_rows_kernel<<<B, 1024, 0, stream>>>(logits, ld, V, vocab_mask, mask_bits, out);
    WB_CHECK_LAUNCH();
}

}  // namespace wb

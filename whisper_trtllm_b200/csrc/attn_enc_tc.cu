// Encoder self-attention on tcgen05 tensor cores: fused flash-style softmax(Q K^T) V over the 1500 frames.
//
// One CTA = one 128-row query tile of one (utterance, head); it walks the keys in 64-wide tiles:
//   S  = Q K_j^T     tcgen05.mma  M=128 N=64 K=64   (Q, K_j K-major in 128B-swizzled smem, S in TMEM)
//   P  = exp2(S - m) online softmax in registers, ONE THREAD PER QUERY ROW (tcgen05.ld 32x32b: lane == row),
//                    written back as bf16 into 128B-swizzled smem (the A operand of the next MMA)
//   O_j = P V_j      tcgen05.mma  M=128 N=64 K=64   (V_j used in place as an MN-major B operand)
//   O  = O * corr + O_j   accumulated in registers by the same thread that owns the row
// Warp roles: warp 0 TMA producer (Q once, K/V ring), warp 1 MMA issuer, warps 2-5 softmax/accumulate.
// S, P and O_j are double buffered so the tensor core works on tile j+1 while the softmax warps handle tile j;
// two CTAs fit per SM (96 KB smem, 256 TMEM columns each).
//
// Semantics: oracle WhisperEncoderAttention.forward (modeling_whisper.py:569-593): no mask, q pre-scaled by
// 0.125 (folded into the packed weights), fp32 softmax statistics.  Replaces the reference's materialised
// [B*H, 1500, 1500] fp32 score tensor (layers/attention.py:332-345).
#include <cstdlib>

#include "wb_internal.h"
#include "wb_ptx.cuh"

namespace wb {

CUtensorMap make_tmap_bf16_2d(const void* ptr, long long ld_elems, int rows, int cols, int box_rows);  // gemm_tc.cu

namespace {

constexpr int TQ = 128, TKV = 64, DH = 64, STAGES = 3;
constexpr uint32_t Q_BYTES = TQ * DH * 2;        // 16 KB
constexpr uint32_t KV_BYTES = TKV * DH * 2;      // 8 KB
constexpr uint32_t P_BYTES = TQ * TKV * 2;       // 16 KB
constexpr uint32_t SMEM_Q = 0;
constexpr uint32_t SMEM_K = SMEM_Q + Q_BYTES;
constexpr uint32_t SMEM_V = SMEM_K + STAGES * KV_BYTES;
constexpr uint32_t SMEM_P = SMEM_V + STAGES * KV_BYTES;
constexpr uint32_t SMEM_BAR = SMEM_P + 2 * P_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 256 + 1024;
constexpr uint32_t TMEM_COLS = 256;              // S0 | S1 | O0 | O1, 64 columns each
constexpr int NUM_THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
enc_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   bf16* __restrict__ out, int S, int H) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;             // STAGES
    uint64_t* v_full = k_full + STAGES;      // STAGES
    uint64_t* kv_empty = v_full + STAGES;    // STAGES
    uint64_t* s_full = kv_empty + STAGES;    // 2
    uint64_t* s_empty = s_full + 2;          // 2
    uint64_t* p_full = s_empty + 2;          // 2
    uint64_t* o_full = p_full + 2;           // 2
    uint64_t* o_empty = o_full + 2;          // 2
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform by construction: uniform role branches
    const int lane = threadIdx.x & 31;
    const uint32_t smem_a = raw_addr + ((1024u - (raw_addr & 1023u)) & 1023u);   // shared-memory address of `smem` (warp-uniform)
    // the barriers by shared-memory address (uniform operands of the converged-warp issue forms below)
    const uint32_t q_full_a = smem_a + SMEM_BAR, k_full_a = q_full_a + 8, v_full_a = k_full_a + STAGES * 8;
    const uint32_t kv_empty_a = v_full_a + STAGES * 8, s_full_a = kv_empty_a + STAGES * 8, s_empty_a = s_full_a + 16;
    const uint32_t p_full_a = s_empty_a + 16, o_full_a = p_full_a + 16, o_empty_a = o_full_a + 16;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * DH;
    const int q0 = qt * TQ;
    const int row_base = b * S;
    const int n_tiles = (S + TKV - 1) / TKV;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmKV);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&kv_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&s_full[s], 1);
            ptx::mbar_init(&s_empty[s], 4);
            ptx::mbar_init(&p_full[s], 4);
            ptx::mbar_init(&o_full[s], 1);
            ptx::mbar_init(&o_empty[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer (converged warp, the elected lane issues) =====================
        ptx::mbar_expect_tx_elect(q_full_a, Q_BYTES);
        ptx::tma_load_2d_elect(smem_a + SMEM_Q, &tmQ, q_full_a, h * DH, row_base + q0);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            ptx::mbar_wait_addr(kv_empty_a + st * 8, ((j / STAGES) & 1) ^ 1);
            ptx::mbar_expect_tx_elect(k_full_a + st * 8, KV_BYTES);
            ptx::tma_load_2d_elect(smem_a + SMEM_K + st * KV_BYTES, &tmKV, k_full_a + st * 8, d + h * DH, row_base + j * TKV);
            ptx::mbar_expect_tx_elect(v_full_a + st * 8, KV_BYTES);
            ptx::tma_load_2d_elect(smem_a + SMEM_V + st * KV_BYTES, &tmKV, v_full_a + st * 8, 2 * d + h * DH, row_base + j * TKV);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, the elected lane issues) =====================
        // tcgen05.mma / commit operands live in uniform registers; issued from an `if (lane == 0)` region every operand was moved
        // across with ELECT + R2UR.BROADCAST (~100 cycles per MMA, 8 MMAs per key tile against 256 cycles of tensor pipe).  The
        // warp stays converged and every operand derives from kernel parameters, block indices and the uniform tile counter.
        constexpr uint32_t idesc_s = ptx::make_idesc_bf16(TQ, TKV, 0, 0);   // S = Q K^T : both operands K-major
        constexpr uint32_t idesc_o = ptx::make_idesc_bf16(TQ, DH, 0, 1);    // O = P V   : V is MN-major (dh contiguous)
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dq = ptx::make_smem_desc_sw128(smem_a + SMEM_Q, 1024, 16);
        auto issue_s = [&](int j) {
            const int st = j % STAGES;
            const uint64_t db = ptx::make_smem_desc_sw128(smem_a + SMEM_K + st * KV_BYTES, 1024, 16);
            const uint32_t dst = tmem_u + (j & 1) * 64;
            ptx::umma_f16_elect(dst, dq, db, idesc_s, 0);
            ptx::umma_f16_elect(dst, dq + 2, db + 2, idesc_s, 1);
            ptx::umma_f16_elect(dst, dq + 4, db + 4, idesc_s, 1);
            ptx::umma_f16_elect(dst, dq + 6, db + 6, idesc_s, 1);
            ptx::umma_commit_elect(s_full_a + (j & 1) * 8);
        };
        ptx::mbar_wait_addr(q_full_a, 0);
        ptx::mbar_wait_addr(k_full_a, 0);
        ptx::tcgen05_fence_after();
        issue_s(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) {
                const int jn = j + 1;
                ptx::mbar_wait_addr(k_full_a + (jn % STAGES) * 8, (jn / STAGES) & 1);
                ptx::mbar_wait_addr(s_empty_a + (jn & 1) * 8, ((jn >> 1) & 1) ^ 1);
                ptx::tcgen05_fence_after();
                issue_s(jn);
            }
            const int st = j % STAGES;
            ptx::mbar_wait_addr(v_full_a + st * 8, (j / STAGES) & 1);
            ptx::mbar_wait_addr(p_full_a + (j & 1) * 8, (j >> 1) & 1);
            ptx::mbar_wait_addr(o_empty_a + (j & 1) * 8, ((j >> 1) & 1) ^ 1);
            ptx::tcgen05_fence_after();
            // A = P [128 x 64 keys] K-major; B = V_j [64 keys x 64 dh], MN-major: key rows are 128 B apart,
            // 8-row swizzle atoms 1024 B apart (SBO); one UMMA_K step = 16 keys = 2048 B
            const uint64_t da = ptx::make_smem_desc_sw128(smem_a + SMEM_P + (j & 1) * P_BYTES, 1024, 16);
            const uint64_t db = ptx::make_smem_desc_sw128(smem_a + SMEM_V + st * KV_BYTES, 1024, KV_BYTES);
            const uint32_t dst = tmem_u + 128 + (j & 1) * 64;
            ptx::umma_f16_elect(dst, da, db, idesc_o, 0);
            ptx::umma_f16_elect(dst, da + 2, db + 128, idesc_o, 1);
            ptx::umma_f16_elect(dst, da + 4, db + 256, idesc_o, 1);
            ptx::umma_f16_elect(dst, da + 6, db + 384, idesc_o, 1);
            ptx::umma_commit_elect(o_full_a + (j & 1) * 8);
            ptx::umma_commit_elect(kv_empty_a + st * 8);
        }
    } else {
        // ===================== softmax + accumulate: thread == query row =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;            // row inside the tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float m_run = -INFINITY, l_run = 0.f, corr_prev = 0.f;
        float o_acc[DH];
#pragma unroll
        for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;

        auto accumulate = [&](int j, float corr) {   // O = O * corr + O_j
            ptx::mbar_wait(&o_full[j & 1], (j >> 1) & 1);
            ptx::tcgen05_fence_after();
            uint32_t v[DH];
            ptx::tmem_ld_32x32(t_lane + 128 + (j & 1) * 64, v);
            ptx::tmem_ld_32x32(t_lane + 128 + (j & 1) * 64 + 32, v + 32);
            ptx::tmem_ld_wait();
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&o_empty[j & 1]);
#pragma unroll
            for (int i = 0; i < DH; ++i) o_acc[i] = fmaf(o_acc[i], corr, __uint_as_float(v[i]));
        };

        for (int j = 0; j < n_tiles; ++j) {
            ptx::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            ptx::tcgen05_fence_after();
            uint32_t sv[TKV];
            ptx::tmem_ld_32x32(t_lane + (j & 1) * 64, sv);
            ptx::tmem_ld_32x32(t_lane + (j & 1) * 64 + 32, sv + 32);
            ptx::tmem_ld_wait();
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_empty[j & 1]);   // S buffer is free for tile j+2

            float* s = reinterpret_cast<float*>(sv);
            if (j == n_tiles - 1) {                              // keys beyond S belong to the next utterance: mask
                const int valid = S - j * TKV;
#pragma unroll
                for (int i = 0; i < TKV; ++i)
                    if (i >= valid) s[i] = -INFINITY;
            }
            // row maximum: four independent chains (a single 63-deep FMNMX chain is 250 cycles of pure latency per tile)
            float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
            for (int i = 4; i < TKV; i += 4) {
                mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const float m_new = fmaxf(m_run, mx);
            const float corr = fast_exp2((m_run - m_new) * LOG2E);   // 0 on the first tile (m_run = -inf)
            const float mb = m_new * LOG2E;
            // (exp2 in bf16x2 was tried: sm_100a has no packed SFU form, `ex2.approx.ftz.bf16x2` becomes two MUFU.EX2.BF16)
            float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;    // four independent chains for the row sum
            uint32_t pk[TKV / 2];
#pragma unroll
            for (int i = 0; i < TKV; i += 4) {
                const float p0 = fast_exp2(fmaf(s[i], LOG2E, -mb)), p1 = fast_exp2(fmaf(s[i + 1], LOG2E, -mb));
                const float p2 = fast_exp2(fmaf(s[i + 2], LOG2E, -mb)), p3 = fast_exp2(fmaf(s[i + 3], LOG2E, -mb));
                ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(p0, p1), t1 = __floats2bfloat162_rn(p2, p3);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t0);
                pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&t1);
            }
            const float ps = (ps0 + ps1) + (ps2 + ps3);
            l_run = l_run * corr + ps;
            m_run = m_new;
            // P row r -> 128B-swizzled K-major tile: 16-byte chunk c of row r lives at chunk (c ^ (r & 7)).
            // The buffer was last read by the P V MMA of tile j-2, whose completion (o_full) this thread
            // already observed in accumulate(j-2) during the previous iteration.
            uint8_t* prow = smem + SMEM_P + (j & 1) * P_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            ptx::fence_proxy_async();   // make the generic-proxy smem writes visible to the tensor core
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[j & 1]);

            if (j >= 1) accumulate(j - 1, corr_prev);           // overlaps with the MMAs of tile j
            corr_prev = corr;
        }
        accumulate(n_tiles - 1, corr_prev);

        const int row = q0 + r;
        if (row < S) {
            const float inv = 1.0f / l_run;
            bf16* o = out + ((size_t)(row_base + row)) * d + h * DH;
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 t = __floats2bfloat162_rn(o_acc[8 * c + 2 * i] * inv, o_acc[8 * c + 2 * i + 1] * inv);
                    w[i] = *reinterpret_cast<uint32_t*>(&t);
                }
                *reinterpret_cast<uint4*>(o + 8 * c) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------------------
// Variant with the output accumulator RESIDENT IN TMEM and lazy rescaling (the default since round 2).
// The kernel above reads every P V product back (tcgen05.ld of 64 columns), rescales and accumulates O in registers: 64 FFMA +
// a 64-register load + two barrier hand-offs per row and key tile, and 64 live accumulator registers per thread.  Here the
// P V MMAs accumulate into ONE TMEM region across all key tiles; P is computed against a reference maximum m_ref that is only
// moved when the row maximum exceeds it by more than 2^8 (then - rarely - the row of O is rescaled IN TMEM: tcgen05.ld, multiply,
// tcgen05.st, between the completion of the previous P V MMA and the hand-over of the next P).  exp2(s - m_ref) <= 256 keeps P
// well inside bf16 / fp32 range; the normalisation by the row sum at the end is exact in the same sense as before.
// ------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t LZ_TMEM_COLS = 256;           // S0 | S1 | O | (unused)
constexpr float LZ_THRESHOLD = 8.0f;             // log2 units

__global__ void __launch_bounds__(NUM_THREADS, 2)
enc_attn_tc_lazy_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        bf16* __restrict__ out, int S, int H) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t smem_a = raw_addr + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    // q_full 1 | k_full 3 | v_full 3 | kv_empty 3 | s_full 2 | s_empty 2 | p_full 2 | o_done 2
    const uint32_t q_full_a = smem_a + SMEM_BAR, k_full_a = q_full_a + 8, v_full_a = k_full_a + STAGES * 8;
    const uint32_t kv_empty_a = v_full_a + STAGES * 8, s_full_a = kv_empty_a + STAGES * 8, s_empty_a = s_full_a + 16;
    const uint32_t p_full_a = s_empty_a + 16, o_done_a = p_full_a + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 1 + 3 * STAGES + 8);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int d = H * DH;
    const int q0 = qt * TQ;
    const int row_base = b * S;
    const int n_tiles = (S + TKV - 1) / TKV;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmKV);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(&bars[0], 1);
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&bars[1 + s], 1);
            ptx::mbar_init(&bars[1 + STAGES + s], 1);
            ptx::mbar_init(&bars[1 + 2 * STAGES + s], 1);
        }
        uint64_t* b2 = bars + 1 + 3 * STAGES;
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&b2[s], 1);          // s_full
            ptx::mbar_init(&b2[2 + s], 4);      // s_empty: one arrive per softmax warp
            ptx::mbar_init(&b2[4 + s], 4);      // p_full
            ptx::mbar_init(&b2[6 + s], 1);      // o_done
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<LZ_TMEM_COLS>(tmem_ptr_smem);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        ptx::mbar_expect_tx_elect(q_full_a, Q_BYTES);
        ptx::tma_load_2d_elect(smem_a + SMEM_Q, &tmQ, q_full_a, h * DH, row_base + q0);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % STAGES;
            ptx::mbar_wait_addr(kv_empty_a + st * 8, ((j / STAGES) & 1) ^ 1);
            ptx::mbar_expect_tx_elect(k_full_a + st * 8, KV_BYTES);
            ptx::tma_load_2d_elect(smem_a + SMEM_K + st * KV_BYTES, &tmKV, k_full_a + st * 8, d + h * DH, row_base + j * TKV);
            ptx::mbar_expect_tx_elect(v_full_a + st * 8, KV_BYTES);
            ptx::tma_load_2d_elect(smem_a + SMEM_V + st * KV_BYTES, &tmKV, v_full_a + st * 8, 2 * d + h * DH, row_base + j * TKV);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_s = ptx::make_idesc_bf16(TQ, TKV, 0, 0);
        constexpr uint32_t idesc_o = ptx::make_idesc_bf16(TQ, DH, 0, 1);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t dq = ptx::make_smem_desc_sw128(smem_a + SMEM_Q, 1024, 16);
        auto issue_s = [&](int j) {
            const int st = j % STAGES;
            const uint64_t db = ptx::make_smem_desc_sw128(smem_a + SMEM_K + st * KV_BYTES, 1024, 16);
            const uint32_t dst = tmem_u + (j & 1) * 64;
            ptx::umma_f16_elect(dst, dq, db, idesc_s, 0);
            ptx::umma_f16_elect(dst, dq + 2, db + 2, idesc_s, 1);
            ptx::umma_f16_elect(dst, dq + 4, db + 4, idesc_s, 1);
            ptx::umma_f16_elect(dst, dq + 6, db + 6, idesc_s, 1);
            ptx::umma_commit_elect(s_full_a + (j & 1) * 8);
        };
        ptx::mbar_wait_addr(q_full_a, 0);
        ptx::mbar_wait_addr(k_full_a, 0);
        ptx::tcgen05_fence_after();
        issue_s(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) {
                const int jn = j + 1;
                ptx::mbar_wait_addr(k_full_a + (jn % STAGES) * 8, (jn / STAGES) & 1);
                ptx::mbar_wait_addr(s_empty_a + (jn & 1) * 8, ((jn >> 1) & 1) ^ 1);
                ptx::tcgen05_fence_after();
                issue_s(jn);
            }
            const int st = j % STAGES;
            ptx::mbar_wait_addr(v_full_a + st * 8, (j / STAGES) & 1);
            ptx::mbar_wait_addr(p_full_a + (j & 1) * 8, (j >> 1) & 1);
            ptx::tcgen05_fence_after();
            const uint64_t da = ptx::make_smem_desc_sw128(smem_a + SMEM_P + (j & 1) * P_BYTES, 1024, 16);
            const uint64_t db = ptx::make_smem_desc_sw128(smem_a + SMEM_V + st * KV_BYTES, 1024, KV_BYTES);
            const uint32_t dst = tmem_u + 128;                      // ONE accumulator for all key tiles
            ptx::umma_f16_elect(dst, da, db, idesc_o, j != 0);
            ptx::umma_f16_elect(dst, da + 2, db + 128, idesc_o, 1);
            ptx::umma_f16_elect(dst, da + 4, db + 256, idesc_o, 1);
            ptx::umma_f16_elect(dst, da + 6, db + 384, idesc_o, 1);
            ptx::umma_commit_elect(o_done_a + (j & 1) * 8);
            ptx::umma_commit_elect(kv_empty_a + st * 8);
        }
    } else {
        // ===================== softmax: thread == query row =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float mref = -INFINITY;      // reference maximum (log2 domain) the exponentials are taken against
        float l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            ptx::mbar_wait_addr(s_full_a + (j & 1) * 8, (j >> 1) & 1);
            ptx::tcgen05_fence_after();
            uint32_t sv[TKV];
            ptx::tmem_ld_32x32(t_lane + (j & 1) * 64, sv);
            ptx::tmem_ld_32x32(t_lane + (j & 1) * 64 + 32, sv + 32);
            ptx::tmem_ld_wait();
            ptx::tcgen05_fence_before();
            __syncwarp();
            ptx::mbar_arrive_elect(s_empty_a + (j & 1) * 8);     // S buffer is free for tile j+2
            float* s = reinterpret_cast<float*>(sv);
            if (j == n_tiles - 1) {
                const int valid = S - j * TKV;
#pragma unroll
                for (int i = 0; i < TKV; ++i)
                    if (i >= valid) s[i] = -INFINITY;
            }
            float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
            for (int i = 4; i < TKV; i += 4) {
                mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]); mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
            // the previous P V MMA is complete: O may be touched, and P[j & 1] (read by the MMA of tile j - 2) is free
            if (j > 0) {
                ptx::mbar_wait_addr(o_done_a + ((j - 1) & 1) * 8, ((j - 1) >> 1) & 1);
                ptx::tcgen05_fence_after();
            }
            const bool grow = mx > mref + LZ_THRESHOLD;          // first tile: mref = -inf -> true
            if (__any_sync(0xffffffffu, grow)) {
                const float mnew = grow ? mx : mref;
                if (j > 0) {                                       // rescale this row of O in TMEM (warp-collective; factor 1 where unchanged)
                    const float f = fast_exp2(mref - mnew);
                    l_run *= f;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t ov[32];
                        ptx::tmem_ld_32x32(t_lane + 128 + c * 32, ov);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
                        ptx::tmem_st_32x32(t_lane + 128 + c * 32, ov);
                    }
                    ptx::tmem_st_wait();
                    ptx::tcgen05_fence_before();
                }
                mref = mnew;
            }
            float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
            uint32_t pk[TKV / 2];
#pragma unroll
            for (int i = 0; i < TKV; i += 4) {
                const float p0 = fast_exp2(fmaf(s[i], LOG2E, -mref)), p1 = fast_exp2(fmaf(s[i + 1], LOG2E, -mref));
                const float p2 = fast_exp2(fmaf(s[i + 2], LOG2E, -mref)), p3 = fast_exp2(fmaf(s[i + 3], LOG2E, -mref));
                ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(p0, p1), t1 = __floats2bfloat162_rn(p2, p3);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t0);
                pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&t1);
            }
            l_run += (ps0 + ps1) + (ps2 + ps3);
            uint8_t* prow = smem + SMEM_P + (j & 1) * P_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            ptx::mbar_arrive_elect(p_full_a + (j & 1) * 8);
        }
        // ---- all key tiles accumulated: O / l
        ptx::mbar_wait_addr(o_done_a + ((n_tiles - 1) & 1) * 8, ((n_tiles - 1) >> 1) & 1);
        ptx::tcgen05_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l_run;
        bf16* o = out + ((size_t)(row_base + row)) * d + h * DH;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t ov[32];
            ptx::tmem_ld_32x32(t_lane + 128 + c2 * 32, ov);
            ptx::tmem_ld_wait();
            if (row < S) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(ov[8 * c + 2 * i]) * inv, __uint_as_float(ov[8 * c + 2 * i + 1]) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&t);
                    }
                    *reinterpret_cast<uint4*>(o + c2 * 32 + 8 * c) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<LZ_TMEM_COLS>(tmem_base);
}

}  // namespace

void encoder_attention_tc(const void* qkv, void* out, int B, int S, int H, cudaStream_t stream) {
    WB_REQUIRE(qkv && out && B > 0 && S > 0 && H > 0, "bad encoder attention arguments");
    const int d = H * DH;
    static PerDeviceOnce configured;
    configured([] {
        WB_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        WB_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_tc_lazy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    });
    static const bool eager_rescale = std::getenv("WB_ENC_ATTN_EAGER") != nullptr;   // (dev) A/B: the round-1 kernel
    const CUtensorMap tmQ = make_tmap_bf16_2d(qkv, 3LL * d, B * S, 3 * d, TQ);
    const CUtensorMap tmKV = make_tmap_bf16_2d(qkv, 3LL * d, B * S, 3 * d, TKV);
    dim3 grid(ceil_div(S, TQ), H, B), block(NUM_THREADS);
    WB_REQUIRE(grid.z <= 65535, "batch too large for one launch");
    if (eager_rescale) enc_attn_tc_kernel<<<grid, block, SMEM_TOTAL, stream>>>(tmQ, tmKV, (bf16*)out, S, H);
    else enc_attn_tc_lazy_kernel<<<grid, block, SMEM_TOTAL, stream>>>(tmQ, tmKV, (bf16*)out, S, H);
    WB_CHECK_LAUNCH();
}

}  // namespace wb

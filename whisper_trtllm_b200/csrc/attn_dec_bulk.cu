// Cross-attention for the decode step as a bulk-copy (TMA 1-D) streaming kernel: bf16, contiguous K/V [B, H, n, 64].
//
// Same arithmetic as decode_attn_kernel<bf16,false> (attn_dec.cu), different data movement: ONE persistent CTA per SM;
// a producer thread streams every (utterance, head) item's K and V rows with cp.async.bulk (8 KB per copy, 64 keys)
// into an 8-stage / 128 KB shared-memory ring, completion on mbarriers; 8 consumer warps do the online softmax out of
// shared memory.  The bytes in flight live in shared memory instead of registers (the LDG kernel needs 4 CTAs x 256
// threads x 64 registers = the whole register file to reach the same depth), so the kernel leaves ~3/4 of the register
// file and 90 KB of shared memory of every SM to kernels of a concurrent stream, and the ring keeps filling across
// item boundaries (the LDG kernel drains its pipeline at every item).
// Reference semantics: WhisperDecoderAttention cross / cache mode (model.py:261-272, 292-300; modeling_whisper.py:474-481).
#include "wb_internal.h"
#include "wb_ptx.cuh"

namespace wb {

namespace {
constexpr int DH = 64;
constexpr int CHUNK_KEYS = 64;
constexpr uint32_t ROW_BYTES = DH * 2;
constexpr uint32_t CHUNK_BYTES = CHUNK_KEYS * ROW_BYTES;      // 8 KB of K (and 8 KB of V) per stage
constexpr uint32_t STAGE_BYTES = 2 * CHUNK_BYTES;
constexpr int STAGES = 8;
constexpr int CW = 8;                                         // consumer warps
static_assert(STAGES == CW, "stage s is owned by consumer warp s");
constexpr int THREADS = (1 + CW) * 32;
constexpr uint32_t SMEM_BAR = STAGES * STAGE_BYTES;
constexpr uint32_t SMEM_PART = SMEM_BAR + 2 * STAGES * 8;
constexpr uint32_t SMEM_TOTAL = SMEM_PART + 2 * CW * (2 + DH) * 4;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1) cross_attn_bulk_kernel(DecAttnArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* empty = full + STAGES;
    float* part_m = reinterpret_cast<float*>(smem + SMEM_PART);   // [2][CW]
    float* part_l = part_m + 2 * CW;                              // [2][CW]
    float* part_o = part_l + 2 * CW;                              // [2][CW][DH]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);   // released by the one consumer warp that owns the stage
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    if (a.active != nullptr && *a.active == 0) return;

    const int n = a.n_keys;
    const int n_chunks = (n + CHUNK_KEYS - 1) / CHUNK_KEYS;
    const int n_items = a.B * a.H;

    if (warp == 0) {
        // ===================== producer: stream K and V of every item through the ring =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / a.H, h = item - b * a.H;
            if (a.row_active != nullptr && a.row_active[b] == 0) continue;   // finished utterance (consumers skip it too)
            const size_t o = (size_t)b * a.kv_bstride + (size_t)h * a.kv_hstride;
            const bf16* kb = reinterpret_cast<const bf16*>(a.k) + o;
            const bf16* vb = reinterpret_cast<const bf16*>(a.v) + o;
            for (int c = 0; c < n_chunks; ++c) {
                const uint32_t bytes = (uint32_t)min(CHUNK_KEYS, n - c * CHUNK_KEYS) * ROW_BYTES;
                ptx::mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* dst = smem + stage * STAGE_BYTES;
                    ptx::mbar_expect_tx(&full[stage], 2 * bytes);
                    bulk_g2s(dst, kb + (size_t)c * CHUNK_KEYS * DH, bytes, &full[stage]);
                    bulk_g2s(dst + CHUNK_BYTES, vb + (size_t)c * CHUNK_KEYS * DH, bytes, &full[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== consumers: online softmax out of shared memory =====================
        const int cw = warp - 1, ct = tid - 32;
        const int sub = lane & 7, grp = lane >> 3;     // 8 lanes per key row (16 B each), 4 rows per warp instruction
        int parity = 0, chunk_base = 0;   // chunk_base: position of the item's first chunk in the ring, mod 8
        uint32_t phase = 0;               // parity of this warp's own stage
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / a.H, h = item - b * a.H;
            if (a.row_active != nullptr && a.row_active[b] == 0) continue;
            parity ^= 1;
            float qf[8];
            if (a.q_parts != nullptr) {   // q = bias + sum of the split-K slabs of the q projection (fp32, fixed order)
                const int c0 = h * DH + sub * 8;
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                    float4 t = a.q_bias != nullptr ? *reinterpret_cast<const float4*>(a.q_bias + c0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int p = 0; p < a.q_n_parts; ++p) {
                        const float4 u = *reinterpret_cast<const float4*>(a.q_parts + (size_t)p * a.q_part_stride + (size_t)b * a.H * DH + c0 + i);
                        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                    }
                    qf[i] = t.x; qf[i + 1] = t.y; qf[i + 2] = t.z; qf[i + 3] = t.w;
                }
            } else {
                ld16(reinterpret_cast<const bf16*>(a.q) + (size_t)b * a.q_stride + h * DH + sub * 8).unpack(qf);
            }
            float m_run = -INFINITY, l_run = 0.f;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;

            // Stage s of the ring belongs to consumer warp s (STAGES == CW): chunk g of this CTA's chunk sequence lands in
            // stage g % 8 and is consumed by warp g % 8 alone, 64 keys per mbarrier wait, 16 rows in flight per lane group.
            const int first = ((cw - chunk_base) % CW + CW) % CW;
            for (int c = first; c < n_chunks; c += CW) {
                const int keys = min(CHUNK_KEYS, n - c * CHUNK_KEYS);
                ptx::mbar_wait(&full[cw], phase);
                const uint8_t* ks = smem + cw * STAGE_BYTES;
                const uint8_t* vs = ks + CHUNK_BYTES;
                for (int r0 = 0; r0 < keys; r0 += 16) {
                    Vec16<bf16> kr[4], vr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = min(r0 + u * 4 + grp, keys - 1);
                        kr[u].raw = *reinterpret_cast<const uint4*>(ks + r * ROW_BYTES + sub * 16);
                        vr[u].raw = *reinterpret_cast<const uint4*>(vs + r * ROW_BYTES + sub * 16);
                    }
                    float sc[4];
                    float mb = -INFINITY;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float kf[8];
                        kr[u].unpack(kf);
                        float dot = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) dot = fmaf(qf[i], kf[i], dot);
                        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                        sc[u] = (r0 + u * 4 + grp < keys) ? dot : -INFINITY;
                        mb = fmaxf(mb, sc[u]);
                    }
                    const float m_new = fmaxf(m_run, mb);
                    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
                    const float scale = __expf(m_run - m_use);
                    l_run *= scale;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] *= scale;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float p = __expf(sc[u] - m_use);
                        l_run += p;
                        float vf[8];
                        vr[u].unpack(vf);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, vf[i], acc[i]);
                    }
                    m_run = m_new;
                }
                __syncwarp();                                    // every lane is done reading this stage
                if (lane == 0) ptx::mbar_arrive(&empty[cw]);
                phase ^= 1;
            }
            chunk_base = (chunk_base + n_chunks) % CW;

            // ---- merge the 4 key groups of the warp, then the 8 warps
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, m_run, o);
                const float ol = __shfl_xor_sync(0xffffffffu, l_run, o);
                const float mm = fmaxf(m_run, om);
                const float s1 = (m_run == -INFINITY) ? 0.f : __expf(m_run - mm);
                const float s2 = (om == -INFINITY) ? 0.f : __expf(om - mm);
                l_run = l_run * s1 + ol * s2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
                    acc[i] = acc[i] * s1 + oa * s2;
                }
                m_run = mm;
            }
            float* pm = part_m + parity * CW;
            float* pl = part_l + parity * CW;
            float* po = part_o + parity * CW * DH;
            if (grp == 0) {
                if (sub == 0) { pm[cw] = m_run; pl[cw] = l_run; }
#pragma unroll
                for (int i = 0; i < 8; ++i) po[cw * DH + sub * 8 + i] = acc[i];
            }
            consumer_barrier();   // partials are double buffered by item parity: one barrier per item
            if (ct < DH) {
                float mm = pm[0];
#pragma unroll
                for (int w = 1; w < CW; ++w) mm = fmaxf(mm, pm[w]);
                float l = 0.f, o = 0.f;
#pragma unroll
                for (int w = 0; w < CW; ++w) {
                    const float s = (pm[w] == -INFINITY) ? 0.f : __expf(pm[w] - mm);
                    l = fmaf(pl[w], s, l);
                    o = fmaf(po[w * DH + ct], s, o);
                }
                reinterpret_cast<bf16*>(a.out)[(size_t)b * a.out_stride + h * DH + ct] = __float2bfloat16_rn(o / l);
            }
        }
    }
}
}  // namespace

bool decode_attention_bulk_supported(const DecAttnArgs& a) {
    if (a.dtype != BF16 || a.k_pages != nullptr || a.state != nullptr || a.n_keys <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.k) & 15) || (reinterpret_cast<uintptr_t>(a.v) & 15)) return false;
    return (a.kv_bstride * 2) % 16 == 0 && (a.kv_hstride * 2) % 16 == 0;
}

void decode_attention_bulk(const DecAttnArgs& a, cudaStream_t stream) {
    WB_REQUIRE(decode_attention_bulk_supported(a), "bulk cross-attention needs bf16, contiguous 16-byte aligned K/V");
    WB_REQUIRE((a.q || a.q_parts) && a.out && a.B > 0 && a.H > 0, "bad decode attention arguments");
    static PerDeviceOnce configured;
    configured([] { WB_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL)); });
    const int sms = device_sm_count();
    const int grid = std::min(a.B * a.H, sms);
    launch_kernel(cross_attn_bulk_kernel, dim3(grid), dim3(THREADS), SMEM_TOTAL, stream, true, a);
}

}  // namespace wb

// tcgen05 GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
//
// Persistent, warp-specialised (one CTA per SM):
//   warp 0   TMA producer   cp.async.bulk.tensor (128-byte swizzle) into a STAGES-deep smem ring
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma (UMMA 128 x BN x 16), tcgen05.commit frees slots
//   warp 2   TMEM allocator 2 x BN fp32 accumulator columns (double buffered so the epilogue of tile i
//                            overlaps the main loop of tile i+1)
//   warps 4-11 epilogue     tcgen05.ld (lane quarter = warp % 4, column half = (warp - 4) / 4) -> 32x32 transpose through
//                            XOR-swizzled smem -> bias / erf-GELU / fp32 residual -> row-contiguous 128-byte stores
//                            (4 rows per warp instruction instead of 32 scattered 16-byte pieces)
//
// Replaces the TensorRT-chosen fp32 tactics behind the reference's ColumnLinear/RowLinear/Conv2d layers
// (tensorrt_llm/layers/linear.py:38-139, models/whisper/model.py:77-79) and the oracle's F.linear/conv1d.
#include "wb_epilogue.cuh"
#include "wb_ptx.cuh"

#include <atomic>
#include <mutex>
#include <unordered_map>

namespace wb {

namespace {

constexpr int BM = 128, BK = 64;
constexpr uint32_t A_BYTES = BM * BK * 2;
// kLean: the decode-step variant that must CO-RESIDE with the bulk-ring cross-attention CTA of a concurrent stream
// (wb_decode_run_multi): 3 stages + 4 epilogue warps = 256 threads, <= 128 registers, <= 90 KB of shared memory.
template <int BN, bool kLean> struct TcCfg {
    static constexpr int EPI_WARPS = kLean ? 4 : 8;
    static constexpr int NUM_THREADS = (4 + EPI_WARPS) * 32;
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = kLean ? 3 : (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
    static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32,64,128,256}
    static constexpr uint32_t STAGING_BYTES = EPI_WARPS * 32 * 32 * 4;  // one 32x32 fp32 transpose tile per epilogue warp
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int BN, bool kLean>
__global__ void __launch_bounds__((TcCfg<BN, kLean>::NUM_THREADS), (kLean ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int K,
               int num_m_tiles, int num_n_tiles, int k_splits, long long split_stride, EpiParams ep,
               const int* __restrict__ active) {
    using Cfg = TcCfg<BN, kLean>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_WARPS = Cfg::EPI_WARPS;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte aligned bases
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    // warp-uniform by construction (the role branches below are then uniform branches: the MMA / TMA issuing code stays on
    // the uniform datapath, see the note at the MMA issuer)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int num_tiles = num_m_tiles * num_n_tiles * k_splits;   // work item = (m tile, n tile, k split)
    const int nk = K / BK / k_splits;                             // k blocks per work item
    const uint32_t smem_a = raw_addr + ((1024u - (raw_addr & 1023u)) & 1023u);   // shared-memory address of the ring
    constexpr uint32_t BAR_OFF = STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES;
    const uint32_t full_a = smem_a + BAR_OFF, empty_a = full_a + STAGES * 8;
    const uint32_t tfull_a = empty_a + STAGES * 8, tempty_a = tfull_a + 16;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmW);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // PDL: everything above overlapped the previous kernel's tail; from here on we touch its outputs
    pdl_wait();
    pdl_trigger();
    // decode loop already stopped (*active == 0): the tile is still computed but nothing is stored — the flag is only
    // needed by the epilogue, so its load latency hides behind the main loop instead of delaying the first TMA
    const bool run = (active == nullptr) || (*active != 0);

    if (warp == 0) {
        // ===================== TMA producer (converged warp, the elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mn = tile / k_splits, ks = tile - mn * k_splits;
            const int m_blk = mn / num_n_tiles, n_blk = mn - m_blk * num_n_tiles;
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait_addr(empty_a + stage * 8, phase ^ 1);
                const uint32_t sa = smem_a + stage * Cfg::STAGE_BYTES;
                ptx::mbar_expect_tx_elect(full_a + stage * 8, Cfg::STAGE_BYTES);
                ptx::tma_load_2d_elect(sa, &tmA, full_a + stage * 8, (ks * nk + kb) * BK, m_blk * BM);
                ptx::tma_load_2d_elect(sa + A_BYTES, &tmW, full_a + stage * 8, (ks * nk + kb) * BK, n_blk * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, the elected lane issues) =====================
        // tcgen05.mma / tcgen05.commit read their operands from UNIFORM registers.  Issued from inside an `if (lane == 0)`
        // region ptxas moved every operand across with ELECT + R2UR.BROADCAST, ~100 cycles per MMA (measured with clock stamps
        // in the decode chains, profiles/r02_chain_trace_*.md): 4 MMAs of a 128 x 256 x 64 k-block need 512 cycles of tensor
        // pipe, their issue took ~650.  With the warp converged and all operands derived from kernel parameters and uniform
        // loop counters the descriptors are computed on the uniform datapath and the MMAs issue back to back.
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            ptx::mbar_wait_addr(tempty_a + acc * 8, acc_phase ^ 1);  // epilogue drained this accumulator
            ptx::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_u + acc * BN;
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait_addr(full_a + stage * 8, phase);  // TMA bytes landed
                ptx::tcgen05_fence_after();
                const uint32_t sa = smem_a + stage * Cfg::STAGE_BYTES;
                const uint64_t da = ptx::make_smem_desc_sw128(sa, 1024, 16);
                const uint64_t db = ptx::make_smem_desc_sw128(sa + A_BYTES, 1024, 16);
                ptx::umma_f16_elect(d_tmem, da, db, idesc, kb != 0);      // advance 16 bf16 = 32 B inside the swizzle atom
                ptx::umma_f16_elect(d_tmem, da + 2, db + 2, idesc, 1);
                ptx::umma_f16_elect(d_tmem, da + 4, db + 4, idesc, 1);
                ptx::umma_f16_elect(d_tmem, da + 6, db + 6, idesc, 1);
                ptx::umma_commit_elect(empty_a + stage * 8);                     // smem slot reusable when MMAs finish
                if (kb == nk - 1) ptx::umma_commit_elect(tfull_a + acc * 8);     // accumulator complete
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;             // the TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;   // which share of the tile's 32-column slabs this warp owns
        constexpr int SLABS = BN / 32;
        constexpr int HALVES = EPI_WARPS / 4;
        constexpr int SLABS_PER_WARP = (SLABS + HALVES - 1) / HALVES;
        float4* st4 = reinterpret_cast<float4*>(staging + (warp - 4) * 32 * 32);   // [32 rows][8 chunks of 16 B], chunk ^= row & 7
        const int rrow = lane >> 3, rchunk = lane & 7;                             // read-back mapping: 4 rows x 128 B per instruction
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mn = tile / k_splits, ks = tile - mn * k_splits;
            const int m_blk = mn / num_n_tiles, n_blk = mn - m_blk * num_n_tiles;
            EpiParams ept = ep;
            ept.out = reinterpret_cast<float*>(ep.out) + ks * split_stride;   // split-K: raw fp32 partial slab (offset 0 otherwise)
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            const int row0 = m_blk * BM + q * 32;
            if (ept.am_val != nullptr) {
                // ---- fused logits processors + argmax (LM head): thread == row, no transposition, no logits in HBM
                const int row = row0 + lane;
                int bits = 1;
                if (row < ept.M) bits |= ((ept.am_row_len != nullptr ? ept.am_row_len[row] : ept.am_state->cur_len) == ept.am_begin) ? 2 : 0;
                float best = -INFINITY;
                int besti = 0x7fffffff;
#pragma unroll 1
                for (int ci = 0; ci < SLABS_PER_WARP; ++ci) {
                    const int c = half * SLABS_PER_WARP + ci;
                    if (c >= SLABS) break;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_row + c * 32, v);
                    ptx::tmem_ld_wait();
                    const int nb = n_blk * BN + c * 32;
                    if (nb < ept.N) {                         // (N % 8 == 0: a slab is valid in whole 8-column groups)
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            if (nb + g8 * 8 < ept.N) {
                                const uint2 mm = __ldg(reinterpret_cast<const uint2*>(ept.am_mask + nb + g8 * 8));   // same address in every lane
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const unsigned mbyte = ((j < 4 ? mm.x : mm.y) >> ((j & 3) * 8)) & 0xffu;
                                    const float val = __uint_as_float(v[g8 * 8 + j]);
                                    if ((mbyte & bits) == 0 && val > best) { best = val; besti = nb + g8 * 8 + j; }
                                }
                            }
                        }
                    }
                }
                if (row < ept.M && run) {
                    const size_t e = (size_t)row * ept.am_stride + (size_t)n_blk * HALVES + half;
                    ept.am_val[e] = best;
                    ept.am_idx[e] = besti;
                }
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            const int g0 = ept.m_period_in > 0 ? row0 / ept.m_period_in : 0;   // one division per tile (period >= 128 rows)
#pragma unroll 1
            for (int ci = 0; ci < SLABS_PER_WARP; ++ci) {
                const int c = half * SLABS_PER_WARP + ci;
                if (c >= SLABS) break;
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_row + c * 32, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)   // thread = row `lane`: conflict-free 16-byte stores (8 lanes hit 8 distinct chunks)
                    st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                __syncwarp();
                const int n0 = n_blk * BN + c * 32 + rchunk * 4;
                // phase 1: everything this slab needs from memory is requested before the first use (the residual may
                // alias the output, so the compiler must not be left to interleave these loads with the stores below)
                EpiRow er[8];
                float4 f[8], rs[8];
                const bool col_ok = n0 < ept.N;
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ept.bias != nullptr && col_ok) bb = __ldg(reinterpret_cast<const float4*>(ept.bias + n0));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + rrow;
                    er[i] = epi_row_near(ept, row0 + rr, g0);
                    er[i].valid = er[i].valid && col_ok && run;
                    f[i] = st4[rr * 8 + (rchunk ^ (rr & 7))];
                    rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ept.res != nullptr && er[i].valid)
                        rs[i] = *reinterpret_cast<const float4*>(ept.res + er[i].res_row * ept.ldres + n0);
                }
                // phase 2: bias -> erf-GELU -> residual -> typed store
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float o[4] = {f[i].x + bb.x, f[i].y + bb.y, f[i].z + bb.z, f[i].w + bb.w};
                    if (ept.act == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = gelu_erf_fast(o[j]);
                    }
                    o[0] += rs[i].x; o[1] += rs[i].y; o[2] += rs[i].z; o[3] += rs[i].w;
                    epi_write<4>(ept, er[i], n0, o);
                }
                __syncwarp();   // staging tile is rewritten by the next slab
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeFn>(p);
    });
    if (fn == nullptr) throw Error(-3, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

struct MapKey {
    const void* ptr; long long ld; int rows, cols, box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && ld == o.ld && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<long long>()(k.ld);
        h = h * 1000003u ^ (size_t)k.rows;
        h = h * 1000003u ^ (size_t)k.cols;
        h = h * 1000003u ^ (size_t)k.box_rows;
        return h;
    }
};

}  // namespace

// 2-D bf16 row-major [rows, cols] tensor map with a {64, box_rows} box and 128-byte swizzle.  Cached:
// the decode loop re-issues the same few dozen GEMMs every step.
CUtensorMap make_tmap_bf16_2d(const void* ptr, long long ld_elems, int rows, int cols, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, ld_elems, rows, cols, box_rows};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    WB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
    WB_REQUIRE((ld_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    return m;
}

// The same row-major bf16 [rows, cols] matrix seen as a 3-D tensor {64 (k inside a k-block), rows, cols / 64 (k-block)} with
// 128-byte swizzle: a box {64, box_rows, box_k} lands in shared memory as box_k consecutive [box_rows x 128 B] K-major swizzled
// slices - box_k k-blocks of a GEMM operand with ONE TMA operation (decode chains, step_chain.cu).
CUtensorMap make_tmap_bf16_kgroups(const void* ptr, long long ld_elems, int rows, int cols, int box_rows, int box_k) {
    WB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
    WB_REQUIRE((ld_elems * 2) % 16 == 0 && cols % 64 == 0, "TMA row pitch must be a multiple of 16 bytes, K a multiple of 64");
    WB_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_k >= 1 && box_k <= cols / 64, "bad TMA box");
    CUtensorMap m;
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)ld_elems * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_k};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(-3, "cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r));
    return m;
}

int sm_count() { return device_sm_count(); }

bool gemm_tc_supported(const GemmArgs& a) {
    if (a.in_dtype != BF16) return false;
    if (a.K % BK != 0 || a.N % 8 != 0) return false;
    if ((a.lda * 2) % 16 != 0 || (a.ldw * 2) % 16 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return false;
    return true;
}

template <int BN, bool kLean>
static void launch_tc(const GemmArgs& a, int k_splits, cudaStream_t stream) {
    using Cfg = TcCfg<BN, kLean>;
    static PerDeviceOnce configured;
    configured([] { WB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, kLean>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); });
    const CUtensorMap tmA = make_tmap_bf16_2d(a.A, a.lda, a.M, a.K, BM);
    const CUtensorMap tmW = make_tmap_bf16_2d(a.W, a.ldw, a.N, a.K, BN);
    const int mt = ceil_div(a.M, BM), nt = ceil_div(a.N, BN);
    const int grid = std::min(mt * nt * k_splits, sm_count());
    EpiParams ep = make_epi(a);
    if (a.am_count != nullptr) *a.am_count = nt * (Cfg::EPI_WARPS / 4);
    launch_kernel(gemm_tc_kernel<BN, kLean>, dim3(grid), dim3(Cfg::NUM_THREADS), Cfg::SMEM_BYTES, stream, true, tmA, tmW, a.K, mt, nt,
                  k_splits, a.split_stride, ep, a.active);
}

// Process-wide A/B switches (wb_set_*): atomics, read once per launch; sessions can override the ones that select their step
// path with wb_session_set_option.  Defaults are the production configuration.
static std::atomic<int> g_force_bn{0};
void set_gemm_tc_block_n(int bn) { g_force_bn = bn; }
static std::atomic<bool> g_lean_decode{false};   // skinny GEMMs use the co-residency-friendly variant (set for multi-stream decoding)
void set_lean_decode_gemm(bool on) { g_lean_decode = on; }

// Tile width / split-K choice.  Large M (encoder): the widest tile that still gives every SM a tile.  Skinny M (decode,
// one or two M tiles): the kernel is bound by how fast one SM can pull operand bytes (~40 B/clk/SM measured: 64 CTAs
// x 320 KB took ~8 us), so minimise bytes per CTA = (BM + BN) * K/S * 2 over the configurations that keep <= one wave,
// charging split-K for the partial slabs its consumer has to re-read (~M*N*4*S bytes each way through L2).
static void pick_config(const GemmArgs& a, int max_splits, int& bn_out, int& splits_out) {
    const int sms = sm_count();
    const int mt = ceil_div(a.M, BM);
    const int nkb = a.K / BK;
    if (mt > 2) {
        int bn = 256;
        while (bn > 32 && mt * ceil_div(a.N, bn) < sms) bn >>= 1;
        bn_out = bn; splits_out = 1;
        return;
    }
    double best = 1e30;
    bn_out = 32; splits_out = 1;
    for (int bn : {32, 64, 128, 256}) {
        if (g_lean_decode && bn > 64 && a.N <= 8192) continue;   // lean variant exists for BN <= 64 (LM head keeps the wide tile)
        for (int s = 1; s <= max_splits; s *= 2) {
            if (nkb % s != 0) continue;
            const long long ctas = (long long)mt * ceil_div(a.N, bn) * s;
            const double waves = (double)((ctas + sms - 1) / sms);
            const double ingest = (double)(BM + bn) * (a.K / s) * 2.0;                  // bytes one CTA pulls per work item
            const double partial = s > 1 ? (double)a.M * a.N * 4.0 * s * 2.0 / 6000.0 : 0.0;   // clk, chip-wide L2 rate
            const double t = waves * (ingest / 40.0 + 1500.0) + partial;                // clk; 1500 = pipeline fill + epilogue
            if (t < best) { best = t; bn_out = bn; splits_out = s; }
        }
    }
}

int gemm_argmax_partials(int N) { return ceil_div(N, 32) * 2; }   // upper bound for every tile width (narrowest tile 32, two halves)

void gemm_tc(const GemmArgs& a, cudaStream_t stream) {
    validate_gemm_common(a);
    if (a.am_val != nullptr)
        WB_REQUIRE(a.am_idx != nullptr && a.am_mask != nullptr && a.am_state != nullptr && a.k_splits == 1 && a.out_mode == 0 &&
                   a.m_period_in == 0 && a.am_stride >= gemm_argmax_partials(a.N),
                   "fused argmax: needs index / mask / state buffers, no split-K, no row remap, am_stride >= gemm_argmax_partials(N)");
    WB_REQUIRE(gemm_tc_supported(a), "shape/alignment not supported by the tcgen05 GEMM");
    int bn = 256, splits = 1;
    const bool auto_split = a.k_splits == 0;
    pick_config(a, auto_split ? std::max(1, a.max_k_splits) : 1, bn, splits);
    if (!auto_split) splits = a.k_splits;
    if (const int forced = g_force_bn.load(std::memory_order_relaxed)) bn = forced;
    WB_REQUIRE(splits >= 1 && (a.K / BK) % splits == 0, "k_splits must divide K / 64");
    if (splits > 1 || auto_split)
        WB_REQUIRE(a.out_dtype == F32 && a.bias == nullptr && a.res == nullptr && a.act == 0 && a.out_mode == 0 && a.out2 == nullptr,
                   "split-K stores raw fp32 partials: bias / activation / residual belong to the consumer");
    if (a.chosen_splits) *a.chosen_splits = splits;
    const bool lean = g_lean_decode && ceil_div(a.M, BM) <= 2 && bn <= 64;
    switch (bn) {
        case 256: launch_tc<256, false>(a, splits, stream); break;
        case 128: launch_tc<128, false>(a, splits, stream); break;
        case 64: if (lean) launch_tc<64, true>(a, splits, stream); else launch_tc<64, false>(a, splits, stream); break;
        case 32: if (lean) launch_tc<32, true>(a, splits, stream); else launch_tc<32, false>(a, splits, stream); break;
        default: WB_REQUIRE(false, "unsupported BLOCK_N");
    }
}

static std::atomic<int> g_gemm_backend{0};
void set_gemm_backend(int backend) { g_gemm_backend = backend; }
int get_gemm_backend() { return g_gemm_backend; }

void gemm(const GemmArgs& a, cudaStream_t stream) {
    if (g_gemm_backend == 0 && gemm_tc_supported(a)) {
        gemm_tc(a, stream);
    } else {
        WB_REQUIRE(a.am_val == nullptr, "the fused argmax epilogue exists on the tcgen05 GEMM only");
        if (a.chosen_splits) *a.chosen_splits = 1;   // the CUDA-core kernel never splits: one "partial" slab = the whole product
        gemm_simt(a, stream);
    }
}

}  // namespace wb

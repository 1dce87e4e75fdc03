// tcgen05 GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
//
// Persistent, warp-specialised (one CTA per SM):
//   warp 0   TMA producer   cp.async.bulk.tensor (128-byte swizzle) into a STAGES-deep smem ring
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma (UMMA 128 x BN x 16), tcgen05.commit frees slots
//   warp 2   TMEM allocator 2 x BN fp32 accumulator columns (double buffered so the epilogue of tile i
//                            overlaps the main loop of tile i+1)
//   warps 4-7 epilogue      tcgen05.ld (lane quarter = warp % 4) -> bias / erf-GELU / fp32 residual -> store
//
// Replaces the TensorRT-chosen fp32 tactics behind the reference's ColumnLinear/RowLinear/Conv2d layers
// (tensorrt_llm/layers/linear.py:38-139, models/whisper/model.py:77-79) and the oracle's F.linear/conv1d.
#include "wb_epilogue.cuh"
#include "wb_ptx.cuh"

#include <mutex>
#include <unordered_map>

namespace wb {

namespace {

constexpr int BM = 128, BK = 64;
constexpr uint32_t A_BYTES = BM * BK * 2;

template <int BN> struct TcCfg {
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
    static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {32,64,128,256}
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int K,
               int num_m_tiles, int num_n_tiles, EpiParams ep, const int* __restrict__ active) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    if (active != nullptr && *active == 0) return;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte aligned bases
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = num_m_tiles * num_n_tiles;
    const int nk = K / BK;

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
            ptx::mbar_init(&tmem_empty_bar[s], 4);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n_tiles, n_blk = tile - m_blk * num_n_tiles;
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sa + A_BYTES, &tmW, &full_bar[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
            ptx::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);  // TMA bytes landed
                ptx::tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = ptx::make_smem_desc_sw128(sa, 1024, 16);
                    const uint64_t db = ptx::make_smem_desc_sw128(sa + A_BYTES, 1024, 16);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)  // advance 16 bf16 = 32 B inside the swizzle atom
                        ptx::umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    ptx::umma_commit(&empty_bar[stage]);                     // smem slot reusable when MMAs finish
                    if (kb == nk - 1) ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n_tiles, n_blk = tile - m_blk * num_n_tiles;
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tcgen05_fence_after();
            const EpiRow r = epi_row(ep, m_blk * BM + q * 32 + lane);
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_row + c * 32, v);
                ptx::tmem_ld_wait();
                const int n0 = n_blk * BN + c * 32;
#pragma unroll
                for (int g = 0; g < 4; ++g) epi_store<8, true>(ep, r, n0 + g * 8, reinterpret_cast<float*>(v) + g * 8);
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

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        WB_CHECK_CUDA(cudaGetDevice(&dev));
        WB_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

bool gemm_tc_supported(const GemmArgs& a) {
    if (a.in_dtype != BF16) return false;
    if (a.K % BK != 0 || a.N % 8 != 0) return false;
    if ((a.lda * 2) % 16 != 0 || (a.ldw * 2) % 16 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) return false;
    return true;
}

template <int BN>
static void launch_tc(const GemmArgs& a, cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    static bool configured = false;
    if (!configured) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const CUtensorMap tmA = make_tmap_bf16_2d(a.A, a.lda, a.M, a.K, BM);
    const CUtensorMap tmW = make_tmap_bf16_2d(a.W, a.ldw, a.N, a.K, BN);
    const int mt = ceil_div(a.M, BM), nt = ceil_div(a.N, BN);
    const int grid = std::min(mt * nt, sm_count());
    EpiParams ep = make_epi(a);
    gemm_tc_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmW, a.K, mt, nt, ep, a.active);
    WB_CHECK_LAUNCH();
}

static int g_force_bn = 0;
void set_gemm_tc_block_n(int bn) { g_force_bn = bn; }

void gemm_tc(const GemmArgs& a, cudaStream_t stream) {
    validate_gemm_common(a);
    WB_REQUIRE(gemm_tc_supported(a), "shape/alignment not supported by the tcgen05 GEMM");
    // pick the widest tile that still yields about one tile per SM
    const int mt = ceil_div(a.M, BM);
    int bn = 256;
    if (g_force_bn) {
        bn = g_force_bn;
    } else {
        const int sms = sm_count();
        while (bn > 32 && mt * ceil_div(a.N, bn) < sms) bn >>= 1;
    }
    switch (bn) {
        case 256: launch_tc<256>(a, stream); break;
        case 128: launch_tc<128>(a, stream); break;
        case 64: launch_tc<64>(a, stream); break;
        case 32: launch_tc<32>(a, stream); break;
        default: WB_REQUIRE(false, "unsupported BLOCK_N");
    }
}

static int g_gemm_backend = 0;
void set_gemm_backend(int backend) { g_gemm_backend = backend; }
int get_gemm_backend() { return g_gemm_backend; }

void gemm(const GemmArgs& a, cudaStream_t stream) {
    if (g_gemm_backend == 0 && gemm_tc_supported(a)) gemm_tc(a, stream);
    else gemm_simt(a, stream);
}

}  // namespace wb

// HBM read-bandwidth probes (measurement tools, not on the product path): what a pure streaming read can reach on
// this GPU, to judge how much headroom the cross-attention kernel (the dominant HBM stream) has left.
//   mode 0  LDG.128 L1-bypassing loads, persistent grid, 8 requests in flight per lane (the attention kernel's pattern)
//   mode 1  cp.async.bulk (TMA 1-D) global -> shared ring, one producer thread per CTA, 8 x 16 KB stages
#include "wb_internal.h"
#include "wb_ptx.cuh"

namespace wb {
namespace {

__global__ void __launch_bounds__(256) probe_ldg_kernel(const uint4* __restrict__ p, size_t n_vec, unsigned* __restrict__ sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n_vec; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n_vec; i += stride) { const uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;   // keep the loads alive
}

constexpr int PB_STAGES = 8;
constexpr uint32_t PB_BYTES = 16384;

__global__ void __launch_bounds__(128) probe_bulk_kernel(const uint8_t* __restrict__ p, size_t n_chunks, unsigned* __restrict__ sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + PB_STAGES * PB_BYTES);
    uint64_t* empty = full + PB_STAGES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < PB_STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 3); }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                ptx::mbar_expect_tx(&full[stage], PB_BYTES);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ptx::smem_u32(smem + stage * PB_BYTES)), "l"(p + c * PB_BYTES), "r"(PB_BYTES),
                               "r"(ptx::smem_u32(&full[stage])) : "memory");
            }
            __syncwarp();
            if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        int stage = 0; uint32_t phase = 0;
        unsigned acc = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            ptx::mbar_wait(&full[stage], phase);
            const uint4* s4 = reinterpret_cast<const uint4*>(smem + stage * PB_BYTES);
            for (int i = (warp - 1) * 32 + lane; i < (int)(PB_BYTES / 16); i += 96) { const uint4 v = s4[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[stage]);
            if (++stage == PB_STAGES) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}
}  // namespace

void bandwidth_probe(const void* buf, size_t bytes, int mode, int ctas_per_sm, unsigned* sink, cudaStream_t stream) {
    WB_REQUIRE(buf != nullptr && sink != nullptr && bytes >= (1u << 20) && (reinterpret_cast<uintptr_t>(buf) & 127) == 0, "bad probe buffer");
    int dev = 0, sms = 0;
    WB_CHECK_CUDA(cudaGetDevice(&dev));
    WB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (ctas_per_sm <= 0) ctas_per_sm = mode == 0 ? 4 : 1;
    if (mode == 0) {
        probe_ldg_kernel<<<sms * ctas_per_sm, 256, 0, stream>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, sink);
    } else {
        const size_t smem = PB_STAGES * PB_BYTES + 2 * PB_STAGES * 8;
        static PerDeviceOnce configured;
        configured([&] { WB_CHECK_CUDA(cudaFuncSetAttribute(probe_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); });
        probe_bulk_kernel<<<sms * ctas_per_sm, 128, smem, stream>>>(reinterpret_cast<const uint8_t*>(buf), bytes / PB_BYTES, sink);
    }
    WB_CHECK_LAUNCH();
}

}  // namespace wb

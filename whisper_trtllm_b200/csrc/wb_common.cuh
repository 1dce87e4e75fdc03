// Shared device/host helpers for the sm_100a Whisper kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <atomic>
#include <mutex>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace wb {

// ---------------------------------------------------------------------------------------------
// error plumbing: kernels launchers throw, the C-ABI boundary (abi.cu) catches and returns codes
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define WB_CHECK_CUDA(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            throw ::wb::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) +    \
                                      " (" __FILE__ ":" + std::to_string(__LINE__) + ")");       \
    } while (0)

#define WB_REQUIRE(cond, msg)                                                                    \
    do {                                                                                         \
        if (!(cond))                                                                             \
            throw ::wb::Error(-1, std::string("invalid argument: ") + (msg) + " [" #cond "] (" + \
                                      __FILE__ ":" + std::to_string(__LINE__) + ")");            \
    } while (0)

// every kernel launch of this library goes through WB_CHECK_LAUNCH, which also counts it (bench `gpu_launches`)
inline std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}
#define WB_CHECK_LAUNCH()                                                \
    do {                                                                 \
        ::wb::launch_counter().fetch_add(1, std::memory_order_relaxed);  \
        WB_CHECK_CUDA(cudaGetLastError());                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Per-DEVICE one-time state.  One process normally drives one GPU (one rank per GPU), but nothing in the C-ABI forbids a
// host thread per device in one process: kernel attributes (cudaFuncSetAttribute) and device properties are therefore kept
// per device ordinal, under a mutex (a lock per launch is ~20 ns against ~2 us of launch).
// ---------------------------------------------------------------------------------------------
constexpr int WB_MAX_DEVICES = 64;
inline int current_device() {
    int d = 0;
    WB_CHECK_CUDA(cudaGetDevice(&d));
    WB_REQUIRE(d >= 0 && d < WB_MAX_DEVICES, "device ordinal out of range");
    return d;
}
// runs `fn` once per device (the current one) and `slot`; later calls return immediately
struct PerDeviceOnce {
    std::mutex mu;
    unsigned long long done = 0;
    template <typename F> void operator()(F&& fn) {
        const unsigned long long bit = 1ull << current_device();
        std::lock_guard<std::mutex> lk(mu);
        if (done & bit) return;
        fn();
        done |= bit;
    }
};
// SM count of the current device
inline int device_sm_count() {
    static std::atomic<int> n[WB_MAX_DEVICES];
    const int d = current_device();
    int v = n[d].load(std::memory_order_relaxed);
    if (v == 0) {
        WB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d));
        n[d].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): kernels of the decode step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that the NEXT kernel's launch latency and prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of the current one.  Every such kernel
// calls pdl_wait() before its first global-memory access (reads of the predecessor's outputs AND writes the
// predecessor might still read) and pdl_trigger() right after, which only allows the dependent grid to start
// being scheduled; its own pdl_wait() still blocks until this grid has completed and flushed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Default OFF: measured on B200 (bench.py A/B, profiles/r01_ab_graph_pdl.md) PDL gave no gain on eager launches
// (4.61 s vs 4.58 s per 256 utterances) and cost 15 % inside CUDA-graph replays; the graph alone is the win.
inline std::atomic<bool>& pdl_enabled() {
    static std::atomic<bool> on{false};
    return on;
}

// launch through cudaLaunchKernelEx; `pdl` adds the programmatic-serialization attribute (only for kernels that
// call pdl_wait()).  Counts the launch like WB_CHECK_LAUNCH.
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                          Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled().load(std::memory_order_relaxed)) ? 1 : 0;
    ::wb::launch_counter().fetch_add(1, std::memory_order_relaxed);
    WB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ---------------------------------------------------------------------------------------------
// dtypes
// ---------------------------------------------------------------------------------------------
enum DType : int { F32 = 0, BF16 = 1 };
using bf16 = __nv_bfloat16;

template <typename T> struct DTypeOf;
template <> struct DTypeOf<float> { static constexpr DType value = F32; };
template <> struct DTypeOf<bf16> { static constexpr DType value = BF16; };

inline size_t dtype_size(int dt) { return dt == F32 ? 4 : 2; }

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
    __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec16<bf16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ void unpack(float* f) const {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void pack(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&p);
        }
        raw = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <typename T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
    Vec16<T> v;
    v.raw = *reinterpret_cast<const decltype(v.raw)*>(p);
    return v;
}
// streaming (read-once) 16-byte load: bypass L1 allocation
template <typename T> __device__ __forceinline__ Vec16<T> ld16_stream(const T* p) {
    Vec16<T> v;
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    v.raw = *reinterpret_cast<decltype(v.raw)*>(&r);
    return v;
}
// 16-byte coherent load the compiler may not sink or merge: keeps a whole batch of independent requests in flight
template <typename T> __device__ __forceinline__ Vec16<T> ld16_issue(const T* p) {
    Vec16<T> v;
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    v.raw = *reinterpret_cast<decltype(v.raw)*>(&r);
    return v;
}
template <typename T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& v) {
    *reinterpret_cast<decltype(v.raw)*>(p) = v.raw;
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// exact-erf GELU, the oracle's activation (transformers/activations.py:214 -> nn.functional.gelu)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf-GELU for the bf16 epilogues: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7) on the SFU fast paths
// (rcp.approx, ex2.approx: ~1e-7 relative each), ~17 instructions per element instead of ~34 with the IEEE
// reciprocal / expf (ncu: the fc1 epilogue was issue-bound, 311 M warp instructions vs 100 M for the same GEMM
// without GELU).  Result error <= ~3e-7 * |x|, three orders below bf16 rounding (4e-3 relative).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float ax = fabsf(x);
    const float z = ax * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * z * -1.4426950408889634f));   // exp(-z^2)
    const float h = 0.5f * ax;
    const float e = fmaf(-p * t, ex, 1.0f);      // erf(|x| / sqrt 2)
    return fmaf(h, e, 0.5f * x);                 // 0.5 x (1 + sign(x) erf) = 0.5 x + 0.5 |x| erf(|x| / sqrt 2)
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Decode-loop kernels read the current length and the "still running" flag from device memory so
// that the whole greedy loop can be enqueued (or graph-replayed) without a host round trip.
struct StepState {
    int cur_len;       // tokens already in ids (>= 1); position of the token being fed = cur_len - 1
    int active;        // 1 while the loop runs; 0 once every row finished or max_length was reached
    int final_len;     // length of ids when the loop stopped
    int done_counter;  // last-block detection in the greedy kernel
    int n_unfinished;  // rows not yet at EOS
    int pad[3];
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace wb

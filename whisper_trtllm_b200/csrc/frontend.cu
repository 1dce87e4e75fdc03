// GPU log-mel front-end: 16 kHz PCM, already padded / trimmed to 30 s -> input_features [B, 80, 3000].
// Reference: WhisperFeatureExtractor._np_extract_fbank_features (feature_extraction_whisper.py:96-109) over
// audio_utils.spectrogram (audio_utils.py:379-433), numpy on one CPU core per utterance (run.py:267).
//
//   frames   reflect-pad 200, frame 400 / hop 160, periodic Hann window           frames_kernel  -> [B*3000, 400] fp32
//   DFT      one fp32 GEMM against the [cos | -sin] basis (402 rows, padded 408)   gemm_simt      -> [B*3000, 408]
//   power    re^2 + im^2, 201 bins (padded 208)                                    power_kernel   -> [B*3000, 208]
//   mel      one fp32 GEMM against the 80 slaney filters                           gemm_simt      -> [B*3000, 80]
//   log      log10(max(., 1e-10)), clamp to (utterance max - 8), (x + 4) / 4,      logmel_max_kernel + logmel_finish_kernel
//            transposed to [B, 80, 3000]
// Only the 3000 kept frames are computed (the reference computes 3001 and drops the last one BEFORE taking the max).
// fp32 CUDA-core GEMMs on purpose: the spectrum spans > 8 decades, bf16 tensor-core inputs would not survive the log.
#include "wb_internal.h"

namespace wb {

namespace {
constexpr int N_FFT = 400, HOP = 160, N_SAMPLES = 480000, N_FRAMES = 3000, N_BINS = 201, N_MELS = 80;
constexpr int SPEC_LD = 408, POW_LD = 208;

__global__ void __launch_bounds__(256) frames_kernel(const float* __restrict__ pcm, const float* __restrict__ window,
                                                     float* __restrict__ frames, int frames_per_block) {
    const int b = blockIdx.y, t0 = blockIdx.x * frames_per_block;
    const float* x = pcm + (size_t)b * N_SAMPLES;
    for (int i = threadIdx.x; i < frames_per_block * N_FFT; i += blockDim.x) {
        const int ft = i / N_FFT, j = i - ft * N_FFT;
        const int t = t0 + ft;
        if (t >= N_FRAMES) break;
        int s = t * HOP + j - N_FFT / 2;                 // index into the un-padded waveform
        if (s < 0) s = -s;                               // np.pad(mode="reflect")
        if (s >= N_SAMPLES) s = 2 * N_SAMPLES - 2 - s;
        frames[((size_t)b * N_FRAMES + t) * N_FFT + j] = x[s] * window[j];
    }
}

__global__ void __launch_bounds__(256) power_kernel(const float* __restrict__ spec, float* __restrict__ pw, long long rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * POW_LD) return;
    const long long r = i / POW_LD;
    const int k = (int)(i - r * POW_LD);
    float v = 0.f;
    if (k < N_BINS) {
        const float re = spec[r * SPEC_LD + k], im = spec[r * SPEC_LD + N_BINS + k];
        v = re * re + im * im;
    }
    pw[i] = v;
}

// per-utterance max of log10(max(mel, 1e-10)): block partials, one slot per block
__global__ void __launch_bounds__(256) logmel_max_kernel(const float* __restrict__ mel, float* __restrict__ partial_max, int per_block) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const float* m = mel + (size_t)b * N_FRAMES * N_MELS;
    const int total = N_FRAMES * N_MELS;
    const int i0 = blockIdx.x * per_block;
    float mx = -INFINITY;
    for (int i = i0 + threadIdx.x; i < min(i0 + per_block, total); i += blockDim.x) mx = fmaxf(mx, log10f(fmaxf(m[i], 1e-10f)));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        partial_max[b * gridDim.x + blockIdx.x] = mx;
    }
}

// out[b, mel, t] = (max(log10(max(v, 1e-10)), utterance_max - 8) + 4) / 4, transposed through shared memory
__global__ void __launch_bounds__(256) logmel_finish_kernel(const float* __restrict__ mel, const float* __restrict__ partial_max,
                                                            int n_partials, float* __restrict__ out) {
    __shared__ float tile[32][N_MELS + 1];
    __shared__ float gmax_s;
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    if (threadIdx.x < 32) {
        float mx = -INFINITY;
        for (int i = threadIdx.x; i < n_partials; i += 32) mx = fmaxf(mx, partial_max[b * n_partials + i]);
        mx = warp_max(mx);
        if (threadIdx.x == 0) gmax_s = mx;
    }
    const float* m = mel + ((size_t)b * N_FRAMES + t0) * N_MELS;
    for (int i = threadIdx.x; i < 32 * N_MELS; i += blockDim.x) {
        const int tt = i / N_MELS, k = i - tt * N_MELS;
        tile[tt][k] = (t0 + tt < N_FRAMES) ? log10f(fmaxf(m[i], 1e-10f)) : 0.f;
    }
    __syncthreads();
    const float floor_v = gmax_s - 8.0f;
    for (int i = threadIdx.x; i < 32 * N_MELS; i += blockDim.x) {
        const int k = i >> 5, tt = i & 31;
        if (t0 + tt < N_FRAMES) out[((size_t)b * N_MELS + k) * N_FRAMES + t0 + tt] = (fmaxf(tile[tt][k], floor_v) + 4.0f) * 0.25f;
    }
}
}  // namespace

size_t log_mel_workspace_bytes(int chunk) {
    const size_t rows = (size_t)chunk * N_FRAMES;
    return rows * (N_FFT + SPEC_LD + POW_LD + N_MELS) * 4 + (size_t)chunk * 64 * 4 + 4096;
}

void log_mel(const float* pcm, int B, const float* window, const float* dft_basis, const float* mel_filters, void* workspace,
             size_t workspace_bytes, float* out, cudaStream_t st) {
    WB_REQUIRE(pcm && window && dft_basis && mel_filters && workspace && out && B > 0, "bad log-mel arguments");
    // largest chunk of utterances the workspace can hold
    int chunk = std::min(B, 64);
    while (chunk > 1 && log_mel_workspace_bytes(chunk) > workspace_bytes) chunk >>= 1;
    WB_REQUIRE(log_mel_workspace_bytes(chunk) <= workspace_bytes, "log-mel workspace too small (see wb_log_mel_workspace_bytes)");
    uint8_t* base = (uint8_t*)workspace;
    const size_t skew = (256 - (reinterpret_cast<uintptr_t>(base) & 255)) & 255;
    float* frames = reinterpret_cast<float*>(base + skew);
    const size_t rows_max = (size_t)chunk * N_FRAMES;
    float* spec = frames + rows_max * N_FFT;
    float* pw = spec + rows_max * SPEC_LD;
    float* mel = pw + rows_max * POW_LD;
    float* pmax = mel + rows_max * N_MELS;
    constexpr int FPB = 8, MAX_PER_BLOCK = N_FRAMES * N_MELS / 60, N_PART = 60;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int bc = std::min(chunk, B - b0);
        const long long rows = (long long)bc * N_FRAMES;
        frames_kernel<<<dim3(ceil_div(N_FRAMES, FPB), bc), 256, 0, st>>>(pcm + (size_t)b0 * N_SAMPLES, window, frames, FPB);
        WB_CHECK_LAUNCH();
        GemmArgs g1;
        g1.A = frames; g1.lda = N_FFT; g1.W = dft_basis; g1.ldw = N_FFT; g1.in_dtype = F32;
        g1.out = spec; g1.ldo = SPEC_LD; g1.out_dtype = F32; g1.M = (int)rows; g1.N = SPEC_LD; g1.K = N_FFT;
        gemm_simt(g1, st);
        power_kernel<<<(unsigned)((rows * POW_LD + 255) / 256), 256, 0, st>>>(spec, pw, rows);
        WB_CHECK_LAUNCH();
        GemmArgs g2;
        g2.A = pw; g2.lda = POW_LD; g2.W = mel_filters; g2.ldw = POW_LD; g2.in_dtype = F32;
        g2.out = mel; g2.ldo = N_MELS; g2.out_dtype = F32; g2.M = (int)rows; g2.N = N_MELS; g2.K = POW_LD;
        gemm_simt(g2, st);
        logmel_max_kernel<<<dim3(N_PART, bc), 256, 0, st>>>(mel, pmax, MAX_PER_BLOCK);
        WB_CHECK_LAUNCH();
        logmel_finish_kernel<<<dim3(ceil_div(N_FRAMES, 32), bc), 256, 0, st>>>(mel, pmax, N_PART, out + (size_t)b0 * N_MELS * N_FRAMES);
        WB_CHECK_LAUNCH();
    }
}

}  // namespace wb
